"""Geometry and sampling (package layout of idrlnet/geo_utils)."""
from .geo import *  # noqa: F401,F403
from .geo_builder import GeometryBuilder  # noqa: F401
from .geo_obj import *  # noqa: F401,F403
from .sympy_np import *  # noqa: F401,F403
