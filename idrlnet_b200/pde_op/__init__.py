"""Equations and operators (package layout of idrlnet/pde_op)."""
from .equations import *  # noqa: F401,F403
from .operator import *  # noqa: F401,F403
