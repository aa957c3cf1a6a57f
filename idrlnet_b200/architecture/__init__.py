"""Network architectures (package layout of idrlnet/architecture: `layer`, `mlp`)."""
from .core import *  # noqa: F401,F403
from .core import __all__  # noqa: F401
