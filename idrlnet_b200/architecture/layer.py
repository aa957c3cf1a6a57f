"""Layer helpers under the reference's module path (idrlnet/architecture/layer.py)."""
from .core import Activation, Initializer, get_activation_layer, get_linear_layer  # noqa: F401

__all__ = ["Activation", "Initializer", "get_activation_layer", "get_linear_layer"]
