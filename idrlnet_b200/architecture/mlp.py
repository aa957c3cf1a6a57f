"""Architectures under the reference's module path (idrlnet/architecture/mlp.py)."""
from .core import MLP, Siren, Arch, get_net_node, get_shared_net_node, SingleVar, BoundedSingleVar  # noqa: F401

__all__ = ['MLP', 'Siren', 'Arch', 'get_net_node', 'get_shared_net_node', 'SingleVar', 'BoundedSingleVar']
