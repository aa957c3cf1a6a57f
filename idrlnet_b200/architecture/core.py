"""Network architectures -- host-side mirror of idrlnet/architecture/{layer,mlp}.py.

Same names, arguments and state-dict keys as the reference (so `model.ckpt`
files stay interchangeable: `layers.mlp_0.weight_g`, ...), written from scratch.
The modules are ordinary torch modules; the fused CUDA step reads their
parameters through `idrlnet_b200.fused.NetPack`.
"""
from __future__ import annotations

import enum
import math
from collections import OrderedDict
from typing import List, Sequence, Union

import torch

from ..header import logger
from ..net import NetNode

__all__ = ["Activation", "Initializer", "get_activation_layer", "get_linear_layer", "MLP", "Siren", "Arch",
           "get_net_node", "get_shared_net_node", "SingleVar", "BoundedSingleVar"]


class Activation(enum.Enum):  # idrlnet/architecture/layer.py:11-20
    relu = "relu"
    silu = "silu"
    selu = "selu"
    sigmoid = "sigmoid"
    tanh = "tanh"
    swish = "swish"
    poly = "poly"
    sin = "sin"
    leaky_relu = "leaky_relu"


class Initializer(enum.Enum):  # idrlnet/architecture/layer.py:23-27
    Xavier_uniform = "Xavier_uniform"
    constant = "constant"
    kaiming_uniform = "kaiming_uniform"
    default = "default"


def _initializer(kind: Initializer, **kwargs):
    """idrlnet/architecture/layer.py:146-164."""
    if kind == Initializer.Xavier_uniform:
        return torch.nn.init.xavier_uniform_
    if kind == Initializer.constant:
        return lambda w: torch.nn.init.constant_(w, kwargs["constant"])
    if kind == Initializer.kaiming_uniform:
        return lambda w: torch.nn.init.kaiming_uniform_(w, mode="fan_in", nonlinearity="relu")
    if kind == Initializer.default:
        return lambda w: w
    raise NotImplementedError(f"initialization {kind} is not supported")


def get_linear_layer(input_dim: int, output_dim: int, weight_norm: bool = False,
                     initializer: Initializer = Initializer.Xavier_uniform, *args, **kwargs) -> torch.nn.Module:
    """Linear layer with the reference's init: chosen initializer on the weight,
    zero bias, optional weight normalisation (idrlnet/architecture/layer.py:30-44)."""
    layer = torch.nn.Linear(input_dim, output_dim)
    _initializer(initializer, **kwargs)(layer.weight)
    torch.nn.init.zeros_(layer.bias)
    if weight_norm:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            layer = torch.nn.utils.weight_norm(layer)  # params: bias, weight_g, weight_v (ckpt-compatible)
    return layer


class _ActivationModule(torch.nn.Module):
    fn = None

    def forward(self, x):
        return type(self).fn(x)


def _make_activation(name: str, fn) -> torch.nn.Module:
    # the class is named after the activation, like the reference's `modularize`
    # (idrlnet/architecture/layer.py:52-65); fused.NetPack recognises it by name.
    return type(name, (_ActivationModule,), {"fn": staticmethod(fn)})()


def _swish(x):
    return x * torch.sigmoid(x)


def _poly(x):
    return torch.cat([x ** 3, x ** 2, x], dim=-1)


def get_activation_layer(activation: Activation = Activation.swish, *args, **kwargs) -> torch.nn.Module:
    """idrlnet/architecture/layer.py:47-92."""
    table = {
        Activation.relu: torch.relu, Activation.selu: torch.selu, Activation.sigmoid: torch.sigmoid,
        Activation.tanh: torch.tanh, Activation.swish: _swish, Activation.silu: torch.nn.functional.silu,
        Activation.sin: torch.sin, Activation.poly: _poly,
    }
    if activation not in table:
        raise NotImplementedError(f"Activation {activation} is not supported")
    return _make_activation(activation.name, table[activation])


class MLP(torch.nn.Module):
    """Multi-layer perceptron (idrlnet/architecture/mlp.py:18-76): Linear
    (+ weight-norm, on by default) and activation after all but the last layer."""

    def __init__(self, n_seq: List[int], activation: Union[Activation, List[Activation]] = Activation.swish,
                 initialization: Initializer = Initializer.kaiming_uniform, weight_norm: bool = True,
                 name: str = "mlp", *args, **kwargs):
        super().__init__()
        assert isinstance(n_seq, list)
        layers = OrderedDict()
        n_lin = len(n_seq) - 1
        for i in range(n_lin):
            layers[f"{name}_{i}"] = get_linear_layer(n_seq[i], n_seq[i + 1], weight_norm, initialization, *args, **kwargs)
            if isinstance(activation, list):
                act = activation[i]
            elif i < n_lin - 1:
                act = activation
            else:
                act = None
            if act is not None and act != "none":
                layers[f"{name}_{i}_activation"] = get_activation_layer(act, *args, **kwargs)
        self.layers = torch.nn.ModuleDict(layers)

    def forward(self, x):
        for layer in self.layers.values():
            x = layer(x)
        return x


class Siren(torch.nn.Module):
    """Sinusoidal representation network (idrlnet/architecture/mlp.py:79-137).
    Runs on the generic torch path (the omega scaling is not a plain Linear stack)."""

    def __init__(self, n_seq: List[int], first_omega: float = 30.0, omega: float = 30.0, name: str = "siren",
                 *args, **kwargs):
        super().__init__()
        self.first_omega, self.omega = first_omega, omega
        layers = OrderedDict()
        for i in range(len(n_seq) - 1):
            lin = torch.nn.Linear(n_seq[i], n_seq[i + 1])
            dim = n_seq[i]
            bound = 1.0 / dim if i == 0 else math.sqrt(6.0 / dim) / omega
            torch.nn.init.uniform_(lin.weight, -bound, bound)
            torch.nn.init.uniform_(lin.bias, -math.sqrt(1 / dim), math.sqrt(1 / dim))
            layers[f"{name}_{i}"] = lin
            if i < len(n_seq) - 2:
                layers[f"{name}_{i}_activation"] = get_activation_layer(Activation.sin)
        self.layers = torch.nn.ModuleDict(layers)

    def forward(self, x):
        n = len(self.layers)
        for i, layer in enumerate(self.layers.values()):
            x = layer(x)
            if isinstance(layer, torch.nn.Linear) and i < n - 1:
                x = (self.first_omega if i == 0 else self.omega) * x
        return x


class Arch(enum.Enum):  # idrlnet/architecture/mlp.py:180-188
    mlp = "mlp"
    toy = "toy"
    mlp_xl = "mlp_xl"
    single_var = "single_var"
    bounded_single_var = "bounded_single_var"
    siren = "siren"


class SingleVar(torch.nn.Module):
    """Unknown scalar coefficient of an inverse problem (idrlnet/architecture/mlp.py:140-155)."""

    def __init__(self, initialization: float = 1.0):
        super().__init__()
        self.value = torch.nn.Parameter(torch.tensor([float(initialization)]))

    def forward(self, x):
        return x[:, :1] * 0.0 + self.value

    def get_value(self):
        return self.value


class BoundedSingleVar(torch.nn.Module):
    """idrlnet/architecture/mlp.py:158-177."""

    def __init__(self, lower_bound, upper_bound):
        super().__init__()
        self.value = torch.nn.Parameter(torch.tensor([0.0]))
        self.ub, self.lb = upper_bound, lower_bound

    def get_value(self):
        return torch.sigmoid(self.value) * (self.ub - self.lb) + self.lb

    def forward(self, x):
        return x[:, :1] * 0.0 + self.get_value()


def get_net_node(inputs: Sequence[str], outputs: Sequence[str], arch: Arch = None, name=None, *args, **kwargs) -> NetNode:
    """NetNode around a pre-configured network (idrlnet/architecture/mlp.py:191-266):
    Arch.mlp = [d,20,20,20,20,o] swish + weight-norm + kaiming;
    Arch.mlp_xl = [d,512x6,o] SiLU + weight-norm + kaiming."""
    arch = Arch.mlp if arch is None else arch
    if "evaluate" in kwargs:
        net = kwargs.pop("evaluate")
    elif arch == Arch.mlp:
        seq = kwargs.get("seq", [len(inputs), 20, 20, 20, 20, len(outputs)])
        net = MLP(seq, Activation.swish, Initializer.kaiming_uniform, weight_norm=True)
    elif arch == Arch.mlp_xl or arch == "fc":
        seq = kwargs.get("seq", [len(inputs)] + [512] * 6 + [len(outputs)])
        net = MLP(seq, Activation.silu, Initializer.kaiming_uniform, weight_norm=True)
    elif arch == Arch.single_var:
        net = SingleVar(initialization=kwargs.get("initialization", 1.0))
    elif arch == Arch.bounded_single_var:
        net = BoundedSingleVar(lower_bound=kwargs["lower_bound"], upper_bound=kwargs["upper_bound"])
    elif arch == Arch.siren:
        seq = kwargs.get("seq", [len(inputs)] + [512] * 6 + [len(outputs)])
        net = Siren(seq)
    else:
        logger.error(f"{arch} is not supported!")
        raise NotImplementedError(f"{arch} is not supported!")
    return NetNode(inputs=inputs, outputs=outputs, net=net, name=name, *args, **kwargs)


def get_shared_net_node(shared_node: NetNode, inputs: Sequence[str], outputs: Sequence[str], name=None,
                        *args, **kwargs) -> NetNode:
    """A second NetNode evaluating the SAME network on other inputs
    (idrlnet/architecture/mlp.py:269-296); its parameters are optimised once."""
    return NetNode(inputs, outputs, shared_node.net, is_reference=True, name=name, *args, **kwargs)
