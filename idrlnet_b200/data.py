"""DataNode / SampleDomain (mirror of idrlnet/data.py): sampling domains that
supply the [N,1] FP32 columns (coordinates, normals, sdf, area) and targets."""
import abc
import functools
import inspect
from typing import Callable, List, Tuple

import numpy
import sympy
import torch

from .header import logger
from .node import Node
from .variable import Variables

__all__ = ["DataNode", "get_data_node", "get_data_nodes", "datanode", "SampleDomain"]


def _lambdify_np(expr, names):
    """Evaluate a target on the sampled numpy columns exactly as idrlnet/data.py:82-92 does: through
    `geo_utils.sympy_np.lambdify_np`, which takes numbers, bools, python callables `f(**columns)` and SymPy
    expressions (compiled once and cached); the result is broadcast to the column shape."""
    from .geo_utils.sympy_np import lambdify_np
    fn = lambdify_np(expr, names)

    def call(**kw):
        like = next(iter(kw.values()))
        out = numpy.asarray(fn(**kw))
        if out.dtype == bool:
            out = out.astype(numpy.float64)
        return out if out.shape == numpy.shape(like) else numpy.broadcast_to(out, numpy.shape(like)).copy()
    return call


def _target_on_device(expr, input_vars, like: torch.Tensor) -> torch.Tensor:
    try:  # a 0-d CPU tensor: broadcasts against device tensors and converts to float without a device sync
        return torch.tensor(float(expr), dtype=torch.float32)
    except (TypeError, ValueError):
        pass
    from .torch_util import torch_lambdify
    e = sympy.sympify(expr)
    used = [s.name for s in e.free_symbols]
    cols = {n: (input_vars[n].detach() if isinstance(input_vars[n], torch.Tensor)
                else torch.as_tensor(numpy.asarray(input_vars[n]), dtype=torch.float32, device=like.device)) for n in used}
    out = torch_lambdify(used, e)(**cols)
    return out if isinstance(out, torch.Tensor) else torch.full((1, 1), float(out), dtype=torch.float32, device=like.device)


class DataNode(Node):
    counter = 0

    def __init__(self, inputs, outputs, sample_fn: Callable, loss_fn: str = "square", lambda_outputs=None, name=None,
                 sigma=1.0, var_sigma=False, *args, **kwargs):
        self.inputs = inputs
        self.outputs = outputs
        self.lambda_outputs = lambda_outputs
        if name is not None:
            self.name = name
        else:
            self.name = "Domain_{}".format(self.counter)
            type(self).counter += 1
        self.sigma = torch.tensor(sigma, dtype=torch.float32, requires_grad=var_sigma)
        self.sample_fn = sample_fn
        self.loss_fn = loss_fn

    def sample(self) -> Variables:
        """Call the user's sampler; evaluate symbolic / constant targets on the
        points; convert to tensors (idrlnet/data.py:75-97)."""
        input_vars, output_vars = self.sample_fn()
        output_vars = dict(output_vars)
        on_device = next((v for v in input_vars.values() if isinstance(v, torch.Tensor)), None)
        np_in = None
        for key, value in output_vars.items():
            if isinstance(value, (torch.Tensor, numpy.ndarray)):
                continue
            try:
                if on_device is not None:
                    # points already live in tensors (e.g. drawn on the GPU): evaluate the target there instead of
                    # copying every column to the host; a constant stays a 0-d scalar
                    output_vars[key] = _target_on_device(value, input_vars, on_device)
                    continue
                if np_in is None:
                    np_in = dict(input_vars)
                output_vars[key] = _lambdify_np(value, list(np_in))(**np_in)
            except Exception:
                logger.error("unsupported constraints type.")
                raise ValueError("unsupported constraints type.")
        return Variables({**input_vars, **output_vars}).to_torch_tensor_()

    def __str__(self):
        return super().__str__() + "DataNode properties:\nlambda_outputs: {}\n".format(self.lambda_outputs)


def get_data_node(fun: Callable, name=None, loss_fn="square", sigma=1.0, var_sigma=False, *args, **kwargs) -> DataNode:
    """Calls `fun()` once to discover the keys (idrlnet/data.py:132-176)."""
    in_, out_ = fun()
    outputs = list(out_.keys())
    lambda_outputs = [k for k in outputs if k.startswith("lambda_")]
    outputs = [k for k in outputs if not k.startswith("lambda_")]
    if name is None:
        name = fun.__name__ if inspect.isfunction(fun) else type(fun).__name__
    return DataNode(inputs=list(in_.keys()), outputs=outputs, sample_fn=fun, lambda_outputs=lambda_outputs,
                    loss_fn=loss_fn, name=name, sigma=sigma, var_sigma=var_sigma, *args, **kwargs)


def datanode(_fun: Callable = None, name=None, loss_fn="square", sigma=1.0, var_sigma=False, **kwargs):
    """Decorator form (idrlnet/data.py:179-211): classes are instantiated at once;
    the decorated name becomes a zero-argument DataNode factory."""

    def wrap(fun):
        if inspect.isclass(fun):
            assert issubclass(fun, SampleDomain), f"{fun} should be subclass of .data.Sample"
            fun = fun()
        assert callable(fun)

        @functools.wraps(fun)
        def wrapped_fun():
            return get_data_node(fun, name=name, loss_fn=loss_fn, sigma=sigma, var_sigma=var_sigma, **kwargs)

        return wrapped_fun

    return wrap if _fun is None else wrap(_fun)


def get_data_nodes(funs: List[Callable], *args, **kwargs) -> Tuple[DataNode]:
    if "names" in kwargs:
        names = kwargs.pop("names")
        return tuple(get_data_node(f, name=n, *args, **kwargs) for f, n in zip(funs, names))
    return tuple(get_data_node(f, *args, **kwargs) for f in funs)


class SampleDomain(metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def sampling(self, *args, **kwargs):
        raise NotImplementedError(f"{type(self)}.sampling method not implemented")

    def __call__(self, *args, **kwargs):
        return self.sampling(self, *args, **kwargs)  # (sic) idrlnet/data.py:233-234
