"""SymPy <-> torch glue (mirror of idrlnet/torch_util.py).

Defines the residual language: a SymPy expression whose derivatives are renamed
to plain symbols ("u__x__x") and which is either lambdified onto torch (generic
path) or compiled to a polynomial for the fused kernel (`compiler.py`)."""
from functools import reduce

import sympy
import torch
from sympy import Derivative, Function, Symbol, lambdify
from sympy.printing.str import StrPrinter
from sympy.utilities.lambdify import implemented_function

from .header import DIFF_SYMBOL

__all__ = ["integral", "torch_lambdify"]


def _integral_fun(x):
    if isinstance(x, torch.Tensor):
        return torch.sum(x, dim=0, keepdim=True) * torch.ones_like(x)
    return x


integral = implemented_function("integral", lambda x: _integral_fun(x))  # idrlnet/torch_util.py:16-22

def _tmax(a, b):
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        return torch.maximum(a, b)
    if isinstance(a, torch.Tensor):
        return torch.clamp(a, min=float(b))
    if isinstance(b, torch.Tensor):
        return torch.clamp(b, min=float(a))
    return max(a, b)


def _tmin(a, b):
    if isinstance(a, torch.Tensor) and isinstance(b, torch.Tensor):
        return torch.minimum(a, b)
    if isinstance(a, torch.Tensor):
        return torch.clamp(a, max=float(b))
    if isinstance(b, torch.Tensor):
        return torch.clamp(b, max=float(a))
    return min(a, b)


TORCH_SYMPY_PRINTER = {  # idrlnet/torch_util.py:43-65 (Min/Max also take python scalars here)
    "sin": torch.sin, "cos": torch.cos, "tan": torch.tan, "exp": torch.exp, "sqrt": torch.sqrt,
    "Abs": torch.abs, "tanh": torch.tanh, "DiracDelta": torch.zeros_like,
    "Heaviside": lambda x, *half: torch.heaviside(x, torch.tensor([0.0], device=x.device, dtype=x.dtype)),
    "amin": lambda x: reduce(_tmin, x), "amax": lambda x: reduce(_tmax, x),
    "Min": lambda *x: reduce(_tmin, x), "Max": lambda *x: reduce(_tmax, x), "sign": torch.sign,
    "equal": lambda x, y: torch.isclose(x, y), "Xor": torch.logical_xor, "log": torch.log,
    "sinh": torch.sinh, "cosh": torch.cosh, "asin": torch.arcsin, "acos": torch.arccos, "atan": torch.arctan,
}


def torch_lambdify(r, f, *args, **kwargs):
    """Callable(**tensors) evaluating `f` over the symbols named in `r`."""
    try:
        f = float(f)
    except (TypeError, ValueError):
        pass
    if isinstance(f, (float, int, bool)):
        const = f
        return lambda **x: torch.zeros_like(next(iter(x.values()))) + const
    return lambdify([k for k in r], f, [TORCH_SYMPY_PRINTER], *args, **kwargs)


def _replace_derivatives(expr):
    """Derivative(u(x,t), x, x) -> Symbol('u__x__x'); u(x,t) -> Symbol('u')
    (idrlnet/torch_util.py:72-88)."""
    expr = sympy.sympify(expr)
    while True:
        ds = expr.atoms(Derivative)
        if not ds:
            break
        d = next(iter(ds))
        expr = expr.subs(d, Function(str(d))(*d.free_symbols))
    while True:
        funs = [f for f in expr.atoms(Function)
                if f.class_key()[1] == 0 and f.class_key()[2] != "integral"]
        if not funs:
            break
        f = funs[0]
        expr = expr.subs(f, Symbol(str(f)))
    return expr


class UnderlineDerivativePrinter(StrPrinter):
    def _print_Function(self, expr):
        return expr.func.__name__

    def _print_Derivative(self, expr):
        return "".join([str(expr.args[0].func)] + [order * (DIFF_SYMBOL + str(key)) for key, order in expr.args[1:]])


def sstr(expr, **settings):
    return UnderlineDerivativePrinter(settings).doprint(expr)


# The reference patches SymPy's str() globally so that str(Derivative(u, x, x)) ==
# "u__x__x" and str(u(x, t)) == "u" (idrlnet/torch_util.py:108); user code such as
# f"burgers_{str(u)}" depends on it, so the drop-in does the same.
sympy.Basic.__str__ = lambda self: sstr(self, order=None)
