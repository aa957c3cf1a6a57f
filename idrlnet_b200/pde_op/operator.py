"""Pointwise and integral operators (mirror of idrlnet/pde_op/operator.py)."""
from typing import List, Union

import numpy as np
import sympy as sp
import torch
from sympy import Function, Number, Symbol, symbols

from ..node import Node
from ..pde import PdeNode
from ..torch_util import _replace_derivatives, integral, torch_lambdify
from ..variable import Variables
from ._common import _X, _Y, _Z, _coords

__all__ = ["NormalGradient", "Difference", "Derivative", "Curl", "Divergence", "ICNode", "Int1DNode", "IntEq"]


class NormalGradient(PdeNode):
    """n . grad T  (idrlnet/pde_op/operator.py:26-54)."""

    def __init__(self, T: Union[str, Symbol, float, int], dim=3, time=True):
        super().__init__()
        self.T, self.dim, self.time = T, dim, time
        nx, ny, nz = symbols("normal_x normal_y normal_z")
        Tf = Function(T)(*_coords(dim, time))
        self.equations = {"normal_gradient_" + T: nx * Tf.diff(_X) + ny * Tf.diff(_Y) + nz * Tf.diff(_Z)}
        self.make_nodes()


class Difference(PdeNode):
    """T - S  (idrlnet/pde_op/operator.py:57-88)."""

    def __init__(self, T, S, dim=3, time=True):
        super().__init__()
        self.T, self.S, self.dim, self.time = T, S, dim, time
        variables = _coords(dim, time)
        self.equations = {"difference_" + T + "_" + S: Function(T)(*variables) - Function(S)(*variables)}
        self.make_nodes()


class Derivative(PdeNode):
    """dT/dp - S  (idrlnet/pde_op/operator.py:91-130)."""

    def __init__(self, T, p, S=0.0, dim=3, time=True):
        super().__init__()
        self.T, self.S, self.dim, self.time = T, S, dim, time
        variables = _coords(dim, time)
        Sf = Function(S)(*variables) if isinstance(S, str) else (Number(S) if isinstance(S, (float, int)) else S)
        p = Symbol(p) if isinstance(p, str) else p
        name = "derivative_" + T + ":" + str(p) + ("_" + str(S) if isinstance(S, str) else "")
        self.equations = {name: Function(T)(*variables).diff(p) - Sf}
        self.make_nodes()


def _component(v, variables):
    return Function(v)(*variables) if isinstance(v, str) else (Number(v) if isinstance(v, (int, float)) else v)


class Curl(PdeNode):
    """curl of (v0, v1, v2) in x, y, z (idrlnet/pde_op/operator.py:133-164); like the reference
    it only fills `equations` -- call `make_nodes()` to use it as a node."""

    def __init__(self, vector, curl_name=None):
        super().__init__()
        names = ["u", "v", "w"] if curl_name is None else curl_name
        a, b, c = (_component(v, (_X, _Y, _Z)) for v in vector)
        self.equations = {names[0]: c.diff(_Y) - b.diff(_Z), names[1]: a.diff(_Z) - c.diff(_X),
                          names[2]: b.diff(_X) - a.diff(_Y)}


class Divergence(PdeNode):
    """Sum of the three components (idrlnet/pde_op/operator.py:167-192; the reference does not
    differentiate them -- pass derivative expressions)."""

    def __init__(self, vector, div_name="div_v"):
        super().__init__()
        a, b, c = (_component(v, (_X, _Y, _Z)) for v in vector)
        self.equations = {div_name: a + b + c}


class ICNode(PdeNode):
    """Monte-Carlo integral over the sampled domain: integral(T * area), or of the flux
    n . (T0, T1[, T2]) for a list (idrlnet/pde_op/operator.py:195-253).  `integral` sums over the
    points of the batch (idrlnet/torch_util.py:16-22)."""

    def __init__(self, T, dim: int = 2, time: bool = False, reduce_name: str = None):
        super().__init__()
        self.T, self.dim, self.time = T, dim, time
        self.reduce_name = str(T) if reduce_name is None else reduce_name
        variables = _coords(dim, time)
        area = Symbol("area")
        normals = symbols("normal_x normal_y normal_z")
        if isinstance(T, list):
            comps = [_component(t, variables) for t in T]
            assert dim in (2, 3), "a flux integral needs dim 2 or 3"
            integrand = sum(n * c for n, c in zip(normals[:dim], comps)) * area
        else:
            integrand = _component(T, variables) * area
        self.equations = {"integral_" + self.reduce_name: integral(integrand)}
        self.make_nodes()


class IntEq:
    """Evaluator of one Int1DNode equation: Gauss-Legendre quadrature in `var` between the
    per-point bounds, nets named in `funs` evaluated at the quadrature nodes."""

    def __init__(self, binding_node, lb_fn, ub_fn, outer_symbols, free_symbols, eq_fn, name):
        self.binding_node = binding_node
        self.lb_fn, self.ub_fn, self.eq_fn = lb_fn, ub_fn, eq_fn
        self.outer_symbols, self.free_symbols, self.name = outer_symbols, free_symbols, name

    def __call__(self, var: Variables):
        node = self.binding_node
        var = dict(var)
        outer = {k: v for k, v in var.items() if k in self.outer_symbols}
        lb, ub = self.lb_fn(**outer), self.ub_fn(**outer)
        like = next(iter(outer.values())) if outer else next(iter(var.values()))
        xi, w = node.quad_s.to(like.device), node.quad_w.to(like.device)
        half = (ub - lb) / 2 + torch.zeros_like(like)
        weights = (half * w).reshape(-1, 1)                      # [N*Q, 1]
        nodes = ((xi + 1) * half + lb).reshape(-1, 1)
        n_pts, q = half.shape[0], xi.numel()
        cols = {k: (torch.ones_like(xi) * var[k]).reshape(-1, 1) for k in self.free_symbols if k in var}
        for fun in node.funs.values():
            feed = dict(cols)
            feed.update({k: nodes for k in fun["input_map"]})
            res = dict(fun["eval"].evaluate(feed))
            cols.update({alias: res[out] for out, alias in fun["output_map"].items()})
        values = weights * self.eq_fn(**{node.var.name: nodes}, **cols)
        return {self.name: values.reshape(n_pts, q).sum(1, keepdim=True)}


class Int1DNode(PdeNode):
    """expression_name(x) = int_{lb}^{ub} expression(x, var) d var  (idrlnet/pde_op/operator.py:256-360).
    `funs = {symbol: {"eval": node, "input_map": {node_input: var}, "output_map": {node_output: symbol}}}`
    names nets to evaluate inside the integrand."""
    counter = 0

    def __init__(self, expression, expression_name, lb, ub, var: Union[str, sp.Symbol] = "s", degree=20, **kwargs):
        self.funs = kwargs.pop("funs", {})
        super().__init__(**kwargs)
        x = sp.Symbol("x")
        self.var = sp.Symbol(var) if isinstance(var, str) else var
        self.degree = degree
        xi, w = np.polynomial.legendre.leggauss(degree)
        self.quad_s = torch.tensor(xi, dtype=torch.float32)
        self.quad_w = torch.tensor(w, dtype=torch.float32)

        def bound(b):
            return sp.Function(b)(x) if isinstance(b, str) else (sp.Number(b) if isinstance(b, (float, int)) else b)
        self.lb, self.ub = bound(lb), bound(ub)
        if isinstance(expression, (float, int)):
            expression = sp.Number(expression)
        if not isinstance(expression, sp.Expr):
            raise TypeError("expression must be a number or a SymPy expression")
        self.equations = {expression_name: expression}
        self.fun_require_input = set()
        for fun in self.funs.values():
            self.fun_require_input |= set(fun["eval"].inputs) - set(fun["input_map"].keys())
        self.make_nodes()

    def make_nodes(self) -> None:
        self.sub_nodes = []
        free, names = set(), set()
        self.lb, self.ub = _replace_derivatives(self.lb), _replace_derivatives(self.ub)
        for name, eq in self.equations.items():
            eq = _replace_derivatives(eq)
            for e in (self.lb, self.ub, eq):
                free.update(s.name for s in e.free_symbols)
            free.update(self.fun_require_input)
            free.discard(self.var.name)
            name = name + self.suffix
            self.sub_nodes.append(self.new_node(name, eq, list(free)))
            names.add(name)
        self.inputs = [s for s in free if s not in self.funs]
        self.derivatives = []
        self.outputs = list(names)

    def new_node(self, name: str = None, tf_eq: sp.Expr = None, free_symbols: List[str] = None, *args, **kwargs):
        outer = [s for s in free_symbols if s not in self.funs]
        node = Node()
        node.evaluate = IntEq(self, torch_lambdify(outer, self.lb), torch_lambdify(outer, self.ub), outer, free_symbols,
                              torch_lambdify([*free_symbols, self.var.name], tf_eq), name)
        node.inputs, node.derivatives, node.outputs, node.name = outer, [], [name], name
        return node
