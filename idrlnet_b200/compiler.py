"""Jet planner and residual compiler.

Turns what the reference evaluates per domain with a node graph
(idrlnet/graph.py:95-121,170-193: longest-prefix matching, autograd for the
leftover derivative symbols; idrlnet/pde.py:65-79 + idrlnet/torch_util.py:25-39:
lambdified SymPy residuals) into the descriptors of include/pinn_b200.h:

* the set of input-derivative channels the net must propagate (`JetPlan`,
  SURVEY.md Appendix C), and
* each loss term as a polynomial over symbol ids (`pinn_term`).

Anything the fused kernels cannot express raises `NotFusable`; the caller then
takes the generic path (jets from the kernel, residual in torch) or plain torch.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Mapping, Sequence, Tuple, Union

import sympy

from . import native as nv

DIFF = "__"  # idrlnet/header.py:6


class NotFusable(Exception):
    pass


def split_name(name: str) -> Tuple[str, Tuple[str, ...]]:
    parts = name.split(DIFF)
    return parts[0], tuple(parts[1:])


@dataclass
class NetSpec:
    name: str
    inputs: Tuple[str, ...]
    outputs: Tuple[str, ...]


@dataclass
class JetPlan:
    """first_in[i] = net-input index of first-order channel i; the first
    n_second of them also carry a pure second-order channel.  Extra channels
    (narrow nets): n_mixed = 1 adds d2/d(first_in[0]) d(first_in[1]); n_high = 1 / 2
    adds the pure third (and fourth) derivative along first_in[0]."""
    first_in: Tuple[int, ...] = ()
    n_second: int = 0
    n_mixed: int = 0
    n_high: int = 0

    @property
    def n_channels(self) -> int:
        return 1 + len(self.first_in) + self.n_second + self.n_mixed + self.n_high

    def channel(self, inputs: Sequence[str], derivs: Tuple[str, ...]) -> int:
        nf = len(self.first_in)
        if len(derivs) == 0:
            return 0
        if len(derivs) == 1:
            return 1 + self.first_in.index(inputs.index(derivs[0]))
        if len(set(derivs)) == 1:
            s = self.first_in.index(inputs.index(derivs[0]))
            if len(derivs) == 2:
                if s >= self.n_second:
                    raise KeyError(derivs)
                return 1 + nf + s
            if s == 0 and len(derivs) - 2 <= self.n_high:
                return 1 + nf + self.n_second + self.n_mixed + (len(derivs) - 3)
        elif len(derivs) == 2 and self.n_mixed and nf >= 2:
            if {inputs.index(v) for v in derivs} == {self.first_in[0], self.first_in[1]}:
                return 1 + nf + self.n_second
        raise NotFusable(f"derivative {DIFF.join(derivs)} has no fused channel")

    def channel_names(self, inputs: Sequence[str]) -> List[Tuple[str, ...]]:
        out: List[Tuple[str, ...]] = [()]
        out += [(inputs[d],) for d in self.first_in]
        out += [(inputs[d], inputs[d]) for d in self.first_in[: self.n_second]]
        if self.n_mixed:
            out.append((inputs[self.first_in[0]], inputs[self.first_in[1]]))
        out += [(inputs[self.first_in[0]],) * (3 + k) for k in range(self.n_high)]
        return out


def plan_jets(inputs: Sequence[str], derivative_tuples: Sequence[Tuple[str, ...]]) -> JetPlan:
    """Channels for a set of derivative symbols.  Beyond first / pure second order the kernels carry ONE mixed second
    derivative (of exactly two differentiated inputs) or the pure third / fourth derivative along ONE input
    (idrlnet_b200/csrc/narrow_warp.h, NarrowWarp's X); anything else is NotFusable."""
    first, second = [], []
    mixed, high_var, n_high = None, None, 0
    for dt in derivative_tuples:
        if len(dt) == 0:
            continue
        for v in dt:
            if v not in inputs:
                raise NotFusable(f"derivative w.r.t. {v}, which is not an input of the net")
        if len(set(dt)) == 1:
            if len(dt) > 4:
                raise NotFusable(f"derivative {DIFF.join(dt)} (order > 4) has no fused channel")
            if dt[0] not in first:
                first.append(dt[0])
            if len(dt) >= 2 and dt[0] not in second:
                second.append(dt[0])
            if len(dt) >= 3:
                if high_var not in (None, dt[0]):
                    raise NotFusable("third/fourth-order derivatives along more than one input")
                high_var, n_high = dt[0], max(n_high, len(dt) - 2)
        elif len(dt) == 2:
            pair = frozenset(dt)
            if mixed not in (None, pair):
                raise NotFusable("more than one mixed second derivative")
            mixed = pair
            for v in dt:
                if v not in first:
                    first.append(v)
        else:
            raise NotFusable(f"derivative {DIFF.join(dt)} (mixed, order > 2) has no fused channel")
    if mixed is not None and n_high:
        raise NotFusable("mixed and third/fourth-order derivatives in one net")
    order = sorted(second, key=inputs.index) + sorted([v for v in first if v not in second], key=inputs.index)
    if high_var is not None:  # the high-order input leads
        order = [high_var] + [v for v in order if v != high_var]
    if len(order) > 3:
        raise NotFusable("more than 3 differentiated inputs")
    if mixed is not None and set(order[:2]) != set(mixed):
        raise NotFusable("the mixed derivative must be of the first two differentiated inputs")
    # (padding an odd channel count to the packed-FMA program with an unused mixed channel was measured on the Poisson
    # step: 2.21 vs 1.95 ms per launch -- not done)
    return JetPlan(tuple(inputs.index(v) for v in order), len(second), 1 if mixed is not None else 0, n_high)


# --------------------------------------------------------------------------- #
def to_expr(e) -> sympy.Expr:
    return sympy.sympify(e) if isinstance(e, (str, int, float)) else e


def resolve(name: str, exprs: Mapping[str, sympy.Expr], net_outputs: Mapping[str, NetSpec],
            sampled: Sequence[str], _depth: int = 0) -> sympy.Expr:
    """Express `name` over leaf symbols (net-output jets and sampled keys) by
    substituting expression nodes, the way the reference's graph chains
    PdeNodes / ExpressionNodes by name (idrlnet/graph.py:95-121)."""
    if _depth > 32:
        raise NotFusable("expression nodes are recursive")
    base, derivs = split_name(name)
    if base in net_outputs or (name in sampled):
        return sympy.Symbol(name)
    if name in exprs:
        e = to_expr(exprs[name])
        sub = {}
        for s in e.free_symbols:
            r = resolve(s.name, exprs, net_outputs, sampled, _depth + 1)
            if r != s:
                sub[s] = r
        return e.subs(sub, simultaneous=True) if sub else e
    if base in exprs and derivs:
        raise NotFusable(f"{name}: derivative of an expression node")
    if base in sampled:
        raise NotFusable(f"{name}: derivative of a sampled key")
    raise NotFusable(f"{name}: cannot be computed")


def polynomial_terms(expr: sympy.Expr) -> List[Tuple[float, List[str]]]:
    """expr as sum of coef * product(symbols); raises NotFusable if it is not a
    polynomial with numeric coefficients."""
    expr = sympy.expand(expr)
    syms = sorted(expr.free_symbols, key=lambda s: s.name)
    if not syms:
        c = float(expr)
        return [(c, [])] if c != 0.0 else []
    try:
        poly = sympy.Poly(expr, *syms)
    except sympy.PolynomialError as e:
        raise NotFusable(f"residual is not a polynomial in its symbols: {e}")
    terms = []
    for monom, coef in poly.terms():
        try:
            c = float(coef)
        except TypeError:
            raise NotFusable(f"non-numeric coefficient {coef}")
        facs: List[str] = []
        for s, p in zip(syms, monom):
            if int(p) != p or p < 0:
                raise NotFusable("non-integer power")
            facs += [s.name] * int(p)
        if len(facs) > nv.MAX_FACTORS:
            raise NotFusable("monomial degree too high")
        terms.append((c, facs))
    return terms


@dataclass
class CompiledEquation:
    name: str
    terms: List[Tuple[float, List[int]]]  # (coef, symbol ids)


@dataclass
class CompiledDomain:
    net: NetSpec
    plan: JetPlan
    aux_keys: List[str]
    equations: List[CompiledEquation] = field(default_factory=list)

    @property
    def n_terms(self) -> int:
        return sum(len(e.terms) for e in self.equations)


def compile_domain(targets: Sequence[str], exprs: Mapping[str, object], nets: Sequence[NetSpec],
                   sampled: Sequence[str]) -> CompiledDomain:
    """Compile the loss terms `targets` of one sampling domain for the fused kernel."""
    net_outputs = {o: n for n in nets for o in n.outputs}
    exprs = {k: to_expr(v) for k, v in exprs.items()}
    resolved = [(t, resolve(t, exprs, net_outputs, sampled)) for t in targets]
    if len(resolved) > nv.MAX_EQ:
        raise NotFusable("too many loss terms in one domain")
    used_nets, jet_syms, other = [], [], []
    for _, e in resolved:
        for s in e.free_symbols:
            base, derivs = split_name(s.name)
            if base in net_outputs and s.name not in sampled:
                if net_outputs[base] not in used_nets:
                    used_nets.append(net_outputs[base])
                jet_syms.append((base, derivs))
            else:
                other.append(s.name)
    if len(used_nets) != 1:
        raise NotFusable(f"domain uses {len(used_nets)} nets (fused kernel handles exactly one)")
    net = used_nets[0]
    if len(net.inputs) > nv.MAX_IN or len(net.outputs) > nv.MAX_OUT:
        raise NotFusable("net has too many inputs/outputs")
    for i in net.inputs:
        if i not in sampled:
            raise NotFusable(f"net input {i} is not sampled directly (upstream expression)")
    plan = plan_jets(net.inputs, [d for _, d in jet_syms])
    aux_keys: List[str] = []
    for name in other:
        if name in net.inputs or name in aux_keys:
            continue
        if name not in sampled:
            raise NotFusable(f"symbol {name} is neither a net output nor sampled")
        aux_keys.append(name)
    if len(aux_keys) > nv.MAX_AUX:
        raise NotFusable("too many auxiliary symbols")

    def sym_id(name: str) -> int:
        base, derivs = split_name(name)
        if base in net_outputs and name not in sampled:
            return plan.channel(net.inputs, derivs) * nv.MAX_OUT + net.outputs.index(base)
        if name in net.inputs:
            return nv.SYM_COORD0 + net.inputs.index(name)
        return nv.SYM_AUX0 + aux_keys.index(name)

    out = CompiledDomain(net, plan, aux_keys)
    for tname, e in resolved:
        terms = [(c, [sym_id(f) for f in facs]) for c, facs in polynomial_terms(e)]
        out.equations.append(CompiledEquation(tname, terms))
    if out.n_terms > nv.MAX_TERMS:
        raise NotFusable("too many monomials in one domain")
    return out


Scalar = Union[int, float]


@dataclass(frozen=True)
class Ptr:
    """An address (device pointer, or host pointer for the test emulator)."""
    addr: int


def build_domain_struct(cd: CompiledDomain, n_points: int, ptr_of: Callable[[str], int],
                        targets: Mapping[str, Union[int, Scalar]], lambdas: Mapping[str, Union[int, Scalar, None]],
                        area: Union[int, Scalar], loss: str, sigma: float) -> nv.PinnDomain:
    """Fill a pinn_domain.  `ptr_of(key)` gives the address of a sampled column;
    targets / lambdas / area values are either `Ptr(address)` (a per-point
    column) or python floats (constants)."""
    d = nv.PinnDomain()
    d.n_points = n_points
    for i, k in enumerate(cd.net.inputs):
        d.coords[i] = ptr_of(k)
    for i, k in enumerate(cd.aux_keys):
        d.aux[i] = ptr_of(k)
    d.n_aux = len(cd.aux_keys)
    if isinstance(area, Ptr):
        d.area, d.area_const = area.addr, 0.0
    else:
        d.area, d.area_const = None, float(area)
    d.sigma = float(sigma)
    d.loss_kind = nv.LOSS[loss]
    d.n_eq = len(cd.equations)
    t = 0
    for i, eq in enumerate(cd.equations):
        q = d.eq[i]
        q.term_begin = t
        for coef, ids in eq.terms:
            tm = d.terms[t]
            tm.coef = coef
            tm.n_fac = len(ids)
            for f, sid in enumerate(ids):
                tm.fac[f] = sid
            t += 1
        q.term_end = t
        tg = targets[eq.name]
        if isinstance(tg, Ptr):
            q.target, q.target_const = tg.addr, 0.0
        else:
            q.target, q.target_const = None, float(tg)
        lm = lambdas.get(eq.name)
        if isinstance(lm, Ptr):
            q.lambda_, q.lambda_const = lm.addr, 0.0
        else:
            q.lambda_, q.lambda_const = None, 1.0 if lm is None else float(lm)
    d.n_terms = t
    return d
