"""CPU oracle for the PINN training step -- TEST INFRASTRUCTURE ONLY.

This file restates, with plain PyTorch autograd on the CPU, the algorithm the
reference (idrl-lab/idrlnet v2.0.0) runs for one training step.  It exists to
check the CUDA path; nothing in the product package (`idrlnet_b200/`) imports
it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import this module.

Pinning: the reference has no tests of its own for this path,
so the oracle is pinned against golden vectors produced by *executing the
reference itself* in the build container (`tests/golden/make_golden.py`,
fixtures in `tests/golden/*.npz`); `tests/test_oracle_golden.py` asserts the
agreement.

Each function cites the reference lines it follows (paths relative to the
reference checkout):

* `mlp_forward`            idrlnet/net.py:14-36, idrlnet/architecture/mlp.py:68-76,
                           idrlnet/architecture/layer.py:68-125
* `differentiate`          idrlnet/variable.py:161-208 (one torch.autograd.grad
                           call per (dependent, independent) pair, one order per
                           sweep, iterate to a fixed point), driven as in
                           idrlnet/graph.py:170-180
* `evaluate_expression`    idrlnet/torch_util.py:25-65 (sympy.lambdify onto a
                           torch namespace), idrlnet/node.py:105-109
* `weighted_loss`          idrlnet/variable.py:40-80, idrlnet/solver.py:353-408
* `training_step`          idrlnet/solver.py:329-339 (forward, loss, backward)

A "problem" is a plain dict (see `tests/golden/make_golden.py` for how one is
built from the reference's own objects):

    problem = {
      "nets":    [{"name", "inputs": [...], "outputs": [...], "act": "tanh|silu|sin",
                   "layers": [{"W": [out,in], "b": [out]}            # plain Linear
                              | {"g": [out,1], "v": [out,in], "b": [out]}],  # weight-norm
                   "shares": <name of the net whose weights it shares> | None}],
      "exprs":   {name: "sympy expr over symbols"},   # every PdeNode/ExpressionNode equation
      "domains": [{"name", "points": {key: [N,1]},
                   "targets": {name: [N,1] | float}, "lambdas": {name: [N,1] | float},
                   "loss": "square|L1|Identity", "sigma": float}],
    }
"""
from __future__ import annotations

import functools
import itertools
from typing import Dict, List, Sequence

import numpy as np
import sympy
import torch

DIFF = "__"  # idrlnet/header.py:6


# --------------------------------------------------------------------------- #
# activations (idrlnet/architecture/layer.py:68-125)
# --------------------------------------------------------------------------- #
def _act(name: str):
    if name == "tanh":
        return torch.tanh
    if name in ("silu", "swish"):
        return lambda z: z * torch.sigmoid(z)  # layer.py:124-125
    if name == "sin":
        return torch.sin
    raise ValueError(f"oracle: unsupported activation {name}")


def mlp_forward(cols: Sequence[torch.Tensor], layers, act: str) -> torch.Tensor:
    """cat(inputs, dim=1) -> Linear/activation stack; last layer linear."""
    h = torch.cat(list(cols), dim=1)  # net.py:18-26 (dict order = NetNode.inputs order)
    f = _act(act)
    for i, (W, b) in enumerate(layers):
        h = torch.nn.functional.linear(h, W, b)  # mlp.py:72
        if i < len(layers) - 1:
            h = f(h)
    return h


# --------------------------------------------------------------------------- #
# derivatives by nested autograd (idrlnet/variable.py:161-208)
# --------------------------------------------------------------------------- #
def _differentiate_one_step(var: Dict[str, torch.Tensor], required: List[str]) -> None:
    required = [d for d in required if d not in var]
    req = set(tuple(r.split(DIFF)) for r in required)
    dep = set(tuple(k.split(DIFF)) for k in var.keys())
    todo = {}
    for dv, rd in itertools.product(dep, req):
        if len(rd) > len(dv) and rd[: len(dv)] == dv and rd[: len(dv) + 1] not in dep:
            todo.setdefault(rd[len(dv)], set()).add(DIFF.join(dv))
    new = {}
    for key, deps in todo.items():
        for v in deps:
            g = torch.autograd.grad(
                var[v], var[key], grad_outputs=torch.ones_like(var[v]),
                retain_graph=True, create_graph=True, allow_unused=True,
            )[0]
            if g is None:
                g = torch.zeros_like(var[v], requires_grad=True)  # variable.py:191-194
            new[DIFF.join([v, key])] = g
    var.update(new)


def differentiate(var: Dict[str, torch.Tensor], required: List[str]) -> None:
    n = -1
    while n != len(var):  # variable.py:198-208
        n = len(var)
        _differentiate_one_step(var, required)


# --------------------------------------------------------------------------- #
# residual expressions (idrlnet/torch_util.py:25-65)
# --------------------------------------------------------------------------- #
_TORCH_NS = {
    "sin": torch.sin, "cos": torch.cos, "tan": torch.tan, "exp": torch.exp,
    "sqrt": torch.sqrt, "Abs": torch.abs, "tanh": torch.tanh, "log": torch.log,
    "sinh": torch.sinh, "cosh": torch.cosh,
    "Min": lambda *x: functools.reduce(torch.minimum, x),
    "Max": lambda *x: functools.reduce(torch.maximum, x),
}


@functools.lru_cache(maxsize=None)
def _compile(expr: str):
    e = sympy.sympify(expr)
    names = sorted(s.name for s in e.free_symbols)
    if not names:  # constant function, torch_util.py:31-36
        c = float(e)
        return names, (lambda **kw: None), c
    return names, sympy.lambdify([sympy.Symbol(n) for n in names], e, [_TORCH_NS]), None


def expression_symbols(expr: str) -> List[str]:
    return list(_compile(expr)[0])


def evaluate_expression(expr: str, var: Dict[str, torch.Tensor], like: torch.Tensor):
    names, fn, const = _compile(expr)
    if const is not None:
        return torch.zeros_like(like) + const
    return fn(*[var[n] for n in names])


# --------------------------------------------------------------------------- #
# loss (idrlnet/variable.py:40-80; idrlnet/solver.py:353-408)
# --------------------------------------------------------------------------- #
def weighted_loss(diff: Dict[str, torch.Tensor], lambdas, area, kind: str):
    f = {"square": lambda v: v ** 2, "L1": torch.abs, "Identity": lambda v: v}[kind]
    loss = 0.0
    for key, val in diff.items():
        lam = lambdas.get(key)
        if lam is not None:
            loss = loss + torch.sum(f(val) * lam * area)
        else:
            loss = loss + torch.sum(f(val) * area)
    return loss


# --------------------------------------------------------------------------- #
# one domain: forward through the node graph (idrlnet/graph.py:170-193)
# --------------------------------------------------------------------------- #
def _as_t(a, dtype, like=None):
    if isinstance(a, torch.Tensor):
        return a.to(dtype)
    a = np.asarray(a)
    if a.ndim == 0:
        assert like is not None
        return torch.zeros_like(like) + float(a)
    return torch.as_tensor(a, dtype=dtype).reshape(-1, 1)


def build_nets(problem, dtype=torch.float32):
    """Leaf tensors for the raw parameters; effective W = g*v/||v||_row for
    weight-norm layers (idrlnet/architecture/layer.py:42-43 ->
    torch.nn.utils.weight_norm, dim=0).  `retain_grad` on the effective W so the
    gradient w.r.t. it (what the CUDA step kernel emits) can be read back."""
    nets = []
    for n in problem["nets"]:
        if n.get("shares"):
            src = next(m for m in nets if m["name"] == n["shares"])
            nets.append({**n, "raw": src["raw"], "layers_t": src["layers_t"]})
            continue
        raw, eff = [], []
        for lay in n["layers"]:
            leaf = {k: torch.as_tensor(np.asarray(a), dtype=dtype).clone().requires_grad_()
                    for k, a in lay.items()}
            if "g" in leaf:
                W = torch._weight_norm(leaf["v"], leaf["g"], 0)
                W.retain_grad()
            else:
                W = leaf["W"]
            raw.append(leaf)
            eff.append((W, leaf["b"]))
        nets.append({**n, "raw": raw, "layers_t": eff})
    return nets


def forward_domain(problem, dom, nets, dtype=torch.float32, want: Sequence[str] = ()):
    """Return the dict of every variable computed on `dom`: sampled keys, net
    outputs, derivative symbols and expression/equation values."""
    var: Dict[str, torch.Tensor] = {}
    for k, v in dom["points"].items():
        t = _as_t(v, dtype).clone()
        if not (k.startswith("lambda_") or k == "area"):
            t.requires_grad_()  # variable.py:109-112
        var[k] = t
    like = next(iter(var.values()))
    exprs = problem.get("exprs", {})

    def ensure(name: str):
        if name in var:
            return
        if DIFF in name:  # derivative symbol: make the base, then differentiate
            ensure(name.split(DIFF)[0])
            differentiate(var, [name])
            if name not in var:
                raise KeyError(f"oracle: can't differentiate {name}")
            return
        if name in exprs:
            syms = expression_symbols(exprs[name])
            for s in syms:
                ensure(s)
            var[name] = evaluate_expression(exprs[name], var, like)
            return
        for net in nets:
            if name in net["outputs"]:
                for i in net["inputs"]:
                    ensure(i)
                out = mlp_forward([var[i] for i in net["inputs"]], net["layers_t"], net["act"])
                for j, o in enumerate(net["outputs"]):
                    var[o] = out[:, j: j + 1]  # net.py:31-35
                return
        raise KeyError(f"oracle: can't compute {name}")

    for name in list(dom.get("targets", {}).keys()) + list(want):
        ensure(name)
    return var


def training_step(problem, dtype=torch.float32, want: Dict[str, Sequence[str]] = None):
    """One forward + loss + backward over every domain.  Returns numpy results:
    {"loss", "loss_components": {domain: float},
     "grads": {net: [{"W"|"g","v","b","W_eff": array|None}, ...]},
     "domains": {name: {symbol: [N,1]}}}"""
    nets = build_nets(problem, dtype)
    want = want or {}
    total = 0.0
    comps, dom_out = {}, {}
    for dom in problem["domains"]:
        var = forward_domain(problem, dom, nets, dtype, want.get(dom["name"], ()))
        like = next(iter(var.values()))
        diff, lambdas = {}, {}
        for key, tgt in dom.get("targets", {}).items():
            diff[key] = var[key] - _as_t(tgt, dtype, like)  # variable.py:84-91
            lam = dom.get("lambdas", {}).get(key)
            if lam is None and ("lambda_" + key) in dom["points"]:
                lam = dom["points"]["lambda_" + key]  # solver.py:371-378
            if lam is not None:
                lambdas[key] = _as_t(lam, dtype, like)
        if diff:
            ld = weighted_loss(diff, lambdas, var["area"], dom.get("loss", "square"))
            comps[dom["name"]] = ld
            total = total + float(dom.get("sigma", 1.0)) * ld  # solver.py:392-398
        dom_out[dom["name"]] = {k: v.detach().numpy().copy() for k, v in var.items()}

    if isinstance(total, torch.Tensor) and total.requires_grad:
        total.backward()  # solver.py:338
    grads = {}
    for n in nets:
        if n.get("shares"):
            continue
        out = []
        for leaf, (W, _) in zip(n["raw"], n["layers_t"]):
            g = {k: (None if t.grad is None else t.grad.numpy().copy()) for k, t in leaf.items()}
            g["W_eff"] = None if W.grad is None else W.grad.numpy().copy()
            out.append(g)
        grads[n["name"]] = out
    return {
        "loss": float(total.detach()) if isinstance(total, torch.Tensor) else float(total),
        "loss_components": {k: float(v.detach()) for k, v in comps.items()},
        "grads": grads,
        "domains": dom_out,
    }


# --------------------------------------------------------------------------- #
# (de)serialisation of problems to a flat npz (used by tests/golden)
# --------------------------------------------------------------------------- #
def problem_to_flat(problem) -> Dict[str, np.ndarray]:
    import json
    flat = {}
    meta = {"nets": [], "exprs": problem.get("exprs", {}), "domains": []}
    for n in problem["nets"]:
        m = {k: n[k] for k in ("name", "inputs", "outputs", "act")}
        m["shares"] = n.get("shares")
        m["n_layers"] = len(n.get("layers", []))
        m["layer_keys"] = [sorted(l.keys()) for l in n.get("layers", [])]
        for i, lay in enumerate(n.get("layers", [])):
            for k, a in lay.items():
                flat[f"net/{n['name']}/{i}/{k}"] = np.asarray(a)
        meta["nets"].append(m)
    for d in problem["domains"]:
        m = {"name": d["name"], "loss": d.get("loss", "square"), "sigma": float(d.get("sigma", 1.0)),
             "points": list(d["points"].keys()), "targets": {}, "lambdas": {}}
        for k, a in d["points"].items():
            flat[f"dom/{d['name']}/points/{k}"] = np.asarray(a)
        for grp in ("targets", "lambdas"):
            for k, a in d.get(grp, {}).items():
                if np.ndim(a) == 0:
                    m[grp][k] = float(a)
                else:
                    m[grp][k] = None
                    flat[f"dom/{d['name']}/{grp}/{k}"] = np.asarray(a)
        meta["domains"].append(m)
    flat["meta"] = np.array(json.dumps(meta))
    return flat


def problem_from_flat(flat) -> dict:
    import json
    meta = json.loads(str(flat["meta"]))
    problem = {"nets": [], "exprs": meta["exprs"], "domains": []}
    for m in meta["nets"]:
        n = {k: m[k] for k in ("name", "inputs", "outputs", "act", "shares")}
        n["layers"] = [{k: np.asarray(flat[f"net/{m['name']}/{i}/{k}"]) for k in keys}
                       for i, keys in enumerate(m["layer_keys"])]
        problem["nets"].append(n)
    for m in meta["domains"]:
        d = {"name": m["name"], "loss": m["loss"], "sigma": m["sigma"],
             "points": {k: np.asarray(flat[f"dom/{m['name']}/points/{k}"]) for k in m["points"]},
             "targets": {}, "lambdas": {}}
        for grp in ("targets", "lambdas"):
            for k, v in m[grp].items():
                d[grp][k] = v if v is not None else np.asarray(flat[f"dom/{m['name']}/{grp}/{k}"])
        problem["domains"].append(d)
    return problem


def topk_abs(values, k):
    """Adaptive re-sampling selection, examples/allen_cahn/allen_cahn.py:137-141:
    `np.argsort(-1.0 * np.abs(residual))[:k]`.  The reference uses NumPy's default (unstable)
    sort, so ties may come in any order there; the restatement fixes them by ascending index
    (`kind="stable"`), which is one of the orders the reference can produce.  NaNs sort last in
    NumPy (i.e. after every finite |r|); the CUDA path ranks them first -- residuals are finite."""
    import numpy as np
    v = np.abs(np.asarray(values, dtype=np.float32).ravel())
    return np.argsort(-v, kind="stable")[:k]
