"""Test helpers: run an oracle-format problem (see oracle/pinn_oracle.py)
through the kernels' descriptors -- on the CPU emulator (tests/emu, host
pointers) -- and build random problems."""
import ctypes as C
import os
import subprocess

import numpy as np

from idrlnet_b200 import compiler as cp
from idrlnet_b200 import native as nv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
_EMU = None


def emu_lib():
    global _EMU
    if _EMU is None:
        so = os.path.join(EMU_DIR, "libpinn_emu.so")
        src = os.path.join(EMU_DIR, "emu.cpp")
        deps = [src, os.path.join(ROOT, "include", "pinn_b200.h")] + [
            os.path.join(ROOT, "idrlnet_b200", "csrc", f) for f in ("narrow_warp.h", "pinn_common.h", "philox.h")]
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-DPINN_EMU",
                                   "-fvisibility=hidden", "-o", so, src])
        h = C.CDLL(so)
        h.pinn_emu_run.restype = C.c_int
        h.pinn_emu_run.argtypes = [C.POINTER(nv.PinnMlp), C.POINTER(nv.PinnJets), C.c_int, C.POINTER(nv.PinnDomain),
                                   C.c_int64, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_int]
        h.pinn_emu_sample_box.restype = None
        h.pinn_emu_sample_box.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                          C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        _EMU = h
    return _EMU


def effective_layers(net):
    """[(W, b)] float32 effective weights of an oracle-format net."""
    out = []
    for lay in net["layers"]:
        if "g" in lay:
            v = np.asarray(lay["v"], np.float64)
            g = np.asarray(lay["g"], np.float64).reshape(-1, 1)
            W = (g * v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
        else:
            W = np.asarray(lay["W"], np.float32)
        out.append((np.ascontiguousarray(W), np.ascontiguousarray(np.asarray(lay["b"], np.float32))))
    return out


def net_specs(problem):
    return [cp.NetSpec(n["name"], tuple(n["inputs"]), tuple(n["outputs"])) for n in problem["nets"]]


def owner_net(problem, name):
    n = next(n for n in problem["nets"] if n["name"] == name)
    if n.get("shares"):
        n = next(m for m in problem["nets"] if m["name"] == n["shares"])
    return n


def flat_grad_from_oracle(res, net_name):
    """Oracle W_eff/b grads of a net, one entry per flat-layout segment (None = no grad)."""
    parts = []
    for g in res["grads"][net_name]:
        W = g["W_eff"] if g.get("W_eff") is not None else g.get("W")
        parts.append(None if W is None else W.ravel())
        parts.append(g["b"].ravel() if g["b"] is not None else None)
    return parts


def flat_layout_sizes(layers):
    return [a.size for Wb in layers for a in Wb]


def mlp_struct(layers, n_in, n_out, act, ptr):
    width = layers[0][0].shape[0]
    n_hidden = len(layers) - 1
    for W, _ in layers[1:-1]:
        assert W.shape == (width, width)
    return nv.make_mlp(n_in, n_out, n_hidden, width, act, [ptr(W) for W, _ in layers], [ptr(b) for _, b in layers])


def domain_columns(dom):
    """float32 contiguous [N] columns of everything sampled on the domain."""
    return {k: np.ascontiguousarray(np.asarray(v, np.float32).reshape(-1)) for k, v in dom["points"].items()}


def compile_problem_domain(problem, dom):
    cols = domain_columns(dom)
    return cp.compile_domain(list(dom["targets"].keys()), problem.get("exprs", {}), net_specs(problem), list(cols))


def run_emu_fused(problem, n_cta=3, n_warps=2, mode=0):
    """Every fusable domain through the emulator.  Returns
    {net: flat grad}, {domain: loss}, [names of domains that were not fusable]."""
    lib = emu_lib()
    grads, losses, skipped = {}, {}, []
    keep = []
    for dom in problem["domains"]:
        if not dom.get("targets"):
            continue
        try:
            cd = compile_problem_domain(problem, dom)
        except cp.NotFusable:
            skipped.append(dom["name"])
            continue
        cols = domain_columns(dom)
        n = len(next(iter(cols.values())))
        net = owner_net(problem, cd.net.name)
        layers = effective_layers(net)
        keep.append(layers)
        mlp = mlp_struct(layers, len(cd.net.inputs), len(cd.net.outputs), net["act"], lambda a: a.ctypes.data)
        jets = nv.make_jets(cd.plan.first_in, cd.plan.n_second, cd.plan.n_mixed, cd.plan.n_high)

        def val(v):
            if np.ndim(v) == 0:
                return float(v)
            a = np.ascontiguousarray(np.asarray(v, np.float32).reshape(-1))
            keep.append(a)
            return cp.Ptr(a.ctypes.data)

        targets = {k: val(v) for k, v in dom["targets"].items()}
        lambdas = {k: val(v) for k, v in dom.get("lambdas", {}).items()}
        for k in dom["targets"]:
            if k not in lambdas and ("lambda_" + k) in cols:
                lambdas[k] = cp.Ptr(cols["lambda_" + k].ctypes.data)
        d = cp.build_domain_struct(cd, n, lambda k: cols[k].ctypes.data, targets, lambdas,
                                   cp.Ptr(cols["area"].ctypes.data), dom.get("loss", "square"),
                                   float(dom.get("sigma", 1.0)))
        owner = net["name"]
        if owner not in grads:
            grads[owner] = np.zeros(sum(flat_layout_sizes(layers)), np.float32)
        loss = np.zeros(1, np.float32)
        rc = lib.pinn_emu_run(C.byref(mlp), C.byref(jets), mode, C.byref(d), n, None, None, None,
                              grads[owner].ctypes.data, loss.ctypes.data, n_cta, n_warps)
        assert rc == 0, rc
        losses[dom["name"]] = float(loss[0])
    return grads, losses, skipped


def random_problem(rng, n_in=2, n_out=1, width=20, n_hidden=3, act="tanh", n_points=100, weight_norm=False,
                   equation="u*u__x + u__t - 0.01*u__x__x", inputs=("x", "t"), outputs=("u",), loss="square",
                   target=0.0, with_lambda=False, sigma=1.0, scale=1.0):
    dims = [n_in] + [width] * n_hidden + [n_out]
    layers = []
    for i in range(len(dims) - 1):
        W = (rng.standard_normal((dims[i + 1], dims[i])) * scale / np.sqrt(dims[i])).astype(np.float32)
        b = (rng.standard_normal(dims[i + 1]) * 0.1).astype(np.float32)
        if weight_norm:
            g = (np.linalg.norm(W, axis=1, keepdims=True) * rng.uniform(0.5, 1.5, (dims[i + 1], 1))).astype(np.float32)
            layers.append({"g": g, "v": W, "b": b})
        else:
            layers.append({"W": W, "b": b})
    pts = {k: rng.uniform(-1, 1, (n_points, 1)).astype(np.float32) for k in inputs}
    pts["area"] = rng.uniform(0.5, 1.5, (n_points, 1)).astype(np.float32) / max(n_points, 1)
    dom = {"name": "dom", "points": pts, "targets": {"eq": target}, "lambdas": {}, "loss": loss, "sigma": sigma}
    if with_lambda:
        dom["lambdas"]["eq"] = rng.uniform(0.0, 2.0, (n_points, 1)).astype(np.float32)
    return {"nets": [{"name": "net", "inputs": list(inputs), "outputs": list(outputs), "act": act,
                      "layers": layers, "shares": None}],
            "exprs": {"eq": equation}, "domains": [dom]}


# --------------------------------------------------------------------------- #
# CUDA side: oracle-format problem -> torch modules -> idrlnet_b200.fused
# --------------------------------------------------------------------------- #
def torch_module_from_net(net, device):
    """An idrlnet_b200.architecture.MLP holding the oracle-format net's parameters."""
    import torch
    from idrlnet_b200 import architecture as arch
    layers = net["layers"]
    wn = "g" in layers[0]
    first = layers[0]["v"] if wn else layers[0]["W"]
    seq = [np.asarray(first).shape[1]] + [np.asarray(l["b"]).shape[0] for l in layers]
    act = {"tanh": arch.Activation.tanh, "silu": arch.Activation.silu, "sin": arch.Activation.sin}[net["act"]]
    m = arch.MLP(seq, activation=act, weight_norm=wn)
    lins = [l for k, l in m.layers.items() if not k.endswith("_activation")]
    with torch.no_grad():
        for lin, lay in zip(lins, layers):
            lin.bias.copy_(torch.as_tensor(np.asarray(lay["b"], np.float32)))
            if wn:
                lin.weight_g.copy_(torch.as_tensor(np.asarray(lay["g"], np.float32)).reshape(lin.weight_g.shape))
                lin.weight_v.copy_(torch.as_tensor(np.asarray(lay["v"], np.float32)))
            else:
                lin.weight.copy_(torch.as_tensor(np.asarray(lay["W"], np.float32)))
    return m.to(device)


class cuda_packs(dict):
    """Lazily built {owner net name: fused.NetPack} (+ .modules) of a problem on the GPU."""

    def __init__(self, problem, device="cuda:0", precision="fp32"):
        super().__init__()
        self.problem, self.device, self.precision, self.modules = problem, device, precision, {}

    def get_pack(self, net_name):
        from idrlnet_b200 import fused
        owner = owner_net(self.problem, net_name)
        if owner["name"] not in self:
            self.modules[owner["name"]] = torch_module_from_net(owner, self.device)
            self[owner["name"]] = fused.NetPack(self.modules[owner["name"]], act=owner["act"], precision=self.precision)
        return owner["name"], self[owner["name"]]


def bind_fused_domain(problem, dom, packs):
    """(fused.FusedDomain bound to the domain's columns on the GPU, its NetPack); raises NotFusable."""
    import torch
    from idrlnet_b200 import fused
    cd = compile_problem_domain(problem, dom)
    _, pack = packs.get_pack(cd.net.name)
    device = packs.device
    cols = {k: torch.as_tensor(v, device=device) for k, v in domain_columns(dom).items()}

    def val(v):
        return float(v) if np.ndim(v) == 0 else torch.as_tensor(np.asarray(v, np.float32).reshape(-1), device=device)

    targets = {k: val(v) for k, v in dom["targets"].items()}
    lambdas = {k: val(v) for k, v in dom.get("lambdas", {}).items()}
    for k in dom["targets"]:
        if k not in lambdas and ("lambda_" + k) in cols:
            lambdas[k] = cols["lambda_" + k]
    fd = fused.FusedDomain(dom["name"], cd, dom.get("loss", "square"), float(dom.get("sigma", 1.0)))
    fd.bind(cols, targets, lambdas, cols["area"])
    return fd, pack


def run_cuda_fused(problem, device="cuda:0", precision="fp32"):
    """All fusable domains through fused.StepEngine (the C ABI).  Returns
    ({net: flat eff grad (numpy)}, {domain: loss}, skipped, {net: module})."""
    import torch
    from idrlnet_b200 import fused
    packs = cuda_packs(problem, device, precision)
    per_net = {}
    skipped = []
    for dom in problem["domains"]:
        if not dom.get("targets"):
            continue
        try:
            fd, pack = bind_fused_domain(problem, dom, packs)
        except cp.NotFusable:
            skipped.append(dom["name"])
            continue
        name = next(k for k, v in packs.items() if v is pack)
        per_net.setdefault(name, []).append(fd)
    grads, losses = {}, {}
    for name, doms in per_net.items():
        eng = fused.StepEngine(packs[name], doms)
        eng.step(scatter=True)
        torch.cuda.synchronize()
        grads[name] = eng.grad.detach().cpu().numpy().copy()
        for d, l in zip(doms, eng.losses.detach().cpu().numpy()):
            losses[d.name] = float(l)
    return grads, losses, skipped, packs.modules


# --------------------------------------------------------------------------- #
# Solver-level helpers
# --------------------------------------------------------------------------- #
def solver_from_problem(problem, network_dir, device=None, **solver_kw):
    """Build an idrlnet_b200 Solver (public API) from an oracle-format problem with FIXED points."""
    import sympy
    import torch
    import idrlnet_b200.shortcut as sc
    nodes, modules = [], {}
    for n in problem["nets"]:
        if n.get("shares"):
            src = next(m for m in nodes if m.name == n["shares"])
            nodes.append(sc.get_shared_net_node(src, inputs=tuple(n["inputs"]), outputs=tuple(n["outputs"]), name=n["name"]))
            continue
        mod = torch_module_from_net(n, "cpu")
        modules[n["name"]] = mod
        nodes.append(sc.NetNode(inputs=tuple(n["inputs"]), outputs=tuple(n["outputs"]), net=mod, name=n["name"]))
    pdes = [sc.ExpressionNode(expression=sympy.sympify(e), name=k) for k, e in problem.get("exprs", {}).items()]
    doms = []
    for d in problem["domains"]:
        pts = {k: np.asarray(v, np.float32) for k, v in d["points"].items()}
        cons = {k: (np.asarray(v, np.float32) if np.ndim(v) else float(v)) for k, v in d["targets"].items()}
        for k, v in d.get("lambdas", {}).items():
            cons["lambda_" + k] = np.asarray(v, np.float32) if np.ndim(v) else float(v)
        doms.append(sc.get_data_node((lambda p=pts, c=cons: (dict(p), dict(c))), name=d["name"],
                                     loss_fn=d.get("loss", "square"), sigma=float(d.get("sigma", 1.0))))
    s = sc.Solver(sample_domains=tuple(doms), netnodes=nodes, pdes=pdes, network_dir=str(network_dir), loading=False,
                  **solver_kw)
    return s, modules


def _oracle_total(po, problem, nets, owners, dtype=None):
    import torch
    dtype = dtype or torch.float32
    for n in owners:
        n["layers_t"][:] = [((torch._weight_norm(l["v"], l["g"], 0) if "g" in l else l["W"]), l["b"]) for l in n["raw"]]
    total = 0.0
    for dom in problem["domains"]:
        var = po.forward_domain(problem, dom, nets, dtype=dtype)
        like = next(iter(var.values()))
        diff = {k: var[k] - po._as_t(t, dtype, like) for k, t in dom["targets"].items()}
        lam = {k: po._as_t(t, dtype, like) for k, t in dom.get("lambdas", {}).items()}
        for k in dom["targets"]:
            if k not in lam and ("lambda_" + k) in dom["points"]:
                lam[k] = po._as_t(dom["points"]["lambda_" + k], dtype, like)
        total = total + float(dom.get("sigma", 1.0)) * po.weighted_loss(diff, lam, var["area"], dom.get("loss", "square"))
    return total


def oracle_train(problem, steps, lr=1e-3, gamma=0.95 ** 0.001, lbfgs=None, dtype=None):
    """The reference training loop on the oracle (CPU, FP32): Adam + ExponentialLR
    (idrlnet/optim.py:70-80), fixed points; returns the per-step (pre-update) losses.
    `lbfgs=dict(lr=, max_iter=)` runs the closure branch instead (idrlnet/solver.py:316-326);
    the value returned per step is then the first closure evaluation's loss, as torch's LBFGS does."""
    import torch
    from oracle import pinn_oracle as po
    nets = po.build_nets(problem, dtype=dtype or torch.float32)
    owners = [n for n in nets if not n.get("shares")]
    leaves = [t for n in owners for lay in n["raw"] for t in lay.values()]
    if lbfgs is not None:
        opt = torch.optim.LBFGS(leaves, **lbfgs)

        def closure():
            opt.zero_grad()
            total = _oracle_total(po, problem, nets, owners, dtype)
            total.backward()
            return total
        return [float(opt.step(closure).detach()) for _ in range(steps)]
    opt = torch.optim.Adam(leaves, lr=lr)
    losses = []
    for step in range(steps):
        opt.zero_grad()
        total = _oracle_total(po, problem, nets, owners, dtype)
        total.backward()
        opt.step()
        for g in opt.param_groups:
            g["lr"] = lr * gamma ** (step + 1)
        losses.append(float(total.detach()))
    return losses
