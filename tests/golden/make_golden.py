"""Generate golden vectors by EXECUTING THE REFERENCE (idrl-lab/idrlnet).

Run in the build container only (`/root/reference` does not exist on the GPU
box):   python tests/golden/make_golden.py

The reference is imported read-only from /root/reference with the five
import-time shims of SURVEY.md section 8(c1) / Appendix B.  For each BASELINE.json
config (scaled down to a few hundred points so the fixtures stay small) the
script builds the problem with the reference's own classes exactly as the
example script does, then runs ONE strict training pass

    forward_through_all_graph -> compute_loss -> loss.backward()

(idrlnet/solver.py:329-338) in FP32 and again in FP64 (same weights, same
points), and stores: the points, the raw + effective net parameters, every
variable the graph computed (net outputs, derivative symbols, residuals), the
per-domain losses, the total loss and the parameter gradients.

Output: tests/golden/<config>.npz  (problem in `oracle.pinn_oracle` flat
format + "ref32/..." and "ref64/..." result arrays).
"""
import collections
import collections.abc
import copy
import json
import math
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


# ----------------------------------------------------------------------- shims
def _install_shims():
    def _stub(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    class _Any:
        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

        def __iter__(self):
            return iter([_Any(), _Any()])

    def _ga(k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Any()

    _stub("pyevtk")
    _stub("pyevtk.hl", pointsToVTK=lambda *a, **k: None)
    mp = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    plt.__getattr__ = _ga
    _stub("matplotlib.tri").__getattr__ = _ga
    mp.pyplot = plt
    collections.Callable = collections.abc.Callable
    sys.path.insert(0, REF)


_install_shims()
os.chdir(tempfile.mkdtemp(prefix="idrl_golden_"))  # reference writes train.log etc. into CWD
import torch  # noqa: E402
import sympy as sp  # noqa: E402
import idrlnet.shortcut as sc  # noqa: E402
from idrlnet.optim import Optimizable, get_available_class  # noqa: E402
import idrlnet.graph as _g  # noqa: E402

Optimizable.SCHEDULE_MAP = get_available_class(torch.optim.lr_scheduler, torch.optim.lr_scheduler.LRScheduler)
_g.VertexTaskPipeline.display = lambda self, filename=None: None

sys.path.insert(0, ROOT)
from oracle.pinn_oracle import problem_to_flat  # noqa: E402


# ------------------------------------------------------------ problem builders
def cfg_simple_poisson():
    """examples/simple_poisson/simple_poisson.py:8-56 at density 64 (as shipped: 1000)."""
    x, y = sp.symbols("x y")
    rec = sc.Rectangle((-1.0, -1.0), (1.0, 1.0))

    @sc.datanode
    class LeftRight(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_boundary(32, sieve=((y > -1.0) & (y < 1.0))), {"T": 0.0}

    @sc.datanode(name="up_down")
    class UpDown(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_boundary(32, sieve=((x > -1.0) & (x < 1.0))), {"normal_gradient_T": 0.0}

    @sc.datanode(name="heat_domain")
    class Heat(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_interior(64), {"diffusion_T": 1.0}

    net = sc.get_net_node(inputs=("x", "y"), outputs=("T",), name="net1", arch=sc.Arch.mlp)
    pde = sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False)
    grad = sc.NormalGradient("T", dim=2, time=False)
    return (Heat(), LeftRight(), UpDown()), [net], [pde, grad]


def cfg_burgers():
    """examples/burgers_equation/burgers_equation.py:7-52, densities /40."""
    x, t = sp.Symbol("x"), sp.Symbol("t")
    time_range = {t: (0, 1)}
    geo = sc.Line1D(-1.0, 1.0)

    @sc.datanode(name="burgers_equation")
    def interior_domain():
        return geo.sample_interior(250, bounds={x: (-1.0, 1.0)}, param_ranges=time_range), {"burgers_u": 0}

    @sc.datanode(name="t_boundary")
    def init_domain():
        return geo.sample_interior(25, param_ranges={t: 0.0}), sc.Variables({"u": -sp.sin(math.pi * x)})

    @sc.datanode(name="x_boundary")
    def boundary_domain():
        return geo.sample_boundary(25, param_ranges=time_range), sc.Variables({"u": 0})

    net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    pde = sc.BurgersNode(u="u", v=0.01 / math.pi)
    return (interior_domain(), init_domain(), boundary_domain()), [net], [pde]


def cfg_allen_cahn():
    """examples/allen_cahn/allen_cahn.py:13-122,197-210, densities /10; the
    plotting-only `allen_test` domain (no targets) is left out."""
    t_symbol, x = sp.Symbol("t"), sp.Symbol("x")
    geo = sc.Line1D(-1.0, 1.0)
    u = sp.Function("u")(x, t_symbol)
    up = sp.Function("up")(x, t_symbol)
    time_range = {t_symbol: (0, 1.0)}

    @sc.datanode
    class AllenInit(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return geo.sample_interior(density=30, param_ranges={t_symbol: 0.0}), {
                "u": x ** 2 * sp.cos(sp.pi * x), "lambda_u": 100}

    @sc.datanode
    class AllenBc(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return geo.sample_boundary(density=20, sieve=sp.Eq(x, -1), param_ranges=time_range), {
                "difference_u_up": 0, "difference_diff_u_diff_up": 0}

    @sc.datanode(name="allen_domain")
    class AllenEq(sc.SampleDomain):
        def __init__(self):
            self.points = geo.sample_interior(density=200, param_ranges=time_range, low_discrepancy=True)

        def sampling(self, *args, **kwargs):
            return self.points, {"AllenCahn_u": 0}

    @sc.datanode(name="re_sampling_domain")
    class SpaceAdaptiveSampling(sc.SampleDomain):
        def __init__(self):
            self.points = geo.sample_interior(density=10, param_ranges=time_range, low_discrepancy=True)
            self.points = sc.Variables(self.points).to_torch_tensor_()
            self.constraints = {"AllenCahn_u": torch.zeros_like(self.points["x"])}

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    net_u = sc.MLP([2, 128, 128, 128, 128, 2], activation=sc.Activation.tanh)
    net_u = sc.NetNode(inputs=("x", "t"), outputs=("u",), name="net1", net=net_u)
    xp = sc.ExpressionNode(name="xp", expression=x + 2)
    tilde = sc.get_shared_net_node(net_u, inputs=("xp", "t"), outputs=("up",), name="net2", arch="mlp")
    diff_u = sc.ExpressionNode(expression=u.diff(x), name="diff_u")
    diff_up = sc.ExpressionNode(expression=up.diff(x), name="diff_up")
    pde = sc.AllenCahnNode(u="u", gamma_1=0.0001, gamma_2=5)
    boundary_up = sc.Difference(T="diff_u", S="diff_up")
    boundary_u = sc.Difference(T="u", S="up")
    return ((AllenInit(), AllenBc(), AllenEq(), SpaceAdaptiveSampling()),
            [net_u, tilde], [pde, xp, diff_up, diff_u, boundary_up, boundary_u])


def cfg_ldc():
    """examples/ldc/ldc.py:13-74, densities /16 (anomaly detection left off)."""
    from idrlnet.pde_op.equations import NavierStokesNode
    x, y = sp.symbols("x y")
    rec = sc.Rectangle((-0.05, -0.05), (0.05, 0.05))

    @sc.datanode(name="flow_domain")
    class Interior(sc.SampleDomain):
        def __init__(self):
            points = rec.sample_interior(25000, bounds={x: (-0.05, 0.05), y: (-0.05, 0.05)})
            n = len(points["x"])
            constraints = sc.Variables({k: torch.zeros(n, 1) for k in ("continuity", "momentum_x", "momentum_y")})
            points["area"] = np.ones_like(points["area"]) * 2.5e-6
            for k in ("continuity", "momentum_x", "momentum_y"):
                constraints["lambda_" + k] = points["sdf"]
            self.points = sc.Variables(points).to_torch_tensor_()
            self.constraints = sc.Variables(constraints).to_torch_tensor_()

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    @sc.datanode(name="left_right_down")
    class LRD(sc.SampleDomain):
        def __init__(self):
            points = rec.sample_boundary(208, sieve=(y < 0.05))
            n = len(points["x"])
            constraints = sc.Variables({"u": torch.zeros(n, 1), "v": torch.zeros(n, 1)})
            points["area"] = np.ones_like(points["area"]) * 1e-4
            self.points = sc.Variables(points).to_torch_tensor_()
            self.constraints = sc.Variables(constraints).to_torch_tensor_()

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    @sc.datanode(name="up")
    class Up(sc.SampleDomain):
        def __init__(self):
            points = rec.sample_boundary(625, sieve=sp.Eq(y, 0.05))
            n = len(points["x"])
            points["area"] = np.ones_like(points["area"]) * 1e-4
            constraints = sc.Variables({"u": torch.ones(n, 1), "v": torch.zeros(n, 1)})
            constraints["lambda_u"] = 1 - 20 * abs(points["x"].copy())
            self.points = sc.Variables(points).to_torch_tensor_()
            self.constraints = sc.Variables(constraints).to_torch_tensor_()

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    net = sc.MLP([2, 100, 100, 100, 100, 3], activation=sc.Activation.tanh,
                 initialization=sc.Initializer.Xavier_uniform, weight_norm=False)
    net_u = sc.NetNode(inputs=("x", "y"), outputs=("u", "v", "p"), net=net, name="net_u")
    pde = NavierStokesNode(nu=0.01, rho=1.0, dim=2, time=False)
    return (Interior(), LRD(), Up()), [net_u], [pde]


def _cfg_ns_unsteady(net_builder):
    """examples/ns_unsteady/NS_unsteady.py:13-49 (Adam instead of LBFGS; first
    64 rows of the example's CSV; density 2000 -> 8)."""
    import pandas as pd
    t = sp.Symbol("t")
    geo = sc.Rectangle((1.0, -2.0), (8.0, 2.0))
    time_range = {t: (0, 20)}
    csv = pd.read_csv(os.path.join(REF, "examples/ns_unsteady/NSexternel_sample.csv")).iloc[:64]

    @sc.datanode(name="NS_domain", loss_fn="L1")
    class NSExternal(sc.SampleDomain):
        def __init__(self):
            self.points = {col: csv[col].to_numpy().reshape(-1, 1) for col in csv.columns}
            self.constraints = {k: self.points.pop(k) for k in ("u", "v", "p")}

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    @sc.datanode(name="NS_external")
    class NSEq(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return geo.sample_interior(density=8, param_ranges=time_range), {
                "continuity": 0, "momentum_x": 0, "momentum_y": 0}

    net = net_builder()
    pde = sc.NavierStokesNode(nu=0.01, rho=1.0, dim=2, time=True)
    return (NSExternal(), NSEq()), [net], [pde]


def cfg_ns_unsteady():
    return _cfg_ns_unsteady(lambda: sc.get_net_node(
        inputs=("x", "y", "t"), outputs=("u", "v", "p"), name="net", arch=sc.Arch.mlp))


def cfg_ns_unsteady_wide():
    """Same problem with the `Arch.mlp_xl` recipe (SiLU + weight-norm,
    idrlnet/architecture/mlp.py:235-246) at width 64x3 so the fixture stays small."""
    return _cfg_ns_unsteady(lambda: sc.get_net_node(
        inputs=("x", "y", "t"), outputs=("u", "v", "p"), name="net", arch=sc.Arch.mlp_xl,
        seq=[3, 64, 64, 64, 3]))


def cfg_ns_unsteady_xl():
    """The BASELINE configs[4] net at its TRUE shape: `Arch.mlp_xl` = 3-512x6-3 SiLU + weight-norm
    (idrlnet/architecture/mlp.py:235-246), 1 319 942 parameters -- the shape `bench.py --config ns_xl` times."""
    return _cfg_ns_unsteady(lambda: sc.get_net_node(
        inputs=("x", "y", "t"), outputs=("u", "v", "p"), name="net", arch=sc.Arch.mlp_xl))


# fixtures whose parameter gradients are stored in float32 only (the FP64 run's, rounded; 5 MB instead of 16 MB)
COMPACT = {"ns_unsteady_xl"}

CONFIGS = {
    "ns_unsteady_xl": cfg_ns_unsteady_xl,
    "simple_poisson": cfg_simple_poisson,
    "burgers": cfg_burgers,
    "allen_cahn": cfg_allen_cahn,
    "ldc": cfg_ldc,
    "ns_unsteady": cfg_ns_unsteady,
    "ns_unsteady_wide": cfg_ns_unsteady_wide,
}


# --------------------------------------------------------- reference execution
def _act_name(mlp):
    for name, layer in mlp.layers.items():
        if name.endswith("_activation"):
            return {"tanh": "tanh", "swish": "silu", "silu": "silu", "sin": "sin"}[type(layer).__name__]
    raise ValueError("no activation")


def _net_layers(mlp):
    out = []
    for name, layer in mlp.layers.items():
        if name.endswith("_activation"):
            continue
        if hasattr(layer, "weight_g"):
            out.append({"g": layer.weight_g.detach().numpy().copy(),
                        "v": layer.weight_v.detach().numpy().copy(),
                        "b": layer.bias.detach().numpy().copy()})
        else:
            out.append({"W": layer.weight.detach().numpy().copy(), "b": layer.bias.detach().numpy().copy()})
    return out


def _strict_pass(solver, samples_np, dtype):
    """solver.py:329-338 on fixed samples."""
    from idrlnet.variable import Variables
    for nn in solver.netnodes:
        nn.net.to(dtype)
        nn.net.zero_grad(set_to_none=True)
    samples = {}
    for dname, d in samples_np.items():
        v = Variables()
        for k, a in d.items():
            t = torch.tensor(np.asarray(a, dtype=np.float32), dtype=dtype)
            if not (k.startswith("lambda_") or k == "area"):
                t.requires_grad_()  # variable.py:109-112
            v[k] = t
        samples[dname] = v
    in_var, true_out, lambda_out = solver.generate_in_out_dict(samples)
    allvars, pred = {}, {}
    for dname, names in solver.outvar_dict_index.items():
        out = copy.copy(in_var[dname])
        solver.vertex_pipelines[dname].operation_order(out)  # graph.py:170-180
        allvars[dname] = {k: v.detach().numpy().copy() for k, v in out.items()}
        pred[dname] = out.subset(names)
    loss = solver.compute_loss(in_var, pred, true_out, lambda_out)
    loss.backward()
    comps = {k: float(v) for k, v in solver.loss_component.items()}
    grads = {}
    for nn in solver.netnodes:
        if nn.is_reference:
            continue
        for pname, p in nn.net.named_parameters():
            grads[f"{nn.name}/{pname}"] = None if p.grad is None else p.grad.detach().numpy().copy()
    return float(loss), comps, allvars, grads


def build(name):
    np.random.seed(0)
    torch.manual_seed(0)
    torch.set_default_dtype(torch.float32)
    domains, nets, pdes = CONFIGS[name]()
    solver = sc.Solver(sample_domains=domains, netnodes=nets, pdes=pdes, max_iter=1,
                       loading=False, network_dir=tempfile.mkdtemp(prefix="net_"))
    # one fixed draw of every domain, rounded to FP32 (what the reference feeds the net)
    samples_np = {}
    for dn in solver.sample_domains:
        s = dn.sample()
        samples_np[dn.name] = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                               .astype(np.float32).reshape(-1, 1) for k, v in s.items()}

    # ---- problem in oracle format
    exprs = {}
    for pde in solver.pdes:
        for sub in pde.sub_nodes:
            exprs[sub.name] = str(sub.evaluate.tf_eq)
    prob_nets = []
    for nn in solver.netnodes:
        entry = {"name": nn.name, "inputs": list(nn.inputs), "outputs": list(nn.outputs),
                 "act": _act_name(nn.net), "shares": None}
        if nn.is_reference:
            entry["shares"] = next(m.name for m in solver.netnodes if m.net is nn.net and not m.is_reference)
            entry["layers"] = []
        else:
            entry["layers"] = _net_layers(nn.net)
        prob_nets.append(entry)
    prob_domains = []
    for dn in solver.sample_domains:
        s = samples_np[dn.name]
        pts = {k: s[k] for k in dn.inputs}
        tg = {k: s[k] for k in dn.outputs}
        lam = {k[len("lambda_"):]: s[k] for k in dn.lambda_outputs}
        prob_domains.append({"name": dn.name, "points": pts, "targets": tg, "lambdas": lam,
                             "loss": dn.loss_fn, "sigma": float(dn.sigma)})
    problem = {"nets": prob_nets, "exprs": exprs, "domains": prob_domains}
    flat = problem_to_flat(problem)

    # ---- run the reference: FP32 then FP64
    for tag, dtype in (("ref32", torch.float32), ("ref64", torch.float64)):
        torch.set_default_dtype(dtype)
        loss, comps, allvars, grads = _strict_pass(solver, samples_np, dtype)
        flat[f"{tag}/loss"] = np.array(loss, dtype=np.float64)
        flat[f"{tag}/loss_components"] = np.array(json.dumps(comps))
        for dname, d in allvars.items():
            for k, a in d.items():
                flat[f"{tag}/dom/{dname}/{k}"] = a
        for k, a in grads.items():
            if a is not None:
                if name in COMPACT:
                    if tag == "ref32":
                        continue
                    a = a.astype(np.float32)
                flat[f"{tag}/grad/{k}"] = a
        flat[f"{tag}/grad_none"] = np.array(json.dumps(sorted(k for k, a in grads.items() if a is None)))
    torch.set_default_dtype(torch.float32)
    out = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(out, **flat)
    n_pts = {d["name"]: len(next(iter(d["points"].values()))) for d in prob_domains}
    print(f"{name}: points {n_pts}  loss32 {float(flat['ref32/loss']):.8e}  loss64 {float(flat['ref64/loss']):.8e}"
          f"  -> {os.path.getsize(out) / 1024:.0f} KiB")


if __name__ == "__main__":
    for cfg in (sys.argv[1:] or [c for c in CONFIGS if c not in COMPACT]):
        build(cfg)
