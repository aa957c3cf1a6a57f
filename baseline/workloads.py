"""The five BASELINE.json workloads as synthetic, pre-sampled problems, written once against
the reference's public API so that the SAME definition drives

  * this package   (`sc` = idrlnet_b200.shortcut; bench.py, the GPU arm), and
  * the unmodified reference (`sc` = idrlnet.shortcut from baseline/_ref; baseline/ref_harness.py:
    the reference's own PyTorch path on CUDA or on the host cores).

Every domain holds tensor points AND tensor targets, the way examples/ldc/ldc.py:13-60 does (the reference
cannot mix tensor points with python-scalar targets, idrlnet/data.py:89).  A `PointSets` object owns the device
columns; the sampling domains just hand out views of the set that is current, so a "step" never touches the host
unless the caller copies new coordinates in.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

F32 = np.float32
NU = 0.01 / math.pi


def _u(rng, lo, hi, n):
    return rng.uniform(lo, hi, n).astype(F32)


def _pm(rng, n, a=-1.0, b=1.0):
    return np.where(rng.uniform(size=n) < 0.5, a, b).astype(F32)


# ------------------------------------------------------------------ samplers: {column: float32 [n]}
def sample_burgers(rng, n):
    ic_x = _u(rng, -1, 1, n["ic"])
    return {"int_x": _u(rng, -1, 1, n["int"]), "int_t": _u(rng, 0, 1, n["int"]),
            "ic_x": ic_x, "ic_t": np.zeros(n["ic"], F32), "ic_u": (-np.sin(np.pi * ic_x)).astype(F32),
            "bc_x": _pm(rng, n["bc"]), "bc_t": _u(rng, 0, 1, n["bc"])}


def sample_poisson(rng, n):
    nb = n["bc"]
    side2 = _pm(rng, nb)
    return {"int_x": _u(rng, -1, 1, n["int"]), "int_y": _u(rng, -1, 1, n["int"]),
            "lr_x": _pm(rng, nb), "lr_y": _u(rng, -1, 1, nb),
            "ud_x": _u(rng, -1, 1, nb), "ud_y": side2, "ud_nx": np.zeros(nb, F32), "ud_ny": side2.copy()}


def sample_ldc(rng, n):
    ix, iy = _u(rng, -0.05, 0.05, n["int"]), _u(rng, -0.05, 0.05, n["int"])
    nw = n["wall"]
    w = rng.integers(0, 3, nw)
    s = _u(rng, -0.05, 0.05, nw)
    lx = _u(rng, -0.05, 0.05, n["lid"])
    return {"int_x": ix, "int_y": iy, "int_sdf": np.minimum(0.05 - np.abs(ix), 0.05 - np.abs(iy)).astype(F32),
            "wall_x": np.where(w == 0, -0.05, np.where(w == 1, 0.05, s)).astype(F32),
            "wall_y": np.where(w == 2, -0.05, s).astype(F32),
            "lid_x": lx, "lid_y": np.full(n["lid"], 0.05, F32), "lid_lam": (1 - 20 * np.abs(lx)).astype(F32)}


def sample_allen(rng, n):
    ic_x = _u(rng, -1, 1, n["ic"])
    return {"int_x": _u(rng, -1, 1, n["int"]), "int_t": _u(rng, 0, 1, n["int"]),
            "ic_x": ic_x, "ic_t": np.zeros(n["ic"], F32), "ic_u": (ic_x ** 2 * np.cos(np.pi * ic_x)).astype(F32),
            "bc_x": np.full(n["bc"], -1.0, F32), "bc_t": _u(rng, 0, 1, n["bc"]),
            "rs_x": _u(rng, -1, 1, n["rs"]), "rs_t": _u(rng, 0, 1, n["rs"])}


def sample_ns(rng, n):
    nd = n["data"]
    return {"int_x": _u(rng, 1, 8, n["int"]), "int_y": _u(rng, -2, 2, n["int"]), "int_t": _u(rng, 0, 20, n["int"]),
            "d_x": _u(rng, 1, 8, nd), "d_y": _u(rng, -2, 2, nd), "d_t": _u(rng, 0, 20, nd),
            "d_u": _u(rng, -1, 1, nd), "d_v": _u(rng, -1, 1, nd), "d_p": _u(rng, -1, 1, nd)}


def flop_per_point(d, H, L, C, pairs):
    """F_step = 3 * 2 * [d*H + C*(L-1)*H^2 + n_pairs*H]."""
    return 3 * 2 * (d * H + C * (L - 1) * H * H + pairs * H)


# ------------------------------------------------------------------ nets and PDE nodes, over a shortcut module
def _net_mlp(inputs, outputs, name="net1"):
    return lambda sc: [sc.get_net_node(inputs=inputs, outputs=outputs, name=name, arch=sc.Arch.mlp)]


def _net_ldc(sc):
    return [sc.NetNode(inputs=("x", "y"), outputs=("u", "v", "p"), name="net_u",
                       net=sc.MLP([2, 100, 100, 100, 100, 3], activation=sc.Activation.tanh,
                                  initialization=sc.Initializer.Xavier_uniform, weight_norm=False))]


def _net_allen(sc):
    """examples/allen_cahn/allen_cahn.py:94-104: the net and its weight-sharing copy on the shifted input."""
    net_u = sc.NetNode(inputs=("x", "t"), outputs=("u",), name="net1",
                       net=sc.MLP([2, 128, 128, 128, 128, 2], activation=sc.Activation.tanh))
    tilde = sc.get_shared_net_node(net_u, inputs=("xp", "t"), outputs=("up",), name="net2", arch="mlp")
    return [net_u, tilde]


def _pdes_allen(sc):
    """examples/allen_cahn/allen_cahn.py:105-118: residual + the periodic boundary condition u(-1,t) = u(1,t),
    u_x(-1,t) = u_x(1,t) written with ExpressionNodes and Difference nodes over the shared net."""
    import sympy as sp
    x, t = sp.Symbol("x"), sp.Symbol("t")
    u, up = sp.Function("u")(x, t), sp.Function("up")(x, t)
    return [sc.AllenCahnNode(u="u", gamma_1=0.0001, gamma_2=5),
            sc.ExpressionNode(name="xp", expression=x + 2),
            sc.ExpressionNode(expression=up.diff(x), name="diff_up"),
            sc.ExpressionNode(expression=u.diff(x), name="diff_u"),
            sc.Difference(T="diff_u", S="diff_up"),
            sc.Difference(T="u", S="up")]


def _net_xl(sc):
    return [sc.get_net_node(inputs=("x", "y", "t"), outputs=("u", "v", "p"), name="net", arch=sc.Arch.mlp_xl)]


ZERO = 0.0  # targets written as numbers below become [n,1] tensor columns (see PointSets)

# domain rows: (name, size key, {symbol: column}, {target: column | number}, {lambda target: column | number},
#               area(sizes) -> float, loss kind)
CONFIGS: Dict[str, dict] = {
    "burgers": dict(
        title="burgers_equation (BASELINE configs[1])", net_desc="Arch.mlp 2-20x4-1 swish weight-norm",
        nets=_net_mlp(("x", "t"), ("u",)), pdes=lambda sc: [sc.BurgersNode(u="u", v=NU)], sample=sample_burgers,
        sizes=lambda world: {"int": 1 << 20, "ic": 10240, "bc": 10240}, scaling="weak",
        domains=[
            ("t_boundary", "ic", {"x": "ic_x", "t": "ic_t"}, {"u": "ic_u"}, {}, lambda n: 2.0 / n["ic"], "square"),
            ("burgers_equation", "int", {"x": "int_x", "t": "int_t"}, {"burgers_u": ZERO}, {}, lambda n: 2.0 / n["int"], "square"),
            ("x_boundary", "bc", {"x": "bc_x", "t": "bc_t"}, {"u": ZERO}, {}, lambda n: 2.0 / n["bc"], "square"),
        ], timed="burgers_equation", flop_per_point=flop_per_point(2, 20, 4, 4, 4), bytes_per_point=16,
        kernel="narrow_kernel<NarrowWarpPair<20,2,1,silu>> (interior domain; channel pairs, packed FMAs)", wide=False),
    "poisson": dict(
        title="simple_poisson (BASELINE configs[0])", net_desc="Arch.mlp 2-20x4-1 swish weight-norm",
        nets=_net_mlp(("x", "y"), ("T",)),
        pdes=lambda sc: [sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False), sc.NormalGradient("T", dim=2, time=False)],
        sample=sample_poisson, sizes=lambda world: {"int": 1 << 20, "bc": 1 << 16}, scaling="weak",
        domains=[
            ("LeftRight", "bc", {"x": "lr_x", "y": "lr_y"}, {"T": ZERO}, {}, lambda n: 4.0 / n["bc"], "square"),
            ("heat_domain", "int", {"x": "int_x", "y": "int_y"}, {"diffusion_T": 1.0}, {}, lambda n: 4.0 / n["int"], "square"),
            ("up_down", "bc", {"x": "ud_x", "y": "ud_y", "normal_x": "ud_nx", "normal_y": "ud_ny"},
             {"normal_gradient_T": ZERO}, {}, lambda n: 4.0 / n["bc"], "square"),
        ], timed="heat_domain", flop_per_point=flop_per_point(2, 20, 4, 5, 5), bytes_per_point=16,
        kernel="narrow_kernel<NarrowWarp<20,2,2,silu>> (interior domain)", wide=False),
    "ldc": dict(
        title="ldc (BASELINE configs[3]: 16 M interior points across 8 GPUs = 2^21 per GPU)",
        net_desc="MLP 2-100x4-3 tanh Xavier", nets=_net_ldc,
        pdes=lambda sc: [sc.NavierStokesNode(nu=0.01, rho=1.0, dim=2, time=False)], sample=sample_ldc,
        sizes=lambda world: {"int": 1 << 21, "wall": 1 << 15, "lid": 1 << 13}, scaling="weak",
        domains=[
            ("left_right_down", "wall", {"x": "wall_x", "y": "wall_y"}, {"u": ZERO, "v": ZERO}, {}, lambda n: 1e-4, "square"),
            ("flow_domain", "int", {"x": "int_x", "y": "int_y"}, {"continuity": ZERO, "momentum_x": ZERO, "momentum_y": ZERO},
             {"continuity": "int_sdf", "momentum_x": "int_sdf", "momentum_y": "int_sdf"}, lambda n: 0.01 / n["int"], "square"),
            ("up", "lid", {"x": "lid_x", "y": "lid_y"}, {"u": 1.0, "v": ZERO}, {"u": "lid_lam"}, lambda n: 1e-4, "square"),
        ], timed="flow_domain", flop_per_point=flop_per_point(2, 100, 4, 5, 12), bytes_per_point=28,
        kernel="wide path (interior domain)", wide=True),
    "allen_cahn": dict(
        title="allen_cahn (BASELINE configs[2]: 4 M interior points; periodic BC over the shared net and the "
              "re-sampling domain included)",
        net_desc="MLP 2-128x4-2 tanh weight-norm + shared copy on (x+2, t)", nets=_net_allen, pdes=_pdes_allen,
        sample=sample_allen, sizes=lambda world: {"int": (1 << 22) // world, "ic": 1 << 14, "bc": 1 << 14, "rs": 1 << 14},
        scaling="strong",
        domains=[
            ("AllenInit", "ic", {"x": "ic_x", "t": "ic_t"}, {"u": "ic_u"}, {"u": 100.0}, lambda n: 2.0 / n["ic"], "square"),
            ("AllenBc", "bc", {"x": "bc_x", "t": "bc_t"}, {"difference_u_up": ZERO, "difference_diff_u_diff_up": ZERO}, {},
             lambda n: 1.0 / n["bc"], "square"),
            ("allen_domain", "int", {"x": "int_x", "t": "int_t"}, {"AllenCahn_u": ZERO}, {}, lambda n: 2.0 / n["int"], "square"),
            ("re_sampling_domain", "rs", {"x": "rs_x", "t": "rs_t"}, {"AllenCahn_u": ZERO}, {}, lambda n: 2.0 / n["rs"], "square"),
        ], timed="allen_domain", flop_per_point=flop_per_point(2, 128, 4, 4, 4), bytes_per_point=16,
        kernel="wide path (interior domain)", wide=True),
    "ns_xl": dict(
        title="ns_unsteady with Arch.mlp_xl (BASELINE configs[4]: 64 M interior points at 2/4/8 GPUs)",
        net_desc="Arch.mlp_xl 3-512x6-3 SiLU weight-norm", nets=_net_xl,
        pdes=lambda sc: [sc.NavierStokesNode(nu=0.01, rho=1.0, dim=2, time=True)], sample=sample_ns,
        sizes=lambda world: {"int": (1 << 20) if world == 1 else (1 << 26) // world, "data": 5000}, scaling="strong",
        domains=[
            ("NS_domain", "data", {"x": "d_x", "y": "d_y", "t": "d_t"}, {"u": "d_u", "v": "d_v", "p": "d_p"}, {},
             lambda n: 1.0 / n["data"], "L1"),
            ("NS_external", "int", {"x": "int_x", "y": "int_y", "t": "int_t"},
             {"continuity": ZERO, "momentum_x": ZERO, "momentum_y": ZERO}, {}, lambda n: 28.0 / n["int"], "square"),
        ], timed="NS_external", flop_per_point=flop_per_point(3, 512, 6, 6, 14), bytes_per_point=28,
        kernel="wide streamed path (interior domain)", wide=True),
}


def scaled_sizes(sizes: Dict[str, int], factor: float) -> Dict[str, int]:
    return {k: max(32, int(v * factor)) for k, v in sizes.items()}


def points_per_step(cfg: dict, sizes: Dict[str, int]) -> int:
    return sum(sizes[d[1]] for d in cfg["domains"])


def coordinate_columns(cfg: dict) -> List[str]:
    """Columns that change from batch to batch (what an end-to-end step copies from the host): the coordinates and
    the per-point lambda columns derived from them; targets that are functions of the coordinates ride along."""
    cols: List[str] = []
    for d in cfg["domains"]:
        for c in list(d[2].values()) + [v for v in d[3].values() if isinstance(v, str)] + [v for v in d[4].values() if isinstance(v, str)]:
            if c not in cols:
                cols.append(c)
    return cols


class PointSets:
    """Device-resident point batches of one workload.  `sets[i]` maps column -> [n,1] tensor; constant columns
    (area, numeric targets / lambdas) are created once and shared by every set."""

    def __init__(self, cfg: dict, sizes: Dict[str, int], device, torch, dtype=None):
        self.cfg, self.sizes, self.device, self.torch = cfg, dict(sizes), device, torch
        self.dtype = dtype or torch.float32
        self.sets: List[Dict[str, "object"]] = []
        self.current = 0
        self._const: Dict[Tuple[str, float], "object"] = {}
        self.columns = coordinate_columns(cfg)

    def const(self, size_key: str, value: float):
        key = (size_key, float(value))
        if key not in self._const:
            self._const[key] = self.torch.full((self.sizes[size_key], 1), float(value), dtype=self.dtype, device=self.device)
        return self._const[key]

    def add_host_batch(self, batch: Dict[str, np.ndarray]) -> int:
        t = self.torch
        self.sets.append({k: t.as_tensor(batch[k].reshape(-1, 1), dtype=self.dtype).to(self.device) for k in self.columns})
        return len(self.sets) - 1

    def add_empty(self) -> int:
        t = self.torch
        ncol = {c: self.sizes[d[1]] for d in self.cfg["domains"] for c in
                list(d[2].values()) + [v for v in d[3].values() if isinstance(v, str)] + [v for v in d[4].values() if isinstance(v, str)]}
        self.sets.append({k: t.empty((ncol[k], 1), dtype=self.dtype, device=self.device) for k in self.columns})
        return len(self.sets) - 1

    def provider(self, row, requires_grad: bool) -> Callable[[], Tuple[dict, dict]]:
        name, skey, colmap, targets, lambdas, area = row[:6]

        def sample():
            cur = self.sets[self.current]
            pts = {}
            for sym, c in colmap.items():
                v = cur[c]
                if requires_grad and not v.requires_grad:
                    v.requires_grad_()  # coordinates are autograd leaves for the reference (idrlnet/variable.py:109-112)
                pts[sym] = v
            pts["area"] = self.const(skey, area(self.sizes))
            cons = {k: (cur[v] if isinstance(v, str) else self.const(skey, v)) for k, v in targets.items()}
            for k, v in lambdas.items():
                cons["lambda_" + k] = cur[v] if isinstance(v, str) else self.const(skey, v)
            return pts, cons
        return sample


def build_problem(sc, cfg: dict, sets: PointSets, requires_grad: bool):
    """(sample_domains, netnodes, pdes) through the shortcut module `sc` (this package's or the reference's)."""
    domains = []
    for row in cfg["domains"]:
        fn = sets.provider(row, requires_grad)

        def make(fn=fn, row=row):
            @sc.datanode(name=row[0], loss_fn=row[6])
            class Domain(sc.SampleDomain):
                def sampling(self, *args, **kwargs):
                    return fn()
            return Domain()
        domains.append(make())
    return tuple(domains), cfg["nets"](sc), cfg["pdes"](sc)
