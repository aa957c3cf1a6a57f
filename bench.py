#!/usr/bin/env python
"""Headline benchmark: collocation points / second per PINN training step
(forward jets + residual + loss + backward + gradient reduction + Adam) on the
Burgers configuration of BASELINE.json (configs[1]: examples/burgers_equation,
default MLP 2-20x4-1 swish + weight-norm, 2^20 interior + 10240 IC + 10240 BC
synthetic points per step per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (see the driver contract in the task statement).
`--impl reference` times the reference's own algorithm (the CPU oracle port,
oracle/pinn_oracle.py: torch autograd on the host cores) on a bounded sample of
the same workload.
"""
import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "collocation points/sec per train step (fwd+derivs+residual+bwd+optimizer)"
NU = 0.01 / math.pi
N_ROTATE = 32  # distinct point batches; 32 * 8.4 MB = 270 MB > 126 MB L2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak_tf32():
    """Dense TF32 tensor throughput: half the BF16 cuBLAS figure (sustained: the kernels are timed inside a step)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        if "bf16_tflops_sustained" in d:
            return d["bf16_tflops_sustained"] / 2, "TF32 = 1/2 of the measured sustained BF16 cuBLAS rate (MEASURED_PEAKS.json)"
    return 2250.0 / 2 / 2, "fallback: 1/2 of half the nominal dense BF16 peak (B200_PROFILING.md)"


# --------------------------------------------------------------------------- #
# synthetic workloads; "burgers" is the headline (BASELINE configs[1])
# --------------------------------------------------------------------------- #
F32 = np.float32


def _u(rng, lo, hi, n):
    return rng.uniform(lo, hi, n).astype(F32)


def _sample_burgers(rng, n):
    ic_x = _u(rng, -1, 1, n["ic"])
    return {"int_x": _u(rng, -1, 1, n["int"]), "int_t": _u(rng, 0, 1, n["int"]),
            "ic_x": ic_x, "ic_t": np.zeros(n["ic"], F32), "ic_u": (-np.sin(np.pi * ic_x)).astype(F32),
            "bc_x": np.where(rng.uniform(size=n["bc"]) < 0.5, -1.0, 1.0).astype(F32), "bc_t": _u(rng, 0, 1, n["bc"])}


def _sample_poisson(rng, n):
    nb = n["bc"]
    side = np.where(rng.uniform(size=nb) < 0.5, -1.0, 1.0).astype(F32)
    side2 = np.where(rng.uniform(size=nb) < 0.5, -1.0, 1.0).astype(F32)
    return {"int_x": _u(rng, -1, 1, n["int"]), "int_y": _u(rng, -1, 1, n["int"]),
            "lr_x": side, "lr_y": _u(rng, -1, 1, nb),
            "ud_x": _u(rng, -1, 1, nb), "ud_y": side2, "ud_nx": np.zeros(nb, F32), "ud_ny": side2.copy()}


def _sample_ldc(rng, n):
    ix, iy = _u(rng, -0.05, 0.05, n["int"]), _u(rng, -0.05, 0.05, n["int"])
    sdf = np.minimum(0.05 - np.abs(ix), 0.05 - np.abs(iy)).astype(F32)
    nw = n["wall"]
    w = rng.integers(0, 3, nw)
    s = _u(rng, -0.05, 0.05, nw)
    wx = np.where(w == 0, -0.05, np.where(w == 1, 0.05, s)).astype(F32)
    wy = np.where(w == 2, -0.05, s).astype(F32)
    lx = _u(rng, -0.05, 0.05, n["lid"])
    return {"int_x": ix, "int_y": iy, "int_sdf": sdf, "wall_x": wx, "wall_y": wy,
            "lid_x": lx, "lid_y": np.full(n["lid"], 0.05, F32), "lid_lam": (1 - 20 * np.abs(lx)).astype(F32)}


def _sample_allen(rng, n):
    ic_x = _u(rng, -1, 1, n["ic"])
    return {"int_x": _u(rng, -1, 1, n["int"]), "int_t": _u(rng, 0, 1, n["int"]),
            "ic_x": ic_x, "ic_t": np.zeros(n["ic"], F32), "ic_u": (ic_x ** 2 * np.cos(np.pi * ic_x)).astype(F32)}


def _sample_ns(rng, n):
    nd = n["data"]
    return {"int_x": _u(rng, 1, 8, n["int"]), "int_y": _u(rng, -2, 2, n["int"]), "int_t": _u(rng, 0, 20, n["int"]),
            "d_x": _u(rng, 1, 8, nd), "d_y": _u(rng, -2, 2, nd), "d_t": _u(rng, 0, 20, nd),
            "d_u": _u(rng, -1, 1, nd), "d_v": _u(rng, -1, 1, nd), "d_p": _u(rng, -1, 1, nd)}


def _flop(d, H, L, C, pairs):  # F_step = 3 * 2 * [d*H + C*(L-1)*H^2 + n_pairs*H]
    return 3 * 2 * (d * H + C * (L - 1) * H * H + pairs * H)


def workload_configs():
    import idrlnet_b200.shortcut as sc
    return {
        "burgers": dict(
            title="burgers_equation (BASELINE configs[1])", net_desc="Arch.mlp 2-20x4-1 swish weight-norm",
            inputs=("x", "t"), outputs=("u",), oracle_act="silu",
            net=lambda: sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp),
            pdes=lambda: [sc.BurgersNode(u="u", v=NU)],
            sizes={"int": 1 << 20, "ic": 10240, "bc": 10240}, sample=_sample_burgers,
            domains=[  # (name, size key, columns {symbol: column}, targets, lambdas, area)
                ("t_boundary", "ic", {"x": "ic_x", "t": "ic_t"}, {"u": "ic_u"}, {}, lambda n: 2.0 / n["ic"]),
                ("burgers_equation", "int", {"x": "int_x", "t": "int_t"}, {"burgers_u": 0.0}, {}, lambda n: 2.0 / n["int"]),
                ("x_boundary", "bc", {"x": "bc_x", "t": "bc_t"}, {"u": 0.0}, {}, lambda n: 2.0 / n["bc"]),
            ], timed=1, flop_per_point=_flop(2, 20, 4, 4, 4), bytes_per_point=8,
            kernel="narrow_kernel<NarrowWarpPair<20,2,1,silu>> (interior domain; channel pairs, packed FMAs)", bound="fp32"),
        "poisson": dict(
            title="simple_poisson (BASELINE configs[0])", net_desc="Arch.mlp 2-20x4-1 swish weight-norm",
            inputs=("x", "y"), outputs=("T",), oracle_act="silu",
            net=lambda: sc.get_net_node(inputs=("x", "y"), outputs=("T",), name="net1", arch=sc.Arch.mlp),
            pdes=lambda: [sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False), sc.NormalGradient("T", dim=2, time=False)],
            sizes={"int": 1 << 20, "bc": 1 << 16}, sample=_sample_poisson,
            domains=[
                ("LeftRight", "bc", {"x": "lr_x", "y": "lr_y"}, {"T": 0.0}, {}, lambda n: 4.0 / n["bc"]),
                ("heat_domain", "int", {"x": "int_x", "y": "int_y"}, {"diffusion_T": 1.0}, {}, lambda n: 4.0 / n["int"]),
                ("up_down", "bc", {"x": "ud_x", "y": "ud_y", "normal_x": "ud_nx", "normal_y": "ud_ny"},
                 {"normal_gradient_T": 0.0}, {}, lambda n: 4.0 / n["bc"]),
            ], timed=1, flop_per_point=_flop(2, 20, 4, 5, 5), bytes_per_point=8,
            kernel="narrow_kernel<20,2,2,silu> (interior domain)", bound="fp32"),
        "ldc": dict(
            title="ldc (BASELINE configs[3])", net_desc="MLP 2-100x4-3 tanh Xavier",
            inputs=("x", "y"), outputs=("u", "v", "p"), oracle_act="tanh",
            net=lambda: sc.NetNode(inputs=("x", "y"), outputs=("u", "v", "p"), name="net_u",
                                   net=sc.MLP([2, 100, 100, 100, 100, 3], activation=sc.Activation.tanh,
                                              initialization=sc.Initializer.Xavier_uniform, weight_norm=False)),
            pdes=lambda: [sc.NavierStokesNode(nu=0.01, rho=1.0, dim=2, time=False)],
            sizes={"int": 1 << 18, "wall": 1 << 14, "lid": 1 << 12}, sample=_sample_ldc,
            domains=[
                ("left_right_down", "wall", {"x": "wall_x", "y": "wall_y"}, {"u": 0.0, "v": 0.0}, {}, lambda n: 1e-4),
                ("flow_domain", "int", {"x": "int_x", "y": "int_y"}, {"continuity": 0.0, "momentum_x": 0.0, "momentum_y": 0.0},
                 {"continuity": "int_sdf", "momentum_x": "int_sdf", "momentum_y": "int_sdf"}, lambda n: 0.01 / n["int"]),
                ("up", "lid", {"x": "lid_x", "y": "lid_y"}, {"u": 1.0, "v": 0.0}, {"u": "lid_lam"}, lambda n: 1e-4),
            ], timed=1, flop_per_point=_flop(2, 100, 4, 5, 12), bytes_per_point=12,
            kernel="wide gemm_rows/gemm_dw kernels (interior domain)", bound="fp32"),
        "allen_cahn": dict(
            title="allen_cahn (BASELINE configs[2])", net_desc="MLP 2-128x4-2 tanh weight-norm",
            inputs=("x", "t"), outputs=("u",), oracle_act="tanh",
            net=lambda: sc.NetNode(inputs=("x", "t"), outputs=("u",), name="net1",
                                   net=sc.MLP([2, 128, 128, 128, 128, 2], activation=sc.Activation.tanh)),
            pdes=lambda: [sc.AllenCahnNode(u="u", gamma_1=0.0001, gamma_2=5)],
            sizes={"int": 1 << 18, "ic": 1 << 14}, sample=_sample_allen,
            domains=[
                ("AllenInit", "ic", {"x": "ic_x", "t": "ic_t"}, {"u": "ic_u"}, {"u": 100.0}, lambda n: 2.0 / n["ic"]),
                ("allen_domain", "int", {"x": "int_x", "t": "int_t"}, {"AllenCahn_u": 0.0}, {}, lambda n: 2.0 / n["int"]),
            ], timed=1, flop_per_point=_flop(2, 128, 4, 4, 4), bytes_per_point=8,
            kernel="wide gemm_rows/gemm_dw kernels (interior domain)", bound="fp32"),
        "ns_xl": dict(
            title="ns_unsteady with Arch.mlp_xl (BASELINE configs[4])", net_desc="Arch.mlp_xl 3-512x6-3 SiLU weight-norm",
            inputs=("x", "y", "t"), outputs=("u", "v", "p"), oracle_act="silu",
            net=lambda: sc.get_net_node(inputs=("x", "y", "t"), outputs=("u", "v", "p"), name="net", arch=sc.Arch.mlp_xl),
            pdes=lambda: [sc.NavierStokesNode(nu=0.01, rho=1.0, dim=2, time=True)],
            sizes={"int": 1 << 15, "data": 5000}, sample=_sample_ns,
            domains=[
                ("NS_domain", "data", {"x": "d_x", "y": "d_y", "t": "d_t"}, {"u": "d_u", "v": "d_v", "p": "d_p"}, {},
                 lambda n: 1.0 / n["data"], "L1"),
                ("NS_external", "int", {"x": "int_x", "y": "int_y", "t": "int_t"},
                 {"continuity": 0.0, "momentum_x": 0.0, "momentum_y": 0.0}, {}, lambda n: 28.0 / n["int"]),
            ], timed=1, flop_per_point=_flop(3, 512, 6, 6, 14), bytes_per_point=12,
            kernel="wide gemm_rows/gemm_dw kernels (interior domain)", bound="fp32"),
    }


class Trainer:
    """A workload wired through the package's public pieces: get_net_node / MLP,
    the PDE nodes, the residual compiler and the fused StepEngine + torch Adam."""

    def __init__(self, cfg, device, world=1, sizes=None, precision="fp32"):
        from idrlnet_b200 import compiler as cp
        from idrlnet_b200 import fused
        self.cfg, self.device, self.world, self.fused = cfg, device, world, fused
        self.sizes = dict(cfg["sizes"] if sizes is None else sizes)
        torch.manual_seed(0)
        self.net = cfg["net"]()
        self.net.net.to(device)
        exprs = {}
        for pde in cfg["pdes"]():
            for sn in pde.sub_nodes:
                exprs[sn.name] = sn.evaluate.tf_eq
        self.exprs = exprs
        spec = [cp.NetSpec(self.net.name, tuple(cfg["inputs"]), tuple(cfg["outputs"]))]
        self.compiled = []
        for dom in cfg["domains"]:
            name, _, cols, targets = dom[0], dom[1], dom[2], dom[3]
            self.compiled.append(cp.compile_domain(list(targets), exprs, spec, list(cols) + ["area"]))
        self.pack = fused.NetPack(self.net.net, precision=precision)
        self.opt = torch.optim.Adam(list(self.net.net.parameters()), lr=1e-3, fused=True)
        self.sched = torch.optim.lr_scheduler.ExponentialLR(self.opt, gamma=math.pow(0.95, 0.001))
        self.engine, self.sets = None, []
        self.points_per_step = sum(self.sizes[d[1]] for d in cfg["domains"])
        self.batch_keys = None

    def sample_host(self, rng):
        b = self.cfg["sample"](rng, self.sizes)
        if self.batch_keys is None:
            self.batch_keys = tuple(b)
        return b

    def bind_set(self, cols):
        doms = []
        for dom, cd in zip(self.cfg["domains"], self.compiled):
            name, _, colmap, targets, lambdas, area = dom[:6]
            loss = dom[6] if len(dom) > 6 else "square"
            fd = self.fused.FusedDomain(name, cd, loss)
            val = lambda v: cols[v] if isinstance(v, str) else v
            fd.bind({k: cols[c] for k, c in colmap.items()}, {k: val(v) for k, v in targets.items()},
                    {k: val(v) for k, v in lambdas.items()}, area(self.sizes))
            doms.append(fd)
        if self.engine is None:
            self.engine = self.fused.StepEngine(self.pack, doms)
        self.sets.append(doms)
        return len(self.sets) - 1

    def step(self, set_idx, ev=None):
        """One full training step on the current stream; returns the loss (0-d device tensor)."""
        eng = self.engine
        eng.domains = self.sets[set_idx]
        eng.launch(timed=None if ev is None else (self.cfg["timed"], ev[0], ev[1]))
        if self.world > 1:
            torch.distributed.all_reduce(eng.flat)
        self.pack.scatter_grad(eng.grad)
        loss = eng.total_loss()
        self.opt.step()
        self.sched.step()
        return loss


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- #
# Reference-algorithm baselines: the oracle (torch autograd restatement of the
# reference, oracle/pinn_oracle.py) on the host cores, or -- for the "reference's
# own PyTorch CUDA path" comparison of BASELINE.md section 3 -- on the GPU.
# --------------------------------------------------------------------------- #
def net_layers_numpy(net_module):
    out = []
    for k, l in net_module.layers.items():
        if k.endswith("_activation"):
            continue
        if hasattr(l, "weight_g"):
            out.append({"g": l.weight_g.detach().cpu().numpy().copy(), "v": l.weight_v.detach().cpu().numpy().copy(),
                        "b": l.bias.detach().cpu().numpy().copy()})
        else:
            out.append({"W": l.weight.detach().cpu().numpy().copy(), "b": l.bias.detach().cpu().numpy().copy()})
    return out


def oracle_problem(cfg, sizes, rng):
    """The workload as an oracle-format problem (see oracle/pinn_oracle.py)."""
    torch.manual_seed(0)
    node = cfg["net"]()
    exprs = {}
    for pde in cfg["pdes"]():
        for sn in pde.sub_nodes:
            exprs[sn.name] = str(sn.evaluate.tf_eq)
    b = cfg["sample"](rng, sizes)
    domains = []
    for dom in cfg["domains"]:
        name, skey, colmap, targets, lambdas, area = dom[:6]
        n = sizes[skey]
        pts = {k: b[c].reshape(-1, 1) for k, c in colmap.items()}
        pts["area"] = np.full((n, 1), area(sizes), np.float32)
        val = lambda v: b[v].reshape(-1, 1) if isinstance(v, str) else float(v)
        domains.append({"name": name, "points": pts, "targets": {k: val(v) for k, v in targets.items()},
                        "lambdas": {k: val(v) for k, v in lambdas.items()},
                        "loss": dom[6] if len(dom) > 6 else "square", "sigma": 1.0})
    return {"nets": [{"name": node.name, "inputs": list(cfg["inputs"]), "outputs": list(cfg["outputs"]),
                      "act": cfg["oracle_act"], "layers": net_layers_numpy(node.net), "shares": None}],
            "exprs": exprs, "domains": domains}


def time_oracle(cfg, sizes, steps, warmup, device="cpu"):
    """Reference algorithm: forward graph (one autograd.grad per derivative symbol) +
    loss + backward + Adam, all domains, on `device`."""
    from oracle import pinn_oracle as po
    rng = np.random.default_rng(0)
    problem = oracle_problem(cfg, sizes, rng)
    dev = torch.device(device)
    nets = po.build_nets(problem)
    for lay in nets[0]["raw"]:
        for k in list(lay):
            lay[k] = lay[k].detach().to(dev).requires_grad_()
    leaves = [t for lay in nets[0]["raw"] for t in lay.values()]
    opt = torch.optim.Adam(leaves, lr=1e-3)
    # points resident on the device before timing
    for dom in problem["domains"]:
        dom["points"] = {k: torch.as_tensor(v, device=dev) for k, v in dom["points"].items()}
        dom["targets"] = {k: (torch.as_tensor(v, device=dev) if not np.isscalar(v) else v) for k, v in dom["targets"].items()}
        dom["lambdas"] = {k: (torch.as_tensor(v, device=dev) if not np.isscalar(v) else v) for k, v in dom["lambdas"].items()}

    def one_step():
        opt.zero_grad()
        eff = []
        for leaf in nets[0]["raw"]:
            eff.append(((torch._weight_norm(leaf["v"], leaf["g"], 0) if "g" in leaf else leaf["W"]), leaf["b"]))
        nets[0]["layers_t"] = eff
        total = 0.0
        for dom in problem["domains"]:
            var = po.forward_domain(problem, dom, nets)
            like = next(iter(var.values()))
            diff = {k: var[k] - po._as_t(t, torch.float32, like) for k, t in dom["targets"].items()}
            lam = {k: po._as_t(t, torch.float32, like) for k, t in dom["lambdas"].items()}
            total = total + po.weighted_loss(diff, lam, var["area"], dom["loss"])
        total.backward()
        opt.step()
        return total

    sync = (lambda: torch.cuda.synchronize()) if dev.type == "cuda" else (lambda: None)
    for _ in range(warmup):
        one_step()
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    sync()
    dt = (time.perf_counter() - t0) / steps
    pts = sum(sizes[d[1]] for d in cfg["domains"])
    return pts / dt, dt


def scaled_sizes(cfg, factor):
    return {k: max(32, int(v * factor)) for k, v in cfg["sizes"].items()}


CPU_SAMPLE = {"burgers": 1 / 8, "poisson": 1 / 8, "ldc": 1 / 32, "allen_cahn": 1 / 32, "ns_xl": 1 / 32}


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is entitled to every host
    thread this process may run on."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


# --------------------------------------------------------------------------- #
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cfg = workload_configs()[args.config]
    device = "cpu"
    if args.baseline_device == "cuda":
        device = f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"
    if args.baseline_tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    factor = args.sample if args.sample else (CPU_SAMPLE[args.config] if device == "cpu" else 1 / 4)
    sizes = scaled_sizes(cfg, factor)
    cores = use_all_host_threads()
    pps, dt = time_oracle(cfg, sizes, args.steps, args.warmup, device)
    sample = (f"{'+'.join(str(sizes[d[1]]) for d in cfg['domains'])} points/step ({factor:g} of the GPU workload), "
              f"{args.steps} steps, torch autograd {'TF32-allowed' if args.baseline_tf32 else 'FP32'} on {device}")
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{cfg['title']}: reference algorithm (torch autograd) on {device}, bounded sample",
                   "net": cfg["net_desc"], "points_per_step": sum(sizes[d[1]] for d in cfg["domains"])},
        "cpu_baseline": {"value": pps, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": pps, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def ncu_traffic(config):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed `ncu --set full` capture (profiles/ncu_traffic.json); None if not captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get(config, {}).get("dram_bytes_per_launch")
    return None


def measure_fp32_peak(device):
    from idrlnet_b200 import native as nv
    lib = nv.lib()
    scratch = torch.zeros(4, device=device)
    st = torch.cuda.current_stream().cuda_stream
    blocks = 148 * 8
    lib.pinn_fp32_fma_probe(2000, blocks, scratch.data_ptr(), st)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        flops = lib.pinn_fp32_fma_probe(20000, blocks, scratch.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def run_ours(args):
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the fused path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device(f"cuda:{local}")
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
    from idrlnet_b200 import native as nv

    cfg = workload_configs()[args.config]
    sizes = scaled_sizes(cfg, args.sample) if args.sample else None
    trainer = Trainer(cfg, device, world=world, sizes=sizes, precision=args.precision)
    wide_tc = args.precision != "fp32" and trainer.pack.width > 32
    rng = np.random.default_rng(1000 + rank)
    pts_per_step = trainer.points_per_step
    n_timed = trainer.sizes[cfg["domains"][cfg["timed"]][1]]
    # ---- device-resident rotating batches (value): together larger than L2 -------
    hb = trainer.sample_host(rng)
    batch_bytes = sum(v.nbytes for v in hb.values())
    n_rotate = max(2, min(N_ROTATE, int(math.ceil(270e6 / batch_bytes))))
    for i in range(n_rotate):
        if i:
            hb = trainer.sample_host(rng)
        trainer.bind_set({k: torch.as_tensor(v, device=device) for k, v in hb.items()})
    keys = trainer.batch_keys

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local)
    clocks.start()  # samples cover warm-up + timed region + e2e region (all under load)
    for i in range(args.warmup):
        trainer.step(i % n_rotate)
    sync_all()
    kernel_events = []
    launches0 = nv.lib().pinn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for i in range(args.steps):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        kernel_events.append(ev)
        loss = trainer.step((args.warmup + i) % n_rotate, ev)
    e1.record()
    sync_all()
    launches = nv.lib().pinn_launch_count() - launches0
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms_per_step = float(ms) / args.steps
    value = pts_per_step * world / (ms_per_step * 1e-3)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    final_loss = float(loss)

    # ---- end to end: host (pinned) points in, loss out, every step -----------
    n_host = 4
    host_sets = []
    for _ in range(n_host):
        hb = trainer.sample_host(rng)
        host_sets.append({k: torch.as_tensor(hb[k]).pin_memory() for k in keys})
    dev_bufs = [{k: torch.empty_like(host_sets[0][k], device=device) for k in keys} for _ in range(2)]
    set_ids = [trainer.bind_set(b) for b in dev_bufs]
    copy_stream = torch.cuda.Stream(device=device)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.zeros(args.steps + args.warmup + 2, dtype=torch.float32).pin_memory()
    h2d_bytes = sum(host_sets[0][k].numel() * 4 for k in keys)

    def upload(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for k in keys:
                dev_bufs[slot][k].copy_(host_sets[i % n_host][k], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_loop(n):
        main = torch.cuda.current_stream()
        upload(0)
        for i in range(n):
            slot = i % 2
            if i + 1 < n:
                upload(i + 1)  # overlaps with this step's compute
            main.wait_event(ready[slot])
            l = trainer.step(set_ids[slot])
            consumed[slot].record(main)
            loss_host[i].copy_(l, non_blocking=True)

    for sl in range(2):
        consumed[sl].record(torch.cuda.current_stream())
    e2e_loop(args.warmup)
    sync_all()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(args.steps)
    t1.record()
    sync_all()
    ms2 = torch.tensor([t0.elapsed_time(t1)], device=device)
    if world > 1:
        torch.distributed.all_reduce(ms2, op=torch.distributed.ReduceOp.MAX)
    e2e_value = pts_per_step * world / (float(ms2) / args.steps * 1e-3)
    clk = clocks.stop()

    if rank == 0:
        hbm_peak, peak_src = peaks()
        fp32_peak = measure_fp32_peak(device)
        achieved = cfg["flop_per_point"] * n_timed / (kern_ms * 1e-3) / 1e12
        if wide_tc:  # tensor-core kernels: the denominator is the TF32 tensor rate (a third of it under 3xTF32 splitting)
            roof_peak, roof_src = tensor_peak_tf32()
            if args.precision == "fp32x3":
                roof_peak, roof_src = roof_peak / 3, roof_src + ", / 3 MMAs per product (3xTF32)"
            roof_bound = "tensor"
            roof_kernel = ("on-chip tcgen05 kernel wide_fused (interior domain)" if args.precision == "tf32" and trainer.pack.width <= 128
                           else "tcgen05 tc_rows / tc_dw2 kernels (interior domain)")
        else:
            roof_peak, roof_src = fp32_peak, "FP32 FMA pipe, pinn_fp32_fma_probe measured in this run"
            roof_bound, roof_kernel = cfg["bound"], cfg["kernel"]
        hbm_achieved = cfg["bytes_per_point"] * n_timed / (kern_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            csz = scaled_sizes(cfg, CPU_SAMPLE[args.config])
            use_all_host_threads()
            pps, dt = time_oracle(cfg, csz, 8, 2, "cpu")
            cpu = {"value": pps, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{'+'.join(str(csz[d[1]]) for d in cfg['domains'])} points/step x 8 steps of the same "
                             f"workload (oracle: reference algorithm, torch autograd FP32)"}
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": ("f32 (3xTF32 split on tensor cores)" if args.precision == "fp32x3" else "tf32") if wide_tc else "f32", "data": "synthetic",
            "config": {"workload": f"{cfg['title']}: {' + '.join(str(trainer.sizes[d[1]]) for d in cfg['domains'])} "
                                   f"points per step per GPU ({', '.join(d[0] for d in cfg['domains'])})",
                       "net": f"{cfg['net_desc']}, Adam 1e-3 + ExponentialLR",
                       "points_per_step_per_gpu": pts_per_step, "parallelism": f"dp{world} (points sharded, 1 all-reduce)",
                       "l2": f"{n_rotate} rotating point batches ({n_rotate * batch_bytes / 1e6:.0f} MB > 126 MB L2)"},
            "roofline": {"bound": roof_bound, "achieved": achieved, "peak": roof_peak, "unit": "TFLOP/s",
                         "frac": achieved / roof_peak if roof_peak else None,
                         "traffic": None if wide_tc else ncu_traffic(args.config),
                         "kernel": roof_kernel, "kernel_ms": kern_ms, "flop_per_point": cfg["flop_per_point"],
                         "peak_source": roof_src, "fp32_fma_peak": fp32_peak,
                         "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": hbm_achieved / hbm_peak, "peak_source": peak_src}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches), "clocks": clk, "final_loss": final_loss,
        }
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="burgers", choices=["burgers", "poisson", "ldc", "allen_cahn", "ns_xl"],
                    help="workload; the driver's default (burgers) is the headline")
    ap.add_argument("--sample", type=float, default=None, help="scale every domain size by this factor")
    ap.add_argument("--baseline-device", default="cpu", choices=["cpu", "cuda"],
                    help="--impl reference: where the reference algorithm (torch autograd) runs")
    ap.add_argument("--baseline-tf32", action="store_true",
                    help="--impl reference --baseline-device cuda: allow TF32 matmuls in torch (the reference's pinned "
                         "torch 1.7.1 did so silently on Ampere; torch 2.x defaults to FP32)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="fp32", choices=["fp32", "tf32", "tf32_streamed", "fp32x3"],
                    help="layer GEMMs of wide nets: FP32 FMA (parity) or tcgen05 TF32; narrow nets are always FP32")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
