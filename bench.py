#!/usr/bin/env python
"""Headline benchmark: collocation points / second per PINN training step (forward jets + residual + loss +
backward + gradient reduction + Adam) on the Burgers configuration of BASELINE.json (configs[1]:
examples/burgers_equation, default MLP 2-20x4-1 swish + weight-norm, 2^20 interior + 10240 IC + 10240 BC synthetic
points per step per GPU), driven through the package's PUBLIC API: `Solver.train_pipe()` on sampling domains that
hand out device-resident (value) or freshly uploaded (e2e) points.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (driver contract).  Beside the headline it carries a `configs` object: the other BASELINE
configurations at their stated sizes (simple_poisson 2^20, ldc 2^21 per GPU = 16 M over 8, allen_cahn 2^22 with its
periodic-BC and re-sampling domains, ns_unsteady `mlp_xl` 2^20 at N=1 / 2^26 over N >= 2 GPUs), each with its
roofline and -- at N=1 -- the UNMODIFIED reference (baseline/_ref, `idrlnet.use_gpu(0)`) timed on the same B200 on
the same workload (baseline/ref_harness.py: strict pass of idrlnet/solver.py:329-339, and train_pipe as shipped).

`--impl reference` times the reference's own CPU implementation (the same unmodified package) on the host cores
on a bounded sample of the headline workload.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

from baseline import workloads as wl

METRIC = "collocation points/sec per train step (fwd+derivs+residual+bwd+optimizer)"
L2_BYTES = 126e6
HARNESS = os.path.join(ROOT, "baseline", "ref_harness.py")
# bounded CPU samples (fraction of the one-GPU workload) for the reference's CPU path
CPU_SAMPLE = {"burgers": 1 / 8, "poisson": 1 / 8, "ldc": 1 / 256, "allen_cahn": 1 / 128, "ns_xl": 1 / 512}
# wide-net arithmetic modes reported per config: (key suffix, precision)
WIDE_MODES = (("", "fp32x3"), ("_tf32", "tf32"))
# steps of the secondary configs (their steps are 50-500 ms long)
SECONDARY_STEPS = {"poisson": (5, 50), "ldc": (3, 10), "allen_cahn": (3, 8), "ns_xl": (3, 6)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 2250.0 / 2}, "fallback (B200_PROFILING.md)"


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
        return self

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


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is entitled to every host thread."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


# --------------------------------------------------------------------------- #
# the reference (unmodified idrlnet from baseline/_ref), in a subprocess
# --------------------------------------------------------------------------- #
def run_harness(config, device, steps, warmup, factor=1.0, pipe_steps=0, threads=0, sizes=None, tf32=False, gpu_index=0,
                timeout=1500):
    cmd = [sys.executable, HARNESS, "--config", config, "--device", device, "--steps", str(steps), "--warmup", str(warmup),
           "--factor", repr(factor), "--pipe-steps", str(pipe_steps), "--threads", str(threads), "--gpu-index", str(gpu_index)]
    if sizes:
        cmd += ["--sizes", ",".join(f"{k}={v}" for k, v in sizes.items())]
    if tf32:
        cmd.append("--tf32")
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=tempfile.mkdtemp(prefix="ref_run_"))
    except subprocess.TimeoutExpired:
        return {"impl": "reference", "unavailable": f"timed out after {timeout} s"}
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            try:
                return json.loads(line)
            except json.JSONDecodeError:
                continue
    tail = (out.stderr or out.stdout).strip().splitlines()[-1:] or ["no output"]
    return {"impl": "reference", "unavailable": f"harness failed (rc {out.returncode}): {tail[0][:200]}"}


def run_reference(args):
    """CPU arm: the reference's own implementation on the host cores, bounded sample of the headline workload."""
    if int(os.environ.get("RANK", 0)) != 0:
        return
    cores = use_all_host_threads()
    factor = args.sample if args.sample else CPU_SAMPLE[args.config]
    res = run_harness(args.config, "cpu", args.steps, args.warmup, factor=factor, threads=cores)
    cfg = wl.CONFIGS[args.config]
    if "strict" not in res:  # baseline/_ref missing: fall back to the oracle port of the same algorithm
        res = time_oracle_port(args.config, factor, args.steps, args.warmup)
        kind = "port"
    else:
        kind = "reference"
    pps, ms = res["strict"]["value"], res["strict"]["ms_per_step"]
    sample = (f"{'+'.join(str(v) for v in res['sizes'].values())} points/step ({factor:g} of the one-GPU workload), "
              f"{args.steps} steps, " + ("unmodified idrlnet 2.0.0 (baseline/_ref), strict pass of idrlnet/solver.py:329-339, "
                                         if kind == "reference" else "oracle port (torch autograd), ") + "FP32 on the host cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{cfg['title']}: reference implementation on the host CPU, bounded sample",
                   "net": cfg["net_desc"], "points_per_step": res["points_per_step"]},
        "cpu_baseline": {"value": pps, "unit": "points/s", "cores": res.get("threads", cores), "kind": kind, "sample": sample},
        "e2e": {"value": pps, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def time_oracle_port(config, factor, steps, warmup):
    """Fallback CPU baseline when baseline/_ref is absent: the oracle's restatement of the reference algorithm."""
    from oracle import pinn_oracle as po
    import idrlnet_b200 as pkg
    pkg.use_cpu()
    import idrlnet_b200.shortcut as sc
    cfg = wl.CONFIGS[config]
    sizes = wl.scaled_sizes(cfg["sizes"](1), factor)
    rng = np.random.default_rng(0)
    b = cfg["sample"](rng, sizes)
    torch.manual_seed(0)
    nets = cfg["nets"](sc)
    exprs = {sn.name: str(sn.evaluate.tf_eq) for pde in cfg["pdes"](sc) for sn in getattr(pde, "sub_nodes", [])
             if hasattr(sn.evaluate, "tf_eq")}
    node = nets[0]
    layers = []
    for k, l in node.net.layers.items():
        if k.endswith("_activation"):
            continue
        if hasattr(l, "weight_g"):
            layers.append({"g": l.weight_g.detach().numpy().copy(), "v": l.weight_v.detach().numpy().copy(), "b": l.bias.detach().numpy().copy()})
        else:
            layers.append({"W": l.weight.detach().numpy().copy(), "b": l.bias.detach().numpy().copy()})
    act = {"burgers": "silu", "poisson": "silu", "ldc": "tanh", "allen_cahn": "tanh", "ns_xl": "silu"}[config]
    domains = []
    for row in cfg["domains"]:
        name, skey, colmap, targets, lambdas, area, loss = row
        if any(t.startswith("difference") for t in targets):
            continue
        n = sizes[skey]
        pts = {k: b[c].reshape(-1, 1) for k, c in colmap.items()}
        pts["area"] = np.full((n, 1), area(sizes), np.float32)
        val = lambda v: b[v].reshape(-1, 1) if isinstance(v, str) else float(v)
        domains.append({"name": name, "points": pts, "targets": {k: val(v) for k, v in targets.items()},
                        "lambdas": {k: val(v) for k, v in lambdas.items()}, "loss": loss, "sigma": 1.0})
    problem = {"nets": [{"name": node.name, "inputs": list(node.inputs), "outputs": list(node.outputs), "act": act,
                         "layers": layers, "shares": None}], "exprs": exprs, "domains": domains}
    onets = po.build_nets(problem)
    leaves = [t for lay in onets[0]["raw"] for t in lay.values()]
    opt = torch.optim.Adam(leaves, lr=1e-3)

    def one_step():
        opt.zero_grad()
        onets[0]["layers_t"] = [((torch._weight_norm(l["v"], l["g"], 0) if "g" in l else l["W"]), l["b"]) for l in onets[0]["raw"]]
        total = 0.0
        for dom in problem["domains"]:
            var = po.forward_domain(problem, dom, onets)
            like = next(iter(var.values()))
            diff = {k: var[k] - po._as_t(t, torch.float32, like) for k, t in dom["targets"].items()}
            lam = {k: po._as_t(t, torch.float32, like) for k, t in dom["lambdas"].items()}
            total = total + po.weighted_loss(diff, lam, var["area"], dom["loss"])
        total.backward()
        opt.step()

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    pts = sum(len(next(iter(d["points"].values()))) for d in domains)
    return {"strict": {"value": pts / dt, "ms_per_step": dt * 1e3}, "sizes": sizes, "points_per_step": pts,
            "threads": torch.get_num_threads()}


# --------------------------------------------------------------------------- #
# this package
# --------------------------------------------------------------------------- #
def ncu_traffic(config, n_points=None):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json; scaled linearly when the capture was taken at another point
    count than the timed launch); None if not captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            e = json.load(f).get(config, {})
        b = e.get("dram_bytes_per_launch")
        if b is not None and n_points and e.get("points"):
            b = int(b * n_points / e["points"])
        return b
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


def measure_tf32_peak(device):
    """Dense TF32 tcgen05 throughput (TFLOP/s) measured by the library's own probe (one persistent CTA per SM issuing
    M=128 N=256 K=8 kind::tf32 MMAs on resident shared-memory operands); falls back to half the measured sustained
    BF16 cuBLAS rate when the probe is not in this build."""
    from idrlnet_b200 import native as nv
    lib = nv.lib()
    if hasattr(lib, "pinn_tf32_mma_probe"):
        scratch = torch.zeros(1024, device=device)
        st = torch.cuda.current_stream().cuda_stream
        if lib.pinn_tf32_mma_probe(200, scratch.data_ptr(), st) > 0:
            torch.cuda.synchronize()
            best = 0.0
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                flops = lib.pinn_tf32_mma_probe(4000, scratch.data_ptr(), st)
                e1.record()
                torch.cuda.synchronize()
                best = max(best, flops / (e0.elapsed_time(e1) * 1e-3) / 1e12)
            if best > 0:
                return best, "dense TF32 tcgen05 rate measured in this run (pinn_tf32_mma_probe)"
    pk, src = measured_peaks()
    return pk.get("bf16_tflops_sustained", 1125.0) / 2, f"1/2 of the sustained BF16 cuBLAS rate, {src}"


class OursRun:
    """One workload on this package's public API: Solver + device point sets."""

    def __init__(self, name, device, world, rank, precision=None, sizes=None):
        import idrlnet_b200.shortcut as sc
        self.name, self.device, self.world, self.rank = name, device, world, rank
        self.cfg = wl.CONFIGS[name]
        self.sizes = dict(sizes) if sizes else self.cfg["sizes"](world)
        self.precision = precision
        self.sets = wl.PointSets(self.cfg, self.sizes, device, torch)
        self.rng = np.random.default_rng(1000 + rank)
        hb = self.cfg["sample"](self.rng, self.sizes)
        self.batch_bytes = sum(hb[c].nbytes for c in self.sets.columns)
        self.n_rotate = max(2, min(32, int(math.ceil(2.15 * L2_BYTES / self.batch_bytes))))
        if self.batch_bytes > 1.2 * L2_BYTES:
            self.n_rotate = 2
        self.sets.add_host_batch(hb)
        for _ in range(self.n_rotate - 1):
            self.sets.add_host_batch(self.cfg["sample"](self.rng, self.sizes))
        torch.manual_seed(0)
        domains, nets, pdes = wl.build_problem(sc, self.cfg, self.sets, requires_grad=False)
        self.solver = sc.Solver(sample_domains=domains, netnodes=nets, pdes=pdes, max_iter=10 ** 9, loading=False,
                                network_dir=tempfile.mkdtemp(prefix="bench_net_"), post_step_loss=False,
                                precision=precision, shard_points=False,
                                opt_config=dict(optimizer="Adam", lr=1e-3, fused=True))
        self.points_per_step = wl.points_per_step(self.cfg, self.sizes)
        self.fused_domains = sorted(self.solver._fused_domains)

    def step(self, set_idx):
        self.sets.current = set_idx
        return self.solver.train_pipe()

    def sync_all(self):
        torch.cuda.synchronize()
        if self.world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def timed(self, steps, warmup):
        """(ms per step [max over ranks], mean kernel ms of the timed domain, final loss, host ms per step, launches)."""
        from idrlnet_b200 import native as nv
        for i in range(warmup):
            self.step(i % self.n_rotate)
        self.sync_all()
        events = self.solver.kernel_events.setdefault(self.cfg["timed"], [])
        events.clear()
        launches0 = nv.lib().pinn_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        loss = None
        for i in range(steps):
            loss = self.step((warmup + i) % self.n_rotate)
        host_ms = (time.perf_counter() - t0) * 1e3 / steps
        e1.record()
        self.sync_all()
        launches = nv.lib().pinn_launch_count() - launches0
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.device)
        if self.world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        kern_ms = float(np.mean([a.elapsed_time(b) for a, b in events])) if events else None
        self.solver.kernel_events.pop(self.cfg["timed"], None)
        return float(ms) / steps, kern_ms, float(loss), host_ms, int(launches)

    def e2e(self, steps, warmup):
        """Same metric end to end: every step's coordinates come from pinned host memory (H2D on a copy stream,
        double-buffered against compute) and the step's loss is read back to the host."""
        n_host = 4
        cols = self.sets.columns
        host_sets = []
        for _ in range(n_host):
            hb = self.cfg["sample"](self.rng, self.sizes)
            host_sets.append({k: torch.as_tensor(hb[k].reshape(-1, 1)).pin_memory() for k in cols})
        slots = [self.sets.add_empty(), self.sets.add_empty()]
        copy_stream = torch.cuda.Stream(device=self.device)
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]
        loss_host = torch.zeros(steps + warmup + 2, dtype=torch.float32).pin_memory()
        h2d_bytes = sum(host_sets[0][k].numel() * 4 for k in cols)

        def upload(i):
            slot = i % 2
            with torch.cuda.stream(copy_stream), torch.no_grad():
                copy_stream.wait_event(consumed[slot])
                dst = self.sets.sets[slots[slot]]
                for k in cols:
                    dst[k].copy_(host_sets[i % n_host][k], non_blocking=True)
                ready[slot].record(copy_stream)

        def loop(n):
            main = torch.cuda.current_stream()
            upload(0)
            for i in range(n):
                slot = i % 2
                if i + 1 < n:
                    upload(i + 1)  # overlaps with this step's compute
                main.wait_event(ready[slot])
                l = self.step(slots[slot])
                consumed[slot].record(main)
                loss_host[i].copy_(l, non_blocking=True)

        for sl in range(2):
            consumed[sl].record(torch.cuda.current_stream())
        loop(warmup)
        self.sync_all()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        loop(steps)
        t1.record()
        self.sync_all()
        ms = torch.tensor([t0.elapsed_time(t1)], device=self.device)
        if self.world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return self.points_per_step * self.world / (float(ms) / steps * 1e-3), h2d_bytes

    def close(self):
        self.solver = None
        self.sets = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()


def sampled_leg(device, world, steps, warmup):
    """SURVEY.md 8 a7 end to end: the Burgers example AS WRITTEN (examples/burgers_equation/burgers_equation.py:7-45, its
    densities raised to the benchmark's 2^20 + 10240 + 10240 points) under `Solver(device_sampling=True)`: every step draws
    fresh points with the Philox kernels on the GPU (rank-offset streams under torchrun) -- no host sampling, no H2D copy --
    then runs the fused step.  Returns whole-job points/s."""
    import sympy as sp
    import idrlnet_b200.shortcut as sc
    from idrlnet_b200.geo_utils import geo as geo_mod
    x, t = sp.Symbol("x"), sp.Symbol("t")
    time_range = {t: (0, 1)}
    line = sc.Line1D(-1.0, 1.0)
    geo_mod.set_device_sampling(True, seed=1234)

    @sc.datanode(name="burgers_equation")
    def interior_domain():
        return line.sample_interior(1 << 19, bounds={x: (-1.0, 1.0)}, param_ranges=time_range), {"burgers_u": 0}

    @sc.datanode(name="t_boundary")
    def init_domain():
        return line.sample_interior(5120, param_ranges={t: 0.0}), sc.Variables({"u": -sp.sin(math.pi * x)})

    @sc.datanode(name="x_boundary")
    def boundary_domain():
        return line.sample_boundary(5120, param_ranges=time_range), sc.Variables({"u": 0})

    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    solver = sc.Solver(sample_domains=(interior_domain(), init_domain(), boundary_domain()), netnodes=[net],
                       pdes=[sc.BurgersNode(u="u", v=wl.NU)], max_iter=10 ** 9, loading=False,
                       network_dir=tempfile.mkdtemp(prefix="bench_sampled_"), post_step_loss=False, device_sampling=True,
                       sampling_seed=1234, opt_config=dict(optimizer="Adam", lr=1e-3, fused=True))
    try:
        assert len(solver._fused_domains) == 3
        pts = sum(int(v["x"].shape[0]) for v in solver.sample_variables_from_domains().values())

        def sync_all():
            torch.cuda.synchronize()
            if world > 1:
                torch.distributed.barrier()
                torch.cuda.synchronize()
        for _ in range(warmup):
            solver.train_pipe()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = solver.train_pipe()
        host_ms = (time.perf_counter() - t0) * 1e3 / steps
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return {"value": pts * world / (float(ms) / steps * 1e-3), "unit": "points/s", "ms_per_step": float(ms) / steps,
                "host_ms_per_step": host_ms, "points_per_step_per_gpu": pts, "final_loss": float(loss),
                "what": "examples/burgers_equation as written under Solver(device_sampling=True): points drawn on the GPU every "
                        "step (Philox, rank-offset streams), h2d_bytes_per_step = 0"}
    finally:
        geo_mod.set_device_sampling(False)


def roofline_of(run: OursRun, kern_ms, fp32_peak, tf32_peak, hbm_peak, hbm_src):
    cfg = run.cfg
    n_timed = run.sizes[next(d[1] for d in cfg["domains"] if d[0] == cfg["timed"])]
    if kern_ms is None:
        return None
    achieved = cfg["flop_per_point"] * n_timed / (kern_ms * 1e-3) / 1e12
    tensor = cfg["wide"] and run.precision in ("tf32", "tf32_streamed", "fp32x3")
    if tensor:
        peak, src = tf32_peak
        if run.precision == "fp32x3":
            peak, src = peak / 3, src + ", / 3 MMAs per product (3xTF32 operand splitting)"
        bound = "tensor"
    else:
        peak, src, bound = fp32_peak, "FP32 FMA pipe, pinn_fp32_fma_probe measured in this run", "fp32"
    hbm = cfg["bytes_per_point"] * n_timed / (kern_ms * 1e-3) / 1e9
    return {"bound": bound, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            "traffic": ncu_traffic(run.name if not run.precision or not cfg["wide"] else f"{run.name}_{run.precision}", n_timed),
            "kernel": cfg["kernel"] + (f" [{run.precision}]" if cfg["wide"] else ""), "kernel_ms": kern_ms,
            "flop_per_point": cfg["flop_per_point"], "peak_source": src, "fp32_fma_peak": fp32_peak,
            "hbm": {"achieved": hbm, "peak": hbm_peak, "unit": "GB/s", "frac": hbm / hbm_peak, "peak_source": hbm_src}}


def torch_cuda_arm(name, local, steps=20, warmup=3, pipe_steps=4):
    """The unmodified reference on this GPU (`idrlnet.use_gpu`), same workload, at the size where it ran fastest in the
    sweep from the BASELINE size down to 1/64 of it (tools/ref_sweep.py -> baseline/ref_best_sizes.json; the reference
    keeps the whole double-backward graph of a step in HBM -- 41 KB / point for ldc, 1.1 MB / point for mlp_xl -- so its
    feasible N ends well below the 16 M / 64 M points of the named configs); without that file: the BASELINE size,
    halved by the harness until it fits."""
    sizes = None
    path = os.path.join(ROOT, "baseline", "ref_best_sizes.json")
    if os.path.exists(path):
        with open(path) as f:
            sizes = json.load(f).get(name, {}).get("sizes")
    clk = ClockSampler(local).start()
    res = run_harness(name, "cuda", steps, warmup, pipe_steps=pipe_steps, gpu_index=local, sizes=sizes)
    res["clocks"] = clk.stop()
    res["size_choice"] = "best of the size sweep (baseline/ref_best_sizes.json)" if sizes else "BASELINE size, halved on OOM"
    return res


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
    import idrlnet_b200 as pkg
    pkg.use_gpu(local)
    from idrlnet_b200 import native as nv
    nv.lib()  # fail loudly if the CUDA library is missing

    cfg = wl.CONFIGS[args.config]
    sizes = wl.scaled_sizes(cfg["sizes"](world), args.sample) if args.sample else None
    head_prec = args.precision if cfg["wide"] else None
    names = [] if args.no_configs else [n for n in ("poisson", "ldc", "allen_cahn", "ns_xl") if n != args.config]
    if world > 1:
        names = [n for n in names if n in ("ldc", "ns_xl")]  # the configs the north star names for multi-GPU runs
    # ---- the reference's own PyTorch CUDA path on this B200 (the >= 20x denominator), FIRST: each arm
    # is a fresh process that needs the whole 180 GB (116 GB for ldc at 2^21 points, 177 GB for mlp_xl at 2^17), so it runs
    # before this process allocates anything on the device
    ref_arms = {}
    if world == 1 and not args.no_torch_cuda:
        for name in [args.config] + names:
            ref_arms[name] = torch_cuda_arm(name, local)
    clocks = ClockSampler(local).start()  # samples cover warm-up + timed region + e2e region (all under load)
    run = OursRun(args.config, device, world, rank, precision=head_prec, sizes=sizes)
    ms_per_step, kern_ms, final_loss, host_ms, launches = run.timed(args.steps, args.warmup)
    value = run.points_per_step * world / (ms_per_step * 1e-3)
    e2e_value, h2d_bytes = run.e2e(args.steps, args.warmup)
    sampled = None
    if args.config == "burgers" and not args.sample:
        try:
            sampled = sampled_leg(device, world, args.steps, args.warmup)
        except Exception as e:  # never take the headline down
            sampled = {"error": f"{type(e).__name__}: {e}"[:300]}
    clk = clocks.stop()
    pk, pk_src = measured_peaks()
    fp32_peak = measure_fp32_peak(device)
    tf32_peak = measure_tf32_peak(device)
    roof = roofline_of(run, kern_ms, fp32_peak, tf32_peak, pk["hbm_gbs"], pk_src)
    head_sizes, head_rot, head_bytes, head_fused = run.sizes, run.n_rotate, run.batch_bytes, run.fused_domains
    head_pts = run.points_per_step
    head_allreduce = None
    if world > 1:
        fused_in_finalize = all(getattr(e, "_peers", None) is not None for e in run.solver._engines.values())
        head_allreduce = ("summed over the ranks' symmetric memory inside the finalize kernel (pinn_step_finalize_allreduce)"
                          if fused_in_finalize and run.solver._engines else "NCCL all_reduce after finalize")
    run.close()

    # ---- the other BASELINE configurations at their stated sizes -------------------------------------------
    configs = {}
    for name in names:
        c = wl.CONFIGS[name]
        modes = WIDE_MODES if c["wide"] else (("", None),)
        for suffix, prec in modes:
            w, k = SECONDARY_STEPS[name]
            if name == "ns_xl" and world > 1:
                k = 3  # 2^26 / N points per GPU: a step takes seconds
            try:
                r = OursRun(name, device, world, rank, precision=prec)
                ck = ClockSampler(local).start()
                ms, kms, loss, hms, _ = r.timed(k, w)
                entry = {"value": r.points_per_step * world / (ms * 1e-3), "unit": "points/s", "ms_per_step": ms,
                         "steps": k, "warmup": w, "scaling": c["scaling"], "precision": prec or "fp32",
                         "workload": f"{c['title']}: {' + '.join(str(r.sizes[d[1]]) for d in c['domains'])} points per step per GPU "
                                     f"({', '.join(d[0] for d in c['domains'])})",
                         "net": c["net_desc"], "points_per_step_per_gpu": r.points_per_step, "fused_domains": r.fused_domains,
                         "l2": f"{r.n_rotate} rotating point batches of {r.batch_bytes / 1e6:.0f} MB",
                         "host_ms_per_step": hms, "final_loss": loss,
                         "roofline": roofline_of(r, kms, fp32_peak, tf32_peak, pk["hbm_gbs"], pk_src), "clocks": ck.stop()}
                r.close()
            except Exception as e:  # a secondary config must never take the headline line down
                entry = {"error": f"{type(e).__name__}: {e}"[:300]}
            configs[name + suffix] = entry

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = use_all_host_threads()
            res = run_harness(args.config, "cpu", 100 if args.config in ("burgers", "poisson") else 6, 2,
                              factor=CPU_SAMPLE[args.config], threads=cores)
            if "strict" in res:
                cpu = {"value": res["strict"]["value"], "unit": "points/s", "cores": res.get("threads", cores), "kind": "reference",
                       "sample": f"{'+'.join(str(v) for v in res['sizes'].values())} points/step x {res['strict']['steps']} steps of the same "
                                 f"workload (unmodified idrlnet 2.0.0 from baseline/_ref, strict pass idrlnet/solver.py:329-339, FP32)"}
            else:
                cpu = {"value": None, "unit": "points/s", "cores": cores, "kind": "reference", "sample": res.get("unavailable", "failed")}
        head_ref = head_ratio = None
        for name, ref in ref_arms.items():
            ref_v = ref.get("strict", {}).get("value")
            if name == args.config:
                head_ref, head_ratio = ref, (value / ref_v if ref_v else None)
                continue
            for key in [k for k in configs if k == name or k.startswith(name + "_")]:
                ours_v = configs[key].get("value")
                configs[key]["torch_cuda"] = ref
                configs[key]["vs_torch_cuda"] = (ours_v / ref_v) if (ours_v and ref_v) else None
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": cfg["scaling"],
            "vs_baseline": None,
            "dtype": "f32" if not cfg["wide"] or head_prec in (None, "fp32") else ("f32 (3xTF32 split on tensor cores)" if head_prec == "fp32x3" else "tf32"),
            "data": "synthetic",
            "config": {"workload": f"{cfg['title']}: {' + '.join(str(head_sizes[d[1]]) for d in cfg['domains'])} "
                                   f"points per step per GPU ({', '.join(d[0] for d in cfg['domains'])})",
                       "net": f"{cfg['net_desc']}, Adam 1e-3 + ExponentialLR",
                       "api": "idrlnet_b200.Solver.train_pipe() (post_step_loss=False: one evaluation per step, the strict step "
                              "of idrlnet/solver.py:329-339)", "fused_domains": head_fused,
                       "points_per_step_per_gpu": head_pts, "parallelism": f"dp{world} (points sharded, 1 all-reduce)", "allreduce": head_allreduce,
                       "l2": f"{head_rot} rotating point batches ({head_rot * head_bytes / 1e6:.0f} MB > 126 MB L2)"},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "points/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "e2e_device_sampling": sampled,
            "gpu_launches": launches, "clocks": clk, "final_loss": final_loss, "host_ms_per_step": host_ms,
            "configs": configs,
        }
        if world == 1 and not args.no_torch_cuda:
            line["torch_cuda"], line["vs_torch_cuda"] = head_ref, head_ratio
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="burgers", choices=list(wl.CONFIGS),
                    help="headline workload; the driver's default (burgers) is BASELINE configs[1]")
    ap.add_argument("--sample", type=float, default=None, help="scale every domain size of the headline by this factor")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only: skip the `configs` object")
    ap.add_argument("--no-torch-cuda", action="store_true", help="skip the reference-on-CUDA arms (N=1 only anyway)")
    ap.add_argument("--precision", default="fp32x3", choices=["fp32", "tf32", "tf32_streamed", "fp32x3"],
                    help="layer GEMMs of a WIDE headline config; narrow nets are always FP32 FMA")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
