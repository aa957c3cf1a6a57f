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
N_INTERIOR = 1 << 20
N_IC = 10240
N_BC = 10240
NU = 0.01 / math.pi
N_ROTATE = 32  # distinct point batches; 32 * 8.4 MB = 270 MB > 126 MB L2
# algorithmic work per point, SURVEY.md 8(d4): F_step = 3 * 2 * [d*H + C*(L-1)*H^2 + n_pairs*H]
FLOP_INTERIOR = 3 * 2 * (2 * 20 + 4 * 3 * 400 + 4 * 20)  # 29 520
BYTES_INTERIOR = 8  # x, t FP32 (area/target constant)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- #
# synthetic Burgers workload
# --------------------------------------------------------------------------- #
def make_batch_host(rng, n_int, n_ic, n_bc):
    """One step's collocation points as float32 numpy columns."""
    f = np.float32
    ic_x = rng.uniform(-1, 1, n_ic).astype(f)
    return {
        "int_x": rng.uniform(-1, 1, n_int).astype(f), "int_t": rng.uniform(0, 1, n_int).astype(f),
        "ic_x": ic_x, "ic_t": np.zeros(n_ic, f), "ic_u": (-np.sin(np.pi * ic_x)).astype(f),
        "bc_x": np.where(rng.uniform(size=n_bc) < 0.5, -1.0, 1.0).astype(f), "bc_t": rng.uniform(0, 1, n_bc).astype(f),
    }


BATCH_KEYS = ("int_x", "int_t", "ic_x", "ic_t", "ic_u", "bc_x", "bc_t")


class BurgersTrainer:
    """Burgers problem wired through the package's public pieces: get_net_node
    (Arch.mlp), BurgersNode, the residual compiler and the fused StepEngine."""

    def __init__(self, device, n_int=N_INTERIOR, n_ic=N_IC, n_bc=N_BC, world=1):
        import idrlnet_b200.shortcut as sc
        from idrlnet_b200 import compiler as cp
        from idrlnet_b200 import fused
        self.device, self.world = device, world
        self.n_int, self.n_ic, self.n_bc = n_int, n_ic, n_bc
        torch.manual_seed(0)
        self.net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
        self.net.net.to(device)
        pde = sc.BurgersNode(u="u", v=NU)
        exprs = {s.name: s.evaluate.tf_eq for s in pde.sub_nodes}
        spec = [cp.NetSpec("net1", ("x", "t"), ("u",))]
        self.cd_int = cp.compile_domain(["burgers_u"], exprs, spec, ["x", "t", "area"])
        self.cd_val = cp.compile_domain(["u"], exprs, spec, ["x", "t", "area"])
        self.pack = fused.NetPack(self.net.net)
        self.fused = fused
        self.params = [p for p in self.net.net.parameters()]
        self.opt = torch.optim.Adam(self.params, lr=1e-3, fused=True)
        self.sched = torch.optim.lr_scheduler.ExponentialLR(self.opt, gamma=math.pow(0.95, 0.001))
        self.engine = None
        self.sets = []

    def bind_set(self, cols):
        """Three FusedDomains bound to one set of device columns."""
        f = self.fused
        d_int = f.FusedDomain("burgers_equation", self.cd_int)
        d_int.bind({"x": cols["int_x"], "t": cols["int_t"]}, {"burgers_u": 0.0}, {}, 1.0 / (self.n_int / 2))
        d_ic = f.FusedDomain("t_boundary", self.cd_val)
        d_ic.bind({"x": cols["ic_x"], "t": cols["ic_t"]}, {"u": cols["ic_u"]}, {}, 2.0 / self.n_ic)
        d_bc = f.FusedDomain("x_boundary", self.cd_val)
        d_bc.bind({"x": cols["bc_x"], "t": cols["bc_t"]}, {"u": 0.0}, {}, 2.0 / self.n_bc)
        doms = [d_ic, d_int, d_bc]  # interior second: its kernel is the one bracketed by events
        if self.engine is None:
            self.engine = f.StepEngine(self.pack, doms)
        self.sets.append(doms)
        return len(self.sets) - 1

    def step(self, set_idx, ev=None):
        """One full training step on the current stream; returns the loss (0-d device tensor)."""
        eng = self.engine
        eng.domains = self.sets[set_idx]
        eng.launch(timed=None if ev is None else (1, ev[0], ev[1]))
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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
# CPU baseline: the oracle (reference algorithm, torch autograd) on host cores
# --------------------------------------------------------------------------- #
def oracle_problem(trainer_state, rng, n_int, n_ic, n_bc):
    b = make_batch_host(rng, n_int, n_ic, n_bc)
    col = lambda a: a.reshape(-1, 1)
    layers = []
    for lin in trainer_state:
        layers.append({k: v.copy() for k, v in lin.items()})
    ones = lambda n, v: np.full((n, 1), v, np.float32)
    return {
        "nets": [{"name": "net1", "inputs": ["x", "t"], "outputs": ["u"], "act": "silu", "layers": layers, "shares": None}],
        "exprs": {"burgers_u": f"u*u__x + u__t - {NU!r}*u__x__x"},
        "domains": [
            {"name": "burgers_equation", "points": {"x": col(b["int_x"]), "t": col(b["int_t"]), "area": ones(n_int, 2.0 / n_int)},
             "targets": {"burgers_u": 0.0}, "lambdas": {}, "loss": "square", "sigma": 1.0},
            {"name": "t_boundary", "points": {"x": col(b["ic_x"]), "t": col(b["ic_t"]), "area": ones(n_ic, 2.0 / n_ic)},
             "targets": {"u": col(b["ic_u"])}, "lambdas": {}, "loss": "square", "sigma": 1.0},
            {"name": "x_boundary", "points": {"x": col(b["bc_x"]), "t": col(b["bc_t"]), "area": ones(n_bc, 2.0 / n_bc)},
             "targets": {"u": 0.0}, "lambdas": {}, "loss": "square", "sigma": 1.0},
        ],
    }


def default_net_state():
    """Arch.mlp weights under torch.manual_seed(0) as oracle-format layers (CPU)."""
    import idrlnet_b200.shortcut as sc
    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp).net
    out = []
    for k, l in net.layers.items():
        if k.endswith("_activation"):
            continue
        out.append({"g": l.weight_g.detach().numpy().copy(), "v": l.weight_v.detach().numpy().copy(),
                    "b": l.bias.detach().numpy().copy()})
    return out


def time_oracle(steps, warmup, n_int, n_ic, n_bc):
    """Reference algorithm on the host: forward graph + loss + backward + Adam."""
    from oracle import pinn_oracle as po
    rng = np.random.default_rng(0)
    state = default_net_state()
    problem = oracle_problem(state, rng, n_int, n_ic, n_bc)
    nets = po.build_nets(problem)
    leaves = [t for lay in nets[0]["raw"] for t in lay.values()]
    opt = torch.optim.Adam(leaves, lr=1e-3)

    def one_step():
        opt.zero_grad()
        # effective weights are functions of the leaves: rebuild them each step
        eff = []
        for leaf in nets[0]["raw"]:
            eff.append((torch._weight_norm(leaf["v"], leaf["g"], 0), leaf["b"]))
        nets[0]["layers_t"] = eff
        total = 0.0
        for dom in problem["domains"]:
            var = po.forward_domain(problem, dom, nets)
            like = next(iter(var.values()))
            diff = {k: var[k] - po._as_t(t, torch.float32, like) for k, t in dom["targets"].items()}
            total = total + po.weighted_loss(diff, {}, var["area"], "square")
        total.backward()
        opt.step()
        return float(total.detach())

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / steps
    return (n_int + n_ic + n_bc) / dt, dt


# --------------------------------------------------------------------------- #
def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = torch.get_num_threads()
    n_int, n_ic, n_bc = 1 << 17, 1280, 1280
    pps, dt = time_oracle(args.steps, args.warmup, n_int, n_ic, n_bc)
    sample = (f"{n_int}+{n_ic}+{n_bc} points/step (1/8 of the GPU workload; throughput is flat in N), "
              f"{args.steps} steps, torch autograd FP32")
    line = {
        "impl": "reference", "metric": METRIC, "value": pps, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "burgers_equation (BASELINE configs[1]), reference algorithm on host CPU, bounded sample",
                   "net": "2-20x4-1 swish weight-norm", "points_per_step": n_int + n_ic + n_bc},
        "cpu_baseline": {"value": pps, "unit": "points/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": pps, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


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

    trainer = BurgersTrainer(device, world=world)
    rng = np.random.default_rng(1000 + rank)
    # ---- device-resident rotating batches (value) ---------------------------
    for _ in range(N_ROTATE):
        hb = make_batch_host(rng, N_INTERIOR, N_IC, N_BC)
        trainer.bind_set({k: torch.as_tensor(hb[k], device=device) for k in BATCH_KEYS})
    pts_per_step = N_INTERIOR + N_IC + N_BC

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        trainer.step(i % N_ROTATE)
    sync_all()
    clocks = ClockSampler(local)
    clocks.start()
    kernel_events = []
    launches0 = nv.lib().pinn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for i in range(args.steps):
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        kernel_events.append(ev)
        loss = trainer.step((args.warmup + i) % N_ROTATE, ev)
    e1.record()
    sync_all()
    launches = nv.lib().pinn_launch_count() - launches0
    clk = clocks.stop()
    ms = torch.tensor([e0.elapsed_time(e1)], device=device)
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
    ms_total = float(ms)
    ms_per_step = ms_total / args.steps
    value = pts_per_step * world / (ms_per_step * 1e-3)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    final_loss = float(loss)

    # ---- end to end: host (pinned) points in, loss out, every step -----------
    n_host = 4
    host_sets = []
    for _ in range(n_host):
        hb = make_batch_host(rng, N_INTERIOR, N_IC, N_BC)
        host_sets.append({k: torch.as_tensor(hb[k]).pin_memory() for k in BATCH_KEYS})
    dev_bufs = [{k: torch.empty_like(host_sets[0][k], device=device) for k in BATCH_KEYS} for _ in range(2)]
    set_ids = [trainer.bind_set(b) for b in dev_bufs]
    copy_stream = torch.cuda.Stream(device=device)
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    loss_host = torch.zeros(args.steps + args.warmup, dtype=torch.float32).pin_memory()
    h2d_bytes = sum(host_sets[0][k].numel() * 4 for k in BATCH_KEYS)

    def upload(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for k in BATCH_KEYS:
                dev_bufs[slot][k].copy_(host_sets[i % n_host][k], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_loop(n, base):
        main = torch.cuda.current_stream()
        upload(base)
        for i in range(n):
            slot = (base + i) % 2
            if i + 1 < n:
                upload(base + i + 1)  # overlaps with this step's compute
            main.wait_event(ready[slot])
            l = trainer.step(set_ids[slot])
            consumed[slot].record(main)
            loss_host[base + i].copy_(l, non_blocking=True)

    for s in range(2):
        consumed[s].record(torch.cuda.current_stream())
    e2e_loop(args.warmup, 0)
    sync_all()
    base = args.warmup + (args.warmup % 2)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(args.steps, 0)
    t1.record()
    sync_all()
    ms2 = torch.tensor([t0.elapsed_time(t1)], device=device)
    if world > 1:
        torch.distributed.all_reduce(ms2, op=torch.distributed.ReduceOp.MAX)
    e2e_value = pts_per_step * world / (float(ms2) / args.steps * 1e-3)

    if rank == 0:
        hbm_peak, peak_src = peaks()
        fp32_peak = measure_fp32_peak(device)
        flops = FLOP_INTERIOR * N_INTERIOR
        achieved = flops / (kern_ms * 1e-3) / 1e12
        hbm_achieved = BYTES_INTERIOR * N_INTERIOR / (kern_ms * 1e-3) / 1e9
        cpu = None
        if world == 1:
            cores = torch.get_num_threads()
            n_int, n_ic, n_bc = 1 << 17, 1280, 1280
            pps, dt = time_oracle(8, 2, n_int, n_ic, n_bc)
            cpu = {"value": pps, "unit": "points/s", "cores": cores, "kind": "port",
                   "sample": f"{n_int}+{n_ic}+{n_bc} points/step x 8 steps of the same workload (oracle, torch autograd FP32)"}
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "burgers_equation (BASELINE configs[1]): 2^20 interior + 10240 IC + 10240 BC points per step per GPU",
                       "net": "Arch.mlp 2-20x4-1 swish weight-norm, Adam 1e-3 + ExponentialLR",
                       "points_per_step_per_gpu": pts_per_step, "parallelism": f"dp{world} (points sharded, 1 all-reduce)",
                       "l2": f"{N_ROTATE} rotating point batches ({N_ROTATE * h2d_bytes / 1e6:.0f} MB > 126 MB L2)"},
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp32_peak if fp32_peak else None, "traffic": None,
                         "kernel": "narrow_kernel<20,2,1,silu> (interior domain)", "kernel_ms": kern_ms,
                         "flop_per_point": FLOP_INTERIOR, "peak_source": "pinn_fp32_fma_probe measured in this run",
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
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
