#!/usr/bin/env python
"""Times the UNMODIFIED reference (idrl-lab/idrlnet v2.0.0, installed from /root/reference into baseline/_ref by
`pip install --no-index --no-deps --target baseline/_ref`) on the workloads of baseline/workloads.py, through its own
public API and stock code path:

    sc.use_gpu(0)                      (idrlnet/__init__.py:20-34; omitted for --device cpu)
    Solver(sample_domains, netnodes, pdes, ...)
    strict pass  = opt.zero_grad(); sample; forward_through_all_graph; compute_loss; backward; opt.step()
                   (idrlnet/solver.py:329-339, the first half of train_pipe: BASELINE.md 3.1 (ii))
    train_pipe   = Solver.train_pipe() as shipped, duplicated forward included (idrlnet/solver.py:309-351; 3.1 (i))

None of this repository's kernels, models or engine is on that path: only the workload table (points, nets and PDE
nodes built with the reference's own classes) is shared.  Import-time shims are the five of SURVEY.md 8(c1) (stub
pyevtk / matplotlib, collections.Callable, LR-scheduler reflection, no graph PNG) -- none touches the compute path.

Prints ONE JSON line.  Run as a subprocess by bench.py (own CUDA context, own default tensor type).
"""
import argparse
import collections
import collections.abc
import json
import os
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")


def install_shims():
    def stub(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    class Anything:
        def __getattr__(self, k):
            return Anything()

        def __call__(self, *a, **k):
            return Anything()

        def __iter__(self):
            return iter([Anything(), Anything()])

    def getattr_any(k):
        if k.startswith("__"):
            raise AttributeError(k)
        return Anything()

    try:
        import pyevtk.hl  # noqa: F401
    except Exception:
        stub("pyevtk")
        stub("pyevtk.hl", pointsToVTK=lambda *a, **k: None)
    try:
        import matplotlib.pyplot  # noqa: F401
        import matplotlib.tri  # noqa: F401
    except Exception:
        mp = stub("matplotlib")
        plt = stub("matplotlib.pyplot")
        plt.__getattr__ = getattr_any
        stub("matplotlib.tri").__getattr__ = getattr_any
        mp.pyplot = plt
    collections.Callable = collections.abc.Callable


def load_reference():
    """idrlnet.shortcut of the unmodified reference (baseline/_ref), or raise ImportError."""
    if not os.path.isdir(os.path.join(REF_DIR, "idrlnet")):
        raise ImportError(f"{REF_DIR}/idrlnet not found (pip install --no-index --no-deps --target baseline/_ref /root/reference)")
    install_shims()
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp(prefix="idrlnet_ref_"))  # the reference writes train.log into the CWD at import
    try:
        sys.path.insert(0, REF_DIR)
        import torch
        import idrlnet
        import idrlnet.shortcut as sc
        from idrlnet.optim import Optimizable, get_available_class
        import idrlnet.graph as g
        assert os.path.realpath(idrlnet.__file__).startswith(os.path.realpath(REF_DIR)), idrlnet.__file__
        Optimizable.SCHEDULE_MAP = get_available_class(torch.optim.lr_scheduler, torch.optim.lr_scheduler.LRScheduler)
        g.VertexTaskPipeline.display = lambda self, filename=None: None
    finally:
        os.chdir(cwd)
    return sc


def strict_step(solver):
    """idrlnet/solver.py:329-339 verbatim in effect: what train_pipe does before its duplicated forward."""
    for opt in solver.optimizers:
        opt.zero_grad()
        samples = solver.sample_variables_from_domains()
        in_var, true_out, lambda_out = solver.generate_in_out_dict(samples)
        pred = solver.forward_through_all_graph(in_var, solver.outvar_dict_index)
        loss = solver.compute_loss(in_var, pred, true_out, lambda_out)
        loss.backward()
        opt.step()
    return loss


def run(args):
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    from baseline import workloads as wl
    sc = load_reference()
    cuda = args.device == "cuda"
    if cuda:
        if not torch.cuda.is_available():
            return {"impl": "reference", "unavailable": "no CUDA device"}
        sc.use_gpu(args.gpu_index)  # idrlnet/__init__.py:20-34: default tensor type -> torch.cuda.FloatTensor
        import idrlnet
        if not idrlnet.GPU_ENABLED:
            return {"impl": "reference", "unavailable": "idrlnet.use_gpu failed"}
        device = torch.device(f"cuda:{args.gpu_index}")
    else:
        device = torch.device("cpu")
        if args.threads:
            torch.set_num_threads(args.threads)
    if args.tf32:
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    torch.autograd.set_detect_anomaly(False)  # examples/ldc/ldc.py:62 turns it on; the strict baseline runs without
    cfg = wl.CONFIGS[args.config]
    base = cfg["sizes"](1)
    if args.sizes:
        base = {k: int(v) for k, v in (kv.split("=") for kv in args.sizes.split(","))}
    factor = args.factor
    result = None
    attempts = []
    while True:
        sizes = wl.scaled_sizes(base, factor) if factor != 1.0 else dict(base)
        try:
            result = time_sizes(sc, wl, cfg, sizes, device, args, np, torch)
            attempts.append({"points": wl.points_per_step(cfg, sizes), "ok": True})
            break
        except torch.cuda.OutOfMemoryError:
            attempts.append({"points": wl.points_per_step(cfg, sizes), "ok": False, "why": "CUDA out of memory"})
            torch.cuda.empty_cache()
            factor *= 0.5
            if factor < 1 / 4096:
                return {"impl": "reference", "unavailable": "out of memory at every size tried", "attempts": attempts}
    result["attempts"] = attempts
    return result


def time_sizes(sc, wl, cfg, sizes, device, args, np, torch):
    rng = np.random.default_rng(0)
    sets = wl.PointSets(cfg, sizes, device, torch)
    sets.add_host_batch(cfg["sample"](rng, sizes))
    torch.manual_seed(0)
    domains, nets, pdes = wl.build_problem(sc, cfg, sets, requires_grad=True)
    solver = sc.Solver(sample_domains=domains, netnodes=nets, pdes=pdes, max_iter=10 ** 9, loading=False,
                       network_dir=tempfile.mkdtemp(prefix="ref_net_"))
    pts = wl.points_per_step(cfg, sizes)
    sync = (lambda: torch.cuda.synchronize(device)) if device.type == "cuda" else (lambda: None)
    out = {"impl": "reference", "config": args.config, "device": str(device), "points_per_step": pts,
           "sizes": sizes, "torch": torch.__version__, "tf32_allowed": bool(args.tf32),
           "what": "unmodified idrlnet 2.0.0 from baseline/_ref" + (", idrlnet.use_gpu(%d)" % args.gpu_index if device.type == "cuda" else ", CPU"),
           "threads": torch.get_num_threads()}
    if device.type == "cuda":
        torch.cuda.reset_peak_memory_stats(device)
    for mode, fn, steps, warmup in (("strict", lambda: strict_step(solver), args.steps, args.warmup),
                                    ("train_pipe", solver.train_pipe, args.pipe_steps, min(args.warmup, 2))):
        if steps <= 0:
            continue
        try:
            for _ in range(warmup):
                fn()
        except torch.cuda.OutOfMemoryError:
            if mode == "strict":
                raise  # the caller halves the problem
            # train_pipe as shipped keeps the graph of its returned (post-step) loss alive into the next iteration: it needs
            # more memory than the strict pass and may not fit at the size where the strict pass runs fastest
            out[mode] = {"unavailable": "CUDA out of memory at this size (the strict pass fits)"}
            torch.cuda.empty_cache()
            continue
        sync()
        if device.type == "cuda":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        loss = None
        for _ in range(steps):
            loss = fn()
        if device.type == "cuda":
            e1.record()
        sync()
        if mode == "strict":
            loss_val = float(loss)
            del loss  # drop the graph before the next mode runs
            loss = loss_val
        wall = (time.perf_counter() - t0) / steps
        dt = e0.elapsed_time(e1) * 1e-3 / steps if device.type == "cuda" else wall
        out[mode] = {"value": pts / dt, "unit": "points/s", "ms_per_step": dt * 1e3, "steps": steps, "warmup": warmup,
                     "wall_ms_per_step": wall * 1e3, "final_loss": float(loss)}
    if device.type == "cuda":
        out["max_memory_gb"] = torch.cuda.max_memory_allocated(device) / 2 ** 30
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="burgers")
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--gpu-index", type=int, default=0)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pipe-steps", type=int, default=0, help="also time Solver.train_pipe() as shipped for this many steps")
    ap.add_argument("--factor", type=float, default=1.0, help="scale every domain size (halved automatically on CUDA OOM)")
    ap.add_argument("--sizes", default=None, help="explicit sizes, e.g. int=65536,data=5000")
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--tf32", action="store_true", help="allow TF32 matmuls (torch 1.7.1, the reference's pin, did so silently)")
    args = ap.parse_args()
    try:
        res = run(args)
    except ImportError as e:
        res = {"impl": "reference", "unavailable": str(e).splitlines()[0][:200]}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
