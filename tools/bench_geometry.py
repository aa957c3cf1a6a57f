"""Sampling throughput of a CSG solid (rectangle minus circle), SURVEY.md 8 rows a7 / f3.

    python tools/bench_geometry.py [density]

device : Geometry.sample_interior_device -> pinn_sample_csg (Philox candidates, sdf program, ordered compaction)
host   : Geometry.sample_interior (NumPy, cached SymPy lowering)
Prints one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import idrlnet_b200.shortcut as sc  # noqa: E402


def main():
    density = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    geo = sc.Rectangle((-1.0, -1.0), (1.0, 1.0)) - sc.Circle((0.2, 0.0), 0.5)
    n_cand = 4 * density
    for _ in range(3):
        out = geo.sample_interior_device(density, seed=1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 20
    for i in range(reps):
        out = geo.sample_interior_device(density, seed=1, offset=i * n_cand)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    kept = out["x"].shape[0]
    host_density = min(density, 1 << 16)
    geo.sample_interior(host_density)
    t0 = time.perf_counter()
    for _ in range(3):
        hp = geo.sample_interior(host_density)
    host_ms = (time.perf_counter() - t0) / 3 * 1e3
    print(json.dumps({
        "solid": "Rectangle((-1,-1),(1,1)) - Circle((0.2,0),0.5)", "candidates": n_cand, "kept": kept,
        "device_ms": ms, "device_candidates_per_s": n_cand / ms * 1e3,
        "device_written_GBps": kept * 3 * 4 / ms / 1e6,
        "host_numpy_candidates": 4 * host_density, "host_numpy_ms": host_ms,
        "host_numpy_candidates_per_s": 4 * host_density / host_ms * 1e3,
        "speedup_vs_host_numpy": (n_cand / ms) / (4 * host_density / host_ms)}))


if __name__ == "__main__":
    main()
