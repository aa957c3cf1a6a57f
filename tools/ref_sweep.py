"""sweep the problem size of the UNMODIFIED reference on CUDA (baseline/ref_harness.py) from the
BASELINE size downwards and keep the best points/s per config -> baseline/ref_best_sizes.json (read by bench.py for
its reference-on-CUDA arm) and a full record under profiles/.   python tools/ref_sweep.py [out.json] [configs...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from baseline import workloads as wl  # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ref_cuda_sweep.json")
configs = sys.argv[2:] or list(wl.CONFIGS)
record, best = {}, {}
for name in configs:
    record[name] = []
    factor = 1.0
    seen = set()
    while factor >= 1 / 64:
        clk = bench.ClockSampler(0).start()
        res = bench.run_harness(name, "cuda", 10, 3, factor=factor)
        res["clocks"] = clk.stop()
        res["requested_factor"] = factor
        pts = res.get("points_per_step")
        if pts in seen:
            factor /= 2
            continue
        seen.add(pts)
        record[name].append(res)
        if "strict" in res:
            print(name, pts, f"{res['strict']['value']:.4g} pts/s", f"{res.get('max_memory_gb', 0):.1f} GB", flush=True)
            if name not in best or res["strict"]["value"] > best[name]["value"]:
                best[name] = {"value": res["strict"]["value"], "sizes": res["sizes"], "points_per_step": pts,
                              "max_memory_gb": res.get("max_memory_gb")}
            # the harness halves on OOM by itself: continue below the size that actually ran
            base = wl.points_per_step(wl.CONFIGS[name], wl.CONFIGS[name]["sizes"](1))
            factor = min(factor, pts / base) / 2
        else:
            print(name, "unavailable:", res.get("unavailable"), flush=True)
            break
os.makedirs(os.path.dirname(out_path), exist_ok=True)
with open(out_path, "w") as f:
    json.dump({"best": best, "runs": record}, f, indent=1)
with open(os.path.join(os.path.dirname(out_path), "ref_best_sizes.json"), "w") as f:
    json.dump(best, f, indent=1)
print(json.dumps(best))
