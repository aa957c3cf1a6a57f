"""Timing of the adaptive re-sampling kernels beside the reference recipe.

    python tools/bench_select.py [n_points] [k]

Ours: pinn_residual_forward (jets + residual, forward only) and pinn_topk_abs on the device.
Reference recipe (examples/allen_cahn/allen_cahn.py:134-141): residual column -> host -> np.argsort.
Prints one JSON line; CUDA-event timing, 3 warm-ups, 20 timed repetitions."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from idrlnet_b200 import compiler as cp, fused  # noqa: E402
import idrlnet_b200.shortcut as sc  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = sc.MLP([2, 128, 128, 128, 128, 1], activation=sc.Activation.tanh).to(dev)
    pack = fused.NetPack(net, precision="fp32")
    pde = sc.AllenCahnNode(u="u", gamma_1=0.0001, gamma_2=5)
    exprs = {sub.name: sub.evaluate.tf_eq for sub in pde.sub_nodes}
    cd = cp.compile_domain(["AllenCahn_u"], exprs, [cp.NetSpec("net", ("x", "t"), ("u",))], ["x", "t"])
    fd = fused.FusedDomain("candidates", cd)
    g = torch.Generator(device=dev).manual_seed(0)
    cols = {"x": torch.rand(n, device=dev, generator=g) * 2 - 1, "t": torch.rand(n, device=dev, generator=g)}
    fd.bind(cols, {"AllenCahn_u": 0.0}, {}, 1.0)

    def timed(fn, reps=20, warm=3):
        for _ in range(warm):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    r = fused.domain_residuals(pack, fd)[0]
    ms_res = timed(lambda: fused.domain_residuals(pack, fd))
    ms_topk = timed(lambda: fused.topk_abs(r, k))
    idx, _ = fused.topk_abs(r, k)
    t0 = time.perf_counter()
    host = r.detach().cpu().numpy().ravel()
    ref_idx = np.argsort(-1.0 * np.abs(host))[:k]
    ms_host = (time.perf_counter() - t0) * 1e3
    same = bool(np.array_equal(np.sort(ref_idx), np.sort(idx.cpu().numpy())))
    # top-k reads the column once per radix pass that still has live candidates (8 passes + gather)
    print(json.dumps({
        "n_points": n, "k": k, "net": "MLP 2-128x4-1 tanh (Allen-Cahn), FP32",
        "residual_forward_ms": ms_res, "residual_forward_points_per_s": n / ms_res * 1e3,
        "topk_ms": ms_topk, "topk_GBps_algorithmic(9 reads of the column)": 9 * 4 * n / ms_topk / 1e6,
        "reference_recipe_d2h_argsort_ms": ms_host, "topk_speedup_vs_host": ms_host / ms_topk,
        "same_selection_as_numpy": same}))


if __name__ == "__main__":
    main()
