"""Observed parity of the CUDA step against the FP64 oracle (reference algorithm) on the golden
configs, per precision mode: relative loss error, norm-wise error of the whole gradient vector
(max|a-b| / max|ref|, the metric of SURVEY.md 8(c3)) and the worst per-layer block.

    python tools/parity_report.py > profiles/r01/parity_report.txt      (needs a B200)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pinn_oracle as po  # noqa: E402
from tests import golden_util as gu  # noqa: E402
from tests import problem_util as pu  # noqa: E402
from idrlnet_b200 import compiler as cp  # noqa: E402

CONFIGS = {"simple_poisson": ["fp32"], "burgers": ["fp32"], "ns_unsteady": ["fp32"],
           "ldc": ["fp32", "fp32x3", "tf32", "tf32_streamed"],
           "allen_cahn": ["fp32", "fp32x3", "tf32", "tf32_streamed"],
           "ns_unsteady_wide": ["fp32", "fp32x3", "tf32_streamed"]}


def main():
    print(f"{'config':18s} {'precision':14s} {'loss rel err':>13s} {'grad normwise':>14s} {'worst block':>12s}")
    for name, precisions in CONFIGS.items():
        problem, _ = gu.load(name)
        fusable = []
        for d in problem["domains"]:
            try:
                pu.compile_problem_domain(problem, d)
                fusable.append(d)
            except cp.NotFusable:
                pass
        problem = dict(problem, domains=fusable)
        ref = po.training_step(problem, dtype=torch.float64)
        for prec in precisions:
            grads, losses, skipped, _ = pu.run_cuda_fused(problem, precision=prec)
            lerr = max(abs(l - ref["loss_components"][k]) / max(abs(ref["loss_components"][k]), 1e-30) for k, l in losses.items())
            gerr, berr = 0.0, 0.0
            for net, flat in grads.items():
                parts = pu.flat_grad_from_oracle(ref, net)
                sizes = pu.flat_layout_sizes(pu.effective_layers(pu.owner_net(problem, net)))
                full = np.concatenate([np.zeros(sz) if p is None else np.asarray(p).ravel() for p, sz in zip(parts, sizes)])
                scale = np.abs(full).max()
                gerr = max(gerr, np.abs(flat - full).max() / scale)
                off = 0
                for sz in sizes:
                    blk = full[off:off + sz]
                    berr = max(berr, np.abs(flat[off:off + sz] - blk).max() / max(np.abs(blk).max(), 1e-3 * scale))
                    off += sz
            print(f"{name:18s} {prec:14s} {lerr:13.2e} {gerr:14.2e} {berr:12.2e}")


if __name__ == "__main__":
    main()
