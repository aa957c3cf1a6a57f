"""Full fused step vs the loss-only evaluation (pinn_loss_fused) that ends every reference-semantics train_pipe."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

for config, prec in (("burgers", None), ("ldc", "tf32"), ("ldc", "fp32x3")):
    dev = torch.device("cuda:0")
    run = bench.OursRun(config, dev, 1, 0, precision=prec)
    run.step(0)
    eng = next(iter(run.solver._engines.values()))

    def t(fn, reps=20):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    print(config, prec, "full %.3f ms  loss-only %.3f ms" % (t(lambda: eng.launch()), t(lambda: eng.launch(loss_only=True))))
    run.close()
