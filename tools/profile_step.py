"""Short run of a bench workload (default Burgers, 2^20 interior points) through Solver.train_pipe() for ncu captures.
    python tools/profile_step.py [steps] [config] [precision] [size factor]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from baseline import workloads as wl  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
config = sys.argv[2] if len(sys.argv) > 2 else "burgers"
precision = sys.argv[3] if len(sys.argv) > 3 else None
factor = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
torch.cuda.set_device(0)
import idrlnet_b200 as pkg  # noqa: E402
pkg.use_gpu(0)
cfg = wl.CONFIGS[config]
sizes = wl.scaled_sizes(cfg["sizes"](1), factor) if factor != 1.0 else None
run = bench.OursRun(config, torch.device("cuda:0"), 1, 0, precision=precision if cfg["wide"] else None, sizes=sizes)
for i in range(steps):
    loss = run.step(i % run.n_rotate)
torch.cuda.synchronize()
print("loss", float(loss), "fused domains", run.fused_domains)
