"""Short run of a bench workload (default Burgers, 2^20 interior points) for ncu captures."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
config = sys.argv[2] if len(sys.argv) > 2 else "burgers"
precision = sys.argv[3] if len(sys.argv) > 3 else "fp32"
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
tr = bench.Trainer(bench.workload_configs()[config], dev, precision=precision)
rng = np.random.default_rng(0)
for _ in range(2):
    hb = tr.sample_host(rng)
    tr.bind_set({k: torch.as_tensor(v, device=dev) for k, v in hb.items()})
for i in range(steps):
    loss = tr.step(i % 2)
torch.cuda.synchronize()
print("loss", float(loss))
