"""Short run of the bench workload (Burgers, 2^20 interior points) for ncu captures."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.cuda.set_device(0)
dev = torch.device("cuda:0")
tr = bench.BurgersTrainer(dev)
rng = np.random.default_rng(0)
for _ in range(2):
    hb = bench.make_batch_host(rng, bench.N_INTERIOR, bench.N_IC, bench.N_BC)
    tr.bind_set({k: torch.as_tensor(hb[k], device=dev) for k in bench.BATCH_KEYS})
for i in range(steps):
    loss = tr.step(i % 2)
torch.cuda.synchronize()
print("loss", float(loss))
