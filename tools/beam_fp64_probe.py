"""Where does the FP64 LBFGS beam separate between CPU and GPU?  Prints the initial loss (no optimiser step) and the loss
after one train_pipe with 1, 2, 4, 8, 20 LBFGS inner iterations, in FP64, on the active device.
    IDRLNET_B200_DEVICE=cpu python tools/beam_fp64_probe.py     (CPU)      python tools/beam_fp64_probe.py   (GPU)"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import idrlnet_b200 as pkg  # noqa: E402
import idrlnet_b200.shortcut as sc  # noqa: E402
from trace_case import CASES  # noqa: E402

os.chdir(tempfile.mkdtemp())
out = {"device": str(pkg.DEVICE)}
for inner in (0, 1, 2, 4, 8, 20):
    net, kw = CASES["beam_lbfgs"](sc, tempfile.mkdtemp(), 12)
    net.net.double()
    torch.set_default_dtype(torch.float64)
    try:
        if inner:
            kw["opt_config"] = dict(optimizer="LBFGS", lr=1, max_iter=inner)
        s = sc.Solver(**kw)
        if inner == 0:
            out["init"] = float(s._loss_and_grads(grads=False))
        else:
            out[f"inner{inner}"] = float(s.train_pipe())
    finally:
        torch.set_default_dtype(torch.float32)
print({k: (v if isinstance(v, str) else float(np.float64(v))) for k, v in out.items()})
for k, v in out.items():
    if k != "device":
        print(k, repr(v))

# PyTorch's fused CUDA weight-norm kernel vs the composite formula, in double (suspected: the kernel takes the row norm
# with sqrtf, i.e. in single precision, whatever the tensor dtype)
if pkg.DEVICE.type == "cuda":
    torch.manual_seed(0)
    v = torch.randn(20, 20, dtype=torch.float64)
    g = torch.rand(20, 1, dtype=torch.float64) + 0.5
    ref = v * (g / torch.norm_except_dim(v, 2, 0))
    fused_cuda = torch._weight_norm(v.cuda(), g.cuda(), 0).cpu()
    comp_cuda = (v.cuda() * (g.cuda() / torch.norm_except_dim(v.cuda(), 2, 0))).cpu()
    print("torch._weight_norm cuda float64 vs CPU composite: max rel", float(((fused_cuda - ref).abs() / ref.abs()).max()))
    print("composite formula on cuda float64 vs CPU composite: max rel", float(((comp_cuda - ref).abs() / ref.abs()).max()))
