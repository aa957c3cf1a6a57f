"""Seeded training traces of the REAL reference (idrlnet) on three cut-down examples (tests/golden/trace_case.py): initial parameters and the
loss `Solver.train_pipe()` returns for 12 iterations (the post-update, re-sampled loss, idrlnet/solver.py:341-344).
Run in the build container only:   python tests/golden/make_trace_golden.py  ->  tests/golden/train_trace.npz"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _install_shims  # noqa: E402
from trace_case import CASES, INFER, INFER_SEED  # noqa: E402


def main():
    _install_shims()
    import idrlnet.shortcut as sc
    out = {}
    for name, case in CASES.items():
        net, kw = case(sc, tempfile.mkdtemp(), 12)
        for k, v in net.net.state_dict().items():
            out[f"{name}/state/{k}"] = v.detach().numpy().copy()
        s = sc.Solver(**kw)
        out[f"{name}/losses"] = np.array([float(s.train_pipe()) for _ in range(12)])
        print(name, out[f"{name}/losses"])
        if name in INFER:  # Solver.infer_step (idrlnet/solver.py:410-420) on freshly drawn points of one domain
            np.random.seed(INFER_SEED)
            dom, cols = INFER[name]
            res = s.infer_step({dom: cols})[dom]
            for c in cols:
                out[f"{name}/infer/{c}"] = res[c].detach().cpu().numpy()
    np.savez(os.path.join(HERE, "train_trace.npz"), **out)


def fp64_traces(names=("beam_lbfgs",)):
    """FP64 tie-breaker traces: the same seeded case with the net cast to double and torch's default dtype
    float64 (so `torch.Tensor(val)` in idrlnet/variable.py:109 keeps the NumPy points in double).  Appended to
    train_trace.npz as `<case>/losses64`.  For the LBFGS beam the reference's OWN FP32 and FP64 traces separate by
    5e-6, 2e-4, 1.6e-2, 2.3 ... per outer iteration: the case is chaotic in FP32, which is why the GPU FP32 run is
    only compared for its first outer iterations while the FP64 run is compared for all twelve."""
    _install_shims()
    import torch
    import idrlnet.shortcut as sc
    path = os.path.join(HERE, "train_trace.npz")
    out = dict(np.load(path))
    for name in names:
        net, kw = CASES[name](sc, tempfile.mkdtemp(), 12)
        net.net.double()
        torch.set_default_dtype(torch.float64)
        try:
            s = sc.Solver(**kw)
            out[f"{name}/losses64"] = np.array([float(s.train_pipe()) for _ in range(12)])
        finally:
            torch.set_default_dtype(torch.float32)
        print(name, "fp64", out[f"{name}/losses64"])
    np.savez(path, **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--fp64":
        fp64_traces(tuple(sys.argv[2:]) or ("beam_lbfgs",))
    else:
        main()
        fp64_traces()
