"""Seeded training trace of the REAL reference (idrlnet) on the simple_poisson example: initial parameters and the
loss `Solver.train_pipe()` returns for 12 iterations (the post-update, re-sampled loss, idrlnet/solver.py:341-344).
Run in the build container only:   python tests/golden/make_trace_golden.py  ->  tests/golden/train_trace.npz"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import _install_shims  # noqa: E402
from trace_case import build  # noqa: E402


def main():
    _install_shims()
    import idrlnet.shortcut as sc
    net, kw = build(sc, tempfile.mkdtemp(), 12)
    state = {k: v.detach().numpy().copy() for k, v in net.net.state_dict().items()}
    s = sc.Solver(**kw)
    losses = [float(s.train_pipe()) for _ in range(12)]
    np.savez(os.path.join(HERE, "train_trace.npz"), losses=np.array(losses), **{"state/" + k: v for k, v in state.items()})
    print(losses)


if __name__ == "__main__":
    main()
