"""Checkpoint interchange with the REAL reference (idrlnet/solver.py:425-460): the reference trains the cut-down Burgers case
(tests/golden/trace_case.py) for 6 iterations and saves `model.ckpt`; a second reference Solver loads it and trains 6 more.
Fixtures: tests/golden/ref_burgers_step6.ckpt (the file as the reference wrote it) and tests/golden/ckpt_trace.npz (the
losses of the resumed run).  The reverse direction -- a checkpoint written by idrlnet_b200 loaded by the reference -- can
only be exercised where the reference exists, so it is asserted here, at generation time.
Run in the build container only:   python tests/golden/make_ckpt_golden.py"""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_golden import _install_shims  # noqa: E402
from trace_case import CASES  # noqa: E402

RESUME_SEED = 101


def resumed_losses(sc, network_dir, n=6, **extra):
    _, kw = CASES["burgers"](sc, network_dir, 12)
    kw["loading"] = True
    s = sc.Solver(**kw, **extra)
    step0 = s.global_step
    np.random.seed(RESUME_SEED)
    return step0, np.array([float(s.train_pipe()) for _ in range(n)])


def main():
    _install_shims()
    import idrlnet.shortcut as ref
    d = tempfile.mkdtemp()
    _, kw = CASES["burgers"](ref, d, 12)
    s = ref.Solver(**kw)
    for _ in range(6):
        s.train_pipe()
    s.save()
    shutil.copy(os.path.join(d, "model.ckpt"), os.path.join(HERE, "ref_burgers_step6.ckpt"))
    step0, losses = resumed_losses(ref, d)
    assert step0 == s.global_step
    print("reference resumed at", step0, losses)
    np.savez(os.path.join(HERE, "ckpt_trace.npz"), global_step=step0, losses=losses)

    # reverse direction: ours writes, the reference reads and continues with the same losses
    import idrlnet_b200 as pkg
    import idrlnet_b200.shortcut as ours
    pkg.use_cpu()
    d2 = tempfile.mkdtemp()
    _, kw = CASES["burgers"](ours, d2, 12)
    s2 = ours.Solver(post_step_loss=True, **kw)
    for _ in range(6):
        s2.train_pipe()
    s2.save()
    step1, losses1 = resumed_losses(ref, d2)
    rel = np.abs(losses1 - losses) / np.abs(losses)
    print("reference resumed from our checkpoint at", step1, "max rel diff", rel.max())
    assert step1 == step0 and rel.max() < 5e-6


if __name__ == "__main__":
    main()
