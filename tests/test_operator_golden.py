"""ICNode / Int1DNode (generic torch path) against outputs of the reference's own
classes on the same inputs and net weights (tests/golden/operators.npz)."""
import os

import numpy as np
import pytest

import idrlnet_b200.shortcut as sc
from tests.golden.operator_cases import CASES

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "operators.npz"))


@pytest.fixture
def on_cpu():
    import idrlnet_b200 as pkg
    was_gpu, dev = pkg.GPU_ENABLED, pkg.DEVICE
    pkg.use_cpu()
    yield
    if was_gpu:
        pkg.use_gpu(dev)


@pytest.mark.parametrize("name", sorted(CASES))
def test_operator_matches_reference(name, on_cpu):
    got = CASES[name](sc)
    for k, v in got.items():
        ref = GOLD[f"{name}/{k}"]
        if ref.dtype.kind in "US":
            assert list(ref) == list(v), k
        else:
            np.testing.assert_allclose(np.asarray(v), ref, rtol=2e-6, atol=1e-6, err_msg=k)
