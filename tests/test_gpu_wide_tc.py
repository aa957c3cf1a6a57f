"""Wide-net layer GEMMs on tcgen05 tensor cores (TF32 multiplicands, FP32 accumulate).
Stated tolerance: 1e-2 norm-wise on losses and gradients against the FP64 oracle
(observed 1e-4 ... 5e-3: TF32 keeps 10 mantissa bits of every multiplicand, and the jets pass
through up to 4 layers and second derivatives); the FP32 parity path of
tests/test_gpu_wide.py holds 1e-5."""
import numpy as np
import pytest
import torch

from oracle import pinn_oracle as po
from tests import golden_util as gu
from tests import problem_util as pu
from tests.test_emu_narrow import _check_grads
from tests.test_gpu_wide import WIDE_SHAPES, _without

pytestmark = pytest.mark.gpu
TOL_TF32 = 1e-2 / 20  # _check_grads allows 20 * tol on gradients


PRECISIONS = ["tf32", "tf32_streamed"]  # on-chip fused kernel (where it applies) / layer-streamed kernels


def _compare(problem, tol=TOL_TF32, precision="tf32"):
    ref = po.training_step(problem, dtype=torch.float64)
    grads, losses, skipped, _ = pu.run_cuda_fused(problem, precision=precision)
    assert not skipped
    worst = 0.0
    for dname, l in losses.items():
        r = ref["loss_components"][dname]
        worst = max(worst, abs(l - r) / max(abs(r), 1e-30))
    assert worst <= 1e-2, worst
    _check_grads(ref, grads, problem, tol)
    # and it must actually differ from the FP32 path (i.e. the tensor-core kernels ran)
    g32, _, _, _ = pu.run_cuda_fused(problem, precision="fp32")
    assert any(not np.array_equal(g32[k], grads[k]) for k in grads)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("name", ["ldc", "ns_unsteady_wide"])
def test_golden_tf32(name, precision):
    problem, _ = gu.load(name)
    _compare(problem, precision=precision)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_golden_allen_cahn_tf32(precision):
    problem, _ = gu.load("allen_cahn")
    _compare(_without(problem, ["AllenBc"]), precision=precision)


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("case", range(len(WIDE_SHAPES)))
def test_random_wide_tf32(case, precision):
    inputs, outputs, eq, act, width, n_hidden, n = WIDE_SHAPES[case]
    rng = np.random.default_rng(200 + case)
    problem = pu.random_problem(rng, n_in=len(inputs), n_out=len(outputs), width=width, n_hidden=n_hidden, act=act,
                                n_points=n, equation=eq, inputs=inputs, outputs=outputs, weight_norm=(case % 2 == 0),
                                with_lambda=(case % 3 == 0), sigma=1.0 + 0.5 * case, target=0.25 if case % 2 else 0.0)
    if "normal_x" in eq:
        d = problem["domains"][0]
        d["points"]["normal_x"] = rng.standard_normal((n, 1)).astype(np.float32)
        d["points"]["normal_y"] = rng.standard_normal((n, 1)).astype(np.float32)
    if n_hidden == 1:
        pytest.skip("single hidden layer: no hidden-to-hidden GEMM, nothing runs on the tensor cores")
    _compare(problem, precision=precision)


def test_fused_many_tiles_and_ragged_tail():
    """On-chip kernel: more tiles than CTAs (persistent loop, TMEM dW accumulation across tiles)
    and a last tile that is only partly filled."""
    rng = np.random.default_rng(31)
    problem = pu.random_problem(rng, n_in=2, n_out=3, inputs=("x", "y"), outputs=("u", "v", "p"), width=100, n_hidden=4,
                                act="tanh", n_points=148 * 16 * 3 + 5,
                                equation="p__x + u*u__x + v*u__y - 0.01*u__x__x - 0.01*u__y__y", with_lambda=True)
    _compare(problem, precision="tf32")


def test_multi_chunk_tf32():
    rng = np.random.default_rng(9)
    problem = pu.random_problem(rng, n_in=3, n_out=1, inputs=("x", "y", "t"), outputs=("u",), width=512, n_hidden=2,
                                act="tanh", n_points=4100, equation="u__x__x + u__y__y + u__t__t + u*u__x")
    _compare(problem)
