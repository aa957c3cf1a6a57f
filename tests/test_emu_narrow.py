"""CPU-tier check of the narrow warp programs: the same per-lane phase functions
the CUDA kernel runs (idrlnet_b200/csrc/narrow_warp.h), executed lane by lane by
tests/emu, against the oracle and the golden vectors of the reference."""
import numpy as np
import pytest
import torch

from oracle import pinn_oracle as po
from tests import golden_util as gu
from tests import problem_util as pu

TOL = 1e-5  # norm-wise, FP32 (BASELINE.json north_star)


def _compare(problem, n_cta=3, n_warps=2, tol=TOL, expect_skipped=()):
    ref = po.training_step(problem, dtype=torch.float64)
    grads, losses, skipped = pu.run_emu_fused(problem, n_cta, n_warps)
    assert sorted(skipped) == sorted(expect_skipped)
    for dname, l in losses.items():
        r = ref["loss_components"][dname]
        assert abs(l - r) <= 10 * tol * max(abs(r), 1e-30), (dname, l, r)
    return ref, grads, losses


def _check_grads(ref, grads, problem, tol=TOL):
    for net, flat in grads.items():
        parts = pu.flat_grad_from_oracle(ref, net)
        layers = pu.effective_layers(pu.owner_net(problem, net))
        sizes = pu.flat_layout_sizes(layers)
        full_ref = np.concatenate([np.zeros(sz) if p is None else p for p, sz in zip(parts, sizes)])
        scale = np.max(np.abs(full_ref))
        off = 0
        for p, sz in zip(parts, sizes):
            got = flat[off:off + sz]
            want = np.zeros(sz) if p is None else p
            err = np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-3 * scale, 1e-30)
            assert err <= 20 * tol, (net, off, err)
            off += sz


@pytest.mark.parametrize("name", ["burgers", "ns_unsteady", "simple_poisson"])
def test_emu_matches_oracle_on_golden(name):
    """Configs whose every domain is fusable: losses and the total gradient."""
    problem, flat = gu.load(name)
    ref, grads, losses = _compare(problem)
    _check_grads(ref, grads, problem)
    # and the reference's own FP32 numbers (golden) for the loss components
    for k, v in gu.ref_components(flat, "ref32").items():
        assert abs(losses[k[:-5]] - v) <= 1e-4 * abs(v)


SHAPES = [
    # (inputs, outputs, equation, act, width, n_hidden, n_points)
    (("x", "t"), ("u",), "u*u__x + u__t - 0.01*u__x__x", "tanh", 20, 4, 77),
    (("x", "t"), ("u",), "5*u**3 - 5*u + u__t - 0.0001*u__x__x", "silu", 20, 4, 64),
    (("x", "y"), ("T",), "-T__x__x - T__y__y", "silu", 20, 4, 33),
    (("x", "y"), ("T",), "normal_x*T__x + normal_y*T__y", "tanh", 20, 3, 40),
    (("x", "y"), ("T",), "T", "sin", 20, 2, 31),
    (("x", "y"), ("u", "v", "p"), "p__x + u*u__x + v*u__y - 0.01*u__x__x - 0.01*u__y__y", "tanh", 20, 4, 50),
    (("x", "y", "t"), ("u", "v", "p"), "u__t + p__y + u*v__x + v*v__y - 0.01*v__x__x - 0.01*v__y__y", "silu", 20, 2, 45),
    (("x", "y", "t"), ("u",), "u__x__x + u__y__y + u__t__t - x*y*u", "sin", 20, 1, 36),
    (("x",), ("u",), "u__x__x + u*u__x", "tanh", 12, 2, 20),
    (("x", "t"), ("u", "v"), "u__t + 0.5*v__x__x + (u**2 + v**2)*v", "tanh", 32, 3, 70),
    (("x", "y", "t"), ("u",), "u__x + u__y + u__t", "silu", 32, 2, 40),
    (("x", "t"), ("u",), "u__t - u__x", "tanh", 8, 5, 100),
    # the mixed second derivative, and the pure third / fourth derivative along one input
    (("x", "y"), ("u",), "u__x__y + u*u__x - 0.3*u__y", "tanh", 20, 3, 45),
    (("x", "y"), ("u", "v"), "u__x__x + u__y__y + 0.5*v__x__y - v*u__y__x", "silu", 20, 3, 40),
    (("x", "y"), ("u",), "u__y__x*u__x__y + u", "sin", 32, 2, 38),
    (("x", "y"), ("u",), "u__x__y + u__x__x - u", "silu", 20, 4, 50),
    (("x",), ("y",), "y__x__x__x + y*y__x__x - y__x", "tanh", 20, 3, 41),
    (("x",), ("y",), "y__x__x__x__x + 1", "silu", 20, 4, 40),
    (("x",), ("y", "w"), "y__x__x__x__x + w__x__x__x*y - 0.1*w__x", "sin", 20, 2, 35),
    (("x",), ("y",), "y__x__x__x__x*y - y__x__x", "tanh", 32, 3, 33),
]


@pytest.mark.parametrize("case", range(len(SHAPES)))
def test_emu_random_problems(case):
    inputs, outputs, eq, act, width, n_hidden, n = SHAPES[case]
    rng = np.random.default_rng(100 + case)
    problem = pu.random_problem(rng, n_in=len(inputs), n_out=len(outputs), width=width, n_hidden=n_hidden, act=act,
                                n_points=n, equation=eq, inputs=inputs, outputs=outputs,
                                weight_norm=(case % 2 == 0), with_lambda=(case % 3 == 0), sigma=1.0 + 0.5 * case,
                                target=0.25 if case % 2 else 0.0)
    if "normal_x" in eq:
        d = problem["domains"][0]
        d["points"]["normal_x"] = rng.standard_normal((n, 1)).astype(np.float32)
        d["points"]["normal_y"] = rng.standard_normal((n, 1)).astype(np.float32)
    ref, grads, losses = _compare(problem, n_cta=2, n_warps=3)
    _check_grads(ref, grads, problem)


@pytest.mark.parametrize("loss", ["L1", "Identity"])
def test_emu_other_losses(loss):
    rng = np.random.default_rng(7)
    problem = pu.random_problem(rng, equation="u__x - 0.3*u", loss=loss, n_points=50, target=0.1, with_lambda=True)
    ref, grads, losses = _compare(problem)
    _check_grads(ref, grads, problem)


def test_emu_empty_and_tiny_domains():
    rng = np.random.default_rng(3)
    for n in (0, 1, 32, 33):
        problem = pu.random_problem(rng, n_points=n)
        grads, losses, _ = pu.run_emu_fused(problem)
        if n == 0:
            assert losses["dom"] == 0.0 and not np.any(grads["net"])
        else:
            ref = po.training_step(problem, dtype=torch.float64)
            _check_grads(ref, grads, problem)


def test_emu_loss_only_mode():
    """MODE_LOSS (pinn_loss_fused): the same losses as the full step, no gradient."""
    problem, _ = gu.load("burgers")
    g_full, l_full, _ = pu.run_emu_fused(problem)
    g_loss, l_loss, _ = pu.run_emu_fused(problem, mode=3)
    assert l_loss == l_full
    assert all(np.abs(g).max() > 0 for g in g_full.values()) and all(np.abs(g).max() == 0 for g in g_loss.values())
