"""End-to-end behavioural parity with the REAL reference: same seeds -> same initial
parameters (architecture init), same sampled points (geometry RNG stream), same per-iteration losses from
`Solver.train_pipe()` (graph, loss, Adam + ExponentialLR, post-update re-evaluation).  Golden:
tests/golden/train_trace.npz, produced by tests/golden/make_trace_golden.py from /root/reference."""
import os

import numpy as np
import pytest
import torch

import idrlnet_b200 as pkg
import idrlnet_b200.shortcut as sc
from tests.golden.trace_case import CASES, INFER, INFER_SEED

# LBFGS (no line search, 20 inner iterations per train_pipe) on the fourth-order beam is chaotic in FP32: the
# reference's OWN FP32 and FP64 traces (train_trace.npz `beam_lbfgs/losses` vs `beam_lbfgs/losses64`, both produced by
# the real reference on the CPU) separate by 5e-6, 2e-4, 1.6e-2, 2.3 ... per outer iteration -- asserted in
# test_lbfgs_beam_is_chaotic_in_fp32_for_the_reference_itself.  A GPU FP32 run (other math libraries, jets from the
# kernel) separates from the CPU FP32 golden at the same rate (observed on a B200: 3e-5, 2.6e-4, 1.5e-2, 0.7 ...), so in
# FP32 only the first two outer iterations are compared there; the tie-breaker is the FP64 run, compared for all twelve
# iterations on the CPU tier and until FP64 rounding itself has been amplified to O(1) on the GPU (test_lbfgs_beam_fp64_*).
GPU_TOL = {"beam_lbfgs": 1e-3}
GPU_STEPS = {"beam_lbfgs": 2}
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "train_trace.npz"))


def _run(name, tmp_path, fused, infer_tol=2e-5):
    net, kw = CASES[name](sc, str(tmp_path), 12)
    prefix = f"{name}/state/"
    for k in GOLD.files:
        if k.startswith(prefix):  # the same torch seed gives the reference's initial parameters
            np.testing.assert_allclose(net.net.state_dict()[k[len(prefix):]].cpu().numpy(), GOLD[k], rtol=0, atol=0)
    s = sc.Solver(fused=fused, **kw)  # defaults = the reference's train_pipe semantics (post-update re-evaluation)
    assert s.post_step_loss
    losses = np.array([float(s.train_pipe()) for _ in range(12)])
    if name in INFER:  # derivative / residual columns of `infer_step` on the trained net, same fresh points
        np.random.seed(INFER_SEED)
        dom, cols = INFER[name]
        res = s.infer_step({dom: cols})[dom]
        for c in cols:
            ref = GOLD[f"{name}/infer/{c}"]
            got = res[c].detach().cpu().numpy()
            assert got.shape == ref.shape, (c, got.shape, ref.shape)
            err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
            assert err <= infer_tol, (name, c, err)
    return losses


@pytest.mark.parametrize("name", sorted(CASES))
def test_generic_path_reproduces_the_reference_trace(name, tmp_path):
    was_gpu, dev = pkg.GPU_ENABLED, pkg.DEVICE
    pkg.use_cpu()
    try:
        losses = _run(name, tmp_path, fused=None)
    finally:
        if was_gpu:
            pkg.use_gpu(dev)
    ref = GOLD[f"{name}/losses"]
    rel = np.abs(losses - ref) / np.abs(ref)
    assert rel.max() <= 5e-6, (rel, losses, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["poisson", "burgers", "cavity"])   # every domain of these runs on the fused kernels
def test_fused_path_reproduces_the_reference_trace(name, tmp_path):
    losses = _run(name, tmp_path, fused=True, infer_tol=2e-3)
    ref = GOLD[f"{name}/losses"]
    rel = np.abs(losses - ref) / np.abs(ref)
    assert rel.max() <= 5e-4, (rel, losses, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_gpu_auto_path_reproduces_the_reference_trace(name, tmp_path):
    """Every case on the GPU the way a switching user runs it (`fused=None`): domains that compile run the fused kernels,
    the others the generic graph with value / derivative columns served by the jet kernel where the net is supported."""
    losses = _run(name, tmp_path, fused=None, infer_tol=2e-3)
    ref = GOLD[f"{name}/losses"]
    rel = (np.abs(losses - ref) / np.abs(ref))[:GPU_STEPS.get(name, 12)]
    print(name, "max rel", rel.max())
    assert rel.max() <= GPU_TOL.get(name, 5e-4), (rel, losses, ref)


def _run_fp64(name, tmp_path, exact_weight_norm=False):
    """The reference's FP64 recipe (tests/golden/make_trace_golden.py::fp64_traces) on this package.

    `exact_weight_norm`: PyTorch's fused CUDA kernel behind `torch._weight_norm` takes the row norm in SINGLE precision
    whatever the tensor dtype (tools/beam_fp64_probe.py on a B200: 6.0e-8 relative against the composite formula
    v * (g / ||v||) that the CPU implementation uses, which the same GPU reproduces to 3e-16) -- a float32-sized
    perturbation that this chaotic case amplifies to 2e-5 after one LBFGS outer iteration.  A tie-breaker run in FP64
    must not contain it, so on the GPU the weight-norm hook is pointed at the composite formula."""
    import sys
    wn_mod = sys.modules["torch.nn.utils.weight_norm"]  # (the attribute of that name on torch.nn.utils is the function)
    real = wn_mod._weight_norm
    if exact_weight_norm:
        wn_mod._weight_norm = lambda v, g, dim=0: v * (g / torch.norm_except_dim(v, 2, dim))
    net, kw = CASES[name](sc, str(tmp_path), 12)
    net.net.double()
    torch.set_default_dtype(torch.float64)
    try:
        s = sc.Solver(**kw)
        assert not s._fused_domains  # the kernels are FP32: an FP64 run takes the torch path everywhere
        return np.array([float(s.train_pipe()) for _ in range(12)])
    finally:
        torch.set_default_dtype(torch.float32)
        wn_mod._weight_norm = real


def test_lbfgs_beam_is_chaotic_in_fp32_for_the_reference_itself():
    """Both traces come from the real reference on the CPU; only the dtype differs."""
    rel = np.abs(GOLD["beam_lbfgs/losses"] - GOLD["beam_lbfgs/losses64"]) / np.abs(GOLD["beam_lbfgs/losses64"])
    assert rel[0] < 1e-5 and rel[1] < 1e-3 and rel[2] > 1e-3 and rel[3] > 0.5, rel


def test_lbfgs_beam_fp64_cpu_matches_the_reference_for_all_iterations(tmp_path):
    was_gpu, dev = pkg.GPU_ENABLED, pkg.DEVICE
    pkg.use_cpu()
    try:
        losses = _run_fp64("beam_lbfgs", tmp_path)
    finally:
        if was_gpu:
            pkg.use_gpu(dev)
    ref = GOLD["beam_lbfgs/losses64"]
    assert np.max(np.abs(losses - ref) / np.abs(ref)) <= 1e-9, (losses, ref)


@pytest.mark.gpu
def test_lbfgs_beam_fp64_gpu_matches_the_reference(tmp_path):
    """the same generic code on the GPU in FP64 reproduces the reference's FP64 CPU trace to rounding
    (1.5e-14 after the first outer iteration = 20 closure evaluations) -- the FP32 separation is rounding chaos, not a bug."""
    assert pkg.DEVICE.type == "cuda"
    losses = _run_fp64("beam_lbfgs", tmp_path, exact_weight_norm=True)
    ref = GOLD["beam_lbfgs/losses64"]
    rel = np.abs(losses - ref) / np.abs(ref)
    print("beam_lbfgs fp64 gpu rel per iteration", rel)
    # Measured on a B200: 1.5e-14, 1.8e-13, 1.6e-11, 2.9e-10, 1.7e-08, 1.1e-04, 2.3e-03, 3.9e-02 ... -- agreement to FP64
    # rounding after the first outer iteration, then the case's own amplification (x10 .. x6000 per outer iteration; a
    # 1e-15 perturbation of the parameters on the CPU alone gives 1.4e-13, 1.3e-12, 2.7e-11, 5.6e-10 after iterations
    # 1-4) until even FP64 rounding reaches O(1) around iteration 11.  So: five outer iterations (100 LBFGS inner
    # iterations) are held to 1e-6 -- against two in FP32 -- and the rest must stay a converging run.
    assert rel[:5].max() <= 1e-6, (rel, losses, ref)
    assert rel[0] <= 1e-12
    assert np.isfinite(losses).all() and losses[-1] < 1e-3 * losses[0]


def test_reference_checkpoint_resumes_with_the_reference_losses(tmp_path):
    """`model.ckpt` written by the REAL reference after 6 Burgers iterations (tests/golden/make_ckpt_golden.py) is loaded by
    `Solver(loading=True)` -- parameters, Adam moments, global_step (idrlnet/solver.py:425-460) -- and the next 6
    `train_pipe()` losses equal those of the reference's own resumed run.  (The generator also asserts the reverse
    direction: the reference resumes from a checkpoint this package wrote with identical losses.)"""
    import shutil
    RESUME_SEED = 101  # tests/golden/make_ckpt_golden.py
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ckpt_trace.npz"))
    shutil.copy(os.path.join(os.path.dirname(__file__), "golden", "ref_burgers_step6.ckpt"), str(tmp_path / "model.ckpt"))
    was_gpu, dev = pkg.GPU_ENABLED, pkg.DEVICE
    pkg.use_cpu()
    try:
        _, kw = CASES["burgers"](sc, str(tmp_path), 12)
        kw["loading"] = True
        s = sc.Solver(**kw)
        assert s.global_step == int(gold["global_step"]) == 6
        np.random.seed(RESUME_SEED)
        losses = np.array([float(s.train_pipe()) for _ in range(6)])
    finally:
        if was_gpu:
            pkg.use_gpu(dev)
    rel = np.abs(losses - gold["losses"]) / np.abs(gold["losses"])
    assert rel.max() <= 5e-6, (rel, losses, gold["losses"])
