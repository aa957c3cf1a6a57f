"""GPU parity of the narrow fused step, called through the C ABI
(idrlnet_b200.fused -> libpinn_b200.so), against the CPU oracle and the golden
vectors produced by the reference."""
import numpy as np
import pytest
import torch

from oracle import pinn_oracle as po
from tests import golden_util as gu
from tests import problem_util as pu
from tests.test_emu_narrow import SHAPES, TOL, _check_grads

pytestmark = pytest.mark.gpu


def _compare_cuda(problem, tol=TOL, expect_skipped=()):
    ref = po.training_step(problem, dtype=torch.float64)
    grads, losses, skipped, modules = pu.run_cuda_fused(problem)
    assert sorted(skipped) == sorted(expect_skipped)
    for dname, l in losses.items():
        r = ref["loss_components"][dname]
        assert abs(l - r) <= 10 * tol * max(abs(r), 1e-30), (dname, l, r)
    _check_grads(ref, grads, problem, tol)
    return ref, grads, losses, modules


@pytest.mark.parametrize("name", ["burgers", "ns_unsteady", "simple_poisson"])
def test_golden_configs(name):
    problem, flat = gu.load(name)
    ref, grads, losses, modules = _compare_cuda(problem)
    # loss components against the reference's own FP32 run
    for k, v in gu.ref_components(flat, "ref32").items():
        assert abs(losses[k[:-5]] - v) <= 1e-4 * abs(v)
    # raw parameter gradients (weight_g / weight_v / bias, i.e. through the weight-norm
    # backward kernel) against the reference's FP64 run
    for pname, g in gu.ref_grads(flat, "ref64").items():
        net, rest = pname.split("/", 1)
        p = dict(modules[net].named_parameters())[rest]
        scale = max(np.max(np.abs(g)), 1e-30)
        err = np.max(np.abs(p.grad.detach().cpu().numpy().reshape(g.shape) - g)) / scale
        assert err <= 20 * TOL, (pname, err)


@pytest.mark.parametrize("case", range(len(SHAPES)))
def test_random_problems(case):
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
    _compare_cuda(problem)


@pytest.mark.parametrize("loss", ["L1", "Identity"])
def test_other_losses(loss):
    rng = np.random.default_rng(7)
    problem = pu.random_problem(rng, equation="u__x - 0.3*u", loss=loss, n_points=50, target=0.1, with_lambda=True)
    _compare_cuda(problem)


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 4097])
def test_ragged_sizes(n):
    rng = np.random.default_rng(n)
    problem = pu.random_problem(rng, n_points=n, act="silu", weight_norm=True)
    if n == 0:
        grads, losses, _, _ = pu.run_cuda_fused(problem)
        assert losses["dom"] == 0.0 and not np.any(grads["net"])
    else:
        _compare_cuda(problem)


def test_oracle_sized_burgers_64k():
    """Burgers shape (BASELINE configs[1]) at 2^16 points against the oracle."""
    rng = np.random.default_rng(11)
    problem = pu.random_problem(rng, n_points=1 << 16, act="silu", weight_norm=True, n_hidden=4,
                                equation="u*u__x + u__t - 0.00318309886183791*u__x__x")
    _compare_cuda(problem)


def test_full_size_properties_1m():
    """At the bench size (2^20 points) the oracle is too slow; check properties:
    bit-reproducibility, and additivity over a split of the points (the loss is a
    SUM over points, idrlnet/variable.py:60-65)."""
    rng = np.random.default_rng(12)
    n = 1 << 20
    problem = pu.random_problem(rng, n_points=n, act="silu", weight_norm=True, n_hidden=4,
                                equation="u*u__x + u__t - 0.00318309886183791*u__x__x")
    g1, l1, _, _ = pu.run_cuda_fused(problem)
    g2, l2, _, _ = pu.run_cuda_fused(problem)
    assert np.array_equal(g1["net"], g2["net"]) and l1 == l2  # deterministic reduction
    dom = problem["domains"][0]
    halves = []
    for sl in (slice(0, n // 2 + 77), slice(n // 2 + 77, n)):
        p2 = dict(problem)
        d2 = dict(dom)
        d2["points"] = {k: v[sl] for k, v in dom["points"].items()}
        p2["domains"] = [d2]
        halves.append(pu.run_cuda_fused(p2))
    gs = halves[0][0]["net"] + halves[1][0]["net"]
    ls = halves[0][1]["dom"] + halves[1][1]["dom"]
    scale = np.max(np.abs(g1["net"]))
    assert np.max(np.abs(gs - g1["net"])) <= 1e-5 * scale
    assert abs(ls - l1["dom"]) <= 1e-5 * abs(l1["dom"])


def test_jet_function_forward_backward():
    """Generic path: jets from the kernel, residual + loss in torch."""
    from idrlnet_b200 import compiler as cp
    from idrlnet_b200 import fused
    rng = np.random.default_rng(5)
    problem = pu.random_problem(rng, n_in=3, n_out=2, inputs=("x", "y", "t"), outputs=("u", "v"), n_points=300,
                                act="tanh", equation="sin(u)*v__x + u__t + u__y__y*v__x__x")
    dom = problem["domains"][0]
    want = ["u", "v", "u__x", "u__y", "u__t", "v__x", "u__y__y", "v__x__x", "u__x__x", "v__y__y"]
    ref = po.training_step(problem, dtype=torch.float64, want={"dom": want})
    net = problem["nets"][0]
    mod = pu.torch_module_from_net(net, "cuda:0")
    pack = fused.NetPack(mod, act="tanh")
    plan = cp.plan_jets(("x", "y", "t"), [tuple(w.split("__")[1:]) for w in want])
    cols = {k: torch.as_tensor(v, device="cuda:0").reshape(-1) for k, v in pu.domain_columns(dom).items()}
    wb = [t for i in range(len(pack.layers)) for t in pack.raw_params(i)]
    jets = fused.jet_mlp(plan, "tanh", [cols[k] for k in ("x", "y", "t")], wb)
    names = plan.channel_names(("x", "y", "t"))
    sym = {}
    for c, dn in enumerate(names):
        for o, on in enumerate(("u", "v")):
            sym["__".join((on,) + dn)] = jets[c * 2 + o]
    for w in want:
        r = ref["domains"]["dom"][w].reshape(-1)
        err = np.max(np.abs(sym[w].detach().cpu().numpy() - r)) / max(np.max(np.abs(r)), 1e-30)
        assert err <= TOL, (w, err)
    res = torch.sin(sym["u"]) * sym["v__x"] + sym["u__t"] + sym["u__y__y"] * sym["v__x__x"]
    loss = torch.sum(res ** 2 * cols["area"])
    loss.backward()
    assert abs(float(loss) - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    parts = pu.flat_grad_from_oracle(ref, "net")
    got = [p.grad for p in wb]
    scale = max(np.max(np.abs(p)) for p in parts if p is not None)
    for g, p in zip(got, parts):
        err = np.max(np.abs(g.detach().cpu().numpy().ravel() - p)) / max(np.max(np.abs(p)), 1e-3 * scale)
        assert err <= 20 * TOL, err


def test_device_sampler_matches_host_philox():
    import ctypes as C
    from idrlnet_b200 import native as nv
    n = 100003
    lo = (C.c_float * 3)(-1.0, 0.25, 0.0)
    hi = (C.c_float * 3)(1.0, 0.75, 20.0)
    dev = [torch.empty(n, device="cuda:0") for _ in range(3)]
    sdf = torch.empty(n, device="cuda:0")
    nv.check(nv.lib().pinn_sample_box(nv.ptr_array([t.data_ptr() for t in dev]), 3, lo, hi, n, 1234, 7, 99,
                                      sdf.data_ptr(), 2, torch.cuda.current_stream().cuda_stream))
    host = [np.empty(n, np.float32) for _ in range(3)]
    hsdf = np.empty(n, np.float32)
    cols = (C.c_void_p * 3)(*[h.ctypes.data for h in host])
    pu.emu_lib().pinn_emu_sample_box(cols, 3, lo, hi, n, 1234, 7, 99, hsdf.ctypes.data, 2)
    for d, h in zip(dev, host):
        assert np.array_equal(d.cpu().numpy(), h)
    assert np.allclose(sdf.cpu().numpy(), hsdf, rtol=0, atol=1e-7)
    x = host[0]
    assert x.min() >= -1.0 and x.max() < 1.0 and abs(x.mean()) < 0.01 and abs(x.var() - 1 / 3) < 0.01
    assert hsdf.min() > 0  # every sampled point is inside the rectangle


def test_loss_only_mode_matches_the_full_step():
    """pinn_loss_fused: forward + residual + loss without the reverse pass (idrlnet/solver.py:341-344) returns
    bit-identical losses to the full step; StepEngine finalizes it into its own buffer, so the gradient of the last
    full step (which `.grad` of the module parameters may alias) survives the re-evaluation untouched."""
    from idrlnet_b200 import fused
    problem, _ = gu.load("burgers")
    packs = pu.cuda_packs(problem)
    doms = []
    for d in problem["domains"]:
        fd, pack = pu.bind_fused_domain(problem, d, packs)
        doms.append(fd)
    eng = fused.StepEngine(pack, doms)
    eng.launch()
    full, grad = eng.losses.clone(), eng.grad.clone()
    assert float(eng.grad.abs().max()) > 0
    eng.launch(loss_only=True)
    torch.cuda.synchronize()
    assert torch.equal(eng.losses, full)
    assert torch.equal(eng.grad, grad)
    # wide nets: FP32 FMA streamed, 3xTF32 streamed and the on-chip TF32 kernel keep the same contract
    problem, _ = gu.load("ldc")
    for precision in ("fp32", "fp32x3", "tf32"):
        packs = pu.cuda_packs(problem, precision=precision)
        doms = []
        for d in problem["domains"]:
            fd, pack = pu.bind_fused_domain(problem, d, packs)
            doms.append(fd)
        eng = fused.StepEngine(pack, doms)
        eng.launch()
        full = eng.losses.clone()
        assert float(eng.grad.abs().max()) > 0
        eng.launch(loss_only=True)
        torch.cuda.synchronize()
        assert torch.equal(eng.losses, full), precision


@pytest.mark.gpu
def test_batched_weight_norm_equals_the_per_layer_calls_and_torch():
    """pinn_weight_norm_forward_batch / _backward_batch (all layers of a net in one launch) give bit-identical results to
    the per-layer entry points and agree with torch.nn.utils.weight_norm's forward and autograd backward."""
    import ctypes as C
    from idrlnet_b200 import native as nv
    lib = nv.lib()
    torch.manual_seed(3)
    shapes = [(20, 2), (20, 20), (37, 20), (1, 37), (64, 129)]
    g = [torch.rand(o, 1, device="cuda") + 0.5 for o, _ in shapes]
    v = [torch.randn(o, i, device="cuda") for o, i in shapes]
    dW = [torch.randn(o, i, device="cuda") for o, i in shapes]
    W1, W2 = [torch.empty_like(x) for x in v], [torch.empty_like(x) for x in v]
    dg1, dg2 = [torch.empty_like(x) for x in g], [torch.empty_like(x) for x in g]
    dv1, dv2 = [torch.empty_like(x) for x in v], [torch.empty_like(x) for x in v]
    st = torch.cuda.current_stream().cuda_stream
    arr = (nv.PinnWnLayer * len(shapes))()
    for k, (o, i) in enumerate(shapes):
        nv.check(lib.pinn_weight_norm_forward(g[k].data_ptr(), v[k].data_ptr(), W1[k].data_ptr(), o, i, st))
        nv.check(lib.pinn_weight_norm_backward(g[k].data_ptr(), v[k].data_ptr(), dW[k].data_ptr(), dg1[k].data_ptr(),
                                               dv1[k].data_ptr(), o, i, st))
        a = arr[k]
        a.g, a.v, a.dW, a.W = g[k].data_ptr(), v[k].data_ptr(), dW[k].data_ptr(), W2[k].data_ptr()
        a.dg, a.dv, a.out_dim, a.in_dim = dg2[k].data_ptr(), dv2[k].data_ptr(), o, i
    nv.check(lib.pinn_weight_norm_forward_batch(arr, len(shapes), st))
    nv.check(lib.pinn_weight_norm_backward_batch(arr, len(shapes), st))
    torch.cuda.synchronize()
    for k in range(len(shapes)):
        assert torch.equal(W1[k], W2[k]) and torch.equal(dg1[k], dg2[k]) and torch.equal(dv1[k], dv2[k])
        gk, vk = g[k].clone().requires_grad_(True), v[k].clone().requires_grad_(True)
        Wt = torch._weight_norm(vk, gk, 0)
        Wt.backward(dW[k])
        torch.testing.assert_close(W2[k], Wt.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(dg2[k], gk.grad, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(dv2[k], vk.grad, rtol=1e-4, atol=1e-5)
    assert lib.pinn_weight_norm_forward_batch(arr, 0, st) < 0          # PINN_E_INVALID: no layers
    assert lib.pinn_weight_norm_backward_batch(arr, 17, st) < 0        # more than PINN_MAX_LAYERS
