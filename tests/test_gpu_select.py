"""forward-only residuals + device top-k (adaptive re-sampling) against the oracle."""
import numpy as np
import pytest
import torch

from tests import golden_util as gu
from tests import problem_util as pu
from oracle import pinn_oracle as po

pytestmark = pytest.mark.gpu


def _check_topk(v, k):
    from idrlnet_b200 import fused
    idx, val = fused.topk_abs(torch.from_numpy(v).cuda(), k)
    ref = po.topk_abs(v, k)
    np.testing.assert_array_equal(idx.cpu().numpy(), ref)          # index work: bit-exact
    np.testing.assert_array_equal(val.cpu().numpy(), v.ravel()[ref])


@pytest.mark.parametrize("n", [1, 5, 257, 1000, 100003, (1 << 20) + 7])
@pytest.mark.parametrize("k", [1, 7, 200, 4096])
def test_topk_random(n, k):
    if k > n:
        pytest.skip("k > n is rejected (separate test)")
    rng = np.random.default_rng(n * 31 + k)
    _check_topk(rng.standard_normal(n).astype(np.float32), k)


@pytest.mark.parametrize("case", ["ties", "zeros", "signs", "inf", "denormal"])
def test_topk_edge_values(case):
    rng = np.random.default_rng(5)
    n = 50000
    if case == "ties":          # heavy ties across the selection threshold -> ascending index decides
        v = rng.integers(-3, 4, n).astype(np.float32)
    elif case == "zeros":
        v = np.zeros(n, np.float32)
    elif case == "signs":       # |-x| == |x|
        v = np.repeat(rng.standard_normal(n // 2).astype(np.float32), 2) * np.tile([1.0, -1.0], n // 2).astype(np.float32)
    elif case == "inf":
        v = rng.standard_normal(n).astype(np.float32)
        v[[3, 77, 40000]] = [np.inf, -np.inf, np.inf]
    else:
        v = (rng.standard_normal(n) * 1e-41).astype(np.float32)
    for k in (1, 200, 1025):
        _check_topk(v, k)


def test_topk_rejects_bad_k():
    from idrlnet_b200 import fused, native
    v = torch.randn(100, device="cuda")
    with pytest.raises(native.PinnError):
        fused.topk_abs(v, 101)
    with pytest.raises(native.PinnError):
        fused.topk_abs(v, 0)
    with pytest.raises(native.PinnError):
        fused.topk_abs(torch.randn(10000, device="cuda"), 4097)


@pytest.mark.parametrize("name", ["burgers", "allen_cahn", "ldc", "simple_poisson"])
def test_residual_forward_matches_oracle(name):
    """pinn_residual_forward on every fusable domain of the golden problems vs the oracle's
    (reference algorithm) residual columns: 1e-5 norm-wise, FP32."""
    from idrlnet_b200 import compiler as cp, fused
    problem, _ = gu.load(name)
    nets = po.build_nets(problem)
    packs = pu.cuda_packs(problem)
    checked = 0
    for dom in problem["domains"]:
        try:
            fd, pack = pu.bind_fused_domain(problem, dom, packs)
        except cp.NotFusable:
            continue
        got = fused.domain_residuals(pack, fd).cpu().numpy()
        var = po.forward_domain(problem, dom, nets)
        for e, eq in enumerate(fd.compiled.equations):
            ref = var[eq.name].detach().numpy().ravel()
            scale = max(np.abs(ref).max(), 1e-6)
            assert np.abs(got[e] - ref).max() <= 1e-5 * scale, (name, dom["name"], eq.name)
            checked += 1
    assert checked > 0


def test_solver_residual_top_k_matches_infer_step(tmp_path):
    problem, _ = gu.load("burgers")
    s, _ = pu.solver_from_problem(problem, tmp_path)
    dom = next(d["name"] for d in problem["domains"] if "burgers_u" in d["targets"])
    top = s.residual_top_k(dom, "burgers_u", 200)
    full = s.infer_step({dom: ["x", "t", "burgers_u"]})[dom]
    r = full["burgers_u"].detach().cpu().numpy().ravel()
    # the generic graph (autograd-free jets + torch residual) and the residual kernel agree to FP32 noise, so
    # compare the selected SET through the values: the k-th largest |r| of both must coincide
    ref_idx = po.topk_abs(r, 200)
    got_idx = top["index"].cpu().numpy().ravel()
    assert len(set(got_idx) ^ set(ref_idx)) <= 4          # only near-ties at the threshold may swap
    np.testing.assert_allclose(np.abs(top["burgers_u"].cpu().numpy().ravel()), np.abs(r[got_idx]), rtol=1e-4, atol=1e-6)
    np.testing.assert_array_equal(top["x"].cpu().numpy().ravel(), full["x"].detach().cpu().numpy().ravel()[got_idx])
    assert np.all(np.diff(np.abs(top["burgers_u"].cpu().numpy().ravel())) <= 0)


def _host_reference(geo, density, seed, offset=0):
    """Candidates of pinn_sample_box for the same (seed, offset), signed distance by the NumPy lowering of the
    geometry's SymPy sdf (the reference's own evaluation path, float64)."""
    import ctypes as C
    from idrlnet_b200 import native as nv
    from idrlnet_b200.geo_utils import lambdify_np
    names = list(geo.bounds)
    n = int(np.prod([hi - lo for lo, hi in geo.bounds.values()]) * density)
    cols = [torch.empty(n, device="cuda") for _ in names]
    lo = (C.c_float * len(names))(*[float(geo.bounds[k][0]) for k in names])
    hi = (C.c_float * len(names))(*[float(geo.bounds[k][1]) for k in names])
    nv.check(nv.lib().pinn_sample_box(nv.ptr_array([c.data_ptr() for c in cols]), len(names), lo, hi, n, seed, 0, offset,
                                      None, 0, torch.cuda.current_stream().cuda_stream))
    pts = {k: c.cpu().numpy().astype(np.float64).reshape(-1, 1) for k, c in zip(names, cols)}
    return pts, lambdify_np(geo.sdf, names)(**pts)


def _solids():
    import idrlnet_b200.shortcut as sc
    r, c = sc.Rectangle((-1.0, -1.0), (1.0, 1.0)), sc.Circle((0.2, 0.0), 0.5)
    c2 = sc.Circle((-0.8, 0.9), 0.6)
    return {
        "rect_minus_circle": r - c,
        "union": c + c2,
        "intersection": r & c2,
        "nested": (r - c) - (c2 & r),
        "channel_minus_holes": sc.Tube2D((-1, -0.5), (1, 0.5)) - sc.Circle((-0.6, 0.0), 0.2) - sc.Circle((0.3, 0.1), 0.25),
        "interval": sc.Line1D(-1.0, 2.5),
        "box_and_sphere": sc.Box((-1, -1, 0), (1, 1, 1)) & sc.Sphere((0.0, 0.0, 0.2), 0.9),
    }


@pytest.mark.parametrize("name", ["rect_minus_circle", "union", "intersection", "nested", "channel_minus_holes", "interval",
                                  "box_and_sphere"])
def test_csg_kernel_matches_host_sdf(name):
    """pinn_sample_csg: same Philox candidates as pinn_sample_box, kept iff sdf > 0, written in
    candidate order; the signed distance agrees with the SymPy expression evaluated on the host in float64."""
    geo = _solids()[name]
    assert geo._tree is not None
    density = 3000
    got = geo.sample_interior_device(density, seed=5, offset=17)
    again = geo.sample_interior_device(density, seed=5, offset=17)
    pts, sdf = _host_reference(geo, density, 5, 17)
    keep = sdf[:, 0] > 0
    axes = list(geo.bounds)
    near = np.abs(sdf[:, 0]) < 1e-6                      # float32 vs float64 may disagree on the sign only here
    n_got = got[axes[0]].shape[0]
    assert abs(n_got - keep.sum()) <= near.sum()
    if near.sum() == 0:
        for k in axes:
            np.testing.assert_array_equal(got[k].cpu().numpy().ravel(), pts[k][keep, 0].astype(np.float32))   # order too
        np.testing.assert_allclose(got["sdf"].cpu().numpy().ravel(), sdf[keep, 0], atol=2e-6)
    assert (got["sdf"] > 0).all()
    for k in got:
        assert torch.equal(got[k], again[k])                                                                   # deterministic
    assert torch.allclose(got["area"], torch.full_like(got["area"], 1.0 / density))


def test_csg_kernel_rejects_bad_programs():
    import ctypes as C
    from idrlnet_b200 import native as nv
    lib = nv.lib()
    cols = [torch.empty(100, device="cuda") for _ in range(2)]
    lo, hi = (C.c_float * 2)(-1, -1), (C.c_float * 2)(1, 1)
    count = torch.zeros(1, dtype=torch.int64, device="cuda")
    ws = torch.empty(64, dtype=torch.uint8, device="cuda")

    def run(ops):
        prog = (nv.PinnShapeOp * len(ops))()
        for o, (op, p) in zip(prog, ops):
            o.op = op
            for i, v in enumerate(p):
                o.p[i] = v
        return lib.pinn_sample_csg(nv.ptr_array([c.data_ptr() for c in cols]), 2, lo, hi, 100, 100, 1, 0, 0, prog, len(ops),
                                   None, count.data_ptr(), ws.data_ptr(), 64, torch.cuda.current_stream().cuda_stream)
    assert run([(1, (-1, -1, 1, 1))]) == 0
    assert run([(17, ())]) < 0                                   # operator without operands
    assert run([(1, (-1, -1, 1, 1)), (2, (0, 0, 0.5))]) < 0     # two values left on the stack
    assert run([(4, (0, 0, 0, 1, 1, 1))]) < 0                    # 3-D primitive with 2 coordinates
    assert run([(99, ())]) < 0


def test_device_geometry_sampler_matches_host_sdf():
    """CSG solids sample on the GPU (Philox box points + SymPy sdf lowered to torch);
    the kept points are exactly those with sdf > 0 and their sdf agrees with the NumPy lowering."""
    import sympy as sp
    import idrlnet_b200.shortcut as sc
    from idrlnet_b200.geo_utils import lambdify_np
    x = sp.Symbol("x")
    geo = (sc.Rectangle((-1.0, -1.0), (1.0, 1.0)) - sc.Circle((0.2, 0.0), 0.5)).translation((0.0, 0.0))  # no CSG tree: torch path
    assert geo._tree is None
    a = geo.sample_interior_device(20000, seed=11)
    b = geo.sample_interior_device(20000, seed=11)
    c = geo.sample_interior_device(20000, seed=12, sieve=(x > 0))
    n = a["x"].shape[0]
    frac = n / 80000
    assert abs(frac - (4 - np.pi * 0.25) / 4) < 0.01
    assert torch.equal(a["x"], b["x"]) and torch.equal(a["sdf"], b["sdf"])          # counter-based: reproducible
    assert (a["sdf"] > 0).all() and (c["x"] > 0).all() and c["x"].shape[0] < n
    host = lambdify_np(geo.sdf, ["x", "y"])(x=a["x"].cpu().numpy().astype(np.float64), y=a["y"].cpu().numpy().astype(np.float64))
    np.testing.assert_allclose(a["sdf"].cpu().numpy(), host, atol=2e-6)
    assert torch.allclose(a["area"], torch.full_like(a["area"], 1 / 20000))


@pytest.mark.gpu
def test_device_boundary_sampler():
    """`sample_boundary` on the device -- every kept point lies on the solid's boundary, normals are
    unit vectors pointing outwards, per-edge counts and areas follow density x measure, the sieve applies."""
    import sympy as sp
    import idrlnet_b200.shortcut as sc
    from idrlnet_b200.geo_utils import lambdify_np
    x, y = sp.symbols("x y")
    rect, circ = sc.Rectangle((-1.0, -1.0), (1.0, 1.0)), sc.Circle((0.2, 0.0), 0.5)
    geo = rect - circ
    b = geo.sample_boundary_device(1000, seed=4)
    n = b["x"].shape[0]
    expect = 4 * 2000 + int(1000 * 2 * np.pi * 0.5)
    assert abs(n - expect) <= 2, (n, expect)                      # nothing of this solid's boundary is sieved away
    assert set(b) == {"x", "y", "normal_x", "normal_y", "area", "sdf"}
    host_sdf = lambdify_np(geo.sdf, ["x", "y"])(x=b["x"].cpu().numpy().astype(np.float64), y=b["y"].cpu().numpy().astype(np.float64))
    assert np.abs(host_sdf).max() < 1e-5
    nrm = (b["normal_x"] ** 2 + b["normal_y"] ** 2).cpu().numpy()
    np.testing.assert_allclose(nrm, 1.0, atol=1e-5)
    out = {"x": (b["x"] + 1e-3 * b["normal_x"]).cpu().numpy().astype(np.float64), "y": (b["y"] + 1e-3 * b["normal_y"]).cpu().numpy().astype(np.float64)}
    assert (lambdify_np(geo.sdf, ["x", "y"])(**out) < 0).all()    # normals leave the solid (the hole's point inwards to it)
    np.testing.assert_allclose(float(b["area"].sum()), 8.0 + 2 * np.pi * 0.5, rtol=1e-3)
    again = geo.sample_boundary_device(1000, seed=4)
    assert all(torch.equal(b[k], again[k]) for k in b)
    left = rect.sample_boundary_device(500, seed=1, sieve=sp.Eq(x, -1.0))
    assert left["x"].shape[0] == 1000 and float(left["x"].max()) == -1.0 and float(left["normal_x"].max()) == -1.0
    inner = (rect + circ).sample_boundary_device(200, seed=2)      # the circle lies inside: its edge is sieved away
    assert inner["x"].shape[0] == 4 * 400
