"""SURVEY.md 8 a7 on the product path: `Solver(device_sampling=True)` makes an UNCHANGED example (its SampleDomains call
`geo.sample_interior` / `geo.sample_boundary` exactly as with the reference) draw every point on the GPU; statistical
checks of the device samplers against the reference's host sampler semantics."""
import math

import numpy as np
import pytest
import sympy as sp
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _reset_switch():
    from idrlnet_b200.geo_utils import geo
    yield
    geo.set_device_sampling(False)


def _burgers_example(sc, density=20000):
    """examples/burgers_equation/burgers_equation.py:7-45, unchanged apart from the density."""
    x, t = sp.Symbol("x"), sp.Symbol("t")
    time_range = {t: (0, 1)}
    geo = sc.Line1D(-1.0, 1.0)

    @sc.datanode(name="burgers_equation")
    def interior_domain():
        return geo.sample_interior(density, bounds={x: (-1.0, 1.0)}, param_ranges=time_range), {"burgers_u": 0}

    @sc.datanode(name="t_boundary")
    def init_domain():
        return geo.sample_interior(100, param_ranges={t: 0.0}), sc.Variables({"u": -sp.sin(math.pi * x)})

    @sc.datanode(name="x_boundary")
    def boundary_domain():
        return geo.sample_boundary(100, param_ranges=time_range), sc.Variables({"u": 0})

    net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    pde = sc.BurgersNode(u="u", v=0.01 / math.pi)
    return (interior_domain(), init_domain(), boundary_domain()), [net], [pde]


def test_unchanged_example_samples_on_the_gpu(tmp_path):
    import idrlnet_b200.shortcut as sc
    from idrlnet_b200.geo_utils import geo
    torch.manual_seed(0)
    np.random.seed(0)
    geo.set_device_sampling(True, seed=7)  # (the datanode factories sample once when they are built)
    domains, nets, pdes = _burgers_example(sc)
    state = np.random.get_state()[1].copy()
    s = sc.Solver(sample_domains=domains, netnodes=nets, pdes=pdes, network_dir=str(tmp_path), loading=False, max_iter=50,
                  device_sampling=True, sampling_seed=7)
    assert sorted(s._fused_domains) == ["burgers_equation", "t_boundary", "x_boundary"]
    samples = s.sample_variables_from_domains()
    pts = samples["burgers_equation"]
    assert all(isinstance(pts[k], torch.Tensor) and pts[k].is_cuda for k in ("x", "t", "sdf", "area"))  # (a constant target
    # stays a 0-d host scalar that the kernel descriptor takes as a constant)
    assert pts["x"].shape == (40000, 1) and pts["t"].shape == (40000, 1)
    again = s.sample_variables_from_domains()["burgers_equation"]
    assert not torch.equal(again["x"], pts["x"])  # fresh points on every call
    bc = samples["x_boundary"]
    assert set(torch.unique(bc["x"]).tolist()) == {-1.0, 1.0} and bc["x"].shape[0] == 200
    assert float(bc["t"].min()) >= 0.0 and float(bc["t"].max()) < 1.0
    ic = samples["t_boundary"]
    assert float(ic["t"].abs().max()) == 0.0
    np.testing.assert_allclose(ic["u"].cpu().numpy(), -np.sin(math.pi * ic["x"].cpu().numpy()), atol=1e-6)
    losses = [float(s.train_pipe()) for _ in range(50)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5])
    assert np.array_equal(np.random.get_state()[1], state)  # the host NumPy stream was never consumed


def test_device_interior_sampler_is_uniform_like_the_host_one():
    """Moments and a 2-D chi-square of the device points against the uniform law the reference samples from
    (idrlnet/geo_utils/geo.py:316-343, np.random.uniform), on rectangle - circle (kernel pinn_sample_csg) and on the plain
    rectangle (no-rejection fast path)."""
    import idrlnet_b200.shortcut as sc
    rect = sc.Rectangle((-1.0, -0.5), (1.0, 0.5))
    for geo_, inside in ((rect, lambda x, y: np.ones_like(x, bool)),
                         (rect - sc.Circle((0.0, 0.0), 0.3), lambda x, y: x * x + y * y > 0.09)):
        dev = geo_.sample_interior_device(100000, seed=11, offset=0)
        np.random.seed(5)
        host = geo_.sample_interior(100000)
        xd, yd = dev["x"].cpu().numpy().ravel(), dev["y"].cpu().numpy().ravel()
        assert inside(xd, yd).all()
        assert abs(len(xd) - len(host["x"])) <= 5 * math.sqrt(len(host["x"]))  # same acceptance rate
        assert abs(float(dev["area"][0]) - float(host["area"][0, 0])) < 1e-12
        # chi-square on a 20 x 10 grid of equal cells, cells cut by the circle excluded
        hx, _, _ = np.histogram2d(xd, yd, bins=(20, 10), range=((-1, 1), (-0.5, 0.5)))
        hh, _, _ = np.histogram2d(host["x"].ravel(), host["y"].ravel(), bins=(20, 10), range=((-1, 1), (-0.5, 0.5)))
        full = hh > 0.95 * hh.max()
        exp = hx[full].sum() / full.sum()
        chi2 = ((hx[full] - exp) ** 2 / exp).sum()
        dof = full.sum() - 1
        assert chi2 < dof + 6 * math.sqrt(2 * dof), (chi2, dof)
        # sdf column agrees with the host evaluation of the symbolic sdf at the same points
        from idrlnet_b200.geo_utils.sympy_np import lambdify_np
        want = lambdify_np(geo_.sdf, ["x", "y"])(x=xd.reshape(-1, 1).astype(np.float64), y=yd.reshape(-1, 1).astype(np.float64))
        np.testing.assert_allclose(dev["sdf"].cpu().numpy(), want, atol=2e-6)


def test_rank_streams_are_disjoint():
    """Rank-offset Philox streams (torch.distributed): emulate two ranks by asking for their stream ids."""
    import idrlnet_b200.shortcut as sc
    from idrlnet_b200.geo_utils import geo
    rect = sc.Rectangle((0.0, 0.0), (1.0, 1.0))
    real = geo._DeviceSampling.rank_world
    try:
        geo._DeviceSampling.rank_world = staticmethod(lambda: (0, 2))
        a = rect.sample_interior_device(5000, seed=3, offset=0)["x"].cpu().numpy().ravel()
        geo._DeviceSampling.rank_world = staticmethod(lambda: (1, 2))
        b = rect.sample_interior_device(5000, seed=3, offset=0)["x"].cpu().numpy().ravel()
    finally:
        geo._DeviceSampling.rank_world = real
    assert len(np.intersect1d(a, b)) < 10 and abs(np.corrcoef(a, b)[0, 1]) < 0.05


def test_boundary_programs_on_the_device():
    """`sample_boundary_device` by compiled edge programs (pinn_sample_program: one launch per edge draws the edge
    parameters and writes coordinates, normals, measure, parameter columns and the composed signed distance): geometric
    invariants, the measure, and agreement with the torch-lowered fallback of the same expressions."""
    import idrlnet_b200.shortcut as sc
    t = sp.Symbol("t")
    circ = sc.Circle((0.2, -0.1), 0.7)
    b = circ.sample_boundary_device(2000, seed=5, offset=0, param_ranges={t: (0.0, 3.0)})
    assert circ.__dict__["_prog_cache"] and all(v[0] is not None for v in circ.__dict__["_prog_cache"].values())
    x, y = b["x"].cpu().numpy().ravel(), b["y"].cpu().numpy().ravel()
    n = len(x)
    assert n == int(2000 * 2 * math.pi * 0.7)
    np.testing.assert_allclose((x - 0.2) ** 2 + (y + 0.1) ** 2, 0.49, atol=2e-6)
    nx, ny = b["normal_x"].cpu().numpy().ravel(), b["normal_y"].cpu().numpy().ravel()
    np.testing.assert_allclose(nx * nx + ny * ny, 1.0, atol=2e-6)
    np.testing.assert_allclose(nx * 0.7, x - 0.2, atol=2e-6)
    assert abs(float(b["area"].sum()) - 2 * math.pi * 0.7) < 1e-3
    assert float(b["sdf"].abs().max()) < 1e-5
    tt = b["t"].cpu().numpy().ravel()
    assert tt.min() >= 0.0 and tt.max() < 3.0 and abs(tt.mean() - 1.5) < 0.1
    ang = np.arctan2(y + 0.1, x - 0.2)
    hist, _ = np.histogram(ang, bins=16, range=(-math.pi, math.pi))
    assert ((hist - n / 16) ** 2 / (n / 16)).sum() < 15 + 6 * math.sqrt(30)  # uniform in the edge parameter
    # CSG with a sieve: the masked path; same keys and counts as the torch-lowered evaluation of the same draws
    rect = sc.Rectangle((-1.0, -1.0), (1.0, 1.0))
    geo_ = rect - sc.Circle((0.0, 0.0), 0.3)
    xs = sp.Symbol("x")
    fast = geo_.sample_boundary_device(500, seed=9, offset=0, sieve=(xs > -0.5))
    cache = geo_.__dict__["_prog_cache"]
    for k in list(cache):
        cache[k] = (None, cache[k][1])  # force the fallback
    slow = geo_.sample_boundary_device(500, seed=9, offset=0, sieve=(xs > -0.5))
    assert set(fast) == set(slow)
    assert abs(fast["x"].shape[0] - slow["x"].shape[0]) <= 0.05 * slow["x"].shape[0]
    assert float(fast["x"].min()) > -0.5 and float(fast["sdf"].abs().max()) < 1e-4
    assert abs(float(fast["area"].sum()) - float(slow["area"].sum())) < 0.05 * float(slow["area"].sum())
