"""Host-side mirror of the reference interface on the CPU (generic torch path):
graph wiring, sampling semantics, Solver loop, checkpoint layout."""
import os

import numpy as np
import pytest
import sympy as sp
import torch

os.environ.setdefault("IDRLNET_B200_LOGLEVEL", "WARNING")
import idrlnet_b200.shortcut as sc  # noqa: E402
from tests import golden_util as gu  # noqa: E402


def _poisson(tmp, max_iter=5):
    x, y = sp.symbols("x y")
    rec = sc.Rectangle((-1.0, -1.0), (1.0, 1.0))

    @sc.datanode
    class LeftRight(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_boundary(50, sieve=((y > -1.0) & (y < 1.0))), {"T": 0.0}

    @sc.datanode(name="up_down")
    class UpDown(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_boundary(50, sieve=((x > -1.0) & (x < 1.0))), {"normal_gradient_T": 0.0}

    @sc.datanode(name="heat_domain")
    class Heat(sc.SampleDomain):
        def __init__(self):
            self.points = 50

        def sampling(self, *args, **kwargs):
            return rec.sample_interior(self.points), {"diffusion_T": 1.0}

    net = sc.get_net_node(inputs=("x", "y"), outputs=("T",), name="net1", arch=sc.Arch.mlp)
    pde = sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False)
    grad = sc.NormalGradient("T", dim=2, time=False)
    return sc.Solver(sample_domains=(Heat(), LeftRight(), UpDown()), netnodes=[net], pdes=[pde, grad],
                     max_iter=max_iter, network_dir=str(tmp), loading=False, save_freq=2)


def test_graph_order_matches_reference(tmp_path):
    """['net1', 'DiffusionNode', '<diffusion_T>'] for the Poisson interior."""
    np.random.seed(0)
    torch.manual_seed(0)
    s = _poisson(tmp_path)
    order = [n.name for n in s.vertex_pipelines["heat_domain"].evaluation_order_list]
    assert order == ["net1", "DiffusionNode", "<diffusion_T>"]
    assert [n.name for n in s.vertex_pipelines["LeftRight"].evaluation_order_list] == ["net1", "<T>"]
    assert [n.name for n in s.vertex_pipelines["up_down"].evaluation_order_list] == ["net1", "NormalGradient", "<normal_gradient_T>"]


def test_solver_trains_and_checkpoints(tmp_path):
    np.random.seed(0)
    torch.manual_seed(0)
    s = _poisson(tmp_path, max_iter=6)
    fired = []

    class Rec(sc.Receiver):
        def receive_notify(self, solver, message):
            fired.extend(message.keys())

    s.register_receiver(Rec())
    l0 = float(s.train_pipe())
    s.solve()
    assert s.global_step == 6 and float(s.loss_component["heat_domain_loss"]) > 0
    assert set(s.loss_component) == {"heat_domain_loss", "LeftRight_loss", "up_down_loss"}
    for sig in (sc.Signal.TRAIN_PIPE_START, sc.Signal.BEFORE_COMPUTE_LOSS, sc.Signal.AFTER_COMPUTE_LOSS,
                sc.Signal.BEFORE_BACKWARD, sc.Signal.TRAIN_PIPE_END, sc.Signal.SOLVE_END):
        assert sig in fired
    ck = torch.load(os.path.join(str(tmp_path), "model.ckpt"), map_location="cpu")
    assert {"net1_dict", "optimizer_0_dict", "global_step"} <= set(ck)
    assert "layers.mlp_0.weight_g" in ck["net1_dict"]  # the reference's state-dict keys
    s.set_domain_parameter("heat_domain", {"points": 100})
    out = s.infer_step({"heat_domain": ["x", "y", "T", "T__x"]})
    assert out["heat_domain"]["T__x"].shape == (400, 1)
    assert l0 > 0


def test_sampling_semantics():
    np.random.seed(1)
    rec = sc.Rectangle((-1.0, -2.0), (1.0, 2.0))
    pts = rec.sample_interior(10)  # measure 8 * density 10
    assert set(pts) == {"x", "y", "sdf", "area"} and len(pts["x"]) == 80
    assert np.all(pts["sdf"] > 0) and np.allclose(pts["area"], 0.1)
    assert np.allclose(pts["sdf"], np.minimum(1 - np.abs(pts["x"]), 2 - np.abs(pts["y"])))
    y = sp.Symbol("y")
    b = rec.sample_boundary(10, sieve=sp.Eq(y, 2.0))
    assert len(b["x"]) == 20 and np.all(b["normal_y"] == 1.0) and np.allclose(b["area"], 0.1)
    t = sp.Symbol("t")
    line = sc.Line1D(-1.0, 1.0)
    lb = line.sample_boundary(7, param_ranges={t: (0, 1)})
    assert len(lb["x"]) == 14 and set(np.unique(lb["normal_x"])) == {-1.0, 1.0} and "t" in lb
    li = line.sample_interior(30, param_ranges={t: 0.0})
    assert len(li["x"]) == 60 and np.all(li["t"] == 0.0)


def test_datanode_evaluates_symbolic_targets():
    x = sp.Symbol("x")
    line = sc.Line1D(-1.0, 1.0)

    @sc.datanode(name="ic")
    def init():
        return line.sample_interior(10, param_ranges={sp.Symbol("t"): 0.0}), {"u": -sp.sin(sp.pi * x), "lambda_u": 100}

    dn = init()
    assert dn.outputs == ["u"] and dn.lambda_outputs == ["lambda_u"] and dn.name == "ic"
    s = dn.sample()
    assert torch.allclose(s["u"], -torch.sin(np.pi * s["x"]).detach(), atol=1e-6)
    assert float(s["lambda_u"][0]) == 100.0 and not s["lambda_u"].requires_grad and s["x"].requires_grad


def test_generic_path_reproduces_reference_golden(tmp_path):
    """The host mirror's generic path (graph + autograd) against the reference's golden
    numbers for Allen-Cahn, which exercises shared nets, ExpressionNodes and Difference."""
    problem, flat = gu.load("allen_cahn")
    x, t = sp.Symbol("x"), sp.Symbol("t")
    u = sp.Function("u")(x, t)
    up = sp.Function("up")(x, t)
    net = sc.MLP([2, 128, 128, 128, 128, 2], activation=sc.Activation.tanh)
    lins = [l for k, l in net.layers.items() if not k.endswith("_activation")]
    with torch.no_grad():
        for lin, lay in zip(lins, problem["nets"][0]["layers"]):
            lin.weight_g.copy_(torch.as_tensor(lay["g"]))
            lin.weight_v.copy_(torch.as_tensor(lay["v"]))
            lin.bias.copy_(torch.as_tensor(lay["b"]))
    net_u = sc.NetNode(inputs=("x", "t"), outputs=("u",), name="net1", net=net)
    xp = sc.ExpressionNode(name="xp", expression=x + 2)
    tilde = sc.get_shared_net_node(net_u, inputs=("xp", "t"), outputs=("up",), name="net2", arch="mlp")
    nodes = [sc.AllenCahnNode(u="u", gamma_1=0.0001, gamma_2=5), xp,
             sc.ExpressionNode(expression=up.diff(x), name="diff_up"), sc.ExpressionNode(expression=u.diff(x), name="diff_u"),
             sc.Difference(T="diff_u", S="diff_up"), sc.Difference(T="u", S="up")]
    doms = []
    for d in problem["domains"]:
        pts = {k: np.asarray(v) for k, v in d["points"].items()}
        cons = {k: (np.asarray(v) if np.ndim(v) else float(v)) for k, v in d["targets"].items()}
        for k, v in d["lambdas"].items():
            cons["lambda_" + k] = np.asarray(v) if np.ndim(v) else float(v)
        doms.append(sc.get_data_node((lambda p=pts, c=cons: (dict(p), dict(c))), name=d["name"]))
    s = sc.Solver(sample_domains=tuple(doms), netnodes=[net_u, tilde], pdes=nodes, network_dir=str(tmp_path),
                  loading=False, max_iter=1)
    s.optimizers[0].zero_grad()
    loss = float(s._loss_and_grads())
    ref = float(flat["ref32/loss"])
    assert abs(loss - ref) <= 2e-5 * abs(ref)
    for k, v in gu.ref_components(flat, "ref32").items():
        assert abs(float(s.loss_component[k]) - v) <= 1e-4 * max(abs(v), 1e-9), k
    for pname, g in gu.ref_grads(flat, "ref32").items():
        p = dict(net.named_parameters())[pname.split("/", 1)[1]]
        assert np.max(np.abs(p.grad.numpy().reshape(g.shape) - g)) <= 2e-4 * np.max(np.abs(g)), pname


def test_datanode_targets_for_tensor_points_stay_off_the_host_path():
    """Points that arrive as tensors get their targets evaluated with torch (no numpy round trip): constants stay 0-d
    scalars, SymPy targets become [N, 1] tensors with the values the numpy path gives."""
    xs = torch.linspace(-1.0, 1.0, 11).reshape(-1, 1)
    x = sp.Symbol("x")

    @sc.datanode(name="tensor_points")
    class Dom(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return {"x": xs, "area": torch.full_like(xs, 0.1)}, {"u": 0.25, "v": -sp.sin(sp.pi * x), "lambda_u": 100.0}

    out = Dom().sample()
    assert out["u"].dim() == 0 and float(out["u"]) == 0.25 and out["u"].device.type == "cpu"
    assert float(out["lambda_u"]) == 100.0
    np.testing.assert_allclose(out["v"].detach().numpy().ravel(), -np.sin(np.pi * xs.numpy().ravel()), atol=1e-6)

    @sc.datanode(name="numpy_points")
    class DomNp(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return {"x": xs.numpy(), "area": np.full_like(xs.numpy(), 0.1)}, {"u": 0.25, "v": -sp.sin(sp.pi * x)}

    ref = DomNp().sample()
    assert tuple(ref["u"].shape) == (11, 1)                      # the reference's shape for host-sampled points
    np.testing.assert_allclose(ref["v"].detach().numpy(), out["v"].detach().numpy(), atol=1e-6)


def test_csg_tree_is_tracked_for_the_device_sampler():
    from idrlnet_b200.geo_utils.geo import _program
    r, c = sc.Rectangle((-1.0, -1.0), (1.0, 1.0)), sc.Circle((0.2, 0.0), 0.5)
    assert _program((r - c)._tree) == [("rect", (-1.0, -1.0, 1.0, 1.0)), ("circle", (0.2, 0.0, 0.5)), ("difference", ())]
    assert _program((~(r & c))._tree)[-2:] == [("intersection", ()), ("complement", ())]
    assert sc.Circle((0.0, 0.0), sp.Symbol("r"))._tree is None          # symbolic parameter: torch-lowered fallback
    assert (r - sc.Circle((0.0, 0.0), sp.Symbol("r")))._tree is None
    assert sc.Rectangle((0.0, 0.0), (1.0, 1.0)).scaling(2.0)._tree is None
    assert sc.Polygon([(0, 0), (1, 0), (0, 1)])._tree is None


def test_boundary_device_sampler_logic_with_a_stub_generator(monkeypatch):
    """`sample_boundary_device` end to end on CPU tensors, with pinn_sample_box replaced by a NumPy stub (the kernel
    itself is checked bit for bit in tests/test_gpu_narrow.py): per-edge counts, expressions, sieve, areas."""
    import ctypes as C
    from idrlnet_b200 import native as nv

    class Stub:
        calls = 0

        def pinn_sample_box(self, cols, n_dims, lo, hi, n, seed, stream_id, offset, sdf, sdf_dims, stream):
            rng = np.random.default_rng(int(seed) * 1000 + int(stream_id))
            for d in range(n_dims):
                buf = (C.c_float * n).from_address(cols[d])
                buf[:] = rng.uniform(lo[d], hi[d], n).astype(np.float32).tolist()
            Stub.calls += 1
            return 0

    monkeypatch.setattr(nv, "lib", lambda: Stub())
    x, y = sp.symbols("x y")
    rect, circ = sc.Rectangle((-1.0, -1.0), (1.0, 1.0)), sc.Circle((0.2, 0.0), 0.5)
    geo = rect - circ
    b = geo.sample_boundary_device(200, seed=4, device="cpu")
    assert Stub.calls == 5
    n = b["x"].shape[0]
    assert abs(n - (4 * 400 + int(200 * 2 * np.pi * 0.5))) <= 2
    assert set(b) == {"x", "y", "normal_x", "normal_y", "area", "sdf"}
    assert float(b["sdf"].abs().max()) < 1e-4
    np.testing.assert_allclose((b["normal_x"] ** 2 + b["normal_y"] ** 2).numpy(), 1.0, atol=1e-5)
    np.testing.assert_allclose(float(b["area"].sum()), 8.0 + 2 * np.pi * 0.5, rtol=1e-3)
    left = rect.sample_boundary_device(100, seed=1, sieve=sp.Eq(x, -1.0), device="cpu")
    assert left["x"].shape[0] == 200 and float(left["x"].max()) == -1.0 and float(left["normal_x"].max()) == -1.0
    inner = (rect + circ).sample_boundary_device(50, seed=2, device="cpu")   # the circle's edge lies inside: sieved away
    assert inner["x"].shape[0] == 4 * 100
    # parameterised solid: the radius is sampled per point and returned as a column
    r = sp.Symbol("r")
    holes = sc.Tube2D((-1, -0.5), (1, 0.5)) - sc.Circle((0.0, 0.0), r)
    pb = holes.sample_boundary_device(100, seed=3, param_ranges={r: (0.1, 0.3)}, device="cpu")
    assert "r" in pb and float(pb["r"].min()) >= 0.1 and float(pb["r"].max()) <= 0.3
    assert float(pb["sdf"].abs().max()) < 1e-4
