"""the mixed second derivative and the pure third / fourth derivative along one input have their own jet
channels in the narrow kernels (idrlnet_b200/csrc/narrow_warp.h, NarrowWarp's X), so such problems take the fused step
instead of nested torch.autograd.grad (idrlnet/variable.py:161-208).  Checked through the Solver against FP64 autograd
on a CPU copy of the same module; what still has no channel must fall back to the generic graph, on the GPU."""
import numpy as np
import pytest
import sympy as sp
import torch

import idrlnet_b200.shortcut as sc

pytestmark = pytest.mark.gpu

LOSS_TOL = 2e-5   # FP32 kernels against FP64 autograd, relative
GRAD_TOL = 2e-4   # raw (weight_g / weight_v / bias) gradients, relative to max(|layer|, 1e-3 |all|)


def _fixed_domain(name, pts, cons, **kw):
    return sc.get_data_node((lambda p=pts, c=cons: (dict(p), dict(c))), name=name, **kw)


def _points(rng, n, keys, lo=-1.0, hi=1.0):
    p = {k: rng.uniform(lo, hi, (n, 1)).astype(np.float32) for k in keys}
    p["area"] = np.full((n, 1), 1.0 / n, np.float32)
    return p


def _grad(a, b):
    return torch.autograd.grad(a, b, torch.ones_like(a), create_graph=True)[0]


class _Fp64Mlp:
    """FP64 CPU restatement of a NetNode's MLP on copies of its raw parameters (weight-norm layers: W = g v / |v| per
    row, idrlnet/architecture/layer.py:42-43), so that nested autograd gives reference gradients for weight_g /
    weight_v / bias.  (copy.deepcopy of a weight-norm module that has run is refused by torch.)"""

    def __init__(self, net, act):
        self.act = {"silu": torch.nn.functional.silu, "tanh": torch.tanh, "sin": torch.sin}[act]
        self.params = {k: p.detach().cpu().double().requires_grad_() for k, p in net.net.named_parameters()}
        self.layers = [k for k, _ in net.net.layers.items() if not k.endswith("_activation")]

    def named_parameters(self):
        return self.params.items()

    def parameters(self):
        return self.params.values()

    def __call__(self, x):
        h = x
        for i, name in enumerate(self.layers):
            q = self.params
            if f"layers.{name}.weight_g" in q:
                v = q[f"layers.{name}.weight_v"]
                w = q[f"layers.{name}.weight_g"] * v / v.norm(dim=1, keepdim=True)
            else:
                w = q[f"layers.{name}.weight"]
            h = h @ w.t() + q[f"layers.{name}.bias"]
            if i + 1 < len(self.layers):
                h = self.act(h)
        return h


def _fp64_copy(net, act="silu"):
    return _Fp64Mlp(net, act)


def _compare_with_fp64_autograd(solver, net, mod, want_fn):
    """want_fn(module64) -> FP64 loss tensor built with nested autograd on the CPU copy `mod` of the module."""
    solver.optimizers[0].zero_grad()
    loss = solver._loss_and_grads()
    assert loss.is_cuda
    got = {k: p.grad.detach().cpu().double().numpy().copy() for k, p in net.net.named_parameters()}
    want = want_fn(mod)
    want.backward()
    assert abs(float(loss) - float(want)) <= LOSS_TOL * abs(float(want)), (float(loss), float(want))
    scale = max(float(p.grad.abs().max()) for p in mod.parameters() if p.grad is not None)
    for k, p in mod.named_parameters():
        if p.grad is None:
            assert np.max(np.abs(got[k])) == 0.0, k
            continue
        w = p.grad.numpy()
        err = np.max(np.abs(got[k] - w)) / max(np.max(np.abs(w)), 1e-3 * scale)
        assert err <= GRAD_TOL, (k, err)


@pytest.mark.parametrize("expr_kind", ["xy_only", "xy_xx", "xy_xx_yy"])
def test_mixed_derivative_runs_on_the_fused_kernels(tmp_path, expr_kind):
    rng = np.random.default_rng(0)
    pts = _points(rng, 200, ("x", "y"))
    x, y = sp.symbols("x y")
    u = sp.Function("u")(x, y)
    expr = {"xy_only": u.diff(x).diff(y) * u + u.diff(y),
            "xy_xx": u.diff(x).diff(y) + u.diff(x, 2) - u,
            "xy_xx_yy": u.diff(x, 2) + u.diff(y, 2) + 0.5 * u.diff(y).diff(x)}[expr_kind]
    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x", "y"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    mod64 = _fp64_copy(net)
    eq = sc.ExpressionNode(expression=expr, name="mixed")
    s = sc.Solver(sample_domains=(_fixed_domain("dom", pts, {"mixed": 0.25}),), netnodes=[net], pdes=[eq],
                  network_dir=str(tmp_path), loading=False, max_iter=3)
    assert sorted(s._fused_domains) == ["dom"]

    def want(mod):
        xt = torch.tensor(pts["x"], dtype=torch.float64, requires_grad=True)
        yt = torch.tensor(pts["y"], dtype=torch.float64, requires_grad=True)
        uu = mod(torch.cat([xt, yt], dim=1))
        ux, uy = _grad(uu, xt), _grad(uu, yt)
        uxy = _grad(ux, yt)
        r = {"xy_only": lambda: uxy * uu + uy,
             "xy_xx": lambda: uxy + _grad(ux, xt) - uu,
             "xy_xx_yy": lambda: _grad(ux, xt) + _grad(uy, yt) + 0.5 * uxy}[expr_kind]()
        return torch.sum((r - 0.25) ** 2 * torch.tensor(pts["area"], dtype=torch.float64))

    _compare_with_fp64_autograd(s, net, mod64, want)
    assert np.isfinite(float(s.train_pipe()))


def test_euler_beam_chain_runs_on_the_fused_kernels(tmp_path):
    """examples/euler_beam/euler_beam.py:47-54: fourth derivative + 1 in the interior; value, first, second and third
    derivative on the boundaries: five domains, five jet plans of the same net, all on the fused step."""
    rng = np.random.default_rng(4)
    x = sp.Symbol("x")
    y = sp.Function("y")(x)
    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x",), outputs=("y",), name="net", arch=sc.Arch.mlp)
    mod64 = _fp64_copy(net)
    pdes = [sc.ExpressionNode(name="dddd_y", expression=y.diff(x).diff(x).diff(x).diff(x) + 1),
            sc.ExpressionNode(name="d_y", expression=y.diff(x)),
            sc.ExpressionNode(name="dd_y", expression=y.diff(x).diff(x)),
            sc.ExpressionNode(name="ddd_y", expression=y.diff(x).diff(x).diff(x))]
    spec = [("interior", _points(rng, 300, ("x",), 0.0, 1.0), "dddd_y"), ("left1", _points(rng, 40, ("x",), 0.0, 0.0), "y"),
            ("left2", _points(rng, 40, ("x",), 0.0, 0.0), "d_y"), ("right1", _points(rng, 40, ("x",), 1.0, 1.0), "dd_y"),
            ("right2", _points(rng, 40, ("x",), 1.0, 1.0), "ddd_y")]
    s = sc.Solver(sample_domains=tuple(_fixed_domain(n, p, {t: 0.0}) for n, p, t in spec), netnodes=[net], pdes=pdes,
                  network_dir=str(tmp_path), loading=False, max_iter=3)
    assert sorted(s._fused_domains) == sorted(n for n, _, _ in spec)

    def want(mod):
        total = 0.0
        for n, p, t in spec:
            xt = torch.tensor(p["x"], dtype=torch.float64, requires_grad=True)
            d = [mod(xt)]
            for _ in range(4):
                d.append(_grad(d[-1], xt))
            r = {"y": d[0], "d_y": d[1], "dd_y": d[2], "ddd_y": d[3], "dddd_y": d[4] + 1.0}[t]
            total = total + torch.sum(r ** 2 * torch.tensor(p["area"], dtype=torch.float64))
        return total

    _compare_with_fp64_autograd(s, net, mod64, want)
    l0, l1 = float(s.train_pipe()), float(s.train_pipe())
    assert np.isfinite(l0) and np.isfinite(l1)


@pytest.mark.parametrize("act", ["tanh", "sin"])
def test_third_order_jets_are_served_to_torch_residuals(tmp_path, act):
    """A non-polynomial residual of a third derivative: the net serves y, y__x, y__x__x__x from ONE jet kernel call
    (JetFunction, generic path) and torch evaluates sin(.) on top -- no nested autograd through the net."""
    from idrlnet_b200 import architecture as arch
    rng = np.random.default_rng(5)
    pts = _points(rng, 150, ("x",))
    x = sp.Symbol("x")
    y = sp.Function("y")(x)
    torch.manual_seed(1)
    mod = arch.MLP([1, 20, 20, 20, 1], activation={"tanh": arch.Activation.tanh, "sin": arch.Activation.sin}[act],
                   weight_norm=False)
    net = sc.NetNode(inputs=("x",), outputs=("y",), net=mod, name="net")
    mod64 = _fp64_copy(net, act)
    eq = sc.ExpressionNode(name="res", expression=sp.sin(0.1 * y.diff(x, 3)) - y.diff(x) * y)
    s = sc.Solver(sample_domains=(_fixed_domain("dom", pts, {"res": 0.0}),), netnodes=[net], pdes=[eq],
                  network_dir=str(tmp_path), loading=False, max_iter=3)
    assert "dom" not in s._fused_domains and s.vertex_pipelines["dom"]._jet_plans  # jets from the kernel, residual in torch

    def want(m64):
        xt = torch.tensor(pts["x"], dtype=torch.float64, requires_grad=True)
        d = [m64(xt)]
        for _ in range(3):
            d.append(_grad(d[-1], xt))
        r = torch.sin(0.1 * d[3]) - d[1] * d[0]
        return torch.sum(r ** 2 * torch.tensor(pts["area"], dtype=torch.float64))

    _compare_with_fp64_autograd(s, net, mod64, want)


def test_third_order_mixed_derivative_falls_back_to_autograd(tmp_path):
    """What still has no fused channel (u__x__x__y) runs through the generic graph with torch autograd, on the GPU."""
    rng = np.random.default_rng(0)
    pts = _points(rng, 100, ("x", "y"))
    x, y = sp.symbols("x y")
    u = sp.Function("u")(x, y)
    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x", "y"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    mod64 = _fp64_copy(net)
    eq = sc.ExpressionNode(expression=u.diff(x, 2).diff(y) - u, name="mixed3")
    s = sc.Solver(sample_domains=(_fixed_domain("dom", pts, {"mixed3": 0.0}),), netnodes=[net], pdes=[eq],
                  network_dir=str(tmp_path), loading=False)
    assert "dom" not in s._fused_domains and not s.vertex_pipelines["dom"]._jet_plans

    def want(mod):
        xt = torch.tensor(pts["x"], dtype=torch.float64, requires_grad=True)
        yt = torch.tensor(pts["y"], dtype=torch.float64, requires_grad=True)
        uu = mod(torch.cat([xt, yt], dim=1))
        r = _grad(_grad(_grad(uu, xt), xt), yt) - uu
        return torch.sum(r ** 2 * torch.tensor(pts["area"], dtype=torch.float64))

    _compare_with_fp64_autograd(s, net, mod64, want)
