"""What the fused kernels cannot express must still run, through the generic graph with the
reference's semantics (torch autograd), on the GPU -- never silently on the CPU."""
import numpy as np
import pytest
import sympy as sp
import torch

import idrlnet_b200.shortcut as sc
from oracle import pinn_oracle as po
from tests import problem_util as pu

pytestmark = pytest.mark.gpu


def _fixed_domain(name, pts, cons, **kw):
    return sc.get_data_node((lambda p=pts, c=cons: (dict(p), dict(c))), name=name, **kw)


def _points(rng, n, keys=("x", "y")):
    p = {k: rng.uniform(-1, 1, (n, 1)).astype(np.float32) for k in keys}
    p["area"] = np.full((n, 1), 1.0 / n, np.float32)
    return p


# (mixed second derivatives now have a fused channel: tests/test_gpu_f4.py holds the remaining fallback case)


@pytest.mark.parametrize("omegas", [(30.0, 30.0), (3.0, 2.0)])
def test_siren_runs_on_the_fused_kernels(tmp_path, omegas):
    """Siren (idrlnet/architecture/mlp.py:79-137) = sin MLP whose hidden pre-activations are scaled by omega;
    the scale is folded into the effective parameters the kernels read, so the fused step applies.  Loss and raw
    parameter gradients against torch autograd in FP64 on the same module."""
    rng = np.random.default_rng(1)
    pts = _points(rng, 128)
    torch.manual_seed(0)
    from idrlnet_b200.architecture.core import Siren
    mod = Siren([2, 32, 32, 1], first_omega=omegas[0], omega=omegas[1])
    net = sc.NetNode(inputs=("x", "y"), outputs=("T",), net=mod, name="siren")
    pde = sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False)
    s = sc.Solver(sample_domains=(_fixed_domain("dom", pts, {"diffusion_T": 1.0}),), netnodes=[net], pdes=[pde],
                  network_dir=str(tmp_path), loading=False, max_iter=3)
    assert sorted(s._fused_domains) == ["dom"]
    s.optimizers[0].zero_grad()
    loss = float(s._loss_and_grads())
    got = {k: p.grad.detach().cpu().double().numpy().copy() for k, p in mod.named_parameters()}
    # FP64 autograd reference on a CPU copy of the module
    import copy
    ref = copy.deepcopy(mod).cpu().double()
    x = torch.tensor(pts["x"], dtype=torch.float64, requires_grad=True)
    y = torch.tensor(pts["y"], dtype=torch.float64, requires_grad=True)
    T = ref(torch.cat([x, y], dim=1))
    g = lambda a, b: torch.autograd.grad(a, b, torch.ones_like(a), create_graph=True)[0]
    r = -g(g(T, x), x) - g(g(T, y), y)
    area = torch.tensor(pts["area"], dtype=torch.float64)
    want = torch.sum((r - 1.0) ** 2 * area)
    want.backward()
    assert abs(loss - float(want)) <= 2e-5 * abs(float(want)), (loss, float(want))
    scale = max(float(p.grad.abs().max()) for p in ref.parameters() if p.grad is not None)
    for k, p in ref.named_parameters():
        if p.grad is None:  # output bias: only derivatives of T enter the residual -> exact zeros here
            assert np.max(np.abs(got[k])) == 0.0, k
            continue
        w = p.grad.numpy()
        err = np.max(np.abs(got[k] - w)) / max(np.max(np.abs(w)), 1e-3 * scale)
        assert err <= 2e-4, (k, err)
    l0 = float(s.train_pipe())
    l1 = float(s.train_pipe())
    assert np.isfinite(l0) and np.isfinite(l1)


def test_derivative_of_a_downstream_expression_keeps_the_chain_rule(tmp_path):
    """a constraint on `flux__x`, where `flux` is an ExpressionNode built from the net output, needs the
    chain rule THROUGH the net (idrlnet/variable.py:161-208 differentiates whatever the graph produced).  A jet-served
    net has no autograd edge to the coordinates, so such a pipeline must keep the torch path for that net instead of
    silently differentiating to zero."""
    import sympy as sp
    rng = np.random.default_rng(3)
    pts = _points(rng, 64)
    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x", "y"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    xs, u = sp.Symbol("x"), sp.Function("u")(sp.Symbol("x"), sp.Symbol("y"))
    flux = sc.ExpressionNode(name="flux", expression=u * u + 2 * xs)
    s = sc.Solver(sample_domains=(_fixed_domain("dom", pts, {"flux__x": 0.0}),), netnodes=[net], pdes=[flux],
                  network_dir=str(tmp_path), loading=False, max_iter=3, post_step_loss=False)
    assert not s._fused_domains and not s.vertex_pipelines["dom"]._jet_plans  # torch path: autograd keeps the chain
    loss = float(s.train_pipe())
    # train_pipe (fast mode) returned the loss at the PRE-update parameters: recompute it in FP64 with autograd
    torch.manual_seed(0)
    net0 = sc.get_net_node(inputs=("x", "y"), outputs=("u",), name="net1b", arch=sc.Arch.mlp).net.cpu().double()
    x = torch.tensor(pts["x"], dtype=torch.float64, requires_grad=True)
    y = torch.tensor(pts["y"], dtype=torch.float64)
    uu = net0(torch.cat([x, y], dim=1))
    fx = torch.autograd.grad(uu * uu + 2 * x, x, torch.ones_like(uu))[0]
    want = float(torch.sum(fx ** 2 * torch.tensor(pts["area"], dtype=torch.float64)))
    assert want > 0 and abs(loss - want) <= 1e-4 * want, (loss, want)


def test_lbfgs_closure_on_fused_step(tmp_path):
    rng = np.random.default_rng(2)
    pts = _points(rng, 512, keys=("x", "t"))
    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    s = sc.Solver(sample_domains=(_fixed_domain("dom", pts, {"burgers_u": 0.0}),
                                  _fixed_domain("ic", {**_points(rng, 64, keys=("x", "t"))}, {"u": 0.5})),
                  netnodes=[net], pdes=[sc.BurgersNode(u="u", v=0.01)], network_dir=str(tmp_path), loading=False,
                  opt_config=dict(optimizer="LBFGS", lr=1, max_iter=5))
    assert len(s._fused_domains) == 2
    losses = [float(s.train_pipe()) for _ in range(4)]
    assert losses[-1] < losses[0]


def test_product_path_refuses_cpu_tensors():
    from idrlnet_b200 import fused
    net = sc.MLP([2, 20, 20, 1])
    with pytest.raises(RuntimeError):
        fused.NetPack(net)  # CPU module: there is no CPU implementation to fall back to
