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


def test_mixed_derivative_falls_back_to_autograd(tmp_path):
    rng = np.random.default_rng(0)
    pts = _points(rng, 200)
    x, y = sp.symbols("x y")
    u = sp.Function("u")(x, y)
    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x", "y"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    eq = sc.ExpressionNode(expression=u.diff(x).diff(y) + u.diff(x, 2) - u, name="mixed")
    s = sc.Solver(sample_domains=(_fixed_domain("dom", pts, {"mixed": 0.0}),), netnodes=[net], pdes=[eq],
                  network_dir=str(tmp_path), loading=False)
    assert "dom" not in s._fused_domains  # u__x__y has no fused channel
    s.optimizers[0].zero_grad()
    loss = s._loss_and_grads()
    assert loss.is_cuda and float(loss) > 0
    assert all(p.grad is not None and p.grad.is_cuda for p in net.net.parameters() if p.requires_grad)
    # same numbers as the oracle (reference algorithm)
    problem = {"nets": [{"name": "net1", "inputs": ["x", "y"], "outputs": ["u"], "act": "silu", "shares": None,
                         "layers": [{"g": l.weight_g.detach().cpu().numpy(), "v": l.weight_v.detach().cpu().numpy(),
                                     "b": l.bias.detach().cpu().numpy()}
                                    for k, l in net.net.layers.items() if not k.endswith("_activation")]}],
               "exprs": {"mixed": "u__x__y + u__x__x - u"},
               "domains": [{"name": "dom", "points": pts, "targets": {"mixed": 0.0}, "lambdas": {}, "loss": "square", "sigma": 1.0}]}
    ref = po.training_step(problem, dtype=torch.float64)
    assert abs(float(loss) - ref["loss"]) <= 1e-4 * abs(ref["loss"])


def test_non_mlp_net_runs_on_torch_path(tmp_path):
    rng = np.random.default_rng(1)
    pts = _points(rng, 128)
    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x", "y"), outputs=("T",), name="siren", arch=sc.Arch.siren, seq=[2, 32, 32, 1])
    pde = sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False)
    s = sc.Solver(sample_domains=(_fixed_domain("dom", pts, {"diffusion_T": 1.0}),), netnodes=[net], pdes=[pde],
                  network_dir=str(tmp_path), loading=False, max_iter=3)
    assert not s._fused_domains  # Siren scales its pre-activations: not a plain Linear stack
    l0 = float(s.train_pipe())
    l1 = float(s.train_pipe())
    assert np.isfinite(l0) and np.isfinite(l1)


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
