"""The public API on the GPU: Solver built from the reference's configurations runs the
fused kernels, matches the oracle on one step, and tracks the reference training loop
(oracle + Adam + ExponentialLR) over 1k steps."""
import numpy as np
import pytest
import torch

from oracle import pinn_oracle as po
from tests import golden_util as gu
from tests import problem_util as pu

pytestmark = pytest.mark.gpu


def _param_grads_vs_golden(modules, flat, tol):
    for pname, g in gu.ref_grads(flat, "ref64").items():
        net, rest = pname.split("/", 1)
        p = dict(modules[net].named_parameters())[rest]
        err = np.max(np.abs(p.grad.detach().cpu().numpy().reshape(g.shape) - g)) / max(np.max(np.abs(g)), 1e-30)
        assert err <= tol, (pname, err)


@pytest.mark.parametrize("name,n_fused", [("simple_poisson", 3), ("burgers", 3), ("ldc", 3), ("ns_unsteady", 2)])
def test_solver_step_matches_reference(name, n_fused, tmp_path):
    problem, flat = gu.load(name)
    s, modules = pu.solver_from_problem(problem, tmp_path)
    assert len(s._fused_domains) == n_fused  # every domain runs the fused kernels
    s.optimizers[0].zero_grad()
    loss = float(s._loss_and_grads())
    ref = float(flat["ref64/loss"])
    assert abs(loss - ref) <= 1e-5 * abs(ref)
    for k, v in gu.ref_components(flat, "ref64").items():
        assert abs(float(s.loss_component[k]) - v) <= 1e-5 * max(abs(v), 1e-12), k
    _param_grads_vs_golden(modules, flat, 2e-4)


def test_solver_mixed_fused_and_generic_allen_cahn(tmp_path):
    """Allen-Cahn: interior / IC / resampling domains fused; the periodic-BC domain (two
    evaluations of a shared net, ExpressionNodes, Difference) on the generic graph with the
    jets served by the kernel (no autograd.grad w.r.t. coordinates)."""
    problem, flat = gu.load("allen_cahn")
    s, modules = pu.solver_from_problem(problem, tmp_path)
    assert sorted(s._fused_domains) == ["AllenInit", "allen_domain", "re_sampling_domain"]
    assert len(s.vertex_pipelines["AllenBc"]._jet_plans) == 2  # net1 and the shared net2 serve jets
    s.optimizers[0].zero_grad()
    loss = float(s._loss_and_grads())
    ref = float(flat["ref64/loss"])
    assert abs(loss - ref) <= 1e-5 * abs(ref)
    _param_grads_vs_golden(modules, flat, 2e-4)


@pytest.mark.parametrize("name", ["simple_poisson", "ldc"])
def test_loss_curve_tracks_reference_training(name, tmp_path):
    """SURVEY 8(c6) / BASELINE.md 3.4: 1k-step loss curves on simple_poisson AND ldc.  The Solver runs with its
    defaults, i.e. the reference's `train_pipe` semantics (idrlnet/solver.py:341-344): the loss returned by step i
    is the gradient-free re-evaluation AFTER update i -- with the fixed points of these problems that is the
    pre-update loss of step i+1 of the reference training loop (oracle + Adam + ExponentialLR on the CPU).

    Tie-breaker: the same loop is also run in FP64.  Wherever the reference algorithm's own FP32 and
    FP64 curves agree (1e-5), the kernels must match the FP32 curve to 2e-4.  On ldc the two reference curves sit on a
    plateau (8.78e-4) until about step 600 and then leave it at rounding-dependent times -- they differ from EACH OTHER
    by up to 47 % between steps 650 and 1000 -- so there the kernels are held to that envelope, not to 2e-4."""
    steps, tol = 1000, 2e-4
    problem, _ = gu.load(name)
    ref32 = np.asarray(pu.oracle_train(problem, steps + 1))
    ref64 = np.asarray(pu.oracle_train(problem, steps + 1, dtype=torch.float64))
    env = np.abs(ref32 - ref64)[1:] / np.abs(ref64[1:])
    stable = int(np.argmax(env > 1e-5)) if (env > 1e-5).any() else steps  # steps before the reference turns chaotic
    assert stable >= 500, (name, stable)
    s, _ = pu.solver_from_problem(problem, tmp_path, max_iter=steps, save_freq=10 ** 9, print_freq=10 ** 9)
    assert s.post_step_loss  # the drop-in default is the reference's
    ours = []
    for _ in range(steps):
        ours.append(s.train_pipe())
    ours = torch.stack([l.reshape(()) for l in ours]).cpu().numpy()
    rel = np.abs(ours - ref32[1:]) / np.abs(ref32[1:])
    print(f"{name}: reference FP32 vs FP64 agree to 1e-5 for {stable} steps (max separation {env.max():.2e}); "
          f"ours vs reference FP32: {rel[:stable].max():.2e} there, {rel.max():.2e} over all {steps} steps")
    assert ours[-1] < ref32[0]
    assert rel[:stable].max() <= tol, (int(rel[:stable].argmax()), float(rel[:stable].max()), ours[:3], ref32[:4])
    if stable < steps:
        rel64 = np.abs(ours - ref64[1:]) / np.abs(ref64[1:])
        assert min(rel.max(), rel64.max()) <= 2.0 * env.max(), (float(rel.max()), float(rel64.max()), float(env.max()))
    else:
        assert rel.max() <= tol


def test_fast_mode_returns_the_pre_update_loss(tmp_path):
    """Solver(post_step_loss=False): one evaluation per step; the returned loss is the one the gradient was taken of."""
    problem, _ = gu.load("burgers")
    ref_losses = np.asarray(pu.oracle_train(problem, 20))
    s, _ = pu.solver_from_problem(problem, tmp_path, post_step_loss=False)
    ours = np.array([float(s.train_pipe()) for _ in range(20)])
    assert np.max(np.abs(ours - ref_losses) / np.abs(ref_losses)) <= 2e-4


def test_grads_survive_the_post_step_evaluation(tmp_path):
    """`.grad` views of the engine's flat buffer must not be garbled by the loss-only re-evaluation
    that ends every default train_pipe (receivers / clipping read the grads at TRAIN_PIPE_END)."""
    problem, flat = gu.load("burgers")
    s, modules = pu.solver_from_problem(problem, tmp_path, opt_config=dict(optimizer="SGD", lr=0.0))
    s.train_pipe()  # lr = 0: parameters unchanged, so the grads must equal the golden ones after the whole pipe
    _param_grads_vs_golden(modules, flat, 2e-4)


def test_lbfgs_closure_tracks_reference(tmp_path):
    """the LBFGS closure branch (idrlnet/solver.py:316-326) on the fused step.  LBFGS
    amplifies rounding differences through its line search and curvature pairs, so the band is wider
    than Adam's: 1e-3 relative on the per-step loss over 6 outer steps x 5 inner iterations."""
    problem, _ = gu.load("burgers")
    cfg = dict(lr=1.0, max_iter=5)
    ref_losses = pu.oracle_train(problem, 6, lbfgs=cfg)
    s, _ = pu.solver_from_problem(problem, tmp_path, opt_config=dict(optimizer="LBFGS", **cfg), post_step_loss=False)
    assert len(s._fused_domains) == len(problem["domains"])
    ours = np.array([float(s.train_pipe()) for _ in range(6)])
    rel = np.abs(ours - np.asarray(ref_losses)) / np.abs(ref_losses)
    assert ours[-1] < ours[0]
    assert rel.max() <= 1e-3, (rel, ours, ref_losses)


def test_infer_step_serves_derivatives(tmp_path):
    problem, flat = gu.load("burgers")
    s, _ = pu.solver_from_problem(problem, tmp_path)
    out = s.infer_step({"burgers_equation": ["x", "t", "u", "u__x", "u__x__x", "burgers_u"]})
    ref = gu.ref_domain(flat, "ref64", "burgers_equation")
    for k in ("u", "u__x", "u__x__x", "burgers_u"):
        got = out["burgers_equation"][k].detach().cpu().numpy()
        assert gu.normwise(got, ref[k]) <= 1e-5, k


def test_solver_tf32_precision_tracks_fp32(tmp_path):
    """Solver(precision="tf32"): tensor-core path through the public API; the loss curve stays
    within the stated TF32 band of the FP32 run (ldc, 100 steps)."""
    problem, _ = gu.load("ldc")
    curves = {}
    for prec in ("fp32", "tf32"):
        s, _ = pu.solver_from_problem(problem, tmp_path / prec, max_iter=100, save_freq=10 ** 9, print_freq=10 ** 9,
                                      precision=prec)
        assert len(s._fused_domains) == 3
        curves[prec] = torch.stack([s.train_pipe().reshape(()) for _ in range(100)]).cpu().numpy()
    rel = np.abs(curves["tf32"] - curves["fp32"]) / np.abs(curves["fp32"])
    assert 0 < rel.max() <= 2e-2, float(rel.max())


def test_solver_trains_on_device_sampled_points(tmp_path):
    """a datanode may return device tensors drawn by the CSG kernel; fresh points every step,
    nothing sampled on the host, every domain on the fused kernels."""
    import idrlnet_b200.shortcut as sc
    geo = sc.Rectangle((-1.0, -1.0), (1.0, 1.0)) - sc.Circle((0.0, 0.0), 0.3)
    state = {"step": 0}

    @sc.datanode(name="interior")
    class Interior(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            state["step"] += 1
            pts = geo.sample_interior_device(2000, seed=3, offset=state["step"] * 8000)
            return pts, {"diffusion_T": 1.0}

    @sc.datanode(name="wall")
    class Wall(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return geo.sample_boundary(100), {"T": 0.0}

    torch.manual_seed(0)
    np.random.seed(0)
    net = sc.get_net_node(inputs=("x", "y"), outputs=("T",), name="net1", arch=sc.Arch.mlp)
    s = sc.Solver(sample_domains=(Interior(), Wall()), netnodes=[net],
                  pdes=[sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False)], network_dir=str(tmp_path), loading=False,
                  max_iter=100)
    assert sorted(s._fused_domains) == ["interior", "wall"]
    losses = [float(s.train_pipe()) for _ in range(60)]
    assert np.isfinite(losses).all() and np.mean(losses[-10:]) < 0.5 * np.mean(losses[:5])


def test_poisson_example_converges_to_the_analytic_solution(tmp_path):
    """End-to-end physics check: the simple_poisson example as shipped (-Laplace T = 1 on [-1,1]^2,
    T = 0 at x = +-1, dT/dn = 0 at y = +-1), 1500 Adam steps on the fused kernels with fresh points every step,
    against T = (1 - x^2) / 2."""
    import sympy as sp
    import idrlnet_b200.shortcut as sc
    x, y = sp.symbols("x y")
    rec = sc.Rectangle((-1.0, -1.0), (1.0, 1.0))

    @sc.datanode
    class LeftRight(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_boundary(1000, sieve=((y > -1.0) & (y < 1.0))), {"T": 0.0}

    @sc.datanode(name="up_down")
    class UpDown(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_boundary(1000, sieve=((x > -1.0) & (x < 1.0))), {"normal_gradient_T": 0.0}

    @sc.datanode(name="heat_domain")
    class Heat(sc.SampleDomain):
        def __init__(self):
            self.points = 1000

        def sampling(self, *args, **kwargs):
            return rec.sample_interior(self.points), {"diffusion_T": 1.0}

    torch.manual_seed(0)
    np.random.seed(0)
    net = sc.get_net_node(inputs=("x", "y"), outputs=("T",), name="net1", arch=sc.Arch.mlp)
    s = sc.Solver(sample_domains=(Heat(), LeftRight(), UpDown()), netnodes=[net],
                  pdes=[sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False), sc.NormalGradient("T", dim=2, time=False)],
                  max_iter=1500, network_dir=str(tmp_path), loading=False, save_freq=10 ** 9, print_freq=10 ** 9)
    assert len(s._fused_domains) == 3
    s.solve()
    s.set_domain_parameter("heat_domain", {"points": 2500})
    out = s.infer_step({"heat_domain": ["x", "y", "T"]})["heat_domain"]
    xs = out["x"].detach().cpu().numpy().ravel()
    T = out["T"].detach().cpu().numpy().ravel()
    err = np.abs(T - (1.0 - xs ** 2) / 2.0).max()
    assert err < 0.05, err
