"""Seeded example runs shared by the reference (golden generator) and idrlnet_b200 (tests): `CASES[name](sc,
network_dir, iters)` returns (net node, Solver kwargs).  Sizes are cut down so the CPU tier stays fast."""
import math

import numpy as np
import sympy as sp
import torch


def poisson(sc, network_dir, iters):
    """examples/simple_poisson: Rectangle boundary with sieves, Neumann condition through NormalGradient."""
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
            self.points = 100

        def sampling(self, *args, **kwargs):
            return rec.sample_interior(self.points), {"diffusion_T": 1.0}

    np.random.seed(7)
    torch.manual_seed(7)
    net = sc.get_net_node(inputs=("x", "y"), outputs=("T",), name="net1", arch=sc.Arch.mlp)
    pde = sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False)
    grad = sc.NormalGradient("T", dim=2, time=False)
    return net, dict(sample_domains=(Heat(), LeftRight(), UpDown()), netnodes=[net], pdes=[pde, grad], max_iter=iters,
                     network_dir=network_dir, loading=False)


def burgers(sc, network_dir, iters):
    """examples/burgers_equation: Line1D with a time parameter range, symbolic initial condition."""
    x, t = sp.Symbol("x"), sp.Symbol("t")
    geo = sc.Line1D(-1.0, 1.0)

    @sc.datanode(name="burgers_equation")
    def interior_domain():
        return geo.sample_interior(400, bounds={x: (-1.0, 1.0)}, param_ranges={t: (0, 1)}), {"burgers_u": 0}

    @sc.datanode(name="t_boundary")
    def init_domain():
        return geo.sample_interior(40, param_ranges={t: 0.0}), sc.Variables({"u": -sp.sin(math.pi * x)})

    @sc.datanode(name="x_boundary")
    def boundary_domain():
        return geo.sample_boundary(40, param_ranges={t: (0, 1)}), sc.Variables({"u": 0})

    np.random.seed(11)
    torch.manual_seed(11)
    net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    return net, dict(sample_domains=(interior_domain(), init_domain(), boundary_domain()), netnodes=[net],
                     pdes=[sc.BurgersNode(u="u", v=0.01 / math.pi)], max_iter=iters, network_dir=network_dir, loading=False)


def cavity(sc, network_dir, iters):
    """examples/ldc in miniature: tanh MLP with Xavier init and no weight-norm, three outputs, Navier-Stokes residuals
    weighted by the signed distance, lid condition with a symbolic weight."""
    x, y = sp.symbols("x y")
    rec = sc.Rectangle((-0.05, -0.05), (0.05, 0.05))

    @sc.datanode(name="flow")
    class Flow(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            pts = rec.sample_interior(40000)
            return pts, {"continuity": 0.0, "momentum_x": 0.0, "momentum_y": 0.0, "lambda_continuity": pts["sdf"],
                         "lambda_momentum_x": pts["sdf"], "lambda_momentum_y": pts["sdf"]}

    @sc.datanode(name="walls")
    class Walls(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_boundary(400, sieve=(y < 0.05)), {"u": 0.0, "v": 0.0}

    @sc.datanode(name="lid")
    class Lid(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_boundary(400, sieve=sp.Eq(y, 0.05)), {"u": 1.0, "v": 0.0, "lambda_u": 1 - 20 * abs(x)}

    np.random.seed(5)
    torch.manual_seed(5)
    mlp = sc.MLP([2, 32, 32, 3], activation=sc.Activation.tanh, initialization=sc.Initializer.Xavier_uniform, weight_norm=False)
    net = sc.NetNode(inputs=("x", "y"), outputs=("u", "v", "p"), net=mlp, name="net_u")
    pde = sc.NavierStokesNode(nu=0.01, rho=1.0, dim=2, time=False)
    return net, dict(sample_domains=(Flow(), Walls(), Lid()), netnodes=[net], pdes=[pde], max_iter=iters,
                     network_dir=network_dir, loading=False)


def beam_lbfgs(sc, network_dir, iters):
    """examples/euler_beam/euler_beam_lbfgs.py: fourth-order derivatives through chained ExpressionNodes, Eq sieves on a
    Line1D boundary, LBFGS closure (idrlnet/solver.py:316-326)."""
    x = sp.symbols("x")
    line = sc.Line1D(0, 1)
    y = sp.Function("y")(x)

    def dom(name, sampler, cons):
        @sc.datanode(name=name)
        class D(sc.SampleDomain):
            def sampling(self, *args, **kwargs):
                return sampler(), dict(cons)
        return D()

    doms = (dom("interior", lambda: line.sample_interior(100), {"dddd_y": 0}),
            dom("left_boundary1", lambda: line.sample_boundary(10, sieve=sp.Eq(x, 0)), {"y": 0}),
            dom("left_boundary2", lambda: line.sample_boundary(10, sieve=sp.Eq(x, 0)), {"d_y": 0}),
            dom("right_boundary1", lambda: line.sample_boundary(10, sieve=sp.Eq(x, 1)), {"dd_y": 0}),
            dom("right_boundary2", lambda: line.sample_boundary(10, sieve=sp.Eq(x, 1)), {"ddd_y": 0}))
    np.random.seed(3)
    torch.manual_seed(3)
    net = sc.get_net_node(inputs=("x",), outputs=("y",), name="net", arch=sc.Arch.mlp)
    pdes = [sc.ExpressionNode(name="dddd_y", expression=y.diff(x).diff(x).diff(x).diff(x) + 1),
            sc.ExpressionNode(name="d_y", expression=y.diff(x)),
            sc.ExpressionNode(name="dd_y", expression=y.diff(x).diff(x)),
            sc.ExpressionNode(name="ddd_y", expression=y.diff(x).diff(x).diff(x))]
    return net, dict(sample_domains=doms, netnodes=[net], pdes=pdes, max_iter=iters, network_dir=network_dir, loading=False,
                     opt_config=dict(optimizer="LBFGS", lr=1))


def periodic(sc, network_dir, iters):
    """examples/allen_cahn in miniature: a net shared under a second name on shifted inputs (`get_shared_net_node`),
    periodic conditions through ExpressionNode + Difference, a weighted initial condition."""
    x, t = sp.Symbol("x"), sp.Symbol("t")
    u, up = sp.Function("u")(x, t), sp.Function("up")(x, t)
    geo = sc.Line1D(-1.0, 1.0)

    @sc.datanode(name="allen_domain")
    def interior():
        return geo.sample_interior(200, bounds={x: (-1.0, 1.0)}, param_ranges={t: (0, 1)}), {"AllenCahn_u": 0}

    @sc.datanode(name="AllenInit")
    def init():
        pts = geo.sample_interior(40, param_ranges={t: 0.0})
        return pts, {"u": x ** 2 * sp.cos(sp.pi * x), "lambda_u": 100}

    @sc.datanode(name="AllenBc")
    def bc():
        pts = geo.sample_boundary(20, sieve=sp.Eq(x, -1.0), param_ranges={t: (0, 1)})
        return pts, {"difference_u_up": 0, "difference_diff_u_diff_up": 0}

    np.random.seed(13)
    torch.manual_seed(13)
    mlp = sc.MLP([2, 24, 24, 2], activation=sc.Activation.tanh)
    net_u = sc.NetNode(inputs=("x", "t"), outputs=("u",), name="net1", net=mlp)
    xp = sc.ExpressionNode(name="xp", expression=x + 2)
    net_up = sc.get_shared_net_node(net_u, inputs=("xp", "t"), outputs=("up",), name="net2", arch="mlp")
    pdes = [xp, sc.ExpressionNode(expression=u.diff(x), name="diff_u"), sc.ExpressionNode(expression=up.diff(x), name="diff_up"),
            sc.AllenCahnNode(u="u", gamma_1=0.0001, gamma_2=5), sc.Difference(T="diff_u", S="diff_up"),
            sc.Difference(T="u", S="up")]
    return net_u, dict(sample_domains=(interior(), init(), bc()), netnodes=[net_u, net_up], pdes=pdes, max_iter=iters,
                       network_dir=network_dir, loading=False)


def inverse_wave(sc, network_dir, iters):
    """examples/inverse_wave_equation: a second "net" that is a single trainable coefficient (`Arch.single_var`), an L1
    data loss on low-discrepancy samples, a PDE whose coefficient is the other net's output."""
    x, t = sp.Symbol("x"), sp.Symbol("t")
    L = float(math.pi)
    geo = sc.Line1D(0, L)
    np.random.seed(17)
    obs = geo.sample_interior(density=4, bounds={x: (0, L)}, param_ranges={t: (0, 2 * L)}, low_discrepancy=True)
    obs_u = np.sin(obs["x"]) * (np.sin(1.54 * obs["t"]) + np.cos(1.54 * obs["t"]))

    @sc.datanode(name="wave_domain", loss_fn="L1")
    class WaveExternal(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return dict(obs), {"u": obs_u}

    @sc.datanode(name="wave_external")
    class WaveEq(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return geo.sample_interior(density=20, bounds={x: (0, L)}, param_ranges={t: (0, 2 * L)}), {"wave_equation": 0.0}

    torch.manual_seed(17)
    net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    var_c = sc.get_net_node(inputs=("x",), outputs=("c",), arch=sc.Arch.single_var)
    pde = sc.WaveNode(c="c", dim=1, time=True, u="u")
    return net, dict(sample_domains=(WaveExternal(), WaveEq()), netnodes=[net, var_c], pdes=[pde], max_iter=iters,
                     network_dir=network_dir, loading=False)


def volterra(sc, network_dir, iters):
    """examples/Volterra_IDE: an integral operator (Int1DNode, Gauss-Legendre quadrature with the net evaluated at the
    nodes) inside the training loop."""
    x, s_, fs = sp.Symbol("x"), sp.Symbol("s"), sp.Symbol("fs")
    f = sp.Function("f")(x)
    geo = sc.Line1D(0, 5)

    @sc.datanode
    def interior():
        return geo.sample_interior(40), {"difference_lhs_rhs": 0}

    @sc.datanode
    def init():
        pts = geo.sample_boundary(1, sieve=sp.Eq(x, 0))
        pts["lambda_f"] = 1000 * np.ones_like(pts["x"])
        return pts, {"f": 1}

    np.random.seed(19)
    torch.manual_seed(19)
    netnode = sc.get_net_node(inputs=("x",), outputs=("f",), name="net")
    lhs = sc.ExpressionNode(expression=f.diff(x) + f, name="lhs")
    rhs = sc.Int1DNode(expression=sp.exp(s_ - x) * fs, var=s_, lb=0, ub=x, expression_name="rhs",
                       funs={"fs": {"eval": netnode, "input_map": {"x": "s"}, "output_map": {"f": "fs"}}}, degree=10)
    diff = sc.Difference(T="lhs", S="rhs", dim=1, time=False)
    return netnode, dict(sample_domains=(interior(), init()), netnodes=[netnode], pdes=[lhs, rhs, diff], max_iter=iters,
                         network_dir=network_dir, loading=False)


def deepritz(sc, network_dir, iters):
    """examples/deepritz: energy functional through ICNode with the "Identity" loss, a sigma-weighted boundary domain,
    fixed point sets drawn once in the constructors."""
    x, y = sp.symbols("x y")
    u = sp.Function("u")(x, y)
    geo = sc.Rectangle((-1, -1), (1.0, 1.0))
    np.random.seed(23)

    @sc.datanode(sigma=1000.0)
    class Boundary(sc.SampleDomain):
        def __init__(self):
            self.points = geo.sample_boundary(25)
            self.constraints = {"u": 0.0}

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    @sc.datanode(loss_fn="Identity")
    class Interior(sc.SampleDomain):
        def __init__(self):
            self.points = geo.sample_interior(60)
            self.constraints = {"integral_dxdy": 0}

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    domains = (Boundary(), Interior())
    torch.manual_seed(23)
    f = 2 * sp.pi ** 2 * sp.sin(sp.pi * x) * sp.sin(sp.pi * y)
    dx_exp = sc.ExpressionNode(expression=0.5 * (u.diff(x) ** 2 + u.diff(y) ** 2) - u * f, name="dxdy")
    net = sc.get_net_node(inputs=("x", "y"), outputs=("u",), name="net", arch=sc.Arch.mlp)
    integral = sc.ICNode("dxdy", dim=2, time=False)
    return net, dict(sample_domains=domains, netnodes=[net], pdes=[dx_exp, integral], max_iter=iters,
                     network_dir=network_dir, loading=False)


def param_poisson(sc, network_dir, iters):
    """examples/parameterized_poisson: a design parameter as a third net input, parameter ranges on boundary samples,
    a boundary target that is the parameter symbol itself, Eq sieves."""
    x, y, temp = sp.symbols("x y temp")
    temp_range = {temp: (-0.2, 0.2)}
    rec = sc.Rectangle((-1.0, -1.0), (1.0, 1.0))

    @sc.datanode
    class Right(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_boundary(40, sieve=(sp.Eq(x, 1.0)), param_ranges=temp_range), sc.Variables({"T": 0.0})

    @sc.datanode
    class Left(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return rec.sample_boundary(40, sieve=(sp.Eq(x, -1.0)), param_ranges=temp_range), sc.Variables({"T": temp})

    @sc.datanode(name="up_down")
    class UpDown(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return (rec.sample_boundary(40, sieve=((x > -1.0) & (x < 1.0)), param_ranges=temp_range),
                    sc.Variables({"normal_gradient_T": 0.0}))

    @sc.datanode(name="heat_domain")
    class Heat(sc.SampleDomain):
        def __init__(self):
            self.points = 60

        def sampling(self, *args, **kwargs):
            return rec.sample_interior(self.points, param_ranges=temp_range), sc.Variables({"diffusion_T": 1.0})

    np.random.seed(29)
    torch.manual_seed(29)
    net = sc.get_net_node(inputs=("x", "y", "temp"), outputs=("T",), name="net1", arch=sc.Arch.mlp)
    pde = sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False)
    grad = sc.NormalGradient("T", dim=2, time=False)
    return net, dict(sample_domains=(Heat(), Left(), Right(), UpDown()), netnodes=[net], pdes=[pde, grad], max_iter=iters,
                     network_dir=network_dir, loading=False)


def min_surface(sc, network_dir, iters):
    """examples/minimal_surface_of_revolution: Abs / sqrt in the integrand, L1 loss on the integral, NumPy-array targets
    on a two-point boundary, a constraint-free inference domain that must not contribute to the loss."""
    x = sp.Symbol("x")
    u = sp.Function("u")(x)
    geo = sc.Line1D(-1, 0.5)
    np.random.seed(31)

    @sc.datanode(sigma=1000.0)
    class Boundary(sc.SampleDomain):
        def __init__(self):
            self.points = geo.sample_boundary(1)
            self.constraints = {"u": np.cosh(self.points["x"])}

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    @sc.datanode(loss_fn="L1")
    class Interior(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return geo.sample_interior(200), {"integral_dx": 0}

    @sc.datanode
    class InteriorInfer(sc.SampleDomain):
        def __init__(self):
            self.points = sc.Variables()
            self.points["x"] = np.linspace(-1, 0.5, 11, endpoint=True).reshape(-1, 1)
            self.points["area"] = np.ones_like(self.points["x"])

        def sampling(self, *args, **kwargs):
            return self.points, {}

    domains = (Boundary(), Interior(), InteriorInfer())
    torch.manual_seed(31)
    dx_exp = sc.ExpressionNode(expression=sp.Abs(u) * sp.sqrt((u.diff(x)) ** 2 + 1), name="dx")
    net = sc.get_net_node(inputs=("x",), outputs=("u",), name="net", arch=sc.Arch.mlp)
    integral = sc.ICNode("dx", dim=1, time=False)
    return net, dict(sample_domains=domains, netnodes=[net], pdes=[dx_exp, integral], max_iter=iters,
                     network_dir=network_dir, loading=False)


def consolidation(sc, network_dir, iters):
    """examples/consolidation: per-point lambda arrays computed with NumPy, array-valued zero targets, a first-order
    time-like term next to a second derivative."""
    x, y = sp.symbols("x y")
    p = sp.Function("p")(x, y)
    geo = sc.Rectangle((0, 0), (1, 1))
    np.random.seed(37)

    @sc.datanode
    class Init(sc.SampleDomain):
        def __init__(self):
            self.points = geo.sample_boundary(30, sieve=(sp.Eq(y, 0.0)))
            self.constraints = {"p": 1.0, "lambda_p": 1 - np.power(self.points["x"], 5)}

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    @sc.datanode
    class Interior(sc.SampleDomain):
        def __init__(self):
            self.points = geo.sample_interior(150)
            self.constraints = {"consolidation": np.zeros_like(self.points["x"])}

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    @sc.datanode
    class UpperBounds(sc.SampleDomain):
        def __init__(self):
            self.points = geo.sample_boundary(30, sieve=(sp.Eq(x, 1.0)))
            self.constraints = {"p": 0.0, "lambda_p": np.power(self.points["y"], 0.2)}

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    @sc.datanode
    class LowerBounds(sc.SampleDomain):
        def __init__(self):
            self.points = geo.sample_boundary(30, sieve=(sp.Eq(x, 0.0)))
            self.constraints = {"normal_gradient_p": 0.0}

        def sampling(self, *args, **kwargs):
            return self.points, self.constraints

    domains = (Init(), Interior(), UpperBounds(), LowerBounds())
    torch.manual_seed(37)
    pde = sc.ExpressionNode(name="consolidation", expression=p.diff(y) - 0.6 * p.diff(x, 2))
    grad = sc.NormalGradient(T="p", dim=2, time=False)
    net = sc.get_net_node(inputs=("x", "y"), outputs=("p",), arch=sc.Arch.mlp, name="net")
    return net, dict(sample_domains=domains, netnodes=[net], pdes=[pde, grad], max_iter=iters,
                     network_dir=network_dir, loading=False)


def ns_steady(sc, network_dir, iters):
    """examples/ns_steady: rectangle minus circle (CSG interior and boundary sampling with sieves), a six-output net in
    the stress formulation, a symbolic inlet profile, two targets in one domain."""
    x, y = sp.Symbol("x"), sp.Symbol("y")
    rec = sc.Rectangle((0.0, 0.0), (1.1, 0.41))
    geo = rec - sc.Circle((0.2, 0.2), 0.05)
    u, v, p = (sp.Function(n)(x, y) for n in ("u", "v", "p"))
    s11, s22, s12 = (sp.Function(n)(x, y) for n in ("s11", "s22", "s12"))
    nu = 0.02

    @sc.datanode
    class Inlet(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return (rec.sample_boundary(60, sieve=(sp.Eq(x, 0.0))),
                    sc.Variables({"u": 4 * (0.41 - y) * y / (0.41 * 0.41)}))

    @sc.datanode
    class Outlet(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return geo.sample_boundary(60, sieve=(sp.Eq(x, 1.1))), sc.Variables({"p": 0.0})

    @sc.datanode
    class Wall(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return geo.sample_boundary(60, sieve=((x > 0.0) & (x < 1.1))), sc.Variables({"u": 0.0, "v": 0.0})

    @sc.datanode(name="NS_external")
    class Interior(sc.SampleDomain):
        def sampling(self, *args, **kwargs):
            return geo.sample_interior(300), {"f_s11": 0.0, "f_s22": 0.0, "f_s12": 0.0, "f_u": 0.0, "f_v": 0.0, "f_p": 0.0}

    np.random.seed(41)
    torch.manual_seed(41)
    net = sc.get_net_node(inputs=("x", "y"), outputs=("u", "v", "p", "s11", "s22", "s12"), name="net", arch=sc.Arch.mlp)
    pdes = [sc.ExpressionNode(name="f_s11", expression=-p + 2 * nu * u.diff(x) - s11),
            sc.ExpressionNode(name="f_s22", expression=-p + 2 * nu * v.diff(y) - s22),
            sc.ExpressionNode(name="f_s12", expression=nu * (u.diff(y) + v.diff(x)) - s12),
            sc.ExpressionNode(name="f_u", expression=u * u.diff(x) + v * u.diff(y) - nu * (s11.diff(x) + s12.diff(y))),
            sc.ExpressionNode(name="f_v", expression=u * v.diff(x) + v * v.diff(y) - nu * (s12.diff(x) + s22.diff(y))),
            sc.ExpressionNode(name="f_p", expression=p + (s11 + s22) / 2)]
    return net, dict(sample_domains=(Inlet(), Outlet(), Wall(), Interior()), netnodes=[net], pdes=pdes, max_iter=iters,
                     network_dir=network_dir, loading=False)


CASES = {"poisson": poisson, "burgers": burgers, "cavity": cavity, "beam_lbfgs": beam_lbfgs, "periodic": periodic,
         "inverse_wave": inverse_wave, "volterra": volterra, "deepritz": deepritz, "param_poisson": param_poisson,
         "min_surface": min_surface, "consolidation": consolidation, "ns_steady": ns_steady}

# after the 12 iterations: `infer_step` of these columns on freshly drawn points (seeded) of one domain
INFER_SEED = 55
INFER = {"poisson": ("heat_domain", ["x", "y", "T", "T__x", "T__y__y", "diffusion_T"]),
         "burgers": ("burgers_equation", ["x", "t", "u", "u__t", "u__x__x", "burgers_u"]),
         "ns_steady": ("NS_external", ["x", "y", "u", "s12__y", "f_u", "f_p"])}


def build(sc, network_dir, iters):  # (kept for callers of the first version)
    return poisson(sc, network_dir, iters)
