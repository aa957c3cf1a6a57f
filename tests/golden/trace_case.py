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


CASES = {"poisson": poisson, "burgers": burgers, "cavity": cavity}


def build(sc, network_dir, iters):  # (kept for callers of the first version)
    return poisson(sc, network_dir, iters)
