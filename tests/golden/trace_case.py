"""A seeded simple_poisson run shared by the reference (golden generator) and idrlnet_b200 (test)."""
import numpy as np
import sympy as sp
import torch


def build(sc, network_dir, iters):
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
