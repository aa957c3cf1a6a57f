"""Quick start: -Laplace(T) = 1 on a plate with a hole, everything on the GPU.

    python examples/poisson_on_device.py [iterations]

The only line an IDRLnet script changes is the import; the rest below uses two opt-in additions:
points drawn on the device (`sample_interior_device` / `sample_boundary_device`) and the top-k
residual query used for adaptive re-sampling."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a source checkout
import idrlnet_b200.shortcut as sc        # instead of: import idrlnet.shortcut as sc

plate = sc.Rectangle((-1.0, -1.0), (1.0, 1.0)) - sc.Circle((0.0, 0.0), 0.3)


@sc.datanode(name="interior")
class Interior(sc.SampleDomain):
    def __init__(self):
        self.step = 0

    def sampling(self, *args, **kwargs):
        self.step += 1                      # a new Philox offset -> fresh points every call, no host work
        return plate.sample_interior_device(20000, seed=0, offset=self.step << 20), {"diffusion_T": 1.0}


@sc.datanode(name="boundary")
class Boundary(sc.SampleDomain):
    def __init__(self):
        self.step = 0

    def sampling(self, *args, **kwargs):
        self.step += 1
        return plate.sample_boundary_device(2000, seed=1, offset=self.step << 20), {"T": 0.0}


def main():
    iterations = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    net = sc.get_net_node(inputs=("x", "y"), outputs=("T",), name="net", arch=sc.Arch.mlp)
    pde = sc.DiffusionNode(T="T", D=1.0, Q=0.0, dim=2, time=False)
    solver = sc.Solver(sample_domains=(Interior(), Boundary()), netnodes=[net], pdes=[pde], max_iter=iterations,
                       network_dir="./network_dir_poisson_on_device", loading=False)
    solver.solve()
    worst = solver.residual_top_k("interior", "diffusion_T", 10)     # where the PDE is violated most
    for x, y, r in zip(worst["x"].flatten().tolist(), worst["y"].flatten().tolist(), worst["diffusion_T"].flatten().tolist()):
        print(f"x = {x:+.3f}  y = {y:+.3f}  -Laplace(T) = {r:+.4f}  (target 1)")


if __name__ == "__main__":
    main()
