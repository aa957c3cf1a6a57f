"""Throughput of the public Solver API (train_pipe) on the Burgers workload with device-resident fixed points,
beside bench.py's trainer loop: what a user of `Solver.solve()` gets, Python orchestration included."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import idrlnet_b200.shortcut as sc  # noqa: E402


def main():
    n_int = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)

    def u(lo, hi, n):
        return (torch.rand(n, 1, device=dev, generator=g) * (hi - lo) + lo)
    xi, ti = u(-1, 1, n_int), u(0, 1, n_int)
    xic = u(-1, 1, 10240)
    pts = {
        "interior": ({"x": xi, "t": ti, "area": torch.full((n_int, 1), 2.0 / n_int, device=dev)}, {"burgers_u": 0.0}),
        "ic": ({"x": xic, "t": torch.zeros_like(xic), "area": torch.full((10240, 1), 2.0 / 10240, device=dev)},
               {"u": -torch.sin(np.pi * xic)}),
        "bc": ({"x": torch.sign(u(-1, 1, 10240)), "t": u(0, 1, 10240), "area": torch.full((10240, 1), 2.0 / 10240, device=dev)},
               {"u": 0.0}),
    }
    doms = [sc.get_data_node((lambda p=p, c=c: (dict(p), dict(c))), name=k) for k, (p, c) in pts.items()]
    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    for post in (True, False):
        s = sc.Solver(sample_domains=tuple(doms), netnodes=[net], pdes=[sc.BurgersNode(u="u", v=0.01 / np.pi)],
                      network_dir="/tmp/bench_solver_api", loading=False, post_step_loss=post)
        for _ in range(5):
            s.train_pipe()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n = 50
        for _ in range(n):
            loss = s.train_pipe()
        float(loss)
        dt = (time.perf_counter() - t0) / n
        print(f"Solver.train_pipe post_step_loss={post}: {dt * 1e3:.3f} ms/step, {(n_int + 20480) / dt:.4g} points/s")
        if os.environ.get("PROFILE"):
            import cProfile
            import pstats
            pr = cProfile.Profile()
            pr.enable()
            for _ in range(20):
                s.train_pipe()
            torch.cuda.synchronize()
            pr.disable()
            pstats.Stats(pr).sort_stats("cumulative").print_stats(28)


if __name__ == "__main__":
    main()
