"""N>1 host logic on the CPU with gloo (world_size 2): every domain is sharded
[rank::world], the gradients are summed with one all-reduce, and the result equals
the single-process step (the loss is a SUM over points, idrlnet/variable.py:60-65)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out, shared=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      IDRLNET_B200_LOGLEVEL="ERROR", IDRLNET_B200_DEVICE="cpu")
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sympy as sp
    import idrlnet_b200.shortcut as sc
    import idrlnet_b200.solver as solver_mod
    rng = np.random.default_rng(0)  # same points on every rank
    n = 101
    pts = {"x": rng.uniform(-1, 1, (n, 1)), "t": rng.uniform(0, 1, (n, 1)), "area": np.full((n, 1), 2.0 / n)}
    bc = {"x": np.where(rng.uniform(size=(33, 1)) < 0.5, -1.0, 1.0), "t": rng.uniform(0, 1, (33, 1)),
          "area": np.full((33, 1), 1.0 / 33)}
    d1 = sc.get_data_node(lambda: (dict(pts), {"burgers_u": 0.0}), name="interior")
    d2 = sc.get_data_node(lambda: (dict(bc), {"u": 0.0, "lambda_u": 3.0}), name="wall", sigma=2.0)
    torch.manual_seed(0)
    net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
    import tempfile
    nets, pdes, domains = [net], [sc.BurgersNode(u="u", v=0.01)], (d1, d2)
    if shared:
        # periodic boundary over a weight-sharing copy of the net on the shifted input (examples/allen_cahn/allen_cahn.py:94-118):
        # the same Parameter objects sit in two NetNodes and must be summed over the ranks ONCE
        x_, t_ = sp.Symbol("x"), sp.Symbol("t")
        nets.append(sc.get_shared_net_node(net, inputs=("xp", "t"), outputs=("up",), name="net2", arch="mlp"))
        pdes += [sc.ExpressionNode(name="xp", expression=x_ + 2),
                 sc.ExpressionNode(name="gap", expression=sp.Function("u")(x_, t_) - sp.Function("up")(x_, t_))]
        per = {"x": np.full((29, 1), -1.0), "t": rng.uniform(0, 1, (29, 1)), "area": np.full((29, 1), 1.0 / 29)}
        domains += (sc.get_data_node(lambda: (dict(per), {"gap": 0.0}), name="periodic"),)
    s = sc.Solver(sample_domains=domains, netnodes=nets, pdes=pdes, network_dir=tempfile.mkdtemp(), loading=False)
    params = list(net.net.parameters())
    # sharded step: each rank sees [rank::world] of every domain, gradients all-reduced
    s.optimizers[0].zero_grad()
    loss_sharded = float(s._loss_and_grads())
    g_sharded = [p.grad.clone() for p in params]
    comps = {k: float(v) for k, v in s.loss_component.items()}
    # single-process reference: same solver with the distributed layer switched off
    real = solver_mod._dist
    solver_mod._dist = lambda: None
    s.optimizers[0].zero_grad()
    loss_full = float(s._loss_and_grads())
    g_full = [p.grad.clone() for p in params]
    solver_mod._dist = real
    # loss components of the sharded run are per-rank partial sums on the generic path
    tot = torch.tensor([loss_sharded])
    dist.all_reduce(tot)
    err = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-12)) for a, b in zip(g_sharded, g_full))
    out.put((rank, err, float(tot), loss_full, comps))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("shared", [False, True], ids=["one_net", "shared_net"])
def test_two_ranks_equal_one(shared):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, shared)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, err, tot, full, comps in res:
        assert err < 1e-5, (rank, err)
        assert abs(tot - full) <= 1e-5 * abs(full), (tot, full)
