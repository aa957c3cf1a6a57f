"""Two GPUs of one box: pinn_step_finalize_allreduce (the deterministic reduction fused with the SUM over the
ranks' symmetric memory, include/pinn_b200.h) against pinn_step_finalize + NCCL all_reduce on the same shards --
bit-identical at world size 2 (a + b in either order), and against the single-process step.  Then a few hundred
steps with one rank delayed at random: the double-buffered slots / flags must never hand out a stale or torn sum.
(idrlnet has no multi-GPU path; the all-reduce is SURVEY.md 8(e)'s.)"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      IDRLNET_B200_LOGLEVEL="ERROR")
    sys.path.insert(0, ROOT)
    import time
    import tempfile
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    import idrlnet_b200 as pkg
    import idrlnet_b200.shortcut as sc
    import idrlnet_b200.solver as solver_mod
    pkg.use_gpu(rank)
    rng = np.random.default_rng(0)  # the same points on every rank; Solver shards them [rank::world]
    n, nb = 4099, 257
    pts = {"x": rng.uniform(-1, 1, (n, 1)), "t": rng.uniform(0, 1, (n, 1)), "area": np.full((n, 1), 2.0 / n)}
    bc = {"x": np.where(rng.uniform(size=(nb, 1)) < 0.5, -1.0, 1.0), "t": rng.uniform(0, 1, (nb, 1)),
          "area": np.full((nb, 1), 1.0 / nb)}
    d1 = sc.get_data_node(lambda: (dict(pts), {"burgers_u": 0.0}), name="interior")
    d2 = sc.get_data_node(lambda: (dict(bc), {"u": 0.0, "lambda_u": 3.0}), name="wall", sigma=2.0)

    def make(peer):
        os.environ["IDRLNET_B200_PEER_ALLREDUCE"] = "1" if peer else "0"
        torch.manual_seed(0)
        net = sc.get_net_node(inputs=("x", "t"), outputs=("u",), name="net1", arch=sc.Arch.mlp)
        s = sc.Solver(sample_domains=(d1, d2), netnodes=[net], pdes=[sc.BurgersNode(u="u", v=0.01)],
                      network_dir=tempfile.mkdtemp(), loading=False)
        return s, list(net.net.parameters())

    def step(s, params):
        s.optimizers[0].zero_grad()
        loss = float(s._loss_and_grads())
        return loss, torch.cat([p.grad.reshape(-1) for p in params]).clone()

    res = {"rank": rank}
    s_peer, p_peer = make(True)
    l_peer, g_peer = step(s_peer, p_peer)
    eng = next(iter(s_peer._engines.values()))
    res["peer_enabled"] = eng._peers is not None and bool(eng.reduced)
    s_nccl, p_nccl = make(False)
    l_nccl, g_nccl = step(s_nccl, p_nccl)
    eng2 = next(iter(s_nccl._engines.values()))
    res["nccl_is_nccl"] = eng2._peers is None and not eng2.reduced
    res["bit_identical"] = bool(torch.equal(g_peer, g_nccl)) and l_peer == l_nccl
    res["max_abs_diff"] = float((g_peer - g_nccl).abs().max())
    # the single-process step on all the points
    real = solver_mod._dist
    solver_mod._dist = lambda: None
    os.environ["IDRLNET_B200_PEER_ALLREDUCE"] = "0"
    s_one, p_one = make(False)
    l_one, g_one = step(s_one, p_one)
    solver_mod._dist = real
    res["err_vs_single"] = float((g_peer - g_one).abs().max() / g_one.abs().max())
    res["loss_vs_single"] = abs(l_peer - l_one) / abs(l_one)
    # stress: the same step again and again, rank (k % world) delayed now and then; every rank must see the same
    # gradient as the first time, every time
    bad = 0
    r = np.random.default_rng(rank + 1)
    for k in range(300):
        if r.uniform() < 0.1:
            time.sleep(float(r.uniform(0, 0.004)))
        _, g = step(s_peer, p_peer)
        if k % 10 == 9:
            bad += int(not torch.equal(g, g_peer))
    torch.cuda.synchronize()
    res["stale_or_torn"] = bad
    # a few optimizer steps through train_pipe: parameters stay identical across ranks
    s_peer.max_iter = s_peer.global_step + 5
    s_peer.solve()
    flat = torch.cat([p.detach().reshape(-1) for p in p_peer])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    res["params_identical"] = all(bool(torch.equal(b, both[0])) for b in both)
    out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(600)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs of one box")
def test_finalize_allreduce_over_peer_memory():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in res:
        print(r)
        assert r["peer_enabled"], r
        assert r["nccl_is_nccl"], r
        assert r["bit_identical"], r
        assert r["err_vs_single"] < 1e-5 and r["loss_vs_single"] < 1e-5, r
        assert r["stale_or_torn"] == 0, r
        assert r["params_identical"], r
