"""CPU tests (gloo, world_size 2) of the N>1 host logic: agent sharding and the result gather that
bench.py runs over NCCL on the GPUs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "fri-reinforcementlearning-c_b200"))
import friq_dist  # noqa: E402


def test_shard_covers_all_agents_contiguously():
    for n, w in [(65536, 1), (65536, 8), (1048576, 8), (10, 4), (7, 8)]:
        nxt = 0
        for r in range(w):
            first, cnt = friq_dist.shard(n, w, r)
            assert first == nxt and cnt >= 0
            nxt += cnt
        assert nxt == n


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, cnt = friq_dist.shard(world * n, world, rank)
    ids = torch.arange(first, first + cnt)
    ret = ids.double() * -10.0; steps = (ids % 1000).int(); rules = (8 + ids % 173).int()
    g_ret, g_steps, g_rules = friq_dist.gather_agent_results(dist, ret, steps, rules)
    stats = friq_dist.reduce_stats(dist, ret, rules)
    x = friq_dist.ResultExchange(dist, cnt, "cpu")           # the one-collective form bench.py uses
    x.ret.copy_(ret); x.steps.copy_(steps); x.rules.copy_(rules)
    x_ret, x_steps, x_rules, x_stats = x.exchange()
    assert torch.equal(x_ret, g_ret) and torch.equal(x_steps, g_steps) and torch.equal(x_rules, g_rules) and torch.equal(x_stats, stats)
    if rank == 0:
        q.put((g_ret.numpy(), g_steps.numpy(), g_rules.numpy(), stats.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_is_in_agent_order_world2():
    world, n = 2, 1000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    g_ret, g_steps, g_rules, stats = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = np.arange(world * n)
    assert np.array_equal(g_ret, ids * -10.0) and np.array_equal(g_steps, ids % 1000) and np.array_equal(g_rules, 8 + ids % 173)
    assert stats[0] == (8 + ids % 173).sum() and stats[1] == (ids * -10.0).sum() and stats[2] == (8 + ids % 173).max()


# ---- knowledge sharing across ranks: one all-gather of rule bases, then the same merge order on every rank
sys.path.insert(0, os.path.join(REPO, "tests"))


def _learned_agent(seed):
    import _oracle
    P = _oracle.Problem(0, trig=1); P.set_no_random(0)
    a = _oracle.Agent(P, 256, seed=seed)
    a.set_hyper(0.5, 1.0, 0.05)
    a.run_steps(700)
    return P, a


def _merge_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, a = _learned_agent(1000 + rank)
    R, rant, rconc = a.rb()
    rbs = friq_dist.gather_rulebases(dist, torch.from_numpy(rant), torch.from_numpy(rconc), 256)
    for r in range(world):                                   # every rank folds the OTHER ranks' rules in, in rank order
        if r != rank:
            a.merge_rb(rbs[r][0].numpy(), rbs[r][1].numpy())
    q.put((rank,) + a.rb())
    dist.barrier()
    dist.destroy_process_group()


def test_rulebase_exchange_and_merge_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_merge_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, R, rant, rconc = q.get(timeout=180)
        got[rank] = (R, rant, rconc)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    agents = [_learned_agent(1000 + r) for r in range(world)]   # the same in one process
    rbs = [a.rb() for _, a in agents]
    for rank, (_, a) in enumerate(agents):
        for r in range(world):
            if r != rank:
                a.merge_rb(rbs[r][1], rbs[r][2])
        R, rant, rconc = a.rb()
        assert got[rank][0] == R and np.array_equal(got[rank][1], rant) and np.array_equal(got[rank][2], rconc)
