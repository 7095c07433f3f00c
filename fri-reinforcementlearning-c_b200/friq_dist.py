"""Agent sharding and result gathering for multi-GPU runs (one process per GPU).

The path shards by agent: agents never interact, so there is no data-path collective.
The only exchange is the gather of per-agent results (return f64, steps i32, rules i32 = 16 B/agent) and a
handful of reduced rule-base statistics - torch.distributed (NCCL over NVLink on GPUs, gloo in CPU tests).
Agent identity = global agent index = Philox key offset, so a sharded run equals one big run agent for agent.
"""
import torch


def shard(n_total, world, rank):
    """Contiguous block of agents owned by `rank`: (first_agent, count).  Blocks differ by at most one."""
    base, extra = divmod(int(n_total), int(world))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def gather_agent_results(dist, ret, steps, rules, out=None):
    """All-gather equal-sized per-agent results into agent order (rank-major == agent-major for contiguous
    shards).  `out` = optional preallocated (ret, steps, rules) of world * n elements."""
    world = dist.get_world_size()
    n = ret.numel()
    if out is None:
        out = (torch.empty(world * n, dtype=ret.dtype, device=ret.device),
               torch.empty(world * n, dtype=steps.dtype, device=steps.device),
               torch.empty(world * n, dtype=rules.dtype, device=rules.device))
    for dst, src in zip(out, (ret, steps, rules)):
        dist.all_gather_into_tensor(dst, src)
    return out


def reduce_stats(dist, ret, rules, stats=None):
    """Global rule-base statistics: [sum of rules, sum of returns, max rules] as float64."""
    if stats is None:
        stats = torch.zeros(3, dtype=torch.float64, device=ret.device)
    stats[0] = rules.sum(); stats[1] = ret.sum(); stats[2] = rules.max()
    dist.all_reduce(stats[:2])
    dist.all_reduce(stats[2:], op=dist.ReduceOp.MAX)
    return stats
