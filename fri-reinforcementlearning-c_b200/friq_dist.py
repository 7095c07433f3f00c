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


def gather_rulebases(dist, rant, rconc, cap):
    """Knowledge sharing across ranks (the reference's MPI mode sends numofrules, rant and rconc point to point,
    frirl_agent.c:427-461): every rank contributes ONE rule base - rant [R][K] float64, rconc [R] float64, R <= cap - and ONE
    all-gather of a fixed-size record (R, antecedents, consequents padded to cap) gives every rank all of them.  Returns
    [(rant_r, rconc_r) for r in range(world)]; each rank then folds them into its own agents in rank order with
    friq_batch_merge_rb, so all ranks apply the same deterministic sequence."""
    world = dist.get_world_size()
    R, K = rant.shape
    assert R <= cap and rconc.shape == (R,)
    rec = torch.zeros(1 + cap * (K + 1), dtype=torch.float64, device=rant.device)
    rec[0] = R
    rec[1:1 + R * K] = rant.reshape(-1)
    rec[1 + cap * K:1 + cap * K + R] = rconc
    allrec = torch.empty(world * rec.numel(), dtype=torch.float64, device=rant.device)
    dist.all_gather_into_tensor(allrec, rec)
    out = []
    for r in range(world):
        v = allrec[r * rec.numel():(r + 1) * rec.numel()]
        n = int(v[0].item())
        out.append((v[1:1 + n * K].reshape(n, K).clone(), v[1 + cap * K:1 + cap * K + n].clone()))
    return out


class ResultExchange:
    """The path's ONE exchange per step, as ONE collective: per-agent (return f64, steps i32, rules i32) live in a single
    16 B/agent device buffer (three SoA segments; `ret`, `steps`, `rules` are views the engine writes into with
    friq_batch_gather_device), which is all-gathered in one call; every rank then holds every agent's results, so the
    rule-base statistics need no further all-reduce.  Outputs are in global agent order (contiguous shards)."""

    def __init__(self, dist, n, device):
        self.dist, self.n = dist, n
        self.world = dist.get_world_size() if dist is not None else 1
        self.local = torch.empty(16 * n, dtype=torch.uint8, device=device)
        self.ret = self.local[:8 * n].view(torch.float64)
        self.steps = self.local[8 * n:12 * n].view(torch.int32)
        self.rules = self.local[12 * n:].view(torch.int32)
        self.all = torch.empty(self.world * 16 * n, dtype=torch.uint8, device=device) if self.world > 1 else self.local
        self.g_ret = torch.empty(self.world * n, dtype=torch.float64, device=device)
        self.g_steps = torch.empty(self.world * n, dtype=torch.int32, device=device)
        self.g_rules = torch.empty(self.world * n, dtype=torch.int32, device=device)
        self.stats = torch.zeros(3, dtype=torch.float64, device=device)

    def exchange(self):
        n, w = self.n, self.world
        if w > 1:
            self.dist.all_gather_into_tensor(self.all, self.local)
        blocks = self.all.view(w, 16 * n)
        self.g_ret.view(w, n).copy_(blocks[:, :8 * n].contiguous().view(torch.float64).view(w, n))
        self.g_steps.view(w, n).copy_(blocks[:, 8 * n:12 * n].contiguous().view(torch.int32).view(w, n))
        self.g_rules.view(w, n).copy_(blocks[:, 12 * n:].contiguous().view(torch.int32).view(w, n))
        self.stats[0] = self.g_rules.sum(); self.stats[1] = self.g_ret.sum(); self.stats[2] = self.g_rules.max()
        return self.g_ret, self.g_steps, self.g_rules, self.stats
