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
    """The path's ONE exchange per step: per-agent (return f64, steps i32, rules i32), gathered from every rank STRAIGHT into
    agent-ordered arrays (contiguous shards: rank-major == agent-major, so `all_gather_into_tensor` needs no un-interleave
    copy), on a SIDE stream so that the collective overlaps the next step's kernels instead of serialising behind them.
    `depth` buffer sets rotate (default 4): the engine writes step i's results into set i % depth while the sets of the steps
    before it are still in flight.  The collective is the only thing that couples the ranks - a rank cannot run further ahead
    of the slowest one than the sets allow - and every rank's step time varies (each learns different rule bases): with two sets
    every step costs the slowest rank's time (8 GPUs: 0.968 of linear), with four the variations average out over the steps.

        buf = x.next_buffers()                  # (ret, steps, rules) views the engine writes into (friq_batch_gather_device)
        ... engine call on `stream` ...
        x.exchange_async(stream)                # side stream: wait for `stream`, 3 all-gathers as one grouped call (+ the statistics kernel)
        ...
        x.wait(stream)                          # `stream` waits for everything in flight; results() is then valid

    CUDA tensors: torch.distributed NCCL; CPU tensors (gloo, tests): the same calls run synchronously.  `stats_fn(ret, rules, n,
    out)` = optional device-side statistics kernel (friq_result_stats), else three torch reductions."""

    def __init__(self, dist, n, device, stats_fn=None, depth=4):
        self.dist, self.n, self.device = dist, n, torch.device(device)
        self.world = dist.get_world_size() if dist is not None else 1
        self.cuda = self.device.type == "cuda"
        self.stats_fn = stats_fn
        mk = lambda m, dt: torch.empty(m, dtype=dt, device=device)
        self.sets = []
        self.depth = max(2, int(depth))
        for _ in range(self.depth):
            loc = (mk(n, torch.float64), mk(n, torch.int32), mk(n, torch.int32))
            glo = tuple(mk(self.world * n, t.dtype) for t in loc) if self.world > 1 else loc
            self.sets.append(dict(local=loc, all=glo, stats=torch.zeros(3, dtype=torch.float64, device=device), done=None, used=False))
        self.cur = -1
        self._no_coalescing = False
        self.side = torch.cuda.Stream(device=self.device, priority=-1) if self.cuda else None
        # legacy single-set views (first set)
        self.ret, self.steps, self.rules = self.sets[0]["local"]

    def next_buffers(self, stream=None):
        """The buffer set of the coming step; `stream` (the engine's) first waits until the exchange that last used it is done."""
        self.cur = (self.cur + 1) % self.depth
        st = self.sets[self.cur]
        if self.cuda and st["done"] is not None:
            (stream or torch.cuda.current_stream(self.device)).wait_event(st["done"])
        return st["local"]

    def _run(self, st):
        if self.world > 1:
            # the three all-gathers as ONE grouped NCCL call (one kernel instead of three: each of them has to find a free SM next to
            # step kernels that fill the GPU, and holds it while it waits for the slowest rank); gloo and older torch: one by one
            cm = getattr(self.dist, "_coalescing_manager", None) if self.cuda else None
            done = False
            if cm is not None and not self._no_coalescing:
                try:
                    with cm(device=self.device):
                        for dst, src in zip(st["all"], st["local"]):
                            self.dist.all_gather_into_tensor(dst, src)
                    done = True
                except (RuntimeError, ValueError, TypeError):
                    self._no_coalescing = True
            if not done:
                for dst, src in zip(st["all"], st["local"]):
                    self.dist.all_gather_into_tensor(dst, src)
        g_ret, _, g_rules = st["all"]
        if self.stats_fn is not None:
            self.stats_fn(g_ret, g_rules, g_ret.numel(), st["stats"])
        else:
            st["stats"][0] = g_rules.sum(); st["stats"][1] = g_ret.sum(); st["stats"][2] = g_rules.max()

    def exchange_async(self, stream=None):
        if self.cur < 0:
            self.cur = 0
        st = self.sets[self.cur]; st["used"] = True
        if not self.cuda:
            self._run(st)
            return
        stream = stream or torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event(); ready.record(stream)
        self.side.wait_event(ready)
        with torch.cuda.stream(self.side):
            self._run(st)
            st["done"] = torch.cuda.Event(); st["done"].record(self.side)

    def wait(self, stream=None):
        if not self.cuda:
            return
        stream = stream or torch.cuda.current_stream(self.device)
        for st in self.sets:
            if st["done"] is not None:
                stream.wait_event(st["done"])

    def results(self):
        """(ret, steps, rules, stats) of the most recent exchange, in global agent order (call wait() first on CUDA)."""
        st = self.sets[max(self.cur, 0)]
        return st["all"] + (st["stats"],)

    def exchange(self):
        """Synchronous form: results written into `ret` / `steps` / `rules` (the first set) are exchanged and returned."""
        self.cur = 0
        self.exchange_async()
        self.wait()
        return self.results()
