"""Worker of tests/test_multi_gpu.py::test_torchrun_nccl_result_exchange_and_rulebase_gather (one process per GPU, torchrun)."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "fri-reinforcementlearning-c_b200"))
sys.path.insert(0, HERE)
import friq_b200 as fq  # noqa: E402
import friq_dist  # noqa: E402
import _oracle  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, steps = 2048, 500
first, _ = friq_dist.shard(world * n, world, rank)
P = fq.Problem("mountaincar", no_random=0)
b = fq.Batch(P, n, 256, device=local, seed0=0x5EED0000, first_agent=first)
stream = torch.cuda.current_stream(); sptr = ctypes.c_void_p(stream.cuda_stream)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
x = friq_dist.ResultExchange(dist, n, "cuda",
                             stats_fn=lambda r, q, m, out: fq.result_stats(local, vp(r), vp(q), m, vp(out), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
for i in range(4):                                   # four steps through both buffer sets, exchanges overlapping the next step
    b.run(steps // 4, fq.RUN_LEARN | fq.RUN_RETILE, stream=sptr, sync=False)
    buf = x.next_buffers(stream)
    b.gather_device(vp(buf[0]), vp(buf[1]), vp(buf[2]), sptr)
    x.exchange_async(stream)
x.wait(stream); torch.cuda.synchronize()
g_ret, g_steps, g_rules, stats = [t.cpu().numpy() for t in x.results()]
# the same agents as ONE batch on this rank's GPU
whole = fq.Batch(P, world * n, 256, device=local, seed0=0x5EED0000, first_agent=0)
for i in range(4):
    whole.run(steps // 4, fq.RUN_LEARN | fq.RUN_RETILE)
o = whole.gather()
assert np.array_equal(g_rules, o["rules"]) and np.array_equal(g_ret, o["last_return"]) and np.array_equal(g_steps, o["last_steps"])
assert stats[0] == o["rules"].sum() and stats[2] == o["rules"].max()
# knowledge sharing across ranks: every rank contributes the rule base of its agent 5; all ranks merge all of them in rank order
# (FIVERB.weights kept exactly as the reference leaves them: merge_rb reads them stale)
bm = fq.Batch(P, 64, 256, device=local, seed0=0x5EED0000, first_agent=first)
bm.run(steps, fq.RUN_LEARN | fq.RUN_KEEP_WEIGHTS)
R, rant, rconc = bm.get_rulebase(5)
rbs = friq_dist.gather_rulebases(dist, torch.from_numpy(rant).cuda(), torch.from_numpy(rconc).cuda(), 256)
OP = _oracle.Problem(0, trig=1); OP.set_no_random(0)
a = _oracle.Agent(OP, 256, seed=0x5EED0000 + first + 9); a.run_steps(steps)
for r in range(world):
    if r != rank:
        bm.merge_rb(rbs[r][0].cpu().numpy(), rbs[r][1].cpu().numpy(), agent=9)
        a.merge_rb(rbs[r][0].cpu().numpy(), rbs[r][1].cpu().numpy())
R9, rant9, rconc9 = bm.get_rulebase(9); oR, orant, orconc = a.rb()
assert R9 == oR and np.array_equal(rant9, orant) and np.array_equal(rconc9, orconc)
dist.barrier()
print(f"NCCL_WORKER_OK rank {rank}", flush=True)
dist.destroy_process_group()
