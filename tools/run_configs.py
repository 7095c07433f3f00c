#!/usr/bin/env python3
"""tools/run_configs.py -- throughput of the BASELINE.json configurations that are NOT the bench.py headline (they are parity-test
cases first; this script only records how fast the same engine runs them on one GPU's share):
  C3a  acrobot, 65,536 learning agents            (config 3's shape on the acrobot problem)
  C3c  cartpole, 16,384 learning agents
  C4   acrobot, 262,144 / N agents against ONE shared 10,240-rule base, greedy inference, 200 steps
  C5   mountaincar hyper-parameter sweep, 1,048,576 / 8 = 131,072 agents per GPU (alpha x gamma x epsilon x seeds)
Prints one JSON line per configuration.  Timing: CUDA events through torch on the launching stream, after a warm-up."""
import ctypes, json, os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "fri-reinforcementlearning-c_b200"))
import friq_b200 as fq


def timed(fn, reps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3


def learning(name, n, cap, chunk, warm, reps, eps):
    P = fq.Problem(name, no_random=0)
    b = fq.Batch(P, n, cap, seed0=0x5EED0000)
    b.set_hyper(np.full(n, P.desc.alpha), np.full(n, P.desc.gamma), np.full(n, eps))
    sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    mode = fq.RUN_LEARN | fq.RUN_RETILE
    for _ in range(warm):
        b.run(chunk, mode, stream=sptr, sync=False)
    o0 = b.gather()
    t = timed(lambda: b.run(chunk, mode, stream=sptr, sync=False), reps)
    o1 = b.gather()
    steps = int((o1["total_steps"] - o0["total_steps"]).sum()); sumR = int((o1["sum_rules"] - o0["sum_rules"]).sum())
    assert steps == n * chunk * reps and not (o1["flags"] & 8).any(), "work skipped or capacity overflow"
    out = dict(config=f"{name} learning", agents=n, env_steps_per_s=steps / t, five_evals_per_s=steps / t * (P.A + 3.5),
               mean_rules=sumR / steps, ms_per_launch_group=t / reps * 1e3, env_steps_per_agent=chunk * (warm + reps))
    b.close(); P.close()
    return out


def c4(n, R, steps):
    rng = np.random.default_rng(1234)
    P = fq.Problem("acrobot")
    u = P.tables()["u"]
    idx = np.unique(rng.integers(0, P.U, size=(3 * R, P.K)), axis=0)
    idx = idx[rng.permutation(len(idx))[:R]]
    rant = np.stack([u[k][idx[:, k]] for k in range(P.K)], axis=1)
    rconc = np.random.default_rng(1235).uniform(-1000, 1000, R)
    b = fq.Batch(P, n, P.ncorner)
    b.set_shared_rulebase(rant, rconc)
    s0 = np.zeros((n, P.S)); s0[:, :2] = np.random.default_rng(4242).uniform(-np.pi / 2, np.pi / 2, size=(n, 2))
    b.set_initial_state(s0)
    sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    b.run(int(os.environ.get("C4_WARM_STEPS", 20)), 0, stream=sptr, sync=False)      # (a short first launch is what tools capture under ncu)
    t = timed(lambda: b.run(steps, 0, stream=sptr, sync=False), 1)
    out = dict(config="C4 acrobot, shared rule base, inference", agents=n, rules=R, env_steps_per_s=n * steps / t,
               five_evals_per_s=n * steps * P.A / t, rule_evals_per_s=n * steps * P.A * R / t, seconds=t)
    b.close(); P.close()
    return out


def c5(n, chunk, reps):
    P = fq.Problem("mountaincar", no_random=0)
    i = np.arange(n)
    alpha = 0.05 * (1 + i % 16); gamma = 0.90 + 0.00625 * (1 + (i // 16) % 16)
    eps = np.array([0, 0.001, 0.002, 0.005, 0.01, 0.02, 0.05, 0.1])[(i // 256) % 8]
    b = fq.Batch(P, n, 192, seed0=0x5EED0000)
    b.set_hyper(alpha, gamma, eps)
    sptr = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    mode = fq.RUN_LEARN | fq.RUN_RETILE
    for _ in range(4):
        b.run(chunk, mode, stream=sptr, sync=False)
    t = timed(lambda: b.run(chunk, mode, stream=sptr, sync=False), reps)
    o = b.gather()
    out = dict(config="C5 mountaincar sweep (alpha x gamma x eps x seeds), one GPU's share", agents=n, env_steps_per_s=n * chunk * reps / t,
               episodes_finished_mean=float(o["episodes"].mean()), rules_mean=float(o["rules"].mean()), overflow=int((o["flags"] & 8).any()))
    b.close(); P.close()
    return out


if __name__ == "__main__":
    which = sys.argv[1:] or ["C3a", "C3c", "C4", "C5"]
    if "C3a" in which:
        for n_ac in [int(x) for x in os.environ.get("C3A_AGENTS", "65536").split(",")]:
            print(json.dumps(learning("acrobot", n_ac, 704, 100, 5, 5, 0.001)), flush=True)
    if "C3c" in which:
        for n_cp in [int(x) for x in os.environ.get("C3C_AGENTS", "65536").split(",")]:    # 16,384 agents are only 128 CTAs: 20 of the 148 SMs idle
            print(json.dumps(learning("cartpole", n_cp, 2304, 50, 5, 5, 0.001)), flush=True)
    if "C4" in which: print(json.dumps(c4(int(os.environ.get("C4_AGENTS", 262144 // int(os.environ.get("C4_GPUS", 2)))), 10240, int(os.environ.get("C4_STEPS", 200)))), flush=True)
    if "C5" in which: print(json.dumps(c5(131072, 250, 6)), flush=True)
