#!/usr/bin/env python3
"""Small run of every fused-kernel policy and the per-call / drop-in kernels, meant to be run under compute-sanitizer
(`compute-sanitizer --tool memcheck python tools/sanitize_case.py`)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "fri-reinforcementlearning-c_b200"))
import friq_b200 as fq

for name, cap, steps in (("mountaincar", 192, 260), ("acrobot", 704, 120), ("cartpole", 512, 40)):
    for policy in (0, 1, 2):
        if name == "cartpole" and policy:
            continue
        P = fq.Problem(name, no_random=0)
        n = 70
        b = fq.Batch(P, n, cap, seed0=3)
        b.set_policy(policy)
        b.set_hyper(np.full(n, 0.5), np.ones(n), np.full(n, 0.1))
        b.run(steps, fq.RUN_LEARN | fq.RUN_RETILE)
        b.run(37, fq.RUN_LEARN)
        o = b.gather()
        st = np.zeros((n, P.S)); hit, conc = b.vag_concl(np.zeros((n, P.K))); best, ac = b.best_action(st)
        print(name, policy, "rules", o["rules"].min(), o["rules"].max(), "steps", int(o["total_steps"].sum()))
        b.close(); P.close()
P = fq.Problem("mountaincar")
b = fq.Batch(P, 40, 192)
b.run(2000, fq.RUN_LEARN | fq.RUN_STOP_AT_EPISODE_END)
R, rant, rconc = b.get_rulebase(0)
r2, q2, eps = fq.reduce_rulebase(P, rant, rconc, width=8)
print("reduce", R, "->", len(q2), "episodes", eps)
b.set_shared_rulebase(rant, rconc); b.run(50, 0)
print("ok")
