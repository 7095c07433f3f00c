#!/usr/bin/env python3
"""Small run of every fused-kernel policy, the shared-rule-base kernels, the any-action-set loop and the per-call kernels, meant to be
run under compute-sanitizer (`compute-sanitizer --tool memcheck python tools/sanitize_case.py`)."""
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
        print(name, policy, "rules", o["rules"].min(), o["rules"].max(), "steps", int(o["total_steps"].sum()), b.work_counters())
        b.close(); P.close()
# another action set: the plain fused loop
d = fq.ProblemDesc(); fq.lib().friq_problem_desc_builtin(0, d)
d.actiondim.values_len = 5; d.actiondim.values = None; d.actiondim.values_div = 0.5; d.no_random = 0
P = fq.Problem(desc=d); b = fq.Batch(P, 40, 192); b.run(150, fq.RUN_LEARN | fq.RUN_RETILE); print("anyA rules", b.gather()["rules"].max()); b.close(); P.close()
P = fq.Problem("mountaincar")
b = fq.Batch(P, 40, 192)
b.run(2000, fq.RUN_LEARN | fq.RUN_STOP_AT_EPISODE_END)
R, rant, rconc = b.get_rulebase(0)
r2, q2, eps = fq.reduce_rulebase(P, rant, rconc, width=8)
print("reduce", R, "->", len(q2), "episodes", eps)
for policy in (0, 2):                       # shared rule base: streaming pass / memoised conclusions (both start-state paths)
    b.set_policy(policy)
    b.set_shared_rulebase(rant, rconc)
    b.set_initial_state(np.random.default_rng(1).uniform([-1.2, -0.05], [0.3, 0.05], size=(40, 2))); b.run(50, 0)
    b.set_initial_state(np.tile([[-0.7, 0.01]], (40, 1))); b.new_episodes(); b.run(50, 0)
print("ok")
