"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C-ABI,
against (a) the per-call vectors recorded from the compiled reference, (b) the reference's own
tests/orig golden rule bases, (c) the oracle port on seeded random-exploration runs.
Everything is compared bit-exactly (FP64 results as np.array_equal) - tighter than the 1e-12
relative bound of the north star."""
import numpy as np
import pytest

import _golden
import _oracle

pytestmark = pytest.mark.gpu
CAP = {"mountaincar": 192, "acrobot": 704, "cartpole": 512}


@pytest.fixture(scope="module")
def fq():
    import friq_b200
    friq_b200.lib()
    assert friq_b200.device_count() >= 1
    return friq_b200


@pytest.fixture(scope="module", params=_golden.ENVS)
def env(request, fq):
    name = request.param
    P = fq.Problem(name)
    yield name, _golden.load(name), P
    P.close()


def test_device_pow_x87_fast_and_exact_paths(fq):
    rng = np.random.default_rng(5)
    L = _oracle.lib()
    for p in (2, 3, 4, 5, 8):
        d = np.concatenate([rng.uniform(1e-3, 60.0, 200000), np.sqrt(rng.uniform(1e-6, 2500.0, 200000)),
                            2.0 ** rng.integers(-8, 8, 64), np.float64([1.0, 2.0, 0.5, 1.0 + 2 ** -52, 1.0 - 2 ** -53])])
        want = np.array([L.fo_pow_x87(float(x), p) for x in d[:60000]])
        got = fq.device_pow_x87(d, p)
        exact = fq.device_pow_x87(d, p, force_exact=True)
        assert np.array_equal(got, exact)              # fast path + filter == integer emulation, everywhere
        assert np.array_equal(got[:60000], want)       # == x87 long double chain (fast_pow.inl:17-27)


def test_branch_free_shepard_weight_equals_library_ops(fq):
    """1/pow_x87(sqrt(s)) through the branch-free sqrt / pow / rcp fast paths (with their fallback)
    must equal the composition of __dsqrt_rn, the integer x87 emulation and __drcp_rn, and the
    fallback must be rare on generic operands."""
    rng = np.random.default_rng(11)
    s = np.concatenate([rng.uniform(1e-6, 3000.0, 1500000), 10.0 ** rng.uniform(-12, 8, 500000),
                        np.float64([1.0, 4.0, 0.25, 2.0, 9.0, 1e-60, 1e60])])
    for p in (3, 5):
        lib_w = fq.device_pow_x87(s, p, force_exact=2)
        fast_w = fq.device_pow_x87(s, p, force_exact=3)
        assert np.array_equal(lib_w, fast_w)
        bad = fq.device_pow_x87(s[:2000000], p, force_exact=4)
        assert (bad != 0).mean() < 0.01


def test_device_trig_is_correctly_rounded(fq):
    rng = np.random.default_rng(6)
    x = np.concatenate([rng.uniform(-10, 10, 200000), rng.uniform(-4.5, 1.5, 100000)])
    s, c = fq.device_trig(x)
    P = _oracle.Problem(0, trig=1)
    L = _oracle.lib()
    idx = rng.integers(0, x.size, 30000)
    assert all(s[i] == L.fo_sin(P.h, float(x[i])) and c[i] == L.fo_cos(P.h, float(x[i])) for i in idx)


def test_get_best_action_vectors(fq, env):
    name, g, P = env
    recs = g["gba"]
    b = fq.Batch(P, len(recs), CAP[name])
    for i, r in enumerate(recs):
        b.set_rulebase(r["rant"], r["rconc"], agent=i)
    best, actconc = b.best_action(np.stack([r["states"] for r in recs]))
    for i, r in enumerate(recs):
        assert best[i] == r["best"]
        assert np.array_equal(actconc[i], r["actconc"])
    b.close()


def test_vag_concl_and_weight_vectors(fq, env):
    name, g, P = env
    recs = [r for r in g["sarsa"]]
    b = fq.Batch(P, len(recs), CAP[name])
    for i, r in enumerate(recs):
        b.set_rulebase(r["rant"], r["rconc"], agent=i)
    for j in range(2):                                  # qp and qnow are evaluated on the unmodified rule base
        hit, conc = b.vag_concl(np.stack([r["concl"][j]["ant"] for r in recs]))
        for i, r in enumerate(recs):
            assert hit[i] == r["concl"][j]["ret"] and conc[i] == r["concl"][j]["conc"]
    wrecs = [(i, r) for i, r in enumerate(recs) if r["weight"] is not None and r["R_after"] == r["R"]]
    ant = np.stack([r["weight"]["ant"] if r["weight"] is not None else r["q_ant"] for r in recs])
    hit, w = b.vag_concl_weight(ant)
    for i, r in wrecs:
        assert hit[i] == r["weight"]["ret"]
        if hit[i] < 0:
            assert np.array_equal(w[i, :r["R"]], r["weight"]["weights"])
    b.close()


def test_update_sarsa_vectors(fq, env):
    name, g, P = env
    recs = g["sarsa"]
    b = fq.Batch(P, len(recs), CAP[name])
    for i, r in enumerate(recs):
        b.set_rulebase(r["rant"], r["rconc"], agent=i)
    b.set_inserted(np.array([r["inserted_before"] for r in recs]))
    b.update_sarsa(np.stack([r["q_ant"] for r in recs]), np.array([r["reward"] for r in recs]),
                   np.stack([r["cur_q_ant"] for r in recs]))
    ins = b.get_inserted()
    for i, r in enumerate(recs):
        R, rant, rconc = b.get_rulebase(i)
        assert R == r["R_after"], (i, r["cat"])
        assert np.array_equal(rant, r["rant_after"])
        assert np.array_equal(rconc, r["rconc_after"]), (i, r["cat"])
        assert ins[i] == r["inserted_after"]
    b.close()


def _construct(fq, b, max_steps):
    """frirl_sequential_run's construct loop (frirl_sequential_run.c:57-161), one launch per episode."""
    log = []
    for _ in range(999):
        b.run(max_steps, fq.RUN_LEARN | fq.RUN_TRACK | fq.RUN_STOP_AT_EPISODE_END)
        o = b.gather()
        log.append((o["episodes"][0], o["last_steps"][0], o["last_return"][0], o["rules"][0]))
        if o["flags"][0] & fq.F_CONVERGED:
            break
    return np.array(log, dtype=np.float64), o


def test_construct_run_reproduces_reference_golden(fq, env):
    """The fused step kernel, stepping the example's dynamics on the device with no_random = 1, must
    reproduce the reference's whole learning trajectory: per-episode (steps, reward, rules) and the
    final rule base, byte-identical to tests/orig (tests/_test.sh:8-17)."""
    name, g, P = env
    n = 40                                              # >1 warp: every agent must agree (whole-batch self-check)
    b = fq.Batch(P, n, CAP[name])
    log, o = _construct(fq, b, 1000)
    ref = np.array(g["episodes"], dtype=np.float64)
    assert log.shape == ref.shape and np.array_equal(log, ref)
    assert np.all(o["episodes"] == o["episodes"][0]) and np.all(o["rules"] == o["rules"][0])
    assert not np.any(o["flags"] & fq.F_OVERFLOW)
    for agent in (0, 17, n - 1):
        R, rant, rconc = b.get_rulebase(agent)
        assert _golden.rb_text(rant, rconc) == _golden.golden_text(name)
    b.close()


@pytest.mark.parametrize("name", ["mountaincar", "acrobot", "cartpole"])
def test_random_exploration_matches_oracle_port(fq, name):
    """epsilon-greedy with the Philox stream: per-agent seeds, per-agent hyper-parameters, chunked
    launches (state must survive kernel boundaries); every agent equals the oracle port bit for bit."""
    P = fq.Problem(name, no_random=0)
    n, first = 24, 1000
    rng = np.random.default_rng(3)
    alpha = rng.uniform(0.2, 0.8, n); gamma = rng.uniform(0.9, 1.0, n); eps = rng.choice([0.0, 0.01, 0.05, 0.2], n)
    b = fq.Batch(P, n, CAP[name], seed0=0x5EED0000, first_agent=first)
    b.set_hyper(alpha, gamma, eps)
    steps = {"mountaincar": 1500, "acrobot": 700, "cartpole": 900}[name]
    for chunk, mode in ((1, fq.RUN_LEARN), (7, fq.RUN_LEARN | fq.RUN_RETILE), (steps - 508, fq.RUN_LEARN),
                        (250, fq.RUN_LEARN | fq.RUN_RETILE), (250, fq.RUN_LEARN | fq.RUN_RETILE)):
        b.run(chunk, mode)                              # re-tiling by rule count must not change any agent's results
    o = b.gather()
    OP = _oracle.Problem(_golden.ENV_ID[name], trig=1); OP.set_no_random(0)
    for i in range(n):
        a = _oracle.Agent(OP, CAP[name], seed=0x5EED0000 + first + i)
        a.set_hyper(alpha[i], gamma[i], eps[i])
        a.run_steps(steps)
        s = a.summary()
        R, rant, rconc = b.get_rulebase(i)
        oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(rant, orant) and np.array_equal(rconc, orconc), i
        assert (o["episodes"][i], o["last_steps"][i], o["last_return"][i], o["total_steps"][i], o["sum_rules"][i]) == \
            (s["episode_num"] - 1, s["last_steps"], s["last_return"], s["total_steps"], s["sum_rules"])
    b.close(); P.close()


@pytest.mark.parametrize("name,policy", [("mountaincar", 0), ("mountaincar", 1), ("mountaincar", 2), ("acrobot", 0),
                                         ("acrobot", 1), ("acrobot", 2), ("cartpole", 0)])
def test_every_evaluation_policy_matches_oracle(fq, name, policy):
    """The fused kernel has three evaluation policies (friq_batch_set_policy): 0 lock-step pass over all actions (+ hit sweep /
    compacted sums when few pairs lack an exact hit; cartpole's A = 21 pass runs in 7 chunks of 3 actions with the Q(s,a) chain
    in chunk 0 or served cooperatively), 1 exact hits first + compacted Shepard chains, 2 the same with memoised weights
    (mountaincar's and acrobot's default).  All of them must reproduce the oracle bit for bit - here on a hard mix: a partial last
    warp (70 agents), random OFF-GRID episode start states (no exact hit possible at episode start: every pair needs its
    Shepard sums), heavy exploration (divergent lanes, different rule counts per lane, partial Q(s,a) masks), per-agent
    hyper-parameters, chunked and re-tiled launches.  Reference: frirl_get_best_action.c:212-281 (the loop over actions),
    examples/cartpole/cartpole.c:114-168 (the off-grid quantiser)."""
    P = fq.Problem(name, no_random=0)
    n, first = 70, 7
    rng = np.random.default_rng(100 + policy)
    alpha = rng.uniform(0.2, 0.8, n); gamma = rng.uniform(0.9, 1.0, n); eps = rng.choice([0.0, 0.02, 0.1, 0.3], n)
    lo = {"mountaincar": [-1.2, -0.05], "acrobot": [-1.0, -1.0, -0.5, -0.5], "cartpole": [-1.0, -0.6, -0.12, -0.6]}[name]
    hi = {"mountaincar": [0.3, 0.05], "acrobot": [1.0, 1.0, 0.5, 0.5], "cartpole": [1.0, 0.6, 0.12, 0.6]}[name]
    s0 = rng.uniform(lo, hi, size=(n, P.S))
    b = fq.Batch(P, n, CAP[name], seed0=0x5EED0000, first_agent=first)
    b.set_policy(policy)
    b.set_hyper(alpha, gamma, eps)
    b.set_initial_state(s0)
    steps = {"mountaincar": 900, "acrobot": 400, "cartpole": 600}[name]
    for chunk, mode in ((3, fq.RUN_LEARN), (steps - 303, fq.RUN_LEARN | fq.RUN_RETILE), (300, fq.RUN_LEARN | fq.RUN_RETILE)):
        b.run(chunk, mode)
    o = b.gather()
    OP = _oracle.Problem(_golden.ENV_ID[name], trig=1); OP.set_no_random(0)
    for i in range(n):
        a = _oracle.Agent(OP, CAP[name], seed=0x5EED0000 + first + i)
        a.set_hyper(alpha[i], gamma[i], eps[i]); a.set_initial_state(s0[i])
        a.run_steps(steps)
        sm = a.summary()
        R, rant, rconc = b.get_rulebase(i)
        oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(rant, orant) and np.array_equal(rconc, orconc), i
        assert (o["episodes"][i], o["last_steps"][i], o["last_return"][i], o["total_steps"][i], o["sum_rules"][i]) == \
            (sm["episode_num"] - 1, sm["last_steps"], sm["last_return"], sm["total_steps"], sm["sum_rules"])
    b.close(); P.close()


@pytest.mark.parametrize("name,A,act_div", [("mountaincar", 5, 0.5), ("acrobot", 7, 1.0 / 3.0), ("cartpole", 5, 0.5)])
def test_fused_kernel_for_other_action_sets_matches_oracle(fq, name, A, act_div):
    """Problems on a built-in environment that are NOT one of the three example shapes - here another action set (the number of
    actions is a run-time value) - run through the plain fused loop k_run_steps_anyA (frirl_types.h:62-100: any frirl_desc):
    random exploration, per-agent hyper-parameters, off-grid start states, chunked launches; every agent equals the oracle."""
    d = fq.ProblemDesc()
    fq.lib().friq_problem_desc_builtin(fq.ENV_ID[name], d)
    d.actiondim.values_len = A; d.actiondim.values = None; d.actiondim.values_div = act_div
    d.no_random = 0
    P = fq.Problem(desc=d)
    assert P.A == A
    OP = _oracle.Problem(_golden.ENV_ID[name], trig=1, actions=A, act_div=act_div); OP.set_no_random(0)
    n, first = 40, 3
    rng = np.random.default_rng(77)
    alpha = rng.uniform(0.2, 0.8, n); gamma = rng.uniform(0.9, 1.0, n); eps = rng.choice([0.0, 0.02, 0.2], n)
    lo = {"mountaincar": [-1.2, -0.05], "acrobot": [-1.0, -1.0, -0.5, -0.5], "cartpole": [-1.0, -0.6, -0.12, -0.6]}[name]
    hi = {"mountaincar": [0.3, 0.05], "acrobot": [1.0, 1.0, 0.5, 0.5], "cartpole": [1.0, 0.6, 0.12, 0.6]}[name]
    s0 = rng.uniform(lo, hi, size=(n, P.S))
    b = fq.Batch(P, n, CAP[name], seed0=0x5EED0000, first_agent=first)
    b.set_hyper(alpha, gamma, eps); b.set_initial_state(s0)
    steps = {"mountaincar": 700, "acrobot": 300, "cartpole": 300}[name]
    for chunk, mode in ((2, fq.RUN_LEARN), (steps - 102, fq.RUN_LEARN | fq.RUN_RETILE), (100, fq.RUN_LEARN)):
        b.run(chunk, mode)
    o = b.gather()
    for i in range(n):
        a = _oracle.Agent(OP, CAP[name], seed=0x5EED0000 + first + i)
        a.set_hyper(alpha[i], gamma[i], eps[i]); a.set_initial_state(s0[i])
        a.run_steps(steps)
        sm = a.summary()
        R, rant, rconc = b.get_rulebase(i)
        oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(rant, orant) and np.array_equal(rconc, orconc), i
        assert (o["episodes"][i], o["last_steps"][i], o["last_return"][i], o["total_steps"][i], o["sum_rules"][i]) == \
            (sm["episode_num"] - 1, sm["last_steps"], sm["last_return"], sm["total_steps"], sm["sum_rules"])
    b.close(); P.close()


def test_hit_sweep_is_disabled_when_dictionary_ve_values_collide(fq):
    """The integer hit sweep is only valid while distinct dictionary entries have distinct VE values.  A loaded rule base
    whose antecedent value lies off the grid but snaps to the same universe point as a grid value creates such a
    collision (distance 0 between DIFFERENT dictionary indices): the engine must notice and fall back to the
    floating-point exact-hit test - results still equal the oracle."""
    name = "mountaincar"
    P = fq.Problem(name, no_random=0)
    g = _golden.load(name)
    rant, rconc = g["final"][1].copy(), g["final"][2].copy()
    u = P.tables()["u"]
    rant[5, 0] = rant[5, 0] + 0.2 * (u[0][1] - u[0][0])      # same nearest universe point, different value
    n = 40
    b = fq.Batch(P, n, CAP[name], seed0=77, first_agent=0)
    b.set_rulebase(rant, rconc)
    eps = np.full(n, 0.05); b.set_hyper(np.full(n, 0.5), np.ones(n), eps)
    b.run(500, fq.RUN_LEARN | fq.RUN_RETILE)
    OP = _oracle.Problem(_golden.ENV_ID[name], trig=1); OP.set_no_random(0)
    for i in (0, 13, 39):
        a = _oracle.Agent(OP, CAP[name], seed=77 + i)
        a.set_rb(rant, rconc); a.set_hyper(0.5, 1.0, 0.05)
        a.run_steps(500)
        R, brant, brconc = b.get_rulebase(i)
        oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(brant, orant) and np.array_equal(brconc, orconc), i
    b.close(); P.close()


@pytest.mark.parametrize("cap", [184, 187])
def test_table_mode_survives_dictionary_growth_and_odd_capacity(fq, cap):
    """Table mode of the fused kernel (memoised Shepard weights between dictionary cells, per-agent rule-cell lists and
    cell-to-rule maps).  A loaded rule base with an antecedent value OFF the grid whose universe point no grid value
    shares makes the dictionary - and with it the weight table and the per-agent indexes - grow after the batch
    exists.  Capacity 184: whole blocks of 8 rules are summed and, for Q(s,a), stored; 187, no multiple of the block size: table
    mode still, but the weights of Q(s,a) are looked up again instead of stored.  Results equal the oracle either way."""
    name = "mountaincar"
    P = fq.Problem(name, no_random=0)
    g = _golden.load(name)
    rant, rconc = g["final"][1].copy(), g["final"][2].copy()
    t = P.tables(); u = t["u"][0]
    used = {int(np.argmin(np.abs(u - v))) for v in np.unique(rant[:, 0])}
    used |= {int(np.argmin(np.abs(u - (-0.5))))}
    free = [i for i in range(3, len(u) - 3) if i not in used]
    rant[7, 0] = u[free[len(free) // 2]]                      # a new dictionary entry with its own VE value
    n = 45
    b = fq.Batch(P, n, cap, seed0=991, first_agent=5)
    b.run(40, fq.RUN_LEARN | fq.RUN_RETILE)                  # indexes sized for the original table ...
    b.reset()
    b.set_rulebase(rant, rconc)                              # ... which now grows
    b.set_hyper(np.full(n, 0.4), np.ones(n), np.full(n, 0.1))
    for chunk in (1, 130, 369):
        b.run(chunk, fq.RUN_LEARN | fq.RUN_RETILE)
    OP = _oracle.Problem(_golden.ENV_ID[name], trig=1); OP.set_no_random(0)
    for i in (0, 9, 31, 44):
        a = _oracle.Agent(OP, cap, seed=991 + 5 + i)
        a.set_rb(rant, rconc); a.set_hyper(0.4, 1.0, 0.1)
        a.run_steps(500)
        R, brant, brconc = b.get_rulebase(i)
        oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(brant, orant) and np.array_equal(brconc, orconc), i
    b.close(); P.close()


@pytest.mark.parametrize("name", ["mountaincar", "acrobot"])
def test_table_mode_equals_computed_chains(fq, name):
    """The looked-up weights ARE the computed ones: a larger batch than the oracle comparisons afford, run once in table
    mode (policy 2, the default) and with policy 1 (integer hit sweep + computed sqrt / x87 power / reciprocal chains); every
    rule base, return and step count must be identical."""
    n, steps = 4096 + 17, 1200 if name == "mountaincar" else 500
    rng = np.random.default_rng(8)
    alpha = rng.uniform(0.2, 0.8, n); gamma = rng.uniform(0.9, 1.0, n); eps = rng.choice([0.0, 0.01, 0.05, 0.2], n)
    res = []
    for policy in (2, 1):
        P = fq.Problem(name, no_random=0)
        b = fq.Batch(P, n, CAP[name], seed0=4242, first_agent=0)
        b.set_policy(policy)
        b.set_hyper(alpha, gamma, eps)
        for chunk in (steps // 3, steps - steps // 3):
            b.run(chunk, fq.RUN_LEARN | fq.RUN_RETILE)
        o = b.gather()
        rbs = [b.get_rulebase(i) for i in range(0, n, 97)]
        res.append((o, rbs))
        b.close(); P.close()
    for (o0, r0), (o1, r1) in zip(res[:-1], res[1:]):
        for k in ("last_return", "last_steps", "rules", "episodes", "flags", "total_steps", "sum_rules"):
            assert np.array_equal(o0[k], o1[k]), k
        for (Ra, ra, qa), (Rb, rb, qb) in zip(r0, r1):
            assert Ra == Rb and np.array_equal(ra, rb) and np.array_equal(qa, qb)


@pytest.mark.parametrize("name,W", [("mountaincar", 3), ("mountaincar", 4), ("acrobot", 3), ("cartpole", 3)])
def test_team_run_reproduces_reference_openmp_build(fq, name, W):
    """SURVEY 8f N3 on the GPU: knowledge sharing between agents (the reference's experimental frirl_omp_run + merge_rb,
    frirl_agent.c, built here with BUILD_OPENMP and run with OMP_NUM_THREADS = W: tests/golden/<env>_omp<W>.*).  Three
    teams of W agents side by side in one batch, every team must reproduce every agent's episode sequence (steps, return,
    rules) and the leader's final rule base byte for byte - including merge_rb's use of STALE FIVERB.weights after exact
    hits, which the fused kernel has to keep exactly as update_rules leaves them (FRIQ_RUN_KEEP_WEIGHTS)."""
    import os
    z = np.load(os.path.join(_golden.GOLDEN, f"{name}_omp{W}.npz"))["episodes"]
    want_txt = open(os.path.join(_golden.GOLDEN, f"{name}_omp{W}.frirlrb.txt")).read()
    P = fq.Problem(name)
    teams = 3
    b = fq.Batch(P, teams * W + 1, {"mountaincar": 256, "acrobot": 768, "cartpole": 512}[name])   # + 1: an incomplete trailing team is left alone
    log, rounds = b.team_run(W)
    assert rounds < 200
    for t in range(teams):
        for a in range(W):
            want = z[z[:, 0] == a][:, 1:]; got = log[log[:, 0] == t * W + a][:, 1:]
            assert _golden.same_episode_rows(want, got), (t, a, len(want), len(got))
        R, rant, rconc = b.get_rulebase(t * W)
        txt = "".join("".join("%.18f " % v for v in rant[i]) + "%.18f \n" % rconc[i] for i in range(R))
        assert txt == want_txt, t
    assert len(log[log[:, 0] == teams * W]) == 0
    b.close(); P.close()


def test_merge_rb_per_call_matches_oracle(fq, tmp_path):
    """merge_rb (frirl_agent.c:58-117) as a per-call entry point: agents that have learned with random exploration (their
    FIVERB.weights kept by FRIQ_RUN_KEEP_WEIGHTS) receive a sender's rule list - the golden rule base with shifted
    consequents, one antecedent moved OFF the grid (snapped before insertion, CHECK_STATES) - and must end with exactly the
    oracle's rule bases; a second merge sees the weights the first one left behind."""
    name = "mountaincar"
    P = fq.Problem(name, no_random=0)
    g = _golden.load(name)
    rant, rconc = g["final"][1].copy(), g["final"][2].copy()
    rng = np.random.default_rng(21)
    rconc = rconc + rng.choice([0.0, 0.5, -3.0, 7.0, -40.0], len(rconc))
    u = P.tables()["u"][0]
    rant[11, 0] = u[17] + 0.25 * (u[18] - u[17])
    n, first, steps = 40, 300, 600
    alpha = rng.uniform(0.2, 0.8, n); gamma = rng.uniform(0.9, 1.0, n); eps = rng.choice([0.0, 0.02, 0.1], n)
    b = fq.Batch(P, n, 256, seed0=0x5EED0000, first_agent=first)
    b.set_hyper(alpha, gamma, eps)
    b.run(steps, fq.RUN_LEARN | fq.RUN_KEEP_WEIGHTS)
    ck = str(tmp_path / "team.ckp")
    b.save_checkpoint(ck)                                    # rule bases + inserted flags + FIVERB.weights
    b.merge_rb(rant, rconc)
    b.merge_rb(rant[::-1][:30], rconc[::-1][:30] * 0.5, agent=7)
    b2 = fq.Batch(P, n, 256, seed0=1, first_agent=0)         # a resumed batch must merge to the same rule bases
    b2.load_checkpoint(ck)
    b2.merge_rb(rant, rconc)
    for i in (0, 7, 22, 39):
        Ra, ra, qa = b.get_rulebase(i) if i != 7 else (None, None, None)
        Rb, rb_, qb = b2.get_rulebase(i)
        if i != 7:
            assert Ra == Rb and np.array_equal(ra, rb_) and np.array_equal(qa, qb), i
    b2.close()
    # a batch with ANOTHER rule capacity re-strides the saved FIVERB.weights (the header carries the capacity at save time);
    # a batch too small for a saved rule base and a file with trailing bytes are rejected
    b3 = fq.Batch(P, n, 301, seed0=1, first_agent=0)
    b3.load_checkpoint(ck)
    b3.merge_rb(rant, rconc)
    for i in (0, 22, 39):
        Ra, ra, qa = b.get_rulebase(i); Rb, rb_, qb = b3.get_rulebase(i)
        assert Ra == Rb and np.array_equal(ra, rb_) and np.array_equal(qa, qb), i
    b3.close()
    b4 = fq.Batch(P, n, 12, seed0=1, first_agent=0)
    with pytest.raises(fq.FriqError):
        b4.load_checkpoint(ck)
    b4.close()
    with open(ck, "ab") as f:
        f.write(b"\0" * 8)
    b5 = fq.Batch(P, n, 256, seed0=1, first_agent=0)
    with pytest.raises(fq.FriqError):
        b5.load_checkpoint(ck)
    b5.close()
    OP = _oracle.Problem(_golden.ENV_ID[name], trig=1); OP.set_no_random(0)
    for i in range(n):
        a = _oracle.Agent(OP, 256, seed=0x5EED0000 + first + i)
        a.set_hyper(alpha[i], gamma[i], eps[i])
        a.run_steps(steps)
        a.merge_rb(rant, rconc)
        if i == 7:
            a.merge_rb(rant[::-1][:30], rconc[::-1][:30] * 0.5)
        R, brant, brconc = b.get_rulebase(i)
        oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(brant, orant) and np.array_equal(brconc, orconc), i
    b.run(200, fq.RUN_LEARN | fq.RUN_KEEP_WEIGHTS)           # and learning goes on from the merged rule bases
    for i in (0, 7, 39):
        a = _oracle.Agent(OP, 256, seed=0x5EED0000 + first + i)
        a.set_hyper(alpha[i], gamma[i], eps[i])
        a.run_steps(steps); a.merge_rb(rant, rconc)
        if i == 7:
            a.merge_rb(rant[::-1][:30], rconc[::-1][:30] * 0.5)
        a.run_steps(200)
        R, brant, brconc = b.get_rulebase(i)
        oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(brant, orant) and np.array_equal(brconc, orconc), i
    b.close(); P.close()


def test_retiled_results_are_reported_in_agent_order(fq):
    """FRIQ_RUN_RETILE moves agents between warps (sorted by rule count); device-side gathers and the
    host-buffer end-to-end call must still report per-agent results in agent order."""
    import ctypes
    import torch
    P = fq.Problem("mountaincar", no_random=0)
    n = 200
    ref = fq.Batch(P, n, 192); ref.run(1200, fq.RUN_LEARN); want = ref.gather()
    b = fq.Batch(P, n, 192)
    for _ in range(4):
        b.run(300, fq.RUN_LEARN | fq.RUN_RETILE)
    d_ret = torch.empty(n, dtype=torch.float64, device="cuda"); d_steps = torch.empty(n, dtype=torch.int32, device="cuda")
    d_rules = torch.empty(n, dtype=torch.int32, device="cuda")
    b.gather_device(ctypes.c_void_p(d_ret.data_ptr()), ctypes.c_void_p(d_steps.data_ptr()), ctypes.c_void_p(d_rules.data_ptr()))
    b.sync()
    assert np.array_equal(d_rules.cpu().numpy(), want["rules"]) and np.array_equal(d_ret.cpu().numpy(), want["last_return"])
    assert len(set(want["rules"])) > 3                  # the sort key really varies
    got = b.gather()
    for k in want:
        assert np.array_equal(got[k], want[k]), k
    for agent in (0, 57, n - 1):
        assert all(np.array_equal(x, y) for x, y in zip(b.get_rulebase(agent), ref.get_rulebase(agent)))
    # end-to-end call with host buffers after further re-tiled steps
    ref.run(100, fq.RUN_LEARN); want = ref.gather()
    h_ret = np.zeros(n); h_steps = np.zeros(n, np.int32); h_rules = np.zeros(n, np.int32)
    eps = np.full(n, 0.01); alpha = np.full(n, 0.5); gamma = np.ones(n)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    b.run_host(100, fq.RUN_LEARN | fq.RUN_RETILE, vp(alpha), vp(gamma), vp(eps), vp(h_ret), vp(h_steps), vp(h_rules))
    assert np.array_equal(h_rules, want["rules"]) and np.array_equal(h_ret, want["last_return"]) and np.array_equal(h_steps, want["last_steps"])
    b.close(); ref.close(); P.close()


def test_inference_mode_leaves_rule_base_untouched(fq):
    """reduction_state = 1 semantics (frirl_test_run.c:65-69): greedy episode, no rule-base writes."""
    g = _golden.load("mountaincar")
    P = fq.Problem("mountaincar")
    b = fq.Batch(P, 33, 192)
    b.set_rulebase(g["final"][1], g["final"][2])          # all agents get the learned rule base
    b.run(1000, fq.RUN_STOP_AT_EPISODE_END)               # no RUN_LEARN
    o = b.gather()
    last = g["episodes"][-1]
    assert np.all(o["last_steps"] == last[1]) and np.all(o["last_return"] == last[2])
    R, rant, rconc = b.get_rulebase(32)
    assert np.array_equal(rant, g["final"][1]) and np.array_equal(rconc, g["final"][2])
    b.close(); P.close()


@pytest.mark.parametrize("policy", [0, 2])
@pytest.mark.parametrize("name,R,n,steps", [("mountaincar", 0, 200, 400), ("acrobot", 700, 150, 120), ("acrobot", 513, 40, 30), ("acrobot", 300, 330, 25)])
def test_shared_rule_base_inference_matches_oracle(fq, name, R, n, steps, policy):
    """BASELINE config 4 in miniature: many agents, ONE read-only rule base, randomised start states, greedy stepping.
    policy 0: one full pass over the rule base per env step (streamed through shared memory by TMA); policy 2 (default):
    conclusions memoised per (observation grid cell, action) + one pass per agent at its off-grid start state.
    R = 0: the learned golden rule base (mountaincar episodes END and restart inside the run); else a synthetic rule base of R
    distinct rules on universe points.  Final raw states, actions and episode results of every agent must
    equal the oracle port's bit for bit."""
    rng = np.random.default_rng(1234)
    P = fq.Problem(name)
    OP = _oracle.Problem(_golden.ENV_ID[name], trig=1)
    g = _golden.load(name)
    if R == 0:
        rant, rconc = g["final"][1], g["final"][2]
    else:
        u = P.tables()["u"]
        idx = np.unique(rng.integers(0, P.U, size=(4 * R, P.K)), axis=0)
        idx = idx[rng.permutation(len(idx))[:R]]
        rant = np.concatenate([g["header"]["rb0"][1], np.stack([u[k][idx[:, k]] for k in range(P.K)], axis=1)])[:R]
        rconc = np.concatenate([np.zeros(len(g["header"]["rb0"][2])), rng.uniform(-1000, 1000, R)])[:R]
    lo = {"mountaincar": [-1.2, -0.05], "acrobot": [-np.pi / 2, -np.pi / 2, 0, 0]}[name]
    hi = {"mountaincar": [0.3, 0.05], "acrobot": [np.pi / 2, np.pi / 2, 0, 0]}[name]
    s0 = rng.uniform(lo, hi, size=(n, P.S))
    if n == 330:                                        # few DISTINCT start states: the de-duplicated start pass (one warp per distinct
        s0 = s0[rng.integers(0, 9, n)]                  # universe-index tuple) instead of one streaming pass per agent
    b = fq.Batch(P, n, max(len(rconc), P.ncorner))
    b.set_policy(policy)
    b.set_shared_rulebase(rant, rconc)
    b.set_initial_state(s0)
    with pytest.raises(fq.FriqError):
        b.run(1, fq.RUN_LEARN)                          # a shared rule base is read-only
    b.run(steps // 3, 0); b.run(steps - steps // 3, 0)
    o = b.gather(); s, a = b.get_state()
    for i in range(n):
        ag = _oracle.Agent(OP, len(rconc) + 8); ag.set_rb(rant, rconc); ag.set_learning(0); ag.set_initial_state(s0[i])
        ag.run_steps(steps)
        sm = ag.summary()
        assert np.array_equal(s[i], sm["s"]), i
        assert (o["total_steps"][i], o["sum_rules"][i]) == (sm["total_steps"], sm["sum_rules"])
        assert (o["episodes"][i], o["last_steps"][i], o["last_return"][i]) == (sm["episode_num"] - 1, sm["last_steps"], sm["last_return"])
    b.close(); P.close()


def test_capacity_overflow_is_flagged(fq):
    P = fq.Problem("mountaincar")
    b = fq.Batch(P, 32, 12)                               # 8 corner rules + room for 4
    b.run(400, fq.RUN_LEARN)
    o = b.gather()
    assert np.all(o["rules"] == 12) and np.all(o["flags"] & fq.F_OVERFLOW)
    b.close(); P.close()


def test_acrobot_table_mode_with_a_capacity_off_the_block_size_matches_oracle(fq):
    """Acrobot in table mode with a rule capacity (701) that is no multiple of the walk's block of 8 rules: many lanes of a warp
    update with weights at every step, and with such a capacity those weights are looked up again instead of being stored in
    whole blocks.  Sampled agents equal the oracle."""
    name, n, cap, steps = "acrobot", 96, 701, 350
    P = fq.Problem(name, no_random=0)
    rng = np.random.default_rng(12)
    alpha = rng.uniform(0.2, 0.8, n); eps = rng.choice([0.0, 0.02, 0.1], n)
    b = fq.Batch(P, n, cap, seed0=31337)
    b.set_hyper(alpha, np.ones(n), eps)
    for chunk in (100, steps - 100):
        b.run(chunk, fq.RUN_LEARN | fq.RUN_RETILE)
    OP = _oracle.Problem(_golden.ENV_ID[name], trig=1); OP.set_no_random(0)
    for i in (0, 17, 40, 95):
        a = _oracle.Agent(OP, cap, seed=31337 + i)
        a.set_hyper(alpha[i], 1.0, eps[i]); a.run_steps(steps)
        R, rant, rconc = b.get_rulebase(i)
        oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(rant, orant) and np.array_equal(rconc, orconc), i
    b.close(); P.close()


def test_full_rule_bases_in_table_mode_equal_computed_chains(fq):
    """Rule bases that fill their capacity to the last slot (capacity 40 = five blocks of 8: the table-mode walk then sums a
    last block without a zero-weight tail and reads the zero cells kept BEHIND the column), with exploration, re-tiling and
    weighted updates going on after the overflow: table mode (policy 2) and the computed chains (policy 1) must agree bit for bit."""
    n = 200
    rng = np.random.default_rng(5)
    alpha = rng.uniform(0.2, 0.8, n); eps = rng.choice([0.0, 0.05, 0.2], n)
    res = []
    for policy in (2, 1):
        P = fq.Problem("mountaincar", no_random=0)
        b = fq.Batch(P, n, 40, seed0=77)
        b.set_policy(policy)
        b.set_hyper(alpha, np.ones(n), eps)
        for chunk in (700, 900):
            b.run(chunk, fq.RUN_LEARN | fq.RUN_RETILE)
        o = b.gather()
        res.append((o, [b.get_rulebase(i) for i in range(0, n, 7)]))
        b.close(); P.close()
    (o0, r0), (o1, r1) = res
    full = o0["rules"] == 40
    assert full.sum() >= 10 and np.all(o0["flags"][full] & fq.F_OVERFLOW)
    for k in ("last_return", "last_steps", "rules", "episodes", "flags", "total_steps", "sum_rules"):
        assert np.array_equal(o0[k], o1[k]), k
    for (Ra, ra, qa), (Rb, rb, qb) in zip(r0, r1):
        assert Ra == Rb and np.array_equal(ra, rb) and np.array_equal(qa, qb)


@pytest.mark.parametrize("width", [1, 7, 64])
def test_batched_rule_base_reduction_equals_reference(fq, width, capfd):
    """Rule-base reduction (frirl_sequential_run.c:170-350) with `width` candidate removals tested speculatively per launch must
    give exactly what the reference's one-candidate-per-episode loop gives: the shipped mountaincar run (`-f <constructed rule
    base>`) reduces 110 rules to 28; the reduced rule base and the log lines are golden fixtures recorded from the reference."""
    import os
    g = _golden.load("mountaincar")
    P = fq.Problem("mountaincar")
    capfd.readouterr()
    rant, rconc, episodes = fq.reduce_rulebase(P, g["final"][1], g["final"][2], strategy=1, width=width, verbose=1)
    out = capfd.readouterr().out
    want = open(os.path.join(_golden.GOLDEN, "mountaincar_reduced.frirlrb.txt"), "rb").read()
    assert _golden.rb_text(rant, rconc) == want
    log = [l for l in out.splitlines() if l.startswith(("Reduction Episode", "Smallest"))]
    ref = [l for l in open(os.path.join(_golden.GOLDEN, "mountaincar_reduction.log")).read().splitlines() if not l.startswith("Loaded")]
    assert log == ref
    assert episodes == 2 + 110                          # the reference replays 1 + 111 episodes; speculation never adds accepted work
    P.close()


def test_batched_rule_base_io_and_checkpoint_resume(fq, tmp_path):
    """Per-agent dumps in the reference's text / binary formats, `-f` warm start for a batch, and a one-file checkpoint: a run
    that is checkpointed, destroyed and resumed in a NEW batch must continue bit-identically to the uninterrupted run (rule
    bases and the sticky fus_is_rule_inserted flag are the learner's whole persistent state between episodes)."""
    import os
    g = _golden.load("mountaincar")
    P = fq.Problem("mountaincar", no_random=0)
    n = 70
    a = fq.Batch(P, n, 192, seed0=99)
    a.run(700, fq.RUN_LEARN | fq.RUN_RETILE | fq.RUN_STOP_AT_EPISODE_END)   # everyone stops at an episode boundary
    a.save_checkpoint(str(tmp_path / "ck.bin"))
    a.save_rb_files(str(tmp_path / "agent%04ld.frirlrb.bin"), first=60, count=10)
    a.save_rb_files(str(tmp_path / "agent%04ld.frirlrb.txt"), first=60, count=10, text=True)
    for i in (60, 69):
        R, rant, rconc = a.get_rulebase(i)
        frant, frconc = fq.load_rb_bin(str(tmp_path / f"agent{i:04d}.frirlrb.bin"), P.K)
        assert np.array_equal(frant, rant) and np.array_equal(frconc, rconc)
        assert open(tmp_path / f"agent{i:04d}.frirlrb.txt", "rb").read() == _golden.rb_text(rant, rconc)
    ep = a.gather()["episodes"]
    # resume in a fresh batch: same seeds; the Philox counter is (episode, step), so the episode numbers must line up -
    # a fresh batch starts at episode 1, hence compare against a reference batch that is ALSO restarted from the checkpoint
    b = fq.Batch(P, n, 192, seed0=99); b.load_checkpoint(str(tmp_path / "ck.bin"))
    c = fq.Batch(P, n, 192, seed0=99)
    for i in range(n):
        R, rant, rconc = a.get_rulebase(i); c.set_rulebase(rant, rconc, agent=i)
    c.set_inserted(a.get_inserted())
    b.run(400, fq.RUN_LEARN); c.run(400, fq.RUN_LEARN)
    for i in (0, 31, 32, 69):
        assert all(np.array_equal(x, y) for x, y in zip(b.get_rulebase(i), c.get_rulebase(i))), i
    assert np.array_equal(b.gather()["last_return"], c.gather()["last_return"])
    # warm start of every agent from the reference's own golden file == set_rulebase of its contents
    d = fq.Batch(P, 33, 192, seed0=5); d.load_rb_file(os.path.join(_golden.GOLDEN, "mountaincar.frirlrb.bin"))
    R, rant, rconc = d.get_rulebase(32)
    assert np.array_equal(rant, g["final"][1]) and np.array_equal(rconc, g["final"][2]) and not d.get_inserted().any()
    for x in (a, b, c, d):
        x.close()
    P.close()


def test_full_size_batch_properties(fq):
    """BASELINE config 3 at its full size (65,536 mountaincar agents on one GPU), checked through size-independent properties:
    (1) with no_random = 1 every one of the 65,536 agents must walk the reference's own trajectory (SURVEY 8d, C3's self-check):
        same rule count, return and steps as the oracle's single agent, and a sampled agent's rule base bit for bit;
    (2) sharding invariance - the batch cut into two shards with first_agent offsets (what two GPUs would run) gives every
        agent exactly the result it gets in the whole batch, with exploration on and re-tiling between launches."""
    n, steps = 65536, 900
    P = fq.Problem("mountaincar")                                   # no_random = 1
    b = fq.Batch(P, n, 192)
    for _ in range(3):
        b.run(steps // 3, fq.RUN_LEARN | fq.RUN_RETILE)
    o = b.gather()
    a = _oracle.Agent(_oracle.Problem(0, trig=1), 192); a.run_steps(steps); sm = a.summary()
    assert np.all(o["rules"] == sm["R"]) and np.all(o["total_steps"] == steps)
    assert np.all(o["last_return"] == sm["last_return"]) and np.all(o["last_steps"] == sm["last_steps"]) and np.all(o["episodes"] == sm["episode_num"] - 1)
    for agent in (0, 31337, n - 1):
        R, rant, rconc = b.get_rulebase(agent); oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(rant, orant) and np.array_equal(rconc, orconc)
    b.close(); P.close()
    P = fq.Problem("mountaincar", no_random=0)
    whole = fq.Batch(P, n, 192, seed0=0x5EED0000, first_agent=0)
    lo = fq.Batch(P, n // 2, 192, seed0=0x5EED0000, first_agent=0)
    hi = fq.Batch(P, n // 2, 192, seed0=0x5EED0000, first_agent=n // 2)
    for x in (whole, lo, hi):
        for _ in range(2):
            x.run(300, fq.RUN_LEARN | fq.RUN_RETILE)
    w, l, h = whole.gather(), lo.gather(), hi.gather()
    for k in ("rules", "last_return", "last_steps", "episodes", "sum_rules"):
        assert np.array_equal(w[k], np.concatenate([l[k], h[k]])), k
    assert len(np.unique(w["rules"])) > 10                          # exploration really diversifies the agents
    for agent in (5, n // 2 + 77):
        src = lo if agent < n // 2 else hi
        assert all(np.array_equal(x, y) for x, y in zip(whole.get_rulebase(agent), src.get_rulebase(agent % (n // 2))))
    for x in (whole, lo, hi):
        x.close()
    P.close()


def test_bench_workload_matches_oracle_on_sampled_agents(fq):
    """The bench workload itself (bench.py, BASELINE configs[2]): 65,536 mountaincar agents, eps = 0.01 (Philox), learning from the
    corner rule base for 6,250 env steps (25 bench steps x 250), re-tiled every 125 steps in two stream groups - exactly the
    driver's `--steps 20 --warmup 5` run.  48 randomly chosen agents (plus the first and the last) must end with the oracle
    port's rule base, return, step and rule-sum counters bit for bit."""
    n, chunk, nchunks = 65536, 250, 25
    P = fq.Problem("mountaincar", no_random=0)
    b = fq.Batch(P, n, 192, seed0=0x5EED0000, first_agent=0)
    b.set_retile_interval(125)
    for _ in range(nchunks):
        b.run(chunk, fq.RUN_LEARN | fq.RUN_RETILE)
    o = b.gather()
    assert int(o["total_steps"].sum()) == n * chunk * nchunks
    assert not np.any(o["flags"] & fq.F_OVERFLOW)
    rng = np.random.default_rng(20261019)
    picks = sorted(set([0, n - 1] + [int(x) for x in rng.integers(0, n, 48)]))
    OP = _oracle.Problem(0, trig=1); OP.set_no_random(0)
    for i in picks:
        a = _oracle.Agent(OP, 192, seed=0x5EED0000 + i)
        a.run_steps(chunk * nchunks)
        sm = a.summary()
        R, rant, rconc = b.get_rulebase(i)
        oR, orant, orconc = a.rb()
        assert R == oR and np.array_equal(rant, orant) and np.array_equal(rconc, orconc), i
        assert (o["episodes"][i], o["last_steps"][i], o["last_return"][i], o["total_steps"][i], o["sum_rules"][i]) == \
            (sm["episode_num"] - 1, sm["last_steps"], sm["last_return"], sm["total_steps"], sm["sum_rules"]), i
    b.close(); P.close()


def test_shared_rule_base_at_config4_size_matches_oracle_samples(fq):
    """BASELINE config 4 at one GPU's share of its full size (131,072 acrobot agents against ONE 10,240-rule base, greedy
    inference): sampled agents must end in exactly the state the oracle port reaches, and the batch must be deterministic."""
    rng = np.random.default_rng(1234)
    P = fq.Problem("acrobot")
    n, R, steps = 131072, 10240, 40
    u = P.tables()["u"]
    idx = np.unique(rng.integers(0, P.U, size=(3 * R, P.K)), axis=0)
    idx = idx[rng.permutation(len(idx))[:R]]
    rant = np.stack([u[k][idx[:, k]] for k in range(P.K)], axis=1)
    rconc = np.random.default_rng(1235).uniform(-1000, 1000, R)
    s0 = np.zeros((n, P.S)); s0[:, :2] = np.random.default_rng(4242).uniform(-np.pi / 2, np.pi / 2, size=(n, 2))
    res = []
    for _ in range(2):
        b = fq.Batch(P, n, P.ncorner)
        b.set_shared_rulebase(rant, rconc); b.set_initial_state(s0)
        b.run(steps, 0)
        res.append(b.get_state()); b.close()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    m = 8192                                            # the memoised path (default, above) == one full pass per step (policy 0)
    b = fq.Batch(P, m, P.ncorner); b.set_policy(0)
    b.set_shared_rulebase(rant, rconc); b.set_initial_state(s0[:m])
    b.run(steps, 0)
    s_full, a_full = b.get_state(); b.close()
    assert np.array_equal(s_full, res[0][0][:m]) and np.array_equal(a_full, res[0][1][:m])
    OP = _oracle.Problem(1, trig=1)
    for i in (0, 65535, n - 1):
        ag = _oracle.Agent(OP, R + 8); ag.set_rb(rant, rconc); ag.set_learning(0); ag.set_initial_state(s0[i])
        ag.run_steps(steps)
        assert np.array_equal(res[0][0][i], ag.summary()["s"]), i
    P.close()
