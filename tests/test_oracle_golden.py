"""CPU tests: pin the oracle port (oracle/friq_oracle.c) against the reference.

Every comparison is bit-exact (np.array_equal on float64): tables built at init,
per-call vectors recorded from the compiled reference, whole learning
trajectories and the final rule bases = the reference's own tests/orig goldens
(tests/_test.sh:8-17 diffs exactly these files)."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

import _golden
import _oracle


@pytest.fixture(scope="module", params=_golden.ENVS)
def env(request):
    name = request.param
    return name, _golden.load(name), _oracle.Problem(_golden.ENV_ID[name], trig=1)


def test_golden_text_is_reference_tests_orig(env):
    name, g, _ = env
    md5 = {"mountaincar": "e58ce96db4996c4b9becca0715bf7d2c", "acrobot": "1e3c1cab9508be0d653830921f15abe6",
           "cartpole": "09a4b7c4ff0de099fcd038f92f889c7c"}[name]
    assert hashlib.md5(_golden.golden_text(name)).hexdigest() == md5 == g["md5"]


def test_init_tables_match_reference(env):
    name, g, P = env
    h = g["header"]
    assert (P.S, P.K, P.A, P.U, P.p) == (h["S"], h["K"], h["A"], h["U"], h["p"])
    assert np.array_equal(P.arr("u", P.K * P.U).reshape(P.K, P.U), h["u"])
    assert np.array_equal(P.arr("ve", P.K * P.U).reshape(P.K, P.U), h["ve"])
    assert np.array_equal(P.arr("udiv", P.K), h["udiv"])
    assert np.array_equal(P.arr("act_val", P.A), h["act_val"])
    assert np.array_equal(P.arr("act_ve", P.A), h["act_ve"])
    for k, d in enumerate(h["statedims"]):
        assert np.array_equal(P.grid(k), d["values"])
    R, rant, rconc = _oracle.Agent(P).rb()
    assert R == h["rb0"][0] and np.array_equal(rant, h["rb0"][1]) and np.array_equal(rconc, h["rb0"][2])
    hy = P.hyper
    assert [hy[0], hy[1], hy[2], hy[3], hy[4], hy[5], hy[6], hy[7], hy[8], hy[9]] == \
        [h[k] for k in ["alpha", "gamma", "epsilon", "qdiff_pos", "qdiff_neg", "qdiff_final_tol",
                        "reward_good_above", "wsig", "skip_rules", "no_random"]]


def test_get_best_action_vectors(env):
    name, g, P = env
    assert len(g["gba"]) >= 20
    for r in g["gba"]:
        a = _oracle.Agent(P); a.set_rb(r["rant"], r["rconc"])
        best, actconc = a.best_action(r["states"])
        assert best == r["best"]
        assert np.array_equal(actconc, r["actconc"])


def test_update_sarsa_vectors(env):
    name, g, P = env
    cats = set()
    for r in g["sarsa"]:
        cats.add(r["cat"])
        a = _oracle.Agent(P); a.set_rb(r["rant"], r["rconc"]); a.set_inserted(r["inserted_before"])
        # nested FIVE_vag_concl calls; the first two (qp, qnow) run on the unmodified rule base
        for c in r["concl"]:
            hit, conc = a.vag_concl(c["ant"])
            assert hit == c["ret"] and conc == c["conc"]
        if r["weight"] is not None and r["R_after"] == r["R"]:
            hit, w = a.vag_concl_weight(r["weight"]["ant"])
            assert hit == r["weight"]["ret"]
            if hit < 0:
                assert np.array_equal(w, r["weight"]["weights"])
        a.update_sarsa(r["q_ant"], r["reward"], r["cur_q_ant"])
        R, rant, rconc = a.rb()
        assert R == r["R_after"]
        assert np.array_equal(rant, r["rant_after"])
        assert np.array_equal(rconc, r["rconc_after"])
        assert a.inserted() == r["inserted_after"]
    assert {0, 2}.issubset(cats)          # insertion and exact-hit update occur in every example


def test_construct_run_reproduces_reference_golden(env):
    name, g, P = env
    a = _oracle.Agent(P)
    log = a.run_construct()
    ref = np.array(g["episodes"], dtype=np.float64)
    assert log.shape == ref.shape
    assert np.array_equal(log, ref)
    R, rant, rconc = a.rb()
    assert np.array_equal(rant, g["final"][1]) and np.array_equal(rconc, g["final"][2])
    assert _golden.rb_text(rant, rconc) == _golden.golden_text(name)


def test_philox_known_answers():
    L = _oracle.lib()
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        L.fo_philox4x32_10(c, k, o)
        assert tuple(o) == want


def test_pow_x87_against_exact_integer_model():
    """fast_pow.inl:17-27: 64-bit-mantissa product chain, each step RNE, then RNE to 53 bits."""
    from fractions import Fraction
    import math

    def rne(x, bits):
        e = math.floor(math.log2(x)); q = Fraction(2) ** (e - bits + 1)
        while x >= q * 2 ** bits: q *= 2
        while x < q * 2 ** (bits - 1): q /= 2
        n = x / q; f = math.floor(n); r = n - f
        if r > Fraction(1, 2) or (r == Fraction(1, 2) and f % 2 == 1): f += 1
        return f * q
    rng = np.random.default_rng(7)
    L = _oracle.lib()
    for p in (3, 5):
        for d in np.concatenate([rng.uniform(1e-3, 50.0, 400), np.sqrt(rng.uniform(1e-6, 900.0, 400))]):
            acc = Fraction(float(d))
            for _ in range(p - 1):
                acc = rne(acc * Fraction(float(d)), 64)
            assert L.fo_pow_x87(float(d), p) == float(rne(acc, 53))


@pytest.mark.parametrize("env,W", [("mountaincar", 3), ("mountaincar", 4), ("acrobot", 3), ("cartpole", 3)])
def test_knowledge_sharing_run_reproduces_reference_openmp_build(env, W):
    """the reference's experimental multi-agent mode (frirl_omp_run + merge_rb, frirl_agent.c; CMake option
    BUILD_OPENMP) was compiled and run here with OMP_NUM_THREADS = W (oracle/Makefile, oracle/make_golden.py).  The port must
    reproduce every agent's episode sequence (steps, return, rules) and agent 0's final rule base - which pins merge_rb
    including its use of STALE FIVERB.weights after exact hits."""
    z = np.load(os.path.join(_golden.GOLDEN, f"{env}_omp{W}.npz"))["episodes"]
    P = _oracle.Problem(_golden.ENV_ID[env], trig=1)
    agents = [_oracle.Agent(P, 4096) for _ in range(W)]
    log = _oracle.team_run(agents)
    for a in range(W):
        want = z[z[:, 0] == a][:, 1:]; got = log[log[:, 0] == a][:, 1:]
        assert _golden.same_episode_rows(want, got), (a, len(want), len(got))
    R, rant, rconc = agents[0].rb()
    txt = "".join("".join("%.18f " % v for v in rant[i]) + "%.18f \n" % rconc[i] for i in range(R))
    assert txt == open(os.path.join(_golden.GOLDEN, f"{env}_omp{W}.frirlrb.txt")).read()
