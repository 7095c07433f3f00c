"""Drop-in tests (GPU): the reference's own C API served by the B200 engine.

(1) The reference's UNMODIFIED example drivers (examples/<env>/<env>.c, compiled against include/dropin and
    linked to libfriq_b200.so by tests/dropin/Makefile) must reproduce the reference's tests/orig golden rule
    bases byte for byte - the reference's own regression test (tests/_test.sh:8-17), run on the GPU.
(2) mountaincar exactly as shipped (`-f <constructed rule base>` -> rule-base REDUCTION,
    frirl_sequential_run.c:170-350) must reproduce the reference's reduced rule base and its log.
(3) tests/dropin/api_probe (our own C program on frirl_* / FIVE_* / five_*) against the oracle, call by call,
    bit-exact."""
import os
import subprocess

import numpy as np
import pytest

import _golden
import _oracle

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "dropin", "_build")


def _exe(name):
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "dropin")], check=True)
    path = os.path.join(BUILD, name)
    assert os.path.exists(path), f"{path} missing: build it where the reference sources exist (make -C tests/dropin)"
    return path


@pytest.mark.parametrize("env", _golden.ENVS)
def test_reference_example_driver_reproduces_tests_orig(env, tmp_path):
    r = subprocess.run([_exe(f"dropin_{env}_crtrig"), "-q"], cwd=tmp_path, env=dict(os.environ, DROPIN_CONSTRUCT="1"),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert open(tmp_path / f"{env}.frirlrb.txt", "rb").read() == _golden.golden_text(env)
    assert open(tmp_path / f"{env}.frirlrb.bin", "rb").read() == open(os.path.join(_golden.GOLDEN, f"{env}.frirlrb.bin"), "rb").read()
    eps = [l for l in r.stdout.splitlines() if " Episode: " in l]
    want = _golden.load(env)["episodes"]
    assert len(eps) == len(want)
    for line, (ep, steps, reward, rules) in zip(eps, want):            # the per-episode printf of sequential_run.c:77-80
        assert f"Episode: {ep}\tSteps: {steps}\t" in line and line.endswith(f"Rules: {rules}") and ("%f" % reward) in line


def test_shipped_mountaincar_rule_base_reduction(tmp_path):
    rb = os.path.join(_golden.GOLDEN, "mountaincar.frirlrb.bin")
    r = subprocess.run([_exe("dropin_mountaincar_crtrig"), "-f", rb], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert open(tmp_path / "mountaincar.frirlrb.txt", "rb").read() == open(os.path.join(_golden.GOLDEN, "mountaincar_reduced.frirlrb.txt"), "rb").read()
    keep = [l for l in r.stdout.splitlines() if l.startswith(("Reduction Episode", "Smallest", "Loaded"))]
    keep = [l.split(" rules from")[0] if l.startswith("Loaded") else l for l in keep]
    assert keep == open(os.path.join(_golden.GOLDEN, "mountaincar_reduction.log")).read().splitlines()


@pytest.mark.parametrize("W", [3, 4])
def test_reference_example_driver_in_omp_mode_reproduces_reference_openmp_build(W, tmp_path):
    """`mountaincar -m omp` (frirl_run -> frirl_omp_run, frirl_agent.c:294-385): the reference needs its experimental CMake option
    BUILD_OPENMP for this; oracle/Makefile built it that way and oracle/make_golden.py recorded runs with OMP_NUM_THREADS = 3 and
    4.  The UNMODIFIED example driver on the drop-in API must print the same per-agent episode sequences and write the same
    final rule base, byte for byte."""
    import re
    r = subprocess.run([_exe("dropin_mountaincar_crtrig"), "-m", "omp"], cwd=tmp_path, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, DROPIN_CONSTRUCT="1", OMP_NUM_THREADS=str(W)))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    eps = {}
    for l in r.stdout.splitlines():
        m = re.match(r"#(\d+) Episode: (\d+)\tSteps: (\d+)\tReward: (?:\x1b\[[0-9;]*m)?(-?[0-9.]+)(?:\x1b\[[0-9;]*m)?\tRules: (\d+)", l)
        if m:
            eps.setdefault(int(m.group(1)), []).append((int(m.group(3)), float(m.group(4)), int(m.group(5))))
    z = np.load(os.path.join(_golden.GOLDEN, f"mountaincar_omp{W}.npz"))["episodes"]
    for a in range(W):
        want = [tuple(x) for x in z[z[:, 0] == a][:, 1:].tolist()]
        assert [tuple(map(float, e)) for e in eps.get(a, [])] == want, a
    assert open(tmp_path / "mountaincar.frirlrb.txt", "rb").read() == open(os.path.join(_golden.GOLDEN, f"mountaincar_omp{W}.frirlrb.txt"), "rb").read()


@pytest.mark.parametrize("W", [3, 5])
def test_reference_example_driver_in_mpi_mode_matches_oracle(W, tmp_path):
    """`mountaincar -m mpi` (frirl_run -> frirl_mpi_run, frirl_agent.c:387-467).  The reference needs BUILD_MPI and an MPI
    launcher, neither of which exists in this image, so there is no recorded reference run: PARITY UNPINNED for this mode - the
    checker is the oracle port with the MPI exchange order (members -> leader in rank order, then leader -> members, no "pended"
    test), whose merge_rb itself IS pinned by the reference's OpenMP runs.  World size from the launcher's OMPI_COMM_WORLD_SIZE; the
    ranks are agents of one process, spread over the box's GPUs."""
    import re
    r = subprocess.run([_exe("dropin_mountaincar_crtrig"), "-m", "mpi"], cwd=tmp_path, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, DROPIN_CONSTRUCT="1", OMPI_COMM_WORLD_SIZE=str(W)))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    eps = {}
    for l in r.stdout.splitlines():
        m = re.match(r"#(\d+) Episode: (\d+)\tSteps: (\d+)\tReward: (?:\x1b\[[0-9;]*m)?(-?[0-9.]+)(?:\x1b\[[0-9;]*m)?\tRules: (\d+)", l)
        if m:
            eps.setdefault(int(m.group(1)), []).append((int(m.group(3)), float(m.group(4)), int(m.group(5))))
    P = _oracle.Problem(0, trig=1)
    agents = [_oracle.Agent(P, 4096) for _ in range(W)]
    log = _oracle.team_run(agents, mpi_order=True)
    for a in range(W):
        want = [tuple(x) for x in log[log[:, 0] == a][:, 1:].tolist()]
        assert [tuple(map(float, e)) for e in eps.get(a, [])] == want, a
    R, rant, rconc = agents[0].rb()
    assert open(tmp_path / "mountaincar.frirlrb.txt", "rb").read() == _golden.rb_text(rant, rconc)
    assert "End" in r.stdout


def test_test_mode_runs_one_greedy_episode(tmp_path):
    """-m test (frirl_test_run.c:20-86): one greedy episode without rule-base updates, exit code 0 on success."""
    rb = os.path.join(_golden.GOLDEN, "acrobot.frirlrb.bin")
    r = subprocess.run([_exe("dropin_acrobot_crtrig"), "-m", "test", "-f", rb], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    OP = _oracle.Problem(1, trig=1)
    g = _golden.load("acrobot")
    a = _oracle.Agent(OP, 1024); a.set_rb(g["final"][1], g["final"][2]); a.set_learning(0)
    # one inference episode on the oracle
    while a.summary()["episode_num"] < 2:
        a.run_steps(1)
    sm = a.summary()
    assert f"Steps:\t{sm['last_steps']}\nRules:\t{sm['last_rules']}\nReward:\t{'%f' % sm['last_return']}\n" in r.stdout, r.stdout[-500:]
    good = sm["last_return"] > 0.0
    assert ("Success!" in r.stdout) == good and r.returncode == (0 if good else 255)
    assert not (tmp_path / "acrobot.frirlrb.txt").exists()              # test mode exits before the dump


def _hex(v):
    return " ".join(float(x).hex() for x in np.atleast_1d(v))


@pytest.mark.parametrize("env", ["mountaincar", "acrobot", "cartpole"])
def test_reference_api_calls_match_oracle(env, tmp_path):
    rng = np.random.default_rng(42)
    g = _golden.load(env)
    h = g["header"]
    K, S, A, U = h["K"], h["S"], h["A"], h["U"]
    OP = _oracle.Problem(_golden.ENV_ID[env], trig=1)
    ag = _oracle.Agent(OP, 4096)
    rant, rconc = g["final"][1], g["final"][2]
    ag.set_rb(rant, rconc)
    lo, hi = h["u"][:, 0], h["u"][:, -1]
    cmds, checks = [f"init {_golden.ENV_ID[env]}", "tables", f"load {os.path.join(_golden.GOLDEN, env + '.frirlrb.bin')}", "rb"], []
    pts = [rng.uniform(lo, hi) for _ in range(12)] + [rant[i] for i in (0, 5, len(rant) - 1)]
    pts += [np.where(rng.random(K) < 0.5, rant[7], rng.uniform(lo, hi)) for _ in range(4)]
    for x in pts:
        cmds += [f"concl {_hex(x)}", f"weight {_hex(x)}", f"dist {_hex(x)}", "bestact"]
    for x in pts[:8]:
        cmds.append(f"best {_hex(x[:S])}")
    for k in range(S):
        for x in rng.uniform(lo[k], hi[k], 5):
            cmds.append(f"snap {k} {_hex(x)}")
    # SARSA updates recorded from the reference run are replayed on top of the FINAL rule base (any state works)
    trans = []
    for r in g["sarsa"][:12]:
        trans.append((r["q_ant"], r["reward"], r["cur_q_ant"]))
        cmds.append(f"sarsa {_hex(r['q_ant'])} {_hex(r['reward'])} {_hex(r['cur_q_ant'])}")
    newrule = np.append(rng.uniform(lo, hi), 12.5)
    cmds += ["remove 3", f"add {_hex(newrule)}", "rb", f"concl {_hex(pts[0])}"]
    r = subprocess.run([_exe("api_probe")], input="\n".join(cmds) + "\n", cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    # frirl_init's WARNING lines (cartpole's grid steps, frirl_init.c:213-217) and five_remove_rule's chatter are not results
    out = [l for l in r.stdout.splitlines() if not l.startswith(("FIVERemoveRule", "frirl_init: WARNING", "WARNING"))]
    it = iter(out)

    def vals(line, skip):
        return np.array([float.fromhex(t) for t in line.split()[skip:]])

    assert next(it).split() == ["init", str(S), str(A), str(U), str(1 << K)]
    t = next(it).split()
    assert t[0] == "tables" and t[1] == "1"                              # frirl_init_ve == the engine's tables
    tv = np.array([float.fromhex(x) for x in t[2:]])
    assert np.array_equal(tv[:2 * K * U:2], h["u"].ravel()) and np.array_equal(tv[1:2 * K * U:2], h["ve"].ravel())
    assert np.array_equal(tv[2 * K * U:], h["act_ve"])
    assert next(it).split() == ["load", str(len(rconc))]

    def check_rb(R, rant_w, rconc_w):
        assert next(it).split() == ["rb", str(R)]
        for i in range(R):
            f = next(it).split()
            for k in range(K):
                idx = int(L.fo_nearest_index(OP.h, k, float(rant_w[i, k]), U))
                assert float.fromhex(f[3 * k]) == rant_w[i, k] and int(f[3 * k + 1]) == idx and float.fromhex(f[3 * k + 2]) == h["ve"][k, idx]
            assert float.fromhex(f[-1]) == rconc_w[i]

    L = _oracle.lib()
    import ctypes as C
    L.fo_nearest_index.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int]
    check_rb(len(rconc), rant, rconc)
    for x in pts:
        hit, conc = ag.vag_concl(x)
        f = next(it).split()
        assert f[0] == "concl" and int(f[1]) == hit and float.fromhex(f[2]) == conc and float.fromhex(f[3]) == conc
        hit_w, w = ag.vag_concl_weight(x)
        f = next(it)
        assert int(f.split()[1]) == hit_w
        if hit_w < 0:
            assert np.array_equal(vals(f, 2), w)
        f = next(it)
        d = vals(f, 2)
        assert int(f.split()[1]) == hit and len(d) == len(rconc)
        if hit >= 0:
            assert d[hit] == 0.0 and np.all(d[:hit] != 0.0)
        else:                                                            # Shepard on these distances == the conclusion
            pw = np.array([L.fo_pow_x87(float(v), OP.p) for v in d])
            acc = ws = 0.0
            for wi, q in zip(1.0 / pw, rconc):
                acc += wi * q; ws += wi
            assert acc / ws == conc
        assert float.fromhex(next(it).split()[1]) == conc                # FIVEVagConcl_FRIRL_BestAct on the same distances
    for x in pts[:8]:
        b, ac = ag.best_action(x[:S])
        f = next(it)
        assert int(f.split()[1]) == b and np.array_equal(vals(f, 2), ac)
    for c in [c for c in cmds if c.startswith("snap")]:
        _, k, xv = c.split()
        assert float.fromhex(next(it).split()[1]) == L.fo_snap(OP.h, int(k), float.fromhex(xv))
    for q_ant, reward, cur in trans:
        ag.update_sarsa(q_ant, reward, cur)
        f = next(it).split()
        assert f[0] == "sarsa" and int(f[1]) == ag.rb()[0] and int(f[2]) == ag.inserted()
    R, orant, orconc = ag.rb()
    orant = np.vstack([np.delete(orant, 3, axis=0), newrule[:K]]); orconc = np.append(np.delete(orconc, 3), newrule[K])
    assert next(it).split() == ["remove", str(R - 1)]
    assert next(it).split() == ["add", str(R)]
    check_rb(R, orant, orconc)
    ag.set_rb(orant, orconc)
    hit, conc = ag.vag_concl(pts[0])
    f = next(it).split()
    assert int(f[1]) == hit and float.fromhex(f[2]) == conc


@pytest.mark.parametrize("tag", ["custom", "custom6", "custom8"])
def test_custom_problem_on_the_reference_api_matches_the_reference(tmp_path, tag):
    """tests/dropin/custom_problem.c - a program on the reference's public API for a problem that is none of the examples
    (3 state dimensions, 5 actions, skip_rules = 0, host-callback environment) - prints, built against the B200 engine, exactly
    what it prints built against the reference (golden recorded by oracle/make_golden.py from oracle/_ref/custom_ref): the
    per-episode log, the learned rule base and per-call results, all as hex floats.  Exercises the generic-shape kernels.
    custom6 / custom8: the same program with 5 / 7 state dimensions = 6 / 8 antecedents (FIVE_MAX_NUM_OF_UNIVERSES, FIVE.h:19:
    5- and 4-bit antecedent indexes in the packed rule word, 64 / 256 corner rules)."""
    r = subprocess.run([_exe(tag + "_dropin")], cwd=tmp_path, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    want = open(os.path.join(_golden.GOLDEN, tag + "_problem.out")).read()
    assert r.stdout.splitlines() == want.splitlines()
    assert open(tmp_path / "custom.frirlrb.txt", "rb").read() == open(os.path.join(_golden.GOLDEN, tag + "_problem.frirlrb.txt"), "rb").read()
