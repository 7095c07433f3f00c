"""CPU tests of the product's host layer (no kernel launches): the C-ABI library loads and
exports every symbol include/friq_b200.h declares, the init-time tables are bit-identical to
the reference's, and the rule-base file formats are byte-compatible."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _golden

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fq():
    import friq_b200
    friq_b200.build()
    return friq_b200


def test_library_exports_every_declared_symbol(fq):
    hdr = open(os.path.join(REPO, "include", "friq_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(friq_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 28
    L = C.CDLL(fq.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/friq_b200.h but not exported"
    assert declared == set(fq.SYMBOLS), "python binding and header disagree"


@pytest.mark.parametrize("env", _golden.ENVS)
def test_host_tables_bit_identical_to_reference(fq, env):
    g = _golden.load(env)["header"]
    P = fq.Problem(env)
    assert (P.S, P.K, P.A, P.U, P.p) == (g["S"], g["K"], g["A"], g["U"], g["p"])
    t = P.tables()
    for name in ["u", "ve", "udiv", "act_val", "act_ve"]:
        assert np.array_equal(t[name], g[name]), name
    d = P.desc
    assert [d.alpha, d.gamma, d.epsilon, d.qdiff_pos_boundary, d.qdiff_neg_boundary, d.qdiff_final_tolerance,
            d.reward_good_above, d.rule_weight_considered_significant_for_update, d.skip_rules, d.no_random,
            d.max_episodes, d.max_steps] == \
        [g[k] for k in ["alpha", "gamma", "epsilon", "qdiff_pos", "qdiff_neg", "qdiff_final_tol", "reward_good_above",
                        "wsig", "skip_rules", "no_random", "max_episodes", "max_steps"]]
    P.close()


@pytest.mark.parametrize("env", _golden.ENVS)
def test_rule_base_file_formats(fq, env, tmp_path):
    K = _golden.load(env)["header"]["K"]
    rant, rconc = fq.load_rb_bin(os.path.join(_golden.GOLDEN, f"{env}.frirlrb.bin"), K)
    final = _golden.load(env)["final"]
    assert np.array_equal(rant, final[1]) and np.array_equal(rconc, final[2])
    fq.save_rb_text(str(tmp_path / "rb.txt"), rant, rconc)
    fq.save_rb_bin(str(tmp_path / "rb.bin"), rant, rconc)
    assert open(tmp_path / "rb.txt", "rb").read() == _golden.golden_text(env)
    assert open(tmp_path / "rb.bin", "rb").read() == open(os.path.join(_golden.GOLDEN, f"{env}.frirlrb.bin"), "rb").read()


def test_rejects_unsupported_problems(fq):
    with pytest.raises(fq.FriqError):
        fq.Problem("mountaincar", statedims_len=7)


def test_no_cpu_fallback_without_gpu(fq):
    """On a box without a CUDA device every compute entry point must fail loudly."""
    try:
        n = fq.device_count()
    except fq.FriqError:
        n = 0
    if n > 0:
        pytest.skip("a GPU is present")
    P = fq.Problem("mountaincar")
    with pytest.raises(fq.FriqError):
        fq.Batch(P, 32, 192)
