"""CPU tests of the drop-in layer: the library exports every function include/dropin/friq_refapi.h declares
(= everything the reference's public headers declare), the forwarding headers exist under the reference's file
names, the host-side table builders of the reference API are bit-identical to the recorded reference tables, and
a program written against the reference API compiles against include/dropin unchanged."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import _golden

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(REPO, "include", "dropin")


@pytest.fixture(scope="module")
def L():
    import friq_b200
    friq_b200.build()
    return C.CDLL(friq_b200.LIB_PATH)


def test_every_reference_prototype_is_exported(L):
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(DROPIN, "friq_refapi.h")).read(), flags=re.S)
    protos = set(re.findall(r"^\s*(?:struct\s+\w+\s*\*|unsigned\s+int|int|void|double\s*\*?|fri_float)\s*\*?\s*((?:frirl|FIVE|five)\w*)\s*\(", hdr, flags=re.M))
    # the 12 + 12 + 17 functions of frirl.h:42-64, frirl_utils.h:19-25, frirl_app_helpers.h:41-46, frirl_test.h:18, FIVE.h:79-103
    for name in ["frirl_init", "frirl_deinit", "frirl_init_ve", "frirl_init_rb", "frirl_episode", "frirl_e_greedy_selection",
                 "frirl_get_best_action", "frirl_check_possible_states", "frirl_update_sarsa", "frirl_sequential_run", "frirl_omp_run",
                 "frirl_mpi_run", "frirl_show_rb", "frirl_show_hex_rb", "frirl_save_rb_to_text_file", "frirl_save_rb_to_bin_file",
                 "frirl_load_rb_from_bin_file", "frirl_print_usage", "frirl_parse_cmdline", "frirl_gen_fixres_arr", "frirl_run",
                 "frirl_visualization_init", "frirl_visualization_deinit", "frirl_test_run",
                 "FIVEInit", "FIVEGScFunc", "FIVEGVagEnv", "FIVEValVag", "FIVEVagConcl", "FIVEVagConclWeight", "FIVEVagConcl_FRIRL_BestAct",
                 "FIVEAddRule", "five_vague_distance", "five_vague_distance_parallel", "five_rule_distance", "five_add_rule",
                 "five_remove_rule", "five_deinit", "FIVE_add_rule", "FIVE_vag_concl_weight", "FIVE_vag_concl", "FIVE_GSc_func"]:
        assert name in protos, f"{name} not declared in friq_refapi.h"
    for name in protos:
        assert hasattr(L, name), f"{name} declared but not exported by libfriq_b200.so"


def test_forwarding_headers_cover_the_reference_file_names():
    for h in ["config.h", "FIVE.h", "frirl.h", "frirl_types.h", "frirl_types_def.h", "frirl_utils.h", "frirl_app_helpers.h", "frirl_test.h"]:
        assert os.path.exists(os.path.join(DROPIN, h)), h


@pytest.mark.parametrize("env", _golden.ENVS)
def test_reference_table_builders_bit_identical(L, env):
    """FIVE_GSc_func + FIVEGVagEnv + frirl_gen_fixres_arr (host side of the reference API) against the tables recorded
    from the compiled reference."""
    h = _golden.load(env)["header"]
    K, S, U, A = h["K"], h["S"], h["U"], h["A"]
    dp = C.POINTER(C.c_double)
    L.FIVE_GSc_func.argtypes = [dp, C.c_int, C.c_int, dp, C.c_int, C.c_int, C.c_double, dp]
    L.FIVEGVagEnv.restype = dp; L.FIVEGVagEnv.argtypes = [dp, C.c_int, C.c_int, dp]
    L.frirl_gen_fixres_arr.argtypes = [dp, C.c_int, C.c_double]
    libc = C.CDLL(None); libc.free.argtypes = [C.c_void_p]
    for k in range(K):
        u = np.zeros(U)
        L.frirl_gen_fixres_arr(u.ctypes.data_as(dp), U, h["statedims"][k]["udiv"] if k < S else h["act_udiv"])
        assert np.array_equal(u, h["u"][k])
        if k < S:
            d = h["statedims"][k]
            pts = np.stack([d["values"], np.full(d["len"], d["steep"]), np.full(d["len"], d["steep"])], axis=1).ravel().copy()
            n = d["len"]
        else:
            pts = np.stack([np.arange(A) * (1.0 / (A - 1) * 2.0) - 1.0, np.full(A, float((A - 1) // 2)), np.full(A, float((A - 1) // 2))], axis=1).ravel().copy()
            n = A
        scf = np.zeros(U)
        assert L.FIVE_GSc_func(u.ctypes.data_as(dp), 1, U, pts.ctypes.data_as(dp), n, 3, float("nan"), scf.ctypes.data_as(dp)) == 0
        ve = L.FIVEGVagEnv(u.ctypes.data_as(dp), 1, U, scf.ctypes.data_as(dp))
        got = np.ctypeslib.as_array(ve, shape=(U,)).copy()
        libc.free(ve)
        assert np.array_equal(got, h["ve"][k]), k
    assert L.FIVE_GSc_func(u.ctypes.data_as(dp), 1, U, pts.ctypes.data_as(dp), n, 3, -1.0, scf.ctypes.data_as(dp)) == -1   # FIVEGScFunc.c:83-88


def test_program_written_against_the_reference_api_compiles(tmp_path):
    """The shape of the reference's example mains: by-value frirl_desc from frirl_desc_default, designated initialisers
    on frirl_dimension_desc, the FRIRL_ALLOC_* / FRIRL_GEN_FIXRES_* macros, callbacks, direct field access."""
    src = tmp_path / "prog.c"
    src.write_text(r'''
#include "frirl.h"
#include "frirl_types.h"
#include "frirl_types_def.h"
#include "frirl_app_helpers.h"
#include <stdio.h>
static void act(struct frirl_desc *f, fri_float a, fri_float *s, int n, fri_float *o) { for (int i = 0; i < n; i++) o[i] = s[i] + a; }
static void rew(struct frirl_desc *f, fri_float *s, int n, struct frirl_reward_desc *r) { r->value = -1; r->success = s[0] > 1; }
static void qnt(struct frirl_desc *f, fri_float *s, int n, fri_float *o) { for (int i = 0; i < n; i++) o[i] = s[i]; }
int main(int argc, char **argv) {
    struct frirl_desc frirl = frirl_desc_default;
    frirl_parse_cmdline(&frirl, argc, argv);
    frirl.do_action_func = act; frirl.get_reward_func = rew; frirl.quant_obs_func = qnt;
    struct frirl_dimension_desc d0 = { .values = (fri_float[]){-1, 0, 1}, .values_len = 3, .values_div = 1.0, .values_steep = 1.0,
                                       .values_def = 0.0, .universe_len = 41, .universe_div = 0.1 };
    FRIRL_ALLOC_UNIVERSE(d0); FRIRL_GEN_FIXRES_UNIVERSE(d0);
    frirl.statedims = (struct frirl_dimension_desc[]){d0}; frirl.statedims_len = 1;
    struct frirl_dimension_desc actiondim = { .values_len = 3, .values_div = 1.0, .universe_len = 41, .universe_div = 0.1 };
    FRIRL_ALLOC_DIM(actiondim); FRIRL_GEN_FIXRES_DIM(actiondim);
    frirl.actiondim = actiondim;
    if (argc > 5) { frirl_init(&frirl); frirl_run(&frirl, 1); printf("%d %f\n", frirl.fiverb->numofrules, frirl.fiverb->rconc[0]);
                    frirl_save_rb_to_text_file(&frirl, "x.txt"); frirl_deinit(&frirl); }
    return 0;
}
''')
    pkg = os.path.join(REPO, "fri-reinforcementlearning-c_b200")
    subprocess.run(["gcc", "-O2", "-I", DROPIN, str(src), "-o", str(tmp_path / "prog"), f"-L{pkg}", "-lfriq_b200", f"-Wl,-rpath,{pkg}", "-lm"], check=True)
    assert subprocess.run([str(tmp_path / "prog")]).returncode == 0
