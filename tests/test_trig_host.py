"""CPU test of the PRODUCT's correctly-rounded sin/cos (csrc/friq_trig.h, the same source nvcc
compiles for the device) against binary128 libquadmath: every result must be the correctly
rounded double, and the double-double value must be accurate to < 2^-90 relative."""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_trig_is_correctly_rounded(tmp_path):
    so = str(tmp_path / "trig_shim.so")
    subprocess.run(["gcc", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared",
                    "-I", os.path.join(REPO, "fri-reinforcementlearning-c_b200", "csrc"), "-o", so,
                    os.path.join(REPO, "tests", "trig_host_shim.c"), "-lquadmath", "-lm"], check=True)
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    L.shim_check.restype = C.c_long; L.shim_check.argtypes = [dp, C.c_long, dp]
    L.shim_check_one.restype = C.c_long; L.shim_check_one.argtypes = [dp, C.c_long, dp]
    L.shim_dd_err.restype = C.c_double; L.shim_dd_err.argtypes = [dp, C.c_long]
    L.shim_quick_err.restype = C.c_double; L.shim_quick_err.argtypes = [dp, C.c_long, dp]
    rng = np.random.default_rng(2024)
    k = np.arange(-12, 13)
    sets = [rng.uniform(-10, 10, 300000), rng.uniform(-4.5, 1.5, 200000), rng.uniform(-1e-3, 1e-3, 50000),
            np.concatenate([k * np.pi / 2 + d for d in (1e-15, -1e-15, 3e-16, 1e-9, -1e-9, 1e-5)])]
    for x in sets:
        x = np.ascontiguousarray(x)
        w = C.c_double()
        assert L.shim_check(x.ctypes.data_as(dp), x.size, C.byref(w)) == 0, w.value
        assert L.shim_check_one(x.ctypes.data_as(dp), x.size, C.byref(w)) == 0, w.value      # friq_sin / friq_cos on their own
        assert L.shim_dd_err(x.ctypes.data_as(dp), min(x.size, 40000)) < 1.0
        # quick phase (Ziv): its error must stay well inside the bound the rounding test assumes (FRIQ_QUICK_ERR = 2^-61),
        # and the test must send only a small share of the calls to the accurate phase
        fb = C.c_double()
        assert L.shim_quick_err(x.ctypes.data_as(dp), min(x.size, 100000), C.byref(fb)) < 2.0 and fb.value < 0.02
