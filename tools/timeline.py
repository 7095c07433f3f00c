"""Experiments only: per-warp timeline of k_run_steps inside one bench step (engine built with -DFRIQ_TIMELINE,
selected with FRIQ_LIB=...).  Writes gpurun_out/timeline_<tag>.npy with rows (t0_ns, t1_ns, smid, warp_id, rmax, rmin, n16)
and prints how busy the SMs were.  Usage: FRIQ_LIB=.../friq_tl.so python tools/timeline.py [tag] [warm_steps]"""
import ctypes as C
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "fri-reinforcementlearning-c_b200"))
import friq_b200 as fq  # noqa: E402


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "t"
    warm = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    n = int(os.environ.get("AGENTS", 65536))
    L = fq.lib()
    L.friqd_timeline_fetch.restype = C.c_int
    L.friqd_timeline_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int]
    P = fq.Problem("mountaincar", no_random=0)
    b = fq.Batch(P, n, 192, seed0=0x5EED0000)
    b.set_retile_interval(int(os.environ.get("RETILE", 125)))
    mode = fq.RUN_LEARN | fq.RUN_RETILE
    for _ in range(warm):
        b.run(250, mode)
    L.friqd_timeline_fetch(None, 0, 1)
    b.run(250, mode)
    buf = np.zeros((1 << 18, 6), dtype=np.uint64)
    m = L.friqd_timeline_fetch(buf.ctypes.data_as(C.c_void_p), 1 << 18, 1)
    buf = buf[:m]
    t0 = buf[:, 0].astype(np.int64); t1 = buf[:, 1].astype(np.int64)
    smid = (buf[:, 2] >> np.uint64(32)).astype(np.int64); wid = (buf[:, 2] & np.uint64(0xffffffff)).astype(np.int64)
    rmax = ((buf[:, 3] >> np.uint64(32)) & np.uint64(0xffff)).astype(np.int64); rmin = (buf[:, 3] & np.uint64(0xffffffff)).astype(np.int64)
    n16 = (buf[:, 3] >> np.uint64(48)).astype(np.int64)
    base = t0.min()
    rows = np.stack([t0 - base, t1 - base, smid, wid, rmax, rmin, n16], axis=1)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    np.save(os.path.join(REPO, "gpurun_out", "timeline_%s.npy" % tag), rows)
    span = (t1.max() - base) / 1e6
    busy = (t1 - t0).sum() / 1e6
    slots = 148 * 2 * (128 // 32)
    print("%s: %d warp records, span %.3f ms, sum of warp times %.1f ms -> %.1f %% of %d warp slots busy" % (tag, m, span, busy, 100 * busy / (span * slots), slots))
    d = (t1 - t0) / 1e3
    for lo, hi in ((0, 20), (20, 30), (30, 40), (40, 50), (50, 60), (60, 80), (80, 100), (100, 200)):
        sel = (rmax >= lo) & (rmax < hi)
        if sel.any():
            kc = (buf[:, 4] >> np.uint64(32)).astype(np.float64)[sel]; pc = (buf[:, 4] & np.uint64(0xffffffff)).astype(np.float64)[sel]
            rounds = (buf[:, 5] >> np.uint64(48)).astype(np.float64)[sel]; pairs = ((buf[:, 5] >> np.uint64(24)) & np.uint64(0xffffff)).astype(np.float64)[sel]
            walk = (buf[:, 5] & np.uint64(0xffffff)).astype(np.float64)[sel]
            print("  rmax [%3d,%3d): %5d warps, mean %.1f us, max %.1f us per launch | kernel %.0f kcyc, in pair sums %.0f kcyc (%.0f %%), rounds %.0f, pairs %.0f, "
                  "walked rules (max per round) %.0f -> %.1f cyc per walked rule, rest %.0f cyc per step"
                  % (lo, hi, sel.sum(), d[sel].mean(), d[sel].max(), kc.mean() / 1e3, pc.mean() / 1e3, 100 * pc.mean() / kc.mean(), rounds.mean(), pairs.mean(),
                     walk.mean(), pc.mean() / max(walk.mean(), 1), (kc.mean() - pc.mean()) / 125))


if __name__ == "__main__":
    main()
