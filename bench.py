#!/usr/bin/env python3
"""bench.py -- benchmark of the B200-native FRIQ-learning engine.

Metric (BASELINE.json): agent env-steps/sec (FIVE evals/sec reported beside it).  One bench "step" = one pass of the
hot path over one batch: EVERY agent advances `env_steps_per_agent_per_step` env steps through the fused step kernel
(dynamics + best action over all actions + epsilon-greedy + SARSA update / rule insertion).

  --workload c3        (default; BASELINE configs[2], the headline) 65,536 mountaincar learners per GPU, eps = 0.01
             acrobot   262,144 acrobot learners per GPU (configs[2]'s shape on the acrobot problem), eps = 0.001
             cartpole  65,536 cartpole learners per GPU (A = 21), eps = 0.001
             c4        configs[3]: 262,144 acrobot agents IN TOTAL against ONE shared 10,240-rule base, greedy inference,
                       200 steps per episode from random start states (strong scaling over --gpus)
             c5        configs[4]: 1,048,576 mountaincar agents IN TOTAL, alpha x gamma x epsilon grid x 512 seeds (strong scaling)
             c1        configs[0]: examples/mountaincar, single agent, construct run until the rule base is complete
             c2        configs[1]: examples/cartpole and examples/acrobot, single agent each
  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU, torchrun for N > 1)
  python bench.py --impl reference ...                            the reference's own CPU implementation of the same workload
                                                                  (oracle/_ref: the unmodified reference, all host cores)
Prints ONE JSON line on rank 0.  `config` is identical in both arms; arm-specific facts live in `stats` / `arm`.
"""
import argparse
import ctypes
import hashlib
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "fri-reinforcementlearning-c_b200")
sys.path.insert(0, PKG)

METRIC = "agent_env_steps_per_sec"
UNIT = "env-steps/s"

# algorithmic work of one env step with R rules (FP64 ops: sub/mul/add/cmp/sqrt/div = 1 each) and the
# index-form bytes of a private rule base streamed once per step (R * (K * 1 B + 8 B) read + R * 8 B written back) + agent state
WORKLOADS = {
    "c3": dict(env="mountaincar", agents=65536, total=None, cap=192, chunk=250, retile=125, eps=0.01, f_rule=82.0, b_rule=19.0,
               evals=6.5, kernel="k_run_steps<2,3,3,3,mountaincar> (table mode)",
               text="mountaincar FRIQ-learning, %d agents/GPU, eps-greedy 0.01, learning from the corner rule base"),
    "acrobot": dict(env="acrobot", agents=262144, total=None, cap=704, chunk=250, retile=125, eps=0.001, f_rule=118.0, b_rule=21.0,
                    evals=6.5, kernel="k_run_steps<4,5,3,3,acrobot> (table mode)",
                    text="acrobot FRIQ-learning, %d agents/GPU, eps-greedy 0.001, learning from the corner rule base"),
    "cartpole": dict(env="cartpole", agents=65536, total=None, cap=512, chunk=250, retile=125, eps=0.001, f_rule=334.0, b_rule=21.0,
                     evals=24.5, kernel="k_run_steps<4,5,21,3,cartpole> (lock-step pass)",
                     text="cartpole FRIQ-learning (21 actions), %d agents/GPU, eps-greedy 0.001, learning from the corner rule base"),
    "c5": dict(env="mountaincar", agents=None, total=1048576, cap=192, chunk=250, retile=125, eps=None, f_rule=82.0, b_rule=19.0,
               evals=6.5, kernel="k_run_steps<2,3,3,3,mountaincar> (table mode)",
               text="mountaincar hyper-parameter sweep, 1,048,576 agents in total (%d/GPU): alpha 16 x gamma 16 x epsilon 8 x 512 seeds"),
    "c4": dict(env="acrobot", agents=None, total=262144, cap=32, chunk=200, retile=0, eps=0.0, f_rule=47.0, b_rule=0.0,
               evals=3.0, rules=10240, kernel="k_infer_memo<4,3,acrobot> + k_shared_start<4,5,3>",
               text="acrobot greedy inference, 262,144 agents in total (%d/GPU) against ONE shared 10,240-rule base, "
                    "200-step episodes from random start states"),
    "c1": dict(envs=["mountaincar"], text="examples/mountaincar, ONE agent, construct run until the rule base is complete (no_random = 1)"),
    "c2": dict(envs=["cartpole", "acrobot"], text="examples/cartpole and examples/acrobot, ONE agent each, construct runs (no_random = 1)"),
}
CAP1 = {"mountaincar": 192, "acrobot": 704, "cartpole": 512}
ENV_A = {"mountaincar": 3, "acrobot": 3, "cartpole": 21}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--agents", type=int, default=0, help="agents per GPU (0 = the workload's own size)")
    ap.add_argument("--chunk", type=int, default=0, help="env steps per agent per bench step (0 = the workload's own)")
    ap.add_argument("--retile", type=int, default=-1, help="env steps between re-tilings by rule count inside a step (-1 = the workload's own)")
    ap.add_argument("--groups", type=int, default=0, help="agent groups on their own streams (0 = the engine's default)")
    ap.add_argument("--nccl-high-priority", type=int, default=1, help="run the result all-gathers on a high-priority NCCL stream")
    ap.add_argument("--policy", type=int, default=-1, help="evaluation policy of the fused kernel (friq_batch_set_policy; -1 = the engine's default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def resolve(args):
    """(workload dict, agents per GPU, chunk, retile) for this run; strong-scaling workloads divide their total by --gpus."""
    w = WORKLOADS[args.workload]
    if "envs" in w:
        return w, 1, 0, 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    gpus = max(args.gpus, world)
    n = args.agents or (w["agents"] if w.get("agents") else w["total"] // gpus)
    chunk = args.chunk or w["chunk"]
    retile = w["retile"] if args.retile < 0 else args.retile
    return w, n, chunk, retile


def make_config(args, w, n, chunk, retile):
    """The workload description both arms print verbatim."""
    if args.workload in ("c1", "c2"):
        return {"workload": w["text"], "agents": 1, "mode": "construct", "no_random": 1}
    return {"workload": w["text"] % n, "agents_per_gpu": n, "env_steps_per_agent_per_step": chunk,
            "steps_per_agent": (args.warmup + args.steps) * chunk if args.workload != "c4" else chunk,
            "scaling": "strong" if w.get("total") else "weak"}


# ------------------------------------------------------------------------------------------ synthetic inputs
def fixres(n, step):
    """frirl_gen_fixres_arr (frirl_app_helpers.c:32-44): symmetric fixed-resolution array."""
    a = np.zeros(n)
    first = -((n - 1) * step) / 2
    for i in range(n // 2 + 1):
        a[i] = first + step * i
    for i in range(n // 2 + 1, n):
        a[i] = a[n - 1 - i] * -1
    return a


def c4_rulebase(R=10240):
    """the 32 acrobot corner rules (frirl_init_rb.c:99-126, q = 0) + distinct random rules on universe points,
    consequents U(-1000, 1000).  Pure numpy (the reference arm must not load the engine)."""
    U = 41
    u = [fixres(U, d) for d in (0.1, 0.1, 0.05, 0.05, 0.1)]           # acrobot.c:261-310 universes
    ext = [np.pi / 2 * 0 + 1.570796326794897, 1.570796326794897, 0.785398163397448, 0.785398163397448, 1.0]
    corners = np.array([[(-ext[k] if (r // (32 >> (k + 1))) % 2 == 0 else ext[k]) for k in range(5)] for r in range(32)])
    rng = np.random.default_rng(1234)
    idx = np.unique(rng.integers(0, U, size=(3 * R, 5)), axis=0)
    idx = idx[rng.permutation(len(idx))[:R - 32]]
    rant = np.concatenate([corners, np.stack([u[k][idx[:, k]] for k in range(5)], axis=1)])
    rconc = np.concatenate([np.zeros(32), np.random.default_rng(1235).uniform(-1000, 1000, R - 32)])
    return np.ascontiguousarray(rant), np.ascontiguousarray(rconc)


def c4_start_states(n, first):
    """theta1, theta2 ~ U(-pi/2, pi/2), speeds 0 (seed 4242, indexed by GLOBAL agent id so that shards agree with one big batch)."""
    rng = np.random.default_rng(4242)
    total = WORKLOADS["c4"]["total"]
    th = rng.uniform(-np.pi / 2, np.pi / 2, size=(max(total, first + n), 2))[first:first + n]
    s0 = np.zeros((n, 4)); s0[:, :2] = th
    return s0


def c5_hyper(first, n):
    i = first + np.arange(n)
    alpha = 0.05 * (1 + i % 16); gamma = 0.90 + 0.00625 * (1 + (i // 16) % 16)
    eps = np.array([0, 0.001, 0.002, 0.005, 0.01, 0.02, 0.05, 0.1])[(i // 256) % 8]
    return alpha, gamma, eps


def save_rb_bin(path, rant, rconc):
    """frirl_save_rb_to_bin_file layout (frirl_utils.c:175-203): int32 R, then per rule K antecedents + consequent, f64."""
    with open(path, "wb") as f:
        f.write(np.int32(len(rconc)).tobytes())
        f.write(np.concatenate([rant, rconc[:, None]], axis=1).astype(np.float64).tobytes())


# ------------------------------------------------------------------------------------------ CPU reference arm
def _ref_exe(env):
    exe = os.path.join(REPO, "oracle", "_ref", "ref_" + env)
    if not os.path.exists(exe):
        raise RuntimeError(f"{exe} missing: run `make -C oracle ref` where /root/reference exists")
    return exe


def _refbench(exe, env):
    out = subprocess.run([exe], env=env, capture_output=True, text=True, check=True).stdout
    m = re.search(r"steps=(\d+) busy_s_per_proc=([\d.]+) wall_s=([\d.]+) steps_per_s=([\d.]+) mean_rules=([\d.]+)", out)
    return float(m.group(1)), float(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))


def run_reference(workload, steps_per_agent, target_seconds=15.0):
    """Time the UNMODIFIED reference (oracle/_ref/ref_<env>, built by oracle/Makefile from the reference sources with its own
    flags -O2 -mavx2) on all host cores: one process per core, each stepping independent agents through frirl_episode() for
    `steps_per_agent` env steps per agent - a bounded sample of the workload.  Learning workloads run epsilon-greedy with the
    example's own epsilon (the reference's libc rand() is unseedable: same exploration rate, not the same stream)."""
    w = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    if workload in ("c1", "c2"):
        tot_steps, tot_busy, notes = 0.0, 0.0, []
        for env_name in w["envs"]:
            # one agent, as many episodes of frirl_episode() as the example's construct run has env steps (SURVEY 8(d) C1/C2:
            # 15,548 / 33,002 / 21,207 steps); time inside the process around the episodes only
            target = {"mountaincar": 15548, "cartpole": 33002, "acrobot": 21207}[env_name]
            steps, busy, _, _, _ = _refbench(_ref_exe(env_name), dict(os.environ, REF_MODE="bench", REF_PROCS="1", REF_STEPS=str(target), REF_REPS="1"))
            tot_steps += steps; tot_busy += busy
            notes.append(f"{env_name}: {int(steps)} steps in {busy:.4f} s")
        return dict(value=tot_steps / tot_busy, unit=UNIT, cores=1, kind="reference", sample="one reference agent per example, frirl_episode() time only; " + "; ".join(notes), mean_rules=None)
    if workload == "c4":
        exe = _ref_exe("acrobot")
        d = tempfile.mkdtemp()
        rant, rconc = c4_rulebase(w["rules"])
        rb = os.path.join(d, "c4.frirlrb.bin"); save_rb_bin(rb, rant, rconc)
        base = dict(os.environ, REF_MODE="infer", REF_PROCS=str(cores), REF_STEPS=str(w["chunk"]), REF_RBFILE=rb)
        _, busy, _, _, _ = _refbench(exe, dict(base, REF_REPS="1"))
        reps = max(1, min(2000, int(target_seconds / max(busy, 1e-3))))
        steps, busy, wall, rate, mr = _refbench(exe, dict(base, REF_REPS=str(reps)))
        return dict(value=rate, unit=UNIT, cores=cores, kind="reference", mean_rules=mr,
                    sample=f"{cores} processes x {reps} greedy {w['chunk']}-step episodes against the 10,240-rule base "
                           f"({int(steps)} steps, {busy:.2f} s busy per process, frirl_episode() time only)")
    exe = _ref_exe(w["env"])
    base = dict(os.environ, REF_MODE="bench", REF_PROCS=str(cores), REF_STEPS=str(steps_per_agent), REF_EGREEDY="1")
    _, busy, _, _, _ = _refbench(exe, dict(base, REF_REPS="2"))
    reps = max(2, min(20000, int(2 * target_seconds / max(busy, 1e-3))))
    steps, busy, wall, rate, mr = _refbench(exe, dict(base, REF_REPS=str(reps)))
    return dict(value=rate, unit=UNIT, cores=cores, kind="reference", mean_rules=mr,
                sample=f"{cores} processes x {reps} independent eps-greedy agents x {steps_per_agent} env steps each "
                       f"({int(steps)} steps, {busy:.2f} s busy per process, frirl_episode() time only)")


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the run by a background thread (NVML; one-shot nvidia-smi calls as a fallback)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period=0.02):
        import threading
        self.index, self.period = index, period
        self.sm, self.mx, self.reasons, self.how = [], [], set(), None
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        try:
            import pynvml as N
            N.nvmlInit()
            h = N.nvmlDeviceGetHandleByIndex(self.index)
            self.how = "nvml"
            bits = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}
            get_reasons = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)
            while not self._stop.is_set():
                self.sm.append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM))); self.mx.append(float(mx))
                r = get_reasons(h)
                self.reasons.update(k for k, b in bits.items() if r & b)
                time.sleep(self.period)
            return
        except Exception:       # noqa: BLE001 - fall back to nvidia-smi
            pass
        self.how = "nvidia-smi"
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout
                c = [x.strip() for x in out.strip().split(",")]
                self.sm.append(float(c[0])); self.mx.append(float(c[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
            except Exception:   # noqa: BLE001
                self.how = "unavailable"
                return

    def stop(self):
        self._stop.set(); self._t.join(timeout=10)
        return dict(sm_mhz=statistics.median(self.sm) if self.sm else None, sm_max_mhz=max(self.mx) if self.mx else None,
                    samples=len(self.sm), reasons=sorted(self.reasons), source=self.how)


def kernel_source_hash():
    """sha1 of the kernel sources with comments and white space removed: what the calibration is tied to is the CODE"""
    h = hashlib.sha1()
    for f in ("friq_engine.cu", "friq_device.cuh", "friq_math.cuh", "friq_trig.h", "friq_devprob.h"):
        t = open(os.path.join(PKG, "csrc", f), "r", errors="replace").read()
        t = re.sub(r"/\*.*?\*/", " ", t, flags=re.S)
        t = re.sub(r"//[^\n]*", " ", t)
        h.update(" ".join(t.split()).encode())
    return h.hexdigest()[:12]


def calibration(kernel_key):
    """Per-kernel constants from the committed ncu captures (profiles/calibration.json, written by tools/calibrate.py from
    `ncu --set full` + instruction-count metrics): DRAM bytes per launch and executed FP64 thread instructions per unit of counted
    work.  Used only when the capture was taken on THIS kernel source (sha1 of csrc/), else null - a stale capture is no claim."""
    try:
        c = json.load(open(os.path.join(REPO, "profiles", "calibration.json")))
        e = c.get(kernel_key)
        if e and e.get("source_sha1") == kernel_source_hash():
            return e
    except (OSError, ValueError):
        pass
    return None


# ------------------------------------------------------------------------------------------ single-agent workloads (c1, c2)
def run_single(args, w):
    """configs[0]/[1]: ONE agent.  `value`: the fused step kernel with a batch of one, one launch per episode (the per-episode
    log line of frirl_sequential_run needs the episode's result on the host) until the device-side "rule base complete"
    predicate fires.  `e2e`: the reference's UNMODIFIED example driver, compiled against the drop-in headers, run as a process on
    the drop-in library (one kernel + one H2D + one D2H per env step: the host callbacks step the environment)."""
    import torch
    import friq_b200 as fq
    torch.cuda.set_device(0)
    tot_steps = 0; tot_s = 0.0; tot_e2e_s = 0.0; tot_e2e_steps = 0; notes = []; launches = 0
    for env_name in w["envs"]:
        P = fq.Problem(env_name)
        b = fq.Batch(P, 1, CAP1[env_name])
        mode = fq.RUN_LEARN | fq.RUN_TRACK | fq.RUN_STOP_AT_EPISODE_END

        def construct():
            b.reset(); n_l = 0
            for _ in range(999):
                b.run(P.desc.max_steps, mode); n_l += 1
                o = b.gather()
                if o["flags"][0] & fq.F_CONVERGED:
                    break
            return int(o["total_steps"][0]), int(o["rules"][0]), int(o["episodes"][0]), n_l
        construct()                                           # warm-up
        t0 = time.perf_counter(); steps, rules, eps_n, n_l = construct(); dt = time.perf_counter() - t0
        tot_steps += steps; tot_s += dt; launches += n_l
        exe = os.path.join(REPO, "tests", "dropin", "_build", "dropin_" + env_name)
        e2e_note = "drop-in driver not built"
        if os.path.exists(exe):
            r = subprocess.run([exe, "-q"], cwd=tempfile.mkdtemp(), capture_output=True, text=True,
                               env=dict(os.environ, DROPIN_CONSTRUCT="1", DROPIN_TIMING="1"))
            m = re.search(r"DROPIN_TIMING frirl_run_s=([\d.]+)", r.stdout)
            ep = [l for l in r.stdout.splitlines() if " Episode: " in l]
            if r.returncode == 0 and m and ep:
                dsteps = sum(int(re.search(r"Steps: (\d+)", l).group(1)) for l in ep)
                tot_e2e_s += float(m.group(1)); tot_e2e_steps += dsteps
                e2e_note = f"{dsteps} steps in {float(m.group(1)):.3f} s of frirl_run()"
            else:
                e2e_note = f"drop-in driver exit {r.returncode}"
        notes.append(f"{env_name}: {steps} steps, {eps_n} episodes, {rules} rules, {dt * 1e3:.1f} ms fused B=1 ({n_l} launches); e2e {e2e_note}")
        b.close(); P.close()
    return tot_steps, tot_s, tot_e2e_s, tot_e2e_steps, notes, launches


# ------------------------------------------------------------------------------------------ our arm
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    w, n, chunk, retile = resolve(args)
    config = make_config(args, w, n, chunk, retile)
    steps_per_agent = (args.warmup + args.steps) * chunk
    evals = w.get("evals", 6.5)

    if args.impl == "reference":
        if rank != 0:
            return
        cb = run_reference(args.workload, steps_per_agent, float(os.environ.get("FRIQ_BENCH_REF_SECONDS", 15.0)))
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True, "scaling": config.get("scaling", "weak"), "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "arm": "the unmodified reference (oracle/_ref, C + AVX2 inline asm, -O2 -mavx2) on all host cores: one process per core "
                   "stepping independent single agents through frirl_episode(); a bounded sample of the workload",
            "stats": {"mean_rules": cb.get("mean_rules")},
            "cpu_baseline": {k: v for k, v in cb.items() if k != "mean_rules"},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "five_evals_per_sec": cb["value"] * evals}))
        return

    if args.workload in ("c1", "c2"):
        if rank != 0:
            return
        tot_steps, tot_s, tot_e2e_s, tot_e2e_steps, notes, launches = run_single(args, w)
        out = {"metric": METRIC, "value": tot_steps / tot_s, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": tot_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "config": config, "stats": {"runs": notes}, "gpu_launches": launches,
               "e2e": {"value": tot_e2e_steps / tot_e2e_s if tot_e2e_s > 0 else None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                       "note": "the reference's UNMODIFIED example drivers (libm trig) on the drop-in library: wall time of frirl_run(); every env "
                               "step is one kernel + one H2D + one D2H because the environment callbacks are host functions"},
               "roofline": None, "five_evals_per_sec": None}
        if not args.no_cpu_baseline:
            try:
                cb = run_reference(args.workload, 0)
                out["cpu_baseline"] = {k: v for k, v in cb.items() if k != "mean_rules"}
            except Exception as e:       # noqa: BLE001
                out["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(out))
        return

    import torch
    import friq_b200 as fq
    import friq_dist

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        # the exchange moves 16 B per agent: one NCCL channel is plenty, and NCCL's persistent CTAs - which sit on SM slots while
        # they wait for the slowest rank - stay out of the step kernels' way (measured at 2 GPUs: 3.91 -> 3.82 ms per step)
        os.environ.setdefault("NCCL_MAX_CTAS", "1")
        # stdout is the ONE JSON line: whatever NCCL prints while the communicator comes up (its version line under NCCL_DEBUG=VERSION)
        # goes to stderr - file descriptor 1 points there until the first collective has run
        sys.stdout.flush(); saved_stdout = os.dup(1); os.dup2(2, 1)
        try:
            try:                                 # the collective runs on NCCL's own stream: make it a high-priority one (it is tiny and must not queue behind the step kernels)
                if not args.nccl_high_priority:
                    raise AttributeError
                opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
                dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
            except (AttributeError, TypeError):
                dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved_stdout, 1); os.close(saved_stdout)
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())

    first_agent, _ = friq_dist.shard(world * n, world, rank)
    shared = args.workload == "c4"
    P = fq.Problem(w["env"], no_random=0)
    b = fq.Batch(P, n, w["cap"], device=local, seed0=0x5EED0000, first_agent=first_agent)
    if args.groups:
        b.set_groups(args.groups)
    if args.policy >= 0:
        b.set_policy(args.policy)
    if args.workload == "c5":
        h_alpha_np, h_gamma_np, h_eps_np = c5_hyper(first_agent, n)
    else:
        h_alpha_np, h_gamma_np, h_eps_np = np.full(n, P.desc.alpha), np.full(n, P.desc.gamma), np.full(n, w["eps"])
    b.set_hyper(h_alpha_np, h_gamma_np, h_eps_np)
    if shared:
        rant, rconc = c4_rulebase(w["rules"])
        b.set_shared_rulebase(rant, rconc)
        s0_np = c4_start_states(n, first_agent)
        h_s0 = torch.from_numpy(s0_np).pin_memory()
        d_s0 = torch.from_numpy(s0_np).cuda()
        mode = 0
    else:
        b.set_retile_interval(retile)
        mode = fq.RUN_LEARN | fq.RUN_RETILE      # agents are re-assigned to warps by rule count before every launch

    # the path's only exchange: per-agent (return f64, steps i32, rules i32), gathered over NCCL on a side stream that overlaps
    # the next step's kernels (double-buffered); see friq_dist.ResultExchange
    def stats_fn(g_ret, g_rules, m, out):        # one small kernel of the engine on the exchange's side stream
        fq.result_stats(local, vp(g_ret), vp(g_rules), m, vp(out), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    xch = friq_dist.ResultExchange(dist if world > 1 else None, n, "cuda", stats_fn=stats_fn)

    def gather_results():
        buf = xch.next_buffers(stream)
        b.gather_device(vp(buf[0]), vp(buf[1]), vp(buf[2]), sptr)
        xch.exchange_async(stream)

    def barrier():
        xch.wait(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, reset_fn, after_fn=None):
        reset_fn()
        clk = ClockSampler(local) if rank == 0 else None     # sampled from the warm-up on (same load)
        for _ in range(args.warmup):
            step_fn()
        barrier()
        o0 = b.gather(); b.work_counters(reset=True)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step_fn()
        if after_fn:
            after_fn()
        xch.wait(stream)
        e1.record(stream)
        barrier()
        clocks = clk.stop() if clk else None
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            per_rank = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(per_rank, ms)                      # every rank's own time (diagnostic: the reported time is their MAX)
            timed.ms_ranks = [float(t.item()) / args.steps for t in per_rank]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        o1 = b.gather(); ctr = b.work_counters()
        dsteps = int((o1["total_steps"] - o0["total_steps"]).sum()); drules = int((o1["sum_rules"] - o0["sum_rules"]).sum())
        tot = torch.tensor([dsteps, drules] + [ctr[k] for k in ("pairs_summed", "table_rules", "chain_rules", "hit_pairs", "appended", "wupd_rules")],
                           dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tot)
        return float(ms.item()), [int(x) for x in tot.tolist()], clocks

    # ---- (1) device-resident: inputs already in HBM
    if shared:
        def step_device():
            b.new_episodes(sptr)                               # every agent starts a fresh episode ...
            b.set_initial_state_device(vp(d_s0), sptr)         # ... from its start state (resident in HBM): one pass per agent + 200 memoised steps
            b.run(chunk, mode, stream=sptr, sync=False)
            gather_results()
    else:
        def step_device():
            b.run(chunk, mode, stream=sptr, sync=False)
            gather_results()

    def reset_batch():                           # friq_batch_reset restores the problem's own hyper-parameters: put the workload's back
        b.reset()
        b.set_hyper(h_alpha_np, h_gamma_np, h_eps_np)

    ms, tot, clocks = timed(step_device, reset_batch)
    ms_ranks = getattr(timed, "ms_ranks", None)
    nsteps, sumR = tot[0], tot[1]
    expect = world * n * chunk * args.steps
    assert nsteps == expect, "work skipped inside the timed region"

    # ---- (2) end to end through the C-ABI with HOST buffers inside the timed region
    h_alpha = torch.from_numpy(h_alpha_np.copy()).pin_memory(); h_gamma = torch.from_numpy(h_gamma_np.copy()).pin_memory()
    h_eps = torch.from_numpy(h_eps_np.copy()).pin_memory()
    # two sets of pinned result buffers: the host reads step i's results while step i + 1 is running (every step's results ARE
    # read, one step late; inputs go host -> device every step)
    h_out = [(torch.empty(n, dtype=torch.float64).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory(),
              torch.empty(n, dtype=torch.int32).pin_memory()) for _ in range(2)]
    pend = [None, None]; e2e_count = [0]; e2e_seen = [0.0]

    def e2e_collect(i):
        if pend[i] is not None:
            pend[i].synchronize(); pend[i] = None
            e2e_seen[0] += float(h_out[i][0][0]) + float(h_out[i][2][n - 1])      # the host really touches the results

    if shared:
        d_s0b = torch.empty_like(d_s0)

        def step_e2e():
            i = e2e_count[0] % 2; e2e_count[0] += 1
            e2e_collect(i)
            b.new_episodes(sptr)
            d_s0b.copy_(h_s0, non_blocking=True)               # H2D of this step's start states
            b.set_initial_state_device(vp(d_s0b), sptr)
            b.run_host_async(chunk, mode, None, None, None, vp(h_out[i][0]), vp(h_out[i][1]), vp(h_out[i][2]), sptr)
            pend[i] = torch.cuda.Event(); pend[i].record(stream)
        h2d, d2h = 8 * 4 * n, 16 * n
    else:
        def step_e2e():
            i = e2e_count[0] % 2; e2e_count[0] += 1
            e2e_collect(i)
            b.run_host_async(chunk, mode, vp(h_alpha), vp(h_gamma), vp(h_eps), vp(h_out[i][0]), vp(h_out[i][1]), vp(h_out[i][2]), sptr)
            pend[i] = torch.cuda.Event(); pend[i].record(stream)
        h2d, d2h = 3 * 8 * n, 16 * n
    ms_e2e, tot_e2e, _ = timed(step_e2e, reset_batch, after_fn=lambda: (e2e_collect(0), e2e_collect(1)))
    nsteps_e2e = tot_e2e[0]

    fp64_peak_g = fq.device_fp64_peak()          # measured DADD/DMUL lane-instr/s of this GPU (G/s)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    t = ms * 1e-3
    f_alg = w["f_rule"] * sumR / world                                   # per GPU
    bytes_alg = (w["b_rule"] * sumR + 256.0 * nsteps) / world
    ach_tflops = f_alg / t / 1e12
    ach_gbs = bytes_alg / t / 1e9
    frac_fp64 = ach_tflops / (fp64_peak_g / 1e3); frac_hbm = ach_gbs / hbm_peak

    # what the kernels EXECUTED in this run: counters accumulated on the device inside the timed region
    cal = calibration(args.workload)
    executed = {"pairs_summed": tot[2], "rules_summed_table": tot[3], "rules_summed_chains": tot[4], "pairs_exact_hit": tot[5],
                "rules_appended": tot[6], "rules_scanned_weighted_update": tot[7],
                "summed_rules_per_env_step": (tot[3] + tot[4]) / max(nsteps, 1),
                "nominal_rule_evaluations_per_env_step": evals * sumR / max(nsteps, 1)}
    traffic = None
    if cal:
        # executed FP64 thread instructions = per-step body + 3 per table rule (DMUL + 2 DADD) + chain constant per computed rule,
        # constants fitted to ncu's dadd/dmul/dfma thread-instruction counts of a capture of THIS kernel source
        fp64_instr = cal["fp64_per_step"] * nsteps + cal["fp64_per_table_rule"] * tot[3] + cal["fp64_per_chain_rule"] * tot[4]
        executed["fp64_thread_instr_est"] = fp64_instr
        executed["frac_executed"] = fp64_instr / world / t / (fp64_peak_g * 1e9)
        executed["calibration"] = cal.get("source")
        traffic = cal.get("dram_bytes_per_launch")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = nsteps / t
    launches_per_step = (-(-chunk // retile) if retile > 0 else 1)
    groups = args.groups or min(8, max(1, n // 32768))
    if shared:
        gpu_launches = args.steps * (1 + 1 + 1 + 1 + 1)            # new_episodes, set_s0, shared_start, infer_memo, result_stats (results are plain copies)
    else:
        gpu_launches = args.steps * ((6 if w["env"] != "cartpole" else 5) * launches_per_step * groups + 3 + 1)   # + 3 result un-permutes + result_stats
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": config["scaling"], "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": config,
        "arm": "libfriq_b200.so (hand-written sm_100a kernels behind the C-ABI), one process per GPU",
        "stats": {"mean_rules": sumR / max(nsteps, 1), "rule_capacity": w["cap"], "retile_interval_steps": retile,
                  "l2": "inputs larger than L2: rule bases + scratch = %.0f MB per GPU" % (n * w["cap"] * 20 / 1e6) if not shared else
                        "shared rule base (0.5 MB) + memo live in L2; agent state 41 MB per GPU streams from HBM",
                  "rng": "Philox4x32-10, key = 0x5EED0000 + agent id"},
        "five_evals_per_sec": value * evals,
        "e2e": {"value": nsteps_e2e / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": gpu_launches,
        "clocks": clocks,
        **({"ms_per_step_by_rank": [round(x, 4) for x in ms_ranks]} if ms_ranks else {}),
        "roofline": {"bound": "fp64" if frac_fp64 >= frac_hbm else "hbm",
                     "achieved": ach_tflops if frac_fp64 >= frac_hbm else ach_gbs,
                     "peak": fp64_peak_g / 1e3 if frac_fp64 >= frac_hbm else hbm_peak,
                     "unit": "TFLOP/s" if frac_fp64 >= frac_hbm else "GB/s",
                     "frac": max(frac_fp64, frac_hbm), "traffic": traffic, "executed": executed,
                     "kernel": w["kernel"],
                     "note": "NOMINAL: algorithmic FP64 ops of the reference path (SURVEY 8d: %g * R per env step, sqrt and / counted as 1, R "
                             "summed on the device) over the non-FMA FP64 issue rate measured in this run (friq_device_fp64_peak).  The "
                             "kernels execute far less - exact hits short-cut the sums like in the reference, table mode looks the weights "
                             "up, config 4 memoises whole conclusions - so frac can exceed 1; `executed` holds the device-side counters of "
                             "this run.  hbm view: %.1f GB/s algorithmic of %.0f GB/s measured = %.3f; time = the whole timed step "
                             "(re-tiling and result gather included)" % (w["f_rule"], ach_gbs, hbm_peak, frac_hbm)},
        "roofline_hbm": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": frac_hbm, "traffic": traffic},
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            cb = run_reference(args.workload, steps_per_agent)
            out["cpu_baseline"] = {k: v for k, v in cb.items() if k != "mean_rules"}
            out["stats"]["cpu_baseline_mean_rules"] = cb.get("mean_rules")
        except Exception as e:       # noqa: BLE001 - reported, never silently dropped
            out["cpu_baseline"] = {"error": str(e)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
