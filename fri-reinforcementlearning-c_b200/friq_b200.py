"""ctypes binding of libfriq_b200.so (the C-ABI declared in include/friq_b200.h).

Thin by design: numpy arrays in, numpy arrays out, every call goes straight to the C-ABI,
which launches the sm_100a kernels.  There is no Python or CPU implementation of the path
here; if the shared library is missing, importing the engine fails loudly.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FRIQ_LIB") or os.path.join(HERE, "libfriq_b200.so")   # FRIQ_LIB: build variants in experiments

ENV_MOUNTAINCAR, ENV_ACROBOT, ENV_CARTPOLE = 0, 1, 2
ENV_ID = {"mountaincar": 0, "acrobot": 1, "cartpole": 2}
RUN_LEARN, RUN_TRACK, RUN_STOP_ON_CONVERGE, RUN_STOP_AT_EPISODE_END, RUN_RETILE, RUN_KEEP_WEIGHTS = 1, 2, 4, 8, 16, 32
F_NEED_RESET, F_INSERTED, F_CONVERGED, F_OVERFLOW, F_EPENDED, F_TEAM_DONE = 1, 2, 4, 8, 16, 32

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)
llp = C.POINTER(C.c_longlong)


class DimDesc(C.Structure):          # friq_dim_desc == struct frirl_dimension_desc (frirl_types.h:34-44)
    _fields_ = [("values_len", C.c_int), ("values", dp), ("values_div", C.c_double), ("values_steep", C.c_double),
                ("values_def", C.c_double), ("universe_len", C.c_int), ("universe", dp), ("universe_div", C.c_double)]


class ProblemDesc(C.Structure):      # friq_problem_desc
    _fields_ = [("env", C.c_int), ("statedims_len", C.c_int), ("statedims", C.POINTER(DimDesc)), ("actiondim", DimDesc),
                ("alpha", C.c_double), ("gamma", C.c_double), ("epsilon", C.c_double),
                ("qdiff_pos_boundary", C.c_double), ("qdiff_neg_boundary", C.c_double), ("qdiff_final_tolerance", C.c_double),
                ("reward_good_above", C.c_double), ("rule_weight_considered_significant_for_update", C.c_double),
                ("skip_rules", C.c_int), ("no_random", C.c_int), ("max_episodes", C.c_int), ("max_steps", C.c_int)]


SYMBOLS = {
    "friq_last_error": (C.c_char_p, []),
    "friq_device_count": (C.c_int, [ip]),
    "friq_problem_desc_builtin": (C.c_int, [C.c_int, C.POINTER(ProblemDesc)]),
    "friq_problem_create": (C.c_int, [C.POINTER(ProblemDesc), C.POINTER(C.c_void_p)]),
    "friq_problem_create_from_tables": (C.c_int, [C.c_int, C.c_int, C.c_int, dp, dp, C.POINTER(C.c_void_p)]),
    "friq_problem_set_learning_params": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]),
    "friq_problem_destroy": (None, [C.c_void_p]),
    "friq_problem_dims": (C.c_int, [C.c_void_p, ip]),
    "friq_problem_tables": (C.c_int, [C.c_void_p, dp, dp, dp, dp, dp]),
    "friq_batch_create": (C.c_int, [C.c_void_p, C.c_int, C.c_long, C.c_int, C.c_uint64, C.c_long, C.POINTER(C.c_void_p)]),
    "friq_batch_destroy": (None, [C.c_void_p]),
    "friq_batch_reset": (C.c_int, [C.c_void_p]),
    "friq_batch_set_hyper": (C.c_int, [C.c_void_p, dp, dp, dp]),
    "friq_batch_set_rulebase": (C.c_int, [C.c_void_p, C.c_long, C.c_int, dp, dp]),
    "friq_batch_get_rulebase": (C.c_int, [C.c_void_p, C.c_long, ip, dp, dp]),
    "friq_batch_set_shared_rulebase": (C.c_int, [C.c_void_p, C.c_int, dp, dp]),
    "friq_batch_set_initial_state": (C.c_int, [C.c_void_p, dp]),
    "friq_batch_set_initial_state_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "friq_batch_get_state": (C.c_int, [C.c_void_p, dp, ip]),
    "friq_batch_set_inserted": (C.c_int, [C.c_void_p, ip]),
    "friq_batch_get_inserted": (C.c_int, [C.c_void_p, ip]),
    "friq_batch_run": (C.c_int, [C.c_void_p, C.c_long, C.c_int, C.c_void_p]),
    "friq_batch_sync": (C.c_int, [C.c_void_p, C.c_void_p]),
    "friq_batch_set_retile_interval": (C.c_int, [C.c_void_p, C.c_long]),
    "friq_batch_set_policy": (C.c_int, [C.c_void_p, C.c_int]),
    "friq_batch_set_groups": (C.c_int, [C.c_void_p, C.c_int]),
    "friq_batch_gather": (C.c_int, [C.c_void_p, dp, ip, ip, ip, ip, llp, llp]),
    "friq_batch_new_episodes": (C.c_int, [C.c_void_p, C.c_void_p]),
    "friq_result_stats": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]),
    "friq_batch_work_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]),
    "friq_batch_gather_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "friq_batch_run_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "friq_batch_run_host_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "friq_batch_vag_concl": (C.c_int, [C.c_void_p, dp, dp, ip]),
    "friq_batch_vag_concl_weight": (C.c_int, [C.c_void_p, dp, ip, dp]),
    "friq_batch_get_best_action": (C.c_int, [C.c_void_p, dp, dp, ip]),
    "friq_batch_update_sarsa": (C.c_int, [C.c_void_p, dp, dp, dp]),
    "friq_batch_classic_step": (C.c_int, [C.c_void_p, dp, dp, dp, ip, C.c_int, ip, dp, ip, ip]),
    "friq_batch_rule_distance": (C.c_int, [C.c_void_p, dp, dp, ip]),
    "friq_batch_bestact_from_dists": (C.c_int, [C.c_void_p, dp, dp]),
    "friq_batch_append_rule": (C.c_int, [C.c_void_p, C.c_long, dp, C.c_double]),
    "friq_batch_remove_rule": (C.c_int, [C.c_void_p, C.c_long, C.c_int]),
    "friq_batch_keep_weights": (C.c_int, [C.c_void_p]),
    "friq_batch_merge_rb": (C.c_int, [C.c_void_p, C.c_long, C.c_int, dp, dp]),
    "friq_team_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int, dp, C.c_long, C.POINTER(C.c_long), ip]),
    "friq_batch_save_rb_files": (C.c_int, [C.c_void_p, C.c_long, C.c_long, C.c_char_p, C.c_int]),
    "friq_batch_load_rb_file": (C.c_int, [C.c_void_p, C.c_long, C.c_char_p]),
    "friq_batch_save_checkpoint": (C.c_int, [C.c_void_p, C.c_char_p]),
    "friq_batch_load_checkpoint": (C.c_int, [C.c_void_p, C.c_char_p]),
    "friq_multi_create": (C.c_int, [C.POINTER(ProblemDesc), C.c_int, C.c_long, C.c_int, C.c_uint64, C.POINTER(C.c_void_p)]),
    "friq_multi_destroy": (None, [C.c_void_p]),
    "friq_multi_devices": (C.c_int, [C.c_void_p]),
    "friq_multi_batch": (C.c_void_p, [C.c_void_p, C.c_int]),
    "friq_multi_first_agent": (C.c_long, [C.c_void_p, C.c_int]),
    "friq_multi_agents": (C.c_long, [C.c_void_p, C.c_int]),
    "friq_multi_set_hyper": (C.c_int, [C.c_void_p, dp, dp, dp]),
    "friq_multi_run": (C.c_int, [C.c_void_p, C.c_long, C.c_int]),
    "friq_multi_sync": (C.c_int, [C.c_void_p]),
    "friq_multi_gather": (C.c_int, [C.c_void_p, dp, ip, ip, dp]),
    "friq_multi_gathered_device": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "friq_multi_merge_from": (C.c_int, [C.c_void_p, C.c_long, C.c_long]),
    "friq_reduce_rulebase": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, ip, dp, dp, C.c_int, ip]),
    "friq_save_rb_to_text_file": (C.c_int, [C.c_char_p, C.c_int, C.c_int, dp, dp]),
    "friq_save_rb_to_bin_file": (C.c_int, [C.c_char_p, C.c_int, C.c_int, dp, dp]),
    "friq_load_rb_from_bin_file": (C.c_int, [C.c_char_p, C.c_int, C.c_int, ip, dp, dp]),
    "friq_device_trig": (C.c_int, [dp, dp, dp, C.c_long]),
    "friq_device_pow_x87": (C.c_int, [dp, dp, C.c_long, C.c_int, C.c_int]),
    "friq_device_fp64_peak": (C.c_int, [dp]),
}

_lib = None


def build():
    """Compile the extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C {HERE}` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


class FriqError(RuntimeError):
    pass


def _ck(rc):
    if rc != 0:
        raise FriqError(f"[{rc}] {lib().friq_last_error().decode()}")


def _p(a, t=dp):
    return None if a is None else a.ctypes.data_as(t)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def device_count():
    n = C.c_int(0)
    _ck(lib().friq_device_count(C.byref(n)))
    return n.value


class Problem:
    """Host tables of one FRIQ problem (frirl_init equivalent)."""

    def __init__(self, env=None, desc=None, **overrides):
        L = lib()
        self.desc = ProblemDesc()
        if desc is not None:
            self.desc = desc
        else:
            _ck(L.friq_problem_desc_builtin(ENV_ID[env] if isinstance(env, str) else env, C.byref(self.desc)))
        for k, v in overrides.items():
            setattr(self.desc, k, v)
        self.h = C.c_void_p()
        _ck(L.friq_problem_create(C.byref(self.desc), C.byref(self.h)))
        d = (C.c_int * 6)()
        L.friq_problem_dims(self.h, d)
        self.S, self.K, self.A, self.U, self.p, self.ncorner = list(d)

    def tables(self):
        u = np.zeros((self.K, self.U)); ve = np.zeros((self.K, self.U)); udiv = np.zeros(self.K)
        av = np.zeros(self.A); ave = np.zeros(self.A)
        _ck(lib().friq_problem_tables(self.h, _p(u), _p(ve), _p(udiv), _p(av), _p(ave)))
        return dict(u=u, ve=ve, udiv=udiv, act_val=av, act_ve=ave)

    def close(self):
        if self.h:
            lib().friq_problem_destroy(self.h); self.h = None


class Batch:
    """n independent agents with private rule bases, resident on one GPU."""

    def __init__(self, problem, n_agents, capacity, device=0, seed0=0x5EED0000, first_agent=0):
        self.P, self.n, self.cap = problem, int(n_agents), int(capacity)
        self.h = C.c_void_p()
        _ck(lib().friq_batch_create(problem.h, device, self.n, self.cap, seed0, first_agent, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().friq_batch_destroy(self.h); self.h = None

    def reset(self):
        _ck(lib().friq_batch_reset(self.h))

    def set_hyper(self, alpha=None, gamma=None, epsilon=None):
        a, g, e = [None if x is None else _f64(np.broadcast_to(x, (self.n,))) for x in (alpha, gamma, epsilon)]
        _ck(lib().friq_batch_set_hyper(self.h, _p(a), _p(g), _p(e)))

    def set_rulebase(self, rant, rconc, agent=-1):
        rant, rconc = _f64(rant), _f64(rconc)
        _ck(lib().friq_batch_set_rulebase(self.h, agent, len(rconc), _p(rant), _p(rconc)))

    def set_shared_rulebase(self, rant, rconc):
        rant, rconc = _f64(rant), _f64(rconc)
        _ck(lib().friq_batch_set_shared_rulebase(self.h, len(rconc), _p(rant), _p(rconc)))

    def set_initial_state(self, s0):
        s0 = None if s0 is None else _f64(s0).reshape(self.n, self.P.S)
        _ck(lib().friq_batch_set_initial_state(self.h, _p(s0)))

    def set_initial_state_device(self, ptr, stream=None):
        _ck(lib().friq_batch_set_initial_state_device(self.h, ptr, stream))

    def get_rulebase(self, agent=0):
        R = C.c_int(0)
        rant = np.zeros((self.cap, self.P.K)); rconc = np.zeros(self.cap)
        _ck(lib().friq_batch_get_rulebase(self.h, agent, C.byref(R), _p(rant), _p(rconc)))
        return R.value, rant[:R.value].copy(), rconc[:R.value].copy()

    def save_rb_files(self, pattern, first=0, count=None, text=False):
        _ck(lib().friq_batch_save_rb_files(self.h, first, self.n - first if count is None else count, pattern.encode(), int(text)))

    def load_rb_file(self, path, agent=-1):
        _ck(lib().friq_batch_load_rb_file(self.h, agent, path.encode()))

    def save_checkpoint(self, path):
        _ck(lib().friq_batch_save_checkpoint(self.h, path.encode()))

    def load_checkpoint(self, path):
        _ck(lib().friq_batch_load_checkpoint(self.h, path.encode()))

    def get_state(self):
        s = np.zeros((self.n, self.P.S)); a = np.zeros(self.n, np.int32)
        _ck(lib().friq_batch_get_state(self.h, _p(s), _p(a, ip)))
        return s, a

    def set_inserted(self, v):
        v = np.ascontiguousarray(np.broadcast_to(v, (self.n,)), dtype=np.int32)
        _ck(lib().friq_batch_set_inserted(self.h, _p(v, ip)))

    def get_inserted(self):
        v = np.zeros(self.n, dtype=np.int32)
        _ck(lib().friq_batch_get_inserted(self.h, _p(v, ip)))
        return v

    def run(self, n_steps, mode=RUN_LEARN, stream=None, sync=True):
        _ck(lib().friq_batch_run(self.h, int(n_steps), int(mode), stream))
        if sync:
            _ck(lib().friq_batch_sync(self.h, stream))

    def set_retile_interval(self, steps):
        _ck(lib().friq_batch_set_retile_interval(self.h, int(steps)))

    def set_policy(self, policy):
        _ck(lib().friq_batch_set_policy(self.h, int(policy)))

    def set_groups(self, groups):
        _ck(lib().friq_batch_set_groups(self.h, int(groups)))

    def sync(self, stream=None):
        _ck(lib().friq_batch_sync(self.h, stream))

    def gather(self):
        n = self.n
        out = dict(last_return=np.zeros(n), last_steps=np.zeros(n, np.int32), rules=np.zeros(n, np.int32),
                   episodes=np.zeros(n, np.int32), flags=np.zeros(n, np.int32),
                   total_steps=np.zeros(n, np.int64), sum_rules=np.zeros(n, np.int64))
        _ck(lib().friq_batch_gather(self.h, _p(out["last_return"]), _p(out["last_steps"], ip), _p(out["rules"], ip),
                                    _p(out["episodes"], ip), _p(out["flags"], ip), _p(out["total_steps"], llp),
                                    _p(out["sum_rules"], llp)))
        return out

    def new_episodes(self, stream=None):
        _ck(lib().friq_batch_new_episodes(self.h, stream))

    def work_counters(self, reset=False):
        v = (C.c_ulonglong * 8)()
        _ck(lib().friq_batch_work_counters(self.h, v, int(reset)))
        return dict(zip(("pairs_summed", "table_rules", "chain_rules", "hit_pairs", "appended", "wupd_rules", "steps"), [int(x) for x in v][:7]))

    def gather_device(self, ret_ptr, steps_ptr, rules_ptr, stream=None):
        _ck(lib().friq_batch_gather_device(self.h, ret_ptr, steps_ptr, rules_ptr, stream))

    def run_host(self, n_steps, mode, alpha_ptr, gamma_ptr, eps_ptr, ret_ptr, steps_ptr, rules_ptr, stream=None):
        """End-to-end call with HOST pointers (e.g. pinned torch tensors): H2D, run, D2H, sync."""
        _ck(lib().friq_batch_run_host(self.h, alpha_ptr, gamma_ptr, eps_ptr, int(n_steps), int(mode), ret_ptr, steps_ptr, rules_ptr, stream))

    def run_host_async(self, n_steps, mode, alpha_ptr, gamma_ptr, eps_ptr, ret_ptr, steps_ptr, rules_ptr, stream=None):
        """The same, enqueued only: the host buffers are valid once `stream` has passed this point (record an event after the call)."""
        _ck(lib().friq_batch_run_host_async(self.h, alpha_ptr, gamma_ptr, eps_ptr, int(n_steps), int(mode), ret_ptr, steps_ptr, rules_ptr, stream))

    def merge_rb(self, rant, rconc, agent=-1):
        """merge_rb (frirl_agent.c:58-117): fold a sender's rule list into one agent's (agent < 0: every agent's) rule base"""
        rant, rconc = _f64(rant), _f64(rconc)
        _ck(lib().friq_batch_merge_rb(self.h, agent, len(rconc), _p(rant), _p(rconc)))

    def team_run(self, team, max_rounds=200, max_log=100000):
        """frirl_omp_run (frirl_agent.c:294-385) for teams of `team` consecutive agents; returns (log rows, rounds)"""
        log = np.zeros((max_log, 4)); n = C.c_long(0); rounds = C.c_int(0)
        _ck(lib().friq_team_run(self.h, int(team), int(max_rounds), _p(log), max_log, C.byref(n), C.byref(rounds)))
        return log[:min(n.value, max_log)].copy(), rounds.value

    # ---- per-call entry points
    def vag_concl(self, ant):
        ant = _f64(ant).reshape(self.n, self.P.K)
        conc = np.zeros(self.n); hit = np.zeros(self.n, np.int32)
        _ck(lib().friq_batch_vag_concl(self.h, _p(ant), _p(conc), _p(hit, ip)))
        return hit, conc

    def vag_concl_weight(self, ant):
        ant = _f64(ant).reshape(self.n, self.P.K)
        hit = np.zeros(self.n, np.int32); w = np.zeros((self.n, self.cap))
        _ck(lib().friq_batch_vag_concl_weight(self.h, _p(ant), _p(hit, ip), _p(w)))
        return hit, w

    def best_action(self, states):
        states = _f64(states).reshape(self.n, self.P.S)
        ac = np.zeros((self.n, self.P.A)); best = np.zeros(self.n, np.int32)
        _ck(lib().friq_batch_get_best_action(self.h, _p(states), _p(ac), _p(best, ip)))
        return best, ac

    def update_sarsa(self, q_ant, reward, cur_q_ant):
        q_ant = _f64(q_ant).reshape(self.n, self.P.K); cur = _f64(cur_q_ant).reshape(self.n, self.P.K)
        reward = _f64(np.broadcast_to(reward, (self.n,)))
        _ck(lib().friq_batch_update_sarsa(self.h, _p(q_ant), _p(reward), _p(cur)))


class Multi:
    """Every GPU of the node behind one handle (friq_multi_*): contiguous agent blocks, NCCL all-gather of the results."""

    def __init__(self, env, n_agents_total, capacity, n_devices=0, seed0=0x5EED0000, **overrides):
        L = lib()
        self.desc = ProblemDesc()
        _ck(L.friq_problem_desc_builtin(ENV_ID[env] if isinstance(env, str) else env, C.byref(self.desc)))
        for k, v in overrides.items():
            setattr(self.desc, k, v)
        self.h = C.c_void_p()
        self.n, self.cap = int(n_agents_total), int(capacity)
        _ck(L.friq_multi_create(C.byref(self.desc), n_devices, self.n, capacity, seed0, C.byref(self.h)))
        self.G = L.friq_multi_devices(self.h)

    def batch(self, gpu, problem_like):
        """A non-owning Batch view of GPU `gpu`'s batch (problem_like supplies S / K / A for array shapes)."""
        b = Batch.__new__(Batch)
        b.P, b.n, b.cap = problem_like, int(lib().friq_multi_agents(self.h, gpu)), self.cap
        b.h = C.c_void_p(lib().friq_multi_batch(self.h, gpu))
        b.close = lambda: None
        return b

    def first_agent(self, gpu):
        return int(lib().friq_multi_first_agent(self.h, gpu))

    def set_hyper(self, alpha=None, gamma=None, epsilon=None):
        a, g, e = [None if x is None else _f64(np.broadcast_to(x, (self.n,))) for x in (alpha, gamma, epsilon)]
        _ck(lib().friq_multi_set_hyper(self.h, _p(a), _p(g), _p(e)))

    def run(self, n_steps, mode=RUN_LEARN, sync=True):
        _ck(lib().friq_multi_run(self.h, int(n_steps), int(mode)))
        if sync:
            _ck(lib().friq_multi_sync(self.h))

    def gather(self):
        ret = np.zeros(self.n); steps = np.zeros(self.n, np.int32); rules = np.zeros(self.n, np.int32); stats = np.zeros(3)
        _ck(lib().friq_multi_gather(self.h, _p(ret), _p(steps, ip), _p(rules, ip), _p(stats)))
        return dict(last_return=ret, last_steps=steps, rules=rules, stats=stats)

    def merge_from(self, src_agent, dst_agent=-1):
        _ck(lib().friq_multi_merge_from(self.h, int(src_agent), int(dst_agent)))

    def close(self):
        if self.h:
            lib().friq_multi_destroy(self.h); self.h = None


def result_stats(device, ret_ptr, rules_ptr, n, out_ptr, stream=None):
    _ck(lib().friq_result_stats(int(device), ret_ptr, rules_ptr, int(n), out_ptr, stream))


def device_trig(x):
    x = _f64(x); s = np.zeros_like(x); c = np.zeros_like(x)
    _ck(lib().friq_device_trig(_p(x), _p(s), _p(c), x.size))
    return s, c


def device_pow_x87(x, p, force_exact=0):
    """force_exact: 0 filter+fallback pow, 1 integer pow, 2 weight via library ops, 3 weight via fast paths, 4 fallback flags"""
    x = _f64(x); out = np.zeros_like(x)
    _ck(lib().friq_device_pow_x87(_p(x), _p(out), x.size, p, int(force_exact)))
    return out


def device_fp64_peak():
    g = C.c_double(0)
    _ck(lib().friq_device_fp64_peak(C.byref(g)))
    return g.value


def reduce_rulebase(problem, rant, rconc, strategy=1, reward_tolerance=0.0, width=32, device=0, verbose=0):
    """friq_reduce_rulebase: returns (rant, rconc, episodes_run) of the reduced rule base."""
    rant = _f64(rant).copy(); rconc = _f64(rconc).copy()
    R = C.c_int(len(rconc)); eps = C.c_int(0)
    _ck(lib().friq_reduce_rulebase(problem.h, device, strategy, reward_tolerance, width, C.byref(R), _p(rant), _p(rconc), verbose, C.byref(eps)))
    return rant[:R.value].copy(), rconc[:R.value].copy(), eps.value


def save_rb_text(path, rant, rconc):
    rant, rconc = _f64(rant), _f64(rconc)
    _ck(lib().friq_save_rb_to_text_file(path.encode(), rant.shape[1], len(rconc), _p(rant), _p(rconc)))


def save_rb_bin(path, rant, rconc):
    rant, rconc = _f64(rant), _f64(rconc)
    _ck(lib().friq_save_rb_to_bin_file(path.encode(), rant.shape[1], len(rconc), _p(rant), _p(rconc)))


def load_rb_bin(path, K, max_rules=16384):
    R = C.c_int(0); rant = np.zeros((max_rules, K)); rconc = np.zeros(max_rules)
    _ck(lib().friq_load_rb_from_bin_file(path.encode(), K, max_rules, C.byref(R), _p(rant), _p(rconc)))
    return rant[:R.value].copy(), rconc[:R.value].copy()
