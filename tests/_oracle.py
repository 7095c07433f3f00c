"""ctypes binding of oracle/libfriq_oracle.so -- TEST INFRASTRUCTURE (the checker)."""
import ctypes as C
import os
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None
dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(REPO, "oracle", "libfriq_oracle.so"))
        L.fo_problem_builtin.restype = C.c_void_p; L.fo_problem_builtin.argtypes = [C.c_int, C.c_int]
        L.fo_problem_builtin_actions.restype = C.c_void_p; L.fo_problem_builtin_actions.argtypes = [C.c_int, C.c_int, C.c_int, dp, C.c_double]
        L.fo_problem_free.argtypes = [C.c_void_p]
        for f in ["fo_problem_u", "fo_problem_ve", "fo_problem_udiv", "fo_problem_act_val", "fo_problem_act_ve"]:
            getattr(L, f).restype = dp; getattr(L, f).argtypes = [C.c_void_p]
        L.fo_problem_grid.restype = dp; L.fo_problem_grid.argtypes = [C.c_void_p, C.c_int]
        L.fo_problem_grid_len.argtypes = [C.c_void_p, C.c_int]
        L.fo_problem_info.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.fo_problem_hyper.argtypes = [C.c_void_p, dp]
        L.fo_problem_set_random.argtypes = [C.c_void_p, C.c_int]
        L.fo_agent_new.restype = C.c_void_p; L.fo_agent_new.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.fo_agent_free.argtypes = [C.c_void_p]
        L.fo_agent_set_rb.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.fo_agent_R.argtypes = [C.c_void_p]
        L.fo_agent_rant.restype = dp; L.fo_agent_rant.argtypes = [C.c_void_p]
        L.fo_agent_rconc.restype = dp; L.fo_agent_rconc.argtypes = [C.c_void_p]
        L.fo_agent_inserted.argtypes = [C.c_void_p]
        L.fo_agent_set_inserted.argtypes = [C.c_void_p, C.c_int]
        L.fo_agent_set_hyper.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.fo_agent_set_learning.argtypes = [C.c_void_p, C.c_int]
        L.fo_agent_summary.argtypes = [C.c_void_p, dp]
        L.fo_agent_set_initial_state.argtypes = [C.c_void_p, dp]
        L.fo_vag_concl.argtypes = [C.c_void_p, dp, dp]
        L.fo_vag_concl_weight.argtypes = [C.c_void_p, dp, dp]
        L.fo_get_best_action.argtypes = [C.c_void_p, dp, dp]
        L.fo_update_sarsa.argtypes = [C.c_void_p, dp, C.c_double, dp]
        L.fo_agent_run_steps.argtypes = [C.c_void_p, C.c_long, C.c_int]
        L.fo_run_construct.argtypes = [C.c_void_p, dp, C.c_int]
        L.fo_philox4x32_10.argtypes = [C.POINTER(C.c_uint32)] * 3
        L.fo_pow_x87.restype = C.c_double; L.fo_pow_x87.argtypes = [C.c_double, C.c_int]
        L.fo_snap.restype = C.c_double; L.fo_snap.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.fo_sin.restype = C.c_double; L.fo_sin.argtypes = [C.c_void_p, C.c_double]
        L.fo_cos.restype = C.c_double; L.fo_cos.argtypes = [C.c_void_p, C.c_double]
        L.fo_env_step.argtypes = [C.c_void_p, C.c_double, dp, dp]
        L.fo_env_reward.argtypes = [C.c_void_p, dp, dp, C.POINTER(C.c_int)]
        L.fo_env_quantize.argtypes = [C.c_void_p, dp, dp]
        L.fo_merge_rb.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.fo_team_run.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, dp, C.c_int]
        L.fo_team_run_mpi.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, dp, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(dp)


class Problem:
    def __init__(self, env, trig=1, actions=None, act_div=0.0):
        """actions: None = the example's own action set; an int A = A generated values with step act_div; an array = those values"""
        if actions is None:
            self.h = lib().fo_problem_builtin(env, trig)
        elif isinstance(actions, int):
            self.h = lib().fo_problem_builtin_actions(env, trig, actions, None, act_div)
        else:
            self._av = np.ascontiguousarray(actions, dtype=np.float64)
            self.h = lib().fo_problem_builtin_actions(env, trig, len(self._av), _p(self._av), act_div)
        info = (C.c_int * 8)(); lib().fo_problem_info(self.h, info)
        self.env, self.S, self.K, self.A, self.U, self.p, self.max_steps, self.max_episodes = list(info)
        hy = np.zeros(32); lib().fo_problem_hyper(self.h, _p(hy)); self.hyper = hy

    def arr(self, name, n):
        return np.ctypeslib.as_array(getattr(lib(), "fo_problem_" + name)(self.h), shape=(n,)).copy()

    def grid(self, k):
        n = lib().fo_problem_grid_len(self.h, k)
        return np.ctypeslib.as_array(lib().fo_problem_grid(self.h, k), shape=(n,)).copy()

    def set_no_random(self, v):
        lib().fo_problem_set_random(self.h, int(v))


class Agent:
    def __init__(self, prob, capacity=4096, seed=0x5EED0000):
        self.P = prob
        self.h = lib().fo_agent_new(prob.h, capacity, seed)

    def __del__(self):
        try:
            lib().fo_agent_free(self.h)
        except Exception:
            pass

    def set_rb(self, rant, rconc):
        rant = np.ascontiguousarray(rant, dtype=np.float64); rconc = np.ascontiguousarray(rconc, dtype=np.float64)
        lib().fo_agent_set_rb(self.h, len(rconc), _p(rant), _p(rconc))

    def rb(self):
        R = lib().fo_agent_R(self.h)
        rant = np.ctypeslib.as_array(lib().fo_agent_rant(self.h), shape=(R * self.P.K,)).copy().reshape(R, self.P.K)
        rconc = np.ctypeslib.as_array(lib().fo_agent_rconc(self.h), shape=(R,)).copy()
        return R, rant, rconc

    def vag_concl(self, ant):
        ant = np.ascontiguousarray(ant, dtype=np.float64); c = C.c_double()
        hit = lib().fo_vag_concl(self.h, _p(ant), C.byref(c))
        return hit, c.value

    def vag_concl_weight(self, ant):
        ant = np.ascontiguousarray(ant, dtype=np.float64); w = np.zeros(lib().fo_agent_R(self.h))
        hit = lib().fo_vag_concl_weight(self.h, _p(ant), _p(w))
        return hit, w

    def best_action(self, states):
        states = np.ascontiguousarray(states, dtype=np.float64); ac = np.zeros(self.P.A)
        b = lib().fo_get_best_action(self.h, _p(states), _p(ac))
        return b, ac

    def update_sarsa(self, q_ant, reward, cur_q_ant):
        q_ant = np.ascontiguousarray(q_ant, dtype=np.float64); cur = np.ascontiguousarray(cur_q_ant, dtype=np.float64)
        lib().fo_update_sarsa(self.h, _p(q_ant), float(reward), _p(cur))

    def set_inserted(self, v):
        lib().fo_agent_set_inserted(self.h, int(v))

    def inserted(self):
        return lib().fo_agent_inserted(self.h)

    def set_hyper(self, alpha, gamma, epsilon):
        lib().fo_agent_set_hyper(self.h, alpha, gamma, epsilon)

    def set_learning(self, v):
        lib().fo_agent_set_learning(self.h, int(v))

    def set_initial_state(self, s0):
        s0 = np.ascontiguousarray(s0, dtype=np.float64)
        lib().fo_agent_set_initial_state(self.h, _p(s0))

    def merge_rb(self, rant, rconc):
        """merge_rb (frirl_agent.c:58-117): fold the sender's rules into this agent's rule base"""
        rant = np.ascontiguousarray(rant, dtype=np.float64); rconc = np.ascontiguousarray(rconc, dtype=np.float64)
        lib().fo_merge_rb(self.h, len(rconc), _p(rant), _p(rconc))

    def run_steps(self, n, stop_on_converge=0):
        lib().fo_agent_run_steps(self.h, n, stop_on_converge)

    def run_construct(self, max_log=1000):
        log = np.zeros(4 * max_log)
        n = lib().fo_run_construct(self.h, _p(log), max_log)
        return log[:4 * n].reshape(n, 4)

    def summary(self):
        o = np.zeros(32); lib().fo_agent_summary(self.h, _p(o))
        K = self.P.K
        return dict(R=int(o[0]), episode_num=int(o[1]), last_return=o[2], last_steps=int(o[3]), last_rules=int(o[4]),
                    converged=int(o[5]), total_steps=int(o[6]), sum_rules=o[7], inserted=int(o[8]), overflow=int(o[9]),
                    ep_return=o[10], ep_steps=int(o[11]), need_reset=int(o[12]), step_in_ep=int(o[13]),
                    s=o[16:16 + K - 1].copy(), qant=o[24:24 + K].copy())


def team_run(agents, max_rounds=200, max_log=20000, mpi_order=False):
    """frirl_omp_run (frirl_agent.c:294-385) over the given agents - or, mpi_order, the exchange order of frirl_mpi_run (:387-467);
    rows (agent, steps, return, rules) per finished episode"""
    hs = (C.c_void_p * len(agents))(*[a.h for a in agents])
    log = np.zeros(4 * max_log)
    n = (lib().fo_team_run_mpi if mpi_order else lib().fo_team_run)(hs, len(agents), max_rounds, _p(log), max_log)
    return log[:4 * n].reshape(n, 4)
