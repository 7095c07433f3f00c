"""Parser for the record streams written by oracle/ref_harness.c (see there) and
committed by oracle/make_golden.py as tests/golden/<env>_calls.npz."""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
ENVS = ["mountaincar", "acrobot", "cartpole"]
ENV_ID = {"mountaincar": 0, "acrobot": 1, "cartpole": 2}
TAG_HEADER, TAG_GBA, TAG_SARSA, TAG_EPISODE, TAG_FINAL = 1, 2, 3, 4, 5


class _Cur:
    def __init__(self, a):
        self.a, self.i = a, 0

    def f(self):
        v = self.a[self.i]; self.i += 1; return float(v)

    def n(self):
        return int(self.f())

    def v(self, n):
        out = np.array(self.a[self.i:self.i + n], dtype=np.float64); self.i += n; return out

    def rb(self, K):
        R = self.n()
        return R, self.v(R * K).reshape(R, K), self.v(R)


def load(env):
    z = np.load(os.path.join(GOLDEN, f"{env}_calls.npz"))
    s = z["stream"]
    recs = {"gba": [], "sarsa": [], "episodes": [], "md5": str(z["md5"])}
    i = 0
    K = S = A = None
    while i < s.size:
        tag, n = int(s[i]), int(s[i + 1])
        c = _Cur(s[i + 2:i + 2 + n]); i += 2 + n
        if tag == TAG_HEADER:
            K, S, A, U, p = c.n(), c.n(), c.n(), c.n(), c.n()
            h = dict(K=K, S=S, A=A, U=U, p=p)
            for name in ["alpha", "gamma", "epsilon", "qdiff_pos", "qdiff_neg", "qdiff_final_tol",
                         "reward_good_above", "wsig", "skip_rules", "no_random", "max_episodes", "max_steps"]:
                h[name] = c.f()
            h["u"] = c.v(K * U).reshape(K, U); h["ve"] = c.v(K * U).reshape(K, U); h["udiv"] = c.v(K)
            h["statedims"] = []
            for _ in range(S):
                ln = c.n()
                d = dict(len=ln, div=c.f(), steep=c.f(), default=c.f(), udiv=c.f())
                d["values"] = c.v(ln); h["statedims"].append(d)
            h["act_div"], h["act_udiv"] = c.f(), c.f()
            h["act_val"], h["act_ve"] = c.v(A), c.v(A)
            h["rb0"] = c.rb(K)
            recs["header"] = h
        elif tag == TAG_GBA:
            R, rant, rconc = c.rb(K)
            recs["gba"].append(dict(R=R, rant=rant, rconc=rconc, states=c.v(S), actconc=c.v(A), best=c.n()))
        elif tag == TAG_SARSA:
            r = dict(cat=c.n(), inserted_before=c.n())
            r["R"], r["rant"], r["rconc"] = c.rb(K)
            r["q_ant"], r["reward"], r["cur_q_ant"] = c.v(K), c.f(), c.v(K)
            r["concl"] = [dict(ant=c.v(K), conc=c.f(), ret=c.n()) for _ in range(c.n())]
            if c.n():
                w = dict(ant=c.v(K), ret=c.n()); w["weights"] = c.v(c.n()); r["weight"] = w
            else:
                r["weight"] = None
            r["R_after"], r["rant_after"], r["rconc_after"] = c.rb(K)
            r["inserted_after"] = c.n()
            recs["sarsa"].append(r)
        elif tag == TAG_EPISODE:
            recs["episodes"].append((c.n(), c.n(), c.f(), c.n()))
        elif tag == TAG_FINAL:
            recs["final"] = c.rb(K)
        assert c.i == n, (tag, c.i, n)
    return recs


def rb_text(rant, rconc):
    """frirl_save_rb_to_text_file format (frirl_utils.c:121-139): '%.18f ' per value, ' \\n' line end."""
    lines = []
    for a, q in zip(rant, rconc):
        lines.append("".join("%.18f " % v for v in a) + "%.18f \n" % q)
    return "".join(lines).encode()


def golden_text(env):
    return open(os.path.join(GOLDEN, f"{env}.frirlrb.txt"), "rb").read()


def same_episode_rows(want, got):
    """rows (steps, return, rules): the golden returns were parsed from the reference's "%f" printout (6 decimals)"""
    if len(want) != len(got):
        return False
    return all(int(w[0]) == int(g[0]) and ("%f" % w[1]) == ("%f" % g[1]) and int(w[2]) == int(g[2]) for w, g in zip(want, got))
