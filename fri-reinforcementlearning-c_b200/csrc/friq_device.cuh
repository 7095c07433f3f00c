// csrc/friq_device.cuh -- device functions of the FRIQ-learning hot path (sm_100a).
//
// One THREAD owns one agent (its state lives in registers for the whole launch) and walks
// its private rule base in rule order, so every sum is accumulated in exactly the order
// the reference uses (sequential over rules i = 0..R-1, left-to-right over dimensions) and
// the results are bit-identical to the reference's C/AVX2 code - no tree reductions.
// The 32 agents of a warp form a rule-base tile laid out [slot][lane], so the warp's loads
// of slot i are one coalesced 128 B (packed antecedents) + 256 B (consequents) transaction.
// All FP64 arithmetic uses the _rn intrinsics: no FMA contraction (the reference is built
// without -mfma; its AVX2 blocks use vmulpd + vaddpd).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "friq_devprob.h"
#include "friq_math.cuh"
#include "friq_trig.h"

#define FRIQ_PI 3.14159265358979323846264338327

// ---------------------------------------------------------------------------------------
// TMA staging of the problem block: one cp.async.bulk (UBLKCP) per CTA, completion on an
// mbarrier.  Called by every thread of the CTA; returns when the bytes have landed.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void friq_stage_problem(const friq_devprob *gprob, void *smem_dst, uint64_t *mbar)
{
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(FRIQ_DEVPROB_SMEM_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(gprob), "r"(FRIQ_DEVPROB_SMEM_BYTES), "r"(bar) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar) : "memory");
    }
}

// ---------------------------------------------------------------------------------------
// nearest universe index, src/inl/min.inl:71-92 (ties -> lower index).  `len` is U except
// for the action VE values of frirl_init.c:157, which pass U-1.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int friq_nearest_index(const friq_devprob &pb, int k, double x, int len)
{
    const double *u = pb.u + (size_t)k * pb.U;
    const int low = (int)__ddiv_rn(__dsub_rn(x, pb.u0[k]), pb.udiv[k]);
    if (low < 0) return 0;
    if (low >= len) return len - 1;
    if (low + 1 >= pb.U) return low;
    const double d1 = fabs(__dsub_rn(__ldg(u + low), x)), d2 = fabs(__dsub_rn(__ldg(u + low + 1), x));
    return d1 <= d2 ? low : low + 1;
}
__device__ __forceinline__ double friq_obs_ve(const friq_devprob &pb, int k, double x)
{
    return __ldg(pb.ve + (size_t)k * pb.U + friq_nearest_index(pb, k, x, pb.U));
}

// squared scaled distance over the first N dimensions of the rule packed in w from the point
// with per-dimension VE values vo[] (five_rule_distance.c:82-100 + :160-212, A1/A2)
template <int N>
__device__ __forceinline__ double friq_dist2(const friq_devprob &pb, uint32_t w, const double *vo)
{
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) {
        const double v = pb.dict_ve[k][FRIQ_FIELD(w, k, N)];
        const double d = __dsub_rn(vo[k], v);
        const double d2 = __dmul_rn(d, d);
        s = (k == 0) ? d2 : __dadd_rn(s, d2);
    }
    return s;
}

template <int P>
__device__ __forceinline__ double friq_shepard_weight(double s)
{
    const double d = __dsqrt_rn(s);                       // vsqrtpd, five_rule_distance.c:205
    return __drcp_rn(friq_pow_x87<P>(d));                 // 1.0 / fast_pow(|d|, p)
}
// the same as a real call: the fall-back of the fused kernels' fast paths is rare, and inlining the library sqrt / rcp
// and the power filter once per chain made those kernels ~100 KB of code (instruction-cache misses in the hot loop)
template <int P>
__device__ __noinline__ double friq_shepard_weight_slow(double s)
{
    return friq_shepard_weight<P>(s);
}
// fall-back of a chain whose square root d is known to be right (SAFE problems: only the power's acceptance test can fail)
template <int P>
__device__ __noinline__ double friq_shepard_weight_from_sqrt(double d)
{
    return __drcp_rn(friq_pow_x87_exact(d, P));
}
// same value, branch-free; bad != 0 => not proven, recompute with friq_shepard_weight
template <int P, bool SAFE>
__device__ __forceinline__ double friq_shepard_weight_fast(double s, unsigned &bad, bool live = true)
{
    const double d = friq_sqrt_fast<SAFE>(s, bad, live);
    return friq_rcp_fast<SAFE>(friq_pow_x87_fast<P>(d, bad, live), bad, live);
}

// first rule whose distance from vo[] (all K dims) is exactly 0, else -1
// (five_rule_distance.c:206-262 early exit + FIVEVagConcl.c:96-105 rescan)
template <int K>
__device__ __forceinline__ int friq_first_hit(const friq_devprob &pb, const uint32_t *rbw, int R, const double *vo)
{
    for (int i = 0; i < R; i++)
        if (friq_dist2<K>(pb, rbw[(size_t)i * FRIQ_TILE], vo) == 0.0) return i;
    return -1;
}

// sequential Shepard sums over all rules (FIVEVagConcl.c:224-235, FIVEVagConclWeight.c:125-136);
// optionally keeps wi in sw[] for the weighted update.
template <int K, int P>
__device__ __forceinline__ void friq_shepard(const friq_devprob &pb, const uint32_t *rbw, const double *rbq, double *sw, int R,
                                             const double *vo, double &vagc, double &ws)
{
    vagc = 0.0; ws = 0.0;
    for (int i = 0; i < R; i++) {
        const double s = friq_dist2<K>(pb, rbw[(size_t)i * FRIQ_TILE], vo);
        const double wi = friq_shepard_weight<P>(s);
        if (sw) sw[(size_t)i * FRIQ_TILE] = wi;
        vagc = __dadd_rn(vagc, __dmul_rn(wi, rbq[(size_t)i * FRIQ_TILE]));
        ws = __dadd_rn(ws, wi);
    }
}

// FIVE_vag_concl (FIVEVagConcl.c:64-351): returns hit index or -1, conclusion in conc
template <int K, int P>
__device__ __forceinline__ int friq_vag_concl(const friq_devprob &pb, const uint32_t *rbw, const double *rbq, double *sw, int R,
                                              const double *vo, double &conc, double &ws)
{
    const int hit = friq_first_hit<K>(pb, rbw, R, vo);
    if (hit >= 0) { conc = rbq[(size_t)hit * FRIQ_TILE]; ws = 0.0; return hit; }
    double vagc;
    friq_shepard<K, P>(pb, rbw, rbq, sw, R, vo, vagc, ws);
    conc = __ddiv_rn(vagc, ws);
    return -1;
}

// The same as real calls, for the fused kernels' rare branches (a snapped point off the observation, an unknown exact-hit
// rule after a per-call state change): inlined they put ~1000 instructions of library sqrt / x87 power / reciprocal into the
// step loop, whose body then no longer fits the 32 KB instruction cache level (no_instruction stalls at 2 warps per scheduler).
template <int K, int P>
__device__ __noinline__ int friq_vag_concl_cold(const friq_devprob *pb, const uint32_t *rbw, const double *rbq, int R, const double *vo, double *conc)
{
    double v[K], c, ws;
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = vo[k];
    const int hit = friq_vag_concl<K, P>(*pb, rbw, rbq, nullptr, R, v, c, ws);
    *conc = c;
    return hit;
}
template <int K>
__device__ __noinline__ int friq_first_hit_cold(const friq_devprob *pb, const uint32_t *rbw, int R, const double *vo)
{
    double v[K];
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = vo[k];
    return friq_first_hit<K>(*pb, rbw, R, v);
}

// ---------------------------------------------------------------------------------------
// frirl_get_best_action (frirl_get_best_action.c:31-341) + FIVEVagConcl_FRIRL_BestAct
// (:56-301): all A actions in one pass over the rule base (ACH actions per sweep when A is
// large), UNR rules per trip, i.e. UNR * ACH independent sqrt / pow / rcp chains in flight;
// the sums are still accumulated strictly in rule order.
//   hidx[a]  first rule at distance 0 from (observation, action a), or -1 (BestAct.c:89-93).
//            Because best action and FIVE_vag_concl see the same VE values (qp_reuse), hidx[a']
//            is also the exact-hit rule of Q(s',a') - the next step's Q(s,a) - so the step
//            kernel never needs a separate exact-hit scan.
//   WITHQ    additionally evaluates the Shepard sums of Q(s,a) at the previous antecedents
//            vo_q[] (FIVEVagConcl.c:224-235) as one more chain for the lanes with needq set,
//            keeping the weights wi in sw[] for the weighted update.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool friq_is_zero(double x)      // x == 0.0 for x >= +0, on the integer pipe
{
    return ((unsigned)__double2hiint(x) | (unsigned)__double2loint(x)) == 0u;
}

// one trip of the pass: UNR rules x ACH actions (+ UNR chains of Q(s,a) when QCH), all chains evaluated together
// stage by stage (friq_shepard_weights_n); the sums are then added in rule order
template <int S, int P, int ACH, bool QCH, int UNR, bool SAFE>
__device__ __forceinline__ void friq_pass_trip(const friq_devprob &pb, const uint32_t (&w)[UNR], const double (&q)[UNR], int i, int R,
                                               const double *vo_obs, const double (&ave)[ACH], const double *vo_q, bool needq, double *sw,
                                               double (&vagc)[ACH], double (&ws)[ACH], unsigned (&hi)[ACH], double &vagq, double &wsq)
{
    constexpr int NA = UNR * ACH, NC = NA + (QCH ? UNR : 0);
    double sv[NC], wv[NC];
    bool live[NC];
#pragma unroll
    for (int u = 0; u < UNR; u++) {
        double ssum = 0.0, sq = 0.0;
#pragma unroll
        for (int k = 0; k < S; k++) {
            const double v = pb.dict_ve[k][FRIQ_FIELD(w[u], k, S + 1)];
            const double d = __dsub_rn(vo_obs[k], v);
            const double d2 = __dmul_rn(d, d);
            ssum = (k == 0) ? d2 : __dadd_rn(ssum, d2);                              // A3/A4
            if (QCH) {
                const double e = __dsub_rn(vo_q[k], v);
                const double e2 = __dmul_rn(e, e);
                sq = (k == 0) ? e2 : __dadd_rn(sq, e2);
            }
        }
        const double va = pb.dict_ve[S][FRIQ_FIELD(w[u], S, S + 1)];
#pragma unroll
        for (int a = 0; a < ACH; a++) {
            const double da = __dsub_rn(ave[a], va);
            sv[u * ACH + a] = __dadd_rn(__dmul_rn(da, da), ssum);                    // A5
            live[u * ACH + a] = !friq_is_zero(sv[u * ACH + a]);                      // s == 0: exact hit, weight discarded (NaN)
        }
        if (QCH) {
            const double e = __dsub_rn(vo_q[S], va);
            sv[NA + u] = __dadd_rn(sq, __dmul_rn(e, e));
            live[NA + u] = needq && !friq_is_zero(sv[NA + u]);
        }
    }
    unsigned bad = 0;
    double dv[NC];
    friq_shepard_weights_n<P, SAFE, NC>(sv, wv, bad, live, dv);
    if (bad) {                                           // rare: a chain outside the proven fast paths -> exact routines, that chain only
#pragma unroll
        for (int n = 0; n < NC; n++)
            if ((bad >> n) & 1u) wv[n] = SAFE ? friq_shepard_weight_from_sqrt<P>(dv[n]) : friq_shepard_weight_slow<P>(live[n] ? sv[n] : 1.0);
    }
#pragma unroll
    for (int u = 0; u < UNR; u++) {
        if (u == 0 || i + u < R) {
#pragma unroll
            for (int a = 0; a < ACH; a++) {
                hi[a] = min(hi[a], live[u * ACH + a] ? 0xffffffffu : (unsigned)(i + u));  // first hit (BestAct.c:89-93): rules come in order
                vagc[a] = __dadd_rn(vagc[a], __dmul_rn(wv[u * ACH + a], q[u]));       // BestAct.c:216
                ws[a] = __dadd_rn(ws[a], wv[u * ACH + a]);                            // :217
            }
            if (QCH) {
                if (needq) sw[(size_t)(i + u) * FRIQ_TILE] = wv[NA + u];
                vagq = __dadd_rn(vagq, __dmul_rn(wv[NA + u], q[u]));                  // FIVEVagConcl.c:234
                wsq = __dadd_rn(wsq, wv[NA + u]);                                     // :235
            }
        }
    }
}

template <int S, int P, int A, int ACH, bool WITHQ, int UNR, bool SAFE>
__device__ __forceinline__ void friq_best_action_pass(const friq_devprob &pb, const uint32_t *rbw, const double *rbq, double *sw, int R,
                                                      const double *vo_obs, double *actconc, int *hidx,
                                                      const double *vo_q, bool needq, double &vagq, double &wsq)
{
    static_assert(A % ACH == 0, "action chunking");
    vagq = 0.0; wsq = 0.0;
#pragma unroll 1
    for (int c0 = 0; c0 < A; c0 += ACH) {
        double vagc[ACH], ws[ACH], ave[ACH];
        unsigned hi[ACH];
#pragma unroll
        for (int a = 0; a < ACH; a++) { vagc[a] = 0.0; ws[a] = 0.0; hi[a] = 0xffffffffu; ave[a] = pb.act_ve[c0 + a]; }
        uint32_t w[UNR]; double q[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) { w[u] = 0; q[u] = 0.0; if (u < R) { w[u] = rbw[(size_t)u * FRIQ_TILE]; q[u] = rbq[(size_t)u * FRIQ_TILE]; } }
#pragma unroll 1
        for (int i = 0; i < R; i += UNR) {
            uint32_t wn[UNR]; double qn[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) {                                                   // prefetch the next trip
                wn[u] = 0; qn[u] = 0.0;
                if (i + UNR + u < R) { wn[u] = rbw[(size_t)(i + UNR + u) * FRIQ_TILE]; qn[u] = rbq[(size_t)(i + UNR + u) * FRIQ_TILE]; }
            }
            if (WITHQ && (A == ACH || c0 == 0))
                friq_pass_trip<S, P, ACH, true, UNR, SAFE>(pb, w, q, i, R, vo_obs, ave, vo_q, needq, sw, vagc, ws, hi, vagq, wsq);
            else
                friq_pass_trip<S, P, ACH, false, UNR, SAFE>(pb, w, q, i, R, vo_obs, ave, vo_q, needq, sw, vagc, ws, hi, vagq, wsq);
#pragma unroll
            for (int u = 0; u < UNR; u++) { w[u] = wn[u]; q[u] = qn[u]; }
        }
#pragma unroll
        for (int a = 0; a < ACH; a++) {
            hidx[c0 + a] = (int)hi[a];
            actconc[c0 + a] = ((int)hi[a] >= 0) ? rbq[(size_t)hi[a] * FRIQ_TILE] : __ddiv_rn(vagc[a], ws[a]);   // :265
        }
    }
}

// ---------------------------------------------------------------------------------------
// One action at a time (the per-call kernels for a run-time number of actions): the whole rule base
// is walked in rule order for ONE action - the sums are accumulated exactly as in
// friq_best_action_pass - with UNR rules (= UNR independent sqrt / pow / rcp chains) per trip.
//   conc   Shepard conclusion of this action, or the consequent of the first rule at distance 0
//   hidx   that rule's index, -1 if none (FIVEVagConcl_FRIRL_BestAct.c:89-93)
// ---------------------------------------------------------------------------------------
template <int S, int P, int UNR, bool SAFE>
__device__ __forceinline__ void friq_action_pass(const friq_devprob &pb, const uint32_t *rbw, const double *rbq, int R,
                                                 const double *vo_obs, double ave, double &conc, int &hidx)
{
    double vagc = 0.0, ws = 0.0;
    unsigned hi = 0xffffffffu;
    uint32_t w[UNR]; double q[UNR];
#pragma unroll
    for (int u = 0; u < UNR; u++) { w[u] = 0; q[u] = 0.0; if (u < R) { w[u] = rbw[(size_t)u * FRIQ_TILE]; q[u] = rbq[(size_t)u * FRIQ_TILE]; } }
#pragma unroll 1
    for (int i = 0; i < R; i += UNR) {
        uint32_t wn[UNR]; double qn[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {                                                       // prefetch the next trip
            wn[u] = 0; qn[u] = 0.0;
            if (i + UNR + u < R) { wn[u] = rbw[(size_t)(i + UNR + u) * FRIQ_TILE]; qn[u] = rbq[(size_t)(i + UNR + u) * FRIQ_TILE]; }
        }
        double wv[UNR], sv[UNR]; bool z[UNR];
        unsigned bad = 0;
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            double ssum = 0.0;
#pragma unroll
            for (int k = 0; k < S; k++) {
                const double d = __dsub_rn(vo_obs[k], pb.dict_ve[k][FRIQ_FIELD(w[u], k, S + 1)]);
                const double d2 = __dmul_rn(d, d);
                ssum = (k == 0) ? d2 : __dadd_rn(ssum, d2);                                  // A3/A4
            }
            const double da = __dsub_rn(ave, pb.dict_ve[S][FRIQ_FIELD(w[u], S, S + 1)]);
            sv[u] = __dadd_rn(__dmul_rn(da, da), ssum);                                      // A5
            z[u] = friq_is_zero(sv[u]);
            wv[u] = friq_shepard_weight_fast<P, SAFE>(sv[u], bad, !z[u]);                    // NaN when s == 0: discarded
        }
        if (bad) {                                       // rare: an operand outside the proven fast paths -> library routines
#pragma unroll
            for (int u = 0; u < UNR; u++) wv[u] = friq_shepard_weight_slow<P>(z[u] ? 1.0 : sv[u]);
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            if (u == 0 || i + u < R) {
                hi = min(hi, z[u] ? (unsigned)(i + u) : 0xffffffffu);                        // first hit: rules come in increasing order
                vagc = __dadd_rn(vagc, __dmul_rn(wv[u], q[u]));                              // BestAct.c:216
                ws = __dadd_rn(ws, wv[u]);                                                   // :217
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) { w[u] = wn[u]; q[u] = qn[u]; }
    }
    hidx = (int)hi;
    conc = (hidx >= 0) ? rbq[(size_t)hidx * FRIQ_TILE] : __ddiv_rn(vagc, ws);               // :265
}

// ---------------------------------------------------------------------------------------
// Cooperative Q(s,a) (FIVEVagConcl.c:224-235) for the FEW lanes of a warp whose pending update has no exact-hit rule.
// Once learning is under way only 1-2 % of the agent-steps need these Shepard sums, but a third of the warp-steps
// contain at least one such lane; carrying the Q(s,a) chain inside the pass (WITHQ) then makes all 32 lanes pay a
// fourth chain per rule for the sake of one.  Here the needy agents are served one after the other: all lanes of the
// warp compute the weights of THAT agent's rules in parallel (lane r takes rules r, r + n, ...) and write them to
// sw[]; the owning lane then adds them up in rule order, so the sums are bit-identical to the sequential loop.
//   wmask  lanes of the warp taking part, qm  lanes that need Q(s,a)
//   rbw / rbq / sw point at the calling lane's column of the tile ([slot][FRIQ_TILE])
// ---------------------------------------------------------------------------------------
template <int K, int P, bool SAFE>
__device__ __forceinline__ void friq_q_cooperative(const friq_devprob &pb, const uint32_t *rbw, const double *rbq, double *sw, int R,
                                                   const double *vo_q, unsigned wmask, unsigned qm, double &vagq, double &wsq)
{
    const int lane = threadIdx.x & 31;
    const int nw = __popc(wmask), rank = __popc(wmask & ((1u << lane) - 1u));
    const uint32_t *tw = rbw - lane;                     // column 0 of the tile
    double *tsw = sw - lane;
    for (unsigned m = qm; m; m &= m - 1) {
        const int j = __ffs(m) - 1;
        const int Rj = __shfl_sync(wmask, R, j);
        double vj[K];
#pragma unroll
        for (int k = 0; k < K; k++) vj[k] = __shfl_sync(wmask, vo_q[k], j);
        for (int i = rank; i < Rj; i += nw) {
            const double sq = friq_dist2<K>(pb, tw[(size_t)i * FRIQ_TILE + j], vj);
            unsigned bad = 0;
            double wq = friq_shepard_weight_fast<P, SAFE>(sq, bad);      // sq != 0: the agent has no rule at this point
            if (bad) wq = friq_shepard_weight_slow<P>(sq);
            tsw[(size_t)i * FRIQ_TILE + j] = wq;
        }
    }
    __syncwarp(wmask);                                   // the weights written by the other lanes are visible
    vagq = 0.0; wsq = 0.0;
    if ((qm >> lane) & 1u)
        for (int i = 0; i < R; i++) {
            const double wi = sw[(size_t)i * FRIQ_TILE];
            vagq = __dadd_rn(vagq, __dmul_rn(wi, rbq[(size_t)i * FRIQ_TILE]));     // FIVEVagConcl.c:234
            wsq = __dadd_rn(wsq, wi);                                              // :235
        }
}

// ---------------------------------------------------------------------------------------
// Exact hits first (FIVEVagConcl_FRIRL_BestAct.c:89-93, FIVEVagConcl.c:96-105): the reference returns the consequent
// of the first rule at distance 0 and never evaluates the Shepard sums then.  Once an agent has learned for a while
// almost every (observation, action) pair it visits has such a rule (mountaincar: 92-98 %), so evaluating all A
// Shepard chains for all 32 lanes and discarding most of them wastes the FP64 pipe.  When the problem block says
// that distance 0 is the same as equal dictionary indices (hit_scan), a first sweep compares packed rule words with
// the packed (observation, action) targets - one integer compare per rule and action - and only the pairs WITHOUT a
// hit go on to the Shepard sums: cooperatively when they are few (friq_actions_cooperative), through the lock-step
// pass when they are many.
// ---------------------------------------------------------------------------------------
#define FRIQ_NO_TARGET 0xfffffffeu                      /* never equals a packed rule word (K * 6 <= 30 bits) nor the padding word */

// packed dictionary indices of the observation with VE values vo[] (grid indices ci[k] when known), or FRIQ_NO_TARGET when
// some dimension's VE value is no dictionary entry's (then no rule can be at distance 0)
//   ci[k] >= 0   the observation is the grid value with that index (quantiser)
//   ci[k] == -1  unknown: search the dictionary for the VE value (custom start states)
//   ci[k] == -2  the default episode start state: the host has done the search (def_hit)
// index of the dictionary entry with VE value v, -1 if none (rare: custom start states; out of line to keep the step loop small)
__device__ __noinline__ int friq_dict_find(const double *ve, int n, double v)
{
    for (int t = 0; t < n; t++) if (ve[t] == v) return t;
    return -1;
}
__device__ __noinline__ double friq_obs_ve_cold(const friq_devprob *pb, int k, double x) { return friq_obs_ve(*pb, k, x); }
template <int S>
__device__ __forceinline__ uint32_t friq_obs_word(const friq_devprob &pb, const double *vo, const int *ci)
{
    uint32_t word = 0; bool ok = true;
#pragma unroll
    for (int k = 0; k < S; k++) {
        int j = ci[k];
        if (j == -2) j = pb.def_hit[k];
        else if (j < 0) j = friq_dict_find(pb.dict_ve[k], pb.dict_len[k], vo[k]);
        ok = ok && j >= 0;
        word |= (uint32_t)(j & (FRIQ_DICT_MAX(S + 1) - 1)) << (FRIQ_BITS_OF(S + 1) * k);
    }
    return ok ? word : FRIQ_NO_TARGET;
}

// first rule equal to each target word, -1 if none
template <int A>
__device__ __forceinline__ void friq_hit_scan(const uint32_t *rbw, int R, const uint32_t (&tgt)[A], int (&hidx)[A])
{
    unsigned hi[A];
#pragma unroll
    for (int a = 0; a < A; a++) hi[a] = 0xffffffffu;
#pragma unroll 1
    for (int i = 0; i < R; i += 8) {
        uint32_t w[8];
#pragma unroll
        for (int u = 0; u < 8; u++) w[u] = (i + u < R) ? rbw[(size_t)(i + u) * FRIQ_TILE] : 0xffffffffu;
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int a = 0; a < A; a++) hi[a] = min(hi[a], (w[u] == tgt[a]) ? (unsigned)(i + u) : 0xffffffffu);
    }
#pragma unroll
    for (int a = 0; a < A; a++) hidx[a] = (int)hi[a];
}

// Shepard sums of the (lane, channel) pairs that have no exact-hit rule - COMPACTED onto the lanes of the warp.
//   channels 0 .. A-1  the conclusion of action a at the current observation vo_obs[] (BestAct.c:212-217, :265), pairs nm[a]
//   channel  A         Q(s,a) of the pending update at the previous antecedents vo_q[] (FIVEVagConcl.c:224-235), lanes qm;
//                      its weights are also written to sw[] for the weighted update
// The lock-step pass spends A chains per rule on every lane although, once learning is under way, only a few of the warp's
// 32 * (A + 1) pairs still need Shepard sums.  Here pair number p of the warp is handed to lane p: the lane fetches the
// owner's point (shuffles), walks the OWNER's rule-base column - the 32 columns of a tile share their cache lines row by
// row, so these reads stay coalesced - and accumulates the sums sequentially in its own registers exactly like the
// lock-step pass does (UNR rules = UNR independent chains per trip, friq_shepard_weights_n).  One round serves as many
// pairs as the warp has lanes for the price of ONE chain per rule; the owners then pick their results up - with shuffles in the
// computed mode, through the warp's scratch in shared memory in table mode (friq_round_smem).  Bit-identical sums.
// dictionary cell of a packed word (row / column of the memoised weights pb.wtab)
template <int K>
__device__ __forceinline__ int friq_cell(const friq_devprob &pb, uint32_t w)
{
    int c = (int)FRIQ_FIELD(w, 0, K);
#pragma unroll
    for (int k = 1; k < K; k++) c += (int)FRIQ_FIELD(w, k, K) * pb.wlin[k];
    return c;
}

// ---- table mode: the per-agent list of rule cells (friq_devstate.rbc) is stored in blocks of FRIQ_CELL_VEC rules per lane,
// [tile][slot / 8][lane][8] u16, so that a lane fetches the cells of 8 consecutive rules of a column with ONE 16-byte load;
// the unused tail of the last block and FRIQ_RBC_AHEAD whole blocks behind it hold the table's zero column (friq_zero_cell).
// Lane-minor on purpose: the lanes of a warp walk columns of the SAME tile, so their loads of a slot share cache lines; three
// agent-major variants of these arrays measured 4-38 % slower (profiles/README.md).
#ifndef FRIQ_CELL_VEC
#define FRIQ_CELL_VEC 8         /* 8 (one 16-byte load per block) or 4 (8-byte loads, half the registers in flight) */
#endif
__host__ __device__ __forceinline__ size_t friq_cap8(int cap) { return (size_t)((cap + FRIQ_CELL_VEC - 1) & ~(FRIQ_CELL_VEC - 1)); }
#define FRIQ_RBC_AHEAD 4        /* blocks of zero cells kept behind a column's last block (the walk requests cell vectors up to 4 blocks ahead, unclamped) */
__host__ __device__ __forceinline__ size_t friq_rbc_slots(int cap) { return friq_cap8(cap) + (size_t)FRIQ_RBC_AHEAD * FRIQ_CELL_VEC; }   // rule slots per column of friq_devstate.rbc
__device__ __forceinline__ size_t friq_rbc_index(int i, int lane) { return ((size_t)(i / FRIQ_CELL_VEC) * FRIQ_TILE + lane) * FRIQ_CELL_VEC + (i % FRIQ_CELL_VEC); }
// The memoised table has one more column than cells, all +0.0 (friq_zero_cell): the cell of every slot between a column's last
// rule and the end of its block, and of FRIQ_RBC_AHEAD whole blocks behind it, so that whole blocks can be summed without a tail.
__device__ __forceinline__ unsigned friq_zero_cell(const friq_devprob &pb) { return (unsigned)pb.wcells; }
#if FRIQ_CELL_VEC == 8
typedef uint4 friq_cellvec;
__device__ __forceinline__ friq_cellvec friq_cellvec_zero() { return make_uint4(0u, 0u, 0u, 0u); }
__device__ __forceinline__ void friq_cellvec_unpack(const friq_cellvec c, unsigned (&cc)[8])
{ cc[0] = __byte_perm(c.x, 0, 0x4410); cc[1] = c.x >> 16; cc[2] = __byte_perm(c.y, 0, 0x4410); cc[3] = c.y >> 16;
  cc[4] = __byte_perm(c.z, 0, 0x4410); cc[5] = c.z >> 16; cc[6] = __byte_perm(c.w, 0, 0x4410); cc[7] = c.w >> 16; }   // PRMT: no shift+mask pair folded into 64-bit address math
#else
typedef uint2 friq_cellvec;
__device__ __forceinline__ friq_cellvec friq_cellvec_zero() { return make_uint2(0u, 0u); }
__device__ __forceinline__ void friq_cellvec_unpack(const friq_cellvec c, unsigned (&cc)[4])
{ cc[0] = c.x & 0xffffu; cc[1] = c.x >> 16; cc[2] = c.y & 0xffffu; cc[3] = c.y >> 16; }
#endif

// the pair's row of the memoised weights (global memory, read-only path)
struct friq_row_f64 {
    const double *__restrict__ row;
    __device__ __forceinline__ double operator()(unsigned c) const { return __ldg(row + c); }
};

// weights (looked up in the pair's table row through the block's cells) and consequents of one block of rules
template <class ROW>
__device__ __forceinline__ void friq_tab_load_t(const ROW &row, const friq_cellvec c, double (&t)[FRIQ_CELL_VEC])
{
    unsigned cc[FRIQ_CELL_VEC];
    friq_cellvec_unpack(c, cc);
#pragma unroll
    for (int u = 0; u < FRIQ_CELL_VEC; u++) t[u] = row(cc[u]);
}
__device__ __forceinline__ void friq_tab_load_q(const double *pq, double (&q)[FRIQ_CELL_VEC])
{
#pragma unroll
    for (int u = 0; u < FRIQ_CELL_VEC; u++) q[u] = pq[(size_t)u * FRIQ_TILE];
}
// the block's terms added in rule order (BestAct.c:216-217 / FIVEVagConcl.c:234-235).  Every slot of the block is added: the
// slots past the column's last rule carry the weight +0.0 (their cell is the table's ZERO column, friq_zero_cell) and a finite
// consequent, and x + (+-0.0) == x bit for bit for every x a sum can hold (a sum that starts at +0.0 is never -0.0).
// KEEP: the weights are also left in sw[], whole blocks, unconditionally (a round of Q(s,a) pairs; the rule capacity is a
// multiple of the block size in table mode, so the zeros of a last block's tail stay inside the column)
template <bool KEEP>
__device__ __forceinline__ void friq_tab_acc(const double (&t)[FRIQ_CELL_VEC], const double (&q)[FRIQ_CELL_VEC], double *psw, double &vag, double &ws)
{
    if constexpr (KEEP) {
#pragma unroll
        for (int u = 0; u < FRIQ_CELL_VEC; u++) psw[(size_t)u * FRIQ_TILE] = t[u];
    }
#pragma unroll
    for (int u = 0; u < FRIQ_CELL_VEC; u++) { vag = __dadd_rn(vag, __dmul_rn(t[u], q[u])); ws = __dadd_rn(ws, t[u]); }
}
// Shepard sums of one pair from the memoised weights: rules 0 .. Rt-1 of column j of the tile (pc = the column's cell blocks,
// pq = its consequents, psw = its weight scratch; c0 / qA = block 0's cells and consequents, already on their way).
// The loop body has NO branch and no tail: ptxas gives the loads of a loop body with control flow in it ONE scoreboard (every
// wait then drains the loads issued just before it and a double buffer overlaps nothing: measured, two exposed load latencies
// per 16 rules); branch-free, the two halves of the trip get their own scoreboards.  A trip takes two blocks of 8 rules; the
// blocks past the column's last one hold zero cells (FRIQ_RBC_AHEAD of them are kept there: k_refresh_index, the append in
// k_run_steps), their consequents are the last block's again; the cell vectors run a whole trip ahead of the loads they
// address (they are the first link of the chain cells -> table entry).
template <bool KEEP, class ROW>
__device__ __forceinline__ void friq_tab_sums(const ROW &row, const friq_cellvec *pc, const double *pq, double *psw,
                                              const friq_cellvec c0, double (&qA)[FRIQ_CELL_VEC], int Rt, double &vag, double &ws)
{
    constexpr int B = FRIQ_CELL_VEC;
    constexpr size_t BS = (size_t)B * FRIQ_TILE;
    if (Rt <= 0) return;
    const int last = (Rt - 1) / B;                       // the column's last block
    double tA[B], tB[B], qB[B];
    friq_tab_load_t(row, c0, tA);
    friq_cellvec c1 = pc[FRIQ_TILE], c2 = pc[2 * FRIQ_TILE];
    int b = 0;
#pragma unroll 1
    for (; b < last; b += 2) {
        friq_tab_load_t(row, c1, tB);
        friq_tab_load_q(pq + (size_t)min(b + 1, last) * BS, qB);
        const friq_cellvec c3 = pc[3 * FRIQ_TILE];
        friq_tab_acc<KEEP>(tA, qA, psw, vag, ws);
        friq_tab_load_t(row, c2, tA);
        friq_tab_load_q(pq + (size_t)min(b + 2, last) * BS, qA);
        const friq_cellvec c4 = pc[4 * FRIQ_TILE];
        friq_tab_acc<KEEP>(tB, qB, psw + BS, vag, ws);
        c1 = c3; c2 = c4; pc += 2 * FRIQ_TILE; psw += 2 * BS;
    }
    if (b == last) friq_tab_acc<KEEP>(tA, qA, psw, vag, ws);      // an odd number of blocks: the last one is already in tA / qA
}
// lane number of the n-th (0-based) set bit of mask: partial warps only (the last warp of a batch); __fns expands to ~220
// instructions per use, so it lives out of line
__device__ __noinline__ int friq_nth_lane(unsigned mask, int n) { return (int)__fns(mask, 0, n + 1); }

// the computed sums of one pair, one rule at a time through the library-exact chain: the table-mode kernel's rare fall-back
// (a point that is no dictionary cell: custom off-grid start states), kept out of line so that it costs the kernel no registers
template <int K, int P>
__device__ __noinline__ void friq_pair_sums_slow(const friq_devprob *pbp, const uint32_t *cw, const double *cq, double *csw, int Rp, int isq,
                                                 const double *pt, double *out)
{
    double vag = 0.0, ws = 0.0, p[K];
#pragma unroll
    for (int k = 0; k < K; k++) p[k] = pt[k];
    for (int i = 0; i < Rp; i++) {
        const double wi = friq_shepard_weight<P>(friq_dist2<K>(*pbp, cw[(size_t)i * FRIQ_TILE], p));
        if (isq) csw[(size_t)i * FRIQ_TILE] = wi;
        vag = __dadd_rn(vag, __dmul_rn(wi, cq[(size_t)i * FRIQ_TILE]));
        ws = __dadd_rn(ws, wi);
    }
    out[0] = vag; out[1] = ws;
}

// a warp's scratch for one round in table mode (shared memory): the owners post one descriptor per pair - x = offset of the
// pair's row in the memoised table (FRIQ_NO_ROW: the point is no dictionary cell), y = owner lane | channel << 8 | rule count
// << 16 - and the lanes that summed a pair post {conclusion (Q(s,a): the weighted sum), sum of the weights} for the owner
#define FRIQ_NO_ROW 0xffffffffu
struct __align__(16) friq_round_smem { uint2 desc[FRIQ_TILE]; double2 res[FRIQ_TILE]; };
template <int S, int P, int A, int UNR, bool SAFE, bool TAB>
__device__ __forceinline__ void friq_pairs_compacted(const friq_devprob &pb, const uint32_t *rbw, const double *rbq, double *sw,
                                                     const unsigned short *rbc /* TAB: the TILE's rule cells (friq_devstate.rbc layout) */, int R,
                                                     const double *vo_obs, const double *vo_q,
                                                     const unsigned (&arow)[A], unsigned qrow /* TAB: table rows of my action pairs / my Q(s,a) pair */,
                                                     friq_round_smem *rs /* TAB: the warp's round scratch */,
                                                     unsigned wmask, const unsigned (&nm)[A], unsigned qm,
                                                     unsigned char *plist /* 32 bytes of shared memory per warp */,
                                                     double (&conc)[A], double &vagq, double &wsq, unsigned &n_tab, unsigned &n_chain, bool keep_sw)
{
    constexpr int K = S + 1, C = A + 1;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t *tw = rbw - lane;                     // column 0 of the tile
    const double *tq = rbq - lane;
    double *tsw = sw - lane;
    unsigned chm[C]; int ofs[C + 1];
#pragma unroll
    for (int a = 0; a < A; a++) chm[a] = nm[a];
    chm[A] = qm;
    const int per_round = __popc(wmask), rank = __popc(wmask & lt);   // the last warp of a batch may be partial
    ofs[0] = 0;
#pragma unroll
    for (int c = 0; c < A; c++) ofs[c + 1] = ofs[c] + __popc(chm[c]);
    // table mode with kept weights: the Q(s,a) pairs get rounds of their own, so that a round either stores every weight it
    // looks up or none (the walk's loop body must stay branch-free, friq_tab_sums)
    const int end_act = ofs[A];
    if (TAB && keep_sw) ofs[A] = ((ofs[A] + per_round - 1) / per_round) * per_round;
    ofs[C] = ofs[A] + __popc(chm[A]);
    const int np = ofs[C];
    vagq = 0.0; wsq = 0.0;
    for (int p0 = 0; p0 < np; p0 += per_round) {
        // ---- which pair is mine?  the owners post (lane, channel) of their pairs of this round, lane number `rank` reads entry `rank`
        __syncwarp(wmask);
#pragma unroll
        for (int t = 0; t < C; t++) {
            const int e = ofs[t] + __popc(chm[t] & lt) - p0;
            if (((chm[t] >> lane) & 1u) && e >= 0 && e < per_round) {
                if constexpr (TAB) rs->desc[e] = make_uint2(t < A ? arow[t < A ? t : 0] : qrow, (unsigned)lane | ((unsigned)t << 8) | ((unsigned)R << 16));
                else plist[e] = (unsigned char)(lane | (t << 5));
            }
        }
        __syncwarp(wmask);
        const bool has = p0 + rank < np && !(p0 + rank >= end_act && p0 + rank < ofs[A]);
        uint2 desc = make_uint2(FRIQ_NO_ROW, 0u);
        int ent = 0;
        if constexpr (TAB) { if (has) desc = rs->desc[rank]; ent = (int)(desc.y & 31u) | (int)(((desc.y >> 8) & 0xffu) << 5); }
        else ent = has ? plist[rank] : 0;
        const int c = ent >> 5, j = ent & 31;
        // ---- the owner's point (table mode needs it only for the rare computed fall-back) and rule count
        double pt[K];
        auto fetch_point = [&]() {
#pragma unroll
            for (int k = 0; k < S; k++) {
                const double x = __shfl_sync(wmask, vo_obs[k], j), y = __shfl_sync(wmask, vo_q[k], j);
                pt[k] = (c == A) ? y : x;
            }
            const double y = __shfl_sync(wmask, vo_q[S], j);
            pt[S] = (c == A) ? y : pb.act_ve[c < A ? c : 0];
        };
        if constexpr (!TAB) fetch_point();
        // table mode: block 0 of the owner's column (cells, consequents) is requested before anything else about the pair is known
        const friq_cellvec *pc = reinterpret_cast<const friq_cellvec *>(rbc) + j;
        friq_cellvec c0 = friq_cellvec_zero(); double q0[FRIQ_CELL_VEC];
        if constexpr (TAB) { c0 = pc[0]; friq_tab_load_q(tq + j, q0); }
        int Rj;
        if constexpr (TAB) Rj = (int)(desc.y >> 16); else Rj = __shfl_sync(wmask, R, j);
        const bool isq = has && c == A;
        double vag = 0.0, ws = 0.0;
        // ---- table mode: a pair whose point is a dictionary cell looks the weight of every rule up (pb.wtab)
        if constexpr (TAB) {
            const bool tab = has && desc.x != FRIQ_NO_ROW;
            if (tab) {
                const double *rowd = pb.wtab + desc.x;
                asm volatile("" : "+l"(rowd));           // one base register for the row (lookups stay base + cell * 8)
                if (keep_sw && p0 >= ofs[A]) friq_tab_sums<true>(friq_row_f64{rowd}, pc, tq + j, tsw + j, c0, q0, Rj, vag, ws);   // a round of Q(s,a) pairs
                else friq_tab_sums<false>(friq_row_f64{rowd}, pc, tq + j, tsw + j, c0, q0, Rj, vag, ws);
                n_tab += (unsigned)Rj;
            }
            const bool slow = has && !tab;
            if (__any_sync(wmask, slow)) {
                fetch_point();
                if (slow) {
                    double out[2];
                    friq_pair_sums_slow<K, P>(&pb, tw + j, tq + j, tsw + j, Rj, isq ? 1 : 0, pt, out);
                    vag = out[0]; ws = out[1]; n_chain += (unsigned)Rj;
                }
            }
        }
        if constexpr (!TAB) {
        const int Rp = has ? Rj : 0;
        n_chain += (unsigned)Rp;
        // ---- sequential Shepard sums over the owner's rules, UNR rules (= UNR independent chains) per trip
        uint32_t w[UNR]; double q[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
            w[u] = 0; q[u] = 0.0;
            if (u < Rp) { w[u] = tw[(size_t)u * FRIQ_TILE + j]; q[u] = tq[(size_t)u * FRIQ_TILE + j]; }
        }
#pragma unroll 1
        for (int i = 0; i < Rp; i += UNR) {
            uint32_t wn[UNR]; double qn[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) {              // prefetch the next trip
                const int r = i + UNR + u;
                wn[u] = 0; qn[u] = 0.0;
                if (r < Rp) { wn[u] = tw[(size_t)r * FRIQ_TILE + j]; qn[u] = tq[(size_t)r * FRIQ_TILE + j]; }
            }
            double sv[UNR], wv[UNR], wq[UNR], dv[UNR]; bool live[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) { sv[u] = friq_dist2<K>(pb, w[u], pt); live[u] = i + u < Rp; }   // != 0 for real rules: no exact hit
            unsigned bad = 0;
            friq_shepard_weights_n<P, SAFE, UNR>(sv, wv, bad, live, dv);
            if (bad) {
#pragma unroll
                for (int u = 0; u < UNR; u++)
                    if ((bad >> u) & 1u) wv[u] = SAFE ? friq_shepard_weight_from_sqrt<P>(dv[u]) : friq_shepard_weight_slow<P>(live[u] ? sv[u] : 1.0);
            }
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                wq[u] = __dmul_rn(wv[u], q[u]);                                               // BestAct.c:216 / FIVEVagConcl.c:234, the product
                if (isq && live[u]) tsw[(size_t)(i + u) * FRIQ_TILE + j] = wv[u];
            }
#pragma unroll
            for (int u = 0; u < UNR; u++)
                if (u == 0 || i + u < Rp) { vag = __dadd_rn(vag, wq[u]); ws = __dadd_rn(ws, wv[u]); }   // :216-217 / :234-235
#pragma unroll
            for (int u = 0; u < UNR; u++) { w[u] = wn[u]; q[u] = qn[u]; }
        }
        }   // !TAB
        const double res = __ddiv_rn(vag, ws);                                               // BestAct.c:265
        // ---- hand the results back to the owners
        if constexpr (TAB) {                             // through the warp's scratch: entry e was summed by the lane of rank e
            rs->res[rank] = make_double2(c == A ? vag : res, ws);
            __syncwarp(wmask);
#pragma unroll
            for (int t = 0; t < C; t++) {
                const int e = ofs[t] + __popc(chm[t] & lt) - p0;
                if (((chm[t] >> lane) & 1u) && e >= 0 && e < per_round) {
                    const double2 v = rs->res[e];
                    if (t < A) conc[t < A ? t : 0] = v.x; else { vagq = v.x; wsq = v.y; }
                }
            }
            continue;
        }
        __syncwarp(wmask);
#pragma unroll
        for (int t = 0; t < C; t++) {
            const int e = ofs[t] + __popc(chm[t] & lt) - p0;         // my pair of channel t is entry e of this round (if I have one)
            const bool mine = ((chm[t] >> lane) & 1u) && e >= 0 && e < per_round;
            const int src = mine ? e : rank;                          // rank of the lane that summed it
            const int sl = (wmask == 0xffffffffu) ? src : friq_nth_lane(wmask, src);
            if (t < A) { const double r = __shfl_sync(wmask, res, sl); if (mine) conc[t < A ? t : 0] = r; }
            else { const double a1 = __shfl_sync(wmask, vag, sl), a2 = __shfl_sync(wmask, ws, sl); if (mine) { vagq = a1; wsq = a2; } }
        }
    }
}

// first index of the maximum (src/inl/max.inl:16-28, strict <)
template <int A>
__device__ __forceinline__ int friq_argmax(const double *v)
{
    int best = 0; double bv = v[0];
#pragma unroll
    for (int i = 1; i < A; i++) if (bv < v[i]) { bv = v[i]; best = i; }
    return best;
}
template <int A>
__device__ __forceinline__ double friq_pick(const double *v, int idx)
{
    double r = v[0];
#pragma unroll
    for (int i = 1; i < A; i++) if (i == idx) r = v[i];
    return r;
}

template <int A>
__device__ __forceinline__ int friq_pick_int(const int *v, int idx)
{
    int r = v[0];
#pragma unroll
    for (int i = 1; i < A; i++) if (i == idx) r = v[i];
    return r;
}

// frirl_e_greedy_selection.c:21-37 with the Philox stream (explore when u1 <= epsilon; the
// random action is min(A-1, floor(u2*A)) - the reference's round(u*A) can return A).
template <int A>
__device__ __forceinline__ int friq_select_action(const friq_devprob &pb, const double *actconc, double epsilon,
                                                  unsigned long long seed, uint32_t episode, uint32_t step)
{
    int best = friq_argmax<A>(actconc);
    if (pb.no_random == 1 || epsilon == 0.0) return best;
    uint32_t r[4];
    friq_philox4x32_10(episode, step, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u1 = (double)((((uint64_t)r[1] << 32) | r[0]) >> 11) * 0x1p-53;
    const double u2 = (double)((((uint64_t)r[3] << 32) | r[2]) >> 11) * 0x1p-53;
    if (u1 > epsilon) return best;
    const int ra = (int)floor(__dmul_rn(u2, (double)A));
    return ra > A - 1 ? A - 1 : ra;
}

// frirl_check_possible_states.c:53-122: index of the grid value the observation snaps to
// (ties -> upper).  The grid-refinement branch (:68) cannot fire when epsilon equals the
// grid step; the host layer rejects problems where it could.
__device__ __forceinline__ int friq_snap_index(const friq_devprob &pb, int k, double x)
{
    const double *g = pb.dict_val[k];
    const int n = pb.grid_len[k];
    int i;
    for (i = 0; i < n; i++) if (x < g[i]) break;
    if (i == n) return n - 1;
    if (i == 0) return 0;
    i--;
    return (__dsub_rn(x, g[i]) < __dsub_rn(g[i + 1], x)) ? i : i + 1;
}

// ---------------------------------------------------------------------------------------
// frirl_update_sarsa (frirl_update_sarsa.c:348-385) + update_rules (:22-143), given
//   qp            Q(s',a')                                  (:355)
//   hitq / qnow   exact-hit rule and value of Q(s,a)        (:356); when hitq < 0, sw[] holds
//                 the wi of that evaluation and ws their sequential sum (re-used by
//                 FIVE_vag_concl_weight, FIVEVagConclWeight.c:125-175, same point, same RB)
//   qs[] / vo_q[] the antecedents (s,a) and their VE values.
// ---------------------------------------------------------------------------------------
#ifndef FRIQ_WUPD_VEC
#define FRIQ_WUPD_VEC 4         /* rules per trip of the weighted-update loop (loads in flight per L2 round trip; 8 and 16 measured no better) */
#endif
struct friq_row_none { __device__ __forceinline__ double operator()(unsigned) const { return 0.0; } };
template <int S, int P, bool COLD = false, class ROW = friq_row_none>
__device__ __forceinline__ void friq_sarsa_apply(const friq_devprob &pb, uint32_t *rbw, double *rbq, double *sw, int cap, int &R, int &flags,
                                                 const double *qs, const int *qi, double reward, double qp, int hitq, double qnow, double ws,
                                                 double alpha, double gamma, double *wref = nullptr, unsigned *n_wupd = nullptr,
                                                 const ROW *wrow = nullptr, const unsigned short *rbc_tile = nullptr)
{   // qi[k] >= 0: qs[k] is known to BE the grid value dict_val[k][qi[k]] (snapping an on-grid value returns it)
    constexpr int K = S + 1;
    const double qdiff = __dmul_rn(alpha, __dsub_rn(__dadd_rn(reward, __dmul_rn(gamma, qp)), qnow));   // :358
    if (qdiff > pb.qdiff_pos || qdiff < pb.qdiff_neg) {
        int gi[K]; bool same = true; uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (qi[k] >= 0) gi[k] = qi[k];
            else { gi[k] = friq_snap_index(pb, k, qs[k]); same = same && (pb.dict_val[k][gi[k]] == qs[k]); }   // :146-170
            packed |= (uint32_t)gi[k] << (FRIQ_BITS_OF(K) * k);
        }
        int hitg = hitq; double qg = qnow;
        if (!same) {                                   // snapped point differs: evaluate Q there (:370)
            double vg[K], wsg;
#pragma unroll
            for (int k = 0; k < K; k++) vg[k] = pb.dict_ve[k][gi[k]];
            if constexpr (COLD) hitg = friq_vag_concl_cold<K, P>(&pb, rbw, rbq, R, vg, &qg);
            else hitg = friq_vag_concl<K, P>(pb, rbw, rbq, nullptr, R, vg, qg, wsg);
        }
        if (hitg < 0) {                                // no rule there: append (five_add_rule.c:47-95)
            flags |= FRIQ_F_INSERTED;
            if (R < cap) { rbw[(size_t)R * FRIQ_TILE] = packed; rbq[(size_t)R * FRIQ_TILE] = __dadd_rn(qg, qdiff); R++; }
            else flags |= FRIQ_F_OVERFLOW;
            return;
        }
        flags &= ~FRIQ_F_INSERTED;                     // :378
    }
    int rules = R - ((flags & FRIQ_F_INSERTED) ? 1 : 0);
    if (hitq >= 0) {
        if (pb.skip_rules == 0 || hitq < rules) rbq[(size_t)hitq * FRIQ_TILE] = __dadd_rn(qnow, qdiff);   // :55
        return;                                                                                          // :61-63
    }
    if (wref) {                                        // FIVERB.weights as the reference leaves it (FIVEVagConclWeight.c:153-175)
#pragma unroll 1
        for (int i = 0; i < R; i++) wref[(size_t)i * FRIQ_TILE] = __ddiv_rn(sw[(size_t)i * FRIQ_TILE], ws);
    }
    if (pb.skip_rules == 0) { rules = R; flags &= ~FRIQ_F_INSERTED; }
    // cheap superset of (wi/ws > wsig): only the few rules that pass it pay the division; four rules per trip, one branch
    // (the weights were stored by ANOTHER lane a moment ago and global stores do not stay in L1: every trip is one L2 round trip -
    // FRIQ_WUPD_VEC rules share it.  ncu, acrobot: this loop holds 24 % of the kernel's samples at 4; measured: 8 = no change, 16 = 19 % slower, profiles/README.md)
    const double pre = (pb.wsig >= 0.05) ? __dmul_rn(0.04, ws) : -1.0;
    if (n_wupd) *n_wupd += (unsigned)rules;
    if (wrow) {
        // table mode, FEW lanes of the warp updating (friq_pairs_compacted then does not store the weights it sums): the owner
        // looks the weights up again - the pair's table row and its own cell blocks are in L1 from the sums - instead of reading
        // what another lane would have had to store for it (mountaincar: +2.7 %; with MANY updating lanes, acrobot, the stored
        // weights win because the owners' reads of one slot coalesce: -10 % when looked up, profiles/README.md)
        const friq_cellvec *pc = reinterpret_cast<const friq_cellvec *>(rbc_tile) + (threadIdx.x & 31);
#pragma unroll 1
        for (int i = 0; i < rules; i += FRIQ_CELL_VEC) {
            unsigned cc[FRIQ_CELL_VEC];
            friq_cellvec_unpack(pc[(size_t)(i / FRIQ_CELL_VEC) * FRIQ_TILE], cc);
            double wi[FRIQ_CELL_VEC]; bool any = false;
#pragma unroll
            for (int u = 0; u < FRIQ_CELL_VEC; u++) { wi[u] = (*wrow)(cc[u]); if (i + u >= rules) wi[u] = -2.0; any = any || (wi[u] > pre); }
            if (!any) continue;
#pragma unroll
            for (int u = 0; u < FRIQ_CELL_VEC; u++) {
                if (wi[u] > pre) {
                    const double wgt = __ddiv_rn(wi[u], ws);                                                         // A6
                    if (wgt > pb.wsig) rbq[(size_t)(i + u) * FRIQ_TILE] = __dadd_rn(qnow, __dmul_rn(qdiff, wgt));    // A7
                }
            }
        }
        return;
    }
#pragma unroll 1
    for (int i = 0; i < rules; i += FRIQ_WUPD_VEC) {
        double wi[FRIQ_WUPD_VEC]; bool any = false;
#pragma unroll
        for (int u = 0; u < FRIQ_WUPD_VEC; u++) { wi[u] = (i + u < rules) ? sw[(size_t)(i + u) * FRIQ_TILE] : -2.0; any = any || (wi[u] > pre); }
        if (!any) continue;
#pragma unroll
        for (int u = 0; u < FRIQ_WUPD_VEC; u++) {
            if (wi[u] > pre) {
                const double wgt = __ddiv_rn(wi[u], ws);                                                         // A6
                if (wgt > pb.wsig) rbq[(size_t)(i + u) * FRIQ_TILE] = __dadd_rn(qnow, __dmul_rn(qdiff, wgt));    // A7
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// merge_rb (frirl_agent.c:58-117, the reference's experimental knowledge-sharing mode): every rule of the sender, in order, is
// folded into the receiver's rule base.  One thread per receiver; the sender is a list of packed rule words and consequents
// with element stride `ss` (32 = a column of another agent's tile, 1 = a dense list).  wref is FIVERB.weights: written only
// when FIVE_vag_concl_weight finds no exact hit (:75-76) and read - possibly STALE - by update_rules (:47-55, :110).
// Library-exact chains; this is a per-chunk operation, not the hot loop.
// ---------------------------------------------------------------------------------------
template <int K, int P>
__device__ __noinline__ void friq_merge_rb(const friq_devprob &pb, uint32_t *rbw, double *rbq, double *sw, double *wref, int cap, int &R, int &flags,
                                           const uint32_t *sw_words, const double *sq_vals, int ss, int Rs)
{
    for (int r = 0; r < Rs; r++) {
        const uint32_t word = sw_words[(size_t)r * ss];
        const double sq = sq_vals[(size_t)r * ss];
        double vo[K], ant[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int j = (int)FRIQ_FIELD(word, k, K);
            vo[k] = pb.dict_ve[k][j]; ant[k] = pb.dict_val[k][j];
        }
        const int hit = friq_first_hit<K>(pb, rbw, R, vo);                       // FIVE_vag_concl_weight, :75-76
        double rc;
        if (hit >= 0) rc = rbq[(size_t)hit * FRIQ_TILE];                         // FIVE_vag_concl, :78-79
        else {
            double vagc, ws;
            friq_shepard<K, P>(pb, rbw, rbq, sw, R, vo, vagc, ws);
            rc = __ddiv_rn(vagc, ws);
            for (int i = 0; i < R; i++) wref[(size_t)i * FRIQ_TILE] = __ddiv_rn(sw[(size_t)i * FRIQ_TILE], ws);
        }
        const double qdiff = __dadd_rn(-rc, sq);                                 // :84
        if (qdiff > pb.qdiff_pos || qdiff < pb.qdiff_neg) {                      // :92
            int gi[K]; bool same = true; uint32_t packed = 0;
#pragma unroll
            for (int k = 0; k < K; k++) {                                        // CHECK_STATES, :96-99 + :19-45
                gi[k] = friq_snap_index(pb, k, ant[k]);
                same = same && (pb.dict_val[k][gi[k]] == ant[k]);
                packed |= (uint32_t)gi[k] << (FRIQ_BITS_OF(K) * k);
            }
            int hitg = hit; double rcg = rc;
            if (!same) {
                double vg[K], wsg;
#pragma unroll
                for (int k = 0; k < K; k++) vg[k] = pb.dict_ve[k][gi[k]];
                hitg = friq_vag_concl<K, P>(pb, rbw, rbq, nullptr, R, vg, rcg, wsg);
            }
            if (hitg < 0) {                                                      // :100-103
                if (R < cap) { rbw[(size_t)R * FRIQ_TILE] = packed; rbq[(size_t)R * FRIQ_TILE] = __dadd_rn(__dmul_rn(0.5, rcg), __dmul_rn(0.5, sq)); R++; }
                else flags |= FRIQ_F_OVERFLOW;
            } else rbq[(size_t)hitg * FRIQ_TILE] = __dadd_rn(__dmul_rn(0.9, rcg), __dmul_rn(0.1, sq));   // :105
            continue;
        }
        const double q = __dadd_rn(__dmul_rn(0.9, rc), __dmul_rn(0.1, sq));      // :110
        for (int w = 0; w < R; w++) {                                            // update_rules, :47-55
            const double wt = wref[(size_t)w * FRIQ_TILE];
            if (wt > pb.wsig) rbq[(size_t)w * FRIQ_TILE] = __dmul_rn(q, wt);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Environments (the example drivers' callbacks).  Expressions keep the reference's
// operation order; constant sub-expressions fold exactly as gcc folds them.
// ---------------------------------------------------------------------------------------
template <int ENV> __device__ __forceinline__ void friq_env_step(double action, const double *s, double *n);
template <int ENV> __device__ __forceinline__ void friq_env_reward(const double *s, double &reward, int &done);

template <> __device__ __forceinline__ void friq_env_step<FRIQ_ENV_MOUNTAINCAR>(double action, const double *s, double *n)
{   // mountaincar.c:37-74
    double v = __dmul_rn(__dadd_rn(__dadd_rn(s[1], __dmul_rn(0.001, action)), __dmul_rn(-0.0025, friq_cos(__dmul_rn(3.0, s[0])))), 0.999);
    if (v < -0.07) v = -0.07;
    if (v > 0.07) v = 0.07;
    double x = __dadd_rn(s[0], v);
    if (x <= -1.5) { x = -1.5; v = 0.0; }
    n[0] = x; n[1] = v;
}
template <> __device__ __forceinline__ void friq_env_reward<FRIQ_ENV_MOUNTAINCAR>(const double *s, double &reward, int &done)
{   // mountaincar.c:77-95
    if (s[0] >= 0.45) { reward = 1000.0; done = 1; } else { reward = -10.0; done = 0; }
}

template <> __device__ __forceinline__ void friq_env_step<FRIQ_ENV_ACROBOT>(double action, const double *s, double *n)
{   // acrobot.c:31-130 (accelerations once, then four Euler sub-steps - "Adam White's loop")
    const double max1 = 4 * FRIQ_PI, max2 = 9 * FRIQ_PI, g = 9.8, dt = 0.05;
    double t1 = s[0], t2 = s[1], w1 = s[2], w2 = s[3];
    double s2, c2;
    friq_sincos(t2, &s2, &c2);
    const double d1 = __dadd_rn(__dadd_rn(__dadd_rn(0.25, __dadd_rn(1.25, c2)), 1.0), 1.0);
    const double d2 = __dadd_rn(__dadd_rn(0.25, __dmul_rn(0.5, c2)), 1.0);
    const double phi2 = __dmul_rn(0.5 * g, friq_cos(__dsub_rn(__dadd_rn(t1, t2), FRIQ_PI / 2)));
    const double phi1 = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(-0.5, w2), s2), __dsub_rn(w2, __dmul_rn(2.0, w1))),
                                            __dmul_rn((1.0 * 0.5 + 1.0 * 1.0) * g, friq_cos(__dsub_rn(t1, FRIQ_PI / 2)))), phi2);
    double acc2 = __dsub_rn(__dsub_rn(__dadd_rn(action, __dmul_rn(phi1, __ddiv_rn(d2, d1))),
                                      __dmul_rn(__dmul_rn(__dmul_rn(0.5, w1), w1), s2)), phi2);
    acc2 = __ddiv_rn(acc2, __dsub_rn(1.25, __ddiv_rn(__dmul_rn(d2, d2), d1)));
    const double acc1 = __ddiv_rn(-__dadd_rn(__dmul_rn(d2, acc2), phi1), d1);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        w1 = __dadd_rn(w1, __dmul_rn(acc1, dt));
        if (w1 < -max1) w1 = -max1;
        if (w1 > max1) w1 = max1;
        t1 = __dadd_rn(t1, __dmul_rn(w1, dt));
        w2 = __dadd_rn(w2, __dmul_rn(acc2, dt));
        if (w2 < -max2) w2 = -max2;
        if (w2 > max2) w2 = max2;
        t2 = __dadd_rn(t2, __dmul_rn(w2, dt));
    }
    if (t1 < -FRIQ_PI) t1 = -FRIQ_PI;
    if (t1 > FRIQ_PI) t1 = FRIQ_PI;
    if (t2 < -FRIQ_PI) t2 = -FRIQ_PI;
    if (t2 > FRIQ_PI) t2 = FRIQ_PI;
    n[0] = t1; n[1] = t2; n[2] = w1; n[3] = w2;
}
template <> __device__ __forceinline__ void friq_env_reward<FRIQ_ENV_ACROBOT>(const double *s, double &reward, int &done)
{   // acrobot.c:133-162
    const double y1 = __dsub_rn(0.0, friq_cos(s[0]));
    const double y2 = __dsub_rn(y1, friq_cos(s[1]));
    if (y2 >= 1.0) { reward = 1000.0; done = 1; } else { reward = -10.0; done = 0; }
}

template <> __device__ __forceinline__ void friq_env_step<FRIQ_ENV_CARTPOLE>(double action, const double *s, double *n)
{   // cartpole.c:35-77 (Euler, tau = 0.02)
    const double g = 9.8, total_mass = 1.0 + 0.1, pml = 0.1 * 0.5, tau = 0.02, fourthirds = 4.0 / 3.0;
    const double x = s[0], xd = s[1], th = s[2], thd = s[3];
    double sn, cs;
    friq_sincos(th, &sn, &cs);
    const double force = __dmul_rn(action, 10.0);
    const double temp = __ddiv_rn(__dadd_rn(force, __dmul_rn(__dmul_rn(__dmul_rn(pml, thd), thd), sn)), total_mass);
    const double thacc = __ddiv_rn(__dsub_rn(__dmul_rn(g, sn), __dmul_rn(cs, temp)),
                                   __dmul_rn(0.5, __dsub_rn(fourthirds, __ddiv_rn(__dmul_rn(__dmul_rn(0.1, cs), cs), total_mass))));
    const double xacc = __dsub_rn(temp, __ddiv_rn(__dmul_rn(__dmul_rn(pml, thacc), cs), total_mass));
    n[0] = __dadd_rn(x, __dmul_rn(tau, xd));
    n[1] = __dadd_rn(xd, __dmul_rn(tau, xacc));
    n[2] = __dadd_rn(th, __dmul_rn(tau, thd));
    n[3] = __dadd_rn(thd, __dmul_rn(tau, thacc));
}
template <> __device__ __forceinline__ void friq_env_reward<FRIQ_ENV_CARTPOLE>(const double *s, double &reward, int &done)
{   // cartpole.c:79-112 ("success" is the failure-termination flag here)
    const double deg45 = FRIQ_PI / 4;
    const double x = s[0], th = s[2], thd = s[3];
    if (x < -4.0 || x > 4.0 || th < (-1 * deg45) || th > deg45) {
        reward = __dsub_rn(__dsub_rn(-10000.0, __dmul_rn(50.0, fabs(x))), __dmul_rn(100.0, fabs(th)));
        done = 1;
    } else {
        reward = __dsub_rn(__dsub_rn(__dsub_rn(10.0, __dmul_rn(__dmul_rn(1000.0, th), th)), __dmul_rn(5.0, fabs(x))), __dmul_rn(10.0, thd));
        done = 0;
    }
}

__device__ __noinline__ int friq_quantize_exact(double y, double div) { return (int)round(__ddiv_rn(y, div)); }   // mountaincar.c:107, acrobot.c:174
// quantise_observations: writes the quantised state q[] and its VE values vo[]
template <int ENV, int S>
__device__ __forceinline__ void friq_env_quantize(const friq_devprob &pb, const double *s, double *q, double *vo, int *qidx)
{
    if (ENV == FRIQ_ENV_CARTPOLE) {                      // cartpole.c:114-168
        const double deg12 = FRIQ_PI / 15, deg3 = FRIQ_PI / 60;
        double x0 = s[0], x1 = round(s[1]), x2 = __dmul_rn(floor(__ddiv_rn(s[2], deg3)), deg3), x3 = s[3];
        if (x0 < 0) x0 = -1;
        if (x0 > 0) x0 = 1;
        if (x1 < -1) x1 = -1;
        if (x1 > 1) x1 = 1;
        if (x2 > deg12) x2 = deg12;
        if (x2 < (-1 * deg12)) x2 = -1 * deg12;
        if (x3 < 0) x3 = -1;
        if (x3 > 0) x3 = 1;
        q[0] = x0; q[1] = x1; q[2] = x2; q[3] = x3;
#pragma unroll
        for (int k = 0; k < S; k++) { vo[k] = friq_obs_ve(pb, k, q[k]); qidx[k] = -1; }      // off-grid values
    } else {                                             // mountaincar.c:98-125, acrobot.c:165-192
#pragma unroll
        for (int k = 0; k < S; k++) {
            // where = (int)round((s + |values[0]|) / div), without the division: ta = y * RN(1 / div) is within 1.5 * 2^-52 relative of
            // RN(y / div), so both round to the same integer unless a half-integer boundary lies within 2^-49 |ta| of ta - only
            // then (never in practice) the exact division decides.  Negative quotients all clamp to index 0 whatever their rounding.
            const double y = __dadd_rn(s[k], pb.q_abs0[k]);
            const double ta = __dmul_rn(y, pb.q_rdiv[k]);
            const double fl = floor(ta), fr = __dsub_rn(ta, fl);
            int where = (int)fl + (fr >= 0.5 ? 1 : 0);
            if (fabs(__dsub_rn(fr, 0.5)) <= __dmul_rn(fabs(ta), 0x1p-49)) where = friq_quantize_exact(y, pb.q_div[k]);
            if (where < 0) where = 0; else if (where > pb.grid_len[k] - 1) where = pb.grid_len[k] - 1;
            q[k] = pb.dict_val[k][where];
            vo[k] = pb.dict_ve[k][where];                // = VE_k[nearest_index(values[where])]
            qidx[k] = where;
        }
    }
}
