// csrc/friq_engine.cu -- CUDA layer of the B200-native FRIQ-learning engine (sm_100a).
//
// Kernels
//   k_run_steps      the fused persistent step kernel: for every agent, n env steps of
//                    frirl_episode.c:86-186 (dynamics, reward, quantiser, best action over all
//                    actions, epsilon-greedy, SARSA update / rule insertion), agent state in
//                    registers, problem block staged to shared memory by TMA.  COOP = 2 is its
//                    TABLE MODE (mountaincar, acrobot): exact hits through a per-agent cell-to-rule
//                    map, Shepard weights looked up in the problem's memoised table.
//   k_build_wtab / k_refresh_index
//                    the table (once per problem) and the per-agent indexes (before every launch).
//   k_vag_concl / k_best_action / k_update_sarsa / k_classic_step / k_rule_distance / k_bestact_dists
//                    one call of FIVE_vag_concl(_weight) / frirl_get_best_action /
//                    frirl_update_sarsa ... per agent: the per-call boundary of the reference API
//                    (and the per-call parity tests); same device functions as the fused kernel.
//   k_infer_shared   inference against ONE shared rule base streamed through shared memory by TMA.
//   k_merge_list / k_team_merge / k_team_ctl
//                    knowledge sharing: merge_rb and the team bookkeeping of frirl_omp_run (frirl_agent.c).
//   k_retile_*       re-assignment of agents to warps by rule count (placement only).
//   k_init_agents    corner-rule bases + episode bookkeeping (frirl_init_rb.c, frirl_init.c).
// The extern "C" friqd_* functions are the thin C-ABI the ANSI-C host layer calls.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "friq_device.cuh"

#ifndef FRIQ_BLOCK
#define FRIQ_BLOCK 128
#endif
#ifndef FRIQ_UNR_MC
#define FRIQ_UNR_MC 3
#endif
#ifndef FRIQ_MINB
#define FRIQ_MINB 2
#endif
#ifndef FRIQ_MINB_COOP
#define FRIQ_MINB_COOP 2
#endif
#ifndef FRIQ_BLOCK_TAB
#define FRIQ_BLOCK_TAB 256   /* table mode (COOP = 2): one 8-warp CTA per SM (measured +2 % over two 4-warp CTAs: one problem block per SM) */
#endif
#ifndef FRIQ_MINB_TAB
#define FRIQ_MINB_TAB 1      /* 8 warps per SM at 254 registers; 12 and 16 warps spill and measured 9-35 % slower (4-rule blocks: 8 % slower) */
#endif
#ifndef FRIQ_COMPACT_UNR
#define FRIQ_COMPACT_UNR 6    /* rules per trip (= independent chains per lane) of friq_pairs_compacted */
#endif
#ifndef FRIQ_COOP_MAXPAIRS
#define FRIQ_COOP_MAXPAIRS 64 /* at most this many (lane, action) pairs of a warp without an exact hit: compacted Shepard sums (<= 2 rounds) */
#endif
#ifndef FRIQ_WUPD_LOOKUP_MAX
#define FRIQ_WUPD_LOOKUP_MAX 8 /* at most this many lanes of a warp with a weighted update pending: they look their weights up again instead of reading sw[] */
#endif
#ifndef FRIQ_QCOOP_MAX
#define FRIQ_QCOOP_MAX 8     /* fewer lanes than this needing Q(s,a): friq_q_cooperative instead of the WITHQ pass */
#endif

#ifdef FRIQ_TIMELINE   /* experiments only (tools/timeline.py): one record per warp of k_run_steps */
#define FRIQ_TL_MAX (1 << 18)
__device__ unsigned long long friq_tl_buf[FRIQ_TL_MAX][6];
__device__ unsigned int friq_tl_n;
__device__ __forceinline__ unsigned long long friq_globaltimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned int friq_smid() { unsigned int v; asm volatile("mov.u32 %0, %smid;" : "=r"(v)); return v; }
#endif

struct __align__(16) friq_smem_t {
    unsigned char prob[FRIQ_DEVPROB_SMEM_BYTES];
    uint64_t mbar;
};

#define FRIQ_SMEM_PROLOGUE_AT(BLOCK_INDEX)                                        \
    __shared__ friq_smem_t sm;                                                    \
    friq_stage_problem(gprob, sm.prob, &sm.mbar);                                 \
    const friq_devprob &pb = *reinterpret_cast<const friq_devprob *>(sm.prob);    \
    const long ag = (long)(BLOCK_INDEX) * blockDim.x + threadIdx.x;               \
    if (ag >= st.n) return;                                                       \
    const size_t tile_off = (size_t)(ag >> 5) * st.cap * FRIQ_TILE + (ag & 31);   \
    uint32_t *rbw = st.rbw + tile_off;                                            \
    double *rbq = st.rbq + tile_off;                                              \
    double *sw = st.sw + tile_off;
#define FRIQ_SMEM_PROLOGUE() FRIQ_SMEM_PROLOGUE_AT(blockIdx.x)

// ---------------------------------------------------------------------------------------
// COOP: exact hits by the integer sweep and cooperative Shepard sums ONLY (problems with hit_scan whose visited
// (observation, action) pairs mostly have a rule - mountaincar); the lock-step pass is not even compiled in, which halves the
// register count and doubles the resident warps of this then latency-bound kernel.
// COOP = 2: the same in TABLE MODE - exact hits through the agent's cell-to-rule map, Shepard weights looked up in the problem's
// memoised table (friq_devprob.wtab); the computed chain survives only as an out-of-line fall-back.
template <int S, int P, int A, int ACH, int ENV, int UNR, bool SAFE, int COOP>
__global__ void __launch_bounds__(COOP == 2 ? FRIQ_BLOCK_TAB : FRIQ_BLOCK, COOP == 2 ? FRIQ_MINB_TAB : COOP ? FRIQ_MINB_COOP : FRIQ_MINB)
k_run_steps(const friq_devprob *__restrict__ gprob, friq_devstate st, long n_steps, int mode)
{
    constexpr int K = S + 1;
    __shared__ unsigned char pair_list[(COOP == 2 ? FRIQ_BLOCK_TAB : FRIQ_BLOCK) / 32][32];   // friq_pairs_compacted, per warp
    __shared__ friq_round_smem round_smem[COOP == 2 ? FRIQ_BLOCK_TAB / 32 : 1];                // the same in table mode
#ifdef FRIQ_TIMELINE
    const unsigned long long tl_t0 = friq_globaltimer();
#endif
    FRIQ_SMEM_PROLOGUE();
    double *prevq = st.prevq ? st.prevq + tile_off : nullptr;
    double *wref = ((mode & FRIQ_M_KEEP_WEIGHTS) && st.wref) ? st.wref + tile_off : nullptr;

    // ---- agent state -> registers
    double s[S], qs[K], vo[K];
#pragma unroll
    for (int k = 0; k < S; k++) s[k] = st.s[(size_t)k * st.n_pad + ag];
#pragma unroll
    for (int k = 0; k < K; k++) { qs[k] = st.qs[(size_t)k * st.n_pad + ag]; vo[k] = st.vo[(size_t)k * st.n_pad + ag]; }
    int aidx = st.aidx[ag], flags = st.flags[ag], R = st.R[ag], step_in_ep = st.step_in_ep[ag], hitq = st.hitq[ag];
    uint32_t ep_num = st.ep_num[ag];
    double ep_return = st.ep_return[ag];
    double prev_reward = st.prev_reward[ag]; int prev_steps = st.prev_steps[ag], prev_numru = st.prev_numru[ag];
    double last_return = st.last_return[ag]; int last_steps = st.last_steps[ag], last_rules = st.last_rules[ag];
    unsigned long long total_steps = st.total_steps[ag], sum_rules = st.sum_rules[ag];
    const unsigned long long seed = st.seed[ag];
    const double alpha = st.alpha[ag], gamma = st.gamma[ag], epsilon = st.epsilon[ag];
    const bool learn = mode & FRIQ_M_LEARN, track = (mode & FRIQ_M_TRACK) && prevq;
    int qi[K];                                           // grid index of qs[k] when it is a grid value, else -1
#pragma unroll
    for (int k = 0; k < K; k++) {
        qi[k] = -1;
        for (int j = 0; j < pb.grid_len[k]; j++) if (pb.dict_val[k][j] == qs[k]) { qi[k] = j; break; }
    }
    // table mode: memoised weights (pb.wtab) + this agent's rule-cell list and cell-to-rule map
    constexpr bool tabmode = COOP == 2;                  // the host launches this variant only with a table and sized indexes
    unsigned short *rbc = tabmode ? st.rbc + (size_t)(ag >> 5) * friq_rbc_slots(st.cap) * FRIQ_TILE : nullptr;      // the TILE's rule cells
    unsigned short *cmap = tabmode ? st.cmap + (size_t)(ag >> 5) * st.idx_cells * FRIQ_TILE + (ag & 31) : nullptr;
    unsigned qrow = FRIQ_NO_ROW;                         // offset of the row of the pending update's point vo[] in the memoised table, if it is a dictionary cell
    if (tabmode) {
        int unk[S];
#pragma unroll
        for (int k = 0; k < S; k++) unk[k] = -1;
        const uint32_t w0 = friq_obs_word<S>(pb, vo, unk);
        int ah = -1;
        for (int j = 0; j < pb.dict_len[S]; j++) if (pb.dict_ve[S][j] == vo[S]) { ah = j; break; }
        if (w0 != FRIQ_NO_TARGET && ah >= 0) qrow = (unsigned)friq_cell<K>(pb, w0 | ((uint32_t)ah << (FRIQ_BITS_OF(K) * S))) * (unsigned)pb.wstride;
    }

    // All lanes of a warp iterate in lockstep (explicit reconvergence at the loop top): a lane that starts
    // an episode spends one iteration on the initial action selection (frirl_episode.c:78) while its
    // neighbours step, and a lane that has finished idles until the slowest lane of the warp is done.
    const unsigned wmask = __activemask();
    long done_steps = 0;
    bool running = true;
    unsigned c_pairs = 0, c_tab = 0, c_chain = 0, c_hit = 0, c_app = 0, c_wupd = 0;    // executed-work counters (friq_devstate.ctr)
#ifdef FRIQ_TIMELINE
    long long tl_cyc = 0; unsigned tl_rounds = 0, tl_pairs = 0, tl_walk = 0; const long long tl_k0 = clock64();
#endif
    while (__any_sync(wmask, running && done_steps < n_steps)) {
        const bool act = running && done_steps < n_steps && !((mode & FRIQ_M_STOP_ON_CONVERGE) && (flags & FRIQ_F_CONVERGED));
        const bool initial = flags & FRIQ_F_NEED_RESET;
        double cur[S], cq[S], vc[S], reward = 0.0; int done = 0, ci[S];
        if (act) {
            if (initial) {                               // frirl_episode.c:46-82
#pragma unroll
                for (int k = 0; k < S; k++) {
                    if (st.s0) { s[k] = qs[k] = st.s0[(size_t)k * st.n_pad + ag]; vo[k] = vc[k] = friq_obs_ve_cold(&pb, k, s[k]); ci[k] = -1; }
                    else { s[k] = qs[k] = pb.values_def[k]; vo[k] = vc[k] = pb.ve_def[k]; ci[k] = -2; }   // -2: dictionary index known to the host (def_hit)
                    cur[k] = cq[k] = s[k]; qi[k] = -1;
                }
                ep_return = 0.0; step_in_ep = 0; flags &= ~FRIQ_F_NEED_RESET;
                prev_numru = R;
                if (track) {
#pragma unroll 1
                    for (int i = 0; i < R; i++) prevq[(size_t)i * FRIQ_TILE] = rbq[(size_t)i * FRIQ_TILE];   // sequential_run.c:72
                }
            } else {                                     // frirl_episode.c:97-112
                friq_env_step<ENV>(pb.act_val[aidx], s, cur);
                friq_env_reward<ENV>(cur, reward, done);
                ep_return = __dadd_rn(ep_return, reward);
                friq_env_quantize<ENV, S>(pb, cur, cq, vc, ci);
                step_in_ep++;
            }
        } else {
#pragma unroll
            for (int k = 0; k < S; k++) { cur[k] = s[k]; cq[k] = qs[k]; vc[k] = vo[k]; ci[k] = qi[k]; }
            running = false;
        }
        // Q(s,a) of the pending update: its exact-hit rule is known from the previous pass (hitq >= 0), or
        // there is none (-1) and its Shepard sums ride along in this pass as one more chain.  -2 = unknown
        // (state set through the per-call API, or act_ve != act_ve_eval): scan for it.
        const bool upd = act && !initial && learn;
        const bool qp_reuse = tabmode || pb.qp_reuse;    // table mode is only launched for problems with qp_reuse
        if (upd && (hitq == -2 || !qp_reuse)) hitq = friq_first_hit_cold<K>(&pb, rbw, R, vo);
        const bool needq = upd && hitq < 0;
        double actconc[A], vagq, wsq; int hidx[A];
        const unsigned qm = __ballot_sync(wmask, needq);
        // exact hits first (BestAct.c:89-93): where distance 0 <=> equal dictionary indices, an integer sweep finds them, and only
        // the (lane, action) pairs without one need Shepard sums - cooperatively when few, else through the lock-step pass
        bool lockstep = true, kept_sw = true;
        uint32_t ow = FRIQ_NO_TARGET;
        unsigned arow[A];                                // table mode: offset of the row of (observation, action a) in the memoised table
#pragma unroll
        for (int a = 0; a < A; a++) arow[a] = FRIQ_NO_ROW;
        if constexpr (COOP || A == ACH) if (COOP || pb.hit_scan) {
            ow = friq_obs_word<S>(pb, vc, ci);
            uint32_t tgt[A]; unsigned nm[A]; int pairs = 0;
#pragma unroll
            for (int a = 0; a < A; a++) tgt[a] = (ow != FRIQ_NO_TARGET && pb.act_hit[a] >= 0) ? (ow | ((uint32_t)pb.act_hit[a] << (FRIQ_BITS_OF(K) * S))) : FRIQ_NO_TARGET;
            if (tabmode) {                               // exact hits in O(1): the first rule of the target's dictionary cell
                const int oc = (ow != FRIQ_NO_TARGET) ? friq_cell<S>(pb, ow) : 0;
#pragma unroll
                for (int a = 0; a < A; a++)
                {
                    const int ca = oc + pb.act_hit[a] * pb.wlin[S];
                    hidx[a] = (act && tgt[a] != FRIQ_NO_TARGET) ? (int)cmap[(size_t)ca * FRIQ_TILE] - 1 : -1;
                    if (tgt[a] != FRIQ_NO_TARGET) arow[a] = (unsigned)ca * (unsigned)pb.wstride;
                }
            } else
                friq_hit_scan<A>(rbw, act ? R : 0, tgt, hidx);
#pragma unroll
            for (int a = 0; a < A; a++) { nm[a] = __ballot_sync(wmask, act && hidx[a] < 0); pairs += __popc(nm[a]); if (act) { if (hidx[a] < 0) c_pairs++; else c_hit++; } }
            if (needq) c_pairs++; else if (upd) c_hit++;
            if (COOP || pairs <= FRIQ_COOP_MAXPAIRS) {
                lockstep = false;
#pragma unroll
                for (int a = 0; a < A; a++) actconc[a] = (hidx[a] >= 0) ? rbq[(size_t)hidx[a] * FRIQ_TILE] : 0.0;
                vagq = 0.0; wsq = 0.0;
                // weighted updates read the weights of Q(s,a): handed over through sw[] when MANY lanes update (their reads of a slot
                // coalesce), looked up again by the owner when few do (FRIQ_WUPD_LOOKUP_MAX); knowledge-sharing runs always keep them
                kept_sw = !tabmode || wref != nullptr || (__popc(qm) > FRIQ_WUPD_LOOKUP_MAX && (st.cap % FRIQ_CELL_VEC) == 0);   // kept weights are stored in whole blocks: a capacity that is no multiple of the block size looks them up again
#ifdef FRIQ_TIMELINE
                const long long tl_c0 = clock64();
                if (pairs | qm) { tl_rounds++; tl_pairs += pairs + __popc(qm); tl_walk += __reduce_max_sync(wmask, (act && (needq || hidx[0] < 0 || hidx[A - 1] < 0 || hidx[A / 2] < 0)) ? R : 0); }
#endif
                if (pairs | qm) friq_pairs_compacted<S, P, A, FRIQ_COMPACT_UNR, SAFE, tabmode>(pb, rbw, rbq, sw, rbc, R, vc, vo, arow, qrow, &round_smem[tabmode ? threadIdx.x >> 5 : 0], wmask, nm, qm, pair_list[threadIdx.x >> 5], actconc, vagq, wsq, c_tab, c_chain, kept_sw);
#ifdef FRIQ_TIMELINE
                tl_cyc += clock64() - tl_c0;
#endif
            }
        }
        if (!COOP && lockstep) {
            if (act) { c_chain += (unsigned)(R * A); if (!(COOP || A == ACH) || !pb.hit_scan) { c_pairs += A + (needq ? 1 : 0); } }
            if (needq) c_chain += (unsigned)R;
            if (__popc(qm) >= FRIQ_QCOOP_MAX)            // many lanes need Q(s,a): its chain rides in the pass
                friq_best_action_pass<S, P, A, ACH, true, UNR, SAFE>(pb, rbw, rbq, sw, act ? R : 0, vc, actconc, hidx, vo, needq, vagq, wsq);
            else {                                       // none or a few: the warp serves them one by one
                friq_best_action_pass<S, P, A, ACH, false, UNR, SAFE>(pb, rbw, rbq, sw, act ? R : 0, vc, actconc, hidx, vo, false, vagq, wsq);
                if (qm) friq_q_cooperative<K, P, SAFE>(pb, rbw, rbq, sw, R, vo, wmask, qm, vagq, wsq);
            }
        }
        const int a2 = friq_select_action<A>(pb, actconc, epsilon, seed, ep_num, (uint32_t)step_in_ep);
        const int next_hit = qp_reuse ? friq_pick_int<A>(hidx, a2) : -2;
        const unsigned qrow_upd = qrow;                  // the pending update's point (its weights are looked up again by the weighted update)
        if (act) qrow = (unsigned)friq_pick_int<A>(reinterpret_cast<const int *>(arow), a2);
        if (act && initial) { aidx = a2; qs[S] = pb.act_val[a2]; vo[S] = pb.act_ve_eval[a2]; qi[S] = a2; hitq = next_hit; }
        if (act && !initial) {
            sum_rules += (unsigned long long)R;
            if (learn) {                                 // frirl_update_sarsa.c:348-385
                double qp, qnow, ws = 0.0;
                if (qp_reuse) qp = friq_pick<A>(actconc, a2);         // Q(s',a') == actconc[a'] bit for bit
                else {
                    double vq[K];
#pragma unroll
                    for (int k = 0; k < S; k++) vq[k] = vc[k];
                    vq[S] = pb.act_ve_eval[a2];
                    friq_vag_concl_cold<K, P>(&pb, rbw, rbq, R, vq, &qp);
                }
                if (hitq >= 0) qnow = rbq[(size_t)hitq * FRIQ_TILE];
                else { ws = wsq; qnow = __ddiv_rn(vagq, wsq); }                               // FIVEVagConcl.c:302
                const int R0 = R;
                if constexpr (tabmode) {
                    const bool look = !kept_sw && qrow_upd != FRIQ_NO_ROW;
                    const friq_row_f64 row{pb.wtab + (look ? qrow_upd : 0u)};
                    friq_sarsa_apply<S, P, true, friq_row_f64>(pb, rbw, rbq, sw, st.cap, R, flags, qs, qi, reward, qp, hitq, qnow, ws, alpha, gamma, wref, &c_wupd,
                                                               look ? &row : nullptr, rbc);
                } else
                    friq_sarsa_apply<S, P, true>(pb, rbw, rbq, sw, st.cap, R, flags, qs, qi, reward, qp, hitq, qnow, ws, alpha, gamma, wref, &c_wupd);
                hitq = next_hit;
                c_app += (unsigned)(R - R0);
                if (R > R0 && tabmode) {                 // a rule was appended: keep the indexes current
                    const int cn = friq_cell<K>(pb, rbw[(size_t)R0 * FRIQ_TILE]);
                    unsigned short *e = rbc + friq_rbc_index(R0, (int)(ag & 31));
                    if ((R0 % FRIQ_CELL_VEC) == 0) {     // a new block: the zero column's cell and a finite consequent in its unused tail (friq_tab_acc),
                        const unsigned z = friq_zero_cell(pb), zz = z | (z << 16);     // and one more block of zero cells FRIQ_RBC_AHEAD blocks on
#if FRIQ_CELL_VEC == 8
                        const friq_cellvec vz = make_uint4(zz, zz, zz, zz);
#else
                        const friq_cellvec vz = make_uint2(zz, zz);
#endif
                        friq_cellvec v = vz; v.x = (unsigned)cn | (z << 16);
                        *reinterpret_cast<friq_cellvec *>(e) = v;
                        *reinterpret_cast<friq_cellvec *>(e + (size_t)FRIQ_RBC_AHEAD * FRIQ_TILE * FRIQ_CELL_VEC) = vz;
#pragma unroll
                        for (int u = 1; u < FRIQ_CELL_VEC; u++) if (R0 + u < st.cap) rbq[(size_t)(R0 + u) * FRIQ_TILE] = 0.0;
                    } else *e = (unsigned short)cn;
                    if (!cmap[(size_t)cn * FRIQ_TILE]) cmap[(size_t)cn * FRIQ_TILE] = (unsigned short)(R0 + 1);
                }
                if (R > R0 && hitq == -1) {              // a rule was appended: is it at the next step's (s,a)?
                    const uint32_t wnew = rbw[(size_t)R0 * FRIQ_TILE];
                    bool same = pb.dict_ve[S][FRIQ_FIELD(wnew, S, K)] == pb.act_ve_eval[a2];
#pragma unroll
                    for (int k = 0; k < S; k++) same = same && (pb.dict_ve[k][FRIQ_FIELD(wnew, k, K)] == vc[k]);
                    if (same) hitq = R0;
                }
            } else hitq = next_hit;
            // frirl_episode.c:163-185
#pragma unroll
            for (int k = 0; k < S; k++) { s[k] = cur[k]; qs[k] = cq[k]; vo[k] = vc[k]; qi[k] = ci[k]; }
            aidx = a2; qs[S] = pb.act_val[a2]; vo[S] = pb.act_ve_eval[a2]; qi[S] = a2;
            total_steps++; done_steps++;
            if (done || step_in_ep >= pb.max_steps) {
                // "RB complete" predicate, frirl_sequential_run.c:83-150
                if (learn && track && prev_numru == R && prev_steps == step_in_ep && ep_return > pb.reward_good_above && prev_reward == ep_return) {
                    flags |= FRIQ_F_EPENDED;                                  // :90
                    bool ok = true;
                    for (int i = 0; i < R && ok; i++)
                        if (fabs(__dsub_rn(rbq[(size_t)i * FRIQ_TILE], prevq[(size_t)i * FRIQ_TILE])) >= pb.qdiff_final_tol) ok = false;
                    if (ok) flags |= FRIQ_F_CONVERGED;
                }
                last_return = ep_return; last_steps = step_in_ep; last_rules = R;
                prev_reward = ep_return; prev_steps = step_in_ep;
                ep_num++;
                flags |= FRIQ_F_NEED_RESET;
                if (mode & FRIQ_M_STOP_AT_EPISODE_END) running = false;
            }
        }
        __syncwarp(wmask);
    }

    // ---- executed-work counters: one atomic per warp and counter
    if (st.ctr) {
        const unsigned v[7] = { c_pairs, c_tab, c_chain, c_hit, c_app, c_wupd, (unsigned)done_steps };
        const int lead = __ffs(wmask) - 1;
#pragma unroll
        for (int c = 0; c < 7; c++) {
            const unsigned t = __reduce_add_sync(wmask, v[c]);
            if ((int)(threadIdx.x & 31) == lead && t) atomicAdd(st.ctr + c, (unsigned long long)t);
        }
    }
    // ---- registers -> agent state
#pragma unroll
    for (int k = 0; k < S; k++) st.s[(size_t)k * st.n_pad + ag] = s[k];
#pragma unroll
    for (int k = 0; k < K; k++) { st.qs[(size_t)k * st.n_pad + ag] = qs[k]; st.vo[(size_t)k * st.n_pad + ag] = vo[k]; }
    st.aidx[ag] = aidx; st.flags[ag] = flags; st.R[ag] = R; st.step_in_ep[ag] = step_in_ep; st.hitq[ag] = hitq;
    st.ep_num[ag] = ep_num; st.ep_return[ag] = ep_return;
    st.prev_reward[ag] = prev_reward; st.prev_steps[ag] = prev_steps; st.prev_numru[ag] = prev_numru;
    st.last_return[ag] = last_return; st.last_steps[ag] = last_steps; st.last_rules[ag] = last_rules;
    st.total_steps[ag] = total_steps; st.sum_rules[ag] = sum_rules;
#ifdef FRIQ_TIMELINE
    {
        int rmax = R, rmin = R;
        for (int o = 16; o; o >>= 1) { rmax = max(rmax, __shfl_xor_sync(wmask, rmax, o)); rmin = min(rmin, __shfl_xor_sync(wmask, rmin, o)); }
        if ((threadIdx.x & 31) == 0) {
            const unsigned int i = atomicAdd(&friq_tl_n, 1u);
            if (i < FRIQ_TL_MAX) {
                friq_tl_buf[i][0] = tl_t0; friq_tl_buf[i][1] = friq_globaltimer();
                friq_tl_buf[i][2] = ((unsigned long long)friq_smid() << 32) | (unsigned long long)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
                friq_tl_buf[i][3] = ((unsigned long long)rmax << 32) | (unsigned long long)rmin | ((unsigned long long)(st.n & 0xffff) << 48);
                friq_tl_buf[i][4] = ((unsigned long long)(clock64() - tl_k0) << 32) | (unsigned long long)(tl_cyc & 0xffffffffll);     // kernel cycles | cycles inside friq_pairs_compacted
                friq_tl_buf[i][5] = ((unsigned long long)tl_rounds << 48) | ((unsigned long long)(tl_pairs & 0xffffff) << 24) | (unsigned long long)(tl_walk & 0xffffff);
            }
        }
    }
#endif
}

// ---------------------------------------------------------------------------------------
// k_run_steps_anyA: the fused step loop for problems on a built-in environment that are NOT one of the three example shapes -
// another action set (run-time A <= FRIQ_MAXA), other grids.  Same step semantics as k_run_steps (frirl_episode.c:86-186), plain
// evaluation: the best-action pass walks the rule base once per action (friq_action_pass), Q(s,a) and Q(s',a') are
// library-exact conclusions (FIVE_vag_concl), no exact-hit forwarding, no tables.  One thread per agent, lanes independent.
// ---------------------------------------------------------------------------------------
template <int S, int P, int ENV>
__global__ void __launch_bounds__(FRIQ_BLOCK, 512 / FRIQ_BLOCK)
k_run_steps_anyA(const friq_devprob *__restrict__ gprob, friq_devstate st, long n_steps, int mode)
{
    constexpr int K = S + 1;
    FRIQ_SMEM_PROLOGUE();
    const int A = pb.A;
    double *prevq = st.prevq ? st.prevq + tile_off : nullptr;
    double *wref = ((mode & FRIQ_M_KEEP_WEIGHTS) && st.wref) ? st.wref + tile_off : nullptr;
    double s[S], qs[K], vo[K];
#pragma unroll
    for (int k = 0; k < S; k++) s[k] = st.s[(size_t)k * st.n_pad + ag];
#pragma unroll
    for (int k = 0; k < K; k++) { qs[k] = st.qs[(size_t)k * st.n_pad + ag]; vo[k] = st.vo[(size_t)k * st.n_pad + ag]; }
    int aidx = st.aidx[ag], flags = st.flags[ag], R = st.R[ag], step_in_ep = st.step_in_ep[ag];
    uint32_t ep_num = st.ep_num[ag];
    double ep_return = st.ep_return[ag];
    double prev_reward = st.prev_reward[ag]; int prev_steps = st.prev_steps[ag], prev_numru = st.prev_numru[ag];
    double last_return = st.last_return[ag]; int last_steps = st.last_steps[ag], last_rules = st.last_rules[ag];
    unsigned long long total_steps = st.total_steps[ag], sum_rules = st.sum_rules[ag];
    const unsigned long long seed = st.seed[ag];
    const double alpha = st.alpha[ag], gamma = st.gamma[ag], epsilon = st.epsilon[ag];
    const bool learn = mode & FRIQ_M_LEARN, track = (mode & FRIQ_M_TRACK) && prevq;
    long done_steps = 0;
    bool running = true;
    while (running && done_steps < n_steps) {
        if ((mode & FRIQ_M_STOP_ON_CONVERGE) && (flags & FRIQ_F_CONVERGED)) break;
        const bool initial = flags & FRIQ_F_NEED_RESET;
        double cur[S], cq[S], vc[K], reward = 0.0; int done = 0, ci[S];
        if (initial) {                                   // frirl_episode.c:46-82
#pragma unroll
            for (int k = 0; k < S; k++) {
                s[k] = qs[k] = st.s0 ? st.s0[(size_t)k * st.n_pad + ag] : pb.values_def[k];
                vo[k] = vc[k] = st.s0 ? friq_obs_ve(pb, k, s[k]) : pb.ve_def[k];
                cur[k] = cq[k] = s[k];
            }
            ep_return = 0.0; step_in_ep = 0; flags &= ~FRIQ_F_NEED_RESET;
            prev_numru = R;
            if (track) for (int i = 0; i < R; i++) prevq[(size_t)i * FRIQ_TILE] = rbq[(size_t)i * FRIQ_TILE];   // sequential_run.c:72
        } else {                                         // frirl_episode.c:97-112
            friq_env_step<ENV>(pb.act_val[aidx], s, cur);
            friq_env_reward<ENV>(cur, reward, done);
            ep_return = __dadd_rn(ep_return, reward);
            friq_env_quantize<ENV, S>(pb, cur, cq, vc, ci);
            step_in_ep++;
        }
        // frirl_get_best_action (frirl_get_best_action.c:31-341): first index of the maximum (max.inl:16-28)
        int best = 0; double bv = 0.0;
        for (int a = 0; a < A; a++) {
            double conc; int hidx;
            friq_action_pass<S, P, 1, false>(pb, rbw, rbq, R, vc, pb.act_ve[a], conc, hidx);
            if (a == 0 || bv < conc) { bv = conc; best = a; }
        }
        int a2 = best;                                   // frirl_e_greedy_selection.c:21-37 with the Philox stream
        if (pb.no_random != 1 && epsilon != 0.0) {
            uint32_t r[4];
            friq_philox4x32_10(ep_num, (uint32_t)step_in_ep, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
            const double u1 = (double)((((uint64_t)r[1] << 32) | r[0]) >> 11) * 0x1p-53;
            const double u2 = (double)((((uint64_t)r[3] << 32) | r[2]) >> 11) * 0x1p-53;
            if (!(u1 > epsilon)) { const int ra = (int)floor(__dmul_rn(u2, (double)A)); a2 = ra > A - 1 ? A - 1 : ra; }
        }
        if (initial) { aidx = a2; qs[S] = pb.act_val[a2]; vo[S] = pb.act_ve_eval[a2]; }
        else {
            sum_rules += (unsigned long long)R;
            if (learn) {                                 // frirl_update_sarsa.c:348-385
                double qp, qnow, wsp, ws;
                vc[S] = pb.act_ve_eval[a2];
                friq_vag_concl<K, P>(pb, rbw, rbq, nullptr, R, vc, qp, wsp);                    // :355
                const int hitq = friq_vag_concl<K, P>(pb, rbw, rbq, sw, R, vo, qnow, ws);       // :356
                int qi[K];
#pragma unroll
                for (int k = 0; k < K; k++) qi[k] = -1;
                friq_sarsa_apply<S, P>(pb, rbw, rbq, sw, st.cap, R, flags, qs, qi, reward, qp, hitq, qnow, ws, alpha, gamma, wref);
            }
#pragma unroll
            for (int k = 0; k < S; k++) { s[k] = cur[k]; qs[k] = cq[k]; vo[k] = vc[k]; }
            aidx = a2; qs[S] = pb.act_val[a2]; vo[S] = pb.act_ve_eval[a2];
            total_steps++; done_steps++;
            if (done || step_in_ep >= pb.max_steps) {
                if (learn && track && prev_numru == R && prev_steps == step_in_ep && ep_return > pb.reward_good_above && prev_reward == ep_return) {
                    flags |= FRIQ_F_EPENDED;             // frirl_sequential_run.c:83-150
                    bool ok = true;
                    for (int i = 0; i < R && ok; i++)
                        if (fabs(__dsub_rn(rbq[(size_t)i * FRIQ_TILE], prevq[(size_t)i * FRIQ_TILE])) >= pb.qdiff_final_tol) ok = false;
                    if (ok) flags |= FRIQ_F_CONVERGED;
                }
                last_return = ep_return; last_steps = step_in_ep; last_rules = R;
                prev_reward = ep_return; prev_steps = step_in_ep;
                ep_num++;
                flags |= FRIQ_F_NEED_RESET;
                if (mode & FRIQ_M_STOP_AT_EPISODE_END) running = false;
            }
        }
    }
    if (st.ctr && done_steps) atomicAdd(st.ctr + FRIQ_CTR_STEPS, (unsigned long long)done_steps);
#pragma unroll
    for (int k = 0; k < S; k++) st.s[(size_t)k * st.n_pad + ag] = s[k];
#pragma unroll
    for (int k = 0; k < K; k++) { st.qs[(size_t)k * st.n_pad + ag] = qs[k]; st.vo[(size_t)k * st.n_pad + ag] = vo[k]; }
    st.aidx[ag] = aidx; st.flags[ag] = flags; st.R[ag] = R; st.step_in_ep[ag] = step_in_ep; st.hitq[ag] = -2;
    st.ep_num[ag] = ep_num; st.ep_return[ag] = ep_return;
    st.prev_reward[ag] = prev_reward; st.prev_steps[ag] = prev_steps; st.prev_numru[ag] = prev_numru;
    st.last_return[ag] = last_return; st.last_steps[ag] = last_steps; st.last_rules[ag] = last_rules;
    st.total_steps[ag] = total_steps; st.sum_rules[ag] = sum_rules;
}

// ---------------------------------------------------------------------------------------
// Shared-rule-base inference (BASELINE config 4: many agents, ONE large read-only rule base,
// reduction_state = 1 semantics of frirl_test_run.c:65-69: greedy stepping, no rule-base writes).
// The rule base is stored as rows of (VE value of each antecedent, consequent) = (K+1) f64 and streams
// through shared memory in tiles of FRIQ_SRB_TILE rules: one elected thread issues cp.async.bulk (TMA)
// into a 2-deep ring, completion on mbarriers; all threads (= agents) of the CTA consume the same tile
// with broadcast shared-memory reads.  One thread per agent, rules visited in order => bit-identical sums.
// ---------------------------------------------------------------------------------------
#define FRIQ_SRB_TILE 256

__device__ __forceinline__ void friq_tma_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *mbar)
{
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar), dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void friq_mbar_wait(uint64_t *mbar, uint32_t parity)
{
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mbar);
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}

// One pass of every thread's observation (VE values vc[]) over the shared rule base: conclusions of all A actions
// (FIVEVagConcl_FRIRL_BestAct.c:89-93, :212-217, :265).  CTA-wide: every thread of the CTA must call it (threads without an
// agent pass any finite vc[]); `parity` carries the ring's mbarrier phases from call to call.
template <int S, int P, int A, int UNR, bool SAFE>
__device__ __forceinline__ void friq_shared_pass(double (*tile)[FRIQ_SRB_TILE * (S + 2)], uint64_t *full, uint32_t &parity,
                                                 const double *__restrict__ srb, int Rs, const double *vc, const double *ave, double *actconc)
{
    constexpr int K = S + 1, RW = K + 1;
    const int ntiles = (Rs + FRIQ_SRB_TILE - 1) / FRIQ_SRB_TILE;
    double vagc[A], ws[A], hq[A];
    unsigned hitmask = 0;
#pragma unroll
    for (int a = 0; a < A; a++) { vagc[a] = 0.0; ws[a] = 0.0; hq[a] = 0.0; }
    if (threadIdx.x == 0)
        for (int t = 0; t < 2 && t < ntiles; t++) {
            const int cnt = min(FRIQ_SRB_TILE, Rs - t * FRIQ_SRB_TILE);
            friq_tma_load(tile[t], srb + (size_t)t * FRIQ_SRB_TILE * RW, (uint32_t)(cnt * RW * sizeof(double)), &full[t]);
        }
    for (int t = 0; t < ntiles; t++) {
        const int bsel = t & 1, cnt = min(FRIQ_SRB_TILE, Rs - t * FRIQ_SRB_TILE);
        friq_mbar_wait(&full[bsel], (parity >> bsel) & 1u);
        parity ^= 1u << bsel;
        const double *tb = tile[bsel];
#pragma unroll 1
        for (int i = 0; i < cnt; i += UNR) {
            constexpr int NC = UNR * A;              // all chains of the trip, evaluated stage by stage
            double sv[NC], wv[NC], dv[NC], q[UNR]; bool live[NC];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const double *r = tb + (size_t)min(i + u, cnt - 1) * RW;
                double ssum = 0.0;
#pragma unroll
                for (int k = 0; k < S; k++) {
                    const double d = __dsub_rn(vc[k], r[k]);
                    const double d2 = __dmul_rn(d, d);
                    ssum = (k == 0) ? d2 : __dadd_rn(ssum, d2);
                }
                q[u] = r[K];
#pragma unroll
                for (int a = 0; a < A; a++) {
                    const double da = __dsub_rn(ave[a], r[S]);
                    sv[u * A + a] = __dadd_rn(__dmul_rn(da, da), ssum);
                    live[u * A + a] = !friq_is_zero(sv[u * A + a]);
                }
            }
            unsigned bad = 0;
            friq_shepard_weights_n<P, SAFE, NC>(sv, wv, bad, live, dv);
            if (bad) {
#pragma unroll
                for (int n = 0; n < NC; n++)
                    if ((bad >> n) & 1u) wv[n] = SAFE ? friq_shepard_weight_from_sqrt<P>(dv[n]) : friq_shepard_weight_slow<P>(live[n] ? sv[n] : 1.0);
            }
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                if (u == 0 || i + u < cnt) {
#pragma unroll
                    for (int a = 0; a < A; a++) {
                        if (!live[u * A + a] && !((hitmask >> a) & 1u)) { hitmask |= 1u << a; hq[a] = q[u]; }
                        vagc[a] = __dadd_rn(vagc[a], __dmul_rn(wv[u * A + a], q[u]));
                        ws[a] = __dadd_rn(ws[a], wv[u * A + a]);
                    }
                }
            }
        }
        __syncthreads();                             // everyone is done with tile[bsel]
        if (threadIdx.x == 0 && t + 2 < ntiles) {
            const int c2 = min(FRIQ_SRB_TILE, Rs - (t + 2) * FRIQ_SRB_TILE);
            friq_tma_load(tile[bsel], srb + (size_t)(t + 2) * FRIQ_SRB_TILE * RW, (uint32_t)(c2 * RW * sizeof(double)), &full[bsel]);
        }
    }
#pragma unroll
    for (int a = 0; a < A; a++) actconc[a] = ((hitmask >> a) & 1u) ? hq[a] : __ddiv_rn(vagc[a], ws[a]);
}

#define FRIQ_SHARED_PROLOGUE()                                                                                          \
    __shared__ friq_smem_t sm;                                                                                        \
    __shared__ __align__(16) double tile[2][FRIQ_SRB_TILE * (S + 2)];                                                 \
    __shared__ uint64_t full[2];                                                                                      \
    friq_stage_problem(gprob, sm.prob, &sm.mbar);                                                                     \
    const friq_devprob &pb = *reinterpret_cast<const friq_devprob *>(sm.prob);                                        \
    if (threadIdx.x == 0) {                                                                                           \
        for (int j = 0; j < 2; j++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full[j]))); \
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");                                            \
    }                                                                                                                 \
    __syncthreads();                                                                                                  \
    const long ag = (long)blockIdx.x * blockDim.x + threadIdx.x;                                                      \
    const bool valid = ag < st.n;                                                                                     \
    const long ia = valid ? ag : 0;                                                                                   \
    uint32_t parity = 0;                                 /* bit j = phase of full[j] */

// The general stepping kernel: one full pass over the shared rule base per env step (policy 0).  Since the quantised
// observations of the built-in environments live on the grid, the default is k_infer_memo below; this kernel stays as the
// reference it is compared with (tests/test_gpu_parity.py) and for the roofline of the streaming pass itself.
template <int S, int P, int A, int ENV, int UNR, bool SAFE>
__global__ void __launch_bounds__(FRIQ_BLOCK, 2)
k_infer_shared(const friq_devprob *__restrict__ gprob, friq_devstate st, const double *__restrict__ srb, int Rs, long n_steps, int mode)
{
    constexpr int K = S + 1;
    FRIQ_SHARED_PROLOGUE();
    double s[S], qs[K], vo[K];                           // vo: VE values of qs[], kept like k_run_steps keeps them (a learning launch may follow)
#pragma unroll
    for (int k = 0; k < S; k++) s[k] = st.s[(size_t)k * st.n_pad + ia];
#pragma unroll
    for (int k = 0; k < K; k++) { qs[k] = st.qs[(size_t)k * st.n_pad + ia]; vo[k] = st.vo[(size_t)k * st.n_pad + ia]; }
    int aidx = st.aidx[ia], flags = st.flags[ia], step_in_ep = st.step_in_ep[ia];
    uint32_t ep_num = st.ep_num[ia];
    double ep_return = st.ep_return[ia], last_return = st.last_return[ia];
    int last_steps = st.last_steps[ia], last_rules = st.last_rules[ia];
    unsigned long long total_steps = st.total_steps[ia], sum_rules = st.sum_rules[ia];
    const unsigned long long seed = st.seed[ia];
    const double epsilon = st.epsilon[ia];
    double ave[A];
#pragma unroll
    for (int a = 0; a < A; a++) ave[a] = pb.act_ve[a];

    long done_steps = 0;
    bool running = valid;
    while (__syncthreads_or(running && done_steps < n_steps)) {
        const bool act = running && done_steps < n_steps;
        const bool initial = flags & FRIQ_F_NEED_RESET;
        double cur[S], cq[S], vc[S], reward = 0.0; int done = 0, ci[S];
        if (act) {
            if (initial) {
#pragma unroll
                for (int k = 0; k < S; k++) {
                    s[k] = qs[k] = st.s0 ? st.s0[(size_t)k * st.n_pad + ag] : pb.values_def[k];
                    vc[k] = st.s0 ? friq_obs_ve(pb, k, s[k]) : pb.ve_def[k];
                    cur[k] = cq[k] = s[k];
                }
                ep_return = 0.0; step_in_ep = 0; flags &= ~FRIQ_F_NEED_RESET;
            } else {
                friq_env_step<ENV>(pb.act_val[aidx], s, cur);
                friq_env_reward<ENV>(cur, reward, done);
                ep_return = __dadd_rn(ep_return, reward);
                friq_env_quantize<ENV, S>(pb, cur, cq, vc, ci);
                step_in_ep++;
            }
        } else {
#pragma unroll
            for (int k = 0; k < S; k++) { cur[k] = s[k]; cq[k] = qs[k]; vc[k] = 0.0; }
            running = false;
        }
        double actconc[A];
        friq_shared_pass<S, P, A, UNR, SAFE>(tile, full, parity, srb, Rs, vc, ave, actconc);
        const int a2 = friq_select_action<A>(pb, actconc, epsilon, seed, ep_num, (uint32_t)step_in_ep);
        if (act && initial) {
            aidx = a2; qs[S] = pb.act_val[a2]; vo[S] = pb.act_ve_eval[a2];
#pragma unroll
            for (int k = 0; k < S; k++) vo[k] = vc[k];
        }
        if (act && !initial) {
            sum_rules += (unsigned long long)Rs;
#pragma unroll
            for (int k = 0; k < S; k++) { s[k] = cur[k]; qs[k] = cq[k]; vo[k] = vc[k]; }
            aidx = a2; qs[S] = pb.act_val[a2]; vo[S] = pb.act_ve_eval[a2];
            total_steps++; done_steps++;
            if (done || step_in_ep >= pb.max_steps) {
                last_return = ep_return; last_steps = step_in_ep; last_rules = Rs;
                ep_num++;
                flags |= FRIQ_F_NEED_RESET;
                if (mode & FRIQ_M_STOP_AT_EPISODE_END) running = false;
            }
        }
    }
    if (!valid) return;
#pragma unroll
    for (int k = 0; k < S; k++) st.s[(size_t)k * st.n_pad + ag] = s[k];
#pragma unroll
    for (int k = 0; k < K; k++) { st.qs[(size_t)k * st.n_pad + ag] = qs[k]; st.vo[(size_t)k * st.n_pad + ag] = vo[k]; }
    st.aidx[ag] = aidx; st.flags[ag] = flags; st.step_in_ep[ag] = step_in_ep; st.hitq[ag] = -2;
    st.ep_num[ag] = ep_num; st.ep_return[ag] = ep_return;
    st.last_return[ag] = last_return; st.last_steps[ag] = last_steps; st.last_rules[ag] = last_rules;
    st.total_steps[ag] = total_steps; st.sum_rules[ag] = sum_rules;
}

// ---- Memoised conclusions (BASELINE config 4).  The shared rule base is read-only and, after the first step of an episode, an
// agent's observation is a GRID point (the quantisers of mountaincar.c:98-125 / acrobot.c:165-192 return values[where]): the
// conclusion of (observation cell, action) against the whole rule base is a constant of the batch.  k_shared_memo_build
// evaluates it once per cell and action - same chains, same rule order, so bit for bit what k_infer_shared computes on every
// step - and k_shared_start does the one pass left per agent: the conclusions at its (off-grid) episode start state.
// Stepping (k_infer_memo) is then dynamics + A table reads.
//
// memo[cell * A + a], cell = sum_k grid_index_k * stride_k (stride_0 = 1).  One WARP per (cell, action): the 32 lanes evaluate
// the weights of 32 consecutive rules in parallel (library-exact chain), then the sums are added in rule order (every lane adds
// the same shuffled values), stopping at the first rule at distance 0.
// conclusion of ONE (point, action) against the shared rule base, computed by a whole warp: every lane returns the value
template <int S, int P>
__device__ __forceinline__ double friq_warp_conclusion(const double *__restrict__ srb, int Rs, const double *vc, double ave)
{
    constexpr int K = S + 1, RW = K + 1;
    const int lane = threadIdx.x & 31;
    double vag = 0.0, ws = 0.0;
    for (int base = 0; base < Rs; base += 32) {
        const int i = base + lane; const bool in = i < Rs;
        const double *r = srb + (size_t)(in ? i : Rs - 1) * RW;
        double ssum = 0.0;
#pragma unroll
        for (int k = 0; k < S; k++) {
            const double d = __dsub_rn(vc[k], r[k]);
            const double d2 = __dmul_rn(d, d);
            ssum = (k == 0) ? d2 : __dadd_rn(ssum, d2);
        }
        const double da = __dsub_rn(ave, r[S]);
        const double sv = __dadd_rn(__dmul_rn(da, da), ssum);
        const double q = r[K];
        const bool zero = in && friq_is_zero(sv);
        const double w = zero ? 0.0 : friq_shepard_weight<P>(sv);
        const double wq = __dmul_rn(w, q);
        const unsigned zm = __ballot_sync(0xffffffffu, zero);
        if (zm) return __shfl_sync(0xffffffffu, q, __ffs(zm) - 1);                             // BestAct.c:89-93: first rule at distance 0
        const int cnt = min(32, Rs - base);
        for (int j = 0; j < cnt; j++) {
            vag = __dadd_rn(vag, __shfl_sync(0xffffffffu, wq, j));                             // :216
            ws = __dadd_rn(ws, __shfl_sync(0xffffffffu, w, j));                                // :217
        }
    }
    return __ddiv_rn(vag, ws);                                                                 // :265
}
template <int S, int P>
__global__ void __launch_bounds__(128)
k_shared_memo_build(const friq_devprob *__restrict__ gprob, const double *__restrict__ srb, int Rs, double *__restrict__ memo, int cells, int A)
{
    const long wi = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wi >= (long)cells * A) return;
    int cell = (int)(wi / A); const int a = (int)(wi % A);
    double vc[S];
#pragma unroll
    for (int k = 0; k < S; k++) { const int n = gprob->grid_len[k]; vc[k] = gprob->dict_ve[k][cell % n]; cell /= n; }
    const double c = friq_warp_conclusion<S, P>(srb, Rs, vc, gprob->act_ve[a]);
    if ((threadIdx.x & 31) == 0) memo[wi] = c;
}

// ---- Episode start states, de-duplicated.  The conclusions at a start state depend on it only through the nearest universe
// index of every dimension (five_rule_distance.c:75: VE_k[idx(x_k)]), and a universe has few points (41 for acrobot): a batch of
// 262,144 random start states holds at most 41 x 41 distinct index tuples in the two dimensions config 4 randomises.  Agents
// insert their tuple into a hash table (k_start_keys), the distinct tuples are numbered (k_start_compact), ONE warp per (tuple,
// action) evaluates the conclusion (k_start_unique) and every agent copies its tuple's values (k_start_scatter).  When more than
// n / FRIQ_START_DEDUPE_DIV tuples are distinct, the thread-per-agent streaming pass (k_shared_start) is cheaper and runs instead;
// both kernels read the count on the device and the wrong one leaves at once, so the host never waits for it.
#define FRIQ_START_DEDUPE_DIV 8
#define FRIQ_KEY_EMPTY 0xffffffffffffffffull
__device__ __forceinline__ unsigned long long friq_mix64(unsigned long long x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
    return x;
}
__global__ void k_start_keys(const friq_devprob *__restrict__ gprob, friq_devstate st, unsigned long long *table, unsigned tmask, int *slot_of_agent)
{
    const long ag = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ag >= st.n) return;
    const int S = gprob->S;
    unsigned long long key = 0;
    for (int k = 0; k < S; k++) {
        const double x = st.s0 ? st.s0[(size_t)k * st.n_pad + ag] : gprob->values_def[k];
        key |= (unsigned long long)friq_nearest_index(*gprob, k, x, gprob->U) << (16 * k);
    }
    unsigned h = (unsigned)friq_mix64(key) & tmask;
    for (;;) {
        const unsigned long long old = atomicCAS(table + h, FRIQ_KEY_EMPTY, key);
        if (old == FRIQ_KEY_EMPTY || old == key) break;
        h = (h + 1) & tmask;
    }
    slot_of_agent[ag] = (int)h;
}
__global__ void k_start_compact(const unsigned long long *__restrict__ table, unsigned tsize, int *rank_of_slot, int *slot_of_rank, int *count)
{
    const unsigned h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= tsize || table[h] == FRIQ_KEY_EMPTY) return;
    const int r = atomicAdd(count, 1);
    rank_of_slot[h] = r; slot_of_rank[r] = (int)h;
}
template <int S, int P>
__global__ void __launch_bounds__(128)
k_start_unique(const friq_devprob *__restrict__ gprob, const double *__restrict__ srb, int Rs, const unsigned long long *__restrict__ table,
               const int *__restrict__ slot_of_rank, const int *__restrict__ count, long n, double *__restrict__ uconc, int A)
{
    const int nu = *count;
    if ((long)nu * FRIQ_START_DEDUPE_DIV > n) return;                    // too many distinct tuples: k_shared_start serves this batch
    const long wi = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wi >= (long)nu * A) return;
    const int u = (int)(wi / A), a = (int)(wi % A);
    const unsigned long long key = table[slot_of_rank[u]];
    double vc[S];
#pragma unroll
    for (int k = 0; k < S; k++) vc[k] = __ldg(gprob->ve + (size_t)k * gprob->U + (int)((key >> (16 * k)) & 0xffffu));
    const double c = friq_warp_conclusion<S, P>(srb, Rs, vc, gprob->act_ve[a]);
    if ((threadIdx.x & 31) == 0) uconc[wi] = c;
}
__global__ void k_start_scatter(friq_devstate st, const int *__restrict__ slot_of_agent, const int *__restrict__ rank_of_slot, const int *__restrict__ count,
                                const double *__restrict__ uconc, double *__restrict__ start, int A)
{
    const long ag = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ag >= st.n || (long)(*count) * FRIQ_START_DEDUPE_DIV > st.n) return;
    const int u = rank_of_slot[slot_of_agent[ag]];
    for (int a = 0; a < A; a++) start[(size_t)a * st.n_pad + ag] = uconc[(size_t)u * A + a];
}

// conclusions of all actions at every agent's episode start state (frirl_episode.c:46-48, :78: the first action is chosen from
// the RAW start state): start[a * n_pad + agent]
template <int S, int P, int A, int UNR, bool SAFE>
__global__ void __launch_bounds__(FRIQ_BLOCK, 2)
k_shared_start(const friq_devprob *__restrict__ gprob, friq_devstate st, const double *__restrict__ srb, int Rs, double *__restrict__ start,
               const int *__restrict__ uniq_count)
{
    if (uniq_count && (long)(*uniq_count) * FRIQ_START_DEDUPE_DIV <= st.n) return;     // the de-duplicated path serves this batch
    FRIQ_SHARED_PROLOGUE();
    double vc[S], ave[A], actconc[A];
#pragma unroll
    for (int k = 0; k < S; k++) vc[k] = st.s0 ? friq_obs_ve(pb, k, st.s0[(size_t)k * st.n_pad + ia]) : pb.ve_def[k];
#pragma unroll
    for (int a = 0; a < A; a++) ave[a] = pb.act_ve[a];
    friq_shared_pass<S, P, A, UNR, SAFE>(tile, full, parity, srb, Rs, vc, ave, actconc);
    if (!valid) return;
#pragma unroll
    for (int a = 0; a < A; a++) start[(size_t)a * st.n_pad + ag] = actconc[a];
}

// greedy / epsilon-greedy stepping on memoised conclusions: per env step the dynamics, the quantiser and A table reads
template <int S, int A, int ENV>
__global__ void __launch_bounds__(FRIQ_BLOCK)
k_infer_memo(const friq_devprob *__restrict__ gprob, friq_devstate st, const double *__restrict__ memo, const double *__restrict__ start,
             int Rs, long n_steps, int mode)
{
    constexpr int K = S + 1;
    FRIQ_SMEM_PROLOGUE();
    (void)rbw; (void)rbq; (void)sw;
    double s[S], qs[K], vo[K];
#pragma unroll
    for (int k = 0; k < S; k++) s[k] = st.s[(size_t)k * st.n_pad + ag];
#pragma unroll
    for (int k = 0; k < K; k++) { qs[k] = st.qs[(size_t)k * st.n_pad + ag]; vo[k] = st.vo[(size_t)k * st.n_pad + ag]; }
    int aidx = st.aidx[ag], flags = st.flags[ag], step_in_ep = st.step_in_ep[ag];
    uint32_t ep_num = st.ep_num[ag];
    double ep_return = st.ep_return[ag], last_return = st.last_return[ag];
    int last_steps = st.last_steps[ag], last_rules = st.last_rules[ag];
    unsigned long long total_steps = st.total_steps[ag], sum_rules = st.sum_rules[ag];
    const unsigned long long seed = st.seed[ag];
    const double epsilon = st.epsilon[ag];
    int stride[S];
    { int m = 1;
#pragma unroll
      for (int k = 0; k < S; k++) { stride[k] = m; m *= pb.grid_len[k]; } }
    long done_steps = 0;
    bool running = true;
    while (running && done_steps < n_steps) {
        const bool initial = flags & FRIQ_F_NEED_RESET;
        double cur[S], cq[S], vc[S], actconc[A], reward = 0.0; int done = 0, ci[S];
        if (initial) {                                   // frirl_episode.c:46-82
#pragma unroll
            for (int k = 0; k < S; k++) {
                s[k] = qs[k] = st.s0 ? st.s0[(size_t)k * st.n_pad + ag] : pb.values_def[k];
                vc[k] = st.s0 ? friq_obs_ve_cold(&pb, k, s[k]) : pb.ve_def[k];
                cur[k] = cq[k] = s[k];
            }
#pragma unroll
            for (int a = 0; a < A; a++) actconc[a] = start[(size_t)a * st.n_pad + ag];
            ep_return = 0.0; step_in_ep = 0; flags &= ~FRIQ_F_NEED_RESET;
        } else {                                         // frirl_episode.c:97-112
            friq_env_step<ENV>(pb.act_val[aidx], s, cur);
            friq_env_reward<ENV>(cur, reward, done);
            ep_return = __dadd_rn(ep_return, reward);
            friq_env_quantize<ENV, S>(pb, cur, cq, vc, ci);
            step_in_ep++;
            int cell = 0;
#pragma unroll
            for (int k = 0; k < S; k++) cell += ci[k] * stride[k];
#pragma unroll
            for (int a = 0; a < A; a++) actconc[a] = __ldg(memo + (size_t)cell * A + a);
        }
        const int a2 = friq_select_action<A>(pb, actconc, epsilon, seed, ep_num, (uint32_t)step_in_ep);
#pragma unroll
        for (int k = 0; k < S; k++) vo[k] = vc[k];
        vo[S] = pb.act_ve_eval[a2];
        if (initial) { aidx = a2; qs[S] = pb.act_val[a2]; }
        else {
            sum_rules += (unsigned long long)Rs;
#pragma unroll
            for (int k = 0; k < S; k++) { s[k] = cur[k]; qs[k] = cq[k]; }
            aidx = a2; qs[S] = pb.act_val[a2];
            total_steps++; done_steps++;
            if (done || step_in_ep >= pb.max_steps) {
                last_return = ep_return; last_steps = step_in_ep; last_rules = Rs;
                ep_num++;
                flags |= FRIQ_F_NEED_RESET;
                if (mode & FRIQ_M_STOP_AT_EPISODE_END) running = false;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < S; k++) st.s[(size_t)k * st.n_pad + ag] = s[k];
#pragma unroll
    for (int k = 0; k < K; k++) { st.qs[(size_t)k * st.n_pad + ag] = qs[k]; st.vo[(size_t)k * st.n_pad + ag] = vo[k]; }
    st.aidx[ag] = aidx; st.flags[ag] = flags; st.step_in_ep[ag] = step_in_ep; st.hitq[ag] = -2;
    st.ep_num[ag] = ep_num; st.ep_return[ag] = ep_return;
    st.last_return[ag] = last_return; st.last_steps[ag] = last_steps; st.last_rules[ag] = last_rules;
    st.total_steps[ag] = total_steps; st.sum_rules[ag] = sum_rules;
}

// ---------------------------------------------------------------------------------------
// per-call kernels
// ---------------------------------------------------------------------------------------
template <int S, int P>
__global__ void __launch_bounds__(FRIQ_BLOCK, 512 / FRIQ_BLOCK)
k_vag_concl(const friq_devprob *__restrict__ gprob, friq_devstate st, const double *__restrict__ ant, double *conc, int *hit, double *weights)
{
    constexpr int K = S + 1;
    FRIQ_SMEM_PROLOGUE();
    double vo[K];
#pragma unroll
    for (int k = 0; k < K; k++) vo[k] = friq_obs_ve(pb, k, ant[(size_t)ag * K + k]);       // five_rule_distance.c:75
    const int R = st.R[ag];
    double c, ws;
    const int h = friq_vag_concl<K, P>(pb, rbw, rbq, sw, R, vo, c, ws);
    if (conc) conc[ag] = c;
    hit[ag] = h;
    if (weights && h < 0)                                                                   // FIVEVagConclWeight.c:153-175
        for (int i = 0; i < R; i++) weights[(size_t)ag * st.cap + i] = __ddiv_rn(sw[(size_t)i * FRIQ_TILE], ws);
}

template <int S, int P, int A, int ACH>
__global__ void __launch_bounds__(FRIQ_BLOCK, 512 / FRIQ_BLOCK)
k_best_action(const friq_devprob *__restrict__ gprob, friq_devstate st, const double *__restrict__ states, double *actconc_out, int *best)
{
    FRIQ_SMEM_PROLOGUE();
    (void)sw;
    double vo[S], actconc[A];
#pragma unroll
    for (int k = 0; k < S; k++) vo[k] = friq_obs_ve(pb, k, states[(size_t)ag * S + k]);    // get_best_action.c:60-63
    int hidx[A]; double vagq, wsq;
    friq_best_action_pass<S, P, A, ACH, false, 1, false>(pb, rbw, rbq, nullptr, st.R[ag], vo, actconc, hidx, vo, false, vagq, wsq);
#pragma unroll
    for (int a = 0; a < A; a++) actconc_out[(size_t)ag * A + a] = actconc[a];
    best[ag] = friq_argmax<A>(actconc);
}

template <int S, int P>
__global__ void __launch_bounds__(FRIQ_BLOCK, 512 / FRIQ_BLOCK)
k_update_sarsa(const friq_devprob *__restrict__ gprob, friq_devstate st, const double *__restrict__ qant, const double *__restrict__ reward,
               const double *__restrict__ curqant)
{
    constexpr int K = S + 1;
    FRIQ_SMEM_PROLOGUE();
    double qs[K], vq[K], vc[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        qs[k] = qant[(size_t)ag * K + k];
        vq[k] = friq_obs_ve(pb, k, qs[k]);
        vc[k] = friq_obs_ve(pb, k, curqant[(size_t)ag * K + k]);
    }
    int R = st.R[ag], flags = st.flags[ag];
    double qp, qnow, wsp, ws;
    friq_vag_concl<K, P>(pb, rbw, rbq, nullptr, R, vc, qp, wsp);                 // :355
    const int hitq = friq_vag_concl<K, P>(pb, rbw, rbq, sw, R, vq, qnow, ws);    // :356
    int qi[K];
#pragma unroll
    for (int k = 0; k < K; k++) qi[k] = -1;
    friq_sarsa_apply<S, P>(pb, rbw, rbq, sw, st.cap, R, flags, qs, qi, reward[ag], qp, hitq, qnow, ws, st.alpha[ag], st.gamma[ag], st.wref ? st.wref + tile_off : nullptr);
    st.R[ag] = R; st.flags[ag] = flags; st.hitq[ag] = -2;
}

// One step of the classic single-agent API (frirl_episode.c:148-160 with HOST environment callbacks):
// proposed action for the quantised observation (frirl_e_greedy_selection -> frirl_get_best_action unless
// the host already drew a random action, forced >= 0), then frirl_update_sarsa(q_ant, reward, cur_q_ant).
// io per agent: in  [0..K) q_ant, [K] reward, [K+1..K+1+S) cur_q_states, [K+1+S] forced action (-1 = greedy),
//                   [K+2+S] learn (reduction_state == 0)
//               out [0] action index, [1] rules, [2] fus_is_rule_inserted, [3] overflow, [4..4+A) actconc
template <int S, int P, int A, int ACH>
__global__ void __launch_bounds__(FRIQ_BLOCK, 512 / FRIQ_BLOCK)
k_classic_step(const friq_devprob *__restrict__ gprob, friq_devstate st, const double *__restrict__ in, double *__restrict__ out)
{
    constexpr int K = S + 1, NIN = K + S + 3, NOUT = 4 + A;
    FRIQ_SMEM_PROLOGUE();
    const double *x = in + (size_t)ag * NIN;
    double *o = out + (size_t)ag * NOUT;
    double qs[K], vq[K], vc[K], cq[K], actconc[A];
#pragma unroll
    for (int k = 0; k < K; k++) { qs[k] = x[k]; vq[k] = friq_obs_ve(pb, k, qs[k]); }
#pragma unroll
    for (int k = 0; k < S; k++) { cq[k] = x[K + 1 + k]; vc[k] = friq_obs_ve(pb, k, cq[k]); }
    const double reward = x[K];
    int a2 = (int)x[K + 1 + S];
    const bool learn = x[K + 2 + S] != 0.0;
    int R = st.R[ag], flags = st.flags[ag];
#pragma unroll
    for (int a = 0; a < A; a++) actconc[a] = 0.0;
    if (a2 < 0) {                                        // frirl_get_best_action.c:31-341
        int hidx[A]; double vagq, wsq;
        friq_best_action_pass<S, P, A, ACH, false, 1, false>(pb, rbw, rbq, nullptr, R, vc, actconc, hidx, vc, false, vagq, wsq);
        a2 = friq_argmax<A>(actconc);
    }
    if (learn) {                                         // frirl_update_sarsa.c:348-385
        cq[S] = pb.act_val[a2]; vc[S] = friq_obs_ve(pb, S, cq[S]);
        double qp, qnow, wsp, ws;
        friq_vag_concl<K, P>(pb, rbw, rbq, nullptr, R, vc, qp, wsp);
        const int hitq = friq_vag_concl<K, P>(pb, rbw, rbq, sw, R, vq, qnow, ws);
        int qi[K];
#pragma unroll
        for (int k = 0; k < K; k++) qi[k] = -1;
        friq_sarsa_apply<S, P>(pb, rbw, rbq, sw, st.cap, R, flags, qs, qi, reward, qp, hitq, qnow, ws, st.alpha[ag], st.gamma[ag], st.wref ? st.wref + tile_off : nullptr);
        st.R[ag] = R; st.flags[ag] = flags; st.hitq[ag] = -2;
    }
    o[0] = (double)a2; o[1] = (double)R; o[2] = (flags & FRIQ_F_INSERTED) ? 1.0 : 0.0; o[3] = (flags & FRIQ_F_OVERFLOW) ? 1.0 : 0.0;
#pragma unroll
    for (int a = 0; a < A; a++) o[4 + a] = actconc[a];
}

// ---- any number of actions (runtime A <= FRIQ_MAXA): the per-call kernels behind the drop-in API for problems that are not
// one of the three example shapes.  One action at a time (friq_action_pass with one rule per trip, library sqrt / rcp range
// checks on: nothing is assumed about the problem), conclusions written straight to global memory.
template <int S, int P>
__global__ void __launch_bounds__(FRIQ_BLOCK, 512 / FRIQ_BLOCK)
k_best_action_anyA(const friq_devprob *__restrict__ gprob, friq_devstate st, const double *__restrict__ states, double *actconc_out, int *best)
{
    FRIQ_SMEM_PROLOGUE();
    (void)sw;
    const int A = pb.A;
    double vo[S];
#pragma unroll
    for (int k = 0; k < S; k++) vo[k] = friq_obs_ve(pb, k, states[(size_t)ag * S + k]);    // get_best_action.c:60-63
    int b = 0; double bv = 0.0;
    for (int a = 0; a < A; a++) {
        double conc; int hidx;
        friq_action_pass<S, P, 1, false>(pb, rbw, rbq, st.R[ag], vo, pb.act_ve[a], conc, hidx);
        actconc_out[(size_t)ag * A + a] = conc;
        if (a == 0 || bv < conc) { bv = conc; b = a; }                                    // first maximum wins (max.inl:16-28)
    }
    best[ag] = b;
}

template <int S, int P>
__global__ void __launch_bounds__(FRIQ_BLOCK, 512 / FRIQ_BLOCK)
k_classic_step_anyA(const friq_devprob *__restrict__ gprob, friq_devstate st, const double *__restrict__ in, double *__restrict__ out)
{
    constexpr int K = S + 1, NIN = K + S + 3;
    FRIQ_SMEM_PROLOGUE();
    const int A = pb.A, NOUT = 4 + A;
    const double *x = in + (size_t)ag * NIN;
    double *o = out + (size_t)ag * NOUT;
    double qs[K], vq[K], vc[K], cq[K];
#pragma unroll
    for (int k = 0; k < K; k++) { qs[k] = x[k]; vq[k] = friq_obs_ve(pb, k, qs[k]); }
#pragma unroll
    for (int k = 0; k < S; k++) { cq[k] = x[K + 1 + k]; vc[k] = friq_obs_ve(pb, k, cq[k]); }
    const double reward = x[K];
    int a2 = (int)x[K + 1 + S];
    const bool learn = x[K + 2 + S] != 0.0;
    int R = st.R[ag], flags = st.flags[ag];
    for (int a = 0; a < A; a++) o[4 + a] = 0.0;
    if (a2 < 0) {
        double bv = 0.0;
        for (int a = 0; a < A; a++) {
            double conc; int hidx;
            friq_action_pass<S, P, 1, false>(pb, rbw, rbq, R, vc, pb.act_ve[a], conc, hidx);
            o[4 + a] = conc;
            if (a == 0 || bv < conc) { bv = conc; a2 = a; }
        }
    }
    if (learn) {
        cq[S] = pb.act_val[a2]; vc[S] = friq_obs_ve(pb, S, cq[S]);
        double qp, qnow, wsp, ws;
        friq_vag_concl<K, P>(pb, rbw, rbq, nullptr, R, vc, qp, wsp);
        const int hitq = friq_vag_concl<K, P>(pb, rbw, rbq, sw, R, vq, qnow, ws);
        int qi[K];
#pragma unroll
        for (int k = 0; k < K; k++) qi[k] = -1;
        friq_sarsa_apply<S, P>(pb, rbw, rbq, sw, st.cap, R, flags, qs, qi, reward, qp, hitq, qnow, ws, st.alpha[ag], st.gamma[ag], st.wref ? st.wref + tile_off : nullptr);
        st.R[ag] = R; st.flags[ag] = flags; st.hitq[ag] = -2;
    }
    o[0] = (double)a2; o[1] = (double)R; o[2] = (flags & FRIQ_F_INSERTED) ? 1.0 : 0.0; o[3] = (flags & FRIQ_F_OVERFLOW) ? 1.0 : 0.0;
}

// five_rule_distance (five_rule_distance.c:63-295): every rule's scaled distance from the point, and the first
// rule at distance 0 (-1: none).  All distances are written (the reference stops at the first exact hit and
// leaves the rest stale).
template <int S>
__global__ void __launch_bounds__(FRIQ_BLOCK, 512 / FRIQ_BLOCK)
k_rule_distance(const friq_devprob *__restrict__ gprob, friq_devstate st, const double *__restrict__ ant, double *__restrict__ dists, int *hit)
{
    constexpr int K = S + 1;
    FRIQ_SMEM_PROLOGUE();
    (void)sw; (void)rbq;
    double vo[K];
#pragma unroll
    for (int k = 0; k < K; k++) vo[k] = friq_obs_ve(pb, k, ant[(size_t)ag * K + k]);
    const int R = st.R[ag];
    int h = -1;
    for (int i = 0; i < R; i++) {
        const double d = __dsqrt_rn(friq_dist2<K>(pb, rbw[(size_t)i * FRIQ_TILE], vo));
        dists[(size_t)ag * st.cap + i] = d;
        if (d == 0.0 && h < 0) h = i;
    }
    hit[ag] = h;
}

// FIVEVagConcl_FRIRL_BestAct (FIVEVagConcl_FRIRL_BestAct.c:56-301) on caller-supplied rule distances
__global__ void k_bestact_dists(friq_devstate st, const double *__restrict__ dists, double *__restrict__ conc, int P)
{
    const long ag = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ag >= st.n) return;
    const double *rbq = st.rbq + (size_t)(ag >> 5) * st.cap * FRIQ_TILE + (ag & 31);
    const double *d = dists + (size_t)ag * st.cap;
    const int R = st.R[ag];
    for (int i = 0; i < R; i++) if (d[i] == 0.0) { conc[ag] = rbq[(size_t)i * FRIQ_TILE]; return; }   // :89-93
    double vagc = 0.0, ws = 0.0;
    for (int i = 0; i < R; i++) {                                                                    // :212-217
        const double wi = __drcp_rn(friq_pow_x87_rt(fabs(d[i]), P));
        vagc = __dadd_rn(vagc, __dmul_rn(wi, rbq[(size_t)i * FRIQ_TILE]));
        ws = __dadd_rn(ws, wi);
    }
    conc[ag] = __ddiv_rn(vagc, ws);                                                                  // :265
}

__global__ void k_init_agents(const friq_devprob *__restrict__ gprob, friq_devstate st, unsigned long long seed0, long first_agent,
                              double alpha, double gamma, double epsilon)
{
    const long ag = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ag >= st.n) return;
    const size_t tile_off = (size_t)(ag >> 5) * st.cap * FRIQ_TILE + (ag & 31);
    const int n = gprob->ncorner;
    for (int i = 0; i < n; i++) { st.rbw[tile_off + (size_t)i * FRIQ_TILE] = gprob->corner[i]; st.rbq[tile_off + (size_t)i * FRIQ_TILE] = 0.0; }
    for (int k = 0; k < FRIQ_MAXK; k++) { st.s[(size_t)k * st.n_pad + ag] = 0.0; st.qs[(size_t)k * st.n_pad + ag] = 0.0; st.vo[(size_t)k * st.n_pad + ag] = 0.0; }
    st.R[ag] = n; st.aidx[ag] = 0; st.flags[ag] = FRIQ_F_NEED_RESET; st.step_in_ep[ag] = 0; st.ep_num[ag] = 1u;
    st.ep_return[ag] = -1.0; st.prev_reward[ag] = -1.0; st.prev_steps[ag] = -1; st.prev_numru[ag] = n;     // frirl_init.c:144-145
    st.last_return[ag] = 0.0; st.last_steps[ag] = 0; st.last_rules[ag] = n;
    st.total_steps[ag] = 0ull; st.sum_rules[ag] = 0ull;
    st.seed[ag] = seed0 + (unsigned long long)(first_agent + ag);
    st.alpha[ag] = alpha; st.gamma[ag] = gamma; st.epsilon[ag] = epsilon;
    st.agent_of_slot[ag] = (int)ag; st.hitq[ag] = -2;
}

// derived indexes of the table mode (friq_devstate.rbc / .cmap), rebuilt from the packed rule words before every fused launch.
// The cell-to-rule map is cleared by a memset beforehand; rules are visited in DESCENDING order and stored unconditionally, so the
// entry that survives in a cell is its FIRST rule (a thread's stores to one address are ordered) and no load sits between stores.
__global__ void k_refresh_index(const friq_devprob *__restrict__ gprob, friq_devstate st)
{
    const long ag = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ag >= st.n) return;
    const int K = gprob->K, cells = st.idx_cells, lane = (int)(ag & 31), bits = FRIQ_BITS_OF(K);
    int wl[FRIQ_MAXK];
    for (int k = 0; k < FRIQ_MAXK; k++) wl[k] = k < K ? gprob->wlin[k] : 0;
    const uint32_t *rbw = st.rbw + (size_t)(ag >> 5) * st.cap * FRIQ_TILE + lane;
    unsigned short *cmap = st.cmap + (size_t)(ag >> 5) * cells * FRIQ_TILE + lane;
    unsigned short *rbc = st.rbc + (size_t)(ag >> 5) * friq_rbc_slots(st.cap) * FRIQ_TILE;
    const int R = st.R[ag], Rb = (int)friq_cap8(R);
    {   // FRIQ_RBC_AHEAD blocks of zero cells behind the column's last block
        const unsigned zz = (unsigned)cells | ((unsigned)cells << 16);
#if FRIQ_CELL_VEC == 8
        const friq_cellvec vz = make_uint4(zz, zz, zz, zz);
#else
        const friq_cellvec vz = make_uint2(zz, zz);
#endif
        for (int k = 0; k < FRIQ_RBC_AHEAD; k++) *reinterpret_cast<friq_cellvec *>(rbc + friq_rbc_index(Rb + k * FRIQ_CELL_VEC, lane)) = vz;
    }
    for (int i0 = Rb - FRIQ_CELL_VEC; i0 >= 0; i0 -= FRIQ_CELL_VEC) {
        uint32_t w[FRIQ_CELL_VEC];
#pragma unroll
        for (int u = 0; u < FRIQ_CELL_VEC; u++) w[u] = (i0 + u < R) ? rbw[(size_t)(i0 + u) * FRIQ_TILE] : 0u;
        unsigned c[FRIQ_CELL_VEC];
#pragma unroll
        for (int u = 0; u < FRIQ_CELL_VEC; u++) {
            int cc = 0;
#pragma unroll
            for (int k = 0; k < FRIQ_MAXK; k++) cc += (int)((w[u] >> ((bits * k) & 31)) & ((1u << bits) - 1u)) * wl[k];
            c[u] = (i0 + u < R && cc < cells) ? (unsigned)cc : (unsigned)cells;   // unused tail of the last block: the table's zero column (friq_zero_cell)
        }
        friq_cellvec v;
#if FRIQ_CELL_VEC == 8
        v = make_uint4(c[0] | (c[1] << 16), c[2] | (c[3] << 16), c[4] | (c[5] << 16), c[6] | (c[7] << 16));
#else
        v = make_uint2(c[0] | (c[1] << 16), c[2] | (c[3] << 16));
#endif
        *reinterpret_cast<friq_cellvec *>(rbc + friq_rbc_index(i0, lane)) = v;
#pragma unroll
        for (int u = FRIQ_CELL_VEC - 1; u >= 0; u--) {
            if (i0 + u < R) cmap[(size_t)c[u] * FRIQ_TILE] = (unsigned short)(i0 + u + 1);
            else if (i0 + u < st.cap) st.rbq[((size_t)(ag >> 5) * st.cap + (size_t)(i0 + u)) * FRIQ_TILE + lane] = 0.0;   // a finite consequent under the zero weight
        }
    }
}

// every agent abandons its current episode and starts a new one at its next step (rule bases, counters and RNG keys untouched)
__global__ void k_new_episodes(friq_devstate st)
{
    const long ag = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ag < st.n) st.flags[ag] |= FRIQ_F_NEED_RESET;
}
// per-agent result statistics of a gathered result set: out[0] = sum of rules, out[1] = sum of returns, out[2] = max rules.
// ONE block, fixed summation tree: deterministic for a given n.
__global__ void __launch_bounds__(1024) k_result_stats(const double *__restrict__ ret, const int *__restrict__ rules, long n, double *__restrict__ out)
{
    __shared__ double s_ret[1024], s_rules[1024]; __shared__ int s_max[1024];
    double a = 0.0, r = 0.0; int m = 0;
    for (long i = threadIdx.x; i < n; i += 1024) { a += ret[i]; r += (double)rules[i]; m = max(m, rules[i]); }
    s_ret[threadIdx.x] = a; s_rules[threadIdx.x] = r; s_max[threadIdx.x] = m;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) { s_ret[threadIdx.x] += s_ret[threadIdx.x + w]; s_rules[threadIdx.x] += s_rules[threadIdx.x + w]; s_max[threadIdx.x] = max(s_max[threadIdx.x], s_max[threadIdx.x + w]); }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = s_rules[0]; out[1] = s_ret[0]; out[2] = (double)s_max[0]; }
}
// per-agent episode start states from a device array [n][S] (agent-major, as a caller holds them) into the engine's [k][n_pad] rows
__global__ void k_set_s0(friq_devstate st, const double *__restrict__ src, int S)
{
    const long ag = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ag >= st.n) return;
    for (int k = 0; k < S; k++) st.s0[(size_t)k * st.n_pad + ag] = src[(size_t)ag * S + k];
}

__global__ void k_broadcast_rb(friq_devstate st, const unsigned int *__restrict__ words, const double *__restrict__ q, int R)
{
    const long ag = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ag >= st.n) return;
    const size_t tile_off = (size_t)(ag >> 5) * st.cap * FRIQ_TILE + (ag & 31);
    for (int i = 0; i < R; i++) { st.rbw[tile_off + (size_t)i * FRIQ_TILE] = words[i]; st.rbq[tile_off + (size_t)i * FRIQ_TILE] = q[i]; }
    st.R[ag] = R; st.hitq[ag] = -2;
}

// ---------------------------------------------------------------------------------------
// Re-tiling: agents are re-assigned to slots in descending order of their rule count so that the
// 32 lanes of a warp walk rule bases of (nearly) equal length - the rule loop is the hot loop and
// its trip count is the per-lane rule count.  Placement never changes an agent's results; slot p
// remembers which agent it holds (agent_of_slot) so results are reported in agent order.
// work[0..cap] = histogram / running offsets, work[cap+1 ..] = new slot of every old slot.
// by_rules == 0 restores the identity placement.
// ---------------------------------------------------------------------------------------
// one atomic per distinct rule count of a warp (after the first sort the lanes of a warp hold nearly equal counts)
__global__ void k_retile_hist(friq_devstate st, int *work)
{
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= st.n) return;
    const int r = st.R[p];
    const unsigned peers = __match_any_sync(__activemask(), r);
    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&work[r], __popc(peers));
}
// exclusive prefix sums of the histogram in DESCENDING rule-count order (one block): work[r] = #agents with more than r rules
__global__ void k_retile_scan(int *work, int cap)
{
    __shared__ int seg[256];
    const int n = cap + 1, t = threadIdx.x, len = (n + 255) / 256;
    const int hi = cap - t * len, lo = max(hi - len + 1, 0);          // thread t owns counts [lo, hi], thread 0 the largest
    int sum = 0;
    for (int r = hi; r >= lo; r--) sum += work[r];
    seg[t] = sum;
    __syncthreads();
    if (t == 0) { int off = 0; for (int i = 0; i < 256; i++) { const int c = seg[i]; seg[i] = off; off += c; } }
    __syncthreads();
    int off = seg[t];
    for (int r = hi; r >= lo; r--) { const int c = work[r]; work[r] = off; off += c; }
}
__global__ void k_retile_assign(friq_devstate st, int *work, int by_rules)
{
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= st.n) return;
    if (!by_rules) { work[st.cap + 1 + p] = st.agent_of_slot[p]; return; }
    const int r = st.R[p], lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(__activemask(), r);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(&work[r], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    work[st.cap + 1 + p] = base + __popc(peers & ((1u << lane) - 1u));
}
// FRIQ_MOVE_SLICES threads per agent: slice 0 moves the scalars, every slice its share of the rules (blockDim.x = 32 agents, blockDim.y = slices)
#define FRIQ_MOVE_SLICES 8
__global__ void k_retile_move(friq_devstate a, friq_devstate b, const int *work)
{
    const long p = (long)blockIdx.x * FRIQ_TILE + threadIdx.x;
    if (p >= a.n) return;
    const long q = work[a.cap + 1 + p];
    const int sl = threadIdx.y;
    if (sl == 0) {
        for (int k = 0; k < FRIQ_MAXK; k++) {
            b.s[(size_t)k * b.n_pad + q] = a.s[(size_t)k * a.n_pad + p];
            b.qs[(size_t)k * b.n_pad + q] = a.qs[(size_t)k * a.n_pad + p];
            b.vo[(size_t)k * b.n_pad + q] = a.vo[(size_t)k * a.n_pad + p];
        }
        b.aidx[q] = a.aidx[p]; b.flags[q] = a.flags[p]; b.R[q] = a.R[p]; b.step_in_ep[q] = a.step_in_ep[p];
        b.last_steps[q] = a.last_steps[p]; b.last_rules[q] = a.last_rules[p]; b.prev_steps[q] = a.prev_steps[p]; b.prev_numru[q] = a.prev_numru[p];
        b.ep_num[q] = a.ep_num[p]; b.ep_return[q] = a.ep_return[p]; b.last_return[q] = a.last_return[p]; b.prev_reward[q] = a.prev_reward[p];
        b.total_steps[q] = a.total_steps[p]; b.sum_rules[q] = a.sum_rules[p]; b.seed[q] = a.seed[p];
        b.alpha[q] = a.alpha[p]; b.gamma[q] = a.gamma[p]; b.epsilon[q] = a.epsilon[p];
        b.agent_of_slot[q] = a.agent_of_slot[p]; b.hitq[q] = a.hitq[p];
        if (a.s0 && b.s0) for (int k = 0; k < FRIQ_MAXK; k++) b.s0[(size_t)k * b.n_pad + q] = a.s0[(size_t)k * a.n_pad + p];
    }
    const size_t po = (size_t)(p >> 5) * a.cap * FRIQ_TILE + (p & 31), qo = (size_t)(q >> 5) * b.cap * FRIQ_TILE + (q & 31);
    const int R = a.R[p];
    const bool pv = a.prevq && b.prevq;
    for (int i = sl; i < R; i += FRIQ_MOVE_SLICES) {
        b.rbw[qo + (size_t)i * FRIQ_TILE] = a.rbw[po + (size_t)i * FRIQ_TILE];
        b.rbq[qo + (size_t)i * FRIQ_TILE] = a.rbq[po + (size_t)i * FRIQ_TILE];
        if (pv) b.prevq[qo + (size_t)i * FRIQ_TILE] = a.prevq[po + (size_t)i * FRIQ_TILE];
    }
}
template <typename T>
__global__ void k_unpermute(friq_devstate st, const T *src, T *dst)
{
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < st.n) dst[st.agent_of_slot[p]] = src[p];
}
// the other direction: per-agent inputs in AGENT order (src0..2, any may be NULL) into the slots the agents live in
__global__ void k_permute_in3(friq_devstate st, const double *src0, double *dst0, const double *src1, double *dst1, const double *src2, double *dst2)
{
    const long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= st.n) return;
    const int a = st.agent_of_slot[p];
    if (src0) dst0[p] = src0[a];
    if (src1) dst1[p] = src1[a];
    if (src2) dst2[p] = src2[a];
}

// ---- knowledge sharing (frirl_agent.c:58-117, 294-385)
// one thread per receiver; receiver < 0: every agent.  The sender is a dense list (d_words, d_q).
template <int S, int P>
__global__ void __launch_bounds__(FRIQ_BLOCK, 512 / FRIQ_BLOCK)
k_merge_list(const friq_devprob *__restrict__ gprob, friq_devstate st, long receiver, const unsigned int *d_words, const double *d_q, int Rs)
{
    FRIQ_SMEM_PROLOGUE();
    if (receiver >= 0 && ag != receiver) return;
    int R = st.R[ag], flags = st.flags[ag];
    friq_merge_rb<S + 1, P>(pb, rbw, rbq, sw, st.wref + tile_off, st.cap, R, flags, d_words, d_q, 1, Rs);
    st.R[ag] = R; st.flags[ag] = flags; st.hitq[ag] = -2;
}
// teams of `team` consecutive agents, member 0 = leader.  phase 0: the leader's rules into every other member (skipped when
// the leader's chunk "pended", frirl_agent.c:338-349); phase 1: every other member's rules, in member order, into the leader
// (each skipped when that member pended, :353-363).  Finished teams are left alone.
template <int S, int P>
__global__ void __launch_bounds__(FRIQ_BLOCK, 512 / FRIQ_BLOCK)
k_team_merge(const friq_devprob *__restrict__ gprob, friq_devstate st, int team, int phase)
{
    FRIQ_SMEM_PROLOGUE();
    const long lead = ag - ag % team;
    if (lead + team > st.n) return;                          // an incomplete trailing team takes no part
    if (st.flags[lead] & FRIQ_F_TEAM_DONE) return;
    int R = st.R[ag], flags = st.flags[ag];
    double *wref = st.wref + tile_off;
    if (phase == 0) {
        if (ag == lead || (st.flags[lead] & FRIQ_F_EPENDED)) return;
        const size_t so = (size_t)(lead >> 5) * st.cap * FRIQ_TILE + (lead & 31);
        friq_merge_rb<S + 1, P>(pb, rbw, rbq, sw, wref, st.cap, R, flags, st.rbw + so, st.rbq + so, FRIQ_TILE, st.R[lead]);
    } else {
        if (ag != lead) return;
        for (int id = 1; id < team; id++) {
            const long m = lead + id;
            if (st.flags[m] & FRIQ_F_EPENDED) continue;
            const size_t so = (size_t)(m >> 5) * st.cap * FRIQ_TILE + (m & 31);
            friq_merge_rb<S + 1, P>(pb, rbw, rbq, sw, wref, st.cap, R, flags, st.rbw + so, st.rbq + so, FRIQ_TILE, st.R[m]);
        }
    }
    st.R[ag] = R; st.flags[ag] = flags; st.hitq[ag] = -2;
}
// op 0, before a chunk: members of running teams forget "converged" / "pended" (frirl_sequential_run.c:34-36), members of finished
// teams are frozen (converged, so FRIQ_M_STOP_ON_CONVERGE skips them).  op 1, after a chunk: a team whose leader found its rule base
// complete (or ran out of episodes, :59-64) is finished; *count = teams still running.
__global__ void k_team_ctl(friq_devstate st, int team, int op, int max_episodes, int *count)
{
    const long ag = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ag >= st.n) return;
    const long lead = ag - ag % team;
    if (lead + team > st.n) { if (op == 0) st.flags[ag] |= FRIQ_F_CONVERGED | FRIQ_F_TEAM_DONE; return; }
    if (op == 0) {
        if (st.flags[lead] & FRIQ_F_TEAM_DONE) st.flags[ag] |= FRIQ_F_CONVERGED;
        else st.flags[ag] &= ~(FRIQ_F_CONVERGED | FRIQ_F_EPENDED);
    } else if (ag == lead && !(st.flags[ag] & FRIQ_F_TEAM_DONE)) {
        if ((st.flags[ag] & FRIQ_F_CONVERGED) || st.ep_num[ag] >= (unsigned)max_episodes) st.flags[ag] |= FRIQ_F_TEAM_DONE;
        else atomicAdd(count, 1);
    }
}

// memoised Shepard weights (friq_devprob.wtab): wtab[q * cells + r] = 1 / d(q, r)^p for every pair of dictionary cells, through the
// same device functions as a computed chain (friq_dist2 left-to-right over the dimensions, sqrt, x87 power, reciprocal)
template <int K, int P>
__global__ void k_build_wtab(const friq_devprob *__restrict__ gprob, double *__restrict__ wtab, int cells, int stride)
{
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)cells * cells) return;
    const int q0 = (int)(t / cells), r0 = (int)(t % cells);
    int q = q0, r = r0;
    double pt[K]; uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int n = gprob->dict_len[k];
        pt[k] = gprob->dict_ve[k][q % n]; q /= n;
        w |= (uint32_t)(r % n) << (FRIQ_BITS_OF(K) * k); r /= n;
    }
    wtab[(size_t)q0 * stride + r0] = friq_shepard_weight<P>(friq_dist2<K>(*gprob, w, pt));
}

__global__ void k_trig(const double *x, double *s, double *c, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) friq_sincos(x[i], &s[i], &c[i]);
}
__global__ void k_pow(const double *x, double *out, long n, int p, int force_exact)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (force_exact == 0) out[i] = friq_pow_x87_rt(x[i], p);
    else if (force_exact == 1) out[i] = friq_pow_x87_exact(x[i], p);
    else {                                   // Shepard weight 1/pow(sqrt(s)): 2 = library ops, 3 = branch-free fast paths
        unsigned bad = 0; double w;
        if (force_exact == 2) w = __drcp_rn(friq_pow_x87_exact(__dsqrt_rn(x[i]), p));
        else {
            w = (p == 3) ? friq_shepard_weight_fast<3, false>(x[i], bad) : friq_shepard_weight_fast<5, false>(x[i], bad);
            if (bad) w = (p == 3) ? friq_shepard_weight<3>(x[i]) : friq_shepard_weight<5>(x[i]);
            if (force_exact == 4) w = (double)bad;
        }
        out[i] = w;
    }
}

// FP64 issue-rate micro-benchmark: 8 independent DADD/DMUL chains per thread (no FMA, like the path)
__global__ void k_fp64_peak(double *out, int iters)
{
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = 1.0 + 1e-9 * (threadIdx.x + j);
    const double m = 1.0000000001, c = 1e-12;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) { a[j] = __dmul_rn(a[j], m); a[j] = __dadd_rn(a[j], c); }
    }
    double r = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) r += a[j];
    if (r == 123.456) out[0] = r;
}

// ---------------------------------------------------------------------------------------
// C-ABI
// ---------------------------------------------------------------------------------------
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

static inline dim3 grid_for(long n) { return dim3((unsigned)((n + FRIQ_BLOCK - 1) / FRIQ_BLOCK)); }

// supported problem shapes (S, p, A): the three example drivers.  ENV_NONE problems of the same
// shape use the per-call kernels only.
#define FRIQ_SHAPE_MC(pb) ((pb)->S == 2 && (pb)->p == 3 && (pb)->A == 3)
#define FRIQ_SHAPE_AC(pb) ((pb)->S == 4 && (pb)->p == 5 && (pb)->A == 3)
#define FRIQ_SHAPE_CP(pb) ((pb)->S == 4 && (pb)->p == 5 && (pb)->A == 21)
#define FRIQ_ERR_SHAPE 9001

extern "C" {

int friqd_set_device(int device) { CK(cudaSetDevice(device)); return 0; }
int friqd_malloc(void **p, unsigned long bytes) { CK(cudaMalloc(p, bytes)); return 0; }
int friqd_free(void *p) { CK(cudaFree(p)); return 0; }
int friqd_memset(void *p, int v, unsigned long bytes) { CK(cudaMemset(p, v, bytes)); return 0; }
int friqd_h2d(void *dst, const void *src, unsigned long bytes) { CK(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); return 0; }
int friqd_d2h(void *dst, const void *src, unsigned long bytes) { CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); return 0; }
int friqd_d2d_async(void *dst, const void *src, unsigned long bytes, void *stream)
{ CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream)); return 0; }
int friqd_h2d_2d(void *dst, unsigned long dpitch, const void *src, unsigned long spitch, unsigned long width, unsigned long height)
{ CK(cudaMemcpy2D(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice)); return 0; }
int friqd_d2h_2d(void *dst, unsigned long dpitch, const void *src, unsigned long spitch, unsigned long width, unsigned long height)
{ CK(cudaMemcpy2D(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost)); return 0; }
int friqd_h2d_async(void *dst, const void *src, unsigned long bytes, void *stream)
{ CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream)); return 0; }
int friqd_d2h_async(void *dst, const void *src, unsigned long bytes, void *stream)
{ CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream)); return 0; }
int friqd_device_count(int *count) { CK(cudaGetDeviceCount(count)); return 0; }
int friqd_broadcast_rb(friq_devstate st, const unsigned int *d_words, const double *d_q, int R, void *stream)
{
    k_broadcast_rb<<<grid_for(st.n), FRIQ_BLOCK, 0, (cudaStream_t)stream>>>(st, d_words, d_q, R);
    CK(cudaGetLastError());
    return 0;
}
int friqd_permute_in3(friq_devstate st, const double *s0, double *d0, const double *s1, double *d1, const double *s2, double *d2, void *stream)
{
    k_permute_in3<<<grid_for(st.n), FRIQ_BLOCK, 0, (cudaStream_t)stream>>>(st, s0, d0, s1, d1, s2, d2);
    CK(cudaGetLastError());
    return 0;
}
int friqd_new_episodes(friq_devstate st, void *stream)
{
    k_new_episodes<<<grid_for(st.n), FRIQ_BLOCK, 0, (cudaStream_t)stream>>>(st);
    CK(cudaGetLastError());
    return 0;
}
int friqd_result_stats(const double *d_ret, const int *d_rules, long n, double *d_out3, void *stream)
{
    k_result_stats<<<1, 1024, 0, (cudaStream_t)stream>>>(d_ret, d_rules, n, d_out3);
    CK(cudaGetLastError());
    return 0;
}
int friqd_set_s0(friq_devstate st, const double *d_s0, int S, void *stream)
{
    k_set_s0<<<grid_for(st.n), FRIQ_BLOCK, 0, (cudaStream_t)stream>>>(st, d_s0, S);
    CK(cudaGetLastError());
    return 0;
}
int friqd_retile(friq_devstate src, friq_devstate dst, int *d_work, int by_rules, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    const dim3 g = grid_for(src.n);
    if (by_rules) {
        CK(cudaMemsetAsync(d_work, 0, sizeof(int) * (src.cap + 1), s));
        k_retile_hist<<<g, FRIQ_BLOCK, 0, s>>>(src, d_work);
        k_retile_scan<<<1, 256, 0, s>>>(d_work, src.cap);
    }
    k_retile_assign<<<g, FRIQ_BLOCK, 0, s>>>(src, d_work, by_rules);
    k_retile_move<<<dim3((unsigned)((src.n + FRIQ_TILE - 1) / FRIQ_TILE)), dim3(FRIQ_TILE, FRIQ_MOVE_SLICES), 0, s>>>(src, dst, d_work);
    CK(cudaGetLastError());
    return 0;
}
int friqd_unpermute(friq_devstate st, const void *src, void *dst, int elem_bytes, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (elem_bytes == 8) k_unpermute<unsigned long long><<<grid_for(st.n), FRIQ_BLOCK, 0, s>>>(st, (const unsigned long long *)src, (unsigned long long *)dst);
    else k_unpermute<unsigned int><<<grid_for(st.n), FRIQ_BLOCK, 0, s>>>(st, (const unsigned int *)src, (unsigned int *)dst);
    CK(cudaGetLastError());
    return 0;
}
#ifdef FRIQ_TIMELINE
int friqd_timeline_fetch(unsigned long long *host, int max_entries, int reset)
{
    unsigned int n = 0; CK(cudaDeviceSynchronize());
    CK(cudaMemcpyFromSymbol(&n, friq_tl_n, sizeof n));
    if (n > FRIQ_TL_MAX) n = FRIQ_TL_MAX;
    if ((int)n > max_entries) n = max_entries;
    if (host && n) CK(cudaMemcpyFromSymbol(host, friq_tl_buf, (size_t)n * 48));
    if (reset) { const unsigned int z = 0; CK(cudaMemcpyToSymbol(friq_tl_n, &z, sizeof z)); }
    return (int)n;
}
#endif
int friqd_stream_create(void **s) { cudaStream_t t; CK(cudaStreamCreate(&t)); *s = (void *)t; return 0; }
int friqd_stream_destroy(void *s) { CK(cudaStreamDestroy((cudaStream_t)s)); return 0; }
int friqd_event_create(void **e) { cudaEvent_t t; CK(cudaEventCreateWithFlags(&t, cudaEventDisableTiming)); *e = (void *)t; return 0; }
int friqd_event_destroy(void *e) { CK(cudaEventDestroy((cudaEvent_t)e)); return 0; }
int friqd_event_record(void *e, void *stream) { CK(cudaEventRecord((cudaEvent_t)e, (cudaStream_t)stream)); return 0; }
int friqd_stream_wait_event(void *stream, void *e) { CK(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)e, 0)); return 0; }
int friqd_device_sync(void) { CK(cudaDeviceSynchronize()); return 0; }
int friqd_sync(void *stream) { CK(cudaStreamSynchronize((cudaStream_t)stream)); return 0; }
const char *friqd_error_string(int err)
{
    if (err == FRIQ_ERR_SHAPE) return "unsupported problem shape: per-call entry points need 1 <= S <= 7 state dimensions, p = S + 1 and A <= 32 actions; the fused kernels a built-in environment with its own number of state dimensions (any action set of <= 32 actions)";
    return cudaGetErrorString((cudaError_t)err);
}

int friqd_init_agents(const friq_devprob *dprob, friq_devstate st, unsigned long long seed0, long first_agent,
                      double alpha, double gamma, double epsilon, void *stream)
{
    k_init_agents<<<grid_for(st.n), FRIQ_BLOCK, 0, (cudaStream_t)stream>>>(dprob, st, seed0, first_agent, alpha, gamma, epsilon);
    CK(cudaGetLastError());
    return 0;
}

int friqd_run_steps(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, long n_steps, int mode, int variant, void *stream)
{   // evaluation policy (friq_batch_set_policy): 0 lock-step pass over all actions (+ hit sweep / compacted sums when few pairs lack
    // a hit); 1 exact hits first + compacted Shepard chains only; 2 the same in table mode (memoised weights, per-agent indexes)
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
    // table mode (memoised weights + per-agent indexes) only where the COOP kernels run and the indexes are sized for this table
    const bool coop = variant >= 1 && hp->hit_scan && hp->safe_range &&
                      ((hp->env == FRIQ_ENV_MOUNTAINCAR && FRIQ_SHAPE_MC(hp)) || (hp->env == FRIQ_ENV_ACROBOT && FRIQ_SHAPE_AC(hp)));
    const bool tab = coop && variant == 2 && hp->wtab && st.rbc && st.idx_cells == hp->wcells && st.cap <= 65535 &&                             /* the rule count rides in 16 bits of a round descriptor */
                     ((st.cap % FRIQ_CELL_VEC) == 0 || !((mode & FRIQ_M_KEEP_WEIGHTS) && st.wref));   /* kept weights (knowledge sharing) are stored in whole blocks of rules */
    if (tab) {
        CK(cudaMemsetAsync(st.cmap, 0, (size_t)((st.n + FRIQ_TILE - 1) / FRIQ_TILE) * st.idx_cells * FRIQ_TILE * sizeof(unsigned short), s));
        k_refresh_index<<<g, FRIQ_BLOCK, 0, s>>>(dprob, st);
    }
#define FRIQ_LAUNCH_STEP(S_, P_, A_, ACH_, ENV_, UNR_)                                                                   \
    do {                                                                                                              \
        if (hp->safe_range) k_run_steps<S_, P_, A_, ACH_, ENV_, UNR_, true, 0><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, n_steps, mode); \
        else k_run_steps<S_, P_, A_, ACH_, ENV_, UNR_, false, 0><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, n_steps, mode);     \
    } while (0)
    const dim3 gt((unsigned)((st.n + FRIQ_BLOCK_TAB - 1) / FRIQ_BLOCK_TAB));
    if (tab && hp->env == FRIQ_ENV_MOUNTAINCAR) k_run_steps<2, 3, 3, 3, FRIQ_ENV_MOUNTAINCAR, 1, true, 2><<<gt, FRIQ_BLOCK_TAB, 0, s>>>(dprob, st, n_steps, mode);
    else if (tab && hp->env == FRIQ_ENV_ACROBOT) k_run_steps<4, 5, 3, 3, FRIQ_ENV_ACROBOT, 1, true, 2><<<gt, FRIQ_BLOCK_TAB, 0, s>>>(dprob, st, n_steps, mode);
    else if (coop && hp->env == FRIQ_ENV_MOUNTAINCAR)
        k_run_steps<2, 3, 3, 3, FRIQ_ENV_MOUNTAINCAR, 1, true, 1><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, n_steps, mode);
    else if (coop && hp->env == FRIQ_ENV_ACROBOT)
        k_run_steps<4, 5, 3, 3, FRIQ_ENV_ACROBOT, 1, true, 1><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, n_steps, mode);
    else if (hp->env == FRIQ_ENV_MOUNTAINCAR && FRIQ_SHAPE_MC(hp)) FRIQ_LAUNCH_STEP(2, 3, 3, 3, FRIQ_ENV_MOUNTAINCAR, FRIQ_UNR_MC);
    else if (hp->env == FRIQ_ENV_ACROBOT && FRIQ_SHAPE_AC(hp)) FRIQ_LAUNCH_STEP(4, 5, 3, 3, FRIQ_ENV_ACROBOT, 2);
    else if (hp->env == FRIQ_ENV_CARTPOLE && FRIQ_SHAPE_CP(hp)) FRIQ_LAUNCH_STEP(4, 5, 21, 3, FRIQ_ENV_CARTPOLE, 1);
    // any other action set / grid on a built-in environment: the plain fused loop
    else if (hp->env == FRIQ_ENV_MOUNTAINCAR && hp->S == 2 && hp->p == 3 && hp->A >= 1 && hp->A <= FRIQ_MAXA)
        k_run_steps_anyA<2, 3, FRIQ_ENV_MOUNTAINCAR><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, n_steps, mode);
    else if (hp->env == FRIQ_ENV_ACROBOT && hp->S == 4 && hp->p == 5 && hp->A >= 1 && hp->A <= FRIQ_MAXA)
        k_run_steps_anyA<4, 5, FRIQ_ENV_ACROBOT><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, n_steps, mode);
    else if (hp->env == FRIQ_ENV_CARTPOLE && hp->S == 4 && hp->p == 5 && hp->A >= 1 && hp->A <= FRIQ_MAXA)
        k_run_steps_anyA<4, 5, FRIQ_ENV_CARTPOLE><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, n_steps, mode);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}

int friqd_run_shared(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, const double *d_srb, int Rs, long n_steps, int mode, void *stream)
{
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
#define FRIQ_LAUNCH_SHARED(S_, P_, A_, ENV_, UNR_)                                                                     \
    do {                                                                                                              \
        if (hp->safe_range) k_infer_shared<S_, P_, A_, ENV_, UNR_, true><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_srb, Rs, n_steps, mode); \
        else k_infer_shared<S_, P_, A_, ENV_, UNR_, false><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_srb, Rs, n_steps, mode);               \
    } while (0)
    if (hp->env == FRIQ_ENV_MOUNTAINCAR && FRIQ_SHAPE_MC(hp)) FRIQ_LAUNCH_SHARED(2, 3, 3, FRIQ_ENV_MOUNTAINCAR, 3);
    else if (hp->env == FRIQ_ENV_ACROBOT && FRIQ_SHAPE_AC(hp)) FRIQ_LAUNCH_SHARED(4, 5, 3, FRIQ_ENV_ACROBOT, 2);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}
// memoised conclusions of a shared rule base (k_shared_memo_build / k_shared_start / k_infer_memo)
int friqd_shared_memo_build(const friq_devprob *hp, const friq_devprob *dprob, const double *d_srb, int Rs, double *d_memo, int cells, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    const long warps = (long)cells * hp->A;
    const unsigned g = (unsigned)((warps * 32 + 127) / 128);
    if (FRIQ_SHAPE_MC(hp)) k_shared_memo_build<2, 3><<<g, 128, 0, s>>>(dprob, d_srb, Rs, d_memo, cells, hp->A);
    else if (FRIQ_SHAPE_AC(hp)) k_shared_memo_build<4, 5><<<g, 128, 0, s>>>(dprob, d_srb, Rs, d_memo, cells, hp->A);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}
// scratch of the de-duplication, one allocation: table u64[T] | slot_of_agent i32[n_pad] | rank_of_slot i32[T] | slot_of_rank i32[T] | count i32[4] | uconc f64[n_pad / DIV * A + A]
static inline unsigned friq_start_tsize(long n) { unsigned t = 1024; while ((long)t < 2 * n) t <<= 1; return t; }
unsigned long friqd_shared_start_scratch_bytes(long n_pad, int A)
{
    const unsigned long T = friq_start_tsize(n_pad);
    return 8ul * T + 4ul * n_pad + 4ul * T + 4ul * T + 16ul + 8ul * ((unsigned long)(n_pad / FRIQ_START_DEDUPE_DIV + 1) * A + A);
}
int friqd_shared_start(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, const double *d_srb, int Rs, double *d_start, void *d_scratch, void *stream)
{
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
    const int A = hp->A;
    int *count = nullptr;
    if (d_scratch && (FRIQ_SHAPE_MC(hp) || FRIQ_SHAPE_AC(hp)) && hp->U <= 65535) {
        const unsigned T = friq_start_tsize(st.n_pad);
        unsigned long long *table = (unsigned long long *)d_scratch;
        int *slot_of_agent = (int *)(table + T), *rank_of_slot = slot_of_agent + st.n_pad, *slot_of_rank = rank_of_slot + T;
        count = slot_of_rank + T;
        double *uconc = (double *)(count + 4);
        CK(cudaMemsetAsync(table, 0xff, 8ul * T, s));
        CK(cudaMemsetAsync(count, 0, 16, s));
        k_start_keys<<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, table, T - 1, slot_of_agent);
        k_start_compact<<<(T + 255) / 256, 256, 0, s>>>(table, T, rank_of_slot, slot_of_rank, count);
        const long maxw = (long)(st.n / FRIQ_START_DEDUPE_DIV + 1) * A;
        const unsigned gu = (unsigned)((maxw * 32 + 127) / 128);
        if (FRIQ_SHAPE_MC(hp)) k_start_unique<2, 3><<<gu, 128, 0, s>>>(dprob, d_srb, Rs, table, slot_of_rank, count, st.n, uconc, A);
        else k_start_unique<4, 5><<<gu, 128, 0, s>>>(dprob, d_srb, Rs, table, slot_of_rank, count, st.n, uconc, A);
        k_start_scatter<<<g, FRIQ_BLOCK, 0, s>>>(st, slot_of_agent, rank_of_slot, count, uconc, d_start, A);
    }
#define FRIQ_LAUNCH_START(S_, P_, A_, UNR_)                                                                            \
    do {                                                                                                              \
        if (hp->safe_range) k_shared_start<S_, P_, A_, UNR_, true><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_srb, Rs, d_start, count); \
        else k_shared_start<S_, P_, A_, UNR_, false><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_srb, Rs, d_start, count);               \
    } while (0)
    if (FRIQ_SHAPE_MC(hp)) FRIQ_LAUNCH_START(2, 3, 3, 3);
    else if (FRIQ_SHAPE_AC(hp)) FRIQ_LAUNCH_START(4, 5, 3, 2);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}
int friqd_run_shared_memo(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, const double *d_memo, const double *d_start,
                          int Rs, long n_steps, int mode, void *stream)
{
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
    if (hp->env == FRIQ_ENV_MOUNTAINCAR && FRIQ_SHAPE_MC(hp)) k_infer_memo<2, 3, FRIQ_ENV_MOUNTAINCAR><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_memo, d_start, Rs, n_steps, mode);
    else if (hp->env == FRIQ_ENV_ACROBOT && FRIQ_SHAPE_AC(hp)) k_infer_memo<4, 3, FRIQ_ENV_ACROBOT><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_memo, d_start, Rs, n_steps, mode);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}

int friqd_vag_concl(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, const double *d_ant, double *d_conc, int *d_hit,
                    double *d_weights, void *stream)
{
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
    if (hp->S == 2 && hp->p == 3) k_vag_concl<2, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_conc, d_hit, d_weights);
    else if (hp->S == 4 && hp->p == 5) k_vag_concl<4, 5><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_conc, d_hit, d_weights);
    else if (hp->S == 1 && hp->p == 2) k_vag_concl<1, 2><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_conc, d_hit, d_weights);
    else if (hp->S == 3 && hp->p == 4) k_vag_concl<3, 4><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_conc, d_hit, d_weights);
    else if (hp->S == 5 && hp->p == 6) k_vag_concl<5, 6><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_conc, d_hit, d_weights);
    else if (hp->S == 6 && hp->p == 7) k_vag_concl<6, 7><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_conc, d_hit, d_weights);
    else if (hp->S == 7 && hp->p == 8) k_vag_concl<7, 8><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_conc, d_hit, d_weights);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}

int friqd_best_action(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, const double *d_states, double *d_actconc,
                      int *d_best, void *stream)
{
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
    if (FRIQ_SHAPE_MC(hp)) k_best_action<2, 3, 3, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_states, d_actconc, d_best);
    else if (FRIQ_SHAPE_AC(hp)) k_best_action<4, 5, 3, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_states, d_actconc, d_best);
    else if (FRIQ_SHAPE_CP(hp)) k_best_action<4, 5, 21, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_states, d_actconc, d_best);
    else if (hp->A < 1 || hp->A > FRIQ_MAXA) return FRIQ_ERR_SHAPE;
    else if (hp->S == 1 && hp->p == 2) k_best_action_anyA<1, 2><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_states, d_actconc, d_best);
    else if (hp->S == 2 && hp->p == 3) k_best_action_anyA<2, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_states, d_actconc, d_best);
    else if (hp->S == 3 && hp->p == 4) k_best_action_anyA<3, 4><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_states, d_actconc, d_best);
    else if (hp->S == 5 && hp->p == 6) k_best_action_anyA<5, 6><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_states, d_actconc, d_best);
    else if (hp->S == 6 && hp->p == 7) k_best_action_anyA<6, 7><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_states, d_actconc, d_best);
    else if (hp->S == 7 && hp->p == 8) k_best_action_anyA<7, 8><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_states, d_actconc, d_best);
    else if (hp->S == 4 && hp->p == 5) k_best_action_anyA<4, 5><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_states, d_actconc, d_best);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}

int friqd_update_sarsa(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, const double *d_qant, const double *d_reward,
                       const double *d_curqant, void *stream)
{
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
    if (hp->S == 2 && hp->p == 3) k_update_sarsa<2, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_qant, d_reward, d_curqant);
    else if (hp->S == 4 && hp->p == 5) k_update_sarsa<4, 5><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_qant, d_reward, d_curqant);
    else if (hp->S == 1 && hp->p == 2) k_update_sarsa<1, 2><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_qant, d_reward, d_curqant);
    else if (hp->S == 3 && hp->p == 4) k_update_sarsa<3, 4><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_qant, d_reward, d_curqant);
    else if (hp->S == 5 && hp->p == 6) k_update_sarsa<5, 6><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_qant, d_reward, d_curqant);
    else if (hp->S == 6 && hp->p == 7) k_update_sarsa<6, 7><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_qant, d_reward, d_curqant);
    else if (hp->S == 7 && hp->p == 8) k_update_sarsa<7, 8><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_qant, d_reward, d_curqant);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}

int friqd_classic_step(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, const double *d_in, double *d_out, void *stream)
{
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
    if (FRIQ_SHAPE_MC(hp)) k_classic_step<2, 3, 3, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_in, d_out);
    else if (FRIQ_SHAPE_AC(hp)) k_classic_step<4, 5, 3, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_in, d_out);
    else if (FRIQ_SHAPE_CP(hp)) k_classic_step<4, 5, 21, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_in, d_out);
    else if (hp->A < 1 || hp->A > FRIQ_MAXA) return FRIQ_ERR_SHAPE;
    else if (hp->S == 1 && hp->p == 2) k_classic_step_anyA<1, 2><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_in, d_out);
    else if (hp->S == 2 && hp->p == 3) k_classic_step_anyA<2, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_in, d_out);
    else if (hp->S == 3 && hp->p == 4) k_classic_step_anyA<3, 4><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_in, d_out);
    else if (hp->S == 5 && hp->p == 6) k_classic_step_anyA<5, 6><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_in, d_out);
    else if (hp->S == 6 && hp->p == 7) k_classic_step_anyA<6, 7><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_in, d_out);
    else if (hp->S == 7 && hp->p == 8) k_classic_step_anyA<7, 8><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_in, d_out);
    else if (hp->S == 4 && hp->p == 5) k_classic_step_anyA<4, 5><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_in, d_out);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}

int friqd_rule_distance(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, const double *d_ant, double *d_dists, int *d_hit, void *stream)
{
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
    if (hp->S == 2) k_rule_distance<2><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_dists, d_hit);
    else if (hp->S == 4) k_rule_distance<4><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_dists, d_hit);
    else if (hp->S == 1) k_rule_distance<1><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_dists, d_hit);
    else if (hp->S == 3) k_rule_distance<3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_dists, d_hit);
    else if (hp->S == 5) k_rule_distance<5><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_dists, d_hit);
    else if (hp->S == 6) k_rule_distance<6><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_dists, d_hit);
    else if (hp->S == 7) k_rule_distance<7><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, d_ant, d_dists, d_hit);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}

int friqd_bestact_dists(const friq_devprob *hp, friq_devstate st, const double *d_dists, double *d_conc, void *stream)
{
    k_bestact_dists<<<grid_for(st.n), FRIQ_BLOCK, 0, (cudaStream_t)stream>>>(st, d_dists, d_conc, hp->p);
    CK(cudaGetLastError());
    return 0;
}

unsigned long friqd_rbc_tile_slots(int cap) { return (unsigned long)(friq_rbc_slots(cap) * FRIQ_TILE); }
int friqd_build_wtab(const friq_devprob *hp, const friq_devprob *dprob, double *d_wtab)
{
    const long n = (long)hp->wcells * hp->wcells;
    const unsigned g = (unsigned)((n + 255) / 256);
    if (hp->K == 3 && hp->p == 3) k_build_wtab<3, 3><<<g, 256>>>(dprob, d_wtab, hp->wcells, hp->wstride);
    else if (hp->K == 5 && hp->p == 5) k_build_wtab<5, 5><<<g, 256>>>(dprob, d_wtab, hp->wcells, hp->wstride);
    else return (int)cudaErrorInvalidValue;
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return 0;
}
int friqd_merge_rb(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, long receiver, const unsigned int *d_words,
                   const double *d_q, int Rs, void *stream)
{
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
    if (!st.wref) return (int)cudaErrorInvalidValue;
    if (hp->S == 2 && hp->p == 3) k_merge_list<2, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, receiver, d_words, d_q, Rs);
    else if (hp->S == 4 && hp->p == 5) k_merge_list<4, 5><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, receiver, d_words, d_q, Rs);
    else if (hp->S == 1 && hp->p == 2) k_merge_list<1, 2><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, receiver, d_words, d_q, Rs);
    else if (hp->S == 3 && hp->p == 4) k_merge_list<3, 4><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, receiver, d_words, d_q, Rs);
    else if (hp->S == 5 && hp->p == 6) k_merge_list<5, 6><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, receiver, d_words, d_q, Rs);
    else if (hp->S == 6 && hp->p == 7) k_merge_list<6, 7><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, receiver, d_words, d_q, Rs);
    else if (hp->S == 7 && hp->p == 8) k_merge_list<7, 8><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, receiver, d_words, d_q, Rs);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}
int friqd_team_merge(const friq_devprob *hp, const friq_devprob *dprob, friq_devstate st, int team, int phase, void *stream)
{
    const dim3 g = grid_for(st.n); cudaStream_t s = (cudaStream_t)stream;
    if (!st.wref || team < 2) return (int)cudaErrorInvalidValue;
    if (hp->S == 2 && hp->p == 3) k_team_merge<2, 3><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, team, phase);
    else if (hp->S == 4 && hp->p == 5) k_team_merge<4, 5><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, team, phase);
    else if (hp->S == 1 && hp->p == 2) k_team_merge<1, 2><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, team, phase);
    else if (hp->S == 3 && hp->p == 4) k_team_merge<3, 4><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, team, phase);
    else if (hp->S == 5 && hp->p == 6) k_team_merge<5, 6><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, team, phase);
    else if (hp->S == 6 && hp->p == 7) k_team_merge<6, 7><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, team, phase);
    else if (hp->S == 7 && hp->p == 8) k_team_merge<7, 8><<<g, FRIQ_BLOCK, 0, s>>>(dprob, st, team, phase);
    else return FRIQ_ERR_SHAPE;
    CK(cudaGetLastError());
    return 0;
}
int friqd_team_ctl(friq_devstate st, int team, int op, int max_episodes, int *d_count, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (op == 1) { CK(cudaMemsetAsync(d_count, 0, sizeof(int), s)); }
    k_team_ctl<<<grid_for(st.n), FRIQ_BLOCK, 0, s>>>(st, team, op, max_episodes, d_count);
    CK(cudaGetLastError());
    return 0;
}
int friqd_trig(const double *d_x, double *d_sin, double *d_cos, long n, void *stream)
{
    k_trig<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_sin, d_cos, n);
    CK(cudaGetLastError());
    return 0;
}
int friqd_pow_x87(const double *d_x, double *d_out, long n, int p, int force_exact, void *stream)
{
    k_pow<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_out, n, p, force_exact);
    CK(cudaGetLastError());
    return 0;
}

// measured FP64 DADD/DMUL issue rate of this GPU, in G lane-instructions per second
int friqd_fp64_peak(double *out_ginstr_per_s, int iters)
{
    double *d; CK(cudaMalloc(&d, 8));
    int dev, sms; CK(cudaGetDevice(&dev)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256;
    k_fp64_peak<<<blocks, threads>>>(d, iters / 8);
    CK(cudaEventRecord(e0));
    k_fp64_peak<<<blocks, threads>>>(d, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    *out_ginstr_per_s = (double)blocks * threads * iters * 16.0 / (ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    return 0;
}

}  // extern "C"
