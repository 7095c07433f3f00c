/* csrc/friq_devprob.h -- layouts shared by the ANSI-C host layer (friq_host.c) and
 * the CUDA layer (friq_engine.cu).  Plain C, no CUDA types.
 *
 * friq_devprob is the read-only "problem block": everything the fused step kernel
 * needs about the FRIQ problem (what struct FIVERB + struct frirl_desc hold on the
 * host in the reference: FIVE.h:24-76, frirl_types.h:62-169), reduced to what the
 * path touches.  It is ~9.7 KB and is staged into shared memory with one TMA bulk
 * copy (cp.async.bulk + mbarrier) per CTA.
 *
 * Rule storage.  A rule's antecedent in dimension k is stored as a 4- to 6-bit index
 * into a per-dimension *value dictionary* dict_val[k][] (the grid values of
 * frirl_init.c:55-102 first - every rule the learner inserts is grid-snapped,
 * frirl_update_sarsa.c:367-375 - then any extra value met in a loaded rule base).
 * dict_ve[k][j] = VE_k[nearest_index(dict_val[k][j])] is exactly what
 * five_add_rule.c:71-84 caches in rseqant_veval.  One u32 per rule (FRIQ_BITS_OF(K)
 * bits per index); the consequent is one f64.
 */
#ifndef FRIQ_DEVPROB_H
#define FRIQ_DEVPROB_H
#include <stddef.h>

#define FRIQ_MAXK 8          /* FIVE_MAX_NUM_OF_UNIVERSES, FIVE.h:19 */
#define FRIQ_MAXDICT 64          /* room per dimension in the dictionaries (the most a 6-bit index addresses) */
/* K antecedent indexes share one 32-bit rule word: 6 bits each up to K = 5 (64 values per dimension), 5 bits for K = 6 (32),
 * 4 bits for K = 7 and 8 = FIVE_MAX_NUM_OF_UNIVERSES (16) */
#define FRIQ_BITS_OF(K) ((K) <= 5 ? 6 : (K) == 6 ? 5 : 4)
#define FRIQ_DICT_MAX(K) (1 << FRIQ_BITS_OF(K))
#define FRIQ_FIELD(w, k, K) (((w) >> (FRIQ_BITS_OF(K) * (k))) & ((1u << FRIQ_BITS_OF(K)) - 1u))
#define FRIQ_MAXA 32
#define FRIQ_TILE 32         /* agents per rule-base tile (= one warp) */

#define FRIQ_ENV_NONE (-1)
#ifndef FRIQ_ENV_MOUNTAINCAR            /* same values as include/friq_b200.h */
#define FRIQ_ENV_MOUNTAINCAR 0
#define FRIQ_ENV_ACROBOT 1
#define FRIQ_ENV_CARTPOLE 2
#endif

/* per-agent flag bits */
enum { FRIQ_F_NEED_RESET = 1, FRIQ_F_INSERTED = 2, FRIQ_F_CONVERGED = 4, FRIQ_F_OVERFLOW = 8,
       FRIQ_F_EPENDED = 16,     /* frirl_desc.epended: the "same as the previous episode" test held (frirl_sequential_run.c:90) */
       FRIQ_F_TEAM_DONE = 32 }; /* knowledge sharing: this agent's team has finished (its leader found the rule base complete) */
/* run mode bits */
enum { FRIQ_M_LEARN = 1, FRIQ_M_TRACK = 2, FRIQ_M_STOP_ON_CONVERGE = 4, FRIQ_M_STOP_AT_EPISODE_END = 8, FRIQ_M_RETILE = 16,
       FRIQ_M_KEEP_WEIGHTS = 32 };   /* maintain FIVERB.weights (friq_devstate.wref) exactly as the reference does: merge_rb reads it stale */

typedef struct friq_devprob {
    int S, K, A, U, p, env, skip_rules, no_random;
    int max_steps, max_episodes, qp_reuse, ncorner;
    int safe_range, hit_scan, pad2, pad3;     /* safe_range: all squared distances in {0} U [2^-200, 2^200] (host-verified);
                                               * hit_scan: distance 0 <=> equal dictionary indices (VE values of distinct dictionary
                                               * entries differ in every dimension), so exact hits can be found by comparing packed words */
    int grid_len[FRIQ_MAXK];
    int dict_len[FRIQ_MAXK];
    int act_dict[FRIQ_MAXA];                  /* dictionary index (dimension S) of action a */
    int act_hit[FRIQ_MAXA];                   /* dictionary index whose VE value equals act_ve[a] (what a rule at distance 0 holds), -1 none */
    int def_hit[FRIQ_MAXK];                   /* the same for the episode start state: index whose VE value equals ve_def[k], -1 none */
    double qdiff_pos, qdiff_neg, wsig, qdiff_final_tol, reward_good_above, pad0;
    double u0[FRIQ_MAXK], udiv[FRIQ_MAXK];
    double values_def[FRIQ_MAXK], ve_def[FRIQ_MAXK];
    double q_abs0[FRIQ_MAXK], q_div[FRIQ_MAXK];   /* quantiser of mountaincar.c:98-125 / acrobot.c:165-192 */
    double act_val[FRIQ_MAXA];
    double act_ve[FRIQ_MAXA];                 /* possible_actions->vevalues (frirl_init.c:156-158, len = U-1) */
    double act_ve_eval[FRIQ_MAXA];            /* VE_act[nearest_index(act_val)] as five_rule_distance sees it */
    double dict_val[FRIQ_MAXK][FRIQ_MAXDICT];
    double dict_ve[FRIQ_MAXK][FRIQ_MAXDICT];
    const double *u, *ve;                     /* device pointers, [K][U] row-major */
    const double *wtab;                       /* device pointer or NULL: memoised Shepard weights 1/d^p between any two dictionary cells,
                                               * [wcells][wstride], cell = sum_k index_k * wlin[k] (built by k_build_wtab with the very device
                                               * functions the kernels use, so a looked-up weight IS the computed one) */
    int wcells, wlin[FRIQ_MAXK], wstride;     /* wstride = wcells rounded up to an even count (rows 16-byte aligned) */
    double q_rdiv[FRIQ_MAXK];                 /* 1 / q_div: the quantiser's quick path (exact result, division only when a rounding boundary is near) */
    /* ---- everything above is staged to shared memory; the tail stays in global ---- */
    unsigned int corner[256];                 /* packed 2^K corner rules, frirl_init_rb.c:99-126 */
} friq_devprob;

/* bytes staged to shared memory (multiple of 16) */
#define FRIQ_DEVPROB_SMEM_BYTES ((unsigned)((offsetof(friq_devprob, corner) + 15ul) & ~15ul))

/* per-agent state, structure of arrays in device memory; n = agents padded to a tile */
typedef struct friq_devstate {
    long n, n_pad;
    int cap;
    unsigned int *rbw;          /* [tiles][cap][32] packed antecedent indices */
    double *rbq;                /* [tiles][cap][32] consequents */
    double *sw;                 /* [tiles][cap][32] scratch: Shepard weights wi of the last Q(s,a) */
    double *prevq;              /* [tiles][cap][32] consequents at episode start (convergence test) */
    double *s;                  /* [FRIQ_MAXK][n_pad] raw state */
    double *qs;                 /* [FRIQ_MAXK][n_pad] quantised antecedents of the previous step (q_ant) */
    double *vo;                 /* [FRIQ_MAXK][n_pad] their VE values */
    int *aidx, *flags, *R, *step_in_ep, *last_steps, *last_rules, *prev_steps, *prev_numru;
    unsigned int *ep_num;
    double *ep_return, *last_return, *prev_reward;
    unsigned long long *total_steps, *sum_rules, *seed;
    double *alpha, *gamma, *epsilon;
    int *hitq;                  /* [n_pad] exact-hit rule of Q(s,a) for the pending update, -1 none, -2 unknown */
    double *s0;                 /* [FRIQ_MAXK][n_pad] optional per-agent episode start state (NULL: values_def) */
    int *agent_of_slot;         /* [n_pad] which agent lives in slot p (identity unless re-tiled by rule count) */
    /* derived per-agent indexes of the fused kernel's table mode (friq_devprob.wtab); rebuilt from rbw before every launch
     * (k_refresh_index) and kept current by the kernel when it appends a rule; NULL when the problem has no table */
    double *wref;               /* [tiles][cap][32] or NULL: FIVERB.weights (FIVE.h:24-76), the NORMALISED weights of the last
                                 * FIVE_vag_concl_weight call without an exact hit (update_rules, frirl_update_sarsa.c:39-40; merge_rb,
                                 * frirl_agent.c:75-76) - persistent, only kept under FRIQ_M_KEEP_WEIGHTS (knowledge sharing) */
    unsigned short *rbc;        /* [tiles][cap][32] dictionary cell of every rule (column of wtab) */
    unsigned short *cmap;       /* [tiles][idx_cells][32] 1 + index of the first rule in that cell, 0 = none: exact hits in O(1) */
    int idx_cells;              /* cells per agent in cmap (= friq_devprob.wcells when the arrays were sized) */
    unsigned long long *ctr;    /* [FRIQ_NCTR] executed-work counters of the fused kernels, summed over all agents and launches (one atomic per warp
                                 * and counter at kernel exit): what bench.py reports as the EXECUTED work beside the nominal roofline */
} friq_devstate;
enum { FRIQ_CTR_PAIRS = 0,      /* (agent, channel) pairs without an exact-hit rule: conclusions that needed their Shepard sums */
       FRIQ_CTR_TAB_RULES = 1,  /* rules summed with a looked-up weight (table mode: 1 DMUL + 2 DADD each) */
       FRIQ_CTR_CHAIN_RULES = 2,/* rules summed with a computed weight (sqrt -> x87 power -> reciprocal chain) */
       FRIQ_CTR_HIT_PAIRS = 3,  /* (agent, channel) pairs answered by an exact-hit rule */
       FRIQ_CTR_APPENDED = 4,   /* rules appended */
       FRIQ_CTR_WUPD_RULES = 5, /* rules scanned by weighted updates */
       FRIQ_CTR_STEPS = 6,      /* env steps executed */
       FRIQ_NCTR = 8 };

#ifdef __cplusplus
extern "C" {
#endif
/* CUDA layer entry points (friq_engine.cu); return 0 or a cudaError_t value */
int friqd_set_device(int device);
int friqd_malloc(void **p, unsigned long bytes);
int friqd_free(void *p);
int friqd_memset(void *p, int v, unsigned long bytes);
int friqd_h2d(void *dst, const void *src, unsigned long bytes);
int friqd_d2h(void *dst, const void *src, unsigned long bytes);
int friqd_d2d_async(void *dst, const void *src, unsigned long bytes, void *stream);
int friqd_h2d_2d(void *dst, unsigned long dpitch, const void *src, unsigned long spitch, unsigned long width, unsigned long height);
int friqd_d2h_2d(void *dst, unsigned long dpitch, const void *src, unsigned long spitch, unsigned long width, unsigned long height);
int friqd_h2d_async(void *dst, const void *src, unsigned long bytes, void *stream);
int friqd_d2h_async(void *dst, const void *src, unsigned long bytes, void *stream);
int friqd_device_count(int *count);
int friqd_broadcast_rb(friq_devstate st, const unsigned int *d_words, const double *d_q, int R, void *stream);
int friqd_permute_in3(friq_devstate st, const double *s0, double *d0, const double *s1, double *d1, const double *s2, double *d2, void *stream);
int friqd_new_episodes(friq_devstate st, void *stream);
int friqd_result_stats(const double *d_ret, const int *d_rules, long n, double *d_out3, void *stream);
int friqd_set_s0(friq_devstate st, const double *d_s0, int S, void *stream);
int friqd_retile(friq_devstate src, friq_devstate dst, int *d_work, int by_rules, void *stream);
int friqd_unpermute(friq_devstate st, const void *src, void *dst, int elem_bytes, void *stream);
int friqd_stream_create(void **s);
int friqd_stream_destroy(void *s);
int friqd_event_create(void **e);
int friqd_event_destroy(void *e);
int friqd_event_record(void *e, void *stream);
int friqd_stream_wait_event(void *stream, void *e);
int friqd_device_sync(void);
int friqd_sync(void *stream);
const char *friqd_error_string(int err);
int friqd_init_agents(const friq_devprob *dprob, friq_devstate st, unsigned long long seed0, long first_agent,
                      double alpha, double gamma, double epsilon, void *stream);
int friqd_run_steps(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, long n_steps, int mode, int variant, void *stream);
int friqd_run_shared(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, const double *d_srb, int Rs, long n_steps, int mode, void *stream);
int friqd_shared_memo_build(const friq_devprob *hprob, const friq_devprob *dprob, const double *d_srb, int Rs, double *d_memo, int cells, void *stream);
unsigned long friqd_shared_start_scratch_bytes(long n_pad, int A);
int friqd_shared_start(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, const double *d_srb, int Rs, double *d_start, void *d_scratch, void *stream);
int friqd_run_shared_memo(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, const double *d_memo, const double *d_start,
                          int Rs, long n_steps, int mode, void *stream);
int friqd_vag_concl(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, const double *d_ant, double *d_conc, int *d_hit,
                    double *d_weights, void *stream);
int friqd_best_action(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, const double *d_states, double *d_actconc,
                      int *d_best, void *stream);
int friqd_update_sarsa(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, const double *d_qant, const double *d_reward,
                       const double *d_curqant, void *stream);
int friqd_classic_step(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, const double *d_in, double *d_out, void *stream);
int friqd_rule_distance(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, const double *d_ant, double *d_dists, int *d_hit, void *stream);
int friqd_bestact_dists(const friq_devprob *hprob, friq_devstate st, const double *d_dists, double *d_conc, void *stream);
unsigned long friqd_rbc_tile_slots(int cap);      /* u16 entries of friq_devstate.rbc per tile of 32 agents */
int friqd_build_wtab(const friq_devprob *hprob, const friq_devprob *dprob, double *d_wtab);
int friqd_merge_rb(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, long receiver, const unsigned int *d_words,
                   const double *d_q, int Rs, void *stream);
int friqd_team_merge(const friq_devprob *hprob, const friq_devprob *dprob, friq_devstate st, int team, int phase, void *stream);
int friqd_team_ctl(friq_devstate st, int team, int op, int max_episodes, int *d_count, void *stream);
int friqd_trig(const double *d_x, double *d_sin, double *d_cos, long n, void *stream);
int friqd_pow_x87(const double *d_x, double *d_out, long n, int p, int force_exact, void *stream);
int friqd_fp64_peak(double *out_ginstr_per_s, int iters);
#ifdef __cplusplus
}
#endif
#endif
