/* csrc/friq_host.c -- ANSI-C host layer of the B200-native FRIQ-learning engine.
 *
 * Implements include/friq_b200.h on top of the thin CUDA C-ABI (friqd_*, friq_engine.cu):
 *   - init-time tables, bit-identical to the reference's (plain sequential FP64; compile with
 *     -ffp-contract=off): universes, scaling functions, vague environments, corner rule base,
 *     action VE values (frirl_init.c, frirl_init_ve.c, frirl_init_rb.c, FIVEGScFunc.c,
 *     FIVEGVagEnv.c, FIVEInit.c, frirl_app_helpers.c:32-44);
 *   - the per-dimension value dictionaries the device rule bases index into;
 *   - batch bookkeeping, host <-> device marshalling, rule-base file formats.
 * No FIVE evaluation or SARSA arithmetic happens here: that is device-only.
 */
#include "../../include/friq_b200.h"
#include "friq_devprob.h"

#include <fcntl.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define FRIQ_MAXGROUPS 8
static __thread char g_err[512] = "";
const char *friq_last_error(void) { return g_err; }
void friq_set_error_(int code, const char *fmt, ...)      /* for the library's other translation units (friq_multi.c) */
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); (void)code;
}
#define FAIL(code, ...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return (code); } while (0)
#define DEV(call) do { int e_ = (call); if (e_) FAIL(e_, "%s failed: %s", #call, friqd_error_string(e_)); } while (0)

struct friq_problem {
    friq_devprob hp;          /* host copy (hp.u / hp.ve hold DEVICE pointers once uploaded) */
    friq_devprob *dp;         /* device copy */
    double *u, *ve;           /* host tables [K][U] */
    double *d_u, *d_ve;
    double *d_wtab;           /* memoised Shepard weights between dictionary cells (friq_devprob.wtab), or NULL */
    int device;               /* device the tables live on, -1 = not uploaded yet */
    double alpha, gamma, epsilon;
};

struct friq_batch {
    friq_problem *P;
    friq_devstate st;         /* current placement */
    friq_devstate st2;        /* alternate buffers used by re-tiling (allocated on first use) */
    int have_st2, permuted;
    long retile_interval;     /* env steps between re-tilings inside one friq_batch_run (FRIQ_RUN_RETILE) */
    double *d_srb; int shared_R;   /* shared read-only rule base: rows of (K VE values, consequent) */
    double *d_memo; int memo_cells;   /* its memoised conclusions per (observation grid cell, action), or NULL */
    void *d_start_scratch;    /* de-duplication of the start states (k_start_keys ...) */
    double *d_start; int start_dirty; /* its conclusions at every agent's episode start state [A][n_pad]; dirty = recompute before the next run */
    int groups;               /* independent agent groups pipelined on their own streams (hides launch tails) */
    void *gstream[FRIQ_MAXGROUPS], *gdone[FRIQ_MAXGROUPS], *gstart;
    int variant;              /* evaluation policy of the fused kernel (friq_batch_set_policy) */
    int *d_work;              /* re-tiling scratch */
    void *d_tmp8, *d_tmp4a, *d_tmp4b;   /* un-permuted result staging */
    double *d_hyp;            /* agent-ordered staging of host hyper-parameter arrays (friq_batch_run_host on a re-tiled batch) */
    int device;
    uint64_t seed0; long first_agent;
    double *d_in0, *d_in1, *d_in2, *d_out0; int *d_iout; size_t cap_in, cap_out, cap_w; double *d_w;
};

/* ------------------------------------------------------------------ built-in problem blocks */
static const double MC_X[10] = {-1.5, -1.295, -1.09, -0.885, -0.68, -0.475, -0.27, -0.065, 0.14, 0.345};
static const double MC_V[6] = {-0.07, -0.042, -0.014, 0.014, 0.042, 0.07};
static const double AC_T[5] = {-1.570796326794897, -0.785398163397448, 0, 0.785398163397448, 1.570796326794897};
static const double CP_T[9] = {-0.2094, -0.1571, -0.1047, -0.0524, 0.0, 0.0524, 0.1047, 0.1571, 0.2094};
static const double CP_A[21] = {-1.0, -0.9, -0.8, -0.7, -0.6, -0.5, -0.3999999999999999, -0.29999999999999992,
                                -0.19999999999999995, -0.09999999999999998, 0.0, +0.09999999999999998, +0.19999999999999995,
                                +0.29999999999999992, +0.3999999999999999, +0.5, +0.6, +0.7, +0.8, +0.9, +1.0};
static const friq_dim_desc MC_DIMS[2] = {   /* mountaincar.c:244-266 */
    {10, MC_X, 0.205, 1.0, -0.5, 41, NULL, 0.1}, {6, MC_V, 0.028, 1.0, 0.0, 41, NULL, 0.005}};
static const friq_dim_desc AC_DIMS[4] = {   /* acrobot.c:261-310 */
    {5, AC_T, 0.785398163397448, 1.0, 0.0, 41, NULL, 0.1}, {5, AC_T, 0.785398163397448, 1.0, 0.0, 41, NULL, 0.1},
    {3, NULL, 0.785398163397448, 1.0, 0.0, 41, NULL, 0.05}, {3, NULL, 0.785398163397448, 1.0, 0.0, 41, NULL, 0.05}};
static const friq_dim_desc CP_DIMS[4] = {   /* cartpole.c:272-384 */
    {2, NULL, 2.0, 1.0, 1.0, 1001, NULL, 0.016}, {3, NULL, 1.0, 1.0, 0.0, 1001, NULL, 0.032},
    {9, CP_T, 0.0524, 21.485917317405871, 0.0, 1001, NULL, 0.0031415926535897933}, {2, NULL, 2.0, 1.0, 0.0, 1001, NULL, 0.016}};

int friq_problem_desc_builtin(int env, friq_problem_desc *d)
{
    memset(d, 0, sizeof *d);
    d->env = env;
    d->gamma = 1.0; d->qdiff_pos_boundary = 1.0; d->rule_weight_considered_significant_for_update = 0.05;   /* frirl_types_def.h:22-77 */
    d->skip_rules = 1; d->no_random = 1; d->max_episodes = 1000; d->max_steps = 1000;
    switch (env) {
    case FRIQ_ENV_MOUNTAINCAR:             /* mountaincar.c:220-279 */
        d->statedims_len = 2; d->statedims = MC_DIMS;
        d->actiondim = (friq_dim_desc){3, NULL, 1.0, 0.0, 0.0, 41, NULL, 0.1};
        d->alpha = 0.5; d->epsilon = 0.01; d->qdiff_neg_boundary = -4.0; d->qdiff_final_tolerance = 500.0; d->reward_good_above = -5000.0;
        return 0;
    case FRIQ_ENV_ACROBOT:                 /* acrobot.c:242-324 */
        d->statedims_len = 4; d->statedims = AC_DIMS;
        d->actiondim = (friq_dim_desc){3, NULL, 1.0, 0.0, 0.0, 41, NULL, 0.1};
        d->alpha = 0.5; d->epsilon = 0.001; d->qdiff_neg_boundary = -200.0; d->qdiff_final_tolerance = 50.0; d->reward_good_above = 0.0;
        return 0;
    case FRIQ_ENV_CARTPOLE:                /* cartpole.c:252-402 */
        d->statedims_len = 4; d->statedims = CP_DIMS;
        d->actiondim = (friq_dim_desc){21, CP_A, 0.1, 0.0, 0.0, 1001, NULL, 0.008};
        d->alpha = 0.3; d->epsilon = 0.001; d->qdiff_neg_boundary = -200.0; d->qdiff_final_tolerance = 250.0; d->reward_good_above = 0.0;
        return 0;
    }
    FAIL(1, "friq_problem_desc_builtin: unknown environment %d", env);
}

/* ------------------------------------------------------------------------ table generation */
/* symmetric fixed-resolution array, frirl_gen_fixres_arr (frirl_app_helpers.c:32-44) */
static void fixres_fill(double *dst, int n, double step)
{
    const double first = -((n - 1) * step) / 2;
    int i = 0;
    while (i < n / 2 + 1) { dst[i] = first + step * i; i++; }
    while (i < n) { dst[i] = dst[n - 1 - i] * -1; i++; }
}

/* nearest universe index, get_vag_abs_min_i_fixres (min.inl:71-92) */
static int nearest_index(const double *u, int U, int len, double div, double x)
{
    int low = (int)((x - u[0]) / div);
    if (low < 0) return 0;
    if (low >= len) return len - 1;
    if (low + 1 >= U) return low;
    return (fabs(u[low] - x) <= fabs(u[low + 1] - x)) ? low : low + 1;
}

/* One row of the vague environment.  pts[] are (position, S_left, S_right) scaling points;
 * the scaling function between two points is constant when S_right(p) == S_left(p+1), else linear
 * (FIVEGScFunc.c:100-146 with nls = NAN); VE is its trapezoid primitive integral (FIVEGVagEnv.c:81-94). */
static void vague_env_row(const double *u, int U, const double *pts, int npts, double *ve)
{
    double *sc = (double *)malloc(sizeof(double) * U);
    const double *last = pts + 3 * (npts - 1);
    for (int j = 0; j < U; j++) {
        const double x = u[j];
        double v = 0.0;
        if (x < pts[0]) v = pts[1];
        for (const double *p = pts; p != last; p += 3)
            if (x >= p[0] && x < p[3]) v = (p[2] == p[4]) ? p[2] : ((p[4] - p[2]) / (p[3] - p[0])) * (x - p[0]) + p[2];
        if (x >= last[0]) v = (j == U - 1 && x == last[0]) ? last[1] : last[2];
        sc[j] = v;
    }
    ve[0] = 0;
    for (int j = 1; j < U; j++) ve[j] = ve[j - 1] + (u[j] - u[j - 1]) * (sc[j - 1] + sc[j]) * 0.5;
    free(sc);
}

/* hit_scan / act_hit of the problem block: an exact hit (scaled distance 0) is the same as equal dictionary indices when, in
 * every dimension, distinct dictionary entries have distinct VE values (and no squared difference can underflow).  Called
 * whenever a dictionary grows. */
static void refresh_hit_scan(friq_problem *P)
{
    friq_devprob *hp = &P->hp;
    int injective = 1;
    for (int k = 0; k < hp->K; k++)
        for (int i = 0; i < hp->dict_len[k]; i++)
            for (int j = i + 1; j < hp->dict_len[k]; j++)
                if (hp->dict_ve[k][i] == hp->dict_ve[k][j]) injective = 0;
    for (int a = 0; a < hp->A; a++) {
        hp->act_hit[a] = -1;
        for (int j = 0; j < hp->dict_len[hp->S]; j++)
            if (hp->dict_ve[hp->S][j] == hp->act_ve[a]) { hp->act_hit[a] = j; break; }
    }
    for (int k = 0; k < hp->S; k++) {
        hp->def_hit[k] = -1;
        for (int j = 0; j < hp->dict_len[k]; j++)
            if (hp->dict_ve[k][j] == hp->ve_def[k]) { hp->def_hit[k] = j; break; }
    }
    hp->hit_scan = (injective && hp->safe_range) ? 1 : 0;
}

static int dict_find_or_add(friq_problem *P, int k, double v)
{
    friq_devprob *hp = &P->hp;
    for (int j = 0; j < hp->dict_len[k]; j++)
        if (memcmp(&hp->dict_val[k][j], &v, sizeof v) == 0) return j;
    if (hp->dict_len[k] >= FRIQ_DICT_MAX(hp->K)) return -1;
    const int j = hp->dict_len[k]++;
    hp->dict_val[k][j] = v;
    hp->dict_ve[k][j] = P->ve[(size_t)k * hp->U + nearest_index(P->u + (size_t)k * hp->U, hp->U, hp->U, hp->udiv[k], v)];   /* five_add_rule.c:71-84 */
    return j;
}

int friq_problem_create(const friq_problem_desc *d, friq_problem **out)
{
    const int S = d->statedims_len, K = S + 1, A = d->actiondim.values_len, U = d->actiondim.universe_len;
    if (S < 1 || K > FRIQ_MAXK) FAIL(2, "friq_problem_create: %d state dimensions unsupported (1..%d: FIVE_MAX_NUM_OF_UNIVERSES is %d)", S, FRIQ_MAXK - 1, FRIQ_MAXK);
    if (A < 2 || A > FRIQ_MAXA) FAIL(2, "friq_problem_create: %d actions unsupported (max %d)", A, FRIQ_MAXA);
    for (int k = 0; k < S; k++) {
        if (d->statedims[k].universe_len != U) FAIL(2, "friq_problem_create: all universes must have the same length (frirl_init.c:37)");
        if (d->statedims[k].values_len < 2 || d->statedims[k].values_len > FRIQ_DICT_MAX(K)) FAIL(2, "friq_problem_create: state dim %d has %d values (2..%d supported with %d antecedents)", k, d->statedims[k].values_len, FRIQ_DICT_MAX(K), K);
    }
    friq_problem *P = (friq_problem *)calloc(1, sizeof *P);
    friq_devprob *hp = &P->hp;
    P->device = -1;
    P->u = (double *)malloc(sizeof(double) * K * U); P->ve = (double *)malloc(sizeof(double) * K * U);
    hp->S = S; hp->K = K; hp->A = A; hp->U = U; hp->p = K;                       /* FIVEInit.c:89-93 called with p = 0 */
    hp->env = d->env; hp->skip_rules = d->skip_rules; hp->no_random = d->no_random;
    hp->max_steps = d->max_steps; hp->max_episodes = d->max_episodes;
    hp->qdiff_pos = d->qdiff_pos_boundary; hp->qdiff_neg = d->qdiff_neg_boundary;
    hp->wsig = d->rule_weight_considered_significant_for_update;
    hp->qdiff_final_tol = d->qdiff_final_tolerance; hp->reward_good_above = d->reward_good_above;
    P->alpha = d->alpha; P->gamma = d->gamma; P->epsilon = d->epsilon;

    double pts[3 * FRIQ_MAXDICT];
    for (int k = 0; k < K; k++) {
        const friq_dim_desc *dim = (k < S) ? &d->statedims[k] : &d->actiondim;
        double *u = P->u + (size_t)k * U, *ve = P->ve + (size_t)k * U;
        if (dim->universe) memcpy(u, dim->universe, sizeof(double) * U); else fixres_fill(u, U, dim->universe_div);
        if (u[U - 1] - u[0] == 0.0) { friq_problem_destroy(P); FAIL(2, "friq_problem_create: universe %d has zero extent (frirl_init.c:160-168)", k); }
        hp->grid_len[k] = hp->dict_len[k] = dim->values_len;
        if (dim->values) memcpy(hp->dict_val[k], dim->values, sizeof(double) * dim->values_len);
        else fixres_fill(hp->dict_val[k], dim->values_len, dim->values_div);
        hp->u0[k] = u[0];
        hp->udiv[k] = (u[U - 1] - u[0]) / (unsigned)(U - 1);                      /* FIVEInit.c:244-248 */
        if (k < S) {                                                              /* frirl_init_ve.c:36-57 */
            for (int i = 0; i < dim->values_len; i++) { pts[3 * i] = hp->dict_val[k][i]; pts[3 * i + 1] = pts[3 * i + 2] = dim->values_steep; }
        } else {                                                                  /* frirl_init_ve.c:82-92: integer steepness (A-1)/2 */
            const double ratio = 1.0 / (A - 1) * 2.0;
            for (int i = 0; i < A; i++) { pts[3 * i] = i * ratio - 1.0; pts[3 * i + 1] = pts[3 * i + 2] = (A - 1) / 2; }
        }
        vague_env_row(u, U, pts, dim->values_len, ve);
        for (int j = 0; j < dim->values_len; j++)
            hp->dict_ve[k][j] = ve[nearest_index(u, U, U, hp->udiv[k], hp->dict_val[k][j])];
        hp->values_def[k] = dim->values_def;
        hp->ve_def[k] = ve[nearest_index(u, U, U, hp->udiv[k], dim->values_def)];
        hp->q_abs0[k] = fabs(hp->dict_val[k][0]); hp->q_div[k] = dim->values_div; hp->q_rdiv[k] = 1.0 / dim->values_div;
        /* frirl_check_possible_states.c:68 would refine the grid when epsilon <= step/2; the device keeps grids fixed */
        for (int j = 1; j < dim->values_len; j++)
            if (dim->values_div <= fabs(hp->dict_val[k][j] - hp->dict_val[k][j - 1]) / 2) {
                friq_problem_destroy(P);
                FAIL(2, "friq_problem_create: dim %d: values_div <= half the grid step would trigger grid refinement (unsupported)", k);
            }
    }
    hp->qp_reuse = 1;
    for (int a = 0; a < A; a++) {
        const double *ua = P->u + (size_t)S * U, *vea = P->ve + (size_t)S * U;
        hp->act_val[a] = hp->dict_val[S][a]; hp->act_dict[a] = a;
        hp->act_ve[a] = vea[nearest_index(ua, U, U - 1, hp->udiv[S], hp->act_val[a])];          /* frirl_init.c:156-158 */
        hp->act_ve_eval[a] = vea[nearest_index(ua, U, U, hp->udiv[S], hp->act_val[a])];
        if (hp->act_ve[a] != hp->act_ve_eval[a]) hp->qp_reuse = 0;
    }
    /* every squared scaled distance the kernels can meet is a sum over dimensions of (VE_k[i] - VE_k[j])^2 */
    {
        double smin = INFINITY, smax = 0.0;
        for (int k = 0; k < K; k++) {
            const double *ve = P->ve + (size_t)k * U;
            double lo = ve[0], hi = ve[0], mind = INFINITY;
            for (int j = 0; j < U; j++) {
                if (ve[j] < lo) lo = ve[j];
                if (ve[j] > hi) hi = ve[j];
                for (int i = j + 1; i < U && i <= j + 1; i++) { const double d = fabs(ve[i] - ve[j]); if (d > 0 && d < mind) mind = d; }
            }
            if (!(ve[U - 1] >= ve[0])) mind = 0;      /* not monotone: give up on the static bound */
            for (int j = 1; j < U; j++) if (ve[j] < ve[j - 1]) mind = 0;
            if (mind * mind < smin) smin = mind * mind;
            smax += (hi - lo) * (hi - lo);
        }
        hp->safe_range = (smin >= 0x1p-200 && smax <= 0x1p200 && isfinite(smax)) ? 1 : 0;
    }
    /* corner rules in binary-counter order (frirl_init_rb.c:99-126) */
    hp->ncorner = 1 << K;
    for (int r = 0; r < hp->ncorner; r++) {
        unsigned w = 0;
        for (int k = 0; k < K; k++) {
            int lo = 0, hi = 0;
            for (int j = 1; j < hp->grid_len[k]; j++) { if (hp->dict_val[k][j] < hp->dict_val[k][lo]) lo = j; if (hp->dict_val[k][j] > hp->dict_val[k][hi]) hi = j; }
            w |= (unsigned)(((r / (hp->ncorner >> (k + 1))) % 2 == 0) ? lo : hi) << (FRIQ_BITS_OF(K) * k);
        }
        hp->corner[r] = w;
    }
    refresh_hit_scan(P);
    /* the episode start state as dictionary entries: its observation word is then always known (friq_obs_word) and the
     * memoised weights cover the first step of an episode too; no rule ever holds these indices unless a loaded rule base does */
    if (hp->hit_scan) {
        int added = 0;
        for (int k = 0; k < S; k++)
            if (hp->def_hit[k] < 0 && hp->dict_len[k] < FRIQ_DICT_MAX(K)) { dict_find_or_add(P, k, hp->values_def[k]); added = 1; }
        if (added) refresh_hit_scan(P);
    }
    *out = P;
    return 0;
}

/* A problem from caller-supplied universes and vague environments (FIVEInit, FIVEInit.c:55-347): no grids, no
 * actions, no environment - only the FIVE-level per-call entry points make sense on it. */
int friq_problem_create_from_tables(int K, int U, int p, const double *u, const double *ve, friq_problem **out)
{
    if (K < 2 || K > FRIQ_MAXK) FAIL(2, "friq_problem_create_from_tables: %d antecedent dimensions unsupported (2..%d)", K, FRIQ_MAXK);
    friq_problem *P = (friq_problem *)calloc(1, sizeof *P);
    friq_devprob *hp = &P->hp;
    P->device = -1;
    P->u = (double *)malloc(sizeof(double) * K * U); P->ve = (double *)malloc(sizeof(double) * K * U);
    memcpy(P->u, u, sizeof(double) * K * U); memcpy(P->ve, ve, sizeof(double) * K * U);
    hp->S = K - 1; hp->K = K; hp->A = 0; hp->U = U; hp->p = p ? p : K;
    hp->env = FRIQ_ENV_NONE; hp->qp_reuse = 0; hp->safe_range = 0; hp->ncorner = 0;
    hp->wsig = 0.05; hp->qdiff_pos = INFINITY; hp->qdiff_neg = -INFINITY;
    for (int k = 0; k < K; k++) {
        hp->u0[k] = u[(size_t)k * U];
        hp->udiv[k] = (u[(size_t)k * U + U - 1] - u[(size_t)k * U]) / (unsigned)(U - 1);
    }
    P->alpha = 0.5; P->gamma = 1.0;
    *out = P;
    return 0;
}

/* learning parameters a caller may change between calls on the classic API (public fields of struct frirl_desc) */
int friq_problem_set_learning_params(friq_problem *P, double qdiff_pos, double qdiff_neg, double wsig, int skip_rules)
{
    friq_devprob *hp = &P->hp;
    if (hp->qdiff_pos == qdiff_pos && hp->qdiff_neg == qdiff_neg && hp->wsig == wsig && hp->skip_rules == skip_rules) return 0;
    hp->qdiff_pos = qdiff_pos; hp->qdiff_neg = qdiff_neg; hp->wsig = wsig; hp->skip_rules = skip_rules;
    if (P->device != -1) { DEV(friqd_set_device(P->device)); DEV(friqd_device_sync()); DEV(friqd_h2d(P->dp, hp, sizeof(friq_devprob))); }
    return 0;
}

/* (Re)build the memoised Shepard weights of the problem on its device and publish the problem block.  A learner's query
 * points and rule antecedents both live on the finite dictionary grid, so the weight 1/d^p of a (query cell, rule cell)
 * pair is a constant of the problem: the fused kernel looks it up instead of running the sqrt -> x87 power -> reciprocal
 * chain per rule and pair.  The table is computed on the device by the same functions (friq_dist2, friq_shepard_weight).
 * Only for problems where exact hits are index equality (hit_scan) and Q(s',a') is actconc[a'] (qp_reuse), i.e. where the
 * fused kernel's policy 2 runs (friq_batch_set_policy(b, 1) runs the computed chains instead: parity tests). */
#define FRIQ_WTAB_MAX_BYTES (64ul << 20)
static int problem_publish(friq_problem *P)
{
    friq_devprob *hp = &P->hp;
    if (P->d_wtab) { DEV(friqd_device_sync()); DEV(friqd_free(P->d_wtab)); P->d_wtab = NULL; }
    hp->wtab = NULL; hp->wcells = 0; hp->wstride = 0;
    if ((hp->env == FRIQ_ENV_MOUNTAINCAR || hp->env == FRIQ_ENV_ACROBOT) && hp->hit_scan && hp->qp_reuse) {
        unsigned long cells = 1; int ok = 1;
        for (int k = 0; k < hp->K; k++) { hp->wlin[k] = (int)cells; cells *= (unsigned long)hp->dict_len[k]; if (cells > 65535ul) { ok = 0; break; } }
        for (int a = 0; a < hp->A; a++) if (hp->act_hit[a] < 0) ok = 0;
        /* one column more than cells, +0.0 in every row: the "cell" of the slots between a rule base's last rule and the end of
         * its block of 8 and of the blocks behind it (friq_device.cuh: friq_zero_cell) */
        const unsigned long stride = (cells + 2ul) & ~1ul;
        if (ok && cells < 65535ul && cells * stride * 8ul <= FRIQ_WTAB_MAX_BYTES) {
            DEV(friqd_h2d(P->dp, hp, sizeof(friq_devprob)));          /* the build kernel reads the dictionaries */
            DEV(friqd_malloc((void **)&P->d_wtab, cells * stride * 8ul));
            DEV(friqd_memset(P->d_wtab, 0, cells * stride * 8ul));
            hp->wcells = (int)cells; hp->wstride = (int)stride;
            DEV(friqd_build_wtab(hp, P->dp, P->d_wtab));
            hp->wtab = P->d_wtab;
        }
    }
    DEV(friqd_h2d(P->dp, hp, sizeof(friq_devprob)));
    return 0;
}

static int problem_upload(friq_problem *P, int device)
{
    friq_devprob *hp = &P->hp;
    const size_t tb = sizeof(double) * hp->K * hp->U;
    if (P->device != -1 && P->device != device) FAIL(3, "friq: a problem is bound to the first device it is used on (%d)", P->device);
    DEV(friqd_set_device(device));
    if (P->device == -1) {
        DEV(friqd_malloc((void **)&P->d_u, tb)); DEV(friqd_malloc((void **)&P->d_ve, tb)); DEV(friqd_malloc((void **)&P->dp, sizeof(friq_devprob)));
        DEV(friqd_h2d(P->d_u, P->u, tb)); DEV(friqd_h2d(P->d_ve, P->ve, tb));
        hp->u = P->d_u; hp->ve = P->d_ve;
        P->device = device;
        return problem_publish(P);
    }
    DEV(friqd_h2d(P->dp, hp, sizeof(friq_devprob)));
    return 0;
}

void friq_problem_destroy(friq_problem *P)
{
    if (!P) return;
    if (P->device != -1) {
        friqd_set_device(P->device); friqd_free(P->d_u); friqd_free(P->d_ve); friqd_free(P->dp);
        if (P->d_wtab) friqd_free(P->d_wtab);
    }
    free(P->u); free(P->ve); free(P);
}

int friq_problem_dims(const friq_problem *P, int *dims)
{
    dims[0] = P->hp.S; dims[1] = P->hp.K; dims[2] = P->hp.A; dims[3] = P->hp.U; dims[4] = P->hp.p; dims[5] = P->hp.ncorner;
    return 0;
}
int friq_problem_tables(const friq_problem *P, double *u, double *ve, double *udiv, double *act_val, double *act_ve)
{
    const size_t n = (size_t)P->hp.K * P->hp.U;
    if (u) memcpy(u, P->u, sizeof(double) * n);
    if (ve) memcpy(ve, P->ve, sizeof(double) * n);
    if (udiv) memcpy(udiv, P->hp.udiv, sizeof(double) * P->hp.K);
    if (act_val) memcpy(act_val, P->hp.act_val, sizeof(double) * P->hp.A);
    if (act_ve) memcpy(act_ve, P->hp.act_ve, sizeof(double) * P->hp.A);
    return 0;
}

/* ------------------------------------------------------------------------------- batches */
int friq_device_count(int *count) { DEV(friqd_device_count(count)); return 0; }

/* device buffers of one placement: rule bases + one slab for the per-agent scalars
 * (3*MAXK + 9 rows of 8 bytes, 12 rows of 4 bytes) */
static int alloc_state(friq_devstate *st, int clear)
{
    const size_t slots = (size_t)st->n_pad * st->cap, np = (size_t)st->n_pad;
    const size_t rows8 = 3 * FRIQ_MAXK + 9, rows4 = 12;
    void *slab;
    DEV(friqd_malloc((void **)&st->rbw, slots * 4));
    DEV(friqd_malloc((void **)&st->rbq, slots * 8 + 8 * FRIQ_TILE * 8));   /* slack: the table-mode loop reads whole blocks of 8 rules */
    DEV(friqd_malloc(&slab, np * (rows8 * 8 + rows4 * 4)));
    if (clear) { DEV(friqd_memset(st->rbw, 0, slots * 4)); DEV(friqd_memset(st->rbq, 0, slots * 8)); }
    DEV(friqd_memset(slab, 0, np * (rows8 * 8 + rows4 * 4)));
    double *p8 = (double *)slab;
    st->s = p8; p8 += FRIQ_MAXK * np; st->qs = p8; p8 += FRIQ_MAXK * np; st->vo = p8; p8 += FRIQ_MAXK * np;
    st->ep_return = p8; p8 += np; st->last_return = p8; p8 += np; st->prev_reward = p8; p8 += np;
    st->alpha = p8; p8 += np; st->gamma = p8; p8 += np; st->epsilon = p8; p8 += np;
    st->total_steps = (unsigned long long *)p8; p8 += np; st->sum_rules = (unsigned long long *)p8; p8 += np; st->seed = (unsigned long long *)p8; p8 += np;
    int *p4 = (int *)p8;
    st->aidx = p4; p4 += np; st->flags = p4; p4 += np; st->R = p4; p4 += np; st->step_in_ep = p4; p4 += np;
    st->last_steps = p4; p4 += np; st->last_rules = p4; p4 += np; st->prev_steps = p4; p4 += np; st->prev_numru = p4; p4 += np;
    st->ep_num = (unsigned int *)p4; p4 += np; st->agent_of_slot = p4; p4 += np; st->hitq = p4; p4 += np;
    return 0;
}
static void free_state(friq_devstate *st)
{
    friqd_free(st->rbw); friqd_free(st->rbq); if (st->prevq) friqd_free(st->prevq);
    if (st->rbc) friqd_free(st->rbc);
    if (st->cmap) friqd_free(st->cmap);
    if (st->wref) friqd_free(st->wref);
    friqd_free(st->s);        /* slab base */
}

/* move every agent to a new slot: by_rules = 1 sorts by descending rule count, 0 restores identity */
static int retile(friq_batch *b, int by_rules, void *stream)
{
    if (!by_rules && !b->permuted) return 0;
    if (!b->have_st2) {
        b->st2 = b->st; b->st2.prevq = NULL; b->st2.s0 = NULL; b->st2.rbc = NULL; b->st2.cmap = NULL; b->st2.wref = NULL;
        int rc = alloc_state(&b->st2, 0); if (rc) return rc;
        if (b->st.s0) DEV(friqd_malloc((void **)&b->st2.s0, 8 * FRIQ_MAXK * (size_t)b->st.n_pad));
        DEV(friqd_malloc((void **)&b->d_work, sizeof(int) * (FRIQ_MAXGROUPS * ((size_t)b->st.cap + 2 + FRIQ_TILE) + (size_t)b->st.n_pad)));
        b->have_st2 = 1;
    }
    if (b->st.prevq && !b->st2.prevq) DEV(friqd_malloc((void **)&b->st2.prevq, (size_t)b->st.n_pad * b->st.cap * 8));
    b->st2.sw = b->st.sw;
    DEV(friqd_retile(b->st, b->st2, b->d_work, by_rules, stream));
    friq_devstate t = b->st; b->st = b->st2; b->st2 = t;
    b->permuted = by_rules;
    return 0;
}
/* the slots [lo, lo + n) of a placement as a stand-alone state (lo is a multiple of the tile size) */
static friq_devstate state_view(const friq_devstate *st, long lo, long n)
{
    friq_devstate v = *st;
    const size_t ro = (size_t)(lo / FRIQ_TILE) * st->cap * FRIQ_TILE;
    v.n = n;
    v.rbw += ro; v.rbq += ro; v.sw += ro; if (v.prevq) v.prevq += ro;
    if (v.rbc) { v.rbc += (size_t)(lo / FRIQ_TILE) * (size_t)friqd_rbc_tile_slots(st->cap); v.cmap += (size_t)(lo / FRIQ_TILE) * st->idx_cells * FRIQ_TILE; }
    v.s += lo; v.qs += lo; v.vo += lo;
    v.aidx += lo; v.flags += lo; v.R += lo; v.step_in_ep += lo; v.last_steps += lo; v.last_rules += lo; v.prev_steps += lo; v.prev_numru += lo;
    v.ep_num += lo; v.ep_return += lo; v.last_return += lo; v.prev_reward += lo;
    v.total_steps += lo; v.sum_rules += lo; v.seed += lo; v.alpha += lo; v.gamma += lo; v.epsilon += lo;
    v.agent_of_slot += lo; v.hitq += lo; if (v.s0) v.s0 += lo;
    return v;
}
/* the per-agent rule-cell / cell-to-rule indexes of the fused kernel's table mode, sized for the problem's current table */
static int ensure_index_one(friq_devstate *st, int cells)
{
    if (st->rbc && st->idx_cells == cells) return 0;
    if (st->rbc) { DEV(friqd_device_sync()); DEV(friqd_free(st->rbc)); DEV(friqd_free(st->cmap)); st->rbc = NULL; st->cmap = NULL; }
    st->idx_cells = cells;
    if (cells > 0) {
        DEV(friqd_malloc((void **)&st->rbc, (size_t)(st->n_pad / FRIQ_TILE) * (size_t)friqd_rbc_tile_slots(st->cap) * 2));
        DEV(friqd_malloc((void **)&st->cmap, (size_t)st->n_pad * cells * 2));
    }
    return 0;
}
static int ensure_index(friq_batch *b)
{
    const int cells = b->P->hp.wtab ? b->P->hp.wcells : 0;
    int rc = ensure_index_one(&b->st, cells); if (rc) return rc;
    if (b->have_st2) { rc = ensure_index_one(&b->st2, cells); if (rc) return rc; }
    return 0;
}
/* FIVERB.weights of every agent (friq_devstate.wref), zero like the reference's fresh allocation (FIVEInit.c:334) */
static int ensure_wref(friq_batch *b)
{
    if (b->st.wref) return 0;
    const size_t bytes = (size_t)b->st.n_pad * b->st.cap * 8;
    DEV(friqd_malloc((void **)&b->st.wref, bytes)); DEV(friqd_memset(b->st.wref, 0, bytes));
    return 0;
}
#define IDENTITY(b) do { if ((b)->permuted) { DEV(friqd_device_sync()); int rc_ = retile((b), 0, NULL); if (rc_) return rc_; DEV(friqd_device_sync()); } } while (0)

int friq_batch_create(friq_problem *P, int device, long n, int cap, uint64_t seed0, long first_agent, friq_batch **out)
{
    if (n < 1) FAIL(4, "friq_batch_create: n_agents must be >= 1");
    if (cap < P->hp.ncorner) FAIL(4, "friq_batch_create: rule_capacity %d < %d initial corner rules", cap, P->hp.ncorner);
    int rc = problem_upload(P, device); if (rc) return rc;
    friq_batch *b = (friq_batch *)calloc(1, sizeof *b);
    b->P = P; b->device = device; b->seed0 = seed0; b->first_agent = first_agent;
    b->retile_interval = 125;
    /* evaluation policy of the fused kernel (friqd_run_steps variant): mountaincar agents find an exact-hit rule for > 90 % of
     * the (observation, action) pairs they visit, so exact hits first + compacted Shepard sums win by a wide margin; for
     * acrobot (13-53 % hits) they still measure 16 % faster than the lock-step pass; cartpole (A = 21, off-grid observations)
     * keeps the lock-step pass.  friq_batch_set_policy() overrides. */
    b->variant = ((P->hp.env == FRIQ_ENV_MOUNTAINCAR || P->hp.env == FRIQ_ENV_ACROBOT) && P->hp.hit_scan) ? 2 : 0;
    friq_devstate *st = &b->st;
    st->n = n; st->n_pad = (n + FRIQ_TILE - 1) / FRIQ_TILE * FRIQ_TILE; st->cap = cap;
    friq_batch_set_groups(b, 0);
    { int rc2 = alloc_state(st, 1); if (rc2) return rc2; }
    const size_t slots = (size_t)st->n_pad * cap;
    DEV(friqd_malloc((void **)&st->sw, slots * 8)); DEV(friqd_memset(st->sw, 0, slots * 8));
    DEV(friqd_malloc((void **)&st->ctr, 8 * FRIQ_NCTR)); DEV(friqd_memset(st->ctr, 0, 8 * FRIQ_NCTR));
    *out = b;
    return friq_batch_reset(b);
}

/* evaluation policy of the fused kernel: 0 lock-step pass over all actions, 1 exact hits first + compacted Shepard chains,
 * 2 the same with memoised weights (table mode).  A policy the problem does not qualify for falls back to the next lower one; results never depend on it
 * (tests/test_gpu_parity.py compares all of them with the oracle). */
int friq_batch_set_policy(friq_batch *b, int policy)
{
    if (policy < 0 || policy > 2) FAIL(4, "friq_batch_set_policy: policy %d (0, 1 or 2)", policy);
    b->variant = policy;
    return 0;
}
/* agent groups of a re-tiled run, each sorted and launched on its own stream; 0 = default (one per 32,768 agents, at most 8) */
int friq_batch_set_groups(friq_batch *b, int groups)
{
    int g = groups > 0 ? groups : (int)(b->st.n ? b->st.n / 32768 : 1);
    if (g < 1) g = 1;
    if (g > FRIQ_MAXGROUPS) g = FRIQ_MAXGROUPS;
    while (g > 1 && b->st.n && (b->st.n / g) < 1024) g--;
    if (b->permuted) { DEV(friqd_set_device(b->device)); IDENTITY(b); }
    b->groups = g;
    return 0;
}

int friq_batch_reset(friq_batch *b)
{
    DEV(friqd_set_device(b->device));
    DEV(friqd_device_sync());
    b->permuted = 0;
    DEV(friqd_init_agents(b->P->dp, b->st, b->seed0, b->first_agent, b->P->alpha, b->P->gamma, b->P->epsilon, NULL));
    DEV(friqd_sync(NULL));
    return 0;
}

void friq_batch_destroy(friq_batch *b)
{
    if (!b) return;
    friqd_set_device(b->device);
    for (int g = 0; g < FRIQ_MAXGROUPS; g++) {
        if (b->gstream[g]) friqd_stream_destroy(b->gstream[g]);
        if (b->gdone[g]) friqd_event_destroy(b->gdone[g]);
    }
    if (b->gstart) friqd_event_destroy(b->gstart);
    friqd_free(b->st.sw);
    friqd_free(b->st.ctr);
    if (b->d_srb) friqd_free(b->d_srb);
    if (b->d_memo) friqd_free(b->d_memo);
    if (b->d_start) friqd_free(b->d_start);
    if (b->d_start_scratch) friqd_free(b->d_start_scratch);
    if (b->st.s0) friqd_free(b->st.s0);
    if (b->have_st2 && b->st2.s0) friqd_free(b->st2.s0);
    free_state(&b->st);
    if (b->have_st2) { free_state(&b->st2); friqd_free(b->d_work); }
    if (b->d_hyp) friqd_free(b->d_hyp);
    if (b->d_tmp8) { friqd_free(b->d_tmp8); friqd_free(b->d_tmp4a); friqd_free(b->d_tmp4b); }
    void *tmp[6] = { b->d_in0, b->d_in1, b->d_in2, b->d_out0, b->d_iout, b->d_w };
    for (int i = 0; i < 6; i++) if (tmp[i]) friqd_free(tmp[i]);
    free(b);
}

int friq_batch_set_hyper(friq_batch *b, const double *alpha, const double *gamma, const double *epsilon)
{
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    const size_t nb = sizeof(double) * b->st.n;
    if (alpha) DEV(friqd_h2d(b->st.alpha, alpha, nb));
    if (gamma) DEV(friqd_h2d(b->st.gamma, gamma, nb));
    if (epsilon) DEV(friqd_h2d(b->st.epsilon, epsilon, nb));
    return 0;
}

int friq_batch_set_rulebase(friq_batch *b, long agent, int R, const double *rant, const double *rconc)
{
    friq_problem *P = b->P; const int K = P->hp.K;
    if (R < 0 || R > b->st.cap) FAIL(5, "friq_batch_set_rulebase: %d rules exceed the capacity %d", R, b->st.cap);
    if (agent >= b->st.n) FAIL(5, "friq_batch_set_rulebase: agent %ld out of range", agent);
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    unsigned int *w = (unsigned int *)malloc(sizeof(unsigned int) * (R + 1));
    int grew = 0;
    for (int i = 0; i < R; i++) {
        unsigned int word = 0;
        for (int k = 0; k < K; k++) {
            const int before = P->hp.dict_len[k];
            const int j = dict_find_or_add(P, k, rant[(size_t)i * K + k]);
            if (j < 0) { free(w); FAIL(5, "friq_batch_set_rulebase: more than %d distinct antecedent values in dimension %d", FRIQ_DICT_MAX(K), k); }
            grew |= (P->hp.dict_len[k] != before);
            word |= (unsigned int)j << (FRIQ_BITS_OF(K) * k);
        }
        w[i] = word;
    }
    if (grew) { refresh_hit_scan(P); int rcp_ = problem_publish(P); if (rcp_) { free(w); return rcp_; } }
    if (agent >= 0) {
        const size_t off = (size_t)(agent / FRIQ_TILE) * b->st.cap * FRIQ_TILE + (agent % FRIQ_TILE);
        if (R) {
            DEV(friqd_h2d_2d(b->st.rbw + off, FRIQ_TILE * 4, w, 4, 4, R));
            DEV(friqd_h2d_2d(b->st.rbq + off, FRIQ_TILE * 8, rconc, 8, 8, R));
        }
        DEV(friqd_h2d(b->st.R + agent, &R, sizeof(int)));
        { const int unknown = -2; DEV(friqd_h2d(b->st.hitq + agent, &unknown, sizeof(int))); }
    } else {
        unsigned int *dw; double *dq;
        DEV(friqd_malloc((void **)&dw, 4 * (R + 1))); DEV(friqd_malloc((void **)&dq, 8 * (R + 1)));
        if (R) { DEV(friqd_h2d(dw, w, 4 * R)); DEV(friqd_h2d(dq, rconc, 8 * R)); }
        DEV(friqd_broadcast_rb(b->st, dw, dq, R, NULL)); DEV(friqd_sync(NULL));
        friqd_free(dw); friqd_free(dq);
    }
    free(w);
    return 0;
}

/* ONE read-only rule base for all agents (BASELINE config 4): antecedent values are resolved to VE values here,
 * exactly as five_add_rule.c:71-84 caches them; R = 0 returns to the private rule bases */
int friq_batch_set_shared_rulebase(friq_batch *b, int R, const double *rant, const double *rconc)
{
    friq_problem *P = b->P; const int K = P->hp.K, U = P->hp.U;
    DEV(friqd_set_device(b->device));
    DEV(friqd_device_sync());
    if (b->d_srb) { friqd_free(b->d_srb); b->d_srb = NULL; }
    if (b->d_memo) { friqd_free(b->d_memo); b->d_memo = NULL; }
    b->shared_R = 0; b->memo_cells = 0; b->start_dirty = 1;
    if (R <= 0) return 0;
    double *rows = (double *)malloc(sizeof(double) * (size_t)R * (K + 1));
    for (int i = 0; i < R; i++) {
        for (int k = 0; k < K; k++)
            rows[(size_t)i * (K + 1) + k] = P->ve[(size_t)k * U + nearest_index(P->u + (size_t)k * U, U, U, P->hp.udiv[k], rant[(size_t)i * K + k])];
        rows[(size_t)i * (K + 1) + K] = rconc[i];
    }
    int rc = friqd_malloc((void **)&b->d_srb, sizeof(double) * (size_t)R * (K + 1));
    if (!rc) rc = friqd_h2d(b->d_srb, rows, sizeof(double) * (size_t)R * (K + 1));
    free(rows);
    if (rc) FAIL(rc, "friq_batch_set_shared_rulebase: %s", friqd_error_string(rc));
    b->shared_R = R;
    /* memoised conclusions: one entry per (observation grid cell, action) - the built-in quantisers return grid points, so after the
     * first step of an episode every conclusion an agent needs is one of these (k_shared_memo_build, k_infer_memo) */
    if (P->hp.env == FRIQ_ENV_MOUNTAINCAR || P->hp.env == FRIQ_ENV_ACROBOT) {
        unsigned long cells = 1;
        for (int k = 0; k < P->hp.S; k++) cells *= (unsigned long)P->hp.grid_len[k];
        if (cells * (unsigned long)P->hp.A * 8ul <= (64ul << 20)) {
            DEV(friqd_malloc((void **)&b->d_memo, cells * (unsigned long)P->hp.A * 8ul));
            int rcm = friqd_shared_memo_build(&P->hp, P->dp, b->d_srb, R, b->d_memo, (int)cells, NULL);
            if (!rcm) rcm = friqd_device_sync();
            if (rcm == 9001) { friqd_free(b->d_memo); b->d_memo = NULL; }           /* no memo kernel for this shape: the streaming pass serves it */
            else if (rcm) FAIL(rcm, "friq_batch_set_shared_rulebase: %s", friqd_error_string(rcm));
            else b->memo_cells = (int)cells;
        }
    }
    return 0;
}

/* per-agent episode start state s0 [n][S] (host, agent order); NULL restores values_def (frirl_episode.c:46-48) */
int friq_batch_set_initial_state(friq_batch *b, const double *s0)
{
    const long n = b->st.n; const size_t np = (size_t)b->st.n_pad; const int S = b->P->hp.S;
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    b->start_dirty = 1;
    if (!s0) { if (b->st.s0) { friqd_free(b->st.s0); b->st.s0 = NULL; } if (b->have_st2 && b->st2.s0) { friqd_free(b->st2.s0); b->st2.s0 = NULL; } return 0; }
    double *t = (double *)calloc(FRIQ_MAXK * np, sizeof(double));
    for (long i = 0; i < n; i++) for (int k = 0; k < S; k++) t[(size_t)k * np + i] = s0[(size_t)i * S + k];
    int rc = 0;
    if (!b->st.s0) rc = friqd_malloc((void **)&b->st.s0, 8 * FRIQ_MAXK * np);
    if (!rc && b->have_st2 && !b->st2.s0) rc = friqd_malloc((void **)&b->st2.s0, 8 * FRIQ_MAXK * np);
    if (!rc) rc = friqd_h2d(b->st.s0, t, 8 * FRIQ_MAXK * np);
    free(t);
    if (rc) FAIL(rc, "friq_batch_set_initial_state: %s", friqd_error_string(rc));
    return 0;
}

/* the same from a DEVICE array d_s0 [n][S] (e.g. a tensor the caller already holds in HBM), asynchronous on `stream` */
int friq_batch_set_initial_state_device(friq_batch *b, const void *d_s0, void *stream)
{
    const size_t np = (size_t)b->st.n_pad;
    if (!d_s0) FAIL(5, "friq_batch_set_initial_state_device: NULL (use friq_batch_set_initial_state(b, NULL) to restore values_def)");
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    b->start_dirty = 1;
    if (!b->st.s0) { DEV(friqd_malloc((void **)&b->st.s0, 8 * FRIQ_MAXK * np)); DEV(friqd_memset(b->st.s0, 0, 8 * FRIQ_MAXK * np)); }
    if (b->have_st2 && !b->st2.s0) DEV(friqd_malloc((void **)&b->st2.s0, 8 * FRIQ_MAXK * np));
    DEV(friqd_set_s0(b->st, (const double *)d_s0, b->P->hp.S, stream));
    return 0;
}

int friq_batch_get_rulebase(friq_batch *b, long agent, int *R_out, double *rant, double *rconc)
{
    const friq_devprob *hp = &b->P->hp; const int K = hp->K;
    if (agent < 0 || agent >= b->st.n) FAIL(5, "friq_batch_get_rulebase: agent %ld out of range", agent);
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    int R; DEV(friqd_d2h(&R, b->st.R + agent, sizeof(int)));
    *R_out = R;
    if (!R || (!rant && !rconc)) return 0;
    const size_t off = (size_t)(agent / FRIQ_TILE) * b->st.cap * FRIQ_TILE + (agent % FRIQ_TILE);
    if (rconc) DEV(friqd_d2h_2d(rconc, 8, b->st.rbq + off, FRIQ_TILE * 8, 8, R));
    if (rant) {
        unsigned int *w = (unsigned int *)malloc(sizeof(unsigned int) * R);
        int rc = friqd_d2h_2d(w, 4, b->st.rbw + off, FRIQ_TILE * 4, 4, R);
        if (rc) { free(w); FAIL(rc, "friq_batch_get_rulebase: %s", friqd_error_string(rc)); }
        for (int i = 0; i < R; i++)
            for (int k = 0; k < K; k++) rant[(size_t)i * K + k] = hp->dict_val[k][FRIQ_FIELD(w[i], k, K)];
        free(w);
    }
    return 0;
}

/* current raw environment state s [n][S] and current action index [n] of every agent (fep_ant, frirl_types.h:155) */
int friq_batch_get_state(friq_batch *b, double *s, int *action_index)
{
    const long n = b->st.n; const size_t np = (size_t)b->st.n_pad; const int S = b->P->hp.S;
    DEV(friqd_set_device(b->device));
    DEV(friqd_device_sync());
    IDENTITY(b);
    if (s) {
        double *t = (double *)malloc(8 * np * S);
        int rc = friqd_d2h(t, b->st.s, 8 * np * S);
        if (!rc) for (long i = 0; i < n; i++) for (int k = 0; k < S; k++) s[(size_t)i * S + k] = t[(size_t)k * np + i];
        free(t);
        if (rc) FAIL(rc, "friq_batch_get_state: %s", friqd_error_string(rc));
    }
    if (action_index) DEV(friqd_d2h(action_index, b->st.aidx, 4 * n));
    return 0;
}

int friq_batch_set_inserted(friq_batch *b, const int *inserted)
{
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    int *f = (int *)malloc(sizeof(int) * b->st.n);
    int rc = friqd_d2h(f, b->st.flags, sizeof(int) * b->st.n);
    if (!rc) { for (long i = 0; i < b->st.n; i++) f[i] = (f[i] & ~FRIQ_F_INSERTED) | (inserted[i] ? FRIQ_F_INSERTED : 0); rc = friqd_h2d(b->st.flags, f, sizeof(int) * b->st.n); }
    free(f);
    if (rc) FAIL(rc, "friq_batch_set_inserted: %s", friqd_error_string(rc));
    return 0;
}
int friq_batch_get_inserted(friq_batch *b, int *inserted)
{
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    DEV(friqd_d2h(inserted, b->st.flags, sizeof(int) * b->st.n));
    for (long i = 0; i < b->st.n; i++) inserted[i] = (inserted[i] & FRIQ_F_INSERTED) ? 1 : 0;
    return 0;
}

int friq_batch_run(friq_batch *b, long n_steps, int mode, void *stream)
{
    DEV(friqd_set_device(b->device));
    if (b->P->hp.env == FRIQ_ENV_HOST_CALLBACKS) FAIL(6, "friq_batch_run: the problem has no built-in environment (host callbacks): use the per-call entry points");
    if (b->shared_R > 0) {
        if (mode & FRIQ_RUN_LEARN) FAIL(6, "friq_batch_run: a shared rule base is read-only (clear it or drop FRIQ_RUN_LEARN)");
        IDENTITY(b);
        if (b->d_memo && b->variant != 0) {          /* memoised conclusions (default); policy 0 = one full pass per env step */
            if (!b->d_start) DEV(friqd_malloc((void **)&b->d_start, 8ul * (unsigned long)b->P->hp.A * (unsigned long)b->st.n_pad));
            if (!b->d_start_scratch) DEV(friqd_malloc(&b->d_start_scratch, friqd_shared_start_scratch_bytes(b->st.n_pad, b->P->hp.A)));
            if (b->start_dirty) { DEV(friqd_shared_start(&b->P->hp, b->P->dp, b->st, b->d_srb, b->shared_R, b->d_start, b->d_start_scratch, stream)); b->start_dirty = 0; }
            DEV(friqd_run_shared_memo(&b->P->hp, b->P->dp, b->st, b->d_memo, b->d_start, b->shared_R, n_steps, mode, stream));
            return 0;
        }
        DEV(friqd_run_shared(&b->P->hp, b->P->dp, b->st, b->d_srb, b->shared_R, n_steps, mode, stream));
        return 0;
    }
    { int rci = ensure_index(b); if (rci) return rci; }
    if (mode & FRIQ_RUN_KEEP_WEIGHTS) {
        if (mode & FRIQ_RUN_RETILE) FAIL(6, "friq_batch_run: FRIQ_RUN_KEEP_WEIGHTS cannot be combined with FRIQ_RUN_RETILE");
        IDENTITY(b);
        int rcw = ensure_wref(b); if (rcw) return rcw;
    }
    if ((mode & FRIQ_RUN_TRACK_CONVERGENCE) && !b->st.prevq) {
        const size_t slots = (size_t)b->st.n_pad * b->st.cap;
        DEV(friqd_malloc((void **)&b->st.prevq, slots * 8)); DEV(friqd_memset(b->st.prevq, 0, slots * 8));
    }
    if ((mode & FRIQ_RUN_RETILE) && !(mode & FRIQ_RUN_STOP_AT_EPISODE_END) && b->retile_interval > 0) {
        /* Rule bases grow in bursts: re-sorting every `retile_interval` steps keeps the lanes of a warp in
         * step.  The batch is cut into independent agent groups, each sorted within its own slot range and
         * issued on its own stream, so that the tail of one group's launch (a few warps with long rule
         * bases) overlaps the next group's launch instead of idling the GPU. */
        const int G = b->groups;
        if (b->permuted && b->permuted != 1 + G) { int rc = retile(b, 0, stream); if (rc) return rc; }   /* group layout changed */
        if (!b->have_st2) { int rc = retile(b, 1, stream); if (rc) return rc; rc = retile(b, 0, stream); if (rc) return rc; rc = ensure_index(b); if (rc) return rc; }  /* allocate buffers */
        if (b->st.prevq && !b->st2.prevq) DEV(friqd_malloc((void **)&b->st2.prevq, (size_t)b->st.n_pad * b->st.cap * 8));
        if (!b->gstart) DEV(friqd_event_create(&b->gstart));
        for (int g = 0; g < G; g++) if (!b->gstream[g]) {
            DEV(friqd_stream_create(&b->gstream[g])); DEV(friqd_event_create(&b->gdone[g]));
        }
        DEV(friqd_event_record(b->gstart, stream));
        const long tiles = b->st.n_pad / FRIQ_TILE, per = (tiles + G - 1) / G * FRIQ_TILE;
        const size_t wstride = (size_t)b->st.cap + 2 + per;
        for (int g = 0; g < G; g++) DEV(friqd_stream_wait_event(b->gstream[g], b->gstart));
        for (long done = 0; done < n_steps; done += b->retile_interval) {
            const long m = n_steps - done < b->retile_interval ? n_steps - done : b->retile_interval;
            b->st2.sw = b->st.sw;
            for (int g = 0; g < G; g++) {
                const long lo = g * per, cnt = (lo + per <= b->st.n ? per : b->st.n - lo);
                if (cnt <= 0) continue;
                friq_devstate src = state_view(&b->st, lo, cnt), dst = state_view(&b->st2, lo, cnt);
                DEV(friqd_retile(src, dst, b->d_work + g * wstride, 1, b->gstream[g]));
                DEV(friqd_run_steps(&b->P->hp, b->P->dp, dst, m, mode, b->variant, b->gstream[g]));
            }
            friq_devstate t = b->st; b->st = b->st2; b->st2 = t;
            b->permuted = 1 + G;
        }
        for (int g = 0; g < G; g++) { DEV(friqd_event_record(b->gdone[g], b->gstream[g])); DEV(friqd_stream_wait_event(stream, b->gdone[g])); }
        return 0;
    }
    if (mode & FRIQ_RUN_RETILE) { int rc = retile(b, 1, stream); if (rc) return rc; }
    DEV(friqd_run_steps(&b->P->hp, b->P->dp, b->st, n_steps, mode, b->variant, stream));
    return 0;
}
int friq_batch_set_retile_interval(friq_batch *b, long steps) { b->retile_interval = steps; return 0; }
int friq_batch_sync(friq_batch *b, void *stream) { DEV(friqd_set_device(b->device)); DEV(friqd_sync(stream)); DEV(friqd_device_sync()); return 0; }

int friq_batch_gather(friq_batch *b, double *last_return, int *last_steps, int *rules, int *episodes, int *flags,
                      long long *total_steps, long long *sum_rules)
{
    const long n = b->st.n;
    DEV(friqd_set_device(b->device));
    DEV(friqd_device_sync());
    IDENTITY(b);
    if (last_return) DEV(friqd_d2h(last_return, b->st.last_return, 8 * n));
    if (last_steps) DEV(friqd_d2h(last_steps, b->st.last_steps, 4 * n));
    if (rules) DEV(friqd_d2h(rules, b->st.R, 4 * n));
    if (episodes) { DEV(friqd_d2h(episodes, b->st.ep_num, 4 * n)); for (long i = 0; i < n; i++) episodes[i] -= 1; }
    if (flags) DEV(friqd_d2h(flags, b->st.flags, 4 * n));
    if (total_steps) DEV(friqd_d2h(total_steps, b->st.total_steps, 8 * n));
    if (sum_rules) DEV(friqd_d2h(sum_rules, b->st.sum_rules, 8 * n));
    return 0;
}

/* every agent starts a NEW episode at its next step (frirl_episode.c:46-82) whatever state its current one is in; rule bases,
 * counters and random streams are untouched.  Asynchronous on `stream`. */
int friq_batch_new_episodes(friq_batch *b, void *stream)
{
    DEV(friqd_set_device(b->device));
    DEV(friqd_new_episodes(b->st, stream));
    return 0;
}
/* statistics of a gathered per-agent result set of n agents (device arrays, e.g. what an all-gather produced): d_out3[0] = sum of
 * rules, [1] = sum of returns, [2] = max rules - one small kernel, deterministic; asynchronous on `stream` */
int friq_result_stats(int device, const void *d_ret_f64, const void *d_rules_i32, long n, void *d_out3_f64, void *stream)
{
    DEV(friqd_set_device(device));
    DEV(friqd_result_stats((const double *)d_ret_f64, (const int *)d_rules_i32, n, (double *)d_out3_f64, stream));
    return 0;
}

/* executed-work counters of the fused kernels since batch creation (or the last call with reset != 0): out[0..7] =
 * pairs that needed Shepard sums, rules summed with looked-up weights, rules summed with computed chains, pairs answered by an
 * exact-hit rule, rules appended, rules scanned by weighted updates, env steps, 0 */
int friq_batch_work_counters(friq_batch *b, unsigned long long *out, int reset)
{
    DEV(friqd_set_device(b->device));
    DEV(friqd_device_sync());
    if (out) DEV(friqd_d2h(out, b->st.ctr, 8 * FRIQ_NCTR));
    if (reset) DEV(friqd_memset(b->st.ctr, 0, 8 * FRIQ_NCTR));
    return 0;
}

int friq_batch_gather_device(friq_batch *b, void *d_ret, void *d_steps, void *d_rules, void *stream)
{
    const long n = b->st.n;
    DEV(friqd_set_device(b->device));
    if (b->permuted) {          /* report in agent order */
        if (d_ret) DEV(friqd_unpermute(b->st, b->st.last_return, d_ret, 8, stream));
        if (d_steps) DEV(friqd_unpermute(b->st, b->st.last_steps, d_steps, 4, stream));
        if (d_rules) DEV(friqd_unpermute(b->st, b->st.R, d_rules, 4, stream));
        return 0;
    }
    if (d_ret) DEV(friqd_d2d_async(d_ret, b->st.last_return, 8 * n, stream));
    if (d_steps) DEV(friqd_d2d_async(d_steps, b->st.last_steps, 4 * n, stream));
    if (d_rules) DEV(friqd_d2d_async(d_rules, b->st.R, 4 * n, stream));
    return 0;
}

int friq_batch_run_host(friq_batch *b, const double *alpha, const double *gamma, const double *epsilon, long n_steps, int mode,
                        double *last_return, int *last_steps, int *rules, void *stream)
{
    int rc = friq_batch_run_host_async(b, alpha, gamma, epsilon, n_steps, mode, last_return, last_steps, rules, stream);
    if (rc) return rc;
    DEV(friqd_sync(stream));
    return 0;
}
/* the same without the final synchronisation: copies in, run, copies out are all enqueued on `stream`; the host buffers (pinned
 * memory for real overlap) are valid once the stream has passed this point (an event recorded after the call) - a caller with two
 * buffer sets reads step i's results while step i + 1 runs */
int friq_batch_run_host_async(friq_batch *b, const double *alpha, const double *gamma, const double *epsilon, long n_steps, int mode,
                              double *last_return, int *last_steps, int *rules, void *stream)
{
    const long n = b->st.n;
    DEV(friqd_set_device(b->device));
    if (alpha || gamma || epsilon) {                     /* host arrays are in agent order */
        if (b->permuted) {                               /* agents live in re-tiled slots: stage in agent order, scatter through the placement */
            const size_t np = (size_t)b->st.n_pad;
            if (!b->d_hyp) DEV(friqd_malloc((void **)&b->d_hyp, 3 * 8 * np));
            double *t0 = b->d_hyp, *t1 = b->d_hyp + np, *t2 = b->d_hyp + 2 * np;
            if (alpha) DEV(friqd_h2d_async(t0, alpha, 8 * n, stream));
            if (gamma) DEV(friqd_h2d_async(t1, gamma, 8 * n, stream));
            if (epsilon) DEV(friqd_h2d_async(t2, epsilon, 8 * n, stream));
            DEV(friqd_permute_in3(b->st, alpha ? t0 : NULL, b->st.alpha, gamma ? t1 : NULL, b->st.gamma, epsilon ? t2 : NULL, b->st.epsilon, stream));
        } else {
            if (alpha) DEV(friqd_h2d_async(b->st.alpha, alpha, 8 * n, stream));
            if (gamma) DEV(friqd_h2d_async(b->st.gamma, gamma, 8 * n, stream));
            if (epsilon) DEV(friqd_h2d_async(b->st.epsilon, epsilon, 8 * n, stream));
        }
    }
    int rc = friq_batch_run(b, n_steps, mode, stream); if (rc) return rc;
    const void *s_ret = b->st.last_return, *s_steps = b->st.last_steps, *s_rules = b->st.R;
    if (b->permuted) {
        if (!b->d_tmp8) { DEV(friqd_malloc(&b->d_tmp8, 8 * b->st.n_pad)); DEV(friqd_malloc(&b->d_tmp4a, 4 * b->st.n_pad)); DEV(friqd_malloc(&b->d_tmp4b, 4 * b->st.n_pad)); }
        rc = friq_batch_gather_device(b, b->d_tmp8, b->d_tmp4a, b->d_tmp4b, stream); if (rc) return rc;
        s_ret = b->d_tmp8; s_steps = b->d_tmp4a; s_rules = b->d_tmp4b;
    }
    if (last_return) DEV(friqd_d2h_async(last_return, s_ret, 8 * n, stream));
    if (last_steps) DEV(friqd_d2h_async(last_steps, s_steps, 4 * n, stream));
    if (rules) DEV(friqd_d2h_async(rules, s_rules, 4 * n, stream));
    return 0;
}

/* ------------------------------------------------------------------- per-call entry points */
static int ensure(void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return 0;
    if (*p) friqd_free(*p);
    *cap = 0;
    int rc = friqd_malloc(p, need); if (rc) return rc;
    *cap = need;
    return 0;
}
static int percall_buffers(friq_batch *b)
{
    const size_t n = (size_t)b->st.n, need_in = n * (2 * FRIQ_MAXK + 4) * 8, need_out = n * (FRIQ_MAXA + 4) * 8;   /* classic step: K + S + 3 in, 4 + A out */
    if (b->cap_in < need_in) {
        size_t c0 = b->cap_in, c1 = b->cap_in, c2 = b->cap_in;
        DEV(ensure((void **)&b->d_in0, &c0, need_in)); DEV(ensure((void **)&b->d_in1, &c1, need_in)); DEV(ensure((void **)&b->d_in2, &c2, need_in));
        b->cap_in = need_in;
    }
    if (b->cap_out < need_out) {
        size_t c0 = b->cap_out, c1 = b->cap_out;
        DEV(ensure((void **)&b->d_out0, &c0, need_out)); DEV(ensure((void **)&b->d_iout, &c1, n * 4 > need_out ? n * 4 : need_out));
        b->cap_out = need_out;
    }
    return 0;
}

int friq_batch_vag_concl(friq_batch *b, const double *ant, double *conc, int *hit)
{
    const long n = b->st.n; const int K = b->P->hp.K;
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    int rc = percall_buffers(b); if (rc) return rc;
    DEV(friqd_h2d(b->d_in0, ant, 8ul * n * K));
    DEV(friqd_vag_concl(&b->P->hp, b->P->dp, b->st, b->d_in0, b->d_out0, b->d_iout, NULL, NULL));
    DEV(friqd_d2h(conc, b->d_out0, 8ul * n)); DEV(friqd_d2h(hit, b->d_iout, 4ul * n));
    return 0;
}

int friq_batch_vag_concl_weight(friq_batch *b, const double *ant, int *hit, double *weights)
{
    const long n = b->st.n; const int K = b->P->hp.K;
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    int rc = percall_buffers(b); if (rc) return rc;
    DEV(ensure((void **)&b->d_w, &b->cap_w, 8ul * n * b->st.cap));
    DEV(friqd_memset(b->d_w, 0, 8ul * n * b->st.cap));
    DEV(friqd_h2d(b->d_in0, ant, 8ul * n * K));
    DEV(friqd_vag_concl(&b->P->hp, b->P->dp, b->st, b->d_in0, NULL, b->d_iout, b->d_w, NULL));
    DEV(friqd_d2h(hit, b->d_iout, 4ul * n)); DEV(friqd_d2h(weights, b->d_w, 8ul * n * b->st.cap));
    return 0;
}

int friq_batch_get_best_action(friq_batch *b, const double *states, double *actconc, int *best)
{
    const long n = b->st.n; const int S = b->P->hp.S, A = b->P->hp.A;
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    int rc = percall_buffers(b); if (rc) return rc;
    DEV(friqd_h2d(b->d_in0, states, 8ul * n * S));
    DEV(friqd_best_action(&b->P->hp, b->P->dp, b->st, b->d_in0, b->d_out0, b->d_iout, NULL));
    DEV(friqd_d2h(actconc, b->d_out0, 8ul * n * A)); DEV(friqd_d2h(best, b->d_iout, 4ul * n));
    return 0;
}

int friq_batch_update_sarsa(friq_batch *b, const double *q_ant, const double *reward, const double *cur_q_ant)
{
    const long n = b->st.n; const int K = b->P->hp.K;
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    int rc = percall_buffers(b); if (rc) return rc;
    DEV(friqd_h2d(b->d_in0, q_ant, 8ul * n * K)); DEV(friqd_h2d(b->d_in1, reward, 8ul * n)); DEV(friqd_h2d(b->d_in2, cur_q_ant, 8ul * n * K));
    DEV(friqd_update_sarsa(&b->P->hp, b->P->dp, b->st, b->d_in0, b->d_in1, b->d_in2, NULL));
    DEV(friqd_sync(NULL));
    return 0;
}

/* one step of the classic API for every agent: proposed action for cur_q_states (forced[i] >= 0: the host's own
 * random draw, frirl_e_greedy_selection.c:28-34), then frirl_update_sarsa when learn != 0 (frirl_episode.c:148-160).
 * One H2D, one kernel, one D2H. */
int friq_batch_classic_step(friq_batch *b, const double *q_ant, const double *reward, const double *cur_q_states, const int *forced,
                            int learn, int *action, double *actconc, int *rules, int *inserted)
{
    const long n = b->st.n; const int S = b->P->hp.S, K = S + 1, A = b->P->hp.A, NIN = K + S + 3, NOUT = 4 + A;
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    int rc = percall_buffers(b); if (rc) return rc;
    double *h = (double *)malloc(sizeof(double) * n * (NIN > NOUT ? NIN : NOUT));
    for (long i = 0; i < n; i++) {
        double *x = h + (size_t)i * NIN;
        memcpy(x, q_ant + (size_t)i * K, sizeof(double) * K);
        x[K] = reward[i];
        memcpy(x + K + 1, cur_q_states + (size_t)i * S, sizeof(double) * S);
        x[K + 1 + S] = forced ? (double)forced[i] : -1.0;
        x[K + 2 + S] = learn ? 1.0 : 0.0;
    }
    rc = friqd_h2d(b->d_in0, h, 8ul * n * NIN);
    if (!rc) rc = friqd_classic_step(&b->P->hp, b->P->dp, b->st, b->d_in0, b->d_out0, NULL);
    if (!rc) rc = friqd_d2h(h, b->d_out0, 8ul * n * NOUT);
    if (!rc)
        for (long i = 0; i < n; i++) {
            const double *o = h + (size_t)i * NOUT;
            if (action) action[i] = (int)o[0];
            if (rules) rules[i] = (int)o[1];
            if (inserted) inserted[i] = (int)o[2];
            if (actconc) memcpy(actconc + (size_t)i * A, o + 4, sizeof(double) * A);
            if (o[3] != 0.0) { snprintf(g_err, sizeof g_err, "friq_batch_classic_step: agent %ld exceeded the rule capacity %d", i, b->st.cap); rc = 7; }
        }
    free(h);
    if (rc && rc != 7) FAIL(rc, "friq_batch_classic_step: %s", friqd_error_string(rc));
    return rc;
}

/* five_rule_distance (FIVE.h:95, five_rule_distance.c:63): ant [n][K] -> dists [n][rule_capacity] (first R valid), hit [n] */
int friq_batch_rule_distance(friq_batch *b, const double *ant, double *dists, int *hit)
{
    const long n = b->st.n; const int K = b->P->hp.K;
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    int rc = percall_buffers(b); if (rc) return rc;
    DEV(ensure((void **)&b->d_w, &b->cap_w, 8ul * n * b->st.cap));
    DEV(friqd_h2d(b->d_in0, ant, 8ul * n * K));
    DEV(friqd_rule_distance(&b->P->hp, b->P->dp, b->st, b->d_in0, b->d_w, b->d_iout, NULL));
    DEV(friqd_d2h(hit, b->d_iout, 4ul * n));
    if (dists) {
        if (n == 1) { int R; DEV(friqd_d2h(&R, b->st.R, sizeof(int))); DEV(friqd_d2h(dists, b->d_w, 8ul * R)); }
        else DEV(friqd_d2h(dists, b->d_w, 8ul * n * b->st.cap));
    }
    return 0;
}

/* FIVEVagConcl_FRIRL_BestAct (FIVE.h:86, FIVEVagConcl_FRIRL_BestAct.c:56): dists [n][rule_capacity] -> conc [n] */
int friq_batch_bestact_from_dists(friq_batch *b, const double *dists, double *conc)
{
    const long n = b->st.n;
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    int rc = percall_buffers(b); if (rc) return rc;
    DEV(ensure((void **)&b->d_w, &b->cap_w, 8ul * n * b->st.cap));
    if (n == 1) { int R; DEV(friqd_d2h(&R, b->st.R, sizeof(int))); DEV(friqd_h2d(b->d_w, dists, 8ul * R)); }
    else DEV(friqd_h2d(b->d_w, dists, 8ul * n * b->st.cap));
    DEV(friqd_bestact_dists(&b->P->hp, b->st, b->d_w, b->d_out0, NULL));
    DEV(friqd_d2h(conc, b->d_out0, 8ul * n));
    return 0;
}

/* FIVE_add_rule (FIVE.h:100, five_add_rule.c:47-95) for one agent: append (ant[K] -> q) */
int friq_batch_append_rule(friq_batch *b, long agent, const double *ant, double q)
{
    friq_problem *P = b->P; const int K = P->hp.K;
    if (agent < 0 || agent >= b->st.n) FAIL(5, "friq_batch_append_rule: agent %ld out of range", agent);
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    int R; DEV(friqd_d2h(&R, b->st.R + agent, sizeof(int)));
    if (R >= b->st.cap) FAIL(5, "friq_batch_append_rule: rule capacity %d exhausted", b->st.cap);
    unsigned int word = 0; int grew = 0;
    for (int k = 0; k < K; k++) {
        const int before = P->hp.dict_len[k];
        const int j = dict_find_or_add(P, k, ant[k]);
        if (j < 0) FAIL(5, "friq_batch_append_rule: more than %d distinct antecedent values in dimension %d", FRIQ_DICT_MAX(K), k);
        grew |= (P->hp.dict_len[k] != before);
        word |= (unsigned int)j << (FRIQ_BITS_OF(K) * k);
    }
    if (grew) { refresh_hit_scan(P); int rcp_ = problem_publish(P); if (rcp_) return rcp_; }
    const size_t off = (size_t)(agent / FRIQ_TILE) * b->st.cap * FRIQ_TILE + (agent % FRIQ_TILE) + (size_t)R * FRIQ_TILE;
    DEV(friqd_h2d(b->st.rbw + off, &word, 4)); DEV(friqd_h2d(b->st.rbq + off, &q, 8));
    R++;
    DEV(friqd_h2d(b->st.R + agent, &R, sizeof(int)));
    { const int unknown = -2; DEV(friqd_h2d(b->st.hitq + agent, &unknown, sizeof(int))); }
    return 0;
}

/* From now on every entry point that runs update_rules for this batch (friq_batch_update_sarsa, friq_batch_classic_step, and
 * friq_batch_run with FRIQ_RUN_KEEP_WEIGHTS) maintains FIVERB.weights the way the reference does (what merge_rb reads). */
int friq_batch_keep_weights(friq_batch *b)
{
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    return ensure_wref(b);
}

/* merge_rb (frirl_agent.c:58-117): the sender's rule list folded into one agent's rule base (agent < 0: into every agent's).
 * The receivers' FIVERB.weights are the ones kept by runs with FRIQ_RUN_KEEP_WEIGHTS (zero before any such run). */
int friq_batch_merge_rb(friq_batch *b, long agent, int Rs, const double *rant, const double *rconc)
{
    friq_problem *P = b->P; const int K = P->hp.K;
    if (agent >= b->st.n) FAIL(5, "friq_batch_merge_rb: agent %ld out of range", agent);
    if (Rs < 0) FAIL(5, "friq_batch_merge_rb: negative rule count");
    if (Rs == 0) return 0;
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    unsigned int *w = (unsigned int *)malloc(sizeof(unsigned int) * Rs);
    int grew = 0;
    for (int i = 0; i < Rs; i++) {
        unsigned int word = 0;
        for (int k = 0; k < K; k++) {
            const int before = P->hp.dict_len[k];
            const int j = dict_find_or_add(P, k, rant[(size_t)i * K + k]);
            if (j < 0) { free(w); FAIL(5, "friq_batch_merge_rb: more than %d distinct antecedent values in dimension %d", FRIQ_DICT_MAX(K), k); }
            grew |= (P->hp.dict_len[k] != before);
            word |= (unsigned int)j << (FRIQ_BITS_OF(K) * k);
        }
        w[i] = word;
    }
    int rc = 0;
    if (grew) { refresh_hit_scan(P); rc = problem_publish(P); }
    if (!rc) rc = ensure_wref(b);
    unsigned int *dw = NULL; double *dq = NULL;
    if (!rc) rc = friqd_malloc((void **)&dw, sizeof(unsigned int) * Rs);
    if (!rc) rc = friqd_malloc((void **)&dq, sizeof(double) * Rs);
    if (!rc) rc = friqd_h2d(dw, w, sizeof(unsigned int) * Rs);
    if (!rc) rc = friqd_h2d(dq, rconc, sizeof(double) * Rs);
    if (!rc) rc = friqd_merge_rb(&P->hp, P->dp, b->st, agent, dw, dq, Rs, NULL);
    if (!rc) rc = friqd_device_sync();
    if (dw) friqd_free(dw);
    if (dq) friqd_free(dq);
    free(w);
    if (rc) FAIL(rc, "friq_batch_merge_rb: %s", friqd_error_string(rc));
    return 0;
}

/* Knowledge-sharing run (the reference's experimental frirl_omp_run, frirl_agent.c:294-385) for a batch of TEAMS of
 * `team` consecutive agents, all teams in parallel on the device:
 *   member 0 (the leader) starts its episodes at values_def, member id >= 1 at the state part of corner rule (id - 1) * gap,
 *   gap = corner rules / (team - 2) (gen_def_states, :120-138; past the last corner rule the reference reads zeros);
 *   a round = every member learns a chunk of FRIRL_AGENT_EPCHUNK - 1 = 9 episodes (fewer when its "rule base complete" test fires,
 *   frirl_sequential_run.c:51-65,152-161), then the leader's rules are merged into every other member and every other member's
 *   rules into the leader (merge_rb, members that "pended" in the chunk are skipped, :338-363);
 *   a team is finished when its leader finds its rule base complete (or runs out of episodes).
 * log (optional): rows (agent, steps, return, rules) of every finished episode, in the order the reference prints them per agent;
 * the caller sizes it (max_log rows; rows beyond it are dropped), *n_log returns the number of finished episodes (= the rows a large
 * enough log receives).  No per-episode device-to-host traffic when both are NULL.  Returns 0; *rounds_out = rounds run. */
int friq_team_run(friq_batch *b, int team, int max_rounds, double *log, long max_log, long *n_log, int *rounds_out)
{
    friq_problem *P = b->P; const friq_devprob *hp = &P->hp;
    const long n = b->st.n; const int S = hp->S;
    if (team < 3) FAIL(6, "friq_team_run: a team needs at least 3 agents (the reference divides by team - 2, frirl_agent.c:132)");
    if (n < team) FAIL(6, "friq_team_run: the batch holds no complete team of %d agents", team);
    if (hp->env == FRIQ_ENV_HOST_CALLBACKS) FAIL(6, "friq_team_run: the problem has no built-in environment");
    DEV(friqd_set_device(b->device));
    IDENTITY(b);
    {   /* start states */
        double *s0 = (double *)calloc((size_t)n * S, sizeof(double));
        const int gap = hp->ncorner / (team - 2);
        for (long i = 0; i < n; i++) {
            const int id = (int)(i % team);
            for (int k = 0; k < S; k++) {
                double v = hp->values_def[k];
                if (id >= 1) {
                    const int idx = (id - 1) * gap;
                    v = idx < hp->ncorner ? hp->dict_val[k][FRIQ_FIELD(hp->corner[idx], k, hp->K)] : 0.0;
                }
                s0[(size_t)i * S + k] = v;
            }
        }
        int rc = friq_batch_set_initial_state(b, s0);
        free(s0);
        if (rc) return rc;
    }
    { int rc = ensure_wref(b); if (rc) return rc; }
    int *d_count = NULL; DEV(friqd_malloc((void **)&d_count, sizeof(int)));
    int *ep_before = NULL, *ep_after = NULL, *h_steps = NULL, *h_rules = NULL; double *h_ret = NULL;
    if (log || n_log) {
        ep_before = (int *)malloc(sizeof(int) * n); ep_after = (int *)malloc(sizeof(int) * n);
        h_steps = (int *)malloc(sizeof(int) * n); h_rules = (int *)malloc(sizeof(int) * n); h_ret = (double *)malloc(sizeof(double) * n);
    }
    /* per-agent logs are kept in per-agent order: collect (round, episode) rows, then emit agent by agent */
    long rows = 0, cap_rows = 0; double *tmp = NULL;
    const int mode = FRIQ_RUN_LEARN | FRIQ_RUN_TRACK_CONVERGENCE | FRIQ_RUN_STOP_AT_EPISODE_END | FRIQ_RUN_STOP_ON_CONVERGE | FRIQ_RUN_KEEP_WEIGHTS;
    int rounds = 0, rc = 0;
    for (; rounds < max_rounds && !rc; rounds++) {
        rc = friqd_team_ctl(b->st, team, 0, hp->max_episodes, d_count, NULL);
        for (int e = 1; e < 10 && !rc; e++) {                                /* epchunk = 1 .. FRIRL_AGENT_EPCHUNK - 1 */
            if (ep_before) rc = friqd_d2h(ep_before, b->st.ep_num, 4 * n);
            if (!rc) rc = friq_batch_run(b, hp->max_steps, mode, NULL);
            if (!rc && ep_before) {
                rc = friqd_device_sync();
                if (!rc) rc = friqd_d2h(ep_after, b->st.ep_num, 4 * n);
                if (!rc) rc = friqd_d2h(h_steps, b->st.last_steps, 4 * n);
                if (!rc) rc = friqd_d2h(h_rules, b->st.last_rules, 4 * n);
                if (!rc) rc = friqd_d2h(h_ret, b->st.last_return, 8 * n);
                for (long i = 0; i < n && !rc; i++) if (ep_after[i] != ep_before[i]) {
                    if (rows == cap_rows) { cap_rows = cap_rows ? cap_rows * 2 : 4096; tmp = (double *)realloc(tmp, sizeof(double) * 4 * cap_rows); }
                    tmp[4 * rows] = (double)i; tmp[4 * rows + 1] = h_steps[i]; tmp[4 * rows + 2] = h_ret[i]; tmp[4 * rows + 3] = h_rules[i];
                    rows++;
                }
            }
        }
        int running = 0;
        if (!rc) rc = friqd_team_ctl(b->st, team, 1, hp->max_episodes, d_count, NULL);
        if (!rc) rc = friqd_device_sync();
        if (!rc) rc = friqd_d2h(&running, d_count, sizeof(int));
        if (rc || running == 0) { rounds++; break; }
        rc = friqd_team_merge(hp, P->dp, b->st, team, 0, NULL);
        if (!rc) rc = friqd_team_merge(hp, P->dp, b->st, team, 1, NULL);
        if (!rc) rc = friqd_device_sync();
    }
    if (!rc && tmp && log) {                                                /* stable counting sort by agent: O(rows + n) */
        long *start = (long *)calloc((size_t)n + 1, sizeof(long));
        for (long r = 0; r < rows; r++) start[(long)tmp[4 * r] + 1]++;
        for (long a = 0; a < n; a++) start[a + 1] += start[a];
        for (long r = 0; r < rows; r++) {
            const long o = start[(long)tmp[4 * r]]++;
            if (o < max_log) memcpy(log + 4 * o, tmp + 4 * r, sizeof(double) * 4);
        }
        free(start);
    }
    if (n_log) *n_log = rows;
    if (rounds_out) *rounds_out = rounds;
    free(tmp); free(ep_before); free(ep_after); free(h_steps); free(h_rules); free(h_ret);
    friqd_free(d_count);
    if (rc) FAIL(rc, "friq_team_run: %s", friqd_error_string(rc));
    return 0;
}

/* five_remove_rule (FIVE.h:97, five_remove_rule.c:29-85) for one agent: later rules move down one slot */
int friq_batch_remove_rule(friq_batch *b, long agent, int rule)
{
    const int K = b->P->hp.K;
    int R = 0;
    int rc = friq_batch_get_rulebase(b, agent, &R, NULL, NULL); if (rc) return rc;
    if (rule < 0 || rule >= R) FAIL(5, "friq_batch_remove_rule: rule %d out of range (rules: %d)", rule, R);
    double *rant = (double *)malloc(sizeof(double) * (size_t)R * K), *rconc = (double *)malloc(sizeof(double) * R);
    rc = friq_batch_get_rulebase(b, agent, &R, rant, rconc);
    if (!rc) {
        memmove(rant + (size_t)rule * K, rant + (size_t)(rule + 1) * K, sizeof(double) * (size_t)(R - 1 - rule) * K);
        memmove(rconc + rule, rconc + rule + 1, sizeof(double) * (R - 1 - rule));
        rc = friq_batch_set_rulebase(b, agent, R - 1, rant, rconc);
    }
    free(rant); free(rconc);
    return rc;
}

/* ------------------------------------------------------------------- rule-base reduction */
/* Reduction strategies 1 (smallest |Q| first) and 2 (largest |Q| first) of frirl_sequential_run.c:170-350 for a problem with a
 * built-in environment: take one rule out, replay one greedy episode without updates (reduction_state = 1), keep the removal
 * when the episode still takes steps_ref steps, its reward is good and within reward_tolerance of the last accepted reward,
 * else put the rule back and mark it important.  The reference tests one candidate per episode; here up to `width`
 * candidates are tested SPECULATIVELY in one launch - agent i replays the rule base with the first i + 1 candidates of the
 * chain removed, assuming the earlier ones were accepted - and the results are then consumed in order up to the first
 * rejection (whose successors are discarded and re-speculated), which yields exactly the sequential result.  With verbose
 * the reference's own log lines are printed.  Deterministic episodes only (no_random = 1 or epsilon = 0). */
static int pick_candidate(const double *cand, int R, int strategy, double *mv_out)
{
    double mv = fabs(cand[0]); int pick = 0;                                  /* frirl_sequential_run.c:262-275 / 305-318 */
    for (int j = 1; j < R; j++) {
        const double a = fabs(cand[j]);
        const int better = (strategy == 1) ? (mv > a) : (mv < a);
        if (better || (mv != mv && cand[j] == cand[j])) { mv = a; pick = j; }
    }
    *mv_out = mv;
    return pick;
}

int friq_reduce_rulebase(friq_problem *P, int device, int strategy, double reward_tolerance, int width,
                         int *R_io, double *rant, double *rconc, int verbose, int *episodes_run)
{
    const friq_devprob *hp = &P->hp; const int K = hp->K, R0 = *R_io;
    if (hp->env == FRIQ_ENV_HOST_CALLBACKS) FAIL(6, "friq_reduce_rulebase: the problem has no built-in environment");
    if (!(hp->no_random == 1 || P->epsilon == 0.0)) FAIL(6, "friq_reduce_rulebase: needs deterministic episodes (no_random = 1 or epsilon = 0)");
    if (strategy != 1 && strategy != 2) FAIL(6, "friq_reduce_rulebase: strategy must be 1 (min |Q| first) or 2 (max |Q| first)");
    if (R0 < 1 || width < 1) FAIL(6, "friq_reduce_rulebase: empty rule base or width < 1");
    friq_batch *b;
    int rc = friq_batch_create(P, device, width, R0 > hp->ncorner ? R0 : hp->ncorner, 0, 0, &b); if (rc) return rc;
    double *ret = (double *)malloc(sizeof(double) * width); int *stp = (int *)malloc(sizeof(int) * width);
    double *cand = (double *)malloc(sizeof(double) * R0), *tant = (double *)malloc(sizeof(double) * (size_t)R0 * K), *tq = (double *)malloc(sizeof(double) * R0);
    double *tcand = (double *)malloc(sizeof(double) * R0), *removed = (double *)malloc(sizeof(double) * (size_t)width * (K + 1));
    int *picks = (int *)malloc(sizeof(int) * width);
    int R = R0, ep = 0, done = 0, total_eps = 0;
#define REDUCE_FAIL(r) do { rc = (r); goto out; } while (0)
    /* the reference first replays the full rule base to learn steps_ref (:193-194), then episode 1 replays it again (:198) */
    if ((rc = friq_batch_set_rulebase(b, -1, R, rant, rconc))) REDUCE_FAIL(rc);
    if ((rc = friq_batch_run(b, hp->max_steps, FRIQ_RUN_STOP_AT_EPISODE_END, NULL))) REDUCE_FAIL(rc);
    if ((rc = friq_batch_gather(b, ret, stp, NULL, NULL, NULL, NULL, NULL))) REDUCE_FAIL(rc);
    total_eps += 2;
    const int steps_ref = stp[0];
    double prev_reward = ret[0];
    const int iterations = R0 + 1;
    memcpy(cand, rconc, sizeof(double) * R);
    ep = 1;
    printf("Reduction Episode: %d\tSteps: %d\tReward: %f\tEpsilon: %f\tRules: %d\n", ep, stp[0], ret[0], P->epsilon, R);
    while (!done) {
        /* ---- speculate: chain of picks from the current state, each assuming the previous ones were accepted */
        int m = 0, tR = R;
        memcpy(tant, rant, sizeof(double) * (size_t)R * K); memcpy(tq, rconc, sizeof(double) * R); memcpy(tcand, cand, sizeof(double) * R);
        if ((rc = friq_batch_reset(b))) REDUCE_FAIL(rc);
        while (m < width && ep + m < iterations) {
            double mv; const int pick = pick_candidate(tcand, tR, strategy, &mv);
            if (mv != mv) break;                                              /* every remaining rule is marked important */
            picks[m] = pick;
            memcpy(removed + (size_t)m * (K + 1), tant + (size_t)pick * K, sizeof(double) * K); removed[(size_t)m * (K + 1) + K] = tcand[pick];
            memmove(tant + (size_t)pick * K, tant + (size_t)(pick + 1) * K, sizeof(double) * (size_t)(tR - 1 - pick) * K);
            memmove(tq + pick, tq + pick + 1, sizeof(double) * (tR - 1 - pick));
            memmove(tcand + pick, tcand + pick + 1, sizeof(double) * (tR - 1 - pick));
            tR--;
            if ((rc = friq_batch_set_rulebase(b, m, tR, tant, tq))) REDUCE_FAIL(rc);
            m++;
        }
        if (m == 0) {                                                         /* nothing left to try (:276-281) or out of iterations */
            double mv; pick_candidate(cand, R, strategy, &mv);
            if (mv != mv && verbose) printf("Smallest rulebase found. Exiting.\n");
            break;
        }
        if ((rc = friq_batch_run(b, hp->max_steps, FRIQ_RUN_STOP_AT_EPISODE_END, NULL))) REDUCE_FAIL(rc);
        if ((rc = friq_batch_gather(b, ret, stp, NULL, NULL, NULL, NULL, NULL))) REDUCE_FAIL(rc);
        /* ---- consume the results in order */
        for (int i = 0; i < m; i++) {
            const int pick = picks[i];
            ep++; total_eps++;
            printf("Reduction Episode: %d\tSteps: %d\tReward: %f\tEpsilon: %f\tRules: %d\n", ep, stp[i], ret[i], P->epsilon, R - 1);
            const double diff = prev_reward - ret[i];
            if (ret[i] > hp->reward_good_above && stp[i] == steps_ref && fabs(diff) <= reward_tolerance) {       /* :206-219 */
                if (verbose) {
                    printf("Reduction Episode: %d\tEliminated rule: no: %d. - ", ep, pick + 1);
                    for (int j = 0; j <= K; j++) printf(" %f", removed[(size_t)i * (K + 1) + j]);
                    printf(" \tReward diff was: %f\n", diff);
                }
                prev_reward = ret[i];
                memmove(rant + (size_t)pick * K, rant + (size_t)(pick + 1) * K, sizeof(double) * (size_t)(R - 1 - pick) * K);
                memmove(rconc + pick, rconc + pick + 1, sizeof(double) * (R - 1 - pick));
                memmove(cand + pick, cand + pick + 1, sizeof(double) * (R - 1 - pick));
                R--;
            } else {                                                          /* :220-237: the rule stays and is marked important */
                cand[pick] = NAN;
                if (verbose) {
                    printf("Reduction Episode: %d\tRule stays: no: %d. - ", ep, pick + 1);
                    for (int j = 0; j < K; j++) printf(" %f", rant[(size_t)pick * K + j]);
                    printf(" %f\n", rconc[pick]);
                }
                break;                                                        /* later speculations assumed this removal */
            }
        }
        if (ep >= iterations) {                                               /* the last iteration still picks (:262-300) ... */
            double mv; const int pick = pick_candidate(cand, R, strategy, &mv);
            if (mv != mv) { if (verbose) printf("Smallest rulebase found. Exiting.\n"); }
            else {                                                            /* ... and its removal is never replayed */
                memmove(rant + (size_t)pick * K, rant + (size_t)(pick + 1) * K, sizeof(double) * (size_t)(R - 1 - pick) * K);
                memmove(rconc + pick, rconc + pick + 1, sizeof(double) * (R - 1 - pick));
                R--;
            }
            done = 1;
        }
    }
    *R_io = R;
    if (episodes_run) *episodes_run = total_eps;
    rc = 0;
out:
    friq_batch_destroy(b);
    free(ret); free(stp); free(cand); free(tant); free(tq); free(tcand); free(removed); free(picks);
    return rc;
#undef REDUCE_FAIL
}

/* ----------------------------------------------------------------------- rule-base files */
/* frirl_save_rb_to_text_file (frirl_utils.c:100-146): "%.18f " per antecedent, "%.18f \n" for Q */
int friq_save_rb_to_text_file(const char *file_name, int K, int R, const double *rant, const double *rconc)
{
    FILE *f = fopen(file_name, "w");
    if (!f) { perror("friq_save_rb_to_text_file: open()"); FAIL(-1, "cannot open %s", file_name); }
    for (int i = 0; i < R; i++) {
        for (int k = 0; k < K; k++) fprintf(f, "%.18f ", rant[(size_t)i * K + k]);
        fprintf(f, "%.18f \n", rconc[i]);
    }
    const int werr = ferror(f);
    if (fclose(f) != 0 || werr) FAIL(-1, "%s: write failed", file_name);
    return 0;
}
/* frirl_save_rb_to_bin_file (frirl_utils.c:151-203): int32 R, then per rule K doubles + Q, host endian */
int friq_save_rb_to_bin_file(const char *file_name, int K, int R, const double *rant, const double *rconc)
{
    FILE *f = fopen(file_name, "wb");
    if (!f) { perror("friq_save_rb_to_bin_file: open()"); FAIL(-1, "cannot open %s", file_name); }
    int32_t n = R;
    fwrite(&n, sizeof n, 1, f);
    for (int i = 0; i < R; i++) { fwrite(rant + (size_t)i * K, sizeof(double), K, f); fwrite(rconc + i, sizeof(double), 1, f); }
    const int werr = ferror(f);
    if (fclose(f) != 0 || werr) FAIL(-1, "%s: write failed", file_name);
    return 0;
}
/* frirl_load_rb_from_bin_file (frirl_utils.c:228-281) */
int friq_load_rb_from_bin_file(const char *file_name, int K, int max_rules, int *R, double *rant, double *rconc)
{
    FILE *f = fopen(file_name, "rb");
    if (!f) { perror("friq_load_rb_from_bin_file: open()"); FAIL(-1, "cannot open %s", file_name); }
    int32_t n = 0;
    if (fread(&n, sizeof n, 1, f) != 1 || n < 0 || n > max_rules) { fclose(f); FAIL(-1, "%s: bad rule count %d (max %d)", file_name, n, max_rules); }
    for (int i = 0; i < n; i++)
        if (fread(rant + (size_t)i * K, sizeof(double), K, f) != (size_t)K || fread(rconc + i, sizeof(double), 1, f) != 1) { fclose(f); FAIL(-1, "%s: truncated", file_name); }
    fclose(f);
    *R = n;
    return 0;
}

/* ------------------------------------------------------------------- batched rule-base I/O */
/* rule bases of the agents [first, first + count) -> host: R [count], rant [count][cap][K], rconc [count][cap] (bulk copies
 * of whole tiles, unpacked on the host) */
static int fetch_rulebases(friq_batch *b, long first, long count, int *R, double *rant, double *rconc)
{
    const friq_devprob *hp = &b->P->hp; const int K = hp->K, cap = b->st.cap;
    if (first < 0 || count < 0 || first + count > b->st.n) FAIL(5, "rule-base range [%ld, %ld) out of range", first, first + count);
    DEV(friqd_set_device(b->device));
    DEV(friqd_device_sync());
    IDENTITY(b);
    const long t0 = first / FRIQ_TILE, t1 = (first + count + FRIQ_TILE - 1) / FRIQ_TILE;
    const size_t tile_slots = (size_t)cap * FRIQ_TILE;
    unsigned int *w = (unsigned int *)malloc(tile_slots * 4); double *q = (double *)malloc(tile_slots * 8);
    int rc = friqd_d2h(R, b->st.R + first, 4ul * count);
    for (long t = t0; t < t1 && !rc; t++) {
        rc = friqd_d2h(w, b->st.rbw + (size_t)t * tile_slots, tile_slots * 4);
        if (!rc) rc = friqd_d2h(q, b->st.rbq + (size_t)t * tile_slots, tile_slots * 8);
        for (int l = 0; l < FRIQ_TILE && !rc; l++) {
            const long ag = t * FRIQ_TILE + l;
            if (ag < first || ag >= first + count) continue;
            const long o = ag - first;
            for (int i = 0; i < R[o]; i++) {
                const unsigned int word = w[(size_t)i * FRIQ_TILE + l];
                for (int k = 0; k < K; k++) rant[((size_t)o * cap + i) * K + k] = hp->dict_val[k][FRIQ_FIELD(word, k, K)];
                rconc[(size_t)o * cap + i] = q[(size_t)i * FRIQ_TILE + l];
            }
        }
    }
    free(w); free(q);
    if (rc) FAIL(rc, "fetch_rulebases: %s", friqd_error_string(rc));
    return 0;
}

/* one rule-base file per agent in the reference's formats (frirl_utils.c:100-203): `pattern` is a printf format with one %ld
 * (the agent index), e.g. "run7/agent%06ld.frirlrb.bin"; text != 0 writes the "%.18f " text format */
int friq_batch_save_rb_files(friq_batch *b, long first, long count, const char *pattern, int text)
{
    const int K = b->P->hp.K, cap = b->st.cap;
    const long chunk = 1024;
    int *R = (int *)malloc(sizeof(int) * chunk);
    double *rant = (double *)malloc(sizeof(double) * (size_t)chunk * cap * K), *rconc = (double *)malloc(sizeof(double) * (size_t)chunk * cap);
    int rc = 0;
    for (long lo = first; lo < first + count && !rc; lo += chunk) {
        const long c = (first + count - lo < chunk) ? first + count - lo : chunk;
        rc = fetch_rulebases(b, lo, c, R, rant, rconc);
        for (long i = 0; i < c && !rc; i++) {
            char name[1024];
            snprintf(name, sizeof name, pattern, lo + i);
            rc = text ? friq_save_rb_to_text_file(name, K, R[i], rant + (size_t)i * cap * K, rconc + (size_t)i * cap)
                      : friq_save_rb_to_bin_file(name, K, R[i], rant + (size_t)i * cap * K, rconc + (size_t)i * cap);
        }
    }
    free(R); free(rant); free(rconc);
    return rc;
}

/* warm start (`-f file`, frirl_app_helpers.c:53-62): the rule base of a reference .frirlrb.bin file for one agent, or for every
 * agent when agent < 0; fus_is_rule_inserted is cleared like frirl_load_rb_from_bin_file does (frirl_utils.c:263) */
int friq_batch_load_rb_file(friq_batch *b, long agent, const char *file_name)
{
    const int K = b->P->hp.K, cap = b->st.cap;
    double *rant = (double *)malloc(sizeof(double) * (size_t)cap * K), *rconc = (double *)malloc(sizeof(double) * cap);
    int R = 0;
    int rc = friq_load_rb_from_bin_file(file_name, K, cap, &R, rant, rconc);
    if (!rc) rc = friq_batch_set_rulebase(b, agent, R, rant, rconc);
    free(rant); free(rconc);
    if (rc) return rc;
    int *f = (int *)malloc(sizeof(int) * b->st.n);
    rc = friq_batch_get_inserted(b, f);
    if (!rc) { if (agent < 0) memset(f, 0, sizeof(int) * b->st.n); else f[agent] = 0; rc = friq_batch_set_inserted(b, f); }
    free(f);
    return rc;
}

/* Checkpoint of a whole batch in ONE file: header {magic "FRIQCKP2", int64 n_agents, int32 K, int32 flags (bit 0: a weights
 * section follows), int32 rule_capacity at save time (= doubles per agent in the weights section), int32 0}, then per agent
 * {int32 R, int32 fus_is_rule_inserted, R x (K antecedents + consequent) f64} - the per-agent record is the reference's binary
 * rule-base format (frirl_utils.c:175-203) with the sticky flag in front - then, with flag bit 0, every agent's FIVERB.weights
 * (rule_capacity f64 each).  A batch with another capacity re-strides the weights on load (every R must fit); the file must
 * end after the last section.  Like the reference, no episode state is saved. */
int friq_batch_save_checkpoint(friq_batch *b, const char *file_name)
{
    const int K = b->P->hp.K, cap = b->st.cap; const long n = b->st.n, chunk = 1024;
    FILE *f = fopen(file_name, "wb");
    if (!f) { perror("friq_batch_save_checkpoint: open()"); FAIL(-1, "cannot open %s", file_name); }
    const long long n64 = n; const int32_t k32 = K, flags32 = b->st.wref ? 1 : 0;   /* bit 0: a FIVERB.weights section follows the records */
    const int32_t cap32 = cap, zero32 = 0;            /* stride of the weights section: rule_capacity doubles per agent */
    fwrite("FRIQCKP2", 1, 8, f); fwrite(&n64, 8, 1, f); fwrite(&k32, 4, 1, f); fwrite(&flags32, 4, 1, f); fwrite(&cap32, 4, 1, f); fwrite(&zero32, 4, 1, f);
    int *R = (int *)malloc(sizeof(int) * chunk), *ins = (int *)malloc(sizeof(int) * n);
    double *rant = (double *)malloc(sizeof(double) * (size_t)chunk * cap * K), *rconc = (double *)malloc(sizeof(double) * (size_t)chunk * cap);
    int rc = friq_batch_get_inserted(b, ins);
    for (long lo = 0; lo < n && !rc; lo += chunk) {
        const long c = (n - lo < chunk) ? n - lo : chunk;
        rc = fetch_rulebases(b, lo, c, R, rant, rconc);
        for (long i = 0; i < c && !rc; i++) {
            const int32_t r32 = R[i], i32 = ins[lo + i];
            fwrite(&r32, 4, 1, f); fwrite(&i32, 4, 1, f);
            for (int r = 0; r < R[i]; r++) { fwrite(rant + ((size_t)i * cap + r) * K, 8, K, f); fwrite(rconc + (size_t)i * cap + r, 8, 1, f); }
        }
    }
    if (!rc && b->st.wref) {          /* knowledge sharing: the persistent FIVERB.weights of every agent (cap doubles each), tile by tile */
        const size_t tile_slots = (size_t)cap * FRIQ_TILE;
        double *t = (double *)malloc(tile_slots * 8), *o = (double *)malloc(sizeof(double) * cap);
        for (long tl = 0; tl * FRIQ_TILE < n && !rc; tl++) {
            rc = friqd_d2h(t, b->st.wref + (size_t)tl * tile_slots, tile_slots * 8);
            for (int l = 0; l < FRIQ_TILE && tl * FRIQ_TILE + l < n && !rc; l++) {
                for (int i = 0; i < cap; i++) o[i] = t[(size_t)i * FRIQ_TILE + l];
                fwrite(o, 8, cap, f);
            }
        }
        free(t); free(o);
    }
    free(R); free(ins); free(rant); free(rconc);
    const int werr = ferror(f);
    if ((fclose(f) != 0 || werr) && !rc) FAIL(-1, "%s: write failed", file_name);
    return rc;
}

int friq_batch_load_checkpoint(friq_batch *b, const char *file_name)
{
    const int K = b->P->hp.K, cap = b->st.cap; const long n = b->st.n;
    FILE *f = fopen(file_name, "rb");
    if (!f) { perror("friq_batch_load_checkpoint: open()"); FAIL(-1, "cannot open %s", file_name); }
    char magic[8]; long long n64 = 0; int32_t k32 = 0, res = 0, wstride = 0, pad32 = 0;
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "FRIQCKP2", 8) || fread(&n64, 8, 1, f) != 1 || fread(&k32, 4, 1, f) != 1 || fread(&res, 4, 1, f) != 1 ||
        fread(&wstride, 4, 1, f) != 1 || fread(&pad32, 4, 1, f) != 1 || wstride < 0) {
        fclose(f); FAIL(-1, "%s: not a FRIQCKP2 checkpoint", file_name);
    }
    if (n64 != n || k32 != K) { fclose(f); FAIL(-1, "%s: checkpoint of %lld agents / K = %d, batch has %ld / %d", file_name, n64, k32, n, K); }
    double *rant = (double *)malloc(sizeof(double) * (size_t)cap * K), *rconc = (double *)malloc(sizeof(double) * cap);
    int *ins = (int *)malloc(sizeof(int) * n);
    int rc = 0;
    for (long a = 0; a < n && !rc; a++) {
        int32_t r32 = 0, i32 = 0;
        if (fread(&r32, 4, 1, f) != 1 || fread(&i32, 4, 1, f) != 1 || r32 < 0 || r32 > cap) { snprintf(g_err, sizeof g_err, "%s: bad record of agent %ld", file_name, a); rc = -1; break; }
        for (int r = 0; r < r32; r++)
            if (fread(rant + (size_t)r * K, 8, K, f) != (size_t)K || fread(rconc + r, 8, 1, f) != 1) { snprintf(g_err, sizeof g_err, "%s: truncated", file_name); rc = -1; break; }
        if (!rc) rc = friq_batch_set_rulebase(b, a, r32, rant, rconc);
        ins[a] = i32;
    }
    if (!rc) rc = friq_batch_set_inserted(b, ins);
    if (!rc && (res & 1)) {           /* FIVERB.weights section: `wstride` doubles per agent as saved, re-strided to this batch's capacity
                                       * (every agent's R <= cap was checked above, so min(wstride, cap) covers all live weights) */
        rc = friq_batch_keep_weights(b);
        const size_t tile_slots = (size_t)cap * FRIQ_TILE;
        const int ncopy = wstride < cap ? wstride : cap;
        double *t = (double *)malloc(tile_slots * 8), *o = (double *)malloc(sizeof(double) * (wstride ? wstride : 1));
        for (long tl = 0; tl * FRIQ_TILE < n && !rc; tl++) {
            memset(t, 0, tile_slots * 8);
            for (int l = 0; l < FRIQ_TILE && tl * FRIQ_TILE + l < n && !rc; l++) {
                if (fread(o, 8, wstride, f) != (size_t)wstride) { snprintf(g_err, sizeof g_err, "%s: truncated weights section", file_name); rc = -1; break; }
                for (int i = 0; i < ncopy; i++) t[(size_t)i * FRIQ_TILE + l] = o[i];
            }
            if (!rc) rc = friqd_h2d(b->st.wref + (size_t)tl * tile_slots, t, tile_slots * 8);
        }
        free(t); free(o);
    }
    if (!rc && fgetc(f) != EOF) { snprintf(g_err, sizeof g_err, "%s: trailing bytes after the last section (corrupt checkpoint)", file_name); rc = -1; }
    fclose(f); free(rant); free(rconc); free(ins);
    return rc;
}

/* --------------------------------------------------------------------- device self-tests */
int friq_device_trig(const double *x, double *s, double *c, long n)
{
    double *dx, *ds, *dc;
    DEV(friqd_malloc((void **)&dx, 8ul * n)); DEV(friqd_malloc((void **)&ds, 8ul * n)); DEV(friqd_malloc((void **)&dc, 8ul * n));
    DEV(friqd_h2d(dx, x, 8ul * n)); DEV(friqd_trig(dx, ds, dc, n, NULL));
    DEV(friqd_d2h(s, ds, 8ul * n)); DEV(friqd_d2h(c, dc, 8ul * n));
    friqd_free(dx); friqd_free(ds); friqd_free(dc);
    return 0;
}
int friq_device_pow_x87(const double *x, double *out, long n, int p, int force_exact)
{
    double *dx, *dout;
    DEV(friqd_malloc((void **)&dx, 8ul * n)); DEV(friqd_malloc((void **)&dout, 8ul * n));
    DEV(friqd_h2d(dx, x, 8ul * n)); DEV(friqd_pow_x87(dx, dout, n, p, force_exact, NULL)); DEV(friqd_d2h(out, dout, 8ul * n));
    friqd_free(dx); friqd_free(dout);
    return 0;
}
int friq_device_fp64_peak(double *g) { DEV(friqd_fp64_peak(g, 1 << 16)); return 0; }
