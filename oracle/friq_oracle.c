/* oracle/friq_oracle.c -- TEST INFRASTRUCTURE ONLY (see friq_oracle.h).
 *
 * Plain scalar C restatement of the reference's FRIQ-learning hot path under its
 * default build flags (FIVE_FIXRES, FIVE_NONAN, FIVE_NOINF, FRIRL_FAST,
 * BUILD_CHECK_STATES, FAST_POW; CMakeLists.txt:9-24).  Every '*' and '+' is a
 * separately rounded FP64 operation (compile with -ffp-contract=off), rule loops
 * run i = 0..R-1 in order, and pow uses the x87 80-bit product chain exactly as
 * the reference does (src/inl/fast_pow.inl:17-27) - valid on any x86-64 host.
 *
 * Citations are to files under the reference tree (/root/reference/).
 */
#include "friq_oracle.h"
#include <math.h>
#include <quadmath.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ helpers */

/* src/inl/fast_pow.inl:17-27 : long double product chain, rounded to double once */
double fo_pow_x87(double b, int p) {
    long double acc = b;
    for (int i = 1; i < p; i++) acc *= b;
    return (double)acc;
}

/* correctly rounded trig (binary128 -> double); the goldens need it */
double fo_sin(const fo_problem *P, double x) { return P->trig == FO_TRIG_CR ? (double)sinq((__float128)x) : sin(x); }
double fo_cos(const fo_problem *P, double x) { return P->trig == FO_TRIG_CR ? (double)cosq((__float128)x) : cos(x); }

/* Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3").
 * The reference's libc rand() seeded from gettimeofday (frirl_init.c:335-336) is
 * unreproducible; the batch engine defines this counter-based stream instead. */
void fo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* src/inl/min.inl:71-92 (nearest universe index; ties -> lower index) */
int fo_nearest_index(const fo_problem *P, int k, double x, int len) {
    const double *u = P->u + (size_t)k * P->U;
    int low = (int)((x - u[0]) / P->udiv[k]);
    if (low < 0) return 0;
    if (low >= len) return len - 1;
    if (low + 1 >= P->U) return low;           /* the reference reads one past the row here */
    double d1 = fabs(u[low] - x), d2 = fabs(u[low + 1] - x);
    return d1 <= d2 ? low : low + 1;
}
static double obs_ve(const fo_problem *P, int k, double x) {
    return P->ve[(size_t)k * P->U + fo_nearest_index(P, k, x, P->U)];
}

/* ------------------------------------------------------- problem construction */

/* src/frirl/frirl_app_helpers.c:32-44 */
static void gen_fixres(double *arr, int len, double div) {
    double from = -((len - 1) * div) / 2;
    int i;
    for (i = 0; i < len / 2 + 1; i++) arr[i] = from + div * i;
    for (; i < len; i++) arr[i] = arr[len - 1 - i] * -1;
}

/* src/five/FIVEGScFunc.c:76-255 restricted to what frirl_init_ve uses: mp > 1
 * scaling points of (position, S_left, S_right) and nls = NAN (always the
 * constant / linear branch, :83-91,135-146). */
static void scaling_function(const double *u, int U, const double *psc, int mp, double *scf) {
    for (int j = 0; j < U; j++) {
        double x = u[j];
        if (x < psc[0]) scf[j] = psc[1];
        for (int p = 0; p < mp - 1; p++) {
            double p1 = psc[p * 3], p2 = psc[(p + 1) * 3];
            if (x >= p1 && x < p2) {
                double s1 = psc[p * 3 + 2], s2 = psc[(p + 1) * 3 + 1];
                scf[j] = (s1 == s2) ? s1 : ((s2 - s1) / (p2 - p1)) * (x - p1) + s1;
            }
        }
        if (x >= psc[(mp - 1) * 3])
            scf[j] = (j == U - 1 && x == psc[(mp - 1) * 3]) ? psc[(mp - 1) * 3 + 1] : psc[(mp - 1) * 3 + 2];
    }
}
/* src/five/FIVEGVagEnv.c:81-94 (no INF scaling factor: primitive integral) */
static void vague_environment(const double *u, int U, const double *scf, double *ve) {
    ve[0] = 0;
    for (int j = 0; j < U - 1; j++) ve[j + 1] = ve[j] + (u[j + 1] - u[j]) * (scf[j] + scf[j + 1]) * 0.5;
}

struct dim_spec { int len; const double *vals; double div, steep, def, udiv; };

static const double MC_POS[] = {-1.5, -1.295, -1.09, -0.885, -0.68, -0.475, -0.27, -0.065, 0.14, 0.345};
static const double MC_VEL[] = {-0.07, -0.042, -0.014, 0.014, 0.042, 0.07};
static const double AC_ANG[] = {-1.570796326794897, -0.785398163397448, 0, 0.785398163397448, 1.570796326794897};
static const double CP_TH[]  = {-0.2094, -0.1571, -0.1047, -0.0524, 0.0, 0.0524, 0.1047, 0.1571, 0.2094};
static const double CP_ACT[] = {-1.0, -0.9, -0.8, -0.7, -0.6, -0.5, -0.3999999999999999, -0.29999999999999992,
    -0.19999999999999995, -0.09999999999999998, 0.0, +0.09999999999999998, +0.19999999999999995,
    +0.29999999999999992, +0.3999999999999999, +0.5, +0.6, +0.7, +0.8, +0.9, +1.0};

fo_problem *fo_problem_builtin(int env, int trig) { return fo_problem_builtin_actions(env, trig, 0, NULL, 0.0); }
/* the example's problem with ANOTHER action dimension (A values: act_vals, or generated with step act_div when NULL - like
 * FRIRL_GEN_FIXRES_VALUES); A = 0 keeps the example's own.  Same table builders, so everything downstream is the general code. */
fo_problem *fo_problem_builtin_actions(int env, int trig, int A, const double *act_vals, double act_div) {
    fo_problem *P = calloc(1, sizeof *P);
    struct dim_spec sd[4], ad;
    P->env = env; P->trig = trig;
    P->gamma = 1.0; P->qdiff_pos = 1.0; P->wsig = 0.05; P->skip_rules = 1; P->no_random = 1;
    P->max_episodes = 1000; P->max_steps = 1000;               /* frirl_types_def.h:22-77 */
    if (env == FO_ENV_MOUNTAINCAR) {                           /* mountaincar.c:220-279 */
        P->S = 2; P->U = 41;
        P->alpha = 0.5; P->epsilon = 0.01; P->qdiff_neg = -4.0; P->qdiff_final_tol = 500.0; P->reward_good_above = -5000.0;
        sd[0] = (struct dim_spec){10, MC_POS, 0.205, 1.0, -0.5, 0.1};
        sd[1] = (struct dim_spec){6, MC_VEL, 0.028, 1.0, 0.0, 0.005};
        ad = (struct dim_spec){3, NULL, 1.0, 0, 0, 0.1};
    } else if (env == FO_ENV_ACROBOT) {                        /* acrobot.c:242-324 */
        P->S = 4; P->U = 41;
        P->alpha = 0.5; P->epsilon = 0.001; P->qdiff_neg = -200.0; P->qdiff_final_tol = 50.0; P->reward_good_above = 0.0;
        sd[0] = (struct dim_spec){5, AC_ANG, 0.785398163397448, 1.0, 0.0, 0.1};
        sd[1] = sd[0];
        sd[2] = (struct dim_spec){3, NULL, 0.785398163397448, 1.0, 0.0, 0.05};
        sd[3] = sd[2];
        ad = (struct dim_spec){3, NULL, 1.0, 0, 0, 0.1};
    } else {                                                   /* cartpole.c:252-402 */
        P->S = 4; P->U = 1001;
        P->alpha = 0.3; P->epsilon = 0.001; P->qdiff_neg = -200.0; P->qdiff_final_tol = 250.0; P->reward_good_above = 0.0;
        sd[0] = (struct dim_spec){2, NULL, 2.0, 1.0, 1.0, 0.016};
        sd[1] = (struct dim_spec){3, NULL, 1.0, 1.0, 0.0, 0.032};
        sd[2] = (struct dim_spec){9, CP_TH, 0.0524, 21.485917317405871, 0.0, 0.0031415926535897933};
        sd[3] = (struct dim_spec){2, NULL, 2.0, 1.0, 0.0, 0.016};
        ad = (struct dim_spec){21, CP_ACT, 0.1, 0, 0, 0.008};
    }
    if (A > 0) { ad.len = A; ad.vals = act_vals; ad.div = act_div; }
    int S = P->S, K = S + 1, U = P->U;
    P->K = K; P->A = ad.len; P->p = K;                          /* FIVEInit.c:89-93 with p = 0 */
    P->u = malloc(sizeof(double) * K * U); P->ve = malloc(sizeof(double) * K * U);
    double *scf = malloc(sizeof(double) * U), *psc = malloc(sizeof(double) * 3 * FO_MAXGRID);
    for (int k = 0; k < K; k++) {
        const struct dim_spec *d = k < S ? &sd[k] : &ad;
        P->grid_len[k] = d->len; P->grid_div[k] = d->div; P->values_def[k] = d->def;
        if (d->vals) memcpy(P->grid[k], d->vals, sizeof(double) * d->len);
        else gen_fixres(P->grid[k], d->len, d->div);
        gen_fixres(P->u + (size_t)k * U, U, d->udiv);
        /* frirl_init_ve.c:36-107 */
        if (k < S) {
            for (int i = 0; i < d->len; i++) { psc[3 * i] = P->grid[k][i]; psc[3 * i + 1] = psc[3 * i + 2] = d->steep; }
        } else {
            double divratio = 1.0 / (d->len - 1) * 2.0;
            for (int i = 0; i < d->len; i++) { psc[3 * i] = i * divratio - 1.0; psc[3 * i + 1] = psc[3 * i + 2] = (d->len - 1) / 2; }
        }
        scaling_function(P->u + (size_t)k * U, U, psc, d->len, scf);
        vague_environment(P->u + (size_t)k * U, U, scf, P->ve + (size_t)k * U);
        P->udiv[k] = (P->u[(size_t)k * U + U - 1] - P->u[(size_t)k * U]) / (unsigned)(U - 1);   /* FIVEInit.c:244-248 */
    }
    for (int a = 0; a < P->A; a++) {                            /* frirl_init.c:156-158 (len = uksize = U-1) */
        P->act_val[a] = P->grid[S][a];
        P->act_ve[a] = P->ve[(size_t)S * U + fo_nearest_index(P, S, P->act_val[a], U - 1)];
    }
    free(scf); free(psc);
    return P;
}
void fo_problem_free(fo_problem *P) { if (P) { free(P->u); free(P->ve); free(P); } }

/* ------------------------------------------------------------------ rule base */

/* src/five/five_add_rule.c:47-95 (no capacity check there; flagged here) */
int fo_add_rule(fo_agent *a, const double *ant, double q) {
    const fo_problem *P = a->P;
    if (a->R >= a->cap) { a->overflow = 1; return -1; }
    for (int k = 0; k < P->K; k++) {
        a->rant[(size_t)a->R * P->K + k] = ant[k];
        a->rve[(size_t)a->R * P->K + k] = obs_ve(P, k, ant[k]);
    }
    a->rconc[a->R] = q;
    a->R++;
    return 0;
}
void fo_agent_set_rb(fo_agent *a, int R, const double *rant, const double *rconc) {
    a->R = 0;
    for (int i = 0; i < R; i++) fo_add_rule(a, rant + (size_t)i * a->P->K, rconc[i]);
}

fo_agent *fo_agent_new(const fo_problem *P, int capacity, uint64_t seed) {
    fo_agent *a = calloc(1, sizeof *a);
    a->P = P; a->cap = capacity; a->seed = seed;
    a->rant = calloc((size_t)capacity * P->K, sizeof(double)); a->rve = calloc((size_t)capacity * P->K, sizeof(double));
    a->rconc = calloc(capacity, sizeof(double)); a->w = calloc(capacity, sizeof(double)); a->prev_rconc = calloc(capacity, sizeof(double));
    a->weights = calloc(capacity, sizeof(double));              /* FIVEInit.c:334 mallocs it uncleared: fresh pages, zeros in practice */
    a->alpha = P->alpha; a->gamma = P->gamma; a->epsilon = P->epsilon; a->learning = 1;
    /* corner rules, binary-counter order, q = 0 (frirl_init_rb.c:99-126) */
    int n = 1 << P->K; double ant[FO_MAXK];
    for (int j = 0; j < n; j++) {
        for (int k = 0; k < P->K; k++) {
            double mn = P->grid[k][0], mx = mn;
            for (int i = 0; i < P->grid_len[k]; i++) { if (P->grid[k][i] > mx) mx = P->grid[k][i]; if (P->grid[k][i] < mn) mn = P->grid[k][i]; }
            int divider = n >> (k + 1);
            ant[k] = ((j / divider) % 2 == 0) ? mn : mx;
        }
        fo_add_rule(a, ant, 0.0);
    }
    a->need_reset = 1; a->episode_num = 1;
    a->ep_return = -1; a->ep_steps = -1;                        /* frirl_init.c:144-145 */
    a->prev_reward = -1; a->prev_steps = -1;
    return a;
}
void fo_agent_free(fo_agent *a) { if (a) { free(a->rant); free(a->rve); free(a->rconc); free(a->w); free(a->weights); free(a->prev_rconc); free(a); } }

/* ------------------------------------------------------------- FIVE evaluation */

/* squared rule distance of rule r from the point whose per-dimension VE values
 * are vo[]: left-to-right sum over dims (five_rule_distance.c:160-212, A1/A2) */
static double dist2(const fo_agent *a, int r, const double *vo, int ndims) {
    const double *rv = a->rve + (size_t)r * a->P->K;
    double s = 0.0;
    for (int k = 0; k < ndims; k++) { double d = vo[k] - rv[k]; d = d * d; s = (k == 0) ? d : s + d; }
    return s;
}

/* FIVE_vag_concl (FIVEVagConcl.c:64-351): first exact hit -> its consequent,
 * otherwise sequential Shepard sum; optionally leaves wi in a->w and returns ws */
static int concl_core(fo_agent *a, const double *vo, double *conc, double *ws_out) {
    const fo_problem *P = a->P;
    for (int i = 0; i < a->R; i++)
        if (dist2(a, i, vo, P->K) == 0.0) { *conc = a->rconc[i]; return i; }
    double vagc = 0, ws = 0;
    for (int i = 0; i < a->R; i++) {
        double d = sqrt(dist2(a, i, vo, P->K));
        double wi = 1.0 / fo_pow_x87(fabs(d), P->p);            /* FIVEVagConcl.c:226 */
        a->w[i] = wi;
        vagc += wi * a->rconc[i];                                /* :234 */
        ws += wi;                                                /* :235 */
    }
    if (ws_out) *ws_out = ws;
    *conc = vagc / ws;                                           /* :302 */
    return -1;
}
int fo_vag_concl(fo_agent *a, const double *ant, double *conc) {
    double vo[FO_MAXK];
    for (int k = 0; k < a->P->K; k++) vo[k] = obs_ve(a->P, k, ant[k]);   /* five_rule_distance.c:75 */
    return concl_core(a, vo, conc, NULL);
}
/* FIVE_vag_concl_weight (FIVEVagConclWeight.c:52-188) */
int fo_vag_concl_weight(fo_agent *a, const double *ant, double *weights) {
    double vo[FO_MAXK], conc, ws;
    for (int k = 0; k < a->P->K; k++) vo[k] = obs_ve(a->P, k, ant[k]);
    int hit = concl_core(a, vo, &conc, &ws);
    if (hit >= 0) return hit;
    for (int i = 0; i < a->R; i++) weights[i] = a->w[i] / ws;    /* :153-175 */
    return -1;
}

/* frirl_get_best_action (frirl_get_best_action.c:31-341) + FIVEVagConcl_FRIRL_BestAct */
int fo_get_best_action(fo_agent *a, const double *states, double *actconc) {
    const fo_problem *P = a->P;
    double vo[FO_MAXK];
    for (int k = 0; k < P->S; k++) vo[k] = obs_ve(P, k, states[k]);
    for (int act = 0; act < P->A; act++) {
        int hit = -1;
        double vagc = 0, ws = 0;
        for (int i = 0; i < a->R && hit < 0; i++) {
            double da = P->act_ve[act] - a->rve[(size_t)i * P->K + P->S];
            if (da * da + dist2(a, i, vo, P->S) == 0.0) hit = i;             /* BestAct.c:89-93 */
        }
        if (hit >= 0) { actconc[act] = a->rconc[hit]; continue; }
        for (int i = 0; i < a->R; i++) {
            double da = P->act_ve[act] - a->rve[(size_t)i * P->K + P->S];
            double d = sqrt(da * da + dist2(a, i, vo, P->S));                /* get_best_action.c:252-275 */
            double wi = 1.0 / fo_pow_x87(fabs(d), P->p);                     /* BestAct.c:215 */
            vagc += wi * a->rconc[i];
            ws += wi;
        }
        actconc[act] = vagc / ws;                                            /* :265 */
    }
    int best = 0;                                                            /* max.inl:16-28: first max wins */
    for (int i = 1; i < P->A; i++) if (actconc[best] < actconc[i]) best = i;
    return best;
}

/* frirl_check_possible_states.c:53-122 (the grid-refinement branch :68 never
 * fires for epsilon == grid step; ties -> upper grid value) */
double fo_snap(const fo_problem *P, int k, double x) {
    const double *g = P->grid[k]; int n = P->grid_len[k], i;
    for (i = 0; i < n; i++) if (x < g[i]) break;
    if (i == n) return g[n - 1];
    if (x < g[0]) return g[0];
    i--;
    return (x - g[i] < g[i + 1] - x) ? g[i] : g[i + 1];
}

/* frirl_update_sarsa (frirl_update_sarsa.c:348-385) + update_rules (:22-143) */
void fo_update_sarsa(fo_agent *a, const double *q_ant, double reward, const double *cur_q_ant) {
    const fo_problem *P = a->P;
    double qp, qnow;
    fo_vag_concl(a, cur_q_ant, &qp);
    int rule_i = fo_vag_concl(a, q_ant, &qnow);
    double qdiff = a->alpha * (reward + a->gamma * qp - qnow);               /* :358 */
    if (qdiff > P->qdiff_pos || qdiff < P->qdiff_neg) {
        double g[FO_MAXK], rconc;
        for (int k = 0; k < P->K; k++) g[k] = fo_snap(P, k, q_ant[k]);      /* :146-170 */
        rule_i = fo_vag_concl(a, g, &rconc);
        if (rule_i < 0) { a->inserted = 1; fo_add_rule(a, g, rconc + qdiff); return; }
        a->inserted = 0;
    }
    int rules = a->R - (a->inserted ? 1 : 0);
    int hit = fo_vag_concl_weight(a, q_ant, a->weights);        /* frb->weights (:39-40): untouched on an exact hit */
    if (hit >= 0) {
        if (P->skip_rules == 0 || hit < rules) a->rconc[hit] = qnow + qdiff;   /* :55 */
        return;                                                               /* :61-63 */
    }
    if (P->skip_rules == 0) { rules = a->R; a->inserted = 0; }
    for (int i = 0; i < rules; i++)
        if (a->weights[i] > P->wsig) a->rconc[i] = qnow + qdiff * a->weights[i];   /* :106-107,138 */
}

/* ---------------------------------------------------------------- environments */

#define FO_PI 3.14159265358979323846264338327

void fo_env_step(const fo_problem *P, double action, const double *s, double *n) {
    if (P->env == FO_ENV_MOUNTAINCAR) {                          /* mountaincar.c:37-74 */
        double v = (s[1] + (0.001 * action) + (-0.0025 * fo_cos(P, 3.0 * s[0]))) * 0.999;
        if (v < -0.07) v = -0.07;
        if (v > 0.07) v = 0.07;
        double x = s[0] + v;
        if (x <= -1.5) { x = -1.5; v = 0.0; }
        n[0] = x; n[1] = v;
    } else if (P->env == FO_ENV_ACROBOT) {                       /* acrobot.c:31-130 */
        const double max1 = 4 * FO_PI, max2 = 9 * FO_PI, g = 9.8, dt = 0.05;
        double t1 = s[0], t2 = s[1], w1 = s[2], w2 = s[3];
        double c2 = fo_cos(P, t2), s2 = fo_sin(P, t2);
        double d1 = 1.0 * 0.25 + 1.0 * (1.0 + 0.25 + 2 * 1.0 * 0.5 * c2) + 1.0 + 1.0;
        double d2 = 1.0 * (0.25 + 1.0 * 0.5 * c2) + 1.0;
        double phi2 = 1.0 * 0.5 * g * fo_cos(P, t1 + t2 - FO_PI / 2);
        double phi1 = -1.0 * 1.0 * 0.5 * w2 * s2 * (w2 - 2 * w1) + (1.0 * 0.5 + 1.0 * 1.0) * g * fo_cos(P, t1 - (FO_PI / 2)) + phi2;
        double acc2 = (action + phi1 * (d2 / d1) - 1.0 * 1.0 * 0.5 * w1 * w1 * s2 - phi2);
        acc2 = acc2 / (1.0 * 0.25 + 1.0 - (d2 * d2 / d1));
        double acc1 = -(d2 * acc2 + phi1) / d1;
        for (int i = 0; i < 4; i++) {                            /* :70-89 */
            w1 = w1 + acc1 * dt;
            if (w1 < -max1) w1 = -max1;
            if (w1 > max1) w1 = max1;
            t1 = t1 + w1 * dt;
            w2 = w2 + acc2 * dt;
            if (w2 < -max2) w2 = -max2;
            if (w2 > max2) w2 = max2;
            t2 = t2 + w2 * dt;
        }
        if (t1 < -FO_PI) t1 = -FO_PI;
        if (t1 > FO_PI) t1 = FO_PI;
        if (t2 < -FO_PI) t2 = -FO_PI;
        if (t2 > FO_PI) t2 = FO_PI;
        n[0] = t1; n[1] = t2; n[2] = w1; n[3] = w2;
    } else {                                                     /* cartpole.c:35-77 */
        const double g = 9.8, total_mass = 1.0 + 0.1, pml = 0.1 * 0.5, tau = 0.02, fourthirds = 4.0 / 3.0;
        double x = s[0], xd = s[1], th = s[2], thd = s[3];
        double sn = fo_sin(P, th), cs = fo_cos(P, th);
        double force = action * 10.0;
        double temp = (force + pml * thd * thd * sn) / total_mass;
        double thacc = (g * sn - cs * temp) / (0.5 * (fourthirds - 0.1 * cs * cs / total_mass));
        double xacc = temp - pml * thacc * cs / total_mass;
        n[0] = x + tau * xd; n[1] = xd + tau * xacc; n[2] = th + tau * thd; n[3] = thd + tau * thacc;
    }
}

void fo_env_reward(const fo_problem *P, const double *s, double *reward, int *done) {
    if (P->env == FO_ENV_MOUNTAINCAR) {                          /* mountaincar.c:77-95 */
        if (s[0] >= 0.45) { *reward = 1000; *done = 1; } else { *reward = -10; *done = 0; }
    } else if (P->env == FO_ENV_ACROBOT) {                       /* acrobot.c:133-162 */
        double y1 = 0.0 - fo_cos(P, s[0]);
        double y2 = y1 - fo_cos(P, s[1]);
        if (y2 >= 0.0 + 1.0) { *reward = 1000; *done = 1; } else { *reward = -10; *done = 0; }
    } else {                                                     /* cartpole.c:79-112 */
        const double deg45 = FO_PI / 4;
        double x = s[0], th = s[2], thd = s[3];
        if (x < -4.0 || x > 4.0 || th < (-1 * deg45) || th > deg45) { *reward = -10000 - 50 * fabs(x) - 100 * fabs(th); *done = 1; }
        else { *reward = 10 - 1000 * th * th - 5 * fabs(x) - 10 * thd; *done = 0; }
    }
}

void fo_env_quantize(const fo_problem *P, const double *s, double *q) {
    if (P->env != FO_ENV_CARTPOLE) {                             /* mountaincar.c:98-125, acrobot.c:165-192 */
        for (int i = 0; i < P->S; i++) {
            int where = round((s[i] + fabs(P->grid[i][0])) / P->grid_div[i]);
            if (where < 0) where = 0; else if (where > P->grid_len[i] - 1) where = P->grid_len[i] - 1;
            q[i] = P->grid[i][where];
        }
    } else {                                                     /* cartpole.c:114-168 */
        const double deg12 = FO_PI / 15, deg3 = FO_PI / 60;
        double x0 = s[0], x1 = round(s[1]), x2 = floor(s[2] / deg3) * deg3, x3 = s[3];
        if (x0 < 0) x0 = -1;
        if (x0 > 0) x0 = 1;
        if (x1 < -1) x1 = -1;
        if (x1 > 1) x1 = 1;
        if (x2 > deg12) x2 = deg12;
        if (x2 < (-1 * deg12)) x2 = -1 * deg12;
        if (x3 < 0) x3 = -1;
        if (x3 > 0) x3 = 1;
        q[0] = x0; q[1] = x1; q[2] = x2; q[3] = x3;
    }
}

/* --------------------------------------------------------------------- stepping */

/* frirl_e_greedy_selection.c:21-37 with the documented replacement RNG: one
 * Philox block per decision, counter = (episode, step, 0, 0), key = agent seed;
 * explore when u1 <= epsilon; random action = min(A-1, floor(u2*A)) (the
 * reference's round(u*A) can return A, one past the action table). */
static int select_action(fo_agent *a, const double *states, uint32_t step, double *actconc) {
    const fo_problem *P = a->P;
    int best = fo_get_best_action(a, states, actconc);
    if (P->no_random == 1 || a->epsilon == 0.0) return best;
    uint32_t ctr[4] = { a->episode_num, step, 0, 0 }, key[2] = { (uint32_t)a->seed, (uint32_t)(a->seed >> 32) }, r[4];
    fo_philox4x32_10(ctr, key, r);
    double u1 = (double)((((uint64_t)r[1] << 32) | r[0]) >> 11) * 0x1p-53;
    double u2 = (double)((((uint64_t)r[3] << 32) | r[2]) >> 11) * 0x1p-53;
    if (u1 > a->epsilon) return best;
    int ra = (int)floor(u2 * P->A);
    return ra > P->A - 1 ? P->A - 1 : ra;
}

/* one env step of frirl_episode.c:86-186 (episode start :46-82 folded in) */
static void agent_step(fo_agent *a) {
    const fo_problem *P = a->P;
    double actconc[FO_MAXA], cur[FO_MAXK], cur_q[FO_MAXK], reward; int done;
    if (a->need_reset) {
        for (int i = 0; i < P->S; i++) a->s[i] = a->qant[i] = a->has_s0 ? a->s0[i] : P->values_def[i];   /* NOT quantised, :46-48 */
        a->ep_return = 0; a->ep_steps = 0; a->step_in_ep = 0; a->need_reset = 0;
        if (a->learning) memcpy(a->prev_rconc, a->rconc, sizeof(double) * a->R);
        a->prev_numru = a->R;
        a->qant[P->S] = P->act_val[select_action(a, a->s, 0, actconc)];
    }
    a->step_in_ep++;
    fo_env_step(P, a->qant[P->S], a->s, cur);
    fo_env_reward(P, cur, &reward, &done);
    a->ep_return += reward;
    fo_env_quantize(P, cur, cur_q);
    cur_q[P->S] = P->act_val[select_action(a, cur_q, (uint32_t)a->step_in_ep, actconc)];
    a->sum_rules += a->R;
    if (a->learning) fo_update_sarsa(a, a->qant, reward, cur_q);
    memcpy(a->s, cur, sizeof(double) * P->S);
    memcpy(a->qant, cur_q, sizeof(double) * P->K);
    a->ep_steps++; a->total_steps++;
    a->success = done;
    if (done || a->step_in_ep >= P->max_steps) {
        /* "RB complete" predicate of frirl_sequential_run.c:83-150 */
        int same = (a->prev_numru == a->R) && (a->prev_steps == a->ep_steps) &&
                   (a->ep_return > P->reward_good_above) && (a->prev_reward == a->ep_return);
        if (same && a->learning) {
            a->epended = 1;                                                  /* :90 */
            int ok = 1;
            for (int i = 0; i < a->R; i++) if (fabs(a->rconc[i] - a->prev_rconc[i]) >= P->qdiff_final_tol) { ok = 0; break; }
            if (ok) a->converged = 1;
        }
        a->last_return = a->ep_return; a->last_steps = a->ep_steps; a->last_rules = a->R;
        a->prev_reward = a->ep_return; a->prev_steps = a->ep_steps;
        a->episode_num++;
        a->need_reset = 1;
    }
}

void fo_agent_run_steps(fo_agent *a, long n_steps, int stop_on_converge) {
    for (long i = 0; i < n_steps; i++) {
        if (stop_on_converge && a->converged) return;
        agent_step(a);
    }
}

int fo_run_construct(fo_agent *a, double *ep_log, int max_log) {
    int episodes = 0;
    for (int epchunk = 1; epchunk < a->P->max_episodes; epchunk++) {   /* frirl_sequential_run.c:51-65 */
        do agent_step(a); while (!a->need_reset);
        if (ep_log && episodes < max_log) {
            ep_log[4 * episodes] = a->episode_num - 1; ep_log[4 * episodes + 1] = a->last_steps;
            ep_log[4 * episodes + 2] = a->last_return; ep_log[4 * episodes + 3] = a->last_rules;
        }
        episodes++;
        if (a->converged) break;
    }
    return episodes;
}

/* ------------------------------------------------------------ knowledge sharing
 * The reference's experimental multi-agent mode (CMake option BUILD_OPENMP): frirl_omp_run, frirl_agent.c:294-385. */

/* merge_rb (frirl_agent.c:58-117): every rule of the sender, in order, is folded into the receiver's rule base */
void fo_merge_rb(fo_agent *r, int Rs, const double *srant, const double *srconc) {
    const fo_problem *P = r->P; const int K = P->K;
    for (int i = 0; i < Rs; i++) {
        const double *sant = srant + (size_t)i * K; const double sq = srconc[i];
        int rule_i = fo_vag_concl_weight(r, sant, r->weights);              /* :75-76: weights stay STALE on an exact hit */
        double rc;
        fo_vag_concl(r, sant, &rc);                                          /* :78-79 */
        double qdiff = -rc + sq;                                             /* :84 */
        if (qdiff > P->qdiff_pos || qdiff < P->qdiff_neg) {                  /* :92 */
            double g[FO_MAXK], rconc = rc;
            for (int k = 0; k < K; k++) g[k] = fo_snap(P, k, sant[k]);      /* CHECK_STATES, :96-99 + :19-45 */
            rule_i = fo_vag_concl(r, g, &rconc);
            if (rule_i < 0) { fo_add_rule(r, g, 0.5 * rconc + 0.5 * sq); continue; }   /* :100-103 */
            r->rconc[rule_i] = 0.9 * rconc + 0.1 * sq;                       /* :105 */
            continue;
        }
        const double q = 0.9 * rc + 0.1 * sq;                                /* :110 */
        for (int w = 0; w < r->R; w++)                                       /* update_rules, :47-55: an ASSIGNMENT of q * weight */
            if (r->weights[w] > P->wsig) r->rconc[w] = q * r->weights[w];
    }
}

/* frirl_omp_run (frirl_agent.c:294-385) for W agents created by fo_agent_new: start states (gen_def_states, :120-138),
 * chunks of FRIRL_AGENT_EPCHUNK - 1 = 9 episodes (frirl_sequential_run.c:51-65,159-161), then agent 0's rules into every
 * other agent and every other agent's rules into agent 0, until agent 0 finds its rule base complete.
 * log rows: (agent, steps, return, rules) per finished episode in per-agent order; returns the number of rows. */
static int team_run(fo_agent **ag, int W, int max_rounds, double *log, int max_log, int mpi_order);
int fo_team_run(fo_agent **ag, int W, int max_rounds, double *log, int max_log) { return team_run(ag, W, max_rounds, log, max_log, 0); }
/* the same with the exchange of frirl_mpi_run (frirl_agent.c:387-467): after every chunk each member's rule base is merged into the
 * leader in rank order (:427-441, no "pended" test), then the leader's - already merged - rule base into every member (:445-461) */
int fo_team_run_mpi(fo_agent **ag, int W, int max_rounds, double *log, int max_log) { return team_run(ag, W, max_rounds, log, max_log, 1); }
static int team_run(fo_agent **ag, int W, int max_rounds, double *log, int max_log, int mpi_order) {
    const fo_problem *P = ag[0]->P; const int K = P->K;
    int n = 0;
    if (W > 2) {
        const int R0 = ag[0]->R, gap = R0 / (W - 2);
        for (int id = 1; id < W; id++) {
            double s0[FO_MAXK];
            const int idx = (id - 1) * gap;                                 /* may run past the corner rules: zeros there */
            for (int i = 0; i < P->S; i++) s0[i] = idx < R0 ? ag[id]->rant[(size_t)idx * K + i] : 0.0;
            fo_agent_set_initial_state(ag[id], s0);
        }
    }
    for (int round = 0; round < max_rounds; round++) {
        int running0 = 1;
        for (int id = 0; id < W; id++) {                                    /* "#pragma omp parallel for": independent */
            fo_agent *a = ag[id];
            a->epended = 0; a->converged = 0;
            for (int epchunk = 1; epchunk < 10; epchunk++) {
                do agent_step(a); while (!a->need_reset);
                if (log && n < max_log) { log[4 * n] = id; log[4 * n + 1] = a->last_steps; log[4 * n + 2] = a->last_return; log[4 * n + 3] = a->last_rules; }
                n++;
                if (a->converged) { if (id == 0) running0 = 0; break; }
            }
            if (id == 0 && !a->converged && !(a->episode_num < (uint32_t)P->max_episodes)) running0 = 0;
        }
        if (!running0) break;
        if (mpi_order) {
            for (int id = 1; id < W; id++) fo_merge_rb(ag[0], ag[id]->R, ag[id]->rant, ag[id]->rconc);
            for (int id = 1; id < W; id++) fo_merge_rb(ag[id], ag[0]->R, ag[0]->rant, ag[0]->rconc);
            continue;
        }
        if (!ag[0]->epended)
            for (int id = 1; id < W; id++) fo_merge_rb(ag[id], ag[0]->R, ag[0]->rant, ag[0]->rconc);
        for (int id = 1; id < W; id++)
            if (!ag[id]->epended) fo_merge_rb(ag[0], ag[id]->R, ag[id]->rant, ag[id]->rconc);
    }
    return n;
}

/* ------------------------------------------------- flat accessors for ctypes use */
void fo_problem_info(const fo_problem *P, int *o) { o[0] = P->env; o[1] = P->S; o[2] = P->K; o[3] = P->A; o[4] = P->U; o[5] = P->p; o[6] = P->max_steps; o[7] = P->max_episodes; }
const double *fo_problem_u(const fo_problem *P) { return P->u; }
const double *fo_problem_ve(const fo_problem *P) { return P->ve; }
const double *fo_problem_udiv(const fo_problem *P) { return P->udiv; }
const double *fo_problem_act_val(const fo_problem *P) { return P->act_val; }
const double *fo_problem_act_ve(const fo_problem *P) { return P->act_ve; }
const double *fo_problem_grid(const fo_problem *P, int k) { return P->grid[k]; }
int fo_problem_grid_len(const fo_problem *P, int k) { return P->grid_len[k]; }
void fo_problem_set_random(fo_problem *P, int no_random) { P->no_random = no_random; }
void fo_problem_hyper(const fo_problem *P, double *o) {
    o[0] = P->alpha; o[1] = P->gamma; o[2] = P->epsilon; o[3] = P->qdiff_pos; o[4] = P->qdiff_neg;
    o[5] = P->qdiff_final_tol; o[6] = P->reward_good_above; o[7] = P->wsig; o[8] = P->skip_rules; o[9] = P->no_random;
    for (int k = 0; k < P->K; k++) { o[10 + k] = P->values_def[k]; o[18 + k] = P->grid_div[k]; }
}
int fo_agent_R(const fo_agent *a) { return a->R; }
const double *fo_agent_rant(const fo_agent *a) { return a->rant; }
const double *fo_agent_rconc(const fo_agent *a) { return a->rconc; }
int fo_agent_inserted(const fo_agent *a) { return a->inserted; }
void fo_agent_set_inserted(fo_agent *a, int v) { a->inserted = v; }
void fo_agent_set_hyper(fo_agent *a, double alpha, double gamma, double epsilon) { a->alpha = alpha; a->gamma = gamma; a->epsilon = epsilon; }
void fo_agent_set_learning(fo_agent *a, int v) { a->learning = v; }
/* per-agent episode start state (BASELINE config 4: randomised acrobot starts); NULL restores values_def */
void fo_agent_set_initial_state(fo_agent *a, const double *s0) { a->has_s0 = s0 != NULL; if (s0) memcpy(a->s0, s0, sizeof(double) * a->P->S); }
void fo_agent_summary(const fo_agent *a, double *o) {
    o[0] = a->R; o[1] = a->episode_num; o[2] = a->last_return; o[3] = a->last_steps; o[4] = a->last_rules;
    o[5] = a->converged; o[6] = (double)a->total_steps; o[7] = a->sum_rules; o[8] = a->inserted; o[9] = a->overflow;
    o[10] = a->ep_return; o[11] = a->ep_steps; o[12] = a->need_reset; o[13] = a->step_in_ep;
    for (int k = 0; k < a->P->K; k++) { o[16 + k] = a->s[k]; o[24 + k] = a->qant[k]; }
}
