/* csrc/frirl_dropin.c -- the reference's C API (include/dropin/friq_refapi.h) on top of the B200 engine.
 *
 * A `struct frirl_desc` / `struct FIVERB` pair is ONE agent: behind it sits a friq_batch of one agent
 * whose rule base lives in device memory.  Every function that evaluates or changes the rule base
 * (five_rule_distance, FIVE_vag_concl, FIVE_vag_concl_weight, FIVEVagConcl_FRIRL_BestAct,
 * frirl_get_best_action, frirl_update_sarsa, and the step loop of frirl_episode) launches the same
 * sm_100a device functions the batched engine uses; this file only marshals arguments, keeps the host
 * mirrors of the fields callers read (numofrules, rant, rconc, rseqant*, ...), runs the host callbacks
 * of the environment, and restates the reference's host-side control flow (episode / run loops, file
 * formats, command line).  Engine failures are fatal in the reference's style: message + exit().
 *
 * Each function cites the reference code whose behaviour it reproduces.
 */
#define _GNU_SOURCE
#include "../../include/dropin/friq_refapi.h"
#include "../../include/friq_b200.h"

#include <fcntl.h>
#include <getopt.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <sys/time.h>
#include <unistd.h>

typedef struct dropin_engine {
    friq_problem *P;
    friq_batch *B;
    int K;
    int mirrored;             /* rules [0, mirrored) have valid antecedent mirrors */
    double alpha, gamma;      /* values last sent to the device */
    double *wbuf;             /* [maxnumofrules] staging for FIVE_vag_concl_weight */
} dropin_engine;

#define ENG(frb) ((dropin_engine *)(frb)->friq_engine)

static void engine_fatal(const char *where, int rc)
{
    printf("%s: FATAL ERROR: B200 engine call failed (%d): %s\n", where, rc, friq_last_error());
    exit(3);
}
#define ENGINE(call) do { int rc_ = (call); if (rc_) engine_fatal(__func__, rc_); } while (0)

/* nearest universe index, get_vag_abs_min_i_fixres (src/inl/min.inl:71-92, ties -> lower) */
static unsigned int nearest_u(const double *u, int len, double x, double div)
{
    int low = (int)((x - u[0]) / div);
    if (low < 0) return 0;
    if (low >= len - 1) return (unsigned int)(len - 1);
    return (fabs(u[low] - x) <= fabs(u[low + 1] - x)) ? (unsigned int)low : (unsigned int)(low + 1);
}

/* ------------------------------------------------------------------------------------------
 * FIVERB: allocation of the host mirrors, FIVEInit.c:55-347
 * ------------------------------------------------------------------------------------------ */
static void mirror_rule(struct FIVERB *frb, int r)        /* five_add_rule.c:71-84 */
{
    const int K = frb->numofantecedents;
    for (int k = 0; k < K; k++) {
        const double v = frb->rant[(size_t)r * K + k];
        const unsigned int idx = nearest_u(frb->uk[k], frb->univlength, v, frb->udivs[k]);
        frb->rseqant[k][r] = v;
        frb->rseqant_uindex[k][r] = frb->rant_uindex[(size_t)r * K + k] = idx;
        frb->rseqant_veval[k][r] = frb->vek[k][idx];
        frb->rant_veval[(size_t)r * K + k] = (unsigned int)(frb->vek[k][idx] > 0 ? frb->vek[k][idx] : 0);   /* declared unsigned int in FIVE.h:48 */
    }
}

static void set_rule_count(struct FIVERB *frb, int R)
{
    frb->numofrules = R;
    frb->newrconc = frb->rconc + R;
    frb->newrant = frb->rant + (size_t)R * frb->numofantecedents;
    frb->avx2_rbsize = (unsigned int)(R / 4 + (R % 4 > 0));
}

/* device -> host mirrors */
static void pull_rulebase(struct FIVERB *frb)
{
    dropin_engine *e = ENG(frb);
    int R = 0;
    ENGINE(friq_batch_get_rulebase(e->B, 0, &R, frb->rant, frb->rconc));
    if (e->mirrored > R) e->mirrored = 0;
    for (int r = e->mirrored; r < R; r++) mirror_rule(frb, r);
    e->mirrored = R;
    set_rule_count(frb, R);
}

static struct FIVERB *fiverb_alloc(double *u, double *ve, int p, int K, int U, int maxrules, double *rant, double *rconc)
{
    if (K > FIVE_MAX_NUM_OF_UNIVERSES) {
        printf("FIVEInit: fatal error: %d universes given, FIVE_MAX_NUM_OF_UNIVERSES is %d\n", K, FIVE_MAX_NUM_OF_UNIVERSES);
        exit(1);
    }
    struct FIVERB *frb = (struct FIVERB *)calloc(1, sizeof *frb);
    if (!frb) return NULL;
    frb->u = u; frb->ve = ve; frb->numofunivs = K; frb->univlength = U;
    frb->maxnumofrules = maxrules; frb->rulelength = K + 1; frb->numofantecedents = K;
    frb->p = p ? p : K;                                                    /* FIVEInit.c:89-93 */
    frb->rant = rant; frb->rconc = rconc;
    frb->uk = (double **)malloc(sizeof(double *) * K); frb->vek = (double **)malloc(sizeof(double *) * K);
    frb->ukdomains = (double *)malloc(sizeof(double) * K); frb->udivs = (double *)malloc(sizeof(double) * K);
    frb->uksize = (unsigned int)(U - 1);
    for (int k = 0; k < K; k++) {
        frb->uk[k] = u + (size_t)k * U; frb->vek[k] = ve + (size_t)k * U;
        frb->ukdomains[k] = frb->uk[k][U - 1] - frb->uk[k][0];
        frb->udivs[k] = frb->ukdomains[k] / frb->uksize;                   /* FIVEInit.c:244-248 */
    }
    frb->rseqant = (double **)malloc(sizeof(double *) * K);
    frb->rseqant_uindex = (unsigned int **)malloc(sizeof(unsigned int *) * K);
    frb->rseqant_veval = (double **)malloc(sizeof(double *) * K);
    for (int k = 0; k < K; k++) {
        frb->rseqant[k] = (double *)calloc(maxrules, sizeof(double));
        frb->rseqant_uindex[k] = (unsigned int *)calloc(maxrules, sizeof(unsigned int));
        frb->rseqant_veval[k] = (double *)calloc(maxrules, sizeof(double));
    }
    frb->rant_uindex = (unsigned int *)calloc((size_t)maxrules * K, sizeof(unsigned int));
    frb->rant_veval = (unsigned int *)calloc((size_t)maxrules * K, sizeof(double));
    frb->ract = frb->rseqant[K - 1]; frb->ract_uindex = frb->rseqant_uindex[K - 1]; frb->ract_veval = frb->rseqant_veval[K - 1];
    frb->ruledists = (double *)calloc(maxrules, sizeof(double));
    frb->weights = (double *)calloc(maxrules, sizeof(double));
    frb->wi = (double *)calloc(maxrules + 4, sizeof(double));
    frb->frd_dists = (double *)calloc((size_t)(maxrules + 4) * FIVE_MAX_NUM_OF_UNIVERSES, sizeof(double));
    frb->fvc_vagdist = (double *)calloc(K, sizeof(double));
    frb->valvagp = (double *)calloc(K, sizeof(double));
    frb->valvagu = frb->uk[K - 1]; frb->valvagve = frb->vek[K - 1]; frb->valvagdims = 1;
    return frb;
}

static int g_device_override = -1;       /* frirl_mpi_run places its agents on the node's GPUs round-robin */
static int device_index(void)
{
    if (g_device_override >= 0) return g_device_override;
    const char *d = getenv("FRIQ_DEVICE");
    return d ? atoi(d) : 0;
}

/* FIVEInit (FIVEInit.c:55): stand-alone rule base over caller-supplied tables */
struct FIVERB *FIVEInit(double *u, double *ve, int p, int numofunivs, int univlength, int numofrules, int maxnumofrules,
                        int rulelength, double *rant, double *rconc)
{
    if (rulelength != numofunivs + 1) {
        printf("FIVEInit: fatal error: rulelength %d != numofunivs + 1 (consequent universes are not supported by the B200 engine)\n", rulelength);
        exit(1);
    }
    struct FIVERB *frb = fiverb_alloc(u, ve, p, numofunivs, univlength, maxnumofrules, rant, rconc);
    if (!frb) return NULL;
    dropin_engine *e = (dropin_engine *)calloc(1, sizeof *e);
    frb->friq_engine = e;
    e->K = numofunivs;
    e->wbuf = (double *)malloc(sizeof(double) * maxnumofrules);
    ENGINE(friq_problem_create_from_tables(numofunivs, univlength, frb->p, u, ve, &e->P));
    ENGINE(friq_batch_create(e->P, device_index(), 1, maxnumofrules, 0, 0, &e->B));
    if (numofrules > 0) ENGINE(friq_batch_set_rulebase(e->B, 0, numofrules, rant, rconc));
    pull_rulebase(frb);
    return frb;
}

void five_deinit(struct FIVERB *frb)                       /* five_deinit.c */
{
    if (!frb) return;
    dropin_engine *e = ENG(frb);
    if (e) { friq_batch_destroy(e->B); friq_problem_destroy(e->P); free(e->wbuf); free(e); }
    for (int k = 0; k < frb->numofantecedents; k++) { free(frb->rseqant[k]); free(frb->rseqant_uindex[k]); free(frb->rseqant_veval[k]); }
    free(frb->rseqant); free(frb->rseqant_uindex); free(frb->rseqant_veval); free(frb->rant_uindex); free(frb->rant_veval);
    free(frb->uk); free(frb->vek); free(frb->ukdomains); free(frb->udivs);
    free(frb->ruledists); free(frb->weights); free(frb->wi); free(frb->frd_dists); free(frb->fvc_vagdist); free(frb->valvagp);
    free(frb);
}

/* ------------------------------------------------------------------------------------------
 * init-time table builders (host; plain sequential FP64, this file is compiled with -ffp-contract=off)
 * ------------------------------------------------------------------------------------------ */
/* FIVE_GSc_func (FIVEGScFunc.c:76-255).  psc holds mp scaling points of np numbers each:
 * np = 1: one scaling value for the whole universe; np = 2: (P, S); np = 3: (P, S_left, S_right).
 * As in the reference the non-linear branch is unreachable (its `c` is forced to NAN, :83-91), so the
 * function between two points is constant when the facing values agree and linear otherwise. */
int FIVE_GSc_func(double *u, int numofunivs, int univlength, double *psc, int mp, int np, double nls, double *scf)
{
    if (nls <= 0) return -1;
    const int total = numofunivs * univlength;
    if (mp <= 1) {
        if (np == 1) { for (int i = 0; i < total; i++) scf[i] = psc[0]; return 0; }
        for (int i = 0; i < numofunivs; i++)
            for (int j = 0; j < univlength; j++) {
                const double x = u[i * univlength + j];
                scf[i * univlength + j] = (x < psc[0]) ? psc[1] : psc[i * np + (np == 2 ? 1 : 2)];
            }
        return 0;
    }
    const double *last = psc + (size_t)(mp - 1) * np;
    for (int i = 0; i < numofunivs; i++)
        for (int j = 0; j < univlength; j++) {
            double *out = &scf[i * univlength + j];
            const double x = u[i * univlength + j];
            if (x < psc[0]) *out = psc[1];
            for (const double *a = psc; a != last; a += np) {
                const double *b = a + np;
                if (!(x >= a[0] && x < b[0])) continue;
                const double s1 = a[np == 2 ? 1 : 2], s2 = b[1];
                *out = (s1 == s2) ? s1 : ((s2 - s1) / (b[0] - a[0])) * (x - a[0]) + s1;
            }
            if (x >= last[0]) {
                if (np == 2) *out = last[1];
                else *out = (j == univlength - 1 && x == last[0]) ? last[1] : last[2];
            }
        }
    return 0;
}

double *FIVEGScFunc(double *u, int numofunivs, int univlength, double *psc, int mp, int np, double nls)
{
    double *scf = (double *)calloc((size_t)numofunivs * univlength, sizeof(double));
    if (scf && FIVE_GSc_func(u, numofunivs, univlength, psc, mp, np, nls, scf) != 0) { free(scf); return NULL; }
    return scf;
}

/* FIVEGVagEnv (FIVEGVagEnv.c:40-102): primitive integral of the scaling function (first element 0), or, when a
 * scaling value is infinite, the trapezoid areas between neighbours (first element -1) */
double *FIVEGVagEnv(double *u, int numofunivs, int univlength, double *scf)
{
    double *ve = (double *)malloc(sizeof(double) * numofunivs * univlength);
    if (!ve) return NULL;
    for (int k = 0; k < numofunivs; k++) {
        const double *uk = u + (size_t)k * univlength, *sk = scf + (size_t)k * univlength;
        double *vk = ve + (size_t)k * univlength;
        int has_inf = 0;
        for (int j = 0; j < univlength; j++) if (sk[j] == INFINITY) { has_inf = 1; break; }
        vk[0] = has_inf ? -1 : 0;
        for (int j = 0; j + 1 < univlength; j++) {
            const double area = (uk[j + 1] - uk[j]) * (sk[j] + sk[j + 1]) * 0.5;
            vk[j + 1] = has_inf ? area : vk[j] + area;
        }
    }
    return ve;
}

/* FIVEValVag (FIVEValVag.c:37-130, FIVE_NOINF): value of the consequent universe at scaled distance vp */
int FIVEValVag(struct FIVERB *frb, double *vp)
{
    const int U = frb->univlength;
    for (int k = 0; k < frb->valvagdims; k++) {
        const double *vve = frb->valvagve + (size_t)k * U;
        int i;
        if (frb->ve[(size_t)k * U] < 0) {
            i = 1;
            double acc = vve[i];
            while (acc < vp[k] && i < U - 1) { i++; acc += vve[i]; }
            if ((acc - vp[k]) > (vp[k] - (acc - vve[i]))) i--;
        } else {
            if (vp[k] < 0 || vp[k] > vve[U - 1]) return -1;
            i = 0;
            double best = fabs(frb->valvagve[0] - vp[k]);
            for (int l = 1; l < U; l++) { const double t = fabs(vve[l] - vp[k]); if (best > t) { best = t; i = l; } }
        }
        frb->valvagp[k] = frb->valvagu[i];
    }
    return 0;
}

/* five_vague_distance (five_vague_distance.c:42-95, default flags): per-dimension scaled distance of two points */
int five_vague_distance(struct FIVERB *frb, fri_float *p1, fri_float *p2, fri_float *d)
{
    for (int k = 0; k < frb->numofunivs; k++) {
        unsigned int i = nearest_u(frb->uk[k], frb->univlength, p1[k], frb->udivs[k]);
        unsigned int j = nearest_u(frb->uk[k], frb->univlength, p2[k], frb->udivs[k]);
        if (i > j) { const unsigned int t = i; i = j; j = t; }
        const double *ve = frb->vek[k];
        if (ve[0] >= 0) d[k] = ve[j] - ve[i];
        else { double s = 0; for (unsigned int l = i + 1; l < j; l++) s += ve[l]; d[k] = s; }
    }
    return 0;
}

/* five_vague_distance_parallel (five_vague_distance_parallel.c:40-64): four points p1[i * p1_offset] against p2 */
int five_vague_distance_parallel(struct FIVERB *frb, fri_float *p1, int p1_offset, fri_float *p2, fri_float *d)
{
    for (int k = 0; k < frb->numofunivs; k++) {
        const double div = (frb->uk[k][frb->univlength - 1] - frb->uk[k][0]) / (frb->univlength - 1);
        const unsigned int j = nearest_u(frb->uk[k], frb->univlength, p2[k], div);
        for (int i = 0; i < 4; i++) {
            unsigned int a = nearest_u(frb->uk[k], frb->univlength, p1[(size_t)i * p1_offset + k], div), b = j;
            if (a > b) { const unsigned int t = a; a = b; b = t; }
            d[k * 4 + i] = frb->vek[k][b] - frb->vek[k][a];
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * FIVE evaluation: device
 * ------------------------------------------------------------------------------------------ */
int five_rule_distance(struct FIVERB *frb, fri_float *x)                        /* five_rule_distance.c:63 */
{
    int hit = -1;
    ENGINE(friq_batch_rule_distance(ENG(frb)->B, x, frb->ruledists, &hit));
    return hit;                                                                  /* ~0 == -1: no exact hit */
}

unsigned int FIVE_vag_concl(struct FIVERB *frb, double *ant, double *conc)      /* FIVEVagConcl.c:64 */
{
    int hit = -1;
    ENGINE(friq_batch_vag_concl(ENG(frb)->B, ant, conc, &hit));
    return hit < 0 ? ~0u : (unsigned int)hit;
}

double FIVEVagConcl(struct FIVERB *frb, double *x)                              /* FIVEVagConcl.c:55-61 */
{
    double conc = 0.0;
    FIVE_vag_concl(frb, x, &conc);
    return conc;
}

unsigned int FIVE_vag_concl_weight(struct FIVERB *frb, double *ant, double *weights)   /* FIVEVagConclWeight.c:52 */
{
    dropin_engine *e = ENG(frb);
    int hit = -1;
    ENGINE(friq_batch_vag_concl_weight(e->B, ant, &hit, e->wbuf));
    if (hit >= 0) return (unsigned int)hit;                                      /* weights untouched, :96-99 */
    memcpy(weights, e->wbuf, sizeof(double) * frb->numofrules);
    return ~0u;
}

unsigned int FIVEVagConclWeight(struct FIVERB *frb, double *x)                   /* FIVEVagConclWeight.c:47-49 */
{
    return FIVE_vag_concl_weight(frb, x, frb->weights);
}

double FIVEVagConcl_FRIRL_BestAct(struct FIVERB *frb, double *ruledists)         /* FIVEVagConcl_FRIRL_BestAct.c:56 */
{
    double conc = 0.0;
    ENGINE(friq_batch_bestact_from_dists(ENG(frb)->B, ruledists, &conc));
    return conc;
}

/* ------------------------------------------------------------------------------------------
 * rule-base edits
 * ------------------------------------------------------------------------------------------ */
int FIVE_add_rule(struct FIVERB *frb, fri_float *rant, fri_float rconc)          /* five_add_rule.c:47-95 */
{
    dropin_engine *e = ENG(frb);
    const int K = frb->numofantecedents, R = frb->numofrules;
    ENGINE(friq_batch_append_rule(e->B, 0, rant, rconc));
    memmove(frb->rant + (size_t)R * K, rant, sizeof(double) * K);
    frb->rconc[R] = rconc;
    mirror_rule(frb, R);
    if (e->mirrored == R) e->mirrored = R + 1;
    set_rule_count(frb, R + 1);
    return 0;
}
int FIVEAddRule(struct FIVERB *frb, double *newrule) { return FIVE_add_rule(frb, newrule, newrule[frb->numofantecedents]); }
int five_add_rule(struct FIVERB *frb, fri_float *ruletoadd) { return FIVE_add_rule(frb, ruletoadd, ruletoadd[frb->numofantecedents]); }

int five_remove_rule(struct FIVERB *frb, unsigned int rulenotoremove)            /* five_remove_rule.c:29-85 */
{
    dropin_engine *e = ENG(frb);
    const int K = frb->numofantecedents;
    if (frb->numofrules < 1 || rulenotoremove > (unsigned int)(frb->numofrules - 1)) {
        printf("FIVERemoveRule - FATAL: Invalid rule: %u max: %d !\n", rulenotoremove, frb->numofrules - 1);
        exit(6);
    }
    printf("FIVERemoveRule: rule to remove: %u/%d. =", rulenotoremove, frb->numofrules - 1);     /* :31-37 (DEBUG is forced on, :16) */
    for (int k = 0; k < K; k++) printf(" %f", frb->rseqant[k][rulenotoremove]);
    printf(" %f\n", frb->rconc[rulenotoremove]);
    ENGINE(friq_batch_remove_rule(e->B, 0, (int)rulenotoremove));
    e->mirrored = 0;
    pull_rulebase(frb);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * FRIRL: init / deinit
 * ------------------------------------------------------------------------------------------ */
void frirl_gen_fixres_arr(fri_float *arr, int len, fri_float div)                /* frirl_app_helpers.c:32-44 */
{
    const fri_float first = -((len - 1) * div) / 2;
    int i = 0;
    for (; i < len / 2 + 1; i++) arr[i] = first + div * i;
    for (; i < len; i++) arr[i] = arr[len - 1 - i] * -1;
}

/* frirl_init_ve (frirl_init_ve.c:25-121): one vague environment per dimension from constant-steepness scaling
 * points at the grid values; the action dimension uses A equidistant points on [-1, 1] with the INTEGER
 * steepness (A - 1) / 2 (:85-92) */
int frirl_init_ve(struct frirl_desc *frirl, fri_float *ve, fri_float *u, int univlength)
{
    const int S = frirl->statedims_len;
    double *scf = (double *)malloc(sizeof(double) * univlength);
    for (int k = 0; k <= S; k++) {
        const int n = (k < S) ? frirl->statedims[k].values_len : frirl->actiondim.values_len;
        double *pts = (double *)malloc(sizeof(double) * 3 * n);
        for (int i = 0; i < n; i++) {
            if (k < S) { pts[3 * i] = frirl->statedims[k].values[i]; pts[3 * i + 1] = pts[3 * i + 2] = frirl->statedims[k].values_steep; }
            else { pts[3 * i] = i * (1.0 / (n - 1) * 2.0) - 1.0; pts[3 * i + 1] = pts[3 * i + 2] = (n - 1) / 2; }
        }
        FIVE_GSc_func(u + (size_t)k * univlength, 1, univlength, pts, n, 3, NAN, scf);
        double *row = FIVEGVagEnv(u + (size_t)k * univlength, 1, univlength, scf);
        memcpy(ve + (size_t)k * univlength, row, sizeof(double) * univlength);
        free(row); free(pts);
    }
    free(scf);
    return 0;
}

/* frirl_init_rb (frirl_init_rb.c:86-147): the 2^K corner rules in binary-counter order (dimension 0 is the most
 * significant bit), all consequents 0 */
int frirl_init_rb(struct frirl_desc *frirl, fri_float *rant, fri_float *rconc, int *numofrules)
{
    const int K = frirl->numofantecedents, S = frirl->statedims_len;
    const int n = 1 << K;
    *numofrules = n;
    for (int r = 0; r < n; r++) rconc[r] = 0.0;
    for (int k = 0; k < K; k++) {
        const struct frirl_dimension_desc *d = (k < S) ? &frirl->statedims[k] : &frirl->actiondim;
        double lo = d->values[0], hi = d->values[0];
        for (int j = 1; j < d->values_len; j++) { if (d->values[j] < lo) lo = d->values[j]; if (d->values[j] > hi) hi = d->values[j]; }
        const int block = n >> (k + 1);
        for (int r = 0; r < n; r++) rant[(size_t)r * K + k] = ((r / block) % 2 == 0) ? lo : hi;
    }
    return 0;
}

static void sync_params(struct frirl_desc *frirl)
{
    dropin_engine *e = ENG(frirl->fiverb);
    if (e->alpha != frirl->alpha || e->gamma != frirl->gamma) {
        ENGINE(friq_batch_set_hyper(e->B, &frirl->alpha, &frirl->gamma, NULL));
        e->alpha = frirl->alpha; e->gamma = frirl->gamma;
    }
    ENGINE(friq_problem_set_learning_params(e->P, frirl->qdiff_pos_boundary, frirl->qdiff_neg_boundary,
                                            frirl->rule_weight_considered_significant_for_update, frirl->skip_rules));
}

int frirl_init(struct frirl_desc *frirl)                                         /* frirl_init.c:29-341 */
{
    const int S = frirl->statedims_len, K = S + 1, U = frirl->statedims[0].universe_len, A = frirl->actiondim.values_len;
    frirl->numofantecedents = K;

    /* the engine builds the tables (bit-identical to frirl_init_ve / FIVEInit, tests/test_oracle_golden.py) */
    friq_dim_desc *dims = (friq_dim_desc *)calloc(S, sizeof *dims);
    friq_problem_desc pd;
    memset(&pd, 0, sizeof pd);
#define COPY_DIM(dst, src) do { (dst).values_len = (src).values_len; (dst).values = (src).values; (dst).values_div = (src).values_div; \
        (dst).values_steep = (src).values_steep; (dst).values_def = (src).values_def; (dst).universe_len = (src).universe_len;      \
        (dst).universe = (src).universe; (dst).universe_div = (src).universe_div; } while (0)
    for (int k = 0; k < S; k++) COPY_DIM(dims[k], frirl->statedims[k]);
    COPY_DIM(pd.actiondim, frirl->actiondim);
    pd.env = FRIQ_ENV_HOST_CALLBACKS; pd.statedims_len = S; pd.statedims = dims;
    pd.alpha = frirl->alpha; pd.gamma = frirl->gamma; pd.epsilon = frirl->epsilon;
    pd.qdiff_pos_boundary = frirl->qdiff_pos_boundary; pd.qdiff_neg_boundary = frirl->qdiff_neg_boundary;
    pd.qdiff_final_tolerance = frirl->qdiff_final_tolerance; pd.reward_good_above = frirl->reward_good_above;
    pd.rule_weight_considered_significant_for_update = frirl->rule_weight_considered_significant_for_update;
    pd.skip_rules = frirl->skip_rules; pd.no_random = frirl->no_random; pd.max_episodes = frirl->max_episodes; pd.max_steps = frirl->max_steps;

    dropin_engine *e = (dropin_engine *)calloc(1, sizeof *e);
    e->K = K; e->alpha = frirl->alpha; e->gamma = frirl->gamma;
    int rc = friq_problem_create(&pd, &e->P);
    free(dims);
    if (rc) { printf("frirl_init: FATAL ERROR: %s\n", friq_last_error()); exit(1); }

    double *u = (double *)malloc(sizeof(double) * K * U), *ve = (double *)malloc(sizeof(double) * K * U);
    double *rant = (double *)calloc((size_t)frirl->five_maxnumofrules * K, sizeof(double));
    double *rconc = (double *)calloc(frirl->five_maxnumofrules, sizeof(double));
    if (!u || !ve || !rant || !rconc) return -1;
    double *act_ve = (double *)malloc(sizeof(double) * A);
    friq_problem_tables(e->P, u, ve, NULL, NULL, act_ve);

    /* the reference's sanity checks (frirl_init.c:160-245) */
    for (int k = 0; k < K; k++)
        if (u[(size_t)k * U + U - 1] - u[(size_t)k * U] == 0.0) {
            printf("frirl_init: FATAL ERROR: U dimension %d max-min=0.0! Did you cast the constant to double explicitly? e.g. 1001.0 instead of 1001.\n", k);
            exit(1);
        }
    for (int k = 0; k <= S; k++) {
        const struct frirl_dimension_desc *d = (k < S) ? &frirl->statedims[k] : &frirl->actiondim;
        const double ulo = u[(size_t)k * U], uhi = u[(size_t)k * U + U - 1], vlo = d->values[0], vhi = d->values[d->values_len - 1];
        if (vlo < ulo || vhi > uhi) {
            if (k < S) printf("frirl_init: FATAL ERROR: The points are out of range for statedim%d. Should be between %lf and %lf, but found %lf - %lf! Check u generation and the statedim boundaries!\n", k, ulo, uhi, vlo, vhi);
            else printf("frirl_init: FATAL ERROR: The points are out of range for the action dimension. Should be between %lf and %lf, but found %lf - %lf! Check u generation and the statedim boundaries for the action dimension!\n", ulo, uhi, vlo, vhi);
            exit(1);
        }
        if (k < S && (vlo > d->values_def || vhi < d->values_def))
            printf("WARNING: default state value for statedim%d is out of range! Should be true: %lf < %lf < %lf\n", k, vlo, d->values_def, vhi);
        for (int j = 1; j < d->values_len; j++) {
            const double step = d->values[j] - d->values[j - 1];
            if (!(step > d->values_div - 0.00001 && step < d->values_div + 0.00001)) {
                if (k < S) printf("frirl_init: WARNING: state dim step or divider is possibly wrong (or innaccurate): statedim%d[%d]=%lf statedim%d[%d]=%lf diff is: %lf and divider is: %lf\n", k, j - 1, d->values[j - 1], k, j, d->values[j], step, d->values_div);
                else printf("frirl_init: WARNING: action dim step or divider is possibly wrong (or innaccurate): actions[%d]=%lf actions[%d]=%lf diff is: %lf and divider is: %lf\n", j - 1, d->values[j - 1], j, d->values[j], step, d->values_div);
            }
        }
    }

    /* possible states / actions (frirl_init.c:55-102, 156-158) */
    frirl->possible_states = (struct frirl_values_desc *)calloc(S, sizeof(struct frirl_values_desc));
    for (int k = 0; k < S; k++) {
        struct frirl_values_desc *ps = &frirl->possible_states[k];
        ps->values_len = frirl->statedims[k].values_len;
        ps->values = (double *)malloc(sizeof(double) * ps->values_len);
        memcpy(ps->values, frirl->statedims[k].values, sizeof(double) * ps->values_len);
        ps->epsilon = frirl->statedims[k].values_div;
    }
    frirl->possible_actions = (struct frirl_values_desc *)calloc(1, sizeof(struct frirl_values_desc));
    frirl->possible_actions->values_len = A;
    frirl->possible_actions->values = (double *)malloc(sizeof(double) * A);
    memcpy(frirl->possible_actions->values, frirl->actiondim.values, sizeof(double) * A);
    frirl->possible_actions->vevalues = act_ve;
    frirl->possible_actions->epsilon = frirl->actiondim.values_div;

    struct FIVERB *frb = fiverb_alloc(u, ve, 0, K, U, frirl->five_maxnumofrules, rant, rconc);
    if (!frb) return -1;
    frb->friq_engine = e;
    e->wbuf = (double *)malloc(sizeof(double) * frirl->five_maxnumofrules);
    rc = friq_batch_create(e->P, device_index(), 1, frirl->five_maxnumofrules, 0, 0, &e->B);   /* corner rule base on the device */
    if (rc) { printf("frirl_init: FATAL ERROR: %s\n", friq_last_error()); exit(1); }
    frirl->fiverb = frb;
    pull_rulebase(frb);
    frirl->fiverb_ua = u + (size_t)U * S; frirl->fiverb_vea = ve + (size_t)U * S;

    frirl->reduction_state = 0;
    frirl->reward.ep_total_value = -1; frirl->reward.ep_total_steps = -1;                      /* frirl_init.c:144-145 */
    frirl->fus_is_rule_inserted = 0;
    const int mr = frirl->five_maxnumofrules;
    frirl->fgba_vagdist_states = (double *)calloc((size_t)(mr + 4) * S, sizeof(double));
    frirl->fgba_ruledist = (double *)calloc(mr, sizeof(double));
    frirl->fgba_actconc = (double *)calloc(mr, sizeof(double));
    frirl->fgba_dists = (double *)calloc((size_t)(mr + 4) * FIVE_MAX_NUM_OF_UNIVERSES, sizeof(double));
    frirl->fgba_statedistsum = (double *)calloc(mr, sizeof(double));
    frirl->fus_values = (double *)calloc(K + 1, sizeof(double));
    frirl->fus_proposed_values = (double *)calloc(K + 1, sizeof(double));
    frirl->fus_check_states = (double *)calloc(K, sizeof(double));
    frirl->fep_ant = (double *)calloc(K, sizeof(double)); frirl->fep_cur_ant = (double *)calloc(K, sizeof(double));
    frirl->fep_q_ant = (double *)calloc(K, sizeof(double)); frirl->fep_cur_q_ant = (double *)calloc(K, sizeof(double));
    frirl->is_running = 1;
    frirl->episode_num = 1;
    struct timeval t;
    gettimeofday(&t, NULL);
    srand((unsigned int)(t.tv_usec * t.tv_sec));                                                /* frirl_init.c:335-336 */
    return 0;
}

void frirl_deinit(struct frirl_desc *frirl)                                      /* frirl_deinit.c:18-60 */
{
    struct FIVERB *frb = frirl->fiverb;
    free(frb->u); free(frb->ve); free(frb->rant); free(frb->rconc);
    five_deinit(frb);
    frirl->fiverb = NULL;
    free(frirl->fus_proposed_values); free(frirl->fus_values); free(frirl->fus_check_states);
    free(frirl->fgba_statedistsum); free(frirl->fgba_vagdist_states); free(frirl->fgba_ruledist); free(frirl->fgba_actconc); free(frirl->fgba_dists);
    free(frirl->fep_ant); free(frirl->fep_cur_ant); free(frirl->fep_q_ant); free(frirl->fep_cur_q_ant);
    for (int k = 0; k < frirl->statedims_len; k++) free(frirl->possible_states[k].values);
    free(frirl->possible_states);
    free(frirl->possible_actions->values); free(frirl->possible_actions->vevalues); free(frirl->possible_actions);
    if (frirl->rbfile) { free(frirl->rbfile); frirl->rbfile = NULL; }
}

/* ------------------------------------------------------------------------------------------
 * FRIRL: action selection, update, episode
 * ------------------------------------------------------------------------------------------ */
unsigned int frirl_get_best_action(struct frirl_desc *frirl, fri_float *states)  /* frirl_get_best_action.c:31 */
{
    int best = 0;
    ENGINE(friq_batch_get_best_action(ENG(frirl->fiverb)->B, states, frirl->fgba_actconc, &best));
    return (unsigned int)best;
}

/* the host's share of frirl_e_greedy_selection.c:21-37: -1 = act greedily, else the drawn action.  The reference's
 * round(u * A) can return A, one past the action table; it is clamped to A - 1 here. */
static int draw_random_action(struct frirl_desc *frirl)
{
    if (frirl->no_random == 1 || frirl->epsilon == 0.0) return -1;
    if ((double)rand() / (double)RAND_MAX > frirl->epsilon) return -1;
    const int a = (int)round((double)rand() / (double)RAND_MAX * frirl->actiondim.values_len);
    return a > frirl->actiondim.values_len - 1 ? frirl->actiondim.values_len - 1 : a;
}

unsigned int frirl_e_greedy_selection(struct frirl_desc *frirl, fri_float *states)
{
    const int forced = draw_random_action(frirl);
    return forced >= 0 ? (unsigned int)forced : frirl_get_best_action(frirl, states);
}

fri_float frirl_check_possible_states(struct frirl_desc *frirl, fri_float observation, struct frirl_values_desc *ps)
{   /* frirl_check_possible_states.c:53-122 (ties -> upper value).  The grid-refinement branch (:68) needs
     * epsilon <= step / 2; the engine keeps grids fixed (friq_problem_create rejects such problems). */
    (void)frirl;
    int i = 0;
    while (i < ps->values_len && !(observation < ps->values[i])) i++;
    if (i == ps->values_len) return ps->values[ps->values_len - 1];
    if (i == 0) return ps->values[0];
    i--;
    return (observation - ps->values[i] < ps->values[i + 1] - observation) ? ps->values[i] : ps->values[i + 1];
}

void frirl_update_sarsa(struct frirl_desc *frirl, fri_float *q_ant, fri_float reward, fri_float *cur_q_ant)
{   /* frirl_update_sarsa.c:348-385 */
    dropin_engine *e = ENG(frirl->fiverb);
    int ins = 0;
    sync_params(frirl);
    ENGINE(friq_batch_update_sarsa(e->B, q_ant, &reward, cur_q_ant));
    ENGINE(friq_batch_get_inserted(e->B, &ins));
    frirl->fus_is_rule_inserted = ins;
    pull_rulebase(frirl->fiverb);
}

void frirl_episode(struct frirl_desc *frirl)                                     /* frirl_episode.c:28-194 */
{
    dropin_engine *e = ENG(frirl->fiverb);
    const int S = frirl->statedims_len, K = frirl->numofantecedents;
    fri_float *q_ant = frirl->fep_q_ant, *cur_q_ant = frirl->fep_cur_q_ant, *states = frirl->fep_ant, *cur_states = frirl->fep_cur_ant;
    sync_params(frirl);
    for (int k = 0; k < S; k++) q_ant[k] = states[k] = frirl->statedims[k].values_def;       /* NOT quantised, :46-48 */
    frirl->reward.ep_total_value = 0;
    frirl->reward.ep_total_steps = 0;
    q_ant[S] = frirl->actiondim.values[frirl_e_greedy_selection(frirl, states)];             /* :78-82 */
    for (int step = 1; step <= frirl->max_steps; step++) {
        frirl->do_action_func(frirl, q_ant[S], states, S, cur_states);                       /* :97 */
        frirl->get_reward_func(frirl, cur_states, S, &frirl->reward);                        /* :106-107 */
        frirl->reward.ep_total_value += frirl->reward.value;
        frirl->quant_obs_func(frirl, cur_states, S, cur_q_ant);                              /* :112 */
        /* :148-160 - proposed action + SARSA update, one kernel */
        const int forced = draw_random_action(frirl);
        int action = 0, rules = 0, inserted = 0;
        ENGINE(friq_batch_classic_step(e->B, q_ant, &frirl->reward.value, cur_q_ant, &forced, frirl->reduction_state == 0,
                                       &action, frirl->fgba_actconc, &rules, &inserted));
        cur_q_ant[S] = frirl->actiondim.values[action];
        frirl->fiverb->numofrules = rules;
        frirl->fus_is_rule_inserted = inserted;
        memcpy(states, cur_states, sizeof(double) * S);                                      /* :163-168 */
        memcpy(q_ant, cur_q_ant, sizeof(double) * K);
        frirl->reward.ep_total_steps++;
        if (frirl->reward.success == 1) break;                                               /* :183-185 */
    }
    pull_rulebase(frirl->fiverb);
}

/* ------------------------------------------------------------------------------------------
 * run loops
 * ------------------------------------------------------------------------------------------ */
static void reduce_rule_base(struct frirl_desc *frirl);

void frirl_sequential_run(struct frirl_desc *frirl)                              /* frirl_sequential_run.c:24-355 */
{
    struct FIVERB *frb = frirl->fiverb;
    frirl->epended = 0;
    if (frirl->construct_rb == 1) {
        double *prev_rconc = (double *)malloc(sizeof(double) * frirl->five_maxnumofrules);
        const int chunk_limit = (frirl->runmode == FRIRL_MPI || frirl->runmode == FRIRL_OMP) ? FRIRL_AGENT_EPCHUNK : frirl->max_episodes;
        for (int chunk = 1; ; chunk++) {
            if (!(chunk < chunk_limit)) {                                                    /* :59-65 */
                if (!(frirl->episode_num < (unsigned int)frirl->max_episodes)) frirl->is_running = 0;
                break;
            }
            const int prev_rules = frb->numofrules, prev_steps = frirl->reward.ep_total_steps;
            const double prev_reward = frirl->reward.ep_total_value;
            memcpy(prev_rconc, frb->rconc, sizeof(double) * frirl->five_maxnumofrules);
            frirl_episode(frirl);
            printf("#%d Episode: %d\tSteps: %d\tReward: %s%f%s\tRules: %d\n", frirl->agent_id, frirl->episode_num, frirl->reward.ep_total_steps,
                   frirl->reward.ep_total_value > frirl->reward_good_above ? TERM_GREEN : TERM_RED, frirl->reward.ep_total_value, TERM_NC, frb->numofrules);
            int complete = 0;
            if ((prev_rules == frb->numofrules && prev_steps == frirl->reward.ep_total_steps &&
                 frirl->reward.ep_total_value > frirl->reward_good_above && prev_reward == frirl->reward.ep_total_value) || frirl->user_exited == 1) {
                complete = 1;                                                                /* :83-150 */
                frirl->epended = 1;
                if (frirl->verbose > 0) printf("Rule-base size and reward are the same as in the previous iteration.\n");
                for (int i = 0; i < frb->numofrules; i++) {
                    if (fabs(frb->rconc[i] - prev_rconc[i]) >= frirl->qdiff_final_tolerance) {
                        complete = 0;
                        if (frirl->verbose <= 0) break;
                        printf("Greater at rule %d. %.18f - %.18f = %.18f (max: %.18f)\n", i, frb->rconc[i], prev_rconc[i], frb->rconc[i] - prev_rconc[i], frirl->qdiff_final_tolerance);
                    }
                }
            }
            if (complete || frirl->user_exited == 1) {
                frirl->is_running = 0;
                printf("-----------------------------------------------------------------\n");
                printf("No more significant changes in rule-base. RB considered complete.\n");
                printf("-----------------------------------------------------------------\n");
                break;
            }
            frirl->episode_num++;
        }
        free(prev_rconc);
    }
    if (frirl->reduce_rb == 1) reduce_rule_base(frirl);
}

/* Rule-base reduction, strategies 1 (smallest |Q| first) and 2 (largest |Q| first),
 * frirl_sequential_run.c:170-350: take one rule out, replay one greedy episode without updates; keep the
 * removal when steps and reward are unchanged, else restore the rule base and mark the rule important (NaN in
 * the candidate list). */
static void reduce_rule_base(struct frirl_desc *frirl)
{
    struct FIVERB *frb = frirl->fiverb;
    const int mr = frirl->five_maxnumofrules, K = frirl->numofantecedents;
    double *cand = (double *)malloc(sizeof(double) * mr), *cand_prev = (double *)malloc(sizeof(double) * mr);
    double *removed = (double *)malloc(sizeof(double) * (K + 1));
    memcpy(cand, frb->rconc, sizeof(double) * mr);
    frirl->original_learning = 1;
    frirl->reduction_state = 1;
    const int by_strategy = (frirl->reduction_strategy == 1 || frirl->reduction_strategy == 2);
    const unsigned int iterations = by_strategy ? (unsigned int)(frb->numofrules + 1) : 10000u;
    frirl_episode(frirl);
    const int steps_ref = frirl->reward.ep_total_steps;
    double prev_reward = 0.0;
    unsigned int pick = 0;
    for (frirl->episode_num = 1; frirl->episode_num <= iterations; frirl->episode_num++) {
        frirl_episode(frirl);
        printf("Reduction Episode: %d\tSteps: %d\tReward: %f\tEpsilon: %f\tRules: %d\n", frirl->episode_num, frirl->reward.ep_total_steps,
               frirl->reward.ep_total_value, frirl->epsilon, frb->numofrules);
        if (!by_strategy) continue;
        if (frirl->episode_num > 1) {
            const double diff = prev_reward - frirl->reward.ep_total_value;
            if (frirl->reward.ep_total_value > frirl->reward_good_above && frirl->reward.ep_total_steps == steps_ref &&
                fabs(diff) <= frirl->reduction_reward_tolerance) {
                if (frirl->verbose > 0) {
                    printf("Reduction Episode: %d\tEliminated rule: no: %d. - ", frirl->episode_num, pick + 1);
                    for (int j = 0; j <= K; j++) printf(" %f", removed[j]);
                    printf(" \tReward diff was: %f\n", diff);
                }
                prev_reward = frirl->reward.ep_total_value;
            } else {
                memcpy(cand, cand_prev, sizeof(double) * mr);
                cand[pick] = NAN;
                if (frirl_load_rb_from_bin_file(frirl, "reduction_tmp.frirlrb.bin") < 0) { printf("Error while loading the binary rule-base file!\n"); exit(-1); }
                frb = frirl->fiverb;
                if (frirl->verbose > 0) {
                    printf("Reduction Episode: %d\tRule stays: no: %d. - ", frirl->episode_num, pick + 1);
                    for (int j = 0; j < K; j++) printf(" %f", frb->rseqant[j][pick]);
                    printf(" %f\n", frb->rconc[pick]);
                }
            }
        } else prev_reward = frirl->reward.ep_total_value;
        /* next candidate: extreme |Q| among the rules not marked important (:262-275, 305-318) */
        double mv = fabs(cand[0]);
        pick = 0;
        for (int j = 1; j < frb->numofrules; j++) {
            const double a = fabs(cand[j]);
            const int better = (frirl->reduction_strategy == 1) ? (mv > a) : (mv < a);
            if (better || (mv != mv && cand[j] == cand[j])) { mv = a; pick = (unsigned int)j; }
        }
        if (mv != mv) {
            if (frirl->verbose > 0) printf("Smallest rulebase found. Exiting.\n");
            break;
        }
        frirl_save_rb_to_bin_file(frirl, "reduction_tmp.frirlrb.bin");
        for (int j = 0; j < K; j++) removed[j] = frb->rseqant[j][pick];
        removed[K] = cand[pick];
        memcpy(cand_prev, cand, sizeof(double) * mr);
        for (int j = (int)pick; j < frb->numofrules - 1; j++) cand[j] = cand[j + 1];
        five_remove_rule(frb, pick);
    }
    free(cand); free(cand_prev); free(removed);
}

/* Knowledge sharing between agents: frirl_omp_run (frirl_agent.c:294-385, the reference's experimental BUILD_OPENMP mode; its
 * default build prints "Built without openmp support!").  World size = OMP_NUM_THREADS (as OpenMP would pick it), else the online
 * CPUs.  Every agent is a full frirl_desc of its own (frirl_init on a copy: own tables, own one-agent device batch), started at
 * gen_def_states' corner-rule states (:120-138); the agents learn their chunks one after the other (the reference's threads are
 * independent between merges, so the order does not matter), merges run on the device (friq_batch_merge_rb). */
static void merge_rule_base(struct frirl_desc *rcvr, int fromid, const struct FIVERB *sndr)              /* merge_rb, :58-117 */
{
    if (rcvr->verbose) printf("Synchronizing the rulebase from #%d to #%d\n", fromid, rcvr->agent_id);
    dropin_engine *e = ENG(rcvr->fiverb);
    sync_params(rcvr);
    ENGINE(friq_batch_merge_rb(e->B, 0, sndr->numofrules, sndr->rant, sndr->rconc));
    pull_rulebase(rcvr->fiverb);
}
void frirl_omp_run(struct frirl_desc *frirl)
{
    const char *nt = getenv("OMP_NUM_THREADS");
    int world = nt ? atoi(nt) : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (world < 1) world = 1;
    if (world == 2) { printf("frirl_omp_run: FATAL ERROR: 2 agents divide by zero in gen_def_states (frirl_agent.c:132): use 1 or >= 3\n"); exit(1); }
    frirl->agent_id = -1; frirl->agent_world_size = world;
    const int S = frirl->statedims_len, K = frirl->numofantecedents;
    struct frirl_desc *agent = (struct frirl_desc *)calloc(world, sizeof *agent);
    for (int id = 0; id < world; id++) {                                                              /* omp_init, :145-263 */
        printf("initializing thread %d\n", id);
        struct frirl_desc *a = &agent[id];
        memcpy(a, frirl, sizeof *a);
        a->agent_id = id; a->agent_world_size = world; a->rbfile = NULL;
        a->statedims = (struct frirl_dimension_desc *)malloc(sizeof(struct frirl_dimension_desc) * S);
        memcpy(a->statedims, frirl->statedims, sizeof(struct frirl_dimension_desc) * S);
        if (id != 0 && world >= 2) {                                                                  /* gen_def_states */
            const struct FIVERB *mf = frirl->fiverb;
            const int gap = mf->numofrules / (world - 2);
            for (int i = 0; i < S; i++) {
                const size_t at = (size_t)(id - 1) * gap * K + i;                                     /* past the corner rules: the buffer's zeros */
                a->statedims[i].values_def = at < (size_t)mf->maxnumofrules * K ? mf->rant[at] : 0.0;
            }
        }
        for (int i = 0; i < S; i++) printf("#%d sdim%d def:%f\n", id, i, a->statedims[i].values_def);
        if (frirl_init(a) != 0) { printf("frirl_omp_run: FATAL ERROR: agent %d could not be initialised\n", id); exit(1); }
        a->agent_id = id; a->agent_world_size = world; a->is_running = 1; a->reduction_state = 0;
        a->episode_num = frirl->episode_num;
        ENGINE(friq_batch_keep_weights(ENG(a->fiverb)->B));
    }
    for (;;) {
        for (int id = 0; id < world; id++) frirl_sequential_run(&agent[id]);
        if (agent[0].is_running == 0) break;
        if (agent[0].epended == 0) {
            for (int id = 1; id < world; id++) merge_rule_base(&agent[id], 0, agent[0].fiverb);
        } else printf("#0 pended, skipping sync to master\n");
        for (int id = 1; id < world; id++) {
            if (agent[id].epended == 0) merge_rule_base(&agent[0], id, agent[id].fiverb);
            else printf("#%d pended, skipping sync\n", id);
        }
    }
    printf("moving rb from thread #0 to main process\n");
    {   /* :370-378: the main process's rule base becomes agent 0's */
        struct FIVERB *frb = frirl->fiverb, *src = agent[0].fiverb;
        while (frb->numofrules > 0) five_remove_rule(frb, frb->numofrules - 1);
        for (int i = 0; i < src->numofrules; i++) FIVE_add_rule(frb, src->rant + (size_t)i * K, src->rconc[i]);
    }
    for (int id = 0; id < world; id++) {
        printf("deinitializing thread %d\n", id);
        struct frirl_dimension_desc *sd = agent[id].statedims;
        frirl_deinit(&agent[id]);
        free(sd);
    }
    free(agent);
}
/* frirl_mpi_run (frirl_agent.c:387-467, the reference's BUILD_MPI mode: one MPI rank per agent).  There is no MPI here: the ranks
 * are agents of THIS process, spread round-robin over the node's GPUs (rank r's one-agent device batch lives on GPU r mod #GPUs),
 * and the rule bases that the reference ships with blocking MPI_Send / MPI_Recv (:427-461) are merged on the receiving agent's
 * GPU (friq_batch_merge_rb).  World size: frirl->agent_world_size when the caller set it (> 1), else the launcher's
 * OMPI_COMM_WORLD_SIZE / PMI_SIZE, else 3 or the number of GPUs, whichever is larger.  Rank 0 is the caller's own frirl_desc, so
 * the merged rule base ends where an MPI rank 0 would hold it.  Exchange order as in the reference: after every chunk of
 * episodes every member's rule base is merged into the leader in rank order (no "pended" test in this mode), then the leader's
 * into every member; the run ends when the leader's rule base is complete (MPI_Bcast of is_running, :421-425). */
void frirl_mpi_run(struct frirl_desc *frirl)
{
    int world = frirl->agent_world_size;
    if (world < 2) { const char *e = getenv("OMPI_COMM_WORLD_SIZE"); if (!e) e = getenv("PMI_SIZE"); world = e ? atoi(e) : 0; }
    int ndev = 1; friq_device_count(&ndev); if (ndev < 1) ndev = 1;
    if (world < 2) world = ndev > 3 ? ndev : 3;
    if (world == 2) { printf("frirl_mpi_run: FATAL ERROR: 2 agents divide by zero in gen_def_states (frirl_agent.c:132): use >= 3\n"); exit(1); }
    const int S = frirl->statedims_len, K = frirl->numofantecedents;
    struct frirl_desc **agent = (struct frirl_desc **)calloc(world, sizeof *agent);
    frirl->agent_id = 0; frirl->agent_world_size = world;
    agent[0] = frirl;
    for (int i = 0; i < S; i++) printf("#%d sdim%d def:%f\n", 0, i, frirl->statedims[i].values_def);
    ENGINE(friq_batch_keep_weights(ENG(frirl->fiverb)->B));
    for (int id = 1; id < world; id++) {
        struct frirl_desc *a = (struct frirl_desc *)malloc(sizeof *a);
        memcpy(a, frirl, sizeof *a);
        a->agent_id = id; a->agent_world_size = world; a->rbfile = NULL;
        a->statedims = (struct frirl_dimension_desc *)malloc(sizeof(struct frirl_dimension_desc) * S);
        memcpy(a->statedims, frirl->statedims, sizeof(struct frirl_dimension_desc) * S);
        {                                                                                             /* gen_def_states, :118-138 */
            const struct FIVERB *mf = frirl->fiverb;
            const int gap = mf->numofrules / (world - 2);
            for (int i = 0; i < S; i++) {
                const size_t at = (size_t)(id - 1) * gap * K + i;
                a->statedims[i].values_def = at < (size_t)mf->maxnumofrules * K ? mf->rant[at] : 0.0;
            }
        }
        for (int i = 0; i < S; i++) printf("#%d sdim%d def:%f\n", id, i, a->statedims[i].values_def);
        g_device_override = id % ndev;
        if (frirl_init(a) != 0) { printf("frirl_mpi_run: FATAL ERROR: agent %d could not be initialised\n", id); exit(1); }
        g_device_override = -1;
        a->agent_id = id; a->agent_world_size = world; a->is_running = 1; a->reduction_state = 0; a->runmode = frirl->runmode;
        a->episode_num = frirl->episode_num;
        ENGINE(friq_batch_keep_weights(ENG(a->fiverb)->B));
        agent[id] = a;
    }
    frirl->is_running = 1;
    for (;;) {
        for (int id = 0; id < world; id++) frirl_sequential_run(agent[id]);
        if (frirl->is_running == 0) { printf("End\n"); break; }                                      /* :421-425 */
        for (int id = 1; id < world; id++) merge_rule_base(frirl, id, agent[id]->fiverb);             /* :427-433 */
        for (int id = 1; id < world; id++) merge_rule_base(agent[id], 0, frirl->fiverb);              /* :445-461 */
    }
    for (int id = 1; id < world; id++) {
        struct frirl_dimension_desc *sd = agent[id]->statedims;
        frirl_deinit(agent[id]);
        free(sd); free(agent[id]);
    }
    free(agent);
}

void frirl_test_run(struct frirl_desc *frirl)                                    /* frirl_test_run.c:20-86 */
{
    if (frirl->reduce_rb) {
        printf("The 'test' mode does not perform reduction. Parameter omitted.\n\n");
        frirl_print_usage();
        exit(-1);
    }
    frirl->construct_rb = 0;
    frirl->reduction_state = 1;                      /* no rule-base updates */
    frirl_episode(frirl);
    if (frirl->verbose != 0)
        printf("Steps:\t%d\nRules:\t%d\nReward:\t%f\n", frirl->reward.ep_total_steps, frirl->fiverb->numofrules, frirl->reward.ep_total_value);
    if (frirl->reward.ep_total_value > frirl->reward_good_above) { printf("\nSuccess!\n"); exit(0); }
    printf("\nInvalid!\n");
    exit(-1);
}

void frirl_run(struct frirl_desc *frirl, int verbose)                            /* frirl_app_helpers.c:50-109 */
{
    (void)verbose;
    if (frirl->rbfile != NULL) {
        if (frirl_load_rb_from_bin_file(frirl, frirl->rbfile) < 0) {
            printf("Error while loading the binary rule-base file: %s!\n", frirl->rbfile);
            exit(-1);
        }
        printf("Loaded %d rules from binary rule-base file: %s\n", frirl->fiverb->numofrules, frirl->rbfile);
        if (frirl->verbose > 1) frirl_show_rb(frirl);
    } else if (frirl->construct_rb == 0 && frirl->reduce_rb == 1) {
        printf("Incrementally constructed rule-base file is missing, \nplease run the construction process first, \nthen supply the constructed rule-base file! \n(example -f example.frirlrb.bin)\n");
        exit(-1);
    }
    switch (frirl->runmode) {
    case FRIRL_SEQ: frirl_sequential_run(frirl); break;
    case FRIRL_OMP: frirl_omp_run(frirl); break;
    case FRIRL_MPI: frirl_mpi_run(frirl); break;
    case FRIRL_TEST: frirl_test_run(frirl); break;
    }
}

void frirl_visualization_init(struct frirl_desc *frirl) { (void)frirl; }         /* built without GLUT */
void frirl_visualization_deinit() {}

/* ------------------------------------------------------------------------------------------
 * rule-base output / files / command line (frirl_utils.c)
 * ------------------------------------------------------------------------------------------ */
void frirl_show_rb(struct frirl_desc *frirl)                                     /* frirl_utils.c:29-56 */
{
    const struct FIVERB *frb = frirl->fiverb;
    const int K = frb->rulelength - 1;
    for (int i = 0; i < frb->numofrules; i++) {
        printf("%d. ", i);
        for (int k = 0; k < K; k++) printf("%.18f ", frb->rant[(size_t)i * K + k]);
        printf("Q: %.18f\n", frb->rconc[i]);
    }
}

void frirl_show_hex_rb(struct frirl_desc *frirl)                                 /* frirl_utils.c:61-93 */
{
    const struct FIVERB *frb = frirl->fiverb;
    const int K = frb->rulelength - 1;
    for (int i = 0; i < frb->numofrules; i++) {
        for (int k = 0; k < K; k++) printf("%.13a ", frb->rant[(size_t)i * K + k]);
        printf("%.13a\n", frb->rconc[i]);
    }
}

int frirl_save_rb_to_text_file(struct frirl_desc *frirl, const char *file_name)  /* frirl_utils.c:100-146 */
{
    const struct FIVERB *frb = frirl->fiverb;
    if (friq_save_rb_to_text_file(file_name, frb->rulelength - 1, frb->numofrules, frb->rant, frb->rconc)) return -1;
    return 0;
}

int frirl_save_rb_to_bin_file(struct frirl_desc *frirl, const char *file_name)   /* frirl_utils.c:151-203 */
{
    const struct FIVERB *frb = frirl->fiverb;
    if (friq_save_rb_to_bin_file(file_name, frirl->numofantecedents, frb->numofrules, frb->rant, frb->rconc)) return -1;
    return 0;
}

int frirl_load_rb_from_bin_file(struct frirl_desc *frirl, const char *file_name) /* frirl_utils.c:228-281 */
{
    struct FIVERB *frb = frirl->fiverb;
    dropin_engine *e = ENG(frb);
    const int K = frb->rulelength - 1;
    int R = 0;
    double *rant = (double *)malloc(sizeof(double) * (size_t)frb->maxnumofrules * K), *rconc = (double *)malloc(sizeof(double) * frb->maxnumofrules);
    if (friq_load_rb_from_bin_file(file_name, K, frb->maxnumofrules, &R, rant, rconc)) { free(rant); free(rconc); return -1; }
    const int zero = 0;
    ENGINE(friq_batch_set_rulebase(e->B, 0, R, rant, rconc));
    ENGINE(friq_batch_set_inserted(e->B, &zero));
    free(rant); free(rconc);
    frirl->fus_is_rule_inserted = 0;                                             /* :263 */
    e->mirrored = 0;
    pull_rulebase(frb);
    return (int)(sizeof(double) * (K + 1));
}

void frirl_print_usage()                                                         /* frirl_utils.c:326-342 */
{
    printf("FRIRL learning usage:\n\n");
    printf("    -m --runmode <runmode>\n");
    printf("\tMode of operation. Possible values: seq, omp, mpi, test. With 'test', a rulebase file should also be specified.\n\tDefault: seq\n\n");
    printf("    -f --rbfile <rule-base file path>\n");
    printf("\tPath to binary rule-base file.\n\n");
    printf("    -r --reductionstrategy <rule-base reduction strategy>\n");
    printf("\tRule-base reduction strategy. Possible values: noreduce, default. Incompatible with 'test' mode.\n\tDefault: noreduce\n\n");
    printf("    -d --draw\n");
    printf("\tEnable visualization.\n\n");
    printf("    -q --quiet\n");
    printf("\tEnable quiet mode.\n\n");
}

void frirl_parse_cmdline(struct frirl_desc *frirl, int argc, char **argv)        /* frirl_utils.c:344-425 */
{
    static const struct option long_opts[] = {
        {"runmode", required_argument, 0, 'm'}, {"rbfile", required_argument, 0, 'f'}, {"reductionstrategy", required_argument, 0, 'r'},
        {"draw", no_argument, 0, 'd'}, {"quiet", no_argument, 0, 'q'}, {0, 0, 0, 0}};
    const char *mode = NULL, *reduce = NULL;
    frirl->argc = argc;
    frirl->argv = argv;
    if (argc == 1) return;
    int opt;
    while ((opt = getopt_long(argc, argv, "m:f:r:dq", long_opts, NULL)) > 0) {
        switch (opt) {
        case 'm': mode = optarg; break;
        case 'f': free(frirl->rbfile); frirl->rbfile = strdup(optarg); break;
        case 'r': reduce = optarg; break;
        case 'd': frirl->visualization = 1; break;
        case 'q': frirl->verbose = 0; break;
        default: frirl_print_usage(); exit(-1);
        }
    }
    if (mode) {
        if (!strcmp(mode, "seq")) frirl->runmode = FRIRL_SEQ;
        else if (!strcmp(mode, "omp")) frirl->runmode = FRIRL_OMP;
        else if (!strcmp(mode, "mpi")) frirl->runmode = FRIRL_MPI;
        else if (!strcmp(mode, "test")) frirl->runmode = FRIRL_TEST;
        else { printf("Invalid runmode!\n\n"); frirl_print_usage(); exit(-1); }
    }
    if (reduce) {
        if (!strcmp(reduce, "noreduce")) { frirl->reduce_rb = 0; frirl->reduction_strategy = FRIRL_REDUCTION_STRATEGY_NOREDUCE; }
        else if (!strcmp(reduce, "default")) { frirl->reduce_rb = 1; frirl->reduction_strategy = FRIRL_REDUCTION_STRATEGY_DEFAULT; }
        else { printf("Invalid reduction strategy!\n\n"); frirl_print_usage(); exit(-1); }
    }
}
