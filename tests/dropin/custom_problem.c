/* tests/dropin/custom_problem.c -- a program written ONLY against the reference's public C API (frirl.h, frirl_types_def.h,
 * frirl_app_helpers.h, FIVE.h), for a problem that is none of the three examples: 3 state dimensions, 5 actions,
 * skip_rules = 0 (the default for new problems), a made-up deterministic environment without libm calls.
 * It is compiled twice from this one source:
 *   oracle/_ref/custom_ref          against the reference's headers and library   -> tests/golden/custom_problem.out
 *   tests/dropin/_build/custom_dropin  against include/dropin and libfriq_b200.so -> must print the same, bit for bit
 * (all numbers are printed as hex floats).  Own code; nothing is taken from the reference's sources.
 * -DNS=5 / -DNS=7 add two / four more state dimensions (K = 6 / 8 antecedents, the reference's FIVE_MAX_NUM_OF_UNIVERSES):
 * custom6_* / custom8_* are the same program for those sizes. */
#include "frirl.h"
#include "frirl_types.h"
#include "frirl_types_def.h"
#include "frirl_app_helpers.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifndef NS
#define NS 3                      /* state dimensions */
#endif
#define NK (NS + 1)

static void do_action(struct frirl_desc *f, fri_float action, fri_float *s, int n, fri_float *o)
{   /* a damped point mass pushed by the action, with a slowly drifting third coordinate */
    (void)f; (void)n;
    double v = s[1] * 0.95 + action * 0.07 - s[0] * 0.02;
    if (v > 0.5) v = 0.5;
    if (v < -0.5) v = -0.5;
    double x = s[0] + v;
    if (x > 2.0) { x = 2.0; v = 0.0; }
    if (x < -2.0) { x = -2.0; v = 0.0; }
    double w = s[2] * 0.9 + (x > 0 ? 0.05 : -0.05);
    o[0] = x; o[1] = v; o[2] = w;
    for (int k = 3; k < NS; k++) o[k] = s[k] * 0.8 + ((k & 1) ? 0.06 : -0.04) * (v > 0 ? 1.0 : -1.0) + 0.01 * s[k - 1];   /* more slow coordinates */
}
static void get_reward(struct frirl_desc *f, fri_float *s, int n, struct frirl_reward_desc *r)
{
    (void)f; (void)n;
    const double ax = s[0] < 0 ? -s[0] : s[0];
    if (s[0] > 1.6 && s[2] > 0.2) { r->value = 500.0; r->success = 1; }
    else { r->value = -1.0 - ax; r->success = 0; }
}
static void quantize(struct frirl_desc *f, fri_float *s, int n, fri_float *o)
{
    for (int i = 0; i < n; i++) o[i] = frirl_check_possible_states(f, s[i], &f->possible_states[i]);
}

int main(void)
{
    struct frirl_desc frirl = frirl_desc_default;
    frirl.alpha = 0.4; frirl.gamma = 0.95; frirl.epsilon = 0.0;
    frirl.qdiff_pos_boundary = 2.0; frirl.qdiff_neg_boundary = -3.0; frirl.qdiff_final_tolerance = 5.0;
    frirl.reward_good_above = 100.0;
    frirl.max_episodes = 25; frirl.max_steps = 150; frirl.five_maxnumofrules = 2048;
    frirl.do_action_func = do_action; frirl.get_reward_func = get_reward; frirl.quant_obs_func = quantize;
    struct frirl_dimension_desc d0 = { .values_len = 5, .values_div = 1.0, .values_steep = 1.0, .values_def = -1.2, .universe_len = 41, .universe_div = 0.1 };
    FRIRL_ALLOC_DIM(d0); FRIRL_GEN_FIXRES_DIM(d0);
    struct frirl_dimension_desc d1 = { .values_len = 5, .values_div = 0.25, .values_steep = 2.0, .values_def = 0.0, .universe_len = 41, .universe_div = 0.025 };
    FRIRL_ALLOC_DIM(d1); FRIRL_GEN_FIXRES_DIM(d1);
    struct frirl_dimension_desc d2 = { .values_len = 3, .values_div = 0.5, .values_steep = 1.0, .values_def = 0.0, .universe_len = 41, .universe_div = 0.025 };
    FRIRL_ALLOC_DIM(d2); FRIRL_GEN_FIXRES_DIM(d2);
    struct frirl_dimension_desc dims[NS] = {d0, d1, d2};
    for (int k = 3; k < NS; k++) {
        struct frirl_dimension_desc dk = { .values_len = 3, .values_div = 0.5, .values_steep = 1.0, .values_def = 0.0, .universe_len = 41, .universe_div = 0.025 };
        FRIRL_ALLOC_DIM(dk); FRIRL_GEN_FIXRES_DIM(dk);
        dims[k] = dk;
    }
    frirl.statedims = dims;
    frirl.statedims_len = NS;
    struct frirl_dimension_desc actiondim = { .values_len = 5, .values_div = 0.5, .universe_len = 41, .universe_div = 0.05 };
    FRIRL_ALLOC_DIM(actiondim); FRIRL_GEN_FIXRES_DIM(actiondim);
    frirl.actiondim = actiondim;
    if (frirl_init(&frirl)) return 1;
    struct FIVERB *frb = frirl.fiverb;
    printf("init rules %d K %d U %d p %d\n", frb->numofrules, frb->numofunivs, frb->univlength, frb->p);
    for (int a = 0; a < 5; a++) printf("act %a ve %a\n", frirl.possible_actions->values[a], frirl.possible_actions->vevalues[a]);
    frirl_sequential_run(&frirl);
    printf("final rules %d\n", frb->numofrules);
    for (int i = 0; i < frb->numofrules; i++) {
        for (int k = 0; k < NK; k++) printf("%a ", frb->rant[i * NK + k]);
        printf("%a\n", frb->rconc[i]);
    }
    /* per-call functions at a few points (pseudo-random, fixed) */
    unsigned long long lcg = 12345;
    for (int t = 0; t < 12; t++) {
        double x[NK];
        for (int k = 0; k < NK; k++) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            const double u01 = (double)(lcg >> 11) * 0x1p-53;
            const double lo = frb->u[k * 41], hi = frb->u[k * 41 + 40];
            x[k] = lo + u01 * (hi - lo);
        }
        if (t % 4 == 3) memcpy(x, frb->rant + (size_t)(t % frb->numofrules) * NK, sizeof x);     /* an exact hit */
        double conc = 0.0;
        const unsigned hit = FIVE_vag_concl(frb, x, &conc);
        const unsigned best = frirl_get_best_action(&frirl, x);
        printf("concl %d %a best %u", (int)hit, conc, best);
        for (int a = 0; a < 5; a++) printf(" %a", frirl.fgba_actconc[a]);
        const unsigned hw = FIVE_vag_concl_weight(frb, x, frb->weights);
        printf(" w %d", (int)hw);
        if (hw == ~0u) printf(" %a %a", frb->weights[0], frb->weights[frb->numofrules - 1]);
        printf("\n");
    }
    frirl_save_rb_to_text_file(&frirl, "custom.frirlrb.txt");
    frirl_deinit(&frirl);
    return 0;
}
