/* tests/dropin/api_probe.c -- a command interpreter over the reference's C API (include/dropin), used by
 * tests/test_dropin_gpu.py to compare the B200 engine's frirl_* / FIVE_* entry points with the oracle call by
 * call.  Reads commands from stdin, prints results as hex floats.  Own code; no reference sources involved. */
#include "frirl.h"
#include "frirl_types_def.h"
#include "frirl_app_helpers.h"
#include "friq_b200.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static void no_reward(struct frirl_desc *f, fri_float *s, int n, struct frirl_reward_desc *r) { (void)f; (void)s; (void)n; r->value = 0; r->success = 1; }
static void no_action(struct frirl_desc *f, fri_float a, fri_float *s, int n, fri_float *o) { (void)f; (void)a; memcpy(o, s, sizeof(double) * n); }
static void no_quant(struct frirl_desc *f, fri_float *s, int n, fri_float *o) { (void)f; memcpy(o, s, sizeof(double) * n); }

static void read_doubles(double *v, int n) { for (int i = 0; i < n; i++) if (scanf("%la", &v[i]) != 1) { printf("parse error\n"); exit(2); } }

int main(void)
{
    struct frirl_desc f = frirl_desc_default;
    struct frirl_dimension_desc dims[8];
    char cmd[64];
    int K = 0, S = 0, A = 0;
    double v[64];
    while (scanf("%63s", cmd) == 1) {
        if (!strcmp(cmd, "init")) {                          /* init <env id>: parameter block of the example driver */
            int env; friq_problem_desc pd;
            if (scanf("%d", &env) != 1 || friq_problem_desc_builtin(env, &pd)) return 2;
            S = pd.statedims_len; K = S + 1; A = pd.actiondim.values_len;
            f.alpha = pd.alpha; f.gamma = pd.gamma; f.epsilon = pd.epsilon;
            f.qdiff_pos_boundary = pd.qdiff_pos_boundary; f.qdiff_neg_boundary = pd.qdiff_neg_boundary; f.qdiff_final_tolerance = pd.qdiff_final_tolerance;
            f.reward_good_above = pd.reward_good_above; f.skip_rules = (unsigned char)pd.skip_rules;
            f.get_reward_func = no_reward; f.do_action_func = no_action; f.quant_obs_func = no_quant;
            for (int k = 0; k <= S; k++) {
                const friq_dim_desc *d = (k < S) ? &pd.statedims[k] : &pd.actiondim;
                struct frirl_dimension_desc *o = &dims[k];
                o->values_len = d->values_len; o->values_div = d->values_div; o->values_steep = d->values_steep; o->values_def = d->values_def;
                o->universe_len = d->universe_len; o->universe_div = d->universe_div;
                o->values = (double *)malloc(sizeof(double) * d->values_len); o->universe = (double *)malloc(sizeof(double) * d->universe_len);
                if (d->values) memcpy(o->values, d->values, sizeof(double) * d->values_len); else frirl_gen_fixres_arr(o->values, o->values_len, o->values_div);
                frirl_gen_fixres_arr(o->universe, o->universe_len, o->universe_div);
            }
            f.statedims = dims; f.statedims_len = S; f.actiondim = dims[S];
            f.five_maxnumofrules = 4096;
            if (frirl_init(&f)) return 3;
            printf("init %d %d %d %d\n", S, A, f.fiverb->univlength, f.fiverb->numofrules);
        } else if (!strcmp(cmd, "load")) {
            char path[512];
            if (scanf("%511s", path) != 1 || frirl_load_rb_from_bin_file(&f, path) < 0) return 4;
            printf("load %d\n", f.fiverb->numofrules);
        } else if (!strcmp(cmd, "concl")) {
            double c = 0; read_doubles(v, K);
            const unsigned int hit = FIVE_vag_concl(f.fiverb, v, &c);
            printf("concl %d %a %a\n", (int)hit, c, FIVEVagConcl(f.fiverb, v));
        } else if (!strcmp(cmd, "weight")) {
            read_doubles(v, K);
            const unsigned int hit = FIVEVagConclWeight(f.fiverb, v);
            printf("weight %d", (int)hit);
            if (hit == ~0u) for (int i = 0; i < f.fiverb->numofrules; i++) printf(" %a", f.fiverb->weights[i]);
            printf("\n");
        } else if (!strcmp(cmd, "dist")) {
            read_doubles(v, K);
            const int hit = five_rule_distance(f.fiverb, v);
            printf("dist %d", hit);
            for (int i = 0; i < f.fiverb->numofrules; i++) printf(" %a", f.fiverb->ruledists[i]);
            printf("\n");
        } else if (!strcmp(cmd, "bestact")) {                /* conclusion from the distances of the last `dist` */
            printf("bestact %a\n", FIVEVagConcl_FRIRL_BestAct(f.fiverb, f.fiverb->ruledists));
        } else if (!strcmp(cmd, "best")) {
            read_doubles(v, S);
            const unsigned int b = frirl_get_best_action(&f, v);
            printf("best %u", b);
            for (int a = 0; a < A; a++) printf(" %a", f.fgba_actconc[a]);
            printf("\n");
        } else if (!strcmp(cmd, "sarsa")) {                  /* q_ant[K] reward cur_q_ant[K] */
            read_doubles(v, 2 * K + 1);
            double cur[8]; memcpy(cur, v + K + 1, sizeof(double) * K);
            frirl_update_sarsa(&f, v, v[K], cur);
            printf("sarsa %d %d\n", f.fiverb->numofrules, (int)f.fus_is_rule_inserted);
        } else if (!strcmp(cmd, "add")) {
            read_doubles(v, K + 1);
            five_add_rule(f.fiverb, v);
            printf("add %d\n", f.fiverb->numofrules);
        } else if (!strcmp(cmd, "remove")) {
            int idx; if (scanf("%d", &idx) != 1) return 2;
            five_remove_rule(f.fiverb, (unsigned int)idx);
            printf("remove %d\n", f.fiverb->numofrules);
        } else if (!strcmp(cmd, "snap")) {
            int k; if (scanf("%d", &k) != 1) return 2;
            read_doubles(v, 1);
            printf("snap %a\n", frirl_check_possible_states(&f, v[0], &f.possible_states[k]));
        } else if (!strcmp(cmd, "rb")) {
            const struct FIVERB *frb = f.fiverb;
            printf("rb %d\n", frb->numofrules);
            for (int i = 0; i < frb->numofrules; i++) {
                for (int k = 0; k < K; k++) printf("%a %u %a ", frb->rant[i * K + k], frb->rseqant_uindex[k][i], frb->rseqant_veval[k][i]);
                printf("%a\n", frb->rconc[i]);
            }
        } else if (!strcmp(cmd, "tables")) {
            const struct FIVERB *frb = f.fiverb;
            double *ve2 = (double *)malloc(sizeof(double) * K * frb->univlength);
            frirl_init_ve(&f, ve2, frb->u, frb->univlength);       /* the public builder must agree with the engine's tables */
            printf("tables %d", memcmp(ve2, frb->ve, sizeof(double) * K * frb->univlength) == 0);
            for (int i = 0; i < K * frb->univlength; i++) printf(" %a %a", frb->u[i], frb->ve[i]);
            for (int a = 0; a < A; a++) printf(" %a", f.possible_actions->vevalues[a]);
            printf("\n");
            free(ve2);
        } else { printf("unknown command %s\n", cmd); return 2; }
        fflush(stdout);
    }
    if (f.fiverb) frirl_deinit(&f);
    return 0;
}
