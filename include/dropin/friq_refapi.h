/* include/dropin/friq_refapi.h -- the reference's C API (frirl_* / FIVE_* / five_*), served by the
 * B200 engine.
 *
 * Source-level drop-in for programs written against szaguldo-kamaz/FRI-ReinforcementLearning-C:
 * the example drivers (examples/{mountaincar,acrobot,cartpole}) compile UNCHANGED against the
 * forwarding headers next to this file (frirl.h, frirl_types.h, frirl_types_def.h,
 * frirl_app_helpers.h, frirl_utils.h, frirl_test.h, FIVE.h, config.h) and link against
 * libfriq_b200.so.  Every type, field, macro and prototype the reference's public headers
 * declare is declared here under the same name and with the same meaning:
 *     src/frirl/frirl_types.h:19-174   enums, frirl_dimension_desc, frirl_values_desc,
 *                                      frirl_reward_desc, frirl_desc (+ the four callbacks)
 *     src/frirl/frirl_types_def.h:22-77 frirl_desc_default
 *     src/frirl/frirl.h:22-64          TERM_* colours, frirl_* prototypes
 *     src/frirl/frirl_utils.h:19-25, frirl_app_helpers.h:25-46, frirl_test.h:18
 *     src/five/FIVE.h:19-103           FIVE_MAX_NUM_OF_UNIVERSES, struct FIVERB, FIVE prototypes
 *     config.h.in:17-73                fri_float, DEBUG_MSG, SIMD_ALIGN, MALLOC, likely/unlikely, ...
 * Field ORDER inside the structs is this library's own (programs are recompiled against these
 * headers; nothing depends on the reference's binary layout), and both structs end in one private
 * pointer to the engine's state.
 *
 * What runs where: all FIVE evaluation and SARSA arithmetic (five_rule_distance, FIVE_vag_concl,
 * FIVE_vag_concl_weight, FIVEVagConcl_FRIRL_BestAct, frirl_get_best_action, frirl_update_sarsa) runs in
 * sm_100a kernels for a batch of ONE agent; the host keeps mirrors of the fields callers read directly
 * (numofrules, rant, rconc, rseqant*, weights, ruledists, fgba_actconc), refreshed after every public
 * call that changes them.  The environment callbacks stay host function pointers and are called from
 * frirl_episode exactly as in the reference.  There is no CPU fallback.
 */
#ifndef FRIQ_REFAPI_H
#define FRIQ_REFAPI_H

#include <stdint.h>
#include <stdlib.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- config.h.in, default option set (CMakeLists.txt:9-24) -------------------------------- */
#define FIVE_FIXRES
#define FIVE_NOINF
#define FIVE_NONAN
#define FRIRL_FAST
#define DOUBLE_PRECISION
#define FAST_ABS
#define FAST_POW
#define FAST_SQRT
#define BUILD_CHECK_STATES
#define CHECK_STATES 1
#define FRIRL_AGENT_EPCHUNK 10
typedef double fri_float;
#ifdef DEBUG
#include <stdio.h>
#define DEBUG_MSG(...) fprintf(stderr, __VA_ARGS__);
#else
#define DEBUG_MSG(...) {};
#endif
#define SIMD_ALIGN __attribute__((aligned(32)))
#define MALLOC(size) malloc(size)
#define likely(x) x
#define unlikely(x) x

/* ---- frirl.h:22-24 ------------------------------------------------------------------------ */
#define TERM_RED   "\033[0;31m"
#define TERM_GREEN "\033[0;32m"
#define TERM_NC    "\033[0m"

/* ---- FIVE.h ------------------------------------------------------------------------------- */
#define FIVE_MAX_NUM_OF_UNIVERSES 8

struct FIVERB {
    /* problem tables (borrowed from the caller of FIVEInit, FIVEInit.c:78-79) */
    int numofunivs, univlength;
    double *u, *ve;
    double **uk, **vek;                 /* per-dimension row pointers into u / ve */
    unsigned int uksize;                /* univlength - 1 */
    double *ukdomains, *udivs;
    double *psc, *scf;                  /* never set by the reference either */
    int p;                              /* Shepard power */
    /* rule base: host mirrors of the device rule base of the one agent behind this struct */
    int numofrules, maxnumofrules, rulelength, numofantecedents;
    double *rant, *rconc;               /* canonical AoS antecedents [R][K] and consequents [R] (borrowed) */
    double *newrant, *newrconc;         /* append cursors */
    double **rseqant;                   /* [K][R] antecedent values per dimension */
    unsigned int **rseqant_uindex;      /* [K][R] nearest universe index */
    double **rseqant_veval;             /* [K][R] VE value at that index */
    unsigned int *rant_uindex, *rant_veval;
    double *ract, *ract_veval;          /* action rows of rseqant / rseqant_veval */
    unsigned int *ract_uindex;
    /* results of the last call */
    double *ruledists;                  /* five_rule_distance */
    double *weights, *wi;               /* FIVE_vag_concl_weight */
    double *frd_dists, *fvc_vagdist;    /* scratch of the reference's AVX2 paths: allocated, unused here */
    /* FIVEValVag */
    double *valvagp, *valvagu, *valvagve;
    double valvagdims;
    unsigned int epno;
    unsigned int avx2_rbsize;           /* ceil(numofrules / 4), kept for programs that read it */
    void *friq_engine;                  /* private: the device batch behind this rule base */
};

struct FIVERB *FIVEInit(double *u, double *ve, int p, int numofunivs, int univlength, int numofrules, int maxnumofrules,
                        int rulelength, double *rant, double *rconc);
void five_deinit(struct FIVERB *frb);
/* init-time table builders (host, plain sequential FP64) */
int FIVE_GSc_func(double *u, int numofunivs, int univlength, double *psc, int mp, int np, double nls, double *scf);
double *FIVEGScFunc(double *u, int numofunivs, int univlength, double *psc, int mp, int np, double nls);
double *FIVEGVagEnv(double *u, int numofunivs, int univlength, double *scf);
int FIVEValVag(struct FIVERB *frb, double *vp);
int five_vague_distance(struct FIVERB *frb, fri_float *p1, fri_float *p2, fri_float *d);
int five_vague_distance_parallel(struct FIVERB *frb, fri_float *p1, int p1_offset, fri_float *p2, fri_float *d);
/* device-backed evaluation */
int five_rule_distance(struct FIVERB *frb, fri_float *x);
unsigned int FIVE_vag_concl(struct FIVERB *frb, double *ant, double *conc);
unsigned int FIVE_vag_concl_weight(struct FIVERB *frb, double *ant, double *weights);
double FIVEVagConcl(struct FIVERB *frb, double *x);
unsigned int FIVEVagConclWeight(struct FIVERB *frb, double *x);
double FIVEVagConcl_FRIRL_BestAct(struct FIVERB *frb, double *ruledists);
/* rule-base edits (host mirror + device) */
int FIVE_add_rule(struct FIVERB *frb, fri_float *rant, fri_float rconc);
int FIVEAddRule(struct FIVERB *frb, double *newrule);
int five_add_rule(struct FIVERB *frb, fri_float *ruletoadd);
int five_remove_rule(struct FIVERB *frb, unsigned int rulenotoremove);

/* ---- frirl_types.h ------------------------------------------------------------------------ */
enum frirl_runmode { FRIRL_SEQ, FRIRL_OMP, FRIRL_MPI, FRIRL_TEST };
enum frirl_reduction_strategy { FRIRL_REDUCTION_STRATEGY_NOREDUCE, FRIRL_REDUCTION_STRATEGY_DEFAULT };

struct frirl_dimension_desc {
    int values_len;
    fri_float *values;
    fri_float values_div, values_steep, values_def;
    int universe_len;
    fri_float *universe;
    fri_float universe_div;
};

struct frirl_values_desc {
    int values_len;
    fri_float *values;
    fri_float *vevalues;                /* VE values of the possible actions (frirl_init.c:156-158) */
    fri_float epsilon;
};

struct frirl_reward_desc {
    fri_float value, ep_total_value;
    int ep_total_steps, success;
};

struct frirl_desc {
    /* command line / run control */
    int argc;
    char **argv;
    int runmode;
    char *rbfile;
    int visualization, gui_width, gui_height, verbose, agent_rnd_init;
    /* problem */
    int statedims_len;
    struct frirl_dimension_desc *statedims;
    struct frirl_dimension_desc actiondim;
    /* learning parameters */
    fri_float alpha, gamma, epsilon;
    fri_float qdiff_pos_boundary, qdiff_neg_boundary, qdiff_final_tolerance;
    fri_float reward_good_above;
    fri_float rule_weight_considered_significant_for_update;
    fri_float reduction_reward_tolerance;
    unsigned char skip_rules, no_random, construct_rb, reduce_rb, reduction_strategy;
    int max_episodes, max_steps;
    unsigned int episode_num;
    int five_maxnumofrules;
    /* environment callbacks (host) */
    void (*get_reward_func)(struct frirl_desc *frirl, fri_float *states, int states_len, struct frirl_reward_desc *reward);
    void (*do_action_func)(struct frirl_desc *frirl, fri_float action, fri_float *states, int states_len, fri_float *new_states);
    void (*quant_obs_func)(struct frirl_desc *frirl, fri_float *states, int states_len, fri_float *new_states);
    void (*draw_func)(struct frirl_desc *frirl, fri_float *new_curr_state, double action, unsigned int steps);
    /* "private" members of the reference, kept because callers read them */
    struct FIVERB *fiverb;
    double *fiverb_ua, *fiverb_vea;
    unsigned int reduction_state;
    int numofantecedents;
    struct frirl_values_desc *possible_states, *possible_actions;
    struct frirl_reward_desc reward;
    fri_float *fgba_vagdist_states, *fgba_ruledist, *fgba_actconc, *fgba_dists, *fgba_statedistsum;
    fri_float *fus_proposed_values, *fus_values, *fus_check_states;
    fri_float fus_is_rule_inserted;
    fri_float *fep_ant, *fep_cur_ant, *fep_q_ant, *fep_cur_q_ant;
    int is_running, agent_id, agent_world_size, epended;
    unsigned int keyaction;
    int valid_simulation, original_learning, user_exited;
};
#define FRIRL frirl_desc

/* ---- frirl.h ------------------------------------------------------------------------------ */
int frirl_init(struct frirl_desc *frirl);
void frirl_deinit(struct frirl_desc *frirl);
int frirl_init_ve(struct frirl_desc *frirl, fri_float *ve, fri_float *u, int univlength);
int frirl_init_rb(struct frirl_desc *frirl, fri_float *rant, fri_float *rconc, int *numofrules);
void frirl_episode(struct frirl_desc *frirl);
unsigned int frirl_e_greedy_selection(struct frirl_desc *frirl, fri_float *states);
unsigned int frirl_get_best_action(struct frirl_desc *frirl, fri_float *states);
fri_float frirl_check_possible_states(struct frirl_desc *frirl, fri_float observation, struct frirl_values_desc *possible_states);
void frirl_update_sarsa(struct frirl_desc *frirl, fri_float *q_ant, fri_float reward, fri_float *cur_q_ant);
void frirl_sequential_run(struct frirl_desc *frirl);
void frirl_omp_run(struct frirl_desc *frirl);
void frirl_mpi_run(struct frirl_desc *frirl);

/* ---- frirl_utils.h, frirl_app_helpers.h, frirl_test.h --------------------------------------- */
void frirl_show_rb(struct frirl_desc *frirl);
void frirl_show_hex_rb(struct frirl_desc *frirl);
int frirl_save_rb_to_text_file(struct frirl_desc *frirl, const char *file_name);
int frirl_save_rb_to_bin_file(struct frirl_desc *frirl, const char *file_name);
int frirl_load_rb_from_bin_file(struct frirl_desc *frirl, const char *file_name);
void frirl_print_usage();
void frirl_parse_cmdline(struct frirl_desc *frirl, int argc, char **argv);
void frirl_gen_fixres_arr(fri_float *arr, int len, fri_float div);
void frirl_run(struct frirl_desc *frirl, int verbose);
void frirl_visualization_init(struct frirl_desc *frirl);
void frirl_visualization_deinit();
void frirl_test_run(struct frirl_desc *frirl);

#define FRIRL_ALLOC_VALUES(x)   fri_float _local_ ## x ## _values[x.values_len];     x.values   = _local_ ## x ## _values;
#define FRIRL_ALLOC_UNIVERSE(x) fri_float _local_ ## x ## _universe[x.universe_len]; x.universe = _local_ ## x ## _universe;
#define FRIRL_ALLOC_DIM(x)  FRIRL_ALLOC_VALUES(x);  FRIRL_ALLOC_UNIVERSE(x);
#define FRIRL_GEN_FIXRES_UNIVERSE(x) frirl_gen_fixres_arr(x.universe, x.universe_len, x.universe_div);
#define FRIRL_GEN_FIXRES_VALUES(x)   frirl_gen_fixres_arr(x.values,   x.values_len,   x.values_div);
#define FRIRL_GEN_FIXRES_DIM(x)  FRIRL_GEN_FIXRES_VALUES(x);  FRIRL_GEN_FIXRES_UNIVERSE(x);

#ifdef __cplusplus
}
#endif
#endif /* FRIQ_REFAPI_H */
