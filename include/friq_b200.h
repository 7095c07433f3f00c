/* include/friq_b200.h -- C-ABI of the B200-native FRIQ-learning engine.
 *
 * Plain C: opaque handles, plain pointers and sizes, no CUDA / torch types.  The
 * library behind it (libfriq_b200.so) is ANSI-C host code on top of hand-written
 * sm_100a kernels; there is NO CPU fallback: every function that evaluates or
 * updates a rule base launches a kernel and fails with a non-zero code when no
 * CUDA device is usable.
 *
 * What it replaces in the reference (szaguldo-kamaz/FRI-ReinforcementLearning-C):
 * the hot path  frirl_sequential_run -> frirl_episode -> { env callbacks,
 * frirl_e_greedy_selection -> frirl_get_best_action -> FIVEVagConcl_FRIRL_BestAct,
 * frirl_update_sarsa -> FIVE_vag_concl / FIVE_vag_concl_weight / FIVE_add_rule ->
 * five_rule_distance }  (SURVEY.md section 3), for a BATCH of independent agents
 * instead of one.  Every entry point cites the reference interface it stands in for.
 *
 * Return value: 0 on success, non-zero on error; friq_last_error() describes it.
 */
#ifndef FRIQ_B200_H
#define FRIQ_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRIQ_ENV_HOST_CALLBACKS (-1)   /* no built-in environment: per-call API only */
#define FRIQ_ENV_MOUNTAINCAR 0         /* examples/mountaincar/mountaincar.c:37-125 */
#define FRIQ_ENV_ACROBOT 1             /* examples/acrobot/acrobot.c:31-192         */
#define FRIQ_ENV_CARTPOLE 2            /* examples/cartpole/cartpole.c:35-168       */

/* run-mode bits of friq_batch_run() */
#define FRIQ_RUN_LEARN 1               /* reduction_state == 0 (frirl_episode.c:155-160); 0 = frirl_test_run.c:65-69 */
#define FRIQ_RUN_TRACK_CONVERGENCE 2   /* evaluate the "RB complete" predicate, frirl_sequential_run.c:83-150 */
#define FRIQ_RUN_STOP_ON_CONVERGE 4    /* converged agents stop stepping */
#define FRIQ_RUN_STOP_AT_EPISODE_END 8 /* every agent stops after finishing its current episode (frirl_episode granularity) */
#define FRIQ_RUN_RETILE 16             /* before the launch, re-assign agents to warps by rule count (results unaffected) */
#define FRIQ_RUN_KEEP_WEIGHTS 32       /* maintain every agent's FIVERB.weights (FIVE.h:24-76) exactly as the reference leaves it after
                                        * update_rules (frirl_update_sarsa.c:39-40): merge_rb reads it stale.  Not with FRIQ_RUN_RETILE. */

/* One dimension of the problem; field for field struct frirl_dimension_desc
 * (src/frirl/frirl_types.h:34-44).  values / universe may be NULL, in which case
 * they are generated like FRIRL_GEN_FIXRES_VALUES / _UNIVERSE do
 * (frirl_app_helpers.h:25-38, frirl_app_helpers.c:32-44). */
typedef struct friq_dim_desc {
    int values_len;
    const double *values;
    double values_div, values_steep, values_def;
    int universe_len;
    const double *universe;
    double universe_div;
} friq_dim_desc;

/* The part of struct frirl_desc (frirl_types.h:62-169) the path reads. */
typedef struct friq_problem_desc {
    int env;                               /* FRIQ_ENV_* (stands in for the 3 callbacks, frirl_types.h:103-120) */
    int statedims_len;
    const friq_dim_desc *statedims;
    friq_dim_desc actiondim;
    double alpha, gamma, epsilon;
    double qdiff_pos_boundary, qdiff_neg_boundary, qdiff_final_tolerance;
    double reward_good_above, rule_weight_considered_significant_for_update;
    int skip_rules, no_random, max_episodes, max_steps;
} friq_problem_desc;

typedef struct friq_problem friq_problem;
typedef struct friq_batch friq_batch;

const char *friq_last_error(void);
int friq_device_count(int *count);

/* ---- problem (host tables: frirl_init.c:29-341, frirl_init_ve.c, FIVEGScFunc.c, FIVEGVagEnv.c, FIVEInit.c) */
/* parameter block of one of the three example drivers (mountaincar.c:220-279, acrobot.c:242-324,
 * cartpole.c:252-402); storage is static, valid for the life of the process */
int friq_problem_desc_builtin(int env, friq_problem_desc *out);
int friq_problem_create(const friq_problem_desc *desc, friq_problem **out);
/* FIVEInit (FIVE.h:79, FIVEInit.c:55): a problem from caller-supplied tables u, ve [K][U]; p = 0 means p = K.
 * No grids / actions / environment: only the FIVE-level per-call entry points apply. */
int friq_problem_create_from_tables(int K, int U, int p, const double *u, const double *ve, friq_problem **out);
/* public learning fields of struct frirl_desc a caller may change between calls (frirl_types.h:84-92) */
int friq_problem_set_learning_params(friq_problem *p, double qdiff_pos, double qdiff_neg, double wsig, int skip_rules);
void friq_problem_destroy(friq_problem *p);
/* dims[0..5] = S, K, A, U, p, 2^K */
int friq_problem_dims(const friq_problem *p, int *dims);
/* host copies of the tables: u, ve [K][U]; udiv [K]; act_val, act_ve [A] (any pointer may be NULL) */
int friq_problem_tables(const friq_problem *p, double *u, double *ve, double *udiv, double *act_val, double *act_ve);

/* ---- batch of independent agents, each with a private rule base (replaces one struct frirl_desc +
 * struct FIVERB per agent).  Agent i uses the Philox key seed0 + first_agent + i. */
int friq_batch_create(friq_problem *p, int device, long n_agents, int rule_capacity, uint64_t seed0, long first_agent, friq_batch **out);
void friq_batch_destroy(friq_batch *b);
/* back to frirl_init state: corner rules (frirl_init_rb.c:99-126), episode 1 */
int friq_batch_reset(friq_batch *b);
/* per-agent hyper-parameters (host arrays of n_agents, NULL keeps the current values) */
int friq_batch_set_hyper(friq_batch *b, const double *alpha, const double *gamma, const double *epsilon);
/* rule base of agent `agent` (or of ALL agents when agent < 0): rant [R][K], rconc [R]
 * (frirl_load_rb_from_bin_file -> five_add_rule, frirl_utils.c:253-276) */
int friq_batch_set_rulebase(friq_batch *b, long agent, int R, const double *rant, const double *rconc);
int friq_batch_get_rulebase(friq_batch *b, long agent, int *R, double *rant, double *rconc);
/* ONE read-only rule base shared by all agents (inference only: reduction_state = 1, frirl_test_run.c:65-69);
 * R = 0 returns to the private rule bases.  Default evaluation: the conclusions are memoised per (observation grid cell,
 * action) once, plus one pass over the rule base (streamed through shared memory by TMA) per agent for its off-grid episode
 * start state; friq_batch_set_policy(b, 0) runs the full pass on every env step instead (same results) */
int friq_batch_set_shared_rulebase(friq_batch *b, int R, const double *rant, const double *rconc);
/* per-agent episode start state s0 [n][S] instead of values_def (frirl_episode.c:46-48); NULL restores it */
int friq_batch_set_initial_state(friq_batch *b, const double *s0);
/* the same from a device array [n][S] (a cudaMalloc'ed buffer / torch tensor already in HBM), asynchronous on `stream` */
int friq_batch_set_initial_state_device(friq_batch *b, const void *d_s0, void *stream);
/* current raw state s [n][S] (fep_ant) and action index [n] of every agent */
int friq_batch_get_state(friq_batch *b, double *s, int *action_index);
int friq_batch_set_inserted(friq_batch *b, const int *inserted);   /* fus_is_rule_inserted, frirl_types.h:153 */
int friq_batch_get_inserted(friq_batch *b, int *inserted);

/* ---- the fused hot path: every agent advances n_steps env steps (frirl_episode.c:86-186 per step,
 * episodes roll over as in frirl_sequential_run.c:57-161).  Asynchronous on `stream`
 * (a cudaStream_t passed as void*, NULL = default stream). */
int friq_batch_run(friq_batch *b, long n_steps, int mode, void *stream);
int friq_batch_sync(friq_batch *b, void *stream);
/* evaluation policy of the fused kernel: 0 lock-step pass over all actions, 1 exact hits first + compacted Shepard chains,
 * 2 the same with memoised weights (table mode; the default where the problem qualifies, else the next lower one).
 * Evaluation strategy only: results are bit-identical under
 * every policy.  With a shared rule base: 0 = one full pass per env step, else memoised conclusions. */
int friq_batch_set_policy(friq_batch *b, int policy);
/* agent groups of a re-tiled run (each sorted by rule count and launched on its own stream); 0 = default */
int friq_batch_set_groups(friq_batch *b, int groups);
/* with FRIQ_RUN_RETILE a run is split into launches of `steps` env steps with a re-tiling before each (default 125; 0 = once per run) */
int friq_batch_set_retile_interval(friq_batch *b, long steps);
/* results -> host arrays of n_agents (any pointer may be NULL): last finished episode's total reward /
 * steps / rule count (the per-episode printf of frirl_sequential_run.c:77-80), episodes finished,
 * converged / overflow flags, env steps executed, and the sum over steps of the rule-base size */
int friq_batch_gather(friq_batch *b, double *last_return, int *last_steps, int *rules, int *episodes, int *flags,
                      long long *total_steps, long long *sum_rules);
/* every agent starts a NEW episode at its next step, whatever state its current one is in (a fresh batch of start states for
 * inference; rule bases, counters and random streams untouched); asynchronous on `stream` */
int friq_batch_new_episodes(friq_batch *b, void *stream);
/* statistics of a gathered result set (device arrays of n agents): d_out3 = {sum of rules, sum of returns, max rules} (f64) */
int friq_result_stats(int device, const void *d_ret_f64, const void *d_rules_i32, long n, void *d_out3_f64, void *stream);
/* executed-work counters of the fused kernels, accumulated on the device since batch creation / the last reset: out[8] = (agent,
 * channel) pairs that needed their Shepard sums, rules summed with looked-up weights (table mode), rules summed with computed
 * sqrt / x87-power / reciprocal chains, pairs answered by an exact-hit rule, rules appended, rules scanned by weighted updates,
 * env steps, 0.  What the path EXECUTES, beside the reference's nominal operation count (bench.py: roofline.executed) */
int friq_batch_work_counters(friq_batch *b, unsigned long long *out, int reset);
/* same results copied device -> device into caller-owned buffers (e.g. torch tensors fed to NCCL) */
int friq_batch_gather_device(friq_batch *b, void *d_last_return_f64, void *d_last_steps_i32, void *d_rules_i32, void *stream);
/* end-to-end convenience: hyper-parameters H2D, run, results D2H, all on `stream` and synchronised */
int friq_batch_run_host(friq_batch *b, const double *alpha, const double *gamma, const double *epsilon, long n_steps, int mode,
                        double *last_return, int *last_steps, int *rules, void *stream);
/* the same, enqueued only (no synchronisation): the host buffers are valid once `stream` has passed this point */
int friq_batch_run_host_async(friq_batch *b, const double *alpha, const double *gamma, const double *epsilon, long n_steps, int mode,
                              double *last_return, int *last_steps, int *rules, void *stream);

/* ---- per-call entry points: one call of the reference function per agent, host buffers in and out */
/* FIVE_vag_concl (FIVE.h:102, FIVEVagConcl.c:64): ant [n][K] -> conc [n], hit [n] (rule index or -1) */
int friq_batch_vag_concl(friq_batch *b, const double *ant, double *conc, int *hit);
/* FIVE_vag_concl_weight (FIVE.h:101, FIVEVagConclWeight.c:52): weights [n][rule_capacity], filled when hit == -1 */
int friq_batch_vag_concl_weight(friq_batch *b, const double *ant, int *hit, double *weights);
/* frirl_get_best_action (frirl.h:54, frirl_get_best_action.c:31): states [n][S] -> actconc [n][A], best [n] */
int friq_batch_get_best_action(friq_batch *b, const double *states, double *actconc, int *best);
/* frirl_update_sarsa (frirl.h:64, frirl_update_sarsa.c:348): q_ant [n][K], reward [n], cur_q_ant [n][K] */
int friq_batch_update_sarsa(friq_batch *b, const double *q_ant, const double *reward, const double *cur_q_ant);

/* one step of the classic API per agent (frirl_episode.c:148-160 with HOST environment callbacks): proposed action for
 * cur_q_states [n][S] (forced[i] >= 0 = the host's own random draw, frirl_e_greedy_selection.c:28-34; NULL = all greedy),
 * then frirl_update_sarsa(q_ant [n][K], reward [n], cur_q_states + action) when learn != 0.  Outputs may be NULL. */
int friq_batch_classic_step(friq_batch *b, const double *q_ant, const double *reward, const double *cur_q_states, const int *forced,
                            int learn, int *action, double *actconc, int *rules, int *inserted);
/* five_rule_distance (FIVE.h:95, five_rule_distance.c:63): ant [n][K] -> dists [n][rule_capacity], hit [n] (-1 none);
 * for n == 1 only the first R distances are written */
int friq_batch_rule_distance(friq_batch *b, const double *ant, double *dists, int *hit);
/* FIVEVagConcl_FRIRL_BestAct (FIVE.h:86, FIVEVagConcl_FRIRL_BestAct.c:56): dists [n][rule_capacity] -> conc [n] */
int friq_batch_bestact_from_dists(friq_batch *b, const double *dists, double *conc);
/* FIVE_add_rule (FIVE.h:100, five_add_rule.c:47) / five_remove_rule (FIVE.h:97, five_remove_rule.c:29) of one agent */
int friq_batch_append_rule(friq_batch *b, long agent, const double *ant, double q);
int friq_batch_remove_rule(friq_batch *b, long agent, int rule);

/* ---- rule-base reduction (frirl_sequential_run.c:170-350, strategy 1: smallest |Q| first, 2: largest |Q| first) of ONE rule
 * base rant [R][K], rconc [R] (reduced in place, *R updated) on a problem with a built-in environment.  Each candidate removal
 * is tested by one greedy episode without updates; up to `width` candidates are tested speculatively per launch and consumed
 * in order, which gives exactly the reference's sequential result.  verbose != 0 prints the reference's log lines. */
int friq_reduce_rulebase(friq_problem *p, int device, int strategy, double reward_tolerance, int width,
                         int *R, double *rant, double *rconc, int verbose, int *episodes_run);

/* ---- knowledge sharing between agents: the reference's experimental multi-agent mode (CMake option BUILD_OPENMP /
 * BUILD_MPI, frirl_agent.c) on the device */
/* merge_rb (frirl_agent.c:58-117): every rule of the sender list rant [Rs][K], rconc [Rs], in order, is folded into the rule base
 * of one agent (agent < 0: of every agent): FIVE_vag_concl_weight + FIVE_vag_concl at the sender's antecedents, then either
 * insertion / replacement at the snapped point (|difference| beyond the qdiff bounds) or the weighted assignment through the
 * receiver's FIVERB.weights - which the reference leaves STALE after an exact hit; the engine keeps that array under
 * FRIQ_RUN_KEEP_WEIGHTS.  The rule lists of agents on other GPUs travel by any transport (friq_dist.gather_rulebases: one
 * all-gather) and are merged on every rank in the same order. */
int friq_batch_merge_rb(friq_batch *b, long agent, int Rs, const double *rant, const double *rconc);
/* allocate FIVERB.weights for the batch: the per-call update entry points (friq_batch_update_sarsa, friq_batch_classic_step) then
 * keep it like the reference's update_rules does; friq_batch_run keeps it under FRIQ_RUN_KEEP_WEIGHTS */
int friq_batch_keep_weights(friq_batch *b);
/* frirl_omp_run (frirl_agent.c:294-385) for a batch of teams of `team` (>= 3) consecutive agents, all teams in parallel: chunks of
 * 9 episodes per member (frirl_sequential_run.c:51-65), then leader -> members and members -> leader merges, until the team's
 * leader (member 0) finds its rule base complete.  log (optional, max_log rows of 4 doubles): (agent, steps, return, rules) of
 * every finished episode, agent by agent; *n_log = rows produced; *rounds_out = rounds run (<= max_rounds). */
int friq_team_run(friq_batch *b, int team, int max_rounds, double *log, long max_log, long *n_log, int *rounds_out);

/* ---- every GPU of the node behind one handle: agents in contiguous blocks, one batch per GPU, no data-path
 * collective; the one exchange - per-agent results - is three grouped ncclAllGather calls over NVLink (NCCL is loaded at run time;
 * one process, ncclCommInitAll).  Stands in for the reference's only distributed mode, frirl_mpi_run (frirl_agent.c:387-467:
 * blocking MPI_Send / MPI_Recv between ranks).  Global agent i uses the Philox key seed0 + i whatever the number of GPUs. */
typedef struct friq_multi friq_multi;
/* n_devices <= 0: all GPUs of the node */
int friq_multi_create(const friq_problem_desc *desc, int n_devices, long n_agents_total, int rule_capacity, uint64_t seed0, friq_multi **out);
void friq_multi_destroy(friq_multi *m);
int friq_multi_devices(const friq_multi *m);
friq_batch *friq_multi_batch(friq_multi *m, int gpu);              /* GPU `gpu`'s batch, for every per-batch entry point above */
long friq_multi_first_agent(const friq_multi *m, int gpu);
long friq_multi_agents(const friq_multi *m, int gpu);              /* gpu < 0: n_agents_total */
/* host arrays of n_agents_total in global agent order (NULL keeps the current values) */
int friq_multi_set_hyper(friq_multi *m, const double *alpha, const double *gamma, const double *epsilon);
/* every GPU advances its agents n_steps env steps (friq_batch_run on the GPU's own stream: the GPUs run side by side) */
int friq_multi_run(friq_multi *m, long n_steps, int mode);
int friq_multi_sync(friq_multi *m);
/* all-gather of (return, steps, rules) into agent-ordered arrays on every GPU, {sum of rules, sum of returns, max rules} by one
 * kernel, then to the host arrays of n_agents_total (any pointer may be NULL); synchronous */
int friq_multi_gather(friq_multi *m, double *last_return, int *last_steps, int *rules, double *stats3);
int friq_multi_gathered_device(friq_multi *m, int gpu, void **d_ret_f64, void **d_steps_i32, void **d_rules_i32);
/* knowledge sharing across GPUs: the rule base of GLOBAL agent src_agent is merged (merge_rb, frirl_agent.c:58-117) into global
 * agent dst_agent, or into every agent of every GPU when dst_agent < 0 */
int friq_multi_merge_from(friq_multi *m, long src_agent, long dst_agent);

/* ---- rule-base files, byte-compatible with frirl_utils.c:100-281 */
int friq_save_rb_to_text_file(const char *file_name, int K, int R, const double *rant, const double *rconc);
int friq_save_rb_to_bin_file(const char *file_name, int K, int R, const double *rant, const double *rconc);
int friq_load_rb_from_bin_file(const char *file_name, int K, int max_rules, int *R, double *rant, double *rconc);

/* ---- batched rule-base I/O: the reference's formats either side of the batch */
/* one file per agent of [first, first + count): `pattern` is a printf format with one %ld (agent index); text != 0: "%.18f " format */
int friq_batch_save_rb_files(friq_batch *b, long first, long count, const char *pattern, int text);
/* warm start (`-f file`, frirl_app_helpers.c:53-62) of one agent, or of all agents when agent < 0 */
int friq_batch_load_rb_file(friq_batch *b, long agent, const char *file_name);
/* all rule bases + fus_is_rule_inserted flags of a batch in ONE file (per-agent records in the reference's binary layout), and -
 * when the batch keeps them (FRIQ_RUN_KEEP_WEIGHTS / friq_batch_keep_weights) - every agent's FIVERB.weights, so that a resumed
 * knowledge-sharing run merges exactly like an uninterrupted one */
int friq_batch_save_checkpoint(friq_batch *b, const char *file_name);
int friq_batch_load_checkpoint(friq_batch *b, const char *file_name);

/* ---- device self-tests / calibration used by tests and bench.py */
int friq_device_trig(const double *x, double *sin_out, double *cos_out, long n);        /* correctly rounded sin/cos */
int friq_device_pow_x87(const double *x, double *out, long n, int p, int force_exact);  /* fast_pow.inl:17-27 */
int friq_device_fp64_peak(double *ginstr_per_s);                                        /* DADD/DMUL lane-instr/s */

#ifdef __cplusplus
}
#endif
#endif
