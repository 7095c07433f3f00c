/* oracle/friq_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the reference's FRIQ-learning hot path, written
 * from the normative pseudo-code in SURVEY.md 3.5 / 8(a).  It exists to CHECK
 * the CUDA engine: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  It is never linked into
 * the product (fri-reinforcementlearning-c_b200/), which has no CPU fallback.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this port against
 *  (a) the reference's own golden rule bases (tests/orig, committed as
 *      tests/golden/<env>.frirlrb.txt after a byte-compare at generation time),
 *  (b) per-call vectors recorded from the compiled reference (oracle/_ref) for
 *      FIVE_vag_concl, FIVE_vag_concl_weight, frirl_get_best_action and
 *      frirl_update_sarsa, and
 *  (c) the reference's init-time tables (u, ve, udivs, action VE values).
 * All comparisons are bit-exact.
 */
#ifndef FRIQ_ORACLE_H
#define FRIQ_ORACLE_H
#include <stdint.h>

#define FO_MAXK 8      /* FIVE_MAX_NUM_OF_UNIVERSES, FIVE.h:19 */
#define FO_MAXGRID 64
#define FO_MAXA 64

enum { FO_ENV_MOUNTAINCAR = 0, FO_ENV_ACROBOT = 1, FO_ENV_CARTPOLE = 2 };
enum { FO_TRIG_LIBM = 0, FO_TRIG_CR = 1 };

typedef struct fo_problem {
    int env, S, K, A, U, p;
    double *u, *ve;                    /* [K][U] */
    double udiv[FO_MAXK];
    int grid_len[FO_MAXK];             /* state grids, then the action grid at index S */
    double grid[FO_MAXK][FO_MAXGRID];
    double grid_div[FO_MAXK];          /* values_div (quantiser step / possible_states epsilon) */
    double values_def[FO_MAXK];
    double act_val[FO_MAXA], act_ve[FO_MAXA];
    double alpha, gamma, epsilon, qdiff_pos, qdiff_neg, qdiff_final_tol, reward_good_above, wsig;
    int skip_rules, no_random, max_episodes, max_steps;
    int trig;
} fo_problem;

typedef struct fo_agent {
    const fo_problem *P;
    int cap, R;
    double *rant;    /* [cap][K] canonical antecedents */
    double *rve;     /* [cap][K] VE value of each antecedent (five_add_rule.c:71-84) */
    double *rconc;   /* [cap] */
    double *w;       /* [cap] scratch: wi */
    double *weights; /* [cap] FIVERB.weights (FIVE.h:24-76): the normalised weights of the LAST FIVE_vag_concl_weight call that
                      * found no exact hit - persistent state: merge_rb (frirl_agent.c:58-117) re-uses it STALE after exact hits */
    int epended;     /* frirl_desc.epended: the "same as the previous episode" predicate held in this chunk (frirl_sequential_run.c:90) */
    double *prev_rconc;
    /* episode state (frirl_episode.c) */
    double s[FO_MAXK], qant[FO_MAXK];
    int need_reset, step_in_ep, inserted, overflow;
    uint32_t episode_num;
    double ep_return; int ep_steps; int success;
    /* results of the last finished episode + convergence bookkeeping */
    double last_return; int last_steps, last_rules; int converged;
    double prev_reward; int prev_steps, prev_numru;
    long total_steps; double sum_rules;
    uint64_t seed;   /* Philox key */
    double alpha, gamma, epsilon; /* per-agent hyper-parameters (sweep) */
    int has_s0; double s0[FO_MAXK];   /* optional per-agent episode start state */
    int learning;    /* 0 = inference (reduction_state=1 semantics, frirl_test_run.c:65-69) */
} fo_agent;

fo_problem *fo_problem_builtin(int env, int trig);
fo_problem *fo_problem_builtin_actions(int env, int trig, int A, const double *act_vals, double act_div);
void fo_problem_free(fo_problem *P);

fo_agent *fo_agent_new(const fo_problem *P, int capacity, uint64_t seed);
void fo_agent_free(fo_agent *a);
void fo_agent_set_rb(fo_agent *a, int R, const double *rant, const double *rconc);
int  fo_add_rule(fo_agent *a, const double *ant, double q);

/* mirrors of the reference entry points (return value: hit rule or -1) */
int  fo_nearest_index(const fo_problem *P, int k, double x, int len);
int  fo_vag_concl(fo_agent *a, const double *ant, double *conc);
int  fo_vag_concl_weight(fo_agent *a, const double *ant, double *weights);
int  fo_get_best_action(fo_agent *a, const double *states, double *actconc);
void fo_update_sarsa(fo_agent *a, const double *q_ant, double reward, const double *cur_q_ant);
double fo_snap(const fo_problem *P, int k, double x);
void fo_agent_set_initial_state(fo_agent *a, const double *s0);
/* knowledge sharing: frirl_agent.c merge_rb / frirl_omp_run */
void fo_merge_rb(fo_agent *rcvr, int Rs, const double *srant, const double *srconc);
int  fo_team_run(fo_agent **ag, int W, int max_rounds, double *log, int max_log);
int  fo_team_run_mpi(fo_agent **ag, int W, int max_rounds, double *log, int max_log);   /* the exchange order of frirl_mpi_run (frirl_agent.c:427-461) */

/* environment */
void fo_env_step(const fo_problem *P, double action, const double *s, double *snew);
void fo_env_reward(const fo_problem *P, const double *s, double *reward, int *done);
void fo_env_quantize(const fo_problem *P, const double *s, double *q);

/* stepping: n env steps, episodes roll over (batch-engine semantics) */
void fo_agent_run_steps(fo_agent *a, long n_steps, int stop_on_converge);
/* the reference's construct loop (frirl_sequential_run.c:57-161); returns #episodes;
 * ep_log (may be NULL) receives 4 doubles per episode: ep, steps, reward, rules */
int  fo_run_construct(fo_agent *a, double *ep_log, int max_log);

void fo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double fo_pow_x87(double b, int p);
double fo_sin(const fo_problem *P, double x);
double fo_cos(const fo_problem *P, double x);

#endif
