/* oracle/ref_harness.c -- TEST INFRASTRUCTURE (never part of the product).
 *
 * Drives the UNMODIFIED reference (compiled from /root/reference by
 * oracle/Makefile into oracle/_ref/libfrirl_ref.so) through its own public API.
 * One binary per example (mountaincar / acrobot / cartpole): the example's
 * main() is compiled where it lies with -Dmain=ref_example_main and
 * -Dfrirl_run=ref_hook_frirl_run, so the example fills its frirl_desc exactly
 * as shipped and then hands control to ref_hook_frirl_run() below.
 *
 * Modes (environment variable REF_MODE):
 *   golden  run the construct phase exactly like the example would (with
 *           construct_rb=1 / reduce_rb=0 forced: mountaincar.c:240-242 ships
 *           the opposite and cannot produce its own golden);
 *           the example then writes <env>.frirlrb.{txt,bin} into the cwd.
 *   record  as golden, and additionally record per-call known-answer vectors
 *           (FIVE_vag_concl, FIVE_vag_concl_weight, frirl_get_best_action,
 *           frirl_update_sarsa, per-episode summary) into $REF_OUT.  The
 *           reference ships no per-call tests; these pin the oracle port and
 *           the CUDA kernels.  Recording works by symbol interposition: this
 *           executable defines the four entry points, the shared library's
 *           internal calls bind to them, and they forward via RTLD_NEXT.
 *   infer   CPU baseline of BASELINE config 4: REF_PROCS processes; each loads the rule base $REF_RBFILE (the reference's binary
 *           format, frirl_load_rb_from_bin_file), switches updates off (reduction_state = 1, frirl_test_run.c:65-69) and runs
 *           REF_REPS greedy episodes of REF_STEPS steps from random start states (theta1, theta2 ~ U(-pi/2, pi/2), the
 *           example's values_def fields are overwritten per episode); prints the same REFBENCH line.
 *   bench   CPU baseline: REF_PROCS processes, each stepping an independent
 *           reference agent (no_random=1, the reference RNG is unseedable) for
 *           >= REF_STEPS env steps of frirl_episode(); prints one line
 *           "REFBENCH procs=.. steps=.. seconds=.. steps_per_s=..".
 *
 * Record stream ($REF_OUT): a sequence of [double tag][double n][n doubles].
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/wait.h>

#include "frirl.h"
#include "frirl_app_helpers.h"

int ref_example_main(int argc, char **argv);
void frirl_run(struct frirl_desc *frirl, int verbose);

enum { TAG_HEADER = 1, TAG_GBA = 2, TAG_SARSA = 3, TAG_EPISODE = 4, TAG_FINAL = 5 };

static const char *g_mode = "golden";
static FILE *g_out = NULL;
static struct frirl_desc *g_frirl = NULL;
static int g_recording = 0;

/* ---- growable double buffer for one record ---- */
static double *g_buf = NULL;
static size_t g_len = 0, g_cap = 0;
static void put(double v) {
    if (g_len == g_cap) { g_cap = g_cap ? g_cap * 2 : 4096; g_buf = realloc(g_buf, g_cap * sizeof(double)); }
    g_buf[g_len++] = v;
}
static void putv(const double *v, size_t n) { for (size_t i = 0; i < n; i++) put(v[i]); }
static void flush_record(int tag) {
    double hdr[2] = { (double)tag, (double)g_len };
    fwrite(hdr, sizeof(double), 2, g_out);
    fwrite(g_buf, sizeof(double), g_len, g_out);
    g_len = 0;
}
static void put_rb(struct FIVERB *frb) {
    int K = frb->numofantecedents;
    put(frb->numofrules);
    putv(frb->rant, (size_t)frb->numofrules * K);
    putv(frb->rconc, frb->numofrules);
}

/* ---- interposed entry points ---- */
typedef unsigned int (*concl_fn)(struct FIVERB *, double *, double *);
typedef unsigned int (*weight_fn)(struct FIVERB *, double *, double *);
typedef unsigned int (*gba_fn)(struct frirl_desc *, fri_float *);
typedef void (*sarsa_fn)(struct frirl_desc *, fri_float *, fri_float, fri_float *);
static concl_fn real_concl; static weight_fn real_weight; static gba_fn real_gba; static sarsa_fn real_sarsa;
static void resolve(void) {
    if (real_concl) return;
    real_concl  = (concl_fn)dlsym(RTLD_NEXT, "FIVE_vag_concl");
    real_weight = (weight_fn)dlsym(RTLD_NEXT, "FIVE_vag_concl_weight");
    real_gba    = (gba_fn)dlsym(RTLD_NEXT, "frirl_get_best_action");
    real_sarsa  = (sarsa_fn)dlsym(RTLD_NEXT, "frirl_update_sarsa");
    if (!real_concl || !real_weight || !real_gba || !real_sarsa) { fprintf(stderr, "ref_harness: dlsym failed\n"); exit(3); }
}

/* nested-call capture while inside a sampled frirl_update_sarsa */
#define MAXNEST 8
static int in_sarsa = 0, n_concl = 0, have_weight = 0;
static double nest_ant[MAXNEST][FIVE_MAX_NUM_OF_UNIVERSES], nest_conc[MAXNEST], nest_ret[MAXNEST];
static double w_ant[FIVE_MAX_NUM_OF_UNIVERSES], w_ret; static double *w_weights = NULL; static int w_n = 0;

unsigned int FIVE_vag_concl(struct FIVERB *frb, double *ant, double *conc) {
    resolve();
    unsigned int r = real_concl(frb, ant, conc);
    if (in_sarsa && n_concl < MAXNEST) {
        memcpy(nest_ant[n_concl], ant, sizeof(double) * frb->numofantecedents);
        nest_conc[n_concl] = *conc; nest_ret[n_concl] = (r == ~0u) ? -1.0 : (double)r; n_concl++;
    }
    return r;
}
unsigned int FIVE_vag_concl_weight(struct FIVERB *frb, double *ant, double *weights) {
    resolve();
    unsigned int r = real_weight(frb, ant, weights);
    if (in_sarsa) {
        have_weight = 1;
        memcpy(w_ant, ant, sizeof(double) * frb->numofantecedents);
        w_ret = (r == ~0u) ? -1.0 : (double)r;
        w_n = frb->numofrules;
        w_weights = realloc(w_weights, sizeof(double) * (w_n + 4));
        if (r == ~0u) memcpy(w_weights, weights, sizeof(double) * w_n); else w_n = 0;
    }
    return r;
}

static long gba_calls = 0, gba_written = 0;
unsigned int frirl_get_best_action(struct frirl_desc *frirl, fri_float *states) {
    resolve();
    unsigned int r = real_gba(frirl, states);
    if (g_recording) {
        long stride = atol(getenv("REF_GBA_STRIDE") ? getenv("REF_GBA_STRIDE") : "400");
        long cap = atol(getenv("REF_GBA_CAP") ? getenv("REF_GBA_CAP") : "60");
        if ((gba_calls < 6 || gba_calls % stride == 0) && gba_written < cap) {
            put_rb(frirl->fiverb);
            putv(states, frirl->statedims_len);
            putv(frirl->fgba_actconc, frirl->actiondim.values_len);
            put(r);
            flush_record(TAG_GBA);
            gba_written++;
        }
        gba_calls++;
    }
    return r;
}

/* categories of update_sarsa outcomes, each sampled with its own stride/cap */
enum { CAT_INSERT, CAT_EXISTS_BOUND, CAT_HIT_SET, CAT_WEIGHTED, CAT_SKIPPED, NCAT };
static long cat_seen[NCAT], cat_written[NCAT];
void frirl_update_sarsa(struct frirl_desc *frirl, fri_float *q_ant, fri_float reward, fri_float *cur_q_ant) {
    resolve();
    if (!g_recording) { real_sarsa(frirl, q_ant, reward, cur_q_ant); return; }
    struct FIVERB *frb = frirl->fiverb;
    int K = frb->numofantecedents;
    int R0 = frb->numofrules;
    static double *rb0 = NULL; static size_t rb0cap = 0;
    size_t need = (size_t)R0 * (K + 1);
    if (need > rb0cap) { rb0cap = need * 2; rb0 = realloc(rb0, rb0cap * sizeof(double)); }
    memcpy(rb0, frb->rant, sizeof(double) * R0 * K);
    memcpy(rb0 + (size_t)R0 * K, frb->rconc, sizeof(double) * R0);
    double ins0 = frirl->fus_is_rule_inserted;
    double qa[FIVE_MAX_NUM_OF_UNIVERSES], cqa[FIVE_MAX_NUM_OF_UNIVERSES];
    memcpy(qa, q_ant, sizeof(double) * K); memcpy(cqa, cur_q_ant, sizeof(double) * K);

    in_sarsa = 1; n_concl = 0; have_weight = 0;
    real_sarsa(frirl, q_ant, reward, cur_q_ant);
    in_sarsa = 0;

    int cat;
    if (frb->numofrules > R0) cat = CAT_INSERT;
    else if (n_concl >= 3) cat = (have_weight && w_ret < 0) ? CAT_WEIGHTED : CAT_EXISTS_BOUND;
    else if (have_weight && w_ret < 0) cat = CAT_WEIGHTED;
    else {
        int changed = 0;
        for (int i = 0; i < R0; i++) if (memcmp(&frb->rconc[i], &rb0[(size_t)R0 * K + i], 8)) changed = 1;
        cat = changed ? CAT_HIT_SET : CAT_SKIPPED;
    }
    static const long stride[NCAT] = { 9, 7, 900, 25, 1 };
    static const long cap[NCAT]    = { 14, 8, 14, 14, 8 };
    long sscale = atol(getenv("REF_SARSA_STRIDE_SCALE") ? getenv("REF_SARSA_STRIDE_SCALE") : "1");
    if ((cat_seen[cat] % (stride[cat] * sscale) == 0 || cat_seen[cat] < 2) && cat_written[cat] < cap[cat]) {
        put(cat); put(ins0);
        put(R0); putv(rb0, (size_t)R0 * K); putv(rb0 + (size_t)R0 * K, R0);
        putv(qa, K); put(reward); putv(cqa, K);
        put(n_concl);
        for (int i = 0; i < n_concl; i++) { putv(nest_ant[i], K); put(nest_conc[i]); put(nest_ret[i]); }
        put(have_weight);
        if (have_weight) { putv(w_ant, K); put(w_ret); put(w_n); putv(w_weights, w_n); }
        put_rb(frb);
        put(frirl->fus_is_rule_inserted);
        flush_record(TAG_SARSA);
        cat_written[cat]++;
    }
    cat_seen[cat]++;
}

/* ---- header: every table the path consumes, as built by the reference ---- */
static void write_header(struct frirl_desc *frirl) {
    struct FIVERB *frb = frirl->fiverb;
    int K = frb->numofunivs, U = frb->univlength, S = frirl->statedims_len, A = frirl->actiondim.values_len;
    put(K); put(S); put(A); put(U); put(frb->p);
    put(frirl->alpha); put(frirl->gamma); put(frirl->epsilon);
    put(frirl->qdiff_pos_boundary); put(frirl->qdiff_neg_boundary); put(frirl->qdiff_final_tolerance);
    put(frirl->reward_good_above); put(frirl->rule_weight_considered_significant_for_update);
    put(frirl->skip_rules); put(frirl->no_random); put(frirl->max_episodes); put(frirl->max_steps);
    putv(frb->u, (size_t)K * U); putv(frb->ve, (size_t)K * U); putv(frb->udivs, K);
    for (int k = 0; k < S; k++) {
        put(frirl->statedims[k].values_len); put(frirl->statedims[k].values_div);
        put(frirl->statedims[k].values_steep); put(frirl->statedims[k].values_def);
        put(frirl->statedims[k].universe_div);
        putv(frirl->statedims[k].values, frirl->statedims[k].values_len);
    }
    put(frirl->actiondim.values_div); put(frirl->actiondim.universe_div);
    putv(frirl->possible_actions->values, A); putv(frirl->possible_actions->vevalues, A);
    put_rb(frb);
    flush_record(TAG_HEADER);
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static long g_bench_steps_target = 0; static int g_bench_pipe = -1;

/* construct loop with per-episode records; same control flow as the reference's
 * own run loop is obtained by simply calling the real frirl_run() - the per
 * episode summary is taken by wrapping frirl_episode(). */
typedef void (*episode_fn)(struct frirl_desc *);
static episode_fn real_episode;
void frirl_episode(struct frirl_desc *frirl) {
    if (!real_episode) real_episode = (episode_fn)dlsym(RTLD_NEXT, "frirl_episode");
    real_episode(frirl);
    if (g_recording) {
        put(frirl->episode_num); put(frirl->reward.ep_total_steps);
        put(frirl->reward.ep_total_value); put(frirl->fiverb->numofrules);
        flush_record(TAG_EPISODE);
    }
}

void ref_hook_frirl_run(struct frirl_desc *frirl, int verbose) {
    g_frirl = frirl;
    if (strcmp(g_mode, "shipped")) {
        frirl->construct_rb = 1;
        frirl->reduce_rb = 0;
    } else {
        /* "shipped": the example exactly as its main() configures it; REF_RBFILE stands in for `-f file`
         * (frirl_utils.c:380 mallocs strlen() bytes for the name and strcpy()s into it, which glibc's
         * fortified strcpy aborts on) */
        if (getenv("REF_RBFILE")) frirl->rbfile = strdup(getenv("REF_RBFILE"));
    }
    if (!strcmp(g_mode, "infer")) {
        const char *rb = getenv("REF_RBFILE");
        if (!rb || frirl_load_rb_from_bin_file(frirl, rb) < 0) { fprintf(stderr, "ref_harness: cannot load REF_RBFILE\n"); exit(7); }
        frirl->reduction_state = 1;                 /* no rule-base updates (frirl_episode.c:155) */
        frirl->max_steps = (int)g_bench_steps_target;
        int reps = atoi(getenv("REF_REPS") ? getenv("REF_REPS") : "1");
        unsigned long long x = 88172645463325252ull + (unsigned long long)getpid();
        long steps = 0; double rule_steps = 0;
        double t0 = now_s();
        for (int r = 0; r < reps; r++) {
            for (int k = 0; k < 2 && k < (int)frirl->statedims_len; k++) {
                x ^= x << 13; x ^= x >> 7; x ^= x << 17;                      /* xorshift64 */
                frirl->statedims[k].values_def = ((double)(x >> 11) * 0x1p-53 - 0.5) * 3.14159265358979323846;
            }
            frirl_episode(frirl);
            steps += frirl->reward.ep_total_steps;
            rule_steps += (double)frirl->fiverb->numofrules * frirl->reward.ep_total_steps;
            frirl->episode_num++;
        }
        double dt = now_s() - t0;
        double msg[4] = { (double)steps, dt, (double)reps, rule_steps };
        if (g_bench_pipe >= 0) { ssize_t w = write(g_bench_pipe, msg, sizeof msg); (void)w; }
        return;
    }
    if (!strcmp(g_mode, "bench")) {
        long steps = 0; int episodes = 0;
        double rule_steps = 0;                     /* sum over episodes of (rules at episode end) x (steps of the episode) */
        /* REF_EGREEDY=1: the example's own epsilon with the reference's libc rand() stream (unseedable, frirl_init.c:335-336) -
         * the same exploration RATE as the GPU arm; default: no_random = 1 as the examples ship */
        if (getenv("REF_EGREEDY") && atoi(getenv("REF_EGREEDY"))) frirl->no_random = 0;
        double t0 = now_s();
        while (steps < g_bench_steps_target) {
            frirl_episode(frirl);
            steps += frirl->reward.ep_total_steps;
            rule_steps += (double)frirl->fiverb->numofrules * frirl->reward.ep_total_steps;
            frirl->episode_num++; episodes++;
        }
        double dt = now_s() - t0;
        double msg[4] = { (double)steps, dt, (double)episodes, rule_steps };
        if (g_bench_pipe >= 0) { ssize_t w = write(g_bench_pipe, msg, sizeof msg); (void)w; }
        return;
    }
    if (!strcmp(g_mode, "record")) {
        write_header(frirl);
        g_recording = 1;
    }
    frirl_run(frirl, verbose);
    if (g_recording) {
        g_recording = 0;
        put_rb(frirl->fiverb);
        flush_record(TAG_FINAL);
    }
}

int main(int argc, char **argv) {
    if (getenv("REF_MODE")) g_mode = getenv("REF_MODE");
    char *one[2] = { argv[0], NULL };
    int quiet = !getenv("REF_VERBOSE");
    if (!strcmp(g_mode, "bench") || !strcmp(g_mode, "infer")) {
        const int infer = !strcmp(g_mode, "infer");
        int procs = atoi(getenv("REF_PROCS") ? getenv("REF_PROCS") : "1");
        g_bench_steps_target = atol(getenv("REF_STEPS") ? getenv("REF_STEPS") : "15000");
        int reps = atoi(getenv("REF_REPS") ? getenv("REF_REPS") : "1");
        int fds[2]; if (pipe(fds)) return 4;
        char tmpl[] = "/tmp/refbenchXXXXXX"; char *dir = mkdtemp(tmpl); if (!dir || chdir(dir)) return 5;
        double t0 = now_s();
        for (int p = 0; p < procs; p++) {
            pid_t pid = fork();
            if (pid == 0) {
                close(fds[0]); g_bench_pipe = fds[1];
                if (!freopen("/dev/null", "w", stdout)) return 6;
                for (int r = 0; r < (infer ? 1 : reps); r++) ref_example_main(1, one);   /* infer: the episodes loop inside the hook */
                _exit(0);
            }
        }
        close(fds[1]);
        double steps = 0, busy = 0, maxbusy = 0, rule_steps = 0; int nmsg = 0; double msg[4];
        while (read(fds[0], msg, sizeof msg) == (ssize_t)sizeof msg) { steps += msg[0]; busy += msg[1]; rule_steps += msg[3]; nmsg++; }
        while (wait(NULL) > 0) {}
        double wall = now_s() - t0;
        /* throughput = steps / (sum of in-frirl_episode time / procs): excludes init + file dump */
        double eff = busy / procs; if (eff > maxbusy) maxbusy = eff;
        printf("REFBENCH env=%s procs=%d reps=%d steps=%.0f busy_s_per_proc=%.6f wall_s=%.6f steps_per_s=%.1f mean_rules=%.2f agents_done=%d\n",
               REF_ENV_NAME, procs, reps, steps, eff, wall, steps / eff, steps > 0 ? rule_steps / steps : 0.0, nmsg);
        char cmd[256]; snprintf(cmd, sizeof cmd, "rm -rf %s", dir); if (system(cmd)) {}
        return 0;
    }
    if (!strcmp(g_mode, "record")) {
        const char *path = getenv("REF_OUT");
        if (!path) { fprintf(stderr, "REF_OUT not set\n"); return 2; }
        g_out = fopen(path, "wb"); if (!g_out) { perror("REF_OUT"); return 2; }
    }
    int saved = -1;
    if (quiet) { fflush(stdout); saved = dup(1); if (!freopen("/dev/null", "w", stdout)) return 6; }
    (void)argc;
    ref_example_main(1, one);
    if (g_out) fclose(g_out);
    return 0;
}
