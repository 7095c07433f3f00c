/* csrc/friq_multi.c -- every GPU of the node behind ONE C handle (friq_multi_*, include/friq_b200.h).
 *
 * The path shards by agent: agents never interact, so a node's GPUs each hold a contiguous block of the agents
 * (block i starts at agent i * ceil(n / G); global agent index = Philox key offset, so the sharded run equals one big batch agent
 * for agent) and step it with the very same kernels - no data-path collective.  The ONE exchange is the gather of per-agent
 * results (return f64, steps i32, rules i32): three grouped ncclAllGather calls straight into agent-ordered arrays on every GPU
 * (one communicator per GPU, ncclCommInitAll, one process), plus the statistics kernel on GPU 0.
 *
 * What it replaces in the reference: its only distributed mode, frirl_mpi_run (frirl_agent.c:387-467), ships rule bases between
 * MPI ranks with blocking MPI_Send / MPI_Recv; here the results of a node's GPUs travel over NVLink by NCCL.  NCCL is loaded at
 * run time (dlopen "libnccl.so.2"): the library has no link-time dependency on it, and callers that never use friq_multi_* never
 * load it.  No NCCL => friq_multi_gather fails loudly (there is no host-memory fallback for the collective).
 */
#define _GNU_SOURCE
#include "../../include/friq_b200.h"
#include "friq_devprob.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define FRIQ_MULTI_MAXDEV 16
extern void friq_set_error_(int code, const char *fmt, ...);      /* friq_host.c: thread-local message behind friq_last_error() */
#define MFAIL(code, ...) do { friq_set_error_((code), __VA_ARGS__); return (code); } while (0)
#define MDEV(call) do { int e_ = (call); if (e_) MFAIL(e_, "%s failed: %s", #call, friqd_error_string(e_)); } while (0)
#define MCALL(call) do { int e_ = (call); if (e_) return e_; } while (0)

/* the slice of nccl.h this file needs (NCCL 2.x ABI) */
typedef struct ncclComm *ncclComm_t;
enum { NCCL_INT32 = 2, NCCL_FLOAT64 = 8 };
static struct {
    void *h;
    int (*CommInitAll)(ncclComm_t *, int, const int *);
    int (*CommDestroy)(ncclComm_t);
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, void *);
    int (*GroupStart)(void);
    int (*GroupEnd)(void);
    const char *(*GetErrorString)(int);
    int (*GetVersion)(int *);
} N;

static int nccl_load(void)
{
    if (N.h) return 0;
    const char *names[] = { "libnccl.so.2", "libnccl.so", NULL };
    void *h = NULL;
    for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!h) MFAIL(7, "friq_multi: NCCL not found (dlopen libnccl.so.2: %s)", dlerror());
    N.CommInitAll = (int (*)(ncclComm_t *, int, const int *))dlsym(h, "ncclCommInitAll");
    N.CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    N.AllGather = (int (*)(const void *, void *, size_t, int, ncclComm_t, void *))dlsym(h, "ncclAllGather");
    N.GroupStart = (int (*)(void))dlsym(h, "ncclGroupStart");
    N.GroupEnd = (int (*)(void))dlsym(h, "ncclGroupEnd");
    N.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    N.GetVersion = (int (*)(int *))dlsym(h, "ncclGetVersion");
    if (!N.CommInitAll || !N.CommDestroy || !N.AllGather || !N.GroupStart || !N.GroupEnd || !N.GetErrorString)
        MFAIL(7, "friq_multi: libnccl lacks a required entry point");
    N.h = h;
    return 0;
}
#define NCCLCK(call) do { int r_ = (call); if (r_) MFAIL(8, "%s failed: %s", #call, N.GetErrorString(r_)); } while (0)

struct friq_multi {
    int G;                                   /* GPUs in use */
    int dev[FRIQ_MULTI_MAXDEV];
    long n_total, per, cnt[FRIQ_MULTI_MAXDEV], first[FRIQ_MULTI_MAXDEV];
    friq_problem *P[FRIQ_MULTI_MAXDEV];      /* a problem's tables live on ONE device: one copy per GPU */
    friq_batch *B[FRIQ_MULTI_MAXDEV];
    void *stream[FRIQ_MULTI_MAXDEV];
    ncclComm_t comm[FRIQ_MULTI_MAXDEV];
    int have_comm;
    /* per GPU: this GPU's block (padded to `per` agents) and the gathered arrays of all G * per agents */
    double *l_ret[FRIQ_MULTI_MAXDEV], *g_ret[FRIQ_MULTI_MAXDEV], *stats[FRIQ_MULTI_MAXDEV];
    int *l_steps[FRIQ_MULTI_MAXDEV], *l_rules[FRIQ_MULTI_MAXDEV], *g_steps[FRIQ_MULTI_MAXDEV], *g_rules[FRIQ_MULTI_MAXDEV];
};

int friq_multi_create(const friq_problem_desc *desc, int n_devices, long n_agents_total, int rule_capacity, uint64_t seed0, friq_multi **out)
{
    int have = 0;
    MCALL(friq_device_count(&have));
    if (have < 1) MFAIL(3, "friq_multi_create: no CUDA device");
    int G = (n_devices <= 0 || n_devices > have) ? have : n_devices;
    if (G > FRIQ_MULTI_MAXDEV) G = FRIQ_MULTI_MAXDEV;
    if (n_agents_total < G) MFAIL(4, "friq_multi_create: %ld agents for %d GPUs", n_agents_total, G);
    friq_multi *m = (friq_multi *)calloc(1, sizeof *m);
    m->G = G; m->n_total = n_agents_total;
    m->per = (n_agents_total + G - 1) / G;
    for (int i = 0; i < G; i++) {
        m->dev[i] = i;
        m->first[i] = (long)i * m->per;
        m->cnt[i] = m->first[i] + m->per <= n_agents_total ? m->per : n_agents_total - m->first[i];
        if (m->cnt[i] < 1) { friq_multi_destroy(m); MFAIL(4, "friq_multi_create: GPU %d would hold no agent (%ld agents over %d GPUs)", i, n_agents_total, G); }
        int rc = friq_problem_create(desc, &m->P[i]);
        if (!rc) rc = friq_batch_create(m->P[i], m->dev[i], m->cnt[i], rule_capacity, seed0, m->first[i], &m->B[i]);
        if (rc) { friq_multi_destroy(m); return rc; }
        MDEV(friqd_set_device(m->dev[i]));
        MDEV(friqd_stream_create(&m->stream[i]));
        const size_t p = (size_t)m->per, t = (size_t)m->per * G;
        MDEV(friqd_malloc((void **)&m->l_ret[i], 8 * p)); MDEV(friqd_malloc((void **)&m->l_steps[i], 4 * p)); MDEV(friqd_malloc((void **)&m->l_rules[i], 4 * p));
        MDEV(friqd_memset(m->l_ret[i], 0, 8 * p)); MDEV(friqd_memset(m->l_steps[i], 0, 4 * p)); MDEV(friqd_memset(m->l_rules[i], 0, 4 * p));
        MDEV(friqd_malloc((void **)&m->g_ret[i], 8 * t)); MDEV(friqd_malloc((void **)&m->g_steps[i], 4 * t)); MDEV(friqd_malloc((void **)&m->g_rules[i], 4 * t));
        MDEV(friqd_malloc((void **)&m->stats[i], 8 * 4));
    }
    *out = m;
    return 0;
}

void friq_multi_destroy(friq_multi *m)
{
    if (!m) return;
    for (int i = 0; i < m->G; i++) {
        friqd_set_device(m->dev[i]);
        friqd_device_sync();
        if (m->have_comm && m->comm[i]) N.CommDestroy(m->comm[i]);
        void *p[] = { m->l_ret[i], m->l_steps[i], m->l_rules[i], m->g_ret[i], m->g_steps[i], m->g_rules[i], m->stats[i] };
        for (unsigned k = 0; k < sizeof p / sizeof p[0]; k++) if (p[k]) friqd_free(p[k]);
        if (m->stream[i]) friqd_stream_destroy(m->stream[i]);
        if (m->B[i]) friq_batch_destroy(m->B[i]);
        if (m->P[i]) friq_problem_destroy(m->P[i]);
    }
    free(m);
}

int friq_multi_devices(const friq_multi *m) { return m->G; }
friq_batch *friq_multi_batch(friq_multi *m, int i) { return (i >= 0 && i < m->G) ? m->B[i] : NULL; }
long friq_multi_first_agent(const friq_multi *m, int i) { return (i >= 0 && i < m->G) ? m->first[i] : -1; }
long friq_multi_agents(const friq_multi *m, int i) { return (i >= 0 && i < m->G) ? m->cnt[i] : (i < 0 ? m->n_total : 0); }

/* per-agent hyper-parameters, host arrays of n_agents_total in GLOBAL agent order (NULL keeps the current values) */
int friq_multi_set_hyper(friq_multi *m, const double *alpha, const double *gamma, const double *epsilon)
{
    for (int i = 0; i < m->G; i++)
        MCALL(friq_batch_set_hyper(m->B[i], alpha ? alpha + m->first[i] : NULL, gamma ? gamma + m->first[i] : NULL, epsilon ? epsilon + m->first[i] : NULL));
    return 0;
}

/* every GPU advances its agents by n_steps env steps; the launches are asynchronous, so the GPUs run side by side */
int friq_multi_run(friq_multi *m, long n_steps, int mode)
{
    for (int i = 0; i < m->G; i++) MCALL(friq_batch_run(m->B[i], n_steps, mode, m->stream[i]));
    return 0;
}

int friq_multi_sync(friq_multi *m)
{
    for (int i = 0; i < m->G; i++) MCALL(friq_batch_sync(m->B[i], m->stream[i]));
    return 0;
}

/* The path's one exchange: every GPU's per-agent results -> agent-ordered arrays on EVERY GPU (three grouped all-gathers over
 * NVLink), statistics {sum of rules, sum of returns, max rules} by one kernel on GPU 0, then one copy to the host arrays of
 * n_agents_total (any pointer may be NULL).  Synchronous. */
int friq_multi_gather(friq_multi *m, double *last_return, int *last_steps, int *rules, double *stats3)
{
    const int G = m->G;
    for (int i = 0; i < G; i++) MCALL(friq_batch_gather_device(m->B[i], m->l_ret[i], m->l_steps[i], m->l_rules[i], m->stream[i]));
    if (G > 1) {
        MCALL(nccl_load());
        if (!m->have_comm) { NCCLCK(N.CommInitAll(m->comm, G, m->dev)); m->have_comm = 1; }
        NCCLCK(N.GroupStart());
        for (int i = 0; i < G; i++) {
            int r = N.AllGather(m->l_ret[i], m->g_ret[i], (size_t)m->per, NCCL_FLOAT64, m->comm[i], m->stream[i]);
            if (!r) r = N.AllGather(m->l_steps[i], m->g_steps[i], (size_t)m->per, NCCL_INT32, m->comm[i], m->stream[i]);
            if (!r) r = N.AllGather(m->l_rules[i], m->g_rules[i], (size_t)m->per, NCCL_INT32, m->comm[i], m->stream[i]);
            if (r) { N.GroupEnd(); MFAIL(8, "ncclAllGather failed: %s", N.GetErrorString(r)); }
        }
        NCCLCK(N.GroupEnd());
    } else {
        MDEV(friqd_set_device(m->dev[0]));
        MDEV(friqd_d2d_async(m->g_ret[0], m->l_ret[0], 8 * (size_t)m->per, m->stream[0]));
        MDEV(friqd_d2d_async(m->g_steps[0], m->l_steps[0], 4 * (size_t)m->per, m->stream[0]));
        MDEV(friqd_d2d_async(m->g_rules[0], m->l_rules[0], 4 * (size_t)m->per, m->stream[0]));
    }
    MDEV(friqd_set_device(m->dev[0]));
    MDEV(friqd_result_stats(m->g_ret[0], m->g_rules[0], m->n_total, m->stats[0], m->stream[0]));
    for (int i = 0; i < G; i++) { MDEV(friqd_set_device(m->dev[i])); MDEV(friqd_sync(m->stream[i])); }
    MDEV(friqd_set_device(m->dev[0]));
    if (last_return) MDEV(friqd_d2h(last_return, m->g_ret[0], 8 * (size_t)m->n_total));
    if (last_steps) MDEV(friqd_d2h(last_steps, m->g_steps[0], 4 * (size_t)m->n_total));
    if (rules) MDEV(friqd_d2h(rules, m->g_rules[0], 4 * (size_t)m->n_total));
    if (stats3) MDEV(friqd_d2h(stats3, m->stats[0], 8 * 3));
    return 0;
}

/* the gathered arrays as they lie on GPU `i` (device pointers, n_agents_total entries each, valid after friq_multi_gather) */
int friq_multi_gathered_device(friq_multi *m, int i, void **d_ret, void **d_steps, void **d_rules)
{
    if (i < 0 || i >= m->G) MFAIL(4, "friq_multi_gathered_device: GPU %d of %d", i, m->G);
    if (d_ret) *d_ret = m->g_ret[i];
    if (d_steps) *d_steps = m->g_steps[i];
    if (d_rules) *d_rules = m->g_rules[i];
    return 0;
}

/* Knowledge sharing across the node's GPUs (what frirl_mpi_run's MPI_Send / MPI_Recv of numofrules, rant, rconc do,
 * frirl_agent.c:427-461): the rule base of GLOBAL agent `src_agent` is folded into every agent of every GPU (dst_agent < 0) or
 * into the one global agent dst_agent, by merge_rb (friq_batch_merge_rb) - every GPU applies the same sender list in the same
 * order, so the result does not depend on the number of GPUs. */
int friq_multi_merge_from(friq_multi *m, long src_agent, long dst_agent)
{
    if (src_agent < 0 || src_agent >= m->n_total || dst_agent >= m->n_total) MFAIL(5, "friq_multi_merge_from: agent out of range");
    const int gs = (int)(src_agent / m->per);
    int dims[6]; friq_problem_dims(m->P[gs], dims);
    const int K = dims[1];
    int R = 0;
    MCALL(friq_batch_get_rulebase(m->B[gs], src_agent - m->first[gs], &R, NULL, NULL));
    double *rant = (double *)malloc(sizeof(double) * (size_t)(R + 1) * K), *rconc = (double *)malloc(sizeof(double) * (size_t)(R + 1));
    int rc = friq_batch_get_rulebase(m->B[gs], src_agent - m->first[gs], &R, rant, rconc);
    for (int i = 0; i < m->G && !rc; i++) {
        if (dst_agent >= 0) {
            if (dst_agent / m->per != i) continue;
            rc = friq_batch_merge_rb(m->B[i], dst_agent - m->first[i], R, rant, rconc);
        } else rc = friq_batch_merge_rb(m->B[i], -1, R, rant, rconc);
    }
    free(rant); free(rconc);
    return rc;
}
