/* tests/dropin/dropin_hook.c -- the example drivers' call to frirl_run() lands here (-Dfrirl_run=...).
 * DROPIN_CONSTRUCT=1 asks for the construction run the goldens in tests/orig were made with
 * (mountaincar's main() hard-codes construct_rb = 0 / reduce_rb = 1); everything else is passed through.
 * DROPIN_TIMING=1 prints the wall time of frirl_run() itself (bench.py --workload c1 / c2: the number BASELINE.md 4.2 asks for,
 * process start and CUDA context creation excluded like the reference arm excludes its process start). */
#include "frirl.h"
#include "frirl_app_helpers.h"
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#undef frirl_run
void frirl_run(struct frirl_desc *frirl, int verbose);

void dropin_hook_frirl_run(struct frirl_desc *frirl, int verbose)
{
    struct timespec t0, t1;
    if (getenv("DROPIN_CONSTRUCT")) { frirl->construct_rb = 1; frirl->reduce_rb = 0; }
    clock_gettime(CLOCK_MONOTONIC, &t0);
    frirl_run(frirl, verbose);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (getenv("DROPIN_TIMING"))
        printf("DROPIN_TIMING frirl_run_s=%.6f\n", (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
}
