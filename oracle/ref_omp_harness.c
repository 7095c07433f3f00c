/* oracle/ref_omp_harness.c -- TEST INFRASTRUCTURE (never part of the product).
 *
 * Drives the UNMODIFIED reference, compiled by oracle/Makefile with its experimental CMake option BUILD_OPENMP, through
 * its knowledge-sharing run mode: frirl_run() -> frirl_omp_run() (frirl_agent.c:294-385): OMP_NUM_THREADS agents, each
 * learning in chunks of FRIRL_AGENT_EPCHUNK - 1 = 9 episodes (frirl_sequential_run.c:51-65), rule bases merged with
 * merge_rb() (frirl_agent.c:58-117) after every chunk.  The example fills its frirl_desc exactly as shipped; like
 * ref_harness.c this hook flips mountaincar's hard-coded construct_rb = 0 / reduce_rb = 1.
 * Output: the library's own "#<agent> Episode: ..." lines on stdout and <env>.frirlrb.txt of agent 0's final rule base.
 */
#include <stdio.h>
#include <stdlib.h>
#include "frirl.h"
#include "frirl_app_helpers.h"

int ref_example_main(int argc, char **argv);
void frirl_run(struct frirl_desc *frirl, int verbose);

void ref_hook_frirl_run(struct frirl_desc *frirl, int verbose) {
    frirl->construct_rb = 1;
    frirl->reduce_rb = 0;
    frirl->runmode = FRIRL_OMP;
    frirl_run(frirl, verbose);
}
int main(int argc, char **argv) { return ref_example_main(argc, argv); }
