/* tests/dropin/dropin_hook.c -- the example drivers' call to frirl_run() lands here (-Dfrirl_run=...).
 * DROPIN_CONSTRUCT=1 asks for the construction run the goldens in tests/orig were made with
 * (mountaincar's main() hard-codes construct_rb = 0 / reduce_rb = 1); everything else is passed through. */
#include "frirl.h"
#include "frirl_app_helpers.h"
#include <stdlib.h>
#undef frirl_run
void frirl_run(struct frirl_desc *frirl, int verbose);

void dropin_hook_frirl_run(struct frirl_desc *frirl, int verbose)
{
    if (getenv("DROPIN_CONSTRUCT")) { frirl->construct_rb = 1; frirl->reduce_rb = 0; }
    frirl_run(frirl, verbose);
}
