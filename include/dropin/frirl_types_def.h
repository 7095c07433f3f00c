/* include/dropin/frirl_types_def.h -- the default parameter block programs start from
 * (`struct frirl_desc frirl = frirl_desc_default;`); same values as the reference's
 * src/frirl/frirl_types_def.h:22-77.  Unlisted members are zero / NULL. */
#ifndef FRIQ_REFAPI_DEFAULTS_H
#define FRIQ_REFAPI_DEFAULTS_H
#include "friq_refapi.h"

static const struct frirl_desc frirl_desc_default = {
    /* run control */
    .argc = 1, .runmode = FRIRL_SEQ, .verbose = 1, .agent_rnd_init = 1,
    .gui_width = 640, .gui_height = 480,
    .construct_rb = 1, .reduce_rb = 0, .reduction_strategy = FRIRL_REDUCTION_STRATEGY_DEFAULT,
    .max_episodes = 1000, .max_steps = 1000, .five_maxnumofrules = 16384,
    /* learning */
    .alpha = 0.5, .gamma = 1.0, .epsilon = 0.001,
    .qdiff_pos_boundary = 1.0, .qdiff_neg_boundary = -250.0, .qdiff_final_tolerance = 250.0,
    .reward_good_above = 0.0, .reduction_reward_tolerance = 0.0,
    .rule_weight_considered_significant_for_update = 0.05,
    .skip_rules = 0, .no_random = 1,
    /* imitation: off */
    .keyaction = (unsigned int)-1, .original_learning = 1,
};
#endif
