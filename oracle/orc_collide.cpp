/* TEST INFRASTRUCTURE — CPU oracle: narrowphase (placeholder). */
#include "orc_env.h"
void orc_collide(OrcEnv* e) { e->n_manifolds = 0; }
