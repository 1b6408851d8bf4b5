/* TEST INFRASTRUCTURE — CPU oracle: inverse kinematics (placeholder until the IK row of SURVEY §8 a10/a11 is built). */
#include "orc_ik.h"
int orc_ik_solve(OrcEnv* e, const double*, const double*, int, double, double* out11) {
    for (int i = 0; i < ORC_NDOF; i++) out11[i] = e->q[i];
    return -1;
}
