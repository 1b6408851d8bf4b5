/* TEST INFRASTRUCTURE — CPU oracle: inverse kinematics restatement. See orc_ik.cpp. */
#ifndef ORC_IK_H
#define ORC_IK_H
#include "orc_env.h"
int orc_ik_solve(OrcEnv* e, const double* pos3, const double* quat4, int max_iter, double residual, double* out11);
#endif
