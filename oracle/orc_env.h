/* TEST INFRASTRUCTURE — CPU oracle: one REALRobot2020 env instance (scalar, fp64 physics).
 * Parity status: UNPINNED (see orc_scene.h). */
#ifndef ORC_ENV_H
#define ORC_ENV_H
#include "orc_scene.h"

#define ORC_MAX_MANIFOLDS 32
#define ORC_MAX_ROWS 1024

struct OrcContactPoint {
    V3 localA, localB;       /* in shape A / B local (COM) frames */
    V3 worldA, worldB;       /* positionWorldOnA / OnB */
    V3 normalB;              /* world normal on B (points from B towards A) */
    double distance;
    double applied_impulse;  /* normal impulse of the last solve (warm start, touch read-out) */
    double applied_f1, applied_f2;
    int lifetime;
};
struct OrcManifold {
    int shapeA, shapeB;      /* scene shape indices; A is always the "movable" / robot side */
    int n;
    OrcContactPoint pt[4];
    int touched;             /* pair visited this step (AABBs overlapped) */
};

struct OrcEnv {
    const OrcScene* sc;
    OrcConfig cfg;
    /* state */
    double q[ORC_NDOF], qd[ORC_NDOF];
    double target[ORC_NDOF];
    OrcObjState obj[ORC_MAXOBJ];
    long timestep;
    /* derived by forward kinematics */
    M3 Rpc[ORC_NL];  V3 rvec[ORC_NL];          /* parent->child rotation, parent COM -> child COM (child frame) */
    OrcFrame link[ORC_NL];                     /* world frames of the link COMs */
    /* ABA cache (articulated inertias etc., reused for unit-impulse responses) */
    double IA[ORC_NL][6][6];
    double h[ORC_NL][6];   double invD[ORC_NL];
    /* contacts */
    OrcManifold manifolds[ORC_MAX_MANIFOLDS];
    int n_manifolds;
    double touch[4];
    /* diagnostics of the last step */
    int last_rows, last_iterations, last_contacts;
    int last_overlaps, manifold_overflow;      /* AABB-overlapping pairs of the last step; a non-empty manifold found no slot */
    double last_residual;
};

void orc_env_init(OrcEnv* e, const OrcScene* sc, const OrcConfig* cfg);
void orc_env_reset(OrcEnv* e);
void orc_fk(OrcEnv* e);
/* pybullet stepSimulation restatement */
void orc_step_simulation(OrcEnv* e);
/* env.py:326-356 step_joints: a2 + a3 + a4 + a5 */
void orc_env_step(OrcEnv* e, const double* joint_cmd9);
void orc_calc_state(const OrcEnv* e, double* joints9);      /* robot.py:203-211 */
/* world frame (COM / inertial frame) of scene shape s, given current state */
OrcFrame orc_shape_frame(const OrcEnv* e, int shape);
/* forward dynamics only (no constraints): qdd for the arm; used by unit tests (ABA vs CRBA) */
void orc_aba(OrcEnv* e, const double* tau, double* qdd);
void orc_mass_matrix(OrcEnv* e, double* M /*11x11*/);
void orc_collide(OrcEnv* e);   /* narrowphase + persistent manifolds (orc_collide.cpp) */
#endif
