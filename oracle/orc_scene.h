/* TEST INFRASTRUCTURE — CPU oracle for the REALRobot2020 step path.
 * Scene description loaded from the asset blob (format: real_robots_b200/blob.py), and the
 * oracle's per-env state.  Parity status: UNPINNED — the arithmetic of the reference path lives in
 * pybullet (un-vendored, unpinned: /root/reference/setup.py:24-32) which cannot be run here; this
 * restates bullet3's published algorithm from the reference's call sites. */
#ifndef ORC_SCENE_H
#define ORC_SCENE_H
#include <cstdint>
#include <string>
#include <vector>
#include "orc_math.h"

#define ORC_NL 16      /* robot links (without the fixed base link_0) */
#define ORC_NDOF 11    /* 7 arm + 4 finger revolute joints */
#define ORC_MAXOBJ 3
#define ORC_W 320
#define ORC_H 240

struct OrcBlobEntry { std::string name; int dtype; int ndim; uint64_t shape[4]; const uint8_t* data; uint64_t nbytes; };
struct OrcBlob {
    std::vector<uint8_t> raw;
    std::vector<OrcBlobEntry> entries;
    bool load(const char* path, std::string& err);
    const OrcBlobEntry* find(const char* name) const;
    const double* f64(const char* name) const;
    const float* f32(const char* name) const;
    const int32_t* i32(const char* name) const;
    const int64_t* i64(const char* name) const;
    const uint8_t* u8(const char* name) const;
    uint64_t dim(const char* name, int i) const;
};

struct OrcShape {
    int body, link;
    int tri_start, tri_count, hull_start, hull_count, tex;
    float rgba[4];
    double friction, restitution, rolling, spinning;
};
struct OrcTexture { int w, h; const uint8_t* data; bool uniform; };

struct OrcConfig {
    double dt;                 /* env.py:202-204 timestep=0.005, frame_skip=1 */
    int solver_iterations;     /* pybullet world default 50 */
    double erp;                /* btContactSolverInfo::m_erp 0.2 */
    double erp2;               /* contact ERP 0.08 (pybullet createEmptyDynamicsWorld) */
    double linear_slop;        /* 1e-5 */
    double warmstart;          /* 0.1 */
    double residual_threshold; /* 1e-7 */
    double lin_damping, ang_damping;   /* btMultiBody 0.04 / 0.04 */
    double max_coord_vel;      /* btMultiBody::m_maxCoordinateVelocity 100 */
    double motor_kp, motor_kd, motor_max_force;   /* pybullet setJointMotorControl2 defaults .1, 1, 1e5 */
    double limit_max_impulse;  /* btMultiBodyConstraint default 100 */
    int use_box_inertia;       /* 1: AABB box inertia for robot links (App. C-2) */
    double gravity;            /* -9.81 (env.py:208) */
    int enable_contacts;
    double contact_breaking;   /* 0.02 */
    double margin;             /* 0.001 per hull */
    double restitution_threshold; /* 0.2 */
    double friction_erp;       /* 0.2 (unused by multibody friction rows) */
    int torsional_friction;    /* 1: spinning / rolling friction rows where the URDFs set the coefficients */
};
void orc_default_config(OrcConfig* c);

struct OrcScene {
    OrcBlob blob;
    int n_objects;
    /* robot tree */
    int parent[ORC_NL], jtype[ORC_NL], dof[ORC_NL];
    V3 axis[ORC_NL], dvec[ORC_NL], evec[ORC_NL];
    M3 rot0[ORC_NL];
    double mass[ORC_NL]; V3 inertia[ORC_NL];
    double damping[ORC_NL], lower[ORC_NL], upper[ORC_NL];
    int link_of_dof[ORC_NDOF];
    V3 base_pos;
    int shape_of_link[ORC_NL + 1];      /* [0] = base link_0 */
    /* table + objects */
    V3 table_pos; Q4 table_quat; V3 table_upper_offset;
    int table_shape[2];
    double obj_mass[ORC_MAXOBJ]; V3 obj_inertia[ORC_MAXOBJ];
    V3 obj_init_pos[ORC_MAXOBJ]; Q4 obj_init_quat[ORC_MAXOBJ];
    int obj_shape[ORC_MAXOBJ];
    /* geometry */
    std::vector<OrcShape> shapes;
    const float* tris;        /* [T][3][8] */
    const double* hull_verts; /* [V][3] */
    std::vector<OrcTexture> textures;
    bool load(const char* path, int n_objects, const OrcConfig& cfg, std::string& err);
};

struct OrcObjState { V3 pos; Q4 quat; V3 lin, ang; };
struct OrcFrame { M3 R; V3 p; };   /* world_from_local rotation and COM position */

#endif
