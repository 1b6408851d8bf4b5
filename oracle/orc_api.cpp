/* TEST INFRASTRUCTURE — CPU oracle: C entry points (loaded with ctypes by tests/, smoke() and
 * bench.py's cpu_baseline leg ONLY; the product path never links or loads this). */
#include "orc_env.h"
#include "orc_render.h"
#include "orc_ik.h"
#include <cstdio>
#include <cstring>
#include <string>

struct OrcHandle {
    OrcScene scene;
    OrcConfig cfg;
    OrcEnv env;
    std::string err;
};

extern "C" {

void orc_config_default(OrcConfig* c) { orc_default_config(c); }
int orc_config_size() { return (int)sizeof(OrcConfig); }

void* orc_create(const char* blob_path, int n_objects, const OrcConfig* cfg) {
    OrcHandle* h = new OrcHandle();
    if (cfg) h->cfg = *cfg; else orc_default_config(&h->cfg);
    if (!h->scene.load(blob_path, n_objects, h->cfg, h->err)) {
        std::fprintf(stderr, "orc_create: %s\n", h->err.c_str());
        delete h; return nullptr;
    }
    orc_env_init(&h->env, &h->scene, &h->cfg);
    return h;
}
void orc_destroy(void* p) { delete (OrcHandle*)p; }
void orc_reset(void* p) { orc_env_reset(&((OrcHandle*)p)->env); }

/* state layout: q[11], qd[11], obj[3][13] = pos3 quat4(xyzw) lin3 ang3 */
void orc_get_state(void* p, double* q, double* qd, double* obj) {
    OrcEnv& e = ((OrcHandle*)p)->env;
    for (int i = 0; i < ORC_NDOF; i++) { q[i] = e.q[i]; qd[i] = e.qd[i]; }
    for (int o = 0; o < ORC_MAXOBJ; o++) {
        double* d = obj + 13 * o; const OrcObjState& s = e.obj[o];
        d[0] = s.pos.x; d[1] = s.pos.y; d[2] = s.pos.z;
        d[3] = s.quat.x; d[4] = s.quat.y; d[5] = s.quat.z; d[6] = s.quat.w;
        d[7] = s.lin.x; d[8] = s.lin.y; d[9] = s.lin.z;
        d[10] = s.ang.x; d[11] = s.ang.y; d[12] = s.ang.z;
    }
}
void orc_set_state(void* p, const double* q, const double* qd, const double* obj) {
    OrcEnv& e = ((OrcHandle*)p)->env;
    if (q) for (int i = 0; i < ORC_NDOF; i++) e.q[i] = q[i];
    if (qd) for (int i = 0; i < ORC_NDOF; i++) e.qd[i] = qd[i];
    if (obj) for (int o = 0; o < ORC_MAXOBJ; o++) {
        const double* d = obj + 13 * o; OrcObjState& s = e.obj[o];
        s.pos = V3(d[0], d[1], d[2]); s.quat = Q4(d[3], d[4], d[5], d[6]);
        s.lin = V3(d[7], d[8], d[9]); s.ang = V3(d[10], d[11], d[12]);
    }
    e.n_manifolds = 0;
    orc_fk(&e);
}
void orc_set_targets(void* p, const double* t11) { OrcEnv& e = ((OrcHandle*)p)->env; for (int i = 0; i < ORC_NDOF; i++) e.target[i] = t11[i]; }
void orc_get_targets(void* p, double* t11) { OrcEnv& e = ((OrcHandle*)p)->env; for (int i = 0; i < ORC_NDOF; i++) t11[i] = e.target[i]; }
void orc_step(void* p, const double* joint_cmd9) { orc_env_step(&((OrcHandle*)p)->env, joint_cmd9); }
void orc_step_sim_only(void* p) { orc_step_simulation(&((OrcHandle*)p)->env); }
void orc_read(void* p, double* joints9, double* touch4) {
    OrcEnv& e = ((OrcHandle*)p)->env;
    orc_calc_state(&e, joints9);
    for (int i = 0; i < 4; i++) touch4[i] = e.touch[i];
}
/* COM-frame pose of robot link (index -1 = base link_0 .. 15), xyz + quat xyzw: BodyPart.get_pose() */
void orc_link_pose(void* p, int link, double* out7) {
    OrcEnv& e = ((OrcHandle*)p)->env;
    OrcFrame f; if (link < 0) { f.R = M3::identity(); f.p = e.sc->base_pos; } else f = e.link[link];
    Q4 q = mat_to_quat(f.R);
    out7[0] = f.p.x; out7[1] = f.p.y; out7[2] = f.p.z; out7[3] = q.x; out7[4] = q.y; out7[5] = q.z; out7[6] = q.w;
}
void orc_aba_qdd(void* p, const double* tau, double* qdd) { orc_aba(&((OrcHandle*)p)->env, tau, qdd); }
void orc_minv(void* p, double* Minv) { orc_mass_matrix(&((OrcHandle*)p)->env, Minv); }
void orc_diag(void* p, int* rows, int* iters, int* contacts, double* residual) {
    OrcEnv& e = ((OrcHandle*)p)->env;
    *rows = e.last_rows; *iters = e.last_iterations; *contacts = e.last_contacts; *residual = e.last_residual;
}
void orc_diag2(void* p, int* overlaps, int* manifolds, int* manifold_overflow) {
    OrcEnv& e = ((OrcHandle*)p)->env;
    *overlaps = e.last_overlaps; *manifolds = e.n_manifolds; *manifold_overflow = e.manifold_overflow;
}
int orc_get_contacts(void* p, int max_pts, int* shapeA, int* shapeB, double* data /* [n][11]: worldA3 worldB3 normal3 dist impulse */) {
    OrcEnv& e = ((OrcHandle*)p)->env;
    int n = 0;
    for (int m = 0; m < e.n_manifolds; m++) for (int k = 0; k < e.manifolds[m].n && n < max_pts; k++) {
        const OrcContactPoint& c = e.manifolds[m].pt[k];
        shapeA[n] = e.manifolds[m].shapeA; shapeB[n] = e.manifolds[m].shapeB;
        double* d = data + 11 * n;
        d[0] = c.worldA.x; d[1] = c.worldA.y; d[2] = c.worldA.z; d[3] = c.worldB.x; d[4] = c.worldB.y; d[5] = c.worldB.z;
        d[6] = c.normalB.x; d[7] = c.normalB.y; d[8] = c.normalB.z; d[9] = c.distance; d[10] = c.applied_impulse;
        n++;
    }
    return n;
}

/* retina render from the current state (env.py:249-255, 536-567): rgb u8[240][320][3], depth f32[240][320], mask i32[240][320] */
void orc_render(void* p, uint8_t* rgb, float* depth, int32_t* mask) {
    OrcHandle* h = (OrcHandle*)p;
    orc_render_eye(&h->env, rgb, depth, mask);
}
/* render from explicit frames: frames[S][12] = world_from_local R (row-major 9) + p (3), as doubles */
void orc_render_frames(void* p, const double* frames, const float* view16, const float* proj16, uint8_t* rgb, float* depth, int32_t* mask) {
    OrcHandle* h = (OrcHandle*)p;
    orc_render_with(&h->scene, frames, view16, proj16, rgb, depth, mask);
}
void orc_shape_frames(void* p, double* frames /* [S][12] */) {
    OrcHandle* h = (OrcHandle*)p;
    int S = (int)h->scene.shapes.size();
    for (int s = 0; s < S; s++) {
        OrcFrame f = orc_shape_frame(&h->env, s);
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) frames[12 * s + 3 * a + b] = f.R.m[a][b];
        frames[12 * s + 9] = f.p.x; frames[12 * s + 10] = f.p.y; frames[12 * s + 11] = f.p.z;
    }
}
void orc_eye_matrices(float* view16, float* proj16) { orc_eye_camera(view16, proj16); }

int orc_ik(void* p, const double* pos3, const double* quat4, int max_iter, double residual, double* out11) {
    return orc_ik_solve(&((OrcHandle*)p)->env, pos3, quat4, max_iter, residual, out11);
}

}  /* extern "C" */

bool orc_closest_points_shapes(const OrcScene* sc, int sa, const OrcFrame& fa, int sb, const OrcFrame& fb, double* out10);
extern "C" int orc_test_closest(void* p, int sa, const double* fa12, int sb, const double* fb12, double* out10) {
    OrcHandle* h = (OrcHandle*)p;
    OrcFrame A, B;
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { A.R.m[a][b] = fa12[3 * a + b]; B.R.m[a][b] = fb12[3 * a + b]; }
    A.p = V3(fa12[9], fa12[10], fa12[11]); B.p = V3(fb12[9], fb12[10], fb12[11]);
    return orc_closest_points_shapes(&h->scene, sa, A, sb, B, out10) ? 1 : 0;
}
