// rr_api.cu — C ABI of libreal_b200.so (include/real_b200.h): context creation from the asset blob, constant
// set-up (robot tree, eye camera, light), device state allocation, and thin launch wrappers.
#include "rr_ctx.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>

static std::string g_create_err;

// ---------------------------------------------------------------------------------------------- blob reader
namespace {
struct Entry { int dtype, ndim; uint64_t shape[4]; const uint8_t* data; uint64_t nbytes; };
struct Blob {
    std::vector<uint8_t> raw;
    std::map<std::string, Entry> ent;
    bool load(const char* path, std::string& err) {
        FILE* f = fopen(path, "rb");
        if (!f) { err = std::string("cannot open asset blob ") + path; return false; }
        fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
        raw.resize((size_t)n);
        size_t got = fread(raw.data(), 1, (size_t)n, f);
        fclose(f);
        if (got != (size_t)n || n < 16 || memcmp(raw.data(), "RRB200\0\0", 8) != 0) { err = "asset blob: bad magic / short read"; return false; }
        uint32_t ver, cnt; memcpy(&ver, raw.data() + 8, 4); memcpy(&cnt, raw.data() + 12, 4);
        if (ver != 1) { err = "asset blob: unsupported version"; return false; }
        if ((uint64_t)cnt > ((uint64_t)n - 16) / 112) { err = "asset blob: entry table out of range"; return false; }
        static const uint64_t item[5] = {1, 4, 4, 8, 8};      // u8 i32 f32 f64 i64
        for (uint32_t i = 0; i < cnt; i++) {
            const uint8_t* e = raw.data() + 16 + 112 * (size_t)i;
            char nm[49]; memcpy(nm, e, 48); nm[48] = 0;
            Entry en; uint32_t dt, nd; memcpy(&dt, e + 48, 4); memcpy(&nd, e + 52, 4);
            en.dtype = (int)dt; en.ndim = (int)nd; memcpy(en.shape, e + 56, 32);
            uint64_t off; memcpy(&off, e + 88, 8); memcpy(&en.nbytes, e + 96, 8);
            if (off > (uint64_t)n || en.nbytes > (uint64_t)n - off) { err = "asset blob: entry out of range"; return false; }
            uint64_t count = 1;
            for (uint32_t d = 0; d < nd && d < 4; d++) count *= en.shape[d];
            if (dt > 4 || nd > 4 || count * item[dt > 4 ? 0 : dt] != en.nbytes) { err = std::string("asset blob: inconsistent entry ") + nm; return false; }
            en.data = raw.data() + off;
            ent[nm] = en;
        }
        return true;
    }
    const Entry* get(const char* name, int dtype, std::string& err) const {
        auto it = ent.find(name);
        if (it == ent.end() || it->second.dtype != dtype) { err = std::string("asset blob: missing entry ") + name; return nullptr; }
        return &it->second;
    }
};

// float helpers mirroring TinyRenderer's mat templates (dot products accumulate from the last component)
float hdot4r(const float* a, const float* b) { float r = 0.f; r += a[3] * b[3]; r += a[2] * b[2]; r += a[1] * b[1]; r += a[0] * b[0]; return r; }
float hdet(const float* m, int n) {
    if (n == 1) return m[0];
    float ret = 0.f;
    for (int i = n; i--;) {
        float mn[9]; int k = 0;
        for (int r = 1; r < n; r++) for (int c = 0; c < n; c++) if (c != i) mn[k++] = m[r * n + c];
        ret += m[i] * (hdet(mn, n - 1) * (float)((i % 2) ? -1 : 1));
    }
    return ret;
}
float hcof4(const float M[4][4], int row, int col) {
    float mn[9]; int k = 0;
    for (int r = 0; r < 4; r++) if (r != row) for (int c = 0; c < 4; c++) if (c != col) mn[k++] = M[r][c];
    return hdet(mn, 3) * (float)(((row + col) % 2) ? -1 : 1);
}

// pybullet computeViewMatrix / computeProjectionMatrixFOV in float (env.py:543-551), column-major out
void eye_camera(float* view, float* proj) {
    float eye[3] = {0.01f, 0.f, 1.2f}, center[3] = {0.f, 0.f, 0.08f}, up[3] = {0.f, 0.f, 1.f};
    float f[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]};
    float fl = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (int i = 0; i < 3; i++) f[i] = f[i] / fl;
    float ul = sqrtf(up[0] * up[0] + up[1] * up[1] + up[2] * up[2]);
    float u[3] = {up[0] / ul, up[1] / ul, up[2] / ul};
    float s[3] = {f[1] * u[2] - f[2] * u[1], f[2] * u[0] - f[0] * u[2], f[0] * u[1] - f[1] * u[0]};
    float sl = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    for (int i = 0; i < 3; i++) s[i] = s[i] / sl;
    u[0] = s[1] * f[2] - s[2] * f[1]; u[1] = s[2] * f[0] - s[0] * f[2]; u[2] = s[0] * f[1] - s[1] * f[0];
    view[0] = s[0]; view[4] = s[1]; view[8] = s[2];
    view[1] = u[0]; view[5] = u[1]; view[9] = u[2];
    view[2] = -f[0]; view[6] = -f[1]; view[10] = -f[2];
    view[3] = view[7] = view[11] = 0.f;
    view[12] = -(s[0] * eye[0] + s[1] * eye[1] + s[2] * eye[2]);
    view[13] = -(u[0] * eye[0] + u[1] * eye[1] + u[2] * eye[2]);
    view[14] = (f[0] * eye[0] + f[1] * eye[1] + f[2] * eye[2]);
    view[15] = 1.f;
    float fov = 80.f, aspect = (float)(320.0 / 240.0), nearVal = 0.1f, farVal = 100.f;
    float yScale = (float)(1.0 / tan((3.1415926535897932384626433832795029 / 180.0) * fov / 2));
    float xScale = yScale / aspect;
    for (int i = 0; i < 16; i++) proj[i] = 0.f;
    proj[0] = xScale; proj[5] = yScale;
    proj[10] = (nearVal + farVal) / (nearVal - farVal);
    proj[11] = -1.f;
    proj[14] = ((2) * farVal * nearVal) / (nearVal - farVal);
}
}  // namespace

// camera-dependent render constants from column-major view / projection matrices (as pybullet returns them)
void rr_set_camera_consts(RenderConst& C, const float* view, const float* proj) {
    float V[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { C.view[4 * i + j] = view[i + 4 * j]; C.proj[4 * i + j] = proj[i + 4 * j]; V[i][j] = view[i + 4 * j]; }
    C.depthNear = proj[14] / (proj[10] - 1);
    C.depthFar = proj[14] / (proj[10] + 1);
    C.nearPlane = C.proj[4 * 2 + 3] / (C.proj[4 * 2 + 2] - 1);
    C.farPlane = C.proj[4 * 2 + 3] / (C.proj[4 * 2 + 2] + 1);
    // camera position = translation column of viewMatrix.invert() (adjugate / determinant, float)
    float adj[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) adj[i][j] = hcof4(V, i, j);
    float tmp = hdot4r(adj[0], V[0]);
    for (int k = 0; k < 3; k++) C.camP[k] = (double)(adj[3][k] / tmp);
}

#define CKC(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(_e); g_create_err = ctx->err; return -1; } } while (0)

static int fill_constants(rr_ctx* ctx, const Blob& b) {
    std::string& err = ctx->err;
    RobotConst& R = ctx->robot;
    memset(&R, 0, sizeof R);
#define GET(var, name, dt) const Entry* var = b.get(name, dt, err); if (!var) return -1
    GET(e_parent, "robot.parent", 1); GET(e_jtype, "robot.jtype", 1); GET(e_dof, "robot.dof", 1);
    GET(e_axis, "robot.axis", 3); GET(e_rot0, "robot.rot0", 3); GET(e_dvec, "robot.dvec", 3); GET(e_evec, "robot.evec", 3);
    GET(e_mass, "robot.mass", 3); GET(e_inb, "robot.inertia_box", 3); GET(e_inu, "robot.inertia_urdf", 3);
    GET(e_damp, "robot.damping", 3); GET(e_lo, "robot.lower", 3); GET(e_hi, "robot.upper", 3); GET(e_bp, "robot.base_pos", 3);
    GET(e_tp, "table.pos", 3); GET(e_tq, "table.quat", 3); GET(e_tu, "table.upper_offset", 3);
    GET(e_om, "obj.mass", 3); GET(e_oi, "obj.inertia", 3); GET(e_op, "obj.init_pose", 3);
    GET(e_sb, "shape.body", 1); GET(e_sl, "shape.link", 1); GET(e_ts, "shape.tri_start", 1); GET(e_tc, "shape.tri_count", 1);
    GET(e_hs, "shape.hull_start", 1); GET(e_hc, "shape.hull_count", 1);
    GET(e_tex, "shape.tex", 1); GET(e_rgba, "shape.rgba", 2); GET(e_amin, "shape.aabb_min", 3); GET(e_amax, "shape.aabb_max", 3);
    GET(e_tri, "tri.data", 2); GET(e_hv, "hull.verts", 3); GET(e_tm, "tex.meta", 4); GET(e_td, "tex.data", 0);
#undef GET
    if (e_parent->shape[0] != RR_NL || e_sb->shape[0] != RR_NSHAPES) { err = "asset blob: unexpected link/shape count"; return -1; }
    const int32_t* par = (const int32_t*)e_parent->data; const int32_t* jt = (const int32_t*)e_jtype->data; const int32_t* df = (const int32_t*)e_dof->data;
    const double* inertia = (const double*)(ctx->cfg.use_box_inertia ? e_inb->data : e_inu->data);
    for (int i = 0; i < RR_NL; i++) {
        R.parent[i] = par[i]; R.jtype[i] = jt[i]; R.dof[i] = df[i];
        if (df[i] >= 0) R.link_of_dof[df[i]] = i;
        for (int k = 0; k < 3; k++) {
            R.axis[i][k] = ((const double*)e_axis->data)[3 * i + k];
            R.dvec[i][k] = ((const double*)e_dvec->data)[3 * i + k];
            R.evec[i][k] = ((const double*)e_evec->data)[3 * i + k];
            R.inertia[i][k] = inertia[3 * i + k];
        }
        for (int k = 0; k < 9; k++) R.rot0[i][k] = ((const double*)e_rot0->data)[9 * i + k];
        const double* a = R.axis[i]; const double* ev = R.evec[i];
        R.S[i][0] = a[0]; R.S[i][1] = a[1]; R.S[i][2] = a[2];
        R.S[i][3] = a[1] * ev[2] - a[2] * ev[1]; R.S[i][4] = a[2] * ev[0] - a[0] * ev[2]; R.S[i][5] = a[0] * ev[1] - a[1] * ev[0];
        if (jt[i] != 1) for (int k = 0; k < 6; k++) R.S[i][k] = 0.0;
        R.mass[i] = ((const double*)e_mass->data)[i]; R.damping[i] = ((const double*)e_damp->data)[i];
        R.lower[i] = ((const double*)e_lo->data)[i]; R.upper[i] = ((const double*)e_hi->data)[i];
    }
    for (int k = 0; k < 3; k++) { R.base_pos[k] = ((const double*)e_bp->data)[k]; R.table_pos[k] = ((const double*)e_tp->data)[k]; }
    {   // table rotation from its quaternion (btMatrix3x3::setRotation)
        const double* q = (const double*)e_tq->data;
        double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3], s = 2.0 / d;
        double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s, wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
        double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs, yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
        double M[9] = {1.0 - (yy + zz), xy - wz, xz + wy, xy + wz, 1.0 - (xx + zz), yz - wx, xz - wy, yz + wx, 1.0 - (xx + yy)};
        memcpy(R.table_R, M, sizeof M);
        const double* off = (const double*)e_tu->data;
        for (int a = 0; a < 3; a++) R.table_upper_pos[a] = R.table_pos[a] + M[3 * a] * off[0] + M[3 * a + 1] * off[1] + M[3 * a + 2] * off[2];
    }
    for (int o = 0; o < RR_MAX_OBJECTS; o++) {
        R.obj_mass[o] = ((const double*)e_om->data)[o];
        for (int k = 0; k < 3; k++) R.obj_inertia[o][k] = ((const double*)e_oi->data)[3 * o + k];
        for (int k = 0; k < 7; k++) R.obj_init[o][k] = ((const double*)e_op->data)[7 * o + k];
    }
    const Entry* e_fr = b.get("shape.friction", 3, err); if (!e_fr) return -1;
    const Entry* e_re = b.get("shape.restitution", 3, err); if (!e_re) return -1;
    const Entry* e_ro = b.get("shape.rolling", 3, err); if (!e_ro) return -1;
    const Entry* e_sp = b.get("shape.spinning", 3, err); if (!e_sp) return -1;
    for (int s = 0; s < RR_NSHAPES; s++) {
        R.shape_body[s] = ((const int32_t*)e_sb->data)[s]; R.shape_link[s] = ((const int32_t*)e_sl->data)[s];
        R.shape_friction[s] = ((const double*)e_fr->data)[s]; R.shape_restitution[s] = ((const double*)e_re->data)[s];
        R.shape_rolling[s] = ((const double*)e_ro->data)[s]; R.shape_spinning[s] = ((const double*)e_sp->data)[s];
    }
    for (int i = 0; i < RR_NL; i++) {
        int mask = 0;
        for (int k = i; k >= 0; k = R.parent[k]) if (R.jtype[k] == 1) mask |= 1 << R.dof[k];
        R.anc[i] = mask;
    }
    R.n_objects = ctx->n_objects; R.cfg = ctx->cfg;

    RenderConst& C = ctx->render;
    memset(&C, 0, sizeof C);
    for (int s = 0; s < RR_NSHAPES; s++) {
        ShapeRenderConst& sh = C.shape[s];
        sh.tri_start = ((const int32_t*)e_ts->data)[s]; sh.tri_count = ((const int32_t*)e_tc->data)[s];
        sh.tex = ((const int32_t*)e_tex->data)[s];
        sh.objectAndLink = R.shape_body[s] + ((R.shape_link[s] + 1) << 24);
        for (int k = 0; k < 4; k++) sh.rgba[k] = ((const float*)e_rgba->data)[4 * s + k];
        for (int k = 0; k < 3; k++) {
            sh.aabb_min[k] = (float)((const double*)e_amin->data)[3 * s + k] - 1e-4f;
            sh.aabb_max[k] = (float)((const double*)e_amax->data)[3 * s + k] + 1e-4f;
        }
        ctx->hull_start[s] = ((const int32_t*)e_hs->data)[s]; ctx->hull_count[s] = ((const int32_t*)e_hc->data)[s];
        for (int k = 0; k < 3; k++) {
            double lo = ((const double*)e_amin->data)[3 * s + k], hi = ((const double*)e_amax->data)[3 * s + k];
            ctx->aabb_c[s][k] = (lo + hi) * 0.5; ctx->aabb_h[s][k] = (hi - lo) * 0.5;
        }
    }
    int K = (int)e_tm->shape[0];
    if (K > RR_MAX_TEX) { err = "too many textures"; return -1; }
    for (int k = 0; k < K; k++) {
        const int64_t* m = (const int64_t*)e_tm->data + 4 * k;
        C.tex[k].w = (int)m[0]; C.tex[k].h = (int)m[1]; C.tex[k].offset = m[2]; C.tex[k].uniform = (int)m[3];
    }
    float view[16], proj[16]; eye_camera(view, proj);
    rr_set_camera_consts(C, view, proj);
    {   // light direction normalize(-50, 30, 100) in double, stored as float (z-up default)
        double ld[3] = {-50.0, 30.0, 100.0};
        double ll = sqrt(ld[0] * ld[0] + ld[1] * ld[1] + ld[2] * ld[2]);
        for (int k = 0; k < 3; k++) C.light[k] = (float)(ld[k] / ll);
    }
    // device geometry
    ctx->n_tris = (int)e_tri->shape[0];
    CKC(cudaMalloc(&ctx->d_tris, e_tri->nbytes));
    CKC(cudaMemcpy(ctx->d_tris, e_tri->data, e_tri->nbytes, cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&ctx->d_tex, e_td->nbytes));
    CKC(cudaMemcpy(ctx->d_tex, e_td->data, e_td->nbytes, cudaMemcpyHostToDevice));
    CKC(cudaMalloc(&ctx->d_hull_verts, e_hv->nbytes));
    CKC(cudaMemcpy(ctx->d_hull_verts, e_hv->data, e_hv->nbytes, cudaMemcpyHostToDevice));
    return 0;
}

extern "C" {

void rr_config_default(rr_config* c) {
    c->dt = 0.005; c->solver_iterations = 50; c->erp = 0.2; c->erp2 = 0.08; c->linear_slop = 1e-5; c->warmstart = 0.1;
    c->residual_threshold = 1e-7; c->lin_damping = 0.04; c->ang_damping = 0.04; c->max_coord_vel = 100.0;
    c->motor_kp = 0.1; c->motor_kd = 1.0; c->motor_max_force = 100000.0; c->limit_max_impulse = 100.0;
    c->use_box_inertia = 1; c->gravity = -9.81; c->enable_contacts = 1; c->contact_breaking = 0.02; c->margin = 0.001;
    c->restitution_threshold = 0.2; c->friction_erp = 0.2; c->torsional_friction = 1;
}
int rr_config_size(void) { return (int)sizeof(rr_config); }

const char* rr_last_error(rr_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
int rr_n_envs(rr_ctx* ctx) { return ctx->N; }
int rr_n_objects(rr_ctx* ctx) { return ctx->n_objects; }
long long rr_launch_count(rr_ctx* ctx) { return ctx->launches; }

int rr_create(rr_ctx** out, int device, int n_envs, int n_objects, const char* blob_path, const rr_config* cfg) {
    *out = nullptr;
    if (n_envs < 1) { g_create_err = "rr_create: n_envs must be >= 1"; return -1; }
    if (n_objects < 1 || n_objects > RR_MAX_OBJECTS) { g_create_err = "rr_create: n_objects must be 1..3"; return -1; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_create_err = "rr_create: no CUDA device (this library has no CPU fallback)"; return -1; }
    if (device < 0 || device >= ndev) { g_create_err = "rr_create: bad device index"; return -1; }
    rr_ctx* ctx = new rr_ctx();
    ctx->device = device; ctx->N = n_envs; ctx->n_objects = n_objects; ctx->launches = 0; ctx->prof_on = false; ctx->host_compute = nullptr; ctx->host_copy = nullptr; ctx->d_rec_scratch = nullptr; ctx->d_list_scratch = nullptr; ctx->d_work_counter = nullptr; ctx->d_tri_shape = nullptr; ctx->raster_grid = 0; ctx->d_planes = nullptr; ctx->d_cull_scratch = nullptr; ctx->h_collide_const = nullptr; ctx->cull_hmax = 0.f;
    for (int k = 0; k < 2; k++) { ctx->d_cmd[k] = nullptr; ctx->d_mask[k] = nullptr; ctx->d_joint[k] = ctx->d_touch[k] = ctx->d_objpose[k] = nullptr; ctx->d_rgb[k] = nullptr; ctx->d_depth[k] = nullptr; ctx->d_seg[k] = nullptr; }
    ctx->img_capacity = 0; ctx->h_pinned = nullptr; ctx->h_pinned_bytes = 0; ctx->host_begun = ctx->host_ended = 0;
    if (cfg) ctx->cfg = *cfg; else rr_config_default(&ctx->cfg);
    // error path: release whatever was allocated and drop the constant-memory ownership records that point at ctx
    auto fail = [&](int rc) { g_create_err = ctx->err; rr_destroy(ctx); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) { ctx->err = "cudaSetDevice failed"; return fail(-1); }
    Blob b;
    if (!b.load(blob_path, ctx->err)) return fail(-1);
    if (fill_constants(ctx, b) != 0) return fail(-1);
    if (rrk_collide_upload_constants(ctx) != 0) return fail(-1);      // also builds the candidate pair table
    for (int i = 0; i < ctx->n_pairs; i++) { ctx->robot.pair_a[i] = ctx->pair_a[i]; ctx->robot.pair_b[i] = ctx->pair_b[i]; }
    if (rrk_upload_constants(ctx) != 0) return fail(-1);
    if (rrk_render_upload_constants(ctx) != 0) return fail(-1);
    DevState& st = ctx->st;
    st.N = n_envs;
    size_t N = (size_t)n_envs;
    cudaError_t e = cudaSuccess;
    auto alloc = [&](void** p, size_t bytes) { if (e == cudaSuccess) { e = cudaMalloc(p, bytes); if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes); } };
    alloc((void**)&st.q, 11 * N * 8); alloc((void**)&st.qd, 11 * N * 8); alloc((void**)&st.target, 11 * N * 8);
    alloc((void**)&st.obj, 39 * N * 8); alloc((void**)&st.touch, 4 * N * 8); alloc((void**)&st.link_frames, 16 * 12 * N * 8);
    alloc((void**)&st.render_mats, N * RR_NSHAPES * RR_MAT_STRIDE * 4); alloc((void**)&st.diag, N * 4 * 4); alloc((void**)&st.col_flags, N * 4); alloc((void**)&st.ready, N * 4); alloc((void**)&st.solve_pos, N * 4); alloc((void**)&st.solve_order, (size_t)RR_SOLVE_CLASSES * (N + 1) * 4); alloc((void**)&st.timestep, N * 8);
    for (int k = 0; k < 2; k++) {
        alloc((void**)&ctx->d_cmd[k], N * 9 * 8); alloc((void**)&ctx->d_mask[k], N); alloc((void**)&ctx->d_joint[k], N * 9 * 8);
        alloc((void**)&ctx->d_touch[k], N * 4 * 8); alloc((void**)&ctx->d_objpose[k], N * 3 * 7 * 8);
    }
    for (int k = 0; k < 2; k++) {
        alloc((void**)&st.man_hdr[k], N * RR_MAX_MANIFOLDS * 2 * 4);
        alloc((void**)&st.man_pts[k], N * RR_MAX_MANIFOLDS * 4 * RR_PT_STRIDE * 8);
    }
    alloc((void**)&st.handoff, N * RR_IMG_DOUBLES * 8);
    if (e != cudaSuccess) { ctx->err = std::string("rr_create: device allocation failed: ") + cudaGetErrorString(e); return fail(-1); }
    ctx->man_parity = 0;
    { const char* ea = getenv("RR_EAGER_COLLIDE"); ctx->eager_collide = !(ea && ea[0] == '0'); }
    { const char* ea = getenv("RR_HEAVY_SOLVER"); ctx->heavy_mode = (ea && ea[0] == '0') ? 0 : (ea && ea[0] == '1') ? 1 : -1; }
    if (rrk_reset(ctx, nullptr, 0) != 0) return fail(-1);
    if (rrk_render_build_static(ctx) != 0) return fail(-1);
    if (cudaDeviceSynchronize() != cudaSuccess) { ctx->err = "rr_create: initial reset failed"; return fail(-1); }
    *out = ctx;
    return 0;
}

int rr_destroy(rr_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    rrk_physics_unbind(ctx); rrk_collide_unbind(ctx); rrk_render_unbind(ctx);
    if (ctx->side_stream) { cudaStreamDestroy(ctx->side_stream); cudaEventDestroy(ctx->ev_solved); cudaEventDestroy(ctx->ev_collided); cudaEventDestroy(ctx->ev_dyn); cudaEventDestroy(ctx->ev_heavy); cudaEventDestroy(ctx->ev_collided_heavy); cudaStreamDestroy(ctx->side_stream2); }
    DevState& st = ctx->st;
    void* ptrs[] = {st.q, st.qd, st.target, st.obj, st.touch, st.link_frames, st.render_mats, st.diag, st.col_flags, st.ready, st.solve_pos, st.solve_order, st.timestep,
                    st.man_hdr[0], st.man_hdr[1], st.man_pts[0], st.man_pts[1], st.handoff,
                    ctx->d_cmd[0], ctx->d_mask[0], ctx->d_joint[0], ctx->d_touch[0], ctx->d_objpose[0], ctx->d_rgb[0], ctx->d_depth[0], ctx->d_seg[0],
                    ctx->d_cmd[1], ctx->d_mask[1], ctx->d_joint[1], ctx->d_touch[1], ctx->d_objpose[1], ctx->d_rgb[1], ctx->d_depth[1], ctx->d_seg[1],
                    ctx->d_tris, ctx->d_tex, ctx->d_hull_verts, ctx->d_rec_scratch, ctx->d_list_scratch, ctx->d_work_counter, ctx->d_tri_shape, ctx->d_planes, ctx->d_cull_scratch,
                    ctx->d_static_keys, ctx->d_static_img};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_late_timeout) { cudaFreeHost(ctx->h_late_timeout); cudaFree(ctx->d_late_timeout); }
    if (ctx->host_compute) {
        for (int i = 0; i < 16; i++) cudaEventDestroy(ctx->host_ev[i]);
        for (int i = 0; i < 2; i++) cudaEventDestroy(ctx->host_done[i]);
        cudaStreamDestroy(ctx->host_compute); cudaStreamDestroy(ctx->host_copy);
    }
    for (int k = 0; k < RR_PROFILE_KINDS; k++) for (cudaEvent_t e : ctx->prof_ev[k]) cudaEventDestroy(e);
    delete ctx;
    return 0;
}

#define ARG(c, msg) do { if (!(c)) { ctx->err = msg; return -3; } } while (0)

int rr_reset(rr_ctx* ctx, const uint8_t* env_mask, void* stream) { return rrk_reset(ctx, env_mask, (cudaStream_t)stream); }
int rr_set_object_pose(rr_ctx* ctx, const uint8_t* env_mask, int obj, const double* pos_quat, void* stream) {
    ARG(obj >= 0 && obj < ctx->n_objects, "rr_set_object_pose: object index out of range");
    ARG(pos_quat != nullptr, "rr_set_object_pose: pos_quat is NULL");
    return rrk_set_object_pose(ctx, env_mask, obj, pos_quat, (cudaStream_t)stream);
}
int rr_set_state(rr_ctx* ctx, const double* q, const double* qd, const double* obj, void* stream) { return rrk_set_state(ctx, q, qd, obj, (cudaStream_t)stream); }
int rr_get_state(rr_ctx* ctx, double* q, double* qd, double* obj, void* stream) { return rrk_get_state(ctx, q, qd, obj, (cudaStream_t)stream); }
int rr_step(rr_ctx* ctx, const double* joint_cmd, void* stream) {
    ARG(joint_cmd != nullptr, "rr_step: joint_cmd is NULL");
    return rrk_step(ctx, joint_cmd, (cudaStream_t)stream);
}
int rr_read_state(rr_ctx* ctx, double* joint_pos, double* touch, double* obj_pos_quat, void* stream) {
    return rrk_read_state(ctx, joint_pos, touch, obj_pos_quat, (cudaStream_t)stream);
}
int rr_render(rr_ctx* ctx, const uint8_t* env_mask, uint8_t* rgb, float* depth, int32_t* mask, void* stream) {
    return rrk_render(ctx, env_mask, nullptr, rgb, depth, mask, (cudaStream_t)stream);
}
int rr_render_frames(rr_ctx* ctx, const double* frames, uint8_t* rgb, float* depth, int32_t* mask, void* stream) {
    ARG(frames != nullptr, "rr_render_frames: frames is NULL");
    return rrk_render(ctx, nullptr, frames, rgb, depth, mask, (cudaStream_t)stream);
}
int rr_render_camera(rr_ctx* ctx, const float* view16_host, const float* proj16_host, const uint8_t* env_mask,
                     uint8_t* rgb, float* depth, int32_t* mask, void* stream) {
    ARG(view16_host && proj16_host, "rr_render_camera: NULL matrix");
    RenderConst C = ctx->render;
    rr_set_camera_consts(C, view16_host, proj16_host);
    return rrk_render_camera(ctx, &C, env_mask, rgb, depth, mask, (cudaStream_t)stream);
}
int rr_shape_frames(rr_ctx* ctx, double* frames, void* stream) { return rrk_shape_frames(ctx, frames, (cudaStream_t)stream); }
int rr_ik(rr_ctx* ctx, const uint8_t* env_mask, const double* pos_quat, double* joints_out, int max_iter, double residual, void* stream) {
    ARG(pos_quat && joints_out, "rr_ik: NULL argument");
    return rrk_ik(ctx, env_mask, pos_quat, joints_out, max_iter, residual, (cudaStream_t)stream);
}
int rr_link_pose(rr_ctx* ctx, int link_index, double* pos_quat, void* stream) {
    ARG(link_index >= -1 && link_index < RR_NL, "rr_link_pose: link index out of range");
    return rrk_link_pose(ctx, link_index, pos_quat, (cudaStream_t)stream);
}
int rr_score(rr_ctx* ctx, const double* goal_xyz, double* score, void* stream) {
    ARG(goal_xyz && score, "rr_score: NULL argument");
    return rrk_score(ctx, goal_xyz, score, (cudaStream_t)stream);
}
int rr_contacts(rr_ctx* ctx, int32_t* hdr, double* pts, int32_t* pairs_host, int* n_pairs_host, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (rrk_join_heavy(ctx, s) != 0) return -1;
    const int p = ctx->man_parity;
    const size_t N = (size_t)ctx->N;
    cudaError_t e = cudaSuccess;
    if (hdr) e = cudaMemcpyAsync(hdr, ctx->st.man_hdr[p], N * RR_MAX_MANIFOLDS * 2 * 4, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && pts) e = cudaMemcpyAsync(pts, ctx->st.man_pts[p], N * RR_MAX_MANIFOLDS * 4 * RR_PT_STRIDE * 8, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) { ctx->err = std::string("rr_contacts: ") + cudaGetErrorString(e); return -1; }
    if (pairs_host) for (int i = 0; i < ctx->n_pairs; i++) { pairs_host[2 * i] = ctx->pair_a[i]; pairs_host[2 * i + 1] = ctx->pair_b[i]; }
    if (n_pairs_host) *n_pairs_host = ctx->n_pairs;
    return 0;
}
int rr_set_contacts(rr_ctx* ctx, const int32_t* hdr, const double* pts, void* stream) {
    ARG(hdr != nullptr && pts != nullptr, "rr_set_contacts: hdr / pts is NULL");
    cudaStream_t s = (cudaStream_t)stream;
    if (rrk_drop_collide_ahead(ctx, s) != 0) return -1;
    const int p = ctx->man_parity;
    const size_t N = (size_t)ctx->N;
    cudaError_t e = cudaMemcpyAsync(ctx->st.man_hdr[p], hdr, N * RR_MAX_MANIFOLDS * 2 * 4, cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->st.man_pts[p], pts, N * RR_MAX_MANIFOLDS * 4 * RR_PT_STRIDE * 8, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) { ctx->err = std::string("rr_set_contacts: ") + cudaGetErrorString(e); return -1; }
    return 0;
}
int rr_profile(rr_ctx* ctx, int enable) {
    for (int k = 0; k < RR_PROFILE_KINDS; k++) { for (cudaEvent_t e : ctx->prof_ev[k]) cudaEventDestroy(e); ctx->prof_ev[k].clear(); }
    ctx->prof_on = enable != 0;
    return 0;
}
int rr_profile_read(rr_ctx* ctx, double* ms_total, long long* launches) {
    cudaSetDevice(ctx->device);
    if (cudaDeviceSynchronize() != cudaSuccess) { ctx->err = "rr_profile_read: device synchronize failed"; return -1; }
    for (int k = 0; k < RR_PROFILE_KINDS; k++) {
        double tot = 0; long long n = 0;
        std::vector<cudaEvent_t>& ev = ctx->prof_ev[k];
        for (size_t i = 0; i + 1 < ev.size(); i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) { tot += ms; n++; }
        }
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        ev.clear();
        ms_total[k] = tot; launches[k] = n;
    }
    return 0;
}
int rr_diag(rr_ctx* ctx, int32_t* out, void* stream) {
    if (rrk_join_heavy(ctx, (cudaStream_t)stream) != 0) return -1;
    cudaError_t e = cudaMemcpyAsync(out, ctx->st.diag, (size_t)ctx->N * 16, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); return -1; }
    return 0;
}

#define CKH(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(_e); return -1; } } while (0)

// Host-buffer path.  rr_step_host_begin enqueues one whole step — H2D of the commands, physics, read-out, render in chunks
// whose D2H copies (on a second stream) overlap the render of the next chunk — and returns; rr_step_host_end waits for the
// oldest step begun.  Two steps may be in flight: the device staging (commands, observations, images) is double buffered,
// so the physics and the first render chunks of step t+1 run while the copy engine still drains step t's images.
int rr_step_host_begin(rr_ctx* ctx, const double* cmd_h, const uint8_t* render_mask_h, double* joint_h, double* touch_h, double* objpose_h,
                       uint8_t* rgb_h, float* depth_h, int32_t* mask_h) {
    ARG(cmd_h != nullptr, "rr_step_host: joint_cmd_host is NULL");
    ARG(ctx->host_begun - ctx->host_ended < 2, "rr_step_host_begin: two steps are already in flight (call rr_step_host_end first)");
    CKH(cudaSetDevice(ctx->device));
    const size_t N = (size_t)ctx->N, PX = (size_t)RR_W * RR_H;
    const bool want_img = render_mask_h && (rgb_h || depth_h || mask_h);
    if (!ctx->host_compute) {
        CKH(cudaStreamCreateWithFlags(&ctx->host_compute, cudaStreamNonBlocking));
        CKH(cudaStreamCreateWithFlags(&ctx->host_copy, cudaStreamNonBlocking));
        for (int i = 0; i < 16; i++) CKH(cudaEventCreateWithFlags(&ctx->host_ev[i], cudaEventDisableTiming));
        for (int i = 0; i < 2; i++) CKH(cudaEventCreateWithFlags(&ctx->host_done[i], cudaEventDisableTiming));
    }
    if (want_img && ctx->img_capacity < N) {
        for (int k = 0; k < 2; k++) { CKH(cudaMalloc(&ctx->d_rgb[k], N * PX * 3)); CKH(cudaMalloc(&ctx->d_depth[k], N * PX * 4)); CKH(cudaMalloc(&ctx->d_seg[k], N * PX * 4)); }
        ctx->img_capacity = N;
    }
    if (ctx->host_begun == ctx->host_ended) CKH(cudaDeviceSynchronize());      // nothing in flight: order after whatever the caller enqueued on other streams
    const int k = ctx->host_begun & 1;
    cudaStream_t s = ctx->host_compute, cs = ctx->host_copy;
    CKH(cudaMemcpyAsync(ctx->d_cmd[k], cmd_h, N * 9 * 8, cudaMemcpyHostToDevice, s));
    if (want_img) CKH(cudaMemcpyAsync(ctx->d_mask[k], render_mask_h, N, cudaMemcpyHostToDevice, s));
    if (rrk_step(ctx, ctx->d_cmd[k], s) != 0) return -1;
    if (rrk_read_state(ctx, ctx->d_joint[k], ctx->d_touch[k], ctx->d_objpose[k], s) != 0) return -1;
    if (joint_h) CKH(cudaMemcpyAsync(joint_h, ctx->d_joint[k], N * 9 * 8, cudaMemcpyDeviceToHost, s));
    if (touch_h) CKH(cudaMemcpyAsync(touch_h, ctx->d_touch[k], N * 4 * 8, cudaMemcpyDeviceToHost, s));
    if (objpose_h) CKH(cudaMemcpyAsync(objpose_h, ctx->d_objpose[k], N * ctx->n_objects * 7 * 8, cudaMemcpyDeviceToHost, s));
    if (want_img) {
        // render in chunks of envs; the device->host copy of chunk c runs on the copy stream while chunk c+1 renders
        const int chunks = N >= 1024 ? 16 : (N >= 64 ? 4 : 1);
        const int per = (int)((N + chunks - 1) / chunks);
        for (int c = 0; c < chunks; c++) {
            const int lo = c * per, cnt = (int)N - lo < per ? (int)N - lo : per;
            if (cnt <= 0) break;
            if (rrk_render_range(ctx, ctx->d_mask[k], nullptr, ctx->d_rgb[k], ctx->d_depth[k], mask_h ? ctx->d_seg[k] : nullptr, lo, cnt, s) != 0) return -1;
            CKH(cudaEventRecord(ctx->host_ev[c], s));
            CKH(cudaStreamWaitEvent(cs, ctx->host_ev[c], 0));
            const size_t o = (size_t)lo * PX, n = (size_t)cnt * PX;
            if (rgb_h) CKH(cudaMemcpyAsync(rgb_h + o * 3, ctx->d_rgb[k] + o * 3, n * 3, cudaMemcpyDeviceToHost, cs));
            if (depth_h) CKH(cudaMemcpyAsync(depth_h + o, ctx->d_depth[k] + o, n * 4, cudaMemcpyDeviceToHost, cs));
            if (mask_h) CKH(cudaMemcpyAsync(mask_h + o, ctx->d_seg[k] + o, n * 4, cudaMemcpyDeviceToHost, cs));
        }
    }
    // the step is done when the compute stream (small copies) and the copy stream (images) both are
    CKH(cudaEventRecord(ctx->host_ev[15], s));
    CKH(cudaStreamWaitEvent(cs, ctx->host_ev[15], 0));
    CKH(cudaEventRecord(ctx->host_done[k], cs));
    ctx->host_begun++;
    return 0;
}
int rr_step_host_end(rr_ctx* ctx) {
    ARG(ctx->host_begun > ctx->host_ended, "rr_step_host_end: no step in flight");
    CKH(cudaSetDevice(ctx->device));
    CKH(cudaEventSynchronize(ctx->host_done[ctx->host_ended & 1]));
    ctx->host_ended++;
    return 0;
}
int rr_step_host(rr_ctx* ctx, const double* cmd_h, const uint8_t* render_mask_h, double* joint_h, double* touch_h, double* objpose_h,
                 uint8_t* rgb_h, float* depth_h, int32_t* mask_h) {
    while (ctx->host_begun > ctx->host_ended) if (rr_step_host_end(ctx) != 0) return -1;
    if (rr_step_host_begin(ctx, cmd_h, render_mask_h, joint_h, touch_h, objpose_h, rgb_h, depth_h, mask_h) != 0) return -1;
    return rr_step_host_end(ctx);
}

}  // extern "C"
