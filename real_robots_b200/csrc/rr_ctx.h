// Internal context of libreal_b200.so (not part of the C ABI).
#pragma once
#ifdef RR_SIMT_EMU_BUILD
// tests/simt: the same sources compiled with g++ against a SIMT emulator (test infrastructure, never shipped or loaded by the package)
#include "simt_emu.h"
#define RR_LAUNCH(kern, grid, block, smem, stream) (simt::Launcher((int)(grid), (int)(block), (size_t)(smem)) << [&](auto&&... a) { kern(a...); })
#define RR_DYN_SMEM(name) unsigned char* name = simt::g_dyn_smem
#else
#include <cuda_runtime.h>
#define RR_LAUNCH(kern, grid, block, smem, stream) kern<<<(grid), (block), (smem), (stream)>>>
#define RR_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/real_b200.h"

#define RR_NSHAPES 22
#define RR_NL RR_NUM_LINKS
#define RR_NDOF RR_NUM_DOF
#define RR_W RR_EYE_W
#define RR_H RR_EYE_H
#define RR_FRAME_STRIDE (RR_NSHAPES * 12)
#define RR_MAX_TEX 16
#define RR_IMG_DOUBLES 16384     // doubles per env of the solver image in global memory (head + slot schedule + row blocks; static_assert in rr_physics.cu)
#define RR_MAT_STRIDE 44       // floats per (env, shape) in render_mats
#define RR_MAX_PAIRS 96          // candidate collision pairs (92 for 3 objects)
#define RR_MAX_MANIFOLDS RR_MAX_MANIFOLDS_PER_ENV      // cached non-empty manifolds per env (<= 32: one lane per slot)
#define RR_PT_STRIDE 17          // doubles per contact point: localA3 localB3 normalB3 dist impulse worldA3 worldB3
#define RR_MAX_CONTACTS 48       // contact points per env that get solver rows
// Solver row ids.  Single rows (phase A of an iteration): joint limits + motors, contact normals, spinning, rolling; then the
// friction pairs (phase B, addressed by contact index).
#define RR_NC_ROWS 24                                        // <= 11 limit rows + 11 motor rows
#define RID_NORMAL RR_NC_ROWS                                // + c
#define RID_SPIN (RID_NORMAL + RR_MAX_CONTACTS)              // + c
#define RID_ROLL (RID_SPIN + RR_MAX_CONTACTS)                // + 2 c + {0, 1}
#define RR_ROWS_A (RID_ROLL + 2 * RR_MAX_CONTACTS)           // 216 single rows
#define RID_FRIC RR_ROWS_A                                   // + 2 c + {0, 1}
#define RR_MAX_ROWS (RID_FRIC + 2 * RR_MAX_CONTACTS)         // 312
#define RR_SOLVE_CLASSES 8     // cost classes of the solver's env order: schedule length above 64, 56, 48, 40, 32, 28, 24 slots per iteration, the rest
#define RR_IDLE 0xff

// ---- constant robot / scene description (uploaded to __constant__ memory once per process+device)
struct RobotConst {
    int parent[RR_NL], jtype[RR_NL], dof[RR_NL], link_of_dof[RR_NDOF];
    double axis[RR_NL][3], dvec[RR_NL][3], evec[RR_NL][3], rot0[RR_NL][9];
    double S[RR_NL][6];          // joint motion axis in the link COM frame: (axis, axis x evec)
    double mass[RR_NL], inertia[RR_NL][3], damping[RR_NL], lower[RR_NL], upper[RR_NL];
    double base_pos[3];
    double table_R[9], table_pos[3], table_upper_pos[3];
    double obj_mass[RR_MAX_OBJECTS], obj_inertia[RR_MAX_OBJECTS][3], obj_init[RR_MAX_OBJECTS][7];
    int shape_body[RR_NSHAPES], shape_link[RR_NSHAPES];
    int n_objects;
    rr_config cfg;
    int anc[RR_NL];              // bit d set: dof d lies on the path from link to the base (incl. the link's own joint)
    int pair_a[RR_MAX_PAIRS], pair_b[RR_MAX_PAIRS];
    double shape_friction[RR_NSHAPES], shape_restitution[RR_NSHAPES], shape_rolling[RR_NSHAPES], shape_spinning[RR_NSHAPES];
};

struct TexConst { int w, h, uniform; long long offset; };
struct ShapeRenderConst {
    int tri_start, tri_count, tex, objectAndLink;
    float rgba[4];
    float aabb_min[3], aabb_max[3];
};
struct RenderConst {
    ShapeRenderConst shape[RR_NSHAPES];
    TexConst tex[RR_MAX_TEX];
    float view[16], proj[16];        // row-major 4x4 (TinyRenderer's Matrix[i][j])
    float light[3];
    float nearPlane, farPlane;       // shader planes
    float depthNear, depthFar;       // planes used by the depth output conversion
    double camP[3];                  // camera position used by back-face culling
    int n_shapes_active;
};

// ---- device state (structure of arrays; index [field][env] so a warp-wide access over envs is coalesced)
struct DevState {
    int N;
    double* q;        // [11][N]
    double* qd;       // [11][N]
    double* target;   // [11][N]
    double* obj;      // [3][13][N]   pos3 quat4 lin3 ang3
    double* touch;    // [4][N]
    double* link_frames;  // [16][12][N]  world_from_link R(9) + p(3), COM frames
    float*  render_mats;  // [N][22][40] per-shape matrices for the rasteriser (see rr_render.cu)
    int*    diag;     // [N][4]
    int*    solve_order; // [RR_SOLVE_CLASSES + RR_SOLVE_CLASSES * N]: per cost class a counter and the envs of that class (k_step_dyn appends,
                        // k_step_solve's warp w takes the w-th env of the concatenation: the longest sweeps start first)
    int*    solve_pos;  // [N] (class << 24) | position in the class list: where solve_order has the env
    int*    ready;      // [N] serial number of the last step whose state k_step_solve_heavy finished for this env (what a rasteriser
                        // launched before that kernel ended waits for, see rrk_render)
    int*    col_flags;  // [N] written by k_collide (1: a non-empty manifold found no slot), merged into diag by k_step_solve
    long long* timestep;  // [N]
    int*    man_hdr[2];   // [N][RR_MAX_MANIFOLDS][2]  (pair index or -1, point count), double buffered
    double* man_pts[2];   // [N][RR_MAX_MANIFOLDS][4][RR_PT_STRIDE]
    double* handoff;      // [N][RR_IMG_DOUBLES] per-env solver image passed from k_step_dyn to k_step_solve (only its used head is touched)
};

// ---- the solver's env order (DevState::solve_order), shared by the kernels of rr_physics.cu, rr_collide.cu and rr_render.cu
// the idx-th env in order of decreasing cost class (every env is in exactly one class list)
__device__ __forceinline__ int solve_pick(const DevState& st, int idx) {
    int cls = 0;
#pragma unroll
    for (int k = 0; k < RR_SOLVE_CLASSES - 1; k++) { const int n = st.solve_order[k]; if (cls == k && idx >= n) { idx -= n; cls = k + 1; } }
    return st.solve_order[RR_SOLVE_CLASSES + cls * st.N + idx];
}
// how many envs the heavy kernel takes: the classes above the last one, as far as its grid goes
__device__ __forceinline__ int solve_n_heavy(const DevState& st, int heavy_cap) {
    int n = 0;
#pragma unroll
    for (int k = 0; k < RR_SOLVE_CLASSES - 1; k++) n += st.solve_order[k];
    return min(n, heavy_cap);
}
// is env e one of them?
__device__ __forceinline__ bool solve_is_heavy(const DevState& st, int e, int heavy_cap) {
    const int cp = st.solve_pos[e], cls = cp >> 24;
    if (cls >= RR_SOLVE_CLASSES - 1) return false;
    int idx = cp & 0xffffff;
    for (int k = 0; k < cls; k++) idx += st.solve_order[k];
    return idx < heavy_cap;
}


struct rr_ctx {
    int device;
    int N, n_objects;
    rr_config cfg;
    DevState st;
    RobotConst robot;          // host copy
    RenderConst render;        // host copy
    // device geometry
    float* d_tris;             // [T][3][8]
    uint8_t* d_tex;
    double* d_hull_verts;      // [V][3]
    int hull_start[RR_NSHAPES], hull_count[RR_NSHAPES];
    double aabb_c[RR_NSHAPES][3], aabb_h[RR_NSHAPES][3];
    int pair_a[RR_MAX_PAIRS], pair_b[RR_MAX_PAIRS], n_pairs;
    int man_parity;            // which manifold buffer holds the current contacts
    // look-ahead narrowphase: the contacts of step t+1 depend on the poses after step t only, so rrk_step launches k_collide for
    // the NEXT step on a side stream right after the solver (it then overlaps whatever the caller does next, normally the
    // rasteriser); the next rrk_step waits for it and flips man_parity.  Anything that edits poses or manifolds drops it.
    cudaStream_t side_stream = nullptr; cudaEvent_t ev_solved = nullptr, ev_collided = nullptr, ev_dyn = nullptr, ev_heavy = nullptr, ev_collided_heavy = nullptr;
    cudaStream_t side_stream2 = nullptr;   // the look-ahead narrowphase of the main solver kernel's envs (the heavy kernel's envs follow that kernel on side_stream)
    bool collide_ahead_split = false;
    bool collide_ahead = false, eager_collide = true;
    int heavy_cap = 0;         // envs k_step_solve_heavy took at most in the last step
    int* d_late_timeout = nullptr; int* h_late_timeout = nullptr;   // k_raster's bounded wait for that kernel ran out (pinned mirror)
    int step_serial = 0;       // rrk_step calls so far (the ticket k_step_solve_heavy leaves in DevState::ready)
    bool heavy_pending = false; // k_step_solve_heavy of the last step has not been joined into the caller's stream yet
    // The heavy kernel pays when a rasteriser follows the step and hides it; back-to-back physics steps are ~0.6 ms faster with every env
    // in the main kernel (the heavy envs' solve -> narrowphase chain is then overlapped by the bulk instead of serialised behind it).
    // heavy_mode -1 (default): split the step iff the previous step was followed FIRST by an eye-camera render of all envs (a caller that
    // reads the state before it renders joins the heavy kernel before the rasteriser can hide it); 0 / 1: never / always
    int heavy_mode = -1; bool render_followed = true, rendered_since_step = false, touched_since_step = false;
    int n_tris;
    void* d_rec_scratch; void* d_list_scratch; int* d_work_counter; uint8_t* d_tri_shape; int raster_grid;   // persistent rasteriser
    void* h_collide_const; float cull_hmax;   // host copies of per-context __constant__ data that are built on the fly
    void* d_planes; void* d_cull_scratch;   // model-space triangle planes (conservative back-face pre-test), per-CTA survivor lists
    // static layer of the eye camera's image (table + robot base never move): raster keys per pixel and the finished static image,
    // built once by k_raster<1> and used as the z-buffer's initial content by k_raster<2> (see rr_render.cu)
    void* d_static_keys = nullptr; void* d_static_img = nullptr; bool static_ready = false; bool static_enabled = true;
    unsigned static_bits = 0; int static_skip_front = 0;
    // staging for the host-buffer path
    // two staging sets: step t+1 may be begun (physics, render into the other image set) while step t's copies still run
    double* d_cmd[2]; uint8_t* d_mask[2]; double* d_joint[2]; double* d_touch[2]; double* d_objpose[2];
    uint8_t* d_rgb[2]; float* d_depth[2]; int32_t* d_seg[2];
    size_t img_capacity;       // envs for which d_rgb/d_depth/d_seg are allocated
    cudaStream_t host_compute, host_copy; cudaEvent_t host_ev[16];   // rr_step_host: render chunk c overlaps the D2H copy of chunk c-1
    cudaEvent_t host_done[2]; int host_begun, host_ended;            // steps begun / ended through rr_step_host_begin / _end
    double* h_pinned; size_t h_pinned_bytes;
    long long launches;
    std::string err;
    // optional event profiling
    bool prof_on;
    std::vector<cudaEvent_t> prof_ev[RR_PROFILE_KINDS];   // start/stop pairs
};

// RAII-free helper: records an event pair around a launch when profiling is on
struct ProfScope {
    rr_ctx* ctx; int kind; cudaStream_t s; cudaEvent_t stop;
    ProfScope(rr_ctx* c, int k, cudaStream_t st) : ctx(c), kind(k), s(st), stop(nullptr) {
        if (!ctx->prof_on || ctx->prof_ev[k].size() >= 16384) return;
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess) return;
        if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return; }
        cudaEventRecord(a, s);
        ctx->prof_ev[k].push_back(a); ctx->prof_ev[k].push_back(b);
        stop = b;
    }
    ~ProfScope() { if (stop) cudaEventRecord(stop, s); }
};

// kernels / launchers (defined in the .cu files)
int rrk_upload_constants(rr_ctx* ctx);
// __constant__ symbols are per device, contexts are not: every launcher re-binds its translation unit's constants when
// another context used them last (stream-ordered copy); rrk_*_unbind forgets a context that is being destroyed.
void rrk_physics_unbind(rr_ctx* ctx);
void rrk_collide_unbind(rr_ctx* ctx);
void rrk_render_unbind(rr_ctx* ctx);
int rrk_reset(rr_ctx* ctx, const uint8_t* mask, cudaStream_t s);
int rrk_set_object_pose(rr_ctx* ctx, const uint8_t* mask, int obj, const double* pq, cudaStream_t s);
int rrk_set_state(rr_ctx* ctx, const double* q, const double* qd, const double* obj, cudaStream_t s);
int rrk_get_state(rr_ctx* ctx, double* q, double* qd, double* obj, cudaStream_t s);
int rrk_step(rr_ctx* ctx, const double* cmd, cudaStream_t s);
int rrk_read_state(rr_ctx* ctx, double* joint, double* touch, double* objpose, cudaStream_t s);
int rrk_link_pose(rr_ctx* ctx, int link, double* out, cudaStream_t s);
int rrk_score(rr_ctx* ctx, const double* goal, double* score, cudaStream_t s);
int rrk_shape_frames(rr_ctx* ctx, double* frames, cudaStream_t s);
int rrk_render_upload_constants(rr_ctx* ctx);
int rrk_render_build_static(rr_ctx* ctx);
int rrk_render(rr_ctx* ctx, const uint8_t* mask, const double* frames_or_null, uint8_t* rgb, float* depth, int32_t* seg, cudaStream_t s);
int rrk_render_range(rr_ctx* ctx, const uint8_t* mask, const double* frames, uint8_t* rgb, float* depth, int32_t* seg, int env_begin, int env_count, cudaStream_t s);
int rrk_render_camera(rr_ctx* ctx, const RenderConst* cam, const uint8_t* mask, uint8_t* rgb, float* depth, int32_t* seg, cudaStream_t s);
void rr_set_camera_consts(RenderConst& C, const float* view16, const float* proj16);
int rrk_collide_upload_constants(rr_ctx* ctx);
int rrk_collide(rr_ctx* ctx, cudaStream_t s, bool flip = true, int mode = 0);   // mode 1 / 2: all but / only the heavy kernel's envs
int rrk_drop_collide_ahead(rr_ctx* ctx, cudaStream_t s);
int rrk_join_heavy(rr_ctx* ctx, cudaStream_t s);     // every reader / writer of the state but the eye-camera rasteriser calls this first
int rrk_clear_manifolds(rr_ctx* ctx, const uint8_t* mask, cudaStream_t s);
int rrk_ik(rr_ctx* ctx, const uint8_t* mask, const double* pos_quat, double* out, int max_iter, double residual, cudaStream_t s);
