// rr_render.cu — batched strip rasteriser with TinyRenderer's conventions, sm_100a.  Compile with -fmad=false.
//
// Replaces getCameraImage(320,240, TinyRenderer) as called by the reference at real_robots/envs/env.py:553-558
// (EyeCamera.renderTarget; view matrix env.py:543-546, projection env.py:548-551) and the output slicing at
// env.py:560-567: RGB u8[240,320,3], depth f32[240,320], segmentation mask i32[240,320].
//
// Pipeline per env (all fp32, same operation order as TinyRenderer's vec/mat templates, no FMA contraction):
//   k_render_setup : one thread per (env, shape): model / projection*view*model / inverse-transpose matrices and a
//                    conservative screen-row range of the shape.
//   k_raster       : persistent CTAs pull envs from a global counter.  Geometry runs once per env (conservative
//                    model-space back-face pre-test, then vertex shader / exact cull / near-plane clip / viewport /
//                    set-up for the survivors) and leaves 80-byte records plus per-strip work lists in an L2-resident
//                    scratch.  The image is then produced in strips of STRIP_ROWS raster rows: a 64-bit key per pixel
//                    in shared memory = (order-preserving depth bits << 32) | (triangle + 1) << 15 | record, and
//                    atomicMax of the key of every covered pixel — the z-test "later triangle wins ties" of the
//                    sequential renderer is exactly max over (depth, index).  After a barrier the winners are shaded
//                    (one fragment-shader evaluation per pixel) and RGB / depth / mask rows leave in packed
//                    12/16/16-byte streaming stores (vertical flip folded in).
// Images are written once; meshes (1.7 MB) and textures (10 MB) are shared by all envs and stay L2 resident.
#include "rr_ctx.h"
#include <cfloat>

__constant__ RenderConst c_render;
__constant__ float c_cull_hmax;      // largest 'smallest altitude' over all triangles (see build_planes)

// Tunables (all overridable with -D for scripts/build_variant.sh sweeps; the defaults are the measured optimum on B200)
#ifndef STRIP_ROWS
#define STRIP_ROWS 40              // raster rows per strip: 40 x 320 x 8 B of keys = 100 KB, two CTAs per SM
#endif
#define NSTRIPS (RR_H / STRIP_ROWS)
#ifndef RTHREADS
#define RTHREADS 512               // threads per CTA
#endif
#ifndef RBLOCKS
#define RBLOCKS 2                  // resident CTAs per SM (shared-memory and register bound: 2 x 512 x 64 registers)
#endif
#define MAT_STRIDE RR_MAT_STRIDE   // floats per (env, shape): pmv[16] model[12] invT[9] ylo yhi flag cam_model[3] cull_bound
#ifndef SMALL_AREA
#define SMALL_AREA 16              // in-strip bounding-box pixels up to which a triangle is rasterised by one lane
#endif
#ifndef PRECULL_AREA
#define PRECULL_AREA 4             // bounding boxes up to this many pixels are coverage-tested in the geometry phase
#endif
#ifndef MED_LANES
#define MED_LANES 8                // lanes that share one medium record (32 / MED_LANES records per warp at a time)
#endif
#ifndef HUGE_AREA
#define HUGE_AREA 256              // in-strip bounding-box pixels above which HUGE_GROUP threads share the triangle
#endif
#ifndef HUGE_GROUP
#define HUGE_GROUP 256
#endif
#ifndef REVERSE_ORDER
#define REVERSE_ORDER 1
#endif
#ifndef TMA_MATS
#define TMA_MATS 0                 // 1: per-env matrix tile staged by a TMA bulk copy (cp.async.bulk + mbarrier) one env ahead.
                                   // Bit-identical and tested, but measured 3 % slower on B200 (4.82 vs 4.68 ms per 4096-env launch):
                                   // the tile base becomes a run-time value in every matrix access and the load it hides is one
                                   // L2 round trip per env.  0: plain coalesced loads + a barrier.
#endif
#ifndef RASTER_LATE_SPARE
#define RASTER_LATE_SPARE 4        // CTAs k_raster leaves out of its grid when other kernels of the step still have to start (see render_range)
#endif
#ifndef GEO_PREFETCH
#define GEO_PREFETCH 1             // geometry passes load the next round's indices / planes one round ahead
#endif

#ifdef RR_STATS   // tuning build only (scripts/raster_stats.py): event counters, read back with rr_debug_raster_stats
__device__ unsigned long long g_stats[16];
#define STAT(i, n) atomicAdd(&g_stats[i], (unsigned long long)(n))
extern "C" int rr_debug_raster_stats(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(out, g_stats, sizeof(g_stats)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_stats, z, sizeof(z)); }
    return 0;
}
#else
#define STAT(i, n)
#endif

// --------------------------------------------------------------------------------------------- TMA bulk copy + mbarrier
// (cp.async.bulk: the 1-D form of the TMA engine; SASS UBLKCP + SYNCS)
#ifndef RR_SIMT_EMU_BUILD
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// one thread: arm the barrier with the byte count, then start the bulk copy global -> shared that completes it
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
#endif

// --------------------------------------------------------------------------------------------- float helpers
__device__ __forceinline__ float dot4r(const float* a, const float* b) { float r = 0.f; r += a[3] * b[3]; r += a[2] * b[2]; r += a[1] * b[1]; r += a[0] * b[0]; return r; }
__device__ __forceinline__ float dot4r(const float* a, float b0, float b1, float b2, float b3) { float r = 0.f; r += a[3] * b3; r += a[2] * b2; r += a[1] * b1; r += a[0] * b0; return r; }
__device__ __forceinline__ float dot3r(float a0, float a1, float a2, const float* b) { float r = 0.f; r += a2 * b[2]; r += a1 * b[1]; r += a0 * b[0]; return r; }

__device__ float det2r(const float* m) {   // 2x2 row-major, expansion along row 0 from the last column
    float ret = 0.f;
    ret += m[1] * (m[2] * -1.f);
    ret += m[0] * (m[3] * 1.f);
    return ret;
}
__device__ float det3r(const float* m) {   // 3x3 row-major
    float ret = 0.f;
#pragma unroll
    for (int i = 2; i >= 0; i--) {
        float mn[4]; int k = 0;
#pragma unroll
        for (int r = 1; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) if (c != i) mn[k++] = m[3 * r + c];
        float cof = det2r(mn) * ((i % 2) ? -1.f : 1.f);
        ret += m[i] * cof;
    }
    return ret;
}
__device__ float cofactor4(const float M[4][4], int row, int col) {
    float mn[9]; int k = 0;
    for (int r = 0; r < 4; r++) if (r != row) for (int c = 0; c < 4; c++) if (c != col) mn[k++] = M[r][c];
    return det3r(mn) * (((row + col) % 2) ? -1.f : 1.f);
}

// --------------------------------------------------------------------------------------------- setup kernel
// frames: f64 [N][22][12] or NULL (then read the simulated state)
// matrices of shape s of env e (see the file header); called by k_render_setup and, for an env whose state arrives late, by k_raster
__device__ void render_setup_shape(const DevState& st, const double* __restrict__ frames, float* __restrict__ mats, const int n_objects,
                                   const double* __restrict__ static_frames /* [22][12] for non-robot, non-object shapes */, const int e, const int s) {
    float* out = mats + ((size_t)e * RR_NSHAPES + s) * MAT_STRIDE;
    const ShapeRenderConst& sh = c_render.shape[s];
    int body = sh.objectAndLink & 0xffffff, link = (sh.objectAndLink >> 24) - 1;
    bool active = !(body >= 2 && body - 2 >= n_objects) && sh.rgba[3] != 0.f;
    double F[12];
    if (frames) {
        for (int k = 0; k < 12; k++) F[k] = frames[((size_t)e * RR_NSHAPES + s) * 12 + k];
    } else if (body == 0 && link >= 0) {
        for (int k = 0; k < 12; k++) F[k] = __ldcg(&st.link_frames[((size_t)link * 12 + k) * st.N + e]);   // L2 reads: k_raster calls this for envs whose
                                                                                                            // neighbours in these arrays may still be written
    } else if (body >= 2) {
        int o = body - 2;
        double q[4];
        for (int k = 0; k < 4; k++) q[k] = __ldcg(&st.obj[((size_t)o * 13 + 3 + k) * st.N + e]);
        double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
        double sc = 2.0 / d;
        double xs = q[0] * sc, ys = q[1] * sc, zs = q[2] * sc;
        double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
        double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
        double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
        F[0] = 1.0 - (yy + zz); F[1] = xy - wz; F[2] = xz + wy;
        F[3] = xy + wz; F[4] = 1.0 - (xx + zz); F[5] = yz - wx;
        F[6] = xz - wy; F[7] = yz + wx; F[8] = 1.0 - (xx + yy);
        for (int k = 0; k < 3; k++) F[9 + k] = __ldcg(&st.obj[((size_t)o * 13 + k) * st.N + e]);
    } else {
        for (int k = 0; k < 12; k++) F[k] = static_frames[s * 12 + k];
    }
    float model[4][4];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) model[i][j] = (float)F[3 * i + j]; model[i][3] = (float)F[9 + i]; }
    model[3][0] = model[3][1] = model[3][2] = 0.f; model[3][3] = 1.f;
    float mv[4][4], pmv[4][4];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++)
        mv[i][j] = dot4r(&c_render.view[4 * i], model[0][j], model[1][j], model[2][j], model[3][j]);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++)
        pmv[i][j] = dot4r(&c_render.proj[4 * i], mv[0][j], mv[1][j], mv[2][j], mv[3][j]);
    // invert_transpose: adjugate / (adj[0] . model[0])
    float adj0[4];
    for (int j = 0; j < 4; j++) adj0[j] = cofactor4(model, 0, j);
    float tmp = dot4r(adj0, model[0]);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[4 * i + j] = pmv[i][j];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 4; j++) out[16 + 4 * i + j] = model[i][j];
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out[28 + 3 * i + j] = ((i == 0) ? adj0[j] : cofactor4(model, i, j)) / tmp;
    // conservative raster-row range from the 8 corners of the local AABB
    float ylo = 0.f, yhi = (float)(RR_H - 1);
    bool all_front = true, all_behind = true; float mn = FLT_MAX, mx = -FLT_MAX;
    for (int c = 0; c < 8; c++) {
        {   // a shape whose 8 AABB corners all have w < -eps lies behind the camera: every triangle is discarded
            float q[4] = {(c & 1) ? sh.aabb_max[0] : sh.aabb_min[0], (c & 2) ? sh.aabb_max[1] : sh.aabb_min[1], (c & 4) ? sh.aabb_max[2] : sh.aabb_min[2], 1.f};
            if (!(dot4r(pmv[3], q) < -1e-3f)) all_behind = false;
        }
        float p[4] = {(c & 1) ? sh.aabb_max[0] : sh.aabb_min[0], (c & 2) ? sh.aabb_max[1] : sh.aabb_min[1], (c & 4) ? sh.aabb_max[2] : sh.aabb_min[2], 1.f};
        float cy = dot4r(pmv[1], p), cw = dot4r(pmv[3], p);
        if (!(cw > 1e-3f)) { all_front = false; continue; }
        float y = (RR_H / 2.f) * cw + (RR_H / 2.f) * cy;
        y = y / cw;
        mn = fminf(mn, y); mx = fmaxf(mx, y);
    }
    if (all_front) { ylo = mn - 2.f; yhi = mx + 2.f; }
    out[37] = ylo; out[38] = yhi; out[39] = (active && !all_behind) ? 1.f : 0.f;
    // Camera position in model coordinates and the distance bound of the conservative back-face pre-test (k_raster,
    // geometry pass 1).  The pre-test is only trusted for a rigid model matrix; otherwise the bound is infinite and every
    // triangle takes the exact test.
    double R[3][3], t[3], orth = 0.0;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) R[i][j] = (double)model[i][j]; t[i] = (double)model[i][3]; }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double d = R[0][i] * R[0][j] + R[1][i] * R[1][j] + R[2][i] * R[2][j] - (i == j ? 1.0 : 0.0);
        orth = fmax(orth, fabs(d));
    }
    double cm[3], cn = 0.0, rad = 0.0;
    for (int j = 0; j < 3; j++) {
        cm[j] = R[0][j] * (c_render.camP[0] - t[0]) + R[1][j] * (c_render.camP[1] - t[1]) + R[2][j] * (c_render.camP[2] - t[2]);
        cn += cm[j] * cm[j];
        const double a = fmax(fabs((double)sh.aabb_min[j]), fabs((double)sh.aabb_max[j]));
        rad += a * a;
    }
    out[40] = (float)cm[0]; out[41] = (float)cm[1]; out[42] = (float)cm[2];
    const double bound = (sqrt(cn) + sqrt(rad) + (double)c_cull_hmax) * 1.001;
    out[43] = (orth < 1e-5 && bound == bound) ? (float)bound : FLT_MAX;
}
__global__ void k_render_setup(DevState st, const double* __restrict__ frames, float* __restrict__ mats, int n_objects, const double* __restrict__ static_frames,
                               int env_begin, int env_count) {
    int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= env_count * RR_NSHAPES) return;
    render_setup_shape(st, frames, mats, n_objects, static_frames, env_begin + gid / RR_NSHAPES, gid % RR_NSHAPES);
}
// rrk_render right after rrk_step: the solver's heavy kernel (rr_physics.cu) may still be running.  The rasteriser then takes the
// envs in the solver's order, lightest first, and for the heavy ones at the end waits for the ticket that kernel leaves per env,
// then computes the env's matrices itself (k_render_setup saw a half-written state for them).  The heavy kernel was
// enqueued before the solver's main kernel, which this launch follows in stream order: its blocks have had idle SMs to start on
// long before a CTA of this kernel can spin on them.  The spin is bounded all the same: a time-out (~2 s) sets *timeout and renders
// what is there, and the host reports it at the next call.
struct LateEnvs {
    const int* order;          // DevState::solve_order, or nullptr: identity order, nobody is late
    const int* ready; int ticket, heavy_cap;
    DevState st; int n_objects; const double* static_frames; float* mats;
    int* timeout;
};

// --------------------------------------------------------------------------------------------- triangle setup
struct Tri {
    float c[4][3];         // clip coordinates: row (x,y,z,w) x vertex
};
struct Screen {
    float p2[3][2];        // screen x/w, y/w
    float w[3];
    int x0, x1, y0, y1;
    bool any;
};

__device__ __forceinline__ void load_corner(const float* __restrict__ tv, int j, float* v) {
    const float4* p = reinterpret_cast<const float4*>(tv + 8 * j);
    float4 a = __ldg(p), b = __ldg(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// vertex shader for the three corners + back-face test.  returns false if culled.
__device__ __forceinline__ bool tri_vertex(const float* __restrict__ M, const float* __restrict__ tv, Tri& T) {
    float wv[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(tv + 8 * j));
        float px = a.x * 1.f, py = a.y * 1.f, pz = a.z * 1.f;
#pragma unroll
        for (int k = 0; k < 4; k++) T.c[k][j] = dot4r(M + 4 * k, px, py, pz, 1.f);
#pragma unroll
        for (int k = 0; k < 3; k++) wv[k][j] = dot4r(M + 16 + 4 * k, px, py, pz, 1.f);
    }
    double v0[3], e1[3], e2[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { v0[k] = (double)wv[k][0]; e1[k] = (double)wv[k][1] - v0[k]; e2[k] = (double)wv[k][2] - v0[k]; }
    double Nx = e1[1] * e2[2] - e1[2] * e2[1], Ny = e1[2] * e2[0] - e1[0] * e2[2], Nz = e1[0] * e2[1] - e1[1] * e2[0];
    double dd = (v0[0] - c_render.camP[0]) * Nx + (v0[1] - c_render.camP[1]) * Ny + (v0[2] - c_render.camP[2]) * Nz;
    return !(dd >= 0);
}

// clip coordinates only (triangles the pre-test has already classified as front-facing)
__device__ __forceinline__ void tri_clip(const float* __restrict__ M, const float* __restrict__ tv, Tri& T) {
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(tv + 8 * j));
        float px = a.x * 1.f, py = a.y * 1.f, pz = a.z * 1.f;
#pragma unroll
        for (int k = 0; k < 4; k++) T.c[k][j] = dot4r(M + 4 * k, px, py, pz, 1.f);
    }
}

__device__ __forceinline__ void screen_setup(const float c[4][3], Screen& S) {
    const float W = RR_W, H = RR_H;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        float cx = c[0][i], cy = c[1][i], cz = c[2][i], cw = c[3][i];
        float x = 0.f; x += (W / 2.f) * cw; x += 0.f * cz; x += 0.f * cy; x += (W / 2.f) * cx;
        float y = 0.f; y += (H / 2.f) * cw; y += 0.f * cz; y += (H / 2.f) * cy; y += 0.f * cx;
        float w = 0.f; w += 1.f * cw; w += 0.f * cz; w += 0.f * cy; w += 0.f * cx;
        S.w[i] = w;
        S.p2[i][0] = x / w; S.p2[i][1] = y / w;
    }
    float mn[2] = {FLT_MAX, FLT_MAX}, mx[2] = {-FLT_MAX, -FLT_MAX};
    const float cl[2] = {W - 1, H - 1};
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            float a = mn[j] < S.p2[i][j] ? mn[j] : S.p2[i][j];
            mn[j] = 0.f > a ? 0.f : a;
            float b = mx[j] > S.p2[i][j] ? mx[j] : S.p2[i][j];
            mx[j] = cl[j] < b ? cl[j] : b;
        }
    S.x0 = (int)mn[0]; S.y0 = (int)mn[1];
    S.any = ((float)S.x0 <= mx[0]) && ((float)S.y0 <= mx[1]);
    S.x1 = S.any ? (int)floorf(mx[0]) : -1;
    S.y1 = S.any ? (int)floorf(mx[1]) : -1;
}

// per-triangle part of TinyRenderer's barycentric(): s[i][0] = C-A, s[i][1] = B-A, and u.z (pixel independent)
struct Edge { float s00, s01, s10, s11, uz; bool degenerate; };
__device__ __forceinline__ void edge_setup(const Screen& S, Edge& E) {
    const float* A = S.p2[0]; const float* B = S.p2[1]; const float* C = S.p2[2];
    E.s00 = C[0] - A[0]; E.s01 = B[0] - A[0];
    E.s10 = C[1] - A[1]; E.s11 = B[1] - A[1];
    E.uz = E.s00 * E.s11 - E.s01 * E.s10;
    E.degenerate = !((double)fabsf(E.uz) > 1e-2);
}
__device__ __forceinline__ void bary_u(const Screen& S, const Edge& E, float px, float py, float& ux, float& uy) {
    float s02 = S.p2[0][0] - px, s12 = S.p2[0][1] - py;
    ux = E.s01 * s12 - s02 * E.s11;
    uy = s02 * E.s10 - E.s00 * s12;
}
// exact sign classification of bc_screen = (1-(ux+uy)/uz, uy/uz, ux/uz) without dividing.
// returns 0 = some component is negative (pixel rejected), 1 = all non-negative, 2 = too close to call (do the divisions)
__device__ __forceinline__ int coverage_class(float ux, float uy, float uz) {
    const bool zn = uz < 0.f;
    // uy/uz < 0  <=>  signs differ and the rounded quotient is not a zero (no underflow when |uy| >= 1e-30, |uz| < 1e7)
    float auy = fabsf(uy), aux = fabsf(ux);
    if (!(fabsf(uz) < 1e7f)) return 2;
    if (auy < 1e-10f || auy > 1e13f) { if (uy != 0.f) return 2; }
    else if ((uy < 0.f) != zn) return 0;
    if (aux < 1e-10f || aux > 1e13f) { if (ux != 0.f) return 2; }
    else if ((ux < 0.f) != zn) return 0;
    // 1 - RN(t/uz) < 0  <=>  RN(t/uz) > 1
    float t = ux + uy;
    float at = zn ? -t : t, az = zn ? -uz : uz;          // az > 0
    if (at <= az) return 1;                               // exact quotient <= 1, rounding is monotonic
    if (at > az * 1.000001f) return 0;                    // exact quotient > 1 + 2^-23: rounds above 1
    return 2;
}
// Correctly rounded a / b from a correctly rounded reciprocal y = RN(1/b) and two FMA residual corrections
// (Markstein): bit-identical to IEEE division inside the guarded range, several times cheaper when y is shared.
// Checked against true division on 2e9 random operand pairs (incl. significands near 1.0 and near all-ones).
__device__ __forceinline__ float div_fast(float a, float b, float y, bool ok) {
    const float aa = fabsf(a);
    if (ok && (aa == 0.f || (aa > 1e-25f && aa < 1e25f))) {
        float q = a * y;
        if (aa == 0.f) return q;
        float r = __fmaf_rn(-b, q, a); q = __fmaf_rn(r, y, q);
        r = __fmaf_rn(-b, q, a); q = __fmaf_rn(r, y, q);
        return q;
    }
    return a / b;
}
// same, without the range guards: only valid when the caller has established the operand ranges (see raster_one)
__device__ __forceinline__ float div_fast_ng(float a, float b, float y) {
    const float q0 = a * y;                 // a == +-0: the quotient is q0 (the corrections below would turn -0 into +0); a select, not a branch
    float r = __fmaf_rn(-b, q0, a), q = __fmaf_rn(r, y, q0);
    r = __fmaf_rn(-b, q, a); q = __fmaf_rn(r, y, q);
    return a == 0.f ? q0 : q;
}
__device__ __forceinline__ bool recip_ok(float b) { const float ab = fabsf(b); return ab > 1e-12f && ab < 1e12f; }
struct Recip { float yuz, yw[3]; bool ok, tight; };
__device__ __forceinline__ bool w_tight(float w) { return w > 1e-6f && w < 1e6f; }
__device__ __forceinline__ void recip_setup(const Screen& S, const Edge& E, Recip& R) {
    R.ok = recip_ok(E.uz) && recip_ok(S.w[0]) && recip_ok(S.w[1]) && recip_ok(S.w[2]);
    R.tight = R.ok && fabsf(E.uz) < 1e7f && w_tight(S.w[0]) && w_tight(S.w[1]) && w_tight(S.w[2]);
    R.yuz = __frcp_rn(E.uz); R.yw[0] = __frcp_rn(S.w[0]); R.yw[1] = __frcp_rn(S.w[1]); R.yw[2] = __frcp_rn(S.w[2]);
}
// bc_screen (bs) and the perspective-correct bc_clip (bc) of pixel (px, py): same quotients as TinyRenderer's triangle()
__device__ __forceinline__ void frag_bar_r(const Screen& S, const Edge& E, const Recip& R, int px, int py, float* bs, float* bc) {
    float ux, uy; bary_u(S, E, (float)px, (float)py, ux, uy);
    if (!E.degenerate) {
        bs[0] = 1.f - div_fast(ux + uy, E.uz, R.yuz, R.ok); bs[1] = div_fast(uy, E.uz, R.yuz, R.ok); bs[2] = div_fast(ux, E.uz, R.yuz, R.ok);
    } else { bs[0] = -1.f; bs[1] = 1.f; bs[2] = 1.f; }
    const float b0 = div_fast(bs[0], S.w[0], R.yw[0], R.ok), b1 = div_fast(bs[1], S.w[1], R.yw[1], R.ok), b2 = div_fast(bs[2], S.w[2], R.yw[2], R.ok);
    const float s = b0 + b1 + b2;
    const bool oks = recip_ok(s);
    const float ys = __frcp_rn(s);
    bc[0] = div_fast(b0, s, ys, oks); bc[1] = div_fast(b1, s, ys, oks); bc[2] = div_fast(b2, s, ys, oks);
}
// Fast variant for pixels whose coverage class is 1 (all three bc_screen components clearly >= 0) on a triangle with
// R.ok: then every numerator is either exactly 0 or lies in [1e-23, 1e20], every quotient and FMA residual is a
// normal float, and the guarded and unguarded divisions coincide.  (ux, uy in {0} U [1e-10, 1e13] by coverage_class,
// uz in [1e-2, 1e7], w in [1e-6, 1e6], bs0 = 1 - q is 0 or >= 2^-24, s >= the largest b_i > 0.)
__device__ __forceinline__ void frag_bar_in(const Screen& S, const Edge& E, const Recip& R, float ux, float uy, float* bc) {
    const float bs0 = 1.f - div_fast_ng(ux + uy, E.uz, R.yuz), bs1 = div_fast_ng(uy, E.uz, R.yuz), bs2 = div_fast_ng(ux, E.uz, R.yuz);
    const float b0 = div_fast_ng(bs0, S.w[0], R.yw[0]), b1 = div_fast_ng(bs1, S.w[1], R.yw[1]), b2 = div_fast_ng(bs2, S.w[2], R.yw[2]);
    const float s = b0 + b1 + b2;
    const float ys = __frcp_rn(s);
    bc[0] = div_fast_ng(b0, s, ys); bc[1] = div_fast_ng(b1, s, ys); bc[2] = div_fast_ng(b2, s, ys);
}
__device__ __forceinline__ void frag_bar(const Screen& S, const Edge& E, int px, int py, float* bs, float* bc) {
    Recip R; recip_setup(S, E, R);
    frag_bar_r(S, E, R, px, py, bs, bc);
}
__device__ __forceinline__ float frag_depth(const Screen& S, const Edge& E, const float* clipz, int px, int py, float* bs, float* bc) {
    frag_bar(S, E, px, py, bs, bc);
    return -1.f * dot3r(clipz[0], clipz[1], clipz[2], bc);
}

__device__ __forceinline__ unsigned int order_bits(float f) {
    unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unorder_bits(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// near-plane clipping (TinyRenderer.cpp clipEdge / clipTriangleAgainstNearplane). returns #triangles, -1 = unclipped
__device__ __noinline__ int clip_near(const float tri[4][3], float out[3][4][3]) {
    if (tri[3][0] < 0 && tri[3][1] < 0 && tri[3][2] < 0) return 0;
    if (tri[3][0] >= 0 && tri[3][1] >= 0 && tri[3][2] >= 0) return -1;
    float verts[8][4]; int nv = 0;
    for (int e = 0; e < 3; e++) {
        int ia = e, ib = (e + 1) % 3;
        float v0[4], v1[4];
        for (int k = 0; k < 4; k++) { v0[k] = tri[k][ia]; v1[k] = tri[k][ib]; }
        bool in0 = v0[3] > 0.f && v0[2] > -v0[3];
        bool in1 = v1[3] > 0.f && v1[2] > -v1[3];
        if (in0 && in1) {
        } else if (in0 || in1) {
            float d0 = v0[2] + v0[3], d1 = v1[2] + v1[3];
            float factor = (float)(1.0 / (double)(d1 - d0));
            float nv4[4];
            for (int k = 0; k < 4; k++) nv4[k] = (v0[k] * d1 - v1[k] * d0) * factor;
            if (in0) { for (int k = 0; k < 4; k++) v1[k] = nv4[k]; } else { for (int k = 0; k < 4; k++) v0[k] = nv4[k]; }
        } else continue;
        if (nv == 0 || !(verts[nv - 1][0] == v0[0] && verts[nv - 1][1] == v0[1] && verts[nv - 1][2] == v0[2] && verts[nv - 1][3] == v0[3])) {
            for (int k = 0; k < 4; k++) verts[nv][k] = v0[k];
            nv++;
        }
        for (int k = 0; k < 4; k++) verts[nv][k] = v1[k];
        nv++;
    }
    if (nv < 3) return 0;
    if (verts[0][0] == verts[nv - 1][0] && verts[0][1] == verts[nv - 1][1] && verts[0][2] == verts[nv - 1][2] && verts[0][3] == verts[nv - 1][3]) nv--;
    int nt = 0;
    for (int i = 1; i < nv - 1 && nt < 3; i++) {
        for (int k = 0; k < 4; k++) { out[nt][k][0] = verts[0][k]; out[nt][k][1] = verts[i][k]; out[nt][k][2] = verts[i + 1][k]; }
        nt++;
    }
    return nt;
}

// Upper bound (as order-preserving bits) of every fragment depth the triangle can produce: fd = -(clipz . bc) with
// bc >= 0 and sum(bc) = 1 up to a few ulp, so fd <= max(-clipz) + 1e-5 max|clipz|.  0xffffffff = no bound.
__device__ __forceinline__ unsigned int depth_bound_bits(const Recip& R, const float* clipz) {
    if (!R.tight) return 0xffffffffu;
    const float zm = fmaxf(fmaxf(-clipz[0], -clipz[1]), -clipz[2]);
    const float am = fmaxf(fmaxf(fabsf(clipz[0]), fabsf(clipz[1])), fabsf(clipz[2]));
    const float b = zm + (am * 1e-5f + 1e-30f);
    return (b == b) ? order_bits(b) : 0xffffffffu;
}

// Rasterise the part of the bounding box inside the strip; this thread takes pixels start, start + stride, ...
// Two nested loops so that the lanes of a warp re-converge before the expensive part: the inner loop advances to the
// thread's next pixel that is (a) covered and (b) not already behind the z-buffer by the triangle's depth bound; the
// outer body then computes the exact barycentrics / depth with as many lanes active as have found such a pixel.
__device__ __forceinline__ void raster_pixels(unsigned long long* keys, const Screen& S, const Edge& E, const Recip& R, const float* clipz,
                                              unsigned int keylo, bool planes, int row0, int start, int stride) {
    const int ya = max(S.y0, row0), yb = min(S.y1, row0 + STRIP_ROWS - 1);
    if (ya > yb) return;
    const int wdt = S.x1 - S.x0 + 1, n = wdt * (yb - ya + 1);
    int sy = 0, dy = 0;
    if (stride != 1) {
        // (k + 0.5) / wdt is at least 0.5 / 320 away from an integer, so the float quotient truncates to the exact k / wdt
        const float rw = __frcp_rn((float)wdt);
        sy = (int)(((float)start + 0.5f) * rw); dy = (int)(((float)stride + 0.5f) * rw);
    } else dy = (wdt == 1) ? 1 : 0;
    int x = S.x0 + (start - sy * wdt), y = ya + sy;
    const int dx = stride - dy * wdt;
    const unsigned int zb = depth_bound_bits(R, clipz);
    int i = start;
    for (;;) {
        float ux = 0.f, uy = 0.f;
        int cls = 0, px = 0, py = 0;
        while (i < n) {
            px = x; py = y;
            i += stride; x += dx; y += dy; if (x > S.x1) { x -= wdt; y++; }
            bary_u(S, E, (float)px, (float)py, ux, uy);
            cls = coverage_class(ux, uy, E.uz);
            STAT(6, 1);
            if (cls) {
                STAT(7, 1);
                const unsigned int cur = reinterpret_cast<volatile unsigned int*>(&keys[(py - row0) * RR_W + px])[1];
                if (cur <= zb) break;
                STAT(8, 1);
                cls = 0;
            }
        }
        if (cls == 0) return;
        float bs[3], bc[3];
        if (cls == 1 && R.tight) frag_bar_in(S, E, R, ux, uy, bc);
        else {
            frag_bar_r(S, E, R, px, py, bs, bc);
            if (cls == 2 && (bs[0] < 0 || bs[1] < 0 || bs[2] < 0)) continue;
        }
        const float fd = -1.f * dot3r(clipz[0], clipz[1], clipz[2], bc);
        if (!(fd >= -1e30f)) continue;                                // z-buffer cleared to -1e30 (also drops NaN)
        if (planes && (fd < -c_render.farPlane || fd > -c_render.nearPlane)) continue;
        const unsigned long long key = ((unsigned long long)order_bits(fd) << 32) | keylo;
        unsigned long long* slot = &keys[(py - row0) * RR_W + px];
        STAT(9, 1);
        if (*reinterpret_cast<volatile unsigned long long*>(slot) < key) { STAT(10, 1); atomicMax(slot, key); }
    }
}


// --------------------------------------------------------------------------------------------- fragment shader
__device__ __forceinline__ void texel_fetch(int tex, const uint8_t* __restrict__ texdata, float u, float v, int* c) {
    if (tex < 0) { c[0] = c[1] = c[2] = 255; return; }
    const TexConst& t = c_render.tex[tex];
    float uf = u, vf = v;
    if (!(u >= 0.f && u < 1.f)) { double ip; uf = (float)modf((double)u, &ip); }
    if (!(v >= 0.f && v < 1.f)) { double ip; vf = (float)modf((double)v, &ip); }
    int x = (int)(uf * (float)t.w), y = (int)(vf * (float)t.h);
    if (x < 0 || y < 0 || x >= t.w || y >= t.h) { c[0] = c[1] = c[2] = 0; return; }
    const uint8_t* p = texdata + t.offset + (t.uniform ? 0 : ((size_t)x + (size_t)y * t.w) * 3);
    c[0] = __ldg(p); c[1] = __ldg(p + 1); c[2] = __ldg(p + 2);
}

// --------------------------------------------------------------------------------------------- raster kernel
// Visible-triangle record written by the geometry phase (20 words = 80 B):
//   bx = x0 | x1 << 16, by = y0 | y1 << 16 (clamped bounding box), keylo = (triangle + 1) << 15 | record index,
//   flags (REC_*), screen coordinates p2[3][2], w[3], clip z[3], and the correctly rounded reciprocals of u.z and w[3].
struct __align__(16) Rec { unsigned int bx, by, keylo, flags; float p2[3][2]; float w[3]; float z[3]; float yuz; float yw[3]; };
static_assert(sizeof(Rec) == 80, "record must be 80 bytes");
#define REC_MAX 24576          // records per env (17 578 triangles + clipped pieces), < 2^15
#define CULL_MAX 32768          // survivor list entries per CTA (triangle index < 2^15, bit 15 = uncertain)
#define REC_CLIPPED 1u         // piece of a near-plane-clipped triangle: no near/far discard, shading from the original triangle
#define REC_OK 2u              // reciprocal operands inside the guarded range
#define REC_TIGHT 4u           // ... and inside the tighter range of the unguarded fast path
#define REC_DEGENERATE 8u      // |u.z| <= 1e-2: TinyRenderer rejects every pixel

__device__ __forceinline__ int agg_inc(int* ctr) {      // warp-aggregated atomic increment of a shared counter
#ifdef RR_SIMT_EMU_BUILD
    return atomicAdd(ctr, 1);           // (the emulator runs one thread at a time: no aggregation to do)
#endif
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(ctr, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane) - 1u));
}

__device__ __forceinline__ void emit_record(Rec* __restrict__ recs, unsigned short* __restrict__ lists, int* nrec, int* nlist,
                                            const Screen& S, const float* z, unsigned int gt, unsigned int flags) {
    if (!S.any) return;
    Edge E; edge_setup(S, E);
    if (E.degenerate) return;                               // bc_screen = (-1,1,1) everywhere: nothing is ever drawn
#if PRECULL_AREA > 0
    {   // most robot triangles are sub-pixel: classify their few candidate pixels now and drop the triangle when none
        // can be covered — fewer records keeps the scratch inside L2
        const int bw = S.x1 - S.x0 + 1, bh = S.y1 - S.y0 + 1;
        if (bw * bh <= PRECULL_AREA) {
            bool any = false;
            for (int py = S.y0; py <= S.y1; py++)
                for (int px = S.x0; px <= S.x1; px++) {
                    float ux, uy; bary_u(S, E, (float)px, (float)py, ux, uy);
                    any |= coverage_class(ux, uy, E.uz) != 0;
                }
            if (!any) return;
        }
    }
#endif
    const int slot = agg_inc(nrec);
    if (slot >= REC_MAX) return;
    STAT(2, 1);
    Recip R; recip_setup(S, E, R);
    Rec r;
    r.bx = (unsigned)S.x0 | ((unsigned)S.x1 << 16); r.by = (unsigned)S.y0 | ((unsigned)S.y1 << 16);
    r.keylo = ((gt + 1u) << 15) | (unsigned)slot;
    r.flags = flags | (R.ok ? REC_OK : 0u) | (R.tight ? REC_TIGHT : 0u);
#pragma unroll
    for (int i = 0; i < 3; i++) { r.p2[i][0] = S.p2[i][0]; r.p2[i][1] = S.p2[i][1]; r.w[i] = S.w[i]; r.z[i] = z[i]; r.yw[i] = R.yw[i]; }
    r.yuz = R.yuz;
    const uint4* src = reinterpret_cast<const uint4*>(&r);
    uint4* dst = reinterpret_cast<uint4*>(recs + slot);
#pragma unroll
    for (int k = 0; k < 5; k++) dst[k] = src[k];
    const int s0 = S.y0 / STRIP_ROWS, s1 = S.y1 / STRIP_ROWS;
    const int bw = S.x1 - S.x0 + 1;
    for (int st = s0; st <= s1; st++) {
        // per strip three work lists by in-strip bounding-box size: small (own lane) grows upwards from entry 0 of the
        // strip's first array, huge (shared by HUGE_GROUP threads) downwards from its end; medium (one warp) fills the second
        const int ya = max(S.y0, st * STRIP_ROWS), yb = min(S.y1, st * STRIP_ROWS + STRIP_ROWS - 1);
        const int area = bw * (yb - ya + 1);
        const int cls = area > HUGE_AREA ? 2 : (area > SMALL_AREA ? 1 : 0);
        STAT(3 + cls, 1);
        // lanes of the warp may be appending to different counters: aggregate per counter
#ifdef RR_SIMT_EMU_BUILD
        const int pos = atomicAdd(&nlist[cls * NSTRIPS + st], 1);
#else
        const unsigned grp = __match_any_sync(__activemask(), st * 4 + cls);
        const int lane = threadIdx.x & 31, leader = __ffs(grp) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&nlist[cls * NSTRIPS + st], __popc(grp));
        base = __shfl_sync(grp, base, leader);
        const int pos = base + __popc(grp & ((1u << lane) - 1u));
#endif
        unsigned short* l = lists + (size_t)st * 2 * REC_MAX;
        l[cls == 0 ? pos : (cls == 2 ? REC_MAX - 1 - pos : REC_MAX + pos)] = (unsigned short)slot;
    }
}

__device__ __noinline__ void emit_clipped(Rec* __restrict__ recs, unsigned short* __restrict__ lists, int* nrec, int* nlist, const Tri& T, unsigned int gt) {
    float out[3][4][3];
    const int nc = clip_near(T.c, out);
    for (int c = 0; c < nc; c++) {
        Screen S; screen_setup(out[c], S);
        emit_record(recs, lists, nrec, nlist, S, out[c][2], gt, REC_CLIPPED);
    }
}

__device__ __forceinline__ void rec_unpack(const Rec& r, Screen& S, Edge& E, Recip& R, float* z) {
#pragma unroll
    for (int i = 0; i < 3; i++) { S.p2[i][0] = r.p2[i][0]; S.p2[i][1] = r.p2[i][1]; S.w[i] = r.w[i]; z[i] = r.z[i]; R.yw[i] = r.yw[i]; }
    S.x0 = r.bx & 0xffff; S.x1 = r.bx >> 16; S.y0 = r.by & 0xffff; S.y1 = r.by >> 16; S.any = true;
    R.yuz = r.yuz; R.ok = (r.flags & REC_OK) != 0; R.tight = (r.flags & REC_TIGHT) != 0;
    E.s00 = S.p2[2][0] - S.p2[0][0]; E.s01 = S.p2[1][0] - S.p2[0][0];
    E.s10 = S.p2[2][1] - S.p2[0][1]; E.s11 = S.p2[1][1] - S.p2[0][1];
    E.uz = E.s00 * E.s11 - E.s01 * E.s10;
    E.degenerate = false;
}
__device__ __forceinline__ Rec load_rec(const Rec* __restrict__ p) {
    Rec r; uint4* d = reinterpret_cast<uint4*>(&r); const uint4* s = reinterpret_cast<const uint4*>(p);
#pragma unroll
    for (int k = 0; k < 5; k++) d[k] = s[k];
    return r;
}

// --------------------------------------------------------------------------------------------- flattened raster (RASTER_FLAT)
// The per-pixel work of a record has two halves of very different weight: the coverage scan of its bounding box (cheap, ~85 %
// of the candidates fail) and the exact barycentrics / depth / z-test of a covered pixel (nine Markstein divisions).  With one
// lane walking one box both halves run at 10-13 active lanes of 32.  Here a warp keeps the records of its current work item
// in registers (one per lane, or one per group of lanes) and
//   scan  : small boxes are FLATTENED — the (record, pixel) pairs of the 32 records of a chunk are numbered consecutively
//           (prefix sum of the box areas) and dealt to the lanes 32 at a time, each lane fetching the few set-up values of its
//           pair's record from the owner lane with shuffles; medium / huge boxes are walked in lockstep by the lanes that
//           share them;
//   queue : candidates that are covered and not already behind the z-buffer are appended to a per-warp queue in shared
//           memory (pixel, coverage class, owner lane);
//   drain : whenever 32 are queued (and at the end of the work item) every lane takes one, fetches the record's depth set-up
//           from the owner lane with shuffles and does the exact evaluation + atomicMax — all 32 lanes busy.
// The arithmetic per fragment is unchanged (same functions, same operands), so the images stay bit-identical.
#ifndef RASTER_FLAT
#define RASTER_FLAT 0
#endif
#define FQ_SIZE 64
struct LaneRec {           // a record unpacked in registers, plus what the scan needs
    float A0x, A0y, s00, s01, s10, s11, uz, w0, w1, w2, yw0, yw1, yw2, yuz, cz0, cz1, cz2;
    unsigned int keylo, flags, zb;
    int x0, x1, ya, yb, wdt, n;           // in-strip part of the bounding box: rows ya..yb, wdt columns from x0, n candidates
};
struct FragQueue { unsigned int* q; int head, count; };

__device__ __forceinline__ void lane_rec_load(const Rec* __restrict__ recs, unsigned int ridx, bool valid, int row0, LaneRec& L) {
    L.n = 0; L.wdt = 1; L.x0 = 0; L.x1 = 0; L.ya = row0; L.yb = row0 - 1; L.zb = 0u; L.keylo = 0u; L.flags = 0u;
    L.A0x = L.A0y = L.s00 = L.s01 = L.s10 = L.s11 = 0.f; L.uz = 1.f; L.w0 = L.w1 = L.w2 = 1.f; L.yw0 = L.yw1 = L.yw2 = 1.f; L.yuz = 1.f; L.cz0 = L.cz1 = L.cz2 = 0.f;
    if (!valid) return;
    const Rec rec = load_rec(recs + ridx);
    Screen S; Edge E; Recip R; float cz[3]; rec_unpack(rec, S, E, R, cz);
    L.A0x = S.p2[0][0]; L.A0y = S.p2[0][1]; L.s00 = E.s00; L.s01 = E.s01; L.s10 = E.s10; L.s11 = E.s11; L.uz = E.uz;
    L.w0 = S.w[0]; L.w1 = S.w[1]; L.w2 = S.w[2]; L.yw0 = R.yw[0]; L.yw1 = R.yw[1]; L.yw2 = R.yw[2]; L.yuz = R.yuz;
    L.cz0 = cz[0]; L.cz1 = cz[1]; L.cz2 = cz[2]; L.keylo = rec.keylo; L.flags = rec.flags;
    L.zb = depth_bound_bits(R, cz);
    L.x0 = S.x0; L.x1 = S.x1; L.ya = max(S.y0, row0); L.yb = min(S.y1, row0 + STRIP_ROWS - 1);
    L.wdt = S.x1 - S.x0 + 1;
    L.n = L.ya > L.yb ? 0 : L.wdt * (L.yb - L.ya + 1);
}
// coverage class of pixel (px, py) for the triangle whose set-up is given; ux, uy out
__device__ __forceinline__ int cover_px(float A0x, float A0y, float s00, float s01, float s10, float s11, float uz, int px, int py) {
    const float s02 = A0x - (float)px, s12 = A0y - (float)py;
    const float ux = s01 * s12 - s02 * s11, uy = s02 * s10 - s00 * s12;
    return coverage_class(ux, uy, uz);
}
// queue entry: px [0,9) | row in strip [9,15) | coverage class [15,17) | owner lane [17,22)
__device__ __forceinline__ void fq_push(FragQueue& Q, bool have, unsigned int ent) {
    const unsigned int m = __ballot_sync(0xffffffffu, have);
    if (have) Q.q[(Q.head + Q.count + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))) & (FQ_SIZE - 1)] = ent;
    Q.count += __popc(m);
}
// every lane takes one queued fragment (the first n_take of the queue), evaluates it exactly and updates the z-buffer
__device__ __forceinline__ void fq_drain(unsigned long long* keys, int row0, FragQueue& Q, int n_take, const LaneRec& L) {
    const int lane = threadIdx.x & 31;
    __syncwarp();
    const bool mine = lane < n_take;
    const unsigned int ent = mine ? Q.q[(Q.head + lane) & (FQ_SIZE - 1)] : 0u;
    const int owner = (int)(ent >> 17) & 31;
    const unsigned int F = 0xffffffffu;
    Screen S; Edge E; Recip R; float cz[3];
    S.p2[0][0] = __shfl_sync(F, L.A0x, owner); S.p2[0][1] = __shfl_sync(F, L.A0y, owner);
    E.s00 = __shfl_sync(F, L.s00, owner); E.s01 = __shfl_sync(F, L.s01, owner); E.s10 = __shfl_sync(F, L.s10, owner); E.s11 = __shfl_sync(F, L.s11, owner);
    E.uz = __shfl_sync(F, L.uz, owner); E.degenerate = false;
    S.w[0] = __shfl_sync(F, L.w0, owner); S.w[1] = __shfl_sync(F, L.w1, owner); S.w[2] = __shfl_sync(F, L.w2, owner);
    R.yw[0] = __shfl_sync(F, L.yw0, owner); R.yw[1] = __shfl_sync(F, L.yw1, owner); R.yw[2] = __shfl_sync(F, L.yw2, owner);
    R.yuz = __shfl_sync(F, L.yuz, owner);
    cz[0] = __shfl_sync(F, L.cz0, owner); cz[1] = __shfl_sync(F, L.cz1, owner); cz[2] = __shfl_sync(F, L.cz2, owner);
    const unsigned int keylo = __shfl_sync(F, L.keylo, owner), flags = __shfl_sync(F, L.flags, owner);
    R.ok = (flags & REC_OK) != 0; R.tight = (flags & REC_TIGHT) != 0;
    if (mine) {
        const int px = (int)(ent & 511u), pr = (int)(ent >> 9) & 63, cls = (int)(ent >> 15) & 3, py = row0 + pr;
        float ux, uy; bary_u(S, E, (float)px, (float)py, ux, uy);
        float bs[3], bc[3];
        bool keep = true;
        if (cls == 1 && R.tight) frag_bar_in(S, E, R, ux, uy, bc);
        else {
            frag_bar_r(S, E, R, px, py, bs, bc);
            if (cls == 2 && (bs[0] < 0 || bs[1] < 0 || bs[2] < 0)) keep = false;
        }
        const float fd = -1.f * dot3r(cz[0], cz[1], cz[2], bc);
        if (!(fd >= -1e30f)) keep = false;                                // z-buffer cleared to -1e30 (also drops NaN)
        if (!(flags & REC_CLIPPED) && (fd < -c_render.farPlane || fd > -c_render.nearPlane)) keep = false;
        if (keep) {
            const unsigned long long key = ((unsigned long long)order_bits(fd) << 32) | keylo;
            unsigned long long* slot = &keys[pr * RR_W + px];
            STAT(9, 1);
            if (*reinterpret_cast<volatile unsigned long long*>(slot) < key) { STAT(10, 1); atomicMax(slot, key); }
        }
    }
    Q.head = (Q.head + n_take) & (FQ_SIZE - 1); Q.count -= n_take;
    __syncwarp();
}
// medium / huge boxes: lane takes candidates start, start + stride, ... of ITS record (L), all lanes in lockstep
__device__ __forceinline__ void scan_lockstep(unsigned long long* keys, int row0, const LaneRec& L, int start, int stride, FragQueue& Q) {
    const int lane = threadIdx.x & 31;
    const int n = L.n, wdt = L.wdt;
    int sy = 0, dy = 0;
    if (stride != 1) {
        // (k + 0.5) / wdt is at least 0.5 / 320 away from an integer, so the float quotient truncates to the exact k / wdt
        const float rw = __frcp_rn((float)wdt);
        sy = (int)(((float)start + 0.5f) * rw); dy = (int)(((float)stride + 0.5f) * rw);
    } else dy = (wdt == 1) ? 1 : 0;
    int x = L.x0 + (start - sy * wdt), y = L.ya + sy;
    const int dx = stride - dy * wdt;
    int i = start;
    while (__any_sync(0xffffffffu, i < n)) {
        bool have = false; unsigned int ent = 0u;
        if (i < n) {
            const int px = x, py = y;
            i += stride; x += dx; y += dy; if (x > L.x1) { x -= wdt; y++; }
            const int cls = cover_px(L.A0x, L.A0y, L.s00, L.s01, L.s10, L.s11, L.uz, px, py);
            STAT(6, 1);
            if (cls) {
                STAT(7, 1);
                const unsigned int cur = reinterpret_cast<volatile unsigned int*>(&keys[(py - row0) * RR_W + px])[1];
                if (cur <= L.zb) { have = true; ent = (unsigned int)px | ((unsigned int)(py - row0) << 9) | ((unsigned int)cls << 15) | ((unsigned int)lane << 17); }
                else STAT(8, 1);
            }
        }
        fq_push(Q, have, ent);
        if (Q.count >= 32) fq_drain(keys, row0, Q, 32, L);
    }
}
// small boxes, one record per lane (n <= SMALL_AREA each): the (record, pixel) pairs of the chunk, 32 at a time
__device__ __forceinline__ void scan_flat(unsigned long long* keys, int row0, const LaneRec& L, FragQueue& Q) {
    const int lane = threadIdx.x & 31;
    const unsigned int F = 0xffffffffu;
    int off = L.n;                                    // inclusive prefix sum of the candidate counts
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(F, off, d); if (lane >= d) off += t; }
    const int total = __shfl_sync(F, off, 31);
    off -= L.n;                                       // exclusive: first flat index of this lane's record
    const int geo = L.x0 | ((L.ya - row0) << 9) | (L.wdt << 15);
    int first = 0;                                    // records that start before `base`
    for (int base = 0; base < total; base += 32) {
        // the records with candidates (a prefix of the lanes, each with n >= 1) start at distinct flat indices: mark the starts that
        // fall into this window; the owner of flat index f is the last record starting at or before f
        const unsigned int mybit = (L.n > 0 && off >= base && off < base + 32) ? 1u << (off - base) : 0u;
        const unsigned int starts = __reduce_or_sync(F, mybit);
        const int f = base + lane;
        const int owner = first - 1 + __popc(starts & (0xffffffffu >> (31 - lane)));
        first += __popc(starts);
        const int o = owner < 0 ? 0 : owner;
        const float A0x = __shfl_sync(F, L.A0x, o), A0y = __shfl_sync(F, L.A0y, o), s00 = __shfl_sync(F, L.s00, o), s01 = __shfl_sync(F, L.s01, o);
        const float s10 = __shfl_sync(F, L.s10, o), s11 = __shfl_sync(F, L.s11, o), uz = __shfl_sync(F, L.uz, o);
        const int g = __shfl_sync(F, geo, o), ooff = __shfl_sync(F, off, o);
        const unsigned int zb = __shfl_sync(F, L.zb, o);
        bool have = false; unsigned int ent = 0u;
        if (f < total) {
            const int k = f - ooff, wdt = g >> 15;
            const int ky = (int)(((float)k + 0.5f) * __frcp_rn((float)wdt));        // exact k / wdt (see scan_lockstep)
            const int px = (g & 511) + (k - ky * wdt), pr = ((g >> 9) & 63) + ky;
            const int cls = cover_px(A0x, A0y, s00, s01, s10, s11, uz, px, row0 + pr);
            STAT(6, 1);
            if (cls) {
                STAT(7, 1);
                const unsigned int cur = reinterpret_cast<volatile unsigned int*>(&keys[pr * RR_W + px])[1];
                if (cur <= zb) { have = true; ent = (unsigned int)px | ((unsigned int)pr << 9) | ((unsigned int)cls << 15) | ((unsigned int)o << 17); }
                else STAT(8, 1);
            }
        }
        fq_push(Q, have, ent);
        if (Q.count >= 32) fq_drain(keys, row0, Q, 32, L);
    }
}

// shading of a pixel from barycentrics of the ORIGINAL triangle (Shader::fragment)
__device__ __forceinline__ unsigned int shade_bar(const float* __restrict__ M, const ShapeRenderConst& sh, const float* __restrict__ tv,
                                                  const uint8_t* __restrict__ texdata, const float* bar) {
    float nrm[3][3], uv[2][3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        float v[8]; load_corner(tv, j, v);
#pragma unroll
        for (int k = 0; k < 3; k++) { float r = 0.f; r += 0.f; r += M[28 + 3 * k + 2] * v[5]; r += M[28 + 3 * k + 1] * v[4]; r += M[28 + 3 * k] * v[3]; nrm[k][j] = r; }
        uv[0][j] = v[6]; uv[1][j] = v[7];
    }
    float bn[3];
#pragma unroll
    for (int i = 0; i < 3; i++) bn[i] = dot3r(nrm[i][0], nrm[i][1], nrm[i][2], bar);
    float nl = sqrtf(bn[0] * bn[0] + bn[1] * bn[1] + bn[2] * bn[2]);
    float k = 1.f / nl;
#pragma unroll
    for (int i = 0; i < 3; i++) bn[i] = bn[i] * k;
    float u = dot3r(uv[0][0], uv[0][1], uv[0][2], bar), v = dot3r(uv[1][0], uv[1][1], uv[1][2], bar);
    float ndl = dot3r(bn[0], bn[1], bn[2], c_render.light);
    float t2 = ndl * 2.f;
    float rd0 = bn[0] * t2 - c_render.light[0], rd1 = bn[1] * t2 - c_render.light[1], rd2 = bn[2] * t2 - c_render.light[2];
    float rl = sqrtf(rd0 * rd0 + rd1 * rd1 + rd2 * rd2);
    float rz = rd2 * (1.f / rl);
    float sp = rz > 0.f ? rz : 0.f;
    float specular = sp * sp;
    float diffuse = ndl > 0.f ? ndl : 0.f;
    int c[3]; texel_fetch(sh.tex, texdata, u, v, c);
    const float ambient = 0.6f, dc = 0.35f, sc = 0.05f, shadow = 1.0f;
    unsigned int packed = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        int ci = (int)((float)c[i] * sh.rgba[i]) & 0xff;
        float fc = ambient * (float)ci + shadow * (dc * diffuse + sc * specular) * (float)ci * 1.0f;
        int oc = 0;
        if (fc == fc) oc = (int)fc;
        oc = oc < 255 ? oc : 255;
        packed |= ((unsigned int)oc & 0xffu) << (8 * i);
    }
    return packed;
}

// slow path (winner is a piece of a near-clipped triangle): barycentrics of the original triangle from scratch
__device__ __noinline__ unsigned int shade_clipped(const float* __restrict__ M, const ShapeRenderConst& sh, const float* __restrict__ tv,
                                                   const uint8_t* __restrict__ texdata, int x, int y) {
    Tri T;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(tv + 8 * j));
        float px = a.x * 1.f, py = a.y * 1.f, pz = a.z * 1.f;
#pragma unroll
        for (int k = 0; k < 4; k++) T.c[k][j] = dot4r(M + 4 * k, px, py, pz, 1.f);
    }
    Screen S; screen_setup(T.c, S);
    Edge E; edge_setup(S, E);
    float bs[3], bar[3];
    frag_depth(S, E, T.c[2], x, y, bs, bar);
    return shade_bar(M, sh, tv, texdata, bar);
}

// colour of the winning fragment of pixel (x, y): barycentrics from the winner's record, then the fragment shader
__device__ __forceinline__ unsigned int shade_pixel(const float* __restrict__ M, const float* __restrict__ tris, const uint8_t* __restrict__ tri_shape,
                                                    const uint8_t* __restrict__ texdata, const Rec* __restrict__ recs, unsigned int lo, int x, int y, int& s_out) {
    const int gt = (int)(lo >> 15) - 1, ri = (int)(lo & 0x7fffu);
    const int s = tri_shape[gt];
    s_out = s;
    const ShapeRenderConst& sh = c_render.shape[s];
    const float* tv = tris + (size_t)gt * 24;
    const Rec rec = load_rec(recs + ri);
    STAT(11, 1);
    if (rec.flags & REC_CLIPPED) return shade_clipped(M + s * MAT_STRIDE, sh, tv, texdata, x, y);
    Screen S; Edge E; Recip R; float cz[3]; rec_unpack(rec, S, E, R, cz);
    float bs[3], bar[3];
    float ux, uy; bary_u(S, E, (float)x, (float)y, ux, uy);
    if (R.tight && coverage_class(ux, uy, E.uz) == 1) frag_bar_in(S, E, R, ux, uy, bar);
    else frag_bar_r(S, E, R, x, y, bs, bar);
    return shade_bar(M + s * MAT_STRIDE, sh, tv, texdata, bar);
}

// Persistent kernel: each CTA takes envs from a global work counter.  Per env:
//   geometry pass 1: one triangle per thread, model-space plane test against the camera: clearly back-facing triangles are
//              dropped, the rest compacted into a survivor list (bit 15 marks "too close to call");
//   geometry pass 2: survivors only — vertex shader, exact fp64 back-face test where pass 1 could not decide, near-plane
//              clip, viewport, bounding box, edge and reciprocal set-up, coverage pre-test of tiny boxes; what is left
//              becomes 80-byte records in the CTA's private scratch and is appended, per strip its box touches, to one
//              of three work lists by in-strip box size (small / medium / huge);
//   per strip (2 barriers): raster — huge records by groups of HUGE_GROUP threads, then medium records one per warp, then
//              small records one per lane in 32-record chunks, all self-scheduled through shared counters so that the
//              coarse items are handed out first; atomicMax on (depth bits, triangle index) with a depth-bound
//              rejection against the current key.  resolve — warps pull groups of 128 pixels: fragment shader for the
//              winners (lane per pixel), then packed row stores; the keys are cleared on the way out.
//
// MODE 0: every shape is rasterised for every env.  The table and the robot base never move and the eye camera is fixed, so
// their part of the image is the same in every env and every step: MODE 1 rasterises only those static shapes (one CTA,
// one env, at rr_create) and leaves, per pixel, the winning raster key (record bits = 0x7fff as the "static" marker) in
// SL.keys and the finished static image in SL.rgb / depth / seg; MODE 2 starts every strip's z-buffer from SL.keys instead
// of zero, rasterises only the moving shapes on top and, for pixels a static key still owns at resolve time, copies the
// finished static pixel.  max over (depth, triangle index) is associative, so the image is bit-identical to MODE 0's.
struct StaticLayer {
    unsigned long long* keys;                                       // [240][320] in raster-row order
    const uint8_t* rgb; const float* depth; const int32_t* seg;    // the finished static image in output layout (MODE 1 writes it as env 0)
    unsigned int bits; int skip_front;                              // bit s: shape s is static; triangles [0, skip_front) all belong to static shapes
};
template <int MODE>
__global__ void __launch_bounds__(RTHREADS, RBLOCKS)
k_raster(const float* __restrict__ mats, const float* __restrict__ tris, const uint8_t* __restrict__ tri_shape, const float4* __restrict__ planes, int n_tris,
         const uint8_t* __restrict__ texdata, const uint8_t* __restrict__ env_mask, Rec* __restrict__ rec_scratch, unsigned short* __restrict__ list_scratch,
         unsigned short* __restrict__ cull_scratch, int* __restrict__ work_counter,
         uint8_t* __restrict__ rgb, float* __restrict__ depth, int32_t* __restrict__ seg, int env_begin, int N, const StaticLayer SL, const LateEnvs LE, const int split) {
    // split > 1 (small batches): an env is `split` work items, item k taking the strips k, k + split, ... — every item does the env's whole
    // geometry (its own scratch), the strips are independent, so a few envs spread over split x as many CTAs.
    unsigned long long* __restrict__ static_keys = SL.keys;
    const unsigned int static_bits = SL.bits; const int skip_front = SL.skip_front;
    RR_DYN_SMEM(smem_raw);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    float* Mbuf = reinterpret_cast<float*>(smem_raw + (size_t)STRIP_ROWS * RR_W * 8);       // two matrix tiles with TMA_MATS
    unsigned int* fq_all = reinterpret_cast<unsigned int*>(smem_raw + (size_t)STRIP_ROWS * RR_W * 8 + (TMA_MATS ? 2 : 1) * RR_NSHAPES * MAT_STRIDE * 4);   // per-warp fragment queues
    __shared__ __align__(8) unsigned long long s_mbar[2];
    __shared__ int s_envq[2];
    __shared__ int s_eid, s_late, s_part;
    __shared__ int s_env, s_nrec, s_ncull, s_chunk, s_medc, s_rgrp, s_nlist[3 * NSTRIPS];   // s_nlist[class][strip]
    const int tid = threadIdx.x, lane = tid & 31;
    Rec* recs = rec_scratch + (size_t)blockIdx.x * REC_MAX;
    unsigned short* lists = list_scratch + (size_t)blockIdx.x * NSTRIPS * 2 * REC_MAX;
    unsigned short* cull = cull_scratch + (size_t)blockIdx.x * CULL_MAX;
    const float nearP = c_render.depthNear, farP = c_render.depthFar;
    const float bg_depth = (float)((double)(farP * (nearP + 1e30f)) / (2. * (double)farP * (double)nearP + (double)(farP * 1e30f) - (double)(nearP * 1e30f)));
    // the key buffer is zero whenever a strip starts: cleared here once, then by the resolve pass that consumes it
    for (int i = tid; i < STRIP_ROWS * RR_W; i += RTHREADS) keys[i] = (MODE == 2) ? static_keys[i] : 0ull;

#if TMA_MATS
    // The per-env matrix tile (22 shapes x 44 floats = 3 872 B) is staged by the TMA engine one env ahead: thread 0 takes
    // the next env from the work counter as soon as the current one starts, arms that tile's mbarrier and issues the bulk
    // copy; everybody waits on the mbarrier of the current tile (phase = use count of the buffer) instead of loading.
    constexpr unsigned MAT_BYTES = RR_NSHAPES * MAT_STRIDE * sizeof(float);
    static_assert(MAT_BYTES % 16 == 0, "bulk copies move multiples of 16 bytes");
    if (tid == 0) {
        mbar_init(&s_mbar[0], 1); mbar_init(&s_mbar[1], 1);
        mbar_fence_init();
        const int first = atomicAdd(work_counter, 1);
        s_envq[0] = first;
        if (first < N) tma_load_1d(Mbuf, mats + (size_t)(env_begin + first) * RR_NSHAPES * MAT_STRIDE, MAT_BYTES, &s_mbar[0]);
    }
#endif
    for (int it = 0;; it++) {
        __syncthreads();
#if TMA_MATS
        const int b = it & 1;
        const int cur_env = s_envq[b];
        float* M = Mbuf + b * (RR_NSHAPES * MAT_STRIDE);
        if (tid == 0) {
            s_nrec = 0; s_ncull = 0; s_chunk = 0; s_medc = 0; s_rgrp = 0;
            // tile b ^ 1 was last read during the previous env, which every thread left before the barrier above
            const int nx = cur_env < N ? atomicAdd(work_counter, 1) : N;
            s_envq[b ^ 1] = nx;
            if (nx < N) tma_load_1d(Mbuf + (b ^ 1) * (RR_NSHAPES * MAT_STRIDE), mats + (size_t)(env_begin + nx) * RR_NSHAPES * MAT_STRIDE, MAT_BYTES, &s_mbar[b ^ 1]);
        }
        if (tid < 3 * NSTRIPS) s_nlist[tid] = 0;
        __syncthreads();
        if (cur_env >= N) break;
        mbar_wait(&s_mbar[b], (unsigned)(it >> 1) & 1u);
        const int e = env_begin + cur_env;
#else
        float* M = Mbuf;
        if (tid == 0) {
            const int item = atomicAdd(work_counter, 1);
            s_nrec = 0; s_ncull = 0; s_chunk = 0; s_medc = 0; s_rgrp = 0;
            s_env = split > 1 ? item / split : item; s_part = split > 1 ? item - s_env * split : 0;
            s_eid = s_env; s_late = 0;
            if (LE.order && s_env < N) {          // the solver's order backwards: lightest envs first, the heavy kernel's envs last
                int idx = N - 1 - s_env, n_heavy = 0, cls = 0;
                for (int k = 0; k < RR_SOLVE_CLASSES - 1; k++) { const int n = LE.order[k]; n_heavy += n; if (cls == k && idx >= n) { idx -= n; cls = k + 1; } }
                const int e_ = LE.order[RR_SOLVE_CLASSES + cls * N + idx];
                s_eid = e_;
                if (N - 1 - s_env < min(n_heavy, LE.heavy_cap)) {
                    s_late = 1;
#ifndef RR_SIMT_EMU_BUILD
                    const long long t0 = clock64();
                    while (*reinterpret_cast<const volatile int*>(LE.ready + e_) != LE.ticket) {
                        __nanosleep(500);
                        if (clock64() - t0 > 4000000000ll) { atomicExch(LE.timeout, 1); break; }      // ~2 s
                    }
                    __threadfence();
#endif
                }
            }
        }
        if (tid < 3 * NSTRIPS) s_nlist[tid] = 0;
        __syncthreads();
        if (s_env >= N) break;
        const int e = env_begin + s_eid;
        if (s_late) {
            if (tid < RR_NSHAPES) render_setup_shape(LE.st, nullptr, LE.mats, LE.n_objects, LE.static_frames, e, tid);
            __threadfence_block();
            __syncthreads();
        }
#endif
        const size_t img = (size_t)e * RR_W * RR_H;
        const bool on = env_mask ? (env_mask[e] != 0) : true;
        if (!on) {   // reference returns cached zero arrays when render is off (env.py:114-122, 273-275)
            for (int i = tid; i < RR_H * RR_W / 4 && s_part == 0; i += RTHREADS) {
                size_t o = img + 4 * (size_t)i;
                if (rgb) { unsigned int* p = reinterpret_cast<unsigned int*>(rgb + o * 3); p[0] = 0; p[1] = 0; p[2] = 0; }
                if (depth) *reinterpret_cast<float4*>(depth + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (seg) *reinterpret_cast<int4*>(seg + o) = make_int4(0, 0, 0, 0);
            }
            continue;
        }
#if !TMA_MATS
        for (int i = tid; i < RR_NSHAPES * MAT_STRIDE; i += RTHREADS) M[i] = mats[(size_t)e * RR_NSHAPES * MAT_STRIDE + i];
        __syncthreads();
#endif
        // ---- geometry pass 1: conservative back-face classification in MODEL space, survivors compacted.
        // planes[t] = (unit normal, offset) of triangle t scaled by (smallest altitude / 1e-5); Ms[40..42] = camera in model
        // coordinates; Ms[43] bounds |camera - any vertex| + altitude.  |sd| > Ms[43] means the camera is further from the
        // triangle's plane than any rounding of the fp32 world coordinates can move that plane, so the exact fp64 test of
        // tri_vertex() would decide the same way; everything else is marked uncertain (bit 15) and takes the exact test.
        const int n_scan = (MODE == 2) ? n_tris - skip_front : n_tris;    // MODE 2: the leading static shapes are not even scanned
#if REVERSE_ORDER
#define PASS1_TRI(i) (n_tris - 1 - (i))                        // last shapes first: objects and the arm's upper links are nearest the eye
#else
#define PASS1_TRI(i) ((n_tris - n_scan) + (i))
#endif
#if GEO_PREFETCH      // the shape id and the plane of the next round are in flight while this round is classified
        int sidx_n = 0; float4 pl_n = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < n_scan) { sidx_n = __ldg(tri_shape + PASS1_TRI(tid)); pl_n = __ldg(planes + PASS1_TRI(tid)); }
#endif
        for (int base = 0; base < n_scan; base += RTHREADS) {
            const int gt = PASS1_TRI(base + tid);
#if GEO_PREFETCH
            const int sidx_c = sidx_n; const float4 pl_c = pl_n;
            if (base + RTHREADS + tid < n_scan) { sidx_n = __ldg(tri_shape + PASS1_TRI(base + RTHREADS + tid)); pl_n = __ldg(planes + PASS1_TRI(base + RTHREADS + tid)); }
#endif
            int code = -1;
            if (base + tid < n_scan) {
#if GEO_PREFETCH
                const int sidx = sidx_c;
#else
                const int sidx = tri_shape[gt];
#endif
                const float* Ms = M + sidx * MAT_STRIDE;
                bool act = Ms[39] != 0.f;
                if (MODE == 1) act = act && ((static_bits >> sidx) & 1u);
                if (MODE == 2) act = act && !((static_bits >> sidx) & 1u);
                if (act) {
#if GEO_PREFETCH
                    const float4 pl = pl_c;
#else
                    const float4 pl = __ldg(planes + gt);
#endif
                    const float sd = pl.x * Ms[40] + pl.y * Ms[41] + pl.z * Ms[42] - pl.w;
                    STAT(0, 1);
                    if (sd > Ms[43]) code = gt;
                    else if (!(sd < -Ms[43])) code = gt | 0x8000;
#ifdef RR_STATS
                    else { Tri T2; if (tri_vertex(Ms, tris + (size_t)gt * 24, T2)) STAT(13, 1); }   // pre-test said back, exact says front: must stay 0
#endif
                }
            }
            const unsigned m = __ballot_sync(0xffffffffu, code >= 0);
            if (m) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&s_ncull, __popc(m));
                pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1u));
                if (code >= 0) cull[pos] = (unsigned short)code;
            }
        }
        __syncthreads();
        // ---- geometry pass 2: vertex shader, (exact cull for the uncertain ones), clip, viewport, set-up, records
        {
            const int ncull = s_ncull;
#if GEO_PREFETCH
            unsigned int code_n = tid < ncull ? cull[tid] : 0u;
#endif
            for (int i = tid; i < ncull; i += RTHREADS) {
#if GEO_PREFETCH
                const unsigned int code = code_n;
                if (i + RTHREADS < ncull) code_n = cull[i + RTHREADS];
#else
                const unsigned int code = cull[i];
#endif
                const unsigned int gt = code & 0x7fffu;
                const float* Ms = M + tri_shape[gt] * MAT_STRIDE;
                const float* tv = tris + (size_t)gt * 24;
                Tri T;
                if (code & 0x8000u) { STAT(15, 1); if (!tri_vertex(Ms, tv, T)) continue; }
                else {
                    tri_clip(Ms, tv, T);
#ifdef RR_STATS
                    { Tri T2; if (!tri_vertex(Ms, tv, T2)) STAT(14, 1); }       // pre-test said front, exact test says back: must stay 0
#endif
                }
                STAT(1, 1);
                const bool front = (T.c[3][0] >= 0 && T.c[3][1] >= 0 && T.c[3][2] >= 0);
                if (front) { Screen S; screen_setup(T.c, S); emit_record(recs, lists, &s_nrec, s_nlist, S, T.c[2], gt, 0u); }
                else if (!(T.c[3][0] < 0 && T.c[3][1] < 0 && T.c[3][2] < 0)) emit_clipped(recs, lists, &s_nrec, s_nlist, T, gt);
            }
        }
        __threadfence_block();
        __syncthreads();
        // ---- strips (two barriers each: raster | shade + store + clear), both phases self-scheduled
        const int part = s_part;
        if (MODE == 2 && split > 1) {          // the previous item's resolve pass left the static keys of some other strip behind
            __syncthreads();
            for (int i = tid; i < STRIP_ROWS * RR_W; i += RTHREADS) keys[i] = static_keys[(size_t)part * STRIP_ROWS * RR_W + i];
            __syncthreads();
        }
        for (int strip = part; strip < NSTRIPS; strip += split) {
            const int row0 = strip * STRIP_ROWS;
            const int next_row0 = strip + split >= NSTRIPS ? 0 : row0 + split * STRIP_ROWS;      // whose static keys the resolve pass leaves in the key buffer
            const int nsmall = min(s_nlist[strip], REC_MAX), nmed = min(s_nlist[NSTRIPS + strip], REC_MAX), nbig = min(s_nlist[2 * NSTRIPS + strip], REC_MAX);
            const unsigned short* list = lists + (size_t)strip * 2 * REC_MAX;
            if (tid == 0) s_rgrp = 0;
            // ---- raster: one loop, one inlined copy of the pixel code (the kernel is instruction-cache sensitive).  A warp
            // works through three phases, coarse items first: 0 = huge boxes (mostly the table) shared by groups of HUGE_GROUP
            // threads, 1 = medium boxes, MED_LANES lanes per record, pulled 32 / MED_LANES at a time, 2 = small boxes, one
            // record per lane, pulled in 32-record chunks.
            {
                int phase = 0, hb = tid / HUGE_GROUP;
                FragQueue FQ; FQ.q = fq_all + (tid >> 5) * FQ_SIZE; FQ.head = 0; FQ.count = 0;
                for (;;) {
                    unsigned int ridx = 0; int start = 0, stride = 1; bool valid = false;
                    if (phase == 0) {
                        if (hb >= nbig) { phase = 1; continue; }
                        ridx = list[REC_MAX - 1 - hb]; start = tid % HUGE_GROUP; stride = HUGE_GROUP; valid = true;
                        hb += RTHREADS / HUGE_GROUP;
                    } else if (phase == 1) {
                        int mi = 0;
                        if (lane == 0) mi = atomicAdd(&s_medc, 32 / MED_LANES);
                        mi = __shfl_sync(0xffffffffu, mi, 0);
                        if (mi >= nmed) { phase = 2; continue; }
                        const int my = mi + lane / MED_LANES;
                        valid = my < nmed;
                        if (valid) ridx = list[REC_MAX + my];
                        start = lane % MED_LANES; stride = MED_LANES;
                    } else {
                        int rb = 0;
                        if (lane == 0) rb = atomicAdd(&s_chunk, 32);
                        rb = __shfl_sync(0xffffffffu, rb, 0);
                        if (rb >= nsmall) break;
                        valid = rb + lane < nsmall;
                        if (valid) ridx = list[rb + lane];
                    }
                    // RASTER_FLAT bits: 1 small boxes flattened, 2 medium boxes through the fragment queue, 4 huge boxes through it
                    if ((phase == 2 && (RASTER_FLAT & 1)) || (phase == 1 && (RASTER_FLAT & 2)) || (phase == 0 && (RASTER_FLAT & 4))) {
                        // (every lane of the warp goes through here: the scan and the drain are warp-wide)
                        LaneRec L; lane_rec_load(recs, ridx, valid, row0, L);
                        if (phase == 2) scan_flat(keys, row0, L, FQ);
                        else scan_lockstep(keys, row0, L, start, stride, FQ);
                        while (FQ.count > 0) fq_drain(keys, row0, FQ, min(FQ.count, 32), L);       // the records change with the next work item
                    } else if (valid) {
                        const Rec rec = load_rec(recs + ridx);
                        Screen S; Edge E; Recip R; float cz[3]; rec_unpack(rec, S, E, R, cz);
                        raster_pixels(keys, S, E, R, cz, rec.keylo, !(rec.flags & REC_CLIPPED), row0, start, stride);
                    }
                }
            }
            __syncthreads();
            if (tid == 0) { s_chunk = 0; s_medc = 0; }
            // ---- resolve: warps pull groups of 128 consecutive pixels.  A: lane per pixel (4 rounds), the winner's colour
            // replaces the low key word (1 << 31 | shape << 24 | rgb); B: 4 consecutive pixels per lane, packed 12/16/16-byte
            // stores with the vertical flip folded in, and the keys are cleared for the next strip.
            for (;;) {
                int g = 0;
                if (lane == 0) g = atomicAdd(&s_rgrp, 1);
                g = __shfl_sync(0xffffffffu, g, 0);
                if (g >= STRIP_ROWS * RR_W / 128) break;
                bool covered = false, moving = false;
                if (MODE == 2) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const unsigned long long key = keys[g * 128 + j * 32 + lane];
                        covered |= key != 0ull; moving |= key != 0ull && (key & 0x7fffull) != 0x7fffull;
                    }
                    if (!__any_sync(0xffffffffu, moving) && __any_sync(0xffffffffu, covered)) {
                        // no moving shape won a pixel of this group: the finished static image is the answer
                        const int i = g * 32 + lane;
                        const int r = i / (RR_W / 4), x4 = i % (RR_W / 4);
                        const size_t po = (size_t)(RR_H - 1 - (row0 + r)) * RR_W + 4 * x4, o = img + po;
                        if (rgb) {
                            const unsigned int* q = reinterpret_cast<const unsigned int*>(SL.rgb + po * 3);
                            unsigned int* p = reinterpret_cast<unsigned int*>(rgb + o * 3);
                            const unsigned int a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
                            __stcs(p, a); __stcs(p + 1, b); __stcs(p + 2, c);
                        }
                        if (depth) __stcs(reinterpret_cast<float4*>(depth + o), __ldg(reinterpret_cast<const float4*>(SL.depth + po)));
                        if (seg) __stcs(reinterpret_cast<int4*>(seg + o), __ldg(reinterpret_cast<const int4*>(SL.seg + po)));
                        const ulonglong2* nk = reinterpret_cast<const ulonglong2*>(static_keys + (size_t)(next_row0 + r) * RR_W + 4 * x4);
                        ulonglong2* kq = reinterpret_cast<ulonglong2*>(keys + r * RR_W + 4 * x4);
                        __syncwarp();                             // the keys replaced here were read by other lanes above
                        kq[0] = __ldg(nk); kq[1] = __ldg(nk + 1);
                        continue;
                    }
                    covered = false;
                }
#pragma unroll 1
                for (int j = 0; j < 4; j++) {
                    const int p = g * 128 + j * 32 + lane;
                    const unsigned long long key = keys[p];
                    if (MODE == 1) static_keys[row0 * RR_W + p] = key ? (key | 0x7fffull) : 0ull;
                    if (key == 0ull) continue;
                    covered = true;
                    if (MODE == 2 && (key & 0x7fffull) == 0x7fffull) continue;   // a static key survived: pass B takes the finished static pixel
                    const int r = p / RR_W, x = p - r * RR_W;
                    int s;
                    const unsigned int col = shade_pixel(M, tris, tri_shape, texdata, recs, (unsigned int)(key & 0xffffffffu), x, row0 + r, s);
                    // bit 31 marks a resolved key (a raster key's low word is (triangle + 1) << 15 | record < 2^30)
                    keys[p] = (key & 0xffffffff00000000ull) | 0x80000000ull | ((unsigned long long)s << 24) | (unsigned long long)(col & 0xffffffu);
                }
                const int i = g * 32 + lane;
                const int r = i / (RR_W / 4), x4 = i % (RR_W / 4);
                // MODE 2: the keys this lane consumes are replaced by the static keys of the same pixels of the next strip
                const ulonglong2* nk = reinterpret_cast<const ulonglong2*>(static_keys + (size_t)(next_row0 + r) * RR_W + 4 * x4);
                if (!__any_sync(0xffffffffu, covered)) {          // nothing but background in these 128 pixels (keys already zero)
                    if (MODE == 2) {
                        ulonglong2* kq = reinterpret_cast<ulonglong2*>(keys + r * RR_W + 4 * x4);
                        __syncwarp();
                        kq[0] = __ldg(nk); kq[1] = __ldg(nk + 1);
                    }
                    const size_t o = img + (size_t)(RR_H - 1 - (row0 + r)) * RR_W + 4 * x4;
                    if (rgb) {
                        unsigned int* p = reinterpret_cast<unsigned int*>(rgb + o * 3);
                        __stcs(p, 0xffffffffu); __stcs(p + 1, 0xffffffffu); __stcs(p + 2, 0xffffffffu);
                    }
                    if (depth) __stcs(reinterpret_cast<float4*>(depth + o), make_float4(bg_depth, bg_depth, bg_depth, bg_depth));
                    if (seg) __stcs(reinterpret_cast<int4*>(seg + o), make_int4(-1, -1, -1, -1));
                    continue;
                }
                __syncwarp();                                     // the colours written above are read by other lanes below
                unsigned int col[4]; float dep[4]; int sg[4];
                ulonglong2* kp = reinterpret_cast<ulonglong2*>(keys + r * RR_W + 4 * x4);
                const ulonglong2 ka = kp[0], kb = kp[1];
                if (MODE == 2) { kp[0] = __ldg(nk); kp[1] = __ldg(nk + 1); }
                else { kp[0] = make_ulonglong2(0ull, 0ull); kp[1] = make_ulonglong2(0ull, 0ull); }
                const unsigned long long k4[4] = {ka.x, ka.y, kb.x, kb.y};
                const size_t po = (size_t)(RR_H - 1 - (row0 + r)) * RR_W + 4 * x4;
                unsigned int scol[4] = {0u, 0u, 0u, 0u}; float4 sdep = make_float4(0.f, 0.f, 0.f, 0.f); int4 sseg = make_int4(0, 0, 0, 0);
                if (MODE == 2 && ((ka.x | ka.y | kb.x | kb.y) & 0x7fffull) == 0x7fffull) {     // (necessary, not sufficient: harmless extra loads)
                    const unsigned int* q = reinterpret_cast<const unsigned int*>(SL.rgb + po * 3);
                    const unsigned int a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
                    scol[0] = a & 0xffffffu; scol[1] = (a >> 24) | ((b & 0xffffu) << 8); scol[2] = (b >> 16) | ((c & 0xffu) << 16); scol[3] = c >> 8;
                    if (depth) sdep = __ldg(reinterpret_cast<const float4*>(SL.depth + po));
                    if (seg) sseg = __ldg(reinterpret_cast<const int4*>(SL.seg + po));
                }
                const float sdep4[4] = {sdep.x, sdep.y, sdep.z, sdep.w}; const int sseg4[4] = {sseg.x, sseg.y, sseg.z, sseg.w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const unsigned long long key = k4[k];
                    if (key == 0ull) { col[k] = 0xffffffu; sg[k] = -1; dep[k] = bg_depth; continue; }
                    if (MODE == 2 && !(key & 0x80000000ull)) { col[k] = scol[k]; sg[k] = sseg4[k]; dep[k] = sdep4[k]; continue; }   // static pixel
                    const unsigned int lo = (unsigned int)(key & 0xffffffffu);
                    col[k] = lo & 0xffffffu;
                    sg[k] = c_render.shape[(lo >> 24) & 31].objectAndLink & 0xffffff;
                    const float z_c = -unorder_bits((unsigned int)(key >> 32));
                    dep[k] = (float)((double)(farP * (nearP + z_c)) / (2. * (double)farP * (double)nearP + (double)(farP * z_c) - (double)(nearP * z_c)));
                }
                const size_t o = img + po;
                if (rgb) {
                    unsigned int* p = reinterpret_cast<unsigned int*>(rgb + o * 3);
                    __stcs(p, (col[0] & 0xffffffu) | (col[1] << 24));
                    __stcs(p + 1, ((col[1] >> 8) & 0xffffu) | (col[2] << 16));
                    __stcs(p + 2, ((col[2] >> 16) & 0xffu) | (col[3] << 8));
                }
                if (depth) __stcs(reinterpret_cast<float4*>(depth + o), make_float4(dep[0], dep[1], dep[2], dep[3]));
                if (seg) __stcs(reinterpret_cast<int4*>(seg + o), make_int4(sg[0], sg[1], sg[2], sg[3]));
            }
            __syncthreads();
        }
    }
}

// --------------------------------------------------------------------------------------------- host side
#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(_e); return -1; } } while (0)

static const size_t RASTER_SMEM = (size_t)STRIP_ROWS * RR_W * 8 + (TMA_MATS ? 2 : 1) * RR_NSHAPES * MAT_STRIDE * 4 + (RASTER_FLAT ? (RTHREADS / 32) * FQ_SIZE * 4 : 0);
static double* g_static_frames[16] = {0};

// Per-triangle model-space plane for the conservative back-face pre-test: unit normal n (orientation of e1 x e2) and
// offset n.v0, both scaled by h / 1e-5 where h is the triangle's smallest altitude.  A vertex rounding error d moves
// the plane seen from distance D by at most d (1 + 2 D / h); fp32 world coordinates carry d <= 1.2e-6, so a camera
// whose scaled plane distance exceeds D + h (the bound k_render_setup stores) is on the same side for the exact test.
static cudaError_t build_planes(rr_ctx* ctx) {
    const int T = ctx->n_tris;
    std::vector<float> tv((size_t)T * 24);
    cudaError_t e = cudaMemcpy(tv.data(), ctx->d_tris, tv.size() * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return e;
    std::vector<float> pl((size_t)T * 4, 0.f);
    double hmax = 0.0;
    for (int t = 0; t < T; t++) {
        double v[3][3];
        for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) v[j][k] = (double)tv[(size_t)t * 24 + 8 * j + k];
        double e1[3], e2[3], e3[3];
        for (int k = 0; k < 3; k++) { e1[k] = v[1][k] - v[0][k]; e2[k] = v[2][k] - v[0][k]; e3[k] = v[2][k] - v[1][k]; }
        const double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double a2 = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        const double l1 = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]), l2 = std::sqrt(e2[0] * e2[0] + e2[1] * e2[1] + e2[2] * e2[2]);
        const double l3 = std::sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
        const double lmax = std::max(l1, std::max(l2, l3));
        if (!(a2 > 0.0) || !(lmax > 0.0)) continue;                 // degenerate: zero plane => always "uncertain"
        const double h = a2 / lmax, sc = h / 1e-5 / a2;
        if (!(h > 1e-9) || !std::isfinite(sc)) continue;
        hmax = std::max(hmax, h);
        for (int k = 0; k < 3; k++) pl[(size_t)t * 4 + k] = (float)(n[k] * sc);
        pl[(size_t)t * 4 + 3] = (float)((n[0] * v[0][0] + n[1] * v[0][1] + n[2] * v[0][2]) * sc);
    }
    const float hm = (float)hmax;
    ctx->cull_hmax = hm;
    e = cudaMemcpyToSymbol(c_cull_hmax, &hm, sizeof(float));
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&ctx->d_planes, pl.size() * sizeof(float));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(ctx->d_planes, pl.data(), pl.size() * sizeof(float), cudaMemcpyHostToDevice);
}

static rr_ctx* g_render_owner[64];
static int bind_render(rr_ctx* ctx, cudaStream_t s) {
    rr_ctx*& owner = g_render_owner[ctx->device & 63];
    if (owner != ctx) {
        CK(cudaDeviceSynchronize());          // kernels of the previous owner may still be reading the old constants
        CK(cudaMemcpyToSymbolAsync(c_render, &ctx->render, sizeof(RenderConst), 0, cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyToSymbolAsync(c_cull_hmax, &ctx->cull_hmax, sizeof(float), 0, cudaMemcpyHostToDevice, s));
        owner = ctx;
    }
    return 0;
}
void rrk_render_unbind(rr_ctx* ctx) { if (g_render_owner[ctx->device & 63] == ctx) g_render_owner[ctx->device & 63] = nullptr; }

int rrk_render_upload_constants(rr_ctx* ctx) {
    CK(cudaMemcpyToSymbol(c_render, &ctx->render, sizeof(RenderConst)));
    g_render_owner[ctx->device & 63] = ctx;
    CK(cudaFuncSetAttribute(k_raster<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RASTER_SMEM));
    CK(cudaFuncSetAttribute(k_raster<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RASTER_SMEM));
    CK(cudaFuncSetAttribute(k_raster<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RASTER_SMEM));
    {
        cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, ctx->device));
        int per_sm = 0;
        int per_sm0 = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm0, k_raster<0>, RTHREADS, RASTER_SMEM));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_raster<2>, RTHREADS, RASTER_SMEM));
        if (per_sm0 < per_sm) per_sm = per_sm0;
        if (per_sm < 1) { ctx->err = "k_raster does not fit on this device"; return -1; }
        ctx->raster_grid = prop.multiProcessorCount * per_sm;           // persistent: one wave of CTAs, envs pulled from a counter
        if (ctx->raster_grid > ctx->N * NSTRIPS) ctx->raster_grid = ctx->N * NSTRIPS;      // (small batches: up to one CTA per strip)
        CK(cudaMalloc(&ctx->d_rec_scratch, (size_t)ctx->raster_grid * REC_MAX * sizeof(Rec)));
        CK(cudaMalloc(&ctx->d_work_counter, sizeof(int)));
        CK(cudaMalloc(&ctx->d_list_scratch, (size_t)ctx->raster_grid * NSTRIPS * 2 * REC_MAX * sizeof(unsigned short)));
        std::vector<uint8_t> ts(ctx->n_tris);
        for (int sidx = 0; sidx < RR_NSHAPES; sidx++)
            for (int t = 0; t < ctx->render.shape[sidx].tri_count; t++) ts[ctx->render.shape[sidx].tri_start + t] = (uint8_t)sidx;
        CK(cudaMalloc(&ctx->d_tri_shape, ts.size()));
        CK(cudaMemcpy(ctx->d_tri_shape, ts.data(), ts.size(), cudaMemcpyHostToDevice));
        if (ctx->n_tris >= 0x8000) { ctx->err = "k_raster: triangle index does not fit 15 bits"; return -1; }
        CK(build_planes(ctx));
        CK(cudaMalloc(&ctx->d_cull_scratch, (size_t)ctx->raster_grid * CULL_MAX * sizeof(unsigned short)));
        // static layer: the robot base and the table (bodies whose frames come from static_frames below)
        ctx->static_bits = 0; ctx->static_skip_front = 0; ctx->static_ready = false;
        bool leading = true;
        for (int sidx = 0; sidx < RR_NSHAPES; sidx++) {
            const bool stat = (ctx->robot.shape_body[sidx] == 0 && ctx->robot.shape_link[sidx] < 0) || ctx->robot.shape_body[sidx] == 1;
            if (stat) ctx->static_bits |= 1u << sidx;
            if (stat && leading && ctx->render.shape[sidx].tri_start == ctx->static_skip_front) ctx->static_skip_front += ctx->render.shape[sidx].tri_count;
            else leading = false;
        }
        const char* sl = getenv("RR_STATIC_LAYER");                  // RR_STATIC_LAYER=0: rasterise every shape for every env (A/B measurements)
        ctx->static_enabled = !(sl && sl[0] == '0');
        CK(cudaMalloc(&ctx->d_static_keys, (size_t)RR_W * RR_H * sizeof(unsigned long long)));
        CK(cudaMalloc(&ctx->d_static_img, (size_t)RR_W * RR_H * 12));      // rgb (3 B, padded to 4) | depth f32 | seg i32
    }
    // static frames (robot base, table) as doubles
    double sf[RR_NSHAPES * 12];
    for (int s = 0; s < RR_NSHAPES; s++) {
        double* f = sf + 12 * s;
        int body = ctx->robot.shape_body[s], link = ctx->robot.shape_link[s];
        for (int k = 0; k < 12; k++) f[k] = 0;
        if (body == 0) { f[0] = f[4] = f[8] = 1.0; for (int k = 0; k < 3; k++) f[9 + k] = ctx->robot.base_pos[k]; }
        else if (body == 1) { for (int k = 0; k < 9; k++) f[k] = ctx->robot.table_R[k]; for (int k = 0; k < 3; k++) f[9 + k] = link < 0 ? ctx->robot.table_pos[k] : ctx->robot.table_upper_pos[k]; }
    }
    if (ctx->device < 16) {
        if (!g_static_frames[ctx->device]) CK(cudaMalloc(&g_static_frames[ctx->device], sizeof sf));
        CK(cudaMemcpy(g_static_frames[ctx->device], sf, sizeof sf, cudaMemcpyHostToDevice));
    }
    return 0;
}

static StaticLayer static_layer(rr_ctx* ctx) {
    StaticLayer SL;
    SL.keys = (unsigned long long*)ctx->d_static_keys;
    SL.rgb = (const uint8_t*)ctx->d_static_img; SL.depth = (const float*)((const uint8_t*)ctx->d_static_img + (size_t)RR_W * RR_H * 4);
    SL.seg = (const int32_t*)((const uint8_t*)ctx->d_static_img + (size_t)RR_W * RR_H * 8);
    SL.bits = ctx->static_bits; SL.skip_front = ctx->static_skip_front;
    return SL;
}
static int render_range(rr_ctx* ctx, const uint8_t* mask, const double* frames, uint8_t* rgb, float* depth, int32_t* seg,
                        int env_begin, int env_count, bool eye_camera, cudaStream_t s) {
    if (env_count <= 0) return 0;
    if (bind_render(ctx, s)) return -1;
    if (ctx->h_late_timeout && *ctx->h_late_timeout) { ctx->err = "k_raster timed out waiting for k_step_solve_heavy (an earlier rr_render returned stale images)"; return -1; }
    // right after rr_step the solver's heavy kernel may still run on the context's side stream: the eye-camera render of all envs
    // does not wait for it here but per env inside k_raster (LateEnvs); everything else joins it first
    // (see rr_ctx::heavy_mode: the split pays when this render is the first thing after the step — a caller that reads the state first has joined the heavy kernel already)
    if (eye_camera && frames == nullptr && env_begin == 0 && env_count == ctx->N && !ctx->touched_since_step) ctx->rendered_since_step = true;
    const bool late = ctx->heavy_pending && eye_camera && frames == nullptr && env_begin == 0 && env_count == ctx->N && !TMA_MATS;
    if (!late && rrk_join_heavy(ctx, s) != 0) return -1;
    LateEnvs LE; memset(&LE, 0, sizeof LE);
    if (late) {
        if (!ctx->h_late_timeout) { CK(cudaMallocHost(&ctx->h_late_timeout, sizeof(int))); *ctx->h_late_timeout = 0; CK(cudaMalloc((void**)&ctx->d_late_timeout, sizeof(int))); CK(cudaMemsetAsync(ctx->d_late_timeout, 0, sizeof(int), s)); }
        LE.order = ctx->st.solve_order; LE.ready = ctx->st.ready; LE.ticket = ctx->step_serial; LE.heavy_cap = ctx->heavy_cap;
        LE.st = ctx->st; LE.n_objects = ctx->n_objects; LE.static_frames = g_static_frames[ctx->device]; LE.mats = ctx->st.render_mats;
        LE.timeout = ctx->d_late_timeout;
    }
    int total = env_count * RR_NSHAPES;
    {
        ProfScope ps(ctx, 1, s);
        RR_LAUNCH(k_render_setup, (total + 127) / 128, 128, 0, s)(ctx->st, frames, ctx->st.render_mats, ctx->n_objects, g_static_frames[ctx->device], env_begin, env_count);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    // the static layer applies to the simulated scene seen by the eye camera; caller-supplied frames (rr_render_frames) and
    // other cameras (rr_render_camera) rasterise everything
    const bool use_static = ctx->static_enabled && ctx->static_ready && eye_camera && frames == nullptr;
    const StaticLayer SL = static_layer(ctx);
#define RASTER_ARGS(RGB, DEP, SEG, MASK, NENV) ctx->st.render_mats, ctx->d_tris, ctx->d_tri_shape, (const float4*)ctx->d_planes, ctx->n_tris, \
        ctx->d_tex, MASK, (Rec*)ctx->d_rec_scratch, (unsigned short*)ctx->d_list_scratch, (unsigned short*)ctx->d_cull_scratch, \
        ctx->d_work_counter, RGB, DEP, SEG, env_begin, NENV, SL, LE, split
    CK(cudaMemsetAsync(ctx->d_work_counter, 0, sizeof(int), s));
    {
        ProfScope ps(ctx, 2, s);
        // small batches: the six strips of an env go to up to six CTAs (see k_raster); not beside the late-env path, whose CTAs write the env's matrices
        int split = 1;
        if (!late && !TMA_MATS) for (int sp : {6, 3, 2}) if (NSTRIPS % sp == 0 && env_count * sp <= ctx->raster_grid) { split = sp; break; }
        int grid = ctx->raster_grid < env_count * split ? ctx->raster_grid : env_count * split;
        // late mode: the heavy kernel's narrowphase follows it while this kernel runs — leave it two SMs (persistent CTAs never give way)
        if (late && grid > 4 * RASTER_LATE_SPARE) grid -= RASTER_LATE_SPARE;
        if (use_static) RR_LAUNCH(k_raster<2>, grid, RTHREADS, RASTER_SMEM, s)(RASTER_ARGS(rgb, depth, seg, mask, env_count));
        else RR_LAUNCH(k_raster<0>, grid, RTHREADS, RASTER_SMEM, s)(RASTER_ARGS(rgb, depth, seg, mask, env_count));
    }
#undef RASTER_ARGS
    ctx->launches++;
    CK(cudaGetLastError());
    if (late) {
        if (rrk_join_heavy(ctx, s) != 0) return -1;          // formally; k_raster has waited for every one of its envs
        CK(cudaMemcpyAsync(ctx->h_late_timeout, ctx->d_late_timeout, sizeof(int), cudaMemcpyDeviceToHost, s));
    }
    return 0;
}
// Once per context (rr_create, after the first reset): matrices of env 0, then the static shapes of that env — their
// matrices are the same in every env — into d_static_keys / d_static_img.  Synchronous, so that every later launch on
// any stream sees the finished layer.
int rrk_render_build_static(rr_ctx* ctx) {
    if (!ctx->static_enabled || ctx->static_ready) return 0;
    if (bind_render(ctx, 0)) return -1;
    const StaticLayer SL = static_layer(ctx);
    RR_LAUNCH(k_render_setup, 1, 128, 0, 0)(ctx->st, nullptr, ctx->st.render_mats, ctx->n_objects, g_static_frames[ctx->device], 0, 1);
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(ctx->d_work_counter, 0, sizeof(int), 0));
    RR_LAUNCH(k_raster<1>, 1, RTHREADS, RASTER_SMEM, 0)(ctx->st.render_mats, ctx->d_tris, ctx->d_tri_shape, (const float4*)ctx->d_planes, ctx->n_tris,
        ctx->d_tex, nullptr, (Rec*)ctx->d_rec_scratch, (unsigned short*)ctx->d_list_scratch, (unsigned short*)ctx->d_cull_scratch,
        ctx->d_work_counter, const_cast<uint8_t*>(SL.rgb), const_cast<float*>(SL.depth), const_cast<int32_t*>(SL.seg), 0, 1, SL, LateEnvs{}, 1);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    ctx->launches += 2;
    ctx->static_ready = true;
    return 0;
}
int rrk_render_range(rr_ctx* ctx, const uint8_t* mask, const double* frames, uint8_t* rgb, float* depth, int32_t* seg,
                     int env_begin, int env_count, cudaStream_t s) {
    return render_range(ctx, mask, frames, rgb, depth, seg, env_begin, env_count, true, s);
}
int rrk_render(rr_ctx* ctx, const uint8_t* mask, const double* frames, uint8_t* rgb, float* depth, int32_t* seg, cudaStream_t s) {
    return rrk_render_range(ctx, mask, frames, rgb, depth, seg, 0, ctx->N, s);
}

// Render with another camera (EnvCamera / render('rgb_array'), reference env.py:221-228, 470-513): the camera part of
// the constants is swapped in stream order around the two kernels and the eye camera restored afterwards.
int rrk_render_camera(rr_ctx* ctx, const RenderConst* cam, const uint8_t* mask, uint8_t* rgb, float* depth, int32_t* seg, cudaStream_t s) {
    if (bind_render(ctx, s)) return -1;
    CK(cudaMemcpyToSymbolAsync(c_render, cam, sizeof(RenderConst), 0, cudaMemcpyHostToDevice, s));
    int rc = render_range(ctx, mask, nullptr, rgb, depth, seg, 0, ctx->N, false, s);
    CK(cudaMemcpyToSymbolAsync(c_render, &ctx->render, sizeof(RenderConst), 0, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));          // `cam` lives on the caller's stack
    return rc;
}
