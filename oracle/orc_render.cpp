/* TEST INFRASTRUCTURE — CPU oracle: restatement of the software rasteriser the reference calls.
 *
 * Reference call sites: real_robots/envs/env.py:249-255 (get_retina), :536-567 (EyeCamera.renderTarget:
 * computeViewMatrix, computeProjectionMatrixFOV(80, 4/3, .1, 100), getCameraImage(320,240) -> TinyRenderer),
 * outputs RGB u8[240,320,3] (RGBA sliced, :560-562), depth f32[240,320] (:564-565), mask i32 (:553,:567).
 * The arithmetic lives in pybullet's examples/TinyRenderer (our_gl.cpp, TinyRenderer.cpp, geometry.h, model.cpp)
 * and examples/SharedMemory/plugins/tinyRendererPlugin/TinyRendererVisualShapeConverter.cpp — un-vendored,
 * unpinned (setup.py:24-32).  PARITY UNPINNED: restated from the published algorithm.
 *
 * Everything is float32 with the same operation order as the original's vec/mat templates (dot products
 * accumulate from the LAST component down to the first, starting at 0).  Compile with -ffp-contract=off.
 */
#include "orc_render.h"
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

typedef float Mat4[4][4];

/* geometry.h: T operator*(vec,vec): ret=0; for(i=DIM;i--;) ret+=l[i]*r[i] */
static inline float dot4r(const float* a, const float* b) { float r = 0.f; r += a[3] * b[3]; r += a[2] * b[2]; r += a[1] * b[1]; r += a[0] * b[0]; return r; }
static inline float dot3r(const float* a, const float* b) { float r = 0.f; r += a[2] * b[2]; r += a[1] * b[1]; r += a[0] * b[0]; return r; }

static void mat4_mul(const Mat4 A, const Mat4 B, Mat4 C) {
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
        float col[4] = {B[0][j], B[1][j], B[2][j], B[3][j]};
        C[i][j] = dot4r(A[i], col);
    }
}
static void mat4_vec(const Mat4 A, const float* v, float* out) { for (int i = 0; i < 4; i++) out[i] = dot4r(A[i], v); }

/* generic determinant by cofactor expansion along row 0, accumulating from the last column (geometry.h dt<DIM>) */
static float det_n(const float* m, int n) {   /* m: n x n row-major */
    if (n == 1) return m[0];
    float ret = 0.f;
    for (int i = n; i--;) {
        float minor[9]; int mi = 0;
        for (int r = 1; r < n; r++) for (int c = 0; c < n; c++) if (c != i) minor[mi++] = m[r * n + c];
        float cof = det_n(minor, n - 1) * (float)((i % 2) ? -1 : 1);
        ret += m[i] * cof;
    }
    return ret;
}
static float cofactor4(const Mat4 M, int row, int col) {
    float minor[9]; int mi = 0;
    for (int r = 0; r < 4; r++) if (r != row) for (int c = 0; c < 4; c++) if (c != col) minor[mi++] = M[r][c];
    return det_n(minor, 3) * (float)(((row + col) % 2) ? -1 : 1);
}
/* mat::invert_transpose(): adjugate / (adj[0] . rows[0]) */
static void mat4_invert_transpose(const Mat4 M, Mat4 out) {
    Mat4 adj;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) adj[i][j] = cofactor4(M, i, j);
    float tmp = dot4r(adj[0], M[0]);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[i][j] = adj[i][j] / tmp;
}

void orc_eye_camera(float* view, float* proj) {
    /* b3ComputeViewMatrixFromPositions (float): eye (0.01,0,1.2) env.py:136, target = table position (0,0,0.08)
     * env.py:253-255, up (0,0,1) env.py:523 */
    float eye[3] = {0.01f, 0.f, 1.2f}, center[3] = {0.f, 0.f, 0.08f}, up[3] = {0.f, 0.f, 1.f};
    float f[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]};
    float fl = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
    for (int i = 0; i < 3; i++) f[i] = f[i] / fl;
    float ul = std::sqrt(up[0] * up[0] + up[1] * up[1] + up[2] * up[2]);
    float u[3] = {up[0] / ul, up[1] / ul, up[2] / ul};
    float s[3] = {f[1] * u[2] - f[2] * u[1], f[2] * u[0] - f[0] * u[2], f[0] * u[1] - f[1] * u[0]};
    float sl = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
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
    /* b3ComputeProjectionMatrixFOV(fov 80, aspect float(320)/240, near .1, far 100) env.py:548-551 */
    float fov = 80.f, aspect = (float)(320.0 / 240.0), nearVal = 0.1f, farVal = 100.f;
    float yScale = (float)(1.0 / std::tan((3.1415926535897932384626433832795029 / 180.0) * fov / 2));
    float xScale = yScale / aspect;
    for (int i = 0; i < 16; i++) proj[i] = 0.f;
    proj[0] = xScale; proj[5] = yScale;
    proj[10] = (nearVal + farVal) / (nearVal - farVal);
    proj[11] = -1.f;
    proj[14] = ((2) * farVal * nearVal) / (nearVal - farVal);
}

struct Raster {
    std::vector<float> z; std::vector<int32_t> seg; std::vector<uint8_t> rgb;   /* raster rows (origin bottom-left) */
};

struct ShaderCtx {
    const OrcTexture* tex; float rgba[4];
    float light[3];
    float nearPlane, farPlane;
    /* per-triangle varyings */
    float tri[4][3];      /* clip coords, column per vertex */
    float nrm[3][3];      /* world normals, column per vertex */
    float uv[2][3];
    float world[3][3];
};

static void texel(const OrcTexture* t, float u, float v, uint8_t* c) {
    if (!t) { c[0] = c[1] = c[2] = 255; return; }
    double ip;
    float uf = (float)std::modf((double)u, &ip);
    float vf = (float)std::modf((double)v, &ip);
    int x = (int)(uf * (float)t->w), y = (int)(vf * (float)t->h);
    if (x < 0 || y < 0 || x >= t->w || y >= t->h) { c[0] = c[1] = c[2] = 0; return; }
    const uint8_t* p = t->uniform ? t->data : t->data + ((size_t)x + (size_t)y * t->w) * 3;
    c[0] = p[0]; c[1] = p[1]; c[2] = p[2];
}

/* Shader::fragment */
static void fragment(const ShaderCtx& S, const float* bar, uint8_t* out) {
    float bn[3], uv[2];
    for (int i = 0; i < 3; i++) bn[i] = dot3r(S.nrm[i], bar);
    float nl = std::sqrt(bn[0] * bn[0] + bn[1] * bn[1] + bn[2] * bn[2]);
    float k = 1.f / nl;
    for (int i = 0; i < 3; i++) bn[i] = bn[i] * k;
    for (int i = 0; i < 2; i++) uv[i] = dot3r(S.uv[i], bar);
    float ndl = dot3r(bn, S.light);
    float t2 = ndl * 2.f;
    float rd[3] = {bn[0] * t2 - S.light[0], bn[1] * t2 - S.light[1], bn[2] * t2 - S.light[2]};
    float rl = std::sqrt(rd[0] * rd[0] + rd[1] * rd[1] + rd[2] * rd[2]);
    float rz = rd[2] * (1.f / rl);
    float sp = rz > 0.f ? rz : 0.f;
    float specular = sp * sp;                    /* pow(x, Model::specular()==2.0) */
    float diffuse = ndl > 0.f ? ndl : 0.f;
    uint8_t c[3]; texel(S.tex, uv[0], uv[1], c);
    const float ambient = 0.6f, dc = 0.35f, sc = 0.05f, shadow = 1.0f;
    for (int i = 0; i < 3; i++) {
        uint8_t ci = (uint8_t)((float)c[i] * S.rgba[i]);
        float fc = ambient * (float)ci + shadow * (dc * diffuse + sc * specular) * (float)ci * 1.0f;
        int oc = 0;
        if (fc == fc) oc = (int)fc;
        out[i] = (uint8_t)(oc < 255 ? oc : 255);
    }
}

static bool barycentric(const float* A, const float* B, const float* C, float px, float py, float* out) {
    float s0[3] = {C[0] - A[0], B[0] - A[0], A[0] - px};
    float s1[3] = {C[1] - A[1], B[1] - A[1], A[1] - py};
    float ux = s0[1] * s1[2] - s0[2] * s1[1];
    float uy = s0[2] * s1[0] - s0[0] * s1[2];
    float uz = s0[0] * s1[1] - s0[1] * s1[0];
    if ((double)std::fabs(uz) > 1e-2) {
        out[0] = 1.f - (ux + uy) / uz; out[1] = uy / uz; out[2] = ux / uz;
        return true;
    }
    out[0] = -1.f; out[1] = 1.f; out[2] = 1.f;
    return false;
}

struct ScreenTri { float pts[3][4]; float p2[3][2]; int x0, x1, y0, y1; bool any; };

static void screen_setup(const float clip[4][3], ScreenTri& T) {
    const float W = ORC_W, H = ORC_H;
    for (int i = 0; i < 3; i++) {
        float cx = clip[0][i], cy = clip[1][i], cz = clip[2][i], cw = clip[3][i];
        /* viewport * clip with reverse-order accumulation; zero entries contribute +-0 */
        float x = 0.f; x += (W / 2.f) * cw; x += 0.f * cz; x += 0.f * cy; x += (W / 2.f) * cx;
        float y = 0.f; y += (H / 2.f) * cw; y += 0.f * cz; y += (H / 2.f) * cy; y += 0.f * cx;
        float z = 0.f; z += .5f * cw; z += .5f * cz; z += 0.f * cy; z += 0.f * cx;
        float w = 0.f; w += 1.f * cw; w += 0.f * cz; w += 0.f * cy; w += 0.f * cx;
        T.pts[i][0] = x; T.pts[i][1] = y; T.pts[i][2] = z; T.pts[i][3] = w;
        T.p2[i][0] = x / w; T.p2[i][1] = y / w;
    }
    float mn[2] = {FLT_MAX, FLT_MAX}, mx[2] = {-FLT_MAX, -FLT_MAX}, cl[2] = {W - 1, H - 1};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 2; j++) {
        float a = mn[j] < T.p2[i][j] ? mn[j] : T.p2[i][j];     /* b3Min(a,b) = a<b?a:b */
        mn[j] = 0.f > a ? 0.f : a;                              /* b3Max(a,b) = a>b?a:b */
        float b = mx[j] > T.p2[i][j] ? mx[j] : T.p2[i][j];
        mx[j] = cl[j] < b ? cl[j] : b;
    }
    T.x0 = (int)mn[0]; T.y0 = (int)mn[1];
    /* loop "P.x <= bboxmax.x" with int P: last value = floor(bboxmax) when bboxmax >= x0 */
    T.any = ((float)T.x0 <= mx[0]) && ((float)T.y0 <= mx[1]);
    T.x1 = T.any ? (int)std::floor(mx[0]) : -1;
    T.y1 = T.any ? (int)std::floor(mx[1]) : -1;
}

static float frag_depth_and_bc(const ScreenTri& T, const float clipz[3], int px, int py, float* bc_screen, float* bc_clip) {
    barycentric(T.p2[0], T.p2[1], T.p2[2], (float)px, (float)py, bc_screen);
    float b[3] = {bc_screen[0] / T.pts[0][3], bc_screen[1] / T.pts[1][3], bc_screen[2] / T.pts[2][3]};
    float s = b[0] + b[1] + b[2];
    for (int i = 0; i < 3; i++) bc_clip[i] = b[i] / s;
    return -1.f * dot3r(clipz, bc_clip);
}

/* our_gl.cpp triangle() */
static void raster_triangle(Raster& R, const ShaderCtx& S, int objectAndLink) {
    ScreenTri T; screen_setup(S.tri, T);
    if (!T.any) return;
    for (int x = T.x0; x <= T.x1; x++) for (int y = T.y0; y <= T.y1; y++) {
        float bs[3], bc[3];
        float fd = frag_depth_and_bc(T, S.tri[2], x, y, bs, bc);
        int idx = x + y * ORC_W;
        if (bs[0] < 0 || bs[1] < 0 || bs[2] < 0 || R.z[idx] > fd) continue;
        if (fd != fd) continue;      /* deviation: a NaN depth (degenerate perspective divide) is dropped, never stored */
        if (fd < -S.farPlane) continue;
        if (fd > -S.nearPlane) continue;
        uint8_t col[3]; fragment(S, bc, col);
        R.z[idx] = fd; R.seg[idx] = objectAndLink;
        R.rgb[3 * idx] = col[0]; R.rgb[3 * idx + 1] = col[1]; R.rgb[3 * idx + 2] = col[2];
    }
}
/* our_gl.cpp triangleClipped(): coverage + depth from the clipped triangle, shading from the original */
static void raster_triangle_clipped(Raster& R, const ShaderCtx& S, const float clipped[4][3], int objectAndLink) {
    ScreenTri T; screen_setup(clipped, T);
    ScreenTri O; screen_setup(S.tri, O);
    if (!T.any) return;
    for (int x = T.x0; x <= T.x1; x++) for (int y = T.y0; y <= T.y1; y++) {
        float bs[3], bc[3];
        float fd = frag_depth_and_bc(T, clipped[2], x, y, bs, bc);
        int idx = x + y * ORC_W;
        if (bs[0] < 0 || bs[1] < 0 || bs[2] < 0 || R.z[idx] > fd) continue;
        if (fd != fd) continue;
        float bs2[3], bc2[3];
        frag_depth_and_bc(O, S.tri[2], x, y, bs2, bc2);
        uint8_t col[3]; fragment(S, bc2, col);
        R.z[idx] = fd; R.seg[idx] = objectAndLink;
        R.rgb[3 * idx] = col[0]; R.rgb[3 * idx + 1] = col[1]; R.rgb[3 * idx + 2] = col[2];
    }
}

static bool vec4_equals(const float* a, const float* b) {   /* TinyRenderer.cpp equals(): component-wise fabs < 1e-4? -> exact compare */
    return a[0] == b[0] && a[1] == b[1] && a[2] == b[2] && a[3] == b[3];
}
/* TinyRenderer.cpp clipEdge / clipTriangleAgainstNearplane; returns number of output triangles, -1 = not clipped (use original) */
static int clip_near(const float tri[4][3], float out[3][4][3]) {
    if (tri[3][0] < 0 && tri[3][1] < 0 && tri[3][2] < 0) return 0;
    if (tri[3][0] >= 0 && tri[3][1] >= 0 && tri[3][2] >= 0) return -1;
    float verts[8][4]; int nv = 0;
    const int ea[3] = {0, 1, 2}, eb[3] = {1, 2, 0};
    for (int e = 0; e < 3; e++) {
        float v0[4], v1[4];
        for (int k = 0; k < 4; k++) { v0[k] = tri[k][ea[e]]; v1[k] = tri[k][eb[e]]; }
        bool in0 = v0[3] > 0.f && v0[2] > -v0[3];
        bool in1 = v1[3] > 0.f && v1[2] > -v1[3];
        if (in0 && in1) {
        } else if (in0 || in1) {
            float d0 = v0[2] + v0[3], d1 = v1[2] + v1[3];
            float factor = (float)(1.0 / (double)(d1 - d0));
            float nv4[4];
            for (int k = 0; k < 4; k++) nv4[k] = (v0[k] * d1 - v1[k] * d0) * factor;
            if (in0) std::memcpy(v1, nv4, sizeof nv4); else std::memcpy(v0, nv4, sizeof nv4);
        } else continue;
        if (nv == 0 || !vec4_equals(verts[nv - 1], v0)) { std::memcpy(verts[nv++], v0, 16); }
        std::memcpy(verts[nv++], v1, 16);
    }
    if (nv < 3) return 0;
    if (vec4_equals(verts[0], verts[nv - 1])) nv--;
    int nt = 0;
    for (int i = 1; i < nv - 1 && nt < 3; i++) {
        for (int k = 0; k < 4; k++) { out[nt][k][0] = verts[0][k]; out[nt][k][1] = verts[i][k]; out[nt][k][2] = verts[i + 1][k]; }
        nt++;
    }
    return nt;
}

void orc_render_with(const OrcScene* sc, const double* frames, const float* view16, const float* proj16,
                     uint8_t* rgb, float* depth, int32_t* mask) {
    const int W = ORC_W, H = ORC_H;
    Raster R;
    R.z.assign(W * H, -1e30f); R.seg.assign(W * H, -1); R.rgb.assign(W * H * 3, 255);
    Mat4 view, proj;
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { view[i][j] = view16[i + 4 * j]; proj[i][j] = proj16[i + 4 * j]; }
    /* near/far recovered from the projection (TinyRendererVisualShapeConverter::render) */
    float nearP = proj16[14] / (proj16[10] - 1);
    float farP = proj16[14] / (proj16[10] + 1);
    /* camera position for back-face culling: viewMatrix.invert() translation column, as doubles */
    Mat4 vit; mat4_invert_transpose(view, vit);            /* invert() = invert_transpose().transpose() */
    double P[3] = {(double)vit[3][0], (double)vit[3][1], (double)vit[3][2]};
    /* light: normalize(-50,30,100) in double, then float */
    double ld[3] = {-50.0, 30.0, 100.0};
    double ll = std::sqrt(ld[0] * ld[0] + ld[1] * ld[1] + ld[2] * ld[2]);
    ShaderCtx S;
    for (int i = 0; i < 3; i++) S.light[i] = (float)(ld[i] / ll);
    S.nearPlane = proj[2][3] / (proj[2][2] - 1);
    S.farPlane = proj[2][3] / (proj[2][2] + 1);
    int nshapes = (int)sc->shapes.size();
    for (int s = 0; s < nshapes; s++) {
        const OrcShape& sh = sc->shapes[s];
        if (sh.body >= 2 && sh.body - 2 >= sc->n_objects) continue;
        if (sh.rgba[3] == 0) continue;
        const double* F = frames + 12 * s;
        Mat4 model;
        for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) model[i][j] = (float)F[3 * i + j]; model[i][3] = (float)F[9 + i]; }
        model[3][0] = model[3][1] = model[3][2] = 0.f; model[3][3] = 1.f;
        Mat4 mv, pmv, invT;
        mat4_mul(view, model, mv);
        mat4_mul(proj, mv, pmv);
        mat4_invert_transpose(model, invT);
        S.tex = sh.tex >= 0 ? &sc->textures[sh.tex] : nullptr;
        for (int k = 0; k < 4; k++) S.rgba[k] = sh.rgba[k];
        int objectAndLink = sh.body + ((sh.link + 1) << 24);
        for (int t = 0; t < sh.tri_count; t++) {
            const float* tv = sc->tris + (size_t)(sh.tri_start + t) * 24;
            for (int j = 0; j < 3; j++) {
                const float* v = tv + 8 * j;
                float p4[4] = {v[0] * 1.f, v[1] * 1.f, v[2] * 1.f, 1.f};        /* localScaling = 1 */
                float n4[4] = {v[3], v[4], v[5], 0.f};
                float o[4];
                mat4_vec(invT, n4, o); S.nrm[0][j] = o[0]; S.nrm[1][j] = o[1]; S.nrm[2][j] = o[2];
                mat4_vec(pmv, p4, o);  for (int k = 0; k < 4; k++) S.tri[k][j] = o[k];
                mat4_vec(model, p4, o); for (int k = 0; k < 3; k++) S.world[k][j] = o[k];
                S.uv[0][j] = v[6]; S.uv[1][j] = v[7];
            }
            /* back-face culling in double on the float world coordinates */
            double v0[3], e1[3], e2[3];
            for (int k = 0; k < 3; k++) { v0[k] = S.world[k][0]; e1[k] = (double)S.world[k][1] - v0[k]; e2[k] = (double)S.world[k][2] - v0[k]; }
            double N[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
            double dd = (v0[0] - P[0]) * N[0] + (v0[1] - P[1]) * N[1] + (v0[2] - P[2]) * N[2];
            if (dd >= 0) continue;
            float clipped[3][4][3];
            int nc = clip_near(S.tri, clipped);
            if (nc < 0) raster_triangle(R, S, objectAndLink);
            else for (int c = 0; c < nc; c++) raster_triangle_clipped(R, S, clipped[c], objectAndLink);
        }
    }
    /* copyCameraImageData: vertical flip, depth conversion, mask low 24 bits */
    for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
        int src = x + (H - 1 - y) * W, dst = x + y * W;
        if (rgb) { rgb[3 * dst] = R.rgb[3 * src]; rgb[3 * dst + 1] = R.rgb[3 * src + 1]; rgb[3 * dst + 2] = R.rgb[3 * src + 2]; }
        if (depth) {
            float z_c = -R.z[src];
            depth[dst] = (float)((double)(farP * (nearP + z_c)) / (2. * (double)farP * (double)nearP + (double)(farP * z_c) - (double)(nearP * z_c)));
        }
        if (mask) { int m = R.seg[src]; if (m >= 0) m &= ((1 << 24) - 1); mask[dst] = m; }
    }
}

void orc_render_eye(const OrcEnv* e, uint8_t* rgb, float* depth, int32_t* mask) {
    float view[16], proj[16];
    orc_eye_camera(view, proj);
    int S = (int)e->sc->shapes.size();
    std::vector<double> frames(12 * S);
    for (int s = 0; s < S; s++) {
        OrcFrame f = orc_shape_frame(e, s);
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) frames[12 * s + 3 * a + b] = f.R.m[a][b];
        frames[12 * s + 9] = f.p.x; frames[12 * s + 10] = f.p.y; frames[12 * s + 11] = f.p.z;
    }
    orc_render_with(e->sc, frames.data(), view, proj, rgb, depth, mask);
}
