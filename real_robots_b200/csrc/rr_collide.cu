// rr_collide.cu — convex-hull narrowphase + persistent 4-point contact manifolds, sm_100a.  Compile with -fmad=false.
//
// Replaces the collision-detection half of pybullet's stepSimulation as called from the reference at
// real_robots/envs/env.py:340 (btCollisionWorld::performDiscreteCollisionDetection: btConvexConvexAlgorithm =
// GJK closest points on the hull cores, EPA when the cores overlap, hull margin 0.001, btPersistentManifold with
// contact breaking threshold 0.02 — SURVEY App. B.4), and control_objects_limits (env.py:257-264) which must run
// before it.
//
// Design: ONE WARP PER ENV, the same mapping as the dynamics kernel.  Lanes test the <= 92 candidate shape pairs
// against each other's world AABBs in parallel (ballot), then the warp walks the overlapping pairs in a fixed
// order.  GJK / EPA control flow is executed redundantly by all 32 lanes on identical values (no divergence, no
// broadcasts); only the support mapping is cooperative: lanes stride the hull's vertices and an argmax shuffle
// picks the FIRST maximum, exactly as a sequential scan would.  The EPA polytope lives in shared memory.
// Manifolds are double-buffered in global memory ([env][slot][...], read from one buffer, written to the other).
#include "rr_ctx.h"
#include <cstdlib>
#include <cstring>
#include <cfloat>

#define CWPB 4
#define FULL 0xffffffffu
#define EPA_MAXV 40
#define EPA_MAXF 80

struct CollideConst {
    int n_pairs;
    int pair_a[RR_MAX_PAIRS], pair_b[RR_MAX_PAIRS];
    int hull_start[RR_NSHAPES], hull_count[RR_NSHAPES];
    int shape_body[RR_NSHAPES], shape_link[RR_NSHAPES];
    double aabb_c[RR_NSHAPES][3], aabb_h[RR_NSHAPES][3];      // local AABB centre / half extents (without margin)
    double base_pos[3], table_R[9], table_pos[3], table_upper_pos[3];
    double obj_init[RR_MAX_OBJECTS][7];
};
__constant__ CollideConst c_col;

#ifndef CT_INLINE
#define CT_INLINE __forceinline__
#endif
#ifndef HS_UNROLL
#define HS_UNROLL 2
#endif
#define RR_DO_PRAGMA(x) _Pragma(#x)
#define RR_UNROLL(n) RR_DO_PRAGMA(unroll n)
#define HS_UNROLL_PRAGMA RR_UNROLL(HS_UNROLL)
#ifndef EPA_INLINE
#define EPA_INLINE __noinline__
#endif
#ifndef HS_INLINE
#define HS_INLINE __forceinline__
#endif
struct V3 { double x, y, z; };
__device__ __forceinline__ V3 mk(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return mk(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ V3 mulR(const double* R, V3 v) { return mk(R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z); }
__device__ __forceinline__ V3 mulRt(const double* R, V3 v) { return mk(R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z); }

struct SupPt { V3 w, a, b; };
struct Hull { const double* v; int n; const double* R; V3 p; };

struct ColMem {
    double R[RR_NSHAPES][9];
    double p[RR_NSHAPES][3];
    double amin[RR_NSHAPES][3], amax[RR_NSHAPES][3];
    // EPA polytope
    double Vw[EPA_MAXV][3], Va[EPA_MAXV][3], Vb[EPA_MAXV][3];
    double Fn[EPA_MAXF][3], Fd[EPA_MAXF];
    int Fv[EPA_MAXF][3]; int Falive[EPA_MAXF];
    int eh[2 * EPA_MAXF][2];      // horizon edges
    int epa_nf, epa_good;         // counters published by lane 0 after a polytope update
    // manifold staging
    double pt[4][RR_PT_STRIDE];
    // GJK simplex (w, a, b per vertex): indexed dynamically, so it lives here rather than in per-thread local memory
    double gs[4][9];
};

// cooperative support vertex: first maximum of v . (R^T d) in vertex order
__device__ HS_INLINE V3 hull_support(const Hull& h, V3 d, int lane) {
    V3 dl = mulRt(h.R, d);
    double best = -DBL_MAX; int idx = 0x7fffffff;
HS_UNROLL_PRAGMA
    for (int i = lane; i < h.n; i += 32) {
        double s = h.v[3 * i] * dl.x + h.v[3 * i + 1] * dl.y + h.v[3 * i + 2] * dl.z;
        if (s > best) { best = s; idx = i; }
    }
    {   // warp arg-max with "lowest index wins ties": the value through an order-preserving 64-bit key and two 32-bit
        // redux steps, then the lowest index among the lanes that hold the maximum (every lane's idx is already the
        // first maximum of its own strided subset)
        const long long bits = __double_as_longlong(best + 0.0);                       // -0 -> +0: equal values, equal keys
        const unsigned long long key = bits < 0 ? ~(unsigned long long)bits : ((unsigned long long)bits | 0x8000000000000000ull);
        const unsigned int hi = (unsigned int)(key >> 32), lo = (unsigned int)key;
        const unsigned int mhi = __reduce_max_sync(FULL, hi);
        const unsigned int mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
        const bool win = hi == mhi && lo == mlo;
        idx = (int)__reduce_min_sync(FULL, win ? (unsigned int)idx : 0xffffffffu);
    }
    V3 vl = mk(h.v[3 * idx], h.v[3 * idx + 1], h.v[3 * idx + 2]);
    return mulR(h.R, vl) + h.p;
}
__device__ __forceinline__ SupPt support(const Hull& A, const Hull& B, V3 d, int lane) {
    SupPt s; s.a = hull_support(A, d, lane); s.b = hull_support(B, -d, lane); s.w = s.a - s.b; return s;
}

__device__ __forceinline__ void closest_segment(V3 a, V3 b, double* lam, int* mask) {
    V3 ab = b - a; double t = -dot(a, ab), den = dot(ab, ab);
    if (t <= 0 || den <= 0) { lam[0] = 1; lam[1] = 0; *mask = 1; return; }
    if (t >= den) { lam[0] = 0; lam[1] = 1; *mask = 2; return; }
    t /= den; lam[0] = 1 - t; lam[1] = t; *mask = 3;
}
__device__ CT_INLINE void closest_triangle(V3 a, V3 b, V3 c, double* lam, int* mask) {
    V3 ab = b - a, ac = c - a, ap = -a;
    double d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0 && d2 <= 0) { lam[0] = 1; lam[1] = lam[2] = 0; *mask = 1; return; }
    V3 bp = -b; double d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0 && d4 <= d3) { lam[1] = 1; lam[0] = lam[2] = 0; *mask = 2; return; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0 && d1 >= 0 && d3 <= 0) { double v = d1 / (d1 - d3); lam[0] = 1 - v; lam[1] = v; lam[2] = 0; *mask = 3; return; }
    V3 cp = -c; double d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0 && d5 <= d6) { lam[2] = 1; lam[0] = lam[1] = 0; *mask = 4; return; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0 && d2 >= 0 && d6 <= 0) { double w = d2 / (d2 - d6); lam[0] = 1 - w; lam[1] = 0; lam[2] = w; *mask = 5; return; }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0 && (d4 - d3) >= 0 && (d5 - d6) >= 0) { double w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); lam[0] = 0; lam[1] = 1 - w; lam[2] = w; *mask = 6; return; }
    double denom = 1.0 / (va + vb + vc);
    double v = vb * denom, w = vc * denom;
    lam[0] = 1 - v - w; lam[1] = v; lam[2] = w; *mask = 7;
}
__device__ __forceinline__ void lam_set(double* lam, int idx, double v) {       // lam[idx] = v without dynamic register indexing
#pragma unroll
    for (int t = 0; t < 4; t++) if (t == idx) lam[t] = v;
}
__device__ bool closest_tetra(const double (*P)[9], double* lam, int* mask) {     // P: the simplex rows in shared memory (w first)
    const int F[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
    const int O[4] = {3, 1, 2, 0};
    double best = DBL_MAX; bool outside_any = false;
    for (int f = 0; f < 4; f++) {
        const int i0 = F[f][0], i1 = F[f][1], i2 = F[f][2], io = O[f];
        V3 a = mk(P[i0][0], P[i0][1], P[i0][2]), b = mk(P[i1][0], P[i1][1], P[i1][2]), c = mk(P[i2][0], P[i2][1], P[i2][2]), d = mk(P[io][0], P[io][1], P[io][2]);
        V3 n = cross(b - a, c - a);
        double sp = dot(-a, n), sd = dot(d - a, n);
        bool outside = (sp * sd < 0) || (sd == 0);
        if (!outside) continue;
        outside_any = true;
        double l3[3]; int m3;
        closest_triangle(a, b, c, l3, &m3);
        V3 q = a * l3[0] + b * l3[1] + c * l3[2];
        double d2 = dot(q, q);
        if (d2 < best) {
            best = d2;
#pragma unroll
            for (int k = 0; k < 4; k++) lam[k] = 0;
            lam_set(lam, i0, l3[0]); lam_set(lam, i1, l3[1]); lam_set(lam, i2, l3[2]);
            *mask = ((m3 & 1) ? 1 << i0 : 0) | ((m3 & 2) ? 1 << i1 : 0) | ((m3 & 4) ? 1 << i2 : 0);
        }
    }
    return outside_any;
}

struct GjkResult { int status; double dist; V3 n, pa, pb; int ns; };      // the simplex itself stays in ColMem::gs

__device__ __forceinline__ V3 gs_w(const ColMem& m, int k) { return mk(m.gs[k][0], m.gs[k][1], m.gs[k][2]); }
__device__ __forceinline__ V3 gs_a(const ColMem& m, int k) { return mk(m.gs[k][3], m.gs[k][4], m.gs[k][5]); }
__device__ __forceinline__ V3 gs_b(const ColMem& m, int k) { return mk(m.gs[k][6], m.gs[k][7], m.gs[k][8]); }
// the simplex is written by lane 0 only (every lane holds the same values) and read by all lanes after a __syncwarp
__device__ __forceinline__ void gs_store(ColMem& m, int k, const SupPt& s, int lane) {
    __syncwarp();
    if (lane == 0) {
        m.gs[k][0] = s.w.x; m.gs[k][1] = s.w.y; m.gs[k][2] = s.w.z; m.gs[k][3] = s.a.x; m.gs[k][4] = s.a.y; m.gs[k][5] = s.a.z;
        m.gs[k][6] = s.b.x; m.gs[k][7] = s.b.y; m.gs[k][8] = s.b.z;
    }
    __syncwarp();
}

__device__ void gjk(const Hull& A, const Hull& B, ColMem& m, GjkResult& r, int lane) {
    r.status = 0; r.ns = 0; r.dist = 0;
    V3 v = A.p - B.p;
    if (dot(v, v) < 1e-12) v = mk(1, 0, 0);
    int ns = 0;
    double lam[4] = {0, 0, 0, 0};                 // barycentric weights of the simplex vertices: registers, static indexing only
    for (int it = 0; it < 64; it++) {
        SupPt w = support(A, B, -v, lane);
        double vv = dot(v, v);
        if (ns > 0 && vv - dot(v, w.w) <= 1e-10 * vv) break;
        bool dup = false;
        for (int k = 0; k < ns; k++) { V3 d = gs_w(m, k) - w.w; if (dot(d, d) < 1e-24) dup = true; }
        if (dup) break;
        gs_store(m, ns, w, lane); ns++;
        int mask = 0;
        if (ns == 1) { lam[0] = 1; mask = 1; }
        else if (ns == 2) closest_segment(gs_w(m, 0), gs_w(m, 1), lam, &mask);
        else if (ns == 3) closest_triangle(gs_w(m, 0), gs_w(m, 1), gs_w(m, 2), lam, &mask);
        else if (!closest_tetra(m.gs, lam, &mask)) { r.status = 1; r.ns = 4; return; }
        // keep the vertices of the closest feature, in order
        int nt = 0;
        double nl[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; k++) if (k < ns && (mask & (1 << k))) {
            if (nt != k) {                          // warp-uniform
                __syncwarp();
                if (lane < 9) m.gs[nt][lane] = m.gs[k][lane];
                __syncwarp();
            }
            lam_set(nl, nt, lam[k]);
            nt++;
        }
        ns = nt;
        V3 nv = mk(0, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; k++) { lam[k] = nl[k]; if (k < ns) nv = nv + gs_w(m, k) * lam[k]; }
        v = nv;
        if (dot(v, v) < 1e-20) { r.status = 1; r.ns = ns; return; }
    }
    V3 pa = mk(0, 0, 0), pb = mk(0, 0, 0);
#pragma unroll
    for (int k = 0; k < 4; k++) if (k < ns) { pa = pa + gs_a(m, k) * lam[k]; pb = pb + gs_b(m, k) * lam[k]; }
    double d = sqrt(dot(v, v));
    r.dist = d; r.n = v * (1.0 / d); r.pa = pa; r.pb = pb; r.ns = ns;
}

__device__ __forceinline__ double det3(V3 a, V3 b, V3 c) { return dot(a, cross(b, c)); }

// EncloseOrigin (btGjkEpa2), unrolled into three non-recursive levels
__device__ bool enclose4(const SupPt* S) { return fabs(det3(S[0].w - S[3].w, S[1].w - S[3].w, S[2].w - S[3].w)) > 0; }
__device__ bool enclose3(const Hull& A, const Hull& B, SupPt* S, int lane) {
    V3 n = cross(S[1].w - S[0].w, S[2].w - S[0].w);
    if (dot(n, n) > 0) {
        S[3] = support(A, B, n, lane); if (enclose4(S)) return true;
        S[3] = support(A, B, -n, lane); if (enclose4(S)) return true;
    }
    return false;
}
__device__ bool enclose2(const Hull& A, const Hull& B, SupPt* S, int lane) {
    V3 d = S[1].w - S[0].w;
    for (int i = 0; i < 3; i++) {
        V3 axis = mk(i == 0, i == 1, i == 2);
        V3 p = cross(d, axis);
        if (dot(p, p) > 0) {
            S[2] = support(A, B, p, lane); if (enclose3(A, B, S, lane)) return true;
            S[2] = support(A, B, -p, lane); if (enclose3(A, B, S, lane)) return true;
        }
    }
    return false;
}
__device__ bool enclose1(const Hull& A, const Hull& B, SupPt* S, int lane) {
    for (int i = 0; i < 3; i++) {
        V3 axis = mk(i == 0, i == 1, i == 2);
        S[1] = support(A, B, axis, lane); if (enclose2(A, B, S, lane)) return true;
        S[1] = support(A, B, -axis, lane); if (enclose2(A, B, S, lane)) return true;
    }
    return false;
}

__device__ __forceinline__ V3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(double* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

__device__ bool epa_add_face(ColMem& m, int& nf, int a, int b, int c) {
    if (nf >= EPA_MAXF) return false;
    V3 wa = ld3(m.Vw[a]), wb = ld3(m.Vw[b]), wc = ld3(m.Vw[c]);
    V3 nn = cross(wb - wa, wc - wa);
    double l = sqrt(dot(nn, nn));
    if (l <= 1e-30) return false;
    V3 n = nn * (1.0 / l); double d = dot(n, wa);
    int v1 = b, v2 = c;
    if (d < 0) { n = -n; d = -d; v1 = c; v2 = b; }
    m.Fv[nf][0] = a; m.Fv[nf][1] = v1; m.Fv[nf][2] = v2; m.Falive[nf] = 1;
    st3(m.Fn[nf], n); m.Fd[nf] = d;
    nf++;
    return true;
}

__device__ EPA_INLINE bool epa(const Hull& A, const Hull& B, ColMem& m, SupPt* S, int ns, double& depth, V3& n, V3& pa, V3& pb, int lane) {
    bool ok = false;
    if (ns == 1) ok = enclose1(A, B, S, lane);
    else if (ns == 2) ok = enclose2(A, B, S, lane);
    else if (ns == 3) ok = enclose3(A, B, S, lane);
    else if (ns == 4) ok = enclose4(S);
    if (!ok) return false;
    if (det3(S[0].w - S[3].w, S[1].w - S[3].w, S[2].w - S[3].w) < 0) { SupPt t = S[0]; S[0] = S[1]; S[1] = t; }
    int nv = 0, nf = 0;
    for (int k = 0; k < 4; k++) { st3(m.Vw[nv], S[k].w); st3(m.Va[nv], S[k].a); st3(m.Vb[nv], S[k].b); nv++; }
    __syncwarp();
    if (!epa_add_face(m, nf, 0, 1, 2) || !epa_add_face(m, nf, 1, 0, 3) || !epa_add_face(m, nf, 2, 1, 3) || !epa_add_face(m, nf, 0, 2, 3)) return false;
    __syncwarp();
    int best = -1;
    for (int it = 0; it < 128; it++) {
        best = -1; double bd = DBL_MAX;
        for (int f = 0; f < nf; f++) if (m.Falive[f] && m.Fd[f] < bd) { bd = m.Fd[f]; best = f; }
        if (best < 0) return false;
        V3 bn = ld3(m.Fn[best]);
        SupPt w = support(A, B, bn, lane);
        double wd = dot(w.w, bn) - m.Fd[best];
        if (wd <= 1e-10 || nv >= EPA_MAXV) break;
        // Horizon and re-triangulation mutate the polytope while reading it: lane 0 does it alone (the other lanes would only
        // repeat the same stores), everybody re-reads the counters after the barrier.
        __syncwarp();
        if (lane == 0) {
            int ne = 0; bool ok2 = true;
            for (int f = 0; f < nf && ok2; f++) {
                if (!m.Falive[f]) continue;
                if (ne + 3 > 2 * EPA_MAXF) { ok2 = false; break; }      // horizon buffer full: stop refining (never seen; memory safety)
                if (dot(ld3(m.Fn[f]), w.w - ld3(m.Vw[m.Fv[f][0]])) > 0) {
                    m.Falive[f] = 0;
                    for (int e = 0; e < 3; e++) {
                        int a = m.Fv[f][e], b = m.Fv[f][(e + 1) % 3];
                        bool cancelled = false;
                        for (int k = 0; k < ne; k++) if (m.eh[k][0] == b && m.eh[k][1] == a) { m.eh[k][0] = m.eh[ne - 1][0]; m.eh[k][1] = m.eh[ne - 1][1]; ne--; cancelled = true; break; }
                        if (!cancelled) { m.eh[ne][0] = a; m.eh[ne][1] = b; ne++; }
                    }
                }
            }
            const int wi = nv; st3(m.Vw[nv], w.w); st3(m.Va[nv], w.a); st3(m.Vb[nv], w.b);
            int nf2 = nf;
            for (int k = 0; k < ne && ok2; k++) if (!epa_add_face(m, nf2, m.eh[k][0], m.eh[k][1], wi)) { ok2 = false; break; }
            m.epa_nf = nf2; m.epa_good = ok2 ? 1 : 0;
        }
        __syncwarp();
        nv++;
        nf = m.epa_nf;
        const bool good = m.epa_good != 0;
        if (!good) break;
    }
    depth = m.Fd[best]; n = ld3(m.Fn[best]);
    int i0 = m.Fv[best][0], i1 = m.Fv[best][1], i2 = m.Fv[best][2];
    V3 pr = n * depth;
    double l[3]; int mm;
    closest_triangle(ld3(m.Vw[i0]) - pr, ld3(m.Vw[i1]) - pr, ld3(m.Vw[i2]) - pr, l, &mm);
    pa = ld3(m.Va[i0]) * l[0] + ld3(m.Va[i1]) * l[1] + ld3(m.Va[i2]) * l[2];
    pb = ld3(m.Vb[i0]) * l[0] + ld3(m.Vb[i1]) * l[1] + ld3(m.Vb[i2]) * l[2];
    __syncwarp();
    return true;
}

__device__ bool closest_points(const Hull& A, const Hull& B, ColMem& m, double& dist, V3& n, V3& pa, V3& pb, int lane) {
    GjkResult g; gjk(A, B, m, g, lane);
    if (g.status == 0) {
        if (!(g.dist > 0)) return false;
        dist = g.dist; n = g.n; pa = g.pa; pb = g.pb; return true;
    }
    double depth; V3 en, ea, eb;
    SupPt S4[4];
    for (int k = 0; k < 4; k++) { S4[k].w = gs_w(m, k); S4[k].a = gs_a(m, k); S4[k].b = gs_b(m, k); }
    if (!epa(A, B, m, S4, g.ns, depth, en, ea, eb, lane)) return false;
    dist = -depth; n = -en; pa = ea; pb = eb;
    return true;
}

// point layout (RR_PT_STRIDE doubles): localA 0..2, localB 3..5, normalB 6..8, distance 9, applied impulse 10, worldA 11..13, worldB 14..16
__device__ int cache_entry(const ColMem& m, int n, V3 localA, double thr) {
    double shortest = thr * thr; int nearest = -1;
    for (int i = 0; i < n; i++) {
        V3 d = ld3(&m.pt[i][0]) - localA; double dd = dot(d, d);
        if (dd < shortest) { shortest = dd; nearest = i; }
    }
    return nearest;
}
__device__ int sort_cached_points(const ColMem& m, V3 pn, double pdist) {
    int maxPenIdx = -1; double maxPen = pdist;
    for (int i = 0; i < 4; i++) if (m.pt[i][9] < maxPen) { maxPenIdx = i; maxPen = m.pt[i][9]; }
    double res[4] = {0, 0, 0, 0};
    V3 p0 = ld3(&m.pt[0][0]), p1 = ld3(&m.pt[1][0]), p2 = ld3(&m.pt[2][0]), p3 = ld3(&m.pt[3][0]);
    if (maxPenIdx != 0) { V3 c = cross(pn - p1, p3 - p2); res[0] = dot(c, c); }
    if (maxPenIdx != 1) { V3 c = cross(pn - p0, p3 - p2); res[1] = dot(c, c); }
    if (maxPenIdx != 2) { V3 c = cross(pn - p0, p3 - p1); res[2] = dot(c, c); }
    if (maxPenIdx != 3) { V3 c = cross(pn - p0, p2 - p1); res[3] = dot(c, c); }
    int mi = -1; double mv = -1e30;
    for (int i = 0; i < 4; i++) if (fabs(res[i]) > mv) { mv = fabs(res[i]); mi = i; }
    return mi;
}

#ifndef COLLIDE_MIN_BLOCKS
#define COLLIDE_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(CWPB * 32, COLLIDE_MIN_BLOCKS)
k_collide(DevState st, const double* __restrict__ hull_verts, const int* __restrict__ man_hdr_in, const double* __restrict__ man_pts_in,
          int* __restrict__ man_hdr_out, double* __restrict__ man_pts_out, rr_config cf, int n_objects, int mode, int heavy_cap) {
    __shared__ ColMem smem[CWPB];
    ColMem& m = smem[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    int e = blockIdx.x * CWPB + (threadIdx.x >> 5);
    if (e >= st.N) return;
    // mode 0: every env.  The look-ahead launches of rrk_step come in two parts: mode 1 skips the envs of the solver's heavy kernel
    // (still running), mode 2 follows that kernel and takes exactly those (warp w: the w-th of them)
    if (mode == 1 && solve_is_heavy(st, e, heavy_cap)) return;
    if (mode == 2) { if (e >= solve_n_heavy(st, heavy_cap)) return; e = solve_pick(st, e); }
    const int N = st.N;
    const double thr = cf.contact_breaking;

    // control_objects_limits (env.py:257-264) runs before the step's collision detection: an object out of bounds collides at its
    // initial pose.  This kernel only READS the state (it may run ahead of the step, beside the rasteriser of the previous one);
    // k_step_dyn applies the same test and writes the teleported state.
    // ---- world frames and AABBs of the 22 shapes (lane = shape)
    if (lane < RR_NSHAPES) {
        const int s = lane, body = c_col.shape_body[s], link = c_col.shape_link[s];
        double R[9], p[3];
        if (body == 0) {
            if (link < 0) { for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.0 : 0.0; for (int k = 0; k < 3; k++) p[k] = c_col.base_pos[k]; }
            else { for (int k = 0; k < 9; k++) R[k] = st.link_frames[((size_t)link * 12 + k) * N + e]; for (int k = 0; k < 3; k++) p[k] = st.link_frames[((size_t)link * 12 + 9 + k) * N + e]; }
        } else if (body == 1) {
            for (int k = 0; k < 9; k++) R[k] = c_col.table_R[k];
            for (int k = 0; k < 3; k++) p[k] = link < 0 ? c_col.table_pos[k] : c_col.table_upper_pos[k];
        } else {
            const int o = body - 2;
            double q[4];
            for (int k = 0; k < 4; k++) q[k] = st.obj[((size_t)o * 13 + 3 + k) * N + e];
            for (int k = 0; k < 3; k++) p[k] = st.obj[((size_t)o * 13 + k) * N + e];
            if (p[2] < c_col.table_pos[2] || (p[0] > 0.11 && p[2] < 0.29)) {
                for (int k = 0; k < 3; k++) p[k] = c_col.obj_init[o][k];
                for (int k = 0; k < 4; k++) q[k] = c_col.obj_init[o][3 + k];
            }
            double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
            double sc = 2.0 / d;
            double xs = q[0] * sc, ys = q[1] * sc, zs = q[2] * sc;
            double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
            double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
            double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
            R[0] = 1.0 - (yy + zz); R[1] = xy - wz; R[2] = xz + wy;
            R[3] = xy + wz; R[4] = 1.0 - (xx + zz); R[5] = yz - wx;
            R[6] = xz - wy; R[7] = yz + wx; R[8] = 1.0 - (xx + yy);
        }
        for (int k = 0; k < 9; k++) m.R[s][k] = R[k];
        for (int k = 0; k < 3; k++) m.p[s][k] = p[k];
        V3 c = ld3(c_col.aabb_c[s]);
        V3 h = mk(c_col.aabb_h[s][0] + cf.margin, c_col.aabb_h[s][1] + cf.margin, c_col.aabb_h[s][2] + cf.margin);
        V3 wc = mulR(R, c) + mk(p[0], p[1], p[2]);
        double ext[3];
        for (int i = 0; i < 3; i++) ext[i] = fabs(R[3 * i]) * h.x + fabs(R[3 * i + 1]) * h.y + fabs(R[3 * i + 2]) * h.z + thr;
        m.amin[s][0] = wc.x - ext[0]; m.amin[s][1] = wc.y - ext[1]; m.amin[s][2] = wc.z - ext[2];
        m.amax[s][0] = wc.x + ext[0]; m.amax[s][1] = wc.y + ext[1]; m.amax[s][2] = wc.z + ext[2];
    }
    __syncwarp();
    // ---- broadphase: lanes test pairs lane, lane+32, lane+64
    unsigned ov[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        int pi = lane + 32 * r; bool o = false;
        if (pi < c_col.n_pairs) {
            int a = c_col.pair_a[pi], b = c_col.pair_b[pi];
            o = m.amin[a][0] <= m.amax[b][0] && m.amax[a][0] >= m.amin[b][0] && m.amin[a][1] <= m.amax[b][1] && m.amax[a][1] >= m.amin[b][1] &&
                m.amin[a][2] <= m.amax[b][2] && m.amax[a][2] >= m.amin[b][2];
        }
        ov[r] = __ballot_sync(FULL, o);
    }
    // ---- old manifold headers: lane = slot
    const int* hin = man_hdr_in + (size_t)e * RR_MAX_MANIFOLDS * 2;
    const double* pin = man_pts_in + (size_t)e * RR_MAX_MANIFOLDS * 4 * RR_PT_STRIDE;
    int* hout = man_hdr_out + (size_t)e * RR_MAX_MANIFOLDS * 2;
    double* pout = man_pts_out + (size_t)e * RR_MAX_MANIFOLDS * 4 * RR_PT_STRIDE;
    int old_pair = -1, old_n = 0;
    if (lane < RR_MAX_MANIFOLDS) { old_pair = hin[2 * lane]; old_n = hin[2 * lane + 1]; }
    int nout = 0, dropped = 0;
    for (int r = 0; r < 3; r++) {
        unsigned bits = ov[r];
        while (bits) {
            const int bpos = __ffs(bits) - 1; bits &= bits - 1;
            const int pi = bpos + 32 * r;
            const int a = c_col.pair_a[pi], b = c_col.pair_b[pi];
            // fetch the cached manifold of this pair, if any
            unsigned hit = __ballot_sync(FULL, old_pair == pi);
            int n = 0;
            if (hit) {
                int slot = __ffs(hit) - 1;
                n = __shfl_sync(FULL, old_n, slot);
                for (int k = lane; k < n * RR_PT_STRIDE; k += 32) m.pt[k / RR_PT_STRIDE][k % RR_PT_STRIDE] = pin[(size_t)slot * 4 * RR_PT_STRIDE + k];
            }
            __syncwarp();
            Hull A{hull_verts + 3 * (size_t)c_col.hull_start[a], c_col.hull_count[a], m.R[a], ld3(m.p[a])};
            Hull B{hull_verts + 3 * (size_t)c_col.hull_start[b], c_col.hull_count[b], m.R[b], ld3(m.p[b])};
            double dcore; V3 nrm, pa, pb;
            if (closest_points(A, B, m, dcore, nrm, pa, pb, lane)) {
                double distance = dcore - 2 * cf.margin;
                V3 onB = pb + nrm * cf.margin;
                if (distance <= thr) {
                    V3 onA = onB + nrm * distance;
                    V3 localA = mulRt(m.R[a], onA - ld3(m.p[a]));
                    V3 localB = mulRt(m.R[b], onB - ld3(m.p[b]));
                    int idx = cache_entry(m, n, localA, thr);
                    double applied = 0.0;
                    int slot;
                    if (idx >= 0) { applied = m.pt[idx][10]; slot = idx; }
                    else if (n < 4) slot = n++;
                    else slot = sort_cached_points(m, localA, distance);
                    __syncwarp();
                    double* P = m.pt[slot];
                    st3(P, localA); st3(P + 3, localB); st3(P + 6, nrm); P[9] = distance; P[10] = applied; st3(P + 11, onA); st3(P + 14, onB);
                    __syncwarp();
                }
            }
            // refreshContactPoints
            for (int i = n - 1; i >= 0; i--) {
                double* P = m.pt[i];
                V3 wa = mulR(m.R[a], ld3(P)) + ld3(m.p[a]);
                V3 wb = mulR(m.R[b], ld3(P + 3)) + ld3(m.p[b]);
                double dist = dot(wa - wb, ld3(P + 6));
                __syncwarp();
                st3(P + 11, wa); st3(P + 14, wb); P[9] = dist;
                __syncwarp();
            }
            for (int i = n - 1; i >= 0; i--) {
                double* P = m.pt[i];
                bool drop = false;
                if (!(P[9] <= thr)) drop = true;
                else {
                    V3 proj = ld3(P + 11) - ld3(P + 6) * P[9];
                    V3 d = ld3(P + 14) - proj;
                    if (dot(d, d) > thr * thr) drop = true;
                }
                if (drop) {
                    __syncwarp();
                    if (i != n - 1) { double t = (lane < RR_PT_STRIDE) ? m.pt[n - 1][lane] : 0.0; __syncwarp(); if (lane < RR_PT_STRIDE) m.pt[i][lane] = t; }
                    n--;
                    __syncwarp();
                }
            }
            // A manifold without points carries no state (a later hit on it finds n = 0, the same as no hit), so only non-empty
            // manifolds take a slot: AABB overlaps without contact cannot crowd out the object x table pairs at the end of the
            // pair order.  A non-empty manifold that still finds no slot is flagged (rr_diag flags bit 4).
            if (n > 0) {
                if (nout < RR_MAX_MANIFOLDS) {
                    for (int k = lane; k < n * RR_PT_STRIDE; k += 32) pout[(size_t)nout * 4 * RR_PT_STRIDE + k] = m.pt[k / RR_PT_STRIDE][k % RR_PT_STRIDE];
                    if (lane == 0) { hout[2 * nout] = pi; hout[2 * nout + 1] = n; }
                    nout++;
                } else dropped = 1;
            }
            __syncwarp();
        }
    }
    for (int k = nout + lane; k < RR_MAX_MANIFOLDS; k += 32) { hout[2 * k] = -1; hout[2 * k + 1] = 0; }
    if (lane == 0) st.col_flags[e] = dropped;
}

// ------------------------------------------------------------------------------------------- host side
#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(_e); return -1; } } while (0)

static rr_ctx* g_col_owner[64];
static int bind_collide(rr_ctx* ctx, cudaStream_t s) {
    rr_ctx*& owner = g_col_owner[ctx->device & 63];
    if (owner != ctx) {
        CK(cudaDeviceSynchronize());          // kernels of the previous owner may still be reading the old constants
        CK(cudaMemcpyToSymbolAsync(c_col, ctx->h_collide_const, sizeof(CollideConst), 0, cudaMemcpyHostToDevice, s));
        owner = ctx;
    }
    return 0;
}
void rrk_collide_unbind(rr_ctx* ctx) {
    if (g_col_owner[ctx->device & 63] == ctx) g_col_owner[ctx->device & 63] = nullptr;
    free(ctx->h_collide_const); ctx->h_collide_const = nullptr;
}

int rrk_collide_upload_constants(rr_ctx* ctx) {
    CollideConst C; memset(&C, 0, sizeof C);
    const RobotConst& R = ctx->robot;
    // candidate pairs in a fixed order: robot x objects, robot x table, objects x table, objects x objects
    int shape_of_link[RR_NL + 1]; int table_shape[2] = {-1, -1}; int obj_shape[RR_MAX_OBJECTS] = {-1, -1, -1};
    for (int s = 0; s < RR_NSHAPES; s++) {
        int body = R.shape_body[s], link = R.shape_link[s];
        if (body == 0) shape_of_link[link + 1] = s;
        else if (body == 1) table_shape[link + 1] = s;
        else obj_shape[body - 2] = s;
        C.shape_body[s] = body; C.shape_link[s] = link;
        C.hull_start[s] = ctx->hull_start[s]; C.hull_count[s] = ctx->hull_count[s];
        for (int k = 0; k < 3; k++) { C.aabb_c[s][k] = ctx->aabb_c[s][k]; C.aabb_h[s][k] = ctx->aabb_h[s][k]; }
    }
    int n = 0;
    for (int l = 0; l <= RR_NL; l++) for (int o = 0; o < ctx->n_objects; o++) { C.pair_a[n] = shape_of_link[l]; C.pair_b[n] = obj_shape[o]; n++; }
    for (int l = 1; l <= RR_NL; l++) for (int t = 0; t < 2; t++) { C.pair_a[n] = shape_of_link[l]; C.pair_b[n] = table_shape[t]; n++; }
    for (int o = 0; o < ctx->n_objects; o++) for (int t = 0; t < 2; t++) { C.pair_a[n] = obj_shape[o]; C.pair_b[n] = table_shape[t]; n++; }
    for (int o = 0; o < ctx->n_objects; o++) for (int p = o + 1; p < ctx->n_objects; p++) { C.pair_a[n] = obj_shape[o]; C.pair_b[n] = obj_shape[p]; n++; }
    C.n_pairs = n;
    for (int k = 0; k < 3; k++) { C.base_pos[k] = R.base_pos[k]; C.table_pos[k] = R.table_pos[k]; C.table_upper_pos[k] = R.table_upper_pos[k]; }
    for (int k = 0; k < 9; k++) C.table_R[k] = R.table_R[k];
    for (int o = 0; o < RR_MAX_OBJECTS; o++) for (int k = 0; k < 7; k++) C.obj_init[o][k] = R.obj_init[o][k];
    for (int i = 0; i < n; i++) { ctx->pair_a[i] = C.pair_a[i]; ctx->pair_b[i] = C.pair_b[i]; }
    ctx->n_pairs = n;
    CK(cudaMemcpyToSymbol(c_col, &C, sizeof C));
    if (!ctx->h_collide_const) ctx->h_collide_const = malloc(sizeof(CollideConst));
    memcpy(ctx->h_collide_const, &C, sizeof(CollideConst));
    g_col_owner[ctx->device & 63] = ctx;
    return 0;
}

// flip = false: the look-ahead launch of rrk_step; the new manifolds become current when the next step adopts them
int rrk_collide(rr_ctx* ctx, cudaStream_t s, bool flip, int mode) {
    if (bind_collide(ctx, s)) return -1;
    const int in = ctx->man_parity, out = 1 - in;
    {
        ProfScope ps(ctx, 3, s);
        const int n = mode == 2 ? (ctx->heavy_cap < ctx->N ? ctx->heavy_cap : ctx->N) : ctx->N;
        RR_LAUNCH(k_collide, (n + CWPB - 1) / CWPB, CWPB * 32, 0, s)(ctx->st, ctx->d_hull_verts, ctx->st.man_hdr[in], ctx->st.man_pts[in],
                                                            ctx->st.man_hdr[out], ctx->st.man_pts[out], ctx->cfg, ctx->n_objects, mode, ctx->heavy_cap);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    if (flip) ctx->man_parity = out;
    return 0;
}

#ifdef RR_SIMT_EMU_BUILD
// test hook of the emulated build (tests/simt): closest points of two scene shapes at explicit frames, one warp
__global__ void k_debug_closest(const double* __restrict__ hull_verts, int sa, int sb, const double* fa, const double* fb, double margin, double* out) {
    __shared__ ColMem smem1;
    ColMem& m = smem1;
    const int lane = threadIdx.x & 31;
    if (lane == 0) { for (int k = 0; k < 9; k++) { m.R[sa][k] = fa[k]; m.R[sb][k] = fb[k]; } for (int k = 0; k < 3; k++) { m.p[sa][k] = fa[9 + k]; m.p[sb][k] = fb[9 + k]; } }
    __syncwarp();
    Hull A{hull_verts + 3 * (size_t)c_col.hull_start[sa], c_col.hull_count[sa], m.R[sa], ld3(m.p[sa])};
    Hull B{hull_verts + 3 * (size_t)c_col.hull_start[sb], c_col.hull_count[sb], m.R[sb], ld3(m.p[sb])};
    double d; V3 n, pa, pb;
    const bool ok = closest_points(A, B, m, d, n, pa, pb, lane);
    if (lane == 0) { out[0] = d; out[1] = n.x; out[2] = n.y; out[3] = n.z; out[4] = pa.x; out[5] = pa.y; out[6] = pa.z; out[7] = pb.x; out[8] = pb.y; out[9] = pb.z; out[10] = ok ? 1.0 : 0.0; }
}
extern "C" int rr_debug_closest(rr_ctx* ctx, int sa, const double* fa12, int sb, const double* fb12, double* out11) {
    if (bind_collide(ctx, 0)) return -1;
    RR_LAUNCH(k_debug_closest, 1, 32, 0, 0)(ctx->d_hull_verts, sa, sb, fa12, fb12, ctx->cfg.margin, out11);
    return 0;
}
#endif
