/* TEST INFRASTRUCTURE — CPU oracle: convex-hull narrowphase + persistent contact manifolds.
 *
 * Restates what pybullet does between the reference's apply_action and the constraint solve inside
 * stepSimulation (env.py:340): btCollisionWorld::performDiscreteCollisionDetection with
 * btConvexConvexAlgorithm (GJK closest points on the hull cores, EPA when the cores overlap, hull margin
 * 0.001 each) feeding a 4-point btPersistentManifold per shape pair.  bullet3 is not vendored
 * by the reference (setup.py:24-32, unpinned) — PARITY UNPINNED, restated from the published algorithm; the
 * pair order, GJK tolerances and EPA bookkeeping are this repo's own choices, documented in DESIGN.md.
 */
#include "orc_env.h"
#include <algorithm>
#include <cfloat>
#include <vector>

namespace {

struct Hull { const double* v; int n; M3 R; V3 p; };     /* local vertices + world frame */

struct SupPt { V3 w, a, b; };    /* Minkowski-difference point w = a - b with its witnesses */

/* support vertex of a hull along world direction d: first maximum in vertex order */
V3 hull_support(const Hull& h, V3 d) {
    V3 dl = transpose(h.R) * d;
    int best = 0; double bv = -DBL_MAX;
    for (int i = 0; i < h.n; i++) {
        double s = h.v[3 * i] * dl.x + h.v[3 * i + 1] * dl.y + h.v[3 * i + 2] * dl.z;
        if (s > bv) { bv = s; best = i; }
    }
    V3 vl(h.v[3 * best], h.v[3 * best + 1], h.v[3 * best + 2]);
    return h.R * vl + h.p;
}
SupPt support(const Hull& A, const Hull& B, V3 d) {
    SupPt s; s.a = hull_support(A, d); s.b = hull_support(B, -d); s.w = s.a - s.b; return s;
}

/* closest point to the origin on segment / triangle; returns barycentric weights and the reduced vertex mask */
void closest_segment(V3 a, V3 b, double* lam, int* mask) {
    V3 ab = b - a; double t = -dot(a, ab), den = dot(ab, ab);
    if (t <= 0 || den <= 0) { lam[0] = 1; lam[1] = 0; *mask = 1; return; }
    if (t >= den) { lam[0] = 0; lam[1] = 1; *mask = 2; return; }
    t /= den; lam[0] = 1 - t; lam[1] = t; *mask = 3;
}
void closest_triangle(V3 a, V3 b, V3 c, double* lam, int* mask) {   /* Ericson, Real-Time Collision Detection 5.1.5, p = origin */
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
/* closest point on tetrahedron abcd to the origin; returns false if the origin is inside */
bool closest_tetra(const V3* p, double* lam, int* mask) {
    static const int F[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
    static const int O[4] = {3, 1, 2, 0};
    double best = DBL_MAX; bool outside_any = false;
    for (int f = 0; f < 4; f++) {
        V3 a = p[F[f][0]], b = p[F[f][1]], c = p[F[f][2]], d = p[O[f]];
        V3 n = cross(b - a, c - a);
        double sp = dot(-a, n), sd = dot(d - a, n);
        bool outside = (sp * sd < 0) || (sd == 0);    /* origin on the other side than the 4th vertex (or degenerate) */
        if (!outside) continue;
        outside_any = true;
        double l3[3]; int m3;
        closest_triangle(a, b, c, l3, &m3);
        V3 q = a * l3[0] + b * l3[1] + c * l3[2];
        double d2 = dot(q, q);
        if (d2 < best) {
            best = d2;
            for (int k = 0; k < 4; k++) lam[k] = 0;
            int mk = 0;
            for (int k = 0; k < 3; k++) { lam[F[f][k]] = l3[k]; if (m3 & (1 << k)) mk |= 1 << F[f][k]; }
            *mask = mk;
        }
    }
    return outside_any;
}

struct GjkResult { int status; /* 0 separated, 1 penetrating */ double dist; V3 n /* B -> A */, pa, pb; SupPt simplex[4]; int ns; };

GjkResult gjk(const Hull& A, const Hull& B) {
    GjkResult r; r.status = 0; r.ns = 0; r.dist = 0;
    V3 v = A.p - B.p;
    if (dot(v, v) < 1e-12) v = V3(1, 0, 0);
    SupPt S[4]; int ns = 0;
    double lam[4] = {0, 0, 0, 0};
    for (int it = 0; it < 64; it++) {
        SupPt w = support(A, B, -v);
        double vv = dot(v, v);
        /* convergence: no more progress towards the origin along v */
        if (ns > 0 && vv - dot(v, w.w) <= 1e-10 * vv) break;
        bool dup = false;
        for (int k = 0; k < ns; k++) { V3 d = S[k].w - w.w; if (dot(d, d) < 1e-24) dup = true; }
        if (dup) break;
        S[ns++] = w;
        int mask = 0;
        if (ns == 1) { lam[0] = 1; mask = 1; }
        else if (ns == 2) closest_segment(S[0].w, S[1].w, lam, &mask);
        else if (ns == 3) closest_triangle(S[0].w, S[1].w, S[2].w, lam, &mask);
        else {
            V3 P[4] = {S[0].w, S[1].w, S[2].w, S[3].w};
            if (!closest_tetra(P, lam, &mask)) { r.status = 1; r.ns = 4; for (int k = 0; k < 4; k++) r.simplex[k] = S[k]; return r; }
        }
        /* reduce the simplex to the supporting face */
        SupPt T[4]; double tl[4]; int nt = 0;
        for (int k = 0; k < ns; k++) if (mask & (1 << k)) { T[nt] = S[k]; tl[nt] = lam[k]; nt++; }
        ns = nt;
        V3 nv;
        for (int k = 0; k < ns; k++) { S[k] = T[k]; lam[k] = tl[k]; nv += S[k].w * lam[k]; }
        v = nv;
        if (dot(v, v) < 1e-20) { r.status = 1; r.ns = ns; for (int k = 0; k < ns; k++) r.simplex[k] = S[k]; return r; }
    }
    V3 pa, pb;
    for (int k = 0; k < ns; k++) { pa += S[k].a * lam[k]; pb += S[k].b * lam[k]; }
    double d = norm(v);
    r.dist = d; r.n = v * (1.0 / d); r.pa = pa; r.pb = pb; r.ns = ns;
    for (int k = 0; k < ns; k++) r.simplex[k] = S[k];
    return r;
}

/* ---- EPA on the hull cores (btGjkEpa2 structure: EncloseOrigin, then polytope expansion) */
struct EpaFace { int v[3]; V3 n; double d; bool alive; };

double det3(V3 a, V3 b, V3 c) { return dot(a, cross(b, c)); }

bool enclose_origin(const Hull& A, const Hull& B, SupPt* S, int& ns) {
    switch (ns) {
    case 1:
        for (int i = 0; i < 3; i++) {
            V3 axis; axis[i] = 1;
            S[1] = support(A, B, axis); ns = 2;
            if (enclose_origin(A, B, S, ns)) return true;
            S[1] = support(A, B, -axis); ns = 2;
            if (enclose_origin(A, B, S, ns)) return true;
            ns = 1;
        }
        break;
    case 2: {
        V3 d = S[1].w - S[0].w;
        for (int i = 0; i < 3; i++) {
            V3 axis; axis[i] = 1;
            V3 p = cross(d, axis);
            if (dot(p, p) > 0) {
                S[2] = support(A, B, p); ns = 3;
                if (enclose_origin(A, B, S, ns)) return true;
                S[2] = support(A, B, -p); ns = 3;
                if (enclose_origin(A, B, S, ns)) return true;
                ns = 2;
            }
        }
        break;
    }
    case 3: {
        V3 n = cross(S[1].w - S[0].w, S[2].w - S[0].w);
        if (dot(n, n) > 0) {
            S[3] = support(A, B, n); ns = 4;
            if (enclose_origin(A, B, S, ns)) return true;
            S[3] = support(A, B, -n); ns = 4;
            if (enclose_origin(A, B, S, ns)) return true;
            ns = 3;
        }
        break;
    }
    case 4:
        if (std::fabs(det3(S[0].w - S[3].w, S[1].w - S[3].w, S[2].w - S[3].w)) > 0) return true;
        break;
    }
    return false;
}

bool epa(const Hull& A, const Hull& B, SupPt* S0, int ns0, double& depth, V3& n, V3& pa, V3& pb) {
    const int MAXV = 40, MAXF = 80;   /* same capacities as the CUDA kernel */
    SupPt V[MAXV]; int nv = 0;
    EpaFace F[MAXF]; int nf = 0;
    SupPt S[4]; int ns = ns0;
    for (int k = 0; k < ns0; k++) S[k] = S0[k];
    if (!enclose_origin(A, B, S, ns) || ns != 4) return false;
    /* orient the tetrahedron so that det > 0 */
    if (det3(S[0].w - S[3].w, S[1].w - S[3].w, S[2].w - S[3].w) < 0) std::swap(S[0], S[1]);
    for (int k = 0; k < 4; k++) V[nv++] = S[k];
    auto add_face = [&](int a, int b, int c) -> bool {
        if (nf >= MAXF) return false;
        EpaFace f; f.v[0] = a; f.v[1] = b; f.v[2] = c; f.alive = true;
        V3 nn = cross(V[b].w - V[a].w, V[c].w - V[a].w);
        double l = norm(nn);
        if (l <= 1e-30) return false;
        f.n = nn * (1.0 / l); f.d = dot(f.n, V[a].w);
        if (f.d < 0) { f.n = -f.n; f.d = -f.d; std::swap(f.v[1], f.v[2]); }   /* outward: origin is inside */
        F[nf++] = f;
        return true;
    };
    if (!add_face(0, 1, 2) || !add_face(1, 0, 3) || !add_face(2, 1, 3) || !add_face(0, 2, 3)) return false;
    int best = -1;
    for (int it = 0; it < 128; it++) {
        best = -1; double bd = DBL_MAX;
        for (int f = 0; f < nf; f++) if (F[f].alive && F[f].d < bd) { bd = F[f].d; best = f; }
        if (best < 0) return false;
        SupPt w = support(A, B, F[best].n);
        double wd = dot(w.w, F[best].n) - F[best].d;
        if (wd <= 1e-10 || nv >= MAXV) break;
        /* remove faces visible from w, collect the horizon */
        int eh[2 * MAXF][2]; int ne = 0;
        for (int f = 0; f < nf; f++) {
            if (!F[f].alive) continue;
            if (dot(F[f].n, w.w - V[F[f].v[0]].w) > 0) {
                F[f].alive = false;
                for (int e = 0; e < 3; e++) {
                    int a = F[f].v[e], b = F[f].v[(e + 1) % 3];
                    bool cancelled = false;
                    for (int k = 0; k < ne; k++) if (eh[k][0] == b && eh[k][1] == a) { eh[k][0] = eh[ne - 1][0]; eh[k][1] = eh[ne - 1][1]; ne--; cancelled = true; break; }
                    if (!cancelled) { eh[ne][0] = a; eh[ne][1] = b; ne++; }
                }
            }
        }
        int wi = nv; V[nv++] = w;
        bool ok = true;
        for (int k = 0; k < ne; k++) if (!add_face(eh[k][0], eh[k][1], wi)) { ok = false; break; }
        if (!ok) break;
    }
    const EpaFace& f = F[best];
    depth = f.d; n = f.n;
    /* witness points: barycentric coordinates of the origin's projection on the best face */
    V3 a = V[f.v[0]].w, b = V[f.v[1]].w, c = V[f.v[2]].w;
    V3 pr = f.n * f.d;
    double l[3]; int m;
    closest_triangle(a - pr, b - pr, c - pr, l, &m);
    pa = V[f.v[0]].a * l[0] + V[f.v[1]].a * l[1] + V[f.v[2]].a * l[2];
    pb = V[f.v[0]].b * l[0] + V[f.v[1]].b * l[1] + V[f.v[2]].b * l[2];
    return true;
}

/* signed core distance, normal on B (B -> A), witness points on the cores */
bool closest_points(const Hull& A, const Hull& B, double& dist, V3& n, V3& pa, V3& pb) {
    GjkResult g = gjk(A, B);
    if (g.status == 0) {                       /* cores separated: GJK witnesses */
        if (!(g.dist > 0)) return false;
        dist = g.dist; n = g.n; pa = g.pa; pb = g.pb; return true;
    }
    double depth; V3 en, ea, eb;
    if (!epa(A, B, g.simplex, g.ns, depth, en, ea, eb)) return false;
    /* EPA normal points from the origin to the boundary of A-B: moving A by -n*depth separates; contact normal on B is -n */
    dist = -depth; n = -en; pa = ea; pb = eb;
    return true;
}

/* ---- persistent manifold (btPersistentManifold) */
int cache_entry(const OrcManifold& m, V3 localA, double thr) {
    double shortest = thr * thr; int nearest = -1;
    for (int i = 0; i < m.n; i++) {
        V3 d = m.pt[i].localA - localA; double dd = dot(d, d);
        if (dd < shortest) { shortest = dd; nearest = i; }
    }
    return nearest;
}
int sort_cached_points(const OrcManifold& m, const OrcContactPoint& pt) {
    int maxPenIdx = -1; double maxPen = pt.distance;
    for (int i = 0; i < 4; i++) if (m.pt[i].distance < maxPen) { maxPenIdx = i; maxPen = m.pt[i].distance; }
    double res[4] = {0, 0, 0, 0};
    const V3 p0 = m.pt[0].localA, p1 = m.pt[1].localA, p2 = m.pt[2].localA, p3 = m.pt[3].localA, pn = pt.localA;
    if (maxPenIdx != 0) { V3 c = cross(pn - p1, p3 - p2); res[0] = dot(c, c); }
    if (maxPenIdx != 1) { V3 c = cross(pn - p0, p3 - p2); res[1] = dot(c, c); }
    if (maxPenIdx != 2) { V3 c = cross(pn - p0, p3 - p1); res[2] = dot(c, c); }
    if (maxPenIdx != 3) { V3 c = cross(pn - p0, p2 - p1); res[3] = dot(c, c); }
    int mi = -1; double mv = -1e30;
    for (int i = 0; i < 4; i++) if (std::fabs(res[i]) > mv) { mv = std::fabs(res[i]); mi = i; }
    return mi;
}

}  /* namespace */

/* candidate pairs in a fixed order: robot x objects, robot x table, objects x table, objects x objects */
static int build_pairs(const OrcScene* sc, int (*pairs)[2]) {
    int n = 0;
    for (int l = 0; l <= ORC_NL; l++) for (int o = 0; o < sc->n_objects; o++) { pairs[n][0] = sc->shape_of_link[l]; pairs[n][1] = sc->obj_shape[o]; n++; }
    for (int l = 1; l <= ORC_NL; l++) for (int t = 0; t < 2; t++) { pairs[n][0] = sc->shape_of_link[l]; pairs[n][1] = sc->table_shape[t]; n++; }
    for (int o = 0; o < sc->n_objects; o++) for (int t = 0; t < 2; t++) { pairs[n][0] = sc->obj_shape[o]; pairs[n][1] = sc->table_shape[t]; n++; }
    for (int o = 0; o < sc->n_objects; o++) for (int p = o + 1; p < sc->n_objects; p++) { pairs[n][0] = sc->obj_shape[o]; pairs[n][1] = sc->obj_shape[p]; n++; }
    return n;
}

void orc_collide(OrcEnv* e) {
    const OrcScene* sc = e->sc; const OrcConfig& cf = e->cfg;
    const double thr = cf.contact_breaking;
    int S = (int)sc->shapes.size();
    std::vector<OrcFrame> fr(S);
    std::vector<V3> amin(S), amax(S);
    const double* lmin = sc->blob.f64("shape.aabb_min"); const double* lmax = sc->blob.f64("shape.aabb_max");
    for (int s = 0; s < S; s++) {
        fr[s] = orc_shape_frame(e, s);
        V3 c((lmin[3 * s] + lmax[3 * s]) * 0.5, (lmin[3 * s + 1] + lmax[3 * s + 1]) * 0.5, (lmin[3 * s + 2] + lmax[3 * s + 2]) * 0.5);
        V3 h((lmax[3 * s] - lmin[3 * s]) * 0.5 + cf.margin, (lmax[3 * s + 1] - lmin[3 * s + 1]) * 0.5 + cf.margin, (lmax[3 * s + 2] - lmin[3 * s + 2]) * 0.5 + cf.margin);
        V3 wc = fr[s].R * c + fr[s].p;
        V3 ext;
        for (int i = 0; i < 3; i++) ext[i] = std::fabs(fr[s].R.m[i][0]) * h.x + std::fabs(fr[s].R.m[i][1]) * h.y + std::fabs(fr[s].R.m[i][2]) * h.z + thr;
        amin[s] = wc - ext; amax[s] = wc + ext;
    }
    int pairs[128][2];
    int np = build_pairs(sc, pairs);
    /* mark: manifolds of pairs whose AABBs no longer overlap are destroyed (broadphase pair removed) */
    OrcManifold old[ORC_MAX_MANIFOLDS]; int nold = e->n_manifolds;
    for (int i = 0; i < nold; i++) old[i] = e->manifolds[i];
    e->n_manifolds = 0; e->last_overlaps = 0; e->manifold_overflow = 0;
    for (int pi = 0; pi < np; pi++) {
        int a = pairs[pi][0], b = pairs[pi][1];
        bool overlap = amin[a].x <= amax[b].x && amax[a].x >= amin[b].x && amin[a].y <= amax[b].y && amax[a].y >= amin[b].y &&
                       amin[a].z <= amax[b].z && amax[a].z >= amin[b].z;
        if (!overlap) continue;
        OrcManifold m; m.shapeA = a; m.shapeB = b; m.n = 0; m.touched = 1;
        for (int i = 0; i < nold; i++) if (old[i].shapeA == a && old[i].shapeB == b) { m = old[i]; break; }
        Hull A{sc->hull_verts + 3 * sc->shapes[a].hull_start, sc->shapes[a].hull_count, fr[a].R, fr[a].p};
        Hull B{sc->hull_verts + 3 * sc->shapes[b].hull_start, sc->shapes[b].hull_count, fr[b].R, fr[b].p};
        double dcore; V3 n, pa, pb;
        if (closest_points(A, B, dcore, n, pa, pb)) {
            double distance = dcore - 2 * cf.margin;
            V3 onB = pb + n * cf.margin;
            if (distance <= thr) {           /* btManifoldResult::addContactPoint */
                OrcContactPoint cp;
                V3 onA = onB + n * distance;
                cp.localA = transpose(fr[a].R) * (onA - fr[a].p);
                cp.localB = transpose(fr[b].R) * (onB - fr[b].p);
                cp.worldA = onA; cp.worldB = onB; cp.normalB = n; cp.distance = distance;
                cp.applied_impulse = 0; cp.applied_f1 = cp.applied_f2 = 0; cp.lifetime = 0;
                int idx = cache_entry(m, cp.localA, thr);
                if (idx >= 0) {      /* replaceContactPoint keeps the warm-start impulse */
                    cp.applied_impulse = m.pt[idx].applied_impulse; cp.lifetime = m.pt[idx].lifetime;
                    m.pt[idx] = cp;
                } else if (m.n < 4) m.pt[m.n++] = cp;
                else m.pt[sort_cached_points(m, cp)] = cp;
            }
        }
        /* refreshContactPoints */
        for (int i = m.n - 1; i >= 0; i--) {
            OrcContactPoint& c = m.pt[i];
            c.worldA = fr[a].R * c.localA + fr[a].p;
            c.worldB = fr[b].R * c.localB + fr[b].p;
            c.distance = dot(c.worldA - c.worldB, c.normalB);
            c.lifetime++;
        }
        for (int i = m.n - 1; i >= 0; i--) {
            OrcContactPoint& c = m.pt[i];
            bool drop = false;
            if (!(c.distance <= thr)) drop = true;
            else {
                V3 proj = c.worldA - c.normalB * c.distance;
                V3 d = c.worldB - proj;
                if (dot(d, d) > thr * thr) drop = true;
            }
            if (drop) { m.pt[i] = m.pt[m.n - 1]; m.n--; }     /* removeContactPoint: swap with last */
        }
        /* A manifold without points carries no state (a later hit on it finds n = 0, the same as no hit), so only
         * non-empty manifolds take one of the ORC_MAX_MANIFOLDS slots: AABB overlaps without contact cannot crowd out the
         * object x table pairs at the end of the pair order.  A non-empty manifold that still finds no slot is flagged. */
        e->last_overlaps++;
        if (m.n > 0) {
            if (e->n_manifolds < ORC_MAX_MANIFOLDS) e->manifolds[e->n_manifolds++] = m;
            else e->manifold_overflow = 1;
        }
    }
}

/* test hook: signed core distance between two scene shapes placed at explicit frames */
bool orc_closest_points_shapes(const OrcScene* sc, int sa, const OrcFrame& fa, int sb, const OrcFrame& fb, double* out10) {
    Hull A{sc->hull_verts + 3 * sc->shapes[sa].hull_start, sc->shapes[sa].hull_count, fa.R, fa.p};
    Hull B{sc->hull_verts + 3 * sc->shapes[sb].hull_start, sc->shapes[sb].hull_count, fb.R, fb.p};
    double d; V3 n, pa, pb;
    if (!closest_points(A, B, d, n, pa, pb)) return false;
    out10[0] = d; out10[1] = n.x; out10[2] = n.y; out10[3] = n.z; out10[4] = pa.x; out10[5] = pa.y; out10[6] = pa.z; out10[7] = pb.x; out10[8] = pb.y; out10[9] = pb.z;
    return true;
}
