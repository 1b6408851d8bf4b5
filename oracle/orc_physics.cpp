/* TEST INFRASTRUCTURE — CPU oracle: restatement of pybullet's stepSimulation for the REALRobot2020 scene.
 *
 * Follows (reference call sites):
 *   real_robots/envs/env.py:340        scene.global_step() -> World.step -> stepSimulation
 *   real_robots/envs/env.py:202-204    gravity 9.81, timestep 0.005, frame_skip 1
 *   real_robots/envs/env.py:314-321    limitActionByJoint
 *   real_robots/envs/env.py:257-264    control_objects_limits
 *   real_robots/envs/robot.py:188-201  apply_action   (POSITION_CONTROL motors)
 *   real_robots/envs/robot.py:203-211  calc_state
 *   real_robots/envs/robot.py:131-163  get_contacts / get_touch_sensors
 * and the third-party algorithm those calls execute: bullet3 (pybullet, UNPINNED, un-vendored:
 * setup.py:24-32) — btMultiBodyDynamicsWorld::solveConstraints, btMultiBody::
 * computeAccelerationsArticulatedBodyAlgorithmMultiDof / calcAccelerationDeltasMultiDof,
 * btMultiBodyJointMotor, btMultiBodyJointLimitConstraint, btMultiBodyConstraintSolver.
 * PARITY UNPINNED for the Bullet arithmetic: written from the published algorithm; no pybullet available to pin against.
 * The action prologue and the object-limit teleport of orc_env_step ARE pinned: tests/test_reference_golden.py checks them
 * against vectors produced by the reference's own env.py / robot.py (tests/golden/make_reference_golden.py).
 */
#include "orc_env.h"
#include <algorithm>
#include <cstdio>

typedef double D6[6][6];

static inline void s6_to_arr(const S6& s, double* a) { a[0] = s.a.x; a[1] = s.a.y; a[2] = s.a.z; a[3] = s.l.x; a[4] = s.l.y; a[5] = s.l.z; }
static inline S6 arr_to_s6(const double* a) { return S6(V3(a[0], a[1], a[2]), V3(a[3], a[4], a[5])); }

/* motion transform parent -> child:  w' = R w ;  v' = R v + (R w) x r  =  R v - [r]x R w */
static void build_X(const M3& R, V3 r, D6 X) {
    M3 K = skew(r) * R;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        X[i][j] = R.m[i][j];      X[i][j + 3] = 0;
        X[i + 3][j] = -K.m[i][j]; X[i + 3][j + 3] = R.m[i][j];
    }
}
static S6 xf_motion(const M3& R, V3 r, S6 v) { V3 w = R * v.a; return S6(w, R * v.l + cross(w, r)); }
/* force transform child -> parent (X^T):  f_p = R^T f ;  t_p = R^T (t + r x f) */
static S6 xf_force_T(const M3& R, V3 r, S6 f) { M3 Rt = transpose(R); return S6(Rt * (f.a + cross(r, f.l)), Rt * f.l); }
static S6 cross_motion(S6 a, S6 b) { return S6(cross(a.a, b.a), cross(a.l, b.a) + cross(a.a, b.l)); }
static S6 mul66(const D6 I, S6 v) {
    double x[6], y[6]; s6_to_arr(v, x);
    for (int i = 0; i < 6; i++) { double s = 0; for (int j = 0; j < 6; j++) s += I[i][j] * x[j]; y[i] = s; }
    return arr_to_s6(y);
}

void orc_env_init(OrcEnv* e, const OrcScene* sc, const OrcConfig* cfg) {
    e->sc = sc; e->cfg = *cfg;
    orc_env_reset(e);
}

void orc_env_reset(OrcEnv* e) {
    /* robot.py:165-185: every joint reset_current_position(0,0); objects at object_poses; env.py:217 timestep=0 */
    for (int i = 0; i < ORC_NDOF; i++) { e->q[i] = 0; e->qd[i] = 0; e->target[i] = 0; }
    for (int o = 0; o < ORC_MAXOBJ; o++) {
        e->obj[o].pos = e->sc->obj_init_pos[o]; e->obj[o].quat = e->sc->obj_init_quat[o];
        e->obj[o].lin = V3(); e->obj[o].ang = V3();
    }
    e->timestep = 0;
    e->n_manifolds = 0;
    for (int i = 0; i < 4; i++) e->touch[i] = 0;
    e->last_rows = e->last_iterations = e->last_contacts = 0; e->last_residual = 0;
    e->last_overlaps = 0; e->manifold_overflow = 0;
    orc_fk(e);
}

static S6 joint_axis(const OrcScene* sc, int i) {   /* btMultiBody::setupRevolute: top = axis, bottom = axis x e */
    return S6(sc->axis[i], cross(sc->axis[i], sc->evec[i]));
}

void orc_fk(OrcEnv* e) {
    const OrcScene* sc = e->sc;
    for (int i = 0; i < ORC_NL; i++) {
        if (sc->jtype[i] == 1) e->Rpc[i] = axis_angle(sc->axis[i], -e->q[sc->dof[i]]) * sc->rot0[i];
        else e->Rpc[i] = sc->rot0[i];
        e->rvec[i] = e->Rpc[i] * sc->dvec[i] + sc->evec[i];
        M3 Rp = M3::identity(); V3 pp = sc->base_pos;
        if (sc->parent[i] >= 0) { Rp = e->link[sc->parent[i]].R; pp = e->link[sc->parent[i]].p; }
        e->link[i].R = Rp * transpose(e->Rpc[i]);
        e->link[i].p = pp + e->link[i].R * e->rvec[i];
    }
}

OrcFrame orc_shape_frame(const OrcEnv* e, int s) {
    const OrcScene* sc = e->sc;
    const OrcShape& sh = sc->shapes[s];
    OrcFrame f;
    if (sh.body == 0) {
        if (sh.link < 0) { f.R = M3::identity(); f.p = sc->base_pos; }
        else f = e->link[sh.link];
    } else if (sh.body == 1) {
        f.R = quat_to_mat(sc->table_quat);
        f.p = sc->table_pos;
        if (sh.link >= 0) f.p = f.p + f.R * sc->table_upper_offset;
    } else {
        const OrcObjState& o = e->obj[sh.body - 2];
        f.R = quat_to_mat(o.quat); f.p = o.pos;
    }
    return f;
}

/* ---- Featherstone ABA in link-COM body coordinates (btMultiBody::computeAccelerationsArticulatedBodyAlgorithmMultiDof)
 * fixed base, gravity as applied link force, per-link damping k(1+|v|), gyroscopic term on. */
static void aba_internal(OrcEnv* e, const double* tau, double* qdd) {
    const OrcScene* sc = e->sc; const OrcConfig& cf = e->cfg;
    S6 v[ORC_NL], c[ORC_NL], p[ORC_NL];
    double Y[ORC_NL];
    for (int i = 0; i < ORC_NL; i++) {
        S6 vp; if (sc->parent[i] >= 0) vp = v[sc->parent[i]];
        v[i] = xf_motion(e->Rpc[i], e->rvec[i], vp);
        S6 vj;
        if (sc->jtype[i] == 1) { vj = joint_axis(sc, i) * e->qd[sc->dof[i]]; v[i] = v[i] + vj; }
        c[i] = cross_motion(v[i], vj);
        M3 Rt = transpose(e->link[i].R);                 /* world -> link */
        V3 I = sc->inertia[i]; double m = sc->mass[i];
        V3 w = v[i].a, l = v[i].l;
        V3 Iw(I.x * w.x, I.y * w.y, I.z * w.z);
        p[i].a = V3(); p[i].l = -(Rt * V3(0, 0, m * cf.gravity));
        p[i].a += Iw * (cf.ang_damping + cf.ang_damping * norm(w));
        p[i].l += (l * m) * (cf.lin_damping + cf.lin_damping * norm(l));
        p[i].a += cross(w, Iw);
        p[i].l += cross(w, l) * m;
        for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) e->IA[i][a][b] = 0;
        e->IA[i][0][0] = I.x; e->IA[i][1][1] = I.y; e->IA[i][2][2] = I.z;
        e->IA[i][3][3] = e->IA[i][4][4] = e->IA[i][5][5] = m;
    }
    for (int i = ORC_NL - 1; i >= 0; i--) {
        D6 Ia; S6 pa;
        S6 Ic = mul66(e->IA[i], c[i]);
        if (sc->jtype[i] == 1) {
            S6 S = joint_axis(sc, i);
            S6 hh = mul66(e->IA[i], S);
            double D = sdot(S, hh);
            e->invD[i] = 1.0 / D;
            s6_to_arr(hh, e->h[i]);
            Y[i] = tau[sc->dof[i]] - sdot(S, p[i]) - sdot(c[i], hh);
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) Ia[a][b] = e->IA[i][a][b] - e->h[i][a] * e->invD[i] * e->h[i][b];
            pa = p[i] + Ic + hh * (e->invD[i] * Y[i]);
        } else {
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) Ia[a][b] = e->IA[i][a][b];
            for (int a = 0; a < 6; a++) e->h[i][a] = 0;
            e->invD[i] = 0; Y[i] = 0;
            pa = p[i] + Ic;
        }
        int par = sc->parent[i];
        if (par >= 0) {
            D6 X, T; build_X(e->Rpc[i], e->rvec[i], X);
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) { double s = 0; for (int k = 0; k < 6; k++) s += Ia[a][k] * X[k][b]; T[a][b] = s; }
            for (int a = 0; a < 6; a++) for (int b = 0; b < 6; b++) { double s = 0; for (int k = 0; k < 6; k++) s += X[k][a] * T[k][b]; e->IA[par][a][b] += s; }
            p[par] = p[par] + xf_force_T(e->Rpc[i], e->rvec[i], pa);
        }
    }
    S6 acc[ORC_NL];
    for (int i = 0; i < ORC_NL; i++) {
        S6 ap; if (sc->parent[i] >= 0) ap = acc[sc->parent[i]];
        S6 a = xf_motion(e->Rpc[i], e->rvec[i], ap);
        if (sc->jtype[i] == 1) {
            double qa = e->invD[i] * (Y[i] - sdot(arr_to_s6(e->h[i]), a));
            qdd[sc->dof[i]] = qa;
            a = a + c[i] + joint_axis(sc, i) * qa;
        } else a = a + c[i];
        acc[i] = a;
    }
}
void orc_aba(OrcEnv* e, const double* tau, double* qdd) { orc_fk(e); aba_internal(e, tau, qdd); }

/* M^-1 * f for a generalised joint force f, using the cached articulated inertias
 * (btMultiBody::calcAccelerationDeltasMultiDof, fixed base) */
static void arm_response(const OrcEnv* e, const double* f, double* out) {
    const OrcScene* sc = e->sc;
    S6 pz[ORC_NL]; double Yd[ORC_NL];
    for (int i = ORC_NL - 1; i >= 0; i--) {
        S6 pa = pz[i];
        if (sc->jtype[i] == 1) {
            Yd[i] = f[sc->dof[i]] - sdot(joint_axis(sc, i), pz[i]);
            pa = pa + arr_to_s6(e->h[i]) * (e->invD[i] * Yd[i]);
        }
        int par = sc->parent[i];
        if (par >= 0) pz[par] = pz[par] + xf_force_T(e->Rpc[i], e->rvec[i], pa);
    }
    S6 acc[ORC_NL];
    for (int i = 0; i < ORC_NL; i++) {
        S6 ap; if (sc->parent[i] >= 0) ap = acc[sc->parent[i]];
        S6 a = xf_motion(e->Rpc[i], e->rvec[i], ap);
        if (sc->jtype[i] == 1) {
            double qa = e->invD[i] * (Yd[i] - sdot(arr_to_s6(e->h[i]), a));
            out[sc->dof[i]] = qa;
            a = a + joint_axis(sc, i) * qa;
        }
        acc[i] = a;
    }
}

void orc_mass_matrix(OrcEnv* e, double* Minv) {   /* returns M^-1 (11x11) from unit responses; tests invert it */
    double tau[ORC_NDOF] = {0}, qdd[ORC_NDOF];
    orc_fk(e); aba_internal(e, tau, qdd);
    for (int j = 0; j < ORC_NDOF; j++) {
        double f[ORC_NDOF] = {0}; f[j] = 1; double r[ORC_NDOF];
        arm_response(e, f, r);
        for (int i = 0; i < ORC_NDOF; i++) Minv[i * ORC_NDOF + j] = r[i];
    }
}

/* ---------------------------------------------------------------- constraint rows + PGS */
struct Row {
    bool arm; double Ja[ORC_NDOF], ra[ORC_NDOF];
    int ob[2]; double Jo[2][6], ro[2][6];           /* object parts: [ang(3), lin(3)] world frame */
    double rhs, cfm, lo, hi, jacDiagABInv, applied;
    int friction_index; double mu;                  /* for friction rows */
    OrcContactPoint* cp; int kind;                  /* 0 generic, 1 normal, 2 friction1, 3 friction2 */
};

struct Solver {
    std::vector<Row> noncontact, normal, spinning, rolling, friction;
    double dva[ORC_NDOF]; double dvo[ORC_MAXOBJ][6];
};

static M3 obj_inv_inertia_world(const OrcEnv* e, int o) {
    M3 R = quat_to_mat(e->obj[o].quat);
    V3 I = e->sc->obj_inertia[o];
    M3 Di = M3::diag(V3(1.0 / I.x, 1.0 / I.y, 1.0 / I.z));
    return R * Di * transpose(R);
}

static void row_finish(const OrcEnv* e, Row& r) {
    /* unit-impulse responses and the diagonal term d = J M^-1 J^T */
    double d = 0;
    if (r.arm) { arm_response(e, r.Ja, r.ra); for (int i = 0; i < ORC_NDOF; i++) d += r.Ja[i] * r.ra[i]; }
    for (int k = 0; k < 2; k++) if (r.ob[k] >= 0) {
        int o = r.ob[k];
        M3 Ii = obj_inv_inertia_world(e, o);
        V3 ang = Ii * V3(r.Jo[k][0], r.Jo[k][1], r.Jo[k][2]);
        double im = 1.0 / e->sc->obj_mass[o];
        r.ro[k][0] = ang.x; r.ro[k][1] = ang.y; r.ro[k][2] = ang.z;
        r.ro[k][3] = r.Jo[k][3] * im; r.ro[k][4] = r.Jo[k][4] * im; r.ro[k][5] = r.Jo[k][5] * im;
        for (int i = 0; i < 6; i++) d += r.Jo[k][i] * r.ro[k][i];
    }
    r.jacDiagABInv = d > 0 ? 1.0 / (d + r.cfm) : 0.0;
}
static double row_relvel(const OrcEnv* e, const Row& r) {
    double v = 0;
    if (r.arm) for (int i = 0; i < ORC_NDOF; i++) v += r.Ja[i] * e->qd[i];
    for (int k = 0; k < 2; k++) if (r.ob[k] >= 0) {
        const OrcObjState& o = e->obj[r.ob[k]];
        v += r.Jo[k][0] * o.ang.x + r.Jo[k][1] * o.ang.y + r.Jo[k][2] * o.ang.z
           + r.Jo[k][3] * o.lin.x + r.Jo[k][4] * o.lin.y + r.Jo[k][5] * o.lin.z;
    }
    return v;
}
static void row_init(Row& r) {
    r.arm = false; r.ob[0] = r.ob[1] = -1;
    for (int i = 0; i < ORC_NDOF; i++) r.Ja[i] = r.ra[i] = 0;
    for (int k = 0; k < 2; k++) for (int i = 0; i < 6; i++) r.Jo[k][i] = r.ro[k][i] = 0;
    r.rhs = r.cfm = 0; r.lo = 0; r.hi = 0; r.jacDiagABInv = 0; r.applied = 0;
    r.friction_index = -1; r.mu = 0; r.cp = nullptr; r.kind = 0;
}
static void apply_impulse(Solver& S, const Row& r, double imp) {
    if (r.arm) for (int i = 0; i < ORC_NDOF; i++) S.dva[i] += r.ra[i] * imp;
    for (int k = 0; k < 2; k++) if (r.ob[k] >= 0) for (int i = 0; i < 6; i++) S.dvo[r.ob[k]][i] += r.ro[k][i] * imp;
}
static double row_dv(const Solver& S, const Row& r) {
    double v = 0;
    if (r.arm) for (int i = 0; i < ORC_NDOF; i++) v += r.Ja[i] * S.dva[i];
    for (int k = 0; k < 2; k++) if (r.ob[k] >= 0) for (int i = 0; i < 6; i++) v += r.Jo[k][i] * S.dvo[r.ob[k]][i];
    return v;
}
/* btMultiBodyConstraintSolver::resolveSingleConstraintRowGeneric */
static double resolve_row(Solver& S, Row& c) {
    double delta = c.rhs - c.applied * c.cfm;
    delta -= row_dv(S, c) * c.jacDiagABInv;
    double sum = c.applied + delta;
    if (sum < c.lo) { delta = c.lo - c.applied; c.applied = c.lo; }
    else if (sum > c.hi) { delta = c.hi - c.applied; c.applied = c.hi; }
    else c.applied = sum;
    apply_impulse(S, c, delta);
    return c.jacDiagABInv != 0 ? delta * (1.0 / c.jacDiagABInv) : 0.0;     /* deltaImpulse * (1. / m_jacDiagABInv) */
}
/* btMultiBodyConstraintSolver::resolveConeFrictionConstraintRows (implicit friction cone over the two directions) */
static double resolve_cone(Solver& S, Row& a, Row& b, double limit) {
    Row* rr[2] = {&a, &b};
    double want[2], res = 0;
    for (int k = 0; k < 2; k++) {
        Row& c = *rr[k];
        double delta = c.rhs - c.applied * c.cfm - row_dv(S, c) * c.jacDiagABInv;
        want[k] = c.applied + delta;
    }
    double len2 = want[0] * want[0] + want[1] * want[1];
    if (len2 > limit * limit) { double s = len2 > 0 ? limit / std::sqrt(len2) : 0; want[0] *= s; want[1] *= s; }
    for (int k = 0; k < 2; k++) {
        Row& c = *rr[k];
        double delta = want[k] - c.applied;
        c.applied = want[k];
        apply_impulse(S, c, delta);
        double r = c.jacDiagABInv != 0 ? delta * (1.0 / c.jacDiagABInv) : 0.0;
        res = std::max(res, r * r);
    }
    return res;
}

/* world-space Jacobian of a point fixed on robot link `li` along direction n (btMultiBody::fillContactJacobianMultiDof) */
static void arm_point_jacobian(const OrcEnv* e, int li, V3 pw, V3 n, double* J) {
    const OrcScene* sc = e->sc;
    for (int i = 0; i < ORC_NDOF; i++) J[i] = 0;
    for (int k = li; k >= 0; k = sc->parent[k]) {
        if (sc->jtype[k] != 1) continue;
        V3 aw = e->link[k].R * sc->axis[k];
        V3 pivot = e->link[k].p - e->link[k].R * sc->evec[k];
        J[sc->dof[k]] = dot(n, cross(aw, pw - pivot));
    }
}
static void arm_angular_jacobian(const OrcEnv* e, int li, V3 n, double* J) {
    const OrcScene* sc = e->sc;
    for (int i = 0; i < ORC_NDOF; i++) J[i] = 0;
    for (int k = li; k >= 0; k = sc->parent[k]) {
        if (sc->jtype[k] != 1) continue;
        J[sc->dof[k]] = dot(n, e->link[k].R * sc->axis[k]);
    }
}

/* btPlaneSpace1 */
static void plane_space(V3 n, V3& p, V3& q) {
    if (std::fabs(n.z) > 0.7071067811865475244008443621048490) {
        double a = n.y * n.y + n.z * n.z; double k = 1.0 / std::sqrt(a);
        p = V3(0, -n.z * k, n.y * k);
        q = V3(a * k, -n.x * p.z, n.x * p.y);
    } else {
        double a = n.x * n.x + n.y * n.y; double k = 1.0 / std::sqrt(a);
        p = V3(-n.y * k, n.x * k, 0);
        q = V3(-n.z * p.y, n.z * p.x, a * k);
    }
}

/* fill the Jacobian of one side of a contact row (side sign = +1 for A, -1 for B) */
static void fill_side(const OrcEnv* e, Row& r, int shape, V3 pw, V3 dir, double sign, bool angular_only) {
    const OrcShape& sh = e->sc->shapes[shape];
    V3 d = dir * sign;
    if (sh.body == 0) {
        if (sh.link < 0) return;                  /* fixed base link_0: static */
        double J[ORC_NDOF];
        if (angular_only) arm_angular_jacobian(e, sh.link, d, J); else arm_point_jacobian(e, sh.link, pw, d, J);
        r.arm = true;
        for (int i = 0; i < ORC_NDOF; i++) r.Ja[i] += J[i];
    } else if (sh.body >= 2) {
        int o = sh.body - 2;
        int k = r.ob[0] < 0 ? 0 : 1;
        r.ob[k] = o;
        if (angular_only) { r.Jo[k][0] = d.x; r.Jo[k][1] = d.y; r.Jo[k][2] = d.z; }
        else {
            V3 t = cross(pw - e->obj[o].pos, d);
            r.Jo[k][0] = t.x; r.Jo[k][1] = t.y; r.Jo[k][2] = t.z;
            r.Jo[k][3] = d.x; r.Jo[k][4] = d.y; r.Jo[k][5] = d.z;
        }
    }   /* body 1 = table: static, no Jacobian */
}

static double restitution_curve(double rel_vel, double restitution, double threshold) {
    if (std::fabs(rel_vel) < threshold) return 0.0;
    return restitution * -rel_vel;
}

static void build_rows(OrcEnv* e, Solver& S) {
    const OrcScene* sc = e->sc; const OrcConfig& cf = e->cfg;
    double dt = cf.dt;
    /* 1. joint limit rows (btMultiBodyJointLimitConstraint), link order; only joints with lower <= upper */
    for (int i = 0; i < ORC_NL; i++) {
        if (sc->jtype[i] != 1 || sc->lower[i] > sc->upper[i]) continue;
        int d = sc->dof[i];
        double pos[2] = {e->q[d] - sc->lower[i], sc->upper[i] - e->q[d]};
        for (int row = 0; row < 2; row++) {
            double pen = pos[row];
            if (pen > 0) continue;
            Row r; row_init(r);
            r.arm = true; r.Ja[d] = row ? -1.0 : 1.0;
            r.lo = 0; r.hi = cf.limit_max_impulse;
            row_finish(e, r);
            double rel_vel = row_relvel(e, r);
            double positional = -pen * cf.erp / dt;     /* pen <= 0; |pen| < split-impulse threshold -> erp */
            double velocity = -rel_vel;
            r.rhs = (positional + velocity) * r.jacDiagABInv;
            S.noncontact.push_back(r);
        }
    }
    /* 2. POSITION_CONTROL motors (btMultiBodyJointMotor), link order */
    for (int i = 0; i < ORC_NL; i++) {
        if (sc->jtype[i] != 1) continue;
        int d = sc->dof[i];
        Row r; row_init(r);
        r.arm = true; r.Ja[d] = 1.0;
        double maxImp = cf.motor_max_force * dt;
        r.lo = -maxImp; r.hi = maxImp;
        row_finish(e, r);
        double cur = e->qd[d];
        double stab = 1.0 * (e->target[d] - e->q[d]) / dt;            /* m_erp = 1 */
        double desired = cf.motor_kp * stab + cur + cf.motor_kd * (0.0 - cur);
        double rel_vel = row_relvel(e, r);
        r.rhs = (desired - rel_vel) * r.jacDiagABInv;
        S.noncontact.push_back(r);
    }
    /* 3. contacts (btMultiBodyConstraintSolver::convertMultiBodyContact) */
    e->last_contacts = 0;
    if (!cf.enable_contacts) return;
    for (int m = 0; m < e->n_manifolds; m++) {
        OrcManifold& mf = e->manifolds[m];
        const OrcShape& sA = sc->shapes[mf.shapeA]; const OrcShape& sB = sc->shapes[mf.shapeB];
        double mu = sA.friction * sB.friction;
        double rest = sA.restitution * sB.restitution;
        /* btManifoldResult::calculateCombinedRollingFriction / ...SpinningFriction: r0*f1 + r1*f0, clamped to +-10 */
        double roll = std::max(-10.0, std::min(10.0, sA.rolling * sB.friction + sB.rolling * sA.friction));
        double spin = std::max(-10.0, std::min(10.0, sA.spinning * sB.friction + sB.spinning * sA.friction));
        if (!cf.torsional_friction) roll = spin = 0.0;
        int rolling_budget = 4;
        for (int k = 0; k < mf.n; k++) {
            OrcContactPoint& cp = mf.pt[k];
            if (cp.distance > cf.contact_breaking) continue;   /* getProcessingThreshold */
            if (e->last_contacts >= 48) continue;              /* solver row capacity shared with the CUDA kernel (RR_MAX_CONTACTS) */
            e->last_contacts++;
            Row r; row_init(r); r.kind = 1; r.cp = &cp;
            fill_side(e, r, mf.shapeA, cp.worldA, cp.normalB, +1.0, false);
            fill_side(e, r, mf.shapeB, cp.worldB, cp.normalB, -1.0, false);
            r.lo = 0; r.hi = 1e10;
            row_finish(e, r);
            double rel_vel = row_relvel(e, r);
            double restitution = restitution_curve(rel_vel, rest, cf.restitution_threshold);
            if (restitution <= 0) restitution = 0;
            double pen = cp.distance + cf.linear_slop;
            double positional = 0, velocity = restitution - rel_vel;
            double erp = (pen > -0.04) ? cf.erp : cf.erp2;   /* split impulse threshold: combine with m_erp when shallow */
            if (pen > 0) velocity -= pen / dt; else positional = -pen * erp / dt;
            r.rhs = (positional + velocity) * r.jacDiagABInv;
            r.applied = cp.applied_impulse * cf.warmstart;
            int nidx = (int)S.normal.size();
            S.normal.push_back(r);
            V3 t1, t2; plane_space(cp.normalB, t1, t2);
            t1 = normalized(t1); t2 = normalized(t2);
            V3 dirs[2] = {t1, t2};
            /* torsional friction (btMultiBodyConstraintSolver::convertMultiBodyContact: `rollingFriction` counts 4 points per
             * manifold; addMultiBodySpinningFrictionConstraint along the normal when the combined spinning friction is > 0,
             * addMultiBodyTorsionalFrictionConstraint along both lateral directions when the combined rolling friction is > 0).
             * Rows are angular only, target velocity 0, no warm start, limits +-(coefficient x normal impulse) set per iteration.
             * Reference data: urdf/cube.urdf:6-7, tomato.urdf:6-7, mustard.urdf:6-7, kuka_gripper.urdf:294-295 (fingers, skins). */
            if (rolling_budget > 0) {
                for (int f = 0; f < 3; f++) {
                    double coef = f == 0 ? spin : roll;
                    if (!(coef > 0)) continue;
                    V3 dir = f == 0 ? cp.normalB : dirs[f - 1];
                    Row tr; row_init(tr); tr.kind = 4 + f; tr.cp = &cp;
                    fill_side(e, tr, mf.shapeA, cp.worldA, dir, +1.0, true);
                    fill_side(e, tr, mf.shapeB, cp.worldB, dir, -1.0, true);
                    tr.friction_index = nidx; tr.mu = coef;
                    row_finish(e, tr);
                    tr.rhs = (0.0 - row_relvel(e, tr)) * tr.jacDiagABInv;
                    tr.lo = -coef; tr.hi = coef;
                    (f == 0 ? S.spinning : S.rolling).push_back(tr);
                }
                rolling_budget--;
            }
            for (int f = 0; f < 2; f++) {
                Row fr; row_init(fr); fr.kind = 2 + f; fr.cp = &cp;
                fill_side(e, fr, mf.shapeA, cp.worldA, dirs[f], +1.0, false);
                fill_side(e, fr, mf.shapeB, cp.worldB, dirs[f], -1.0, false);
                fr.friction_index = nidx; fr.mu = mu;
                row_finish(e, fr);
                double rv = row_relvel(e, fr);
                fr.rhs = (0.0 - rv) * fr.jacDiagABInv;
                fr.lo = -mu; fr.hi = mu;
                S.friction.push_back(fr);
            }
        }
    }
}

static void solve(OrcEnv* e, Solver& S) {
    const OrcConfig& cf = e->cfg;
    for (int i = 0; i < ORC_NDOF; i++) S.dva[i] = 0;
    for (int o = 0; o < ORC_MAXOBJ; o++) for (int i = 0; i < 6; i++) S.dvo[o][i] = 0;
    /* warm start (normal rows only) */
    for (size_t j = 0; j < S.normal.size(); j++) if (S.normal[j].applied != 0) apply_impulse(S, S.normal[j], S.normal[j].applied);
    int it = 0; double res = 0;
    for (it = 0; it < cf.solver_iterations; it++) {
        res = 0;
        int n = (int)S.noncontact.size();
        for (int j = 0; j < n; j++) {
            int idx = (it & 1) ? j : n - 1 - j;
            double r = resolve_row(S, S.noncontact[idx]);
            res = std::max(res, r * r);
        }
        for (size_t j = 0; j < S.normal.size(); j++) { double r = resolve_row(S, S.normal[j]); res = std::max(res, r * r); }
        for (int pass = 0; pass < 2; pass++) {       /* spinning rows, then rolling rows: generic rows whose limits follow the normal impulse */
            std::vector<Row>& T = pass == 0 ? S.spinning : S.rolling;
            for (size_t j = 0; j < T.size(); j++) {
                double total = S.normal[T[j].friction_index].applied;
                if (total > 0) {
                    T[j].lo = -(T[j].mu * total); T[j].hi = T[j].mu * total;
                    double r = resolve_row(S, T[j]); res = std::max(res, r * r);
                }
            }
        }
        for (size_t j = 0; j + 1 < S.friction.size(); j += 2) {
            Row& a = S.friction[j]; Row& b = S.friction[j + 1];
            double total = S.normal[a.friction_index].applied;
            double lim = a.mu * total;
            a.lo = b.lo = -lim; a.hi = b.hi = lim;
            res = std::max(res, resolve_cone(S, a, b, lim));
        }
        if (res <= cf.residual_threshold) { it++; break; }
    }
    e->last_iterations = it; e->last_residual = res;
    e->last_rows = (int)(S.noncontact.size() + S.normal.size() + S.spinning.size() + S.rolling.size() + S.friction.size());
    /* write back applied impulses (solveGroupCacheFriendlyFinish) */
    for (size_t j = 0; j < S.normal.size(); j++) S.normal[j].cp->applied_impulse = S.normal[j].applied;
    for (size_t j = 0; j < S.friction.size(); j++) {
        if (S.friction[j].kind == 2) S.friction[j].cp->applied_f1 = S.friction[j].applied;
        else S.friction[j].cp->applied_f2 = S.friction[j].applied;
    }
}

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* exponential-map quaternion integration (btMultiBody::stepPositionsMultiDof, base body, world-frame omega) */
static Q4 integrate_quat(Q4 q, V3 w, double dt) {
    double ang = norm(w);
    if (ang * dt > 0.7853981633974483) ang = 0.5 * 1.5707963267948966 / dt;    /* ANGULAR_MOTION_THRESHOLD */
    V3 ax;
    if (ang < 0.001) ax = w * (0.5 * dt - (dt * dt * dt) * 0.020833333333 * ang * ang);
    else ax = w * (std::sin(0.5 * ang * dt) / ang);
    Q4 dq(ax.x, ax.y, ax.z, std::cos(ang * dt * 0.5));
    return qnormalize(qmul(dq, q));
}

void orc_step_simulation(OrcEnv* e) {
    const OrcScene* sc = e->sc; const OrcConfig& cf = e->cfg;
    double dt = cf.dt;
    orc_fk(e);
    /* collision detection on the poses at the start of the step */
    if (cf.enable_contacts) orc_collide(e); else e->n_manifolds = 0;
    /* forward dynamics: joint damping torque (pybullet adds -damping*qd), ABA, v += a*dt */
    double tau[ORC_NDOF], qdd[ORC_NDOF];
    for (int d = 0; d < ORC_NDOF; d++) tau[d] = -sc->damping[sc->link_of_dof[d]] * e->qd[d];
    aba_internal(e, tau, qdd);
    for (int d = 0; d < ORC_NDOF; d++) e->qd[d] = clampd(e->qd[d] + qdd[d] * dt, -cf.max_coord_vel, cf.max_coord_vel);
    for (int o = 0; o < sc->n_objects; o++) {
        OrcObjState& b = e->obj[o];
        M3 R = quat_to_mat(b.quat), Rt = transpose(R);
        V3 wl = Rt * b.ang, vl = Rt * b.lin;            /* body-frame velocities */
        V3 I = sc->obj_inertia[o]; double m = sc->obj_mass[o];
        V3 Iw(I.x * wl.x, I.y * wl.y, I.z * wl.z);
        /* zero-acceleration force (same terms as the links); the m*(w x v) term cancels in the world-frame output */
        V3 pt = Iw * (cf.ang_damping + cf.ang_damping * norm(wl)) + cross(wl, Iw);
        V3 pf = -(Rt * V3(0, 0, m * cf.gravity)) + (vl * m) * (cf.lin_damping + cf.lin_damping * norm(vl));
        V3 al(-pt.x / I.x, -pt.y / I.y, -pt.z / I.z);
        V3 ll = pf * (-1.0 / m);
        V3 wdot = R * al, vdot = R * ll;
        for (int k = 0; k < 3; k++) {
            b.ang[k] = clampd(b.ang[k] + wdot[k] * dt, -cf.max_coord_vel, cf.max_coord_vel);
            b.lin[k] = clampd(b.lin[k] + vdot[k] * dt, -cf.max_coord_vel, cf.max_coord_vel);
        }
    }
    /* constraints */
    Solver S;
    build_rows(e, S);
    solve(e, S);
    for (int d = 0; d < ORC_NDOF; d++) e->qd[d] = clampd(e->qd[d] + S.dva[d], -cf.max_coord_vel, cf.max_coord_vel);
    for (int o = 0; o < sc->n_objects; o++) {
        OrcObjState& b = e->obj[o];
        for (int k = 0; k < 3; k++) {
            b.ang[k] = clampd(b.ang[k] + S.dvo[o][k], -cf.max_coord_vel, cf.max_coord_vel);
            b.lin[k] = clampd(b.lin[k] + S.dvo[o][3 + k], -cf.max_coord_vel, cf.max_coord_vel);
        }
    }
    /* touch sensors: max normal force over contacts on skin links whose other side is a base link
     * (robot.py:131-163; getContactPoints(body,-1,link,-1)); normal force = applied impulse / dt */
    static const int skin_link[4] = {11, 10, 15, 14};    /* skin_00, skin_01, skin_10, skin_11 */
    for (int s = 0; s < 4; s++) e->touch[s] = 0;
    for (int m = 0; m < e->n_manifolds; m++) {
        const OrcManifold& mf = e->manifolds[m];
        const OrcShape& sA = sc->shapes[mf.shapeA]; const OrcShape& sB = sc->shapes[mf.shapeB];
        if (sA.body != 0 || sB.body == 0 || sB.link != -1) continue;
        for (int s = 0; s < 4; s++) if (sA.link == skin_link[s]) {
            bool any = false; double best = 0;
            for (int k = 0; k < mf.n; k++) if (std::fabs(mf.pt[k].distance) < 0.1) {
                double f = mf.pt[k].applied_impulse / dt;
                if (!any || f > best) { best = f; any = true; }
            }
            if (any && best > e->touch[s]) e->touch[s] = best;
        }
    }
    /* integrate positions (semi-implicit Euler) */
    for (int d = 0; d < ORC_NDOF; d++) e->q[d] += e->qd[d] * dt;
    for (int o = 0; o < sc->n_objects; o++) {
        OrcObjState& b = e->obj[o];
        b.pos = b.pos + b.lin * dt;
        b.quat = integrate_quat(b.quat, b.ang, dt);
    }
    orc_fk(e);
}

void orc_calc_state(const OrcEnv* e, double* j) {
    for (int i = 0; i < 7; i++) j[i] = e->q[i];
    j[7] = e->q[7];          /* base_to_finger00_joint */
    j[8] = -e->q[8];         /* -finger00_to_finger01_joint */
}

void orc_env_step(OrcEnv* e, const double* cmd) {
    static const double maxDiff[9] = {0.2, 0.2, 0.2, 0.2, 0.2, 0.3, 0.3, 0.1, 0.1};
    const double PI = 3.141592653589793;
    double lo[9], hi[9];
    for (int i = 0; i < 9; i++) { lo[i] = -PI * 0.944; hi[i] = PI * 0.944; }
    lo[0] = -PI * 0.666; hi[0] = PI * 0.666;
    for (int i = 1; i < 9; i += 2) { lo[i] = -PI * 0.666; hi[i] = PI * 0.666; }
    lo[6] = -PI * 0.972; hi[6] = PI * 0.972;
    lo[7] = lo[8] = 0; hi[7] = hi[8] = PI / 2;
    double cur[9], a[9];
    orc_calc_state(e, cur);
    for (int i = 0; i < 9; i++) {            /* limitActionByJoint env.py:314-321 */
        double diff = std::min(maxDiff[i], cmd[i] - cur[i]);
        diff = std::max(-maxDiff[i], diff);
        a[i] = cur[i] + diff;
    }
    /* control_objects_limits env.py:257-264 (table never moves; objects teleport home, velocities zeroed) */
    for (int o = 0; o < e->sc->n_objects; o++) {
        V3 p = e->obj[o].pos;
        if (p.z < e->sc->table_pos.z || (p.x > 0.11 && p.z < 0.29)) {
            e->obj[o].pos = e->sc->obj_init_pos[o]; e->obj[o].quat = e->sc->obj_init_quat[o];
            e->obj[o].lin = V3(); e->obj[o].ang = V3();
        }
    }
    /* apply_action robot.py:188-201 */
    for (int i = 0; i < 9; i++) a[i] = std::max(lo[i], std::min(a[i], hi[i]));
    a[8] = std::max(0.0, std::min(2 * a[7], a[8]));
    for (int i = 0; i < 7; i++) e->target[i] = a[i];
    e->target[7] = a[7];      /* base_to_finger00 */
    e->target[8] = -a[8];     /* finger00_to_finger01 */
    e->target[9] = a[7];      /* base_to_finger10 */
    e->target[10] = -a[8];    /* finger10_to_finger11 */
    orc_step_simulation(e);
    e->timestep++;
}
