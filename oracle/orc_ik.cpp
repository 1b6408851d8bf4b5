/* TEST INFRASTRUCTURE — CPU oracle: inverse kinematics.
 *
 * Restates pybullet.calculateInverseKinematics(0, 7, pos, orn, maxNumIterations=1000, residualThreshold=0.001) as the
 * reference calls it (real_robots/envs/env.py:372-375 step_cartesian, :423-426 generate_plan/goToPosXY): damped
 * least squares on position + orientation (BussIK, IK2_VEL_DLS_WITH_ORIENTATION), seeded with the
 * env's current joint positions, joint damping 0.5 on every dof, step clamp 45 degrees, stop when the POSITION
 * residual is below the threshold.  pybullet is un-vendored and unpinned (setup.py:24-32): PARITY UNPINNED.
 * Deviation: the target quaternion is normalised first (the reference's cartesian space allows un-normalised input).
 */
#include "orc_ik.h"
#include <cmath>

static void fk_chain(const OrcScene* sc, const double* q, M3* Rw, V3* pw) {
    for (int i = 0; i <= 7; i++) {
        M3 Rpc = sc->jtype[i] == 1 ? axis_angle(sc->axis[i], -q[sc->dof[i]]) * sc->rot0[i] : sc->rot0[i];
        V3 r = Rpc * sc->dvec[i] + sc->evec[i];
        M3 Rp = M3::identity(); V3 pp = sc->base_pos;
        if (sc->parent[i] >= 0) { Rp = Rw[sc->parent[i]]; pp = pw[sc->parent[i]]; }
        Rw[i] = Rp * transpose(Rpc);
        pw[i] = pp + Rw[i] * r;
    }
}

int orc_ik_solve(OrcEnv* e, const double* pos3, const double* quat4, int max_iter, double residual, double* out11) {
    const OrcScene* sc = e->sc;
    double q[ORC_NDOF];
    for (int i = 0; i < ORC_NDOF; i++) q[i] = e->q[i];
    V3 target(pos3[0], pos3[1], pos3[2]);
    Q4 tq(quat4[0], quat4[1], quat4[2], quat4[3]);
    double qn = std::sqrt(tq.x * tq.x + tq.y * tq.y + tq.z * tq.z + tq.w * tq.w);
    if (qn > 0) tq = Q4(tq.x / qn, tq.y / qn, tq.z / qn, tq.w / qn); else tq = Q4(0, 0, 0, 1);
    const double damp = 0.5, max_angle = 45.0 * 3.14159265358979323846 / 180.0;
    int it = 0;
    for (it = 0; it < max_iter; it++) {
        M3 Rw[8]; V3 pw[8];
        fk_chain(sc, q, Rw, pw);
        V3 ee = pw[7];
        V3 dp = target - ee;
        if (norm(dp) <= residual) break;            /* loop runs while currentDiff > residualThreshold */
        /* orientation error: deltaQ = target * current^-1, angle-axis */
        Q4 cq = mat_to_quat(Rw[7]);
        Q4 dq = qmul(tq, Q4(-cq.x, -cq.y, -cq.z, cq.w));
        double w = dq.w > 1 ? 1 : (dq.w < -1 ? -1 : dq.w);
        double angle = 2.0 * std::acos(w);
        double s2 = 1.0 - dq.w * dq.w;
        V3 axis(1, 0, 0);
        if (s2 >= 10 * 2.220446049250313e-16) { double s = 1.0 / std::sqrt(s2); axis = V3(dq.x * s, dq.y * s, dq.z * s); }
        const double PI = 3.14159265358979323846;
        if (angle > PI) angle -= 2.0 * PI; else if (angle < -PI) angle += 2.0 * PI;
        V3 dr = normalized(axis) * angle;
        double dS[6] = {dp.x, dp.y, dp.z, dr.x, dr.y, dr.z};
        /* Jacobian (6 x 7): linear on top, angular below; finger dofs do not move link 7 */
        double J[6][7];
        for (int k = 0; k < 7; k++) {
            V3 aw = Rw[k] * sc->axis[k];
            V3 piv = pw[k] - Rw[k] * sc->evec[k];
            V3 lin = cross(aw, ee - piv);
            J[0][k] = lin.x; J[1][k] = lin.y; J[2][k] = lin.z; J[3][k] = aw.x; J[4][k] = aw.y; J[5][k] = aw.z;
        }
        /* (J^T J + diag(damp)) dTheta = J^T dS   — Cholesky on the 7x7 SPD system */
        double U[7][7], b[7];
        for (int i = 0; i < 7; i++) {
            for (int j = 0; j < 7; j++) { double s = 0; for (int r = 0; r < 6; r++) s += J[r][i] * J[r][j]; U[i][j] = s; }
            U[i][i] += damp;
            double s = 0; for (int r = 0; r < 6; r++) s += J[r][i] * dS[r]; b[i] = s;
        }
        double L[7][7] = {{0}};
        for (int i = 0; i < 7; i++) for (int j = 0; j <= i; j++) {
            double s = U[i][j];
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            L[i][j] = (i == j) ? std::sqrt(s) : s / L[j][j];
        }
        double y[7], dth[7];
        for (int i = 0; i < 7; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
        for (int i = 6; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < 7; k++) s -= L[k][i] * dth[k]; dth[i] = s / L[i][i]; }
        double mx = 0; for (int i = 0; i < 7; i++) mx = std::fmax(mx, std::fabs(dth[i]));
        if (mx > max_angle) for (int i = 0; i < 7; i++) dth[i] *= max_angle / mx;
        for (int i = 0; i < 7; i++) q[i] += dth[i];
    }
    for (int i = 0; i < ORC_NDOF; i++) out11[i] = q[i];
    return it;
}
