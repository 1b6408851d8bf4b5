// rr_physics.cu — batched articulated dynamics for the REALRobot2020 scene, sm_100a.
//
// Replaces pybullet's stepSimulation as called from the reference at real_robots/envs/env.py:340
// (scene.global_step()), together with the Python prologue of step_joints (env.py:314-321 limitActionByJoint,
// env.py:257-264 control_objects_limits, robot.py:188-201 apply_action) and the state read-out
// (robot.py:203-211 calc_state, robot.py:152-163 get_touch_sensors).
//
// Design: ONE WARP PER ENV.  All link state of the env lives in shared memory (struct WarpMem); lanes split the
// work inside each Featherstone step (the 36 entries of a 6x6 articulated inertia, the 6 components of a spatial
// vector, the 16 links for per-link independent work, the 11 unit-impulse responses).  fp64 throughout, as in
// pybullet's double-precision build.  Global state is structure-of-arrays [field][env] so the warps of a block
// touch consecutive addresses.  No tensor cores: nothing here is a dense contraction.
#include "rr_ctx.h"
#include <cstddef>

__constant__ RobotConst c_robot;

#define WPB 4                 // envs (warps) per block
#define FULL 0xffffffffu

// Kinematics workspace (both kernels)
struct KinWs {
    double Rpc[RR_NL][9];     // parent -> link rotation
    double rvec[RR_NL][3];    // parent COM -> link COM, link frame
    double Rw[RR_NL][9];      // world_from_link
    double pw[RR_NL][3];      // link COM, world
};
#define RR_SLOTS_A RR_ROWS_A
#define RR_SLOTS_B RR_MAX_CONTACTS
// What k_step_dyn hands to k_step_solve through DevState::handoff — the solver IMAGE of an env: a fixed head (state after
// the unconstrained update), the SLOT SCHEDULE of one Gauss-Seidel iteration, and the constraint rows as compact BLOCKS.
//
// The lanes of a warp form four groups — the arm dofs (lanes 0..10) and the three free bodies (lanes 11+6o .. 16+6o) — and a
// constraint row only touches the groups of the bodies it couples, so rows of disjoint groups commute.  A row is stored
// once per group it touches, as a block: (J, M^-1 J^T) per lane of the group, then rhs, 1/d, d, lo-or-coefficient, hi, applied
// impulse.  The schedule says, per slot and group, which block that group resolves (in the order of the sequential sweep
// restricted to the group; the joint rows run backwards on even iterations, hence two schedules for the single rows); the
// groups of one slot execute the same instructions at the same time.  A row that needs a reduction wider than one group's
// own shuffles (an arm contact row, or a row coupling two groups) is the slot's single `wide` row.  Friction pairs (the
// implicit cone) have their own slots after the single rows.  The solver keeps the image in shared memory (the part that
// does not fit stays in global memory and is addressed there).
#define RR_ARM_BLK 28            // doubles: 11 x (J, rr), rhs, inv, rinv, lo | coef, hi, app
#define RR_OBJ_BLK 18            // doubles:  6 x (J, rr), rhs, inv, rinv, lo | coef, hi, app
#define ENT_IDLE 0xffffffffu     // schedule entry, word 0: bits 0..14 block offset (16-byte units, relative to the first block), bit 15 wide,
#define ENT_NORMAL 0xfffeu       // bits 16..31: 0xfffe normal row, 0xffe0 | dof joint row, else (0) a row whose limits follow its normal row's impulse;
#define ENT_JOINT 0xffe0u        // word 1: row id | row id of the contact's normal row << 16 (the applied impulses live in shared memory by row id)
struct SolveHead {
    double q[RR_NDOF], qd[RR_NDOF], target[RR_NDOF];
    double objreg[RR_MAX_OBJECTS][13];                  // free bodies after the unconstrained update (pos3 quat4 lin3 ang3)
    int n_slotsA[2], n_slotsB;
    int n_nc, nc, overflow, fault, simple, n_rows_total, n_blk_doubles, pad_[2];   // fault bit 0: a non-finite joint command arrived this step
    unsigned short nrm_rel[RR_MAX_CONTACTS];            // per contact: its normal row's first block (bit 15: an arm block), for the impulse write-back
};
#define HEAD_DOUBLES ((int)(sizeof(SolveHead) / 8))
#define GSCHED_A(par) (8 * RR_SLOTS_A * (par))                          // word offsets of the schedules in the image (fixed places; 2 words per slot and group)
#define GSCHED_B (16 * RR_SLOTS_A)
#define GBLK_BASE (HEAD_DOUBLES + (16 * RR_SLOTS_A + 8 * RR_SLOTS_B) / 2)   // first block, in doubles
static_assert(sizeof(SolveHead) % 16 == 0, "the schedule words and blocks must stay 16-byte aligned");
static_assert(GBLK_BASE % 2 == 0 && GBLK_BASE + RR_NC_ROWS * RR_ARM_BLK + RR_MAX_CONTACTS * 6 * (RR_ARM_BLK + RR_OBJ_BLK) <= RR_IMG_DOUBLES, "solver image too small");
#ifndef SOLVE_HEAVY_ABOVE
#define SOLVE_HEAVY_ABOVE 24
#endif
// cost class of a schedule (slots per iteration, a pair slot counted twice; an env whose contacts are all object-vs-static: 16 + 2 * 4 =
// 24): the classes above the last one are "heavy" (k_step_solve_heavy takes the first HEAVY_GRID of them, longest first)
__device__ __forceinline__ int solve_cost_class(int cost) {
    return cost > 64 ? 0 : cost > 56 ? 1 : cost > 48 ? 2 : cost > 40 ? 3 : cost > 32 ? 4 : cost > 28 ? 5 : cost > SOLVE_HEAVY_ABOVE ? 6 : 7;
}
// k_step_dyn's own workspace
// 14 144 B per env: four blocks of four envs per SM.  The forward-dynamics workspace (ABA: dead once aba_minv has the M^-1 columns) and
// the workspace of the constraint rows (first written after that) share their bytes.
struct WarpMem {
    double q[RR_NDOF], qd[RR_NDOF], target[RR_NDOF];
    double Minv[RR_NDOF][RR_NDOF];
    int fault;
    KinWs k;
    union {
        struct {      // forward dynamics
            double v[RR_NL][6];       // spatial velocity (ang, lin) in link frame
            double c[RR_NL][6];       // Coriolis acceleration
            double p[RR_NL][6];       // zero-acceleration force
            double IA[RR_NL][36];     // articulated inertia
            double h[RR_NL][6];
            double invD[RR_NL];
            double Y[RR_NL];
            double acc[RR_NL][6];
            double X[36], T[36], Ia[36], pa[6], Ic[6];
            double tau[RR_NDOF], qdd[RR_NDOF];
        };
        struct {      // constraint rows
            // non-contact rows (limits then motors) while they are being set up
            double r_rhs[RR_NC_ROWS], r_inv[RR_NC_ROWS], r_rinv[RR_NC_ROWS], r_lo[RR_NC_ROWS], r_hi[RR_NC_ROWS], r_sign[RR_NC_ROWS];
            int r_dof[RR_NC_ROWS];
            int n_rows;
            unsigned char c_grp[RR_MAX_CONTACTS], c_flags[RR_MAX_CONTACTS];   // per contact: group mask (bit 0 arm, bit 1+o object o); bit 0 spinning row, bit 1 rolling rows, bit 2 wide
            unsigned short c_base[RR_MAX_CONTACTS];                           // first block of the contact (16-byte units)
            double dof_axis[RR_NDOF][3], dof_piv[RR_NDOF][3];   // world joint axis / pivot per dof
            double opos[RR_MAX_OBJECTS][3], ovel[RR_MAX_OBJECTS][6], oIinv[RR_MAX_OBJECTS][9], oinvm[RR_MAX_OBJECTS];
            double Jtmp[32];
        };
    };
};
static_assert(sizeof(WarpMem) <= 14336, "k_step_dyn: four blocks of WPB = 4 envs per SM need <= 56 KB of shared memory per block");
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}

// two warp reductions for the price of one: after the first exchange the lower half-warp carries a, the upper b
__device__ __forceinline__ void warp_sum2(double a, double b, double& sa, double& sb) {
    const bool hi = (threadIdx.x & 16) != 0;
    double keep = hi ? b : a, send = hi ? a : b;
    keep += __shfl_xor_sync(0xffffffffu, send, 16);
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, off);
    sa = __shfl_sync(0xffffffffu, keep, 0); sb = __shfl_sync(0xffffffffu, keep, 16);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// ------------------------------------------------------------------------------------------- forward kinematics
// lane i < 16 computes its link's joint transform; then a serial sweep down the tree builds world frames.
__device__ __forceinline__ void fk_local(const double* q, KinWs& m, int lane) {
    if (lane < RR_NL) {
        const int i = lane;
        double R0[9];
#pragma unroll
        for (int k = 0; k < 9; k++) R0[k] = c_robot.rot0[i][k];
        double R[9];
        if (c_robot.jtype[i] == 1) {
            double th = -q[c_robot.dof[i]];
            double s, c; sincos(th, &s, &c);
            double ax = c_robot.axis[i][0], ay = c_robot.axis[i][1], az = c_robot.axis[i][2];
            double t = 1.0 - c;
            double A[9] = {c + ax * ax * t, ax * ay * t - az * s, ax * az * t + ay * s,
                           ay * ax * t + az * s, c + ay * ay * t, ay * az * t - ax * s,
                           az * ax * t - ay * s, az * ay * t + ax * s, c + az * az * t};
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) R[3 * a + b] = A[3 * a] * R0[b] + A[3 * a + 1] * R0[3 + b] + A[3 * a + 2] * R0[6 + b];
        } else {
#pragma unroll
            for (int k = 0; k < 9; k++) R[k] = R0[k];
        }
#pragma unroll
        for (int k = 0; k < 9; k++) m.Rpc[i][k] = R[k];
#pragma unroll
        for (int a = 0; a < 3; a++)
            m.rvec[i][a] = R[3 * a] * c_robot.dvec[i][0] + R[3 * a + 1] * c_robot.dvec[i][1] + R[3 * a + 2] * c_robot.dvec[i][2] + c_robot.evec[i][a];
    }
    __syncwarp();
}

// serial sweep: world frames (lanes 0..8 rotation entries, then lanes 0..2 position) and, if vel is given,
// spatial velocities (lanes 16..21)
__device__ __forceinline__ void fk_sweep(KinWs& m, int lane, double (*vel)[6], const double* qd) {
    const bool with_vel = vel != nullptr;
    for (int i = 0; i < RR_NL; i++) {
        const int par = c_robot.parent[i];
        if (lane < 9) {
            int a = lane / 3, b = lane % 3;
            double r;
            if (par >= 0) r = m.Rw[par][3 * a] * m.Rpc[i][3 * b] + m.Rw[par][3 * a + 1] * m.Rpc[i][3 * b + 1] + m.Rw[par][3 * a + 2] * m.Rpc[i][3 * b + 2];
            else r = m.Rpc[i][3 * b + a];       // identity base: Rw = Rpc^T
            m.Rw[i][lane] = r;
        } else if (with_vel && lane >= 16 && lane < 22) {
            int k = lane - 16;
            double wp[3] = {0, 0, 0}, vp[3] = {0, 0, 0};
            if (par >= 0) {
#pragma unroll
                for (int a = 0; a < 3; a++) { wp[a] = vel[par][a]; vp[a] = vel[par][3 + a]; }
            }
            double w[3];
#pragma unroll
            for (int a = 0; a < 3; a++) w[a] = m.Rpc[i][3 * a] * wp[0] + m.Rpc[i][3 * a + 1] * wp[1] + m.Rpc[i][3 * a + 2] * wp[2];
            double qdj = (c_robot.jtype[i] == 1) ? qd[c_robot.dof[i]] : 0.0;
            double out;
            if (k < 3) out = w[k] + c_robot.S[i][k] * qdj;
            else {
                int a = k - 3;
                double Rv = m.Rpc[i][3 * a] * vp[0] + m.Rpc[i][3 * a + 1] * vp[1] + m.Rpc[i][3 * a + 2] * vp[2];
                const double* r = m.rvec[i];
                double wxr = (a == 0) ? (w[1] * r[2] - w[2] * r[1]) : (a == 1) ? (w[2] * r[0] - w[0] * r[2]) : (w[0] * r[1] - w[1] * r[0]);
                out = Rv + wxr + c_robot.S[i][k] * qdj;
            }
            vel[i][k] = out;
        }
        __syncwarp();
        if (lane < 3) {
            double pp = (par >= 0) ? m.pw[par][lane] : c_robot.base_pos[lane];
            m.pw[i][lane] = pp + m.Rw[i][3 * lane] * m.rvec[i][0] + m.Rw[i][3 * lane + 1] * m.rvec[i][1] + m.Rw[i][3 * lane + 2] * m.rvec[i][2];
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------- ABA
__device__ __forceinline__ double dot6(const double* a, const double* b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
}

// per-link bias terms (lane = link): Coriolis acceleration c, zero-acceleration force p, rigid-body inertia
__device__ __forceinline__ void aba_bias(WarpMem& m, int lane, const rr_config& cf) {
    if (lane < RR_NL) {
        const int i = lane;
        double w[3] = {m.v[i][0], m.v[i][1], m.v[i][2]}, l[3] = {m.v[i][3], m.v[i][4], m.v[i][5]};
        double qdj = (c_robot.jtype[i] == 1) ? m.qd[c_robot.dof[i]] : 0.0;
        double ja[3] = {c_robot.S[i][0] * qdj, c_robot.S[i][1] * qdj, c_robot.S[i][2] * qdj};
        double jl[3] = {c_robot.S[i][3] * qdj, c_robot.S[i][4] * qdj, c_robot.S[i][5] * qdj};
        // c = v x_motion vJ : top = w x ja ; bottom = l x ja + w x jl
        m.c[i][0] = w[1] * ja[2] - w[2] * ja[1];
        m.c[i][1] = w[2] * ja[0] - w[0] * ja[2];
        m.c[i][2] = w[0] * ja[1] - w[1] * ja[0];
        m.c[i][3] = (l[1] * ja[2] - l[2] * ja[1]) + (w[1] * jl[2] - w[2] * jl[1]);
        m.c[i][4] = (l[2] * ja[0] - l[0] * ja[2]) + (w[2] * jl[0] - w[0] * jl[2]);
        m.c[i][5] = (l[0] * ja[1] - l[1] * ja[0]) + (w[0] * jl[1] - w[1] * jl[0]);
        double I[3] = {c_robot.inertia[i][0], c_robot.inertia[i][1], c_robot.inertia[i][2]};
        double mass = c_robot.mass[i];
        double Iw[3] = {I[0] * w[0], I[1] * w[1], I[2] * w[2]};
        double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        double ln = sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
        double ka = cf.ang_damping + cf.ang_damping * wn, kl = cf.lin_damping + cf.lin_damping * ln;
        double fz = mass * cf.gravity;    // world force (0,0,fz); in link frame: R^T * f = column... Rw^T row 2
        double pa0 = Iw[0] * ka + (w[1] * Iw[2] - w[2] * Iw[1]);
        double pa1 = Iw[1] * ka + (w[2] * Iw[0] - w[0] * Iw[2]);
        double pa2 = Iw[2] * ka + (w[0] * Iw[1] - w[1] * Iw[0]);
        double pl0 = -(m.k.Rw[i][6] * fz) + l[0] * mass * kl + (w[1] * l[2] - w[2] * l[1]) * mass;
        double pl1 = -(m.k.Rw[i][7] * fz) + l[1] * mass * kl + (w[2] * l[0] - w[0] * l[2]) * mass;
        double pl2 = -(m.k.Rw[i][8] * fz) + l[2] * mass * kl + (w[0] * l[1] - w[1] * l[0]) * mass;
        m.p[i][0] = pa0; m.p[i][1] = pa1; m.p[i][2] = pa2; m.p[i][3] = pl0; m.p[i][4] = pl1; m.p[i][5] = pl2;
    }
    for (int e = lane; e < RR_NL * 36; e += 32) {
        int i = e / 36, k = e % 36, r = k / 6, cc = k % 6;
        double val = 0.0;
        if (r == cc) val = (r < 3) ? c_robot.inertia[i][r] : c_robot.mass[i];
        m.IA[i][k] = val;
    }
    __syncwarp();
}

// dense motion transform X (parent -> link): [[R,0],[-[r]x R, R]]
__device__ __forceinline__ double X_entry(const double* R, const double* r, int a, int b) {
    if (a < 3) return (b < 3) ? R[3 * a + b] : 0.0;
    int i = a - 3;
    if (b >= 3) return R[3 * i + (b - 3)];
    // -(skew(r) R)[i][b] ; skew(r) row i = (i==0: 0,-rz,ry) (1: rz,0,-rx) (2: -ry,rx,0)
    double k;
    if (i == 0) k = -r[2] * R[3 + b] + r[1] * R[6 + b];
    else if (i == 1) k = r[2] * R[b] - r[0] * R[6 + b];
    else k = -r[1] * R[b] + r[0] * R[3 + b];
    return -k;
}

__device__ __forceinline__ void aba_backward(WarpMem& m, int lane) {
    for (int i = RR_NL - 1; i >= 0; i--) {
        const bool rev = c_robot.jtype[i] == 1;
        const int par = c_robot.parent[i];
        // Ic = IA*c (lanes 0..5), h = IA*S (lanes 8..13), X entries (all lanes)
        if (lane < 6) m.Ic[lane] = dot6(&m.IA[i][6 * lane], m.c[i]);
        else if (lane >= 8 && lane < 14) { int k = lane - 8; m.h[i][k] = rev ? dot6(&m.IA[i][6 * k], c_robot.S[i]) : 0.0; }
        if (par >= 0) for (int e = lane; e < 36; e += 32) m.X[e] = X_entry(m.k.Rpc[i], m.k.rvec[i], e / 6, e % 6);
        __syncwarp();
        double invD = 0.0, Y = 0.0;
        if (rev) {
            double D = dot6(c_robot.S[i], m.h[i]);
            invD = 1.0 / D;
            Y = m.tau[c_robot.dof[i]] - dot6(c_robot.S[i], m.p[i]) - dot6(m.c[i], m.h[i]);
        }
        if (lane == 0) { m.invD[i] = invD; m.Y[i] = Y; }
        if (par >= 0) {
            for (int e = lane; e < 36; e += 32) { int a = e / 6, b = e % 6; m.Ia[e] = m.IA[i][e] - m.h[i][a] * invD * m.h[i][b]; }
            if (lane < 6) m.pa[lane] = m.p[i][lane] + m.Ic[lane] + m.h[i][lane] * (invD * Y);
            __syncwarp();
            // T = Ia * X
            for (int e = lane; e < 36; e += 32) {
                int a = e / 6, b = e % 6; double s = 0.0;
#pragma unroll
                for (int k = 0; k < 6; k++) s += m.Ia[6 * a + k] * m.X[6 * k + b];
                m.T[e] = s;
            }
            __syncwarp();
            // IA[par] += X^T T ; p[par] += X^T pa
            for (int e = lane; e < 36; e += 32) {
                int a = e / 6, b = e % 6; double s = 0.0;
#pragma unroll
                for (int k = 0; k < 6; k++) s += m.X[6 * k + a] * m.T[6 * k + b];
                m.IA[par][e] += s;
            }
            if (lane < 6) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 6; k++) s += m.X[6 * k + lane] * m.pa[k];
                m.p[par][lane] += s;
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void aba_forward(WarpMem& m, int lane) {
    for (int i = 0; i < RR_NL; i++) {
        const int par = c_robot.parent[i];
        const bool rev = c_robot.jtype[i] == 1;
        double a6[6];
        // every lane computes the transformed parent acceleration redundantly (6x6 on 6 values is cheap)
        if (par >= 0) {
            const double* R = m.k.Rpc[i]; const double* r = m.k.rvec[i]; const double* ap = m.acc[par];
            double w[3];
#pragma unroll
            for (int a = 0; a < 3; a++) w[a] = R[3 * a] * ap[0] + R[3 * a + 1] * ap[1] + R[3 * a + 2] * ap[2];
            a6[0] = w[0]; a6[1] = w[1]; a6[2] = w[2];
            a6[3] = R[0] * ap[3] + R[1] * ap[4] + R[2] * ap[5] + (w[1] * r[2] - w[2] * r[1]);
            a6[4] = R[3] * ap[3] + R[4] * ap[4] + R[5] * ap[5] + (w[2] * r[0] - w[0] * r[2]);
            a6[5] = R[6] * ap[3] + R[7] * ap[4] + R[8] * ap[5] + (w[0] * r[1] - w[1] * r[0]);
        } else {
#pragma unroll
            for (int k = 0; k < 6; k++) a6[k] = 0.0;
        }
        double qa = 0.0;
        if (rev) qa = m.invD[i] * (m.Y[i] - dot6(m.h[i], a6));
        __syncwarp();
        if (lane < 6) m.acc[i][lane] = a6[lane] + m.c[i][lane] + (rev ? c_robot.S[i][lane] * qa : 0.0);
        if (rev && lane == 6) m.qdd[c_robot.dof[i]] = qa;
        __syncwarp();
    }
}

// M^-1 columns: lane j < 11 runs the articulated-body delta pass (calcAccelerationDeltasMultiDof) for a unit
// generalised force on dof j, using the cached h / invD / transforms.  Private accumulators stay in local memory.
__device__ __forceinline__ void aba_minv(WarpMem& m, int lane) {
    if (lane < RR_NDOF) {
        double pz[RR_NL][6];
        double Yd[RR_NL];
#pragma unroll 1
        for (int i = 0; i < RR_NL; i++) {
#pragma unroll
            for (int k = 0; k < 6; k++) pz[i][k] = 0.0;
        }
#pragma unroll 1
        for (int i = RR_NL - 1; i >= 0; i--) {
            double pa[6];
#pragma unroll
            for (int k = 0; k < 6; k++) pa[k] = pz[i][k];
            if (c_robot.jtype[i] == 1) {
                double f = (c_robot.dof[i] == lane) ? 1.0 : 0.0;
                double y = f - dot6(c_robot.S[i], pz[i]);
                Yd[i] = y;
                double s = m.invD[i] * y;
#pragma unroll
                for (int k = 0; k < 6; k++) pa[k] += m.h[i][k] * s;
            } else Yd[i] = 0.0;
            int par = c_robot.parent[i];
            if (par >= 0) {
                // X^T force transform: f_p = R^T f ; t_p = R^T (t + r x f)
                const double* R = m.k.Rpc[i]; const double* r = m.k.rvec[i];
                double t[3] = {pa[0] + (r[1] * pa[5] - r[2] * pa[4]), pa[1] + (r[2] * pa[3] - r[0] * pa[5]), pa[2] + (r[0] * pa[4] - r[1] * pa[3])};
#pragma unroll
                for (int a = 0; a < 3; a++) {
                    pz[par][a] += R[a] * t[0] + R[3 + a] * t[1] + R[6 + a] * t[2];
                    pz[par][3 + a] += R[a] * pa[3] + R[3 + a] * pa[4] + R[6 + a] * pa[5];
                }
            }
        }
        // forward: reuse pz as the acceleration store
#pragma unroll 1
        for (int i = 0; i < RR_NL; i++) {
            int par = c_robot.parent[i];
            double a6[6];
            if (par >= 0) {
                const double* R = m.k.Rpc[i]; const double* r = m.k.rvec[i]; const double* ap = pz[par];
                double w[3];
#pragma unroll
                for (int a = 0; a < 3; a++) w[a] = R[3 * a] * ap[0] + R[3 * a + 1] * ap[1] + R[3 * a + 2] * ap[2];
                a6[0] = w[0]; a6[1] = w[1]; a6[2] = w[2];
                a6[3] = R[0] * ap[3] + R[1] * ap[4] + R[2] * ap[5] + (w[1] * r[2] - w[2] * r[1]);
                a6[4] = R[3] * ap[3] + R[4] * ap[4] + R[5] * ap[5] + (w[2] * r[0] - w[0] * r[2]);
                a6[5] = R[6] * ap[3] + R[7] * ap[4] + R[8] * ap[5] + (w[0] * r[1] - w[1] * r[0]);
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) a6[k] = 0.0;
            }
            if (c_robot.jtype[i] == 1) {
                double qa = m.invD[i] * (Yd[i] - dot6(m.h[i], a6));
                m.Minv[c_robot.dof[i]][lane] = qa;
#pragma unroll
                for (int k = 0; k < 6; k++) a6[k] += c_robot.S[i][k] * qa;
            }
#pragma unroll
            for (int k = 0; k < 6; k++) pz[i][k] = a6[k];
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------- free rigid bodies
struct ObjReg { double pos[3], quat[4], lin[3], ang[3]; };

__device__ __forceinline__ void quat_to_mat(const double* q, double* R) {
    double d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    double s = 2.0 / d;
    double xs = q[0] * s, ys = q[1] * s, zs = q[2] * s;
    double wx = q[3] * xs, wy = q[3] * ys, wz = q[3] * zs;
    double xx = q[0] * xs, xy = q[0] * ys, xz = q[0] * zs;
    double yy = q[1] * ys, yz = q[1] * zs, zz = q[2] * zs;
    R[0] = 1.0 - (yy + zz); R[1] = xy - wz; R[2] = xz + wy;
    R[3] = xy + wz; R[4] = 1.0 - (xx + zz); R[5] = yz - wx;
    R[6] = xz - wy; R[7] = yz + wx; R[8] = 1.0 - (xx + yy);
}

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

// unconstrained velocity update of a free body (base-only btMultiBody ABA: gravity, k(1+|v|) damping, gyroscopic term)
__device__ __forceinline__ void obj_free_update(ObjReg& b, int o, const rr_config& cf) {
    double R[9]; quat_to_mat(b.quat, R);
    double wl[3], vl[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
        wl[a] = R[a] * b.ang[0] + R[3 + a] * b.ang[1] + R[6 + a] * b.ang[2];
        vl[a] = R[a] * b.lin[0] + R[3 + a] * b.lin[1] + R[6 + a] * b.lin[2];
    }
    double I[3] = {c_robot.obj_inertia[o][0], c_robot.obj_inertia[o][1], c_robot.obj_inertia[o][2]};
    double mass = c_robot.obj_mass[o];
    double Iw[3] = {I[0] * wl[0], I[1] * wl[1], I[2] * wl[2]};
    double wn = sqrt(wl[0] * wl[0] + wl[1] * wl[1] + wl[2] * wl[2]);
    double ln = sqrt(vl[0] * vl[0] + vl[1] * vl[1] + vl[2] * vl[2]);
    double ka = cf.ang_damping + cf.ang_damping * wn, kl = cf.lin_damping + cf.lin_damping * ln;
    double fz = mass * cf.gravity;
    double pt[3] = {Iw[0] * ka + (wl[1] * Iw[2] - wl[2] * Iw[1]), Iw[1] * ka + (wl[2] * Iw[0] - wl[0] * Iw[2]), Iw[2] * ka + (wl[0] * Iw[1] - wl[1] * Iw[0])};
    double pf[3] = {-(R[6] * fz) + vl[0] * mass * kl, -(R[7] * fz) + vl[1] * mass * kl, -(R[8] * fz) + vl[2] * mass * kl};
    double al[3] = {-pt[0] / I[0], -pt[1] / I[1], -pt[2] / I[2]};
    double ll[3] = {pf[0] * (-1.0 / mass), pf[1] * (-1.0 / mass), pf[2] * (-1.0 / mass)};
#pragma unroll
    for (int a = 0; a < 3; a++) {
        double wdot = R[3 * a] * al[0] + R[3 * a + 1] * al[1] + R[3 * a + 2] * al[2];
        double vdot = R[3 * a] * ll[0] + R[3 * a + 1] * ll[1] + R[3 * a + 2] * ll[2];
        b.ang[a] = clampd(b.ang[a] + wdot * cf.dt, -cf.max_coord_vel, cf.max_coord_vel);
        b.lin[a] = clampd(b.lin[a] + vdot * cf.dt, -cf.max_coord_vel, cf.max_coord_vel);
    }
}

__device__ __forceinline__ void obj_integrate(ObjReg& b, double dt) {
#pragma unroll
    for (int a = 0; a < 3; a++) b.pos[a] += b.lin[a] * dt;
    double w[3] = {b.ang[0], b.ang[1], b.ang[2]};
    double ang = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    if (ang * dt > 0.7853981633974483) ang = 0.5 * 1.5707963267948966 / dt;
    double k;
    if (ang < 0.001) k = 0.5 * dt - (dt * dt * dt) * 0.020833333333 * ang * ang;
    else k = sin(0.5 * ang * dt) / ang;
    double dq[4] = {w[0] * k, w[1] * k, w[2] * k, cos(ang * dt * 0.5)};
    const double* q = b.quat;
    double r[4] = {dq[3] * q[0] + dq[0] * q[3] + dq[1] * q[2] - dq[2] * q[1],
                   dq[3] * q[1] + dq[1] * q[3] + dq[2] * q[0] - dq[0] * q[2],
                   dq[3] * q[2] + dq[2] * q[3] + dq[0] * q[1] - dq[1] * q[0],
                   dq[3] * q[3] - dq[0] * q[0] - dq[1] * q[1] - dq[2] * q[2]};
    double n = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]);
#pragma unroll
    for (int a = 0; a < 4; a++) b.quat[a] = r[a] / n;
}

// ------------------------------------------------------------------------------------------- step kernel
__device__ __forceinline__ void store_frames(const KinWs& m, const DevState& st, int e, int lane) {
    for (int k = lane; k < RR_NL * 12; k += 32) {
        int i = k / 12, j = k % 12;
        st.link_frames[(size_t)k * st.N + e] = (j < 9) ? m.Rw[i][j] : m.pw[i][j - 9];
    }
}

// lane -> group: 0 = arm dofs (lanes 0..10), 1 + o = free body o (lanes 11 + 6 o .. 16 + 6 o), 4 = idle lanes 29..31
__device__ __forceinline__ int lane_group(int lane) { return lane < RR_NDOF ? 0 : (lane < 29 ? 1 + (lane - RR_NDOF) / 6 : 4); }

// lane 0: place one row (or friction pair) of group mask gm into the earliest slot all its groups are free in; a wide row also
// needs a slot without another wide row.  ent[g] = word 0 of group g's schedule entry for this row, w1 = word 1.
__device__ __forceinline__ void sched_place(unsigned int* sched, int* cur, unsigned int* wide_used, int gm, bool is_wide, const unsigned int* ent, unsigned int w1) {
    int slot = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) if ((gm >> g) & 1) slot = max(slot, cur[g]);
    if (is_wide) { while ((wide_used[slot >> 5] >> (slot & 31)) & 1u) slot++; wide_used[slot >> 5] |= 1u << (slot & 31); }
#pragma unroll
    for (int g = 0; g < 4; g++) if ((gm >> g) & 1) { sched[8 * slot + 2 * g] = ent[g] | (is_wide ? 0x8000u : 0u); sched[8 * slot + 2 * g + 1] = w1; cur[g] = slot + 1; }
}
// word 0 of the entries of contact row `rowidx` (0 normal, 1 2 friction, 3.. spinning / rolling) for the groups of gm; blocks of
// one contact are laid out group by group, `nr` rows each
__device__ __forceinline__ void contact_entries(unsigned int* ent, int gm, int base16, int nr, int rowidx) {
    int off = base16;        // 16-byte units
#pragma unroll
    for (int g = 0; g < 4; g++) if ((gm >> g) & 1) {
        const int sz = (g == 0 ? RR_ARM_BLK : RR_OBJ_BLK) / 2;
        ent[g] = (unsigned int)(off + rowidx * sz) | ((rowidx == 0 ? ENT_NORMAL : 0u) << 16);
        off += nr * sz;
    }
}

// One env step is two kernels.  k_step_dyn: action prologue, kinematics, ABA, M^-1, constraint rows -> the global row scratch,
// and the slot schedule of the solver sweep -> DevState::handoff.  k_step_solve: PGS, integration, final kinematics,
// observations.  The solver needs a third of the dynamics' shared memory, so on its own it runs at more than twice the occupancy.
__global__ void __launch_bounds__(WPB * 32) k_step_dyn(DevState st, const double* __restrict__ cmd, int* __restrict__ man_hdr, double* __restrict__ man_pts, rr_config cf, int n_objects) {
    RR_DYN_SMEM(smem_raw);
    WarpMem& m = reinterpret_cast<WarpMem*>(smem_raw)[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * WPB + (threadIdx.x >> 5);
    if (e >= st.N) return;
    const int N = st.N;
    const double dt = cf.dt;

    if (lane < RR_NDOF) { m.q[lane] = st.q[(size_t)lane * N + e]; m.qd[lane] = st.qd[(size_t)lane * N + e]; }
    __syncwarp();
    // ---- action prologue (env.py:314-321, robot.py:188-201): lanes 0..8
    {
        double a = 0.0;
        bool badcmd = false;
        if (lane < RR_NUM_JOINTS) {
            const double PI = 3.141592653589793;
            double cur = (lane < 8) ? m.q[lane] : -m.q[8];
            double md = (lane < 5) ? 0.2 : (lane < 7 ? 0.3 : 0.1);
            const double want = cmd[(size_t)e * RR_NUM_JOINTS + lane];
            // the reference asserts np.isfinite(a).all() (robot.py:189); a batched step cannot raise per env: fmin / fmax turn a NaN
            // or infinite command into a full +-max step (never into a NaN target) and the env is flagged in rr_diag
            badcmd = !isfinite(want);
            double diff = fmin(md, want - cur);
            diff = fmax(-md, diff);
            a = cur + diff;
            double lo = -PI * 0.944, hi = PI * 0.944;
            if (lane == 0 || (lane & 1)) { lo = -PI * 0.666; hi = PI * 0.666; }
            if (lane == 6) { lo = -PI * 0.972; hi = PI * 0.972; }
            if (lane >= 7) { lo = 0.0; hi = PI / 2; }
            a = fmax(lo, fmin(a, hi));
        }
        const bool anybad = __any_sync(FULL, badcmd);
        if (lane == 0) m.fault = anybad ? 1 : 0;
        double a7 = __shfl_sync(FULL, a, 7);
        if (lane == 8) a = fmax(0.0, fmin(2 * a7, a));
        double a8 = __shfl_sync(FULL, a, 8);
        if (lane < 7) m.target[lane] = a;
        if (lane == 7) { m.target[7] = a7; m.target[9] = a7; }
        if (lane == 8) { m.target[8] = -a8; m.target[10] = -a8; }
    }
    // ---- objects: lanes 0..n_objects-1 each own one free body (registers)
    ObjReg ob;
    const bool has_obj = lane < n_objects;
    if (has_obj) {
        const double* base = st.obj + (size_t)lane * 13 * N + e;
#pragma unroll
        for (int k = 0; k < 3; k++) ob.pos[k] = base[(size_t)k * N];
#pragma unroll
        for (int k = 0; k < 4; k++) ob.quat[k] = base[(size_t)(3 + k) * N];
#pragma unroll
        for (int k = 0; k < 3; k++) { ob.lin[k] = base[(size_t)(7 + k) * N]; ob.ang[k] = base[(size_t)(10 + k) * N]; }
        // control_objects_limits (env.py:257-264)
        if (ob.pos[2] < c_robot.table_pos[2] || (ob.pos[0] > 0.11 && ob.pos[2] < 0.29)) {
#pragma unroll
            for (int k = 0; k < 3; k++) { ob.pos[k] = c_robot.obj_init[lane][k]; ob.lin[k] = 0.0; ob.ang[k] = 0.0; }
#pragma unroll
            for (int k = 0; k < 4; k++) ob.quat[k] = c_robot.obj_init[lane][3 + k];
        }
    }
    __syncwarp();

    // ---- forward dynamics
    fk_local(m.q, m.k, lane);
    fk_sweep(m.k, lane, m.v, m.qd);
    if (lane < RR_NDOF) m.tau[lane] = -c_robot.damping[c_robot.link_of_dof[lane]] * m.qd[lane];
    aba_bias(m, lane, cf);
    aba_backward(m, lane);
    aba_forward(m, lane);
    if (lane < RR_NDOF) m.qd[lane] = clampd(m.qd[lane] + m.qdd[lane] * dt, -cf.max_coord_vel, cf.max_coord_vel);
    if (has_obj) obj_free_update(ob, lane, cf);
    aba_minv(m, lane);

    double* img = st.handoff + (size_t)e * RR_IMG_DOUBLES;                      // the solver image of this env
    SolveHead* hand = reinterpret_cast<SolveHead*>(img);
    unsigned int* gsched = reinterpret_cast<unsigned int*>(img + HEAD_DOUBLES);
    double* gblk = img + GBLK_BASE;
    const int mygrp = lane_group(lane);
    const int seg0 = mygrp == 0 ? 0 : (mygrp < 4 ? 11 + 6 * (mygrp - 1) : 31);
    const int myk2 = 2 * (lane - seg0);                                          // offset of this lane's (J, rr) pair inside a block of its group
    const int mysc = mygrp == 0 ? 2 * RR_NDOF : 12;                              // offset of the scalars
    const bool leader = mygrp < 4 && lane == seg0;
    // ---- constraint rows: joint limits (link order) then motors (link order)
    if (lane == 0) {
        int n = 0;
        for (int i = 0; i < RR_NL; i++) {
            if (c_robot.jtype[i] != 1 || c_robot.lower[i] > c_robot.upper[i]) continue;
            int d = c_robot.dof[i];
            double pos[2] = {m.q[d] - c_robot.lower[i], c_robot.upper[i] - m.q[d]};
            for (int row = 0; row < 2; row++) {
                double pen = pos[row];
                if (pen > 0 || n >= RR_NC_ROWS - RR_NDOF) continue;
                double sgn = row ? -1.0 : 1.0;
                double dd = m.Minv[d][d];
                double inv = dd > 0 ? 1.0 / dd : 0.0;
                double rel = sgn * m.qd[d];
                m.r_dof[n] = d; m.r_sign[n] = sgn; m.r_inv[n] = inv; m.r_rinv[n] = inv != 0.0 ? 1.0 / inv : 0.0; m.r_lo[n] = 0.0; m.r_hi[n] = cf.limit_max_impulse;
                m.r_rhs[n] = (-pen * cf.erp / dt - rel) * inv;
                n++;
            }
        }
        for (int d = 0; d < RR_NDOF; d++) {
            double dd = m.Minv[d][d];
            double inv = dd > 0 ? 1.0 / dd : 0.0;
            double cur = m.qd[d];
            double stab = 1.0 * (m.target[d] - m.q[d]) / dt;
            double desired = cf.motor_kp * stab + cur + cf.motor_kd * (0.0 - cur);
            double maxImp = cf.motor_max_force * dt;
            m.r_dof[n] = d; m.r_sign[n] = 1.0; m.r_inv[n] = inv; m.r_rinv[n] = inv != 0.0 ? 1.0 / inv : 0.0; m.r_lo[n] = -maxImp; m.r_hi[n] = maxImp;
            m.r_rhs[n] = (desired - cur) * inv;
            n++;
        }
        m.n_rows = n;
    }
    __syncwarp();
    // ... written out as arm blocks: J = sgn e_d, M^-1 J^T = sgn * column d of M^-1
    const int n_nc = m.n_rows;
    for (int idx = 0; idx < n_nc; idx++) {
        double* blk = gblk + idx * RR_ARM_BLK;
        const int d = m.r_dof[idx]; const double sgn = m.r_sign[idx];
        if (lane < RR_NDOF) { blk[2 * lane] = lane == d ? sgn : 0.0; blk[2 * lane + 1] = m.Minv[lane][d] * sgn; }
        if (lane == 0) { blk[22] = m.r_rhs[idx]; blk[23] = m.r_inv[idx]; blk[24] = m.r_rinv[idx]; blk[25] = m.r_lo[idx]; blk[26] = m.r_hi[idx]; blk[27] = 0.0; }
    }
    int blk_top = n_nc * RR_ARM_BLK;               // doubles used by the blocks so far
    // ---- contact rows (btMultiBodyConstraintSolver::convertMultiBodyContact): per manifold point a normal row, two friction
    //      directions and, where the combined coefficients are > 0, a spinning row (about the normal) and two rolling rows (about
    //      the friction directions).  A row is a 29-vector in "owner" layout: lanes 0..10 = arm dofs, lane 11+6o+k = component k
    //      (ang xyz, lin xyz) of object o; each group the row touches gets a block with its lanes' (J, M^-1 J^T) pairs.
    const double vel = (lane < RR_NDOF) ? m.qd[lane] : 0.0;
    double ovel = 0.0;      // object velocity component owned by this lane (lanes 11..28)
    int nc = 0, overflow = 0, simple = 1, n_rows_total = n_nc;
    int* hdr = man_hdr + (size_t)e * RR_MAX_MANIFOLDS * 2;
    double* pts = man_pts + (size_t)e * RR_MAX_MANIFOLDS * 4 * RR_PT_STRIDE;
    if (cf.enable_contacts) {
        if (lane < RR_NDOF) {
            const int i = c_robot.link_of_dof[lane];
            const double* R = m.k.Rw[i];
#pragma unroll
            for (int a = 0; a < 3; a++) {
                m.dof_axis[lane][a] = R[3 * a] * c_robot.axis[i][0] + R[3 * a + 1] * c_robot.axis[i][1] + R[3 * a + 2] * c_robot.axis[i][2];
                m.dof_piv[lane][a] = m.k.pw[i][a] - (R[3 * a] * c_robot.evec[i][0] + R[3 * a + 1] * c_robot.evec[i][1] + R[3 * a + 2] * c_robot.evec[i][2]);
            }
        }
        if (has_obj) {
            double R[9]; quat_to_mat(ob.quat, R);
            const double ii[3] = {1.0 / c_robot.obj_inertia[lane][0], 1.0 / c_robot.obj_inertia[lane][1], 1.0 / c_robot.obj_inertia[lane][2]};
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++) m.oIinv[lane][3 * a + b] = R[3 * a] * ii[0] * R[3 * b] + R[3 * a + 1] * ii[1] * R[3 * b + 1] + R[3 * a + 2] * ii[2] * R[3 * b + 2];
            m.oinvm[lane] = 1.0 / c_robot.obj_mass[lane];
#pragma unroll
            for (int a = 0; a < 3; a++) { m.opos[lane][a] = ob.pos[a]; m.ovel[lane][a] = ob.ang[a]; m.ovel[lane][3 + a] = ob.lin[a]; }
        }
        __syncwarp();
        const int myo = (lane >= 11 && lane < 29) ? (lane - 11) / 6 : -1, myk = (lane >= 11 && lane < 29) ? (lane - 11) % 6 : 0;
        if (myo >= 0 && myo < n_objects) ovel = m.ovel[myo][myk];
        for (int slot = 0; slot < RR_MAX_MANIFOLDS; slot++) {
            const int pi = hdr[2 * slot];
            if (pi < 0) break;
            const int npt = hdr[2 * slot + 1];
            const int sa = c_robot.pair_a[pi], sb = c_robot.pair_b[pi];
            const int bodyA = c_robot.shape_body[sa], linkA = c_robot.shape_link[sa], bodyB = c_robot.shape_body[sb];
            const bool armA = bodyA == 0 && linkA >= 0;
            const int oA = bodyA >= 2 ? bodyA - 2 : -1, oB = bodyB >= 2 ? bodyB - 2 : -1;
            const double fa = c_robot.shape_friction[sa], fb = c_robot.shape_friction[sb];
            const double mu = fa * fb;
            const double rest = c_robot.shape_restitution[sa] * c_robot.shape_restitution[sb];
            // btManifoldResult::calculateCombinedRollingFriction / ...SpinningFriction: r0 f1 + r1 f0, clamped to +-10
            double roll = fmax(-10.0, fmin(10.0, c_robot.shape_rolling[sa] * fb + c_robot.shape_rolling[sb] * fa));
            double spin = fmax(-10.0, fmin(10.0, c_robot.shape_spinning[sa] * fb + c_robot.shape_spinning[sb] * fa));
            if (!cf.torsional_friction) { roll = 0.0; spin = 0.0; }
            const int gm = (armA ? 1 : 0) | (oA >= 0 ? 2 << oA : 0) | (oB >= 0 ? 2 << oB : 0);
            const bool in_row = mygrp < 4 && ((gm >> mygrp) & 1);
            const int nr = 3 + (spin > 0 ? 1 : 0) + (roll > 0 ? 2 : 0);          // rows of each point of this manifold
            int csize = 0, mygoff = 0;                                           // doubles per point; offset of this lane's group inside it
#pragma unroll
            for (int g = 0; g < 4; g++) if ((gm >> g) & 1) { if (g == mygrp) mygoff = csize; csize += nr * (g == 0 ? RR_ARM_BLK : RR_OBJ_BLK); }
            const int mysz = mygrp == 0 ? RR_ARM_BLK : RR_OBJ_BLK;
            for (int k = 0; k < npt; k++) {
                const double* P = pts + ((size_t)slot * 4 + k) * RR_PT_STRIDE;
                const double dist = P[9];
                if (dist > cf.contact_breaking) continue;
                if (nc >= RR_MAX_CONTACTS) { overflow = 1; continue; }
                const double nB[3] = {P[6], P[7], P[8]}, wA[3] = {P[11], P[12], P[13]}, wB[3] = {P[14], P[15], P[16]};
                double t1[3], t2[3];
                if (fabs(nB[2]) > 0.7071067811865475244008443621048490) {        // btPlaneSpace1
                    double a = nB[1] * nB[1] + nB[2] * nB[2], kk = 1.0 / sqrt(a);
                    t1[0] = 0; t1[1] = -nB[2] * kk; t1[2] = nB[1] * kk;
                    t2[0] = a * kk; t2[1] = -nB[0] * t1[2]; t2[2] = nB[0] * t1[1];
                } else {
                    double a = nB[0] * nB[0] + nB[1] * nB[1], kk = 1.0 / sqrt(a);
                    t1[0] = -nB[1] * kk; t1[1] = nB[0] * kk; t1[2] = 0;
                    t2[0] = -nB[2] * t1[1]; t2[1] = nB[2] * t1[0]; t2[2] = a * kk;
                }
                {
                    double l1 = 1.0 / sqrt(t1[0] * t1[0] + t1[1] * t1[1] + t1[2] * t1[2]), l2 = 1.0 / sqrt(t2[0] * t2[0] + t2[1] * t2[1] + t2[2] * t2[2]);
                    for (int a = 0; a < 3; a++) { t1[a] *= l1; t2[a] *= l2; }
                }
                // di: 0 normal, 1 2 friction, 3 spinning (angular, about the normal), 4 5 rolling (angular, about t1 / t2)
                int rowidx = 0;
                for (int di = 0; di < 6; di++) {
                    if (di == 3 && !(spin > 0)) continue;
                    if (di >= 4 && !(roll > 0)) break;
                    const bool angular = di >= 3;
                    const int dd_ = angular ? di - 3 : di;
                    const double* d = dd_ == 0 ? nB : (dd_ == 1 ? t1 : t2);
                    double J = 0.0;
                    if (lane < RR_NDOF) {
                        if (armA && ((c_robot.anc[linkA] >> lane) & 1)) {
                            const double* aw = m.dof_axis[lane]; const double* pv = m.dof_piv[lane];
                            if (angular) J = d[0] * aw[0] + d[1] * aw[1] + d[2] * aw[2];
                            else {
                                double r0 = wA[0] - pv[0], r1 = wA[1] - pv[1], r2 = wA[2] - pv[2];
                                J = d[0] * (aw[1] * r2 - aw[2] * r1) + d[1] * (aw[2] * r0 - aw[0] * r2) + d[2] * (aw[0] * r1 - aw[1] * r0);
                            }
                        }
                    } else if (myo >= 0 && (myo == oA || myo == oB)) {
                        const double sg = myo == oA ? 1.0 : -1.0;
                        const double dd0 = d[0] * sg, dd1 = d[1] * sg, dd2 = d[2] * sg;
                        if (angular) J = myk == 0 ? dd0 : (myk == 1 ? dd1 : (myk == 2 ? dd2 : 0.0));
                        else {
                            const double* w = myo == oA ? wA : wB;
                            double r0 = w[0] - m.opos[myo][0], r1 = w[1] - m.opos[myo][1], r2 = w[2] - m.opos[myo][2];
                            if (myk == 0) J = r1 * dd2 - r2 * dd1; else if (myk == 1) J = r2 * dd0 - r0 * dd2; else if (myk == 2) J = r0 * dd1 - r1 * dd0;
                            else J = myk == 3 ? dd0 : (myk == 4 ? dd1 : dd2);
                        }
                    }
                    m.Jtmp[lane] = J;
                    __syncwarp();
                    double rr = 0.0;
                    if (lane < RR_NDOF) {
                        if (armA) {
#pragma unroll
                            for (int j = 0; j < RR_NDOF; j++) rr += m.Minv[lane][j] * m.Jtmp[j];
                        }
                    } else if (myo >= 0 && (myo == oA || myo == oB)) {
                        if (myk < 3) rr = m.oIinv[myo][3 * myk] * m.Jtmp[11 + 6 * myo] + m.oIinv[myo][3 * myk + 1] * m.Jtmp[12 + 6 * myo] + m.oIinv[myo][3 * myk + 2] * m.Jtmp[13 + 6 * myo];
                        else rr = J * m.oinvm[myo];
                    }
                    double den, relv; warp_sum2(J * rr, J * (lane < RR_NDOF ? vel : ovel), den, relv);      // two reductions for the price of one
                    const double inv = den > 0 ? 1.0 / den : 0.0;
                    double rhs, app0 = 0.0, lo = 0.0, hi = 1e10;         // normal rows: [0, 1e10]; the others: lo = the coefficient, limits follow the normal impulse
                    if (di == 0) {
                        double restitution = (fabs(relv) < cf.restitution_threshold) ? 0.0 : rest * -relv;
                        if (restitution <= 0) restitution = 0;
                        const double pen = dist + cf.linear_slop;
                        double positional = 0, velocity = restitution - relv;
                        const double erp = (pen > -0.04) ? cf.erp : cf.erp2;
                        if (pen > 0) velocity -= pen / dt; else positional = -pen * erp / dt;
                        rhs = (positional + velocity) * inv;
                        app0 = P[10] * cf.warmstart;
                    } else { rhs = (0.0 - relv) * inv; lo = di < 3 ? mu : (di == 3 ? spin : roll); hi = 0.0; }
                    if (in_row) {
                        double* blk = gblk + blk_top + mygoff + rowidx * mysz;
                        blk[myk2] = J; blk[myk2 + 1] = rr;
                        if (leader) { blk[mysc] = rhs; blk[mysc + 1] = inv; blk[mysc + 2] = inv != 0.0 ? 1.0 / inv : 0.0; blk[mysc + 3] = lo; blk[mysc + 4] = hi; blk[mysc + 5] = app0; }
                    }
                    rowidx++;
                    __syncwarp();
                }
                if (lane == 0) {
                    const bool wide = armA || (gm & (gm - 1)) != 0;       // needs a sum over more lanes than one group's own shuffles cover
                    m.c_grp[nc] = (unsigned char)gm;
                    m.c_flags[nc] = (unsigned char)((spin > 0 ? 1 : 0) | (roll > 0 ? 2 : 0) | (wide ? 4 : 0));
                    m.c_base[nc] = (unsigned short)(blk_top / 2);
                    hand->nrm_rel[nc] = (unsigned short)((blk_top / 2) | (armA ? 0x8000 : 0));
                }
                if (armA || oA < 0 || oB >= 0) simple = 0;
                blk_top += csize;
                n_rows_total += nr;
                nc++;
            }
        }
    }
    __syncwarp();
    // ---- slot schedule of one Gauss-Seidel iteration (see SolveHead), for both parities of the joint-row order
    const int n_single = n_nc + 4 * nc;          // upper bound of the slots of the single rows
    for (int k = lane; k < n_single * 8; k += 32) { gsched[GSCHED_A(0) + k] = ENT_IDLE; gsched[GSCHED_A(1) + k] = ENT_IDLE; }
    for (int k = lane; k < nc * 8; k += 32) gsched[GSCHED_B + k] = ENT_IDLE;
    __syncwarp();
    if (lane == 0) {
        unsigned int ent[4] = {0u, 0u, 0u, 0u};
        for (int par = 0; par < 2; par++) {
            int cur[4] = {0, 0, 0, 0};
            unsigned int wide_used[(RR_SLOTS_A + 31) / 32];
            for (int k = 0; k < (RR_SLOTS_A + 31) / 32; k++) wide_used[k] = 0u;
            unsigned int* sched = gsched + GSCHED_A(par);
            for (int j = 0; j < n_nc; j++) {
                const int idx = par ? j : n_nc - 1 - j;                               // it & 1 ? j : n - 1 - j
                ent[0] = (unsigned int)(idx * (RR_ARM_BLK / 2)) | ((ENT_JOINT | (unsigned int)m.r_dof[idx]) << 16);
                sched_place(sched, cur, wide_used, 1, false, ent, (unsigned int)idx);
            }
            for (int c = 0; c < nc; c++) {
                const int nr = 3 + (m.c_flags[c] & 1) + (m.c_flags[c] & 2);
                contact_entries(ent, m.c_grp[c], m.c_base[c], nr, 0);
                sched_place(sched, cur, wide_used, m.c_grp[c], (m.c_flags[c] & 4) != 0, ent, (unsigned int)(RID_NORMAL + c) | ((unsigned int)(RID_NORMAL + c) << 16));
            }
            for (int c = 0; c < nc; c++) if (m.c_flags[c] & 1) {
                const int nr = 3 + (m.c_flags[c] & 1) + (m.c_flags[c] & 2);
                contact_entries(ent, m.c_grp[c], m.c_base[c], nr, 3);
                sched_place(sched, cur, wide_used, m.c_grp[c], (m.c_flags[c] & 4) != 0, ent, (unsigned int)(RID_SPIN + c) | ((unsigned int)(RID_NORMAL + c) << 16));
            }
            for (int c = 0; c < nc; c++) if (m.c_flags[c] & 2) {
                const int nr = 3 + (m.c_flags[c] & 1) + (m.c_flags[c] & 2);
                for (int r2 = 0; r2 < 2; r2++) {
                    contact_entries(ent, m.c_grp[c], m.c_base[c], nr, 3 + (m.c_flags[c] & 1) + r2);
                    sched_place(sched, cur, wide_used, m.c_grp[c], (m.c_flags[c] & 4) != 0, ent, (unsigned int)(RID_ROLL + 2 * c + r2) | ((unsigned int)(RID_NORMAL + c) << 16));
                }
            }
            hand->n_slotsA[par] = max(max(cur[0], cur[1]), max(cur[2], cur[3]));
        }
        int cur[4] = {0, 0, 0, 0};
        unsigned int wide_used[(RR_SLOTS_B + 31) / 32];
        for (int k = 0; k < (RR_SLOTS_B + 31) / 32; k++) wide_used[k] = 0u;
        for (int c = 0; c < nc; c++) {
            const int nr = 3 + (m.c_flags[c] & 1) + (m.c_flags[c] & 2);
            contact_entries(ent, m.c_grp[c], m.c_base[c], nr, 1);                     // first friction row; the second follows it
            sched_place(gsched + GSCHED_B, cur, wide_used, m.c_grp[c], (m.c_flags[c] & 4) != 0, ent, (unsigned int)(RID_FRIC + 2 * c) | ((unsigned int)(RID_NORMAL + c) << 16));
        }
        hand->n_slotsB = max(max(cur[0], cur[1]), max(cur[2], cur[3]));
        hand->n_nc = n_nc; hand->nc = nc; hand->overflow = overflow; hand->fault = m.fault; hand->simple = simple;
        hand->n_rows_total = n_rows_total; hand->n_blk_doubles = blk_top;
        // cost class of the sweep (slots per iteration; a pair slot counts twice): k_step_solve starts the longest chains first
        const int cost = (hand->n_slotsA[0] + hand->n_slotsA[1]) / 2 + 2 * hand->n_slotsB;
        const int cls = solve_cost_class(cost);
        const int pos = atomicAdd(&st.solve_order[cls], 1);
        st.solve_order[RR_SOLVE_CLASSES + cls * N + pos] = e;
        st.solve_pos[e] = (cls << 24) | pos;
    }
    // ---- hand the state over
    if (lane < RR_NDOF) { hand->q[lane] = m.q[lane]; hand->qd[lane] = m.qd[lane]; hand->target[lane] = m.target[lane]; }
    if (has_obj) {
#pragma unroll
        for (int k = 0; k < 3; k++) { hand->objreg[lane][k] = ob.pos[k]; hand->objreg[lane][7 + k] = ob.lin[k]; hand->objreg[lane][10 + k] = ob.ang[k]; }
#pragma unroll
        for (int k = 0; k < 4; k++) hand->objreg[lane][3 + k] = ob.quat[k];
    }
}

// ------------------------------------------------------------------------------------------- the solver kernel
#ifdef RR_SIMT_EMU_BUILD
#define RR_PREFETCH_L1(p)
#else
#define RR_PREFETCH_L1(p) asm volatile("prefetch.global.L1 [%0];" :: "l"(p))
#endif
// sum of the six products of a lane's own group (an arm lane of a joint row: the product of the row's dof), all lanes at once
__device__ __forceinline__ double group_sum6(double pr, int s0, int s1, int s2, int s3, int s4, int s5) {
    const double a = __shfl_sync(FULL, pr, s0), b = __shfl_sync(FULL, pr, s1), c = __shfl_sync(FULL, pr, s2);
    const double d = __shfl_sync(FULL, pr, s3), e = __shfl_sync(FULL, pr, s4), f = __shfl_sync(FULL, pr, s5);
    return ((a + b) + (c + d)) + (e + f);
}
#ifndef SWEEP_WAR_FENCE
#define SWEEP_WAR_FENCE() __syncwarp()
#endif
// The Gauss-Seidel sweep of one env (see SolveHead).  SS: the schedule pointers point into shared memory.  The blocks are streamed
// from L2, prefetched three slots ahead.
template <bool SS>
__device__ __forceinline__ void pgs_sweep(const unsigned int* const schedA0, const unsigned int* const schedA1, const unsigned int* const schedB0, const double* const gblk,
                                          double* const sapp, const int nA0, const int nA1, const int nB, const int lane, const rr_config& cf, double& dv_out, int& iters_out) {
    // dv in owner layout: lane k < 11 owns the arm dof k, lane 11+6o+c the object component.
    const int grp = lane_group(lane);
    const bool ingrp = grp < 4;
    const int g4 = ingrp ? grp : 0;                            // idle lanes read group 0's entry and ignore it
    const int seg0 = grp == 0 ? 0 : (ingrp ? 11 + 6 * (grp - 1) : 31);   // first lane of the own group
    const bool leader = ingrp && lane == seg0;
    const int myk2 = 2 * (lane - seg0), mysc = grp == 0 ? 2 * RR_NDOF : 12, mysz = grp == 0 ? RR_ARM_BLK : RR_OBJ_BLK;
    // shuffle sources of the group sum: the six lanes of a free body; for arm lanes the row's dof, then lane 31 (always 0)
    const bool isobj = grp >= 1 && grp < 4;
    const int s1 = isobj ? seg0 + 1 : 31, s2 = isobj ? seg0 + 2 : 31, s3 = isobj ? seg0 + 3 : 31, s4 = isobj ? seg0 + 4 : 31, s5 = isobj ? seg0 + 5 : 31;
    double dv = 0.0;
    int iters = 0; double res = 0.0;
    {
        for (int slot = 0; slot < nA0; slot++) {             // warm start: dv of the normal rows' initial impulses
            const uint2 w = *reinterpret_cast<const uint2*>(schedA0 + 8 * slot + 2 * g4);
            if (ingrp && w.x != ENT_IDLE && (w.x >> 16) == ENT_NORMAL) {
                const double a = sapp[w.y & 0xffffu];
                if (a != 0.0) dv += gblk[2 * (w.x & 0x7fffu) + myk2 + 1] * a;
            }
        }
        for (int it = 0; it < cf.solver_iterations; it++) {
            res = 0.0;
            const int par = it & 1;
            // ---- phase A: single rows (joint limits + motors, normals, spinning, rolling), one slot at a time
            const unsigned int* sched = (par ? schedA1 : schedA0) + 2 * g4;
            const int nA = par ? nA1 : nA0;
            unsigned int nx = nA > 3 ? sched[8 * 3] : ENT_IDLE;                 // word 0 of the entry three slots ahead (loaded one slot early)
            for (int slot = 0; slot < nA; slot++) {
                const uint2 w = *reinterpret_cast<const uint2*>(sched + 8 * slot);
                const bool act = ingrp && w.x != ENT_IDLE;
                const unsigned int kind = w.x >> 16;
                const double* b = gblk + 2 * (act ? (w.x & 0x7fffu) : 0u);
                if (ingrp && nx != ENT_IDLE) RR_PREFETCH_L1(gblk + 2 * (nx & 0x7fffu) + myk2);      // the blocks three slots ahead are on their way into L1
                nx = slot + 4 < nA ? sched[8 * (slot + 4)] : ENT_IDLE;
                double J = 0.0, rr = 0.0, rhs = 0.0, inv = 0.0, rinv = 0.0, lo = 0.0, hi = 0.0, app = 0.0;
                bool on = act;
                const int rid = act ? (int)(w.y & 0xffffu) : 0;
                if (act) {
                    const double2 jr = *reinterpret_cast<const double2*>(b + myk2);
                    const double2 c0 = *reinterpret_cast<const double2*>(b + mysc), c1 = *reinterpret_cast<const double2*>(b + mysc + 2);
                    J = jr.x; rr = jr.y; rhs = c0.x; inv = c0.y; rinv = c1.x; lo = c1.y; hi = b[mysc + 4]; app = sapp[rid];
                    if (kind < ENT_JOINT) {           // spinning / rolling row: limits +-(coefficient x normal impulse), solved only while that impulse is positive
                        const double total = sapp[w.y >> 16];
                        on = total > 0.0;
                        hi = lo * total; lo = -hi;
                    }
                }
                const int s0 = isobj ? seg0 : ((act && kind >= ENT_JOINT && kind < ENT_NORMAL) ? (int)(kind & 15u) : 31);
                const double pr = J * dv;
                double s = group_sum6(pr, s0, s1, s2, s3, s4, s5);
                const bool inw = act && (w.x & 0x8000u);
                if (__any_sync(FULL, inw)) {                 // the slot's wide row sums over all of its lanes
                    const double t = warp_sum(inw ? pr : 0.0);
                    if (inw) s = t;
                }
                double delta = rhs - s * inv;
                const double sum = app + delta;
                if (sum < lo) { delta = lo - app; app = lo; }
                else if (sum > hi) { delta = hi - app; app = hi; }
                else app = sum;
                if (on) {
                    dv += rr * delta;
                    const double r = delta * rinv;
                    res = fmax(res, r * r);
                }
                SWEEP_WAR_FENCE();                            // every lane has read this slot's impulses (the shuffles above converge the
                if (on && leader) sapp[rid] = app;           // warp; this orders the shared-memory accesses as well)
                __syncwarp();
            }
            // ---- phase B: friction pairs with the implicit cone
            const unsigned int* schedb = schedB0 + 2 * g4;
            for (int slot = 0; slot < nB; slot++) {
                const uint2 w = *reinterpret_cast<const uint2*>(schedb + 8 * slot);
                const bool act = ingrp && w.x != ENT_IDLE;
                const double* ba = gblk + 2 * (act ? (w.x & 0x7fffu) : 0u);
                const double* bb = ba + mysz;
                const int r0 = act ? (int)(w.y & 0xffffu) : 0;
                double Ja = 0.0, rra = 0.0, Jb = 0.0, rrb = 0.0, rhsa = 0.0, inva = 0.0, rhsb = 0.0, invb = 0.0, rinva = 0.0, rinvb = 0.0, lim = 0.0, appa = 0.0, appb = 0.0;
                if (act) {
                    const double2 ja = *reinterpret_cast<const double2*>(ba + myk2), jb = *reinterpret_cast<const double2*>(bb + myk2);
                    const double2 a0 = *reinterpret_cast<const double2*>(ba + mysc), a1 = *reinterpret_cast<const double2*>(ba + mysc + 2);
                    const double2 b0 = *reinterpret_cast<const double2*>(bb + mysc);
                    Ja = ja.x; rra = ja.y; Jb = jb.x; rrb = jb.y;
                    rhsa = a0.x; inva = a0.y; rinva = a1.x; rhsb = b0.x; invb = b0.y; rinvb = bb[mysc + 2];
                    lim = a1.y * sapp[w.y >> 16];
                    appa = sapp[r0]; appb = sapp[r0 + 1];
                }
                const int s0 = isobj ? seg0 : 31;
                const double pa = Ja * dv, pb = Jb * dv;
                double sa_ = group_sum6(pa, s0, s1, s2, s3, s4, s5), sb_ = group_sum6(pb, s0, s1, s2, s3, s4, s5);
                const bool inw = act && (w.x & 0x8000u);
                if (__any_sync(FULL, inw)) {
                    double ta, tb; warp_sum2(inw ? pa : 0.0, inw ? pb : 0.0, ta, tb);
                    if (inw) { sa_ = ta; sb_ = tb; }
                }
                double wa = appa + (rhsa - sa_ * inva), wb = appb + (rhsb - sb_ * invb);
                const double len2 = wa * wa + wb * wb;
                if (len2 > lim * lim) { const double scl = len2 > 0 ? lim / sqrt(len2) : 0.0; wa *= scl; wb *= scl; }
                const double da = wa - appa, db = wb - appb;
                if (act) {
                    dv += rra * da; dv += rrb * db;
                    const double qa = da * rinva, qb = db * rinvb;
                    res = fmax(res, fmax(qa * qa, qb * qb));
                }
                SWEEP_WAR_FENCE();
                if (act && leader) { sapp[r0] = wa; sapp[r0 + 1] = wb; }
                __syncwarp();
            }
            res = warp_max(res);
            iters = it + 1;
            if (res <= cf.residual_threshold) break;
        }
    }
    dv_out = dv; iters_out = iters;
}

// One warp per env, 24 warps per SM.  The schedule and the blocks are read-only here and stream from L2 (the images of the
// envs in flight — ~14 KB each — stay L2-resident; a block is one or two 128-byte lines); the applied impulses, the only
// thing the sweep updates besides dv, live in shared memory by row id.  (Measured alternatives, same results: the whole image
// in shared memory, one env per warp — 12 warps per SM, and lane-per-(env, group) streams, 7 envs per warp — 2 warps per SM:
// both residency-bound at 1.9 ms for 4096 envs.)
#ifndef SOLVE_WPB
#define SOLVE_WPB 4
#endif
#ifndef SOLVE_MIN_BLOCKS
#define SOLVE_MIN_BLOCKS 6           // blocks per SM of the main kernel when it has every env (24 warps, 80 registers): the long chains of the
#define SOLVE_MIN_BLOCKS_SPLIT 7     // heavy envs want latency; when they have their own kernel, 7 (28 warps, 72 registers) — measured best of
#endif                               // 5 / 6 / 7 / 8 blocks x 4 warps, 14 x 2, 28 x 1, 3 x 8 in either mode (back-to-back steps 3.52 vs 4.05 ms, bench step 6.55 vs 6.45)
struct SolveEpi { SolveHead head; KinWs kin; };
#define SOLVE_SMEM_DOUBLES ((int)(sizeof(SolveEpi) / 8) > RR_MAX_ROWS ? (int)(sizeof(SolveEpi) / 8) : RR_MAX_ROWS)
#ifndef SOLVE_SCHED_WORDS
#define SOLVE_SCHED_WORDS 512       // words of the slot schedule kept in shared memory (64 slots; longer schedules are read from the image)
#endif
#define SOLVE_SMEM_BYTES ((size_t)SOLVE_SMEM_DOUBLES * 8 + SOLVE_SCHED_WORDS * 4)
// Heavy envs: the few envs (2-5 % under a random policy) whose arm touches something have schedules 1.5-3.5x as long as the rest,
// and a Gauss-Seidel chain cannot be split: started together with the others, the longest of them alone set the kernel's duration
// (2.2 ms against 1.0 ms for a batch without them).  They get a kernel of their own on a side stream, launched beside the main one:
// one warp per block, the whole schedule in shared memory, one env each; same code, same arithmetic.  The chain is not faster there
// (a slot is ~100 dependent fp64 / shuffle instructions: keeping the row blocks in shared memory too was measured and changed
// nothing) — the point is that nothing on the step's critical path waits for it: rr_render's
// rasteriser starts when the MAIN kernel is done, takes the heavy envs last and waits for each one's ticket (DevState::ready);
// whoever else touches the state joins the side stream first (rrk_join_heavy).
#define HEAVY_SCHED_WORDS (16 * RR_SLOTS_A + 8 * RR_SLOTS_B)
#define HEAVY_SMEM_BYTES ((size_t)SOLVE_SMEM_DOUBLES * 8 + HEAVY_SCHED_WORDS * 4)
#ifndef HEAVY_GRID
#define HEAVY_GRID 192        // one env per block, all resident at once.  Not more: their narrowphase (4 envs per block, ~0.6 ms each) has
#define HEAVY_PER_BLOCK 1     // to fit under the rasteriser on the SMs that kernel leaves free; the envs beyond go to the main kernel, first
#endif
template <bool HEAVY>
__device__ __forceinline__ void solve_env(const DevState& st, const int e, double* const smem, unsigned int* const ssched, const int sched_cap,
                                          int* __restrict__ man_hdr, double* __restrict__ man_pts, const rr_config& cf, const int n_objects, const int lane) {
    double* const sapp = smem;                                 // applied impulses by row id (during the sweep)
    const int N = st.N;
    const double dt = cf.dt;
    const double* gimg = st.handoff + (size_t)e * RR_IMG_DOUBLES;
    const SolveHead* gh = reinterpret_cast<const SolveHead*>(gimg);
    const int nA0 = gh->n_slotsA[0], nA1 = gh->n_slotsA[1], nB = gh->n_slotsB, nc = gh->nc;
    const unsigned int* gsched = reinterpret_cast<const unsigned int*>(gimg + HEAD_DOUBLES);
    const double* gblk = gimg + GBLK_BASE;
    // ---- applied impulses: zero, except the warm start of the normal rows
    for (int k = lane; k < RR_MAX_ROWS; k += 32) sapp[k] = 0.0;
    __syncwarp();
    for (int c = lane; c < nc; c += 32) {
        const unsigned int nr = gh->nrm_rel[c];
        sapp[RID_NORMAL + c] = gblk[2 * (nr & 0x7fffu) + ((nr & 0x8000u) ? 2 * RR_NDOF + 5 : 17)];
    }
    __syncwarp();
    // ---- the slot schedule comes into shared memory when it fits (it is read by every lane in every slot of every iteration)
    const bool sched_smem = 8 * (nA0 + nA1 + nB) <= sched_cap;
    if (sched_smem) {
        for (int k = lane; k < 8 * nA0; k += 32) ssched[k] = gsched[GSCHED_A(0) + k];
        for (int k = lane; k < 8 * nA1; k += 32) ssched[8 * nA0 + k] = gsched[GSCHED_A(1) + k];
        for (int k = lane; k < 8 * nB; k += 32) ssched[8 * (nA0 + nA1) + k] = gsched[GSCHED_B + k];
    }
    __syncwarp();
    double dv = 0.0;
    int iters = 0;
    if (sched_smem) pgs_sweep<true>(ssched, ssched + 8 * nA0, ssched + 8 * (nA0 + nA1), gblk, sapp, nA0, nA1, nB, lane, cf, dv, iters);
    else pgs_sweep<false>(gsched + GSCHED_A(0), gsched + GSCHED_A(1), gsched + GSCHED_B, gblk, sapp, nA0, nA1, nB, lane, cf, dv, iters);
    // ---- impulse write-back (warm start of the next step) and gripper touch sensors (robot.py:131-163)
    int* hdr = man_hdr + (size_t)e * RR_MAX_MANIFOLDS * 2;
    double* pts = man_pts + (size_t)e * RR_MAX_MANIFOLDS * 4 * RR_PT_STRIDE;
    double touch[4] = {0.0, 0.0, 0.0, 0.0};
    if (cf.enable_contacts) {
        int c = 0;
        for (int slot = 0; slot < RR_MAX_MANIFOLDS; slot++) {
            const int pi = hdr[2 * slot];
            if (pi < 0) break;
            const int npt = hdr[2 * slot + 1];
            const int sa = c_robot.pair_a[pi], sb = c_robot.pair_b[pi];
            const int linkA = c_robot.shape_link[sa];
            const bool sens = c_robot.shape_body[sa] == 0 && c_robot.shape_body[sb] != 0 && c_robot.shape_link[sb] == -1;
            const int sensor = linkA == 11 ? 0 : (linkA == 10 ? 1 : (linkA == 15 ? 2 : (linkA == 14 ? 3 : -1)));
            bool any = false; double best = 0.0;
            for (int k = 0; k < npt; k++) {
                double* P = pts + ((size_t)slot * 4 + k) * RR_PT_STRIDE;
                double imp = P[10];
                if (!(P[9] > cf.contact_breaking) && c < RR_MAX_CONTACTS) { imp = sapp[RID_NORMAL + c]; if (lane == 0) P[10] = imp; c++; }
                if (sens && sensor >= 0 && fabs(P[9]) < 0.1) { const double f = imp / dt; if (!any || f > best) { best = f; any = true; } }
            }
            if (any && sens && sensor >= 0 && best > touch[sensor]) touch[sensor] = best;
        }
    }
    __syncwarp();
    // ---- the sweep's shared memory now holds the head of the image and the kinematics workspace
    SolveEpi& ep = *reinterpret_cast<SolveEpi*>(smem);
    for (int k = lane; k < HEAD_DOUBLES; k += 32) reinterpret_cast<double*>(&ep.head)[k] = gimg[k];
    __syncwarp();
    SolveHead& m = ep.head;
    ObjReg ob;
    const bool has_obj = lane < n_objects;
    if (has_obj) {
#pragma unroll
        for (int k = 0; k < 3; k++) { ob.pos[k] = m.objreg[lane][k]; ob.lin[k] = m.objreg[lane][7 + k]; ob.ang[k] = m.objreg[lane][10 + k]; }
#pragma unroll
        for (int k = 0; k < 4; k++) ob.quat[k] = m.objreg[lane][3 + k];
    }
    const int overflow = m.overflow, simple = m.simple, rows_live = m.n_rows_total;
    int fault = m.fault;
    if (lane < RR_NDOF) {
        m.qd[lane] = clampd(m.qd[lane] + dv, -cf.max_coord_vel, cf.max_coord_vel);
        m.q[lane] += m.qd[lane] * dt;
    }
    {   // hand the object delta-velocities back to the lanes that own the bodies
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const double da = __shfl_sync(FULL, dv, 11 + 6 * (lane % 3) + k), dl = __shfl_sync(FULL, dv, 14 + 6 * (lane % 3) + k);
            if (has_obj) {
                ob.ang[k] = clampd(ob.ang[k] + da, -cf.max_coord_vel, cf.max_coord_vel);
                ob.lin[k] = clampd(ob.lin[k] + dl, -cf.max_coord_vel, cf.max_coord_vel);
            }
        }
    }
    if (has_obj) obj_integrate(ob, dt);
    // ---- failure detection: a non-finite state cannot recover by itself (and the reference has nothing to say about it);
    // the env is put back into its initial state as rr_reset would, keeps stepping, and is flagged in rr_diag (bit 3)
    {
        bool bad = lane < RR_NDOF && !(isfinite(m.q[lane]) && isfinite(m.qd[lane]));
        if (has_obj) {
#pragma unroll
            for (int k = 0; k < 3; k++) bad |= !(isfinite(ob.pos[k]) && isfinite(ob.lin[k]) && isfinite(ob.ang[k]));
#pragma unroll
            for (int k = 0; k < 4; k++) bad |= !isfinite(ob.quat[k]);
        }
        if (__any_sync(FULL, bad)) {
            fault |= 2;
            if (lane < RR_NDOF) { m.q[lane] = 0.0; m.qd[lane] = 0.0; m.target[lane] = 0.0; }
            if (has_obj) {
#pragma unroll
                for (int k = 0; k < 3; k++) { ob.pos[k] = c_robot.obj_init[lane][k]; ob.lin[k] = 0.0; ob.ang[k] = 0.0; }
#pragma unroll
                for (int k = 0; k < 4; k++) ob.quat[k] = c_robot.obj_init[lane][3 + k];
            }
            for (int slot = lane; slot < RR_MAX_MANIFOLDS; slot += 32) { hdr[2 * slot] = -1; hdr[2 * slot + 1] = 0; }
#pragma unroll
            for (int k = 0; k < 4; k++) touch[k] = 0.0;
        }
    }
    __syncwarp();
    // ---- final kinematics, state write-back, observations
    fk_local(m.q, ep.kin, lane);
    fk_sweep(ep.kin, lane, nullptr, nullptr);
    store_frames(ep.kin, st, e, lane);
    if (lane < RR_NDOF) {
        st.q[(size_t)lane * N + e] = m.q[lane]; st.qd[(size_t)lane * N + e] = m.qd[lane];
        st.target[(size_t)lane * N + e] = m.target[lane];
    }
    if (has_obj) {
        double* base = st.obj + (size_t)lane * 13 * N + e;
#pragma unroll
        for (int k = 0; k < 3; k++) base[(size_t)k * N] = ob.pos[k];
#pragma unroll
        for (int k = 0; k < 4; k++) base[(size_t)(3 + k) * N] = ob.quat[k];
#pragma unroll
        for (int k = 0; k < 3; k++) { base[(size_t)(7 + k) * N] = ob.lin[k]; base[(size_t)(10 + k) * N] = ob.ang[k]; }
    }
    if (lane < RR_NUM_TOUCH) st.touch[(size_t)lane * N + e] = lane == 0 ? touch[0] : (lane == 1 ? touch[1] : (lane == 2 ? touch[2] : touch[3]));
    if (lane == 0) {
        st.diag[4 * e] = rows_live; st.diag[4 * e + 1] = iters; st.diag[4 * e + 2] = nc; st.diag[4 * e + 3] = overflow | (simple ? 2 : 0) | (fault << 2) | (cf.enable_contacts ? (st.col_flags[e] << 4) : 0) | ((((nA0 + nA1) >> 1) + 2 * nB) << 16);   // bit 0: contact cap hit, bit 1: every contact is object-vs-static, bits 2-3: faults, bit 4: manifold slots exhausted, bits 16-31: sweep cost (slots per iteration, pair slots twice)
        st.timestep[e] += 1;
    }
}
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(SOLVE_WPB * 32, MIN_BLOCKS) k_step_solve(DevState st, int* __restrict__ man_hdr, double* __restrict__ man_pts, rr_config cf, int n_objects, int heavy_cap) {
    RR_DYN_SMEM(smem_raw);
    double* const smem = reinterpret_cast<double*>(smem_raw + (size_t)(threadIdx.x >> 5) * SOLVE_SMEM_BYTES);
    const int idx = blockIdx.x * SOLVE_WPB + (threadIdx.x >> 5) + solve_n_heavy(st, heavy_cap);      // the heavy kernel has the first ones
    if (idx >= st.N) return;
    solve_env<false>(st, solve_pick(st, idx), smem, reinterpret_cast<unsigned int*>(smem + SOLVE_SMEM_DOUBLES), SOLVE_SCHED_WORDS, man_hdr, man_pts, cf, n_objects, threadIdx.x & 31);
}
__global__ void __launch_bounds__(32) k_step_solve_heavy(DevState st, int* __restrict__ man_hdr, double* __restrict__ man_pts, rr_config cf, int n_objects, int heavy_cap, int ticket) {
    RR_DYN_SMEM(smem_raw);
    double* const smem = reinterpret_cast<double*>(smem_raw);
    unsigned int* const ssched = reinterpret_cast<unsigned int*>(smem + SOLVE_SMEM_DOUBLES);
    const int n = solve_n_heavy(st, heavy_cap);
    for (int idx = blockIdx.x; idx < n; idx += gridDim.x) {
        const int e = solve_pick(st, idx);
        solve_env<true>(st, e, smem, ssched, HEAVY_SCHED_WORDS, man_hdr, man_pts, cf, n_objects, threadIdx.x);
        __threadfence();                      // the env's state, frames and observations are out before the ticket is
        __syncwarp();
        if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(st.ready + e) = ticket;
    }
}

// FK only (after reset / set_state): refresh link_frames
struct FkMem { double q[RR_NDOF]; KinWs k; };
__global__ void __launch_bounds__(WPB * 32) k_fk(DevState st) {
    RR_DYN_SMEM(smem_raw);
    FkMem& m = *reinterpret_cast<FkMem*>(smem_raw + (size_t)(threadIdx.x >> 5) * sizeof(FkMem));
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * WPB + (threadIdx.x >> 5);
    if (e >= st.N) return;
    if (lane < RR_NDOF) m.q[lane] = st.q[(size_t)lane * st.N + e];
    __syncwarp();
    fk_local(m.q, m.k, lane);
    fk_sweep(m.k, lane, nullptr, nullptr);
    store_frames(m.k, st, e, lane);
}

__global__ void k_reset(DevState st, const uint8_t* __restrict__ mask, int n_objects, int* __restrict__ man_hdr) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    if (mask && !mask[e]) return;
    for (int k = 0; k < RR_MAX_MANIFOLDS; k++) { man_hdr[((size_t)e * RR_MAX_MANIFOLDS + k) * 2] = -1; man_hdr[((size_t)e * RR_MAX_MANIFOLDS + k) * 2 + 1] = 0; }
    const int N = st.N;
    for (int d = 0; d < RR_NDOF; d++) { st.q[(size_t)d * N + e] = 0.0; st.qd[(size_t)d * N + e] = 0.0; st.target[(size_t)d * N + e] = 0.0; }
    for (int o = 0; o < RR_MAX_OBJECTS; o++) {
        double* base = st.obj + (size_t)o * 13 * N + e;
        for (int k = 0; k < 7; k++) base[(size_t)k * N] = c_robot.obj_init[o][k];
        for (int k = 7; k < 13; k++) base[(size_t)k * N] = 0.0;
    }
    for (int k = 0; k < RR_NUM_TOUCH; k++) st.touch[(size_t)k * N + e] = 0.0;
    st.timestep[e] = 0;
    for (int k = 0; k < 4; k++) st.diag[4 * e + k] = 0;
}

__global__ void k_set_object_pose(DevState st, const uint8_t* __restrict__ mask, int obj, const double* __restrict__ pq) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    if (mask && !mask[e]) return;
    double* base = st.obj + (size_t)obj * 13 * st.N + e;
    for (int k = 0; k < 7; k++) base[(size_t)k * st.N] = pq[(size_t)e * 7 + k];
    for (int k = 7; k < 13; k++) base[(size_t)k * st.N] = 0.0;
}

// AoS [N,...] <-> SoA [...][N] transposes for the state hooks
__global__ void k_set_state(DevState st, const double* __restrict__ q, const double* __restrict__ qd, const double* __restrict__ obj) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    const int N = st.N;
    if (q) for (int d = 0; d < RR_NDOF; d++) st.q[(size_t)d * N + e] = q[(size_t)e * RR_NDOF + d];
    if (qd) for (int d = 0; d < RR_NDOF; d++) st.qd[(size_t)d * N + e] = qd[(size_t)e * RR_NDOF + d];
    if (obj) for (int k = 0; k < 39; k++) st.obj[(size_t)k * N + e] = obj[(size_t)e * 39 + k];
}
__global__ void k_get_state(DevState st, double* __restrict__ q, double* __restrict__ qd, double* __restrict__ obj) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    const int N = st.N;
    if (q) for (int d = 0; d < RR_NDOF; d++) q[(size_t)e * RR_NDOF + d] = st.q[(size_t)d * N + e];
    if (qd) for (int d = 0; d < RR_NDOF; d++) qd[(size_t)e * RR_NDOF + d] = st.qd[(size_t)d * N + e];
    if (obj) for (int k = 0; k < 39; k++) obj[(size_t)e * 39 + k] = st.obj[(size_t)k * N + e];
}

// robot.py:203-211 calc_state, robot.py:152-163 touch, env.py:240-244 object positions
__global__ void k_read_state(DevState st, double* __restrict__ joint, double* __restrict__ touch, double* __restrict__ objpose, int n_objects) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    const int N = st.N;
    if (joint) {
        for (int i = 0; i < 8; i++) joint[(size_t)e * 9 + i] = st.q[(size_t)i * N + e];
        joint[(size_t)e * 9 + 8] = -st.q[(size_t)8 * N + e];
    }
    if (touch) for (int k = 0; k < RR_NUM_TOUCH; k++) touch[(size_t)e * 4 + k] = st.touch[(size_t)k * N + e];
    if (objpose) for (int o = 0; o < n_objects; o++) for (int k = 0; k < 7; k++)
        objpose[((size_t)e * n_objects + o) * 7 + k] = st.obj[((size_t)o * 13 + k) * N + e];
}

__device__ __forceinline__ void mat_to_quat(const double* R, double* q) {
    double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        double s = sqrt(tr + 1.0);
        q[3] = s * 0.5; s = 0.5 / s;
        q[0] = (R[7] - R[5]) * s; q[1] = (R[2] - R[6]) * s; q[2] = (R[3] - R[1]) * s;
    } else {
        int i = R[0] < R[4] ? (R[4] < R[8] ? 2 : 1) : (R[0] < R[8] ? 2 : 0);
        int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        q[i] = s * 0.5; s = 0.5 / s;
        q[3] = (R[3 * k + j] - R[3 * j + k]) * s; q[j] = (R[3 * j + i] + R[3 * i + j]) * s; q[k] = (R[3 * k + i] + R[3 * i + k]) * s;
    }
}

__global__ void k_link_pose(DevState st, int link, double* __restrict__ out) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    double R[9], p[3];
    if (link < 0) { for (int k = 0; k < 9; k++) R[k] = (k % 4 == 0) ? 1.0 : 0.0; for (int k = 0; k < 3; k++) p[k] = c_robot.base_pos[k]; }
    else {
        for (int k = 0; k < 9; k++) R[k] = st.link_frames[((size_t)link * 12 + k) * st.N + e];
        for (int k = 0; k < 3; k++) p[k] = st.link_frames[((size_t)link * 12 + 9 + k) * st.N + e];
    }
    double q[4]; mat_to_quat(R, q);
    for (int k = 0; k < 3; k++) out[(size_t)e * 7 + k] = p[k];
    for (int k = 0; k < 4; k++) out[(size_t)e * 7 + 3 + k] = q[k];
}

// env.py:181-200 evaluateGoal
__global__ void k_score(DevState st, const double* __restrict__ goal, double* __restrict__ score, int n_objects) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    const double pos_const = -log(0.25) / 0.10;
    double s = 0.0;
    for (int o = 0; o < n_objects; o++) {
        double d2 = 0.0;
        for (int k = 0; k < 3; k++) {
            double d = goal[((size_t)e * n_objects + o) * 3 + k] - st.obj[((size_t)o * 13 + k) * st.N + e];
            d2 += d * d;
        }
        s += exp(-pos_const * sqrt(d2));
    }
    score[e] = s;
}

// world frames of all 22 scene shapes, f64 [N][22][12] (render input / test hook)
__global__ void k_shape_frames(DevState st, double* __restrict__ frames, int n_objects) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    const int N = st.N;
    for (int s = 0; s < RR_NSHAPES; s++) {
        double* f = frames + ((size_t)e * RR_NSHAPES + s) * 12;
        int body = c_robot.shape_body[s], link = c_robot.shape_link[s];
        if (body == 0) {
            if (link < 0) { for (int k = 0; k < 9; k++) f[k] = (k % 4 == 0) ? 1.0 : 0.0; for (int k = 0; k < 3; k++) f[9 + k] = c_robot.base_pos[k]; }
            else for (int k = 0; k < 12; k++) f[k] = st.link_frames[((size_t)link * 12 + k) * N + e];
        } else if (body == 1) {
            for (int k = 0; k < 9; k++) f[k] = c_robot.table_R[k];
            for (int k = 0; k < 3; k++) f[9 + k] = (link < 0) ? c_robot.table_pos[k] : c_robot.table_upper_pos[k];
        } else {
            int o = body - 2;
            double q[4], R[9];
            for (int k = 0; k < 4; k++) q[k] = st.obj[((size_t)o * 13 + 3 + k) * N + e];
            quat_to_mat(q, R);
            for (int k = 0; k < 9; k++) f[k] = R[k];
            for (int k = 0; k < 3; k++) f[9 + k] = st.obj[((size_t)o * 13 + k) * N + e];
        }
    }
}

// ------------------------------------------------------------------------------------------- host launchers
#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { ctx->err = std::string(#x) + ": " + cudaGetErrorString(_e); return -1; } } while (0)

static rr_ctx* g_phys_owner[64];
static int bind_physics(rr_ctx* ctx, cudaStream_t s) {
    rr_ctx*& owner = g_phys_owner[ctx->device & 63];
    if (owner != ctx) {
        CK(cudaDeviceSynchronize());          // kernels of the previous owner may still be reading the old constants
        CK(cudaMemcpyToSymbolAsync(c_robot, &ctx->robot, sizeof(RobotConst), 0, cudaMemcpyHostToDevice, s));
        owner = ctx;
    }
    return 0;
}
void rrk_physics_unbind(rr_ctx* ctx) { if (g_phys_owner[ctx->device & 63] == ctx) g_phys_owner[ctx->device & 63] = nullptr; }

int rrk_upload_constants(rr_ctx* ctx) {
    CK(cudaMemcpyToSymbol(c_robot, &ctx->robot, sizeof(RobotConst)));
    g_phys_owner[ctx->device & 63] = ctx;
    CK(cudaFuncSetAttribute(k_step_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(WarpMem) * WPB)));
    CK(cudaFuncSetAttribute(k_step_solve<SOLVE_MIN_BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SOLVE_SMEM_BYTES * SOLVE_WPB)));
    CK(cudaFuncSetAttribute(k_step_solve<SOLVE_MIN_BLOCKS_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SOLVE_SMEM_BYTES * SOLVE_WPB)));
    CK(cudaFuncSetAttribute(k_step_solve_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAVY_SMEM_BYTES));
    CK(cudaFuncSetAttribute(k_fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(FkMem) * WPB)));
    return 0;
}
static inline int nblk(int n, int b) { return (n + b - 1) / b; }

static int launch_fk(rr_ctx* ctx, cudaStream_t s) {
    if (bind_physics(ctx, s)) return -1;
    RR_LAUNCH(k_fk, nblk(ctx->N, WPB), WPB * 32, sizeof(FkMem) * WPB, s)(ctx->st);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
int rrk_reset(rr_ctx* ctx, const uint8_t* mask, cudaStream_t s) {
    if (rrk_drop_collide_ahead(ctx, s)) return -1;
    if (bind_physics(ctx, s)) return -1;
    RR_LAUNCH(k_reset, nblk(ctx->N, 128), 128, 0, s)(ctx->st, mask, ctx->n_objects, ctx->st.man_hdr[ctx->man_parity]);
    ctx->launches++;
    CK(cudaGetLastError());
    return launch_fk(ctx, s);
}
int rrk_set_object_pose(rr_ctx* ctx, const uint8_t* mask, int obj, const double* pq, cudaStream_t s) {
    if (rrk_drop_collide_ahead(ctx, s)) return -1;
    if (bind_physics(ctx, s)) return -1;
    RR_LAUNCH(k_set_object_pose, nblk(ctx->N, 128), 128, 0, s)(ctx->st, mask, obj, pq);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
__global__ void k_clear_manifolds(int N, const uint8_t* __restrict__ mask, int* __restrict__ man_hdr) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N) return;
    if (mask && !mask[e]) return;
    for (int k = 0; k < RR_MAX_MANIFOLDS; k++) { man_hdr[((size_t)e * RR_MAX_MANIFOLDS + k) * 2] = -1; man_hdr[((size_t)e * RR_MAX_MANIFOLDS + k) * 2 + 1] = 0; }
}
int rrk_clear_manifolds(rr_ctx* ctx, const uint8_t* mask, cudaStream_t s) {
    if (rrk_drop_collide_ahead(ctx, s)) return -1;
    RR_LAUNCH(k_clear_manifolds, nblk(ctx->N, 128), 128, 0, s)(ctx->N, mask, ctx->st.man_hdr[ctx->man_parity]);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
int rrk_set_state(rr_ctx* ctx, const double* q, const double* qd, const double* obj, cudaStream_t s) {
    if (rrk_drop_collide_ahead(ctx, s)) return -1;
    if (bind_physics(ctx, s)) return -1;
    if (rrk_clear_manifolds(ctx, nullptr, s) != 0) return -1;
    RR_LAUNCH(k_set_state, nblk(ctx->N, 128), 128, 0, s)(ctx->st, q, qd, obj);
    ctx->launches++;
    CK(cudaGetLastError());
    return launch_fk(ctx, s);
}
int rrk_get_state(rr_ctx* ctx, double* q, double* qd, double* obj, cudaStream_t s) {
    if (rrk_join_heavy(ctx, s)) return -1;
    RR_LAUNCH(k_get_state, nblk(ctx->N, 128), 128, 0, s)(ctx->st, q, qd, obj);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
// An editor of poses or manifolds on stream s invalidates the look-ahead narrowphase: wait for it (it reads what is about to be
// written), then forget it — the next step collides in line.
int rrk_drop_collide_ahead(rr_ctx* ctx, cudaStream_t s) {
    if (rrk_join_heavy(ctx, s)) return -1;
    if (!ctx->collide_ahead) return 0;
    CK(cudaStreamWaitEvent(s, ctx->ev_collided, 0));
    if (ctx->collide_ahead_split) CK(cudaStreamWaitEvent(s, ctx->ev_collided_heavy, 0));
    ctx->collide_ahead = false;
    return 0;
}
// the context's side streams and their events
#ifndef SIDE_PRIORITY_MASK
#define SIDE_PRIORITY_MASK 0     // measured on B200: a high-priority heavy kernel slows the main solver kernel by what it gains itself (6.80 vs 6.72 ms per bench step)
#endif
static int side_streams(rr_ctx* ctx) {
    if (ctx->side_stream) return 0;
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    int mask = SIDE_PRIORITY_MASK;          // bit 0: the heavy kernel's stream, bit 1: the look-ahead narrowphase's
    { const char* ea = getenv("RR_SIDE_PRIORITY"); if (ea && ea[0] >= '0' && ea[0] <= '3') mask = ea[0] - '0'; }
    CK(cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, (mask & 1) ? hi : 0));
    CK(cudaEventCreateWithFlags(&ctx->ev_solved, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_collided, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_dyn, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_heavy, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_collided_heavy, cudaEventDisableTiming));
    CK(cudaStreamCreateWithPriority(&ctx->side_stream2, cudaStreamNonBlocking, (mask & 2) ? hi : 0));
    return 0;
}
int rrk_join_heavy(rr_ctx* ctx, cudaStream_t s) {
    ctx->touched_since_step = true;
    if (!ctx->heavy_pending) return 0;
    CK(cudaStreamWaitEvent(s, ctx->ev_heavy, 0));
    ctx->heavy_pending = false;
    return 0;
}
int rrk_step(rr_ctx* ctx, const double* cmd, cudaStream_t s) {
    if (bind_physics(ctx, s)) return -1;
    if (rrk_join_heavy(ctx, s)) return -1;
    ctx->step_serial++;
    if (ctx->cfg.enable_contacts) {
        if (ctx->collide_ahead) {            // the previous step already collided at the poses it produced
            CK(cudaStreamWaitEvent(s, ctx->ev_collided, 0));
            if (ctx->collide_ahead_split) CK(cudaStreamWaitEvent(s, ctx->ev_collided_heavy, 0));
            ctx->man_parity ^= 1;
            ctx->collide_ahead = false;
        } else if (rrk_collide(ctx, s) != 0) return -1;
    }
    if (ctx->step_serial > 1) ctx->render_followed = ctx->rendered_since_step;      // (step_serial was just incremented)
    ctx->rendered_since_step = false; ctx->touched_since_step = false;
    const bool heavy = ctx->cfg.enable_contacts && (ctx->heavy_mode == 1 || (ctx->heavy_mode < 0 && ctx->render_followed));
    if ((heavy || ctx->eager_collide) && side_streams(ctx)) return -1;
    const int heavy_cap = heavy ? HEAVY_GRID * HEAVY_PER_BLOCK : 0;
    ctx->heavy_cap = heavy_cap;
    {
        ProfScope ps(ctx, 0, s);
        CK(cudaMemsetAsync(ctx->st.solve_order, 0, RR_SOLVE_CLASSES * sizeof(int), s));
        RR_LAUNCH(k_step_dyn, nblk(ctx->N, WPB), WPB * 32, sizeof(WarpMem) * WPB, s)(ctx->st, cmd, ctx->st.man_hdr[ctx->man_parity], ctx->st.man_pts[ctx->man_parity], ctx->cfg, ctx->n_objects);
        if (heavy) {         // enqueued before the main kernel: its blocks are placed first
            CK(cudaEventRecord(ctx->ev_dyn, s));
            CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_dyn, 0));
            RR_LAUNCH(k_step_solve_heavy, HEAVY_GRID, 32, HEAVY_SMEM_BYTES, ctx->side_stream)(ctx->st, ctx->st.man_hdr[ctx->man_parity], ctx->st.man_pts[ctx->man_parity], ctx->cfg, ctx->n_objects, heavy_cap, ctx->step_serial);
            CK(cudaEventRecord(ctx->ev_heavy, ctx->side_stream));
            ctx->heavy_pending = true;     // joined by whoever touches the state next (rrk_join_heavy); the eye-camera rasteriser waits per env instead
            ctx->launches++;
        }
        if (heavy) RR_LAUNCH(k_step_solve<SOLVE_MIN_BLOCKS_SPLIT>, nblk(ctx->N, SOLVE_WPB), SOLVE_WPB * 32, SOLVE_SMEM_BYTES * SOLVE_WPB, s)(ctx->st, ctx->st.man_hdr[ctx->man_parity], ctx->st.man_pts[ctx->man_parity], ctx->cfg, ctx->n_objects, heavy_cap);
        else RR_LAUNCH(k_step_solve<SOLVE_MIN_BLOCKS>, nblk(ctx->N, SOLVE_WPB), SOLVE_WPB * 32, SOLVE_SMEM_BYTES * SOLVE_WPB, s)(ctx->st, ctx->st.man_hdr[ctx->man_parity], ctx->st.man_pts[ctx->man_parity], ctx->cfg, ctx->n_objects, heavy_cap);
        ctx->launches += 2;
        CK(cudaGetLastError());
    }
    if (ctx->cfg.enable_contacts && ctx->eager_collide) {
        // the narrowphase of the NEXT step (it only reads the state): for the heavy kernel's envs behind that kernel on its stream,
        // for the others (or for all, without a heavy kernel) behind the main solver kernel on a stream of its own
        CK(cudaEventRecord(ctx->ev_solved, s));
        if (heavy) {
            CK(cudaStreamWaitEvent(ctx->side_stream2, ctx->ev_solved, 0));
            if (rrk_collide(ctx, ctx->side_stream2, false, 1) != 0) return -1;
            CK(cudaEventRecord(ctx->ev_collided, ctx->side_stream2));
            // (behind the main solver kernel as well, so that it does not compete with it for SMs)
            CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_solved, 0));
            if (rrk_collide(ctx, ctx->side_stream, false, 2) != 0) return -1;
            CK(cudaEventRecord(ctx->ev_collided_heavy, ctx->side_stream));
        } else {
            CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_solved, 0));
            if (rrk_collide(ctx, ctx->side_stream, false, 0) != 0) return -1;
            CK(cudaEventRecord(ctx->ev_collided, ctx->side_stream));
        }
        ctx->collide_ahead = true; ctx->collide_ahead_split = heavy;
    }
    return 0;
}
int rrk_read_state(rr_ctx* ctx, double* joint, double* touch, double* objpose, cudaStream_t s) {
    if (rrk_join_heavy(ctx, s)) return -1;
    if (bind_physics(ctx, s)) return -1;
    RR_LAUNCH(k_read_state, nblk(ctx->N, 128), 128, 0, s)(ctx->st, joint, touch, objpose, ctx->n_objects);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
int rrk_link_pose(rr_ctx* ctx, int link, double* out, cudaStream_t s) {
    if (rrk_join_heavy(ctx, s)) return -1;
    if (bind_physics(ctx, s)) return -1;
    RR_LAUNCH(k_link_pose, nblk(ctx->N, 128), 128, 0, s)(ctx->st, link, out);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
int rrk_score(rr_ctx* ctx, const double* goal, double* score, cudaStream_t s) {
    if (rrk_join_heavy(ctx, s)) return -1;
    if (bind_physics(ctx, s)) return -1;
    RR_LAUNCH(k_score, nblk(ctx->N, 128), 128, 0, s)(ctx->st, goal, score, ctx->n_objects);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
int rrk_shape_frames(rr_ctx* ctx, double* frames, cudaStream_t s) {
    if (rrk_join_heavy(ctx, s)) return -1;
    if (bind_physics(ctx, s)) return -1;
    RR_LAUNCH(k_shape_frames, nblk(ctx->N, 128), 128, 0, s)(ctx->st, frames, ctx->n_objects);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
// ------------------------------------------------------------------------------------------- inverse kinematics
// Replaces pybullet.calculateInverseKinematics(0, 7, pos, orn, maxNumIterations, residualThreshold) (env.py:372-375,
// 423-426): damped least squares on position + orientation of the gripper base link, seeded from the env's current
// joints, damping 0.5, 45 degree step clamp, stops on the position residual.  One thread per request: the problem is a
// 7-dof serial chain and is off the per-step path (it only runs when the cartesian request changes or a macro plan is built).
__global__ void k_ik(DevState st, const uint8_t* __restrict__ mask, const double* __restrict__ pos_quat, double* __restrict__ out,
                     int max_iter, double residual) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= st.N) return;
    const int N = st.N;
    double q[RR_NDOF];
    for (int d = 0; d < RR_NDOF; d++) q[d] = st.q[(size_t)d * N + e];
    if (mask && !mask[e]) { for (int d = 0; d < RR_NDOF; d++) out[(size_t)e * RR_NDOF + d] = q[d]; return; }
    const double* pq = pos_quat + (size_t)e * 7;
    const double tgt[3] = {pq[0], pq[1], pq[2]};
    double tq[4] = {pq[3], pq[4], pq[5], pq[6]};
    {
        double n = sqrt(tq[0] * tq[0] + tq[1] * tq[1] + tq[2] * tq[2] + tq[3] * tq[3]);
        if (n > 0) { for (int k = 0; k < 4; k++) tq[k] /= n; } else { tq[0] = tq[1] = tq[2] = 0; tq[3] = 1; }
    }
    const double damp = 0.5, max_angle = 45.0 * 3.14159265358979323846 / 180.0, PI = 3.14159265358979323846;
    for (int it = 0; it < max_iter; it++) {
        double Rw[8][9], pw[8][3];
        for (int i = 0; i <= 7; i++) {
            double R[9];
            for (int k = 0; k < 9; k++) R[k] = c_robot.rot0[i][k];
            if (c_robot.jtype[i] == 1) {
                double s, c; sincos(-q[c_robot.dof[i]], &s, &c);
                double ax = c_robot.axis[i][0], ay = c_robot.axis[i][1], az = c_robot.axis[i][2], t = 1.0 - c;
                double A[9] = {c + ax * ax * t, ax * ay * t - az * s, ax * az * t + ay * s,
                               ay * ax * t + az * s, c + ay * ay * t, ay * az * t - ax * s,
                               az * ax * t - ay * s, az * ay * t + ax * s, c + az * az * t};
                double T[9];
                for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) T[3 * a + b] = A[3 * a] * R[b] + A[3 * a + 1] * R[3 + b] + A[3 * a + 2] * R[6 + b];
                for (int k = 0; k < 9; k++) R[k] = T[k];
            }
            double r[3];
            for (int a = 0; a < 3; a++) r[a] = R[3 * a] * c_robot.dvec[i][0] + R[3 * a + 1] * c_robot.dvec[i][1] + R[3 * a + 2] * c_robot.dvec[i][2] + c_robot.evec[i][a];
            const int par = c_robot.parent[i];
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
                if (par >= 0) Rw[i][3 * a + b] = Rw[par][3 * a] * R[3 * b] + Rw[par][3 * a + 1] * R[3 * b + 1] + Rw[par][3 * a + 2] * R[3 * b + 2];
                else Rw[i][3 * a + b] = R[3 * b + a];
            }
            for (int a = 0; a < 3; a++) pw[i][a] = (par >= 0 ? pw[par][a] : c_robot.base_pos[a]) + Rw[i][3 * a] * r[0] + Rw[i][3 * a + 1] * r[1] + Rw[i][3 * a + 2] * r[2];
        }
        const double* ee = pw[7];
        double dp[3] = {tgt[0] - ee[0], tgt[1] - ee[1], tgt[2] - ee[2]};
        if (sqrt(dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2]) <= residual) break;
        double cq[4]; mat_to_quat(Rw[7], cq);
        const double ix = -cq[0], iy = -cq[1], iz = -cq[2], iw = cq[3];
        double dq[4] = {tq[3] * ix + tq[0] * iw + tq[1] * iz - tq[2] * iy,
                        tq[3] * iy + tq[1] * iw + tq[2] * ix - tq[0] * iz,
                        tq[3] * iz + tq[2] * iw + tq[0] * iy - tq[1] * ix,
                        tq[3] * iw - tq[0] * ix - tq[1] * iy - tq[2] * iz};
        double w = dq[3] > 1 ? 1 : (dq[3] < -1 ? -1 : dq[3]);
        double angle = 2.0 * acos(w);
        double s2 = 1.0 - dq[3] * dq[3];
        double axis[3] = {1, 0, 0};
        if (s2 >= 10 * 2.220446049250313e-16) { double s = 1.0 / sqrt(s2); axis[0] = dq[0] * s; axis[1] = dq[1] * s; axis[2] = dq[2] * s; }
        if (angle > PI) angle -= 2.0 * PI; else if (angle < -PI) angle += 2.0 * PI;
        double an = sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
        double k_ = an > 0 ? 1.0 / an : 1.0;
        double dS[6] = {dp[0], dp[1], dp[2], axis[0] * k_ * angle, axis[1] * k_ * angle, axis[2] * k_ * angle};
        double J[6][7];
        for (int k = 0; k < 7; k++) {
            double aw[3], piv[3];
            for (int a = 0; a < 3; a++) {
                aw[a] = Rw[k][3 * a] * c_robot.axis[k][0] + Rw[k][3 * a + 1] * c_robot.axis[k][1] + Rw[k][3 * a + 2] * c_robot.axis[k][2];
                piv[a] = pw[k][a] - (Rw[k][3 * a] * c_robot.evec[k][0] + Rw[k][3 * a + 1] * c_robot.evec[k][1] + Rw[k][3 * a + 2] * c_robot.evec[k][2]);
            }
            double r0 = ee[0] - piv[0], r1 = ee[1] - piv[1], r2 = ee[2] - piv[2];
            J[0][k] = aw[1] * r2 - aw[2] * r1; J[1][k] = aw[2] * r0 - aw[0] * r2; J[2][k] = aw[0] * r1 - aw[1] * r0;
            J[3][k] = aw[0]; J[4][k] = aw[1]; J[5][k] = aw[2];
        }
        double L[7][7], b[7];
        for (int i = 0; i < 7; i++) {
            for (int j = 0; j <= i; j++) { double s = 0; for (int r = 0; r < 6; r++) s += J[r][i] * J[r][j]; L[i][j] = s; }
            L[i][i] += damp;
            double s = 0; for (int r = 0; r < 6; r++) s += J[r][i] * dS[r]; b[i] = s;
        }
        for (int i = 0; i < 7; i++) for (int j = 0; j <= i; j++) {       // in-place Cholesky of the lower triangle
            double s = L[i][j];
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            L[i][j] = (i == j) ? sqrt(s) : s / L[j][j];
        }
        double y[7], dth[7];
        for (int i = 0; i < 7; i++) { double s = b[i]; for (int k = 0; k < i; k++) s -= L[i][k] * y[k]; y[i] = s / L[i][i]; }
        for (int i = 6; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < 7; k++) s -= L[k][i] * dth[k]; dth[i] = s / L[i][i]; }
        double mx = 0; for (int i = 0; i < 7; i++) mx = fmax(mx, fabs(dth[i]));
        if (mx > max_angle) for (int i = 0; i < 7; i++) dth[i] *= max_angle / mx;
        for (int i = 0; i < 7; i++) q[i] += dth[i];
    }
    for (int d = 0; d < RR_NDOF; d++) out[(size_t)e * RR_NDOF + d] = q[d];
}

int rrk_ik(rr_ctx* ctx, const uint8_t* mask, const double* pos_quat, double* out, int max_iter, double residual, cudaStream_t s) {
    if (rrk_join_heavy(ctx, s)) return -1;
    if (bind_physics(ctx, s)) return -1;
    RR_LAUNCH(k_ik, nblk(ctx->N, 64), 64, 0, s)(ctx->st, mask, pos_quat, out, max_iter, residual);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}
