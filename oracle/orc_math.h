/* TEST INFRASTRUCTURE — CPU oracle for the REALRobot2020 step path (see oracle/README.md).
 * Small fp64 vector / matrix / quaternion helpers used by the oracle only.
 * Nothing under real_robots_b200/ may include this file. */
#ifndef ORC_MATH_H
#define ORC_MATH_H
#include <cmath>
#include <cstring>

struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    double& operator[](int i) { return (&x)[i]; }
    double operator[](int i) const { return (&x)[i]; }
};
static inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }
static inline V3 operator*(V3 a, double s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3 operator*(double s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }
static inline V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
static inline V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
static inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
static inline V3 normalized(V3 a) { double n = norm(a); return n > 0 ? a * (1.0 / n) : a; }

struct M3 {
    double m[3][3];
    M3() { std::memset(m, 0, sizeof m); }
    static M3 identity() { M3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1; return r; }
    static M3 diag(V3 d) { M3 r; r.m[0][0] = d.x; r.m[1][1] = d.y; r.m[2][2] = d.z; return r; }
    V3 row(int i) const { return V3(m[i][0], m[i][1], m[i][2]); }
    V3 col(int j) const { return V3(m[0][j], m[1][j], m[2][j]); }
};
static inline V3 operator*(const M3& A, V3 v) { return V3(dot(A.row(0), v), dot(A.row(1), v), dot(A.row(2), v)); }
static inline M3 operator*(const M3& A, const M3& B) {
    M3 r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
        double s = 0; for (int k = 0; k < 3; k++) s += A.m[i][k] * B.m[k][j]; r.m[i][j] = s; }
    return r;
}
static inline M3 operator+(const M3& A, const M3& B) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = A.m[i][j] + B.m[i][j]; return r; }
static inline M3 operator-(const M3& A, const M3& B) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = A.m[i][j] - B.m[i][j]; return r; }
static inline M3 operator*(const M3& A, double s) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = A.m[i][j] * s; return r; }
static inline M3 transpose(const M3& A) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = A.m[j][i]; return r; }
static inline M3 skew(V3 v) { M3 r; r.m[0][1] = -v.z; r.m[0][2] = v.y; r.m[1][0] = v.z; r.m[1][2] = -v.x; r.m[2][0] = -v.y; r.m[2][1] = v.x; return r; }
static inline M3 outer(V3 a, V3 b) { M3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a[i] * b[j]; return r; }
/* rotation by angle th about unit axis a (Rodrigues) */
static inline M3 axis_angle(V3 a, double th) {
    double c = std::cos(th), s = std::sin(th), t = 1 - c;
    M3 r;
    r.m[0][0] = c + a.x * a.x * t;       r.m[0][1] = a.x * a.y * t - a.z * s; r.m[0][2] = a.x * a.z * t + a.y * s;
    r.m[1][0] = a.y * a.x * t + a.z * s; r.m[1][1] = c + a.y * a.y * t;       r.m[1][2] = a.y * a.z * t - a.x * s;
    r.m[2][0] = a.z * a.x * t - a.y * s; r.m[2][1] = a.z * a.y * t + a.x * s; r.m[2][2] = c + a.z * a.z * t;
    return r;
}

struct Q4 { double x, y, z, w; Q4() : x(0), y(0), z(0), w(1) {} Q4(double a, double b, double c, double d) : x(a), y(b), z(c), w(d) {} };
static inline Q4 qmul(Q4 a, Q4 b) {
    return Q4(a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
              a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x,
              a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z);
}
static inline Q4 qnormalize(Q4 q) { double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w); return Q4(q.x / n, q.y / n, q.z / n, q.w / n); }
static inline M3 quat_to_mat(Q4 q) {   /* btMatrix3x3::setRotation */
    double d = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    double s = 2.0 / d;
    double xs = q.x * s, ys = q.y * s, zs = q.z * s;
    double wx = q.w * xs, wy = q.w * ys, wz = q.w * zs;
    double xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
    double yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
    M3 r;
    r.m[0][0] = 1.0 - (yy + zz); r.m[0][1] = xy - wz;         r.m[0][2] = xz + wy;
    r.m[1][0] = xy + wz;         r.m[1][1] = 1.0 - (xx + zz); r.m[1][2] = yz - wx;
    r.m[2][0] = xz - wy;         r.m[2][1] = yz + wx;         r.m[2][2] = 1.0 - (xx + yy);
    return r;
}
static inline Q4 mat_to_quat(const M3& R) {   /* btMatrix3x3::getRotation */
    double tr = R.m[0][0] + R.m[1][1] + R.m[2][2];
    double t[4];
    if (tr > 0) {
        double s = std::sqrt(tr + 1.0);
        t[3] = s * 0.5; s = 0.5 / s;
        t[0] = (R.m[2][1] - R.m[1][2]) * s; t[1] = (R.m[0][2] - R.m[2][0]) * s; t[2] = (R.m[1][0] - R.m[0][1]) * s;
    } else {
        int i = R.m[0][0] < R.m[1][1] ? (R.m[1][1] < R.m[2][2] ? 2 : 1) : (R.m[0][0] < R.m[2][2] ? 2 : 0);
        int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = std::sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + 1.0);
        t[i] = s * 0.5; s = 0.5 / s;
        t[3] = (R.m[k][j] - R.m[j][k]) * s; t[j] = (R.m[j][i] + R.m[i][j]) * s; t[k] = (R.m[k][i] + R.m[i][k]) * s;
    }
    return Q4(t[0], t[1], t[2], t[3]);
}

/* spatial 6-vectors: top = angular, bottom = linear (Bullet's btSpatialMotionVector / btSpatialForceVector) */
struct S6 { V3 a, l; S6() {} S6(V3 A, V3 L) : a(A), l(L) {} };
static inline S6 operator+(S6 p, S6 q) { return S6(p.a + q.a, p.l + q.l); }
static inline S6 operator-(S6 p, S6 q) { return S6(p.a - q.a, p.l - q.l); }
static inline S6 operator*(S6 p, double s) { return S6(p.a * s, p.l * s); }
static inline double sdot(S6 p, S6 q) { return dot(p.a, q.a) + dot(p.l, q.l); }

/* 6x6 articulated inertia as blocks: force = [ torque ; linear ] = [A B; B^T... ] acting on motion [ang; lin]
 *   torque = A*ang + B*lin ;  force = Bt*ang + C*lin       (A, C symmetric; Bt = B^T) */
struct I6 { M3 A, B, C; };
static inline S6 operator*(const I6& I, S6 v) { return S6(I.A * v.a + I.B * v.l, transpose(I.B) * v.a + I.C * v.l); }
#endif
