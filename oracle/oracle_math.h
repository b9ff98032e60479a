// TEST INFRASTRUCTURE ONLY -- CPU oracle for the CubeSLAM hot path.
// Nothing under oracle/ is part of the shipped product: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, link or call it.
//
// Small fixed-size FP64 math used by the oracle.  Eigen is not available in this
// container, so the Eigen-internal algorithms the reference relies on are restated
// here.  Summation order is left-to-right everywhere; all
// translation units are compiled with -O2 -ffp-contract=off so that +,-,*,/,sqrt
// are single correctly-rounded IEEE-754 operations (no FMA contraction).
//
// PARITY STATUS: pinned through its users -- the online-mode replay of the reference's own TUM sequence against its committed output
// files (tests/test_reference_replay.py; see oracle_proposal.cpp / oracle_ba.cpp) and the ray_plane_interact worked example printed in
// detect_3d_cuboid/src/object_3d_util.cpp:884-905 (tests/test_oracle_golden.py).
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

struct V2 { double x, y; };
struct V3 { double x, y, z; };
struct V4 { double a[4]; };
struct M3 { double m[9]; /* row-major */ double& operator()(int r, int c) { return m[r * 3 + c]; } double operator()(int r, int c) const { return m[r * 3 + c]; } };
struct M4 { double m[16]; double& operator()(int r, int c) { return m[r * 4 + c]; } double operator()(int r, int c) const { return m[r * 4 + c]; } };
struct Quat { double w, x, y, z; };

static inline V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
static inline V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
static inline V2 operator*(double s, V2 a) { return {s * a.x, s * a.y}; }
static inline double norm(V2 a) { return std::sqrt(a.x * a.x + a.y * a.y); }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline double norm(V3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

static inline M3 mul(const M3& A, const M3& B) {
    M3 C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C(i, j) = (A(i, 0) * B(0, j) + A(i, 1) * B(1, j)) + A(i, 2) * B(2, j);
    return C;
}
static inline V3 mul(const M3& A, V3 v) {
    return {(A(0, 0) * v.x + A(0, 1) * v.y) + A(0, 2) * v.z, (A(1, 0) * v.x + A(1, 1) * v.y) + A(1, 2) * v.z,
            (A(2, 0) * v.x + A(2, 1) * v.y) + A(2, 2) * v.z};
}
static inline M3 transpose(const M3& A) {
    M3 T;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T(i, j) = A(j, i);
    return T;
}
static inline M3 identity3() { M3 I; std::memset(I.m, 0, sizeof I.m); I(0, 0) = I(1, 1) = I(2, 2) = 1; return I; }

// Eigen Matrix3d::inverse(): cofactor / determinant (Eigen/src/LU/InverseImpl.h, size-3 path;
// recalled).  result(i,j) = cofactor(j,i) / det.
static inline double cof3(const M3& m, int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
}
static inline M3 inverse3(const M3& m) {
    double c0 = cof3(m, 0, 0), c1 = cof3(m, 1, 0), c2 = cof3(m, 2, 0);
    double det = (c0 * m(0, 0) + c1 * m(1, 0)) + c2 * m(2, 0);
    double invdet = 1.0 / det;
    M3 r;
    r(0, 0) = c0 * invdet; r(0, 1) = c1 * invdet; r(0, 2) = c2 * invdet;
    r(1, 0) = cof3(m, 0, 1) * invdet; r(1, 1) = cof3(m, 1, 1) * invdet; r(1, 2) = cof3(m, 2, 1) * invdet;
    r(2, 0) = cof3(m, 0, 2) * invdet; r(2, 1) = cof3(m, 1, 2) * invdet; r(2, 2) = cof3(m, 2, 2) * invdet;
    return r;
}
// generic 4x4 inverse by cofactors (only feeds cam_pose.projectionMatrix, which the
// scored path never reads: object_3d_util.cpp:941-1011 takes it but does not use it).
static inline M4 inverse4(const M4& a) {
    const double* m = a.m;
    double inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    det = 1.0 / det;
    M4 r;
    for (int i = 0; i < 16; i++) r.m[i] = inv[i] * det;
    return r;
}

// Eigen Quaterniond(Matrix3d) (Eigen/src/Geometry/Quaternion.h quaternionbase_assign_impl)
static inline Quat quat_from_rot(const M3& m) {
    Quat q;
    double t = (m(0, 0) + m(1, 1)) + m(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m(2, 1) - m(1, 2)) * t;
        q.y = (m(0, 2) - m(2, 0)) * t;
        q.z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (m(k, j) - m(j, k)) * t;
        v[j] = (m(j, i) + m(i, j)) * t;
        v[k] = (m(k, i) + m(i, k)) * t;
        q.x = v[0]; q.y = v[1]; q.z = v[2];
    }
    return q;
}
// Eigen QuaternionBase::toRotationMatrix
static inline M3 quat_to_rot(const Quat& q) {
    double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3 r;
    r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
    r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
    return r;
}
// Eigen quaternion * vector:  uv = q.vec x v; uv += uv; v + w*uv + q.vec x uv
static inline V3 quat_rot(const Quat& q, V3 v) {
    V3 qv{q.x, q.y, q.z};
    V3 uv = cross(qv, v);
    uv = uv + uv;
    V3 c = cross(qv, uv);
    return {(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}
// Eigen quaternion product (scalar path, quat_product<Arch::Target,...>)
static inline Quat quat_mul(const Quat& a, const Quat& b) {
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
static inline Quat quat_conj(const Quat& q) { return {q.w, -q.x, -q.y, -q.z}; }
// coeffs order in Eigen is (x,y,z,w); squaredNorm sums in that order
static inline void quat_normalize(Quat& q) {
    double n = std::sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

// detect_3d_cuboid/src/matrix_utils.cpp:19-33
static inline Quat zyx_euler_to_quat(double roll, double pitch, double yaw) {
    double sy = std::sin(yaw * 0.5), cy = std::cos(yaw * 0.5);
    double sp = std::sin(pitch * 0.5), cp = std::cos(pitch * 0.5);
    double sr = std::sin(roll * 0.5), cr = std::cos(roll * 0.5);
    Quat q;
    q.w = cr * cp * cy + sr * sp * sy;
    q.x = sr * cp * cy - cr * sp * sy;
    q.y = cr * sp * cy + sr * cp * sy;
    q.z = cr * cp * sy - sr * sp * cy;
    return q;
}
// detect_3d_cuboid/src/matrix_utils.cpp:38-49
static inline void quat_to_euler_zyx(const Quat& q, double& roll, double& pitch, double& yaw) {
    double qw = q.w, qx = q.x, qy = q.y, qz = q.z;
    roll = std::atan2(2 * (qw * qx + qy * qz), 1 - 2 * (qx * qx + qy * qy));
    pitch = std::asin(2 * (qw * qy - qz * qx));
    yaw = std::atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz));
}
// detect_3d_cuboid/src/matrix_utils.cpp:81-96
static inline M3 euler_zyx_to_rot(double roll, double pitch, double yaw) {
    double cp = std::cos(pitch), sp = std::sin(pitch), sr = std::sin(roll), cr = std::cos(roll), sy = std::sin(yaw), cy = std::cos(yaw);
    M3 R;
    R(0, 0) = cp * cy; R(0, 1) = (sr * sp * cy) - (cr * sy); R(0, 2) = (cr * sp * cy) + (sr * sy);
    R(1, 0) = cp * sy; R(1, 1) = (sr * sp * sy) + (cr * cy); R(1, 2) = (cr * sp * sy) - (sr * cy);
    R(2, 0) = -sp; R(2, 1) = sr * cp; R(2, 2) = cr * cp;
    return R;
}
// detect_3d_cuboid/src/matrix_utils.cpp:344-353
static inline double normalize_to_pi(double angle) {
    if (angle > M_PI / 2) return angle - M_PI;
    else if (angle < -M_PI / 2) return angle + M_PI;
    else return angle;
}


// ---- deterministic atan2 -----------------------------------------------------------------------------------------
// The reference calls libm's atan2 (object_3d_util.cpp:273,490,573,702; box_proposal_detail.cpp:313).  libm results are
// not reproducible across implementations (glibc vs CUDA differ in the last ulp), and the proposal ranking contains
// structural near-ties (object yaw samples 90 degrees apart describe the same cuboid with vp1/vp2 swapped), so a 1-ulp
// difference flips ranking indices.  Both the oracle and the device path therefore evaluate atan2 with the same
// specified algorithm: the classic argument-reduction + degree-11 odd polynomial scheme of Sun's fdlibm (s_atan.c /
// e_atan2.c, error < 1 ulp), written with plain IEEE-754 double operations only, so that it is bit-reproducible anywhere
// FMA contraction is off.  tests/test_oracle_golden.py checks it against glibc's atan2 to <= 1 ulp.

static inline double det_atan(double x, unsigned hx_bits) {
    const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01, 1.57079632679489655800e+00};
    const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17, 6.12323399573676603587e-17};
    const double aT[11] = {3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01, -1.11111104054623557880e-01,
                           9.09088713343650656196e-02, -7.69187620504482999495e-02, 6.66107313738753120669e-02, -5.83357013379057348645e-02,
                           4.97687799461593236017e-02, -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    // x >= 0 here (called with |y/x|); hx_bits = high word of x
    const unsigned ix = hx_bits & 0x7fffffffu;
    int id;
    if (ix >= 0x44100000u) {  // |x| >= 2^66 (or inf/nan: callers filter those)
        return atanhi[3] + atanlo[3];
    }
    if (ix < 0x3fdc0000u) {  // |x| < 0.4375
        if (ix < 0x3e200000u) return x;  // |x| < 2^-29
        id = -1;
    } else if (ix < 0x3ff30000u) {  // |x| < 1.1875
        if (ix < 0x3fe60000u) { id = 0; x = (2.0 * x - 1.0) / (2.0 + x); }  // 7/16 <= |x| < 11/16
        else { id = 1; x = (x - 1.0) / (x + 1.0); }                          // 11/16 <= |x| < 19/16
    } else if (ix < 0x40038000u) { id = 2; x = (x - 1.5) / (1.0 + 1.5 * x); }  // |x| < 2.4375
    else { id = 3; x = -1.0 / x; }
    double z = x * x;
    double w = z * z;
    double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    return atanhi[id] - ((x * (s1 + s2) - atanlo[id]) - x);
}

static inline double det_atan2(double y, double x) {
    const double tiny = 1.0e-300, pi_o_4 = 7.8539816339744827900E-01, pi_o_2 = 1.5707963267948965580E+00, pi = 3.1415926535897931160E+00,
                 pi_lo = 1.2246467991473531772E-16;
    if (x != x || y != y) return x + y;
    unsigned hx, lx, hy, ly;
    { unsigned long long ux, uy; std::memcpy(&ux, &x, 8); std::memcpy(&uy, &y, 8); hx = (unsigned)(ux >> 32); lx = (unsigned)ux; hy = (unsigned)(uy >> 32); ly = (unsigned)uy; }
    const unsigned ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    const int m = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);  // 2*sign(x) + sign(y)
    if ((iy | ly) == 0) {  // y = 0
        if (m == 0 || m == 1) return y;
        return (m == 2) ? pi + tiny : -pi - tiny;
    }
    if ((ix | lx) == 0) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;  // x = 0
    if (ix == 0x7ff00000u) {  // x = inf
        if (iy == 0x7ff00000u) {
            if (m == 0) return pi_o_4 + tiny;
            if (m == 1) return -pi_o_4 - tiny;
            if (m == 2) return 3.0 * pi_o_4 + tiny;
            return -3.0 * pi_o_4 - tiny;
        }
        if (m == 0) return 0.0;
        if (m == 1) return -0.0;
        if (m == 2) return pi + tiny;
        return -pi - tiny;
    }
    if (iy == 0x7ff00000u) return (hy >> 31) ? -pi_o_2 - tiny : pi_o_2 + tiny;  // y = inf
    const int k = ((int)iy - (int)ix) >> 20;
    double z;
    if (k > 60) z = pi_o_2 + 0.5 * pi_lo;            // |y/x| > 2^60
    else if ((hx >> 31) && k < -60) z = 0.0;          // |y|/x < -2^60
    else {
        double q = fabs(y / x);
        unsigned hq, lq;
        { unsigned long long uq; std::memcpy(&uq, &q, 8); hq = (unsigned)(uq >> 32); lq = (unsigned)uq; }
        (void)lq;
        z = det_atan(q, hq);
    }
    if (m == 0) return z;
    if (m == 1) return -z;
    if (m == 2) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}

}  // namespace orc
