// csb_math.cuh -- fixed-size FP64 math shared by the host planner and the sm_100a kernels.
//
// Everything is plain IEEE-754 double arithmetic evaluated left to right; the library is compiled with
// -fmad=false (device) and -ffp-contract=off (host) so that a*b+c is two correctly rounded operations on
// both sides and results do not depend on where a function runs.  std::min/std::max semantics of the
// reference (`(b<a)?b:a`, NaN-propagating on the first argument) are kept via cmin/cmax -- CUDA's
// fmin/fmax treat NaN differently.
//
// Reference types restated: Eigen::Quaterniond(Matrix3d), toRotationMatrix, q*v, Matrix3d::inverse
// (SURVEY.md App. B); matrix_utils.cpp:19-98,344-353; g2o se3quat.h:41-362, se3_ops.hpp:28-48.
#pragma once
#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define CSB_HD __host__ __device__ __forceinline__
#else
#define CSB_HD inline
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace csb {

struct V2 { double x, y; };
struct V3 { double x, y, z; };
struct M3 { double m[9]; };  // row-major
struct Quat { double w, x, y, z; };
struct SE3 { Quat r; V3 t; };
struct Cube { SE3 pose; V3 scale; };

CSB_HD double cmin(double a, double b) { return (b < a) ? b : a; }
CSB_HD double cmax(double a, double b) { return (a < b) ? b : a; }

CSB_HD V2 sub(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
CSB_HD double norm2(V2 a) { return sqrt(a.x * a.x + a.y * a.y); }
CSB_HD double dist2(V2 a, V2 b) { return norm2(sub(a, b)); }
CSB_HD V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
CSB_HD V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
CSB_HD V3 scl(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
CSB_HD double norm3(V3 a) { return sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }
CSB_HD V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

CSB_HD V3 mul(const M3& A, V3 v) {
    return {(A.m[0] * v.x + A.m[1] * v.y) + A.m[2] * v.z, (A.m[3] * v.x + A.m[4] * v.y) + A.m[5] * v.z, (A.m[6] * v.x + A.m[7] * v.y) + A.m[8] * v.z};
}
CSB_HD M3 mul(const M3& A, const M3& B) {
    M3 C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C.m[i * 3 + j] = (A.m[i * 3] * B.m[j] + A.m[i * 3 + 1] * B.m[3 + j]) + A.m[i * 3 + 2] * B.m[6 + j];
    return C;
}
CSB_HD M3 identity3() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }

CSB_HD double cof3(const M3& a, int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return a.m[i1 * 3 + j1] * a.m[i2 * 3 + j2] - a.m[i1 * 3 + j2] * a.m[i2 * 3 + j1];
}
// Eigen Matrix3d::inverse(): cofactors of column 0 give the determinant; result(i,j) = cofactor(j,i)/det
CSB_HD M3 inverse3(const M3& a) {
    double c0 = cof3(a, 0, 0), c1 = cof3(a, 1, 0), c2 = cof3(a, 2, 0);
    double det = (c0 * a.m[0] + c1 * a.m[3]) + c2 * a.m[6];
    double invdet = 1.0 / det;
    M3 r;
    r.m[0] = c0 * invdet; r.m[1] = c1 * invdet; r.m[2] = c2 * invdet;
    r.m[3] = cof3(a, 0, 1) * invdet; r.m[4] = cof3(a, 1, 1) * invdet; r.m[5] = cof3(a, 2, 1) * invdet;
    r.m[6] = cof3(a, 0, 2) * invdet; r.m[7] = cof3(a, 1, 2) * invdet; r.m[8] = cof3(a, 2, 2) * invdet;
    return r;
}

CSB_HD Quat quat_from_rot(const M3& R) {
    Quat q;
    double t = (R.m[0] + R.m[4]) + R.m[8];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R.m[7] - R.m[5]) * t;
        q.y = (R.m[2] - R.m[6]) * t;
        q.z = (R.m[3] - R.m[1]) * t;
    } else {
        // i = index of the largest diagonal entry, (j, k) its cyclic successors; written out per case so that nothing is
        // dynamically indexed (keeps the matrix in registers on the device)
        int i = 0;
        if (R.m[4] > R.m[0]) i = 1;
        if (R.m[8] > (i == 0 ? R.m[0] : R.m[4])) i = 2;
        if (i == 0) {  // j = 1, k = 2
            t = sqrt(R.m[0] - R.m[4] - R.m[8] + 1.0);
            q.x = 0.5 * t;
            t = 0.5 / t;
            q.w = (R.m[7] - R.m[5]) * t;
            q.y = (R.m[3] + R.m[1]) * t;
            q.z = (R.m[6] + R.m[2]) * t;
        } else if (i == 1) {  // j = 2, k = 0
            t = sqrt(R.m[4] - R.m[8] - R.m[0] + 1.0);
            q.y = 0.5 * t;
            t = 0.5 / t;
            q.w = (R.m[2] - R.m[6]) * t;
            q.z = (R.m[7] + R.m[5]) * t;
            q.x = (R.m[1] + R.m[3]) * t;
        } else {  // j = 0, k = 1
            t = sqrt(R.m[8] - R.m[0] - R.m[4] + 1.0);
            q.z = 0.5 * t;
            t = 0.5 / t;
            q.w = (R.m[3] - R.m[1]) * t;
            q.x = (R.m[2] + R.m[6]) * t;
            q.y = (R.m[5] + R.m[7]) * t;
        }
    }
    return q;
}
CSB_HD M3 quat_to_rot(const Quat& q) {
    double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3 r;
    r.m[0] = 1 - (tyy + tzz); r.m[1] = txy - twz; r.m[2] = txz + twy;
    r.m[3] = txy + twz; r.m[4] = 1 - (txx + tzz); r.m[5] = tyz - twx;
    r.m[6] = txz - twy; r.m[7] = tyz + twx; r.m[8] = 1 - (txx + tyy);
    return r;
}
CSB_HD V3 quat_rot(const Quat& q, V3 v) {
    V3 qv{q.x, q.y, q.z};
    V3 uv = cross(qv, v);
    uv = add(uv, uv);
    V3 c = cross(qv, uv);
    return {(v.x + q.w * uv.x) + c.x, (v.y + q.w * uv.y) + c.y, (v.z + q.w * uv.z) + c.z};
}
CSB_HD Quat quat_mul(const Quat& a, const Quat& b) {
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return r;
}
CSB_HD Quat quat_conj(const Quat& q) { return {q.w, -q.x, -q.y, -q.z}; }
CSB_HD void quat_normalize(Quat& q) {
    double n = sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
}

// matrix_utils.cpp:38-49
CSB_HD void quat_to_euler_zyx(const Quat& q, double& roll, double& pitch, double& yaw) {
    double qw = q.w, qx = q.x, qy = q.y, qz = q.z;
    roll = atan2(2 * (qw * qx + qy * qz), 1 - 2 * (qx * qx + qy * qy));
    pitch = asin(2 * (qw * qy - qz * qx));
    yaw = atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz));
}
// matrix_utils.cpp:81-96
CSB_HD M3 euler_zyx_to_rot(double roll, double pitch, double yaw) {
    double cp = cos(pitch), sp = sin(pitch), sr = sin(roll), cr = cos(roll), sy = sin(yaw), cy = cos(yaw);
    M3 R;
    R.m[0] = cp * cy; R.m[1] = (sr * sp * cy) - (cr * sy); R.m[2] = (cr * sp * cy) + (sr * sy);
    R.m[3] = cp * sy; R.m[4] = (sr * sp * sy) + (cr * cy); R.m[5] = (cr * sp * sy) - (sr * cy);
    R.m[6] = -sp; R.m[7] = sr * cp; R.m[8] = cr * cp;
    return R;
}
// matrix_utils.cpp:344-353
CSB_HD double normalize_to_pi(double a) {
    if (a > M_PI / 2) return a - M_PI;
    else if (a < -M_PI / 2) return a + M_PI;
    else return a;
}


// ---- deterministic atan2 -----------------------------------------------------------------------------------------
// The reference calls libm's atan2 (object_3d_util.cpp:273,490,573,702; box_proposal_detail.cpp:313).  libm results are
// not reproducible across implementations (glibc vs CUDA differ in the last ulp), and the proposal ranking contains
// structural near-ties (object yaw samples 90 degrees apart describe the same cuboid with vp1/vp2 swapped), so a 1-ulp
// difference flips ranking indices.  Both the oracle and the device path therefore evaluate atan2 with the same
// specified algorithm: the classic argument-reduction + degree-11 odd polynomial scheme of Sun's fdlibm (s_atan.c /
// e_atan2.c, error < 1 ulp), written with plain IEEE-754 double operations only, so that it is bit-reproducible anywhere
// FMA contraction is off.  The oracle carries its own copy (oracle/oracle_math.h); tests check both against glibc's atan2 to <= 1 ulp.

CSB_HD double det_atan(double x, unsigned hx_bits) {
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

CSB_HD double det_atan2(double y, double x) {
    const double tiny = 1.0e-300, pi_o_4 = 7.8539816339744827900E-01, pi_o_2 = 1.5707963267948965580E+00, pi = 3.1415926535897931160E+00,
                 pi_lo = 1.2246467991473531772E-16;
    if (x != x || y != y) return x + y;
    unsigned hx, lx, hy, ly;
    #if defined(__CUDA_ARCH__)
    hx = (unsigned)__double2hiint(x); lx = (unsigned)__double2loint(x); hy = (unsigned)__double2hiint(y); ly = (unsigned)__double2loint(y);
#else
    { unsigned long long ux, uy; memcpy(&ux, &x, 8); memcpy(&uy, &y, 8); hx = (unsigned)(ux >> 32); lx = (unsigned)ux; hy = (unsigned)(uy >> 32); ly = (unsigned)uy; }
#endif
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
        #if defined(__CUDA_ARCH__)
        hq = (unsigned)__double2hiint(q); lq = (unsigned)__double2loint(q);
#else
        { unsigned long long uq; memcpy(&uq, &q, 8); hq = (unsigned)(uq >> 32); lq = (unsigned)uq; }
#endif
        (void)lq;
        z = det_atan(q, hq);
    }
    if (m == 0) return z;
    if (m == 1) return -z;
    if (m == 2) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}

// Six det_atan2 evaluations at once, for operands that take det_atan2's generic path (finite, non-zero, |y/x| in [2^-29, 2^66), no
// exponent shortcut): the same operations in the same order per element, but with the interval choice of det_atan expressed as
// selects of (numerator, denominator) feeding ONE division, so that the six dependency chains are independent straight-line code the
// compiler can interleave.  Returns false (and computes nothing) if any element needs one of det_atan2's special paths; the caller then
// falls back to det_atan2 element by element.  Bit-identical to det_atan2 (tests/test_proposal_gpu.py::test_batched_atan2).
#if defined(__CUDACC__)
__device__ __forceinline__ bool det_atan2_x6(const double* y, const double* x, double* out) {
    const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01, 9.82793723247329054082e-01, 1.57079632679489655800e+00};
    const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17, 1.39033110312309984516e-17, 6.12323399573676603587e-17};
    const double aT[11] = {3.33333333333329318027e-01, -1.99999999998764832476e-01, 1.42857142725034663711e-01, -1.11111104054623557880e-01,
                           9.09088713343650656196e-02, -7.69187620504482999495e-02, 6.66107313738753120669e-02, -5.83357013379057348645e-02,
                           4.97687799461593236017e-02, -3.65315727442169155270e-02, 1.62858201153657823623e-02};
    const double pi = 3.1415926535897931160E+00, pi_lo = 1.2246467991473531772E-16;
    bool ok = true;
    double q[6];
    int m[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const unsigned hx = (unsigned)__double2hiint(x[i]), lx = (unsigned)__double2loint(x[i]);
        const unsigned hy = (unsigned)__double2hiint(y[i]), ly = (unsigned)__double2loint(y[i]);
        const unsigned ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
        m[i] = (int)((hy >> 31) & 1u) | (int)((hx >> 30) & 2u);
        const int k = ((int)iy - (int)ix) >> 20;
        // NaN / inf (exponent all ones), zeros, and the exponent shortcuts of det_atan2
        ok = ok && (ix < 0x7ff00000u) && (iy < 0x7ff00000u) && ((ix | lx) != 0) && ((iy | ly) != 0) && (k <= 60) && (k >= -60);
    }
    if (!ok) return false;
#pragma unroll
    for (int i = 0; i < 6; i++) q[i] = fabs(y[i] / x[i]);
    double num[6], den[6], hi[6], lo[6];
    bool small[6];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const unsigned ixq = (unsigned)__double2hiint(q[i]) & 0x7fffffffu;
        ok = ok && (ixq < 0x44100000u) && (ixq >= 0x3e200000u);
        const double v = q[i];
        small[i] = ixq < 0x3fdc0000u;
        const bool i0 = !small[i] && ixq < 0x3fe60000u, i1 = !small[i] && !i0 && ixq < 0x3ff30000u, i2 = !small[i] && !i0 && !i1 && ixq < 0x40038000u;
        num[i] = small[i] ? v : i0 ? (2.0 * v - 1.0) : i1 ? (v - 1.0) : i2 ? (v - 1.5) : -1.0;
        den[i] = small[i] ? 1.0 : i0 ? (2.0 + v) : i1 ? (v + 1.0) : i2 ? (1.0 + 1.5 * v) : v;
        hi[i] = i0 ? atanhi[0] : i1 ? atanhi[1] : i2 ? atanhi[2] : atanhi[3];
        lo[i] = i0 ? atanlo[0] : i1 ? atanlo[1] : i2 ? atanlo[2] : atanlo[3];
    }
    if (!ok) return false;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        const double xr = num[i] / den[i];  // den == 1 in the small interval: xr == q exactly
        const double z = xr * xr;
        const double w = z * z;
        const double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
        const double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
        const double p = xr * (s1 + s2);
        const double zz = small[i] ? (xr - p) : (hi[i] - ((p - lo[i]) - xr));
        const double r2 = pi - (zz - pi_lo), r3 = (zz - pi_lo) - pi;
        out[i] = (m[i] == 0) ? zz : (m[i] == 1) ? -zz : (m[i] == 2) ? r2 : r3;
    }
    return true;
}
#endif

// ---- SE(3), g2o se3quat.h ---------------------------------------------------------------------
CSB_HD void se3_normalize(SE3& s) {  // :346-351
    if (s.r.w < 0) { s.r.w = -s.r.w; s.r.x = -s.r.x; s.r.y = -s.r.y; s.r.z = -s.r.z; }
    quat_normalize(s.r);
}
CSB_HD SE3 se3_make(const Quat& q, V3 t) { SE3 s; s.r = q; s.t = t; se3_normalize(s); return s; }
CSB_HD SE3 se3_from_vec7(const double* v) { return se3_make(Quat{v[6], v[3], v[4], v[5]}, V3{v[0], v[1], v[2]}); }
CSB_HD SE3 se3_mul(const SE3& a, const SE3& b) {  // :110-116
    SE3 r;
    r.t = add(a.t, quat_rot(a.r, b.t));
    r.r = quat_mul(a.r, b.r);
    se3_normalize(r);
    return r;
}
CSB_HD SE3 se3_inverse(const SE3& a) {  // :129-134
    SE3 r;
    r.r = quat_conj(a.r);
    r.t = quat_rot(r.r, V3{a.t.x * -1., a.t.y * -1., a.t.z * -1.});
    return r;
}
CSB_HD M3 skew(V3 v) { return M3{{0, -v.z, v.y, v.z, 0, -v.x, -v.y, v.x, 0}}; }
CSB_HD void se3_log(const SE3& s, double* res) {  // :230-267
    M3 R = quat_to_rot(s.r);
    double d = 0.5 * (R.m[0] + R.m[4] + R.m[8] - 1);
    V3 dR{R.m[7] - R.m[5], R.m[2] - R.m[6], R.m[3] - R.m[1]};
    V3 omega;
    double c;
    if (d > 0.99999) {
        omega = scl(0.5, dR);
        c = (1. / 12.);
    } else {
        double theta = acos(d);
        omega = scl(theta / (2 * sqrt(1 - d * d)), dR);
        c = (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
    }
    M3 Om = skew(omega);
    M3 Om2 = mul(Om, Om);
    M3 I = identity3();
    M3 Vinv;
    for (int i = 0; i < 9; i++) Vinv.m[i] = I.m[i] - 0.5 * Om.m[i] + c * Om2.m[i];
    V3 ups = mul(Vinv, s.t);
    res[0] = omega.x; res[1] = omega.y; res[2] = omega.z; res[3] = ups.x; res[4] = ups.y; res[5] = ups.z;
}
CSB_HD SE3 se3_exp(const double* u) {  // :275-323
    V3 omega{u[0], u[1], u[2]}, upsilon{u[3], u[4], u[5]};
    double theta = norm3(omega);
    M3 Om = skew(omega);
    M3 Om2 = mul(Om, Om);
    M3 I = identity3();
    M3 R, V;
    if (theta < 0.00001) {
        for (int i = 0; i < 9; i++) R.m[i] = I.m[i] + Om.m[i] + Om2.m[i];
        V = R;
    } else {
        double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / (pow(theta, 3));
        for (int i = 0; i < 9; i++) R.m[i] = I.m[i] + a * Om.m[i] + b * Om2.m[i];
        for (int i = 0; i < 9; i++) V.m[i] = I.m[i] + b * Om.m[i] + c * Om2.m[i];
    }
    return se3_make(quat_from_rot(R), mul(V, upsilon));
}

}  // namespace csb
