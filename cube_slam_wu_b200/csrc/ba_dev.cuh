// ba_dev.cuh -- device restatement of the g2o value types / edge residuals used by the BA kernels (ba.cu, ba_solve.cu).
// Reference: object_slam/include/object_slam/g2o_Object.h:23-292, Thirdparty/g2o/g2o/types/types_six_dof_expmap.h:59-99.
#pragma once
#include "csb_math.cuh"

namespace csb {

// ---- g2o::cuboid (g2o_Object.h) ----------------------------------------------------------------
__device__ __forceinline__ Cube cube_from_vec10(const double* v) {  // fromVector :45-48 (no normalisation)
    Cube c;
    c.pose.r = Quat{v[6], v[3], v[4], v[5]};
    c.pose.t = V3{v[0], v[1], v[2]};
    c.scale = V3{v[7], v[8], v[9]};
    return c;
}
__device__ __forceinline__ Cube cube_exp_update(const Cube& c, const double* u) {  // :57-63
    Cube r;
    r.pose = se3_mul(c.pose, se3_exp(u));
    r.scale = V3{c.scale.x + u[6], c.scale.y + u[7], c.scale.z + u[8]};
    return r;
}
__device__ __forceinline__ void cube_log_error(const Cube& self, const Cube& newone, double* res) {  // :66-73
    SE3 pose_diff = se3_mul(se3_inverse(newone.pose), self.pose);
    se3_log(pose_diff, res);
    res[6] = self.scale.x - newone.scale.x; res[7] = self.scale.y - newone.scale.y; res[8] = self.scale.z - newone.scale.z;
}
// min_log_error :76-101 with rotate_cuboid :104-114 (yaw -90, 0, 90, 180 deg; x/y scales swap at +-90)
__device__ __forceinline__ void cube_min_log_error(const Cube& self, const Cube& newone, double* res) {
    double best_n = 0;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        double yaw_angle = (double)(i - 1) * M_PI / 2.0;
        Cube rc;
        SE3 rot = se3_make(Quat{cos(yaw_angle * 0.5), 0, 0, sin(yaw_angle * 0.5)}, V3{0, 0, 0});
        rc.pose = se3_mul(newone.pose, rot);
        rc.scale = newone.scale;
        if ((yaw_angle == M_PI / 2.0) || (yaw_angle == -M_PI / 2.0) || (yaw_angle == 3 * M_PI / 2.0)) { double t = rc.scale.x; rc.scale.x = rc.scale.y; rc.scale.y = t; }
        double e[9];
        cube_log_error(self, rc, e);
        double s = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) s += e[k] * e[k];
        double n = sqrt(s);
        if (i == 0 || n < best_n) {  // Eigen minCoeff: first minimum
            best_n = n;
#pragma unroll
            for (int k = 0; k < 9; k++) res[k] = e[k];
        }
    }
}
// the same, also reporting which of the four yaw variants won and the measurement pose it was compared with (analytic Jacobians)
__device__ __forceinline__ void cube_min_log_error_k(const Cube& self, const Cube& newone, double* res, SE3& best_pose) {
    double best_n = 0;
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
        double yaw_angle = (double)(i - 1) * M_PI / 2.0;
        Cube rc;
        SE3 rot = se3_make(Quat{cos(yaw_angle * 0.5), 0, 0, sin(yaw_angle * 0.5)}, V3{0, 0, 0});
        rc.pose = se3_mul(newone.pose, rot);
        rc.scale = newone.scale;
        if ((yaw_angle == M_PI / 2.0) || (yaw_angle == -M_PI / 2.0) || (yaw_angle == 3 * M_PI / 2.0)) { double t = rc.scale.x; rc.scale.x = rc.scale.y; rc.scale.y = t; }
        double e[9];
        cube_log_error(self, rc, e);
        double s = 0;
#pragma unroll
        for (int k = 0; k < 9; k++) s += e[k] * e[k];
        double n = sqrt(s);
        if (i == 0 || n < best_n) {
            best_n = n;
            best_pose = rc.pose;
#pragma unroll
            for (int k = 0; k < 9; k++) res[k] = e[k];
        }
    }
}

// ---- closed-form derivatives of the SE(3) logarithm ----------------------------------------------------------
// g2o orders a tangent vector as [omega (rotation); upsilon (translation)] (se3quat.h:230-323).  For xi = log(T):
//   log(exp(eps) T) = xi + Jl^-1(xi) eps + O(eps^2),   log(T exp(eps)) = xi + Jl^-1(-xi) eps + O(eps^2)
//   Jl^-1(xi) = [ J^-1 0 ; -J^-1 Q J^-1  J^-1 ]  with J the SO(3) left Jacobian of omega and Q(upsilon, omega) as in Barfoot,
//   "State Estimation for Robotics", eq. 7.86 (there in [rho; phi] order).  se3_jl_inv_apply returns Jl^-1(xi) v.
__device__ __forceinline__ M3 m3_add(const M3& a, const M3& b, double sb) { M3 r; for (int i = 0; i < 9; i++) r.m[i] = a.m[i] + sb * b.m[i]; return r; }
__device__ __forceinline__ void se3_jl_inv_apply(const double* xi, const double* v, double* out) {
    const V3 w{xi[0], xi[1], xi[2]}, u{xi[3], xi[4], xi[5]};
    const M3 W = skew(w), P = skew(u);
    const double th2 = w.x * w.x + w.y * w.y + w.z * w.z, th = sqrt(th2);
    const M3 WW = mul(W, W), WP = mul(W, P), PW = mul(P, W), WPW = mul(WP, W);
    M3 Ji, Q;
    const M3 t1 = m3_add(m3_add(WP, PW, 1.0), WPW, 1.0);                       // WP + PW + WPW
    const M3 t2 = m3_add(m3_add(mul(WW, P), mul(P, WW), 1.0), WPW, -3.0);      // WWP + PWW - 3 WPW
    const M3 t3 = m3_add(mul(WPW, W), mul(W, WPW), 1.0);                        // WPWW + WWPW
    if (th < 1e-4) {
        Ji = m3_add(m3_add(identity3(), W, -0.5), WW, 1.0 / 12.0);
        Q = m3_add(m3_add(m3_add(M3{{0, 0, 0, 0, 0, 0, 0, 0, 0}}, P, 0.5), t1, 1.0 / 6.0), t2, -1.0 / 24.0);
    } else {
        const double half = 0.5 * th, cot = half / tan(half);
        // J^-1 = cot I + (1 - cot) a a^T - (theta / 2) a^  with a = w / theta;  a a^T = I + a^ a^
        Ji = m3_add(m3_add(identity3(), WW, (1.0 - cot) / th2), W, -0.5);
        const double sn = sin(th), cs = cos(th);
        const double c1 = (th - sn) / (th2 * th);
        const double c2 = (1.0 - 0.5 * th2 - cs) / (th2 * th2);
        const double c3 = 0.5 * (c2 - 3.0 * (th - sn - th2 * th / 6.0) / (th2 * th2 * th));
        Q = m3_add(m3_add(m3_add(m3_add(M3{{0, 0, 0, 0, 0, 0, 0, 0, 0}}, P, 0.5), t1, c1), t2, -c2), t3, -c3);
    }
    const V3 a = mul(Ji, V3{v[0], v[1], v[2]});
    const V3 b = mul(Ji, V3{v[3], v[4], v[5]});
    const V3 qa = mul(Ji, mul(Q, a));
    out[0] = a.x; out[1] = a.y; out[2] = a.z;
    out[3] = b.x - qa.x; out[4] = b.y - qa.y; out[5] = b.z - qa.z;
}
// column c of Ad_T = [ R 0 ; t^ R  R ]  (T exp(x) T^-1 = exp(Ad_T x))
__device__ __forceinline__ void se3_adjoint_col(const SE3& T, int c, double* col) {
    const M3 R = quat_to_rot(T.r);
    if (c < 3) {
        const V3 r{R.m[c], R.m[3 + c], R.m[6 + c]};
        const V3 tr = cross(T.t, r);
        col[0] = r.x; col[1] = r.y; col[2] = r.z; col[3] = tr.x; col[4] = tr.y; col[5] = tr.z;
    } else {
        const int k = c - 3;
        col[0] = col[1] = col[2] = 0.0;
        col[3] = R.m[k]; col[4] = R.m[3 + k]; col[5] = R.m[6 + k];
    }
}

// projectOntoImageBbox :156-197
__device__ __forceinline__ void cube_project_bbox(const Cube& c, const SE3& Tcw, const double* K, double* out) {
    const M3 R = quat_to_rot(c.pose.r);
    const M3 Rc = quat_to_rot(Tcw.r);
    const double k0 = K[0], k1 = K[1], k2 = K[2], k3 = K[3], k4 = K[4], k5 = K[5], k6 = K[6], k7 = K[7], k8 = K[8];
    // similarityTransform(): R * diag(scale) | t
    const double s00 = R.m[0] * c.scale.x, s01 = R.m[1] * c.scale.y, s02 = R.m[2] * c.scale.z;
    const double s10 = R.m[3] * c.scale.x, s11 = R.m[4] * c.scale.y, s12 = R.m[5] * c.scale.z;
    const double s20 = R.m[6] * c.scale.x, s21 = R.m[7] * c.scale.y, s22 = R.m[8] * c.scale.z;
    double minx = 0, miny = 0, maxx = 0, maxy = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        // corners_body (g2o_Object.h:169-171): x: 1 1 -1 -1 1 1 -1 -1 | y: 1 -1 -1 1 1 -1 -1 1 | z: -1 -1 -1 -1 1 1 1 1
        const double bx = ((k >> 1) & 1) ? -1.0 : 1.0;
        const double by = (((k & 3) == 0) || ((k & 3) == 3)) ? 1.0 : -1.0;
        const double bz = (k < 4) ? -1.0 : 1.0;
        const double w3 = ((0.0 * bx + 0.0 * by) + 0.0 * bz) + 1.0 * 1.0;
        const double cwx = (((s00 * bx + s01 * by) + s02 * bz) + c.pose.t.x * 1.0) / w3;
        const double cwy = (((s10 * bx + s11 * by) + s12 * bz) + c.pose.t.y * 1.0) / w3;
        const double cwz = (((s20 * bx + s21 * by) + s22 * bz) + c.pose.t.z * 1.0) / w3;
        const double p3 = ((0.0 * cwx + 0.0 * cwy) + 0.0 * cwz) + 1.0 * 1.0;
        const double pcx = (((Rc.m[0] * cwx + Rc.m[1] * cwy) + Rc.m[2] * cwz) + Tcw.t.x * 1.0) / p3;
        const double pcy = (((Rc.m[3] * cwx + Rc.m[4] * cwy) + Rc.m[5] * cwz) + Tcw.t.y * 1.0) / p3;
        const double pcz = (((Rc.m[6] * cwx + Rc.m[7] * cwy) + Rc.m[8] * cwz) + Tcw.t.z * 1.0) / p3;
        const double ux = (k0 * pcx + k1 * pcy) + k2 * pcz, uy = (k3 * pcx + k4 * pcy) + k5 * pcz, uz = (k6 * pcx + k7 * pcy) + k8 * pcz;
        const double u = ux / uz, v = uy / uz;
        if (k == 0) { minx = maxx = u; miny = maxy = v; }
        else { if (u > maxx) maxx = u; if (u < minx) minx = u; if (v > maxy) maxy = v; if (v < miny) miny = v; }
    }
    out[0] = (maxx + minx) / 2; out[1] = (maxy + miny) / 2; out[2] = maxx - minx; out[3] = maxy - miny;
}

// ---- edge types ---------------------------------------------------------------------------------
enum { EDGE_CUBOID = 0, EDGE_PROJ = 1, EDGE_ODOM = 2 };
template <int TYPE> struct EdgeDims;
template <> struct EdgeDims<EDGE_CUBOID> { static constexpr int D = 9, Di = 6, Dj = 9, REC = 132; };
template <> struct EdgeDims<EDGE_PROJ> { static constexpr int D = 4, Di = 6, Dj = 9, REC = 132; };
template <> struct EdgeDims<EDGE_ODOM> { static constexpr int D = 6, Di = 6, Dj = 6, REC = 84; };

struct EdgeCtx {
    SE3 cam;          // vertex 0 estimate (world -> camera)
    SE3 cam2;         // odometry: vertex 1 estimate
    Cube cube;        // vertex 1 estimate
    Cube meas_cube;   // EdgeSE3Cuboid measurement
    SE3 meas_se3;     // EdgeSE3Expmap measurement
    double meas4[4];  // EdgeSE3CuboidProj measurement
    const double* K;
};

// computeError() of the three edge classes with vertex 0 / vertex 1 replaced by (possibly perturbed) estimates
template <int TYPE>
__device__ __forceinline__ void edge_error(const EdgeCtx& x, const SE3& v0, const Cube& v1c, const SE3& v1s, double* e) {
    if (TYPE == EDGE_CUBOID) {  // g2o_Object.h:250-259
        SE3 Twc = se3_inverse(v0);
        Cube esti;
        esti.pose = se3_mul(Twc, x.meas_cube.pose);  // transform_from :117-122
        esti.scale = x.meas_cube.scale;
        cube_min_log_error(v1c, esti, e);
    } else if (TYPE == EDGE_PROJ) {  // g2o_Object.h:279-290
        double r[4];
        cube_project_bbox(v1c, v0, x.K, r);
        for (int i = 0; i < 4; i++) e[i] = r[i] - x.meas4[i];
    } else {  // types_six_dof_expmap.h:90-99
        SE3 err = se3_mul(se3_mul(x.meas_se3, v0), se3_inverse(v1s));
        se3_log(err, e);
    }
}

}  // namespace csb
