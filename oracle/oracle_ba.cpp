// TEST INFRASTRUCTURE ONLY -- CPU oracle, BA half of the CubeSLAM hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build or call this file; the shipped library never links it.
//
// Restates (paths relative to /root/reference/object_slam):
//   g2o::cuboid / VertexCuboid / EdgeSE3Cuboid / EdgeSE3CuboidProj   include/object_slam/g2o_Object.h:23-292
//   SE3Quat                                                           Thirdparty/g2o/g2o/types/se3quat.h:41-362, se3_ops.hpp:28-48
//   VertexSE3Expmap::oplusImpl, EdgeSE3Expmap::computeError           Thirdparty/g2o/g2o/types/types_six_dof_expmap.h:59-99
//   BaseBinaryEdge::linearizeOplus (central differences, delta=1e-9)  Thirdparty/g2o/g2o/core/base_binary_edge.hpp:130-205
//   BaseBinaryEdge::constructQuadraticForm                            Thirdparty/g2o/g2o/core/base_binary_edge.hpp:54-120
//   BlockSolver::buildSystem                                          Thirdparty/g2o/g2o/core/block_solver.hpp:501-560
//   OptimizationAlgorithmLevenberg::solve                             Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189
//   LinearSolverDense::solve (Eigen LDLT -> plain Cholesky LDL^T here) Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:65-113
//
// PARITY STATUS: PINNED against outputs of the reference itself: tests/test_reference_replay.py re-runs object_slam's online mode on the
// bundled TUM sequence (graph recipe of main_obj.cpp:738-803, optimize(5) after every frame with this file's numeric Jacobians and
// Levenberg-Marquardt) and reproduces the committed output_obj_poses.txt (landmark after every one of the 58 frames to the printed
// digits) and output_cam_poses.txt (0.07 mm in the median).  The offline fixture is covered by tests/test_oracle_golden.py.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "oracle_math.h"

using namespace orc;

namespace {

struct SE3 {
    Quat r{1, 0, 0, 0};
    V3 t{0, 0, 0};
};
// se3quat.h:346-351
void normalizeRotation(SE3& s) {
    if (s.r.w < 0) { s.r.w = -s.r.w; s.r.x = -s.r.x; s.r.y = -s.r.y; s.r.z = -s.r.z; }
    quat_normalize(s.r);
}
SE3 se3_from(const Quat& q, V3 t) { SE3 s; s.r = q; s.t = t; normalizeRotation(s); return s; }
// se3quat.h:165-168 (fromVector: x y z qx qy qz qw) followed by the Vector7d ctor's normalize (:66-69)
SE3 se3_from_vec7(const double* v) { return se3_from(Quat{v[6], v[3], v[4], v[5]}, V3{v[0], v[1], v[2]}); }
void se3_to_vec7(const SE3& s, double* v) { v[0] = s.t.x; v[1] = s.t.y; v[2] = s.t.z; v[3] = s.r.x; v[4] = s.r.y; v[5] = s.r.z; v[6] = s.r.w; }
// se3quat.h:110-116
SE3 operator*(const SE3& a, const SE3& b) {
    SE3 r;
    r.t = a.t + quat_rot(a.r, b.t);
    r.r = quat_mul(a.r, b.r);
    normalizeRotation(r);
    return r;
}
// se3quat.h:129-134
SE3 inverse(const SE3& a) {
    SE3 r;
    r.r = quat_conj(a.r);
    r.t = quat_rot(r.r, V3{a.t.x * -1., a.t.y * -1., a.t.z * -1.});
    return r;
}
// se3_ops.hpp:28-48
M3 skew(V3 v) {
    M3 m; std::memset(m.m, 0, sizeof m.m);
    m(0, 1) = -v.z; m(0, 2) = v.y; m(1, 2) = -v.x; m(1, 0) = v.z; m(2, 0) = -v.y; m(2, 1) = v.x;
    return m;
}
V3 deltaR(const M3& R) { return {R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1)}; }
// se3quat.h:230-267
void se3_log(const SE3& s, double res[6]) {
    M3 R = quat_to_rot(s.r);
    double d = 0.5 * (R(0, 0) + R(1, 1) + R(2, 2) - 1);
    V3 omega;
    V3 dR = deltaR(R);
    M3 V_inv;
    M3 I = identity3();
    if (d > 0.99999) {
        omega = 0.5 * dR;
        M3 Om = skew(omega);
        M3 Om2 = mul(Om, Om);
        for (int i = 0; i < 9; i++) V_inv.m[i] = I.m[i] - 0.5 * Om.m[i] + (1. / 12.) * Om2.m[i];
    } else {
        double theta = std::acos(d);
        omega = (theta / (2 * std::sqrt(1 - d * d))) * dR;
        M3 Om = skew(omega);
        M3 Om2 = mul(Om, Om);
        double c = (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta);
        for (int i = 0; i < 9; i++) V_inv.m[i] = I.m[i] - 0.5 * Om.m[i] + c * Om2.m[i];
    }
    V3 ups = mul(V_inv, s.t);
    res[0] = omega.x; res[1] = omega.y; res[2] = omega.z; res[3] = ups.x; res[4] = ups.y; res[5] = ups.z;
}
// se3quat.h:275-323
SE3 se3_exp(const double u[6]) {
    V3 omega{u[0], u[1], u[2]}, upsilon{u[3], u[4], u[5]};
    double theta = norm(omega);
    M3 Om = skew(omega);
    M3 R, V;
    M3 I = identity3();
    M3 Om2 = mul(Om, Om);
    if (theta < 0.00001) {
        for (int i = 0; i < 9; i++) R.m[i] = I.m[i] + Om.m[i] + Om2.m[i];
        V = R;
    } else {
        double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta), c = (theta - std::sin(theta)) / (std::pow(theta, 3));
        for (int i = 0; i < 9; i++) R.m[i] = I.m[i] + a * Om.m[i] + b * Om2.m[i];
        for (int i = 0; i < 9; i++) V.m[i] = I.m[i] + b * Om.m[i] + c * Om2.m[i];
    }
    return se3_from(quat_from_rot(R), mul(V, upsilon));
}
// se3quat.h:336-344
M4 to_homogeneous(const SE3& s) {
    M4 h; std::memset(h.m, 0, sizeof h.m); h(3, 3) = 1;
    M3 R = quat_to_rot(s.r);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) h(i, j) = R(i, j);
    h(0, 3) = s.t.x; h(1, 3) = s.t.y; h(2, 3) = s.t.z;
    return h;
}

// g2o_Object.h:23-199
struct Cuboid {
    SE3 pose;
    V3 scale{0, 0, 0};
};
Cuboid cuboid_from_vec10(const double* v) {  // :45-48 (fromVector: no normalisation of the quaternion!)
    Cuboid c;
    c.pose.r = Quat{v[6], v[3], v[4], v[5]};
    c.pose.t = V3{v[0], v[1], v[2]};
    c.scale = V3{v[7], v[8], v[9]};
    return c;
}
void cuboid_to_vec10(const Cuboid& c, double* v) { se3_to_vec7(c.pose, v); v[7] = c.scale.x; v[8] = c.scale.y; v[9] = c.scale.z; }
Cuboid cuboid_from_minimal(const double* v) {  // :36-41  xyz rpy scale
    Cuboid c;
    c.pose = se3_from(zyx_euler_to_quat(v[3], v[4], v[5]), V3{v[0], v[1], v[2]});
    c.scale = V3{v[6], v[7], v[8]};
    return c;
}
Cuboid exp_update(const Cuboid& c, const double u[9]) {  // :57-63
    Cuboid r;
    r.pose = c.pose * se3_exp(u);
    r.scale = V3{c.scale.x + u[6], c.scale.y + u[7], c.scale.z + u[8]};
    return r;
}
Cuboid rotate_cuboid(const Cuboid& c, double yaw_angle) {  // :104-114
    Cuboid r;
    SE3 rot = se3_from(Quat{std::cos(yaw_angle * 0.5), 0, 0, std::sin(yaw_angle * 0.5)}, V3{0, 0, 0});
    r.pose = c.pose * rot;
    r.scale = c.scale;
    if ((yaw_angle == M_PI / 2.0) || (yaw_angle == -M_PI / 2.0) || (yaw_angle == 3 * M_PI / 2.0)) std::swap(r.scale.x, r.scale.y);
    return r;
}
void cube_log_error(const Cuboid& self, const Cuboid& newone, double res[9]) {  // :66-73
    SE3 pose_diff = inverse(newone.pose) * self.pose;
    se3_log(pose_diff, res);
    res[6] = self.scale.x - newone.scale.x; res[7] = self.scale.y - newone.scale.y; res[8] = self.scale.z - newone.scale.z;
}
void min_log_error(const Cuboid& self, const Cuboid& newone, double res[9]) {  // :76-101
    const double rotate_angles[4] = {-1, 0, 1, 2};
    double errs[4][9], nrm[4];
    for (int i = 0; i < 4; i++) {
        Cuboid rc = rotate_cuboid(newone, rotate_angles[i] * M_PI / 2.0);
        cube_log_error(self, rc, errs[i]);
        double s = 0;
        for (int k = 0; k < 9; k++) s += errs[i][k] * errs[i][k];
        nrm[i] = std::sqrt(s);
    }
    int min_label = 0;  // Eigen minCoeff(&idx): first minimum, strict <
    for (int i = 1; i < 4; i++) if (nrm[i] < nrm[min_label]) min_label = i;
    std::memcpy(res, errs[min_label], 9 * sizeof(double));
}
Cuboid transform_from(const Cuboid& c, const SE3& Twc) { Cuboid r; r.pose = Twc * c.pose; r.scale = c.scale; return r; }  // :117-122
Cuboid transform_to(const Cuboid& c, const SE3& Twc) { Cuboid r; r.pose = inverse(Twc) * c.pose; r.scale = c.scale; return r; }  // :126-132
// :156-197
void projectOntoImageBbox(const Cuboid& c, const SE3& Tcw, const M3& K, double out[4]) {
    static const double body[3][8] = {{1, 1, -1, -1, 1, 1, -1, -1}, {1, -1, -1, 1, 1, -1, -1, 1}, {-1, -1, -1, -1, 1, 1, 1, 1}};
    M4 S = to_homogeneous(c.pose);
    double sc[3] = {c.scale.x, c.scale.y, c.scale.z};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) S(i, j) = S(i, j) * sc[j];
    M4 Tc = to_homogeneous(Tcw);
    double minx = 0, miny = 0, maxx = 0, maxy = 0;
    for (int k = 0; k < 8; k++) {
        double w[4], p[4];
        for (int r = 0; r < 4; r++) w[r] = ((S(r, 0) * body[0][k] + S(r, 1) * body[1][k]) + S(r, 2) * body[2][k]) + S(r, 3) * 1.0;
        double cw[3] = {w[0] / w[3], w[1] / w[3], w[2] / w[3]};
        for (int r = 0; r < 4; r++) p[r] = ((Tc(r, 0) * cw[0] + Tc(r, 1) * cw[1]) + Tc(r, 2) * cw[2]) + Tc(r, 3) * 1.0;
        V3 pc{p[0] / p[3], p[1] / p[3], p[2] / p[3]};
        V3 uv = mul(K, pc);
        double u = uv.x / uv.z, v = uv.y / uv.z;
        if (k == 0) { minx = maxx = u; miny = maxy = v; }
        else { if (u > maxx) maxx = u; if (u < minx) minx = u; if (v > maxy) maxy = v; if (v < miny) miny = v; }
    }
    out[0] = (maxx + minx) / 2; out[1] = (maxy + miny) / 2; out[2] = maxx - minx; out[3] = maxy - miny;
}

// residuals -------------------------------------------------------------------------------------
// EdgeSE3Cuboid::computeError, g2o_Object.h:250-259
void err_se3cuboid(const SE3& Tcw, const Cuboid& cube, const Cuboid& meas, double e[9]) {
    SE3 Twc = inverse(Tcw);
    Cuboid esti = transform_from(meas, Twc);
    min_log_error(cube, esti, e);
}
// EdgeSE3CuboidProj::computeError, g2o_Object.h:279-290
void err_proj(const SE3& Tcw, const Cuboid& cube, const M3& K, const double meas[4], double e[4]) {
    double r[4];
    projectOntoImageBbox(cube, Tcw, K, r);
    for (int i = 0; i < 4; i++) e[i] = r[i] - meas[i];
}
// EdgeSE3Expmap::computeError, types_six_dof_expmap.h:90-99
void err_odom(const SE3& T1, const SE3& T2, const SE3& C, double e[6]) {
    SE3 err = C * T1 * inverse(T2);
    se3_log(err, e);
}
// VertexSE3Expmap::oplusImpl, types_six_dof_expmap.h:73-76
SE3 cam_oplus(const SE3& T, const double u[6]) { return se3_exp(u) * T; }

// delta = 1e-9 is the reference's (base_binary_edge.hpp:147).  Test knob (orc_ba_set_delta): a larger step gives central differences whose
// round-off is far below the reference's (1e-16 / delta), i.e. an independent, accurate Jacobian to hold the closed-form ones against.
double kDelta = 1e-9;
double kScalar = 1.0 / (2 * kDelta);

// Generic numeric Jacobian (base_binary_edge.hpp:130-205).  f(vi, vj, err) evaluates the residual.
template <int D, int Di, int Dj, class VI, class VJ, class F, class OI, class OJ>
void numeric_jacobian(const VI& vi, const VJ& vj, bool i_free, bool j_free, F f, OI oplus_i, OJ oplus_j, double* Ji /*D x Di col-major*/, double* Jj) {
    double ep[D], em[D];
    if (i_free) {
        double add[Di];
        std::fill(add, add + Di, 0.0);
        for (int d = 0; d < Di; d++) {
            add[d] = kDelta; f(oplus_i(vi, add), vj, ep);
            add[d] = -kDelta; f(oplus_i(vi, add), vj, em);
            add[d] = 0.0;
            for (int r = 0; r < D; r++) Ji[d * D + r] = kScalar * (ep[r] - em[r]);
        }
    }
    if (j_free) {
        double add[Dj];
        std::fill(add, add + Dj, 0.0);
        for (int d = 0; d < Dj; d++) {
            add[d] = kDelta; f(vi, oplus_j(vj, add), ep);
            add[d] = -kDelta; f(vi, oplus_j(vj, add), em);
            add[d] = 0.0;
            for (int r = 0; r < D; r++) Jj[d * D + r] = kScalar * (ep[r] - em[r]);
        }
    }
}

// constructQuadraticForm (base_binary_edge.hpp:54-120), no robust kernel.
// J col-major (D x Di), info row-major D x D (symmetric).  Hii Di x Di col-major, Hij Di x Dj col-major.
template <int D, int Di, int Dj>
void quadratic_form(const double* A, const double* B, const double* omega, const double* err, bool i_free, bool j_free, double* Hii, double* bi,
                    double* Hjj, double* bj, double* Hij) {
    double omega_r[D];
    for (int r = 0; r < D; r++) {
        double s = 0;
        for (int k = 0; k < D; k++) s += omega[r * D + k] * err[k];
        omega_r[r] = -s;
    }
    if (i_free) {
        double AtO[Di * D];  // Di x D row-major
        for (int i = 0; i < Di; i++)
            for (int c = 0; c < D; c++) {
                double s = 0;
                for (int k = 0; k < D; k++) s += A[i * D + k] * omega[k * D + c];
                AtO[i * D + c] = s;
            }
        for (int i = 0; i < Di; i++) {
            double s = 0;
            for (int k = 0; k < D; k++) s += A[i * D + k] * omega_r[k];
            bi[i] += s;
        }
        for (int c = 0; c < Di; c++)
            for (int i = 0; i < Di; i++) {
                double s = 0;
                for (int k = 0; k < D; k++) s += AtO[i * D + k] * A[c * D + k];
                Hii[c * Di + i] += s;
            }
        if (j_free)
            for (int c = 0; c < Dj; c++)
                for (int i = 0; i < Di; i++) {
                    double s = 0;
                    for (int k = 0; k < D; k++) s += AtO[i * D + k] * B[c * D + k];
                    Hij[c * Di + i] += s;
                }
    }
    if (j_free) {
        double BtO[Dj * D];
        for (int i = 0; i < Dj; i++)
            for (int c = 0; c < D; c++) {
                double s = 0;
                for (int k = 0; k < D; k++) s += B[i * D + k] * omega[k * D + c];
                BtO[i * D + c] = s;
            }
        for (int i = 0; i < Dj; i++) {
            double s = 0;
            for (int k = 0; k < D; k++) s += B[i * D + k] * omega_r[k];
            bj[i] += s;
        }
        for (int c = 0; c < Dj; c++)
            for (int i = 0; i < Dj; i++) {
                double s = 0;
                for (int k = 0; k < D; k++) s += BtO[i * D + k] * B[c * D + k];
                Hjj[c * Dj + i] += s;
            }
    }
}

struct Graph {
    int n_cam = 0, n_cube = 0, n_ec = 0, n_ep = 0, n_eo = 0;
    std::vector<SE3> cams; std::vector<int> cam_fixed;
    std::vector<Cuboid> cubes; std::vector<int> cube_fixed;
    const int *ec_cam = nullptr, *ec_cube = nullptr; const double *ec_meas = nullptr, *ec_info = nullptr;
    const int *ep_cam = nullptr, *ep_cube = nullptr; const double *ep_meas = nullptr, *ep_info = nullptr, *ep_K = nullptr;
    const int *eo_i = nullptr, *eo_j = nullptr; const double *eo_meas = nullptr, *eo_info = nullptr;
};

struct Lin {
    std::vector<double> ec_err, ec_Ji, ec_Jj, ep_err, ep_Ji, ep_Jj, eo_err, eo_Ji, eo_Jj;
    std::vector<double> H_cam, b_cam, H_cube, b_cube, ec_Hij, ep_Hij, eo_Hij;
};

// computeActiveErrors (sparse_optimizer.cpp:61-88)
void compute_errors(const Graph& g, Lin& L) {
    L.ec_err.assign((size_t)g.n_ec * 9, 0); L.ep_err.assign((size_t)g.n_ep * 4, 0); L.eo_err.assign((size_t)g.n_eo * 6, 0);
    for (int e = 0; e < g.n_ec; e++) err_se3cuboid(g.cams[g.ec_cam[e]], g.cubes[g.ec_cube[e]], cuboid_from_vec10(g.ec_meas + 10 * e), &L.ec_err[9 * e]);
    for (int e = 0; e < g.n_ep; e++) { M3 K; std::memcpy(K.m, g.ep_K + 9 * e, sizeof K.m); err_proj(g.cams[g.ep_cam[e]], g.cubes[g.ep_cube[e]], K, g.ep_meas + 4 * e, &L.ep_err[4 * e]); }
    for (int e = 0; e < g.n_eo; e++) err_odom(g.cams[g.eo_i[e]], g.cams[g.eo_j[e]], se3_from_vec7(g.eo_meas + 7 * e), &L.eo_err[6 * e]);
}
double chi2_of(const double* e, const double* info, int D) {  // BaseEdge::chi2 = e^T * Omega * e
    double c = 0;
    for (int r = 0; r < D; r++) { double s = 0; for (int k = 0; k < D; k++) s += info[r * D + k] * e[k]; c += e[r] * s; }
    return c;
}
double active_chi2(const Graph& g, const Lin& L) {
    double chi = 0;
    for (int e = 0; e < g.n_ec; e++) chi += chi2_of(&L.ec_err[9 * e], g.ec_info + 81 * e, 9);
    for (int e = 0; e < g.n_ep; e++) chi += chi2_of(&L.ep_err[4 * e], g.ep_info + 16 * e, 4);
    for (int e = 0; e < g.n_eo; e++) chi += chi2_of(&L.eo_err[6 * e], g.eo_info + 36 * e, 6);
    return chi;
}

// BlockSolver::buildSystem (block_solver.hpp:501-560): edge order = cuboid edges, projection edges, odometry edges
// (matches main_obj.cpp:768,788 where cuboid-edge ids < odometry-edge ids and g2o sorts active edges by id).
void build_system(const Graph& g, Lin& L, int n_threads = 1) {
    L.ec_Ji.assign((size_t)g.n_ec * 54, 0); L.ec_Jj.assign((size_t)g.n_ec * 81, 0);
    L.ep_Ji.assign((size_t)g.n_ep * 24, 0); L.ep_Jj.assign((size_t)g.n_ep * 36, 0);
    L.eo_Ji.assign((size_t)g.n_eo * 36, 0); L.eo_Jj.assign((size_t)g.n_eo * 36, 0);
    L.H_cam.assign((size_t)g.n_cam * 36, 0); L.b_cam.assign((size_t)g.n_cam * 6, 0);
    L.H_cube.assign((size_t)g.n_cube * 81, 0); L.b_cube.assign((size_t)g.n_cube * 9, 0);
    L.ec_Hij.assign((size_t)g.n_ec * 54, 0); L.ep_Hij.assign((size_t)g.n_ep * 54, 0); L.eo_Hij.assign((size_t)g.n_eo * 36, 0);
    auto cube_oplus = [](const Cuboid& c, const double* u) { return exp_update(c, u); };
    auto cam_op = [](const SE3& T, const double* u) { return cam_oplus(T, u); };
    // linearizeOplus for every edge (independent -> may be threaded for the all-cores baseline)
    auto lin_range = [&](int tid, int nt) {
        for (int e = tid; e < g.n_ec; e += nt) {
            int ci = g.ec_cam[e], cj = g.ec_cube[e];
            Cuboid meas = cuboid_from_vec10(g.ec_meas + 10 * e);
            auto f = [&](const SE3& T, const Cuboid& c, double* out) { err_se3cuboid(T, c, meas, out); };
            numeric_jacobian<9, 6, 9>(g.cams[ci], g.cubes[cj], !g.cam_fixed[ci], !g.cube_fixed[cj], f, cam_op, cube_oplus, &L.ec_Ji[54 * (size_t)e], &L.ec_Jj[81 * (size_t)e]);
        }
        for (int e = tid; e < g.n_ep; e += nt) {
            int ci = g.ep_cam[e], cj = g.ep_cube[e];
            M3 K; std::memcpy(K.m, g.ep_K + 9 * e, sizeof K.m);
            const double* meas = g.ep_meas + 4 * e;
            auto f = [&](const SE3& T, const Cuboid& c, double* out) { err_proj(T, c, K, meas, out); };
            numeric_jacobian<4, 6, 9>(g.cams[ci], g.cubes[cj], !g.cam_fixed[ci], !g.cube_fixed[cj], f, cam_op, cube_oplus, &L.ep_Ji[24 * (size_t)e], &L.ep_Jj[36 * (size_t)e]);
        }
        for (int e = tid; e < g.n_eo; e += nt) {
            int ci = g.eo_i[e], cj = g.eo_j[e];
            SE3 C = se3_from_vec7(g.eo_meas + 7 * e);
            auto f = [&](const SE3& T1, const SE3& T2, double* out) { err_odom(T1, T2, C, out); };
            numeric_jacobian<6, 6, 6>(g.cams[ci], g.cams[cj], !g.cam_fixed[ci], !g.cam_fixed[cj], f, cam_op, cam_op, &L.eo_Ji[36 * (size_t)e], &L.eo_Jj[36 * (size_t)e]);
        }
    };
    if (n_threads <= 1) lin_range(0, 1);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(lin_range, t, n_threads);
        for (auto& t : th) t.join();
    }
    // constructQuadraticForm, sequential in edge order
    for (int e = 0; e < g.n_ec; e++) {
        int ci = g.ec_cam[e], cj = g.ec_cube[e];
        quadratic_form<9, 6, 9>(&L.ec_Ji[54 * (size_t)e], &L.ec_Jj[81 * (size_t)e], g.ec_info + 81 * (size_t)e, &L.ec_err[9 * (size_t)e], !g.cam_fixed[ci], !g.cube_fixed[cj],
                                &L.H_cam[36 * (size_t)ci], &L.b_cam[6 * (size_t)ci], &L.H_cube[81 * (size_t)cj], &L.b_cube[9 * (size_t)cj], &L.ec_Hij[54 * (size_t)e]);
    }
    for (int e = 0; e < g.n_ep; e++) {
        int ci = g.ep_cam[e], cj = g.ep_cube[e];
        quadratic_form<4, 6, 9>(&L.ep_Ji[24 * (size_t)e], &L.ep_Jj[36 * (size_t)e], g.ep_info + 16 * (size_t)e, &L.ep_err[4 * (size_t)e], !g.cam_fixed[ci], !g.cube_fixed[cj],
                                &L.H_cam[36 * (size_t)ci], &L.b_cam[6 * (size_t)ci], &L.H_cube[81 * (size_t)cj], &L.b_cube[9 * (size_t)cj], &L.ep_Hij[54 * (size_t)e]);
    }
    for (int e = 0; e < g.n_eo; e++) {
        int ci = g.eo_i[e], cj = g.eo_j[e];
        quadratic_form<6, 6, 6>(&L.eo_Ji[36 * (size_t)e], &L.eo_Jj[36 * (size_t)e], g.eo_info + 36 * (size_t)e, &L.eo_err[6 * (size_t)e], !g.cam_fixed[ci], !g.cam_fixed[cj],
                                &L.H_cam[36 * (size_t)ci], &L.b_cam[6 * (size_t)ci], &L.H_cam[36 * (size_t)cj], &L.b_cam[6 * (size_t)cj], &L.eo_Hij[36 * (size_t)e]);
    }
}

void load_graph(Graph& g, int n_cam, const double* cams7, const int* cam_fixed, int n_cube, const double* cubes10, const int* cube_fixed) {
    g.n_cam = n_cam; g.n_cube = n_cube;
    g.cams.resize(n_cam); g.cam_fixed.assign(cam_fixed, cam_fixed + n_cam);
    g.cubes.resize(n_cube); g.cube_fixed.assign(cube_fixed, cube_fixed + n_cube);
    for (int i = 0; i < n_cam; i++) g.cams[i] = se3_from_vec7(cams7 + 7 * i);
    for (int i = 0; i < n_cube; i++) g.cubes[i] = cuboid_from_vec10(cubes10 + 10 * i);
}

// Dense system in g2o's ordering for main_obj.cpp's graph: free cuboid vertices first (ids 0..), then free cameras.
struct Dense {
    int n = 0;
    std::vector<int> cube_col, cam_col;
    std::vector<double> H, b;
};
void assemble_dense(const Graph& g, const Lin& L, Dense& D) {
    D.cube_col.assign(g.n_cube, -1); D.cam_col.assign(g.n_cam, -1);
    int n = 0;
    for (int i = 0; i < g.n_cube; i++) if (!g.cube_fixed[i]) { D.cube_col[i] = n; n += 9; }
    for (int i = 0; i < g.n_cam; i++) if (!g.cam_fixed[i]) { D.cam_col[i] = n; n += 6; }
    D.n = n; D.H.assign((size_t)n * n, 0); D.b.assign(n, 0);
    auto H = [&](int r, int c) -> double& { return D.H[(size_t)r * n + c]; };
    for (int i = 0; i < g.n_cube; i++) if (D.cube_col[i] >= 0) {
        int o = D.cube_col[i];
        for (int c = 0; c < 9; c++) { for (int r = 0; r < 9; r++) H(o + r, o + c) = L.H_cube[81 * (size_t)i + c * 9 + r]; D.b[o + c] = L.b_cube[9 * (size_t)i + c]; }
    }
    for (int i = 0; i < g.n_cam; i++) if (D.cam_col[i] >= 0) {
        int o = D.cam_col[i];
        for (int c = 0; c < 6; c++) { for (int r = 0; r < 6; r++) H(o + r, o + c) = L.H_cam[36 * (size_t)i + c * 6 + r]; D.b[o + c] = L.b_cam[6 * (size_t)i + c]; }
    }
    auto add_off = [&](int oi, int di, int oj, int dj, const double* blk) {
        if (oi < 0 || oj < 0) return;
        for (int c = 0; c < dj; c++) for (int r = 0; r < di; r++) { H(oi + r, oj + c) += blk[c * di + r]; H(oj + c, oi + r) += blk[c * di + r]; }
    };
    for (int e = 0; e < g.n_ec; e++) add_off(D.cam_col[g.ec_cam[e]], 6, D.cube_col[g.ec_cube[e]], 9, &L.ec_Hij[54 * (size_t)e]);
    for (int e = 0; e < g.n_ep; e++) add_off(D.cam_col[g.ep_cam[e]], 6, D.cube_col[g.ep_cube[e]], 9, &L.ep_Hij[54 * (size_t)e]);
    for (int e = 0; e < g.n_eo; e++) add_off(D.cam_col[g.eo_i[e]], 6, D.cam_col[g.eo_j[e]], 6, &L.eo_Hij[36 * (size_t)e]);
}
// LDL^T without pivoting (the reference uses Eigen::LDLT with pivoting; same solution up to round-off).
bool ldlt_solve(std::vector<double> A, int n, const std::vector<double>& b, std::vector<double>& x) {
    std::vector<double> d(n);
    for (int j = 0; j < n; j++) {
        double s = A[(size_t)j * n + j];
        for (int k = 0; k < j; k++) s -= A[(size_t)j * n + k] * A[(size_t)j * n + k] * d[k];
        d[j] = s;
        if (!(s > 0)) return false;
        for (int i = j + 1; i < n; i++) {
            double t = A[(size_t)i * n + j];
            for (int k = 0; k < j; k++) t -= A[(size_t)i * n + k] * A[(size_t)j * n + k] * d[k];
            A[(size_t)i * n + j] = t / s;
        }
    }
    x = b;
    for (int i = 0; i < n; i++) for (int k = 0; k < i; k++) x[i] -= A[(size_t)i * n + k] * x[k];
    for (int i = 0; i < n; i++) x[i] /= d[i];
    for (int i = n - 1; i >= 0; i--) for (int k = i + 1; k < n; k++) x[i] -= A[(size_t)k * n + i] * x[k];
    return true;
}

struct LMState { double lambda = -1, ni = 2; int nBad = 0; };

// OptimizationAlgorithmLevenberg::solve, one outer iteration.  returns 0 OK, 1 Terminate, -1 Fail
int lm_iteration(Graph& g, int iteration, LMState& st, double* chi_out) {
    const double tau = 1e-5, goodUp = 2. / 3., goodLow = 1. / 3.;
    const int maxTrials = 10;
    Lin L;
    compute_errors(g, L);
    double currentChi = active_chi2(g, L), tempChi = currentChi, iniChi = currentChi;
    build_system(g, L);
    Dense D;
    assemble_dense(g, L, D);
    if (iteration == 0) {
        double maxDiag = 0;
        for (int i = 0; i < D.n; i++) maxDiag = std::max(std::fabs(D.H[(size_t)i * D.n + i]), maxDiag);
        st.lambda = tau * maxDiag; st.ni = 2; st.nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
        std::vector<SE3> cam_bak = g.cams; std::vector<Cuboid> cube_bak = g.cubes;  // push
        std::vector<double> Hl = D.H;
        for (int i = 0; i < D.n; i++) Hl[(size_t)i * D.n + i] += st.lambda;
        std::vector<double> x;
        bool ok2 = ldlt_solve(Hl, D.n, D.b, x);
        if (!ok2) x.assign(D.n, 0.0);
        for (int i = 0; i < g.n_cube; i++) if (D.cube_col[i] >= 0) g.cubes[i] = exp_update(g.cubes[i], &x[D.cube_col[i]]);
        for (int i = 0; i < g.n_cam; i++) if (D.cam_col[i] >= 0) g.cams[i] = cam_oplus(g.cams[i], &x[D.cam_col[i]]);
        Lin L2;
        compute_errors(g, L2);
        tempChi = active_chi2(g, L2);
        if (!ok2) tempChi = std::numeric_limits<double>::max();
        rho = (currentChi - tempChi);
        double scale = 0;
        for (int j = 0; j < D.n; j++) scale += x[j] * (st.lambda * x[j] + D.b[j]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(tempChi)) {
            double alpha = 1. - std::pow((2 * rho - 1), 3);
            alpha = std::min(alpha, goodUp);
            double scaleFactor = std::max(goodLow, alpha);
            st.lambda *= scaleFactor; st.ni = 2; currentChi = tempChi;
        } else {
            st.lambda *= st.ni; st.ni *= 2;
            g.cams = cam_bak; g.cubes = cube_bak;  // pop
        }
        qmax++;
    } while (rho < 0 && qmax < maxTrials);
    if (chi_out) *chi_out = currentChi;
    if (qmax == maxTrials || rho == 0) return 1;
    if ((iniChi - currentChi) * 1e3 < iniChi) st.nBad++; else st.nBad = 0;
    if (st.nBad >= 3) return 1;
    return 0;
}

}  // namespace

extern "C" {

struct orc_ba_edges {
    int n_ec; const int* ec_cam; const int* ec_cube; const double* ec_meas10; const double* ec_info81;
    int n_ep; const int* ep_cam; const int* ep_cube; const double* ep_meas4; const double* ep_info16; const double* ep_K9;
    int n_eo; const int* eo_i; const int* eo_j; const double* eo_meas7; const double* eo_info36;
};
struct orc_ba_out {
    double *ec_err, *ec_Ji, *ec_Jj, *ep_err, *ep_Ji, *ep_Jj, *eo_err, *eo_Ji, *eo_Jj;
    double *H_cam, *b_cam, *H_cube, *b_cube, *ec_Hij, *ep_Hij, *eo_Hij;
};

static void bind_edges(Graph& g, const orc_ba_edges* E) {
    g.n_ec = E->n_ec; g.ec_cam = E->ec_cam; g.ec_cube = E->ec_cube; g.ec_meas = E->ec_meas10; g.ec_info = E->ec_info81;
    g.n_ep = E->n_ep; g.ep_cam = E->ep_cam; g.ep_cube = E->ep_cube; g.ep_meas = E->ep_meas4; g.ep_info = E->ep_info16; g.ep_K = E->ep_K9;
    g.n_eo = E->n_eo; g.eo_i = E->eo_i; g.eo_j = E->eo_j; g.eo_meas = E->eo_meas7; g.eo_info = E->eo_info36;
}

void orc_ba_set_delta(double delta) { kDelta = delta; kScalar = 1.0 / (2 * kDelta); }

// One computeActiveErrors + buildSystem pass.  Any output pointer may be NULL.  Returns chi2.
double orc_ba_linearize(int n_cam, const double* cams7, const int* cam_fixed, int n_cube, const double* cubes10, const int* cube_fixed,
                        const orc_ba_edges* E, const orc_ba_out* O, int n_threads) {
    Graph g; load_graph(g, n_cam, cams7, cam_fixed, n_cube, cubes10, cube_fixed); bind_edges(g, E);
    Lin L; compute_errors(g, L); build_system(g, L, n_threads);
    auto cp = [](double* dst, const std::vector<double>& v) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(double)); };
    if (O) {
        cp(O->ec_err, L.ec_err); cp(O->ec_Ji, L.ec_Ji); cp(O->ec_Jj, L.ec_Jj); cp(O->ep_err, L.ep_err); cp(O->ep_Ji, L.ep_Ji); cp(O->ep_Jj, L.ep_Jj);
        cp(O->eo_err, L.eo_err); cp(O->eo_Ji, L.eo_Ji); cp(O->eo_Jj, L.eo_Jj); cp(O->H_cam, L.H_cam); cp(O->b_cam, L.b_cam); cp(O->H_cube, L.H_cube);
        cp(O->b_cube, L.b_cube); cp(O->ec_Hij, L.ec_Hij); cp(O->ep_Hij, L.ep_Hij); cp(O->eo_Hij, L.eo_Hij);
    }
    return active_chi2(g, L);
}

// SparseOptimizer::optimize(iterations) with LM + dense solve; vertices updated in place.  Returns iterations done.
int orc_ba_optimize(int n_cam, double* cams7, const int* cam_fixed, int n_cube, double* cubes10, const int* cube_fixed, const orc_ba_edges* E,
                    int iterations, double* final_chi2) {
    Graph g; load_graph(g, n_cam, cams7, cam_fixed, n_cube, cubes10, cube_fixed); bind_edges(g, E);
    LMState st; int it = 0; double chi = 0;
    for (int i = 0; i < iterations; i++) { int r = lm_iteration(g, i, st, &chi); it++; if (r != 0) break; }
    for (int i = 0; i < n_cam; i++) se3_to_vec7(g.cams[i], cams7 + 7 * i);
    for (int i = 0; i < n_cube; i++) cuboid_to_vec10(g.cubes[i], cubes10 + 10 * i);
    if (final_chi2) *final_chi2 = chi;
    return it;
}

// helpers used by fixtures/tests to build graphs the way main_obj.cpp does
void orc_cuboid_from_minimal(const double* v9, double* out10) { cuboid_to_vec10(cuboid_from_minimal(v9), out10); }
void orc_cuboid_transform_to(const double* cube10, const double* Twc7, double* out10) { cuboid_to_vec10(transform_to(cuboid_from_vec10(cube10), se3_from_vec7(Twc7)), out10); }
void orc_cuboid_transform_from(const double* cube10, const double* Twc7, double* out10) { cuboid_to_vec10(transform_from(cuboid_from_vec10(cube10), se3_from_vec7(Twc7)), out10); }
void orc_se3_inverse(const double* a7, double* out7) { se3_to_vec7(inverse(se3_from_vec7(a7)), out7); }
void orc_se3_mul(const double* a7, const double* b7, double* out7) { se3_to_vec7(se3_from_vec7(a7) * se3_from_vec7(b7), out7); }
void orc_se3_log(const double* a7, double* out6) { se3_log(se3_from_vec7(a7), out6); }
void orc_se3_exp(const double* u6, double* out7) { se3_to_vec7(se3_exp(u6), out7); }
void orc_cuboid_min_log_error(const double* self10, const double* other10, double* out9) { min_log_error(cuboid_from_vec10(self10), cuboid_from_vec10(other10), out9); }
void orc_cuboid_project_bbox(const double* cube10, const double* Tcw7, const double* K9, double* out4) {
    M3 K; std::memcpy(K.m, K9, sizeof K.m);
    projectOntoImageBbox(cuboid_from_vec10(cube10), se3_from_vec7(Tcw7), K, out4);
}
}  // extern "C"

// Observation record of one detected cuboid for graph assembly, object_slam/src/main_obj.cpp:643-679 and :732.
// in: pos3, rotY, scale3, normalized_error, camera_roll_delta, camera_pitch_delta, transToWolrd (row-major 4x4), sampled flag.
// out: meas_quality, cube_local_meas (10)
extern "C" void orc_observation(const double* pos, double rotY, const double* scale, double normalized_error, double roll_delta, double pitch_delta,
                                const double* T16, int sampled, double* quality, double* local10) {
    double v9[9] = {pos[0], pos[1], pos[2], 0, 0, rotY, scale[0], scale[1], scale[2]};
    Cuboid ground = cuboid_from_minimal(v9);
    M3 R;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R(i, j) = T16[i * 4 + j];
    if (sampled) {
        double e0, e1, e2;
        quat_to_euler_zyx(quat_from_rot(R), e0, e1, e2);  // cam_pose_raw.euler_angle
        R = euler_zyx_to_rot(e0 + roll_delta, e1 + pitch_delta, e2);
    }
    SE3 Twc = se3_from(quat_from_rot(R), V3{T16[3], T16[7], T16[11]});
    Cuboid local = transform_to(ground, Twc);
    cuboid_to_vec10(local, local10);
    *quality = (1 - normalized_error + 0.5) / 2;
}
