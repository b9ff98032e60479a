// abi_smoke.cpp -- the C ABI used from plain C++ (no Eigen / OpenCV / Python): what a maintainer's adapter does, reduced to one file.
//
//   g++ -std=c++17 -I include integration/examples/abi_smoke.cpp cube_slam_wu_b200/libcubeslam_b200.so -Wl,-rpath,$PWD/cube_slam_wu_b200 -o abi_smoke
//
// Draws one synthetic gray frame, runs the line detector (csb_lsd_detect_batch) and the line descriptors (csb_lbd_*), feeds the segments and the gray frame to the cuboid
// proposal path (csb_detect_plan + csb_detect_batch_gray), then linearises a two-camera / one-cuboid graph in both Jacobian modes
// (csb_ba_set_graph / csb_ba_linearize).  Exit code 0 = every call returned CSB_OK and the results are sane; 77 = no CUDA device
// (there is no CPU fallback: csb_create fails); anything else = failure.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cubeslam_b200.h"

#define CHECK(call)                                                                                   \
    do {                                                                                              \
        int rc__ = (call);                                                                            \
        if (rc__ != CSB_OK) {                                                                         \
            std::fprintf(stderr, "%s -> %d (%s)\n", #call, rc__, ctx ? csb_last_error(ctx) : "");   \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

static void fill_quad(std::vector<uint8_t>& img, int W, int H, const double (*p)[2], uint8_t v) {  // convex quadrilateral, scan conversion
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            bool in = true;
            for (int k = 0; k < 4 && in; k++) {
                const double ax = p[k][0], ay = p[k][1], bx = p[(k + 1) & 3][0], by = p[(k + 1) & 3][1];
                in = (bx - ax) * (y - ay) - (by - ay) * (x - ax) >= 0;
            }
            if (in) img[(size_t)y * W + x] = v;
        }
}

int main() {
    csb_context* ctx = nullptr;
    if (csb_create(&ctx, 0) != CSB_OK) {
        std::printf("abi_smoke: no usable CUDA device -- csb_create failed, nothing was computed (%s)\n", csb_version());
        return 77;
    }
    const int W = 1242, H = 375;
    std::vector<uint8_t> gray((size_t)W * H, 90);
    const double car[4][2] = {{420, 150}, {640, 160}, {630, 300}, {410, 285}}, roof[4][2] = {{460, 110}, {600, 118}, {640, 160}, {420, 150}};
    const double wall[4][2] = {{800, 40}, {1100, 60}, {1090, 330}, {790, 300}};
    fill_quad(gray, W, H, car, 190);
    fill_quad(gray, W, H, roof, 40);
    fill_quad(gray, W, H, wall, 140);

    // ---- line_lbd_detect::detect_filter_lines
    csb_lsd_params lp{15.0f, 1, 1024, 0};
    std::vector<float> seg((size_t)lp.max_lines * 4);
    int32_t n_seg = 0;
    csb_lsd_stats ls{};
    CHECK(csb_lsd_detect_batch(ctx, gray.data(), 1, W, H, &lp, seg.data(), &n_seg, &ls));
    std::printf("abi_smoke: %d line segments (%lld regions grown, %d kernel launches)\n", n_seg, (long long)ls.n_regions, ls.n_kernel_launches);
    if (n_seg < 8) return 2;

    // ---- line_lbd_detect::detect_descrip_lines: descriptors of those segments, read in place on the device; the same rows again through
    //      the host-buffer entry point must give the same bytes
    csb_lbd_stats bs{};
    std::vector<uint8_t> desc((size_t)n_seg * 32), desc2((size_t)n_seg * 32);
    CHECK(csb_lbd_run_on_lsd(ctx, 0, 0));
    CHECK(csb_lbd_download(ctx, desc.data(), nullptr, nullptr, nullptr, n_seg, &bs));
    const int32_t off[2] = {0, n_seg};
    CHECK(csb_lbd_describe_batch(ctx, gray.data(), 1, W, H, seg.data(), off, desc2.data(), nullptr, nullptr));
    std::printf("abi_smoke: %lld LBD descriptors (%lld gradient samples)\n", (long long)bs.n_lines, (long long)bs.n_samples);
    if (bs.n_lines != n_seg || std::memcmp(desc.data(), desc2.data(), desc.size()) != 0) return 6;

    // ---- detect_3d_cuboid::detect_cuboid
    csb_detect_params dp{1, 1, 1, 0, 1, 0, 1.0, 3.0};
    csb_frame fr{};
    const double K[9] = {718.856, 0, 607.19, 0, 718.856, 185.22, 0, 0, 1};
    const double T[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, -1, 0, 1.65, 0, 0, 0, 1};  // camera 1.65 m above the ground, looking along world y
    std::memcpy(fr.Kalib, K, sizeof K);
    std::memcpy(fr.transToWolrd, T, sizeof T);
    fr.img_width = W; fr.img_height = H;
    fr.box_begin = 0; fr.box_end = 1; fr.line_begin = 0; fr.line_end = n_seg;
    const double box[5] = {405, 105, 240, 200, 0.9};
    std::vector<double> lines((size_t)n_seg * 4);
    for (int i = 0; i < 4 * n_seg; i++) lines[i] = seg[i];  // float -> double like main_obj.cpp:596-599
    int n_tasks = 0;
    int64_t n_map = 0;
    CHECK(csb_detect_plan(&fr, 1, box, 1, &dp, nullptr, 0, &n_tasks, &n_map));
    std::vector<csb_task> tasks(n_tasks > 0 ? n_tasks : 1);
    CHECK(csb_detect_plan(&fr, 1, box, 1, &dp, tasks.data(), n_tasks, &n_tasks, &n_map));
    csb_cuboid cub{};
    int32_t n_cub = 0;
    csb_detect_stats ds{};
    CHECK(csb_detect_batch_gray(ctx, &fr, 1, box, 1, lines.data(), n_seg, tasks.data(), n_tasks, gray.data(), (int64_t)gray.size(), &dp, &cub, &n_cub, &ds));
    std::printf("abi_smoke: %lld hypotheses enumerated, %lld scored, %d cuboid(s)", (long long)ds.n_enumerated, (long long)ds.n_scored, n_cub);
    if (n_cub) std::printf("; best: pos %.3f %.3f %.3f scale %.3f %.3f %.3f yaw %.3f", cub.pos[0], cub.pos[1], cub.pos[2], cub.scale[0], cub.scale[1], cub.scale[2], cub.rotY);
    std::printf("\n");
    if (ds.n_enumerated <= 0) return 3;

    // ---- BlockSolver::buildSystem on a tiny graph: 2 cameras, 1 cuboid, 2 EdgeSE3Cuboid, 1 EdgeSE3Expmap
    const int32_t cam_fixed[2] = {1, 0}, cube_fixed[1] = {0};
    const int32_t ec_cam[2] = {0, 1}, ec_cube[2] = {0, 0}, eo_i[1] = {0}, eo_j[1] = {1};
    const double cams[14] = {0, 0, 0, 0, 0, 0, 1, 0.1, 0, 0, 0, 0, 0, 1};
    const double cubes[10] = {0.2, 0.1, 3.0, 0, 0, 0.0998334166, 0.9950041653, 0.5, 0.4, 0.3};
    double ec_meas[20], ec_info[162] = {0}, eo_meas[7] = {0.1, 0, 0, 0, 0, 0, 1}, eo_info[36] = {0};
    for (int e = 0; e < 2; e++) {
        for (int k = 0; k < 10; k++) ec_meas[10 * e + k] = cubes[k] + 0.01 * (k + e);
        for (int k = 0; k < 9; k++) ec_info[81 * e + 10 * k] = 4.0;
    }
    for (int k = 0; k < 6; k++) eo_info[7 * k] = 1.0;
    csb_ba_graph g{};
    g.n_cam = 2; g.n_cube = 1; g.cam_fixed = cam_fixed; g.cube_fixed = cube_fixed;
    g.n_ec = 2; g.ec_cam = ec_cam; g.ec_cube = ec_cube; g.ec_meas = ec_meas; g.ec_info = ec_info;
    g.n_eo = 1; g.eo_cam_i = eo_i; g.eo_cam_j = eo_j; g.eo_meas = eo_meas; g.eo_info = eo_info;
    CHECK(csb_ba_set_graph(ctx, &g));
    double Hn[81], Ha[81], chi_n = 0, chi_a = 0;
    csb_ba_output o{};
    o.H_cube = Hn; o.chi2 = &chi_n;
    CHECK(csb_ba_linearize(ctx, cams, cubes, &o));
    CHECK(csb_ba_set_jacobian_mode(ctx, CSB_BA_JACOBIAN_ANALYTIC));
    o.H_cube = Ha; o.chi2 = &chi_a;
    CHECK(csb_ba_linearize(ctx, cams, cubes, &o));
    double worst = 0, scale = 1;
    for (int k = 0; k < 81; k++) { worst = std::fmax(worst, std::fabs(Hn[k] - Ha[k])); scale = std::fmax(scale, std::fabs(Hn[k])); }
    std::printf("abi_smoke: BA chi2 %.6g; cuboid Hessian block, numeric vs closed-form Jacobians: %.3g of its scale\n", chi_n, worst / scale);
    if (!(chi_n > 0) || chi_n != chi_a || worst / scale > 1e-4) return 4;
    csb_destroy(ctx);
    std::printf("abi_smoke: ok\n");
    return 0;
}
