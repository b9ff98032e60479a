// edlines_emul.cpp -- TEST INFRASTRUCTURE: runs the device functions of cube_slam_wu_b200/csrc/edlines_dev.cuh on the host.
//
// Every EDLines stage is a one-thread-per-item function without shared memory or synchronisation, so a loop over the items executes
// exactly the code the CUDA kernels of csrc/edlines.cu wrap.  Built by tests/test_edlines_emul.py with
//   g++ -O2 -std=c++17 -ffp-contract=off -I/usr/local/cuda/include -shared -fPIC
// and compared with oracle/oracle_edlines.cpp bit for bit.  The input is the {dx, dy} gradient image (the oracle's, pinned against cv2).
#include <vector_types.h>
#include <vector_functions.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "../../cube_slam_wu_b200/csrc/edlines_dev.cuh"
#include "../../cube_slam_wu_b200/csrc/lbd_dev.cuh"

using namespace csb;

// desc32_out (n x 32 bytes) / desc72_out (n x 72 floats) optional: detect_descrip_lines with use_LSD = false (lbd_dev.cuh on the emitted key lines)
extern "C" int emul_edlines_detect(const int16_t* dx, const int16_t* dy, int w, int h, int filter, float length_thres, float* lines_out, int max_lines,
                                   long long* stats4, int* n_chains_out, uint8_t* desc32_out, float* desc72_out) {
    const EdDims d = ed_make_dims(w, h, 1);
    const size_t npx = (size_t)w * h;
    std::vector<short2> grad(npx);
    for (size_t i = 0; i < npx; i++) grad[i] = make_short2(dx[i], dy[i]);
    std::vector<uint16_t> gd(npx);
    std::vector<uint32_t> anchors(d.anchor_words, 0u), edge(d.edge_words, 0u);
    std::vector<ushort2> p1(d.part_cap), p2(d.part_cap), chain(d.chain_cap), line(d.chain_cap);
    std::vector<int> sid(d.max_edges + 2, 0);
    std::vector<EdLine> stage(d.stage_cap);
    std::vector<float> lines((size_t)max_lines * 4, 0.f);
    std::vector<float2> keyl((size_t)max_lines);
    int n_chains = 0, n_lines = 0;
    unsigned long long stats[4] = {0, 0, 0, 0};
    EdBuffers B{};
    B.grad = grad.data(); B.gd = gd.data(); B.anchors = anchors.data(); B.edge = edge.data(); B.part1 = p1.data(); B.part2 = p2.data();
    B.chain_px = chain.data(); B.line_px = line.data(); B.chain_sid = sid.data(); B.n_chains = &n_chains; B.stage = stage.data();
    B.lines = lines.data(); B.keyl = keyl.data(); B.n_lines = &n_lines; B.stats = stats;
    for (size_t i = 0; i < npx; i++) ed_pixel(B, d, i);                         // k_ed_pixel
    for (int c = 0; c < d.nxc * d.nyc; c++)                                     // k_ed_anchor
        if (ed_anchor(B, d, 0, c)) anchors[c >> 5] |= 1u << (c & 31);
    ed_draw(B, d, 0);                                                           // k_ed_draw
    for (int c = 0; c < (n_chains > 0 ? n_chains : 0); c++) ed_fit(B, d, 0, c); // k_ed_fit
    ed_emit(B, d, 0, filter, length_thres, max_lines);                          // k_ed_emit
    const int n = n_lines < max_lines ? n_lines : max_lines;
    if (lines_out && n > 0) std::memcpy(lines_out, lines.data(), (size_t)n * 16);
    if (desc32_out && n > 0) {
        float G[LBDK_ROWS], L[3 * LBDK_BAND_W];
        lbdk_weights(G, L);
        for (int i = 0; i < n; i++)                                             // k_lbdk_line: item = line
            lbdk_line(grad.data(), w, h, lines.data() + 4 * (size_t)i, keyl[i].x, (int)keyl[i].y, G, L, desc72_out ? desc72_out + 72 * (size_t)i : nullptr,
                      desc32_out + 32 * (size_t)i);
    }
    if (stats4) for (int i = 0; i < 4; i++) stats4[i] = (long long)stats[i];
    if (n_chains_out) *n_chains_out = n_chains;
    return n_lines;
}
