// distmap.cu -- per-ROI Canny + 3x3 chamfer distance transform on the GPU (SURVEY.md 8 "next" row f-1).
//
// Replaces the two OpenCV calls of detect_3d_cuboid/src/box_proposal_detail.cpp:320-327
//     cv::Canny(gray_img(object_bbox), im_canny, 80, 200);
//     cv::distanceTransform(255 - im_canny, dist_map, CV_DIST_L2, 3);
// so that the batch entry points can start from the gray frames instead of caller-computed distance maps.
// OpenCV is not part of the reference tree; the algorithms restated here are OpenCV's own open-source ones
// (imgproc/src/canny.cpp: Sobel 3x3 -> L1 magnitude -> NMS with the tan(22.5) integer test -> hysteresis;
//  imgproc/src/distransform.cpp distanceTransform_3x3: 16.16 fixed-point two-pass chamfer, a = 0.955, b = 1.3693),
// pinned bit-exactly against python cv2 4.13 with IPP disabled (tests/test_distmap_gpu.py).  (A cv2 build WITH IPP replaces
// the distance transform by a closed-source float variant that differs by <= 4e-4.)
// Semantics of cv::Canny on a cv::Mat ROI view: the Sobel filter sees the real neighbours of the ROI inside the parent image
// (BORDER_REPLICATE only at the image border); magnitude outside the ROI counts as 0; edges do not connect across the ROI bound.
//
//   k_canny   : one CTA per task.  The ROI is processed in strips of CN_ROWS rows: gray strip (2-pixel halo) -> shared memory;
//               Sobel + L1 magnitude of the strip (one thread per 4 pixels, column sums shared between neighbours) -> shared
//               memory; non-maximum suppression -> 2 bits per pixel in a packed map that stays in shared memory for the whole
//               ROI; then breadth-first hysteresis on that map (work list in global memory, shared-memory atomics); the packed
//               map goes out with coalesced word stores.
//   k_dist3x3 : one CTA (4 warps) per task: forward / backward chamfer passes; each row is a min-plus prefix scan
//               d[j] = min_m (c[m] + a (j - m)) done with warp shuffles + one shared-memory exchange, the previous row stays in
//               registers.  The edge bits come from shared memory (packed map loaded once); the forward result is written to
//               global memory row by row and read back for the backward pass in strips of DT_STRIP rows by TMA bulk copies
//               (cp.async.bulk + mbarrier, double buffered), so no row ever waits for a global load.
//
// Packed map layout (global `cmap` and shared): row pitch = WPR = ceil(W / 16) words, pixel (r, c) = bits [2 (c & 15), +2) of word
// r * WPR + (c >> 4); 0 = weak candidate, 1 = no edge, 2 = edge.  A task's words start at byte 4 * map_offset of `cmap`.
#include <cuda_runtime.h>

#include <cstdint>

#include "context.h"
#include "ptx_helpers.cuh"

namespace csb {

constexpr int CN_THREADS = 256;
constexpr int CN_ROWS = 16;

// Sobel 3x3 (cv::Sobel ksize 3, scale 1) at strip position (y, x) of the byte strip g
__device__ __forceinline__ void sobel_s(const uint8_t* g, int pitch, int y, int x, int& dx, int& dy) {
    const uint8_t* p = g + (y - 1) * pitch + (x - 1);
    const int a = p[0], b = p[1], c = p[2];
    const int d = p[pitch], f = p[pitch + 2];
    const int q = p[2 * pitch], h = p[2 * pitch + 1], i = p[2 * pitch + 2];
    dx = (c + 2 * f + i) - (a + 2 * d + q);
    dy = (q + 2 * h + i) - (a + 2 * b + c);
}

__global__ void __launch_bounds__(CN_THREADS) k_canny(DetectBuffers B, const uint8_t* gray, uint8_t* cmap, int* queue, int low, int high, int pm_words_cap, int wp_cap) {
    extern __shared__ __align__(16) unsigned char cn_smem[];
    const int task = blockIdx.x, tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;  // 8 row groups x 32 column groups
    const TaskTab t = B.ttab[task];
    const FrameTab& ft = B.ftab[t.frame_id];
    const uint8_t* img = gray + ft.gray_offset;
    const int IW = ft.img_w, IH = ft.img_h;
    const int W = t.roi_w, H = t.roi_h;
    const int Wp = (W + 15) & ~15, WPR = Wp >> 4;
    const int GP = wp_cap + 8;   // gray strip pitch (bytes): column index = ROI column + 4
    const int MP = wp_cap + 8;   // magnitude strip pitch (u16)
    unsigned* pm = reinterpret_cast<unsigned*>(cn_smem);                                   // packed map, H * WPR words
    uint8_t* gs = cn_smem + 4 * (size_t)pm_words_cap;                                      // (CN_ROWS + 4) x GP bytes
    unsigned short* ms = reinterpret_cast<unsigned short*>(gs + (size_t)(CN_ROWS + 4) * GP);  // (CN_ROWS + 2) x MP u16
    unsigned* gs32 = reinterpret_cast<unsigned*>(gs);
    unsigned* ms32 = reinterpret_cast<unsigned*>(ms);
    uint8_t* pm8 = reinterpret_cast<uint8_t*>(pm);
    __shared__ int s_tail, s_head, s_end;
    int* q = queue + t.map_offset;
    const int TG22 = (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5);
    if (tid == 0) { s_tail = 0; s_head = 0; }
    const int G4 = (Wp + 8) >> 2;  // 4-byte groups per gray strip row that are actually used
    const int NG = Wp >> 2;        // 4-pixel groups per ROI row
    for (int r0 = 0; r0 < H; r0 += CN_ROWS) {
        const int rows = min(CN_ROWS, H - r0);
        __syncthreads();  // previous strip fully consumed
        // (1) gray strip: strip row sr <-> ROI row r0 - 2 + sr, byte column u <-> ROI column u - 4; BORDER_REPLICATE at the image border
        for (int sr = ty; sr < rows + 4; sr += 8) {
            const int y = min(max(t.roi_top + r0 - 2 + sr, 0), IH - 1);
            const uint8_t* row = img + (size_t)y * IW;
            for (int g = tx; g < G4; g += 32) {
                const int x0 = t.roi_left + 4 * g - 4;
                unsigned w = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) w |= (unsigned)row[min(max(x0 + k, 0), IW - 1)] << (8 * k);
                gs32[(sr * GP >> 2) + g] = w;
            }
        }
        __syncthreads();
        // (2) L1 gradient magnitude of ROI rows r0 - 1 .. r0 + rows (0 outside the ROI: the zero border of Canny's magnitude buffer);
        //     strip row mr <-> ROI row r0 - 1 + mr, u16 column index = ROI column + 4
        for (int mr = ty; mr < rows + 2; mr += 8) {
            const int r = r0 - 1 + mr;
            const bool row_in = r >= 0 && r < H;
            if (tx == 0) { ms32[(mr * MP >> 1) + 0] = 0; ms32[(mr * MP >> 1) + 1] = 0; }  // columns -4 .. -1
            for (int g = tx; g < NG + 1; g += 32) {
                unsigned lo = 0, hi = 0;
                if (row_in && g < NG) {
                    int s1[6], tt[6];
                    const unsigned* r0p = gs32 + ((mr + 0) * GP >> 2) + g;  // gray rows r-1, r, r+1 are strip rows mr, mr+1, mr+2
                    const unsigned* r1p = gs32 + ((mr + 1) * GP >> 2) + g;
                    const unsigned* r2p = gs32 + ((mr + 2) * GP >> 2) + g;
                    const unsigned a0 = r0p[0], a1 = r0p[1], a2 = r0p[2], b0 = r1p[0], b1 = r1p[1], b2 = r1p[2], c0w = r2p[0], c1 = r2p[1], c2 = r2p[2];
#pragma unroll
                    for (int x = 0; x < 6; x++) {
                        // window column x <-> ROI column 4g - 1 + x: byte 3 of word g, bytes 0..3 of word g+1, byte 0 of word g+2
                        const int top = (x == 0) ? (a0 >> 24) : (x == 5) ? (a2 & 255u) : ((a1 >> (8 * (x - 1))) & 255u);
                        const int mid = (x == 0) ? (b0 >> 24) : (x == 5) ? (b2 & 255u) : ((b1 >> (8 * (x - 1))) & 255u);
                        const int bot = (x == 0) ? (c0w >> 24) : (x == 5) ? (c2 & 255u) : ((c1 >> (8 * (x - 1))) & 255u);
                        s1[x] = top + 2 * mid + bot;
                        tt[x] = bot - top;
                    }
                    unsigned m4[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const int dx = s1[k + 2] - s1[k], dy = tt[k] + 2 * tt[k + 1] + tt[k + 2];
                        m4[k] = (4 * g + k < W) ? (unsigned)(abs(dx) + abs(dy)) : 0u;
                    }
                    lo = m4[0] | (m4[1] << 16); hi = m4[2] | (m4[3] << 16);
                }
                const int wi = (mr * MP >> 1) + 2 + 2 * g;  // u16 column 4g + 4
                ms32[wi] = lo; ms32[wi + 1] = hi;           // g == NG writes the zero columns Wp .. Wp + 3
            }
        }
        __syncthreads();
        // (3) non-maximum suppression of ROI rows r0 .. r0 + rows - 1; one thread per 4 pixels = one byte of the packed map
        for (int rr = ty; rr < rows; rr += 8) {
            const int r = r0 + rr, mr = rr + 1;
            for (int g = tx; g < NG; g += 32) {
                // magnitudes of columns 4g - 2 .. 4g + 5 on rows r-1, r, r+1 (u16 columns 4g + 2 .. 4g + 9 -> words 2g + 1 .. 2g + 4)
                unsigned up[4], md[4], dn[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    up[k] = ms32[((mr - 1) * MP >> 1) + 2 * g + 1 + k];
                    md[k] = ms32[(mr * MP >> 1) + 2 * g + 1 + k];
                    dn[k] = ms32[((mr + 1) * MP >> 1) + 2 * g + 1 + k];
                }
                auto at = [](const unsigned* w, int x) -> int { return (int)((w[x >> 1] >> (16 * (x & 1))) & 0xffffu); };  // x = column - (4g - 2)
                unsigned bits = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int c = 4 * g + k, x = k + 2;
                    const int m = at(md, x);
                    unsigned v = 1;
                    if (m > low) {  // implies c < W (magnitude 0 beyond the ROI)
                        int xs, ys;
                        sobel_s(gs, GP, rr + 2, c + 4, xs, ys);
                        const int ax = abs(xs), ay = abs(ys) << 15;
                        const int tg22x = ax * TG22;
                        bool keep;
                        if (ay < tg22x) keep = (m > at(md, x - 1)) && (m >= at(md, x + 1));
                        else {
                            const int tg67x = tg22x + (ax << 16);
                            if (ay > tg67x) keep = (m > at(up, x)) && (m >= at(dn, x));
                            else {
                                const bool neg = (xs ^ ys) < 0;  // s = -1
                                const int mu = neg ? at(up, x + 1) : at(up, x - 1), mdn = neg ? at(dn, x - 1) : at(dn, x + 1);
                                keep = (m > mu) && (m > mdn);
                            }
                        }
                        if (keep) {
                            v = (m > high) ? 2u : 0u;
                            if (v == 2u) q[atomicAdd(&s_tail, 1)] = (r << 16) | c;
                        }
                    }
                    bits |= v << (2 * k);
                }
                pm8[(size_t)r * (WPR * 4) + g] = (uint8_t)bits;
            }
        }
    }
    __syncthreads();
    // (4) hysteresis: breadth-first promotion of weak pixels 8-connected to edge pixels
    while (true) {
        if (tid == 0) s_end = s_tail;
        __syncthreads();
        const int head = s_head, end = s_end;
        if (head >= end) break;
        for (int e = head + tid; e < end; e += CN_THREADS) {
            const int rc = q[e];
            const int r = rc >> 16, c = rc & 0xffff;
#pragma unroll
            for (int dr = -1; dr <= 1; dr++)
#pragma unroll
                for (int dc = -1; dc <= 1; dc++) {
                    const int rr = r + dr, cc = c + dc;
                    if ((dr | dc) == 0 || rr < 0 || cc < 0 || rr >= H || cc >= W) continue;
                    unsigned* w = pm + rr * WPR + (cc >> 4);
                    const int sh = 2 * (cc & 15);
                    if (((*w >> sh) & 3u) != 0) continue;
                    const unsigned old = atomicOr(w, 2u << sh);
                    if (((old >> sh) & 3u) == 0) q[atomicAdd(&s_tail, 1)] = (rr << 16) | cc;
                }
        }
        __syncthreads();
        if (tid == 0) s_head = end;
        __syncthreads();
    }
    // (5) packed map -> global
    unsigned* out = reinterpret_cast<unsigned*>(cmap + 4 * (size_t)t.map_offset);
    for (int i = tid; i < H * WPR; i += CN_THREADS) out[i] = pm[i];
}

constexpr unsigned DT_HV = 62587u;                 // cvRound(0.955f  * 65536)
constexpr unsigned DT_DG = 89738u;                 // cvRound(1.3693f * 65536)
constexpr unsigned DT_MAX = 0xffffffffu - DT_DG;   // DIST_MAX; also used for the border cells (behaves like OpenCV's INIT_DIST0)
constexpr int DT_WARPS = 4;            // warps per CTA = per task
constexpr int DT_THREADS = 32 * DT_WARPS;
constexpr int DT_STRIP = 8;            // rows per TMA strip of the backward pass
constexpr int DT_INF = 0x3fffffff;     // "no path yet" inside the kernel (32-bit arithmetic); becomes DT_MAX on output

// One CTA (4 warps) per task; thread t owns columns [t*CH, t*CH + CH).  The previous row lives in registers; each row is a
// min-plus scan:  forward  d[j] = min_{m<=j} (c[m] - a m) + a j,   backward  d[j] = min_{m>=j} (c[m] + a m) - a j
// (a = DT_HV; c = 0 on edge pixels, else the 3-neighbour minimum over the already finished adjacent row): sequential inside the
// thread's chunk, warp shuffles across lanes, one shared-memory exchange across the 4 warps; two barriers per row.
// 32-bit arithmetic: OpenCV saturates unreachable cells at DIST_MAX (~2^32); with at least one edge pixel in the ROI every final
// value is a real path length (< 2^27 for ROIs up to 1280 px), and saturated cells only ever lose comparisons, so any "infinity"
// that survives the additions gives the same result.  A ROI without edge pixels ends at DT_INF everywhere -> DIST_MAX, as in OpenCV.
// The chunk size CH (2 / 3 / 4 / 10 columns per thread) is chosen per task from its ROI width; all classes run in one launch.
template <int CH>
__device__ __forceinline__ void dist3x3_cta(const TaskTab& t, const uint8_t* cmap, int* dtmp, float* maps, int tid, int (*s_tot)[DT_WARPS], int (*s_edge)[2 * DT_WARPS],
                                            unsigned* s_pm, int* s_strip, int buf_stride, uint64_t* s_bar) {
    const int W = t.roi_w, H = t.roi_h;
    const int WPR = (W + 15) >> 4;
    int* tmp = dtmp + t.map_offset;
    float* out = maps + t.map_offset;
    const float scale = 1.f / 65536.f;
    const int HV = (int)DT_HV, DG = (int)DT_DG;
    const int lane = tid & 31, warp = tid >> 5;
    const int j0 = tid * CH;
    const unsigned FULL = 0xffffffffu;
    {
        const unsigned* gpm = reinterpret_cast<const unsigned*>(cmap + 4 * (size_t)t.map_offset);
        for (int i = tid; i < H * WPR; i += DT_THREADS) s_pm[i] = gpm[i];
    }
    int prev[CH];
    // ---- forward pass
#pragma unroll
    for (int k = 0; k < CH; k++) prev[k] = DT_INF;  // row -1 = border
    if (tid < 2 * DT_WARPS) { s_edge[0][tid] = DT_INF; s_edge[1][tid] = DT_INF; }
    __syncthreads();
    for (int i = 0; i < H; i++) {
        const int par = i & 1;
        // neighbours' boundary cells of the previous row: by shuffle inside the warp, through s_edge across warps
        int pl = __shfl_up_sync(FULL, prev[CH - 1], 1), pr = __shfl_down_sync(FULL, prev[0], 1);
        if (lane == 0) pl = (warp == 0) ? DT_INF : s_edge[par ^ 1][2 * (warp - 1) + 1];
        if (lane == 31) pr = (warp == DT_WARPS - 1) ? DT_INF : s_edge[par ^ 1][2 * (warp + 1)];
        int v[CH];
        int run = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int j = j0 + k;
            int x = 0x7fffffff;
            if (j < W) {
                int c;
                if (((s_pm[i * WPR + (j >> 4)] >> (2 * (j & 15))) & 3u) == 2u) c = 0;
                else {
                    const int ul = (k > 0) ? prev[k - 1] : pl;
                    const int ur = (j + 1 < W) ? ((k + 1 < CH) ? prev[k + 1] : pr) : DT_INF;
                    c = min(min(ul + DG, prev[k] + HV), min(ur + DG, DT_INF));
                }
                x = c - HV * j;
            }
            run = min(run, x);
            v[k] = run;  // inclusive prefix min inside the chunk
        }
        int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc = min(inc, y); }
        if (lane == 31) s_tot[par][warp] = inc;
        int excl = __shfl_up_sync(FULL, inc, 1);
        if (lane == 0) excl = 0x7fffffff;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < DT_WARPS - 1; w++) if (w < warp) excl = min(excl, s_tot[par][w]);
        excl = min(excl, DT_INF + HV);  // left border cell (column -1)
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int j = j0 + k;
            const int dv = min(min(excl, v[k]) + HV * j, DT_INF);
            prev[k] = (j < W) ? dv : DT_INF;
            if (j < W) tmp[(size_t)i * W + j] = dv;
        }
        if (lane == 0) s_edge[par][2 * warp] = prev[0];
        if (lane == 31) s_edge[par][2 * warp + 1] = prev[CH - 1];
        __syncthreads();
    }
    // ---- backward pass: the forward result comes back in strips of DT_STRIP rows (bottom strip first), double buffered.
    // Strip s covers rows [s*DT_STRIP, min(H, (s+1)*DT_STRIP)); its first int sits at tmp + s*DT_STRIP*W: 16-byte aligned because
    // map_offset and DT_STRIP*W are multiples of 4 ints.  Byte counts are rounded up to 16 (the map slots are padded to 16 bytes).
    const int n_strips = (H + DT_STRIP - 1) / DT_STRIP;
    auto issue = [&](int s, int b) {
        const int r_lo = s * DT_STRIP, r_hi = min(H, r_lo + DT_STRIP);
        const uint32_t bytes = ((uint32_t)((r_hi - r_lo) * W) * 4u + 15u) & ~15u;
        fence_proxy_async();  // generic-proxy reads of this buffer (previous strip) are ordered before the async-proxy write
        mbar_expect_tx(&s_bar[b], bytes);
        tma_bulk_g2s(s_strip + b * buf_stride, tmp + (size_t)r_lo * W, bytes, &s_bar[b]);
    };
#pragma unroll
    for (int k = 0; k < CH; k++) prev[k] = DT_INF;  // row H = border
    if (tid < 2 * DT_WARPS) { s_edge[0][tid] = DT_INF; s_edge[1][tid] = DT_INF; }
    __threadfence();
    fence_proxy_async_all();  // this thread's forward stores (generic proxy, global) before the bulk reads of the async proxy
    __syncthreads();
    uint32_t ph0 = 0, ph1 = 0;
    if (tid == 0) issue(n_strips - 1, 0);
    for (int s = n_strips - 1, it = 0; s >= 0; s--, it++) {
        const int b = it & 1;
        if (tid == 0 && s > 0) issue(s - 1, b ^ 1);  // buffer b^1 was released by the trailing barrier of the previous strip
        if (b == 0) { mbar_wait(&s_bar[0], ph0); ph0 ^= 1; } else { mbar_wait(&s_bar[1], ph1); ph1 ^= 1; }
        const int r_lo = s * DT_STRIP, r_hi = min(H, r_lo + DT_STRIP);
        const int* sb = s_strip + b * buf_stride;
        for (int i = r_hi - 1; i >= r_lo; i--) {
            const int par = i & 1;
            int pl = __shfl_up_sync(FULL, prev[CH - 1], 1), pr = __shfl_down_sync(FULL, prev[0], 1);
            if (lane == 0) pl = (warp == 0) ? DT_INF : s_edge[par ^ 1][2 * (warp - 1) + 1];
            if (lane == 31) pr = (warp == DT_WARPS - 1) ? DT_INF : s_edge[par ^ 1][2 * (warp + 1)];
            int v[CH];
            int run = 0x7fffffff;
#pragma unroll
            for (int k = CH - 1; k >= 0; k--) {
                const int j = j0 + k;
                int x = 0x7fffffff;
                if (j < W) {
                    const int self = sb[(i - r_lo) * W + j];
                    const int dl = (k > 0) ? prev[k - 1] : pl;
                    const int dr = (j + 1 < W) ? ((k + 1 < CH) ? prev[k + 1] : pr) : DT_INF;
                    const int t0 = min(min(self, dr + DG), min(prev[k] + HV, dl + DG));
                    x = t0 + HV * j;
                }
                run = min(run, x);
                v[k] = run;  // inclusive suffix min inside the chunk
            }
            int inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_down_sync(FULL, inc, o); if (lane + o < 32) inc = min(inc, y); }
            if (lane == 0) s_tot[par][warp] = inc;
            int excl = __shfl_down_sync(FULL, inc, 1);
            if (lane == 31) excl = 0x7fffffff;
            __syncthreads();
#pragma unroll
            for (int w = 1; w < DT_WARPS; w++) if (w > warp) excl = min(excl, s_tot[par][w]);
            excl = min(excl, DT_INF + HV * W);  // right border cell (column W)
#pragma unroll
            for (int k = 0; k < CH; k++) {
                const int j = j0 + k;
                const int dv = min(min(excl, v[k]) - HV * j, DT_INF);
                prev[k] = (j < W) ? dv : DT_INF;
                if (j < W) out[(size_t)i * W + j] = (float)((dv >= DT_INF) ? DT_MAX : (unsigned)dv) * scale;
            }
            if (lane == 0) s_edge[par][2 * warp] = prev[0];
            if (lane == 31) s_edge[par][2 * warp + 1] = prev[CH - 1];
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(DT_THREADS) k_dist3x3(DetectBuffers B, const uint8_t* cmap, int* dtmp, float* maps, int pm_words_cap, int strip_cap) {
    extern __shared__ __align__(128) unsigned char dt_smem[];
    __shared__ int s_tot[2][DT_WARPS];
    __shared__ int s_edge[2][2 * DT_WARPS];
    __shared__ uint64_t s_bar[2];
    int* s_strip = reinterpret_cast<int*>(dt_smem);                          // 2 buffers of strip_cap ints (16-byte aligned)
    unsigned* s_pm = reinterpret_cast<unsigned*>(dt_smem) + 2 * strip_cap;   // packed map
    const int task = blockIdx.x, tid = threadIdx.x;
    const TaskTab t = B.ttab[task];
    if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); fence_mbar_init(); }
    __syncthreads();
    if (t.roi_w <= DT_THREADS * 2) dist3x3_cta<2>(t, cmap, dtmp, maps, tid, s_tot, s_edge, s_pm, s_strip, strip_cap, s_bar);
    else if (t.roi_w <= DT_THREADS * 3) dist3x3_cta<3>(t, cmap, dtmp, maps, tid, s_tot, s_edge, s_pm, s_strip, strip_cap, s_bar);
    else if (t.roi_w <= DT_THREADS * 4) dist3x3_cta<4>(t, cmap, dtmp, maps, tid, s_tot, s_edge, s_pm, s_strip, strip_cap, s_bar);
    else dist3x3_cta<10>(t, cmap, dtmp, maps, tid, s_tot, s_edge, s_pm, s_strip, strip_cap, s_bar);
}

// ---- gray-frame upload without the copy engine: only what the ROIs read -----------------------------------------------------------
// The packed gray frames of a batch are one linear byte array, cut into 128-byte segments (8 lanes x 16 bytes).  k_gray_mark sets
// the bit of every segment that holds a pixel of some task's ROI or of its one-pixel Sobel halo (clamped to the image); k_gray_gather
// copies the marked segments from the caller's pinned (device-mapped) buffer into the device frame buffer -- SM-initiated PCIe reads
// run at the copy engine's rate (measured: 50 GB/s against 55), so the upload shrinks with the share of the frames the boxes cover.
// Everything outside the marked segments is never read by k_canny in a way that reaches a result (magnitudes outside the ROI count as 0).
// The gather is a persistent kernel of small CTAs (64 threads, <= 32 registers, no shared memory) so that it finds room next to the
// one-CTA-per-SM scoring kernel of another context and the transfer overlaps compute the way a copy-engine transfer would.
constexpr int GSEG = 128;
constexpr int GG_THREADS = 64;

__global__ void __launch_bounds__(128) k_gray_mark(DetectBuffers B, unsigned* seg_bits) {
    const int task = blockIdx.x;
    const TaskTab t = B.ttab[task];
    const FrameTab& ft = B.ftab[t.frame_id];
    const int IW = ft.img_w, IH = ft.img_h;
    const int y0 = max(t.roi_top - 1, 0), y1 = min(t.roi_top + t.roi_h, IH - 1);      // inclusive
    const int x0 = max(t.roi_left - 1, 0), x1 = min(t.roi_left + t.roi_w, IW - 1);
    for (int y = y0 + threadIdx.x; y <= y1; y += blockDim.x) {
        const long long a = (ft.gray_offset + (long long)y * IW + x0) / GSEG, b = (ft.gray_offset + (long long)y * IW + x1) / GSEG;
        // consecutive segments of one row: set whole runs of bits per bitmap word
        for (long long w = a >> 5; w <= (b >> 5); w++) {
            const int lo = (w == (a >> 5)) ? (int)(a & 31) : 0, hi = (w == (b >> 5)) ? (int)(b & 31) : 31;
            const unsigned m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
            if ((seg_bits[w] & m) != m) atomicOr(seg_bits + w, m);
        }
    }
}

// Each warp walks bitmap words (32 segments = 4 KB of frame); a load instruction moves four segments (lane / 8 = segment, lane % 8 = chunk),
// two load instructions are in flight per lane before the stores.
__global__ void __launch_bounds__(GG_THREADS, 8) k_gray_gather(const uint4* __restrict__ src, uint4* __restrict__ dst, const unsigned* __restrict__ seg_bits, int n_words,
                                                             long long n_chunks16, int* n_segments_out) {
    const int lane = threadIdx.x & 31, sub = lane >> 3, ch = lane & 7;
    const int warp = (blockIdx.x * GG_THREADS + threadIdx.x) >> 5, n_warps = (gridDim.x * GG_THREADS) >> 5;
    int n_mine = 0;
    for (int w = warp; w < n_words; w += n_warps) {
        unsigned m = seg_bits[w];
        n_mine += __popc(m);
        while (m) {
            long long o[2];
            uint4 v[2];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                // the (sub + 1)-th set bit of m, then drop up to four bits
                const unsigned b = __fns(m, 0, sub + 1);
                o[k] = (b < 32u) ? ((long long)w * 32 + b) * (GSEG / 16) + ch : -1;
                if (o[k] >= n_chunks16) o[k] = -1;
#pragma unroll
                for (int q = 0; q < 4; q++) m &= m - 1;
                if (o[k] >= 0) v[k] = src[o[k]];
            }
#pragma unroll
            for (int k = 0; k < 2; k++)
                if (o[k] >= 0) dst[o[k]] = v[k];
        }
    }
    if (lane == 0 && n_mine) atomicAdd(n_segments_out, n_mine);
}

cudaError_t launch_gray_gather(const DetectBuffers& B, const uint8_t* gray_host_mapped, uint8_t* gray_dev, long long n_bytes, unsigned* seg_bits, int* n_segments_out,
                               int num_sms, cudaStream_t st) {
    if (B.n_tasks == 0 || n_bytes <= 0) return cudaSuccess;
    const long long n_seg = (n_bytes + GSEG - 1) / GSEG;
    const int n_words = (int)((n_seg + 31) / 32);
    cudaError_t e = cudaMemsetAsync(seg_bits, 0, 4 * (size_t)n_words, st);
    if (e != cudaSuccess) return e;
    k_gray_mark<<<B.n_tasks, 128, 0, st>>>(B, seg_bits);
    k_gray_gather<<<2 * num_sms, GG_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(gray_host_mapped), reinterpret_cast<uint4*>(gray_dev), seg_bits, n_words,
                                                     n_bytes / 16, n_segments_out);
    return cudaGetLastError();
}

cudaError_t launch_distmaps(const DetectBuffers& B, const uint8_t* gray, uint8_t* cmap, int* queue, unsigned* dtmp, float* maps, int max_roi_w, int max_pm_words, cudaStream_t st) {
    if (B.n_tasks == 0) return cudaSuccess;
    if (max_roi_w > DT_THREADS * 10) return cudaErrorInvalidValue;  // ROI wider than 1280 px
    const int wp_cap = (max_roi_w + 15) & ~15;
    const int pm_cap = (max_pm_words + 3) & ~3;
    const size_t smem_c = 4 * (size_t)pm_cap + (size_t)(CN_ROWS + 4) * (wp_cap + 8) + 2 * (size_t)(CN_ROWS + 2) * (wp_cap + 8);
    if (smem_c > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_canny, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c);
    if (e != cudaSuccess) return e;
    k_canny<<<B.n_tasks, CN_THREADS, smem_c, st>>>(B, gray, cmap, queue, 80, 200, pm_cap, wp_cap);
    const int strip_cap = (DT_STRIP * max_roi_w + 3) & ~3;
    const size_t smem_d = 4 * (size_t)(2 * strip_cap + pm_cap);
    if (smem_d > 200 * 1024) return cudaErrorInvalidValue;
    e = cudaFuncSetAttribute(k_dist3x3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d);
    if (e != cudaSuccess) return e;
    k_dist3x3<<<B.n_tasks, DT_THREADS, smem_d, st>>>(B, cmap, reinterpret_cast<int*>(dtmp), maps, pm_cap, strip_cap);
    return cudaGetLastError();
}

}  // namespace csb
