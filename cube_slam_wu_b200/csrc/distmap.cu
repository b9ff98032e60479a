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

#include <cstdlib>
#include <cstring>
#include "context.h"
#include "ptx_helpers.cuh"

namespace csb {

constexpr int DM_THREADS_BATCH = 256;   // CTA of a batch launch: 4 ROIs per SM
constexpr int DM_THREADS_FEW = 1024;    // few ROIs (a single frame): the whole SM works on one ROI, for latency
constexpr int DM_COLS = 120;  // ROI columns a warp produces per row: lanes 1..30 x 4 pixels (lanes 0 and 31 only supply the magnitude halo)

constexpr unsigned DT_HV = 62587u;                 // cvRound(0.955f  * 65536)
constexpr unsigned DT_DG = 89738u;                 // cvRound(1.3693f * 65536)
constexpr unsigned DT_MAX = 0xffffffffu - DT_DG;   // DIST_MAX; also used for the border cells (behaves like OpenCV's INIT_DIST0)
constexpr int DT_INF = 0x3fffffff;     // "no path yet" inside the kernel (32-bit arithmetic); becomes DT_MAX on output

// gray bytes (y, x .. x+3) of a frame as one word, BORDER_REPLICATE at the image border, assembled byte by byte: only for words that
// straddle the left / right image border (everything else goes through two aligned word loads + a funnel shift, see k_distmap)
__device__ __noinline__ unsigned load_gray4_border(const uint8_t* __restrict__ row, int IW, int x) {
    unsigned w = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) w |= (unsigned)row[min(max(x + k, 0), IW - 1)] << (8 * k);
    return w;
}

// 16 bits -> the even bit positions of a word
__device__ __forceinline__ unsigned spread16(unsigned x) {
    x &= 0xffffu;
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// The two vertical sweeps of the distance transform (see k_distmap), thread = column x = tid + k * DM_THREADS.
template <int CPT, int DM_THREADS>
__device__ __forceinline__ void dt_sweeps(int W, int H, int WPE, const unsigned* Sb, int* rowD, int* rowU, int WB, int* __restrict__ tmpD, int* __restrict__ tmpU, int tid) {
    const int HV = (int)DT_HV, DG = (int)DT_DG;
    int dprev[CPT], uprev[CPT];
#pragma unroll
    for (int k = 0; k < CPT; k++) { dprev[k] = DT_INF; uprev[k] = DT_INF; }
    const unsigned* ed = Sb + (tid >> 5);                  // word of column tid in row 0; column tid + DM_THREADS k is DM_THREADS / 32 k words further
    const unsigned* eu = Sb + (H - 1) * WPE + (tid >> 5);
    const int sh = tid & 31;
    int* gd = tmpD + tid;
    int* gu = tmpU + (size_t)(H - 1) * W + tid;
    int off_c = 0, off_p = WB;  // row buffers: current / previous parity
    for (int i = 0; i < H; i++) {
#pragma unroll
        for (int k = 0; k < CPT; k++) {
            const int x = tid + k * DM_THREADS;
            if (x < W) {
                // cell x of the previous row sits at index x + 1 of the row buffer (indices 0 and W + 1 = border, always DT_INF)
                int d = min(dprev[k] + HV, min(rowD[off_p + x], rowD[off_p + x + 2]) + DG);
                int u = min(uprev[k] + HV, min(rowU[off_p + x], rowU[off_p + x + 2]) + DG);
                d = min(d, DT_INF); u = min(u, DT_INF);
                if ((ed[(DM_THREADS / 32) * k] >> sh) & 1u) d = 0;
                if ((eu[(DM_THREADS / 32) * k] >> sh) & 1u) u = 0;
                dprev[k] = d; uprev[k] = u;
                rowD[off_c + x + 1] = d; rowU[off_c + x + 1] = u;
                gd[k * DM_THREADS] = d;
                gu[k * DM_THREADS] = u;
            }
        }
        ed += WPE; eu -= WPE; gd += W; gu -= W;
        off_c ^= WB; off_p ^= WB;
        __syncthreads();
    }
}

// The same two sweeps when the ROI is at most half a CTA wide: warps 0 .. nw - 1 run the downward sweep, warps nw .. 2 nw - 1 the upward
// one, each group on its own named barrier (two shorter dependency chains side by side; the other warps wait at the caller's barrier).
__device__ __forceinline__ void dt_sweeps_split(int W, int H, int WPE, const unsigned* Sb, int* rowD, int* rowU, int WB, int* __restrict__ tmpD, int* __restrict__ tmpU, int tid) {
    const int HV = (int)DT_HV, DG = (int)DT_DG;
    const int nw = (W + 31) >> 5, grp = (tid >> 5) / nw;
    if (grp >= 2) return;
    const int x = tid - grp * nw * 32;
    const bool up = grp == 1;
    const int step = up ? -1 : 1, y0 = up ? H - 1 : 0;
    const unsigned* e = Sb + y0 * WPE + (x >> 5);
    const int sh = x & 31;
    int* row = up ? rowU : rowD;
    int* g = (up ? tmpU : tmpD) + (size_t)y0 * W + x;
    int prev = DT_INF, off_c = 0, off_p = WB;
    for (int i = 0; i < H; i++) {
        if (x < W) {
            int d = min(prev + HV, min(row[off_p + x], row[off_p + x + 2]) + DG);
            d = min(d, DT_INF);
            if ((*e >> sh) & 1u) d = 0;
            prev = d;
            row[off_c + x + 1] = d;
            *g = d;
        }
        e += step * WPE; g += step * W;
        off_c ^= WB; off_p ^= WB;
        if (up) asm volatile("bar.sync 2, %0;" ::"r"(nw * 32) : "memory");
        else asm volatile("bar.sync 1, %0;" ::"r"(nw * 32) : "memory");
    }
}

// Row pass of the distance transform, one warp per row: buf[x] = V(x) - a x on entry, dist(x) = min_x' (V(x') + a |x - x'|) on exit.
// Lane l owns the chunk [l CH, (l + 1) CH): prefix-min of V - a x and suffix-min of V + a x inside the chunk, the lanes' totals
// combined with shuffles.  Chunk in registers (CH <= CHT):
template <int CHT>
__device__ __forceinline__ void row_scan_regs(int* buf, int CH, int lane) {
    const unsigned FULL = 0xffffffffu;
    const int HV = (int)DT_HV;
    const int i0 = lane * CH;
    int p[CHT];
    int incP = 0x7fffffff, incQ = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < CHT; k++) {
        p[k] = (k < CH) ? buf[i0 + k] : DT_INF;   // a missing element can never win a minimum against a real one
        incP = min(incP, p[k]);
        incQ = min(incQ, p[k] + 2 * HV * (i0 + k));
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(FULL, incP, o), b = __shfl_down_sync(FULL, incQ, o);
        if (lane >= o) incP = min(incP, a);
        if (lane + o < 32) incQ = min(incQ, b);
    }
    int runP = __shfl_up_sync(FULL, incP, 1), runQ = __shfl_down_sync(FULL, incQ, 1);
    if (lane == 0) runP = 0x7fffffff;
    if (lane == 31) runQ = 0x7fffffff;
    int f[CHT];
#pragma unroll
    for (int k = 0; k < CHT; k++) { runP = min(runP, p[k]); f[k] = runP + HV * (i0 + k); }
#pragma unroll
    for (int k = CHT - 1; k >= 0; k--) {
        runQ = min(runQ, p[k] + 2 * HV * (i0 + k));
        if (k < CH) buf[i0 + k] = min(f[k], runQ - HV * (i0 + k));
    }
}
// buf[x] = min(Down, Up)(x) - a x for x < 32 CH (DT_INF beyond the row), all global loads of the lane in flight together
template <int CHT>
__device__ __forceinline__ void row_load(int* buf, const int* rd, const int* ru, int W, int CH, int lane) {
    const int HV = (int)DT_HV;
    int a[CHT], b[CHT];
#pragma unroll
    for (int j = 0; j < CHT; j++) {
        const int x = lane + 32 * j;
        a[j] = (x < W) ? rd[x] : DT_INF;
        b[j] = (x < W) ? ru[x] : DT_INF;
    }
#pragma unroll
    for (int j = 0; j < CHT; j++) {
        const int x = lane + 32 * j;
        if (j < CH) buf[x] = min(a[j], b[j]) - HV * x;
    }
    __syncwarp();
}
// any chunk length, through a second shared-memory buffer
__device__ __forceinline__ void row_scan_smem(int* bufP, int* bufQ, int CH, int lane) {
    const unsigned FULL = 0xffffffffu;
    const int HV = (int)DT_HV;
    const int i0 = lane * CH;
    int runP = 0x7fffffff, runQ = 0x7fffffff;
    for (int k = CH - 1; k >= 0; k--) { runQ = min(runQ, bufP[i0 + k] + 2 * HV * (i0 + k)); bufQ[i0 + k] = runQ; }
    for (int k = 0; k < CH; k++) { runP = min(runP, bufP[i0 + k]); bufP[i0 + k] = runP; }
    int incP = runP, incQ = runQ;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(FULL, incP, o), b = __shfl_down_sync(FULL, incQ, o);
        if (lane >= o) incP = min(incP, a);
        if (lane + o < 32) incQ = min(incQ, b);
    }
    int cP = __shfl_up_sync(FULL, incP, 1), cQ = __shfl_down_sync(FULL, incQ, 1);
    if (lane == 0) cP = 0x7fffffff;
    if (lane == 31) cQ = 0x7fffffff;
    for (int k = 0; k < CH; k++) {
        const int x = i0 + k;
        bufP[x] = min(min(cP, bufP[x]) + HV * x, min(cQ, bufQ[x]) - HV * x);
    }
}

// Per-task stage cycles of k_distmap (thread 0), compiled in only with -DCSB_DM_PHASES (CSB_DM_PHASES=1 python -m cube_slam_wu_b200.build --force;
// tools/distmap_phases.py reads them through csb_debug_distmap_phases).
#ifdef CSB_DM_PHASES
constexpr int DM_PHASE_TASKS = 8192;
__device__ long long g_dm_phase[DM_PHASE_TASKS][8];
#define DM_PHASE(idx) do { if (tid == 0 && task < DM_PHASE_TASKS) { const long long now__ = clock64(); g_dm_phase[task][idx] = now__ - t_prev__; t_prev__ = now__; } } while (0)
#else
#define DM_PHASE(idx) do { } while (0)
#endif

// One CTA per task: Canny (Sobel -> L1 magnitude -> non-maximum suppression -> hysteresis) and the 3x3 chamfer distance transform.
//
// (1) Sobel + NMS, register sliding window, no shared-memory staging: a warp takes a unit = (120 ROI columns, strip of rows); lane l owns the
//     four columns 120 wc + 4 (l - 1) .. + 3 and walks down the strip holding two gray rows (6 columns each, packed u16 pairs; the two
//     outer columns come from the neighbour lanes' words by shuffle), three magnitude rows (4 + the neighbours' adjacent magnitudes) and the
//     gradients of the middle row.  Column sums / differences are formed on packed pairs with plain integer adds (biased so that no borrow
//     crosses the halves); the suppression test is branch-free (the candidate density of real frames makes every branch divergent).
//     Result: one bit per pixel in two planes that stay in shared memory, S = edge (kept, magnitude > high) and Wk = weak candidate
//     (kept, low < magnitude <= high); two lanes combine their nibbles into one byte per plane.
// (2) hysteresis = flood fill on the bit planes: S |= dilate3x3(S) & Wk, word-parallel and in place, until nothing changes.
// (3) distance transform.  OpenCV's two raster passes compute, for every pixel, the exact 3x3 chamfer distance (a, b) to the nearest edge
//     pixel of the ROI (integer min / + only, so the order of evaluation cannot change a bit).  A shortest chamfer path can always be
//     taken as vertical / diagonal moves that all go down or all go up, followed by horizontal moves inside the target pixel's row.  So:
//     Down(y, x) = edge ? 0 : min(Down(y-1, x) + a, Down(y-1, x +- 1) + b) for rows top to bottom and Up(y, x) the same bottom to top --
//     two independent sweeps run in the same loop, thread = column, ONE barrier per row, no scan inside a row -- then per row
//     dist(x) = min_x' (V(x') + a |x - x'|) with V = min(Down, Up): a prefix-min of V - a x and a suffix-min of V + a x, one warp per
//     row (every lane scans a contiguous chunk in shared memory, the lanes' totals are combined with shuffles).  Checked against cv2 on
//     the CPU (numpy model) before it was written, and bit for bit on the GPU by tests/test_distmap_gpu.py.
template <int DM_THREADS>
__global__ void __launch_bounds__(DM_THREADS, 1024 / DM_THREADS) k_distmap(DetectBuffers B, const uint8_t* __restrict__ gray, uint8_t* cmap, int* queue, int* dtmp, float* maps,
                                                           int low, int high, int plane_words_cap, int work_ints) {
    constexpr int DM_WARPS = DM_THREADS / 32;
    extern __shared__ __align__(16) unsigned char dm_smem[];
    unsigned* Sb = reinterpret_cast<unsigned*>(dm_smem);                 // edge bits, H * WPE words; the weak plane follows it
    int* work = reinterpret_cast<int*>(Sb + plane_words_cap);            // sweeps: 2 x 2 row buffers; row pass: 2 buffers per warp

    const int task = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    const TaskTab t = B.ttab[task];
    const FrameTab& ft = B.ftab[t.frame_id];
    const uint8_t* img = gray + ft.gray_offset;
    const int img_mis = (int)(reinterpret_cast<uintptr_t>(img) & 3);
    const unsigned* imgA = reinterpret_cast<const unsigned*>(img - img_mis);
    const int IW = ft.img_w, IH = ft.img_h;
    const int W = t.roi_w, H = t.roi_h;
    const int WPE = (W + 31) >> 5, nE = H * WPE, RB = 4 * WPE;  // RB = bytes per plane row
    unsigned* Wb = Sb + nE;
    uint8_t* S8 = reinterpret_cast<uint8_t*>(Sb);
    uint8_t* W8 = reinterpret_cast<uint8_t*>(Wb);
    const int TG22 = (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5);
#ifdef CSB_DM_PHASES
    long long t_prev__ = clock64();
    int n_levels__ = 0;
#endif

    // ---- (1) Sobel, magnitude, non-maximum suppression
    {
        const int n_wc = (W + DM_COLS - 1) / DM_COLS;
        // plane bytes past the last warp column (columns >= W, padding of the 32-pixel words): no edge, no candidate
        {
            const int b0 = n_wc * (DM_COLS / 8), nb = RB - b0;
            for (int i = tid; i < H * nb; i += DM_THREADS) {
                const int y = i / nb, b = b0 + (i - y * nb);
                S8[(size_t)y * RB + b] = 0; W8[(size_t)y * RB + b] = 0;
            }
        }
        // number of row strips: the one that wastes least -- idle warps in the last round of units against the two extra rows a strip
        // starts with (score = units / (8 ceil(units / 8)) x SR / (SR + 2), compared as integers)
        int n_st;
        {
            // (every warp finds it for itself: candidate c = lane + 1 + 32 i, best of the warp by shuffles, the smallest c on ties)
            unsigned best = 0;  // score << 8 | (255 - c)
            for (int c = lane + 1; c <= 2 * DM_WARPS && c <= H; c += 32) {
                const unsigned sr = (H + c - 1) / c, u = n_wc * ((H + sr - 1) / sr);
                const unsigned score = u * sr * 1024u / (((u + DM_WARPS - 1) / DM_WARPS) * DM_WARPS * (sr + 2));  // <= 1024
                best = max(best, score << 8 | (unsigned)(255 - c));
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
            n_st = 255 - (int)(best & 255u);
        }
        const int SR = (H + n_st - 1) / n_st;
        const int n_units = n_wc * n_st;
        for (int u = warp; u < n_units; u += DM_WARPS) {
            const int wc = u % n_wc, s = u / n_wc;
            const int r_begin = s * SR, r_end = min(H, r_begin + SR);
            if (r_begin >= r_end) continue;  // warp-uniform
            const int cR = wc * DM_COLS + 4 * (lane - 1);  // first ROI column of this lane
            const bool need = (cR + 3 >= -1) && (cR <= W);   // columns -1 .. W feed a result
            bool cin[4];
#pragma unroll
            for (int k = 0; k < 4; k++) cin[k] = (cR + k >= 0) && (cR + k < W);
            const int bidx = wc * (DM_COLS / 8) + ((lane - 1) >> 1);  // byte of the plane row that this lane pair fills
            const bool store_ok = (lane & 1) && lane <= 29 && bidx < RB;
            const int gx = t.roi_left + cR;
            const bool fast = gx >= 0 && gx + 3 < IW;  // the lane's four bytes lie inside the image row: two aligned word loads + funnel shift
            const int goff = gx + img_mis;
            // the lane's gray word of ROI row rr (issued one row ahead of its use: a row step is one global-load latency otherwise)
            auto gray_word = [&](int rr) -> unsigned {
                rr = min(max(rr, -1), H);  // rows outside -1 .. H never reach a result
                const int yc = min(max(t.roi_top + rr, 0), IH - 1);
                unsigned w = 0u;
                if (need) {
                    if (fast) {
                        const int off = yc * IW + goff;  // byte offset from the 4-byte aligned base (rows are packed: any alignment)
                        const unsigned lo = __ldg(imgA + (off >> 2)), hi = __ldg(imgA + (off >> 2) + 1);  // (the frame buffer has 64 bytes of slack)
                        w = __funnelshift_r(lo, hi, (unsigned)(off & 3) * 8u);
                    } else
                        w = load_gray4_border(img + (size_t)yc * IW, IW, gx);
                }
                return w;
            };
            // word -> three packed pairs: columns (-1, 0), (1, 2), (3, 4) relative to cR
            auto unpack_row = [&](unsigned w, unsigned& P0, unsigned& P1, unsigned& P2) {
                const unsigned wl = __shfl_up_sync(FULL, w, 1), wr = __shfl_down_sync(FULL, w, 1);
                P0 = (wl >> 24) | ((w & 0xffu) << 16);
                P1 = __byte_perm(w, 0u, 0x4241);
                P2 = (w >> 24) | ((wr & 0xffu) << 16);
            };
            unsigned A0, A1, A2, B0, B1, B2;
            unpack_row(gray_word(r_begin - 2), A0, A1, A2);
            unpack_row(gray_word(r_begin - 1), B0, B1, B2);
            unsigned w_next = gray_word(r_begin);
            // magnitude rows m - 2 (M0) and m - 1 (M1): index 0 .. 5 = columns cR - 1 .. cR + 4; gradients of row m - 1
            int M0[6], M1[6], px[4], py[4];
#pragma unroll
            for (int j = 0; j < 6; j++) { M0[j] = 0; M1[j] = 0; }
#pragma unroll
            for (int k = 0; k < 4; k++) { px[k] = 0; py[k] = 0; }
            for (int m = r_begin - 1; m <= r_end; m++) {
                unsigned C0, C1, C2;
                const unsigned w_cur = w_next;
                w_next = gray_word(m + 2);
                unpack_row(w_cur, C0, C1, C2);
                // Sobel of row m: column sums S = top + 2 mid + bot, differences T = bot - top (+1024 per half)
                const unsigned S0 = A0 + C0 + 2u * B0, S1 = A1 + C1 + 2u * B1, S2 = A2 + C2 + 2u * B2;
                const unsigned T0 = C0 - A0 + 0x04000400u, T1 = C1 - A1 + 0x04000400u, T2 = C2 - A2 + 0x04000400u;
                const unsigned dx01 = S1 - S0 + 0x08000800u, dx23 = S2 - S1 + 0x08000800u;              // dx + 2048
                const unsigned T01 = __funnelshift_r(T0, T1, 16), T12 = __funnelshift_r(T1, T2, 16);
                const unsigned dy01 = T0 + 2u * T01 + T1, dy23 = T1 + 2u * T12 + T2;                    // dy + 4096
                int cx[4], cy[4], M2[6];
                cx[0] = (int)(dx01 & 0xffffu) - 2048; cx[1] = (int)(dx01 >> 16) - 2048; cx[2] = (int)(dx23 & 0xffffu) - 2048; cx[3] = (int)(dx23 >> 16) - 2048;
                cy[0] = (int)(dy01 & 0xffffu) - 4096; cy[1] = (int)(dy01 >> 16) - 4096; cy[2] = (int)(dy23 & 0xffffu) - 4096; cy[3] = (int)(dy23 >> 16) - 4096;
                const bool row_in = m >= 0 && m < H;
#pragma unroll
                for (int k = 0; k < 4; k++) M2[k + 1] = (row_in && cin[k]) ? abs(cx[k]) + abs(cy[k]) : 0;  // magnitude 0 outside the ROI
                M2[0] = __shfl_up_sync(FULL, M2[4], 1);
                M2[5] = __shfl_down_sync(FULL, M2[1], 1);
                const int r = m - 1;
                if (r >= r_begin) {
                    unsigned sbits = 0, wbits = 0;
                    const bool cand = (M1[1] > low) || (M1[2] > low) || (M1[3] > low) || (M1[4] > low);
                    if (__any_sync(FULL, cand)) {
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const int j = k + 1;
                            const int mg = M1[j];
                            const int xs = px[k], ys = py[k];
                            const int ax = abs(xs), ay = abs(ys) << 15;
                            const int tg22x = ax * TG22;
                            const int tg67x = tg22x + (ax << 16);
                            const bool horiz = ay < tg22x, vert = ay > tg67x;
                            const bool neg = (xs ^ ys) < 0;  // s = -1
                            int na = neg ? M0[j + 1] : M0[j - 1];
                            int nb = neg ? M2[j - 1] : M2[j + 1];
                            na = vert ? M0[j] : na; nb = vert ? M2[j] : nb;
                            na = horiz ? M1[j - 1] : na; nb = horiz ? M1[j + 1] : nb;
                            // horizontal / vertical: m > a && m >= b; diagonal: m > a && m > b
                            const bool diag = !horiz && !vert;
                            const bool keep = (mg > low) && (mg > na) && (mg + (diag ? 0 : 1) > nb);
                            const bool strong = keep && (mg > high);
                            sbits |= (strong ? 1u : 0u) << k;
                            wbits |= ((keep && !strong) ? 1u : 0u) << k;
                        }
                    }
                    const unsigned mine = sbits | (wbits << 4);
                    const unsigned other = __shfl_down_sync(FULL, mine, 1);
                    if (store_ok) {
                        S8[(size_t)r * RB + bidx] = (uint8_t)((mine & 15u) | ((other & 15u) << 4));
                        W8[(size_t)r * RB + bidx] = (uint8_t)((mine >> 4) | (other & 0xf0u));
                    }
                }
                A0 = B0; A1 = B1; A2 = B2; B0 = C0; B1 = C1; B2 = C2;
#pragma unroll
                for (int j = 0; j < 6; j++) { M0[j] = M1[j]; M1[j] = M2[j]; }
#pragma unroll
                for (int k = 0; k < 4; k++) { px[k] = cx[k]; py[k] = cy[k]; }
            }
        }
    }
    __syncthreads();
    DM_PHASE(0);
    // ---- (2) hysteresis: weak pixels 8-connected (through weak pixels) to an edge pixel become edge pixels.  In place: a word only ever
    //      gains bits, a stale neighbour word just postpones a promotion to the next round.
    while (true) {
        int changed = 0;
        for (int idx = tid; idx < nE; idx += DM_THREADS) {
            const unsigned wk = Wb[idx];
            if (wk == 0u) continue;
            const int y = idx / WPE, w = idx - y * WPE;
            unsigned acc = 0;
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
                const int yy = y + dy;
                if (yy < 0 || yy >= H) continue;
                const unsigned* row = Sb + yy * WPE;
                const unsigned c = row[w], l = (w > 0) ? row[w - 1] : 0u, rr = (w + 1 < WPE) ? row[w + 1] : 0u;
                acc |= c | (c << 1) | (c >> 1) | (l >> 31) | (rr << 31);
            }
            const unsigned nw = acc & wk;
            if (nw) { Sb[idx] |= nw; Wb[idx] = wk & ~nw; changed = 1; }
        }
#ifdef CSB_DM_PHASES
        n_levels__++;
#endif
        if (!__syncthreads_or(changed)) break;
    }
    DM_PHASE(1);
    // packed 2-bit map -> global (csb_detect_debug_map reads it): 0 weak candidate, 1 no edge, 2 edge; 16 pixels per word
    {
        const int WPR = (W + 15) >> 4;
        unsigned* out = reinterpret_cast<unsigned*>(cmap + 4 * (size_t)t.map_offset);
        for (int i = tid; i < H * WPR; i += DM_THREADS) {
            const int y = i / WPR, w = i - y * WPR;
            const unsigned sw = Sb[y * WPE + (w >> 1)] >> (16 * (w & 1)), ww = Wb[y * WPE + (w >> 1)] >> (16 * (w & 1));
            out[i] = spread16(~(sw | ww)) | (spread16(sw) << 1);
        }
    }
    // ---- (3) distance transform: vertical sweeps
    const int WB = W + 2;
    int* rowD = work;            // 2 x WB
    int* rowU = work + 2 * WB;   // 2 x WB
    for (int i = tid; i < 2 * WB; i += DM_THREADS) { rowD[i] = DT_INF; rowU[i] = DT_INF; }
    __syncthreads();
    DM_PHASE(2);
    int* tmpD = dtmp + t.map_offset;
    int* tmpU = queue + t.map_offset;
    if (2 * ((W + 31) & ~31) <= DM_THREADS) {
        dt_sweeps_split(W, H, WPE, Sb, rowD, rowU, WB, tmpD, tmpU, tid);
        __syncthreads();
    } else if (W <= DM_THREADS) dt_sweeps<1, DM_THREADS>(W, H, WPE, Sb, rowD, rowU, WB, tmpD, tmpU, tid);
    else if (W <= 2 * DM_THREADS) dt_sweeps<2, DM_THREADS>(W, H, WPE, Sb, rowD, rowU, WB, tmpD, tmpU, tid);
    else if (W <= 3 * DM_THREADS) dt_sweeps<3, DM_THREADS>(W, H, WPE, Sb, rowD, rowU, WB, tmpD, tmpU, tid);
    else dt_sweeps<5, DM_THREADS>(W, H, WPE, Sb, rowD, rowU, WB, tmpD, tmpU, tid);
    DM_PHASE(3);
    // ---- row pass (the sweeps end with a barrier: the block's global writes are visible to all its threads).
    // Scaled like OpenCV: (float)(unsigned) * 2^-16; unreachable -> DIST_MAX.
    {
        const int HV = (int)DT_HV;
        const int CH = ((W + 31) >> 5) | 1;   // odd chunk length: conflict-free shared-memory access at stride CH
        const int WP = 32 * CH;
        const int n_rw = min(DM_WARPS, work_ints / (2 * WP));  // warps that get a buffer pair (all of them unless the ROI is very wide)
        int* buf = work + warp * 2 * WP;
        float* out = maps + t.map_offset;
        const float scale = 1.f / 65536.f;
        for (int y = warp; y < H && warp < n_rw; y += n_rw) {
            const int* rd = tmpD + (size_t)y * W;
            const int* ru = tmpU + (size_t)y * W;
            if (CH <= 5) { row_load<5>(buf, rd, ru, W, CH, lane); row_scan_regs<5>(buf, CH, lane); }
            else if (CH <= 9) { row_load<9>(buf, rd, ru, W, CH, lane); row_scan_regs<9>(buf, CH, lane); }
            else if (CH <= 13) { row_load<13>(buf, rd, ru, W, CH, lane); row_scan_regs<13>(buf, CH, lane); }
            else {
                for (int x = lane; x < WP; x += 32) buf[x] = ((x < W) ? min(rd[x], ru[x]) : DT_INF) - HV * x;
                __syncwarp();
                row_scan_smem(buf, buf + WP, CH, lane);
            }
            __syncwarp();
            float* o = out + (size_t)y * W;
            for (int x = lane; x < W; x += 32) {
                const int v = buf[x];
                o[x] = (float)((v >= DT_INF) ? DT_MAX : (unsigned)v) * scale;
            }
            __syncwarp();
        }
    }
    DM_PHASE(4);
#ifdef CSB_DM_PHASES
    if (tid == 0 && task < DM_PHASE_TASKS) { g_dm_phase[task][5] = n_levels__; g_dm_phase[task][6] = 0; g_dm_phase[task][7] = (long long)W << 32 | H; }
#endif
}

// ---- gray-frame upload without the copy engine: only what the ROIs read -----------------------------------------------------------
// The packed gray frames of a batch are one linear byte array, cut into segments of CSB_GRAY_SEG bytes.  k_gray_mark sets the bit of every
// segment that holds a pixel of some task's ROI or of its one-pixel Sobel halo (clamped to the image); k_gray_gather copies the marked
// segments from the caller's pinned (device-mapped) buffer into the device frame buffer -- SM-initiated PCIe reads run at the copy engine's
// rate (measured: 50 GB/s against 55), so the upload shrinks with the share of the frames the boxes cover.  A ROI row is ~220 bytes at an
// arbitrary offset: with 128-byte segments it costs 2.7 segments (346 bytes), with 32-byte segments (one sector) 7.8 (250 bytes): 18 MB ->
// 13 MB per bench step, which is what counts when eight ranks pull from one host.
// Everything outside the marked segments is never read by k_distmap in a way that reaches a result (magnitudes outside the ROI count as 0).
// The gather is a persistent kernel of small CTAs (64 threads, <= 32 registers, no shared memory) so that it finds room next to the
// one-CTA-per-SM scoring kernel of another context and the transfer overlaps compute the way a copy-engine transfer would.
constexpr int GSEG = CSB_GRAY_SEG;
constexpr int GG_THREADS = 64;
constexpr int GG_LPS = GSEG / 16;       // lanes per segment (16 bytes each)
constexpr int GG_SPI = 32 / GG_LPS;     // segments per load instruction of a warp
static_assert(GSEG >= 16 && GSEG <= 128 && (GSEG & (GSEG - 1)) == 0, "CSB_GRAY_SEG");

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

// Each warp walks bitmap words (32 segments); a load instruction moves GG_SPI segments (lane / GG_LPS = segment, lane % GG_LPS = 16-byte
// chunk), two load instructions are in flight per lane before the stores.
__global__ void __launch_bounds__(GG_THREADS, 8) k_gray_gather(const uint4* __restrict__ src, uint4* __restrict__ dst, const unsigned* __restrict__ seg_bits, int n_words,
                                                             long long n_chunks16, int* n_segments_out) {
    const int lane = threadIdx.x & 31, sub = lane / GG_LPS, ch = lane % GG_LPS;
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
                // the (sub + 1)-th set bit of m, then drop up to GG_SPI bits
                const unsigned b = __fns(m, 0, sub + 1);
                o[k] = (b < 32u) ? ((long long)w * 32 + b) * (GSEG / 16) + ch : -1;
                if (o[k] >= n_chunks16) o[k] = -1;
#pragma unroll
                for (int q = 0; q < GG_SPI; q++) m &= m - 1;
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
    if (max_roi_w > DM_THREADS_BATCH * 5) return cudaErrorInvalidValue;  // ROI wider than 1280 px
    const int plane_cap = (max_pm_words + 3) & ~3;  // >= 2 * H * ceil(W / 32) words for every task (capi_detect.cu)
    const int wp = 32 * (((max_roi_w + 31) >> 5) | 1);
    // few ROIs (one frame per call): one 1024-thread CTA per ROI -- four times the warps on the suppression units and on the row pass
    bool few = B.n_tasks <= 148;  // at most one ROI per SM
    if (const char* force = getenv("CSB_DISTMAP_CTA")) {  // the tests run every case through both variants
        if (!strcmp(force, "1024")) few = true;
        else if (!strcmp(force, "256")) few = false;
    }
    const int warps = (few ? DM_THREADS_FEW : DM_THREADS_BATCH) / 32;
    // work space: row pass (2 buffers of wp ints per warp, as many warps as fit in ~96 KB, at least 8) | sweeps (4 * (W + 2)) | nothing else
    int rw = warps;
    while (rw > 8 && ((size_t)2 * rw * wp * 4 > 96 * 1024 || 4 * (size_t)plane_cap + (size_t)2 * rw * wp * 4 > 220 * 1024)) rw >>= 1;
    const int work_ints = 2 * rw * wp;
    const size_t smem = 4 * (size_t)plane_cap + 4 * (size_t)work_ints;
    if (smem > 220 * 1024) return cudaErrorInvalidValue;
    cudaError_t e;
    if (few) {
        e = cudaFuncSetAttribute(k_distmap<DM_THREADS_FEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_distmap<DM_THREADS_FEW><<<B.n_tasks, DM_THREADS_FEW, smem, st>>>(B, gray, cmap, queue, reinterpret_cast<int*>(dtmp), maps, 80, 200, plane_cap, work_ints);
    } else {
        e = cudaFuncSetAttribute(k_distmap<DM_THREADS_BATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_distmap<DM_THREADS_BATCH><<<B.n_tasks, DM_THREADS_BATCH, smem, st>>>(B, gray, cmap, queue, reinterpret_cast<int*>(dtmp), maps, 80, 200, plane_cap, work_ints);
    }
    return cudaGetLastError();
}

}  // namespace csb

#ifdef CSB_DM_PHASES
extern "C" int csb_debug_distmap_phases(long long* out, int n_tasks) {
    if (n_tasks > csb::DM_PHASE_TASKS) n_tasks = csb::DM_PHASE_TASKS;
    return (int)cudaMemcpyFromSymbol(out, csb::g_dm_phase, sizeof(long long) * 8 * (size_t)n_tasks);
}
#endif
