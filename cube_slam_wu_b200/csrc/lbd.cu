// lbd.cu -- LBD line descriptors on the GPU (SURVEY.md 8 "next" row f-2, the descriptor half of BASELINE config #3).
//
// Replaces, for a batch of equally sized 8-bit gray frames and their line segments, what
//     line_lbd_detect::detect_descrip_lines(gray_img, lines_mat, line_descrips)      line_lbd/class/line_lbd_allclass.cpp:239-281
// does after the detector: lbd->compute(gray_img, keylines, descrips) = BinaryDescriptor::computeImpl
// (line_lbd/libs/binary_descriptor.cpp:607-794, one octave):
//   computeGaussianPyramid / computeSobel (:347-402)  GaussianBlur 5x5 sigma 1 on the u8 frame, Sobel 3x3 -> two int16 images
//   computeLBD (:1150-1512)                            per line: 63 rows (9 bands x 7) x numOfPixels samples of the gradient, rotated into the
//                                                      line's frame, signed row sums, Gaussian band weights, mean / std per band -> 72 floats,
//                                                      two normalisations with the 0.4 clip
//   binaryConversion (:405-417, :766-773)              32 bytes from the 32 band pairs of `combinations` (:74-107)
// with the key-line fields as LSDDetector::detectImpl fills them (line_lbd/libs/LSDDetector.cpp:80-101 clamp, :239-245 numOfPixels, angle).
//
//   k_lbd_grad     : (64x16 tile, frame).  u8 tile + 3-pixel halo (BORDER_REFLECT_101) -> shared memory; the 5-tap blur as OpenCV's CV_8U
//                    fixed-point path does it (taps {14, 62, 104, 62, 14} / 256: exact 16-bit horizontal sums, exact 32-bit vertical sums, one
//                    rounding); Sobel on the blurred tile; one 4-byte {dx, dy} record per pixel (the descriptor reads both with one gather).
//                    HBM bound: 1 B read + 4 B written per pixel.
//   k_lbd_prefix   : exclusive prefix sum of the per-frame line counts (work list of the next kernel; the counts may come straight from the
//                    line detector's device buffers).
//   k_lbd_describe : persistent (8 CTAs x 4 warps per SM), one warp per line (global atomic work counter).  Lane l owns rows l and l + 32 of the line's support
//                    region and walks them sample by sample: the reference accumulates the sample position (sCorX += dL[0]) and the four row
//                    sums in FLOAT, sample after sample, so a row is a sequential chain -- but the 63 rows are independent, and the
//                    position chain does not depend on the gathered values, so the gathers of one row are issued several samples ahead.
//                    The band sums add the 21 (14) rows of a band in row order from shared memory (lane = (band, quantity)); the two
//                    normalisations are sequential 36- / 72-term float sums, as in the reference.  Gather bound (L1/L2: neighbouring rows
//                    and consecutive samples share 32-byte sectors).
//
// Arithmetic: float, the reference's operation order, no FMA contraction (-fmad=false, IEEE sqrt/div).  atan2 / cos / sin of the reference
// (libm on float operands, float results) are det_atan2 / det_sincos in double rounded once to float -- the specified variants the oracle
// uses too (oracle/oracle_lbd.cpp); the band weights are computed on the host with the reference's expressions.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "context.h"
#include "csb_math.cuh"
#include "lbd.h"

namespace csb {

constexpr int LBD_BANDS = 9, LBD_BAND_W = 7, LBD_ROWS = LBD_BANDS * LBD_BAND_W;
constexpr int LG_TW = 64, LG_TH = 16, LG_THREADS = 256;
constexpr int LD_WARPS = 4;
constexpr int LD_CTAS_PER_SM = 8;  // 32 warps per SM; 16 CTAs per SM measured slower (0.27 vs 0.26 ms for 22.8 k lines)

__constant__ float c_lbd_G[LBD_ROWS];
__constant__ float c_lbd_L[3 * LBD_BAND_W];
__constant__ unsigned char c_lbd_comb[32][2] = {{0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {0, 6}, {1, 2}, {1, 3}, {1, 4}, {1, 5}, {1, 6},
                                                {2, 3}, {2, 4}, {2, 5}, {2, 6}, {2, 7}, {2, 8}, {3, 4}, {3, 5}, {3, 6}, {3, 7}, {3, 8},
                                                {4, 5}, {4, 6}, {4, 7}, {4, 8}, {5, 6}, {5, 7}, {5, 8}, {6, 7}, {6, 8}, {7, 8}};

__device__ __forceinline__ int lbd_reflect101(int p, int len) {
    if (p < 0) p = -p;
    if (p >= len) p = 2 * len - 2 - p;
    return min(max(p, 0), len - 1);
}

// same sequence as oracle_lsd.cpp / lsd.cu det_sincos
__device__ __forceinline__ void lbd_sincos(double x, double& s, double& c) {
    const double two_over_pi = 6.36619772367581382433e-01, pio2_hi = 1.57079632673412561417e+00, pio2_lo = 6.07710050650619224932e-11;
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double kf = floor(x * two_over_pi + 0.5);
    const int k = (int)kf;
    double r = x - kf * pio2_hi;
    r = r - kf * pio2_lo;
    const double z = r * r;
    const double sn = r + (z * r) * (S1 + z * (S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)))));
    const double cs = 1.0 - (0.5 * z - z * (z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))))));
    switch (k & 3) {
        case 0: s = sn; c = cs; break;
        case 1: s = cs; c = -sn; break;
        case 2: s = -sn; c = -cs; break;
        default: s = -cs; c = sn; break;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// blur 5x5 (fixed point) + Sobel -> {dx, dy} int16 pairs
// ---------------------------------------------------------------------------------------------------------------------------------
// K0, K1, K2: the integer taps {K0, K1, K2, K1, K0} / 256 of the 8-bit 5x5 Gaussian; SAT: saturate the blurred value to 255 (needed when the taps
// sum to more than 256).  <14, 62, 104, false> = cv2 4.x's fixed-point path; <14, 63, 103, true> = OpenCV <= 3.4.0 (sepFilter2D with 8 fractional
// bits, every tap rounded on its own), the generation the reference's committed outputs were produced with (oracle/oracle_lbd.cpp).
template <int K0, int K1, int K2, bool SAT>
__global__ void __launch_bounds__(LG_THREADS) k_lbd_grad(const uint8_t* __restrict__ gray, short2* __restrict__ grad, int w, int h) {
    __shared__ uint8_t s_g[LG_TH + 6][LG_TW + 8];          // gray rows y0-3 .. y0+TH+2, columns x0-3 .. x0+TW+2
    __shared__ unsigned short s_h[LG_TH + 6][LG_TW + 2];   // horizontal sums, columns x0-1 .. x0+TW
    __shared__ uint8_t s_b[LG_TH + 2][LG_TW + 4];          // blurred rows y0-1 .. y0+TH, columns x0-1 .. x0+TW
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * LG_TW, y0 = blockIdx.y * LG_TH;
    const size_t fo = (size_t)blockIdx.z * w * h;
    const uint8_t* img = gray + fo;
    for (int i = tid; i < (LG_TH + 6) * (LG_TW + 6); i += LG_THREADS) {
        const int r = i / (LG_TW + 6), c = i - r * (LG_TW + 6);
        s_g[r][c] = img[(size_t)lbd_reflect101(y0 - 3 + r, h) * w + lbd_reflect101(x0 - 3 + c, w)];
    }
    __syncthreads();
    for (int i = tid; i < (LG_TH + 6) * (LG_TW + 2); i += LG_THREADS) {
        const int r = i / (LG_TW + 2), c = i - r * (LG_TW + 2);
        const uint8_t* p = &s_g[r][c];
        s_h[r][c] = (unsigned short)(K0 * (p[0] + p[4]) + K1 * (p[1] + p[3]) + K2 * p[2]);   // <= 257 * 255 = 65535
    }
    __syncthreads();
    for (int i = tid; i < (LG_TH + 2) * (LG_TW + 2); i += LG_THREADS) {
        const int r = i / (LG_TW + 2), c = i - r * (LG_TW + 2);
        const unsigned v = (unsigned)K0 * (s_h[r][c] + s_h[r + 4][c]) + (unsigned)K1 * (s_h[r + 1][c] + s_h[r + 3][c]) + (unsigned)K2 * s_h[r + 2][c];
        const unsigned bv = (v + 32768u) >> 16;
        s_b[r][c] = (uint8_t)(SAT ? min(bv, 255u) : bv);
    }
    __syncthreads();
    for (int i = tid; i < LG_TH * LG_TW; i += LG_THREADS) {
        const int r = i / LG_TW, c = i - r * LG_TW;
        const int x = x0 + c, y = y0 + r;
        if (x >= w || y >= h) continue;
        const int a = s_b[r][c], b = s_b[r][c + 1], cc = s_b[r][c + 2];
        const int d = s_b[r + 1][c], f = s_b[r + 1][c + 2];
        const int q = s_b[r + 2][c], hh = s_b[r + 2][c + 1], ii = s_b[r + 2][c + 2];
        grad[fo + (size_t)y * w + x] = make_short2((short)((cc + 2 * f + ii) - (a + 2 * d + q)), (short)((q + 2 * hh + ii) - (a + 2 * b + cc)));
    }
}

// The same for frames whose width is a multiple of 4 (every row 4-byte aligned): 128x32 tile, four pixels per thread in every stage,
// 32-bit loads of the gray rows, packed 4 x u16 / 4 x u8 intermediates in shared memory, one 16-byte store of four {dx, dy} records.
// Column groups: group g of a tile covers columns x0 - 4 + 4g .. x0 - 1 + 4g (g = 0 .. 33; 1 .. 32 are the tile's own pixels).
constexpr int L4_TW = 128, L4_TH = 32, L4_G = L4_TW / 4 + 2, L4_THREADS = 256;

__global__ void __launch_bounds__(L4_THREADS) k_lbd_grad4(const uint8_t* __restrict__ gray, short2* __restrict__ grad, int w, int h) {
    __shared__ uint32_t s_g[L4_TH + 6][L4_G];   // gray rows y0-3 .. y0+TH+2
    __shared__ uint2 s_h[L4_TH + 6][L4_G];      // horizontal sums (4 x u16)
    __shared__ uint32_t s_b[L4_TH + 2][L4_G];   // blurred rows y0-1 .. y0+TH (4 x u8)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = blockIdx.x * L4_TW, y0 = blockIdx.y * L4_TH;
    const size_t fo = (size_t)blockIdx.z * w * h;
    const uint8_t* img = gray + fo;
    for (int r = warp; r < L4_TH + 6; r += 8) {
        const uint8_t* row = img + (size_t)lbd_reflect101(y0 - 3 + r, h) * w;
        for (int g = lane; g < L4_G; g += 32) {
            const int xw = x0 - 4 + 4 * g;
            uint32_t v;
            if (xw >= 0 && xw + 3 < w) v = *reinterpret_cast<const uint32_t*>(row + xw);
            else {
                v = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) v |= (uint32_t)row[lbd_reflect101(xw + k, w)] << (8 * k);
            }
            s_g[r][g] = v;
        }
    }
    __syncthreads();
    for (int r = warp; r < L4_TH + 6; r += 8)
        for (int g = lane; g < L4_G; g += 32) {
            const uint32_t wl = g > 0 ? s_g[r][g - 1] : 0u, wc = s_g[r][g], wr = g < L4_G - 1 ? s_g[r][g + 1] : 0u;
            int b[8];  // columns 4g-2 .. 4g+5 of the group's frame
            b[0] = (wl >> 16) & 255; b[1] = wl >> 24;
            b[2] = wc & 255; b[3] = (wc >> 8) & 255; b[4] = (wc >> 16) & 255; b[5] = wc >> 24;
            b[6] = wr & 255; b[7] = (wr >> 8) & 255;
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) o[k] = 14 * (b[k] + b[k + 4]) + 62 * (b[k + 1] + b[k + 3]) + 104 * b[k + 2];
            s_h[r][g] = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
        }
    __syncthreads();
    for (int r = warp; r < L4_TH + 2; r += 8)
        for (int g = lane; g < L4_G; g += 32) {
            const uint2 h0 = s_h[r][g], h1 = s_h[r + 1][g], h2 = s_h[r + 2][g], h3 = s_h[r + 3][g], h4 = s_h[r + 4][g];
            auto tap = [](uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, int sh) -> uint32_t {
                const uint32_t m = 0xffffu;
                return (14u * (((a0 >> sh) & m) + ((a4 >> sh) & m)) + 62u * (((a1 >> sh) & m) + ((a3 >> sh) & m)) + 104u * ((a2 >> sh) & m) + 32768u) >> 16;
            };
            s_b[r][g] = tap(h0.x, h1.x, h2.x, h3.x, h4.x, 0) | (tap(h0.x, h1.x, h2.x, h3.x, h4.x, 16) << 8) | (tap(h0.y, h1.y, h2.y, h3.y, h4.y, 0) << 16) |
                        (tap(h0.y, h1.y, h2.y, h3.y, h4.y, 16) << 24);
        }
    __syncthreads();
    for (int r = warp; r < L4_TH; r += 8) {
        const int y = y0 + r, g = lane + 1, x = x0 + 4 * lane;
        if (y >= h || x >= w) continue;
        int cs[6], dd[6];  // column sums top + 2 mid + bot and differences bot - top over columns x-1 .. x+4
#pragma unroll
        for (int j = 0; j < 6; j++) {
            const int gi = (j == 0) ? g - 1 : (j == 5) ? g + 1 : g, sh = (j == 0) ? 24 : (j == 5) ? 0 : 8 * (j - 1);
            const int t = (s_b[r][gi] >> sh) & 255, m = (s_b[r + 1][gi] >> sh) & 255, bt = (s_b[r + 2][gi] >> sh) & 255;
            cs[j] = t + 2 * m + bt;
            dd[j] = bt - t;
        }
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int dx = cs[k + 2] - cs[k], dy = dd[k] + 2 * dd[k + 1] + dd[k + 2];
            o[k] = ((uint32_t)dx & 0xffffu) | ((uint32_t)dy << 16);
        }
        *reinterpret_cast<uint4*>(grad + fo + (size_t)y * w + x) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// work list: exclusive prefix of min(count, stride) over frames; prefix[n_frames] = total.  Also resets the work counter.
// ---------------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_lbd_prefix(const int* __restrict__ counts, int n_frames, int stride, int* __restrict__ prefix, unsigned long long* ctr) {
    __shared__ int s_w[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_base = 0; ctr[0] = 0; ctr[1] = 0; }
    __syncthreads();
    for (int f0 = 0; f0 < n_frames; f0 += 1024) {
        const int f = f0 + tid;
        const int v = (f < n_frames) ? min(max(counts[f], 0), stride) : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int t = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += y; }
            s_w[lane] = t;
        }
        __syncthreads();
        const int base = s_base + (warp ? s_w[warp - 1] : 0);
        if (f < n_frames) prefix[f] = base + inc - v;
        __syncthreads();
        if (tid == 1023) s_base = base + inc;
        __syncthreads();
    }
    if (tid == 0) prefix[n_frames] = s_base;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------------------------------------------------------------
struct LbdArgs {
    const short2* grad;    // n_frames x h x w
    const float* lines;    // n_frames x stride x 4
    const int* prefix;     // n_frames + 1
    unsigned long long* ctr;  // [0] work counter, [1] samples
    uint8_t* desc;         // n_frames x stride x 32
    float* descf;          // n_frames x stride x 72 or nullptr
    float* keyl;           // n_frames x stride x 4 {angle, numOfPixels, lineLength, 0} or nullptr
    const float2* keyl_in; // nullptr: key-line fields as LSDDetector fills them; else n_frames x stride {lineDirection_, numOfPixels} from the detector (EDLines)
    int w, h, n_frames, stride;
};

__global__ void __launch_bounds__(32 * LD_WARPS) k_lbd_describe(LbdArgs A) {
    __shared__ float s_rs[LD_WARPS][4][64];  // row sums {pgdL, ngdL, pgdO, ngdO} x row, already weighted by gaussCoefG_
    __shared__ float s_des[LD_WARPS][72];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned FULL = 0xffffffffu;
    float(*rs)[64] = s_rs[warp];
    float* des = s_des[warp];
    const int total = A.prefix[A.n_frames];
    const int w = A.w, h = A.h;
    while (true) {
        int item = 0;
        if (lane == 0) item = (int)atomicAdd(A.ctr, 1ull);
        item = __shfl_sync(FULL, item, 0);
        if (item >= total) break;
        // frame of the item: last f with prefix[f] <= item
        int lo = 0, hi = A.n_frames - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (A.prefix[mid] <= item) lo = mid; else hi = mid - 1;
        }
        const int frame = lo, idx = item - A.prefix[lo];
        const size_t row = (size_t)frame * A.stride + idx;
        const float4 ln = *reinterpret_cast<const float4*>(A.lines + 4 * row);
        // key-line fields (LSDDetector.cpp:80-101, 239-245)
        float e0 = ln.x, e1 = ln.y, e2 = ln.z, e3 = ln.w;
        if (e0 < 0) e0 = 0;
        if (e0 >= w) e0 = (float)w - 1.0f;
        if (e2 < 0) e2 = 0;
        if (e2 >= w) e2 = (float)w - 1.0f;
        if (e1 < 0) e1 = 0;
        if (e1 >= h) e1 = (float)h - 1.0f;
        if (e3 < 0) e3 = 0;
        if (e3 >= h) e3 = (float)h - 1.0f;
        const int px1 = __float2int_rn(e0), py1 = __float2int_rn(e1), px2 = __float2int_rn(e2), py2 = __float2int_rn(e3);  // cvRound
        int len = max(abs(px2 - px1), abs(py2 - py1)) + 1;  // cv::LineIterator count, 8-connected
        const float fy = e3 - e1, fx = e2 - e0;
        float angle;
        if (A.keyl_in) {
            // detect_descrip_lines with use_LSD = false: the EDLines detector supplies direction = lineDirection_ and numOfPixels = pixels of
            // the fitted chain segment, and its end points are used as projected, unclamped (binary_descriptor.cpp:1045-1140)
            const float2 k = A.keyl_in[row];
            angle = k.x;
            len = (int)(short)(int)k.y;
            e0 = ln.x; e1 = ln.y; e2 = ln.z; e3 = ln.w;
        } else
            angle = (float)det_atan2((double)fy, (double)fx);
        float dL0, dL1;
        {
            double sn, cs;
            lbd_sincos((double)angle, sn, cs);
            dL0 = (float)cs;
            dL1 = (float)sn;
        }
        const float dO0 = -dL1, dO1 = dL0;
        const int halfWidth = (len - 1) / 2, halfHeight = (LBD_ROWS - 1) / 2;
        const float midX = (float)(0.5 * (double)(e0 + e2)), midY = (float)(0.5 * (double)(e1 + e3));
        // computeLBD :1262-1263; the row origin moves by (-dL1, +dL0) per row, accumulated in float like the reference (:1322-1323)
        float ox = (-dL0 * (float)halfWidth + dL1 * (float)halfHeight) + midX;
        float oy = (-dL1 * (float)halfWidth - dL0 * (float)halfHeight) + midY;
        float xA = 0.f, yA = 0.f, xB = 0.f, yB = 0.f;  // origins of rows lane and lane + 32 (lane 31 has no second row)
        for (int r = 0; r < LBD_ROWS; r++) {
            if (r == lane) { xA = ox; yA = oy; }
            if (r == lane + 32) { xB = ox; yB = oy; }
            ox = __fsub_rn(ox, dL1);
            oy = __fadd_rn(oy, dL0);
        }
        const bool hasB = lane + 32 < LBD_ROWS;
        const short2* g = A.grad + (size_t)frame * w * h;
        float pLA = 0.f, nLA = 0.f, pOA = 0.f, nOA = 0.f, pLB = 0.f, nLB = 0.f, pOB = 0.f, nOB = 0.f;
        const int wm = w - 1, hm = h - 1;
#pragma unroll 4
        for (int s = 0; s < len; s++) {
            const int xa = min(max((int)(short)(int)roundf(xA), 0), wm), ya = min(max((int)(short)(int)roundf(yA), 0), hm);
            const int xb = min(max((int)(short)(int)roundf(xB), 0), wm), yb = min(max((int)(short)(int)roundf(yB), 0), hm);
            const short2 ga = __ldg(g + ya * w + xa);
            const short2 gb = __ldg(g + yb * w + xb);
            {
                const float dx = (float)ga.x, dy = (float)ga.y;
                const float gDL = __fadd_rn(__fmul_rn(dx, dL0), __fmul_rn(dy, dL1));
                const float gDO = __fadd_rn(__fmul_rn(dx, dO0), __fmul_rn(dy, dO1));
                // `if (g > 0) p += g; else n -= g;` without the branch: the untouched sum gets + 0
                pLA = __fadd_rn(pLA, fmaxf(gDL, 0.f));
                nLA = __fadd_rn(nLA, fmaxf(-gDL, 0.f));
                pOA = __fadd_rn(pOA, fmaxf(gDO, 0.f));
                nOA = __fadd_rn(nOA, fmaxf(-gDO, 0.f));
            }
            {
                const float dx = (float)gb.x, dy = (float)gb.y;
                const float gDL = __fadd_rn(__fmul_rn(dx, dL0), __fmul_rn(dy, dL1));
                const float gDO = __fadd_rn(__fmul_rn(dx, dO0), __fmul_rn(dy, dO1));
                pLB = __fadd_rn(pLB, fmaxf(gDL, 0.f));
                nLB = __fadd_rn(nLB, fmaxf(-gDL, 0.f));
                pOB = __fadd_rn(pOB, fmaxf(gDO, 0.f));
                nOB = __fadd_rn(nOB, fmaxf(-gDO, 0.f));
            }
            xA = __fadd_rn(xA, dL0); yA = __fadd_rn(yA, dL1);
            xB = __fadd_rn(xB, dL0); yB = __fadd_rn(yB, dL1);
        }
        {
            const float cg = c_lbd_G[lane];
            rs[0][lane] = cg * pLA; rs[1][lane] = cg * nLA; rs[2][lane] = cg * pOA; rs[3][lane] = cg * nOA;
            if (hasB) {
                const float cb = c_lbd_G[lane + 32];
                rs[0][lane + 32] = cb * pLB; rs[1][lane + 32] = cb * nLB; rs[2][lane + 32] = cb * pOB; rs[3][lane + 32] = cb * nOB;
            }
        }
        __syncwarp();
        // band sums (:1339-1379): accumulator (band b, quantity q) receives the rows of bands b-1, b, b+1 in row order
        for (int a = lane; a < LBD_BANDS * 4; a += 32) {
            const int b = a >> 2, q = a & 3;
            const int r_lo = max(0, (b - 1) * LBD_BAND_W), r_hi = min(LBD_ROWS, (b + 2) * LBD_BAND_W);
            float S = 0.f, S2 = 0.f;
            for (int r = r_lo; r < r_hi; r++) {
                const int rb = r / LBD_BAND_W, j = r - rb * LBD_BAND_W;
                const float c = (rb == b) ? c_lbd_L[j + LBD_BAND_W] : (rb == b + 1) ? c_lbd_L[j + 2 * LBD_BAND_W] : c_lbd_L[j];
                const float v = rs[q][r];
                S = __fadd_rn(S, __fmul_rn(c, v));
                S2 = __fadd_rn(S2, __fmul_rn(__fmul_rn(c, c), __fmul_rn(v, v)));
            }
            const float invN = (b == 0 || b == LBD_BANDS - 1) ? (float)(1.0 / (LBD_BAND_W * 2.0)) : (float)(1.0 / (LBD_BAND_W * 3.0));
            const float mean = __fmul_rn(S, invN);
            // desVec[b*8 + {0,1,2,3}] = means of pgdL, ngdL, pgdO, ngdO; [b*8 + {4,5,6,7}] = their stds (:1404-1419)
            des[b * 8 + q] = mean;
            des[b * 8 + 4 + q] = __fsqrt_rn(__fsub_rn(__fmul_rn(S2, invN), __fmul_rn(mean, mean)));
        }
        __syncwarp();
        // first normalisation (:1424-1458): lane 0 sums the squared means, lane 1 the squared stds, band after band
        float nrm = 0.f;
        if (lane < 2) {
            const int off = lane * 4;
            for (int b = 0; b < LBD_BANDS; b++)
#pragma unroll
                for (int i = 0; i < 4; i++) { const float v = des[b * 8 + off + i]; nrm = __fadd_rn(nrm, __fmul_rn(v, v)); }
            nrm = __fdiv_rn(1.f, __fsqrt_rn(nrm));
        }
        const float tM = __shfl_sync(FULL, nrm, 0), tS = __shfl_sync(FULL, nrm, 1);
        __syncwarp();
        for (int i = lane; i < 72; i += 32) {
            float v = __fmul_rn(des[i], ((i & 7) < 4) ? tM : tS);
            if (v > 0.4f) v = 0.4f;  // :1465 compares with the double 0.4: float values >= 0.4f end up as 0.4f either way
            des[i] = v;
        }
        __syncwarp();
        float t2 = 0.f;
        if (lane == 0) {
            for (int i = 0; i < 72; i++) { const float v = des[i]; t2 = __fadd_rn(t2, __fmul_rn(v, v)); }
            t2 = __fdiv_rn(1.f, __fsqrt_rn(t2));
        }
        t2 = __shfl_sync(FULL, t2, 0);
        __syncwarp();
        for (int i = lane; i < 72; i += 32) {
            const float v = __fmul_rn(des[i], t2);
            des[i] = v;
            if (A.descf) A.descf[72 * row + i] = v;
        }
        __syncwarp();
        {   // binaryConversion: bit i of byte c = desVec[8 a + i] > desVec[8 b + i]
            const float* f1 = des + 8 * c_lbd_comb[lane][0];
            const float* f2 = des + 8 * c_lbd_comb[lane][1];
            unsigned r = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) r |= (f1[i] > f2[i]) ? (1u << i) : 0u;
            A.desc[32 * row + lane] = (uint8_t)r;
        }
        if (lane == 0) {
            atomicAdd(A.ctr + 1, (unsigned long long)LBD_ROWS * (unsigned long long)len);
            if (A.keyl) {
                const double ddx = (double)(e0 - e2), ddy = (double)(e1 - e3);
                *reinterpret_cast<float4*>(A.keyl + 4 * row) = make_float4(angle, (float)len, (float)sqrt(ddx * ddx + ddy * ddy), 0.f);
            }
        }
        __syncwarp();
    }
}

// BinaryDescriptor ctor (binary_descriptor.cpp:232-258), integer divisions as written there -> the constant-memory band weights
cudaError_t lbd_upload_weights(cudaStream_t st) {
    float G[LBD_ROWS], L[3 * LBD_BAND_W];
    double u = (LBD_BAND_W * 3 - 1) / 2;
    double sigma = (LBD_BAND_W * 2 + 1) / 2;
    double invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < LBD_BAND_W * 3; i++) {
        const double dis = i - u;
        L[i] = (float)std::exp(dis * dis * invsigma2);
    }
    u = (LBD_BANDS * LBD_BAND_W - 1) / 2;
    sigma = u;
    invsigma2 = -1 / (2 * sigma * sigma);
    for (int i = 0; i < LBD_ROWS; i++) {
        const double dis = i - u;
        G[i] = (float)std::exp(dis * dis * invsigma2);
    }
    cudaError_t e = cudaMemcpyToSymbolAsync(c_lbd_G, G, sizeof G, 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbolAsync(c_lbd_L, L, sizeof L, 0, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(st);  // G / L live on this stack frame
}

// Descriptors of detector-supplied key lines (EDLines) with the warp-cooperative kernel: lines / keyl_in / desc rows are
// n_frames x stride, counts[f] lines in frame f (clamped to stride); prefix (n_frames + 1 ints) and ctr (2 x u64) are scratch.
cudaError_t lbd_describe_keylines(const short2* grad, const float* lines, const float2* keyl_in, const int* counts, int n_frames, int stride, int w, int h,
                                  uint8_t* desc, float* descf, int* prefix, unsigned long long* ctr, int num_sms, cudaStream_t st) {
    k_lbd_prefix<<<1, 1024, 0, st>>>(counts, n_frames, stride, prefix, ctr);
    LbdArgs A{};
    A.grad = grad; A.lines = lines; A.prefix = prefix; A.ctr = ctr; A.desc = desc; A.descf = descf; A.keyl = nullptr; A.keyl_in = keyl_in;
    A.w = w; A.h = h; A.n_frames = n_frames; A.stride = stride;
    k_lbd_describe<<<num_sms * LD_CTAS_PER_SM, 32 * LD_WARPS, 0, st>>>(A);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------------
struct LbdState {
    bool uploaded = false, ran = false, timed_last = false, from_lsd = false, weights_set = false;
    int w = 0, h = 0, n_frames = 0, stride = 0;
    int want_float = 0;
    DevBuf d_gray, d_lines, d_counts, d_prefix, d_grad, d_desc, d_descf, d_keyl, d_ctr;
    HostBuf h_in, h_out;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int launches_last = 0;
    // sources of the last run (own buffers or the line detector's)
    const uint8_t* src_gray = nullptr;
    const float* src_lines = nullptr;
    const int* src_counts = nullptr;
};

void lbd_launch_grad(const uint8_t* gray, short2* grad, int w, int h, int n_frames, cudaStream_t st, int blur_generation) {
    const dim3 gg((w + LG_TW - 1) / LG_TW, (h + LG_TH - 1) / LG_TH, n_frames);
    if (blur_generation == 3)  // the taps of OpenCV <= 3.4.0 (csb_set_blur_generation); one pixel per thread, any width
        k_lbd_grad<14, 63, 103, true><<<gg, LG_THREADS, 0, st>>>(gray, grad, w, h);
    else if (w % 4 == 0)
        k_lbd_grad4<<<dim3((w + L4_TW - 1) / L4_TW, (h + L4_TH - 1) / L4_TH, n_frames), L4_THREADS, 0, st>>>(gray, grad, w, h);
    else
        k_lbd_grad<14, 62, 104, false><<<gg, LG_THREADS, 0, st>>>(gray, grad, w, h);
}

void lbd_release(LbdState*& s) {
    if (!s) return;
    DevBuf* bufs[] = {&s->d_gray, &s->d_lines, &s->d_counts, &s->d_prefix, &s->d_grad, &s->d_desc, &s->d_descf, &s->d_keyl, &s->d_ctr};
    for (DevBuf* b : bufs) b->release();
    s->h_in.release();
    s->h_out.release();
    for (auto& e : s->ev)
        if (e) cudaEventDestroy(e);
    delete s;
    s = nullptr;
}

}  // namespace csb

using namespace csb;

static int lbd_state(csb_context* c) {
    CSB_CUDA(c, cudaSetDevice(c->device));
    if (!c->lbd) {
        c->lbd = new LbdState();
        for (auto& e : c->lbd->ev) CSB_CUDA(c, cudaEventCreate(&e));
    }
    LbdState& s = *c->lbd;
    if (!s.weights_set) {
        CSB_CUDA(c, lbd_upload_weights(c->stream));
        s.weights_set = true;
    }
    return CSB_OK;
}

static int lbd_outputs(csb_context* c, LbdState& s) {
    const size_t rows = (size_t)s.n_frames * s.stride;
    CSB_CUDA(c, s.d_prefix.ensure((size_t)(s.n_frames + 1) * 4));
    CSB_CUDA(c, s.d_grad.ensure((size_t)s.n_frames * s.w * s.h * 4));
    CSB_CUDA(c, s.d_desc.ensure_zeroed(std::max<size_t>(rows, 1) * 32, c->stream));
    CSB_CUDA(c, s.d_keyl.ensure_zeroed(std::max<size_t>(rows, 1) * 16, c->stream));
    if (s.want_float) CSB_CUDA(c, s.d_descf.ensure_zeroed(std::max<size_t>(rows, 1) * 72 * 4, c->stream));
    CSB_CUDA(c, s.d_ctr.ensure(64));
    return CSB_OK;
}

static int lbd_launch(csb_context* c, LbdState& s, int timed) {
    cudaStream_t st = c->stream;
    if (timed) CSB_CUDA(c, cudaEventRecord(s.ev[0], st));
    lbd_launch_grad(s.src_gray, s.d_grad.as<short2>(), s.w, s.h, s.n_frames, st, c->blur_generation);
    if (timed) CSB_CUDA(c, cudaEventRecord(s.ev[1], st));
    k_lbd_prefix<<<1, 1024, 0, st>>>(s.src_counts, s.n_frames, s.stride, s.d_prefix.as<int>(), s.d_ctr.as<unsigned long long>());
    LbdArgs A{};
    A.grad = s.d_grad.as<short2>();
    A.lines = s.src_lines;
    A.prefix = s.d_prefix.as<int>();
    A.ctr = s.d_ctr.as<unsigned long long>();
    A.desc = s.d_desc.as<uint8_t>();
    A.descf = s.want_float ? s.d_descf.as<float>() : nullptr;
    A.keyl = s.d_keyl.as<float>();
    A.w = s.w; A.h = s.h; A.n_frames = s.n_frames; A.stride = s.stride;
    k_lbd_describe<<<c->num_sms * LD_CTAS_PER_SM, 32 * LD_WARPS, 0, st>>>(A);
    if (timed) CSB_CUDA(c, cudaEventRecord(s.ev[2], st));
    CSB_CUDA(c, cudaGetLastError());
    s.launches_last = 3;
    s.timed_last = timed != 0;
    s.ran = true;
    return CSB_OK;
}

extern "C" {

int csb_lbd_upload(csb_context* c, const uint8_t* gray, int n_frames, int width, int height, const float* lines, const int32_t* line_offsets,
                   int want_float) {
    if (!c || !gray || !line_offsets || n_frames <= 0 || width < 8 || height < 8 || width > 32767 || height > 32767) return CSB_ERR_INVALID;
    int stride = 0;
    for (int f = 0; f < n_frames; f++) {
        const int n = line_offsets[f + 1] - line_offsets[f];
        if (n < 0) return CSB_ERR_INVALID;
        stride = std::max(stride, n);
    }
    if (line_offsets[n_frames] > line_offsets[0] && !lines) return CSB_ERR_INVALID;
    int rc = lbd_state(c);
    if (rc != CSB_OK) return rc;
    LbdState& s = *c->lbd;
    s.uploaded = false;
    s.ran = false;
    s.from_lsd = false;
    s.w = width; s.h = height; s.n_frames = n_frames; s.stride = std::max(stride, 1);
    s.want_float = want_float != 0;
    const size_t gbytes = (size_t)width * height * n_frames, lbytes = (size_t)n_frames * s.stride * 16, cbytes = (size_t)n_frames * 4;
    CSB_CUDA(c, s.d_gray.ensure(gbytes));
    CSB_CUDA(c, s.d_lines.ensure(lbytes));
    CSB_CUDA(c, s.d_counts.ensure(cbytes));
    rc = lbd_outputs(c, s);
    if (rc != CSB_OK) return rc;
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));  // the staging buffer may still feed an earlier upload
    cudaPointerAttributes pa{};
    const bool pinned = cudaPointerGetAttributes(&pa, gray) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    CSB_CUDA(c, s.h_in.ensure((pinned ? 0 : gbytes) + lbytes + cbytes));
    char* hp = s.h_in.as<char>();
    float* hl = reinterpret_cast<float*>(hp);
    int32_t* hc = reinterpret_cast<int32_t*>(hp + lbytes);
    std::memset(hl, 0, lbytes);
    for (int f = 0; f < n_frames; f++) {
        const int n = line_offsets[f + 1] - line_offsets[f];
        hc[f] = n;
        if (n) std::memcpy(hl + (size_t)f * s.stride * 4, lines + 4 * (size_t)(line_offsets[f] - line_offsets[0]), (size_t)n * 16);
    }
    CSB_CUDA(c, cudaMemcpyAsync(s.d_lines.p, hl, lbytes, cudaMemcpyHostToDevice, c->stream));
    CSB_CUDA(c, cudaMemcpyAsync(s.d_counts.p, hc, cbytes, cudaMemcpyHostToDevice, c->stream));
    if (pinned) {
        CSB_CUDA(c, cudaMemcpyAsync(s.d_gray.p, gray, gbytes, cudaMemcpyHostToDevice, c->stream));
    } else {
        std::memcpy(hp + lbytes + cbytes, gray, gbytes);
        CSB_CUDA(c, cudaMemcpyAsync(s.d_gray.p, hp + lbytes + cbytes, gbytes, cudaMemcpyHostToDevice, c->stream));
    }
    s.h2d_bytes = (int64_t)(gbytes + lbytes + cbytes);
    s.src_gray = s.d_gray.as<uint8_t>();
    s.src_lines = s.d_lines.as<float>();
    s.src_counts = s.d_counts.as<int>();
    s.uploaded = true;
    return CSB_OK;
}

int csb_lbd_run(csb_context* c, int timed) {
    if (!c || !c->lbd || !c->lbd->uploaded) {
        if (c) c->err = "csb_lbd_run before csb_lbd_upload";
        return CSB_ERR_STATE;
    }
    CSB_CUDA(c, cudaSetDevice(c->device));
    return lbd_launch(c, *c->lbd, timed);
}

int csb_lbd_run_on_lsd(csb_context* c, int want_float, int timed) {
    if (!c) return CSB_ERR_INVALID;
    LsdView v{};
    if (!lsd_view(c->lsd, v)) {
        c->err = "csb_lbd_run_on_lsd before csb_lsd_run";
        return CSB_ERR_STATE;
    }
    if (v.w > 32767 || v.h > 32767) return CSB_ERR_INVALID;
    int rc = lbd_state(c);
    if (rc != CSB_OK) return rc;
    LbdState& s = *c->lbd;
    s.uploaded = false;  // the resident inputs of csb_lbd_upload (if any) are not the source of this run
    s.from_lsd = true;
    s.w = v.w; s.h = v.h; s.n_frames = v.n_frames; s.stride = v.max_lines;
    s.want_float = want_float != 0;
    rc = lbd_outputs(c, s);
    if (rc != CSB_OK) return rc;
    s.h2d_bytes = 0;
    s.src_gray = v.gray;
    s.src_lines = v.lines;
    s.src_counts = v.n_lines;
    return lbd_launch(c, s, timed);
}

int csb_lbd_download(csb_context* c, uint8_t* desc_out, float* desc_float_out, float* keylines_out, int32_t* n_lines_out, int64_t capacity_rows,
                     csb_lbd_stats* stats) {
    if (!c || !c->lbd || !c->lbd->ran) {
        if (c) c->err = "csb_lbd_download before csb_lbd_run";
        return CSB_ERR_STATE;
    }
    LbdState& s = *c->lbd;
    if (desc_float_out && !s.want_float) {
        c->err = "csb_lbd_download: float descriptors were not requested at upload / run time";
        return CSB_ERR_STATE;
    }
    CSB_CUDA(c, cudaSetDevice(c->device));
    const size_t nb = (size_t)(s.n_frames + 1) * 4;
    CSB_CUDA(c, s.h_out.ensure(nb + 64));
    CSB_CUDA(c, cudaMemcpyAsync(s.h_out.p, s.d_prefix.p, nb, cudaMemcpyDeviceToHost, c->stream));
    CSB_CUDA(c, cudaMemcpyAsync(s.h_out.as<char>() + nb, s.d_ctr.p, 16, cudaMemcpyDeviceToHost, c->stream));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<int32_t> prefix(s.n_frames + 1);
    std::memcpy(prefix.data(), s.h_out.p, nb);
    unsigned long long ctr[2];
    std::memcpy(ctr, s.h_out.as<char>() + nb, 16);
    const int64_t total = prefix[s.n_frames];
    int rows = 0;
    for (int f = 0; f < s.n_frames; f++) rows = std::max(rows, prefix[f + 1] - prefix[f]);
    const bool fits = total <= capacity_rows;
    s.d2h_bytes = (int64_t)nb + 16;
    if (rows > 0 && fits) {
        const size_t per_row = 32 + (desc_float_out ? 288 : 0) + (keylines_out ? 16 : 0);
        CSB_CUDA(c, s.h_out.ensure((size_t)rows * s.n_frames * per_row));
        char* h = s.h_out.as<char>();
        size_t off = 0;
        auto pull = [&](const void* dev, size_t rb, void* out) -> cudaError_t {
            if (!out) return cudaSuccess;
            char* base = h + off;
            cudaError_t e = cudaMemcpy2DAsync(base, (size_t)rows * rb, dev, (size_t)s.stride * rb, (size_t)rows * rb, s.n_frames, cudaMemcpyDeviceToHost, c->stream);
            if (e != cudaSuccess) return e;
            e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) return e;
            for (int f = 0; f < s.n_frames; f++)
                std::memcpy(reinterpret_cast<char*>(out) + (size_t)prefix[f] * rb, base + (size_t)f * rows * rb, (size_t)(prefix[f + 1] - prefix[f]) * rb);
            off += (size_t)rows * s.n_frames * rb;
            s.d2h_bytes += (int64_t)((size_t)rows * s.n_frames * rb);
            return cudaSuccess;
        };
        CSB_CUDA(c, pull(s.d_desc.p, 32, desc_out));
        CSB_CUDA(c, pull(s.d_descf.p, 288, desc_float_out));
        CSB_CUDA(c, pull(s.d_keyl.p, 16, keylines_out));
    }
    if (n_lines_out)
        for (int f = 0; f < s.n_frames; f++) n_lines_out[f] = prefix[f + 1] - prefix[f];
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->n_lines = total;
        stats->n_samples = (int64_t)ctr[1];
        stats->h2d_bytes = s.h2d_bytes;
        stats->d2h_bytes = s.d2h_bytes;
        stats->n_kernel_launches = s.launches_last;
        if (s.timed_last) {
            cudaEventElapsedTime(&stats->gpu_ms_grad, s.ev[0], s.ev[1]);
            cudaEventElapsedTime(&stats->gpu_ms_describe, s.ev[1], s.ev[2]);
        }
    }
    if (!fits) {
        c->err = "csb_lbd_download: more lines than capacity_rows";
        return CSB_ERR_CAPACITY;
    }
    return CSB_OK;
}

int csb_lbd_describe_batch(csb_context* c, const uint8_t* gray, int n_frames, int width, int height, const float* lines, const int32_t* line_offsets,
                           uint8_t* desc_out, float* desc_float_out, csb_lbd_stats* stats) {
    int rc = csb_lbd_upload(c, gray, n_frames, width, height, lines, line_offsets, desc_float_out != nullptr);
    if (rc != CSB_OK) return rc;
    rc = csb_lbd_run(c, stats != nullptr);
    if (rc != CSB_OK) return rc;
    return csb_lbd_download(c, desc_out, desc_float_out, nullptr, nullptr, (int64_t)line_offsets[n_frames] - line_offsets[0], stats);
}

int csb_lbd_debug_gradients(csb_context* c, int frame, int16_t* dx_out, int16_t* dy_out) {
    if (!c || !c->lbd || !c->lbd->ran) return CSB_ERR_STATE;
    LbdState& s = *c->lbd;
    if (frame < 0 || frame >= s.n_frames) return CSB_ERR_INVALID;
    CSB_CUDA(c, cudaSetDevice(c->device));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)s.w * s.h;
    std::vector<int16_t> both(2 * n);
    CSB_CUDA(c, cudaMemcpy(both.data(), s.d_grad.as<short2>() + n * frame, n * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) {
        if (dx_out) dx_out[i] = both[2 * i];
        if (dy_out) dy_out[i] = both[2 * i + 1];
    }
    return CSB_OK;
}

}  // extern "C"
