// lsd.cu -- LSD line segment detection for a batch of gray frames (SURVEY.md 8 "next" row f-2).
//
// Replaces the LSD branch of line_lbd_detect::detect_filter_lines (reference: line_lbd/class/line_lbd_allclass.cpp:130-149, 200-235 ->
// LSDDetector::detectImpl line_lbd/libs/LSDDetector.cpp:154-293 -> LineSegmentDetectorImpl line_lbd/libs/lsd.cpp:414-1167).  The OpenCV
// calls inside (convertTo, GaussianBlur 7x7 on CV_64F, resize x0.8 INTER_LINEAR, fastAtan2) are restated from OpenCV's algorithms; see
// oracle/oracle_lsd.cpp for how each piece is pinned against cv2 4.13.
//
//   k_lsd_maps  : (64x16 tile of the scaled frame, frame).  u8 tile + halo -> shared memory; separable 7-tap Gaussian in FP64 with
//                 OpenCV's summation order (row filter left to right, column filter centre tap then symmetric pairs); bilinear x0.8
//                 down-scaling with float coefficients into a shared-memory tile; ll_angle (lsd.cpp:538-590) on that tile: 2x2 gradient,
//                 norm, fastAtan2 level-line angle.  Per pixel: a 16-byte record {angle in degrees, cosf(angle), sinf(angle), label slot}
//                 (the two terms region_grow adds per accepted pixel, lsd.cpp:679-680, evaluated once here with the specified
//                 det_sincos), the gradient norm, a compact angle copy, the initial label; the scaled image is also kept (debug API).
//   k_lsd_merge / _flatten / _contact / _units : the work partition -- "units" = defined pixels linked by 8-adjacency and similar
//                 level-line angles (union-find, root = first pixel in raster order), their sizes / bounding boxes, the
//                 defined-neighbour mask of every pixel, the unit list (see the comment above k_lsd_merge for why units are independent
//                 and how that is checked at run time).
//   k_lsd_grow  : one CTA per frame (4 warps, four frames per SM), one warp per unit (dynamic queue, big
//                 units first).  flsd's seed loop (lsd.cpp:474-535) visits seeds in raster order and every region depends on the
//                 `used` map left by the previous ones -- but only inside one unit; units run concurrently and the segments are put back
//                 into seed order at the end.  The `used` map lives in shared memory as a bitmap (24 KB for 512x384; red.or / red.and:
//                 words are shared between units).  A warp first lists its unit's pixels in raster order (label scan of the bounding
//                 box), then walks that list.  region_grow expands one queue entry at a time: lanes 0..8 hold its 3x3 neighbourhood
//                 (defined-neighbour mask from the entry itself, so no bounds tests), the first aligned lane is accepted, the running
//                 angle is updated, the later lanes are tested again -- the reference's sequential order.  The order-dependent FP64 sums
//                 of region2rect / get_theta / refine are accumulated in region order from per-lane products staged in shared memory;
//                 reduce_region_radius' swap-with-last removal is done as two ordered compactions that yield the same permutation;
//                 rect_nfa counts aligned pixels lane-parallel (closed-form scan-line bounds, integer counts).  One seed is a small state
//                 machine with a single call site per building block, which keeps the kernel inside the instruction cache.  Finished
//                 rectangles go through a shared-memory job queue: warps without a unit run the NFA search (rect_improve), which does
//                 not touch the used map.
//
// Arithmetic: FP64, -fmad=false, no libm in anything that decides region membership (det_sincos / fast_atan2f are specified,
// csb_math-style); log/exp/pow/sinh of the NFA use CUDA's libm (they only feed `>` comparisons against 0 and each other).
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

#include "context.h"

namespace csb {

constexpr double LSD_PI = 3.1415926535897932384626433832795;
constexpr double LSD_3_2_PI = (3 * LSD_PI) / 2;
constexpr double LSD_2PI = 2 * LSD_PI;
constexpr double LSD_DEG2RAD = LSD_PI / 180;
constexpr float LSD_NOTDEF_DEG = -1024.f;  // sentinel in the degree map (fastAtan2 returns [0, 360])
constexpr double LSD_SCALE = 0.8;
constexpr int LSD_RING = 512;

struct LsdDims {
    int w, h;     // source frame
    int W, H;     // scaled frame
    int n_frames;
};

struct LsdConst {
    double k[7];       // Gaussian taps (cv::getGaussianKernel(7, 0.6 / 0.8))
    double inv_scale;  // 1 / 0.8
    double rho;        // gradient threshold QUANT / sin(prec)
    double prec, p;    // angle tolerance (rad), its probability
    double log_nt;
    int min_reg_size;
    float length_thres;
    int filter, max_lines;
};

struct LsdBuffers {
    const uint8_t* gray;
    double* scaled;
    float4* pix;       // {deg, cosf, sinf, (unit label << 9 | defined-neighbour mask) as int bits}; mask bit k <-> neighbour (k / 3 - 1, k % 3 - 1)
    float* deg;        // compact copy of pix.x for the NFA counts and the labelling
    double* modgrad;
    // work arenas, W*H entries per frame each.  Arena A serves the units of the first round, arena B the merged units of later rounds.
    uint32_t *regA, *tmpA, *lstA, *regB, *tmpB, *lstB;  // region list (x | y << 16), scratch of reduce_region_radius, a unit's pixels in raster order
    int* label;        // unit label of every pixel: root = smallest pixel index of its unit; -1 = undefined
    int* csize;        // per root pixel: unit size, min x, max x, max y (aggregated into the leader when units are merged)
    int* cminx;
    int* cmaxx;
    int* cmaxy;
    int* cflag;        // per root pixel: bit 0 = some pixel of the unit touches a defined pixel of another unit
    int* cgrp;         // per root pixel: union-find parent over units (merging of interacting units)
    int* cmark;        // per root pixel: round in which the root became the leader of a merged unit
    int* units;        // n_frames x W*H: roots of the first-round units, big ones from the front, small ones from the back
    int* units2;       // n_frames x W*H: leaders of the merged units of the current round
    int* viol;         // n_frames x 2 x W*H: (unit leader, foreign root) pairs recorded in the current round
    int* ncomp;        // n_frames x 4: [0] big units, [1] small units
    float4* stage;     // n_frames x stage_cap segments in completion order ...
    int* stage_key;    // ... the pixel index of the seed each one grew from (the reference's output order; -1 = discarded) ...
    int* stage_owner;  // ... and the root of the unit that produced it
    float* lines;      // n_frames x max_lines x 4
    int* n_lines;      // n_frames
    unsigned long long* stats;  // [0] regions, [1] region pixels, [2..5] SM cycles per phase, [6] SM cycles of the whole seed loop (summed over warps), [7] merge rounds, [8] violations
    int stage_cap;
};

// ---------------------------------------------------------------------------------------------------------------------------------
// specified arithmetic (same sequences as oracle/oracle_lsd.cpp)
// ---------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void det_sincos(double x, double& s, double& c) {
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

// cv::fastAtan2, degrees in [0, 360]
__device__ __forceinline__ float fast_atan2f(float y, float x) {
    constexpr float p1 = 0.9997878412794807f * (float)(180 / LSD_PI), p3 = -0.3258083974640975f * (float)(180 / LSD_PI),
                    p5 = 0.1555786518463281f * (float)(180 / LSD_PI), p7 = -0.04432655554792128f * (float)(180 / LSD_PI);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
    return i;
}

// cv::resize's source column for destination column d (xofs / alpha of resizeGeneric): fraction zeroed when clamped
__device__ __forceinline__ int resize_src_x(int d, double inv, int n, float& f) {
    f = (float)((d + 0.5) * inv - 0.5);
    int s = (int)floorf(f);
    f -= s;
    if (s < 0) { f = 0; s = 0; }
    if (s >= n - 1) { f = 0; s = n - 1; }
    return s;
}
// ... and source row (yofs / beta): rows are clamped when read, the fraction is kept
__device__ __forceinline__ int resize_src_y(int d, double inv, float& f) {
    f = (float)((d + 0.5) * inv - 0.5);
    const int s = (int)floorf(f);
    f -= s;
    return s;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// k_lsd_maps (cv::GaussianBlur + cv::resize of flsd, lsd.cpp:449-460, and ll_angle, :538-590)
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr int SC_TW = 64, SC_TH = 16, SC_THREADS = 256;
constexpr int SC_BC = 88, SC_BR = 26;  // blurred tile capacity: source columns / rows feeding (SC_TW + 1) x (SC_TH + 1) scaled pixels
constexpr int SC_GP = SC_BC + 8;       // u8 tile pitch
constexpr int SC_SP = SC_TW + 1;       // scaled tile pitch

// One CTA per 64 x 16 tile of the scaled frame: u8 tile -> row filter -> column filter -> bilinear x0.8 -> (SC_TW + 1) x (SC_TH + 1)
// scaled pixels in shared memory -> the 2x2 gradient of ll_angle -> per-pixel outputs.  Thread (tx, ty) = (tid & 31, tid >> 5) strides a
// tile stage by 32 columns and 8 rows, so no stage needs a division.
__global__ void __launch_bounds__(SC_THREADS) k_lsd_maps(LsdBuffers B, LsdDims d, LsdConst C) {
    __shared__ uint8_t g[(SC_BR + 6) * SC_GP];
    __shared__ double rowf[(SC_BR + 6) * SC_BC];  // re-used for the scaled tile once the column filter is done
    __shared__ double blur[SC_BR * SC_BC];
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5, f = blockIdx.z;
    const int X0 = blockIdx.x * SC_TW, Y0 = blockIdx.y * SC_TH;
    const int X1 = min(X0 + SC_TW, d.W - 1), Y1 = min(Y0 + SC_TH, d.H - 1);  // one scaled column / row beyond the tile for the gradient
    const uint8_t* src = B.gray + (size_t)f * d.w * d.h;
    float fr;
    const int sx_lo = resize_src_x(X0, C.inv_scale, d.w, fr);
    const int sx_hi = min(resize_src_x(X1, C.inv_scale, d.w, fr) + 1, d.w - 1);
    const int sy_lo = min(max(resize_src_y(Y0, C.inv_scale, fr), 0), d.h - 1);
    const int sy_hi = min(max(resize_src_y(Y1, C.inv_scale, fr) + 1, 0), d.h - 1);
    const int nc = sx_hi - sx_lo + 1, nr = sy_hi - sy_lo + 1;  // <= SC_BC, SC_BR
    // (1) u8 tile: virtual rows sy_lo-3 .. sy_hi+3, virtual columns sx_lo-3 .. sx_hi+3, BORDER_REFLECT_101
    for (int r = ty; r < nr + 6; r += 8) {
        const uint8_t* row = src + (size_t)reflect101(sy_lo - 3 + r, d.h) * d.w;
        for (int c = tx; c < nc + 6; c += 32) g[r * SC_GP + c] = row[reflect101(sx_lo - 3 + c, d.w)];
    }
    __syncthreads();
    // (2) row filter, taps summed left to right (cv::RowFilter)
    for (int r = ty; r < nr + 6; r += 8)
        for (int c = tx; c < nc; c += 32) {
            const uint8_t* s = g + r * SC_GP + c;
            double acc = C.k[0] * (double)s[0];
#pragma unroll
            for (int t = 1; t < 7; t++) acc += C.k[t] * (double)s[t];
            rowf[r * SC_BC + c] = acc;
        }
    __syncthreads();
    // (3) column filter (cv::SymmColumnFilter): centre tap, then f_k (S[+k] + S[-k])
    for (int r = ty; r < nr; r += 8)
        for (int c = tx; c < nc; c += 32) {
            const double* s = rowf + (r + 3) * SC_BC + c;
            double acc = C.k[3] * s[0];
#pragma unroll
            for (int t = 1; t <= 3; t++) acc += C.k[3 + t] * (s[t * SC_BC] + s[-t * SC_BC]);
            blur[r * SC_BC + c] = acc;
        }
    __syncthreads();
    // (4) bilinear down-scaling (HResizeLinear then VResizeLinear, float coefficients) of scaled rows Y0 .. Y1, columns X0 .. X1
    double* sc = rowf;
    for (int yy = ty; yy <= Y1 - Y0; yy += 8) {
        const int dy = Y0 + yy;
        float fy;
        const int sy = resize_src_y(dy, C.inv_scale, fy);
        const float b1 = fy, b0 = 1.f - fy;
        const int y0 = min(max(sy, 0), d.h - 1) - sy_lo, y1 = min(max(sy + 1, 0), d.h - 1) - sy_lo;
        for (int xx = tx; xx <= X1 - X0; xx += 32) {
            const int dx = X0 + xx;
            float fx;
            const int sx = resize_src_x(dx, C.inv_scale, d.w, fx);
            const float a1 = fx, a0 = 1.f - fx;
            const double* r0p = blur + y0 * SC_BC + (sx - sx_lo);
            const double* r1p = blur + y1 * SC_BC + (sx - sx_lo);
            double r0, r1;
            if (sx + 1 < d.w) {
                r0 = r0p[0] * (double)a0 + r0p[1] * (double)a1;
                r1 = r1p[0] * (double)a0 + r1p[1] * (double)a1;
            } else {
                r0 = r0p[0];
                r1 = r1p[0];
            }
            const double v = r0 * (double)b0 + r1 * (double)b1;
            sc[yy * SC_SP + xx] = v;
            if (B.scaled && xx < SC_TW && yy < SC_TH) B.scaled[((size_t)f * d.H + dy) * d.W + dx] = v;  // only for csb_lsd_debug_maps (8 B per scaled pixel otherwise)
        }
    }
    __syncthreads();
    // (5) ll_angle (lsd.cpp:538-590): 2x2 gradient, norm, fastAtan2 level-line angle; the records region_grow reads
    const size_t fo = (size_t)f * d.W * d.H;
    for (int yy = ty; yy < SC_TH; yy += 8) {
        const int y = Y0 + yy;
        if (y >= d.H) break;
        for (int xx = tx; xx < SC_TW; xx += 32) {
            const int x = X0 + xx;
            if (x >= d.W) break;
            float deg = LSD_NOTDEF_DEG, cv = 0.f, sv = 0.f;
            double mg = 0.0;
            bool def = false;
            if (x < d.W - 1 && y < d.H - 1) {
                const double* S = sc + yy * SC_SP + xx;
                const double a = S[0], b = S[1], c = S[SC_SP], e = S[SC_SP + 1];
                const double DA = e - a, BC = b - c;
                const double gx = DA + BC, gy = DA - BC;
                const double norm = sqrt((gx * gx + gy * gy) / 4);
                mg = norm;
                if (norm > C.rho) {
                    def = true;
                    deg = fast_atan2f((float)gx, (float)(-gy));
                    const double ang = (double)deg * LSD_DEG2RAD;
                    double s, c2;
                    det_sincos((double)(float)ang, s, c2);
                    cv = (float)c2;
                    sv = (float)s;
                }
            }
            const size_t pi = fo + (size_t)y * d.W + x;
            B.pix[pi] = make_float4(deg, cv, sv, 0.f);
            B.deg[pi] = deg;
            B.modgrad[pi] = mg;
            B.label[pi] = def ? y * d.W + x : -1;
            if (def) {
                B.csize[pi] = 0;
                B.cminx[pi] = 0x7fffffff;
                B.cmaxx[pi] = -1;
                B.cmaxy[pi] = -1;
                B.cflag[pi] = 0;
                B.cmark[pi] = 0;
                B.cgrp[pi] = y * d.W + x;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Units of independent work.  region_grow (lsd.cpp:637-688) only ever adds DEFINED 8-neighbours that are aligned with the running
// region angle, and the `used` state of a pixel is only read and written by regions that test it.  The defined pixels are therefore
// partitioned into "units": connected sets under the relation "8-adjacent, both defined, level-line angles within LSD_LINK_DEG of
// each other" (a straight edge is one unit; at a corner or a crossing the angle jumps and the units split).  A unit is processed in
// isolation by one warp, seeds in raster order, and is VALID iff none of its regions ever finds a pixel of another unit aligned --
// whatever that pixel's used state -- because then neither side can observe the other.  A unit that does find one records the pair
// and stops; interacting units are merged (union-find over unit roots) and the merged unit is redone, round after round, until every
// unit is valid; the result is then exactly the reference's sequential one.  Units smaller than min_reg_size that touch no other unit
// are whole connected components that cannot yield a segment (lsd.cpp:489) and are skipped.
//   k_lsd_merge   : union-find with atomicMin (root = smallest pixel index) over the W / NW / N / NE neighbours
//   k_lsd_flatten : label <- root (also into pix.w); per-root size and bounding box (warp-aggregated atomics)
//   k_lsd_contact : per-root flag "touches another unit"
//   k_lsd_units   : roots of the first-round units, big ones first (they bound the frame's critical path)
// ---------------------------------------------------------------------------------------------------------------------------------
constexpr float LSD_LINK_DEG = 45.f;  // default; csb_lsd_params::unit_link_deg overrides it (tests use small values to force merge rounds)

__device__ __forceinline__ int ccl_find(const int* L, int a) {
    int r = L[a];
    while (r != a) {
        a = r;
        r = L[a];
    }
    return a;
}
__device__ __forceinline__ void ccl_union(int* L, int a, int b) {
    while (true) {
        a = ccl_find(L, a);
        b = ccl_find(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}
__device__ __forceinline__ bool lsd_linked(float a, float b, float link_deg) {  // both defined
    float dd = fabsf(a - b);
    if (dd > 180.f) dd = 360.f - dd;
    return dd <= link_deg;
}

__global__ void __launch_bounds__(256) k_lsd_merge(LsdBuffers B, LsdDims d, float link_deg) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, f = blockIdx.z;
    if (x >= d.W || y >= d.H) return;
    const size_t fo = (size_t)f * d.W * d.H;
    int* L = B.label + fo;
    const float* A = B.deg + fo;
    const int p = y * d.W + x;
    const float a = A[p];
    if (a == LSD_NOTDEF_DEG) return;
    if (x > 0 && A[p - 1] != LSD_NOTDEF_DEG && lsd_linked(a, A[p - 1], link_deg)) ccl_union(L, p, p - 1);
    if (y > 0) {
        const int q = p - d.W;
        if (x > 0 && A[q - 1] != LSD_NOTDEF_DEG && lsd_linked(a, A[q - 1], link_deg)) ccl_union(L, p, q - 1);
        if (A[q] != LSD_NOTDEF_DEG && lsd_linked(a, A[q], link_deg)) ccl_union(L, p, q);
        if (x + 1 < d.W && A[q + 1] != LSD_NOTDEF_DEG && lsd_linked(a, A[q + 1], link_deg)) ccl_union(L, p, q + 1);
    }
}

__global__ void __launch_bounds__(256) k_lsd_flatten(LsdBuffers B, LsdDims d) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, f = blockIdx.z;
    const size_t fo = (size_t)f * d.W * d.H;
    int* L = B.label + fo;
    int r = -1;
    if (x < d.W && y < d.H) {
        const int p = y * d.W + x;
        if (L[p] >= 0) {
            r = ccl_find(L, p);
            L[p] = r;
            unsigned mask = 0;  // defined 8-neighbours (a negative label never turns non-negative or back while labels are flattened)
            for (int k = 0; k < 9; k++) {
                const int xx = x + k % 3 - 1, yy = y + k / 3 - 1;
                if (k != 4 && xx >= 0 && xx < d.W && yy >= 0 && yy < d.H && L[yy * d.W + xx] >= 0) mask |= 1u << k;
            }
            B.pix[fo + p].w = __uint_as_float(((unsigned)r << 9) | mask);
        }
    }
    const unsigned act = __ballot_sync(0xffffffffu, r >= 0);
    if (r < 0) return;
    const unsigned grp = __match_any_sync(act, r);
    const int mn = __reduce_min_sync(grp, x), mx = __reduce_max_sync(grp, x);
    if ((int)(threadIdx.x) == __ffs(grp) - 1) {
        atomicAdd(&B.csize[fo + r], __popc(grp));
        atomicMin(&B.cminx[fo + r], mn);
        atomicMax(&B.cmaxx[fo + r], mx);
        atomicMax(&B.cmaxy[fo + r], y);
    }
}

__global__ void __launch_bounds__(256) k_lsd_contact(LsdBuffers B, LsdDims d) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, f = blockIdx.z;
    if (x >= d.W || y >= d.H) return;
    const size_t fo = (size_t)f * d.W * d.H;
    const int* L = B.label + fo;
    const int p = y * d.W + x, r = L[p];
    if (r < 0) return;
    bool touch = false;
    for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            const int xx = x + dx, yy = y + dy;
            if (xx < 0 || xx >= d.W || yy < 0 || yy >= d.H) continue;
            const int l = L[yy * d.W + xx];
            touch |= l >= 0 && l != r;
        }
    if (touch) B.cflag[fo + r] = 1;
}

constexpr int LSD_BIG_UNIT = 128;

__global__ void __launch_bounds__(256) k_lsd_units(LsdBuffers B, LsdDims d, int min_reg_size) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, f = blockIdx.z;
    if (x >= d.W || y >= d.H) return;
    const size_t fo = (size_t)f * d.W * d.H;
    const int p = y * d.W + x;
    if (B.label[fo + p] != p) return;
    const int size = B.csize[fo + p];
    if (size < min_reg_size && !B.cflag[fo + p]) return;
    int* nc = B.ncomp + 4 * f;
    const int slot = size >= LSD_BIG_UNIT ? atomicAdd(&nc[0], 1) : d.W * d.H - 1 - atomicAdd(&nc[1], 1);
    B.units[fo + slot] = p;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// k_lsd_grow
// ---------------------------------------------------------------------------------------------------------------------------------
struct LRect {
    double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p;
};

__device__ __forceinline__ double l_distSq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
__device__ __forceinline__ double l_dist(double x1, double y1, double x2, double y2) { return sqrt(l_distSq(x1, y1, x2, y2)); }
__device__ __forceinline__ double l_angle_diff_signed(double a, double b) {
    double diff = a - b;
    while (diff <= -LSD_PI) diff += LSD_2PI;
    while (diff > LSD_PI) diff -= LSD_2PI;
    return diff;
}
__device__ __forceinline__ bool l_double_equal(double a, double b) {
    if (a == b) return true;
    const double abs_diff = fabs(a - b), aa = fabs(a), bb = fabs(b);
    double abs_max = (aa > bb) ? aa : bb;
    if (abs_max < DBL_MIN) abs_max = DBL_MIN;
    return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
// lsd.cpp:1151-1167 on a defined pixel whose angle is deg degrees
__device__ __forceinline__ bool l_aligned_deg(float deg, double theta, double prec) {
    const double a = (double)deg * LSD_DEG2RAD;
    double n_theta = theta - a;
    if (n_theta < 0) n_theta = -n_theta;
    if (n_theta > LSD_3_2_PI) {
        n_theta -= LSD_2PI;
        if (n_theta < 0) n_theta = -n_theta;
    }
    return n_theta <= prec;
}
// The same decision as l_aligned_deg((deg), (double)reg_deg * DEG2RAD, prec), taken in float degrees when it is not within 1e-3 degrees
// of a threshold (the FP64 expressions are accurate to 1e-13 degrees, the float difference to 4e-5).
__device__ __forceinline__ bool l_aligned_fast(float deg, float reg_deg, float prec_deg, double prec) {
    float dd = fabsf(reg_deg - deg);
    const bool wrap = dd > 270.f;
    const float d2 = wrap ? fabsf(dd - 360.f) : dd;
    if (fabsf(d2 - prec_deg) > 1e-3f && fabsf(dd - 270.f) > 1e-3f) return d2 < prec_deg;
    return l_aligned_deg(deg, (double)reg_deg * LSD_DEG2RAD, prec);
}
__device__ double l_log_gamma(double x) {
    if (x > 15.0) return 0.918938533204673 + (x - 0.5) * log(x) - x + 0.5 * x * log(x * sinh(1 / x) + 1 / (810.0 * pow(x, 6.0)));
    const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
    double a = (x + 0.5) * log(x + 5.5) - (x + 5.5);
    double b = 0;
    for (int n = 0; n < 7; ++n) {
        a -= log(x + double(n));
        b += q[n] * pow(x, double(n));
    }
    return a + log(b);
}
// lsd.cpp:1100-1149 (the first term is n + 1, not log_gamma(n + 1), in the reference's copy)
__device__ __noinline__ double l_nfa(int n, int k, double p, double LOG_NT) {
    if (n == 0 || k == 0) return -LOG_NT;
    if (n == k) return -LOG_NT - double(n) * log10(p);
    const double p_term = p / (1 - p);
    const double log1term = (double(n) + 1) - l_log_gamma(double(k) + 1) - l_log_gamma(double(n - k) + 1) + double(k) * log(p) +
                            double(n - k) * log(1.0 - p);
    double term = exp(log1term);
    if (l_double_equal(term, 0)) {
        if (k > n * p) return -log1term / 2.30258509299404568402 - LOG_NT;
        return -LOG_NT;
    }
    double bin_tail = term;
    const double tolerance = 0.1;
    for (int i = k + 1; i <= n; ++i) {
        const double bin_term = double(n - i + 1) / double(i);
        const double mult_term = bin_term * p_term;
        term *= mult_term;
        bin_tail += term;
        if (bin_term < 1) {
            const double err = term * ((1 - pow(mult_term, double(n - i + 1))) / (1 - mult_term) - 1);
            if (err < tolerance * fabs(-log10(bin_tail) - LOG_NT) * bin_tail) break;
        }
    }
    return -log10(bin_tail) - LOG_NT;
}

// explicit shared-window accesses (generic pointers into dynamic shared memory cost a window-base computation per access)
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void atoms_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void atoms_and(uint32_t a, uint32_t v) { asm volatile("red.shared.and.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

struct Grow {
    const float4* pix;
    const float* deg;
    const double* mg;
    const int* cgrp;    // union-find parents over unit roots (merged units)
    uint32_t* reg;      // region list: pixel indices (y * W + x)
    uint32_t* tmp;
    uint32_t U;         // shared address of the used bitmap: bit i <-> pixel index i
    uint32_t ring;      // shared address of the last LSD_RING queue entries (pixel index | defined-neighbour mask << 23)
    double* sc;         // shared: 3 x 32 doubles
    int W, H, lane;
    int leader;         // root of the unit being processed
    int mode;           // 0: first-round unit (foreign <=> label != leader); 1: merged unit (find(label) != leader); 2: whole frame (nothing is foreign)
    int foreign_root;   // set when a region found a pixel of another unit aligned: the unit must be merged with that one and redone
    double LOG_NT;
    unsigned long long n_regions, n_px;

    __device__ __forceinline__ bool used(int idx) const { return (lds32(U + 4u * (unsigned)(idx >> 5)) >> (idx & 31)) & 1u; }
    __device__ __forceinline__ void set_used(int idx) const { atoms_or(U + 4u * (unsigned)(idx >> 5), 1u << (idx & 31)); }
    __device__ __forceinline__ void clear_used(int idx) const { atoms_and(U + 4u * (unsigned)(idx >> 5), ~(1u << (idx & 31))); }
    __device__ __forceinline__ bool is_foreign(int lbl) const {
        if (mode == 0) return lbl != leader;
        if (mode == 1) return lbl != leader && ccl_find(cgrp, lbl) != leader;
        return false;
    }

    // lsd.cpp:637-688.  Returns the region size (the region is reg[0 .. size)), or -1 if a pixel of another unit was found aligned.
    // A queue entry is a pixel index plus the mask of its defined 8-neighbours (precomputed per pixel, so the expansion needs neither
    // bounds tests nor a "defined" lookup); the queue head is read from a shared-memory ring (the last LSD_RING entries; older ones are
    // rebuilt from the region list).  Lanes 0..8 hold the 3x3 neighbourhood of the entry being expanded; every defined neighbour is
    // loaded: its record carries the label that tells own pixels (candidates while unused) from foreign ones (tested whatever their used
    // state: finding one aligned invalidates the unit).  The accept loop replays the reference's order: the first aligned lane is
    // accepted, the running angle is updated, the lanes after it are tested again with the new angle.
    __device__ __forceinline__ int region_grow(int seed, double prec, double& reg_angle_out) {
        const float4 vs = pix[seed];
        if (lane == 0) {
            reg[0] = (uint32_t)seed;
            sts32(ring, (uint32_t)seed | (__float_as_uint(vs.w) << 23));
            set_used(seed);
        }
        int n = 1;
        float reg_deg = vs.x;  // the region angle is always a fastAtan2 value: degrees in float, radians = deg * (pi / 180) in double
        const float prec_deg = (float)(prec * (180.0 / LSD_PI));
        double s0, c0;
        det_sincos((double)reg_deg * LSD_DEG2RAD, s0, c0);
        float sumdx = (float)c0, sumdy = (float)s0;
        const int off = (lane / 3 - 1) * W + (lane % 3 - 1);
        const uint32_t lanebit = lane < 9 ? 1u << (23 + lane) : 0u;
        __syncwarp();
        for (int i = 0; i < n; ++i) {
            uint32_t e;
            if (n - i <= LSD_RING) e = lds32(ring + 4u * (unsigned)(i & (LSD_RING - 1)));
            else {
                const uint32_t r = reg[i];
                e = r | (__float_as_uint(pix[r].w) << 23);
            }
            const bool def = e & lanebit;
            const int idx = (int)(e & 0x7fffffu) + off;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (def) v = pix[idx];
            const uint32_t wv = __float_as_uint(v.w);
            const bool frn = def && is_foreign((int)(wv >> 9));
            const bool cand = def && (frn || !used(idx));
            unsigned m = __ballot_sync(0xffffffffu, cand);
            while (m) {
                const bool al = cand && l_aligned_fast(v.x, reg_deg, prec_deg, prec);
                const unsigned am = __ballot_sync(0xffffffffu, al) & m;
                if (!am) break;
                const int l = __ffs(am) - 1;
                if (__shfl_sync(0xffffffffu, (int)frn, l)) {
                    foreign_root = (int)(__shfl_sync(0xffffffffu, wv, l) >> 9);
                    break;
                }
                if (lane == l) {
                    set_used(idx);
                    reg[n] = (uint32_t)idx;
                    sts32(ring + 4u * (unsigned)(n & (LSD_RING - 1)), (uint32_t)idx | (wv << 23));
                }
                sumdx += __shfl_sync(0xffffffffu, v.y, l);
                sumdy += __shfl_sync(0xffffffffu, v.z, l);
                reg_deg = fast_atan2f(sumdy, sumdx);
                ++n;
                m &= ~((2u << l) - 1u);
            }
            __syncwarp();
            if (foreign_root >= 0) return -1;
        }
        reg_angle_out = (double)reg_deg * LSD_DEG2RAD;
        return n;
    }

    // lsd.cpp:690-746 + get_theta :748-784
    __device__ __forceinline__ void region2rect(int n, double reg_angle, double prec, double p, LRect& rec) {
        double x = 0, y = 0, sum = 0;
        for (int base = 0; base < n; base += 32) {
            const int idx = base + lane, cnt = min(32, n - base);
            if (idx < n) {
                const int q = (int)reg[idx];
                const int py = q / W, px = q - py * W;
                const double w = mg[q];
                sc[lane] = double(px) * w;
                sc[32 + lane] = double(py) * w;
                sc[64 + lane] = w;
            }
            __syncwarp();
            for (int j = 0; j < cnt; j++) {
                x += sc[j];
                y += sc[32 + j];
                sum += sc[64 + j];
            }
            __syncwarp();
        }
        x /= sum;
        y /= sum;
        double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
        for (int base = 0; base < n; base += 32) {
            const int idx = base + lane, cnt = min(32, n - base);
            if (idx < n) {
                const int q = (int)reg[idx];
                const int py = q / W, px = q - py * W;
                const double w = mg[q];
                const double ddx = double(px) - x, ddy = double(py) - y;
                sc[lane] = ddy * ddy * w;
                sc[32 + lane] = ddx * ddx * w;
                sc[64 + lane] = ddx * ddy * w;
            }
            __syncwarp();
            for (int j = 0; j < cnt; j++) {
                Ixx += sc[j];
                Iyy += sc[32 + j];
                Ixy -= sc[64 + j];
            }
            __syncwarp();
        }
        const double lambda = 0.5 * (Ixx + Iyy - sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
        double theta = (fabs(Ixx) > fabs(Iyy)) ? double(fast_atan2f(float(lambda - Ixx), float(Ixy))) : double(fast_atan2f(float(Ixy), float(lambda - Iyy)));
        theta *= LSD_DEG2RAD;
        if (fabs(l_angle_diff_signed(theta, reg_angle)) > prec) theta += LSD_PI;
        double dx, dy;
        det_sincos(theta, dy, dx);
        double l_min = 0, l_max = 0, w_min = 0, w_max = 0;  // the reference's running min / max from 0 (order-free)
        for (int idx = lane; idx < n; idx += 32) {
            const int q = (int)reg[idx];
            const int py = q / W, px = q - py * W;
            const double regdx = double(px) - x, regdy = double(py) - y;
            const double l = regdx * dx + regdy * dy;
            const double w = -regdx * dy + regdy * dx;
            l_max = fmax(l_max, l);
            l_min = fmin(l_min, l);
            w_max = fmax(w_max, w);
            w_min = fmin(w_min, w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l_max = fmax(l_max, __shfl_xor_sync(0xffffffffu, l_max, o));
            l_min = fmin(l_min, __shfl_xor_sync(0xffffffffu, l_min, o));
            w_max = fmax(w_max, __shfl_xor_sync(0xffffffffu, w_max, o));
            w_min = fmin(w_min, __shfl_xor_sync(0xffffffffu, w_min, o));
        }
        rec.x1 = x + l_min * dx;
        rec.y1 = y + l_min * dy;
        rec.x2 = x + l_max * dx;
        rec.y2 = y + l_max * dy;
        rec.width = w_max - w_min;
        rec.x = x;
        rec.y = y;
        rec.theta = theta;
        rec.dx = dx;
        rec.dy = dy;
        rec.prec = prec;
        rec.p = p;
        if (rec.width < 1.0) rec.width = 1.0;
    }

    // One pass of the loop of reduce_region_radius (lsd.cpp:845-868): drop the points farther than sqrt(radSq) from the seed.  The
    // reference removes them one by one with swap(reg[i], reg[size - 1]); the resulting order is: kept points stay, holes among the
    // first m = #kept slots (ascending) are filled by the kept points of slots >= m in descending slot order.  Returns the new size.
    __device__ __forceinline__ int reduce_step(int n, double xc, double yc, double radSq) {
        int m = 0;
        for (int base = 0; base < n; base += 32) {
            const int idx = base + lane;
            bool keep = false;
            if (idx < n) {
                const int q = (int)reg[idx];
                const int py = q / W, px = q - py * W;
                keep = !(l_distSq(xc, yc, double(px), double(py)) > radSq);
                if (!keep) clear_used(q);
            }
            m += __popc(__ballot_sync(0xffffffffu, keep));
        }
        if (m != n) {
            int r = 0;
            for (int top = n - 1; top >= m; top -= 32) {  // kept points of slots >= m, descending
                const int pos = top - lane;
                bool keep = false;
                uint32_t q = 0;
                if (pos >= m) {
                    q = reg[pos];
                    keep = !(l_distSq(xc, yc, double((int)q % W), double((int)q / W)) > radSq);
                }
                const unsigned b = __ballot_sync(0xffffffffu, keep);
                if (keep) tmp[r + __popc(b & ((1u << lane) - 1u))] = q;
                r += __popc(b);
            }
            __syncwarp();
            r = 0;
            for (int base = 0; base < m; base += 32) {  // holes of slots < m, ascending
                const int idx = base + lane;
                bool far = false;
                if (idx < m) {
                    const int q = (int)reg[idx];
                    far = l_distSq(xc, yc, double(q % W), double(q / W)) > radSq;
                }
                const unsigned b = __ballot_sync(0xffffffffu, far);
                if (far) reg[idx] = tmp[r + __popc(b & ((1u << lane) - 1u))];
                r += __popc(b);
            }
            __syncwarp();
        }
        return m;
    }

    // First half of refine() (lsd.cpp:794-813): release the region and derive the tighter tolerance tau from the angle spread near the seed.
    __device__ __forceinline__ double tau_step(int n, double width) {
        const int q0 = (int)reg[0];
        const double xc = double(q0 % W), yc = double(q0 / W);
        const double ang_c = (double)deg[q0] * LSD_DEG2RAD;
        double sum = 0, s_sum = 0;
        int cnt = 0;
        for (int base = 0; base < n; base += 32) {
            const int idx = base + lane;
            bool near = false;
            if (idx < n) {
                const int q = (int)reg[idx];
                const int py = q / W, px = q - py * W;
                clear_used(q);
                if (l_dist(xc, yc, double(px), double(py)) < width) {
                    near = true;
                    sc[lane] = l_angle_diff_signed((double)deg[q] * LSD_DEG2RAD, ang_c);
                }
            }
            unsigned m = __ballot_sync(0xffffffffu, near);
            while (m) {
                const int l = __ffs(m) - 1;
                m &= m - 1;
                const double d = sc[l];
                sum += d;
                s_sum += d * d;
                ++cnt;
            }
            __syncwarp();
        }
        const double mean_angle = sum / double(cnt);
        return 2.0 * sqrt((s_sum - 2.0 * mean_angle * sum) / double(cnt) + mean_angle * mean_angle);
    }

};

// Everything after refine(): the NFA search over rectangle variants and the output filters.  None of it touches the used map, so a
// finished rectangle can be handed to any warp of the CTA (job queue in k_lsd_grow); the producing warp goes on growing regions.
struct NfaCtx {
    const float* deg;
    int W, H, w, h, lane;
    double LOG_NT;
    int filter;
    float length_thres;
    float4* stage;
    int* stage_key;
    int* stage_owner;
    int stage_cap;
    int* nout;  // shared counter of staged segments

    // lsd.cpp:977-1098.  Integer-division slopes and the tailp->p.x comparisons are the reference's; since every step is an integer the
    // scan-line bounds of row y have a closed form, so rows can be counted in any order.
    __device__ __forceinline__ double rect_nfa(const LRect& rec) const {
        const double half_width = rec.width / 2.0;
        const double dyhw = rec.dy * half_width, dxhw = rec.dx * half_width;
        int ex[4], ey[4];
        ex[0] = int(rec.x1 - dyhw); ey[0] = int(rec.y1 + dxhw);
        ex[1] = int(rec.x2 - dyhw); ey[1] = int(rec.y2 + dxhw);
        ex[2] = int(rec.x2 + dyhw); ey[2] = int(rec.y2 - dxhw);
        ex[3] = int(rec.x1 + dyhw); ey[3] = int(rec.y1 - dxhw);
#define LSD_CSWAP(a, b)                                                              \
    if (ex[b] < ex[a] || (ex[b] == ex[a] && ey[b] < ey[a])) {                        \
        int t_ = ex[a]; ex[a] = ex[b]; ex[b] = t_; t_ = ey[a]; ey[a] = ey[b]; ey[b] = t_; \
    }
        LSD_CSWAP(0, 1) LSD_CSWAP(2, 3) LSD_CSWAP(0, 2) LSD_CSWAP(1, 3) LSD_CSWAP(1, 2)
#undef LSD_CSWAP
        int mi = 0, miy = ey[0], mix = ex[0], may = ey[0];
#pragma unroll
        for (int i = 1; i < 4; i++) {
            if (miy > ey[i]) { mi = i; miy = ey[i]; mix = ex[i]; }
            if (may < ey[i]) may = ey[i];
        }
        unsigned taken = 1u << mi;
        int lx_ = 0, ly_ = 0, li = -1;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (!(taken >> i & 1u)) {
                if (li < 0 || lx_ > ex[i]) { li = i; lx_ = ex[i]; ly_ = ey[i]; }
            }
        taken |= 1u << li;
        int rx_ = 0, ry_ = 0, ri = -1;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (!(taken >> i & 1u)) {
                if (ri < 0 || rx_ < ex[i]) { ri = i; rx_ = ex[i]; ry_ = ey[i]; }
            }
        taken |= 1u << ri;
        int tx_ = 0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (!(taken >> i & 1u)) tx_ = ex[i];
        const int flstep = (miy != ly_) ? (mix - lx_) / (miy - ly_) : 0;
        const int slstep = (ly_ != tx_) ? (lx_ - tx_) / (ly_ - tx_) : 0;
        const int frstep = (miy != ry_) ? (mix - rx_) / (miy - ry_) : 0;
        const int srstep = (ry_ != tx_) ? (rx_ - tx_) / (ry_ - tx_) : 0;
        const int y0 = max(miy, 0), y1 = min(may, H - 1);
        int total = 0, alg = 0;
        const double theta = rec.theta, prec = rec.prec;
        const int rows = y1 - y0 + 1;
        if (rows >= 16) {  // one lane per row
            for (int y = y0 + lane; y <= y1; y += 32) {
                const long long c = y - y0;
                const long long cfl = min(max((long long)ly_ - y0, 0ll), c), cfr = min(max((long long)ry_ - y0, 0ll), c);
                const int xl = max((int)(mix + cfl * flstep + (c - cfl) * slstep), 0);
                const int xr = min((int)(mix + cfr * frstep + (c - cfr) * srstep), W - 1);
                if (xr >= xl) total += xr - xl + 1;
                const float* row = deg + (size_t)y * W;
#pragma unroll 4
                for (int x = xl; x <= xr; x++) {
                    const float a = row[x];
                    if (a != LSD_NOTDEF_DEG && l_aligned_deg(a, theta, prec)) ++alg;
                }
            }
        } else {
            for (int y = y0; y <= y1; y++) {
                const long long c = y - y0;
                const long long cfl = min(max((long long)ly_ - y0, 0ll), c), cfr = min(max((long long)ry_ - y0, 0ll), c);
                const int xl = max((int)(mix + cfl * flstep + (c - cfl) * slstep), 0);
                const int xr = min((int)(mix + cfr * frstep + (c - cfr) * srstep), W - 1);
                if (lane == 0 && xr >= xl) total += xr - xl + 1;
                const float* row = deg + (size_t)y * W;
                for (int x = xl + lane; x <= xr; x += 32) {
                    const float a = row[x];
                    if (a != LSD_NOTDEF_DEG && l_aligned_deg(a, theta, prec)) ++alg;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            total += __shfl_xor_sync(0xffffffffu, total, o);
            alg += __shfl_xor_sync(0xffffffffu, alg, o);
        }
        return l_nfa(total, alg, rec.p, LOG_NT);
    }

    // lsd.cpp:873-975 as one loop (a single rect_nfa call site keeps the code small): trial 0 is the rectangle itself, then five
    // stages of five trials each -- finer precision, narrower, one side in, the other side in, finer precision again.
    __device__ __forceinline__ double rect_improve(LRect& rec) const {
        const double LOG_EPS = 0, delta = 0.5, delta_2 = delta / 2.0;
        double log_nfa = -DBL_MAX;
        LRect r = rec;
        for (int t = 0; t <= 25; ++t) {
            const int stage = t == 0 ? -1 : (t - 1) / 5;
            if (t > 0 && (t - 1) % 5 == 0) {  // a new stage starts from the best rectangle so far, unless it is already meaningful
                if (log_nfa > LOG_EPS) return log_nfa;
                r = rec;
            }
            if (stage == 0) {
                r.p /= 2;
                r.prec = r.p * LSD_PI;
            } else if (stage > 0) {
                if (!((r.width - delta) >= 0.5)) continue;
                if (stage == 2) {
                    r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2; r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
                } else if (stage == 3) {
                    r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2; r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
                }
                if (stage == 4) {
                    r.p /= 2;
                    r.prec = r.p * LSD_PI;
                } else {
                    r.width -= delta;
                }
            }
            const double v = rect_nfa(r);
            if (t == 0) log_nfa = v;
            else if (v > log_nfa) { log_nfa = v; rec = r; }
        }
        return log_nfa;
    }

    // rect_improve (lsd.cpp:500-504), the +0.5 / rescale (:509-519), the LSDDetector / line_lbd filters, staging
    __device__ __noinline__ void finish(LRect rec, int seed, int owner) const {
        const double log_nfa = rect_improve(rec);
        if (log_nfa <= 0) return;
        rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
        rec.x1 /= LSD_SCALE; rec.y1 /= LSD_SCALE; rec.x2 /= LSD_SCALE; rec.y2 /= LSD_SCALE;
        float e0 = float(rec.x1), e1 = float(rec.y1), e2 = float(rec.x2), e3 = float(rec.y2);
        if (filter) {
            // LSDDetector.cpp:80-101, 219-232 (octaveScale = 1), line_lbd_allclass.cpp:206
            const int w_ = w, h_ = h;
            if (e0 < 0) e0 = 0;
            if (e0 >= w_) e0 = (float)w_ - 1.0f;
            if (e2 < 0) e2 = 0;
            if (e2 >= w_) e2 = (float)w_ - 1.0f;
            if (e1 < 0) e1 = 0;
            if (e1 >= h_) e1 = (float)h_ - 1.0f;
            if (e3 < 0) e3 = 0;
            if (e3 >= h_) e3 = (float)h_ - 1.0f;
            const float thr = 10;
            if (((e0 < thr) && (e2 < thr)) || ((e0 > w_ - thr) && (e2 > w_ - thr)) || ((e1 < thr) && (e3 < thr)) || ((e1 > h_ - thr) && (e3 > h_ - thr)))
                return;
            const double ddx = double(e0 - e2), ddy = double(e1 - e3);
            const float len = (float)sqrt(ddx * ddx + ddy * ddy);
            if (!(len > length_thres)) return;
        }
        if (lane == 0) {
            const int slot = atomicAdd(nout, 1);
            if (slot < stage_cap) {
                stage[slot] = make_float4(e0, e1, e2, e3);
                stage_key[slot] = seed;
                stage_owner[slot] = owner;
            }
        }
    }
};

constexpr int LSD_JOBQ = 32;  // finished rectangles waiting for their NFA search
struct NfaJob {
    double r[12];
    int seed, owner;
};

// Warps per frame.  Four (four frames per SM) measured best at every batch size: 15.5 ms vs 16.5 ms with eight at 256 frames (the
// extra warps mostly poll the rectangle queue), and twice the residency for large batches; two was slower (36 vs 33 ms at 1024 frames).
constexpr int LSD_WARPS_MAX = 4;
constexpr int LSD_WARP_SMEM = 96 * sizeof(double) + LSD_RING * sizeof(uint32_t);  // per warp: staging of the ordered sums + queue ring
constexpr int LSD_MAX_ROUNDS = 12;  // merge rounds before the rest of the frame is redone as one unit

template <int LSD_WARPS>
__global__ void __launch_bounds__(LSD_WARPS * 32, 512 / (LSD_WARPS * 32)) k_lsd_grow(LsdBuffers B, LsdDims d, LsdConst C) {
    extern __shared__ __align__(16) unsigned char lsd_smem[];
    __shared__ int s_next, s_nout, s_nviol, s_nunits, s_alloc, s_qtail, s_qhead, s_qdone, s_active;
    __shared__ NfaJob s_job[LSD_JOBQ];
    __shared__ int s_job_ready[LSD_JOBQ];
    const int f = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npx = d.W * d.H, nw = (npx + 31) / 32;
    Grow G;
    G.sc = reinterpret_cast<double*>(lsd_smem + (size_t)warp * LSD_WARP_SMEM);
    G.ring = (uint32_t)__cvta_generic_to_shared(G.sc + 96);
    uint32_t* Ubits = reinterpret_cast<uint32_t*>(lsd_smem + (size_t)LSD_WARPS * LSD_WARP_SMEM);
    G.U = (uint32_t)__cvta_generic_to_shared(Ubits);
    const size_t fo = (size_t)f * npx;
    G.pix = B.pix + fo;
    G.deg = B.deg + fo;
    G.mg = B.modgrad + fo;
    int* cgrp = B.cgrp + fo;
    G.cgrp = cgrp;
    G.W = d.W; G.H = d.H; G.lane = lane;
    G.LOG_NT = C.log_nt;
    G.n_regions = 0; G.n_px = 0;
    for (int i = threadIdx.x; i < nw; i += blockDim.x) Ubits[i] = 0;
    const int n_big = B.ncomp[4 * f], n_small = B.ncomp[4 * f + 1];
    if (threadIdx.x == 0) { s_next = 0; s_nout = 0; s_nviol = 0; s_alloc = 0; s_nunits = n_big + n_small; s_qtail = 0; s_qhead = 0; s_qdone = 0; s_active = LSD_WARPS; }
    if (threadIdx.x < LSD_JOBQ) s_job_ready[threadIdx.x] = 0;
    __syncthreads();
    const int* L = B.label + fo;
    int *csize = B.csize + fo, *cminx = B.cminx + fo, *cmaxx = B.cmaxx + fo, *cmaxy = B.cmaxy + fo, *cmark = B.cmark + fo;
    const int* units = B.units + fo;
    int* units2 = B.units2 + fo;
    int* viol = B.viol + 2 * fo;
    float4* stage = B.stage + (size_t)f * B.stage_cap;
    int* stage_key = B.stage_key + (size_t)f * B.stage_cap;
    int* stage_owner = B.stage_owner + (size_t)f * B.stage_cap;
    NfaCtx N;
    N.deg = G.deg; N.W = d.W; N.H = d.H; N.w = d.w; N.h = d.h; N.lane = lane; N.LOG_NT = C.log_nt; N.filter = C.filter; N.length_thres = C.length_thres;
    N.stage = stage; N.stage_key = stage_key; N.stage_owner = stage_owner; N.stage_cap = B.stage_cap; N.nout = &s_nout;
    long long cyc[4] = {0, 0, 0, 0};  // region_grow, region2rect, refine (with re-growing), rect_improve
    const long long t_begin = clock64();
    int round = 0;
    for (;; round++) {
        const int n_units = s_nunits;
        const int mode = round == 0 ? 0 : (round <= LSD_MAX_ROUNDS ? 1 : 2);
        uint32_t *regX = round == 0 ? B.regA : B.regB, *tmpX = round == 0 ? B.tmpA : B.tmpB, *lstX = round == 0 ? B.lstA : B.lstB;
        while (true) {
            int ci = 0;
            if (lane == 0) ci = atomicAdd(&s_next, 1);
            ci = __shfl_sync(0xffffffffu, ci, 0);
            if (ci >= n_units) break;
            const int root = mode == 0 ? units[ci < n_big ? ci : npx - 1 - (ci - n_big)] : (mode == 1 ? units2[ci] : 0);
            int usize, minx, maxx, y0, maxy;
            if (mode == 2) { usize = npx; minx = 0; maxx = d.W - 1; y0 = 0; maxy = d.H - 1; }
            else { usize = csize[root]; minx = cminx[root]; maxx = cmaxx[root]; y0 = root / d.W; maxy = cmaxy[root]; }
            int off = 0;
            if (lane == 0) off = atomicAdd(&s_alloc, usize);
            off = __shfl_sync(0xffffffffu, off, 0);
            G.reg = regX + fo + off;
            G.tmp = tmpX + fo + off;
            uint32_t* lst = lstX + fo + off;
            G.leader = root;
            G.mode = mode;
            G.foreign_root = -1;
            // (1) the unit's pixels in raster order: scan its bounding box for the label.  A merged unit starts from a clean used map.
            int cnt = 0;
            for (int y = y0; y <= maxy && cnt < usize; y++)
                for (int xb = minx & ~31; xb <= maxx; xb += 128) {
                    int lbl[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {  // four independent loads in flight
                        const int x = xb + 32 * k + lane;
                        lbl[k] = x <= maxx ? L[y * d.W + x] : -1;
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const bool mine = mode == 0 ? lbl[k] == root : (lbl[k] >= 0 && (mode == 2 || lbl[k] == root || ccl_find(cgrp, lbl[k]) == root));
                        const unsigned m = __ballot_sync(0xffffffffu, mine);
                        if (mine) {
                            const int pi = y * d.W + xb + 32 * k + lane;
                            lst[cnt + __popc(m & ((1u << lane) - 1u))] = (uint32_t)pi;
                            if (mode != 0) G.clear_used(pi);
                        }
                        cnt += __popc(m);
                    }
                }
            __syncwarp();
            // (2) flsd's seed loop (lsd.cpp:474-535) over this unit
            for (int base = 0; base < cnt && G.foreign_root < 0; base += 32) {
                const int idx = base + lane;
                const int q = idx < cnt ? (int)lst[idx] : 0;
                unsigned todo = 0xffffffffu;
                while (G.foreign_root < 0) {
                    const bool unused = idx < cnt && !G.used(q);
                    const unsigned m = __ballot_sync(0xffffffffu, unused) & todo;
                    if (!m) break;
                    const int l = __ffs(m) - 1;
                    todo = ~((2u << l) - 1u);
                    const int seed = __shfl_sync(0xffffffffu, q, l);
                    // ---- one seed (lsd.cpp:478-534 with refine :786-832 and reduce_region_radius :834-871 unrolled into one loop, so
                    //      that region_grow and region2rect have a single call site each)
                    LRect rec;
                    double reg_angle = 0, tol = C.prec, radSq = 0, xc = 0, yc = 0;
                    int n = 0, attempt = 0;
                    bool reducing = false, good = false;
                    long long t0 = clock64();
                    while (true) {
                        if (!reducing) {
                            n = G.region_grow(seed, tol, reg_angle);
                            const long long t1 = clock64();
                            cyc[attempt == 0 ? 0 : 2] += t1 - t0;
                            t0 = t1;
                            if (n < 0) break;  // a foreign pixel: the unit ends here
                            if (attempt == 0) G.n_regions++;
                            G.n_px += n;
                            if (n < (attempt == 0 ? C.min_reg_size : 2)) break;
                        } else {
                            radSq *= 0.75 * 0.75;
                            n = G.reduce_step(n, xc, yc, radSq);
                            if (n < 2) break;
                        }
                        G.region2rect(n, reg_angle, C.prec, C.p, rec);
                        const double density = double(n) / (l_dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
                        long long t1 = clock64();
                        cyc[attempt == 0 ? 1 : 2] += t1 - t0;
                        t0 = t1;
                        if (density >= 0.7) { good = true; break; }
                        if (attempt == 0) {  // refine(): try a tighter angle tolerance first
                            tol = G.tau_step(n, rec.width);
                            attempt = 1;
                        } else if (!reducing) {  // then shrink the region around the seed
                            xc = double(seed % d.W); yc = double(seed / d.W);
                            const double radSq1 = l_distSq(xc, yc, rec.x1, rec.y1), radSq2 = l_distSq(xc, yc, rec.x2, rec.y2);
                            radSq = radSq1 > radSq2 ? radSq1 : radSq2;
                            reducing = true;
                        }
                        t1 = clock64();
                        cyc[2] += t1 - t0;
                        t0 = t1;
                    }
                    if (n < 0) break;
                    if (!good) continue;
                    // ---- hand the rectangle to whichever warp is free (the NFA search does not touch the used map); keep it when the queue is full
                    int ticket = -1;
                    if (lane == 0) {
                        const int tl = *(volatile int*)&s_qtail, dn = *(volatile int*)&s_qdone;
                        if (tl - dn < LSD_JOBQ - 2 * LSD_WARPS_MAX) ticket = atomicAdd(&s_qtail, 1);
                    }
                    ticket = __shfl_sync(0xffffffffu, ticket, 0);
                    if (ticket >= 0) {
                        NfaJob& jb = s_job[ticket % LSD_JOBQ];
                        const double rv[12] = {rec.x1, rec.y1, rec.x2, rec.y2, rec.width, rec.x, rec.y, rec.theta, rec.dx, rec.dy, rec.prec, rec.p};
                        if (lane < 12) jb.r[lane] = rv[lane];
                        if (lane == 12) { jb.seed = seed; jb.owner = root; }
                        __syncwarp();
                        if (lane == 0) {
                            __threadfence_block();
                            *(volatile int*)&s_job_ready[ticket % LSD_JOBQ] = ticket + 1;
                        }
                    } else {
                        t0 = clock64();
                        N.finish(rec, seed, root);
                        cyc[3] += clock64() - t0;
                    }
                }
                __syncwarp();
            }
            if (G.foreign_root >= 0 && lane == 0) {  // one pair per unit and round
                const int k = atomicAdd(&s_nviol, 1);
                viol[2 * k] = root;
                viol[2 * k + 1] = G.foreign_root;
            }
        }
        // ---- no unit left for this warp: serve the rectangle queue until every warp is here and the queue is empty
        if (lane == 0) atomicSub(&s_active, 1);
        while (true) {
            int h = -1;
            if (lane == 0) {
                const int act = *(volatile int*)&s_active;  // read before the queue: a producer pushes before it retires
                const int hd = *(volatile int*)&s_qhead, tl = *(volatile int*)&s_qtail;
                if (hd < tl) h = atomicCAS(&s_qhead, hd, hd + 1) == hd ? hd : -2;
                else if (act == 0) h = -3;
            }
            h = __shfl_sync(0xffffffffu, h, 0);
            if (h == -3) break;
            if (h < 0) { __nanosleep(200); continue; }
            const int slot = h % LSD_JOBQ;
            if (lane == 0) while (*(volatile int*)&s_job_ready[slot] != h + 1) __nanosleep(50);
            __syncwarp();
            const NfaJob& jb = s_job[slot];
            LRect rec;
            rec.x1 = jb.r[0]; rec.y1 = jb.r[1]; rec.x2 = jb.r[2]; rec.y2 = jb.r[3]; rec.width = jb.r[4]; rec.x = jb.r[5]; rec.y = jb.r[6];
            rec.theta = jb.r[7]; rec.dx = jb.r[8]; rec.dy = jb.r[9]; rec.prec = jb.r[10]; rec.p = jb.r[11];
            const int jseed = jb.seed, jowner = jb.owner;
            __syncwarp();
            if (lane == 0) atomicAdd(&s_qdone, 1);
            const long long t0 = clock64();
            N.finish(rec, jseed, jowner);
            cyc[3] += clock64() - t0;
        }
        __syncthreads();
        const int nv = s_nviol;
        __syncthreads();  // everybody has read the count before thread 0 resets it
        if (nv == 0) break;
        // ---- merge the interacting units (one thread: the lists are short), then redo the merged ones in the next round
        if (threadIdx.x == 0) {
            int k = 0;
            if (round + 1 > LSD_MAX_ROUNDS) {
                k = 1;  // last resort: the whole frame as one unit, processed sequentially by one warp
            } else {
                for (int v = 0; v < nv; v++) {
                    int ra = ccl_find(cgrp, viol[2 * v]), rb = ccl_find(cgrp, viol[2 * v + 1]);
                    if (ra == rb) continue;
                    if (ra > rb) { const int t = ra; ra = rb; rb = t; }
                    cgrp[rb] = ra;
                    csize[ra] += csize[rb];
                    cminx[ra] = min(cminx[ra], cminx[rb]);
                    cmaxx[ra] = max(cmaxx[ra], cmaxx[rb]);
                    cmaxy[ra] = max(cmaxy[ra], cmaxy[rb]);
                }
                for (int v = 0; v < nv; v++) {
                    const int r = ccl_find(cgrp, viol[2 * v]);
                    if (cmark[r] != round + 1) {
                        cmark[r] = round + 1;
                        units2[k++] = r;
                    }
                }
            }
            s_nunits = k;
            s_next = 0;
            s_nviol = 0;
            s_alloc = 0;
            s_active = LSD_WARPS;
            atomicAdd(&B.stats[7], 1ull);
            atomicAdd(&B.stats[8], (unsigned long long)nv);
        }
        __syncthreads();
        // segments of units that are being redone are dropped
        const int n_st = min(s_nout, B.stage_cap);
        const bool all = round + 1 > LSD_MAX_ROUNDS;
        for (int i = threadIdx.x; i < n_st; i += blockDim.x)
            if (stage_key[i] >= 0 && (all || cmark[ccl_find(cgrp, stage_owner[i])] == round + 1)) stage_key[i] = -1;
        __syncthreads();
    }
    if (lane == 0) {
        atomicAdd(&B.stats[0], G.n_regions);
        atomicAdd(&B.stats[1], G.n_px);
        for (int i = 0; i < 4; i++) atomicAdd(&B.stats[2 + i], (unsigned long long)cyc[i]);
        atomicAdd(&B.stats[6], (unsigned long long)(clock64() - t_begin));
    }
    // (3) the reference pushes segments in seed order (lsd.cpp:523): rank every kept segment by its seed's pixel index
    const int n_raw = s_nout, n_st = min(n_raw, B.stage_cap);
    float4* lines = reinterpret_cast<float4*>(B.lines) + (size_t)f * C.max_lines;
    __shared__ int s_kept;
    if (threadIdx.x == 0) s_kept = 0;
    __syncthreads();
    int kept = 0;
    for (int i = threadIdx.x; i < n_st; i += blockDim.x) {
        const int key = stage_key[i];
        if (key < 0) continue;
        int rank = 0;
        for (int j = 0; j < n_st; j++) {
            const int kj = stage_key[j];
            rank += kj >= 0 && kj < key;
        }
        if (rank < C.max_lines) lines[rank] = stage[i];
        kept++;
    }
    if (kept) atomicAdd(&s_kept, kept);
    __syncthreads();
    if (threadIdx.x == 0) B.n_lines[f] = n_raw > B.stage_cap ? 0x7fffffff : s_kept;  // staging overflow is reported as a capacity error
}

// ---------------------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------------------
struct LsdState {
    bool uploaded = false, ran = false, timed_last = false;
    LsdDims d{};
    LsdConst C{};
    csb_lsd_params params{};
    DevBuf d_gray, d_scaled, d_pix, d_deg, d_mg, d_arena, d_roots, d_ncomp, d_stage, d_stage_key, d_stage_owner, d_lines, d_nlines, d_stats;
    int stage_cap = 0;
    HostBuf h_gray, h_out;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    size_t grow_smem = 0;
    int64_t h2d_bytes = 0, d2h_bytes = 0;
    int launches_last = 0;
};

void lsd_release(LsdState*& s) {
    if (!s) return;
    DevBuf* bufs[] = {&s->d_gray, &s->d_scaled, &s->d_pix, &s->d_deg, &s->d_mg, &s->d_arena, &s->d_roots, &s->d_ncomp, &s->d_stage,
                      &s->d_stage_key, &s->d_stage_owner, &s->d_lines, &s->d_nlines, &s->d_stats};
    for (DevBuf* b : bufs) b->release();
    s->h_gray.release();
    s->h_out.release();
    for (auto& e : s->ev)
        if (e) cudaEventDestroy(e);
    delete s;
    s = nullptr;
}

bool lsd_view(LsdState* s, LsdView& v) {
    if (!s || !s->ran) return false;
    v.gray = s->d_gray.as<uint8_t>();
    v.lines = s->d_lines.as<float>();
    v.n_lines = s->d_nlines.as<int>();
    v.w = s->d.w; v.h = s->d.h; v.n_frames = s->d.n_frames; v.max_lines = s->params.max_lines;
    return true;
}

static LsdBuffers lsd_buffers(LsdState& s) {
    LsdBuffers B{};
    B.gray = s.d_gray.as<uint8_t>();
    B.scaled = s.d_scaled.as<double>();
    B.pix = s.d_pix.as<float4>();
    B.deg = s.d_deg.as<float>();
    B.modgrad = s.d_mg.as<double>();
    const size_t npx = (size_t)s.d.W * s.d.H * s.d.n_frames;
    uint32_t* ar = s.d_arena.as<uint32_t>();  // 6 work arenas
    B.regA = ar; B.tmpA = ar + npx; B.lstA = ar + 2 * npx; B.regB = ar + 3 * npx; B.tmpB = ar + 4 * npx; B.lstB = ar + 5 * npx;
    int* ro = s.d_roots.as<int>();            // 12 per-pixel int arrays (viol counts twice)
    B.label = ro; B.csize = ro + npx; B.cminx = ro + 2 * npx; B.cmaxx = ro + 3 * npx; B.cmaxy = ro + 4 * npx; B.cflag = ro + 5 * npx;
    B.cgrp = ro + 6 * npx; B.cmark = ro + 7 * npx; B.units = ro + 8 * npx; B.units2 = ro + 9 * npx; B.viol = ro + 10 * npx;
    B.ncomp = s.d_ncomp.as<int>();
    B.stage = s.d_stage.as<float4>();
    B.stage_key = s.d_stage_key.as<int>();
    B.stage_owner = s.d_stage_owner.as<int>();
    B.stage_cap = s.stage_cap;
    B.lines = s.d_lines.as<float>();
    B.n_lines = s.d_nlines.as<int>();
    B.stats = s.d_stats.as<unsigned long long>();
    return B;
}

}  // namespace csb

using namespace csb;

static int lsd_prepare(csb_context* c, int n_frames, int width, int height, const csb_lsd_params* params) {
    if (!c || !params || n_frames <= 0 || width < 8 || height < 8 || (int64_t)width * height > (1 << 23) || params->max_lines <= 0) return CSB_ERR_INVALID;
    CSB_CUDA(c, cudaSetDevice(c->device));
    if (!c->lsd) {
        c->lsd = new LsdState();
        for (auto& e : c->lsd->ev) CSB_CUDA(c, cudaEventCreate(&e));
    }
    LsdState& s = *c->lsd;
    s.uploaded = false;
    s.ran = false;
    s.params = *params;
    LsdDims& d = s.d;
    d.w = width;
    d.h = height;
    d.W = (int)std::lrint(width * LSD_SCALE);   // cvRound (lsd.cpp:459 -> cv::resize dsize)
    d.H = (int)std::lrint(height * LSD_SCALE);
    d.n_frames = n_frames;
    LsdConst& C = s.C;
    {   // cv::getGaussianKernel(7, sigma, CV_64F) with sigma = SIGMA_SCALE / SCALE (lsd.cpp:453-457); ksize = 1 + 2 ceil(sigma sqrt(2 * 3 ln 10)) = 7
        const double sigma = 0.6 / 0.8, scale2X = -0.5 / (sigma * sigma);
        double sum = 0;
        for (int i = 0; i < 7; i++) {
            const double x = i - 3.0;
            C.k[i] = std::exp(scale2X * x * x);
            sum += C.k[i];
        }
        sum = 1. / sum;
        for (int i = 0; i < 7; i++) C.k[i] *= sum;
    }
    C.inv_scale = 1.0 / LSD_SCALE;
    C.prec = LSD_PI * 22.5 / 180;
    C.p = 22.5 / 180;
    C.rho = 2.0 / std::sin(C.prec);
    C.log_nt = 5 * (std::log10(double(d.W)) + std::log10(double(d.H))) / 2 + std::log10(11.0);
    C.min_reg_size = int(-C.log_nt / std::log10(C.p));
    C.length_thres = params->line_length_thres;
    C.filter = params->filter;
    C.max_lines = params->max_lines;
    s.grow_smem = (size_t)LSD_WARPS_MAX * LSD_WARP_SMEM + (size_t)((d.W * d.H + 31) / 32) * sizeof(uint32_t);
    s.stage_cap = std::max((d.W * d.H) / 4, params->max_lines);
    if (s.grow_smem > (size_t)c->max_smem_optin) {
        c->err = "csb_lsd: frame too large for the shared-memory used/defined bitmaps";
        return CSB_ERR_CAPACITY;
    }
    const size_t npx = (size_t)d.W * d.H * n_frames;
    CSB_CUDA(c, s.d_gray.ensure((size_t)width * height * n_frames));
    CSB_CUDA(c, s.d_scaled.ensure(npx * 8));
    CSB_CUDA(c, s.d_pix.ensure(npx * 16));
    CSB_CUDA(c, s.d_deg.ensure(npx * 4));
    CSB_CUDA(c, s.d_mg.ensure(npx * 8));
    CSB_CUDA(c, s.d_arena.ensure(npx * 4 * 6));
    CSB_CUDA(c, s.d_roots.ensure(npx * 4 * 12));
    CSB_CUDA(c, s.d_ncomp.ensure((size_t)n_frames * 16));
    CSB_CUDA(c, s.d_stage.ensure((size_t)n_frames * s.stage_cap * 16));
    CSB_CUDA(c, s.d_stage_key.ensure((size_t)n_frames * s.stage_cap * 4));
    CSB_CUDA(c, s.d_stage_owner.ensure((size_t)n_frames * s.stage_cap * 4));
    CSB_CUDA(c, s.d_lines.ensure_zeroed((size_t)n_frames * params->max_lines * 16, c->stream));
    CSB_CUDA(c, s.d_nlines.ensure((size_t)n_frames * 4));
    CSB_CUDA(c, s.d_stats.ensure(128));
    CSB_CUDA(c, cudaFuncSetAttribute(k_lsd_grow<LSD_WARPS_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s.grow_smem));
    return CSB_OK;
}

extern "C" {

int csb_lsd_upload(csb_context* c, const uint8_t* gray, int n_frames, int width, int height, const csb_lsd_params* params) {
    if (!gray) return CSB_ERR_INVALID;
    int rc = lsd_prepare(c, n_frames, width, height, params);
    if (rc != CSB_OK) return rc;
    LsdState& s = *c->lsd;
    const size_t bytes = (size_t)width * height * n_frames;
    cudaPointerAttributes pa{};
    const bool pinned = cudaPointerGetAttributes(&pa, gray) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();  // an unregistered pageable pointer is not an error here
    if (pinned) {  // caller-pinned frames go straight to the device (the caller keeps them alive until the next synchronising call)
        CSB_CUDA(c, cudaMemcpyAsync(s.d_gray.p, gray, bytes, cudaMemcpyHostToDevice, c->stream));
    } else {
        CSB_CUDA(c, s.h_gray.ensure(bytes));
        std::memcpy(s.h_gray.p, gray, bytes);
        CSB_CUDA(c, cudaMemcpyAsync(s.d_gray.p, s.h_gray.p, bytes, cudaMemcpyHostToDevice, c->stream));
    }
    s.h2d_bytes = (int64_t)bytes;
    s.uploaded = true;
    return CSB_OK;
}

int csb_lsd_run(csb_context* c, int timed) {
    if (!c || !c->lsd || !c->lsd->uploaded) {
        if (c) c->err = "csb_lsd_run before csb_lsd_upload";
        return CSB_ERR_STATE;
    }
    LsdState& s = *c->lsd;
    CSB_CUDA(c, cudaSetDevice(c->device));
    const LsdDims& d = s.d;
    LsdBuffers B = lsd_buffers(s);
    B.scaled = nullptr;  // the scaled image itself is no input of any later stage: written only on request (csb_lsd_debug_maps)
    cudaStream_t st = c->stream;
    CSB_CUDA(c, cudaMemsetAsync(s.d_stats.p, 0, 128, st));
    if (timed) CSB_CUDA(c, cudaEventRecord(s.ev[0], st));
    k_lsd_maps<<<dim3((d.W + SC_TW - 1) / SC_TW, (d.H + SC_TH - 1) / SC_TH, d.n_frames), SC_THREADS, 0, st>>>(B, d, s.C);
    const dim3 pg((d.W + 31) / 32, (d.H + 7) / 8, d.n_frames), pb(32, 8);
    CSB_CUDA(c, cudaMemsetAsync(s.d_ncomp.p, 0, (size_t)d.n_frames * 16, st));
    k_lsd_merge<<<pg, pb, 0, st>>>(B, d, s.params.unit_link_deg > 0 ? (float)s.params.unit_link_deg : LSD_LINK_DEG);
    k_lsd_flatten<<<pg, pb, 0, st>>>(B, d);
    k_lsd_contact<<<pg, pb, 0, st>>>(B, d);
    k_lsd_units<<<pg, pb, 0, st>>>(B, d, s.C.min_reg_size);
    if (timed) CSB_CUDA(c, cudaEventRecord(s.ev[1], st));
    k_lsd_grow<LSD_WARPS_MAX><<<d.n_frames, 32 * LSD_WARPS_MAX, s.grow_smem, st>>>(B, d, s.C);
    if (timed) CSB_CUDA(c, cudaEventRecord(s.ev[2], st));
    CSB_CUDA(c, cudaGetLastError());
    s.launches_last = 6;
    s.timed_last = timed != 0;
    s.ran = true;
    return CSB_OK;
}

int csb_lsd_download(csb_context* c, float* lines_out, int32_t* n_lines_out, csb_lsd_stats* stats) {
    if (!c || !c->lsd || !c->lsd->ran) {
        if (c) c->err = "csb_lsd_download before csb_lsd_run";
        return CSB_ERR_STATE;
    }
    LsdState& s = *c->lsd;
    CSB_CUDA(c, cudaSetDevice(c->device));
    const size_t lb = (size_t)s.d.n_frames * s.params.max_lines * 16, nb = (size_t)s.d.n_frames * 4;
    CSB_CUDA(c, s.h_out.ensure(lb + nb + 128));
    char* h = s.h_out.as<char>();
    // counts first, then only the rows that are in use (the per-frame capacity is usually far larger than the segment count)
    CSB_CUDA(c, cudaMemcpyAsync(h + lb, s.d_nlines.p, nb, cudaMemcpyDeviceToHost, c->stream));
    CSB_CUDA(c, cudaMemcpyAsync(h + lb + nb, s.d_stats.p, 128, cudaMemcpyDeviceToHost, c->stream));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    int rows = 0;
    for (int f = 0; f < s.d.n_frames; f++) rows = std::max(rows, std::min(reinterpret_cast<const int32_t*>(h + lb)[f], s.params.max_lines));
    const size_t pitch = (size_t)s.params.max_lines * 16;
    if (rows > 0) CSB_CUDA(c, cudaMemcpy2DAsync(h, pitch, s.d_lines.p, pitch, (size_t)rows * 16, s.d.n_frames, cudaMemcpyDeviceToHost, c->stream));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    s.d2h_bytes = (int64_t)((size_t)rows * 16 * s.d.n_frames + nb + 128);
    const int32_t* nl = reinterpret_cast<const int32_t*>(h + lb);
    const unsigned long long* st = reinterpret_cast<const unsigned long long*>(h + lb + nb);
    bool overflow = false;
    int64_t total = 0;
    for (int f = 0; f < s.d.n_frames; f++) {
        overflow |= nl[f] > s.params.max_lines;
        total += std::min(nl[f], s.params.max_lines);
    }
    if (lines_out)
        for (int f = 0; f < s.d.n_frames; f++)
            std::memcpy(reinterpret_cast<char*>(lines_out) + f * pitch, h + f * pitch, (size_t)std::min(nl[f], s.params.max_lines) * 16);
    if (n_lines_out) std::memcpy(n_lines_out, nl, nb);
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->n_lines = total;
        stats->n_regions = (int64_t)st[0];
        stats->n_region_px = (int64_t)st[1];
        stats->h2d_bytes = s.h2d_bytes;
        stats->d2h_bytes = s.d2h_bytes;
        stats->scaled_width = s.d.W;
        stats->scaled_height = s.d.H;
        stats->n_kernel_launches = s.launches_last;
        for (int i = 0; i < 5; i++) stats->grow_cycles[i] = (int64_t)st[2 + i];
        stats->n_merge_rounds = (int64_t)st[7];
        stats->n_unit_conflicts = (int64_t)st[8];
        if (s.timed_last) {
            cudaEventElapsedTime(&stats->gpu_ms_maps, s.ev[0], s.ev[1]);
            cudaEventElapsedTime(&stats->gpu_ms_grow, s.ev[1], s.ev[2]);
        }
    }
    if (overflow) {
        c->err = "csb_lsd: more segments than max_lines in at least one frame";
        return CSB_ERR_CAPACITY;
    }
    return CSB_OK;
}

int csb_lsd_detect_batch(csb_context* c, const uint8_t* gray, int n_frames, int width, int height, const csb_lsd_params* params, float* lines_out,
                         int32_t* n_lines_out, csb_lsd_stats* stats) {
    int rc = csb_lsd_upload(c, gray, n_frames, width, height, params);
    if (rc != CSB_OK) return rc;
    rc = csb_lsd_run(c, stats != nullptr);
    if (rc != CSB_OK) return rc;
    return csb_lsd_download(c, lines_out, n_lines_out, stats);
}

int csb_lsd_debug_maps(csb_context* c, int frame, double* scaled_out, double* modgrad_out, double* angles_out) {
    if (!c || !c->lsd || !c->lsd->ran) return CSB_ERR_STATE;
    LsdState& s = *c->lsd;
    if (frame < 0 || frame >= s.d.n_frames) return CSB_ERR_INVALID;
    CSB_CUDA(c, cudaSetDevice(c->device));
    CSB_CUDA(c, cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)s.d.W * s.d.H, off = n * frame;
    if (scaled_out) {
        // the run does not keep the scaled image: the first stage is run again for it (the frames are still resident; it rewrites the same
        // per-pixel maps, nothing a later download reads)
        LsdBuffers B = lsd_buffers(s);
        k_lsd_maps<<<dim3((s.d.W + SC_TW - 1) / SC_TW, (s.d.H + SC_TH - 1) / SC_TH, s.d.n_frames), SC_THREADS, 0, c->stream>>>(B, s.d, s.C);
        CSB_CUDA(c, cudaGetLastError());
        CSB_CUDA(c, cudaStreamSynchronize(c->stream));
        CSB_CUDA(c, cudaMemcpy(scaled_out, s.d_scaled.as<double>() + off, n * 8, cudaMemcpyDeviceToHost));
    }
    if (modgrad_out) CSB_CUDA(c, cudaMemcpy(modgrad_out, s.d_mg.as<double>() + off, n * 8, cudaMemcpyDeviceToHost));
    if (angles_out) {
        std::vector<float> deg(n);
        CSB_CUDA(c, cudaMemcpy(deg.data(), s.d_deg.as<float>() + off, n * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; i++) angles_out[i] = deg[i] == LSD_NOTDEF_DEG ? -1024.0 : (double)deg[i] * LSD_DEG2RAD;
    }
    return CSB_OK;
}

}  // extern "C"
