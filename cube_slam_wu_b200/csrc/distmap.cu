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
//   k_canny_nms  : per ROI pixel: Sobel at the pixel and at the two neighbours its gradient direction selects -> 0 weak / 1 no / 2 strong
//   k_canny_hyst : per task: breadth-first promotion of weak pixels 8-connected to strong ones (work list in global memory)
//   k_dist3x3    : one warp per task: forward / backward chamfer passes; each row is a min-plus prefix scan
//                  d[j] = min_m (c[m] + a (j - m)) done with warp shuffles, previous row in registers, rows sequential
#include <cuda_runtime.h>

#include <cstdint>

#include "context.h"

namespace csb {

__device__ __forceinline__ int gray_at(const uint8_t* img, int W, int H, int y, int x) {
    y = min(max(y, 0), H - 1);
    x = min(max(x, 0), W - 1);
    return (int)img[(size_t)y * W + x];
}
// Sobel 3x3 (cv::Sobel ksize 3, scale 1) at image position (y, x)
__device__ __forceinline__ void sobel_at(const uint8_t* img, int W, int H, int y, int x, int& dx, int& dy) {
    const int a = gray_at(img, W, H, y - 1, x - 1), b = gray_at(img, W, H, y - 1, x), c = gray_at(img, W, H, y - 1, x + 1);
    const int d = gray_at(img, W, H, y, x - 1), f = gray_at(img, W, H, y, x + 1);
    const int g = gray_at(img, W, H, y + 1, x - 1), h = gray_at(img, W, H, y + 1, x), i = gray_at(img, W, H, y + 1, x + 1);
    dx = (c + 2 * f + i) - (a + 2 * d + g);
    dy = (g + 2 * h + i) - (a + 2 * b + c);
}
// L1 gradient magnitude of ROI pixel (r, c); 0 outside the ROI (the zero border of Canny's magnitude buffer)
__device__ __forceinline__ int mag_at(const uint8_t* img, int W, int H, const TaskTab& t, int r, int c) {
    if (r < 0 || c < 0 || r >= t.roi_h || c >= t.roi_w) return 0;
    int dx, dy;
    sobel_at(img, W, H, t.roi_top + r, t.roi_left + c, dx, dy);
    return abs(dx) + abs(dy);
}

// grid (n_tasks, NMS_Y)
__global__ void __launch_bounds__(256) k_canny_nms(DetectBuffers B, const uint8_t* gray, uint8_t* cmap, int low, int high) {
    const int task = blockIdx.x;
    const TaskTab t = B.ttab[task];
    const FrameTab& ft = B.ftab[t.frame_id];
    const uint8_t* img = gray + ft.gray_offset;
    const int W = ft.img_w, H = ft.img_h;
    const int n = t.roi_w * t.roi_h;
    uint8_t* out = cmap + t.map_offset;
    const int TG22 = (int)(0.4142135623730950488016887242097 * (1 << 15) + 0.5);
    for (int p = blockIdx.y * blockDim.x + threadIdx.x; p < n; p += gridDim.y * blockDim.x) {
        const int r = p / t.roi_w, c = p - r * t.roi_w;
        int xs, ys;
        sobel_at(img, W, H, t.roi_top + r, t.roi_left + c, xs, ys);
        const int m = abs(xs) + abs(ys);
        uint8_t v = 1;
        if (m > low) {
            const int x = abs(xs), y = abs(ys) << 15;
            const int tg22x = x * TG22;
            bool keep;
            if (y < tg22x) keep = (m > mag_at(img, W, H, t, r, c - 1)) && (m >= mag_at(img, W, H, t, r, c + 1));
            else {
                const int tg67x = tg22x + (x << 16);
                if (y > tg67x) keep = (m > mag_at(img, W, H, t, r - 1, c)) && (m >= mag_at(img, W, H, t, r + 1, c));
                else {
                    const int s = ((xs ^ ys) < 0) ? -1 : 1;
                    keep = (m > mag_at(img, W, H, t, r - 1, c - s)) && (m > mag_at(img, W, H, t, r + 1, c + s));
                }
            }
            if (keep) v = (m > high) ? 2 : 0;
        }
        out[p] = v;
    }
}

// one CTA per task
__global__ void __launch_bounds__(256) k_canny_hyst(DetectBuffers B, uint8_t* cmap, int* queue) {
    const int task = blockIdx.x, tid = threadIdx.x;
    const TaskTab t = B.ttab[task];
    const int W = t.roi_w, Hh = t.roi_h, n = W * Hh;
    uint8_t* map = cmap + t.map_offset;  // map_offset is a multiple of 4: 32-bit words are aligned
    int* q = queue + t.map_offset;
    __shared__ int s_tail, s_head, s_end;
    if (tid == 0) { s_tail = 0; s_head = 0; }
    __syncthreads();
    for (int p = tid; p < n; p += blockDim.x)
        if (map[p] == 2) q[atomicAdd(&s_tail, 1)] = p;
    __syncthreads();
    while (true) {
        if (tid == 0) s_end = s_tail;
        __syncthreads();
        const int head = s_head, end = s_end;
        if (head >= end) break;
        for (int e = head + tid; e < end; e += blockDim.x) {
            const int p = q[e];
            const int r = p / W, c = p - r * W;
#pragma unroll
            for (int dr = -1; dr <= 1; dr++)
#pragma unroll
                for (int dc = -1; dc <= 1; dc++) {
                    const int rr = r + dr, cc = c + dc;
                    if ((dr | dc) == 0 || rr < 0 || cc < 0 || rr >= Hh || cc >= W) continue;
                    const int np = rr * W + cc;
                    if (map[np] != 0) continue;
                    unsigned* w = reinterpret_cast<unsigned*>(map + (np & ~3));
                    const int sh = 8 * (np & 3);
                    const unsigned old = atomicOr(w, 2u << sh);
                    if (((old >> sh) & 0xffu) == 0) q[atomicAdd(&s_tail, 1)] = np;
                }
        }
        __syncthreads();
        if (tid == 0) s_head = end;
        __syncthreads();
    }
}

constexpr unsigned DT_HV = 62587u;                 // cvRound(0.955f  * 65536)
constexpr unsigned DT_DG = 89738u;                 // cvRound(1.3693f * 65536)
constexpr unsigned DT_MAX = 0xffffffffu - DT_DG;   // DIST_MAX; also used for the border cells (behaves like OpenCV's INIT_DIST0)
constexpr int DT_WARPS = 4;            // warps (= tasks) per CTA
constexpr int DT_INF = 0x3fffffff;     // "no path yet" inside the kernel (32-bit arithmetic); becomes DT_MAX on output

// One warp per task; lane l owns columns [l*CH, l*CH + CH).  The previous row lives in registers, neighbours' boundary cells come
// by shuffle, each row is a min-plus scan:  forward  d[j] = min_{m<=j} (c[m] - a m) + a j,   backward  d[j] = min_{m>=j} (c[m] + a m) - a j
// (a = DT_HV; c = 0 on edge pixels, else the 3-neighbour minimum over the already finished adjacent row).  No block barriers.
// 32-bit arithmetic: OpenCV saturates unreachable cells at DIST_MAX (~2^32); with at least one edge pixel in the ROI every final
// value is a real path length (< 2^27 for ROIs up to 1280 px), and saturated cells only ever lose comparisons, so any "infinity"
// that survives the additions gives the same result.  A ROI without edge pixels ends at DT_INF everywhere -> DIST_MAX, as in OpenCV.
// The chunk size CH (8 / 16 / 40 columns per lane) is chosen per task from its ROI width; all classes run in one launch so that
// the launch lasts as long as the slowest warp, not the sum of the classes.
template <int CH>
__device__ __forceinline__ void dist3x3_warp(const TaskTab& t, const uint8_t* cmap, int* dtmp, float* maps, int lane) {
    const int W = t.roi_w, H = t.roi_h;
    const uint8_t* map = cmap + t.map_offset;
    int* tmp = dtmp + t.map_offset;
    float* out = maps + t.map_offset;
    const float scale = 1.f / 65536.f;
    const int HV = (int)DT_HV, DG = (int)DT_DG;
    const int j0 = lane * CH;
    int prev[CH];
    // ---- forward pass
#pragma unroll
    for (int k = 0; k < CH; k++) prev[k] = DT_INF;  // row -1 = border
    uint8_t feat[CH], nfeat[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) { const int j = j0 + k; feat[k] = (j < W) ? map[j] : 1; }
    for (int i = 0; i < H; i++) {
#pragma unroll
        for (int k = 0; k < CH; k++) { const int j = j0 + k; nfeat[k] = (i + 1 < H && j < W) ? map[(size_t)(i + 1) * W + j] : 1; }  // prefetch
        const int pl = __shfl_up_sync(0xffffffffu, prev[CH - 1], 1), pr = __shfl_down_sync(0xffffffffu, prev[0], 1);
        const int left_in = (lane == 0) ? DT_INF : pl, right_in = (lane == 31) ? DT_INF : pr;
        int v[CH];
        int run = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int j = j0 + k;
            int x = 0x7fffffff;
            if (j < W) {
                int c;
                if (feat[k] == 2) c = 0;
                else {
                    const int ul = (k > 0) ? prev[k - 1] : left_in;
                    const int ur = (k + 1 < CH) ? ((j + 1 < W) ? prev[k + 1] : DT_INF) : ((j + 1 < W) ? right_in : DT_INF);
                    c = min(min(ul + DG, prev[k] + HV), min(ur + DG, DT_INF));
                }
                x = c - HV * j;
            }
            run = min(run, x);
            v[k] = run;  // inclusive prefix min inside the chunk
        }
        int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = min(inc, y); }
        int excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0x7fffffff;
        excl = min(excl, DT_INF + HV);  // left border cell (column -1)
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int j = j0 + k;
            const int dv = min(min(excl, v[k]) + HV * j, DT_INF);
            prev[k] = (j < W) ? dv : DT_INF;
            if (j < W) tmp[(size_t)i * W + j] = dv;
            feat[k] = nfeat[k];
        }
    }
    // ---- backward pass
#pragma unroll
    for (int k = 0; k < CH; k++) prev[k] = DT_INF;  // row H = border
    int self[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) { const int j = j0 + k; self[k] = (j < W) ? tmp[(size_t)(H - 1) * W + j] : DT_INF; }
    for (int i = H - 1; i >= 0; i--) {
        int nself[CH];  // prefetch the row above while this one is processed
#pragma unroll
        for (int k = 0; k < CH; k++) { const int j = j0 + k; nself[k] = (i > 0 && j < W) ? tmp[(size_t)(i - 1) * W + j] : DT_INF; }
        const int pl = __shfl_up_sync(0xffffffffu, prev[CH - 1], 1), pr = __shfl_down_sync(0xffffffffu, prev[0], 1);
        const int left_in = (lane == 0) ? DT_INF : pl, right_in = (lane == 31) ? DT_INF : pr;
        int v[CH];
        int run = 0x7fffffff;
#pragma unroll
        for (int k = CH - 1; k >= 0; k--) {
            const int j = j0 + k;
            int x = 0x7fffffff;
            if (j < W) {
                const int dl = (k > 0) ? prev[k - 1] : left_in;
                const int dr = (k + 1 < CH) ? ((j + 1 < W) ? prev[k + 1] : DT_INF) : ((j + 1 < W) ? right_in : DT_INF);
                const int t0 = min(min(self[k], dr + DG), min(prev[k] + HV, dl + DG));
                x = t0 + HV * j;
            }
            run = min(run, x);
            v[k] = run;  // inclusive suffix min inside the chunk
        }
        int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_down_sync(0xffffffffu, inc, o); if (lane + o < 32) inc = min(inc, y); }
        int excl = __shfl_down_sync(0xffffffffu, inc, 1);
        if (lane == 31) excl = 0x7fffffff;
        excl = min(excl, DT_INF + HV * W);  // right border cell (column W)
#pragma unroll
        for (int k = 0; k < CH; k++) {
            const int j = j0 + k;
            const int dv = min(min(excl, v[k]) - HV * j, DT_INF);
            prev[k] = (j < W) ? dv : DT_INF;
            if (j < W) out[(size_t)i * W + j] = (float)((dv >= DT_INF) ? DT_MAX : (unsigned)dv) * scale;
            self[k] = nself[k];
        }
    }
}

__global__ void __launch_bounds__(32 * DT_WARPS) k_dist3x3(DetectBuffers B, const uint8_t* cmap, int* dtmp, float* maps) {
    const int task = blockIdx.x * DT_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (task >= B.n_tasks) return;
    const TaskTab t = B.ttab[task];
    if (t.roi_w <= 32 * 8) dist3x3_warp<8>(t, cmap, dtmp, maps, lane);
    else if (t.roi_w <= 32 * 16) dist3x3_warp<16>(t, cmap, dtmp, maps, lane);
    else dist3x3_warp<40>(t, cmap, dtmp, maps, lane);
}

cudaError_t launch_distmaps(const DetectBuffers& B, const uint8_t* gray, uint8_t* cmap, int* queue, unsigned* dtmp, float* maps, int max_roi_w, cudaStream_t st) {
    if (B.n_tasks == 0) return cudaSuccess;
    dim3 g(B.n_tasks, 8);
    k_canny_nms<<<g, 256, 0, st>>>(B, gray, cmap, 80, 200);
    k_canny_hyst<<<B.n_tasks, 256, 0, st>>>(B, cmap, queue);
    const int grid = (B.n_tasks + DT_WARPS - 1) / DT_WARPS;
    if (max_roi_w > 32 * 40) return cudaErrorInvalidValue;  // ROI wider than 1280 px
    k_dist3x3<<<grid, 32 * DT_WARPS, 0, st>>>(B, cmap, reinterpret_cast<int*>(dtmp), maps);
    return cudaGetLastError();
}

}  // namespace csb
