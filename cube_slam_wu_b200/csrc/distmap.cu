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
//   k_dist3x3    : per task: forward / backward chamfer passes; each row is a min-plus prefix scan
//                  d[j] = min_m (c[m] + a (j - m)) done with warp shuffles, rows are sequential
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
constexpr int DT_THREADS = 256;

// block-wide inclusive prefix-min of one int64 per thread; carry = min of everything before this segment
__device__ __forceinline__ long long block_prefix_min(long long v, long long carry, long long* s_agg, int tid, long long& seg_min) {
    const unsigned FULL = 0xffffffffu;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        long long t = __shfl_up_sync(FULL, v, o);
        if (lane >= o && t < v) v = t;
    }
    if (lane == 31) s_agg[warp] = v;
    __syncthreads();
    long long pre = carry, tot = carry;
#pragma unroll
    for (int w = 0; w < DT_THREADS / 32; w++) {
        const long long a = s_agg[w];
        if (w < warp && a < pre) pre = a;
        if (a < tot) tot = a;
    }
    __syncthreads();
    seg_min = tot;
    return (pre < v) ? pre : v;
}

// one CTA per task.  tmp = 32-bit fixed-point distances of the whole ROI (global, L2 resident).
__global__ void __launch_bounds__(DT_THREADS) k_dist3x3(DetectBuffers B, const uint8_t* cmap, unsigned* dtmp, float* maps) {
    const int task = blockIdx.x, tid = threadIdx.x;
    const TaskTab t = B.ttab[task];
    const int W = t.roi_w, H = t.roi_h;
    const uint8_t* map = cmap + t.map_offset;
    unsigned* tmp = dtmp + t.map_offset;
    float* out = maps + t.map_offset;
    __shared__ long long s_agg[DT_THREADS / 32];
    const float scale = 1.f / 65536.f;
    // forward pass: rows top -> bottom, within a row left -> right
    for (int i = 0; i < H; i++) {
        const unsigned* up = tmp + (size_t)(i - 1) * W;
        long long carry = (long long)DT_MAX + DT_HV;  // left border cell at position -1:  INIT - HV * (-1)
        for (int j0 = 0; j0 < W; j0 += DT_THREADS) {
            const int j = j0 + tid;
            long long v = 0x7fffffffffffffffLL;
            if (j < W) {
                unsigned c;
                if (map[(size_t)i * W + j] == 2) c = 0;
                else {
                    const unsigned ul = (i > 0 && j > 0) ? up[j - 1] : DT_MAX, u = (i > 0) ? up[j] : DT_MAX, ur = (i > 0 && j + 1 < W) ? up[j + 1] : DT_MAX;
                    unsigned long long t0 = (unsigned long long)ul + DT_DG, tt = (unsigned long long)u + DT_HV;
                    if (t0 > tt) t0 = tt;
                    tt = (unsigned long long)ur + DT_DG;
                    if (t0 > tt) t0 = tt;
                    c = (t0 > DT_MAX) ? DT_MAX : (unsigned)t0;
                }
                v = (long long)c - (long long)DT_HV * j;
            }
            long long seg;
            const long long pm = block_prefix_min(v, carry, s_agg, tid, seg);
            if (j < W) {
                long long d = pm + (long long)DT_HV * j;
                tmp[(size_t)i * W + j] = (d > (long long)DT_MAX) ? DT_MAX : (unsigned)d;
            }
            carry = seg;
        }
        __syncthreads();  // row i visible before row i+1 reads it
    }
    // backward pass: rows bottom -> top, within a row right -> left (scan over q = W-1-j)
    for (int i = H - 1; i >= 0; i--) {
        const unsigned* dn = tmp + (size_t)(i + 1) * W;
        long long carry = (long long)DT_MAX + DT_HV;
        for (int q0 = 0; q0 < W; q0 += DT_THREADS) {
            const int q = q0 + tid, j = W - 1 - q;
            long long v = 0x7fffffffffffffffLL;
            if (q < W) {
                const unsigned self = tmp[(size_t)i * W + j];
                const unsigned dr = (i + 1 < H && j + 1 < W) ? dn[j + 1] : DT_MAX, d = (i + 1 < H) ? dn[j] : DT_MAX, dl = (i + 1 < H && j > 0) ? dn[j - 1] : DT_MAX;
                unsigned long long t0 = self, tt = (unsigned long long)dr + DT_DG;
                if (t0 > tt) t0 = tt;
                tt = (unsigned long long)d + DT_HV;
                if (t0 > tt) t0 = tt;
                tt = (unsigned long long)dl + DT_DG;
                if (t0 > tt) t0 = tt;
                v = (long long)t0 - (long long)DT_HV * q;
            }
            long long seg;
            const long long pm = block_prefix_min(v, carry, s_agg, tid, seg);
            if (q < W) {
                long long d = pm + (long long)DT_HV * q;
                const unsigned t0 = (d > (long long)DT_MAX) ? DT_MAX : (unsigned)d;
                tmp[(size_t)i * W + j] = t0;
                out[(size_t)i * W + j] = (float)t0 * scale;
            }
            carry = seg;
        }
        __syncthreads();
    }
}

cudaError_t launch_distmaps(const DetectBuffers& B, const uint8_t* gray, uint8_t* cmap, int* queue, unsigned* dtmp, float* maps, cudaStream_t st) {
    if (B.n_tasks == 0) return cudaSuccess;
    dim3 g(B.n_tasks, 8);
    k_canny_nms<<<g, 256, 0, st>>>(B, gray, cmap, 80, 200);
    k_canny_hyst<<<B.n_tasks, 256, 0, st>>>(B, cmap, queue);
    k_dist3x3<<<B.n_tasks, DT_THREADS, 0, st>>>(B, cmap, dtmp, maps);
    return cudaGetLastError();
}

}  // namespace csb
