// proposal.cu -- sm_100a kernels of the proposal half (detect_3d_cuboid::detect_cuboid hot path).
//
//   k_prep_lines : per task (one warp): ROI line filter, merge_break_lines, length filter, angle / midpoint tables
//   k_score      : per task (one persistent CTA per SM): distance map staged in shared memory by a TMA bulk copy
//                  (cp.async.bulk + mbarrier), VP-support tables, hypothesis sweep -> ordered compaction ->
//                  dense scoring (distance-map gathers + edge-angle error)
//   k_select     : per task: fuse_normalize_scores_v2 (2/3 selection by both scores with std::partial_sort
//                  semantics), 3D recovery of the kept proposals, skew-augmented score
//   k_rank       : per 2D box: final top-k over the height samples, cuboid records
//
// Compiled with -fmad=false: FP64 results must be the plain IEEE sequence (see csb_math.cuh).
#include <cuda_runtime.h>

#include <cstdint>

#include "proposal.h"
#include "proposal_dev.cuh"

namespace csb {

// ------------------------------------------------------------------------------------------------
// small PTX wrappers: mbarrier + 1-D TMA bulk copy global -> shared
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// k_prep_lines
// ------------------------------------------------------------------------------------------------
// Merge predicate of merge_break_lines (object_3d_util.cpp:463-497) for rows a < b.  On success returns the merged
// segment and its angle.
struct LineSM {
    double *x1, *y1, *x2, *y2, *ang;
};
__device__ __forceinline__ bool merge_pred(const LineSM& L, int a, int b, double thr_ang, double thr_dist, double* out /*x1 y1 x2 y2 ang*/) {
    double diff = fabs(L.ang[a] - L.ang[b]);
    double angle_diff = cmin(diff, M_PI - diff);
    if (!(angle_diff < thr_ang)) return false;
    double d12 = norm2(V2{L.x2[a] - L.x1[b], L.y2[a] - L.y1[b]});
    double d21 = norm2(V2{L.x2[b] - L.x1[a], L.y2[b] - L.y1[a]});
    if (!((d12 < thr_dist) || (d21 < thr_dist))) return false;
    double msx, msy, mex, mey;
    if (L.x1[a] < L.x1[b]) { msx = L.x1[a]; msy = L.y1[a]; } else { msx = L.x1[b]; msy = L.y1[b]; }
    if (L.x2[a] > L.x2[b]) { mex = L.x2[a]; mey = L.y2[a]; } else { mex = L.x2[b]; mey = L.y2[b]; }
    double merged_angle = atan2(mey - msy, mex - msx);
    double temp = fabs(L.ang[a] - merged_angle);
    double merge_angle_diff = cmin(temp, M_PI - temp);
    if (!(merge_angle_diff < thr_ang)) return false;
    out[0] = msx; out[1] = msy; out[2] = mex; out[3] = mey; out[4] = merged_angle;
    return true;
}

// One warp per task.  Dynamic shared memory: 5 * cap doubles.
__global__ void __launch_bounds__(32) k_prep_lines(DetectBuffers B, int cap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int task = blockIdx.x;
    const int lane = threadIdx.x;
    const TaskTab& tt = B.ttab[task];
    const FrameTab& ft = B.ftab[tt.frame_id];
    LineSM L;
    L.x1 = reinterpret_cast<double*>(smem_raw);
    L.y1 = L.x1 + cap; L.x2 = L.y1 + cap; L.y2 = L.x2 + cap; L.ang = L.y2 + cap;
    const unsigned FULL = 0xffffffffu;

    // (1) align left->right (object_3d_util.cpp:246-258) and keep lines with both endpoints inside the expanded ROI
    //     (box_proposal_detail.cpp:271-283); order-preserving compaction
    const double rl = (double)tt.roi_left, rt = (double)tt.roi_top, rr = (double)tt.roi_right, rd = (double)tt.roi_down;
    int total = 0;
    const int n_raw = ft.line_end - ft.line_begin;
    for (int base = 0; base < n_raw; base += 32) {
        int i = base + lane;
        bool in = false;
        double a = 0, b = 0, c = 0, d = 0;
        if (i < n_raw) {
            const double* p = B.lines + 4 * (size_t)(ft.line_begin + i);
            a = p[0]; b = p[1]; c = p[2]; d = p[3];
            if (c < a) { double t0 = a, t1 = b; a = c; b = d; c = t0; d = t1; }
            in = inside_box(V2{a, b}, rl, rt, rr, rd) && inside_box(V2{c, d}, rl, rt, rr, rd);
        }
        unsigned bal = __ballot_sync(FULL, in);
        if (in) {
            int pos = total + __popc(bal & ((1u << lane) - 1));
            L.x1[pos] = a; L.y1[pos] = b; L.x2[pos] = c; L.y2[pos] = d;
            L.ang[pos] = atan2(d - b, c - a);
        }
        total += __popc(bal);
    }
    __syncwarp();

    // (2) merge_break_lines (object_3d_util.cpp:431-511): repeatedly merge the first (seg1, seg2) pair in lexicographic
    // order that passes the angle / gap / merged-angle tests, then restart.  The restart only needs to revisit pairs
    // whose rows changed: after a merge at (m, s2), rows < m failed against every row except the new content of row m
    // (the row moved into slot s2 already failed against them), so the next pass tests (a, m) for a < m, then scans
    // seg1 = m, m+1, ... in full.  Angles are cached per row (atan2 of unchanged endpoints is unchanged).
    const double thr_ang = 5.0 / 180.0 * M_PI, thr_dist = 20.0;
    bool can_force_merge = true;
    int counter = 0;
    int marker = -1;  // seg1 of the previous merge; -1: nothing known yet
    while (can_force_merge && counter < 500) {
        counter++;
        can_force_merge = false;
        int hit_a = -1, hit_b = -1;
        double mg[5];
        if (marker >= 0) {
            for (int base = 0; base < marker && hit_a < 0; base += 32) {
                int a = base + lane;
                double o[5];
                bool ok = (a < marker) && merge_pred(L, a, marker, thr_ang, thr_dist, o);
                unsigned bal = __ballot_sync(FULL, ok);
                if (bal) {
                    int src = __ffs(bal) - 1;
                    hit_a = base + src; hit_b = marker;
                    for (int k = 0; k < 5; k++) mg[k] = __shfl_sync(FULL, o[k], src);
                }
            }
        }
        if (hit_a < 0) {
            for (int s1 = (marker < 0 ? 0 : marker); s1 < total - 1 && hit_a < 0; s1++) {
                for (int base = s1 + 1; base < total && hit_a < 0; base += 32) {
                    int b = base + lane;
                    double o[5];
                    bool ok = (b < total) && merge_pred(L, s1, b, thr_ang, thr_dist, o);
                    unsigned bal = __ballot_sync(FULL, ok);
                    if (bal) {
                        int src = __ffs(bal) - 1;
                        hit_a = s1; hit_b = base + src;
                        for (int k = 0; k < 5; k++) mg[k] = __shfl_sync(FULL, o[k], src);
                    }
                }
            }
        }
        if (hit_a >= 0) {
            if (lane == 0) {
                L.x1[hit_a] = mg[0]; L.y1[hit_a] = mg[1]; L.x2[hit_a] = mg[2]; L.y2[hit_a] = mg[3]; L.ang[hit_a] = mg[4];
                // fast_RemoveRow (matrix_utils.cpp:183-187)
                L.x1[hit_b] = L.x1[total - 1]; L.y1[hit_b] = L.y1[total - 1]; L.x2[hit_b] = L.x2[total - 1]; L.y2[hit_b] = L.y2[total - 1];
                L.ang[hit_b] = L.ang[total - 1];
            }
            total--;
            marker = hit_a;
            can_force_merge = true;
            __syncwarp();
        }
    }

    // (3) length filter > 30 px (object_3d_util.cpp:517-539) + angle / midpoint tables (box_proposal_detail.cpp:309-315)
    const size_t ob = (size_t)tt.line_cap_offset;
    int n_out = 0;
    for (int base = 0; base < total; base += 32) {
        int i = base + lane;
        bool keep = false;
        if (i < total) keep = norm2(V2{L.x2[i] - L.x1[i], L.y2[i] - L.y1[i]}) > 30.0;
        unsigned bal = __ballot_sync(FULL, keep);
        if (keep) {
            size_t pos = ob + n_out + __popc(bal & ((1u << lane) - 1));
            B.ml_seg[4 * pos + 0] = L.x1[i]; B.ml_seg[4 * pos + 1] = L.y1[i]; B.ml_seg[4 * pos + 2] = L.x2[i]; B.ml_seg[4 * pos + 3] = L.y2[i];
            B.ml_ang[pos] = L.ang[i];  // == atan2(y2-y1, x2-x1) of the stored endpoints
            B.ml_mid[2 * pos + 0] = (L.x1[i] + L.x2[i]) / 2;
            B.ml_mid[2 * pos + 1] = (L.y1[i] + L.y2[i]) / 2;
        }
        n_out += __popc(bal);
    }
    if (lane == 0) B.n_merged[task] = n_out;
}

// ------------------------------------------------------------------------------------------------
// k_score
// ------------------------------------------------------------------------------------------------
__constant__ double c_t[11] = {0 / 10.0, 1 / 10.0, 2 / 10.0, 3 / 10.0, 4 / 10.0, 5 / 10.0, 6 / 10.0, 7 / 10.0, 8 / 10.0, 9 / 10.0, 10 / 10.0};
__constant__ double c_1mt[11] = {1 - 0 / 10.0, 1 - 1 / 10.0, 1 - 2 / 10.0, 1 - 3 / 10.0, 1 - 4 / 10.0, 1 - 5 / 10.0,
                                 1 - 6 / 10.0, 1 - 7 / 10.0, 1 - 8 / 10.0, 1 - 9 / 10.0, 1 - 10 / 10.0};

// 11 samples along one box edge, object_3d_util.cpp:642-664.  WEIGHT: 0 none, 1 x1.5 (edges 4,5 of config 2), 2 x2 (edge 6)
template <bool SMEM, int WEIGHT>
__device__ __forceinline__ float edge_samples(float sum_dist, const float* __restrict__ map, int cols, int last, V2 c1, V2 c2) {
#pragma unroll 1
    for (int k = 0; k < 11; k++) {
        double sx = c_t[k] * c1.x + c_1mt[k] * c2.x;
        double sy = c_t[k] * c1.y + c_1mt[k] * c2.y;
        int li = __double2int_rz(sy) * cols + __double2int_rz(sx);
        li = max(0, min(li, last));  // defined behaviour for samples on the ROI's right/bottom bound (reference: UB)
        float d1 = SMEM ? map[li] : __ldg(map + li);
        if (WEIGHT == 1) d1 = (float)((double)d1 * 3.0 / 2.0);
        if (WEIGHT == 2) d1 = (float)((double)d1 * 2.0);
        sum_dist = sum_dist + d1;
    }
    return sum_dist;
}

// box_edge_sum_dists, object_3d_util.cpp:622-667 with the visible-edge tables of box_proposal_detail.cpp:646, 663
template <bool SMEM>
__device__ __forceinline__ double box_edge_sum_dists(const float* __restrict__ map, int rows, int cols, const V2* c, int config_id) {
    const int last = rows * cols - 1;
    float s = 0;
    s = edge_samples<SMEM, 0>(s, map, cols, last, c[0], c[1]);
    s = edge_samples<SMEM, 0>(s, map, cols, last, c[1], c[2]);
    s = edge_samples<SMEM, 0>(s, map, cols, last, c[2], c[3]);
    s = edge_samples<SMEM, 0>(s, map, cols, last, c[3], c[0]);
    if (config_id == 1) {
        s = edge_samples<SMEM, 0>(s, map, cols, last, c[1], c[5]);
        s = edge_samples<SMEM, 0>(s, map, cols, last, c[2], c[4]);
        s = edge_samples<SMEM, 0>(s, map, cols, last, c[3], c[7]);
        s = edge_samples<SMEM, 0>(s, map, cols, last, c[4], c[7]);
        s = edge_samples<SMEM, 0>(s, map, cols, last, c[4], c[5]);
    } else {
        s = edge_samples<SMEM, 1>(s, map, cols, last, c[1], c[5]);
        s = edge_samples<SMEM, 1>(s, map, cols, last, c[2], c[4]);
        s = edge_samples<SMEM, 2>(s, map, cols, last, c[4], c[5]);
    }
    return (double)s;
}

// one box edge against the (<=2) supporting line angles of its VP, object_3d_util.cpp:696-715
__device__ __forceinline__ double edge_angle_diff(V2 a, V2 b, double v0, double v1) {
    double box_edge_angle = normalize_to_pi(atan2(b.y - a.y, b.x - a.x));
    double angle_diff_temp = 100;
    if (!isnan(v0)) {
        double temp = fabs(box_edge_angle - v0);
        temp = cmin(temp, M_PI - temp);
        if (temp < angle_diff_temp) angle_diff_temp = temp;
    }
    if (!isnan(v1)) {
        double temp = fabs(box_edge_angle - v1);
        temp = cmin(temp, M_PI - temp);
        if (temp < angle_diff_temp) angle_diff_temp = temp;
    }
    return angle_diff_temp;
}
// box_edge_alignment_angle_error, object_3d_util.cpp:670-723 with the tables of box_proposal_detail.cpp:651, 665
__device__ __forceinline__ double box_edge_alignment_angle_error(const double* sup /*6*/, const V2* c, int config_id) {
    const double not_found_penalty = 30.0 / 180.0 * M_PI * 2;
    double total = 0;
    // VP 1: edges (1,2) and (8,5) | (3,4)
    if (!isnan(sup[0]) || !isnan(sup[1])) {
        total = total + edge_angle_diff(c[0], c[1], sup[0], sup[1]);
        total = total + (config_id == 1 ? edge_angle_diff(c[7], c[4], sup[0], sup[1]) : edge_angle_diff(c[2], c[3], sup[0], sup[1]));
    } else
        total = total + not_found_penalty;
    // VP 2: edges (4,1) and (5,6)
    if (!isnan(sup[2]) || !isnan(sup[3])) {
        total = total + edge_angle_diff(c[3], c[0], sup[2], sup[3]);
        total = total + edge_angle_diff(c[4], c[5], sup[2], sup[3]);
    } else
        total = total + not_found_penalty;
    // VP 3: edges (4,8)|(3,5) and (2,6)
    if (!isnan(sup[4]) || !isnan(sup[5])) {
        total = total + (config_id == 1 ? edge_angle_diff(c[3], c[7], sup[4], sup[5]) : edge_angle_diff(c[2], c[4], sup[4], sup[5]));
        total = total + edge_angle_diff(c[1], c[5], sup[4], sup[5]);
    } else
        total = total + not_found_penalty;
    return total;
}

// VP_support_edge_infos (object_3d_util.cpp:548-619) for one group, executed by one warp; lanes stride over lines.
__device__ __forceinline__ void vp_support_warp(const double* vp, int n, const double* ang, const double* mid, double* sup_out, int lane) {
    const unsigned FULL = 0xffffffffu;
    for (int vp_id = 0; vp_id < 3; vp_id++) {
        const double thr = (vp_id != 2 ? 15.0 : 10.0) / 180.0 * M_PI;
        const double vx = vp[2 * vp_id], vy = vp[2 * vp_id + 1];
        bool have_base = false;
        double base = 0;
        // lane-local extrema of the smoothed inlier angles; ties keep the lowest line index (Eigen max/minCoeff)
        double vmax = 0, vmin = 0;
        int imax = -1, imin = -1;
        for (int b0 = 0; b0 < n; b0 += 32) {
            int e = b0 + lane;
            bool inl = false;
            double raw = 0;
            if (e < n) {
                raw = atan2(mid[2 * e + 1] - vy, mid[2 * e] - vx);
                double nrm = normalize_to_pi(raw);
                double d = fabs(ang[e] - nrm);
                d = cmin(d, M_PI - d);
                inl = d < thr;
            }
            unsigned bal = __ballot_sync(FULL, inl);
            if (!have_base && bal) {
                base = __shfl_sync(FULL, raw, __ffs(bal) - 1);  // smooth_jump_angles: base = first inlier (:285)
                have_base = true;
            }
            if (inl) {
                double v = raw;
                if ((raw - base) < -M_PI) v = raw + 2 * M_PI;
                else if ((raw - base) > M_PI) v = raw - 2 * M_PI;
                if (imax < 0) { vmax = vmin = v; imax = imin = e; }
                else {
                    if (v > vmax) { vmax = v; imax = e; }
                    if (v < vmin) { vmin = v; imin = e; }
                }
            }
        }
        // warp reduction, first index wins on ties
        for (int off = 16; off > 0; off >>= 1) {
            double ov = __shfl_down_sync(FULL, vmax, off); int oi = __shfl_down_sync(FULL, imax, off);
            if (oi >= 0 && (imax < 0 || ov > vmax || (ov == vmax && oi < imax))) { vmax = ov; imax = oi; }
            ov = __shfl_down_sync(FULL, vmin, off); oi = __shfl_down_sync(FULL, imin, off);
            if (oi >= 0 && (imin < 0 || ov < vmin || (ov == vmin && oi < imin))) { vmin = ov; imin = oi; }
        }
        imax = __shfl_sync(FULL, imax, 0);
        imin = __shfl_sync(FULL, imin, 0);
        if (lane == 0) {
            if (imax >= 0) {
                int low = imax, top = imin;
                if (vp_id > 0) { int t = low; low = top; top = t; }
                sup_out[2 * vp_id] = ang[low];
                sup_out[2 * vp_id + 1] = ang[top];
            } else {
                sup_out[2 * vp_id] = nan("");
                sup_out[2 * vp_id + 1] = nan("");
            }
        }
    }
}

constexpr int SCORE_THREADS = 512;
constexpr int LINE_SMEM_CAP = 256;

struct ScoreSmemLayout {
    int groups_cap;    // groups with tables in shared memory
    int map_cap_floats;
    size_t total_bytes;
};

// Persistent CTA: loops over tasks handed out by an atomic counter (largest first).
__global__ void __launch_bounds__(SCORE_THREADS, 1) k_score(DetectBuffers B, int groups_cap, int map_cap_floats) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [map floats][vp tables 6*G][support tables 6*G][line ang][line mid][queue 2T][warp counts][misc]
    float* s_map = reinterpret_cast<float*>(smem_raw);
    double* s_vp = reinterpret_cast<double*>(smem_raw + (size_t)map_cap_floats * 4);
    double* s_sup = s_vp + 6 * (size_t)groups_cap;
    double* s_lang = s_sup + 6 * (size_t)groups_cap;
    double* s_lmid = s_lang + LINE_SMEM_CAP;
    int* s_queue = reinterpret_cast<int*>(s_lmid + 2 * LINE_SMEM_CAP);
    int* s_wcnt = s_queue + 2 * SCORE_THREADS;
    __shared__ uint64_t s_bar;
    __shared__ int s_task;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned FULL = 0xffffffffu;
    constexpr int NW = SCORE_THREADS / 32;
    constexpr int QCAP = 2 * SCORE_THREADS;

    if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t bar_parity = 0;

    while (true) {
        if (tid == 0) s_task = atomicAdd(B.counters, 1);
        __syncthreads();
        const int slot = s_task;
        if (slot >= B.n_tasks) break;
        const int task = B.task_order[slot];
        const TaskTab tt = B.ttab[task];
        const FrameTab& ft = B.ftab[tt.frame_id];
        const TaskGeo geo = make_geo(tt);
        const int n_groups = ft.n_roll * ft.n_pitch * ft.n_yaw;
        const int n_lines = B.n_merged[task];
        const int map_floats = tt.roi_w * tt.roi_h;
        const bool map_smem = map_floats <= map_cap_floats;
        const float* gmap = B.maps + tt.map_offset;

        // (a) kick off the TMA bulk copy of the distance map; it lands while the VP tables are built
        if (map_smem && tid == 0) {
            uint32_t bytes = ((uint32_t)map_floats * 4u + 15u) & ~15u;
            fence_proxy_async();  // previous task's generic reads of s_map are ordered before the async write
            mbar_expect_tx(&s_bar, bytes);
            tma_bulk_g2s(s_map, gmap, bytes, &s_bar);
        }
        // (b) merged-line tables -> shared memory
        const double* lang = B.ml_ang + tt.line_cap_offset;
        const double* lmid = B.ml_mid + 2 * (size_t)tt.line_cap_offset;
        const bool lines_smem = n_lines <= LINE_SMEM_CAP;
        if (lines_smem) {
            for (int i = tid; i < n_lines; i += SCORE_THREADS) { s_lang[i] = lang[i]; s_lmid[2 * i] = lmid[2 * i]; s_lmid[2 * i + 1] = lmid[2 * i + 1]; }
        }
        __syncthreads();
        // (c) per-group vanishing points and VP-support angles: one warp per group
        for (int g = warp; g < n_groups; g += NW) {
            int yaw_id = g % ft.n_yaw, pair = g / ft.n_yaw;
            double vp[6];
            vanishing_points(ft.KinvR[pair], ft.cosy[yaw_id], ft.siny[yaw_id], vp);
            if (lane == 0) { for (int q = 0; q < 6; q++) s_vp[6 * g + q] = vp[q]; }
            if (n_lines > 0) vp_support_warp(vp, n_lines, lines_smem ? s_lang : lang, lines_smem ? s_lmid : lmid, s_sup + 6 * g, lane);
            else if (lane < 6) s_sup[6 * g + lane] = nan("");
        }
        __syncthreads();

        // (d) sweep.  Phase 1: every thread tests one hypothesis (corner construction + rejection cascade); survivors'
        // ids are appended to a ring buffer in enumeration order.  Phase 2: whenever a full block's worth is queued
        // (or at the end), each thread scores one survivor with all lanes active.
        int q_head = 0, q_count = 0, n_done = 0;
        bool map_ready = !map_smem;
        const int n_hyp = tt.n_hyp;
        for (int base = 0; base < n_hyp || q_count > 0;) {
            if (base < n_hyp) {
                int h = base + tid;
                bool valid = false;
                if (h < n_hyp) {
                    int group, top, cfg;
                    decode_hyp(h, tt.n_top, group, top, cfg);
                    if (tt.cfg_mask & cfg) {  // cfg 1 -> bit0, cfg 2 -> bit1
                        V2 c[8];
                        valid = construct_corners(geo, s_vp + 6 * group, (double)(tt.top_x0 + top * tt.top_step), cfg, c) > 0;
                    }
                }
                unsigned bal = __ballot_sync(FULL, valid);
                if (lane == 0) s_wcnt[warp] = __popc(bal);
                __syncthreads();
                int pre = 0, tot = 0;
#pragma unroll
                for (int w = 0; w < NW; w++) { int cw = s_wcnt[w]; if (w < warp) pre += cw; tot += cw; }
                if (valid) s_queue[(q_head + q_count + pre + __popc(bal & ((1u << lane) - 1))) % QCAP] = h;
                q_count += tot;
                base += SCORE_THREADS;
                __syncthreads();
            }
            const bool flush = (base >= n_hyp);
            if (q_count >= SCORE_THREADS || (flush && q_count > 0)) {
                if (!map_ready) { mbar_wait(&s_bar, bar_parity); bar_parity ^= 1; map_ready = true; }
                const int m = min(q_count, SCORE_THREADS);
                if (tid < m) {
                    int h = s_queue[(q_head + tid) % QCAP];
                    int group, top, cfg;
                    decode_hyp(h, tt.n_top, group, top, cfg);
                    V2 c[8];
                    construct_corners(geo, s_vp + 6 * group, (double)(tt.top_x0 + top * tt.top_step), cfg, c);
                    double total_angle_diff = box_edge_alignment_angle_error(s_sup + 6 * group, c, cfg);
                    V2 cs[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) cs[i] = V2{c[i].x - geo.roi_l, c[i].y - geo.roi_t};
                    double sum_dist = map_smem ? box_edge_sum_dists<true>(s_map, tt.roi_h, tt.roi_w, cs, cfg)
                                               : box_edge_sum_dists<false>(gmap, tt.roi_h, tt.roi_w, cs, cfg);
                    size_t o = (size_t)tt.out_offset + n_done + tid;
                    B.p_dist[o] = sum_dist / tt.diag;
                    B.p_angle[o] = total_angle_diff;
                    B.p_hyp[o] = h;
                }
                q_head = (q_head + m) % QCAP;
                q_count -= m;
                n_done += m;
                __syncthreads();
            }
        }
        if (!map_ready) { mbar_wait(&s_bar, bar_parity); bar_parity ^= 1; }  // drain the copy before the buffer is reused
        if (tid == 0) B.n_valid[task] = n_done;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// k_select
// ------------------------------------------------------------------------------------------------
constexpr int SELECT_THREADS = 256;

// bitonic sort of idx[0..npad) by key[idx] ascending, ties by index; idx < 0 are +inf pads
__device__ __forceinline__ bool key_less(const double* key, int a, int b) {
    if (b < 0) return a >= 0;
    if (a < 0) return false;
    double ka = key[a], kb = key[b];
    return (ka < kb) || (ka == kb && a < b);
}
__device__ void bitonic_sort_idx(int* idx, int npad, const double* key, int tid, int nthreads) {
    for (int k = 2; k <= npad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < npad; i += nthreads) {
                int ixj = i ^ j;
                if (ixj > i) {
                    int a = idx[i], b = idx[ixj];
                    bool up = ((i & k) == 0);
                    bool swap = up ? key_less(key, b, a) : key_less(key, a, b);
                    if (swap) { idx[i] = b; idx[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

// block-wide exclusive scan of one int per thread (SELECT_THREADS threads); returns prefix, total via smem
__device__ __forceinline__ int block_excl_scan(int v, int* s_w, int tid, int& total) {
    const unsigned FULL = 0xffffffffu;
    int lane = tid & 31, warp = tid >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int pre = 0, tot = 0;
    for (int w = 0; w < SELECT_THREADS / 32; w++) { int c = s_w[w]; if (w < warp) pre += c; tot += c; }
    __syncthreads();
    total = tot;
    return pre + inc - v;
}

// Recompute corners + 3D object for proposal (task tt, hypothesis h).  Returns vp_1_position.
__device__ __forceinline__ int recover_object(const TaskTab& tt, const FrameTab& ft, int h, V2* c, Obj3D& o, int& cfg, int& group) {
    int top;
    decode_hyp(h, tt.n_top, group, top, cfg);
    int yaw_id = group % ft.n_yaw, pair = group / ft.n_yaw;
    double vp[6];
    vanishing_points(ft.KinvR[pair], ft.cosy[yaw_id], ft.siny[yaw_id], vp);
    TaskGeo geo = make_geo(tt);
    int vp1 = construct_corners(geo, vp, (double)(tt.top_x0 + top * tt.top_step), cfg, c);
    corners_to_3d(c, ft.Tnew[pair], ft.invK, o);
    return vp1;
}

__global__ void __launch_bounds__(SELECT_THREADS) k_select(DetectBuffers B, int n_cap /* smem capacity in proposals */) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_w[SELECT_THREADS / 32];
    __shared__ double s_red[4 * (SELECT_THREADS / 32)];
    const int task = blockIdx.x, tid = threadIdx.x;
    const TaskTab tt = B.ttab[task];
    const FrameTab& ft = B.ftab[tt.frame_id];
    const int N = B.n_valid[task];
    const size_t ob = (size_t)tt.out_offset;
    if (N == 0) {
        if (tid == 0) { B.n_keep[task] = 0; B.n_cand[task] = 0; }
        return;
    }
    int npad = 1;
    while (npad < N) npad <<= 1;
    // working arrays: shared memory when the task fits, else the task's global scratch slots
    double *vd, *va;
    int* idx;
    unsigned char* flag;
    if (N <= n_cap) {
        vd = reinterpret_cast<double*>(smem_raw);
        va = vd + n_cap;
        idx = reinterpret_cast<int*>(va + n_cap);
        int cap_pad = 1; while (cap_pad < n_cap) cap_pad <<= 1;
        flag = reinterpret_cast<unsigned char*>(idx + cap_pad);
        for (int i = tid; i < N; i += SELECT_THREADS) { vd[i] = B.p_dist[ob + i]; va[i] = B.p_angle[ob + i]; }
    } else {
        vd = B.p_dist + ob; va = B.p_angle + ob;
        idx = B.sel_idx + 2 * ob;  // 2x slots: room for the power-of-two padding (npad < 2N)
        flag = B.sel_flag + ob;
    }
    __syncthreads();

    int* keep = B.keep + ob;
    int n_keep = 0;
    if (N > 4) {
        // fuse_normalize_scores_v2, object_3d_util.cpp:736-787
        const int k = (int)round((double)(float)N / 3.0 * 2.0);  // breaking_num
        // angle list: only order statistics k-1, k-2 and (if there is a strict gap) the k-1 smallest as a set
        for (int i = tid; i < npad; i += SELECT_THREADS) idx[i] = (i < N) ? i : -1;
        __syncthreads();
        bitonic_sort_idx(idx, npad, va, tid, SELECT_THREADS);
        const bool angle_active = va[idx[k - 1]] > va[idx[k - 2]];
        for (int i = tid; i < N; i += SELECT_THREADS) flag[i] = 0;
        __syncthreads();
        if (angle_active)
            for (int p = tid; p < k - 1; p += SELECT_THREADS) flag[idx[p]] = 1;
        __syncthreads();
        // distance list
        for (int i = tid; i < npad; i += SELECT_THREADS) idx[i] = (i < N) ? i : -1;
        __syncthreads();
        bitonic_sort_idx(idx, npad, vd, tid, SELECT_THREADS);
        const double vk = vd[idx[k - 1]];
        // does the unstable partial_sort matter?  (a) several elements equal the k-th smallest value -> membership / which
        // one is dropped depends on the heap; (b) with the angle filter off the kept list keeps partial_sort's ORDER, so any
        // tie inside the first k positions matters too.
        int local = 0;
        for (int i = tid; i < N; i += SELECT_THREADS) local += (vd[i] == vk) ? 1 : 0;
        int mult;
        block_excl_scan(local, s_w, tid, mult);
        int local2 = 0;
        if (!angle_active)
            for (int p = tid; p < k - 1; p += SELECT_THREADS) local2 += (vd[idx[p]] == vd[idx[p + 1]]) ? 1 : 0;
        int inner_ties;
        block_excl_scan(local2, s_w, tid, inner_ties);
        const bool need_emul = (mult > 1) || (!angle_active && inner_ties > 0);
        if (need_emul) {
            // literal std::partial_sort on the iota vector (matrix_utils.cpp:327-335), one thread
            for (int i = tid; i < N; i += SELECT_THREADS) idx[i] = i;
            __syncthreads();
            if (tid == 0) {
                auto less = [vd](int a, int b) { return vd[a] < vd[b]; };
                heap_select(idx, k, N, less);
                if (!angle_active) heap_sort(idx, k, less);
            }
            __syncthreads();
        }
        if (angle_active) {
            // dist_keep = sorted[0..k-1) as a SET (after heap_select the excluded k-th element sits at the heap top, idx[0])
            if (need_emul) { for (int p = tid + 1; p < k; p += SELECT_THREADS) flag[idx[p]] |= 2; }
            else { for (int p = tid; p < k - 1; p += SELECT_THREADS) flag[idx[p]] |= 2; }
            __syncthreads();
            // std::set_intersection of the two index-sorted sets == ascending indices with both flags
            for (int b0 = 0; b0 < N; b0 += SELECT_THREADS) {
                int i = b0 + tid;
                int v = (i < N && flag[i] == 3) ? 1 : 0;
                int tot;
                int pre = block_excl_scan(v, s_w, tid, tot);
                if (v) keep[n_keep + pre] = i;
                n_keep += tot;
            }
        } else {
            // final_keep_inds = dist_keep_inds in partial_sort order (:783-786)
            n_keep = k - 1;
            for (int p = tid; p < n_keep; p += SELECT_THREADS) keep[p] = idx[p];
        }
    } else {
        n_keep = N;
        for (int p = tid; p < N; p += SELECT_THREADS) keep[p] = p;
    }
    __syncthreads();

    // min / max of the kept errors (:798-817) and combined score (:820-836)
    double mn_d = 1e300, mx_d = -1e300, mn_a = 1e300, mx_a = -1e300;
    for (int j = tid; j < n_keep; j += SELECT_THREADS) {
        double d = vd[keep[j]], a = va[keep[j]];
        mn_d = cmin(mn_d, d); mx_d = cmax(mx_d, d); mn_a = cmin(mn_a, a); mx_a = cmax(mx_a, a);
    }
    {
        const unsigned FULL = 0xffffffffu;
        for (int o = 16; o > 0; o >>= 1) {
            mn_d = cmin(mn_d, __shfl_xor_sync(FULL, mn_d, o)); mx_d = cmax(mx_d, __shfl_xor_sync(FULL, mx_d, o));
            mn_a = cmin(mn_a, __shfl_xor_sync(FULL, mn_a, o)); mx_a = cmax(mx_a, __shfl_xor_sync(FULL, mx_a, o));
        }
        int lane = tid & 31, warp = tid >> 5;
        if (lane == 0) { s_red[4 * warp] = mn_d; s_red[4 * warp + 1] = mx_d; s_red[4 * warp + 2] = mn_a; s_red[4 * warp + 3] = mx_a; }
        __syncthreads();
        mn_d = 1e6; mx_d = -1; mn_a = 1e6; mx_a = -1;  // the reference's initial values (:798-801)
        for (int w = 0; w < SELECT_THREADS / 32; w++) {
            mn_d = cmin(mn_d, s_red[4 * w]); mx_d = cmax(mx_d, s_red[4 * w + 1]); mn_a = cmin(mn_a, s_red[4 * w + 2]); mx_a = cmax(mx_a, s_red[4 * w + 3]);
        }
    }
    const double w_ang = 0.8;
    double* norm_score = B.norm_score + ob;
    for (int j = tid; j < n_keep; j += SELECT_THREADS) {
        double d = vd[keep[j]], a = va[keep[j]];
        double comb;
        if (n_keep > 1) {
            comb = (d - mn_d) / (mx_d - mn_d);
            double ak = a;
            if ((mx_a - mn_a) > 0) ak = (a - mn_a) / (mx_a - mn_a);
            comb = (comb + w_ang * ak) / (1 + w_ang);
        } else
            comb = (d + w_ang * a) / (1 + w_ang);
        norm_score[j] = comb;
    }
    if (tid == 0) B.n_keep[task] = n_keep;
    __syncthreads();

    // 3D recovery of the kept proposals + skew-augmented score (box_proposal_detail.cpp:723-822); candidates keep list order
    const double weight_skew_error = 1.5;
    int n_cand = 0;
    for (int b0 = 0; b0 < n_keep; b0 += SELECT_THREADS) {
        int j = b0 + tid;
        int ok = 0;
        double score = 0;
        if (j < n_keep) {
            int h = B.p_hyp[ob + keep[j]];
            V2 c[8];
            Obj3D o;
            int cfg, group;
            recover_object(tt, ft, h, c, o, cfg, group);
            if (!((o.scale[0] < 0) || (o.scale[1] < 0) || (o.scale[2] < 0))) {
                ok = 1;
                double skew_ratio = cmax(o.scale[0], o.scale[1]) / cmin(o.scale[0], o.scale[1]);
                double skew_error = weight_skew_error * cmax(skew_ratio - B.dc.nominal_skew_ratio, 0.0);
                if (skew_ratio > B.dc.max_cut_skew) skew_error = 100;
                score = norm_score[j] + weight_skew_error * skew_error;
            }
        }
        int tot;
        int pre = block_excl_scan(ok, s_w, tid, tot);
        if (ok) { B.cand_score[ob + n_cand + pre] = score; B.cand_keeppos[ob + n_cand + pre] = j; }
        n_cand += tot;
    }
    if (tid == 0) B.n_cand[task] = n_cand;
}

// ------------------------------------------------------------------------------------------------
// k_rank : one warp per 2D box
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_rank(DetectBuffers B) {
    const int box = blockIdx.x, lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const int t0 = B.box_task_begin[box], t1 = B.box_task_begin[box + 1];
    int M = 0;
    for (int t = t0; t < t1; t++) M += B.n_cand[t];
    const int kmax = B.dc.max_cuboid_num;
    const int k = min(kmax, M);
    if (lane == 0) B.n_cuboids[box] = k;
    if (k == 0) return;
    // candidate r of the box -> (task, position in that task's candidate list); tasks of a box are contiguous
    auto locate = [&](int r, int& t, int& c) {
        t = t0;
        while (r >= B.n_cand[t]) { r -= B.n_cand[t]; t++; }
        c = r;
    };
    auto score_of = [&](int r) {
        int t, c;
        locate(r, t, c);
        return B.cand_score[(size_t)B.ttab[t].out_offset + c];
    };
    int* ridx = B.rank_idx + (size_t)B.ttab[t0].out_offset;  // scratch: >= M slots
    if (k == 1) {
        // partial_sort(idx, idx+1, end): start with element 0, replace by every later strictly smaller one.  That is the first
        // index of the minimum over the non-NaN scores -- unless score[0] is NaN, which nothing can replace.
        double best = 0; int bi = -1;
        for (int r = lane; r < M; r += 32) {
            double s = score_of(r);
            if (s == s && (bi < 0 || s < best)) { best = s; bi = r; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(FULL, best, o); int oi = __shfl_down_sync(FULL, bi, o);
            if (oi >= 0 && (bi < 0 || ov < best || (ov == best && oi < bi))) { best = ov; bi = oi; }
        }
        bi = __shfl_sync(FULL, bi, 0);
        double s0 = score_of(0);
        if (s0 != s0 || bi < 0) bi = 0;
        if (lane == 0) ridx[0] = bi;
    } else {
        for (int r = lane; r < M; r += 32) ridx[r] = r;
        __syncwarp();
        if (lane == 0) {
            auto less = [&](int a, int b) { return score_of(a) < score_of(b); };
            heap_select(ridx, k, M, less);
            heap_sort(ridx, k, less);
        }
    }
    __syncwarp();
    // write the k cuboid records (lanes over winners)
    for (int w = lane; w < k; w += 32) {
        int r = ridx[w];
        int t, c;
        locate(r, t, c);
        const TaskTab tt = B.ttab[t];
        const FrameTab& ft = B.ftab[tt.frame_id];
        const size_t ob = (size_t)tt.out_offset;
        int j = B.cand_keeppos[ob + c];
        int vidx = B.keep[ob + j];
        int h = B.p_hyp[ob + vidx];
        V2 cr[8];
        Obj3D o;
        int cfg, group;
        int vp1 = recover_object(tt, ft, h, cr, o, cfg, group);
        int yaw_id = group % ft.n_yaw, pair = group / ft.n_yaw;
        int roll_id = pair / ft.n_pitch, pitch_id = pair % ft.n_pitch;
        csb_cuboid& out = B.cuboids[(size_t)box * kmax + w];
        for (int i = 0; i < 3; i++) { out.pos[i] = o.pos[i]; out.scale[i] = o.scale[i]; }
        out.rotY = ft.yaw[yaw_id];
        out.box_config_type[0] = (double)cfg; out.box_config_type[1] = (double)vp1;
        // corner re-indexing, object_3d_util.cpp:994-1007
        const int ids_l[8] = {6, 5, 8, 7, 2, 3, 4, 1}, ids_r[8] = {5, 6, 7, 8, 3, 2, 1, 4};
        for (int i = 0; i < 8; i++) {
            int id = (vp1 == 1 ? ids_l[i] : ids_r[i]) - 1;
            out.box_corners_2d[i] = __double2int_rz(cr[id].x);
            out.box_corners_2d[8 + i] = __double2int_rz(cr[id].y);
        }
        // compute3D_BoxCorner / similarityTransformation, object_3d_util.cpp:15-73 (cos/sin(rotY) from the host yaw table)
        {
            const double body[3][8] = {{1, 1, -1, -1, 1, 1, -1, -1}, {1, -1, -1, 1, 1, -1, -1, 1}, {-1, -1, -1, -1, 1, 1, 1, 1}};
            double cy = ft.cosy[yaw_id], sy = ft.siny[yaw_id];
            double rot[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
            double res[12];
            for (int i = 0; i < 3; i++) {
                for (int jj = 0; jj < 3; jj++) {
                    double sc0 = (jj == 0) ? o.scale[0] : 0.0, sc1 = (jj == 1) ? o.scale[1] : 0.0, sc2 = (jj == 2) ? o.scale[2] : 0.0;
                    res[i * 4 + jj] = (rot[i * 3] * sc0 + rot[i * 3 + 1] * sc1) + rot[i * 3 + 2] * sc2;
                }
                res[i * 4 + 3] = o.pos[i];
            }
            for (int kk = 0; kk < 8; kk++) {
                double h3 = ((0.0 * body[0][kk] + 0.0 * body[1][kk]) + 0.0 * body[2][kk]) + 1.0 * 1.0;
                for (int rr = 0; rr < 3; rr++) {
                    double hv = ((res[rr * 4] * body[0][kk] + res[rr * 4 + 1] * body[1][kk]) + res[rr * 4 + 2] * body[2][kk]) + res[rr * 4 + 3] * 1.0;
                    out.box_corners_3d_world[rr * 8 + kk] = hv / h3;
                }
            }
        }
        out.rect_detect_2d[0] = (double)tt.left_x_raw; out.rect_detect_2d[1] = (double)tt.top_y_raw;
        out.rect_detect_2d[2] = (double)tt.obj_width_raw; out.rect_detect_2d[3] = (double)tt.obj_height_raw;
        out.edge_distance_error = B.p_dist[ob + vidx];
        out.edge_angle_error = B.p_angle[ob + vidx];
        out.normalized_error = B.norm_score[ob + j];
        out.skew_ratio = cmax(o.scale[0], o.scale[1]) / cmin(o.scale[0], o.scale[1]);
        out.down_expand_height = (double)tt.down_expand;
        if (ft.sample_rp) {
            out.camera_roll_delta = ft.roll[roll_id] - ft.euler_raw[0];
            out.camera_pitch_delta = ft.pitch[pitch_id] - ft.euler_raw[1];
        } else { out.camera_roll_delta = 0; out.camera_pitch_delta = 0; }
        out.task_id = t; out.raw_cube_ind = vidx; out.rank_index = r; out.reserved = 0;
    }
}

// debug: recompute the 2x8 corners of every valid proposal of one task
__global__ void k_debug_corners(DetectBuffers B, int task, double* out) {
    const TaskTab tt = B.ttab[task];
    const FrameTab& ft = B.ftab[tt.frame_id];
    int N = B.n_valid[task];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int h = B.p_hyp[(size_t)tt.out_offset + i];
    V2 c[8];
    Obj3D o;
    int cfg, group;
    recover_object(tt, ft, h, c, o, cfg, group);
    for (int k = 0; k < 8; k++) { out[16 * (size_t)i + k] = c[k].x; out[16 * (size_t)i + 8 + k] = c[k].y; }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static size_t score_smem_bytes(int groups_cap, int map_cap_floats) {
    return (size_t)map_cap_floats * 4 + (size_t)groups_cap * 12 * 8 + (size_t)LINE_SMEM_CAP * 3 * 8 + (size_t)(2 * SCORE_THREADS + SCORE_THREADS / 32) * 4 + 64;
}

cudaError_t launch_prep_lines(const DetectBuffers& B, int max_lines_per_frame, cudaStream_t st) {
    int cap = max_lines_per_frame < 1 ? 1 : max_lines_per_frame;
    size_t smem = (size_t)cap * 5 * 8;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_prep_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_prep_lines<<<B.n_tasks, 32, smem, st>>>(B, cap);
    return cudaGetLastError();
}

cudaError_t launch_score(const DetectBuffers& B, int max_groups, int num_sms, int max_smem_optin, int* map_cap_floats_out, cudaStream_t st) {
    int groups_cap = max_groups;
    size_t fixed = score_smem_bytes(groups_cap, 0);
    size_t budget = (size_t)max_smem_optin - 1024;  // static __shared__ + slack
    if (fixed + 16 * 1024 > budget) return cudaErrorInvalidValue;
    int map_cap = (int)((budget - fixed) / 4) & ~31;
    size_t smem = score_smem_bytes(groups_cap, map_cap);
    cudaError_t e = cudaFuncSetAttribute(k_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (map_cap_floats_out) *map_cap_floats_out = map_cap;
    int grid = B.n_tasks < num_sms ? B.n_tasks : num_sms;
    k_score<<<grid, SCORE_THREADS, smem, st>>>(B, groups_cap, map_cap);
    return cudaGetLastError();
}

cudaError_t launch_select(const DetectBuffers& B, int max_hyp_per_task, int max_smem_optin, cudaStream_t st) {
    // per proposal: 2 doubles + idx (power-of-two padded: <= 2 ints) + 1 flag byte
    int n_cap = max_hyp_per_task;
    auto bytes = [](int n) { int p = 1; while (p < n) p <<= 1; return (size_t)n * 16 + (size_t)p * 4 + (size_t)n + 16; };
    size_t budget = (size_t)max_smem_optin - 4096;
    if (n_cap > 8192) n_cap = 8192;
    while (bytes(n_cap) > budget) n_cap -= 256;
    size_t smem = bytes(n_cap);
    cudaError_t e = cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_select<<<B.n_tasks, SELECT_THREADS, smem, st>>>(B, n_cap);
    return cudaGetLastError();
}

cudaError_t launch_rank(const DetectBuffers& B, int n_boxes, cudaStream_t st) {
    if (n_boxes == 0) return cudaSuccess;
    k_rank<<<n_boxes, 32, 0, st>>>(B);
    return cudaGetLastError();
}

cudaError_t launch_debug_corners(const DetectBuffers& B, int task, int n_valid, double* out, cudaStream_t st) {
    if (n_valid == 0) return cudaSuccess;
    k_debug_corners<<<(n_valid + 127) / 128, 128, 0, st>>>(B, task, out);
    return cudaGetLastError();
}

}  // namespace csb
