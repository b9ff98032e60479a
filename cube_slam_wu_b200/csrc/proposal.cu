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
#include <cstdio>

#include "proposal.h"
#include "proposal_score.cuh"
#include "ptx_helpers.cuh"

namespace csb {

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
    double merged_angle = det_atan2(mey - msy, mex - msx);
    double temp = fabs(L.ang[a] - merged_angle);
    double merge_angle_diff = cmin(temp, M_PI - temp);
    if (!(merge_angle_diff < thr_ang)) return false;
    out[0] = msx; out[1] = msy; out[2] = mex; out[3] = mey; out[4] = merged_angle;
    return true;
}

// Per-phase cycle counters of k_score (thread 0 of every CTA, summed over CTAs and tasks), compiled in only with -DCSB_SCORE_PHASES
// (CSB_SCORE_PHASES=1 python -m cube_slam_wu_b200.build --force; tools/score_phases.py).  Read and reset through csb_detect_debug_score_phases().
__device__ unsigned long long g_score_phase_cycles[12];  // [8..10]: VP-support units decided by the float / double / exact tier
#ifdef CSB_SCORE_PHASES
#define SCORE_PHASE(idx)                                                            \
    do {                                                                            \
        if (tid == 0) {                                                             \
            const long long now__ = clock64();                                      \
            atomicAdd(&g_score_phase_cycles[idx], (unsigned long long)(now__ - t_prev)); \
            t_prev = now__;                                                         \
        }                                                                           \
    } while (0)
#else
#define SCORE_PHASE(idx) do { } while (0)
#endif

cudaError_t score_phase_cycles(unsigned long long* out12, bool reset) {
#ifndef CSB_SCORE_PHASES
    (void)reset;
    for (int i = 0; i < 12; i++) out12[i] = 0;
    return cudaErrorNotSupported;  // the counters are not compiled into this build
#else
    cudaError_t e = cudaMemcpyFromSymbol(out12, g_score_phase_cycles, sizeof(unsigned long long) * 12);
    if (e != cudaSuccess) return e;
    if (reset) {
        unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        e = cudaMemcpyToSymbol(g_score_phase_cycles, z, sizeof z);
    }
    return e;
#endif
}

constexpr int SCORE_CTAS_PER_SM = 1;  // 2 x 256 threads was measured slower (maps no longer fit shared memory, coarser tail)
constexpr int SCORE_THREADS = 768;  // 24 warps at 80 registers (some spills) beat 16 warps at 124 and 28 / 32 warps with more spills (bench value: 640 -> 0.275, 768 -> 0.269, 896 -> 0.269 ms per step)
constexpr int SUBW = 8;  // lanes per VP-support unit

// VP_support_edge_infos (object_3d_util.cpp:548-619) for one (vanishing point, line table) unit, executed by an 8-lane
// sub-group; all 32 lanes of the warp run this in lock step (ballots / shuffles are warp-wide), `active` masks units that
// do not exist.  Returns the two supporting line angles (NaN if none) on every lane of the sub-group.
//
// lcs (optional, n <= 64): cos/sin of every line angle.  A line supports the VP iff the angle between its direction and the
// ray VP -> midpoint is below thr modulo pi, i.e. |cross(u_line, d)|^2 < sin^2(thr) |d|^2.  Lines that fail this test with a
// margin of 1e-6 rad are certain outliers and never reach atan2; the remaining candidates (typically 10-20 %) are packed
// densely onto the lanes and decided exactly as the reference does (atan2 -> normalize_to_pi -> compare).
__device__ __forceinline__ void vp_support_unit(bool active, double vx, double vy, double thr, double s2_margin, int n, const double* ang, const double* mid,
                                                const double* lcs, int lane, int swap_lt, double& out_low, double& out_top) {
    const unsigned FULL = 0xffffffffu;
    const int sl = lane & (SUBW - 1), sbase = lane & ~(SUBW - 1), sshift = sbase;
    bool have_base = false;
    double base = 0;
    // lane-local extrema of the smoothed inlier angles; ties keep the lowest line index (Eigen max/minCoeff: first wins)
    double vmax = 0, vmin = 0;
    int imax = -1, imin = -1;
    const bool packed = (lcs != nullptr) && (n <= 64);
    unsigned cm_lo = 0, cm_hi = 0;  // candidate lines of this unit (bit e)
    int n_iter = n;
    if (packed) {
        for (int b0 = 0; b0 < n; b0 += SUBW) {
            const int e = b0 + sl;
            bool cand = false;
            if (active && e < n) {
                const double dx = mid[2 * e] - vx, dy = mid[2 * e + 1] - vy;
                const double cr = dx * lcs[2 * e + 1] - dy * lcs[2 * e];
                cand = !(cr * cr > s2_margin * (dx * dx + dy * dy));  // NaN / inf -> candidate (decided exactly below)
            }
            const unsigned sub = (__ballot_sync(FULL, cand) >> sshift) & ((1u << SUBW) - 1);
            if (b0 < 32) cm_lo |= sub << b0; else cm_hi |= sub << (b0 - 32);
        }
        n_iter = __popc(cm_lo) + __popc(cm_hi);
        // all sub-groups of the warp must run the same number of (ballot / shuffle) rounds
        for (int off = 16; off >= SUBW; off >>= 1) n_iter = max(n_iter, __shfl_xor_sync(FULL, n_iter, off));
    }
    const int n_lo = __popc(cm_lo), n_cand = n_lo + __popc(cm_hi);
    for (int b0 = 0; b0 < n_iter; b0 += SUBW) {
        int e = b0 + sl;
        bool have = active && e < n;
        if (packed) {
            have = active && e < n_cand;
            if (have) e = (e < n_lo) ? (int)__fns(cm_lo, 0, e + 1) : 32 + (int)__fns(cm_hi, 0, e - n_lo + 1);
        }
        bool inl = false;
        double raw = 0;
        if (have) {
            raw = det_atan2(mid[2 * e + 1] - vy, mid[2 * e] - vx);
            double nrm = normalize_to_pi(raw);
            double d = fabs(ang[e] - nrm);
            d = cmin(d, M_PI - d);
            inl = d < thr;
        }
        const unsigned sub = (__ballot_sync(FULL, inl) >> sshift) & ((1u << SUBW) - 1);
        const double cand = __shfl_sync(FULL, raw, sbase + (sub ? (__ffs(sub) - 1) : 0));
        if (!have_base && sub) { base = cand; have_base = true; }  // smooth_jump_angles: base = first inlier (:285)
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
#pragma unroll
    for (int off = SUBW / 2; off > 0; off >>= 1) {
        double ov = __shfl_xor_sync(FULL, vmax, off); int oi = __shfl_xor_sync(FULL, imax, off);
        if (oi >= 0 && (imax < 0 || ov > vmax || (ov == vmax && oi < imax))) { vmax = ov; imax = oi; }
        ov = __shfl_xor_sync(FULL, vmin, off); oi = __shfl_xor_sync(FULL, imin, off);
        if (oi >= 0 && (imin < 0 || ov < vmin || (ov == vmin && oi < imin))) { vmin = ov; imin = oi; }
    }
    if (imax >= 0) {
        int low = imax, top = imin;
        if (swap_lt) { int t = low; low = top; top = t; }  // "match matlab code" (:609-610)
        out_low = ang[low];
        out_top = ang[top];
    } else {
        out_low = nan("");
        out_top = nan("");
    }
}

// Fast path of VP_support_edge_infos (vp_support_mixed), one THREAD per unit, no atan2.  The reference's decisions are all angle comparisons:
//  * inlier: acute angle between the line direction u and the ray d = mid - VP below thr  <=>  cross(u, d)^2 < sin^2(thr) |d|^2;
//  * smooth_jump_angles + max/minCoeff: with d0 the ray of the first inlier, v - base is the signed angle from d0 to d wrapped to
//    (-pi, pi]; the largest v is the most counter-clockwise ray of the open upper half plane cross(d0, d) > 0 (d0 itself if that is
//    empty), the smallest v the most clockwise ray of the lower half plane; inside a half plane "more counter-clockwise" is the
//    sign of one cross product.
// Every comparison is taken with a guard band that is orders of magnitude wider than the rounding of the arithmetic used here and of
// the reference's atan2 / subtractions, so a decision taken here equals the reference's; anything inside a guard band (also NaN /
// inf from a vanishing point at infinity, duplicate rays, rays opposite to d0) is escalated:
//   float  (guards: 1e-4 rad around thr and between rays, rays shorter than 10 px refused)   -> every test, FP32 pipe
//   double (guards: 1e-6 rad around thr, 1e-9 rad between rays)                              -> the single tests float refused, inline
//   vp_support_unit(), the reference's expressions literally (atan2, normalize_to_pi, smooth_jump_angles) -> units with a test that
//   double refused as well (function returns false)
// +1: ray e is counter-clockwise of ray r, -1: clockwise, 0: cannot tell (inside the guard band even in double)
__device__ __forceinline__ int ray_ccw(int r, float rx, float ry, float rn2, int e, float ex, float ey, float en2, double vxd, double vyd, const double* mid) {
    const float cf = rx * ey - ry * ex;
    if (cf * cf > 1e-8f * rn2 * en2 && rn2 > 100.0f && en2 > 100.0f) return cf > 0 ? 1 : -1;  // > 1e-4 rad apart, rays longer than 10 px
    const double ax = mid[2 * r] - vxd, ay = mid[2 * r + 1] - vyd, bx = mid[2 * e] - vxd, by = mid[2 * e + 1] - vyd;
    const double c = ax * by - ay * bx;
    if (c * c > 1e-18 * (ax * ax + ay * ay) * (bx * bx + by * by)) return c > 0 ? 1 : -1;        // > 1e-9 rad apart
    return 0;
}

__device__ __forceinline__ bool vp_support_mixed(double vxd, double vyd, float s2_lo_f, float s2_hi_f, double s2_lo_d, double s2_hi_d, int n, const double* ang,
                                                 const double* mid, const double* lcs, const float* midf, const float* lcsf, int swap_lt, double& out_low,
                                                 double& out_top) {
    const float vx = (float)vxd, vy = (float)vyd;
    const float2* midf2 = reinterpret_cast<const float2*>(midf);
    const float2* lcsf2 = reinterpret_cast<const float2*>(lcsf);
    bool amb = false;
    int ibase = -1, imax = -1, imin = -1;
    float bx = 0, by = 0, bn2 = 0, ux = 0, uy = 0, un2 = 0, lx = 0, ly = 0, ln2 = 0;
    for (int e0 = 0; e0 < n; e0 += 32) {
        // inlier bits of 32 lines (registers only), then the ordering logic on the set bits, in line order
        unsigned m = 0;
        const int cnt = (n - e0 < 32) ? (n - e0) : 32;
#pragma unroll 4
        for (int b = 0; b < cnt; b++) {
            const int e = e0 + b;
            const float2 mp = midf2[e], cs = lcsf2[e];
            const float dx = mp.x - vx, dy = mp.y - vy;
            const float cr = __fmaf_rn(dx, cs.y, -dy * cs.x);  // guarded test: contraction is harmless here
            const float n2 = __fmaf_rn(dx, dx, dy * dy), cr2 = cr * cr;
            bool in = cr2 < s2_lo_f * n2;
            const bool out = cr2 > s2_hi_f * n2;
            if (!(in || out) || !(n2 > 100.0f)) {
                const double ddx = mid[2 * e] - vxd, ddy = mid[2 * e + 1] - vyd;
                const double dcr = ddx * lcs[2 * e + 1] - ddy * lcs[2 * e];
                const double dn2 = ddx * ddx + ddy * ddy, dcr2 = dcr * dcr;
                const bool in_d = dcr2 < s2_lo_d * dn2, out_d = dcr2 > s2_hi_d * dn2;
                if (!(in_d || out_d)) amb = true;
                in = in_d;
            }
            m |= (in ? 1u : 0u) << b;
        }
        while (m) {
            const int e = e0 + __ffs(m) - 1;
            m &= m - 1;
            const float2 mp = midf2[e];
            const float dx = mp.x - vx, dy = mp.y - vy;
            const float n2 = __fmaf_rn(dx, dx, dy * dy);
            // float crosses against the base ray and both current extremes (select-only fast path; the branchy generic path below
            // runs when a needed comparison falls inside the float guard band)
            const float cb = __fmaf_rn(bx, dy, -by * dx), cu = __fmaf_rn(ux, dy, -uy * dx), cl = __fmaf_rn(lx, dy, -ly * dx);
            const bool longe = n2 > 100.0f;
            const bool okb = longe && bn2 > 100.0f && cb * cb > 1e-8f * bn2 * n2;
            const bool oku = longe && un2 > 100.0f && cu * cu > 1e-8f * un2 * n2;
            const bool okl = longe && ln2 > 100.0f && cl * cl > 1e-8f * ln2 * n2;
            const bool first = ibase < 0;
            const bool upper = cb > 0;
            const bool need_u = !first && upper && imax >= 0, need_l = !first && !upper && imin >= 0;
            if (first || (okb && (!need_u || oku) && (!need_l || okl))) {
                const bool take_u = !first && upper && (imax < 0 || cu > 0);
                const bool take_l = !first && !upper && (imin < 0 || cl < 0);
                if (first) { ibase = e; bx = dx; by = dy; bn2 = n2; }
                if (take_u) { imax = e; ux = dx; uy = dy; un2 = n2; }
                if (take_l) { imin = e; lx = dx; ly = dy; ln2 = n2; }
            } else {
                const int side = ray_ccw(ibase, bx, by, bn2, e, dx, dy, n2, vxd, vyd, mid);  // upper / lower half plane of d0
                if (side == 0) amb = true;  // along d0 or opposite to it
                else if (side > 0) {
                    if (imax < 0) { imax = e; ux = dx; uy = dy; un2 = n2; }
                    else {
                        const int o = ray_ccw(imax, ux, uy, un2, e, dx, dy, n2, vxd, vyd, mid);
                        if (o == 0) amb = true;
                        else if (o > 0) { imax = e; ux = dx; uy = dy; un2 = n2; }
                    }
                } else {
                    if (imin < 0) { imin = e; lx = dx; ly = dy; ln2 = n2; }
                    else {
                        const int o = ray_ccw(imin, lx, ly, ln2, e, dx, dy, n2, vxd, vyd, mid);
                        if (o == 0) amb = true;
                        else if (o < 0) { imin = e; lx = dx; ly = dy; ln2 = n2; }
                    }
                }
            }
        }
    }
    if (amb) return false;
    if (ibase < 0) { out_low = nan(""); out_top = nan(""); return true; }
    int low = imax >= 0 ? imax : ibase, top = imin >= 0 ? imin : ibase;
    if (swap_lt) { int t = low; low = top; top = t; }  // "match matlab code" (:609-610)
    out_low = ang[low];
    out_top = ang[top];
    return true;
}

constexpr int PREP_THREADS = 256;
constexpr int PREP_WARPS = PREP_THREADS / 32;

// order-preserving block compaction step: returns this thread's output slot (valid when `in`) and adds the chunk's count to total
__device__ __forceinline__ int prep_compact(bool in, int* s_cnt, int tid, int& total) {
    const unsigned bal = __ballot_sync(0xffffffffu, in);
    const int lane = tid & 31, warp = tid >> 5;
    if (lane == 0) s_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < PREP_WARPS; w++) { int c = s_cnt[w]; if (w < warp) before += c; all += c; }
    const int pos = total + before + __popc(bal & ((1u << lane) - 1));
    total += all;
    __syncthreads();
    return pos;
}

// One block (8 warps) per task: merged long-line table of the task's ROI.  Dynamic shared memory: cap * (5 doubles + 3 ints).
__global__ void __launch_bounds__(PREP_THREADS, 4) k_prep_lines(DetectBuffers B, int cap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_cnt[PREP_WARPS];
    __shared__ int s_hit, s_njobs;
    const int task = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TaskTab& tt = B.ttab[task];
    const FrameTab& ft = B.ftab[tt.frame_id];
    LineSM L;
    L.x1 = reinterpret_cast<double*>(smem_raw);
    L.y1 = L.x1 + cap; L.x2 = L.y1 + cap; L.y2 = L.x2 + cap; L.ang = L.y2 + cap;
    int* first = reinterpret_cast<int*>(L.ang + cap);
    int* jobs = first + cap;  // rows to re-scan this round: two ints (row, from) per job
    const unsigned FULL = 0xffffffffu;

    // (1) align left->right (object_3d_util.cpp:246-258) and keep lines with both endpoints inside the expanded ROI
    //     (box_proposal_detail.cpp:271-283); order-preserving compaction
    const double rl = (double)tt.roi_left, rt = (double)tt.roi_top, rr = (double)tt.roi_right, rd = (double)tt.roi_down;
    int total = 0;
    const int n_raw = ft.line_end - ft.line_begin;
    for (int base = 0; base < n_raw; base += PREP_THREADS) {
        int i = base + tid;
        bool in = false;
        double a = 0, b = 0, c = 0, d = 0;
        if (i < n_raw) {
            const double* p = B.lines + 4 * (size_t)(ft.line_begin + i);
            a = p[0]; b = p[1]; c = p[2]; d = p[3];
            if (c < a) { double t0 = a, t1 = b; a = c; b = d; c = t0; d = t1; }
            in = inside_box(V2{a, b}, rl, rt, rr, rd) && inside_box(V2{c, d}, rl, rt, rr, rd);
        }
        const int pos = prep_compact(in, s_cnt, tid, total);
        if (in) {
            L.x1[pos] = a; L.y1[pos] = b; L.x2[pos] = c; L.y2[pos] = d;
            L.ang[pos] = det_atan2(d - b, c - a);
        }
    }
    __syncthreads();

    // (2) merge_break_lines (object_3d_util.cpp:431-511): repeatedly merge the first (seg1, seg2) pair in lexicographic
    // order that passes the angle / gap / merged-angle tests, then restart from scratch (<= 500 rounds).
    // The restart is replayed exactly but incrementally: first[s1] memoises the smallest seg2 > s1 that passes against
    // s1 (-1: none).  A merge at (m, s2) rewrites row m, moves the last row into slot s2 and drops the last slot; every
    // other pair is unchanged, so a row only (i) re-tests the two rewritten slots and (ii) re-scans from its old first hit
    // if that hit was one of the rewritten / dropped slots.  Each round costs O(rows) pair tests spread over the block.
    const double thr_ang = 5.0 / 180.0 * M_PI, thr_dist = 20.0;
    // warp-cooperative scan of row s1 from slot `from`: smallest passing partner or -1 (uniform over the warp)
    auto scan_row = [&](int s1, int from) -> int {
        for (int base = from; base < total; base += 32) {
            int bb = base + lane;
            double o[5];
            bool ok = (bb < total) && merge_pred(L, s1, bb, thr_ang, thr_dist, o);
            unsigned bal = __ballot_sync(FULL, ok);
            if (bal) return base + __ffs(bal) - 1;
        }
        return -1;
    };
    for (int s1 = warp; s1 < total; s1 += PREP_WARPS) {
        int f = scan_row(s1, s1 + 1);
        if (lane == 0) first[s1] = f;
    }
    __syncthreads();
    int counter = 0;
    while (counter < 500) {
        counter++;
        // the first row that has a partner (warp 0 scans, everyone reads)
        if (warp == 0) {
            int hit = -1;
            for (int base = 0; base < total && hit < 0; base += 32) {
                int r = base + lane;
                unsigned bal = __ballot_sync(FULL, r < total && first[r] >= 0);
                if (bal) hit = base + __ffs(bal) - 1;
            }
            if (lane == 0) { s_hit = hit; s_njobs = 0; }
        }
        __syncthreads();
        const int hit_a = s_hit;
        if (hit_a < 0) break;
        const int hit_b = first[hit_a];
        __syncthreads();  // everyone has read first[hit_a] before the tables change
        if (tid == 0) {
            double mg[5];
            merge_pred(L, hit_a, hit_b, thr_ang, thr_dist, mg);
            L.x1[hit_a] = mg[0]; L.y1[hit_a] = mg[1]; L.x2[hit_a] = mg[2]; L.y2[hit_a] = mg[3]; L.ang[hit_a] = mg[4];
            // fast_RemoveRow (matrix_utils.cpp:183-187)
            L.x1[hit_b] = L.x1[total - 1]; L.y1[hit_b] = L.y1[total - 1]; L.x2[hit_b] = L.x2[total - 1]; L.y2[hit_b] = L.y2[total - 1];
            L.ang[hit_b] = L.ang[total - 1];
        }
        total--;
        const int dropped = total;  // index of the slot that no longer exists
        __syncthreads();
        // thread-parallel memo update of the untouched rows; rows that must re-scan are queued as jobs (row, from)
        for (int r = tid; r < total; r += PREP_THREADS) {
            int f, rescan_from = -1;
            if (r == hit_a || r == hit_b) {
                rescan_from = r + 1;  // rewritten row: full scan
                f = -1;
            } else {
                f = first[r];
                if (f == hit_a || f == hit_b) { rescan_from = f; f = -1; }   // that slot holds a different line now
                else if (f == dropped) f = -1;                               // its line lives in slot hit_b now (re-tested below)
                double o[5];
                // the two rewritten slots, smaller index first, only where they could precede the current first hit
                int c0 = hit_a < hit_b ? hit_a : hit_b, c1 = hit_a < hit_b ? hit_b : hit_a;
                if (c1 >= total) c1 = -1;  // hit_b was the last slot: nothing moved there
                bool found = false;
                if (c0 > r && (rescan_from < 0 ? (f < 0 || c0 < f) : c0 < rescan_from)) {
                    if (merge_pred(L, r, c0, thr_ang, thr_dist, o)) { f = c0; found = true; rescan_from = -1; }
                }
                if (!found && c1 > r && (rescan_from < 0 ? (f < 0 || c1 < f) : c1 < rescan_from)) {
                    if (merge_pred(L, r, c1, thr_ang, thr_dist, o)) { f = c1; rescan_from = -1; }
                }
            }
            first[r] = f;
            if (rescan_from >= 0) {
                int slot = atomicAdd(&s_njobs, 1);
                jobs[2 * slot] = r; jobs[2 * slot + 1] = rescan_from;
            }
        }
        __syncthreads();
        const int nj = s_njobs;
        for (int q = warp; q < nj; q += PREP_WARPS) {
            const int row = jobs[2 * q], from = jobs[2 * q + 1];
            int f = scan_row(row, from);
            if (lane == 0) first[row] = f;
        }
        __syncthreads();
    }

    // (3) length filter > 30 px (object_3d_util.cpp:517-539) + angle / midpoint tables (box_proposal_detail.cpp:309-315)
    const size_t ob = (size_t)tt.line_cap_offset;
    int n_out = 0;
    for (int base = 0; base < total; base += PREP_THREADS) {
        int i = base + tid;
        bool keep = false;
        if (i < total) keep = norm2(V2{L.x2[i] - L.x1[i], L.y2[i] - L.y1[i]}) > 30.0;
        const size_t pos = ob + prep_compact(keep, s_cnt, tid, n_out);
        if (keep) {
            const double a = L.ang[i];  // == det_atan2(y2-y1, x2-x1) of the stored endpoints
            const double mx = (L.x1[i] + L.x2[i]) / 2, my = (L.y1[i] + L.y2[i]) / 2;
            B.ml_seg[4 * pos + 0] = L.x1[i]; B.ml_seg[4 * pos + 1] = L.y1[i]; B.ml_seg[4 * pos + 2] = L.x2[i]; B.ml_seg[4 * pos + 3] = L.y2[i];
            B.ml_ang[pos] = a;
            B.ml_mid[2 * pos + 0] = mx;
            B.ml_mid[2 * pos + 1] = my;
        }
    }
    if (tid == 0) B.n_merged[task] = n_out;
}

// ------------------------------------------------------------------------------------------------
// k_vp_support : VP_support_edge_infos for every unit of every task -- (group, vp1), (group, vp2), (roll-pitch pair, vp3) --
// against the task's merged-line table; B.vp_sup gets 6 doubles per group (low/top of vp1, vp2, vp3).
// grid = (unit chunks, tasks), VPS_THREADS units per block, one thread per unit.  Dynamic shared memory: cap * (5 doubles + 4 floats).
// ------------------------------------------------------------------------------------------------
constexpr int VPS_THREADS = 128;

__global__ void __launch_bounds__(VPS_THREADS) k_vp_support(DetectBuffers B, int cap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_amb[VPS_THREADS];
    __shared__ int s_namb;
    const int task = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31;
    const TaskTab& tt = B.ttab[task];
    const FrameTab& ft = B.ftab[tt.frame_id];
    const int n_yaw = ft.n_yaw, n_pairs = ft.n_roll * ft.n_pitch, n_groups = n_pairs * n_yaw;
    const int n_units = 2 * n_groups + n_pairs;
    const int u_begin = blockIdx.x * VPS_THREADS;
    if (u_begin >= n_units) return;
    double* k_ang = reinterpret_cast<double*>(smem_raw);  // kept lines: angle | midpoint | cos,sin | float copies
    double* k_mid = k_ang + cap;
    double* k_cs = k_mid + 2 * cap;
    float* k_midf = reinterpret_cast<float*>(k_cs + 2 * cap);
    float* k_csf = k_midf + 2 * cap;
    const int n_lines = B.n_merged[task];
    {
        const double* lang = B.ml_ang + tt.line_cap_offset;
        const double* lmid = B.ml_mid + 2 * (size_t)tt.line_cap_offset;
        for (int i = tid; i < n_lines; i += VPS_THREADS) {
            const double a = lang[i], mx = lmid[2 * i], my = lmid[2 * i + 1];
            const double ca = cos(a), sa = sin(a);  // only feed guarded tests (vp_support_mixed) / the prefilter of vp_support_unit
            k_ang[i] = a; k_mid[2 * i] = mx; k_mid[2 * i + 1] = my; k_cs[2 * i] = ca; k_cs[2 * i + 1] = sa;
            k_midf[2 * i] = (float)mx; k_midf[2 * i + 1] = (float)my; k_csf[2 * i] = (float)ca; k_csf[2 * i + 1] = (float)sa;
        }
        if (tid == 0) s_namb = 0;
    }
    __syncthreads();
    double* sup = B.vp_sup + (size_t)task * B.sup_stride;
    auto unit_of = [&](int u, int& g, int& vp_id) {
        if (u < 2 * n_groups) { g = u >> 1; vp_id = u & 1; }
        else { g = (u - 2 * n_groups) * n_yaw; vp_id = 2; }
    };
    auto unit_vp = [&](int g, int vp_id, double& vx, double& vy) {
        const int yaw_id = g % n_yaw, pair = g / n_yaw;
        double vp[6];
        vanishing_points(ft.KinvR[pair], ft.cosy[yaw_id], ft.siny[yaw_id], vp);
        vx = vp_id == 0 ? vp[0] : (vp_id == 1 ? vp[2] : vp[4]);
        vy = vp_id == 0 ? vp[1] : (vp_id == 1 ? vp[3] : vp[5]);
    };
    auto store_unit = [&](int g, int vp_id, double lo, double tp, int first_lane, int stride) {
        if (vp_id < 2) { if (first_lane == 0) { sup[6 * g + 2 * vp_id] = lo; sup[6 * g + 2 * vp_id + 1] = tp; } }
        else for (int y = first_lane; y < n_yaw; y += stride) { sup[6 * (g + y) + 4] = lo; sup[6 * (g + y) + 5] = tp; }  // vp3 is shared by the pair's yaw samples
    };
    // sin^2 of the guard-band edges around the 15 / 10 degree thresholds
    const float fl12 = sinf((float)(15.0 / 180.0 * M_PI) - 1e-4f), fh12 = sinf((float)(15.0 / 180.0 * M_PI) + 1e-4f);
    const float fl3 = sinf((float)(10.0 / 180.0 * M_PI) - 1e-4f), fh3 = sinf((float)(10.0 / 180.0 * M_PI) + 1e-4f);
    const double dl12 = sin(15.0 / 180.0 * M_PI - 1e-6), dh12 = sin(15.0 / 180.0 * M_PI + 1e-6);
    const double dl3 = sin(10.0 / 180.0 * M_PI - 1e-6), dh3 = sin(10.0 / 180.0 * M_PI + 1e-6);
    {
        const int u = u_begin + tid;
        if (u < n_units) {
            int g, vp_id;
            unit_of(u, g, vp_id);
            double vx, vy, lo, tp;
            unit_vp(g, vp_id, vx, vy);
            const bool v3 = vp_id == 2;
            const bool ok = vp_support_mixed(vx, vy, v3 ? fl3 * fl3 : fl12 * fl12, v3 ? fh3 * fh3 : fh12 * fh12, v3 ? dl3 * dl3 : dl12 * dl12, v3 ? dh3 * dh3 : dh12 * dh12,
                                             n_lines, k_ang, k_mid, k_cs, k_midf, k_csf, vp_id > 0, lo, tp);
            if (ok) store_unit(g, vp_id, lo, tp, 0, 1);
            else s_amb[atomicAdd(&s_namb, 1)] = u;
        }
    }
    __syncthreads();
    // exact tier: 8-lane sub-groups replay the reference's expressions for the queued units
    const int n_amb = s_namb;
    constexpr int SUBS = VPS_THREADS / SUBW;
    const int sg = tid / SUBW, sl = tid & (SUBW - 1);
    const double* lcs = (n_lines <= 64) ? k_cs : nullptr;
    for (int u0 = 0; u0 < n_amb; u0 += SUBS) {
        const bool active = u0 + sg < n_amb;
        int g = 0, vp_id = 0;
        if (active) unit_of(s_amb[u0 + sg], g, vp_id);
        double vx, vy, lo, tp;
        unit_vp(g, vp_id, vx, vy);
        vp_support_unit(active, vx, vy, (vp_id != 2 ? 15.0 : 10.0) / 180.0 * M_PI, (vp_id != 2 ? dh12 * dh12 : dh3 * dh3), n_lines, k_ang, k_mid, lcs, lane, vp_id > 0, lo, tp);
        if (active) store_unit(g, vp_id, lo, tp, sl, SUBW);
    }
#ifdef CSB_SCORE_PHASES
    if (tid == 0) {
        const int mine = (n_units - u_begin < VPS_THREADS) ? (n_units - u_begin) : VPS_THREADS;
        atomicAdd(&g_score_phase_cycles[8], (unsigned long long)(mine - n_amb));
        atomicAdd(&g_score_phase_cycles[10], (unsigned long long)n_amb);
    }
#endif
}

// block-wide exclusive scan of one int per thread; s_w holds one slot per warp
template <int THREADS>
__device__ __forceinline__ int block_excl_scan(int v, int* s_w, int tid, int& total) {
    const unsigned FULL = 0xffffffffu;
    const int lane = tid & 31, warp = tid >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(FULL, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    int pre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) { int c = s_w[w]; if (w < warp) pre += c; tot += c; }
    __syncthreads();
    total = tot;
    return pre + inc - v;
}

// Persistent CTA: loops over tasks handed out by an atomic counter (largest first).
//   (a) TMA bulk copy of the task's distance map into shared memory (lands while (b)-(d) run)
//   (b) vanishing points per (roll,pitch,yaw) group -> shared memory
//   (c) VP-support angles of the groups (computed by k_prep_lines) -> shared memory
//   (d) phase 1: every hypothesis through the corner construction / rejection cascade -> validity bitmask (enumeration order)
//   (e) prefix sums over the bitmask words: proposal i of the compacted list <-> hypothesis id
//   (f) phase 2: one thread per surviving proposal (all lanes busy): corners again, 99/77 distance-map gathers, edge-angle error
__global__ void __launch_bounds__(SCORE_THREADS, SCORE_CTAS_PER_SM) k_score(DetectBuffers B, int groups_cap, int map_cap_floats, int words_cap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_map = reinterpret_cast<float*>(smem_raw);
    double* s_vp = reinterpret_cast<double*>(smem_raw + (size_t)map_cap_floats * 4);
    double* s_sup = s_vp + 6 * (size_t)groups_cap;
    unsigned* s_mask = reinterpret_cast<unsigned*>(s_sup + 6 * (size_t)groups_cap);
    int* s_wpre = reinterpret_cast<int*>(s_mask + words_cap);  // words_cap + 1 entries
    __shared__ uint64_t s_bar;
    __shared__ int s_task, s_chunk;
    __shared__ int s_w[SCORE_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned FULL = 0xffffffffu;

    if (tid == 0) { mbar_init(&s_bar, 1); fence_mbar_init(); }
    __syncthreads();
    uint32_t bar_parity = 0;
#ifdef CSB_SCORE_PHASES
    long long t_prev = clock64();
#endif

    while (true) {
        if (tid == 0) {
            const int sl = atomicAdd(B.counters, 1);
            s_task = sl;
            // Host-buffer entry point: the distance maps arrive chunk by chunk on a copy stream while this kernel already runs;
            // a chunk is usable once its flag (copied right after it, same stream) carries the current epoch.
            if (B.ready_flags && sl < B.n_tasks) {
                const TaskTab& tq = B.ttab[B.task_order[sl]];
            const long long last = (long long)tq.map_offset + (long long)tq.roi_w * tq.roi_h - 1;
            int chunk = 0;
            while (chunk < B.n_chunks - 1 && B.chunk_end[chunk] <= last) chunk++;
            const volatile unsigned* fl = B.ready_flags + chunk;
                while (*fl != B.epoch) __nanosleep(256);
                __threadfence();
            }
        }
        __syncthreads();
        const int slot = s_task;
        if (slot >= B.n_tasks) break;
        SCORE_PHASE(0);  // task fetch (+ chunk wait)
        const int task = B.task_order[slot];
        const TaskTab tt = B.ttab[task];
        const FrameTab& ft = B.ftab[tt.frame_id];
        const TaskGeo geo = make_geo(tt);
        const int n_yaw = ft.n_yaw, n_pairs = ft.n_roll * ft.n_pitch;
        const int n_groups = n_pairs * n_yaw;
        const int map_floats = tt.roi_w * tt.roi_h;
        const bool map_smem = map_floats <= map_cap_floats;
        const float* gmap = B.maps + tt.map_offset;
        // a map larger than the shared-memory budget: its first rows (as many whole rows as fit) are staged, the rest is gathered from L2
        const int smem_floats = map_smem ? map_floats : (map_cap_floats / tt.roi_w) * tt.roi_w;

        // (a)
        if (smem_floats > 0 && tid == 0) {
            uint32_t bytes = map_smem ? (((uint32_t)map_floats * 4u + 15u) & ~15u) : (((uint32_t)smem_floats * 4u) & ~15u);
            fence_proxy_async();  // previous task's generic reads of s_map are ordered before the async write
            mbar_expect_tx(&s_bar, bytes);
            tma_bulk_g2s(s_map, gmap, bytes, &s_bar);
        }
        // (b) vanishing points per group + the VP-support angles k_prep_lines left in global memory
        for (int g = tid; g < n_groups; g += SCORE_THREADS) {
            int yaw_id = g % n_yaw, pair = g / n_yaw;
            double vp[6];
            vanishing_points(ft.KinvR[pair], ft.cosy[yaw_id], ft.siny[yaw_id], vp);
#pragma unroll
            for (int q = 0; q < 6; q++) s_vp[6 * g + q] = vp[q];
        }
        {
            const double* sup = B.vp_sup + (size_t)task * B.sup_stride;
            for (int i = tid; i < 6 * n_groups; i += SCORE_THREADS) s_sup[i] = sup[i];
        }
        if (tid == 0) s_chunk = 0;
        SCORE_PHASE(1);
        __syncthreads();
        SCORE_PHASE(2);  // VP support

        // (d) phase 1 -> validity bitmask
        const int n_hyp = tt.n_hyp;
        const int n_words = (n_hyp + 31) >> 5;
        // one thread per (group, top sample): corner 2 once, then both configurations; hypothesis id = 2 * pair + (cfg - 1), so a
        // warp's 32 pairs fill two mask words (bits interleaved: even = configuration 1, odd = configuration 2)
        const int n_pairs_gt = n_hyp >> 1;
        // warps fetch chunks of 32 pairs from a shared counter: the cost of a pair depends on where the cascade rejects it
        while (true) {
            int cb = 0;
            if (lane == 0) cb = atomicAdd(&s_chunk, 32);
            cb = __shfl_sync(FULL, cb, 0);
            if (cb >= n_pairs_gt) break;
            const int p = cb + lane;
            bool v1 = false, v2 = false;
            if (p < n_pairs_gt) {
                const int group = p / tt.n_top, top = p - group * tt.n_top;
                const double c1x = (double)(tt.top_x0 + top * tt.top_step);
                V2 c2;
                const int vp1 = construct_corner2(geo, s_vp + 6 * group, c1x, c2);
                if (vp1 > 0) {
                    V2 c[8];
                    if (tt.cfg_mask & 1) v1 = construct_rest(geo, s_vp + 6 * group, c1x, c2, vp1, 1, c) > 0;
                    if (tt.cfg_mask & 2) v2 = construct_rest(geo, s_vp + 6 * group, c1x, c2, vp1, 2, c) > 0;
                }
            }
            const unsigned b1 = __ballot_sync(FULL, v1), b2 = __ballot_sync(FULL, v2);
            if (lane < 2) {
                // spread 16 bits of each ballot to even / odd positions
                unsigned x = (lane == 0) ? (b1 & 0xffffu) : (b1 >> 16), y = (lane == 0) ? (b2 & 0xffffu) : (b2 >> 16);
                x = (x | (x << 8)) & 0x00ff00ffu; x = (x | (x << 4)) & 0x0f0f0f0fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
                y = (y | (y << 8)) & 0x00ff00ffu; y = (y | (y << 4)) & 0x0f0f0f0fu; y = (y | (y << 2)) & 0x33333333u; y = (y | (y << 1)) & 0x55555555u;
                const int w = (cb >> 4) + lane;  // first hypothesis of the chunk = 2 * pair -> word (2 * pair) / 32
                if (w < n_words) s_mask[w] = x | (y << 1);
            }
        }
        __syncthreads();
        SCORE_PHASE(3);  // phase 1
        // (e) exclusive prefix of the word popcounts
        int n_valid = 0;
        for (int w0 = 0; w0 < n_words; w0 += SCORE_THREADS) {
            const int w = w0 + tid;
            const int v = (w < n_words) ? __popc(s_mask[w]) : 0;
            int tot;
            const int pre = block_excl_scan<SCORE_THREADS>(v, s_w, tid, tot);
            if (w < n_words) s_wpre[w] = n_valid + pre;
            n_valid += tot;
        }
        __syncthreads();
        // (f) phase 2
        SCORE_PHASE(4);  // prefix
        if (smem_floats > 0) { mbar_wait(&s_bar, bar_parity); bar_parity ^= 1; }
        SCORE_PHASE(5);  // wait for the map
        for (int i = tid; i < n_valid; i += SCORE_THREADS) {
            // word containing the i-th set bit: last w with s_wpre[w] <= i
            int lo = 0, hi = n_words - 1;
            while (lo < hi) {
                int mid = (lo + hi + 1) >> 1;
                if (s_wpre[mid] <= i) lo = mid; else hi = mid - 1;
            }
            const int h = (lo << 5) + (int)__fns(s_mask[lo], 0, i - s_wpre[lo] + 1);
            int group, top, cfg;
            decode_hyp(h, tt.n_top, group, top, cfg);
            V2 c[8];
            construct_corners<false>(geo, s_vp + 6 * group, (double)(tt.top_x0 + top * tt.top_step), cfg, c);
            const double total_angle_diff = box_edge_alignment_angle_error(s_sup + 6 * group, c, cfg);
            // (the bulk copy of a partial map is rounded down to 16 bytes: up to three floats short of smem_floats)
            const double sum_dist = map_smem ? box_edge_sum_dists<1>(s_map, gmap, 0, tt.roi_h, tt.roi_w, c, geo.roi_l, geo.roi_t, cfg)
                                             : box_edge_sum_dists<2>(s_map, gmap, smem_floats & ~3, tt.roi_h, tt.roi_w, c, geo.roi_l, geo.roi_t, cfg);
            const size_t o = (size_t)tt.out_offset + i;
            B.p_dist[o] = sum_dist / tt.diag;
            B.p_angle[o] = total_angle_diff;
            B.p_hyp[o] = h;
        }
        if (tid == 0) B.n_valid[task] = n_valid;
        __syncthreads();
        SCORE_PHASE(6);  // phase 2
    }
    SCORE_PHASE(7);  // idle tail: exit of this CTA relative to its last task (total kernel time is the slowest CTA)
}

// ------------------------------------------------------------------------------------------------
// k_select : fuse_normalize_scores_v2 (object_3d_util.cpp:726-837), one CTA per task
// ------------------------------------------------------------------------------------------------
constexpr int SELECT_THREADS = 128;

// order-preserving map double -> uint64 (negative values reversed, sign bit flipped)
__device__ __forceinline__ unsigned long long ordered_key(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double ordered_key_inv(unsigned long long k) {
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// k-th smallest (1-based) of v[0..N) by an MSB radix select over the ordered keys (8 bits per pass; the passes end as soon as one
// candidate is left -- usually after two or three -- or after the eighth); also returns how many elements are strictly smaller.  All
// threads of the block call it; hist = 256 ints of shared memory, s_bc = 3 ints, s_key = one 64-bit word.
template <int THREADS>
__device__ __forceinline__ double radix_select_kth(const double* v, int N, int k, int* hist, int* s_bc, unsigned long long* s_key, int tid, int& n_less) {
    unsigned long long prefix = 0, mask = 0;
    int rank = k - 1, below = 0;  // 0-based rank inside the current candidate set
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        for (int b = tid; b < 256; b += THREADS) hist[b] = 0;
        __syncthreads();
        for (int i = tid; i < N; i += THREADS) {
            unsigned long long key = ordered_key(v[i]);
            if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 255ull)], 1);
        }
        __syncthreads();
        if (tid < 32) {
            // lane owns bins [8*lane, 8*lane+8): find the bin where the cumulative count passes `rank`
            int c[8], s = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) { c[q] = hist[8 * tid + q]; s += c[q]; }
            int inc = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, inc, o); if (tid >= o) inc += t; }
            int excl = inc - s;
            if (rank >= excl && rank < inc) {
                int acc = excl, bin = 0, cnt = c[7];
#pragma unroll
                for (int q = 0; q < 8; q++) { if (rank >= acc + c[q]) { acc += c[q]; } else { bin = q; cnt = c[q]; break; } }
                s_bc[0] = 8 * tid + bin;
                s_bc[1] = acc;  // elements of the candidate set below the chosen bin
                s_bc[2] = cnt;  // elements in the chosen bin
            }
        }
        __syncthreads();
        const int bin = s_bc[0], acc = s_bc[1], cnt = s_bc[2];
        prefix |= ((unsigned long long)bin) << shift;
        mask |= 255ull << shift;
        rank -= acc;
        below += acc;
        if (cnt == 1 && pass < 7) {
            // one candidate left: it is the k-th smallest, nothing else shares its remaining bits
            for (int i = tid; i < N; i += THREADS) {
                const unsigned long long key = ordered_key(v[i]);
                if ((key & mask) == prefix) *s_key = key;
            }
            __syncthreads();
            prefix = *s_key;
            __syncthreads();
            break;
        }
        __syncthreads();
    }
    n_less = below;
    return ordered_key_inv(prefix);
}

// Tasks with n_lo < N <= n_hi are handled by this launch (two launches: a small-footprint variant for the common case and a
// large one).  Working arrays live in shared memory when N <= n_cap, else in the task's global scratch.
__global__ void __launch_bounds__(SELECT_THREADS) k_select(DetectBuffers B, int n_lo, int n_hi, int n_cap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_w[SELECT_THREADS / 32];
    __shared__ double s_red[4 * (SELECT_THREADS / 32)];
    __shared__ int s_hist[256];
    __shared__ int s_bc[3];
    __shared__ unsigned long long s_key;
    const int task = blockIdx.x, tid = threadIdx.x;
    const int N = B.n_valid[task];
    if (N == 0 && n_lo == 0 && tid == 0) B.n_keep[task] = 0;  // (the first launch also covers the tasks without a valid proposal)
    if (N <= n_lo || N > n_hi) return;
    const TaskTab tt = B.ttab[task];
    const size_t ob = (size_t)tt.out_offset;
    // layout: vd[n_cap] | va[n_cap] | ibuf[n_cap] | flag[n_cap].  The partial_sort replay, when needed, aliases va (heap values)
    // and ibuf (heap indices + larger-child table): k <= 2N/3 + 1, so 8k <= 8N and 4(k + k/2) <= 4N.
    double *vd, *va;
    int* ibuf;
    unsigned char* flag;
    WarpHeap heap;
    if (N <= n_cap) {
        vd = reinterpret_cast<double*>(smem_raw);
        va = vd + n_cap;
        ibuf = reinterpret_cast<int*>(va + n_cap);
        flag = reinterpret_cast<unsigned char*>(ibuf + n_cap);
        heap.hv = va;
        for (int i = tid; i < N; i += SELECT_THREADS) { vd[i] = B.p_dist[ob + i]; va[i] = B.p_angle[ob + i]; }
    } else {
        vd = B.p_dist + ob; va = B.p_angle + ob;
        ibuf = B.sel_idx + 2 * ob;
        flag = B.sel_flag + ob;
        heap.hv = B.sel_heap + 2 * ob;
    }
    heap.hi = ibuf;
    heap.big = nullptr;
    __syncthreads();

    int* keep = B.keep + ob;
    int n_keep = 0;
    if (N > 4) {
        // fuse_normalize_scores_v2, object_3d_util.cpp:736-787
        const int k = (int)round((double)(float)N / 3.0 * 2.0);  // breaking_num (:739)
        // angle list: sorted[k-1] > sorted[k-2] (:766)  <=>  exactly k-1 elements are strictly below the k-th smallest value;
        // then the kept k-1 are precisely those elements
        int lessA;
        const double vkA = radix_select_kth<SELECT_THREADS>(va, N, k, s_hist, s_bc, &s_key, tid, lessA);
        const bool angle_active = (lessA == k - 1);
        // distance list: dist_keep = first k-1 of std::partial_sort(iota, iota+k, end) (matrix_utils.cpp:327-335, unstable).
        int lessD;
        const double vk = radix_select_kth<SELECT_THREADS>(vd, N, k, s_hist, s_bc, &s_key, tid, lessD);
        // With the angle filter on, dist_keep is used as a SET = heap minus its top after __heap_select.  If only one element of
        // value vk is inside the heap (lessD == k-1) it is the top (the dropped k-th) and the set is exactly {d < vk}; otherwise
        // which of the tied elements stay depends on the heap -> replay.  With the angle filter off the kept list keeps
        // partial_sort's ORDER (:783-786): replay (it also yields the sorted order).
        bool need_emul = angle_active ? (lessD != k - 1) : true;
        if (angle_active) {
            for (int i = tid; i < N; i += SELECT_THREADS) flag[i] = ((va[i] < vkA) ? 1 : 0) | ((!need_emul && vd[i] < vk) ? 2 : 0);
        }
        __syncthreads();
        // Angle filter off, no tie at the boundary (exactly k - 1 distances below the k-th smallest) and no two equal values among them:
        // std::partial_sort's prefix is then the ONE ascending order of those k - 1 elements, whatever the heap does -- sort them with the
        // whole CTA (bitonic, keys in the angle array's storage) instead of replaying k sequential heap pops on one warp.  Any duplicate
        // key sends the task to the literal replay below.
        if (!angle_active && lessD == k - 1 && N <= n_cap) {
            int m = 1;
            while (m < k - 1) m <<= 1;
            if (m <= n_cap) {
                unsigned long long* keys = reinterpret_cast<unsigned long long*>(va);
                if (tid == 0) s_bc[0] = 0;
                __syncthreads();
                const unsigned long long kk = ordered_key(vk);  // the same order the radix select counted in: exactly lessD = k - 1 keys are below
                for (int i = tid; i < N; i += SELECT_THREADS) {
                    const unsigned long long key = ordered_key(vd[i]);
                    if (key < kk) { const int pos = atomicAdd(&s_bc[0], 1); keys[pos] = key; ibuf[pos] = i; }
                }
                for (int i = k - 1 + tid; i < m; i += SELECT_THREADS) { keys[i] = ~0ull; ibuf[i] = -1; }
                __syncthreads();
                for (int size = 2; size <= m; size <<= 1)
                    for (int stride = size >> 1; stride > 0; stride >>= 1) {
                        for (int t = tid; t < (m >> 1); t += SELECT_THREADS) {
                            const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1)), hi = lo | stride;
                            const bool up = (lo & size) == 0;
                            const unsigned long long a = keys[lo], b = keys[hi];
                            if ((a > b) == up) { keys[lo] = b; keys[hi] = a; const int ti = ibuf[lo]; ibuf[lo] = ibuf[hi]; ibuf[hi] = ti; }
                        }
                        __syncthreads();
                    }
                int dup = 0;
                for (int p = tid; p + 1 < k - 1; p += SELECT_THREADS) dup |= (keys[p] == keys[p + 1]) ? 1 : 0;
                if (!__syncthreads_or(dup)) {
                    for (int p = tid; p < k - 1; p += SELECT_THREADS) keep[p] = ibuf[p];
                    need_emul = false;
                }
                __syncthreads();
            }
        }
        if (need_emul) {
            // literal libstdc++ __heap_select (+ __sort_heap), warp-cooperative (proposal_dev.cuh); va / ibuf provide the storage
            heap.big = heap.hi + k;
            if (tid < 32) wh_partial_sort(heap, vd, k, N, !angle_active, tid);
            __syncthreads();
            // after __heap_select the k-th (excluded) element is the heap top, hi[0]; after __sort_heap it is hi[k-1]
            if (angle_active) { for (int p = tid + 1; p < k; p += SELECT_THREADS) flag[heap.hi[p]] |= 2; }
            else { for (int p = tid; p < k - 1; p += SELECT_THREADS) keep[p] = heap.hi[p]; }
            __syncthreads();
        }
        if (angle_active) {
            // std::set_intersection of the two index-sorted sets == ascending indices carrying both flags (:773-781)
            for (int b0 = 0; b0 < N; b0 += SELECT_THREADS) {
                int i = b0 + tid;
                int v = (i < N && flag[i] == 3) ? 1 : 0;
                int tot;
                int pre = block_excl_scan<SELECT_THREADS>(v, s_w, tid, tot);
                if (v) keep[n_keep + pre] = i;
                n_keep += tot;
            }
        } else
            n_keep = k - 1;
    } else {
        n_keep = N;
        for (int p = tid; p < N; p += SELECT_THREADS) keep[p] = p;
    }
    __syncthreads();

    // min / max of the kept errors (:798-817) and combined score (:820-836); angle values re-read from global (va may be clobbered)
    const double* ga = B.p_angle + ob;
    double mn_d = 1e300, mx_d = -1e300, mn_a = 1e300, mx_a = -1e300;
    for (int j = tid; j < n_keep; j += SELECT_THREADS) {
        double d = vd[keep[j]], a = ga[keep[j]];
        mn_d = cmin(mn_d, d); mx_d = cmax(mx_d, d); mn_a = cmin(mn_a, a); mx_a = cmax(mx_a, a);
    }
    {
        const unsigned FULL = 0xffffffffu;
#pragma unroll
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
        double d = vd[keep[j]], a = ga[keep[j]];
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
}

// ------------------------------------------------------------------------------------------------
// k_recover : 3D recovery of every kept proposal + skew-augmented score (box_proposal_detail.cpp:723-822)
// ------------------------------------------------------------------------------------------------
// Recompute corners + 3D object for proposal (task tt, hypothesis h).  Returns vp_1_position.
__device__ __forceinline__ int recover_object(const TaskTab& tt, const FrameTab& ft, int h, V2* c, Obj3D& o, int& cfg, int& group) {
    int top;
    decode_hyp(h, tt.n_top, group, top, cfg);
    int yaw_id = group % ft.n_yaw, pair = group / ft.n_yaw;
    double vp[6];
    vanishing_points(ft.KinvR[pair], ft.cosy[yaw_id], ft.siny[yaw_id], vp);
    TaskGeo geo = make_geo(tt);
    int vp1 = construct_corners<false>(geo, vp, (double)(tt.top_x0 + top * tt.top_step), cfg, c);
    corners_to_3d(c, ft.Tnew[pair], ft.invK, o);
    return vp1;
}

constexpr int RECOVER_THREADS = 128;
// grid (n_tasks, RECOVER_Y): blocks of a task stride over its kept list.  cand_score[j] / cand_ok[j] are indexed by the
// position j in the kept list (raw_obj_proposals = the ok ones, in list order).
__global__ void __launch_bounds__(RECOVER_THREADS) k_recover(DetectBuffers B) {
    const int task = blockIdx.x;
    const int n_keep = B.n_keep[task];
    const int j0 = blockIdx.y * RECOVER_THREADS + threadIdx.x;
    if (j0 >= n_keep) return;
    const TaskTab tt = B.ttab[task];
    const FrameTab& ft = B.ftab[tt.frame_id];
    const size_t ob = (size_t)tt.out_offset;
    const double weight_skew_error = 1.5;
    for (int j = j0; j < n_keep; j += gridDim.y * RECOVER_THREADS) {
        int h = B.p_hyp[ob + B.keep[ob + j]];
        V2 c[8];
        Obj3D o;
        int cfg, group;
        recover_object(tt, ft, h, c, o, cfg, group);
        unsigned char ok = 0;
        double score = 0;
        if (!((o.scale[0] < 0) || (o.scale[1] < 0) || (o.scale[2] < 0))) {  // (:766)
            ok = 1;
            double skew_ratio = cmax(o.scale[0], o.scale[1]) / cmin(o.scale[0], o.scale[1]);
            double skew_error = weight_skew_error * cmax(skew_ratio - B.dc.nominal_skew_ratio, 0.0);
            if (skew_ratio > B.dc.max_cut_skew) skew_error = 100;
            score = B.norm_score[ob + j] + weight_skew_error * skew_error;  // (:813-820)
        }
        B.cand_score[ob + j] = score;
        B.cand_ok[ob + j] = ok;
    }
}

// ------------------------------------------------------------------------------------------------
// k_rank : final ranking per 2D box (box_proposal_detail.cpp:804-838), one warp per box
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_rank(DetectBuffers B) {
    const int box = blockIdx.x, lane = threadIdx.x;
    const unsigned FULL = 0xffffffffu;
    const int t0 = B.box_task_begin[box], t1 = B.box_task_begin[box + 1];
    const int kmax = B.dc.max_cuboid_num;
    // candidates of the box = kept proposals with ok flag, in (task, kept position) order; "slot" = position in the
    // concatenated kept lists, "rank index" = position among the ok ones (index into raw_obj_proposals)
    int S = 0;
    for (int t = t0; t < t1; t++) S += B.n_keep[t];
    auto locate = [&](int slot, int& t, int& j) {
        t = t0;
        while (slot >= B.n_keep[t]) { slot -= B.n_keep[t]; t++; }
        j = slot;
    };
    int* ridx = (t1 > t0) ? B.rank_idx + 2 * (size_t)B.ttab[t0].out_offset : nullptr;  // scratch: 2 x the box's capacity (ridx[M] | win[M], M <= S <= capacity)
    // compact the ok slots (ordered) and count them
    int M = 0;
    for (int s0 = 0; s0 < S; s0 += 32) {
        int s = s0 + lane;
        bool ok = false;
        if (s < S) { int t, j; locate(s, t, j); ok = B.cand_ok[(size_t)B.ttab[t].out_offset + j] != 0; }
        unsigned bal = __ballot_sync(FULL, ok);
        if (ok) ridx[M + __popc(bal & ((1u << lane) - 1))] = s;
        M += __popc(bal);
    }
    __syncwarp();
    const int k = min(kmax, M);
    if (lane == 0) B.n_cuboids[box] = k;
    if (k == 0) return;
    auto score_of_slot = [&](int s) { int t, j; locate(s, t, j); return B.cand_score[(size_t)B.ttab[t].out_offset + j]; };
    // winners[w] = rank index (position in ridx) of the w-th best
    int* win = ridx + M;  // scratch behind the compacted list
    if (k == 1) {
        // partial_sort(idx, idx+1, end): start with element 0, replace by every later strictly smaller one.  That is the first
        // index of the minimum over the non-NaN scores -- unless score[0] is NaN, which nothing can replace.
        double best = 0; int bi = -1;
        for (int r = lane; r < M; r += 32) {
            double sc = score_of_slot(ridx[r]);
            if (sc == sc && (bi < 0 || sc < best)) { best = sc; bi = r; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_down_sync(FULL, best, o); int oi = __shfl_down_sync(FULL, bi, o);
            if (oi >= 0 && (bi < 0 || ov < best || (ov == best && oi < bi))) { best = ov; bi = oi; }
        }
        bi = __shfl_sync(FULL, bi, 0);
        double s0v = score_of_slot(ridx[0]);
        if (s0v != s0v || bi < 0) bi = 0;
        if (lane == 0) win[0] = bi;
    } else {
        for (int r = lane; r < M; r += 32) win[r] = r;
        __syncwarp();
        if (lane == 0) {
            auto less = [&](int a, int b) { return score_of_slot(ridx[a]) < score_of_slot(ridx[b]); };
            heap_select(win, k, M, less);
            heap_sort(win, k, less);
        }
    }
    __syncwarp();
    // write the k cuboid records (lanes over winners)
    for (int w = lane; w < k; w += 32) {
        const int r = win[w];
        int t, j;
        locate(ridx[r], t, j);
        const TaskTab tt = B.ttab[t];
        const FrameTab& ft = B.ftab[tt.frame_id];
        const size_t ob = (size_t)tt.out_offset;
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
// parity probe of det_atan2_x6: thread t evaluates elements [6t, 6t + 6) the way box_edge_alignment_angle_error does
__global__ void k_debug_atan2(const double* y, const double* x, double* out, int n6, int* n_fallback) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n6) return;
    double yy[6], xx[6], o[6];
#pragma unroll
    for (int i = 0; i < 6; i++) { yy[i] = y[6 * t + i]; xx[i] = x[6 * t + i]; }
    if (!det_atan2_x6(yy, xx, o)) {
        atomicAdd(n_fallback, 1);
#pragma unroll
        for (int i = 0; i < 6; i++) o[i] = det_atan2(yy[i], xx[i]);
    }
#pragma unroll
    for (int i = 0; i < 6; i++) out[6 * t + i] = o[i];
}
cudaError_t launch_debug_atan2(const double* y, const double* x, double* out, int n6, int* n_fallback, cudaStream_t st) {
    if (n6 == 0) return cudaSuccess;
    k_debug_atan2<<<(n6 + 127) / 128, 128, 0, st>>>(y, x, out, n6, n_fallback);
    return cudaGetLastError();
}

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
static size_t score_smem_bytes(int groups_cap, int map_cap_floats, int words_cap) {
    return (size_t)map_cap_floats * 4 + (size_t)groups_cap * 12 * 8 + (size_t)(2 * words_cap + 1) * 4 + 64;
}

cudaError_t launch_prep_lines(const DetectBuffers& B, int max_lines_per_frame, int max_groups, cudaStream_t st) {
    int cap = max_lines_per_frame < 1 ? 1 : max_lines_per_frame;
    size_t smem = (size_t)cap * (5 * 8 + 3 * 4) + 16;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_prep_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_prep_lines<<<B.n_tasks, PREP_THREADS, smem, st>>>(B, cap);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t smem2 = (size_t)cap * (5 * 8 + 4 * 4) + 16;
    e = cudaFuncSetAttribute(k_vp_support, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
    if (e != cudaSuccess) return e;
    const int n_units_max = 2 * max_groups + MAX_RP * MAX_RP;
    dim3 grid((n_units_max + VPS_THREADS - 1) / VPS_THREADS, B.n_tasks);
    k_vp_support<<<grid, VPS_THREADS, smem2, st>>>(B, cap);
    return cudaGetLastError();
}

cudaError_t launch_score(const DetectBuffers& B, int max_groups, int max_hyp_per_task, int num_sms, int max_smem_optin, int* map_cap_floats_out, cudaStream_t st) {
    const int groups_cap = max_groups;
    const int words_cap = (max_hyp_per_task + 31) / 32 + 1;
    size_t fixed = score_smem_bytes(groups_cap, 0, words_cap);
    // static __shared__ + slack; with two CTAs per SM each gets half of the SM's 228 KB (1 KB per CTA is reserved by the system)
    size_t budget = SCORE_CTAS_PER_SM == 1 ? (size_t)max_smem_optin - 1024 : (size_t)(228 * 1024) / SCORE_CTAS_PER_SM - 2048;
    if (fixed + 16 * 1024 > budget) return cudaErrorInvalidValue;
    int map_cap = (int)((budget - fixed) / 4) & ~31;
    size_t smem = score_smem_bytes(groups_cap, map_cap, words_cap);
    cudaError_t e = cudaFuncSetAttribute(k_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (map_cap_floats_out) *map_cap_floats_out = map_cap;
    int grid = B.n_tasks < SCORE_CTAS_PER_SM * num_sms ? B.n_tasks : SCORE_CTAS_PER_SM * num_sms;
    k_score<<<grid, SCORE_THREADS, smem, st>>>(B, groups_cap, map_cap, words_cap);
    return cudaGetLastError();
}

// per proposal: 2 doubles + idx (power-of-two padded ints) + 1 flag byte
static size_t select_smem_bytes(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return (size_t)n * 16 + (size_t)p * 4 + (size_t)n + 16;
}

cudaError_t launch_select(const DetectBuffers& B, int max_hyp_per_task, int max_smem_optin, cudaStream_t st, int* n_launches) {
    int L = 0;
    // small-footprint variant: tasks with up to 2048 valid proposals (several CTAs per SM)
    const int small_cap = 2048;
    cudaError_t e = cudaFuncSetAttribute(k_select, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin - 2048);
    if (e != cudaSuccess) return e;
    k_select<<<B.n_tasks, SELECT_THREADS, select_smem_bytes(small_cap), st>>>(B, 0, small_cap, small_cap);
    L++;
    if (max_hyp_per_task > small_cap) {
        int big_cap = max_hyp_per_task > 8192 ? 8192 : max_hyp_per_task;
        while (select_smem_bytes(big_cap) > (size_t)max_smem_optin - 4096) big_cap -= 256;
        k_select<<<B.n_tasks, SELECT_THREADS, select_smem_bytes(big_cap), st>>>(B, small_cap, 0x7fffffff, big_cap);
        L++;
    }
    if (n_launches) *n_launches = L;
    return cudaGetLastError();
}

cudaError_t launch_recover(const DetectBuffers& B, cudaStream_t st) {
    dim3 grid(B.n_tasks, 8);
    k_recover<<<grid, RECOVER_THREADS, 0, st>>>(B);
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
