// proposal_score.cuh -- per-proposal scoring device functions (distance-map gathers, edge-angle error).
#pragma once
#include "proposal_dev.cuh"

namespace csb {

__constant__ double c_t[11] = {0 / 10.0, 1 / 10.0, 2 / 10.0, 3 / 10.0, 4 / 10.0, 5 / 10.0, 6 / 10.0, 7 / 10.0, 8 / 10.0, 9 / 10.0, 10 / 10.0};
__constant__ double c_1mt[11] = {1 - 0 / 10.0, 1 - 1 / 10.0, 1 - 2 / 10.0, 1 - 3 / 10.0, 1 - 4 / 10.0, 1 - 5 / 10.0,
                                 1 - 6 / 10.0, 1 - 7 / 10.0, 1 - 8 / 10.0, 1 - 9 / 10.0, 1 - 10 / 10.0};

// 11 samples along one box edge, object_3d_util.cpp:642-664.  WEIGHT: 0 none, 1 x1.5 (edges 4,5 of config 2), 2 x2 (edge 6)
template <bool SMEM, int WEIGHT>
__device__ __forceinline__ float edge_samples(float sum_dist, const float* __restrict__ map, int cols, int last, V2 c1, V2 c2) {
#pragma unroll 4
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
// c: corners in image coordinates; (ox, oy) = ROI origin.  The reference shifts all eight corners first (box_proposal_detail.cpp:634-636);
// shifting the two corners of an edge when the edge is sampled gives the same values and keeps only one corner set live.
template <bool SMEM>
__device__ __forceinline__ double box_edge_sum_dists(const float* __restrict__ map, int rows, int cols, const V2* cc, double ox, double oy, int config_id) {
    const int last = rows * cols - 1;
    float s = 0;
    struct Sh { const V2* c; double ox, oy; __device__ __forceinline__ V2 operator[](int i) const { return V2{c[i].x - ox, c[i].y - oy}; } } c{cc, ox, oy};
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
    double box_edge_angle = normalize_to_pi(det_atan2(b.y - a.y, b.x - a.x));
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

}  // namespace csb
