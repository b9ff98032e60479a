// proposal_score.cuh -- per-proposal scoring device functions (distance-map gathers, edge-angle error).
#pragma once
#include "proposal_dev.cuh"

namespace csb {

__constant__ double c_t[11] = {0 / 10.0, 1 / 10.0, 2 / 10.0, 3 / 10.0, 4 / 10.0, 5 / 10.0, 6 / 10.0, 7 / 10.0, 8 / 10.0, 9 / 10.0, 10 / 10.0};
__constant__ double c_1mt[11] = {1 - 0 / 10.0, 1 - 1 / 10.0, 1 - 2 / 10.0, 1 - 3 / 10.0, 1 - 4 / 10.0, 1 - 5 / 10.0,
                                 1 - 6 / 10.0, 1 - 7 / 10.0, 1 - 8 / 10.0, 1 - 9 / 10.0, 1 - 10 / 10.0};

// 11 samples along one box edge, object_3d_util.cpp:642-664.  WEIGHT: 0 none, 1 x1.5 (edges 4,5 of config 2), 2 x2 (edge 6).
// MODE: 1 = the whole distance map is in shared memory; 2 = its first n_smem floats (whole rows) are, the rest is gathered from
// global memory (maps larger than the shared-memory budget); 0 = global memory only.
template <int MODE, int WEIGHT>
__device__ __forceinline__ float edge_samples(float sum_dist, const float* __restrict__ smap, const float* __restrict__ gmap, int n_smem, int cols, int last,
                                              V2 c1, V2 c2) {
#pragma unroll 4
    for (int k = 0; k < 11; k++) {
        double sx = c_t[k] * c1.x + c_1mt[k] * c2.x;
        double sy = c_t[k] * c1.y + c_1mt[k] * c2.y;
        int li = __double2int_rz(sy) * cols + __double2int_rz(sx);
        li = max(0, min(li, last));  // defined behaviour for samples on the ROI's right/bottom bound (reference: UB)
        float d1;
        if (MODE == 1) d1 = smap[li];
        else if (MODE == 2) d1 = (li < n_smem) ? smap[li] : __ldg(gmap + li);
        else d1 = __ldg(gmap + li);
        if (WEIGHT == 1) d1 = (float)((double)d1 * 3.0 / 2.0);
        if (WEIGHT == 2) d1 = (float)((double)d1 * 2.0);
        sum_dist = sum_dist + d1;
    }
    return sum_dist;
}

// box_edge_sum_dists, object_3d_util.cpp:622-667 with the visible-edge tables of box_proposal_detail.cpp:646, 663
// c: corners in image coordinates; (ox, oy) = ROI origin.  The reference shifts all eight corners first (box_proposal_detail.cpp:634-636);
// shifting the two corners of an edge when the edge is sampled gives the same values and keeps only one corner set live.
template <int MODE>
__device__ __forceinline__ double box_edge_sum_dists(const float* __restrict__ smap, const float* __restrict__ gmap, int n_smem, int rows, int cols, const V2* cc,
                                                     double ox, double oy, int config_id) {
    const int last = rows * cols - 1;
    float s = 0;
    struct Sh { const V2* c; double ox, oy; __device__ __forceinline__ V2 operator[](int i) const { return V2{c[i].x - ox, c[i].y - oy}; } } c{cc, ox, oy};
#define CSB_EDGE(W, A, B) s = edge_samples<MODE, W>(s, smap, gmap, n_smem, cols, last, c[A], c[B]);
    CSB_EDGE(0, 0, 1) CSB_EDGE(0, 1, 2) CSB_EDGE(0, 2, 3) CSB_EDGE(0, 3, 0)
    if (config_id == 1) {
        CSB_EDGE(0, 1, 5) CSB_EDGE(0, 2, 4) CSB_EDGE(0, 3, 7) CSB_EDGE(0, 4, 7) CSB_EDGE(0, 4, 5)
    } else {
        CSB_EDGE(1, 1, 5) CSB_EDGE(1, 2, 4) CSB_EDGE(2, 4, 5)
    }
#undef CSB_EDGE
    return (double)s;
}

// one box edge (its angle already evaluated) against the (<=2) supporting line angles of its VP, object_3d_util.cpp:696-715
__device__ __forceinline__ double edge_angle_diff(double raw_atan2, double v0, double v1) {
    double box_edge_angle = normalize_to_pi(raw_atan2);
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
// box_edge_alignment_angle_error, object_3d_util.cpp:670-723 with the tables of box_proposal_detail.cpp:651, 665.
// The six edge angles are evaluated together (det_atan2_x6: six independent chains instead of six calls one after the other);
// operands on one of det_atan2's special paths make the whole proposal take the scalar route.
__device__ __forceinline__ double box_edge_alignment_angle_error(const double* sup /*6*/, const V2* c, int config_id) {
    const double not_found_penalty = 30.0 / 180.0 * M_PI * 2;
    // VP 1: edges (1,2) and (8,5) | (3,4);  VP 2: edges (4,1) and (5,6);  VP 3: edges (4,8)|(3,5) and (2,6)
    double dy[6], dx[6], ang[6];
    dy[0] = c[1].y - c[0].y; dx[0] = c[1].x - c[0].x;
    dy[1] = config_id == 1 ? c[4].y - c[7].y : c[3].y - c[2].y; dx[1] = config_id == 1 ? c[4].x - c[7].x : c[3].x - c[2].x;
    dy[2] = c[0].y - c[3].y; dx[2] = c[0].x - c[3].x;
    dy[3] = c[5].y - c[4].y; dx[3] = c[5].x - c[4].x;
    dy[4] = config_id == 1 ? c[7].y - c[3].y : c[4].y - c[2].y; dx[4] = config_id == 1 ? c[7].x - c[3].x : c[4].x - c[2].x;
    dy[5] = c[5].y - c[1].y; dx[5] = c[5].x - c[1].x;
    if (!det_atan2_x6(dy, dx, ang)) {
#pragma unroll
        for (int i = 0; i < 6; i++) ang[i] = det_atan2(dy[i], dx[i]);
    }
    double total = 0;
    if (!isnan(sup[0]) || !isnan(sup[1])) {
        total = total + edge_angle_diff(ang[0], sup[0], sup[1]);
        total = total + edge_angle_diff(ang[1], sup[0], sup[1]);
    } else
        total = total + not_found_penalty;
    if (!isnan(sup[2]) || !isnan(sup[3])) {
        total = total + edge_angle_diff(ang[2], sup[2], sup[3]);
        total = total + edge_angle_diff(ang[3], sup[2], sup[3]);
    } else
        total = total + not_found_penalty;
    if (!isnan(sup[4]) || !isnan(sup[5])) {
        total = total + edge_angle_diff(ang[4], sup[4], sup[5]);
        total = total + edge_angle_diff(ang[5], sup[4], sup[5]);
    } else
        total = total + not_found_penalty;
    return total;
}

}  // namespace csb
