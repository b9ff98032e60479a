// proposal_dev.cuh -- device functions of the proposal half (geometry, scoring, 3D recovery, heap emulation).
//
// Each function states the reference lines it implements (paths relative to detect_3d_cuboid/src/).
// Arithmetic is written in the same order as the reference so that, with -fmad=false, the FP64 results are
// the same IEEE sequence a scalar x86 build produces.
#pragma once
#include "csb_internal.h"
#include "csb_math.cuh"

namespace csb {

struct TaskGeo {
    double left, top, right, down;          // left_x_raw, top_y_raw, right_x_raw, down_y_expan
    double roi_l, roi_t, roi_r, roi_d;      // expan_distmap bounds (inclusive)
};

__device__ __forceinline__ TaskGeo make_geo(const TaskTab& t) {
    TaskGeo g;
    g.left = (double)t.left_x_raw; g.top = (double)t.top_y_raw; g.right = (double)t.right_x_raw; g.down = (double)t.down_y_expan;
    g.roi_l = (double)t.roi_left; g.roi_t = (double)t.roi_top; g.roi_r = (double)t.roi_right; g.roi_d = (double)t.roi_down;
    return g;
}

// object_3d_util.cpp:239-242
__device__ __forceinline__ bool inside_box(V2 p, double l, double t, double r, double d) { return l <= p.x && p.x <= r && t <= p.y && p.y <= d; }

// `norm(a - b) < shorted_edge_thre` (20 px) of the rejection cascade without the square root: sqrt is correctly rounded and monotone, and the
// largest double whose rounded root is below 20 is 400 - 2^-43 (the root of 400 - 2^-44 already rounds up to 20), so
// sqrt(s) < 20  <=>  s < 400 - 2^-44 for every double s -- the same decisions as the reference (checked on the host around 400 and on
// 3e6 random values; k_score: 0.183 -> 0.168 ms).
__device__ __forceinline__ bool shorter_than_20(V2 a, V2 b) {
    const double dx = a.x - b.x, dy = a.y - b.y;
    return (dx * dx + dy * dy) < __longlong_as_double(0x4078FFFFFFFFFFFFLL);
}
// object_3d_util.cpp:309-353
__device__ __forceinline__ V2 seg_hit_boundary(V2 ps, V2 pe, double bx0, double by0, double bx1, double by1) {
    V2 direc = sub(pe, ps);
    V2 hit{-1.0, -1.0};
    if (by0 == by1) {
        double lambd = (by0 - ps.y) / direc.y;
        if (lambd >= 0) {
            V2 tmp{ps.x + lambd * direc.x, ps.y + lambd * direc.y};
            if ((bx0 <= tmp.x) && (tmp.x <= bx1)) { hit = tmp; hit.y = by0; }
        }
    }
    if (bx0 == bx1) {
        double lambd = (bx0 - ps.x) / direc.x;
        if (lambd >= 0) {
            V2 tmp{ps.x + lambd * direc.x, ps.y + lambd * direc.y};
            if ((by0 <= tmp.y) && (tmp.y <= by1)) { hit = tmp; hit.x = bx0; }
        }
    }
    return hit;
}

// object_3d_util.cpp:357-382 with infinite_line == true (u_b is dead there)
__device__ __forceinline__ V2 line_intersect(V2 p1s, V2 p1e, V2 p2s, V2 p2e) {
    double X2_X1 = p1e.x - p1s.x, Y2_Y1 = p1e.y - p1s.y;
    double X4_X3 = p2e.x - p2s.x, Y4_Y3 = p2e.y - p2s.y;
    double X1_X3 = p1s.x - p2s.x, Y1_Y3 = p1s.y - p2s.y;
    double u_a = (X4_X3 * Y1_Y3 - Y4_Y3 * X1_X3) / (Y4_Y3 * X2_X1 - X4_X3 * Y2_Y1);
    double INT_X = p1s.x + X2_X1 * u_a;
    double INT_Y = p1s.y + Y2_Y1 * u_a;
    return {INT_X * 1.0, INT_Y * 1.0};
}

// getVanishingPoints, object_3d_util.cpp:928-937.  KinvR row-major; c,s = cos/sin(yaw) from the host table.
__device__ __forceinline__ void vanishing_points(const double* K, double c, double s, double* vp /*6: x1 y1 x2 y2 x3 y3*/) {
    double ax = (K[0] * c + K[1] * s) + K[2] * 0.0, ay = (K[3] * c + K[4] * s) + K[5] * 0.0, az = (K[6] * c + K[7] * s) + K[8] * 0.0;
    double ms = -s;
    double bx = (K[0] * ms + K[1] * c) + K[2] * 0.0, by = (K[3] * ms + K[4] * c) + K[5] * 0.0, bz = (K[6] * ms + K[7] * c) + K[8] * 0.0;
    double cx = (K[0] * 0.0 + K[1] * 0.0) + K[2] * 1.0, cy = (K[3] * 0.0 + K[4] * 0.0) + K[5] * 1.0, cz = (K[6] * 0.0 + K[7] * 0.0) + K[8] * 1.0;
    vp[0] = ax / az; vp[1] = ay / az; vp[2] = bx / bz; vp[3] = by / bz; vp[4] = cx / cz; vp[5] = cy / cz;
}

// Corner construction + rejection cascade, box_proposal_detail.cpp:413-625.
// Returns vp_1_position (1 left, 2 right) or 0 if the hypothesis is rejected.  c[0..7] = corners 1..8.
// Split in two: corner 2 (box_proposal_detail.cpp:413-461) does not depend on the configuration, so the sweep computes it once
// per (group, top sample) and runs the rest (:467-625) for both configurations.
// CHECK = false recomputes the corners of a hypothesis that is already known to pass: same arithmetic, rejection tests dropped.
template <bool CHECK = true>
__device__ __forceinline__ int construct_corner2(const TaskGeo& g, const double* vp, double c1x, V2& corner_2_top) {
    V2 vp_1{vp[0], vp[1]};
    V2 corner_1_top{c1x, g.top};
    int vp_1_position = 0;
    corner_2_top = seg_hit_boundary(vp_1, corner_1_top, g.right, g.top, g.right, g.down);
    if (corner_2_top.x == -1) {
        corner_2_top = seg_hit_boundary(vp_1, corner_1_top, g.left, g.top, g.left, g.down);
        if (corner_2_top.x != -1) vp_1_position = 2;
    } else
        vp_1_position = 1;
    if (CHECK && (!(vp_1_position > 0))) return 0;
    if (CHECK && shorter_than_20(corner_1_top, corner_2_top)) return 0;
    return vp_1_position;
}

template <bool CHECK = true>
__device__ __forceinline__ int construct_rest(const TaskGeo& g, const double* vp, double c1x, V2 corner_2_top, int vp_1_position, int config_id, V2* c);

template <bool CHECK = true>
__device__ __forceinline__ int construct_corners(const TaskGeo& g, const double* vp, double c1x, int config_id, V2* c) {
    V2 corner_2_top;
    int vp_1_position = construct_corner2<CHECK>(g, vp, c1x, corner_2_top);
    if (CHECK && (vp_1_position == 0)) return 0;
    return construct_rest<CHECK>(g, vp, c1x, corner_2_top, vp_1_position, config_id, c);
}

template <bool CHECK>
__device__ __forceinline__ int construct_rest(const TaskGeo& g, const double* vp, double c1x, V2 corner_2_top, int vp_1_position, int config_id, V2* c) {
    V2 vp_1{vp[0], vp[1]}, vp_2{vp[2], vp[3]}, vp_3{vp[4], vp[5]};
    V2 corner_1_top{c1x, g.top};
    V2 corner_3_top, corner_4_top;
    if (config_id == 1) {
        if (vp_1_position == 1) corner_4_top = seg_hit_boundary(vp_2, corner_1_top, g.left, g.top, g.left, g.down);
        else corner_4_top = seg_hit_boundary(vp_2, corner_1_top, g.right, g.top, g.right, g.down);
        if (CHECK && (corner_4_top.y == -1)) return 0;
        if (CHECK && shorter_than_20(corner_1_top, corner_4_top)) return 0;
        corner_3_top = line_intersect(vp_2, corner_2_top, vp_1, corner_4_top);
        if (CHECK && (!inside_box(corner_3_top, g.left, g.top, g.right, g.down))) return 0;
        if (CHECK && (shorter_than_20(corner_3_top, corner_4_top) || shorter_than_20(corner_3_top, corner_2_top))) return 0;
    } else {
        if (vp_1_position == 1) corner_3_top = seg_hit_boundary(vp_2, corner_2_top, g.left, g.top, g.left, g.down);
        else corner_3_top = seg_hit_boundary(vp_2, corner_2_top, g.right, g.top, g.right, g.down);
        if (CHECK && (corner_3_top.y == -1)) return 0;
        if (CHECK && shorter_than_20(corner_2_top, corner_3_top)) return 0;
        corner_4_top = line_intersect(vp_1, corner_3_top, vp_2, corner_1_top);
        if (CHECK && (!inside_box(corner_4_top, g.left, g.roi_t, g.right, g.roi_d))) return 0;  // sic: x from the raw box, y from the ROI (:558)
        if (CHECK && (shorter_than_20(corner_3_top, corner_4_top) || shorter_than_20(corner_4_top, corner_1_top))) return 0;
    }
    V2 corner_5_down = seg_hit_boundary(vp_3, corner_3_top, g.left, g.down, g.right, g.down);
    if (CHECK && (corner_5_down.y == -1)) return 0;
    if (CHECK && shorter_than_20(corner_3_top, corner_5_down)) return 0;
    V2 corner_6_down = line_intersect(vp_2, corner_5_down, vp_3, corner_2_top);
    if (CHECK && (!inside_box(corner_6_down, g.roi_l, g.roi_t, g.roi_r, g.roi_d))) return 0;
    if (CHECK && (shorter_than_20(corner_6_down, corner_2_top) || shorter_than_20(corner_6_down, corner_5_down))) return 0;
    V2 corner_7_down = line_intersect(vp_1, corner_6_down, vp_3, corner_1_top);
    if (CHECK && (!inside_box(corner_7_down, g.roi_l, g.roi_t, g.roi_r, g.roi_d))) return 0;
    if (CHECK && (shorter_than_20(corner_7_down, corner_1_top) || shorter_than_20(corner_7_down, corner_6_down))) return 0;
    V2 corner_8_down = line_intersect(vp_1, corner_5_down, vp_2, corner_7_down);
    if (CHECK && (!inside_box(corner_8_down, g.roi_l, g.roi_t, g.roi_r, g.roi_d))) return 0;
    if (CHECK && (shorter_than_20(corner_8_down, corner_4_top) || shorter_than_20(corner_8_down, corner_5_down) ||
                  shorter_than_20(corner_8_down, corner_7_down)))
        return 0;
    c[0] = corner_1_top; c[1] = corner_2_top; c[2] = corner_3_top; c[3] = corner_4_top;
    c[4] = corner_5_down; c[5] = corner_6_down; c[6] = corner_7_down; c[7] = corner_8_down;
    return vp_1_position;
}

// hypothesis id -> (group, top, config).  id = (group * n_top + top) * 2 + (config - 1)
__device__ __forceinline__ void decode_hyp(int h, int n_top, int& group, int& top, int& cfg) {
    cfg = (h & 1) + 1;
    int r = h >> 1;
    group = r / n_top;
    top = r - group * n_top;
}

// ---- 3D recovery -------------------------------------------------------------------------------
// plane_hits_3d (object_3d_util.cpp:853-875) for one pixel; T = top 3 rows of transToWolrd (3x4 row-major)
__device__ __forceinline__ V3 plane_hit_3d(const double* T, const double* invK, const double* plane, V2 px) {
    double rx = (invK[0] * px.x + invK[1] * px.y) + invK[2] * 1.0;
    double ry = (invK[3] * px.x + invK[4] * px.y) + invK[5] * 1.0;
    double rz = (invK[6] * px.x + invK[7] * px.y) + invK[8] * 1.0;
    double denom = (plane[0] * rx + plane[1] * ry) + plane[2] * rz;
    double frac = -plane[3] / denom;
    double sx = frac * rx, sy = frac * ry, sz = frac * rz;
    double h0 = ((T[0] * sx + T[1] * sy) + T[2] * sz) + T[3] * 1.0;
    double h1 = ((T[4] * sx + T[5] * sy) + T[6] * sz) + T[7] * 1.0;
    double h2 = ((T[8] * sx + T[9] * sy) + T[10] * sz) + T[11] * 1.0;
    double h3 = ((0.0 * sx + 0.0 * sy) + 0.0 * sz) + 1.0 * 1.0;
    return {h0 / h3, h1 / h3, h2 / h3};
}

struct Obj3D {
    double pos[3], scale[3];
};
// change_2d_corner_to_3d_object (object_3d_util.cpp:941-990): pose/scale part
__device__ __forceinline__ void corners_to_3d(const V2* c, const double* T, const double* invK, Obj3D& o) {
    // ground_plane_sensor = transToWolrd^T * (0,0,1,0) = third row of transToWolrd (box_proposal_detail.cpp:130-131, 376)
    double ground[4];
    for (int i = 0; i < 4; i++) ground[i] = ((T[i] * 0.0 + T[4 + i] * 0.0) + T[8 + i] * 1.0) + (i == 3 ? 1.0 : 0.0) * 0.0;
    V3 g0 = plane_hit_3d(T, invK, ground, c[4]), g1 = plane_hit_3d(T, invK, ground, c[5]);
    V3 g2 = plane_hit_3d(T, invK, ground, c[6]), g3 = plane_hit_3d(T, invK, ground, c[7]);
    double length_half = norm3(sub(g0, g3)) / 2;
    double width_half = norm3(sub(g0, g1)) / 2;
    // get_wall_plane_equation, :909-925
    V3 n = cross(sub(g0, g1), V3{0, 0, 1});
    double nn = norm3(n);
    n = {n.x / nn, n.y / nn, n.z / nn};
    double dist = ((-n.x) * g0.x + (-n.y) * g0.y) + (-n.z) * g0.z;
    double wall_w[4] = {n.x, n.y, n.z, dist};
    if (dist < 0)
        for (int i = 0; i < 4; i++) wall_w[i] = -wall_w[i];
    double wall_s[4];
    for (int i = 0; i < 4; i++) wall_s[i] = ((T[i] * wall_w[0] + T[4 + i] * wall_w[1]) + T[8 + i] * wall_w[2]) + (i == 3 ? 1.0 : 0.0) * wall_w[3];
    V3 topw = plane_hit_3d(T, invK, wall_s, c[1]);
    double height_half = topw.z / 2;
    double mean_x = (((g0.x + g1.x) + g2.x) + g3.x) / 4;
    double mean_y = (((g0.y + g1.y) + g2.y) + g3.y) / 4;
    o.pos[0] = mean_x; o.pos[1] = mean_y; o.pos[2] = height_half;
    o.scale[0] = length_half; o.scale[1] = width_half; o.scale[2] = height_half;
}

// ---- libstdc++ heap algorithms, restated (bits/stl_heap.h, bits/stl_algo.h __heap_select/__partial_sort) ----
// The reference ranks with std::partial_sort (matrix_utils.cpp:327-335), which is not stable: which of several
// equal keys ends up inside the kept prefix depends on these exact sift sequences, so they are reproduced literally.
// Generic over the element type so that the heap can carry (value, index) pairs.
template <class T, class Less>
__device__ __forceinline__ void heap_adjust(T* first, int holeIndex, int len, T value, Less less) {
    const int topIndex = holeIndex;
    int secondChild = holeIndex;
    while (secondChild < (len - 1) / 2) {
        secondChild = 2 * (secondChild + 1);
        if (less(first[secondChild], first[secondChild - 1])) secondChild--;
        first[holeIndex] = first[secondChild];
        holeIndex = secondChild;
    }
    if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
        secondChild = 2 * (secondChild + 1);
        first[holeIndex] = first[secondChild - 1];
        holeIndex = secondChild - 1;
    }
    // __push_heap
    int parent = (holeIndex - 1) / 2;
    while (holeIndex > topIndex && less(first[parent], value)) {
        first[holeIndex] = first[parent];
        holeIndex = parent;
        parent = (holeIndex - 1) / 2;
    }
    first[holeIndex] = value;
}
template <class T, class Less>
__device__ __forceinline__ void heap_make(T* first, int len, Less less) {
    if (len < 2) return;
    int parent = (len - 2) / 2;
    while (true) {
        T value = first[parent];
        heap_adjust(first, parent, len, value, less);
        if (parent == 0) return;
        parent--;
    }
}
// __sort_heap(first, first+k)
template <class T, class Less>
__device__ __forceinline__ void heap_sort(T* first, int k, Less less) {
    int last = k;
    while (last > 1) {
        --last;
        T value = first[last];
        first[last] = first[0];
        heap_adjust(first, 0, last, value, less);
    }
}
// __heap_select(first, first+k, first+n) for an index array (elements beyond k live in the same array)
template <class Less>
__device__ __forceinline__ void heap_select(int* first, int k, int n, Less less) {
    heap_make(first, k, less);
    for (int i = k; i < n; i++)
        if (less(first[i], first[0])) {  // __pop_heap(first, middle, i)
            int value = first[i];
            first[i] = first[0];
            heap_adjust(first, 0, k, value, less);
        }
}

// ---- warp-cooperative version of the same algorithms (bit-identical results) -------------------------------------
// The heap lives in three shared/global arrays: hv (value), hi (index) and big (for every node with two children, the
// child __adjust_heap would move into the hole: the right one unless right < left).  Facts used:
//  * __make_heap sifts parents in descending index order; nodes of one tree level own disjoint subtrees, so a level can be
//    sifted by many lanes at once with the same result.
//  * __adjust_heap first walks the hole from the root to a leaf along `big` (independent of the inserted value), then
//    __push_heap lifts the value: along that path the old contents are non-increasing, so the final slot is found with one
//    ballot, and the shift of the path elements is done by all lanes in parallel.
struct WarpHeap {
    double* hv;
    int* hi;
    int* big;
};

__device__ __forceinline__ int wh_bigger_child(const WarpHeap& H, int h) {
    const int r = 2 * h + 2;
    return (H.hv[r] < H.hv[r - 1]) ? r - 1 : r;
}

// literal single-lane __adjust_heap on the SoA arrays (used inside the level-parallel make_heap)
__device__ __forceinline__ void wh_adjust_serial(const WarpHeap& H, int holeIndex, int len, double value, int vidx) {
    const int topIndex = holeIndex;
    int secondChild = holeIndex;
    while (secondChild < (len - 1) / 2) {
        secondChild = 2 * (secondChild + 1);
        if (H.hv[secondChild] < H.hv[secondChild - 1]) secondChild--;
        H.hv[holeIndex] = H.hv[secondChild]; H.hi[holeIndex] = H.hi[secondChild];
        holeIndex = secondChild;
    }
    if ((len & 1) == 0 && secondChild == (len - 2) / 2) {
        secondChild = 2 * (secondChild + 1);
        H.hv[holeIndex] = H.hv[secondChild - 1]; H.hi[holeIndex] = H.hi[secondChild - 1];
        holeIndex = secondChild - 1;
    }
    int parent = (holeIndex - 1) / 2;
    while (holeIndex > topIndex && H.hv[parent] < value) {
        H.hv[holeIndex] = H.hv[parent]; H.hi[holeIndex] = H.hi[parent];
        holeIndex = parent;
        parent = (holeIndex - 1) / 2;
    }
    H.hv[holeIndex] = value; H.hi[holeIndex] = vidx;
}

// __make_heap(first, first+len), one warp
__device__ __forceinline__ void wh_make(const WarpHeap& H, int len, int lane) {
    if (len >= 2) {
        const int last_parent = (len - 2) / 2;
        int lvl = 0;
        while ((2 << lvl) - 2 < last_parent) lvl++;  // deepest level holding a parent: nodes [2^lvl - 1, 2^(lvl+1) - 2]
        for (; lvl >= 0; lvl--) {
            const int lo = (1 << lvl) - 1;
            int hi_node = (2 << lvl) - 2;
            if (hi_node > last_parent) hi_node = last_parent;
            for (int p = hi_node - lane; p >= lo; p -= 32) wh_adjust_serial(H, p, len, H.hv[p], H.hi[p]);
            __syncwarp();
        }
    }
    for (int h = lane; h < (len - 1) / 2; h += 32) H.big[h] = wh_bigger_child(H, h);
    __syncwarp();
}

// __adjust_heap(first, 0, len, value) executed by one warp; keeps `big` consistent for nodes with two children (< (len-1)/2)
__device__ __forceinline__ void wh_replace_root(const WarpHeap& H, int len, double value, int vidx, int lane) {
    const unsigned FULL = 0xffffffffu;
    // (a) hole path root -> leaf; lane l remembers the node of level l
    int h = 0, d = 0, mine = 0;
    const int half = (len - 1) / 2;
    while (h < half) {
        h = H.big[h];
        d++;
        if (lane == d) mine = h;
    }
    if ((len & 1) == 0 && h == (len - 2) / 2) {
        h = 2 * h + 1;
        d++;
        if (lane == d) mine = h;
    }
    // (b) old contents along the path; j = number of levels 1..d whose old value is not < value (a prefix, by the heap property)
    double ov = 0; int oi = 0;
    if (lane <= d) { ov = H.hv[mine]; oi = H.hi[mine]; }
    const unsigned stay = __ballot_sync(FULL, lane >= 1 && lane <= d && !(ov < value));
    const int j = __popc(stay);
    const double nv = __shfl_down_sync(FULL, ov, 1);
    const int ni = __shfl_down_sync(FULL, oi, 1);
    if (lane < j) { H.hv[mine] = nv; H.hi[mine] = ni; }
    else if (lane == j) { H.hv[mine] = value; H.hi[mine] = vidx; }
    __syncwarp();
    // (c) nodes on the path above the final slot had one child rewritten
    if (lane < j && mine < half) H.big[mine] = wh_bigger_child(H, mine);
    __syncwarp();
}

// ---- preferred-leaf variant used by the __heap_select phase (fixed heap length) ----------------------------------------
// H.big doubles as pl[]: for every internal node h (2h+1 < len) the leaf reached from h by always stepping to the bigger
// child (the single child for the last parent of an even-length heap).  The hole path of __adjust_heap(first, 0, len, ..)
// is then the ancestor chain of pl[0] -- computed arithmetically by all lanes at once instead of a 10-step pointer chase.
__device__ __forceinline__ int wh_depth(int node) { return 31 - __clz(node + 1); }
__device__ __forceinline__ int wh_ancestor(int leaf, int leaf_depth, int depth) { return ((leaf + 1) >> (leaf_depth - depth)) - 1; }
__device__ __forceinline__ int wh_pl_of(const WarpHeap& H, int node, int half) { return node < half ? H.big[node] : node; }

__device__ __forceinline__ void wh_build_pl(const WarpHeap& H, int len, int lane) {
    const int half = len >> 1;  // internal nodes are [0, half)
    if (half == 0) return;
    int lvl = wh_depth(half - 1);
    for (; lvl >= 0; lvl--) {
        const int lo = (1 << lvl) - 1;
        int hi_node = (2 << lvl) - 2;
        if (hi_node > half - 1) hi_node = half - 1;
        for (int h = lo + lane; h <= hi_node; h += 32) {
            const int bc = (2 * h + 2 < len) ? wh_bigger_child(H, h) : 2 * h + 1;
            H.big[h] = wh_pl_of(H, bc, half);
        }
        __syncwarp();
    }
}

// pl[] -> big[] (bigger child of every node with two children), in place
__device__ __forceinline__ void wh_pl_to_big(const WarpHeap& H, int len, int lane) {
    for (int h = lane; h < (len - 1) / 2; h += 32) {
        const int leaf = H.big[h];
        H.big[h] = wh_ancestor(leaf, wh_depth(leaf), wh_depth(h) + 1);
    }
    __syncwarp();
}

// __adjust_heap(first, 0, len, value) with the preferred-leaf table; len >= 2
__device__ __forceinline__ void wh_replace_root_pl(const WarpHeap& H, int len, double value, int vidx, int lane) {
    const unsigned FULL = 0xffffffffu;
    const int half = len >> 1;
    const int leaf = H.big[0];
    const int d = wh_depth(leaf);
    const int mine = (lane <= d) ? wh_ancestor(leaf, d, lane) : 0;
    // old contents along the path; j = number of levels 1..d whose old value is not < value (a prefix, by the heap property)
    double ov = 0; int oi = 0;
    if (lane <= d) { ov = H.hv[mine]; oi = H.hi[mine]; }
    const unsigned stay = __ballot_sync(FULL, lane >= 1 && lane <= d && !(ov < value));
    const int j = __popc(stay);
    const double nv = __shfl_down_sync(FULL, ov, 1);
    const int ni = __shfl_down_sync(FULL, oi, 1);
    if (lane < j) { H.hv[mine] = nv; H.hi[mine] = ni; }
    else if (lane == j) { H.hv[mine] = value; H.hi[mine] = vidx; }
    __syncwarp();
    if (j == 0) return;  // the value stays at the root: no child changed anywhere
    // path nodes above the final slot had one child rewritten: new bigger child, then the new preferred leaf.  A node whose
    // bigger child leaves the path ("terminal") takes that child's (unchanged) leaf; the node in the final slot keeps its own;
    // the others inherit from the next terminal level below them.
    bool terminal = false;
    int tval = 0;
    if (lane < j) {
        const int bc = (2 * mine + 2 < len) ? wh_bigger_child(H, mine) : 2 * mine + 1;
        if (bc != wh_ancestor(leaf, d, lane + 1)) { terminal = true; tval = wh_pl_of(H, bc, half); }
    } else if (lane == j) {
        terminal = true;
        tval = wh_pl_of(H, mine, half);
    }
    const unsigned tb = __ballot_sync(FULL, terminal);
    const int src = __ffs(tb & ~((1u << lane) - 1u)) - 1;  // first terminal level at or below this one (lane j always is)
    const int nl = __shfl_sync(FULL, tval, src < 0 ? 0 : src);
    if (lane < j) H.big[mine] = nl;
    __syncwarp();
}

// std::partial_sort(iota, iota + k, iota + N) by vd with the libstdc++ algorithms above; on return hi[0..k) holds the heap
// (sorted == false: hi[0] is the excluded k-th element) or the sorted prefix (sorted == true).  One warp.
__device__ __forceinline__ void wh_partial_sort(const WarpHeap& H, const double* vd, int k, int N, bool sorted, int lane) {
    const unsigned FULL = 0xffffffffu;
    for (int i = lane; i < k; i += 32) { H.hv[i] = vd[i]; H.hi[i] = i; }
    __syncwarp();
    wh_make(H, k, lane);
    wh_build_pl(H, k, lane);  // H.big holds the preferred-leaf table during __heap_select
    // __heap_select: 32 candidates at a time; everything before the first one that beats the root is a no-op
    int i = k;
    while (i < N) {
        const int c = i + lane;
        const double root = H.hv[0];
        const double cv = (c < N) ? vd[c] : 0.0;
        const unsigned hit = __ballot_sync(FULL, c < N && cv < root);
        if (!hit) { i += 32; continue; }
        const int src = __ffs(hit) - 1;
        const double v = __shfl_sync(FULL, cv, src);
        wh_replace_root_pl(H, k, v, i + src, lane);
        i += src + 1;
    }
    if (sorted) {
        wh_pl_to_big(H, k, lane);
        // __sort_heap: repeatedly move the root behind the shrinking heap and re-insert the former last leaf
        for (int last = k - 1; last >= 1; last--) {
            const double v = H.hv[last]; const int vi = H.hi[last];
            const double rv = H.hv[0]; const int ri = H.hi[0];
            __syncwarp();
            if (lane == 0) { H.hv[last] = rv; H.hi[last] = ri; }
            __syncwarp();
            if (last >= 2) wh_replace_root(H, last, v, vi, lane);
            else if (lane == 0) { H.hv[0] = v; H.hi[0] = vi; }
            __syncwarp();
        }
    }
}

}  // namespace csb
