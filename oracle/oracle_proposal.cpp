// TEST INFRASTRUCTURE ONLY -- CPU oracle, proposal half of the CubeSLAM hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may build or call this file; the shipped library never links it.
//
// A plain C++ restatement of detect_3d_cuboid::detect_cuboid() and the primitives it
// calls, with plotting / printing removed.  Each function cites the reference lines it
// follows (paths relative to /root/reference).  Canny + distance transform are third-party
// OpenCV in the reference (box_proposal_detail.cpp:320-327) and are INPUTS here: the caller
// supplies one float32 distance map per (box, height-sample) task, in orc_plan() order.
//
// PARITY STATUS: PINNED against outputs of the reference itself.  The reference has no tests and does not compile here (needs Eigen,
// OpenCV, ROS), but it commits what its object_slam node wrote in online mode for the bundled TUM sequence
// (object_slam/data/output_obj_poses.txt, output_cam_poses.txt), and tests/test_reference_replay.py re-runs that mode with this file
// (after oracle_edlines.cpp + cv2 Canny / distance transform, before oracle_ba.cpp): the landmark pose after every one of the 58 frames
// comes out to the files' printed digits, frame 0 being a single detect_cuboid() result, frames 1.. with camera roll / pitch sampling on.
// Also pinned:
//   * ray_plane_interact worked example, object_3d_util.cpp:884-905 (tests/test_oracle_golden.py)
//   * enumeration counts on the bundled demo (320/111, 6400/1799)
// Defined-by-oracle behaviour where the reference has UB: dist_map.at<float>(r,c) is
// evaluated as the linear index r*cols+c into the continuous buffer, clamped to the last
// element (reference reads out of bounds when a corner lies on the ROI's right/bottom bound).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <numeric>
#include <thread>
#include <vector>

#include "oracle_math.h"

using namespace orc;

extern "C" {
struct orc_params {
    int consider_config_1, consider_config_2;
    int whether_sample_cam_roll_pitch, whether_sample_bbox_height;
    int max_cuboid_num;
    int leak_cam_state;  // 1: literal reference (cam_pose member leaks across boxes); 0: boxes independent
    double nominal_skew_ratio, max_cut_skew;
    int libm_atan2;      // 1: std::atan2 like the reference; 0 (default): the specified det_atan2 (oracle_math.h) the GPU path also uses
    int reserved;
};
struct orc_cuboid {  // mirror of class cuboid, detect_3d_cuboid.h:20-41
    double pos[3], scale[3], rotY;
    double box_config_type[2];
    double box_corners_3d_world[24];  // 3x8 row-major
    double rect_detect_2d[4];
    double edge_distance_error, edge_angle_error, normalized_error, skew_ratio;
    double down_expand_height, camera_roll_delta, camera_pitch_delta;
    int box_corners_2d[16];  // 2x8 row-major
    int task_id, raw_cube_ind;  // provenance: task and compacted valid index
};
struct orc_task {
    int box_id, hs_id, down_expand;
    int left, top, width, height;  // dist-map ROI (cv::Rect, box_proposal_detail.cpp:320)
    int n_top;
    long long map_offset;  // float offset of this task's map in the packed buffer
};
}

namespace {

// atan2 used on the scored path (see oracle_math.h "deterministic atan2"); set once per call from orc_params::libm_atan2
bool g_libm_atan2 = false;
inline double sp_atan2(double y, double x) { return g_libm_atan2 ? std::atan2(y, x) : det_atan2(y, x); }

// detect_3d_cuboid.h:59-71
struct CamPose {
    M4 transToWolrd;
    M3 Kalib, rotationToWorld, invR, invK, KinvR;
    double euler_angle[3];
    double projectionMatrix[12];
    double camera_yaw;
};

// box_proposal_detail.cpp:38-42
void set_calibration(CamPose& c, const M3& K) { c.Kalib = K; c.invK = inverse3(K); }

// box_proposal_detail.cpp:45-56
void set_cam_pose(CamPose& c, const M4& T) {
    c.transToWolrd = T;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) c.rotationToWorld(i, j) = T(i, j);
    quat_to_euler_zyx(quat_from_rot(c.rotationToWorld), c.euler_angle[0], c.euler_angle[1], c.euler_angle[2]);
    c.invR = inverse3(c.rotationToWorld);
    M4 Ti = inverse4(T);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 4; j++)
            c.projectionMatrix[i * 4 + j] = (c.Kalib(i, 0) * Ti(0, j) + c.Kalib(i, 1) * Ti(1, j)) + c.Kalib(i, 2) * Ti(2, j);
    c.KinvR = mul(c.Kalib, c.invR);
    c.camera_yaw = c.euler_angle[2];
}

// matrix_utils.cpp:368-380 (accumulating linspace; rounding decides the sample count)
template <class T>
void linespace(T starting, T ending, T step, std::vector<T>& res) {
    while (starting <= ending) {
        res.push_back(starting);
        starting += step;
        if (res.size() > 1000) break;
    }
}

// object_3d_util.cpp:239-242
bool check_inside_box(V2 pt, V2 lt, V2 rb) { return lt.x <= pt.x && pt.x <= rb.x && lt.y <= pt.y && pt.y <= rb.y; }

// object_3d_util.cpp:309-353
V2 seg_hit_boundary(V2 pt_start, V2 pt_end, double bx0, double by0, double bx1, double by1) {
    V2 direc = pt_end - pt_start;
    V2 hit{-1, -1};
    if (by0 == by1) {  // horizontal edge
        double lambd = (by0 - pt_start.y) / direc.y;
        if (lambd >= 0) {
            V2 tmp = pt_start + lambd * direc;
            if ((bx0 <= tmp.x) && (tmp.x <= bx1)) {
                hit = tmp;
                hit.y = by0;
            }
        }
    }
    if (bx0 == bx1) {  // vertical edge
        double lambd = (bx0 - pt_start.x) / direc.x;
        if (lambd >= 0) {
            V2 tmp = pt_start + lambd * direc;
            if ((by0 <= tmp.y) && (tmp.y <= by1)) {
                hit = tmp;
                hit.x = bx0;
            }
        }
    }
    return hit;
}

// object_3d_util.cpp:357-382 (always called with infinite_line = true on this path)
V2 lineSegmentIntersect(V2 p1s, V2 p1e, V2 p2s, V2 p2e) {
    double X2_X1 = p1e.x - p1s.x, Y2_Y1 = p1e.y - p1s.y;
    double X4_X3 = p2e.x - p2s.x, Y4_Y3 = p2e.y - p2s.y;
    double X1_X3 = p1s.x - p2s.x, Y1_Y3 = p1s.y - p2s.y;
    double u_a = (X4_X3 * Y1_Y3 - Y4_Y3 * X1_X3) / (Y4_Y3 * X2_X1 - X4_X3 * Y2_Y1);
    double INT_X = p1s.x + X2_X1 * u_a;
    double INT_Y = p1s.y + Y2_Y1 * u_a;
    double INT_B = 1;
    return {INT_X * INT_B, INT_Y * INT_B};
}

struct Line { double x1, y1, x2, y2; };

// object_3d_util.cpp:431-543 (+ matrix_utils.cpp:183-187 fast_RemoveRow, object_3d_util.cpp:269-274)
void merge_break_lines(const std::vector<Line>& all_lines, std::vector<Line>& out, double pre_merge_dist_thre,
                       double pre_merge_angle_thre_degree, double edge_length_threshold) {
    bool can_force_merge = true;
    std::vector<Line> L = all_lines;
    int total = (int)L.size();
    int counter = 0;
    double pre_merge_angle_thre = pre_merge_angle_thre_degree / 180.0 * M_PI;
    std::vector<double> ang(L.size());
    while (can_force_merge && (counter < 500)) {
        counter++;
        can_force_merge = false;
        for (int i = 0; i < total; i++) ang[i] = sp_atan2(L[i].y2 - L[i].y1, L[i].x2 - L[i].x1);
        for (int seg1 = 0; seg1 < total - 1; seg1++) {
            for (int seg2 = seg1 + 1; seg2 < total; seg2++) {
                double diff = std::abs(ang[seg1] - ang[seg2]);
                double angle_diff = std::min(diff, M_PI - diff);
                if (angle_diff < pre_merge_angle_thre) {
                    double d12 = norm(V2{L[seg1].x2 - L[seg2].x1, L[seg1].y2 - L[seg2].y1});
                    double d21 = norm(V2{L[seg2].x2 - L[seg1].x1, L[seg2].y2 - L[seg1].y1});
                    if ((d12 < pre_merge_dist_thre) || (d21 < pre_merge_dist_thre)) {
                        V2 ms, me;
                        if (L[seg1].x1 < L[seg2].x1) ms = {L[seg1].x1, L[seg1].y1};
                        else ms = {L[seg2].x1, L[seg2].y1};
                        if (L[seg1].x2 > L[seg2].x2) me = {L[seg1].x2, L[seg1].y2};
                        else me = {L[seg2].x2, L[seg2].y2};
                        double merged_angle = sp_atan2(me.y - ms.y, me.x - ms.x);
                        double temp = std::abs(ang[seg1] - merged_angle);
                        double merge_angle_diff = std::min(temp, M_PI - temp);
                        if (merge_angle_diff < pre_merge_angle_thre) {
                            L[seg1] = {ms.x, ms.y, me.x, me.y};
                            L[seg2] = L[total - 1];  // fast_RemoveRow
                            total--;
                            can_force_merge = true;
                            break;
                        }
                    }
                }
            }
            if (can_force_merge) break;
        }
    }
    out.clear();
    if (edge_length_threshold > 0) {
        for (int i = 0; i < total; i++) {
            double len = norm(V2{L[i].x2 - L[i].x1, L[i].y2 - L[i].y1});
            if (len > edge_length_threshold) out.push_back(L[i]);
        }
    } else {
        out.assign(L.begin(), L.begin() + total);
    }
}

// object_3d_util.cpp:548-619 (+ smooth_jump_angles :278-302).  out[vp*2+{0,1}], NaN if unsupported.
void VP_support_edge_infos(const V2 VPs[3], const std::vector<V2>& mid, const std::vector<double>& edge_angles,
                           double thre12_deg, double thre3_deg, double out[6]) {
    for (int i = 0; i < 6; i++) out[i] = std::nan("");
    int n = (int)edge_angles.size();
    if (n == 0) return;
    std::vector<int> inlier_id(n);
    std::vector<double> inlier_raw(n);
    for (int vp_id = 0; vp_id < 3; vp_id++) {
        double vp_angle_thre = (vp_id != 2 ? thre12_deg : thre3_deg) / 180.0 * M_PI;
        int cnt = 0;
        for (int e = 0; e < n; e++) {
            double raw = sp_atan2(mid[e].y - VPs[vp_id].y, mid[e].x - VPs[vp_id].x);
            double nrm = normalize_to_pi(raw);
            double d = std::abs(edge_angles[e] - nrm);
            d = std::min(d, M_PI - d);
            if (d < vp_angle_thre) {
                inlier_raw[cnt] = raw;
                inlier_id[cnt] = e;
                cnt++;
            }
        }
        if (cnt > 0) {
            // smooth_jump_angles
            double base = inlier_raw[0];
            int imax = 0, imin = 0;
            double vmax = 0, vmin = 0;
            for (int i = 0; i < cnt; i++) {
                double v = inlier_raw[i];
                if ((inlier_raw[i] - base) < -M_PI) v = inlier_raw[i] + 2 * M_PI;
                else if ((inlier_raw[i] - base) > M_PI) v = inlier_raw[i] - 2 * M_PI;
                // Eigen maxCoeff/minCoeff(&idx): first extremum wins (strict comparison)
                if (i == 0) { vmax = vmin = v; }
                else {
                    if (v > vmax) { vmax = v; imax = i; }
                    if (v < vmin) { vmin = v; imin = i; }
                }
            }
            int low = imax, top = imin;
            if (vp_id > 0) std::swap(low, top);
            out[vp_id * 2 + 0] = edge_angles[inlier_id[low]];
            out[vp_id * 2 + 1] = edge_angles[inlier_id[top]];
        }
    }
}

// corners: c[0..7].{x,y}  (0-based ids of the reference's 1..8)
// object_3d_util.cpp:622-667
double box_edge_sum_dists(const float* dist_map, int rows, int cols, const V2 c[8], const int (*edges)[2], int n_edges,
                          bool reweight) {
    float sum_dist = 0;
    const long long last = (long long)rows * cols - 1;
    for (int e = 0; e < n_edges; e++) {
        V2 c1 = c[edges[e][0]], c2 = c[edges[e][1]];
        for (double s = 0; s < 11; s++) {
            double sx = s / 10.0 * c1.x + (1 - s / 10.0) * c2.x;
            double sy = s / 10.0 * c1.y + (1 - s / 10.0) * c2.y;
            long long li = (long long)int(sy) * cols + int(sx);
            if (li < 0) li = 0;
            if (li > last) li = last;  // oracle-defined clamp (reference: UB)
            float d1 = dist_map[li];
            if (reweight) {
                if ((4 <= e) && (e <= 5)) d1 = d1 * 3.0 / 2.0;
                if (6 == e) d1 = d1 * 2.0;
            }
            sum_dist = sum_dist + d1;
        }
    }
    return double(sum_dist);
}

// object_3d_util.cpp:670-723
double box_edge_alignment_angle_error(const double vp_bound[6], const int (*vpe)[4], const V2 c[8]) {
    double total = 0;
    double not_found_penalty = 30.0 / 180.0 * M_PI * 2;
    for (int vp = 0; vp < 3; vp++) {
        double valid[2];
        int nv = 0;
        for (int i = 0; i < 2; i++)
            if (!std::isnan(vp_bound[vp * 2 + i])) valid[nv++] = vp_bound[vp * 2 + i];
        if (nv > 0) {
            for (int ee = 0; ee < 2; ee++) {
                V2 a = c[vpe[vp][2 * ee]], b = c[vpe[vp][2 * ee + 1]];
                double box_edge_angle = normalize_to_pi(sp_atan2(b.y - a.y, b.x - a.x));
                double angle_diff_temp = 100;
                for (int i = 0; i < nv; i++) {
                    double temp = std::abs(box_edge_angle - valid[i]);
                    temp = std::min(temp, M_PI - temp);
                    if (temp < angle_diff_temp) angle_diff_temp = temp;
                }
                total = total + angle_diff_temp;
            }
        } else
            total = total + not_found_penalty;
    }
    return total;
}

// matrix_utils.cpp:327-335 -- must be the real (unstable) std::partial_sort
void sort_indexes(const std::vector<double>& vec, std::vector<int>& idx, int top_k) {
    std::partial_sort(idx.begin(), idx.begin() + top_k, idx.end(), [&vec](int i1, int i2) { return vec[i1] < vec[i2]; });
}

// object_3d_util.cpp:726-837
void fuse_normalize_scores_v2(const std::vector<double>& dist_error, const std::vector<double>& angle_error,
                              std::vector<double>& combined, std::vector<int>& final_keep, double weight_vp_angle,
                              bool whether_normalize) {
    int raw_data_size = (int)dist_error.size();
    final_keep.clear();
    if (raw_data_size > 4) {
        int breaking_num = (int)round(float(raw_data_size) / 3.0 * 2.0);
        std::vector<int> dist_sorted(raw_data_size);
        std::iota(dist_sorted.begin(), dist_sorted.end(), 0);
        std::vector<int> angle_sorted = dist_sorted;
        sort_indexes(dist_error, dist_sorted, breaking_num);
        sort_indexes(angle_error, angle_sorted, breaking_num);
        std::vector<int> dist_keep(dist_sorted.begin(), dist_sorted.begin() + breaking_num - 1);
        if (angle_error[angle_sorted[breaking_num - 1]] > angle_error[angle_sorted[breaking_num - 2]]) {
            std::vector<int> angle_keep(angle_sorted.begin(), angle_sorted.begin() + breaking_num - 1);
            std::sort(dist_keep.begin(), dist_keep.end());
            std::sort(angle_keep.begin(), angle_keep.end());
            std::set_intersection(dist_keep.begin(), dist_keep.end(), angle_keep.begin(), angle_keep.end(),
                                  std::back_inserter(final_keep));
        } else {
            final_keep = dist_keep;  // NOTE: left in partial_sort order (not index order), as the reference does
        }
    } else {
        final_keep.resize(raw_data_size);
        std::iota(final_keep.begin(), final_keep.end(), 0);
    }
    int new_data_size = (int)final_keep.size();
    double min_d = 1e6, max_d = -1, min_a = 1e6, max_a = -1;
    std::vector<double> dk(new_data_size), ak(new_data_size);
    for (int i = 0; i < new_data_size; i++) {
        double td = dist_error[final_keep[i]], ta = angle_error[final_keep[i]];
        min_d = std::min(min_d, td); max_d = std::max(max_d, td);
        min_a = std::min(min_a, ta); max_a = std::max(max_a, ta);
        dk[i] = td; ak[i] = ta;
    }
    combined.resize(new_data_size);
    if (whether_normalize && (new_data_size > 1)) {
        for (int i = 0; i < new_data_size; i++) combined[i] = (dk[i] - min_d) / (max_d - min_d);
        if ((max_a - min_a) > 0) {
            for (int i = 0; i < new_data_size; i++) ak[i] = (ak[i] - min_a) / (max_a - min_a);
            for (int i = 0; i < new_data_size; i++) combined[i] = (combined[i] + weight_vp_angle * ak[i]) / (1 + weight_vp_angle);
        } else
            for (int i = 0; i < new_data_size; i++) combined[i] = (combined[i] + weight_vp_angle * ak[i]) / (1 + weight_vp_angle);
    } else
        for (int i = 0; i < new_data_size; i++) combined[i] = (dk[i] + weight_vp_angle * ak[i]) / (1 + weight_vp_angle);
}

// object_3d_util.cpp:841-847, 853-906
void plane_hits_3d(const M4& T, const M3& invK, const double plane[4], const V2* pixels, int n, V3* out_world) {
    for (int i = 0; i < n; i++) {
        V3 ray = mul(invK, V3{pixels[i].x, pixels[i].y, 1.0});
        double denom = (plane[0] * ray.x + plane[1] * ray.y) + plane[2] * ray.z;
        double frac = -plane[3] / denom;
        V3 ps{frac * ray.x, frac * ray.y, frac * ray.z};
        double h[4];
        for (int r = 0; r < 4; r++) h[r] = ((T(r, 0) * ps.x + T(r, 1) * ps.y) + T(r, 2) * ps.z) + T(r, 3) * 1.0;
        out_world[i] = {h[0] / h[3], h[1] / h[3], h[2] / h[3]};
    }
}

// object_3d_util.cpp:909-925
void get_wall_plane_equation(V3 p1, V3 p2, double plane[4]) {
    V3 n = cross(p1 - p2, V3{0, 0, 1});
    double nn = norm(n);
    n = {n.x / nn, n.y / nn, n.z / nn};
    double dist = ((-n.x) * p1.x + (-n.y) * p1.y) + (-n.z) * p1.z;
    plane[0] = n.x; plane[1] = n.y; plane[2] = n.z; plane[3] = dist;
    if (dist < 0)
        for (int i = 0; i < 4; i++) plane[i] = -plane[i];
}

// object_3d_util.cpp:15-73 (similarityTransformation + compute3D_BoxCorner)
void compute3D_BoxCorner(const orc_cuboid& o, double out[24]) {
    static const double body[3][8] = {{1, 1, -1, -1, 1, 1, -1, -1}, {1, -1, -1, 1, 1, -1, -1, 1}, {-1, -1, -1, -1, 1, 1, 1, 1}};
    double c = std::cos(o.rotY), s = std::sin(o.rotY);
    M3 rot; rot(0, 0) = c; rot(0, 1) = -s; rot(0, 2) = 0; rot(1, 0) = s; rot(1, 1) = c; rot(1, 2) = 0; rot(2, 0) = 0; rot(2, 1) = 0; rot(2, 2) = 1;
    M3 sc; std::memset(sc.m, 0, sizeof sc.m); sc(0, 0) = o.scale[0]; sc(1, 1) = o.scale[1]; sc(2, 2) = o.scale[2];
    M3 rs = mul(rot, sc);
    M4 res; std::memset(res.m, 0, sizeof res.m); res(3, 3) = 1;
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) res(i, j) = rs(i, j); res(i, 3) = o.pos[i]; }
    for (int k = 0; k < 8; k++) {
        double h[4];
        for (int r = 0; r < 4; r++) h[r] = ((res(r, 0) * body[0][k] + res(r, 1) * body[1][k]) + res(r, 2) * body[2][k]) + res(r, 3) * 1.0;
        for (int r = 0; r < 3; r++) out[r * 8 + k] = h[r] / h[3];
    }
}

// object_3d_util.cpp:941-1011
void change_2d_corner_to_3d_object(const V2 c[8], double cfg, double vp_1_position, double yaw_esti, const double ground_plane_sensor[4],
                                   const M4& T, const M3& invK, orc_cuboid& o) {
    V3 gnd[4];
    plane_hits_3d(T, invK, ground_plane_sensor, c + 4, 4, gnd);
    double length_half = norm(gnd[0] - gnd[3]) / 2;
    double width_half = norm(gnd[0] - gnd[1]) / 2;
    double wall_w[4], wall_s[4];
    get_wall_plane_equation(gnd[0], gnd[1], wall_w);
    for (int i = 0; i < 4; i++) wall_s[i] = ((T(0, i) * wall_w[0] + T(1, i) * wall_w[1]) + T(2, i) * wall_w[2]) + T(3, i) * wall_w[3];
    V3 topw;
    plane_hits_3d(T, invK, wall_s, c + 1, 1, &topw);
    double height_half = topw.z / 2;
    double mean_x = (((gnd[0].x + gnd[1].x) + gnd[2].x) + gnd[3].x) / 4;
    double mean_y = (((gnd[0].y + gnd[1].y) + gnd[2].y) + gnd[3].y) / 4;
    o.pos[0] = mean_x; o.pos[1] = mean_y; o.pos[2] = height_half;
    o.rotY = yaw_esti;
    o.scale[0] = length_half; o.scale[1] = width_half; o.scale[2] = height_half;
    o.box_config_type[0] = cfg; o.box_config_type[1] = vp_1_position;
    static const int ids_left[8] = {6, 5, 8, 7, 2, 3, 4, 1}, ids_right[8] = {5, 6, 7, 8, 3, 2, 1, 4};
    const int* ids = (vp_1_position == 1) ? ids_left : ids_right;
    for (int i = 0; i < 8; i++) {
        o.box_corners_2d[i] = (int)c[ids[i] - 1].x;
        o.box_corners_2d[8 + i] = (int)c[ids[i] - 1].y;
    }
    compute3D_BoxCorner(o, o.box_corners_3d_world);
}

struct TaskResult {
    orc_task task;
    int n_enum = 0;
    std::vector<Line> merged;
    std::vector<double> rows;     // N x 9  (all_configs_error_one_objH)
    std::vector<double> corners;  // N x 16 (x0..x7,y0..y7)  (all_box_corners_2d_one_objH)
    std::vector<int> hyp_id;      // N enumeration ids
    std::vector<int> keep;        // good_proposal_ids
    std::vector<double> norm_score;
};
struct BoxResult {
    std::vector<orc_cuboid> raw;       // raw_obj_proposals
    std::vector<double> combined;      // all_combined_score
    std::vector<int> sorted;           // sort_idx_small[0..k)
};
struct FrameResult {
    std::vector<TaskResult> tasks;
    std::vector<BoxResult> boxes;
    long long n_scored = 0, n_enum = 0;
};

struct BoxGeom {
    int left_x_raw, top_y_raw, obj_width_raw, obj_height_raw, right_x_raw, down_y_raw;
};

// box_proposal_detail.cpp:143-172
BoxGeom box_geom(const double* b) {
    BoxGeom g;
    g.left_x_raw = (int)b[0]; g.top_y_raw = (int)b[1]; g.obj_width_raw = (int)b[2]; g.obj_height_raw = (int)b[3];
    g.right_x_raw = (int)(g.left_x_raw + b[2]);
    g.down_y_raw = g.top_y_raw + g.obj_height_raw;
    return g;
}
std::vector<int> down_expand_samples(const BoxGeom& g, int img_height, bool sample_height) {
    std::vector<int> v; v.push_back(0);
    if (sample_height) {
        int r = std::max(std::min(20, g.obj_height_raw - 90), 20);
        r = std::min(r, img_height - g.top_y_raw - g.obj_height_raw - 1);
        if (r > 10) v.push_back((int)round(r / 2));
        v.push_back(r);
    }
    return v;
}
struct TaskGeom {
    int down_expand, obj_height_expan, down_y_expan, top_sample_resolution;
    int left_e, right_e, top_e, down_e, height_e, width_e;
    double diag;
};
// box_proposal_detail.cpp:202-248
TaskGeom task_geom(const BoxGeom& g, int down_expand, int img_width, int img_height) {
    TaskGeom t;
    t.down_expand = down_expand;
    t.obj_height_expan = g.obj_height_raw + down_expand;
    t.down_y_expan = g.top_y_raw + t.obj_height_expan;
    t.diag = std::sqrt((double)(g.obj_width_raw * g.obj_width_raw + t.obj_height_expan * t.obj_height_expan));
    t.top_sample_resolution = (int)round(std::min(20, g.obj_width_raw / 10));
    int w = std::min(std::max(std::min(20, g.obj_width_raw - 100), 10), std::max(std::min(20, t.obj_height_expan - 100), 10));
    t.left_e = std::max(0, g.left_x_raw - w);
    t.right_e = std::min(img_width - 1, g.right_x_raw + w);
    t.top_e = std::max(0, g.top_y_raw - w);
    t.down_e = std::min(img_height - 1, t.down_y_expan + w);
    t.height_e = t.down_e - t.top_e;
    t.width_e = t.right_e - t.left_e;
    return t;
}

int plan_tasks(const double* boxes, int n_boxes, int img_w, int img_h, bool sample_height, std::vector<orc_task>& out) {
    long long off = 0;
    for (int b = 0; b < n_boxes; b++) {
        BoxGeom g = box_geom(boxes + 5 * b);
        std::vector<int> de = down_expand_samples(g, img_h, sample_height);
        for (int hs = 0; hs < (int)de.size(); hs++) {
            TaskGeom t = task_geom(g, de[hs], img_w, img_h);
            if (t.top_sample_resolution < 1) break;  // box_proposal_detail.cpp:215-216
            std::vector<int> tops;
            linespace<int>(g.left_x_raw + 5, g.right_x_raw - 5, t.top_sample_resolution, tops);
            orc_task k;
            k.box_id = b; k.hs_id = hs; k.down_expand = de[hs];
            k.left = t.left_e; k.top = t.top_e; k.width = t.width_e; k.height = t.height_e;
            k.n_top = (int)tops.size();
            k.map_offset = off;
            off += (long long)std::max(0, t.width_e) * std::max(0, t.height_e);
            out.push_back(k);
        }
    }
    return (int)out.size();
}

// box_proposal_detail.cpp:65-861
void detect_cuboid(const double* Kin, const double* Tin, int img_width, int img_height, const double* boxes, int num_2d_objs,
                   const double* lines_in, int n_lines, const float* dist_maps, const orc_params& P, FrameResult& R) {
    CamPose cam_pose, cam_pose_raw;
    M3 K; std::memcpy(K.m, Kin, sizeof K.m);
    M4 transToWolrd; std::memcpy(transToWolrd.m, Tin, sizeof transToWolrd.m);
    set_calibration(cam_pose, K);
    set_cam_pose(cam_pose, transToWolrd);
    cam_pose_raw = cam_pose;
    R.boxes.resize(num_2d_objs);

    bool all_configs[2] = {P.consider_config_1 != 0, P.consider_config_2 != 0};
    const double vp12_edge_angle_thre = 15, vp3_edge_angle_thre = 10, shorted_edge_thre = 20;
    const bool reweight_edge_distance = true, whether_normalize_two_errors = true;
    const double weight_vp_angle = 0.8, weight_skew_error = 1.5;

    // align_left_right_edges, object_3d_util.cpp:246-258
    std::vector<Line> all_lines_raw(n_lines);
    for (int i = 0; i < n_lines; i++) {
        Line l{lines_in[4 * i], lines_in[4 * i + 1], lines_in[4 * i + 2], lines_in[4 * i + 3]};
        if (l.x2 < l.x1) l = {l.x2, l.y2, l.x1, l.y1};
        all_lines_raw[i] = l;
    }
    // ground_plane_sensor = T^T * (0,0,1,0)  (box_proposal_detail.cpp:130-131)
    auto ground_from = [](const M4& T, double g[4]) {
        for (int i = 0; i < 4; i++) g[i] = ((T(0, i) * 0.0 + T(1, i) * 0.0) + T(2, i) * 1.0) + T(3, i) * 0.0;
    };
    double ground_plane_sensor[4];
    ground_from(cam_pose.transToWolrd, ground_plane_sensor);

    long long map_off = 0;
    for (int object_id = 0; object_id < num_2d_objs; object_id++) {
        BoxGeom g = box_geom(boxes + 5 * object_id);
        const int left_x_raw = g.left_x_raw, top_y_raw = g.top_y_raw, obj_width_raw = g.obj_width_raw, obj_height_raw = g.obj_height_raw,
                  right_x_raw = g.right_x_raw;
        std::vector<int> down_expand_sample_all = down_expand_samples(g, img_height, P.whether_sample_bbox_height != 0);

        const CamPose& yaw_src = P.leak_cam_state ? cam_pose : cam_pose_raw;
        double yaw_init = yaw_src.camera_yaw - 90.0 / 180.0 * M_PI;
        std::vector<double> obj_yaw_samples;
        linespace<double>(yaw_init - 45.0 / 180.0 * M_PI, yaw_init + 45.0 / 180.0 * M_PI, 6.0 / 180.0 * M_PI, obj_yaw_samples);

        BoxResult& BR = R.boxes[object_id];
        for (int hs = 0; hs < (int)down_expand_sample_all.size(); hs++) {
            int down_expand_sample = down_expand_sample_all[hs];
            TaskGeom tg = task_geom(g, down_expand_sample, img_width, img_height);
            const int down_y_expan = tg.down_y_expan;
            const double obj_diaglength_expan = tg.diag;
            if (tg.top_sample_resolution < 1) break;
            std::vector<int> top_x_samples;
            linespace<int>(left_x_raw + 5, right_x_raw - 5, tg.top_sample_resolution, top_x_samples);

            const int left_x_expan_distmap = tg.left_e, right_x_expan_distmap = tg.right_e, top_y_expan_distmap = tg.top_e,
                      down_y_expan_distmap = tg.down_e, height_expan_distmap = tg.height_e, width_expan_distmap = tg.width_e;
            V2 expan_lt{(double)left_x_expan_distmap, (double)top_y_expan_distmap}, expan_rb{(double)right_x_expan_distmap, (double)down_y_expan_distmap};

            R.tasks.emplace_back();
            TaskResult& TR = R.tasks.back();
            int task_id = (int)R.tasks.size() - 1;
            TR.task.box_id = object_id; TR.task.hs_id = hs; TR.task.down_expand = down_expand_sample;
            TR.task.left = left_x_expan_distmap; TR.task.top = top_y_expan_distmap; TR.task.width = width_expan_distmap; TR.task.height = height_expan_distmap;
            TR.task.n_top = (int)top_x_samples.size();
            TR.task.map_offset = map_off;
            const float* dist_map = dist_maps + map_off;
            map_off += (long long)std::max(0, width_expan_distmap) * std::max(0, height_expan_distmap);

            // lines inside the expanded ROI (box_proposal_detail.cpp:271-283)
            std::vector<Line> inside;
            for (int e = 0; e < n_lines; e++)
                if (check_inside_box(V2{all_lines_raw[e].x1, all_lines_raw[e].y1}, expan_lt, expan_rb))
                    if (check_inside_box(V2{all_lines_raw[e].x2, all_lines_raw[e].y2}, expan_lt, expan_rb)) inside.push_back(all_lines_raw[e]);
            merge_break_lines(inside, TR.merged, 20, 5, 30);
            int nl = (int)TR.merged.size();
            std::vector<double> lines_inobj_angles(nl);
            std::vector<V2> edge_mid_pts(nl);
            for (int i = 0; i < nl; i++) {
                const Line& l = TR.merged[i];
                lines_inobj_angles[i] = sp_atan2(l.y2 - l.y1, l.x2 - l.x1);
                edge_mid_pts[i] = {(l.x1 + l.x2) / 2, (l.y1 + l.y2) / 2};
            }

            std::vector<double> cam_roll_samples, cam_pitch_samples;
            if (P.whether_sample_cam_roll_pitch) {
                linespace<double>(cam_pose_raw.euler_angle[0] - 6.0 / 180.0 * M_PI, cam_pose_raw.euler_angle[0] + 6.0 / 180.0 * M_PI, 3.0 / 180.0 * M_PI, cam_roll_samples);
                linespace<double>(cam_pose_raw.euler_angle[1] - 6.0 / 180.0 * M_PI, cam_pose_raw.euler_angle[1] + 6.0 / 180.0 * M_PI, 3.0 / 180.0 * M_PI, cam_pitch_samples);
            } else {
                cam_roll_samples.push_back(cam_pose_raw.euler_angle[0]);
                cam_pitch_samples.push_back(cam_pose_raw.euler_angle[1]);
            }

            int n_top = (int)top_x_samples.size();
            int valid_n = 0;
            for (int cam_roll_id = 0; cam_roll_id < (int)cam_roll_samples.size(); cam_roll_id++)
                for (int cam_pitch_id = 0; cam_pitch_id < (int)cam_pitch_samples.size(); cam_pitch_id++)
                    for (int obj_yaw_id = 0; obj_yaw_id < (int)obj_yaw_samples.size(); obj_yaw_id++) {
                        if (P.whether_sample_cam_roll_pitch) {
                            M4 Tn = transToWolrd;
                            M3 Rn = euler_zyx_to_rot(cam_roll_samples[cam_roll_id], cam_pitch_samples[cam_pitch_id], cam_pose_raw.euler_angle[2]);
                            for (int i = 0; i < 3; i++)
                                for (int j = 0; j < 3; j++) Tn(i, j) = Rn(i, j);
                            set_cam_pose(cam_pose, Tn);
                            ground_from(cam_pose.transToWolrd, ground_plane_sensor);
                        }
                        double obj_yaw_esti = obj_yaw_samples[obj_yaw_id];
                        // getVanishingPoints, object_3d_util.cpp:928-937
                        V2 vps[3];
                        {
                            V3 a = mul(cam_pose.KinvR, V3{std::cos(obj_yaw_esti), std::sin(obj_yaw_esti), 0});
                            V3 b = mul(cam_pose.KinvR, V3{-std::sin(obj_yaw_esti), std::cos(obj_yaw_esti), 0});
                            V3 c3 = mul(cam_pose.KinvR, V3{0, 0, 1});
                            vps[0] = {a.x / a.z, a.y / a.z}; vps[1] = {b.x / b.z, b.y / b.z}; vps[2] = {c3.x / c3.z, c3.y / c3.z};
                        }
                        V2 vp_1 = vps[0], vp_2 = vps[1], vp_3 = vps[2];
                        double vp_bound[6];
                        VP_support_edge_infos(vps, edge_mid_pts, lines_inobj_angles, vp12_edge_angle_thre, vp3_edge_angle_thre, vp_bound);
                        int group_id = (cam_roll_id * (int)cam_pitch_samples.size() + cam_pitch_id) * (int)obj_yaw_samples.size() + obj_yaw_id;

                        for (int sample_top_pt_id = 0; sample_top_pt_id < n_top; sample_top_pt_id++) {
                            TR.n_enum += (int)all_configs[0] + (int)all_configs[1];
                            V2 corner_1_top{(double)top_x_samples[sample_top_pt_id], (double)top_y_raw};
                            int vp_1_position = 0;
                            V2 corner_2_top = seg_hit_boundary(vp_1, corner_1_top, right_x_raw, top_y_raw, right_x_raw, down_y_expan);
                            if (corner_2_top.x == -1) {
                                corner_2_top = seg_hit_boundary(vp_1, corner_1_top, left_x_raw, top_y_raw, left_x_raw, down_y_expan);
                                if (corner_2_top.x != -1) vp_1_position = 2;
                            } else
                                vp_1_position = 1;
                            if (!(vp_1_position > 0)) continue;
                            if (norm(corner_1_top - corner_2_top) < shorted_edge_thre) continue;

                            for (int config_id = 1; config_id < 3; config_id++) {
                                if (!all_configs[config_id - 1]) continue;
                                V2 corner_3_top, corner_4_top;
                                if (config_id == 1) {
                                    if (vp_1_position == 1) corner_4_top = seg_hit_boundary(vp_2, corner_1_top, left_x_raw, top_y_raw, left_x_raw, down_y_expan);
                                    else corner_4_top = seg_hit_boundary(vp_2, corner_1_top, right_x_raw, top_y_raw, right_x_raw, down_y_expan);
                                    if (corner_4_top.y == -1) continue;
                                    if (norm(corner_1_top - corner_4_top) < shorted_edge_thre) continue;
                                    corner_3_top = lineSegmentIntersect(vp_2, corner_2_top, vp_1, corner_4_top);
                                    if (!check_inside_box(corner_3_top, V2{(double)left_x_raw, (double)top_y_raw}, V2{(double)right_x_raw, (double)down_y_expan})) continue;
                                    if ((norm(corner_3_top - corner_4_top) < shorted_edge_thre) || (norm(corner_3_top - corner_2_top) < shorted_edge_thre)) continue;
                                }
                                if (config_id == 2) {
                                    if (vp_1_position == 1) corner_3_top = seg_hit_boundary(vp_2, corner_2_top, left_x_raw, top_y_raw, left_x_raw, down_y_expan);
                                    else corner_3_top = seg_hit_boundary(vp_2, corner_2_top, right_x_raw, top_y_raw, right_x_raw, down_y_expan);
                                    if (corner_3_top.y == -1) continue;
                                    if (norm(corner_2_top - corner_3_top) < shorted_edge_thre) continue;
                                    corner_4_top = lineSegmentIntersect(vp_1, corner_3_top, vp_2, corner_1_top);
                                    if (!check_inside_box(corner_4_top, V2{(double)left_x_raw, (double)top_y_expan_distmap}, V2{(double)right_x_raw, (double)down_y_expan_distmap})) continue;
                                    if ((norm(corner_3_top - corner_4_top) < shorted_edge_thre) || (norm(corner_4_top - corner_1_top) < shorted_edge_thre)) continue;
                                }
                                V2 corner_5_down = seg_hit_boundary(vp_3, corner_3_top, left_x_raw, down_y_expan, right_x_raw, down_y_expan);
                                if (corner_5_down.y == -1) continue;
                                if (norm(corner_3_top - corner_5_down) < shorted_edge_thre) continue;
                                V2 corner_6_down = lineSegmentIntersect(vp_2, corner_5_down, vp_3, corner_2_top);
                                if (!check_inside_box(corner_6_down, expan_lt, expan_rb)) continue;
                                if ((norm(corner_6_down - corner_2_top) < shorted_edge_thre) || (norm(corner_6_down - corner_5_down) < shorted_edge_thre)) continue;
                                V2 corner_7_down = lineSegmentIntersect(vp_1, corner_6_down, vp_3, corner_1_top);
                                if (!check_inside_box(corner_7_down, expan_lt, expan_rb)) continue;
                                if ((norm(corner_7_down - corner_1_top) < shorted_edge_thre) || (norm(corner_7_down - corner_6_down) < shorted_edge_thre)) continue;
                                V2 corner_8_down = lineSegmentIntersect(vp_1, corner_5_down, vp_2, corner_7_down);
                                if (!check_inside_box(corner_8_down, expan_lt, expan_rb)) continue;
                                if ((norm(corner_8_down - corner_4_top) < shorted_edge_thre) || (norm(corner_8_down - corner_5_down) < shorted_edge_thre) ||
                                    (norm(corner_8_down - corner_7_down) < shorted_edge_thre))
                                    continue;

                                V2 c[8] = {corner_1_top, corner_2_top, corner_3_top, corner_4_top, corner_5_down, corner_6_down, corner_7_down, corner_8_down};
                                V2 cs[8];
                                for (int i = 0; i < 8; i++) cs[i] = {c[i].x - left_x_expan_distmap, c[i].y - top_y_expan_distmap};
                                double sum_dist, total_angle_diff;
                                if (config_id == 1) {
                                    static const int vis[9][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {1, 5}, {2, 4}, {3, 7}, {4, 7}, {4, 5}};
                                    static const int vpe[3][4] = {{0, 1, 7, 4}, {3, 0, 4, 5}, {3, 7, 1, 5}};
                                    sum_dist = box_edge_sum_dists(dist_map, height_expan_distmap, width_expan_distmap, cs, vis, 9, false);
                                    total_angle_diff = box_edge_alignment_angle_error(vp_bound, vpe, c);
                                } else {
                                    static const int vis[7][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {1, 5}, {2, 4}, {4, 5}};
                                    static const int vpe[3][4] = {{0, 1, 2, 3}, {3, 0, 4, 5}, {2, 4, 1, 5}};
                                    sum_dist = box_edge_sum_dists(dist_map, height_expan_distmap, width_expan_distmap, cs, vis, 7, reweight_edge_distance);
                                    total_angle_diff = box_edge_alignment_angle_error(vp_bound, vpe, c);
                                }
                                double row[9] = {(double)config_id, (double)vp_1_position, obj_yaw_esti, (double)sample_top_pt_id, sum_dist / obj_diaglength_expan,
                                                 total_angle_diff, (double)down_expand_sample, 0, 0};
                                if (P.whether_sample_cam_roll_pitch) { row[7] = cam_roll_samples[cam_roll_id]; row[8] = cam_pitch_samples[cam_pitch_id]; }
                                else { row[7] = cam_pose_raw.euler_angle[0]; row[8] = cam_pose_raw.euler_angle[1]; }
                                TR.rows.insert(TR.rows.end(), row, row + 9);
                                for (int i = 0; i < 8; i++) TR.corners.push_back(c[i].x);
                                for (int i = 0; i < 8; i++) TR.corners.push_back(c[i].y);
                                TR.hyp_id.push_back((group_id * n_top + sample_top_pt_id) * 2 + (config_id - 1));
                                valid_n++;
                            }
                        }
                    }

            std::vector<double> dist_col(valid_n), angle_col(valid_n);
            for (int i = 0; i < valid_n; i++) { dist_col[i] = TR.rows[9 * i + 4]; angle_col[i] = TR.rows[9 * i + 5]; }
            fuse_normalize_scores_v2(dist_col, angle_col, TR.norm_score, TR.keep, weight_vp_angle, whether_normalize_two_errors);
            R.n_scored += valid_n;
            R.n_enum += TR.n_enum;

            for (int box_id = 0; box_id < (int)TR.keep.size(); box_id++) {
                int raw_cube_ind = TR.keep[box_id];
                const double* row = &TR.rows[9 * raw_cube_ind];
                if (P.whether_sample_cam_roll_pitch) {
                    M4 Tn = transToWolrd;
                    M3 Rn = euler_zyx_to_rot(row[7], row[8], cam_pose_raw.euler_angle[2]);
                    for (int i = 0; i < 3; i++)
                        for (int j = 0; j < 3; j++) Tn(i, j) = Rn(i, j);
                    set_cam_pose(cam_pose, Tn);
                    ground_from(cam_pose.transToWolrd, ground_plane_sensor);
                }
                orc_cuboid o;
                std::memset(&o, 0, sizeof o);
                V2 c[8];
                for (int i = 0; i < 8; i++) c[i] = {TR.corners[16 * raw_cube_ind + i], TR.corners[16 * raw_cube_ind + 8 + i]};
                change_2d_corner_to_3d_object(c, row[0], row[1], row[2], ground_plane_sensor, cam_pose.transToWolrd, cam_pose.invK, o);
                if ((o.scale[0] < 0) || (o.scale[1] < 0) || (o.scale[2] < 0)) continue;
                o.rect_detect_2d[0] = left_x_raw; o.rect_detect_2d[1] = top_y_raw; o.rect_detect_2d[2] = obj_width_raw; o.rect_detect_2d[3] = obj_height_raw;
                o.edge_distance_error = row[4];
                o.edge_angle_error = row[5];
                o.normalized_error = TR.norm_score[box_id];
                o.skew_ratio = std::max(o.scale[0], o.scale[1]) / std::min(o.scale[0], o.scale[1]);
                o.down_expand_height = row[6];
                if (P.whether_sample_cam_roll_pitch) {
                    o.camera_roll_delta = row[7] - cam_pose_raw.euler_angle[0];
                    o.camera_pitch_delta = row[8] - cam_pose_raw.euler_angle[1];
                } else { o.camera_roll_delta = 0; o.camera_pitch_delta = 0; }
                o.task_id = task_id; o.raw_cube_ind = raw_cube_ind;
                BR.raw.push_back(o);
            }
        }  // height samples

        // final ranking, box_proposal_detail.cpp:804-838
        int actual_cuboid_num_small = std::min(P.max_cuboid_num, (int)BR.raw.size());
        BR.combined.resize(BR.raw.size());
        for (int i = 0; i < (int)BR.raw.size(); i++) {
            const orc_cuboid& o = BR.raw[i];
            double skew_error = weight_skew_error * std::max(o.skew_ratio - P.nominal_skew_ratio, 0.0);
            if (o.skew_ratio > P.max_cut_skew) skew_error = 100;
            BR.combined[i] = o.normalized_error + weight_skew_error * skew_error;
        }
        std::vector<int> sort_idx_small(BR.raw.size());
        std::iota(sort_idx_small.begin(), sort_idx_small.end(), 0);
        sort_indexes(BR.combined, sort_idx_small, actual_cuboid_num_small);
        BR.sorted.assign(sort_idx_small.begin(), sort_idx_small.begin() + actual_cuboid_num_small);
    }
}

}  // namespace

extern "C" {

int orc_plan(const double* boxes, int n_boxes, int img_w, int img_h, int sample_height, orc_task* out, int max_tasks) {
    std::vector<orc_task> t;
    plan_tasks(boxes, n_boxes, img_w, img_h, sample_height != 0, t);
    if ((int)t.size() > max_tasks) return -(int)t.size();
    for (size_t i = 0; i < t.size(); i++) out[i] = t[i];
    return (int)t.size();
}

void* orc_detect_frame(const double* K, const double* T, int img_w, int img_h, const double* boxes, int n_boxes, const double* lines, int n_lines,
                       const float* dist_maps, const orc_params* P) {
    FrameResult* R = new FrameResult();
    g_libm_atan2 = P->libm_atan2 != 0;
    detect_cuboid(K, T, img_w, img_h, boxes, n_boxes, lines, n_lines, dist_maps, *P, *R);
    return R;
}
void orc_free(void* h) { delete (FrameResult*)h; }
int orc_num_tasks(void* h) { return (int)((FrameResult*)h)->tasks.size(); }
long long orc_num_scored(void* h) { return ((FrameResult*)h)->n_scored; }
long long orc_num_enum(void* h) { return ((FrameResult*)h)->n_enum; }
void orc_task_info(void* h, int t, orc_task* out, int* n_valid, int* n_enum, int* n_merged, int* n_keep) {
    TaskResult& TR = ((FrameResult*)h)->tasks[t];
    *out = TR.task; *n_valid = (int)TR.hyp_id.size(); *n_enum = TR.n_enum; *n_merged = (int)TR.merged.size(); *n_keep = (int)TR.keep.size();
}
void orc_task_data(void* h, int t, double* rows, double* corners, int* hyp_id, double* merged, int* keep, double* norm_score) {
    TaskResult& TR = ((FrameResult*)h)->tasks[t];
    if (rows) std::copy(TR.rows.begin(), TR.rows.end(), rows);
    if (corners) std::copy(TR.corners.begin(), TR.corners.end(), corners);
    if (hyp_id) std::copy(TR.hyp_id.begin(), TR.hyp_id.end(), hyp_id);
    if (merged) for (size_t i = 0; i < TR.merged.size(); i++) { merged[4 * i] = TR.merged[i].x1; merged[4 * i + 1] = TR.merged[i].y1; merged[4 * i + 2] = TR.merged[i].x2; merged[4 * i + 3] = TR.merged[i].y2; }
    if (keep) std::copy(TR.keep.begin(), TR.keep.end(), keep);
    if (norm_score) std::copy(TR.norm_score.begin(), TR.norm_score.end(), norm_score);
}
void orc_box_info(void* h, int b, int* n_raw, int* n_sorted) {
    BoxResult& B = ((FrameResult*)h)->boxes[b];
    *n_raw = (int)B.raw.size(); *n_sorted = (int)B.sorted.size();
}
void orc_box_data(void* h, int b, orc_cuboid* raw, double* combined, int* sorted) {
    BoxResult& B = ((FrameResult*)h)->boxes[b];
    if (raw) std::copy(B.raw.begin(), B.raw.end(), raw);
    if (combined) std::copy(B.combined.begin(), B.combined.end(), combined);
    if (sorted) std::copy(B.sorted.begin(), B.sorted.end(), sorted);
}

// Batch driver for the CPU baseline: frames are independent (the reference processes them one
// after another on one thread; n_threads>1 distributes frames over std::threads).
// frame f: K[9f..], T[16f..], boxes[box_off[f]..box_off[f+1]) x5, lines[line_off[f]..line_off[f+1]) x4,
// dist maps packed per frame starting at map_off[f] (floats).  Returns total scored proposals; per-box best
// cuboid written to best[] (n_total_boxes entries; task_id = -1 if none).
long long orc_detect_batch(int n_frames, const double* K, const double* T, int img_w, int img_h, const double* boxes, const int* box_off,
                           const double* lines, const int* line_off, const float* dist_maps, const long long* map_off, const orc_params* P,
                           int n_threads, orc_cuboid* best, long long* n_enum_out) {
    g_libm_atan2 = P->libm_atan2 != 0;
    std::vector<long long> scored(n_frames, 0), enumd(n_frames, 0);
    auto work = [&](int tid) {
        for (int f = tid; f < n_frames; f += n_threads) {
            FrameResult R;
            int nb = box_off[f + 1] - box_off[f];
            detect_cuboid(K + 9 * f, T + 16 * f, img_w, img_h, boxes + 5 * (size_t)box_off[f], nb, lines + 4 * (size_t)line_off[f],
                          line_off[f + 1] - line_off[f], dist_maps + map_off[f], *P, R);
            scored[f] = R.n_scored; enumd[f] = R.n_enum;
            if (best)
                for (int b = 0; b < nb; b++) {
                    orc_cuboid& o = best[box_off[f] + b];
                    if (!R.boxes[b].sorted.empty()) o = R.boxes[b].raw[R.boxes[b].sorted[0]];
                    else { std::memset(&o, 0, sizeof o); o.task_id = -1; }
                }
        }
    };
    if (n_threads <= 1) { n_threads = 1; work(0); }
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(work, t);
        for (auto& t : th) t.join();
    }
    long long tot = 0, tote = 0;
    for (int f = 0; f < n_frames; f++) { tot += scored[f]; tote += enumd[f]; }
    if (n_enum_out) *n_enum_out = tote;
    return tot;
}

// KAT hooks (tests/test_oracle_golden.py)
double orc_det_atan2(double y, double x) { return det_atan2(y, x); }
void orc_kat_ray_plane(const double* rays3xn, int n, const double* plane4, double* out3xn) {
    for (int i = 0; i < n; i++) {
        double rx = rays3xn[i], ry = rays3xn[n + i], rz = rays3xn[2 * n + i];
        double frac = -plane4[3] / ((plane4[0] * rx + plane4[1] * ry) + plane4[2] * rz);
        out3xn[i] = frac * rx; out3xn[n + i] = frac * ry; out3xn[2 * n + i] = frac * rz;
    }
}
void orc_kat_invK_pixels(const double* K9, const double* pix2xn, int n, double* rays3xn) {
    M3 K; std::memcpy(K.m, K9, sizeof K.m);
    M3 iK = inverse3(K);
    for (int i = 0; i < n; i++) {
        V3 r = mul(iK, V3{pix2xn[i], pix2xn[n + i], 1.0});
        rays3xn[i] = r.x; rays3xn[n + i] = r.y; rays3xn[2 * n + i] = r.z;
    }
}
void orc_kat_set_cam_pose(const double* K9, const double* T16, double* euler3, double* KinvR9, double* invK9) {
    CamPose c; M3 K; std::memcpy(K.m, K9, sizeof K.m); M4 T; std::memcpy(T.m, T16, sizeof T.m);
    set_calibration(c, K); set_cam_pose(c, T);
    std::memcpy(euler3, c.euler_angle, 3 * sizeof(double));
    std::memcpy(KinvR9, c.KinvR.m, sizeof c.KinvR.m);
    std::memcpy(invK9, c.invK.m, sizeof c.invK.m);
}
}  // extern "C"
