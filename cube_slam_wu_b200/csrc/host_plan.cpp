// host_plan.cpp -- host side of the proposal half: integer task planning and per-frame sweep tables.
//
// This is the product's own host logic (C++, like the reference); it never calls into oracle/.
// Follows detect_3d_cuboid/src/box_proposal_detail.cpp:45-56 (set_cam_pose), :143-256 (box / ROI integers),
// :180-184 (yaw samples), :344-355 (roll/pitch samples), :368-377 (per-sample camera pose) and
// matrix_utils.cpp:368-380 (accumulating linespace).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "csb_internal.h"
#include "csb_math.cuh"
#include "host_plan.h"

namespace csb {

template <class T>
static void linespace(T starting, T ending, T step, std::vector<T>& res) {
    while (starting <= ending) {
        res.push_back(starting);
        starting += step;
        if (res.size() > 1000) break;
    }
}

// Sample angles of a frame (cheap part of the table: one quaternion->Euler conversion and the accumulating linespaces).
static int frame_angles(const csb_frame& f, const csb_detect_params& p, double euler_raw[3], std::vector<double>& yaws, std::vector<double>& rolls,
                        std::vector<double>& pitches) {
    M3 Rraw;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rraw.m[i * 3 + j] = f.transToWolrd[i * 4 + j];
    quat_to_euler_zyx(quat_from_rot(Rraw), euler_raw[0], euler_raw[1], euler_raw[2]);
    double yaw_init = euler_raw[2] - 90.0 / 180.0 * M_PI;
    linespace<double>(yaw_init - 45.0 / 180.0 * M_PI, yaw_init + 45.0 / 180.0 * M_PI, 6.0 / 180.0 * M_PI, yaws);
    if (p.whether_sample_cam_roll_pitch) {
        linespace<double>(euler_raw[0] - 6.0 / 180.0 * M_PI, euler_raw[0] + 6.0 / 180.0 * M_PI, 3.0 / 180.0 * M_PI, rolls);
        linespace<double>(euler_raw[1] - 6.0 / 180.0 * M_PI, euler_raw[1] + 6.0 / 180.0 * M_PI, 3.0 / 180.0 * M_PI, pitches);
    } else {
        rolls.push_back(euler_raw[0]);
        pitches.push_back(euler_raw[1]);
    }
    if ((int)yaws.size() > MAX_YAW || (int)rolls.size() > MAX_RP || (int)pitches.size() > MAX_RP) return CSB_ERR_CAPACITY;
    return CSB_OK;
}

int frame_group_count(const csb_frame& f, const csb_detect_params& p, int* n_groups) {
    double e[3];
    std::vector<double> yaws, rolls, pitches;
    int rc = frame_angles(f, p, e, yaws, rolls, pitches);
    if (rc != CSB_OK) return rc;
    *n_groups = (int)(yaws.size() * rolls.size() * pitches.size());
    return CSB_OK;
}

int build_frame_tab(const csb_frame& f, const csb_detect_params& p, FrameTab& ft) {
    std::memset(&ft, 0, sizeof ft);
    M3 K;
    std::memcpy(K.m, f.Kalib, sizeof K.m);
    M3 invK = inverse3(K);
    std::memcpy(ft.invK, invK.m, sizeof ft.invK);
    M3 Rraw;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Rraw.m[i * 3 + j] = f.transToWolrd[i * 4 + j];
    quat_to_euler_zyx(quat_from_rot(Rraw), ft.euler_raw[0], ft.euler_raw[1], ft.euler_raw[2]);

    // yaw samples: facing the camera +-45 deg every 6 deg (box_proposal_detail.cpp:180-184).  The reference reads
    // cam_pose.camera_yaw, which after the first box holds the yaw re-extracted from the last sampled pose (a
    // <=1-ulp state leak); boxes are independent here and use the raw camera yaw (DESIGN.md "known deviations").
    double yaw_init = ft.euler_raw[2] - 90.0 / 180.0 * M_PI;
    std::vector<double> yaws;
    linespace<double>(yaw_init - 45.0 / 180.0 * M_PI, yaw_init + 45.0 / 180.0 * M_PI, 6.0 / 180.0 * M_PI, yaws);
    if ((int)yaws.size() > MAX_YAW) return CSB_ERR_CAPACITY;
    ft.n_yaw = (int)yaws.size();
    for (int i = 0; i < ft.n_yaw; i++) {
        ft.yaw[i] = yaws[i];
        ft.cosy[i] = std::cos(yaws[i]);
        ft.siny[i] = std::sin(yaws[i]);
    }
    std::vector<double> rolls, pitches;
    ft.sample_rp = p.whether_sample_cam_roll_pitch ? 1 : 0;
    if (ft.sample_rp) {
        linespace<double>(ft.euler_raw[0] - 6.0 / 180.0 * M_PI, ft.euler_raw[0] + 6.0 / 180.0 * M_PI, 3.0 / 180.0 * M_PI, rolls);
        linespace<double>(ft.euler_raw[1] - 6.0 / 180.0 * M_PI, ft.euler_raw[1] + 6.0 / 180.0 * M_PI, 3.0 / 180.0 * M_PI, pitches);
    } else {
        rolls.push_back(ft.euler_raw[0]);
        pitches.push_back(ft.euler_raw[1]);
    }
    if ((int)rolls.size() > MAX_RP || (int)pitches.size() > MAX_RP) return CSB_ERR_CAPACITY;
    ft.n_roll = (int)rolls.size();
    ft.n_pitch = (int)pitches.size();
    for (int r = 0; r < ft.n_roll; r++) ft.roll[r] = rolls[r];
    for (int q = 0; q < ft.n_pitch; q++) ft.pitch[q] = pitches[q];
    for (int r = 0; r < ft.n_roll; r++)
        for (int q = 0; q < ft.n_pitch; q++) {
            int pair = r * ft.n_pitch + q;
            M3 Rn = ft.sample_rp ? euler_zyx_to_rot(rolls[r], pitches[q], ft.euler_raw[2]) : Rraw;
            M3 invR = inverse3(Rn);
            M3 KinvR = mul(K, invR);
            std::memcpy(ft.KinvR[pair], KinvR.m, sizeof KinvR.m);
            for (int i = 0; i < 3; i++) {
                for (int j = 0; j < 3; j++) ft.Tnew[pair][i * 4 + j] = Rn.m[i * 3 + j];
                ft.Tnew[pair][i * 4 + 3] = f.transToWolrd[i * 4 + 3];
            }
        }
    ft.line_begin = f.line_begin;
    ft.line_end = f.line_end;
    ft.img_w = f.img_width;
    ft.img_h = f.img_height;
    return CSB_OK;
}


int plan_tasks(const csb_frame* frames, int n_frames, const double* boxes, int n_boxes, const csb_detect_params& p, std::vector<csb_task>& tasks,
               std::vector<TaskTab>* tabs, int64_t* n_map_floats) {
    int64_t map_off = 0, out_off = 0;
    int line_cap_off = 0;
    const int n_cfg = (p.consider_config_1 ? 1 : 0) + (p.consider_config_2 ? 1 : 0);
    for (int f = 0; f < n_frames; f++) {
        const csb_frame& fr = frames[f];
        if (fr.box_begin < 0 || fr.box_end > n_boxes || fr.box_begin > fr.box_end) return CSB_ERR_INVALID;
        // frames own disjoint, ascending box ranges: a box's tasks must be contiguous (box -> task prefix sums, k_rank, k_observe)
        if (f > 0 && fr.box_begin < frames[f - 1].box_end) return CSB_ERR_INVALID;
        int n_groups = 0;
        int rc = frame_group_count(fr, p, &n_groups);
        if (rc != CSB_OK) return rc;
        if (n_groups > MAX_GROUPS) return CSB_ERR_CAPACITY;
        const int img_width = fr.img_width, img_height = fr.img_height;
        for (int b = fr.box_begin; b < fr.box_end; b++) {
            const double* bx = boxes + 5 * (size_t)b;
            // box_proposal_detail.cpp:143-149
            int left_x_raw = (int)bx[0], top_y_raw = (int)bx[1], obj_width_raw = (int)bx[2], obj_height_raw = (int)bx[3];
            int right_x_raw = (int)(left_x_raw + bx[2]);
            // :160-172
            std::vector<int> down_expand_sample_all;
            down_expand_sample_all.push_back(0);
            if (p.whether_sample_bbox_height) {
                int r = std::max(std::min(20, obj_height_raw - 90), 20);
                r = std::min(r, img_height - top_y_raw - obj_height_raw - 1);
                if (r > 10) down_expand_sample_all.push_back((int)std::round(r / 2));
                down_expand_sample_all.push_back(r);
            }
            for (int hs = 0; hs < (int)down_expand_sample_all.size(); hs++) {
                int down_expand_sample = down_expand_sample_all[hs];
                int obj_height_expan = obj_height_raw + down_expand_sample;  // :204-207
                int down_y_expan = top_y_raw + obj_height_expan;
                double diag = std::sqrt((double)(obj_width_raw * obj_width_raw + obj_height_expan * obj_height_expan));
                int top_sample_resolution = (int)std::round(std::min(20, obj_width_raw / 10));  // :212
                if (top_sample_resolution < 1) break;                                            // :215-216
                int n_tops = 0;  // linespace<int>(left_x_raw + 5, right_x_raw - 5, top_sample_resolution) :219, counted
                for (int x = left_x_raw + 5; x <= right_x_raw - 5; x += top_sample_resolution)
                    if (++n_tops > 1000) break;
                // :242-248
                int w = std::min(std::max(std::min(20, obj_width_raw - 100), 10), std::max(std::min(20, obj_height_expan - 100), 10));
                int left_e = std::max(0, left_x_raw - w), right_e = std::min(img_width - 1, right_x_raw + w);
                int top_e = std::max(0, top_y_raw - w), down_e = std::min(img_height - 1, down_y_expan + w);
                int height_e = down_e - top_e, width_e = right_e - left_e;
                if (width_e <= 0 || height_e <= 0) return CSB_ERR_INVALID;  // cv::Rect would be empty; the reference would throw in cv::Canny

                csb_task t;
                std::memset(&t, 0, sizeof t);
                t.frame_id = f; t.box_id = b; t.hs_id = hs; t.down_expand = down_expand_sample;
                t.roi_left = left_e; t.roi_top = top_e; t.roi_width = width_e; t.roi_height = height_e;
                t.n_top = n_tops;
                t.n_enum = n_groups * t.n_top * n_cfg;
                t.map_offset = map_off;
                tasks.push_back(t);
                if (tabs) {
                    TaskTab tt;
                    std::memset(&tt, 0, sizeof tt);
                    tt.frame_id = f; tt.box_id = b; tt.hs_id = hs; tt.down_expand = down_expand_sample;
                    tt.left_x_raw = left_x_raw; tt.top_y_raw = top_y_raw; tt.right_x_raw = right_x_raw; tt.down_y_expan = down_y_expan;
                    tt.obj_width_raw = obj_width_raw; tt.obj_height_raw = obj_height_raw;
                    tt.roi_left = left_e; tt.roi_top = top_e; tt.roi_right = right_e; tt.roi_down = down_e; tt.roi_w = width_e; tt.roi_h = height_e;
                    tt.n_top = t.n_top; tt.top_x0 = left_x_raw + 5; tt.top_step = top_sample_resolution;
                    tt.n_hyp = n_groups * t.n_top * 2;
                    tt.n_enum = t.n_enum;
                    tt.cfg_mask = (p.consider_config_1 ? 1 : 0) | (p.consider_config_2 ? 2 : 0);
                    tt.line_cap_offset = line_cap_off;
                    tt.diag = diag;
                    tt.map_offset = map_off;
                    tt.out_offset = out_off;
                    tabs->push_back(tt);
                }
                int64_t px = (int64_t)width_e * height_e;
                map_off += (px + 3) & ~(int64_t)3;  // keep every map 16-byte aligned for cp.async.bulk
                out_off += (int64_t)n_groups * t.n_top * 2;
                line_cap_off += (fr.line_end - fr.line_begin);
            }
        }
    }
    if (n_map_floats) *n_map_floats = map_off;
    return CSB_OK;
}

}  // namespace csb
