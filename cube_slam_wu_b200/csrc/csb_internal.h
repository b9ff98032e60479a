// csb_internal.h -- tables shared between the host planner and the kernels (not part of the C ABI).
#pragma once
#include <cstdint>

#include "../../include/cubeslam_b200.h"

#define CSB_MAX_CHUNKS 16  // slices of the streamed distance-map upload (csb_detect_batch)

namespace csb {

constexpr int MAX_RP = 8;      // max camera roll (or pitch) samples; the reference yields 4..5 (matrix_utils.cpp:368-380)
constexpr int MAX_PAIRS = MAX_RP * MAX_RP;
constexpr int MAX_YAW = 32;    // object yaw samples; the reference yields 15..16 (box_proposal_detail.cpp:180-184)
constexpr int MAX_GROUPS = 800;  // (roll,pitch,yaw) groups staged in shared memory per task

// Per-frame sweep tables.  All trigonometry that defines the sweep is evaluated on the host (glibc), so the
// kernels only do +,-,*,/,sqrt on these values (SURVEY.md 7 "libm vs CUDA math").
struct FrameTab {
    double invK[9];
    double euler_raw[3];  // cam_pose_raw.euler_angle
    double roll[MAX_RP], pitch[MAX_RP];
    double yaw[MAX_YAW], cosy[MAX_YAW], siny[MAX_YAW];
    double KinvR[MAX_PAIRS][9];   // cam_pose.KinvR after set_cam_pose(transToWolrd_new) for pair (roll_id*n_pitch+pitch_id)
    double Tnew[MAX_PAIRS][12];   // top three rows of transToWolrd_new (row-major 3x4); row 3 is 0 0 0 1
    int32_t n_roll, n_pitch, n_yaw, sample_rp;
    int32_t line_begin, line_end, img_w, img_h;
    int64_t gray_offset;  // byte offset of this frame in the packed gray buffer (csb_detect_upload_gray)
};

// Per-(box, height sample) task geometry: the integer logic of box_proposal_detail.cpp:143-256.
struct TaskTab {
    int32_t frame_id, box_id, hs_id, down_expand;
    int32_t left_x_raw, top_y_raw, right_x_raw, down_y_expan, obj_width_raw, obj_height_raw;
    int32_t roi_left, roi_top, roi_right, roi_down, roi_w, roi_h;  // *_expan_distmap
    int32_t n_top, top_x0, top_step;  // top_x_samples[i] = top_x0 + i*top_step
    int32_t n_hyp;                    // enumeration space: n_groups * n_top * 2
    int32_t n_enum;                   // enabled hypotheses (n_groups * n_top * #enabled configs)
    int32_t cfg_mask;                 // bit0: config 1, bit1: config 2
    int32_t line_cap_offset;          // offset of this task's merged-line slots
    int32_t pad0;
    double diag;                      // obj_diaglength_expan
    int64_t map_offset;               // floats, 16-byte aligned
    int64_t out_offset;               // offset of this task's per-proposal slots (capacity n_hyp)
};

struct DetectConst {
    int32_t max_cuboid_num;
    int32_t whether_sample_cam_roll_pitch;
    double nominal_skew_ratio, max_cut_skew;
};

}  // namespace csb
