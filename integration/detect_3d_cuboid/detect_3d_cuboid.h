// Drop-in replacement for detect_3d_cuboid/include/detect_3d_cuboid/detect_3d_cuboid.h of the reference.
//
// Same class names, members, public flags and detect_cuboid() signature (reference header lines 20-41, 59-71, 74-118), so
// object_slam/src/main_obj.cpp:492-500, 633-669 and detect_3d_cuboid/src/main.cpp:62-72 compile unchanged.  Everything inside
// detect_cuboid() -- cv::Canny + cv::distanceTransform per box ROI (box_proposal_detail.cpp:320-327), the sweep, scoring, selection and
// 3D recovery -- runs on the GPU through csb_detect_batch_gray() of include/cubeslam_b200.h; this header only converts types.
// box_proposal_detail.cpp drops out of the build: set_calibration / set_cam_pose (:36-56) are defined here.
//
// Compiled by tests/test_adapters_compile.py against the interface stubs under tests/stubs/ (the build container has no Eigen / OpenCV).
#pragma once

#include <cmath>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Dense>
#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>

#include "cubeslam_b200.h"

class cuboid  // identical to the reference (detect_3d_cuboid.h:20-41)
{
public:
    Eigen::Vector3d pos;
    Eigen::Vector3d scale;
    double rotY;
    Eigen::Vector2d box_config_type;
    Eigen::Matrix2Xi box_corners_2d;
    Eigen::Matrix3Xd box_corners_3d_world;
    Eigen::Vector4d rect_detect_2d;
    double edge_distance_error;
    double edge_angle_error;
    double normalized_error;
    double skew_ratio;
    double down_expand_height;
    double camera_roll_delta;
    double camera_pitch_delta;
    void print_cuboid();
};
typedef std::vector<cuboid*> ObjectSet;

struct cam_pose_infos  // detect_3d_cuboid.h:59-71
{
    Eigen::Matrix4d transToWolrd;
    Eigen::Matrix3d Kalib;
    Eigen::Matrix3d rotationToWorld;
    Eigen::Vector3d euler_angle;
    Eigen::Matrix3d invR;
    Eigen::Matrix3d invK;
    Eigen::Matrix<double, 3, 4> projectionMatrix;
    Eigen::Matrix3d KinvR;
    double camera_yaw;
};

class detect_3d_cuboid
{
public:
    cam_pose_infos cam_pose;
    cam_pose_infos cam_pose_raw;

    // The reference's constructor cannot fail; this one needs a CUDA device (there is no CPU fallback) and says so at once instead of
    // returning empty ObjectSets later.
    detect_3d_cuboid() {
        if (csb_create(&ctx_, 0) != CSB_OK) throw std::runtime_error("detect_3d_cuboid (B200): no usable CUDA device");
    }
    ~detect_3d_cuboid() { csb_destroy(ctx_); }
    detect_3d_cuboid(const detect_3d_cuboid&) = delete;
    detect_3d_cuboid& operator=(const detect_3d_cuboid&) = delete;

    void set_calibration(const Eigen::Matrix3d& Kalib) { cam_pose.Kalib = Kalib; cam_pose.invK = Kalib.inverse(); }  // box_proposal_detail.cpp:36-40
    // box_proposal_detail.cpp:45-56 with quat_to_euler_zyx (matrix_utils.cpp:38-51) written out
    void set_cam_pose(const Eigen::Matrix4d& transToWolrd)
    {
        cam_pose.transToWolrd = transToWolrd;
        cam_pose.rotationToWorld = transToWolrd.topLeftCorner<3, 3>();
        const Eigen::Quaterniond q(cam_pose.rotationToWorld);
        const double qw = q.w(), qx = q.x(), qy = q.y(), qz = q.z();
        Eigen::Vector3d euler_angles;
        euler_angles(0) = atan2(2 * (qw * qx + qy * qz), 1 - 2 * (qx * qx + qy * qy));
        euler_angles(1) = asin(2 * (qw * qy - qz * qx));
        euler_angles(2) = atan2(2 * (qw * qz + qx * qy), 1 - 2 * (qy * qy + qz * qz));
        cam_pose.euler_angle = euler_angles;
        cam_pose.invR = cam_pose.rotationToWorld.inverse();
        cam_pose.projectionMatrix = cam_pose.Kalib * transToWolrd.inverse().topRows<3>();  // project world coordinate to camera
        cam_pose.KinvR = cam_pose.Kalib * cam_pose.invR;
        cam_pose.camera_yaw = cam_pose.euler_angle(2);
    }

    // Same contract as the reference: all_object_cuboids is resized to the number of boxes; each ObjectSet holds up to
    // max_cuboid_num new-ed cuboids (caller-owned, as in the reference); an empty ObjectSet means "nothing found".
    void detect_cuboid(const cv::Mat& rgb_img, const Eigen::Matrix4d& transToWolrd, const Eigen::MatrixXd& obj_bbox_coors, Eigen::MatrixXd edges,
                       std::vector<ObjectSet>& all_object_cuboids)
    {
        set_cam_pose(transToWolrd);
        cam_pose_raw = cam_pose;
        cv::Mat gray;
        if (rgb_img.channels() == 3) cv::cvtColor(rgb_img, gray, CV_BGR2GRAY); else gray = rgb_img;

        const int n_boxes = (int)obj_bbox_coors.rows(), n_lines = (int)edges.rows();
        all_object_cuboids.assign(n_boxes, ObjectSet());
        csb_frame fr;
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) fr.Kalib[i * 3 + j] = cam_pose.Kalib(i, j);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) fr.transToWolrd[i * 4 + j] = transToWolrd(i, j);
        fr.img_width = rgb_img.cols; fr.img_height = rgb_img.rows;
        fr.box_begin = 0; fr.box_end = n_boxes; fr.line_begin = 0; fr.line_end = n_lines;
        std::vector<double> boxes(5 * (size_t)n_boxes), lines(4 * (size_t)n_lines);  // Eigen is column-major: repack row-major
        for (int b = 0; b < n_boxes; b++) for (int j = 0; j < 5; j++) boxes[5 * b + j] = obj_bbox_coors(b, j);
        for (int l = 0; l < n_lines; l++) for (int j = 0; j < 4; j++) lines[4 * l + j] = edges(l, j);

        csb_detect_params p;
        p.consider_config_1 = consider_config_1; p.consider_config_2 = consider_config_2;
        p.whether_sample_cam_roll_pitch = whether_sample_cam_roll_pitch; p.whether_sample_bbox_height = whether_sample_bbox_height;
        p.max_cuboid_num = max_cuboid_num; p.reserved = 0; p.nominal_skew_ratio = nominal_skew_ratio; p.max_cut_skew = max_cut_skew;

        int n_tasks = 0; int64_t n_map = 0;
        if (csb_detect_plan(&fr, 1, boxes.data(), n_boxes, &p, nullptr, 0, &n_tasks, &n_map) != CSB_OK) return;
        std::vector<csb_task> tasks(n_tasks > 0 ? n_tasks : 1);
        csb_detect_plan(&fr, 1, boxes.data(), n_boxes, &p, tasks.data(), n_tasks, &n_tasks, &n_map);
        if (!gray.isContinuous()) gray = gray.clone();  // the library takes the whole frame, row-major, no padding
        std::vector<csb_cuboid> out((size_t)std::max(1, n_boxes * max_cuboid_num));
        std::vector<int32_t> n_out(std::max(1, n_boxes));
        // Canny (80, 200) + 3x3 L2 distance transform of every box ROI + the whole proposal path (box_proposal_detail.cpp:143-838) on the GPU
        if (csb_detect_batch_gray(ctx_, &fr, 1, boxes.data(), n_boxes, lines.data(), n_lines, tasks.data(), n_tasks, gray.data, (int64_t)gray.total(), &p,
                                  out.data(), n_out.data(), nullptr) != CSB_OK)
            return;  // like the reference: no exception, empty ObjectSets (csb_last_error(ctx_) holds the reason)
        for (int b = 0; b < n_boxes; b++)
            for (int r = 0; r < n_out[b]; r++) {
                const csb_cuboid& c = out[(size_t)b * max_cuboid_num + r];
                cuboid* o = new cuboid();
                o->pos = Eigen::Vector3d(c.pos[0], c.pos[1], c.pos[2]);
                o->scale = Eigen::Vector3d(c.scale[0], c.scale[1], c.scale[2]);
                o->rotY = c.rotY;
                o->box_config_type = Eigen::Vector2d(c.box_config_type[0], c.box_config_type[1]);
                o->box_corners_2d.resize(2, 8); o->box_corners_3d_world.resize(3, 8);
                for (int k = 0; k < 8; k++) {
                    o->box_corners_2d(0, k) = c.box_corners_2d[k]; o->box_corners_2d(1, k) = c.box_corners_2d[8 + k];
                    for (int rr = 0; rr < 3; rr++) o->box_corners_3d_world(rr, k) = c.box_corners_3d_world[rr * 8 + k];
                }
                o->rect_detect_2d = Eigen::Vector4d(c.rect_detect_2d[0], c.rect_detect_2d[1], c.rect_detect_2d[2], c.rect_detect_2d[3]);
                o->edge_distance_error = c.edge_distance_error; o->edge_angle_error = c.edge_angle_error; o->normalized_error = c.normalized_error;
                o->skew_ratio = c.skew_ratio; o->down_expand_height = c.down_expand_height;
                o->camera_roll_delta = c.camera_roll_delta; o->camera_pitch_delta = c.camera_pitch_delta;
                all_object_cuboids[b].push_back(o);
            }
        // whether_plot_* / whether_save_final_images: drawing is unchanged reference code (plot_image_with_cuboid) and is left to the caller's build.
    }

    bool whether_plot_detail_images = false;  // the reference defaults to true and blocks on cv::waitKey; plotting is not part of the hot path
    bool whether_plot_final_images = false;
    bool whether_save_final_images = false;
    cv::Mat cuboids_2d_img;
    bool print_details = false;
    bool consider_config_1 = true;
    bool consider_config_2 = true;
    bool whether_sample_cam_roll_pitch = true;
    bool whether_sample_bbox_height = false;
    int max_cuboid_num = 1;
    double nominal_skew_ratio = 1;
    double max_cut_skew = 3;

private:
    csb_context* ctx_ = nullptr;
};
