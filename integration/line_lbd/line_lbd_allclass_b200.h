// line_lbd_allclass_b200.h -- drop-in for class line_lbd_detect
// (reference: line_lbd/include/line_lbd/line_lbd_allclass.h:20-60, line_lbd/class/line_lbd_allclass.cpp:130-149, 200-281).
//
// Same class name, same public members the callers touch (use_LSD, line_length_thres; object_slam/src/main_obj.cpp:503-505,
// line_lbd/src/detect_lines.cpp:61-66) and the same detect_filter_lines(const cv::Mat&, cv::Mat&) /
// detect_descrip_lines(const cv::Mat&, cv::Mat&, cv::Mat&) signatures; the work is done by csb_lsd_* / csb_edlines_* / csb_lbd_* of
// libcubeslam_b200.so: use_LSD = true -> the LSD branch, use_LSD = false (the reference's default, what object_slam selects) -> EDLines.
// Compiled by tests/test_adapters_compile.py against the interface stubs under tests/stubs/ (the build container has no OpenCV).
#pragma once
#include <opencv2/core.hpp>

#include <cstring>
#include <stdexcept>
#include <vector>

#include "cubeslam_b200.h"

class line_lbd_detect {
public:
    line_lbd_detect(int numoctaves = 1, float octaveratio = 2.0) : numoctaves_(numoctaves), octaveratio_(octaveratio) {
        if (numoctaves != 1) throw std::runtime_error("line_lbd_detect (B200): one octave only, as every caller in the reference uses");
        if (csb_create(&ctx_, 0) != CSB_OK) throw std::runtime_error("line_lbd_detect (B200): no usable CUDA device (there is no CPU fallback)");
    }
    ~line_lbd_detect() { csb_destroy(ctx_); }
    line_lbd_detect(const line_lbd_detect&) = delete;
    line_lbd_detect& operator=(const line_lbd_detect&) = delete;

    bool use_LSD = false;           // line_lbd_allclass.h:32, line_lbd_allclass.cpp:125: EDLines unless the caller switches
    float line_length_thres = 50;   // line_lbd_allclass.h:35, :126; both callers set 15

    // line_lbd_allclass.cpp:221-235: gray image in, n x 4 CV_32F [x1 y1 x2 y2] out (keylines_to_mat, :28-38)
    void detect_filter_lines(const cv::Mat& gray_img, cv::Mat& linesmat_out) {
        if (gray_img.type() != CV_8UC1) throw std::runtime_error("Error, depth image!= 0");  // LSDDetector.cpp:163-164
        cv::Mat gray = gray_img.isContinuous() ? gray_img : gray_img.clone();
        csb_lsd_params p{line_length_thres, 1, max_lines_, 0};
        lines_.resize((size_t)max_lines_ * 4);
        int32_t n = 0;
        const int rc = use_LSD ? csb_lsd_detect_batch(ctx_, gray.data, 1, gray.cols, gray.rows, &p, lines_.data(), &n, nullptr)
                               : csb_edlines_detect_batch(ctx_, gray.data, 1, gray.cols, gray.rows, &p, lines_.data(), &n, nullptr);
        if (rc == CSB_ERR_CAPACITY) {  // more segments than rows: grow once and repeat
            max_lines_ *= 4;
            return detect_filter_lines(gray_img, linesmat_out);
        }
        if (rc != CSB_OK) throw std::runtime_error(csb_last_error(ctx_));
        linesmat_out.create(n, 4, CV_32FC1);
        if (n) std::memcpy(linesmat_out.data, lines_.data(), (size_t)n * 16);
    }

    // line_lbd_allclass.cpp:239-260: lines and their 32-byte LBD descriptors (lbd->compute on the detector's key lines, octave 0).
    // The segments never leave the device between the two stages.  The reference's Mat overload describes every key line the detector
    // returns (no length filter): line_length_thres is not applied here either (csb_lsd_params.line_length_thres = -1).
    void detect_descrip_lines(const cv::Mat& gray_img, cv::Mat& lines_mat, cv::Mat& line_descrips) {
        if (gray_img.type() != CV_8UC1) throw std::runtime_error("Error, depth image!= 0");
        cv::Mat gray = gray_img.isContinuous() ? gray_img : gray_img.clone();
        csb_lsd_params p{-1.0f, 1, max_lines_, 0};
        lines_.resize((size_t)max_lines_ * 4);
        int32_t n = 0;
        int rc;
        if (use_LSD) {
            rc = csb_lsd_upload(ctx_, gray.data, 1, gray.cols, gray.rows, &p);
            if (rc == CSB_OK) rc = csb_lsd_run(ctx_, 0);
            if (rc == CSB_OK) rc = csb_lbd_run_on_lsd(ctx_, /*want_float=*/0, 0);
            if (rc == CSB_OK) rc = csb_lsd_download(ctx_, lines_.data(), &n, nullptr);
        } else {  // EDLines key lines described from the detector's own fields (binary_descriptor.cpp:1045-1140)
            rc = csb_edlines_upload(ctx_, gray.data, 1, gray.cols, gray.rows, &p);
            if (rc == CSB_OK) rc = csb_edlines_run(ctx_, 0);
            if (rc == CSB_OK) rc = csb_edlines_describe(ctx_, /*want_float=*/0);
            if (rc == CSB_OK) rc = csb_edlines_download(ctx_, lines_.data(), &n, nullptr);
        }
        if (rc == CSB_ERR_CAPACITY) {
            max_lines_ *= 4;
            return detect_descrip_lines(gray_img, lines_mat, line_descrips);
        }
        if (rc != CSB_OK) throw std::runtime_error(csb_last_error(ctx_));
        lines_mat.create(n, 4, CV_32FC1);
        line_descrips.create(n, 32, CV_8UC1);
        if (n) {
            std::memcpy(lines_mat.data, lines_.data(), (size_t)n * 16);
            rc = use_LSD ? csb_lbd_download(ctx_, line_descrips.data, nullptr, nullptr, nullptr, n, nullptr)
                         : csb_edlines_download_descriptors(ctx_, line_descrips.data, nullptr, nullptr, n);
            if (rc != CSB_OK) throw std::runtime_error(csb_last_error(ctx_));
        }
    }

private:
    int numoctaves_;
    float octaveratio_;
    csb_context* ctx_ = nullptr;
    int max_lines_ = 4096;
    std::vector<float> lines_;
};
