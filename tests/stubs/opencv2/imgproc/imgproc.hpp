// TEST STUB: cv::cvtColor BGR -> gray with OpenCV's 8-bit fixed-point weights (R 4899, G 9617, B 1868, shift 14).
#pragma once
#include "../core.hpp"
namespace cv {
inline void cvtColor(const Mat& src, Mat& dst, int /*code*/) {
    dst.create(src.rows, src.cols, CV_8UC1);
    for (size_t i = 0; i < src.total(); i++) {
        const unsigned char* p = src.data + 3 * i;
        dst.data[i] = (unsigned char)((p[0] * 1868 + p[1] * 9617 + p[2] * 4899 + (1 << 13)) >> 14);
    }
}
}  // namespace cv
