// TEST STUB (tests/stubs): the part of cv::Mat the adapter headers use (8-bit / float matrices, ROI views are not needed by the GPU path).
#pragma once
#include <cstddef>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_BGR2GRAY 6
namespace cv {
class Mat {
public:
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * elemSize() + 16);
        data = buf_->data();
    }
    int type() const { return type_; }
    int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
    size_t elemSize() const { return type_ == CV_32FC1 ? 4 : (type_ == CV_8UC3 ? 3 : 1); }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return true; }
    Mat clone() const { Mat m(rows, cols, type_); if (data) std::memcpy(m.data, data, (size_t)rows * cols * elemSize()); return m; }
    template <class T> T* ptr(int r) { return reinterpret_cast<T*>(data + (size_t)r * cols * elemSize()); }
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
private:
    int type_ = CV_8UC1;
    std::shared_ptr<std::vector<unsigned char>> buf_;
};
}  // namespace cv
