// Mat.h -- the sliver of cv::Mat the tracking path uses (rows, cols, ptr<T>(row), zeros, clone), so that host code
// written against the reference's frame / PixelWisePyramid members compiles without OpenCV.  With OpenCV available a
// maintainer keeps cv::Mat and passes .data / .ptr<T>(0) to the C-ABI directly.
#pragma once

#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

namespace ellc_host {

enum { CV_8UC1 = 0, CV_32FC1 = 5 };
typedef unsigned char uchar;

class Mat {
public:
    int rows, cols, type_;
    Mat() : rows(0), cols(0), type_(CV_8UC1) {}
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf_(std::make_shared<std::vector<uint8_t> >((size_t)r * c * elem(type))) {}
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }        // vector value-initialises to 0
    bool empty() const { return !buf_ || buf_->empty(); }
    size_t elemSize() const { return elem(type_); }
    Mat clone() const { Mat m(rows, cols, type_); if (buf_) *m.buf_ = *buf_; return m; }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(buf_->data() + (size_t)r * cols * elemSize()); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(buf_->data() + (size_t)r * cols * elemSize()); }
    uint8_t* data() { return buf_ ? buf_->data() : nullptr; }
    const uint8_t* data() const { return buf_ ? buf_->data() : nullptr; }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }

private:
    static size_t elem(int type) { return type == CV_32FC1 ? 4 : 1; }
    std::shared_ptr<std::vector<uint8_t> > buf_;      // shallow copies share pixels, like cv::Mat headers
};

}  // namespace ellc_host
