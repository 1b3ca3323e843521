// Stand-in for the subset of OpenCV 3.0 (core / imgproc / highgui) that the reference's tracking path uses
// (src/Frame.cpp, src/Frame.h, src/PixelWisePyramid.cpp, src/Pyramid.cpp, src/UserDefinedFunc.cpp, src/ImageFunc.cpp).
//
// TEST INFRASTRUCTURE ONLY.  The image has no OpenCV C++ headers or libraries, so the reference cannot be built as it is.
// These headers let the reference's OWN, UNMODIFIED sources compile where they lie under /root/reference (recipe:
// oracle/Makefile, target _ref/libellc_ref.so), so that the oracle restatement is checked against the reference's own
// per-pixel code, sampler quirks, band split, termination logic and pose bookkeeping.  The third-party arithmetic below is
// OUR restatement of the published OpenCV algorithms (the same ones the oracle restates; they are pinned separately against
// cv2 4.13 in tests/golden/): cv::pyrDown, Mat::inv(DECOMP_LU), the small-matrix gemm with its double accumulator, the
// per-element scalar multiply.  Everything is evaluated eagerly (no MatExpr), single-channel unless noted.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_BGR2GRAY 6
#define CV_GRAY2BGR 8
#define CV_COMP_KL_DIV 5
#define CV_8UC4 CV_MAKETYPE(CV_8U, 4)

namespace cv {

typedef unsigned char uchar;
typedef std::string String;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };

enum { DECOMP_LU = 0 };
enum { INTER_LINEAR = 1 };

class Mat {
public:
    int flags, rows, cols;
    size_t step;                       // bytes per row
    uchar* data;
    std::shared_ptr<std::vector<uchar> > buf;

    Mat() : flags(0), rows(0), cols(0), step(0), data(nullptr) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& v) {
        create(r, c, type);
        const int cn = channels();
        for (int y = 0; y < r; ++y)
            for (int x = 0; x < c * cn; ++x) {
                const double q = v.val[x % cn];
                if (depth() == CV_8U) ptr<uchar>(y)[x] = (uchar)q; else if (depth() == CV_32F) ptr<float>(y)[x] = (float)q; else ptr<double>(y)[x] = q;
            }
    }
    Mat(int r, int c, int type, void* ext) : flags(type), rows(r), cols(c), step((size_t)c * esz(type)), data((uchar*)ext) {}

    void create(int r, int c, int type) {
        flags = type; rows = r; cols = c; step = (size_t)c * esz(type);
        buf = std::make_shared<std::vector<uchar> >((size_t)r * step);      // value-initialised: zeros
        data = buf->data();
    }
    static size_t esz(int type) {
        const int depth = type & 7, cn = (type >> CV_CN_SHIFT) + 1;
        return (size_t)cn * (depth == CV_8U ? 1 : depth == CV_32F ? 4 : 8);
    }
    int type() const { return flags; }
    int depth() const { return flags & 7; }
    int channels() const { return (flags >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return esz(flags); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return true; }
    Size size() const { return Size(cols, rows); }
    size_t total() const { return (size_t)rows * cols; }

    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }

    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    template <typename T> T& at(int i) { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }        // vector-shaped Mat
    template <typename T> const T& at(int i) const { return rows == 1 ? ptr<T>(0)[i] : ptr<T>(i)[0]; }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }

    Mat clone() const {
        Mat m(rows, cols, flags);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }

    double getd(int r, int c) const {
        switch (depth()) {
            case CV_8U: return at<uchar>(r, c);
            case CV_32F: return at<float>(r, c);
            default: return at<double>(r, c);
        }
    }
    // convertTo: saturate_cast<dst>(src*alpha + beta); only the conversions the reference needs
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const {
        Mat out(rows, cols, rtype);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                const double v = getd(r, c) * alpha + beta;
                switch (out.depth()) {
                    case CV_8U: { long q = std::lrint(v); out.at<uchar>(r, c) = (uchar)std::min(255L, std::max(0L, q)); break; }
                    case CV_32F: out.at<float>(r, c) = (float)v; break;
                    default: out.at<double>(r, c) = v;
                }
            }
        dst = out;
    }

    Mat t() const {
        assert(depth() == CV_32F);
        Mat m(cols, rows, flags);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) m.at<float>(c, r) = at<float>(r, c);
        return m;
    }
    // per-element product (scale = 1)
    Mat mul(const Mat& o, double scale = 1) const {
        assert(depth() == CV_32F && o.depth() == CV_32F && rows == o.rows && cols == o.cols && scale == 1);
        Mat m(rows, cols, flags);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) m.at<float>(r, c) = at<float>(r, c) * o.at<float>(r, c);
        return m;
    }
    // a.mul(scalar): cv::multiply(Mat, Scalar) on CV_32F converts the scalar to float and multiplies in float
    Mat mul(double s) const {
        assert(depth() == CV_32F);
        const float f = (float)s;
        Mat m(rows, cols, flags);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) m.at<float>(r, c) = at<float>(r, c) * f;
        return m;
    }
    // Mat::inv(DECOMP_LU) on CV_32F: LUImpl<float> on [A | I] (partial pivoting, threshold FLT_EPSILON*10, elimination with
    // d = -1/pivot, back-substitution multiplying by the stored reciprocal); a singular matrix yields all zeros.
    Mat inv(int method = DECOMP_LU) const {
        (void)method;
        assert(depth() == CV_32F && rows == cols);
        const int m = rows;
        std::vector<float> A((size_t)m * m), B((size_t)m * m, 0.f);
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) { A[i * m + j] = at<float>(i, j); B[i * m + j] = (i == j) ? 1.f : 0.f; }
        const float eps = 1.1920929e-07f * 10;
        bool ok = true;
        for (int i = 0; i < m && ok; ++i) {
            int k = i;
            for (int j = i + 1; j < m; ++j)
                if (std::abs(A[j * m + i]) > std::abs(A[k * m + i])) k = j;
            if (std::abs(A[k * m + i]) < eps) { ok = false; break; }
            if (k != i) {
                for (int j = i; j < m; ++j) std::swap(A[i * m + j], A[k * m + j]);
                for (int j = 0; j < m; ++j) std::swap(B[i * m + j], B[k * m + j]);
            }
            const float d = -1 / A[i * m + i];
            for (int j = i + 1; j < m; ++j) {
                const float alpha = A[j * m + i] * d;
                for (int c = i + 1; c < m; ++c) A[j * m + c] += alpha * A[i * m + c];
                for (int c = 0; c < m; ++c) B[j * m + c] += alpha * B[i * m + c];
            }
            A[i * m + i] = -d;
        }
        Mat out(m, m, flags);
        if (!ok) return out;
        for (int i = m - 1; i >= 0; --i)
            for (int j = 0; j < m; ++j) {
                float s = B[i * m + j];
                for (int c = i + 1; c < m; ++c) s -= A[i * m + c] * B[c * m + j];
                B[i * m + j] = s * A[i * m + i];
            }
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) out.at<float>(i, j) = B[i * m + j];
        return out;
    }

    Mat& operator+=(const Mat& o) {
        assert(depth() == CV_32F && o.depth() == CV_32F && rows == o.rows && cols == o.cols);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) at<float>(r, c) += o.at<float>(r, c);
        return *this;
    }
    Mat& operator-=(const Mat& o) {
        assert(depth() == CV_32F && o.depth() == CV_32F && rows == o.rows && cols == o.cols);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) at<float>(r, c) -= o.at<float>(r, c);
        return *this;
    }
};

// cv::gemm for CV_32F (GEMMSingleMul<float,double>): double accumulator over k, rounded to float once
inline Mat operator*(const Mat& a, const Mat& b) {
    assert(a.depth() == CV_32F && b.depth() == CV_32F && a.cols == b.rows);
    Mat m(a.rows, b.cols, a.flags);
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < b.cols; ++j) {
            double s = 0;
            for (int k = 0; k < a.cols; ++k) s += (double)a.at<float>(i, k) * (double)b.at<float>(k, j);
            m.at<float>(i, j) = (float)s;
        }
    return m;
}
inline Mat operator*(const Mat& a, double s) { return a.mul(s); }
inline Mat operator*(double s, const Mat& a) { return a.mul(s); }
// Mat / scalar is NOT a division in OpenCV: matop.cpp builds MatOp_AddEx with alpha = 1./s, and the assignment runs
// convertTo(dst, type, alpha), whose CV_32F -> CV_32F kernel multiplies by (float)alpha
inline Mat operator/(const Mat& a, double s) { return a.mul(1. / s); }
inline Mat operator+(const Mat& a, const Mat& b) { Mat m = a.clone(); m += b; return m; }
inline Mat operator-(const Mat& a, const Mat& b) { Mat m = a.clone(); m -= b; return m; }
inline Mat operator-(const Mat& a) { return a.mul(-1.0); }
// compare: 255 where true
inline Mat operator>(const Mat& a, double v) {
    Mat m(a.rows, a.cols, CV_8UC1);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) m.at<uchar>(r, c) = (a.getd(r, c) > v) ? 255 : 0;
    return m;
}
inline Mat operator!=(const Mat& a, double v) {
    Mat m(a.rows, a.cols, CV_8UC1);
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) m.at<uchar>(r, c) = (a.getd(r, c) != v) ? 255 : 0;
    return m;
}
inline void sqrt(const Mat& src, Mat& dst) {
    assert(src.depth() == CV_32F);
    Mat out(src.rows, src.cols, src.flags);
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) out.at<float>(r, c) = std::sqrt(src.at<float>(r, c));
    dst = out;
}
inline std::ostream& operator<<(std::ostream& os, const Mat& m) {
    os << "[";
    for (int r = 0; r < m.rows; ++r) {
        for (int c = 0; c < m.cols; ++c) os << m.getd(r, c) << (c + 1 < m.cols ? ", " : "");
        os << (r + 1 < m.rows ? ";\n " : "");
    }
    return os << "]";
}

template <typename T> struct DepthOf;
template <> struct DepthOf<float> { enum { value = CV_32F }; };
template <> struct DepthOf<double> { enum { value = CV_64F }; };
template <> struct DepthOf<uchar> { enum { value = CV_8U }; };

template <typename T> class MatCommaInitializer_ {
public:
    Mat m;
    size_t idx;
    explicit MatCommaInitializer_(const Mat& mm) : m(mm), idx(0) {}
    template <typename V> MatCommaInitializer_& operator,(V v) {
        m.ptr<T>(0)[idx++] = (T)v;
        return *this;
    }
    operator Mat() const { return m; }
};
template <typename T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, CV_MAKETYPE(DepthOf<T>::value, 1)) {}
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
};
template <typename T, typename V> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v) {
    MatCommaInitializer_<T> ci(m);
    return (ci, v);
}

inline int countNonZero(const Mat& m) {
    int n = 0;
    for (int r = 0; r < m.rows; ++r)
        for (int c = 0; c < m.cols; ++c) n += (m.getd(r, c) != 0);
    return n;
}
inline void multiply(const Mat& a, const Mat& b, Mat& dst, double scale = 1) { dst = a.mul(b, scale); }
inline void add(const Mat& a, const Mat& b, Mat& dst) { dst = a + b; }

// cv::pyrDown on CV_8UC1: separable [1 4 6 4 1]/16 in both directions, BORDER_REFLECT_101, dst = ((w+1)/2, (h+1)/2),
// (sum + 128) >> 8
inline void pyrDown(const Mat& src, Mat& dst, const Size& = Size()) {
    assert(src.type() == CV_8UC1);
    const int w = src.cols, h = src.rows, dw = (w + 1) / 2, dh = (h + 1) / 2;
    Mat out(dh, dw, CV_8UC1);
    auto refl = [](int p, int n) { if (n == 1) return 0; while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p; return p; };
    static const int kw[5] = {1, 4, 6, 4, 1};
    std::vector<int> rowbuf((size_t)h * dw);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < dw; ++x) {
            int s = 0;
            for (int k = -2; k <= 2; ++k) s += kw[k + 2] * src.at<uchar>(y, refl(2 * x + k, w));
            rowbuf[(size_t)y * dw + x] = s;
        }
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            int s = 0;
            for (int k = -2; k <= 2; ++k) s += kw[k + 2] * rowbuf[(size_t)refl(2 * y + k, h) * dw + x];
            out.at<uchar>(y, x) = (uchar)((s + 128) >> 8);
        }
    dst = out;
}

// cv::resize(INTER_LINEAR) on CV_8U, 1 or 3 channels, by scale factors: pixel-centre mapping sx = (dx + 0.5)/fx - 0.5.
// (Only the frame constructor calls it, on the 4x camera image; the driver feeds 4x pixel-replicated images, for which
// every bilinear resampling returns the original pixels exactly.)
inline void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int = INTER_LINEAR) {
    assert(src.depth() == CV_8U);
    int dw = dsize.width, dh = dsize.height;
    if (dw == 0 || dh == 0) { dw = (int)std::lrint(src.cols * fx); dh = (int)std::lrint(src.rows * fy); }
    else { fx = (double)dw / src.cols; fy = (double)dh / src.rows; }
    const int cn = src.channels();
    Mat out(dh, dw, src.type());
    for (int y = 0; y < dh; ++y) {
        double sy = (y + 0.5) / fy - 0.5;
        int y0 = (int)std::floor(sy); double wy = sy - y0;
        int y1 = std::min(std::max(y0 + 1, 0), src.rows - 1); y0 = std::min(std::max(y0, 0), src.rows - 1);
        for (int x = 0; x < dw; ++x) {
            double sx = (x + 0.5) / fx - 0.5;
            int x0 = (int)std::floor(sx); double wx = sx - x0;
            int x1 = std::min(std::max(x0 + 1, 0), src.cols - 1); x0 = std::min(std::max(x0, 0), src.cols - 1);
            for (int c = 0; c < cn; ++c) {
                const double v = (1 - wy) * ((1 - wx) * src.ptr<uchar>(y0)[x0 * cn + c] + wx * src.ptr<uchar>(y0)[x1 * cn + c]) +
                                 wy * ((1 - wx) * src.ptr<uchar>(y1)[x0 * cn + c] + wx * src.ptr<uchar>(y1)[x1 * cn + c]);
                out.ptr<uchar>(y)[x * cn + c] = (uchar)std::min(255L, std::max(0L, std::lrint(v)));
            }
        }
    }
    dst = out;
}

// cv::cvtColor(BGR2GRAY) on CV_8UC3: (B*1868 + G*9617 + R*4899 + 8192) >> 14
inline void cvtColor(const Mat& src, Mat& dst, int code) {
    assert(code == CV_BGR2GRAY && src.type() == CV_8UC3);
    (void)code;
    Mat out(src.rows, src.cols, CV_8UC1);
    for (int y = 0; y < src.rows; ++y)
        for (int x = 0; x < src.cols; ++x) {
            const uchar* p = src.ptr<uchar>(y) + 3 * x;
            out.at<uchar>(y, x) = (uchar)((p[0] * 1868 + p[1] * 9617 + p[2] * 4899 + 8192) >> 14);
        }
    dst = out;
}

// Lens undistortion is outside the tracking path (the driver runs with FLAG_DO_UNDISTORTION = false, as SURVEY 8d specifies).
inline Mat getOptimalNewCameraMatrix(const Mat&, const Mat&, Size, double) {
    std::fprintf(stderr, "shim: getOptimalNewCameraMatrix is not available\n"); std::abort();
}
inline void undistort(const Mat&, Mat&, const Mat&, const Mat&, const Mat&) {
    std::fprintf(stderr, "shim: undistort is not available\n"); std::abort();
}

typedef Mat MatND;
inline int cvRound(double v) { return (int)std::lrint(v); }
struct Point { int x, y; Point() : x(0), y(0) {} Point(int xx, int yy) : x(xx), y(yy) {} };
enum { COLORMAP_JET = 2 };
// drawing / colour maps: display only, never on a compared path
inline void applyColorMap(const Mat& src, Mat& dst, int) { dst = Mat(src.rows, src.cols, CV_8UC3); }
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
inline void rectangle(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
inline void circle(Mat&, Point, int, const Scalar&, int = 1, int = 8, int = 0) {}
inline void putText(Mat&, const String&, Point, int, double, Scalar, int = 1, int = 8, bool = false) {}

// cv::calcHist as the reference calls it (src/GlobalOptimize.cpp:68): one CV_8UC1 image, no mask, `hsize` uniform bins over
// [ranges[0][0], ranges[0][1]), float counts in an hsize x 1 Mat
inline void calcHist(const Mat* images, int nimages, const int* channels, const Mat& mask, Mat& hist, int dims, const int* histSize,
                     const float** ranges, bool uniform = true, bool accumulate = false) {
    assert(nimages == 1 && dims == 1 && uniform && !accumulate && images[0].type() == CV_8UC1 && mask.empty());
    (void)channels; (void)nimages; (void)dims; (void)uniform; (void)accumulate;
    const int n = histSize[0];
    const double lo = ranges[0][0], hi = ranges[0][1];
    Mat h(n, 1, CV_32FC1);
    for (int y = 0; y < images[0].rows; ++y)
        for (int x = 0; x < images[0].cols; ++x) {
            const double v = images[0].at<uchar>(y, x);
            if (v < lo || v >= hi) continue;
            const int b = std::min(n - 1, (int)std::floor((v - lo) * n / (hi - lo)));
            h.at<float>(b, 0) += 1.f;
        }
    hist = h;
}
// cv::compareHist(CV_COMP_KL_DIV) on CV_32F histograms: sum p log(p / q) in double, skipping bins where p or q is (nearly) zero
inline double compareHist(const Mat& H1, const Mat& H2, int method) {
    assert(method == CV_COMP_KL_DIV && H1.depth() == CV_32F && H2.depth() == CV_32F && H1.total() == H2.total());
    (void)method;
    double result = 0;
    for (int r = 0; r < H1.rows; ++r)
        for (int c = 0; c < H1.cols; ++c) {
            const double p = H1.at<float>(r, c), q = H2.at<float>(r, c);
            if (std::fabs(p) <= 2.2204460492503131e-16) continue;
            const double qq = std::fabs(q) <= 2.2204460492503131e-16 ? 1e-10 : q;
            result += p * std::log(p / qq);
        }
    return result;
}

// The driver hands images to frame::frame(VideoCapture) through this stand-in.
class VideoCapture {
public:
    const Mat* next;
    VideoCapture() : next(nullptr) {}
    explicit VideoCapture(const Mat* m) : next(m) {}
    VideoCapture& operator>>(Mat& m) { m = next ? next->clone() : Mat(); return *this; }
    bool isOpened() const { return next != nullptr; }
};

inline int waitKey(int = 0) { return -1; }
inline void imshow(const String&, const Mat&) {}
inline void namedWindow(const String&, int = 0) {}
inline void moveWindow(const String&, int, int) {}
inline void destroyWindow(const String&) {}
inline bool imwrite(const String&, const Mat&) { return false; }

}  // namespace cv
