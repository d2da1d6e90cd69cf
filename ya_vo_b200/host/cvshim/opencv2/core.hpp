// Minimal stand-in for the OpenCV core types the YA_VO front-end classes expose in their public
// signatures (cv::Mat, cv::Point, cv::Size, cv::Scalar, cv::Vec3b, cv::Range, cv::Mat_).
// Used ONLY when the real OpenCV headers are not installed (this image has no OpenCV C++); with
// OpenCV present, ya_vo_b200/host/include/yavo_cv.hpp includes <opencv2/core.hpp> instead.
// Not a general OpenCV replacement: just enough surface for Image / FastDetector / Brief and their tests.
#ifndef YAVO_CVSHIM_CORE_HPP
#define YAVO_CVSHIM_CORE_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <ostream>
#include <stdexcept>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

class Exception : public std::runtime_error {
   public:
    explicit Exception(const std::string &m) : std::runtime_error(m) {}
};

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    bool operator==(const Point_ &o) const { return x == o.x && y == o.y; }
    bool operator!=(const Point_ &o) const { return !(*this == o); }
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
template <typename T>
std::ostream &operator<<(std::ostream &os, const Point_<T> &p) {
    return os << "[" << p.x << ", " << p.y << "]";
}

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

struct TermCriteria {  // cv::TermCriteria (type: COUNT = 1, EPS = 2)
    enum Type { COUNT = 1, MAX_ITER = COUNT, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria() : type(0), maxCount(0), epsilon(0) {}
    TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    double operator[](int i) const { return val[i]; }
};

struct Vec3b {
    uchar val[3];
    Vec3b() { val[0] = val[1] = val[2] = 0; }
    Vec3b(uchar a, uchar b, uchar c) { val[0] = a; val[1] = b; val[2] = c; }
    uchar &operator[](int i) { return val[i]; }
    const uchar &operator[](int i) const { return val[i]; }
};

struct Range {
    int start, end;
    Range(int s, int e) : start(s), end(e) {}
};

template <typename T>
inline T saturate_cast(double v);
template <>
inline uchar saturate_cast<uchar>(double v) {
    const double r = std::nearbyint(v);  // round half to even, as cvRound
    return (uchar)(r < 0 ? 0 : (r > 255 ? 255 : r));
}
template <>
inline float saturate_cast<float>(double v) { return (float)v; }
template <>
inline double saturate_cast<double>(double v) { return v; }

class Mat {
   public:
    int rows, cols;
    uchar *data;
    size_t step;  // bytes per row

    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(CV_8UC1) {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, const Scalar &s) { create(r, c, type); setTo(s); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    Mat(Size sz, int type, const Scalar &s) { create(sz.height, sz.width, type); setTo(s); }
    // header over caller-owned pixels (no copy), as cv::Mat(rows, cols, type, void*, step)
    Mat(int r, int c, int type, void *ext, size_t stp = 0) : rows(r), cols(c), data((uchar *)ext), type_(type) {
        step = stp ? stp : (size_t)c * elemSize();
    }

    static Mat zeros(int r, int c, int type) { return Mat(r, c, type, Scalar(0)); }
    static Mat zeros(Size sz, int type) { return Mat(sz, type, Scalar(0)); }

    void create(int r, int c, int type) { allocate(r, c, type, true); }
    void allocate(int r, int c, int type, bool zero_pixels) {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        // 64 KiB of zeroed slack after the pixels: Image::getPixelVal-style unchecked linear reads slightly past
        // the end (the reference's BRIEF does this when row + 8 == H) see zeros instead of heap contents
        const size_t n = step * r;
        buf_ = std::shared_ptr<uchar>(new uchar[n + 65536], std::default_delete<uchar[]>());
        data = buf_.get();
        std::memset(zero_pixels ? data : data + n, 0, zero_pixels ? n + 65536 : 65536);
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize1() const { return depth() == CV_8U ? 1 : (depth() == CV_32F ? 4 : 8); }
    size_t elemSize() const { return elemSize1() * channels(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return (size_t)rows * cols; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    Size size() const { return Size(cols, rows); }

    template <typename T> T &at(int r, int c) { return *reinterpret_cast<T *>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T &at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T &at(Point p) const { return at<T>(p.y, p.x); }
    template <typename T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + (size_t)r * step); }

    void setTo(const Scalar &s) {
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++)
                for (int k = 0; k < channels(); k++) put(r, c, k, s.val[k < 4 ? k : 3]);
    }
    Mat clone() const {
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat &dst) const {
        if (dst.rows != rows || dst.cols != cols || dst.type_ != type_ || !dst.data) dst.create(rows, cols, type_);
        for (int r = 0; r < rows; r++) std::memcpy(dst.data + (size_t)r * dst.step, data + (size_t)r * step, (size_t)cols * elemSize());
    }
    void convertTo(Mat &dst, int type) const {
        Mat out(rows, cols, CV_MAKETYPE(type & 7, channels()));
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++)
                for (int k = 0; k < channels(); k++) out.put(r, c, k, get(r, c, k));
        dst = out;
    }
    Mat mul(const Mat &o) const {  // element-wise product
        if (type_ == CV_32FC1 && o.type_ == CV_32FC1) {  // the only case on the hot path: plain float rows, no zero fill
            Mat out;
            out.allocate(rows, cols, type_, false);
            for (int r = 0; r < rows; r++) {
                const float *a = ptr<float>(r), *b = o.ptr<float>(r);
                float *d = out.ptr<float>(r);
                for (int c = 0; c < cols; c++) d[c] = a[c] * b[c];
            }
            return out;
        }
        Mat out(rows, cols, type_);
        for (int r = 0; r < rows; r++)
            for (int c = 0; c < cols; c++)
                for (int k = 0; k < channels(); k++) out.put(r, c, k, mulElem(get(r, c, k), o.get(r, c, k)));
        return out;
    }
    Mat operator()(const Range &rr, const Range &cr) const {  // sub-matrix view sharing the buffer
        Mat m;
        m.rows = rr.end - rr.start; m.cols = cr.end - cr.start; m.type_ = type_; m.step = step; m.buf_ = buf_;
        m.data = data + (size_t)rr.start * step + (size_t)cr.start * elemSize();
        return m;
    }
    // element access through double, for the generic helpers above
    double get(int r, int c, int k = 0) const {
        const uchar *p = data + (size_t)r * step + ((size_t)c * channels() + k) * elemSize1();
        switch (depth()) {
            case CV_8U: return *p;
            case CV_32F: { float f; std::memcpy(&f, p, 4); return f; }
            default: { double d; std::memcpy(&d, p, 8); return d; }
        }
    }
    void put(int r, int c, int k, double v) {
        uchar *p = data + (size_t)r * step + ((size_t)c * channels() + k) * elemSize1();
        switch (depth()) {
            case CV_8U: *p = saturate_cast<uchar>(v); break;
            case CV_32F: { float f = (float)v; std::memcpy(p, &f, 4); break; }
            default: std::memcpy(p, &v, 8);
        }
    }

   protected:
    double mulElem(double a, double b) const { return depth() == CV_32F ? (double)((float)a * (float)b) : a * b; }
    int type_;
    std::shared_ptr<uchar> buf_;
};

// matrix product (float / double single channel), as cv::Mat operator*
inline Mat operator*(const Mat &a, const Mat &b) {
    if (a.cols != b.rows) throw Exception("Mat operator*: size mismatch");
    Mat out = Mat::zeros(a.rows, b.cols, a.type());
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            double s = 0;
            for (int k = 0; k < a.cols; k++) s += a.get(i, k) * b.get(k, j);
            out.put(i, j, 0, s);
        }
    return out;
}

template <typename T> struct DataDepth;
template <> struct DataDepth<uchar> { enum { value = CV_8U }; };
template <> struct DataDepth<float> { enum { value = CV_32F }; };
template <> struct DataDepth<double> { enum { value = CV_64F }; };

template <typename T>
class MatCommaInit_;

template <typename T>
class Mat_ : public Mat {
   public:
    Mat_() {}
    Mat_(int r, int c) : Mat(r, c, CV_MAKETYPE(DataDepth<T>::value, 1)) {}
    T &operator()(int r, int c) { return this->template at<T>(r, c); }
};

// (cv::Mat_<float>(3,3) << a, b, c, ...) comma initialiser
template <typename T>
class MatCommaInit_ {
   public:
    MatCommaInit_(const Mat_<T> &m, T first) : m_(m), i_(0) { push(first); }
    MatCommaInit_ &operator,(T v) { push(v); return *this; }
    operator Mat() const { return m_; }
    operator Mat_<T>() const { return m_; }

   private:
    void push(T v) {
        if (i_ < (int)m_.total()) m_.template at<T>(i_ / m_.cols, i_ % m_.cols) = v;
        i_++;
    }
    Mat_<T> m_;
    int i_;
};
template <typename T, typename V>
MatCommaInit_<T> operator<<(const Mat_<T> &m, V v) { return MatCommaInit_<T>(m, (T)v); }

enum { BORDER_CONSTANT = 0, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { COLOR_GRAY2RGB = 8, COLOR_GRAY2BGR = 8 };

}  // namespace cv
#endif
