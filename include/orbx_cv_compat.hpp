// include/orbx_cv_compat.hpp
//
// Minimal stand-in for the handful of OpenCV *types* that appear in the ORBextractor interface
// (reference: /root/reference/inc/ORBextractor.h:24-25 pulls in <opencv/cv.h>, <opencv2/opencv.hpp>).
// It exists only so that include/ORBextractor.h compiles on machines without OpenCV headers (this
// build image has none).  When real OpenCV headers are on the include path they are used instead
// and nothing below is seen.  No image-processing *functions* live here: the product does all pixel
// work on the GPU behind include/orbx.h.
#ifndef ORBX_CV_COMPAT_HPP_
#define ORBX_CV_COMPAT_HPP_

#if !defined(ORBX_FORCE_CV_COMPAT) && defined(__has_include)
#if __has_include(<opencv2/core.hpp>)
#define ORBX_HAVE_OPENCV 1
#endif
#endif

#ifdef ORBX_HAVE_OPENCV
#include <opencv2/core.hpp>
#else

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#ifndef CV_PI
#define CV_PI 3.1415926535897932384626433832795
#endif
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAT_DEPTH(t) ((t) & 7)
#define CV_MAT_CN(t) ((((t) >> CV_CN_SHIFT) & 511) + 1)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {

template <typename T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
    template <typename U>
    explicit Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
    Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
    bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

template <typename T>
struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
    bool empty() const { return width <= 0 || height <= 0; }
    bool operator==(const Size_& o) const { return width == o.width && height == o.height; }
};
typedef Size_<int> Size;

template <typename T>
struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T _x, T _y, T w, T h) : x(_x), y(_y), width(w), height(h) {}
};
typedef Rect_<int> Rect;

// Same 28-byte layout as cv::KeyPoint: {pt.x, pt.y, size, angle, response, octave, class_id}.
class KeyPoint {
public:
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(Point2f _pt, float _size, float _angle = -1, float _response = 0, int _octave = 0,
             int _class_id = -1)
        : pt(_pt), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0,
             int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
    Point2f pt;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

class _InputArray;
class _OutputArray;
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

// Reference-counted 2-D matrix header with ROI support (the subset of cv::Mat the interface uses).
class Mat {
public:
    enum { AUTO_STEP = 0 };
    int rows, cols;
    uchar* data;
    size_t step;  // bytes per row

    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(CV_8UC1) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), step(0), type_(type) { create(r, c, type); }
    Mat(Size sz, int type) : rows(0), cols(0), data(nullptr), step(0), type_(type) { create(sz.height, sz.width, type); }
    // Non-owning header over user memory.
    Mat(int r, int c, int type, void* ext, size_t ext_step = AUTO_STEP)
        : rows(r), cols(c), data((uchar*)ext), step(ext_step ? ext_step : (size_t)c * esz(type)), type_(type) {}
    Mat(const Mat& m, const Rect& roi)
        : rows(roi.height), cols(roi.width), data(m.data + (size_t)roi.y * m.step + (size_t)roi.x * m.elemSize()),
          step(m.step), type_(m.type_), buf_(m.buf_) {}

    void create(int r, int c, int type) {
        if (data && r == rows && c == cols && type == type_) return;
        rows = r; cols = c; type_ = type;
        step = (size_t)c * esz(type);
        size_t total = step * (size_t)r;
        buf_ = std::shared_ptr<uchar>(new uchar[total ? total : 1], std::default_delete<uchar[]>());
        data = buf_.get();
    }
    void create(Size sz, int type) { create(sz.height, sz.width, type); }
    void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }

    Mat operator()(const Rect& roi) const { return Mat(*this, roi); }
    Mat rowRange(int r0, int r1) const { return Mat(*this, Rect(0, r0, cols, r1 - r0)); }
    Mat colRange(int c0, int c1) const { return Mat(*this, Rect(c0, 0, c1 - c0, rows)); }
    Mat row(int r) const { return Mat(*this, Rect(0, r, cols, 1)); }

    Mat clone() const { Mat m; copyToMat(m); return m; }
    inline void copyTo(OutputArray dst) const;
    void copyToMat(Mat& m) const {
        m.create(rows, cols, type_);
        const size_t rowBytes = (size_t)cols * elemSize();
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, rowBytes);
    }
    static Mat zeros(int r, int c, int type) {
        Mat m(r, c, type);
        if (m.data) std::memset(m.data, 0, m.step * (size_t)r);
        return m;
    }

    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize() const { return esz(type_); }
    size_t elemSize1() const { return esz(CV_MAT_DEPTH(type_)); }
    size_t step1() const { return step / elemSize1(); }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }

    // cv::Mat::reshape(cn): same data, `cn` channels per element (rows unchanged; the matrix must be continuous).
    Mat reshape(int cn) const {
        Mat m(*this);
        const int total_per_row = cols * CV_MAT_CN(type_);
        m.type_ = CV_MAKETYPE(CV_MAT_DEPTH(type_), cn);
        m.cols = total_per_row / cn;
        return m;
    }

    // Single-index access for row / column vectors (cv::Mat::at<T>(int i0)).
    template <typename T> T& at(int i) { return rows == 1 ? ((T*)data)[i] : *(T*)(data + (size_t)i * step); }
    template <typename T> const T& at(int i) const { return rows == 1 ? ((const T*)data)[i] : *(const T*)(data + (size_t)i * step); }
    template <typename T> T& at(int r, int c) { return ((T*)(data + (size_t)r * step))[c]; }
    template <typename T> const T& at(int r, int c) const { return ((const T*)(data + (size_t)r * step))[c]; }
    uchar* ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }

private:
    static size_t esz(int type) {
        static const size_t d[8] = {1, 1, 2, 2, 4, 4, 8, 2};
        return d[CV_MAT_DEPTH(type)] * (size_t)CV_MAT_CN(type);
    }
    int type_;
    std::shared_ptr<uchar> buf_;
};

// Thin proxies: only Mat-backed arrays are supported (all the reference passes on this path).
class _InputArray {
public:
    _InputArray() : m_(nullptr) {}
    _InputArray(const Mat& m) : m_(&m) {}
    bool empty() const { return !m_ || m_->empty(); }
    Mat getMat() const { return m_ ? *m_ : Mat(); }
protected:
    const Mat* m_;
};

class _OutputArray : public _InputArray {
public:
    _OutputArray() : out_(nullptr) {}
    _OutputArray(Mat& m) : _InputArray(m), out_(&m) {}
    // A temporary header (e.g. Mat::row(i)) used as a fixed-size destination.
    _OutputArray(const Mat& m) : _InputArray(m), out_(const_cast<Mat*>(&m)) {}
    void create(int rows, int cols, int type) const { if (out_) out_->create(rows, cols, type); }
    void create(Size sz, int type) const { create(sz.height, sz.width, type); }
    void release() const { if (out_) out_->release(); }
    Mat getMat() const { return out_ ? *out_ : Mat(); }
    Mat& getMatRef() const { return *out_; }
    bool needed() const { return out_ != nullptr; }
private:
    Mat* out_;
};

inline void Mat::copyTo(OutputArray dst) const { copyToMat(dst.getMatRef()); }

inline InputArray noArray() { static _OutputArray none; return none; }

}  // namespace cv

#endif  // ORBX_HAVE_OPENCV
#endif  // ORBX_CV_COMPAT_HPP_
