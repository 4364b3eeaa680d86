// tests/cpp/opencv_stub/opencv2/core.hpp -- TEST INFRASTRUCTURE.  Declarations only (no definitions), spelled like the
// public OpenCV 3 / 4 core API (modules/core/include/opencv2/core/{types,mat,cvdef}.hpp): just the part of cv:: that
// include/*.h, extractorb_b200/csrc/ORBextractor.cpp and tests/cpp/dropin_main.cpp touch.  This image has no OpenCV C++
// headers, so the `__has_include(<opencv2/core.hpp>)` branch of include/orbx_cv_compat.hpp would otherwise never be
// parsed by any test; tests/test_abi.py compiles the product sources against this tree with -fsyntax-only.
#ifndef OPENCV_CORE_HPP_STUB
#define OPENCV_CORE_HPP_STUB

#include <cstddef>
#include <vector>

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

namespace cv {

template <typename _Tp> class Point_ {
public:
    Point_();
    Point_(_Tp _x, _Tp _y);
    _Tp x, y;
};
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
typedef Point2i Point;

template <typename _Tp> class Size_ {
public:
    Size_();
    Size_(_Tp _width, _Tp _height);
    _Tp width, height;
};
typedef Size_<int> Size2i;
typedef Size2i Size;

template <typename _Tp> class Rect_ {
public:
    Rect_();
    Rect_(_Tp _x, _Tp _y, _Tp _width, _Tp _height);
    _Tp x, y, width, height;
};
typedef Rect_<int> Rect2i;
typedef Rect2i Rect;

class KeyPoint {
public:
    KeyPoint();
    KeyPoint(Point2f _pt, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1);
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1);
    Point2f pt;
    float size;
    float angle;
    float response;
    int octave;
    int class_id;
};

struct MatStep {
    MatStep();
    explicit MatStep(size_t s);
    operator size_t() const;
    size_t* p;
    size_t buf[2];
};

class MatExpr;

class Mat {
public:
    Mat();
    Mat(int rows, int cols, int type);
    Mat(Size size, int type);
    Mat(int rows, int cols, int type, void* data, size_t step = 0);
    Mat(const Mat& m);
    ~Mat();
    Mat& operator=(const Mat& m);
    Mat& operator=(const MatExpr& e);
    Mat(const MatExpr& e);
    Mat clone() const;
    void create(int rows, int cols, int type);
    void release();
    Mat operator()(const Rect& roi) const;
    bool empty() const;
    int type() const;
    uchar* ptr(int i0 = 0);
    const uchar* ptr(int i0 = 0) const;
    template <typename _Tp> _Tp& at(int row, int col);
    template <typename _Tp> const _Tp& at(int row, int col) const;
    static MatExpr zeros(int rows, int cols, int type);
    int flags;
    int dims;
    int rows, cols;
    uchar* data;
    MatStep step;
};

class MatExpr {
public:
    operator Mat() const;
};

class _InputArray {
public:
    _InputArray();
    _InputArray(const Mat& m);
    _InputArray(const MatExpr& expr);
    Mat getMat(int idx = -1) const;
    bool empty() const;
};

class _OutputArray : public _InputArray {
public:
    _OutputArray();
    _OutputArray(Mat& m);
    void create(int rows, int cols, int type, int i = -1, bool allowTransposed = false, int fixedDepthMask = 0) const;
    void release() const;
};

typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

}  // namespace cv

#endif  // OPENCV_CORE_HPP_STUB
