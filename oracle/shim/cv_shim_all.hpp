// oracle/shim/cv_shim_all.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Everything /root/reference/src/orb_extractor/ORBextractor.cc needs from OpenCV so that it compiles
// VERBATIM here (the image has no OpenCV C++): the cv types from include/orbx_cv_compat.hpp plus the
// five primitives (restated in oracle/cv_prims.c and pinned against cv2 4.13.0), cvRound/cvFloor/
// cvCeil and a few constants.  API surface per SURVEY.md section 8(c).
#ifndef ORACLE_CV_SHIM_ALL_HPP_
#define ORACLE_CV_SHIM_ALL_HPP_

#define ORBX_FORCE_CV_COMPAT 1
#include "../../include/orbx_cv_compat.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../cv_prims.h"

static inline int cvRound(double v) { return ocv_round_d(v); }
static inline int cvRound(float v) { return ocv_round_f(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3,
       BORDER_REFLECT_101 = 4, BORDER_REFLECT101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };

struct Scalar { double v[4]; Scalar() { v[0] = v[1] = v[2] = v[3] = 0; } };

inline float fastAtan2(float y, float x) { return ocv_fast_atan2(y, x); }

inline void resize(InputArray _src, OutputArray _dst, Size dsize, double fx = 0, double fy = 0,
                   int interpolation = INTER_LINEAR) {
    (void)fx; (void)fy;
    assert(interpolation == INTER_LINEAR);
    Mat src = _src.getMat();
    assert(src.type() == CV_8UC1 && !dsize.empty());
    _dst.create(dsize, src.type());
    Mat dst = _dst.getMat();
    ocv_resize_linear_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.cols, dst.rows, dst.step);
}

// The input is always treated as isolated (pixels outside a sub-matrix are never read).
inline void copyMakeBorder(InputArray _src, OutputArray _dst, int top, int bottom, int left, int right,
                           int borderType, const Scalar& = Scalar()) {
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
    Mat src = _src.getMat();
    assert(src.type() == CV_8UC1);
    _dst.create(src.rows + top + bottom, src.cols + left + right, src.type());
    Mat dst = _dst.getMat();
    ocv_copy_make_border_reflect101_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step,
                                       top, bottom, left, right);
}

inline void FAST(InputArray _img, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true) {
    assert(nonmaxSuppression);
    Mat img = _img.getMat();
    keypoints.clear();
    const int cap = ((img.cols + 1) / 2) * ((img.rows + 1) / 2) + 16;
    std::vector<int> xs(cap), ys(cap), sc(cap);
    int n = ocv_fast9_16_nms(img.data, img.cols, img.rows, img.step, threshold, xs.data(), ys.data(), sc.data(), cap);
    assert(n <= cap);
    keypoints.reserve(n);
    for (int i = 0; i < n; ++i)
        keypoints.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
}

inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY = 0,
                         int borderType = BORDER_DEFAULT) {
    assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    Mat src = _src.getMat().clone();
    _dst.create(src.rows, src.cols, src.type());
    Mat dst = _dst.getMat();
    ocv_gaussian_blur_7x7_s2_u8(src.data, src.cols, src.rows, src.step, dst.data, dst.step);
}

// Only referenced by dead code (ComputeKeyPointsOld, ORBextractor.cc:890-1067).
struct KeyPointsFilter {
    static void retainBest(std::vector<KeyPoint>&, int) { std::abort(); }
};

}  // namespace cv

#endif
