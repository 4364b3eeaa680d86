// oracle/shim/cv_shim_stereo.hpp -- TEST INFRASTRUCTURE.  The extra OpenCV surface that
// Frame::ComputeStereoMatches (reference src/Frame.cc:813-990) needs on top of cv_shim_all.hpp:
// Mat::convertTo(CV_16S), Mat - scalar on CV_16S, cv::norm(a, b, NORM_L1).
#ifndef ORACLE_CV_SHIM_STEREO_HPP_
#define ORACLE_CV_SHIM_STEREO_HPP_

#include "cv_shim_all.hpp"

#include <climits>
#include <cstdint>

namespace cv {

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

// cv::Mat::convertTo for 8U -> 16S (no scaling); `dst` may alias `src` (the reference converts in place).
inline void convertTo16S(const Mat& src, Mat& dst) {
    assert(src.type() == CV_8UC1);
    Mat tmp(src.rows, src.cols, CV_MAKETYPE(CV_16S, 1));
    for (int r = 0; r < src.rows; ++r)
        for (int c = 0; c < src.cols; ++c) tmp.at<short>(r, c) = (short)src.at<uchar>(r, c);
    dst = tmp;
}

// Mat - scalar for CV_16S (saturating like cv::subtract).
inline Mat operator-(const Mat& a, int s) {
    assert(a.type() == CV_MAKETYPE(CV_16S, 1));
    Mat out(a.rows, a.cols, a.type());
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) {
            int v = (int)a.at<short>(r, c) - s;
            out.at<short>(r, c) = (short)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
        }
    return out;
}

inline double norm(const Mat& a, const Mat& b, int normType) {
    assert(normType == NORM_L1 && a.type() == CV_MAKETYPE(CV_16S, 1) && b.type() == a.type() && a.rows == b.rows && a.cols == b.cols);
    long long acc = 0;
    for (int r = 0; r < a.rows; ++r)
        for (int c = 0; c < a.cols; ++c) {
            int d = (int)a.at<short>(r, c) - (int)b.at<short>(r, c);
            acc += d < 0 ? -d : d;
        }
    return (double)acc;
}

}  // namespace cv

#endif
