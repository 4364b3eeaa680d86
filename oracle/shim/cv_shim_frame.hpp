// oracle/shim/cv_shim_frame.hpp -- TEST INFRASTRUCTURE.  The extra OpenCV surface that Frame::UndistortKeyPoints /
// ComputeImageBounds (reference src/Frame.cc:748-812) need on top of cv_shim_all.hpp: cv::undistortPoints on an
// N x 1 CV_32FC2 matrix with K == P and no rectification (restated in oracle/cv_prims.c, pinned against cv2 4.13.0).
#ifndef ORACLE_CV_SHIM_FRAME_HPP_
#define ORACLE_CV_SHIM_FRAME_HPP_

#include "cv_shim_all.hpp"

namespace cv {

inline void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& dist, const Mat& R, const Mat& P) {
    assert(src.type() == CV_MAKETYPE(CV_32F, 2) && src.cols == 1 && R.empty());
    assert(K.type() == CV_32FC1 && P.type() == CV_32FC1 && K.rows == 3 && P.rows == 3 && dist.type() == CV_32FC1);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) assert(K.at<float>(r, c) == P.at<float>(r, c));
    const int nd = dist.rows * dist.cols;
    float d[8] = {0};
    for (int i = 0; i < nd && i < 8; ++i) d[i] = dist.at<float>(i);
    const int n = src.rows;
    std::vector<float> in(2 * (size_t)n), out(2 * (size_t)n);
    for (int i = 0; i < n; ++i) { in[2 * i] = src.at<float>(i, 0); in[2 * i + 1] = src.at<float>(i, 1); }
    ocv_undistort_points_f32(in.data(), n, K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), d, nd, out.data());
    if (dst.data != src.data) dst.create(n, 1, CV_MAKETYPE(CV_32F, 2));
    for (int i = 0; i < n; ++i) { dst.at<float>(i, 0) = out[2 * i]; dst.at<float>(i, 1) = out[2 * i + 1]; }
}

}  // namespace cv

#endif
