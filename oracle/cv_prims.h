/*
 * oracle/cv_prims.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatements (plain C) of the five OpenCV primitives the reference's hot path calls.
 * OpenCV itself is a third-party dependency that is absent from /root/reference
 * (CMakeLists.txt:23 `find_package(OpenCV 3)`, minor version unpinned), so the published
 * fixed-point algorithms are restated here and pinned against python cv2 4.13.0 by
 * tests/test_oracle_prims.py (live, when cv2 is importable) and tests/golden/prims_*.npz.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * use anything in oracle/.  The product (extractorb_b200/, include/) never links or calls it.
 *
 * Reference call sites (all in /root/reference/src/orb_extractor/ORBextractor.cc):
 *   resize           :1183      copyMakeBorder :1193, :1213     FAST :818, :837
 *   GaussianBlur     :1127      fastAtan2      :101             cvRound :79,:113,:117,:447,:1171
 */
#ifndef ORACLE_CV_PRIMS_H_
#define ORACLE_CV_PRIMS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cvRound(float/double): round half to even (SSE cvtsd2si semantics). */
int ocv_round_f(float v);
int ocv_round_d(double v);

/* cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR) for CV_8UC1. */
void ocv_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep,
                          uint8_t* dst, int dw, int dh, size_t dstep);

/* cv::copyMakeBorder(..., BORDER_REFLECT_101 [| BORDER_ISOLATED]) for CV_8UC1.
 * dst must hold (h+top+bottom) rows of (w+left+right) bytes; src may alias the interior of dst. */
void ocv_copy_make_border_reflect101_u8(const uint8_t* src, int w, int h, size_t sstep,
                                        uint8_t* dst, size_t dstep,
                                        int top, int bottom, int left, int right);

/* cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16.
 * Writes up to cap (x, y, score) triples in emission order (row-major); returns the number found
 * (which may exceed cap; only the first cap are stored). */
int ocv_fast9_16_nms(const uint8_t* img, int w, int h, size_t step, int threshold,
                     int* xs, int* ys, int* scores, int cap);

/* FAST corner measure used by the restatement: best(p) = max over the 16 arcs of 9 contiguous ring
 * pixels of min(d) and min(-d).  corner <=> best > threshold; score = best - 1. */
int ocv_fast9_16_best(const uint8_t* p, size_t step);

/* cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 (src != dst). */
void ocv_gaussian_blur_7x7_s2_u8(const uint8_t* src, int w, int h, size_t sstep,
                                 uint8_t* dst, size_t dstep);

/* cv::fastAtan2(y, x) scalar path, degrees in [0, 360]. */
float ocv_fast_atan2(float y, float x);

/* cv::undistortPoints(src, dst, K, distCoeffs, noArray(), P) for CV_32FC2 points with K = P =
 * [fx 0 cx; 0 fy cy; 0 0 1] given as floats and n_dist in {4, 5} coefficients (k1 k2 p1 p2 [k3]): the default
 * 5 fixed-point iterations in double precision (no FMA), result rounded to float.  Call sites:
 * Frame::UndistortKeyPoints src/Frame.cc:767, Frame::ComputeImageBounds :797.  xy_in may alias xy_out. */
void ocv_undistort_points_f32(const float* xy_in, int n, float fx, float fy, float cx, float cy, const float* dist,
                              int n_dist, float* xy_out);

/* cv::createCLAHE(clip_limit, Size(tiles_x, tiles_y))->apply(src, dst) for CV_8UC1: per-tile clipped histograms ->
 * LUTs, bilinear interpolation between the four surrounding tile LUTs in float (no FMA), round half to even.  Images
 * whose sides are not multiples of the tile grid are extended with BORDER_REFLECT_101 for the LUTs only.  Call sites:
 * src/orb_extractor/main_orb_extractor.cpp:19-22, src/clahe/main_clahe.cpp:7-11, main_show_clahe_keypoint.cpp:19-22.
 * Returns 0, or -1 for bad arguments. */
int ocv_clahe_u8(const uint8_t* src, int w, int h, size_t sstep, double clip_limit, int tiles_x, int tiles_y,
                 uint8_t* dst, size_t dstep);

#ifdef __cplusplus
}
#endif
#endif
