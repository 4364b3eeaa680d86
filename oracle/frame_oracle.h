/*
 * oracle/frame_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * CPU restatement of the per-frame post-processing that every Frame constructor runs right after
 * ExtractORB, and of the matcher that consumes it during monocular initialisation:
 *   Frame::UndistortKeyPoints        /root/reference/src/Frame.cc:748-782   (cv::undistortPoints -> oracle/cv_prims.c)
 *   Frame::ComputeImageBounds        src/Frame.cc:784-812
 *   Frame::AssignFeaturesToGrid      src/Frame.cc:383-417   + PosInGrid :726-736, FRAME_GRID 64 x 48 (inc/Frame.h:39-40)
 *   Frame::GetFeaturesInArea         src/Frame.cc:655-724
 *   ORBmatcher::SearchForInitialization  src/ORBmatcher.cc:705-814  + ComputeThreeMaxima :2303-2344,
 *                                        DescriptorDistance :2349-2365, TH_LOW 50, HISTO_LENGTH 30 (:36-38)
 * Pinned against those functions compiled unmodified from /root/reference (oracle/_ref/ref_frame, recipe in
 * oracle/Makefile) through tests/golden/frame_*.npz, and against cv2 4.13.0 for undistortPoints.
 */
#ifndef ORACLE_FRAME_ORACLE_H_
#define ORACLE_FRAME_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#include "orb_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif

#define ORB_ORACLE_GRID_COLS 64
#define ORB_ORACLE_GRID_ROWS 48

typedef struct OrbOracleCalib {
    float fx, fy, cx, cy;   /* mK */
    float dist[5];          /* mDistCoef: k1 k2 p1 p2 [k3] */
    int32_t n_dist;         /* 4 or 5 */
    float min_x, max_x, min_y, max_y; /* mnMinX, mnMaxX, mnMinY, mnMaxY */
} OrbOracleCalib;

/* Frame::ComputeImageBounds: fills min_x..max_y of `c` for a width x height image. */
void orb_oracle_image_bounds(OrbOracleCalib* c, int width, int height);
/* Frame::UndistortKeyPoints: keys -> keys_un (copies when dist[0] == 0). */
void orb_oracle_undistort_keypoints(const OrbOracleCalib* c, const OrbOracleKeyPoint* keys, int n, OrbOracleKeyPoint* keys_un);
/* Frame::AssignFeaturesToGrid (mono, Nleft == -1): mGrid[ix][iy] = cell_items[cell_start[ix*48+iy] .. cell_start[ix*48+iy+1]),
 * ascending keypoint index inside a cell.  cell_start has 64*48+1 entries.  Returns the number of keypoints placed. */
int orb_oracle_assign_grid(const OrbOracleCalib* c, const OrbOracleKeyPoint* keys_un, int n, int32_t* cell_start, int32_t* cell_items);
/* Frame::GetFeaturesInArea (bRight = false).  Returns the count; out receives up to cap indices in the reference's order. */
int orb_oracle_features_in_area(const OrbOracleCalib* c, const OrbOracleKeyPoint* keys_un, const int32_t* cell_start,
                                const int32_t* cell_items, float x, float y, float r, int min_level, int max_level,
                                int32_t* out, int cap);
/* ORBmatcher::SearchForInitialization.  prev_matched: n1 (x, y) pairs, updated in place like vbPrevMatched;
 * matches12: n1 ints (-1 = none).  Returns nmatches. */
int orb_oracle_search_for_initialization(const OrbOracleCalib* c, const OrbOracleKeyPoint* keys_un1, const uint8_t* desc1, int n1,
                                         const OrbOracleKeyPoint* keys_un2, const uint8_t* desc2, int n2,
                                         const int32_t* cell_start2, const int32_t* cell_items2, float* prev_matched,
                                         int window_size, float nn_ratio, int check_orientation, int32_t* matches12);
#ifdef __cplusplus
}
#endif
#endif
