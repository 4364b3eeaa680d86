// include/ORBframe.h -- GPU replacements for what every ORB-SLAM3 Frame constructor does with the extractor's output
// (reference src/Frame.cc:318-347) and for the matcher of monocular initialisation (src/ORBmatcher.cc:705-814):
//
//   Frame::ComputeImageBounds     src/Frame.cc:784-812   -> ORB_SLAM3::ComputeImageBounds
//   Frame::UndistortKeyPoints     src/Frame.cc:748-782   \_ ORB_SLAM3::UndistortAndAssignToGrid
//   Frame::AssignFeaturesToGrid   src/Frame.cc:383-417   /
//   ORBmatcher::SearchForInitialization (with Frame::GetFeaturesInArea :655-724, DescriptorDistance, ComputeThreeMaxima)
//                                                        -> ORB_SLAM3::SearchForInitialization
//
// Data members keep the reference's shapes: mvKeysUn is a std::vector<cv::KeyPoint>, mGrid is
// std::vector<std::size_t>[FRAME_GRID_COLS][FRAME_GRID_ROWS] (inc/Frame.h:195).  A Frame would call
//
//   ORB_SLAM3::ComputeImageBounds(*mpORBextractorLeft, mK, mDistCoef, imGray.cols, imGray.rows, calib);   // first frame only
//   ORB_SLAM3::UndistortAndAssignToGrid(*mpORBextractorLeft, calib, mvKeys, mvKeysUn, mGrid);
//
// and ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) becomes
//
//   ORB_SLAM3::SearchForInitialization(*F1.mpORBextractorLeft, calib, F1.mvKeysUn, F1.mDescriptors, F2.mvKeysUn, F2.mDescriptors,
//                                      F2.mGrid, vbPrevMatched, vnMatches12, windowSize, mfNNratio, mbCheckOrientation);
#ifndef ORBFRAME_H
#define ORBFRAME_H

#include <cstddef>
#include <vector>

#include "ORBextractor.h"
#include "orbx.h"

#ifndef FRAME_GRID_ROWS
#define FRAME_GRID_ROWS 48  // inc/Frame.h:39
#endif
#ifndef FRAME_GRID_COLS
#define FRAME_GRID_COLS 64  // inc/Frame.h:40
#endif

namespace ORB_SLAM3 {

typedef std::vector<std::size_t> FrameGridCells[FRAME_GRID_COLS][FRAME_GRID_ROWS];

// K: 3x3 CV_32F (mK), distCoef: 4x1 or 5x1 CV_32F (mDistCoef).  Fills calib (intrinsics + mnMinX .. mnMaxY).
// Returns false on error (see ext.LastError()).
bool ComputeImageBounds(ORBextractor& ext, const cv::Mat& K, const cv::Mat& distCoef, int cols, int rows, OrbxFrameCalib& calib);

// mvKeys -> mvKeysUn and mGrid (cleared first).  Returns the number of keypoints placed in the grid, -1 on error.
int UndistortAndAssignToGrid(ORBextractor& ext, const OrbxFrameCalib& calib, const std::vector<cv::KeyPoint>& mvKeys,
                             std::vector<cv::KeyPoint>& mvKeysUn, FrameGridCells& mGrid);

// Frame::ExtractORB(0, im, x0, x1) + UndistortKeyPoints + AssignFeaturesToGrid in one call (what the monocular Frame constructor
// does with its image, src/Frame.cc:307-347); the keypoints stay on the GPU between the extraction and the grid.  Returns the
// extractor's return value (monoIndex), -1 on error.  Note: mvImagePyramid of `ext` is not refreshed by this call.
int ExtractFrame(ORBextractor& ext, const OrbxFrameCalib& calib, const cv::Mat& im, int x0, int x1, std::vector<cv::KeyPoint>& mvKeys,
                 cv::Mat& mDescriptors, std::vector<cv::KeyPoint>& mvKeysUn, FrameGridCells& mGrid);

// Returns nmatches (or -1 on error); vnMatches12 is resized to mvKeysUn1.size(), vbPrevMatched updated in place.
int SearchForInitialization(ORBextractor& ext, const OrbxFrameCalib& calib, const std::vector<cv::KeyPoint>& mvKeysUn1,
                            const cv::Mat& mDescriptors1, const std::vector<cv::KeyPoint>& mvKeysUn2, const cv::Mat& mDescriptors2,
                            const FrameGridCells& mGrid2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                            int windowSize = 10, float nnratio = 0.9f, bool checkOrientation = true);

}  // namespace ORB_SLAM3

#endif
