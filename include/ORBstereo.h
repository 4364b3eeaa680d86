// include/ORBstereo.h -- GPU replacement for the body of Frame::ComputeStereoMatches (reference
// src/Frame.cc:813-990), the first consumer of the extractor's outputs on the stereo path.  The two extractors
// keep their pyramids resident on the GPU, so mvImagePyramid does not have to be downloaded for matching.
//
//   void Frame::ComputeStereoMatches() {
//       ORB_SLAM3::ComputeStereoMatches(*mpORBextractorLeft, *mpORBextractorRight, mvKeys, mDescriptors, mvKeysRight,
//                                       mDescriptorsRight, mb, mbf, mvuRight, mvDepth);
//   }
#ifndef ORBSTEREO_H
#define ORBSTEREO_H

#include <vector>

#include "ORBextractor.h"

namespace ORB_SLAM3 {

// Returns the number of matches kept, or -1 on error (see left.LastError()).  mvuRight / mvDepth are resized to
// mvKeys.size() and filled with -1 where no match was kept, exactly like the reference.
int ComputeStereoMatches(ORBextractor& left, ORBextractor& right, const std::vector<cv::KeyPoint>& mvKeys,
                         const cv::Mat& mDescriptors, const std::vector<cv::KeyPoint>& mvKeysRight,
                         const cv::Mat& mDescriptorsRight, float mb, float mbf, std::vector<float>& mvuRight,
                         std::vector<float>& mvDepth);

}  // namespace ORB_SLAM3

#endif
