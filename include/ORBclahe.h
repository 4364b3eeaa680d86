// include/ORBclahe.h -- GPU replacement for the pre-processing step of the reference's demos:
//
//   cv::Ptr<cv::CLAHE> clahe = cv::createCLAHE(3.0, cv::Size(8, 8));          // src/orb_extractor/main_orb_extractor.cpp:19
//   clahe->apply(image, im_clahe);                                           // :22, src/clahe/main_clahe.cpp:11
//
// becomes
//
//   ORB_SLAM3::ApplyCLAHE(extractor, image, im_clahe, 3.0, cv::Size(8, 8));
//
// with im_clahe byte-identical to OpenCV's result (cv2 4.13.0 is the pin; the reference tree holds no CLAHE code).
#ifndef ORBCLAHE_H
#define ORBCLAHE_H

#include "ORBextractor.h"

namespace ORB_SLAM3 {

// src: CV_8UC1; dst is (re)allocated to src's size.  Returns false on error (see ext.LastError()).
bool ApplyCLAHE(ORBextractor& ext, const cv::Mat& src, cv::Mat& dst, double clipLimit = 40.0, cv::Size tileGridSize = cv::Size(8, 8));

}  // namespace ORB_SLAM3

#endif
