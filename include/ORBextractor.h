// include/ORBextractor.h -- drop-in replacement for the reference's ORB_SLAM3::ORBextractor
// (reference inc/ORBextractor.h:31-111; the 5-argument operator() of inc/ORBExtractor.h:55-56 is offered
// as an overload).  Same constructor, same public members, same return conventions; every pixel is
// processed on the GPU through the C-ABI of include/orbx.h.  There is no CPU code path.
//
// Differences a caller can observe (all documented in DESIGN.md section 7):
//   * an image that is not CV_8UC1 makes operator() return -1 (the reference asserts, ORBextractor.cc:1087);
//   * levels too small for the 30-px cell grid / aspect ratio < 0.5 return -1 (undefined behaviour there);
//   * equal-size quadtree nodes are split "later-created first", i.e. the reference's behaviour under a
//     monotonic allocator (its own tie-break compares heap pointers, ORBextractor.cc:689);
//   * a sub-matrix input is treated as isolated at level 0 (never reads pixels outside the ROI).
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <list>
#include <string>
#include <vector>

#include "orbx_cv_compat.hpp"

struct OrbxHandle;

namespace ORB_SLAM3 {

class ExtractorNode {
public:
    ExtractorNode() : bNoMore(false) {}

    // Host-side helper kept for source compatibility (reference ORBextractor.cc:486-542); the extractor
    // itself runs the quadtree on the GPU.
    void DivideNode(ExtractorNode& n1, ExtractorNode& n2, ExtractorNode& n3, ExtractorNode& n4);

    std::vector<cv::KeyPoint> vKeys;
    cv::Point2i UL, UR, BL, BR;
    std::list<ExtractorNode>::iterator lit;
    bool bNoMore;
};

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image.  Mask is ignored, as in the reference.
    // Returns the number of non-lapping ("mono") keypoints, or -1 for an empty/unsupported image.
    int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                   cv::OutputArray _descriptors, std::vector<int>& vLappingArea);
    int operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                   cv::OutputArray _descriptors, std::vector<int>& vLappingArea,
                   std::vector<std::vector<cv::KeyPoint> >& allLevelsKeypoints);

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Level l is a w_l x h_l ROI at (19,19) of a bordered buffer, like the reference (:1173-1177).  The buffers belong to the
    // extractor (one pinned block, refilled by every call that rebuilds the pyramid); clone() a level to keep it longer.
    std::vector<cv::Mat> mvImagePyramid;
    const std::vector<cv::Mat>& GetPyramid() const { return mvImagePyramid; }

    // Stage-wise methods the reference leaves public (its "//protected:" is commented out) and its demos
    // call directly (src/orb_extractor/main_orb_extractor.cpp:44-46).
    void ComputePyramid(cv::Mat image);
    void ComputeKeyPointsOctTree(std::vector<std::vector<cv::KeyPoint> >& allKeypoints);
    std::vector<cv::KeyPoint> DistributeOctTree(const std::vector<cv::KeyPoint>& vToDistributeKeys, const int& minX,
                                                const int& maxX, const int& minY, const int& maxY, const int& nFeatures,
                                                const int& level);

    std::vector<cv::Point> pattern;

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;
    std::vector<int> umax;

    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    // ---- additions (not in the reference) ----
    // Select the CUDA device before the first extraction (default: ORBX_DEVICE env var, else 0).
    void SetDevice(int device);
    // Skip the device-to-host copy of the pyramid after operator() (mvImagePyramid keeps its last content).
    void SetPyramidDownload(bool enable) { mbDownloadPyramid = enable; }
    // Text of the last error (empty when the last call succeeded).
    const std::string& LastError() const { return mLastError; }
    OrbxHandle* NativeHandle();

private:
    bool EnsureHandle();
    bool DownloadPyramid();
    bool FetchAllLevels(std::vector<std::vector<cv::KeyPoint> >& all, int cap);
    int Extract(cv::InputArray image, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors,
                std::vector<int>& vLappingArea, std::vector<std::vector<cv::KeyPoint> >* allLevels);

    OrbxHandle* mpHandle;
    int mnDevice;
    bool mbDownloadPyramid;
    unsigned char* mpPyramidHost;      // pinned block holding the bordered planes of mvImagePyramid (one copy per call)
    size_t mnPyramidHostBytes;
    std::string mLastError;
};

}  // namespace ORB_SLAM3

#endif  // ORBEXTRACTOR_H
