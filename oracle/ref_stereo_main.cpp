// oracle/ref_stereo_main.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Driver around the UNMODIFIED reference functions Frame::ComputeStereoMatches (src/Frame.cc:813-990) and
// ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2349-2365).  oracle/Makefile extracts those two function
// bodies from /root/reference where they lie into oracle/_ref/gen_stereo_body.inc (git-ignored, never copied
// into the repo) and compiles them here against a minimal stand-in for the Frame / ORBmatcher / extractor
// members the function touches.
//
// Input file : int32 'STIN', nlevels, nL, nR; float mb, mbf; float scale[nlevels], invscale[nlevels];
//              nL*28 B left keypoints, nL*32 B left descriptors, nR*28 B, nR*32 B;
//              per level: int32 w, h, then w*h bytes left level, w*h bytes right level.
// Output file: int32 'STOU', nL; nL floats mvuRight; nL floats mvDepth.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cv_shim_stereo.hpp"

using namespace std;

namespace ORB_SLAM3 {

class ORBmatcher {
public:
    static const int TH_LOW;
    static const int TH_HIGH;
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
};
const int ORBmatcher::TH_HIGH = 100;  // src/ORBmatcher.cc:36
const int ORBmatcher::TH_LOW = 50;    // src/ORBmatcher.cc:37

struct FakeExtractor {
    std::vector<cv::Mat> mvImagePyramid;
};

class Frame {
public:
    void ComputeStereoMatches();
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<float> mvuRight, mvDepth;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    float mb, mbf;
    FakeExtractor *mpORBextractorLeft, *mpORBextractorRight;
};

}  // namespace ORB_SLAM3

// cv::Mat::convertTo is a member in OpenCV; the compat Mat has none, so the extracted text is compiled with
// `X.convertTo(X,CV_16S)` rewritten by the preprocessor-free sed step in the Makefile to cv::convertTo16S(X,X).
namespace ORB_SLAM3 {
#include "gen_stereo_body.inc"
}

static bool rd(FILE* f, void* p, size_t n) { return std::fread(p, 1, n, f) == n; }

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s <in.stin> <out.stou>\n", argv[0]); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 1;
    int32_t hdr[4];
    if (!rd(f, hdr, 16) || hdr[0] != 0x4e495453) return 1;
    const int nlevels = hdr[1], nL = hdr[2], nR = hdr[3];
    ORB_SLAM3::Frame fr;
    ORB_SLAM3::FakeExtractor exL, exR;
    fr.mpORBextractorLeft = &exL; fr.mpORBextractorRight = &exR;
    if (!rd(f, &fr.mb, 4) || !rd(f, &fr.mbf, 4)) return 1;
    fr.mvScaleFactors.resize(nlevels); fr.mvInvScaleFactors.resize(nlevels);
    if (!rd(f, fr.mvScaleFactors.data(), 4 * nlevels) || !rd(f, fr.mvInvScaleFactors.data(), 4 * nlevels)) return 1;
    fr.N = nL;
    fr.mvKeys.resize(nL); fr.mvKeysRight.resize(nR);
    fr.mDescriptors = cv::Mat(std::max(nL, 1), 32, CV_8UC1); fr.mDescriptorsRight = cv::Mat(std::max(nR, 1), 32, CV_8UC1);
    if (!rd(f, fr.mvKeys.data(), 28 * (size_t)nL) || !rd(f, fr.mDescriptors.data, 32 * (size_t)nL)) return 1;
    if (!rd(f, fr.mvKeysRight.data(), 28 * (size_t)nR) || !rd(f, fr.mDescriptorsRight.data, 32 * (size_t)nR)) return 1;
    for (int l = 0; l < nlevels; ++l) {
        int32_t wh[2];
        if (!rd(f, wh, 8)) return 1;
        cv::Mat a(wh[1], wh[0], CV_8UC1), b(wh[1], wh[0], CV_8UC1);
        if (!rd(f, a.data, (size_t)wh[0] * wh[1]) || !rd(f, b.data, (size_t)wh[0] * wh[1])) return 1;
        exL.mvImagePyramid.push_back(a); exR.mvImagePyramid.push_back(b);
    }
    std::fclose(f);
    fr.ComputeStereoMatches();
    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 1;
    int32_t oh[2] = {0x554f5453, nL};
    std::fwrite(oh, 4, 2, o);
    std::fwrite(fr.mvuRight.data(), 4, (size_t)nL, o);
    std::fwrite(fr.mvDepth.data(), 4, (size_t)nL, o);
    std::fclose(o);
    return 0;
}
