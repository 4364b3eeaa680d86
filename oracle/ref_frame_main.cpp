// oracle/ref_frame_main.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Driver around the UNMODIFIED reference functions Frame::UndistortKeyPoints, ComputeImageBounds,
// AssignFeaturesToGrid, PosInGrid, GetFeaturesInArea (src/Frame.cc:383-417, :655-812) and
// ORBmatcher::SearchForInitialization, ComputeThreeMaxima, DescriptorDistance (src/ORBmatcher.cc:705-814,
// :2303-2365).  oracle/Makefile extracts those function bodies from /root/reference where they lie into
// oracle/_ref/gen_frame_body.inc (git-ignored, never copied into the repo) and compiles them here, untouched,
// against a minimal stand-in for the Frame / ORBmatcher members they use.
//
// Input : int32 'FRIN', w, h, n1, n2, window, check_orientation, n_dist, has_prev; float nnratio, fx, fy, cx, cy,
//         dist[5]; n1*28 B keys1, n1*32 B desc1, n2*28 B keys2, n2*32 B desc2; [n1*2 floats vbPrevMatched].
// Output: int32 'FROU', n1, n2; 4 floats mnMinX mnMaxX mnMinY mnMaxY; n1*28 B mvKeysUn(1); n2*28 B mvKeysUn(2);
//         per frame: 3073 int32 cell offsets (cell = ix*48+iy), then the items; int32 nmatches; n1 int32 vnMatches12;
//         n1*2 floats vbPrevMatched (updated).
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cv_shim_frame.hpp"

using namespace std;

#define FRAME_GRID_ROWS 48  // inc/Frame.h:39
#define FRAME_GRID_COLS 64  // inc/Frame.h:40

namespace ORB_SLAM3 {

class GeometricCamera { public: virtual ~GeometricCamera() {} };
class Pinhole : public GeometricCamera {
public:
    cv::Mat K;
    cv::Mat toK() { return K.clone(); }
};

class Frame {
public:
    void AssignFeaturesToGrid();
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1,
                                     const bool bRight = false) const;
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    void UndistortKeyPoints();
    void ComputeImageBounds(const cv::Mat& imLeft);

    int N, Nleft;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    cv::Mat mDescriptors;
    cv::Mat mK, mDistCoef;
    GeometricCamera* mpCamera;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    std::vector<std::size_t> mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;  // src/Frame.cc:40
float Frame::mnMinX, Frame::mnMinY, Frame::mnMaxX, Frame::mnMaxY;  // :39

class ORBmatcher {
public:
    ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}  // src/ORBmatcher.cc:40
    static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                int windowSize = 10);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};
const int ORBmatcher::TH_HIGH = 100;     // src/ORBmatcher.cc:36
const int ORBmatcher::TH_LOW = 50;       // :37
const int ORBmatcher::HISTO_LENGTH = 30; // :38

#include "gen_frame_body.inc"

}  // namespace ORB_SLAM3

static bool rd(FILE* f, void* p, size_t n) { return n == 0 || std::fread(p, 1, n, f) == n; }

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: %s <in.frin> <out.frou>\n", argv[0]); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) return 1;
    int32_t hdr[9];
    float fl[10];
    if (!rd(f, hdr, sizeof hdr) || hdr[0] != 0x4e495246 || !rd(f, fl, sizeof fl)) return 1;
    const int w = hdr[1], h = hdr[2], n1 = hdr[3], n2 = hdr[4], window = hdr[5], checkOri = hdr[6], nDist = hdr[7], hasPrev = hdr[8];
    ORB_SLAM3::Pinhole cam;
    cam.K = cv::Mat::zeros(3, 3, CV_32FC1);
    cam.K.at<float>(0, 0) = fl[1]; cam.K.at<float>(1, 1) = fl[2]; cam.K.at<float>(0, 2) = fl[3]; cam.K.at<float>(1, 2) = fl[4];
    cam.K.at<float>(2, 2) = 1.f;
    cv::Mat dist(nDist, 1, CV_32FC1);
    for (int i = 0; i < nDist; ++i) dist.at<float>(i) = fl[5 + i];
    ORB_SLAM3::Frame F[2];
    const int ns[2] = {n1, n2};
    for (int k = 0; k < 2; ++k) {
        ORB_SLAM3::Frame& fr = F[k];
        fr.N = ns[k]; fr.Nleft = -1; fr.mpCamera = &cam; fr.mK = cam.toK(); fr.mDistCoef = dist.clone();
        fr.mvKeys.resize(ns[k]);
        fr.mDescriptors = cv::Mat(std::max(ns[k], 1), 32, CV_8UC1);
        if (!rd(f, fr.mvKeys.data(), 28 * (size_t)ns[k]) || !rd(f, fr.mDescriptors.data, 32 * (size_t)ns[k])) return 1;
    }
    std::vector<cv::Point2f> prev(n1);
    if (hasPrev && !rd(f, prev.data(), 8 * (size_t)n1)) return 1;
    std::fclose(f);

    // What the monocular Frame constructor does after ExtractORB (src/Frame.cc:318-347).
    cv::Mat im(h, w, CV_8UC1);
    for (int k = 0; k < 2; ++k) {
        F[k].UndistortKeyPoints();
        if (k == 0) {
            F[k].ComputeImageBounds(im);
            ORB_SLAM3::Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(ORB_SLAM3::Frame::mnMaxX - ORB_SLAM3::Frame::mnMinX);
            ORB_SLAM3::Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(ORB_SLAM3::Frame::mnMaxY - ORB_SLAM3::Frame::mnMinY);
        }
        F[k].AssignFeaturesToGrid();
    }
    if (!hasPrev)  // Tracking::MonocularInitialization: mvbPrevMatched[i] = mCurrentFrame.mvKeysUn[i].pt
        for (int i = 0; i < n1; ++i) prev[i] = F[0].mvKeysUn[i].pt;
    ORB_SLAM3::ORBmatcher matcher(fl[0], checkOri != 0);
    std::vector<int> m12;
    const int nmatches = matcher.SearchForInitialization(F[0], F[1], prev, m12, window);

    FILE* o = std::fopen(argv[2], "wb");
    if (!o) return 1;
    int32_t oh[3] = {0x554f5246, n1, n2};
    std::fwrite(oh, 4, 3, o);
    float b[4] = {ORB_SLAM3::Frame::mnMinX, ORB_SLAM3::Frame::mnMaxX, ORB_SLAM3::Frame::mnMinY, ORB_SLAM3::Frame::mnMaxY};
    std::fwrite(b, 4, 4, o);
    for (int k = 0; k < 2; ++k) std::fwrite(F[k].mvKeysUn.data(), 28, (size_t)ns[k], o);
    for (int k = 0; k < 2; ++k) {
        std::vector<int32_t> start(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, 0), items;
        for (int ix = 0; ix < FRAME_GRID_COLS; ++ix)
            for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) {
                for (size_t j = 0; j < F[k].mGrid[ix][iy].size(); ++j) items.push_back((int32_t)F[k].mGrid[ix][iy][j]);
                start[ix * FRAME_GRID_ROWS + iy + 1] = (int32_t)items.size();
            }
        std::fwrite(start.data(), 4, start.size(), o);
        std::fwrite(items.data(), 4, items.size(), o);
    }
    int32_t nm = nmatches;
    std::fwrite(&nm, 4, 1, o);
    std::vector<int32_t> m(m12.begin(), m12.end());
    std::fwrite(m.data(), 4, m.size(), o);
    std::fwrite(prev.data(), 8, (size_t)n1, o);
    std::fclose(o);
    return 0;
}
