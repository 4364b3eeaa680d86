// extractorb_b200/csrc/ORBextractor.cpp -- host side of the drop-in class declared in include/ORBextractor.h.
// It only unwraps cv::InputArray / cv::OutputArray, sizes the caller's containers and forwards to the
// C-ABI (include/orbx.h); all pixel work happens in libextractorb_cuda.so.  Reference being replaced:
// /root/reference/src/orb_extractor/ORBextractor.cc (operator() :1078-1162, constructor :408-475).
#include "../../include/ORBextractor.h"
#include "../../include/ORBstereo.h"
#include "../../include/ORBframe.h"
#include "../../include/ORBclahe.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/orbx.h"

namespace ORB_SLAM3 {

namespace {
const signed char kPatternTable[1024] = {
#include "orb_pattern.inc"
};
static_assert(sizeof(cv::KeyPoint) == sizeof(OrbxKeyPoint), "cv::KeyPoint must be the 28-byte layout the C-ABI writes");
const int kEdge = 19;  // EDGE_THRESHOLD, reference :72
}  // namespace

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      mpHandle(nullptr), mnDevice(0), mbDownloadPyramid(true), mpPyramidHost(nullptr), mnPyramidHostBytes(0) {
    if (const char* env = std::getenv("ORBX_DEVICE")) mnDevice = std::atoi(env);
    // Constructor tables (reference :419-474): plain host arithmetic through the library's device-free entry point, so the
    // accessors are right even when no GPU is present -- ORB-SLAM3 constructs extractors while parsing its settings
    // (reference src/Tracking.cc:768-774) and copies the tables into every Frame (src/Frame.cc:97-103).  Only the
    // handle (device state) is created lazily.
    const size_t nl = (size_t)(nlevels > 0 ? nlevels : 0);
    mvScaleFactor.assign(nl, 1.f);
    mvInvScaleFactor = mvLevelSigma2 = mvInvLevelSigma2 = mvScaleFactor;
    mnFeaturesPerLevel.assign(nl, 0);
    umax.assign(16, 0);
    mvImagePyramid.resize(nl);
    {
        OrbxParams prm;
        std::memset(&prm, 0, sizeof(prm));
        prm.nfeatures = nfeatures; prm.scale_factor = (float)scaleFactor; prm.nlevels = nlevels;
        if (orbx_ctor_tables(&prm, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                             mnFeaturesPerLevel.data(), umax.data()) != ORBX_OK)
            mLastError = "constructor arguments out of range (nlevels 1..16, scaleFactor > 1, nfeatures >= 0)";
    }
    const cv::Point* p0 = nullptr;
    (void)p0;
    pattern.reserve(512);
    for (int i = 0; i < 512; ++i) pattern.push_back(cv::Point(kPatternTable[2 * i], kPatternTable[2 * i + 1]));
    EnsureHandle();
}

ORBextractor::~ORBextractor() {
    mvImagePyramid.clear();
    if (mpPyramidHost) orbx_host_free(mpPyramidHost);
    if (mpHandle) orbx_destroy(mpHandle);
}

void ORBextractor::SetDevice(int device) {
    if (device == mnDevice && mpHandle) return;
    if (mpHandle) { orbx_destroy(mpHandle); mpHandle = nullptr; }
    mnDevice = device;
    EnsureHandle();
}

OrbxHandle* ORBextractor::NativeHandle() {
    EnsureHandle();
    return mpHandle;
}

bool ORBextractor::EnsureHandle() {
    if (mpHandle) return true;
    OrbxParams prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.nfeatures = nfeatures; prm.scale_factor = (float)scaleFactor; prm.nlevels = nlevels;
    prm.ini_th_fast = iniThFAST; prm.min_th_fast = minThFAST;
    const int rc = orbx_create(&prm, mnDevice, &mpHandle);
    if (rc != ORBX_OK) {
        mpHandle = nullptr;
        mLastError = std::string("orbx_create: ") + orbx_status_string(rc);
        return false;
    }
    mLastError.clear();
    return true;
}

// mvImagePyramid (reference :1173-1177): ONE device-to-host copy of the frame's whole bordered block into a pinned buffer the
// extractor owns and reuses; the level Mats are headers into it (ROI at (19,19) of a (w+38) x (h+38) plane whose row step is
// the device pitch), valid until the next call that rebuilds the pyramid -- the reference also replaces them on every call.
bool ORBextractor::DownloadPyramid() {
    int w0 = 0, h0 = 0;
    if (orbx_get_level_size(mpHandle, 0, &w0, &h0) != ORBX_OK) return false;
    size_t bytes = 0;
    std::vector<size_t> off((size_t)nlevels);
    std::vector<int32_t> pitch((size_t)nlevels), lw((size_t)nlevels), lh((size_t)nlevels);
    if (orbx_get_pyramid_layout(mpHandle, w0, h0, &bytes, off.data(), pitch.data(), lw.data(), lh.data()) != ORBX_OK) return false;
    if (bytes > mnPyramidHostBytes) {
        for (int l = 0; l < nlevels; ++l) mvImagePyramid[l] = cv::Mat();   // headers into the old buffer
        orbx_host_free(mpPyramidHost);
        mpPyramidHost = static_cast<unsigned char*>(orbx_host_alloc(bytes, 0));
        mnPyramidHostBytes = mpPyramidHost ? bytes : 0;
        if (!mpPyramidHost) return false;
    }
    if (orbx_download_pyramid(mpHandle, 0, mpPyramidHost, mnPyramidHostBytes) != ORBX_OK) return false;
    for (int l = 0; l < nlevels; ++l) {
        unsigned char* plane = mpPyramidHost + off[(size_t)l] + (ORBX_PLANE_PADL - kEdge);      // byte of border column -19 in plane row 0
        cv::Mat temp(lh[(size_t)l] + 2 * kEdge, lw[(size_t)l] + 2 * kEdge, CV_8UC1, plane, (size_t)pitch[(size_t)l]);
        mvImagePyramid[l] = temp(cv::Rect(kEdge, kEdge, lw[(size_t)l], lh[(size_t)l]));   // :1177
    }
    return true;
}

int ORBextractor::Extract(cv::InputArray _image, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors,
                          std::vector<int>& vLappingArea, std::vector<std::vector<cv::KeyPoint> >* allLevels) {
    if (_image.empty()) return -1;                                                       // :1083
    cv::Mat image = _image.getMat();
    if (image.type() != CV_8UC1) { mLastError = "image must be CV_8UC1"; return -1; }     // reference: assert, :1087
    if (vLappingArea.size() < 2) { mLastError = "vLappingArea needs two entries"; return -1; }
    if (!EnsureHandle()) return -1;
    const int cap = orbx_max_keypoints(mpHandle, image.cols, image.rows);
    if (cap < 0) { mLastError = orbx_last_error(mpHandle); return -1; }
    std::vector<cv::KeyPoint> kps((size_t)cap);
    std::vector<unsigned char> desc((size_t)cap * 32);
    int n = 0, mono = 0;
    const int rc = orbx_extract(mpHandle, image.data, image.cols, image.rows, (size_t)image.step, vLappingArea[0], vLappingArea[1],
                                reinterpret_cast<OrbxKeyPoint*>(kps.data()), desc.data(), cap, &n, &mono);
    if (rc != ORBX_OK) { mLastError = orbx_last_error(mpHandle); return -1; }
    mLastError.clear();
    if (n == 0) {
        _descriptors.release();                                                          // :1102-1103
    } else {
        _descriptors.create(n, 32, CV_8U);                                               // :1106
        cv::Mat d = _descriptors.getMat();
        for (int r = 0; r < n; ++r) std::memcpy(d.ptr(r), desc.data() + (size_t)r * 32, 32);
    }
    kps.resize((size_t)n);
    _keypoints.swap(kps);                                                                // fresh vector, :1112
    if (allLevels && !FetchAllLevels(*allLevels, cap)) { mLastError = orbx_last_error(mpHandle); return -1; }   // :1094
    if (mbDownloadPyramid && !DownloadPyramid()) { mLastError = orbx_last_error(mpHandle); return -1; }
    return mono;                                                                         // monoIndex, :1161
}

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray, std::vector<cv::KeyPoint>& _keypoints,
                             cv::OutputArray _descriptors, std::vector<int>& vLappingArea) {
    return Extract(_image, _keypoints, _descriptors, vLappingArea, nullptr);
}

int ORBextractor::operator()(cv::InputArray _image, cv::InputArray, std::vector<cv::KeyPoint>& _keypoints,
                             cv::OutputArray _descriptors, std::vector<int>& vLappingArea,
                             std::vector<std::vector<cv::KeyPoint> >& allLevelsKeypoints) {
    return Extract(_image, _keypoints, _descriptors, vLappingArea, &allLevelsKeypoints);
}

void ORBextractor::ComputePyramid(cv::Mat image) {
    if (image.empty() || image.type() != CV_8UC1 || !EnsureHandle()) return;
    if (orbx_compute_pyramid(mpHandle, image.data, image.cols, image.rows, (size_t)image.step) != ORBX_OK) {
        mLastError = orbx_last_error(mpHandle);
        return;
    }
    mLastError.clear();
    DownloadPyramid();
}

// allKeypoints (level coordinates, angle set) of the resident frame: one read-back for all levels.
bool ORBextractor::FetchAllLevels(std::vector<std::vector<cv::KeyPoint> >& all, int cap) {
    all.assign((size_t)nlevels, std::vector<cv::KeyPoint>());
    if (cap <= 0) {
        int w0 = 0, h0 = 0;
        if (orbx_get_level_size(mpHandle, 0, &w0, &h0) != ORBX_OK) return false;
        cap = orbx_max_keypoints(mpHandle, w0, h0);
        if (cap <= 0) return false;
    }
    std::vector<cv::KeyPoint> flat((size_t)cap);
    std::vector<int32_t> counts((size_t)nlevels, 0);
    int total = 0;
    if (orbx_get_all_level_keypoints(mpHandle, 0, reinterpret_cast<OrbxKeyPoint*>(flat.data()), cap, counts.data(), &total) != ORBX_OK) return false;
    size_t at = 0;
    for (int l = 0; l < nlevels; ++l) {
        all[(size_t)l].assign(flat.begin() + at, flat.begin() + at + (size_t)counts[(size_t)l]);
        at += (size_t)counts[(size_t)l];
    }
    return true;
}

void ORBextractor::ComputeKeyPointsOctTree(std::vector<std::vector<cv::KeyPoint> >& allKeypoints) {
    allKeypoints.assign((size_t)nlevels, std::vector<cv::KeyPoint>());                   // :775
    if (!EnsureHandle()) return;
    if (orbx_compute_keypoints_octtree(mpHandle) != ORBX_OK) { mLastError = orbx_last_error(mpHandle); return; }
    mLastError.clear();
    if (!FetchAllLevels(allKeypoints, 0)) mLastError = orbx_last_error(mpHandle);
}

std::vector<cv::KeyPoint> ORBextractor::DistributeOctTree(const std::vector<cv::KeyPoint>& vToDistributeKeys, const int& minX,
                                                          const int& maxX, const int& minY, const int& maxY, const int& N,
                                                          const int& /*level*/) {
    std::vector<cv::KeyPoint> out;
    if (!EnsureHandle()) return out;
    const int nIni = (int)std::lround((double)(maxX - minX) / (double)(maxY - minY > 0 ? maxY - minY : 1));
    const int cap = std::max(N + 2, 4 * std::max(nIni, 1)) + 8;
    out.resize((size_t)cap);
    int n = 0;
    const int rc = orbx_distribute_octtree(mpHandle, reinterpret_cast<const OrbxKeyPoint*>(vToDistributeKeys.data()),
                                           (int)vToDistributeKeys.size(), minX, maxX, minY, maxY, N,
                                           reinterpret_cast<OrbxKeyPoint*>(out.data()), cap, &n);
    if (rc != ORBX_OK) { mLastError = orbx_last_error(mpHandle); out.clear(); return out; }
    mLastError.clear();
    out.resize((size_t)n);
    return out;
}

// Frame::ComputeStereoMatches (reference src/Frame.cc:813-990) through the C-ABI.
int ComputeStereoMatches(ORBextractor& left, ORBextractor& right, const std::vector<cv::KeyPoint>& mvKeys,
                         const cv::Mat& mDescriptors, const std::vector<cv::KeyPoint>& mvKeysRight,
                         const cv::Mat& mDescriptorsRight, float mb, float mbf, std::vector<float>& mvuRight,
                         std::vector<float>& mvDepth) {
    const int N = (int)mvKeys.size(), Nr = (int)mvKeysRight.size();
    mvuRight.assign((size_t)N, -1.0f);                                                   // :815-816
    mvDepth.assign((size_t)N, -1.0f);
    if (N == 0) return 0;
    OrbxHandle* hl = left.NativeHandle();
    OrbxHandle* hr = right.NativeHandle();
    if (!hl || !hr) return -1;
    // descriptors as contiguous n x 32 bytes (cv::Mat rows may be strided)
    std::vector<unsigned char> dl((size_t)N * 32), dr((size_t)std::max(Nr, 1) * 32);
    for (int i = 0; i < N; ++i) std::memcpy(&dl[(size_t)i * 32], mDescriptors.ptr(i), 32);
    for (int i = 0; i < Nr; ++i) std::memcpy(&dr[(size_t)i * 32], mDescriptorsRight.ptr(i), 32);
    int kept = 0;
    const int rc = orbx_stereo_match(hl, hr, reinterpret_cast<const OrbxKeyPoint*>(mvKeys.data()), dl.data(), N,
                                     reinterpret_cast<const OrbxKeyPoint*>(mvKeysRight.data()), dr.data(), Nr, mb, mbf,
                                     mvuRight.data(), mvDepth.data(), &kept);
    return rc == ORBX_OK ? kept : -1;
}

// ---- include/ORBclahe.h: cv::CLAHE::apply through the C-ABI ----
bool ApplyCLAHE(ORBextractor& ext, const cv::Mat& src, cv::Mat& dst, double clipLimit, cv::Size tileGridSize) {
    OrbxHandle* h = ext.NativeHandle();
    if (!h || src.empty() || src.type() != CV_8UC1) return false;
    dst.create(src.rows, src.cols, CV_8UC1);
    return orbx_clahe(h, src.data, ORBX_MEM_HOST, 1, src.cols, src.rows, src.step, src.step * (size_t)src.rows, clipLimit, tileGridSize.width,
                      tileGridSize.height, dst.data, ORBX_MEM_HOST, dst.step, dst.step * (size_t)dst.rows, nullptr) == ORBX_OK;
}

// ---- include/ORBframe.h: Frame post-processing and SearchForInitialization through the C-ABI ----
bool ComputeImageBounds(ORBextractor& ext, const cv::Mat& K, const cv::Mat& distCoef, int cols, int rows, OrbxFrameCalib& calib) {
    OrbxHandle* h = ext.NativeHandle();
    if (!h || K.rows != 3 || K.cols != 3 || K.type() != CV_32FC1 || distCoef.type() != CV_32FC1) return false;
    const int nd = distCoef.rows * distCoef.cols;
    if (nd < 4 || nd > 5) return false;
    std::memset(&calib, 0, sizeof(calib));
    calib.fx = K.at<float>(0, 0); calib.fy = K.at<float>(1, 1); calib.cx = K.at<float>(0, 2); calib.cy = K.at<float>(1, 2);
    for (int i = 0; i < nd; ++i) calib.dist[i] = distCoef.rows == 1 ? distCoef.at<float>(0, i) : distCoef.at<float>(i, 0);
    calib.n_dist = nd;
    return orbx_frame_image_bounds(h, &calib, cols, rows) == ORBX_OK;
}

int UndistortAndAssignToGrid(ORBextractor& ext, const OrbxFrameCalib& calib, const std::vector<cv::KeyPoint>& mvKeys,
                             std::vector<cv::KeyPoint>& mvKeysUn, FrameGridCells& mGrid) {
    OrbxHandle* h = ext.NativeHandle();
    if (!h) return -1;
    const int N = (int)mvKeys.size();
    mvKeysUn.resize((size_t)N);
    std::vector<int32_t> start(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, 0), items((size_t)std::max(N, 1));
    int placed = 0;
    const int rc = orbx_frame_undistort_grid(h, &calib, reinterpret_cast<const OrbxKeyPoint*>(mvKeys.data()), N,
                                             reinterpret_cast<OrbxKeyPoint*>(mvKeysUn.data()), start.data(), items.data(), &placed);
    if (rc != ORBX_OK) return -1;
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) {
            const int c = i * FRAME_GRID_ROWS + j;
            mGrid[i][j].assign(items.begin() + start[c], items.begin() + start[c + 1]);
        }
    return placed;
}

int ExtractFrame(ORBextractor& ext, const OrbxFrameCalib& calib, const cv::Mat& im, int x0, int x1, std::vector<cv::KeyPoint>& mvKeys,
                 cv::Mat& mDescriptors, std::vector<cv::KeyPoint>& mvKeysUn, FrameGridCells& mGrid) {
    OrbxHandle* h = ext.NativeHandle();
    if (!h || im.empty() || im.type() != CV_8UC1) return -1;
    const int cap = orbx_max_keypoints(h, im.cols, im.rows);
    if (cap <= 0) return -1;
    std::vector<OrbxKeyPoint> k((size_t)cap), u((size_t)cap);
    std::vector<unsigned char> d((size_t)cap * 32);
    std::vector<int32_t> start(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, 0), items((size_t)cap);
    int n = 0, mono = 0, placed = 0;
    if (orbx_extract_frame(h, im.data, im.cols, im.rows, im.step, x0, x1, &calib, k.data(), d.data(), cap, &n, &mono, u.data(), start.data(),
                           items.data(), &placed) != ORBX_OK)
        return -1;
    static_assert(sizeof(cv::KeyPoint) == sizeof(OrbxKeyPoint), "cv::KeyPoint layout");
    mvKeys.resize((size_t)n); mvKeysUn.resize((size_t)n);
    if (n > 0) {
        std::memcpy(mvKeys.data(), k.data(), (size_t)n * sizeof(OrbxKeyPoint));
        std::memcpy(mvKeysUn.data(), u.data(), (size_t)n * sizeof(OrbxKeyPoint));
        mDescriptors.create(n, 32, CV_8UC1);
        for (int i = 0; i < n; ++i) std::memcpy(mDescriptors.ptr(i), &d[(size_t)i * 32], 32);
    } else {
        mDescriptors.release();
    }
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) {
            const int c = i * FRAME_GRID_ROWS + j;
            mGrid[i][j].assign(items.begin() + start[c], items.begin() + start[c + 1]);
        }
    return mono;
}

int SearchForInitialization(ORBextractor& ext, const OrbxFrameCalib& calib, const std::vector<cv::KeyPoint>& mvKeysUn1,
                            const cv::Mat& mDescriptors1, const std::vector<cv::KeyPoint>& mvKeysUn2, const cv::Mat& mDescriptors2,
                            const FrameGridCells& mGrid2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                            int windowSize, float nnratio, bool checkOrientation) {
    OrbxHandle* h = ext.NativeHandle();
    const int n1 = (int)mvKeysUn1.size(), n2 = (int)mvKeysUn2.size();
    vnMatches12.assign((size_t)n1, -1);                                                         // src/ORBmatcher.cc:708
    if (!h || (int)vbPrevMatched.size() < n1) return -1;
    if (n1 == 0) return 0;
    std::vector<unsigned char> d1((size_t)n1 * 32), d2((size_t)std::max(n2, 1) * 32);
    for (int i = 0; i < n1; ++i) std::memcpy(&d1[(size_t)i * 32], mDescriptors1.ptr(i), 32);
    for (int i = 0; i < n2; ++i) std::memcpy(&d2[(size_t)i * 32], mDescriptors2.ptr(i), 32);
    std::vector<int32_t> start(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, 0), items;
    items.reserve((size_t)n2);
    for (int i = 0; i < FRAME_GRID_COLS; ++i)
        for (int j = 0; j < FRAME_GRID_ROWS; ++j) {
            for (size_t k = 0; k < mGrid2[i][j].size(); ++k) items.push_back((int32_t)mGrid2[i][j][k]);
            start[i * FRAME_GRID_ROWS + j + 1] = (int32_t)items.size();
        }
    if (items.empty()) items.push_back(0);
    int nmatches = 0;
    static_assert(sizeof(cv::Point2f) == 8, "cv::Point2f layout");
    const int rc = orbx_search_for_initialization(h, &calib, reinterpret_cast<const OrbxKeyPoint*>(mvKeysUn1.data()), d1.data(), n1,
                                                  reinterpret_cast<const OrbxKeyPoint*>(mvKeysUn2.data()), d2.data(), n2, start.data(),
                                                  items.data(), reinterpret_cast<float*>(vbPrevMatched.data()), windowSize, nnratio,
                                                  checkOrientation ? 1 : 0, vnMatches12.data(), &nmatches);
    return rc == ORBX_OK ? nmatches : -1;
}

// Geometry of one quadtree split (reference :486-542); host-side, for callers that use the node type.
void ExtractorNode::DivideNode(ExtractorNode& n1, ExtractorNode& n2, ExtractorNode& n3, ExtractorNode& n4) {
    const int halfX = (int)std::ceil(static_cast<float>(UR.x - UL.x) / 2);
    const int halfY = (int)std::ceil(static_cast<float>(BR.y - UL.y) / 2);
    const int xm = UL.x + halfX, ym = UL.y + halfY;
    ExtractorNode* child[4] = {&n1, &n2, &n3, &n4};
    n1.UL = UL;                     n1.UR = cv::Point2i(xm, UL.y);   n1.BL = cv::Point2i(UL.x, ym);  n1.BR = cv::Point2i(xm, ym);
    n2.UL = n1.UR;                  n2.UR = UR;                      n2.BL = n1.BR;                  n2.BR = cv::Point2i(UR.x, ym);
    n3.UL = n1.BL;                  n3.UR = n1.BR;                   n3.BL = BL;                     n3.BR = cv::Point2i(xm, BL.y);
    n4.UL = n3.UR;                  n4.UR = n2.BR;                   n4.BL = n3.BR;                  n4.BR = BR;
    for (int c = 0; c < 4; ++c) { child[c]->vKeys.clear(); child[c]->vKeys.reserve(vKeys.size()); child[c]->bNoMore = false; }
    for (size_t i = 0; i < vKeys.size(); ++i) {
        const cv::KeyPoint& kp = vKeys[i];
        const int c = (kp.pt.x < xm ? 0 : 1) + (kp.pt.y < ym ? 0 : 2);
        child[c]->vKeys.push_back(kp);
    }
    for (int c = 0; c < 4; ++c) if (child[c]->vKeys.size() == 1) child[c]->bNoMore = true;
}

}  // namespace ORB_SLAM3
