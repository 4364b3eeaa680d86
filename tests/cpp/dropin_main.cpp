// tests/cpp/dropin_main.cpp -- drives the drop-in C++ class exactly like oracle/ref_main.cpp drives the
// unmodified reference (same frame / result file formats), so that tests/test_cpp_dropin.py can compare the
// two programs' outputs byte for byte.  Also exercises the accessors, the stage-wise public methods and the
// two-thread stereo pattern of the reference's caller (src/Frame.cc:109-112).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "ORBExtractor.h"   // global-namespace spelling (reference inc/ORBExtractor.h), includes ORBextractor.h
#include "ORBstereo.h"
#include "ORBframe.h"
#include "ORBclahe.h"

static bool load_frames(const char* path, int& n, int& w, int& h, std::vector<uint8_t>& data) {
    FILE* fp = std::fopen(path, "rb");
    if (!fp) return false;
    int32_t hdr[4];
    if (std::fread(hdr, 4, 4, fp) != 4 || hdr[0] != 0x4642524f) { std::fclose(fp); return false; }
    n = hdr[1]; w = hdr[2]; h = hdr[3];
    data.resize((size_t)n * w * h);
    bool ok = std::fread(data.data(), 1, data.size(), fp) == data.size();
    std::fclose(fp);
    return ok;
}

// tables nfeatures scale nlevels ini min: the constructor tables through the class's accessors and public members, in the
// format of `oracle/_ref/ref_extract tables` (reference inc/ORBextractor.h:63-83).  Needs no GPU: the tables are host arithmetic.
static int cmd_tables(int argc, char** argv) {
    if (argc < 7) return 2;
    ORB_SLAM3::ORBextractor ex(std::atoi(argv[2]), (float)std::atof(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), std::atoi(argv[6]));
    auto pf = [](const char* name, const std::vector<float>& v) {
        std::printf("%s", name);
        for (size_t i = 0; i < v.size(); ++i) { uint32_t u; std::memcpy(&u, &v[i], 4); std::printf(" %08x", u); }
        std::printf("\n");
    };
    auto pi = [](const char* name, const std::vector<int>& v) {
        std::printf("%s", name);
        for (size_t i = 0; i < v.size(); ++i) std::printf(" %d", v[i]);
        std::printf("\n");
    };
    std::printf("levels %d\n", ex.GetLevels());
    { float f = ex.GetScaleFactor(); uint32_t u; std::memcpy(&u, &f, 4); std::printf("scaleFactor %08x\n", u); }
    pf("mvScaleFactor", ex.GetScaleFactors());
    pf("mvInvScaleFactor", ex.GetInverseScaleFactors());
    pf("mvLevelSigma2", ex.GetScaleSigmaSquares());
    pf("mvInvLevelSigma2", ex.GetInverseScaleSigmaSquares());
    pi("mnFeaturesPerLevel", ex.mnFeaturesPerLevel);
    pi("umax", ex.umax);
    return 0;
}

// latency <frames.orbf> iters nfeatures scale nlevels ini min: wall-clock latency of the drop-in operator() itself (what a
// Frame constructor pays, reference src/Frame.cc:419-427) -- 5- and 6-argument overloads, with and without the download of
// mvImagePyramid.  Prints one JSON object (microseconds).
static int cmd_latency(int argc, char** argv) {
    if (argc < 9) return 2;
    int n, w, h;
    std::vector<uint8_t> frames;
    if (!load_frames(argv[2], n, w, h, frames)) { std::fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    const int iters = std::atoi(argv[3]);
    ORB_SLAM3::ORBextractor ex(std::atoi(argv[4]), (float)std::atof(argv[5]), std::atoi(argv[6]), std::atoi(argv[7]), std::atoi(argv[8]));
    std::printf("{\"width\": %d, \"height\": %d, \"iters\": %d", w, h, iters);
    const char* names[4] = {"five_arg_with_pyramid", "five_arg_no_pyramid", "six_arg_with_pyramid", "six_arg_no_pyramid"};
    for (int mode = 0; mode < 4; ++mode) {
        ex.SetPyramidDownload((mode & 1) == 0);
        std::vector<double> us;
        size_t nk = 0;
        for (int it = 0; it < iters + 20; ++it) {
            cv::Mat img(h, w, CV_8UC1, frames.data() + (size_t)(it % n) * w * h);
            std::vector<cv::KeyPoint> kps;
            cv::Mat desc;
            std::vector<int> lap = {0, 0};
            std::vector<std::vector<cv::KeyPoint> > lvl;
            const auto t0 = std::chrono::steady_clock::now();
            const int ret = mode < 2 ? ex(img, cv::Mat(), kps, desc, lap) : ex(img, cv::Mat(), kps, desc, lap, lvl);
            const auto t1 = std::chrono::steady_clock::now();
            if (ret < 0) { std::fprintf(stderr, "latency: %s\n", ex.LastError().c_str()); return 3; }
            if (it >= 20) us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
            nk = kps.size();
        }
        std::sort(us.begin(), us.end());
        std::printf(", \"%s\": {\"p50_us\": %.2f, \"p99_us\": %.2f, \"keypoints\": %zu}", names[mode], us[us.size() / 2],
                    us[std::min(us.size() - 1, (size_t)(us.size() * 0.99))], nk);
    }
    std::printf("}\n");
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !std::strcmp(argv[1], "tables")) return cmd_tables(argc, argv);
    if (argc >= 2 && !std::strcmp(argv[1], "latency")) return cmd_latency(argc, argv);
    if (argc < 12 || std::strcmp(argv[1], "run")) {
        std::fprintf(stderr, "usage: %s run <frames.orbf> <out.orbr> nfeatures scale nlevels ini min lap0 lap1 dump_pyr [mode]\n", argv[0]);
        return 2;
    }
    int n, w, h;
    std::vector<uint8_t> frames;
    if (!load_frames(argv[2], n, w, h, frames)) { std::fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    const int nfeatures = std::atoi(argv[4]); const float scale = (float)std::atof(argv[5]); const int nlevels = std::atoi(argv[6]);
    const int ini = std::atoi(argv[7]), mn = std::atoi(argv[8]), lap0 = std::atoi(argv[9]), lap1 = std::atoi(argv[10]);
    const int dump = std::atoi(argv[11]);
    const char* mode = argc > 12 ? argv[12] : "six";
    FILE* out = std::fopen(argv[3], "wb");
    if (!out) return 1;
    int32_t hdr[4] = {0x5242524f, n, nlevels, dump};
    std::fwrite(hdr, 4, 4, out);

    if (!std::strcmp(mode, "stereo")) {
        // frames 0 / 1 = left / right image; the two extractors run on two host threads like the reference's stereo
        // Frame constructor (src/Frame.cc:109-112), then Frame::ComputeStereoMatches' replacement.  Output:
        // int32 nL, nL floats mvuRight, nL floats mvDepth, int32 kept.  argv[9] / argv[10] carry bf and fx here.
        if (n < 2) return 7;
        const float mbf = (float)std::atof(argv[9]), mb = mbf / (float)std::atof(argv[10]);
        ORB_SLAM3::ORBextractor exl(nfeatures, scale, nlevels, ini, mn), exr(nfeatures, scale, nlevels, ini, mn);
        cv::Mat iml(h, w, CV_8UC1, frames.data()), imr(h, w, CV_8UC1, frames.data() + (size_t)w * h);
        std::vector<cv::KeyPoint> kl, kr; cv::Mat dl, dr; std::vector<int> lapl = {0, 0}, lapr = {0, 0};
        std::thread tl([&]() { exl(iml, cv::Mat(), kl, dl, lapl); });
        std::thread tr([&]() { exr(imr, cv::Mat(), kr, dr, lapr); });
        tl.join(); tr.join();
        std::vector<float> ur, dp;
        const int kept = ORB_SLAM3::ComputeStereoMatches(exl, exr, kl, dl, kr, dr, mb, mbf, ur, dp);
        if (kept < 0) { std::fprintf(stderr, "stereo: %s\n", exl.LastError().c_str()); return 8; }
        std::fclose(out);
        out = std::fopen(argv[3], "wb");
        int32_t nl32 = (int32_t)kl.size(), kept32 = kept;
        std::fwrite(&nl32, 4, 1, out);
        std::fwrite(ur.data(), 4, ur.size(), out);
        std::fwrite(dp.data(), 4, dp.size(), out);
        std::fwrite(&kept32, 4, 1, out);
        std::fclose(out);
        return 0;
    }

    if (!std::strcmp(mode, "clahe")) {
        // cv::createCLAHE(clip, Size(tx, ty))->apply on every frame (include/ORBclahe.h); argv[9] / argv[10] carry tx / ty,
        // argv[5] the clip limit.  Output: the processed frames, w*h bytes each.
        ORB_SLAM3::ORBextractor ex(nfeatures, 1.2f, nlevels, ini, mn);
        std::fclose(out);
        out = std::fopen(argv[3], "wb");
        for (int i = 0; i < n; ++i) {
            cv::Mat im(h, w, CV_8UC1, frames.data() + (size_t)i * w * h), res;
            if (!ORB_SLAM3::ApplyCLAHE(ex, im, res, (double)scale, cv::Size(lap0, lap1))) { std::fprintf(stderr, "clahe: %s\n", ex.LastError().c_str()); return 13; }
            for (int r = 0; r < h; ++r) std::fwrite(res.ptr(r), 1, (size_t)w, out);
        }
        std::fclose(out);
        return 0;
    }

    if (!std::strcmp(mode, "frame")) {
        // frames 0 / 1 = two views.  What the monocular Frame constructor does after ExtractORB (src/Frame.cc:307-347) and
        // Tracking::MonocularInitialization's matcher call (src/Tracking.cc: ORBmatcher matcher(0.9,true);
        // matcher.SearchForInitialization(mInitialFrame, mCurrentFrame, mvbPrevMatched, mvIniMatches, 100)), through
        // include/ORBframe.h.  argv[13..]: fx fy cx cy n_dist d0..d4 window check_orientation.  The output has the layout of
        // oracle/ref_frame_main.cpp so that one parser reads both programs.
        if (n < 2 || argc < 25) return 7;
        cv::Mat K = cv::Mat::zeros(3, 3, CV_32FC1);
        K.at<float>(0, 0) = (float)std::atof(argv[13]); K.at<float>(1, 1) = (float)std::atof(argv[14]);
        K.at<float>(0, 2) = (float)std::atof(argv[15]); K.at<float>(1, 2) = (float)std::atof(argv[16]); K.at<float>(2, 2) = 1.f;
        const int nd = std::atoi(argv[17]);
        cv::Mat dist(nd, 1, CV_32FC1);
        for (int i = 0; i < nd; ++i) dist.at<float>(i, 0) = (float)std::atof(argv[18 + i]);
        const int window = std::atoi(argv[23]), check = std::atoi(argv[24]);
        ORB_SLAM3::ORBextractor ex(nfeatures, scale, nlevels, ini, mn);
        std::vector<cv::KeyPoint> keys[2], un[2];
        cv::Mat desc[2];
        static ORB_SLAM3::FrameGridCells grid[2];
        OrbxFrameCalib calib;
        std::vector<int> lap = {lap0, lap1};
        for (int k = 0; k < 2; ++k) {
            cv::Mat im(h, w, CV_8UC1, frames.data() + (size_t)k * w * h);
            if (k == 0) {
                // first frame: the three separate calls
                if (ex(im, cv::Mat(), keys[k], desc[k], lap) < 0) { std::fprintf(stderr, "frame: %s\n", ex.LastError().c_str()); return 8; }
                desc[k] = desc[k].clone();
                if (!ORB_SLAM3::ComputeImageBounds(ex, K, dist, w, h, calib)) return 9;
                if (ORB_SLAM3::UndistortAndAssignToGrid(ex, calib, keys[k], un[k], grid[k]) < 0) return 10;
            } else {
                // second frame: extraction + undistort + grid in one call (keypoints stay on the GPU in between)
                if (ORB_SLAM3::ExtractFrame(ex, calib, im, lap0, lap1, keys[k], desc[k], un[k], grid[k]) < 0) return 12;
            }
        }
        std::vector<cv::Point2f> prev(un[0].size());
        for (size_t i = 0; i < un[0].size(); ++i) prev[i] = un[0][i].pt;
        std::vector<int> m12;
        const int nm = ORB_SLAM3::SearchForInitialization(ex, calib, un[0], desc[0], un[1], desc[1], grid[1], prev, m12, window, 0.9f, check != 0);
        if (nm < 0) { std::fprintf(stderr, "frame: %s\n", ex.LastError().c_str()); return 11; }
        std::fclose(out);
        out = std::fopen(argv[3], "wb");
        int32_t oh[3] = {0x554f5246, (int32_t)un[0].size(), (int32_t)un[1].size()};
        std::fwrite(oh, 4, 3, out);
        float b[4] = {calib.min_x, calib.max_x, calib.min_y, calib.max_y};
        std::fwrite(b, 4, 4, out);
        for (int k = 0; k < 2; ++k) std::fwrite(un[k].data(), sizeof(cv::KeyPoint), un[k].size(), out);
        for (int k = 0; k < 2; ++k) {
            std::vector<int32_t> start(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, 0), items;
            for (int ix = 0; ix < FRAME_GRID_COLS; ++ix)
                for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) {
                    for (size_t j = 0; j < grid[k][ix][iy].size(); ++j) items.push_back((int32_t)grid[k][ix][iy][j]);
                    start[ix * FRAME_GRID_ROWS + iy + 1] = (int32_t)items.size();
                }
            std::fwrite(start.data(), 4, start.size(), out);
            std::fwrite(items.data(), 4, items.size(), out);
        }
        int32_t nm32 = nm;
        std::fwrite(&nm32, 4, 1, out);
        std::vector<int32_t> m(m12.begin(), m12.end());
        std::fwrite(m.data(), 4, m.size(), out);
        std::fwrite(prev.data(), 8, prev.size(), out);
        std::fclose(out);
        return 0;
    }

    ORBextractor ex(nfeatures, scale, nlevels, ini, mn);          // global alias of ORB_SLAM3::ORBextractor
    if (ex.GetLevels() != nlevels || (int)ex.GetScaleFactors().size() != nlevels || ex.GetScaleFactor() != scale) return 3;
    for (int i = 0; i < n; ++i) {
        cv::Mat img(h, w, CV_8UC1, frames.data() + (size_t)i * w * h);
        std::vector<cv::KeyPoint> kps;
        cv::Mat desc;
        std::vector<int> lap = {lap0, lap1};
        std::vector<std::vector<cv::KeyPoint> > lvl;
        int ret;
        if (!std::strcmp(mode, "five")) {
            // the north-star signature (reference inc/ORBExtractor.h:55-56) + the demo's direct stage calls
            ret = ex(img, cv::Mat(), kps, desc, lap);
            ORBextractor ex2(nfeatures, scale, nlevels, ini, mn);
            ex2.ComputePyramid(img);
            ex2.ComputeKeyPointsOctTree(lvl);
        } else if (!std::strcmp(mode, "threads")) {
            // stereo pattern: two instances, two host threads, same image -> must agree with each other
            ORB_SLAM3::ORBextractor right(nfeatures, scale, nlevels, ini, mn);
            std::vector<cv::KeyPoint> kps_r; cv::Mat desc_r; std::vector<std::vector<cv::KeyPoint> > lvl_r;
            std::vector<int> lap_r = lap;
            int ret_r = -7;
            std::thread tl([&]() { ret = ex(img, cv::Mat(), kps, desc, lap, lvl); });
            std::thread tr([&]() { ret_r = right(img, cv::Mat(), kps_r, desc_r, lap_r, lvl_r); });
            tl.join(); tr.join();
            if (ret != ret_r || kps.size() != kps_r.size() ||
                (kps.size() && std::memcmp(kps.data(), kps_r.data(), kps.size() * sizeof(cv::KeyPoint)))) return 4;
        } else {
            ret = ex(img, cv::Mat(), kps, desc, lap, lvl);
        }
        if (ret < 0 && !ex.LastError().empty()) std::fprintf(stderr, "extractor: %s\n", ex.LastError().c_str());
        int32_t cnt = (int32_t)kps.size();
        std::fwrite(&ret, 4, 1, out);
        std::fwrite(&cnt, 4, 1, out);
        for (int l = 0; l < nlevels; ++l) {
            int32_t c = l < (int)lvl.size() ? (int32_t)lvl[l].size() : 0;
            std::fwrite(&c, 4, 1, out);
        }
        if (cnt) {
            std::fwrite(kps.data(), sizeof(cv::KeyPoint), (size_t)cnt, out);
            if (desc.rows != cnt || desc.cols != 32) return 5;
            for (int r = 0; r < cnt; ++r) std::fwrite(desc.ptr(r), 1, 32, out);
        } else if (!desc.empty()) return 6;
        for (size_t l = 0; l < lvl.size(); ++l)
            if (!lvl[l].empty()) std::fwrite(lvl[l].data(), sizeof(cv::KeyPoint), lvl[l].size(), out);
        if (dump) {
            for (int l = 0; l < nlevels; ++l) {
                const cv::Mat& m = ex.mvImagePyramid[l];           // ROI at (19,19) of the bordered buffer
                int32_t wh[2] = {m.cols, m.rows};
                std::fwrite(wh, 4, 2, out);
                const uint8_t* base = m.data - 19 * m.step - 19;
                for (int r = 0; r < m.rows + 38; ++r) std::fwrite(base + (size_t)r * m.step, 1, (size_t)m.cols + 38, out);
            }
        }
    }
    std::fclose(out);
    return 0;
}
