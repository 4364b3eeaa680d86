// oracle/ref_main.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Driver around the UNMODIFIED reference extractor (/root/reference/src/orb_extractor/ORBextractor.cc,
// compiled where it lies by oracle/Makefile against oracle/shim).  Two uses:
//   * `run`   : golden generation.  Built with -DORBX_BUMP_ALLOC the global operator new is a
//               monotonic bump allocator, which makes the reference's pointer tie-break in
//               DistributeOctTree (ORBextractor.cc:689, sort of pair<int,ExtractorNode*>) deterministic:
//               "equal size => later-created node first" (SURVEY.md section 8(a) A3-detail.7).
//   * `bench` : CPU baseline timing with the normal allocator, one extractor + one frame per thread.
//
// Frame file  : int32 'ORBF', nframes, w, h ; nframes*h*w bytes.
// Result file : int32 'ORBR', nframes, nlevels, dump_pyr ; per frame:
//               int32 ret, n, counts[nlevels]; n*28 B keypoints; n*32 B descriptors;
//               sum(counts)*28 B level keypoints (level coordinates, ORBextractor.cc:1094);
//               if dump_pyr: per level int32 w, h then (h+38)*(w+38) bytes (bordered plane).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "ORBextractor.h"

#ifdef ORBX_BUMP_ALLOC
#include <new>
#include <sys/mman.h>
static char* g_arena = nullptr;
static size_t g_arena_size = (size_t)48 << 30;
static size_t g_arena_off = 0;
static void* bump_alloc(size_t n) {
    if (!g_arena) {
        g_arena = (char*)mmap(nullptr, g_arena_size, PROT_READ | PROT_WRITE,
                              MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (g_arena == (char*)MAP_FAILED) { std::fprintf(stderr, "bump arena mmap failed\n"); std::abort(); }
    }
    size_t off = (g_arena_off + 15) & ~(size_t)15;
    if (off + n > g_arena_size) { std::fprintf(stderr, "bump arena exhausted\n"); std::abort(); }
    g_arena_off = off + n;
    return g_arena + off;
}
void* operator new(size_t n) { return bump_alloc(n); }
void* operator new[](size_t n) { return bump_alloc(n); }
void operator delete(void*) noexcept {}
void operator delete[](void*) noexcept {}
void operator delete(void*, size_t) noexcept {}
void operator delete[](void*, size_t) noexcept {}
static size_t bump_mark() { return g_arena_off; }
static void bump_reset(size_t m) { g_arena_off = m; }
#else
static size_t bump_mark() { return 0; }
static void bump_reset(size_t) {}
#endif

struct Frames {
    int n = 0, w = 0, h = 0;
    uint8_t* data = nullptr;
};

static bool load_frames(const char* path, Frames& f) {
    FILE* fp = std::fopen(path, "rb");
    if (!fp) return false;
    int32_t hdr[4];
    if (std::fread(hdr, 4, 4, fp) != 4 || hdr[0] != 0x4642524f) { std::fclose(fp); return false; }
    f.n = hdr[1]; f.w = hdr[2]; f.h = hdr[3];
    size_t total = (size_t)f.n * f.w * f.h;
    f.data = (uint8_t*)std::malloc(total);
    bool ok = std::fread(f.data, 1, total, fp) == total;
    std::fclose(fp);
    return ok;
}

struct Cfg {
    int nfeatures = 1000; float scale = 1.2f; int nlevels = 8; int ini = 20; int mn = 7; int lap0 = 0; int lap1 = 0;
};

static Cfg parse_cfg(char** a) {
    Cfg c;
    c.nfeatures = std::atoi(a[0]); c.scale = (float)std::atof(a[1]); c.nlevels = std::atoi(a[2]);
    c.ini = std::atoi(a[3]); c.mn = std::atoi(a[4]); c.lap0 = std::atoi(a[5]); c.lap1 = std::atoi(a[6]);
    return c;
}

static int cmd_run(int argc, char** argv) {
    if (argc < 11) return 2;
    Frames fr;
    if (!load_frames(argv[2], fr)) { std::fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    Cfg c = parse_cfg(argv + 4);
    const int dump = argc > 11 ? std::atoi(argv[11]) : 0;
    FILE* out = std::fopen(argv[3], "wb");
    if (!out) return 1;
    int32_t hdr[4] = {0x5242524f, fr.n, c.nlevels, dump};
    std::fwrite(hdr, 4, 4, out);
    const size_t mark = bump_mark();
    for (int i = 0; i < fr.n; ++i) {
        bump_reset(mark);
        ORB_SLAM3::ORBextractor ex(c.nfeatures, c.scale, c.nlevels, c.ini, c.mn);
        cv::Mat img(fr.h, fr.w, CV_8UC1, fr.data + (size_t)i * fr.w * fr.h);
        std::vector<cv::KeyPoint> kps;
        cv::Mat desc;
        std::vector<int> lap = {c.lap0, c.lap1};
        std::vector<std::vector<cv::KeyPoint> > lvl;
        int ret = ex(img, cv::Mat(), kps, desc, lap, lvl);
        int32_t n = (int32_t)kps.size();
        std::fwrite(&ret, 4, 1, out);
        std::fwrite(&n, 4, 1, out);
        for (int l = 0; l < c.nlevels; ++l) {
            int32_t cnt = l < (int)lvl.size() ? (int32_t)lvl[l].size() : 0;
            std::fwrite(&cnt, 4, 1, out);
        }
        if (n) {
            std::fwrite(kps.data(), sizeof(cv::KeyPoint), (size_t)n, out);
            for (int r = 0; r < n; ++r) std::fwrite(desc.ptr(r), 1, 32, out);
        }
        for (size_t l = 0; l < lvl.size(); ++l)
            if (!lvl[l].empty()) std::fwrite(lvl[l].data(), sizeof(cv::KeyPoint), lvl[l].size(), out);
        if (dump) {
            for (int l = 0; l < c.nlevels; ++l) {
                const cv::Mat& m = ex.mvImagePyramid[l];
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

// bench <frames> <threads> <seconds_budget> <cfg x7>: every thread owns one extractor and walks the
// frame set round-robin (thread t takes frames t, t+T, ...) repeatedly until the wall budget is spent.
static int cmd_bench(int argc, char** argv) {
    if (argc < 12) return 2;
    Frames fr;
    if (!load_frames(argv[2], fr)) { std::fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    const int T = std::max(1, std::atoi(argv[3]));
    const double budget = std::atof(argv[4]);
    Cfg c = parse_cfg(argv + 5);
    std::vector<std::vector<double> > lat(T);
    std::vector<long> kpsum(T, 0);
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
            ORB_SLAM3::ORBextractor ex(c.nfeatures, c.scale, c.nlevels, c.ini, c.mn);
            std::vector<int> lap = {c.lap0, c.lap1};
            int i = t % fr.n;
            for (;;) {
                auto a = std::chrono::steady_clock::now();
                if (std::chrono::duration<double>(a - t0).count() >= budget && !lat[t].empty()) break;
                cv::Mat img(fr.h, fr.w, CV_8UC1, fr.data + (size_t)i * fr.w * fr.h);
                std::vector<cv::KeyPoint> kps;
                cv::Mat desc;
                std::vector<std::vector<cv::KeyPoint> > lvl;
                ex(img, cv::Mat(), kps, desc, lap, lvl);
                auto b = std::chrono::steady_clock::now();
                lat[t].push_back(std::chrono::duration<double, std::milli>(b - a).count());
                kpsum[t] += (long)kps.size();
                i = (i + T) % fr.n;
            }
        });
    for (auto& x : th) x.join();
    double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<double> all;
    long kp = 0;
    for (int t = 0; t < T; ++t) { all.insert(all.end(), lat[t].begin(), lat[t].end()); kp += kpsum[t]; }
    std::sort(all.begin(), all.end());
    double p50 = all[all.size() / 2], p99 = all[std::min(all.size() - 1, (size_t)(all.size() * 0.99))];
    std::printf("{\"frames\": %zu, \"wall_s\": %.6f, \"fps\": %.3f, \"threads\": %d, \"p50_ms\": %.4f, \"p99_ms\": %.4f, "
                "\"mean_keypoints\": %.2f}\n",
                all.size(), wall, all.size() / wall, T, p50, p99, (double)kp / all.size());
    return 0;
}

// tables nfeatures scale nlevels ini min: the constructor tables of the unmodified class (ORBextractor.cc:419-474) through
// its own accessors (inc/ORBextractor.h:63-83) and public members, one line of hex-exact numbers per table.
static int cmd_tables(int argc, char** argv) {
    if (argc < 7) return 2;
    ORB_SLAM3::ORBextractor ex(std::atoi(argv[2]), (float)std::atof(argv[3]), std::atoi(argv[4]), std::atoi(argv[5]), std::atoi(argv[6]));
    auto pf = [](const char* name, const std::vector<float>& v) {
        std::printf("%s", name);
        for (float f : v) { uint32_t u; std::memcpy(&u, &f, 4); std::printf(" %08x", u); }
        std::printf("\n");
    };
    auto pi = [](const char* name, const std::vector<int>& v) {
        std::printf("%s", name);
        for (int i : v) std::printf(" %d", i);
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

int main(int argc, char** argv) {
    if (argc >= 2 && !std::strcmp(argv[1], "tables")) return cmd_tables(argc, argv);
    if (argc >= 2 && !std::strcmp(argv[1], "run")) return cmd_run(argc, argv);
    if (argc >= 2 && !std::strcmp(argv[1], "bench")) return cmd_bench(argc, argv);
    std::fprintf(stderr,
                 "usage: %s run <frames.orbf> <out.orbr> nfeatures scale nlevels ini min lap0 lap1 [dump_pyr]\n"
                 "       %s bench <frames.orbf> threads seconds nfeatures scale nlevels ini min lap0 lap1\n",
                 argv[0], argv[0]);
    return 2;
}
