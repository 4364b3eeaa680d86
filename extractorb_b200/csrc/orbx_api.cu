// extractorb_b200/csrc/orbx_api.cu -- host side of libextractorb_cuda.so: constructor tables, per-size
// plan, workspace in HBM, launch orchestration and the extern "C" boundary declared in include/orbx.h.
//
// Reference being replaced: ORB_SLAM3::ORBextractor (/root/reference/src/orb_extractor/ORBextractor.cc,
// twin ORBExtractor.cpp).  Line citations below refer to ORBextractor.cc.
#include <cuda.h>               // CUtensorMap types only: cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost a pointer test when no profiler is attached

#include <cfloat>
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <utility>
#include <vector>

#include "../../include/orbx.h"
#include "orbx_kernels.cuh"
#include "orbx_frame.cuh"
#include "orbx_clahe.cuh"

static const int8_t kPatternHost[1024] = {
#include "orb_pattern.inc"
};

namespace {

inline int cv_round_f(float v) { return (int)lrintf(v); }      // cvRound: round half to even
inline int cv_round_d(double v) { return (int)lrint(v); }
inline int cv_floor_f(float v) { int i = (int)v; return i - (i > v); }
inline int cv_ceil_f(float v) { int i = (int)v; return i + (i < v); }
inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

struct PlanEntry {
    OrbxPlan plan;
    std::vector<int2> xtab, ytab;
    std::vector<OrbxCell> cells;
    std::vector<OrbxFastTile> tiles;
    OrbxFastTile* d_tiles = nullptr;
    uint8_t* d_slot_level = nullptr;          // level of every kept-keypoint slot (kp_total bytes), k_describe
    size_t ft_smem = 0;
    int2* d_xtab = nullptr;
    int2* d_ytab = nullptr;
    OrbxCell* d_cells = nullptr;
    uint32_t* d_blur_tiles = nullptr;
    long long pyr_stride = 0, blur_stride = 0, cand_stride = 0;
    size_t fast_smem = 0, qt_smem = 0;
    bool qt_global = false;                 // quadtree node tables in HBM (ws.qt_scratch) instead of shared memory
    int blur_tiles = 0;
    int rs_rows[ORBX_MAX_LEVELS] = {0};     // resize kernel: staged source rows / row pitch (bytes) per level
    int rs_pitch[ORBX_MAX_LEVELS] = {0};
    int rs_pairs[ORBX_MAX_LEVELS] = {0};    // 1: source columns advance by <= 2 per destination column (two-column horizontal pass)
    int rs_launch_pitch[ORBX_MAX_LEVELS] = {0};  // staging pitch actually used at launch (ORBX_RS_PITCH for the specialised instance)
    // CUDA graph of the whole launch sequence for small launch groups (latency path), keyed by its arguments
    struct GraphKey {
        const void* imgs; long long rs, fs; int nf, lap0, lap1; void* kps; void* desc; int cap; void* counts; int fo, stages;
        bool border;          // the graph includes k_pyr_border (one or two frames, or a pyramid sink is set)
        bool operator==(const GraphKey& o) const {
            return imgs == o.imgs && rs == o.rs && fs == o.fs && nf == o.nf && lap0 == o.lap0 && lap1 == o.lap1 && kps == o.kps &&
                   desc == o.desc && cap == o.cap && counts == o.counts && fo == o.fo && stages == o.stages && border == o.border;
        }
    };
    struct GraphSlot { GraphKey key; cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
    GraphSlot graphs[4];
    int graph_next = 0;
};

struct StageEvents { cudaEvent_t ev[ORBX_NUM_STAGES + 1]; };

}  // namespace

struct OrbxHandle {
    OrbxParams prm;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // constructor tables (:408-475)
    std::vector<float> sf, inv_sf, sigma2, inv_sigma2;
    std::vector<int> quota;
    int umax[16];
    OrbxFloatConsts fc;
    int cand_per_cell = 64;
    int n_slots = 4;               // staging slots of the host-buffer pipeline (ORBX_SLOTS=2..4)
    bool ramp = true;              // ramp the launch-group size of host-buffer calls up / down (ORBX_RAMP=0 disables)
    bool fast_tma = true;          // ORBX_FAST_TMA=0: stage the FAST tiles with cp.async instead of one TMA bulk tensor copy
    bool fast_v1 = false;          // ORBX_FAST_V1=1: the round-1 warp-per-cell FAST kernel (A/B measurements only)
    // plans keyed by image size
    std::map<std::pair<int, int>, PlanEntry*> plans;
    PlanEntry* cur = nullptr;      // plan of the resident frames
    int resident_frames = 0;
    // workspace (sized for ws_frames frames of ws_plan)
    PlanEntry* ws_plan = nullptr;
    int ws_frames = 0;
    OrbxWs ws{};                   // workspace set 0 (also the view the accessors read unless res_set == 1)
    OrbxWs ws2{};                  // workspace set 1: consecutive launch groups alternate sets and compute streams
    int ws2_frames = 0;
    int res_set = 0;               // which set holds the resident (last) group
    bool borders_valid[2] = {false, false};   // the resident planes of each set have their 19-px border (k_pyr_border ran)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_s2 = nullptr;
    cudaStream_t s_side = nullptr;      // side branch of captured graphs: border + blur run beside FAST + quadtree
    // per-level branches of the single-frame graph: FAST + quadtree of level l start as soon as level l exists
    cudaStream_t s_lvl[ORBX_MAX_LEVELS] = {};
    cudaEvent_t ev_lvl_ready[ORBX_MAX_LEVELS] = {}, ev_lvl_done[ORBX_MAX_LEVELS] = {};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    float* d_pattern_f = nullptr;
    int2* d_angle_w = nullptr;
    uint8_t* d_in = nullptr; size_t d_in_bytes = 0;        // input staging (two slots when pipelining host frames)
    void* d_out = nullptr; size_t d_out_bytes = 0;          // output staging (two slots)
    cudaStream_t s_in = nullptr, s_out = nullptr;          // copy streams of the host-buffer pipeline
    int* h_flag = nullptr;                                  // pinned copy of the overflow flag word
    uint8_t* pyr_out = nullptr; size_t pyr_out_stride = 0;  // orbx_set_pyramid_output: host sink of every frame's bordered pyramid
    cudaEvent_t ev_pyr_ready[2] = {nullptr, nullptr}, ev_pyr_done[2] = {nullptr, nullptr};
    bool pyr_pending[2] = {false, false};
    uint8_t* d_stereo = nullptr; size_t d_stereo_bytes = 0;  // scratch of orbx_stereo_match / orbx_frame_* / orbx_search_for_initialization
    uint8_t* d_frame = nullptr; size_t d_frame_bytes = 0;    // device-resident outputs of orbx_extract_frame
    uint8_t* h_out1 = nullptr; size_t h_out1_bytes = 0;      // pinned read-back buffer of single-frame host calls
    uint8_t* h_frame = nullptr; size_t h_frame_bytes = 0;    // pinned read-back buffer of orbx_extract_frame
    uint8_t* d_clahe = nullptr; size_t d_clahe_bytes = 0;    // LUTs + staging of orbx_clahe
    int last_init_fallbacks = 0;                             // filtered re-enumerations of the last SearchForInitialization (diagnostic)
    // host-buffer pipeline: ORBX_IN_SLOTS input staging slots (the copy engine runs ahead of the two compute
    // streams), two output staging slots
    cudaEvent_t ev_h2d[4] = {nullptr, nullptr, nullptr, nullptr}, ev_in_free[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_done[4] = {nullptr, nullptr, nullptr, nullptr}, ev_d2h[4] = {nullptr, nullptr, nullptr, nullptr};
    // profiling
    std::vector<StageEvents> events;
    size_t events_used = 0;
    double stage_ms[ORBX_NUM_STAGES] = {0, 0, 0, 0, 0};
    int64_t stage_launches = 0;
    int64_t total_launches = 0;
};

namespace {

int fail(OrbxHandle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

// number of leading pyramid levels whose quadtrees run with ORBX_QT_THREADS_BIG threads
inline int qt_big_levels(const OrbxPlan& P, long long min_pixels = ORBX_QT_BIG_PIXELS) {
    int nbig = 0;
    while (nbig < P.nlevels && (long long)P.lv[nbig].w * P.lv[nbig].h >= min_pixels) ++nbig;
    return nbig;
}

#define ORBX_CUDA(call)                                                                              \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            char buf_[512];                                                                          \
            snprintf(buf_, sizeof(buf_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return fail(h, ORBX_ERR_CUDA, buf_);                                                     \
        }                                                                                            \
    } while (0)

// ORBextractor::ORBextractor, :408-475.  Float/double mix reproduced literally.  Pure host arithmetic: no CUDA call, so the
// tables exist (orbx_ctor_tables) even where no device does.
struct CtorTables {
    std::vector<float> sf, inv_sf, sigma2, inv_sigma2;
    std::vector<int> quota;
    int umax[16];
};

void compute_ctor_tables(int nfeatures, float scale_factor, int L, CtorTables& t) {
    const double scaleFactor = (double)scale_factor;  // double member initialised from float (inc/ORBextractor.h:98)
    t.sf.assign(L, 1.f); t.inv_sf.assign(L, 1.f); t.sigma2.assign(L, 1.f); t.inv_sigma2.assign(L, 1.f);
    for (int i = 1; i < L; ++i) {
        t.sf[i] = (float)(t.sf[i - 1] * scaleFactor);
        t.sigma2[i] = t.sf[i] * t.sf[i];
    }
    for (int i = 0; i < L; ++i) {
        t.inv_sf[i] = 1.0f / t.sf[i];
        t.inv_sigma2[i] = 1.0f / t.sigma2[i];
    }
    t.quota.assign(L, 0);
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)L));
    int sum = 0;
    for (int l = 0; l < L - 1; ++l) {
        t.quota[l] = cv_round_f(nDesired);
        sum += t.quota[l];
        nDesired *= factor;
    }
    if (L > 0) t.quota[L - 1] = std::max(nfeatures - sum, 0);
    // umax, :459-474
    int v, v0;
    const int vmax = cv_floor_f(ORBX_HALF_PATCH * sqrtf(2.f) / 2 + 1);
    const int vmin = cv_ceil_f(ORBX_HALF_PATCH * sqrtf(2.f) / 2);
    const double hp2 = ORBX_HALF_PATCH * ORBX_HALF_PATCH;
    for (v = 0; v < 16; ++v) t.umax[v] = 0;
    for (v = 0; v <= vmax; ++v) t.umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = ORBX_HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (t.umax[v0] == t.umax[v0 + 1]) ++v0;
        t.umax[v] = v0;
        ++v0;
    }
}

void build_ctor_tables(OrbxHandle* h) {
    CtorTables t;
    compute_ctor_tables(h->prm.nfeatures, h->prm.scale_factor, h->prm.nlevels, t);
    h->sf = t.sf; h->inv_sf = t.inv_sf; h->sigma2 = t.sigma2; h->inv_sigma2 = t.inv_sigma2; h->quota = t.quota;
    for (int v = 0; v < 16; ++v) h->umax[v] = t.umax[v];
    // cv::fastAtan2 constants (float products, as OpenCV computes them) and factorPI (:104)
    const float s = (float)(180.0 / 3.1415926535897932384626433832795);
    h->fc.atan_p1 = 0.9997878412794807f * s;
    h->fc.atan_p3 = -0.3258083974640975f * s;
    h->fc.atan_p5 = 0.1555786518463281f * s;
    h->fc.atan_p7 = -0.04432655554792128f * s;
    h->fc.atan_eps = (float)DBL_EPSILON;
    h->fc.deg2rad = (float)(3.1415926535897932384626433832795 / 180.f);
}

// cv::resize INTER_LINEAR coefficient tables (per axis), OpenCV resize.cpp model.
void linear_axis_table(int ssize, int dsize, std::vector<int2>& out) {
    const double inv_scale = (double)dsize / ssize;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = cv_floor_f(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        const int a0 = cv_round_f((1.f - f) * 2048.f), a1 = cv_round_f(f * 2048.f);
        out.push_back(make_int2(s, (a0 & 0xffff) | (a1 << 16)));
    }
}

void free_plan(PlanEntry* p) {
    if (!p) return;
    for (auto& g : p->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    cudaFree(p->d_xtab); cudaFree(p->d_ytab); cudaFree(p->d_cells); cudaFree(p->d_tiles); cudaFree(p->d_slot_level); cudaFree(p->d_blur_tiles);
    delete p;
}

int build_plan(OrbxHandle* h, int width, int height, PlanEntry** out) {
    const int L = h->prm.nlevels;
    const float W = (float)(h->prm.cell_size > 0 ? h->prm.cell_size : 30);   // :777
    PlanEntry* pe = new PlanEntry();
    OrbxPlan& P = pe->plan;
    memset(&P, 0, sizeof(P));
    P.nlevels = L; P.width = width; P.height = height;
    P.ini_th = std::min(std::max(h->prm.ini_th_fast, 0), 255);   // cv::FAST clamps the threshold
    P.min_th = std::min(std::max(h->prm.min_th_fast, 0), 255);
    for (int i = 0; i < 16; ++i) P.umax[i] = h->umax[i];
    long long plane_off = 0, blur_off = 0, cand_off = 0;
    int kp_off = 0, maxcw = 7, maxch = 7, qt_nc = 8, ft_maxtw = 7, ft_maxth = 7, ft_qcap = 1, ft_scap = 1;
    for (int l = 0; l < L; ++l) {
        OrbxLevel& V = P.lv[l];
        V.w = cv_round_f((float)width * h->inv_sf[l]);     // :1171
        V.h = cv_round_f((float)height * h->inv_sf[l]);
        if (V.w > ORBX_MAX_LEVEL_DIM || V.h > ORBX_MAX_LEVEL_DIM) { delete pe; return fail(h, ORBX_ERR_IMAGE_TOO_LARGE, "level larger than 4127 px"); }
        if (V.w <= 2 * ORBX_EDGE || V.h <= 2 * ORBX_EDGE) { delete pe; return fail(h, ORBX_ERR_LEVEL_TOO_SMALL, "pyramid level smaller than the 19-px border"); }
        V.pitch = (int)align_up(ORBX_PADL + V.w + ORBX_EDGE, 64);
        V.plane_rows = V.h + 2 * ORBX_EDGE;
        V.plane_off = plane_off;
        plane_off += align_up((long long)V.pitch * V.plane_rows, 256);
        V.blur_pitch = (int)align_up(V.w, 16);
        V.blur_off = blur_off;
        blur_off += align_up((long long)V.blur_pitch * V.h, 256);
        if (l > 0) {
            const OrbxLevel& S = P.lv[l - 1];
            V.xtab_off = (int)pe->xtab.size();
            V.ytab_off = (int)pe->ytab.size();
            if (S.w == 2 * V.w && S.h == 2 * V.h) {       // OpenCV: exact 2x decimation -> INTER_AREA fast path
                for (int x = 0; x < V.w; ++x) pe->xtab.push_back(make_int2(2 * x, -1));
                for (int y = 0; y < V.h; ++y) pe->ytab.push_back(make_int2(2 * y, 0));
            } else {
                linear_axis_table(S.w, V.w, pe->xtab);
                linear_axis_table(S.h, V.h, pe->ytab);
            }
            // shared-memory window of the resize kernel: exact maxima over its 128x64 tiles (the 16-row tiles of the latency
            // instance are subsets of them)
            int rows = 1, cols = 1;
            for (int y0 = 0; y0 < V.h; y0 += ORBX_RS_TH) {
                const int yl = std::min(y0 + ORBX_RS_TH, V.h) - 1;
                rows = std::max(rows, std::min(pe->ytab[V.ytab_off + yl].x + 1, S.h - 1) - pe->ytab[V.ytab_off + y0].x + 1);
            }
            for (int x0 = 0; x0 < V.w; x0 += ORBX_RS_TW) {
                const int xl = std::min(x0 + ORBX_RS_TW, V.w) - 1;
                cols = std::max(cols, std::min(pe->xtab[V.xtab_off + xl].x + 1, S.w - 1) - pe->xtab[V.xtab_off + x0].x + 1);
            }
            pe->rs_rows[l] = rows;
            pe->rs_pairs[l] = 1;
            for (int x = 0; x + 1 < V.w; ++x) {
                const int dxs = pe->xtab[V.xtab_off + x + 1].x - pe->xtab[V.xtab_off + x].x;
                if (dxs < 0 || dxs > 2) pe->rs_pairs[l] = 0;
            }
            {
                const bool area_l = pe->xtab[V.xtab_off].y == -1;
                const int pp = (int)align_up(cols + 15 + 15, 16);
                pe->rs_launch_pitch[l] = (!area_l && pp <= ORBX_RS_PITCH) ? ORBX_RS_PITCH : pp;
            }
            pe->rs_pitch[l] = (int)align_up(cols + 15 + 15, 16);     // 16-byte aligned window start + whole 16-byte vectors
        }
        // FAST cell grid, :781-814
        const int minBX = ORBX_FAST_BORDER, minBY = ORBX_FAST_BORDER;
        const int maxBX = V.w - ORBX_FAST_BORDER, maxBY = V.h - ORBX_FAST_BORDER;
        const float fwidth = (float)(maxBX - minBX), fheight = (float)(maxBY - minBY);
        V.nCols = (int)(fwidth / W); V.nRows = (int)(fheight / W);
        if (V.nCols < 1 || V.nRows < 1) { delete pe; return fail(h, ORBX_ERR_LEVEL_TOO_SMALL, "pyramid level narrower than one FAST cell (division by zero in the reference, ORBextractor.cc:794)"); }
        V.wCell = (int)ceilf(fwidth / V.nCols); V.hCell = (int)ceilf(fheight / V.nRows);
        V.cell_off = (int)pe->cells.size();
        V.tile_off = (int)pe->tiles.size();
        for (int i = 0; i < V.nRows; ++i) {
            const float iniY = (float)(minBY + i * V.hCell);
            float maxY = iniY + V.hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = (float)maxBY;
            const size_t row_first = pe->cells.size();
            for (int j = 0; j < V.nCols; ++j) {
                const float iniX = (float)(minBX + j * V.wCell);
                float maxX = iniX + V.wCell + 6;
                if (iniX >= maxBX - 6) continue;          // -6 here vs -3 for rows: reference quirk, kept
                if (maxX > maxBX) maxX = (float)maxBX;
                const int cw = (int)maxX - (int)iniX, ch = (int)maxY - (int)iniY;
                if (cw < 7 || ch < 7) continue;           // cv::FAST finds nothing in images < 7 px
                if (cw > ORBX_MAX_CELL_DIM || ch > ORBX_MAX_CELL_DIM) { delete pe; return fail(h, ORBX_ERR_BAD_ARGUMENT, "cell_size too large (cell image > 127 px)"); }
                OrbxCell c;
                c.x0 = (uint16_t)(int)iniX; c.y0 = (uint16_t)(int)iniY; c.cw = (uint8_t)cw; c.ch = (uint8_t)ch;
                c.level = (uint8_t)l; c.pad = 0;
                c.ordbase = (uint32_t)(i * V.nCols + j) << ORBX_ORD_CELL_SHIFT;
                c.xoff = (uint16_t)(j * V.wCell); c.yoff = (uint16_t)(i * V.hCell);
                {
                    const int al = (ORBX_PADL + (int)iniX) & 3;
                    const int nwords = (al + cw + 3) >> 2;
                    const int wq0 = (al + 3) >> 2;
                    const int ngrp = ((al + cw - 4) >> 2) - wq0 + 1;
                    c.nwords = (uint8_t)nwords; c.ngrp = (uint8_t)ngrp; c.wq0 = (uint8_t)wq0;
                    c.wmagic = (1u << 20) / (unsigned)nwords + 1u;
                    c.gmagic = (1u << 24) / (unsigned)ngrp + 1u;
                    const unsigned first_mask = (0xFu << ((al + 3) & 3)) & 0xFu;        // interior starts at byte al+3
                    const unsigned last_mask = 0xFu >> (3 - ((al + cw - 4) & 3));        // interior ends at byte al+cw-4
                    c.masks = (uint8_t)(first_mask | (last_mask << 4));
                    c.reserved = 0;
                }
                if ((long long)V.nRows * V.nCols >= (1 << (32 - ORBX_ORD_CELL_SHIFT))) { delete pe; return fail(h, ORBX_ERR_IMAGE_TOO_LARGE, "too many cells"); }
                pe->cells.push_back(c);
                maxcw = std::max(maxcw, cw); maxch = std::max(maxch, ch);
            }
            // k_fast_tiles: the row's cells (consecutive columns from 0: cells are only ever skipped at the right end) in
            // runs of at most ORBX_FT_MAXC, balanced
            const int nv = (int)(pe->cells.size() - row_first);
            // a tile row (alignment slack + cells + 6 halo columns + the word right of the last pair) fits the 256-byte TMA box
            const int maxc = std::max(1, std::min(ORBX_FT_MAXC, (256 - 15 - 6 - 8) / V.wCell));
            const int ntr = (nv + maxc - 1) / maxc;
            for (int t = 0, j0 = 0; t < ntr; ++t) {
                const int nc = nv / ntr + (t < nv % ntr ? 1 : 0);
                const OrbxCell& a = pe->cells[row_first + j0];
                const OrbxCell& z = pe->cells[row_first + j0 + nc - 1];
                OrbxFastTile T;
                memset(&T, 0, sizeof(T));
                T.x0 = a.x0; T.y0 = a.y0; T.tw = (uint16_t)(z.x0 + z.cw - a.x0); T.th = a.ch;
                T.level = (uint8_t)l; T.ncells = (uint8_t)nc; T.wcell = (uint16_t)V.wCell;
                T.ordbase = a.ordbase; T.xoff = a.xoff; T.yoff = a.yoff;
                T.cmagic = 0xffffffffu / (unsigned)V.wCell + 1u;
                const int a16 = (ORBX_PADL + T.x0) & 15;
                const int lo = a16 + 3, hi = a16 + T.tw - 4;                  // first / last interior byte of a tile row
                const int pq0 = lo >> 3, pq1 = hi >> 3, npairs = pq1 - pq0 + 1;
                T.pq0 = (uint8_t)pq0; T.npairs = (uint8_t)npairs;
                T.pmagic = 0xffffffffu / (unsigned)npairs + 1u;
                T.nitems = (uint16_t)(npairs * (T.th - 6));
                auto spread8 = [](unsigned m) {                               // pixel k of a pair -> bit 8k+7 (k < 4), 8(k-4)+3 (k >= 4)
                    unsigned sp = 0;
                    for (int k = 0; k < 8; ++k) sp |= ((m >> k) & 1u) << (k < 4 ? 8 * k + 7 : 8 * (k - 4) + 3);
                    return sp;
                };
                T.vmagic = 0xffffffffu / (unsigned)((a16 + T.tw + 15) >> 4) + 1u;
                T.first_mask = spread8((0xffu << (lo - 8 * pq0)) & 0xffu);
                T.last_mask = spread8(0xffu >> (8 * pq1 + 7 - hi));
                for (int c = 0; c < nc; ++c) {                                // the layout k_fast_tiles relies on
                    const OrbxCell& q = pe->cells[row_first + j0 + c];
                    if (q.x0 != a.x0 + c * V.wCell || q.y0 != a.y0 || q.ch != a.ch || (c < nc - 1 && q.cw != V.wCell + 6) || q.cw > V.wCell + 6) {
                        delete pe; return fail(h, ORBX_ERR_BAD_ARGUMENT, "internal: irregular FAST cell row");
                    }
                }
                if (npairs > 255 || npairs * (T.th - 6) > 65535) { delete pe; return fail(h, ORBX_ERR_BAD_ARGUMENT, "cell_size too large"); }
                pe->tiles.push_back(T);
                ft_maxtw = std::max(ft_maxtw, (int)T.tw); ft_maxth = std::max(ft_maxth, (int)T.th);
                ft_qcap = std::max(ft_qcap, (T.tw - 6) * (T.th - 6));
                if (npairs > 32) { delete pe; return fail(h, ORBX_ERR_BAD_ARGUMENT, "internal: more than 32 pairs per FAST tile row"); }
                // strict 3x3 maxima inside a cell: at most one per 2x2 block of its interior
                ft_scap = std::max(ft_scap, nc * ((V.wCell + 1) / 2) * ((T.th - 6 + 1) / 2));
                j0 += nc;
            }
        }
        V.ncells = (int)pe->cells.size() - V.cell_off;
        V.ntiles = (int)pe->tiles.size() - V.tile_off;
        // quadtree, :548-550
        V.N = h->quota[l];
        V.nIni = (int)roundf((float)(maxBX - minBX) / (maxBY - minBY));
        if (V.nIni < 1) { delete pe; return fail(h, ORBX_ERR_LEVEL_TOO_SMALL, "aspect ratio < 0.5 (nIni == 0 is UB in the reference, ORBextractor.cc:548)"); }
        V.hX = (float)(maxBX - minBX) / V.nIni;
        V.span_y = maxBY - minBY;
        const long long worst = (long long)std::max(V.ncells, 1) * ((V.wCell + 1) / 2) * ((V.hCell + 1) / 2);
        long long cap = std::max<long long>(1024, (long long)V.ncells * h->cand_per_cell);
        cap = std::min(cap, std::max<long long>(worst, 16));
        cap = align_up(cap, 4);
        if (cap >= (1 << 24)) { delete pe; return fail(h, ORBX_ERR_IMAGE_TOO_LARGE, "candidate capacity exceeds 2^24"); }
        V.cand_cap = (int)cap;
        V.cand_off = cand_off;
        cand_off += cap;
        V.kp_cap = std::max(V.N + 2, 4 * V.nIni) + 2;
        V.kp_off = kp_off;
        kp_off += V.kp_cap;
        qt_nc = std::max(qt_nc, V.kp_cap + 2);
        V.sf = h->sf[l];
        V.kp_size = (float)(int)(31 * h->sf[l]);           // scaledPatchSize, :872
    }
    P.ncells_total = (int)pe->cells.size();
    P.kp_total = kp_off;
    P.qt_nc = qt_nc;
    P.fast_tp = (int)align_up(maxcw + 6, 4);
    P.fast_trows = maxch;
    P.fast_qcap = (int)align_up((long long)(maxcw - 6) * (maxch - 6), 2);
    P.ntiles_total = (int)pe->tiles.size();
    P.ft_tp = (int)align_up(15 + ft_maxtw + 8, 16);        // alignment slack + the word right of the last pair (<= 256 for cells up to 227 px)
    P.ft_trows = ft_maxth;
    P.ft_qcap = (int)align_up(ft_qcap, 8);
    P.ft_scap = (int)align_up(ft_scap, 8);
    P.ft_tpmagic = 0xffffffffu / (unsigned)P.ft_tp + 1u;
    pe->ft_smem = 2 * align_up((size_t)P.ft_tp * P.ft_trows, 128) + 2 * (size_t)P.ft_qcap;    // tile image + score map (128-byte aligned) + queue
    if (pe->ft_smem > 200 * 1024 || (long long)P.ft_tp * P.ft_trows > 65535) { delete pe; return fail(h, ORBX_ERR_BAD_ARGUMENT, "cell_size too large"); }
    {   // p / ft_tp by __umulhi in k_fast_tiles: exact over the tile's byte range
        const unsigned m = P.ft_tpmagic;
        for (unsigned pp = 0; pp < (unsigned)(P.ft_tp * P.ft_trows); ++pp)
            if ((unsigned)(((unsigned long long)pp * m) >> 32) != pp / (unsigned)P.ft_tp) { delete pe; return fail(h, ORBX_ERR_BAD_ARGUMENT, "internal: tile pitch magic"); }
    }
    pe->pyr_stride = align_up(plane_off, 256);
    pe->blur_stride = align_up(blur_off, 256);
    pe->cand_stride = align_up(cand_off, 4);
    pe->fast_smem = (size_t)ORBX_FAST_WARPS * (size_t)align_up(2 * align_up((long long)P.fast_tp * P.fast_trows, 16) + 2ll * P.fast_qcap, 16);
    P.qt_sk = 1;
    while (P.qt_sk < qt_nc) P.qt_sk <<= 1;
    P.qt_bytes = align_up((long long)orbx_qt_bytes(qt_nc), 16);
    // node tables in shared memory while they fit (with room for several CTAs per SM); beyond that -- tens of thousands of features
    // on one level -- every (frame, level) tree gets a block of HBM instead (alloc_ws_set), the kernel is the same
    pe->qt_global = P.qt_bytes > 160 * 1024;
    pe->qt_smem = pe->qt_global ? 0 : (size_t)P.qt_bytes;
    if (qt_nc > 65535) { delete pe; return fail(h, ORBX_ERR_BAD_ARGUMENT, "more than 65533 features on one level (quadtree nodes are indexed with 16 bits)"); }
    if (pe->fast_smem > 200 * 1024) { delete pe; return fail(h, ORBX_ERR_BAD_ARGUMENT, "cell_size too large"); }
    std::vector<uint32_t> btiles;
    for (int l = 0; l < L; ++l)
        for (int ty = 0; ty < (P.lv[l].h + ORBX_BLUR_TH - 1) / ORBX_BLUR_TH; ++ty)
            for (int tx = 0; tx < (P.lv[l].w + ORBX_BLUR_TW - 1) / ORBX_BLUR_TW; ++tx)
                btiles.push_back((uint32_t)l | ((uint32_t)tx << 8) | ((uint32_t)ty << 20));
    pe->blur_tiles = (int)btiles.size();
    ORBX_CUDA(cudaMalloc(&pe->d_blur_tiles, btiles.size() * sizeof(uint32_t)));
    ORBX_CUDA(cudaMemcpy(pe->d_blur_tiles, btiles.data(), btiles.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    // device tables
    const size_t nx = std::max<size_t>(pe->xtab.size(), 1), ny = std::max<size_t>(pe->ytab.size(), 1);
    ORBX_CUDA(cudaMalloc(&pe->d_xtab, nx * sizeof(int2)));
    ORBX_CUDA(cudaMalloc(&pe->d_ytab, ny * sizeof(int2)));
    ORBX_CUDA(cudaMalloc(&pe->d_cells, std::max<size_t>(pe->cells.size(), 1) * sizeof(OrbxCell)));
    if (!pe->xtab.empty()) ORBX_CUDA(cudaMemcpy(pe->d_xtab, pe->xtab.data(), pe->xtab.size() * sizeof(int2), cudaMemcpyHostToDevice));
    if (!pe->ytab.empty()) ORBX_CUDA(cudaMemcpy(pe->d_ytab, pe->ytab.data(), pe->ytab.size() * sizeof(int2), cudaMemcpyHostToDevice));
    if (!pe->cells.empty()) ORBX_CUDA(cudaMemcpy(pe->d_cells, pe->cells.data(), pe->cells.size() * sizeof(OrbxCell), cudaMemcpyHostToDevice));
    {
        std::vector<uint8_t> sl((size_t)std::max(P.kp_total, 1), 0);
        for (int l = 0; l < L; ++l)
            for (int i = 0; i < P.lv[l].kp_cap; ++i) sl[(size_t)P.lv[l].kp_off + i] = (uint8_t)l;
        ORBX_CUDA(cudaMalloc(&pe->d_slot_level, sl.size()));
        ORBX_CUDA(cudaMemcpy(pe->d_slot_level, sl.data(), sl.size(), cudaMemcpyHostToDevice));
    }
    ORBX_CUDA(cudaMalloc(&pe->d_tiles, std::max<size_t>(pe->tiles.size(), 1) * sizeof(OrbxFastTile)));
    if (!pe->tiles.empty()) ORBX_CUDA(cudaMemcpy(pe->d_tiles, pe->tiles.data(), pe->tiles.size() * sizeof(OrbxFastTile), cudaMemcpyHostToDevice));
    *out = pe;
    return ORBX_OK;
}

void free_ws_set(OrbxWs& w, bool own_flags) {
    cudaFree(w.pyr); cudaFree(w.blur); cudaFree(w.cand); cudaFree(w.keynode); cudaFree(const_cast<uint8_t*>(w.tmaps)); cudaFree(const_cast<uint8_t*>(w.tmaps_blur)); cudaFree(const_cast<uint8_t*>(w.tmaps_b7)); cudaFree(const_cast<uint8_t*>(w.tmaps_rs)); cudaFree(w.qt_scratch);
    cudaFree(w.kprec); cudaFree(w.cand_count); cudaFree(w.level_count);
    if (own_flags) cudaFree(w.flags);
    memset(&w, 0, sizeof(w));
}

void free_workspace(OrbxHandle* h) {
    for (auto& kv : h->plans)      // captured graphs hold workspace pointers
        for (auto& g : kv.second->graphs)
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    free_ws_set(h->ws2, false);
    free_ws_set(h->ws, true);
    h->ws_plan = nullptr; h->ws_frames = 0; h->ws2_frames = 0; h->resident_frames = 0; h->cur = nullptr; h->res_set = 0;
}

// TMA tensor maps, one per level, 3-D: (byte column, row, frame) with strides (row pitch, frame stride) and a fixed box.
//   planes : over a workspace's pyramid block, box = ft_tp x ft_trows -- the FAST tile image of k_fast_tiles<true>
//   blur   : over its blurred levels, box = 48 x 37 -- the descriptor window of k_describe<true>
// Returns false when the driver entry point is missing or refuses a map: the cp.async instances are used instead.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled tensor_map_encoder() {
    static PFN_encodeTiled encode = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            encode = (PFN_encodeTiled)fn;
        else
            cudaGetLastError();
        if (!encode && getenv("ORBX_DEBUG")) fprintf(stderr, "orbx: cuTensorMapEncodeTiled entry point not found\n");
    }
    return encode;
}

// box_w / box_h <= 0: per level, the resize kernel's staging window (entry l describes the plane of level l - 1)
bool encode_level_maps(PlanEntry* pe, int frames, bool blur, uint8_t* base, long long frame_stride, int box_w_all, int box_h_all,
                       std::vector<CUtensorMap>& out) {
    const OrbxPlan& P = pe->plan;
    PFN_encodeTiled encode = tensor_map_encoder();
    if (!encode) return false;
    const bool resize = box_w_all <= 0;
    out.resize((size_t)P.nlevels);
    memset(out.data(), 0, out.size() * sizeof(CUtensorMap));
    for (int l = resize ? 1 : 0; l < P.nlevels; ++l) {
        const OrbxLevel& V = P.lv[resize ? l - 1 : l];
        const int box_w = resize ? pe->rs_launch_pitch[l] : box_w_all, box_h = resize ? pe->rs_rows[l] : box_h_all;
        if (box_w > 256 || box_h > 256 || (box_w & 15) || box_w < 16 || box_h < 1) return false;
        const int pitch = blur ? V.blur_pitch : V.pitch, rows = blur ? V.h : V.plane_rows;
        const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)frames};
        const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frame_stride};
        const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        const CUresult cr = encode(&out[(size_t)l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base + (blur ? V.blur_off : V.plane_off), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) {
            if (getenv("ORBX_DEBUG")) fprintf(stderr, "orbx: cuTensorMapEncodeTiled(level %d, %s) failed with CUresult %d\n", l, blur ? "blur" : "planes", (int)cr);
            return false;
        }
    }
    return true;
}

int upload_maps(OrbxHandle* h, const std::vector<CUtensorMap>& maps, const uint8_t** out) {
    uint8_t* d = nullptr;
    ORBX_CUDA(cudaMalloc(&d, maps.size() * sizeof(CUtensorMap)));
    ORBX_CUDA(cudaMemcpy(d, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    *out = d;
    return ORBX_OK;
}

int alloc_ws_set(OrbxHandle* h, PlanEntry* pe, int frames, OrbxWs& w, int* shared_flags) {
    const OrbxPlan& P = pe->plan;
    ORBX_CUDA(cudaMalloc(&w.pyr, (size_t)pe->pyr_stride * frames));
    ORBX_CUDA(cudaMalloc(&w.blur, (size_t)pe->blur_stride * frames));
    ORBX_CUDA(cudaMalloc(&w.cand, (size_t)pe->cand_stride * frames * sizeof(uint2)));
    ORBX_CUDA(cudaMalloc(&w.keynode, (size_t)pe->cand_stride * frames * sizeof(uint16_t)));
    if (getenv("ORBX_POISON")) {   // test aid: stale workspace bytes (unwritten borders, padding) must never reach a result
        ORBX_CUDA(cudaMemset(w.pyr, 0xA5, (size_t)pe->pyr_stride * frames));
        ORBX_CUDA(cudaMemset(w.blur, 0x5A, (size_t)pe->blur_stride * frames));
    }
    ORBX_CUDA(cudaMalloc(&w.kprec, (size_t)P.kp_total * frames * sizeof(OrbxKpRec)));
    w.qt_scratch = nullptr;
    if (pe->qt_global) ORBX_CUDA(cudaMalloc(&w.qt_scratch, (size_t)P.qt_bytes * P.nlevels * frames));
    ORBX_CUDA(cudaMalloc(&w.cand_count, (size_t)P.nlevels * frames * sizeof(int)));
    ORBX_CUDA(cudaMalloc(&w.level_count, (size_t)P.nlevels * frames * sizeof(int2)));
    if (shared_flags) {
        w.flags = shared_flags;
    } else {
        ORBX_CUDA(cudaMalloc(&w.flags, sizeof(int)));
        ORBX_CUDA(cudaMemset(w.flags, 0, sizeof(int)));
    }
    w.tmaps = nullptr; w.tmaps_blur = nullptr; w.tmaps_b7 = nullptr; w.tmaps_rs = nullptr;
    if (h->fast_tma) {
        std::vector<CUtensorMap> maps;
        if (P.ntiles_total > 0 && encode_level_maps(pe, frames, false, w.pyr, pe->pyr_stride, P.ft_tp, P.ft_trows, maps)) {
            const int ru = upload_maps(h, maps, &w.tmaps);
            if (ru != ORBX_OK) return ru;
        }
        if (encode_level_maps(pe, frames, true, w.blur, pe->blur_stride, ORBX_DESC_PP, 37, maps)) {
            const int ru = upload_maps(h, maps, &w.tmaps_blur);
            if (ru != ORBX_OK) return ru;
        }
        if (encode_level_maps(pe, frames, false, w.pyr, pe->pyr_stride, ORBX_BLUR_TW + 32, ORBX_BLUR_TH + 6, maps)) {
            const int ru = upload_maps(h, maps, &w.tmaps_b7);
            if (ru != ORBX_OK) return ru;
        }
        if (P.nlevels > 1 && encode_level_maps(pe, frames, false, w.pyr, pe->pyr_stride, 0, 0, maps)) {
            const int ru = upload_maps(h, maps, &w.tmaps_rs);
            if (ru != ORBX_OK) return ru;
        }
    }
    w.slot_level = pe->d_slot_level;
    w.pyr_stride = pe->pyr_stride; w.blur_stride = pe->blur_stride; w.cand_stride = pe->cand_stride;
    w.kp_stride = P.kp_total;
    w.xtab = pe->d_xtab; w.ytab = pe->d_ytab; w.cells = pe->d_cells; w.tiles = pe->d_tiles;
    w.pattern_f = h->d_pattern_f; w.angle_w = h->d_angle_w; w.blur_tiles = pe->d_blur_tiles;
    return ORBX_OK;
}

// `sets` = 2 allocates the second workspace set used to overlap consecutive launch groups.
int ensure_workspace(OrbxHandle* h, PlanEntry* pe, int frames, int sets = 1) {
    if (h->ws_plan != pe || h->ws_frames < frames) {
        ORBX_CUDA(cudaStreamSynchronize(h->stream));
        ORBX_CUDA(cudaStreamSynchronize(h->stream2));
        free_workspace(h);
        int rc = alloc_ws_set(h, pe, frames, h->ws, nullptr);
        if (rc != ORBX_OK) return rc;
        h->ws_plan = pe; h->ws_frames = frames;
    }
    if (sets > 1 && h->ws2_frames < frames) {
        ORBX_CUDA(cudaStreamSynchronize(h->stream2));
        free_ws_set(h->ws2, false);
        h->ws2_frames = 0;
        int rc = alloc_ws_set(h, pe, frames, h->ws2, h->ws.flags);
        if (rc != ORBX_OK) return rc;
        h->ws2_frames = frames;
    }
    return ORBX_OK;
}

int get_plan(OrbxHandle* h, int width, int height, PlanEntry** out) {
    auto key = std::make_pair(width, height);
    auto it = h->plans.find(key);
    if (it != h->plans.end()) { *out = it->second; return ORBX_OK; }
    PlanEntry* pe = nullptr;
    int rc = build_plan(h, width, height, &pe);
    if (rc != ORBX_OK) return rc;
    h->plans[key] = pe;
    *out = pe;
    return ORBX_OK;
}

void drop_plans(OrbxHandle* h) {
    free_workspace(h);
    for (auto& kv : h->plans) free_plan(kv.second);
    h->plans.clear();
}

enum { STAGES_PYRAMID = 1, STAGES_KEYPOINTS = 2, STAGES_ALL = 3 };

// dynamic shared memory of k_pyr_resize: staged source window + 16-bit horizontal sums + destination-row descriptors
size_t rs_smem_bytes(int rows, int pitch) { return (size_t)rows * pitch + (size_t)rows * ORBX_RS_TW * 2 + (size_t)ORBX_RS_TH * 16; }

int launch_border(OrbxHandle* h, PlanEntry* pe, const OrbxWs& ws, int nf, cudaStream_t st) {
    const OrbxPlan& P = pe->plan;
    int max_rows = 0;
    for (int l = 0; l < P.nlevels; ++l) max_rows = std::max(max_rows, P.lv[l].plane_rows);
    k_pyr_border<<<dim3((max_rows + ORBX_BORDER_ROWS - 1) / ORBX_BORDER_ROWS, P.nlevels, nf), dim3(16, 16), 0, st>>>(P, ws);
    (void)h;
    return ORBX_OK;
}

// Launch the stages for `nf` device-resident frames.  Outputs (device pointers, may be NULL) are written
// at frame index frame_out0 + f.
int launch_group_raw(OrbxHandle* h, PlanEntry* pe, cudaStream_t st, const uint8_t* d_imgs, long long row_stride,
                     long long frame_stride, int nf, int lap0, int lap1, void* d_kps, uint8_t* d_desc, int cap_per_frame,
                     int32_t* d_counts, int frame_out0, int stages, bool in_capture, int set = 0) {
    OrbxPlan P = pe->plan;
    P.lap0 = lap0; P.lap1 = lap1;
    const OrbxWs& ws = set ? h->ws2 : h->ws;
    if (nf < 1 || nf > (set ? h->ws2_frames : h->ws_frames) || h->ws_plan != pe)
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "internal: launch group larger than the workspace");
    const bool prof = (h->prm.flags & ORBX_FLAG_PROFILE) != 0 && !in_capture;
    StageEvents* se = nullptr;
    if (prof) {
        if (h->events_used == h->events.size()) {
            StageEvents e;
            for (auto& x : e.ev) ORBX_CUDA(cudaEventCreate(&x));
            h->events.push_back(e);
        }
        se = &h->events[h->events_used++];
        ORBX_CUDA(cudaEventRecord(se->ev[0], st));
    }
    int64_t launches = 0;
    // NVTX ranges around the launches of each stage (the reference's only tracing is a std::chrono pair
    // around ExtractORB, src/Frame.cc:106-117); they show up on the host timeline of nsys / ncu and cost nothing otherwise.
    nvtxRangePushA("orbx:pyramid");
    struct NvtxPop { ~NvtxPop() { nvtxRangePop(); } } nvtx_pop;
    auto nvtx_stage = [](const char* name) { nvtxRangePop(); nvtxRangePushA(name); };
    // Single-frame graphs (one or two frames, everything captured): FAST and the quadtree of level l form their own branch
    // that starts as soon as level l exists, instead of waiting for the whole pyramid -- level 0's quadtree is the longest
    // kernel of a frame and now runs beside the resize cascade.
    const bool per_level = in_capture && stages == STAGES_ALL && nf <= 2;
    if (per_level) {
        ORBX_CUDA(cudaMemsetAsync(ws.cand_count, 0, (size_t)P.nlevels * nf * sizeof(int), st));
    }
    auto level_branch = [&](int l) -> int {
        const OrbxLevel& V = P.lv[l];
        cudaStream_t sl = h->s_lvl[l];
        ORBX_CUDA(cudaEventRecord(h->ev_lvl_ready[l], st));
        ORBX_CUDA(cudaStreamWaitEvent(sl, h->ev_lvl_ready[l], 0));
        if (h->fast_v1)
            k_fast_cells<<<dim3((V.ncells + ORBX_FAST_WARPS - 1) / ORBX_FAST_WARPS, nf), ORBX_FAST_WARPS * 32, pe->fast_smem, sl>>>(P, ws, V.cell_off,
                                                                                                                              V.cell_off + V.ncells);
        else if (V.ntiles > 0 && ws.tmaps)
            (P.ft_tp == 256 ? k_fast_tiles<true, 256> : k_fast_tiles<true, 0>)<<<dim3(V.ntiles, nf), ORBX_FT_THREADS, pe->ft_smem, sl>>>(P, ws, V.tile_off);
        else if (V.ntiles > 0)
            k_fast_tiles<false><<<dim3(V.ntiles, nf), ORBX_FT_THREADS, pe->ft_smem, sl>>>(P, ws, V.tile_off);
        k_octree<ORBX_QT_THREADS_BIG><<<dim3(nf, 1), ORBX_QT_THREADS_BIG, pe->qt_smem, sl>>>(P, ws, l);
        launches += 2;
        ORBX_CUDA(cudaEventRecord(h->ev_lvl_done[l], sl));
        return ORBX_OK;
    };
    if (stages & STAGES_PYRAMID) {
        {
            const OrbxLevel& V = P.lv[0];
            const dim3 blk(32, 8);
            const dim3 grd(((V.w + 15) / 16 + 31) / 32, (V.h + 7) / 8, nf);
            const int aligned16 = ((uintptr_t)d_imgs % 16 == 0 && row_stride % 16 == 0 && frame_stride % 16 == 0) ? 1 : 0;
            k_pyr_level0<<<grd, blk, 0, st>>>(P, ws, d_imgs, row_stride, frame_stride, aligned16);
            ++launches;
            if (per_level) { const int rb = level_branch(0); if (rb != ORBX_OK) return rb; }
        }
        for (int l = 1; l < P.nlevels; ++l) {
            const OrbxLevel& V = P.lv[l];
            const bool lat = nf <= 2;                                         // latency mode: 128x16 tiles
            const int tile_h = lat ? ORBX_RS_TH_LAT : ORBX_RS_TH;
            const dim3 grd((V.w + ORBX_RS_TW - 1) / ORBX_RS_TW, (V.h + tile_h - 1) / tile_h, nf);
            const bool area = pe->xtab[V.xtab_off].y == -1;
            const bool fixed = !area && pe->rs_pitch[l] <= ORBX_RS_PITCH;     // specialised instance: constant staging pitch
            const int pitch = fixed ? ORBX_RS_PITCH : pe->rs_pitch[l];
            const size_t smem = rs_smem_bytes(pe->rs_rows[l], pitch);
            if (lat) {
                if (area) k_pyr_resize<true, 0, ORBX_RS_TH_LAT><<<grd, 256, smem, st>>>(P, ws, l, pe->rs_rows[l], pitch, pe->rs_pairs[l]);
                else if (fixed) k_pyr_resize<false, ORBX_RS_PITCH, ORBX_RS_TH_LAT><<<grd, 256, smem, st>>>(P, ws, l, pe->rs_rows[l], pitch, pe->rs_pairs[l]);
                else k_pyr_resize<false, 0, ORBX_RS_TH_LAT><<<grd, 256, smem, st>>>(P, ws, l, pe->rs_rows[l], pitch, pe->rs_pairs[l]);
            } else {
                const bool tma = ws.tmaps_rs != nullptr && !area;
                if (area) k_pyr_resize<true, 0, ORBX_RS_TH><<<grd, 256, smem, st>>>(P, ws, l, pe->rs_rows[l], pitch, pe->rs_pairs[l]);
                else if (fixed && tma) k_pyr_resize<false, ORBX_RS_PITCH, ORBX_RS_TH, true><<<grd, 256, smem, st>>>(P, ws, l, pe->rs_rows[l], pitch, pe->rs_pairs[l]);
                else if (fixed) k_pyr_resize<false, ORBX_RS_PITCH, ORBX_RS_TH><<<grd, 256, smem, st>>>(P, ws, l, pe->rs_rows[l], pitch, pe->rs_pairs[l]);
                else if (tma) k_pyr_resize<false, 0, ORBX_RS_TH, true><<<grd, 256, smem, st>>>(P, ws, l, pe->rs_rows[l], pitch, pe->rs_pairs[l]);
                else k_pyr_resize<false, 0, ORBX_RS_TH><<<grd, 256, smem, st>>>(P, ws, l, pe->rs_rows[l], pitch, pe->rs_pairs[l]);
            }
            ++launches;
            if (per_level) { const int rb = level_branch(l); if (rb != ORBX_OK) return rb; }
        }
        // In a captured graph the border fill and the blur form a side branch: neither FAST nor the quadtree
        // reads the border or the blurred levels, only k_describe (and pyramid downloads) do.
        const bool fork = in_capture && stages == STAGES_ALL;
        cudaStream_t sb = st;
        if (fork) {
            ORBX_CUDA(cudaEventRecord(h->ev_fork, st));
            ORBX_CUDA(cudaStreamWaitEvent(h->s_side, h->ev_fork, 0));
            sb = h->s_side;
        }
        // The 19-px border is read by nobody on this path (k_pyr_border): it is written here only for one or two frames (the
        // drop-in operator() hands mvImagePyramid to its caller; off the critical path on the side branch) and when a pyramid
        // sink is set; otherwise on demand, by the accessors that take planes out of the device (ensure_borders).
        const bool want_border = nf <= 2 || h->pyr_out != nullptr;
        if (want_border) {
            launch_border(h, pe, ws, nf, sb);
            ++launches;
        }
        h->borders_valid[set] = want_border;
        if (fork) {
            if (ws.tmaps_b7) k_blur7<true><<<dim3(pe->blur_tiles, nf), 256, 0, sb>>>(P, ws, 0);
            else k_blur7<false><<<dim3(pe->blur_tiles, nf), 256, 0, sb>>>(P, ws, 0);
            ++launches;
            ORBX_CUDA(cudaEventRecord(h->ev_join, sb));
        }
    }
    const bool forked = in_capture && stages == STAGES_ALL;
    if (se) ORBX_CUDA(cudaEventRecord(se->ev[1], st));
    nvtx_stage("orbx:fast");
    if (per_level) {
        for (int l = 0; l < P.nlevels; ++l) ORBX_CUDA(cudaStreamWaitEvent(st, h->ev_lvl_done[l], 0));
        ORBX_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
        if (ws.tmaps_blur)
            k_describe<true><<<dim3((P.kp_total + 2 * ORBX_DESC_WARPS - 1) / (2 * ORBX_DESC_WARPS), nf), ORBX_DESC_WARPS * 32, 0, st>>>(
                P, ws, h->fc, d_kps, d_desc, cap_per_frame, d_counts, frame_out0);
        else
            k_describe<false><<<dim3((P.kp_total + 2 * ORBX_DESC_WARPS - 1) / (2 * ORBX_DESC_WARPS), nf), ORBX_DESC_WARPS * 32, 0, st>>>(
                P, ws, h->fc, d_kps, d_desc, cap_per_frame, d_counts, frame_out0);
        ++launches;
    } else if (stages & STAGES_KEYPOINTS) {
        ORBX_CUDA(cudaMemsetAsync(ws.cand_count, 0, (size_t)P.nlevels * nf * sizeof(int), st));
        if (h->fast_v1)
            k_fast_cells<<<dim3((P.ncells_total + ORBX_FAST_WARPS - 1) / ORBX_FAST_WARPS, nf), ORBX_FAST_WARPS * 32, pe->fast_smem, st>>>(P, ws, 0, P.ncells_total);
        else if (P.ntiles_total > 0 && ws.tmaps)
            (P.ft_tp == 256 ? k_fast_tiles<true, 256> : k_fast_tiles<true, 0>)<<<dim3(P.ntiles_total, nf), ORBX_FT_THREADS, pe->ft_smem, st>>>(P, ws, 0);
        else if (P.ntiles_total > 0)
            k_fast_tiles<false><<<dim3(P.ntiles_total, nf), ORBX_FT_THREADS, pe->ft_smem, st>>>(P, ws, 0);
        ++launches;
        if (se) ORBX_CUDA(cudaEventRecord(se->ev[2], st));
        nvtx_stage("orbx:octree");
        // levels of a megapixel or more hold tens of thousands of candidates each: their quadtrees get 1024-thread CTAs
        const int nbig = qt_big_levels(P);
        if (nbig > 0) {
            k_octree<ORBX_QT_THREADS_BIG><<<dim3(nf, nbig), ORBX_QT_THREADS_BIG, pe->qt_smem, st>>>(P, ws, 0);
            ++launches;
        }
        if (nbig < P.nlevels) {
            if (nf <= 2) k_octree<ORBX_QT_THREADS_BIG><<<dim3(nf, P.nlevels - nbig), ORBX_QT_THREADS_BIG, pe->qt_smem, st>>>(P, ws, nbig);
            else if (nf <= 8) k_octree<ORBX_QT_THREADS_LAT><<<dim3(nf, P.nlevels - nbig), ORBX_QT_THREADS_LAT, pe->qt_smem, st>>>(P, ws, nbig);
            else if ((long long)P.lv[nbig].w * P.lv[nbig].h >= ORBX_QT_MID_PIXELS)
                k_octree<ORBX_QT_THREADS_MID><<<dim3(nf, P.nlevels - nbig), ORBX_QT_THREADS_MID, pe->qt_smem, st>>>(P, ws, nbig);
            else   // VGA-class levels (a few thousand candidates, ~250 nodes): 128-thread CTAs, 122 us instead of 137 us per 256 frames
                k_octree<ORBX_QT_THREADS><<<dim3(nf, P.nlevels - nbig), ORBX_QT_THREADS, pe->qt_smem, st>>>(P, ws, nbig);
            ++launches;
        }
        if (se) ORBX_CUDA(cudaEventRecord(se->ev[3], st));
        nvtx_stage("orbx:blur");
        if (forked) {
            ORBX_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
        } else {
            if (ws.tmaps_b7) k_blur7<true><<<dim3(pe->blur_tiles, nf), 256, 0, st>>>(P, ws, 1);
            else k_blur7<false><<<dim3(pe->blur_tiles, nf), 256, 0, st>>>(P, ws, 1);
            ++launches;
        }
        if (se) ORBX_CUDA(cudaEventRecord(se->ev[4], st));
        nvtx_stage("orbx:describe");
        if (ws.tmaps_blur)
            k_describe<true><<<dim3((P.kp_total + 2 * ORBX_DESC_WARPS - 1) / (2 * ORBX_DESC_WARPS), nf), ORBX_DESC_WARPS * 32, 0, st>>>(
                P, ws, h->fc, d_kps, d_desc, cap_per_frame, d_counts, frame_out0);
        else
            k_describe<false><<<dim3((P.kp_total + 2 * ORBX_DESC_WARPS - 1) / (2 * ORBX_DESC_WARPS), nf), ORBX_DESC_WARPS * 32, 0, st>>>(
                P, ws, h->fc, d_kps, d_desc, cap_per_frame, d_counts, frame_out0);
        ++launches;
        if (se) ORBX_CUDA(cudaEventRecord(se->ev[5], st));
    } else if (se) {
        for (int i = 2; i <= ORBX_NUM_STAGES; ++i) ORBX_CUDA(cudaEventRecord(se->ev[i], st));
    }
    if (!in_capture) ORBX_CUDA(cudaGetLastError());
    h->stage_launches += launches;
    h->total_launches += launches;
    h->cur = pe;
    h->resident_frames = nf;
    h->res_set = set;
    return ORBX_OK;
}

// Small launch groups are launch-latency bound (a single frame is 13 dependent kernels): replay them as one
// CUDA graph, captured once per argument set.  Large groups and profiled runs launch directly.
int launch_group(OrbxHandle* h, PlanEntry* pe, cudaStream_t st, const uint8_t* d_imgs, long long row_stride,
                 long long frame_stride, int nf, int lap0, int lap1, void* d_kps, uint8_t* d_desc, int cap_per_frame,
                 int32_t* d_counts, int frame_out0, int stages, int set = 0, bool single_group = true) {
    // Graphs pay only when the same argument set comes back (one frame / one small group per call).  The groups of a multi-group
    // call differ in their image / output pointers, so each would be captured and instantiated anew: launched directly instead.
    const bool use_graph = single_group && set == 0 && nf <= 8 && !(h->prm.flags & ORBX_FLAG_PROFILE) && !(h->prm.flags & ORBX_FLAG_NO_GRAPH);
    if (!use_graph)
        return launch_group_raw(h, pe, st, d_imgs, row_stride, frame_stride, nf, lap0, lap1, d_kps, d_desc, cap_per_frame, d_counts,
                                frame_out0, stages, false, set);
    const bool border = (stages & STAGES_PYRAMID) && (nf <= 2 || h->pyr_out != nullptr);
    const PlanEntry::GraphKey key{d_imgs, row_stride, frame_stride, nf, lap0, lap1, d_kps, d_desc, cap_per_frame, d_counts, frame_out0, stages, border};
    for (auto& g : pe->graphs)
        if (g.exec && g.key == key) {
            ORBX_CUDA(cudaGraphLaunch(g.exec, st));
            h->stage_launches += g.launches; h->total_launches += g.launches;
            h->cur = pe; h->resident_frames = nf; h->res_set = 0;
            if (stages & STAGES_PYRAMID) h->borders_valid[0] = border;
            return ORBX_OK;
        }
    cudaGraph_t graph = nullptr;
    const int64_t launches_before = h->total_launches;
    ORBX_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    int rc = launch_group_raw(h, pe, st, d_imgs, row_stride, frame_stride, nf, lap0, lap1, d_kps, d_desc, cap_per_frame, d_counts,
                              frame_out0, stages, true);
    cudaError_t ce = cudaStreamEndCapture(st, &graph);
    if (rc != ORBX_OK || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return rc != ORBX_OK ? rc : fail(h, ORBX_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(ce));
    }
    PlanEntry::GraphSlot& slot = pe->graphs[pe->graph_next];
    pe->graph_next = (pe->graph_next + 1) % 4;
    if (slot.exec) { cudaGraphExecDestroy(slot.exec); slot.exec = nullptr; }
    ce = cudaGraphInstantiate(&slot.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { slot.exec = nullptr; return fail(h, ORBX_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce)); }
    slot.key = key;
    slot.launches = h->total_launches - launches_before;
    ORBX_CUDA(cudaGraphLaunch(slot.exec, st));
    return ORBX_OK;
}

int collect_events(OrbxHandle* h) {
    for (size_t i = 0; i < h->events_used; ++i)
        for (int s = 0; s < ORBX_NUM_STAGES; ++s) {
            float ms = 0.f;
            ORBX_CUDA(cudaEventElapsedTime(&ms, h->events[i].ev[s], h->events[i].ev[s + 1]));
            h->stage_ms[s] += ms;
        }
    h->events_used = 0;
    return ORBX_OK;
}

int ensure_bytes(OrbxHandle* h, void** p, size_t* have, size_t need, bool pinned_host) {
    if (*have >= need) return ORBX_OK;
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    if (pinned_host) { if (*p) cudaFreeHost(*p); } else { if (*p) cudaFree(*p); }
    *p = nullptr; *have = 0;
    if (pinned_host) ORBX_CUDA(cudaMallocHost(p, need)); else ORBX_CUDA(cudaMalloc(p, need));
    *have = need;
    return ORBX_OK;
}

// Dynamic shared memory limits.  cudaFuncSetAttribute is per function and per device, shared by every handle of the process:
// two handles with different nfeatures / cell_size on different threads must not lower each other's limit between the call
// and the launch (left / right / 5x-nfeatures init extractors coexist in ORB-SLAM3).  So every kernel that uses dynamic shared
// memory gets the device's opt-in maximum (minus its static part) exactly once per device, under a process-wide mutex.
template <typename K>
cudaError_t raise_smem_limit(K kernel, int optin) {
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes);
}

int set_kernel_attrs_device(OrbxHandle* h) {
    static std::mutex mu;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lock(mu);
    if (h->device < 64 && done[h->device]) return ORBX_OK;
    int optin = 0;
    ORBX_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
    ORBX_CUDA(raise_smem_limit(k_fast_cells, optin));
    ORBX_CUDA(raise_smem_limit(k_fast_tiles<true>, optin));
    ORBX_CUDA(raise_smem_limit(k_fast_tiles<true, 256>, optin));
    ORBX_CUDA(raise_smem_limit(k_fast_tiles<false>, optin));
    ORBX_CUDA(raise_smem_limit(k_octree<ORBX_QT_THREADS>, optin));
    ORBX_CUDA(raise_smem_limit(k_octree<ORBX_QT_THREADS_MID>, optin));
    ORBX_CUDA(raise_smem_limit(k_octree<ORBX_QT_THREADS_LAT>, optin));
    ORBX_CUDA(raise_smem_limit(k_octree<ORBX_QT_THREADS_BIG>, optin));
    ORBX_CUDA(raise_smem_limit(k_pyr_resize<false, 0, ORBX_RS_TH>, optin));
    ORBX_CUDA(raise_smem_limit(k_pyr_resize<false, ORBX_RS_PITCH, ORBX_RS_TH>, optin));
    ORBX_CUDA(raise_smem_limit(k_pyr_resize<false, ORBX_RS_PITCH, ORBX_RS_TH, true>, optin));
    ORBX_CUDA(raise_smem_limit(k_pyr_resize<false, 0, ORBX_RS_TH, true>, optin));
    ORBX_CUDA(raise_smem_limit(k_pyr_resize<true, 0, ORBX_RS_TH>, optin));
    ORBX_CUDA(raise_smem_limit(k_pyr_resize<false, 0, ORBX_RS_TH_LAT>, optin));
    ORBX_CUDA(raise_smem_limit(k_pyr_resize<false, ORBX_RS_PITCH, ORBX_RS_TH_LAT>, optin));
    ORBX_CUDA(raise_smem_limit(k_pyr_resize<true, 0, ORBX_RS_TH_LAT>, optin));
    ORBX_CUDA(raise_smem_limit(k_init_resolve, optin));
    ORBX_CUDA(raise_smem_limit(k_clahe_apply, optin));
    if (h->device < 64) done[h->device] = true;
    return ORBX_OK;
}

int set_kernel_attrs(OrbxHandle* h, PlanEntry* pe) {
    size_t rs = 0;
    for (int l = 1; l < pe->plan.nlevels; ++l)
        rs = std::max(rs, rs_smem_bytes(pe->rs_rows[l], std::max(pe->rs_pitch[l], ORBX_RS_PITCH)));
    if (rs > 200 * 1024) return fail(h, ORBX_ERR_BAD_ARGUMENT, "scale factor too large for the resize kernel's shared memory");
    return set_kernel_attrs_device(h);
}

// Did any frame since the last check overflow its candidate workspace?  The flag word is fetched into pinned
// memory by fetch_overflow (queued on the compute stream before its final synchronize) and cleared when set.
int fetch_overflow(OrbxHandle* h, cudaStream_t st) {
    ORBX_CUDA(cudaMemcpyAsync(h->h_flag, h->ws.flags, sizeof(int), cudaMemcpyDeviceToHost, st));
    return ORBX_OK;
}
int check_overflow(OrbxHandle* h, bool* overflow, cudaStream_t st) {
    const int flag = *h->h_flag;
    *overflow = (flag & 1) != 0;
    if (flag) {   // cleared on the compute stream (ordered before the next launch there), then waited for: the caller may re-plan
        ORBX_CUDA(cudaMemsetAsync(h->ws.flags, 0, sizeof(int), st));
        ORBX_CUDA(cudaStreamSynchronize(st));
    }
    return ORBX_OK;
}

}  // namespace

static const OrbxWs& res_ws(const OrbxHandle* h) { return h->res_set ? h->ws2 : h->ws; }

// mvImagePyramid leaves the device with its border (reference :1173-1177, :1193): written now if the extraction skipped it.
static int ensure_borders(OrbxHandle* h) {
    if (!h->cur || h->resident_frames < 1 || h->borders_valid[h->res_set]) return ORBX_OK;
    launch_border(h, h->cur, res_ws(h), h->resident_frames, h->stream);
    h->total_launches += 1; h->stage_launches += 1;
    ORBX_CUDA(cudaGetLastError());
    h->borders_valid[h->res_set] = true;
    return ORBX_OK;
}

extern "C" {

const char* orbx_status_string(int s) {
    switch (s) {
        case ORBX_OK: return "ok";
        case ORBX_ERR_EMPTY_IMAGE: return "empty image";
        case ORBX_ERR_BAD_ARGUMENT: return "bad argument";
        case ORBX_ERR_LEVEL_TOO_SMALL: return "pyramid level too small";
        case ORBX_ERR_IMAGE_TOO_LARGE: return "image too large";
        case ORBX_ERR_CUDA: return "CUDA error";
        case ORBX_ERR_CAPACITY: return "output capacity too small";
        case ORBX_ERR_CANDIDATE_OVERFLOW: return "FAST candidate overflow";
        case ORBX_ERR_NO_FRAME: return "no resident frame";
        case ORBX_ERR_NO_DEVICE: return "no CUDA device";
        default: return "unknown status";
    }
}

const char* orbx_last_error(const OrbxHandle* h) { return h ? h->err.c_str() : "null handle"; }

int orbx_create(const OrbxParams* prm, int device, OrbxHandle** out) {
    if (!prm || !out) return ORBX_ERR_BAD_ARGUMENT;
    *out = nullptr;
    if (prm->nlevels < 1 || prm->nlevels > ORBX_MAX_LEVELS || prm->nfeatures < 0 || !(prm->scale_factor > 1.0f) ||
        prm->cell_size < 0 || prm->cell_size > 60 || prm->max_batch < 0)
        return ORBX_ERR_BAD_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ORBX_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return ORBX_ERR_BAD_ARGUMENT;
    OrbxHandle* h = new OrbxHandle();
    h->prm = *prm;
    if (h->prm.cell_size == 0) h->prm.cell_size = 30;
    if (h->prm.max_batch == 0) h->prm.max_batch = 1;
    h->cand_per_cell = prm->cand_per_cell > 0 ? prm->cand_per_cell : 64;
    h->device = device;
    if (const char* e1 = getenv("ORBX_FAST_V1")) h->fast_v1 = atoi(e1) != 0;
    if (const char* e4 = getenv("ORBX_FAST_TMA")) h->fast_tma = atoi(e4) != 0;
    if (const char* e2 = getenv("ORBX_SLOTS")) h->n_slots = std::min(4, std::max(2, atoi(e2)));
    if (const char* e3 = getenv("ORBX_RAMP")) h->ramp = atoi(e3) != 0;
    build_ctor_tables(h);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost(&h->h_flag, sizeof(int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_s2, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    for (int l = 0; l < prm->nlevels && l < ORBX_MAX_LEVELS && e == cudaSuccess; ++l) {
        e = cudaStreamCreateWithFlags(&h->s_lvl[l], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_lvl_ready[l], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_lvl_done[l], cudaEventDisableTiming);
    }
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&h->ev_pyr_ready[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_pyr_done[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_in_free[i], cudaEventDisableTiming);
    }
    {
        std::vector<float> pf(1024);
        // device layout [k][hl][4]: test t = 16*hl + k (descriptor bytes 2 hl, 2 hl + 1 = bits 0..15 of lane hl's word) so that a
        // half-warp's loads coalesce and the two half-warps of k_describe read the same words
        for (int t = 0; t < 256; ++t)
            for (int c = 0; c < 4; ++c) {
                const int slot = c == 1 ? 2 : (c == 2 ? 1 : c);      // stored as (x0, x1, y0, y1): the two points side by side for the packed FP32 ops
                pf[(size_t)((t & 15) * 16 + (t >> 4)) * 4 + slot] = (float)kPatternHost[4 * t + c];
            }
        // IC_Angle weight words: byte k of the aligned row window is patch column u = k - al - 15 (:82-97)
        // {u bytes, v bytes}, both signed and zero outside the circle; rows 31, 32 stay zero (k_describe steps three rows at a time)
        std::vector<int2> aw((size_t)4 * ORBX_ANGLE_ROWS * ORBX_ANGLE_WORDS, make_int2(0, 0));
        for (int al = 0; al < 4; ++al)
            for (int row = 0; row < 31; ++row)
                for (int wd = 0; wd < ORBX_ANGLE_WORDS; ++wd) {
                    unsigned wu = 0, wm = 0;
                    const int v = row - ORBX_HALF_PATCH;
                    for (int bb = 0; bb < 4; ++bb) {
                        const int uu = 4 * wd + bb - al - ORBX_HALF_PATCH;
                        const int au = uu < 0 ? -uu : uu;
                        if (au <= ORBX_HALF_PATCH && au <= h->umax[v < 0 ? -v : v]) {
                            wu |= (unsigned)(uu & 0xff) << (8 * bb);
                            wm |= (unsigned)(v & 0xff) << (8 * bb);
                        }
                    }
                    aw[((size_t)al * ORBX_ANGLE_ROWS + row) * ORBX_ANGLE_WORDS + wd] = make_int2((int)wu, (int)wm);
                }
        if (e == cudaSuccess) e = cudaMalloc(&h->d_pattern_f, pf.size() * sizeof(float));
        if (e == cudaSuccess) e = cudaMemcpy(h->d_pattern_f, pf.data(), pf.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMalloc(&h->d_angle_w, aw.size() * sizeof(int2));
        if (e == cudaSuccess) e = cudaMemcpy(h->d_angle_w, aw.data(), aw.size() * sizeof(int2), cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) { delete h; return ORBX_ERR_CUDA; }
    *out = h;
    return ORBX_OK;
}

void orbx_destroy(OrbxHandle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    drop_plans(h);
    cudaFree(h->d_stereo); cudaFree(h->d_clahe); cudaFree(h->d_frame); if (h->h_frame) cudaFreeHost(h->h_frame); if (h->h_out1) cudaFreeHost(h->h_out1); cudaFree(h->d_pattern_f); cudaFree(h->d_angle_w); cudaFree(h->d_in); cudaFree(h->d_out);
    for (auto& e : h->events) for (auto& x : e.ev) cudaEventDestroy(x);
    for (int i = 0; i < 4; ++i) {
        if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]);
        if (h->ev_d2h[i]) cudaEventDestroy(h->ev_d2h[i]);
        if (h->ev_h2d[i]) cudaEventDestroy(h->ev_h2d[i]);
        if (h->ev_in_free[i]) cudaEventDestroy(h->ev_in_free[i]);
    }
    if (h->h_flag) cudaFreeHost(h->h_flag);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_pyr_ready[i]) cudaEventDestroy(h->ev_pyr_ready[i]);
        if (h->ev_pyr_done[i]) cudaEventDestroy(h->ev_pyr_done[i]);
    }
    if (h->ev_s2) cudaEventDestroy(h->ev_s2);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->s_side) cudaStreamDestroy(h->s_side);
    for (int l = 0; l < ORBX_MAX_LEVELS; ++l) {
        if (h->s_lvl[l]) cudaStreamDestroy(h->s_lvl[l]);
        if (h->ev_lvl_ready[l]) cudaEventDestroy(h->ev_lvl_ready[l]);
        if (h->ev_lvl_done[l]) cudaEventDestroy(h->ev_lvl_done[l]);
    }
    if (h->stream2) cudaStreamDestroy(h->stream2);
    if (h->s_in) cudaStreamDestroy(h->s_in);
    if (h->s_out) cudaStreamDestroy(h->s_out);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int orbx_get_tables(const OrbxHandle* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                    int32_t* fpl, int32_t* umax16) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    const size_t L = (size_t)h->prm.nlevels;
    if (scale) memcpy(scale, h->sf.data(), L * sizeof(float));
    if (inv_scale) memcpy(inv_scale, h->inv_sf.data(), L * sizeof(float));
    if (sigma2) memcpy(sigma2, h->sigma2.data(), L * sizeof(float));
    if (inv_sigma2) memcpy(inv_sigma2, h->inv_sigma2.data(), L * sizeof(float));
    if (fpl) memcpy(fpl, h->quota.data(), L * sizeof(int32_t));
    if (umax16) memcpy(umax16, h->umax, 16 * sizeof(int32_t));
    return ORBX_OK;
}

int orbx_ctor_tables(const OrbxParams* prm, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* fpl,
                     int32_t* umax16) {
    if (!prm || prm->nlevels < 1 || prm->nlevels > 64 || prm->nfeatures < 0 || !(prm->scale_factor > 1.0f))   // tables only: no ORBX_MAX_LEVELS limit
        return ORBX_ERR_BAD_ARGUMENT;
    CtorTables t;
    compute_ctor_tables(prm->nfeatures, prm->scale_factor, prm->nlevels, t);
    const size_t L = (size_t)prm->nlevels;
    if (scale) memcpy(scale, t.sf.data(), L * sizeof(float));
    if (inv_scale) memcpy(inv_scale, t.inv_sf.data(), L * sizeof(float));
    if (sigma2) memcpy(sigma2, t.sigma2.data(), L * sizeof(float));
    if (inv_sigma2) memcpy(inv_sigma2, t.inv_sigma2.data(), L * sizeof(float));
    if (fpl) memcpy(fpl, t.quota.data(), L * sizeof(int32_t));
    if (umax16) memcpy(umax16, t.umax, 16 * sizeof(int32_t));
    return ORBX_OK;
}

int orbx_max_keypoints(const OrbxHandle* hc, int width, int height) {
    OrbxHandle* h = const_cast<OrbxHandle*>(hc);
    if (!h || width <= 0 || height <= 0) return ORBX_ERR_BAD_ARGUMENT;
    if (cudaSetDevice(h->device) != cudaSuccess) return ORBX_ERR_CUDA;
    PlanEntry* pe = nullptr;
    int rc = get_plan(h, width, height, &pe);
    if (rc != ORBX_OK) return rc;
    return pe->plan.kp_total;
}

}  // extern "C"

namespace {

// The frames of one call, handed out in launch groups.  One consumer (orbx_extract_batch) walks them in order; several
// consumers (orbx_extract_batch_multi: one host thread + handle per device) pull groups from the shared cursor, so a
// device that is fed more slowly (GPUs behind a shared PCIe switch get unequal shares) simply takes fewer groups and all
// devices finish together.
struct FramePool {
    std::atomic<int> cursor{0};
    int n_frames = 0;
    int consumers = 1;
};

// Next launch group for a consumer that has already taken `gi` groups.  Host-buffer pipelines ramp the group size up at
// the start of the call and down at its end, so that the first H2D copy (nothing to overlap with yet) and the last
// kernels + D2H copy (nothing left to overlap) are short.  Never more than `group` frames: that is what the workspace
// and the staging slots hold.
int next_group(FramePool& pool, int group, int gi, bool ramp, int* f0_out) {
    int f0 = pool.cursor.load(std::memory_order_relaxed);
    for (;;) {
        const int remaining = pool.n_frames - f0;
        if (remaining <= 0) return 0;
        int nf = std::min(group, remaining);
        if (ramp && group > 32) {
            const int up = gi < 4 ? (32 << gi) : group;
            nf = std::min(nf, up);
            const int share = (remaining + pool.consumers - 1) / pool.consumers;       // what is left for this consumer
            if (share <= group) nf = std::min(nf, std::max(32, (share + 1) / 2));
            if (remaining - nf < 16 && remaining <= group) nf = remaining;             // no tiny tail group
        }
        if (pool.cursor.compare_exchange_weak(f0, f0 + nf, std::memory_order_relaxed)) { *f0_out = f0; return nf; }
    }
}

struct BatchArgs {
    const uint8_t* images; int in_mem; int n_frames; int width, height; size_t row_stride, frame_stride; int lap0, lap1;
    OrbxKeyPoint* kps; uint8_t* desc; int cap_per_frame; int32_t* counts; int out_mem;
};

// One pass of a consumer over the pool.  *overflow: some frame overflowed its FAST-candidate workspace (the caller regrows
// and repeats the call); *taken: frames this consumer processed.
int extract_pass(OrbxHandle* h, FramePool& pool, const BatchArgs& a, cudaStream_t st, bool* overflow, int* taken) {
    *overflow = false;
    *taken = 0;
    ORBX_CUDA(cudaSetDevice(h->device));
    const int n_frames = a.n_frames, width = a.width, height = a.height, cap_per_frame = a.cap_per_frame;
    PlanEntry* pe = nullptr;
    int rc = get_plan(h, width, height, &pe);
    if (rc != ORBX_OK) return rc;
    // Host-buffer pipelines of many groups run with launch groups of at most 128 frames: the finer grain shortens the fill and
    // drain of the three-stream pipeline (166.6 k vs 161.1 k frames/s end to end at 4096 frames per call, profiles/r02_e2e_knobs.txt),
    // while device-resident calls keep max_batch (larger groups are faster there).
    const bool host_side = a.in_mem == ORBX_MEM_HOST || a.out_mem == ORBX_MEM_HOST;
    const int group = std::min(host_side && n_frames >= 1024 ? std::min(h->prm.max_batch, 128) : h->prm.max_batch, n_frames);
    const bool multi_group = n_frames > group || pool.consumers > 1;
    // Consecutive launch groups alternate between two workspace sets and two compute streams, so that the
    // latency-bound kernels and the tail of every kernel of one group overlap the next group's kernels.
    // (Profiled runs stay on one stream: their per-stage event times must not overlap.)
    const bool dual = n_frames > group && !(h->prm.flags & ORBX_FLAG_PROFILE) && !(h->prm.flags & ORBX_FLAG_SINGLE_STREAM);
    rc = ensure_workspace(h, pe, group, dual ? 2 : 1);
    if (rc != ORBX_OK) return rc;
    rc = set_kernel_attrs(h, pe);
    if (rc != ORBX_OK) return rc;
    // Host buffers are pipelined through staging slots: copy-in (s_in), kernels (st / stream2), copy-out (s_out).
    const bool host_in = a.in_mem == ORBX_MEM_HOST, host_out = a.out_mem == ORBX_MEM_HOST;
    const bool copy_only = (h->prm.flags & ORBX_FLAG_COPY_ONLY) != 0;
    const size_t kp_bytes = (size_t)cap_per_frame * sizeof(OrbxKeyPoint), ds_bytes = (size_t)cap_per_frame * 32;
    const size_t in_slot = (size_t)align_up((long long)width * height * group, 256);
    const size_t o_kps_off = 0, o_desc_off = (size_t)align_up((long long)kp_bytes * group, 256);
    const size_t o_cnt_off = o_desc_off + (size_t)align_up((long long)ds_bytes * group, 256);
    const size_t out_slot = o_cnt_off + (size_t)align_up(8ll * group, 256);
    const int n_out_slots = multi_group ? h->n_slots : 1;
    if (host_out) { rc = ensure_bytes(h, &h->d_out, &h->d_out_bytes, (size_t)n_out_slots * out_slot, false); if (rc != ORBX_OK) return rc; }
    const bool single_out = host_out && n_frames == 1;
    if (single_out) { rc = ensure_bytes(h, (void**)&h->h_out1, &h->h_out1_bytes, out_slot, true); if (rc != ORBX_OK) return rc; }
    const int n_in_slots = multi_group ? h->n_slots : 1;
    if (host_in) { rc = ensure_bytes(h, (void**)&h->d_in, &h->d_in_bytes, (size_t)n_in_slots * in_slot, false); if (rc != ORBX_OK) return rc; }
    const bool ramp = (host_in || host_out) && h->ramp;
    OrbxKeyPoint* kps = a.kps; uint8_t* desc = a.desc; int32_t* counts = a.counts;
    const uint8_t* images = a.images;
    for (int gi = 0;; ++gi) {
        int f0 = 0;
        const int nf = next_group(pool, group, gi, ramp, &f0);
        if (nf == 0) break;
        *taken += nf;
        const int slot = gi & 1;
        const int set = dual ? slot : 0;
        cudaStream_t cs = set ? h->stream2 : st;
        const uint8_t* d_imgs;
        long long rs, fs;
        const int islot = gi % n_in_slots;
        if (host_in) {
            uint8_t* dst = h->d_in + (size_t)islot * in_slot;
            if (gi >= n_in_slots) ORBX_CUDA(cudaStreamWaitEvent(h->s_in, h->ev_in_free[islot], 0));   // slot's previous group consumed
            if (a.row_stride == (size_t)width && (nf == 1 || a.frame_stride == (size_t)width * height)) {
                ORBX_CUDA(cudaMemcpyAsync(dst, images + (size_t)f0 * a.frame_stride, (size_t)width * height * nf, cudaMemcpyHostToDevice, h->s_in));
            } else {
                for (int f = 0; f < nf; ++f)
                    ORBX_CUDA(cudaMemcpy2DAsync(dst + (size_t)f * width * height, (size_t)width, images + (size_t)(f0 + f) * a.frame_stride,
                                                a.row_stride, (size_t)width, (size_t)height, cudaMemcpyHostToDevice, h->s_in));
            }
            ORBX_CUDA(cudaEventRecord(h->ev_h2d[islot], h->s_in));
            ORBX_CUDA(cudaStreamWaitEvent(cs, h->ev_h2d[islot], 0));
            d_imgs = dst; rs = width; fs = (long long)width * height;
        } else {
            d_imgs = images + (size_t)f0 * a.frame_stride; rs = (long long)a.row_stride; fs = (long long)a.frame_stride;
        }
        // pyramid sink (orbx_set_pyramid_output): this workspace set's previous pyramids must have left before they are overwritten
        if (h->pyr_out && h->pyr_pending[set]) ORBX_CUDA(cudaStreamWaitEvent(cs, h->ev_pyr_done[set], 0));
        if (host_out) {
            const int oslot = gi % n_out_slots;
            uint8_t* ob = (uint8_t*)h->d_out + (size_t)oslot * out_slot;
            if (gi >= n_out_slots) ORBX_CUDA(cudaStreamWaitEvent(cs, h->ev_d2h[oslot], 0));   // slot's previous results copied out
            if (!copy_only) {
                rc = launch_group(h, pe, cs, d_imgs, rs, fs, nf, a.lap0, a.lap1, kps ? ob + o_kps_off : nullptr, desc ? ob + o_desc_off : nullptr,
                                  cap_per_frame, (int32_t*)(ob + o_cnt_off), 0, STAGES_ALL, set, !multi_group);
                if (rc != ORBX_OK) return rc;
            }
            ORBX_CUDA(cudaEventRecord(h->ev_done[oslot], cs));
            if (host_in) ORBX_CUDA(cudaEventRecord(h->ev_in_free[islot], cs));
            ORBX_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_done[oslot], 0));
            if (single_out) {
                // one frame: a single read-back into pinned memory; the caller's (usually pageable) arrays are filled after
                // the final synchronisation -- three copies into pageable memory would block the host one after the other
                ORBX_CUDA(cudaMemcpyAsync(h->h_out1, ob, out_slot, cudaMemcpyDeviceToHost, h->s_out));
            } else {
                if (kps) ORBX_CUDA(cudaMemcpyAsync(kps + (size_t)f0 * cap_per_frame, ob + o_kps_off, kp_bytes * nf, cudaMemcpyDeviceToHost, h->s_out));
                if (desc) ORBX_CUDA(cudaMemcpyAsync(desc + (size_t)f0 * cap_per_frame * 32, ob + o_desc_off, ds_bytes * nf, cudaMemcpyDeviceToHost, h->s_out));
                if (counts) ORBX_CUDA(cudaMemcpyAsync(counts + 2 * (size_t)f0, ob + o_cnt_off, 8 * (size_t)nf, cudaMemcpyDeviceToHost, h->s_out));
            }
            ORBX_CUDA(cudaEventRecord(h->ev_d2h[oslot], h->s_out));
        } else {
            if (!copy_only) {
                rc = launch_group(h, pe, cs, d_imgs, rs, fs, nf, a.lap0, a.lap1, kps, desc, cap_per_frame, counts, f0, STAGES_ALL, set, !multi_group);
                if (rc != ORBX_OK) return rc;
            }
            if (host_in) ORBX_CUDA(cudaEventRecord(h->ev_in_free[islot], cs));
        }
        if (h->pyr_out && !copy_only) {
            // mvImagePyramid to the host (reference :1173-1177 leaves it there): the group's bordered planes, one copy
            const OrbxWs& w = set ? h->ws2 : h->ws;
            ORBX_CUDA(cudaEventRecord(h->ev_pyr_ready[set], cs));
            ORBX_CUDA(cudaStreamWaitEvent(h->s_out, h->ev_pyr_ready[set], 0));
            ORBX_CUDA(cudaMemcpy2DAsync(h->pyr_out + (size_t)f0 * h->pyr_out_stride, h->pyr_out_stride, w.pyr, (size_t)pe->pyr_stride,
                                        std::min((size_t)pe->pyr_stride, h->pyr_out_stride), (size_t)nf, cudaMemcpyDeviceToHost, h->s_out));
            ORBX_CUDA(cudaEventRecord(h->ev_pyr_done[set], h->s_out));
            h->pyr_pending[set] = true;
        }
    }
    if (dual) {   // join the second compute stream into the caller's stream
        ORBX_CUDA(cudaEventRecord(h->ev_s2, h->stream2));
        ORBX_CUDA(cudaStreamWaitEvent(st, h->ev_s2, 0));
    }
    rc = fetch_overflow(h, st);
    if (rc != ORBX_OK) return rc;
    ORBX_CUDA(cudaStreamSynchronize(st));
    if (host_out || h->pyr_out) ORBX_CUDA(cudaStreamSynchronize(h->s_out));
    h->pyr_pending[0] = h->pyr_pending[1] = false;
    if (single_out && *taken == 1) {
        int32_t c2[2];
        std::memcpy(c2, h->h_out1 + o_cnt_off, 8);
        const size_t nk = (size_t)std::min(std::max(c2[0], 0), cap_per_frame);     // entries beyond n are unspecified padding
        if (kps) std::memcpy(kps, h->h_out1 + o_kps_off, nk * sizeof(OrbxKeyPoint));
        if (desc) std::memcpy(desc, h->h_out1 + o_desc_off, nk * 32);
        if (counts) std::memcpy(counts, c2, 8);
    }
    if (h->prm.flags & ORBX_FLAG_PROFILE) { rc = collect_events(h); if (rc != ORBX_OK) return rc; }
    return check_overflow(h, overflow, st);
}

int check_batch_args(OrbxHandle* h, const BatchArgs& a) {
    if (!a.images || a.width <= 0 || a.height <= 0 || a.n_frames <= 0) return fail(h, ORBX_ERR_EMPTY_IMAGE, "empty image");
    if (a.row_stride < (size_t)a.width || a.cap_per_frame < 0 || (a.n_frames > 1 && a.frame_stride < a.row_stride * (size_t)a.height))
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "bad strides or capacity");
    return ORBX_OK;
}

}  // namespace

extern "C" {

// The launch-group sizes a host-buffer call of n_frames frames is cut into (device-free: the same next_group the pipeline uses,
// consumers taking turns), for tests and for callers that size their own staging.
int orbx_plan_groups(int n_frames, int max_batch, int consumers, int ramp, int32_t* sizes, int capacity) {
    if (n_frames < 0 || max_batch < 1 || consumers < 1) return ORBX_ERR_BAD_ARGUMENT;
    FramePool pool;
    pool.n_frames = n_frames;
    pool.consumers = consumers;
    const int group = std::min(max_batch, std::max(n_frames, 1));
    std::vector<int> gi((size_t)consumers, 0);
    int n = 0;
    for (int c = 0;; c = (c + 1) % consumers) {
        int f0 = 0;
        const int nf = next_group(pool, group, gi[(size_t)c]++, ramp != 0, &f0);
        if (nf == 0) break;
        if (sizes && n < capacity) sizes[n] = nf;
        ++n;
    }
    return n;
}

int orbx_extract_batch(OrbxHandle* h, const uint8_t* images, int in_mem, int n_frames, int width, int height,
                       size_t row_stride, size_t frame_stride, int lap0, int lap1, OrbxKeyPoint* kps, uint8_t* desc,
                       int cap_per_frame, int32_t* counts, int out_mem, void* stream) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    const BatchArgs a{images, in_mem, n_frames, width, height, row_stride, frame_stride, lap0, lap1, kps, desc, cap_per_frame, counts, out_mem};
    int rc = check_batch_args(h, a);
    if (rc != ORBX_OK) return rc;
    ORBX_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    for (int attempt = 0;; ++attempt) {
        FramePool pool;
        pool.n_frames = n_frames;
        bool overflow = false;
        int taken = 0;
        rc = extract_pass(h, pool, a, st, &overflow, &taken);
        if (rc != ORBX_OK) return rc;
        if (!overflow) return ORBX_OK;
        // grow the candidate workspace and run the call again (every output is rewritten)
        if (attempt >= 6) return fail(h, ORBX_ERR_CANDIDATE_OVERFLOW, "FAST candidate workspace overflow after regrowth");
        h->cand_per_cell *= 4;
        drop_plans(h);
    }
}

// "Independent frames sharded across the GPUs of one box on per-GPU streams" from ONE process: handles[i] owns device i's
// streams and workspace, one host thread per handle runs the same three-stream pipeline as orbx_extract_batch and pulls its
// launch groups from a cursor shared by all of them (the reference's own concurrency is this pattern with two threads and
// two extractor instances, src/Frame.cc:109-112).  Host buffers only: feeding several devices from host memory is the point.
int orbx_extract_batch_multi(OrbxHandle* const* handles, int n_handles, const uint8_t* images, int n_frames, int width, int height,
                             size_t row_stride, size_t frame_stride, int lap0, int lap1, OrbxKeyPoint* kps, uint8_t* desc,
                             int cap_per_frame, int32_t* counts, int32_t* frames_per_handle) {
    if (!handles || n_handles < 1 || n_handles > 64) return ORBX_ERR_BAD_ARGUMENT;
    for (int i = 0; i < n_handles; ++i) {
        if (!handles[i]) return ORBX_ERR_BAD_ARGUMENT;
        for (int j = 0; j < i; ++j) if (handles[j] == handles[i]) return fail(handles[0], ORBX_ERR_BAD_ARGUMENT, "the same handle twice");
    }
    const BatchArgs a{images, ORBX_MEM_HOST, n_frames, width, height, row_stride, frame_stride, lap0, lap1, kps, desc, cap_per_frame, counts,
                      ORBX_MEM_HOST};
    int rc = check_batch_args(handles[0], a);
    if (rc != ORBX_OK) return rc;
    for (int attempt = 0;; ++attempt) {
        FramePool pool;
        pool.n_frames = n_frames;
        pool.consumers = n_handles;
        std::vector<int> rcs((size_t)n_handles, ORBX_OK), taken((size_t)n_handles, 0);
        std::vector<char> ovf((size_t)n_handles, 0);
        auto work = [&](int i) {
            bool o = false;
            rcs[(size_t)i] = extract_pass(handles[i], pool, a, handles[i]->stream, &o, &taken[(size_t)i]);
            ovf[(size_t)i] = o ? 1 : 0;
        };
        std::vector<std::thread> th;
        for (int i = 1; i < n_handles; ++i) th.emplace_back(work, i);
        work(0);
        for (auto& t : th) t.join();
        bool overflow = false;
        for (int i = 0; i < n_handles; ++i) {
            if (rcs[(size_t)i] != ORBX_OK) {
                if (i != 0) handles[0]->err = handles[i]->err;
                return rcs[(size_t)i];
            }
            overflow = overflow || ovf[(size_t)i];
            if (frames_per_handle) frames_per_handle[i] = taken[(size_t)i];
        }
        if (!overflow) return ORBX_OK;
        if (attempt >= 6) return fail(handles[0], ORBX_ERR_CANDIDATE_OVERFLOW, "FAST candidate workspace overflow after regrowth");
        for (int i = 0; i < n_handles; ++i) {
            if (cudaSetDevice(handles[i]->device) != cudaSuccess) return fail(handles[0], ORBX_ERR_CUDA, "cudaSetDevice");
            handles[i]->cand_per_cell *= 4;
            drop_plans(handles[i]);
        }
    }
}

int orbx_extract(OrbxHandle* h, const uint8_t* image, int width, int height, size_t stride, int lap0, int lap1,
                 OrbxKeyPoint* kps, uint8_t* desc, int capacity, int* n_out, int* mono_out) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    if (n_out) *n_out = 0;
    if (mono_out) *mono_out = 0;
    int32_t cnt[2] = {0, 0};
    int rc = orbx_extract_batch(h, image, ORBX_MEM_HOST, 1, width, height, stride, stride * (size_t)height, lap0, lap1, kps, desc,
                                capacity, cnt, ORBX_MEM_HOST, nullptr);
    if (rc != ORBX_OK) return rc;
    if (n_out) *n_out = cnt[0];
    if (mono_out) *mono_out = cnt[1];
    if (cnt[0] > capacity && (kps || desc)) return fail(h, ORBX_ERR_CAPACITY, "keypoint capacity too small");
    return ORBX_OK;
}

static int run_stages_single(OrbxHandle* h, const uint8_t* image, int width, int height, size_t stride, int stages) {
    ORBX_CUDA(cudaSetDevice(h->device));
    PlanEntry* pe = nullptr;
    int rc;
    if (stages & STAGES_PYRAMID) {
        if (!image || width <= 0 || height <= 0) return fail(h, ORBX_ERR_EMPTY_IMAGE, "empty image");
        rc = get_plan(h, width, height, &pe);
        if (rc != ORBX_OK) return rc;
        rc = ensure_workspace(h, pe, std::max(1, std::min(h->prm.max_batch, 1)));
        if (rc != ORBX_OK) return rc;
        rc = ensure_bytes(h, (void**)&h->d_in, &h->d_in_bytes, (size_t)width * height, false);
        if (rc != ORBX_OK) return rc;
        ORBX_CUDA(cudaMemcpy2DAsync(h->d_in, (size_t)width, image, stride, (size_t)width, (size_t)height, cudaMemcpyHostToDevice, h->stream));
    } else {
        pe = h->cur;
        if (!pe || h->resident_frames < 1) return fail(h, ORBX_ERR_NO_FRAME, "no resident pyramid");
    }
    rc = set_kernel_attrs(h, pe);
    if (rc != ORBX_OK) return rc;
    for (int attempt = 0;; ++attempt) {
        // ComputeKeyPointsOctTree alone runs on the resident pyramid: the workspace set the last group left it in
        const int set = (stages & STAGES_PYRAMID) ? 0 : h->res_set;
        rc = launch_group(h, pe, h->stream, h->d_in, width, (long long)width * height, 1, 0, 0, nullptr, nullptr, 0, nullptr, 0, stages, set);
        if (rc != ORBX_OK) return rc;
        rc = fetch_overflow(h, h->stream);
        if (rc != ORBX_OK) return rc;
        ORBX_CUDA(cudaStreamSynchronize(h->stream));
        if (h->prm.flags & ORBX_FLAG_PROFILE) { rc = collect_events(h); if (rc != ORBX_OK) return rc; }
        if (!(stages & STAGES_KEYPOINTS)) return ORBX_OK;
        bool ov = false;
        rc = check_overflow(h, &ov, h->stream);
        if (rc != ORBX_OK) return rc;
        if (!ov) return ORBX_OK;
        (void)attempt;
        return fail(h, ORBX_ERR_CANDIDATE_OVERFLOW, "FAST candidate workspace overflow (raise OrbxParams.cand_per_cell)");
    }
}

int orbx_compute_pyramid(OrbxHandle* h, const uint8_t* image, int width, int height, size_t stride) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    return run_stages_single(h, image, width, height, stride, STAGES_PYRAMID);
}

int orbx_compute_keypoints_octtree(OrbxHandle* h) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    return run_stages_single(h, nullptr, 0, 0, 0, STAGES_KEYPOINTS);
}

int orbx_distribute_octtree(OrbxHandle* h, const OrbxKeyPoint* keys, int n, int min_x, int max_x, int min_y, int max_y,
                            int n_features, OrbxKeyPoint* out, int capacity, int* n_out) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    if (n_out) *n_out = 0;
    if (n < 0 || (n > 0 && !keys) || max_x <= min_x || max_y <= min_y || n_features < 0)
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "bad DistributeOctTree arguments");
    ORBX_CUDA(cudaSetDevice(h->device));
    const int nIni = (int)roundf((float)(max_x - min_x) / (max_y - min_y));
    if (nIni < 1) return fail(h, ORBX_ERR_LEVEL_TOO_SMALL, "aspect ratio < 0.5 (nIni == 0)");
    const int CM = (1 << ORBX_COORD_BITS) - 1;
    std::vector<uint2> cand((size_t)std::max(n, 1));
    // Keys live in box coordinates, 0 <= x < max_x - min_x (the cell loop hands them over that way, :855-860).  A key
    // outside the box would index past vpIniNodes in the reference (:574, undefined behaviour) and past the node table
    // here: rejected up front, with the kernel's own float arithmetic for the initial-node index.
    const float hX0 = (float)(max_x - min_x) / nIni;
    for (int i = 0; i < n; ++i) {
        const float x = keys[i].x, y = keys[i].y, r = keys[i].response;
        if (!(x >= 0 && x <= CM && y >= 0 && y <= CM && x == floorf(x) && y == floorf(y) && r >= 0 && r <= 255 && r == floorf(r)))
            return fail(h, ORBX_ERR_BAD_ARGUMENT, "DistributeOctTree keys must have integer coordinates in [0,4095] and integer responses in [0,255]");
        if (!(x < (float)(max_x - min_x) && y <= (float)(max_y - min_y) && (int)(x / hX0) < nIni))
            return fail(h, ORBX_ERR_BAD_ARGUMENT, "DistributeOctTree key outside the box [0, maxX-minX) x [0, maxY-minY]");
        cand[i] = make_uint2((uint32_t)x | ((uint32_t)y << ORBX_COORD_BITS) | ((uint32_t)r << 24), (uint32_t)i);
    }
    if (n >= (1 << 24)) return fail(h, ORBX_ERR_BAD_ARGUMENT, "too many keys");
    // one-level plan + private workspace
    OrbxPlan P;
    memset(&P, 0, sizeof(P));
    P.nlevels = 1; P.lap0 = 0; P.lap1 = -1;
    OrbxLevel& V = P.lv[0];
    V.N = n_features; V.nIni = nIni; V.hX = (float)(max_x - min_x) / nIni; V.span_y = max_y - min_y;
    V.cand_cap = std::max(n, 1); V.cand_off = 0;
    V.kp_cap = std::max(n_features + 2, 4 * nIni) + 2; V.kp_off = 0; V.sf = 1.f;
    P.kp_total = V.kp_cap; P.qt_nc = V.kp_cap + 2;
    if (P.qt_nc > 65535) return fail(h, ORBX_ERR_BAD_ARGUMENT, "n_features too large (quadtree nodes are indexed with 16 bits)");
    P.qt_sk = 1;
    while (P.qt_sk < P.qt_nc) P.qt_sk <<= 1;
    P.qt_bytes = align_up((long long)orbx_qt_bytes(P.qt_nc), 16);
    const bool qt_global = P.qt_bytes > 160 * 1024;
    const size_t smem = qt_global ? 0 : (size_t)P.qt_bytes;
    OrbxWs w;
    memset(&w, 0, sizeof(w));
    int rc = ORBX_OK;
    uint2* d_cand = nullptr; uint16_t* d_kn = nullptr; OrbxKpRec* d_rec = nullptr; int* d_cnt = nullptr; int2* d_lc = nullptr; uint8_t* d_qt = nullptr;
    auto cleanup = [&]() { cudaFree(d_cand); cudaFree(d_kn); cudaFree(d_rec); cudaFree(d_cnt); cudaFree(d_lc); cudaFree(d_qt); };
#define ORBX_CUDA_L(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); return fail(h, ORBX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
    ORBX_CUDA_L(cudaMalloc(&d_cand, cand.size() * sizeof(uint2)));
    ORBX_CUDA_L(cudaMalloc(&d_kn, cand.size() * sizeof(uint16_t)));
    ORBX_CUDA_L(cudaMalloc(&d_rec, (size_t)V.kp_cap * sizeof(OrbxKpRec)));
    ORBX_CUDA_L(cudaMalloc(&d_cnt, sizeof(int)));
    ORBX_CUDA_L(cudaMalloc(&d_lc, sizeof(int2)));
    ORBX_CUDA_L(cudaMemcpyAsync(d_cand, cand.data(), cand.size() * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
    ORBX_CUDA_L(cudaMemcpyAsync(d_cnt, &n, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (qt_global) ORBX_CUDA_L(cudaMalloc(&d_qt, (size_t)P.qt_bytes));
    w.cand = d_cand; w.keynode = d_kn; w.kprec = d_rec; w.cand_count = d_cnt; w.level_count = d_lc; w.qt_scratch = d_qt;
    w.cand_stride = (long long)cand.size(); w.kp_stride = V.kp_cap;
    { const int ra = set_kernel_attrs_device(h); if (ra != ORBX_OK) { cleanup(); return ra; } }
    k_octree<ORBX_QT_THREADS_LAT><<<dim3(1, 1), ORBX_QT_THREADS_LAT, smem, h->stream>>>(P, w, 0);
    h->total_launches += 1; h->stage_launches += 1;
    ORBX_CUDA_L(cudaGetLastError());
    int2 lc;
    ORBX_CUDA_L(cudaMemcpyAsync(&lc, d_lc, sizeof(int2), cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA_L(cudaStreamSynchronize(h->stream));
    std::vector<OrbxKpRec> rec((size_t)std::max(lc.x, 1));
    ORBX_CUDA_L(cudaMemcpy(rec.data(), d_rec, (size_t)lc.x * sizeof(OrbxKpRec), cudaMemcpyDeviceToHost));
#undef ORBX_CUDA_L
    cleanup();
    if (n_out) *n_out = lc.x;
    for (int i = 0; i < lc.x && i < capacity; ++i) out[i] = keys[rec[i].src];
    if (lc.x > capacity) rc = fail(h, ORBX_ERR_CAPACITY, "keypoint capacity too small");
    return rc;
}

int orbx_stereo_match(OrbxHandle* h, OrbxHandle* right, const OrbxKeyPoint* keys_l, const uint8_t* desc_l, int n_l,
                      const OrbxKeyPoint* keys_r, const uint8_t* desc_r, int n_r, float mb, float mbf, float* u_right,
                      float* depth, int* n_matched) {
    if (!h || !right) return ORBX_ERR_BAD_ARGUMENT;
    if (n_matched) *n_matched = 0;
    if (n_l < 0 || n_r < 0 || (n_l > 0 && (!keys_l || !desc_l || !u_right || !depth)) || (n_r > 0 && (!keys_r || !desc_r)) || n_r >= 65535 ||
        !(mb > 0.f))
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "bad stereo arguments");
    if (!h->cur || !right->cur || h->resident_frames < 1 || right->resident_frames < 1)
        return fail(h, ORBX_ERR_NO_FRAME, "both extractors must hold a resident pyramid");
    if (h->device != right->device || h->cur->plan.width != right->cur->plan.width || h->cur->plan.height != right->cur->plan.height ||
        h->cur->plan.nlevels != right->cur->plan.nlevels)
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "left and right extractors must share device, image size and level count");
    if (n_l == 0) return ORBX_OK;
    ORBX_CUDA(cudaSetDevice(h->device));
    // scratch layout: kl | dl | kr | dr | u | d | sad | n
    const size_t o_kl = 0, o_dl = o_kl + align_up((long long)n_l * sizeof(OrbxKeyPoint), 256), o_kr = o_dl + align_up((long long)n_l * 32, 256);
    const size_t o_dr = o_kr + align_up((long long)std::max(n_r, 1) * sizeof(OrbxKeyPoint), 256), o_u = o_dr + align_up((long long)std::max(n_r, 1) * 32, 256);
    const size_t o_d = o_u + align_up((long long)n_l * 4, 256), o_s = o_d + align_up((long long)n_l * 4, 256), o_n = o_s + align_up((long long)n_l * 4, 256);
    int rc = ensure_bytes(h, (void**)&h->d_stereo, &h->d_stereo_bytes, o_n + 256, false);
    if (rc != ORBX_OK) return rc;
    uint8_t* b = h->d_stereo;
    cudaStream_t st = h->stream;
    ORBX_CUDA(cudaStreamSynchronize(right->stream));            // the right pyramid must be complete
    ORBX_CUDA(cudaMemcpyAsync(b + o_kl, keys_l, (size_t)n_l * sizeof(OrbxKeyPoint), cudaMemcpyHostToDevice, st));
    ORBX_CUDA(cudaMemcpyAsync(b + o_dl, desc_l, (size_t)n_l * 32, cudaMemcpyHostToDevice, st));
    if (n_r > 0) {
        ORBX_CUDA(cudaMemcpyAsync(b + o_kr, keys_r, (size_t)n_r * sizeof(OrbxKeyPoint), cudaMemcpyHostToDevice, st));
        ORBX_CUDA(cudaMemcpyAsync(b + o_dr, desc_r, (size_t)n_r * 32, cudaMemcpyHostToDevice, st));
    }
    OrbxStereoArgs a;
    a.kl = (const OrbxKeyPoint*)(b + o_kl); a.dl = (const uint32_t*)(b + o_dl); a.nl = n_l;
    a.kr = (const OrbxKeyPoint*)(b + o_kr); a.dr = (const uint32_t*)(b + o_dr); a.nr = n_r;
    a.kp_stride = 0; a.pyr_stride = 0; a.counts_l = nullptr; a.counts_r = nullptr; a.cap = 0;
    a.pyr_l = res_ws(h).pyr; a.pyr_r = res_ws(right).pyr;     // frame 0 of each handle's resident group
    a.mbf = mbf; a.min_d = 0.f; a.max_d = mbf / mb;           // minZ = mb, maxD = mbf / minZ (:844-846)
    a.th_mul = 1.5f * 1.4f;
    a.u_right = (float*)(b + o_u); a.depth = (float*)(b + o_d); a.sad = (int*)(b + o_s); a.n_matched = (int*)(b + o_n);
    k_stereo_match<<<(n_l + 7) / 8, 256, 0, st>>>(h->cur->plan, a);
    k_stereo_filter<<<1, 256, 0, st>>>(a);
    h->total_launches += 2; h->stage_launches += 2;
    ORBX_CUDA(cudaGetLastError());
    int nm = 0;
    ORBX_CUDA(cudaMemcpyAsync(u_right, b + o_u, (size_t)n_l * 4, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaMemcpyAsync(depth, b + o_d, (size_t)n_l * 4, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaMemcpyAsync(&nm, b + o_n, 4, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaStreamSynchronize(st));
    if (n_matched) *n_matched = nm;
    return ORBX_OK;
}

// The same row for every stereo pair of the launch group the two handles hold resident, inputs and outputs on the device:
// what orbx_extract_batch(..., ORBX_MEM_DEVICE) wrote goes straight into the matcher, nothing is downloaded and re-uploaded.
int orbx_stereo_match_batch(OrbxHandle* h, OrbxHandle* right, int n_pairs, const OrbxKeyPoint* kps_l, const uint8_t* desc_l,
                            const int32_t* counts_l, const OrbxKeyPoint* kps_r, const uint8_t* desc_r, const int32_t* counts_r,
                            int cap_per_frame, float mb, float mbf, float* u_right, float* depth, int32_t* n_matched, void* stream) {
    if (!h || !right) return ORBX_ERR_BAD_ARGUMENT;
    if (n_pairs < 1 || cap_per_frame < 1 || cap_per_frame >= 65535 || !kps_l || !desc_l || !counts_l || !kps_r || !desc_r || !counts_r ||
        !u_right || !depth || !n_matched || !(mb > 0.f))
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "bad stereo arguments");
    if (!h->cur || !right->cur || h->resident_frames < n_pairs || right->resident_frames < n_pairs)
        return fail(h, ORBX_ERR_NO_FRAME, "both extractors must hold the pairs' pyramids resident (one launch group)");
    if (h->device != right->device || h->cur->plan.width != right->cur->plan.width || h->cur->plan.height != right->cur->plan.height ||
        h->cur->plan.nlevels != right->cur->plan.nlevels)
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "left and right extractors must share device, image size and level count");
    ORBX_CUDA(cudaSetDevice(h->device));
    int rc = ensure_bytes(h, (void**)&h->d_stereo, &h->d_stereo_bytes, (size_t)n_pairs * cap_per_frame * 4 + 256, false);
    if (rc != ORBX_OK) return rc;
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    OrbxStereoArgs a;
    a.kl = kps_l; a.dl = (const uint32_t*)desc_l; a.nl = 0; a.kr = kps_r; a.dr = (const uint32_t*)desc_r; a.nr = 0;
    a.kp_stride = cap_per_frame; a.pyr_stride = h->cur->pyr_stride; a.counts_l = counts_l; a.counts_r = counts_r; a.cap = cap_per_frame;
    a.pyr_l = res_ws(h).pyr; a.pyr_r = res_ws(right).pyr;
    a.mbf = mbf; a.min_d = 0.f; a.max_d = mbf / mb; a.th_mul = 1.5f * 1.4f;
    a.u_right = u_right; a.depth = depth; a.sad = (int*)h->d_stereo; a.n_matched = n_matched;
    k_stereo_match<<<dim3((cap_per_frame + 7) / 8, n_pairs), 256, 0, st>>>(h->cur->plan, a);
    k_stereo_filter<<<n_pairs, 256, 0, st>>>(a);
    h->total_launches += 2; h->stage_launches += 2;
    ORBX_CUDA(cudaGetLastError());
    return ORBX_OK;
}

// ---- Frame post-processing + SearchForInitialization (orbx_frame.cuh) ----------------------------------------
static bool calib_ok(const OrbxFrameCalib* c, bool need_bounds) {
    if (!c || !(c->fx != 0.f) || !(c->fy != 0.f) || c->n_dist < 0 || c->n_dist > 5) return false;
    if (need_bounds && (!(c->max_x > c->min_x) || !(c->max_y > c->min_y))) return false;
    return true;
}

int orbx_frame_image_bounds(OrbxHandle* h, OrbxFrameCalib* calib, int width, int height) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    if (!calib_ok(calib, false) || width <= 0 || height <= 0) return fail(h, ORBX_ERR_BAD_ARGUMENT, "bad calibration");
    if (calib->dist[0] == 0.0f) {                                        // src/Frame.cc:805-811
        calib->min_x = 0.0f; calib->max_x = (float)width; calib->min_y = 0.0f; calib->max_y = (float)height;
        return ORBX_OK;
    }
    ORBX_CUDA(cudaSetDevice(h->device));
    int rc = ensure_bytes(h, (void**)&h->d_stereo, &h->d_stereo_bytes, 256, false);
    if (rc != ORBX_OK) return rc;
    float m[8] = {0.f, 0.f, (float)width, 0.f, 0.f, (float)height, (float)width, (float)height};   // :788-791
    cudaStream_t st = h->stream;
    ORBX_CUDA(cudaMemcpyAsync(h->d_stereo, m, sizeof(m), cudaMemcpyHostToDevice, st));
    k_undistort_points<<<1, 128, 0, st>>>(*calib, (const float2*)h->d_stereo, (float2*)(h->d_stereo + 64), 4);
    h->total_launches += 1; h->stage_launches += 1;
    ORBX_CUDA(cudaGetLastError());
    ORBX_CUDA(cudaMemcpyAsync(m, h->d_stereo + 64, sizeof(m), cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaStreamSynchronize(st));
    calib->min_x = std::min(m[0], m[4]); calib->max_x = std::max(m[2], m[6]);                       // :798-802
    calib->min_y = std::min(m[1], m[3]); calib->max_y = std::max(m[5], m[7]);
    return ORBX_OK;
}

int orbx_frame_undistort_grid(OrbxHandle* h, const OrbxFrameCalib* calib, const OrbxKeyPoint* keys, int n, OrbxKeyPoint* keys_un,
                              int32_t* cell_start, int32_t* cell_items, int* n_in_grid) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    if (n_in_grid) *n_in_grid = 0;
    if (!calib_ok(calib, true) || n < 0 || !cell_start || (n > 0 && (!keys || !keys_un || !cell_items)))
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "bad frame-grid arguments");
    ORBX_CUDA(cudaSetDevice(h->device));
    const int nn = std::max(n, 1);
    const size_t o_k = 0, o_u = o_k + align_up((long long)nn * sizeof(OrbxKeyPoint), 256), o_c = o_u + align_up((long long)nn * sizeof(OrbxKeyPoint), 256);
    const size_t o_s = o_c + align_up((long long)nn * 4, 256), o_i = o_s + align_up((ORBX_GRID_CELLS + 1) * 4, 256), o_n = o_i + align_up((long long)nn * 4, 256);
    int rc = ensure_bytes(h, (void**)&h->d_stereo, &h->d_stereo_bytes, o_n + 256, false);
    if (rc != ORBX_OK) return rc;
    uint8_t* b = h->d_stereo;
    cudaStream_t st = h->stream;
    if (n > 0) ORBX_CUDA(cudaMemcpyAsync(b + o_k, keys, (size_t)n * sizeof(OrbxKeyPoint), cudaMemcpyHostToDevice, st));
    k_frame_undistort_grid<<<1, 1024, 0, st>>>(*calib, (const OrbxKeyPoint*)(b + o_k), n, nullptr, (OrbxKeyPoint*)(b + o_u), (int*)(b + o_c),
                                               (int*)(b + o_s), (int*)(b + o_i), (int*)(b + o_n));
    h->total_launches += 1; h->stage_launches += 1;
    ORBX_CUDA(cudaGetLastError());
    int placed = 0;
    if (n > 0) ORBX_CUDA(cudaMemcpyAsync(keys_un, b + o_u, (size_t)n * sizeof(OrbxKeyPoint), cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaMemcpyAsync(cell_start, b + o_s, (ORBX_GRID_CELLS + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (n > 0) ORBX_CUDA(cudaMemcpyAsync(cell_items, b + o_i, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaMemcpyAsync(&placed, b + o_n, 4, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaStreamSynchronize(st));
    if (n_in_grid) *n_in_grid = placed;
    return ORBX_OK;
}

// What the monocular Frame constructor does with one image (src/Frame.cc:307-347): ExtractORB, UndistortKeyPoints,
// AssignFeaturesToGrid -- the keypoints stay on the device between the extraction and the grid kernel.
int orbx_extract_frame(OrbxHandle* h, const uint8_t* image, int width, int height, size_t stride, int lap0, int lap1,
                       const OrbxFrameCalib* calib, OrbxKeyPoint* kps, uint8_t* desc, int capacity, int* n_out, int* mono_out,
                       OrbxKeyPoint* kps_un, int32_t* cell_start, int32_t* cell_items, int* n_in_grid) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    if (n_out) *n_out = 0;
    if (mono_out) *mono_out = 0;
    if (n_in_grid) *n_in_grid = 0;
    if (!calib_ok(calib, true) || capacity <= 0 || !kps || !desc || !kps_un || !cell_start || !cell_items)
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "bad frame arguments");
    ORBX_CUDA(cudaSetDevice(h->device));
    const size_t cap = (size_t)capacity;
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (size_t)align_up((long long)bytes, 256); return at; };
    const size_t o_k = take(cap * sizeof(OrbxKeyPoint)), o_d = take(cap * 32), o_cnt = take(8), o_u = take(cap * sizeof(OrbxKeyPoint));
    const size_t o_c = take(cap * 4), o_s = take((ORBX_GRID_CELLS + 1) * 4), o_i = take(cap * 4), o_n = take(4);
    int rc = ensure_bytes(h, (void**)&h->d_frame, &h->d_frame_bytes, o, false);
    if (rc != ORBX_OK) return rc;
    uint8_t* b = h->d_frame;
    rc = orbx_extract_batch(h, image, ORBX_MEM_HOST, 1, width, height, stride, stride * (size_t)height, lap0, lap1, (OrbxKeyPoint*)(b + o_k),
                            b + o_d, capacity, (int32_t*)(b + o_cnt), ORBX_MEM_DEVICE, nullptr);
    if (rc != ORBX_OK) return rc;
    cudaStream_t st = h->stream;
    k_frame_undistort_grid<<<1, 1024, 0, st>>>(*calib, (const OrbxKeyPoint*)(b + o_k), capacity, (const int*)(b + o_cnt), (OrbxKeyPoint*)(b + o_u),
                                               (int*)(b + o_c), (int*)(b + o_s), (int*)(b + o_i), (int*)(b + o_n));
    h->total_launches += 1; h->stage_launches += 1;
    ORBX_CUDA(cudaGetLastError());
    // one contiguous read-back into pinned memory (copies into pageable memory block the host one by one), then the
    // `n` valid entries of each array are handed to the caller
    rc = ensure_bytes(h, (void**)&h->h_frame, &h->h_frame_bytes, o, true);
    if (rc != ORBX_OK) return rc;
    ORBX_CUDA(cudaMemcpyAsync(h->h_frame, b, o, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaStreamSynchronize(st));
    const uint8_t* hb = h->h_frame;
    int32_t cnt[2];
    int placed = 0;
    std::memcpy(cnt, hb + o_cnt, 8);
    std::memcpy(&placed, hb + o_n, 4);
    const size_t n = (size_t)std::min(std::max(cnt[0], 0), capacity);
    std::memcpy(kps, hb + o_k, n * sizeof(OrbxKeyPoint));
    std::memcpy(desc, hb + o_d, n * 32);
    std::memcpy(kps_un, hb + o_u, n * sizeof(OrbxKeyPoint));
    std::memcpy(cell_start, hb + o_s, (ORBX_GRID_CELLS + 1) * 4);
    std::memcpy(cell_items, hb + o_i, (size_t)std::min(std::max(placed, 0), capacity) * 4);
    if (n_out) *n_out = cnt[0];
    if (mono_out) *mono_out = cnt[1];
    if (n_in_grid) *n_in_grid = placed;
    if (cnt[0] > capacity) return fail(h, ORBX_ERR_CAPACITY, "keypoint capacity too small");
    return ORBX_OK;
}

int orbx_search_for_initialization_mem(OrbxHandle* h, const OrbxFrameCalib* calib, const OrbxKeyPoint* keys_un1, const uint8_t* desc1, int n1,
                                       const OrbxKeyPoint* keys_un2, const uint8_t* desc2, int n2, const int32_t* cell_start2,
                                       const int32_t* cell_items2, float* prev_matched, int window_size, float nn_ratio,
                                       int check_orientation, int32_t* matches12, int* n_matches, int mem) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    if (n_matches) *n_matches = 0;
    if (!calib_ok(calib, true) || n1 < 0 || n2 < 0 || n2 > 32768 || !cell_start2 || (mem != ORBX_MEM_HOST && mem != ORBX_MEM_DEVICE) ||
        (n1 > 0 && (!keys_un1 || !desc1 || !prev_matched || !matches12)) || (n2 > 0 && (!keys_un2 || !desc2 || !cell_items2)))
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "bad SearchForInitialization arguments");
    if (n1 == 0) return ORBX_OK;
    const bool dev = mem == ORBX_MEM_DEVICE;
    int n_items = 0;
    if (!dev) {
        n_items = cell_start2[ORBX_GRID_CELLS];
        if (n_items < 0 || n_items > n2) return fail(h, ORBX_ERR_BAD_ARGUMENT, "grid does not belong to frame 2");
    }
    ORBX_CUDA(cudaSetDevice(h->device));
    const int m2 = std::max(n2, 1);
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o += (size_t)align_up((long long)bytes, 256); return at; };
    // device-resident inputs are used where they lie; only the kernels' scratch is allocated
    const size_t o_k1 = dev ? 0 : take((size_t)n1 * sizeof(OrbxKeyPoint)), o_d1 = dev ? 0 : take((size_t)n1 * 32);
    const size_t o_k2 = dev ? 0 : take((size_t)m2 * sizeof(OrbxKeyPoint)), o_d2 = dev ? 0 : take((size_t)m2 * 32);
    const size_t o_cs = dev ? 0 : take((ORBX_GRID_CELLS + 1) * 4), o_ci = dev ? 0 : take((size_t)m2 * 4), o_pv = dev ? 0 : take((size_t)n1 * 8);
    const size_t o_m = dev ? 0 : take((size_t)n1 * 4);
    const size_t o_sk = take((size_t)n1 * 32), o_si = take((size_t)n1 * 32), o_sc = take((size_t)n1 * 4);
    const size_t o_p = take((size_t)n1 * 4), o_al = take((size_t)n1 * 4), o_n = take(8);
    int rc = ensure_bytes(h, (void**)&h->d_stereo, &h->d_stereo_bytes, o, false);
    if (rc != ORBX_OK) return rc;
    uint8_t* b = h->d_stereo;
    cudaStream_t st = h->stream;
    if (!dev) {
        ORBX_CUDA(cudaMemcpyAsync(b + o_k1, keys_un1, (size_t)n1 * sizeof(OrbxKeyPoint), cudaMemcpyHostToDevice, st));
        ORBX_CUDA(cudaMemcpyAsync(b + o_d1, desc1, (size_t)n1 * 32, cudaMemcpyHostToDevice, st));
        if (n2 > 0) {
            ORBX_CUDA(cudaMemcpyAsync(b + o_k2, keys_un2, (size_t)n2 * sizeof(OrbxKeyPoint), cudaMemcpyHostToDevice, st));
            ORBX_CUDA(cudaMemcpyAsync(b + o_d2, desc2, (size_t)n2 * 32, cudaMemcpyHostToDevice, st));
            if (n_items > 0) ORBX_CUDA(cudaMemcpyAsync(b + o_ci, cell_items2, (size_t)n_items * 4, cudaMemcpyHostToDevice, st));
        }
        ORBX_CUDA(cudaMemcpyAsync(b + o_cs, cell_start2, (ORBX_GRID_CELLS + 1) * 4, cudaMemcpyHostToDevice, st));
        ORBX_CUDA(cudaMemcpyAsync(b + o_pv, prev_matched, (size_t)n1 * 8, cudaMemcpyHostToDevice, st));
    }
    OrbxInitArgs a;
    a.calib = *calib;
    a.k1 = dev ? keys_un1 : (const OrbxKeyPoint*)(b + o_k1); a.d1 = dev ? (const uint32_t*)desc1 : (const uint32_t*)(b + o_d1); a.n1 = n1;
    a.k2 = dev ? keys_un2 : (const OrbxKeyPoint*)(b + o_k2); a.d2 = dev ? (const uint32_t*)desc2 : (const uint32_t*)(b + o_d2); a.n2 = n2;
    a.cell_start2 = dev ? (const int*)cell_start2 : (const int*)(b + o_cs); a.cell_items2 = dev ? (const int*)cell_items2 : (const int*)(b + o_ci);
    a.prev = dev ? prev_matched : (float*)(b + o_pv); a.r = (float)window_size; a.nn_ratio = nn_ratio; a.check_orientation = check_orientation ? 1 : 0;
    a.sl_key = (uint4*)(b + o_sk); a.sl_idx = (uint4*)(b + o_si); a.sl_count = (int*)(b + o_sc);
    a.matches12 = dev ? (int*)matches12 : (int*)(b + o_m); a.pushed = (int*)(b + o_p); a.act_list = (int*)(b + o_al); a.n_matches = (int*)(b + o_n);
    const size_t smem = (size_t)m2 * 6 + 16;
    { const int ra = set_kernel_attrs_device(h); if (ra != ORBX_OK) return ra; }
    k_init_shortlist<<<(n1 + 7) / 8, 256, 0, st>>>(a);
    k_init_resolve<<<1, 256, smem, st>>>(a);
    h->total_launches += 2; h->stage_launches += 2;
    ORBX_CUDA(cudaGetLastError());
    int nm[2] = {0, 0};
    if (!dev) {
        ORBX_CUDA(cudaMemcpyAsync(matches12, b + o_m, (size_t)n1 * 4, cudaMemcpyDeviceToHost, st));
        ORBX_CUDA(cudaMemcpyAsync(prev_matched, b + o_pv, (size_t)n1 * 8, cudaMemcpyDeviceToHost, st));
    }
    ORBX_CUDA(cudaMemcpyAsync(nm, b + o_n, 8, cudaMemcpyDeviceToHost, st));
    ORBX_CUDA(cudaStreamSynchronize(st));
    if (n_matches) *n_matches = nm[0];
    h->last_init_fallbacks = nm[1];
    return ORBX_OK;
}

int orbx_search_for_initialization(OrbxHandle* h, const OrbxFrameCalib* calib, const OrbxKeyPoint* keys_un1, const uint8_t* desc1, int n1,
                                   const OrbxKeyPoint* keys_un2, const uint8_t* desc2, int n2, const int32_t* cell_start2,
                                   const int32_t* cell_items2, float* prev_matched, int window_size, float nn_ratio,
                                   int check_orientation, int32_t* matches12, int* n_matches) {
    return orbx_search_for_initialization_mem(h, calib, keys_un1, desc1, n1, keys_un2, desc2, n2, cell_start2, cell_items2, prev_matched,
                                              window_size, nn_ratio, check_orientation, matches12, n_matches, ORBX_MEM_HOST);
}

// ---- CLAHE (orbx_clahe.cuh) ----
int orbx_clahe(OrbxHandle* h, const uint8_t* images, int in_mem, int n_frames, int width, int height, size_t row_stride,
               size_t frame_stride, double clip_limit, int tiles_x, int tiles_y, uint8_t* out, int out_mem, size_t out_row_stride,
               size_t out_frame_stride, void* stream) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    if (n_frames < 0 || width <= 0 || height <= 0 || tiles_x <= 0 || tiles_y <= 0 || tiles_x > 64 || tiles_y > 4096 ||
        (n_frames > 0 && (!images || !out)) || row_stride < (size_t)width || out_row_stride < (size_t)width ||
        (n_frames > 1 && (frame_stride < row_stride * (size_t)(height - 1) + width || out_frame_stride < out_row_stride * (size_t)(height - 1) + width)))
        return fail(h, ORBX_ERR_BAD_ARGUMENT, "bad CLAHE arguments");
    if (n_frames == 0) return ORBX_OK;
    ORBX_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    // geometry exactly as cv::CLAHE::apply: extend to multiples of the tile grid (for the LUTs only)
    int ew = width, eh = height;
    if (!(width % tiles_x == 0 && height % tiles_y == 0)) { ew = width + (tiles_x - width % tiles_x); eh = height + (tiles_y - height % tiles_y); }
    OrbxClaheArgs a;
    a.w = width; a.h = height; a.tiles_x = tiles_x; a.tiles_y = tiles_y; a.tw = ew / tiles_x; a.th = eh / tiles_y;
    const int area = a.tw * a.th;
    a.lut_scale = (float)(256 - 1) / area;
    a.clip = 0;
    if (clip_limit > 0.0) a.clip = std::max((int)(clip_limit * area / 256), 1);
    a.inv_tw = 1.0f / a.tw; a.inv_th = 1.0f / a.th;
    // frames per launch group: input + output of a group stay in L2 between the two kernels
    const size_t frame_bytes = (size_t)width * height;
    int G = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_frames, (40u << 20) / frame_bytes));
    G = std::min(G, 65535);
    const bool host_in = in_mem == ORBX_MEM_HOST, host_out = out_mem == ORBX_MEM_HOST;
    const size_t lut_bytes = (size_t)G * tiles_x * tiles_y * 256;
    const size_t o_lut = 0, o_in = align_up((long long)lut_bytes, 256), o_out = o_in + (host_in ? align_up((long long)(G * frame_bytes), 256) : 0);
    const size_t need = o_out + (host_out ? align_up((long long)(G * frame_bytes), 256) : 0);
    int rc = ensure_bytes(h, (void**)&h->d_clahe, &h->d_clahe_bytes, need, false);
    if (rc != ORBX_OK) return rc;
    a.lut = h->d_clahe + o_lut;
    // two rows of tile LUTs + per-column interpolation terms (xa as float, two LUT bases as 16-bit halves)
    const size_t smem = (size_t)2 * tiles_x * 256 + (size_t)((width + 3) & ~3) * 8;
    if (width > 16384) return fail(h, ORBX_ERR_IMAGE_TOO_LARGE, "CLAHE: image wider than 16384 pixels");
    { const int ra = set_kernel_attrs_device(h); if (ra != ORBX_OK) return ra; }
    for (int f0 = 0; f0 < n_frames; f0 += G) {
        const int nf = std::min(G, n_frames - f0);
        if (host_in) {
            ORBX_CUDA(cudaMemcpy2DAsync(h->d_clahe + o_in, width, images + (size_t)f0 * frame_stride, row_stride, width,
                                        frame_stride == row_stride * (size_t)height ? (size_t)height * nf : (size_t)height, cudaMemcpyHostToDevice, st));
            if (frame_stride != row_stride * (size_t)height)
                for (int f = 1; f < nf; ++f)
                    ORBX_CUDA(cudaMemcpy2DAsync(h->d_clahe + o_in + (size_t)f * frame_bytes, width, images + (size_t)(f0 + f) * frame_stride, row_stride,
                                                width, height, cudaMemcpyHostToDevice, st));
            a.src = h->d_clahe + o_in; a.src_row = width; a.src_frame = (long long)frame_bytes;
        } else {
            a.src = images + (size_t)f0 * frame_stride; a.src_row = (long long)row_stride; a.src_frame = (long long)frame_stride;
        }
        if (host_out) { a.dst = h->d_clahe + o_out; a.dst_row = width; a.dst_frame = (long long)frame_bytes; }
        else { a.dst = out + (size_t)f0 * out_frame_stride; a.dst_row = (long long)out_row_stride; a.dst_frame = (long long)out_frame_stride; }
        k_clahe_lut<<<dim3(tiles_x * tiles_y, nf), 256, 0, st>>>(a);
        k_clahe_apply<<<dim3(tiles_y + 1, nf), 256, smem, st>>>(a);
        h->total_launches += 2; h->stage_launches += 2;
        ORBX_CUDA(cudaGetLastError());
        if (host_out) {
            if (out_frame_stride == out_row_stride * (size_t)height)
                ORBX_CUDA(cudaMemcpy2DAsync(out + (size_t)f0 * out_frame_stride, out_row_stride, h->d_clahe + o_out, width, width, (size_t)height * nf,
                                            cudaMemcpyDeviceToHost, st));
            else
                for (int f = 0; f < nf; ++f)
                    ORBX_CUDA(cudaMemcpy2DAsync(out + (size_t)(f0 + f) * out_frame_stride, out_row_stride, h->d_clahe + o_out + (size_t)f * frame_bytes, width,
                                                width, height, cudaMemcpyDeviceToHost, st));
        }
        if ((host_in || host_out) && f0 + G < n_frames) ORBX_CUDA(cudaStreamSynchronize(st));   // staging buffers are reused by the next group
    }
    if (host_in || host_out) ORBX_CUDA(cudaStreamSynchronize(st));
    return ORBX_OK;
}

int orbx_last_init_fallbacks(const OrbxHandle* h) { return h ? h->last_init_fallbacks : 0; }

int orbx_get_pyramid_layout(OrbxHandle* h, int width, int height, size_t* frame_bytes, size_t* plane_offset, int32_t* pitch,
                            int32_t* level_w, int32_t* level_h) {
    if (!h || width <= 0 || height <= 0) return ORBX_ERR_BAD_ARGUMENT;
    ORBX_CUDA(cudaSetDevice(h->device));
    PlanEntry* pe = nullptr;
    int rc = get_plan(h, width, height, &pe);
    if (rc != ORBX_OK) return rc;
    if (frame_bytes) *frame_bytes = (size_t)pe->pyr_stride;
    for (int l = 0; l < pe->plan.nlevels; ++l) {
        const OrbxLevel& V = pe->plan.lv[l];
        if (plane_offset) plane_offset[l] = (size_t)V.plane_off;
        if (pitch) pitch[l] = V.pitch;
        if (level_w) level_w[l] = V.w;
        if (level_h) level_h[l] = V.h;
    }
    return ORBX_OK;
}

int orbx_set_pyramid_output(OrbxHandle* h, uint8_t* host_dst, size_t frame_stride) {
    if (!h || (host_dst && frame_stride == 0)) return ORBX_ERR_BAD_ARGUMENT;
    h->pyr_out = host_dst;
    h->pyr_out_stride = host_dst ? frame_stride : 0;
    return ORBX_OK;
}

int orbx_download_pyramid(OrbxHandle* h, int frame, uint8_t* host_dst, size_t capacity) {
    if (!h || !host_dst) return ORBX_ERR_BAD_ARGUMENT;
    if (!h->cur || frame < 0 || frame >= h->resident_frames) return fail(h, ORBX_ERR_NO_FRAME, "frame not resident");
    if (capacity < (size_t)h->cur->pyr_stride) return fail(h, ORBX_ERR_CAPACITY, "pyramid buffer too small (orbx_get_pyramid_layout)");
    ORBX_CUDA(cudaSetDevice(h->device));
    { const int rb = ensure_borders(h); if (rb != ORBX_OK) return rb; }
    ORBX_CUDA(cudaMemcpyAsync(host_dst, res_ws(h).pyr + (size_t)frame * res_ws(h).pyr_stride, (size_t)h->cur->pyr_stride, cudaMemcpyDeviceToHost,
                              h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

void* orbx_host_alloc(size_t bytes, int write_combined) {
    void* p = nullptr;
    if (bytes == 0) return nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void orbx_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int orbx_get_level_size(const OrbxHandle* h, int level, int* width, int* height) {
    if (!h || !h->cur) return ORBX_ERR_NO_FRAME;
    if (level < 0 || level >= h->cur->plan.nlevels) return ORBX_ERR_BAD_ARGUMENT;
    if (width) *width = h->cur->plan.lv[level].w;
    if (height) *height = h->cur->plan.lv[level].h;
    return ORBX_OK;
}

static int check_frame(OrbxHandle* h, int frame, int level) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    if (!h->cur || frame < 0 || frame >= h->resident_frames) return fail(h, ORBX_ERR_NO_FRAME, "frame not resident");
    if (level < 0 || level >= h->cur->plan.nlevels) return fail(h, ORBX_ERR_BAD_ARGUMENT, "bad level");
    return ORBX_OK;
}

int orbx_get_pyramid_level(OrbxHandle* h, int frame, int level, uint8_t* dst, size_t dst_stride, int with_border) {
    int rc = check_frame(h, frame, level);
    if (rc != ORBX_OK) return rc;
    if (!dst) return fail(h, ORBX_ERR_BAD_ARGUMENT, "null destination");
    ORBX_CUDA(cudaSetDevice(h->device));
    const OrbxLevel& V = h->cur->plan.lv[level];
    const int b = with_border ? ORBX_EDGE : 0;
    const uint8_t* src = res_ws(h).pyr + (size_t)frame * res_ws(h).pyr_stride + V.plane_off + (size_t)(ORBX_EDGE - b) * V.pitch + (ORBX_PADL - b);
    if (with_border) { rc = ensure_borders(h); if (rc != ORBX_OK) return rc; }
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    ORBX_CUDA(cudaMemcpy2D(dst, dst_stride, src, (size_t)V.pitch, (size_t)V.w + 2 * b, (size_t)V.h + 2 * b, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbx_get_blurred_level(OrbxHandle* h, int frame, int level, uint8_t* dst, size_t dst_stride) {
    int rc = check_frame(h, frame, level);
    if (rc != ORBX_OK) return rc;
    if (!dst) return fail(h, ORBX_ERR_BAD_ARGUMENT, "null destination");
    ORBX_CUDA(cudaSetDevice(h->device));
    const OrbxLevel& V = h->cur->plan.lv[level];
    const uint8_t* src = res_ws(h).blur + (size_t)frame * res_ws(h).blur_stride + V.blur_off;
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    ORBX_CUDA(cudaMemcpy2D(dst, dst_stride, src, (size_t)V.blur_pitch, (size_t)V.w, (size_t)V.h, cudaMemcpyDeviceToHost));
    return ORBX_OK;
}

int orbx_get_level_keypoints(OrbxHandle* h, int frame, int level, OrbxKeyPoint* kps, int capacity, int* n_out) {
    int rc = check_frame(h, frame, level);
    if (rc != ORBX_OK) return rc;
    ORBX_CUDA(cudaSetDevice(h->device));
    const OrbxPlan& P = h->cur->plan;
    const OrbxLevel& V = P.lv[level];
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    int2 lc;
    ORBX_CUDA(cudaMemcpy(&lc, res_ws(h).level_count + (size_t)frame * P.nlevels + level, sizeof(int2), cudaMemcpyDeviceToHost));
    if (n_out) *n_out = lc.x;
    const int n = std::min(lc.x, capacity);
    if (n > 0 && kps) {
        std::vector<OrbxKpRec> rec((size_t)n);
        ORBX_CUDA(cudaMemcpy(rec.data(), res_ws(h).kprec + (size_t)frame * res_ws(h).kp_stride + V.kp_off, (size_t)n * sizeof(OrbxKpRec), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; ++i) {
            kps[i].x = rec[i].x; kps[i].y = rec[i].y; kps[i].size = V.kp_size; kps[i].angle = rec[i].angle;
            kps[i].response = rec[i].response; kps[i].octave = level; kps[i].class_id = -1;
        }
    }
    return ORBX_OK;
}

// allKeypoints of every level with one read-back (the 6-argument operator() hands them all to its caller, :1094): the frame's
// whole keypoint-record block and its level counts travel in two copies into pinned memory instead of two per level.
int orbx_get_all_level_keypoints(OrbxHandle* h, int frame, OrbxKeyPoint* kps, int capacity, int32_t* counts, int* n_total) {
    int rc = check_frame(h, frame, 0);
    if (rc != ORBX_OK) return rc;
    if (n_total) *n_total = 0;
    ORBX_CUDA(cudaSetDevice(h->device));
    const OrbxPlan& P = h->cur->plan;
    const size_t rec_bytes = (size_t)P.kp_total * sizeof(OrbxKpRec), cnt_bytes = (size_t)P.nlevels * sizeof(int2);
    const size_t o_cnt = (size_t)align_up((long long)rec_bytes, 256);
    rc = ensure_bytes(h, (void**)&h->h_frame, &h->h_frame_bytes, o_cnt + cnt_bytes, true);
    if (rc != ORBX_OK) return rc;
    const OrbxWs& w = res_ws(h);
    ORBX_CUDA(cudaMemcpyAsync(h->h_frame, w.kprec + (size_t)frame * w.kp_stride, rec_bytes, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaMemcpyAsync(h->h_frame + o_cnt, w.level_count + (size_t)frame * P.nlevels, cnt_bytes, cudaMemcpyDeviceToHost, h->stream));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    const OrbxKpRec* rec = reinterpret_cast<const OrbxKpRec*>(h->h_frame);
    const int2* lc = reinterpret_cast<const int2*>(h->h_frame + o_cnt);
    int total = 0;
    for (int l = 0; l < P.nlevels; ++l) {
        const OrbxLevel& V = P.lv[l];
        const int n = std::min(std::max(lc[l].x, 0), V.kp_cap);
        if (counts) counts[l] = n;
        for (int i = 0; i < n; ++i, ++total) {
            if (!kps || total >= capacity) continue;
            const OrbxKpRec& r = rec[V.kp_off + i];
            OrbxKeyPoint& k = kps[total];
            k.x = r.x; k.y = r.y; k.size = V.kp_size; k.angle = r.angle; k.response = r.response; k.octave = l; k.class_id = -1;
        }
    }
    if (n_total) *n_total = total;
    if (kps && total > capacity) return fail(h, ORBX_ERR_CAPACITY, "keypoint capacity too small");
    return ORBX_OK;
}

int orbx_get_level_candidates(OrbxHandle* h, int frame, int level, int32_t* xs, int32_t* ys, int32_t* scores, uint32_t* order,
                              int capacity, int* n_out) {
    int rc = check_frame(h, frame, level);
    if (rc != ORBX_OK) return rc;
    ORBX_CUDA(cudaSetDevice(h->device));
    const OrbxPlan& P = h->cur->plan;
    const OrbxLevel& V = P.lv[level];
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    int cnt = 0;
    ORBX_CUDA(cudaMemcpy(&cnt, res_ws(h).cand_count + (size_t)frame * P.nlevels + level, sizeof(int), cudaMemcpyDeviceToHost));
    cnt = std::min(cnt, V.cand_cap);
    if (n_out) *n_out = cnt;
    const int n = std::min(cnt, capacity);
    if (n > 0) {
        std::vector<uint2> c((size_t)n);
        ORBX_CUDA(cudaMemcpy(c.data(), res_ws(h).cand + (size_t)frame * res_ws(h).cand_stride + V.cand_off, (size_t)n * sizeof(uint2), cudaMemcpyDeviceToHost));
        const uint32_t CM = (1u << ORBX_COORD_BITS) - 1;
        for (int i = 0; i < n; ++i) {
            if (xs) xs[i] = (int32_t)(c[i].x & CM);
            if (ys) ys[i] = (int32_t)((c[i].x >> ORBX_COORD_BITS) & CM);
            if (scores) scores[i] = (int32_t)(c[i].x >> 24);
            if (order) order[i] = c[i].y;
        }
    }
    return ORBX_OK;
}

int orbx_stage_times(OrbxHandle* h, float* ms, int64_t* launches) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    for (int s = 0; s < ORBX_NUM_STAGES; ++s) { if (ms) ms[s] = (float)h->stage_ms[s]; h->stage_ms[s] = 0; }
    if (launches) *launches = h->stage_launches;
    h->stage_launches = 0;
    return ORBX_OK;
}

int64_t orbx_launch_count(const OrbxHandle* h) { return h ? h->total_launches : 0; }

int orbx_synchronize(OrbxHandle* h) {
    if (!h) return ORBX_ERR_BAD_ARGUMENT;
    ORBX_CUDA(cudaSetDevice(h->device));
    ORBX_CUDA(cudaStreamSynchronize(h->stream));
    return ORBX_OK;
}

void* orbx_get_stream(const OrbxHandle* h) { return h ? (void*)h->stream : nullptr; }

int orbx_uses_tma(const OrbxHandle* h) { return h && h->ws_plan && h->ws.tmaps && h->ws.tmaps_blur && h->ws.tmaps_b7 ? 1 : 0; }

}  // extern "C"
