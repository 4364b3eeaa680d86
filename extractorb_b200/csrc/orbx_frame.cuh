// extractorb_b200/csrc/orbx_frame.cuh -- sm_100a kernels for the rows that follow the extractor in every Frame
// constructor and in monocular initialisation:
//
//   k_frame_undistort_grid   Frame::UndistortKeyPoints (reference src/Frame.cc:748-782, cv::undistortPoints in
//                            double precision) + Frame::AssignFeaturesToGrid / PosInGrid (:383-417, :726-736)
//   k_undistort_points       the same point arithmetic on bare (x, y) pairs: Frame::ComputeImageBounds (:784-812)
//   k_init_shortlist         Frame::GetFeaturesInArea (:655-724) + ORBmatcher::DescriptorDistance
//                            (src/ORBmatcher.cc:2349-2365) for every level-0 keypoint of the first frame
//   k_init_resolve           the order-dependent part of ORBmatcher::SearchForInitialization (:705-814) and its
//                            rotation-consistency filter (ComputeThreeMaxima :2303-2344)
//
// Every float / double step is written with explicit round-to-nearest intrinsics (no FMA contraction): the
// reference is compiled without FMA (CMakeLists.txt:4-5), and results are compared bit for bit.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orbx.h"

#define ORBX_GRID_COLS ORBX_FRAME_GRID_COLS
#define ORBX_GRID_ROWS ORBX_FRAME_GRID_ROWS
#define ORBX_GRID_CELLS (ORBX_GRID_COLS * ORBX_GRID_ROWS)
#define ORBX_MATCH_TH_LOW 50        // ORBmatcher::TH_LOW, src/ORBmatcher.cc:37
#define ORBX_MATCH_HISTO 30         // ORBmatcher::HISTO_LENGTH, :38
#define ORBX_INIT_TAGS 2048
#define ORBX_SHORT_K 8              // shortlist length per keypoint (two uint4 of keys, two of indices)

// cv::undistortPoints for one point, K == P, no rectification: 5 fixed-point iterations (the default
// TermCriteria(MAX_ITER, 5, 0.01) tests the count only), every operation rounded separately.
__device__ __forceinline__ float2 undistort_point(float u, float v, const OrbxFrameCalib& c) {
    double k[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) k[i] = 0.0;
    for (int i = 0; i < c.n_dist && i < 5; ++i) k[i] = (double)c.dist[i];
    const double fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy;
    const double ifx = __ddiv_rn(1.0, fx), ify = __ddiv_rn(1.0, fy);
    double x = __dmul_rn(__dsub_rn((double)u, cx), ifx), y = __dmul_rn(__dsub_rn((double)v, cy), ify);
    const double x0 = x, y0 = y;
    if (c.n_dist > 0) {
        for (int j = 0; j < 5; ++j) {
            const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y);
            const double r2 = __dadd_rn(xx, yy);
            // (1 + ((k7 r2 + k6) r2 + k5) r2) / (1 + ((k4 r2 + k1) r2 + k0) r2)
            const double num = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[7], r2), k[6]), r2), k[5]), r2));
            const double den = __dadd_rn(1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(k[4], r2), k[1]), r2), k[0]), r2));
            const double icdist = __ddiv_rn(num, den);
            if (icdist < 0) { x = x0; y = y0; break; }
            // deltaX = 2 k2 x y + k3 (r2 + 2 x x) + k8 r2 + k9 r2 r2
            const double two_x = __dmul_rn(2.0, x), two_y = __dmul_rn(2.0, y);
            double dX = __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[2]), x), y);
            dX = __dadd_rn(dX, __dmul_rn(k[3], __dadd_rn(r2, __dmul_rn(two_x, x))));
            dX = __dadd_rn(dX, __dmul_rn(k[8], r2));
            dX = __dadd_rn(dX, __dmul_rn(__dmul_rn(k[9], r2), r2));
            // deltaY = k2 (r2 + 2 y y) + 2 k3 x y + k10 r2 + k11 r2 r2
            double dY = __dmul_rn(k[2], __dadd_rn(r2, __dmul_rn(two_y, y)));
            dY = __dadd_rn(dY, __dmul_rn(__dmul_rn(__dmul_rn(2.0, k[3]), x), y));
            dY = __dadd_rn(dY, __dmul_rn(k[10], r2));
            dY = __dadd_rn(dY, __dmul_rn(__dmul_rn(k[11], r2), r2));
            x = __dmul_rn(__dsub_rn(x0, dX), icdist);
            y = __dmul_rn(__dsub_rn(y0, dY), icdist);
        }
    }
    // [xx yy ww] = P [x y 1], P = [fx 0 cx; 0 fy cy; 0 0 1]
    const double px = __dadd_rn(__dadd_rn(__dmul_rn(fx, x), __dmul_rn(0.0, y)), cx);
    const double py = __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(fy, y)), cy);
    const double ww = __ddiv_rn(1.0, __dadd_rn(__dadd_rn(__dmul_rn(0.0, x), __dmul_rn(0.0, y)), 1.0));
    return make_float2((float)__dmul_rn(px, ww), (float)__dmul_rn(py, ww));
}

__global__ void __launch_bounds__(128)
k_undistort_points(const OrbxFrameCalib calib, const float2* __restrict__ in, float2* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = undistort_point(in[i].x, in[i].y, calib);
}

// mfGridElementWidthInv / HeightInv (src/Frame.cc:339-340)
__device__ __forceinline__ float grid_w_inv(const OrbxFrameCalib& c) { return __fdiv_rn((float)ORBX_GRID_COLS, __fsub_rn(c.max_x, c.min_x)); }
__device__ __forceinline__ float grid_h_inv(const OrbxFrameCalib& c) { return __fdiv_rn((float)ORBX_GRID_ROWS, __fsub_rn(c.max_y, c.min_y)); }

// One CTA per frame: undistort every keypoint, bin it (PosInGrid), and emit mGrid as CSR with the items of a
// cell in ascending keypoint index (the reference push_backs in index order).
//   cell_start[ix * 48 + iy .. +1] delimit mGrid[ix][iy] inside cell_items.
__global__ void __launch_bounds__(1024)
k_frame_undistort_grid(const OrbxFrameCalib calib, const OrbxKeyPoint* __restrict__ keys, int n, const int* __restrict__ n_dev,
                       OrbxKeyPoint* __restrict__ keys_un, int* __restrict__ cell_of, int* __restrict__ cell_start,
                       int* __restrict__ cell_items, int* __restrict__ n_in_grid) {
    // n_dev: the keypoint count still lives on the device (the extraction that produced `keys` has not been read back)
    if (n_dev) n = min(max(*n_dev, 0), n);
    __shared__ int s_cnt[ORBX_GRID_CELLS];
    __shared__ int s_start[ORBX_GRID_CELLS + 1];
    __shared__ int s_part[32];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int c = tid; c < ORBX_GRID_CELLS; c += nt) s_cnt[c] = 0;
    __syncthreads();
    const bool distorted = calib.dist[0] != 0.0f;                          // :750
    const float wInv = grid_w_inv(calib), hInv = grid_h_inv(calib);
    for (int i = tid; i < n; i += nt) {
        OrbxKeyPoint kp = keys[i];
        if (distorted) {
            const float2 p = undistort_point(kp.x, kp.y, calib);
            kp.x = p.x; kp.y = p.y;
        }
        keys_un[i] = kp;
        // PosInGrid: round() is half away from zero
        const int px = (int)roundf(__fmul_rn(__fsub_rn(kp.x, calib.min_x), wInv));
        const int py = (int)roundf(__fmul_rn(__fsub_rn(kp.y, calib.min_y), hInv));
        int c = -1;
        if (px >= 0 && px < ORBX_GRID_COLS && py >= 0 && py < ORBX_GRID_ROWS) {
            c = px * ORBX_GRID_ROWS + py;
            atomicAdd(&s_cnt[c], 1);
        }
        cell_of[i] = c;
    }
    __syncthreads();
    // exclusive scan of the 3072 counts: 3 cells per thread, warp scan, scan of warp totals
    {
        const int b = tid * 3;
        int v0 = 0, v1 = 0, v2 = 0;
        if (b < ORBX_GRID_CELLS) { v0 = s_cnt[b]; v1 = s_cnt[b + 1]; v2 = s_cnt[b + 2]; }
        const int tot = v0 + v1 + v2;
        int inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if ((tid & 31) >= o) inc += t;
        }
        if ((tid & 31) == 31) s_part[tid >> 5] = inc;
        __syncthreads();
        if (tid < 32) {
            int p = s_part[tid], pi = p;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, pi, o);
                if (tid >= o) pi += t;
            }
            s_part[tid] = pi - p;
        }
        __syncthreads();
        const int ex = s_part[tid >> 5] + inc - tot;
        if (b < ORBX_GRID_CELLS) { s_start[b] = ex; s_start[b + 1] = ex + v0; s_start[b + 2] = ex + v0 + v1; }
        if (b + 3 == ORBX_GRID_CELLS) s_start[ORBX_GRID_CELLS] = ex + tot;
    }
    __syncthreads();
    for (int c = tid; c < ORBX_GRID_CELLS; c += nt) s_cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const int c = cell_of[i];
        if (c >= 0) cell_items[s_start[c] + atomicAdd(&s_cnt[c], 1)] = i;
    }
    __syncthreads();
    // ascending index inside each cell (segments are tiny: insertion sort by the cell's thread)
    for (int c = tid; c < ORBX_GRID_CELLS; c += nt) {
        const int s = s_start[c], e = s_start[c + 1];
        for (int a = s + 1; a < e; ++a) {
            const int v = cell_items[a];
            int b = a - 1;
            while (b >= s && cell_items[b] > v) { cell_items[b + 1] = cell_items[b]; --b; }
            cell_items[b + 1] = v;
        }
    }
    for (int c = tid; c <= ORBX_GRID_CELLS; c += nt) cell_start[c] = s_start[c];
    if (tid == 0) *n_in_grid = s_start[ORBX_GRID_CELLS];
}

// ---------------------------------------------------------------------------------------------------------------
// SearchForInitialization
// ---------------------------------------------------------------------------------------------------------------
struct OrbxInitArgs {
    OrbxFrameCalib calib;
    const OrbxKeyPoint* k1; const uint32_t* d1; int n1;
    const OrbxKeyPoint* k2; const uint32_t* d2; int n2;
    const int* cell_start2; const int* cell_items2;
    float* prev;            // n1 x 2, in/out (vbPrevMatched)
    float r;                // (float)windowSize
    float nn_ratio;
    int check_orientation;
    // shortlist: for keypoint i1 up to ORBX_SHORT_K smallest (distance << 22 | enumeration position) candidates in
    // ascending order (0xffffffff = none) and the number of candidates enumerated; a list shorter than
    // min(count, ORBX_SHORT_K) means the warp merge could not prove the next entry (see init_enumerate)
    uint4* sl_key; uint4* sl_idx; int* sl_count;
    int* matches12;         // n1
    int* pushed;            // n1: bestIdx2 at the time i1 was matched (never reset), -1 otherwise
    int* act_list;          // n1: scratch, the keypoints k_init_resolve has to visit, ascending
    int* n_matches;
};

struct Top4 {
    unsigned k0, k1, k2, k3;
    int j0, j1, j2, j3;
    bool dropped;   // an entry fell off the end: once this list runs empty its next key is unknown
    __device__ __forceinline__ void init() { k0 = k1 = k2 = k3 = 0xffffffffu; j0 = j1 = j2 = j3 = -1; dropped = false; }
    __device__ __forceinline__ void insert(unsigned key, int j) {
        dropped = dropped || k3 != 0xffffffffu;
        if (key < k3) {
            k3 = key; j3 = j;
            if (k3 < k2) { unsigned t = k2; k2 = k3; k3 = t; int u = j2; j2 = j3; j3 = u; }
            if (k2 < k1) { unsigned t = k1; k1 = k2; k2 = t; int u = j1; j1 = j2; j2 = u; }
            if (k1 < k0) { unsigned t = k0; k0 = k1; k1 = t; int u = j0; j0 = j1; j1 = u; }
        }
    }
    __device__ __forceinline__ void pop() { k0 = k1; j0 = j1; k1 = k2; j1 = j2; k2 = k3; j2 = j3; k3 = 0xffffffffu; j3 = -1; }
};

// GetFeaturesInArea(x, y, r, 0, 0) of frame 2 in the reference's enumeration order (ix ascending, iy ascending,
// push order inside a cell) + DescriptorDistance against descriptor `dq`, by one warp.  A column of cells
// mGrid[ix][iyMin..iyMax] is one contiguous CSR segment.  With FILTER, candidates whose current
// vMatchedDistance is <= their distance are dropped (src/ORBmatcher.cc:744).  Every lane returns the warp-wide
// smallest (distance << 22 | position) keys in ascending order -- up to ORBX_SHORT_K of them; the merge of the per-lane
// top-4 lists stops early when a lane that had to drop entries runs empty (its next key is unknown), so every key
// returned is exactly the next smallest.  `count` = candidates enumerated (after the filter, if any).
template <bool FILTER>
__device__ __forceinline__ void init_enumerate(const OrbxInitArgs& a, float x, float y, const uint32_t (&dq)[8], const uint16_t* vmd,
                                               unsigned (&okey)[ORBX_SHORT_K], int (&oidx)[ORBX_SHORT_K], int& count) {
    const int lane = threadIdx.x & 31;
    const OrbxFrameCalib& c = a.calib;
    const float r = a.r;
    const float wInv = grid_w_inv(c), hInv = grid_h_inv(c);
    count = 0;
#pragma unroll
    for (int q = 0; q < ORBX_SHORT_K; ++q) { okey[q] = 0xffffffffu; oidx[q] = -1; }
    const float xr = __fsub_rn(x, c.min_x), yr = __fsub_rn(y, c.min_y);
    const int minCX = max(0, (int)floorf(__fmul_rn(__fsub_rn(xr, r), wInv)));
    if (minCX >= ORBX_GRID_COLS) return;
    const int maxCX = min(ORBX_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(xr, r), wInv)));
    if (maxCX < 0) return;
    const int minCY = max(0, (int)floorf(__fmul_rn(__fsub_rn(yr, r), hInv)));
    if (minCY >= ORBX_GRID_ROWS) return;
    const int maxCY = min(ORBX_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(yr, r), hInv)));
    if (maxCY < 0) return;
    Top4 t;
    t.init();
    int running = 0;
    // the CSR ranges of all (at most 64) cell columns are fetched at once: lane c holds columns c and c + 32
    const int ncol = maxCX - minCX + 1;
    int cs0 = 0, ce0 = 0, cs1 = 0, ce1 = 0;
    if (lane < ncol) {
        cs0 = __ldg(a.cell_start2 + (minCX + lane) * ORBX_GRID_ROWS + minCY);
        ce0 = __ldg(a.cell_start2 + (minCX + lane) * ORBX_GRID_ROWS + maxCY + 1);
    }
    if (lane + 32 < ncol) {
        cs1 = __ldg(a.cell_start2 + (minCX + lane + 32) * ORBX_GRID_ROWS + minCY);
        ce1 = __ldg(a.cell_start2 + (minCX + lane + 32) * ORBX_GRID_ROWS + maxCY + 1);
    }
    for (int c = 0; c < ncol; ++c) {
        const int s0 = __shfl_sync(0xffffffffu, cs0, c & 31), e0 = __shfl_sync(0xffffffffu, ce0, c & 31);
        const int s1 = __shfl_sync(0xffffffffu, cs1, c & 31), e1 = __shfl_sync(0xffffffffu, ce1, c & 31);
        const int s = c < 32 ? s0 : s1, e = c < 32 ? e0 : e1;
        for (int base = s; base < e; base += 32) {
            const int p = base + lane;
            bool ok = false;
            int i2 = -1;
            if (p < e) {
                i2 = __ldg(a.cell_items2 + p);
                const OrbxKeyPoint* kp = a.k2 + i2;
                const int oct = __ldg(&kp->octave);
                const float dx = __fsub_rn(__ldg(&kp->x), x), dy = __fsub_rn(__ldg(&kp->y), y);
                ok = oct == 0 && fabsf(dx) < r && fabsf(dy) < r;          // bCheckLevels with minLevel = maxLevel = 0
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int pos = running + __popc(bal & ((1u << lane) - 1u));
                const uint4* q = reinterpret_cast<const uint4*>(a.d2 + (size_t)i2 * 8);
                const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
                const int dist = __popc(q0.x ^ dq[0]) + __popc(q0.y ^ dq[1]) + __popc(q0.z ^ dq[2]) + __popc(q0.w ^ dq[3]) +
                                 __popc(q1.x ^ dq[4]) + __popc(q1.y ^ dq[5]) + __popc(q1.z ^ dq[6]) + __popc(q1.w ^ dq[7]);
                bool keep = true;
                if (FILTER) keep = !((int)vmd[i2] <= dist);
                if (keep) t.insert(((unsigned)dist << 22) | (unsigned)pos, i2);
            }
            running += __popc(bal);
        }
    }
    count = running;
    // warp-wide merge of the per-lane sorted lists
#pragma unroll
    for (int q = 0; q < ORBX_SHORT_K; ++q) {
        if (__any_sync(0xffffffffu, t.dropped && t.k0 == 0xffffffffu)) break;
        const unsigned m = __reduce_min_sync(0xffffffffu, t.k0);
        if (m == 0xffffffffu) break;
        const unsigned own = __ballot_sync(0xffffffffu, t.k0 == m);
        const int src = __ffs(own) - 1;
        okey[q] = m;
        oidx[q] = __shfl_sync(0xffffffffu, t.j0, src);
        if (lane == src) t.pop();
    }
}

// One warp per keypoint of frame 1.
__global__ void __launch_bounds__(256)
k_init_shortlist(const OrbxInitArgs a) {
    const int lane = threadIdx.x & 31;
    const int i1 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i1 >= a.n1) return;
    if (lane == 0) { a.matches12[i1] = -1; a.pushed[i1] = -1; if (i1 == 0) a.n_matches[1] = 0; }
    unsigned key[ORBX_SHORT_K];
    int idx[ORBX_SHORT_K];
    int count = 0;
#pragma unroll
    for (int q = 0; q < ORBX_SHORT_K; ++q) { key[q] = 0xffffffffu; idx[q] = -1; }
    if (__ldg(&a.k1[i1].octave) <= 0) {                                  // `if (level1 > 0) continue;` :723-725
        uint32_t dq[8];
        const uint4* q = reinterpret_cast<const uint4*>(a.d1 + (size_t)i1 * 8);
        const uint4 q0 = __ldg(q), q1 = __ldg(q + 1);
        dq[0] = q0.x; dq[1] = q0.y; dq[2] = q0.z; dq[3] = q0.w; dq[4] = q1.x; dq[5] = q1.y; dq[6] = q1.z; dq[7] = q1.w;
        if (__ldg(&a.k1[i1].octave) == 0)
            init_enumerate<false>(a, a.prev[2 * i1], a.prev[2 * i1 + 1], dq, nullptr, key, idx, count);
    }
    if (lane == 0) {
        a.sl_key[2 * i1] = make_uint4(key[0], key[1], key[2], key[3]);
        a.sl_key[2 * i1 + 1] = make_uint4(key[4], key[5], key[6], key[7]);
        a.sl_idx[2 * i1] = make_uint4((unsigned)idx[0], (unsigned)idx[1], (unsigned)idx[2], (unsigned)idx[3]);
        a.sl_idx[2 * i1 + 1] = make_uint4((unsigned)idx[4], (unsigned)idx[5], (unsigned)idx[6], (unsigned)idx[7]);
        a.sl_count[i1] = count;
    }
}

// One CTA of 256 threads = 32 keypoint slots x 8 shortlist entries.  The reference's loop carries vMatchedDistance and
// vnMatches21 from one keypoint to the next; here the 32 keypoints of a chunk are evaluated together against the
// current state, and the longest prefix for which that provably equals the one-at-a-time result is applied: a keypoint
// "clashes" when an earlier pending keypoint of the chunk matches a candidate that is still live in its own shortlist
// (that match would lower vMatchedDistance under it), or when its shortlist ran dry (it then has to be re-enumerated
// against the exact state).  Everything before the first clash is applied in parallel, then the rest is evaluated again;
// the first pending keypoint can never clash, so every round makes progress.  A keypoint whose best unfiltered candidate
// is already above TH_LOW can never match (filtering only removes candidates) and is not visited at all.  The rotation
// histogram, the three-maxima filter, the vbPrevMatched update and the final count run on all threads afterwards.
__global__ void __launch_bounds__(256)
k_init_resolve(const OrbxInitArgs a) {
    extern __shared__ __align__(16) uint8_t smem_init[];
    int* m21 = reinterpret_cast<int*>(smem_init);                       // vnMatches21
    uint16_t* vmd = reinterpret_cast<uint16_t*>(m21 + a.n2);            // vMatchedDistance (0xffff = INT_MAX)
    __shared__ int s_hist[ORBX_MATCH_HISTO];
    __shared__ int s_ind[3];
    __shared__ int s_nm;
    __shared__ int s_tag[ORBX_INIT_TAGS];    // bucket (candidate index mod ORBX_INIT_TAGS) -> lowest slot whose speculative match it holds (32: none)
    __shared__ unsigned s_pending;    // slots not resolved yet
    __shared__ int s_stop;            // first slot that clashes in this round (32: none)
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < a.n2; i += blockDim.x) { m21[i] = -1; vmd[i] = 0xffffu; }
    if (tid < ORBX_MATCH_HISTO) s_hist[tid] = 0;
    if (tid == 0) { s_nm = 0; s_pending = 0u; }
    for (int i = tid; i < ORBX_INIT_TAGS; i += blockDim.x) s_tag[i] = 32;
    int n_fallback = 0;
    const int slot = tid >> 3, q = tid & 7, gl = lane & 24;
    const unsigned* sl_key = reinterpret_cast<const unsigned*>(a.sl_key);
    const int* sl_idx = reinterpret_cast<const int*>(a.sl_idx);
    // ---- ordered list of the keypoints that can match at all: shortlist not empty and its best distance <= TH_LOW ----
    __shared__ int s_scan[8];
    __shared__ int s_nact;
    {
        // 256 consecutive keypoints per step (coalesced loads, four steps' loads in flight), ballot + warp totals give the rank
        int total = 0;
        for (int base = 0; base < a.n1; base += 4 * 256) {
            bool f[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = base + j * 256 + tid;
                f[j] = i < a.n1 && a.sl_count[i] > 0 && (int)(sl_key[(size_t)i * ORBX_SHORT_K] >> 22) <= ORBX_MATCH_TH_LOW;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned bal = __ballot_sync(0xffffffffu, f[j]);
                if (lane == 0) s_scan[tid >> 5] = __popc(bal);
                __syncthreads();
                int off = total, all = 0;
#pragma unroll
                for (int w = 0; w < 8; ++w) { const int c = s_scan[w]; all += c; if (w < (tid >> 5)) off += c; }
                if (f[j]) a.act_list[off + __popc(bal & ((1u << lane) - 1u))] = base + j * 256 + tid;
                total += all;
                __syncthreads();
            }
        }
        if (tid == 0) s_nact = total;
        __syncthreads();
    }
    const int nact = s_nact;
    unsigned nkey = 0xffffffffu;
    int nidx = 0, ncnt = 0, ni1 = -1;
    if (slot < nact) {
        ni1 = a.act_list[slot];
        nkey = sl_key[(size_t)ni1 * ORBX_SHORT_K + q]; nidx = sl_idx[(size_t)ni1 * ORBX_SHORT_K + q]; ncnt = a.sl_count[ni1];
    }
    for (int base = 0; base < nact; base += 32) {
        const int i1 = ni1;
        const unsigned key = nkey;
        const int idx = nidx, cnt = ncnt;
        {   // the next chunk's shortlist entries are in flight while this chunk is resolved
            const int nx = base + 32 + slot;
            nkey = 0xffffffffu; nidx = 0; ncnt = 0; ni1 = -1;
            if (nx < nact) {
                ni1 = a.act_list[nx];
                nkey = sl_key[(size_t)ni1 * ORBX_SHORT_K + q]; nidx = sl_idx[(size_t)ni1 * ORBX_SHORT_K + q]; ncnt = a.sl_count[ni1];
            }
        }
        const bool valid = key != 0xffffffffu;
        const int dist = (int)(key >> 22);
        const int listed = __popc((__ballot_sync(0xffffffffu, valid) >> gl) & 0xffu);
        __syncthreads();                                                  // previous chunk fully applied; s_pending == 0
        if (q == 0 && i1 >= 0) atomicOr(&s_pending, 1u << slot);
        __syncthreads();
        while (true) {
            const unsigned pending = s_pending;
            if (pending == 0u) break;
            const bool mine = (pending >> slot) & 1u;
            const bool first = mine && (pending & ((1u << slot) - 1u)) == 0u;
            // ---- evaluate every pending slot against the current state ----
            const bool live = mine && valid && (int)vmd[valid ? idx : 0] > dist;                        // :744
            const unsigned lm = (__ballot_sync(0xffffffffu, live) >> gl) & 0xffu;
            const int b = __ffs(lm) - 1, s2 = __ffs(lm & (lm - 1u)) - 1;
            const bool dry = mine && cnt > listed && s2 < 0;
            int bestDist = 0x7fffffff, bestDist2 = 0x7fffffff, bestIdx2 = -1;
            {
                const unsigned bk = __shfl_sync(0xffffffffu, key, gl | (b & 7));
                const int bi = __shfl_sync(0xffffffffu, idx, gl | (b & 7));
                const unsigned sk = __shfl_sync(0xffffffffu, key, gl | (s2 & 7));
                if (b >= 0) { bestDist = (int)(bk >> 22); bestIdx2 = bi; }
                if (s2 >= 0) bestDist2 = (int)(sk >> 22);
            }
            // the first pending slot sees the exact state: if its shortlist ran dry, its warp enumerates again with the filter
            if (__any_sync(0xffffffffu, first && dry)) {
                const int fl = __ffs(__ballot_sync(0xffffffffu, first && dry)) - 1;           // a lane of that slot
                const int fi1 = __shfl_sync(0xffffffffu, i1, fl);
                uint32_t dq[8];
                const uint4* qd = reinterpret_cast<const uint4*>(a.d1 + (size_t)fi1 * 8);
                const uint4 q0 = __ldg(qd), q1 = __ldg(qd + 1);
                dq[0] = q0.x; dq[1] = q0.y; dq[2] = q0.z; dq[3] = q0.w; dq[4] = q1.x; dq[5] = q1.y; dq[6] = q1.z; dq[7] = q1.w;
                unsigned fkey[ORBX_SHORT_K];
                int fidx[ORBX_SHORT_K];
                int cnt2;
                init_enumerate<true>(a, a.prev[2 * fi1], a.prev[2 * fi1 + 1], dq, vmd, fkey, fidx, cnt2);
                if (first) {
                    bestDist = 0x7fffffff; bestDist2 = 0x7fffffff; bestIdx2 = -1;
                    if (fkey[0] != 0xffffffffu) { bestDist = (int)(fkey[0] >> 22); bestIdx2 = fidx[0]; }
                    if (fkey[1] != 0xffffffffu) bestDist2 = (int)(fkey[1] >> 22);
                }
                if (lane == fl) ++n_fallback;
            }
            const bool match = mine && bestIdx2 >= 0 && bestDist <= ORBX_MATCH_TH_LOW &&
                               (float)bestDist < __fmul_rn((float)bestDist2, a.nn_ratio);                 // :758-760
            // every speculative match tags its candidate's bucket with its slot (lowest slot wins); two candidates sharing a
            // bucket can only cause a spurious clash, i.e. one more round, never a missed one
            if (q == 0 && match) atomicMin(&s_tag[bestIdx2 & (ORBX_INIT_TAGS - 1)], slot);
            if (tid == 0) s_stop = 32;
            __syncthreads();
            // ---- clash detection: did an earlier slot's speculative match land on this live candidate? ----
            const bool clash = !first && (dry || (live && s_tag[idx & (ORBX_INIT_TAGS - 1)] < slot));
            if (clash) atomicMin(&s_stop, slot);
            __syncthreads();
            if (q == 0 && match) s_tag[bestIdx2 & (ORBX_INIT_TAGS - 1)] = 32;             // all tags back to "none" for the next round
            // ---- apply the clash-free prefix ----
            const int stop = s_stop;
            if (mine && slot < stop && q == 0) {
                if (match) {
                    const int old = m21[bestIdx2];
                    if (old >= 0) a.matches12[old] = -1;                                    // :762-766
                    a.matches12[i1] = bestIdx2;
                    a.pushed[i1] = bestIdx2;
                    m21[bestIdx2] = i1;
                    vmd[bestIdx2] = (uint16_t)bestDist;
                }
                atomicAnd(&s_pending, ~(1u << slot));
            }
            __syncthreads();
        }
    }
    if (n_fallback) atomicAdd(&a.n_matches[1], n_fallback);
    __threadfence_block();
    __syncthreads();
    // ---- rotation histogram over every keypoint that was ever matched (:772-783; entries of matches that were
    // later taken over stay in the histogram, exactly like rotHist) ----
    if (a.check_orientation) {
        const float factor = __fdiv_rn(1.0f, (float)ORBX_MATCH_HISTO);
        for (int i1 = tid; i1 < a.n1; i1 += blockDim.x) {
            const int b2 = a.pushed[i1];
            if (b2 < 0) continue;
            float rot = __fsub_rn(a.k1[i1].angle, a.k2[b2].angle);
            if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
            int bin = (int)roundf(__fmul_rn(rot, factor));
            if (bin == ORBX_MATCH_HISTO) bin = 0;
            a.pushed[i1] = bin;
            atomicAdd(&s_hist[bin], 1);
        }
        __syncthreads();
        if (tid == 0) {   // ComputeThreeMaxima, :2303-2344
            int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
            for (int i = 0; i < ORBX_MATCH_HISTO; i++) {
                const int s = s_hist[i];
                if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
                else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
                else if (s > max3) { max3 = s; ind3 = i; }
            }
            if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
            else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) { ind3 = -1; }
            s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
        }
        __syncthreads();
        for (int i1 = tid; i1 < a.n1; i1 += blockDim.x) {
            const int bin = a.pushed[i1];
            if (bin < 0) continue;
            if (bin == s_ind[0] || bin == s_ind[1] || bin == s_ind[2]) continue;
            a.matches12[i1] = -1;                                                           // :797-801
        }
        __syncthreads();
    }
    // ---- vbPrevMatched update (:808-811) and the match count ----
    int local = 0;
    for (int i1 = tid; i1 < a.n1; i1 += blockDim.x) {
        const int m = a.matches12[i1];
        if (m >= 0) {
            a.prev[2 * i1] = a.k2[m].x; a.prev[2 * i1 + 1] = a.k2[m].y;
            ++local;
        }
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if (lane == 0 && local) atomicAdd(&s_nm, local);
    __syncthreads();
    if (tid == 0) a.n_matches[0] = s_nm;   // n_matches[1]: filtered re-enumerations (diagnostic), cleared by k_init_shortlist
}
