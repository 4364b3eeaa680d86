// extractorb_b200/csrc/orbx_kernels.cuh -- sm_100a kernels of the ORB extraction path.
//
//   k_pyr_level0 / k_pyr_resize : ComputePyramid          (reference ORBextractor.cc:1164-1219)
//   k_fast_cells                : cell loop + cv::FAST    (:797-864)
//   k_octree                    : DistributeOctTree       (:544-771, DivideNode :486-542)
//   k_blur7                     : GaussianBlur 7x7 s=2    (:1126-1127)
//   k_describe                  : IC_Angle, rBRIEF, scale-back + two-ended scatter (:75-145, :1131-1159)
//
// All pixel arithmetic is integer / fixed point and bit-exact with the OpenCV primitives the reference
// calls; the float steps (fastAtan2, pattern rotation, scale-back) use __f*_rn intrinsics so that nvcc
// never contracts them into FMAs (the reference is built without FMA, CMakeLists.txt:4-11).
#ifndef ORBX_KERNELS_CUH_
#define ORBX_KERNELS_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "orbx_plan.h"

#define ORBX_FULL_MASK 0xffffffffu

__device__ __forceinline__ int reflect101(int p, int len) {
    p = p < 0 ? -p : p;
    return p >= len ? 2 * (len - 1) - p : p;
}

// =================================================================================================
// K1  ComputePyramid.  One thread produces one aligned 32-bit word (4 pixels) of the bordered plane,
// border included: a border pixel is the level pixel at the reflect-101 coordinate, so it is simply
// recomputed there (no second pass, no dependency on neighbouring CTAs).
// =================================================================================================
__global__ void __launch_bounds__(256)
k_pyr_level0(const __grid_constant__ OrbxPlan plan, const OrbxWs ws, const uint8_t* __restrict__ imgs,
             long long row_stride, long long frame_stride) {
    const OrbxLevel& L = plan.lv[0];
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    const int frame = blockIdx.z;
    if (wx * 4 >= L.pitch || r >= L.plane_rows) return;
    const uint8_t* src = imgs + (long long)frame * frame_stride;
    const int y = reflect101(r - ORBX_EDGE, L.h);
    const uint8_t* srow = src + (long long)y * row_stride;
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int dx = wx * 4 + b - ORBX_PADL;
        uint32_t v = 0;
        if (dx >= -ORBX_EDGE && dx < L.w + ORBX_EDGE) v = __ldg(srow + reflect101(dx, L.w));
        word |= v << (8 * b);
    }
    uint8_t* dst = ws.pyr + (long long)frame * ws.pyr_stride + L.plane_off;
    *reinterpret_cast<uint32_t*>(dst + (long long)r * L.pitch + wx * 4) = word;
}

// cv::resize INTER_LINEAR, 8UC1 fixed point (11-bit coefficients): see oracle/cv_prims.c for the model.
__global__ void __launch_bounds__(256)
k_pyr_resize(const __grid_constant__ OrbxPlan plan, const OrbxWs ws, int level) {
    const OrbxLevel& L = plan.lv[level];
    const OrbxLevel& S = plan.lv[level - 1];
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    const int frame = blockIdx.z;
    if (wx * 4 >= L.pitch || r >= L.plane_rows) return;
    uint8_t* fbase = ws.pyr + (long long)frame * ws.pyr_stride;
    const uint8_t* sroi = fbase + S.plane_off + (long long)ORBX_EDGE * S.pitch + ORBX_PADL;
    const int y = reflect101(r - ORBX_EDGE, L.h);
    const int2 yt = __ldg(ws.ytab + L.ytab_off + y);
    const int b0 = yt.y & 0xffff, b1 = yt.y >> 16;
    const uint8_t* s0 = sroi + (long long)yt.x * S.pitch;
    const uint8_t* s1 = s0 + S.pitch;  // row h_src is the border row; its weight b1 is 0 there
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int dx = wx * 4 + b - ORBX_PADL;
        uint32_t v = 0;
        if (dx >= -ORBX_EDGE && dx < L.w + ORBX_EDGE) {
            const int x = reflect101(dx, L.w);
            const int2 xt = __ldg(ws.xtab + L.xtab_off + x);
            if (xt.y == -1) {  // exact 2x2 decimation: OpenCV takes the INTER_AREA fast path
                v = (s0[xt.x] + s0[xt.x + 1] + s1[xt.x] + s1[xt.x + 1] + 2) >> 2;
            } else {
                const int a0 = xt.y & 0xffff, a1 = xt.y >> 16;
                const int h0 = s0[xt.x] * a0 + s0[xt.x + 1] * a1;
                const int h1 = s1[xt.x] * a0 + s1[xt.x + 1] * a1;
                v = (uint32_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
            }
        }
        word |= (v & 0xffu) << (8 * b);
    }
    uint8_t* dst = fbase + L.plane_off;
    *reinterpret_cast<uint32_t*>(dst + (long long)r * L.pitch + wx * 4) = word;
}

// =================================================================================================
// K2  FAST-9/16 + 3x3 NMS + ini/min threshold choice, one WARP per 30-px grid cell.
//
// Threshold-free formulation: best(p) = max over the 16 arcs of 9 contiguous ring pixels of
// min(centre-ring) and min(ring-centre); p is a corner at threshold t  <=>  best(p) > t, its score is
// best(p)-1, and p survives OpenCV's NMS at any t  <=>  best(p) > best(q) for its 8 neighbours q inside
// the cell interior.  So one pass at t = min(ini,min) yields both candidate sets, and the reference's
// "re-run with minThFAST iff empty" (:835-838) is a per-cell ballot.
// =================================================================================================
#define ORBX_FAST_WARPS 4

__device__ __forceinline__ int fast_best(const uint8_t* __restrict__ t, int p, int tp) {
    const int c = t[p];
    int d[16];
    d[0] = c - t[p + 3 * tp];      d[1] = c - t[p + 3 * tp + 1];  d[2] = c - t[p + 2 * tp + 2];
    d[3] = c - t[p + tp + 3];      d[4] = c - t[p + 3];           d[5] = c - t[p - tp + 3];
    d[6] = c - t[p - 2 * tp + 2];  d[7] = c - t[p - 3 * tp + 1];  d[8] = c - t[p - 3 * tp];
    d[9] = c - t[p - 3 * tp - 1];  d[10] = c - t[p - 2 * tp - 2]; d[11] = c - t[p - tp - 3];
    d[12] = c - t[p - 3];          d[13] = c - t[p + tp - 3];     d[14] = c - t[p + 2 * tp - 2];
    d[15] = c - t[p + 3 * tp - 1];
    int mn3[16], mx3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        mn3[k] = __vimin3_s32(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
        mx3[k] = __vimax3_s32(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
    }
    int bright = -256, dark = 256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        bright = max(bright, __vimin3_s32(mn3[k], mn3[(k + 3) & 15], mn3[(k + 6) & 15]));
        dark = min(dark, __vimax3_s32(mx3[k], mx3[(k + 3) & 15], mx3[(k + 6) & 15]));
    }
    return max(bright, -dark);
}

// Upper bound of best(p) from the 8 antipodal ring pairs: every arc of 9 contiguous ring pixels contains
// at least one pixel of each pair, so  best <= max( c - max_pairs(min(a,b)),  min_pairs(max(a,b)) - c ).
// The first two pairs (vertical, horizontal) reject most pixels of smooth regions after 5 loads.
__device__ __forceinline__ bool fast_quick(const uint8_t* __restrict__ t, int p, int tp, int th) {
    const int c = t[p];
    const int a0 = t[p + 3 * tp], a8 = t[p - 3 * tp], a4 = t[p + 3], a12 = t[p - 3];
    int dk = max(min(a0, a8), min(a4, a12));   // dark side:   need c - dk > th
    int br = min(max(a0, a8), max(a4, a12));   // bright side: need br - c > th
    if (c - dk <= th && br - c <= th) return false;
    const int a2 = t[p + 2 * tp + 2], a10 = t[p - 2 * tp - 2], a6 = t[p - 2 * tp + 2], a14 = t[p + 2 * tp - 2];
    dk = __vimax3_s32(dk, min(a2, a10), min(a6, a14));
    br = __vimin3_s32(br, max(a2, a10), max(a6, a14));
    if (c - dk <= th && br - c <= th) return false;
    const int a1 = t[p + 3 * tp + 1], a9 = t[p - 3 * tp - 1], a3 = t[p + tp + 3], a11 = t[p - tp - 3];
    const int a5 = t[p - tp + 3], a13 = t[p + tp - 3], a7 = t[p - 3 * tp + 1], a15 = t[p + 3 * tp - 1];
    dk = __vimax3_s32(dk, min(a1, a9), min(a3, a11));
    br = __vimin3_s32(br, max(a1, a9), max(a3, a11));
    dk = __vimax3_s32(dk, min(a5, a13), min(a7, a15));
    br = __vimin3_s32(br, max(a5, a13), max(a7, a15));
    return c - dk > th || br - c > th;
}

__global__ void __launch_bounds__(ORBX_FAST_WARPS * 32)
k_fast_cells(const __grid_constant__ OrbxPlan plan, const OrbxWs ws) {
    extern __shared__ __align__(16) uint8_t smem_fast[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ci = blockIdx.x * ORBX_FAST_WARPS + wib;
    const int frame = blockIdx.y;
    if (ci >= plan.ncells_total) return;
    const OrbxCell cell = ws.cells[ci];
    const OrbxLevel& L = plan.lv[cell.level];

    const int tp = plan.fast_tp;
    const int tile_bytes = tp * plan.fast_trows;
    uint8_t* tile = smem_fast + (size_t)wib * (2 * tile_bytes + 2 * plan.fast_qcap);
    uint8_t* score = tile + tile_bytes;
    uint16_t* queue = reinterpret_cast<uint16_t*>(score + tile_bytes);

    const int cw = cell.cw, ch = cell.ch;
    // ---- stage the cell image (aligned 32-bit loads; pixel (r,c) lands at tile[r*tp + a + c]) ----
    const uint8_t* plane = ws.pyr + (long long)frame * ws.pyr_stride + L.plane_off;
    const int gx = ORBX_PADL + cell.x0;
    const int a = gx & 3;
    const int nwords = (a + cw + 3) >> 2;
    const uint8_t* g0 = plane + (long long)(ORBX_EDGE + cell.y0) * L.pitch + (gx - a);
    {
        const int rows_per_it = 32 / nwords;
        if (rows_per_it >= 1) {
            const int lr = lane / nwords, lw = lane - lr * nwords;
            if (lr < rows_per_it)
                for (int r = lr; r < ch; r += rows_per_it) {
                    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(g0 + (long long)r * L.pitch) + lw);
                    *reinterpret_cast<uint32_t*>(tile + r * tp + 4 * lw) = v;
                }
        } else {
            for (int r = 0; r < ch; ++r)
                for (int wd = lane; wd < nwords; wd += 32) {
                    const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(g0 + (long long)r * L.pitch) + wd);
                    *reinterpret_cast<uint32_t*>(tile + r * tp + 4 * wd) = v;
                }
        }
    }
    for (int i = lane; i < (tile_bytes >> 2); i += 32) reinterpret_cast<uint32_t*>(score)[i] = 0;
    __syncwarp();

    const int iw = cw - 6, ih = ch - 6;
    const int npix = iw * ih;
    const unsigned magic = (1u << 24) / (unsigned)iw + 1u;  // idx / iw == (idx * magic) >> 24 (idx < 2^14, iw < 2^7)
    // The reference detects at iniThFAST and re-runs the cell at minThFAST only if that found nothing
    // (:818-838); the same two phases here, the second one skipped when it cannot add anything.
    int qn = 0, n_out = 0, use_th = plan.ini_th;
    for (int phase = 0; phase < 2; ++phase) {
        if (phase == 1) {
            if (plan.min_th >= plan.ini_th) break;
            use_th = plan.min_th;
        }
        // ---- quick rejection over the interior, survivors compacted into `queue` ----
        qn = 0;
        for (int base = 0; base < npix; base += 32) {
            const int idx = base + lane;
            bool pass = false;
            int p = 0;
            if (idx < npix) {
                const int r = (int)(((unsigned)idx * magic) >> 24);
                p = (r + 3) * tp + a + (idx - r * iw) + 3;
                pass = fast_quick(tile, p, tp, use_th);
            }
            const unsigned bal = __ballot_sync(ORBX_FULL_MASK, pass);
            if (pass) queue[qn + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)p;
            qn += __popc(bal);
        }
        __syncwarp();
        // ---- exact corner measure for the queued pixels ----
        for (int e = lane; e < qn; e += 32) {
            const int p = queue[e];
            const int best = fast_best(tile, p, tp);
            if (best > use_th) score[p] = (uint8_t)best;
            else queue[e] = 0xffff;
        }
        __syncwarp();
        // ---- strict 3x3 non-max suppression (pixels at or below the threshold count as 0, like cv::FAST) ----
        n_out = 0;
        for (int base = 0; base < qn; base += 32) {
            const int e = base + lane;
            bool lm = false;
            if (e < qn) {
                const int p = queue[e];
                if (p != 0xffff) {
                    const int s = score[p];
                    lm = s > score[p - 1] && s > score[p + 1] && s > score[p - tp - 1] && s > score[p - tp] &&
                         s > score[p - tp + 1] && s > score[p + tp - 1] && s > score[p + tp] && s > score[p + tp + 1];
                    if (!lm) queue[e] = 0xffff;
                }
            }
            n_out += __popc(__ballot_sync(ORBX_FULL_MASK, lm));
        }
        if (n_out > 0) break;
        __syncwarp();
    }
    if (n_out == 0) return;

    // ---- emission: one atomic per cell reserves the slots; entries carry the emission-order key ----
    int* counter = ws.cand_count + frame * plan.nlevels + cell.level;
    int slot0 = 0;
    if (lane == 0) slot0 = atomicAdd(counter, n_out);
    slot0 = __shfl_sync(ORBX_FULL_MASK, slot0, 0);
    if (slot0 + n_out > L.cand_cap) {
        if (lane == 0) atomicOr(ws.flags + frame, 1);
    }
    uint2* cand = ws.cand + (long long)frame * ws.cand_stride + L.cand_off;
    const unsigned tpmagic = (1u << 24) / (unsigned)tp + 1u;  // p / tp (p < 2^14, tp <= 136)
    int written = 0;
    for (int base = 0; base < qn; base += 32) {
        const int e = base + lane;
        const int p = e < qn ? queue[e] : 0xffff;
        const bool keep = p != 0xffff;
        const unsigned bal = __ballot_sync(ORBX_FULL_MASK, keep);
        if (keep) {
            const int slot = slot0 + written + __popc(bal & ((1u << lane) - 1u));
            if (slot < L.cand_cap) {
                const int r = (int)(((unsigned)p * tpmagic) >> 24);
                const int c = p - r * tp - a;  // cell-image coordinates (>= 3)
                const uint32_t x = (uint32_t)(c + cell.xoff), y = (uint32_t)(r + cell.yoff);
                const uint32_t resp = (uint32_t)score[p] - 1u;
                cand[slot] = make_uint2(x | (y << ORBX_COORD_BITS) | (resp << 24),
                                        cell.ordbase | ((uint32_t)r << 7) | (uint32_t)c);
            }
        }
        written += __popc(bal);
    }
}

// =================================================================================================
// K3  DistributeOctTree as a level-synchronous array rebuild, one CTA per (frame, level).
//
// The reference keeps a std::list of nodes, pushes children to the FRONT and erases parents in place,
// so at any time the list is "live nodes by creation time, newest first" (initial nodes last, in
// order).  One pass therefore maps list -> reverse(children in creation order) ++ untouched nodes.
// The careful phase (:678-743) splits nodes in (size desc, pointer desc) order until the list holds
// >= N nodes; under a monotonic allocator pointer order == creation order == list position (asc).
// Keys never move: each key carries the list position of its node and the split geometry is
// recomputed from the node rectangle.
// =================================================================================================
#define ORBX_QT_THREADS 256

struct QtShared {
    short4* rect[2];  // x0, x1, y0, y1
    int* cnt[2];
    int* cc;          // 4 child counts per node (reused as 64-bit argmax slots at the end)
    int* seq;         // processing sequence (list positions)
    int* tmp;         // children per sequence entry -> exclusive scan
    int* cbase;       // creation index of the first child, -1 if the node is not split this pass
    int* surv;        // survivor flag -> new list position
    unsigned long long* skey;  // careful-phase sort keys
};

__device__ __forceinline__ int qt_quadrant(short4 r, int x, int y) {
    const int mx = r.x + ((r.y - r.x + 1) >> 1);  // UL.x + ceil((UR.x-UL.x)/2), :488
    const int my = r.z + ((r.w - r.z + 1) >> 1);
    return (x < mx ? 0 : 1) + (y < my ? 0 : 2);   // n1, n2, n3, n4 of :517-531
}

// Exclusive scan of v[0..n) in shared memory (in place); returns the total to every thread.
__device__ int block_excl_scan(int* v, int n, int* scratch) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int ipt = (n + nt - 1) / nt;
    const int beg = min(tid * ipt, n), end = min(beg + ipt, n);
    int sum = 0;
    for (int i = beg; i < end; ++i) sum += v[i];
    const int lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(ORBX_FULL_MASK, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();  // scratch may still be read from a previous call
    if (lane == 31) scratch[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const int w = lane < nw ? scratch[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(ORBX_FULL_MASK, winc, o);
            if (lane >= o) winc += t;
        }
        if (lane < nw) scratch[lane] = winc - w;
        if (lane == 31) scratch[32] = winc;
    }
    __syncthreads();
    int run = scratch[wid] + inc - sum;
    const int total = scratch[32];
    for (int i = beg; i < end; ++i) {
        const int x = v[i];
        v[i] = run;
        run += x;
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(ORBX_QT_THREADS)
k_octree(const __grid_constant__ OrbxPlan plan, const OrbxWs ws) {
    extern __shared__ __align__(16) uint8_t smem_qt[];
    __shared__ int s_scratch[33];
    __shared__ int s_nexp, s_cut, s_state;

    const int level = blockIdx.x, frame = blockIdx.y;
    const OrbxLevel& L = plan.lv[level];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const int NC = plan.qt_nc;
    const int N = L.N;

    QtShared q;
    {
        uint8_t* p = smem_qt;
        q.skey = reinterpret_cast<unsigned long long*>(p); p += (size_t)NC * 8;
        q.rect[0] = reinterpret_cast<short4*>(p); p += (size_t)NC * 8;
        q.rect[1] = reinterpret_cast<short4*>(p); p += (size_t)NC * 8;
        q.cc = reinterpret_cast<int*>(p); p += (size_t)NC * 16;
        q.cnt[0] = reinterpret_cast<int*>(p); p += (size_t)NC * 4;
        q.cnt[1] = reinterpret_cast<int*>(p); p += (size_t)NC * 4;
        q.seq = reinterpret_cast<int*>(p); p += (size_t)NC * 4;
        q.tmp = reinterpret_cast<int*>(p); p += (size_t)NC * 4;
        q.cbase = reinterpret_cast<int*>(p); p += (size_t)NC * 4;
        q.surv = reinterpret_cast<int*>(p);
    }

    const int ncand = min(ws.cand_count[frame * plan.nlevels + level], L.cand_cap);
    const uint2* cand = ws.cand + (long long)frame * ws.cand_stride + L.cand_off;
    uint16_t* keynode = ws.keynode + (long long)frame * ws.cand_stride + L.cand_off;
    const int CM = (1 << ORBX_COORD_BITS) - 1;

    // ---- initial nodes (:548-590): assign by float division, drop the empty ones ----
    int cur = 0;
    for (int i = tid; i < L.nIni; i += nt) {
        q.rect[0][i] = make_short4((short)(int)__fmul_rn(L.hX, (float)i), (short)(int)__fmul_rn(L.hX, (float)(i + 1)),
                                   0, (short)L.span_y);
        q.cnt[0][i] = 0;
    }
    __syncthreads();
    for (int k = tid; k < ncand; k += nt) {
        const int x = cand[k].x & CM;
        const int r = (int)__fdiv_rn((float)x, L.hX);
        keynode[k] = (uint16_t)r;
        atomicAdd(&q.cnt[0][r], 1);
    }
    __syncthreads();
    for (int i = tid; i < L.nIni; i += nt) q.surv[i] = q.cnt[0][i] > 0;
    __syncthreads();
    int nlive = block_excl_scan(q.surv, L.nIni, s_scratch);
    for (int i = tid; i < L.nIni; i += nt)
        if (q.cnt[0][i] > 0) {
            q.rect[1][q.surv[i]] = q.rect[0][i];
            q.cnt[1][q.surv[i]] = q.cnt[0][i];
        }
    for (int k = tid; k < ncand; k += nt) keynode[k] = (uint16_t)q.surv[keynode[k]];
    cur = 1;
    __syncthreads();

    bool careful = false;
    for (;;) {
        const int prev_size = nlive;
        const short4* rect = q.rect[cur];
        const int* cnt = q.cnt[cur];
        // ---- processing sequence: live nodes holding > 1 key, in list order ----
        for (int i = tid; i < nlive; i += nt) {
            q.tmp[i] = cnt[i] > 1;
            q.cc[4 * i + 0] = 0; q.cc[4 * i + 1] = 0; q.cc[4 * i + 2] = 0; q.cc[4 * i + 3] = 0;
        }
        if (tid == 0) { s_nexp = 0; s_cut = -1; }
        __syncthreads();
        const int ns = block_excl_scan(q.tmp, nlive, s_scratch);
        if (ns == 0) break;  // every node is a single key: list size unchanged -> finish (:674)
        for (int i = tid; i < nlive; i += nt)
            if (cnt[i] > 1) q.seq[q.tmp[i]] = i;
        __syncthreads();
        if (careful) {
            // (size desc, later-created first) == (size desc, list position asc); rank sort
            for (int s = tid; s < ns; s += nt) {
                const int pos = q.seq[s];
                q.skey[s] = ((unsigned long long)(unsigned)cnt[pos] << 32) | (unsigned)(0x7fffffff - pos);
            }
            __syncthreads();
            for (int s = tid; s < ns; s += nt) {
                const unsigned long long key = q.skey[s];
                int rank = 0;
                for (int t = 0; t < ns; ++t) rank += q.skey[t] > key;
                q.tmp[rank] = 0x7fffffff - (int)(unsigned)(key & 0xffffffffu);
            }
            __syncthreads();
            for (int s = tid; s < ns; s += nt) q.seq[s] = q.tmp[s];
            __syncthreads();
        }
        // ---- child occupancy of every node in the sequence ----
        for (int base = 0; base < ncand; base += nt) {
            const int k = base + tid;
            int key = -1;
            if (k < ncand) {
                const int pos = keynode[k];
                if (cnt[pos] > 1) {
                    const uint32_t v = cand[k].x;
                    key = 4 * pos + qt_quadrant(rect[pos], (int)(v & CM), (int)((v >> ORBX_COORD_BITS) & CM));
                }
            }
            const unsigned peers = __match_any_sync(ORBX_FULL_MASK, key);
            if (key >= 0 && lane == __ffs(peers) - 1) atomicAdd(&q.cc[key], __popc(peers));
        }
        __syncthreads();
        for (int s = tid; s < ns; s += nt) {
            const int pos = q.seq[s];
            q.tmp[s] = (q.cc[4 * pos] > 0) + (q.cc[4 * pos + 1] > 0) + (q.cc[4 * pos + 2] > 0) + (q.cc[4 * pos + 3] > 0);
        }
        if (tid == 0) q.tmp[ns] = 0;
        __syncthreads();
        const int tall = block_excl_scan(q.tmp, ns + 1, s_scratch);  // tmp[s] = children created before split s
        // ---- how many splits are applied (careful phase stops once the list reaches N, :735) ----
        int nsplit = ns;
        if (careful) {
            for (int j = tid + 1; j <= ns; j += nt) {
                const bool now = prev_size + q.tmp[j] - j >= N;
                const bool before = prev_size + q.tmp[j - 1] - (j - 1) >= N;
                if (now && !before) s_cut = j;
            }
            __syncthreads();
            if (s_cut > 0) nsplit = s_cut;
        }
        const int T = q.tmp[nsplit];  // children created this pass
        (void)tall;
        for (int i = tid; i < nlive; i += nt) { q.cbase[i] = -1; q.surv[i] = 1; }
        __syncthreads();
        for (int s = tid; s < nsplit; s += nt) {
            const int pos = q.seq[s];
            q.cbase[pos] = q.tmp[s];
            q.surv[pos] = 0;
        }
        __syncthreads();
        const int nsurv = block_excl_scan(q.surv, nlive, s_scratch);
        short4* nrect = q.rect[cur ^ 1];
        int* ncnt = q.cnt[cur ^ 1];
        // children: list position T-1-(creation index); DivideNode geometry :488-514
        for (int s = tid; s < nsplit; s += nt) {
            const int pos = q.seq[s];
            const short4 r = rect[pos];
            const int mx = r.x + ((r.y - r.x + 1) >> 1), my = r.z + ((r.w - r.z + 1) >> 1);
            int ci = q.tmp[s];
            int nexp = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int n = q.cc[4 * pos + c];
                if (n > 0) {
                    const int np = T - 1 - ci;
                    nrect[np] = make_short4((short)((c & 1) ? mx : r.x), (short)((c & 1) ? r.y : mx),
                                            (short)((c & 2) ? my : r.z), (short)((c & 2) ? r.w : my));
                    ncnt[np] = n;
                    nexp += n > 1;
                    ++ci;
                }
            }
            if (nexp) atomicAdd(&s_nexp, nexp);
        }
        for (int i = tid; i < nlive; i += nt)
            if (q.cbase[i] < 0) {
                nrect[T + q.surv[i]] = rect[i];
                ncnt[T + q.surv[i]] = cnt[i];
            }
        // keys follow their node
        for (int k = tid; k < ncand; k += nt) {
            const int pos = keynode[k];
            const int cb = q.cbase[pos];
            int np;
            if (cb >= 0) {
                const uint32_t v = cand[k].x;
                const int c = qt_quadrant(rect[pos], (int)(v & CM), (int)((v >> ORBX_COORD_BITS) & CM));
                int before = 0;
                for (int j = 0; j < c; ++j) before += q.cc[4 * pos + j] > 0;
                np = T - 1 - (cb + before);
            } else {
                np = T + q.surv[pos];
            }
            keynode[k] = (uint16_t)np;
        }
        __syncthreads();
        cur ^= 1;
        nlive = T + nsurv;
        // ---- termination / phase switch (:674-678, :739-740) ----
        if (tid == 0) {
            int st = 0;
            if (nlive >= N || nlive == prev_size) st = 2;
            else if (!careful && nlive + 3 * s_nexp > N) st = 1;
            s_state = st;
        }
        __syncthreads();
        const int st = s_state;
        __syncthreads();
        if (st == 2) break;
        if (st == 1) careful = true;
    }

    // ---- best key per node: largest response, earliest emission order on ties (:751-767) ----
    unsigned long long* best = reinterpret_cast<unsigned long long*>(q.cc);
    __syncthreads();
    for (int i = tid; i < nlive; i += nt) best[i] = 0ull;
    __syncthreads();
    for (int k = tid; k < ncand; k += nt) {
        const uint2 v = cand[k];
        const unsigned long long key =
            ((unsigned long long)(v.x >> 24) << 56) | ((unsigned long long)(0xffffffffu - v.y) << 24) | (unsigned)k;
        atomicMax(&best[keynode[k]], key);
    }
    __syncthreads();
    // ---- kept keypoints in list order + lapping prefix for the two-ended output fill (:1147-1156) ----
    OrbxKpRec* rec = ws.kprec + (long long)frame * ws.kp_stride + L.kp_off;
    const int nout = min(nlive, L.kp_cap);
    for (int i = tid; i < nout; i += nt) {
        const int k = (int)(best[i] & 0xffffffull);
        const uint32_t v = cand[k].x;
        const float x = (float)((int)(v & CM) + ORBX_FAST_BORDER);
        const float y = (float)((int)((v >> ORBX_COORD_BITS) & CM) + ORBX_FAST_BORDER);
        const float xs = level != 0 ? __fmul_rn(x, L.sf) : x;
        q.tmp[i] = (xs >= (float)plan.lap0 && xs <= (float)plan.lap1) ? 1 : 0;
        OrbxKpRec o;
        o.x = x; o.y = y; o.response = (float)(v >> 24); o.angle = -1.f; o.lap_before = 0; o.src = k;
        rec[i] = o;
    }
    __syncthreads();
    const int nlap = block_excl_scan(q.tmp, nout, s_scratch);
    for (int i = tid; i < nout; i += nt) rec[i].lap_before = q.tmp[i];
    if (tid == 0) ws.level_count[frame * plan.nlevels + level] = make_int2(nout, nlap);
}

// =================================================================================================
// K5a  GaussianBlur 7x7 sigma 2, OpenCV fixed-point path: 8.8 kernel {18,34,48,56,48,34,18},
// exact accumulation, one rounding (V + 32768) >> 16.  Reads the bordered plane (its reflect-101
// border is exactly the BORDER_REFLECT_101 extension of the borderless clone the reference blurs).
// =================================================================================================
#define ORBX_BLUR_TW 128
#define ORBX_BLUR_TH 32

__global__ void __launch_bounds__(256)
k_blur7(const __grid_constant__ OrbxPlan plan, const OrbxWs ws) {
    __shared__ __align__(16) uint8_t s_src[(ORBX_BLUR_TH + 6) * (ORBX_BLUR_TW + 8)];
    __shared__ __align__(16) uint16_t s_h[(ORBX_BLUR_TH + 6) * ORBX_BLUR_TW];
    int level = 0, tile = blockIdx.x;
    for (; level < plan.nlevels; ++level) {
        const int nt = ((plan.lv[level].w + ORBX_BLUR_TW - 1) / ORBX_BLUR_TW) * ((plan.lv[level].h + ORBX_BLUR_TH - 1) / ORBX_BLUR_TH);
        if (tile < nt) break;
        tile -= nt;
    }
    if (level >= plan.nlevels) return;
    const OrbxLevel& L = plan.lv[level];
    const int frame = blockIdx.y;
    if (ws.level_count[frame * plan.nlevels + level].x == 0) return;  // reference skips empty levels (:1122)
    const int tiles_x = (L.w + ORBX_BLUR_TW - 1) / ORBX_BLUR_TW;
    const int x0 = (tile % tiles_x) * ORBX_BLUR_TW, y0 = (tile / tiles_x) * ORBX_BLUR_TH;
    const uint8_t* plane = ws.pyr + (long long)frame * ws.pyr_stride + L.plane_off;
    const int tid = threadIdx.x;
    constexpr int SW = (ORBX_BLUR_TW + 8) / 4;  // words per staged row
    const int max_word = L.pitch / 4 - 1;
    for (int i = tid; i < (ORBX_BLUR_TH + 6) * SW; i += 256) {
        const int r = i / SW, wd = i - r * SW;
        const int pr = min(ORBX_EDGE + y0 + r - 3, L.plane_rows - 1);
        const int pw = min((ORBX_PADL + x0 - 4) / 4 + wd, max_word);
        reinterpret_cast<uint32_t*>(s_src)[i] = __ldg(reinterpret_cast<const uint32_t*>(plane + (long long)pr * L.pitch) + pw);
    }
    __syncthreads();
    // horizontal pass: s_h[r][x] = sum k[i] * src[r][x + i - 3]; src column of output x is x + 4
    for (int i = tid; i < (ORBX_BLUR_TH + 6) * (ORBX_BLUR_TW / 4); i += 256) {
        const int r = i / (ORBX_BLUR_TW / 4), xq = i - r * (ORBX_BLUR_TW / 4);
        const uint8_t* s = s_src + r * (ORBX_BLUR_TW + 8) + xq * 4 + 1;  // src[x-3] for x = 4*xq
        int v[10];
#pragma unroll
        for (int j = 0; j < 10; ++j) v[j] = s[j];
        uint16_t* hrow = s_h + r * ORBX_BLUR_TW + xq * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            hrow[j] = (uint16_t)(18 * (v[j] + v[j + 6]) + 34 * (v[j + 1] + v[j + 5]) + 48 * (v[j + 2] + v[j + 4]) + 56 * v[j + 3]);
    }
    __syncthreads();
    uint8_t* out = ws.blur + (long long)frame * ws.blur_stride + L.blur_off;
    for (int i = tid; i < ORBX_BLUR_TH * (ORBX_BLUR_TW / 4); i += 256) {
        const int r = i / (ORBX_BLUR_TW / 4), xq = i - r * (ORBX_BLUR_TW / 4);
        const int y = y0 + r, x = x0 + xq * 4;
        if (y >= L.h || x >= L.blur_pitch) continue;
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint16_t* h = s_h + r * ORBX_BLUR_TW + xq * 4 + j;
            const uint32_t acc = 18u * ((uint32_t)h[0] + h[6 * ORBX_BLUR_TW]) + 34u * ((uint32_t)h[ORBX_BLUR_TW] + h[5 * ORBX_BLUR_TW]) +
                                 48u * ((uint32_t)h[2 * ORBX_BLUR_TW] + h[4 * ORBX_BLUR_TW]) + 56u * (uint32_t)h[3 * ORBX_BLUR_TW];
            word |= ((acc + 32768u) >> 16) << (8 * j);
        }
        *reinterpret_cast<uint32_t*>(out + (long long)y * L.blur_pitch + x) = word;
    }
}

// =================================================================================================
// K4+K5b+K6  One warp per kept keypoint: IC_Angle (:75-102), rBRIEF (:105-145), scale-back and the
// two-ended scatter of operator() (:1143-1156).
// =================================================================================================
struct OrbxFloatConsts {
    float atan_p1, atan_p3, atan_p5, atan_p7;  // fastAtan2 polynomial, scaled to degrees
    float atan_eps;                            // (float)DBL_EPSILON
    float deg2rad;                             // factorPI = (float)(CV_PI/180.f), :104
};

__device__ __forceinline__ float fast_atan2_deg(float y, float x, const OrbxFloatConsts& fc) {
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, fc.atan_eps));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(fc.atan_p7, c2), fc.atan_p5), c2), fc.atan_p3), c2), fc.atan_p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, fc.atan_eps));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(fc.atan_p7, c2), fc.atan_p5), c2), fc.atan_p3), c2), fc.atan_p1), c));
    }
    if (x < 0.f) a = __fsub_rn(180.f, a);
    if (y < 0.f) a = __fsub_rn(360.f, a);
    return a;
}

#define ORBX_DESC_WARPS 8

__global__ void __launch_bounds__(ORBX_DESC_WARPS * 32)
k_describe(const __grid_constant__ OrbxPlan plan, const OrbxWs ws, const OrbxFloatConsts fc,
           void* __restrict__ kps_out, uint8_t* __restrict__ desc_out,
           int cap_per_frame, int32_t* __restrict__ counts, int frame_out0) {
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * ORBX_DESC_WARPS + (threadIdx.x >> 5);
    const int frame = blockIdx.y;
    const int2* lc = ws.level_count + frame * plan.nlevels;
    // level of this slot + totals
    int level = -1, idx = 0, n_before = 0, lap_before = 0, n_total = 0, lap_total = 0;
    for (int l = 0; l < plan.nlevels; ++l) {
        const int2 c = lc[l];
        const int off = plan.lv[l].kp_off;
        if (slot >= off && slot < off + plan.lv[l].kp_cap) {
            level = l; idx = slot - off; n_before = n_total; lap_before = lap_total;
        }
        n_total += c.x; lap_total += c.y;
    }
    const long long fo = (long long)(frame_out0 + frame);
    if (slot == 0 && lane == 0 && counts) {
        counts[2 * fo] = n_total;
        counts[2 * fo + 1] = n_total - lap_total;  // monoIndex, the reference's return value (:1161)
    }
    if (level < 0 || idx >= lc[level].x) return;
    const OrbxLevel& L = plan.lv[level];
    OrbxKpRec* recp = ws.kprec + (long long)frame * ws.kp_stride + L.kp_off + idx;
    const OrbxKpRec rec = *recp;
    const int cx = __float2int_rn(rec.x), cy = __float2int_rn(rec.y);

    // ---- IC_Angle on the un-blurred level ----
    const uint8_t* plane = ws.pyr + (long long)frame * ws.pyr_stride + L.plane_off;
    const uint8_t* ctr = plane + (long long)(ORBX_EDGE + cy) * L.pitch + ORBX_PADL + cx;
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - ORBX_HALF_PATCH;
        const int au = u < 0 ? -u : u;
        int colsum = 0;
#pragma unroll
        for (int v = -ORBX_HALF_PATCH; v <= ORBX_HALF_PATCH; ++v) {
            const int av = v < 0 ? -v : v;
            if (au <= plan.umax[av]) {
                const int val = ctr[(long long)v * L.pitch + u];
                colsum += val;
                m01 += v * val;
            }
        }
        m10 = u * colsum;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(ORBX_FULL_MASK, m10, o);
        m01 += __shfl_xor_sync(ORBX_FULL_MASK, m01, o);
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10, fc);

    // ---- rBRIEF on the blurred level: lane i computes descriptor byte i ----
    const float rad = __fmul_rn(angle, fc.deg2rad);
    const float a = (float)cos((double)rad), b = (float)sin((double)rad);
    const uint8_t* bl = ws.blur + (long long)frame * ws.blur_stride + L.blur_off + (long long)cy * L.blur_pitch + cx;
    const int4* pat = reinterpret_cast<const int4*>(ws.pattern) + lane * 2;
    const int4 p0 = __ldg(pat), p1 = __ldg(pat + 1);
    const int pw[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    int val = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x0 = (float)(int)(signed char)(pw[k] & 0xff), y0 = (float)(int)(signed char)((pw[k] >> 8) & 0xff);
        const float x1 = (float)(int)(signed char)((pw[k] >> 16) & 0xff), y1 = (float)(int)(signed char)((pw[k] >> 24) & 0xff);
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int c0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int c1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int t0 = bl[(long long)r0 * L.blur_pitch + c0];
        const int t1 = bl[(long long)r1 * L.blur_pitch + c1];
        val |= (t0 < t1) << k;
    }

    // ---- output slot: lapping keypoints fill from the back, the rest from the front ----
    const float xs = level != 0 ? __fmul_rn(rec.x, L.sf) : rec.x;
    const float ys = level != 0 ? __fmul_rn(rec.y, L.sf) : rec.y;
    const bool lapping = xs >= (float)plan.lap0 && xs <= (float)plan.lap1;
    const int laps_before = lap_before + rec.lap_before;
    const int out_idx = lapping ? (n_total - 1 - laps_before) : (n_before + idx - laps_before);
    if (lane == 0) recp->angle = angle;
    if (out_idx < cap_per_frame) {
        if (desc_out) desc_out[(fo * cap_per_frame + out_idx) * 32 + lane] = (uint8_t)val;
        if (kps_out && lane < 7) {
            float f;
            switch (lane) {
                case 0: f = xs; break;
                case 1: f = ys; break;
                case 2: f = L.kp_size; break;
                case 3: f = angle; break;
                case 4: f = rec.response; break;
                case 5: f = __int_as_float(level); break;
                default: f = __int_as_float(-1); break;
            }
            reinterpret_cast<float*>(kps_out)[(fo * cap_per_frame + out_idx) * 7 + lane] = f;
        }
    }
}

#endif  // ORBX_KERNELS_CUH_
