// extractorb_b200/csrc/orbx_kernels.cuh -- sm_100a kernels of the ORB extraction path.
//
//   k_pyr_level0 / k_pyr_resize : ComputePyramid          (reference ORBextractor.cc:1164-1219)
//   k_fast_tiles (k_fast_cells) : cell loop + cv::FAST    (:797-864); the round-1 kernel stays as an A/B reference (ORBX_FAST_V1=1)
//   k_octree                    : DistributeOctTree       (:544-771, DivideNode :486-542)
//   k_blur7                     : GaussianBlur 7x7 s=2    (:1126-1127)
//   k_describe                  : IC_Angle, rBRIEF, scale-back + two-ended scatter (:75-145, :1131-1159)
//   k_pyr_border                : copyMakeBorder of every level, only when a plane leaves the device (:1193, :1213)
//
// All pixel arithmetic is integer / fixed point and bit-exact with the OpenCV primitives the reference
// calls; the float steps (fastAtan2, pattern rotation, scale-back) use __f*_rn intrinsics so that nvcc
// never contracts them into FMAs (the reference is built without FMA, CMakeLists.txt:4-11).
#ifndef ORBX_KERNELS_CUH_
#define ORBX_KERNELS_CUH_

#include <cuda_pipeline.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/orbx.h"
#include "orbx_plan.h"

#define ORBX_FULL_MASK 0xffffffffu

__device__ __forceinline__ int reflect101(int p, int len) {
    p = p < 0 ? -p : p;
    return p >= len ? 2 * (len - 1) - p : p;
}

// TMA + mbarrier helpers (sm_90+ PTX; SASS: UTMALDG / SYNCS).  One thread arms the barrier with the byte count of the box and
// issues the bulk tensor copy; the copy engine lands the box in shared memory and completes the barrier's transaction count.
__device__ __forceinline__ void orbx_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void orbx_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void orbx_tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ bool orbx_mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}


// =================================================================================================
// K1  ComputePyramid.
// =================================================================================================
// Level 0 = copyMakeBorder(image, BORDER_REFLECT_101) (:1213): this kernel copies the image into the level-0
// ROI (16 bytes per thread; 128-bit loads when the source rows are 16-byte aligned), the 19-px border is written by k_pyr_border, and only when somebody will read it (see there).
__global__ void __launch_bounds__(256)
k_pyr_level0(const __grid_constant__ OrbxPlan plan, const OrbxWs ws, const uint8_t* __restrict__ imgs,
             long long row_stride, long long frame_stride, int aligned16) {
    const OrbxLevel& L = plan.lv[0];
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int frame = blockIdx.z;
    if (x >= L.w || y >= L.h) return;
    const uint8_t* srow = imgs + (long long)frame * frame_stride + (long long)y * row_stride + x;
    uint8_t* drow = ws.pyr + (long long)frame * ws.pyr_stride + L.plane_off + (long long)(ORBX_EDGE + y) * L.pitch + ORBX_PADL + x;
    if (aligned16 && x + 16 <= L.w) {
        *reinterpret_cast<uint4*>(drow) = __ldg(reinterpret_cast<const uint4*>(srow));
    } else {
        for (int j = 0; j < 16 && x + j < L.w; ++j) drow[j] = __ldg(srow + j);
    }
}

// cv::resize INTER_LINEAR, 8UC1 fixed point (11-bit coefficients), the model pinned in oracle/cv_prims.c:
//   H(y', x) = S[y'][sx]*a0 + S[y'][sx+1]*a1                      (int32, coefficients x2048)
//   dst(y,x) = ( ((b0*(H(sy,x)>>4))>>16) + ((b1*(H(sy+1,x)>>4))>>16) + 2 ) >> 2
// Separable and staged through shared memory: a CTA owns a 128x64 tile of the level (ROI only), stages the
// source window with 16-byte cp.async, runs the horizontal pass once per staged source row (H>>4 fits 16
// bits) and the vertical pass from shared memory (IMAD.HI against b<<16), and stores aligned words.
// Borders are filled afterwards by k_pyr_border (level l+1 never reads level l's border: edge taps clamp).
// AREA: exact 2x2 decimation, where OpenCV switches INTER_LINEAR to the INTER_AREA fast path.
#define ORBX_RS_TW 128
#define ORBX_RS_TH 64
#define ORBX_RS_TH_LAT 16

// Horizontal pass of one thread: destination column fixed, staged rows g, g+NG, ...  (NG = 0: run-time `ng`).
// PITCH = 0: run-time staging pitch; otherwise the pitch is a constant and every load has an immediate offset.
template <bool AREA, int NG, int PITCH>
__device__ __forceinline__ void rs_hpass(const uint8_t* __restrict__ q0, const uint8_t* __restrict__ q1, uint16_t* __restrict__ hq,
                                         int a0, int a1, int g, int ng_rt, int nrows, int pitch_rt) {
    const int ng = NG ? NG : ng_rt;
    const int step = ng * (PITCH ? PITCH : pitch_rt);
    int r = g;
#pragma unroll 1
    for (; r + 3 * ng < nrows; r += 4 * ng) {                     // four rows per iteration: loads first, then the math
        const int u00 = q0[0], u01 = q1[0], u10 = q0[step], u11 = q1[step];
        const int u20 = q0[2 * step], u21 = q1[2 * step], u30 = q0[3 * step], u31 = q1[3 * step];
        hq[0] = (uint16_t)(AREA ? (u00 + u01) : ((u00 * a0 + u01 * a1) >> 4));
        hq[ng * ORBX_RS_TW] = (uint16_t)(AREA ? (u10 + u11) : ((u10 * a0 + u11 * a1) >> 4));
        hq[2 * ng * ORBX_RS_TW] = (uint16_t)(AREA ? (u20 + u21) : ((u20 * a0 + u21 * a1) >> 4));
        hq[3 * ng * ORBX_RS_TW] = (uint16_t)(AREA ? (u30 + u31) : ((u30 * a0 + u31 * a1) >> 4));
        q0 += 4 * step; q1 += 4 * step; hq += 4 * ng * ORBX_RS_TW;
    }
    for (; r < nrows; r += ng) {
        const int u0 = q0[0], u1 = q1[0];
        hq[0] = (uint16_t)(AREA ? (u0 + u1) : ((u0 * a0 + u1 * a1) >> 4));
        q0 += step; q1 += step; hq += ng * ORBX_RS_TW;
    }
}

// Horizontal pass, two adjacent destination columns per thread (levels whose source columns advance by at most 2 per destination
// column, i.e. every scale factor <= 2): the <= 7 source bytes both columns need lie in two aligned words; one PRMT (selector fixed
// per thread) lines up {u0, u1} of column A in the low half and of column B in the high half, and each H = u0*a0 + u1*a1 is ONE
// IDP.2A against the column's {a0, a1} 16-bit pair (the x table's own format).  Nine instructions per pair and row instead of twelve.
template <int NG, int PITCH>
__device__ __forceinline__ void rs_hpass2(const uint32_t* __restrict__ q, uint32_t* __restrict__ hq, unsigned sel, unsigned ca, unsigned cb,
                                          int g, int ng_rt, int nrows, int pitch_rt) {
    const int ng = NG ? NG : ng_rt;
    const int step = ng * ((PITCH ? PITCH : pitch_rt) >> 2);         // words
    auto h2 = [&](uint32_t w0, uint32_t w1) -> uint32_t {
        const uint32_t v = __byte_perm(w0, w1, sel);
        const uint32_t ha = __dp2a_lo(ca, v, 0u), hb = __dp2a_hi(cb, v, 0u);
        return (ha >> 4) | ((hb << 12) & 0xffff0000u);
    };
    int r = g;
#pragma unroll 1
    for (; r + 3 * ng < nrows; r += 4 * ng) {                     // four rows per iteration: loads first, then the math
        const uint32_t a0 = q[0], a1 = q[1], b0 = q[step], b1 = q[step + 1];
        const uint32_t c0 = q[2 * step], c1 = q[2 * step + 1], d0 = q[3 * step], d1 = q[3 * step + 1];
        hq[0] = h2(a0, a1);
        hq[ng * (ORBX_RS_TW / 2)] = h2(b0, b1);
        hq[2 * ng * (ORBX_RS_TW / 2)] = h2(c0, c1);
        hq[3 * ng * (ORBX_RS_TW / 2)] = h2(d0, d1);
        q += 4 * step; hq += 4 * ng * (ORBX_RS_TW / 2);
    }
    for (; r < nrows; r += ng) {
        hq[0] = h2(q[0], q[1]);
        q += step; hq += ng * (ORBX_RS_TW / 2);
    }
}

#define ORBX_RS_PITCH 192     // staging pitch of the specialised instance (covers scale factors up to ~1.25)

// TH: tile height.  ORBX_RS_TH for throughput; ORBX_RS_TH_LAT for one or two frames, where a level is a handful of tiles and
// the kernel time is one CTA's latency (more, smaller CTAs finish sooner).
// Reflect-101 border words of one plane row (copyMakeBorder, :1193).  srow / drow: the source level row and the destination
// plane row as words, word 0 = level column 0.  A border word (4 pixels) is a byte-reversed, funnel-shifted window of the
// level row: left-strip pixels dx..dx+3 (dx < 0) mirror level pixels -dx-3..-dx, right-strip pixels mirror 2(w-1)-dx-3..2(w-1)-dx.
__device__ __forceinline__ void border_left_word(const uint32_t* __restrict__ srow, uint32_t* __restrict__ drow, int wx /* -5..-1 */) {
    const int k = -wx - 1;                                     // level words k, k+1 hold pixels 4k+1 .. 4k+4
    const uint32_t v = __funnelshift_r(srow[k], srow[k + 1], 8);
    drow[wx] = __byte_perm(v, 0, 0x0123);                      // pixel -20 (byte 0 of word -5) is padding
}
__device__ __forceinline__ void border_right_word(const uint32_t* __restrict__ srow, uint32_t* __restrict__ drow, int wx, int w) {
    const int dx0 = 4 * wx;
    const int s0 = 2 * (w - 1) - dx0 - 3;                      // first mirrored level pixel (>= 0 since w > 22)
    const int k = s0 >> 2;
    const uint32_t v = __funnelshift_r(srow[k], srow[k + 1], 8 * (s0 & 3));
    uint32_t word = __byte_perm(v, 0, 0x0123);
    const int keep = w - dx0;                                  // level bytes at the start of a word straddling level / border
    if (keep > 0) {
        const uint32_t orig = srow[wx];
        const uint32_t sel = keep == 1 ? 0x7650u : (keep == 2 ? 0x7610u : 0x7210u);
        word = __byte_perm(orig, word, sel);
    }
    drow[wx] = word;
}

// TMA: the source window (box = staging pitch x src_rows_max bytes of the previous level's plane, ws.tmaps_rs[level]) arrives as
// one bulk tensor copy instead of ~1000 16-byte cp.async.
template <bool AREA, int PITCH, int TH, bool TMA = false>
__global__ void __launch_bounds__(256)
k_pyr_resize(const __grid_constant__ OrbxPlan plan, const OrbxWs ws, int level, int src_rows_max, int src_pitch_rt, int pairs) {
    extern __shared__ __align__(128) uint8_t smem_rs[];
    __shared__ __align__(8) unsigned long long s_bar;
    const int src_pitch_s = PITCH ? PITCH : src_pitch_rt;
    const OrbxLevel& L = plan.lv[level];
    const OrbxLevel& S = plan.lv[level - 1];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * ORBX_RS_TW, y0 = blockIdx.y * TH;
    const int tw = min(ORBX_RS_TW, L.w - x0), th = min(TH, L.h - y0);   // this tile's extent
    const int frame = blockIdx.z;
    uint8_t* fbase = ws.pyr + (long long)frame * ws.pyr_stride;
    const uint8_t* splane = fbase + S.plane_off;
    const int2* xtab = ws.xtab + L.xtab_off;
    const int2* ytab = ws.ytab + L.ytab_off;
    // source window (inclusive), second taps clamped into the level (their weight is 0 when clamped)
    const int sx_lo = __ldg(&xtab[x0]).x, sx_hi = min(__ldg(&xtab[x0 + tw - 1]).x + 1, S.w - 1);
    const int sy_lo = __ldg(&ytab[y0]).x, sy_hi = min(__ldg(&ytab[y0 + th - 1]).x + 1, S.h - 1);
    const int nrows = sy_hi - sy_lo + 1;
    const int gcol = ORBX_PADL + sx_lo;            // plane byte column of the first staged pixel
    const int al = gcol & 15;
    const int nvec = (al + (sx_hi - sx_lo + 1) + 15) >> 4;
    uint8_t* s_src = smem_rs;                                           // [src_rows_max][src_pitch_s]
    uint8_t* s_hb = smem_rs + (size_t)src_rows_max * src_pitch_s;       // [src_rows_max][TW] uint16
    int4* s_yt = reinterpret_cast<int4*>(s_hb + (size_t)src_rows_max * ORBX_RS_TW * 2);   // [TH] row descriptors
    // ---- stage: one TMA box, or 16-byte cp.async (plane rows are 64-byte aligned), 16 or 32 lanes per source row ----
    {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        if (TMA) {
            if (tid == 0) {
                orbx_mbar_init(bar, 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                orbx_mbar_expect_tx(bar, (uint32_t)(src_rows_max * src_pitch_s));
                orbx_tma_load_3d((uint32_t)__cvta_generic_to_shared(s_src), ws.tmaps_rs + 128 * level, bar, gcol - al, ORBX_EDGE + sy_lo, frame);
            }
        } else {
            const uint8_t* g = splane + (long long)(ORBX_EDGE + sy_lo) * S.pitch + (gcol - al);
            const int sh = nvec <= 16 ? 4 : 5;
            const int v0 = tid & ((1 << sh) - 1), rstep = 256 >> sh;
            for (int v = v0; v < nvec; v += 1 << sh)
                for (int r = tid >> sh; r < nrows; r += rstep)
                    __pipeline_memcpy_async(s_src + r * src_pitch_s + 16 * v, g + (long long)r * S.pitch + 16 * v, 16);
            __pipeline_commit();
        }
        // destination-row descriptors: byte offsets of the two source rows inside s_h, vertical coefficients << 16
        if (tid < th) {
            const int2 yt = __ldg(&ytab[y0 + tid]);
            const int r0 = yt.x - sy_lo, r1 = min(yt.x + 1, S.h - 1) - sy_lo;
            s_yt[tid] = make_int4(r0 * (ORBX_RS_TW * 2), r1 * (ORBX_RS_TW * 2), (int)((unsigned)(yt.y & 0xffff) << 16),
                                  (int)((unsigned)(yt.y >> 16) << 16));
        }
        if (TMA) {
            __syncthreads();                               // the barrier is initialised before anybody polls it
            unsigned spins = 0;
            while (!orbx_mbar_try_wait(bar, 0)) {
                if (++spins > (1u << 18)) __trap();
            }
        } else {
            __pipeline_wait_prior(0);
        }
    }
    __syncthreads();
    // ---- horizontal pass: thread owns one destination column, walks the staged rows.  Full tiles: two row
    // groups of 128 columns; partial tiles: 256 / tw row groups of tw columns, so that no lane idles ----
    {
        uint16_t* s_h = reinterpret_cast<uint16_t*>(s_hb);
        if (!AREA && pairs) {
            // column pairs: 64 pairs x 4 row groups on a full tile, 256 / pairs row groups on a partial one
            const int np = (tw + 1) >> 1;
            const bool full = tw == ORBX_RS_TW;
            const int ng = full ? 4 : 256 / np;
            const int g = full ? tid >> 6 : tid / np, c2 = full ? tid & 63 : tid - g * np;
            if (g < ng) {
                const int ca = 2 * c2, cb = min(ca + 1, tw - 1);
                const int2 xa = __ldg(&xtab[x0 + ca]), xb = __ldg(&xtab[x0 + cb]);
                const int off = xa.x - sx_lo + al;                 // staged byte of column A's first tap
                const unsigned pa0 = off & 3, pa1 = pa0 + (min(xa.x + 1, S.w - 1) - xa.x);
                const unsigned pb0 = pa0 + (xb.x - xa.x), pb1 = pb0 + (min(xb.x + 1, S.w - 1) - xb.x);      // <= 6
                const unsigned sel = pa0 | (pa1 << 4) | (pb0 << 8) | (pb1 << 12);
                const uint32_t* q = reinterpret_cast<const uint32_t*>(s_src + g * src_pitch_s) + (off >> 2);
                uint32_t* hq = reinterpret_cast<uint32_t*>(s_h) + g * (ORBX_RS_TW / 2) + c2;
                if (full) rs_hpass2<4, PITCH>(q, hq, sel, (unsigned)xa.y, (unsigned)xb.y, g, 4, nrows, src_pitch_s);
                else rs_hpass2<0, PITCH>(q, hq, sel, (unsigned)xa.y, (unsigned)xb.y, g, ng, nrows, src_pitch_s);
            }
        } else if (tw == ORBX_RS_TW) {
            const int c = tid & (ORBX_RS_TW - 1), g = tid >> 7;
            const int2 xt = __ldg(&xtab[x0 + c]);
            const uint8_t* q0 = s_src + (xt.x - sx_lo + al) + g * src_pitch_s;
            const uint8_t* q1 = q0 + (min(xt.x + 1, S.w - 1) - xt.x);
            rs_hpass<AREA, 2, PITCH>(q0, q1, s_h + c + g * ORBX_RS_TW, xt.y & 0xffff, (xt.y >> 16) & 0xffff, g, 2, nrows, src_pitch_s);
        } else {
            const int ng = 256 / tw;
            const int g = tid / tw, c = tid - g * tw;
            if (g < ng) {
                const int2 xt = __ldg(&xtab[x0 + c]);
                const uint8_t* q0 = s_src + (xt.x - sx_lo + al) + g * src_pitch_s;
                const uint8_t* q1 = q0 + (min(xt.x + 1, S.w - 1) - xt.x);
                rs_hpass<AREA, 0, PITCH>(q0, q1, s_h + c + g * ORBX_RS_TW, xt.y & 0xffff, (xt.y >> 16) & 0xffff, g, ng, nrows, src_pitch_s);
            }
        }
    }
    __syncthreads();
    // ---- vertical pass: one aligned word of 4 destination pixels per item.  t = hi(p0*(b0<<16)) + hi(p1*(b1<<16)) + 2
    // < 1024, so the four (t >> 2) are packed as two 16-bit pairs, shifted once each and merged with one PRMT.
    // Full-width tiles: lane = word, warp w owns rows w, w+8, ... (pointers advance by constants).  Partial tiles:
    // items dealt out linearly so that every lane stays busy; the word that straddles the right edge spills into
    // the border columns, which k_pyr_border rewrites afterwards. ----
    {
        auto vword = [](uint2 h0, uint2 h1, unsigned b0, unsigned b1) -> uint32_t {
            unsigned t0, t1, t2, t3;
            if (AREA) {
                t0 = (h0.x & 0xffffu) + (h1.x & 0xffffu) + 2u; t1 = (h0.x >> 16) + (h1.x >> 16) + 2u;
                t2 = (h0.y & 0xffffu) + (h1.y & 0xffffu) + 2u; t3 = (h0.y >> 16) + (h1.y >> 16) + 2u;
            } else {
                t0 = __umulhi(h0.x & 0xffffu, b0) + __umulhi(h1.x & 0xffffu, b1) + 2u;
                t1 = __umulhi(h0.x >> 16, b0) + __umulhi(h1.x >> 16, b1) + 2u;
                t2 = __umulhi(h0.y & 0xffffu, b0) + __umulhi(h1.y & 0xffffu, b1) + 2u;
                t3 = __umulhi(h0.y >> 16, b0) + __umulhi(h1.y >> 16, b1) + 2u;
            }
            const unsigned lo = (t1 * 65536u + t0) >> 2, hi = (t3 * 65536u + t2) >> 2;
            return __byte_perm(lo, hi, 0x6420);
        };
        uint8_t* dtile = fbase + L.plane_off + (long long)(ORBX_EDGE + y0) * L.pitch + ORBX_PADL + x0;
        const unsigned dpitch = (unsigned)L.pitch;
        if (tw == ORBX_RS_TW) {
            const int lane = tid & 31, wrp = tid >> 5;
            const uint8_t* hb = s_hb + 8 * lane;
            const int4* ytp = s_yt + wrp;
            uint8_t* d = dtile + wrp * dpitch + 4 * lane;
#pragma unroll 2
            for (int row = wrp; row < th; row += 8) {
                const int4 yt = *ytp;
                const uint2 h0 = *reinterpret_cast<const uint2*>(hb + yt.x);
                const uint2 h1 = *reinterpret_cast<const uint2*>(hb + yt.y);
                *reinterpret_cast<uint32_t*>(d) = vword(h0, h1, (unsigned)yt.z, (unsigned)yt.w);
                ytp += 8; d += 8 * dpitch;
            }
        } else {
            const int tw4 = (tw + 3) >> 2, nitems = tw4 * th;
            const unsigned m = (1u << 20) / (unsigned)tw4 + 1u;           // i / tw4 for i < 2^11
            for (int i = tid; i < nitems; i += 256) {
                const unsigned row = ((unsigned)i * m) >> 20;
                const unsigned cg = (unsigned)i - row * (unsigned)tw4;
                const int4 yt = s_yt[row];
                const uint2 h0 = *reinterpret_cast<const uint2*>(s_hb + yt.x + 8 * cg);
                const uint2 h1 = *reinterpret_cast<const uint2*>(s_hb + yt.y + 8 * cg);
                *reinterpret_cast<uint32_t*>(dtile + (row * dpitch + 4 * cg)) = vword(h0, h1, (unsigned)yt.z, (unsigned)yt.w);
            }
        }
    }
}

// copyMakeBorder(BORDER_REFLECT_101) of every level (:1193, :1213).  Nothing on the extraction path reads the border: FAST, IC_Angle
// and the stereo windows stay inside the level, and k_blur7 mirrors its own 3-px halo.  Only a caller that takes mvImagePyramid
// out of the device sees it (the reference's level Mats are ROIs of bordered buffers, :1173-1177), so the host side launches
// this kernel on demand -- batch extraction never writes those 200 KB per frame.  16 lanes per plane row.  A border word
// (4 pixels) is a byte-reversed, funnel-shifted window of the level row: for the left strip pixels dx..dx+3
// (dx < 0) mirror level pixels -dx-3..-dx, for the right strip they mirror 2(w-1)-dx-3..2(w-1)-dx.  The word
// that straddles level and border keeps its level bytes.  Band rows (the 19 rows above / below the level)
// additionally copy the level columns from the reflected row.
#define ORBX_BORDER_ROWS 64
__global__ void __launch_bounds__(256)
k_pyr_border(const __grid_constant__ OrbxPlan plan, const OrbxWs ws) {
    const int level = blockIdx.y;
    const int frame = blockIdx.z;
    const OrbxLevel& L = plan.lv[level];
    const int lane = threadIdx.x;                              // 0..15
    uint8_t* plane = ws.pyr + (long long)frame * ws.pyr_stride + L.plane_off;
    // a CTA walks ORBX_BORDER_ROWS plane rows, 16 at a time (few, fatter CTAs: the kernel was bound by the CTA launch rate)
    const int r_end = min((int)(blockIdx.x + 1) * ORBX_BORDER_ROWS, L.plane_rows);
    for (int r = blockIdx.x * ORBX_BORDER_ROWS + threadIdx.y; r < r_end; r += 16) {
    const bool band = r < ORBX_EDGE || r >= ORBX_EDGE + L.h;
    // level row this plane row mirrors (itself for middle rows), as words relative to level column 0
    const uint32_t* srow = reinterpret_cast<const uint32_t*>(plane + (long long)(ORBX_EDGE + reflect101(r - ORBX_EDGE, L.h)) * L.pitch + ORBX_PADL);
    uint32_t* drow = reinterpret_cast<uint32_t*>(plane + (long long)r * L.pitch + ORBX_PADL);   // word 0 = level column 0
    if (band) {
        // level columns of a band row: 16-byte copies (level column 0 sits at plane byte 32 of a 64-byte aligned row), three
        // independent loads in flight per lane before the stores, then the last words one by one
        const int nw = L.w >> 2;                               // words fully inside the level
        const int nv = L.w >> 4;
        const uint4* s4 = reinterpret_cast<const uint4*>(srow);
        uint4* d4 = reinterpret_cast<uint4*>(drow);
        for (int base = lane; base < nv; base += 48) {
            const bool h1 = base + 16 < nv, h2 = base + 32 < nv;
            const uint4 t0 = s4[base];
            uint4 t1 = t0, t2 = t0;
            if (h1) t1 = s4[base + 16];
            if (h2) t2 = s4[base + 32];
            d4[base] = t0;
            if (h1) d4[base + 16] = t1;
            if (h2) d4[base + 32] = t2;
        }
        for (int wx = 4 * nv + lane; wx < nw; wx += 16) drow[wx] = srow[wx];
    }
    const int rw0 = L.w >> 2;                                  // first right-strip word (may straddle level/border)
    const int rwl = (L.w + ORBX_EDGE - 1) >> 2;                // last word holding border pixels
    const int nright = rwl - rw0 + 1;
    if (lane < 5) border_left_word(srow, drow, lane - 5);      // left strip: words -5..-1 (pixels -20..-1; pixel -20 is padding)
    else if (lane - 5 < nright) border_right_word(srow, drow, rw0 + (lane - 5), L.w);
    }
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

// best(p) with both sides evaluated at once: X_k = (255 + c - v_k) | (255 + v_k - c) << 16 is one IMAD per
// ring pixel (v_k * 65535 + B; both halves stay in [0, 510], so no borrow crosses the halves), and the
// max-over-arcs-of-min chain runs on 16-bit SIMD (VIMNMX3.S16x2): low half = "ring darker" side, high half =
// "ring brighter" side.  Arcs of 9 = min3 of three consecutive min3 triples.
__device__ __forceinline__ int fast_best(const uint8_t* __restrict__ t, int p, int tp) {
    const unsigned c = t[p];
    const unsigned B = 255u + (255u << 16) + c * (1u - 65536u);
    unsigned x[16];
    x[0] = t[p + 3 * tp] * 65535u + B;      x[1] = t[p + 3 * tp + 1] * 65535u + B;  x[2] = t[p + 2 * tp + 2] * 65535u + B;
    x[3] = t[p + tp + 3] * 65535u + B;      x[4] = t[p + 3] * 65535u + B;           x[5] = t[p - tp + 3] * 65535u + B;
    x[6] = t[p - 2 * tp + 2] * 65535u + B;  x[7] = t[p - 3 * tp + 1] * 65535u + B;  x[8] = t[p - 3 * tp] * 65535u + B;
    x[9] = t[p - 3 * tp - 1] * 65535u + B;  x[10] = t[p - 2 * tp - 2] * 65535u + B; x[11] = t[p - tp - 3] * 65535u + B;
    x[12] = t[p - 3] * 65535u + B;          x[13] = t[p + tp - 3] * 65535u + B;     x[14] = t[p + 2 * tp - 2] * 65535u + B;
    x[15] = t[p + 3 * tp - 1] * 65535u + B;
    unsigned m3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m3[k] = __vimin3_s16x2(x[k], x[(k + 1) & 15], x[(k + 2) & 15]);
    unsigned m9[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) m9[k] = __vimin3_s16x2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
    unsigned r0 = __vimax3_s16x2(m9[0], m9[1], m9[2]), r1 = __vimax3_s16x2(m9[3], m9[4], m9[5]);
    unsigned r2 = __vimax3_s16x2(m9[6], m9[7], m9[8]), r3 = __vimax3_s16x2(m9[9], m9[10], m9[11]);
    unsigned r4 = __vimax3_s16x2(m9[12], m9[13], m9[14]);
    r0 = __vimax3_s16x2(r0, r1, r2);
    r3 = __vimax3_s16x2(r3, r4, m9[15]);
    r0 = __vmaxs2(r0, r3);
    return (int)max(r0 & 0xffffu, r0 >> 16) - 255;
}

__global__ void __launch_bounds__(ORBX_FAST_WARPS * 32)
k_fast_cells(const __grid_constant__ OrbxPlan plan, const OrbxWs ws, const int cell_base, const int cell_end) {
    extern __shared__ __align__(16) uint8_t smem_fast[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ci = cell_base + blockIdx.x * ORBX_FAST_WARPS + wib;     // [cell_base, cell_end): all cells, or one level's slice
    const int frame = blockIdx.y;
    if (ci >= cell_end) return;
    OrbxCell cell;
    {
        const uint4* cp = reinterpret_cast<const uint4*>(ws.cells + ci);
        const uint4 c0 = __ldg(cp), c1 = __ldg(cp + 1);
        *reinterpret_cast<uint4*>(&cell) = c0;
        *(reinterpret_cast<uint4*>(&cell) + 1) = c1;
    }
    const OrbxLevel& L = plan.lv[cell.level];

    const int tp = plan.fast_tp;
    const int tile_bytes = (tp * plan.fast_trows + 15) & ~15;            // 16-byte aligned sub-buffers
    uint8_t* tile = smem_fast + (size_t)wib * ((2 * tile_bytes + 2 * plan.fast_qcap + 15) & ~15);
    uint8_t* score = tile + tile_bytes;
    uint16_t* queue = reinterpret_cast<uint16_t*>(score + tile_bytes);

    const int ch = cell.ch;
    // ---- stage the cell image (aligned 32-bit cp.async; pixel (r,c) lands at tile[r*tp + a + c]) while the
    // score map is cleared with 128-bit stores ----
    const uint8_t* plane = ws.pyr + (long long)frame * ws.pyr_stride + L.plane_off;
    const int gx = ORBX_PADL + cell.x0;
    const int a = gx & 3;
    const int nwords = cell.nwords;
    {
        const uint32_t* g32 = reinterpret_cast<const uint32_t*>(plane + (long long)(ORBX_EDGE + cell.y0) * L.pitch + (gx - a));
        const int pw = L.pitch >> 2, tpw4 = tp >> 2;
        const int total = nwords * ch;                                   // < 2^13
        const unsigned wmagic = cell.wmagic;                             // i / nwords == (i * wmagic) >> 20
        for (int i = lane; i < total; i += 32) {
            const int r = (int)(((unsigned)i * wmagic) >> 20);
            const int wd = i - r * nwords;
            __pipeline_memcpy_async(reinterpret_cast<uint32_t*>(tile) + r * tpw4 + wd, g32 + r * pw + wd, 4);
        }
        __pipeline_commit();
        const int nz = (ch * tp + 15) >> 4;
        for (int i = lane; i < nz; i += 32) reinterpret_cast<uint4*>(score)[i] = make_uint4(0, 0, 0, 0);
        __pipeline_wait_prior(0);
    }
    __syncwarp();

    const int ih = ch - 6;
    // 4-pixel groups (aligned words) covering the interior columns [a+3, a+cw-3)
    const int wq0 = cell.wq0;
    const int ngrp = cell.ngrp;
    const int ngroups = ngrp * ih;
    const unsigned gmagic = cell.gmagic;                       // idx / ngrp == (idx * gmagic) >> 24 (idx < 2^13)
    const int wq_last = wq0 + ngrp - 1;
    // per-word pixel flags live at bits {0, 16, 1, 17} (pixels 0..3): spread the host's 4-bit edge masks likewise
    auto spread = [](unsigned m) { return (m & 1u) | ((m & 2u) << 15) | ((m & 4u) >> 1) | ((m & 8u) << 14); };
    const unsigned first_mask = spread(cell.masks & 0xFu), last_mask = spread(cell.masks >> 4), full_mask = 0x00030003u;
    // The reference detects at iniThFAST and re-runs the cell at minThFAST only if that found nothing
    // (:818-838); the same two phases here, the second one skipped when it cannot add anything.
    int qn = 0, n_out = 0, use_th = plan.ini_th;
    for (int phase = 0; phase < 2; ++phase) {
        if (phase == 1) {
            if (plan.min_th >= plan.ini_th) break;
            use_th = plan.min_th;
        }
        // ---- stage 1: branch-free bound from the vertical + horizontal ring pairs, 4 pixels per thread.
        // A group is one aligned 32-bit word of row r; its up/down neighbours (rows r-3, r+3) are the same
        // word of those rows, its left/right neighbours (x-3, x+3) come from a funnel shift of the row's
        // adjacent words.  Bounds are evaluated two pixels at a time with 16-bit SIMD min/max.
        qn = 0;
        {
            const unsigned T1 = (unsigned)(use_th + 1) * 0x00010001u;
            const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tile);
            const int tpw = tp >> 2;
            for (int base = 0; base < ngroups; base += 64) {
                unsigned mk[2] = {0u, 0u};
                int pb[2] = {0, 0};
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int idx = base + 32 * g + lane;
                    if (idx < ngroups) {
                        const int gr = (int)(((unsigned)idx * gmagic) >> 24);
                        const int wq = wq0 + (idx - gr * ngrp);
                        const int wi = (gr + 3) * tpw + wq;
                        const unsigned wc = t32[wi], wl = t32[wi - 1], wr = t32[wi + 1];
                        const unsigned wu = t32[wi - 3 * tpw], wd = t32[wi + 3 * tpw];
                        const unsigned lft = __funnelshift_r(wl, wc, 8), rgt = __funnelshift_r(wc, wr, 24);
                        unsigned z[2];
#pragma unroll
                        for (int hlf = 0; hlf < 2; ++hlf) {
                            const unsigned sel = hlf ? 0x4342u : 0x4140u;
                            const unsigned c2 = __byte_perm(wc, 0, sel), u2 = __byte_perm(wu, 0, sel), d2 = __byte_perm(wd, 0, sel);
                            const unsigned l2 = __byte_perm(lft, 0, sel), r2 = __byte_perm(rgt, 0, sel);
                            const unsigned dk = __vmaxs2(__vmins2(u2, d2), __vmins2(l2, r2));
                            const unsigned br = __vmins2(__vmaxs2(u2, d2), __vmaxs2(l2, r2));
                            // lane-wise (c - dk > th) | (br - c > th): bit 15 of (x|0x8000) - (y + th + 1) is set iff x - y > th
                            z[hlf] = (((c2 | 0x80008000u) - (dk + T1)) | ((br | 0x80008000u) - (c2 + T1))) & 0x80008000u;
                        }
                        // flags of pixels 0,1 at bits 0,16 and of pixels 2,3 at bits 1,17; keep interior pixels only
                        const unsigned valid = (wq == wq0 ? first_mask : full_mask) & (wq == wq_last ? last_mask : full_mask);
                        mk[g] = ((z[0] >> 15) | (z[1] >> 14)) & valid;
                        pb[g] = (gr + 3) * tp + 4 * wq;
                    }
                }
                // queue slots by warp prefix sum of the per-lane counts (queue order is irrelevant: every entry
                // carries its own position)
                const int cnt = __popc(mk[0]) + __popc(mk[1]);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t_ = __shfl_up_sync(ORBX_FULL_MASK, inc, o);
                    if (lane >= o) inc += t_;
                }
                int pos = qn + inc - cnt;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (mk[g] & 0x00000001u) queue[pos++] = (uint16_t)(pb[g] + 0);
                    if (mk[g] & 0x00010000u) queue[pos++] = (uint16_t)(pb[g] + 1);
                    if (mk[g] & 0x00000002u) queue[pos++] = (uint16_t)(pb[g] + 2);
                    if (mk[g] & 0x00020000u) queue[pos++] = (uint16_t)(pb[g] + 3);
                }
                qn += __shfl_sync(ORBX_FULL_MASK, inc, 31);
            }
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
                    const int nb = max(max(max((int)score[p - 1], (int)score[p + 1]), max((int)score[p - tp - 1], (int)score[p - tp])),
                                       max(max((int)score[p - tp + 1], (int)score[p + tp - 1]), max((int)score[p + tp], (int)score[p + tp + 1])));
                    lm = s > nb;
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
        if (lane == 0) atomicOr(ws.flags, 1);
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
// K2 (v2)  The same cell loop, one CTA per TILE = up to ORBX_FT_MAXC consecutive cells of one cell row.
//
// k_fast_cells spends most of its instructions around the arithmetic: a warp per cell stages 36x36 pixels for a 30x30
// interior with 4-byte copies and a division per word, every per-cell prologue instruction costs a full warp for 900
// pixels, queues of ~150 entries and candidate lists of ~17 leave lanes idle, and the queue is walked three times.
// Here a CTA of 256 threads owns one image of ~223x38 pixels (16-byte cp.async rows, no per-word index arithmetic),
// shares one queue (full warps in every phase), and replaces the queue walks of NMS / emission by one dense pass
// over the score map (95 % of its words are zero: one LDS + one test per 4 pixels).
//
// Semantics kept exactly (ORBextractor.cc:797-864): cv::FAST(iniThFAST, nonmax) per cell; a cell that yields no keypoint is
// run again with minThFAST.  NMS never looks across a cell's detection interior (neighbours outside count as 0): the dense
// pass masks the left / right neighbours at cell boundary columns; the rows above / below the interior are never scored.
// "Cell yields nothing at iniThFAST" is decided after NMS (two equal adjacent maxima suppress each other), per cell, by a
// flag word; only the empty cells are then scanned again at minThFAST.
// =================================================================================================
#define ORBX_FT_THREADS 256

// Pair-bound flags of the 4 pixels of one aligned word: bit 15 / 31 of the two returned words = pixels (0,1) / (2,3).
// c, u, d: the word and the same word 3 rows up / down; l, r: the word shifted by 3 pixels left / right.
// KT = (0x8000 - (th + 1)) in both halves: bit 15 of x - y + KT is set iff x - y > th (|x - y| <= 255, no borrow crosses halves).
__device__ __forceinline__ void ft_word_flags(unsigned c, unsigned u, unsigned d, unsigned l, unsigned r, unsigned KT, unsigned& z0, unsigned& z1) {
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
        const unsigned sel = hlf ? 0x4342u : 0x4140u;
        const unsigned c2 = __byte_perm(c, 0, sel), u2 = __byte_perm(u, 0, sel), d2 = __byte_perm(d, 0, sel);
        const unsigned l2 = __byte_perm(l, 0, sel), r2 = __byte_perm(r, 0, sel);
        const unsigned dk = __vmaxs2(__vmins2(u2, d2), __vmins2(l2, r2));     // every arc of 9 holds one pixel of each antipodal pair
        const unsigned br = __vmins2(__vmaxs2(u2, d2), __vmaxs2(l2, r2));
        const unsigned z = (c2 - dk + KT) | (br - c2 + KT);
        if (hlf) z1 = z; else z0 = z;
    }
}

// Flags of one aligned 8-byte pair (8 pixels) at 32-bit word index b of the tile: pixel k of the pair -> bit 8k+7 (k < 4),
// bit 8(k-4)+3 (k >= 4).  The sign bits (15, 31) of the four flag words sit in their bytes 1 and 3: two PRMTs gather them as
// the top bits of 4 + 4 bytes, one shift interleaves the second word, and the caller's valid-pixel mask (a subset of
// 0x88888888) removes everything else.
#define ORBX_FT_FULL 0x88888888u
__device__ __forceinline__ unsigned ft_pair_flags(const uint32_t* __restrict__ t32, int b, int tpw, unsigned KT, unsigned valid) {
    const unsigned wl = t32[b - 1], wr = t32[b + 2];
    const uint2 wc = *reinterpret_cast<const uint2*>(t32 + b);
    const uint2 wu = *reinterpret_cast<const uint2*>(t32 + b - 3 * tpw);
    const uint2 wd = *reinterpret_cast<const uint2*>(t32 + b + 3 * tpw);
    unsigned z00, z01, z10, z11;
    ft_word_flags(wc.x, wu.x, wd.x, __funnelshift_r(wl, wc.x, 8), __funnelshift_r(wc.x, wc.y, 24), KT, z00, z01);
    ft_word_flags(wc.y, wu.y, wd.y, __funnelshift_r(wc.x, wc.y, 8), __funnelshift_r(wc.y, wr, 24), KT, z10, z11);
    const unsigned A = __byte_perm(z00, z01, 0x7531), B = __byte_perm(z10, z11, 0x7531) >> 4;
    return ((A & 0x80808080u) | (B & 0x08080808u)) & valid;
}

// 8-bit pixel mask of a pair -> the flag layout above
__device__ __forceinline__ unsigned ft_spread8(unsigned m) {
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s |= ((m >> k) & 1u) << (k < 4 ? 8 * k + 7 : 8 * (k - 4) + 3);
    return s;
}

// Queue the flagged pixels of one pair; p0 = tile byte index of the pair's first pixel, qa = shared-memory byte address of the
// next free queue slot (a running address: test, value, predicated store, predicated bump per pixel).
__device__ __forceinline__ unsigned ft_queue_pair(unsigned qa, unsigned mk, int p0) {
#define ORBX_FT_PUSH(bit, k)                                                                             \
    if (mk & (bit)) {                                                                                    \
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(qa), "h"((unsigned short)(p0 + (k))) : "memory"); \
        qa += 2;                                                                                         \
    }
    ORBX_FT_PUSH(0x00000080u, 0) ORBX_FT_PUSH(0x00008000u, 1) ORBX_FT_PUSH(0x00800000u, 2) ORBX_FT_PUSH(0x80000000u, 3)
    ORBX_FT_PUSH(0x00000008u, 4) ORBX_FT_PUSH(0x00000800u, 5) ORBX_FT_PUSH(0x00080000u, 6) ORBX_FT_PUSH(0x08000000u, 7)
#undef ORBX_FT_PUSH
    return qa;
}

// TMA: the tile image arrives as ONE bulk tensor copy (box = ft_tp bytes x ft_trows rows of the level's plane, described by
// ws.tmaps[level]; columns / rows past the plane are zero-filled) instead of ~570 16-byte cp.async with their index arithmetic.
// TP: the tile row pitch as a compile-time constant (256 bytes for the reference's 30-pixel cells: the ring and neighbour offsets of
// the exact measure and the NMS become immediates, row / column of a queue entry a shift and a mask); 0 = the plan's value.
template <bool TMA, int TP = 0>
__global__ void __launch_bounds__(ORBX_FT_THREADS)
k_fast_tiles(const __grid_constant__ OrbxPlan plan, const OrbxWs ws, const int tile_base) {
    extern __shared__ __align__(128) uint8_t smem_ft[];
    __shared__ int s_qn, s_flags;
    __shared__ __align__(8) unsigned long long s_bar;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int frame = blockIdx.y;
    OrbxFastTile T;
    {
        const uint4* tp4 = reinterpret_cast<const uint4*>(ws.tiles + tile_base + blockIdx.x);
        uint4* d = reinterpret_cast<uint4*>(&T);
        d[0] = __ldg(tp4); d[1] = __ldg(tp4 + 1); d[2] = __ldg(tp4 + 2);
    }
    const OrbxLevel& L = plan.lv[T.level];
    const int tp = TP ? TP : plan.ft_tp, tpw = tp >> 2;
    const int map_bytes = tp * plan.ft_trows;                       // multiple of 16
    uint8_t* tile = smem_ft;
    const int map_pitch = (map_bytes + 127) & ~127;                 // both maps are TMA destinations: 128-byte aligned
    uint8_t* score = smem_ft + map_pitch;
    uint16_t* queue = reinterpret_cast<uint16_t*>(smem_ft + 2 * map_pitch);
    const unsigned queue_sa = (unsigned)__cvta_generic_to_shared(queue);
    const int th_rows = T.th, tw = T.tw;
    const int a16 = (ORBX_PADL + T.x0) & 15;

    // ---- stage the tile image: whole 16-byte chunks of the plane rows (64-byte aligned), score map cleared meanwhile ----
    if (TMA) {
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        if (tid == 0) {
            orbx_mbar_init(bar, 1);                                 // init (+ its fence) and the copies come from the same thread
            // two boxes on one barrier: the tile image, and a box that lies entirely below the plane -- the copy engine fills what is
            // outside the tensor with zeros, which clears the score map without a single store instruction
            orbx_mbar_expect_tx(bar, 2u * (uint32_t)map_bytes);
            orbx_tma_load_3d((uint32_t)__cvta_generic_to_shared(tile), ws.tmaps + 128 * T.level, bar, ORBX_PADL + T.x0 - a16, ORBX_EDGE + T.y0, frame);
            orbx_tma_load_3d((uint32_t)__cvta_generic_to_shared(score), ws.tmaps + 128 * T.level, bar, 0, 1 << 20, frame);
            s_qn = 0; s_flags = 0;
        }
        __syncthreads();                                            // the barrier exists (and the counters are set) before anybody polls it
        unsigned spins = 0;
        while (!orbx_mbar_try_wait(bar, 0)) {
            if (++spins > (1u << 18)) __trap();                     // a copy that never lands must fail loudly, not hang the device
        }
        // every thread has seen the barrier complete: both maps are visible to it, no second block barrier needed
    } else {
        const uint8_t* g = ws.pyr + (long long)frame * ws.pyr_stride + L.plane_off + (long long)(ORBX_EDGE + T.y0) * L.pitch + (ORBX_PADL + T.x0 - a16);
        const int nvec = (a16 + tw + 15) >> 4;
        const int total = nvec * th_rows;
        for (int i = tid; i < total; i += ORBX_FT_THREADS) {
            const int r = (int)__umulhi((unsigned)i, T.vmagic);
            const int v = i - r * nvec;
            __pipeline_memcpy_async(tile + r * tp + 16 * v, g + (long long)r * L.pitch + 16 * v, 16);
        }
        __pipeline_commit();
        const int nz = (th_rows * tp) >> 4;
        for (int i = tid; i < nz; i += ORBX_FT_THREADS) reinterpret_cast<uint4*>(score)[i] = make_uint4(0, 0, 0, 0);
        if (tid == 0) { s_qn = 0; s_flags = 0; }
        __pipeline_wait_prior(0);
        __syncthreads();
    }

    const uint32_t* t32 = reinterpret_cast<const uint32_t*>(tile);
    const int nrows = th_rows - 6;
    const int ncells = T.ncells, wcell = T.wcell;
    const unsigned all_cells = (1u << ncells) - 1u;
    const unsigned lt_mask = (1u << lane) - 1u;

    unsigned empty = all_cells;                                        // cells still without a keypoint
    for (int phase = 0; phase < 2; ++phase) {
        int use_th = plan.ini_th;
        if (phase == 1) {
            if (plan.min_th >= plan.ini_th || empty == 0u) break;      // a lower threshold cannot add to a cell that has keypoints
            use_th = plan.min_th;
        }
        const unsigned KT = (0x8000u - (unsigned)(use_th + 1)) * 0x00010001u;
        // ---- stage 1: pair bound on aligned 8-pixel pairs, survivors queued (two pairs per thread and scan) ----
        if (phase == 0) {
            const int nitems = T.nitems, npairs = T.npairs, pq0 = T.pq0;
            for (int base = 0; base < nitems; base += 2 * ORBX_FT_THREADS) {
                unsigned mk[2] = {0u, 0u};
                int p0[2] = {0, 0};
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int i = base + g * ORBX_FT_THREADS + tid;
                    if (i < nitems) {
                        const int r = (int)__umulhi((unsigned)i, T.pmagic);
                        const int k = i - r * npairs;
                        const int b = (r + 3) * tpw + 2 * (pq0 + k);
                        const unsigned valid = (k == 0 ? T.first_mask : ORBX_FT_FULL) & (k == npairs - 1 ? T.last_mask : ORBX_FT_FULL);
                        mk[g] = ft_pair_flags(t32, b, tpw, KT, valid);
                        p0[g] = 4 * b;
                    }
                }
                const int cnt = __popc(mk[0]) + __popc(mk[1]);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t_ = __shfl_up_sync(ORBX_FULL_MASK, inc, o);
                    if (lane >= o) inc += t_;
                }
                int wbase = 0;
                if (lane == 31 && inc > 0) wbase = atomicAdd(&s_qn, inc);
                wbase = __shfl_sync(ORBX_FULL_MASK, wbase, 31);
                unsigned qa = queue_sa + 2u * (unsigned)(wbase + inc - cnt);
                qa = ft_queue_pair(qa, mk[0], p0[0]);
                ft_queue_pair(qa, mk[1], p0[1]);
            }
        } else {
            // only the cells that came out empty: items = (empty cell, row, pair of the cell's column span)
            const int ne = __popc(empty);
            const int ppc = (wcell >> 3) + 2;                          // a span of wcell bytes touches at most this many 8-byte pairs
            const unsigned per_cell = (unsigned)(nrows * ppc);
            const unsigned m_cell = 0xffffffffu / per_cell + 1u, m_ppc = 0xffffffffu / (unsigned)ppc + 1u;   // exact for i < 2^15
            const int nitems = ne * (int)per_cell;
            const int hi_all = a16 + tw - 4;
            for (int base = 0; base < nitems; base += ORBX_FT_THREADS) {
                const int i = base + tid;
                unsigned mk = 0u;
                int p0 = 0;
                if (i < nitems) {
                    const int e = (int)__umulhi((unsigned)i, m_cell);
                    const int rem = i - e * (int)per_cell;
                    const int r = (int)__umulhi((unsigned)rem, m_ppc);
                    const int k = rem - r * ppc;
                    const int c = (int)__fns(empty, 0, e + 1);          // e-th empty cell
                    const int lo = a16 + 3 + c * wcell, hi = min(lo + wcell - 1, hi_all);
                    const int pq = (lo >> 3) + k;
                    if (8 * pq <= hi) {
                        const unsigned m8 = (0xffu << max(lo - 8 * pq, 0)) & (0xffu >> max(8 * pq + 7 - hi, 0)) & 0xffu;
                        const int b = (r + 3) * tpw + 2 * pq;
                        mk = ft_pair_flags(t32, b, tpw, KT, ft_spread8(m8));
                        p0 = 4 * b;
                    }
                }
                const int cnt = __popc(mk);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t_ = __shfl_up_sync(ORBX_FULL_MASK, inc, o);
                    if (lane >= o) inc += t_;
                }
                int wbase = 0;
                if (lane == 31 && inc > 0) wbase = atomicAdd(&s_qn, inc);
                wbase = __shfl_sync(ORBX_FULL_MASK, wbase, 31);
                ft_queue_pair(queue_sa + 2u * (unsigned)(wbase + inc - cnt), mk, p0);
            }
        }
        __syncthreads();
        // From here on every warp owns one contiguous chunk of the queue (whole rounds of 32 slots) and compacts it in place:
        // the pixels that pass the exact measure move to the front of the chunk, then the NMS survivors do.  A round writes only
        // slots it (or an earlier round) has already read, so no second buffer and no block-wide counters are needed, and the
        // lanes of the next pass are dense again.
        const int qn = s_qn;
        const int chunk = ((qn + ORBX_FT_THREADS - 1) >> 8) << 5;
        const int cbeg = wid * chunk, cend = min(cbeg + chunk, qn);
        // ---- stage 2: exact corner measure of the queued pixels ----
        int na = 0;
        for (int e0 = cbeg; e0 < cend; e0 += 32) {
            const int e = e0 + lane;
            int p = 0;
            bool alive = false;
            if (e < cend) {
                p = queue[e];
                const int best = fast_best(tile, p, tp);
                alive = best > use_th;
                if (alive) score[p] = (uint8_t)best;
            }
            const unsigned bal = __ballot_sync(ORBX_FULL_MASK, alive);
            if (alive) queue[cbeg + na + __popc(bal & lt_mask)] = (uint16_t)p;
            na += __popc(bal);
        }
        __syncthreads();                                                    // the score map is complete
        // ---- strict 3x3 non-max suppression inside the cell's detection interior (pixels at or below the threshold count as 0,
        // like cv::FAST; the neighbours across a cell boundary column do not count) ----
        int nsv = 0;
        unsigned cflags = 0u;
        for (int i0 = 0; i0 < na; i0 += 32) {
            const int i = i0 + lane;
            int p = 0;
            bool keep = false;
            if (i < na) {
                p = queue[cbeg + i];
                const int s = score[p];
                const int r = TP == 256 ? (p >> 8) : (int)__umulhi((unsigned)p, plan.ft_tpmagic);
                const int xi = p - r * tp - a16 - 3;                         // interior column
                const int c = (int)__umulhi((unsigned)xi, T.cmagic);
                const int rem = xi - c * wcell;
                int nb = max((int)score[p - tp], (int)score[p + tp]);
                if (rem != 0) nb = max(nb, max((int)score[p - 1], max((int)score[p - tp - 1], (int)score[p + tp - 1])));
                if (rem != wcell - 1) nb = max(nb, max((int)score[p + 1], max((int)score[p - tp + 1], (int)score[p + tp + 1])));
                keep = s > nb;
                if (keep) cflags |= 1u << c;
            }
            const unsigned bal = __ballot_sync(ORBX_FULL_MASK, keep);
            if (keep) queue[cbeg + nsv + __popc(bal & lt_mask)] = (uint16_t)p;
            nsv += __popc(bal);
        }
        cflags = __reduce_or_sync(ORBX_FULL_MASK, cflags);
        if (lane == 0 && cflags) atomicOr(&s_flags, (int)cflags);
        // ---- emission: one atomic per warp reserves the slots; entries carry the emission-order key (:855-860).  (One atomic
        // per tile through a shared counter and two more barriers was slower, also at 4K: 507 vs 478 us, 2.09 vs 1.98 ms.) ----
        if (nsv > 0) {
            int slot0 = 0;
            if (lane == 0) {
                slot0 = atomicAdd(ws.cand_count + frame * plan.nlevels + T.level, nsv);
                if (slot0 + nsv > L.cand_cap) atomicOr(ws.flags, 1);
            }
            slot0 = __shfl_sync(ORBX_FULL_MASK, slot0, 0);
            __syncwarp();
            uint2* cand = ws.cand + (long long)frame * ws.cand_stride + L.cand_off;
            for (int i = lane; i < nsv; i += 32) {
                const int slot = slot0 + i;
                if (slot >= L.cand_cap) break;
                const int p = queue[cbeg + i];
                const int r = TP == 256 ? (p >> 8) : (int)__umulhi((unsigned)p, plan.ft_tpmagic);
                const int x = p - r * tp - a16;                              // tile image column (>= 3)
                const int c = (int)__umulhi((unsigned)(x - 3), T.cmagic);
                const int xc = x - c * wcell;                                // column inside the cell image
                const uint32_t resp = (uint32_t)score[p] - 1u;
                cand[slot] = make_uint2((uint32_t)(x + T.xoff) | ((uint32_t)(r + T.yoff) << ORBX_COORD_BITS) | (resp << 24),
                                        (T.ordbase + ((uint32_t)c << ORBX_ORD_CELL_SHIFT)) | ((uint32_t)r << 7) | (uint32_t)xc);
            }
        }
        __syncthreads();                                                     // every warp has reported its cells
        empty = all_cells & ~(unsigned)s_flags;
        if (phase == 0 && plan.min_th < plan.ini_th && empty != 0u) {
            __syncthreads();                                                 // everyone has read s_qn / s_flags
            if (tid == 0) s_qn = 0;
            __syncthreads();
        }
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
#define ORBX_QT_THREADS 128      // large launch groups (throughput) whose largest remaining level is below ORBX_QT_MID_PIXELS
#define ORBX_QT_THREADS_MID 256  // large launch groups otherwise (the mid levels of a 4K pyramid hold thousands of candidates each)
#define ORBX_QT_THREADS_LAT 512  // small launch groups: one CTA per level is latency bound
#define ORBX_QT_THREADS_BIG 1024 // levels of >= ORBX_QT_BIG_PIXELS pixels (tens of thousands of candidates per quadtree)
#define ORBX_QT_BIG_PIXELS 1000000
#define ORBX_QT_MID_PIXELS 600000

__device__ __forceinline__ int qt_quadrant(short4 r, int x, int y) {
    const int mx = r.x + ((r.y - r.x + 1) >> 1);  // UL.x + ceil((UR.x-UL.x)/2), :488
    const int my = r.z + ((r.w - r.z + 1) >> 1);
    return (x < mx ? 0 : 1) + (y < my ? 0 : 2);   // n1, n2, n3, n4 of :517-531
}

// Exclusive scan of v[0..n) (in place); returns the total to every thread.  Two block barriers: the warp totals go through one of
// two scratch rows (`*flip` alternates them, so a call never overwrites the row the slowest warp of the previous call may still be
// reading), and every warp scans the <= 32 warp totals itself with shuffles.
__device__ int block_excl_scan(int* v, int n, int* scratch /* [2][32] */, int* flip) {
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
    int* row = scratch + 32 * (*flip);
    *flip ^= 1;
    if (lane == 31) row[wid] = inc;
    __syncthreads();
    const int w = lane < nw ? row[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(ORBX_FULL_MASK, winc, o);
        if (lane >= o) winc += t;
    }
    const int total = __shfl_sync(ORBX_FULL_MASK, winc, 31);
    int run = __shfl_sync(ORBX_FULL_MASK, winc - w, wid) + inc - sum;
    for (int i = beg; i < end; ++i) {
        const int x = v[i];
        v[i] = run;
        run += x;
    }
    __syncthreads();
    return total;
}

// Bytes of node storage per quadtree (rect x2, cnt x2, child counts x2, seq / tmp / cbase / surv, sort keys padded to a power
// of two).  Small trees keep it in shared memory; a tree too big for that (tens of thousands of features on one level) runs
// from a per-(frame, level) block of ws.qt_scratch instead -- same code, generic pointers.
#define ORBX_QT_NODE_BYTES 72
__host__ __device__ inline size_t orbx_qt_bytes(int nc) {
    int sk = 1;
    while (sk < nc) sk <<= 1;
    return (size_t)nc * ORBX_QT_NODE_BYTES + (size_t)sk * 8;
}

// One sweep over a tree's candidates with several independent global loads in flight per thread (a quadtree is one CTA: its
// candidate list lives in L2 / HBM, and the sweep is bound by load latency, not by instructions).  f(k, node, xy) per candidate.
#define ORBX_QT_MLP 4
template <int NT, typename F>
__device__ __forceinline__ void qt_sweep(const uint2* __restrict__ cand, const uint16_t* __restrict__ keynode, int ncand, F f) {
    for (int b0 = 0; b0 < ncand; b0 += ORBX_QT_MLP * NT) {           // uniform trip count: f may use warp-wide primitives
        const int base = b0 + (int)threadIdx.x;
        int pos[ORBX_QT_MLP];
        uint32_t v[ORBX_QT_MLP];
#pragma unroll
        for (int j = 0; j < ORBX_QT_MLP; ++j) {
            const int k = base + j * NT;
            pos[j] = k < ncand ? (int)keynode[k] : -1;
            v[j] = k < ncand ? cand[k].x : 0u;
        }
#pragma unroll
        for (int j = 0; j < ORBX_QT_MLP; ++j) f(base + j * NT, pos[j], v[j]);     // f is called by every lane (pos < 0: no candidate)
    }
}

// Warp-aggregated atomicAdd(&cc[key], 1): lanes holding the same key add once.  key < 0: nothing.  All lanes of the warp call it.
__device__ __forceinline__ void qt_count(int* cc, int key) {
    const unsigned peers = __match_any_sync(ORBX_FULL_MASK, key);
    if (key >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&cc[key], __popc(peers));
}

template <int NT>
__global__ void __launch_bounds__(NT)
k_octree(const __grid_constant__ OrbxPlan plan, const OrbxWs ws, const int level_base) {
    extern __shared__ __align__(16) uint8_t smem_qt[];
    __shared__ int s_scratch[64];
    __shared__ int s_nexp, s_cut, s_state;
    int flip = 0;

    // grid (frames, levels): CTAs are handed out level-major, the biggest trees (level 0) first -- the longest jobs start first
    const int level = level_base + blockIdx.y, frame = blockIdx.x;
    const OrbxLevel& L = plan.lv[level];
    const int tid = threadIdx.x;
    constexpr int nt = NT;
    const int NC = plan.qt_nc;
    const int N = L.N;

    // node storage: shared memory, or this tree's block of the global scratch when the plan's trees do not fit
    uint8_t* p = ws.qt_scratch ? ws.qt_scratch + ((size_t)frame * plan.nlevels + level) * plan.qt_bytes : smem_qt;
    int* ccA = reinterpret_cast<int*>(p); p += (size_t)NC * 16;      // 4 child counts per node of the current list (int4 reads)
    int* ccB = reinterpret_cast<int*>(p); p += (size_t)NC * 16;      // ... of the list under construction
    unsigned long long* skey = reinterpret_cast<unsigned long long*>(p); p += (size_t)plan.qt_sk * 8;
    short4* rectA = reinterpret_cast<short4*>(p); p += (size_t)NC * 8;   // x0, x1, y0, y1
    short4* rectB = reinterpret_cast<short4*>(p); p += (size_t)NC * 8;
    int* cntA = reinterpret_cast<int*>(p); p += (size_t)NC * 4;
    int* cntB = reinterpret_cast<int*>(p); p += (size_t)NC * 4;
    int* seq = reinterpret_cast<int*>(p); p += (size_t)NC * 4;       // processing sequence (list positions)
    int* tmp = reinterpret_cast<int*>(p); p += (size_t)NC * 4;       // children per sequence entry -> exclusive scan
    int* cbase = reinterpret_cast<int*>(p); p += (size_t)NC * 4;     // creation index of the first child, -1 if the node is not split this pass
    int* surv = reinterpret_cast<int*>(p);                           // survivor flag -> new list position

    const int ncand = min(ws.cand_count[frame * plan.nlevels + level], L.cand_cap);
    const uint2* cand = ws.cand + (long long)frame * ws.cand_stride + L.cand_off;
    uint16_t* keynode = ws.keynode + (long long)frame * ws.cand_stride + L.cand_off;
    const int CM = (1 << ORBX_COORD_BITS) - 1;

    // ---- initial nodes (:548-590): assign by float division, drop the empty ones ----
    for (int i = tid; i < L.nIni; i += nt) {
        rectA[i] = make_short4((short)(int)__fmul_rn(L.hX, (float)i), (short)(int)__fmul_rn(L.hX, (float)(i + 1)), 0, (short)L.span_y);
        cntA[i] = 0;
    }
    __syncthreads();
    for (int b0 = 0; b0 < ncand; b0 += ORBX_QT_MLP * nt) {
        const int base = b0 + tid;
        uint32_t v[ORBX_QT_MLP];
#pragma unroll
        for (int j = 0; j < ORBX_QT_MLP; ++j) v[j] = base + j * nt < ncand ? cand[base + j * nt].x : 0u;
#pragma unroll
        for (int j = 0; j < ORBX_QT_MLP; ++j) {
            const int k = base + j * nt;
            int r = -1;
            if (k < ncand) {
                r = (int)__fdiv_rn((float)(int)(v[j] & CM), L.hX);
                keynode[k] = (uint16_t)r;
            }
            qt_count(cntA, r);
        }
    }
    __syncthreads();
    for (int i = tid; i < L.nIni; i += nt) surv[i] = cntA[i] > 0;
    __syncthreads();
    int nlive = block_excl_scan(surv, L.nIni, s_scratch, &flip);
    for (int i = tid; i < L.nIni; i += nt)
        if (cntA[i] > 0) {
            rectB[surv[i]] = rectA[i];
            cntB[surv[i]] = cntA[i];
        }
    for (int i = tid; i < 4 * nlive; i += nt) ccB[i] = 0;
    __syncthreads();
    // keys move to the compacted initial list; child occupancy of every node that can still be split
    qt_sweep<NT>(cand, keynode, ncand, [&](int k, int pos, uint32_t v) {
        int key = -1;
        if (pos >= 0) {
            const int np = surv[pos];
            keynode[k] = (uint16_t)np;
            if (cntB[np] > 1) key = 4 * np + qt_quadrant(rectB[np], (int)(v & CM), (int)((v >> ORBX_COORD_BITS) & CM));
        }
        qt_count(ccB, key);
    });
    __syncthreads();
    // current list: (rect, cnt, cc); list under construction: (nrect, ncnt, ncc)
    short4 *rect = rectB, *nrect = rectA;
    int *cnt = cntB, *ncnt = cntA, *cc = ccB, *ncc = ccA;

    bool careful = false;
    for (;;) {
        const int prev_size = nlive;
        // ---- processing sequence: live nodes holding > 1 key, in list order ----
        for (int i = tid; i < nlive; i += nt) tmp[i] = cnt[i] > 1;
        if (tid == 0) { s_nexp = 0; s_cut = -1; }
        __syncthreads();
        const int ns = block_excl_scan(tmp, nlive, s_scratch, &flip);
        if (ns == 0) break;  // every node is a single key: list size unchanged -> finish (:674)
        for (int i = tid; i < nlive; i += nt) {
            if (cnt[i] > 1) seq[tmp[i]] = i;
            surv[i] = i - tmp[i];        // position among the nodes that stay, valid whenever every node with > 1 key splits (no cut)
            cbase[i] = -1;
        }
        __syncthreads();
        if (careful && NT >= 512 && ns <= 512) {
            // (size desc, later-created first) == (size desc, list position asc).  Latency instances (512 / 1024 threads, one or a few
            // frames) with a few hundred nodes: a rank sort -- every thread owns at most one element and walks the keys once, two
            // barriers -- instead of the ~40 barriers of a bitonic network (with 256-thread CTAs in big groups the network is faster)
            for (int s_ = tid; s_ < ns; s_ += nt) {
                const int pos = seq[s_];
                skey[s_] = ((unsigned long long)(unsigned)cnt[pos] << 32) | (unsigned)(0x7fffffff - pos);
            }
            __syncthreads();
            for (int s_ = tid; s_ < ns; s_ += nt) {
                const unsigned long long key = skey[s_];
                int rank = 0;
                for (int t = 0; t < ns; ++t) rank += skey[t] > key;
                tmp[rank] = 0x7fffffff - (int)(unsigned)(key & 0xffffffffu);
            }
            __syncthreads();
            for (int s_ = tid; s_ < ns; s_ += nt) seq[s_] = tmp[s_];
            __syncthreads();
        } else if (careful) {
            // bitonic sort of the 64-bit keys, zero-padded to a power of two
            int n2 = 1;
            while (n2 < ns) n2 <<= 1;
            for (int s_ = tid; s_ < n2; s_ += nt) {
                unsigned long long key = 0ull;
                if (s_ < ns) {
                    const int pos = seq[s_];
                    key = ((unsigned long long)(unsigned)cnt[pos] << 32) | (unsigned)(0x7fffffff - pos);
                }
                skey[s_] = key;
            }
            __syncthreads();
            for (int size = 2; size <= n2; size <<= 1)
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int t = tid; t < (n2 >> 1); t += nt) {
                        const int i = 2 * t - (t & (stride - 1));          // lower index of the pair
                        const int j = i + stride;
                        const unsigned long long a = skey[i], b = skey[j];
                        const bool desc = (i & size) == 0;                 // descending blocks first: the whole array ends descending
                        if ((a < b) == desc) { skey[i] = b; skey[j] = a; }
                    }
                    __syncthreads();
                }
            for (int s_ = tid; s_ < ns; s_ += nt) seq[s_] = 0x7fffffff - (int)(unsigned)(skey[s_] & 0xffffffffu);
            __syncthreads();
        }
        // ---- children of every node in the sequence (occupancy was counted when the node was created) ----
        for (int s_ = tid; s_ < ns; s_ += nt) {
            const int4 c4 = *reinterpret_cast<const int4*>(cc + 4 * seq[s_]);
            tmp[s_] = (c4.x > 0) + (c4.y > 0) + (c4.z > 0) + (c4.w > 0);
        }
        if (tid == 0) tmp[ns] = 0;
        __syncthreads();
        block_excl_scan(tmp, ns + 1, s_scratch, &flip);  // tmp[s] = children created before split s
        // ---- how many splits are applied (careful phase stops once the list reaches N, :735) ----
        int nsplit = ns;
        if (careful) {
            for (int j = tid + 1; j <= ns; j += nt) {
                const bool now = prev_size + tmp[j] - j >= N;
                const bool before = prev_size + tmp[j - 1] - (j - 1) >= N;
                if (now && !before) s_cut = j;
            }
            __syncthreads();
            if (s_cut > 0) nsplit = s_cut;
        }
        const int T = tmp[nsplit];  // children created this pass
        int nsurv;
        if (nsplit == ns) {
            // every node with more than one key splits: the survivors' positions were derived from the first scan
            for (int s_ = tid; s_ < ns; s_ += nt) cbase[seq[s_]] = tmp[s_];
            nsurv = nlive - ns;
            __syncthreads();
        } else {
            // the careful phase stopped early: nodes behind the cut stay as they are
            for (int i = tid; i < nlive; i += nt) surv[i] = 1;
            __syncthreads();
            for (int s_ = tid; s_ < nsplit; s_ += nt) {
                const int pos = seq[s_];
                cbase[pos] = tmp[s_];
                surv[pos] = 0;
            }
            __syncthreads();
            nsurv = block_excl_scan(surv, nlive, s_scratch, &flip);
        }
        // children: list position T-1-(creation index); DivideNode geometry :488-514
        for (int s_ = tid; s_ < nsplit; s_ += nt) {
            const int pos = seq[s_];
            const short4 r = rect[pos];
            const int mx = r.x + ((r.y - r.x + 1) >> 1), my = r.z + ((r.w - r.z + 1) >> 1);
            const int4 c4 = *reinterpret_cast<const int4*>(cc + 4 * pos);
            const int nn[4] = {c4.x, c4.y, c4.z, c4.w};
            int ci = tmp[s_];
            int nexp = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int n = nn[c];
                if (n > 0) {
                    const int np = T - 1 - ci;
                    nrect[np] = make_short4((short)((c & 1) ? mx : r.x), (short)((c & 1) ? r.y : mx),
                                            (short)((c & 2) ? my : r.z), (short)((c & 2) ? r.w : my));
                    ncnt[np] = n;
                    *reinterpret_cast<int4*>(ncc + 4 * np) = make_int4(0, 0, 0, 0);
                    nexp += n > 1;
                    ++ci;
                }
            }
            if (nexp) atomicAdd(&s_nexp, nexp);
        }
        // untouched nodes keep their rectangle, count and child occupancy
        for (int i = tid; i < nlive; i += nt)
            if (cbase[i] < 0) {
                const int np = T + surv[i];
                nrect[np] = rect[i];
                ncnt[np] = cnt[i];
                *reinterpret_cast<int4*>(ncc + 4 * np) = *reinterpret_cast<const int4*>(cc + 4 * i);
            }
        __syncthreads();
        // keys follow their node, and count the occupancy of the new node's own children on the way
        qt_sweep<NT>(cand, keynode, ncand, [&](int k, int pos, uint32_t v) {
            int key = -1;
            if (pos >= 0) {
                const int cb = cbase[pos];
                int np;
                if (cb >= 0) {
                    const int x = (int)(v & CM), y = (int)((v >> ORBX_COORD_BITS) & CM);
                    const int c = qt_quadrant(rect[pos], x, y);
                    const int4 cc4 = *reinterpret_cast<const int4*>(cc + 4 * pos);
                    const unsigned occ = (cc4.x > 0) | ((cc4.y > 0) << 1) | ((cc4.z > 0) << 2) | ((cc4.w > 0) << 3);
                    np = T - 1 - (cb + __popc(occ & ((1u << c) - 1u)));
                    if (ncnt[np] > 1) key = 4 * np + qt_quadrant(nrect[np], x, y);
                } else {
                    np = T + surv[pos];
                }
                keynode[k] = (uint16_t)np;
            }
            qt_count(ncc, key);
        });
        __syncthreads();
        { short4* t = rect; rect = nrect; nrect = t; }
        { int* t = cnt; cnt = ncnt; ncnt = t; }
        { int* t = cc; cc = ncc; ncc = t; }
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
    unsigned long long* best = reinterpret_cast<unsigned long long*>(ncc);     // the list under construction is free now
    __syncthreads();
    for (int i = tid; i < nlive; i += nt) best[i] = 0ull;
    __syncthreads();
    for (int base = tid; base < ncand; base += ORBX_QT_MLP * nt) {
        uint2 v[ORBX_QT_MLP];
        int pos[ORBX_QT_MLP];
#pragma unroll
        for (int j = 0; j < ORBX_QT_MLP; ++j) {
            const int k = base + j * nt;
            v[j] = k < ncand ? cand[k] : make_uint2(0u, 0u);
            pos[j] = k < ncand ? (int)keynode[k] : -1;
        }
#pragma unroll
        for (int j = 0; j < ORBX_QT_MLP; ++j)
            if (pos[j] >= 0) {
                const unsigned long long key = ((unsigned long long)(v[j].x >> 24) << 56) | ((unsigned long long)(0xffffffffu - v[j].y) << 24) |
                                               (unsigned)(base + j * nt);
                atomicMax(&best[pos[j]], key);
            }
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
        tmp[i] = (xs >= (float)plan.lap0 && xs <= (float)plan.lap1) ? 1 : 0;
        OrbxKpRec o;
        o.x = x; o.y = y; o.response = (float)(v >> 24); o.angle = -1.f; o.lap_before = 0; o.src = k;
        rec[i] = o;
    }
    __syncthreads();
    const int nlap = block_excl_scan(tmp, nout, s_scratch, &flip);
    for (int i = tid; i < nout; i += nt) rec[i].lap_before = tmp[i];
    if (tid == 0) ws.level_count[frame * plan.nlevels + level] = make_int2(nout, nlap);
}

// =================================================================================================
// K5a  GaussianBlur 7x7 sigma 2, OpenCV fixed-point path: 8.8 kernel {18,34,48,56,48,34,18},
// exact accumulation, one rounding (V + 32768) >> 16.  Reads the bordered plane (its reflect-101
// border is exactly the BORDER_REFLECT_101 extension of the borderless clone the reference blurs).
// =================================================================================================
#define ORBX_BLUR_TW 128
#define ORBX_BLUR_TH 64

// Horizontal taps via IDP.4A on byte-aligned windows (funnel shifts of the staged words), vertical taps via
// IDP.2A on 16-bit horizontal sums stored as row pairs (rows 2k, 2k+1 share one 32-bit word per column).
// blockIdx.x indexes a host-built tile table (level | tile_x << 8 | tile_y << 20).
// TMA: the 160 x 70 byte source window arrives as one bulk tensor copy (ws.tmaps_b7, the level's plane) instead of 700 cp.async.
template <bool TMA>
__global__ void __launch_bounds__(256, 7)
k_blur7(const __grid_constant__ OrbxPlan plan, const OrbxWs ws, int skip_empty) {
    constexpr int SROWS = ORBX_BLUR_TH + 6;          // 70 staged rows = 35 row pairs
    constexpr int SPB = ORBX_BLUR_TW + 32;           // staged bytes per row: columns x0-16 .. x0+143 (16-byte chunks)
    constexpr int SW = SPB / 4;
    __shared__ __align__(128) uint32_t s_src[SROWS * SW];
    __shared__ __align__(16) uint32_t s_h2[(SROWS / 2) * ORBX_BLUR_TW];
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t tdesc = __ldg(ws.blur_tiles + blockIdx.x);
    const int level = tdesc & 0xff;
    const OrbxLevel& L = plan.lv[level];
    const int frame = blockIdx.y;
    // the reference skips levels without keypoints (:1122); when blur runs concurrently with the quadtree (graph
    // replay of small groups) every level is blurred instead -- the extra output is never read
    if (skip_empty && ws.level_count[frame * plan.nlevels + level].x == 0) return;
    const int x0 = ((tdesc >> 8) & 0xfff) * ORBX_BLUR_TW, y0 = (tdesc >> 20) * ORBX_BLUR_TH;
    const uint8_t* plane = ws.pyr + (long long)frame * ws.pyr_stride + L.plane_off;
    const int tid = threadIdx.x;
    if (TMA) {
        // the box holds plane rows 19 + y0 - 3 ..: halo rows above / below the level are the (possibly unwritten) border, so the
        // tiles on the top / bottom edge mirror their three halo rows in shared memory, like the halo columns below
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
        if (tid == 0) {
            orbx_mbar_init(bar, 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            orbx_mbar_expect_tx(bar, SROWS * SPB);
            orbx_tma_load_3d((uint32_t)__cvta_generic_to_shared(s_src), ws.tmaps_b7 + 128 * level, bar, ORBX_PADL + x0 - 16, ORBX_EDGE + y0 - 3, frame);
        }
        __syncthreads();                                   // the barrier is initialised before anybody polls it
        unsigned spins = 0;
        while (!orbx_mbar_try_wait(bar, 0)) {
            if (++spins > (1u << 18)) __trap();
        }
        const bool top = y0 == 0, bottom = L.h <= y0 + ORBX_BLUR_TH + 2;
        if (top || bottom) {
            // staged row r holds level row y0 - 3 + r
            for (int i = tid; i < 6 * SW; i += 256) {
                const int k = i / SW, wd = i - k * SW;     // k = 0..2: rows above the level, 3..5: rows below it
                if (k < 3) {
                    if (top) s_src[(2 - k) * SW + wd] = s_src[(4 + k) * SW + wd];          // level row -(k+1) <- row k+1
                } else if (bottom) {
                    const int e = L.h - y0 + 3 + (k - 3);  // staged index of level row h + (k-3)
                    if (e < SROWS) s_src[e * SW + wd] = s_src[(e - 2 - 2 * (k - 3)) * SW + wd];   // level row h+j <- row h-2-j
                }
            }
        }
    } else {
        // BORDER_REFLECT_101 of the borderless clone the reference blurs (:1126-1127), without reading the plane's border (it may
        // not have been written, see k_pyr_border): halo rows come from the mirrored level row, the 3 halo columns of the tiles
        // on the left / right edge are mirrored in shared memory below.
        const int max_chunk = (L.pitch >> 4) - 1, c0 = (ORBX_PADL + x0 - 16) >> 4;
        for (int i = tid; i < SROWS * (SPB / 16); i += 256) {
            const int r = i / (SPB / 16), c = i - r * (SPB / 16);
            const int lr = reflect101(min(y0 + r - 3, L.h + 2), L.h);
            __pipeline_memcpy_async(reinterpret_cast<uint8_t*>(s_src) + r * SPB + 16 * c,
                                    plane + (long long)(ORBX_EDGE + lr) * L.pitch + 16 * min(c0 + c, max_chunk), 16);
        }
        __pipeline_commit();
        __pipeline_wait_prior(0);
    }
    __syncthreads();
    {
        // staged byte 16 + (x - x0) of a row holds level column x
        // right: the level ends inside this tile or inside its 3-column halo (a level of 257 columns: the tile [128, 256) reads 256..258)
        const bool left = x0 == 0, right = L.w <= x0 + ORBX_BLUR_TW + 2;
        if (left || right) {
            uint8_t* sb = reinterpret_cast<uint8_t*>(s_src);
            if (tid < 2 * SROWS) {
                const int r = tid >> 1, side = tid & 1;
                uint8_t* row = sb + r * SPB + 16;
                if (side == 0 && left) { row[-1] = row[1]; row[-2] = row[2]; row[-3] = row[3]; }
                if (side == 1 && right) {
                    const int e = L.w - x0;                              // first column past the level, 1 <= e <= 130
                    row[e] = row[e - 2]; row[e + 1] = row[e - 3]; row[e + 2] = row[e - 4];
                }
            }
            __syncthreads();
        }
    }
    // ---- horizontal: item = (row pair, 4-pixel group); output x reads staged bytes x+13 .. x+19 ----
    const unsigned K0 = 18u | (34u << 8) | (48u << 16) | (56u << 24), K1 = 48u | (34u << 8) | (18u << 16);
    // partial tiles (right / bottom edge of a level): horizontal sums nobody reads are not computed
    const int xq_end = min(ORBX_BLUR_TW / 4, (L.blur_pitch - x0 + 3) >> 2);          // 4-pixel groups with an output column
    const int rp_end = min(SROWS / 2, ((min(ORBX_BLUR_TH, L.h - y0) + 1) >> 1) + 3);   // row pairs feeding an output row
    auto hitem = [&](int rp, int xq) {
        uint32_t h[2][4];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t* s = s_src + (2 * rp + k) * SW + xq + 3;
            const uint32_t a = s[0], b = s[1], c = s[2];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t lo = j == 3 ? b : __funnelshift_r(a, b, 8 * (j + 1));
                const uint32_t hi = j == 3 ? c : __funnelshift_r(b, c, 8 * (j + 1));
                h[k][j] = __dp4a(lo, K0, __dp4a(hi, K1, 0u));
            }
        }
        uint4 o;
        o.x = __byte_perm(h[0][0], h[1][0], 0x5410); o.y = __byte_perm(h[0][1], h[1][1], 0x5410);      // sums < 2^16: rows 2k | 2k+1 << 16
        o.z = __byte_perm(h[0][2], h[1][2], 0x5410); o.w = __byte_perm(h[0][3], h[1][3], 0x5410);
        *reinterpret_cast<uint4*>(s_h2 + rp * ORBX_BLUR_TW + 4 * xq) = o;
    };
    // tiles on a level's right edge hold xq_end < 32 groups per row: items are dealt out linearly there, so that no lane idles
    // (a level of 257 columns has a tile column with ONE group per row)
    const unsigned xq_magic = (1u << 20) / (unsigned)xq_end + 1u;      // i / xq_end for i < 2^11
    if (xq_end == ORBX_BLUR_TW / 4) {
        for (int i = tid; i < (SROWS / 2) * (ORBX_BLUR_TW / 4); i += 256) {
            if ((i >> 5) >= rp_end) break;
            hitem(i >> 5, i & 31);
        }
    } else {
        for (int i = tid; i < rp_end * xq_end; i += 256) {
            const int rp = (int)(((unsigned)i * xq_magic) >> 20);
            hitem(rp, i - rp * xq_end);
        }
    }
    __syncthreads();
    // ---- vertical: item = (two output rows 2q, 2q+1) x (4 pixels); both rows use row pairs q .. q+3 ----
    const unsigned WE01 = 18u | (34u << 8) | (48u << 16) | (56u << 24), WE23 = 48u | (34u << 8) | (18u << 16);
    const unsigned WO01 = (18u << 8) | (34u << 16) | (48u << 24), WO23 = 56u | (48u << 8) | (34u << 16) | (18u << 24);
    uint8_t* out = ws.blur + (long long)frame * ws.blur_stride + L.blur_off;
    auto vitem = [&](int q, int xq) {
        const int y = y0 + 2 * q, x = x0 + 4 * xq;
        const uint4 p0 = *reinterpret_cast<const uint4*>(s_h2 + (q + 0) * ORBX_BLUR_TW + 4 * xq);
        const uint4 p1 = *reinterpret_cast<const uint4*>(s_h2 + (q + 1) * ORBX_BLUR_TW + 4 * xq);
        const uint4 p2 = *reinterpret_cast<const uint4*>(s_h2 + (q + 2) * ORBX_BLUR_TW + 4 * xq);
        const uint4 p3 = *reinterpret_cast<const uint4*>(s_h2 + (q + 3) * ORBX_BLUR_TW + 4 * xq);
        const uint32_t a0[4] = {p0.x, p0.y, p0.z, p0.w}, a1[4] = {p1.x, p1.y, p1.z, p1.w};
        const uint32_t a2[4] = {p2.x, p2.y, p2.z, p2.w}, a3[4] = {p3.x, p3.y, p3.z, p3.w};
        uint32_t ve[4], vo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ve[j] = __dp2a_hi(a3[j], WE23, __dp2a_lo(a2[j], WE23, __dp2a_hi(a1[j], WE01, __dp2a_lo(a0[j], WE01, 32768u))));
            vo[j] = __dp2a_hi(a3[j], WO23, __dp2a_lo(a2[j], WO23, __dp2a_hi(a1[j], WO01, __dp2a_lo(a0[j], WO01, 32768u))));
        }
        // V + 32768 < 2^24: the result byte is byte 2 of each sum; three PRMTs gather four of them
        const uint32_t we = __byte_perm(__byte_perm(ve[0], ve[1], 0x0062), __byte_perm(ve[2], ve[3], 0x0062), 0x5410);
        const uint32_t wo = __byte_perm(__byte_perm(vo[0], vo[1], 0x0062), __byte_perm(vo[2], vo[3], 0x0062), 0x5410);
        *reinterpret_cast<uint32_t*>(out + (long long)y * L.blur_pitch + x) = we;
        if (y + 1 < L.h) *reinterpret_cast<uint32_t*>(out + (long long)(y + 1) * L.blur_pitch + x) = wo;
    };
    const int nq = (min(ORBX_BLUR_TH, L.h - y0) + 1) >> 1;             // output row pairs of this tile
    if (xq_end == ORBX_BLUR_TW / 4) {
#pragma unroll
        for (int it = 0; it < (ORBX_BLUR_TH / 2) * (ORBX_BLUR_TW / 4) / 256; ++it) {
            const int i = tid + it * 256;
            if ((i >> 5) < nq) vitem(i >> 5, i & 31);
        }
    } else {
        for (int i = tid; i < nq * xq_end; i += 256) {
            const int q = (int)(((unsigned)i * xq_magic) >> 20);
            vitem(q, i - q * xq_end);
        }
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

// Blackwell's packed FP32 multiply (FMUL2: two IEEE round-to-nearest products per instruction), spelled as PTX with the explicit .rn
// qualifier.  Only the products are packed: a packed mul feeding a packed add is fused by ptxas into FFMA2 -- one rounding instead of
// two, which would change descriptor bits -- even with .rn on both and -fmad=false (the __fmul2_rn / __fadd2_rn intrinsics of
// sm_100_rt.h behave the same); a scalar add of the two halves is left alone.
__device__ __forceinline__ float2 orbx_mul2(float2 a, float2 b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float2 orbx_add2(float2 a, float2 b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&d);
}

// Shared-memory byte load from a 32-bit shared address.  Not volatile: the scheduler may move it, its address operand orders it.
__device__ __forceinline__ unsigned orbx_lds_u8(unsigned saddr) {
    unsigned v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

#define ORBX_DESC_WARPS 8
#define ORBX_ANGLE_WORDS 9   // 31 patch columns + up to 3 bytes of alignment slack = 9 aligned words per row
#define ORBX_ANGLE_ROWS 33   // 31 patch rows + 2 all-zero rows: eleven steps of three rows

// glibc's sinf / cosf (sysdeps/ieee754/flt-32/s_sincosf.h, the ARM optimized-routines algorithm) for 0 <= y < 120:
// quadrant n = round(y * 2/pi) through a scaled float->int conversion, x = y - n * pi/2 in double, then a degree-7 sine or
// degree-8 cosine polynomial in double, rounded once to float (inputs below 2^-12 return y and 1).  Non-negative inputs only.
__device__ __forceinline__ float orbx_sinf_poly(double x, double x2, int n, bool neg_table) {
    // cosine c0..c4 and sine s1..s3 coefficients of __sincosf_table[0]; table[1] negates the cosine ones
    const double sg = neg_table ? -1.0 : 1.0;
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double s1 = __fma_rn(x2, -0x1.994eb3774cf24p-13, 0x1.1107605230bc4p-7);
        const double x7 = x3 * x2;
        const double s = __fma_rn(x3, -0x1.555545995a603p-3, x);
        return (float)__fma_rn(x7, s1, s);
    }
    const double x4 = x2 * x2;
    const double c2 = __fma_rn(x2, sg * 0x1.99343027bf8c3p-16, sg * -0x1.6c087e89a359dp-10);
    const double c1 = __fma_rn(x2, sg * -0x1.ffffffd0c621cp-2, sg * 0x1p0);
    const double x6 = x4 * x2;
    const double c = __fma_rn(x4, sg * 0x1.55553e1068f19p-5, c1);
    return (float)__fma_rn(x6, c2, c);
}

__device__ __forceinline__ void orbx_glibc_sincosf(float y, float* sinp, float* cosp) {
    const double x = (double)y;
    const unsigned top = __float_as_uint(y) >> 20;                    // abstop12 of a non-negative float
    if (top < (0x3f490fdbu >> 20)) {                                  // abstop12(y) < abstop12(pi/4 as float, 0x1.921FB6p-1)
        if (top < (0x39800000u >> 20)) { *sinp = y; *cosp = 1.0f; return; }   // |y| < 2^-12
        const double x2 = x * x;
        *sinp = orbx_sinf_poly(x, x2, 0, false);
        *cosp = orbx_sinf_poly(x, x2, 1, false);
        return;
    }
    const double r = x * 0x1.45F306DC9C883p+23;                      // 2/pi * 2^24
    const int n = ((int)r + 0x800000) >> 24;
    const double xr = __fma_rn(-(double)n, 0x1.921FB54442D18p0, x);
    const double x2 = xr * xr;
    // sign[q] = {1, -1, -1, 1}[q & 3]; sine uses quadrant n, cosine quadrant n + 1
    const double ss = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const double cs = (((n + 1) & 3) == 1 || ((n + 1) & 3) == 2) ? -1.0 : 1.0;
    *sinp = orbx_sinf_poly(xr * ss, x2, n, (n & 2) != 0);
    *cosp = orbx_sinf_poly(xr * cs, x2, n ^ 1, ((n + 1) & 2) != 0);
}

__device__ __forceinline__ int dp4a_u8_s8(unsigned a, int b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// cvRound for |v| < 2^22 without F2I: adding 1.5*2^23 rounds to nearest-even at integer granularity; the
// integer is the low mantissa bits (bias removed by the caller).
#define ORBX_RND_MAGIC 12582912.0f
#define ORBX_RND_BIAS 0x4B400000

// TMA: the 37-row window of the blurred level arrives as one bulk tensor copy per keypoint (box 64 B x 37 rows from the level's
// tensor map over the blur buffer, one mbarrier per keypoint) instead of 13 rounds of 4-byte cp.async with their index arithmetic.
// The box starts at a 16-byte aligned column (a box starting at an arbitrary byte never completed on the B200: trap), so it is
// 64 bytes wide: up to 15 bytes of alignment slack + the 37 columns.
//
// A warp owns TWO keypoints (slots 2w, 2w+1): lanes 0..15 carry the first through everything that is per keypoint (level
// bookkeeping, fastAtan2, sincos, the output records), lanes 16..31 the second, so those ~330 instructions are paid once per pair;
// the moments of both run on 27 lanes each, one after the other, and meet in a reduction that costs what one did; every lane then
// computes 16 tests = two descriptor bytes of its keypoint, and the two half-warps read the same pattern words (half the
// pattern traffic per keypoint -- the kernel is bound by L1 / shared-memory wavefronts, not by issue slots).
#define ORBX_DESC_PP 64        // patch row pitch in shared memory = TMA box width
#define ORBX_DESC_PATCH (37 * ORBX_DESC_PP + 64)   // 2432 B: 128-byte aligned slots
template <bool TMA>
__global__ void __launch_bounds__(ORBX_DESC_WARPS * 32, 5)
k_describe(const __grid_constant__ OrbxPlan plan, const OrbxWs ws, const OrbxFloatConsts fc,
           void* __restrict__ kps_out, uint8_t* __restrict__ desc_out,
           int cap_per_frame, int32_t* __restrict__ counts, int frame_out0) {
    __shared__ __align__(128) uint8_t s_patch[ORBX_DESC_WARPS][2][ORBX_DESC_PATCH];
    __shared__ __align__(8) unsigned long long s_bar[ORBX_DESC_WARPS][2];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int half = lane >> 4, hl = lane & 15;
    const int slot0 = 2 * (blockIdx.x * ORBX_DESC_WARPS + wib);
    const int frame = blockIdx.y;
    if (slot0 >= plan.kp_total) return;
    // ---- this half's keypoint; a half without one (odd total, or a level that kept fewer keypoints than it has slots) shadows the
    // other half -- same loads, no stores -- so that every warp-wide step below stays uniform ----
    const int2* lc = ws.level_count + frame * plan.nlevels;
    int2 mine = make_int2(0, 0);
    if (lane < plan.nlevels) mine = lc[lane];
    // per-frame sums of {keypoints, lapping keypoints} over all levels (REDUX); the frame's first warp reports them
    const int n_total = __reduce_add_sync(ORBX_FULL_MASK, mine.x), lap_total = __reduce_add_sync(ORBX_FULL_MASK, mine.y);
    const long long fo = (long long)(frame_out0 + frame);
    if (slot0 == 0 && lane == 0 && counts) {
        counts[2 * fo] = n_total;
        counts[2 * fo + 1] = n_total - lap_total;  // monoIndex, the reference's return value (:1161)
    }
    int slot = slot0 + half;
    int level = slot < plan.kp_total ? (int)__ldg(ws.slot_level + slot) : 0;
    const int n_level = __shfl_sync(ORBX_FULL_MASK, mine.x, level);
    const bool valid = slot < plan.kp_total && slot - plan.lv[level].kp_off < n_level;
    {
        const unsigned vb = __ballot_sync(ORBX_FULL_MASK, valid);
        if (vb == 0u) return;
        const int o_slot = __shfl_xor_sync(ORBX_FULL_MASK, slot, 16), o_level = __shfl_xor_sync(ORBX_FULL_MASK, level, 16);
        if (!valid) { slot = o_slot; level = o_level; }
    }
    const OrbxLevel& L = plan.lv[level];
    const int idx = slot - L.kp_off;
    OrbxKpRec* recp = ws.kprec + (long long)frame * ws.kp_stride + slot;
    const OrbxKpRec rec = *recp;
    // the sums over the levels below each half's level
    const int lvl_a = __shfl_sync(ORBX_FULL_MASK, level, 0), lvl_b = __shfl_sync(ORBX_FULL_MASK, level, 16);
    const int nb_a = __reduce_add_sync(ORBX_FULL_MASK, lane < lvl_a ? mine.x : 0), nb_b = __reduce_add_sync(ORBX_FULL_MASK, lane < lvl_b ? mine.x : 0);
    const int lb_a = __reduce_add_sync(ORBX_FULL_MASK, lane < lvl_a ? mine.y : 0), lb_b = __reduce_add_sync(ORBX_FULL_MASK, lane < lvl_b ? mine.y : 0);
    const int n_before = half ? nb_b : nb_a, lap_before = half ? lb_b : lb_a;
    const int cx = (int)rec.x, cy = (int)rec.y;  // integral by construction

    // Stage the 37x37 window of the blurred level (|rotated offset| <= 18) into shared memory, issued first so that the copy
    // overlaps the moment computation below: one TMA box per keypoint, or aligned 32-bit cp.async copies.
    const int bp = L.blur_pitch;
    const int bx0 = cx - 18, o0 = TMA ? (bx0 & 15) : (bx0 & 3);
    uint8_t* patch = s_patch[wib][half];
    const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar[wib][half]);
    if (TMA) {
        if (hl == 0) {
            orbx_mbar_init(bar, 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            orbx_mbar_expect_tx(bar, 37 * ORBX_DESC_PP);
            orbx_tma_load_3d((uint32_t)__cvta_generic_to_shared(patch), ws.tmaps_blur + 128 * level, bar, bx0 - o0, cy - 18, frame);
        }
    } else {
        const uint32_t* b32 = reinterpret_cast<const uint32_t*>(ws.blur + (long long)frame * ws.blur_stride + L.blur_off +
                                                                  (long long)(cy - 18) * bp + (bx0 - o0));
        const int bpw = bp >> 2;
        for (int i = hl; i < 37 * 11; i += 16) {
            const int row = (i * 373) >> 12;                   // i / 11 for i < 407
            const int wd = i - row * 11;
            __pipeline_memcpy_async(reinterpret_cast<uint32_t*>(patch) + row * (ORBX_DESC_PP / 4) + wd, b32 + row * bpw + wd, 4);
        }
        __pipeline_commit();
    }

    // ---- IC_Angle on the un-blurred level: 31 rows x 9 aligned words, IDP.4A against per-alignment weight words: u inside the
    // circle (else 0) for m10, v = row - 15 inside the circle (else 0) for m01.  Lanes 0..26 hold three rows of nine words and step
    // three rows at a time: every address is the lane's first one plus a constant.  The table has 33 rows (31, 32: all zero), so
    // the last step needs no row test; the pixels it multiplies by zero are plane rows cy+16, cy+17 (inside the bordered plane).
    // Both keypoints of the warp go through the same 27 lanes: their patch corner, pitch and alignment come from lanes 0 / 16.
    int m10 = 0, m01 = 0;
    {
        const int col0 = ORBX_PADL + cx - ORBX_HALF_PATCH;
        const int al_h = col0 & 3;
        const unsigned long long corner = (unsigned long long)(ws.pyr + (long long)frame * ws.pyr_stride + L.plane_off +
                                                               (long long)(ORBX_EDGE + cy - ORBX_HALF_PATCH) * L.pitch + (col0 - al_h));
        const int r0 = (lane * 57) >> 9;                       // lane / 9
        const int wd4 = 4 * (lane - ORBX_ANGLE_WORDS * r0);
        int sum10[2], sum01[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const unsigned long long c_k = __shfl_sync(ORBX_FULL_MASK, corner, 16 * k);
            const int pitch_k = __shfl_sync(ORBX_FULL_MASK, L.pitch, 16 * k), al_k = __shfl_sync(ORBX_FULL_MASK, al_h, 16 * k);
            int a10 = 0, a01 = 0;
            if (lane < 3 * ORBX_ANGLE_WORDS) {
                const uint8_t* p8 = reinterpret_cast<const uint8_t*>(c_k) + (long long)r0 * pitch_k + wd4;
                const long long step = 3LL * pitch_k;
                const int2* wt = ws.angle_w + al_k * (ORBX_ANGLE_ROWS * ORBX_ANGLE_WORDS) + lane;
#pragma unroll
                for (int it = 0; it < ORBX_ANGLE_ROWS / 3; ++it) {
                    const unsigned px = __ldg(reinterpret_cast<const uint32_t*>(p8));
                    p8 += step;
                    const int2 w = __ldg(wt + it * 3 * ORBX_ANGLE_WORDS);
                    a10 = dp4a_u8_s8(px, w.x, a10);
                    a01 = dp4a_u8_s8(px, w.y, a01);
                }
            }
            sum10[k] = a10; sum01[k] = a01;
        }
        // each half keeps the partial sums of ITS keypoint and receives the partner lane's: five shuffle steps reduce both keypoints
        m10 = (half ? sum10[1] : sum10[0]) + __shfl_xor_sync(ORBX_FULL_MASK, half ? sum10[0] : sum10[1], 16);
        m01 = (half ? sum01[1] : sum01[0]) + __shfl_xor_sync(ORBX_FULL_MASK, half ? sum01[0] : sum01[1], 16);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            m10 += __shfl_xor_sync(ORBX_FULL_MASK, m10, o);
            m01 += __shfl_xor_sync(ORBX_FULL_MASK, m01, o);
        }
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10, fc);

    // ---- rBRIEF on the blurred level: lane hl of a half computes descriptor bytes 2 hl, 2 hl + 1 of its keypoint ----
    const float rad = __fmul_rn(angle, fc.deg2rad);
    // The reference calls glibc's cosf / sinf (:111).  orbx_glibc_sincosf restates that algorithm (a double-precision
    // polynomial after a quadrant reduction); it equals glibc 2.39 on every float in [0, 6.5] (checked exhaustively on the
    // host, tools/check_sincosf.c), so the sample coordinates -- and the descriptors -- are bit-identical, not just close.
    float a, b;
    orbx_glibc_sincosf(rad, &b, &a);
    if (TMA) {
        __syncwarp();                                           // lanes 0 / 16 have initialised the barriers and issued the copies
        unsigned spins = 0;
        while (!orbx_mbar_try_wait(bar, 0)) {
            if (++spins > (1u << 18)) __trap();                 // a copy that never lands must fail loudly, not hang the device
        }
    } else {
        __pipeline_wait_prior(0);
    }
    __syncwarp();
    // Sample address = patch + (round(r)+18)*64 + round(c)+18+o0.  The constant parts ride in the rounding constants: adding
    // 1.5*2^23 + k (k an EVEN integer: ties still go to the even neighbour, like cvRound) leaves BIAS + k + round(v) in the low bits,
    // so the row term carries the +18, the column term the patch's shared address + 18 + o0 less its odd bit, and one register
    // holds the rest (the two biases, mod 2^32, and that odd bit): address = (rbits << 6) + cbits + kreg, two integer instructions.
    const unsigned patch_sa = (unsigned)__cvta_generic_to_shared(patch) + 18u + (unsigned)o0;
    const float mr = ORBX_RND_MAGIC + 18.0f, mc = (float)(12582912u + (patch_sa & ~1u));      // exact: integers below 2^24
    unsigned kreg = (patch_sa & 1u) - 65u * (unsigned)ORBX_RND_BIAS;
    asm volatile("" : "+r"(kreg) :: "memory");              // the loads below (plain asm, free to be scheduled) stay behind the wait above
    const float4* pat = reinterpret_cast<const float4*>(ws.pattern_f) + hl;   // layout [k][hl]: both halves read the same 256 bytes
    // Both points of a test share Blackwell's packed FP32 instructions: FMUL2 for the four products, scalar adds (see orbx_mul2),
    // FADD2 for the rounding constant -- the same bits as sixteen scalar operations in ten.  x*a - y*b is x*a + y*(-b): negation is exact.
    const float2 aa = make_float2(a, a), bb = make_float2(b, b), nb = make_float2(-b, -b), mmr = make_float2(mr, mr), mmc = make_float2(mc, mc);
    unsigned val = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const float4 t = __ldg(pat + k * 16);   // test 16*hl + k: x0, x1, y0, y1
        const float2 px = make_float2(t.x, t.y), py = make_float2(t.z, t.w);
        const float2 xb = orbx_mul2(px, bb), ya = orbx_mul2(py, aa), xa = orbx_mul2(px, aa), yb = orbx_mul2(py, nb);
        const float2 r = orbx_add2(make_float2(__fadd_rn(xb.x, ya.x), __fadd_rn(xb.y, ya.y)), mmr);
        const float2 c = orbx_add2(make_float2(__fadd_rn(xa.x, yb.x), __fadd_rn(xa.y, yb.y)), mmc);
        const unsigned t0 = orbx_lds_u8(((unsigned)__float_as_int(r.x) << 6) + (unsigned)__float_as_int(c.x) + kreg);
        const unsigned t1 = orbx_lds_u8(((unsigned)__float_as_int(r.y) << 6) + (unsigned)__float_as_int(c.y) + kreg);
        val |= (unsigned)(t0 < t1) << k;
    }

    // ---- output slot: lapping keypoints fill from the back, the rest from the front ----
    if (!valid) return;
    const float xs = level != 0 ? __fmul_rn(rec.x, L.sf) : rec.x;
    const float ys = level != 0 ? __fmul_rn(rec.y, L.sf) : rec.y;
    const bool lapping = xs >= (float)plan.lap0 && xs <= (float)plan.lap1;
    const int laps_before = lap_before + rec.lap_before;
    const int out_idx = lapping ? (n_total - 1 - laps_before) : (n_before + idx - laps_before);
    if (hl == 0) recp->angle = angle;
    if (out_idx < cap_per_frame) {
        const long long orow = fo * cap_per_frame + out_idx;
        if (desc_out) {
            uint8_t* d = desc_out + orow * 32 + 2 * hl;
            if (((uintptr_t)desc_out & 1) == 0) *reinterpret_cast<uint16_t*>(d) = (uint16_t)val;
            else { d[0] = (uint8_t)val; d[1] = (uint8_t)(val >> 8); }
        }
        if (kps_out && hl < 7) {
            float f = xs;                                      // cv::KeyPoint: pt.x, pt.y, size, angle, response, octave, class_id
            f = hl == 1 ? ys : f;
            f = hl == 2 ? L.kp_size : f;
            f = hl == 3 ? angle : f;
            f = hl == 4 ? rec.response : f;
            f = hl == 5 ? __int_as_float(level) : f;
            f = hl == 6 ? __int_as_float(-1) : f;
            reinterpret_cast<float*>(kps_out)[orow * 7 + hl] = f;
        }
    }
}

// =================================================================================================
// S1/S2  Frame::ComputeStereoMatches (reference src/Frame.cc:813-990), the first consumer of the path's outputs.
// k_stereo_match: one warp per left keypoint -- Hamming search over the right keypoints whose row band covers
// the left keypoint's row (DescriptorDistance, src/ORBmatcher.cc:2349), then the 11x11 SAD sliding window on the
// two resident pyramids and the parabola sub-pixel fit.  k_stereo_filter: median-based outlier rejection.
// =================================================================================================
struct OrbxStereoArgs {
    const OrbxKeyPoint* kl; const uint32_t* dl; int nl;
    const OrbxKeyPoint* kr; const uint32_t* dr; int nr;
    // batched form (blockIdx.y = stereo pair): per-pair strides of the keypoint / descriptor / output arrays (entries) and of the
    // resident pyramids (bytes); counts_* = the {n, mono} pairs orbx_extract_batch wrote on the device (NULL: nl / nr above)
    long long kp_stride; long long pyr_stride;
    const int32_t* counts_l; const int32_t* counts_r; int cap;
    const uint8_t* pyr_l; const uint8_t* pyr_r;    // frame base of each handle's resident pyramid
    float mbf, min_d, max_d;                        // minD = 0, maxD = mbf / mb (:844-846)
    float th_mul;                                   // 1.5f * 1.4f (:979)
    float* u_right; float* depth; int* sad;         // sad: window distance of an accepted match, else -1
    int* n_matched;
};

#define ORBX_TH_HIGH 100   // ORBmatcher::TH_HIGH, src/ORBmatcher.cc:36
#define ORBX_TH_LOW 50     // ORBmatcher::TH_LOW, :37

__global__ void __launch_bounds__(256)
k_stereo_match(const __grid_constant__ OrbxPlan plan, OrbxStereoArgs a) {
    const int lane = threadIdx.x & 31;
    const int iL = blockIdx.x * 8 + (threadIdx.x >> 5);
    {   // this block's stereo pair
        const long long pair = blockIdx.y;
        if (a.counts_l) { a.nl = min(max(a.counts_l[2 * pair], 0), a.cap); a.nr = min(max(a.counts_r[2 * pair], 0), a.cap); }
        a.kl += pair * a.kp_stride; a.dl += 8 * pair * a.kp_stride; a.kr += pair * a.kp_stride; a.dr += 8 * pair * a.kp_stride;
        a.pyr_l += pair * a.pyr_stride; a.pyr_r += pair * a.pyr_stride;
        a.u_right += pair * a.kp_stride; a.depth += pair * a.kp_stride; a.sad += pair * a.kp_stride;
    }
    if (iL >= a.nl) return;
    const OrbxKeyPoint kpL = a.kl[iL];
    float out_u = -1.0f, out_d = -1.0f;
    int out_sad = -1;
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;                                   // vRowIndices[vL], :858
    const float minU = __fsub_rn(uL, a.max_d), maxU = __fsub_rn(uL, a.min_d);
    if (maxU >= 0.f && levelL >= 0 && levelL < plan.nlevels) { // :866
        // ---- best right candidate: smallest (distance, index), distance < TH_HIGH (:869-897) ----
        const uint4 l0 = *reinterpret_cast<const uint4*>(a.dl + 8 * (size_t)iL), l1 = *reinterpret_cast<const uint4*>(a.dl + 8 * (size_t)iL + 4);
        unsigned best = ((unsigned)ORBX_TH_HIGH << 16) | 0xffffu;
        for (int iR = lane; iR < a.nr; iR += 32) {
            const OrbxKeyPoint kr = a.kr[iR];
            if (kr.octave < 0 || kr.octave >= plan.nlevels) continue;
            const float r = __fmul_rn(2.0f, plan.lv[kr.octave].sf);
            const int maxr = (int)ceilf(__fadd_rn(kr.y, r)), minr = (int)floorf(__fsub_rn(kr.y, r));   // :836-837
            if (row < minr || row > maxr) continue;
            if (kr.octave < levelL - 1 || kr.octave > levelL + 1) continue;
            if (!(kr.x >= minU && kr.x <= maxU)) continue;
            const uint4 r0 = *reinterpret_cast<const uint4*>(a.dr + 8 * (size_t)iR), r1 = *reinterpret_cast<const uint4*>(a.dr + 8 * (size_t)iR + 4);
            const int dist = __popc(l0.x ^ r0.x) + __popc(l0.y ^ r0.y) + __popc(l0.z ^ r0.z) + __popc(l0.w ^ r0.w) +
                             __popc(l1.x ^ r1.x) + __popc(l1.y ^ r1.y) + __popc(l1.z ^ r1.z) + __popc(l1.w ^ r1.w);
            best = min(best, ((unsigned)dist << 16) | (unsigned)iR);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(ORBX_FULL_MASK, best, o));
        const int bestDist = (int)(best >> 16);
        if (bestDist < (ORBX_TH_HIGH + ORBX_TH_LOW) / 2) {     // thOrbDist, :818, :900
            const int bestIdxR = (int)(best & 0xffffu);
            const OrbxLevel& L = plan.lv[levelL];
            const float uR0 = a.kr[bestIdxR].x;
            const float sfi = __fdiv_rn(1.0f, L.sf);           // mvInvScaleFactors[octave], :431-435 of the extractor
            const float scaleduL = roundf(__fmul_rn(uL, sfi)), scaledvL = roundf(__fmul_rn(vL, sfi));
            const float scaleduR0 = roundf(__fmul_rn(uR0, sfi));
            const int w = 5, Lh = 5;
            const float iniu = scaleduR0 + Lh - w, endu = scaleduR0 + Lh + w + 1;
            if (!(iniu < 0 || endu >= (float)L.w)) {           // :923-924
                const uint8_t* imL = a.pyr_l + L.plane_off + (long long)ORBX_EDGE * L.pitch + ORBX_PADL;
                const uint8_t* imR = a.pyr_r + L.plane_off + (long long)ORBX_EDGE * L.pitch + ORBX_PADL;
                const int y0 = (int)(scaledvL - w), x0 = (int)(scaleduL - w), xr0 = (int)(scaleduR0 - w);
                const int cL = imL[(y0 + w) * L.pitch + x0 + w];
                int il[4], pr_[4], pc_[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int p = lane + 32 * q;
                    const int pr = p < 121 ? p / 11 : 0, pc = p < 121 ? p - 11 * (p / 11) : 0;
                    pr_[q] = pr; pc_[q] = pc;
                    il[q] = (int)imL[(y0 + pr) * L.pitch + x0 + pc] - cL;      // IL - IL(w,w), :912-913
                }
                int bestW = 0x7fffffff, bestinc = 0;
                float d_prev = 0.f, d_best = 0.f, d_next = 0.f, d_last = 0.f;
                bool want_next = false;
#pragma unroll 1
                for (int inc = -Lh; inc <= Lh; ++inc) {                          // :926-943
                    const int xr = xr0 + inc;
                    const int cR = imR[(y0 + w) * L.pitch + xr + w];
                    int acc = 0;
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (lane + 32 * q < 121) acc += abs(il[q] - ((int)imR[(y0 + pr_[q]) * L.pitch + xr + pc_[q]] - cR));
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(ORBX_FULL_MASK, acc, o);
                    const float dist = (float)acc;
                    if (want_next) { d_next = dist; want_next = false; }
                    if (acc < bestW) { bestW = acc; bestinc = inc; d_prev = d_last; d_best = dist; want_next = true; }
                    d_last = dist;
                }
                if (bestinc != -Lh && bestinc != Lh) {                           // :945
                    const float deltaR = __fdiv_rn(__fsub_rn(d_prev, d_next),
                                                   __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d_prev, d_next), __fmul_rn(2.0f, d_best))));   // :953
                    if (!(deltaR < -1.f || deltaR > 1.f)) {
                        float bestuR = __fmul_rn(L.sf, __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));   // :959
                        float disparity = __fsub_rn(uL, bestuR);
                        if (disparity >= a.min_d && disparity < a.max_d) {       // :963
                            if (disparity <= 0.f) {
                                disparity = 0.01f;
                                bestuR = (float)((double)uL - 0.01);
                            }
                            out_d = __fdiv_rn(a.mbf, disparity);
                            out_u = bestuR;
                            out_sad = bestW;
                        }
                    }
                }
            }
        }
    }
    if (lane == 0) { a.u_right[iL] = out_u; a.depth[iL] = out_d; a.sad[iL] = out_sad; }
}

// Outlier rejection (:977-990): median of the accepted window distances (element size/2 of the ascending
// order), threshold 1.5 * 1.4 * median, matches at or above it are dropped.  One CTA.
__global__ void __launch_bounds__(256)
k_stereo_filter(OrbxStereoArgs a) {
    __shared__ int s_cnt[8];
    __shared__ int s_total;
    const int tid = threadIdx.x;
    {
        const long long pair = blockIdx.x;
        if (a.counts_l) a.nl = min(max(a.counts_l[2 * pair], 0), a.cap);
        a.u_right += pair * a.kp_stride; a.depth += pair * a.kp_stride; a.sad += pair * a.kp_stride; a.n_matched += pair;
    }
    auto block_sum = [&](int v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(ORBX_FULL_MASK, v, o);
        __syncthreads();
        if ((tid & 31) == 0) s_cnt[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) { int t = 0; for (int i = 0; i < 8; ++i) t += s_cnt[i]; s_total = t; }
        __syncthreads();
        return s_total;
    };
    int mine = 0;
    for (int i = tid; i < a.nl; i += 256) mine += a.sad[i] >= 0;
    const int n = block_sum(mine);
    if (n == 0) { if (tid == 0) *a.n_matched = 0; return; }   // the reference indexes an empty vector here (UB)
    const int k = n / 2;                                        // vDistIdx[size/2] of the ascending order
    int lo = 0, hi = 65535;                                     // smallest v with #(sad <= v) >= k + 1
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        int c = 0;
        for (int i = tid; i < a.nl; i += 256) { const int s = a.sad[i]; c += (s >= 0 && s <= mid); }
        if (block_sum(c) >= k + 1) hi = mid; else lo = mid + 1;
    }
    const float thDist = __fmul_rn(a.th_mul, (float)lo);
    int kept = 0;
    for (int i = tid; i < a.nl; i += 256) {
        const int s = a.sad[i];
        if (s < 0) continue;
        if ((float)s < thDist) { ++kept; } else { a.u_right[i] = -1.0f; a.depth[i] = -1.0f; }
    }
    kept = block_sum(kept);
    if (tid == 0) *a.n_matched = kept;
}

#endif  // ORBX_KERNELS_CUH_
