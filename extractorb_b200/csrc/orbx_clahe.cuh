// extractorb_b200/csrc/orbx_clahe.cuh -- sm_100a kernels for the pre-processing step of the reference's demos
// (SURVEY.md section 8(f), rank 4): cv::createCLAHE(3.0, Size(8, 8))->apply(image, im_clahe)
// (reference src/orb_extractor/main_orb_extractor.cpp:19-22, src/clahe/main_clahe.cpp:7-11,
// src/clahe/main_show_clahe_keypoint.cpp:19-22).  The arithmetic is OpenCV's (imgproc clahe.cpp, 8-bit path):
//
//   k_clahe_lut    one CTA per (tile, frame): 256-bin histogram of the tile (of the BORDER_REFLECT_101-extended image
//                  when the sides are not multiples of the tile grid), clip at clipLimit, redistribute the excess
//                  (equal share + one extra to every residualStep-th bin), prefix sum, lut = round(sum * lutScale)
//   k_clahe_apply  one CTA per (band of rows that interpolates between the same two tile rows, frame): the two rows of
//                  tile LUTs are staged in shared memory; every pixel blends its four LUT entries in float
//                  ((p1*xa1 + p2*xa)*ya1 + (q1*xa1 + q2*xa)*ya, every operation rounded separately, no FMA) and rounds
//                  half to even.
//
// A launch group is sized so that its frames stay in the 126 MB L2 between the two kernels: the image is read from HBM once.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

struct OrbxClaheArgs {
    const uint8_t* src; long long src_row, src_frame;   // bytes
    uint8_t* dst; long long dst_row, dst_frame;
    uint8_t* lut;                                        // [frame][tiles_y * tiles_x][256]
    int w, h, tiles_x, tiles_y, tw, th;                  // tw, th: tile size in the extended image
    int clip;                                            // clipLimit in pixels (0: no clipping)
    float lut_scale, inv_tw, inv_th;
};

__device__ __forceinline__ int clahe_reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

__global__ void __launch_bounds__(256)
k_clahe_lut(const OrbxClaheArgs a) {
    __shared__ int s_hist[8][256];   // one histogram per warp
    __shared__ int s_part[8];
    __shared__ int s_clipped;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.x, frame = blockIdx.y;
    const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
    for (int i = tid; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint8_t* img = a.src + (long long)frame * a.src_frame;
    const int x0 = tx * a.tw, y0 = ty * a.th;
    int* hist = s_hist[wid];
    const bool inside = x0 + a.tw <= a.w && y0 + a.th <= a.h;
    if (inside && ((a.tw | x0) & 3) == 0 && (a.src_row & 3) == 0 && (((uintptr_t)img) & 3) == 0) {
        // aligned words, 4 pixels per load
        const int wpr = a.tw >> 2, total = wpr * a.th;
        for (int i = tid; i < total; i += 256) {
            const int r = i / wpr, c = i - r * wpr;
            const uint32_t v = *reinterpret_cast<const uint32_t*>(img + (long long)(y0 + r) * a.src_row + x0 + 4 * c);
            atomicAdd(&hist[v & 0xffu], 1); atomicAdd(&hist[(v >> 8) & 0xffu], 1);
            atomicAdd(&hist[(v >> 16) & 0xffu], 1); atomicAdd(&hist[v >> 24], 1);
        }
    } else {
        const int total = a.tw * a.th;
        for (int i = tid; i < total; i += 256) {
            const int r = i / a.tw, c = i - r * a.tw;
            const int y = clahe_reflect101(y0 + r, a.h), x = clahe_reflect101(x0 + c, a.w);
            atomicAdd(&hist[img[(long long)y * a.src_row + x]], 1);
        }
    }
    __syncthreads();
    // bin `tid`: sum of the warp histograms, clip, redistribute
    int hv = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) hv += s_hist[k][tid];
    if (a.clip > 0) {
        int excess = max(hv - a.clip, 0);
        hv = min(hv, a.clip);
        excess = __reduce_add_sync(0xffffffffu, excess);
        if (lane == 0) s_part[wid] = excess;
        __syncthreads();
        if (tid == 0) {
            int c = 0;
            for (int k = 0; k < 8; ++k) c += s_part[k];
            s_clipped = c;
        }
        __syncthreads();
        const int clipped = s_clipped;
        const int batch = clipped / 256;
        const int residual = clipped - batch * 256;
        hv += batch;
        if (residual != 0) {
            const int step = max(256 / residual, 1);      // bins 0, step, 2*step, ... get one more, `residual` of them at most
            if (tid % step == 0 && tid / step < residual) hv += 1;
        }
        __syncthreads();
    }
    // inclusive prefix sum over the 256 bins
    int inc = hv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_part[wid] = inc;
    __syncthreads();
    int off = 0;
    for (int k = 0; k < wid; ++k) off += s_part[k];
    const int sum = inc + off;
    const int v = __float2int_rn(__fmul_rn((float)sum, a.lut_scale));           // saturate_cast<uchar>(sum * lutScale)
    a.lut[((long long)frame * (a.tiles_x * a.tiles_y) + tile) * 256 + tid] = (uint8_t)min(max(v, 0), 255);
}

// Rows y with the same (ty1, ty2) = clamp(floor(y / th - 0.5) + {0, 1}) form a band; band b covers the rows whose
// unclamped ty1 is b - 1: y in [ceil((b - 0.5) * th), ceil((b + 0.5) * th)).  The host passes nothing but the geometry;
// the band limits are found by evaluating the reference's float expression on the candidate rows.
__global__ void __launch_bounds__(256)
k_clahe_apply(const OrbxClaheArgs a) {
    extern __shared__ __align__(16) uint8_t s_lut[];     // [2][tiles_x][256]
    const int tid = threadIdx.x;
    const int band = blockIdx.x, frame = blockIdx.y;      // band = unclamped ty1 + 1, in [0, tiles_y]
    // first / last row of the band (float arithmetic identical to the per-pixel one below)
    int r0 = max(band * a.th - a.th / 2 - 2, 0), r1;
    while (r0 < a.h && (int)floorf(__fsub_rn(__fmul_rn((float)r0, a.inv_th), 0.5f)) < band - 1) ++r0;
    r1 = r0;
    while (r1 < a.h && (int)floorf(__fsub_rn(__fmul_rn((float)r1, a.inv_th), 0.5f)) == band - 1) ++r1;
    if (r0 >= r1) return;
    const int ty1 = max(band - 1, 0), ty2 = min(band, a.tiles_y - 1);
    {
        const uint8_t* lut = a.lut + (long long)frame * (a.tiles_x * a.tiles_y) * 256;
        const uint4* p1 = reinterpret_cast<const uint4*>(lut + (long long)ty1 * a.tiles_x * 256);
        const uint4* p2 = reinterpret_cast<const uint4*>(lut + (long long)ty2 * a.tiles_x * 256);
        const int n16 = a.tiles_x * 16;
        uint4* s = reinterpret_cast<uint4*>(s_lut);
        for (int i = tid; i < n16; i += 256) { s[i] = p1[i]; s[n16 + i] = p2[i]; }
    }
    // per-column interpolation terms, once per CTA: xa = frac(x / tw - 0.5) and the two LUT bases (tile column * 256)
    float* s_xa = reinterpret_cast<float*>(s_lut + 2 * a.tiles_x * 256);
    uint32_t* s_xi = reinterpret_cast<uint32_t*>(s_xa + ((a.w + 3) & ~3));
    for (int x = tid; x < a.w; x += 256) {
        const float txf = __fsub_rn(__fmul_rn((float)x, a.inv_tw), 0.5f);
        const int t1 = (int)floorf(txf);
        s_xa[x] = __fsub_rn(txf, (float)t1);
        s_xi[x] = (uint32_t)(max(t1, 0) * 256) | ((uint32_t)(min(t1 + 1, a.tiles_x - 1) * 256) << 16);
    }
    __syncthreads();
    const uint8_t* plane1 = s_lut;
    const uint8_t* plane2 = s_lut + a.tiles_x * 256;
    const uint8_t* img = a.src + (long long)frame * a.src_frame;
    uint8_t* out = a.dst + (long long)frame * a.dst_frame;
    const bool words = (a.w & 3) == 0 && ((a.src_row | a.dst_row) & 3) == 0 && ((((uintptr_t)img) | ((uintptr_t)out)) & 3) == 0;
    const int lane = tid & 31, wid = tid >> 5;
    // u8 -> float without the conversion unit: 2^23 + b is exact, subtracting 2^23 leaves b
    auto u8f = [](uint32_t b) { return __fsub_rn(__uint_as_float(0x4b000000u | b), 8388608.0f); };
    for (int y = r0 + wid; y < r1; y += 8) {              // one warp per row
        const float tyf = __fsub_rn(__fmul_rn((float)y, a.inv_th), 0.5f);
        const float ya = __fsub_rn(tyf, (float)(band - 1)), ya1 = __fsub_rn(1.0f, ya);
        const uint8_t* srow = img + (long long)y * a.src_row;
        uint8_t* drow = out + (long long)y * a.dst_row;
        if (words) {
            for (int c = lane; c < (a.w >> 2); c += 32) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(srow + 4 * c);
                const float4 xa4 = *reinterpret_cast<const float4*>(s_xa + 4 * c);
                const uint4 xi4 = *reinterpret_cast<const uint4*>(s_xi + 4 * c);
                const float xas[4] = {xa4.x, xa4.y, xa4.z, xa4.w};
                const uint32_t xis[4] = {xi4.x, xi4.y, xi4.z, xi4.w};
                uint32_t o = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t sv = (v >> (8 * k)) & 0xffu;
                    const uint32_t i1 = (xis[k] & 0xffffu) + sv, i2 = (xis[k] >> 16) + sv;
                    const float xa = xas[k], xa1 = __fsub_rn(1.0f, xa);
                    const float top = __fadd_rn(__fmul_rn(u8f(plane1[i1]), xa1), __fmul_rn(u8f(plane1[i2]), xa));
                    const float bot = __fadd_rn(__fmul_rn(u8f(plane2[i1]), xa1), __fmul_rn(u8f(plane2[i2]), xa));
                    const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
                    // round half to even through the 1.5 * 2^23 addition; res lies in [0, 255]
                    o |= min(__float_as_uint(__fadd_rn(res, 12582912.0f)) & 0x1ffu, 255u) << (8 * k);
                }
                *reinterpret_cast<uint32_t*>(drow + 4 * c) = o;
            }
        } else {
            for (int x = lane; x < a.w; x += 32) {
                const uint32_t sv = srow[x];
                const uint32_t xi = s_xi[x];
                const uint32_t i1 = (xi & 0xffffu) + sv, i2 = (xi >> 16) + sv;
                const float xa = s_xa[x], xa1 = __fsub_rn(1.0f, xa);
                const float top = __fadd_rn(__fmul_rn(u8f(plane1[i1]), xa1), __fmul_rn(u8f(plane1[i2]), xa));
                const float bot = __fadd_rn(__fmul_rn(u8f(plane2[i1]), xa1), __fmul_rn(u8f(plane2[i2]), xa));
                const float res = __fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya));
                drow[x] = (uint8_t)min(__float_as_uint(__fadd_rn(res, 12582912.0f)) & 0x1ffu, 255u);
            }
        }
    }
}
