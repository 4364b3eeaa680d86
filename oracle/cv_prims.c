/*
 * oracle/cv_prims.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see cv_prims.h).
 *
 * Plain-C restatements of the OpenCV primitives called by the reference hot path.  They follow the
 * published OpenCV 4.x algorithms (modules/imgproc/src/resize.cpp, modules/features2d/src/fast.cpp +
 * fast_score.cpp, modules/imgproc/src/smooth.dispatch.cpp fixed-point path, modules/core/src/
 * mathfuncs_core.simd.hpp fastAtan2) and are pinned against python cv2 4.13.0 by the test-suite.
 * Build with -ffp-contract=off: the reference is built without FMA (CMakeLists.txt:4-11).
 */
#include "cv_prims.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

int ocv_round_f(float v) { return (int)lrintf(v); }
int ocv_round_d(double v) { return (int)lrint(v); }

static int ocv_floor_f(float v) {
    int i = (int)v;
    return i - (i > v);
}

/* ------------------------------------------------------------------------------------------------
 * resize, INTER_LINEAR, 8UC1.  Call site: ORBextractor.cc:1183 (ComputePyramid).
 * Fixed point: 11 coefficient bits (x2048) per axis, int32 accumulate, the vertical pass drops 4
 * bits before and 16+2 after the multiply with a +2 rounding term.
 * ---------------------------------------------------------------------------------------------- */
static void linear_axis_tables(int ssize, int dsize, int* ofs, short* coef) {
    const double inv_scale = (double)dsize / ssize;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = ocv_floor_f(f);
        f -= (float)s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
        ofs[d] = s;
        coef[2 * d + 0] = (short)ocv_round_f((1.f - f) * 2048.f);
        coef[2 * d + 1] = (short)ocv_round_f(f * 2048.f);
    }
}

void ocv_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep,
                          uint8_t* dst, int dw, int dh, size_t dstep) {
    if (sw == dw && sh == dh) {
        for (int y = 0; y < dh; ++y) memcpy(dst + (size_t)y * dstep, src + (size_t)y * sstep, (size_t)dw);
        return;
    }
    /* OpenCV switches an exact 2x2 decimation from INTER_LINEAR to the INTER_AREA fast path. */
    if (sw == 2 * dw && sh == 2 * dh) {
        for (int y = 0; y < dh; ++y) {
            const uint8_t* s0 = src + (size_t)(2 * y) * sstep;
            const uint8_t* s1 = s0 + sstep;
            uint8_t* d = dst + (size_t)y * dstep;
            for (int x = 0; x < dw; ++x)
                d[x] = (uint8_t)((s0[2 * x] + s0[2 * x + 1] + s1[2 * x] + s1[2 * x + 1] + 2) >> 2);
        }
        return;
    }
    int* xofs = (int*)malloc(sizeof(int) * (size_t)dw);
    int* yofs = (int*)malloc(sizeof(int) * (size_t)dh);
    short* xa = (short*)malloc(sizeof(short) * 2 * (size_t)dw);
    short* ya = (short*)malloc(sizeof(short) * 2 * (size_t)dh);
    int* row0 = (int*)malloc(sizeof(int) * (size_t)dw);
    int* row1 = (int*)malloc(sizeof(int) * (size_t)dw);
    linear_axis_tables(sw, dw, xofs, xa);
    linear_axis_tables(sh, dh, yofs, ya);
    for (int y = 0; y < dh; ++y) {
        int sy0 = yofs[y];
        int sy1 = sy0 + 1 < sh ? sy0 + 1 : sh - 1;
        const uint8_t* s0 = src + (size_t)sy0 * sstep;
        const uint8_t* s1 = src + (size_t)sy1 * sstep;
        for (int x = 0; x < dw; ++x) {
            int sx0 = xofs[x];
            int sx1 = sx0 + 1 < sw ? sx0 + 1 : sw - 1;
            row0[x] = s0[sx0] * xa[2 * x] + s0[sx1] * xa[2 * x + 1];
            row1[x] = s1[sx0] * xa[2 * x] + s1[sx1] * xa[2 * x + 1];
        }
        const int b0 = ya[2 * y], b1 = ya[2 * y + 1];
        uint8_t* d = dst + (size_t)y * dstep;
        for (int x = 0; x < dw; ++x) {
            int v = (((b0 * (row0[x] >> 4)) >> 16) + ((b1 * (row1[x] >> 4)) >> 16) + 2) >> 2;
            d[x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(xofs); free(yofs); free(xa); free(ya); free(row0); free(row1);
}

/* ------------------------------------------------------------------------------------------------
 * copyMakeBorder, BORDER_REFLECT_101.  Call sites: ORBextractor.cc:1193, :1213.
 * gfedcb|abcdefgh|gfedcba  (edge pixel not repeated).
 * ---------------------------------------------------------------------------------------------- */
static int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

void ocv_copy_make_border_reflect101_u8(const uint8_t* src, int w, int h, size_t sstep,
                                        uint8_t* dst, size_t dstep,
                                        int top, int bottom, int left, int right) {
    const int dw = w + left + right;
    int* map = (int*)malloc(sizeof(int) * (size_t)dw);
    for (int x = 0; x < dw; ++x) map[x] = reflect101(x - left, w);
    uint8_t* tmp = (uint8_t*)malloc((size_t)w);
    /* interior rows first (src may alias dst's interior, hence the row copy) */
    for (int y = 0; y < h; ++y) {
        memcpy(tmp, src + (size_t)y * sstep, (size_t)w);
        uint8_t* d = dst + (size_t)(y + top) * dstep;
        for (int x = 0; x < dw; ++x) d[x] = tmp[map[x]];
    }
    for (int y = 0; y < top; ++y)
        memcpy(dst + (size_t)y * dstep, dst + (size_t)(top + reflect101(y - top, h)) * dstep, (size_t)dw);
    for (int y = 0; y < bottom; ++y)
        memcpy(dst + (size_t)(top + h + y) * dstep, dst + (size_t)(top + reflect101(h + y, h)) * dstep, (size_t)dw);
    free(map); free(tmp);
}

/* ------------------------------------------------------------------------------------------------
 * FAST-9/16 with non-max suppression.  Call sites: ORBextractor.cc:818, :837.
 * ---------------------------------------------------------------------------------------------- */
static const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

int ocv_fast9_16_best(const uint8_t* p, size_t step) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[(ptrdiff_t)kRingDy[k] * (ptrdiff_t)step + kRingDx[k]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int i = 1; i < 9; ++i) {
            if (d[k + i] < mn) mn = d[k + i];
            if (d[k + i] > mx) mx = d[k + i];
        }
        if (mn > best) best = mn;   /* arc all darker than the centre by at least mn   */
        if (-mx > best) best = -mx; /* arc all brighter than the centre by at least -mx */
    }
    return best;
}

/* Score of a pixel already known to be a corner at `threshold`: max(threshold, best) - 1, evaluated
 * with the pairwise min/max chain over the 16 arcs (d[] holds the 16 differences, wrapped to 25). */
static int fast_score_chain(const int* d, int threshold) {
    int a0 = threshold;
    for (int k = 0; k < 16; k += 2) {
        int a = d[k + 1] < d[k + 2] ? d[k + 1] : d[k + 2];
        if (d[k + 3] < a) a = d[k + 3];
        if (a <= a0) continue;
        for (int i = 4; i <= 8; ++i) if (d[k + i] < a) a = d[k + i];
        int e0 = a < d[k] ? a : d[k];
        int e1 = a < d[k + 9] ? a : d[k + 9];
        if (e0 > a0) a0 = e0;
        if (e1 > a0) a0 = e1;
    }
    int b0 = -a0;
    for (int k = 0; k < 16; k += 2) {
        int b = d[k + 1] > d[k + 2] ? d[k + 1] : d[k + 2];
        if (d[k + 3] > b) b = d[k + 3];
        if (b >= b0) continue;
        for (int i = 4; i <= 8; ++i) if (d[k + i] > b) b = d[k + i];
        int e0 = b > d[k] ? b : d[k];
        int e1 = b > d[k + 9] ? b : d[k + 9];
        if (e0 < b0) b0 = e0;
        if (e1 < b0) b0 = e1;
    }
    return -b0 - 1;
}

int ocv_fast9_16_nms(const uint8_t* img, int w, int h, size_t step, int threshold,
                     int* xs, int* ys, int* scores, int cap) {
    if (w < 7 || h < 7) return 0;
    ptrdiff_t off[25];
    for (int k = 0; k < 25; ++k) off[k] = (ptrdiff_t)kRingDy[k & 15] * (ptrdiff_t)step + kRingDx[k & 15];
    if (threshold < 0) threshold = 0;
    if (threshold > 255) threshold = 255;
    /* cls[255 + (ring - centre)]: 1 = ring darker than centre - t, 2 = brighter than centre + t */
    uint8_t cls[512];
    for (int i = -255; i <= 255; ++i) cls[i + 255] = (uint8_t)(i < -threshold ? 1 : (i > threshold ? 2 : 0));
    /* three rolling rows of scores, zero outside the 3-pixel interior */
    uint8_t* rows = (uint8_t*)calloc((size_t)w * 3, 1);
    int n = 0;
    for (int y = 3; y < h - 2; ++y) {
        uint8_t* cur = rows + (size_t)((y - 3) % 3) * w;
        memset(cur, 0, (size_t)w);
        if (y < h - 3) {
            const uint8_t* p = img + (size_t)y * step + 3;
            for (int x = 3; x < w - 3; ++x, ++p) {
                const int v = p[0];
                const uint8_t* t = cls + 255 - v;
                int m = t[p[off[0]]] | t[p[off[8]]];
                if (!m) continue;
                m &= t[p[off[2]]] | t[p[off[10]]];
                m &= t[p[off[4]]] | t[p[off[12]]];
                m &= t[p[off[6]]] | t[p[off[14]]];
                if (!m) continue;
                m &= t[p[off[1]]] | t[p[off[9]]];
                m &= t[p[off[3]]] | t[p[off[11]]];
                m &= t[p[off[5]]] | t[p[off[13]]];
                m &= t[p[off[7]]] | t[p[off[15]]];
                if (!m) continue;
                int corner = 0;
                if (m & 1) { /* look for 9 contiguous ring pixels darker than v - t */
                    const int lim = v - threshold;
                    int run = 0;
                    for (int k = 0; k < 25 && !corner; ++k) {
                        if (p[off[k]] < lim) { if (++run > 8) corner = 1; } else run = 0;
                    }
                }
                if (!corner && (m & 2)) {
                    const int lim = v + threshold;
                    int run = 0;
                    for (int k = 0; k < 25 && !corner; ++k) {
                        if (p[off[k]] > lim) { if (++run > 8) corner = 1; } else run = 0;
                    }
                }
                if (corner) {
                    int d[25];
                    for (int k = 0; k < 25; ++k) d[k] = v - p[off[k]];
                    cur[x] = (uint8_t)fast_score_chain(d, threshold);
                }
            }
        }
        if (y == 3) continue;
        /* non-max suppression of the previous row (y-1) against rows y-2, y-1, y */
        const uint8_t* prev = rows + (size_t)((y - 4) % 3) * w;
        const uint8_t* pprev = rows + (size_t)((y - 5 + 3) % 3) * w;
        for (int x = 3; x < w - 3; ++x) {
            const int s = prev[x];
            if (!s) continue;
            if (s > prev[x - 1] && s > prev[x + 1] && s > pprev[x - 1] && s > pprev[x] && s > pprev[x + 1] &&
                s > cur[x - 1] && s > cur[x] && s > cur[x + 1]) {
                if (n < cap) { xs[n] = x; ys[n] = y - 1; scores[n] = s; }
                ++n;
            }
        }
    }
    free(rows);
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * GaussianBlur 7x7, sigma 2, BORDER_REFLECT_101, 8UC1.  Call site: ORBextractor.cc:1127.
 * OpenCV >= 3.4 fixed-point path: 8.8 kernel {18,34,48,56,48,34,18}, exact 16.16 accumulate,
 * one rounding at the end.
 * ---------------------------------------------------------------------------------------------- */
void ocv_gaussian_blur_7x7_s2_u8(const uint8_t* src, int w, int h, size_t sstep,
                                 uint8_t* dst, size_t dstep) {
    uint16_t* hbuf = (uint16_t*)malloc(sizeof(uint16_t) * (size_t)w * (size_t)h);
    uint8_t* pad = (uint8_t*)malloc((size_t)w + 6);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * sstep;
        for (int i = 0; i < 3; ++i) { pad[i] = s[reflect101(i - 3, w)]; pad[w + 3 + i] = s[reflect101(w + i, w)]; }
        memcpy(pad + 3, s, (size_t)w);
        uint16_t* hr = hbuf + (size_t)y * w;
        for (int x = 0; x < w; ++x)
            hr[x] = (uint16_t)(18 * (pad[x] + pad[x + 6]) + 34 * (pad[x + 1] + pad[x + 5]) +
                               48 * (pad[x + 2] + pad[x + 4]) + 56 * pad[x + 3]);
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* r[7];
        for (int j = 0; j < 7; ++j) r[j] = hbuf + (size_t)reflect101(y + j - 3, h) * w;
        uint8_t* d = dst + (size_t)y * dstep;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 18u * ((uint32_t)r[0][x] + r[6][x]) + 34u * ((uint32_t)r[1][x] + r[5][x]) +
                           48u * ((uint32_t)r[2][x] + r[4][x]) + 56u * (uint32_t)r[3][x];
            d[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(hbuf); free(pad);
}

/* ------------------------------------------------------------------------------------------------
 * fastAtan2 (scalar).  Call site: ORBextractor.cc:101 (IC_Angle).  All float32, no FMA.
 * ---------------------------------------------------------------------------------------------- */
float ocv_fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.1415926535897932384626433832795);
    const float p1 = 0.9997878412794807f * scale;
    const float p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale;
    const float p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------------------------------------------
 * undistortPoints (OpenCV calib3d cvUndistortPointsInternal, default TermCriteria(MAX_ITER, 5, 0.01) so only the
 * iteration count applies; no rotation, no tilt, P = K).  Call sites: src/Frame.cc:767, :797.
 * Pinned against cv2 4.13.0 (tests/test_oracle_frame.py, tests/golden/frame_kat.npz).
 * ---------------------------------------------------------------------------------------------- */
void ocv_undistort_points_f32(const float* xy_in, int n, float fxf, float fyf, float cxf, float cyf, const float* dist,
                              int n_dist, float* xy_out) {
    double k[14] = {0};
    for (int i = 0; i < n_dist && i < 14; ++i) k[i] = (double)dist[i];
    const double fx = fxf, fy = fyf, cx = cxf, cy = cyf;
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; ++i) {
        double x = ((double)xy_in[2 * i] - cx) * ifx, y = ((double)xy_in[2 * i + 1] - cy) * ify;
        const double x0 = x, y0 = y;
        if (n_dist > 0) {
            for (int j = 0; j < 5; ++j) {
                const double r2 = x * x + y * y;
                const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                if (icdist < 0) { x = x0; y = y0; break; }
                const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
                const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
                x = (x0 - deltaX) * icdist;
                y = (y0 - deltaY) * icdist;
            }
        }
        /* RR = P * I: [fx 0 cx; 0 fy cy; 0 0 1] */
        const double xx = fx * x + 0. * y + cx, yy = 0. * x + fy * y + cy, ww = 1. / (0. * x + 0. * y + 1.);
        xy_out[2 * i] = (float)(xx * ww);
        xy_out[2 * i + 1] = (float)(yy * ww);
    }
}

/* ------------------------------------------------------------------------------------------------
 * CLAHE (OpenCV imgproc clahe.cpp: CLAHE_CalcLut_Body + CLAHE_Interpolation_Body, 8-bit path).
 * Call sites: src/orb_extractor/main_orb_extractor.cpp:19-22, src/clahe/main_clahe.cpp:7-11.
 * Pinned against cv2 4.13.0 (tests/test_clahe.py, tests/golden/clahe_kat.npz).
 * ---------------------------------------------------------------------------------------------- */
static int reflect101_idx(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

int ocv_clahe_u8(const uint8_t* src, int w, int h, size_t sstep, double clip_limit, int tilesX, int tilesY,
                 uint8_t* dst, size_t dstep) {
    if (!src || !dst || w <= 0 || h <= 0 || tilesX <= 0 || tilesY <= 0) return -1;
    const int histSize = 256;
    /* extended size for the LUTs (copyMakeBorder(..., 0, tilesY - rows % tilesY, 0, tilesX - cols % tilesX, REFLECT_101)) */
    int ew = w, eh = h;
    if (!(w % tilesX == 0 && h % tilesY == 0)) { ew = w + (tilesX - w % tilesX); eh = h + (tilesY - h % tilesY); }
    const int tw = ew / tilesX, th = eh / tilesY;
    const int tileSizeTotal = tw * th;
    const float lutScale = (float)(histSize - 1) / tileSizeTotal;
    int clipLimit = 0;
    if (clip_limit > 0.0) {
        clipLimit = (int)(clip_limit * tileSizeTotal / histSize);
        if (clipLimit < 1) clipLimit = 1;
    }
    uint8_t* lut = (uint8_t*)malloc((size_t)tilesX * tilesY * histSize);
    for (int k = 0; k < tilesX * tilesY; ++k) {
        const int ty = k / tilesX, tx = k % tilesX;
        int hist[256];
        memset(hist, 0, sizeof(hist));
        for (int y = ty * th; y < (ty + 1) * th; ++y) {
            const uint8_t* row = src + (size_t)reflect101_idx(y, h) * sstep;
            for (int x = tx * tw; x < (tx + 1) * tw; ++x) hist[row[reflect101_idx(x, w)]]++;
        }
        if (clipLimit > 0) {
            int clipped = 0;
            for (int i = 0; i < histSize; ++i)
                if (hist[i] > clipLimit) { clipped += hist[i] - clipLimit; hist[i] = clipLimit; }
            const int redistBatch = clipped / histSize;
            int residual = clipped - redistBatch * histSize;
            for (int i = 0; i < histSize; ++i) hist[i] += redistBatch;
            if (residual != 0) {
                const int residualStep = (histSize / residual) > 1 ? (histSize / residual) : 1;
                for (int i = 0; i < histSize && residual > 0; i += residualStep, residual--) hist[i]++;
            }
        }
        int sum = 0;
        for (int i = 0; i < histSize; ++i) {
            sum += hist[i];
            int v = ocv_round_f((float)sum * lutScale);
            lut[(size_t)k * histSize + i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
    for (int y = 0; y < h; ++y) {
        const float tyf = y * inv_th - 0.5f;
        int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
        const float ya = tyf - ty1, ya1 = 1.0f - ya;
        if (ty1 < 0) ty1 = 0;
        if (ty2 > tilesY - 1) ty2 = tilesY - 1;
        const uint8_t* lutPlane1 = lut + (size_t)ty1 * tilesX * histSize;
        const uint8_t* lutPlane2 = lut + (size_t)ty2 * tilesX * histSize;
        for (int x = 0; x < w; ++x) {
            const float txf = x * inv_tw - 0.5f;
            int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
            const float xa = txf - tx1, xa1 = 1.0f - xa;
            if (tx1 < 0) tx1 = 0;
            if (tx2 > tilesX - 1) tx2 = tilesX - 1;
            const int srcVal = src[(size_t)y * sstep + x];
            const int ind1 = tx1 * histSize + srcVal, ind2 = tx2 * histSize + srcVal;
            const float res = (lutPlane1[ind1] * xa1 + lutPlane1[ind2] * xa) * ya1 + (lutPlane2[ind1] * xa1 + lutPlane2[ind2] * xa) * ya;
            int v = ocv_round_f(res);
            dst[(size_t)y * dstep + x] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
    free(lut);
    return 0;
}
