/*
 * oracle/stereo_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see stereo_oracle.h).
 * Follows /root/reference/src/Frame.cc:813-990 statement by statement; build with -ffp-contract=off.
 */
#include "stereo_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TH_HIGH 100 /* src/ORBmatcher.cc:36 */
#define TH_LOW 50   /* :37 */

/* ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:2349-2365: popcount of the XOR over 8 x int32 */
int orb_oracle_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t x, y;
        memcpy(&x, a + 4 * i, 4); memcpy(&y, b + 4 * i, 4);
        uint32_t v = x ^ y;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

typedef struct { int dist, idx; } DistIdx;
static int cmp_distidx(const void* a, const void* b) { /* pair<int,int> operator< */
    const DistIdx* x = (const DistIdx*)a; const DistIdx* y = (const DistIdx*)b;
    if (x->dist != y->dist) return x->dist < y->dist ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

int orb_oracle_stereo(int N, const OrbOracleKeyPoint* kL, const uint8_t* dL, int Nr, const OrbOracleKeyPoint* kR,
                      const uint8_t* dR, int nlevels, const float* scale, const float* inv_scale, const int* lw,
                      const int* lh, const uint8_t* const* pyrL, const size_t* pitchL, const uint8_t* const* pyrR,
                      const size_t* pitchR, float mb, float mbf, float* mvuRight, float* mvDepth) {
    (void)nlevels;
    for (int i = 0; i < N; ++i) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }         /* :815-816 */
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;                                      /* :818 */
    const int nRows = lh[0];                                                           /* :820 */
    /* row table, :823-841 */
    int* cnt = (int*)calloc((size_t)nRows, sizeof(int));
    int* minr = (int*)malloc(sizeof(int) * (size_t)(Nr > 0 ? Nr : 1));
    int* maxr = (int*)malloc(sizeof(int) * (size_t)(Nr > 0 ? Nr : 1));
    for (int iR = 0; iR < Nr; ++iR) {
        const float kpY = kR[iR].y;
        const float r = 2.0f * scale[kR[iR].octave];
        maxr[iR] = (int)ceilf(kpY + r);
        minr[iR] = (int)floorf(kpY - r);
        if (minr[iR] < 0 || maxr[iR] >= nRows) { free(cnt); free(minr); free(maxr); return -1; }
        for (int yi = minr[iR]; yi <= maxr[iR]; ++yi) cnt[yi]++;
    }
    int** rows = (int**)malloc(sizeof(int*) * (size_t)nRows);
    for (int y = 0; y < nRows; ++y) { rows[y] = (int*)malloc(sizeof(int) * (size_t)(cnt[y] > 0 ? cnt[y] : 1)); cnt[y] = 0; }
    for (int iR = 0; iR < Nr; ++iR)
        for (int yi = minr[iR]; yi <= maxr[iR]; ++yi) rows[yi][cnt[yi]++] = iR;

    const float minZ = mb, minD = 0, maxD = mbf / minZ;                                /* :844-846 */
    DistIdx* vDistIdx = (DistIdx*)malloc(sizeof(DistIdx) * (size_t)(N > 0 ? N : 1));
    int nDist = 0;
    for (int iL = 0; iL < N; ++iL) {                                                   /* :852 */
        const int levelL = kL[iL].octave;
        const float vL = kL[iL].y, uL = kL[iL].x;
        const int row = (int)vL;                                                       /* vRowIndices[vL] */
        if (row < 0 || row >= nRows || cnt[row] == 0) continue;                        /* :861 */
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;                                                        /* :867 */
        int bestDist = TH_HIGH;
        int bestIdxR = 0;
        for (int iC = 0; iC < cnt[row]; ++iC) {                                        /* :876-897 */
            const int iR = rows[row][iC];
            if (kR[iR].octave < levelL - 1 || kR[iR].octave > levelL + 1) continue;
            const float uR = kR[iR].x;
            if (uR >= minU && uR <= maxU) {
                const int dist = orb_oracle_descriptor_distance(dL + 32 * (size_t)iL, dR + 32 * (size_t)iR);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {                                                    /* :900 */
            const float uR0 = kR[bestIdxR].x;
            const float scaleFactor = inv_scale[levelL];
            const float scaleduL = roundf(uL * scaleFactor);
            const float scaledvL = roundf(vL * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5, L = 5;
            const uint8_t* imL = pyrL[levelL]; const uint8_t* imR = pyrR[levelL];
            const size_t pL = pitchL[levelL], pR = pitchR[levelL];
            const int y0 = (int)(scaledvL - w), x0 = (int)(scaleduL - w);
            short IL[11][11];
            const short cL = (short)imL[(size_t)(y0 + w) * pL + (x0 + w)];
            for (int r = 0; r < 11; ++r)
                for (int c = 0; c < 11; ++c) IL[r][c] = (short)((short)imL[(size_t)(y0 + r) * pL + (x0 + c)] - cL);  /* :912-913 */
            int bestDistW = INT_MAX, bestincR = 0;
            float vDists[11];
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= lw[levelL]) continue;                              /* :923-924 */
            for (int incR = -L; incR <= L; ++incR) {                                   /* :926-943 */
                const int xr0 = (int)(scaleduR0 + incR - w);
                const short cR = (short)imR[(size_t)(y0 + w) * pR + (xr0 + w)];
                long long acc = 0;
                for (int r = 0; r < 11; ++r)
                    for (int c = 0; c < 11; ++c) {
                        const int d = (int)IL[r][c] - ((int)imR[(size_t)(y0 + r) * pR + (xr0 + c)] - (int)cR);
                        acc += d < 0 ? -d : d;
                    }
                const float dist = (float)(double)acc;
                if (dist < bestDistW) { bestDistW = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;                             /* :945-946 */
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));   /* :953 */
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scale[levelL] * ((float)scaleduR0 + (float)bestincR + deltaR);      /* :959 */
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {                               /* :963 */
                if (disparity <= 0) {
                    disparity = 0.01;
                    bestuR = uL - 0.01;
                }
                mvDepth[iL] = mbf / disparity;
                mvuRight[iL] = bestuR;
                vDistIdx[nDist].dist = bestDistW; vDistIdx[nDist].idx = iL; nDist++;
            }
        }
    }
    int kept = nDist;
    if (nDist > 0) {                                     /* the reference indexes an empty vector otherwise (UB) */
        qsort(vDistIdx, (size_t)nDist, sizeof(DistIdx), cmp_distidx);                  /* :977 */
        const float median = (float)vDistIdx[nDist / 2].dist;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = nDist - 1; i >= 0; --i) {                                         /* :981-990 */
            if (vDistIdx[i].dist < thDist) break;
            mvuRight[vDistIdx[i].idx] = -1;
            mvDepth[vDistIdx[i].idx] = -1;
            kept--;
        }
    }
    for (int y = 0; y < nRows; ++y) free(rows[y]);
    free(rows); free(cnt); free(minr); free(maxr); free(vDistIdx);
    return kept;
}
