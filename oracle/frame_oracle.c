/*
 * oracle/frame_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see frame_oracle.h).
 * Follows /root/reference/src/Frame.cc and src/ORBmatcher.cc statement by statement; build with -ffp-contract=off.
 */
#include "frame_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "cv_prims.h"
#include "stereo_oracle.h"

#define GCOLS ORB_ORACLE_GRID_COLS
#define GROWS ORB_ORACLE_GRID_ROWS
#define TH_LOW 50       /* src/ORBmatcher.cc:37 */
#define HISTO_LENGTH 30 /* :38 */

static float grid_w_inv(const OrbOracleCalib* c) { return (float)GCOLS / (c->max_x - c->min_x); } /* src/Frame.cc:339 */
static float grid_h_inv(const OrbOracleCalib* c) { return (float)GROWS / (c->max_y - c->min_y); } /* :340 */

/* Frame::ComputeImageBounds, src/Frame.cc:784-812 */
void orb_oracle_image_bounds(OrbOracleCalib* c, int width, int height) {
    if (c->dist[0] != 0.0f) {
        float m[8] = {0.f, 0.f, (float)width, 0.f, 0.f, (float)height, (float)width, (float)height};
        ocv_undistort_points_f32(m, 4, c->fx, c->fy, c->cx, c->cy, c->dist, c->n_dist, m);
        c->min_x = fminf(m[0], m[4]);
        c->max_x = fmaxf(m[2], m[6]);
        c->min_y = fminf(m[1], m[3]);
        c->max_y = fmaxf(m[5], m[7]);
    } else {
        c->min_x = 0.0f; c->max_x = (float)width; c->min_y = 0.0f; c->max_y = (float)height;
    }
}

/* Frame::UndistortKeyPoints, src/Frame.cc:748-782 */
void orb_oracle_undistort_keypoints(const OrbOracleCalib* c, const OrbOracleKeyPoint* keys, int n, OrbOracleKeyPoint* keys_un) {
    if (keys_un != keys) memcpy(keys_un, keys, sizeof(OrbOracleKeyPoint) * (size_t)n);
    if (c->dist[0] == 0.0f || n == 0) return;
    float* m = (float*)malloc(sizeof(float) * 2 * (size_t)n);
    for (int i = 0; i < n; ++i) { m[2 * i] = keys[i].x; m[2 * i + 1] = keys[i].y; }
    ocv_undistort_points_f32(m, n, c->fx, c->fy, c->cx, c->cy, c->dist, c->n_dist, m);
    for (int i = 0; i < n; ++i) { keys_un[i].x = m[2 * i]; keys_un[i].y = m[2 * i + 1]; }
    free(m);
}

/* Frame::PosInGrid, src/Frame.cc:726-736 */
static int pos_in_grid(const OrbOracleCalib* c, const OrbOracleKeyPoint* kp, int* px, int* py) {
    const int posX = (int)roundf((kp->x - c->min_x) * grid_w_inv(c));
    const int posY = (int)roundf((kp->y - c->min_y) * grid_h_inv(c));
    if (posX < 0 || posX >= GCOLS || posY < 0 || posY >= GROWS) return 0;
    *px = posX; *py = posY;
    return 1;
}

/* Frame::AssignFeaturesToGrid, src/Frame.cc:383-417 (Nleft == -1) */
int orb_oracle_assign_grid(const OrbOracleCalib* c, const OrbOracleKeyPoint* keys_un, int n, int32_t* cell_start, int32_t* cell_items) {
    const int ncell = GCOLS * GROWS;
    int* cell = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    memset(cell_start, 0, sizeof(int32_t) * (size_t)(ncell + 1));
    int placed = 0;
    for (int i = 0; i < n; ++i) {
        int gx, gy;
        cell[i] = -1;
        if (pos_in_grid(c, keys_un + i, &gx, &gy)) { cell[i] = gx * GROWS + gy; cell_start[cell[i] + 1]++; placed++; }
    }
    for (int k = 0; k < ncell; ++k) cell_start[k + 1] += cell_start[k];
    int* fill = (int*)calloc((size_t)ncell, sizeof(int));
    for (int i = 0; i < n; ++i)   /* push_back in index order */
        if (cell[i] >= 0) cell_items[cell_start[cell[i]] + fill[cell[i]]++] = i;
    free(fill); free(cell);
    return placed;
}

/* Frame::GetFeaturesInArea, src/Frame.cc:655-724 */
int orb_oracle_features_in_area(const OrbOracleCalib* c, const OrbOracleKeyPoint* keys_un, const int32_t* cell_start,
                                const int32_t* cell_items, float x, float y, float r, int minLevel, int maxLevel,
                                int32_t* out, int cap) {
    int n = 0;
    const float factorX = r, factorY = r;
    const float wInv = grid_w_inv(c), hInv = grid_h_inv(c);
    int nMinCellX = (int)floorf((x - c->min_x - factorX) * wInv);
    if (nMinCellX < 0) nMinCellX = 0;
    if (nMinCellX >= GCOLS) return 0;
    int nMaxCellX = (int)ceilf((x - c->min_x + factorX) * wInv);
    if (nMaxCellX > GCOLS - 1) nMaxCellX = GCOLS - 1;
    if (nMaxCellX < 0) return 0;
    int nMinCellY = (int)floorf((y - c->min_y - factorY) * hInv);
    if (nMinCellY < 0) nMinCellY = 0;
    if (nMinCellY >= GROWS) return 0;
    int nMaxCellY = (int)ceilf((y - c->min_y + factorY) * hInv);
    if (nMaxCellY > GROWS - 1) nMaxCellY = GROWS - 1;
    if (nMaxCellY < 0) return 0;
    const int bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
        for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
            const int cidx = ix * GROWS + iy;
            for (int j = cell_start[cidx]; j < cell_start[cidx + 1]; ++j) {
                const OrbOracleKeyPoint* kp = keys_un + cell_items[j];
                if (bCheckLevels) {
                    if (kp->octave < minLevel) continue;
                    if (maxLevel >= 0 && kp->octave > maxLevel) continue;
                }
                const float distx = kp->x - x, disty = kp->y - y;
                if (fabsf(distx) < factorX && fabsf(disty) < factorY) {
                    if (n < cap) out[n] = cell_items[j];
                    n++;
                }
            }
        }
    return n;
}

/* ORBmatcher::ComputeThreeMaxima, src/ORBmatcher.cc:2303-2344 (on bin sizes) */
static void three_maxima(const int* histo, int L, int* ind1, int* ind2, int* ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = histo[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; *ind3 = *ind2; *ind2 = *ind1; *ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; *ind3 = *ind2; *ind2 = i; }
        else if (s > max3) { max3 = s; *ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { *ind2 = -1; *ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { *ind3 = -1; }
}

/* ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:705-814 */
int orb_oracle_search_for_initialization(const OrbOracleCalib* c, const OrbOracleKeyPoint* k1, const uint8_t* d1, int n1,
                                         const OrbOracleKeyPoint* k2, const uint8_t* d2, int n2, const int32_t* cell_start2,
                                         const int32_t* cell_items2, float* prev, int windowSize, float nnratio,
                                         int checkOrientation, int32_t* vnMatches12) {
    int nmatches = 0;
    for (int i = 0; i < n1; ++i) vnMatches12[i] = -1;
    int* rotBin = (int*)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1));   /* push order inside a bin is irrelevant */
    int rotCount[HISTO_LENGTH] = {0};
    int nPushed = 0;
    int* pushedIdx = (int*)malloc(sizeof(int) * (size_t)(n1 > 0 ? n1 : 1));
    const float factor = 1.0f / HISTO_LENGTH;
    int* vMatchedDistance = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    int* vnMatches21 = (int*)malloc(sizeof(int) * (size_t)(n2 > 0 ? n2 : 1));
    int32_t* vIndices2 = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n2 > 0 ? n2 : 1));
    for (int i = 0; i < n2; ++i) { vMatchedDistance[i] = INT_MAX; vnMatches21[i] = -1; }

    for (int i1 = 0; i1 < n1; i1++) {
        const int level1 = k1[i1].octave;
        if (level1 > 0) continue;
        const int nInd = orb_oracle_features_in_area(c, k2, cell_start2, cell_items2, prev[2 * i1], prev[2 * i1 + 1],
                                                     (float)windowSize, level1, level1, vIndices2, n2);
        if (nInd == 0) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int v = 0; v < nInd; ++v) {
            const int i2 = vIndices2[v];
            const int dist = orb_oracle_descriptor_distance(d1 + 32 * (size_t)i1, d2 + 32 * (size_t)i2);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) { bestDist2 = dist; }
        }
        if (bestDist <= TH_LOW) {
            if ((float)bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                vnMatches12[i1] = bestIdx2;
                vnMatches21[bestIdx2] = i1;
                vMatchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOrientation) {
                    float rot = k1[i1].angle - k2[bestIdx2].angle;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)roundf(rot * factor);
                    if (bin == HISTO_LENGTH) bin = 0;
                    rotBin[nPushed] = bin; pushedIdx[nPushed] = i1; nPushed++;
                    rotCount[bin]++;
                }
            }
        }
    }
    if (checkOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotCount, HISTO_LENGTH, &ind1, &ind2, &ind3);
        for (int p = 0; p < nPushed; ++p) {
            const int b = rotBin[p];
            if (b == ind1 || b == ind2 || b == ind3) continue;
            const int idx1 = pushedIdx[p];
            if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)
        if (vnMatches12[i1] >= 0) { prev[2 * i1] = k2[vnMatches12[i1]].x; prev[2 * i1 + 1] = k2[vnMatches12[i1]].y; }
    free(rotBin); free(pushedIdx); free(vMatchedDistance); free(vnMatches21); free(vIndices2);
    return nmatches;
}
