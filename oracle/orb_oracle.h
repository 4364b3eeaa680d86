/*
 * oracle/orb_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference hot path (Frame::ExtractORB -> ORBextractor::operator()),
 * following /root/reference/src/orb_extractor/ORBextractor.cc line by line in behaviour (citations in
 * orb_oracle.c).  Oracle definition: reference code + cv2-4.13.0-equivalent
 * primitives (oracle/cv_prims.c) + the monotonic-allocator tie rule in DistributeOctTree
 * ("equal node size => later-created node is split first").
 *
 * PINNING: the reference ships no golden vectors for this path.  This restatement is pinned against
 * (a) the unmodified reference compiled here (oracle/_ref/ref_extract_bump, outputs committed as
 * tests/golden/ref_*.npz by tests/golden/make_golden.py) and (b) cv2 4.13.0 for the primitives.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 */
#ifndef ORACLE_ORB_ORACLE_H_
#define ORACLE_ORB_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrbOracleKeyPoint { /* cv::KeyPoint layout, 28 bytes */
    float x, y, size, angle, response;
    int32_t octave, class_id;
} OrbOracleKeyPoint;

typedef struct OrbOracle OrbOracle;

/* cell_w: the constant W of ComputeKeyPointsOctTree (30 in the reference, ORBextractor.cc:777). */
OrbOracle* orb_oracle_create(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int cell_w);
void orb_oracle_destroy(OrbOracle* o);

/* Constructor tables (ORBextractor.cc:408-475). which: 0 mvScaleFactor, 1 mvInvScaleFactor,
 * 2 mvLevelSigma2, 3 mvInvLevelSigma2. */
void orb_oracle_scale_table(const OrbOracle* o, int which, float* out);
void orb_oracle_quota(const OrbOracle* o, int* out);   /* mnFeaturesPerLevel */
void orb_oracle_umax(const OrbOracle* o, int* out16);  /* umax[0..15] */

/* operator() (ORBextractor.cc:1078-1162).  Returns the reference's return value (monoIndex, or -1 for
 * an empty image), or -2 if a level is too small for the cell grid (UB in the reference).
 * kps/desc receive min(n, cap) entries; *n_out the total count. */
int orb_oracle_extract(OrbOracle* o, const uint8_t* img, int w, int h, size_t stride, int lap0, int lap1,
                       OrbOracleKeyPoint* kps, uint8_t* desc, int cap, int* n_out);

/* Per-stage state of the last orb_oracle_extract call. */
int orb_oracle_level_size(const OrbOracle* o, int level, int* w, int* h);
/* bordered plane, (h+38) rows of (w+38) bytes, pitch returned */
const uint8_t* orb_oracle_level_plane(const OrbOracle* o, int level, size_t* pitch);
/* blurred level (w x h, pitch w); NULL if the level had no keypoints (blur skipped, :1122) */
const uint8_t* orb_oracle_level_blur(const OrbOracle* o, int level);
/* FAST candidates of the cell loop in emission order, coordinates relative to (16,16) (:855-860) */
int orb_oracle_level_candidates(const OrbOracle* o, int level, int* xs, int* ys, int* scores, int cap);
/* allKeypoints[level] after orientation (level coordinates; == allLevelsKeypoints, :1094) */
int orb_oracle_level_keypoints(const OrbOracle* o, int level, OrbOracleKeyPoint* kps, int cap);

/* Stand-alone stages for unit tests. */
int orb_oracle_distribute(const int* xs, const int* ys, const int* scores, int n,
                          int minX, int maxX, int minY, int maxY, int N, int* kept_idx, int cap);
float orb_oracle_ic_angle(const uint8_t* center, size_t step, const int* umax16);
void orb_oracle_descriptor(const uint8_t* center, size_t step, float angle_deg, uint8_t* desc32);

#ifdef __cplusplus
}
#endif
#endif
