/*
 * oracle/stereo_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * CPU restatement of Frame::ComputeStereoMatches (reference src/Frame.cc:813-990) with
 * ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2349-2365), TH_HIGH = 100, TH_LOW = 50 (:36-37).
 * Pinned against the unmodified function compiled by oracle/Makefile (oracle/_ref/ref_stereo) through
 * tests/golden/stereo_*.npz.
 */
#ifndef ORACLE_STEREO_ORACLE_H_
#define ORACLE_STEREO_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#include "orb_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Level l of the left/right pyramid: lw[l] x lh[l] pixels at pyrL[l]/pyrR[l] with row pitch pitchL[l]/pitchR[l].
 * Returns the number of matches kept (depth > 0), or -1 if a keypoint's row band leaves the image (UB there). */
int orb_oracle_stereo(int nL, const OrbOracleKeyPoint* kL, const uint8_t* dL, int nR, const OrbOracleKeyPoint* kR,
                      const uint8_t* dR, int nlevels, const float* scale, const float* inv_scale, const int* lw,
                      const int* lh, const uint8_t* const* pyrL, const size_t* pitchL, const uint8_t* const* pyrR,
                      const size_t* pitchR, float mb, float mbf, float* uRight, float* depth);
int orb_oracle_descriptor_distance(const uint8_t* a, const uint8_t* b);
#ifdef __cplusplus
}
#endif
#endif
