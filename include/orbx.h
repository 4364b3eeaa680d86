/*
 * include/orbx.h -- C-ABI of libextractorb_cuda.so (sm_100a CUDA kernels behind plain pointers).
 *
 * This is the drop-in boundary for the ONE hot path this project accelerates: the path
 * Frame::ExtractORB (reference src/Frame.cc:419-427) drives through ORBextractor::operator()
 * (reference src/orb_extractor/ORBextractor.cc:1078-1162, 5-argument twin ORBExtractor.cpp:980-1112).
 * The C++ class in include/ORBextractor.h only unwraps cv::InputArray/OutputArray and forwards here.
 * There is NO CPU fallback: every entry point fails with an error code when CUDA is unavailable.
 *
 * Conventions: every function returns ORBX_OK (0) or a negative OrbxStatus; nothing throws, aborts or
 * writes to stdout.  A handle owns one CUDA stream + workspace on one device and is NOT re-entrant
 * (like a reference ORBextractor instance, which mutates mvImagePyramid); distinct handles may be used
 * concurrently from different host threads (stereo left/right, reference src/Frame.cc:109-112).
 */
#ifndef ORBX_H_
#define ORBX_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ORBX_API __attribute__((visibility("default")))
#else
#define ORBX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum OrbxStatus {
    ORBX_OK = 0,
    ORBX_ERR_EMPTY_IMAGE = -1,        /* reference returns -1 for an empty image (ORBextractor.cc:1083) */
    ORBX_ERR_BAD_ARGUMENT = -2,
    ORBX_ERR_LEVEL_TOO_SMALL = -3,    /* a pyramid level has no 30-px cell / aspect < 0.5: UB in the reference */
    ORBX_ERR_IMAGE_TOO_LARGE = -4,    /* a level is wider/taller than 4127 px */
    ORBX_ERR_CUDA = -5,               /* see orbx_last_error() */
    ORBX_ERR_CAPACITY = -6,           /* caller's keypoint capacity too small; *n_out holds the need */
    ORBX_ERR_CANDIDATE_OVERFLOW = -7, /* more FAST candidates than the workspace holds even after regrowth */
    ORBX_ERR_NO_FRAME = -8,           /* accessor called before any extraction / frame index not resident */
    ORBX_ERR_NO_DEVICE = -9
} OrbxStatus;

/* Constructor arguments of ORBextractor (reference ORBextractor.cc:408-411) + build-side knobs. */
typedef struct OrbxParams {
    int32_t nfeatures;
    float scale_factor;
    int32_t nlevels;
    int32_t ini_th_fast;
    int32_t min_th_fast;
    int32_t cell_size;     /* W of ComputeKeyPointsOctTree; 0 -> 30 (reference ORBextractor.cc:777) */
    int32_t max_batch;     /* frames processed per internal launch group; 0 -> 1 */
    int32_t cand_per_cell; /* initial FAST-candidate capacity per grid cell; 0 -> 64 (grows on overflow) */
    int32_t flags;         /* ORBX_FLAG_* */
} OrbxParams;

#define ORBX_FLAG_PROFILE 1 /* record CUDA-event time per stage (orbx_stage_times) */
#define ORBX_FLAG_NO_GRAPH 2 /* never replay small launch groups as CUDA graphs */
#define ORBX_FLAG_SINGLE_STREAM 4 /* do not overlap consecutive launch groups on two compute streams */
#define ORBX_FLAG_COPY_ONLY 8 /* measurement aid: orbx_extract_batch moves the same bytes through the same staging slots, streams
                                 and events but launches no kernel (outputs are unspecified): the copy bound of the host pipeline */

/* Same 28-byte layout as cv::KeyPoint. */
typedef struct OrbxKeyPoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} OrbxKeyPoint;

typedef struct OrbxHandle OrbxHandle;

enum { ORBX_MEM_HOST = 0, ORBX_MEM_DEVICE = 1 };

/* Pipeline stages, in launch order (index into orbx_stage_times output). */
enum {
    ORBX_STAGE_PYRAMID = 0, /* ComputePyramid            ORBextractor.cc:1164-1219 */
    ORBX_STAGE_FAST = 1,    /* cell loop + cv::FAST      ORBextractor.cc:797-864   */
    ORBX_STAGE_OCTREE = 2,  /* DistributeOctTree         ORBextractor.cc:544-771   */
    ORBX_STAGE_BLUR = 3,    /* GaussianBlur 7x7          ORBextractor.cc:1126-1127 */
    ORBX_STAGE_DESCRIBE = 4,/* IC_Angle + rBRIEF + scatter  :75-145, :1131-1159     */
    ORBX_NUM_STAGES = 5
};

ORBX_API const char* orbx_status_string(int status);
ORBX_API const char* orbx_last_error(const OrbxHandle* h);

/* ORBextractor::ORBextractor (ORBextractor.cc:408-475). */
ORBX_API int orbx_create(const OrbxParams* params, int device, OrbxHandle** out);
ORBX_API void orbx_destroy(OrbxHandle* h);

/* Constructor tables: mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2 (nlevels floats
 * each), mnFeaturesPerLevel (nlevels ints), umax (16 ints).  Any pointer may be NULL. */
ORBX_API int orbx_get_tables(const OrbxHandle* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                    int32_t* features_per_level, int32_t* umax16);
/* The same tables straight from the constructor arguments (nfeatures, scale_factor, nlevels of *params), computed on
 * the host without touching CUDA: the reference builds them unconditionally in its constructor (ORBextractor.cc:419-474)
 * and ORB-SLAM3 copies them into every Frame (src/Frame.cc:97-103), so they must exist even when no device does. */
ORBX_API int orbx_ctor_tables(const OrbxParams* params, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                     int32_t* features_per_level, int32_t* umax16);
/* Upper bound of keypoints one frame can yield (sum over levels of max(N_l + 2, 4 * nIni)). */
ORBX_API int orbx_max_keypoints(const OrbxHandle* h, int width, int height);

/* ORBextractor::operator() (ORBextractor.cc:1078-1162) on one host image (CV_8UC1, `stride` bytes per
 * row).  kps/desc receive the keypoints and 32-byte descriptors in the reference's two-ended order
 * (lapping keypoints filled from the back); *n_out = total, *mono_out = the reference's return value. */
ORBX_API int orbx_extract(OrbxHandle* h, const uint8_t* image, int width, int height, size_t stride,
                 int lap0, int lap1, OrbxKeyPoint* kps, uint8_t* desc, int capacity,
                 int* n_out, int* mono_out);

/* The same path over n_frames independent frames (the data-parallel form: one launch group per
 * max_batch frames).  images/kps/desc/counts live in `in_mem`/`out_mem` memory (ORBX_MEM_*); frame f's
 * outputs start at kps + f*cap_per_frame, desc + f*cap_per_frame*32, counts + 2*f ({n, mono}).
 * `stream` is the cudaStream_t the kernels are launched on (NULL -> the handle's own stream).  Host
 * buffers (pinned for full speed) are pipelined group by group: H2D of group g+1, kernels of group g and D2H
 * of group g-1 overlap on three streams.  The call returns after all results have landed. */
ORBX_API int orbx_extract_batch(OrbxHandle* h, const uint8_t* images, int in_mem, int n_frames, int width, int height,
                       size_t row_stride, size_t frame_stride, int lap0, int lap1,
                       OrbxKeyPoint* kps, uint8_t* desc, int cap_per_frame, int32_t* counts, int out_mem,
                       void* stream);

/* The same call over several devices of one box from one process: handles[i] (each created on its own device) is driven by
 * its own host thread through the pipeline above, and the threads pull launch groups from one shared cursor -- the reference's
 * own concurrency pattern (two extractor instances on two host threads, src/Frame.cc:109-112) widened to n devices.  A device
 * that is fed more slowly takes fewer groups, so all finish together.  images / kps / desc / counts are HOST buffers (pinned
 * for full speed), laid out as for orbx_extract_batch; frames_per_handle (n_handles ints, may be NULL) receives how many
 * frames each handle processed.  Results do not depend on which device processed a frame. */
ORBX_API int orbx_extract_batch_multi(OrbxHandle* const* handles, int n_handles, const uint8_t* images, int n_frames, int width,
                             int height, size_t row_stride, size_t frame_stride, int lap0, int lap1, OrbxKeyPoint* kps,
                             uint8_t* desc, int cap_per_frame, int32_t* counts, int32_t* frames_per_handle);

/* The launch-group sizes a host-buffer call of n_frames frames is cut into (sizes[0..return value), at most `capacity` written):
 * groups ramp up at the start and down at the end of a call, never above max_batch.  `consumers` > 1: the handles of
 * orbx_extract_batch_multi taking turns.  Needs no device. */
ORBX_API int orbx_plan_groups(int n_frames, int max_batch, int consumers, int ramp, int32_t* sizes, int capacity);

/* Stage-wise entry points mirroring the reference's public methods (inc/ORBextractor.h:89-92), which the
 * demos call directly (src/orb_extractor/main_orb_extractor.cpp:44-46). */
ORBX_API int orbx_compute_pyramid(OrbxHandle* h, const uint8_t* image, int width, int height, size_t stride);
ORBX_API int orbx_compute_keypoints_octtree(OrbxHandle* h); /* on the resident pyramid; results via orbx_get_level_keypoints */
/* ORBextractor::DistributeOctTree (ORBextractor.cc:544-771) on caller-supplied keypoints (host).
 * Ties on response keep the earlier input keypoint; equal-size nodes split later-created first. */
ORBX_API int orbx_distribute_octtree(OrbxHandle* h, const OrbxKeyPoint* keys, int n, int min_x, int max_x, int min_y,
                            int max_y, int n_features, OrbxKeyPoint* out, int capacity, int* n_out);

/* State of the most recent extraction.  `frame` indexes the frames of the last launch group
 * (0 for orbx_extract). */
ORBX_API int orbx_get_level_size(const OrbxHandle* h, int level, int* width, int* height);
/* mvImagePyramid[level] (inc/ORBextractor.h:85): with_border=0 copies the w x h level, 1 the
 * (w+38) x (h+38) plane including the 19-px reflect-101 border. */
ORBX_API int orbx_get_pyramid_level(OrbxHandle* h, int frame, int level, uint8_t* dst, size_t dst_stride, int with_border);
/* mvImagePyramid in one piece.  The library keeps a frame's bordered planes in one block of *frame_bytes bytes: level l starts
 * at plane_offset[l], has level_h[l] + 2 * ORBX_PLANE_BORDER rows of pitch[l] bytes, and its level pixel (0, 0) sits at row
 * ORBX_PLANE_BORDER, byte ORBX_PLANE_PADL of a row (the 19-px reflect-101 border lies around it, reference :1173-1177, :1193).
 * Arrays take nlevels entries; any pointer may be NULL. */
#define ORBX_PLANE_BORDER 19
#define ORBX_PLANE_PADL 32
ORBX_API int orbx_get_pyramid_layout(OrbxHandle* h, int width, int height, size_t* frame_bytes, size_t* plane_offset, int32_t* pitch,
                            int32_t* level_w, int32_t* level_h);
/* One copy of resident frame `frame`'s whole block into host_dst (capacity >= *frame_bytes; pinned memory for full speed). */
ORBX_API int orbx_download_pyramid(OrbxHandle* h, int frame, uint8_t* host_dst, size_t capacity);
/* Sink for the pyramids of every frame of the following orbx_extract* calls: frame f's block is copied to
 * host_dst + f * frame_stride inside the call's own pipeline (the reference leaves mvImagePyramid on the host after every
 * operator() call).  NULL switches it off (the default: keypoints and descriptors only). */
ORBX_API int orbx_set_pyramid_output(OrbxHandle* h, uint8_t* host_dst, size_t frame_stride);
/* Pinned (page-locked, portable) host memory for frame / result buffers; write_combined != 0 for buffers the CPU only writes. */
ORBX_API void* orbx_host_alloc(size_t bytes, int write_combined);
ORBX_API void orbx_host_free(void* p);
/* allKeypoints[level] in level coordinates with angle set (== allLevelsKeypoints, ORBextractor.cc:1094). */
ORBX_API int orbx_get_level_keypoints(OrbxHandle* h, int frame, int level, OrbxKeyPoint* kps, int capacity, int* n_out);
/* The same for every level in one call and one read-back: kps receives the levels back to back (level l contributes counts[l]
 * entries), *n_total their sum; capacity >= orbx_max_keypoints() always suffices. */
ORBX_API int orbx_get_all_level_keypoints(OrbxHandle* h, int frame, OrbxKeyPoint* kps, int capacity, int32_t* counts, int* n_total);
/* FAST candidates of the cell loop (ORBextractor.cc:855-860): x, y relative to (16,16), score, and the
 * emission-order key (cell row, cell col, y, x).  Storage order is unspecified; sort by `order`. */
ORBX_API int orbx_get_level_candidates(OrbxHandle* h, int frame, int level, int32_t* xs, int32_t* ys, int32_t* scores,
                              uint32_t* order, int capacity, int* n_out);
/* The 7x7 sigma-2 blurred level the descriptors were sampled from (ORBextractor.cc:1126-1127). */
ORBX_API int orbx_get_blurred_level(OrbxHandle* h, int frame, int level, uint8_t* dst, size_t dst_stride);

/* ---- next row of the path --------------------------------------------------
 * Frame::ComputeStereoMatches (reference src/Frame.cc:813-990): for every left keypoint, Hamming search among the
 * right keypoints of its row band (ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:2349; TH_HIGH 100, TH_LOW 50),
 * 11x11 SAD refinement on the two pyramids, parabola sub-pixel fit, median-based outlier rejection.
 * `left` / `right` are the handles that just extracted the two images (their pyramids stay resident on the GPU, so
 * mvImagePyramid never travels); keys/desc are the host arrays those extractions returned; mb / mbf as in Frame.
 * u_right / depth (n_l floats each) receive mvuRight / mvDepth (-1 where unmatched); *n_matched the kept matches. */
ORBX_API int orbx_stereo_match(OrbxHandle* left, OrbxHandle* right, const OrbxKeyPoint* keys_l, const uint8_t* desc_l, int n_l,
                               const OrbxKeyPoint* keys_r, const uint8_t* desc_r, int n_r, float mb, float mbf,
                               float* u_right, float* depth, int* n_matched);

/* The same row for all stereo pairs of one launch group, device-resident end to end: `left` / `right` just ran
 * orbx_extract_batch(..., ORBX_MEM_DEVICE) on n_pairs left / right images as ONE launch group (n_pairs <= max_batch), kps_* /
 * desc_* / counts_* are the device arrays those calls wrote (cap_per_frame entries per frame), u_right / depth (cap_per_frame
 * floats per pair) and n_matched (one int per pair) are device arrays too.  Entries of u_right / depth beyond a pair's left
 * keypoint count are not written.  Asynchronous on `stream` (NULL: left's own stream); nothing travels to the host. */
ORBX_API int orbx_stereo_match_batch(OrbxHandle* left, OrbxHandle* right, int n_pairs, const OrbxKeyPoint* kps_l, const uint8_t* desc_l,
                                     const int32_t* counts_l, const OrbxKeyPoint* kps_r, const uint8_t* desc_r, const int32_t* counts_r,
                                     int cap_per_frame, float mb, float mbf, float* u_right, float* depth, int32_t* n_matched, void* stream);

/* ---- next rows: what every Frame constructor does right after ExtractORB,
 * and the matcher that consumes it during monocular initialisation -------------------------------------------- */
#define ORBX_FRAME_GRID_COLS 64 /* FRAME_GRID_COLS, reference inc/Frame.h:40 */
#define ORBX_FRAME_GRID_ROWS 48 /* FRAME_GRID_ROWS, inc/Frame.h:39 */

typedef struct OrbxFrameCalib {
    float fx, fy, cx, cy;             /* mK (Pinhole::toK()) */
    float dist[5];                    /* mDistCoef: k1 k2 p1 p2 [k3]; dist[0] == 0 means "already undistorted" (src/Frame.cc:750) */
    int32_t n_dist;                   /* 4 or 5 */
    float min_x, max_x, min_y, max_y; /* mnMinX, mnMaxX, mnMinY, mnMaxY (Frame::ComputeImageBounds) */
} OrbxFrameCalib;

/* Frame::ComputeImageBounds (src/Frame.cc:784-812): fills min_x .. max_y of *calib for a width x height image
 * (the four image corners undistorted on the GPU with the arithmetic of cv::undistortPoints). */
ORBX_API int orbx_frame_image_bounds(OrbxHandle* h, OrbxFrameCalib* calib, int width, int height);
/* Frame::UndistortKeyPoints (src/Frame.cc:748-782) + Frame::AssignFeaturesToGrid / PosInGrid (:383-417, :726-736),
 * mono case.  keys -> keys_un (n records); mGrid[ix][iy] = cell_items[cell_start[ix*48+iy] .. cell_start[ix*48+iy+1])
 * with ascending keypoint index inside a cell; cell_start has 64*48+1 entries, cell_items room for n.
 * *n_in_grid = keypoints that fell inside the grid. */
ORBX_API int orbx_frame_undistort_grid(OrbxHandle* h, const OrbxFrameCalib* calib, const OrbxKeyPoint* keys, int n,
                                       OrbxKeyPoint* keys_un, int32_t* cell_start, int32_t* cell_items, int* n_in_grid);
/* The monocular Frame constructor's use of one image in one call (src/Frame.cc:307-347): ExtractORB, UndistortKeyPoints,
 * AssignFeaturesToGrid.  Same outputs as orbx_extract followed by orbx_frame_undistort_grid, but the keypoints stay on the
 * device in between.  kps / desc / kps_un / cell_items need room for `capacity` entries (orbx_max_keypoints). */
ORBX_API int orbx_extract_frame(OrbxHandle* h, const uint8_t* image, int width, int height, size_t stride, int lap0, int lap1,
                                const OrbxFrameCalib* calib, OrbxKeyPoint* kps, uint8_t* desc, int capacity, int* n_out, int* mono_out,
                                OrbxKeyPoint* kps_un, int32_t* cell_start, int32_t* cell_items, int* n_in_grid);
/* ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:705-814) with GetFeaturesInArea (src/Frame.cc:655-724),
 * DescriptorDistance (:2349-2365) and ComputeThreeMaxima (:2303-2344).  keys_un*, desc*: mvKeysUn / mDescriptors of
 * the two frames; cell_start2 / cell_items2: frame 2's grid from orbx_frame_undistort_grid; prev_matched: n1 (x, y)
 * pairs, vbPrevMatched, updated in place; matches12: n1 ints (vnMatches12, -1 = none); *n_matches = return value.
 * nn_ratio / check_orientation are the ORBmatcher constructor arguments (:40).  n2 <= 32768. */
ORBX_API int orbx_search_for_initialization(OrbxHandle* h, const OrbxFrameCalib* calib, const OrbxKeyPoint* keys_un1,
                                            const uint8_t* desc1, int n1, const OrbxKeyPoint* keys_un2, const uint8_t* desc2, int n2,
                                            const int32_t* cell_start2, const int32_t* cell_items2, float* prev_matched,
                                            int window_size, float nn_ratio, int check_orientation, int32_t* matches12,
                                            int* n_matches);

/* The same call with every array (keypoints, descriptors, grid, prev_matched, matches12) in `mem` memory (ORBX_MEM_HOST or
 * ORBX_MEM_DEVICE): device-resident frames -- e.g. what a fused extraction left on the GPU -- are matched where they lie,
 * and matches12 / prev_matched stay on the device; only *n_matches (host) is read back. */
ORBX_API int orbx_search_for_initialization_mem(OrbxHandle* h, const OrbxFrameCalib* calib, const OrbxKeyPoint* keys_un1,
                                                const uint8_t* desc1, int n1, const OrbxKeyPoint* keys_un2, const uint8_t* desc2, int n2,
                                                const int32_t* cell_start2, const int32_t* cell_items2, float* prev_matched,
                                                int window_size, float nn_ratio, int check_orientation, int32_t* matches12,
                                                int* n_matches, int mem);

/* ----  the pre-processing step of the reference's demos: cv::createCLAHE(clip_limit,
 * Size(tiles_x, tiles_y))->apply(image, out) for CV_8UC1 (src/orb_extractor/main_orb_extractor.cpp:19-22,
 * src/clahe/main_clahe.cpp:7-11), on n_frames independent frames.  images / out live in in_mem / out_mem memory
 * (ORBX_MEM_*); strides in bytes; the two buffers must not overlap.  tiles_x <= 64, width <= 16384.  `stream`: the
 * cudaStream_t to launch on (NULL -> the handle's own); the call returns after the results have landed when either side is
 * host memory, and is asynchronous on `stream` otherwise. */
ORBX_API int orbx_clahe(OrbxHandle* h, const uint8_t* images, int in_mem, int n_frames, int width, int height, size_t row_stride,
                        size_t frame_stride, double clip_limit, int tiles_x, int tiles_y, uint8_t* out, int out_mem,
                        size_t out_row_stride, size_t out_frame_stride, void* stream);

/* Diagnostic: how many keypoints of the last orbx_search_for_initialization call had to be re-enumerated because their
 * 8-entry shortlist was exhausted by already-matched candidates (the slow, still exact, path). */
ORBX_API int orbx_last_init_fallbacks(const OrbxHandle* h);

/* Sum of CUDA-event milliseconds per stage and number of kernel launches since the last call
 * (requires ORBX_FLAG_PROFILE; resets the accumulators). */
ORBX_API int orbx_stage_times(OrbxHandle* h, float* ms_per_stage, int64_t* launches);
/* Total kernel launches issued by this handle since creation. */
ORBX_API int64_t orbx_launch_count(const OrbxHandle* h);
ORBX_API int orbx_synchronize(OrbxHandle* h);
/* 1 if the current workspace stages its FAST tiles with TMA (cp.async.bulk.tensor through per-level tensor maps), 0 if it
 * fell back to cp.async (tensor maps unavailable, or a cell size whose tile row exceeds the 256-byte TMA box). */
ORBX_API int orbx_uses_tma(const OrbxHandle* h);
/* The handle's cudaStream_t (for callers that time with their own events). */
ORBX_API void* orbx_get_stream(const OrbxHandle* h);

#ifdef __cplusplus
}
#endif
#endif /* ORBX_H_ */
