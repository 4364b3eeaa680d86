// extractorb_b200/csrc/orbx_plan.h -- geometry ("plan") and workspace descriptors shared by host and device.
//
// A Plan is everything that depends only on (image size, constructor arguments): level sizes, plane
// layout in HBM, resize tables, the FAST cell grid and the quadtree quotas.  It is built once on the
// host (orbx_api.cu: build_plan) and passed to every kernel by value as a __grid_constant__ parameter.
#ifndef ORBX_PLAN_H_
#define ORBX_PLAN_H_

#include <stddef.h>
#include <stdint.h>

#define ORBX_MAX_LEVELS 16
#define ORBX_EDGE 19          // EDGE_THRESHOLD, reference ORBextractor.cc:72
#define ORBX_PADL 32          // byte column of ROI x=0 inside a plane row (>= 19, 32-byte aligned rows)
#define ORBX_HALF_PATCH 15    // HALF_PATCH_SIZE, :71
#define ORBX_FAST_BORDER 16   // EDGE_THRESHOLD-3, :781
#define ORBX_COORD_BITS 12    // candidate x/y packing: x | y<<12 | score<<24
#define ORBX_MAX_LEVEL_DIM (4095 + 2 * ORBX_FAST_BORDER)
#define ORBX_ORD_CELL_SHIFT 14  // emission-order key: cellseq<<14 | ycell<<7 | xcell
#define ORBX_MAX_CELL_DIM 127

struct OrbxLevel {
    int w, h;              // level (ROI) size, reference ORBextractor.cc:1171
    int pitch;             // plane row pitch in bytes (multiple of 64)
    int plane_rows;        // h + 38
    long long plane_off;   // byte offset of the plane (its top border row) inside a frame's pyramid block
    int blur_pitch;        // blurred plane row pitch (multiple of 16), no border
    long long blur_off;    // byte offset inside a frame's blur block
    int xtab_off, ytab_off;  // offsets into the resize tables (entries), level >= 1
    // FAST cell grid, reference :781-814
    int nCols, nRows, wCell, hCell;
    int cell_off, ncells;  // slice of the cell table
    int tile_off, ntiles;  // slice of the FAST tile table (k_fast_tiles)
    // quadtree, reference :548-568
    int N;                 // mnFeaturesPerLevel[level]
    int nIni;
    float hX;
    int span_y;            // maxY - minY
    int cand_cap;          // candidate capacity (entries)
    long long cand_off;    // entry offset inside a frame's candidate block
    int kp_cap;            // kept-keypoint capacity: max(N + 2, 4 * nIni) + slack
    int kp_off;            // record offset inside a frame's keypoint staging block
    float sf;              // mvScaleFactor[level]
    float kp_size;         // (float)(int)(31 * sf), :872
};

struct OrbxPlan {
    int nlevels;
    int width, height;
    int ini_th, min_th;
    int ncells_total;
    int kp_total;          // sum of kp_cap
    int qt_nc;             // node capacity of the quadtree kernel (max over levels)
    int qt_sk;             // sort-key slots of the careful phase (qt_nc rounded up to a power of two)
    long long qt_bytes;    // node storage of one quadtree (orbx_qt_bytes(qt_nc))
    int fast_tp;           // FAST smem tile pitch (bytes)
    int fast_trows;        // FAST smem tile rows
    int fast_qcap;         // FAST smem queue capacity (entries)
    int ntiles_total;      // FAST tiles of all levels (k_fast_tiles)
    int ft_tp;             // k_fast_tiles: smem tile pitch (bytes, multiple of 16), rows, queue capacity (entries)
    int ft_trows;
    int ft_qcap;
    int ft_scap;           // NMS survivor list capacity (entries)
    unsigned ft_tpmagic;   // 2^32 / ft_tp + 1: tile byte index -> row (__umulhi)
    int lap0, lap1;
    int umax[ORBX_HALF_PATCH + 1];
    OrbxLevel lv[ORBX_MAX_LEVELS];
};

// One FAST cell (reference ORBextractor.cc:797-814): the cell image is [x0, x0+cw) x [y0, y0+ch) in
// level coordinates, FAST evaluates its pixels >= 3 px inside; xoff/yoff = (j*wCell, i*hCell).
struct OrbxCell {          // 32 bytes, read as two uint4
    uint16_t x0, y0;
    uint8_t cw, ch;
    uint8_t level, pad;
    uint32_t ordbase;      // (i * nCols + j) << ORBX_ORD_CELL_SHIFT
    uint16_t xoff, yoff;
    // precomputed by the host for k_fast_cells (a = (ORBX_PADL + x0) & 3 is the tile's byte alignment)
    uint32_t wmagic;       // (1<<20)/nwords + 1, nwords = (a + cw + 3) >> 2: staged words per tile row
    uint32_t gmagic;       // (1<<24)/ngrp + 1,  ngrp = 4-pixel groups per interior row
    uint8_t nwords, ngrp, wq0, masks;   // wq0 = first group's word; masks = first_mask | last_mask << 4
    uint32_t reserved;
};

static_assert(sizeof(OrbxCell) == 32, "OrbxCell is read as two uint4");

// One FAST tile of k_fast_tiles: `ncells` consecutive cells of one cell row (reference ORBextractor.cc:797-814), processed by
// one CTA as a single image [x0, x0+tw) x [y0, y0+th) in level coordinates.  The cells' detection interiors tile the columns
// [3, tw-3) without gaps: cell c owns interior columns [c*wcell, (c+1)*wcell) (the last cell of a row may be narrower).
// In shared memory pixel (r, x) sits at byte r*ft_tp + a16 + x, a16 = (ORBX_PADL + x0) & 15.
#define ORBX_FT_MAXC 7          // cells per tile (7 x 31 px: at most 29 aligned 8-byte pairs per row -> <= 1024 work items)
struct OrbxFastTile {           // 48 bytes, read as three uint4
    uint16_t x0, y0;
    uint16_t tw, th;
    uint8_t level, ncells;
    uint16_t wcell;
    uint32_t ordbase;           // (cell row * nCols + first cell column) << ORBX_ORD_CELL_SHIFT
    uint16_t xoff, yoff;        // first cell column * wCell, cell row * hCell
    uint32_t cmagic;            // 2^32 / wcell + 1: interior column -> cell (__umulhi)
    uint32_t pmagic;            // 2^32 / npairs + 1: work item -> row (__umulhi)
    uint8_t pq0, npairs;        // first 8-byte pair holding an interior pixel, pairs per row
    uint16_t nitems;            // npairs * (th - 6)
    uint32_t first_mask, last_mask;   // interior-pixel flags of the first / last pair, in the flag layout of k_fast_tiles
    uint32_t vmagic;            // 2^32 / (16-byte chunks per tile row) + 1: staging item -> row
    uint32_t reserved;
};
static_assert(sizeof(OrbxFastTile) == 48, "OrbxFastTile is read as three uint4");

// Kept keypoint record written by the quadtree kernel, completed by the describe kernel.
struct OrbxKpRec {
    float x, y;            // level coordinates (ROI), +16 border already added (:875-876)
    float response;
    float angle;
    int lap_before;        // lapping keypoints earlier in this level's list
    int src;               // index of the winning candidate
};

// Device workspace for one launch group of up to `frames` frames.
struct OrbxWs {
    uint8_t* pyr;          long long pyr_stride;    // per-frame byte strides
    uint8_t* blur;         long long blur_stride;
    uint2* cand;           long long cand_stride;   // entries
    uint16_t* keynode;                                // same indexing as cand
    OrbxKpRec* kprec;      int kp_stride;             // records
    int* cand_count;       // [frame][level]
    int2* level_count;     // [frame][level] = {n, n_lapping}
    int* flags;            // one word, bit0: some frame overflowed its candidate workspace
    uint8_t* qt_scratch;   // NULL: quadtree node tables live in shared memory; else qt_bytes per (frame, level) in HBM
    const int2* xtab;      // resize: {src offset, a0 | a1<<16}
    const int2* ytab;
    const OrbxCell* cells;
    const OrbxFastTile* tiles;
    const uint8_t* tmaps;     // one 128-byte CUtensorMap per level over this workspace's planes (k_fast_tiles<true>), or NULL
    const uint8_t* tmaps_blur; // ... over its blurred levels, box 64 x 37 (k_describe<true>), or NULL
    const uint8_t* tmaps_b7;   // ... over its planes, box 160 x 70 (k_blur7<true>), or NULL
    const uint8_t* tmaps_rs;   // ... entry l over the plane of level l-1, box = staging pitch x rows of k_pyr_resize<.., true>, or NULL
    const uint8_t* slot_level; // level of every kept-keypoint slot of a frame (kp_total bytes)
    const float* pattern_f;   // rBRIEF tests as floats, layout [bit k][descriptor byte][x0, x1, y0, y1]
    const int2* angle_w;      // IC_Angle weights [4 alignments][33 rows][9 words] = {u bytes, v bytes} (signed, 0 outside the circle)
    const uint32_t* blur_tiles; // blur tile table: level | tile_x << 8 | tile_y << 20
};

#endif
