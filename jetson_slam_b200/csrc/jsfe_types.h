// jsfe_types.h -- plain structs shared by the host side and the kernels of libjsfe.so.
#ifndef JSFE_TYPES_H
#define JSFE_TYPES_H

#include <cuda.h>
#include <stdint.h>

#define JSFE_MAXL 16
#define JSFE_B 20               // BORDER_SKIP (reference include/cuda/orb_gpu.hpp:17)
#define JSFE_PATCH_R 21         // 18 (max rotated rBRIEF offset) + 3 (blur radius)
#define JSFE_PATCH_ROWS 43      // 2*21+1
#define JSFE_PATCH_PITCH 48     // bytes per staged patch row (12 aligned words)

namespace jsfe {

// Per-level geometry + device pointers (slot 0; slot s = base + s * slot_stride).
struct LevelGeom {
    int h, w, pitch;                 // pitch: bytes per row, multiple of 16; bytes [w, pitch) are zero
    int tile_h, tile_w, n_tile_h, n_tile_w;
    int cell_offset;                 // level_offset_[l] in the reference
    int tile_row_offset;             // running sum of n_tile_h (index into row_start)
    int T;                           // y-lanes of the reference's NMS launch (tie-break rule)
    int cells_per_block;             // NMS cells one k_fast_cells block owns (adjacent in a tile row)
    int tile_pw;                     // shared-memory pixel-tile pitch of k_fast_cells = TMA box width (JSFE_FAST_PW for every level)
    int blocks_per_row;              // ceil(n_tile_w / cells_per_block)
    int block_offset;                // first k_fast_cells work item of this level
    int fast_ngx, fast_nrl;          // k_fast_cells phase A thread grid: 8-pixel column groups x row lanes (ngx*nrl <= 256)
    unsigned fast_ngx_inv;           // floor(65536/ngx)+1: tid/ngx == (tid*inv)>>16 for tid < 256
    int fast_cap;                    // k_fast_cells work-list slots of a block of this level (<= score positions of the block)
    float scale, inv_scale, rscale;  // rscale = 1.0f / inv_scale (what the reference's resize kernel uses)
    unsigned long long slot_stride;  // bytes between consecutive slots of this level's image
    uint8_t* img;                    // level image of slot 0
    uint8_t* blur;                   // 7x7-blurred level image of slot 0 (same pitch/stride; zero outside [B,h-B)x[B,w-B))
    const uint8_t* mask;             // [h][pitch] or nullptr (all pass); shared by all slots
};

// Small read-only tables in global memory (one copy per handle).
struct DevTables {
    uint32_t lut_bits[2048];   // FAST arc LUT, 1 bit per 16-bit ring mask; bit 0xFFFF = 0
    uint32_t lut_perm[2048];   // the same LUT indexed by k_fast_cells' merged flag word: ring point 4j+k is index bit perm(j,k)
                               // (byte k of ring word j): k=0 -> bits 3..0, k=1 -> 7..4, k=2 -> 11..8, k=3 -> 15..12, j=0 highest
    int umax[16];              // radius-15 disc half-widths
    float gauss[49];           // 7x7 sigma=10 weights, row-major
    float sep_a[7], sep_b[7];  // separable factors: gauss[i*7+j] ~= sep_a[i]*sep_b[j] (|err| < 4e-9)
    int8_t pat_x[512], pat_y[512];
    float2 pat_f[512];  // the same offsets as floats (x, y), transposed: entry [j*32 + b] = sample j (0..15) of descriptor byte b
    uint8_t col_rank[JSFE_MAXL][128];     // column priority of the reference's smem tree (0 wins ties)
    uint8_t col_by_rank[JSFE_MAXL][128];  // inverse permutation
    uint16_t colkey[JSFE_MAXL][192];      // k_fast_cells: per owned column (127 - priority rank) << 8 | cell index in the block
    uint16_t rowkey[JSFE_MAXL][256];      // per owned row dy: (7 - dy % T) << 8 | (255 - dy)
};

struct Params {
    int L;
    int cap;        // max keypoints per slot (= number of NMS cells over all levels)
    int n_tile_rows;  // sum of n_tile_h
    int threshold;  // th_FAST_MAX
    const uint32_t* pyr_map;         // per k_pyramid block: level << 28 | tile row << 14 | tile column
    const uint32_t* fast_map;        // per k_fast_cells work item: level << 28 | tile row << 14 | block index in the row
    unsigned long long vmax_packed;  // nibble |u| (0..15) = largest |v| of the radius-15 disc whose row contains column u
    int use_tma;       // 1: tiles/windows are staged by TMA (cp.async.bulk.tensor), 0: by the threads (JSFE_NO_TMA=1)
    int compass_mode;  // k_fast_cells pre-test: adjacent compass points every accepted arc must cover (0..3)
    int fast_q_thresh; // k_fast_cells phase A: m = (t-3)/4 + 1 (0 for t < 3): p - v > t implies (p>>2) - (v>>2) >= m
    int H0, W0;
    int pyr_blocks_total;             // k_pyramid blocks (128x32 pixel tiles) over levels 1..L-1
    int pyr_block_start[JSFE_MAXL + 1];
    int fast_items_total;             // k_fast_cells work items over all levels
    int blur_items_total;             // k_blur threads (4 columns x 32 rows each) over all levels
    int blur_item_start[JSFE_MAXL + 1];
    LevelGeom lv[JSFE_MAXL];
    const DevTables* tab;
    // per-slot arrays, slot stride = cap unless noted
    int *cell_x, *cell_y, *cell_s;                    // per NMS cell candidates
    int *kp_x, *kp_y, *kp_s, *kp_l;                   // compacted, level coordinates, final order
    float* kp_angle;                                  // radians
    int* n_kp;                                        // [slot]
    int* n_per_level;                                 // [slot][JSFE_MAXL]
    int* row_start;                                   // [slot][n_tile_rows + 1] first keypoint index of each (level, tile row)
    int* kps;                                         // [slot][6][cap] output planes
    uint8_t* desc;                                    // [slot][cap][32]
    float *u_right, *depth;                           // [slot][cap]
    int *best_idx, *best_dist;                        // [slot][cap]
    int* sad_best;                                    // [slot][cap]: accepted SAD minimum or -1
    // cross-scale NMS scratch (allocated only when apply_nms_ms): hash table of occupied level-0 pixels
    int ms_table_size;                                // power of two >= 2*cap
    int *ms_keys, *ms_sums, *ms_cnts;                 // [slot][ms_table_size]
    uint8_t* ms_drop;                                 // [slot][cap]
    unsigned blur_amb_units;                          // JSFE_BLUR_AMB_UNITS (18 x 2^-15); see DESIGN.md 4.2
};

// TMA descriptors (cuTensorMapEncodeTiled, 3-D u8 tensors {pitch, h, slots}), passed as one __grid_constant__ parameter.
//   tile[l]: box {JSFE_FAST_PW, tile_h+8, 1} of level l     -> k_fast_cells pixel tile (+4 px halo), out-of-image = 0
//   disc[l]: box {48, 31, 1} of level l                      -> k_orient_desc intensity-centroid disc
//   win[l] : box {64, 37, 1} of the BLURRED level l          -> k_orient_desc rBRIEF sample window
// Box origins must be 16-byte aligned in x (u8 elements): an unaligned innermost coordinate traps (measured on B200).
struct TmaMaps {
    CUtensorMap tile[JSFE_MAXL];
    CUtensorMap disc[JSFE_MAXL];
    CUtensorMap win[JSFE_MAXL];
};

// The right eye of a stereo match: normally slots 2p+1 of the same handle (left = 2p), or a slot of ANOTHER handle
// with identical geometry (the reference keeps one extractor per eye).
struct RightSide {
    const int* kps;          // [slot][6][cap]
    const uint8_t* desc;     // [slot][cap][32]
    const int* row_start;    // [slot][n_tile_rows + 1]
    const uint8_t* img[JSFE_MAXL];
    int left_mul, left_add, right_mul, right_add;   // slot = mul * pair + add
};

}  // namespace jsfe

#endif
