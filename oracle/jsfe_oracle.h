/*
 * jsfe_oracle.h -- CPU restatement of the Jetson-SLAM stereo front-end.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product (jetson_slam_b200/,
 * include/jsfe.h, libjsfe.so) may include, link or call this code: it exists so
 * that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg can check (and time beside) the CUDA path.
 *
 * It restates, stage by stage, what the reference's hot path computes
 * (reference = ashishkumar822/Jetson-SLAM, paths relative to /root/reference):
 *   geometry/tables  src/cuda/orb_gpu.cpp:22-441
 *   pyramid          src/cuda/orb_pyramid.cu:18-68
 *   FAST score       src/cuda/orb_FAST_compute_score.cu:1412-1560
 *   cell NMS/argmax  src/cuda/orb_FAST_apply_NMS_G.cu:1178-1483
 *   cross-scale NMS  src/cuda/orb_FAST_apply_NMS_MS.cu:18-467, orb_FAST_apply_NMS_MS.cpp:15-122
 *   compaction       src/cuda/orb_FAST_obtain_keypoints.cpp:12-56
 *   orientation      src/cuda/orb_FAST_orientation.cu:17-65
 *   7x7 blur         src/cuda/orb_gaussian.cu:21-138
 *   rBRIEF           src/cuda/orb_descriptor.cu:12-69 (+ OpenCV bit_pattern_31_)
 *   output packing   src/cuda/orb_copy_output.cu:12-45
 *   stereo match     src/cuda/orb_stereo_match.cu:28-580
 *
 * Parity pin: the reference ships NO tests, fixtures or golden vectors for this
 * path.  This oracle is pinned against the reference's own
 * src/cuda compiled unmodified for sm_100a (oracle/ref_build -> oracle/_ref)
 * and run on a B200; the resulting vectors are committed under tests/golden/ref_*.npz
 * (tools/make_golden.py; report: tests/golden/pin_report_r1_run*.json -- every stage bit-identical).
 * Status: PINNED for the extraction + stereo path.  The three adjacent helper kernels (projection, Hamming pairs,
 * frustum) are checked against the reference kernels on the GPU box by tests/test_helpers.py (no committed fixture).
 * orc_search_by_projection (row f1) restates reference HOST code that cannot be built here: parity UNPINNED for it.
 */
#ifndef JSFE_ORACLE_H
#define JSFE_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 16
#define ORC_BORDER 20 /* BORDER_SKIP, include/cuda/orb_gpu.hpp:17 */

typedef struct orc_config {
    int32_t height, width;
    int32_t n_levels;
    float scale_factor;
    int32_t fast_n_min, fast_n_max;
    int32_t th_fast_min, th_fast_max; /* th_fast_min is ignored, as in the reference (orb_gpu.cpp:42-47) */
    int32_t tile_h, tile_w;
    int32_t fixed_multi_scale_tile_size;
    int32_t apply_nms_ms;    /* effective only when n_levels > 1 (orb_gpu.cpp:37) */
    int32_t nms_ms_mode_gpu; /* 1: dense s0-volume variant (two-phase deterministic), 0: CPU bucket variant */
} orc_config;

typedef struct orc_ctx orc_ctx;

/* mask: NULL (all 255) or height*width bytes (level-0 mask; levels are NN-resized + thresholded >10). */
orc_ctx* orc_create(const orc_config* cfg, const uint8_t* mask);
void orc_destroy(orc_ctx* c);

int orc_max_kp(const orc_ctx* c);
int orc_n_levels(const orc_ctx* c);
/* per level: h,w,tile_h,tile_w,n_tile_h,n_tile_w,level_offset  (7 ints per level) */
void orc_level_geometry(const orc_ctx* c, int32_t* out7);
void orc_scales(const orc_ctx* c, float* scale, float* inv_scale);
const uint8_t* orc_lut(const orc_ctx* c);       /* 65536 bytes, [0xFFFF] = 0 */
const int32_t* orc_umax(const orc_ctx* c);      /* 16 */
const float* orc_gauss(const orc_ctx* c);       /* 49 */
const int8_t* orc_pattern_x(const orc_ctx* c);  /* 512 */
const int8_t* orc_pattern_y(const orc_ctx* c);  /* 512 */
/* column priority permutation of the NMS-G smem tree for a given tile_w: rank[j] (0 = wins ties) */
void orc_column_rank(int tile_w, int32_t* rank);

/*
 * Full extraction of one eye.  image: height*width contiguous u8.
 * kps_soa: 6*N int32 planes [x|y|score|angle(f32 bits)|octave|size] laid out with stride N (= return value),
 * exactly as ORB_GPU::extract leaves `out_keypoints` (orb_gpu.cpp:784-816).  Caller provides 6*max_kp ints.
 * desc: 32*N bytes.  Returns N (total keypoints).
 */
int orc_extract(orc_ctx* c, const uint8_t* image, int32_t* kps_soa, uint8_t* desc);

/* stage accessors (valid after orc_extract) */
const uint8_t* orc_level_image(const orc_ctx* c, int level);
const uint8_t* orc_level_blur(const orc_ctx* c, int level);
const int32_t* orc_level_score(const orc_ctx* c, int level);
/* per-cell candidates before compaction (max_kp each, level_offset-indexed) */
const int32_t* orc_cell_x(const orc_ctx* c);
const int32_t* orc_cell_y(const orc_ctx* c);
const int32_t* orc_cell_score(const orc_ctx* c);
/* compacted level-coordinate keypoints: n per level, arrays indexed level_offset[l]+j */
const int32_t* orc_n_keypoints(const orc_ctx* c);
const int32_t* orc_kp_x(const orc_ctx* c);
const int32_t* orc_kp_y(const orc_ctx* c);
const int32_t* orc_kp_score(const orc_ctx* c);
const float* orc_kp_angle(const orc_ctx* c);

/* individual stages, exposed for unit tests */
void orc_stage_pyramid(orc_ctx* c, const uint8_t* image);
void orc_stage_fast(orc_ctx* c);
void orc_stage_cells(orc_ctx* c);
void orc_stage_nms_ms(orc_ctx* c);
void orc_stage_compact(orc_ctx* c);
void orc_stage_orient(orc_ctx* c);
void orc_stage_blur(orc_ctx* c);
void orc_stage_describe(orc_ctx* c);

/* libdevice transcriptions (CUDA 12.9 libdevice, fast path |x| < 105615) */
float orc_atan2f(float y, float x);
float orc_cosf(float x);
float orc_sinf(float x);

/*
 * Stereo match (orb_stereo_match.cu:105-580).  cl / cr hold the left / right pyramids of the last
 * orc_extract.  kps_* are the 6-plane SoAs with stride n_*.  Outputs u_right[n_left], depth[n_left]
 * (-1 = no match).  Optional debug outputs (may be NULL): best_idx_r[n_left] (Hamming arg-min right
 * index or -1), best_dist[n_left] (Hamming distance or th_high).
 * Returns the number of left keypoints that keep a depth.
 */
int orc_stereo_match(const orc_ctx* cl, const orc_ctx* cr, int th_high, int th_low, float mb, float mbf,
                     int n_left, const int32_t* kps_left, const uint8_t* desc_left,
                     int n_right, const int32_t* kps_right, const uint8_t* desc_right,
                     float* u_right, float* depth, int32_t* best_idx_r, int32_t* best_dist);

/* Convenience for the CPU baseline: extract both eyes (optionally on 2 threads) + match. Returns n_left. */
int orc_stereo_pair(orc_ctx* cl, orc_ctx* cr, const uint8_t* img_l, const uint8_t* img_r,
                    float mb, float mbf, int32_t* kps_l, uint8_t* desc_l, int32_t* n_r_out,
                    int32_t* kps_r, uint8_t* desc_r, float* u_right, float* depth, int threads);

/* ---- adjacent rows: helpers of ORBmatcher::SearchByProjection and Tracking::SearchLocalPoints ---- */
float orc_logf(float x); /* libdevice logf transcription */
/* src/cuda/orb_matcher.cu:17-64: Pc = R*Pw + t, pinhole projection, bounds check */
void orc_project_points(int n, const float* px, const float* py, const float* pz, const float* rcw9, const float* tcw3,
                        float fx, float fy, float cx, float cy, float min_x, float max_x, float min_y, float max_y,
                        float* u, float* v, float* invz, uint8_t* is_valid);
/* src/cuda/orb_matcher.cu:95-118: 256-bit Hamming distance of indexed descriptor pairs */
void orc_hamming_pairs(int n, const int32_t* idx_l, const int32_t* idx_r, const uint8_t* desc_l, const uint8_t* desc_r,
                       int32_t* dist);
/* src/cuda/tracking_isinfrustum.cu:19-107: Frame::isInFrustum per map point.  Outputs other than is_infrustum are
 * written only for points inside the frustum (as in the reference). */
void orc_in_frustum(int n, const float* px, const float* py, const float* pz, const float* pnx, const float* pny,
                    const float* pnz, const float* max_distance, const float* inv_max_distance,
                    const float* inv_min_distance, const float* rcw9, const float* tcw3, const float* ow3, float fx,
                    float fy, float cx, float cy, int min_x, int max_x, int min_y, int max_y, int n_scale_levels,
                    float log_scale_factor, float view_cos_angle, float* invz, float* u, float* v,
                    int32_t* predicted_level, float* view_cos, uint8_t* is_infrustum);

/* ---- the live branch of ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)
 * (src/ORBmatcher.cpp:1647-1963) with Frame::AssignFeaturesToGrid / GetFeaturesInArea (src/Frame.cpp:464-479, 569-639) and
 * ComputeThreeMaxima (src/ORBmatcher.cpp:2097-2138).  Restatement only (the host code around the two pinned kernels cannot
 * be compiled here): parity UNPINNED for the host loops. */
void orc_assign_features_to_grid(int n, const float* x, const float* y, float min_x, float min_y, float winv, float hinv,
                                 int32_t* cell_start /* 64*48+1 */, int32_t* cell_items /* n */);
int orc_search_by_projection(int n_last, const float* px, const float* py, const float* pz, const int32_t* last_octave,
                             const float* last_angle, const uint8_t* last_desc, const float* rcw9, const float* tcw3,
                             float fx, float fy, float cx, float cy, float min_x, float max_x, float min_y, float max_y,
                             float mbf, float th, const float* scale_factors, int level_mode, int n_cur,
                             const float* cur_x, const float* cur_y, const int32_t* cur_octave, const float* cur_angle,
                             const float* cur_uright, const uint8_t* cur_occupied, const uint8_t* cur_desc, int th_high,
                             int check_orientation, int32_t* best_idx2, int32_t* best_dist, int32_t* rot_bin,
                             int32_t* cur_match, int32_t* hist /* 30 */);

/* ---- cv::remap(INTER_LINEAR, CV_32FC1 maps, BORDER_CONSTANT 0) and cv::cvtColor(*2GRAY) for 8-bit images
 * (Examples/Stereo/stereo_euroc.cpp:106-107,145-146; src/Tracking.cpp:260-285).  OpenCV's published algorithm, PINNED
 * bit-exactly against the cv2 4.13 wheel of this image. */
void orc_remap_bilinear_u8(const uint8_t* src, int src_h, int src_w, int64_t src_pitch, const float* map_x, const float* map_y,
                           int dst_h, int dst_w, uint8_t* dst, int64_t dst_pitch);
void orc_cvt_gray_u8(const uint8_t* src, int64_t n_pixels, int channels, int blue_idx, uint8_t* dst);

#ifdef __cplusplus
}
#endif
#endif
