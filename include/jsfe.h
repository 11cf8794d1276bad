/*
 * jsfe.h -- C ABI of the B200-native stereo front-end (libjsfe.so).
 *
 * This is the drop-in boundary for the reference's per-frame hot path.  The reference
 * (ashishkumar822/Jetson-SLAM, paths relative to its root) has no plugin/FFI layer: its SLAM core
 * includes five C++ entities directly.  Each entry point below names the reference interface it
 * replaces; compat/ re-implements those C++ headers as thin shims over this ABI.
 *
 *   extern "C", plain pointers and sizes, no C++/torch types.  Every call returns 0 (JSFE_OK) or a
 *   negative jsfe_status; jsfe_last_error() gives a thread-local message.  There is NO CPU fallback:
 *   without a CUDA device jsfe_create fails with JSFE_ERR_CUDA.
 *
 * Model: a handle owns `max_images` image SLOTS on one GPU.  A slot is one eye: its level-0 image,
 * its pyramid, its keypoints/descriptors.  Stereo pair p uses slot 2p (left) and 2p+1 (right).
 * The reference's "one ORB_GPU per eye" maps to a handle with max_images = 1 (compat/) and the
 * batched B200 path to one handle with max_images = 2 * pairs.  All device memory of the extraction /
 * matching path is allocated in jsfe_create (the upload staging buffer and the three streams of
 * jsfe_process_host_pairs are created at its first use); nothing is allocated per frame.  A handle is single-threaded; distinct
 * handles may be driven concurrently from distinct host threads (the reference does this for L/R,
 * src/Frame.cpp:107-110).  Work is enqueued on the caller's CUDA stream (cudaStream_t passed as
 * void*; NULL = the legacy default stream); only the jsfe_get_* / jsfe_download_* calls synchronise.
 */
#ifndef JSFE_H
#define JSFE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JSFE_MAX_LEVELS 16
#define JSFE_BORDER 20 /* BORDER_SKIP, include/cuda/orb_gpu.hpp:17 */

typedef enum jsfe_status {
    JSFE_OK = 0,
    JSFE_ERR_INVALID = -1, /* bad argument / configuration                 */
    JSFE_ERR_CUDA = -2,    /* CUDA runtime error (message has the details) */
    JSFE_ERR_CAPACITY = -3 /* more images/pairs than the handle's slots    */
} jsfe_status;

/* Mirrors the 15 constructor parameters of Jetson_SLAM::ORBExtractor (include/ORBextractor.h:25-35)
 * = orb_cuda::ORB_GPU (include/cuda/orb_gpu.hpp:26-35), plus device/slot sizing.
 * The reference loads its mask from a file path (src/cuda/orb_gpu.cpp:64-91); here the caller
 * passes the level-0 mask bytes (NULL = all 255); levels are nearest-neighbour resized and
 * thresholded at 10 exactly as the reference does. */
typedef struct jsfe_config {
    int32_t height, width;
    int32_t n_levels;
    float scale_factor;
    int32_t fast_n_min, fast_n_max;
    int32_t th_fast_min, th_fast_max; /* th_fast_min is accepted and ignored, as in the reference (orb_gpu.cpp:42-47) */
    int32_t tile_h, tile_w;
    int32_t fixed_multi_scale_tile_size;
    int32_t apply_nms_ms;    /* effective only for n_levels > 1 (orb_gpu.cpp:37) */
    int32_t nms_ms_mode_gpu; /* 1: dense-volume rule (deterministic two-phase), 0: bucket rule (orb_FAST_apply_NMS_MS.cpp) */
    const uint8_t* mask;     /* host pointer, height x mask_pitch, or NULL */
    int64_t mask_pitch;
    int32_t device_id;
    int32_t max_images; /* image slots; >= 1 */
} jsfe_config;

typedef struct jsfe_handle jsfe_handle;

/* Per-level geometry (orb_gpu.cpp:49-62, 224-258, 305-327). */
typedef struct jsfe_level_info {
    int32_t height, width;
    int32_t pitch; /* bytes per row of the device level image (multiple of 16) */
    int32_t tile_h, tile_w, n_tile_h, n_tile_w;
    int32_t cell_offset; /* level_offset_[i] */
    float scale, inv_scale;
} jsfe_level_info;

/* Device-side view of one slot's results (all pointers are device pointers owned by the handle).
 * kps: 6 planes [x | y | score | angle_deg(f32 bits) | octave | size], plane stride = capacity.
 * Same plane order as ORB_GPU::extract's out_keypoints (orb_gpu.cpp:784-816); the reference packs
 * planes with stride N, jsfe_pack_keypoints produces exactly that layout. */
typedef struct jsfe_slot_view {
    const int32_t* n_keypoints;  /* 1 int                           */
    const int32_t* n_per_level;  /* n_levels ints                   */
    const int32_t* kps;          /* 6 * capacity                    */
    const uint8_t* desc;         /* 32 * capacity                   */
    const float* u_right;        /* capacity (left slots, after jsfe_stereo_match) */
    const float* depth;          /* capacity                        */
    const int32_t* best_idx_r;   /* capacity: Hamming arg-min right index or -1    */
    const int32_t* best_dist;    /* capacity: its distance or TH_HIGH              */
    int32_t capacity;
} jsfe_slot_view;

const char* jsfe_last_error(void);

/* ---- lifetime: replaces ORBExtractor::ORBExtractor / ~ORBExtractor (src/ORBextractor.cpp:27-95)
 *      and ORB_GPU::ORB_GPU / ~ORB_GPU (src/cuda/orb_gpu.cpp:22-453). */
int jsfe_create(const jsfe_config* cfg, jsfe_handle** out);
int jsfe_destroy(jsfe_handle* h);

/* ---- geometry: replaces ORBExtractor::get_levels / get_scale_factors / get_inverse_scale_factors
 *      (include/ORBextractor.h:44-72) and the public ORB_GPU::height_/width_ (orb_gpu.hpp:242-243). */
int jsfe_max_keypoints(const jsfe_handle* h); /* max_kp_count_ (orb_gpu.cpp:321) = per-slot capacity */
int jsfe_num_levels(const jsfe_handle* h);
int jsfe_get_level_info(const jsfe_handle* h, int level, jsfe_level_info* out);

/* ---- input: replaces the cudaMemcpy2D upload at orb_gpu.cpp:497.
 * Copies n images into slots [first_slot, first_slot+n).  src may be host (pinned for async) or
 * device memory; row_pitch / image_stride in bytes.  Asynchronous on `stream`. */
int jsfe_set_images(jsfe_handle* h, int first_slot, int n, const uint8_t* src, int64_t row_pitch,
                    int64_t image_stride, int src_is_device, void* stream);
/* Zero-copy alternative: device pointer + pitch of a slot's level-0 image, for producers that write
 * frames straight into the slot (capture DMA, rectification kernel). */
int jsfe_slot_image(jsfe_handle* h, int slot, uint8_t** dev_ptr, int64_t* pitch);

/* ---- extraction: replaces ORBExtractor::extract (src/ORBextractor.cpp:97-105) ->
 *      ORB_GPU::extract (src/cuda/orb_gpu.cpp:489-841) for slots [first_slot, first_slot+n):
 *      pyramid, FAST score, per-cell NMS arg-max, optional cross-scale NMS, ordered compaction,
 *      IC angle, 7x7 blur, rBRIEF, output packing.  Asynchronous on `stream`. */
int jsfe_extract(jsfe_handle* h, int first_slot, int n, void* stream);

/* ---- stereo: replaces ORB_GPU::ORB_compute_stereo_match (include/cuda/orb_gpu.hpp:218-229,
 *      src/cuda/orb_stereo_match.cu:105-580) as called by Frame::ComputeStereoMatches
 *      (src/Frame.cpp:780-803) for pairs [first_pair, first_pair+n): left = slot 2p, right = 2p+1.
 *      mb is explicit (the reference reads Frame::mb before assigning it; benchmarks use mbf/fx).
 *      Results land in the LEFT slot's u_right/depth/best_idx_r/best_dist.  Asynchronous. */
int jsfe_stereo_match(jsfe_handle* h, int first_pair, int n, int th_high, int th_low, float mb, float mbf,
                      void* stream);

/* Same, for ONE pair whose eyes live in two handles with identical geometry on the same device (the reference keeps one
 * ORBExtractor per eye and hands both pyramids to the matcher, src/Frame.cpp:784-800): left = (hl, slot_l), right =
 * (hr, slot_r).  Results land in hl's slot_l.  `stream` must be ordered after both extractions. */
int jsfe_stereo_match_cross(jsfe_handle* hl, int slot_l, jsfe_handle* hr, int slot_r, int th_high, int th_low, float mb,
                            float mbf, void* stream);

/* ---- results */
int jsfe_slot_view_get(const jsfe_handle* h, int slot, jsfe_slot_view* out);
/* Level image left on the device by the last extract: replaces the public ORB_GPU::image_
 * (orb_gpu.hpp:247) that Frame::ComputeStereoMatches hands to the matcher. */
int jsfe_level_image(const jsfe_handle* h, int slot, int level, const uint8_t** dev_ptr, int32_t* height,
                     int32_t* width, int64_t* pitch);
/* Pack a slot's keypoints the way the reference returns them (6 planes, stride N; descriptors 32N)
 * into CALLER-owned device buffers (the compat SyncedMem): orb_gpu.cpp:784-831.  *n_out receives N
 * (this call synchronises `stream` to learn N). dst_kps needs 6*capacity ints, dst_desc 32*capacity. */
int jsfe_pack_keypoints(jsfe_handle* h, int slot, int32_t* dst_kps_dev, uint8_t* dst_desc_dev, int32_t* n_out,
                        void* stream);
/* The same packing with ONE synchronisation: the buffers must hold the capacity (6*jsfe_max_keypoints() ints, 32*jsfe_max_keypoints()
 * bytes); the pack kernel reads N on the device, N is copied back behind it and `stream` is synchronised once.  On return the
 * device buffers are complete and *n_out = N (planes with stride N, as the reference returns them). */
int jsfe_pack_keypoints_once(jsfe_handle* h, int slot, int32_t* dst_kps_dev, uint8_t* dst_desc_dev, int32_t* n_out, void* stream);
/* Host copies (synchronise `stream`).  kps_host: 6*N ints with stride N (reference layout), desc_host 32*N;
 * either may be NULL.  Capacity of the host buffers must be jsfe_max_keypoints(). */
int jsfe_get_keypoints(jsfe_handle* h, int slot, int32_t* kps_host, uint8_t* desc_host, int32_t* n_out, void* stream);
/* jsfe_get_stereo reads the LEFT slot 2*pair; jsfe_get_stereo_slot reads an explicit left slot (cross-handle use). */
int jsfe_get_stereo_slot(jsfe_handle* h, int left_slot, float* u_right_host, float* depth_host, int32_t* best_idx_r_host,
                         int32_t* best_dist_host, int32_t* n_left_out, void* stream);
int jsfe_get_stereo(jsfe_handle* h, int pair, float* u_right_host, float* depth_host, int32_t* best_idx_r_host,
                    int32_t* best_dist_host, int32_t* n_left_out, void* stream);

/* Bulk D2H of n slots' result slabs into the handle's pinned staging area (one async copy per
 * array kind, then one synchronise): the per-step readback of the end-to-end path.
 * After it returns, jsfe_host_results gives host pointers into the staging area. */
typedef struct jsfe_host_results {
    const int32_t* n_keypoints; /* [n]                      */
    const int32_t* kps;         /* [n][6][capacity]         */
    const uint8_t* desc;        /* [n][capacity][32]        */
    const float* u_right;       /* [n][capacity] (left slots meaningful) */
    const float* depth;         /* [n][capacity]            */
    int32_t capacity;
    int64_t bytes; /* bytes copied D2H by the last jsfe_download_results */
} jsfe_host_results;
int jsfe_download_results(jsfe_handle* h, int first_slot, int n, jsfe_host_results* out, void* stream);

/* ---- end-to-end batch call with HOST buffers (the call a batch user makes; bench.py's `e2e` times exactly this):
 * replaces, per pair, the whole hot path of Frame::Frame (src/Frame.cpp:103-122,219): image upload, extract x2,
 * stereo match, download of keypoints/descriptors/uRight/depth.  `images` = [L0,R0,L1,R1,...] contiguous u8
 * [2*n_pairs][height][width] (pinned memory recommended).  Internally pipelined in chunks of `chunk_pairs` pairs
 * (<= 0: default) over three CUDA streams: contiguous H2D | re-pitch + kernels | D2H, so transfers hide behind
 * compute.  Synchronous: on return `out` points at the handle's pinned result slabs for slots [0, 2*n_pairs). */
int jsfe_process_host_pairs(jsfe_handle* h, int n_pairs, const uint8_t* images, int chunk_pairs, int th_high, int th_low,
                            float mb, float mbf, jsfe_host_results* out);
/* The same call split at its synchronisation point: _begin enqueues the whole batch and returns, _end waits and hands out the
 * result slabs.  One batch per handle may be in flight; with two handles a single host thread keeps two batches in flight
 * (begin(h1), end(h0), begin(h0), end(h1), ...), so the upload of one batch overlaps the kernels of the other and sustained
 * host-to-host throughput approaches the device-resident rate.  `images` must stay valid (pinned recommended) until _end. */
int jsfe_process_host_pairs_begin(jsfe_handle* h, int n_pairs, const uint8_t* images, int chunk_pairs, int th_high, int th_low,
                                  float mb, float mbf);
int jsfe_process_host_pairs_end(jsfe_handle* h, jsfe_host_results* out);
/* Pinned host memory for the `images` argument above (what SyncedMem's cudaMallocHost does in the reference,
 * src/cuda/synced_mem_holder.cpp:46,65).  write_combined != 0 allocates it cudaHostAllocWriteCombined: the copy engine's reads are
 * then not snooped through the CPU caches, which is worth ~5 % of host-to-host throughput once several GPUs of one socket upload
 * at the same time (bench.py, 4 GPUs: 184 k -> 193 k pairs/s).  Write-combined memory is for buffers the host only WRITES
 * (a capture / decode target); host reads of it are slow.  Free with jsfe_host_free. */
int jsfe_host_alloc(void** ptr, size_t bytes, int write_combined);
int jsfe_host_free(void* ptr);

/* ---- adjacent rows: stateless helpers of the tracking thread.  All pointers are DEVICE pointers,
 * work is enqueued on `stream` and NOT synchronised (the compat shims synchronise, as the reference does).
 * jsfe_project_points replaces orb_cuda::ORB_Search_by_projection_project_on_frame (include/cuda/orb_matcher.hpp:11-17),
 * jsfe_hamming_pairs  replaces orb_cuda::ORB_compute_distances                     (include/cuda/orb_matcher.hpp:19-23),
 * jsfe_in_frustum     replaces tracking_cuda::compute_isInFrustum_GPU              (include/cuda/tracking_gpu.hpp:13-28). */
int jsfe_project_points(int n, const float* px, const float* py, const float* pz, const float* rcw9, const float* tcw3, float fx,
                        float fy, float cx, float cy, float min_x, float max_x, float min_y, float max_y, float* u, float* v,
                        float* invz, uint8_t* is_valid, void* stream);
int jsfe_hamming_pairs(int n, const int32_t* idx_left, const int32_t* idx_right, const uint8_t* desc_left,
                       const uint8_t* desc_right, int32_t* distance, void* stream);
int jsfe_in_frustum(int n, const float* px, const float* py, const float* pz, const float* pnx, const float* pny, const float* pnz,
                    const float* max_distance, const float* invariance_max_distance, const float* invariance_min_distance,
                    const float* rcw9, const float* tcw3, const float* ow3, float fx, float fy, float cx, float cy, int min_x,
                    int max_x, int min_y, int max_y, int n_scale_levels, float log_scale_factor, float view_cos_angle, float* invz,
                    float* u, float* v, int32_t* predicted_level, float* view_cos, uint8_t* is_infrustum, void* stream);

/* ---- SURVEY.md 8(f2), the resident form: Tracking::SearchLocalPoints re-packs nine float arrays of every local map point and uploads
 * them for EVERY frame (src/Tracking.cpp:1486-1561) although map points change far less often than frames arrive.  A jsfe_mappool
 * keeps that SoA on the device under stable slot ids (e.g. MapPoint::mnId modulo the capacity, or an index the caller maintains):
 *   jsfe_mappool_update     uploads n new / moved map points (HOST arrays, what GetWorldPosNormalExp and GetDistanceInvariances return)
 *                           into slots ids[i]: called when LocalMapping changes them;
 *   jsfe_mappool_in_frustum compute_isInFrustum_GPU (include/cuda/tracking_gpu.hpp:13-28) for the n pool points `ids_dev` (DEVICE int32
 *                           list) under the HOST pose: per frame 4 bytes per point and 15 floats cross PCIe instead of 36 bytes per point.
 * Outputs are DEVICE arrays indexed by the position in ids_dev, identical to jsfe_in_frustum on the gathered arrays. */
typedef struct jsfe_mappool jsfe_mappool;
int jsfe_mappool_create(int capacity, int device_id, jsfe_mappool** out);
int jsfe_mappool_destroy(jsfe_mappool* p);
int jsfe_mappool_update(jsfe_mappool* p, int n, const int32_t* ids, const float* px, const float* py, const float* pz, const float* pnx,
                        const float* pny, const float* pnz, const float* max_distance, const float* invariance_max_distance,
                        const float* invariance_min_distance, void* stream);
int jsfe_mappool_in_frustum(jsfe_mappool* p, int n, const int32_t* ids_dev, const float* rcw9_host, const float* tcw3_host,
                            const float* ow3_host, float fx, float fy, float cx, float cy, int min_x, int max_x, int min_y, int max_y,
                            int n_scale_levels, float log_scale_factor, float view_cos_angle, float* invz, float* u, float* v,
                            int32_t* predicted_level, float* view_cos, uint8_t* is_infrustum, void* stream);

/* ---- ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) as ONE device
 * pass (replaces the live branch src/ORBmatcher.cpp:1647-1963: 2 kernels + 3 host loops + 9 copies per call).
 * All pointers are DEVICE pointers; work is enqueued on `stream`, not synchronised.
 *
 * jsfe_build_frame_grid replaces Frame::AssignFeaturesToGrid (src/Frame.cpp:464-479, PosInGrid :696-706): the 64x48 grid of the
 * current frame as CSR -- cell_start[64*48+1], cell = ix*48+iy (mGrid[ix][iy]); cell_items[n_cur] lists keypoint indices,
 * ascending inside a cell (the host's push_back order).  Build once per frame, reuse for every search against it. */
#define JSFE_FRAME_GRID_COLS 64
#define JSFE_FRAME_GRID_ROWS 48
#define JSFE_HISTO_LENGTH 30
int jsfe_build_frame_grid(int n_cur, const float* cur_x, const float* cur_y, float min_x, float max_x, float min_y, float max_y,
                          int32_t* cell_start, int32_t* cell_items, void* stream);

typedef struct jsfe_sbp_args {
    /* last frame: the points the host loop keeps (map point present, not an outlier), in LastFrame order */
    int32_t n_last;
    const float *px, *py, *pz;      /* world positions (MapPoint::GetWorldPos)                        */
    const int32_t* last_octave;     /* LastFrame.mvKeys[i].octave                                      */
    const float* last_angle;        /* LastFrame.mvKeysUn[i].angle (degrees)                           */
    const uint8_t* last_desc;       /* [n_last][32] MapPoint::GetDescriptor                            */
    const float *rcw9, *tcw3;       /* CurrentFrame.mTcw rotation (row-major) and translation          */
    float fx, fy, cx, cy, min_x, max_x, min_y, max_y; /* CurrentFrame intrinsics and image bounds (mnMinX ...) */
    float mbf, th;                  /* CurrentFrame.mbf; search window factor (radius = th * scale_factors[octave]) */
    float scale_factors[16];        /* CurrentFrame.mvScaleFactors                                     */
    int32_t level_mode;             /* 0: levels octave-1..octave+1; 1: bForward (>= octave); 2: bBackward (<= octave) */
    /* current frame */
    int32_t n_cur;                  /* <= 65535                                                        */
    const float *cur_x, *cur_y;     /* mvKeysUn[i].pt                                                  */
    const int32_t* cur_octave;
    const float* cur_angle;         /* degrees                                                         */
    const float* cur_uright;        /* mvuRight (<= 0: monocular keypoint)                             */
    const uint8_t* cur_occupied;    /* != 0 <=> mvpMapPoints[i] && Observations() > 0; NULL = none     */
    const uint8_t* cur_desc;        /* [n_cur][32]                                                     */
    const int32_t *cell_start, *cell_items; /* from jsfe_build_frame_grid                              */
    int32_t th_high;                /* ORBmatcher::TH_HIGH (100); must be < 256                        */
    int32_t check_orientation;      /* mbCheckOrientation                                              */
    /* outputs */
    int32_t* best_idx2;             /* [n_last] matched current keypoint, -1 = none (before the rotation cull) */
    int32_t* best_dist;             /* [n_last] its Hamming distance, 256 = none                       */
    int32_t* rot_bin;               /* [n_last] rotation-histogram bin, -1 = none                      */
    int32_t* cur_match;             /* [n_cur] last-frame point whose map point the keypoint ends up with, -1 = none */
    int32_t* hist;                  /* [30] rotation histogram sizes                                   */
    int32_t* n_matches;             /* [1] the function's return value                                 */
} jsfe_sbp_args;
int jsfe_search_by_projection(const jsfe_sbp_args* args, void* stream);

/* ---- the output side on the device.  Frame::Frame unpacks the 6-plane SoA into
 * std::vector<cv::KeyPoint> on the host, one field at a time (src/Frame.cpp:116-196: pt.x/pt.y = int -> float, response =
 * score, angle in degrees, octave, size).  jsfe_frame_view does that unpack on the device for slot `slot`:
 *   keys  : [capacity] records with cv::KeyPoint's memory layout (pt.x, pt.y, size, angle, response, octave, class_id = -1),
 *           so ONE device->host copy of n*28 bytes fills mvKeys' storage;
 *   x, y, angle (f32), octave (i32): [capacity] planes = the mvKeysUn arrays jsfe_build_frame_grid / jsfe_search_by_projection
 *           read, so the tracking-thread searches run on extractor output without a host round trip.
 * Any output pointer may be NULL.  Device pointers; the first n_keypoints entries are valid.  Not synchronised. */
typedef struct jsfe_cv_keypoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} jsfe_cv_keypoint;
int jsfe_frame_view(jsfe_handle* h, int slot, jsfe_cv_keypoint* keys, float* x, float* y, int32_t* octave, float* angle, void* stream);

/* ---- the input side on the device.  DEVICE pointers, enqueued on `stream`, not synchronised.
 * jsfe_remap_bilinear replaces cv::remap(im, rect, M1, M2, cv::INTER_LINEAR) of the reference's stereo examples
 * (Examples/Stereo/stereo_euroc.cpp:106-107,145-146): 8-bit single channel, CV_32FC1 maps [dst_h][dst_w] (contiguous) as
 * cv::initUndistortRectifyMap(..., CV_32F, ...) returns them, BORDER_CONSTANT 0, OpenCV's fixed-point bilinear scheme
 * (bit-exact with OpenCV 4.x).  `n_images` source images `src_stride` bytes apart share the maps (frames of one camera);
 * image i is written at dst + i*dst_stride -- e.g. straight into the level-0 slots (jsfe_slot_image) of every second slot.
 * jsfe_cvt_gray replaces cv::cvtColor(im, im, CV_RGB2GRAY | CV_BGR2GRAY | CV_RGBA2GRAY | CV_BGRA2GRAY) of
 * Tracking::GrabImageStereo (src/Tracking.cpp:260-285): channels = 3 or 4 interleaved, rgb_order != 0 for RGB(A). */
int jsfe_remap_bilinear(const uint8_t* src, int src_h, int src_w, int64_t src_pitch, int64_t src_stride, int n_images,
                        const float* map_x, const float* map_y, int dst_h, int dst_w, uint8_t* dst, int64_t dst_pitch,
                        int64_t dst_stride, void* stream);
int jsfe_cvt_gray(const uint8_t* src, int h, int w, int64_t src_pitch, int channels, int rgb_order, uint8_t* dst, int64_t dst_pitch,
                  void* stream);

/* ---- the results of a batch from every GPU on one rank ("NCCL gather of keypoint/descriptor buffers only when a
 * batch is requested", BASELINE config C5).  The reference has no counterpart (one device, src/cuda/orb_gpu.cpp:24).
 * One process per GPU; each rank owns a handle and a jsfe_gather bound to it.  jsfe_gather_begin packs the results of pairs
 * [first_pair, first_pair + n_pairs) -- trimmed to their keypoint counts -- into ONE region and moves it to the root on the gather's
 * own stream, ordered after `compute_stream`.  Nothing is made to wait at this point: the next call that overwrites this handle's
 * results (jsfe_extract, jsfe_stereo_match*, jsfe_process_host_pairs*) first waits, on the device, for the local packing (tens of
 * microseconds), so the transfer overlaps the following extraction and work of OTHER handles queued on the same stream is not held up.  Transport: if the non-root ranks mapped the root's landing buffers (CUDA IPC:
 * jsfe_gather_ipc_export on the root, jsfe_gather_ipc_import everywhere else, jsfe_gather_set_peers_mapped(root, 1)), a copy kernel
 * stores the trimmed region straight into the root's memory over NVLink and NCCL only carries two 4-byte all-reduces (double-buffer
 * credit in front of the stores, completion behind them); otherwise the regions travel as one ncclSend/ncclRecv group, padded to
 * the capacity bound.  The root's view of a batch stays valid until the root calls jsfe_gather_begin for the batch after the next.
 * `nccl_comm` is an ncclComm_t the host owns (e.g. torch.distributed's); libnccl is resolved with dlopen.  world == 1 needs none.
 *
 * Region layout (all little endian; jsfe_gather_region_bytes() apart on the root, in rank order):
 *    0  int32 magic 'JSG1', rank, n_pairs, capacity;  16  int64 payload_bytes;  24  int64 sequence;
 *   32  int32 n_keypoints[2 * n_pairs]            (slot order L0, R0, L1, R1, ...), zero-padded to a multiple of 16 bytes
 *   then one section per slot, each padded to 16 bytes:   int32 kps[6][n] | uint8 desc[n][32] | left slots only: float u_right[n], depth[n]
 * The gathered bytes are a pure function of the inputs: the slot sections of a pair are identical whatever the number of ranks. */
typedef struct jsfe_gather jsfe_gather;
typedef struct jsfe_gathered {
    const uint8_t* data;   /* root: DEVICE pointer to `world` regions; other ranks: NULL */
    int64_t region_stride; /* bytes between the regions of consecutive ranks             */
    int32_t world, root;
    int32_t n_pairs;       /* pairs per rank in this gather                               */
    int32_t transport;     /* 0: single rank, 1: peer-memory stores over NVLink, 2: NCCL send/recv */
} jsfe_gathered;
int64_t jsfe_gather_region_bytes(const jsfe_handle* h, int max_pairs);
int jsfe_gather_create(jsfe_handle* h, void* nccl_comm, int rank, int world, int root, int max_pairs, jsfe_gather** out);
int jsfe_gather_ipc_export(jsfe_gather* g, int buffer /* 0 | 1 */, uint8_t handle64[64]);
int jsfe_gather_ipc_import(jsfe_gather* g, int buffer /* 0 | 1 */, const uint8_t handle64[64]);
int jsfe_gather_set_peers_mapped(jsfe_gather* g, int all_mapped);
int jsfe_gather_begin(jsfe_gather* g, int first_pair, int n_pairs, void* compute_stream);
int jsfe_gather_end(jsfe_gather* g, jsfe_gathered* out); /* waits for the gather started by jsfe_gather_begin */
/* Optional stage timing of this rank's side of a gather (CUDA events on the gather's stream).  After jsfe_gather_profile(g, 1) every
 * jsfe_gather_begin is timed; jsfe_gather_stage_times gives, for the gather jsfe_gather_end returned last, microseconds of
 *   [0] the local pack, [1] the credit all-reduce (it ends when the SLOWEST rank has joined: rank skew shows up here),
 *   [2] the peer-memory stores (non-root ranks; NCCL transport: the send/recv group), [3] the completion all-reduce. */
int jsfe_gather_profile(jsfe_gather* g, int enable);
int jsfe_gather_stage_times(jsfe_gather* g, float us[4]);
int jsfe_gather_destroy(jsfe_gather* g);

/* Stage inspection for tests (device -> host, synchronous): level image, candidate cells, level keypoints. */
int jsfe_debug_level_image(jsfe_handle* h, int slot, int level, uint8_t* host_dst /* h*w contiguous */);
/* the 7x7-blurred level (descriptor input); zero outside [20,h-20)x[20,w-20) like the reference's image_gaussian_ */
int jsfe_debug_level_blur(jsfe_handle* h, int slot, int level, uint8_t* host_dst /* h*w contiguous */);
int jsfe_debug_cells(jsfe_handle* h, int slot, int32_t* x, int32_t* y, int32_t* score /* capacity each */);
int jsfe_debug_level_keypoints(jsfe_handle* h, int slot, int32_t* x, int32_t* y, int32_t* score, int32_t* level,
                               float* angle_rad /* capacity each, first N valid */);
/* Optional per-kernel timing: when enabled, every kernel launch is bracketed by CUDA events on the launching
 * stream.  jsfe_profile_read synchronises those events, returns the accumulated milliseconds and launch counts
 * per stage (JSFE_STAGE_*) since the last read, and resets the accumulators. */
enum { JSFE_STAGE_PYRAMID = 0, JSFE_STAGE_FAST_CELLS = 1, JSFE_STAGE_COMPACT = 2, JSFE_STAGE_ORIENT_DESC = 3,
       JSFE_STAGE_STEREO_MATCH = 4, JSFE_STAGE_STEREO_OUTLIER = 5, JSFE_STAGE_NMS_MS = 6, JSFE_STAGE_BLUR = 7, JSFE_NUM_STAGES = 8 };
int jsfe_profile_enable(jsfe_handle* h, int on);
int jsfe_profile_read(jsfe_handle* h, float* stage_ms, int64_t* stage_launches, int n_stages);
/* number of kernels the library has launched since creation (bench.py's gpu_launches) */
int64_t jsfe_launch_count(const jsfe_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* JSFE_H */
