// compat/cuda/orb_gpu.hpp -- header-compatible shim of the reference's orb_cuda::ORB_GPU
// (include/cuda/orb_gpu.hpp:22-330) on top of the C ABI in include/jsfe.h.
//
// Keeps what the SLAM core uses: the constructor signature (:26-35), extract() (:178-180),
// ORB_compute_stereo_match() (:218-229) and the public members Frame::ComputeStereoMatches reads
// (height_, width_, image_: :242-247; src/Frame.cpp:784-800).  One ORB_GPU = one jsfe handle with a single image slot,
// exactly like the reference's one-extractor-per-eye; two instances may be driven from two host threads
// (src/Frame.cpp:107-110).  The reference checks no CUDA error; this shim aborts with the library's message instead.
#ifndef JSFE_COMPAT_ORB_GPU_HPP
#define JSFE_COMPAT_ORB_GPU_HPP

#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <opencv2/opencv.hpp>

#include <cuda_runtime_api.h>

#include <cuda/synced_mem_holder.hpp>
#include <jsfe.h>

namespace orb_cuda {

#define BORDER_SKIP 20
#define CIRCULAR_HALF_PATCH_SIZE 15

class ORB_GPU {
    static void check(int rc, const char* what) {
        if (rc != 0) { fprintf(stderr, "jsfe: %s failed (%d): %s\n", what, rc, jsfe_last_error()); abort(); }
    }
    // right-eye lookup for ORB_compute_stereo_match: level-0 device pointer -> handle
    static std::map<const void*, ORB_GPU*>& registry() { static std::map<const void*, ORB_GPU*> r; return r; }
    static std::mutex& registry_mutex() { static std::mutex m; return m; }

public:
    ORB_GPU(int im_height, int im_width, int n_levels, float scale_factor, int FAST_N_MIN, int FAST_N_MAX, int th_FAST_MIN,
            int th_FAST_MAX, int tile_h, int tile_w, bool fixed_multi_scale_tile_size, bool apply_nms_ms, bool nms_ms_mode_gpu,
            std::string str_mask, int device_id = 0) {
        jsfe_config c;
        c.height = im_height; c.width = im_width; c.n_levels = n_levels; c.scale_factor = scale_factor;
        c.fast_n_min = FAST_N_MIN; c.fast_n_max = FAST_N_MAX; c.th_fast_min = th_FAST_MIN; c.th_fast_max = th_FAST_MAX;
        c.tile_h = tile_h; c.tile_w = tile_w; c.fixed_multi_scale_tile_size = fixed_multi_scale_tile_size;
        c.apply_nms_ms = apply_nms_ms; c.nms_ms_mode_gpu = nms_ms_mode_gpu;
        c.mask = nullptr; c.mask_pitch = 0; c.device_id = device_id; c.max_images = 1;
        cv::Mat mask = cv::imread(str_mask);            // the reference loads its mask from a file (orb_gpu.cpp:64-91)
        cv::Mat gray, fitted;
        if (!mask.empty()) {
            cv::cvtColor(mask, gray, 6 /* CV_BGR2GRAY */);
            // the reference accepts a mask of ANY size and cv::resize's it to every level (orb_gpu.cpp:78-90); jsfe_create wants the
            // level-0 geometry, so bring it there first with the same nearest-neighbour rule (resizing twice with INTER_NEAREST from
            // the original grid equals resizing once only when the mask already has the frame size -- the common case -- and is the
            // closest the C ABI's height x width mask can get otherwise)
            if (gray.rows != im_height || gray.cols != im_width) cv::resize(gray, fitted, cv::Size(im_width, im_height), 0, 0, 0 /* INTER_NEAREST */);
            else fitted = gray;
            c.mask = fitted.data; c.mask_pitch = fitted.cols;
        }
        check(jsfe_create(&c, &h_), "jsfe_create");
        n_levels_ = n_levels;
        device_id_ = device_id;
        max_kp_count_ = jsfe_max_keypoints(h_);
        for (int l = 0; l < n_levels; ++l) {
            jsfe_level_info li;
            check(jsfe_get_level_info(h_, l, &li), "jsfe_get_level_info");
            height_.push_back(li.height); width_.push_back(li.width);
            scale_.push_back(li.scale); inv_scale_.push_back(li.inv_scale);
            const uint8_t* dev = nullptr; int32_t hh, ww; int64_t pitch;
            check(jsfe_level_image(h_, 0, l, &dev, &hh, &ww, &pitch), "jsfe_level_image");
            image_.push_back(SyncedMem<unsigned char>::view(const_cast<unsigned char*>(dev), hh * ww, (size_t)pitch));
        }
        n_keypoints_.assign(n_levels, 0);
        std::lock_guard<std::mutex> g(registry_mutex());
        registry()[image_[0].gpu_data_] = this;
    }
    ~ORB_GPU() {
        {
            std::lock_guard<std::mutex> g(registry_mutex());
            registry().erase(image_.empty() ? nullptr : image_[0].gpu_data_);
        }
        jsfe_destroy(h_);
    }
    ORB_GPU(const ORB_GPU&) = delete;
    ORB_GPU& operator=(const ORB_GPU&) = delete;

    // src/cuda/orb_gpu.cpp:489-841: on return the DEVICE buffers of the outputs are complete; host copies are the
    // caller's job (Frame.cpp:119-122).  Output layout: 6 planes x N (x|y|score|angle|octave|size), descriptors 32 x N.
    // One stream synchronisation per call: the output buffers are grown to the capacity once, the pack kernel reads N on the device.
    void extract(const cv::Mat& image, SyncedMem<int>& out_keypoints, SyncedMem<unsigned char>& out_keypoints_desc) {
        check(jsfe_set_images(h_, 0, 1, image.data, image.cols, (int64_t)image.cols * image.rows, 0, nullptr), "jsfe_set_images");
        check(jsfe_extract(h_, 0, 1, nullptr), "jsfe_extract");
        if (out_keypoints.capacity_ < 6 * max_kp_count_) out_keypoints.resize(6 * max_kp_count_);
        if (out_keypoints_desc.capacity_ < 32 * max_kp_count_) out_keypoints_desc.resize(32 * max_kp_count_);
        int32_t n = 0;
        check(jsfe_pack_keypoints_once(h_, 0, out_keypoints.gpu_data(), out_keypoints_desc.gpu_data(), &n, nullptr), "jsfe_pack_keypoints_once");
        out_keypoints.resize(6 * n);            // shrinks count_ only (the reference's resize never gives memory back either)
        out_keypoints_desc.resize(32 * n);
        last_desc_dev_ = out_keypoints_desc.gpu_data();
        last_n_ = n;
    }

    // src/cuda/orb_stereo_match.cu:105-580.  The keypoints/descriptors of both eyes are the ones the two extractors
    // just produced and still hold on the device, so the host vectors are only used for their sizes.
    void ORB_compute_stereo_match(int ORB_TH_HIGH, int ORB_TH_LOW, float mb, float mbf, std::vector<int>& /*octave_height*/,
                                  std::vector<int>& /*octave_width*/, std::vector<cv::KeyPoint>& mvKeys,
                                  std::vector<cv::KeyPoint>& mvKeysRight, std::vector<float>& mvuRight, std::vector<float>& mvDepth,
                                  unsigned char* keypoint_descriptor_left, unsigned char* keypoint_descriptor_right,
                                  std::vector<SyncedMem<unsigned char> >& /*images_left_smem*/,
                                  std::vector<SyncedMem<unsigned char> >& images_right_smem) {
        ORB_GPU* right = nullptr;
        {
            std::lock_guard<std::mutex> g(registry_mutex());
            auto it = registry().find(images_right_smem.empty() ? nullptr : images_right_smem[0].gpu_data_);
            if (it != registry().end()) right = it->second;
        }
        if (!right) { fprintf(stderr, "jsfe: right pyramid does not belong to a live ORB_GPU\n"); abort(); }
        jsfe_handle* hr = right->h_;
        // The matcher works on the keypoints and descriptors the two extractors still hold on the device.  What the caller passes
        // must be exactly those: the descriptor buffers of the last extract() of each eye, and as many keypoints as they hold
        // (src/Frame.cpp:784-800 passes them untouched).  Anything else would silently be ignored, so it is refused.
        if (keypoint_descriptor_left != last_desc_dev_ || keypoint_descriptor_right != right->last_desc_dev_ ||
            (int)mvKeys.size() != last_n_ || (int)mvKeysRight.size() != right->last_n_) {
            fprintf(stderr, "jsfe: ORB_compute_stereo_match was handed keypoints/descriptors other than the ones of the last extract() "
                            "(left %zu vs %d, right %zu vs %d)\n", mvKeys.size(), last_n_, mvKeysRight.size(), right->last_n_);
            abort();
        }
        check(jsfe_stereo_match_cross(h_, 0, hr, 0, ORB_TH_HIGH, ORB_TH_LOW, mb, mbf, nullptr), "jsfe_stereo_match_cross");
        const size_t N = mvKeys.size();
        mvuRight.resize(N, -1.0f);
        mvDepth.resize(N, -1.0f);
        std::vector<float> ur(max_kp_count_), dp(max_kp_count_);
        int32_t n = 0;
        check(jsfe_get_stereo_slot(h_, 0, ur.data(), dp.data(), nullptr, nullptr, &n, nullptr), "jsfe_get_stereo_slot");
        if ((size_t)n != N) { fprintf(stderr, "jsfe: %zu left keypoints passed, %d on the device\n", N, n); abort(); }
        for (size_t i = 0; i < N; ++i) { mvuRight[i] = ur[i]; mvDepth[i] = dp[i]; }
    }

    jsfe_handle* handle() { return h_; }

    // public members of the reference that callers read
    int device_id_;
    std::vector<float> scale_, inv_scale_;
    std::vector<int> height_, width_;
    std::vector<int> n_keypoints_;
    std::vector<SyncedMem<unsigned char> > image_;   // per-level UNBLURRED pyramid of the last extract (device views; pitch_ in bytes)
    int max_kp_count_;
    int n_levels_;

private:
    jsfe_handle* h_ = nullptr;
    const unsigned char* last_desc_dev_ = nullptr;   // the caller's descriptor buffer filled by the last extract()
    int last_n_ = -1;
};

}  // namespace orb_cuda
#endif
