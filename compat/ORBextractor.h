// compat/ORBextractor.h -- header-compatible shim of Jetson_SLAM::ORBExtractor (reference include/ORBextractor.h:21-98,
// src/ORBextractor.cpp:27-105): same constructor, extract(), scale getters and the public `orb_gpu_` member that
// Frame::ComputeStereoMatches dereferences (src/Frame.cpp:784-785).
#ifndef JSFE_COMPAT_ORBEXTRACTOR_H
#define JSFE_COMPAT_ORBEXTRACTOR_H

#include <string>
#include <vector>

#include <opencv2/opencv.hpp>

#include "cuda/orb_gpu.hpp"

namespace Jetson_SLAM {

class ORBExtractor {
public:
    ORBExtractor(int im_height, int im_width, float scale_factor, int n_levels, int FAST_N_MIN, int FAST_N_MAX, int th_FAST_MIN,
                 int th_FAST_MAX, std::string str_mask, int tile_h, int tile_w, bool fixed_multi_scale_tile_size, bool apply_nms_ms,
                 bool nms_ms_mode_gpu, bool use_gpu = false)
        : n_levels_(n_levels), scale_factor_(scale_factor), use_gpu_(use_gpu) {
        scale_.resize(n_levels); inv_scale_.resize(n_levels); level_sigma2_.resize(n_levels); inv_level_sigma2_.resize(n_levels);
        scale_[0] = 1.0f; level_sigma2_[0] = 1.0f;
        for (int i = 1; i < n_levels; ++i) { scale_[i] = scale_[i - 1] * scale_factor; level_sigma2_[i] = scale_[i] * scale_[i]; }
        for (int i = 0; i < n_levels; ++i) { inv_scale_[i] = 1.0f / scale_[i]; inv_level_sigma2_[i] = 1.0f / level_sigma2_[i]; }
        // like the reference, the GPU path is unconditional (`use_gpu` is stored and ignored, src/ORBextractor.cpp:74-87)
        orb_gpu_ = new orb_cuda::ORB_GPU(im_height, im_width, n_levels, scale_factor, FAST_N_MIN, FAST_N_MAX, th_FAST_MIN, th_FAST_MAX,
                                         tile_h, tile_w, fixed_multi_scale_tile_size, apply_nms_ms, nms_ms_mode_gpu, str_mask, 0);
    }
    ~ORBExtractor() { delete orb_gpu_; }
    ORBExtractor(const ORBExtractor&) = delete;
    ORBExtractor& operator=(const ORBExtractor&) = delete;

    void extract(const cv::Mat& image, orb_cuda::SyncedMem<int>& keypoints, orb_cuda::SyncedMem<unsigned char>& keypoints_desc) {
        orb_gpu_->extract(image, keypoints, keypoints_desc);
    }
    int get_levels() { return n_levels_; }
    float get_scale_factor() { return scale_factor_; }
    const std::vector<float> get_scale_factors() { return scale_; }
    std::vector<float> get_inverse_scale_factors() { return inv_scale_; }
    std::vector<float> get_scale_sigma_squares() { return level_sigma2_; }
    std::vector<float> get_inverse_scale_sigma_squares() { return inv_level_sigma2_; }

    orb_cuda::ORB_GPU* orb_gpu_;

protected:
    std::vector<float> scale_, inv_scale_, level_sigma2_, inv_level_sigma2_;
    int n_levels_;
    float scale_factor_;
    bool use_gpu_;
};

}  // namespace Jetson_SLAM
#endif
