// compat/frame_view.hpp -- SURVEY.md 8(f4): the output side of Frame::Frame as a lazily-materialised host view.
//
// The reference brings four arrays to the host after every extraction and unpacks the left and right SoA into
// std::vector<cv::KeyPoint> one field at a time (src/Frame.cpp:116-196), whether or not the tracking thread ever looks at them, and
// Frame::ComputeBoW later splits mDescriptors into one cv::Mat per keypoint (src/Frame.cpp:709-716 -> Converter::toDescriptorVector).
// LazyFrameView keeps one eye's results on the device (they stay in the extractor's slot, where jsfe_build_frame_grid and
// jsfe_search_by_projection read them) and materialises
//     keys()         -> std::vector<cv::KeyPoint>   ONE device-to-host copy of n * 28 bytes (jsfe_frame_view writes cv::KeyPoint records)
//     descriptors()  -> cv::Mat n x 32 CV_8UC1       ONE copy of n * 32 bytes
//     bow_vector()   -> std::vector<cv::Mat>         row headers into descriptors() (what Converter::toDescriptorVector returns), no copy
// on first use only.  Same field values as the reference's loop: pt = (float)int, response = (float)score, angle in degrees,
// octave, size = (float)int, class_id = -1.
#ifndef JSFE_COMPAT_FRAME_VIEW_HPP
#define JSFE_COMPAT_FRAME_VIEW_HPP

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <opencv2/opencv.hpp>

#include <cuda_runtime_api.h>

#include <jsfe.h>

namespace jsfe_compat {

class LazyFrameView {
    static_assert(sizeof(cv::KeyPoint) == sizeof(jsfe_cv_keypoint), "cv::KeyPoint layout (pt.x, pt.y, size, angle, response, octave, class_id)");
    static void check(int rc, const char* what) {
        if (rc != 0) { fprintf(stderr, "jsfe: %s failed (%d): %s\n", what, rc, jsfe_last_error()); abort(); }
    }

public:
    // Call after jsfe_extract of `slot` on `stream`; nothing is copied here.
    LazyFrameView(jsfe_handle* h, int slot, cudaStream_t stream = nullptr) : h_(h), slot_(slot), stream_(stream) {
        cap_ = jsfe_max_keypoints(h);
        check(jsfe_slot_view_get(h, slot, &view_), "jsfe_slot_view_get");
    }
    ~LazyFrameView() { if (d_keys_) cudaFree(d_keys_); }
    LazyFrameView(const LazyFrameView&) = delete;
    LazyFrameView& operator=(const LazyFrameView&) = delete;

    int size() {
        if (n_ < 0) {
            cudaMemcpyAsync(&n_, view_.n_keypoints, sizeof(int), cudaMemcpyDeviceToHost, stream_);
            cudaStreamSynchronize(stream_);
        }
        return n_;
    }
    const std::vector<cv::KeyPoint>& keys() {
        if (!have_keys_) {
            if (!d_keys_ && cudaMalloc((void**)&d_keys_, sizeof(jsfe_cv_keypoint) * (size_t)cap_) != cudaSuccess) { fprintf(stderr, "jsfe: cudaMalloc failed\n"); abort(); }
            check(jsfe_frame_view(h_, slot_, d_keys_, nullptr, nullptr, nullptr, nullptr, stream_), "jsfe_frame_view");
            const int n = size();
            keys_.resize(n);
            if (n) cudaMemcpyAsync(keys_.data(), d_keys_, sizeof(jsfe_cv_keypoint) * (size_t)n, cudaMemcpyDeviceToHost, stream_);
            cudaStreamSynchronize(stream_);
            have_keys_ = true;
        }
        return keys_;
    }
    const cv::Mat& descriptors() {
        if (!have_desc_) {
            const int n = size();
            desc_store_.resize((size_t)n * 32);
            if (n) cudaMemcpyAsync(desc_store_.data(), view_.desc, (size_t)n * 32, cudaMemcpyDeviceToHost, stream_);
            cudaStreamSynchronize(stream_);
            desc_ = cv::Mat(n, 32, CV_8UC1, desc_store_.data());
            have_desc_ = true;
        }
        return desc_;
    }
    // Converter::toDescriptorVector(mDescriptors): one 1 x 32 header per keypoint, sharing the storage of descriptors()
    std::vector<cv::Mat> bow_vector() {
        const cv::Mat& d = descriptors();
        std::vector<cv::Mat> v;
        v.reserve(d.rows);
        for (int i = 0; i < d.rows; ++i) v.push_back(cv::Mat(1, 32, CV_8UC1, desc_store_.data() + (size_t)i * 32));
        return v;
    }
    const jsfe_slot_view& device() const { return view_; }   // u_right, depth, descriptors, ... still on the device

private:
    jsfe_handle* h_;
    int slot_;
    cudaStream_t stream_;
    int cap_ = 0, n_ = -1;
    jsfe_slot_view view_;
    jsfe_cv_keypoint* d_keys_ = nullptr;
    bool have_keys_ = false, have_desc_ = false;
    std::vector<cv::KeyPoint> keys_;
    std::vector<unsigned char> desc_store_;
    cv::Mat desc_;
};

}  // namespace jsfe_compat
#endif
