// ref_driver.cpp -- C entry points around the reference's OWN classes (TEST INFRASTRUCTURE ONLY).
//
// The reference sources under /root/reference/src/cuda are compiled unmodified (see Makefile);
// this file is the only code of ours in oracle/_ref/libjsref.so.  It drives orb_cuda::ORB_GPU the
// way Frame::Frame / Frame::ComputeStereoMatches do (src/Frame.cpp:80-250,780-803) and exposes
// intermediate buffers so the CPU oracle and the new CUDA path can be pinned against them.
#include <cuda/orb_gpu.hpp>
#include <cuda/orb_matcher.hpp>
#include <cuda/tracking_gpu.hpp>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

using orb_cuda::ORB_GPU;
using orb_cuda::SyncedMem;

namespace {
struct RefEye {
    ORB_GPU* gpu = nullptr;
    SyncedMem<int> kps;            // out_keypoints of the last extract (6*N)
    SyncedMem<unsigned char> desc; // out descriptors (32*N)
    int H = 0, W = 0, L = 0;
};

// The reference reads buffers it never initialises (score maps outside the border ring, blurred
// image border, nms_s_level_: SURVEY.md App. A.9).  cudaMalloc'd pages are not guaranteed zero, so
// define them once here; the kernels themselves are untouched.
void zero_uninitialised(ORB_GPU* g) {
    for (int i = 0; i < g->n_levels_; ++i) {
        cudaMemset(g->score_[i].gpu_data(), 0, sizeof(int) * g->height_[i] * g->width_[i]);
        cudaMemset(g->score_nms_[i].gpu_data(), 0, sizeof(int) * g->height_[i] * g->width_[i]);
        cudaMemset(g->image_gaussian_[i].gpu_data(), 0, (size_t)g->height_[i] * g->width_[i]);
        cudaMemset(g->image_[i].gpu_data(), 0, (size_t)g->height_[i] * g->width_[i]);
    }
    // LUT[0xFFFF] lies one int past the reference's 65535-entry table (orb_gpu.cpp:373) and is read
    // whenever all 16 ring pixels are brighter/darker than the centre.  cudaMalloc rounds the 262140-byte
    // table up to its 256-byte granule, so that word exists; pin it to 0 (its true value for N_MAX < 16).
    for (int i = 0; i < g->n_levels_; ++i)
        cudaMemset((char*)g->lookup_table_[i].gpu_data() + sizeof(int) * 0xFFFF, 0, sizeof(int));
    if (g->nms_s_level_.gpu_data()) g->nms_s_level_.set_zero_gpu();
    if (g->nms_s_score_.gpu_data()) g->nms_s_score_.set_zero_gpu();
    cudaMemset(g->keypoints_.gpu_data(), 0, sizeof(int) * g->keypoints_.count_);
    cudaDeviceSynchronize();
}

void unpack(SyncedMem<int>& k, std::vector<cv::KeyPoint>& out) {  // src/Frame.cpp:124-157
    const int N = k.count_ / 6;
    const int* d = k.cpu_data();
    out.resize(N);
    for (int i = 0; i < N; ++i) {
        out[i].pt.x = d[0 * N + i];
        out[i].pt.y = d[1 * N + i];
        out[i].response = d[2 * N + i];
        out[i].angle = ((const float*)d)[3 * N + i];
        out[i].octave = d[4 * N + i];
        out[i].size = d[5 * N + i];
    }
}
}  // namespace

extern "C" {

void* jsref_create(int H, int W, int L, float scale, int n_min, int n_max, int th_min, int th_max,
                   int tile_h, int tile_w, int fixed_tile, int nms_ms, int nms_ms_gpu) {
    RefEye* e = new RefEye;
    e->H = H; e->W = W; e->L = L;
    e->gpu = new ORB_GPU(H, W, L, scale, n_min, n_max, th_min, th_max, tile_h, tile_w, fixed_tile != 0,
                         nms_ms != 0, nms_ms_gpu != 0, std::string(""), 0);
    zero_uninitialised(e->gpu);
    return e;
}

void jsref_destroy(void* h) {
    RefEye* e = (RefEye*)h;
    delete e->gpu;
    delete e;
}

int jsref_max_kp(void* h) { return ((RefEye*)h)->gpu->max_kp_count_; }

// ORB_GPU::extract + the caller's D2H (Frame.cpp:119-122).  Returns N.
int jsref_extract(void* h, const uint8_t* img, int32_t* kps_out, uint8_t* desc_out) {
    RefEye* e = (RefEye*)h;
    cv::Mat m(e->H, e->W, CV_8UC1, (void*)img);
    e->gpu->extract(m, e->kps, e->desc);
    e->kps.to_cpu();
    e->desc.to_cpu();
    const int N = e->kps.count_ / 6;
    if (kps_out) memcpy(kps_out, e->kps.cpu_data(), sizeof(int) * 6 * N);
    if (desc_out) memcpy(desc_out, e->desc.cpu_data(), 32 * N);
    return N;
}

// intermediate buffers of the last extract
void jsref_level_dims(void* h, int32_t* hh, int32_t* ww) {
    RefEye* e = (RefEye*)h;
    for (int i = 0; i < e->L; ++i) { hh[i] = e->gpu->height_[i]; ww[i] = e->gpu->width_[i]; }
}
void jsref_level_image(void* h, int l, uint8_t* out) {
    ORB_GPU* g = ((RefEye*)h)->gpu;
    cudaMemcpy(out, g->image_[l].gpu_data(), (size_t)g->height_[l] * g->width_[l], cudaMemcpyDeviceToHost);
}
void jsref_level_blur(void* h, int l, uint8_t* out) {
    ORB_GPU* g = ((RefEye*)h)->gpu;
    cudaMemcpy(out, g->image_gaussian_[l].gpu_data(), (size_t)g->height_[l] * g->width_[l], cudaMemcpyDeviceToHost);
}
void jsref_level_score(void* h, int l, int32_t* out) {
    ORB_GPU* g = ((RefEye*)h)->gpu;
    cudaMemcpy(out, g->score_[l].gpu_data(), sizeof(int) * g->height_[l] * g->width_[l], cudaMemcpyDeviceToHost);
}
// level-coordinate keypoints after compaction: x|y|score|angle planes of max_kp, plus n per level / offsets
void jsref_level_keypoints(void* h, int32_t* x, int32_t* y, int32_t* s, float* a, int32_t* n_per_level,
                           int32_t* level_offset) {
    ORB_GPU* g = ((RefEye*)h)->gpu;
    const int M = g->max_kp_count_;
    std::vector<int> buf(g->keypoints_.count_);
    cudaMemcpy(buf.data(), g->keypoints_.gpu_data(), sizeof(int) * buf.size(), cudaMemcpyDeviceToHost);
    memcpy(x, buf.data() + g->x_offset_, sizeof(int) * M);
    memcpy(y, buf.data() + g->y_offset_, sizeof(int) * M);
    memcpy(s, buf.data() + g->s_offset_, sizeof(int) * M);
    memcpy(a, buf.data() + g->a_offset_, sizeof(int) * M);
    for (int i = 0; i < g->n_levels_; ++i) { n_per_level[i] = g->n_keypoints_[i]; level_offset[i] = g->level_offset_[i]; }
}
void jsref_tables(void* h, int32_t* lut65535, int32_t* umax16, float* gauss49, int8_t* patx512, int8_t* paty512) {
    ORB_GPU* g = ((RefEye*)h)->gpu;
    memcpy(lut65535, g->lookup_table_[0].cpu_data(), sizeof(int) * 0xFFFF);
    memcpy(umax16, g->umax_[0].cpu_data(), sizeof(int) * 16);
    memcpy(gauss49, g->gaussian_weights_[0].cpu_data(), sizeof(float) * 49);
    memcpy(patx512, g->pattern_x_[0].cpu_data(), 512);
    memcpy(paty512, g->pattern_y_[0].cpu_data(), 512);
}

// Frame::ComputeStereoMatches on the last extracts of the two eyes.  mb is passed explicitly
// (the reference reads Frame::mb before it is assigned).
int jsref_stereo_match(void* hl, void* hr, int th_high, int th_low, float mb, float mbf, float* u_right, float* depth) {
    RefEye* l = (RefEye*)hl; RefEye* r = (RefEye*)hr;
    std::vector<cv::KeyPoint> kl, kr;
    unpack(l->kps, kl);
    unpack(r->kps, kr);
    std::vector<float> ur, dp;
    l->gpu->ORB_compute_stereo_match(th_high, th_low, mb, mbf, l->gpu->height_, l->gpu->width_, kl, kr, ur, dp,
                                     l->desc.gpu_data(), r->desc.gpu_data(), l->gpu->image_, r->gpu->image_);
    memcpy(u_right, ur.data(), sizeof(float) * ur.size());
    memcpy(depth, dp.data(), sizeof(float) * dp.size());
    return (int)ur.size();
}

// Timing of the reference's per-frame stereo path exactly as Frame::Frame sequences it:
// two host threads extract L/R, 4 sync D2H, AoS unpack, stereo match.  Returns seconds for `iters` pairs.
double jsref_time_pairs(void* hl, void* hr, const uint8_t* img_l, const uint8_t* img_r, float mb, float mbf, int iters,
                        int two_threads) {
    RefEye* l = (RefEye*)hl; RefEye* r = (RefEye*)hr;
    cv::Mat ml(l->H, l->W, CV_8UC1, (void*)img_l), mr(r->H, r->W, CV_8UC1, (void*)img_r);
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < iters; ++it) {
        if (two_threads) {
            std::thread a([&] { l->gpu->extract(ml, l->kps, l->desc); });
            std::thread b([&] { r->gpu->extract(mr, r->kps, r->desc); });
            a.join(); b.join();
        } else {
            l->gpu->extract(ml, l->kps, l->desc);
            r->gpu->extract(mr, r->kps, r->desc);
        }
        l->kps.to_cpu(); r->kps.to_cpu(); l->desc.to_cpu(); r->desc.to_cpu();
        std::vector<cv::KeyPoint> kl, kr;
        unpack(l->kps, kl);
        unpack(r->kps, kr);
        std::vector<float> ur, dp;
        if (!kl.empty())
            l->gpu->ORB_compute_stereo_match(100, 50, mb, mbf, l->gpu->height_, l->gpu->width_, kl, kr, ur, dp,
                                             l->desc.gpu_data(), r->desc.gpu_data(), l->gpu->image_, r->gpu->image_);
    }
    cudaDeviceSynchronize();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// ---- adjacent helpers: thin pass-throughs to the reference's own functions (device pointers) ----
void jsref_project_points(int n, float* px, float* py, float* pz, float* R, float* t, float fx, float fy, float cx, float cy,
                          float minx, float maxx, float miny, float maxy, float* u, float* v, float* invz, unsigned char* ok) {
    orb_cuda::ORB_Search_by_projection_project_on_frame(n, px, py, pz, R, t, fx, fy, cx, cy, minx, maxx, miny, maxy, u, v, invz, ok);
}
void jsref_hamming_pairs(int n, int* il, int* ir, unsigned char* dl, unsigned char* dr, int* dist) {
    orb_cuda::ORB_compute_distances(n, il, ir, dl, dr, dist);
}
void jsref_in_frustum(int n, float* px, float* py, float* pz, float* pnx, float* pny, float* pnz, float* md, float* imax, float* imin,
                      float* R, float* t, float* ow, float fx, float fy, float cx, float cy, int minx, int maxx, int miny, int maxy,
                      int nlev, float logsf, float vca, float* invz, float* u, float* v, int* lvl, float* vc, unsigned char* in) {
    tracking_cuda::compute_isInFrustum_GPU(n, px, py, pz, pnx, pny, pnz, md, imax, imin, R, t, ow, fx, fy, cx, cy, minx, maxx, miny, maxy,
                                           nlev, logsf, vca, invz, u, v, lvl, vc, in);
}

}  // extern "C"
