// tests/cpp/frame_hotpath.cpp -- drives the compat/ shims exactly the way the reference's Frame::Frame does on the hot
// path (src/Frame.cpp:103-122 two extractor threads + 4 D2H, :124-196 SoA unpack, :780-803 ComputeStereoMatches) and dumps
// the results for tests/test_compat_cpp.py to compare with the oracle.  Compiled against the OpenCV stand-in header of
// oracle/ref_build/stub (the image has no OpenCV C++ headers); with real OpenCV the same source compiles unchanged.
//
// usage: frame_hotpath H W levels scale nmin nmax thmin thmax tile_h tile_w left.raw right.raw mb mbf out.bin [--time N]
//   --time N: after the checked run, repeat the frame body (two extractor threads, 4 D2H, unpack, stereo match) N times and print
//             "fps <frames per second>": the reference-API path a Jetson-SLAM user calls (bench.py: compat_api)
#include <ORBextractor.h>
#include <frame_view.hpp>

#include <chrono>
#include <cstring>

#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

using namespace Jetson_SLAM;
using orb_cuda::SyncedMem;

static std::vector<unsigned char> read_raw(const char* path, size_t n) {
    std::vector<unsigned char> v(n);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return v;
}

static void unpack(SyncedMem<int>& k, std::vector<cv::KeyPoint>& out) {   // src/Frame.cpp:124-157
    const int N = k.count_ / 6;
    const int* d = k.cpu_data();
    out.resize(N);
    for (int i = 0; i < N; ++i) {
        out[i].pt.x = d[0 * N + i]; out[i].pt.y = d[1 * N + i]; out[i].response = d[2 * N + i];
        out[i].angle = ((const float*)d)[3 * N + i]; out[i].octave = d[4 * N + i]; out[i].size = d[5 * N + i];
    }
}

int main(int argc, char** argv) {
    int time_frames = 0;
    if (argc == 18 && !strcmp(argv[16], "--time")) { time_frames = atoi(argv[17]); argc = 16; }
    if (argc != 16) { fprintf(stderr, "bad usage\n"); return 2; }
    const int H = atoi(argv[1]), W = atoi(argv[2]), L = atoi(argv[3]);
    const float scale = atof(argv[4]);
    const int nmin = atoi(argv[5]), nmax = atoi(argv[6]), thmin = atoi(argv[7]), thmax = atoi(argv[8]), th = atoi(argv[9]), tw = atoi(argv[10]);
    std::vector<unsigned char> l = read_raw(argv[11], (size_t)H * W), r = read_raw(argv[12], (size_t)H * W);
    const float mb = atof(argv[13]), mbf = atof(argv[14]);
    ORBExtractor exl(H, W, scale, L, nmin, nmax, thmin, thmax, "", th, tw, false, false, true, true);
    ORBExtractor exr(H, W, scale, L, nmin, nmax, thmin, thmax, "", th, tw, false, false, true, true);
    cv::Mat iml(H, W, CV_8UC1, l.data()), imr(H, W, CV_8UC1, r.data());
    std::vector<float> ur, dp;
    std::vector<cv::KeyPoint> kl, kr;
    SyncedMem<int> kps_l, kps_r;
    SyncedMem<unsigned char> desc_l, desc_r;
    auto frame = [&] {                            // src/Frame.cpp:103-122, 124-196, 780-803
        std::thread a([&] { exl.extract(iml, kps_l, desc_l); });
        std::thread b([&] { exr.extract(imr, kps_r, desc_r); });
        a.join(); b.join();
        kps_l.to_cpu(); kps_r.to_cpu(); desc_l.to_cpu(); desc_r.to_cpu();
        unpack(kps_l, kl); unpack(kps_r, kr);
        ur.clear(); dp.clear();
        orb_cuda::ORB_GPU& gl = *exl.orb_gpu_; orb_cuda::ORB_GPU& gr = *exr.orb_gpu_;
        gl.ORB_compute_stereo_match(100, 50, mb, mbf, gl.height_, gl.width_, kl, kr, ur, dp, desc_l.gpu_data(), desc_r.gpu_data(),
                                    gl.image_, gr.image_);
    };
    for (int rep = 0; rep < 2; ++rep) frame();    // twice: buffers are reused from frame to frame
    {   // the lazily-materialised view gives the same mvKeys / mDescriptors as the reference's unpack loop
        jsfe_compat::LazyFrameView view(exl.orb_gpu_->handle(), 0);
        const std::vector<cv::KeyPoint>& vk = view.keys();
        const cv::Mat& vd = view.descriptors();
        bool same = vk.size() == kl.size() && vd.rows == (int)kl.size();
        for (size_t i = 0; same && i < kl.size(); ++i)
            same = vk[i].pt.x == kl[i].pt.x && vk[i].pt.y == kl[i].pt.y && vk[i].response == kl[i].response && vk[i].octave == kl[i].octave &&
                   vk[i].size == kl[i].size && !memcmp(&vk[i].angle, &kl[i].angle, 4) && vk[i].class_id == -1;
        same = same && !memcmp(vd.data, desc_l.cpu_data(), 32 * kl.size());
        std::vector<cv::Mat> bow = view.bow_vector();
        same = same && bow.size() == kl.size() && (bow.empty() || (bow.back().data == vd.data + 32 * (kl.size() - 1) && bow.back().cols == 32));
        if (!same) { fprintf(stderr, "LazyFrameView differs from the reference unpack loop\n"); return 3; }
    }
    if (time_frames > 0) {
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < time_frames; ++i) frame();
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("fps %.2f\n", time_frames / dt);
    }
    SyncedMem<int> copy = kps_l;                  // the SLAM core copies these objects (Tracking.cpp:292); must not double-free
    FILE* f = fopen(argv[15], "wb");
    const int nl = kps_l.count_ / 6, nr = kps_r.count_ / 6;
    fwrite(&nl, 4, 1, f); fwrite(&nr, 4, 1, f);
    fwrite(kps_l.cpu_data(), 4, 6 * nl, f); fwrite(desc_l.cpu_data(), 1, 32 * nl, f);
    fwrite(kps_r.cpu_data(), 4, 6 * nr, f); fwrite(desc_r.cpu_data(), 1, 32 * nr, f);
    fwrite(ur.data(), 4, nl, f); fwrite(dp.data(), 4, nl, f);
    fclose(f);
    printf("ok %d %d\n", nl, nr);
    return 0;
}
