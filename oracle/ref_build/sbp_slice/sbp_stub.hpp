// sbp_stub.hpp -- the minimum of Frame / MapPoint / cv::Mat / SyncedMem that the reference's own source lines of
// ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (src/ORBmatcher.cpp:1647-1963), ORBmatcher::ComputeThreeMaxima
// (:2097-2138), Frame::AssignFeaturesToGrid (src/Frame.cpp:464-479), Frame::GetFeaturesInArea (:569-639) and Frame::PosInGrid
// (:696-706) need to compile UNMODIFIED on the CPU.  TEST INFRASTRUCTURE (oracle/): it exists to produce reference fixtures for
// SURVEY.md 8(f1); nothing under jetson_slam_b200/ uses it.  Member names follow include/Frame.h, include/MapPoint.h,
// include/ORBmatcher.h and include/cuda/synced_mem_holder.hpp of the reference; only what those lines touch is declared.
// The two device calls of the sliced code (orb_cuda::ORB_Search_by_projection_project_on_frame, orb_cuda::ORB_compute_distances) are
// routed to the oracle's restatements of those kernels (orc_project_points / orc_hamming_pairs), which are pinned bit-for-bit
// against the reference's kernels on a B200 (tests/test_helpers.py).
#pragma once
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

#include "../../jsfe_oracle.h"

using namespace std;

#define FRAME_GRID_ROWS 48   /* include/Frame.h:46-47 */
#define FRAME_GRID_COLS 64

namespace cv {
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };

// dense float matrix with value semantics for the handful of operations the sliced lines use on the 4x4 pose
class Mat {
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    size_t step[2] = {0, 0};
    Mat() {}
    Mat(int r, int c) : rows(r), cols(c), buf_(new vector<float>((size_t)r * c, 0.f)) { bind(); }
    static Mat bytes(unsigned char* p, int r, int c) { Mat m; m.rows = r; m.cols = c; m.data = p; m.step[0] = (size_t)c; m.step[1] = 1; return m; }
    float& f(int i, int j) { return (*buf_)[(size_t)i * cols + j]; }
    float f(int i, int j) const { return (*buf_)[(size_t)i * cols + j]; }
    template <typename T> T& at(int i) { return reinterpret_cast<T*>(data)[i]; }
    template <typename T> const T& at(int i) const { return reinterpret_cast<const T*>(data)[i]; }
    Mat rowRange(int a, int b) const { Mat m(b - a, cols); for (int i = a; i < b; ++i) for (int j = 0; j < cols; ++j) m.f(i - a, j) = f(i, j); return m; }
    Mat colRange(int a, int b) const { Mat m(rows, b - a); for (int i = 0; i < rows; ++i) for (int j = a; j < b; ++j) m.f(i, j - a) = f(i, j); return m; }
    Mat col(int c) const { return colRange(c, c + 1); }
    Mat t() const { Mat m(cols, rows); for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) m.f(j, i) = f(i, j); return m; }
private:
    shared_ptr<vector<float>> buf_;
    void bind() { data = reinterpret_cast<unsigned char*>(buf_->data()); step[0] = (size_t)cols * sizeof(float); step[1] = sizeof(float); }
};
inline Mat operator*(const Mat& a, const Mat& b) {     // cv::gemm accumulates CV_32F products in double
    Mat m(a.rows, b.cols);
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < b.cols; ++j) { double s = 0; for (int k = 0; k < a.cols; ++k) s += (double)a.f(i, k) * (double)b.f(k, j); m.f(i, j) = (float)s; }
    return m;
}
inline Mat operator+(const Mat& a, const Mat& b) { Mat m(a.rows, a.cols); for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) m.f(i, j) = a.f(i, j) + b.f(i, j); return m; }
inline Mat operator-(const Mat& a) { Mat m(a.rows, a.cols); for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) m.f(i, j) = -a.f(i, j); return m; }
}  // namespace cv

// include/cuda/synced_mem_holder.hpp:10-65 on the host only: resize keeps the buffer unless it must grow (src/cuda/synced_mem_holder.cpp:52-69)
template <typename T>
struct SyncedMem {
    int count_ = 0, capacity_ = 0;
    T* cpu_data_ = nullptr;
    ~SyncedMem() { delete[] cpu_data_; }
    void resize(int count) { count_ = count; if (capacity_ < count_) { capacity_ = count_; delete[] cpu_data_; cpu_data_ = new T[capacity_ > 0 ? capacity_ : 1]; } }
    T* cpu_data() { return cpu_data_; }
    T* gpu_data() { return cpu_data_; }
    void to_gpu_async() {} void to_cpu_async() {} void sync_stream() {} void to_gpu() {} void to_cpu() {}
};

namespace orb_cuda {
// include/cuda/orb_matcher.hpp:11-23 -> the oracle's restatements of the two kernels (pinned against the reference's kernels)
inline void ORB_Search_by_projection_project_on_frame(int n, float* Px, float* Py, float* Pz, float* Rcw, float* tcw, float fx, float fy, float cx,
                                                      float cy, float min_x, float max_x, float min_y, float max_y, float* u, float* v, float* invz,
                                                      unsigned char* is_valid) {
    orc_project_points(n, Px, Py, Pz, Rcw, tcw, fx, fy, cx, cy, min_x, max_x, min_y, max_y, u, v, invz, is_valid);
}
inline void ORB_compute_distances(int n, int* idx_last, int* idx_curr, unsigned char* desc_last, unsigned char* desc_curr, int* dist) {
    orc_hamming_pairs(n, idx_last, idx_curr, desc_last, desc_curr, dist);
}
}  // namespace orb_cuda
using namespace orb_cuda;

class MapPoint {   // include/MapPoint.h: the three accessors the sliced lines call
public:
    float x = 0, y = 0, z = 0;
    unsigned char desc[32];
    int nObs = 0;
    void GetWorldPosExp(float& px, float& py, float& pz) { px = x; py = y; pz = z; }
    void GetDescriptorExp(unsigned char* d) { memcpy(d, desc, 32); }
    int Observations() { return nObs; }
};

class Frame {      // include/Frame.h: members used by the sliced lines
public:
    cv::Mat mTcw;
    float mb = 0, mbf = 0;
    int N = 0;
    vector<MapPoint*> mvpMapPoints;
    vector<bool> mvbOutlier;
    vector<cv::KeyPoint> mvKeys, mvKeysUn;
    vector<float> mvScaleFactors, mvuRight;
    cv::Mat mDescriptors;
    static float fx, fy, cx, cy, mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    void GetFeaturesInArea(const float& x, const float& y, const float& invzc, const float& r, std::vector<int>& nbr_indices,
                           const int minLevel = -1, const int maxLevel = -1);   // defaults: include/Frame.h
};

class ORBmatcher {  // include/ORBmatcher.h:91-93 and src/ORBmatcher.cpp:36-38
public:
    static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;
    bool mbCheckOrientation = true;
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    void ComputeThreeMaxima(vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
};
