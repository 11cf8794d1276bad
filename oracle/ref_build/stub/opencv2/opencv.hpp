// Minimal stand-in for <opencv2/opencv.hpp>, TEST INFRASTRUCTURE ONLY.
//
// The reference's src/cuda translation units include OpenCV only for (a) the
// cv::Mat / cv::KeyPoint value types at the ORB_GPU boundary, (b) mask loading
// (imread/cvtColor/resize/threshold) in the ORB_GPU constructor and (c) the
// cvFloor/cvCeil/cvRound helpers used to build the umax table.  OpenCV's C++
// headers are absent from this image, so the oracle build (oracle/ref_build/
// Makefile) supplies this stub via -I; the reference sources themselves are
// compiled unmodified from /root/reference.  Nothing in the shipped product
// includes this file.
#ifndef JSFE_ORACLE_OPENCV_STUB_HPP
#define JSFE_ORACLE_OPENCV_STUB_HPP

#include <cmath>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <memory>
#include <iostream>
#include <algorithm>
#include <climits>

#define CV_8UC1 0
#define CV_8UC3 16
#define CV_BGR2GRAY 6
#define CV_INTER_NN 0
#define CV_THRESH_BINARY 0

static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v)  { int i = (int)v; return i + (i < v); }
static inline int cvRound(double v) { return (int)lrint(v); }

namespace cv {

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
};

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
};
typedef Point_<float> Point2f;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};

// Single-channel 8-bit matrix with shared ownership (enough for the mask path
// and for handing an image to ORB_GPU::extract).
class Mat {
public:
    int rows, cols;
    unsigned char* data;
    Mat() : rows(0), cols(0), data(nullptr) {}
    Mat(int r, int c, int /*type*/) : rows(r), cols(c) {
        buf_.reset(new std::vector<unsigned char>((size_t)r * c));
        data = buf_->data();
    }
    Mat(int r, int c, int /*type*/, void* ext) : rows(r), cols(c), data((unsigned char*)ext) {}
    bool empty() const { return data == nullptr || rows * cols == 0; }
private:
    std::shared_ptr<std::vector<unsigned char> > buf_;
};

static inline Mat imread(const std::string&) { return Mat(); }   // no mask files in the oracle runs
static inline void cvtColor(const Mat& src, Mat& dst, int) { dst = src; }
static inline void resize(const Mat& src, Mat& dst, Size sz, double, double, int) {
    Mat out(sz.height, sz.width, CV_8UC1);
    // nearest neighbour, OpenCV convention: sx = min(floor(x * inv_scale), cols-1)
    const double fx = (double)src.cols / sz.width, fy = (double)src.rows / sz.height;
    for (int y = 0; y < sz.height; ++y) {
        int sy = std::min((int)std::floor(y * fy), src.rows - 1);
        for (int x = 0; x < sz.width; ++x) {
            int sx = std::min((int)std::floor(x * fx), src.cols - 1);
            out.data[(size_t)y * sz.width + x] = src.data[(size_t)sy * src.cols + sx];
        }
    }
    dst = out;
}
static inline void threshold(const Mat& src, Mat& dst, double th, double maxv, int) {
    Mat out(src.rows, src.cols, CV_8UC1);
    for (size_t i = 0; i < (size_t)src.rows * src.cols; ++i)
        out.data[i] = src.data[i] > th ? (unsigned char)maxv : 0;
    dst = out;
}

}  // namespace cv
#endif
