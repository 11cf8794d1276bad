// compat/cuda/orb_matcher.hpp -- header-compatible shim of the reference's include/cuda/orb_matcher.hpp:11-23 over the C ABI.
// Same signatures; like the reference (src/cuda/orb_matcher.cu:88,142) the calls are synchronous on return.
#ifndef JSFE_COMPAT_ORB_MATCHER_HPP
#define JSFE_COMPAT_ORB_MATCHER_HPP

#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime_api.h>
#include <jsfe.h>

namespace orb_cuda {

inline void ORB_Search_by_projection_project_on_frame(int n_points, float* Px_gpu, float* Py_gpu, float* Pz_gpu, float* Rcw_gpu,
                                                      float* tcw_gpu, float& fx, float& fy, float& cx, float& cy, float& minX,
                                                      float& maxX, float& minY, float& maxY, float* u_gpu, float* v_gpu,
                                                      float* invz_gpu, unsigned char* is_valid_gpu) {
    if (jsfe_project_points(n_points, Px_gpu, Py_gpu, Pz_gpu, Rcw_gpu, tcw_gpu, fx, fy, cx, cy, minX, maxX, minY, maxY, u_gpu, v_gpu,
                            invz_gpu, is_valid_gpu, nullptr) != 0) { fprintf(stderr, "jsfe: %s\n", jsfe_last_error()); abort(); }
    cudaStreamSynchronize(nullptr);
}

inline void ORB_compute_distances(int n_points, int* idx_left, int* idx_right, unsigned char* descriptor_left,
                                  unsigned char* descriptor_right, int* distance) {
    if (jsfe_hamming_pairs(n_points, idx_left, idx_right, descriptor_left, descriptor_right, distance, nullptr) != 0) {
        fprintf(stderr, "jsfe: %s\n", jsfe_last_error()); abort();
    }
    cudaStreamSynchronize(nullptr);
}

}  // namespace orb_cuda
#endif
