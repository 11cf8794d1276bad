// compat/cuda/tracking_gpu.hpp -- header-compatible shim of the reference's include/cuda/tracking_gpu.hpp:13-28 over the C ABI
// (the STATIC_MEM_IS_IN variant, which is the one the reference compiles).  Synchronous on return like
// src/cuda/tracking_isinfrustum.cu:150.
#ifndef JSFE_COMPAT_TRACKING_GPU_HPP
#define JSFE_COMPAT_TRACKING_GPU_HPP

#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime_api.h>
#include <jsfe.h>

namespace tracking_cuda {

#define STATIC_MEM_IS_IN

inline void compute_isInFrustum_GPU(int n_points, float* Px_gpu, float* Py_gpu, float* Pz_gpu, float* Pnx_gpu, float* Pny_gpu,
                                    float* Pnz_gpu, float* MaxDistance_gpu, float* invariance_maxDistance_gpu,
                                    float* invariance_minDistance_gpu, float* Rcw_gpu, float* tcw_gpu, float* Ow_gpu, float& fx,
                                    float& fy, float& cx, float& cy, int& minX, int& maxX, int& minY, int& maxY, int& nScaleLevels,
                                    float& logScaleFactor, float& viewCosAngle, float* invz_gpu, float* u_gpu, float* v_gpu,
                                    int* predictedlevel_gpu, float* viewCos_gpu, unsigned char* is_infrustum_gpu) {
    if (jsfe_in_frustum(n_points, Px_gpu, Py_gpu, Pz_gpu, Pnx_gpu, Pny_gpu, Pnz_gpu, MaxDistance_gpu, invariance_maxDistance_gpu,
                        invariance_minDistance_gpu, Rcw_gpu, tcw_gpu, Ow_gpu, fx, fy, cx, cy, minX, maxX, minY, maxY, nScaleLevels,
                        logScaleFactor, viewCosAngle, invz_gpu, u_gpu, v_gpu, predictedlevel_gpu, viewCos_gpu, is_infrustum_gpu,
                        nullptr) != 0) { fprintf(stderr, "jsfe: %s\n", jsfe_last_error()); abort(); }
    cudaStreamSynchronize(nullptr);
}

}  // namespace tracking_cuda
#endif
