// jsfe_gather.cuh -- SURVEY.md 8(e): the results of a batch, trimmed to their keypoint counts, from every GPU to one rank.
// (included once by jsfe.cu)
//
// The reference has no multi-GPU path (device 0 is hard-wired, src/cuda/orb_gpu.cpp:24); this is the B200-side exchange step a
// batch consumer asks for (BASELINE config C5).  Two kernels, both on the gather stream of the rank that owns the results:
//   k_gather_pack  one block per slot: packs [n | kps 6 x n | desc 32 x n | (left eye) u_right n, depth n] of every slot of the
//                  batch, back to back, 16-byte aligned, into a staging region in LOCAL HBM (the compute stream only waits for
//                  this kernel before it may overwrite the results);
//   k_gather_put   copies the region -- header and payload, payload_bytes read from the header ON THE DEVICE, so the wire carries
//                  the trimmed size without a host round trip -- into this rank's region of the ROOT's landing buffer through a
//                  peer mapping (CUDA IPC): 16-byte stores over NVLink 5 / NVSwitch, no copy engine, no intermediate buffer.
// The only NCCL call in that mode is a 4-byte all-reduce behind the put: it tells the root that every region has landed and
// tells every rank that the root has released the landing buffer of two batches ago (credit for the double buffer).  Without
// peer mappings the regions travel as one ncclSend/ncclRecv group instead (padded to the capacity bound, sizes are not known on
// the host).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "jsfe_types.h"

namespace jsfe {

#define JSFE_GATHER_MAGIC 0x3147534A   // "JSG1"

// Region header (32 bytes), followed by int32 n_keypoints[2 * n_pairs], padded to 16 bytes, then the slot sections.
struct GatherHeader {
    int32_t magic, rank, n_pairs, capacity;
    int64_t payload_bytes;   // bytes after the header block (sum of the slot sections)
    int64_t sequence;        // batch counter of the sending rank
};

__host__ __device__ __forceinline__ size_t gather_header_bytes(int n_pairs) { return (size_t)((32 + 8 * n_pairs + 15) & ~15); }
__host__ __device__ __forceinline__ size_t gather_slot_bytes(int n, int left) { return (size_t)(((left ? 64 : 56) * n + 15) & ~15); }

__global__ void __launch_bounds__(256) k_gather_pack(const __grid_constant__ Params p, int first_pair, int n_pairs, int rank, long long sequence,
                                                     uint8_t* __restrict__ region) {
    __shared__ unsigned long long s_part[8];
    const int slot_in_batch = blockIdx.x;                 // 0 .. 2*n_pairs-1
    const int slot = 2 * first_pair + slot_in_batch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // section offset = sum of the sections in front of this slot (every block re-derives it: at most 2*n_pairs loads)
    unsigned long long before = 0, total = 0;
    for (int t = tid; t < 2 * n_pairs; t += 256) {
        const unsigned long long b = gather_slot_bytes(p.n_kp[2 * first_pair + t], (t & 1) == 0);
        if (t < slot_in_batch) before += b;
        total += b;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { before += __shfl_xor_sync(0xffffffffu, before, d); total += __shfl_xor_sync(0xffffffffu, total, d); }
    if (lane == 0) s_part[warp] = before;
    __syncthreads();
    before = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) before += s_part[w];
    __syncthreads();
    if (lane == 0) s_part[warp] = total;
    __syncthreads();
    total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) total += s_part[w];
    const size_t hdr = gather_header_bytes(n_pairs);
    if (slot_in_batch == 0) {                             // header + the keypoint counts
        if (tid == 0) {
            GatherHeader h;
            h.magic = JSFE_GATHER_MAGIC; h.rank = rank; h.n_pairs = n_pairs; h.capacity = p.cap;
            h.payload_bytes = (long long)total; h.sequence = sequence;
            *reinterpret_cast<GatherHeader*>(region) = h;
        }
        int32_t* cnt = reinterpret_cast<int32_t*>(region + 32);
        for (int t = tid; t < 2 * n_pairs; t += 256) cnt[t] = p.n_kp[2 * first_pair + t];
        for (int t = 32 + 8 * n_pairs + 4 * tid; t < (int)hdr; t += 1024) *reinterpret_cast<int32_t*>(region + t) = 0;   // pad words
    }
    const int n = p.n_kp[slot];
    const bool left = (slot_in_batch & 1) == 0;
    uint32_t* dst = reinterpret_cast<uint32_t*>(region + hdr + before);
    const size_t cap = (size_t)p.cap;
    const uint32_t* kps = reinterpret_cast<const uint32_t*>(p.kps + (size_t)slot * 6 * cap);
    for (int pl = 0; pl < 6; ++pl)
        for (int i = tid; i < n; i += 256) dst[pl * n + i] = kps[pl * cap + i];
    const uint32_t* desc = reinterpret_cast<const uint32_t*>(p.desc + (size_t)slot * cap * 32);
    for (int i = tid; i < 8 * n; i += 256) dst[6 * n + i] = desc[i];
    int words = 14 * n;
    if (left) {
        const uint32_t* ur = reinterpret_cast<const uint32_t*>(p.u_right + (size_t)slot * cap);
        const uint32_t* dp = reinterpret_cast<const uint32_t*>(p.depth + (size_t)slot * cap);
        for (int i = tid; i < n; i += 256) { dst[14 * n + i] = ur[i]; dst[15 * n + i] = dp[i]; }
        words = 16 * n;
    }
    const int padded = (int)(gather_slot_bytes(n, left) >> 2);
    for (int i = words + tid; i < padded; i += 256) dst[i] = 0;   // alignment words: the gathered bytes are deterministic
}

// src: a packed region in local HBM; dst: the same region in the root's landing buffer (peer mapping).  Grid-stride 16-byte copy
// of header + payload; the grid is sized for the capacity bound, blocks beyond the payload leave at once.
__global__ void __launch_bounds__(256) k_gather_put(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n_pairs) {
    const GatherHeader* h = reinterpret_cast<const GatherHeader*>(src);
    const size_t bytes = gather_header_bytes(n_pairs) + (size_t)h->payload_bytes;
    const size_t nvec = bytes >> 4;                       // both terms are multiples of 16
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    // few blocks (the launch takes at most 32): four loads in flight per thread before their stores
    const size_t stride = (size_t)gridDim.x * 256;
    size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    for (; i + 3 * stride < nvec; i += 4 * stride) {
        const uint4 a = s[i], b = s[i + stride], c = s[i + 2 * stride], e = s[i + 3 * stride];
        d[i] = a; d[i + stride] = b; d[i + 2 * stride] = c; d[i + 3 * stride] = e;
    }
    for (; i < nvec; i += stride) d[i] = s[i];
}

}  // namespace jsfe
