// jsfe_kernels.cuh -- sm_100a kernels of the stereo front-end (included once by jsfe.cu).
//
// Float expressions whose rounding decides an output bit are written with __fmul_rn/__fmaf_rn/
// __fadd_rn/__fdiv_rn intrinsics (never contracted or reordered by nvcc) in the association the
// reference's kernels have on sm_100a; transcendental calls are the plain libdevice atan2f/cosf/sinf
// (no --use_fast_math), the same code the reference links.  Reference citations are relative to
// /root/reference.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "jsfe_types.h"

namespace jsfe {

// =================================================================================================
// K1  k_pyramid: every level l >= 1 is a bilinear resample of level 0.
//     replaces imresize_GPU_pitched (src/cuda/orb_pyramid.cu:18-68), one launch for all levels and
//     all image slots; one thread = 4 adjacent output pixels = one 32-bit store; pad bytes = 0.
// =================================================================================================
__global__ void __launch_bounds__(256) k_pyramid(const __grid_constant__ Params p, int slot0) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.pyr_groups_total) return;
    int l = 1;
    while (l + 1 < p.L && g >= p.pyr_group_start[l + 1]) ++l;
    const LevelGeom& lv = p.lv[l];
    const int slot = slot0 + blockIdx.y;
    const int gl = g - p.pyr_group_start[l];
    const int gpr = lv.pitch >> 2;  // groups per row
    const int y = gl / gpr;
    const int x4 = (gl - y * gpr) << 2;
    const uint8_t* __restrict__ src = p.lv[0].img + (size_t)slot * p.lv[0].slot_stride;
    const int sp = p.lv[0].pitch;
    const float s = lv.rscale;
    const float fy = __fmul_rn(s, (float)y);
    const int yt = (int)floorf(fy);
    const float wyt = __fsub_rn((float)(yt + 1), fy), wyb = __fsub_rn(1.0f, wyt);
    const uint8_t* r0 = src + (size_t)yt * sp;
    const uint8_t* r1 = r0 + sp;
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x4 + k;
        if (x < lv.w) {
            const float fx = __fmul_rn(s, (float)x);
            const int xl = (int)floorf(fx);
            const float wxl = __fsub_rn((float)(xl + 1), fx), wxr = __fsub_rn(1.0f, wxl);
            // FMUL,FMUL,FFMA,FFMA,FFMA,F2I.TRUNC -- the contraction nvcc emits for the reference expression
            float acc = __fmul_rn(__fmul_rn(wyt, wxr), (float)__ldg(r0 + xl + 1));
            acc = __fmaf_rn(__fmul_rn(wyt, wxl), (float)__ldg(r0 + xl), acc);
            acc = __fmaf_rn(__fmul_rn(wyb, wxl), (float)__ldg(r1 + xl), acc);
            acc = __fmaf_rn(__fmul_rn(wyb, wxr), (float)__ldg(r1 + xl + 1), acc);
            packed |= (__float2uint_rz(acc) & 0xFFu) << (8 * k);
        }
    }
    uint8_t* dst = lv.img + (size_t)slot * lv.slot_stride + (size_t)y * lv.pitch + x4;
    *reinterpret_cast<uint32_t*>(dst) = packed;
}

// =================================================================================================
// K2  k_fast_cells: FAST ring test + SAD score + fused 3x3 NMS + per-cell arg-max, one block per
//     group of adjacent NMS cells.  The int32 score map of the reference never exists in HBM.
//     replaces FASTComputeScoreGPU_patternSize_16_lookup_mask (src/cuda/orb_FAST_compute_score.cu:1412-1560)
//          and Tile_unrolling_reduction_kernel_v2        (src/cuda/orb_FAST_apply_NMS_G.cu:1178-1397)
// =================================================================================================
__device__ __forceinline__ int fast_score(const uint8_t* c, int pw, int t, const uint32_t* __restrict__ lut) {
    const int v = c[0], vt = v + t, v_t = v - t;
    const int p4 = c[3], p12 = c[-3];
    if (p4 <= vt && p4 >= v_t && p12 <= vt && p12 >= v_t) return 0;
    const int p0 = c[3 * pw], p8 = c[-3 * pw];
    if (p0 <= vt && p0 >= v_t && p8 <= vt && p8 >= v_t) return 0;
    int r[16];
    r[0] = p0; r[4] = p4; r[8] = p8; r[12] = p12;
    r[1] = c[3 * pw + 1];  r[2] = c[2 * pw + 2];   r[3] = c[pw + 3];
    r[5] = c[-pw + 3];     r[6] = c[-2 * pw + 2];  r[7] = c[-3 * pw + 1];
    r[9] = c[-3 * pw - 1]; r[10] = c[-2 * pw - 2]; r[11] = c[-pw - 3];
    r[13] = c[pw - 3];     r[14] = c[2 * pw - 2];  r[15] = c[3 * pw - 1];
    unsigned bright = 0, dark = 0;
    int sad = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        bright |= (unsigned)(r[k] > vt) << k;
        dark |= (unsigned)(r[k] < v_t) << k;
        sad += abs(r[k] - v);
    }
    const unsigned hit = ((__ldg(lut + (bright >> 5)) >> (bright & 31)) | (__ldg(lut + (dark >> 5)) >> (dark & 31))) & 1u;
    return hit ? sad : 0;
}

__global__ void __launch_bounds__(256) k_fast_cells(const __grid_constant__ Params p, int slot0) {
    extern __shared__ __align__(16) uint8_t smem[];
    int l = 0;
    while (l + 1 < p.L && (int)blockIdx.x >= p.lv[l + 1].block_offset) ++l;
    const LevelGeom& lv = p.lv[l];
    const int slot = slot0 + blockIdx.y;
    const int item = blockIdx.x - lv.block_offset;
    const int ty = item / lv.blocks_per_row;
    const int tx0 = (item - ty * lv.blocks_per_row) * lv.cells_per_block;
    const int ncells = min(lv.cells_per_block, lv.n_tile_w - tx0);
    const int X0 = tx0 * lv.tile_w, GW = ncells * lv.tile_w, y0 = ty * lv.tile_h;
    const int gx0 = ((X0 - 4) >> 4) << 4, gy0 = y0 - 4;
    const int PR = lv.tile_h + 8;
    const int PW = ((X0 + GW + 4 - gx0) + 15) & ~15;
    const int SW = (GW + 2 + 1) & ~1;  // score row length (u16 elements)
    uint8_t* pix = smem;
    uint16_t* sc = reinterpret_cast<uint16_t*>(smem + (size_t)PR * PW);
    const uint8_t* __restrict__ img = lv.img + (size_t)slot * lv.slot_stride;

    // stage the pixel tile (16-byte vectors; out-of-image = 0)
    {
        const int vpr = PW >> 4, nvec = PR * vpr;
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
            const int row = i / vpr, v = i - row * vpr;
            const int gy = gy0 + row, gx = gx0 + (v << 4);
            uint4 val = make_uint4(0, 0, 0, 0);
            if (gy >= 0 && gy < lv.h && gx >= 0 && gx < lv.pitch)
                val = __ldg(reinterpret_cast<const uint4*>(img + (size_t)gy * lv.pitch + gx));
            *reinterpret_cast<uint4*>(pix + (size_t)row * PW + (v << 4)) = val;
        }
    }
    __syncthreads();

    // scores on the cell group + 1 px halo
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const uint32_t* lut = p.tab->lut_bits;
    for (int ry = warp; ry < lv.tile_h + 2; ry += nwarps) {
        const int y = y0 - 1 + ry;
        const bool yin = (y >= JSFE_B) && (y < lv.h - JSFE_B);
        for (int rx = lane; rx < GW + 2; rx += 32) {
            const int x = X0 - 1 + rx;
            int s = 0;
            if (yin && x >= JSFE_B && x < lv.w - JSFE_B) {
                if (lv.mask == nullptr || lv.mask[(size_t)y * lv.pitch + x])
                    s = fast_score(pix + (size_t)(y - gy0) * PW + (x - gx0), PW, p.threshold, lut);
            }
            sc[ry * SW + rx] = (uint16_t)s;
        }
    }
    __syncthreads();

    // per-cell arg-max of NMS survivors under the reference's tie-break order:
    // (score desc, column priority of the smem tree asc, y-lane (y-y0)%T asc, y asc) packed into one key.
    const uint8_t* rank = p.tab->col_rank[l];
    const uint8_t* by_rank = p.tab->col_by_rank[l];
    const int ymin = max(y0, JSFE_B), ymax = min(y0 + lv.tile_h, lv.h - JSFE_B);
    for (int c = warp; c < ncells; c += nwarps) {
        const int x0c = X0 + c * lv.tile_w;
        const int cw = min(lv.tile_w, lv.w - x0c);
        unsigned best = 0;
        for (int y = ymin; y < ymax; ++y) {
            const uint16_t* row = sc + (y - (y0 - 1)) * SW + (x0c - (X0 - 1));
            for (int j = lane; j < cw; j += 32) {
                const int s = row[j];
                if (s == 0) continue;
                const uint16_t* up = row + j - SW;
                const uint16_t* dn = row + j + SW;
                const bool ok = s >= up[-1] && s >= up[0] && s >= up[1] && s >= row[j - 1] && s >= row[j + 1] &&
                                s >= dn[-1] && s >= dn[0] && s >= dn[1];
                if (ok) {
                    const int dy = y - y0;
                    const unsigned key = ((unsigned)s << 18) | ((127u - rank[j]) << 11) | ((7u - (unsigned)(dy % lv.T)) << 8) |
                                         (255u - (unsigned)dy);
                    best = max(best, key);
                }
            }
        }
        best = __reduce_max_sync(0xffffffffu, best);
        if (lane == 0) {
            int bx = x0c, by = y0, bs = 0;
            if (best) {
                bs = (int)(best >> 18);
                bx = x0c + by_rank[127 - ((best >> 11) & 127u)];
                by = y0 + 255 - (int)(best & 255u);
            }
            const size_t o = (size_t)slot * p.cap + lv.cell_offset + ty * lv.n_tile_w + tx0 + c;
            p.cell_x[o] = bx;
            p.cell_y[o] = by;
            p.cell_s[o] = bs;
        }
    }
}

// =================================================================================================
// K3  k_compact: ordered stream compaction of the per-cell candidates (level-major, cell row-major),
//     on the device.  replaces the D2H -> host loop -> H2D bounce of ORB_GPU::FAST_obtain_keypoints
//     (src/cuda/orb_FAST_obtain_keypoints.cpp:12-56).  One block per slot.
//     Also records, for the stereo matcher, the first keypoint index of every (level, tile row).
// =================================================================================================
__global__ void __launch_bounds__(1024) k_compact(const __grid_constant__ Params p, int slot0) {
    __shared__ int warp_tot[32];
    __shared__ int s_base;
    const int slot = slot0 + blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int* cs = p.cell_s + (size_t)slot * p.cap;
    const int* cx = p.cell_x + (size_t)slot * p.cap;
    const int* cy = p.cell_y + (size_t)slot * p.cap;
    int* row_start = p.row_start + (size_t)slot * (p.n_tile_rows + 1);
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < p.cap; c0 += blockDim.x) {
        const int c = c0 + threadIdx.x;
        int flag = 0, l = 0, s = 0;
        if (c < p.cap) {
            while (l + 1 < p.L && c >= p.lv[l + 1].cell_offset) ++l;
            s = cs[c];
            flag = s > 0;
        }
        const unsigned b = __ballot_sync(0xffffffffu, flag);
        const int wpre = __popc(b & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(b);
        __syncthreads();
        if (warp == 0) {
            int v = warp_tot[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, v, d);
                if (lane >= d) v += n;
            }
            warp_tot[lane] = v;  // inclusive
        }
        __syncthreads();
        const int base = s_base;
        const int excl = base + (warp ? warp_tot[warp - 1] : 0) + wpre;
        if (c < p.cap) {
            const int rel = c - p.lv[l].cell_offset;
            if (rel % p.lv[l].n_tile_w == 0) row_start[p.lv[l].tile_row_offset + rel / p.lv[l].n_tile_w] = excl;
            if (flag) {
                const size_t o = (size_t)slot * p.cap + excl;
                p.kp_x[o] = cx[c];
                p.kp_y[o] = cy[c];
                p.kp_s[o] = s;
                p.kp_l[o] = l;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        row_start[p.n_tile_rows] = s_base;
        p.n_kp[slot] = s_base;
    }
    __syncthreads();
    if (threadIdx.x < p.L) {
        const int l = threadIdx.x;
        const int a = row_start[p.lv[l].tile_row_offset];
        const int e = (l + 1 < p.L) ? row_start[p.lv[l + 1].tile_row_offset] : s_base;
        p.n_per_level[(size_t)slot * JSFE_MAXL + l] = e - a;
    }
}

// =================================================================================================
// K4  k_orient_desc: one warp per keypoint.  Stages the 43x43 level-image patch in shared memory,
//     computes the intensity-centroid angle, evaluates the 7x7 blur ONLY at the 512 rotated sample
//     points (same 49-FFMA chain, so bit-identical to blurring the whole level), forms the 256-bit
//     descriptor and writes the final output planes.
//     replaces FASTComputeOrientationGPU (src/cuda/orb_FAST_orientation.cu:17-65), imgaussian_GPU
//     (src/cuda/orb_gaussian.cu:21-138), ORB_compute_descriptorGPU (src/cuda/orb_descriptor.cu:12-69),
//     ORB_copy_output_GPU (src/cuda/orb_copy_output.cu:12-45) and the D2D descriptor copies
//     (src/cuda/orb_gpu.cpp:819-831).
// =================================================================================================
__device__ __forceinline__ int blur_at(const uint8_t* pc, const float* __restrict__ gw) {
    // pc -> staged patch at the sample centre; 49 sequential FFMA, row-major taps, trunc to u8
    float acc = 0.0f;
#pragma unroll
    for (int i = -3; i <= 3; ++i) {
#pragma unroll
        for (int j = -3; j <= 3; ++j) acc = __fmaf_rn(gw[(i + 3) * 7 + (j + 3)], (float)pc[i * JSFE_PATCH_PITCH + j], acc);
    }
    return (int)(__float2uint_rz(acc) & 0xFFu);
}

__global__ void __launch_bounds__(256) k_orient_desc(const __grid_constant__ Params p, int slot0) {
    __shared__ __align__(16) uint8_t s_patch[8][JSFE_PATCH_ROWS * JSFE_PATCH_PITCH];
    __shared__ float s_gw[49];
    __shared__ int8_t s_px[512], s_py[512];
    const int slot = slot0 + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) { s_px[i] = p.tab->pat_x[i]; s_py[i] = p.tab->pat_y[i]; }
    if (threadIdx.x < 49) s_gw[threadIdx.x] = p.tab->gauss[threadIdx.x];
    __syncthreads();
    const int n = p.n_kp[slot];
    const int o = blockIdx.x * 8 + warp;
    if (o >= n) return;
    const size_t so = (size_t)slot * p.cap + o;
    const int x = p.kp_x[so], y = p.kp_y[so], l = p.kp_l[so], score = p.kp_s[so];
    const LevelGeom& lv = p.lv[l];
    const uint8_t* __restrict__ img = lv.img + (size_t)slot * lv.slot_stride;
    uint8_t* patch = s_patch[warp];
    // stage rows y-21..y+21, 48 bytes starting at the aligned address below x-21
    const int px0 = (x - JSFE_PATCH_R) & ~3, py0 = y - JSFE_PATCH_R;
    for (int i = lane; i < JSFE_PATCH_ROWS * (JSFE_PATCH_PITCH / 4); i += 32) {
        const int row = i / (JSFE_PATCH_PITCH / 4), wv = i - row * (JSFE_PATCH_PITCH / 4);
        const int gy = py0 + row, gx = px0 + 4 * wv;
        uint32_t v = 0;
        if (gy >= 0 && gy < lv.h && gx >= 0 && gx < lv.pitch) v = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)gy * lv.pitch + gx));
        *reinterpret_cast<uint32_t*>(patch + row * JSFE_PATCH_PITCH + 4 * wv) = v;
    }
    __syncwarp();
    const uint8_t* ctr = patch + JSFE_PATCH_R * JSFE_PATCH_PITCH + (x - px0);

    // intensity centroid over the radius-15 disc (integer moments; any summation order is exact)
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - 15, au = abs(u);
#pragma unroll 1
        for (int v = -15; v <= 15; ++v) {
            if (au <= p.tab->umax[abs(v)]) {
                const int I = ctr[v * JSFE_PATCH_PITCH + u];
                m10 += u * I;
                m01 += v * I;
            }
        }
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    const float angle = atan2f((float)m01, (float)m10);
    const float a = cosf(angle), b = sinf(angle);

    // descriptor byte `lane`: 8 comparisons of blurred samples
    unsigned val = 0;
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        int t[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int pi = lane * 16 + 2 * i + k;
            const float fpx = (float)s_px[pi], fpy = (float)s_py[pi];
            const int row = (int)rintf(__fmaf_rn(b, fpx, __fmul_rn(a, fpy)));
            const int col = __float2int_rn(__fmaf_rn(a, fpx, -__fmul_rn(b, fpy)));
            const int sx = x + col, sy = y + row;
            // the reference blurs only [B, h-B) x [B, w-B); everything else of its blurred image is 0
            int tv = 0;
            if (sx >= JSFE_B && sx < lv.w - JSFE_B && sy >= JSFE_B && sy < lv.h - JSFE_B)
                tv = blur_at(ctr + row * JSFE_PATCH_PITCH + col, s_gw);
            t[k] = tv;
        }
        val |= (unsigned)(t[0] < t[1]) << i;
    }
    p.desc[so * 32 + lane] = (uint8_t)val;

    if (lane == 0) {
        int* kp = p.kps + (size_t)slot * 6 * p.cap;
        const float sc = lv.scale;
        kp[0 * p.cap + o] = __float2int_rz(__fmul_rn((float)x, sc));
        kp[1 * p.cap + o] = __float2int_rz(__fmul_rn((float)y, sc));
        kp[2 * p.cap + o] = score;
        kp[3 * p.cap + o] = __float_as_int((float)((double)angle * (180.0 / 3.14159265358979323846)));
        kp[4 * p.cap + o] = l;
        kp[5 * p.cap + o] = __float2int_rz(__fmul_rn(31.0f, sc));
        p.kp_angle[so] = angle;
    }
}

// =================================================================================================
// K5  k_stereo_match: one warp per left keypoint.  Candidate right keypoints come straight from the
//     compacted list via the (level, tile row) index of k_compact -- no host row table, no candidate
//     pair list, no distance vector in HBM.  Hamming arg-min (popc + packed (dist,idx) min), SAD strip
//     with integer accumulation and warp reduction, parabola, disparity/depth.
//     replaces the host loops + ORBGetDistanceStereoGPU + Compute_L1_distance_GPU + cublasSgemv of
//     ORB_GPU::ORB_compute_stereo_match (src/cuda/orb_stereo_match.cu:105-561).
// =================================================================================================
__global__ void __launch_bounds__(256) k_stereo_match(const __grid_constant__ Params p, int pair0, int th_high, int th_low,
                                                      float mb, float mbf) {
    const int pair = pair0 + blockIdx.y;
    const int sl = 2 * pair, sr = sl + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    const int nL = p.n_kp[sl];
    if (i >= nL) return;
    const int cap = p.cap;
    const int* kL = p.kps + (size_t)sl * 6 * cap;
    const int* kR = p.kps + (size_t)sr * 6 * cap;
    const size_t oL = (size_t)sl * cap + i;
    const int XL = kL[i], YL = kL[cap + i], lvl = kL[4 * cap + i];
    const float uL = (float)XL;
    const float maxD = __fdiv_rn(mbf, mb);
    const float minU = __fsub_rn(uL, maxD), maxU = uL;

    // left descriptor in registers (every lane holds all 8 words)
    const uint4* dl = reinterpret_cast<const uint4*>(p.desc + oL * 32);
    const uint4 l0 = __ldg(dl), l1 = __ldg(dl + 1);
    const uint8_t* descR = p.desc + (size_t)sr * cap * 32;
    const int* rs = p.row_start + (size_t)sr * (p.n_tile_rows + 1);

    unsigned best = ((unsigned)th_high << 16) | 0xFFFFu;  // strict-min scan from TH_HIGH; ties -> lowest right index
    for (int lr = max(0, lvl - 1); lr <= min(p.L - 1, lvl + 1); ++lr) {
        const LevelGeom& g = p.lv[lr];
        const float rho = __fmul_rn(2.0f, g.scale);
        // conservative tile-row window of level lr that can reach row YL (exact test below)
        int ylo = (int)floorf(((float)YL - rho - 1.0f) / g.scale) - 1;
        int yhi = (int)ceilf(((float)YL + rho + 2.0f) / g.scale) + 1;
        ylo = max(ylo, 0);
        yhi = min(yhi, g.h - 1);
        if (ylo > yhi) continue;
        const int r0 = rs[g.tile_row_offset + ylo / g.tile_h];
        const int r1 = rs[g.tile_row_offset + yhi / g.tile_h + 1];
        for (int r = r0 + lane; r < r1; r += 32) {
            const float yR = (float)kR[cap + r];
            const int maxr = (int)ceilf(__fadd_rn(yR, rho)), minr = (int)floorf(__fsub_rn(yR, rho));
            if (YL < minr || YL > maxr) continue;
            const float uR = (float)kR[r];
            if (!(uR >= minU && uR <= maxU)) continue;
            const uint4* dr = reinterpret_cast<const uint4*>(descR + (size_t)r * 32);
            const uint4 a = __ldg(dr), b = __ldg(dr + 1);
            const int d = __popc(l0.x ^ a.x) + __popc(l0.y ^ a.y) + __popc(l0.z ^ a.z) + __popc(l0.w ^ a.w) +
                          __popc(l1.x ^ b.x) + __popc(l1.y ^ b.y) + __popc(l1.z ^ b.z) + __popc(l1.w ^ b.w);
            best = min(best, ((unsigned)d << 16) | (unsigned)r);
        }
    }
    best = __reduce_min_sync(0xffffffffu, best);
    const int bestD = (int)(best >> 16);
    const int bestIdx = (bestD < th_high) ? (int)(best & 0xFFFFu) : -1;
    float out_u = -1.0f, out_d = -1.0f;
    int out_sad = -1;
    const int th = (th_high + th_low) / 2;
    if (bestIdx >= 0 && bestD < th) {
        const LevelGeom& g = p.lv[lvl];
        const float inv = g.inv_scale;
        const float suR0 = roundf(__fmul_rn((float)kR[bestIdx], inv));
        const float suL = roundf(__fmul_rn(uL, inv));
        const float svL = roundf(__fmul_rn((float)YL, inv));
        if (!(suR0 - 10.0f < 0.0f || suR0 + 10.0f >= (float)g.w)) {
            const uint8_t* imL = g.img + (size_t)sl * g.slot_stride + (size_t)(int)svL * g.pitch + (int)suL;
            const uint8_t* imR = g.img + (size_t)sr * g.slot_stride + (size_t)(int)svL * g.pitch + (int)suR0;
            const int lc = imL[0];
            int rc[11];
#pragma unroll
            for (int s = 0; s < 11; ++s) rc[s] = imR[s - 5];
            int acc[11];
#pragma unroll
            for (int s = 0; s < 11; ++s) acc[s] = 0;
            for (int q = lane; q < 121; q += 32) {
                const int dy = q / 11 - 5, dx = q % 11 - 5;
                const int lvv = (int)imL[dy * g.pitch + dx] - lc;
                const uint8_t* rr = imR + dy * g.pitch + dx;
#pragma unroll
                for (int s = 0; s < 11; ++s) acc[s] += abs(lvv - ((int)rr[s - 5] - rc[s]));
            }
#pragma unroll
            for (int s = 0; s < 11; ++s) acc[s] = __reduce_add_sync(0xffffffffu, acc[s]);
            int bd = acc[0], bR = 0;
#pragma unroll
            for (int s = 1; s < 11; ++s)
                if (acc[s] < bd) { bd = acc[s]; bR = s; }
            if (bR != 0 && bR != 10) {
                float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                for (int s = 1; s < 10; ++s)
                    if (s == bR) { d1 = (float)acc[s - 1]; d2 = (float)acc[s]; d3 = (float)acc[s + 1]; }
                const float num = __fsub_rn(d1, d3);
                const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2)));
                const float deltaR = __fdiv_rn(num, den);
                if (!(deltaR < -1.0f || deltaR > 1.0f)) {
                    float bestuR = __fmul_rn(g.scale, __fadd_rn(__fsub_rn(__fadd_rn(suR0, (float)bR), 5.0f), deltaR));
                    float disparity = __fsub_rn(uL, bestuR);
                    if (disparity >= 0.0f && disparity < maxD) {
                        if (disparity <= 0.0f) {
                            disparity = (float)0.01;
                            bestuR = (float)((double)uL - 0.01);
                        }
                        out_d = __fdiv_rn(mbf, disparity);
                        out_u = bestuR;
                        out_sad = bd;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        p.best_idx[oL] = bestIdx;
        p.best_dist[oL] = (bestIdx >= 0) ? bestD : th_high;
        p.u_right[oL] = out_u;
        p.depth[oL] = out_d;
        p.sad_best[oL] = out_sad;
    }
}

// K6  k_stereo_outlier: median-of-SAD cut (src/cuda/orb_stereo_match.cu:565-578): drop matches whose SAD
//     minimum is >= 1.5*1.4*median, median = element n/2 of the ascending list.  Two-pass 8-bit radix
//     select in shared memory, one block per pair.
__global__ void __launch_bounds__(1024) k_stereo_outlier(const __grid_constant__ Params p, int pair0) {
    __shared__ int hist[256];
    __shared__ int s_n, s_hi, s_k2, s_med;
    const int sl = 2 * (pair0 + blockIdx.x);
    const int nL = p.n_kp[sl];
    const int* sad = p.sad_best + (size_t)sl * p.cap;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < nL; i += blockDim.x) {
        const int v = sad[i];
        if (v >= 0) { ++cnt; atomicAdd(&hist[(v >> 8) & 255], 1); }
    }
    if (cnt) atomicAdd(&s_n, cnt);
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;  // the reference is undefined here (reads vDistIdx[0] of an empty vector)
    if (threadIdx.x == 0) {
        int k = n / 2, b = 0;
        while (k >= hist[b]) { k -= hist[b]; ++b; }
        s_hi = b;
        s_k2 = k;
    }
    __syncthreads();
    const int hi = s_hi;
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < nL; i += blockDim.x) {
        const int v = sad[i];
        if (v >= 0 && ((v >> 8) & 255) == hi) atomicAdd(&hist[v & 255], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int k = s_k2, b = 0;
        while (k >= hist[b]) { k -= hist[b]; ++b; }
        s_med = (hi << 8) | b;
    }
    __syncthreads();
    const float median = (float)s_med;
    const float thDist = __fmul_rn(1.5f * 1.4f, median);
    for (int i = threadIdx.x; i < nL; i += blockDim.x) {
        const int v = sad[i];
        if (v >= 0 && !((float)v < thDist)) {
            p.u_right[(size_t)sl * p.cap + i] = -1.0f;
            p.depth[(size_t)sl * p.cap + i] = -1.0f;
        }
    }
}

// K7  k_pack: reference output layout (6 planes with stride N, descriptors 32N) into caller buffers
//     (src/cuda/orb_gpu.cpp:784-831); used by the C++ compat shim and jsfe_get_keypoints.
__global__ void k_pack(const __grid_constant__ Params p, int slot, int n, int* dst_kps, uint8_t* dst_desc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 6 * n && dst_kps) {
        const int plane = i / n, j = i - plane * n;
        dst_kps[i] = p.kps[(size_t)slot * 6 * p.cap + (size_t)plane * p.cap + j];
    }
    if (i < 8 * n && dst_desc)
        reinterpret_cast<uint32_t*>(dst_desc)[i] = reinterpret_cast<const uint32_t*>(p.desc + (size_t)slot * p.cap * 32)[i];
}

}  // namespace jsfe
