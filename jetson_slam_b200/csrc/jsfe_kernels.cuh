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

// ---- TMA / mbarrier helpers (sm_90+ PTX; SASS: UTMALDG + SYNCS) ---------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 3-D tiled load: box at element coordinates (x, y, z) of the tensor described by `map` -> dense rows in shared memory
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// ---- programmatic dependent launch (sm_90+): every kernel of the per-frame chain lets its successor in the stream start as soon as
// all of its own blocks are running (launch_dependents at entry) and itself waits for the completion -- and the memory -- of its
// predecessor right before it first touches global memory.  Between the 8 small kernels of a single stereo pair this hides the
// launch latency of each edge; launched without the attribute (or behind a copy / an event) both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// =================================================================================================
// K1  k_pyramid: every level l >= 1 is a bilinear resample of level 0.
//     replaces imresize_GPU_pitched (src/cuda/orb_pyramid.cu:18-68): one launch for all levels and all
//     image slots; a block is a 128 x 32 pixel tile of ONE level (level/tile decode is block-uniform); a thread owns
//     4 adjacent columns (their x terms are computed once) and 4 rows; 4 pixels = one 32-bit store; pad bytes
//     [w, pitch) are written as 0.
//     u8->f32 uses the 2^23 magic (LOP + FADD) so that only floor/trunc (F2I) land on the quarter-rate XU pipe.
// =================================================================================================
__device__ __forceinline__ float u8_to_f32(unsigned v) { return __uint_as_float(v | 0x4B000000u) - 8388608.0f; }

// K0  k_repitch: device-side copy of n images with contiguous rows (row pitch = w, as a host buffer has them) into the
//     16-byte-pitched level-0 slots; pad bytes [w, pitch) are written as 0.  A kernel rather than cudaMemcpy2DAsync so that
//     it never queues behind another batch's long host-to-device copy on a copy engine (two batches in flight).
//     Thread = one 16-byte chunk of a destination row; the source is read as aligned 32-bit words and funnel-shifted.
__global__ void __launch_bounds__(256) k_repitch(const uint8_t* __restrict__ src, long long src_row_pitch, long long src_image_stride,
                                                 const uint8_t* __restrict__ src_begin, const uint8_t* __restrict__ src_end, uint8_t* __restrict__ dst, int pitch,
                                                 unsigned long long slot_stride, int h, int w) {
    pdl_launch_dependents();
    pdl_wait();
    const int cpr = pitch >> 4;                                   // chunks per row
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= h * cpr) return;
    const int row = id / cpr, c = id - row * cpr, x = c << 4;
    const uint8_t* s = src + (size_t)blockIdx.y * src_image_stride + (size_t)row * src_row_pitch + x;
    const int nb = min(16, w - x);                                // valid bytes of this chunk (<= 0: pure padding)
    uint32_t o[4] = {0u, 0u, 0u, 0u};
    if (nb > 0) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(s);
        const uint32_t* q = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        if (reinterpret_cast<const uint8_t*>(q) >= src_begin && reinterpret_cast<const uint8_t*>(q + 5) <= src_end) {
            const unsigned sh = (unsigned)(a & 3u) * 8u;
            uint32_t wv[5];
#pragma unroll
            for (int k = 0; k < 5; ++k) wv[k] = __ldg(q + k);
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = __funnelshift_r(wv[k], wv[k + 1], sh);
        } else {                                                  // first / last words of the buffer: stay inside it
            for (int b = 0; b < nb; ++b) o[b >> 2] |= (uint32_t)__ldg(s + b) << (8 * (b & 3));
        }
        if (nb < 16) {                                            // zero the bytes beyond the image width
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int keep = nb - 4 * k;                      // bytes of word k that belong to the row
                o[k] = keep >= 4 ? o[k] : keep <= 0 ? 0u : (o[k] & ((1u << (8 * keep)) - 1u));
            }
        }
    }
    *reinterpret_cast<uint4*>(dst + (size_t)blockIdx.y * slot_stride + (size_t)row * pitch + x) = make_uint4(o[0], o[1], o[2], o[3]);
}

#define JSFE_PYR_ROWS 32   // rows per block tile (8 y-lanes x 4 rows each)

__global__ void __launch_bounds__(256) k_pyramid(const __grid_constant__ Params p, int slot0) {
    pdl_launch_dependents();
    const uint32_t item = __ldg(p.pyr_map + blockIdx.x);   // level << 28 | tile row << 14 | tile column (a table written at create time)
    pdl_wait();
    const int l = (int)(item >> 28);
    const LevelGeom& lv = p.lv[l];
    const int slot = slot0 + blockIdx.y;
    const int tyb = (int)((item >> 14) & 0x3fffu), txb = (int)(item & 0x3fffu);
    const int x4 = (txb << 7) + ((threadIdx.x & 31) << 2);
    if (x4 >= lv.pitch) return;
    const uint8_t* __restrict__ src = p.lv[0].img + (size_t)slot * p.lv[0].slot_stride;
    const int sp = p.lv[0].pitch;
    const float s = lv.rscale;
    // column terms are shared by every row this thread produces
    int xl[4];
    float wxl[4], wxr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float fx = __fmul_rn(s, (float)(x4 + k));
        xl[k] = (int)floorf(fx);
        wxl[k] = __fsub_rn((float)(xl[k] + 1), fx);
        wxr[k] = __fsub_rn(1.0f, wxl[k]);
        if (x4 + k >= lv.w) { xl[k] = 0; wxl[k] = 0.0f; wxr[k] = 0.0f; }   // pad columns: weights 0 -> byte 0
    }
    uint8_t* dst = lv.img + (size_t)slot * lv.slot_stride + x4;
    const int y_begin = tyb * JSFE_PYR_ROWS + (threadIdx.x >> 5);
#pragma unroll 2
    for (int y = y_begin; y < min(lv.h, (tyb + 1) * JSFE_PYR_ROWS); y += 8) {
        const float fy = __fmul_rn(s, (float)y);
        const int yt = (int)floorf(fy);
        const float wyt = __fsub_rn((float)(yt + 1), fy), wyb = __fsub_rn(1.0f, wyt);
        const uint8_t* r0 = src + (size_t)yt * sp;
        const uint8_t* r1 = r0 + sp;
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // FMUL,FMUL,FFMA,FFMA,FFMA,F2I.TRUNC -- the contraction nvcc emits for the reference expression
            float acc = __fmul_rn(__fmul_rn(wyt, wxr[k]), u8_to_f32(__ldg(r0 + xl[k] + 1)));
            acc = __fmaf_rn(__fmul_rn(wyt, wxl[k]), u8_to_f32(__ldg(r0 + xl[k])), acc);
            acc = __fmaf_rn(__fmul_rn(wyb, wxl[k]), u8_to_f32(__ldg(r1 + xl[k])), acc);
            acc = __fmaf_rn(__fmul_rn(wyb, wxr[k]), u8_to_f32(__ldg(r1 + xl[k] + 1)), acc);
            // trunc(acc) for 0 <= acc < 256 = the low mantissa byte of RZ(acc + 2^23): an FADD instead of F2I (quarter-rate pipe)
            packed |= (__float_as_uint(__fadd_rz(acc, 8388608.0f)) & 0xFFu) << (8 * k);
        }
        *reinterpret_cast<uint32_t*>(dst + (size_t)y * lv.pitch) = packed;
    }
}

// =================================================================================================
// K2  k_fast_cells: FAST ring test + SAD score + fused 3x3 NMS + per-cell arg-max.  One block owns a
//     group of adjacent NMS cells, stages their pixels (+4 px halo) in shared memory once and keeps the
//     scores there -- the reference's int32 score map (and its 9 reads per pixel) never exist in HBM.
//     replaces FASTComputeScoreGPU_patternSize_16_lookup_mask (src/cuda/orb_FAST_compute_score.cu:1412-1560)
//          and Tile_unrolling_reduction_kernel_v2              (src/cuda/orb_FAST_apply_NMS_G.cu:1178-1397)
//
//  FAST is integer-ALU-bound, not HBM-bound, so the work is cut before it is spread:
//   phase A  a CONSERVATIVE compass pre-test on 6-bit pixels, 8 pixels per thread and iteration.  With q = p >> 2,
//            p - v > t implies q_p - q_v >= m (m = (t-3)/4 + 1), and a byte lane q_p - q_v + (128 - m) can neither
//            overflow nor borrow, so "ring point brighter" for 4 pixels is ONE 32-bit add whose byte MSBs are the
//            flags (darker: the mirrored subtraction).  A pixel survives for polarity bright (dark) iff enough adjacent
//            compass points (ring 0,4,8,12) pass -- a necessary condition for any arc the LUT accepts, derived from
//            FAST_N_MIN and verified against the LUT at create time.  Flags of up to 8 row iterations are kept in
//            registers (one byte lane per pixel, shifted in from the MSB) and emitted ONCE per thread: warp scan of
//            the counts, one shared-memory atomic per warp, then a loop over the set bits.  Bright survivors fill the
//            work list from the front, dark survivors from the back (the position of an entry is its polarity).
//   phase B  the work list is evaluated densely and EXACTLY for the entry's polarity only: 16 ring bytes packed
//            4 per word, one packed compare per word, SAD by VABSDIFF4.ACC, the 16 flags merged into one
//            permuted index of the LUT bitmap.  Hits store their score and become positives: a warp writes them
//            over work-list entries it has already consumed, so the list only has to hold the candidates and is
//            sized for six resident blocks per SM (LevelGeom::fast_cap).
//   phase C  every warp walks its positives: 3x3 NMS test and one shared-memory atomicMax per cell on a key that
//            encodes the reference's tie-break order:
//            (score desc, column priority of the reference's smem tree asc, y-lane (y-y0)%T asc, y asc).
//  Every pixel rejected in phase A has score 0 in the reference too; phase B is the reference's arithmetic
//  (early-outs included), so scores are bit-identical.  A tile with more survivors than list slots (adversarial
//  or very textured tiles) falls back to a dense, exact evaluation of the whole tile.
// =================================================================================================
#ifndef JSFE_FAST_SWPAD
#define JSFE_FAST_SWPAD 0     // extra u16 columns per score row: shifts the shared-memory banks from row to row
#endif
// score row length (u16) of a k_fast_cells block that owns gw columns
__host__ __device__ __forceinline__ int fast_score_pitch(int gw) { return ((gw + 2 + 1) & ~1) + JSFE_FAST_SWPAD; }
#ifndef JSFE_FAST_PW
#define JSFE_FAST_PW 224   // shared-memory pitch of the pixel tile = TMA box width: covers floor16 slack 15 + 4 + 192 + 4 (+ pad)
#endif

// per-byte MSB = (a > b), unsigned bytes; nb7 = ~b & 0x7f7f7f7f (hoisted when b is loop-invariant)
__device__ __forceinline__ unsigned msb_gt(unsigned a, unsigned b, unsigned nb7) {
    const unsigned t = (a & 0x7f7f7f7fu) + nb7;
    return (a & ~b) | (~(a ^ b) & t);
}

__device__ __forceinline__ unsigned q6(unsigned w) { return (w >> 2) & 0x3f3f3f3fu; }   // 4 pixels -> 4 six-bit lanes

// (a & ~mask) | (b & mask) as ONE LOP3 (nvcc splits the two-constant form into two)
__device__ __forceinline__ unsigned bitsel(unsigned a, unsigned b, unsigned mask) {
    unsigned r;
    asm("lop3.b32 %0, %1, %2, %3, 0xD8;" : "=r"(r) : "r"(a), "r"(b), "r"(mask));
    return r;
}
__device__ __forceinline__ unsigned xor_forced(unsigned a, unsigned m) {   // keeps nvcc from turning x ^ (c ? ~0 : 0) into NOT + SEL
    unsigned r;
    asm("xor.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(m));
    return r;
}
// a * b + c on the FMA pipe (IMAD): byte packing there instead of PRMT keeps the ALU pipe, the kernel's limiter, free
__device__ __forceinline__ unsigned imad(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ unsigned pack4(unsigned b0, unsigned b1, unsigned b2, unsigned b3) {   // b0 | b1 << 8 | b2 << 16 | b3 << 24
    return imad(imad(b3, 256u, b2), 65536u, imad(b1, 256u, b0));
}
// shared-memory atomic add without nvcc's warp-aggregation preamble (the callers already elect one lane)
__device__ __forceinline__ unsigned atom_add_shared(unsigned* addr, unsigned v) {
    unsigned r;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(smem_u32(addr)), "r"(v) : "memory");
    return r;
}

__device__ __forceinline__ unsigned sad4_acc(unsigned a, unsigned b, unsigned acc) {    // acc + sum of |a_k - b_k| over the 4 bytes
    unsigned r;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
    return r;
}

// phase B for ONE work-list entry: exact FAST decision of ONE polarity (DARK = 0: ring point brighter than v + t, 1: darker than
// v - t) for the pixel whose centre byte c points at in the staged tile (pitch JSFE_FAST_PW); returns the SAD score or 0.
// The bytes are packed 4 ring points per word with IMADs (FMA pipe; the ALU pipe is this kernel's limiter), compared with one
// packed compare per word against the saturated bound (p > min(v+t,255), resp. p < max(v-t,0): saturation never changes the
// outcome), and the 16 flags are merged into one index of the permuted LUT bitmap.
template <int DARK>
__device__ __forceinline__ unsigned fast_eval(const uint8_t* __restrict__ c, int threshold, int compass_mode, const uint32_t* __restrict__ lutp) {
    constexpr int PW = JSFE_FAST_PW;
    const int v = c[0];
    // ring point k -> byte k%4 of word k/4   (offsets: orb_FAST_compute_score.cu:24-48)
    const unsigned R0 = pack4(c[3 * PW], c[3 * PW + 1], c[2 * PW + 2], c[PW + 3]);
    const unsigned R1 = pack4(c[3], c[-PW + 3], c[-2 * PW + 2], c[-3 * PW + 1]);
    const unsigned R2 = pack4(c[-3 * PW], c[-3 * PW - 1], c[-2 * PW - 2], c[-PW - 3]);
    const unsigned R3 = pack4(c[-3], c[PW - 3], c[2 * PW - 2], c[3 * PW - 1]);
    unsigned f0, f1, f2, f3;
    if (DARK) {
        const unsigned L4 = (unsigned)max(v - threshold, 0) * 0x01010101u, l7 = L4 & 0x7f7f7f7fu;
        auto lt = [&](unsigned r) { const unsigned t = l7 + (~r & 0x7f7f7f7fu); return (L4 & ~r) | (~(L4 ^ r) & t); };   // MSB: r < L
        f0 = lt(R0); f1 = lt(R1); f2 = lt(R2); f3 = lt(R3);
    } else {
        const unsigned H4 = (unsigned)min(v + threshold, 255) * 0x01010101u, nH7 = ~H4 & 0x7f7f7f7fu;
        f0 = msb_gt(R0, H4, nH7); f1 = msb_gt(R1, H4, nH7); f2 = msb_gt(R2, H4, nH7); f3 = msb_gt(R3, H4, nH7);
    }
    // merge the 16 flags (byte MSBs) into the high nibbles of one word: byte k = [ring k, 4+k, 8+k, 12+k | garbage]
    unsigned W = bitsel(f1 >> 1, f0, 0x80808080u);
    W = bitsel(f2 >> 2, W, 0xC0C0C0C0u);
    W = bitsel(f3 >> 3, W, 0xE0E0E0E0u);
    const unsigned u = bitsel(W >> 8, W >> 4, 0x0f0f0f0fu);
    const unsigned idx = __byte_perm(u, 0u, 0x4420);      // 16-bit index in the permuted bit order of DevTables::lut_perm
    unsigned hit = (__ldg(lutp + (idx >> 5)) >> (idx & 31u)) & 1u;
    if (compass_mode < 2) {
        // arcs shorter than 8 need not cover two adjacent compass points: apply the reference's early-outs explicitly
        // (orb_FAST_compute_score.cu:1452-1470): (4 and 12 both similar) or (0 and 8 both similar) -> score 0
        const int p0 = (int)(R0 & 255u), p4 = (int)(R1 & 255u), p8 = (int)(R2 & 255u), p12 = (int)(R3 & 255u);
        const bool d0 = abs(p0 - v) > threshold, d4 = abs(p4 - v) > threshold, d8 = abs(p8 - v) > threshold, d12 = abs(p12 - v) > threshold;
        if (!((d4 || d12) && (d0 || d8))) hit = 0u;
    }
    if (!hit) return 0u;
    const unsigned V4 = (unsigned)v * 0x01010101u;
    return sad4_acc(R3, V4, sad4_acc(R2, V4, sad4_acc(R1, V4, sad4_acc(R0, V4, 0u))));
}

// MODE = Params::compass_mode (0..3), HAS_MASK = the handle has a mask image: compile-time so that the 8-pixel loop carries
// neither the other modes' predicated-off formulas nor the mask pointer bookkeeping
template <int MODE, bool HAS_MASK>
__global__ void __launch_bounds__(256) k_fast_cells(const __grid_constant__ Params p, const __grid_constant__ TmaMaps tm, int slot0) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ unsigned s_best[192];
    __shared__ unsigned s_ncand;      // work-list counters, packed: bright | dark << 16
    __shared__ int s_npos[8];         // positives per warp (phase C walks a warp's own positives)
    constexpr int PW = JSFE_FAST_PW;
    pdl_launch_dependents();
    const uint32_t item = __ldg(p.fast_map + blockIdx.x);   // level << 28 | tile row << 14 | block in row
    const int l = (int)(item >> 28);
    const LevelGeom& lv = p.lv[l];
    const int slot = slot0 + blockIdx.y;
    const int ty = (int)((item >> 14) & 0x3fffu);
    const int tx0 = (int)(item & 0x3fffu) * lv.cells_per_block;
    const int ncells = min(lv.cells_per_block, lv.n_tile_w - tx0);
    const int X0 = tx0 * lv.tile_w, GW = ncells * lv.tile_w, y0 = ty * lv.tile_h;
    const int gx0 = ((X0 - 4) >> 4) << 4, gy0 = y0 - 4;
    const int PR = lv.tile_h + 8;
    const int SW = fast_score_pitch(GW);   // score row length (u16)
    const int SR = lv.tile_h + 2;
    const int cs0 = X0 - 1 - gx0;       // tile column of score column 0
    uint8_t* pix = smem;
    uint16_t* sc = reinterpret_cast<uint16_t*>(smem + (size_t)PR * PW);
    // work list of codes (score row << 8 | score column), lv.fast_cap slots (the host sizes it so that the wanted number of blocks
    // fits an SM; tiles with more survivors take the dense fallback): bright survivors from the front, dark ones from the back.
    // Phase B later overwrites entries it has consumed with the positives.
    uint16_t* cand = sc + (size_t)SR * SW;
    const int cap = lv.fast_cap;
    const uint8_t* __restrict__ img = lv.img + (size_t)slot * lv.slot_stride;
    const int tid = threadIdx.x, lane = tid & 31;
    // pixels whose score can be non-zero: interior [B, w-B) x [B, h-B) (score positions of this block: rows ry, columns rx)
    const int ry_lo = max(0, JSFE_B - (y0 - 1)), ry_hi = min(SR, lv.h - JSFE_B - (y0 - 1));
    const int xlo = max(X0 - 1, JSFE_B), xhi = min(X0 + GW, lv.w - JSFE_B - 1);  // inclusive valid x range

    // ---- stage the pixel tile: one TMA box load (out-of-image rows/columns arrive as 0); meanwhile clear the scores
    pdl_wait();                           // the level images are the predecessor's output
    if (p.use_tma) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            mbar_expect_tx(&s_bar, (uint32_t)(PR * PW));
            tma_load_3d(pix, &tm.tile[l], &s_bar, gx0, gy0, slot);
        }
    } else {
        constexpr int vpr = PW >> 4;
        const int nvec = PR * vpr;
        for (int i = tid; i < nvec; i += 256) {
            const int row = i / vpr, v = i - row * vpr;
            const int gy = gy0 + row, gx = gx0 + (v << 4);
            uint4 val = make_uint4(0, 0, 0, 0);
            if (gy >= 0 && gy < lv.h && gx >= 0 && gx < lv.pitch)
                val = __ldg(reinterpret_cast<const uint4*>(img + (size_t)gy * lv.pitch + gx));
            *reinterpret_cast<uint4*>(pix + (size_t)row * PW + (v << 4)) = val;
        }
    }
    {
        uint4* z = reinterpret_cast<uint4*>(sc);
        const int nz = (SR * SW + 7) >> 3;    // a ragged tail spills into the (still unused) work list
        for (int i = tid; i < nz; i += 256) z[i] = make_uint4(0, 0, 0, 0);
        if (tid < 192) s_best[tid] = 0;
        if (tid == 0) s_ncand = 0;
    }
    __syncthreads();                      // also makes the mbarrier init visible to every thread
    if (p.use_tma) mbar_wait(&s_bar, 0);  // TMA bytes have landed

    // ---- phase A: conservative compass pre-test on 6-bit pixels; thread = (column group g, row lane rl) on the level's
    //      fixed ngx x nrl grid, 8 pixels (two 32-bit words) per iteration, up to 8 iterations between two emissions
    {
        const int g0 = cs0 >> 3;                    // first 8-pixel column group that holds a score column
        const int nrl = lv.fast_nrl;
        const int rl = (int)(((unsigned)tid * lv.fast_ngx_inv) >> 16), g = tid - rl * lv.fast_ngx;
        const int c = (g0 + g) << 3, xb = gx0 + c;
        // column validity of this thread's 2 x 4 pixels (0xFF per valid byte), hoisted out of the row loop
        unsigned vxa = 0, vxb = 0;
        {
            const int lo = min(max(xlo - xb, 0), 8), hi = min(max(xhi - xb + 1, 0), 8);
            if (hi > lo && rl < nrl) {
                const unsigned long long m = (hi == 8 ? ~0ull : ((1ull << (8 * hi)) - 1ull)) & ~((1ull << (8 * lo)) - 1ull);
                vxa = (unsigned)m;
                vxb = (unsigned)(m >> 32);
            }
        }
        const unsigned C4 = (unsigned)(128 - p.fast_q_thresh) * 0x01010101u;
        int ry = ry_lo + rl;
        const uint8_t* rp = pix + (ry + 3) * PW + c;
        const uint8_t* mp = HAS_MASK ? lv.mask + (size_t)(y0 - 1 + ry) * lv.pitch + xb : nullptr;
        const int rstep = nrl * PW, cstep = nrl << 8;
        const size_t mstep = (size_t)nrl * lv.pitch;
        // one 4-pixel group on 6-bit lanes: V = centre pixels, P0/P8 = rows +3/-3, P4/P12 = columns +3/-3.  Returns the
        // bright flags in fb and the dark flags in fd (byte MSBs; other bits are garbage)
        auto compass = [&](unsigned V, unsigned P0, unsigned P4, unsigned P8, unsigned P12, unsigned& fb, unsigned& fd) {
            const unsigned A = C4 - V, B = C4 + V;
            const unsigned b0 = P0 + A, b4 = P4 + A, b8 = P8 + A, b12 = P12 + A;       // MSB: q_p - q_v >= m
            const unsigned d0 = B - P0, d4 = B - P4, d8 = B - P8, d12 = B - P12;       // MSB: q_v - q_p >= m
            if (MODE == 2) {         // an arc of >= 8 covers two ADJACENT compass points: one of {0,8} and one of {4,12}
                fb = (b0 | b8) & (b4 | b12);
                fd = (d0 | d8) & (d4 | d12);
            } else if (MODE == 3) {  // three adjacent compass points = at least three of the four
                fb = (b0 & b8 & (b4 | b12)) | (b4 & b12 & (b0 | b8));
                fd = (d0 & d8 & (d4 | d12)) | (d4 & d12 & (d0 | d8));
            } else {                 // no usable arc condition: only the reference's two early-outs, either polarity
                fb = fd = ((b4 | d4) | (b12 | d12)) & ((b0 | d0) | (b8 | d8));
            }
        };
        for (int r0 = ry_lo; r0 < ry_hi; r0 += 8 * nrl) {    // block-uniform trip counts (the emission is warp-collective)
            const int nit = min(8, (ry_hi - r0 + nrl - 1) / nrl);
            const int code_first = (ry << 8) + (c - cs0);     // work-list code of this thread's first pixel in this chunk
            unsigned aba = 0, abb = 0, ada = 0, adb = 0;      // flag accumulators: byte lane = pixel, newest iteration in the MSB
            for (int it = 0; it < nit; ++it, ry += nrl, rp += rstep) {
                unsigned fba = 0, fbb = 0, fda = 0, fdb = 0;
                if ((vxa | vxb) && ry < ry_hi) {
                    const unsigned Q0 = q6(*reinterpret_cast<const unsigned*>(rp - 4));
                    const uint2 W12 = *reinterpret_cast<const uint2*>(rp);          // c is a multiple of 8
                    const unsigned Q3 = q6(*reinterpret_cast<const unsigned*>(rp + 8));
                    const uint2 D = *reinterpret_cast<const uint2*>(rp + 3 * PW);    // ring 0  (0,+3)
                    const uint2 U = *reinterpret_cast<const uint2*>(rp - 3 * PW);    // ring 8  (0,-3)
                    const unsigned Q1 = q6(W12.x), Q2 = q6(W12.y);
                    compass(Q1, q6(D.x), __byte_perm(Q1, Q2, 0x6543), q6(U.x), __byte_perm(Q0, Q1, 0x4321), fba, fda);
                    compass(Q2, q6(D.y), __byte_perm(Q2, Q3, 0x6543), q6(U.y), __byte_perm(Q1, Q2, 0x4321), fbb, fdb);
                    if (HAS_MASK) {
                        const uint2 mw = __ldg(reinterpret_cast<const uint2*>(mp));
                        const unsigned ma = msb_gt(mw.x, 0u, 0x7f7f7f7fu), mb = msb_gt(mw.y, 0u, 0x7f7f7f7fu);
                        fba &= ma; fda &= ma; fbb &= mb; fdb &= mb;
                    }
                }
                if (HAS_MASK) mp += mstep;
                aba = bitsel(aba >> 1, fba, 0x80808080u);
                abb = bitsel(abb >> 1, fbb, 0x80808080u);
                ada = bitsel(ada >> 1, fda, 0x80808080u);
                adb = bitsel(adb >> 1, fdb, 0x80808080u);
            }
            // emission: the flag of iteration `it` sits at bit 8 - nit + it of its pixel's byte lane
            aba &= vxa; abb &= vxb; ada &= vxa; adb &= vxb;
            const int cb = __popc(aba) + __popc(abb), cd = __popc(ada) + __popc(adb);
            unsigned incl = (unsigned)cb | ((unsigned)cd << 16);
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned n = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += n;
            }
            const unsigned tot = __shfl_sync(0xffffffffu, incl, 31);
            if (tot) {
                unsigned base = 0;
                if (lane == 31) base = atom_add_shared(&s_ncand, tot);
                base = __shfl_sync(0xffffffffu, base, 31);
                int ob = (int)((base & 0xffffu) + (incl & 0xffffu)) - cb;     // first bright slot of this thread
                int od = (int)((base >> 16) + (incl >> 16)) - cd;            // first dark slot (counted from the back)
                const int code_base = code_first - (8 - nit) * cstep;
                // one entry per set bit
                uint16_t* pb = cand + ob;                 // bright: upwards from the front
                uint16_t* pd = cand + (cap - 1 - od);     // dark: downwards from the back
                // a list that does not fit is never read (the counters still say so): keep the stores inside the array
                if (ob + cb > cap) aba = abb = 0;
                if (od + cd > cap) ada = adb = 0;
                auto emit = [&](unsigned m, int col_off, uint16_t*& ptr, int dir) {
                    const int cb0 = code_base + col_off;
                    // lowest set bit first: the loop-carried chain is m &= m - 1, the bit index and its code offset
                    // (b & 7) rows of nrl, (b >> 3) columns are off it
                    while (m) {
                        const unsigned b = (unsigned)__ffs((int)m) - 1u;
                        m &= m - 1u;
                        *ptr = (uint16_t)(cb0 + (int)imad(b & 7u, (unsigned)cstep, b >> 3));
                        ptr += dir;
                    }
                };
                emit(aba, 0, pb, 1);
                emit(abb, 4, pb, 1);
                emit(ada, 0, pd, -1);
                emit(adb, 4, pd, -1);
            }
        }
    }
    __syncthreads();

    // ---- phase B: exact evaluation of the work list (one polarity per entry); hits store their score and become positives.
    //      Warp w evaluates entries [256 it + 32 w, +32) in its it-th pass and writes its q-th positive over the (q & 31)-th entry of
    //      its (q >> 5)-th pass: positives <= entries consumed, so they always fit, without atomics or a second array.
    const unsigned ncand2 = s_ncand;
    const int nb = (int)(ncand2 & 0xffffu), nd = (int)(ncand2 >> 16), ntot = nb + nd;
    const bool list_ok = ntot <= cap;
    auto slot_of = [&](int j) { return j < nb ? j : cap - 1 - (j - nb); };     // list index of the j-th entry
    {
        const uint32_t* __restrict__ lutp = p.tab->lut_perm;
        const uint8_t* cbase = pix + 3 * PW + cs0;
        if (list_ok) {
            const int nit = (ntot + 255) >> 8;        // warp-uniform trip count: the positives append is warp-collective
            const unsigned lt = (1u << lane) - 1u;
            const int wbase = tid - lane;             // first entry of this warp in pass 0
            int wpos = 0;                              // positives of this warp so far (warp-uniform)
            for (int it = 0, i = tid; it < nit; ++it, i += 256) {
                unsigned score = 0;
                int code = 0;
                const int w0 = i - lane;              // first entry of this warp: the polarity is warp-uniform except in one warp per block
                if (w0 + 31 < nb) {
                    code = cand[i];
                    score = fast_eval<0>(cbase + (code >> 8) * PW + (code & 255), p.threshold, MODE, lutp);
                } else if (w0 >= nb) {
                    if (i < ntot) {
                        code = cand[cap - 1 - (i - nb)];
                        score = fast_eval<1>(cbase + (code >> 8) * PW + (code & 255), p.threshold, MODE, lutp);
                    }
                } else if (i < ntot) {
                    const bool dark = i >= nb;
                    code = cand[dark ? cap - 1 - (i - nb) : i];
                    const uint8_t* c = cbase + (code >> 8) * PW + (code & 255);
                    score = dark ? fast_eval<1>(c, p.threshold, MODE, lutp) : fast_eval<0>(c, p.threshold, MODE, lutp);
                }
                if (score) sc[(code >> 8) * SW + (code & 255)] = (uint16_t)score;
                const unsigned pb = __ballot_sync(0xffffffffu, score != 0u);   // every lane has read its entry by now
                if (score != 0u) {
                    const int q = wpos + __popc(pb & lt);
                    cand[slot_of(((q >> 5) << 8) + wbase + (q & 31))] = (uint16_t)code;
                }
                wpos += __popc(pb);
            }
            if (lane == 0) s_npos[tid >> 5] = wpos;
        } else {
            // work list overflow (more survivors than score positions: adversarial input): every interior pixel, both polarities
            const int n = SR * SW;
            for (int i = tid; i < n; i += 256) {
                const int ry = i / SW, rx = i - ry * SW;
                const int x = X0 - 1 + rx;
                if (ry < ry_lo || ry >= ry_hi || x < xlo || x > xhi) continue;
                if (HAS_MASK && lv.mask[(size_t)(y0 - 1 + ry) * lv.pitch + x] == 0) continue;
                const uint8_t* c = cbase + ry * PW + rx;
                unsigned score = fast_eval<0>(c, p.threshold, MODE, lutp);
                if (!score) score = fast_eval<1>(c, p.threshold, MODE, lutp);
                if (score) sc[i] = (uint16_t)score;
            }
        }
    }
    __syncthreads();

    // ---- phase C: NMS + per-cell arg-max.  Every warp walks its own positives; if the work list overflowed, the whole score
    //      tile is walked instead (exact, slower: noise-like or adversarial tiles only)
    {
        const int ymin = max(y0, JSFE_B), ymax = min(y0 + lv.tile_h, lv.h - JSFE_B);
        const int wv = min(GW, lv.w - X0);          // owned columns actually inside the image
        // one candidate: 3x3 NMS (ties survive) and the packed arg-max key of its cell
        auto nms = [&](int ry, int rx, unsigned s) {
            const int dy = ry - 1, xx = rx - 1, y = y0 + dy;
            const uint16_t* row = sc + ry * SW + rx;
            const uint16_t* up = row - SW;
            const uint16_t* dn = row + SW;
            const unsigned m = max(max(max((unsigned)up[-1], (unsigned)up[0]), max((unsigned)up[1], (unsigned)row[-1])),
                                   max(max((unsigned)row[1], (unsigned)dn[-1]), max((unsigned)dn[0], (unsigned)dn[1])));
            if (s >= m && y >= ymin && y < ymax && xx >= 0 && xx < wv) {
                const unsigned ck = __ldg(&p.tab->colkey[l][xx]);
                const unsigned key = (s << 18) | ((ck >> 8) << 11) | (unsigned)__ldg(&p.tab->rowkey[l][dy]);
                atomicMax(&s_best[ck & 0xFFu], key);
            }
        };
        if (list_ok) {
            const int npos = s_npos[tid >> 5];
            for (int q = lane, j = tid; q < npos; q += 32, j += 256) {      // its q-th positive sits over its entry of pass q >> 5
                const int code = cand[slot_of(j)];
                const int ry = code >> 8, rx = code & 255;     // a positive: score > 0, row and column >= 1 by construction of phase A
                nms(ry, rx, sc[ry * SW + rx]);
            }
        } else {
            for (int ry = 1 + (tid >> 5); ry < SR - 1; ry += 8)       // warp = row, lane = column: no division
                for (int rx = 1 + lane; rx < SW - 1; rx += 32) {
                    const unsigned s = sc[ry * SW + rx];
                    if (s) nms(ry, rx, s);
                }
        }
    }
    __syncthreads();
    if (tid < ncells) {
        const uint8_t* by_rank = p.tab->col_by_rank[l];
        const unsigned best = s_best[tid];
        const int x0c = X0 + tid * lv.tile_w;
        int bx = x0c, by = y0, bs = 0;
        if (best) {
            bs = (int)(best >> 18);
            bx = x0c + by_rank[127 - ((best >> 11) & 127u)];
            by = y0 + 255 - (int)(best & 255u);
        }
        const size_t o = (size_t)slot * p.cap + lv.cell_offset + ty * lv.n_tile_w + tx0 + tid;
        p.cell_x[o] = bx;
        p.cell_y[o] = by;
        p.cell_s[o] = bs;
    }
}

// =================================================================================================
// K2c k_blur: 7x7 sigma=10 blur of every level (descriptor input), interior [20,h-20)x[20,w-20) only; the
//     rest of the blurred level stays 0 exactly like the reference's never-written border.
//     replaces imgaussian_GPU (src/cuda/orb_gaussian.cu:21-138).
//
//  The reference's value is trunc(chain of 49 sequential FFMA).  We evaluate the separable form
//  (7+7 FFMA per pixel; one thread owns 4 columns x 32 rows and keeps the row sums of the last 7 input rows
//  in a rotating register window), whose distance to the chain is bounded by 5.3e-4;
//  a pixel whose separable value lies within 18 x 2^-15 = 5.49e-4 of an integer is ambiguous and gets the
//  exact chain instead, so the stored byte is always the reference's.  The ambiguity flags never leave the
//  block: each thread keeps the 4-bit mask of every row it produced in its own 32 bytes of shared memory,
//  after the strip loop the flagged pixels (about 0.1 % on textured images, all of a flat region) are
//  compacted into a block-wide list and re-evaluated densely, one thread per listed pixel -- no mask in HBM,
//  no second kernel, no atomics inside the FFMA loop.  Input words come through L1 (each 32-bit word of a row
//  is shared by 3 neighbouring threads), bytes are widened with PRMT + FADD (0x4B000000 trick).
// =================================================================================================
#define JSFE_BLUR_AMB_UNITS 18u   // ambiguity margin in units of 2^-15: 18 * 2^-15 = 5.49e-4 >= the 5.3e-4 bound on |separable - chain|
#ifndef JSFE_BLUR_ROWS
#define JSFE_BLUR_ROWS 32
#endif
#ifndef JSFE_BLUR_BLOCKS
#define JSFE_BLUR_BLOCKS 3        // resident blocks per SM the register allocation aims at
#endif
#define JSFE_FIX_LIST 2048        // block-wide list of ambiguous pixels; beyond it (flat regions) the owner decides its pixels in place

// separable factors of the 7x7 weights; the same for every handle (sigma is fixed at 10 in the reference), kept in
// constant memory so the FFMAs read them as c[][] operands instead of holding 14 registers (74 -> ~60: 4 blocks/SM)
__constant__ float c_sep_a[7];
__constant__ float c_sep_b[7];

// the reference's blur value: 49 sequential FFMA in row-major tap order, truncated (orb_gaussian.cu:37-135)
__device__ __forceinline__ unsigned blur_exact(const uint8_t* __restrict__ pc, int pitch, const float* __restrict__ gw) {
    float acc = 0.0f;
#pragma unroll 1
    for (int i = -3; i <= 3; ++i) {
#pragma unroll
        for (int j = -3; j <= 3; ++j) acc = __fmaf_rn(__ldg(gw + (i + 3) * 7 + (j + 3)), (float)(unsigned)__ldg(pc + i * pitch + j), acc);
    }
    return __float2uint_rz(acc) & 0xFFu;
}

__device__ __forceinline__ float byte_f(unsigned w, unsigned sel) {   // exact u8 -> f32 without I2F
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - 8388608.0f;
}

__global__ void __launch_bounds__(256, JSFE_BLUR_BLOCKS) k_blur(const __grid_constant__ Params p, int slot0) {
    __shared__ __align__(16) uint8_t s_flag[256 * JSFE_BLUR_ROWS];   // [thread][row of its strip]: 4-bit ambiguity mask of its 4 pixels
    __shared__ unsigned s_list[JSFE_FIX_LIST];                       // level << 28 | y << 14 | x
    __shared__ int s_n;
    const int tid = threadIdx.x;
    const int gi = blockIdx.x * blockDim.x + tid;
    const bool active = gi < p.blur_items_total;
    const int slot = slot0 + blockIdx.y;
    uint4* myflags = reinterpret_cast<uint4*>(s_flag + tid * JSFE_BLUR_ROWS);
    myflags[0] = make_uint4(0, 0, 0, 0);
    myflags[1] = make_uint4(0, 0, 0, 0);
    if (tid == 0) s_n = 0;
    int l = 0, xg = 0, r0 = 0, xend = 0;
    if (active) {
        while (l + 1 < p.L && gi >= p.blur_item_start[l + 1]) ++l;
        const LevelGeom& lv = p.lv[l];
        const int ncg = (lv.w - 2 * JSFE_B + 3) >> 2;               // 4-column groups over [20, w-20)
        const int li = gi - p.blur_item_start[l];
        const int strip = li / ncg, cg = li - strip * ncg;
        xg = JSFE_B + (cg << 2);                                     // 20 is a multiple of 4: groups are word-aligned
        r0 = JSFE_B + strip * JSFE_BLUR_ROWS;
        const int r1 = min(r0 + JSFE_BLUR_ROWS, lv.h - JSFE_B);
        xend = lv.w - JSFE_B;
        const uint8_t* __restrict__ src = lv.img + (size_t)slot * lv.slot_stride;
        uint8_t* __restrict__ dst = lv.blur + (size_t)slot * lv.slot_stride;
        uint8_t* fl = s_flag + tid * JSFE_BLUR_ROWS;
        const int nin = r1 - r0 + 6;                                 // input rows r0-3 .. r1+2
        const uint8_t* rp = src + (size_t)(r0 - 3) * lv.pitch + xg;
        float q[4][7];
        for (int base = 0; base < nin; base += 7) {
            // issue the loads of the next 7 input rows back to back (21 independent 32-bit loads in flight per thread):
            // the rows arrive from L2/HBM while the other warps of the SM are in their FFMA phase
            unsigned wr[7][3];
#pragma unroll
            for (int ph = 0; ph < 7; ++ph) {
                if (base + ph < nin) {
                    const uint8_t* r = rp + (size_t)ph * lv.pitch;
                    wr[ph][0] = __ldg(reinterpret_cast<const unsigned*>(r - 4));
                    wr[ph][1] = __ldg(reinterpret_cast<const unsigned*>(r));
                    wr[ph][2] = __ldg(reinterpret_cast<const unsigned*>(r + 4));
                }
            }
            rp += (size_t)7 * lv.pitch;
#pragma unroll
            for (int ph = 0; ph < 7; ++ph) {
                const int ir = base + ph;
                if (ir < nin) {
                    const unsigned W0 = wr[ph][0], W1 = wr[ph][1], W2 = wr[ph][2];
                    float f[10];  // pixels xg-3 .. xg+6
                    f[0] = byte_f(W0, 0x7441); f[1] = byte_f(W0, 0x7442); f[2] = byte_f(W0, 0x7443);
                    f[3] = byte_f(W1, 0x7440); f[4] = byte_f(W1, 0x7441); f[5] = byte_f(W1, 0x7442); f[6] = byte_f(W1, 0x7443);
                    f[7] = byte_f(W2, 0x7440); f[8] = byte_f(W2, 0x7441); f[9] = byte_f(W2, 0x7442);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float r = 0.0f;
#pragma unroll
                        for (int j = 0; j < 7; ++j) r = __fmaf_rn(c_sep_b[j], f[k + j], r);
                        q[k][ph] = r;
                    }
                    if (ir >= 6) {
                        const int y = r0 + ir - 6;
                        unsigned out = 0, amb = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float A = 0.0f;
#pragma unroll
                            for (int j = 0; j < 7; ++j) A = __fmaf_rn(c_sep_a[j], q[k][(ph + 1 + j) % 7], A);
                            // A in [0, 256): RZ(A + 256) = 256 + floor(A * 2^15) / 2^15, so mantissa bits 15..22 are trunc(A) and bits
                            // 0..14 the fraction in units of 2^-15.  Within blur_amb_units of an integer the truncation of the
                            // reference's chain E may differ from trunc(A) (|A - E| <= 5.3e-4): flag the pixel.
                            const unsigned tb = __float_as_uint(__fadd_rz(A, 256.0f));
                            amb |= (((tb + p.blur_amb_units) & 0x7FFFu) < 2u * p.blur_amb_units ? 1u : 0u) << k;
                            out |= ((tb >> 15) & 0xFFu) << (8 * k);
                        }
                        fl[y - r0] = (uint8_t)amb;
                        uint8_t* o = dst + (size_t)y * lv.pitch + xg;
                        if (xg + 3 < xend) {
                            *reinterpret_cast<unsigned*>(o) = out;
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                if (xg + k < xend) o[k] = (uint8_t)(out >> (8 * k));
                        }
                    }
                }
            }
        }
    }
    __syncthreads();          // s_n = 0 is visible; the separable bytes of this block are ordered before the corrections below
    if (active) {
        const uint4 fa = myflags[0], fb = myflags[1];
        if (fa.x | fa.y | fa.z | fa.w | fb.x | fb.y | fb.z | fb.w) {
            const LevelGeom& lv = p.lv[l];
            const unsigned* fw = reinterpret_cast<const unsigned*>(myflags);
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                unsigned m = fw[k] & 0x0F0F0F0Fu;                // byte j = row 4k + j, low nibble = its 4 pixels
                while (m) {
                    const int b = __ffs((int)m) - 1;
                    m &= m - 1;
                    const int y = r0 + 4 * k + (b >> 3), x = xg + (b & 7);
                    if (x >= xend) continue;                      // pad columns of the last group
                    const int pos = atomicAdd(&s_n, 1);
                    if (pos < JSFE_FIX_LIST) {
                        s_list[pos] = ((unsigned)l << 28) | ((unsigned)y << 14) | (unsigned)x;
                    } else {                                       // list full (flat region: every lane is here): decide in place
                        const size_t o = (size_t)slot * lv.slot_stride + (size_t)y * lv.pitch + x;
                        lv.blur[o] = (uint8_t)blur_exact(lv.img + o, lv.pitch, p.tab->gauss);
                    }
                }
            }
        }
    }
    __syncthreads();
    const int n = min(s_n, JSFE_FIX_LIST);
    for (int i = tid; i < n; i += 256) {
        const unsigned code = s_list[i];
        const int ll = code >> 28, y = (code >> 14) & 0x3FFF, x = code & 0x3FFF;
        const LevelGeom& lv = p.lv[ll];
        const size_t o = (size_t)slot * lv.slot_stride + (size_t)y * lv.pitch + x;
        lv.blur[o] = (uint8_t)blur_exact(lv.img + o, lv.pitch, p.tab->gauss);
    }
}

// =================================================================================================
// K2b cross-scale NMS on the per-cell candidates (optional, apply_nms_ms && L > 1), one block per slot.
//
//  k_nms_ms_dense   -- the rule of the reference's GPU mode (src/cuda/orb_FAST_apply_NMS_MS.cu:18-467):
//     every candidate is projected to its level-0 pixel (Y,X) = trunc(y*s), trunc(x*s); per pixel the
//     scores over levels give sum and zeros = #levels without a candidate; a candidate survives iff
//     sum*zeros at its pixel is >= sum*zeros at all 8 neighbouring pixels.  The reference materialises an
//     L x H0 x W0 int32 volume (15 MB at KITTI size) and races on it; here the <= cap occupied pixels go
//     into an open-addressing hash table in L2-resident scratch and all reads happen after all writes
//     (the race-free two-phase semantics the oracle defines).
//  k_nms_ms_buckets -- the rule of the reference's CPU mode (src/cuda/orb_FAST_apply_NMS_MS.cpp:15-122):
//     candidates are bucketed by level-0 NMS tile of (trunc(x*s - 20), trunc(y*s - 20)); inside a bucket,
//     in (level, cell) insertion order, every ordered pair of different levels within +-1 px zeroes the
//     lower score (sequential, order-dependent -> one thread walks one bucket).
// =================================================================================================
__device__ __forceinline__ unsigned hash_px(int key) { return (unsigned)key * 2654435761u; }

__global__ void __launch_bounds__(1024) k_nms_ms_dense(const __grid_constant__ Params p, int slot0) {
    pdl_launch_dependents();
    pdl_wait();
    const int slot = slot0 + blockIdx.x;
    const int ts = p.ms_table_size, mask = ts - 1;
    int* keys = p.ms_keys + (size_t)slot * ts;
    int* sums = p.ms_sums + (size_t)slot * ts;
    int* cnts = p.ms_cnts + (size_t)slot * ts;
    int* cs = p.cell_s + (size_t)slot * p.cap;
    const int* cx = p.cell_x + (size_t)slot * p.cap;
    const int* cy = p.cell_y + (size_t)slot * p.cap;
    for (int i = threadIdx.x; i < ts; i += blockDim.x) { keys[i] = -1; sums[i] = 0; cnts[i] = 0; }
    __syncthreads();
    for (int c = threadIdx.x; c < p.cap; c += blockDim.x) {
        const int s = cs[c];
        if (!s) continue;
        int l = 0;
        while (l + 1 < p.L && c >= p.lv[l + 1].cell_offset) ++l;
        const int Y = __float2int_rz(__fmul_rn((float)cy[c], p.lv[l].scale));
        const int X = __float2int_rz(__fmul_rn((float)cx[c], p.lv[l].scale));
        const int key = Y * p.W0 + X;
        unsigned h = hash_px(key) & mask;
        for (;;) {
            const int prev = atomicCAS(&keys[h], -1, key);
            if (prev == -1 || prev == key) { atomicAdd(&sums[h], s); atomicAdd(&cnts[h], 1); break; }
            h = (h + 1) & mask;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.cap; c += blockDim.x) {
        const int s = cs[c];
        if (!s) continue;
        int l = 0;
        while (l + 1 < p.L && c >= p.lv[l + 1].cell_offset) ++l;
        const int Y = __float2int_rz(__fmul_rn((float)cy[c], p.lv[l].scale));
        const int X = __float2int_rz(__fmul_rn((float)cx[c], p.lv[l].scale));
        int prod[9];
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const int key = (Y + q / 3 - 1) * p.W0 + (X + q % 3 - 1);
            unsigned h = hash_px(key) & mask;
            int v = 0;
            for (;;) {
                const int k = keys[h];
                if (k == key) { v = sums[h] * (p.L - cnts[h]); break; }
                if (k == -1) break;
                h = (h + 1) & mask;
            }
            prod[q] = v;
        }
        bool valid = true;
#pragma unroll
        for (int q = 0; q < 9; ++q) valid &= prod[4] >= prod[q];
        if (!valid) p.ms_drop[(size_t)slot * p.cap + c] = 1; else p.ms_drop[(size_t)slot * p.cap + c] = 0;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < p.cap; c += blockDim.x)
        if (cs[c] && p.ms_drop[(size_t)slot * p.cap + c]) cs[c] = 0;
}

__global__ void __launch_bounds__(256) k_nms_ms_buckets(const __grid_constant__ Params p, int slot0) {
    pdl_launch_dependents();
    pdl_wait();
    extern __shared__ int s_bucket[];  // bucket id per cell (-1: no candidate)
    const int slot = slot0 + blockIdx.x;
    int* cs = p.cell_s + (size_t)slot * p.cap;
    const int* cx = p.cell_x + (size_t)slot * p.cap;
    const int* cy = p.cell_y + (size_t)slot * p.cap;
    int* bx = p.ms_keys + (size_t)slot * p.ms_table_size;  // scratch: projected x / y per cell
    int* by = p.ms_sums + (size_t)slot * p.ms_table_size;
    const int nb = p.lv[0].n_tile_h * p.lv[0].n_tile_w;
    for (int c = threadIdx.x; c < p.cap; c += blockDim.x) {
        int b = -1;
        if (cs[c] > 0) {
            int l = 0;
            while (l + 1 < p.L && c >= p.lv[l + 1].cell_offset) ++l;
            const int x0 = __float2int_rz(__fsub_rn(__fmul_rn((float)cx[c], p.lv[l].scale), (float)JSFE_B));
            const int y0 = __float2int_rz(__fsub_rn(__fmul_rn((float)cy[c], p.lv[l].scale), (float)JSFE_B));
            bx[c] = x0;
            by[c] = y0;
            b = (y0 / p.lv[0].tile_h) * p.lv[0].n_tile_w + x0 / p.lv[0].tile_w;
        }
        s_bucket[c] = b;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        int ex[64], ey[64], es[64], el[64], ec[64];
        int n = 0, l = 0;
        for (int c = 0; c < p.cap; ++c) {
            while (l + 1 < p.L && c >= p.lv[l + 1].cell_offset) ++l;
            if (s_bucket[c] == b && n < 64) { ex[n] = bx[c]; ey[n] = by[c]; es[n] = cs[c]; el[n] = l; ec[n] = c; ++n; }
        }
        for (int j = 0; j < n; ++j)
            for (int k = 0; k < n; ++k) {
                if (j == k || el[j] == el[k]) continue;
                if (es[j] && es[k]) {
                    const int dx = ex[j] - ex[k], dy = ey[j] - ey[k];
                    if (dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1) {
                        if (es[j] < es[k]) es[j] = 0; else es[k] = 0;
                    }
                }
            }
        for (int j = 0; j < n; ++j) cs[ec[j]] = es[j];
    }
}

// =================================================================================================
// K3  k_compact: ordered stream compaction of the per-cell candidates (level-major, cell row-major),
//     on the device.  replaces the D2H -> host loop -> H2D bounce of ORB_GPU::FAST_obtain_keypoints
//     (src/cuda/orb_FAST_obtain_keypoints.cpp:12-56).  One block per slot.
//     Also records, for the stereo matcher, the first keypoint index of every (level, tile row).
// =================================================================================================
__global__ void __launch_bounds__(1024) k_compact(const __grid_constant__ Params p, int slot0) {
    pdl_launch_dependents();
    pdl_wait();
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
// K4  k_orient_desc: one warp per keypoint.  Intensity-centroid angle from the level image (lanes =
//     columns of the radius-15 disc), then the 37x37 window of the BLURRED level around the keypoint is
//     staged in shared memory (40-byte rows, aligned words) and the 256 steered BRIEF tests sample it;
//     finally the output planes are written at the keypoint's final (compacted) index.
//     replaces FASTComputeOrientationGPU (src/cuda/orb_FAST_orientation.cu:17-65), ORB_compute_descriptorGPU
//     (src/cuda/orb_descriptor.cu:12-69), ORB_copy_output_GPU (src/cuda/orb_copy_output.cu:12-45) and the
//     per-level D2D descriptor copies (src/cuda/orb_gpu.cpp:819-831).
// =================================================================================================
#define JSFE_DP_R 18      // max |rotated rBRIEF offset|: 13*sqrt(2) = 18.4 -> rint <= 18
#define JSFE_DP_ROWS 37
#define JSFE_DP_PITCH 64  // TMA box {64, 37}: the box origin must be 16-byte aligned in x (measured: unaligned u8 x -> illegal
                          // instruction), so the box starts at floor16(x-18) and must span 15 + 37 columns
#define JSFE_DISC_ROWS 31
#define JSFE_DISC_PITCH 48  // box {48, 31} from floor16(x-15): 15 + 31 columns
#define JSFE_WIN_BYTES (JSFE_DP_ROWS * JSFE_DP_PITCH)       // 2368
#define JSFE_DISC_BYTES (JSFE_DISC_ROWS * JSFE_DISC_PITCH)  // 1488
#define JSFE_WIN_SLOT 2432                                   // 2368 rounded up to 128
#define JSFE_WARP_SMEM (JSFE_WIN_SLOT + 1536)                // + disc (1488 rounded up to 128)

#ifndef JSFE_KP_PER_WARP
#define JSFE_KP_PER_WARP 4   // keypoints a warp processes one after the other (amortises the pattern staging)
#endif

__global__ void __launch_bounds__(256) k_orient_desc(const __grid_constant__ Params p, const __grid_constant__ TmaMaps tm, int slot0) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ __align__(128) uint8_t s_buf[8][JSFE_WARP_SMEM];
    __shared__ __align__(8) uint64_t s_bar[8];
    __shared__ float2 s_pat[512];                    // rBRIEF sample offsets as floats (x, y)
    const int slot = slot0 + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = p.n_kp[slot];
    const int o0 = (blockIdx.x * 8 + warp) * JSFE_KP_PER_WARP;    // this warp's keypoints: o0 .. o0+3
    if ((int)blockIdx.x * 8 * JSFE_KP_PER_WARP >= n) return;
    uint8_t* win = s_buf[warp];                     // blurred window rows y-18..y+18, columns from wx0 (pitch 64)
    uint8_t* disc = s_buf[warp] + JSFE_WIN_SLOT;    // level image rows y-15..y+15, columns from dx0 (pitch 48)
    uint64_t* bar = &s_bar[warp];

    // keypoint record of iteration k (lane-uniform); loads for the next keypoint are issued as soon as the buffers are free
    auto issue = [&](int o, int& x, int& y, int& l, int& score) {
        const size_t so = (size_t)slot * p.cap + o;
        x = p.kp_x[so]; y = p.kp_y[so]; l = p.kp_l[so]; score = p.kp_s[so];
        if (p.use_tma) {
            if (lane == 0) {
                mbar_expect_tx(bar, JSFE_WIN_BYTES + JSFE_DISC_BYTES);
                tma_load_3d(win, &tm.win[l], bar, (x - JSFE_DP_R) & ~15, y - JSFE_DP_R, slot);
                tma_load_3d(disc, &tm.disc[l], bar, (x - 15) & ~15, y - 15, slot);
            }
        } else {
            const LevelGeom& lv = p.lv[l];
            const uint8_t* __restrict__ img = lv.img + (size_t)slot * lv.slot_stride;
            const uint8_t* __restrict__ blr = lv.blur + (size_t)slot * lv.slot_stride;
            const int wx0 = (x - JSFE_DP_R) & ~15, dx0 = (x - 15) & ~15;
            for (int i = lane; i < JSFE_DP_ROWS * (JSFE_DP_PITCH / 4); i += 32) {
                const int row = i / (JSFE_DP_PITCH / 4), wv = i - row * (JSFE_DP_PITCH / 4);
                const int gy = y - JSFE_DP_R + row, gx = wx0 + 4 * wv;
                uint32_t v = 0;
                if (gy >= 0 && gy < lv.h && gx < lv.pitch) v = __ldg(reinterpret_cast<const uint32_t*>(blr + (size_t)gy * lv.pitch + gx));
                *reinterpret_cast<uint32_t*>(win + row * JSFE_DP_PITCH + 4 * wv) = v;
            }
            for (int i = lane; i < JSFE_DISC_ROWS * (JSFE_DISC_PITCH / 4); i += 32) {
                const int row = i / (JSFE_DISC_PITCH / 4), wv = i - row * (JSFE_DISC_PITCH / 4);
                const int gy = y - 15 + row, gx = dx0 + 4 * wv;
                uint32_t v = 0;
                if (gy >= 0 && gy < lv.h && gx < lv.pitch) v = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)gy * lv.pitch + gx));
                *reinterpret_cast<uint32_t*>(disc + row * JSFE_DISC_PITCH + 4 * wv) = v;
            }
            __syncwarp();
        }
    };

    int x = 0, y = 0, l = 0, score = 0;
    if (p.use_tma && lane == 0) mbar_init(bar, 1);
    __syncwarp();
    if (o0 < n) issue(o0, x, y, l, score);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) s_pat[i] = __ldg(&p.tab->pat_f[i]);
    __syncthreads();
    // half-width table of the radius-15 disc, packed: nibble |u| = largest |v| whose row still contains column u
    const unsigned long long vmax_tab = p.vmax_packed;

#pragma unroll 1
    for (int k = 0; k < JSFE_KP_PER_WARP; ++k) {
        const int o = o0 + k;
        if (o >= n) break;
        if (p.use_tma) mbar_wait(bar, (uint32_t)(k & 1));
        const LevelGeom& lv = p.lv[l];
        const int wx0 = (x - JSFE_DP_R) & ~15, dx0 = (x - 15) & ~15;   // x >= 20, so both are >= 0

        // intensity centroid over the radius-15 disc (integer moments; any summation order is exact); lane = column
        int m10 = 0, m01 = 0;
        if (lane < 31) {
            const int u = lane - 15, au = abs(u);
            const uint8_t* col = disc + 15 * JSFE_DISC_PITCH + (x - 15 - dx0) + lane;
            const int vmax = (int)((vmax_tab >> (4 * au)) & 15ull);
#pragma unroll
            for (int v = -15; v <= 15; ++v) {
                const int I = (abs(v) <= vmax) ? (int)col[v * JSFE_DISC_PITCH] : 0;
                m10 += I;
                m01 += v * I;
            }
            m10 *= u;
        }
        m10 = __reduce_add_sync(0xffffffffu, m10);
        m01 = __reduce_add_sync(0xffffffffu, m01);
        const float angle = atan2f((float)m01, (float)m10);
        const float a = cosf(angle), b = sinf(angle);

        // descriptor byte `lane`: 8 comparisons of blurred samples.  rintf via the 1.5*2^23 magic add (exact, no F2I)
        const uint8_t* ctrb = win + JSFE_DP_R * JSFE_DP_PITCH + (x - wx0);
        unsigned val = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int t[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float2 pt = s_pat[(2 * i + q) * 32 + lane];
                const float fr = __fadd_rn(__fmaf_rn(b, pt.x, __fmul_rn(a, pt.y)), 12582912.0f);
                const float fc = __fadd_rn(__fmaf_rn(a, pt.x, -__fmul_rn(b, pt.y)), 12582912.0f);
                const int row = __float_as_int(fr) - 0x4B400000, col = __float_as_int(fc) - 0x4B400000;
                t[q] = ctrb[row * JSFE_DP_PITCH + col];
            }
            val |= (unsigned)(t[0] < t[1]) << i;
        }
        // outputs of this keypoint
        const size_t so = (size_t)slot * p.cap + o;
        const int ox = x, oy = y, ol = l, oscore = score;
        __syncwarp();                                   // every lane has finished reading win/disc
        if (o + 1 < n && k + 1 < JSFE_KP_PER_WARP) issue(o + 1, x, y, l, score);   // next keypoint's boxes are in flight during the stores
        p.desc[so * 32 + lane] = (uint8_t)val;
        if (lane == 0) {
            int* kp = p.kps + (size_t)slot * 6 * p.cap;
            const float sc = lv.scale;
            kp[0 * p.cap + o] = __float2int_rz(__fmul_rn((float)ox, sc));
            kp[1 * p.cap + o] = __float2int_rz(__fmul_rn((float)oy, sc));
            kp[2 * p.cap + o] = oscore;
            kp[3 * p.cap + o] = __float_as_int((float)((double)angle * (180.0 / 3.14159265358979323846)));
            kp[4 * p.cap + o] = ol;
            kp[5 * p.cap + o] = __float2int_rz(__fmul_rn(31.0f, sc));
            p.kp_angle[so] = angle;
        }
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
#ifndef JSFE_SM_BLOCKS
#define JSFE_SM_BLOCKS 8
#endif
__global__ void __launch_bounds__(256, JSFE_SM_BLOCKS) k_stereo_match(const __grid_constant__ Params p, const __grid_constant__ RightSide rsd,
                                                      int pair0, int th_high, int th_low, float mb, float mbf) {
    pdl_launch_dependents();
    pdl_wait();
    const int pair = pair0 + blockIdx.y;
    const int sl = rsd.left_mul * pair + rsd.left_add, sr = rsd.right_mul * pair + rsd.right_add;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + warp;
    const int nL = p.n_kp[sl];
    if (i >= nL) return;
    const int cap = p.cap;
    const int* kL = p.kps + (size_t)sl * 6 * cap;
    const int* kR = rsd.kps + (size_t)sr * 6 * cap;
    const size_t oL = (size_t)sl * cap + i;
    const int XL = kL[i], YL = kL[cap + i], lvl = kL[4 * cap + i];
    const float uL = (float)XL;
    const float maxD = __fdiv_rn(mbf, mb);
    const float minU = __fsub_rn(uL, maxD), maxU = uL;

    // left descriptor in registers (every lane holds all 8 words)
    const uint4* dl = reinterpret_cast<const uint4*>(p.desc + oL * 32);
    const uint4 l0 = __ldg(dl), l1 = __ldg(dl + 1);
    const uint8_t* descR = rsd.desc + (size_t)sr * cap * 32;
    const int* rs = rsd.row_start + (size_t)sr * (p.n_tile_rows + 1);

    // lanes 0..2 work out the candidate range of levels lvl-1, lvl, lvl+1 in parallel (tile rows that can reach row YL;
    // conservative, the exact row test is below); the three ranges are then scanned as one concatenated list
    int c0 = 0, cn = 0;
    float rho_l = 0.0f;
    {
        const int lr = lvl - 1 + lane;
        if (lane < 3 && lr >= 0 && lr < p.L) {
            const LevelGeom& g = p.lv[lr];
            rho_l = __fmul_rn(2.0f, g.scale);
            int ylo = (int)floorf(((float)YL - rho_l - 1.0f) / g.scale) - 1;
            int yhi = (int)ceilf(((float)YL + rho_l + 2.0f) / g.scale) + 1;
            ylo = max(ylo, 0);
            yhi = min(yhi, g.h - 1);
            if (ylo <= yhi) {
                c0 = rs[g.tile_row_offset + ylo / g.tile_h];
                cn = rs[g.tile_row_offset + yhi / g.tile_h + 1] - c0;
            }
        }
    }
    const int b0 = __shfl_sync(0xffffffffu, c0, 0), b1 = __shfl_sync(0xffffffffu, c0, 1), b2 = __shfl_sync(0xffffffffu, c0, 2);
    const int n0 = __shfl_sync(0xffffffffu, cn, 0), n1 = __shfl_sync(0xffffffffu, cn, 1), n2 = __shfl_sync(0xffffffffu, cn, 2);
    const float rho0 = __shfl_sync(0xffffffffu, rho_l, 0), rho1 = __shfl_sync(0xffffffffu, rho_l, 1), rho2 = __shfl_sync(0xffffffffu, rho_l, 2);
    const int n01 = n0 + n1, ntot = n01 + n2;

    unsigned best = ((unsigned)th_high << 16) | 0xFFFFu;  // strict-min scan from TH_HIGH; ties -> lowest right index
    for (int t = lane; t < ntot; t += 32) {
        const int r = (t < n0) ? b0 + t : (t < n01) ? b1 + (t - n0) : b2 + (t - n01);
        const float rho = (t < n0) ? rho0 : (t < n01) ? rho1 : rho2;
        const int yRi = kR[cap + r], uRi = kR[r];
        const float yR = (float)yRi, uR = (float)uRi;
        const int maxr = (int)ceilf(__fadd_rn(yR, rho)), minr = (int)floorf(__fsub_rn(yR, rho));
        if (YL < minr || YL > maxr) continue;
        if (!(uR >= minU && uR <= maxU)) continue;
        const uint4* dr = reinterpret_cast<const uint4*>(descR + (size_t)r * 32);
        const uint4 a = __ldg(dr), b = __ldg(dr + 1);
        const int d = __popc(l0.x ^ a.x) + __popc(l0.y ^ a.y) + __popc(l0.z ^ a.z) + __popc(l0.w ^ a.w) +
                      __popc(l1.x ^ b.x) + __popc(l1.y ^ b.y) + __popc(l1.z ^ b.z) + __popc(l1.w ^ b.w);
        best = min(best, ((unsigned)d << 16) | (unsigned)r);
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
            const uint8_t* imR = rsd.img[lvl] + (size_t)sr * g.slot_stride + (size_t)(int)svL * g.pitch + (int)suR0;
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
__global__ void __launch_bounds__(1024) k_stereo_outlier(const __grid_constant__ Params p, int pair0, int left_mul, int left_add) {
    pdl_launch_dependents();
    pdl_wait();
    __shared__ int hist[256];
    __shared__ int s_n, s_hi, s_k2, s_med;
    const int sl = left_mul * (pair0 + blockIdx.x) + left_add;
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
    if (n < 0) n = p.n_kp[slot];          // the count stays on the device: the caller learns it from the same stream later
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 6 * n && dst_kps) {
        const int plane = i / n, j = i - plane * n;
        dst_kps[i] = p.kps[(size_t)slot * 6 * p.cap + (size_t)plane * p.cap + j];
    }
    if (i < 8 * n && dst_desc)
        reinterpret_cast<uint32_t*>(dst_desc)[i] = reinterpret_cast<const uint32_t*>(p.desc + (size_t)slot * p.cap * 32)[i];
}

// =================================================================================================
// Adjacent rows: the three helper kernels the tracking thread calls.  Straight SoA kernels, stream
// ordered; float associations pinned with intrinsics to what nvcc emits for the reference (see oracle).
//   k_project_points  replaces ORB_Search_by_projection_project_on_GPU (src/cuda/orb_matcher.cu:17-64)
//   k_hamming_pairs   replaces ORB_compute_descriptor_Distance_GPU     (src/cuda/orb_matcher.cu:95-118)
//   k_in_frustum      replaces isInFrustum_GPU                         (src/cuda/tracking_isinfrustum.cu:19-107)
// =================================================================================================
__device__ __forceinline__ float dot3_plus(float a, float ra, float b, float rb, float c, float rc, float t) {
    return __fadd_rn(__fmaf_rn(c, rc, __fmaf_rn(b, rb, __fmul_rn(a, ra))), t);
}

__global__ void __launch_bounds__(256) k_project_points(int n, const float* __restrict__ px, const float* __restrict__ py,
                                                        const float* __restrict__ pz, const float* __restrict__ Rcw,
                                                        const float* __restrict__ tcw, float fx, float fy, float cx, float cy,
                                                        float min_x, float max_x, float min_y, float max_y, float* __restrict__ u,
                                                        float* __restrict__ v, float* __restrict__ invz, uint8_t* __restrict__ is_valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = px[i], y = py[i], z = pz[i];
    const float X = dot3_plus(x, __ldg(Rcw + 0), y, __ldg(Rcw + 1), z, __ldg(Rcw + 2), __ldg(tcw + 0));
    const float Y = dot3_plus(x, __ldg(Rcw + 3), y, __ldg(Rcw + 4), z, __ldg(Rcw + 5), __ldg(tcw + 1));
    const float Z = dot3_plus(x, __ldg(Rcw + 6), y, __ldg(Rcw + 7), z, __ldg(Rcw + 8), __ldg(tcw + 2));
    float iz = -1.0f, uu = -1.0f, vv = -1.0f;
    uint8_t ok = 0;
    if (Z > 0.0f) {
        iz = __frcp_rn(Z);
        uu = __fmaf_rn(__fmul_rn(X, fx), iz, cx);
        vv = __fmaf_rn(__fmul_rn(Y, fy), iz, cy);
        if (!(uu < min_x || uu > max_x || vv < min_y || vv > max_y)) ok = 1;
    }
    u[i] = uu; v[i] = vv; invz[i] = iz; is_valid[i] = ok;
}

// =================================================================================================
// the whole of ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) on the device
// (src/ORBmatcher.cpp:1647-1963): projection -> 64x48 grid window (Frame::GetFeaturesInArea, src/Frame.cpp:569-639)
// -> Hamming arg-min -> rotation histogram -> ComputeThreeMaxima cull.  The reference does this with 2 kernels,
// 3 host loops and 9 copies per call; here: k_frame_grid once per frame, then k_sbp_match + k_sbp_finish per call.
// The host's sequential semantics are kept by construction: arg-min ties go to the lowest CSR position (= the host's
// candidate order: cell column, cell row, insertion order), a keypoint claimed by several points keeps the highest
// point index (= the last assignment), and the cull runs in a second kernel after all assignments.
// =================================================================================================
#define JSFE_GRID_COLS JSFE_FRAME_GRID_COLS
#define JSFE_GRID_ROWS JSFE_FRAME_GRID_ROWS
#define JSFE_GRID_CELLS (JSFE_GRID_COLS * JSFE_GRID_ROWS)

typedef ::jsfe_sbp_args SbpArgs;   // the public argument block (include/jsfe.h) is passed to the kernels by value

__device__ __forceinline__ int grid_cell_of(float x, float y, float min_x, float min_y, float winv, float hinv) {
    const int cx = (int)roundf(__fmul_rn(__fsub_rn(x, min_x), winv)), cy = (int)roundf(__fmul_rn(__fsub_rn(y, min_y), hinv));
    return (cx < 0 || cx >= JSFE_GRID_COLS || cy < 0 || cy >= JSFE_GRID_ROWS) ? -1 : cx * JSFE_GRID_ROWS + cy;
}

// Frame::AssignFeaturesToGrid (src/Frame.cpp:464-479) as CSR: one block; counts by the whole block, exclusive scan,
// then warp 0 walks the keypoints in order (match_any gives the rank among equal cells) so every cell lists ascending indices
__global__ void __launch_bounds__(1024) k_frame_grid(int n, const float* __restrict__ x, const float* __restrict__ y, float min_x,
                                                     float min_y, float winv, float hinv, int* __restrict__ cell_start,
                                                     int* __restrict__ cell_items) {
    __shared__ int s_cnt[JSFE_GRID_CELLS];
    __shared__ int s_warp[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int c = tid; c < JSFE_GRID_CELLS; c += 1024) s_cnt[c] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int c = grid_cell_of(x[i], y[i], min_x, min_y, winv, hinv);
        if (c >= 0) atomicAdd(&s_cnt[c], 1);
    }
    __syncthreads();
    // exclusive scan of 3072 counters: 3 per thread, warp scan, then scan of the 32 warp totals
    const int c0 = s_cnt[3 * tid], c1 = s_cnt[3 * tid + 1], c2 = s_cnt[3 * tid + 2];
    int incl = c0 + c1 + c2;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += t;
        }
        s_warp[lane] = w;   // inclusive
    }
    __syncthreads();
    const int base = (warp ? s_warp[warp - 1] : 0) + incl - (c0 + c1 + c2);
    cell_start[3 * tid] = base;
    cell_start[3 * tid + 1] = base + c0;
    cell_start[3 * tid + 2] = base + c0 + c1;
    if (tid == 1023) cell_start[JSFE_GRID_CELLS] = base + c0 + c1 + c2;
    __syncthreads();
    s_cnt[3 * tid] = base;              // becomes the running fill position of every cell
    s_cnt[3 * tid + 1] = base + c0;
    s_cnt[3 * tid + 2] = base + c0 + c1;
    __syncthreads();
    if (warp == 0) {
        for (int b = 0; b < n; b += 32) {
            const int i = b + lane;
            const int c = i < n ? grid_cell_of(x[i], y[i], min_x, min_y, winv, hinv) : -1;
            const unsigned peers = __match_any_sync(0xffffffffu, c);
            const int leader = __ffs(peers) - 1;
            int pos = 0;
            if (c >= 0 && lane == leader) { pos = s_cnt[c]; s_cnt[c] = pos + __popc(peers); }
            pos = __shfl_sync(0xffffffffu, pos, leader);
            if (c >= 0) cell_items[pos + __popc(peers & ((1u << lane) - 1u))] = i;
            __syncwarp();
        }
    }
}

// one warp per last-frame point
__global__ void __launch_bounds__(256) k_sbp_match(const __grid_constant__ SbpArgs a) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= a.n_last) return;
    int out_idx = -1, out_dist = 256, out_bin = -1;
    // projection: the arithmetic of k_project_points (src/cuda/orb_matcher.cu:17-64)
    const float X = dot3_plus(a.px[i], __ldg(a.rcw9 + 0), a.py[i], __ldg(a.rcw9 + 1), a.pz[i], __ldg(a.rcw9 + 2), __ldg(a.tcw3 + 0));
    const float Y = dot3_plus(a.px[i], __ldg(a.rcw9 + 3), a.py[i], __ldg(a.rcw9 + 4), a.pz[i], __ldg(a.rcw9 + 5), __ldg(a.tcw3 + 1));
    const float Z = dot3_plus(a.px[i], __ldg(a.rcw9 + 6), a.py[i], __ldg(a.rcw9 + 7), a.pz[i], __ldg(a.rcw9 + 8), __ldg(a.tcw3 + 2));
    bool ok = false;
    float x = 0.f, y = 0.f, iz = 0.f;
    if (Z > 0.0f) {
        iz = __frcp_rn(Z);
        x = __fmaf_rn(__fmul_rn(X, a.fx), iz, a.cx);
        y = __fmaf_rn(__fmul_rn(Y, a.fy), iz, a.cy);
        ok = !(x < a.min_x || x > a.max_x || y < a.min_y || y > a.max_y);
    }
    if (ok) {
        const int lo = a.last_octave[i];
        const float r = __fmul_rn(a.th, a.scale_factors[lo]);
        const int min_level = a.level_mode == 1 ? lo : a.level_mode == 2 ? 0 : lo - 1;
        const int max_level = a.level_mode == 1 ? -1 : a.level_mode == 2 ? lo : lo + 1;
        const float winv = __fdiv_rn((float)JSFE_GRID_COLS, __fsub_rn(a.max_x, a.min_x));
        const float hinv = __fdiv_rn((float)JSFE_GRID_ROWS, __fsub_rn(a.max_y, a.min_y));
        const int cx0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(x, a.min_x), r), winv)));
        const int cx1 = min(JSFE_GRID_COLS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(x, a.min_x), r), winv)));
        const int cy0 = max(0, (int)floorf(__fmul_rn(__fsub_rn(__fsub_rn(y, a.min_y), r), hinv)));
        const int cy1 = min(JSFE_GRID_ROWS - 1, (int)ceilf(__fmul_rn(__fadd_rn(__fsub_rn(y, a.min_y), r), hinv)));
        if (cx0 < JSFE_GRID_COLS && cx1 >= 0 && cy0 < JSFE_GRID_ROWS && cy1 >= 0) {
            const bool check_levels = (min_level > 0) || (max_level >= 0);
            const uint4* dl = reinterpret_cast<const uint4*>(a.last_desc + (size_t)i * 32);
            const uint4 l0 = __ldg(dl), l1 = __ldg(dl + 1);
            const float ur = __fsub_rn(x, __fmul_rn(a.mbf, iz));
            unsigned best = (256u << 16) | 0xFFFFu;
            for (int ix = cx0; ix <= cx1; ++ix) {      // cells (ix, cy0..cy1) are contiguous in the CSR
                const int j0 = __ldg(a.cell_start + ix * JSFE_GRID_ROWS + cy0), j1 = __ldg(a.cell_start + ix * JSFE_GRID_ROWS + cy1 + 1);
                for (int j = j0 + lane; j < j1; j += 32) {
                    const int idx = __ldg(a.cell_items + j);
                    const int oc = a.cur_octave[idx];
                    if (check_levels && (oc < min_level || (max_level >= 0 && oc > max_level))) continue;
                    const float dx = __fsub_rn(a.cur_x[idx], x), dy = __fsub_rn(a.cur_y[idx], y);
                    if (!(fabsf(dx) < r && fabsf(dy) < r)) continue;
                    if (a.cur_occupied != nullptr && a.cur_occupied[idx]) continue;
                    const float cu = a.cur_uright[idx];
                    if (cu > 0.0f && fabsf(__fsub_rn(ur, cu)) > r) continue;
                    const uint4* dr = reinterpret_cast<const uint4*>(a.cur_desc + (size_t)idx * 32);
                    const uint4 p = __ldg(dr), q = __ldg(dr + 1);
                    const int d = __popc(l0.x ^ p.x) + __popc(l0.y ^ p.y) + __popc(l0.z ^ p.z) + __popc(l0.w ^ p.w) +
                                  __popc(l1.x ^ q.x) + __popc(l1.y ^ q.y) + __popc(l1.z ^ q.z) + __popc(l1.w ^ q.w);
                    best = min(best, ((unsigned)d << 16) | (unsigned)j);   // ties -> lowest CSR position = host order
                }
            }
            best = __reduce_min_sync(0xffffffffu, best);
            const int bd = (int)(best >> 16);
            if (bd <= a.th_high && (best & 0xFFFFu) != 0xFFFFu) {
                out_idx = __ldg(a.cell_items + (best & 0xFFFFu));
                out_dist = bd;
                if (a.check_orientation) {
                    float rot = __fsub_rn(a.last_angle[i], a.cur_angle[out_idx]);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    int bin = (int)roundf(__fmul_rn(rot, 1.0f / JSFE_HISTO_LENGTH));
                    if (bin == JSFE_HISTO_LENGTH) bin = 0;
                    out_bin = bin;
                }
            }
        }
    }
    if (lane == 0) {
        a.best_idx2[i] = out_idx;
        a.best_dist[i] = out_dist;
        a.rot_bin[i] = out_bin;
        if (out_idx >= 0) {
            atomicMax(a.cur_match + out_idx, i);      // the host's last assignment wins
            atomicAdd(a.n_matches, 1);
            if (out_bin >= 0) atomicAdd(a.hist + out_bin, 1);
        }
    }
}

// ComputeThreeMaxima (src/ORBmatcher.cpp:2097-2138) + the cull loop (:1940-1955); one block
__global__ void __launch_bounds__(1024) k_sbp_finish(const __grid_constant__ SbpArgs a) {
    __shared__ int s_ind[3];
    __shared__ int s_culled;
    if (threadIdx.x == 0) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int b = 0; b < JSFE_HISTO_LENGTH; ++b) {
            const int s = a.hist[b];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
            else if (s > max3) { max3 = s; ind3 = b; }
        }
        if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
        else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
        s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
        s_culled = 0;
    }
    __syncthreads();
    const int i1 = s_ind[0], i2 = s_ind[1], i3 = s_ind[2];
    int culled = 0;
    for (int i = threadIdx.x; i < a.n_last; i += blockDim.x) {
        const int b = a.rot_bin[i];
        if (b >= 0 && b != i1 && b != i2 && b != i3) {
            a.cur_match[a.best_idx2[i]] = -1;
            ++culled;
        }
    }
    if (culled) atomicAdd(&s_culled, culled);
    __syncthreads();
    if (threadIdx.x == 0) *a.n_matches -= s_culled;
}

// Frame::Frame's SoA -> cv::KeyPoint unpack (src/Frame.cpp:116-196) on the device
__global__ void __launch_bounds__(256) k_frame_view(const __grid_constant__ Params p, int slot, ::jsfe_cv_keypoint* __restrict__ keys,
                                                    float* __restrict__ x, float* __restrict__ y, int* __restrict__ octave,
                                                    float* __restrict__ angle) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_kp[slot]) return;
    const int* kp = p.kps + (size_t)slot * 6 * p.cap;
    const float fx = (float)kp[i], fy = (float)kp[p.cap + i], ang = __int_as_float(kp[3 * p.cap + i]);
    const int oct = kp[4 * p.cap + i];
    if (keys) {
        ::jsfe_cv_keypoint k;
        k.x = fx; k.y = fy; k.size = (float)kp[5 * p.cap + i]; k.angle = ang; k.response = (float)kp[2 * p.cap + i];
        k.octave = oct; k.class_id = -1;
        keys[i] = k;
    }
    if (x) x[i] = fx;
    if (y) y[i] = fy;
    if (octave) octave[i] = oct;
    if (angle) angle[i] = ang;
}

// =================================================================================================
// the input side on the device, so a raw camera frame crosses PCIe once.
//   k_remap_bilinear  cv::remap(src, dst, map_x, map_y, INTER_LINEAR), CV_32FC1 maps, BORDER_CONSTANT 0, 8-bit, as the
//                     reference's stereo example rectifies (Examples/Stereo/stereo_euroc.cpp:106-107,145-146); OpenCV's
//                     fixed-point scheme: coordinates rounded to 1/32 px, int16 weights scaled by 2^15 (imgwarp.cpp).
//                     Thread = 4 output pixels of one row; the decoded map entry (tap address, 4 weights) is reused for
//                     every image of the batch (frames of one camera share the maps), so the 8 B/px of map traffic is
//                     paid once per batch and each image costs ~1 B read (L2-gathered taps) + 1 B write per pixel.
//                     Two per-thread paths, chosen once when the map entries are decoded:
//                       window  (taken by a warp when, for every lane, the 4 pixels read one source row pair, at most 7 columns apart,
//                                all taps inside the image): per image 6 aligned 32-bit loads fetch a 12-byte window of both rows, two funnel
//                                shifts per row align it to the first tap, PRMT picks each pixel's 2 x 2 bytes and two IDP.2A
//                                (u16 x u8 dot products) apply the four weights -- 1.5 loads and ~8 ALU ops per pixel;
//                       gather  (anything else: borders, row crossings, wild maps): four byte loads per pixel.
//                     Both are the same integer arithmetic, i.e. bit-exact with OpenCV.
//   k_cvt_gray        cv::cvtColor(*2GRAY) for 8-bit BGR/RGB(A) (src/Tracking.cpp:260-285): 15-bit coefficients.
// =================================================================================================
#ifndef JSFE_REMAP_UNROLL
#define JSFE_REMAP_UNROLL 1   // gather path: images whose taps are loaded before any is used; measured on B200: 1 -> 0.58, 2 -> 0.6, 4 -> 1.1 us/image
#endif                        // (752x480): the register cost of deeper unrolling outweighs the extra loads in flight
__global__ void __launch_bounds__(256) k_remap_bilinear(const uint8_t* __restrict__ src, int src_h, int src_w, long long src_pitch,
                                                        long long src_stride, int n_images, const float* __restrict__ map_x,
                                                        const float* __restrict__ map_y, int dst_h, int dst_w,
                                                        uint8_t* __restrict__ dst, long long dst_pitch, long long dst_stride,
                                                        int word_stores, int word_loads) {
    const int x4 = (blockIdx.x * 64 + (threadIdx.x & 63)) << 2, y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x4 >= dst_w || y >= dst_h) return;
    long long off[4];      // offset of tap (iy, ix) in the source image (may point outside: guarded by the valid bits)
    unsigned w01[4], w23[4];   // packed int16 weights: (w00 | w01 << 16), (w10 | w11 << 16)
    unsigned valid = 0;    // 4 bits per pixel: tap k of pixel j inside the source
    int ixs[4], iys[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        off[j] = 0; w01[j] = 0; w23[j] = 0; ixs[j] = 0; iys[j] = 0;
        if (x4 + j < dst_w) {
            const size_t m = (size_t)y * dst_w + x4 + j;
            const int sx = __float2int_rn(__fmul_rn(__ldg(map_x + m), 32.0f)), sy = __float2int_rn(__fmul_rn(__ldg(map_y + m), 32.0f));
            const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767), fx = sx & 31, fy = sy & 31;
            int a = (32 - fx) * (32 - fy) * 32, b = fx * (32 - fy) * 32, c = (32 - fx) * fy * 32, d = fx * fy * 32;
            if ((fx | fy) == 0) { a = 32767; d = 1; }     // BilinearTab_i[0]: 32768 saturates to short, the 1 goes to tap (1,1)
            w01[j] = (unsigned)a | ((unsigned)b << 16);
            w23[j] = (unsigned)c | ((unsigned)d << 16);
            off[j] = (long long)iy * src_pitch + ix;
            ixs[j] = ix; iys[j] = iy;
            const unsigned x0 = (unsigned)ix < (unsigned)src_w, x1 = (unsigned)(ix + 1) < (unsigned)src_w;
            const unsigned y0 = (unsigned)iy < (unsigned)src_h, y1 = (unsigned)(iy + 1) < (unsigned)src_h;
            valid |= ((x0 & y0) | ((x1 & y0) << 1) | ((x0 & y1) << 2) | ((x1 & y1) << 3)) << (4 * j);
        }
    }
    // window path: all 4 pixels present, every tap inside the image, one source row pair, columns within 7 of the leftmost tap, and
    // the 12-byte window [base, base + 12) inside the row (base = leftmost tap rounded down to a word)
    const int ixmin = min(min(ixs[0], ixs[1]), min(ixs[2], ixs[3])), ixmax = max(max(ixs[0], ixs[1]), max(ixs[2], ixs[3]));
    const int wbase = ixmin & ~3;
    const bool fits = word_loads && word_stores && x4 + 3 < dst_w && valid == 0xFFFFu && iys[0] == iys[1] && iys[0] == iys[2] && iys[0] == iys[3] &&
                      ixmax - ixmin <= 6 && wbase + 12 <= src_w;
    // The choice is made per WARP: a warp whose lanes disagree would run both loops one after the other for every image (measured on
    // a radial-distortion map, where a row crossing hits ~4 of 32 lanes: 0.85 us/image against 0.59 for the gather path alone).
    const bool window = __all_sync(__activemask(), fits);
    if (window) {
        const unsigned sh = (unsigned)(ixmin & 3) * 8u;
        unsigned sel[4];      // PRMT selector of pixel j on the aligned 8-byte window: bytes (d, d+1), d = ix_j - ixmin
#pragma unroll
        for (int j = 0; j < 4; ++j) { const unsigned d = (unsigned)(ixs[j] - ixmin); sel[j] = d | ((d + 1u) << 4); }
        const long long woff = (long long)iys[0] * src_pitch + wbase;
        for (int img = blockIdx.z; img < n_images; img += gridDim.z) {
            const unsigned* __restrict__ t = reinterpret_cast<const unsigned*>(src + (size_t)img * src_stride + woff);
            const unsigned* __restrict__ b = reinterpret_cast<const unsigned*>(src + (size_t)img * src_stride + woff + src_pitch);
            const unsigned t0 = __ldg(t), t1 = __ldg(t + 1), t2 = __ldg(t + 2), b0 = __ldg(b), b1 = __ldg(b + 1), b2 = __ldg(b + 2);
            const unsigned T0 = __funnelshift_r(t0, t1, sh), T1 = __funnelshift_r(t1, t2, sh);   // 8 bytes from the leftmost tap
            const unsigned B0 = __funnelshift_r(b0, b1, sh), B1 = __funnelshift_r(b1, b2, sh);
            unsigned out = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned taps = __byte_perm(__byte_perm(T0, T1, sel[j]), __byte_perm(B0, B1, sel[j]), 0x5410);   // p00 p01 p10 p11
                const unsigned acc = __dp2a_hi(w23[j], taps, __dp2a_lo(w01[j], taps, 0u));
                out |= min((acc + (1u << 14)) >> 15, 255u) << (8 * j);
            }
            *reinterpret_cast<unsigned*>(dst + (size_t)img * dst_stride + (size_t)y * dst_pitch + x4) = out;
        }
        return;
    }
    // gather path; images in groups of JSFE_REMAP_UNROLL: all 16 x U tap loads are issued before the first is used
    for (int img0 = blockIdx.z * JSFE_REMAP_UNROLL; img0 < n_images; img0 += gridDim.z * JSFE_REMAP_UNROLL) {
        int v[JSFE_REMAP_UNROLL][4][4];
#pragma unroll
        for (int u = 0; u < JSFE_REMAP_UNROLL; ++u) {
            const uint8_t* __restrict__ s = src + (size_t)min(img0 + u, n_images - 1) * src_stride;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint8_t* p = s + off[j];
                const unsigned vb = valid >> (4 * j);
                v[u][j][0] = (vb & 1u) ? (int)__ldg(p) : 0;
                v[u][j][1] = (vb & 2u) ? (int)__ldg(p + 1) : 0;
                v[u][j][2] = (vb & 4u) ? (int)__ldg(p + src_pitch) : 0;
                v[u][j][3] = (vb & 8u) ? (int)__ldg(p + src_pitch + 1) : 0;
            }
        }
#pragma unroll
        for (int u = 0; u < JSFE_REMAP_UNROLL; ++u) {
            if (img0 + u >= n_images) break;
            unsigned out = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int acc = v[u][j][0] * (int)(w01[j] & 0xFFFFu) + v[u][j][1] * (int)(w01[j] >> 16) +
                                v[u][j][2] * (int)(w23[j] & 0xFFFFu) + v[u][j][3] * (int)(w23[j] >> 16);
                out |= (unsigned)min((acc + (1 << 14)) >> 15, 255) << (8 * j);
            }
            uint8_t* o = dst + (size_t)(img0 + u) * dst_stride + (size_t)y * dst_pitch + x4;
            if (word_stores && x4 + 3 < dst_w) *reinterpret_cast<unsigned*>(o) = out;
            else
                for (int j = 0; j < 4 && x4 + j < dst_w; ++j) o[j] = (uint8_t)(out >> (8 * j));
        }
    }
}

__global__ void __launch_bounds__(256) k_cvt_gray(const uint8_t* __restrict__ src, int h, int w, long long src_pitch, int channels,
                                                  int blue_idx, uint8_t* __restrict__ dst, long long dst_pitch) {
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* p = src + (size_t)y * src_pitch + (size_t)x * channels;
    const int b = p[blue_idx], g = p[1], r = p[blue_idx ^ 2];
    dst[(size_t)y * dst_pitch + x] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
}

__global__ void __launch_bounds__(256) k_hamming_pairs(int n, const int* __restrict__ idx_l, const int* __restrict__ idx_r,
                                                       const uint8_t* __restrict__ desc_l, const uint8_t* __restrict__ desc_r,
                                                       int* __restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* a = reinterpret_cast<const uint4*>(desc_l + (size_t)idx_l[i] * 32);
    const uint4* b = reinterpret_cast<const uint4*>(desc_r + (size_t)idx_r[i] * 32);
    const uint4 a0 = __ldg(a), a1 = __ldg(a + 1), b0 = __ldg(b), b1 = __ldg(b + 1);
    dist[i] = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
              __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

__global__ void __launch_bounds__(256) k_in_frustum(int n, const float* __restrict__ px, const float* __restrict__ py,
                                                    const float* __restrict__ pz, const float* __restrict__ pnx,
                                                    const float* __restrict__ pny, const float* __restrict__ pnz,
                                                    const float* __restrict__ max_distance, const float* __restrict__ inv_max,
                                                    const float* __restrict__ inv_min, const float* __restrict__ Rcw,
                                                    const float* __restrict__ tcw, const float* __restrict__ Ow, float fx, float fy,
                                                    float cx, float cy, int min_x, int max_x, int min_y, int max_y, int n_levels,
                                                    float log_sf, float view_cos_angle, float* __restrict__ invz, float* __restrict__ u,
                                                    float* __restrict__ v, int* __restrict__ level, float* __restrict__ view_cos,
                                                    uint8_t* __restrict__ in, const int* __restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t ok = 0;
    // ids != nullptr: query i is map point ids[i] of a RESIDENT pool (jsfe_mappool_*); the outputs stay indexed by the query
    const int s = ids ? ids[i] : i;
    const float x = px[s], y = py[s], z = pz[s];
    const float X = dot3_plus(x, __ldg(Rcw + 0), y, __ldg(Rcw + 1), z, __ldg(Rcw + 2), __ldg(tcw + 0));
    const float Y = dot3_plus(x, __ldg(Rcw + 3), y, __ldg(Rcw + 4), z, __ldg(Rcw + 5), __ldg(tcw + 1));
    const float Z = dot3_plus(x, __ldg(Rcw + 6), y, __ldg(Rcw + 7), z, __ldg(Rcw + 8), __ldg(tcw + 2));
    if (Z > 0.0f) {
        const float iz = __frcp_rn(Z);
        const float uu = __fmaf_rn(__fmul_rn(X, fx), iz, cx), vv = __fmaf_rn(__fmul_rn(Y, fy), iz, cy);
        if (!(uu < (float)min_x || uu > (float)max_x || vv < (float)min_y || vv > (float)max_y)) {
            const float ox = __fsub_rn(x, __ldg(Ow + 0)), oy = __fsub_rn(y, __ldg(Ow + 1)), oz = __fsub_rn(z, __ldg(Ow + 2));
            // association measured on the reference's sm_100a build: the SECOND product is the plain FMUL (tests/test_helpers.py)
            const float dist = __fsqrt_rn(__fmaf_rn(oz, oz, __fmaf_rn(ox, ox, __fmul_rn(oy, oy))));
            if (!(dist < inv_min[s] || dist > inv_max[s])) {
                const float vc = __fdiv_rn(__fmaf_rn(oz, pnz[s], __fmaf_rn(ox, pnx[s], __fmul_rn(oy, pny[s]))), dist);
                if (!(vc < view_cos_angle)) {
                    const float ratio = __fdiv_rn(max_distance[s], dist);
                    int ns = (int)ceilf(__fdiv_rn(logf(ratio), log_sf));
                    if (ns < 0) ns = 0; else if (ns >= n_levels) ns = n_levels - 1;
                    u[i] = uu; v[i] = vv; invz[i] = iz; level[i] = ns; view_cos[i] = vc;
                    ok = 1;
                }
            }
        }
    }
    in[i] = ok;
}

// jsfe_mappool_update: scatter n refreshed map points (staged contiguously) to their slots of the resident SoA
__global__ void __launch_bounds__(256) k_mappool_scatter(int n, const int* __restrict__ ids, const float* __restrict__ staged, int stage_stride,
                                                         float* __restrict__ pool, int pool_stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = ids[i];
#pragma unroll
    for (int a = 0; a < 9; ++a) pool[(size_t)a * pool_stride + s] = staged[(size_t)a * stage_stride + i];
}

}  // namespace jsfe
