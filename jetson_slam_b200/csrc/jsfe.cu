// jsfe.cu -- host side of libjsfe.so: handle, geometry/tables, device arenas, kernel sequencing and
// the C ABI declared in include/jsfe.h.  No CPU fallback: every compute entry point launches CUDA.
//
// Geometry and tables restate src/cuda/orb_gpu.cpp:22-441 of the reference (float32 level scales,
// truncating level sizes, per-level tile sizes, FAST arc LUT built by the reference's scan procedure,
// umax, 7x7 sigma=10 weights, OpenCV rBRIEF pattern).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/jsfe.h"
#include "jsfe_kernels.cuh"
#include "jsfe_gather.cuh"

#include <dlfcn.h>

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) return fail(JSFE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

const int kPattern[256 * 4] = {
#include "orb_pattern_31.inc"
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// The reference's smem tree over a tile's columns (orb_FAST_apply_NMS_G.cu:1331-1364), run on slot ids:
// column j beats column k on equal scores iff the tree returns j when only j and k hold the maximum.
void column_rank(int tile_w, uint8_t* rank, uint8_t* by_rank) {
    std::vector<int> val(tile_w), id(tile_w);
    const int steps = (int)std::ceil(std::log2((float)tile_w));
    for (int j = 0; j < tile_w; ++j) {
        int wins = 0;
        for (int k = 0; k < tile_w; ++k) {
            if (k == j) continue;
            for (int q = 0; q < tile_w; ++q) { val[q] = (q == j || q == k); id[q] = q; }
            int g = (tile_w - 1) / 2 + 1;
            for (int it = 0; it < steps; ++it) {
                for (int q = 0; q < g && q < tile_w; ++q)
                    if (q + g < tile_w && val[q] < val[q + g]) { val[q] = val[q + g]; id[q] = id[q + g]; }
                g = (g - 1) / 2 + 1;
            }
            if (id[0] == j) ++wins;
        }
        rank[j] = (uint8_t)(tile_w - 1 - wins);
    }
    for (int j = 0; j < tile_w; ++j) by_rank[rank[j]] = (uint8_t)j;
}

}  // namespace

struct jsfe_handle {
    jsfe::TmaMaps tma;    // TMA descriptors (second __grid_constant__ kernel parameter)
    jsfe_config cfg;
    int device = 0;
    int max_images = 0;
    jsfe::Params P;       // host copy of the kernel parameter block
    size_t fast_smem = 0; // dynamic shared memory of k_fast_cells
    void (*fast_kernel)(jsfe::Params, jsfe::TmaMaps, int) = nullptr;   // the k_fast_cells<compass mode, mask> instance of this handle
    std::vector<void*> dev_allocs;
    // pinned staging for jsfe_download_results / jsfe_get_*
    uint8_t* d_arena = nullptr;    // n_kp | kps | desc | u_right | depth of every slot (one allocation; Params points into it)
    uint8_t* h_arena = nullptr;    // the pinned mirror, same offsets
    size_t arena_off[5] = {0, 0, 0, 0, 0}, arena_bytes = 0;
    int32_t* h_n = nullptr;        // views into h_arena
    int32_t* h_kps = nullptr;
    uint8_t* h_desc = nullptr;
    float* h_ur = nullptr;
    float* h_dp = nullptr;
    int32_t* h_misc = nullptr;  // cap ints scratch x 4
    int64_t launches = 0;
    // optional per-kernel timing (CUDA events on the launching stream)
    bool profiling = false;
    struct Span { int stage; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    // end-to-end pipeline (jsfe_process_host_pairs): unpitched H2D staging + three streams
    int chunk_images = 0;
    // k_blur (FP32-pipe bound) runs beside k_fast_cells (integer-ALU bound) on an auxiliary stream
    cudaStream_t st_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int overlap_blur = 1;
    // jsfe_gather_begin packs this handle's results on its own stream; the next call that overwrites results waits for this event
    // first (deferred, so that the caller's stream is not stalled while OTHER handles still have work to run on it)
    cudaEvent_t pack_pending = nullptr;
    bool pdl = true;      // programmatic dependent launch between the kernels of the chain (JSFE_NO_PDL=1: plain launches)
    uint8_t* d_stage = nullptr;
    cudaStream_t st_h2d = nullptr, st_comp = nullptr, st_d2h = nullptr;
    std::vector<cudaEvent_t> ev_up, ev_done;
    // single-chunk calls of jsfe_process_host_pairs replay one CUDA graph (re-pitch + kernels + D2H); key = what is baked in
    cudaGraphExec_t pair_graph = nullptr;
    int pg_pairs = 0, pg_th_high = 0, pg_th_low = 0;
    float pg_mb = 0.f, pg_mbf = 0.f;
    bool pg_disabled = false;
    bool repitch_kernel = true;   // JSFE_NO_REPITCH_KERNEL=1: device-to-device 2-D copies instead (copy engine)
    int pending_pairs = 0;   // batch enqueued by jsfe_process_host_pairs_begin and not yet collected
    bool pipe_ready = false; // d_stage and the three pipeline streams exist
};

namespace {

template <typename T>
int dev_alloc(jsfe_handle* h, T** out, size_t count, bool zero = true) {
    void* ptr = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(T), 16);
    CU(cudaMalloc(&ptr, bytes));
    h->dev_allocs.push_back(ptr);
    if (zero) CU(cudaMemset(ptr, 0, bytes));
    *out = (T*)ptr;
    return JSFE_OK;
}

int check_slots(const jsfe_handle* h, int first, int n) {
    if (!h) return fail(JSFE_ERR_INVALID, "null handle");
    if (first < 0 || n < 0 || first + n > h->max_images)
        return fail(JSFE_ERR_CAPACITY, "slots [%d,%d) exceed the handle's %d image slots", first, first + n, h->max_images);
    return JSFE_OK;
}

cudaEvent_t take_event(jsfe_handle* h) {
    if (!h->event_pool.empty()) { cudaEvent_t e = h->event_pool.back(); h->event_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// RAII span around one kernel launch when profiling is on
struct StageTimer {
    jsfe_handle* h; cudaStream_t st; int stage; cudaEvent_t a = nullptr;
    StageTimer(jsfe_handle* h_, cudaStream_t st_, int stage_) : h(h_), st(st_), stage(stage_) {
        if (h->profiling) { a = take_event(h); cudaEventRecord(a, st); }
    }
    ~StageTimer() {
        if (a) { cudaEvent_t b = take_event(h); cudaEventRecord(b, st); h->spans.push_back({stage, a, b}); }
    }
};

// Launch with the programmatic-stream-serialization attribute: the kernel may start while its predecessor in the stream is still
// running; every kernel launched this way executes griddepcontrol.wait before it touches global memory (jsfe_kernels.cuh).
template <typename... KArgs, typename... Args>
cudaError_t launch_k(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

int post_launch(jsfe_handle* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(JSFE_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    ++h->launches;
    return JSFE_OK;
}

}  // namespace

extern "C" {

const char* jsfe_last_error(void) { return g_err.c_str(); }

int jsfe_create(const jsfe_config* cfg, jsfe_handle** out) {
    if (!cfg || !out) return fail(JSFE_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->n_levels < 1 || cfg->n_levels > JSFE_MAXL) return fail(JSFE_ERR_INVALID, "n_levels must be in [1,%d]", JSFE_MAXL);
    if (cfg->height < 1 || cfg->width < 1 || cfg->max_images < 1) return fail(JSFE_ERR_INVALID, "bad image size / max_images");
    if (!(cfg->scale_factor >= 1.0f)) return fail(JSFE_ERR_INVALID, "scale_factor must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(JSFE_ERR_CUDA, "no CUDA device: libjsfe has no CPU fallback");
    if (cfg->device_id < 0 || cfg->device_id >= ndev) return fail(JSFE_ERR_INVALID, "device_id %d out of range", cfg->device_id);
    CU(cudaSetDevice(cfg->device_id));

    jsfe_handle* h = new (std::nothrow) jsfe_handle;
    if (!h) return fail(JSFE_ERR_INVALID, "out of host memory");
    h->cfg = *cfg;
    h->cfg.mask = nullptr;
    h->device = cfg->device_id;
    h->max_images = cfg->max_images;
    jsfe::Params& P = h->P;
    memset(&P, 0, sizeof P);
    memset(&h->tma, 0, sizeof h->tma);
    P.L = cfg->n_levels;
    P.threshold = cfg->th_fast_max;  // the reference overwrites th_FAST_MIN_ and uses th_FAST_MAX only (orb_gpu.cpp:42-47)
    P.H0 = cfg->height;
    P.W0 = cfg->width;

    // ---- level geometry (orb_gpu.cpp:49-62, 224-258, 305-327), all float32
    float scale[JSFE_MAXL], inv[JSFE_MAXL];
    scale[0] = 1.0f;
    inv[0] = 1.0f;
    for (int i = 1; i < P.L; ++i) {
        scale[i] = cfg->scale_factor * scale[i - 1];
        inv[i] = 1.0f / scale[i];
    }
    int cells = 0, tile_rows = 0, items = 0;
    size_t smem_max = 0;
    int debug_cap = 0;
    if (const char* e = getenv("JSFE_DEBUG_FAST_CAP")) debug_cap = std::max(0, atoi(e));   // tests: force the work-list overflow path
    P.pyr_block_start[0] = 0;
    P.pyr_block_start[1] = 0;
    for (int i = 0; i < P.L; ++i) {
        jsfe::LevelGeom& g = P.lv[i];
        g.h = i ? (int)((float)cfg->height * inv[i]) : cfg->height;
        g.w = i ? (int)((float)cfg->width * inv[i]) : cfg->width;
        if (g.h < 1 || g.w < 1) { delete h; return fail(JSFE_ERR_INVALID, "level %d is empty", i); }
        g.pitch = (int)align_up(g.w, 16);
        g.scale = scale[i];
        g.inv_scale = inv[i];
        g.rscale = 1.0f / inv[i];
        if (cfg->fixed_multi_scale_tile_size) {
            g.tile_h = cfg->tile_h;
            g.tile_w = cfg->tile_w;
        } else {
            g.tile_h = (int)((float)cfg->tile_h * inv[i]);
            g.tile_w = (int)((float)cfg->tile_w * inv[i]);
        }
        // the reference divides by 128/tile_w (orb_FAST_apply_NMS_G.cu:1434); tile_h <= 254 is this library's key / work-list code width
        if (g.tile_w < 1 || g.tile_w > 128 || g.tile_h < 1 || g.tile_h > 254) {
            delete h;
            return fail(JSFE_ERR_INVALID, "level %d tile %dx%d unsupported (need 1<=tile_w<=128, 1<=tile_h<=254)", i, g.tile_h, g.tile_w);
        }
        g.n_tile_h = (g.h - 1) / g.tile_h + 1;
        g.n_tile_w = (g.w - 1) / g.tile_w + 1;
        g.cell_offset = cells;
        cells += g.n_tile_h * g.n_tile_w;
        g.tile_row_offset = tile_rows;
        tile_rows += g.n_tile_h;
        // y-lane count of the reference's launch (orb_FAST_apply_NMS_G.cu:1405-1431)
        int n_loc = std::max(1, std::min(10, g.tile_w / 3));
        n_loc = std::min(n_loc, g.tile_h);
        int T = (g.tile_h - 1) / n_loc + 1;
        if (T * 128 > 1024) T = 1024 / 128;
        g.T = T;
#ifndef JSFE_FAST_BLOCKS
#define JSFE_FAST_BLOCKS 6    // k_fast_cells blocks per SM the shared-memory budget aims at (6 x 256 threads x 40 registers fit)
#endif
#ifndef JSFE_FAST_GW
#define JSFE_FAST_GW 192      // widest cell group of a k_fast_cells block (A/B knob; s_best and the colkey table hold 192 columns)
#endif
        int fast_width = std::min(std::min(192, JSFE_FAST_GW), JSFE_FAST_PW - 23);   // pixels of cell group per k_fast_cells block (s_best and the colkey table hold 192 columns)
        if (const char* e = getenv("JSFE_FAST_WIDTH")) fast_width = std::max(16, std::min(fast_width, atoi(e)));
        g.cells_per_block = std::max(1, std::min(g.n_tile_w, fast_width / g.tile_w));
        g.blocks_per_row = (g.n_tile_w + g.cells_per_block - 1) / g.cells_per_block;
        g.block_offset = items;
        items += g.blocks_per_row * g.n_tile_h;
        const size_t gw = (size_t)g.cells_per_block * g.tile_w;
        const size_t pw = JSFE_FAST_PW, pr = g.tile_h + 8;   // fixed pitch: covers X0+GW+4-gx0 (<= 19 + 192 + 4) with gx0 = floor16(X0-4)
        if (gw + 8 + 15 > pw) { delete h; return fail(JSFE_ERR_INVALID, "internal: cell group wider than the staged tile"); }
        g.tile_pw = (int)pw;
        // phase A thread grid: 8-pixel column groups covering score columns [cs0, cs0+gw+1], cs0 = X0-1-floor16(X0-4) in [3,18]
        int ngx = 1;
        for (int cs0 = 3; cs0 <= 18; ++cs0) ngx = std::max(ngx, (int)((cs0 + gw + 1) >> 3) - (cs0 >> 3) + 1);
        g.fast_ngx = ngx;
        g.fast_nrl = std::max(1, 256 / ngx);
        g.fast_ngx_inv = 65536u / (unsigned)ngx + 1u;
        for (unsigned t = 0; t < 256; ++t)
            if (((t * g.fast_ngx_inv) >> 16) != t / (unsigned)ngx) { delete h; return fail(JSFE_ERR_INVALID, "internal: phase A reciprocal is not exact"); }
        // shared memory of a block: pixels + scores + work list.  The list gets what is left of the per-block budget that lets
        // JSFE_FAST_BLOCKS blocks share an SM (228 KB, 1 KB reserved per block, ~1 KB static), at most one slot per score position;
        // a tile with more survivors than slots is evaluated densely (exact, slower).  If the budget leaves less than a third of
        // the positions, the list gets a third and fewer blocks fit.
        const size_t sw = (size_t)jsfe::fast_score_pitch((int)gw);
        const size_t positions = (size_t)(g.tile_h + 2) * sw, fixed = pr * pw + positions * 2 + 64;
        const size_t budget = (228 * 1024) / JSFE_FAST_BLOCKS - 2048;
        size_t cap = budget > fixed ? (budget - fixed) / 2 : 0;
        cap = std::min(positions, std::max(cap, positions / 3));
        if (debug_cap > 0) cap = std::min(cap, (size_t)debug_cap);
        g.fast_cap = (int)cap;
        smem_max = std::max(smem_max, fixed + cap * 2);
        g.slot_stride = align_up((size_t)g.h * g.pitch, 256);
        if (i >= 1) P.pyr_block_start[i + 1] = P.pyr_block_start[i] + ((g.pitch + 127) / 128) * ((g.h + 31) / 32);
        if (g.w >= 16384 || g.h >= 16384) { delete h; return fail(JSFE_ERR_INVALID, "images larger than 16383 pixels are not supported"); }
    }
    P.pyr_blocks_total = P.L > 1 ? P.pyr_block_start[P.L] : 0;
    P.blur_item_start[0] = 0;
    for (int i = 0; i < P.L; ++i) {
        const jsfe::LevelGeom& g = P.lv[i];
        int items = 0;
        if (g.w > 2 * JSFE_B && g.h > 2 * JSFE_B)
            items = ((g.w - 2 * JSFE_B + 3) / 4) * ((g.h - 2 * JSFE_B + JSFE_BLUR_ROWS - 1) / JSFE_BLUR_ROWS);
        P.blur_item_start[i + 1] = P.blur_item_start[i] + items;
    }
    P.blur_items_total = P.blur_item_start[P.L];
    P.fast_items_total = items;
    P.cap = cells;
    P.n_tile_rows = tile_rows;
    h->fast_smem = smem_max;
    if (cells >= 65535) { delete h; return fail(JSFE_ERR_INVALID, "more than 65534 NMS cells per image"); }

    // ---- tables
    jsfe::DevTables* T = new jsfe::DevTables;
    memset(T, 0, sizeof *T);
    for (int m = 0; m < 0xFFFF; ++m) {  // scan procedure of orb_gpu.cpp:366-436; entry 0xFFFF stays 0
        int run = 0, accepted = 0;
        for (int bit = 15; bit >= 0; --bit) {
            if (m & (1 << bit)) ++run;
            else {
                if (run >= cfg->fast_n_min && run <= cfg->fast_n_max) { accepted = 1; break; }
                run = 0;
            }
        }
        if (!accepted)
            for (int bit = 15; bit >= 0; --bit) { if (m & (1 << bit)) ++run; else break; }
        if (run >= cfg->fast_n_min && run <= cfg->fast_n_max) T->lut_bits[m >> 5] |= 1u << (m & 31);
    }
    {  // umax (orb_gpu.cpp:161-182)
        const int R = 15;
        const double half = (double)((float)R * std::sqrt(2.f) / 2);
        const int vmax = (int)std::floor(half + 1), vmin = (int)std::ceil(half);
        for (int v = 0; v <= vmax; ++v) T->umax[v] = (int)std::lrint(std::sqrt((double)(R * R) - (double)(v * v)));
        for (int v = R, v0 = 0; v >= vmin; --v) {
            while (T->umax[v0] == T->umax[v0 + 1]) ++v0;
            T->umax[v] = v0;
            ++v0;
        }
    }
    P.vmax_packed = 0;
    for (int au = 0; au <= 15; ++au) {
        int vmax = 0;
        for (int v = 0; v <= 15; ++v) if (au <= T->umax[v]) vmax = v;
        P.vmax_packed |= (unsigned long long)vmax << (4 * au);
    }
    {  // 7x7 sigma=10 weights (orb_gpu.cpp:196-218): double exp of a float argument, float sum, float divide
        const float sigma2 = 10.0f * 10.0f;
        float sum = 0;
        int n = 0;
        for (int j = -3; j <= 3; ++j)
            for (int k = -3; k <= 3; ++k) {
                const float arg = (float)(-(j * j + k * k)) / (2 * sigma2);
                T->gauss[n] = (float)std::exp((double)arg);
                sum += T->gauss[n];
                ++n;
            }
        for (int i = 0; i < 49; ++i) T->gauss[i] /= sum;
    }
    {  // separable factors of the 7x7 weights: w[j][k] = g[j]*g[k] in exact arithmetic, g = exp(-j^2/200)/sum
        double g[7], gs = 0;
        for (int j = 0; j < 7; ++j) { g[j] = std::exp(-(double)((j - 3) * (j - 3)) / 200.0); gs += g[j]; }
        double worst = 0;
        for (int j = 0; j < 7; ++j) { T->sep_a[j] = (float)(g[j] / gs); T->sep_b[j] = (float)(g[j] / gs); }
        for (int j = 0; j < 7; ++j)
            for (int k = 0; k < 7; ++k)
                worst = std::max(worst, std::fabs((double)T->gauss[j * 7 + k] - (double)T->sep_a[j] * (double)T->sep_b[k]));
        if (worst > 1e-8) { delete T; delete h; return fail(JSFE_ERR_INVALID, "separable blur factors off by %g", worst); }
    }
    {  // compass pre-test of k_fast_cells: every accepted ring mask must contain `mode` adjacent compass bits
        auto compass_ok = [](int m, int mode) {
            const int c[4] = {(m >> 0) & 1, (m >> 4) & 1, (m >> 8) & 1, (m >> 12) & 1};
            for (int s0 = 0; s0 < 4; ++s0) {
                int all = 1;
                for (int k = 0; k < mode; ++k) all &= c[(s0 + k) & 3];
                if (all) return true;
            }
            return false;
        };
        int mode = std::min(3, std::max(0, cfg->fast_n_min / 4));
        for (; mode > 0; --mode) {
            bool ok = true;
            for (int m = 0; m < 0xFFFF && ok; ++m)
                if ((T->lut_bits[m >> 5] >> (m & 31)) & 1u) ok = compass_ok(m, mode);
            if (ok) break;
        }
        P.compass_mode = mode;
    }
    P.fast_q_thresh = P.threshold >= 3 ? (P.threshold - 3) / 4 + 1 : 0;
    if (P.threshold < 0 || P.threshold > 255) { delete T; delete h; return fail(JSFE_ERR_INVALID, "th_fast_max must be in [0,255]"); }
    for (int m = 0; m < 0x10000; ++m) {   // LUT in the bit order k_fast_cells' flag merge produces (see fast_eval)
        if (!((T->lut_bits[m >> 5] >> (m & 31)) & 1u)) continue;
        unsigned idx = 0;
        for (int k = 0; k < 16; ++k)
            if (m & (1 << k)) idx |= 1u << (4 * (k & 3) + 3 - (k >> 2));   // ring k = byte k%4 of word k/4
        T->lut_perm[idx >> 5] |= 1u << (idx & 31);
    }
    for (int i = 0; i < 512; ++i) {  // orb_bitpattern.cpp:266-273
        T->pat_x[i] = (int8_t)kPattern[2 * i];
        T->pat_y[i] = (int8_t)kPattern[2 * i + 1];
        T->pat_f[(i & 15) * 32 + (i >> 4)] = make_float2((float)T->pat_x[i], (float)T->pat_y[i]);  // [sample j][byte b]: bank-conflict-free per lane
    }
    for (int i = 0; i < P.L; ++i) {
        const jsfe::LevelGeom& g = P.lv[i];
        column_rank(g.tile_w, T->col_rank[i], T->col_by_rank[i]);
        for (int c = 0; c < 192 && c < g.cells_per_block * g.tile_w; ++c)
            T->colkey[i][c] = (uint16_t)(((127u - T->col_rank[i][c % g.tile_w]) << 8) | (unsigned)(c / g.tile_w));
        for (int dy = 0; dy < g.tile_h && dy < 256; ++dy)
            T->rowkey[i][dy] = (uint16_t)(((7u - (unsigned)(dy % g.T)) << 8) | (255u - (unsigned)dy));
    }

    int rc = JSFE_OK;
    auto bail = [&](int code) { delete T; jsfe_destroy(h); return code; };
    jsfe::DevTables* dT = nullptr;
    if ((rc = dev_alloc(h, &dT, 1)) != JSFE_OK) return bail(rc);
    if (cudaMemcpy(dT, T, sizeof *T, cudaMemcpyHostToDevice) != cudaSuccess) return bail(fail(JSFE_ERR_CUDA, "table upload failed"));
    if (cudaMemcpyToSymbol(jsfe::c_sep_a, T->sep_a, sizeof T->sep_a) != cudaSuccess ||
        cudaMemcpyToSymbol(jsfe::c_sep_b, T->sep_b, sizeof T->sep_b) != cudaSuccess)
        return bail(fail(JSFE_ERR_CUDA, "constant upload failed"));
    P.tab = dT;
    delete T;
    T = nullptr;
    {  // k_fast_cells work item -> (level, tile row, block in row)
        std::vector<uint32_t> map((size_t)P.fast_items_total);
        for (int i = 0; i < P.L; ++i) {
            const jsfe::LevelGeom& g = P.lv[i];
            for (int ty = 0; ty < g.n_tile_h; ++ty)
                for (int bx = 0; bx < g.blocks_per_row; ++bx)
                    map[(size_t)g.block_offset + (size_t)ty * g.blocks_per_row + bx] = ((uint32_t)i << 28) | ((uint32_t)ty << 14) | (uint32_t)bx;
        }
        uint32_t* d_map = nullptr;
        if ((rc = dev_alloc(h, &d_map, map.size())) != JSFE_OK) return bail(rc);
        if (cudaMemcpy(d_map, map.data(), map.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(JSFE_ERR_CUDA, "work map upload failed"));
        P.fast_map = d_map;
        std::vector<uint32_t> pm((size_t)std::max(P.pyr_blocks_total, 1));
        for (int i = 1; i < P.L; ++i) {
            const jsfe::LevelGeom& g = P.lv[i];
            const int tiles_x = (g.pitch + 127) / 128, tiles_y = (g.h + 31) / 32;
            for (int ty = 0; ty < tiles_y; ++ty)
                for (int tx = 0; tx < tiles_x; ++tx)
                    pm[(size_t)P.pyr_block_start[i] + (size_t)ty * tiles_x + tx] = ((uint32_t)i << 28) | ((uint32_t)ty << 14) | (uint32_t)tx;
        }
        uint32_t* d_pm = nullptr;
        if ((rc = dev_alloc(h, &d_pm, pm.size())) != JSFE_OK) return bail(rc);
        if (cudaMemcpy(d_pm, pm.data(), pm.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess)
            return bail(fail(JSFE_ERR_CUDA, "work map upload failed"));
        P.pyr_map = d_pm;
    }

    // ---- images + masks
    const size_t M = (size_t)h->max_images, cap = (size_t)P.cap;
    for (int i = 0; i < P.L; ++i) {
        jsfe::LevelGeom& g = P.lv[i];
        if ((rc = dev_alloc(h, &g.img, g.slot_stride * M)) != JSFE_OK) return bail(rc);
        if ((rc = dev_alloc(h, &g.blur, g.slot_stride * M)) != JSFE_OK) return bail(rc);  // stays 0 outside the blurred interior
        g.mask = nullptr;
        if (cfg->mask) {  // INTER_NEAREST + THRESH_BINARY(10), orb_gpu.cpp:78-90
            std::vector<uint8_t> m((size_t)g.h * g.pitch, 0);
            const double ifx = 1.0 / ((double)g.w / cfg->width), ify = 1.0 / ((double)g.h / cfg->height);
            for (int y = 0; y < g.h; ++y) {
                const int sy = std::min((int)std::floor(y * ify), cfg->height - 1);
                for (int x = 0; x < g.w; ++x) {
                    const int sx = std::min((int)std::floor(x * ifx), cfg->width - 1);
                    m[(size_t)y * g.pitch + x] = cfg->mask[(size_t)sy * cfg->mask_pitch + sx] > 10 ? 255 : 0;
                }
            }
            uint8_t* dm = nullptr;
            if ((rc = dev_alloc(h, &dm, m.size())) != JSFE_OK) return bail(rc);
            if (cudaMemcpy(dm, m.data(), m.size(), cudaMemcpyHostToDevice) != cudaSuccess) return bail(fail(JSFE_ERR_CUDA, "mask upload failed"));
            g.mask = dm;
        }
    }
    // ---- TMA descriptors: 3-D u8 tensors {pitch, h, slots} over every level image / blurred level
    P.use_tma = (getenv("JSFE_NO_TMA") && atoi(getenv("JSFE_NO_TMA"))) ? 0 : 1;
    for (int i = 0; i < P.L && P.use_tma; ++i)
        if (P.lv[i].tile_h + 8 > 256 || P.lv[i].tile_pw > 256) P.use_tma = 0;   // box extents are limited to 256
    if (P.use_tma) {
        typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn)
            return bail(fail(JSFE_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver"));
        auto encode = [&](CUtensorMap* m, void* base, const jsfe::LevelGeom& g, unsigned bw, unsigned bh) -> bool {
            const cuuint64_t dims[3] = {(cuuint64_t)g.pitch, (cuuint64_t)g.h, (cuuint64_t)M};
            const cuuint64_t strides[2] = {(cuuint64_t)g.pitch, (cuuint64_t)g.slot_stride};
            const cuuint32_t box[3] = {bw, bh, 1}, estr[3] = {1, 1, 1};
            return ((encode_fn)fn)(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        };
        for (int i = 0; i < P.L; ++i) {
            const jsfe::LevelGeom& g = P.lv[i];
            if (!encode(&h->tma.tile[i], g.img, g, (unsigned)g.tile_pw, (unsigned)(g.tile_h + 8)) ||
                !encode(&h->tma.disc[i], g.img, g, 48, 31) || !encode(&h->tma.win[i], g.blur, g, 64, 37))
                return bail(fail(JSFE_ERR_CUDA, "cuTensorMapEncodeTiled failed for level %d", i));
        }
    }
    // ---- per-slot arrays
    if ((rc = dev_alloc(h, &P.cell_x, M * cap)) || (rc = dev_alloc(h, &P.cell_y, M * cap)) || (rc = dev_alloc(h, &P.cell_s, M * cap)) ||
        (rc = dev_alloc(h, &P.kp_x, M * cap)) || (rc = dev_alloc(h, &P.kp_y, M * cap)) || (rc = dev_alloc(h, &P.kp_s, M * cap)) ||
        (rc = dev_alloc(h, &P.kp_l, M * cap)) || (rc = dev_alloc(h, &P.kp_angle, M * cap)) ||
        (rc = dev_alloc(h, &P.n_per_level, M * JSFE_MAXL)) || (rc = dev_alloc(h, &P.row_start, M * (size_t)(tile_rows + 1))) ||
        (rc = dev_alloc(h, &P.best_idx, M * cap)) || (rc = dev_alloc(h, &P.best_dist, M * cap)) ||
        (rc = dev_alloc(h, &P.sad_best, M * cap)))
        return bail(rc);
    {   // the five result arrays a caller downloads live in ONE arena (device and pinned host, same offsets), so that the results
        // of the whole handle -- the single-pair latency path -- cross PCIe as one copy instead of five
        size_t off = 0;
        auto carve = [&](size_t bytes) { const size_t o = off; off = align_up(off + bytes, 256); return o; };
        h->arena_off[0] = carve(M * 4);
        h->arena_off[1] = carve(M * 6 * cap * 4);
        h->arena_off[2] = carve(M * cap * 32);
        h->arena_off[3] = carve(M * cap * 4);
        h->arena_off[4] = carve(M * cap * 4);
        h->arena_bytes = off;
        uint8_t* d = nullptr;
        if ((rc = dev_alloc(h, &d, off)) != JSFE_OK) return bail(rc);
        h->d_arena = d;
        P.n_kp = (int*)(d + h->arena_off[0]);
        P.kps = (int*)(d + h->arena_off[1]);
        P.desc = d + h->arena_off[2];
        P.u_right = (float*)(d + h->arena_off[3]);
        P.depth = (float*)(d + h->arena_off[4]);
    }
    P.blur_amb_units = JSFE_BLUR_AMB_UNITS;
    if (const char* e = getenv("JSFE_DEBUG_BLUR_UNITS")) P.blur_amb_units = (unsigned)atoi(e);   // experiments only: < 18 breaks exactness
    if (cfg->apply_nms_ms && P.L > 1 && !cfg->nms_ms_mode_gpu) {
        // k_nms_ms_buckets walks at most 64 candidates per level-0 bucket (the reference's host rule has no cap,
        // orb_FAST_apply_NMS_MS.cpp:15-122).  A bucket is tile_h0 x tile_w0 level-0 pixels; level l has one candidate per cell of
        // c_l = tile_l * scale_l level-0 pixels, and an interval of tile_0 pixels meets at most floor((tile_0 - 1) / c_l) + 2 such cells per axis.
        // Geometries whose bound exceeds the scratch are refused instead of silently diverging.
        long bound = 0;
        for (int i = 0; i < P.L; ++i) {
            const double ch = (double)P.lv[i].tile_h * P.lv[i].scale, cw = (double)P.lv[i].tile_w * P.lv[i].scale;
            bound += (long)(std::floor((P.lv[0].tile_h - 1) / ch) + 2) * (long)(std::floor((P.lv[0].tile_w - 1) / cw) + 2);
        }
        if (bound > 64) return bail(fail(JSFE_ERR_INVALID, "cross-scale NMS (CPU rule): up to %ld candidates per level-0 bucket, the kernel holds 64", bound));
    }
    if (cfg->apply_nms_ms && P.L > 1) {
        int ts = 64;
        while (ts < 2 * P.cap) ts <<= 1;
        P.ms_table_size = ts;
        if ((rc = dev_alloc(h, &P.ms_keys, M * (size_t)ts)) || (rc = dev_alloc(h, &P.ms_sums, M * (size_t)ts)) ||
            (rc = dev_alloc(h, &P.ms_cnts, M * (size_t)ts)) || (rc = dev_alloc(h, &P.ms_drop, M * cap)))
            return bail(rc);
        if ((size_t)P.cap * 4 > 48 * 1024 &&
            cudaFuncSetAttribute(jsfe::k_nms_ms_buckets, cudaFuncAttributeMaxDynamicSharedMemorySize, P.cap * 4) != cudaSuccess)
            return bail(fail(JSFE_ERR_CUDA, "k_nms_ms_buckets needs %d bytes of shared memory", P.cap * 4));
    }
    // ---- pinned staging
    if (cudaMallocHost((void**)&h->h_arena, h->arena_bytes + 16) != cudaSuccess ||
        cudaMallocHost((void**)&h->h_misc, 5 * cap * sizeof(int32_t) + 16) != cudaSuccess)
        return bail(fail(JSFE_ERR_CUDA, "pinned host allocation failed: %s", cudaGetErrorString(cudaGetLastError())));
    h->h_n = (int32_t*)(h->h_arena + h->arena_off[0]);
    h->h_kps = (int32_t*)(h->h_arena + h->arena_off[1]);
    h->h_desc = h->h_arena + h->arena_off[2];
    h->h_ur = (float*)(h->h_arena + h->arena_off[3]);
    h->h_dp = (float*)(h->h_arena + h->arena_off[4]);
    {
        using jsfe::k_fast_cells;
        void (*const table[4][2])(jsfe::Params, jsfe::TmaMaps, int) = {
            {k_fast_cells<0, false>, k_fast_cells<0, true>}, {k_fast_cells<1, false>, k_fast_cells<1, true>},
            {k_fast_cells<2, false>, k_fast_cells<2, true>}, {k_fast_cells<3, false>, k_fast_cells<3, true>}};
        h->fast_kernel = table[P.compass_mode][cfg->mask ? 1 : 0];
    }
    if (h->fast_smem + 4096 > 48 * 1024) {  // dynamic + static shared memory beyond 48 KB needs the opt-in
        if (cudaFuncSetAttribute(h->fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fast_smem) != cudaSuccess)
            return bail(fail(JSFE_ERR_CUDA, "k_fast_cells needs %zu bytes of shared memory", h->fast_smem));
    }
    {   // sub-batch size: images whose level-0 + levels + blurred levels fit in ~60% of L2
        size_t per_image = 0;
        for (int i = 0; i < P.L; ++i) per_image += P.lv[i].slot_stride * (i ? 2 : 1) + (i ? 0 : P.lv[i].slot_stride);
        int l2 = 0;
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, h->device);
        (void)per_image; (void)l2;
        // Measured on B200 (profiles/r01_chunk_sweep.txt): every kernel of the path is instruction-issue-bound, not
        // memory-bound, so L2-sized sub-batches only add launch tails (30.5k pairs/s unchunked vs 24.5k..29.5k chunked).
        // Default: no sub-batching; JSFE_CHUNK_IMAGES overrides for experiments.
        h->chunk_images = 0;
        if (const char* e = getenv("JSFE_CHUNK_IMAGES")) h->chunk_images = atoi(e);
    }
    if (const char* e = getenv("JSFE_NO_OVERLAP")) h->overlap_blur = atoi(e) ? 0 : 1;
    if (const char* e = getenv("JSFE_NO_PDL")) h->pdl = atoi(e) == 0;
    if (const char* e = getenv("JSFE_NO_REPITCH_KERNEL")) h->repitch_kernel = atoi(e) == 0;
    if (cudaStreamCreateWithFlags(&h->st_aux, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(JSFE_ERR_CUDA, "stream/event creation failed"));
    CU(cudaDeviceSynchronize());
    *out = h;
    return JSFE_OK;
}

int jsfe_destroy(jsfe_handle* h) {
    if (!h) return JSFE_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (void* p : h->dev_allocs) cudaFree(p);
    for (auto& sp : h->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_up) cudaEventDestroy(e);
    for (cudaEvent_t e : h->ev_done) cudaEventDestroy(e);
    if (h->st_aux) cudaStreamDestroy(h->st_aux);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->pair_graph) cudaGraphExecDestroy(h->pair_graph);
    if (h->st_h2d) cudaStreamDestroy(h->st_h2d);
    if (h->st_comp) cudaStreamDestroy(h->st_comp);
    if (h->st_d2h) cudaStreamDestroy(h->st_d2h);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->h_arena) cudaFreeHost(h->h_arena);
    if (h->h_misc) cudaFreeHost(h->h_misc);
    delete h;
    return JSFE_OK;
}

int jsfe_max_keypoints(const jsfe_handle* h) { return h ? h->P.cap : fail(JSFE_ERR_INVALID, "null handle"); }
int jsfe_num_levels(const jsfe_handle* h) { return h ? h->P.L : fail(JSFE_ERR_INVALID, "null handle"); }
int64_t jsfe_launch_count(const jsfe_handle* h) { return h ? h->launches : 0; }

int jsfe_get_level_info(const jsfe_handle* h, int level, jsfe_level_info* out) {
    if (!h || !out || level < 0 || level >= h->P.L) return fail(JSFE_ERR_INVALID, "bad level");
    const jsfe::LevelGeom& g = h->P.lv[level];
    out->height = g.h; out->width = g.w; out->pitch = g.pitch;
    out->tile_h = g.tile_h; out->tile_w = g.tile_w; out->n_tile_h = g.n_tile_h; out->n_tile_w = g.n_tile_w;
    out->cell_offset = g.cell_offset; out->scale = g.scale; out->inv_scale = g.inv_scale;
    return JSFE_OK;
}

int jsfe_set_images(jsfe_handle* h, int first_slot, int n, const uint8_t* src, int64_t row_pitch, int64_t image_stride,
                    int src_is_device, void* stream) {
    int rc = check_slots(h, first_slot, n);
    if (rc) return rc;
    if (!src || row_pitch < h->P.W0) return fail(JSFE_ERR_INVALID, "bad source image pointer / pitch");
    CU(cudaSetDevice(h->device));
    const jsfe::LevelGeom& g = h->P.lv[0];
    if (src_is_device && row_pitch == g.w && h->repitch_kernel) {
        // contiguous rows on the device (the staging buffer of jsfe_process_host_pairs, or any packed device frame)
        const int chunks = g.h * (g.pitch / 16);
        CU(launch_k(h->pdl, jsfe::k_repitch, dim3((chunks + 255) / 256, n), dim3(256), 0, (cudaStream_t)stream, src, (long long)row_pitch,
                    (long long)image_stride, src, src + (size_t)(n - 1) * image_stride + (size_t)g.h * row_pitch,
                    g.img + (size_t)first_slot * g.slot_stride, g.pitch, (unsigned long long)g.slot_stride, g.h, g.w));
        return JSFE_OK;
    }
    const cudaMemcpyKind kind = src_is_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (image_stride == row_pitch * (int64_t)g.h && g.slot_stride == (size_t)g.pitch * g.h) {
        // rows of consecutive images are equally spaced on both sides: one 2-D copy for the whole batch
        CU(cudaMemcpy2DAsync(g.img + (size_t)first_slot * g.slot_stride, g.pitch, src, (size_t)row_pitch, g.w, (size_t)g.h * n, kind,
                             (cudaStream_t)stream));
    } else {
        for (int i = 0; i < n; ++i)
            CU(cudaMemcpy2DAsync(g.img + (size_t)(first_slot + i) * g.slot_stride, g.pitch, src + (size_t)i * image_stride,
                                 (size_t)row_pitch, g.w, g.h, kind, (cudaStream_t)stream));
    }
    return JSFE_OK;
}

int jsfe_slot_image(jsfe_handle* h, int slot, uint8_t** dev_ptr, int64_t* pitch) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    if (dev_ptr) *dev_ptr = h->P.lv[0].img + (size_t)slot * h->P.lv[0].slot_stride;
    if (pitch) *pitch = h->P.lv[0].pitch;
    return JSFE_OK;
}

static int extract_chunk(jsfe_handle* h, int first_slot, int n, void* stream);

// results of this handle are about to be overwritten on `st`: a gather's pack of the previous results has to be through
static int wait_pending_pack(jsfe_handle* h, cudaStream_t st) {
    if (h->pack_pending) {
        cudaEvent_t e = h->pack_pending;
        h->pack_pending = nullptr;
        if (cudaStreamWaitEvent(st, e, 0) != cudaSuccess) return fail(JSFE_ERR_CUDA, "waiting for the gather's pack failed");
    }
    return JSFE_OK;
}

int jsfe_extract(jsfe_handle* h, int first_slot, int n, void* stream) {
    int rc = check_slots(h, first_slot, n);
    if (rc) return rc;
    if ((rc = wait_pending_pack(h, (cudaStream_t)stream))) return rc;
    // Large batches are cut into sub-batches whose pyramid + blurred pyramid stay L2-resident between kernels
    // (126 MB L2): the levels written by k_pyramid are then read by k_fast_cells / k_blur / k_orient_desc from L2.
    const int chunk = h->chunk_images > 0 ? h->chunk_images : n;
    for (int s = 0; s < n; s += chunk)
        if ((rc = extract_chunk(h, first_slot + s, std::min(chunk, n - s), stream))) return rc;
    return JSFE_OK;
}

static int extract_chunk(jsfe_handle* h, int first_slot, int n, void* stream) {
    int rc = JSFE_OK;
    if (n == 0) return JSFE_OK;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const jsfe::Params& P = h->P;
    if (P.pyr_blocks_total > 0) {
        StageTimer t(h, st, 0);
        const bool pdl = h->pdl && !h->profiling;
        CU(launch_k(pdl, jsfe::k_pyramid, dim3(P.pyr_blocks_total, n), dim3(256), 0, st, P, first_slot));
        if ((rc = post_launch(h, "k_pyramid"))) return rc;
    }
    // fork: the blur of every level depends only on the pyramid, not on FAST; when profiling per kernel, stay serial
    const bool overlap = h->overlap_blur && !h->profiling && P.blur_items_total > 0;
    cudaStream_t sb = overlap ? h->st_aux : st;
    if (overlap) {
        CU(cudaEventRecord(h->ev_fork, st));
        CU(cudaStreamWaitEvent(sb, h->ev_fork, 0));
    }
    {
        StageTimer t(h, st, 1);
        CU(launch_k(h->pdl && !h->profiling, h->fast_kernel, dim3(P.fast_items_total, n), dim3(256), h->fast_smem, st, P, h->tma, first_slot));
    }
    if ((rc = post_launch(h, "k_fast_cells"))) return rc;
    if (P.blur_items_total > 0) {
        {
            StageTimer t(h, sb, 7);
            jsfe::k_blur<<<dim3((P.blur_items_total + 255) / 256, n), 256, 0, sb>>>(P, first_slot);
        }
        if ((rc = post_launch(h, "k_blur"))) return rc;
        if (overlap) CU(cudaEventRecord(h->ev_join, sb));
    }
    if (h->cfg.apply_nms_ms && P.L > 1) {  // orb_gpu.cpp:665-712
        {
            StageTimer t(h, st, 6);
            if (h->cfg.nms_ms_mode_gpu) CU(launch_k(h->pdl && !h->profiling, jsfe::k_nms_ms_dense, dim3(n), dim3(1024), 0, st, P, first_slot));
            else CU(launch_k(h->pdl && !h->profiling, jsfe::k_nms_ms_buckets, dim3(n), dim3(256), (size_t)P.cap * 4, st, P, first_slot));
        }
        if ((rc = post_launch(h, "k_nms_ms"))) return rc;
    }
    {
        StageTimer t(h, st, 2);
        CU(launch_k(h->pdl && !h->profiling, jsfe::k_compact, dim3(n), dim3(1024), 0, st, P, first_slot));
    }
    if ((rc = post_launch(h, "k_compact"))) return rc;
    if (overlap) CU(cudaStreamWaitEvent(st, h->ev_join, 0));   // join: the descriptor samples the blurred levels
    {
        StageTimer t(h, st, 3);
        CU(launch_k(h->pdl && !h->profiling, jsfe::k_orient_desc, dim3((P.cap + 8 * JSFE_KP_PER_WARP - 1) / (8 * JSFE_KP_PER_WARP), n), dim3(256), 0,
                    st, P, h->tma, first_slot));
    }
    if ((rc = post_launch(h, "k_orient_desc"))) return rc;
    return JSFE_OK;
}

namespace {
jsfe::RightSide right_side_of(const jsfe_handle* hr, int left_mul, int left_add, int right_mul, int right_add) {
    jsfe::RightSide r;
    memset(&r, 0, sizeof r);
    r.kps = hr->P.kps;
    r.desc = hr->P.desc;
    r.row_start = hr->P.row_start;
    for (int i = 0; i < hr->P.L; ++i) r.img[i] = hr->P.lv[i].img;
    r.left_mul = left_mul; r.left_add = left_add; r.right_mul = right_mul; r.right_add = right_add;
    return r;
}

int launch_stereo(jsfe_handle* h, const jsfe::RightSide& r, int first_pair, int n, int th_high, int th_low, float mb, float mbf, cudaStream_t st) {
    int rc;
    const jsfe::Params& P = h->P;
    {
        StageTimer t(h, st, 4);
        CU(launch_k(h->pdl && !h->profiling, jsfe::k_stereo_match, dim3((P.cap + 7) / 8, n), dim3(256), 0, st, P, r, first_pair, th_high, th_low, mb, mbf));
    }
    if ((rc = post_launch(h, "k_stereo_match"))) return rc;
    StageTimer t(h, st, 5);
    CU(launch_k(h->pdl && !h->profiling, jsfe::k_stereo_outlier, dim3(n), dim3(1024), 0, st, P, first_pair, r.left_mul, r.left_add));
    return post_launch(h, "k_stereo_outlier");
}
}  // namespace

int jsfe_stereo_match(jsfe_handle* h, int first_pair, int n, int th_high, int th_low, float mb, float mbf, void* stream) {
    int rc = check_slots(h, 2 * first_pair, 2 * n);
    if (rc) return rc;
    if (n == 0) return JSFE_OK;
    if (!(mb > 0.0f) || th_high < 0 || th_high > 32767) return fail(JSFE_ERR_INVALID, "bad stereo parameters");
    CU(cudaSetDevice(h->device));
    if ((rc = wait_pending_pack(h, (cudaStream_t)stream))) return rc;
    return launch_stereo(h, right_side_of(h, 2, 0, 2, 1), first_pair, n, th_high, th_low, mb, mbf, (cudaStream_t)stream);
}

int jsfe_stereo_match_cross(jsfe_handle* hl, int slot_l, jsfe_handle* hr, int slot_r, int th_high, int th_low, float mb, float mbf,
                            void* stream) {
    int rc = check_slots(hl, slot_l, 1);
    if (rc) return rc;
    if ((rc = check_slots(hr, slot_r, 1))) return rc;
    if (!(mb > 0.0f) || th_high < 0 || th_high > 32767) return fail(JSFE_ERR_INVALID, "bad stereo parameters");
    const jsfe::Params &A = hl->P, &B = hr->P;
    bool same = hl->device == hr->device && A.L == B.L && A.cap == B.cap && A.n_tile_rows == B.n_tile_rows;
    for (int i = 0; same && i < A.L; ++i)
        same = A.lv[i].h == B.lv[i].h && A.lv[i].w == B.lv[i].w && A.lv[i].pitch == B.lv[i].pitch &&
               A.lv[i].slot_stride == B.lv[i].slot_stride && A.lv[i].tile_h == B.lv[i].tile_h && A.lv[i].tile_w == B.lv[i].tile_w;
    if (!same) return fail(JSFE_ERR_INVALID, "left and right handles must share device and geometry");
    CU(cudaSetDevice(hl->device));
    if ((rc = wait_pending_pack(hl, (cudaStream_t)stream))) return rc;
    return launch_stereo(hl, right_side_of(hr, 0, slot_l, 0, slot_r), 0, 1, th_high, th_low, mb, mbf, (cudaStream_t)stream);
}

int jsfe_slot_view_get(const jsfe_handle* h, int slot, jsfe_slot_view* out) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    if (!out) return fail(JSFE_ERR_INVALID, "null out");
    const jsfe::Params& P = h->P;
    const size_t cap = P.cap, s = slot;
    out->n_keypoints = P.n_kp + s;
    out->n_per_level = P.n_per_level + s * JSFE_MAXL;
    out->kps = P.kps + s * 6 * cap;
    out->desc = P.desc + s * cap * 32;
    out->u_right = P.u_right + s * cap;
    out->depth = P.depth + s * cap;
    out->best_idx_r = P.best_idx + s * cap;
    out->best_dist = P.best_dist + s * cap;
    out->capacity = P.cap;
    return JSFE_OK;
}

int jsfe_level_image(const jsfe_handle* h, int slot, int level, const uint8_t** dev_ptr, int32_t* height, int32_t* width, int64_t* pitch) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    if (level < 0 || level >= h->P.L) return fail(JSFE_ERR_INVALID, "bad level");
    const jsfe::LevelGeom& g = h->P.lv[level];
    if (dev_ptr) *dev_ptr = g.img + (size_t)slot * g.slot_stride;
    if (height) *height = g.h;
    if (width) *width = g.w;
    if (pitch) *pitch = g.pitch;
    return JSFE_OK;
}

int jsfe_pack_keypoints(jsfe_handle* h, int slot, int32_t* dst_kps_dev, uint8_t* dst_desc_dev, int32_t* n_out, void* stream) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaMemcpyAsync(h->h_n, h->P.n_kp + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int n = h->h_n[0];
    if (n_out) *n_out = n;
    if (n > 0 && (dst_kps_dev || dst_desc_dev)) {
        jsfe::k_pack<<<(8 * n + 255) / 256, 256, 0, st>>>(h->P, slot, n, dst_kps_dev, dst_desc_dev);
        if ((rc = post_launch(h, "k_pack"))) return rc;
    }
    return JSFE_OK;
}

int jsfe_host_alloc(void** ptr, size_t bytes, int write_combined) {
    if (!ptr || bytes == 0) return fail(JSFE_ERR_INVALID, "jsfe_host_alloc: null pointer or zero size");
    *ptr = nullptr;
    if (cudaHostAlloc(ptr, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)) != cudaSuccess) {
        cudaGetLastError();
        return fail(JSFE_ERR_CUDA, "cudaHostAlloc(%zu bytes%s) failed", bytes, write_combined ? ", write-combined" : "");
    }
    return JSFE_OK;
}

int jsfe_host_free(void* ptr) {
    if (ptr && cudaFreeHost(ptr) != cudaSuccess) {
        cudaGetLastError();
        return fail(JSFE_ERR_CUDA, "cudaFreeHost failed");
    }
    return JSFE_OK;
}

int jsfe_pack_keypoints_once(jsfe_handle* h, int slot, int32_t* dst_kps_dev, uint8_t* dst_desc_dev, int32_t* n_out, void* stream) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    if (!dst_kps_dev || !dst_desc_dev || !n_out) return fail(JSFE_ERR_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    jsfe::k_pack<<<(8 * h->P.cap + 255) / 256, 256, 0, st>>>(h->P, slot, -1, dst_kps_dev, dst_desc_dev);
    if ((rc = post_launch(h, "k_pack"))) return rc;
    CU(cudaMemcpyAsync(h->h_n, h->P.n_kp + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_out = h->h_n[0];
    return JSFE_OK;
}

int jsfe_get_keypoints(jsfe_handle* h, int slot, int32_t* kps_host, uint8_t* desc_host, int32_t* n_out, void* stream) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cap = h->P.cap;
    CU(cudaMemcpyAsync(h->h_n, h->P.n_kp + slot, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_kps, h->P.kps + (size_t)slot * 6 * cap, 6 * cap * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_desc, h->P.desc + (size_t)slot * cap * 32, cap * 32, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int n = h->h_n[0];
    if (n_out) *n_out = n;
    if (kps_host)
        for (int pl = 0; pl < 6; ++pl) memcpy(kps_host + (size_t)pl * n, h->h_kps + (size_t)pl * cap, (size_t)n * sizeof(int32_t));
    if (desc_host) memcpy(desc_host, h->h_desc, (size_t)n * 32);
    return JSFE_OK;
}

int jsfe_get_stereo(jsfe_handle* h, int pair, float* u_right_host, float* depth_host, int32_t* best_idx_r_host,
                    int32_t* best_dist_host, int32_t* n_left_out, void* stream) {
    int rc = check_slots(h, 2 * pair, 2);
    if (rc) return rc;
    return jsfe_get_stereo_slot(h, 2 * pair, u_right_host, depth_host, best_idx_r_host, best_dist_host, n_left_out, stream);
}

int jsfe_get_stereo_slot(jsfe_handle* h, int left_slot, float* u_right_host, float* depth_host, int32_t* best_idx_r_host,
                         int32_t* best_dist_host, int32_t* n_left_out, void* stream) {
    int rc = check_slots(h, left_slot, 1);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cap = h->P.cap, s = (size_t)left_slot;
    CU(cudaMemcpyAsync(h->h_n, h->P.n_kp + s, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_misc, h->P.u_right + s * cap, cap * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_misc + cap, h->P.depth + s * cap, cap * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_misc + 2 * cap, h->P.best_idx + s * cap, cap * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_misc + 3 * cap, h->P.best_dist + s * cap, cap * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const int n = h->h_n[0];
    if (n_left_out) *n_left_out = n;
    if (u_right_host) memcpy(u_right_host, h->h_misc, (size_t)n * 4);
    if (depth_host) memcpy(depth_host, h->h_misc + cap, (size_t)n * 4);
    if (best_idx_r_host) memcpy(best_idx_r_host, h->h_misc + 2 * cap, (size_t)n * 4);
    if (best_dist_host) memcpy(best_dist_host, h->h_misc + 3 * cap, (size_t)n * 4);
    return JSFE_OK;
}

int jsfe_download_results(jsfe_handle* h, int first_slot, int n, jsfe_host_results* out, void* stream) {
    int rc = check_slots(h, first_slot, n);
    if (rc) return rc;
    if (!out) return fail(JSFE_ERR_INVALID, "null out");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cap = h->P.cap, f = first_slot, N = n;
    CU(cudaMemcpyAsync(h->h_n, h->P.n_kp + f, N * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_kps, h->P.kps + f * 6 * cap, N * 6 * cap * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_desc, h->P.desc + f * cap * 32, N * cap * 32, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_ur, h->P.u_right + f * cap, N * cap * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h->h_dp, h->P.depth + f * cap, N * cap * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    out->n_keypoints = h->h_n;
    out->kps = h->h_kps;
    out->desc = h->h_desc;
    out->u_right = h->h_ur;
    out->depth = h->h_dp;
    out->capacity = h->P.cap;
    out->bytes = (int64_t)(N * 4 + N * 6 * cap * 4 + N * cap * 32 + 2 * N * cap * 4);
    return JSFE_OK;
}

int jsfe_profile_enable(jsfe_handle* h, int on) {
    if (!h) return fail(JSFE_ERR_INVALID, "null handle");
    h->profiling = on != 0;
    return JSFE_OK;
}

int jsfe_profile_read(jsfe_handle* h, float* stage_ms, int64_t* stage_launches, int n_stages) {
    if (!h || !stage_ms || !stage_launches || n_stages < 1) return fail(JSFE_ERR_INVALID, "bad argument");
    CU(cudaSetDevice(h->device));
    for (int i = 0; i < n_stages; ++i) { stage_ms[i] = 0.f; stage_launches[i] = 0; }
    for (auto& sp : h->spans) {
        CU(cudaEventSynchronize(sp.b));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, sp.a, sp.b));
        if (sp.stage < n_stages) { stage_ms[sp.stage] += ms; ++stage_launches[sp.stage]; }
        h->event_pool.push_back(sp.a);
        h->event_pool.push_back(sp.b);
    }
    h->spans.clear();
    return JSFE_OK;
}

int jsfe_process_host_pairs_begin(jsfe_handle* h, int n_pairs, const uint8_t* images, int chunk_pairs, int th_high, int th_low,
                                  float mb, float mbf) {
    int rc = check_slots(h, 0, 2 * n_pairs);
    if (rc) return rc;
    if (!images || n_pairs < 1) return fail(JSFE_ERR_INVALID, "bad argument");
    if (h->pending_pairs) return fail(JSFE_ERR_INVALID, "a batch is already in flight on this handle (call jsfe_process_host_pairs_end)");
    CU(cudaSetDevice(h->device));
    const jsfe::Params& P = h->P;
    const jsfe::LevelGeom& g = P.lv[0];
    const size_t img_bytes = (size_t)g.h * g.w;
    if (!h->pipe_ready) {  // first use: staging buffer in the host layout (contiguous rows) and the three streams; all or nothing
        cudaError_t e = cudaMalloc((void**)&h->d_stage, img_bytes * h->max_images);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st_h2d, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st_comp, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->st_d2h, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            if (h->d_stage) { cudaFree(h->d_stage); h->d_stage = nullptr; }
            if (h->st_h2d) { cudaStreamDestroy(h->st_h2d); h->st_h2d = nullptr; }
            if (h->st_comp) { cudaStreamDestroy(h->st_comp); h->st_comp = nullptr; }
            if (h->st_d2h) { cudaStreamDestroy(h->st_d2h); h->st_d2h = nullptr; }
            return fail(JSFE_ERR_CUDA, "pipeline set-up failed: %s", cudaGetErrorString(e));
        }
        h->pipe_ready = true;
    }
    if ((rc = wait_pending_pack(h, h->st_comp))) return rc;
    if (chunk_pairs < 1) chunk_pairs = 32;
    // chunk schedule: a short first chunk (C/4, then C/2) so that compute starts early, full chunks after, and a short
    // last chunk (C/2, C/4) so that little D2H is left when compute ends
    std::vector<int> sizes;
    {
        int left = n_pairs;
        const int ramp[2] = {std::max(1, chunk_pairs / 4), std::max(1, chunk_pairs / 2)};
        // the ramp needs room for its head and its tail plus at least one pair in between
        const bool ramped = n_pairs >= 3 * chunk_pairs && n_pairs > 2 * (ramp[0] + ramp[1]) && !getenv("JSFE_NO_RAMP");
        if (ramped)
            for (int r = 0; r < 2; ++r) { sizes.push_back(ramp[r]); left -= ramp[r]; }
        const int tail = ramped ? ramp[0] + ramp[1] : 0;
        while (left - tail > 0) { const int c = std::min(chunk_pairs, left - tail); sizes.push_back(c); left -= c; }
        if (ramped)
            for (int r = 1; r >= 0; --r) { sizes.push_back(ramp[r]); left -= ramp[r]; }
        int total = 0;
        for (int c : sizes) total += c;
        if (left != 0 || total != n_pairs) return fail(JSFE_ERR_INVALID, "internal: chunk schedule covers %d of %d pairs", total, n_pairs);
    }
    const int n_chunks = (int)sizes.size();
    while ((int)h->ev_up.size() < n_chunks) {
        cudaEvent_t a, b;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        h->ev_up.push_back(a);
        h->ev_done.push_back(b);
    }
    const size_t cap = P.cap;
    struct ProfilingOff {   // per-kernel event spans are not recorded inside the pipeline; restored on every exit path
        jsfe_handle* h;
        bool saved;
        explicit ProfilingOff(jsfe_handle* hh) : h(hh), saved(hh->profiling) { hh->profiling = false; }
        ~ProfilingOff() { h->profiling = saved; }
    } profiling_off(h);
    auto enqueue_results = [&](int s0, int ns, cudaStream_t st) -> int {
        if (s0 == 0 && ns == h->max_images) {   // the whole handle (the single-pair latency path): one copy of the result arena
            CU(cudaMemcpyAsync(h->h_arena, h->d_arena, h->arena_bytes, cudaMemcpyDeviceToHost, st));
            return JSFE_OK;
        }
        CU(cudaMemcpyAsync(h->h_n + s0, P.n_kp + s0, (size_t)ns * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h->h_kps + (size_t)s0 * 6 * cap, P.kps + (size_t)s0 * 6 * cap, (size_t)ns * 6 * cap * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h->h_desc + (size_t)s0 * cap * 32, P.desc + (size_t)s0 * cap * 32, (size_t)ns * cap * 32, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h->h_ur + (size_t)s0 * cap, P.u_right + (size_t)s0 * cap, (size_t)ns * cap * 4, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(h->h_dp + (size_t)s0 * cap, P.depth + (size_t)s0 * cap, (size_t)ns * cap * 4, cudaMemcpyDeviceToHost, st));
        return JSFE_OK;
    };
    bool done_by_graph = false;
    if (n_chunks == 1 && !h->pg_disabled && !getenv("JSFE_NO_GRAPH")) {
        // latency path (one frame at a time): everything after the upload is one graph launch
        const bool hit = h->pair_graph && h->pg_pairs == n_pairs && h->pg_th_high == th_high && h->pg_th_low == th_low &&
                         h->pg_mb == mb && h->pg_mbf == mbf;
        if (!hit) {
            if (h->pair_graph) { cudaGraphExecDestroy(h->pair_graph); h->pair_graph = nullptr; }
            cudaGraph_t g_cap = nullptr;
            bool ok = cudaStreamBeginCapture(h->st_comp, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                ok = jsfe_set_images(h, 0, 2 * n_pairs, h->d_stage, g.w, (int64_t)img_bytes, 1, h->st_comp) == JSFE_OK &&
                     jsfe_extract(h, 0, 2 * n_pairs, h->st_comp) == JSFE_OK &&
                     jsfe_stereo_match(h, 0, n_pairs, th_high, th_low, mb, mbf, h->st_comp) == JSFE_OK &&
                     enqueue_results(0, 2 * n_pairs, h->st_comp) == JSFE_OK;
                ok = (cudaStreamEndCapture(h->st_comp, &g_cap) == cudaSuccess) && ok && g_cap;
            }
            if (ok) ok = cudaGraphInstantiate(&h->pair_graph, g_cap, 0) == cudaSuccess;
            if (g_cap) cudaGraphDestroy(g_cap);
            if (!ok) {   // capture is an optimisation: fall back to stream launches for good
                cudaGetLastError();
                h->pair_graph = nullptr;
                h->pg_disabled = true;
            } else {
                h->pg_pairs = n_pairs; h->pg_th_high = th_high; h->pg_th_low = th_low; h->pg_mb = mb; h->pg_mbf = mbf;
            }
        }
        if (h->pair_graph) {
            cudaError_t e = cudaMemcpyAsync(h->d_stage, images, img_bytes * 2 * n_pairs, cudaMemcpyHostToDevice, h->st_comp);
            if (e == cudaSuccess) e = cudaGraphLaunch(h->pair_graph, h->st_comp);
            if (e != cudaSuccess) {   // drain whatever was enqueued, then report
                cudaStreamSynchronize(h->st_comp);
                return fail(JSFE_ERR_CUDA, "graph launch failed: %s", cudaGetErrorString(e));
            }
            done_by_graph = true;
        }
    }
    // from here on nothing returns before the drain below: a failed enqueue must not leave work in flight with pending_pairs == 0
    auto cu = [&](cudaError_t e, const char* what) -> bool {
        if (e == cudaSuccess) return true;
        rc = fail(JSFE_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        return false;
    };
    for (int c = 0, p0 = 0; c < n_chunks && !done_by_graph; p0 += sizes[c], ++c) {
        const int np = sizes[c];
        const int s0 = 2 * p0, ns = 2 * np;
        // 1. one contiguous H2D per chunk (2-D copies with odd row widths run far below PCIe speed)
        if (!cu(cudaMemcpyAsync(h->d_stage + img_bytes * s0, images + img_bytes * s0, img_bytes * ns, cudaMemcpyHostToDevice, h->st_h2d), "upload") ||
            !cu(cudaEventRecord(h->ev_up[c], h->st_h2d), "cudaEventRecord"))
            break;
        // 2. re-pitch into the slots (device-to-device), extract, match
        if (!cu(cudaStreamWaitEvent(h->st_comp, h->ev_up[c], 0), "cudaStreamWaitEvent")) break;
        if ((rc = jsfe_set_images(h, s0, ns, h->d_stage + img_bytes * s0, g.w, (int64_t)img_bytes, 1, h->st_comp))) break;
        if ((rc = jsfe_extract(h, s0, ns, h->st_comp))) break;
        if ((rc = jsfe_stereo_match(h, p0, np, th_high, th_low, mb, mbf, h->st_comp))) break;
        if (!cu(cudaEventRecord(h->ev_done[c], h->st_comp), "cudaEventRecord")) break;
        // 3. results of this chunk back to pinned host memory while the next chunk computes
        if (!cu(cudaStreamWaitEvent(h->st_d2h, h->ev_done[c], 0), "cudaStreamWaitEvent")) break;
        if ((rc = enqueue_results(s0, ns, h->st_d2h))) break;
    }
    if (rc) {   // something failed while enqueueing: drain what was enqueued and report
        cudaStreamSynchronize(h->st_h2d); cudaStreamSynchronize(h->st_comp); cudaStreamSynchronize(h->st_d2h);
        return rc;
    }
    h->pending_pairs = n_pairs;
    return JSFE_OK;
}

int jsfe_process_host_pairs_end(jsfe_handle* h, jsfe_host_results* out) {
    if (!h || !out) return fail(JSFE_ERR_INVALID, "bad argument");
    if (!h->pending_pairs) return fail(JSFE_ERR_INVALID, "no batch in flight on this handle");
    CU(cudaSetDevice(h->device));
    const int n_pairs = h->pending_pairs;
    h->pending_pairs = 0;
    cudaError_t e1 = cudaStreamSynchronize(h->st_h2d), e2 = cudaStreamSynchronize(h->st_comp), e3 = cudaStreamSynchronize(h->st_d2h);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return fail(JSFE_ERR_CUDA, "pipeline failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
    const size_t N = (size_t)2 * n_pairs, cap = (size_t)h->P.cap;
    out->n_keypoints = h->h_n;
    out->kps = h->h_kps;
    out->desc = h->h_desc;
    out->u_right = h->h_ur;
    out->depth = h->h_dp;
    out->capacity = h->P.cap;
    out->bytes = (int64_t)(N * 4 + N * 6 * cap * 4 + N * cap * 32 + 2 * N * cap * 4);
    return JSFE_OK;
}

int jsfe_process_host_pairs(jsfe_handle* h, int n_pairs, const uint8_t* images, int chunk_pairs, int th_high, int th_low,
                            float mb, float mbf, jsfe_host_results* out) {
    if (!out) return fail(JSFE_ERR_INVALID, "bad argument");
    const int rc = jsfe_process_host_pairs_begin(h, n_pairs, images, chunk_pairs, th_high, th_low, mb, mbf);
    return rc ? rc : jsfe_process_host_pairs_end(h, out);
}

int jsfe_project_points(int n, const float* px, const float* py, const float* pz, const float* rcw9, const float* tcw3, float fx,
                        float fy, float cx, float cy, float min_x, float max_x, float min_y, float max_y, float* u, float* v,
                        float* invz, uint8_t* is_valid, void* stream) {
    if (n < 0 || (n && (!px || !py || !pz || !rcw9 || !tcw3 || !u || !v || !invz || !is_valid))) return fail(JSFE_ERR_INVALID, "bad argument");
    if (n == 0) return JSFE_OK;
    jsfe::k_project_points<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, px, py, pz, rcw9, tcw3, fx, fy, cx, cy, min_x, max_x,
                                                                                 min_y, max_y, u, v, invz, is_valid);
    CU(cudaGetLastError());
    return JSFE_OK;
}

int jsfe_hamming_pairs(int n, const int32_t* idx_left, const int32_t* idx_right, const uint8_t* desc_left,
                       const uint8_t* desc_right, int32_t* distance, void* stream) {
    if (n < 0 || (n && (!idx_left || !idx_right || !desc_left || !desc_right || !distance))) return fail(JSFE_ERR_INVALID, "bad argument");
    if (n == 0) return JSFE_OK;
    jsfe::k_hamming_pairs<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, idx_left, idx_right, desc_left, desc_right, distance);
    CU(cudaGetLastError());
    return JSFE_OK;
}

int jsfe_in_frustum(int n, const float* px, const float* py, const float* pz, const float* pnx, const float* pny, const float* pnz,
                    const float* max_distance, const float* invariance_max_distance, const float* invariance_min_distance,
                    const float* rcw9, const float* tcw3, const float* ow3, float fx, float fy, float cx, float cy, int min_x,
                    int max_x, int min_y, int max_y, int n_scale_levels, float log_scale_factor, float view_cos_angle, float* invz,
                    float* u, float* v, int32_t* predicted_level, float* view_cos, uint8_t* is_infrustum, void* stream) {
    if (n < 0 || (n && (!px || !py || !pz || !pnx || !pny || !pnz || !max_distance || !invariance_max_distance || !invariance_min_distance || !rcw9 ||
                        !tcw3 || !ow3 || !invz || !u || !v || !predicted_level || !view_cos || !is_infrustum)))
        return fail(JSFE_ERR_INVALID, "bad argument");
    if (n == 0) return JSFE_OK;
    jsfe::k_in_frustum<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(n, px, py, pz, pnx, pny, pnz, max_distance, invariance_max_distance,
                                                                            invariance_min_distance, rcw9, tcw3, ow3, fx, fy, cx, cy, min_x, max_x,
                                                                            min_y, max_y, n_scale_levels, log_scale_factor, view_cos_angle, invz, u, v,
                                                                            predicted_level, view_cos, is_infrustum, nullptr);
    CU(cudaGetLastError());
    return JSFE_OK;
}

int jsfe_build_frame_grid(int n_cur, const float* cur_x, const float* cur_y, float min_x, float max_x, float min_y, float max_y,
                          int32_t* cell_start, int32_t* cell_items, void* stream) {
    if (n_cur < 0 || n_cur > 65535 || !cell_start || (n_cur && (!cur_x || !cur_y || !cell_items)) || !(max_x > min_x) || !(max_y > min_y))
        return fail(JSFE_ERR_INVALID, "bad argument");
    const float winv = (float)JSFE_FRAME_GRID_COLS / (max_x - min_x), hinv = (float)JSFE_FRAME_GRID_ROWS / (max_y - min_y);  // Frame.cpp:234-235
    jsfe::k_frame_grid<<<1, 1024, 0, (cudaStream_t)stream>>>(n_cur, cur_x, cur_y, min_x, min_y, winv, hinv, cell_start, cell_items);
    CU(cudaGetLastError());
    return JSFE_OK;
}

int jsfe_search_by_projection(const jsfe_sbp_args* a, void* stream) {
    if (!a || a->n_last < 0 || a->n_cur < 0 || a->n_cur > 65535 || a->th_high < 0 || a->th_high >= 256 || a->level_mode < 0 ||
        a->level_mode > 2 || !a->n_matches || !a->hist || !(a->max_x > a->min_x) || !(a->max_y > a->min_y))
        return fail(JSFE_ERR_INVALID, "bad argument");
    if (a->n_last && (!a->px || !a->py || !a->pz || !a->last_octave || !a->last_angle || !a->last_desc || !a->rcw9 || !a->tcw3 ||
                      !a->best_idx2 || !a->best_dist || !a->rot_bin))
        return fail(JSFE_ERR_INVALID, "null last-frame array");
    if (a->n_cur && (!a->cur_x || !a->cur_y || !a->cur_octave || !a->cur_angle || !a->cur_uright || !a->cur_desc || !a->cell_start ||
                     !a->cell_items || !a->cur_match))
        return fail(JSFE_ERR_INVALID, "null current-frame array");
    cudaStream_t st = (cudaStream_t)stream;
    if (a->n_cur) CU(cudaMemsetAsync(a->cur_match, 0xFF, (size_t)a->n_cur * sizeof(int32_t), st));   // -1
    CU(cudaMemsetAsync(a->hist, 0, JSFE_HISTO_LENGTH * sizeof(int32_t), st));
    CU(cudaMemsetAsync(a->n_matches, 0, sizeof(int32_t), st));
    if (a->n_last == 0) return JSFE_OK;
    if (a->n_cur == 0) {   // nothing to match against: every point reports "none"
        CU(cudaMemsetAsync(a->best_idx2, 0xFF, (size_t)a->n_last * sizeof(int32_t), st));
        CU(cudaMemsetAsync(a->rot_bin, 0xFF, (size_t)a->n_last * sizeof(int32_t), st));
        std::vector<int32_t> none((size_t)a->n_last, 256);
        CU(cudaMemcpyAsync(a->best_dist, none.data(), none.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));   // `none` is a stack-owned buffer
        return JSFE_OK;
    }
    jsfe::k_sbp_match<<<(a->n_last + 7) / 8, 256, 0, st>>>(*a);
    CU(cudaGetLastError());
    if (a->check_orientation) {
        jsfe::k_sbp_finish<<<1, 1024, 0, st>>>(*a);
        CU(cudaGetLastError());
    }
    return JSFE_OK;
}

int jsfe_frame_view(jsfe_handle* h, int slot, jsfe_cv_keypoint* keys, float* x, float* y, int32_t* octave, float* angle, void* stream) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    jsfe::k_frame_view<<<(h->P.cap + 255) / 256, 256, 0, (cudaStream_t)stream>>>(h->P, slot, keys, x, y, octave, angle);
    CU(cudaGetLastError());
    return JSFE_OK;
}

int jsfe_remap_bilinear(const uint8_t* src, int src_h, int src_w, int64_t src_pitch, int64_t src_stride, int n_images,
                        const float* map_x, const float* map_y, int dst_h, int dst_w, uint8_t* dst, int64_t dst_pitch,
                        int64_t dst_stride, void* stream) {
    if (!src || !map_x || !map_y || !dst || src_h < 1 || src_w < 1 || dst_h < 1 || dst_w < 1 || src_pitch < src_w || dst_pitch < dst_w ||
        n_images < 0 || src_h > 32767 || src_w > 32767)
        return fail(JSFE_ERR_INVALID, "bad argument");
    if (n_images == 0) return JSFE_OK;
    const int word_stores = ((uintptr_t)dst % 4 == 0) && (dst_pitch % 4 == 0) && (dst_stride % 4 == 0);
    const int word_loads = ((uintptr_t)src % 4 == 0) && (src_pitch % 4 == 0) && (src_stride % 4 == 0);   // the window path reads aligned words
    // images in flight per output tile: enough blocks to fill the GPU, few enough that the decoded maps are reused
    const int tiles = ((dst_w + 255) / 256) * ((dst_h + 3) / 4);
    const int groups = (n_images + JSFE_REMAP_UNROLL - 1) / JSFE_REMAP_UNROLL;   // a thread handles its images JSFE_REMAP_UNROLL at a time
    const int gz = std::max(1, std::min(groups, (148 * 8 + tiles - 1) / tiles));
    jsfe::k_remap_bilinear<<<dim3((dst_w + 255) / 256, (dst_h + 3) / 4, gz), 256, 0, (cudaStream_t)stream>>>(
        src, src_h, src_w, (long long)src_pitch, (long long)src_stride, n_images, map_x, map_y, dst_h, dst_w, dst, (long long)dst_pitch,
        (long long)dst_stride, word_stores, word_loads);
    CU(cudaGetLastError());
    return JSFE_OK;
}

int jsfe_cvt_gray(const uint8_t* src, int h, int w, int64_t src_pitch, int channels, int rgb_order, uint8_t* dst, int64_t dst_pitch,
                  void* stream) {
    if (!src || !dst || h < 1 || w < 1 || (channels != 3 && channels != 4) || src_pitch < (int64_t)w * channels || dst_pitch < w || h > 65535)
        return fail(JSFE_ERR_INVALID, "bad argument");
    jsfe::k_cvt_gray<<<dim3((w + 255) / 256, h), 256, 0, (cudaStream_t)stream>>>(src, h, w, (long long)src_pitch, channels,
                                                                               rgb_order ? 2 : 0, dst, (long long)dst_pitch);
    CU(cudaGetLastError());
    return JSFE_OK;
}

int jsfe_debug_level_image(jsfe_handle* h, int slot, int level, uint8_t* host_dst) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    if (level < 0 || level >= h->P.L || !host_dst) return fail(JSFE_ERR_INVALID, "bad level / dst");
    CU(cudaSetDevice(h->device));
    const jsfe::LevelGeom& g = h->P.lv[level];
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy2D(host_dst, g.w, g.img + (size_t)slot * g.slot_stride, g.pitch, g.w, g.h, cudaMemcpyDeviceToHost));
    return JSFE_OK;
}

int jsfe_debug_level_blur(jsfe_handle* h, int slot, int level, uint8_t* host_dst) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    if (level < 0 || level >= h->P.L || !host_dst) return fail(JSFE_ERR_INVALID, "bad level / dst");
    CU(cudaSetDevice(h->device));
    const jsfe::LevelGeom& g = h->P.lv[level];
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy2D(host_dst, g.w, g.blur + (size_t)slot * g.slot_stride, g.pitch, g.w, g.h, cudaMemcpyDeviceToHost));
    return JSFE_OK;
}

int jsfe_debug_cells(jsfe_handle* h, int slot, int32_t* x, int32_t* y, int32_t* score) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    const size_t cap = h->P.cap, o = (size_t)slot * cap;
    if (x) CU(cudaMemcpy(x, h->P.cell_x + o, cap * 4, cudaMemcpyDeviceToHost));
    if (y) CU(cudaMemcpy(y, h->P.cell_y + o, cap * 4, cudaMemcpyDeviceToHost));
    if (score) CU(cudaMemcpy(score, h->P.cell_s + o, cap * 4, cudaMemcpyDeviceToHost));
    return JSFE_OK;
}

int jsfe_debug_level_keypoints(jsfe_handle* h, int slot, int32_t* x, int32_t* y, int32_t* score, int32_t* level, float* angle_rad) {
    int rc = check_slots(h, slot, 1);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    CU(cudaDeviceSynchronize());
    const size_t cap = h->P.cap, o = (size_t)slot * cap;
    if (x) CU(cudaMemcpy(x, h->P.kp_x + o, cap * 4, cudaMemcpyDeviceToHost));
    if (y) CU(cudaMemcpy(y, h->P.kp_y + o, cap * 4, cudaMemcpyDeviceToHost));
    if (score) CU(cudaMemcpy(score, h->P.kp_s + o, cap * 4, cudaMemcpyDeviceToHost));
    if (level) CU(cudaMemcpy(level, h->P.kp_l + o, cap * 4, cudaMemcpyDeviceToHost));
    if (angle_rad) CU(cudaMemcpy(angle_rad, h->P.kp_angle + o, cap * 4, cudaMemcpyDeviceToHost));
    return JSFE_OK;
}


// ================================================================================================= SURVEY.md 8(e): gather
// NCCL is resolved at run time (dlopen): libjsfe.so has no link-time dependency on it, and inside a process that already loaded
// an NCCL (torch's bundled one) the same library -- hence the same ncclComm_t -- is found by its soname.
typedef struct ncclComm* jsfe_nccl_comm_t;
struct NcclApi {
    void* lib = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, jsfe_nccl_comm_t, cudaStream_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, jsfe_nccl_comm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, jsfe_nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
    if (g_nccl.lib) return JSFE_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(JSFE_ERR_INVALID, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
    NcclApi a;
    a.lib = lib;
    *(void**)&a.AllReduce = dlsym(lib, "ncclAllReduce");
    *(void**)&a.Send = dlsym(lib, "ncclSend");
    *(void**)&a.Recv = dlsym(lib, "ncclRecv");
    *(void**)&a.GroupStart = dlsym(lib, "ncclGroupStart");
    *(void**)&a.GroupEnd = dlsym(lib, "ncclGroupEnd");
    *(void**)&a.GetErrorString = dlsym(lib, "ncclGetErrorString");
    if (!a.AllReduce || !a.Send || !a.Recv || !a.GroupStart || !a.GroupEnd || !a.GetErrorString)
        return fail(JSFE_ERR_INVALID, "libnccl lacks an expected symbol");
    g_nccl = a;
    return JSFE_OK;
}
#define NC(call)                                                                                                  \
    do {                                                                                                          \
        int r__ = (call);                                                                                         \
        if (r__ != 0) return fail(JSFE_ERR_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r__));              \
    } while (0)

struct jsfe_gather {
    jsfe_handle* h = nullptr;
    jsfe_nccl_comm_t comm = nullptr;
    int rank = 0, world = 1, root = 0, max_pairs = 0;
    size_t region_bytes = 0;            // stride between rank regions (capacity bound, 256-aligned)
    uint8_t* stage = nullptr;           // this rank's packed region in local HBM (the root packs straight into its landing buffer)
    uint8_t* landing[2] = {nullptr, nullptr};   // root: [world][region_bytes] x 2 (double buffer)
    uint8_t* peer[2] = {nullptr, nullptr};      // non-root: the root's landing buffers mapped through CUDA IPC (nullptr: NCCL transport)
    int32_t* token = nullptr;           // 2 ints: all-reduce send / receive
    cudaStream_t st = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_packed = nullptr, ev_done[2] = {nullptr, nullptr};
    long long seq = 0;
    int last_pairs = 0, in_flight = 0;
    bool peer_all_mapped = false;       // root: every other rank writes its region through a peer mapping (no NCCL receive posted)
    // optional stage timing (jsfe_gather_profile): events on the gather stream in front of the pack and behind every stage
    bool timing = false;
    cudaEvent_t ev_t[2][5] = {{nullptr, nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr, nullptr}};
    int last_ended = -1;                // buffer of the last gather jsfe_gather_end returned
};

int64_t jsfe_gather_region_bytes(const jsfe_handle* h, int max_pairs) {
    if (!h || max_pairs < 1) return fail(JSFE_ERR_INVALID, "bad argument");
    const size_t cap = (size_t)h->P.cap;
    return (int64_t)align_up(jsfe::gather_header_bytes(max_pairs) + (size_t)max_pairs * (jsfe::gather_slot_bytes((int)cap, 1) + jsfe::gather_slot_bytes((int)cap, 0)), 256);
}

int jsfe_gather_destroy(jsfe_gather* g) {
    if (!g) return JSFE_OK;
    cudaSetDevice(g->h->device);
    if (g->st) cudaStreamSynchronize(g->st);
    if (g->h->pack_pending == g->ev_packed) g->h->pack_pending = nullptr;   // the pack is through
    for (int b = 0; b < 2; ++b) {
        if (g->peer[b]) cudaIpcCloseMemHandle(g->peer[b]);
        if (g->landing[b]) cudaFree(g->landing[b]);
        if (g->ev_done[b]) cudaEventDestroy(g->ev_done[b]);
        for (cudaEvent_t e : g->ev_t[b])
            if (e) cudaEventDestroy(e);
    }
    if (g->stage) cudaFree(g->stage);
    if (g->token) cudaFree(g->token);
    if (g->ev_ready) cudaEventDestroy(g->ev_ready);
    if (g->ev_packed) cudaEventDestroy(g->ev_packed);
    if (g->st) cudaStreamDestroy(g->st);
    delete g;
    return JSFE_OK;
}

int jsfe_gather_create(jsfe_handle* h, void* nccl_comm, int rank, int world, int root, int max_pairs, jsfe_gather** out) {
    if (!h || !out || world < 1 || rank < 0 || rank >= world || root < 0 || root >= world || max_pairs < 1 || 2 * max_pairs > h->max_images)
        return fail(JSFE_ERR_INVALID, "bad gather geometry");
    *out = nullptr;
    if (world > 1) {
        if (!nccl_comm) return fail(JSFE_ERR_INVALID, "a communicator (ncclComm_t) is required for world > 1");
        int rc = nccl_load();
        if (rc) return rc;
    }
    CU(cudaSetDevice(h->device));
    jsfe_gather* g = new (std::nothrow) jsfe_gather;
    if (!g) return fail(JSFE_ERR_INVALID, "out of host memory");
    g->h = h; g->comm = (jsfe_nccl_comm_t)nccl_comm; g->rank = rank; g->world = world; g->root = root; g->max_pairs = max_pairs;
    g->region_bytes = (size_t)jsfe_gather_region_bytes(h, max_pairs);
    // highest priority: the pack, the stores and NCCL's small kernels take the first block slots that free up instead of queueing
    // behind the whole grid of the extraction kernel that runs beside them (measured: pack 530 -> see DESIGN.md section 8)
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaError_t e = getenv("JSFE_GATHER_NO_PRIORITY") ? cudaStreamCreateWithFlags(&g->st, cudaStreamNonBlocking)
                                                      : cudaStreamCreateWithPriority(&g->st, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_ready, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_packed, cudaEventDisableTiming);
    for (int b = 0; b < 2 && e == cudaSuccess; ++b) e = cudaEventCreateWithFlags(&g->ev_done[b], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc((void**)&g->token, 2 * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemset(g->token, 0, 2 * sizeof(int32_t));
    if (e == cudaSuccess && rank != root) e = cudaMalloc((void**)&g->stage, g->region_bytes);
    for (int b = 0; b < 2 && e == cudaSuccess && rank == root; ++b) {
        e = cudaMalloc((void**)&g->landing[b], g->region_bytes * (size_t)world);
        if (e == cudaSuccess) e = cudaMemset(g->landing[b], 0, g->region_bytes * (size_t)world);
    }
    if (e != cudaSuccess) {
        jsfe_gather_destroy(g);
        return fail(JSFE_ERR_CUDA, "gather set-up failed: %s", cudaGetErrorString(e));
    }
    CU(cudaDeviceSynchronize());
    *out = g;
    return JSFE_OK;
}

int jsfe_gather_ipc_export(jsfe_gather* g, int buffer, uint8_t handle64[64]) {
    if (!g || !handle64 || buffer < 0 || buffer > 1 || g->rank != g->root) return fail(JSFE_ERR_INVALID, "only the root exports its landing buffers");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    CU(cudaSetDevice(g->h->device));
    cudaIpcMemHandle_t hnd;
    CU(cudaIpcGetMemHandle(&hnd, g->landing[buffer]));
    memcpy(handle64, &hnd, 64);
    return JSFE_OK;
}

int jsfe_gather_ipc_import(jsfe_gather* g, int buffer, const uint8_t handle64[64]) {
    if (!g || !handle64 || buffer < 0 || buffer > 1 || g->rank == g->root) return fail(JSFE_ERR_INVALID, "only non-root ranks import");
    CU(cudaSetDevice(g->h->device));
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, handle64, 64);
    void* ptr = nullptr;
    CU(cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess));
    g->peer[buffer] = (uint8_t*)ptr;
    return JSFE_OK;
}

int jsfe_gather_begin(jsfe_gather* g, int first_pair, int n_pairs, void* compute_stream) {
    if (!g || n_pairs < 1 || n_pairs > g->max_pairs) return fail(JSFE_ERR_INVALID, "bad pair range for this gather");
    jsfe_handle* h = g->h;
    int rc = check_slots(h, 2 * first_pair, 2 * n_pairs);
    if (rc) return rc;
    if (g->in_flight) return fail(JSFE_ERR_INVALID, "a gather is already in flight (call jsfe_gather_end)");
    CU(cudaSetDevice(h->device));
    cudaStream_t cs = (cudaStream_t)compute_stream;
    const int b = (int)(g->seq & 1);
    const bool is_root = g->rank == g->root;
    const bool p2p = !is_root && g->peer[0] && g->peer[1];
    CU(cudaEventRecord(g->ev_ready, cs));
    CU(cudaStreamWaitEvent(g->st, g->ev_ready, 0));
    auto mark = [&](int i) { return g->timing ? cudaEventRecord(g->ev_t[b][i], g->st) : cudaSuccess; };
    CU(mark(0));
    // pack in local HBM (the root: straight into its own region of the landing buffer)
    uint8_t* region = is_root ? g->landing[b] + (size_t)g->rank * g->region_bytes : g->stage;
    jsfe::k_gather_pack<<<2 * n_pairs, 256, 0, g->st>>>(h->P, first_pair, n_pairs, g->rank, g->seq, region);
    if ((rc = post_launch(h, "k_gather_pack"))) return rc;
    CU(cudaEventRecord(g->ev_packed, g->st));
    CU(mark(1));
    h->pack_pending = g->ev_packed;   // the results may be overwritten once they are packed: the next extraction / match waits for it
    if (g->world > 1) {
        const size_t bound = jsfe::gather_header_bytes(n_pairs) + (size_t)n_pairs * (jsfe::gather_slot_bytes(h->P.cap, 1) + jsfe::gather_slot_bytes(h->P.cap, 0));
        // every rank uses the same transport: the peer mappings are made on all non-root ranks or on none (the root is told with
        // jsfe_gather_set_peers_mapped)
        const bool use_p2p = is_root ? g->peer_all_mapped : p2p;
        if (use_p2p) {
            // credit: the root joins this all-reduce in ITS jsfe_gather_begin of this batch, i.e. after its consumer released the landing
            // buffer of two batches ago -- no rank stores into that buffer earlier
            NC(g_nccl.AllReduce(g->token, g->token + 1, 1, 2 /* ncclInt32 */, 0 /* ncclSum */, g->comm, g->st));
            CU(mark(2));
            if (!is_root) {
                // a few blocks are enough (the stores only have to be through within a step) and, on the high-priority stream, all
                // that may be taken from the extraction running beside them: with 148 x 8 blocks the stores of seven ranks into one
                // root held every SM of every rank for 0.3 ms per step (8 GPUs: gather efficiency 0.93 instead of 0.96)
                const int blocks = (int)std::max<size_t>(1, std::min<size_t>((bound / 16 + 255) / 256, 32));
                jsfe::k_gather_put<<<blocks, 256, 0, g->st>>>(g->stage, g->peer[b] + (size_t)g->rank * g->region_bytes, n_pairs);
                if ((rc = post_launch(h, "k_gather_put"))) return rc;
            }
            CU(mark(3));
            // completion: behind this all-reduce every rank's stores of this batch have been performed (kernel completion makes them
            // visible system-wide), so the root may read all regions
            NC(g_nccl.AllReduce(g->token, g->token + 1, 1, 2 /* ncclInt32 */, 0 /* ncclSum */, g->comm, g->st));
        } else {
            // NCCL transport: one send/recv group, padded to the capacity bound (the trimmed size is known on the device only)
            CU(mark(2));
            NC(g_nccl.GroupStart());
            if (is_root) {
                for (int r = 0; r < g->world; ++r)
                    if (r != g->root) NC(g_nccl.Recv(g->landing[b] + (size_t)r * g->region_bytes, bound, 1 /* ncclUint8 */, r, g->comm, g->st));
            } else {
                NC(g_nccl.Send(g->stage, bound, 1 /* ncclUint8 */, g->root, g->comm, g->st));
            }
            NC(g_nccl.GroupEnd());
            CU(mark(3));
        }
    } else {
        CU(mark(2));
        CU(mark(3));
    }
    CU(mark(4));
    CU(cudaEventRecord(g->ev_done[b], g->st));
    g->last_pairs = n_pairs;
    g->in_flight = 1;
    ++g->seq;
    return JSFE_OK;
}

int jsfe_gather_end(jsfe_gather* g, jsfe_gathered* out) {
    if (!g || !out) return fail(JSFE_ERR_INVALID, "bad argument");
    if (!g->in_flight) return fail(JSFE_ERR_INVALID, "no gather in flight");
    CU(cudaSetDevice(g->h->device));
    const int b = (int)((g->seq - 1) & 1);
    g->in_flight = 0;
    CU(cudaEventSynchronize(g->ev_done[b]));
    g->last_ended = b;
    const bool is_root = g->rank == g->root;
    out->data = is_root ? g->landing[b] : nullptr;
    out->region_stride = (int64_t)g->region_bytes;
    out->world = g->world;
    out->root = g->root;
    out->n_pairs = g->last_pairs;
    out->transport = g->world == 1 ? 0 : (is_root ? (g->peer_all_mapped ? 1 : 2) : ((g->peer[0] && g->peer[1]) ? 1 : 2));
    return JSFE_OK;
}

int jsfe_gather_profile(jsfe_gather* g, int enable) {
    if (!g) return fail(JSFE_ERR_INVALID, "null gather");
    if (g->in_flight) return fail(JSFE_ERR_INVALID, "a gather is in flight");
    CU(cudaSetDevice(g->h->device));
    if (enable)
        for (int b = 0; b < 2; ++b)
            for (cudaEvent_t& e : g->ev_t[b])
                if (!e) CU(cudaEventCreate(&e));
    g->timing = enable != 0;
    g->last_ended = -1;
    return JSFE_OK;
}

int jsfe_gather_stage_times(jsfe_gather* g, float us[4]) {
    if (!g || !us) return fail(JSFE_ERR_INVALID, "bad argument");
    if (!g->timing || g->last_ended < 0) return fail(JSFE_ERR_INVALID, "no timed gather has ended (jsfe_gather_profile, then begin/end)");
    CU(cudaSetDevice(g->h->device));
    for (int i = 0; i < 4; ++i) {
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, g->ev_t[g->last_ended][i], g->ev_t[g->last_ended][i + 1]));
        us[i] = ms * 1000.0f;
    }
    return JSFE_OK;
}

int jsfe_gather_set_peers_mapped(jsfe_gather* g, int all_mapped) {
    if (!g) return fail(JSFE_ERR_INVALID, "null gather");
    g->peer_all_mapped = all_mapped != 0;
    return JSFE_OK;
}


// ================================================================================================= SURVEY.md 8(f2): resident map points
// Tracking::SearchLocalPoints (src/Tracking.cpp:1346-1806) re-packs nine float arrays of every local map point on the host and
// uploads them for each frame; map points change far less often than frames arrive.  The pool keeps that SoA on the device under
// stable slot ids: a frame sends its list of ids (4 bytes per point instead of 36) and the pose.
struct jsfe_mappool {
    int device = 0, capacity = 0;
    float* soa = nullptr;        // [9][capacity]: x y z | nx ny nz | max_distance | invariance_max | invariance_min
    float* stage = nullptr;      // device staging of an update: [9][capacity]
    int32_t* stage_ids = nullptr;
    float* h_stage = nullptr;    // pinned mirror of the staging area
    int32_t* h_ids = nullptr;
    float* pose = nullptr;       // 15 floats on the device: Rcw (9) | tcw (3) | Ow (3)
    float* h_pose = nullptr;
};

int jsfe_mappool_destroy(jsfe_mappool* p) {
    if (!p) return JSFE_OK;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    if (p->soa) cudaFree(p->soa);
    if (p->stage) cudaFree(p->stage);
    if (p->stage_ids) cudaFree(p->stage_ids);
    if (p->pose) cudaFree(p->pose);
    if (p->h_stage) cudaFreeHost(p->h_stage);
    if (p->h_ids) cudaFreeHost(p->h_ids);
    if (p->h_pose) cudaFreeHost(p->h_pose);
    delete p;
    return JSFE_OK;
}

int jsfe_mappool_create(int capacity, int device_id, jsfe_mappool** out) {
    if (!out || capacity < 1) return fail(JSFE_ERR_INVALID, "bad argument");
    *out = nullptr;
    CU(cudaSetDevice(device_id));
    jsfe_mappool* p = new (std::nothrow) jsfe_mappool;
    if (!p) return fail(JSFE_ERR_INVALID, "out of host memory");
    p->device = device_id; p->capacity = capacity;
    const size_t n9 = (size_t)9 * capacity * sizeof(float);
    cudaError_t e = cudaMalloc((void**)&p->soa, n9);
    if (e == cudaSuccess) e = cudaMemset(p->soa, 0, n9);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->stage, n9);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->stage_ids, (size_t)capacity * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&p->pose, 15 * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&p->h_stage, n9);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&p->h_ids, (size_t)capacity * 4);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&p->h_pose, 15 * sizeof(float));
    if (e != cudaSuccess) { jsfe_mappool_destroy(p); return fail(JSFE_ERR_CUDA, "map pool allocation failed: %s", cudaGetErrorString(e)); }
    *out = p;
    return JSFE_OK;
}

int jsfe_mappool_update(jsfe_mappool* p, int n, const int32_t* ids, const float* px, const float* py, const float* pz, const float* pnx,
                        const float* pny, const float* pnz, const float* max_distance, const float* invariance_max_distance,
                        const float* invariance_min_distance, void* stream) {
    if (!p || n < 0 || n > p->capacity || (n && (!ids || !px || !py || !pz || !pnx || !pny || !pnz || !max_distance || !invariance_max_distance || !invariance_min_distance)))
        return fail(JSFE_ERR_INVALID, "bad argument");
    if (n == 0) return JSFE_OK;
    for (int i = 0; i < n; ++i)
        if (ids[i] < 0 || ids[i] >= p->capacity) return fail(JSFE_ERR_CAPACITY, "map point id %d outside the pool of %d", ids[i], p->capacity);
    CU(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaStreamSynchronize(st));     // the pinned staging area is reused from call to call
    const float* src[9] = {px, py, pz, pnx, pny, pnz, max_distance, invariance_max_distance, invariance_min_distance};
    for (int a = 0; a < 9; ++a) memcpy(p->h_stage + (size_t)a * n, src[a], (size_t)n * sizeof(float));
    memcpy(p->h_ids, ids, (size_t)n * 4);
    CU(cudaMemcpyAsync(p->stage, p->h_stage, (size_t)9 * n * sizeof(float), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(p->stage_ids, p->h_ids, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    jsfe::k_mappool_scatter<<<(n + 255) / 256, 256, 0, st>>>(n, p->stage_ids, p->stage, n, p->soa, p->capacity);
    CU(cudaGetLastError());
    return JSFE_OK;
}

int jsfe_mappool_in_frustum(jsfe_mappool* p, int n, const int32_t* ids_dev, const float* rcw9_host, const float* tcw3_host, const float* ow3_host,
                            float fx, float fy, float cx, float cy, int min_x, int max_x, int min_y, int max_y, int n_scale_levels,
                            float log_scale_factor, float view_cos_angle, float* invz, float* u, float* v, int32_t* predicted_level,
                            float* view_cos, uint8_t* is_infrustum, void* stream) {
    if (!p || n < 0 || n > p->capacity || (n && (!ids_dev || !rcw9_host || !tcw3_host || !ow3_host || !invz || !u || !v || !predicted_level || !view_cos || !is_infrustum)))
        return fail(JSFE_ERR_INVALID, "bad argument");
    if (n == 0) return JSFE_OK;
    CU(cudaSetDevice(p->device));
    cudaStream_t st = (cudaStream_t)stream;
    CU(cudaStreamSynchronize(st));     // the pinned pose block is reused from call to call
    memcpy(p->h_pose, rcw9_host, 9 * sizeof(float));
    memcpy(p->h_pose + 9, tcw3_host, 3 * sizeof(float));
    memcpy(p->h_pose + 12, ow3_host, 3 * sizeof(float));
    CU(cudaMemcpyAsync(p->pose, p->h_pose, 15 * sizeof(float), cudaMemcpyHostToDevice, st));
    const size_t c = (size_t)p->capacity;
    const float* a = p->soa;
    jsfe::k_in_frustum<<<(n + 255) / 256, 256, 0, st>>>(n, a, a + c, a + 2 * c, a + 3 * c, a + 4 * c, a + 5 * c, a + 6 * c, a + 7 * c, a + 8 * c, p->pose,
                                                         p->pose + 9, p->pose + 12, fx, fy, cx, cy, min_x, max_x, min_y, max_y, n_scale_levels,
                                                         log_scale_factor, view_cos_angle, invz, u, v, predicted_level, view_cos, is_infrustum, ids_dev);
    CU(cudaGetLastError());
    return JSFE_OK;
}

}  // extern "C"
