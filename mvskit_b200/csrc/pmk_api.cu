// mvskit_b200/csrc/pmk_api.cu -- C ABI (include/pmk.h): context, per-view constants, kernel launches.
//
// Host-side float math in this file (camera constants, level table) follows the same operation order as
// the reference; the file is compiled with -Xcompiler -ffp-contract=off so x86 never fuses it.
#include <chrono>
#include <cmath>
#include <climits>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <nvjpeg.h>

#include "pmk_filter.cuh"

using namespace pmk;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CUDA_TRY(expr)                                                                                         \
    do {                                                                                                       \
        cudaError_t e__ = (expr);                                                                              \
        if (e__ != cudaSuccess) return fail(PMK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// ---- host copies of the reference's small linear algebra (left-to-right, one rounding per op) ----------
inline float hdot3(const float* a, const float* b) { float acc = a[0] * b[0]; acc = acc + a[1] * b[1]; acc = acc + a[2] * b[2]; return acc; }
inline float hdot4(const float* a, const float* b) { float acc = a[0] * b[0]; acc = acc + a[1] * b[1]; acc = acc + a[2] * b[2]; acc = acc + a[3] * b[3]; return acc; }
inline float hnorm3(const float* a) { return std::sqrt(hdot3(a, a)); }
inline void hcross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
// adjugate / determinant, cofactor expansion along the first row
inline void hinv3(const float* m, float* r) {
    const float c00 = m[4] * m[8] - m[5] * m[7];
    const float c01 = m[5] * m[6] - m[3] * m[8];
    const float c02 = m[3] * m[7] - m[4] * m[6];
    const float det = (m[0] * c00 + m[1] * c01) + m[2] * c02;
    r[0] = c00 / det;
    r[1] = (m[2] * m[7] - m[1] * m[8]) / det;
    r[2] = (m[1] * m[5] - m[2] * m[4]) / det;
    r[3] = c01 / det;
    r[4] = (m[0] * m[8] - m[2] * m[6]) / det;
    r[5] = (m[2] * m[3] - m[0] * m[5]) / det;
    r[6] = c02 / det;
    r[7] = (m[1] * m[6] - m[0] * m[7]) / det;
    r[8] = (m[0] * m[4] - m[1] * m[3]) / det;
}

// (int)floorf(log(ratio) / log(2.0f) + 0.5f) with the unqualified log resolving to double (optim.cpp:808)
inline int level_diff_host(float ratio) {
    const double v = std::log((double)ratio) / std::log((double)2.0f) + (double)0.5f;
    const float fv = (float)v;
    if (!(fv > -1.0e9f)) return INT_MIN;
    return (int)std::floor(fv);
}

// smallest float r with level_diff_host(r) >= k
float level_threshold(int k) {
    float r = std::exp2f((float)k - 0.5f);
    for (int guard = 0; guard < 64 && level_diff_host(r) >= k; ++guard) r = std::nextafterf(r, 0.0f);
    for (int guard = 0; guard < 128 && level_diff_host(r) < k; ++guard) r = std::nextafterf(r, INFINITY);
    return r;
}

struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct pmk_store;

struct pmk_ctx {
    pmk_config cfg;
    pmk_store* store = nullptr;             // device patch store (pmk_store_host.cuh), created on first use
    cudaStream_t stream = nullptr;
    cudaStream_t s_in = nullptr, s_out = nullptr;        // copy streams of the chunked host-buffer NCC call (s_in doubles as the sweep's second stream)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::vector<cudaEvent_t> ev_in, ev_k;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0;
    Params params;
    std::vector<ViewConst> h_views;          // working-level constants (device mirror in d_views)
    std::vector<std::vector<float> > P_all;  // per view: nlevels * 12
    std::vector<char> view_set;
    std::vector<std::vector<uint8_t*> > masks;   // per view: the mask pyramid (nlevels device arrays), empty = no mask
    std::vector<void*> owned;                // every device allocation of the context
    ViewConst* d_views = nullptr;
    unsigned int* d_counters = nullptr;      // work-queue heads
    bool views_dirty = true;
    Scratch s_coord, s_normal, s_views, s_nviews, s_incc, s_ncc, s_levels, s_misc[8];
    Scratch s_ready;                         // per-chunk arrival words of the streamed host-buffer NCC call
    unsigned int* h_epoch = nullptr;         // pinned source of those words
    size_t h_epoch_cap = 0;
    unsigned int epoch = 0;
    void* flush_buf = nullptr;
    size_t flush_bytes = 0;
    uint64_t launches = 0;
    unsigned k1_attr_done = 0;
    bool k1_prefetch = false;    // set by upload_views
    std::vector<Scratch> pool;               // staging buffers of the host-pointer entry points
    float* tex_scratch = nullptr;
    float* mat_scratch = nullptr;
    int cand_grid = 0;
    // PmMvps thresholds (pmmvps.cpp:54-67)
    float angle_threshold0, angle_threshold1, neighbor_threshold, neighbor_threshold1, neighbor_threshold2;
    float ncc_threshold, ncc_threshold_before;
    int depth = 0;
};

namespace {

int ensure(pmk_ctx* ctx, Scratch& s, size_t bytes) {
    if (bytes <= s.cap) return PMK_OK;
    if (s.p) CUDA_TRY(cudaFree(s.p));
    s.p = nullptr; s.cap = 0;
    const size_t cap = bytes + bytes / 4 + 256;
    CUDA_TRY(cudaMalloc(&s.p, cap));
    s.cap = cap;
    (void)ctx;
    return PMK_OK;
}

int upload_views(pmk_ctx* ctx) {
    if (!ctx->views_dirty) return PMK_OK;
    for (int v = 0; v < ctx->cfg.nviews; ++v)
        if (!ctx->view_set[v]) return fail(PMK_ERR_STATE, "pmk: view " + std::to_string(v) + " has not been uploaded (pmk_set_view)");
    CUDA_TRY(cudaMemcpyAsync(ctx->d_views, ctx->h_views.data(), sizeof(ViewConst) * ctx->cfg.nviews, cudaMemcpyHostToDevice, ctx->stream));
    // K1 prefetches its gathers through L2 when the levels it samples (the working level and coarser) cannot stay L2-resident
    size_t hot = 0;
    for (const ViewConst& vc : ctx->h_views)
        for (int l = ctx->cfg.level; l < ctx->cfg.level + 3 && l < PMK_MAX_LEVELS; ++l) hot += (size_t)vc.w[l] * vc.h[l] * sizeof(Texel);
    int l2_bytes = 0;
    cudaDeviceGetAttribute(&l2_bytes, cudaDevAttrL2CacheSize, ctx->cfg.device);
    const char* pf = getenv("PMK_K1_PREFETCH");
    ctx->k1_prefetch = pf ? atoi(pf) != 0 : hot > (size_t)l2_bytes * 3 / 4;
    ctx->views_dirty = false;
    return PMK_OK;
}

void refresh_params(pmk_ctx* ctx) {
    Params& p = ctx->params;
    p.views = ctx->d_views;
    p.nviews = ctx->cfg.nviews;
    p.level = ctx->cfg.level;
    p.nlevels = ctx->cfg.level + 3;
    p.csize = ctx->cfg.csize;
    p.wsize = ctx->cfg.wsize;
    p.min_image_num = ctx->cfg.min_image_num;
    p.tau = std::min(ctx->cfg.min_image_num * 2, ctx->cfg.nviews);        // pmmvps.cpp:32
    p.depth = ctx->depth;
    p.cos_angle1 = cosf(ctx->angle_threshold1);
    p.cos_angle0 = cosf(ctx->angle_threshold0);
    p.ncc_threshold = ctx->ncc_threshold;
    p.ncc_threshold_before = ctx->ncc_threshold_before;
    p.level_scale = (float)(1 << ctx->cfg.level);
    p.has_masks = 0;
    for (const auto& m : ctx->masks) if (!m.empty()) p.has_masks = 1;
    for (int k = 0; k < PMK_MAX_LEVELS; ++k) p.level_thr[k] = INFINITY;
    int slot = 0;
    for (int k = -ctx->cfg.level + 1; k <= 2 && slot < PMK_MAX_LEVELS; ++k) p.level_thr[slot++] = level_threshold(k);
}

template <int WS, int MINB>
int launch_k1(pmk_ctx* ctx, int n, const void* coord, const void* normal, const void* views, const void* nviews, int stride,
              void* incc, void* ncc, void* levels, const unsigned int* ready, unsigned int epoch, int chunk_shift, int packed) {
    const int fstride = ctx->params.tau * K1_FRAME_WORDS + 4;
    const size_t smem = (size_t)K1_WARPS * 32 * fstride * sizeof(float);
    if (!(ctx->k1_attr_done & (1u << MINB))) {                     // once per compiled variant
        CUDA_TRY(cudaFuncSetAttribute(k1_ncc<WS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        CUDA_TRY(cudaFuncSetAttribute(k1_ncc<WS, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        ctx->k1_attr_done |= 1u << MINB;
    }
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k1_ncc<WS, MINB>, K1_WARPS * 32, smem));
    if (per_sm < 1) return fail(PMK_ERR_CUDA, "pmk: k1_ncc does not fit on an SM");
    const int nbatch = (n + 31) / 32;
    const int want = (nbatch + K1_WARPS - 1) / K1_WARPS;
    const int grid = std::max(1, std::min(want, ctx->sm_count * per_sm));
    CUDA_TRY(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned int), ctx->stream));
    k1_ncc<WS, MINB><<<grid, K1_WARPS * 32, smem, ctx->stream>>>(ctx->params, n, (const float4*)coord, (const float4*)normal, (const int*)views,
                                                          (const int*)nviews, stride, (float*)incc, (float*)ncc, (int*)levels, ctx->d_counters,
                                                          ready, epoch, chunk_shift, (packed ? 1 : 0) | (ctx->k1_prefetch ? 2 : 0));
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return PMK_OK;
}

int dispatch_k1(pmk_ctx* ctx, int n, const void* coord, const void* normal, const void* views, const void* nviews, int stride,
                void* incc, void* ncc, void* levels, const unsigned int* ready = nullptr, unsigned int epoch = 0, int chunk_shift = 0, int packed = 0) {
    static const int minb = getenv("PMK_K1_MINB") ? atoi(getenv("PMK_K1_MINB")) : 4;   // tuning knob: CTAs/SM the kernel is compiled for (the 53 KB frame store caps it at 4)
#ifdef PMK_WS_ONLY
    if (ctx->cfg.wsize == PMK_WS_ONLY) return launch_k1<PMK_WS_ONLY, 4>(ctx, n, coord, normal, views, nviews, stride, incc, ncc, levels, ready, epoch, chunk_shift, packed);
    (void)minb;
#else
    switch (ctx->cfg.wsize) {
        case 5: return launch_k1<5, 3>(ctx, n, coord, normal, views, nviews, stride, incc, ncc, levels, ready, epoch, chunk_shift, packed);
        case 7: return minb == 3 ? launch_k1<7, 3>(ctx, n, coord, normal, views, nviews, stride, incc, ncc, levels, ready, epoch, chunk_shift, packed)
                                 : launch_k1<7, 4>(ctx, n, coord, normal, views, nviews, stride, incc, ncc, levels, ready, epoch, chunk_shift, packed);
        case 9: return launch_k1<9, 3>(ctx, n, coord, normal, views, nviews, stride, incc, ncc, levels, ready, epoch, chunk_shift, packed);
        case 11: return launch_k1<11, 2>(ctx, n, coord, normal, views, nviews, stride, incc, ncc, levels, ready, epoch, chunk_shift, packed);
    }
#endif
    return fail(PMK_ERR_ARG, "pmk: wsize must be 5, 7, 9 or 11");
}


// ---- candidate-kernel plumbing ------------------------------------------------------------------------------------
// order-preserving map between floats and unsigned ints, for bisection over float values
inline uint32_t f2o(float f) { uint32_t u; std::memcpy(&u, &f, 4); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
inline float o2f(uint32_t o) { uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o; float f; std::memcpy(&f, &u, 4); return f; }
// const float angle = acos(dot) with the unqualified (double) acos, narrowed (photoSet.cpp:92)
inline float acos_ref(float d) { return (float)std::acos((double)d); }

AngleGate make_gate(float min_angle, float max_angle) {
    AngleGate g;
    {   // smallest d in [-1, 1] with acos_ref(d) < max_angle (predicate is monotone: false ... true)
        uint32_t lo = f2o(-1.0f), hi = f2o(1.0f);
        if (acos_ref(-1.0f) < max_angle) g.dot_ge = -1.0f;
        else if (!(acos_ref(1.0f) < max_angle)) g.dot_ge = 2.0f;
        else { while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (acos_ref(o2f(mid)) < max_angle) hi = mid; else lo = mid; } g.dot_ge = o2f(hi); }
    }
    {   // largest d in [-1, 1] with min_angle < acos_ref(d) (true ... false)
        uint32_t lo = f2o(-1.0f), hi = f2o(1.0f);
        if (min_angle < acos_ref(1.0f)) g.dot_le = 1.0f;
        else if (!(min_angle < acos_ref(-1.0f))) g.dot_le = -2.0f;
        else { while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (min_angle < acos_ref(o2f(mid))) lo = mid; else hi = mid; } g.dot_le = o2f(lo); }
    }
    return g;
}

int cand_params(pmk_ctx* ctx, CandParams& cp, uint64_t seed) {
    if (ctx->cfg.nviews > CAND_MAXV) return fail(PMK_ERR_ARG, "pmk: the candidate kernels support at most 128 views");
    const int ws = ctx->cfg.wsize, texw = ws * ws * 3 + 4;
    if (!ctx->cand_grid) {
        ctx->cand_grid = ctx->sm_count * 4;
        const size_t warps = (size_t)ctx->cand_grid * CAND_WARPS;       // one slice per warp of the candidate kernels / per CTA of the sweep
        CUDA_TRY(cudaMalloc((void**)&ctx->tex_scratch, warps * ctx->cfg.nviews * texw * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&ctx->mat_scratch, warps * ctx->cfg.nviews * ctx->cfg.nviews * sizeof(float)));
        ctx->owned.push_back(ctx->tex_scratch);
        ctx->owned.push_back(ctx->mat_scratch);
    }
    cp.p = ctx->params;
    cp.gate = make_gate(ctx->cfg.max_angle_threshold, ctx->angle_threshold1);           // optim.cpp:153-156
    cp.sort_threshold = (float)(1.0f - std::cos(10.0f * M_PI / 180.0f));                // optim.cpp:222
    cp.ascale = (float)(M_PI / 48.0f);                                                   // optim.cpp:487
    cp.tex_scratch = ctx->tex_scratch;
    cp.mat_scratch = ctx->mat_scratch;
    cp.seed = seed;
    return PMK_OK;
}

// stage a host array into pool slot `slot`; returns the device pointer through *out
int stage_in(pmk_ctx* ctx, int slot, const void* host, size_t bytes, void** out) {
    if ((int)ctx->pool.size() <= slot) ctx->pool.resize(slot + 1);
    int rc = ensure(ctx, ctx->pool[slot], bytes ? bytes : 16);
    if (rc) return rc;
    if (host && bytes) CUDA_TRY(cudaMemcpyAsync(ctx->pool[slot].p, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *out = ctx->pool[slot].p;
    return PMK_OK;
}

template <typename K>
int cand_grid_for(pmk_ctx* ctx, int nwarps_needed) {
    (void)sizeof(K);
    return std::max(1, std::min(ctx->cand_grid, (nwarps_needed + CAND_WARPS - 1) / CAND_WARPS));
}

// PMK_WS_ONLY=<w> (dev builds: python -m mvskit_b200.build --fast) compiles the kernels for one window size only
#ifdef PMK_WS_ONLY
#define WS_DISPATCH(ws, CALL)                                 \
    switch (ws) {                                             \
        case PMK_WS_ONLY: { constexpr int WS = PMK_WS_ONLY; CALL; } break; \
        default: return fail(PMK_ERR_ARG, "pmk: this development build only has wsize " + std::to_string(PMK_WS_ONLY)); \
    }
#else
#define WS_DISPATCH(ws, CALL)                                 \
    switch (ws) {                                             \
        case 5: { constexpr int WS = 5; CALL; } break;        \
        case 7: { constexpr int WS = 7; CALL; } break;        \
        case 9: { constexpr int WS = 9; CALL; } break;        \
        case 11: { constexpr int WS = 11; CALL; } break;      \
        default: return fail(PMK_ERR_ARG, "pmk: wsize must be 5, 7, 9 or 11"); \
    }
#endif

}  // namespace

#include "pmk_store_host.cuh"

extern "C" {

int pmk_abi_version(void) { return PMK_ABI_VERSION; }
const char* pmk_last_error(void) { return g_err.c_str(); }

void pmk_default_config(pmk_config* c) {
    if (!c) return;
    std::memset(c, 0, sizeof(*c));
    c->device = 0;
    c->nviews = 0;
    c->level = 1; c->csize = 2; c->wsize = 7; c->min_image_num = 3;       // option.cpp:19-33
    c->ncc_threshold = 0.7f;
    c->max_angle_threshold = 10.0f * M_PI / 180.0f;
    c->quad_threshold = 2.5f;
    c->max_patches = 0;
    c->cell_capacity = 0;
    c->jitter_mode = 0;
    c->sweep_group = 1;
}

int pmk_create(const pmk_config* cfg, pmk_ctx** out) {
    if (!cfg || !out) return fail(PMK_ERR_ARG, "pmk_create: null argument");
    *out = nullptr;
    if (cfg->nviews < 1 || cfg->nviews > 4096) return fail(PMK_ERR_ARG, "pmk_create: nviews out of range");
    if (cfg->level < 0 || cfg->level + 3 > PMK_MAX_LEVELS) return fail(PMK_ERR_ARG, "pmk_create: level out of range");
    if (cfg->csize < 1) return fail(PMK_ERR_ARG, "pmk_create: csize < 1");
    if (cfg->wsize != 5 && cfg->wsize != 7 && cfg->wsize != 9 && cfg->wsize != 11) return fail(PMK_ERR_ARG, "pmk_create: wsize must be 5, 7, 9 or 11");
    if (cfg->min_image_num < 1 || std::min(cfg->min_image_num * 2, cfg->nviews) > PMK_MAX_TAU)
        return fail(PMK_ERR_ARG, "pmk_create: tau = min(2*minImageNum, nviews) exceeds PMK_MAX_TAU");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(PMK_ERR_CUDA, "pmk_create: no CUDA device (this library has no CPU path)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(PMK_ERR_ARG, "pmk_create: bad device ordinal");
    CUDA_TRY(cudaSetDevice(cfg->device));
    pmk_ctx* ctx = new pmk_ctx();
    ctx->cfg = *cfg;
    struct Guard { pmk_ctx* c; ~Guard() { if (c) pmk_destroy(c); } } guard{ctx};      // a failed step below releases what the earlier ones made
    CUDA_TRY(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, cfg->device));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreate(&ctx->ev0));
    CUDA_TRY(cudaEventCreate(&ctx->ev1));
    ctx->h_views.resize(cfg->nviews);
    std::memset(ctx->h_views.data(), 0, sizeof(ViewConst) * cfg->nviews);
    ctx->P_all.resize(cfg->nviews);
    ctx->view_set.assign(cfg->nviews, 0);
    ctx->masks.assign(cfg->nviews, std::vector<uint8_t*>());
    CUDA_TRY(cudaMalloc((void**)&ctx->d_views, sizeof(ViewConst) * cfg->nviews));
    CUDA_TRY(cudaMalloc((void**)&ctx->d_counters, 64 * sizeof(unsigned int)));
    // thresholds, pmmvps.cpp:54-67
    ctx->angle_threshold0 = 60.0f * M_PI / 180.0f;
    ctx->angle_threshold1 = 60.0f * M_PI / 180.0f;
    ctx->neighbor_threshold = 0.5f;
    ctx->neighbor_threshold1 = 1.0f;
    ctx->neighbor_threshold2 = 1.0f;
    ctx->ncc_threshold = cfg->ncc_threshold;
    ctx->ncc_threshold_before = cfg->ncc_threshold - 0.3f;
    ctx->depth = 0;
    refresh_params(ctx);
    guard.c = nullptr;
    *out = ctx;
    return PMK_OK;
}

void pmk_destroy(pmk_ctx* ctx) {
    if (!ctx) return;
    pmk_comm_destroy(ctx);
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (void* p : ctx->owned) cudaFree(p);
    Scratch* all[] = {&ctx->s_coord, &ctx->s_normal, &ctx->s_views, &ctx->s_nviews, &ctx->s_incc, &ctx->s_ncc, &ctx->s_levels, &ctx->s_ready};
    if (ctx->h_epoch) cudaFreeHost(ctx->h_epoch);
    for (Scratch* s : all) if (s->p) cudaFree(s->p);
    for (Scratch& s : ctx->s_misc) if (s.p) cudaFree(s.p);
    for (Scratch& s : ctx->pool) if (s.p) cudaFree(s.p);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    if (ctx->d_views) cudaFree(ctx->d_views);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    for (cudaEvent_t e : {ctx->ev0, ctx->ev1, ctx->ev_fork, ctx->ev_join}) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_in) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_k) cudaEventDestroy(e);
    for (cudaStream_t q : {ctx->s_in, ctx->s_out, ctx->stream}) if (q) cudaStreamDestroy(q);
    if (ctx->store && ctx->store->adj) cudaFree(ctx->store->adj);
    if (ctx->store && ctx->store->h_hdr) cudaFreeHost(ctx->store->h_hdr);
    delete ctx->store;
    delete ctx;
}

static int set_view_impl(pmk_ctx* ctx, int view, const float* P, const uint8_t* rgb, bool on_device, int width, int height) {
    if (!ctx || !P || (!rgb && !on_device)) return fail(PMK_ERR_ARG, "pmk_set_view: null argument");
    if (view < 0 || view >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_set_view: view out of range");
    const int nlevels = ctx->cfg.level + 3, level = ctx->cfg.level;
    if ((width >> (nlevels - 1)) < 8 || (height >> (nlevels - 1)) < 8) return fail(PMK_ERR_ARG, "pmk_set_view: image too small for the pyramid (the coarsest level must keep 8 pixels per side)");
    if (ctx->view_set[view]) return fail(PMK_ERR_STATE, "pmk_set_view: view already uploaded");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    ViewConst& vc = ctx->h_views[view];
    // ---- camera: updateProjection / updateCamera (camera.cpp:65-100), getCameraCenter (:295-308) ----
    std::vector<float>& Pl = ctx->P_all[view];
    Pl.assign((size_t)nlevels * 12, 0.0f);
    std::memcpy(Pl.data(), P, 12 * sizeof(float));
    for (int l = 1; l < nlevels; ++l) {
        float* q = Pl.data() + l * 12;
        std::memcpy(q, q - 12, 12 * sizeof(float));
        for (int c = 0; c < 8; ++c) q[c] = q[c] / 2.0f;
    }
    const float* P0 = Pl.data();
    std::memcpy(vc.P, Pl.data() + level * 12, 12 * sizeof(float));
    const float on = hnorm3(P0 + 8);
    for (int c = 0; c < 4; ++c) vc.oaxis[c] = P0[8 + c] / on;
    {
        const float M[9] = {P0[0], P0[1], P0[2], P0[4], P0[5], P0[6], P0[8], P0[9], P0[10]};
        float Mi[9];
        hinv3(M, Mi);
        const float q[3] = {P0[3], P0[7], P0[11]};
        for (int r = 0; r < 3; ++r) {
            float acc = (-Mi[3 * r]) * q[0];
            acc = acc + (-Mi[3 * r + 1]) * q[1];
            acc = acc + (-Mi[3 * r + 2]) * q[2];
            vc.center[r] = acc;
        }
        vc.center[3] = 1.0f;
    }
    {
        const float* Pw = vc.P;   // Camera::unproject inverts the working-level 3x3 (camera.cpp:331-335)
        const float M[9] = {Pw[0], Pw[1], Pw[2], Pw[4], Pw[5], Pw[6], Pw[8], Pw[9], Pw[10]};
        float Mi[9];
        hinv3(M, Mi);
        for (int r = 0; r < 3; ++r) { vc.Minv[4 * r] = Mi[3 * r]; vc.Minv[4 * r + 1] = Mi[3 * r + 1]; vc.Minv[4 * r + 2] = Mi[3 * r + 2]; vc.Minv[4 * r + 3] = 0.0f; }
    }
    // ---- Optim::setAxesScales (optim.cpp:43-65) ----
    {
        float za[3] = {vc.oaxis[0], vc.oaxis[1], vc.oaxis[2]}, x0[3] = {P0[0], P0[1], P0[2]}, ya[3], xa[3];
        hcross3(za, x0, ya);
        const float yn = hnorm3(ya);
        ya[0] = ya[0] / yn; ya[1] = ya[1] / yn; ya[2] = ya[2] / yn;
        hcross3(ya, za, xa);
        for (int i = 0; i < 3; ++i) { vc.xaxis[i] = xa[i]; vc.yaxis[i] = ya[i]; vc.zaxis[i] = za[i]; }
        vc.xaxis[3] = vc.yaxis[3] = vc.zaxis[3] = 0.0f;
        const float xa4[4] = {xa[0], xa[1], xa[2], 0.0f}, ya4[4] = {ya[0], ya[1], ya[2], 0.0f};
        vc.ipscale = hdot4(P0, xa4) + hdot4(P0 + 4, ya4);
    }
    // ---- image: dims (image.cpp:135-138), grid (patch_manager.cpp:36-37), pyramid (K0) ----
    vc.w[0] = width; vc.h[0] = height;
    for (int l = 1; l < nlevels; ++l) { vc.w[l] = vc.w[l - 1] / 2; vc.h[l] = vc.h[l - 1] / 2; }
    for (int l = nlevels; l < PMK_MAX_LEVELS; ++l) { vc.w[l] = 0; vc.h[l] = 0; vc.img[l] = nullptr; }
    vc.gw = (vc.w[level] + ctx->cfg.csize - 1) / ctx->cfg.csize;
    vc.gh = (vc.h[level] + ctx->cfg.csize - 1) / ctx->cfg.csize;
    const size_t npix0 = (size_t)width * height;
    if (!on_device) {
        { const int rc = ensure(ctx, ctx->s_misc[0], npix0 * 3); if (rc) return rc; }
        CUDA_TRY(cudaMemcpyAsync(ctx->s_misc[0].p, rgb, npix0 * 3, cudaMemcpyHostToDevice, ctx->stream));
    }
    for (int l = 0; l < nlevels; ++l) {
        void* d = nullptr;
        CUDA_TRY(cudaMalloc(&d, (size_t)vc.w[l] * vc.h[l] * sizeof(Texel)));
        ctx->owned.push_back(d);
        vc.img[l] = (const Texel*)d;
    }
    k0_u8_to_rgbx<<<(unsigned)(((npix0 + 3) / 4 + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t*)ctx->s_misc[0].p, (Texel*)vc.img[0], (int)npix0);
    ctx->launches++;
    for (int l = 1; l < nlevels; ++l) {
        dim3 blk(32, 8), grd((vc.w[l] + 31) / 32, (vc.h[l] + 7) / 8);
        k0_downsample<<<grd, blk, 0, ctx->stream>>>(vc.img[l - 1], vc.w[l - 1], vc.h[l - 1], (Texel*)vc.img[l], vc.w[l], vc.h[l]);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));       // rgb is a caller buffer; the staging copy must land
    ctx->view_set[view] = 1;
    ctx->views_dirty = true;
    return PMK_OK;
}

int pmk_set_view(pmk_ctx* ctx, int view, const float* P, const uint8_t* rgb, int width, int height) {
    if (!rgb) return fail(PMK_ERR_ARG, "pmk_set_view: null argument");
    return set_view_impl(ctx, view, P, rgb, false, width, height);
}

// ---- masks: Image::alloc's mask branch (image/image.cpp:143-161) + buildMaskPyramid (:717-747) -----------------------------------
int pmk_set_view_mask(pmk_ctx* ctx, int view, const uint8_t* grey, int width, int height) {
    if (!ctx || !grey) return fail(PMK_ERR_ARG, "pmk_set_view_mask: null argument");
    if (view < 0 || view >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_set_view_mask: view out of range");
    if (!ctx->view_set[view]) return fail(PMK_ERR_STATE, "pmk_set_view_mask: upload the view first (pmk_set_view)");
    if (!ctx->masks[view].empty()) return fail(PMK_ERR_STATE, "pmk_set_view_mask: mask already uploaded");
    ViewConst& vc = ctx->h_views[view];
    // the reference lets the mask's header overwrite m_widths[0] / m_heights[0] (image.cpp:146) and then indexes the image with them
    if (width != vc.w[0] || height != vc.h[0]) return fail(PMK_ERR_ARG, "pmk_set_view_mask: mask and image dimensions differ");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    const int nlevels = ctx->cfg.level + 3;
    const size_t npix0 = (size_t)width * height;
    { const int rc = ensure(ctx, ctx->s_misc[0], npix0); if (rc) return rc; }
    CUDA_TRY(cudaMemcpyAsync(ctx->s_misc[0].p, grey, npix0, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<uint8_t*> lv(nlevels, nullptr);
    for (int l = 0; l < nlevels; ++l) {
        void* d = nullptr;
        CUDA_TRY(cudaMalloc(&d, (size_t)vc.w[l] * vc.h[l]));
        ctx->owned.push_back(d);
        lv[l] = (uint8_t*)d;
    }
    k0_mask_threshold<<<(unsigned)((npix0 + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t*)ctx->s_misc[0].p, lv[0], (int)npix0);
    ctx->launches++;
    for (int l = 1; l < nlevels; ++l) {
        dim3 blk(32, 8), grd((vc.w[l] + 31) / 32, (vc.h[l] + 7) / 8);
        k0_mask_downsample<<<grd, blk, 0, ctx->stream>>>(lv[l - 1], vc.w[l - 1], vc.h[l - 1], lv[l], vc.w[l], vc.h[l]);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));       // grey is a caller buffer
    ctx->masks[view] = lv;
    vc.mask = lv[ctx->cfg.level];
    ctx->views_dirty = true;
    ctx->params.has_masks = 1;
    return PMK_OK;
}

int pmk_get_level_mask(pmk_ctx* ctx, int view, int level, uint8_t* mask_out, int* has_mask) {
    int w = 0, h = 0;
    const int rc = pmk_get_level_dims(ctx, view, level, &w, &h);
    if (rc) return rc;
    if (!has_mask) return fail(PMK_ERR_ARG, "pmk_get_level_mask: null argument");
    *has_mask = ctx->masks[view].empty() ? 0 : 1;
    if (!*has_mask || !mask_out) return PMK_OK;
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(mask_out, ctx->masks[view][level], (size_t)w * h, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_probe_mask(pmk_ctx* ctx, int n, int view, const float* coord4, int* out) {
    if (!ctx || !coord4 || !out) return fail(PMK_ERR_ARG, "pmk_probe_mask: null argument");
    if (view >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_probe_mask: view out of range");
    if (n <= 0) return PMK_OK;
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = upload_views(ctx);
    if (rc) return rc;
    const size_t N = (size_t)n;
    Scratch* s = ctx->s_misc;
    if ((rc = ensure(ctx, s[1], N * 16)) || (rc = ensure(ctx, s[4], N * 4))) return rc;
    CUDA_TRY(cudaMemcpyAsync(s[1].p, coord4, N * 16, cudaMemcpyHostToDevice, ctx->stream));
    k_probe_mask<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->params, n, view, (const float4*)s[1].p, (int*)s[4].p);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, s[4].p, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

// ---- JPEG ingest: nvJPEG decodes straight into the staging buffer K0 reads (Image::readJpeg, image/image.cpp:827-879) ----------------
namespace {
struct NvjpegApi {
    void* handle = nullptr;
    nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
    nvjpegStatus_t (*Destroy)(nvjpegHandle_t) = nullptr;
    nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
    nvjpegStatus_t (*JpegStateDestroy)(nvjpegJpegState_t) = nullptr;
    nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
    nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*, cudaStream_t) = nullptr;
};
NvjpegApi* nvjpeg_api() {
    static NvjpegApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so"};
        for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW); if (api.handle) break; }
        if (api.handle) {
            api.CreateSimple = (decltype(api.CreateSimple))dlsym(api.handle, "nvjpegCreateSimple");
            api.Destroy = (decltype(api.Destroy))dlsym(api.handle, "nvjpegDestroy");
            api.JpegStateCreate = (decltype(api.JpegStateCreate))dlsym(api.handle, "nvjpegJpegStateCreate");
            api.JpegStateDestroy = (decltype(api.JpegStateDestroy))dlsym(api.handle, "nvjpegJpegStateDestroy");
            api.GetImageInfo = (decltype(api.GetImageInfo))dlsym(api.handle, "nvjpegGetImageInfo");
            api.Decode = (decltype(api.Decode))dlsym(api.handle, "nvjpegDecode");
        }
    }
    return (api.handle && api.CreateSimple && api.Destroy && api.JpegStateCreate && api.JpegStateDestroy && api.GetImageInfo && api.Decode) ? &api : nullptr;
}
}  // namespace

int pmk_set_view_jpeg(pmk_ctx* ctx, int view, const float* P, const uint8_t* jpeg, uint64_t nbytes, int* width_out, int* height_out) {
    if (!ctx || !P || !jpeg || nbytes < 4) return fail(PMK_ERR_ARG, "pmk_set_view_jpeg: null argument");
    if (jpeg[0] != 0xFF || jpeg[1] != 0xD8) return fail(PMK_ERR_ARG, "pmk_set_view_jpeg: not a JPEG stream (no SOI marker)");
    NvjpegApi* api = nvjpeg_api();
    if (!api) return fail(PMK_ERR_STATE, "pmk_set_view_jpeg: libnvjpeg.so.12 not found");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    nvjpegHandle_t h = nullptr;
    nvjpegJpegState_t st = nullptr;
    if (api->CreateSimple(&h) != NVJPEG_STATUS_SUCCESS) return fail(PMK_ERR_CUDA, "nvjpegCreateSimple failed");
    int rc = PMK_OK, comps = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
    nvjpegChromaSubsampling_t sub;
    do {
        if (api->JpegStateCreate(h, &st) != NVJPEG_STATUS_SUCCESS) { rc = fail(PMK_ERR_CUDA, "nvjpegJpegStateCreate failed"); break; }
        if (api->GetImageInfo(h, jpeg, (size_t)nbytes, &comps, &sub, ws, hs) != NVJPEG_STATUS_SUCCESS) { rc = fail(PMK_ERR_ARG, "pmk_set_view_jpeg: cannot parse the JPEG header"); break; }
        const int w = ws[0], hgt = hs[0];
        if ((rc = ensure(ctx, ctx->s_misc[0], (size_t)w * hgt * 3))) break;
        nvjpegImage_t out;
        std::memset(&out, 0, sizeof(out));
        out.channel[0] = (unsigned char*)ctx->s_misc[0].p;
        out.pitch[0] = (size_t)w * 3;
        if (api->Decode(h, st, jpeg, (size_t)nbytes, NVJPEG_OUTPUT_RGBI, &out, ctx->stream) != NVJPEG_STATUS_SUCCESS) { rc = fail(PMK_ERR_CUDA, "nvjpegDecode failed"); break; }
        if (width_out) *width_out = w;
        if (height_out) *height_out = hgt;
        rc = set_view_impl(ctx, view, P, nullptr, true, w, hgt);      // synchronises the stream
    } while (false);
    if (st) api->JpegStateDestroy(st);
    api->Destroy(h);
    return rc;
}

int pmk_get_thresholds(pmk_ctx* ctx, pmk_thresholds* t) {
    if (!ctx || !t) return fail(PMK_ERR_ARG, "pmk_get_thresholds: null argument");
    t->tau = ctx->params.tau; t->depth = ctx->depth;
    t->ncc_threshold = ctx->ncc_threshold; t->ncc_threshold_before = ctx->ncc_threshold_before;
    t->angle_threshold0 = ctx->angle_threshold0; t->angle_threshold1 = ctx->angle_threshold1;
    t->max_angle_threshold = ctx->cfg.max_angle_threshold; t->quad_threshold = ctx->cfg.quad_threshold;
    t->neighbor_threshold = ctx->neighbor_threshold; t->neighbor_threshold1 = ctx->neighbor_threshold1; t->neighbor_threshold2 = ctx->neighbor_threshold2;
    return PMK_OK;
}

int pmk_set_depth(pmk_ctx* ctx, int depth) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_set_depth: null ctx");
    ctx->depth = depth;
    refresh_params(ctx);
    return PMK_OK;
}

int pmk_set_ncc_thresholds(pmk_ctx* ctx, float ncc, float before) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_set_ncc_thresholds: null ctx");
    ctx->ncc_threshold = ncc;
    ctx->ncc_threshold_before = before;
    refresh_params(ctx);
    return PMK_OK;
}

int pmk_update_threshold(pmk_ctx* ctx) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_update_threshold: null ctx");
    ctx->ncc_threshold -= 0.05f;              // pmmvps.cpp:70-74
    ctx->ncc_threshold_before -= 0.05f;
    ctx->depth += 1;                          // pmmvps.cpp:106
    refresh_params(ctx);
    return PMK_OK;
}

int pmk_get_camera(pmk_ctx* ctx, int view, int level, pmk_camera* out) {
    if (!ctx || !out) return fail(PMK_ERR_ARG, "pmk_get_camera: null argument");
    if (view < 0 || view >= ctx->cfg.nviews || !ctx->view_set[view]) return fail(PMK_ERR_ARG, "pmk_get_camera: view not set");
    if (level < 0 || level >= ctx->cfg.level + 3) return fail(PMK_ERR_ARG, "pmk_get_camera: level out of range");
    const ViewConst& vc = ctx->h_views[view];
    std::memcpy(out->P, ctx->P_all[view].data() + level * 12, 12 * sizeof(float));
    std::memcpy(out->center, vc.center, sizeof(out->center));
    std::memcpy(out->oaxis, vc.oaxis, sizeof(out->oaxis));
    for (int i = 0; i < 3; ++i) { out->xaxis[i] = vc.xaxis[i]; out->yaxis[i] = vc.yaxis[i]; out->zaxis[i] = vc.zaxis[i]; }
    out->ipscale = vc.ipscale;
    return PMK_OK;
}

int pmk_get_level_dims(pmk_ctx* ctx, int view, int level, int* width, int* height) {
    if (!ctx || !width || !height) return fail(PMK_ERR_ARG, "pmk_get_level_dims: null argument");
    if (view < 0 || view >= ctx->cfg.nviews || !ctx->view_set[view]) return fail(PMK_ERR_ARG, "pmk_get_level_dims: view not set");
    if (level < 0 || level >= ctx->cfg.level + 3) return fail(PMK_ERR_ARG, "pmk_get_level_dims: level out of range");
    *width = ctx->h_views[view].w[level]; *height = ctx->h_views[view].h[level];
    return PMK_OK;
}

int pmk_get_grid_dims(pmk_ctx* ctx, int view, int* gw, int* gh) {
    if (!ctx || !gw || !gh) return fail(PMK_ERR_ARG, "pmk_get_grid_dims: null argument");
    if (view < 0 || view >= ctx->cfg.nviews || !ctx->view_set[view]) return fail(PMK_ERR_ARG, "pmk_get_grid_dims: view not set");
    *gw = ctx->h_views[view].gw; *gh = ctx->h_views[view].gh;
    return PMK_OK;
}

int pmk_get_level_image(pmk_ctx* ctx, int view, int level, uint8_t* rgb_out) {
    int w = 0, h = 0;
    int rc = pmk_get_level_dims(ctx, view, level, &w, &h);
    if (rc) return rc;
    if (!rgb_out) return fail(PMK_ERR_ARG, "pmk_get_level_image: null output");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    const size_t npix = (size_t)w * h;
    if ((rc = ensure(ctx, ctx->s_misc[0], npix * 3))) return rc;
    k0_rgbx_to_u8<<<(unsigned)((npix + 255) / 256), 256, 0, ctx->stream>>>(ctx->h_views[view].img[level], (uint8_t*)ctx->s_misc[0].p, (int)npix);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(rgb_out, ctx->s_misc[0].p, npix * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_ncc_eval_dev(pmk_ctx* ctx, int n, const void* d_coord4, const void* d_normal4, const void* d_views, const void* d_nviews, int stride,
                     void* d_incc, void* d_ncc, void* d_levels) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_ncc_eval: null ctx");
    if (n < 0 || stride < 1) return fail(PMK_ERR_ARG, "pmk_ncc_eval: bad n/stride");
    if (n == 0) return PMK_OK;
    if (!d_coord4 || !d_normal4 || !d_views || !d_nviews || !d_incc) return fail(PMK_ERR_ARG, "pmk_ncc_eval: null buffer");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = upload_views(ctx);
    if (rc) return rc;
    return dispatch_k1(ctx, n, d_coord4, d_normal4, d_views, d_nviews, stride, d_incc, d_ncc, d_levels);
}

static int ncc_eval_host(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const int* views, const int* nviews, int stride,
                         float* incc_out, float* ncc_out, int* levels_out, int packed);

int pmk_ncc_eval(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const int* views, const int* nviews, int stride,
                 float* incc_out, float* ncc_out, int* levels_out) {
    return ncc_eval_host(ctx, n, coord4, normal4, views, nviews, stride, incc_out, ncc_out, levels_out, 0);
}

int pmk_ncc_eval_packed(pmk_ctx* ctx, int n, const float* coord3, const float* normal3, const uint8_t* views8, const uint8_t* nviews8, int stride,
                        float* incc_out, float* ncc_out, int* levels_out) {
    if (ctx && ctx->cfg.nviews > 255) return fail(PMK_ERR_ARG, "pmk_ncc_eval_packed: byte view ids need nviews <= 255 (255 marks an invalid entry)");
    return ncc_eval_host(ctx, n, coord3, normal3, (const int*)views8, (const int*)nviews8, stride, incc_out, ncc_out, levels_out, 1);
}

// element sizes of the four input arrays: float4 / float4 / int rows / int, or (packed) 3 floats / 3 floats / byte rows / byte
static int ncc_eval_host(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const int* views, const int* nviews, int stride,
                         float* incc_out, float* ncc_out, int* levels_out, int packed) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_ncc_eval: null ctx");
    if (n < 0 || stride < 1) return fail(PMK_ERR_ARG, "pmk_ncc_eval: bad n/stride");
    if (n == 0) return PMK_OK;
    if (!coord4 || !normal4 || !views || !nviews || !incc_out) return fail(PMK_ERR_ARG, "pmk_ncc_eval: null buffer");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    const size_t N = (size_t)n;
    const int tau = ctx->params.tau;
    int rc;
    const size_t ec = packed ? 12 : 16, ev = (packed ? 1 : 4) * (size_t)stride, en = packed ? 1 : 4;      // bytes per hypothesis: coord = normal, view row, nviews
    const char *h_c = (const char*)coord4, *h_n = (const char*)normal4, *h_v = (const char*)views, *h_nv = (const char*)nviews;
    if ((rc = ensure(ctx, ctx->s_coord, N * 16)) || (rc = ensure(ctx, ctx->s_normal, N * 16)) || (rc = ensure(ctx, ctx->s_views, N * stride * 4)) ||
        (rc = ensure(ctx, ctx->s_nviews, N * 4)) || (rc = ensure(ctx, ctx->s_incc, N * 4)) || (rc = ensure(ctx, ctx->s_ncc, N * 4)) ||
        (rc = ensure(ctx, ctx->s_levels, N * tau * 4)))
        return rc;
    cudaStream_t st = ctx->stream;
    static const char* mode_env = getenv("PMK_E2E_MODE");
    static const int mode_chunks = mode_env && !strcmp(mode_env, "chunks");
    static const int chunk_log2 = getenv("PMK_E2E_CHUNK_LOG2") ? std::min(24, std::max(14, atoi(getenv("PMK_E2E_CHUNK_LOG2")))) : 17;
    static const int nstreams = getenv("PMK_E2E_STREAMS") ? std::min(2, std::max(1, atoi(getenv("PMK_E2E_STREAMS")))) : 2;
    const size_t chunk = (size_t)1 << chunk_log2;
    const int nchunks = (int)((N + chunk - 1) / chunk);
    // an error return after the first enqueue must not leave copies in flight that read the caller's buffers / h_epoch
    struct Drain {
        pmk_ctx* c; bool ok;
        ~Drain() { if (!ok) { cudaStreamSynchronize(c->s_in); cudaStreamSynchronize(c->s_out); cudaStreamSynchronize(c->stream); } }
    } drain{ctx, false};
    CUDA_TRY(cudaEventRecord(ctx->ev1, st));                      // the copies must not overtake earlier work on the context stream
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_in, ctx->ev1, 0));
    CUDA_TRY(cudaStreamWaitEvent(ctx->s_out, ctx->ev1, 0));
    if (!mode_chunks || packed) {
        // Streamed: ONE K1 launch over the whole batch consumes the hypotheses while they cross PCIe.  The copy streams move the
        // inputs chunk by chunk and, after each chunk, a copy-engine write of this call's epoch into the arrival words of the chunk's
        // slots (one word per 2^14 hypotheses); a warp that draws a batch whose slot has not landed waits on that word (k1_ncc).
        // Every copy is enqueued BEFORE the kernel is launched (a failed enqueue returns before any launch), so the kernel never
        // depends on a later host action and tools that make launches synchronous cannot hang it.  Scores go straight into the caller's buffers when those are mapped pinned memory
        // (coalesced 128-byte posted writes), else through device staging and one D2H; the strided levels_out always takes staging.
        constexpr int SLOT_LOG2 = 14;
        const size_t slot = (size_t)1 << SLOT_LOG2, nslots = (N + slot - 1) / slot;
        {
            const void* before = ctx->s_ready.p;
            if ((rc = ensure(ctx, ctx->s_ready, nslots * 4))) return rc;
            // fresh device memory may hold anything (a freed array of small ints, say): no word may equal a future epoch
            if (ctx->s_ready.p != before) CUDA_TRY(cudaMemset(ctx->s_ready.p, 0, ctx->s_ready.cap));
        }
        if (ctx->h_epoch_cap < nslots) {
            if (ctx->h_epoch) CUDA_TRY(cudaFreeHost(ctx->h_epoch));
            ctx->h_epoch = nullptr; ctx->h_epoch_cap = 0;
            CUDA_TRY(cudaHostAlloc((void**)&ctx->h_epoch, (nslots + 64) * sizeof(unsigned int), cudaHostAllocDefault));
            ctx->h_epoch_cap = nslots + 64;
        }
        if (++ctx->epoch == 0) ctx->epoch = 1;                    // 0 is what cleared words hold
        const unsigned int epoch = ctx->epoch;
        for (size_t k = 0; k < nslots; ++k) ctx->h_epoch[k] = epoch;
        if ((rc = upload_views(ctx))) return rc;
        cudaStream_t sin[2] = {ctx->s_in, ctx->s_out};
        unsigned int* d_ready = (unsigned int*)ctx->s_ready.p;
        const size_t cslots = chunk >> SLOT_LOG2;
        int c = 0;
        for (size_t s0 = 0; s0 < nslots; s0 += cslots, ++c) {
            const size_t take = std::min(cslots, nslots - s0), o = s0 << SLOT_LOG2, m = std::min(take << SLOT_LOG2, N - o);
            cudaStream_t si = sin[c % nstreams];                            // two copy streams: one's set-up gaps hide under the other's transfer
            CUDA_TRY(cudaMemcpyAsync((char*)ctx->s_coord.p + o * ec, h_c + o * ec, m * ec, cudaMemcpyHostToDevice, si));
            CUDA_TRY(cudaMemcpyAsync((char*)ctx->s_normal.p + o * ec, h_n + o * ec, m * ec, cudaMemcpyHostToDevice, si));
            CUDA_TRY(cudaMemcpyAsync((char*)ctx->s_views.p + o * ev, h_v + o * ev, m * ev, cudaMemcpyHostToDevice, si));
            CUDA_TRY(cudaMemcpyAsync((char*)ctx->s_nviews.p + o * en, h_nv + o * en, m * en, cudaMemcpyHostToDevice, si));
            CUDA_TRY(cudaMemcpyAsync(d_ready + s0, ctx->h_epoch + s0, take * 4, cudaMemcpyHostToDevice, si));
        }
        void *d_incc = nullptr, *d_ncc = nullptr;
        bool direct = cudaHostGetDevicePointer(&d_incc, incc_out, 0) == cudaSuccess && (!ncc_out || cudaHostGetDevicePointer(&d_ncc, ncc_out, 0) == cudaSuccess);
        if (!direct) { (void)cudaGetLastError(); d_incc = ctx->s_incc.p; d_ncc = ncc_out ? ctx->s_ncc.p : nullptr; }
        rc = dispatch_k1(ctx, n, ctx->s_coord.p, ctx->s_normal.p, ctx->s_views.p, ctx->s_nviews.p, stride, d_incc, d_ncc,
                         levels_out ? ctx->s_levels.p : nullptr, (const unsigned int*)d_ready, epoch, SLOT_LOG2, packed);
        if (rc) return rc;
        if (!direct) {
            CUDA_TRY(cudaMemcpyAsync(incc_out, ctx->s_incc.p, N * 4, cudaMemcpyDeviceToHost, st));
            if (ncc_out) CUDA_TRY(cudaMemcpyAsync(ncc_out, ctx->s_ncc.p, N * 4, cudaMemcpyDeviceToHost, st));
        }
        if (levels_out) CUDA_TRY(cudaMemcpyAsync(levels_out, ctx->s_levels.p, N * tau * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaStreamSynchronize(ctx->s_in));
        CUDA_TRY(cudaStreamSynchronize(ctx->s_out));
        drain.ok = true;
        return PMK_OK;
    }
    // PMK_E2E_MODE=chunks (kept for A/B measurements): three-stage pipeline over chunks of the batch, H2D on s_in, one K1 launch per
    // chunk on the context stream, D2H on s_out, chained by events.
    while ((int)ctx->ev_in.size() < nchunks) {
        cudaEvent_t a, b;
        CUDA_TRY(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        ctx->ev_in.push_back(a); ctx->ev_k.push_back(b);
    }
    for (int c = 0; c < nchunks; ++c) {
        const size_t o = (size_t)c * chunk, m = std::min(chunk, N - o);
        CUDA_TRY(cudaMemcpyAsync((char*)ctx->s_coord.p + o * 16, coord4 + o * 4, m * 16, cudaMemcpyHostToDevice, ctx->s_in));
        CUDA_TRY(cudaMemcpyAsync((char*)ctx->s_normal.p + o * 16, normal4 + o * 4, m * 16, cudaMemcpyHostToDevice, ctx->s_in));
        CUDA_TRY(cudaMemcpyAsync((char*)ctx->s_views.p + o * stride * 4, views + o * stride, m * stride * 4, cudaMemcpyHostToDevice, ctx->s_in));
        CUDA_TRY(cudaMemcpyAsync((char*)ctx->s_nviews.p + o * 4, nviews + o, m * 4, cudaMemcpyHostToDevice, ctx->s_in));
        CUDA_TRY(cudaEventRecord(ctx->ev_in[c], ctx->s_in));
        CUDA_TRY(cudaStreamWaitEvent(st, ctx->ev_in[c], 0));
        rc = pmk_ncc_eval_dev(ctx, (int)m, (char*)ctx->s_coord.p + o * 16, (char*)ctx->s_normal.p + o * 16, (char*)ctx->s_views.p + o * stride * 4,
                              (char*)ctx->s_nviews.p + o * 4, stride, (char*)ctx->s_incc.p + o * 4, ncc_out ? (char*)ctx->s_ncc.p + o * 4 : nullptr,
                              levels_out ? (char*)ctx->s_levels.p + o * tau * 4 : nullptr);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(ctx->ev_k[c], st));
        CUDA_TRY(cudaStreamWaitEvent(ctx->s_out, ctx->ev_k[c], 0));
        CUDA_TRY(cudaMemcpyAsync(incc_out + o, (char*)ctx->s_incc.p + o * 4, m * 4, cudaMemcpyDeviceToHost, ctx->s_out));
        if (ncc_out) CUDA_TRY(cudaMemcpyAsync(ncc_out + o, (char*)ctx->s_ncc.p + o * 4, m * 4, cudaMemcpyDeviceToHost, ctx->s_out));
        if (levels_out) CUDA_TRY(cudaMemcpyAsync(levels_out + o * tau, (char*)ctx->s_levels.p + o * tau * 4, m * tau * 4, cudaMemcpyDeviceToHost, ctx->s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->s_out));
    CUDA_TRY(cudaStreamSynchronize(st));
    drain.ok = true;
    return PMK_OK;
}

int pmk_probe(pmk_ctx* ctx, int n, const int* view, const float* coord4, const float* normal4, float* project3, float* unit1,
              float* px4, float* py4, int* cell_ixy2, int* cell_ok) {
    if (!ctx || !view || !coord4) return fail(PMK_ERR_ARG, "pmk_probe: null argument");
    if (n <= 0) return PMK_OK;
    for (int i = 0; i < n; ++i) if (view[i] < 0 || view[i] >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_probe: view out of range");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = upload_views(ctx);
    if (rc) return rc;
    const size_t N = (size_t)n;
    Scratch* s = ctx->s_misc;
    if ((rc = ensure(ctx, s[0], N * 4)) || (rc = ensure(ctx, s[1], N * 16)) || (rc = ensure(ctx, s[2], N * 16)) || (rc = ensure(ctx, s[3], N * 12)) ||
        (rc = ensure(ctx, s[4], N * 4)) || (rc = ensure(ctx, s[5], N * 16)) || (rc = ensure(ctx, s[6], N * 16)) || (rc = ensure(ctx, s[7], N * 12)))
        return rc;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(s[0].p, view, N * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s[1].p, coord4, N * 16, cudaMemcpyHostToDevice, st));
    if (normal4) CUDA_TRY(cudaMemcpyAsync(s[2].p, normal4, N * 16, cudaMemcpyHostToDevice, st));
    int* cells = (int*)s[7].p;
    k_probe<<<(n + 127) / 128, 128, 0, st>>>(ctx->params, n, (const int*)s[0].p, (const float4*)s[1].p, normal4 ? (const float4*)s[2].p : nullptr,
                                             project3 ? (float*)s[3].p : nullptr, unit1 ? (float*)s[4].p : nullptr,
                                             (px4 && py4) ? (float4*)s[5].p : nullptr, (px4 && py4) ? (float4*)s[6].p : nullptr,
                                             cell_ixy2 ? cells : nullptr, cell_ok ? cells + 2 * N : nullptr);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    if (project3) CUDA_TRY(cudaMemcpyAsync(project3, s[3].p, N * 12, cudaMemcpyDeviceToHost, st));
    if (unit1) CUDA_TRY(cudaMemcpyAsync(unit1, s[4].p, N * 4, cudaMemcpyDeviceToHost, st));
    if (px4 && py4) {
        CUDA_TRY(cudaMemcpyAsync(px4, s[5].p, N * 16, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(py4, s[6].p, N * 16, cudaMemcpyDeviceToHost, st));
    }
    if (cell_ixy2) CUDA_TRY(cudaMemcpyAsync(cell_ixy2, cells, N * 8, cudaMemcpyDeviceToHost, st));
    if (cell_ixy2 && cell_ok) CUDA_TRY(cudaMemcpyAsync(cell_ok, cells + 2 * N, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PMK_OK;
}

// ---- candidate entry points (host pointers in, host pointers out) ----------------------------------------------------
int pmk_probe_unproject(pmk_ctx* ctx, int n, const int* view, const float* icoord3, float* coord4_out) {
    if (!ctx || !view || !icoord3 || !coord4_out) return fail(PMK_ERR_ARG, "pmk_probe_unproject: null argument");
    if (n <= 0) return PMK_OK;
    for (int i = 0; i < n; ++i) if (view[i] < 0 || view[i] >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_probe_unproject: view out of range");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = upload_views(ctx);
    if (rc) return rc;
    const size_t N = (size_t)n;
    void *dv, *di, *dout;
    if ((rc = stage_in(ctx, 0, view, N * 4, &dv)) || (rc = stage_in(ctx, 1, icoord3, N * 12, &di)) || (rc = stage_in(ctx, 2, nullptr, N * 16, &dout))) return rc;
    k_probe_unproject<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->params, n, (const int*)dv, (const float*)di, (float4*)dout);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(coord4_out, dout, N * 16, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_set_inccs(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const int* views, const int* nviews, int stride,
                  int robust, int pairwise, float* out) {
    if (!ctx || !coord4 || !normal4 || !views || !nviews || !out) return fail(PMK_ERR_ARG, "pmk_set_inccs: null argument");
    if (n <= 0) return PMK_OK;
    if (stride < 1 || stride > CAND_MAXV) return fail(PMK_ERR_ARG, "pmk_set_inccs: stride out of range");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = upload_views(ctx);
    if (rc) return rc;
    CandParams cp;
    if ((rc = cand_params(ctx, cp, 0))) return rc;
    const size_t N = (size_t)n, osz = N * stride * (pairwise ? stride : 1) * sizeof(float);
    void *dc, *dn, *dv, *dnv, *dout;
    if ((rc = stage_in(ctx, 0, coord4, N * 16, &dc)) || (rc = stage_in(ctx, 1, normal4, N * 16, &dn)) || (rc = stage_in(ctx, 2, views, N * stride * 4, &dv)) ||
        (rc = stage_in(ctx, 3, nviews, N * 4, &dnv)) || (rc = stage_in(ctx, 4, nullptr, osz, &dout)))
        return rc;
    CUDA_TRY(cudaMemsetAsync(dout, 0, osz, ctx->stream));
    const int grid = std::max(1, std::min(ctx->cand_grid, (n + CAND_WARPS - 1) / CAND_WARPS));
    const size_t smem = CAND_WARPS * sizeof(WarpScratch);
    WS_DISPATCH(ctx->cfg.wsize, (k2_set_inccs<WS><<<grid, CAND_WARPS * 32, smem, ctx->stream>>>(cp, n, (const float4*)dc, (const float4*)dn, (const int*)dv,
                                                                                            (const int*)dnv, stride, robust, pairwise, (float*)dout)));
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, dout, osz, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_pre_process(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const int* views, const int* nviews, int stride, int maxv,
                    int* ret, int* images_out, int* nimages_out, float* dscale_out, float* ascale_out) {
    if (!ctx || !coord4 || !normal4 || !views || !nviews || !ret || !images_out || !nimages_out || !dscale_out || !ascale_out)
        return fail(PMK_ERR_ARG, "pmk_pre_process: null argument");
    if (n <= 0) return PMK_OK;
    if (stride < 1 || maxv < 1) return fail(PMK_ERR_ARG, "pmk_pre_process: bad stride/maxv");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = upload_views(ctx);
    if (rc) return rc;
    CandParams cp;
    if ((rc = cand_params(ctx, cp, 0))) return rc;
    const size_t N = (size_t)n;
    void *dc, *dn, *dv, *dnv, *dret, *dimg, *dni, *dds, *das;
    if ((rc = stage_in(ctx, 0, coord4, N * 16, &dc)) || (rc = stage_in(ctx, 1, normal4, N * 16, &dn)) || (rc = stage_in(ctx, 2, views, N * stride * 4, &dv)) ||
        (rc = stage_in(ctx, 3, nviews, N * 4, &dnv)) || (rc = stage_in(ctx, 4, nullptr, N * 4, &dret)) || (rc = stage_in(ctx, 5, nullptr, N * maxv * 4, &dimg)) ||
        (rc = stage_in(ctx, 6, nullptr, N * 4, &dni)) || (rc = stage_in(ctx, 7, nullptr, N * 4, &dds)) || (rc = stage_in(ctx, 8, nullptr, N * 4, &das)))
        return rc;
    const int grid = std::max(1, std::min(ctx->cand_grid, (n + CAND_WARPS - 1) / CAND_WARPS));
    const size_t smem = CAND_WARPS * sizeof(WarpScratch);
    WS_DISPATCH(ctx->cfg.wsize, (k_pre_process<WS><<<grid, CAND_WARPS * 32, smem, ctx->stream>>>(cp, n, (const float4*)dc, (const float4*)dn, (const int*)dv,
                                 (const int*)dnv, stride, maxv, (int*)dret, (int*)dimg, (int*)dni, (float*)dds, (float*)das)));
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    cudaStream_t st = ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(ret, dret, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(images_out, dimg, N * maxv * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(nimages_out, dni, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(dscale_out, dds, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ascale_out, das, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PMK_OK;
}

int pmk_cost_func(pmk_ctx* ctx, int npatches, const float* coord4, const float* normal4, const float* dscale, const int* views, const int* nviews,
                  int stride, int nitems, const int* patch_of_item, const double* x3, double* cost_out) {
    if (!ctx || !coord4 || !normal4 || !dscale || !views || !nviews || !patch_of_item || !x3 || !cost_out) return fail(PMK_ERR_ARG, "pmk_cost_func: null argument");
    if (npatches <= 0 || nitems <= 0) return PMK_OK;
    for (int i = 0; i < nitems; ++i) if (patch_of_item[i] < 0 || patch_of_item[i] >= npatches) return fail(PMK_ERR_ARG, "pmk_cost_func: patch index out of range");
    for (int i = 0; i < npatches; ++i) if (nviews[i] < 1 || views[(size_t)i * stride] < 0 || views[(size_t)i * stride] >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_cost_func: bad reference view");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = upload_views(ctx);
    if (rc) return rc;
    CandParams cp;
    if ((rc = cand_params(ctx, cp, 0))) return rc;
    const size_t N = (size_t)npatches, M = (size_t)nitems;
    void *dc, *dn, *dd, *dv, *dnv, *dp, *dx, *dcost;
    if ((rc = stage_in(ctx, 0, coord4, N * 16, &dc)) || (rc = stage_in(ctx, 1, normal4, N * 16, &dn)) || (rc = stage_in(ctx, 2, views, N * stride * 4, &dv)) ||
        (rc = stage_in(ctx, 3, nviews, N * 4, &dnv)) || (rc = stage_in(ctx, 4, dscale, N * 4, &dd)) || (rc = stage_in(ctx, 5, patch_of_item, M * 4, &dp)) ||
        (rc = stage_in(ctx, 6, x3, M * 24, &dx)) || (rc = stage_in(ctx, 7, nullptr, M * 8, &dcost)))
        return rc;
    const int grid = std::max(1, std::min(ctx->cand_grid, (nitems + CAND_WARPS * 4 - 1) / (CAND_WARPS * 4)));
    WS_DISPATCH(ctx->cfg.wsize, (k_cost_func<WS><<<grid, CAND_WARPS * 32, 0, ctx->stream>>>(cp, nitems, (const int*)dp, (const float4*)dc, (const float4*)dn,
                                 (const float*)dd, (const int*)dv, (const int*)dnv, stride, (const double*)dx, (double*)dcost)));
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(cost_out, dcost, M * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_refine(pmk_ctx* ctx, int n, float* coord4, float* normal4, const float* dscale, const int* views, const int* nviews, int stride,
               const uint64_t* streams, uint64_t seed, float* ncc_out, double* trace_out) {
    if (!ctx || !coord4 || !normal4 || !dscale || !views || !nviews || !streams || !ncc_out) return fail(PMK_ERR_ARG, "pmk_refine: null argument");
    if (n <= 0) return PMK_OK;
    for (int i = 0; i < n; ++i) if (nviews[i] < 1 || views[(size_t)i * stride] < 0 || views[(size_t)i * stride] >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_refine: bad reference view");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = upload_views(ctx);
    if (rc) return rc;
    CandParams cp;
    if ((rc = cand_params(ctx, cp, seed))) return rc;
    const size_t N = (size_t)n;
    void *dc, *dn, *dd, *dv, *dnv, *dst, *dncc, *dtr = nullptr;
    if ((rc = stage_in(ctx, 0, coord4, N * 16, &dc)) || (rc = stage_in(ctx, 1, normal4, N * 16, &dn)) || (rc = stage_in(ctx, 2, views, N * stride * 4, &dv)) ||
        (rc = stage_in(ctx, 3, nviews, N * 4, &dnv)) || (rc = stage_in(ctx, 4, dscale, N * 4, &dd)) || (rc = stage_in(ctx, 5, streams, N * 8, &dst)) ||
        (rc = stage_in(ctx, 6, nullptr, N * 4, &dncc)))
        return rc;
    if (trace_out && (rc = stage_in(ctx, 7, nullptr, N * PMR1_EVALS * 32, &dtr))) return rc;
    const int grid = std::max(1, std::min(ctx->cand_grid, (n + CAND_WARPS - 1) / CAND_WARPS));
    const size_t smem = CAND_WARPS * sizeof(WarpScratch);
    WS_DISPATCH(ctx->cfg.wsize, (k3_refine<WS><<<grid, CAND_WARPS * 32, smem, ctx->stream>>>(cp, n, (float4*)dc, (float4*)dn, (const float*)dd, (const int*)dv,
                                 (const int*)dnv, stride, (const uint64_t*)dst, (float*)dncc, (double*)dtr)));
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    cudaStream_t st = ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(coord4, dc, N * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(normal4, dn, N * 16, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(ncc_out, dncc, N * 4, cudaMemcpyDeviceToHost, st));
    if (trace_out) CUDA_TRY(cudaMemcpyAsync(trace_out, dtr, N * PMR1_EVALS * 32, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PMK_OK;
}

int pmk_post_process(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const float* ncc, const int* views, const int* nviews, int stride,
                     int maxv, int* ret, int* images_out, int* nimages_out, int* grids_out, float* tmp_out) {
    if (!ctx || !coord4 || !normal4 || !ncc || !views || !nviews || !ret || !images_out || !nimages_out || !grids_out || !tmp_out)
        return fail(PMK_ERR_ARG, "pmk_post_process: null argument");
    if (n <= 0) return PMK_OK;
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = upload_views(ctx);
    if (rc) return rc;
    CandParams cp;
    if ((rc = cand_params(ctx, cp, 0))) return rc;
    const size_t N = (size_t)n;
    void *dc, *dn, *dq, *dv, *dnv, *dret, *dimg, *dni, *dgr, *dtmp;
    if ((rc = stage_in(ctx, 0, coord4, N * 16, &dc)) || (rc = stage_in(ctx, 1, normal4, N * 16, &dn)) || (rc = stage_in(ctx, 2, views, N * stride * 4, &dv)) ||
        (rc = stage_in(ctx, 3, nviews, N * 4, &dnv)) || (rc = stage_in(ctx, 4, ncc, N * 4, &dq)) || (rc = stage_in(ctx, 5, nullptr, N * 4, &dret)) ||
        (rc = stage_in(ctx, 6, nullptr, N * maxv * 4, &dimg)) || (rc = stage_in(ctx, 7, nullptr, N * 4, &dni)) || (rc = stage_in(ctx, 8, nullptr, N * maxv * 8, &dgr)) ||
        (rc = stage_in(ctx, 9, nullptr, N * 4, &dtmp)))
        return rc;
    const int grid = std::max(1, std::min(ctx->cand_grid, (n + CAND_WARPS - 1) / CAND_WARPS));
    const size_t smem = CAND_WARPS * sizeof(WarpScratch);
    WS_DISPATCH(ctx->cfg.wsize, (k_post_process<WS><<<grid, CAND_WARPS * 32, smem, ctx->stream>>>(cp, n, (const float4*)dc, (const float4*)dn, (const float*)dq,
                                 (const int*)dv, (const int*)dnv, stride, maxv, (int*)dret, (int*)dimg, (int*)dni, (int*)dgr, (float*)dtmp)));
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    cudaStream_t st = ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(ret, dret, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(images_out, dimg, N * maxv * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(nimages_out, dni, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(grids_out, dgr, N * maxv * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(tmp_out, dtmp, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PMK_OK;
}

int pmk_sync(pmk_ctx* ctx) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_sync: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_device_alloc(pmk_ctx* ctx, uint64_t bytes, void** out) {
    if (!ctx || !out) return fail(PMK_ERR_ARG, "pmk_device_alloc: null argument");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    CUDA_TRY(cudaMalloc(out, bytes ? bytes : 1));
    return PMK_OK;
}

int pmk_device_free(pmk_ctx* ctx, void* p) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_device_free: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    CUDA_TRY(cudaFree(p));
    return PMK_OK;
}

int pmk_memcpy_h2d(pmk_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_memcpy_h2d: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return PMK_OK;
}

int pmk_memcpy_d2h(pmk_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_memcpy_d2h: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PMK_OK;
}

int pmk_host_alloc_pinned(uint64_t bytes, void** out) {
    if (!out) return fail(PMK_ERR_ARG, "pmk_host_alloc_pinned: null argument");
    CUDA_TRY(cudaMallocHost(out, bytes ? bytes : 1));
    return PMK_OK;
}

int pmk_host_free_pinned(void* p) {
    CUDA_TRY(cudaFreeHost(p));
    return PMK_OK;
}

int pmk_timer_begin(pmk_ctx* ctx) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_timer_begin: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    CUDA_TRY(cudaEventRecord(ctx->ev0, ctx->stream));
    return PMK_OK;
}

int pmk_timer_end(pmk_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(PMK_ERR_ARG, "pmk_timer_end: null argument");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    CUDA_TRY(cudaEventRecord(ctx->ev1, ctx->stream));
    CUDA_TRY(cudaEventSynchronize(ctx->ev1));
    CUDA_TRY(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return PMK_OK;
}

int pmk_launch_count(pmk_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return fail(PMK_ERR_ARG, "pmk_launch_count: null argument");
    *out = ctx->launches;
    return PMK_OK;
}

int pmk_flush_l2(pmk_ctx* ctx) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_flush_l2: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    if (!ctx->flush_buf) {
        ctx->flush_bytes = (size_t)256 << 20;       // > 126 MB L2
        CUDA_TRY(cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
    }
    CUDA_TRY(cudaMemsetAsync(ctx->flush_buf, 0, ctx->flush_bytes, ctx->stream));
    return PMK_OK;
}

// ---- device patch store ---------------------------------------------------------------------------------------------------------
int pmk_store_clear(pmk_ctx* ctx) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_store_clear: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    const StoreDev& d = s->d;
    CUDA_TRY(cudaMemsetAsync(d.counters, 0, SC_COUNT * sizeof(int), ctx->stream));
    k_reset_cells<<<(d.total_cells + 255) / 256, 256, 0, ctx->stream>>>(d);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemsetAsync(d.state, 0, ((size_t)d.cap + d.stage_cap) * sizeof(int), ctx->stream));
    s->n = 0;
    s->canonical = true;
    return PMK_OK;
}

int pmk_store_add(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const float* scal4, const int* images, const int* nimages, int stride) {
    if (!ctx || !coord4 || !normal4 || !scal4 || !images || !nimages) return fail(PMK_ERR_ARG, "pmk_store_add: null argument");
    if (n <= 0) return PMK_OK;
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    StoreDev& d = s->d;
    if ((rc = store_check_overflow(ctx))) return rc;
    if (s->n + n > d.cap) return fail(PMK_ERR_CAPACITY, "pmk_store_add: patch store full (raise pmk_config.max_patches)");
    for (int i = 0; i < n; ++i) {
        if (nimages[i] < 1 || nimages[i] > std::min(stride, d.maxv)) return fail(PMK_ERR_ARG, "pmk_store_add: bad image count");
        for (int k = 0; k < nimages[i]; ++k) {
            const int v = images[(size_t)i * stride + k];
            if (v < 0 || v >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_store_add: image index out of range");
        }
    }
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    cudaStream_t st = ctx->stream;
    const int first = s->n;
    CUDA_TRY(cudaMemcpyAsync(d.coord + first, coord4, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d.normal + first, normal4, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d.scal + first, scal4, (size_t)n * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d.nimg + first, nimages, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpy2DAsync(d.images + (size_t)first * d.maxv, (size_t)d.maxv * 4, images, (size_t)stride * 4, (size_t)std::min(stride, d.maxv) * 4, n, cudaMemcpyHostToDevice, st));
    int birth0 = 0;
    CUDA_TRY(cudaMemcpyAsync(&birth0, d.counters + SC_BIRTH, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    k_store_add<<<std::max(1, std::min(ctx->sm_count * 4, (n + 3) / 4)), 128, 0, st>>>(sp, first, n, (unsigned int)birth0);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    s->n += n;
    const int c2[2] = {s->n, birth0 + n};
    CUDA_TRY(cudaMemcpyAsync(d.counters + SC_N, c2, sizeof(c2), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    s->canonical = false;
    return store_check_overflow(ctx);
}

int pmk_store_count(pmk_ctx* ctx, int* n_out) {
    if (!ctx || !n_out) return fail(PMK_ERR_ARG, "pmk_store_count: null argument");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    if ((rc = store_check_overflow(ctx))) return rc;
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    return store_order(ctx, sp, n_out);
}

int pmk_store_get(pmk_ctx* ctx, int nmax, int maxv, float* coord4, float* normal4, float* scal4, int* images, int* nimages, int* grids,
                  int* vimages, int* nvimages, int* vgrids, int* n_out) {
    if (!ctx || !n_out) return fail(PMK_ERR_ARG, "pmk_store_get: null argument");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    if ((rc = store_check_overflow(ctx))) return rc;
    pmk_store* s = ctx->store;
    const StoreDev& d = s->d;
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    int nalive = 0;
    if ((rc = store_order(ctx, sp, &nalive))) return rc;
    *n_out = nalive;
    const int n = std::min(nalive, nmax);
    if (n <= 0) return PMK_OK;
    cudaStream_t st = ctx->stream;
    auto fetch = [&](auto* src, void* dst, int row, int out_row, size_t elem) -> int {
        if (!dst) return PMK_OK;
        const long long total = (long long)n * row;
        typedef typename std::remove_pointer<decltype(src)>::type T;
        k_gather_rows<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, (T*)s->gather_tmp, s->vals2, total, row);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpy2DAsync(dst, (size_t)out_row * elem, s->gather_tmp, (size_t)row * elem, (size_t)std::min(row, out_row) * elem, n, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return PMK_OK;
    };
    if ((rc = fetch(d.coord, coord4, 1, 1, 16)) || (rc = fetch(d.normal, normal4, 1, 1, 16)) || (rc = fetch(d.scal, scal4, 1, 1, 16)) ||
        (rc = fetch(d.nimg, nimages, 1, 1, 4)) || (rc = fetch(d.nvimg, nvimages, 1, 1, 4)) ||
        (rc = fetch(d.images, images, d.maxv, maxv, 4)) || (rc = fetch(d.vimages, vimages, d.maxv, maxv, 4)))
        return rc;
    // grids come back as (ix, iy) pairs like Patch::m_grids
    std::vector<int> packed((size_t)n * maxv);
    for (int which = 0; which < 2; ++which) {
        int* out = which ? vgrids : grids;
        if (!out) continue;
        std::fill(packed.begin(), packed.end(), 0);
        if ((rc = fetch(which ? d.vcells : d.cells, packed.data(), d.maxv, maxv, 4))) return rc;
        for (size_t i = 0; i < packed.size(); ++i) { out[2 * i] = packed[i] & 0xffff; out[2 * i + 1] = (int)((unsigned)packed[i] >> 16); }
    }
    return PMK_OK;
}

int pmk_store_depth_map(pmk_ctx* ctx, int view, int* ids) {
    if (!ctx || !ids) return fail(PMK_ERR_ARG, "pmk_store_depth_map: null argument");
    if (view < 0 || view >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_store_depth_map: view out of range");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    const int c0 = s->cell_base[view], nc = s->cell_base[view + 1] - c0;
    std::vector<unsigned long long> keys(nc);
    CUDA_TRY(cudaMemcpyAsync(keys.data(), s->d.dmap + c0, (size_t)nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < nc; ++i) ids[i] = keys[i] == ~0ull ? -1 : (int)(keys[i] & 0xffffffffu);
    return PMK_OK;
}

int pmk_store_cell_counts(pmk_ctx* ctx, int view, int which, int* counts) {
    if (!ctx || !counts) return fail(PMK_ERR_ARG, "pmk_store_cell_counts: null argument");
    if (view < 0 || view >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_store_cell_counts: view out of range");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    const int c0 = s->cell_base[view], nc = s->cell_base[view + 1] - c0, cap = s->d.cell_cap;
    std::vector<int> cnt(nc), slots((size_t)nc * cap);
    CUDA_TRY(cudaMemcpyAsync(cnt.data(), s->d.ccount + c0, (size_t)nc * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(slots.data(), s->d.cslots + (size_t)c0 * cap, (size_t)nc * cap * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int c = 0; c < nc; ++c) {
        int k = 0;
        for (int i = 0; i < std::min(cnt[c], cap); ++i) {
            const int e = slots[(size_t)c * cap + i];
            if ((e & 0x7fffffff) == SLOT_TOMB) continue;
            if ((which != 0) == (e < 0)) ++k;
        }
        counts[c] = k;
    }
    return PMK_OK;
}

int pmk_store_colors(pmk_ctx* ctx, int nmax, uint8_t* rgb) {
    if (!ctx || !rgb) return fail(PMK_ERR_ARG, "pmk_store_colors: null argument");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    int nalive = 0;
    if ((rc = store_order(ctx, sp, &nalive))) return rc;
    const int n = std::min(nalive, nmax);
    if (n <= 0) return PMK_OK;
    k_patch_colors<<<(n + 127) / 128, 128, 0, ctx->stream>>>(sp, n, s->vals2, (unsigned char*)s->gather_tmp);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(rgb, s->gather_tmp, (size_t)n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

static int read_stats(pmk_ctx* ctx, uint64_t* stats16) {
    if (!stats16) return PMK_OK;
    CUDA_TRY(cudaMemcpyAsync(stats16, ctx->store->stats, SS_COUNT * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_propagate_diagonals(pmk_ctx* ctx, int iter, int image, int diag_first, int diag_count, uint64_t seed, uint64_t* stats16) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_propagate_diagonals: null ctx");
    if (image < 0 || image >= ctx->cfg.nviews || diag_first < 0 || diag_count < 0) return fail(PMK_ERR_ARG, "pmk_propagate_diagonals: bad argument");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(ctx->store->stats, 0, SS_COUNT * sizeof(uint64_t), ctx->stream));
    if ((rc = sweep_views(ctx, iter, image, 1, diag_first, diag_count, seed))) return rc;
    if ((rc = read_stats(ctx, stats16))) return rc;
    return store_check_overflow(ctx);
}

int pmk_propagate_forced(pmk_ctx* ctx, int iter, int image, int x, int y, const pmk_forced_io* io, uint64_t* stats16) {
    if (!ctx || !io) return fail(PMK_ERR_ARG, "pmk_propagate_forced: null argument");
    if (image < 0 || image >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_propagate_forced: image out of range");
    const int T = io->ntries, S = io->stride;
    if (T < 0 || T > 2 * SRC_MAX || S < ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_propagate_forced: bad ntries / stride (stride must hold every view)");
    if (!io->code || !io->ncc0 || !io->coord4 || !io->normal4 || !io->scal4 || !io->nimages || !io->images || !io->ntries_out || !io->outcome ||
        !io->branch_full || !io->post_ret || !io->nimages_out || !io->images_out || !io->grids_out || !io->nvimages_out || !io->vimages_out ||
        !io->vgrids_out || !io->tmp_out)
        return fail(PMK_ERR_ARG, "pmk_propagate_forced: null buffer");
    for (int t = 0; t < T; ++t) {
        if (io->code[t] < 0 || io->code[t] > 3) return fail(PMK_ERR_ARG, "pmk_propagate_forced: bad code");
        if (io->code[t] != 3) continue;
        if (io->nimages[t] < 1 || io->nimages[t] > S) return fail(PMK_ERR_ARG, "pmk_propagate_forced: bad image count");
        for (int k = 0; k < io->nimages[t]; ++k)
            if (io->images[(size_t)t * S + k] < 0 || io->images[(size_t)t * S + k] >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_propagate_forced: image index out of range");
    }
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    const ViewConst& vc = ctx->h_views[image];
    if (x < 0 || x >= vc.gw || y < 0 || y >= vc.gh) return fail(PMK_ERR_ARG, "pmk_propagate_forced: dest cell outside the grid");
    if (ctx->store->nranks > 1) return fail(PMK_ERR_STATE, "pmk_propagate_forced: single-GPU contexts only");
    const size_t Tn = (size_t)std::max(T, 1);
    void *d_code, *d_ncc0, *d_coord, *d_normal, *d_scal, *d_nimg, *d_images, *d_oi, *d_of, *d_ol;
    // o_i: ntries, outcome[T], full[T], post[T], nimg[T], nvimg[T]; o_l: images, cells, vimages, vcells [T][S] each
    if ((rc = stage_in(ctx, 0, io->code, Tn * 4, &d_code)) || (rc = stage_in(ctx, 1, io->ncc0, Tn * 4, &d_ncc0)) || (rc = stage_in(ctx, 2, io->coord4, Tn * 16, &d_coord)) ||
        (rc = stage_in(ctx, 3, io->normal4, Tn * 16, &d_normal)) || (rc = stage_in(ctx, 4, io->scal4, Tn * 16, &d_scal)) || (rc = stage_in(ctx, 5, io->nimages, Tn * 4, &d_nimg)) ||
        (rc = stage_in(ctx, 6, io->images, Tn * S * 4, &d_images)) || (rc = stage_in(ctx, 7, nullptr, (1 + 5 * Tn) * 4, &d_oi)) || (rc = stage_in(ctx, 8, nullptr, Tn * 4, &d_of)) ||
        (rc = stage_in(ctx, 9, nullptr, 4 * Tn * S * 4, &d_ol)))
        return rc;
    CUDA_TRY(cudaMemsetAsync(d_oi, 0xff, (1 + 5 * Tn) * 4, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d_of, 0, Tn * 4, ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d_ol, 0xff, 4 * Tn * S * 4, ctx->stream));
    ForceIO f;
    f.ntries = T; f.stride = S;
    f.code = (const int*)d_code; f.ncc0 = (const float*)d_ncc0; f.coord = (const float4*)d_coord; f.normal = (const float4*)d_normal; f.scal = (const float4*)d_scal;
    f.nimg = (const int*)d_nimg; f.images = (const int*)d_images;
    int* oi = (int*)d_oi;
    f.o_ntries = oi; f.o_outcome = oi + 1; f.o_full = oi + 1 + Tn; f.o_post = oi + 1 + 2 * Tn; f.o_nimg = oi + 1 + 3 * Tn; f.o_nvimg = oi + 1 + 4 * Tn;
    f.o_tmp = (float*)d_of;
    int* ol = (int*)d_ol;
    f.o_images = ol; f.o_cells = ol + Tn * S; f.o_vimages = ol + 2 * Tn * S; f.o_vcells = ol + 3 * Tn * S;
    CUDA_TRY(cudaMemsetAsync(ctx->store->stats, 0, SS_COUNT * sizeof(uint64_t), ctx->stream));
    const int inc = (iter % 2 == 1) ? -1 : 1, ndiag = vc.gw + vc.gh - 1, d = x + y;
    const int step = inc > 0 ? d : ndiag - 1 - d;
    if ((rc = sweep_views(ctx, iter, image, 1, step, 1, 0, x, &f))) return rc;
    std::vector<int> hi(1 + 5 * Tn), hl(4 * Tn * S);
    CUDA_TRY(cudaMemcpyAsync(hi.data(), d_oi, hi.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(hl.data(), d_ol, hl.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(io->tmp_out, d_of, Tn * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    *io->ntries_out = hi[0];
    for (int t = 0; t < T; ++t) {
        io->outcome[t] = hi[1 + t]; io->branch_full[t] = hi[1 + Tn + t]; io->post_ret[t] = hi[1 + 2 * Tn + t];
        io->nimages_out[t] = std::max(0, hi[1 + 3 * Tn + t]); io->nvimages_out[t] = std::max(0, hi[1 + 4 * Tn + t]);
        for (int k = 0; k < S; ++k) {
            const size_t o = (size_t)t * S + k;
            io->images_out[o] = hl[o]; io->vimages_out[o] = hl[2 * Tn * S + o];
            const int c = hl[Tn * S + o], vcell = hl[3 * Tn * S + o];
            io->grids_out[2 * o] = c == -1 ? -1 : (c & 0xffff); io->grids_out[2 * o + 1] = c == -1 ? -1 : (int)((unsigned)c >> 16);
            io->vgrids_out[2 * o] = vcell == -1 ? -1 : (vcell & 0xffff); io->vgrids_out[2 * o + 1] = vcell == -1 ? -1 : (int)((unsigned)vcell >> 16);
        }
    }
    if ((rc = read_stats(ctx, stats16))) return rc;
    return store_check_overflow(ctx);
}

int pmk_propagate(pmk_ctx* ctx, int iter, uint64_t seed, uint64_t* stats16) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_propagate: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    CUDA_TRY(cudaMemsetAsync(ctx->store->stats, 0, SS_COUNT * sizeof(uint64_t), ctx->stream));
    if (getenv("PMK_VERBOSE") && !ctx->store->laps) {
        if ((rc = dalloc(ctx, &ctx->store->laps, 17))) return rc;
    }
    if (ctx->store->laps) CUDA_TRY(cudaMemsetAsync(ctx->store->laps, 0, 17 * sizeof(unsigned long long), ctx->stream));
    ctx->store->host_wait_s = 0.0; ctx->store->host_waits = 0;
    const auto t_prop0 = std::chrono::steady_clock::now();
    const int group = ctx->store->group;
    for (int image = 0; image < ctx->cfg.nviews; image += group) {                        // propagate.cpp:73, `group` views at a time
        if ((rc = sweep_views(ctx, iter, image, std::min(group, ctx->cfg.nviews - image), 0, 1 << 30, seed))) return rc;
        if ((rc = store_check_overflow(ctx))) return rc;
    }
    if (ctx->store->laps) {
        unsigned long long l[16];
        CUDA_TRY(cudaMemcpyAsync(l, ctx->store->laps, sizeof(l), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_prop0).count();
        fprintf(stderr, "pmk_propagate[rank %d/%d] iter %d: wall %.3f s; stream laps (s): plan+sweep %.3f, apply %.3f, pack %.3f, header gather(+wait for ranks) %.3f, "
                        "host trip + payload gather %.3f, removals %.3f, keys+sort %.3f, scan+add %.3f; host blocked %.3f s in %lld syncs\n",
                ctx->store->rank, ctx->store->nranks, iter, wall, l[0] / 1e9, l[1] / 1e9, l[2] / 1e9, l[3] / 1e9, l[4] / 1e9, l[5] / 1e9, l[6] / 1e9, l[7] / 1e9,
                ctx->store->host_wait_s, ctx->store->host_waits);
    }
    return read_stats(ctx, stats16);
}

int pmk_filter_rebuild(pmk_ctx* ctx, int additive, int* n_out) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_filter_rebuild: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    if ((rc = store_rebuild(ctx, additive))) return rc;
    if (n_out) *n_out = ctx->store->n;
    return PMK_OK;
}

int pmk_filter_stage(pmk_ctx* ctx, int stage, int nmax, float* f_out, int* i_out, int* i_out2, int* killed_out) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_filter_stage: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    if (!s->canonical) return fail(PMK_ERR_STATE, "pmk_filter_stage: call pmk_filter_rebuild first");
    const int n = std::min(s->n, nmax);
    int killed = 0;
    if ((rc = filter_stage(ctx, stage, &killed))) return rc;
    if (killed_out) *killed_out = killed;
    if (n > 0) {
        if (f_out && (stage == 1 || stage == 3)) CUDA_TRY(cudaMemcpyAsync(f_out, s->f_tmp, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (i_out && stage != 1) CUDA_TRY(cudaMemcpyAsync(i_out, s->i_tmp, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (i_out2 && stage == 3) CUDA_TRY(cudaMemcpyAsync(i_out2, s->i_tmp3, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (i_out2 && stage == 2) CUDA_TRY(cudaMemcpyAsync(i_out2, s->d.nimg, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return store_check_overflow(ctx);
}

int pmk_filter(pmk_ctx* ctx, int* counts6) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_filter: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    int c[6] = {0, 0, 0, 0, 0, 0};
    static const bool verbose = getenv("PMK_VERBOSE") != nullptr;
    auto now = [&]() { cudaStreamSynchronize(ctx->stream); return std::chrono::steady_clock::now(); };
    auto t0 = verbose ? now() : std::chrono::steady_clock::time_point();
    auto lap = [&](const char* what) {
        if (!verbose) return;
        const auto t1 = now();
        fprintf(stderr, "pmk_filter: %-22s %8.2f ms (%d patches)\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count(), ctx->store->n);
        t0 = t1;
    };
    if ((rc = store_rebuild(ctx, 0))) return rc;                     // filter.cpp:26
    lap("rebuild(0)");
    c[0] = ctx->store->n;
    static const char* names[5] = {"", "filterOutside", "filterExact", "filterNeighbor", "filterSmallGroups"};
    for (int stage = 1; stage <= 4; ++stage) {                       // filterOutside, filterExact, filterNeighbor(1), filterSmallGroups
        if ((rc = filter_stage(ctx, stage, &c[stage]))) return rc;
        lap(names[stage]);
        if ((rc = store_rebuild(ctx, 1))) return rc;                 // filter.cpp:31,36,41,46
        lap("rebuild(1)");
    }
    c[5] = ctx->store->n;
    if (counts6) std::memcpy(counts6, c, sizeof(c));
    return PMK_OK;
}

// ---- multi-GPU ------------------------------------------------------------------------------------------------------------------
int pmk_step_share(int step_tasks, int rank, int nranks, int* count) {
    if (!count || step_tasks < 0 || nranks < 1 || rank < 0 || rank >= nranks) return fail(PMK_ERR_ARG, "pmk_step_share: bad argument");
    *count = step_tasks > rank ? (step_tasks - rank + nranks - 1) / nranks : 0;      // global tasks G of the step with G % nranks == rank
    return PMK_OK;
}

int pmk_comm_unique_id(char* id128) {
    if (!id128) return fail(PMK_ERR_ARG, "pmk_comm_unique_id: null argument");
    NcclApi* api = nccl_api();
    if (!api) return fail(PMK_ERR_STATE, "pmk_comm_unique_id: libnccl.so.2 not found");
    ncclUniqueId id;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    const ncclResult_t r = api->GetUniqueId(&id);
    if (r != ncclSuccess) return fail(PMK_ERR_CUDA, std::string("ncclGetUniqueId: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
    std::memcpy(id128, &id, 128);
    return PMK_OK;
}

int pmk_comm_init(pmk_ctx* ctx, int rank, int nranks, const char* id128) {
    if (!ctx || nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) return fail(PMK_ERR_ARG, "pmk_comm_init: bad argument (1 <= nranks <= 64)");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    if (s->nccl_comm) return fail(PMK_ERR_STATE, "pmk_comm_init: communicator already set");
    s->rank = rank; s->nranks = nranks;
    if (nranks == 1) return PMK_OK;
    if (!id128) return fail(PMK_ERR_ARG, "pmk_comm_init: null id");
    NcclApi* api = nccl_api();
    if (!api) return fail(PMK_ERR_STATE, "pmk_comm_init: libnccl.so.2 not found");
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    ncclComm_t comm = nullptr;
    const ncclResult_t r = api->CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return fail(PMK_ERR_CUDA, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
    s->nccl_comm = comm;
    // capacities of one rank's step message (the gather itself only moves what a step produced): a dest cell rarely stores more than
    // two patches per step and trims a handful; 2 x / 8 x the widest step, bounded so that (1 + nranks) buffers stay below ~10 GB
    s->ml.rec_cap = std::max(4096, std::min(1 << 18, s->max_tasks * 2));
    s->ml.rem_cap = std::max(16384, std::min(1 << 20, s->max_tasks * 8));
    s->ml.rec_words = 16 + 4 * s->d.maxv;
    if ((rc = dalloc(ctx, &s->msg, s->ml.words())) || (rc = dalloc(ctx, &s->all_msgs, s->ml.words() * nranks)) ||
        (rc = dalloc(ctx, &s->pack_ids, s->ml.rec_cap)) || (rc = dalloc(ctx, &s->rec_base, 2 * nranks + 4)) || (rc = dalloc(ctx, &s->all_hdr, 4 * nranks)))
        return rc;
    if (!s->h_hdr) CUDA_TRY(cudaMallocHost((void**)&s->h_hdr, 4 * 64 * sizeof(int)));
    s->ml.stride = s->ml.words();
    // NCCL sets its channels up on the first collective (hundreds of ms): pay that here, not inside the first wavefront step
    CUDA_TRY(cudaMemsetAsync(s->msg, 0, 4 * sizeof(int), ctx->stream));
    for (int warm = 0; warm < 2; ++warm) {
        const ncclResult_t wr = api->AllGather(s->msg, s->all_hdr, 4, ncclInt32, comm, ctx->stream);
        if (wr != ncclSuccess) return fail(PMK_ERR_CUDA, std::string("ncclAllGather (warm-up): ") + (api->GetErrorString ? api->GetErrorString(wr) : "error"));
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    const size_t nrec = (size_t)nranks * s->ml.rec_cap;
    if ((rc = dalloc(ctx, &s->mg_keys, nrec)) || (rc = dalloc(ctx, &s->mg_keys2, nrec)) || (rc = dalloc(ctx, &s->mg_vals, nrec)) || (rc = dalloc(ctx, &s->mg_vals2, nrec)))
        return rc;
    cub::DeviceRadixSort::SortPairs(nullptr, s->mg_cub_bytes, s->mg_keys, s->mg_keys2, s->mg_vals, s->mg_vals2, (int)nrec, 0, 64, ctx->stream);
    CUDA_TRY(cudaMalloc(&s->mg_cub, s->mg_cub_bytes));
    ctx->owned.push_back(s->mg_cub);
    return PMK_OK;
}

int pmk_comm_destroy(pmk_ctx* ctx) {
    if (!ctx || !ctx->store) return PMK_OK;
    pmk_store* s = ctx->store;
    if (s->nccl_comm) {
        CUDA_TRY(cudaSetDevice(ctx->cfg.device));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        NcclApi* api = nccl_api();
        if (api) api->CommDestroy((ncclComm_t)s->nccl_comm);
        s->nccl_comm = nullptr;
    }
    s->rank = 0; s->nranks = 1;
    return PMK_OK;
}

int pmk_store_checksum(pmk_ctx* ctx, uint64_t* out2) {
    if (!ctx || !out2) return fail(PMK_ERR_ARG, "pmk_store_checksum: null argument");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    if ((rc = store_check_overflow(ctx))) return rc;
    pmk_store* s = ctx->store;
    unsigned long long* d = (unsigned long long*)s->keys;
    CUDA_TRY(cudaMemsetAsync(d, 0, 16, ctx->stream));
    if (s->n > 0) {
        k_store_checksum<<<(s->n + 255) / 256, 256, 0, ctx->stream>>>(s->d, s->n, d);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    CUDA_TRY(cudaMemcpyAsync(out2, d, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_probe_neighbor(pmk_ctx* ctx, int n, const float* lhs10, const float* rhs10, const float* hunit, const float* radius, float threshold, int* out) {
    if (!ctx || !lhs10 || !rhs10 || !out) return fail(PMK_ERR_ARG, "pmk_probe_neighbor: null argument");
    if (radius && !hunit) return fail(PMK_ERR_ARG, "pmk_probe_neighbor: radius needs hunit");
    if (n <= 0) return PMK_OK;
    for (int i = 0; i < n; ++i) {
        const int lr = (int)lhs10[(size_t)i * 10 + 9], rr = (int)rhs10[(size_t)i * 10 + 9];
        if (lr < 0 || lr >= ctx->cfg.nviews || rr < 0 || rr >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_probe_neighbor: reference view out of range");
    }
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    const size_t N = (size_t)n;
    void *dl, *dr, *dh = nullptr, *dd = nullptr, *dout;
    if ((rc = stage_in(ctx, 0, lhs10, N * 40, &dl)) || (rc = stage_in(ctx, 1, rhs10, N * 40, &dr)) || (rc = stage_in(ctx, 2, nullptr, N * 4, &dout))) return rc;
    if (hunit && (rc = stage_in(ctx, 3, hunit, N * 4, &dh))) return rc;
    if (radius && (rc = stage_in(ctx, 4, radius, N * 4, &dd))) return rc;
    k_probe_neighbor<<<(n + 127) / 128, 128, 0, ctx->stream>>>(sp, n, (const float*)dl, (const float*)dr, (const float*)dh, (const float*)dd, threshold, (int*)dout);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, dout, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

// ---- pass-throughs for the rest of PatchManager's public surface (mvskit_b200/host) --------------------------------------------------
int pmk_contour2_to_projection(const float* intrinsics6, const float* extrinsics6, float* P12) {
    if (!intrinsics6 || !extrinsics6 || !P12) return fail(PMK_ERR_ARG, "pmk_contour2_to_projection: null argument");
    // Camera::setProjection, txtType 2 (camera.cpp:116-131) with Camera::quat2proj (:241-261); float arithmetic in the reference's
    // order, 4 x 4 products summed left to right (the order the oracle's Matrix4f shim defines)
    const float* in = intrinsics6; const float* q = extrinsics6;
    const float K[4][4] = {{in[0], in[2], in[3], 0.0f}, {0.0f, in[1], in[4], 0.0f}, {0.0f, 0.0f, 1.0f, 0.0f}, {0.0f, 0.0f, 0.0f, 1.0f}};
    const float a = (float)(q[0] * M_PI / 180.0), b = (float)(q[1] * M_PI / 180.0), g = (float)(q[2] * M_PI / 180.0);
    const float s1 = sinf(a), s2 = sinf(b), s3 = sinf(g), c1 = cosf(a), c2 = cosf(b), c3 = cosf(g);
    float R[4][4];
    R[0][0] = c2 * c3; R[0][1] = c3 * s2 * s1 - s3 * c1; R[0][2] = c3 * s2 * c1 + s3 * s1; R[0][3] = q[3];
    R[1][0] = s3 * c2; R[1][1] = s3 * s2 * s1 + c3 * c1; R[1][2] = s3 * s2 * c1 - c3 * s1; R[1][3] = q[4];
    R[2][0] = -s2;     R[2][1] = c2 * s1;                R[2][2] = c2 * c1;                R[2][3] = q[5];
    R[3][0] = R[3][1] = R[3][2] = 0.0f; R[3][3] = 1.0f;
    for (int y = 0; y < 3; ++y)
        for (int x = 0; x < 4; ++x) {
            float acc = K[y][0] * R[0][x];
            acc = acc + K[y][1] * R[1][x];
            acc = acc + K[y][2] * R[2][x];
            acc = acc + K[y][3] * R[3][x];
            P12[4 * y + x] = acc;
        }
    return PMK_OK;
}

int pmk_probe_visible(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const int* image, const int* cell_ixy, float strict, int* out, int* cell_ixy_out) {
    if (!ctx || !coord4 || !normal4 || !image || !out) return fail(PMK_ERR_ARG, "pmk_probe_visible: null argument");
    if (n <= 0) return PMK_OK;
    for (int i = 0; i < n; ++i) if (image[i] < 0 || image[i] >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_probe_visible: image out of range");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    const size_t N = (size_t)n;
    void *dc, *dn, *di, *dcell = nullptr, *dout, *dco;
    if ((rc = stage_in(ctx, 0, coord4, N * 16, &dc)) || (rc = stage_in(ctx, 1, normal4, N * 16, &dn)) || (rc = stage_in(ctx, 2, image, N * 4, &di)) ||
        (rc = stage_in(ctx, 3, nullptr, N * 4, &dout)) || (rc = stage_in(ctx, 4, nullptr, N * 8, &dco)))
        return rc;
    if (cell_ixy && (rc = stage_in(ctx, 5, cell_ixy, N * 8, &dcell))) return rc;
    k_probe_visible<<<(n + 127) / 128, 128, 0, ctx->stream>>>(sp, n, (const float4*)dc, (const float4*)dn, (const int*)di, (const int*)dcell, strict, (int*)dout, (int*)dco);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, dout, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (cell_ixy_out) CUDA_TRY(cudaMemcpyAsync(cell_ixy_out, dco, N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_probe_scales(pmk_ctx* ctx, int n, const float* coord4, const int* images, const int* nimages, int stride, float* dscale, float* ascale) {
    if (!ctx || !coord4 || !images || !nimages || !dscale || !ascale) return fail(PMK_ERR_ARG, "pmk_probe_scales: null argument");
    if (n <= 0) return PMK_OK;
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < std::min(nimages[i], stride); ++k)
            if (images[(size_t)i * stride + k] < 0 || images[(size_t)i * stride + k] >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_probe_scales: image out of range");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = upload_views(ctx);
    if (rc) return rc;
    CandParams cp;
    if ((rc = cand_params(ctx, cp, 0))) return rc;
    const size_t N = (size_t)n;
    void *dc, *di, *dni, *dd, *da;
    if ((rc = stage_in(ctx, 0, coord4, N * 16, &dc)) || (rc = stage_in(ctx, 1, images, N * stride * 4, &di)) || (rc = stage_in(ctx, 2, nimages, N * 4, &dni)) ||
        (rc = stage_in(ctx, 3, nullptr, N * 4, &dd)) || (rc = stage_in(ctx, 4, nullptr, N * 4, &da)))
        return rc;
    k_probe_scales<<<std::max(1, std::min(ctx->cand_grid, (n + CAND_WARPS - 1) / CAND_WARPS)), CAND_WARPS * 32, 0, ctx->stream>>>(cp, n, (const float4*)dc, (const int*)di, (const int*)dni, stride, (float*)dd, (float*)da);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(dscale, dd, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(ascale, da, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_probe_neighbors(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const float* scal4, const int* images, const int* nimages, int stride,
                        float scale, int margin, int cap, int* ids_out, int* count_out) {
    if (!ctx || !coord4 || !normal4 || !scal4 || !images || !nimages || !ids_out || !count_out) return fail(PMK_ERR_ARG, "pmk_probe_neighbors: null argument");
    if (n <= 0) return PMK_OK;
    if (margin < 0 || margin > 2 || cap < 1) return fail(PMK_ERR_ARG, "pmk_probe_neighbors: margin must be 0..2 and cap >= 1");
    for (int i = 0; i < n; ++i) {
        if (nimages[i] < 1 || nimages[i] > stride) return fail(PMK_ERR_ARG, "pmk_probe_neighbors: bad image count");
        for (int k = 0; k < nimages[i]; ++k) if (images[(size_t)i * stride + k] < 0 || images[(size_t)i * stride + k] >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_probe_neighbors: image out of range");
    }
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    const size_t N = (size_t)n;
    void *dc, *dn, *ds, *di, *dni, *dids, *dcnt;
    if ((rc = stage_in(ctx, 0, coord4, N * 16, &dc)) || (rc = stage_in(ctx, 1, normal4, N * 16, &dn)) || (rc = stage_in(ctx, 2, scal4, N * 16, &ds)) ||
        (rc = stage_in(ctx, 3, images, N * stride * 4, &di)) || (rc = stage_in(ctx, 4, nimages, N * 4, &dni)) || (rc = stage_in(ctx, 5, nullptr, N * cap * 4, &dids)) ||
        (rc = stage_in(ctx, 6, nullptr, N * 4, &dcnt)))
        return rc;
    const int grid = std::max(1, std::min(ctx->cand_grid, (n + CAND_WARPS - 1) / CAND_WARPS));
    k_probe_neighbors<<<grid, CAND_WARPS * 32, CAND_WARPS * 2 * CAND_MAXV * sizeof(int), ctx->stream>>>(sp, n, (const float4*)dc, (const float4*)dn, (const float4*)ds, (const int*)di,
                                                                                                     (const int*)dni, stride, scale, margin, cap, (int*)dids, (int*)dcnt);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(ids_out, dids, N * cap * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(count_out, dcnt, N * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return store_check_overflow(ctx);
}

int pmk_store_ids(pmk_ctx* ctx, int nmax, int* ids_out, int* n_out) {
    if (!ctx || !n_out || (nmax > 0 && !ids_out)) return fail(PMK_ERR_ARG, "pmk_store_ids: null argument");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    if ((rc = store_check_overflow(ctx))) return rc;
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    int nalive = 0;
    if ((rc = store_order(ctx, sp, &nalive))) return rc;         // live patches first, in collect order: vals2 = their store ids
    *n_out = nalive;
    const int n = std::min(nalive, nmax);
    if (n > 0) {
        CUDA_TRY(cudaMemcpyAsync(ids_out, ctx->store->vals2, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    }
    return PMK_OK;
}

int pmk_store_remove(pmk_ctx* ctx, int n, const int* ids) {
    if (!ctx || (n > 0 && !ids)) return fail(PMK_ERR_ARG, "pmk_store_remove: null argument");
    if (n <= 0) return PMK_OK;
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    if ((rc = store_check_overflow(ctx))) return rc;
    for (int i = 0; i < n; ++i) if (ids[i] < 0 || ids[i] >= s->n) return fail(PMK_ERR_ARG, "pmk_store_remove: id out of range");
    if (n > s->d.cap) return fail(PMK_ERR_ARG, "pmk_store_remove: too many ids");
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    CUDA_TRY(cudaMemcpyAsync(s->rem_list, ids, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(cudaMemcpyAsync(s->d.counters + SC_REM, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    k4_apply_remove<<<std::max(1, std::min(ctx->sm_count, (n + 3) / 4)), 128, 0, ctx->stream>>>(sp, s->rem_list, s->d.cap);   // PatchManager::removePatch (patch_manager.cpp:303-325)
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    s->canonical = false;
    return PMK_OK;
}

int pmk_store_update_depth_maps(pmk_ctx* ctx, int n, const int* ids) {
    if (!ctx || (n > 0 && !ids)) return fail(PMK_ERR_ARG, "pmk_store_update_depth_maps: null argument");
    if (n <= 0) return PMK_OK;
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    void* dids;
    if ((rc = stage_in(ctx, 0, ids, (size_t)n * 4, &dids))) return rc;
    k_store_update_depth<<<std::max(1, std::min(ctx->sm_count * 4, (n + 3) / 4)), 128, 0, ctx->stream>>>(sp, n, (const int*)dids);   // patch_manager.cpp:191-221
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return PMK_OK;
}

int pmk_store_cell_ids(pmk_ctx* ctx, int view, int which, int* offsets, int* ids, int ids_cap, int* total_out) {
    if (!ctx || !offsets || !total_out) return fail(PMK_ERR_ARG, "pmk_store_cell_ids: null argument");
    if (view < 0 || view >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_store_cell_ids: view out of range");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    const int c0 = s->cell_base[view], nc = s->cell_base[view + 1] - c0;
    if (nc + 1 > s->d.cap) return fail(PMK_ERR_CAPACITY, "pmk_store_cell_ids: scratch too small for this grid");
    cudaStream_t st = ctx->stream;
    k_store_cell_ids<<<(nc + 255) / 256, 256, 0, st>>>(s->d, c0, nc, which, 0, s->i_tmp, nullptr, nullptr);
    CUDA_TRY(cudaMemsetAsync(s->i_tmp + nc, 0, sizeof(int), st));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(s->cub_tmp, s->cub_bytes, s->i_tmp, s->i_tmp2, nc + 1, st));
    CUDA_TRY(cudaMemcpyAsync(offsets, s->i_tmp2, (size_t)(nc + 1) * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const int total = offsets[nc];
    *total_out = total;
    ctx->launches += 2;
    if (!ids || total <= 0) return PMK_OK;
    if (total > ids_cap) return fail(PMK_ERR_CAPACITY, "pmk_store_cell_ids: ids buffer too small (see total_out)");
    if ((size_t)total * 4 > s->gather_bytes) return fail(PMK_ERR_CAPACITY, "pmk_store_cell_ids: scratch too small");
    k_store_cell_ids<<<(nc + 255) / 256, 256, 0, st>>>(s->d, c0, nc, which, 1, nullptr, s->i_tmp2, (int*)s->gather_tmp);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(ids, s->gather_tmp, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return PMK_OK;
}

int pmk_debug_cell_times(pmk_ctx* ctx, float* out_total_cells) {
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_debug_cell_times: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    const size_t n = (size_t)s->d.total_cells;
    if (!s->cell_ns) {
        if ((rc = dalloc(ctx, &s->cell_ns, n))) return rc;
        CUDA_TRY(cudaMemsetAsync(s->cell_ns, 0, n * sizeof(float), ctx->stream));
    }
    if (out_total_cells) {
        CUDA_TRY(cudaMemcpyAsync(out_total_cells, s->cell_ns, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaMemsetAsync(s->cell_ns, 0, n * sizeof(float), ctx->stream));
    }
    return PMK_OK;
}

int pmk_debug_phase_times(pmk_ctx* ctx, uint64_t* out8) {       // out8: 16 words (8 try phases, 8 sub-phases of PMK_SUBPHASE builds)
    if (!ctx) return fail(PMK_ERR_ARG, "pmk_debug_phase_times: null ctx");
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    if (!s->phase_ns) {
        if ((rc = dalloc(ctx, &s->phase_ns, 16))) return rc;
        CUDA_TRY(cudaMemsetAsync(s->phase_ns, 0, 16 * sizeof(unsigned long long), ctx->stream));
    }
    if (out8) {
        CUDA_TRY(cudaMemcpyAsync(out8, s->phase_ns, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(cudaMemsetAsync(s->phase_ns, 0, 16 * sizeof(unsigned long long), ctx->stream));
    }
    return PMK_OK;
}

int pmk_probe_check(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const float* scal4, const int* images, const int* nimages, int stride,
                    int* ret, float* gain, int* nneighbors, int* vimages_out, int* nvimages_out) {
    if (!ctx || !coord4 || !normal4 || !scal4 || !images || !nimages || !ret || !gain || !nneighbors || !vimages_out || !nvimages_out)
        return fail(PMK_ERR_ARG, "pmk_probe_check: null argument");
    if (n <= 0) return PMK_OK;
    if (stride < ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_probe_check: stride must hold every view (vimages_out)");
    for (int i = 0; i < n; ++i) {
        if (nimages[i] < 1 || nimages[i] > stride) return fail(PMK_ERR_ARG, "pmk_probe_check: bad image count");
        for (int k = 0; k < nimages[i]; ++k) if (images[(size_t)i * stride + k] < 0 || images[(size_t)i * stride + k] >= ctx->cfg.nviews) return fail(PMK_ERR_ARG, "pmk_probe_check: image index out of range");
    }
    CUDA_TRY(cudaSetDevice(ctx->cfg.device));
    int rc = store_init(ctx);
    if (rc) return rc;
    StoreParams sp;
    if ((rc = store_params(ctx, sp, 0))) return rc;
    const size_t N = (size_t)n;
    void *dc, *dn, *ds, *di, *dni, *dret, *dg, *dnn, *dvi, *dnv;
    if ((rc = stage_in(ctx, 0, coord4, N * 16, &dc)) || (rc = stage_in(ctx, 1, normal4, N * 16, &dn)) || (rc = stage_in(ctx, 2, scal4, N * 16, &ds)) ||
        (rc = stage_in(ctx, 3, images, N * stride * 4, &di)) || (rc = stage_in(ctx, 4, nimages, N * 4, &dni)) || (rc = stage_in(ctx, 5, nullptr, N * 4, &dret)) ||
        (rc = stage_in(ctx, 6, nullptr, N * 4, &dg)) || (rc = stage_in(ctx, 7, nullptr, N * 4, &dnn)) || (rc = stage_in(ctx, 8, nullptr, N * stride * 4, &dvi)) ||
        (rc = stage_in(ctx, 9, nullptr, N * 4, &dnv)))
        return rc;
    const int grid = std::max(1, std::min(ctx->cand_grid, (n + CAND_WARPS - 1) / CAND_WARPS));
    const size_t smem = CAND_WARPS * (sizeof(WarpScratch) + 3 * CAND_MAXV * sizeof(int));
    k_probe_check<<<grid, CAND_WARPS * 32, smem, ctx->stream>>>(sp, n, (const float4*)dc, (const float4*)dn, (const float4*)ds, (const int*)di, (const int*)dni, stride,
                                                               (int*)dret, (float*)dg, (int*)dnn, (int*)dvi, (int*)dnv);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    cudaStream_t st = ctx->stream;
    CUDA_TRY(cudaMemcpyAsync(ret, dret, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(gain, dg, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(nneighbors, dnn, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(vimages_out, dvi, N * stride * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(nvimages_out, dnv, N * 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return store_check_overflow(ctx);
}

}  // extern "C"
