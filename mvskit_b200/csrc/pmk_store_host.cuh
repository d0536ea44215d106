// mvskit_b200/csrc/pmk_store_host.cuh -- host side of the device patch store: allocation, the rebuild
// (collect + compact + register + depth maps + visible lists), the sweep driver (Propagate::run) and Filter::run.
// Included by pmk_api.cu after pmk_ctx is defined.
#pragma once

#include <algorithm>
#include <random>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <dlfcn.h>
#include <nccl.h>

#include "pmk_filter.cuh"

struct pmk_store {
    pmk::StoreDev d;                    // device pointers
    int n = 0;                          // patches allocated (host mirror of SC_N)
    int max_tasks = 0;
    int rank = 0, nranks = 1;             // multi-GPU: this context sweeps every nranks-th dest cell of a step, starting at `rank`
    int group = 1;                        // views swept concurrently (pmk_config.sweep_group)
    int* cell_base_d = nullptr;
    std::vector<int> cell_base;         // host copy
    // scratch
    int* rem_list = nullptr; int* task_new = nullptr; int* final_id = nullptr; int* order = nullptr; int* alive_list = nullptr;
    unsigned long long* stats = nullptr; unsigned long long* step_max = nullptr;
    float* cell_ns = nullptr;            // per-cell sweep time of the last pass (allocated on first pmk_debug_cell_times call)
    unsigned long long* phase_ns = nullptr;   // warp time by try phase (allocated on first pmk_debug_phase_times call)
    unsigned long long* keys = nullptr; unsigned long long* keys2 = nullptr;
    int* vals = nullptr; int* vals2 = nullptr;
    void* cub_tmp = nullptr; size_t cub_bytes = 0;
    void* gather_tmp = nullptr; size_t gather_bytes = 0;
    float* f_tmp = nullptr; int* i_tmp = nullptr; int* i_tmp2 = nullptr; int* i_tmp3 = nullptr;
    int* nb_scratch = nullptr;
    int* adj = nullptr; size_t adj_cap = 0;   // filterSmallGroups: directed neighbour lists (CSR values), grown on demand
    int* small = nullptr;               // a few device ints for results
    float jitter[4];
    // multi-GPU exchange of the step mutations (pmk_comm_init)
    void* nccl_comm = nullptr;
    pmk::MsgLayout ml = {0, 0, 0, 0};
    int* msg = nullptr; int* all_msgs = nullptr; int* pack_ids = nullptr; int* rec_base = nullptr;
    unsigned long long* laps = nullptr;      // PMK_VERBOSE: [16] lap accumulators + [16] = last stamp
    double host_wait_s = 0.0; long long host_waits = 0;
    int* all_hdr = nullptr; int* h_hdr = nullptr;      // the ranks' 4-word message headers: device copy, pinned host copy
    unsigned long long* mg_keys = nullptr; unsigned long long* mg_keys2 = nullptr; int* mg_vals = nullptr; int* mg_vals2 = nullptr;
    void* mg_cub = nullptr; size_t mg_cub_bytes = 0;
    bool canonical = false;             // patch ids are the reference's m_ppatches indices (collect order, no holes)
};

namespace {

using namespace pmk;

template <typename T>
int dalloc(pmk_ctx* ctx, T** p, size_t count) {
    CUDA_TRY(cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T)));
    ctx->owned.push_back(*p);
    return PMK_OK;
}

int store_params(pmk_ctx* ctx, StoreParams& sp, uint64_t seed) {
    int rc = cand_params(ctx, sp.cp, seed);
    if (rc) return rc;
    pmk_store* s = ctx->store;
    sp.st = s->d;
    sp.neighbor_cos = cosf(120.0f / M_PI * 180.0f);                 // pmmvps.cpp:124, value preserved
    sp.neighbor_radius_cos = cos(120.0f * M_PI / 180.0f);           // pmmvps.cpp:150
    sp.neighbor_threshold = ctx->neighbor_threshold;
    sp.neighbor_threshold1 = ctx->neighbor_threshold1;
    sp.neighbor_threshold2 = ctx->neighbor_threshold2;
    sp.quad_threshold = ctx->cfg.quad_threshold;
    sp.max_patches_cell = 2 * ctx->cfg.csize * ctx->cfg.csize;     // propagate.cpp:24-25
    sp.nb_scratch = s->nb_scratch;
    return PMK_OK;
}

int store_init(pmk_ctx* ctx) {
    if (ctx->store) return PMK_OK;
    int rc = upload_views(ctx);
    if (rc) return rc;
    if (ctx->cfg.nviews > CAND_MAXV) return fail(PMK_ERR_ARG, "pmk: the patch store supports at most 128 views");
    // a dest cell receives the patches of two source cells (<= MAX_NUM_OF_PATCHES each) with two tries per patch: SRC_MAX sources and
    // NEW_MAX staged patches per cell bound MAX_NUM_OF_PATCHES = 2 * csize^2 by 16
    if (2 * ctx->cfg.csize * ctx->cfg.csize > SRC_MAX / 2 || 4 * ctx->cfg.csize * ctx->cfg.csize > NEW_MAX / 2 * 2)
        return fail(PMK_ERR_ARG, "pmk: csize > 2 is not supported by the patch store (MAX_NUM_OF_PATCHES = 2 * csize^2 must be <= 16)");
    pmk_store* s = new pmk_store();
    ctx->store = s;
    const int nv = ctx->cfg.nviews;
    s->cell_base.assign(nv + 1, 0);
    int max_diag = 1;
    for (int v = 0; v < nv; ++v) {
        const ViewConst& vc = ctx->h_views[v];
        if (vc.gw >= 65535 || vc.gh >= 32767) return fail(PMK_ERR_ARG, "pmk: cell grid too large for packed cell indices");
        if ((long long)vc.gw * vc.gh >= (1ll << 24)) return fail(PMK_ERR_ARG, "pmk: more than 2^24 cells per view (the collect key keeps 24 bits of cell index)");
        s->cell_base[v + 1] = s->cell_base[v] + vc.gw * vc.gh;
        max_diag = std::max(max_diag, std::min(vc.gw, vc.gh));
    }
    StoreDev& d = s->d;
    d.total_cells = s->cell_base[nv];
    d.maxv = nv;
    // a cell holds its m_pgrids entries (every view that matches the surface adds its patches until the cell's own sweep trims
    // it) and its m_vpgrids entries (views that see the surface without matching it): both grow with the number of views
    d.cell_cap = ctx->cfg.cell_capacity > 0 ? ctx->cfg.cell_capacity : std::min(1024, std::max(96, 4 * nv));
    if (d.cell_cap > 1024) return fail(PMK_ERR_ARG, "pmk: cell_capacity exceeds 1024");
    // grids first (their size is fixed by the views), then the patch arrays from what is left of the device memory
    if ((rc = dalloc(ctx, &d.ccount, d.total_cells)) || (rc = dalloc(ctx, &d.cslots, (size_t)d.total_cells * d.cell_cap)) ||
        (rc = dalloc(ctx, &d.dmap, d.total_cells)))
        return rc;
    s->group = ctx->cfg.sweep_group <= 0 ? 1 : std::min(std::min(ctx->cfg.sweep_group, nv), (int)GROUP_MAX);
    s->max_tasks = max_diag * s->group;
    d.stage_cap = s->max_tasks * NEW_MAX;
    const size_t per_patch = 3 * 16 + 4 * 4 + 4 * (size_t)nv * 4 + 64;          // store arrays + sort / scratch words per patch
    if (ctx->cfg.max_patches > 0) d.cap = ctx->cfg.max_patches;
    else {
        // one iteration creates up to ~1.5 patches per cell before Filter::run compacts the store; removed patches keep their
        // slot until then.  Default: 4 per cell of all views, within 70 % of the memory that is still free after the grids,
        // the staging region and ~8 GB of per-warp scratch.
        size_t free_b = 0, total_b = 0;
        CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
        const size_t reserve = (size_t)d.stage_cap * per_patch + ((size_t)8 << 30);
        const size_t usable = free_b > reserve ? (size_t)((free_b - reserve) * 0.7) : 0;
        const size_t by_mem = usable / (per_patch + (size_t)nv * 4);             // + the gather buffer row
        d.cap = (int)std::min<size_t>(std::min<size_t>((size_t)4 * d.total_cells, by_mem), (size_t)0x3fffffff);
        if (d.cap < 1024) return fail(PMK_ERR_CUDA, "pmk: not enough device memory for the patch store");
    }
    const size_t tot = (size_t)d.cap + d.stage_cap;
    if ((rc = dalloc(ctx, &d.coord, tot)) || (rc = dalloc(ctx, &d.normal, tot)) || (rc = dalloc(ctx, &d.scal, tot)) ||
        (rc = dalloc(ctx, &d.nimg, tot)) || (rc = dalloc(ctx, &d.nvimg, tot)) || (rc = dalloc(ctx, &d.state, tot)) || (rc = dalloc(ctx, &d.birth, tot)) ||
        (rc = dalloc(ctx, &d.images, tot * d.maxv)) || (rc = dalloc(ctx, &d.cells, tot * d.maxv)) ||
        (rc = dalloc(ctx, &d.vimages, tot * d.maxv)) || (rc = dalloc(ctx, &d.vcells, tot * d.maxv)) ||
        (rc = dalloc(ctx, &d.counters, SC_COUNT)) || (rc = dalloc(ctx, &s->cell_base_d, nv + 1)))
        return rc;
    d.cell_base = s->cell_base_d;
    CUDA_TRY(cudaMemcpyAsync(s->cell_base_d, s->cell_base.data(), (nv + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = dalloc(ctx, &s->rem_list, d.cap)) || (rc = dalloc(ctx, &s->task_new, s->max_tasks)) || (rc = dalloc(ctx, &s->order, s->max_tasks + 1)) || (rc = dalloc(ctx, &s->final_id, d.stage_cap)) || (rc = dalloc(ctx, &s->alive_list, d.stage_cap)) ||
        (rc = dalloc(ctx, &s->stats, SS_COUNT)) || (rc = dalloc(ctx, &s->step_max, 4)) || (rc = dalloc(ctx, &s->keys, d.cap)) || (rc = dalloc(ctx, &s->keys2, d.cap)) ||
        (rc = dalloc(ctx, &s->vals, d.cap)) || (rc = dalloc(ctx, &s->vals2, d.cap)) || (rc = dalloc(ctx, &s->f_tmp, d.cap)) ||
        (rc = dalloc(ctx, &s->i_tmp, d.cap + 1)) || (rc = dalloc(ctx, &s->i_tmp2, d.cap + 1)) || (rc = dalloc(ctx, &s->i_tmp3, d.cap + 1)) || (rc = dalloc(ctx, &s->small, 16)))
        return rc;
    CandParams cp;
    if ((rc = cand_params(ctx, cp, 0))) return rc;                       // sizes cand_grid and the pairwise scratch
    if ((rc = dalloc(ctx, &s->nb_scratch, (size_t)ctx->cand_grid * CAND_WARPS * NB_STRIDE))) return rc;
    CUDA_TRY(cudaMemsetAsync(s->nb_scratch, 0, (size_t)ctx->cand_grid * CAND_WARPS * NB_STRIDE * sizeof(int), ctx->stream));
    s->gather_bytes = (size_t)d.cap * d.maxv * sizeof(int);
    s->gather_bytes = std::max(s->gather_bytes, (size_t)d.cap * sizeof(float4));
    CUDA_TRY(cudaMalloc(&s->gather_tmp, s->gather_bytes));
    ctx->owned.push_back(s->gather_tmp);
    size_t b1 = 0, b2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, s->keys, s->keys2, s->vals, s->vals2, d.cap, 0, 64, ctx->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, s->i_tmp, s->i_tmp2, d.cap + 1, ctx->stream);
    s->cub_bytes = std::max(b1, b2);
    CUDA_TRY(cudaMalloc(&s->cub_tmp, s->cub_bytes));
    ctx->owned.push_back(s->cub_tmp);
    // the reference re-constructs std::default_random_engine on every propagatePatch call (propagate.cpp:139-141), so
    // its jitter is always the first four draws; same libstdc++, same values
    {
        std::default_random_engine generator;
        std::uniform_real_distribution<float> distribution(-0.5, 0.5);
        for (int i = 0; i < 4; ++i) s->jitter[i] = distribution(generator);
    }
    CUDA_TRY(cudaMemsetAsync(d.counters, 0, SC_COUNT * sizeof(int), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d.ccount, 0, (size_t)d.total_cells * sizeof(int), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d.cslots, 0xff, (size_t)d.total_cells * d.cell_cap * sizeof(int), ctx->stream));   // SLOT_FREE everywhere, once
    CUDA_TRY(cudaMemsetAsync(d.dmap, 0xff, (size_t)d.total_cells * sizeof(unsigned long long), ctx->stream));
    CUDA_TRY(cudaMemsetAsync(d.state, 0, tot * sizeof(int), ctx->stream));
    s->n = 0;
    s->canonical = true;
    return PMK_OK;
}

int store_check_overflow(pmk_ctx* ctx) {
    pmk_store* s = ctx->store;
    int c[SC_COUNT];
    CUDA_TRY(cudaMemcpyAsync(c, s->d.counters, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    s->n = c[SC_N];
    if (c[SC_OVERFLOW]) return fail(PMK_ERR_CAPACITY, "pmk: " + std::to_string(c[SC_OVERFLOW]) + " cell registrations dropped (raise pmk_config.cell_capacity)");
    if (c[SC_FULL]) return fail(PMK_ERR_CAPACITY, "pmk: patch store full, " + std::to_string(c[SC_FULL]) + " patches dropped (raise pmk_config.max_patches)");
    if (c[SC_NBOVER]) return fail(PMK_ERR_CAPACITY, "pmk: findNeighbors scratch overflow");
    if (c[SC_MSGOVER]) return fail(PMK_ERR_CAPACITY, "pmk: a multi-GPU step produced more new patches / removals than one rank's message holds");
    return PMK_OK;
}

template <typename T>
int gather_rows(pmk_ctx* ctx, T* arr, const int* perm, int nalive, int row);

template <typename T>
__global__ void k_gather_rows(const T* __restrict__ src, T* __restrict__ dst, const int* __restrict__ perm, long long total, int row) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const long long q = t / row; const int i = (int)(t % row);
    dst[t] = src[(size_t)perm[q] * row + i];
}

template <typename T>
int gather_rows(pmk_ctx* ctx, T* arr, const int* perm, int nalive, int row) {
    if (nalive <= 0) return PMK_OK;
    pmk_store* s = ctx->store;
    const long long total = (long long)nalive * row;
    k_gather_rows<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(arr, (T*)s->gather_tmp, perm, total, row);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(arr, s->gather_tmp, (size_t)total * sizeof(T), cudaMemcpyDeviceToDevice, ctx->stream));
    return PMK_OK;
}

__global__ void k_fill_state(int* state, int nalive, int n) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) state[q] = q < nalive ? 1 : 0;
}

// PatchManager::collectPatches order (patch_manager.cpp:75-104): s->vals2[0..nalive) = patch ids in m_ppatches order
int store_order(pmk_ctx* ctx, const StoreParams& sp, int* nalive_out) {
    pmk_store* s = ctx->store;
    cudaStream_t st = ctx->stream;
    const int n = s->n;
    int nalive = 0;
    if (n > 0) {
        k5_collect_keys<<<(n + 255) / 256, 256, 0, st>>>(sp, n, s->keys, s->vals);
        ctx->launches++;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(s->cub_tmp, s->cub_bytes, s->keys, s->keys2, s->vals, s->vals2, n, 0, 64, st));
        ctx->launches++;
        CUDA_TRY(cudaMemsetAsync(s->small, 0, sizeof(int), st));
        k5_count_alive<<<(n + 255) / 256, 256, 0, st>>>(s->keys2, n, s->small);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(&nalive, s->small, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    *nalive_out = nalive;
    return PMK_OK;
}

// collectPatches + setDepthMapsVGridsVPGridsAddPatchV(additive): after this, patch ids are the reference's m_ppatches indices
int store_rebuild(pmk_ctx* ctx, int additive) {
    pmk_store* s = ctx->store;
    StoreParams sp;
    int rc = store_params(ctx, sp, 0);
    if (rc) return rc;
    StoreDev& d = s->d;
    cudaStream_t st = ctx->stream;
    if ((rc = store_check_overflow(ctx))) return rc;
    const int n = s->n;
    int nalive = 0;
    if ((rc = store_order(ctx, sp, &nalive))) return rc;
    if (n > 0) {
        if ((rc = gather_rows(ctx, d.coord, s->vals2, nalive, 1)) || (rc = gather_rows(ctx, d.normal, s->vals2, nalive, 1)) ||
            (rc = gather_rows(ctx, d.scal, s->vals2, nalive, 1)) || (rc = gather_rows(ctx, d.nimg, s->vals2, nalive, 1)) ||
            (rc = gather_rows(ctx, d.nvimg, s->vals2, nalive, 1)) || (rc = gather_rows(ctx, d.birth, s->vals2, nalive, 1)) ||
            (rc = gather_rows(ctx, d.images, s->vals2, nalive, d.maxv)) || (rc = gather_rows(ctx, d.cells, s->vals2, nalive, d.maxv)) ||
            (rc = gather_rows(ctx, d.vimages, s->vals2, nalive, d.maxv)) || (rc = gather_rows(ctx, d.vcells, s->vals2, nalive, d.maxv)))
            return rc;
        k_fill_state<<<(n + 255) / 256, 256, 0, st>>>(d.state, nalive, n);
        ctx->launches++;
    }
    s->n = nalive;
    CUDA_TRY(cudaMemcpyAsync(d.counters + SC_N, &s->n, sizeof(int), cudaMemcpyHostToDevice, st));
    k_reset_cells<<<(d.total_cells + 255) / 256, 256, 0, st>>>(d);
    ctx->launches++;
    if (nalive > 0) {
        const int blocks = std::min(ctx->sm_count * 8, (nalive + 3) / 4);
        k5_register<<<blocks, 128, 0, st>>>(sp, nalive, 0, 0, 0);
        ctx->launches++;
        const long long tot = (long long)nalive * ctx->cfg.nviews;
        k5_depth_maps<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(sp, nalive);
        ctx->launches++;
        const int grid = std::max(1, std::min(ctx->cand_grid, (nalive + CAND_WARPS - 1) / CAND_WARPS));
        k5_set_vimages<<<grid, CAND_WARPS * 32, CAND_WARPS * sizeof(WarpScratch), st>>>(sp, nalive, additive);
        ctx->launches++;
    }
    CUDA_TRY(cudaGetLastError());
    s->canonical = true;
    return store_check_overflow(ctx);
}

int kill_flagged(pmk_ctx* ctx, const float* gains, const int* flags, int* killed) {
    pmk_store* s = ctx->store;
    if (s->n <= 0) { *killed = 0; return PMK_OK; }
    CUDA_TRY(cudaMemsetAsync(s->small, 0, sizeof(int), ctx->stream));
    k_kill_flagged<<<(s->n + 255) / 256, 256, 0, ctx->stream>>>(s->d, s->n, gains, flags, s->small);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(killed, s->small, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    s->canonical = false;
    return PMK_OK;
}

// Filter::filterSmallGroups (filter.cpp:432-525): directed isNeighbor edges (k9_group_edges), then the reference's order-dependent
// labelling as a min-ancestor fixed point on the device (k9_label_*, pmk_filter.cuh) -- no adjacency ever leaves the GPU
int small_groups(pmk_ctx* ctx, const StoreParams& sp, int* flags_dev, int* removed) {
    pmk_store* s = ctx->store;
    const int n = s->n;
    *removed = 0;
    if (n <= 0) return PMK_OK;
    cudaStream_t st = ctx->stream;
    const int grid = std::max(1, std::min(ctx->cand_grid, (n + CAND_WARPS - 1) / CAND_WARPS));
    k9_group_edges<<<grid, CAND_WARPS * 32, 0, st>>>(sp, n, 0, s->i_tmp, nullptr, nullptr);
    ctx->launches++;
    CUDA_TRY(cudaMemsetAsync(s->i_tmp + n, 0, sizeof(int), st));
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(s->cub_tmp, s->cub_bytes, s->i_tmp, s->i_tmp2, n + 1, st));
    ctx->launches++;
    int nedges = 0;
    CUDA_TRY(cudaMemcpyAsync(&nedges, s->i_tmp2 + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if ((size_t)nedges > s->adj_cap) {                       // adjacency buffer, grown on demand and kept
        if (s->adj) CUDA_TRY(cudaFree(s->adj));
        s->adj = nullptr; s->adj_cap = 0;
        const size_t cap = 2 * (size_t)nedges + 1024;        // the store grows ~1.3x per iteration: one allocation serves a whole run
        CUDA_TRY(cudaMalloc((void**)&s->adj, cap * sizeof(int)));
        s->adj_cap = cap;
    }
    k9_group_edges<<<grid, CAND_WARPS * 32, 0, st>>>(sp, n, 1, nullptr, s->i_tmp2, s->adj);
    ctx->launches++;
    int* L = s->i_tmp3;                                        // labels; i_tmp (degrees) becomes the group sizes afterwards
    const int tb = (n + 255) / 256;
    k9_label_init<<<tb, 256, 0, st>>>(n, L);
    ctx->launches++;
    const int rgrid = std::max(1, std::min(ctx->sm_count * 8, (n + 7) / 8));
    for (int round = 0; round < 4096; ++round) {
        CUDA_TRY(cudaMemsetAsync(s->small, 0, sizeof(int), st));
        k9_label_relax<<<rgrid, 256, 0, st>>>(n, s->i_tmp2, s->adj, L, s->small);
        k9_label_jump<<<tb, 256, 0, st>>>(n, L, s->small);
        ctx->launches += 2;
        int changed = 0;
        CUDA_TRY(cudaMemcpyAsync(&changed, s->small, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (!changed) break;
    }
    const int threshold = std::max(20, n / 10000);
    CUDA_TRY(cudaMemsetAsync(s->i_tmp, 0, (size_t)n * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(s->small, 0, sizeof(int), st));
    k9_group_sizes<<<tb, 256, 0, st>>>(s->d, n, L, s->i_tmp);
    k9_group_flags<<<tb, 256, 0, st>>>(s->d, n, L, s->i_tmp, threshold, s->i_tmp2, s->small);
    ctx->launches += 2;
    CUDA_TRY(cudaMemcpyAsync(flags_dev, s->i_tmp2, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(removed, s->small, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    return PMK_OK;
}

// one filter stage on a clean store; outputs stay in s->f_tmp / s->i_tmp / s->i_tmp3 for pmk_filter_stage
int filter_stage(pmk_ctx* ctx, int stage, int* killed) {
    pmk_store* s = ctx->store;
    StoreParams sp;
    int rc = store_params(ctx, sp, 0);
    if (rc) return rc;
    *killed = 0;
    const int n = s->n;
    if (n <= 0) return PMK_OK;
    cudaStream_t st = ctx->stream;
    const int grid = std::max(1, std::min(ctx->cand_grid, (n + CAND_WARPS - 1) / CAND_WARPS));
    const size_t smem = CAND_WARPS * sizeof(WarpScratch);
    switch (stage) {
        case 1:
            k6_gains<<<grid, CAND_WARPS * 32, 0, st>>>(sp, n, s->f_tmp);
            ctx->launches++;
            CUDA_TRY(cudaGetLastError());
            return kill_flagged(ctx, s->f_tmp, nullptr, killed);
        case 2:
            WS_DISPATCH(ctx->cfg.wsize, (k7_exact<WS><<<grid, CAND_WARPS * 32, smem, st>>>(sp, n, s->i_tmp)));
            ctx->launches++;
            CUDA_TRY(cudaGetLastError());
            return kill_flagged(ctx, nullptr, s->i_tmp, killed);
        case 3:
            k8_neighbor<<<grid, CAND_WARPS * 32, 0, st>>>(sp, n, s->i_tmp, s->i_tmp3, s->f_tmp);
            ctx->launches++;
            CUDA_TRY(cudaGetLastError());
            return kill_flagged(ctx, nullptr, s->i_tmp, killed);
        case 4:
            if ((rc = small_groups(ctx, sp, s->i_tmp, killed))) return rc;
            { int k2 = 0; return kill_flagged(ctx, nullptr, s->i_tmp, &k2); }
    }
    return fail(PMK_ERR_ARG, "pmk_filter_stage: unknown stage");
}

// NCCL is resolved at run time (dlopen) so that the library neither links a second NCCL into a process that already has
// torch's, nor needs one when it runs on a single GPU.
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
        if (api.handle) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
            api.AllGather = (decltype(api.AllGather))dlsym(api.handle, "ncclAllGather");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        }
    }
    return (api.handle && api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather) ? &api : nullptr;
}

// one CTA of NW warps per dest cell (pmk_cell.cuh)
template <int WS, int NW, int MINB>
int launch_cells(pmk_ctx* ctx, const StoreParams& sp, const SweepArgs& a) {
    typedef CellGeom<WS, NW> Gm;
    const size_t csmem = ((sizeof(WarpScratch) + sizeof(SweepScratch) + sizeof(CellCta) + 15) & ~(size_t)15) +
                         (size_t)Gm::nslots(ctx->params.tau) * Gm::SLOT * sizeof(float);
    // the slot count depends on tau (a context property): opt in to the largest size once, ask the occupancy for the size at hand
    static bool attr_done = false;
    if (!attr_done) {
        const size_t cmax = ((sizeof(WarpScratch) + sizeof(SweepScratch) + sizeof(CellCta) + 15) & ~(size_t)15) + (size_t)Gm::nslots(PMK_MAX_TAU) * Gm::SLOT * sizeof(float);
        CUDA_TRY(cudaFuncSetAttribute(k4_cells<WS, NW, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cmax));
        attr_done = true;
    }
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k4_cells<WS, NW, MINB>, NW * 32, csmem));
    if (per_sm < 1) return fail(PMK_ERR_CUDA, "pmk: k4_cells does not fit on an SM");
    if (getenv("PMK_VERBOSE")) { static bool said = false; if (!said) { said = true; fprintf(stderr, "pmk: k4_cells<%d,%d,%d> %zu B smem, %d CTAs/SM\n", WS, NW, MINB, csmem, per_sm); } }
    const int cgrid = std::max(1, std::min(a.ntasks, std::min(ctx->sm_count * per_sm, ctx->cand_grid * CAND_WARPS)));
    k4_cells<WS, NW, MINB><<<cgrid, NW * 32, csmem, ctx->stream>>>(sp, a);
    return PMK_OK;
}

template <int WS>
int launch_sweep(pmk_ctx* ctx, const StoreParams& sp, const SweepArgs& sa) {
    pmk_store* s = ctx->store;
    cudaStream_t st = ctx->stream;
    CUDA_TRY(cudaMemsetAsync(s->d.counters + SC_REM, 0, sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(s->step_max, 0, sizeof(unsigned long long), st));
    static const bool vlap = getenv("PMK_VERBOSE") != nullptr;
#define PMK_LAP(slot) do { if (vlap && s->laps) k_lap<<<1, 1, 0, st>>>(s->laps, s->laps + 16, slot); } while (0)
    PMK_LAP(-1);
    if (sa.ntasks > 0) {
        // longest-first order (k4_plan), then one CTA per dest cell, handed out through SC_NEXT (pmk_cell.cuh)
        k4_plan<<<1, 1024, 0, st>>>(sp, sa, s->order);
        CUDA_TRY(cudaMemsetAsync(s->d.counters + SC_NEXT, 0, sizeof(int), st));
        // 8 warps per CTA, 3 CTAs per SM (80 registers): measured best on config 2 against 4 / 12 / 16 warps and 2 / 4 / 5 CTAs per SM
        const int rc_cells = launch_cells<WS, 8, 3>(ctx, sp, sa);
        if (rc_cells) return rc_cells;
        ctx->launches += 2;
    }
    k_fold_step<<<1, 1, 0, st>>>(s->stats, s->step_max);
    PMK_LAP(0);                                   // plan + sweep kernel
    if (s->nranks <= 1) {
        k4_apply_remove<<<std::max(1, std::min(ctx->sm_count, (sa.ntasks + 3) / 4)), 128, 0, st>>>(sp, s->rem_list, s->d.cap);
        k4_apply_scan<<<1, 1024, 0, st>>>(sp, s->task_new, sa.ntasks, s->final_id, s->alive_list, s->small + 8);
        k4_apply_add<<<std::max(1, std::min(ctx->sm_count * 2, sa.ntasks)), 128, 0, st>>>(sp, s->final_id, s->alive_list, s->small + 8);
        ctx->launches += 3;
        PMK_LAP(1);                               // apply (single GPU)
    } else {
        // pack this rank's mutations, all-gather over NVLink, apply every rank's in rank order
        NcclApi* api = nccl_api();
        if (!api || !s->nccl_comm) return fail(PMK_ERR_STATE, "pmk: multi-GPU sweep without a communicator (pmk_comm_init)");
        k4_pack_scan<<<1, 1024, 0, st>>>(sp, s->task_new, sa.ntasks, s->rem_list, s->ml, s->msg, s->pack_ids);
        k4_pack_copy<<<ctx->sm_count, 128, 0, st>>>(sp, sa, s->ml, s->msg, s->pack_ids);
        k_stamp<<<1, 1, 0, st>>>(s->step_max + 1);
        PMK_LAP(2);                               // pack
        // headers first: every rank learns how much the others produced, the payload gather then moves max-over-ranks words
        ncclResult_t nr = api->AllGather(s->msg, s->all_hdr, 4, ncclInt32, (ncclComm_t)s->nccl_comm, st);
        if (nr != ncclSuccess) return fail(PMK_ERR_CUDA, std::string("ncclAllGather: ") + (api->GetErrorString ? api->GetErrorString(nr) : "error"));
        CUDA_TRY(cudaMemcpyAsync(s->h_hdr, s->all_hdr, (size_t)s->nranks * 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        PMK_LAP(3);                               // header all-gather (includes waiting for the slowest rank)
        { const auto h0 = std::chrono::steady_clock::now(); CUDA_TRY(cudaStreamSynchronize(st)); s->host_wait_s += std::chrono::duration<double>(std::chrono::steady_clock::now() - h0).count(); s->host_waits++; }
        size_t need = 4;
        for (int r = 0; r < s->nranks; ++r) need = std::max(need, (size_t)4 + (size_t)s->h_hdr[4 * r + 1] + (size_t)s->h_hdr[4 * r] * s->ml.rec_words);
        need = std::min((need + 255) & ~(size_t)255, s->ml.words());
        s->ml.stride = need;
        nr = api->AllGather(s->msg, s->all_msgs, need, ncclInt32, (ncclComm_t)s->nccl_comm, st);
        if (nr != ncclSuccess) return fail(PMK_ERR_CUDA, std::string("ncclAllGather: ") + (api->GetErrorString ? api->GetErrorString(nr) : "error"));
        k_fold_exchange<<<1, 1, 0, st>>>(s->stats, s->step_max + 1, (unsigned long long)(need + 4) * 4ull * s->nranks);
        PMK_LAP(4);                               // host round trip + payload all-gather
        k4_unpack_remove<<<ctx->sm_count, 128, 0, st>>>(sp, s->ml, s->all_msgs, s->nranks);
        PMK_LAP(5);                               // removals
        RankOff ro;
        ro.off[0] = 0;
        for (int r = 0; r < s->nranks; ++r) ro.off[r + 1] = ro.off[r] + s->h_hdr[4 * r];
        const int nrec = ro.off[s->nranks];
        if (nrec > 0) {
            k4_unpack_keys<<<(nrec + 255) / 256, 256, 0, st>>>(s->ml, s->all_msgs, s->nranks, ro, s->mg_keys, s->mg_vals);
            CUDA_TRY(cub::DeviceRadixSort::SortPairs(s->mg_cub, s->mg_cub_bytes, s->mg_keys, s->mg_keys2, s->mg_vals, s->mg_vals2, nrec, 0, 40, st));
        }
        PMK_LAP(6);                               // keys + sort
        k4_unpack_scan<<<1, 32, 0, st>>>(sp, s->ml, s->all_msgs, s->nranks, s->rec_base);
        if (nrec > 0) k4_unpack_add<<<ctx->sm_count * 2, 128, 0, st>>>(sp, s->ml, s->all_msgs, s->nranks, ro, s->rec_base, s->mg_vals2);
        ctx->launches += 7;
        PMK_LAP(7);                               // scan + add
    }
#undef PMK_LAP
    CUDA_TRY(cudaGetLastError());
    return PMK_OK;
}

// Wavefront steps [step_first, step_first + step_count) of Propagate::propagatePmImage for views [img_first, img_first + nimg):
// step k carries anti-diagonal k (from the far corner on odd iterations, propagate.cpp:80-86) of every view of the group.
// (rank, nranks): this GPU takes the dest cells of a step whose global number G has G % nranks == rank (multi-GPU partition).
int sweep_views(pmk_ctx* ctx, int iter, int img_first, int nimg, int step_first, int step_count, uint64_t seed, int only_x = -1, const ForceIO* force = nullptr) {
    pmk_store* s = ctx->store;
    StoreParams sp;
    int rc = store_params(ctx, sp, seed);
    if (rc) return rc;
    if (nimg > GROUP_MAX) return fail(PMK_ERR_ARG, "pmk: sweep group too large");
    const int inc = (iter % 2 == 1) ? -1 : 1;
    SweepArgs sa;
    std::memset(&sa, 0, sizeof(sa));
    sa.inc = inc; sa.iter = iter;
    sa.jitter_mode = ctx->cfg.jitter_mode;
    for (int i = 0; i < 4; ++i) sa.jitter[i] = s->jitter[i];
    if (force) sa.force = *force;
    sa.wslot_base = 0;
    sa.rem_list = s->rem_list; sa.task_new = s->task_new; sa.stats = s->stats; sa.order = s->order; sa.step_max = s->step_max; sa.cell_ns = s->cell_ns; sa.phase_ns = s->phase_ns;
    int max_steps = 0;
    for (int g = 0; g < nimg; ++g) { const ViewConst& vc = ctx->h_views[img_first + g]; max_steps = std::max(max_steps, vc.gw + vc.gh - 1); }
    sa.rank = s->rank; sa.nranks = s->nranks;
    for (int k = step_first; k < step_first + step_count && k < max_steps; ++k) {
        sa.ngroup = 0;
        int gtasks = 0;                                              // tasks of the whole step, all ranks
        for (int g = 0; g < nimg; ++g) {
            const ViewConst& vc = ctx->h_views[img_first + g];
            const int gw = vc.gw, gh = vc.gh, ndiag = gw + gh - 1;
            if (k >= ndiag) continue;
            const int d = inc > 0 ? k : ndiag - 1 - k;
            int xlo = std::max(0, d - gh + 1), xhi = std::min(gw - 1, d);                    // the whole anti-diagonal
            if (only_x >= 0) { xlo = std::max(xlo, only_x); xhi = std::min(xhi, only_x); }   // a single dest cell (pmk_propagate_forced)
            if (xhi < xlo) continue;
            const int m = sa.ngroup++;
            sa.g_img[m] = img_first + g; sa.g_diag[m] = d; sa.g_xlo[m] = xlo; sa.g_off[m] = gtasks;
            gtasks += xhi - xlo + 1;
        }
        sa.g_off[sa.ngroup] = gtasks;
        sa.ntasks = gtasks > s->rank ? (gtasks - s->rank + s->nranks - 1) / s->nranks : 0;   // global tasks G with G % nranks == rank
        if (sa.ntasks <= 0 && s->nranks <= 1) continue;
        if (sa.ntasks > s->max_tasks) return fail(PMK_ERR_CAPACITY, "pmk: sweep step exceeds the staging capacity");
        {
            static const int room_weight = getenv("PMK_ROOM_WEIGHT") ? atoi(getenv("PMK_ROOM_WEIGHT")) : 2;
            sa.room_weight = room_weight;
        }
        WS_DISPATCH(ctx->cfg.wsize, { if ((rc = launch_sweep<WS>(ctx, sp, sa))) return rc; });
    }
    s->canonical = false;
    return PMK_OK;
}

}  // namespace
