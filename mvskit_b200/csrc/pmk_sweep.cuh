// mvskit_b200/csrc/pmk_sweep.cuh -- K4: Propagate::propagatePmImage / propagatePatch / generatePatch
// (pmmvps/propagate.cpp:72-237) as an anti-diagonal wavefront over DEST cells.
//
// Schedule "PMS1".  The reference sweeps the cells of one view in raster order (Gauss-Seidel): source cell k hands its
// patches to cells k + inc and k + inc * gwidth.  Everything written into dest cell D = (x, y) therefore comes from
// (x, y - inc) [visited first] and (x - inc, y), both on the previous anti-diagonal, and nothing on D's own anti-diagonal
// reads or writes D.  One wavefront step = one anti-diagonal of one view; one CTA (k4_cells, pmk_cell.cuh) owns one dest cell
// and replays, in the reference's order, every propagatePatch call that targets it: sort, trim to MAX_NUM_OF_PATCHES, two tries per call
// (fill an empty slot at the jittered cell centre, or challenge the worst patch at its own pixel), generatePatch,
// computeNcc, preProcess, refinePatch (PMR1), postProcess, removePatch / addPatch.  All grid mutations other than D's own
// list are staged and applied between steps (k4_apply_*), so a step reads one consistent snapshot and is deterministic.
// Differences from the raster order, by construction: (a) the row wrap of `index + inc` at a row end and the out-of-range
// `index + inc * gwidth` on the last row (propagate.cpp:106-108, an out-of-bounds access in the reference) are skipped;
// (b) postProcess's store reads (setVImagesVGrids, check) see the snapshot of the step start plus D's own changes.
#pragma once

#include "pmk_store.cuh"

namespace pmk {

enum SweepStat { SS_CALLS = 0, SS_TRIES, SS_GEN_NULL, SS_NCC_LOSE, SS_FAIL0, SS_FAIL1, SS_ADDED, SS_REPLACED, SS_TRIMMED, SS_EVALS,
                 SS_CELL_NS, SS_STEP_MAX_NS, SS_STEPS, SS_NCCL_NS, SS_MSG_BYTES, SS_COUNT = 16 };   // ..., summed dest-cell time, summed per-step slowest cell,
                                                                                                 // steps, time inside the step exchanges (multi-GPU), bytes gathered

constexpr int GROUP_MAX = 128;         // views whose wavefronts one step can carry
constexpr int LKEEP = 40;              // entries of a cell list the sweep keeps (MAX_NUM_OF_PATCHES <= 32, plus slack)
constexpr int REM_OVERLAY = 128;       // removals of a dest cell that the step's own check() calls can see

// Teacher forcing (tests only, pmk_propagate_forced): the tries of the step's single dest cell start from recorded hypotheses --
// the patch as the reference's refinePatch left it -- so that everything decided AFTER the hypothesis (the m_ncc test against the
// worst patch, postProcess, setVImagesVGrids, check, add / replace) can be compared try by try.  code: 0 generatePatch NULL,
// 1 lost, 2 preProcess == -1, 3 refined (coord / normal / scal = {ncc, dscale, ascale} / images are the refined patch).
struct ForceIO {
    int ntries, stride;
    const int* code; const float* ncc0; const float4* coord; const float4* normal; const float4* scal; const int* nimg; const int* images;
    int* o_ntries; int* o_outcome; int* o_full; int* o_post; int* o_nimg; int* o_images; int* o_cells; int* o_nvimg; int* o_vimages; int* o_vcells; float* o_tmp;
};

struct SweepArgs {
    // dest cells of this step, enumerated over the whole step (every view of the group, every cell of its anti-diagonal): for group
    // member g, view g_img[g], cells (g_xlo[g] + t, g_diag[g] - g_xlo[g] - t), t in [0, g_off[g + 1] - g_off[g]); global task = g_off[g] + t.
    // Multi-GPU: rank r of n takes the global tasks G with G % n == r (interleaved: neighbouring cells of a diagonal go to different GPUs,
    // so heavy regions spread evenly); its local task L stands for G = L * n + r.  ids and creation numbers follow the global order.
    int ngroup, ntasks, inc;               // ntasks = LOCAL tasks of this rank
    int rank, nranks;
    int g_img[GROUP_MAX], g_diag[GROUP_MAX], g_xlo[GROUP_MAX], g_off[GROUP_MAX + 1];
    int iter;                              // Propagate::run(iter)
    int jitter_mode;                       // 0: the reference's four constant draws (propagate.cpp:139-141), 1: Philox per try
    float jitter[4];
    int* rem_list;                         // patches removed this step (SC_REM counts them)
    int* task_new;                         // [max_tasks] staged entries used by each task
    const int* order;                      // [ntasks] task ids, most expensive first (k4_plan)
    int wslot_base;                        // first per-CTA scratch slot of this launch
    int room_weight;                       // cost weight of a try into a cell that still has room (k4_plan)
    unsigned long long* stats;             // SweepStat
    unsigned long long* step_max;          // slowest dest cell of this step, ns (one word per step, zeroed by the host)
    float* cell_ns;                        // optional [total_cells]: time spent on each dest cell in its last visit (profiling)
    ForceIO force;                         // code == nullptr in production
    unsigned long long* phase_ns;          // optional [8]: warp time by phase of a try (profiling, pmk_debug_phase_times): generatePatch + computeNcc,
                                           // preProcess, refinePatch, postProcess, its store-reading tail, waiting for the turn / commit, tries, refined tries
};

__device__ __forceinline__ int sweep_global_task(const SweepArgs& sa, int task) { return task * sa.nranks + sa.rank; }
// group member of global task G (linear scan: the group is the number of views swept together, at most GROUP_MAX)
__device__ __forceinline__ int sweep_group_of(const SweepArgs& sa, int G) {
    int g = 0;
    while (g + 1 < sa.ngroup && G >= sa.g_off[g + 1]) ++g;
    return g;
}

struct SweepScratch {                  // per warp
    int l_snap[LKEEP];                 // snapshot of the dest cell's list the try was evaluated against
    int vimg[CAND_MAXV];
    int vcell[CAND_MAXV];
    int cells[CAND_MAXV];
};

// PatchManager::sortPatches (patch_manager.cpp:406-433), descending: the reference's O(n^2) swap sort, verbatim semantics
__device__ __forceinline__ void swap_sort_desc(int* id, float* ncc, int n) {
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (ncc[i] < ncc[j]) { const float t = ncc[i]; ncc[i] = ncc[j]; ncc[j] = t; const int u = id[i]; id[i] = id[j]; id[j] = u; }
}

// sortPatches(desc) + "keep the first `keep`" (propagate.cpp:88-97,124-134) in one streaming pass over the m_pgrids entries
// of global cell c: id[] / ncc[] / birth[] end up holding the `keep` best patches by (m_ncc descending, creation ascending),
// in that order.  A cell can hold hundreds of entries (every view that sees the surface registers its patches there until the
// cell's own sweep trims it), so nothing longer than `keep` is ever materialised.  TRIM: every other live entry is a patch
// the reference removes (removePatch): the first REM_OVERLAY go to removed[] (visible to this cell's check() calls), the
// rest straight to the step's removal list.  Returns the kept count; *ntrim counts the trimmed ones.
template <bool TRIM>
__device__ __forceinline__ int warp_load_cell_topk(const StoreParams& sp, int c, int keep, int* id, float* ncc, unsigned int* birth,
                                                   int* removed, int* nremoved, int* rem_list, int* ntrim, int lane) {
    const StoreDev& st = sp.st;
    const int n = min(st.ccount[c], st.cell_cap);
    int m = 0, trimmed = 0, nrem = TRIM ? *nremoved : 0;
    // short cells (the common case): the reference's own procedure, verbatim -- vector order (creation), its O(n^2) swap sort,
    // then everything past `keep` is removed -- so ties in m_ncc resolve exactly as in the reference
    int live = 0;
    for (int base = 0; base < n; base += 32) {
        const int s = base + lane;
        int e = SLOT_TOMB;
        if (s < n) e = st.cslots[(size_t)c * st.cell_cap + s];
        const bool ok = e != SLOT_TOMB && e >= 0 && st.state[e] == 1;
        const unsigned msk = __ballot_sync(0xffffffffu, ok);
        if (ok) { const int pos = live + __popc(msk & ((1u << lane) - 1u)); if (pos < LKEEP) { id[pos] = e; ncc[pos] = st.scal[e].x; birth[pos] = st.birth[e]; } }
        live += __popc(msk);
    }
    __syncwarp();
    if (live <= LKEEP) {
        if (lane == 0) {
            for (int i = 1; i < live; ++i) {                 // vector order = creation order
                const int e = id[i]; const float v = ncc[i]; const unsigned int b = birth[i];
                int j = i - 1;
                while (j >= 0 && birth[j] > b) { id[j + 1] = id[j]; ncc[j + 1] = ncc[j]; birth[j + 1] = birth[j]; --j; }
                id[j + 1] = e; ncc[j + 1] = v; birth[j + 1] = b;
            }
            swap_sort_desc(id, ncc, live);
            if (TRIM) for (int i = keep; i < live; ++i) {
                if (nrem < REM_OVERLAY) removed[nrem++] = id[i];
                else rem_list[atomicAdd(st.counters + SC_REM, 1)] = id[i];
                ++trimmed;
            }
        }
        m = min(live, keep);
        if (TRIM) { *nremoved = __shfl_sync(0xffffffffu, nrem, 0); *ntrim = __shfl_sync(0xffffffffu, trimmed, 0); }
        __syncwarp();
        return m;
    }
    // long cells: stream, keeping only the `keep` best by (m_ncc descending, creation ascending)
    for (int base = 0; base < n; base += 32) {
        const int s = base + lane;
        int e = SLOT_TOMB;
        if (s < n) e = st.cslots[(size_t)c * st.cell_cap + s];
        const bool ok = e != SLOT_TOMB && e >= 0 && st.state[e] == 1;
        const float v = ok ? st.scal[e].x : 0.0f;
        const unsigned int b = ok ? st.birth[e] : 0u;
        unsigned msk = __ballot_sync(0xffffffffu, ok);
        while (msk) {
            const int l = __ffs(msk) - 1;
            msk &= msk - 1;
            const int el = __shfl_sync(0xffffffffu, e, l);
            const float vl = __shfl_sync(0xffffffffu, v, l);
            const unsigned int bl = __shfl_sync(0xffffffffu, b, l);
            int out = -1;                                    // the entry that falls off the list, if any
            if (lane == 0) {
                int pos = m;                                 // insertion point: after every better entry
                while (pos > 0 && (ncc[pos - 1] < vl || (ncc[pos - 1] == vl && birth[pos - 1] > bl))) --pos;
                if (pos >= keep) out = el;
                else {
                    if (m == keep) { out = id[keep - 1]; } else ++m;
                    for (int j = m - 1; j > pos; --j) { id[j] = id[j - 1]; ncc[j] = ncc[j - 1]; birth[j] = birth[j - 1]; }
                    id[pos] = el; ncc[pos] = vl; birth[pos] = bl;
                }
                if (TRIM && out >= 0) {
                    if (nrem < REM_OVERLAY) removed[nrem++] = out;
                    else rem_list[atomicAdd(st.counters + SC_REM, 1)] = out;
                    ++trimmed;
                }
            }
        }
    }
    m = __shfl_sync(0xffffffffu, m, 0);
    if (TRIM) { *nremoved = __shfl_sync(0xffffffffu, nrem, 0); *ntrim = __shfl_sync(0xffffffffu, trimmed, 0); }
    __syncwarp();
    return m;
}

enum TryOutcome { TRY_GEN_NULL = 0, TRY_LOSE, TRY_FAIL0, TRY_FAIL1, TRY_ACCEPT, TRY_DIVERGED };

struct Cand {                    // a candidate after Optim::refinePatch (its image list is in shared memory)
    V4 X, N;
    int nv, nvv;
    float ncc, dscale, ascale, tmp;
};

// ---- longest-first schedule of a step ---------------------------------------------------------------------------------------
// A step ends with its slowest dest cell, so the cells are started in decreasing order of a cost estimate: the number of
// propagatePatch calls aimed at the cell (patches of its two source cells whose reference view is the swept view), times
// `room_weight` while the cell still has room (every try then runs the full optimisation instead of first having to beat the
// worst patch's NCC, which most tries fail).
// One block; counting sort on the estimate.  The order only changes WHEN a cell runs, never what it computes.

__global__ void __launch_bounds__(1024) k4_plan(const StoreParams sp, const SweepArgs sa, int* __restrict__ order) {
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    constexpr int NBIN = 8 * SRC_MAX + 2;
    __shared__ int hist[NBIN], offs[NBIN];
    for (int i = threadIdx.x; i < NBIN; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    const int maxp = sp.max_patches_cell;
    for (int base = 0; base < sa.ntasks; base += blockDim.x) {
        const int task = base + threadIdx.x;
        int est = 0;
        if (task < sa.ntasks) {
            const int G = sweep_global_task(sa, task), g = sweep_group_of(sa, G);
            const int img = sa.g_img[g];
            const int gw = p.views[img].gw, gh = p.views[img].gh;
            const int x = sa.g_xlo[g] + (G - sa.g_off[g]), y = sa.g_diag[g] - x;
            int nsrc = 0;
            for (int side = 0; side < 2; ++side) {
                const int sx = side == 0 ? x : x - sa.inc, sy = side == 0 ? y - sa.inc : y;
                if (sx < 0 || gw <= sx || sy < 0 || gh <= sy) continue;
                const int c = st.cell_base[img] + sy * gw + sx;
                const int n = min(st.ccount[c], st.cell_cap);
                int k = 0;
                for (int s2 = 0; s2 < n; ++s2) {
                    const int e = st.cslots[(size_t)c * st.cell_cap + s2];
                    if (e == SLOT_TOMB || e < 0) continue;
                    if (st.images[(size_t)e * st.maxv] == img) ++k;
                }
                nsrc += min(k, maxp);
            }
            const int cD = st.cell_base[img] + y * gw + x;
            int nd = 0;
            const int n = min(st.ccount[cD], st.cell_cap);
            for (int s2 = 0; s2 < n; ++s2) { const int e = st.cslots[(size_t)cD * st.cell_cap + s2]; if (e != SLOT_TOMB && e >= 0) ++nd; }
            est = min(NBIN - 1, nsrc * (nd < maxp ? sa.room_weight : 1));
            atomicAdd(&hist[est], 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int b = NBIN - 1; b >= 0; --b) { offs[b] = acc; acc += hist[b]; }
    }
    __syncthreads();
    for (int base = 0; base < sa.ntasks; base += blockDim.x) {
        const int task = base + threadIdx.x;
        if (task >= sa.ntasks) continue;
        // recompute the estimate (cheaper than keeping it: tasks can exceed the block size)
        const int G = sweep_global_task(sa, task), g = sweep_group_of(sa, G);
        const int img = sa.g_img[g];
        const int gw = p.views[img].gw, gh = p.views[img].gh;
        const int x = sa.g_xlo[g] + (G - sa.g_off[g]), y = sa.g_diag[g] - x;
        int nsrc = 0;
        for (int side = 0; side < 2; ++side) {
            const int sx = side == 0 ? x : x - sa.inc, sy = side == 0 ? y - sa.inc : y;
            if (sx < 0 || gw <= sx || sy < 0 || gh <= sy) continue;
            const int c = st.cell_base[img] + sy * gw + sx;
            const int n = min(st.ccount[c], st.cell_cap);
            int k = 0;
            for (int s2 = 0; s2 < n; ++s2) {
                const int e = st.cslots[(size_t)c * st.cell_cap + s2];
                if (e == SLOT_TOMB || e < 0) continue;
                if (st.images[(size_t)e * st.maxv] == img) ++k;
            }
            nsrc += min(k, maxp);
        }
        const int cD = st.cell_base[img] + y * gw + x;
        int nd = 0;
        const int n = min(st.ccount[cD], st.cell_cap);
        for (int s2 = 0; s2 < n; ++s2) { const int e = st.cslots[(size_t)cD * st.cell_cap + s2]; if (e != SLOT_TOMB && e >= 0) ++nd; }
        const int est = min(NBIN - 1, nsrc * (nd < maxp ? sa.room_weight : 1));
        order[atomicAdd(&offs[est], 1)] = task;
    }
}

__global__ void k_fold_step(unsigned long long* stats, const unsigned long long* step_max) {
    stats[SS_STEP_MAX_NS] += *step_max;
    stats[SS_STEPS] += 1;
}

// multi-GPU: time stamps around a step's exchange (device clock, on the stream) and the bytes it gathered
__global__ void k_stamp(unsigned long long* t) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(*t)); }
// profiling (PMK_VERBOSE): acc[slot] += now - *last; *last = now -- laps on the stream's own timeline
__global__ void k_lap(unsigned long long* acc, unsigned long long* last, int slot) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (slot >= 0) acc[slot] += now - *last;
    *last = now;
}
__global__ void k_fold_exchange(unsigned long long* stats, const unsigned long long* t0, unsigned long long bytes) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    stats[SS_NCCL_NS] += now - *t0;
    stats[SS_MSG_BYTES] += bytes;
}

// ---- apply: removals ---------------------------------------------------------------------------------------------------------------
// PatchManager::removePatch (patch_manager.cpp:303-325): erase the patch from every cell it is registered in.  One warp per patch.
__device__ __forceinline__ void erase_from_cell(const StoreDev& st, int c, int entry) {
    const int n = min(st.ccount[c], st.cell_cap);
    for (int s = 0; s < n; ++s)
        if (st.cslots[(size_t)c * st.cell_cap + s] == entry) { st.cslots[(size_t)c * st.cell_cap + s] = SLOT_TOMB; }
}

__global__ void k4_apply_remove(const StoreParams sp, const int* __restrict__ rem_list, int nrem_max) {
    const StoreDev& st = sp.st;
    const int nrem = min(st.counters[SC_REM], nrem_max);
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = gwarp; r < nrem; r += nwarps) {
        const int id = rem_list[r];
        if (st.state[id] != 1) continue;                   // listed twice (cannot happen within one cell; defensive)
        __syncwarp();
        const int ni = st.nimg[id], nv = st.nvimg[id];
        for (int i = lane; i < ni; i += 32) {
            const int img = st.images[(size_t)id * st.maxv + i], c = st.cells[(size_t)id * st.maxv + i];
            erase_from_cell(st, cell_global(sp, img, cell_x(c), cell_y(c)), id);
        }
        for (int i = lane; i < nv; i += 32) {
            const int img = st.vimages[(size_t)id * st.maxv + i], c = st.vcells[(size_t)id * st.maxv + i];
            erase_from_cell(st, cell_global(sp, img, cell_x(c), cell_y(c)), (int)((unsigned)id | SLOT_V));
        }
        __syncwarp();
        if (lane == 0) st.state[id] = 0;
    }
}

// Slots that were never written hold SLOT_FREE, which no scan takes for an erased slot: an appender reserves slot s with
// atomicAdd(ccount) and stores it afterwards, and a concurrent insert into the same cell already sees s < ccount -- a stale
// SLOT_TOMB left there by an earlier epoch would be claimed by its CAS and then overwritten by the appender's store (one
// registration lost, silently).  k_reset_cells restores SLOT_FREE wherever ccount is zeroed.
constexpr int SLOT_FREE = -1;          // 0xffffffff: reads as an erased m_vpgrids entry, so readers skip it too

// PatchManager::init's empty grids (patch_manager.cpp:38-51): every used slot back to SLOT_FREE, count 0, depth map = m_MAXDEPTH
__global__ void k_reset_cells(const StoreDev st) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= st.total_cells) return;
    const int n = min(st.ccount[c], st.cell_cap);
    int* slots = st.cslots + (size_t)c * st.cell_cap;
    for (int s = 0; s < n; ++s) slots[s] = SLOT_FREE;
    st.ccount[c] = 0;
    st.dmap[c] = ~0ull;
}

// register `entry` in cell c: reuse an erased slot, else append
__device__ __forceinline__ void insert_into_cell(const StoreDev& st, int c, int entry) {
    int* slots = st.cslots + (size_t)c * st.cell_cap;
    const int n = min(st.ccount[c], st.cell_cap);
    for (int s = 0; s < n; ++s)
        if (slots[s] == SLOT_TOMB && atomicCAS(slots + s, SLOT_TOMB, entry) == SLOT_TOMB) return;
    const int s = atomicAdd(st.ccount + c, 1);
    if (s < st.cell_cap) slots[s] = entry;
    else { atomicSub(st.ccount + c, 1); atomicAdd(st.counters + SC_OVERFLOW, 1); }
}

// PatchManager::addPatch (patch_manager.cpp:158-189) for patch `id` whose lists are already in the store
__device__ __forceinline__ void warp_register_patch(const StoreParams& sp, int id, bool with_v, bool with_depth, int lane) {
    const StoreDev& st = sp.st;
    const int ni = st.nimg[id];
    for (int i = lane; i < ni; i += 32) {
        const int img = st.images[(size_t)id * st.maxv + i], c = st.cells[(size_t)id * st.maxv + i];
        insert_into_cell(st, cell_global(sp, img, cell_x(c), cell_y(c)), id);
    }
    if (with_v) {
        const int nv = st.nvimg[id];
        for (int i = lane; i < nv; i += 32) {
            const int img = st.vimages[(size_t)id * st.maxv + i], c = st.vcells[(size_t)id * st.maxv + i];
            insert_into_cell(st, cell_global(sp, img, cell_x(c), cell_y(c)), (int)((unsigned)id | SLOT_V));
        }
    }
    if (with_depth) {
        const V4 X = f4v(st.coord[id]);
        for (int v = lane; v < sp.cp.p.nviews; v += 32) update_depth_map(sp, id, X, v);
    }
}

// one block: exclusive scan of the staged-and-alive counts -> final ids (task order); bumps SC_N / SC_BIRTH.
// final_id[task * NEW_MAX + j] = id or -1.  Launched with 1024 threads; thread t owns tasks t, t + 1024, ...
__global__ void __launch_bounds__(1024) k4_apply_scan(const StoreParams sp, const int* __restrict__ task_new, int ntasks, int* __restrict__ final_id,
                                                      int* __restrict__ alive_list, int* __restrict__ alive_count) {
    const StoreDev& st = sp.st;
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    const int n0 = st.counters[SC_N];
    const int b0 = st.counters[SC_BIRTH];
    for (int base = 0; base < ntasks; base += 1024) {
        const int t = base + tid;
        int cnt = 0;
        if (t < ntasks) {
            const int k = task_new[t];
            for (int j = 0; j < k; ++j) cnt += st.state[st.cap + t * NEW_MAX + j] == 1 ? 1 : 0;
        }
        int incl = cnt;                                     // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += v; }
            s_warp[lane] = w;                               // inclusive over warps
        }
        __syncthreads();
        const int carry = s_carry;
        int excl = carry + (wid > 0 ? s_warp[wid - 1] : 0) + incl - cnt;
        if (t < ntasks) {
            const int k = task_new[t];
            for (int j = 0; j < k; ++j) {
                const int sid = st.cap + t * NEW_MAX + j;
                int fid = -1;
                if (st.state[sid] == 1) {
                    if (n0 + excl < st.cap) { fid = n0 + excl; st.birth[sid] = (unsigned int)(b0 + excl); }
                    else atomicAdd(st.counters + SC_FULL, 1);
                    alive_list[excl] = t * NEW_MAX + j;             // compact list of the surviving staged entries, in (task, slot) order
                    ++excl;
                }
                final_id[t * NEW_MAX + j] = fid;
            }
        }
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (tid == 0) {
        const int total = s_carry;
        st.counters[SC_N] = min(st.cap, n0 + total);
        st.counters[SC_BIRTH] = b0 + total;
        *alive_count = total;
    }
}

// copy every staged, surviving patch to its final slot and register it (one warp per entry of the scan's compact list)
__global__ void k4_apply_add(const StoreParams sp, const int* __restrict__ final_id, const int* __restrict__ alive_list, const int* __restrict__ alive_count) {
    const StoreDev& st = sp.st;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int total = *alive_count;
    for (int pos = gwarp; pos < total; pos += nwarps) {
        const int w = alive_list[pos];
        const int fid = final_id[w];
        if (fid < 0) continue;
        const int sid = st.cap + w;
        if (lane == 0) {
            st.coord[fid] = st.coord[sid]; st.normal[fid] = st.normal[sid]; st.scal[fid] = st.scal[sid];
            st.nimg[fid] = st.nimg[sid]; st.nvimg[fid] = st.nvimg[sid]; st.state[fid] = 1; st.birth[fid] = st.birth[sid];
        }
        const int ni = st.nimg[sid], nv = st.nvimg[sid];
        for (int i = lane; i < ni; i += 32) { st.images[(size_t)fid * st.maxv + i] = st.images[(size_t)sid * st.maxv + i]; st.cells[(size_t)fid * st.maxv + i] = st.cells[(size_t)sid * st.maxv + i]; }
        for (int i = lane; i < nv; i += 32) { st.vimages[(size_t)fid * st.maxv + i] = st.vimages[(size_t)sid * st.maxv + i]; st.vcells[(size_t)fid * st.maxv + i] = st.vcells[(size_t)sid * st.maxv + i]; }
        __syncwarp();
        const bool deep = sp.cp.p.depth != 0;                 // addPatch returns before m_vpgrids / depth maps at m_depth == 0 (:173-175)
        warp_register_patch(sp, fid, deep, deep, lane);
        __syncwarp();
    }
}

}  // namespace pmk

// =====================================================================================================================
// Multi-GPU: the store is replicated, the dest cells of a step are dealt out to the ranks in turn (global task G -> rank G % nranks), and the step's mutations
// (new patches, removals) travel between the ranks as one fixed-layout message per rank (ncclAllGather over NVLink).
// Every rank then applies ALL messages, so ids, creation numbers and grids stay identical everywhere.
//   message = int hdr[4] {n_new, n_rem, overflow, 0} | int rem[n_rem] | rec[n_new][16 + 4 * maxv]      (compact: only what the step produced)
//   record  = coord4, normal4, scal4 (as int bits), nimg, nvimg, global task index, slot in the task,
//             images[maxv], cells[maxv], vimages[maxv], vcells[maxv]
// The ranks first gather their 4-word headers; the host reads them and gathers max-over-ranks words of payload, so a step moves what
// it produced instead of the worst case (34 MB per rank at 128 views).
// Final ids and creation numbers are assigned in (global task, slot) order -- the order the single-GPU scan uses -- so an
// N-GPU run produces the single-GPU store bit for bit.
// =====================================================================================================================
namespace pmk {

struct MsgLayout {
    int rem_cap, rec_cap, rec_words;
    size_t stride;                       // words between two ranks' messages in the gathered buffer (set per step)
    __host__ __device__ size_t words() const { return 4 + (size_t)rem_cap + (size_t)rec_cap * rec_words; }        // capacity of one message
    __device__ const int* msg_of(const int* all, int rk) const { return all + (size_t)rk * stride; }
    __device__ static const int* rec_of(const int* msg, int r, int rec_words) { return msg + 4 + msg[1] + (size_t)r * rec_words; }
};

// one block: list the staged patches that survived the step (any order: the receivers sort by (task, slot)), copy the removals, fill the header
__global__ void __launch_bounds__(1024) k4_pack_scan(const StoreParams sp, const int* __restrict__ task_new, int ntasks, const int* __restrict__ rem_list, MsgLayout ml,
                                                     int* __restrict__ msg, int* __restrict__ pack_ids) {
    const StoreDev& st = sp.st;
    __shared__ int s_n, s_over;
    if (threadIdx.x == 0) { s_n = 0; s_over = 0; }
    __syncthreads();
    for (int t = threadIdx.x; t < ntasks; t += blockDim.x) {
        const int k = task_new[t];
        for (int j = 0; j < k; ++j) {
            const int sid = st.cap + t * NEW_MAX + j;
            if (st.state[sid] != 1) continue;
            const int pos = atomicAdd(&s_n, 1);
            if (pos < ml.rec_cap) pack_ids[pos] = sid; else s_over = 1;
        }
    }
    __syncthreads();
    int nrem = st.counters[SC_REM];
    if (nrem > ml.rem_cap) { nrem = ml.rem_cap; if (threadIdx.x == 0) s_over = 1; }
    for (int i = threadIdx.x; i < nrem; i += blockDim.x) msg[4 + i] = rem_list[i];
    __syncthreads();
    if (threadIdx.x == 0) { msg[0] = min(s_n, ml.rec_cap); msg[1] = nrem; msg[2] = s_over; msg[3] = 0; }
}

__global__ void k4_pack_copy(const StoreParams sp, const SweepArgs sa, MsgLayout ml, int* __restrict__ msg, const int* __restrict__ pack_ids) {
    const StoreDev& st = sp.st;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n = msg[0];
    for (int r = gwarp; r < n; r += nwarps) {
        const int sid = pack_ids[r];
        int* rec = msg + 4 + msg[1] + (size_t)r * ml.rec_words;
        if (lane == 0) {
            const float4 c = st.coord[sid], m = st.normal[sid], s = st.scal[sid];
            rec[0] = __float_as_int(c.x); rec[1] = __float_as_int(c.y); rec[2] = __float_as_int(c.z); rec[3] = __float_as_int(c.w);
            rec[4] = __float_as_int(m.x); rec[5] = __float_as_int(m.y); rec[6] = __float_as_int(m.z); rec[7] = __float_as_int(m.w);
            rec[8] = __float_as_int(s.x); rec[9] = __float_as_int(s.y); rec[10] = __float_as_int(s.z); rec[11] = __float_as_int(s.w);
            rec[12] = st.nimg[sid]; rec[13] = st.nvimg[sid];
            const int task = (sid - st.cap) / NEW_MAX;
            rec[14] = sweep_global_task(sa, task);
            rec[15] = (sid - st.cap) % NEW_MAX;
        }
        const int ni = st.nimg[sid], nv = st.nvimg[sid], mv = st.maxv;
        for (int i = lane; i < ni; i += 32) { rec[16 + i] = st.images[(size_t)sid * mv + i]; rec[16 + mv + i] = st.cells[(size_t)sid * mv + i]; }
        for (int i = lane; i < nv; i += 32) { rec[16 + 2 * mv + i] = st.vimages[(size_t)sid * mv + i]; rec[16 + 3 * mv + i] = st.vcells[(size_t)sid * mv + i]; }
    }
}

// removals of every rank's message (a patch may be listed by several ranks: the first warp to see it erases it)
__global__ void k4_unpack_remove(const StoreParams sp, MsgLayout ml, const int* __restrict__ all, int nranks) {
    const StoreDev& st = sp.st;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int rk = 0; rk < nranks; ++rk) {
        const int* msg = ml.msg_of(all, rk);
        const int nrem = msg[1];
        for (int r = gwarp; r < nrem; r += nwarps) {
            const int id = msg[4 + r];
            const int ni = st.nimg[id], nv = st.nvimg[id];
            for (int i = lane; i < ni; i += 32) {
                const int img = st.images[(size_t)id * st.maxv + i], c = st.cells[(size_t)id * st.maxv + i];
                erase_from_cell(st, cell_global(sp, img, cell_x(c), cell_y(c)), id);
            }
            for (int i = lane; i < nv; i += 32) {
                const int img = st.vimages[(size_t)id * st.maxv + i], c = st.vcells[(size_t)id * st.maxv + i];
                erase_from_cell(st, cell_global(sp, img, cell_x(c), cell_y(c)), (int)((unsigned)id | SLOT_V));
            }
            __syncwarp();
            if (lane == 0) st.state[id] = 0;
        }
    }
}

struct RankOff { int off[65]; };        // off[r] = records of the ranks before r (host-side prefix of the gathered headers)
__device__ __forceinline__ int rank_of_record(const RankOff& ro, int nranks, int i) {
    int rk = 0;
    while (rk + 1 < nranks && i >= ro.off[rk + 1]) ++rk;
    return rk;
}

// sort keys of every rank's records: (global task, slot).  One thread per record.
__global__ void k4_unpack_keys(MsgLayout ml, const int* __restrict__ all, int nranks, RankOff ro, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ro.off[nranks]) return;
    const int rk = rank_of_record(ro, nranks, i), r = i - ro.off[rk];
    const int* rec = MsgLayout::rec_of(ml.msg_of(all, rk), r, ml.rec_words);
    keys[i] = (unsigned long long)(unsigned int)rec[14] * NEW_MAX + (unsigned int)rec[15];
    vals[i] = i;
}

// one thread: how many records there are in all, capacity check, counters
__global__ void k4_unpack_scan(const StoreParams sp, MsgLayout ml, const int* __restrict__ all, int nranks, int* __restrict__ rec_base) {
    const StoreDev& st = sp.st;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int total = 0;
    for (int rk = 0; rk < nranks; ++rk) {
        const int* msg = ml.msg_of(all, rk);
        if (msg[2]) atomicAdd(st.counters + SC_MSGOVER, 1);
        total += msg[0];
    }
    const int n = st.counters[SC_N], b = st.counters[SC_BIRTH];
    int take = total;
    if (n + total > st.cap) { take = max(0, st.cap - n); atomicAdd(st.counters + SC_FULL, total - take); }
    rec_base[0] = n; rec_base[1] = b; rec_base[2] = take;
    st.counters[SC_N] = n + take;
    st.counters[SC_BIRTH] = b + total;
}

// record at sorted position pos gets id n0 + pos and creation number b0 + pos (one warp per record)
__global__ void k4_unpack_add(const StoreParams sp, MsgLayout ml, const int* __restrict__ all, int nranks, RankOff ro, const int* __restrict__ rec_base,
                              const int* __restrict__ sorted_vals) {
    const StoreDev& st = sp.st;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool deep = sp.cp.p.depth != 0;
    const int n0 = rec_base[0], b0 = rec_base[1], take = rec_base[2];
    for (int pos = gwarp; pos < take; pos += nwarps) {
        const int i = sorted_vals[pos];
        const int rk = rank_of_record(ro, nranks, i), r = i - ro.off[rk];
        const int* rec = MsgLayout::rec_of(ml.msg_of(all, rk), r, ml.rec_words);
        const int fid = n0 + pos, mv = st.maxv;
        const int ni = rec[12], nv = rec[13];
        if (lane == 0) {
            st.coord[fid] = make_float4(__int_as_float(rec[0]), __int_as_float(rec[1]), __int_as_float(rec[2]), __int_as_float(rec[3]));
            st.normal[fid] = make_float4(__int_as_float(rec[4]), __int_as_float(rec[5]), __int_as_float(rec[6]), __int_as_float(rec[7]));
            st.scal[fid] = make_float4(__int_as_float(rec[8]), __int_as_float(rec[9]), __int_as_float(rec[10]), __int_as_float(rec[11]));
            st.nimg[fid] = ni; st.nvimg[fid] = nv; st.state[fid] = 1; st.birth[fid] = (unsigned int)(b0 + pos);
        }
        for (int k = lane; k < ni; k += 32) { st.images[(size_t)fid * mv + k] = rec[16 + k]; st.cells[(size_t)fid * mv + k] = rec[16 + mv + k]; }
        for (int k = lane; k < nv; k += 32) { st.vimages[(size_t)fid * mv + k] = rec[16 + 2 * mv + k]; st.vcells[(size_t)fid * mv + k] = rec[16 + 3 * mv + k]; }
        __syncwarp();
        warp_register_patch(sp, fid, deep, deep, lane);
        __syncwarp();
    }
}

// order-independent digest of the live store (replica consistency checks): sum over patches of a hash of coord, ncc and lists
__global__ void k_store_checksum(const StoreDev st, int n, unsigned long long* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n || st.state[q] != 1) return;
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](unsigned int v) { h ^= v; h *= 1099511628211ull; };
    const float4 c = st.coord[q], m = st.normal[q], s = st.scal[q];
    mix(__float_as_uint(c.x)); mix(__float_as_uint(c.y)); mix(__float_as_uint(c.z));
    mix(__float_as_uint(m.x)); mix(__float_as_uint(m.y)); mix(__float_as_uint(m.z));
    mix(__float_as_uint(s.x)); mix(__float_as_uint(s.y));
    const int ni = st.nimg[q], nv = st.nvimg[q];
    for (int i = 0; i < ni; ++i) { mix((unsigned)st.images[(size_t)q * st.maxv + i]); mix((unsigned)st.cells[(size_t)q * st.maxv + i]); }
    mix(0xffffffffu);
    for (int i = 0; i < nv; ++i) { mix((unsigned)st.vimages[(size_t)q * st.maxv + i]); mix((unsigned)st.vcells[(size_t)q * st.maxv + i]); }
    atomicAdd(out, h);
    atomicAdd(out + 1, 1ull);
}

}  // namespace pmk
