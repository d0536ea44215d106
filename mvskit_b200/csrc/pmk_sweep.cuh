// mvskit_b200/csrc/pmk_sweep.cuh -- K4: Propagate::propagatePmImage / propagatePatch / generatePatch
// (pmmvps/propagate.cpp:72-237) as an anti-diagonal wavefront over DEST cells.
//
// Schedule "PMS1".  The reference sweeps the cells of one view in raster order (Gauss-Seidel): source cell k hands its
// patches to cells k + inc and k + inc * gwidth.  Everything written into dest cell D = (x, y) therefore comes from
// (x, y - inc) [visited first] and (x - inc, y), both on the previous anti-diagonal, and nothing on D's own anti-diagonal
// reads or writes D.  One wavefront step = one anti-diagonal of one view; one warp owns one dest cell and replays, in the
// reference's order, every propagatePatch call that targets it: sort, trim to MAX_NUM_OF_PATCHES, two tries per call
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
                 SS_CELL_NS, SS_STEP_MAX_NS, SS_STEPS, SS_COOP_CELLS, SS_COOP_REFINES, SS_COUNT = 16 };   // ..., summed dest-cell time, summed per-step slowest cell, steps

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
    // dest cells of this step: for group member g, view g_img[g], cells (g_xlo[g] + t, g_diag[g] - g_xlo[g] - t),
    // t in [0, g_off[g + 1] - g_off[g]); task index = g_off[g] + t
    int ngroup, ntasks, inc;
    int g_img[GROUP_MAX], g_diag[GROUP_MAX], g_xlo[GROUP_MAX], g_off[GROUP_MAX + 1];
    // the same enumeration without the row-band restriction (multi-GPU: ids and creation numbers follow THIS order on every rank)
    int g_gxlo[GROUP_MAX], g_goff[GROUP_MAX];
    int iter;                              // Propagate::run(iter)
    int wpc;                               // warps cooperating on one dest cell: 1, 2 or 4 (CAND_WARPS / wpc cells per CTA)
    int jitter_mode;                       // 0: the reference's four constant draws (propagate.cpp:139-141), 1: Philox per try
    float jitter[4];
    int* rem_list;                         // patches removed this step (SC_REM counts them)
    int* task_new;                         // [max_tasks] staged entries used by each task
    const int* order;                      // [ntasks] task ids, most expensive first (k4_plan); order[max_tasks] = number of heavy tasks
    int heavy_slot;                        // index of that count inside `order`
    int range_sel;                         // 0: every task, 1: the heavy prefix, 2: the light rest
    int wslot_base;                        // first per-warp scratch slot of this launch
    int split;                             // host: run heavy / light cells as two concurrent launches on wide steps
    int heavy_est;                         // dest cells with at least this many estimated full-cost tries count as heavy
    int coop;                              // 1: full cells, 2: every cell with four warps refines one candidate with all of them (coop_refine)
    int room_weight;                       // cost weight of a try into a cell that still has room (k4_plan)
    unsigned long long* stats;             // SweepStat
    unsigned long long* step_max;          // slowest dest cell of this step, ns (one word per step, zeroed by the host)
    float* cell_ns;                        // optional [total_cells]: time spent on each dest cell in its last visit (profiling)
    ForceIO force;                         // code == nullptr in production
    unsigned long long* phase_ns;          // optional [8]: warp time by phase of a try (profiling, pmk_debug_phase_times): generatePatch + computeNcc,
                                           // preProcess, refinePatch, postProcess, its store-reading tail, waiting for the turn / commit, tries, refined tries
};

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

// PatchManager::computeNcc (patch_manager.cpp:401-404) by the whole warp: views spread over the evaluator groups
template <int WS>
__device__ __forceinline__ float warp_compute_ncc(const Params& p, WarpScratch& ws, V4 X, V4 N, int nv, int lane) {
    constexpr int GW = WS <= 8 ? 8 : 16;
    compute_weights(p, X, N, ws.images, nv, ws.units, lane);
    float incc = 2.0f;
    if (nv >= 2) {
        const int sz = min(p.tau, nv);
        warp_set_inccs<WS, GW>(p, X, N, ws.images, sz, 1, ws.inccs, lane);
        float score = 0.0f, tw = 0.0f;
        bool ref_ok = false;
        for (int i = 1; i < sz; ++i) {
            const float v = ws.inccs[i];
            if (v != 2.0f) { ref_ok = true; tw = xadd(tw, ws.units[i]); score = xadd(score, xmul(v, ws.units[i])); }
        }
        (void)ref_ok;
        incc = (tw == 0.0f) ? 2.0f : xdiv(score, tw);
    }
    __syncwarp();
    return xsub(1.0f, unrobustincc(incc));
}

// ---- one dest cell per CTA: the tries run speculatively on the CTA's warps and commit in order ------------------------------
// The tries of a dest cell (call c, try k) depend on each other only through the cell's list: whether it still has room, and
// otherwise which patch is the worst.  Each warp claims the next try, snapshots the list (seqlock), evaluates the try against
// that snapshot, waits for its turn (commit_ptr == try index), and commits if the snapshot's assumptions still hold -- the
// branch (room / full), the worst patch when full; the store-reading tail (stage B) is redone when the list changed at all.
// Otherwise it re-evaluates, now against the final state since it holds the turn.  The outcome is the sequential one.
struct CellShared {
    int l_id[LKEEP];
    float l_ncc[LKEEP];
    unsigned int l_birth[LKEEP];
    int src_id[SRC_MAX];
    int removed[REM_OVERLAY + NEW_MAX];
    int nl, nrem, nnew, nsrc;
    int version;                 // seqlock: odd while a commit is in progress
    int next_try, commit_ptr;
    // cooperative refinement (all warps of the cell work on ONE candidate): command, context and per-candidate results
    int coop, cmd;
    float rs_center[4], rs_ray[4], rs_x[4], rs_n[4];
    float rs_dscale;
    int rs_nv, rs_warp;          // rs_warp: CTA warp whose WarpScratch holds the image list
    unsigned long long rs_stream;
    double res_cost[PMR1_CANDS];
    double res_x[PMR1_CANDS][3];
};

struct TrySnap {
    int ver, np, w, nrem;
    float wncc;
};

__device__ __forceinline__ int ld_shared_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// consistent copy of the dest-cell list into warp-private scratch
__device__ __forceinline__ TrySnap take_snapshot(const CellShared& cs, int* my_id, int maxp, int lane) {
    TrySnap sn;
    while (true) {
        const int v0 = ld_shared_volatile(&cs.version);
        if (v0 & 1) { __nanosleep(50); continue; }
        const int np = ld_shared_volatile(&cs.nl);
        for (int i = lane; i < min(np, LKEEP); i += 32) my_id[i] = ld_shared_volatile(&cs.l_id[i]);
        sn.np = np;
        sn.w = np >= maxp ? ld_shared_volatile(&cs.l_id[maxp - 1]) : -1;
        sn.wncc = np >= maxp ? __int_as_float(ld_shared_volatile(reinterpret_cast<const int*>(&cs.l_ncc[maxp - 1]))) : 0.0f;
        sn.nrem = ld_shared_volatile(&cs.nrem);
        __syncwarp();
        const int v1 = ld_shared_volatile(&cs.version);
        sn.ver = v0;
        if (v0 == v1) break;
    }
    return sn;
}

enum TryOutcome { TRY_GEN_NULL = 0, TRY_LOSE, TRY_FAIL0, TRY_FAIL1, TRY_ACCEPT, TRY_DIVERGED };

struct Cand {                    // a candidate after stage A (its image list is in ws.images)
    V4 X, N;
    int nv, nvv;
    float ncc, dscale, ascale, tmp;
};

// ---- Optim::refinePatch by ALL warps of a dest cell --------------------------------------------------------------------------
// Same schedule, same arithmetic and therefore the same result as warp_refine, but the 8 candidates of a PMR1 level are spread
// over the 16 evaluator groups of four warps: one candidate per pair of groups, the pair splitting the non-reference views
// (both grab the reference texture).  The per-view terms travel to one lane and are summed there in view order, exactly as
// Optim::cost_func does, so the cost is bit-identical to the single-warp evaluation.  A level costs 1 + 3 texture grabs of
// latency instead of 2 x 6.  Every participating warp calls this with the same arguments (read from the cell's shared block).
template <int WS, typename Sync>
__device__ __forceinline__ float coop_refine(const CandParams& cp, CellShared& cs, const int* images, float* weights, V4& X, V4& N, int nv,
                                             float dscale, uint64_t stream, int cell_warp, int nwarps, int lane, Sync& cell_sync) {
    constexpr int GW = WS <= 8 ? 8 : 16;
    constexpr int G = 32 / GW;
    const Params& p = cp.p;
    const int grp = lane / GW, col = lane % GW;
    const float cmask = col < WS ? 1.0f : 0.0f;
    const unsigned gm = group_mask<GW>(lane);
    const double lb[3] = {-(double)__int_as_float(0x7f800000), -23.99999, -23.99999};
    const double ub[3] = {(double)__int_as_float(0x7f800000), 23.99999, 23.99999};
    RefineCtx rc;
    rc.center = X;
    rc.ref = images[0];
    rc.ray = sub4(X, ld4(p.views[rc.ref].center));
    rc.ray = div4(rc.ray, norm4(rc.ray));
    rc.dscale = dscale;
    if (cell_warp == 0) compute_weights(p, X, N, images, nv, weights, lane);      // m_weights of the UNREFINED patch (optim.cpp:490)
    double best[3];
    encode(cp, rc, X, N, best);
#pragma unroll
    for (int i = 0; i < 3; ++i) best[i] = fmax(fmin(best[i], ub[i]), lb[i]);
    double fbest = group_cost<WS, GW>(cp, rc, best, images, nv, col, cmask, gm);  // every group: the same value
    __syncwarp();
    const int tg = nwarps * G;                         // evaluator groups of the cell
    const int halves = tg >= 2 * PMR1_CANDS ? 2 : 1;   // groups per candidate
    const int gidx = cell_warp * G + grp;              // this group's index in the cell
    const int sz = min(p.tau, nv);
    const int minimum = min(p.min_image_num, sz);
    const int n_o = sz - 1;                            // non-reference views
    const int h0 = halves == 2 ? n_o / 2 : n_o;        // views [1, 1 + h0) go to half 0, the rest to half 1
    const unsigned pairmask = halves == 2 ? (GW == 8 ? (0xffffu << (16 * (lane / 16))) : 0xffffffffu) : gm;
    double r[3] = {4.0, 4.0, 4.0};
#pragma unroll 1
    for (int level = 0; level < PMR1_LEVELS; ++level) {
#pragma unroll 1
        for (int cb = 0; cb < PMR1_CANDS; cb += tg / halves) {
            const int cnd = cb + gidx / halves, half = gidx % halves;
            const bool live = cnd < PMR1_CANDS;
            uint32_t ctr[4] = {(uint32_t)stream, (uint32_t)(stream >> 32), (uint32_t)level, (uint32_t)cnd};
            philox4x32_10((uint32_t)cp.seed, (uint32_t)(cp.seed >> 32), ctr);
            double xc[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) xc[i] = fmax(fmin(__dadd_rn(best[i], __dmul_rn(r[i], uniform_pm1(ctr[i]))), ub[i]), lb[i]);
            // Optim::cost_func, this group's share of the views
            V4 coord, normal, px, py;
            decode(cp, rc, xc, coord, normal);
            get_paxes(p.views[rc.ref], coord, normal, p.level_scale, px, py);
            float t0[WS][3], t[WS][3];
            float inv0, inv;
            const bool ref_ok = group_grab<WS, GW>(p, images[0], coord, normal, px, py, col, cmask, t0, inv0, gm) >= 0;
            const int i_first = half == 0 ? 1 : 1 + h0, i_last = half == 0 ? 1 + h0 : sz;
            float val[PMK_MAX_TAU];
#pragma unroll
            for (int k = 0; k < PMK_MAX_TAU; ++k) val[k] = 0.0f;
            int okmask = 0;
#pragma unroll 1
            for (int i = i_first; i < i_last; ++i) {
                if (group_grab<WS, GW>(p, images[i], coord, normal, px, py, col, cmask, t, inv, gm) < 0) continue;
                const float d = group_dot<WS, GW>(t0, inv0, t, inv, gm);
                const float v = robustincc(__double2float_rn(1.0 - (double)d));
                const int k = i - i_first;
#pragma unroll
                for (int kk = 0; kk < PMK_MAX_TAU; ++kk) if (kk == k) val[kk] = v;
                okmask |= 1 << k;
            }
            // the second half hands its terms to the first (lanes GW apart)
            float oval[PMK_MAX_TAU];
#pragma unroll
            for (int k = 0; k < PMK_MAX_TAU; ++k) oval[k] = 0.0f;
            int ookmask = 0;
            if (halves == 2) {
#pragma unroll
                for (int k = 0; k < PMK_MAX_TAU / 2; ++k) oval[k] = __shfl_xor_sync(pairmask, val[k], GW);
                ookmask = __shfl_xor_sync(pairmask, okmask, GW);
            }
            if (live && half == 0 && col == 0) {
                double fc = 2.0;
                if (ref_ok) {
                    double ans = 0.0;
                    int denom = 0;
#pragma unroll
                    for (int k = 0; k < PMK_MAX_TAU; ++k) if (k < h0 && ((okmask >> k) & 1)) { ans += (double)val[k]; ++denom; }
                    if (halves == 2) {
#pragma unroll
                        for (int k = 0; k < PMK_MAX_TAU / 2; ++k) if (k < n_o - h0 && ((ookmask >> k) & 1)) { ans += (double)oval[k]; ++denom; }
                    }
                    fc = denom < minimum - 1 ? 2.0 : ans / (double)denom;
                }
                cs.res_cost[cnd] = fc;
                cs.res_x[cnd][0] = xc[0]; cs.res_x[cnd][1] = xc[1]; cs.res_x[cnd][2] = xc[2];
            }
        }
        cell_sync();
        // argmin over the level's candidates in index order, lowest index on ties (strict <), as the sequential loop does
        double fwin = cs.res_cost[0];
        int cwin = 0;
        for (int c = 1; c < PMR1_CANDS; ++c) { const double fc = cs.res_cost[c]; if (fc < fwin) { fwin = fc; cwin = c; } }
        if (fwin < fbest) { fbest = fwin; best[0] = cs.res_x[cwin][0]; best[1] = cs.res_x[cwin][1]; best[2] = cs.res_x[cwin][2]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = __dmul_rn(r[i], 0.6);
        cell_sync();
    }
    // optim.cpp:534-541: decode, normal.w = 0, ncc = 1.0 - unrobustincc(computeINCC(...)) with the stale weights (the caller's warp)
    V4 Xf, Nf;
    decode(cp, rc, best, Xf, Nf);
    float ncc = 0.0f;
    if (cell_warp == 0) {
        const float incc = group_incc<WS, GW>(p, Xf, Nf, images, nv, weights, col, cmask, gm);
        ncc = __double2float_rn(1.0 - (double)unrobustincc(incc));
    }
    __syncwarp();
    X = Xf;
    N = V4{Nf.x, Nf.y, Nf.z, 0.0f};
    return ncc;
}

#ifndef PMK_SWEEP_MINB
#define PMK_SWEEP_MINB 2
#endif
template <int WS>
__global__ void __launch_bounds__(CAND_WARPS * 32, PMK_SWEEP_MINB) k4_sweep(const StoreParams sp, const SweepArgs sa) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& ws = warp_scratch(smem_raw);
    SweepScratch& ss = reinterpret_cast<SweepScratch*>(smem_raw + CAND_WARPS * sizeof(WarpScratch))[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31, cta_warp = threadIdx.x >> 5;
    const int wpc = sa.wpc, cpc = CAND_WARPS / wpc;                   // warps per cell, cells per CTA
    const int cell_slot = cta_warp / wpc, warp = cta_warp % wpc;      // `warp` = rank among the warps of this dest cell
    CellShared& cs = reinterpret_cast<CellShared*>(smem_raw + CAND_WARPS * (sizeof(WarpScratch) + sizeof(SweepScratch)))[cell_slot];
    const CandParams& cp = sp.cp;
    const Params& p = cp.p;
    const StoreDev& st = sp.st;
    const int gwarp = sa.wslot_base + blockIdx.x * CAND_WARPS + cta_warp;
    const int nheavy = sa.range_sel ? sa.order[sa.heavy_slot] : 0;
    const int t_first = sa.range_sel == 2 ? nheavy : 0, t_last = sa.range_sel == 1 ? nheavy : sa.ntasks;
    // barrier over the warps of this cell only (the cells of a CTA run different numbers of tasks)
    auto cell_sync = [&]() {
        if (wpc == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(1 + cell_slot), "r"(wpc * 32) : "memory");
    };
    const int inc = sa.inc;
    const int maxp = sp.max_patches_cell;
    unsigned int stat[SS_COUNT];
#pragma unroll
    for (int i = 0; i < SS_COUNT; ++i) stat[i] = 0;

    for (int tslot = t_first + blockIdx.x * cpc + cell_slot; tslot < t_last; tslot += gridDim.x * cpc) {
        const int task = sa.order[tslot];
        int g = 0;
        while (g + 1 < sa.ngroup && task >= sa.g_off[g + 1]) ++g;
        const int img = sa.g_img[g];
        const ViewConst& vimgc = p.views[img];
        const int gw = vimgc.gw, gh = vimgc.gh;
        const int x = sa.g_xlo[g] + (task - sa.g_off[g]), y = sa.g_diag[g] - x;
        const int cD = st.cell_base[img] + y * gw + x;
        unsigned long long t_begin = 0;
        if (warp == 0 && lane == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
        // ================= preamble (warp 0): D's list, trim, sources =================
        if (warp == 0) {
            int nrem = 0;
            // ---- D's list: sortPatches + trim (propagate.cpp:123-134) ----
            {   // sortPatches recomputes a negative m_ncc first (patch_manager.cpp:411-415)
                const int n = min(st.ccount[cD], st.cell_cap);
                for (int base = 0; base < n; base += 32) {
                    const int s2 = base + lane;
                    int e = SLOT_TOMB;
                    if (s2 < n) e = st.cslots[(size_t)cD * st.cell_cap + s2];
                    const bool neg = e != SLOT_TOMB && e >= 0 && st.state[e] == 1 && st.scal[e].x < 0.0f;
                    unsigned msk = __ballot_sync(0xffffffffu, neg);
                    while (msk) {
                        const int l = __ffs(msk) - 1;
                        msk &= msk - 1;
                        const int el = __shfl_sync(0xffffffffu, e, l);
                        const int nv = min(st.nimg[el], CAND_MAXV);
                        for (int k = lane; k < nv; k += 32) ws.images[k] = st.images[(size_t)el * st.maxv + k];
                        __syncwarp();
                        const float v = warp_compute_ncc<WS>(p, ws, f4v(st.coord[el]), f4v(st.normal[el]), nv, lane);
                        if (lane == 0) st.scal[el].x = v;
                        __syncwarp();
                    }
                }
            }
            int ntrim = 0;
            const int nl = warp_load_cell_topk<true>(sp, cD, maxp, cs.l_id, cs.l_ncc, cs.l_birth, cs.removed, &nrem, sa.rem_list, &ntrim, lane);
            stat[SS_TRIMMED] += ntrim;
            // ---- sources: (x, y - inc) first, then (x - inc, y); each cell's sorted top-maxp, reference view == img ----
            int nsrc = 0;
            for (int side = 0; side < 2; ++side) {
                const int sx = side == 0 ? x : x - inc, sy = side == 0 ? y - inc : y;
                if (sx < 0 || gw <= sx || sy < 0 || gh <= sy) continue;
                int* tid = ss.vimg; float* tncc = reinterpret_cast<float*>(ss.vcell);       // scratch, free until stage B
                unsigned int* tbirth = reinterpret_cast<unsigned int*>(ss.cells);
                const int m = warp_load_cell_topk<false>(sp, st.cell_base[img] + sy * gw + sx, maxp, tid, tncc, tbirth, nullptr, nullptr, nullptr, nullptr, lane);
                for (int i = 0; i < m && nsrc < SRC_MAX; ++i) {
                    const int e = tid[i];
                    if (st.images[(size_t)e * st.maxv] == img) { if (lane == 0) cs.src_id[nsrc] = e; ++nsrc; }
                }
                __syncwarp();
            }
            if (lane == 0) {
                cs.nl = nl; cs.nrem = nrem; cs.nnew = 0; cs.nsrc = nsrc; cs.version = 0; cs.next_try = 0; cs.commit_ptr = 0;
                // a full cell's tries form a serial chain (each replacement moves the next try): all warps refine ONE candidate
                cs.coop = (sa.coop && wpc > 1 && (sa.coop == 2 || nl >= maxp)) ? 1 : 0;
                cs.cmd = 0;
            }
        }
        cell_sync();
        const int ntries = 2 * cs.nsrc;                                                    // MAX_NUM_OF_PROPAG tries per call
        const bool coop = cs.coop != 0;
        if (coop && warp == 0) stat[SS_COOP_CELLS] += 1;
        if (coop && warp != 0) {
            // helper warps of a cooperative cell: join every refinement the cell's first warp announces
            while (true) {
                cell_sync();
                if (cs.cmd == 2) break;
                V4 hX{cs.rs_x[0], cs.rs_x[1], cs.rs_x[2], cs.rs_x[3]}, hN{cs.rs_n[0], cs.rs_n[1], cs.rs_n[2], cs.rs_n[3]};
                const WarpScratch& mws = reinterpret_cast<const WarpScratch*>(smem_raw)[cs.rs_warp];
                coop_refine<WS>(cp, cs, mws.images, nullptr, hX, hN, cs.rs_nv, cs.rs_dscale, cs.rs_stream, warp, wpc, lane, cell_sync);
            }
        }
        // ================= the propagatePatch tries (propagate.cpp:122-218), speculative, committed in order =================
        while (!(coop && warp != 0)) {
            int t = 0;
            if (lane == 0) t = atomicAdd(&cs.next_try, 1);
            t = __shfl_sync(0xffffffffu, t, 0);
            if (t >= ntries) break;
            const int call = t >> 1, k = t & 1;
            const int src = cs.src_id[call];
            const V4 sX = f4v(st.coord[src]), sN = f4v(st.normal[src]);
            const int sref = st.images[(size_t)src * st.maxv];
            const int snv = min(st.nimg[src], CAND_MAXV);
            // one PMR1 stream per call: (iter, view, dest cell, call ordinal)
            const uint64_t stream = ((uint64_t)(unsigned)sa.iter << 56) ^ ((uint64_t)(unsigned)img << 40) ^ ((uint64_t)(unsigned)(y * gw + x) << 8) ^ (uint64_t)call;
            Cand cd;
            cd.nv = 0; cd.nvv = 0; cd.ncc = 0.f; cd.dscale = 0.f; cd.ascale = 0.f; cd.tmp = 0.f;
            cd.X = V4{0.f, 0.f, 0.f, 1.f}; cd.N = sN;
            int outcome = TRY_GEN_NULL;
            bool have_turn = false;
            TrySnap sn;
            unsigned long long ph_t = 0;
            // profiling only: close the phase that ends here (lane 0 of the warp, one global atomic)
#define PMK_PHASE(slot)                                                                                        \
            if (sa.phase_ns != nullptr && lane == 0) {                                                         \
                unsigned long long now_;                                                                       \
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));                                       \
                if ((slot) >= 0) atomicAdd(sa.phase_ns + (slot), now_ - ph_t);                                 \
                ph_t = now_;                                                                                   \
            }
            PMK_PHASE(-1)
            if (sa.phase_ns != nullptr && lane == 0) atomicAdd(sa.phase_ns + 6, 1ull);
            for (int attempt = 0; attempt < 3; ++attempt) {
                sn = take_snapshot(cs, ss.l_snap, maxp, lane);
                // ---------------- stage A: everything up to postProcess's store-independent part ----------------
                {
                    V3 ic;
                    if (sn.np < maxp) {
                        float jx = sa.jitter[2 * k], jy = sa.jitter[2 * k + 1];
                        if (sa.jitter_mode == 1) {
                            uint32_t ctr[4] = {(uint32_t)stream, (uint32_t)(stream >> 32), 0x4a495454u, (uint32_t)k};
                            philox4x32_10((uint32_t)cp.seed, (uint32_t)(cp.seed >> 32), ctr);
                            jx = (float)(0.5 * uniform_pm1(ctr[0])); jy = (float)(0.5 * uniform_pm1(ctr[1]));
                        }
                        const float cxf = (float)(p.csize * (2 * x + 1) - 1) / 2.0f, cyf = (float)(p.csize * (2 * y + 1) - 1) / 2.0f;
                        ic = V3{xadd(cxf, xmul(jx, (float)p.csize)), xadd(cyf, xmul(jy, (float)p.csize)), xadd(1.0f, 0.0f)};
                    } else ic = project(vimgc.P, f4v(st.coord[sn.w]));
                    // ---- generatePatch (propagate.cpp:220-237) ----
                    const ViewConst& vr = p.views[sref];
                    const float depth = dot4(ld4(vr.oaxis), sX);
                    const float b0 = xsub(xmul(depth, ic.x), vr.P[3]), b1 = xsub(xmul(depth, ic.y), vr.P[7]), b2 = xsub(xmul(depth, ic.z), vr.P[11]);
                    V4 X, N = sN;                                                              // Camera::unproject (camera.cpp:329-337)
                    X.x = xadd(xadd(xmul(vr.Minv[0], b0), xmul(vr.Minv[1], b1)), xmul(vr.Minv[2], b2));
                    X.y = xadd(xadd(xmul(vr.Minv[4], b0), xmul(vr.Minv[5], b1)), xmul(vr.Minv[6], b2));
                    X.z = xadd(xadd(xmul(vr.Minv[8], b0), xmul(vr.Minv[9], b1)), xmul(vr.Minv[10], b2));
                    X.w = 1.0f;
                    int nv = 0;                                                                // setGridsImages (patch_manager.cpp:223-239)
                    for (int base = 0; base < snv; base += 32) {
                        const int i = base + lane;
                        bool keep = false;
                        int v = 0;
                        if (i < snv) {
                            v = st.images[(size_t)src * st.maxv + i];
                            const V3 q = project(p.views[v].P, X);
                            const int ix = cell_of(q.x, p.csize), iy = cell_of(q.y, p.csize);
                            keep = 0 <= ix && ix < p.views[v].gw && 0 <= iy && iy < p.views[v].gh;
                        }
                        const unsigned msk = __ballot_sync(0xffffffffu, keep);
                        if (keep) ws.images[nv + __popc(msk & ((1u << lane) - 1u))] = v;
                        nv += __popc(msk);
                    }
                    __syncwarp();
                    outcome = TRY_GEN_NULL;
                    if (nv > 0) {
                        float ncc = warp_compute_ncc<WS>(p, ws, X, N, nv, lane);
                        ncc = __shfl_sync(0xffffffffu, ncc, 0);
                        PMK_PHASE(0)
                        if (sn.np >= maxp && ncc < sn.wncc) outcome = TRY_LOSE;
                        else {
                            // ---- patch optimisation (propagate.cpp:176-193) ----
                            float dscale, ascale;
                            const int pre = warp_pre_process<WS>(cp, ws, X, N, nv, dscale, ascale, lane);
                            PMK_PHASE(1)
                            if (pre == -1) outcome = TRY_FAIL0;
                            else {
                                if (sa.phase_ns != nullptr && lane == 0) atomicAdd(sa.phase_ns + 7, 1ull);
                                if (coop) {
                                    if (lane == 0) {
                                        cs.rs_x[0] = X.x; cs.rs_x[1] = X.y; cs.rs_x[2] = X.z; cs.rs_x[3] = X.w;
                                        cs.rs_n[0] = N.x; cs.rs_n[1] = N.y; cs.rs_n[2] = N.z; cs.rs_n[3] = N.w;
                                        cs.rs_dscale = dscale; cs.rs_nv = nv; cs.rs_warp = cta_warp; cs.rs_stream = stream; cs.cmd = 1;
                                    }
                                    cell_sync();
                                    stat[SS_COOP_REFINES] += 1;
                                    ncc = coop_refine<WS>(cp, cs, ws.images, ws.units, X, N, nv, dscale, stream, 0, wpc, lane, cell_sync);
                                } else
                                ncc = warp_refine<WS>(cp, ws, X, N, nv, dscale, stream, nullptr, lane);
                                ncc = __shfl_sync(0xffffffffu, ncc, 0);
                                X = V4{__shfl_sync(0xffffffffu, X.x, 0), __shfl_sync(0xffffffffu, X.y, 0), __shfl_sync(0xffffffffu, X.z, 0), __shfl_sync(0xffffffffu, X.w, 0)};
                                N = V4{__shfl_sync(0xffffffffu, N.x, 0), __shfl_sync(0xffffffffu, N.y, 0), __shfl_sync(0xffffffffu, N.z, 0), 0.0f};
                                PMK_PHASE(2)
                                const int r = warp_post_process<WS>(cp, ws, X, N, nv, gwarp, lane);
                                PMK_PHASE(3)
                                outcome = r == 0 ? TRY_ACCEPT : TRY_FAIL1;
                                cd.X = X; cd.N = N; cd.nv = nv; cd.ncc = ncc; cd.dscale = dscale; cd.ascale = ascale;
                                if (r == 0) {
                                    bool outside = false;
                                    for (int i = lane; i < nv; i += 32) {                      // setGrids (optim.cpp:285)
                                        const V3 q = project(p.views[ws.images[i]].P, X);
                                        const int ix = cell_of(q.x, p.csize), iy = cell_of(q.y, p.csize);
                                        outside |= ix < 0 || p.views[ws.images[i]].gw <= ix || iy < 0 || p.views[ws.images[i]].gh <= iy;
                                        ss.cells[i] = pack_cell(ix, iy);
                                    }
                                    __syncwarp();
                                    // Views inherited from the source patch are not re-checked by addImages, and the refinement moves
                                    // the patch: a view can end up seeing it outside its grid.  The reference then writes m_pgrids out of
                                    // bounds (addPatch, patch_manager.cpp:164-170); here the candidate is rejected like any postProcess failure.
                                    if (__any_sync(0xffffffffu, outside)) outcome = TRY_FAIL1;
                                }
                            }
                        }
                    }
                }
                bool post_ok = outcome == TRY_ACCEPT;
                // ---------------- stage B (needs the list): setVImagesVGrids, check (optim.cpp:288-296) ----------------
                bool stage_b_done = false;
                while (true) {
                    if (post_ok && !stage_b_done) {
                        outcome = TRY_ACCEPT;
                        cd.tmp = xmul(max_std(0.0f, xsub(cd.ncc, p.ncc_threshold)), (float)cd.nv);      // m_tmp = score2
                        cd.nvv = 0;
                        if (p.depth) cd.nvv = warp_set_vimages(sp, ws, cd.X, cd.N, ws.images, cd.nv, ss.vimg, ss.vcell, 0, lane);
                        if (2 <= p.depth) {                                                    // Optim::check (optim.cpp:300-323)
                            PGeo me; me.X = cd.X; me.N = cd.N; me.dscale = cd.dscale; me.ref = ws.images[0];
                            const PatchLists pl{ws.images, ss.cells, cd.nv, ss.vimg, ss.vcell, cd.nvv};
                            const Overlay ov{cD, ss.l_snap, min(sn.np, LKEEP), cs.removed, sn.nrem};
                            const float gain = warp_compute_gain(sp, me, cd.ncc, pl, ov, lane);
                            cd.tmp = gain;
                            if (gain < 0.0f) outcome = TRY_FAIL1;
                            else {
                                int* nb = sp.nb_scratch + (size_t)gwarp * NB_STRIDE;
                                const int nn = warp_find_neighbors(sp, me, pl, 4.0f, 2, ov, nb, lane);
                                if (6 < nn && warp_filter_quad(sp, me, pl, nb, nn, nullptr, lane)) outcome = TRY_FAIL1;
                            }
                        }
                        stage_b_done = true;
                        PMK_PHASE(4)
                    }
                    if (have_turn) break;
                    // ---------------- wait for this try's turn ----------------
                    while (ld_shared_volatile(&cs.commit_ptr) != t) __nanosleep(100);
                    have_turn = true;
                    __syncwarp();
                    if (ld_shared_volatile(&cs.version) == sn.ver) break;                    // nothing committed since the snapshot
                    // the list changed: are the snapshot's assumptions still true?
                    const int np_now = ld_shared_volatile(&cs.nl);
                    const bool same_branch = (sn.np < maxp) == (np_now < maxp);
                    const bool same_worst = np_now < maxp || ld_shared_volatile(&cs.l_id[maxp - 1]) == sn.w;
                    if (!(same_branch && same_worst)) { stage_b_done = false; post_ok = false; outcome = -1; break; }   // redo stage A
                    sn = take_snapshot(cs, ss.l_snap, maxp, lane);                            // stage A stands; stage B reads the list
                    if (post_ok && p.depth >= 2) stage_b_done = false; else break;
                }
                if (outcome >= 0) break;
            }
            PMK_PHASE(5)
#undef PMK_PHASE
            // ================= commit (this warp holds the turn) =================
            stat[SS_CALLS] += (k == 0) ? 1 : 0;
            stat[SS_TRIES] += 1;
            if (outcome != TRY_GEN_NULL) stat[SS_EVALS] += 1;
            if (outcome == TRY_FAIL1 || outcome == TRY_ACCEPT) stat[SS_EVALS] += PMR1_EVALS + 1;
            if (outcome == TRY_GEN_NULL) stat[SS_GEN_NULL] += 1;
            else if (outcome == TRY_LOSE) stat[SS_NCC_LOSE] += 1;
            else if (outcome == TRY_FAIL0) stat[SS_FAIL0] += 1;
            else if (outcome == TRY_FAIL1) stat[SS_FAIL1] += 1;
            else {
                // ---- removePatch(worst) / addPatch(new) (propagate.cpp:195-207); grid updates are staged ----
                int nl = cs.nl;
                if (lane == 0) { cs.version = sn.ver + 1; }
                __threadfence_block();
                if (nl == maxp) {
                    const int w = cs.l_id[maxp - 1];
                    if (w >= st.cap) { if (lane == 0) st.state[w] = 0; }                       // staged this step: never reaches the grids
                    else if (lane == 0) { cs.removed[cs.nrem] = w; cs.nrem = cs.nrem + 1; }
                    --nl;
                    stat[SS_REPLACED] += 1;
                } else stat[SS_ADDED] += 1;
                __syncwarp();
                const int sid = st.cap + task * NEW_MAX + cs.nnew;
                __syncwarp();
                if (lane == 0) {
                    cs.nnew = cs.nnew + 1;
                    st.coord[sid] = v4f(cd.X); st.normal[sid] = v4f(cd.N);
                    st.scal[sid] = make_float4(cd.ncc, cd.dscale, cd.ascale, cd.tmp);
                    st.nimg[sid] = cd.nv; st.nvimg[sid] = cd.nvv; st.state[sid] = 1;
                    st.birth[sid] = 0xffffffffu;
                }
                bool inD = false;
                for (int i = lane; i < cd.nv; i += 32) {
                    st.images[(size_t)sid * st.maxv + i] = ws.images[i];
                    st.cells[(size_t)sid * st.maxv + i] = ss.cells[i];
                    if (ws.images[i] == img && cell_x(ss.cells[i]) == x && cell_y(ss.cells[i]) == y) inD = true;
                }
                for (int i = lane; i < cd.nvv; i += 32) {
                    st.vimages[(size_t)sid * st.maxv + i] = ss.vimg[i];
                    st.vcells[(size_t)sid * st.maxv + i] = ss.vcell[i];
                }
                inD = __any_sync(0xffffffffu, inD);
                if (lane == 0) {
                    if (inD) { cs.l_id[nl] = sid; cs.l_ncc[nl] = cd.ncc; ++nl; swap_sort_desc(cs.l_id, cs.l_ncc, nl); }
                    cs.nl = nl;
                    __threadfence_block();
                    cs.version = sn.ver + 2;
                }
                __syncwarp();
            }
            __threadfence_block();
            if (lane == 0) cs.commit_ptr = t + 1;
            __syncwarp();
        }
        if (coop && warp == 0) { if (lane == 0) cs.cmd = 2; cell_sync(); }
        cell_sync();
        // ---- hand the step's mutations to k4_apply ----
        if (warp == 0) {
            const int nrem = cs.nrem;
            if (lane == 0) sa.task_new[task] = cs.nnew;
            if (nrem > 0) {
                int base = 0;
                if (lane == 0) base = atomicAdd(st.counters + SC_REM, nrem);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (int i = lane; i < nrem; i += 32) sa.rem_list[base + i] = cs.removed[i];
            }
            if (lane == 0) {
                unsigned long long t_end;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
                atomicAdd(sa.stats + SS_CELL_NS, t_end - t_begin);
                atomicMax(sa.step_max, t_end - t_begin);
                if (sa.cell_ns) sa.cell_ns[cD] = (float)(t_end - t_begin);
            }
        }
        cell_sync();
    }
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < SS_COUNT; ++i) if (stat[i]) atomicAdd(sa.stats + i, (unsigned long long)stat[i]);
}

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
            int g = 0;
            while (g + 1 < sa.ngroup && task >= sa.g_off[g + 1]) ++g;
            const int img = sa.g_img[g];
            const int gw = p.views[img].gw, gh = p.views[img].gh;
            const int x = sa.g_xlo[g] + (task - sa.g_off[g]), y = sa.g_diag[g] - x;
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
        int acc = 0, heavy = 0;
        for (int b = NBIN - 1; b >= 0; --b) { offs[b] = acc; acc += hist[b]; if (b >= sa.heavy_est) heavy = acc; }
        order[sa.heavy_slot] = heavy;
    }
    __syncthreads();
    for (int base = 0; base < sa.ntasks; base += blockDim.x) {
        const int task = base + threadIdx.x;
        if (task >= sa.ntasks) continue;
        // recompute the estimate (cheaper than keeping it: tasks can exceed the block size)
        int g = 0;
        while (g + 1 < sa.ngroup && task >= sa.g_off[g + 1]) ++g;
        const int img = sa.g_img[g];
        const int gw = p.views[img].gw, gh = p.views[img].gh;
        const int x = sa.g_xlo[g] + (task - sa.g_off[g]), y = sa.g_diag[g] - x;
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
__global__ void __launch_bounds__(1024) k4_apply_scan(const StoreParams sp, const int* __restrict__ task_new, int ntasks, int* __restrict__ final_id) {
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
    }
}

// copy every staged, surviving patch to its final slot and register it (one warp per staged entry)
__global__ void k4_apply_add(const StoreParams sp, const int* __restrict__ task_new, int ntasks, const int* __restrict__ final_id) {
    const StoreDev& st = sp.st;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int w = gwarp; w < ntasks * NEW_MAX; w += nwarps) {
        const int t = w / NEW_MAX, j = w % NEW_MAX;
        if (j >= task_new[t]) continue;
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
// Multi-GPU: the store is replicated, the dest cells of a step are partitioned by row band, and the step's mutations
// (new patches, removals) travel between the ranks as one fixed-layout message per rank (ncclAllGather over NVLink).
// Every rank then applies ALL messages in rank order, so ids, creation numbers and grids stay identical everywhere.
//   message = int hdr[4] {n_new, n_rem, overflow, 0} | int rem[rem_cap] | rec[rec_cap][16 + 4 * maxv]
//   record  = coord4, normal4, scal4 (as int bits), nimg, nvimg, global task index, slot in the task,
//             images[maxv], cells[maxv], vimages[maxv], vcells[maxv]
// Final ids and creation numbers are assigned in (global task, slot) order -- the order the single-GPU scan uses -- so an
// N-GPU run produces the single-GPU store bit for bit.
// =====================================================================================================================
namespace pmk {

struct MsgLayout {
    int rem_cap, rec_cap, rec_words;
    __host__ __device__ size_t words() const { return 4 + (size_t)rem_cap + (size_t)rec_cap * rec_words; }
};

// one thread: list the staged patches that survived the step, fill the header
__global__ void k4_pack_scan(const StoreParams sp, const int* __restrict__ task_new, int ntasks, const int* __restrict__ rem_list, MsgLayout ml,
                             int* __restrict__ msg, int* __restrict__ pack_ids) {
    const StoreDev& st = sp.st;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int n = 0, over = 0;
    for (int t = 0; t < ntasks; ++t) {
        const int k = task_new[t];
        for (int j = 0; j < k; ++j) {
            const int sid = st.cap + t * NEW_MAX + j;
            if (st.state[sid] != 1) continue;
            if (n < ml.rec_cap) pack_ids[n++] = sid; else over = 1;
        }
    }
    int nrem = st.counters[SC_REM];
    if (nrem > ml.rem_cap) { nrem = ml.rem_cap; over = 1; }
    msg[0] = n; msg[1] = nrem; msg[2] = over; msg[3] = 0;
    for (int i = 0; i < nrem; ++i) msg[4 + i] = rem_list[i];
}

__global__ void k4_pack_copy(const StoreParams sp, const SweepArgs sa, MsgLayout ml, int* __restrict__ msg, const int* __restrict__ pack_ids) {
    const StoreDev& st = sp.st;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const int n = msg[0];
    for (int r = gwarp; r < n; r += nwarps) {
        const int sid = pack_ids[r];
        int* rec = msg + 4 + ml.rem_cap + (size_t)r * ml.rec_words;
        if (lane == 0) {
            const float4 c = st.coord[sid], m = st.normal[sid], s = st.scal[sid];
            rec[0] = __float_as_int(c.x); rec[1] = __float_as_int(c.y); rec[2] = __float_as_int(c.z); rec[3] = __float_as_int(c.w);
            rec[4] = __float_as_int(m.x); rec[5] = __float_as_int(m.y); rec[6] = __float_as_int(m.z); rec[7] = __float_as_int(m.w);
            rec[8] = __float_as_int(s.x); rec[9] = __float_as_int(s.y); rec[10] = __float_as_int(s.z); rec[11] = __float_as_int(s.w);
            rec[12] = st.nimg[sid]; rec[13] = st.nvimg[sid];
            const int task = (sid - st.cap) / NEW_MAX;
            int g = 0;
            while (g + 1 < sa.ngroup && task >= sa.g_off[g + 1]) ++g;
            rec[14] = sa.g_goff[g] + (sa.g_xlo[g] + (task - sa.g_off[g]) - sa.g_gxlo[g]);
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
        const int* msg = all + (size_t)rk * ml.words();
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

// sort keys of every rank's records: (global task, slot); unused tail = ~0.  One thread per record slot.
__global__ void k4_unpack_keys(MsgLayout ml, const int* __restrict__ all, int nranks, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nranks * ml.rec_cap) return;
    const int rk = i / ml.rec_cap, r = i % ml.rec_cap;
    const int* msg = all + (size_t)rk * ml.words();
    unsigned long long k = ~0ull;
    if (r < msg[0]) {
        const int* rec = msg + 4 + ml.rem_cap + (size_t)r * ml.rec_words;
        k = (unsigned long long)(unsigned int)rec[14] * NEW_MAX + (unsigned int)rec[15];
    }
    keys[i] = k;
    vals[i] = i;
}

// one thread: how many records there are in all, capacity check, counters
__global__ void k4_unpack_scan(const StoreParams sp, MsgLayout ml, const int* __restrict__ all, int nranks, int* __restrict__ rec_base) {
    const StoreDev& st = sp.st;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int total = 0;
    for (int rk = 0; rk < nranks; ++rk) {
        const int* msg = all + (size_t)rk * ml.words();
        if (msg[2]) atomicAdd(st.counters + SC_OVERFLOW, 1);
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
__global__ void k4_unpack_add(const StoreParams sp, MsgLayout ml, const int* __restrict__ all, int nranks, const int* __restrict__ rec_base,
                              const int* __restrict__ sorted_vals) {
    const StoreDev& st = sp.st;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const bool deep = sp.cp.p.depth != 0;
    const int n0 = rec_base[0], b0 = rec_base[1], take = rec_base[2];
    for (int pos = gwarp; pos < take; pos += nwarps) {
        const int i = sorted_vals[pos];
        const int rk = i / ml.rec_cap, r = i % ml.rec_cap;
        const int* rec = all + (size_t)rk * ml.words() + 4 + ml.rem_cap + (size_t)r * ml.rec_words;
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
