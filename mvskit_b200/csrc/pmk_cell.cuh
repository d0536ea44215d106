// mvskit_b200/csrc/pmk_cell.cuh -- K4 "PMS1c": the wavefront sweep of pmk_sweep.cuh with ONE CTA PER DEST CELL.
//
// Same schedule, same arithmetic and therefore the same store as the one-warp-per-cell kernel of round 1 (the tests compare
// the checksums), but every propagatePatch try (pmmvps/propagate.cpp:123-218) is spread over all evaluator groups of a CTA
// instead of being walked by one warp:
//   * a texture grab (Optim::getTex + normalize, optim.cpp:790-844,917-940) is still the work of one 8-lane group, but it lands
//     in a SHARED-MEMORY slot (centred lattice, column-major per lane) instead of registers, so any group can pair any two;
//   * Optim::setINCCs (optim.cpp:708-746) -> one group per view, all views of a candidate at once;
//   * Optim::refinePatch / PMR1 (optim.cpp:470-547) -> the 8 candidates x tau views of a level are 48 independent grabs, then
//     40 dots from shared memory, then every thread sums the per-view terms in view order (Optim::cost_func, :401-468): a
//     level costs one grab latency instead of twelve;
//   * Optim::setRefImage (optim.cpp:348-383) -> grabs one group per view, the pairs spread over all threads;
//   * Optim::check (optim.cpp:300-323) -> computeGain one warp per registration, findNeighbors one warp per view with a
//     CTA-wide hash table, filterQuad on the first warp.
// The list logic (addImages, constraintImages, sortImages, ...) stays on warp 0 and is tiny.  The kernel needs <= 80 registers
// (the monolith needed 254), runs 3 CTAs x 8 warps per SM, and a dest cell's chain of tries -- which is what bounds a wavefront
// step -- is about half as long (measured on config 2: Propagate::run 6.5 s -> 3.2-3.5 s; DESIGN.md section 4 has the phase times).
#pragma once

#include "pmk_sweep.cuh"

namespace pmk {

#ifndef PMK_CELL_MINB
#define PMK_CELL_MINB 3
#endif
constexpr int EVAL_SLOTS = PMR1_CANDS * PMK_MAX_TAU;      // (candidate, view) items of one PMR1 level
constexpr int FRAME_SLOTS = EVAL_SLOTS > CAND_MAXV ? EVAL_SLOTS : CAND_MAXV;

template <int WS, int NW>
struct CellGeom {
    static constexpr int GW = WS <= 8 ? 8 : 16;              // lanes of an evaluator group
    static constexpr int G = 32 / GW;                         // groups per warp
    static constexpr int TG = NW * G;                         // groups per CTA
    static constexpr int NT = NW * 32;                        // threads per CTA
    static constexpr int SLOT = WS * 3 * WS;                  // floats of one texture slot: [row][channel][column < WS]
    __host__ __device__ static constexpr int nslots(int tau) { return (PMR1_CANDS * tau > TG + 1) ? PMR1_CANDS * tau : TG + 1; }
};

#ifdef PMK_SUBPHASE
#define PMK_SUBT(slot) do { if ((int)threadIdx.x == 0) { unsigned long long n_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(n_)); if ((slot) >= 0) cs.sub[slot] += n_ - cs.sub_t; cs.sub_t = n_; } } while (0)
#else
#define PMK_SUBT(slot) do { } while (0)
#endif

struct CandGeo { V4 X, N, px, py; };                     // a candidate's decoded point and its patch axes (Optim::getPAxes)
struct ItemFrame { float tlx, tly, dxx, dxy, dyx, dyy; int level, view; };    // what Optim::getTex decides before it samples

struct CellCta {                       // one per CTA, shared memory
    int l_id[LKEEP];                   // the dest cell's m_pgrids list, sorted (sortPatches)
    float l_ncc[LKEEP];
    unsigned int l_birth[LKEEP];
    int src_id[SRC_MAX];
    int removed[REM_OVERLAY + NEW_MAX];
    int nl, nrem, nnew, nsrc;
    int bi[8];                         // small broadcasts from warp 0
    float bf[8];
    RefineCtx rc;                      // Optim::m_center / m_ray / m_dscale of the refinement in flight
    int fr_ok[PMK_MAX_TAU];            // view index valid
    ViewConst vc[PMK_MAX_TAU];         // the refinement's views (m_indexes), copied next to the SM: every cost evaluation starts from them
    CandGeo cg[PMR1_CANDS];
    ItemFrame fr[FRAME_SLOTS];         // per item: the sampling frame (level < 0: getTex == -1)
    float inv[FRAME_SLOTS];            //           1 / msd of Optim::normalize
    float val[EVAL_SLOTS];             //           robustincc(1 - dot(reference, view))
    int wl[FRAME_SLOTS];               // setINCCs: work list of views whose frame is valid
    int nwl;
    float mp[2 * CAND_MAXV];           // computeGain: per registration, the strongest non-neighbour of its cell
    int negs[32];                      // preamble: list entries whose m_ncc has to be recomputed
    int nneg;
    int nb_count, nb_over;             // findNeighbors
    unsigned int nb_token;
    int task;
    unsigned int stat[SS_COUNT];
    unsigned long long t_begin, ph_t;
    unsigned long long sub_t, sub[8];  // PMK_SUBPHASE builds: time between the barriers of cta_costs (profiling)
};

// ---- the sampling half of a texture grab, by one evaluator group, into a shared-memory slot -------------------------------------
// Optim::getTex's 49 samples + Optim::normalize (optim.cpp:835-842, 917-940) for a frame make_frame accepted.  Arithmetic identical
// to group_grab (pmk_cand.cuh); the centred lattice column of each lane (col < WS) goes to slot[(row * 3 + channel) * WS + col].
// Returns 1 / sqrt(ssd / (3 n)) (1 when ssd == 0).
template <int WS, int GW>
__device__ __noinline__ float sample_slot(const ViewConst& vc, const ItemFrame& f, int col, unsigned gm, float* __restrict__ slot) {
    constexpr int NSAMP = WS * WS;
    constexpr float INV_NSAMP = 1.0f / (float)NSAMP, INV_3NSAMP = 1.0f / (float)(3 * NSAMP);
    const float cmask = col < WS ? 1.0f : 0.0f;
    const Texel* img = vc.img[f.level];
    const int W = vc.w[f.level];
    const float fcol = (float)(col < WS ? col : WS - 1);
    const float dyx = f.dyx, dyy = f.dyy;
    const float bx = fmaf(f.dxx, fcol, f.tlx), by = fmaf(f.dxy, fcol, f.tly);
    float t[WS][3];
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int y = 0; y < WS; ++y) {
        bilinear(img, W, fmaf(dyx, (float)y, bx), fmaf(dyy, (float)y, by), t[y][0], t[y][1], t[y][2]);
        s0 = fmaf(t[y][0], cmask, s0); s1 = fmaf(t[y][1], cmask, s1); s2 = fmaf(t[y][2], cmask, s2);
    }
    const float m0 = -group_sum<GW>(s0, gm) * INV_NSAMP * cmask, m1 = -group_sum<GW>(s1, gm) * INV_NSAMP * cmask, m2 = -group_sum<GW>(s2, gm) * INV_NSAMP * cmask;
    float ssd = 0.f;
#pragma unroll
    for (int y = 0; y < WS; ++y) {
        t[y][0] = fmaf(t[y][0], cmask, m0); t[y][1] = fmaf(t[y][1], cmask, m1); t[y][2] = fmaf(t[y][2], cmask, m2);
        ssd = fmaf(t[y][0], t[y][0], fmaf(t[y][1], t[y][1], fmaf(t[y][2], t[y][2], ssd)));
        if (col < WS) { slot[(y * 3 + 0) * WS + col] = t[y][0]; slot[(y * 3 + 1) * WS + col] = t[y][1]; slot[(y * 3 + 2) * WS + col] = t[y][2]; }
    }
    const float var = group_sum<GW>(ssd, gm) * INV_3NSAMP;
    return var > 0.0f ? rsqrtf(var) : 1.0f;
}

// the deciding half (optim.cpp:790-833 + getTexSafe :895-915), by ONE thread per (candidate, view) item
__device__ __noinline__ void frame_item(const Params& p, const ViewConst* vc, int view, const CandGeo& cg, ItemFrame* out) {
    ItemFrame fr;
    fr.level = -1; fr.view = 0;
    fr.tlx = fr.tly = fr.dxx = fr.dxy = fr.dyx = fr.dyy = 0.0f;
    if (vc != nullptr) {
        const Frame f = make_frame(p, *vc, cg.X, cg.N, cg.px, cg.py);
        fr.level = f.level; fr.view = view;
        fr.tlx = f.tlx; fr.tly = f.tly; fr.dxx = f.dxx; fr.dxy = f.dxy; fr.dyx = f.dyx; fr.dyy = f.dyy;
    }
    *out = fr;
}
__device__ __forceinline__ const ViewConst* view_or_null(const Params& p, int view) { return (view >= 0 && view < p.nviews) ? p.views + view : nullptr; }

// Optim::dot (optim.cpp:601-609) of two slots; same order of operations as group_dot(a = reference, b = view)
template <int WS, int GW>
__device__ __forceinline__ float dot_slots(const float* __restrict__ a, float inv_a, const float* __restrict__ b, float inv_b, int col, unsigned gm) {
    // lanes past the lattice width hold zeros (group_grab masks them): they add nothing but take part in the reduction tree
    float dp = 0.f;
    if (col < WS) {
#pragma unroll
        for (int y = 0; y < WS; ++y)
            dp = fmaf(a[(y * 3 + 0) * WS + col], b[(y * 3 + 0) * WS + col],
                      fmaf(a[(y * 3 + 1) * WS + col], b[(y * 3 + 1) * WS + col], fmaf(a[(y * 3 + 2) * WS + col], b[(y * 3 + 2) * WS + col], dp)));
    }
    return group_sum<GW>(dp, gm) * inv_a * inv_b * (1.0f / (float)(3 * WS * WS));
}

// per-thread coordinates inside the CTA
template <int WS, int NW>
struct CellLane {
    int tid, warp, lane, col, g;
    unsigned gm;
    __device__ __forceinline__ CellLane() {
        constexpr int GW = CellGeom<WS, NW>::GW;
        tid = (int)threadIdx.x; warp = tid >> 5; lane = tid & 31; col = lane % GW;
        g = warp * CellGeom<WS, NW>::G + lane / GW;
        gm = group_mask<GW>(lane);
    }
};

// ---- Optim::setINCCs 1-vs-all (optim.cpp:708-746) by the CTA ---------------------------------------------------------------------
// One THREAD per view decides the frame (projections, pyramid level, getTexSafe), one GROUP per accepted view samples, then the
// dots against the reference slot.  images / inccs live in shared memory; every thread holds the same X, N, n.  Ends with a barrier.
template <int WS, int NW>
__device__ __noinline__ void cta_set_inccs(const Params& p, CellCta& cs, float* tex, V4 X, V4 N, const int* images, int n, int robust, float* inccs) {
    typedef CellGeom<WS, NW> Gm;
    const CellLane<WS, NW> L;
    if (L.tid == 0) cs.nwl = 1;                                   // work list entry 0 = the reference view
    if (L.tid < ((n + 31) & ~31)) {                              // the warps that hold a view compute the patch axes (no barrier needed)
        CandGeo cg;
        cg.X = X; cg.N = N;
        get_paxes(p.views[images[0]], X, N, p.level_scale, cg.px, cg.py);
        if (L.tid < n) frame_item(p, view_or_null(p, images[L.tid]), images[L.tid], cg, &cs.fr[L.tid]);
    }
    __syncthreads();
    if (cs.fr[0].level < 0) {
        for (int k = L.tid; k < n; k += Gm::NT) inccs[k] = 2.0f;
        __syncthreads();
        return;
    }
    for (int i = L.tid; i < n; i += Gm::NT) {
        if (i == 0) { inccs[0] = 0.0f; cs.wl[0] = 0; }
        else if (cs.fr[i].level < 0) inccs[i] = 2.0f;
        else cs.wl[atomicAdd(&cs.nwl, 1)] = i;
    }
    __syncthreads();
    const int nw = cs.nwl;
#pragma unroll 1
    for (int base = 0; base < nw; base += Gm::TG) {
        const int k = base + L.g;
        float* mine = tex + (size_t)(k == 0 ? 0 : 1 + L.g) * Gm::SLOT;       // slot 0: the reference view; slot 1 + g: this group's view
        float inv = 1.0f;
        if (k < nw) {
            inv = sample_slot<WS, Gm::GW>(p.views[cs.fr[cs.wl[k]].view], cs.fr[cs.wl[k]], L.col, L.gm, mine);
            if (k == 0 && L.col == 0) cs.inv[0] = inv;
        }
        __syncthreads();
        if (k >= 1 && k < nw) {
            const float d = dot_slots<WS, Gm::GW>(tex, cs.inv[0], mine, inv, L.col, L.gm);
            float r = xsub(1.0f, d);
            if (robust) r = robustincc(r);
            if (L.col == 0) inccs[cs.wl[k]] = r;
        }
        if (base + Gm::TG < nw) __syncthreads();                  // the next round overwrites the slots
    }
    __syncthreads();
}

// PatchManager::computeNcc (patch_manager.cpp:401-404) on {X, N, ws.images[0..nv)}: every thread returns m_ncc
template <int WS, int NW>
__device__ __noinline__ float cta_compute_ncc(const Params& p, WarpScratch& ws, CellCta& cs, float* tex, V4 X, V4 N, int nv) {
    const int tid = (int)threadIdx.x;
    if (tid < 32) compute_weights(p, X, N, ws.images, nv, ws.units, tid);
    float incc = 2.0f;
    if (nv >= 2) {
        const int sz = min(p.tau, nv);
        cta_set_inccs<WS, NW>(p, cs, tex, X, N, ws.images, sz, 1, ws.inccs);
        float score = 0.0f, tw = 0.0f;
        for (int i = 1; i < sz; ++i) {
            const float v = ws.inccs[i];
            if (v != 2.0f) { tw = xadd(tw, ws.units[i]); score = xadd(score, xmul(v, ws.units[i])); }
        }
        incc = (tw == 0.0f) ? 2.0f : xdiv(score, tw);
    }
    __syncthreads();
    return xsub(1.0f, unrobustincc(incc));
}

// ---- Optim::cost_func (optim.cpp:401-468) for `ncand` encoded points at once -------------------------------------------------------
// (1) one thread per candidate: its point, decode + getPAxes; (2) one thread per (candidate c, view i) item: the frame; (3) one group
// per accepted item: sample into slot c * sz + i; (4) the dots against the candidate's reference slot.  The points: level < 0 -> the
// single point `best`; else candidate c of PMR1 level `level` around `best` with radii `r` (pmr1_point).  best / r only need to be
// valid in warp 0.  On return (after a CTA barrier) cs.fr[].level / cs.val hold what eval_cost needs.  One copy, out of line: the
// sweep kernel is instruction-cache bound.
__device__ __forceinline__ void pmr1_point(const CandParams& cp, uint64_t stream, int level, int cnd, const double best[3], const double r[3], double xc[3]) {
    const double lb[3] = {-(double)__int_as_float(0x7f800000), -23.99999, -23.99999};
    const double ub[3] = {(double)__int_as_float(0x7f800000), 23.99999, 23.99999};
    uint32_t ctr[4] = {(uint32_t)stream, (uint32_t)(stream >> 32), (uint32_t)level, (uint32_t)cnd};
    philox4x32_10((uint32_t)cp.seed, (uint32_t)(cp.seed >> 32), ctr);
#pragma unroll
    for (int i = 0; i < 3; ++i) xc[i] = fmax(fmin(__dadd_rn(best[i], __dmul_rn(r[i], uniform_pm1(ctr[i]))), ub[i]), lb[i]);
}

template <int WS, int NW>
__device__ __noinline__ void cta_costs(const CandParams& cp, CellCta& cs, float* tex, const int* images, int sz, int ncand, uint64_t stream, int level,
                                       double b0, double b1, double b2, double r0, double r1, double r2) {
    typedef CellGeom<WS, NW> Gm;
    const Params& p = cp.p;
    const CellLane<WS, NW> L;
    const int nitems = ncand * sz;
    PMK_SUBT(-1);
    if (L.tid < ncand) {
        const double best[3] = {b0, b1, b2}, r[3] = {r0, r1, r2};
        double xc[3] = {b0, b1, b2};
        if (level >= 0) pmr1_point(cp, stream, level, L.tid, best, r, xc);
        CandGeo cg;
        decode(cp, cs.vc[0], cs.rc, xc, cg.X, cg.N);
        get_paxes(cs.vc[0], cg.X, cg.N, p.level_scale, cg.px, cg.py);
        cs.cg[L.tid] = cg;
    }
    __syncthreads();
    PMK_SUBT(0);
    for (int item = L.tid; item < nitems; item += Gm::NT) {
        const int c = item / sz, i = item - c * sz;
        frame_item(p, cs.fr_ok[i] ? &cs.vc[i] : nullptr, i, cs.cg[c], &cs.fr[item]);              // ItemFrame::view = index into cs.vc
    }
    __syncthreads();
    PMK_SUBT(1);
#pragma unroll 1
    for (int item = L.g; item < nitems; item += Gm::TG) {
        if (cs.fr[item].level < 0) continue;
        const float inv = sample_slot<WS, Gm::GW>(cs.vc[cs.fr[item].view], cs.fr[item], L.col, L.gm, tex + (size_t)item * Gm::SLOT);
        if (L.col == 0) cs.inv[item] = inv;
    }
    __syncthreads();
    PMK_SUBT(2);
#pragma unroll 1
    for (int item = L.g; item < nitems; item += Gm::TG) {
        const int c = item / sz, i = item - c * sz;
        if (i == 0 || cs.fr[c * sz].level < 0 || cs.fr[item].level < 0) continue;
        const float d = dot_slots<WS, Gm::GW>(tex + (size_t)(c * sz) * Gm::SLOT, cs.inv[c * sz], tex + (size_t)item * Gm::SLOT, cs.inv[item], L.col, L.gm);
        if (L.col == 0) cs.val[item] = robustincc(__double2float_rn(1.0 - (double)d));
    }
    __syncthreads();
    PMK_SUBT(3);
}
// the cost of candidate c from the per-view terms, summed in view order as cost_func does
__device__ __forceinline__ double eval_cost(const CellCta& cs, int c, int sz, int minimum) {
    if (cs.fr[c * sz].level < 0) return 2.0;
    double ans = 0.0;
    int denom = 0;
    for (int i = 1; i < sz; ++i) if (cs.fr[c * sz + i].level >= 0) { ans += (double)cs.val[c * sz + i]; ++denom; }
    if (denom < minimum - 1) return 2.0;
    return ans / (double)denom;
}

// ---- Optim::refinePatch (optim.cpp:470-547), schedule PMR1, by the CTA ------------------------------------------------------------------
// Same points, same costs, same argmin as warp_refine (pmk_cand.cuh).  The search state (best point, its cost, the radii) lives in
// warp 0 only: its first lanes decode the level's candidates (cta_costs step 1) and it alone reads the costs back; the other warps
// only supply frames, samples and dots.  Every thread returns the same X, N and m_ncc (broadcast through shared memory).
template <int WS, int NW>
__device__ __forceinline__ float cta_refine(const CandParams& cp, WarpScratch& ws, CellCta& cs, float* tex, V4& X, V4& N, int nv, float dscale, uint64_t stream) {
    const Params& p = cp.p;
    const int tid = (int)threadIdx.x;
    const bool w0 = tid < 32;
    double best[3] = {0.0, 0.0, 0.0}, fbest = 0.0;
    double r[3] = {4.0, 4.0, 4.0};
    if (w0) {
        const double lb[3] = {-(double)__int_as_float(0x7f800000), -23.99999, -23.99999};
        const double ub[3] = {(double)__int_as_float(0x7f800000), 23.99999, 23.99999};
        RefineCtx rc;
        rc.center = X;
        rc.ref = ws.images[0];
        rc.ray = sub4(X, ld4(p.views[rc.ref].center));
        rc.ray = div4(rc.ray, norm4(rc.ray));
        rc.dscale = dscale;
        if (tid == 0) cs.rc = rc;
        compute_weights(p, X, N, ws.images, nv, ws.units, tid);                   // m_weights of the UNREFINED patch (optim.cpp:490)
        encode(cp, rc, X, N, best);
#pragma unroll
        for (int i = 0; i < 3; ++i) best[i] = fmax(fmin(best[i], ub[i]), lb[i]);
    }
    const int sz = min(p.tau, nv);
    const int minimum = min(p.min_image_num, sz);
    {   // the sz views every cost evaluation of this refinement reads: per-view constants into shared memory, word by word
        constexpr int VW = sizeof(ViewConst) / 4;
        for (int w = tid; w < sz * VW; w += NW * 32) {
            const int i = w / VW, o = w - i * VW, v = ws.images[i];
            const bool ok = v >= 0 && v < p.nviews;
            if (o == 0) cs.fr_ok[i] = ok ? 1 : 0;
            reinterpret_cast<unsigned int*>(&cs.vc[i])[o] = reinterpret_cast<const unsigned int*>(p.views + (ok ? v : 0))[o];
        }
    }
    __syncthreads();
    cta_costs<WS, NW>(cp, cs, tex, ws.images, sz, 1, stream, -1, best[0], best[1], best[2], 0.0, 0.0, 0.0);
    if (w0) fbest = eval_cost(cs, 0, sz, minimum);
#pragma unroll 1
    for (int level = 0; level < PMR1_LEVELS; ++level) {
        cta_costs<WS, NW>(cp, cs, tex, ws.images, sz, PMR1_CANDS, stream, level, best[0], best[1], best[2], r[0], r[1], r[2]);
        if (w0) {
            // argmin over the level's candidates in index order, lowest index on ties (strict <)
            const double mine = eval_cost(cs, tid < PMR1_CANDS ? tid : 0, sz, minimum);        // lane c: the cost of candidate c
            double fwin = __shfl_sync(0xffffffffu, mine, 0);
            int cwin = 0;
            for (int c = 1; c < PMR1_CANDS; ++c) { const double fc = __shfl_sync(0xffffffffu, mine, c); if (fc < fwin) { fwin = fc; cwin = c; } }
            if (fwin < fbest) { fbest = fwin; double xw[3]; pmr1_point(cp, stream, level, cwin, best, r, xw); best[0] = xw[0]; best[1] = xw[1]; best[2] = xw[2]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) r[i] = __dmul_rn(r[i], 0.6);
        }
        PMK_SUBT(4);
    }
    // optim.cpp:534-541: decode, normal.w = 0, ncc = 1.0 - unrobustincc(computeINCC(...)) with the stale weights
    if (nv >= 2) cta_costs<WS, NW>(cp, cs, tex, ws.images, sz, 1, stream, -1, best[0], best[1], best[2], 0.0, 0.0, 0.0);
    if (tid == 0) {
        // the final point was decoded by cta_costs' first step (candidate 0); with fewer than two views nothing was evaluated
        V4 Xf, Nf;
        decode(cp, cs.vc[0], cs.rc, best, Xf, Nf);
        float incc = 2.0f;
        if (nv >= 2 && cs.fr[0].level >= 0) {
            float score = 0.0f, tw = 0.0f;
            for (int i = 1; i < sz; ++i)
                if (cs.fr[i].level >= 0) { tw = xadd(tw, ws.units[i]); score = xadd(score, xmul(cs.val[i], ws.units[i])); }
            incc = (tw == 0.0f) ? 2.0f : xdiv(score, tw);
        }
        cs.bf[0] = Xf.x; cs.bf[1] = Xf.y; cs.bf[2] = Xf.z; cs.bf[3] = Xf.w;
        cs.bf[4] = Nf.x; cs.bf[5] = Nf.y; cs.bf[6] = Nf.z;
        cs.bf[7] = __double2float_rn(1.0 - (double)unrobustincc(incc));
    }
    __syncthreads();
    X = V4{cs.bf[0], cs.bf[1], cs.bf[2], cs.bf[3]};
    N = V4{cs.bf[4], cs.bf[5], cs.bf[6], 0.0f};
    const float ncc = cs.bf[7];
    __syncthreads();
    return ncc;
}

// ---- Optim::setRefImage (optim.cpp:348-383) by the CTA; same arithmetic as warp_set_ref_image ------------------------------------------
template <int WS, int NW>
__device__ __noinline__ void cta_set_ref_image(const CandParams& cp, WarpScratch& ws, CellCta& cs, float* tex, V4 X, V4 N, int nv, int wslot) {
    typedef CellGeom<WS, NW> Gm;
    constexpr int TEXW = WS * WS * 3;
    const Params& p = cp.p;
    const CellLane<WS, NW> L;
    float* gtex = cp.tex_scratch + (size_t)wslot * p.nviews * (TEXW + 4);
    float* mat = cp.mat_scratch + (size_t)wslot * p.nviews * p.nviews;
    if (L.tid < ((nv + 31) & ~31)) {
        CandGeo cg;
        cg.X = X; cg.N = N;
        get_paxes(p.views[ws.images[0]], X, N, p.level_scale, cg.px, cg.py);
        if (L.tid < nv) frame_item(p, view_or_null(p, ws.images[L.tid]), ws.images[L.tid], cg, &cs.fr[L.tid]);
    }
    __syncthreads();
    float* mine = tex + (size_t)L.g * Gm::SLOT;
#pragma unroll 1
    for (int i = L.g; i < nv; i += Gm::TG) {
        const int lv = cs.fr[i].level;
        float* dst = gtex + (size_t)i * (TEXW + 4);
        if (lv >= 0) {
            const float inv = sample_slot<WS, Gm::GW>(p.views[cs.fr[i].view], cs.fr[i], L.col, L.gm, mine);
            if (L.col < WS) {
#pragma unroll
                for (int y = 0; y < WS; ++y) {
                    float* q = dst + (y * WS + L.col) * 3;
                    q[0] = mine[(y * 3 + 0) * WS + L.col] * inv; q[1] = mine[(y * 3 + 1) * WS + L.col] * inv; q[2] = mine[(y * 3 + 2) * WS + L.col] * inv;
                }
            }
        }
        if (L.col == 0) dst[TEXW] = lv >= 0 ? 1.0f : 0.0f;
        __syncwarp(L.gm);
    }
    __syncthreads();
#pragma unroll 1
    for (int q = L.tid; q < nv * nv; q += Gm::NT) {
        const int i = q / nv, j = q % nv;
        if (j < i) continue;
        float v = 0.0f;
        if (j > i) {
            const float* a = gtex + (size_t)i * (TEXW + 4);
            const float* b = gtex + (size_t)j * (TEXW + 4);
            v = 2.0f;
            if (a[TEXW] != 0.0f && b[TEXW] != 0.0f) {
                float dp = 0.0f;
                for (int e = 0; e < TEXW; ++e) dp = fmaf(a[e], b[e], dp);
                v = robustincc(xsub(1.0f, dp * (1.0f / (float)TEXW)));
            }
        }
        mat[i * nv + j] = v; mat[j * nv + i] = v;
    }
    __syncthreads();
    if (L.warp == 0) {
        float best = 1073741824.0f;
        int bi = -1;
        for (int i = L.lane; i < nv; i += 32) {
            float sum = 0.0f;
            for (int j = 0; j < nv; ++j) sum = xadd(sum, mat[i * nv + j]);
            if (sum < best) { best = sum; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi >= 0 && (bi < 0 || ob < best || (ob == best && oi < bi))) { best = ob; bi = oi; }
        }
        if (L.lane == 0 && bi > 0) { const int tmp = ws.images[0]; ws.images[0] = ws.images[bi]; ws.images[bi] = tmp; }
    }
    __syncthreads();
}

// ---- Optim::preProcess (optim.cpp:137-163) by the CTA; every thread gets the same return value, nv, dscale, ascale ------------------
template <int WS, int NW>
__device__ __forceinline__ int cta_pre_process(const CandParams& cp, WarpScratch& ws, CellCta& cs, float* tex, V4 X, V4 N, int& nv, float& dscale, float& ascale) {
    const Params& p = cp.p;
    const int tid = (int)threadIdx.x;
    dscale = 0.0f; ascale = 0.0f;
    if (nv < 1) { nv = 0; return -1; }
    if (tid < 32) {
        const int n2 = warp_add_images(p, X, N, ws.images, nv, CAND_MAXV, ws.mark, tid);                // optim.cpp:139
        if (tid == 0) cs.bi[0] = n2;
    }
    __syncthreads();
    nv = cs.bi[0];
    cta_set_inccs<WS, NW>(p, cs, tex, X, N, ws.images, nv, 0, ws.inccs);                                    // constraintImages, :141
    if (tid < 32) {
        int n2 = warp_constraint(ws.images, ws.inccs, nv, p.ncc_threshold_before, tid);
        __syncwarp();
        n2 = warp_sort_images(cp, X, N, ws, n2, tid);                                                   // :143
        __syncwarp();
        float ds = 0.0f, as = 0.0f;
        int r = -1;
        if (n2 > 0) warp_set_scales(p, X, ws.images, n2, ds, as, ws.units, tid);                         // :145-147
        if (n2 >= p.min_image_num) {                                                                    // :149
            if (warp_check_angles(cp, X, ws, n2, tid)) r = 0;                                           // :153-160
            else n2 = 0;
        }
        if (tid == 0) { cs.bi[0] = n2; cs.bi[1] = r; cs.bf[0] = ds; cs.bf[1] = as; }
    }
    __syncthreads();
    nv = cs.bi[0]; dscale = cs.bf[0]; ascale = cs.bf[1];
    const int r = cs.bi[1];
    __syncthreads();
    return r;
}

// ---- Optim::postProcess (optim.cpp:260-290), the store-independent part, by the CTA ---------------------------------------------------
template <int WS, int NW>
__device__ __forceinline__ int cta_post_process(const CandParams& cp, WarpScratch& ws, CellCta& cs, float* tex, V4 X, V4 N, int& nv, int wslot) {
    const Params& p = cp.p;
    const int tid = (int)threadIdx.x;
    if (nv < p.min_image_num) return -1;                                                                // :261
    if (tid < 32) {
        const int m = warp_get_mask(p, X, tid);                                                         // :265
        int n2 = nv;
        if (m != 0) n2 = warp_add_images(p, X, N, ws.images, nv, CAND_MAXV, ws.mark, tid);               // :268
        if (tid == 0) { cs.bi[0] = n2; cs.bi[1] = m; }
    }
    __syncthreads();
    nv = cs.bi[0];
    const int masked = cs.bi[1];
    __syncthreads();
    if (masked == 0) return -1;
    cta_set_inccs<WS, NW>(p, cs, tex, X, N, ws.images, nv, 0, ws.inccs);                                    // :269
    if (tid < 32) {
        int n2 = warp_constraint(ws.images, ws.inccs, nv, p.ncc_threshold, tid);
        __syncwarp();
        n2 = warp_filter_by_angle(p, X, N, ws.images, n2, tid);                                         // :270
        if (tid == 0) cs.bi[0] = n2;
    }
    __syncthreads();
    nv = cs.bi[0];
    __syncthreads();
    if (nv < p.min_image_num) return -1;                                                                // :272
    cta_set_ref_image<WS, NW>(cp, ws, cs, tex, X, N, nv, wslot);                                            // :277
    cta_set_inccs<WS, NW>(p, cs, tex, X, N, ws.images, nv, 0, ws.inccs);                                    // :279
    if (tid < 32) {
        const int n2 = warp_constraint(ws.images, ws.inccs, nv, p.ncc_threshold, tid);
        if (tid == 0) cs.bi[0] = n2;
    }
    __syncthreads();
    nv = cs.bi[0];
    __syncthreads();
    if (nv < p.min_image_num) return -1;                                                                // :281
    return 0;
}

// ---- Filter::computeGain (filter.cpp:108-146) by the CTA: one warp per registration, lanes over the cell's slots --------------------
// The maximum over a cell does not depend on the order; the subtractions run in list order like the reference.
template <int NW>
__device__ __forceinline__ float cta_compute_gain(const StoreParams& sp, CellCta& cs, const PGeo& me, float ncc, const PatchLists& pl, const Overlay& ov) {
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tot = pl.nimg + pl.nvimg;
    for (int i = warp; i < tot; i += NW) {
        const bool isv = i >= pl.nimg;
        const int img = isv ? pl.vimages[i - pl.nimg] : pl.images[i], pc = isv ? pl.vcells[i - pl.nimg] : pl.cells[i];
        const float pdepth = isv ? dot4(ld4(p.views[img].oaxis), me.X) : 0.0f;
        const int c = cell_global(sp, img, cell_x(pc), cell_y(pc));
        const bool local = (c == ov.cell);
        const int n = local ? ov.n : min(st.ccount[c], st.cell_cap);
        float mp = 0.0f;
        for (int s = lane; s < n; s += 32) {
            const int e = local ? ov.ids[s] : st.cslots[(size_t)c * st.cell_cap + s];
            if (e == SLOT_TOMB || e < 0 || (!local && overlay_removed(ov, e))) continue;
            const PGeo q = load_geo(st, e);
            if (isv && !(pdepth < dot4(ld4(p.views[img].oaxis), q.X))) continue;                   // Camera::computeDepth (camera.cpp:339-346)
            if (!is_neighbor(sp, me, q, sp.neighbor_threshold1)) mp = max_std(mp, xsub(st.scal[e].x, p.ncc_threshold));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mp = max_std(mp, __shfl_xor_sync(0xffffffffu, mp, o));
        if (lane == 0) cs.mp[i] = mp;
    }
    __syncthreads();
    float gain = xmul(max_std(0.0f, xsub(ncc, p.ncc_threshold)), (float)pl.nimg);                 // score2 (patch.cpp:27-29)
    for (int i = 0; i < tot; ++i) gain = xsub(gain, cs.mp[i]);
    __syncthreads();
    return gain;
}

// ---- PatchManager::findNeighbors (patch_manager.cpp:671-728) by the CTA: one warp per view of m_images ----------------------------------
// Same cells, same tests, same (sorted) result as warp_find_neighbors; the (token, id) table and the output list are shared by the warps.
template <int NW>
__device__ __forceinline__ int cta_find_neighbors(const StoreParams& sp, CellCta& cs, const PGeo& me, const PatchLists& pl, float scale, int margin,
                                                  const Overlay& ov, int* out) {
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float u1 = __int_as_float(0x7f800000), u2 = u1;
    float usum = 0.0f;
    for (int i = 0; i < pl.nimg; ++i) {                                 // Propagate::computeRadius (propagate.cpp:474-481)
        const ViewConst& vc = p.views[pl.images[i]];
        const float gu = get_unit(vc, me.X, p.level_scale);
        usum = xadd(usum, gu);
        V4 ray = sub4(ld4(vc.center), me.X);
        ray = div4(ray, norm4(ray));
        const float d = dot4(ray, me.N);
        const float u = (0.0f < d) ? xdiv(gu, d) : 1073741824.0f;
        if (u < u1) { u2 = u1; u1 = u; } else if (u < u2) u2 = u;
    }
    const float rad0 = xmul(u2, (float)p.csize);
    const float radius = __double2float_rn(__dmul_rn(1.5 * (double)margin, (double)rad0));
    const float unit = xmul(xdiv(usum, (float)pl.nimg), (float)p.csize);
    const float thr = xmul(sp.neighbor_threshold, scale);
    unsigned long long* table = reinterpret_cast<unsigned long long*>(out + 2 * NB_CAP);
    unsigned int* tokp = reinterpret_cast<unsigned int*>(table + NB_HASH);
    if (tid == 0) {
        unsigned int token = *tokp + 1;
        if (token == 0) token = 1;
        *tokp = token;
        cs.nb_token = token; cs.nb_count = 0; cs.nb_over = 0;
    }
    __syncthreads();
    const unsigned int token = cs.nb_token;
    bool overflow = false;
    const int side = 2 * margin + 1, ncell = side * side;
    for (int i = warp; i < pl.nimg; i += NW) {
        const int img = pl.images[i];
        const ViewConst& vc = p.views[img];
        const int ix = cell_x(pl.cells[i]), iy = cell_y(pl.cells[i]);
        int myc = -1, myn = 0, mynslot = 0;
        if (lane < ncell) {
            const int yt = iy + lane / side - margin, xt = ix + lane % side - margin;
            if (!(yt < 0 || vc.gh <= yt || xt < 0 || vc.gw <= xt)) {
                myc = cell_global(sp, img, xt, yt);
                mynslot = min(st.ccount[myc], st.cell_cap);
                myn = mynslot + (myc == ov.cell ? ov.n : 0);
            }
        }
        int incl = myn;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int base = 0; base < total; base += 32) {
            const int f = min(base + lane, total - 1);
            int j = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) { const int v = __shfl_sync(0xffffffffu, incl, j + step - 1); if (v <= f) j += step; }
            const int c = __shfl_sync(0xffffffffu, myc, j), nslot = __shfl_sync(0xffffffffu, mynslot, j);
            const int slot = f - (__shfl_sync(0xffffffffu, incl, j) - __shfl_sync(0xffffffffu, myn, j));
            int id = -1;
            if (base + lane < total && c >= 0) {
                if (slot < nslot) {
                    const int e = st.cslots[(size_t)c * st.cell_cap + slot];
                    if ((e & 0x7fffffff) != SLOT_TOMB) {
                        const bool isv = e < 0;
                        const int q = e & 0x7fffffff;
                        if (!(c == ov.cell && !isv) && !overlay_removed(ov, q)) id = q;
                    }
                } else id = ov.ids[slot - nslot];
            }
            bool hit = false;
            if (id >= 0 && nb_insert(table, token, id, overflow)) hit = is_neighbor_radius(sp, me, load_geo(st, id), unit, thr, radius) != 0;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            int wbase = 0;
            if (m != 0 && lane == 0) wbase = atomicAdd(&cs.nb_count, __popc(m));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (hit) { const int pos = wbase + __popc(m & ((1u << lane) - 1u)); if (pos < NB_CAP) out[pos] = id; }
        }
    }
    if (overflow) cs.nb_over = 1;
    __syncthreads();
    int nuni = cs.nb_count;
    if (nuni > NB_CAP) { nuni = NB_CAP; if (tid == 0) cs.nb_over = 1; }
    // ascending id order, as warp_find_neighbors leaves it (the quadric fit sums over the neighbours in this order)
    int* tmp = out + NB_CAP;
    for (int k = tid; k < nuni; k += NW * 32) {
        const int id = out[k];
        int rank = 0;
        for (int j = 0; j < nuni; ++j) rank += out[j] < id ? 1 : 0;
        tmp[rank] = id;
    }
    __syncthreads();
    for (int k = tid; k < nuni; k += NW * 32) out[k] = tmp[k];
    if (tid == 0 && cs.nb_over) atomicAdd(st.counters + SC_NBOVER, 1);
    __syncthreads();
    return nuni;
}

// warp-0 pieces kept out of line so that their registers (double-precision normal equations, the sorted cell lists) do not set
// the budget of the whole kernel
__device__ __noinline__ int quad_out_of_line(const StoreParams& sp, const PGeo& me, const PatchLists& pl, const int* nb, int n, int lane) {
    return warp_filter_quad(sp, me, pl, nb, n, nullptr, lane);
}
__device__ __noinline__ int topk_trim_out_of_line(const StoreParams& sp, int c, int keep, int* id, float* ncc, unsigned int* birth, int* removed, int* nremoved,
                                                  int* rem_list, int* ntrim, int lane) {
    return warp_load_cell_topk<true>(sp, c, keep, id, ncc, birth, removed, nremoved, rem_list, ntrim, lane);
}
__device__ __noinline__ int topk_out_of_line(const StoreParams& sp, int c, int keep, int* id, float* ncc, unsigned int* birth, int lane) {
    return warp_load_cell_topk<false>(sp, c, keep, id, ncc, birth, nullptr, nullptr, nullptr, nullptr, lane);
}

// =====================================================================================================================================
// the kernel: dest cells of one wavefront step, handed out longest-first through a counter; one CTA per cell at a time
// =====================================================================================================================================
template <int WS, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) k4_cells(const __grid_constant__ StoreParams sp, const __grid_constant__ SweepArgs sa) {
    typedef CellGeom<WS, NW> Gm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& ws = *reinterpret_cast<WarpScratch*>(smem_raw);
    SweepScratch& ss = *reinterpret_cast<SweepScratch*>(smem_raw + sizeof(WarpScratch));
    CellCta& cs = *reinterpret_cast<CellCta*>(smem_raw + sizeof(WarpScratch) + sizeof(SweepScratch));
    float* tex = reinterpret_cast<float*>(smem_raw + ((sizeof(WarpScratch) + sizeof(SweepScratch) + sizeof(CellCta) + 15) & ~(size_t)15));
    const CandParams& cp = sp.cp;
    const Params& p = cp.p;
    const StoreDev& st = sp.st;
    const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wslot = sa.wslot_base + blockIdx.x;              // this CTA's slice of the pairwise / findNeighbors scratch
    const int inc = sa.inc;
    const int maxp = sp.max_patches_cell;
    if (tid < SS_COUNT) cs.stat[tid] = 0;
    if (tid < 8) cs.sub[tid] = 0;
    __syncthreads();
#define PMK_STAT(slot, v) do { if (tid == 0) cs.stat[slot] += (v); } while (0)
    // profiling only: thread 0 closes the phase that ends here
#define PMK_PHASE(slot)                                                                                    \
    if (sa.phase_ns != nullptr && tid == 0) {                                                              \
        unsigned long long now_;                                                                           \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));                                           \
        if ((slot) >= 0) atomicAdd(sa.phase_ns + (slot), now_ - cs.ph_t);                                  \
        cs.ph_t = now_;                                                                                    \
    }

    for (;;) {
        if (tid == 0) cs.task = atomicAdd(st.counters + SC_NEXT, 1);
        __syncthreads();
        const int tslot = cs.task;
        if (tslot >= sa.ntasks) break;
        const int task = sa.order[tslot];
        const int G = sweep_global_task(sa, task), g = sweep_group_of(sa, G);
        const int img = sa.g_img[g];
        const ViewConst& vimgc = p.views[img];
        const int gw = vimgc.gw, gh = vimgc.gh;
        const int x = sa.g_xlo[g] + (G - sa.g_off[g]), y = sa.g_diag[g] - x;
        const int cD = st.cell_base[img] + y * gw + x;
        if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(cs.t_begin));
        // ================= preamble: sortPatches' recomputation of negative m_ncc (patch_manager.cpp:411-415), by the CTA =================
        {
            const int n = min(st.ccount[cD], st.cell_cap);
            for (int base = 0; base < n; base += 32) {
                if (warp == 0) {
                    const int s2 = base + lane;
                    int e = SLOT_TOMB;
                    if (s2 < n) e = st.cslots[(size_t)cD * st.cell_cap + s2];
                    const bool neg = e != SLOT_TOMB && e >= 0 && st.state[e] == 1 && st.scal[e].x < 0.0f;
                    const unsigned msk = __ballot_sync(0xffffffffu, neg);
                    if (neg) cs.negs[__popc(msk & ((1u << lane) - 1u))] = e;
                    if (lane == 0) cs.nneg = __popc(msk);
                }
                __syncthreads();
                const int nneg = cs.nneg;
                for (int j = 0; j < nneg; ++j) {
                    const int el = cs.negs[j];
                    const int nv = min(st.nimg[el], CAND_MAXV);
                    for (int k = tid; k < nv; k += NW * 32) ws.images[k] = st.images[(size_t)el * st.maxv + k];
                    __syncthreads();
                    const float v = cta_compute_ncc<WS, NW>(p, ws, cs, tex, f4v(st.coord[el]), f4v(st.normal[el]), nv);
                    if (tid == 0) st.scal[el].x = v;
                }
                __syncthreads();
            }
        }
        // ================= preamble (warp 0): D's list, trim, sources (propagate.cpp:88-99,123-134) =================
        if (warp == 0) {
            int nrem = 0, ntrim = 0;
            const int nl = topk_trim_out_of_line(sp, cD, maxp, cs.l_id, cs.l_ncc, cs.l_birth, cs.removed, &nrem, sa.rem_list, &ntrim, lane);
            if (lane == 0) cs.stat[SS_TRIMMED] += ntrim;
            int nsrc = 0;
            for (int side = 0; side < 2; ++side) {              // (x, y - inc) first, then (x - inc, y); sorted top-maxp, reference view == img
                const int sx = side == 0 ? x : x - inc, sy = side == 0 ? y - inc : y;
                if (sx < 0 || gw <= sx || sy < 0 || gh <= sy) continue;
                int* tid_ = ss.vimg; float* tncc = reinterpret_cast<float*>(ss.vcell);       // scratch, free until stage B
                unsigned int* tbirth = reinterpret_cast<unsigned int*>(ss.cells);
                const int m = topk_out_of_line(sp, st.cell_base[img] + sy * gw + sx, maxp, tid_, tncc, tbirth, lane);
                for (int i = 0; i < m && nsrc < SRC_MAX; ++i) {
                    const int e = tid_[i];
                    if (st.images[(size_t)e * st.maxv] == img) { if (lane == 0) cs.src_id[nsrc] = e; ++nsrc; }
                }
                __syncwarp();
            }
            if (lane == 0) { cs.nl = nl; cs.nrem = nrem; cs.nnew = 0; cs.nsrc = nsrc; }
        }
        __syncthreads();
        const int ntries = 2 * cs.nsrc;                                                    // MAX_NUM_OF_PROPAG tries per call
        const bool forced = sa.force.code != nullptr;
        bool fill0 = false;                                                                // try 0 of the current call drew its jitter
        if (forced && tid == 0) *sa.force.o_ntries = ntries;
        // ================= the propagatePatch tries (propagate.cpp:122-218), in order, each one spread over the CTA =================
#pragma unroll 1
        for (int t = 0; t < ntries; ++t) {
            const int call = t >> 1, k = t & 1;
            const int src = cs.src_id[call];
            const V4 sX = f4v(st.coord[src]), sN = f4v(st.normal[src]);
            const int sref = st.images[(size_t)src * st.maxv];
            const int snv = min(st.nimg[src], CAND_MAXV);
            // one PMR1 stream per call: (iter, view, dest cell, call ordinal)
            const uint64_t stream = ((uint64_t)(unsigned)sa.iter << 56) ^ ((uint64_t)(unsigned)img << 40) ^ ((uint64_t)(unsigned)(y * gw + x) << 8) ^ (uint64_t)call;
            const int np = cs.nl;
            const int wid = np >= maxp ? cs.l_id[maxp - 1] : -1;
            const float wncc = np >= maxp ? cs.l_ncc[maxp - 1] : 0.0f;
            PMK_PHASE(-1)
            if (sa.phase_ns != nullptr && tid == 0) atomicAdd(sa.phase_ns + 6, 1ull);
            Cand cd;
            cd.nv = 0; cd.nvv = 0; cd.ncc = 0.f; cd.dscale = 0.f; cd.ascale = 0.f; cd.tmp = 0.f;
            cd.X = V4{0.f, 0.f, 0.f, 1.f}; cd.N = sN;
            int outcome = TRY_GEN_NULL;
            int post_r = -2;                                                           // postProcess' return once it has run
            bool refined = false;                                                      // cd holds the patch as refinePatch left it
            if (!forced) {
                V3 ic;
                if (np < maxp) {
                    // the reference draws from an engine it re-constructs per propagatePatch call (propagate.cpp:139-141): the
                    // first fill-branch try of a call gets draws 0, 1 and the second one draws 2, 3
                    const int dr = (k == 1 && fill0) ? 2 : 0;
                    float jx = sa.jitter[dr], jy = sa.jitter[dr + 1];
                    if (sa.jitter_mode == 1) {
                        uint32_t ctr[4] = {(uint32_t)stream, (uint32_t)(stream >> 32), 0x4a495454u, (uint32_t)k};
                        philox4x32_10((uint32_t)cp.seed, (uint32_t)(cp.seed >> 32), ctr);
                        jx = (float)(0.5 * uniform_pm1(ctr[0])); jy = (float)(0.5 * uniform_pm1(ctr[1]));
                    }
                    const float cxf = (float)(p.csize * (2 * x + 1) - 1) / 2.0f, cyf = (float)(p.csize * (2 * y + 1) - 1) / 2.0f;
                    ic = V3{xadd(cxf, xmul(jx, (float)p.csize)), xadd(cyf, xmul(jy, (float)p.csize)), xadd(1.0f, 0.0f)};
                } else ic = project(vimgc.P, f4v(st.coord[wid]));
                // ---- generatePatch (propagate.cpp:220-237) ----
                const ViewConst& vr = p.views[sref];
                const float depth = dot4(ld4(vr.oaxis), sX);
                V4 X = unproject_rows(vr.P, vr.Minv, V3{xmul(depth, ic.x), xmul(depth, ic.y), xmul(depth, ic.z)});         // Camera::unproject (camera.cpp:329-337)
                V4 N = sN;
                if (warp == 0) {                                                           // setGridsImages (patch_manager.cpp:223-239)
                    int nv0 = 0;
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
                        if (keep) ws.images[nv0 + __popc(msk & ((1u << lane) - 1u))] = v;
                        nv0 += __popc(msk);
                    }
                    if (lane == 0) cs.bi[0] = nv0;
                }
                __syncthreads();
                int nv = cs.bi[0];
                __syncthreads();
                if (nv > 0) {
                    float ncc = cta_compute_ncc<WS, NW>(p, ws, cs, tex, X, N, nv);
                    PMK_PHASE(0)
                    if (np >= maxp && ncc < wncc) outcome = TRY_LOSE;
                    else {
                        // ---- patch optimisation (propagate.cpp:176-193) ----
                        float dscale, ascale;
                        const int pre = cta_pre_process<WS, NW>(cp, ws, cs, tex, X, N, nv, dscale, ascale);
                        PMK_PHASE(1)
                        if (pre == -1) outcome = TRY_FAIL0;
                        else {
                            if (sa.phase_ns != nullptr && tid == 0) atomicAdd(sa.phase_ns + 7, 1ull);
                            ncc = cta_refine<WS, NW>(cp, ws, cs, tex, X, N, nv, dscale, stream);
                            PMK_PHASE(2)
                            refined = true;
                            cd.X = X; cd.N = N; cd.nv = nv; cd.ncc = ncc; cd.dscale = dscale; cd.ascale = ascale;
                        }
                    }
                }
            } else {
                // ---- teacher forcing (tests): the try starts from a recorded hypothesis instead of generating and refining one ----
                const int fcode = t < sa.force.ntries ? sa.force.code[t] : 0;
                if (fcode != 0) {
                    if (np >= maxp && sa.force.ncc0[t] < wncc) outcome = TRY_LOSE;       // propagate.cpp:170, on the recorded m_ncc
                    else if (fcode == 1) outcome = TRY_DIVERGED;                          // the recorded run lost here; nothing to continue from
                    else if (fcode == 2) outcome = TRY_FAIL0;
                    else {
                        const float4 sc = sa.force.scal[t];
                        cd.X = f4v(sa.force.coord[t]); cd.N = f4v(sa.force.normal[t]);
                        cd.ncc = sc.x; cd.dscale = sc.y; cd.ascale = sc.z;
                        cd.nv = min(sa.force.nimg[t], CAND_MAXV);
                        for (int i = tid; i < cd.nv; i += NW * 32) ws.images[i] = sa.force.images[(size_t)t * sa.force.stride + i];
                        __syncthreads();
                        refined = true;
                    }
                }
            }
            if (refined) {
                int nv = cd.nv;
                const int r = cta_post_process<WS, NW>(cp, ws, cs, tex, cd.X, cd.N, nv, wslot);
                PMK_PHASE(3)
                post_r = r;
                outcome = r == 0 ? TRY_ACCEPT : TRY_FAIL1;
                cd.nv = nv;
                if (r == 0) {
                    bool outside = false;
                    for (int i = tid; i < nv; i += NW * 32) {                          // setGrids (optim.cpp:285)
                        const V3 q = project(p.views[ws.images[i]].P, cd.X);
                        const int ix = cell_of(q.x, p.csize), iy = cell_of(q.y, p.csize);
                        outside |= ix < 0 || p.views[ws.images[i]].gw <= ix || iy < 0 || p.views[ws.images[i]].gh <= iy;
                        ss.cells[i] = pack_cell(ix, iy);
                    }
                    // Views inherited from the source patch are not re-checked by addImages, and the refinement moves
                    // the patch: a view can end up seeing it outside its grid.  The reference then writes m_pgrids out of
                    // bounds (addPatch, patch_manager.cpp:164-170); here the candidate is rejected like any postProcess failure.
                    if (__syncthreads_or(outside ? 1 : 0)) outcome = TRY_FAIL1;
                }
            }
            if (k == 0) fill0 = np < maxp;
            // ---------------- postProcess's store-reading tail: setVImagesVGrids, check (optim.cpp:288-296) ----------------
            if (outcome == TRY_ACCEPT) {
                cd.tmp = xmul(max_std(0.0f, xsub(cd.ncc, p.ncc_threshold)), (float)cd.nv);      // m_tmp = score2
                cd.nvv = 0;
                if (p.depth) {
                    if (warp == 0) {
                        const int nvv = warp_set_vimages(sp, ws, cd.X, cd.N, ws.images, cd.nv, ss.vimg, ss.vcell, 0, lane);
                        if (lane == 0) cs.bi[0] = nvv;
                    }
                    __syncthreads();
                    cd.nvv = cs.bi[0];
                    __syncthreads();
                }
                if (2 <= p.depth) {                                                    // Optim::check (optim.cpp:300-323)
                    PGeo me; me.X = cd.X; me.N = cd.N; me.dscale = cd.dscale; me.ref = ws.images[0];
                    const PatchLists pl{ws.images, ss.cells, cd.nv, ss.vimg, ss.vcell, cd.nvv};
                    const Overlay ov{cD, cs.l_id, min(cs.nl, LKEEP), cs.removed, cs.nrem};
                    const float gain = cta_compute_gain<NW>(sp, cs, me, cd.ncc, pl, ov);
                    cd.tmp = gain;
                    if (gain < 0.0f) outcome = TRY_FAIL1;
                    else {
                        int* nb = sp.nb_scratch + (size_t)wslot * NB_STRIDE;
                        const int nn = cta_find_neighbors<NW>(sp, cs, me, pl, 4.0f, 2, ov, nb);
                        if (6 < nn) {
                            if (warp == 0) { const int rej = quad_out_of_line(sp, me, pl, nb, nn, lane); if (lane == 0) cs.bi[0] = rej; }
                            __syncthreads();
                            if (cs.bi[0]) outcome = TRY_FAIL1;
                            __syncthreads();
                        }
                    }
                }
                PMK_PHASE(4)
            }
            if (forced && t < sa.force.ntries) {
                // what the try decided, for the comparison with the recorded run
                const ForceIO& fo = sa.force;
                if (tid == 0) {
                    fo.o_outcome[t] = outcome; fo.o_full[t] = np >= maxp ? 1 : 0; fo.o_post[t] = post_r;
                    fo.o_nimg[t] = post_r == 0 ? cd.nv : 0; fo.o_nvimg[t] = outcome == TRY_ACCEPT || post_r == 0 ? cd.nvv : 0; fo.o_tmp[t] = cd.tmp;
                }
                if (post_r == 0) {
                    for (int i = tid; i < cd.nv; i += NW * 32) { fo.o_images[(size_t)t * fo.stride + i] = ws.images[i]; fo.o_cells[(size_t)t * fo.stride + i] = ss.cells[i]; }
                    for (int i = tid; i < cd.nvv; i += NW * 32) { fo.o_vimages[(size_t)t * fo.stride + i] = ss.vimg[i]; fo.o_vcells[(size_t)t * fo.stride + i] = ss.vcell[i]; }
                }
            }
            // ================= commit =================
            PMK_STAT(SS_CALLS, (k == 0) ? 1 : 0);
            PMK_STAT(SS_TRIES, 1);
            if (outcome != TRY_GEN_NULL) PMK_STAT(SS_EVALS, 1);
            if (outcome == TRY_FAIL1 || outcome == TRY_ACCEPT) PMK_STAT(SS_EVALS, PMR1_EVALS + 1);
            if (outcome == TRY_GEN_NULL) PMK_STAT(SS_GEN_NULL, 1);
            else if (outcome == TRY_LOSE) PMK_STAT(SS_NCC_LOSE, 1);
            else if (outcome == TRY_FAIL0) PMK_STAT(SS_FAIL0, 1);
            else if (outcome == TRY_FAIL1) PMK_STAT(SS_FAIL1, 1);
            else if (outcome == TRY_ACCEPT && warp == 0) {
                // ---- removePatch(worst) / addPatch(new) (propagate.cpp:195-207); grid updates are staged ----
                int nl = cs.nl;
                if (nl == maxp) {
                    const int w = cs.l_id[maxp - 1];
                    if (w >= st.cap) { if (lane == 0) st.state[w] = 0; }                       // staged this step: never reaches the grids
                    else if (lane == 0) { cs.removed[cs.nrem] = w; cs.nrem = cs.nrem + 1; }
                    --nl;
                    if (lane == 0) cs.stat[SS_REPLACED] += 1;
                } else if (lane == 0) cs.stat[SS_ADDED] += 1;
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
                }
            }
            PMK_PHASE(5)
            __syncthreads();
        }
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
                atomicAdd(sa.stats + SS_CELL_NS, t_end - cs.t_begin);
                atomicMax(sa.step_max, t_end - cs.t_begin);
                if (sa.cell_ns) sa.cell_ns[cD] = (float)(t_end - cs.t_begin);
            }
        }
        __syncthreads();
    }
#undef PMK_PHASE
#undef PMK_STAT
    if (tid < SS_COUNT && cs.stat[tid]) atomicAdd(sa.stats + tid, (unsigned long long)cs.stat[tid]);
#ifdef PMK_SUBPHASE
    if (sa.phase_ns != nullptr && tid < 8 && cs.sub[tid]) atomicAdd(sa.phase_ns + 8 + tid, cs.sub[tid]);
#endif
}

}  // namespace pmk
