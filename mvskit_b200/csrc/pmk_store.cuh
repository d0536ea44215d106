// mvskit_b200/csrc/pmk_store.cuh -- the device patch store (PatchManager's grids as structure-of-arrays in HBM)
// and the store-reading decisions of the reference: isNeighbor / isNeighborRadius, isVisible, updateDepthMaps,
// setVImagesVGrids, computeGain, findNeighbors + filterQuad.
//
// Layout:
//   patch p (0 <= p < cap, plus a staging region [cap, cap + stage_cap) used by the sweep):
//       coord[p], normal[p]            float4        Patch::m_coord / m_normal                    (patch.hpp:33-35)
//       scal[p] = {ncc, dscale, ascale, tmp}         Patch::m_ncc / m_dscale / m_ascale / m_tmp   (patch.hpp:44-66)
//       images[p][maxv], cells[p][maxv], nimg[p]     Patch::m_images / m_grids (ix | iy << 16)    (patch.hpp:38-39)
//       vimages[p][maxv], vcells[p][maxv], nvimg[p]  Patch::m_vimages / m_vgrids                  (patch.hpp:41-42)
//       state[p]  1 = registered in the grids, 0 = removed;  birth[p] = creation sequence number
//   cell c of view v (global index cell_base[v] + iy * gw + ix):
//       cslots[c][cell_cap], ccount[c]   PatchManager::m_pgrids AND m_vpgrids in one slot array: entry = patch id,
//                                        bit 31 set for an m_vpgrids entry; SLOT_TOMB = erased (patch_manager.hpp:89-96)
//       dmap[c]                          PatchManager::m_dpgrids as (ordered depth bits << 32 | patch id), ~0 = m_MAXDEPTH
// A cell's std::vector order in the reference is the order of the addPatch calls, i.e. ascending creation order, so
// it is recovered from birth[] wherever it matters (sortPatches ties, collectPatches ids).
#pragma once

#include "pmk_cand.cuh"

namespace pmk {

constexpr int SLOT_TOMB = 0x7fffffff;
constexpr unsigned SLOT_V = 0x80000000u;
constexpr int NEW_MAX = 32;            // new patches one dest cell can stage in one wavefront step (2 per call, 16 calls)
constexpr int SRC_MAX = 32;            // source patches feeding one dest cell (two cells of <= MAX_NUM_OF_PATCHES)
constexpr int NB_CAP = 4096;           // findNeighbors scratch per warp (2 * NB_CAP ints: list + sort buffer)

enum StoreCounter { SC_N = 0, SC_BIRTH = 1, SC_OVERFLOW = 2, SC_FULL = 3, SC_REM = 4, SC_NBOVER = 5, SC_NEXT = 6, SC_MSGOVER = 7, SC_COUNT = 16 };   // SC_NEXT: work queue of a sweep step

struct StoreDev {
    int cap, stage_cap, maxv, cell_cap, total_cells;
    float4* coord; float4* normal; float4* scal;
    int* nimg; int* images; int* cells;
    int* nvimg; int* vimages; int* vcells;
    int* state;
    unsigned int* birth;
    int* counters;                 // StoreCounter
    const int* cell_base;          // [nviews + 1]
    int* ccount; int* cslots;
    unsigned long long* dmap;
};

struct StoreParams {
    CandParams cp;
    StoreDev st;
    float neighbor_cos;            // cosf(120.0f / M_PI * 180.0f): the value PmMvps::isNeighbor really tests (pmmvps.cpp:124)
    double neighbor_radius_cos;    // cos(120.0f * M_PI / 180.0f), double compare                          (pmmvps.cpp:150)
    float neighbor_threshold, neighbor_threshold1, neighbor_threshold2, quad_threshold;
    int max_patches_cell;          // Propagate::MAX_NUM_OF_PATCHES = 2 * csize * csize                       (propagate.cpp:24-25)
    int* nb_scratch;               // per warp: NB_CAP ints
};

__device__ __forceinline__ int pack_cell(int ix, int iy) { return (ix & 0xffff) | (iy << 16); }
__device__ __forceinline__ int cell_x(int c) { return c & 0xffff; }
__device__ __forceinline__ int cell_y(int c) { return (int)((unsigned)c >> 16); }
__device__ __forceinline__ int cell_global(const StoreParams& sp, int img, int ix, int iy) {
    return sp.st.cell_base[img] + iy * sp.cp.p.views[img].gw + ix;
}
__device__ __forceinline__ V4 f4v(float4 a) { return V4{a.x, a.y, a.z, a.w}; }
__device__ __forceinline__ float4 v4f(V4 a) { return make_float4(a.x, a.y, a.z, a.w); }

// order-preserving float -> uint (for atomicMin on depths)
__device__ __forceinline__ unsigned int ford(float f) { const unsigned int u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ unsigned long long depth_key(float depth, int id) { return ((unsigned long long)ford(depth) << 32) | (unsigned int)id; }

// the geometry isNeighbor needs from a patch
struct PGeo { V4 X, N; float dscale; int ref; };
__device__ __forceinline__ PGeo load_geo(const StoreDev& st, int id) {
    PGeo g;
    g.X = f4v(st.coord[id]); g.N = f4v(st.normal[id]);
    g.dscale = st.scal[id].y;
    g.ref = st.images[(size_t)id * st.maxv];
    return g;
}

// ---- PmMvps::isNeighbor (pmmvps.cpp:117-147), decision arithmetic in the reference's order -----------------------
__device__ __forceinline__ int is_neighbor_h(const StoreParams& sp, const PGeo& l, const PGeo& r, float hunit, float thr) {
    if (dot4(l.N, r.N) < sp.neighbor_cos) return 0;
    const V4 diff = sub4(l.X, r.X);
    const float vunit = xadd(l.dscale, r.dscale);
    const float f0 = dot4(l.N, diff), f1 = dot4(r.N, diff);
    float ftmp = xdiv(xadd(fabsf(f0), fabsf(f1)), 2.0f);
    ftmp = xdiv(ftmp, vunit);
    // (diff - f0 * lN + diff - f1 * rN).norm() / 2.0f / hunit
    const V4 h = sub4(add4(sub4(diff, mul4(l.N, f0)), diff), mul4(r.N, f1));
    const float hsize = xdiv(xdiv(norm4(h), 2.0f), hunit);
    if (1.0f < hsize) ftmp = xdiv(ftmp, min_std(2.0f, hsize));
    return ftmp < thr ? 1 : 0;
}
__device__ __forceinline__ int is_neighbor(const StoreParams& sp, const PGeo& l, const PGeo& r, float thr) {
    const Params& p = sp.cp.p;
    const float hunit = xmul(xdiv(xadd(get_unit(p.views[l.ref], l.X, p.level_scale), get_unit(p.views[r.ref], r.X, p.level_scale)), 2.0f), (float)p.csize);
    return is_neighbor_h(sp, l, r, hunit, thr);
}

// ---- PmMvps::isNeighborRadius (pmmvps.cpp:149-180) ------------------------------------------------------------------
__device__ __forceinline__ int is_neighbor_radius(const StoreParams& sp, const PGeo& l, const PGeo& r, float hunit, float thr, float radius) {
    if ((double)dot4(l.N, r.N) < sp.neighbor_radius_cos) return 0;
    const V4 diff = sub4(r.X, l.X);
    const float vunit = xadd(l.dscale, r.dscale);
    const float f0 = dot4(l.N, diff), f1 = dot4(r.N, diff);
    float ftmp = xdiv(xadd(fabsf(f0), fabsf(f1)), 2.0f);
    ftmp = xdiv(ftmp, vunit);
    const V4 h = sub4(sub4(mul4(diff, 2.0f), mul4(l.N, f0)), mul4(r.N, f1));
    const float hsize = xdiv(xdiv(norm4(h), 2.0f), hunit);
    if (xdiv(radius, hunit) < hsize) return 0;
    if (1.0f < hsize) ftmp = xdiv(ftmp, min_std(2.0f, hsize));
    return ftmp < thr ? 1 : 0;
}

// ---- PatchManager::isVisible (patch_manager.cpp:335-376) ---------------------------------------------------------
__device__ __forceinline__ int is_visible(const StoreParams& sp, V4 X, V4 N, int img, int ix, int iy, float strict) {
    const Params& p = sp.cp.p;
    const ViewConst& vc = p.views[img];
    if (ix < 0 || vc.gw <= ix || iy < 0 || vc.gh <= iy) return 0;
    if (p.depth == 0) return 1;
    const unsigned long long k = sp.st.dmap[cell_global(sp, img, ix, iy)];
    if (k == ~0ull) return 1;
    const V4 Xd = f4v(sp.st.coord[(int)(k & 0xffffffffu)]);
    V4 ray = sub4(X, ld4(vc.center));
    ray = div4(ray, norm4(ray));
    const float diff = dot4(ray, sub4(X, Xd));
    const float factor = __double2float_rn(fmin(2.0, 2.0 + (double)dot4(ray, N)));
    const float lim = xmul(xmul(xmul(get_unit(vc, X, p.level_scale), (float)p.csize), strict), factor);
    return diff < lim ? 1 : 0;
}
// PatchManager::isVisible0 (:327-333)
__device__ __forceinline__ int is_visible0(const StoreParams& sp, V4 X, V4 N, int img, int& ix, int& iy, float strict) {
    const Params& p = sp.cp.p;
    const V3 ic = project(p.views[img].P, X);
    ix = cell_of(ic.x, p.csize); iy = cell_of(ic.y, p.csize);
    return is_visible(sp, X, N, img, ix, iy, strict);
}

// ---- PatchManager::updateDepthMaps (patch_manager.cpp:191-221) / Filter::setDepthMapsSub (filter.cpp:587-626) for one view --------
__device__ __forceinline__ void update_depth_map(const StoreParams& sp, int id, V4 X, int img) {
    const Params& p = sp.cp.p;
    const ViewConst& vc = p.views[img];
    const V3 ic = project(vc.P, X);
    const float fx = xdiv(ic.x, (float)p.csize), fy = xdiv(ic.y, (float)p.csize);
    const int xs[2] = {(int)floorf(fx), (int)ceilf(fx)}, ys[2] = {(int)floorf(fy), (int)ceilf(fy)};
    const float depth = dot4(ld4(vc.oaxis), X);
    const unsigned long long key = depth_key(depth, id);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            if (xs[i] < 0 || vc.gw <= xs[i] || ys[j] < 0 || vc.gh <= ys[j]) continue;
            atomicMin(sp.st.dmap + cell_global(sp, img, xs[i], ys[j]), key);
        }
}

// ---- PatchManager::setVImagesVGrids (patch_manager.cpp:267-301): views outside m_images / m_vimages that pass isVisible0 ----------
// On entry the candidate's m_images are ws.images[0..nv) and `vimg/vcell` hold nvv existing entries; appends in view order.
__device__ __forceinline__ int warp_set_vimages(const StoreParams& sp, WarpScratch& ws, V4 X, V4 N, const int* images, int nv, int* vimg, int* vcell, int nvv, int lane) {
    const Params& p = sp.cp.p;
    for (int v = lane; v < p.nviews; v += 32) ws.mark[v] = 0;
    __syncwarp();
    for (int i = lane; i < nv; i += 32) ws.mark[images[i]] = 1;
    for (int i = lane; i < nvv; i += 32) ws.mark[vimg[i]] = 1;
    __syncwarp();
    for (int base = 0; base < p.nviews; base += 32) {
        const int v = base + lane;
        int ix = 0, iy = 0;
        bool add = false;
        if (v < p.nviews && !ws.mark[v]) add = is_visible0(sp, X, N, v, ix, iy, sp.neighbor_threshold) != 0;
        const unsigned m = __ballot_sync(0xffffffffu, add);
        if (add) {
            const int pos = nvv + __popc(m & ((1u << lane) - 1u));
            if (pos < sp.st.maxv) { vimg[pos] = v; vcell[pos] = pack_cell(ix, iy); }
        }
        nvv = min(sp.st.maxv, nvv + __popc(m));
    }
    __syncwarp();
    return nvv;
}

// ---- a view of one patch's lists for the gain / neighbour tests ---------------------------------------------------------------------
struct PatchLists {
    const int* images; const int* cells; int nimg;
    const int* vimages; const int* vcells; int nvimg;
};

// Dest-cell overlay of the sweep: while a warp works on dest cell `cell` its list lives in shared memory and replaces the
// (stale) global slots of that cell; `removed` are patches the warp has already erased this step.
struct Overlay {
    int cell;                 // global cell index the overlay replaces, or -1
    const int* ids; int n;    // current m_pgrids content of that cell
    const int* removed; int nremoved;
};

__device__ __forceinline__ bool overlay_removed(const Overlay& ov, int id) {
    for (int i = 0; i < ov.nremoved; ++i) if (ov.removed[i] == id) return true;
    return false;
}

// max over the m_pgrids entries q of cell c of (q.ncc - thr) subject to `front` (pdepth < bdepth, vimages only) and !isNeighbor;
// one LANE walks the whole cell (the warp spreads over the patch's registrations)
__device__ __forceinline__ float lane_cell_pressure(const StoreParams& sp, const PGeo& me, int img, int c, bool need_front, float pdepth, const Overlay& ov) {
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    float mp = 0.0f;
    const bool local = (c == ov.cell);
    const int n = local ? ov.n : min(st.ccount[c], st.cell_cap);
    for (int s = 0; s < n; ++s) {
        const int e = local ? ov.ids[s] : st.cslots[(size_t)c * st.cell_cap + s];
        if (e == SLOT_TOMB || e < 0 || (!local && overlay_removed(ov, e))) continue;
        const PGeo q = load_geo(st, e);
        if (need_front && !(pdepth < dot4(ld4(p.views[img].oaxis), q.X))) continue;            // Camera::computeDepth (camera.cpp:339-346)
        if (!is_neighbor(sp, me, q, sp.neighbor_threshold1)) mp = max_std(mp, xsub(st.scal[e].x, p.ncc_threshold));
    }
    return mp;
}

// ---- Filter::computeGain (filter.cpp:108-146) ----------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_compute_gain(const StoreParams& sp, const PGeo& me, float ncc, const PatchLists& pl, const Overlay& ov, int lane) {
    const Params& p = sp.cp.p;
    float gain = xmul(max_std(0.0f, xsub(ncc, p.ncc_threshold)), (float)pl.nimg);              // score2 (patch.cpp:27-29)
    const int tot = pl.nimg + pl.nvimg;
    for (int base = 0; base < tot; base += 32) {
        const int i = base + lane;
        float mp = 0.0f;
        if (i < tot) {
            const bool isv = i >= pl.nimg;
            const int img = isv ? pl.vimages[i - pl.nimg] : pl.images[i], c = isv ? pl.vcells[i - pl.nimg] : pl.cells[i];
            const float pdepth = isv ? dot4(ld4(p.views[img].oaxis), me.X) : 0.0f;
            mp = lane_cell_pressure(sp, me, img, cell_global(sp, img, cell_x(c), cell_y(c)), isv, pdepth, ov);
        }
        const int m = min(32, tot - base);
        for (int k = 0; k < m; ++k) gain = xsub(gain, __shfl_sync(0xffffffffu, mp, k));          // in list order, like the reference
    }
    return gain;
}

// ---- PatchManager::findNeighbors (patch_manager.cpp:671-728): unique patches of the +-margin cells of every view of m_images
// (m_pgrids and m_vpgrids) that pass isNeighborRadius.  Ids go to `out` in ascending order; returns the count.
// Per view the (2 margin + 1)^2 cells are flattened into one slot range spread over the lanes (binary search on a prefix held in
// the lanes); a warp-private open-addressing table of (call token, id) words in global scratch makes every patch be tested once,
// however many views' cells it shows up in (sort + unique by pointer in the reference).
constexpr int NB_HASH = 16384;         // table entries per warp (power of two): every patch id SEEN in the searched cells goes in
constexpr int NB_STRIDE = 2 * NB_CAP + 2 * NB_HASH + 4;   // ints of findNeighbors scratch per warp

__device__ __forceinline__ bool nb_insert(unsigned long long* table, unsigned int token, int id, bool& full) {
    unsigned int h = ((unsigned int)id * 2654435761u) >> 18;                  // 14 bits
    const unsigned long long mine = ((unsigned long long)token << 32) | (unsigned int)id;
    for (int probe = 0; probe < NB_HASH; ++probe) {
        const unsigned long long cur = __ldcg(table + h);                    // L2: the table is written with atomics
        if ((unsigned int)(cur >> 32) == token) {
            if ((unsigned int)cur == (unsigned int)id) return false;
        } else {
            const unsigned long long old = atomicCAS(table + h, cur, mine);
            if (old == cur) return true;
            if (old == mine) return false;
            continue;                                                          // another lane claimed the slot: look at it again
        }
        h = (h + 1) & (NB_HASH - 1);
    }
    full = true;                                                               // reported by the caller, never silent
    return false;
}

__device__ __forceinline__ int warp_find_neighbors(const StoreParams& sp, const PGeo& me, const PatchLists& pl, float scale, int margin,
                                                   const Overlay& ov, int* out, int lane) {
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    // Propagate::computeRadius (propagate.cpp:474-481): second smallest computeUnits entry * csize
    float u1 = __int_as_float(0x7f800000), u2 = u1;
    float usum = 0.0f;
    for (int i = 0; i < pl.nimg; ++i) {
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
    // scratch of this warp: [0, NB_CAP) list, [NB_CAP, 2 NB_CAP) sort buffer, then the table and its call counter
    unsigned long long* table = reinterpret_cast<unsigned long long*>(out + 2 * NB_CAP);
    unsigned int* tokp = reinterpret_cast<unsigned int*>(table + NB_HASH);
    unsigned int token = 0;
    if (lane == 0) { token = *tokp + 1; if (token == 0) token = 1; *tokp = token; }
    token = __shfl_sync(0xffffffffu, token, 0);
    int nuni = 0;
    bool overflow = false;
    const int side = 2 * margin + 1, ncell = side * side;           // <= 25 cells, one per lane
    for (int i = 0; i < pl.nimg; ++i) {
        const int img = pl.images[i];
        const ViewConst& vc = p.views[img];
        const int ix = cell_x(pl.cells[i]), iy = cell_y(pl.cells[i]);
        int myc = -1, myn = 0, mynslot = 0;
        if (lane < ncell) {
            const int yt = iy + lane / side - margin, xt = ix + lane % side - margin;
            if (!(yt < 0 || vc.gh <= yt || xt < 0 || vc.gw <= xt)) {
                myc = cell_global(sp, img, xt, yt);
                mynslot = min(st.ccount[myc], st.cell_cap);
                myn = mynslot + (myc == ov.cell ? ov.n : 0);       // the overlay supplies the m_pgrids entries, the slots the m_vpgrids ones
            }
        }
        int incl = myn;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        for (int base = 0; base < total; base += 32) {
            const int f = min(base + lane, total - 1);
            // which cell does flat index f fall into: binary search on the inclusive prefix held by the lanes
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
            // a patch shows up in the cells of every view it is registered in: look at it once (the test does not depend on the view)
            bool hit = false;
            if (id >= 0 && nb_insert(table, token, id, overflow)) hit = is_neighbor_radius(sp, me, load_geo(st, id), unit, thr, radius) != 0;
            const unsigned m = __ballot_sync(0xffffffffu, hit);
            if (hit) { const int pos = nuni + __popc(m & ((1u << lane) - 1u)); if (pos < NB_CAP) out[pos] = id; }
            nuni += __popc(m);
        }
    }
    __syncwarp();
    if (nuni > NB_CAP) { overflow = true; nuni = NB_CAP; }
    if (__any_sync(0xffffffffu, overflow) && lane == 0) atomicAdd(st.counters + SC_NBOVER, 1);
    // ascending id order (ids follow creation order, identical on every replica and for every number of GPUs): the quadric fit
    // sums over the neighbours and must do so in one order everywhere
    int* tmp = out + NB_CAP;
    for (int k = lane; k < nuni; k += 32) {
        const int id = out[k];
        int rank = 0;
        for (int j = 0; j < nuni; ++j) rank += out[j] < id ? 1 : 0;
        tmp[rank] = id;
    }
    __syncwarp();
    for (int k = lane; k < nuni; k += 32) out[k] = tmp[k];
    __syncwarp();
    return nuni;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- Filter::filterQuad (filter.cpp:338-409) + ortho (:411-425) + lls (:427-446) ----------------------------------------------------
// The reference solves the n x 5 least-squares problem with Eigen's JacobiSVD (third party, absent); here: normal equations
// in double with partial pivoting.  The decision `residual < m_quadThreshold` is tolerance-bound (summation order in the
// reference follows heap addresses).  Returns 1 = reject.
__device__ __forceinline__ int warp_filter_quad(const StoreParams& sp, const PGeo& me, const PatchLists& pl, const int* nb, int n, float* residual_out, int lane) {
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    V4 xd, yd;
    const V4 z = me.N;
    if (fabsf(z.x) > 0.5f) xd = V4{z.y, -z.x, 0.f, 0.f};
    else if (fabsf(z.y) > 0.5f) xd = V4{0.f, z.z, -z.y, 0.f};
    else xd = V4{-z.z, 0.f, z.x, 0.f};
    xd = div4(xd, norm4(xd));
    yd = V4{xsub(xmul(z.y, xd.z), xmul(z.z, xd.y)), xsub(xmul(z.z, xd.x), xmul(z.x, xd.z)), xsub(xmul(z.x, xd.y), xmul(z.y, xd.x)), 0.f};
    double hs = 0.0;
    for (int k = lane; k < n; k += 32) hs += (double)norm4(sub4(f4v(st.coord[nb[k]]), me.X));
    const float h = (float)(warp_sum_d(hs) / (double)n);
    // accumulate A^T A (15 unique entries) and A^T b
    double ata[15], atb[5];
#pragma unroll
    for (int i = 0; i < 15; ++i) ata[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) atb[i] = 0.0;
    for (int k = lane; k < n; k += 32) {
        const V4 diff = sub4(f4v(st.coord[nb[k]]), me.X);
        const float fx = xdiv(dot4(diff, xd), h), fy = xdiv(dot4(diff, yd), h), fz = dot4(diff, me.N);
        const double a[5] = {(double)xmul(fx, fx), (double)xmul(fy, fy), (double)xmul(fx, fy), (double)fx, (double)fy};
        int q = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = i; j < 5; ++j) ata[q++] += a[i] * a[j];
            atb[i] += a[i] * (double)fz;
        }
    }
#pragma unroll
    for (int i = 0; i < 15; ++i) ata[i] = warp_sum_d(ata[i]);
#pragma unroll
    for (int i = 0; i < 5; ++i) atb[i] = warp_sum_d(atb[i]);
    double M[5][6];
    {
        int q = 0;
        for (int i = 0; i < 5; ++i) for (int j = i; j < 5; ++j) { M[i][j] = ata[q]; M[j][i] = ata[q]; ++q; }
        for (int i = 0; i < 5; ++i) M[i][5] = atb[i];
    }
    double x[5] = {0, 0, 0, 0, 0};
    bool singular = false;
    for (int c = 0; c < 5; ++c) {
        int piv = c;
        for (int r = c + 1; r < 5; ++r) if (fabs(M[r][c]) > fabs(M[piv][c])) piv = r;
        if (fabs(M[piv][c]) < 1e-30) { singular = true; break; }
        if (piv != c) for (int j = 0; j < 6; ++j) { const double t = M[c][j]; M[c][j] = M[piv][j]; M[piv][j] = t; }
        for (int r = c + 1; r < 5; ++r) {
            const double f = M[r][c] / M[c][c];
            for (int j = c; j < 6; ++j) M[r][j] -= f * M[c][j];
        }
    }
    if (!singular)
        for (int r = 4; r >= 0; --r) {
            double s = M[r][5];
            for (int j = r + 1; j < 5; ++j) s -= M[r][j] * x[j];
            x[r] = s / M[r][r];
        }
    const int inum = min(p.tau, pl.nimg);
    float unit = 0.0f;
    for (int i = 0; i < inum; ++i) unit = xadd(unit, get_unit(p.views[pl.images[i]], me.X, p.level_scale));
    unit = xdiv(unit, (float)inum);
    double rs = 0.0;
    for (int k = lane; k < n; k += 32) {
        const V4 diff = sub4(f4v(st.coord[nb[k]]), me.X);
        const float fx = xdiv(dot4(diff, xd), h), fy = xdiv(dot4(diff, yd), h), fz = dot4(diff, me.N);
        const float res = (float)(x[0] * (double)xmul(fx, fx) + x[1] * (double)xmul(fy, fy) + x[2] * (double)xmul(fx, fy) + x[3] * (double)fx + x[4] * (double)fy - (double)fz);
        rs += (double)xdiv(fabsf(res), unit);
    }
    const float residual = (float)(warp_sum_d(rs) / (double)(n - 5));
    if (residual_out) *residual_out = residual;
    return residual < sp.quad_threshold ? 0 : 1;
}

}  // namespace pmk
