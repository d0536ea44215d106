// mvskit_b200/csrc/pmk_filter.cuh -- K5..K10: the store rebuild and the four filters of Filter::run (pmmvps/filter.cpp:25-49).
//
//   rebuild  = PatchManager::collectPatches (patch_manager.cpp:75-104) + Filter::setDepthMapsVGridsVPGridsAddPatchV (filter.cpp:628-655):
//              K5a collect keys -> radix sort -> K5b gather (ids become the reference's m_ppatches indices), K5c register m_pgrids,
//              K5d depth maps (atomicMin on depth|id: nearest patch, first in collect order on ties), K5e setVImagesVGrids + m_vpgrids
//   K6  filterOutside  (filter.cpp:51-106)   computeGain for every patch against one snapshot, then remove gain < 0
//   K7  filterExact    (filter.cpp:148-263)  per registration: visible in its cell or a 4-neighbour; rebuild m_images in view order,
//                                            setRefImage (pairwise INCC), setGrids; drop below minImageNum
//   K8  filterNeighbor (filter.cpp:265-336)  < 6 neighbours or quadric residual >= m_quadThreshold
//   K9  filterSmallGroups (filter.cpp:432-578) directed isNeighbor edges in the +-1 cells of the reference view (device),
//                                            breadth-first labelling in m_ppatches order (host, order-dependent in the reference)
#pragma once

#include "pmk_cell.cuh"

namespace pmk {

// ---- K5a: collect key = first appearance in the reference's scan (view, cell, creation order) ------------------------------------
__global__ void k5_collect_keys(const StoreParams sp, int n, unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const StoreDev& st = sp.st;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    unsigned long long key = ~0ull;
    if (st.state[p] == 1) {
        const int ni = st.nimg[p];
        for (int i = 0; i < ni; ++i) {
            const int img = st.images[(size_t)p * st.maxv + i], c = st.cells[(size_t)p * st.maxv + i];
            const unsigned long long k = ((unsigned long long)img << 56) | ((unsigned long long)(cell_y(c) * sp.cp.p.views[img].gw + cell_x(c)) << 32) | st.birth[p];
            key = k < key ? k : key;
        }
        if (ni == 0) key = ~0ull - 1;           // registered nowhere: the reference can no longer reach it
    }
    keys[p] = key;
    vals[p] = p;
}

__global__ void k5_count_alive(const unsigned long long* __restrict__ keys, int n, int* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    // keys are sorted ascending: the alive prefix ends where the key reaches the sentinels
    const bool alive = keys[p] < ~0ull - 1;
    const bool next_alive = (p + 1 < n) && keys[p + 1] < ~0ull - 1;
    if (alive && !next_alive) *out = p + 1;
}

// ---- K5b: the gather into collect order is k_gather_rows (pmk_store_host.cuh): one array at a time through a scratch buffer ----

// ---- K5c: m_pgrids from the patch lists ----------------------------------------------------------------------------------------------
__global__ void k5_register(const StoreParams sp, int n, int first, int with_v, int with_depth) {
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int q = first + gwarp; q < n; q += nwarps) {
        if (sp.st.state[q] != 1) continue;
        warp_register_patch(sp, q, with_v != 0, with_depth != 0, lane);
    }
}

// ---- K5d: Filter::setDepthMapsSub (filter.cpp:587-626), one thread per (patch, view) -------------------------------------------------
__global__ void k5_depth_maps(const StoreParams sp, int n) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nviews = sp.cp.p.nviews;
    if (t >= (long long)n * nviews) return;
    const int q = (int)(t / nviews), v = (int)(t % nviews);
    if (sp.st.state[q] != 1) return;
    update_depth_map(sp, q, f4v(sp.st.coord[q]), v);
}

// ---- K5e: Filter::setVGridsVPGrids + addPatchV (filter.cpp:657-687) ------------------------------------------------------------------
__global__ void __launch_bounds__(CAND_WARPS * 32) k5_set_vimages(const StoreParams sp, int n, int additive) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& ws = warp_scratch(smem_raw);
    const StoreDev& st = sp.st;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    for (int q = gwarp; q < n; q += gridDim.x * CAND_WARPS) {
        if (st.state[q] != 1) continue;
        const int nvv0 = additive ? st.nvimg[q] : 0;
        const int nvv = warp_set_vimages(sp, ws, f4v(st.coord[q]), f4v(st.normal[q]), st.images + (size_t)q * st.maxv, st.nimg[q],
                                         st.vimages + (size_t)q * st.maxv, st.vcells + (size_t)q * st.maxv, nvv0, lane);
        if (lane == 0) st.nvimg[q] = nvv;
        __syncwarp();
        for (int i = lane; i < nvv; i += 32) {
            const int img = st.vimages[(size_t)q * st.maxv + i], c = st.vcells[(size_t)q * st.maxv + i];
            insert_into_cell(st, cell_global(sp, img, cell_x(c), cell_y(c)), (int)((unsigned)q | SLOT_V));
        }
        __syncwarp();
    }
}

__device__ __forceinline__ PatchLists lists_of(const StoreDev& st, int q) {
    return PatchLists{st.images + (size_t)q * st.maxv, st.cells + (size_t)q * st.maxv, st.nimg[q],
                      st.vimages + (size_t)q * st.maxv, st.vcells + (size_t)q * st.maxv, st.nvimg[q]};
}

// ---- K6: Filter::filterOutsideSub (filter.cpp:98-106) ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(CAND_WARPS * 32) k6_gains(const StoreParams sp, int n, float* __restrict__ gains) {
    const StoreDev& st = sp.st;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    const Overlay none{-1, nullptr, 0, nullptr, 0};
    for (int q = gwarp; q < n; q += gridDim.x * CAND_WARPS) {
        if (st.state[q] != 1) { if (lane == 0) gains[q] = 0.0f; continue; }
        const float g = warp_compute_gain(sp, load_geo(st, q), st.scal[q].x, lists_of(st, q), none, lane);
        if (lane == 0) gains[q] = g;
    }
}

// remove[q] != 0 -> the patch leaves the store at the next rebuild (PatchManager::removePatch; the grids are rebuilt from
// the patch lists before anything reads them again, filter.cpp:29-48)
__global__ void k_kill_flagged(const StoreDev st, int n, const float* __restrict__ gains, const int* __restrict__ flags, int* __restrict__ killed) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n || st.state[q] != 1) return;
    const bool kill = gains ? gains[q] < 0.0f : flags[q] != 0;
    if (kill) { st.state[q] = 0; atomicAdd(killed, 1); }
}

// ---- K7: Filter::filterExact (filter.cpp:148-263) -------------------------------------------------------------------------------------
template <int WS>
__global__ void __launch_bounds__(CAND_WARPS * 32) k7_exact(const StoreParams sp, int n, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& ws = warp_scratch(smem_raw);
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    for (int q = gwarp; q < n; q += gridDim.x * CAND_WARPS) {
        if (st.state[q] != 1) { if (lane == 0) flags[q] = 0; continue; }
        const V4 X = f4v(st.coord[q]), N = f4v(st.normal[q]);
        const int ni = min(st.nimg[q], CAND_MAXV);
        const float thr = sp.neighbor_threshold1;
        // safe[i] (filter.cpp:226-246): visible in its own cell or in one of the 4 neighbours
        for (int i = lane; i < ni; i += 32) {
            const int img = st.images[(size_t)q * st.maxv + i], c = st.cells[(size_t)q * st.maxv + i];
            const int x = cell_x(c), y = cell_y(c), w = p.views[img].gw, h = p.views[img].gh;
            int safe = 0;
            if (is_visible(sp, X, N, img, x, y, thr)) safe = 1;
            else if (0 < x && is_visible(sp, X, N, img, x - 1, y, thr)) safe = 1;
            else if (x < w - 1 && is_visible(sp, X, N, img, x + 1, y, thr)) safe = 1;
            else if (0 < y && is_visible(sp, X, N, img, x, y - 1, thr)) safe = 1;
            else if (y < h - 1 && is_visible(sp, X, N, img, x, y + 1, thr)) safe = 1;
            ws.idx[i] = img;
            ws.alive[i] = (unsigned char)safe;
        }
        __syncwarp();
        // m_newimages: the safe views in ascending view order (the outer loop of filterExactSub runs over views)
        int cnt = 0;
        for (int i = lane; i < ni; i += 32) {
            if (!ws.alive[i]) continue;
            int rank = 0;
            for (int j = 0; j < ni; ++j) if (ws.alive[j] && (ws.idx[j] < ws.idx[i] || (ws.idx[j] == ws.idx[i] && j < i))) ++rank;
            ws.images[rank] = ws.idx[i];
        }
        for (int i = 0; i < ni; ++i) cnt += ws.alive[i] ? 1 : 0;
        __syncwarp();
        int kill = 0;
        if (p.min_image_num <= cnt) {
            warp_set_ref_image<WS>(sp.cp, ws, X, N, cnt, gwarp, lane);                           // :252
        } else kill = 1;                                                                        // :256-259
        for (int i = lane; i < cnt; i += 32) {                                                  // setGrids (:253)
            const int img = ws.images[i];
            const V3 ic = project(p.views[img].P, X);
            st.images[(size_t)q * st.maxv + i] = img;
            st.cells[(size_t)q * st.maxv + i] = pack_cell(cell_of(ic.x, p.csize), cell_of(ic.y, p.csize));
        }
        if (lane == 0) { st.nimg[q] = cnt; flags[q] = kill; }
        __syncwarp();
    }
}

// ---- K8: Filter::filterNeighborSub (filter.cpp:316-336) -------------------------------------------------------------------------------
__global__ void __launch_bounds__(CAND_WARPS * 32) k8_neighbor(const StoreParams sp, int n, int* __restrict__ flags, int* __restrict__ nn_out, float* __restrict__ res_out) {
    const StoreDev& st = sp.st;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    const Overlay none{-1, nullptr, 0, nullptr, 0};
    int* nb = sp.nb_scratch + (size_t)gwarp * NB_STRIDE;
    for (int q = gwarp; q < n; q += gridDim.x * CAND_WARPS) {
        if (st.state[q] != 1) { if (lane == 0) { flags[q] = 0; if (nn_out) nn_out[q] = 0; if (res_out) res_out[q] = 0.0f; } continue; }
        const PGeo me = load_geo(st, q);
        const PatchLists pl = lists_of(st, q);
        const int nn = warp_find_neighbors(sp, me, pl, 4.0f, 2, none, nb, lane);
        int rej = 0;
        float res = -1.0f;
        if (nn < 6) rej = 1;
        else if (warp_filter_quad(sp, me, pl, nb, nn, &res, lane)) rej = 1;
        if (lane == 0) { flags[q] = rej; if (nn_out) nn_out[q] = nn; if (res_out) res_out[q] = res; }
        __syncwarp();
    }
}

// ---- K9: directed isNeighbor edges for Filter::filterSmallGroupsSub (filter.cpp:527-578).  pass 0 counts, pass 1 fills. -----------------
__global__ void __launch_bounds__(CAND_WARPS * 32) k9_group_edges(const StoreParams sp, int n, int pass, int* __restrict__ deg, const int* __restrict__ offs, int* __restrict__ adj) {
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    for (int q = gwarp; q < n; q += gridDim.x * CAND_WARPS) {
        if (st.state[q] != 1) { if (pass == 0 && lane == 0) deg[q] = 0; continue; }
        const PGeo me = load_geo(st, q);
        const int img = me.ref, c0 = st.cells[(size_t)q * st.maxv];
        const int ix = cell_x(c0), iy = cell_y(c0);
        const ViewConst& vc = p.views[img];
        int cnt = 0;
        const int o0 = pass ? offs[q] : 0;
        for (int w = 0; w < 9; ++w) {
            const int yt = iy + w / 3 - 1, xt = ix + w % 3 - 1;
            if (yt < 0 || vc.gh <= yt || xt < 0 || vc.gw <= xt) continue;
            const int c = cell_global(sp, img, xt, yt);
            const int m = min(st.ccount[c], st.cell_cap);
            for (int base = 0; base < m; base += 32) {
                const int s = base + lane;
                bool hit = false;
                int id = 0;
                if (s < m) {
                    const int e = st.cslots[(size_t)c * st.cell_cap + s];
                    if ((e & 0x7fffffff) != SLOT_TOMB) {
                        id = e & 0x7fffffff;
                        if (id != q && st.state[id] == 1) hit = is_neighbor(sp, me, load_geo(st, id), sp.neighbor_threshold2) != 0;
                    }
                }
                const unsigned msk = __ballot_sync(0xffffffffu, hit);
                if (pass && hit) adj[o0 + cnt + __popc(msk & ((1u << lane) - 1u))] = id;
                cnt += __popc(msk);
            }
        }
        if (pass == 0 && lane == 0) deg[q] = cnt;
        __syncwarp();
    }
}

// ---- Filter::filterSmallGroups' labelling (filter.cpp:455-473) on the device ------------------------------------------------------------
// The reference labels breadth-first in m_ppatches order over a DIRECTED relation (a patch's neighbours are looked up in the grid of
// ITS reference view only): for pid = 0, 1, ...: if unlabelled, start a group and claim everything reachable that is still unlabelled.
// Reachability is transitive, so the patch that ends up labelling v is m(v) = the smallest id among all patches that can reach v
// (v included): nobody smaller reaches m(v), hence it is still unlabelled at its turn and starts a group; no earlier group start
// reaches v.  label(v) = min over ancestors is the fixed point of  L[v] = min(L[v], L[u]) over edges u -> v,  which shortcutting
// L[v] = L[L[v]] (valid because L[v] reaches v and L[L[v]] reaches L[v]) brings down to a few rounds.  Groups, sizes and therefore
// the removals are exactly the reference's; only the group numbers differ (seed ids instead of consecutive integers).
__global__ void k9_label_init(int n, int* __restrict__ L) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) L[q] = q;
}
// one warp per source patch u: push L[u] along its out-edges
__global__ void k9_label_relax(int n, const int* __restrict__ offs, const int* __restrict__ adj, int* __restrict__ L, int* __restrict__ changed) {
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    bool any = false;
    for (int u = gwarp; u < n; u += nwarps) {
        const int lu = L[u];
        for (int k = offs[u] + lane; k < offs[u + 1]; k += 32) {
            const int v = adj[k];
            if (lu < L[v]) { atomicMin(L + v, lu); any = true; }
        }
    }
    if (__any_sync(0xffffffffu, any) && lane == 0) *changed = 1;
}
__global__ void k9_label_jump(int n, int* __restrict__ L, int* __restrict__ changed) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    int l = L[q];
    int ll = L[l];
    if (ll < l) {
        while (ll < l) { l = ll; ll = L[l]; }
        L[q] = l;
        *changed = 1;
    }
}
__global__ void k9_group_sizes(const StoreDev st, int n, const int* __restrict__ L, int* __restrict__ size) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n && st.state[q] == 1) atomicAdd(size + L[q], 1);
}
__global__ void k9_group_flags(const StoreDev st, int n, const int* __restrict__ L, const int* __restrict__ size, int threshold, int* __restrict__ flags, int* __restrict__ removed) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const int f = (st.state[q] == 1 && size[L[q]] < threshold) ? 1 : 0;
    flags[q] = f;
    if (f) atomicAdd(removed, 1);
}

// PatchManager::writePly colour (patch_manager.cpp:566-581): mean over m_images of Image::getColor at the projection, rounded
__global__ void k_patch_colors(const StoreParams sp, int n, const int* __restrict__ perm, unsigned char* __restrict__ rgb) {
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int q = perm[k];
    const V4 X = f4v(st.coord[q]);
    float r = 0.f, g = 0.f, b = 0.f;
    const int ni = st.nimg[q];
    for (int i = 0; i < ni; ++i) {
        const ViewConst& vc = p.views[st.images[(size_t)q * st.maxv + i]];
        const V3 ic = project(vc.P, X);
        const float x = fminf(fmaxf(ic.x, 0.0f), (float)(vc.w[p.level] - 2)), y = fminf(fmaxf(ic.y, 0.0f), (float)(vc.h[p.level] - 2));
        float cr, cg, cb;
        bilinear(vc.img[p.level], vc.w[p.level], x, y, cr, cg, cb);
        r += cr; g += cg; b += cb;
    }
    const float inv = ni > 0 ? 1.0f / (float)ni : 0.0f;
    rgb[3 * k] = (unsigned char)min(255, (int)floorf(r * inv + 0.5f));
    rgb[3 * k + 1] = (unsigned char)min(255, (int)floorf(g * inv + 0.5f));
    rgb[3 * k + 2] = (unsigned char)min(255, (int)floorf(b * inv + 0.5f));
}

// PatchManager::readPatches body (patch_manager.cpp:450-462) for patches already copied to slots [first, first + n):
// m_tmp = score2(m_nccThreshold), m_vimages cleared, setGrids, addPatch.  One warp per patch.
__global__ void k_store_add(const StoreParams sp, int first, int n, unsigned int birth0) {
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int k = gwarp; k < n; k += nwarps) {
        const int q = first + k;
        const V4 X = f4v(st.coord[q]);
        // setGrids (patch_manager.cpp:241-249).  The reference then indexes m_pgrids with whatever cell comes out; a view whose
        // projection falls outside its grid would be an out-of-bounds write there, so such views are dropped from the list the way
        // setGridsImages does (patch_manager.cpp:223-239), keeping the order of the rest.
        int ni = 0;
        const int ni0 = min(st.nimg[q], st.maxv);
        for (int base = 0; base < ni0; base += 32) {
            const int i = base + lane;
            bool keep = false;
            int v = 0, cell = 0;
            if (i < ni0) {
                v = st.images[(size_t)q * st.maxv + i];
                const V3 ic = project(p.views[v].P, X);
                const int ix = cell_of(ic.x, p.csize), iy = cell_of(ic.y, p.csize);
                keep = 0 <= ix && ix < p.views[v].gw && 0 <= iy && iy < p.views[v].gh;
                cell = pack_cell(ix, iy);
            }
            const unsigned msk = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) { const int pos = ni + __popc(msk & ((1u << lane) - 1u)); st.images[(size_t)q * st.maxv + pos] = v; st.cells[(size_t)q * st.maxv + pos] = cell; }
            ni += __popc(msk);
            __syncwarp();
        }
        if (lane == 0) st.nimg[q] = ni;
        if (lane == 0) {
            float4 sc = st.scal[q];
            sc.w = xmul(max_std(0.0f, xsub(sc.x, p.ncc_threshold)), (float)ni);
            st.scal[q] = sc;
            st.nvimg[q] = 0; st.state[q] = ni > 0 ? 1 : 0; st.birth[q] = birth0 + (unsigned int)k;
        }
        __syncwarp();
        const bool deep = p.depth != 0;
        warp_register_patch(sp, q, deep, deep, lane);
        __syncwarp();
    }
}

// probe of PmMvps::isNeighbor / isNeighborRadius (pmmvps.cpp:117-180) on free-standing patch pairs; rec = coord4, normal4, dscale, ref
__global__ void k_probe_neighbor(const StoreParams sp, int n, const float* __restrict__ lrec, const float* __restrict__ rrec,
                                 const float* __restrict__ hunit, const float* __restrict__ radius, float thr, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PGeo l, r;
    const float* a = lrec + (size_t)i * 10; const float* b = rrec + (size_t)i * 10;
    l.X = V4{a[0], a[1], a[2], a[3]}; l.N = V4{a[4], a[5], a[6], a[7]}; l.dscale = a[8]; l.ref = (int)a[9];
    r.X = V4{b[0], b[1], b[2], b[3]}; r.N = V4{b[4], b[5], b[6], b[7]}; r.dscale = b[8]; r.ref = (int)b[9];
    if (radius) out[i] = is_neighbor_radius(sp, l, r, hunit[i], thr, radius[i]);
    else if (hunit) out[i] = is_neighbor_h(sp, l, r, hunit[i], thr);
    else out[i] = is_neighbor(sp, l, r, thr);
}

// PatchManager::setGrids + setVImagesVGrids + Optim::check (optim.cpp:285-323) on free-standing candidates against the current store:
// ret = check's return (1 = rejected), gain = m_tmp, nn = findNeighbors count, vimages/vcells = the visible lists it used.
__global__ void __launch_bounds__(CAND_WARPS * 32) k_probe_check(const StoreParams sp, int n, const float4* __restrict__ coord, const float4* __restrict__ normal,
                                                                  const float4* __restrict__ scal, const int* __restrict__ images, const int* __restrict__ nimg, int stride,
                                                                  int* __restrict__ ret, float* __restrict__ gain_out, int* __restrict__ nn_out,
                                                                  int* __restrict__ vimages_out, int* __restrict__ nvimg_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& ws = warp_scratch(smem_raw);
    int* cells = reinterpret_cast<int*>(smem_raw + CAND_WARPS * sizeof(WarpScratch)) + (threadIdx.x >> 5) * 3 * CAND_MAXV;
    int* vimg = cells + CAND_MAXV;
    int* vcell = vimg + CAND_MAXV;
    const StoreDev& st = sp.st;
    const Params& p = sp.cp.p;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    const Overlay none{-1, nullptr, 0, nullptr, 0};
    int* nb = sp.nb_scratch + (size_t)gwarp * NB_STRIDE;
    for (int h = gwarp; h < n; h += gridDim.x * CAND_WARPS) {
        const V4 X = f4v(coord[h]), N = f4v(normal[h]);
        const float4 sc = scal[h];
        const int nv = min(min(nimg[h], stride), CAND_MAXV);
        bool outside = false;
        for (int i = lane; i < nv; i += 32) {
            const int v = images[(size_t)h * stride + i];
            ws.images[i] = v;
            const V3 q = project(p.views[v].P, X);
            const int ix = cell_of(q.x, p.csize), iy = cell_of(q.y, p.csize);
            outside |= ix < 0 || p.views[v].gw <= ix || iy < 0 || p.views[v].gh <= iy;
            cells[i] = pack_cell(ix, iy);
        }
        __syncwarp();
        if (__any_sync(0xffffffffu, outside)) {
            // the reference would index m_pgrids out of bounds here (computeGain, filter.cpp:113-118); such a candidate cannot come out
            // of postProcess (addImages keeps only views whose projection lies inside the image): reported as ret = -2
            if (lane == 0) { ret[h] = -2; gain_out[h] = 0.0f; nn_out[h] = -1; nvimg_out[h] = 0; }
            for (int i = lane; i < stride; i += 32) vimages_out[(size_t)h * stride + i] = -1;
            __syncwarp();
            continue;
        }
        const int nvv = warp_set_vimages(sp, ws, X, N, ws.images, nv, vimg, vcell, 0, lane);
        PGeo me; me.X = X; me.N = N; me.dscale = sc.y; me.ref = ws.images[0];
        const PatchLists pl{ws.images, cells, nv, vimg, vcell, nvv};
        const float gain = warp_compute_gain(sp, me, sc.x, pl, none, lane);
        int r = 0, nn = -1;
        if (gain < 0.0f) r = 1;
        else {
            nn = warp_find_neighbors(sp, me, pl, 4.0f, 2, none, nb, lane);
            if (6 < nn && warp_filter_quad(sp, me, pl, nb, nn, nullptr, lane)) r = 1;
        }
        if (lane == 0) { ret[h] = r; gain_out[h] = gain; nn_out[h] = nn; nvimg_out[h] = nvv; }
        for (int i = lane; i < stride; i += 32) vimages_out[(size_t)h * stride + i] = i < nvv ? vimg[i] : -1;
        __syncwarp();
    }
}

// ---- probes behind the host mirror's PatchManager pass-throughs ------------------------------------------------------------------
// PatchManager::isVisible0 / isVisible (patch_manager.cpp:327-376) for free-standing points against the store's depth maps;
// cell_in == nullptr: isVisible0 (the cell is computed from the projection and returned in cell_out)
__global__ void k_probe_visible(const StoreParams sp, int n, const float4* __restrict__ coord, const float4* __restrict__ normal, const int* __restrict__ image,
                                const int* __restrict__ cell_in, float strict, int* __restrict__ out, int* __restrict__ cell_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V4 X = f4v(coord[i]), N = f4v(normal[i]);
    int ix, iy, r;
    if (cell_in) { ix = cell_in[2 * i]; iy = cell_in[2 * i + 1]; r = is_visible(sp, X, N, image[i], ix, iy, strict); }
    else r = is_visible0(sp, X, N, image[i], ix, iy, strict);
    out[i] = r;
    if (cell_out) { cell_out[2 * i] = ix; cell_out[2 * i + 1] = iy; }
}

// PatchManager::setScales (patch_manager.cpp:378-399) for fresh patches (m_dscale starts at 0); one warp per patch
__global__ void __launch_bounds__(CAND_WARPS * 32) k_probe_scales(const CandParams cp, int n, const float4* __restrict__ coord, const int* __restrict__ images,
                                                                  const int* __restrict__ nimg, int stride, float* __restrict__ dscale, float* __restrict__ ascale) {
    __shared__ float tmp[CAND_WARPS][PMK_MAX_TAU + 1];
    __shared__ int img[CAND_WARPS][PMK_MAX_TAU + 1];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int gwarp = blockIdx.x * CAND_WARPS + w;
    for (int h = gwarp; h < n; h += gridDim.x * CAND_WARPS) {
        const int nv = min(nimg[h], stride);
        if (lane < min(nv, PMK_MAX_TAU)) img[w][lane] = images[(size_t)h * stride + lane];
        __syncwarp();
        float ds = 0.0f, as = 0.0f;
        if (nv >= 1) warp_set_scales(cp.p, f4v(coord[h]), img[w], nv, ds, as, tmp[w], lane);
        if (lane == 0) { dscale[h] = ds; ascale[h] = as; }
        __syncwarp();
    }
}

// PatchManager::findNeighbors (patch_manager.cpp:671-728) for free-standing patches {coord, normal, dscale, images}: ids of the
// neighbours (store ids = m_ppatches indices after a rebuild), ascending, and their number
__global__ void __launch_bounds__(CAND_WARPS * 32) k_probe_neighbors(const StoreParams sp, int n, const float4* __restrict__ coord, const float4* __restrict__ normal,
                                                                     const float4* __restrict__ scal, const int* __restrict__ images, const int* __restrict__ nimg, int stride,
                                                                     float scale, int margin, int cap, int* __restrict__ ids_out, int* __restrict__ count_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* simg = reinterpret_cast<int*>(smem_raw) + (threadIdx.x >> 5) * 2 * CAND_MAXV;
    int* cells = simg + CAND_MAXV;
    const Params& p = sp.cp.p;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    const Overlay none{-1, nullptr, 0, nullptr, 0};
    int* nb = sp.nb_scratch + (size_t)gwarp * NB_STRIDE;
    for (int h = gwarp; h < n; h += gridDim.x * CAND_WARPS) {
        const V4 X = f4v(coord[h]);
        const int nv = min(min(nimg[h], stride), CAND_MAXV);
        for (int i = lane; i < nv; i += 32) {
            const int v = images[(size_t)h * stride + i];
            simg[i] = v;
            const V3 q = project(p.views[v].P, X);
            cells[i] = pack_cell(cell_of(q.x, p.csize), cell_of(q.y, p.csize));
        }
        __syncwarp();
        PGeo me; me.X = X; me.N = f4v(normal[h]); me.dscale = scal[h].y; me.ref = simg[0];
        const PatchLists pl{simg, cells, nv, nullptr, nullptr, 0};
        const int nn = warp_find_neighbors(sp, me, pl, scale, margin, none, nb, lane);
        for (int k = lane; k < min(nn, cap); k += 32) ids_out[(size_t)h * cap + k] = nb[k];
        if (lane == 0) count_out[h] = nn;
        __syncwarp();
    }
}

// PatchManager::updateDepthMaps (patch_manager.cpp:191-221) for stored patches; one warp per patch, lanes over the views
__global__ void k_store_update_depth(const StoreParams sp, int n, const int* __restrict__ ids) {
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int k = gwarp; k < n; k += nwarps) {
        const int id = ids[k];
        if (id < 0 || id >= sp.st.cap || sp.st.state[id] != 1) continue;
        const V4 X = f4v(sp.st.coord[id]);
        for (int v = lane; v < sp.cp.p.nviews; v += 32) update_depth_map(sp, id, X, v);
    }
}

// entries of the cells of one view, flattened: pass 0 counts per cell (which = 0: m_pgrids, 1: m_vpgrids), pass 1 writes ids at offs[cell]
__global__ void k_store_cell_ids(const StoreDev st, int c0, int ncell, int which, int pass, int* __restrict__ count, const int* __restrict__ offs, int* __restrict__ ids) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    const int n = min(st.ccount[c0 + c], st.cell_cap);
    int k = 0;
    for (int s2 = 0; s2 < n; ++s2) {
        const int e = st.cslots[(size_t)(c0 + c) * st.cell_cap + s2];
        if ((e & 0x7fffffff) == SLOT_TOMB) continue;
        if ((which != 0) != (e < 0)) continue;
        if (pass) ids[offs[c] + k] = e & 0x7fffffff;
        ++k;
    }
    if (!pass) count[c] = k;
}

}  // namespace pmk
