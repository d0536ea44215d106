// mvskit_b200/csrc/pmk_cand.cuh -- candidate-patch kernels: K2 (setINCCs), view selection (preProcess /
// postProcess), K3 (PMR1 refinement over Optim::cost_func) and hypothesis generation.
//
// Execution model: one warp owns one candidate patch.  A warp is split into G = 32/GW "evaluator groups" of
// GW = 8 lanes (16 for wsize > 8); a group grabs one wsize x wsize texture at a time (lane = lattice column,
// loop over rows, 3-step shuffle reductions) exactly like phase C of k1_ncc.  Depending on the step the four
// groups work on four different VIEWS of the same patch (setINCCs) or on four different HYPOTHESES of the
// refinement schedule (cost_func), so no lane idles while the list logic (sortImages etc., tiny) is serial.
#pragma once

#include "pmk_ncc.cuh"

namespace pmk {

#ifndef PMK_CAND_INLINE
#define PMK_CAND_INLINE __forceinline__
#endif

constexpr int CAND_WARPS = 4;          // warps per CTA in the candidate kernels
constexpr int PMR1_LEVELS = 12;        // refinement schedule "PMR1" (CPU twin: oracle/shim/nlopt.hpp)
constexpr int PMR1_CANDS = 8;
constexpr int PMR1_EVALS = 1 + PMR1_LEVELS * PMR1_CANDS;

// host-built decision thresholds on the ray-ray dot product so that PhotoSet::checkAngles' acos() comparison
// (photoSet.cpp:90-96) is reproduced without evaluating acos on the device
struct AngleGate {
    float dot_ge;     // (float)acos(dot) < maxAngle  <=>  dot >= dot_ge
    float dot_le;     // minAngle < (float)acos(dot)  <=>  dot <= dot_le
};

struct CandParams {
    Params p;
    AngleGate gate;               // checkAngles(minAngle = m_maxAngleThreshold, maxAngle = m_angleThreshold1)
    float sort_threshold;         // 1.0f - cos(10 deg), sortImages (optim.cpp:222)
    float ascale;                 // (float)(M_PI / 48.0f), refinePatch (optim.cpp:487)
    float* tex_scratch;           // per warp: nviews * TEXW floats (pairwise setINCCs)
    float* mat_scratch;           // per warp: nviews * nviews floats
    uint64_t seed;                // PMR1 Philox key
};

// ---- Philox4x32-10 (counter-based RNG of the PMR1 schedule) ----------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = h1 ^ c[1] ^ k0, n1 = l1, n2 = h0 ^ c[3] ^ k1, n3 = l0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ double uniform_pm1(uint32_t v) { return ((double)(v >> 8) + 0.5) * (1.0 / 8388608.0) - 1.0; }

// ---- one texture grab by one evaluator group ---------------------------------------------------------------------
// Optim::getTex + Optim::normalize for (X, N, px, py) in `view` (optim.cpp:790-844, 917-940).  Every lane of the
// group holds the same arguments.  On return: t = this lane's lattice column, centred (tex - mean) and masked;
// inv_msd = 1 / sqrt(ssd / (3 n)) (1 when ssd == 0).  Returns the pyramid level sampled or -1 (getTex == -1).
#ifndef PMK_GRAB_INLINE
#define PMK_GRAB_INLINE __noinline__     // one copy of the texture grab: the sweep kernel is I-cache bound with it inlined 15 times
#endif
template <int WS, int GW>
__device__ __forceinline__ int group_grab_body(const Params& p, int view, V4 X, V4 N, V4 px, V4 py, int col, float cmask,
                                               float t[WS][3], float& inv_msd, unsigned gm) {
    constexpr int NSAMP = WS * WS;
    constexpr float INV_NSAMP = 1.0f / (float)NSAMP, INV_3NSAMP = 1.0f / (float)(3 * NSAMP);
    Frame f;
    f.level = -1;
    f.tlx = f.tly = 2.0f; f.dxx = f.dxy = f.dyx = f.dyy = 0.0f;
    int vsafe = 0;
    if (view >= 0 && view < p.nviews) {
        f = make_frame(p, p.views[view], X, N, px, py);
        if (f.level >= 0) vsafe = view;
    }
    const int level = f.level;
    if (level < 0) { f.tlx = f.tly = 2.0f; f.dxx = f.dxy = f.dyx = f.dyy = 0.0f; }     // harmless target, see k1_ncc
    const ViewConst& vc = p.views[vsafe];
    const int lv = level < 0 ? p.level : level;
    const Texel* img = vc.img[lv];
    const int W = vc.w[lv];
    const float fcol = (float)(col < WS ? col : WS - 1);
    const float bx = fmaf(f.dxx, fcol, f.tlx), by = fmaf(f.dxy, fcol, f.tly);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int y = 0; y < WS; ++y) {
        bilinear(img, W, fmaf(f.dyx, (float)y, bx), fmaf(f.dyy, (float)y, by), t[y][0], t[y][1], t[y][2]);
        s0 = fmaf(t[y][0], cmask, s0); s1 = fmaf(t[y][1], cmask, s1); s2 = fmaf(t[y][2], cmask, s2);
    }
    const float m0 = -group_sum<GW>(s0, gm) * INV_NSAMP * cmask, m1 = -group_sum<GW>(s1, gm) * INV_NSAMP * cmask, m2 = -group_sum<GW>(s2, gm) * INV_NSAMP * cmask;
    float ssd = 0.f;
#pragma unroll
    for (int y = 0; y < WS; ++y) {
        t[y][0] = fmaf(t[y][0], cmask, m0); t[y][1] = fmaf(t[y][1], cmask, m1); t[y][2] = fmaf(t[y][2], cmask, m2);
        ssd = fmaf(t[y][0], t[y][0], fmaf(t[y][1], t[y][1], fmaf(t[y][2], t[y][2], ssd)));
    }
    const float var = group_sum<GW>(ssd, gm) * INV_3NSAMP;
    inv_msd = var > 0.0f ? rsqrtf(var) : 1.0f;
    return level;
}
template <int WS, int GW>
__device__ PMK_GRAB_INLINE int group_grab(const Params& p, int view, V4 X, V4 N, V4 px, V4 py, int col, float cmask,
                                          float t[WS][3], float& inv_msd, unsigned gm = 0xffffffffu) {
    return group_grab_body<WS, GW>(p, view, X, N, px, py, col, cmask, t, inv_msd, gm);
}

// Optim::dot of two grabbed textures (optim.cpp:601-609); both centred, scales applied here
template <int WS, int GW>
__device__ __forceinline__ float group_dot(const float a[WS][3], float inv_a, const float b[WS][3], float inv_b, unsigned gm = 0xffffffffu) {
    float dp = 0.f;
#pragma unroll
    for (int y = 0; y < WS; ++y) dp = fmaf(a[y][0], b[y][0], fmaf(a[y][1], b[y][1], fmaf(a[y][2], b[y][2], dp)));
    return group_sum<GW>(dp, gm) * inv_a * inv_b * (1.0f / (float)(3 * WS * WS));
}

// ---- Optim::decode (optim.cpp:582-599) --------------------------------------------------------------------------------
struct RefineCtx {
    V4 center, ray;      // m_center, m_ray
    float dscale;        // m_dscale
    int ref;
};

__device__ __forceinline__ void decode(const CandParams& cp, const ViewConst& vc, const RefineCtx& rc, const double x[3], V4& coord, V4& normal) {
    const float s = __double2float_rn(__dmul_rn((double)rc.dscale, x[0]));        // double scalar narrowed, then Vector4f * float
    coord = add4(rc.center, mul4(rc.ray, s));
    const float angle1 = __double2float_rn(__dmul_rn(x[1], (double)cp.ascale));
    const float angle2 = __double2float_rn(__dmul_rn(x[2], (double)cp.ascale));
    const float fx = xmul(sinf(angle1), cosf(angle2));
    const float fy = sinf(angle2);
    const float fz = xmul(-cosf(angle1), cosf(angle2));
    const V4 xa = ld4(vc.xaxis), ya = ld4(vc.yaxis), za = ld4(vc.zaxis);
    normal.x = xadd(xadd(xmul(xa.x, fx), xmul(ya.x, fy)), xmul(za.x, fz));
    normal.y = xadd(xadd(xmul(xa.y, fx), xmul(ya.y, fy)), xmul(za.y, fz));
    normal.z = xadd(xadd(xmul(xa.z, fx), xmul(ya.z, fy)), xmul(za.z, fz));
    normal.w = 0.0f;
}
__device__ __forceinline__ void decode(const CandParams& cp, const RefineCtx& rc, const double x[3], V4& coord, V4& normal) {
    decode(cp, cp.p.views[rc.ref], rc, x, coord, normal);
}

// ---- Optim::encode (optim.cpp:549-580) -----------------------------------------------------------------------------------
__device__ __forceinline__ void encode(const CandParams& cp, const RefineCtx& rc, V4 coord, V4 normal, double x[3]) {
    x[0] = (double)xdiv(dot4(sub4(coord, rc.center), rc.ray), rc.dscale);
    const ViewConst& vc = cp.p.views[rc.ref];
    const V4 xa = ld4(vc.xaxis), ya = ld4(vc.yaxis), za = ld4(vc.zaxis);
    const V3 n3{normal.x, normal.y, normal.z};
    const float fx = dot3(V3{xa.x, xa.y, xa.z}, n3), fy = dot3(V3{ya.x, ya.y, ya.z}, n3), fz = dot3(V3{za.x, za.y, za.z}, n3);
    const float a2 = asinf(fmaxf(-1.0f, fminf(1.0f, fy)));
    const float cosb = __double2float_rn(cos((double)a2));
    double a1 = 0.0;
    if (cosb != 0.0f) {
        const float sina = xdiv(fx, cosb), cosa = xdiv(-fz, cosb);
        float v = acosf(fmaxf(-1.0f, fminf(1.0f, cosa)));
        if (sina < 0.0f) v = -v;
        a1 = (double)v;
    }
    x[1] = a1 / (double)cp.ascale;
    x[2] = (double)a2 / (double)cp.ascale;
}

// ---- Optim::cost_func (optim.cpp:401-468) by one evaluator group -------------------------------------------------------------
// views/sz: m_indexes (first min(tau, n) entries are used).  Returns the cost as the reference's double.
#ifndef PMK_COST_INLINE
#define PMK_COST_INLINE __forceinline__
#endif
template <int WS, int GW>
__device__ PMK_COST_INLINE double group_cost(const CandParams& cp, const RefineCtx& rc, const double x[3], const int* views, int nimages,
                                             int col, float cmask, unsigned gm) {
    const Params& p = cp.p;
    V4 coord, normal, px, py;
    decode(cp, rc, x, coord, normal);
    get_paxes(p.views[rc.ref], coord, normal, p.level_scale, px, py);
    const int sz = min(p.tau, nimages);
    const int minimum = min(p.min_image_num, sz);
    float t0[WS][3], t[WS][3];
    float inv0, inv;
    if (group_grab<WS, GW>(p, views[0], coord, normal, px, py, col, cmask, t0, inv0, gm) < 0) return 2.0;
    double ans = 0.0;
    int denom = 0;
#pragma unroll 1
    for (int i = 1; i < sz; ++i) {
        if (group_grab<WS, GW>(p, views[i], coord, normal, px, py, col, cmask, t, inv, gm) < 0) continue;
        const float d = group_dot<WS, GW>(t0, inv0, t, inv, gm);
        ans += (double)robustincc(__double2float_rn(1.0 - (double)d));
        ++denom;
    }
    if (denom < minimum - 1) return 2.0;
    return ans / (double)denom;
}

// ---- Optim::computeINCC (optim.cpp:630-706) with given weights, by one evaluator group ----------------------------------------
template <int WS, int GW>
__device__ __forceinline__ float group_incc(const Params& p, V4 X, V4 N, const int* views, int nimages, const float* weights,
                                            int col, float cmask, unsigned gm) {
    if (nimages < 2) return 2.0f;
    V4 px, py;
    get_paxes(p.views[views[0]], X, N, p.level_scale, px, py);
    const int sz = min(p.tau, nimages);
    float t0[WS][3], t[WS][3];
    float inv0, inv;
    if (group_grab<WS, GW>(p, views[0], X, N, px, py, col, cmask, t0, inv0, gm) < 0) return 2.0f;
    float score = 0.0f, tw = 0.0f;
#pragma unroll 1
    for (int i = 1; i < sz; ++i) {
        if (group_grab<WS, GW>(p, views[i], X, N, px, py, col, cmask, t, inv, gm) < 0) continue;
        const float d = group_dot<WS, GW>(t0, inv0, t, inv, gm);
        tw = xadd(tw, weights[i]);
        score = xadd(score, xmul(robustincc(__double2float_rn(1.0 - (double)d)), weights[i]));
    }
    return (tw == 0.0f) ? 2.0f : xdiv(score, tw);
}

// Optim::computeWeights (optim.cpp:109-132, 942-948) for images[0..n), serial per lane (n is small)
__device__ __forceinline__ void compute_weights(const Params& p, V4 X, V4 N, const int* views, int n, float* w, int lane) {
    for (int i = lane; i < n; i += 32) {
        const ViewConst& vc = p.views[views[i]];
        V4 ray = sub4(ld4(vc.center), X);
        const float dist = norm4(ray);
        ray = div4(ray, dist);
        const float d = dot4(ray, N);
        w[i] = (0.0f < d) ? xdiv(unit_from_dist(dist, vc.ipscale, p.level_scale), d) : 1073741824.0f;
    }
    __syncwarp();
    const float w0 = n > 0 ? w[0] : 1.0f;
    __syncwarp();
    for (int i = lane; i < n; i += 32) w[i] = (i == 0) ? 1.0f : min_std(1.0f, xdiv(w0, w[i]));
    __syncwarp();
}

// ---- Optim::setINCCs 1-vs-all (optim.cpp:708-746): inccs[i] for images[0..n), views spread over the groups -------------------
#ifndef PMK_INCCS_INLINE
#define PMK_INCCS_INLINE __forceinline__
#endif
template <int WS, int GW>
__device__ PMK_INCCS_INLINE void warp_set_inccs(const Params& p, V4 X, V4 N, const int* views, int n, int robust, float* inccs,
                                               int lane) {
    constexpr int G = 32 / GW;
    const int grp = lane / GW, col = lane % GW;
    const float cmask = col < WS ? 1.0f : 0.0f;
    V4 px, py;
    get_paxes(p.views[views[0]], X, N, p.level_scale, px, py);
    float t0[WS][3], t[WS][3];
    float inv0, inv;
    const int l0 = group_grab<WS, GW>(p, views[0], X, N, px, py, col, cmask, t0, inv0);
    if (l0 < 0) {
        for (int i = lane; i < n; i += 32) inccs[i] = 2.0f;
        __syncwarp();
        return;
    }
    if (lane == 0) inccs[0] = 0.0f;
#pragma unroll 1
    for (int base = 1; base < n; base += G) {
        const int i = base + grp;
        const int v = i < n ? views[i] : -1;
        const int lv = group_grab<WS, GW>(p, v, X, N, px, py, col, cmask, t, inv);
        const float d = group_dot<WS, GW>(t0, inv0, t, inv);
        if (i < n && col == 0) {
            float r = 2.0f;
            if (lv >= 0) { r = xsub(1.0f, d); if (robust) r = robustincc(r); }
            inccs[i] = r;
        }
    }
    __syncwarp();
}

// ---- Optim::addImages (optim.cpp:165-205): append every other view that sees the patch within 60 degrees ------------------------
// `images` holds n entries on entry; returns the new count.  visdata2[ref] is "all other views in index order"
// (option.cpp:151-166, useVisData == 0), so the appended views come out in ascending index order.
__device__ __forceinline__ int warp_add_images(const Params& p, V4 X, V4 N, int* images, int n, int cap, unsigned char* mark, int lane) {
    for (int v = lane; v < p.nviews; v += 32) mark[v] = 0;
    __syncwarp();
    for (int i = lane; i < n; i += 32) mark[images[i]] = 1;
    __syncwarp();
    const int ref = images[0];
    for (int base = 0; base < p.nviews; base += 32) {
        const int v = base + lane;
        bool add = false;
        if (v < p.nviews && v != ref && !mark[v]) {
            const ViewConst& vc = p.views[v];
            const V3 ic = project(vc.P, X);
            const float W = (float)(vc.w[p.level] - 1), H = (float)(vc.h[p.level] - 1);
            if (!(ic.x < 0.0f || W <= ic.x || ic.y < 0.0f || H <= ic.y)) {
                V4 ray = sub4(ld4(vc.center), X);
                ray = div4(ray, norm4(ray));
                add = p.cos_angle0 <= dot4(ray, N);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, add);
        if (add) {
            const int pos = n + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) images[pos] = v;
        }
        n = min(cap, n + __popc(m));
    }
    __syncwarp();
    return n;
}

// keep images[0] and every images[i] with inccs[i] < 1 - thr  (Optim::constraintImages, optim.cpp:207-219); stable
__device__ __forceinline__ int warp_constraint(int* images, const float* inccs, int n, float thr, int lane) {
    const float lim = xsub(1.0f, thr);
    int out = n > 0 ? 1 : 0;
    for (int base = 1; base < n; base += 32) {
        const int i = base + lane;
        const bool keep = i < n && inccs[i] < lim;
        const int v = i < n ? images[i] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) images[out + __popc(m & ((1u << lane) - 1u))] = v;     // out + rank <= i: never overtakes unread entries of later chunks
        out += __popc(m);
        __syncwarp();
    }
    return out;
}

}  // namespace pmk

// =====================================================================================================================
// Warp-level list logic of Optim::preProcess / postProcess.  All arrays live in warp-private shared memory.
// =====================================================================================================================
namespace pmk {

constexpr int CAND_MAXV = 128;         // longest view list a candidate can carry (== max nviews of these kernels)

struct WarpScratch {
    int images[CAND_MAXV];
    float inccs[CAND_MAXV];
    float units[CAND_MAXV];
    float rays[CAND_MAXV][4];
    int idx[CAND_MAXV];
    unsigned char mark[CAND_MAXV];
    unsigned char alive[CAND_MAXV];
};

// Optim::sortImages(patch, isFixed = 1) (optim.cpp:221-258).  Returns the new image count (0 when fewer than two
// views face the patch, exactly as the reference clears m_images).
__device__ __forceinline__ int warp_sort_images(const CandParams& cp, V4 X, V4 N, WarpScratch& ws, int n, int lane) {
    const Params& p = cp.p;
    // computeUnits (optim.cpp:86-107): keep views with ray . n > 0, in order
    int m = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        bool keep = false;
        float unit = 0.0f;
        V4 ray{0.f, 0.f, 0.f, 0.f};
        int v = 0;
        if (i < n) {
            v = ws.images[i];
            const ViewConst& vc = p.views[v];
            ray = sub4(ld4(vc.center), X);
            const float dist = norm4(ray);
            ray = div4(ray, dist);
            const float d = dot4(ray, N);
            keep = !(d <= 0.0f);
            unit = xdiv(unit_from_dist(dist, vc.ipscale, p.level_scale), d);
        }
        const unsigned msk = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = m + __popc(msk & ((1u << lane) - 1u));
            ws.idx[pos] = v; ws.units[pos] = unit;
            ws.rays[pos][0] = ray.x; ws.rays[pos][1] = ray.y; ws.rays[pos][2] = ray.z; ws.rays[pos][3] = ray.w;
            ws.alive[pos] = 1;
        }
        m += __popc(msk);
    }
    __syncwarp();
    if (m < 2) return 0;
    if (lane == 0) ws.units[0] = 0.0f;                    // isFixed: the reference image stays first
    __syncwarp();
    const float thr = cp.sort_threshold, half = xdiv(thr, 2.0f);
    int out = 0;
    for (int step = 0; step < m; ++step) {
        // min_element over the surviving entries: smallest unit, first index on ties
        float best = __int_as_float(0x7f800000);
        int bi = 0x7fffffff;
        for (int i = lane; i < m; i += 32)
            if (ws.alive[i]) { const float u = ws.units[i]; if (bi == 0x7fffffff || u < best) { best = u; bi = i; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (oi != 0x7fffffff && (bi == 0x7fffffff || ob < best || (ob == best && oi < bi))) { best = ob; bi = oi; }
        }
        if (lane == 0) { ws.images[out] = ws.idx[bi]; ws.alive[bi] = 0; }
        ++out;
        __syncwarp();
        const V4 rs{ws.rays[bi][0], ws.rays[bi][1], ws.rays[bi][2], ws.rays[bi][3]};
        for (int i = lane; i < m; i += 32) {
            if (!ws.alive[i]) continue;
            const V4 ri{ws.rays[i][0], ws.rays[i][1], ws.rays[i][2], ws.rays[i][3]};
            const float ftmp = min_std(thr, max_std(half, xsub(1.0f, dot4(rs, ri))));
            ws.units[i] = xdiv(xmul(ws.units[i], thr), ftmp);
        }
        __syncwarp();
    }
    return out;
}

// PatchManager::setScales (patch_manager.cpp:378-399) for a fresh patch (m_dscale starts at 0)
__device__ __forceinline__ void warp_set_scales(const Params& p, V4 X, const int* images, int n, float& dscale, float& ascale, float* tmp, int lane) {
    const ViewConst& v0 = p.views[images[0]];
    const float unit = get_unit(v0, X, p.level_scale);
    const float unit2 = xmul(2.0f, unit);
    V4 ray = sub4(X, ld4(v0.center));
    ray = div4(ray, norm4(ray));
    const int num = min(p.tau, n);
    if (lane >= 1 && lane < num) {
        const ViewConst& vc = p.views[images[lane]];
        const Proj P = load_proj(vc.P);
        const V3 a = project(P, X), b = project(P, sub4(X, mul4(ray, unit2)));
        tmp[lane] = norm3(sub3(a, b));
    }
    __syncwarp();
    float ds = 0.0f;
    for (int i = 1; i < num; ++i) ds = xadd(ds, tmp[i]);
    ds = xdiv(ds, (float)(num - 1));
    ds = xdiv(unit2, ds);
    dscale = ds;
    // atan() is unqualified in the reference: double overload, narrowed on assignment
    ascale = __double2float_rn(atan((double)xdiv(ds, xdiv(xmul(unit, (float)p.wsize), 2.0f))));
    __syncwarp();
}

// PhotoSet::checkAngles (photoSet.cpp:77-103): at least one pair of rays with minAngle < angle < maxAngle
__device__ __forceinline__ bool warp_check_angles(const CandParams& cp, V4 X, WarpScratch& ws, int n, int lane) {
    const Params& p = cp.p;
    for (int i = lane; i < n; i += 32) {
        V4 ray = sub4(ld4(p.views[ws.images[i]].center), X);
        ray = div4(ray, norm4(ray));
        ws.rays[i][0] = ray.x; ws.rays[i][1] = ray.y; ws.rays[i][2] = ray.z; ws.rays[i][3] = ray.w;
    }
    __syncwarp();
    bool hit = false;
    const int npairs = n * n;
    for (int q = lane; q < npairs; q += 32) {
        const int i = q / n, j = q % n;
        if (j <= i) continue;
        const V4 a{ws.rays[i][0], ws.rays[i][1], ws.rays[i][2], ws.rays[i][3]}, b{ws.rays[j][0], ws.rays[j][1], ws.rays[j][2], ws.rays[j][3]};
        const float d = max_std(-1.0f, min_std(1.0f, dot4(a, b)));
        if (d >= cp.gate.dot_ge && d <= cp.gate.dot_le) hit = true;
    }
    return __any_sync(0xffffffffu, hit);
}

// Optim::filterImagesByAngle (optim.cpp:325-346); returns new count (0 = reference image rejected)
__device__ __forceinline__ int warp_filter_by_angle(const Params& p, V4 X, V4 N, int* images, int n, int lane) {
    int out = 0;
    bool ref_bad = false;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        bool keep = false;
        int v = 0;
        if (i < n) {
            v = images[i];
            V4 ray = sub4(ld4(p.views[v].center), X);
            ray = div4(ray, norm4(ray));
            keep = !(dot4(ray, N) < p.cos_angle1);
            if (i == 0 && !keep) ref_bad = true;
        }
        ref_bad = __any_sync(0xffffffffu, ref_bad);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) images[out + __popc(m & ((1u << lane) - 1u))] = v;
        out += __popc(m);
        __syncwarp();
    }
    return ref_bad ? 0 : out;
}

// =====================================================================================================================
// kernels
// =====================================================================================================================
__device__ __forceinline__ WarpScratch& warp_scratch(unsigned char* smem_raw) {
    return reinterpret_cast<WarpScratch*>(smem_raw)[threadIdx.x >> 5];
}

// K2: Optim::setINCCs, 1-vs-all (pairwise == 0) or all pairs (pairwise == 1)
template <int WS>
__global__ void __launch_bounds__(CAND_WARPS * 32) k2_set_inccs(const CandParams cp, int n, const float4* __restrict__ coord,
                                                                const float4* __restrict__ normal, const int* __restrict__ views,
                                                                const int* __restrict__ nviews, int stride, int robust, int pairwise,
                                                                float* __restrict__ out) {
    constexpr int GW = WS <= 8 ? 8 : 16;
    constexpr int G = 32 / GW;
    constexpr int TEXW = WS * WS * 3;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& ws = warp_scratch(smem_raw);
    const Params& p = cp.p;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    const int grp = lane / GW, col = lane % GW;
    const float cmask = col < WS ? 1.0f : 0.0f;
    for (int h = gwarp; h < n; h += gridDim.x * CAND_WARPS) {
        const float4 c = __ldg(coord + h), m = __ldg(normal + h);
        const V4 X{c.x, c.y, c.z, c.w}, N{m.x, m.y, m.z, m.w};
        const int nv = min(min(__ldg(nviews + h), stride), CAND_MAXV);
        for (int i = lane; i < nv; i += 32) ws.images[i] = __ldg(views + (size_t)h * stride + i);
        __syncwarp();
        if (nv < 1) continue;
        if (!pairwise) {
            warp_set_inccs<WS, GW>(p, X, N, ws.images, nv, robust, ws.inccs, lane);
            for (int i = lane; i < nv; i += 32) out[(size_t)h * stride + i] = ws.inccs[i];
        } else {
            // optim.cpp:748-783: grab + normalise every view once (kept in L2-resident scratch), then every pair
            float* tex = cp.tex_scratch + (size_t)gwarp * p.nviews * (TEXW + 4);
            V4 px, py;
            get_paxes(p.views[ws.images[0]], X, N, p.level_scale, px, py);
            float t[WS][3];
            float inv;
            for (int base = 0; base < nv; base += G) {
                const int i = base + grp;
                const int lv = group_grab<WS, GW>(p, i < nv ? ws.images[i] : -1, X, N, px, py, col, cmask, t, inv);
                if (i < nv) {
                    float* dst = tex + (size_t)i * (TEXW + 4);
                    if (col < WS)
#pragma unroll
                        for (int y = 0; y < WS; ++y) { float* q = dst + (y * WS + col) * 3; q[0] = t[y][0] * inv; q[1] = t[y][1] * inv; q[2] = t[y][2] * inv; }
                    if (col == 0) dst[TEXW] = lv >= 0 ? 1.0f : 0.0f;
                }
            }
            __syncwarp();
            float* o = out + (size_t)h * stride * stride;
            for (int q = lane; q < nv * nv; q += 32) {
                const int i = q / nv, j = q % nv;
                if (j < i) continue;
                float r = 0.0f;
                if (j > i) {
                    const float* a = tex + (size_t)i * (TEXW + 4);
                    const float* b = tex + (size_t)j * (TEXW + 4);
                    r = 2.0f;
                    if (a[TEXW] != 0.0f && b[TEXW] != 0.0f) {
                        float dp = 0.0f;
                        for (int e = 0; e < TEXW; ++e) dp = fmaf(a[e], b[e], dp);
                        r = xsub(1.0f, dp * (1.0f / (float)TEXW));
                        if (robust) r = robustincc(r);
                    }
                }
                o[i * stride + j] = r;
                o[j * stride + i] = r;
            }
            __syncwarp();
        }
        __syncwarp();
    }
}


// Optim::preProcess (optim.cpp:137-163) on the candidate {X, N, ws.images[0..nv)}: returns the reference's return value;
// nv / dscale / ascale are the patch's m_images.size(), m_dscale, m_ascale afterwards (nv = 0 when checkAngles clears the list)
template <int WS>
__device__ PMK_CAND_INLINE int warp_pre_process(const CandParams& cp, WarpScratch& ws, V4 X, V4 N, int& nv, float& dscale, float& ascale, int lane) {
    constexpr int GW = WS <= 8 ? 8 : 16;
    const Params& p = cp.p;
    int r = -1;
    dscale = 0.0f; ascale = 0.0f;
    if (nv >= 1) {
        nv = warp_add_images(p, X, N, ws.images, nv, CAND_MAXV, ws.mark, lane);                     // optim.cpp:139
        warp_set_inccs<WS, GW>(p, X, N, ws.images, nv, 0, ws.inccs, lane);                           // constraintImages, :141
        nv = warp_constraint(ws.images, ws.inccs, nv, p.ncc_threshold_before, lane);
        __syncwarp();
        nv = warp_sort_images(cp, X, N, ws, nv, lane);                                               // :143
        __syncwarp();
        if (nv > 0) warp_set_scales(p, X, ws.images, nv, dscale, ascale, ws.units, lane);            // :145-147
        if (nv >= p.min_image_num) {                                                                 // :149
            if (warp_check_angles(cp, X, ws, nv, lane)) r = 0;                                       // :153-160
            else nv = 0;
        }
    } else nv = 0;
    __syncwarp();
    return r;
}

// Optim::preProcess (optim.cpp:137-163) for fresh candidates {coord, normal, images}
template <int WS>
__global__ void __launch_bounds__(CAND_WARPS * 32) k_pre_process(const CandParams cp, int n, const float4* __restrict__ coord,
                                                                 const float4* __restrict__ normal, const int* __restrict__ views,
                                                                 const int* __restrict__ nviews, int stride, int maxv,
                                                                 int* __restrict__ ret, int* __restrict__ images_out, int* __restrict__ nimages_out,
                                                                 float* __restrict__ dscale_out, float* __restrict__ ascale_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& ws = warp_scratch(smem_raw);
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    for (int h = gwarp; h < n; h += gridDim.x * CAND_WARPS) {
        const float4 c = __ldg(coord + h), m = __ldg(normal + h);
        const V4 X{c.x, c.y, c.z, c.w}, N{m.x, m.y, m.z, m.w};
        int nv = min(min(__ldg(nviews + h), stride), CAND_MAXV);
        for (int i = lane; i < nv; i += 32) ws.images[i] = __ldg(views + (size_t)h * stride + i);
        __syncwarp();
        float dscale, ascale;
        const int r = warp_pre_process<WS>(cp, ws, X, N, nv, dscale, ascale, lane);
        if (lane == 0) { ret[h] = r; nimages_out[h] = nv; dscale_out[h] = dscale; ascale_out[h] = ascale; }
        for (int i = lane; i < maxv; i += 32) images_out[(size_t)h * maxv + i] = i < nv ? ws.images[i] : -1;
        __syncwarp();
    }
}

// Optim::cost_func (optim.cpp:401-468) at given encoded points: item i evaluates x[i] for patch pid[i]
template <int WS>
__global__ void __launch_bounds__(CAND_WARPS * 32) k_cost_func(const CandParams cp, int nitems, const int* __restrict__ pid,
                                                               const float4* __restrict__ coord, const float4* __restrict__ normal,
                                                               const float* __restrict__ dscale, const int* __restrict__ views,
                                                               const int* __restrict__ nviews, int stride, const double* __restrict__ x,
                                                               double* __restrict__ cost) {
    constexpr int GW = WS <= 8 ? 8 : 16;
    constexpr int G = 32 / GW;
    const Params& p = cp.p;
    const int lane = threadIdx.x & 31;
    const int grp = lane / GW, col = lane % GW;
    const float cmask = col < WS ? 1.0f : 0.0f;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    for (int base = gwarp * G; base < nitems; base += gridDim.x * CAND_WARPS * G) {
        const int it = min(base + grp, nitems - 1);
        const int h = __ldg(pid + it);
        const float4 c = __ldg(coord + h);
        const int* vrow = views + (size_t)h * stride;
        RefineCtx rc;
        rc.center = V4{c.x, c.y, c.z, c.w};
        rc.ref = __ldg(vrow);
        rc.ray = sub4(rc.center, ld4(p.views[rc.ref].center));
        rc.ray = div4(rc.ray, norm4(rc.ray));
        rc.dscale = __ldg(dscale + h);
        const double xs[3] = {x[3 * (size_t)it], x[3 * (size_t)it + 1], x[3 * (size_t)it + 2]};
        const double f = group_cost<WS, GW>(cp, rc, xs, vrow, min(__ldg(nviews + h), stride), col, cmask, group_mask<GW>(lane));
        if (col == 0 && base + grp < nitems) cost[base + grp] = f;
    }
}

// Optim::refinePatch (optim.cpp:470-547) with the PMR1 schedule in place of NLopt BOBYQA, on {X, N, ws.images[0..nv)}.
// X / N are updated in place; returns m_ncc = 1 - unrobustincc(computeINCC) with the pre-refinement weights (:539).
template <int WS>
__device__ PMK_CAND_INLINE float warp_refine(const CandParams& cp, WarpScratch& ws, V4& X, V4& N, int nv, float dscale, uint64_t stream,
                                             double* tr, int lane) {
    constexpr int GW = WS <= 8 ? 8 : 16;
    constexpr int G = 32 / GW;
    const Params& p = cp.p;
    const int grp = lane / GW, col = lane % GW;
    const float cmask = col < WS ? 1.0f : 0.0f;
    const double lb[3] = {-(double)__int_as_float(0x7f800000), -23.99999, -23.99999};
    const double ub[3] = {(double)__int_as_float(0x7f800000), 23.99999, 23.99999};
    RefineCtx rc;
    rc.center = X;
    rc.ref = ws.images[0];
    rc.ray = sub4(X, ld4(p.views[rc.ref].center));
    rc.ray = div4(rc.ray, norm4(rc.ray));
    rc.dscale = dscale;
    compute_weights(p, X, N, ws.images, nv, ws.units, lane);          // m_weights of the UNREFINED patch (optim.cpp:490)
    double best[3];
    encode(cp, rc, X, N, best);
#pragma unroll
    for (int i = 0; i < 3; ++i) best[i] = fmax(fmin(best[i], ub[i]), lb[i]);
    const unsigned gm = group_mask<GW>(lane);
    double fbest = group_cost<WS, GW>(cp, rc, best, ws.images, nv, col, cmask, gm);
    __syncwarp();
    if (tr && lane == 0) { tr[0] = best[0]; tr[1] = best[1]; tr[2] = best[2]; tr[3] = fbest; }
    double r[3] = {4.0, 4.0, 4.0};
#pragma unroll 1
    for (int level = 0; level < PMR1_LEVELS; ++level) {
        double fwin = 0.0, xwin[3] = {0.0, 0.0, 0.0};
        int cwin = -1;
#pragma unroll 1
        for (int cb = 0; cb < PMR1_CANDS; cb += G) {
            const int cnd = cb + grp;                 // this group's candidate of the level
            uint32_t ctr[4] = {(uint32_t)stream, (uint32_t)(stream >> 32), (uint32_t)level, (uint32_t)cnd};
            philox4x32_10((uint32_t)cp.seed, (uint32_t)(cp.seed >> 32), ctr);
            double xc[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) xc[i] = fmax(fmin(__dadd_rn(best[i], __dmul_rn(r[i], uniform_pm1(ctr[i]))), ub[i]), lb[i]);
            const double fc = group_cost<WS, GW>(cp, rc, xc, ws.images, nv, col, cmask, gm);
            if (tr && col == 0 && cnd < PMR1_CANDS) { double* q = tr + (size_t)(1 + level * PMR1_CANDS + cnd) * 4; q[0] = xc[0]; q[1] = xc[1]; q[2] = xc[2]; q[3] = fc; }
            // argmin over the candidates seen so far, lowest candidate index on ties (strict <, in index order)
            if (cnd < PMR1_CANDS && (cwin < 0 || fc < fwin)) { fwin = fc; cwin = cnd; xwin[0] = xc[0]; xwin[1] = xc[1]; xwin[2] = xc[2]; }
        }
        __syncwarp();
        // combine the groups: smallest cost, then smallest candidate index
#pragma unroll
        for (int o = GW; o < 32; o <<= 1) {
            const double of = __shfl_xor_sync(0xffffffffu, fwin, o);
            const int oc = __shfl_xor_sync(0xffffffffu, cwin, o);
            const double o0 = __shfl_xor_sync(0xffffffffu, xwin[0], o), o1 = __shfl_xor_sync(0xffffffffu, xwin[1], o), o2 = __shfl_xor_sync(0xffffffffu, xwin[2], o);
            if (oc >= 0 && (cwin < 0 || of < fwin || (of == fwin && oc < cwin))) { fwin = of; cwin = oc; xwin[0] = o0; xwin[1] = o1; xwin[2] = o2; }
        }
        if (fwin < fbest) { fbest = fwin; best[0] = xwin[0]; best[1] = xwin[1]; best[2] = xwin[2]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = __dmul_rn(r[i], 0.6);
    }
    // optim.cpp:534-541: decode, normal.w = 0, ncc = 1.0 - unrobustincc(computeINCC(...)) with the stale weights
    V4 Xf, Nf;
    decode(cp, rc, best, Xf, Nf);
    const float incc = group_incc<WS, GW>(p, Xf, Nf, ws.images, nv, ws.units, col, cmask, gm);
    __syncwarp();
    X = Xf;
    N = V4{Nf.x, Nf.y, Nf.z, 0.0f};
    return __double2float_rn(1.0 - (double)unrobustincc(incc));
}

template <int WS>
__global__ void __launch_bounds__(CAND_WARPS * 32) k3_refine(const CandParams cp, int n, float4* __restrict__ coord, float4* __restrict__ normal,
                                                             const float* __restrict__ dscale, const int* __restrict__ views,
                                                             const int* __restrict__ nviews, int stride, const uint64_t* __restrict__ streams,
                                                             float* __restrict__ ncc_out, double* __restrict__ trace) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& ws = warp_scratch(smem_raw);
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    for (int h = gwarp; h < n; h += gridDim.x * CAND_WARPS) {
        const float4 c = coord[h], m = normal[h];
        V4 X{c.x, c.y, c.z, c.w}, N{m.x, m.y, m.z, m.w};
        const int nv = min(min(__ldg(nviews + h), stride), CAND_MAXV);
        for (int i = lane; i < nv; i += 32) ws.images[i] = __ldg(views + (size_t)h * stride + i);
        __syncwarp();
        const float ncc = warp_refine<WS>(cp, ws, X, N, nv, __ldg(dscale + h), __ldg(streams + h), trace ? trace + (size_t)h * PMR1_EVALS * 4 : nullptr, lane);
        if (lane == 0) {
            coord[h] = make_float4(X.x, X.y, X.z, X.w);
            normal[h] = make_float4(N.x, N.y, N.z, 0.0f);
            ncc_out[h] = ncc;
        }
        __syncwarp();
    }
}

// Optim::setRefImage (optim.cpp:348-383) on ws.images[0..nv): pairwise robust INCC, reference = argmin of the row sums
// (accumulate(row, 0.0f) in index order; first minimum wins, start value INT_MAX / 2), swapped into slot 0.
template <int WS>
__device__ PMK_CAND_INLINE void warp_set_ref_image(const CandParams& cp, WarpScratch& ws, V4 X, V4 N, int nv, int wslot, int lane) {
    constexpr int GW = WS <= 8 ? 8 : 16;
    constexpr int G = 32 / GW;
    constexpr int TEXW = WS * WS * 3;
    const Params& p = cp.p;
    const int grp = lane / GW, col = lane % GW;
    const float cmask = col < WS ? 1.0f : 0.0f;
    float* tex = cp.tex_scratch + (size_t)wslot * p.nviews * (TEXW + 4);
    float* mat = cp.mat_scratch + (size_t)wslot * p.nviews * p.nviews;
    V4 px, py;
    get_paxes(p.views[ws.images[0]], X, N, p.level_scale, px, py);
    float t[WS][3];
    float inv;
    for (int base = 0; base < nv; base += G) {
        const int i = base + grp;
        const int lv = group_grab<WS, GW>(p, i < nv ? ws.images[i] : -1, X, N, px, py, col, cmask, t, inv);
        if (i < nv) {
            float* dst = tex + (size_t)i * (TEXW + 4);
            if (col < WS)
#pragma unroll
                for (int y = 0; y < WS; ++y) { float* q = dst + (y * WS + col) * 3; q[0] = t[y][0] * inv; q[1] = t[y][1] * inv; q[2] = t[y][2] * inv; }
            if (col == 0) dst[TEXW] = lv >= 0 ? 1.0f : 0.0f;
        }
    }
    __syncwarp();
    for (int q = lane; q < nv * nv; q += 32) {
        const int i = q / nv, j = q % nv;
        if (j < i) continue;
        float v = 0.0f;
        if (j > i) {
            const float* a = tex + (size_t)i * (TEXW + 4);
            const float* b = tex + (size_t)j * (TEXW + 4);
            v = 2.0f;
            if (a[TEXW] != 0.0f && b[TEXW] != 0.0f) {
                float dp = 0.0f;
                for (int e = 0; e < TEXW; ++e) dp = fmaf(a[e], b[e], dp);
                v = robustincc(xsub(1.0f, dp * (1.0f / (float)TEXW)));
            }
        }
        mat[i * nv + j] = v; mat[j * nv + i] = v;
    }
    __syncwarp();
    float best = 1073741824.0f;
    int bi = -1;
    for (int i = lane; i < nv; i += 32) {
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
    __syncwarp();
    if (lane == 0 && bi > 0) { const int tmp = ws.images[0]; ws.images[0] = ws.images[bi]; ws.images[bi] = tmp; }
    __syncwarp();
}

// Optim::postProcess (optim.cpp:260-290), the part that does not read the patch store (everything before
// setVImagesVGrids / check), on {X, N, ws.images[0..nv)}.  Returns 0 / -1; nv = m_images.size() on success.
// `wslot` selects this warp's slice of the pairwise scratch.
template <int WS>
__device__ PMK_CAND_INLINE int warp_post_process(const CandParams& cp, WarpScratch& ws, V4 X, V4 N, int& nv, int wslot, int lane) {
    constexpr int GW = WS <= 8 ? 8 : 16;
    const Params& p = cp.p;
    int r = -1;
    do {
        if (nv < p.min_image_num) break;                                                             // :261
        if (warp_get_mask(p, X, lane) == 0) break;                                                   // :265
        nv = warp_add_images(p, X, N, ws.images, nv, CAND_MAXV, ws.mark, lane);                      // :268
        warp_set_inccs<WS, GW>(p, X, N, ws.images, nv, 0, ws.inccs, lane);                           // :269
        nv = warp_constraint(ws.images, ws.inccs, nv, p.ncc_threshold, lane);
        __syncwarp();
        nv = warp_filter_by_angle(p, X, N, ws.images, nv, lane);                                     // :270
        __syncwarp();
        if (nv < p.min_image_num) break;                                                             // :272
        warp_set_ref_image<WS>(cp, ws, X, N, nv, wslot, lane);                                       // :277
        warp_set_inccs<WS, GW>(p, X, N, ws.images, nv, 0, ws.inccs, lane);                           // :279
        nv = warp_constraint(ws.images, ws.inccs, nv, p.ncc_threshold, lane);
        __syncwarp();
        if (nv < p.min_image_num) break;                                                             // :281
        r = 0;
    } while (false);
    __syncwarp();
    return r;
}

// tmp_out = Patch::score2(nccThreshold).
template <int WS>
__global__ void __launch_bounds__(CAND_WARPS * 32) k_post_process(const CandParams cp, int n, const float4* __restrict__ coord,
                                                                  const float4* __restrict__ normal, const float* __restrict__ ncc,
                                                                  const int* __restrict__ views, const int* __restrict__ nviews, int stride, int maxv,
                                                                  int* __restrict__ ret, int* __restrict__ images_out, int* __restrict__ nimages_out,
                                                                  int* __restrict__ grids_out, float* __restrict__ tmp_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch& ws = warp_scratch(smem_raw);
    const Params& p = cp.p;
    const int lane = threadIdx.x & 31;
    const int gwarp = blockIdx.x * CAND_WARPS + (threadIdx.x >> 5);
    for (int h = gwarp; h < n; h += gridDim.x * CAND_WARPS) {
        const float4 c = __ldg(coord + h), m = __ldg(normal + h);
        const V4 X{c.x, c.y, c.z, c.w}, N{m.x, m.y, m.z, m.w};
        int nv = min(min(__ldg(nviews + h), stride), CAND_MAXV);
        for (int i = lane; i < nv; i += 32) ws.images[i] = __ldg(views + (size_t)h * stride + i);
        __syncwarp();
        const int r = warp_post_process<WS>(cp, ws, X, N, nv, gwarp, lane);
        const int nout = r == 0 ? nv : 0;
        if (lane == 0) {
            ret[h] = r; nimages_out[h] = nout;
            // m_tmp = score2(nccThreshold) = max(0, ncc - thr) * nimages   (patch.cpp:27-29)
            tmp_out[h] = r == 0 ? xmul(max_std(0.0f, xsub(__ldg(ncc + h), p.ncc_threshold)), (float)nv) : 0.0f;
        }
        for (int i = lane; i < maxv; i += 32) {
            int v = -1, ix = 0, iy = 0;
            if (i < nout) {
                v = ws.images[i];
                const V3 ic = project(p.views[v].P, X);                                                   // setGrids, :285
                ix = cell_of(ic.x, p.csize); iy = cell_of(ic.y, p.csize);
            }
            images_out[(size_t)h * maxv + i] = v;
            grids_out[((size_t)h * maxv + i) * 2] = ix; grids_out[((size_t)h * maxv + i) * 2 + 1] = iy;
        }
        __syncwarp();
    }
}

}  // namespace pmk
