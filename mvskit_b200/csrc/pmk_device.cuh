// mvskit_b200/csrc/pmk_device.cuh -- device-side building blocks shared by every pmk kernel.
//
// Float-order contract.  Everything that can change an INTEGER decision of the reference
// (projected cell index, pyramid level, getTexSafe accept/reject, view-angle gates, visibility) is
// computed with the reference's operation order, one IEEE rounding per operation, through the
// explicit round-to-nearest intrinsics below -- nvcc never contracts those into FMAs, whatever
// -fmad says.  Quantities the north-star only bounds by tolerance (bilinear blend, the 147-term
// normalise/dot sums) use FMAs and warp-shuffle trees.
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/pmk.h"

namespace pmk {

// ---- exact arithmetic ----------------------------------------------------------------------------
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xsqrt(float a) { return __fsqrt_rn(a); }
// std::min / std::max argument-order semantics (NaN handling differs from fminf/fmaxf)
__device__ __forceinline__ float min_std(float a, float b) { return (b < a) ? b : a; }
__device__ __forceinline__ float max_std(float a, float b) { return (a < b) ? b : a; }

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

__device__ __forceinline__ V4 ld4(const float* p) { const float4 v = *reinterpret_cast<const float4*>(p); return V4{v.x, v.y, v.z, v.w}; }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
__device__ __forceinline__ float dot4(V4 a, V4 b) { return xadd(xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)), xmul(a.w, b.w)); }
__device__ __forceinline__ float norm3(V3 a) { return xsqrt(dot3(a, a)); }
__device__ __forceinline__ float norm4(V4 a) { return xsqrt(dot4(a, a)); }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return V3{xsub(xmul(a.y, b.z), xmul(a.z, b.y)), xsub(xmul(a.z, b.x), xmul(a.x, b.z)), xsub(xmul(a.x, b.y), xmul(a.y, b.x))};
}
__device__ __forceinline__ V3 sub3(V3 a, V3 b) { return V3{xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)}; }
__device__ __forceinline__ V4 sub4(V4 a, V4 b) { return V4{xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z), xsub(a.w, b.w)}; }
__device__ __forceinline__ V4 add4(V4 a, V4 b) { return V4{xadd(a.x, b.x), xadd(a.y, b.y), xadd(a.z, b.z), xadd(a.w, b.w)}; }
__device__ __forceinline__ V4 div4(V4 a, float s) { return V4{xdiv(a.x, s), xdiv(a.y, s), xdiv(a.z, s), xdiv(a.w, s)}; }
__device__ __forceinline__ V3 div3(V3 a, float s) { return V3{xdiv(a.x, s), xdiv(a.y, s), xdiv(a.z, s)}; }
__device__ __forceinline__ V4 mul4(V4 a, float s) { return V4{xmul(a.x, s), xmul(a.y, s), xmul(a.z, s), xmul(a.w, s)}; }

// ---- image texel: R, G, B, X as four IEEE half floats (8 bytes).  The pyramid only ever holds the integers
// 0..255 (every level is re-rounded to u8, image.cpp:308-310), which binary16 represents exactly, so widening
// a texel to fp32 reproduces the reference's (float)uchar bit for bit at half the L1/L2/HBM traffic of fp32.
typedef uint2 Texel;
__device__ __forceinline__ Texel make_texel(float r, float g, float b) {
    const __half2 lo = __floats2half2_rn(r, g), hi = __floats2half2_rn(b, 0.0f);
    return make_uint2(*reinterpret_cast<const unsigned int*>(&lo), *reinterpret_cast<const unsigned int*>(&hi));
}
__device__ __forceinline__ void texel_rgb(Texel t, float& r, float& g, float& b) {
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    r = lo.x; g = lo.y;
    b = __low2float(*reinterpret_cast<const __half2*>(&t.y));
}

// ---- per-view constants (one per image, in global memory; 16-byte aligned rows) -------------------
struct __align__(16) ViewConst {
    float P[12];                          // working-level projection rows (camera.cpp:91-100)
    float center[4];                      // camera.cpp:295-308
    float oaxis[4];                       // camera.cpp:68-69
    float xaxis[4];                       // optim.cpp:43-54 (w unused)
    float yaxis[4];
    float zaxis[4];
    float Minv[12];                       // inverse of P[:, :3] at the working level, 3 rows padded to 4 (camera.cpp:331-335)
    float ipscale;                        // optim.cpp:56-64
    int gw, gh;                           // cell grid (patch_manager.cpp:36-37)
    int pad0;
    int w[PMK_MAX_LEVELS];                // image.cpp:135-138
    int h[PMK_MAX_LEVELS];
    const Texel* img[PMK_MAX_LEVELS];     // RGBX half-float texels holding the u8-rounded pyramid (image.cpp:245-315)
    const uint8_t* mask;                  // working-level silhouette mask (0 / 255), NULL = the view has none (image.cpp:143-161,717-747)
    uint64_t pad1;
};

// scalars shared by all kernels (passed by value)
struct Params {
    const ViewConst* views;
    int nviews;
    int level;            // Option::m_level
    int nlevels;          // level + 3
    int csize, wsize, tau, min_image_num, depth;
    float cos_angle1;     // cosf(m_angleThreshold1)   (optim.cpp:797)
    float cos_angle0;     // cosf(m_angleThreshold0)   (optim.cpp:183)
    float ncc_threshold, ncc_threshold_before;
    float level_scale;    // (float)(1 << level)
    // levelDiff = -level + #{k : ratio >= level_thr[k]}, thresholds found on the host with the same
    // libm expression the reference evaluates (optim.cpp:808); PMK_MAX_LEVELS-1 entries, +inf padded
    float level_thr[PMK_MAX_LEVELS];
    int has_masks;        // some view carries a mask (pmk_set_view_mask); 0 skips PhotoSet::getMask's loop altogether
};

// ---- Camera::project (camera.cpp:310-326) ------------------------------------------------------------
__device__ __forceinline__ V3 project(const float* __restrict__ P, V4 X) {
    const float i0 = dot4(ld4(P), X), i1 = dot4(ld4(P + 4), X), i2 = dot4(ld4(P + 8), X);
    if (i2 <= 0.0f) return V3{-65535.0f, -65535.0f, -1.0f};
    const float lo = -2147483648.0f, hi = 2147483648.0f;   // (float)(INT_MIN + 3.0f), (float)(INT_MAX - 3.0f)
    V3 r;
    r.x = max_std(lo, min_std(hi, xdiv(i0, i2)));
    r.y = max_std(lo, min_std(hi, xdiv(i1, i2)));
    r.z = (i2 < __int_as_float(0x7f800000)) ? 1.0f : __int_as_float(0x7fc00000);   // z / z for z > 0: 1, or NaN at +inf
    return r;
}

// project with the rows already in registers
struct Proj { V4 r0, r1, r2; };
__device__ __forceinline__ Proj load_proj(const float* __restrict__ P) { return Proj{ld4(P), ld4(P + 4), ld4(P + 8)}; }
__device__ __forceinline__ V3 project(const Proj& P, V4 X) {
    const float i0 = dot4(P.r0, X), i1 = dot4(P.r1, X), i2 = dot4(P.r2, X);
    if (i2 <= 0.0f) return V3{-65535.0f, -65535.0f, -1.0f};
    const float lo = -2147483648.0f, hi = 2147483648.0f;
    V3 r;
    r.x = max_std(lo, min_std(hi, xdiv(i0, i2)));
    r.y = max_std(lo, min_std(hi, xdiv(i1, i2)));
    r.z = (i2 < __int_as_float(0x7f800000)) ? 1.0f : __int_as_float(0x7fc00000);   // z / z for z > 0: 1, or NaN at +inf
    return r;
}

// ---- Camera::unproject (camera.cpp:329-337) at the working level: M^-1 (icoord - p4), icoord = depth * (u, v, 1) -------------------
// Minv is the host's adjugate / determinant inverse of P[:, :3] (the definition the oracle's Matrix3f::inverse shim uses).
struct ViewConst;
__device__ __forceinline__ V4 unproject_rows(const float* __restrict__ P, const float* __restrict__ Minv, V3 ic) {
    const float b0 = xsub(ic.x, P[3]), b1 = xsub(ic.y, P[7]), b2 = xsub(ic.z, P[11]);
    V4 X;
    X.x = xadd(xadd(xmul(Minv[0], b0), xmul(Minv[1], b1)), xmul(Minv[2], b2));
    X.y = xadd(xadd(xmul(Minv[4], b0), xmul(Minv[5], b1)), xmul(Minv[6], b2));
    X.z = xadd(xadd(xmul(Minv[8], b0), xmul(Minv[9], b1)), xmul(Minv[10], b2));
    X.w = 1.0f;
    return X;
}

// ---- Photo::getMask(coord, m_level) (photo.cpp:44-52) -> Image::getMask(fx, fy, level) (image.cpp:749-781) ------------
// -1: the view has no mask, or the rounded pixel lies outside the image; else the mask value (0 outside / 255 inside).
// The reference converts floorf(f + 0.5f) to int first (x86: INT_MIN for NaN and out-of-range, i.e. "outside"); comparing the
// floored floats against the image size decides the same way without the conversion.
__device__ __forceinline__ int view_mask(const ViewConst& vc, int level, V4 X) {
    if (vc.mask == nullptr) return -1;
    const V3 ic = project(vc.P, X);
    const float fx = floorf(xadd(ic.x, 0.5f)), fy = floorf(xadd(ic.y, 0.5f));
    const int W = vc.w[level], H = vc.h[level];
    if (!(fx >= 0.0f && fx < (float)W && fy >= 0.0f && fy < (float)H)) return -1;
    return (int)__ldg(vc.mask + (size_t)(int)fy * W + (int)fx);
}

// ---- PhotoSet::getMask(coord, m_level) (photoSet.cpp:223-233), warp-cooperative: 0 as soon as one view says outside, else -1 ----
__device__ __forceinline__ int warp_get_mask(const Params& p, V4 X, int lane) {
    if (!p.has_masks) return -1;
    bool outside = false;
    for (int v = lane; v < p.nviews; v += 32) outside |= view_mask(p.views[v], p.level, X) == 0;
    return __any_sync(0xffffffffu, outside) ? 0 : -1;
}

// ---- Optim::getUnit (optim.cpp:34-41): (float)(2.0 * fz * (1 << level) / ipscale), evaluated in double --------
// 2.0 * fz * 2^level is a power-of-two scaling (exact in float as well), and a correctly rounded double
// quotient of two floats narrows to the correctly rounded float quotient (double rounding is innocuous for
// division when the wide format has >= 2p + 2 = 50 bits; double has 53), so one fp32 divide is bit-identical.
__device__ __forceinline__ float unit_from_dist(float fz, float ipscale, float level_scale) {
    if (ipscale == 0.0f) return 1.0f;
    return xdiv(xmul(fz, 2.0f * level_scale), ipscale);
}
__device__ __forceinline__ float get_unit(const ViewConst& vc, V4 X, float level_scale) {
    return unit_from_dist(norm4(sub4(X, ld4(vc.center))), vc.ipscale, level_scale);
}

// ---- Optim::getPAxes (optim.cpp:67-84) ------------------------------------------------------------------
__device__ __forceinline__ void get_paxes(const ViewConst& vc, V4 X, V4 N, float level_scale, V4& px, V4& py) {
    const float pscale = get_unit(vc, X, level_scale);
    const V3 n3{N.x, N.y, N.z};
    const V4 xa = ld4(vc.xaxis);
    V3 y3 = cross3(n3, V3{xa.x, xa.y, xa.z});
    y3 = div3(y3, norm3(y3));
    const V3 x3 = cross3(y3, n3);
    px = V4{xmul(x3.x, pscale), xmul(x3.y, pscale), xmul(x3.z, pscale), xmul(0.0f, pscale)};
    py = V4{xmul(y3.x, pscale), xmul(y3.y, pscale), xmul(y3.z, pscale), xmul(0.0f, pscale)};
    const Proj P = load_proj(vc.P);
    const V3 c0 = project(P, X);
    const float xdis = norm3(sub3(project(P, add4(X, px)), c0));
    const float ydis = norm3(sub3(project(P, add4(X, py)), c0));
    px = div4(px, xdis);
    py = div4(py, ydis);
}

// ---- per (hypothesis, view) sampling frame: everything Optim::getTex decides before it samples ----------
// (optim.cpp:790-833 + getTexSafe :895-915).  level < 0 means getTex returns -1.
struct Frame {
    float tlx, tly, dxx, dxy, dyx, dyy;
    int level;
    float unit;      // Optim::computeUnits entry for this view (optim.cpp:109-132)
};

__device__ __forceinline__ Frame make_frame(const Params& p, const ViewConst& vc, V4 X, V4 N, V4 px, V4 py) {
    Frame f;
    f.level = -1;
    f.tlx = f.tly = f.dxx = f.dxy = f.dyx = f.dyy = 0.0f;
    V4 ray = sub4(ld4(vc.center), X);
    const float dist = norm4(ray);          // == ||X - center|| bit for bit (squares of negated terms)
    ray = div4(ray, dist);
    const float rn = dot4(ray, N);
    // computeUnits: getUnit / (ray . n), INT_MAX / 2 when back-facing
    f.unit = (0.0f < rn) ? xdiv(unit_from_dist(dist, vc.ipscale, p.level_scale), rn) : 1073741824.0f;
    const float weight = max_std(0.0f, rn);
    if (weight < p.cos_angle1) return f;
    const Proj P = load_proj(vc.P);
    V3 c = project(P, X);
    V3 dx = sub3(project(P, add4(X, px)), c);
    V3 dy = sub3(project(P, add4(X, py)), c);
    const float ratio = xdiv(xadd(norm3(dx), norm3(dy)), 2.0f);
    int ld = -p.level;                                   // optim.cpp:808-809 through the host-built table
#pragma unroll
    for (int k = 0; k < PMK_MAX_LEVELS; ++k) ld += (ratio >= p.level_thr[k]) ? 1 : 0;
    const int newLevel = p.level + ld;
    // myPow2(levelDiff) is a power of two: the divide is exact, so is the multiply by its inverse
    const float inv = __int_as_float((127 - ld) << 23);
    c.x = xmul(c.x, inv); c.y = xmul(c.y, inv);
    dx.x = xmul(dx.x, inv); dx.y = xmul(dx.y, inv);
    dy.x = xmul(dy.x, inv); dy.y = xmul(dy.y, inv);
    const float m = (float)(p.wsize / 2);
    // getTexSafe corners (optim.cpp:898-901)
    const float ax = xmul(dx.x, m), ay = xmul(dx.y, m), bx = xmul(dy.x, m), by = xmul(dy.y, m);
    const float tlx = xsub(xsub(c.x, ax), bx), tly = xsub(xsub(c.y, ay), by);
    const float trx = xsub(xadd(c.x, ax), bx), try_ = xsub(xadd(c.y, ay), by);
    const float blx = xadd(xsub(c.x, ax), bx), bly = xadd(xsub(c.y, ay), by);
    const float brx = xadd(xadd(c.x, ax), bx), bry = xadd(xadd(c.y, ay), by);
    const float minx = min_std(tlx, min_std(trx, min_std(blx, brx)));
    const float maxx = max_std(tlx, max_std(trx, max_std(blx, brx)));
    const float miny = min_std(tly, min_std(try_, min_std(bly, bry)));
    const float maxy = max_std(tly, max_std(try_, max_std(bly, bry)));
    const float W = (float)(vc.w[newLevel] - 1 - 2), H = (float)(vc.h[newLevel] - 1 - 2);
    if (minx < 2.0f || W <= maxx || miny < 2.0f || H <= maxy) return f;
    f.level = newLevel;
    f.tlx = tlx; f.tly = tly; f.dxx = dx.x; f.dxy = dx.y; f.dyx = dy.x; f.dyy = dy.y;
    return f;
}

// ---- Optim::computeUnits entry (optim.cpp:109-132): getUnit / (ray . n), INT_MAX/2 when back-facing ------
__device__ __forceinline__ float view_unit(const Params& p, const ViewConst& vc, V4 X, V4 N) {
    float unit = get_unit(vc, X, p.level_scale);
    V4 ray = sub4(ld4(vc.center), X);
    ray = div4(ray, norm4(ray));
    const float d = dot4(ray, N);
    return (0.0f < d) ? xdiv(unit, d) : 1073741824.0f;
}

// ---- PatchManager::setGrids cell index (patch_manager.cpp:229-231,245-247) --------------------------------
__device__ __forceinline__ int cell_of(float u, int csize) { return ((int)floorf(xadd(u, 0.5f))) / csize; }

// ---- Image::getColor bilinear (image.cpp:448-471) on RGBX half-float texels --------------------------------------
// Truncating index, reference tap weights; the blend itself is tolerance-bound and uses FMAs.
__device__ __forceinline__ void bilinear(const Texel* __restrict__ img, int W, float x, float y, float& r, float& g, float& b) {
    const int lx = (int)x, ly = (int)y;
    const float dx1 = xsub(x, (float)lx), dx0 = xsub(1.0f, dx1);
    const float dy1 = xsub(y, (float)ly), dy0 = xsub(1.0f, dy1);
    const float f00 = dx0 * dy0, f01 = dx0 * dy1, f10 = dx1 * dy0, f11 = dx1 * dy1;
    const Texel* p0 = img + (size_t)ly * W + lx;
    const Texel ta = __ldg(p0), tc = __ldg(p0 + 1), td = __ldg(p0 + W), te = __ldg(p0 + W + 1);
    float ar, ag, ab, cr, cg, cb, dr, dg, db, er, eg, eb;
    texel_rgb(ta, ar, ag, ab); texel_rgb(tc, cr, cg, cb); texel_rgb(td, dr, dg, db); texel_rgb(te, er, eg, eb);
    r = fmaf(er, f11, fmaf(cr, f10, fmaf(dr, f01, ar * f00)));
    g = fmaf(eg, f11, fmaf(cg, f10, fmaf(dg, f01, ag * f00)));
    b = fmaf(eb, f11, fmaf(cb, f10, fmaf(db, f01, ab * f00)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float robustincc(float x) { return xdiv(x, xadd(1.0f, xmul(3.0f, x))); }      // optim.cpp:622-624
__device__ __forceinline__ float unrobustincc(float x) { return xdiv(x, xsub(1.0f, xmul(3.0f, x))); }    // optim.cpp:626-628

}  // namespace pmk
