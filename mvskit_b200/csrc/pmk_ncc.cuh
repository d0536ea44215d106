// mvskit_b200/csrc/pmk_ncc.cuh -- K0 (pyramid) and K1 (batched hypothesis NCC) kernels.
#pragma once

#include "pmk_device.cuh"

namespace pmk {

// =====================================================================================================
// K0: image pyramid.  Image::buildImagePyramid (image/image.cpp:245-315), filter == 0.
// The 4x4 [1 3 3 1]x[1 3 3 1]/64 taps times u8 values are exact in fp32 and so are their partial sums
// (multiples of 1/64 below 256), so any summation order reproduces the reference bit for bit; the
// result is re-rounded to an integer (floorf(c + 0.5f)) and stored as a half-float RGBX texel (exact for 0..255).
// =====================================================================================================
__global__ void k0_u8_to_rgbx(const uint8_t* __restrict__ src, Texel* __restrict__ dst, int npix) {
    // four pixels = three aligned 32-bit words in, four texels out: coalesced both ways (the staging buffer is 256-byte aligned)
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = 4 * q;
    if (i >= npix) return;
    if (i + 3 < npix) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(src) + 3 * (size_t)q;
        const uint32_t a = __ldg(w), b = __ldg(w + 1), c = __ldg(w + 2);
        dst[i] = make_texel((float)(a & 255u), (float)((a >> 8) & 255u), (float)((a >> 16) & 255u));
        dst[i + 1] = make_texel((float)(a >> 24), (float)(b & 255u), (float)((b >> 8) & 255u));
        dst[i + 2] = make_texel((float)((b >> 16) & 255u), (float)(b >> 24), (float)(c & 255u));
        dst[i + 3] = make_texel((float)((c >> 8) & 255u), (float)((c >> 16) & 255u), (float)(c >> 24));
    } else {
        for (int k = i; k < npix; ++k) { const uint8_t* p = src + (size_t)k * 3; dst[k] = make_texel((float)p[0], (float)p[1], (float)p[2]); }
    }
}

__global__ void k0_downsample(const Texel* __restrict__ src, int Wp, int Hp, Texel* __restrict__ dst, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    const float k4[4] = {1.0f / 8.0f, 3.0f / 8.0f, 3.0f / 8.0f, 1.0f / 8.0f};
    float r = 0.0f, g = 0.0f, b = 0.0f;
#pragma unroll
    for (int i = -1; i < 3; ++i) {
        const int yt = 2 * y + i;
        if (yt < 0 || Hp - 1 < yt) continue;
#pragma unroll
        for (int j = -1; j < 3; ++j) {
            const int xt = 2 * x + j;
            if (xt < 0 || Wp - 1 < xt) continue;
            float tr, tg, tb;
            texel_rgb(__ldg(src + (size_t)yt * Wp + xt), tr, tg, tb);
            const float wgt = k4[i + 1] * k4[j + 1];
            r += wgt * tr; g += wgt * tg; b += wgt * tb;
        }
    }
    dst[(size_t)y * W + x] = make_texel(floorf(r + 0.5f), floorf(g + 0.5f), floorf(b + 0.5f));
}

__global__ void k0_rgbx_to_u8(const Texel* __restrict__ src, uint8_t* __restrict__ dst, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    float r, g, b;
    texel_rgb(src[i], r, g, b);
    dst[(size_t)i * 3] = (uint8_t)r; dst[(size_t)i * 3 + 1] = (uint8_t)g; dst[(size_t)i * 3 + 2] = (uint8_t)b;
}

// K0m: silhouette masks.  Image::alloc thresholds the level-0 mask (grey > 127 -> 255, else 0; image.cpp:149-156) and
// buildMaskPyramid (:717-747) marks a coarse pixel inside when any of its 2x2 fine pixels is.  u8 in, u8 out, exact.
__global__ void k0_mask_threshold(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npix) dst[i] = src[i] > 127 ? 255 : 0;
}

__global__ void k0_mask_downsample(const uint8_t* __restrict__ src, int Wp, int Hp, uint8_t* __restrict__ dst, int W, int H) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || y >= H) return;
    // W = Wp / 2, so 2x + 1 <= Wp - 1: the reference's min(m_widths[level - 1], 2x + 1) (:725) never binds
    const uint8_t* r0 = src + (size_t)(2 * y) * Wp + 2 * x;
    const uint8_t* r1 = r0 + Wp;
    dst[(size_t)y * W + x] = (r0[0] | r0[1] | r1[0] | r1[1]) ? 255 : 0;
}

// PhotoSet::getMask(coord, m_level) (view < 0) or PhotoSet::getMask(view, coord, m_level), one thread per point
__global__ void k_probe_mask(const Params p, int n, int view, const float4* __restrict__ coord, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = coord[i];
    const V4 X{c.x, c.y, c.z, c.w};
    int r = -1;
    if (view >= 0) r = view_mask(p.views[view], p.level, X);
    else for (int v = 0; v < p.nviews; ++v) if (view_mask(p.views[v], p.level, X) == 0) { r = 0; break; }
    out[i] = r;
}

// =====================================================================================================
// K1: batched hypothesis NCC = PatchManager::computeNcc (patch_manager.cpp:401-404).
//
// Work decomposition (one warp owns a batch of 32 hypotheses, no block-level synchronisation):
//   phase A  lane = hypothesis      getPAxes in the reference view                      (optim.cpp:67-84)
//   phase B  lane = hypothesis,     loop over its <= tau views: sampling frame (angle gate, 3 projections,
//            pyramid level, getTexSafe) and computeUnits weight; frames go to warp-private shared memory
//                                                                                      (optim.cpp:790-833,895-915,109-132)
//   phase C  32/GW hypotheses at a time, GW = 8 lanes each (16 for wsize > 8): lane = lattice column, loop over
//            rows: bilinear gather, normalise, dot against the reference view; sums are 3-step shuffle
//            reductions inside each lane group, shared by all groups         (optim.cpp:835-842,917-940,601-609)
//   phase D  lane = hypothesis      robust weighted mean                               (optim.cpp:686-703)
// Phases A, B and D keep all 32 lanes busy with distinct hypotheses; only phase C is per-hypothesis.
// =====================================================================================================
constexpr int K1_WARPS = 8;            // warps per CTA
constexpr int K1_FRAME_WORDS = 8;      // tlx tly dxx dxy dyx dyy (level | view << 4) weight

// sum over the aligned group of GW lanes this lane belongs to (every lane of the group gets the total)
// `mask` names the lanes that execute the call together: the whole warp, or just this lane's group when the
// groups may sit in different control flow (candidate kernels)
template <int GW>
__device__ __forceinline__ float group_sum(float v, unsigned mask = 0xffffffffu) {
#pragma unroll
    for (int o = GW / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}
template <int GW>
__device__ __forceinline__ unsigned group_mask(int lane) { return GW >= 32 ? 0xffffffffu : (((1u << GW) - 1u) << ((lane / GW) * GW)); }

// L2 prefetch of the texels one sampling frame will gather (pyramids that do not fit L2: every gather of phase C would
// otherwise pay a DRAM round trip inside a dependent chain).  The GW lanes of a group share the frame's bounding box:
// rows [y0, y1] x texels [x0, x1], one prefetch per 32-byte sector, sectors dealt out over the lanes.  No effect on results.
template <int WS, int GW>
__device__ __forceinline__ void prefetch_frame(const Params& p, const float* __restrict__ fr, int col) {
    const float4 fa = *reinterpret_cast<const float4*>(fr);            // tlx tly dxx dxy
    const float4 fb = *reinterpret_cast<const float4*>(fr + 4);        // dyx dyy (level | view << 4) weight
    const int pk = __float_as_int(fb.z);
    const ViewConst& vc = p.views[pk >> 4];
    const int lv = pk & 15, W = vc.w[lv], H = vc.h[lv];
    constexpr float E = (float)(WS - 1);
    const float ax = E * fa.z, bx = E * fb.x, ay = E * fa.w, by = E * fb.y;
    const int x0 = max(0, (int)(fa.x + fminf(ax, 0.f) + fminf(bx, 0.f))), x1 = min(W - 1, (int)(fa.x + fmaxf(ax, 0.f) + fmaxf(bx, 0.f)) + 1);
    const int y0 = max(0, (int)(fa.y + fminf(ay, 0.f) + fminf(by, 0.f))), y1 = min(H - 1, (int)(fa.y + fmaxf(ay, 0.f) + fmaxf(by, 0.f)) + 1);
    if (x1 < x0 || y1 < y0) return;
    const char* base = reinterpret_cast<const char*>(vc.img[lv]);
    const int nr = min(y1 - y0 + 1, 16);
    for (int t = col; t < nr * 4; t += GW) {
        const size_t rowb = (size_t)(y0 + (t >> 2)) * W;
        const size_t b0 = (rowb + x0) * sizeof(Texel), b1 = (rowb + x1 + 1) * sizeof(Texel);
        const size_t a = (b0 & ~(size_t)31) + 32 * (t & 3);
        if (a < b1) asm volatile("prefetch.global.L2 [%0];" :: "l"(base + a));
    }
}

template <int WS, int MINB>
__global__ void __launch_bounds__(K1_WARPS * 32, MINB) k1_ncc(const Params p, int n, const float4* __restrict__ coord,
                                                              const float4* __restrict__ normal, const int* __restrict__ views,
                                                              const int* __restrict__ nviews, int stride,
                                                              float* __restrict__ incc_out, float* __restrict__ ncc_out,
                                                              int* __restrict__ levels_out, unsigned int* __restrict__ next_batch,
                                                              const unsigned int* __restrict__ ready, unsigned int epoch, int chunk_shift,
                                                              int flags) {
    const int packed = flags & 1;                         // pmk_ncc_eval_packed's wire format
    const bool prefetch = (flags & 2) != 0;               // pyramid larger than L2: prefetch the next view's footprint while sampling this one
    // packed != 0 (pmk_ncc_eval_packed): coord / normal are rows of 3 floats (w = 1 / w = 0 implied), views rows of `stride` bytes,
    // nviews bytes -- 31 B per hypothesis over PCIe instead of 60; everything downstream is unchanged
    constexpr int NSAMP = WS * WS;
    constexpr int GW = WS <= 8 ? 8 : 16;                  // lanes that share one hypothesis in phase C
    constexpr int G = 32 / GW;                            // hypotheses sampled concurrently by a warp
    constexpr float INV_NSAMP = 1.0f / (float)NSAMP;
    constexpr float INV_3NSAMP = 1.0f / (float)(3 * NSAMP);
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    // warp-private frame store: [hyp 0..31][view 0..tau-1][8 words], hypothesis stride padded by 4 words
    // so the float4 stores of phase B (lane = hypothesis) are bank-conflict free per quarter warp
    const int fstride = p.tau * K1_FRAME_WORDS + 4;
    float* frames = smem + (size_t)warp_in_cta * 32 * fstride;
    const int nbatch = (n + 31) / 32;
    const int grp = lane / GW, col = lane % GW;           // phase C: hypothesis slot and lattice column of this lane
    const float cmask = col < WS ? 1.0f : 0.0f;           // lanes past the lattice width contribute nothing
    const float fcol = (float)(col < WS ? col : WS - 1);

    // batches are handed out through a global counter: hypotheses differ a lot in cost (0..tau valid views)
    for (;;) {
        int batch = 0;
        if (lane == 0) batch = (int)atomicAdd(next_batch, 1u);
        batch = __shfl_sync(0xffffffffu, batch, 0);
        if (batch >= nbatch) break;
        // streamed input (pmk_ncc_eval with host buffers): the hypotheses arrive over PCIe in chunks of 1 << chunk_shift while this
        // kernel runs; a copy-engine write of `epoch` into ready[chunk] follows each chunk's copies on the same stream.  The queue
        // hands batches out in input order, so a warp only waits when the kernel has caught up with the link.
        if (ready != nullptr) {
            if (lane == 0) {
                const unsigned int* f = ready + ((batch * 32) >> chunk_shift);
                unsigned int seen;
                for (;;) {
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
                    if (seen == epoch) break;
                    __nanosleep(256);
                }
            }
            __syncwarp();
        }
        const int h = batch * 32 + lane;
        const bool live = h < n;
        // ---------------- phase A/B: lane = hypothesis ----------------
        V4 X{0.f, 0.f, 0.f, 1.f}, N{0.f, 0.f, 1.f, 0.f};
        int nv = 0;
        if (live) {
            if (packed) {
                const float* c3 = reinterpret_cast<const float*>(coord) + 3 * (size_t)h;
                const float* n3 = reinterpret_cast<const float*>(normal) + 3 * (size_t)h;
                X = V4{__ldcg(c3), __ldcg(c3 + 1), __ldcg(c3 + 2), 1.0f}; N = V4{__ldcg(n3), __ldcg(n3 + 1), __ldcg(n3 + 2), 0.0f};
                nv = (int)__ldcg(reinterpret_cast<const unsigned char*>(nviews) + h);
            } else {
                const float4 c = __ldcg(coord + h), m = __ldcg(normal + h);   // read once; L2 only (the streamed inputs land during the kernel)
                X = V4{c.x, c.y, c.z, c.w}; N = V4{m.x, m.y, m.z, m.w};
                nv = __ldcg(nviews + h);
            }
        }
        const int sz = min(p.tau, min(nv, stride));       // a count beyond the row length must not read the next hypothesis' views
        const bool usable = live && nv >= 2;              // optim.cpp:631,643: fewer than 2 images -> 2.0
        int valid_mask = 0;
        float* myrow = frames + (size_t)lane * fstride;
        {
            const int* vrow = views + (size_t)h * stride;
            const unsigned char* vrow8 = reinterpret_cast<const unsigned char*>(views) + (size_t)h * stride;
            V4 px{0.f, 0.f, 0.f, 0.f}, py{0.f, 0.f, 0.f, 0.f};
            bool ref_ok = false;
            if (usable) {
                const int ref = packed ? (int)__ldcg(vrow8) : ready ? __ldcg(vrow) : __ldg(vrow);              // streamed rows must not take the non-coherent path
                ref_ok = ref >= 0 && ref < p.nviews;
                if (ref_ok) get_paxes(p.views[ref], X, N, p.level_scale, px, py);
            }
            float unit0 = 1.0f;
#pragma unroll 1
            for (int k = 0; k < p.tau; ++k) {
                Frame f; f.level = -1; f.tlx = f.tly = f.dxx = f.dxy = f.dyx = f.dyy = 0.0f; f.unit = 1.0f;
                float w = 0.0f;
                int vk = 0;
                if (ref_ok && k < sz) {
                    const int v = packed ? (int)__ldcg(vrow8 + k) : ready ? __ldcg(vrow + k) : __ldg(vrow + k);
                    if (v >= 0 && v < p.nviews) {
                        vk = v;
                        const ViewConst& vc = p.views[v];
                        f = make_frame(p, vc, X, N, px, py);
                        if (k == 0) { unit0 = f.unit; w = 1.0f; }
                        else w = min_std(1.0f, xdiv(unit0, f.unit));            // optim.cpp:944-946
                    }
                }
                if (live && levels_out) levels_out[(size_t)h * p.tau + k] = f.level;
                if (f.level >= 0) valid_mask |= 1 << k;
                else {
                    // frames nobody may sample still get a harmless target (texel (2,2) of view 0 at the working
                    // level, zero pitch) so that phase C can run branch-free over every lane group
                    f.tlx = f.tly = 2.0f; f.dxx = f.dxy = f.dyx = f.dyy = 0.0f; f.level = p.level; vk = 0;
                }
                float* fr = myrow + k * K1_FRAME_WORDS;
                *reinterpret_cast<float4*>(fr) = make_float4(f.tlx, f.tly, f.dxx, f.dxy);
                *reinterpret_cast<float4*>(fr + 4) = make_float4(f.dyx, f.dyy, __int_as_float((f.level & 15) | (vk << 4)), w);
            }
        }
        __syncwarp();

        // ---------------- phase C: G hypotheses per pass, GW lanes each; lane = lattice column, loop = row ----------------
        // Branch-free over the warp: lanes past the lattice width re-sample the last column and invalid frames
        // point at a harmless texel (phase B); both are kept out of the sums arithmetically (cmask) or simply never
        // stored, so there is no divergence inside the gather.
        // A hypothesis needs the gather only if its reference view and at least one other are valid.
        const bool need = usable && (valid_mask & 1) && (valid_mask & ~1);
        const unsigned need_mask = __ballot_sync(0xffffffffu, need);
#pragma unroll 1
        for (int i = 0; i < 32 / G; ++i) {
            if (!((need_mask >> (i * G)) & ((1u << G) - 1u))) continue;      // warp-uniform: nobody in this pass
            const int j = i * G + grp;                                         // this lane's hypothesis
            const int vmj = __shfl_sync(0xffffffffu, valid_mask, j);
            const int vm = ((need_mask >> j) & 1u) ? vmj : 0;
            const unsigned any_vm = __reduce_or_sync(0xffffffffu, (unsigned)vm);
            float* row = frames + (size_t)j * fstride;
            float t0[WS][3];                                   // centred reference texture: this lane's column
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;                // reference-view channel means (times cmask)
            float c0 = 0.f, c1 = 0.f, c2 = 0.f;                // residual sums of the centred reference texture
            float inv_msd0 = 1.0f;
#pragma unroll 1
            for (int k = 0; k < p.tau; ++k) {
                if (prefetch) {                                 // warp-uniform: the frame sampled after this one (next view, or the next pass' first)
                    if (k + 1 < p.tau) prefetch_frame<WS, GW>(p, row + (k + 1) * K1_FRAME_WORDS, col);
                    else if (i + 1 < 32 / G) prefetch_frame<WS, GW>(p, row + (size_t)G * fstride, col);
                }
                if (!((any_vm >> k) & 1u)) continue;           // warp-uniform
                const float* fr = row + k * K1_FRAME_WORDS;
                const float4 fa = *reinterpret_cast<const float4*>(fr);
                const float4 fb = *reinterpret_cast<const float4*>(fr + 4);
                const int packed = __float_as_int(fb.z);
                const ViewConst& vc = p.views[packed >> 4];
                const Texel* img = vc.img[packed & 15];
                const int W = vc.w[packed & 15];
                // Vector3f samp = tl + dx * x + dy * y   (optim.cpp:837); tolerance domain -> FMAs
                const float bx = fmaf(fa.z, fcol, fa.x), by = fmaf(fa.w, fcol, fa.y);
                if (k == 0) {
                    // Optim::normalize (optim.cpp:917-940) of the reference texture: per-channel mean, joint variance
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int y = 0; y < WS; ++y) {
                        bilinear(img, W, fmaf(fb.x, (float)y, bx), fmaf(fb.y, (float)y, by), t0[y][0], t0[y][1], t0[y][2]);
                        s0 = fmaf(t0[y][0], cmask, s0); s1 = fmaf(t0[y][1], cmask, s1); s2 = fmaf(t0[y][2], cmask, s2);
                    }
                    a0 = group_sum<GW>(s0) * INV_NSAMP; a1 = group_sum<GW>(s1) * INV_NSAMP; a2 = group_sum<GW>(s2) * INV_NSAMP;
                    const float m0 = -a0 * cmask, m1 = -a1 * cmask, m2 = -a2 * cmask;
                    float ssd = 0.f;
                    c0 = c1 = c2 = 0.f;
#pragma unroll
                    for (int y = 0; y < WS; ++y) {
                        t0[y][0] = fmaf(t0[y][0], cmask, m0); t0[y][1] = fmaf(t0[y][1], cmask, m1); t0[y][2] = fmaf(t0[y][2], cmask, m2);
                        c0 += t0[y][0]; c1 += t0[y][1]; c2 += t0[y][2];
                        ssd = fmaf(t0[y][0], t0[y][0], fmaf(t0[y][1], t0[y][1], fmaf(t0[y][2], t0[y][2], ssd)));
                    }
                    ssd = group_sum<GW>(ssd);
                    c0 = group_sum<GW>(c0); c1 = group_sum<GW>(c1); c2 = group_sum<GW>(c2);
                    const float var = ssd * INV_3NSAMP;
                    inv_msd0 = var > 0.0f ? rsqrtf(var) : 1.0f;            // msd == 0 -> 1 (optim.cpp:934-936)
                } else {
                    // one pass over view k, centred on the REFERENCE view's channel means (photo-consistent views
                    // differ little from them, so the moment subtraction below does not cancel):
                    //   u = tex_k - a_ref;  mean_k = a_ref + sum(u)/n;  ssd_k = sum(u^2) - n * |sum(u)/n|^2
                    //   sum(t0 * (tex_k - mean_k)) = sum(t0 * u) - (sum(u)/n) . sum(t0)
                    const float m0 = -a0 * cmask, m1 = -a1 * cmask, m2 = -a2 * cmask;
                    float u0 = 0.f, u1 = 0.f, u2 = 0.f, uu = 0.f, tu = 0.f;
#pragma unroll
                    for (int y = 0; y < WS; ++y) {
                        float r, g, b;
                        bilinear(img, W, fmaf(fb.x, (float)y, bx), fmaf(fb.y, (float)y, by), r, g, b);
                        r = fmaf(r, cmask, m0); g = fmaf(g, cmask, m1); b = fmaf(b, cmask, m2);
                        u0 += r; u1 += g; u2 += b;
                        uu = fmaf(r, r, fmaf(g, g, fmaf(b, b, uu)));
                        tu = fmaf(t0[y][0], r, fmaf(t0[y][1], g, fmaf(t0[y][2], b, tu)));
                    }
                    u0 = group_sum<GW>(u0) * INV_NSAMP; u1 = group_sum<GW>(u1) * INV_NSAMP; u2 = group_sum<GW>(u2) * INV_NSAMP;
                    uu = group_sum<GW>(uu);
                    tu = group_sum<GW>(tu);
                    if (col == 0 && ((vm >> k) & 1)) {
                        const float ssd = fmaxf(uu - (float)NSAMP * fmaf(u0, u0, fmaf(u1, u1, u2 * u2)), 0.0f);
                        const float var = ssd * INV_3NSAMP;
                        const float inv_msd = var > 0.0f ? rsqrtf(var) : 1.0f;
                        const float dp = tu - fmaf(u0, c0, fmaf(u1, c1, u2 * c2));
                        // Optim::dot (optim.cpp:601-609): sum(tex0 . tex_k) / (3 * sz); overwrites tlx (consumed)
                        row[k * K1_FRAME_WORDS] = dp * inv_msd0 * inv_msd * INV_3NSAMP;
                    }
                }
            }
        }
        __syncwarp();

        // ---------------- phase D: lane = hypothesis ----------------
        if (live) {
            float score = 2.0f;                                   // optim.cpp:631,654-655
            if (usable && (valid_mask & 1)) {
                float acc = 0.0f, tw = 0.0f;
#pragma unroll 1
                for (int k = 1; k < p.tau; ++k) {
                    if ((valid_mask >> k) & 1) {
                        const float d = myrow[k * K1_FRAME_WORDS], w = myrow[k * K1_FRAME_WORDS + 7];
                        tw = xadd(tw, w);
                        // robustincc(1.0 - dot) * weight: the subtraction is done in double, narrowed (optim.cpp:690)
                        const float x = __double2float_rn(1.0 - (double)d);
                        acc = xadd(acc, xmul(robustincc(x), w));
                    }
                }
                score = (tw == 0.0f) ? 2.0f : xdiv(acc, tw);
            }
            incc_out[h] = score;
            if (ncc_out) ncc_out[h] = xsub(1.0f, unrobustincc(score));
        }
        __syncwarp();
    }
}

// ---- probes for parity tests ------------------------------------------------------------------------
__global__ void k_probe(const Params p, int n, const int* __restrict__ view, const float4* __restrict__ coord,
                        const float4* __restrict__ normal, float* project3, float* unit1, float4* px4, float4* py4,
                        int* cell_ixy2, int* cell_ok) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = view[i];
    const ViewConst& vc = p.views[v];
    const float4 c = coord[i];
    const V4 X{c.x, c.y, c.z, c.w};
    const V3 ic = project(vc.P, X);
    if (project3) { project3[3 * i] = ic.x; project3[3 * i + 1] = ic.y; project3[3 * i + 2] = ic.z; }
    if (unit1) unit1[i] = get_unit(vc, X, p.level_scale);
    if (px4 && py4 && normal) {
        const float4 m = normal[i];
        V4 px, py;
        get_paxes(vc, X, V4{m.x, m.y, m.z, m.w}, p.level_scale, px, py);
        px4[i] = make_float4(px.x, px.y, px.z, px.w);
        py4[i] = make_float4(py.x, py.y, py.z, py.w);
    }
    if (cell_ixy2) {
        const int ix = cell_of(ic.x, p.csize), iy = cell_of(ic.y, p.csize);
        cell_ixy2[2 * i] = ix; cell_ixy2[2 * i + 1] = iy;
        if (cell_ok) cell_ok[i] = (0 <= ix && ix < vc.gw && 0 <= iy && iy < vc.gh) ? 1 : 0;
    }
}

// Camera::unproject (camera.cpp:329-337) at the working level, one thread per point: the routine Propagate::generatePatch goes through
__global__ void k_probe_unproject(const Params p, int n, const int* __restrict__ view, const float* __restrict__ icoord3, float4* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const ViewConst& vc = p.views[view[i]];
    const V4 X = unproject_rows(vc.P, vc.Minv, V3{icoord3[3 * i], icoord3[3 * i + 1], icoord3[3 * i + 2]});
    out[i] = make_float4(X.x, X.y, X.z, X.w);
}

}  // namespace pmk
