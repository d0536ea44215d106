// mvskit_b200/csrc/pmk_ncc.cuh -- K0 (pyramid) and K1 (batched hypothesis NCC) kernels.
#pragma once

#include "pmk_device.cuh"

namespace pmk {

// =====================================================================================================
// K0: image pyramid.  Image::buildImagePyramid (image/image.cpp:245-315), filter == 0.
// The 4x4 [1 3 3 1]x[1 3 3 1]/64 taps times u8 values are exact in fp32 and so are their partial sums
// (multiples of 1/64 below 256), so any summation order reproduces the reference bit for bit; the
// result is re-rounded to an integer (floorf(c + 0.5f)) and stored as float RGBX.
// =====================================================================================================
__global__ void k0_u8_to_rgbx(const uint8_t* __restrict__ src, float4* __restrict__ dst, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const uint8_t* p = src + (size_t)i * 3;
    dst[i] = make_float4((float)p[0], (float)p[1], (float)p[2], 0.0f);
}

__global__ void k0_downsample(const float4* __restrict__ src, int Wp, int Hp, float4* __restrict__ dst, int W, int H) {
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
            const float4 t = __ldg(src + (size_t)yt * Wp + xt);
            const float wgt = k4[i + 1] * k4[j + 1];
            r += wgt * t.x; g += wgt * t.y; b += wgt * t.z;
        }
    }
    dst[(size_t)y * W + x] = make_float4(floorf(r + 0.5f), floorf(g + 0.5f), floorf(b + 0.5f), 0.0f);
}

__global__ void k0_rgbx_to_u8(const float4* __restrict__ src, uint8_t* __restrict__ dst, int npix) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 t = src[i];
    dst[(size_t)i * 3] = (uint8_t)t.x; dst[(size_t)i * 3 + 1] = (uint8_t)t.y; dst[(size_t)i * 3 + 2] = (uint8_t)t.z;
}

// =====================================================================================================
// K1: batched hypothesis NCC = PatchManager::computeNcc (patch_manager.cpp:401-404).
//
// Work decomposition (one warp owns a batch of 32 hypotheses, no block-level synchronisation):
//   phase A  lane = hypothesis      getPAxes in the reference view                      (optim.cpp:67-84)
//   phase B  lane = hypothesis,     loop over its <= tau views: sampling frame (angle gate, 3 projections,
//            pyramid level, getTexSafe) and computeUnits weight; frames go to warp-private shared memory
//                                                                                      (optim.cpp:790-833,895-915,109-132)
//   phase C  warp = hypothesis,     lane = sample of the wsize x wsize lattice: bilinear gather, normalise,
//            dot against the reference view with warp-shuffle reductions               (optim.cpp:835-842,917-940,601-609)
//   phase D  lane = hypothesis      robust weighted mean                               (optim.cpp:686-703)
// Phases A, B and D keep all 32 lanes busy with distinct hypotheses; only phase C is per-hypothesis.
// =====================================================================================================
constexpr int K1_WARPS = 8;            // warps per CTA
constexpr int K1_FRAME_WORDS = 8;      // tlx tly dxx dxy dyx dyy (level | view << 4) weight

// Butterfly reduction of four values at once: 6 shuffles leave the totals of (a, b, c, d) in lanes
// 0, 8, 16 and 24; three more broadcast the first three to every lane.
__device__ __forceinline__ void warp_sum3_all(int lane, float& a, float& b, float& c) {
    const bool h16 = lane & 16, h8 = lane & 8;
    float k0 = h16 ? c : a, k1 = h16 ? 0.0f : b;
    const float s0 = h16 ? a : c, s1 = h16 ? b : 0.0f;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    float k = h8 ? k1 : k0;
    const float s = h8 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, s, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    a = __shfl_sync(0xffffffffu, k, 0);
    b = __shfl_sync(0xffffffffu, k, 8);
    c = __shfl_sync(0xffffffffu, k, 16);
}

// Two values: 5 shuffles; the total of `a` lands in lanes 0-15, the total of `b` in lanes 16-31.
__device__ __forceinline__ float warp_sum2_split(int lane, float a, float b) {
    const bool h16 = lane & 16;
    float k = h16 ? b : a;
    const float s = h16 ? a : b;
    k += __shfl_xor_sync(0xffffffffu, s, 16);
    k += __shfl_xor_sync(0xffffffffu, k, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

template <int WS>
__global__ void __launch_bounds__(K1_WARPS * 32, 3) k1_ncc(const Params p, int n, const float4* __restrict__ coord,
                                                           const float4* __restrict__ normal, const int* __restrict__ views,
                                                           const int* __restrict__ nviews, int stride,
                                                           float* __restrict__ incc_out, float* __restrict__ ncc_out,
                                                           int* __restrict__ levels_out, unsigned int* __restrict__ next_batch) {
    constexpr int NSAMP = WS * WS;
    constexpr int NS = (NSAMP + 31) / 32;                 // sample slots per lane
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

    // lattice coordinates of this lane's sample slots; slots past the lattice re-sample the last point
    // (same address for those lanes: a broadcast) and are masked out of every sum
    float lx[NS], ly[NS], lm[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        const int s = lane + 32 * q;
        const int sc = s < NSAMP ? s : NSAMP - 1;
        lx[q] = (float)(sc % WS); ly[q] = (float)(sc / WS); lm[q] = s < NSAMP ? 1.0f : 0.0f;
    }

    // batches are handed out through a global counter: hypotheses differ a lot in cost (0..tau valid views)
    for (;;) {
        int batch = 0;
        if (lane == 0) batch = (int)atomicAdd(next_batch, 1u);
        batch = __shfl_sync(0xffffffffu, batch, 0);
        if (batch >= nbatch) break;
        const int h = batch * 32 + lane;
        const bool live = h < n;
        // ---------------- phase A/B: lane = hypothesis ----------------
        V4 X{0.f, 0.f, 0.f, 1.f}, N{0.f, 0.f, 1.f, 0.f};
        int nv = 0;
        if (live) {
            const float4 c = __ldg(coord + h), m = __ldg(normal + h);
            X = V4{c.x, c.y, c.z, c.w}; N = V4{m.x, m.y, m.z, m.w};
            nv = __ldg(nviews + h);
        }
        const int sz = min(p.tau, nv);
        const bool usable = live && nv >= 2;              // optim.cpp:631,643: fewer than 2 images -> 2.0
        int valid_mask = 0;
        float* myrow = frames + (size_t)lane * fstride;
        {
            const int* vrow = views + (size_t)h * stride;
            V4 px{0.f, 0.f, 0.f, 0.f}, py{0.f, 0.f, 0.f, 0.f};
            bool ref_ok = false;
            if (usable) {
                const int ref = __ldg(vrow);
                ref_ok = ref >= 0 && ref < p.nviews;
                if (ref_ok) get_paxes(p.views[ref], X, N, p.level_scale, px, py);
            }
            float unit0 = 1.0f;
#pragma unroll 1
            for (int k = 0; k < p.tau; ++k) {
                Frame f; f.level = -1; f.tlx = f.tly = f.dxx = f.dxy = f.dyx = f.dyy = 0.0f;
                float w = 0.0f;
                int vk = 0;
                if (ref_ok && k < sz) {
                    const int v = __ldg(vrow + k);
                    if (v >= 0 && v < p.nviews) {
                        vk = v;
                        const ViewConst& vc = p.views[v];
                        f = make_frame(p, vc, X, N, px, py);
                        const float u = view_unit(p, vc, X, N);
                        if (k == 0) { unit0 = u; w = 1.0f; }
                        else w = min_std(1.0f, xdiv(unit0, u));                 // optim.cpp:944-946
                    }
                }
                if (live && levels_out) levels_out[(size_t)h * p.tau + k] = f.level;
                if (f.level >= 0) valid_mask |= 1 << k;
                float* fr = myrow + k * K1_FRAME_WORDS;
                *reinterpret_cast<float4*>(fr) = make_float4(f.tlx, f.tly, f.dxx, f.dxy);
                *reinterpret_cast<float4*>(fr + 4) = make_float4(f.dyx, f.dyy, __int_as_float((f.level & 15) | (vk << 4)), w);
            }
        }
        __syncwarp();

        // ---------------- phase C: warp = hypothesis, lane = sample ----------------
        // a hypothesis needs the gather only if its reference view and at least one other are valid
        const bool need = usable && (valid_mask & 1) && (valid_mask & ~1);
        unsigned todo = __ballot_sync(0xffffffffu, need);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int vm = __shfl_sync(0xffffffffu, valid_mask, j);
            float* row = frames + (size_t)j * fstride;
            float t0[NS][3];                                   // centred (not yet scaled) reference texture
            float inv_msd0 = 1.0f;
#pragma unroll 1
            for (int k = 0; k < p.tau; ++k) {
                if (!((vm >> k) & 1)) continue;                // warp-uniform
                const float* fr = row + k * K1_FRAME_WORDS;
                const float4 fa = *reinterpret_cast<const float4*>(fr);
                const float4 fb = *reinterpret_cast<const float4*>(fr + 4);
                const int packed = __float_as_int(fb.z);
                const int lvl = packed & 15;
                const ViewConst& vc = p.views[packed >> 4];
                const float4* img = vc.img[lvl];
                const int W = vc.w[lvl];
                float t[NS][3];
                float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int q = 0; q < NS; ++q) {
                    // Vector3f samp = tl + dx * x + dy * y   (optim.cpp:837); tolerance domain -> FMAs
                    const float sx = fmaf(fb.x, ly[q], fmaf(fa.z, lx[q], fa.x));
                    const float sy = fmaf(fb.y, ly[q], fmaf(fa.w, lx[q], fa.y));
                    bilinear(img, W, sx, sy, t[q][0], t[q][1], t[q][2]);
                    s0 = fmaf(t[q][0], lm[q], s0); s1 = fmaf(t[q][1], lm[q], s1); s2 = fmaf(t[q][2], lm[q], s2);
                }
                // Optim::normalize (optim.cpp:917-940): per-channel mean, joint variance
                warp_sum3_all(lane, s0, s1, s2);
                const float a0 = s0 * INV_NSAMP, a1 = s1 * INV_NSAMP, a2 = s2 * INV_NSAMP;
                float ssd = 0.f, dp = 0.f;
#pragma unroll
                for (int q = 0; q < NS; ++q) {
                    t[q][0] = (t[q][0] - a0) * lm[q]; t[q][1] = (t[q][1] - a1) * lm[q]; t[q][2] = (t[q][2] - a2) * lm[q];
                    ssd = fmaf(t[q][0], t[q][0], fmaf(t[q][1], t[q][1], fmaf(t[q][2], t[q][2], ssd)));
                }
                if (k == 0) {
                    ssd = warp_sum(ssd);
                    const float var = ssd * INV_3NSAMP;
                    inv_msd0 = var > 0.0f ? rsqrtf(var) : 1.0f;            // msd == 0 -> 1 (optim.cpp:934-936)
#pragma unroll
                    for (int q = 0; q < NS; ++q) { t0[q][0] = t[q][0]; t0[q][1] = t[q][1]; t0[q][2] = t[q][2]; }
                } else {
                    // Optim::dot (optim.cpp:601-609): sum(tex0 . tex_k) / (3 * sz)
#pragma unroll
                    for (int q = 0; q < NS; ++q)
                        dp = fmaf(t0[q][0], t[q][0], fmaf(t0[q][1], t[q][1], fmaf(t0[q][2], t[q][2], dp)));
                    const float r = warp_sum2_split(lane, ssd, dp);        // lanes 0-15: ssd, lanes 16-31: dp
                    const float dpt = __shfl_sync(0xffffffffu, r, 16);
                    if (lane == 0) {
                        const float var = r * INV_3NSAMP;
                        const float inv_msd = var > 0.0f ? rsqrtf(var) : 1.0f;
                        row[k * K1_FRAME_WORDS] = dpt * inv_msd0 * inv_msd * INV_3NSAMP;   // overwrites tlx (consumed)
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

}  // namespace pmk
