/* oracle/pm_oracle.c -- TEST INFRASTRUCTURE ONLY.  See pm_oracle.h for scope and pinning.
 *
 * Arithmetic contract (matches oracle/shim/Eigen/Dense, which defines it for the reference build):
 * every reduction is strict left-to-right with one IEEE rounding per operation (compile with
 * -ffp-contract=off), vector/scalar is a per-component divide, and wherever the reference's C++
 * promotes to double (literals such as 2.0 / 1.0, unqualified log()) this file does the same.
 */
#include "pm_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- small vector helpers (left-to-right) ------------------------------------------------------- */
static float dot3(const float* a, const float* b) { float acc = a[0] * b[0]; acc = acc + a[1] * b[1]; acc = acc + a[2] * b[2]; return acc; }
static float dot4(const float* a, const float* b) { float acc = a[0] * b[0]; acc = acc + a[1] * b[1]; acc = acc + a[2] * b[2]; acc = acc + a[3] * b[3]; return acc; }
static float norm3(const float* a) { return sqrtf(dot3(a, a)); }
static float norm4(const float* a) { return sqrtf(dot4(a, a)); }
static void cross3(const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static float fmin_std(float a, float b) { return (b < a) ? b : a; }   /* std::min(a, b) */
static float fmax_std(float a, float b) { return (a < b) ? b : a; }   /* std::max(a, b) */

/* 3x3 inverse exactly as oracle/shim/Eigen/Dense Mat::inverse() */
static void inv3(const float* m, float* r) {
    const float c00 = m[4] * m[8] - m[5] * m[7];
    const float c01 = m[5] * m[6] - m[3] * m[8];
    const float c02 = m[3] * m[7] - m[4] * m[6];
    const float det = (m[0] * c00 + m[1] * c01) + m[2] * c02;
    r[0] = c00 / det;
    r[1] = (m[2] * m[7] - m[1] * m[8]) / det;
    r[2] = (m[1] * m[5] - m[2] * m[4]) / det;
    r[3] = c01 / det;
    r[4] = (m[0] * m[8] - m[2] * m[6]) / det;
    r[5] = (m[2] * m[3] - m[0] * m[5]) / det;
    r[6] = c02 / det;
    r[7] = (m[1] * m[6] - m[0] * m[7]) / det;
    r[8] = (m[0] * m[4] - m[1] * m[3]) / det;
}

/* ---- scene ------------------------------------------------------------------------------------- */
pmo_scene* pmo_scene_create(int nviews, int level, int csize, int wsize, int min_image_num, float ncc_threshold) {
    pmo_scene* s = (pmo_scene*)calloc(1, sizeof(pmo_scene));
    s->nviews = nviews; s->level = level; s->nlevels = level + 3;           /* pmmvps.cpp:36 */
    s->csize = csize; s->wsize = wsize; s->min_image_num = min_image_num;
    s->tau = (min_image_num * 2 < nviews) ? min_image_num * 2 : nviews;     /* pmmvps.cpp:32 */
    s->depth = 0;
    /* thresholds, pmmvps.cpp:54-67 and option.cpp:30-31 */
    s->ncc_threshold = ncc_threshold;
    s->ncc_threshold_before = ncc_threshold - 0.3f;
    s->angle_threshold0 = 60.0f * M_PI / 180.0f;
    s->angle_threshold1 = 60.0f * M_PI / 180.0f;
    s->max_angle_threshold = 10.0f * M_PI / 180.0f;
    s->quad_threshold = 2.5f;
    s->neighbor_threshold = 0.5f; s->neighbor_threshold1 = 1.0f; s->neighbor_threshold2 = 1.0f;
    const size_t nl = (size_t)nviews * s->nlevels;
    s->P = (float*)calloc(nl * 12, sizeof(float));
    s->Minv = (float*)calloc(nl * 9, sizeof(float));
    s->center = (float*)calloc((size_t)nviews * 4, sizeof(float));
    s->oaxis = (float*)calloc((size_t)nviews * 4, sizeof(float));
    s->xaxis = (float*)calloc((size_t)nviews * 3, sizeof(float));
    s->yaxis = (float*)calloc((size_t)nviews * 3, sizeof(float));
    s->zaxis = (float*)calloc((size_t)nviews * 3, sizeof(float));
    s->ipscale = (float*)calloc((size_t)nviews, sizeof(float));
    s->img = (unsigned char**)calloc(nl, sizeof(unsigned char*));
    s->w = (int*)calloc(nl, sizeof(int)); s->h = (int*)calloc(nl, sizeof(int));
    s->gw = (int*)calloc((size_t)nviews, sizeof(int)); s->gh = (int*)calloc((size_t)nviews, sizeof(int));
    s->mask = (unsigned char**)calloc(nl, sizeof(unsigned char*));
    return s;
}

void pmo_scene_destroy(pmo_scene* s) {
    if (!s) return;
    for (int i = 0; i < s->nviews * s->nlevels; ++i) { free(s->img[i]); free(s->mask[i]); }
    free(s->mask);
    free(s->P); free(s->Minv); free(s->center); free(s->oaxis); free(s->xaxis); free(s->yaxis); free(s->zaxis);
    free(s->ipscale); free(s->img); free(s->w); free(s->h); free(s->gw); free(s->gh); free(s);
}

void pmo_update_threshold(pmo_scene* s) {   /* pmmvps.cpp:70-74, :106 */
    s->ncc_threshold -= 0.05f;
    s->ncc_threshold_before -= 0.05f;
    s->depth += 1;
}

void pmo_set_camera(pmo_scene* s, int view, const float* P12) {
    float* P = s->P + (size_t)view * s->nlevels * 12;
    memcpy(P, P12, 12 * sizeof(float));
    for (int l = 1; l < s->nlevels; ++l) {                                  /* camera.cpp:95-99 */
        float* q = P + l * 12;
        memcpy(q, q - 12, 12 * sizeof(float));
        for (int c = 0; c < 8; ++c) q[c] = q[c] / 2.0f;
    }
    for (int l = 0; l < s->nlevels; ++l) {                                  /* camera.cpp:331-332 */
        const float* q = P + l * 12;
        const float M[9] = {q[0], q[1], q[2], q[4], q[5], q[6], q[8], q[9], q[10]};
        inv3(M, s->Minv + ((size_t)view * s->nlevels + l) * 9);
    }
    /* camera.cpp:68-69: m_oaxis = row(2) / ||row(2).head(3)|| */
    float* oa = s->oaxis + view * 4;
    const float on = norm3(P + 8);
    for (int c = 0; c < 4; ++c) oa[c] = P[8 + c] / on;
    /* camera.cpp:295-308: centre = (-M^-1) * q, w = 1 */
    {
        const float* Mi = s->Minv + (size_t)view * s->nlevels * 9;
        const float q[3] = {P[3], P[7], P[11]};
        float* c = s->center + view * 4;
        for (int r = 0; r < 3; ++r) {
            float acc = (-Mi[3 * r]) * q[0];
            acc = acc + (-Mi[3 * r + 1]) * q[1];
            acc = acc + (-Mi[3 * r + 2]) * q[2];
            c[r] = acc;
        }
        c[3] = 1.0f;
    }
    /* optim.cpp:47-54 */
    float* xa = s->xaxis + view * 3; float* ya = s->yaxis + view * 3; float* za = s->zaxis + view * 3;
    za[0] = oa[0]; za[1] = oa[1]; za[2] = oa[2];
    const float x0[3] = {P[0], P[1], P[2]};
    cross3(za, x0, ya);
    const float yn = norm3(ya);
    ya[0] = ya[0] / yn; ya[1] = ya[1] / yn; ya[2] = ya[2] / yn;
    cross3(ya, za, xa);
    /* optim.cpp:57-64 */
    const float xa4[4] = {xa[0], xa[1], xa[2], 0.0f};
    const float ya4[4] = {ya[0], ya[1], ya[2], 0.0f};
    const float fx = dot4(P, xa4);
    const float fy = dot4(P + 4, ya4);
    s->ipscale[view] = fx + fy;
}

void pmo_set_image(pmo_scene* s, int view, const unsigned char* rgb, int w, int h) {
    const size_t base = (size_t)view * s->nlevels;
    s->w[base] = w; s->h[base] = h;
    for (int l = 1; l < s->nlevels; ++l) {                                  /* image.cpp:135-138 */
        s->w[base + l] = s->w[base + l - 1] / 2;
        s->h[base + l] = s->h[base + l - 1] / 2;
    }
    free(s->img[base]);
    s->img[base] = (unsigned char*)malloc((size_t)w * h * 3);
    memcpy(s->img[base], rgb, (size_t)w * h * 3);
    /* image.cpp:245-315, filter == 0.  mask = [1 3 3 1]x[1 3 3 1] / 64 (exact), divided again by
       mask.sum() == 1.0f exactly; every product and partial sum is a multiple of 1/64 below 2^8, so
       the filter is exact in float and order-independent. */
    static const float k4[4] = {1.0f, 3.0f, 3.0f, 1.0f};
    float mask[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) mask[i][j] = (k4[i] * k4[j]) / 64.0f;
    for (int l = 1; l < s->nlevels; ++l) {
        const int W = s->w[base + l], H = s->h[base + l], Wp = s->w[base + l - 1], Hp = s->h[base + l - 1];
        const unsigned char* src = s->img[base + l - 1];
        free(s->img[base + l]);
        unsigned char* dst = s->img[base + l] = (unsigned char*)malloc((size_t)W * H * 3 + 1);
        for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) {
            float col[3] = {0.0f, 0.0f, 0.0f};
            for (int i = -1; i < 3; ++i) {
                const int yt = 2 * y + i;
                if (yt < 0 || Hp - 1 < yt) continue;
                for (int j = -1; j < 3; ++j) {
                    const int xt = 2 * x + j;
                    if (xt < 0 || Wp - 1 < xt) continue;
                    const unsigned char* p = src + ((size_t)yt * Wp + xt) * 3;
                    col[0] += mask[i + 1][j + 1] * (float)p[0];
                    col[1] += mask[i + 1][j + 1] * (float)p[1];
                    col[2] += mask[i + 1][j + 1] * (float)p[2];
                }
            }
            unsigned char* q = dst + ((size_t)y * W + x) * 3;
            for (int c = 0; c < 3; ++c) q[c] = (unsigned char)((int)floorf(col[c] / 1.0f + 0.5f));
        }
    }
    /* patch_manager.cpp:36-37 */
    s->gw[view] = (s->w[base + s->level] + s->csize - 1) / s->csize;
    s->gh[view] = (s->h[base + s->level] + s->csize - 1) / s->csize;
}

/* ---- camera ------------------------------------------------------------------------------------ */
void pmo_project(const pmo_scene* s, int view, const float* X, int level, float* out) {
    const float* P = s->P + ((size_t)view * s->nlevels + level) * 12;
    float ic[3];
    for (int r = 0; r < 3; ++r) ic[r] = dot4(P + 4 * r, X);
    if (ic[2] <= 0.0) { out[0] = -65535.0f; out[1] = -65535.0f; out[2] = -1.0f; return; }   /* camera.cpp:313-316 */
    const float z = ic[2];
    out[0] = ic[0] / z; out[1] = ic[1] / z; out[2] = ic[2] / z;
    const float lo = (float)(INT_MIN + 3.0f), hi = (float)(INT_MAX - 3.0f);                    /* camera.cpp:322-323 */
    out[0] = fmax_std(lo, fmin_std(hi, out[0]));
    out[1] = fmax_std(lo, fmin_std(hi, out[1]));
}

void pmo_unproject(const pmo_scene* s, int view, const float* ic, int level, float* out) {
    const float* P = s->P + ((size_t)view * s->nlevels + level) * 12;
    const float* Mi = s->Minv + ((size_t)view * s->nlevels + level) * 9;
    const float b[3] = {ic[0] - P[3], ic[1] - P[7], ic[2] - P[11]};
    for (int r = 0; r < 3; ++r) out[r] = dot3(Mi + 3 * r, b);
    out[3] = 1.0f;
}

float pmo_get_unit(const pmo_scene* s, int view, const float* X) {
    const float* c = s->center + view * 4;
    const float d[4] = {X[0] - c[0], X[1] - c[1], X[2] - c[2], X[3] - c[3]};
    const float fz = norm4(d);
    const float ipscale = s->ipscale[view];
    if (ipscale == 0.0f) return 1.0;
    return (float)(2.0 * fz * (0x0001 << s->level) / ipscale);
}

void pmo_get_paxes(const pmo_scene* s, int view, const float* X, const float* N, float* px, float* py) {
    const float pscale = pmo_get_unit(s, view, X);
    float y3[3], x3[3];
    cross3(N, s->xaxis + view * 3, y3);
    const float yn = norm3(y3);
    y3[0] = y3[0] / yn; y3[1] = y3[1] / yn; y3[2] = y3[2] / yn;
    cross3(y3, N, x3);
    for (int i = 0; i < 3; ++i) { px[i] = x3[i] * pscale; py[i] = y3[i] * pscale; }
    px[3] = 0.0f * pscale; py[3] = 0.0f * pscale;
    float c0[3], c1[3], Xp[4], d[3];
    pmo_project(s, view, X, s->level, c0);
    for (int i = 0; i < 4; ++i) Xp[i] = X[i] + px[i];
    pmo_project(s, view, Xp, s->level, c1);
    for (int i = 0; i < 3; ++i) d[i] = c1[i] - c0[i];
    const float xdis = norm3(d);
    for (int i = 0; i < 4; ++i) Xp[i] = X[i] + py[i];
    pmo_project(s, view, Xp, s->level, c1);
    for (int i = 0; i < 3; ++i) d[i] = c1[i] - c0[i];
    const float ydis = norm3(d);
    for (int i = 0; i < 4; ++i) { px[i] = px[i] / xdis; py[i] = py[i] / ydis; }
}

/* ---- image ------------------------------------------------------------------------------------- */
void pmo_get_color(const pmo_scene* s, int view, float x, float y, int level, float* rgb) {
    const size_t b = (size_t)view * s->nlevels + level;
    const int W = s->w[b];
    const unsigned char* img = s->img[b];
    const int lx = (int)x, ly = (int)y;
    const float dx1 = x - lx, dx0 = 1.0f - dx1;
    const float dy1 = y - ly, dy0 = 1.0f - dy1;
    const float f00 = dx0 * dy0, f01 = dx0 * dy1, f10 = dx1 * dy0, f11 = dx1 * dy1;
    const unsigned char* p0 = img + 3 * ((size_t)ly * W + lx);
    const unsigned char* p1 = p0 + 3 * (size_t)W;
    for (int c = 0; c < 3; ++c) {
        float v = 0.0f;
        v += p0[c] * f00 + p1[c] * f01;
        v += p0[3 + c] * f10 + p1[3 + c] * f11;
        rgb[c] = v;
    }
}

int pmo_level_diff(float ratio) { return (int)floorf(log(ratio) / log(2.0f) + 0.5f); }

static float my_pow2(int d) {   /* optim.cpp:785-788 */
    static const float scales[] = {0.0625, 0.125, 0.25, 0.5, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024};
    return scales[d + 4];
}

static int get_tex_safe(const pmo_scene* s, int view, int size, const float* c, const float* dx, const float* dy, int level) {
    const int margin = size / 2;
    float tl[2], tr[2], bl[2], br[2];
    for (int i = 0; i < 2; ++i) {
        tl[i] = c[i] - dx[i] * margin - dy[i] * margin;
        tr[i] = c[i] + dx[i] * margin - dy[i] * margin;
        bl[i] = c[i] - dx[i] * margin + dy[i] * margin;
        br[i] = c[i] + dx[i] * margin + dy[i] * margin;
    }
    const float minx = fmin_std(tl[0], fmin_std(tr[0], fmin_std(bl[0], br[0])));
    const float maxx = fmax_std(tl[0], fmax_std(tr[0], fmax_std(bl[0], br[0])));
    const float miny = fmin_std(tl[1], fmin_std(tr[1], fmin_std(bl[1], br[1])));
    const float maxy = fmax_std(tl[1], fmax_std(tr[1], fmax_std(bl[1], br[1])));
    const int margin2 = 2;
    const size_t b = (size_t)view * s->nlevels + level;
    if (minx < margin2 || s->w[b] - 1 - margin2 <= maxx || miny < margin2 || s->h[b] - 1 - margin2 <= maxy) return -1;
    return 0;
}

int pmo_get_tex(const pmo_scene* s, const float* X, const float* px, const float* py, const float* N, int view,
                float* tex, int* level_out) {
    if (level_out) *level_out = -1;
    const float* cen = s->center + view * 4;
    float ray[4] = {cen[0] - X[0], cen[1] - X[1], cen[2] - X[2], cen[3] - X[3]};
    const float rn = norm4(ray);
    for (int i = 0; i < 4; ++i) ray[i] = ray[i] / rn;
    const float weight = fmax_std(0.0f, dot4(ray, N));
    if (weight < cosf(s->angle_threshold1)) return -1;                       /* optim.cpp:796-798 */
    const int size = s->wsize, margin = size / 2;
    float c[3], dx[3], dy[3], t[3], Xp[4];
    pmo_project(s, view, X, s->level, c);
    for (int i = 0; i < 4; ++i) Xp[i] = X[i] + px[i];
    pmo_project(s, view, Xp, s->level, t);
    for (int i = 0; i < 3; ++i) dx[i] = t[i] - c[i];
    for (int i = 0; i < 4; ++i) Xp[i] = X[i] + py[i];
    pmo_project(s, view, Xp, s->level, t);
    for (int i = 0; i < 3; ++i) dy[i] = t[i] - c[i];
    const float ratio = (norm3(dx) + norm3(dy)) / 2.0f;
    int levelDiff = pmo_level_diff(ratio);                                    /* optim.cpp:808 */
    { const int a = (2 < levelDiff) ? 2 : levelDiff; levelDiff = (-s->level < a) ? a : -s->level; }
    const float scale = my_pow2(levelDiff);
    const int newLevel = s->level + levelDiff;
    for (int i = 0; i < 3; ++i) { c[i] = c[i] / scale; dx[i] = dx[i] / scale; dy[i] = dy[i] / scale; }
    if (get_tex_safe(s, view, size, c, dx, dy, newLevel) == -1) return -1;
    if (level_out) *level_out = newLevel;
    float tl[2];
    for (int i = 0; i < 2; ++i) tl[i] = c[i] - dx[i] * margin - dy[i] * margin;
    for (int y = 0; y < size; ++y) for (int x = 0; x < size; ++x) {
        const float sx = tl[0] + dx[0] * x + dy[0] * y;
        const float sy = tl[1] + dx[1] * x + dy[1] * y;
        pmo_get_color(s, view, sx, sy, newLevel, tex + 3 * (y * size + x));
    }
    return 0;
}

void pmo_normalize(float* tex, int sz) {
    float ave[3] = {0.0f, 0.0f, 0.0f};
    for (int i = 0; i < sz; ++i) for (int c = 0; c < 3; ++c) ave[c] = ave[c] + tex[3 * i + c];
    for (int c = 0; c < 3; ++c) ave[c] = ave[c] / sz;
    float ssd = 0.0f;
    for (int i = 0; i < sz; ++i) {
        const float d[3] = {tex[3 * i] - ave[0], tex[3 * i + 1] - ave[1], tex[3 * i + 2] - ave[2]};
        ssd += dot3(d, d);
    }
    float msd = sqrtf(ssd / (3 * sz));
    if (msd == 0.0f) msd = 1.0f;
    for (int i = 0; i < sz; ++i) for (int c = 0; c < 3; ++c) tex[3 * i + c] = (tex[3 * i + c] - ave[c]) / msd;
}

float pmo_dot(const float* t0, const float* t1, int sz) {
    float ssd = 0.0f;
    for (int i = 0; i < sz; ++i) ssd += dot3(t0 + 3 * i, t1 + 3 * i);
    return ssd / (3 * sz);
}

float pmo_robustincc(float incc) { return incc / (1 + 3 * incc); }
float pmo_unrobustincc(float rincc) { return rincc / (1 - 3 * rincc); }

void pmo_compute_weights(const pmo_scene* s, const float* X, const float* N, const int* views, int nviews, float* w) {
    for (int i = 0; i < nviews; ++i) {                                       /* optim.cpp:109-132 */
        const int v = views[i];
        float unit = pmo_get_unit(s, v, X);
        const float* cen = s->center + v * 4;
        float ray[4] = {cen[0] - X[0], cen[1] - X[1], cen[2] - X[2], cen[3] - X[3]};
        const float rn = norm4(ray);
        for (int k = 0; k < 4; ++k) ray[k] = ray[k] / rn;
        const float d = dot4(ray, N);
        if (0.0f < d) unit /= d; else unit = INT_MAX / 2;
        w[i] = unit;
    }
    for (int i = 1; i < nviews; ++i) w[i] = fmin_std(1.0f, w[0] / w[i]);    /* optim.cpp:942-948 */
    if (nviews > 0) w[0] = 1.0f;
}

#define PMO_MAXTEX (11 * 11 * 3)

float pmo_compute_incc(const pmo_scene* s, const float* X, const float* N, const int* views, int nviews,
                       const float* weights, int robust, int* levels_out) {
    if (nviews < 2) return 2.0;
    float px[4], py[4];
    pmo_get_paxes(s, views[0], X, N, px, py);
    const int sz = (s->tau < nviews) ? s->tau : nviews;
    const int tsz = s->wsize * s->wsize;
    float* texs = (float*)malloc((size_t)sz * tsz * 3 * sizeof(float));
    int ok[64] = {0};
    for (int i = 0; i < sz; ++i) {
        int lvl;
        ok[i] = pmo_get_tex(s, X, px, py, N, views[i], texs + (size_t)i * tsz * 3, &lvl) == 0;
        if (levels_out) levels_out[i] = lvl;
        if (ok[i]) pmo_normalize(texs + (size_t)i * tsz * 3, tsz);
    }
    float score = 0.0f;
    if (!ok[0]) { free(texs); return 2.0; }
    float totalWeight = 0.0f;
    for (int i = 1; i < sz; ++i) {
        if (!ok[i]) continue;
        totalWeight += weights[i];
        const float d = pmo_dot(texs, texs + (size_t)i * tsz * 3, tsz);
        if (robust) score += pmo_robustincc(1.0 - d) * weights[i];
        else score += (1.0 - d) * weights[i];
    }
    if (totalWeight == 0.0f) score = 2.0f; else score /= totalWeight;
    free(texs);
    return score;
}

void pmo_compute_ncc(const pmo_scene* s, int n, const float* X, const float* N, const int* views, const int* nviews,
                     int stride, float* incc, float* ncc, int* levels) {
    float w[256];
    for (int i = 0; i < n; ++i) {
        const int* v = views + (size_t)i * stride;
        if (levels) for (int k = 0; k < s->tau; ++k) levels[(size_t)i * s->tau + k] = -1;
        pmo_compute_weights(s, X + 4 * i, N + 4 * i, v, nviews[i], w);
        const float sc = pmo_compute_incc(s, X + 4 * i, N + 4 * i, v, nviews[i], w, 1, levels ? levels + (size_t)i * s->tau : 0);
        if (incc) incc[i] = sc;
        if (ncc) ncc[i] = 1.0f - pmo_unrobustincc(sc);
    }
}

static void grab_all(const pmo_scene* s, const float* X, const float* N, const int* views, int nviews, float* texs, int* ok) {
    float px[4], py[4];
    pmo_get_paxes(s, views[0], X, N, px, py);
    const int tsz = s->wsize * s->wsize;
    for (int i = 0; i < nviews; ++i) {
        ok[i] = pmo_get_tex(s, X, px, py, N, views[i], texs + (size_t)i * tsz * 3, 0) == 0;
        if (ok[i]) pmo_normalize(texs + (size_t)i * tsz * 3, tsz);
    }
}

void pmo_set_inccs(const pmo_scene* s, const float* X, const float* N, const int* views, int nviews, int robust, float* out) {
    const int tsz = s->wsize * s->wsize;
    float* texs = (float*)malloc((size_t)nviews * tsz * 3 * sizeof(float));
    int* ok = (int*)malloc((size_t)nviews * sizeof(int));
    grab_all(s, X, N, views, nviews, texs, ok);
    if (!ok[0]) { for (int i = 0; i < nviews; ++i) out[i] = 2.0f; free(texs); free(ok); return; }
    for (int i = 0; i < nviews; ++i) {
        if (i == 0) out[i] = 0.0f;
        else if (ok[i]) {
            const float d = pmo_dot(texs, texs + (size_t)i * tsz * 3, tsz);
            out[i] = robust ? pmo_robustincc(1.0f - d) : 1.0f - d;
        } else out[i] = 2.0f;
    }
    free(texs); free(ok);
}

void pmo_set_inccs_pair(const pmo_scene* s, const float* X, const float* N, const int* views, int nviews, int robust, float* out) {
    const int tsz = s->wsize * s->wsize;
    float* texs = (float*)malloc((size_t)nviews * tsz * 3 * sizeof(float));
    int* ok = (int*)malloc((size_t)nviews * sizeof(int));
    grab_all(s, X, N, views, nviews, texs, ok);
    for (int i = 0; i < nviews; ++i) {
        out[i * nviews + i] = 0.0f;
        for (int j = i + 1; j < nviews; ++j) {
            float v = 2.0f;
            if (ok[i] && ok[j]) {
                const float d = pmo_dot(texs + (size_t)i * tsz * 3, texs + (size_t)j * tsz * 3, tsz);
                v = robust ? pmo_robustincc(1.0f - d) : 1.0f - d;
            }
            out[i * nviews + j] = out[j * nviews + i] = v;
        }
    }
    free(texs); free(ok);
}

int pmo_cell(const pmo_scene* s, int view, const float* X, int* ix, int* iy) {
    float ic[3];
    pmo_project(s, view, X, s->level, ic);
    *ix = ((int)floorf(ic[0] + 0.5f)) / s->csize;
    *iy = ((int)floorf(ic[1] + 0.5f)) / s->csize;
    return (0 <= *ix && *ix < s->gw[view] && 0 <= *iy && *iy < s->gh[view]) ? 1 : 0;
}

/* ---- masks -------------------------------------------------------------------------------------------- */
void pmo_set_mask(pmo_scene* s, int view, const unsigned char* grey, int w, int h) {
    const size_t base = (size_t)view * s->nlevels;
    if (w != s->w[base] || h != s->h[base]) return;      /* the reference lets the mask's header overwrite m_widths[0] (image.cpp:146) */
    free(s->mask[base]);
    unsigned char* m0 = s->mask[base] = (unsigned char*)malloc((size_t)w * h);
    for (size_t i = 0; i < (size_t)w * h; ++i) m0[i] = (127 < (int)grey[i]) ? 255 : 0;            /* image.cpp:149-156 */
    for (int l = 1; l < s->nlevels; ++l) {                                                       /* image.cpp:717-747 */
        const int W = s->w[base + l], H = s->h[base + l], Wp = s->w[base + l - 1], Hp = s->h[base + l - 1];
        const unsigned char* src = s->mask[base + l - 1];
        free(s->mask[base + l]);
        unsigned char* dst = s->mask[base + l] = (unsigned char*)malloc((size_t)W * H + 1);
        for (int y = 0; y < H; ++y) {
            const int ys[2] = {2 * y, (Hp < 2 * y + 1) ? Hp : 2 * y + 1};                        /* :723 (min against the height itself) */
            for (int x = 0; x < W; ++x) {
                const int xs[2] = {2 * x, (Wp < 2 * x + 1) ? Wp : 2 * x + 1};
                int inside = 0;
                for (int j = 0; j < 2; ++j) for (int i = 0; i < 2; ++i) if (src[(size_t)ys[j] * Wp + xs[i]]) ++inside;
                dst[(size_t)y * W + x] = (0 < inside) ? 255 : 0;
            }
        }
    }
}

int pmo_get_mask_view(const pmo_scene* s, int view, const float* X, int level) {
    const size_t idx = (size_t)view * s->nlevels + level;
    if (!s->mask[idx]) return -1;                                                                 /* photo.cpp:45-47 */
    float ic[3];
    pmo_project(s, view, X, level, ic);
    const int ix = (int)floorf(ic[0] + 0.5f), iy = (int)floorf(ic[1] + 0.5f);                     /* image.cpp:759-760 */
    if (ix < 0 || s->w[idx] <= ix || iy < 0 || s->h[idx] <= iy) return -1;                        /* :775-778 */
    return s->mask[idx][(size_t)iy * s->w[idx] + ix];
}

int pmo_get_mask(const pmo_scene* s, const float* X, int level) {
    for (int v = 0; v < s->nviews; ++v) if (pmo_get_mask_view(s, v, X, level) == 0) return 0;     /* photoSet.cpp:224-231 */
    return -1;
}
