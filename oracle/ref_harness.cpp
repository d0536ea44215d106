// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (checker + CPU baseline; never on the product path).
//
// C-ABI harness around the UNMODIFIED reference classes (imkaywu/MVSKit, compiled from
// /root/reference by oracle/Makefile against oracle/shim/*).  Everything below calls the
// reference's own member functions; nothing re-implements them.  Where the reference keeps a
// value in a local variable (the pyramid level chosen inside Optim::getTex, optim.cpp:807-811)
// the harness re-derives it with the reference's own project()/myPow2 calls and says so.
//
// Used by: tests/ (parity), tests/golden/make_golden.py (fixture generation),
//          bench.py --impl reference and bench.py's cpu_baseline leg.
#include <algorithm>
#include <cmath>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <chrono>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <numeric>
#include <queue>
#include <random>
#include <sstream>
#include <string>
#include <vector>

// Access to Optim's per-view axes / scratch (declared protected in optim.hpp:118-139).  Member
// layout does not depend on access specifiers, and every reference TU is built without this.
#define protected public
#define private public
#include "pmmvps/pmmvps.hpp"
#undef protected
#undef private
#include "nlopt.hpp"

using Eigen::Vector3f;
using Eigen::Vector4f;

namespace {

struct NullBuf : public std::streambuf {
    int overflow(int c) override { return c; }
};
NullBuf g_nullbuf;
std::streambuf* g_cerr_saved = nullptr;

Option* g_option = nullptr;
PmMvps* g_pm = nullptr;

inline Vector4f v4(const float* p) { return Vector4f(p[0], p[1], p[2], p[3]); }

void fill_patch(Patch& patch, const float* coord, const float* normal, const int* views, int nviews) {
    patch.m_coord = v4(coord);
    patch.m_normal = v4(normal);
    patch.m_images.assign(views, views + nviews);
}

}  // namespace

extern "C" {

// 1 if an unqualified log(float) in a reference-like TU resolves to the double overload
// (decides how optim.cpp:808 rounds; see oracle/shim/compat.h).
int pmref_log_is_double(void) { return sizeof(decltype(log(1.0f))) == sizeof(double) ? 1 : 0; }

void pmref_quiet(int on) {
    if (on && !g_cerr_saved) g_cerr_saved = std::cerr.rdbuf(&g_nullbuf);
    if (!on && g_cerr_saved) { std::cerr.rdbuf(g_cerr_saved); g_cerr_saved = nullptr; }
}

void pmref_shutdown(void) {
    delete g_pm; g_pm = nullptr;
    delete g_option; g_option = nullptr;
}

// Option::init (option.cpp:35) + PmMvps::init (pmmvps.cpp:18).  prefix must end with '/'.
int pmref_init(const char* prefix, const char* option_name) {
    pmref_shutdown();
    pmref_quiet(1);
    g_option = new Option();
    g_option->init(prefix, option_name);
    g_pm = new PmMvps();
    g_pm->init(*g_option);
    return 0;
}

// what: 0 nimages, 1 level, 2 csize, 3 wsize, 4 tau, 5 minImageNum, 6 depth, 7 maxLevel(allocated)
int pmref_info(int what) {
    switch (what) {
        case 0: return g_pm->m_nimages;
        case 1: return g_pm->m_level;
        case 2: return g_pm->m_csize;
        case 3: return g_pm->m_wsize;
        case 4: return g_pm->m_tau;
        case 5: return g_pm->m_minImageNumThreshold;
        case 6: return g_pm->m_depth;
        case 7: return g_pm->m_level + 3;
    }
    return -1;
}

// what: 0 nccThreshold, 1 nccThresholdBefore, 2 angleThreshold0, 3 angleThreshold1, 4 maxAngleThreshold,
//       5 quadThreshold, 6 neighborThreshold, 7 neighborThreshold1, 8 neighborThreshold2
float pmref_threshold(int what) {
    switch (what) {
        case 0: return g_pm->m_nccThreshold;
        case 1: return g_pm->m_nccThresholdBefore;
        case 2: return g_pm->m_angleThreshold0;
        case 3: return g_pm->m_angleThreshold1;
        case 4: return g_pm->m_maxAngleThreshold;
        case 5: return g_pm->m_quadThreshold;
        case 6: return g_pm->m_neighborThreshold;
        case 7: return g_pm->m_neighborThreshold1;
        case 8: return g_pm->m_neighborThreshold2;
    }
    return 0.0f;
}

void pmref_set_depth(int depth) { g_pm->m_depth = depth; }
// PmMvps::updateThreshold (pmmvps.cpp:70-74), the reference's own
void pmref_update_threshold(void) { g_pm->updateThreshold(); }
void pmref_set_ncc_thresholds(float ncc, float before) { g_pm->m_nccThreshold = ncc; g_pm->m_nccThresholdBefore = before; }

void pmref_image_dims(int view, int level, int* w, int* h) {
    *w = g_pm->m_photoSet.getWidth(view, level);
    *h = g_pm->m_photoSet.getHeight(view, level);
}

void pmref_grid_dims(int view, int* gw, int* gh) {
    *gw = g_pm->m_patchManager.m_gwidths[view];
    *gh = g_pm->m_patchManager.m_gheights[view];
}

// u8 interleaved RGB of one pyramid level, as built by Image::buildImagePyramid (image.cpp:245-315)
void pmref_get_image(int view, int level, unsigned char* out) {
    const std::vector<unsigned char>& img = g_pm->m_photoSet.m_photos[view].m_images[level];
    std::memcpy(out, img.data(), img.size());
}

// per-view constants: level-`level` projection (camera.cpp:91-100), centre (:295-308), oaxis (:68-69),
// Optim axes + ipscale (optim.cpp:43-65)
void pmref_get_camera(int view, int level, float* P12, float* center4, float* oaxis4, float* xaxis3,
                      float* yaxis3, float* zaxis3, float* ipscale) {
    const Photo& ph = g_pm->m_photoSet.m_photos[view];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) P12[r * 4 + c] = ph.m_projections[level](r, c);
    for (int i = 0; i < 4; ++i) { center4[i] = ph.m_center(i); oaxis4[i] = ph.m_oaxis(i); }
    for (int i = 0; i < 3; ++i) {
        xaxis3[i] = g_pm->m_optim.m_xaxes[view](i);
        yaxis3[i] = g_pm->m_optim.m_yaxes[view](i);
        zaxis3[i] = g_pm->m_optim.m_zaxes[view](i);
    }
    *ipscale = g_pm->m_optim.m_ipscales[view];
}

void pmref_project(int n, const int* view, const float* coord4, int level, float* out3) {
    for (int i = 0; i < n; ++i) {
        const Vector3f ic = g_pm->m_photoSet.project(view[i], v4(coord4 + 4 * i), level);
        out3[3 * i] = ic(0); out3[3 * i + 1] = ic(1); out3[3 * i + 2] = ic(2);
    }
}

// Camera::unproject (camera.cpp:329-337)
void pmref_unproject(int n, const int* view, const float* icoord3, int level, float* out4) {
    for (int i = 0; i < n; ++i) {
        const Vector4f c = g_pm->m_photoSet.m_photos[view[i]].unproject(
            Vector3f(icoord3[3 * i], icoord3[3 * i + 1], icoord3[3 * i + 2]), level);
        for (int k = 0; k < 4; ++k) out4[4 * i + k] = c(k);
    }
}

void pmref_get_unit(int n, const int* view, const float* coord4, float* out) {
    for (int i = 0; i < n; ++i) out[i] = g_pm->m_optim.getUnit(view[i], v4(coord4 + 4 * i));
}

// Optim::getPAxes (optim.cpp:67-84)
void pmref_get_paxes(int n, const int* view, const float* coord4, const float* normal4, float* px4, float* py4) {
    for (int i = 0; i < n; ++i) {
        Vector4f px, py;
        g_pm->m_optim.getPAxes(view[i], v4(coord4 + 4 * i), v4(normal4 + 4 * i), px, py);
        for (int k = 0; k < 4; ++k) { px4[4 * i + k] = px(k); py4[4 * i + k] = py(k); }
    }
}

// Image::getColor bilinear (image.cpp:448-471)
void pmref_get_color(int n, const int* view, const float* xy, int level, float* rgb) {
    for (int i = 0; i < n; ++i) {
        const Vector3f c = g_pm->m_photoSet.getColor(view[i], xy[2 * i], xy[2 * i + 1], level);
        rgb[3 * i] = c(0); rgb[3 * i + 1] = c(1); rgb[3 * i + 2] = c(2);
    }
}

// PatchManager::setGridsImages cell indices (patch_manager.cpp:223-239); ok[i]=0 when the view is dropped
void pmref_cells(int n, const int* view, const float* coord4, int* ixy, int* ok) {
    for (int i = 0; i < n; ++i) {
        Patch p;
        p.m_coord = v4(coord4 + 4 * i);
        std::vector<int> images(1, view[i]);
        g_pm->m_patchManager.setGridsImages(p, images);
        ok[i] = p.m_images.empty() ? 0 : 1;
        Patch q;
        q.m_coord = p.m_coord;
        q.m_images.assign(1, view[i]);
        g_pm->m_patchManager.setGrids(q);     // unclipped index (patch_manager.cpp:241-249)
        ixy[2 * i] = q.m_grids[0](0); ixy[2 * i + 1] = q.m_grids[0](1);
    }
}

// Raw (un-normalised) texture of one hypothesis in one view: Optim::getPAxes in `refview`, then
// Optim::getTex (optim.cpp:790-844).  flag = getTex's return; level = pyramid level it sampled,
// re-derived here with the reference's own project()/myPow2 because getTex keeps it local.
void pmref_get_tex(const float* coord4, const float* normal4, int refview, int view, float* tex, int* flag, int* level) {
    Optim& op = g_pm->m_optim;
    const Vector4f coord = v4(coord4), normal = v4(normal4);
    Vector4f px, py;
    op.getPAxes(refview, coord, normal, px, py);
    std::vector<Vector3f> t;
    *flag = op.getTex(coord, px, py, normal, view, g_pm->m_wsize, t);
    for (size_t i = 0; i < t.size(); ++i) { tex[3 * i] = t[i](0); tex[3 * i + 1] = t[i](1); tex[3 * i + 2] = t[i](2); }
    Vector3f center = g_pm->m_photoSet.project(view, coord, g_pm->m_level);
    Vector3f dx = g_pm->m_photoSet.project(view, coord + px, g_pm->m_level) - center;
    Vector3f dy = g_pm->m_photoSet.project(view, coord + py, g_pm->m_level) - center;
    const float ratio = (dx.norm() + dy.norm()) / 2.0f;
    int levelDiff = (int)floorf(log(ratio) / log(2.0f) + 0.5f);
    levelDiff = std::max(-g_pm->m_level, std::min(2, levelDiff));
    *level = g_pm->m_level + levelDiff;
}

// levelDiff rounding as compiled into this library (optim.cpp:808), for building the product's
// threshold table test: returns (int)floorf(log(ratio)/log(2.0f)+0.5f) unclamped.
int pmref_level_diff(float ratio) { return (int)floorf(log(ratio) / log(2.0f) + 0.5f); }

// The "hypothesis NCC eval" unit: PatchManager::computeNcc (patch_manager.cpp:401-404) =
// Optim::computeWeights (optim.cpp:942-948) + Optim::computeINCC(...,1) (optim.cpp:630-706).
// views: n rows of `stride` ints, nviews[i] valid per row, [0] = reference view.
void pmref_compute_ncc(int n, const float* coord4, const float* normal4, const int* views, const int* nviews,
                       int stride, float* incc, float* ncc) {
    Patch patch;
    for (int i = 0; i < n; ++i) {
        fill_patch(patch, coord4 + 4 * i, normal4 + 4 * i, views + (size_t)i * stride, nviews[i]);
        g_pm->m_optim.computeWeights(patch);
        const float s = g_pm->m_optim.computeINCC(patch.m_coord, patch.m_normal, patch.m_images, 1);
        if (incc) incc[i] = s;
        if (ncc) ncc[i] = 1.0f - g_pm->m_optim.unrobustincc(s);
    }
}

// Same loop, timed (seconds).  Output is accumulated so the work cannot be elided.
double pmref_time_compute_ncc(int n, const float* coord4, const float* normal4, const int* views, const int* nviews,
                              int stride, int repeats, double* checksum) {
    Patch patch;
    double acc = 0.0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < repeats; ++r) {
        for (int i = 0; i < n; ++i) {
            fill_patch(patch, coord4 + 4 * i, normal4 + 4 * i, views + (size_t)i * stride, nviews[i]);
            g_pm->m_optim.computeWeights(patch);
            acc += g_pm->m_optim.computeINCC(patch.m_coord, patch.m_normal, patch.m_images, 1);
        }
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (checksum) *checksum = acc;
    return std::chrono::duration<double>(t1 - t0).count();
}

// Optim::computeWeights (optim.cpp:942-948) -> m_weights[0..nviews)
void pmref_weights(const float* coord4, const float* normal4, const int* views, int nviews, float* w) {
    Patch patch;
    fill_patch(patch, coord4, normal4, views, nviews);
    g_pm->m_optim.computeWeights(patch);
    for (int i = 0; i < nviews; ++i) w[i] = g_pm->m_optim.m_weights[i];
}

// Optim::setINCCs 1-vs-all (optim.cpp:708-746)
void pmref_set_inccs(const float* coord4, const float* normal4, const int* views, int nviews, int robust, float* out) {
    Patch patch;
    fill_patch(patch, coord4, normal4, views, nviews);
    std::vector<float> inccs;
    std::vector<int> idx(views, views + nviews);
    g_pm->m_optim.setINCCs(patch, inccs, idx, robust);
    for (int i = 0; i < nviews; ++i) out[i] = inccs[i];
}

// Optim::setINCCs pairwise (optim.cpp:748-783); out is nviews x nviews row-major
void pmref_set_inccs_pair(const float* coord4, const float* normal4, const int* views, int nviews, int robust, float* out) {
    Patch patch;
    fill_patch(patch, coord4, normal4, views, nviews);
    std::vector<std::vector<float> > inccs;
    std::vector<int> idx(views, views + nviews);
    g_pm->m_optim.setINCCs(patch, inccs, idx, robust);
    for (int i = 0; i < nviews; ++i) for (int j = 0; j < nviews; ++j) out[i * nviews + j] = inccs[i][j];
}

// ---- patch record marshalling for the multi-step functions ----------------------------------------
// A "patch record" crosses this ABI as: coord4, normal4, scal[4] = {ncc, dscale, ascale, tmp},
// images[maxv] + nimages, grids[2*maxv], vimages[maxv] + nvimages, vgrids[2*maxv].
struct PatchIO {
    float* coord4; float* normal4; float* scal4;
    int* images; int* nimages; int* grids;
    int* vimages; int* nvimages; int* vgrids;
    int maxv;
};

static void patch_out(const Patch& p, const PatchIO& io, int i) {
    for (int k = 0; k < 4; ++k) { io.coord4[4 * i + k] = p.m_coord(k); io.normal4[4 * i + k] = p.m_normal(k); }
    io.scal4[4 * i] = p.m_ncc; io.scal4[4 * i + 1] = p.m_dscale; io.scal4[4 * i + 2] = p.m_ascale; io.scal4[4 * i + 3] = p.m_tmp;
    const int ni = std::min((int)p.m_images.size(), io.maxv);
    io.nimages[i] = (int)p.m_images.size();
    for (int k = 0; k < ni; ++k) {
        io.images[(size_t)i * io.maxv + k] = p.m_images[k];
        if (k < (int)p.m_grids.size()) {
            io.grids[((size_t)i * io.maxv + k) * 2] = p.m_grids[k](0);
            io.grids[((size_t)i * io.maxv + k) * 2 + 1] = p.m_grids[k](1);
        }
    }
    const int nv = std::min((int)p.m_vimages.size(), io.maxv);
    io.nvimages[i] = (int)p.m_vimages.size();
    for (int k = 0; k < nv; ++k) {
        io.vimages[(size_t)i * io.maxv + k] = p.m_vimages[k];
        if (k < (int)p.m_vgrids.size()) {
            io.vgrids[((size_t)i * io.maxv + k) * 2] = p.m_vgrids[k](0);
            io.vgrids[((size_t)i * io.maxv + k) * 2 + 1] = p.m_vgrids[k](1);
        }
    }
}

// Optim::preProcess (optim.cpp:137-163) on fresh patches {coord, normal, images}; ret[i] = its return.
// Output record = the patch after the call (m_images order, m_dscale/m_ascale from setScales).
void pmref_pre_process(int n, const float* coord4, const float* normal4, const int* views, const int* nviews, int stride,
                       int* ret, PatchIO* out) {
    for (int i = 0; i < n; ++i) {
        Patch patch;
        fill_patch(patch, coord4 + 4 * i, normal4 + 4 * i, views + (size_t)i * stride, nviews[i]);
        ret[i] = g_pm->m_optim.preProcess(patch);
        patch_out(patch, *out, i);
    }
}

// PMR1 refinement stream/seed (oracle/shim/nlopt.hpp)
void pmref_refine_seed(unsigned long long seed) { pmr1::state().seed = seed; }
void pmref_refine_stream(unsigned long long stream) { pmr1::state().stream = stream; }

// Optim::refinePatch (optim.cpp:470-547) on patches whose images/dscale are already set (post-preProcess).
// trace (optional): 97 * 4 doubles per patch = every evaluated {x0,x1,x2,f}.
void pmref_refine(int n, float* coord4, float* normal4, const float* dscale, const int* views, const int* nviews, int stride,
                  const unsigned long long* streams, float* ncc_out, double* trace) {
    std::vector<double> tr;
    for (int i = 0; i < n; ++i) {
        Patch patch;
        fill_patch(patch, coord4 + 4 * i, normal4 + 4 * i, views + (size_t)i * stride, nviews[i]);
        patch.m_dscale = dscale[i];
        pmr1::state().stream = streams[i];
        tr.clear();
        pmr1::state().trace = trace ? &tr : nullptr;
        g_pm->m_optim.refinePatch(patch, 100);
        pmr1::state().trace = nullptr;
        for (int k = 0; k < 4; ++k) { coord4[4 * i + k] = patch.m_coord(k); normal4[4 * i + k] = patch.m_normal(k); }
        ncc_out[i] = patch.m_ncc;
        if (trace) for (size_t k = 0; k < tr.size() && k < 97 * 4; ++k) trace[(size_t)i * 97 * 4 + k] = tr[k];
    }
}

// Optim::cost_func (optim.cpp:401-468) at given encoded points, for a patch context set up exactly as
// refinePatch does (:481-490).  x: n x 3 doubles.
void pmref_cost_func(const float* coord4, const float* normal4, float dscale, const int* views, int nviews,
                     int n, const double* x, double* cost) {
    Optim& op = g_pm->m_optim;
    Patch patch;
    fill_patch(patch, coord4, normal4, views, nviews);
    op.m_center = patch.m_coord;
    op.m_ray = patch.m_coord - g_pm->m_photoSet.m_photos[patch.m_images[0]].m_center;
    op.m_ray /= op.m_ray.norm();
    op.m_indexes = patch.m_images;
    op.m_dscale = dscale;
    op.m_ascale = M_PI / 48.0f;
    for (int i = 0; i < n; ++i) cost[i] = Optim::cost_func(3, x + 3 * i, nullptr, nullptr);
}

// Optim::encode / decode (optim.cpp:549-599) in the same context
void pmref_encode(const float* coord4, const float* normal4, float dscale, int refview, double* x3) {
    Optim& op = g_pm->m_optim;
    op.m_center = v4(coord4);
    op.m_ray = op.m_center - g_pm->m_photoSet.m_photos[refview].m_center;
    op.m_ray /= op.m_ray.norm();
    op.m_indexes.assign(1, refview);
    op.m_dscale = dscale;
    op.m_ascale = M_PI / 48.0f;
    op.encode(v4(coord4), v4(normal4), x3);
}

// Optim::postProcess (optim.cpp:260-298) on patches in their post-refine state.
void pmref_post_process(int n, const float* coord4, const float* normal4, const float* scal4, const int* views, const int* nviews,
                        int stride, int* ret, PatchIO* out) {
    for (int i = 0; i < n; ++i) {
        Patch patch;
        fill_patch(patch, coord4 + 4 * i, normal4 + 4 * i, views + (size_t)i * stride, nviews[i]);
        patch.m_ncc = scal4[4 * i]; patch.m_dscale = scal4[4 * i + 1]; patch.m_ascale = scal4[4 * i + 2];
        ret[i] = g_pm->m_optim.postProcess(patch);
        patch_out(patch, *out, i);
    }
}

// ---- masks: Image::getMask (image.cpp:749-781) over Image::alloc's thresholded level 0 (:143-161) and
// buildMaskPyramid (:717-747); Photo::getMask (photo.cpp:44-52); PhotoSet::getMask (photoSet.cpp:215-233) ----
// One pyramid level of a view's mask through Image::getMask(ix, iy, level); returns 0 when the view has no mask.
int pmref_get_mask_level(int view, int level, unsigned char* out) {
    const int w = g_pm->m_photoSet.getWidth(view, level), h = g_pm->m_photoSet.getHeight(view, level);
    if (g_pm->m_photoSet.getMask(view, 0, 0, level) == -1) return 0;
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) out[(size_t)y * w + x] = (unsigned char)g_pm->m_photoSet.getMask(view, x, y, level);
    return 1;
}

// PhotoSet::getMask(index, coord, level) per view (view >= 0) or PhotoSet::getMask(coord, level) over all views (view < 0)
void pmref_get_mask(int n, int view, const float* coord4, int level, int* out) {
    for (int i = 0; i < n; ++i)
        out[i] = view >= 0 ? g_pm->m_photoSet.getMask(view, v4(coord4 + 4 * i), level) : g_pm->m_photoSet.getMask(v4(coord4 + 4 * i), level);
}

// ---- patch store -------------------------------------------------------------------------------------
void pmref_clear_patches(void) { g_pm->m_patchManager.init(); }

// PatchManager::addPatch (patch_manager.cpp:158-189) after setGrids (:241-249); like readPatches (:450-462)
void pmref_add_patches(int n, const float* coord4, const float* normal4, const float* scal4, const int* views, const int* nviews, int stride) {
    for (int i = 0; i < n; ++i) {
        Ppatch pp(new Patch());
        fill_patch(*pp, coord4 + 4 * i, normal4 + 4 * i, views + (size_t)i * stride, nviews[i]);
        pp->m_ncc = scal4[4 * i]; pp->m_dscale = scal4[4 * i + 1]; pp->m_ascale = scal4[4 * i + 2];
        pp->m_tmp = pp->score2(g_pm->m_nccThreshold);
        g_pm->m_patchManager.setGrids(*pp);
        g_pm->m_patchManager.addPatch(pp);
    }
}

// DepthNormInit::createPatches (depth_normal_init.cpp:29-33) -> readPatches of <prefix>ply/00000000.patch
void pmref_create_patches(void) { g_pm->m_dnInit.createPatches(); }

// PatchManager::collectPatches(target) (patch_manager.cpp:75-104); returns count
int pmref_collect(int target) {
    g_pm->m_patchManager.collectPatches(target);
    return (int)g_pm->m_patchManager.m_ppatches.size();
}

void pmref_get_patches(PatchIO* out) {
    const std::vector<Ppatch>& pp = g_pm->m_patchManager.m_ppatches;
    for (size_t i = 0; i < pp.size(); ++i) patch_out(*pp[i], *out, (int)i);
}

// m_dpgrids of one view as m_ppatches ids (-1 = m_MAXDEPTH); requires a prior pmref_collect
void pmref_get_depth_map(int view, int* ids) {
    const std::vector<Ppatch>& g = g_pm->m_patchManager.m_dpgrids[view];
    for (size_t i = 0; i < g.size(); ++i) ids[i] = (g[i] == PatchManager::m_MAXDEPTH) ? -1 : g[i]->m_id;
}

// cell occupancy counts of m_pgrids / m_vpgrids for one view
void pmref_get_cell_counts(int view, int which, int* counts) {
    const std::vector<std::vector<Ppatch> >& g = which ? g_pm->m_patchManager.m_vpgrids[view] : g_pm->m_patchManager.m_pgrids[view];
    for (size_t i = 0; i < g.size(); ++i) counts[i] = (int)g[i].size();
}

// Propagate::run (propagate.cpp:28-64): the reference's own raster Gauss-Seidel sweep, unmodified
void pmref_propagate_run(int iter) { g_pm->m_propagate.run(iter); }
// Filter::run (filter.cpp:25-49)
void pmref_filter_run(void) { g_pm->m_filter.run(); }

// Filter::computeGain (filter.cpp:108-146) for every collected patch (call pmref_collect first)
void pmref_gains(float* gains) {
    const std::vector<Ppatch>& pp = g_pm->m_patchManager.m_ppatches;
    for (size_t i = 0; i < pp.size(); ++i) gains[i] = g_pm->m_filter.computeGain(*pp[i]);
}

// PmMvps::isNeighbor (pmmvps.cpp:117-147) between collected patches a[i], b[i]
void pmref_is_neighbor(int n, const int* a, const int* b, float thr, int* out) {
    const std::vector<Ppatch>& pp = g_pm->m_patchManager.m_ppatches;
    for (int i = 0; i < n; ++i) out[i] = g_pm->isNeighbor(*pp[a[i]], *pp[b[i]], thr);
}

// PmMvps::run (pmmvps.cpp:76-114), unmodified; returns wall seconds; *alive = patches after the last Filter::run
double pmref_run(int* alive) {
    const auto t0 = std::chrono::steady_clock::now();
    g_pm->run();
    const auto t1 = std::chrono::steady_clock::now();
    g_pm->m_patchManager.collectPatches(0);
    if (alive) *alive = (int)g_pm->m_patchManager.m_ppatches.size();
    return std::chrono::duration<double>(t1 - t0).count();
}

// ---- schedule PMS1 driven through the reference's own propagatePatch ---------------------------------------------------------
// One dest cell (x, y) of `image`: trim it like propagatePatch / propagatePmImage do (sortPatches, removePatch beyond
// MAX_NUM_OF_PATCHES), then replay every Propagate::propagatePatch call that the raster sweep would aim at it: the sorted
// top-MAX patches of (x, y - inc), then of (x - inc, y), whose reference image is `image` (propagate.cpp:87-108).  The PMR1
// stream of a call is (iter << 56) ^ (image << 40) ^ (cell << 8) ^ call, the same function the CUDA sweep uses.
int pmref_propagate_dest(int image, int x, int y, int inc, int iter) {
    PatchManager& pm = g_pm->m_patchManager;
    const int gw = pm.m_gwidths[image], gh = pm.m_gheights[image];
    const int maxp = g_pm->m_propagate.MAX_NUM_OF_PATCHES;
    const int index = y * gw + x;
    {
        std::vector<Ppatch> cur = pm.m_pgrids[image][index];          // copy: removePatch edits the cell
        pm.sortPatches(cur, 0);
        for (int i = (int)cur.size() - 1; i >= maxp; --i) pm.removePatch(cur[i]);
    }
    std::vector<Ppatch> sources;
    for (int side = 0; side < 2; ++side) {
        const int sx = side == 0 ? x : x - inc, sy = side == 0 ? y - inc : y;
        if (sx < 0 || gw <= sx || sy < 0 || gh <= sy) continue;
        std::vector<Ppatch> src = pm.m_pgrids[image][sy * gw + sx];
        pm.sortPatches(src, 0);
        if ((int)src.size() > maxp) src.resize(maxp);
        for (size_t n = 0; n < src.size(); ++n) if (src[n]->m_images[0] == image) sources.push_back(src[n]);
    }
    for (size_t call = 0; call < sources.size(); ++call) {
        pmr1::state().stream = ((unsigned long long)(unsigned)iter << 56) ^ ((unsigned long long)(unsigned)image << 40) ^
                               ((unsigned long long)(unsigned)index << 8) ^ (unsigned long long)call;
        g_pm->m_propagate.propagatePatch(sources[call], image, index);
    }
    return (int)sources.size();
}

// ---- teacher forcing: Propagate::propagatePatch replayed with every intermediate recorded ---------------------------------------
// pmref_trace_dest does for one dest cell what pmref_propagate_dest does, but instead of calling the reference's propagatePatch
// it walks that function's control flow (propagate.cpp:126-218) through the reference's OWN member functions -- sortPatches,
// removePatch, generatePatch, preProcess, refinePatch, postProcess, addPatch -- and records, per try: where the try ended, the
// candidate's m_ncc out of generatePatch, the patch as refinePatch left it (the "hypothesis" a teacher-forced device run starts
// from), postProcess' return value and the patch as postProcess left it, and what happened to the cell.
// tests/test_trace_cpu.py pins this walk to the real propagatePatch: both leave bit-identical stores from the same state.
//   code:     0 generatePatch returned NULL, 1 lost against the worst patch's m_ncc, 2 preProcess == -1, 3 refined
//   decision: 0 nothing stored, 1 added into free room, 2 replaced the worst patch
struct TraceIO {
    int cap;
    int* code; float* ncc0; int* post_ret; int* decision; int* branch_full;
    PatchIO mid;         // after refinePatch (m_images as preProcess left them, m_ncc / m_dscale / m_ascale in scal4)
    PatchIO fin;         // after postProcess (lists, grids, visible lists, m_tmp)
};

int pmref_trace_dest(int image, int x, int y, int inc, int iter, TraceIO* io) {
    PatchManager& pm = g_pm->m_patchManager;
    Propagate& pr = g_pm->m_propagate;
    const int gw = pm.m_gwidths[image], gh = pm.m_gheights[image];
    const int maxp = pr.MAX_NUM_OF_PATCHES;
    const int index = y * gw + x;
    {
        std::vector<Ppatch> cur = pm.m_pgrids[image][index];
        pm.sortPatches(cur, 0);
        for (int i = (int)cur.size() - 1; i >= maxp; --i) pm.removePatch(cur[i]);
    }
    std::vector<Ppatch> sources;
    for (int side = 0; side < 2; ++side) {
        const int sx = side == 0 ? x : x - inc, sy = side == 0 ? y - inc : y;
        if (sx < 0 || gw <= sx || sy < 0 || gh <= sy) continue;
        std::vector<Ppatch> src = pm.m_pgrids[image][sy * gw + sx];
        pm.sortPatches(src, 0);
        if ((int)src.size() > maxp) src.resize(maxp);
        for (size_t n = 0; n < src.size(); ++n) if (src[n]->m_images[0] == image) sources.push_back(src[n]);
    }
    int ntry = 0;
    for (size_t call = 0; call < sources.size(); ++call) {
        pmr1::state().stream = ((unsigned long long)(unsigned)iter << 56) ^ ((unsigned long long)(unsigned)image << 40) ^
                               ((unsigned long long)(unsigned)index << 8) ^ (unsigned long long)call;
        const Ppatch& ppatch = sources[call];
        // ---- propagate.cpp:126-136 ----
        std::vector<Ppatch>& ppatches = pm.m_pgrids[image][index];
        pm.sortPatches(ppatches, 0);
        int npatches = (int)ppatches.size();
        if (npatches > maxp) { for (int i = npatches - 1; i >= maxp; --i) pm.removePatch(ppatches[i]); }
        // ---- :138-150 ----
        std::default_random_engine generator;
        std::uniform_real_distribution<float> distribution(-0.5, 0.5);
        const int cx = index % gw, cy = index / gw;
        Vector3f icoord;
        icoord << (g_pm->m_csize * (2 * cx + 1) - 1) / 2.0f, (g_pm->m_csize * (2 * cy + 1) - 1) / 2.0f, 1.0f;
        for (int it = 0; it < pr.MAX_NUM_OF_PROPAG; ++it, ++ntry) {                          // :153
            const bool rec = ntry < io->cap;
            if (rec) { io->code[ntry] = 0; io->ncc0[ntry] = 0.0f; io->post_ret[ntry] = -2; io->decision[ntry] = 0; }
            npatches = (int)ppatches.size();
            if (rec) io->branch_full[ntry] = npatches < maxp ? 0 : 1;
            Ppatch newppatch;
            if (npatches < maxp) {                                                           // :156-165
                Vector3f d;
                d << distribution(generator) * g_pm->m_csize, distribution(generator) * g_pm->m_csize, 0.0f;
                const Vector3f nic = icoord + d;
                newppatch = pr.generatePatch(ppatch, nic);
                if (!newppatch) continue;
                if (rec) io->ncc0[ntry] = newppatch->m_ncc;
            } else {                                                                          // :166-173
                pm.sortPatches(ppatches, 0);
                const Vector3f ic = g_pm->m_photoSet.project(image, ppatches[maxp - 1]->m_coord, g_pm->m_level);
                newppatch = pr.generatePatch(ppatch, ic);
                if (newppatch && rec) io->ncc0[ntry] = newppatch->m_ncc;
                if (!newppatch) continue;
                if (newppatch->m_ncc < ppatches[maxp - 1]->m_ncc) { if (rec) io->code[ntry] = 1; continue; }
            }
            if (g_pm->m_optim.preProcess(*newppatch) == -1) { if (rec) io->code[ntry] = 2; continue; }   // :183-187
            g_pm->m_optim.refinePatch(*newppatch, 100);                                                  // :190
            if (rec) { io->code[ntry] = 3; patch_out(*newppatch, io->mid, ntry); }
            const int pr_ret = g_pm->m_optim.postProcess(*newppatch);                                     // :193
            if (rec) { io->post_ret[ntry] = pr_ret; patch_out(*newppatch, io->fin, ntry); }
            if (pr_ret == -1) continue;
            if (npatches == maxp) { pm.removePatch(ppatches[maxp - 1]); if (rec) io->decision[ntry] = 2; }   // :199-202
            else if (rec) io->decision[ntry] = 1;
            pm.addPatch(newppatch);                                                                       // :209
        }
    }
    return ntry;
}

// one wavefront step: every dest cell of anti-diagonal `diag` of `image`
int pmref_propagate_diag(int image, int diag, int inc, int iter) {
    PatchManager& pm = g_pm->m_patchManager;
    const int gw = pm.m_gwidths[image], gh = pm.m_gheights[image];
    int calls = 0;
    for (int x = std::max(0, diag - gh + 1); x <= std::min(gw - 1, diag); ++x) calls += pmref_propagate_dest(image, x, diag - x, inc, iter);
    return calls;
}

// ids (m_ppatches indexes after the last pmref_collect) of the m_pgrids / m_vpgrids entries of one cell
int pmref_cell_ids(int view, int index, int which, int* ids, int cap) {
    const std::vector<Ppatch>& g = which ? g_pm->m_patchManager.m_vpgrids[view][index] : g_pm->m_patchManager.m_pgrids[view][index];
    for (size_t i = 0; i < g.size() && (int)i < cap; ++i) ids[i] = g[i]->m_id;
    return (int)g.size();
}

// ---- Filter::run, stage by stage (the stage functions are protected in filter.hpp:36-52; see the access note above) --------------
static std::vector<Ppatch> g_prev;

// Filter::setDepthMapsVGridsVPGridsAddPatchV (filter.cpp:628-655)
void pmref_filter_rebuild(int additive) { g_pm->m_filter.setDepthMapsVGridsVPGridsAddPatchV(additive); }

// collectPatches(0) and remember m_ppatches so per-patch results can be reported in this order after a stage
int pmref_stage_begin(void) {
    g_pm->m_patchManager.collectPatches(0);
    g_prev = g_pm->m_patchManager.m_ppatches;
    return (int)g_prev.size();
}

// stage: 1 filterOutside (:51-106), 2 filterExact (:148-263), 3 filterNeighbor(1) (:265-336), 4 filterSmallGroups (:432-525)
void pmref_filter_stage(int stage) {
    switch (stage) {
        case 1: g_pm->m_filter.filterOutside(); break;
        case 2: g_pm->m_filter.filterExact(); break;
        case 3: g_pm->m_filter.filterNeighbor(1); break;
        case 4: g_pm->m_filter.filterSmallGroups(); break;
    }
}

// alive[i] = 1 when the i-th patch of the last pmref_stage_begin is still reachable from the grids
void pmref_stage_alive(int* alive) {
    g_pm->m_patchManager.collectPatches(0);
    std::map<const Patch*, int> now;
    const std::vector<Ppatch>& pp = g_pm->m_patchManager.m_ppatches;
    for (size_t i = 0; i < pp.size(); ++i) now[pp[i].get()] = 1;
    for (size_t i = 0; i < g_prev.size(); ++i) alive[i] = now.count(g_prev[i].get()) ? 1 : 0;
}

void pmref_stage_patches(PatchIO* out) { for (size_t i = 0; i < g_prev.size(); ++i) patch_out(*g_prev[i], *out, (int)i); }
void pmref_stage_gains(float* g) { for (size_t i = 0; i < g_pm->m_filter.m_gains.size(); ++i) g[i] = g_pm->m_filter.m_gains[i]; }
void pmref_stage_rejects(int* r) { for (size_t i = 0; i < g_pm->m_filter.m_rejects.size(); ++i) r[i] = g_pm->m_filter.m_rejects[i]; }

// PatchManager::findNeighbors(patch, 4, 2, 1) sizes and Filter::filterQuad decisions for the remembered patches
void pmref_stage_neighbors(int* count, int* quad) {
    for (size_t i = 0; i < g_prev.size(); ++i) {
        std::vector<Ppatch> nb;
        g_pm->m_patchManager.findNeighbors(*g_prev[i], nb, 4, 2, 1);
        count[i] = (int)nb.size();
        quad[i] = nb.size() >= 6 ? g_pm->m_filter.filterQuad(*g_prev[i], nb) : -1;
    }
}

// PatchManager::writePly (patch_manager.cpp:542-633) of the current store to `path` (ASCII PLY with the per-patch colours); returns the count
int pmref_write_ply(const char* path) {
    g_pm->m_patchManager.collectPatches(0);
    g_pm->m_patchManager.writePly(g_pm->m_patchManager.m_ppatches, std::string(path));
    return (int)g_pm->m_patchManager.m_ppatches.size();
}

// ---- PatchManager's public surface, for the pass-through tests ---------------------------------------------------------------
// isVisible0 (cells == NULL; the cells come back) / isVisible (patch_manager.cpp:327-376)
void pmref_is_visible(int n, const float* coord4, const float* normal4, const int* image, const int* cells, float strict, int* out, int* cells_out) {
    for (int i = 0; i < n; ++i) {
        Patch patch;
        patch.m_coord = v4(coord4 + 4 * i); patch.m_normal = v4(normal4 + 4 * i);
        int ix = 0, iy = 0;
        if (cells) { ix = cells[2 * i]; iy = cells[2 * i + 1]; out[i] = g_pm->m_patchManager.isVisible(patch, image[i], ix, iy, strict); }
        else out[i] = g_pm->m_patchManager.isVisible0(patch, image[i], ix, iy, strict);
        if (cells_out) { cells_out[2 * i] = ix; cells_out[2 * i + 1] = iy; }
    }
}
// setScales (patch_manager.cpp:378-399) on fresh patches
void pmref_set_scales(int n, const float* coord4, const int* views, const int* nviews, int stride, float* dscale, float* ascale) {
    for (int i = 0; i < n; ++i) {
        Patch patch;
        fill_patch(patch, coord4 + 4 * i, coord4 + 4 * i, views + (size_t)i * stride, nviews[i]);
        g_pm->m_patchManager.setScales(patch);
        dscale[i] = patch.m_dscale; ascale[i] = patch.m_ascale;
    }
}
// findNeighbors (patch_manager.cpp:671-728) of a free-standing patch: the neighbours' m_ppatches indices (after pmref_collect), ascending
int pmref_find_neighbors(const float* coord4, const float* normal4, const float* scal4, const int* views, int nviews, float scale, int margin, int* ids, int cap) {
    Patch patch;
    fill_patch(patch, coord4, normal4, views, nviews);
    patch.m_ncc = scal4[0]; patch.m_dscale = scal4[1]; patch.m_ascale = scal4[2];
    g_pm->m_patchManager.setGrids(patch);
    std::vector<Ppatch> nb;
    g_pm->m_patchManager.findNeighbors(patch, nb, scale, margin, 0);
    std::vector<int> out;
    for (size_t i = 0; i < nb.size(); ++i) out.push_back(nb[i]->m_id);
    std::sort(out.begin(), out.end());
    for (size_t i = 0; i < out.size() && (int)i < cap; ++i) ids[i] = out[i];
    return (int)out.size();
}
// removePatch / updateDepthMaps (patch_manager.cpp:303-325, 191-221) for patches of the last pmref_collect, by index
void pmref_remove_patches(int n, const int* ids) {
    std::vector<Ppatch> pp = g_pm->m_patchManager.m_ppatches;
    for (int i = 0; i < n; ++i) g_pm->m_patchManager.removePatch(pp[ids[i]]);
}
void pmref_update_depth_maps(int n, const int* ids) {
    std::vector<Ppatch> pp = g_pm->m_patchManager.m_ppatches;
    for (int i = 0; i < n; ++i) g_pm->m_patchManager.updateDepthMaps(pp[ids[i]]);
}

// PatchManager::setVImagesVGrids / Optim::check on a free-standing patch record (optim.cpp:290-323)
int pmref_check(const float* coord4, const float* normal4, const float* scal4, const int* views, int nviews, float* gain_out, PatchIO* out) {
    Patch patch;
    fill_patch(patch, coord4, normal4, views, nviews);
    patch.m_ncc = scal4[0]; patch.m_dscale = scal4[1]; patch.m_ascale = scal4[2];
    g_pm->m_patchManager.setGrids(patch);
    g_pm->m_patchManager.setVImagesVGrids(patch);
    const int r = g_pm->m_optim.check(patch);
    if (gain_out) *gain_out = patch.m_tmp;
    if (out) patch_out(patch, *out, 0);
    return r;
}

}  // extern "C"
