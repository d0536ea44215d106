/* oracle/pm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's (imkaywu/MVSKit) PatchMatch hot path, one function per
 * reference function, each citing the file:line it follows.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this; the product (mvskit_b200/) never does.
 *
 * Pinning: the reference has no tests, golden vectors or fixtures, so this
 * restatement is pinned against the reference ITSELF, compiled unmodified into
 * oracle/_ref/libpmref.so (oracle/Makefile): tests/test_oracle_vs_ref.py requires bit-identical
 * outputs, and tests/golden/ holds vectors generated from libpmref.so (tests/golden/make_golden.py).
 * Third-party arithmetic that is absent and unpinned in the reference (Eigen summation order,
 * NLopt BOBYQA) is defined by oracle/shim/ -- see the headers there; at the NLopt boundary parity
 * is UNPINNED by construction and the refinement schedule is the PMR1 definition.
 */
#ifndef PM_ORACLE_H
#define PM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmo_scene {
    int nviews;
    int level;          /* working pyramid level (Option::m_level, option.cpp:20) */
    int nlevels;        /* level + 3 (pmmvps.cpp:36) */
    int csize, wsize, tau, min_image_num;
    int depth;          /* PmMvps::m_depth */
    float ncc_threshold, ncc_threshold_before;
    float angle_threshold0, angle_threshold1, max_angle_threshold, quad_threshold;
    float neighbor_threshold, neighbor_threshold1, neighbor_threshold2;
    /* per view */
    float* P;           /* nviews * nlevels * 12 */
    float* center;      /* nviews * 4 */
    float* oaxis;       /* nviews * 4 */
    float* xaxis;       /* nviews * 3  (Optim::m_xaxes) */
    float* yaxis;
    float* zaxis;
    float* ipscale;     /* nviews      (Optim::m_ipscales) */
    float* Minv;        /* nviews * nlevels * 9: inverse of P[:, :3] per level (Camera::unproject) */
    unsigned char** img;/* nviews * nlevels pointers to interleaved u8 RGB */
    int* w;             /* nviews * nlevels */
    int* h;
    int* gw;            /* nviews: cell grid (patch_manager.cpp:36-37) */
    int* gh;
    unsigned char** mask;/* nviews * nlevels pointers to the u8 mask pyramid (0 / 255), NULL where a view has no mask */
} pmo_scene;

pmo_scene* pmo_scene_create(int nviews, int level, int csize, int wsize, int min_image_num, float ncc_threshold);
void pmo_scene_destroy(pmo_scene* s);
/* camera.cpp:65-100,295-308 + optim.cpp:43-65 from a level-0 3x4 P */
void pmo_set_camera(pmo_scene* s, int view, const float* P12);
/* image.cpp:92-192 (dims) + :245-315 (pyramid) from level-0 u8 RGB (copied) */
void pmo_set_image(pmo_scene* s, int view, const unsigned char* rgb, int w, int h);
void pmo_update_threshold(pmo_scene* s);              /* pmmvps.cpp:70-74 (+ ++m_depth, :106) */

void  pmo_project(const pmo_scene* s, int view, const float* X4, int level, float* out3);     /* camera.cpp:310-326 */
void  pmo_unproject(const pmo_scene* s, int view, const float* ic3, int level, float* out4);  /* camera.cpp:329-337 */
float pmo_get_unit(const pmo_scene* s, int view, const float* X4);                            /* optim.cpp:34-41 */
void  pmo_get_paxes(const pmo_scene* s, int view, const float* X4, const float* N4, float* px4, float* py4); /* optim.cpp:67-84 */
void  pmo_get_color(const pmo_scene* s, int view, float x, float y, int level, float* rgb);   /* image.cpp:448-471 */
int   pmo_level_diff(float ratio);                                                            /* optim.cpp:808 */
/* optim.cpp:790-844; returns 0 / -1; tex = wsize*wsize*3 floats; *level_out = pyramid level sampled (or -1) */
int   pmo_get_tex(const pmo_scene* s, const float* X4, const float* px4, const float* py4, const float* N4,
                  int view, float* tex, int* level_out);
void  pmo_normalize(float* tex, int sz);                                                      /* optim.cpp:917-940 */
float pmo_dot(const float* t0, const float* t1, int sz);                                      /* optim.cpp:601-609 */
float pmo_robustincc(float x);                                                                /* optim.cpp:622-624 */
float pmo_unrobustincc(float x);                                                              /* optim.cpp:626-628 */
void  pmo_compute_weights(const pmo_scene* s, const float* X4, const float* N4, const int* views, int nviews, float* w); /* optim.cpp:109-132,942-948 */
/* optim.cpp:630-706 given weights; levels_out (tau ints, may be NULL) */
float pmo_compute_incc(const pmo_scene* s, const float* X4, const float* N4, const int* views, int nviews,
                       const float* weights, int robust, int* levels_out);
/* patch_manager.cpp:401-404 batch: weights from the hypothesis itself, then computeINCC(...,1) */
void  pmo_compute_ncc(const pmo_scene* s, int n, const float* X4, const float* N4, const int* views, const int* nviews,
                      int stride, float* incc, float* ncc, int* levels);
/* optim.cpp:708-746 (1-vs-all) and :748-783 (pairwise, out nviews*nviews) */
void  pmo_set_inccs(const pmo_scene* s, const float* X4, const float* N4, const int* views, int nviews, int robust, float* out);
void  pmo_set_inccs_pair(const pmo_scene* s, const float* X4, const float* N4, const int* views, int nviews, int robust, float* out);
/* patch_manager.cpp:223-249: cell index of X in a view; returns 1 if inside the grid */
int   pmo_cell(const pmo_scene* s, int view, const float* X4, int* ix, int* iy);

/* image.cpp:143-161 (grey > 127 -> 255, else 0) + buildMaskPyramid :717-747 (a coarse pixel is inside when any of its
 * 2x2 fine pixels is) from a level-0 u8 mask (copied); call after pmo_set_image, same dimensions */
void  pmo_set_mask(pmo_scene* s, int view, const unsigned char* grey, int w, int h);
/* Photo::getMask(coord, level) photo.cpp:44-52 -> Image::getMask(float, float, level) image.cpp:749-781:
 * -1 without a mask or outside the image, else 0 / 255 at the rounded pixel */
int   pmo_get_mask_view(const pmo_scene* s, int view, const float* X4, int level);
/* PhotoSet::getMask(coord, level) photoSet.cpp:223-233: 0 as soon as one view's mask says outside, else -1 */
int   pmo_get_mask(const pmo_scene* s, const float* X4, int level);

#ifdef __cplusplus
}
#endif
#endif
