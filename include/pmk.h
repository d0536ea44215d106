/* include/pmk.h -- C ABI of the B200-native PatchMatch-MVS hot path ("pmk").
 *
 * Drop-in boundary for imkaywu/MVSKit's PM-MVS inner loop.  The reference has no FFI layer; its
 * boundary is the C++ class surface (PmMvps / Option / PatchManager / Patch).  The host-side mirror
 * of those classes lives in mvskit_b200/host/ and calls ONLY the functions below, each of which
 * replaces the reference member functions cited next to it (paths relative to the reference tree).
 *
 * Conventions
 *   - Every function returns 0 on success or a negative pmk_status; pmk_last_error() describes the
 *     last failure on the calling thread.  No exceptions cross this boundary.
 *   - One pmk_ctx per GPU.  Calls on a context are serialised by the caller and are stream-ordered
 *     on the context's CUDA stream; functions taking HOST pointers return after their results are
 *     in the caller's buffers; `_dev` variants take DEVICE pointers and only enqueue work.
 *   - The caller owns all host buffers; the context owns all device memory it allocates.
 *   - No CPU fallback exists: without a CUDA device pmk_create fails (PMK_ERR_CUDA).
 *   - Layouts: coord/normal are float[4] per item (x,y,z,w) like Eigen::Vector4f in Patch
 *     (pmmvps/patch.hpp:33-35); view lists are rows of `stride` ints, [0] = reference image
 *     (Patch::m_images, patch.hpp:38), entries past nviews[i] ignored.
 */
#ifndef PMK_H
#define PMK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMK_ABI_VERSION 3
#define PMK_MAX_LEVELS 6
#define PMK_MAX_TAU 8

typedef enum pmk_status {
    PMK_OK = 0,
    PMK_ERR_ARG = -1,      /* bad argument (the reference would exit(1) or index out of range)   */
    PMK_ERR_CUDA = -2,     /* CUDA runtime failure, or no device                                   */
    PMK_ERR_STATE = -3,    /* call order violated (e.g. evaluating before every view is uploaded)  */
    PMK_ERR_CAPACITY = -4  /* a fixed-capacity device structure overflowed                         */
} pmk_status;

typedef struct pmk_ctx pmk_ctx;

/* Scalars of Option (pmmvps/option.hpp:20-73, defaults option.cpp:19-33) that the path consumes;
 * PmMvps::init derives the remaining thresholds from them exactly as pmmvps.cpp:32,54-67. */
typedef struct pmk_config {
    int device;                 /* CUDA device ordinal                                             */
    int nviews;                 /* Option::m_nimages                                               */
    int level;                  /* Option::m_level   (working pyramid level; level+3 are built)    */
    int csize;                  /* Option::m_csize                                                 */
    int wsize;                  /* Option::m_wsize   (5, 7, 9 or 11)                               */
    int min_image_num;          /* Option::m_minImageNum                                           */
    float ncc_threshold;        /* Option::m_nccThreshold                                          */
    float max_angle_threshold;  /* Option::m_maxAngleThreshold, radians                            */
    float quad_threshold;       /* Option::m_quadThreshold                                         */
    int max_patches;            /* capacity of the device patch store; 0 = 4 per cell of all views */
    int cell_capacity;          /* slots per grid cell (m_pgrids + m_vpgrids entries); 0 = max(96, 4 * nviews) capped at 1024 */
    int jitter_mode;            /* 0: the reference's pixel jitter (propagate.cpp:139-141 re-seeds its engine on every
                                   call, so it is the same four draws each time); 1: Philox4x32 per (iter, view, cell, call, try) */
    int sweep_group;            /* views whose wavefronts advance together in Propagate::run: 1 = one view after the other like
                                   the reference (propagate.cpp:73); g > 1 = g views per pass (more parallel work per step) */
} pmk_config;

/* Thresholds held by PmMvps (pmmvps/pmmvps.hpp:36-87); read back for parity checks. */
typedef struct pmk_thresholds {
    int tau, depth;
    float ncc_threshold, ncc_threshold_before;
    float angle_threshold0, angle_threshold1, max_angle_threshold, quad_threshold;
    float neighbor_threshold, neighbor_threshold1, neighbor_threshold2;
} pmk_thresholds;

/* Per-view constants the device uses, as derived by the host side of the library. */
typedef struct pmk_camera {
    float P[12];        /* level-`level` projection, row-major 3x4 (image/camera.cpp:91-100)        */
    float center[4];    /* Camera::getCameraCenter                       (camera.cpp:295-308)       */
    float oaxis[4];     /* Camera::m_oaxis                               (camera.cpp:68-69)         */
    float xaxis[3], yaxis[3], zaxis[3];   /* Optim::m_xaxes/m_yaxes/m_zaxes (optim.cpp:43-54)        */
    float ipscale;      /* Optim::m_ipscales                             (optim.cpp:56-64)          */
} pmk_camera;

void pmk_default_config(pmk_config* cfg);                   /* Option::Option   (option.cpp:19-33)  */
int pmk_create(const pmk_config* cfg, pmk_ctx** out);       /* PmMvps::init     (pmmvps.cpp:18-68)  */
void pmk_destroy(pmk_ctx* ctx);
const char* pmk_last_error(void);
int pmk_abi_version(void);

/* Photo::init + Image::alloc + buildImagePyramid for one view (image/photoSet.cpp:20-61,
 * image/camera.cpp:27-100, image/image.cpp:92-192,245-315).  `P` is the level-0 3x4 projection
 * ("CONTOUR" camera file), `rgb` interleaved u8 of width*height*3.  Builds the level+3 level
 * pyramid on the device (kernel K0) as 4 x fp16 RGBX texels holding the u8-rounded values (exact). */
int pmk_set_view(pmk_ctx* ctx, int view, const float* P, const uint8_t* rgb, int width, int height);
/* The same from a JPEG byte stream: Image::readJpeg (image/image.cpp:827-879; CImg + ImageMagick in the reference, third party and
 * absent) becomes an nvJPEG decode straight into the staging buffer K0 reads.  JPEG decoders differ in the last bit (IDCT, chroma
 * upsampling): against libjpeg the mean pixel difference is below 0.6 grey levels on 4:4:4 streams.  width_out / height_out may be NULL. */
int pmk_set_view_jpeg(pmk_ctx* ctx, int view, const float* P, const uint8_t* jpeg, uint64_t nbytes, int* width_out, int* height_out);

/* The silhouette mask of one view: Image::alloc's mask branch (image/image.cpp:143-161: readPGMImage / readPBMImage, grey > 127 ->
 * 255 else 0) + Image::buildMaskPyramid (:717-747) on the device (kernel K0m).  `grey` is width*height u8, same dimensions as the
 * view's image; call after pmk_set_view(view).  Views without a mask answer -1 to getMask, like the reference without a mask file.
 * From then on Optim::postProcess rejects candidates for which PhotoSet::getMask(coord, m_level) == 0 (optim.cpp:265,
 * photoSet.cpp:223-233), in pmk_post_process and inside pmk_propagate alike. */
int pmk_set_view_mask(pmk_ctx* ctx, int view, const uint8_t* grey, int width, int height);
/* Image::m_masks[level] of one view (*has_mask = 0 and mask_out untouched when the view has none; mask_out may be NULL) */
int pmk_get_level_mask(pmk_ctx* ctx, int view, int level, uint8_t* mask_out, int* has_mask);
/* PhotoSet::getMask(coord, m_level) over all views (view < 0: 0 or -1; photoSet.cpp:223-233) or PhotoSet::getMask(view, coord, m_level)
 * = Photo::getMask (photo.cpp:44-52; -1, 0 or 255) on n points */
int pmk_probe_mask(pmk_ctx* ctx, int n, int view, const float* coord4, int* out);

int pmk_get_thresholds(pmk_ctx* ctx, pmk_thresholds* out);
int pmk_set_depth(pmk_ctx* ctx, int depth);                 /* PmMvps::m_depth (pmmvps.hpp:61)      */
int pmk_set_ncc_thresholds(pmk_ctx* ctx, float ncc_threshold, float ncc_threshold_before); /* PmMvps::m_nccThreshold / m_nccThresholdBefore (public, pmmvps.hpp:36,81) */
int pmk_update_threshold(pmk_ctx* ctx);                     /* PmMvps::updateThreshold + ++m_depth (pmmvps.cpp:70-74,106) */
int pmk_get_camera(pmk_ctx* ctx, int view, int level, pmk_camera* out);
int pmk_get_level_dims(pmk_ctx* ctx, int view, int level, int* width, int* height); /* Image::getWidth/getHeight */
int pmk_get_grid_dims(pmk_ctx* ctx, int view, int* gwidth, int* gheight);  /* PatchManager::m_gwidths/m_gheights (patch_manager.cpp:36-37) */
int pmk_get_level_image(pmk_ctx* ctx, int view, int level, uint8_t* rgb_out);       /* Image::m_images[level] */

/* K1 -- the "hypothesis NCC eval" unit: PatchManager::computeNcc (patch_manager.cpp:401-404) =
 * Optim::computeWeights (optim.cpp:942-948) + Optim::computeINCC(coord, normal, images, 1)
 * (optim.cpp:630-706) with getPAxes / getTex / getTexSafe / normalize / dot / robustincc inside.
 *   incc_out[i]   : computeINCC's return (2.0f = invalid sentinel)
 *   ncc_out[i]    : 1.0f - unrobustincc(incc)                  (may be NULL)
 *   levels_out    : n x tau ints, pyramid level each view was sampled at, -1 where getTex
 *                   returned -1 or the slot is unused              (may be NULL)
 * Host-pointer variant: copies inputs H2D, runs, copies results D2H, returns when they landed. */
int pmk_ncc_eval(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const int* views,
                 const int* nviews, int stride, float* incc_out, float* ncc_out, int* levels_out);
/* Byte-lean form of the host-pointer call for PCIe-bound callers (31 B instead of 60 B per hypothesis at stride 6): coord3 / normal3 are
 * rows of 3 floats (w = 1 and w = 0 implied -- what Patch::m_coord and a refined Patch::m_normal hold, patch.hpp:33-35), views8 rows of
 * `stride` bytes (view ids; the context must have nviews <= 255, so that 255 is an id no view has), nviews8 one byte each.  Results are bit-identical to pmk_ncc_eval on the widened inputs. */
int pmk_ncc_eval_packed(pmk_ctx* ctx, int n, const float* coord3, const float* normal3, const uint8_t* views8, const uint8_t* nviews8, int stride,
                        float* incc_out, float* ncc_out, int* levels_out);
/* Device-pointer variant: enqueues on the context stream and returns. */
int pmk_ncc_eval_dev(pmk_ctx* ctx, int n, const void* d_coord4, const void* d_normal4, const void* d_views,
                     const void* d_nviews, int stride, void* d_incc_out, void* d_ncc_out, void* d_levels_out);

/* K2 -- Optim::setINCCs (optim.cpp:708-783) for n patches {coord, normal, images}: robust = isRobust.
 *   pairwise == 0: 1-vs-all, out[n][stride]          (out[i][0] = 0, 2.0f where a texture is missing)
 *   pairwise == 1: all pairs, out[n][stride][stride] (symmetric, zero diagonal)
 * Unlike computeINCC this looks at EVERY listed image, not just the first tau (stride <= 128). */
int pmk_set_inccs(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const int* views, const int* nviews,
                  int stride, int robust, int pairwise, float* out);

/* Optim::preProcess (optim.cpp:137-163) on fresh candidates {coord, normal, images}: addImages, constraintImages at
 * m_nccThresholdBefore, sortImages, PatchManager::setScales (patch_manager.cpp:378-399), PhotoSet::checkAngles
 * (photoSet.cpp:77-103).  ret[i] = the reference's return value (0 / -1); images_out[i][maxv] (-1 padded) and
 * nimages_out[i] = m_images afterwards (cleared, like the reference, when checkAngles fails). */
int pmk_pre_process(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const int* views, const int* nviews,
                    int stride, int maxv, int* ret, int* images_out, int* nimages_out, float* dscale_out, float* ascale_out);

/* Optim::cost_func (optim.cpp:401-468): objective of the refinement, at nitems encoded points x3[i] (depth along the
 * reference ray / dscale, two normal angles / (pi/48)); item i belongs to patch patch_of_item[i], whose context
 * (m_center, m_ray, m_indexes, m_dscale) is set up as refinePatch does (optim.cpp:481-490). */
int pmk_cost_func(pmk_ctx* ctx, int npatches, const float* coord4, const float* normal4, const float* dscale, const int* views,
                  const int* nviews, int stride, int nitems, const int* patch_of_item, const double* x3, double* cost_out);

/* K3 -- Optim::refinePatch (optim.cpp:470-547).  The reference drives NLopt BOBYQA (third party, absent, unpinned);
 * this path runs the seeded counter-based schedule PMR1 over the same 3 variables, bounds and objective:
 * 1 + 12 levels x 8 candidates = 97 cost_func evaluations, Philox4x32-10 keyed by `seed`, counter (streams[i], level,
 * candidate).  coord4 / normal4 are updated in place; ncc_out = 1 - unrobustincc(computeINCC) with the pre-refinement
 * weights (optim.cpp:539).  trace_out (optional): n x 97 x {x0, x1, x2, cost} doubles, every evaluated point. */
int pmk_refine(pmk_ctx* ctx, int n, float* coord4, float* normal4, const float* dscale, const int* views, const int* nviews,
               int stride, const uint64_t* streams, uint64_t seed, float* ncc_out, double* trace_out);

/* Optim::postProcess (optim.cpp:260-290) up to and including m_tmp = score2: addImages, constraintImages at
 * m_nccThreshold, filterImagesByAngle, setRefImage (pairwise INCC), constraintImages, PatchManager::setGrids.
 * The tail that reads the patch store (setVImagesVGrids, check; optim.cpp:291-296) runs in pmk_propagate. */
int pmk_post_process(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const float* ncc, const int* views,
                     const int* nviews, int stride, int maxv, int* ret, int* images_out, int* nimages_out, int* grids_out,
                     float* tmp_out);

/* ---- device patch store: PatchManager's m_pgrids / m_vpgrids / m_dpgrids and m_ppatches as structure-of-arrays in HBM ----
 * A patch record crosses this boundary as coord4, normal4, scal4 = {m_ncc, m_dscale, m_ascale, m_tmp}, images[maxv] + nimages,
 * grids[maxv][2] = Patch::m_grids, and the same for m_vimages / m_vgrids (patch.hpp:33-66). */
int pmk_store_clear(pmk_ctx* ctx);                           /* PatchManager::init (patch_manager.cpp:24-52) */
/* PatchManager::readPatches body (patch_manager.cpp:450-462): m_tmp = score2(nccThreshold), m_vimages cleared, setGrids
 * (:241-249), addPatch (:158-189).  Image entries are view indexes (image2index already applied). */
int pmk_store_add(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const float* scal4, const int* images,
                  const int* nimages, int stride);
/* PatchManager::collectPatches (patch_manager.cpp:75-104): number of patches reachable from the grids */
int pmk_store_count(pmk_ctx* ctx, int* n_out);
/* m_ppatches after collectPatches, in the reference's collect order; any output pointer may be NULL */
int pmk_store_get(pmk_ctx* ctx, int nmax, int maxv, float* coord4, float* normal4, float* scal4, int* images, int* nimages,
                  int* grids, int* vimages, int* nvimages, int* vgrids, int* n_out);
int pmk_store_depth_map(pmk_ctx* ctx, int view, int* ids);   /* m_dpgrids[view] as patch ids, -1 = m_MAXDEPTH */
int pmk_store_cell_counts(pmk_ctx* ctx, int view, int which, int* counts);  /* sizes of m_pgrids (0) / m_vpgrids (1) cells */
int pmk_store_colors(pmk_ctx* ctx, int nmax, uint8_t* rgb);  /* PatchManager::writePly colours (patch_manager.cpp:566-581) */

/* K4 -- Propagate::run(iter) (propagate.cpp:28-64): propagatePmImage / propagatePatch / generatePatch (:72-237) for every view in
 * order, each as an anti-diagonal wavefront over dest cells (schedule PMS1), with preProcess, refinePatch (PMR1, `seed`)
 * and postProcess including its store-reading tail (setVImagesVGrids, check; optim.cpp:290-296).  stats16 (optional): calls, tries,
 * generatePatch == NULL, lost to the worst patch's ncc, preProcess failures (m_fcount0), postProcess failures (m_fcount1), added,
 * replaced, trimmed, hypothesis NCC evaluations. */
int pmk_propagate(pmk_ctx* ctx, int iter, uint64_t seed, uint64_t* stats16);
/* the same for wavefront steps [diag_first, diag_first + diag_count) of one view (parity tests step the sweep) */
int pmk_propagate_diagonals(pmk_ctx* ctx, int iter, int image, int diag_first, int diag_count, uint64_t seed, uint64_t* stats16);

/* Teacher-forced replay of ONE dest cell (x, y) of `image` (parity tests; BASELINE north_star: "accept/reject decisions given identical
 * hypotheses").  The sweep kernel runs exactly as in pmk_propagate_diagonals for that cell, except that try t does not generate and
 * refine its own hypothesis but starts from record t -- the patch as the reference's Optim::refinePatch left it (pmmvps/propagate.cpp:190)
 * -- and then decides everything that follows on the device: the m_ncc test against the cell's worst patch (:170), Optim::postProcess
 * (optim.cpp:260-298) with setVImagesVGrids and check, removePatch(worst) / addPatch(new) (:199-209).
 *   in : code[t] 0 generatePatch NULL / 1 lost to the worst patch / 2 preProcess == -1 / 3 refined; ncc0[t] = m_ncc out of
 *        generatePatch; for code 3: coord4, normal4, scal4 = {m_ncc, m_dscale, m_ascale, -}, images[t][stride] + nimages[t]
 *   out: ntries_out = tries the device's own sources give (must equal ntries), outcome[t] 0 NULL / 1 lost / 2 preProcess failed /
 *        3 postProcess or check failed / 4 stored / 5 diverged (the record says "lost" but the device's worst patch does not beat
 *        it); branch_full[t]; post_ret[t] = postProcess' store-independent return (0 / -1, -2 = not reached); for post_ret 0 the
 *        lists as stored: images_out / grids_out (ix, iy pairs) / vimages_out / vgrids_out [t][stride], counts, tmp_out = m_tmp. */
typedef struct pmk_forced_io {
    int ntries, stride;
    const int* code; const float* ncc0; const float* coord4; const float* normal4; const float* scal4; const int* nimages; const int* images;
    int* ntries_out; int* outcome; int* branch_full; int* post_ret;
    int* nimages_out; int* images_out; int* grids_out; int* nvimages_out; int* vimages_out; int* vgrids_out; float* tmp_out;
} pmk_forced_io;
int pmk_propagate_forced(pmk_ctx* ctx, int iter, int image, int x, int y, const pmk_forced_io* io, uint64_t* stats16);

/* K5 -- PatchManager::collectPatches + Filter::setDepthMapsVGridsVPGridsAddPatchV(additive) (filter.cpp:628-655): patch ids become
 * the reference's m_ppatches indices; depth maps, m_vimages / m_vgrids and m_vpgrids are rebuilt. */
int pmk_filter_rebuild(pmk_ctx* ctx, int additive, int* n_out);
/* One filter of Filter::run on a rebuilt store; per-patch results in collect order (any may be NULL):
 *   stage 1 filterOutside (filter.cpp:51-106)       f_out = computeGain (:108-146)
 *   stage 2 filterExact (:148-263)                  i_out = removed flag, i_out2 = new m_images.size()
 *   stage 3 filterNeighbor(1) (:265-336)            i_out = m_rejects, i_out2 = findNeighbors count, f_out = filterQuad residual (-1: not run)
 *   stage 4 filterSmallGroups (:432-525)            i_out = removed flag
 * killed_out = patches removed.  The grids are rebuilt by the next pmk_filter_rebuild, as in Filter::run. */
int pmk_filter_stage(pmk_ctx* ctx, int stage, int nmax, float* f_out, int* i_out, int* i_out2, int* killed_out);
/* K5..K9 -- Filter::run (filter.cpp:25-49).  counts6 (optional): patches before, removed by each of the four filters, patches after. */
int pmk_filter(pmk_ctx* ctx, int* counts6);

/* ---- multi-GPU: one process and one pmk_ctx per GPU; images and the patch store are replicated, the dest cells of every
 * wavefront step are dealt out to the ranks in turn (the step's cells are numbered view by view along the anti-diagonals; rank r of n
 * takes the cells whose number G has G % n == r, so neighbouring cells -- and with them the heavy regions -- spread evenly), and after
 * each step the ranks exchange the step's new and removed patches (ncclAllGather over NVLink/NVSwitch, headers first, then exactly the
 * payload the step produced) and all apply all of them in global cell order, so every replica holds the single-GPU store.  The
 * reference is single-threaded: there is no counterpart to cite. */
int pmk_step_share(int step_tasks, int rank, int nranks, int* count);        /* how many of a step's cells rank r takes (host only) */
int pmk_comm_unique_id(char* id128);                        /* rank 0: ncclGetUniqueId; ship the 128 bytes to every rank */
int pmk_comm_init(pmk_ctx* ctx, int rank, int nranks, const char* id128);     /* ncclCommInitRank; nranks == 1 needs no id */
int pmk_comm_destroy(pmk_ctx* ctx);
int pmk_store_checksum(pmk_ctx* ctx, uint64_t* out2);       /* {order-independent digest of the live patches, their count} */

/* Probes of the device-side building blocks, for parity tests (each item independent):
 *   project  : Camera::project at the working level        (camera.cpp:310-326)   -> out3
 *   unit     : Optim::getUnit                              (optim.cpp:34-41)      -> out1
 *   paxes    : Optim::getPAxes                             (optim.cpp:67-84)      -> px4, py4
 *   cell     : PatchManager::setGrids index + in-grid flag (patch_manager.cpp:223-249) -> ixy2, ok */
int pmk_probe(pmk_ctx* ctx, int n, const int* view, const float* coord4, const float* normal4,
              float* project3, float* unit1, float* px4, float* py4, int* cell_ixy2, int* cell_ok);
/* Camera::unproject (image/camera.cpp:329-337) at the working level, as Propagate::generatePatch uses it (propagate.cpp:224-226):
 * icoord3 = depth * (u, v, 1); coord4_out = (M^-1 (icoord - p4), 1). */
int pmk_probe_unproject(pmk_ctx* ctx, int n, const int* view, const float* icoord3, float* coord4_out);

/* PmMvps::isNeighbor (pmmvps.cpp:117-147; hunit NULL: computed from the two reference views as in :117-121) and isNeighborRadius
 * (:149-180; radius non-NULL) on n free-standing pairs.  A patch is 10 floats: coord4, normal4, m_dscale, (float)m_images[0]. */
int pmk_probe_neighbor(pmk_ctx* ctx, int n, const float* lhs10, const float* rhs10, const float* hunit, const float* radius, float threshold, int* out);

/* The store-reading tail of Optim::postProcess on n free-standing candidates against the current store (optim.cpp:285-323):
 * PatchManager::setGrids, setVImagesVGrids (patch_manager.cpp:267-301) and Optim::check = Filter::computeGain (filter.cpp:108-146), and,
 * when the gain is not negative, PatchManager::findNeighbors(patch, 4, 2) + Filter::filterQuad for more than 6 neighbours.
 * ret = check's return (1 = reject; -2 = a listed view does not see the candidate inside its grid), gain = m_tmp, nneighbors = findNeighbors' size (-1 when the gain already rejected),
 * vimages_out[n][stride] (-1 padded) / nvimages_out = the visible lists.  stride >= nviews. */
int pmk_probe_check(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const float* scal4, const int* images, const int* nimages, int stride,
                    int* ret, float* gain, int* nneighbors, int* vimages_out, int* nvimages_out);

/* ---- the rest of PatchManager's / Camera's public surface, for the host mirror's pass-throughs (mvskit_b200/host) ----------------
 * Camera::setProjection for "CONTOUR2" camera files: 6 intrinsics {fx, fy, skew, cx, cy, -} and 6 extrinsics {Euler angles in degrees,
 * translation} -> the level-0 3 x 4 projection K [R | t] (image/camera.cpp:116-131, quat2proj :241-261).  Host arithmetic only. */
int pmk_contour2_to_projection(const float* intrinsics6, const float* extrinsics6, float* P12);
/* PatchManager::isVisible0 (cell_ixy == NULL; the cells come back in cell_ixy_out) / isVisible (patch_manager.cpp:327-376) of n
 * free-standing points {coord4, normal4} in view image[i], against the store's depth maps, with the given strictness. */
int pmk_probe_visible(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const int* image, const int* cell_ixy, float strict,
                      int* out, int* cell_ixy_out);
/* PatchManager::setScales (patch_manager.cpp:378-399) for fresh patches: m_dscale, m_ascale from {coord4, images[n][stride], nimages}. */
int pmk_probe_scales(pmk_ctx* ctx, int n, const float* coord4, const int* images, const int* nimages, int stride, float* dscale, float* ascale);
/* PatchManager::findNeighbors(patch, neighbors, scale, margin) (patch_manager.cpp:671-728) for free-standing patches (scal4 = {ncc,
 * dscale, ascale, -}): the neighbours' store ids in ascending order (ids_out[n][cap], first min(count, cap)) and their number. */
int pmk_probe_neighbors(pmk_ctx* ctx, int n, const float* coord4, const float* normal4, const float* scal4, const int* images, const int* nimages,
                        int stride, float scale, int margin, int cap, int* ids_out, int* count_out);
/* Store ids of the live patches in collect order: ids_out[i] is the store id of the i-th record pmk_store_get returns (= the reference's
 * m_ppatches[i], patch_manager.cpp:75-104).  Right after pmk_filter_rebuild / pmk_filter the store is compact and ids_out[i] == i; after
 * pmk_store_add / pmk_propagate / pmk_store_remove it is not.  pmk_store_remove, pmk_store_update_depth_maps, pmk_probe_neighbors,
 * pmk_store_cell_ids and pmk_store_depth_map all speak store ids. */
int pmk_store_ids(pmk_ctx* ctx, int nmax, int* ids_out, int* n_out);
/* PatchManager::removePatch (patch_manager.cpp:303-325) for stored patches, by store id (pmk_store_ids). */
int pmk_store_remove(pmk_ctx* ctx, int n, const int* ids);
/* PatchManager::updateDepthMaps (patch_manager.cpp:191-221) for stored patches. */
int pmk_store_update_depth_maps(pmk_ctx* ctx, int n, const int* ids);
/* m_pgrids (which = 0) / m_vpgrids (1) of one view as CSR: offsets[cells + 1] and the patch ids of every cell in slot order.  ids may be
 * NULL to ask for the total only (total_out). */
int pmk_store_cell_ids(pmk_ctx* ctx, int view, int which, int* offsets, int* ids, int ids_cap, int* total_out);

/* Profiling aid: nanoseconds the last sweep spent on every dest cell (all views, view-major, row-major cells).  The first call
 * (out may be NULL) switches the recording on; every later call returns and clears the times. */
int pmk_debug_cell_times(pmk_ctx* ctx, float* out_total_cells);

/* Profiling aid: nanoseconds of CTA time the sweeps spent in each phase of a propagatePatch try, summed over all dest cells since the
 * last call; out16 holds 16 words: [0] generatePatch + computeNcc, [1] preProcess, [2] refinePatch, [3] postProcess (store-independent
 * part), [4] its store-reading tail (setVImagesVGrids, check), [5] commit, [6] tries, [7] tries that reached refinePatch, [8..15] the
 * steps of one cost evaluation inside refinePatch (builds with -DPMK_SUBPHASE only, else 0).  The first call (out16 may be NULL)
 * switches the recording on; every later call returns and clears the sums. */
int pmk_debug_phase_times(pmk_ctx* ctx, uint64_t* out16);

/* Stream control / timing helpers for bench.py (no reference counterpart). */
int pmk_sync(pmk_ctx* ctx);
int pmk_device_alloc(pmk_ctx* ctx, uint64_t bytes, void** out);
int pmk_device_free(pmk_ctx* ctx, void* p);
int pmk_memcpy_h2d(pmk_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int pmk_memcpy_d2h(pmk_ctx* ctx, void* dst, const void* src, uint64_t bytes);
int pmk_host_alloc_pinned(uint64_t bytes, void** out);
int pmk_host_free_pinned(void* p);
/* CUDA-event timer on the context stream: begin, ..., end -> milliseconds */
int pmk_timer_begin(pmk_ctx* ctx);
int pmk_timer_end(pmk_ctx* ctx, float* ms);
/* number of kernel launches issued by this context since creation */
int pmk_launch_count(pmk_ctx* ctx, uint64_t* out);
/* write at least `bytes` of device memory to evict L2 between timed iterations */
int pmk_flush_l2(pmk_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PMK_H */
