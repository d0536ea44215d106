// mvskit_b200/host/pmmvps.hpp -- host-side mirror of the reference's class surface for the PatchMatch path:
// Option (pmmvps/option.hpp:20-73), Patch (patch.hpp:23-74), PatchManager (patch_manager.hpp:31-107), PmMvps
// (pmmvps.hpp:25-107) and the sub-objects its driver and tests poke (m_photoSet, m_dnInit, m_propagate, m_optim, m_filter).
// Same names, same public members, same argument meaning, same error behaviour (fatal configuration errors print to
// cerr and exit(1); soft failures return -1; "reject" clears Patch::m_images).  Every method body is a call into the
// C ABI of include/pmk.h -- the grids, depth maps and all arithmetic live on the GPU; there is no CPU path.
//
// Eigen is a third-party dependency the reference does not vendor.  With -DPMK_HOST_USE_EIGEN the vector types below are
// Eigen's; otherwise a minimal stand-in with the same element access (v(i), v[i]) is used.
#ifndef PMK_HOST_PMMVPS_HPP
#define PMK_HOST_PMMVPS_HPP

#include <iostream>
#include <map>
#include <unordered_map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pmk.h"

#ifdef PMK_HOST_USE_EIGEN
#include "Eigen/Dense"
using Eigen::Vector2i;
using Eigen::Vector3f;
using Eigen::Vector3i;
using Eigen::Vector4f;
#else
namespace pmk_host {
template <typename T, int N>
struct Vec {
    T d[N];
    Vec() { for (int i = 0; i < N; ++i) d[i] = T(0); }
    Vec(T a, T b) { static_assert(N == 2, ""); d[0] = a; d[1] = b; }
    Vec(T a, T b, T c) { static_assert(N == 3, ""); d[0] = a; d[1] = b; d[2] = c; }
    Vec(T a, T b, T c, T e) { static_assert(N == 4, ""); d[0] = a; d[1] = b; d[2] = c; d[3] = e; }
    T& operator()(int i) { return d[i]; }
    const T& operator()(int i) const { return d[i]; }
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
};
}  // namespace pmk_host
typedef pmk_host::Vec<int, 2> Vector2i;
typedef pmk_host::Vec<float, 3> Vector3f;
typedef pmk_host::Vec<int, 3> Vector3i;
typedef pmk_host::Vec<float, 4> Vector4f;
#endif

using std::map;
using std::shared_ptr;
using std::string;
using std::vector;

// ---- Option (option.hpp:20-73) ---------------------------------------------------------------------------------------------------
struct Option {
public:
    Option();
    void init(const string prefix, const string option);

    int m_nimages, m_nillums, m_level, m_csize;
    float m_nccThreshold;
    int m_wsize, m_minImageNum, m_cpu, m_setEdge, m_useBound, m_useVisData, m_sequence;
    float m_maxAngleThreshold, m_quadThreshold;
    string m_prefix, m_option;
    int m_flag;
    vector<int> m_images;
    map<int, int> m_dict;
    vector<vector<int> > m_visdata, m_visdata2;

protected:
    void initVisdata();
};

// ---- Patch (patch.hpp:23-74) -----------------------------------------------------------------------------------------------------
class Patch {
public:
    Patch();
    float score2(const float threshold) const;

    Vector4f m_coord, m_normal;
    vector<int> m_images;
    vector<Vector2i> m_grids;
    vector<int> m_vimages;
    vector<Vector2i> m_vgrids;
    float m_ncc;
    int m_nimages, m_iter, m_collected, m_flag;
    unsigned char m_dflag;
    int m_fix, m_id;
    float m_dscale, m_ascale, m_tmp;
};
typedef shared_ptr<Patch> Ppatch;

std::istream& operator>>(std::istream& istr, Patch& rhs);
std::ostream& operator<<(std::ostream& ostr, const Patch& rhs);
std::istream& operator>>(std::istream& istr, Vector4f& v);
std::ostream& operator<<(std::ostream& ostr, const Vector4f& v);

class PmMvps;

// ---- PhotoSet: the part of image/photoSet.hpp the path's callers use ---------------------------------------------------------------------
class PhotoSet {
public:
    // reads <prefix>txt/%08d.txt ("CONTOUR" + 12 floats, camera.cpp:27-63) and <prefix>image/%04d0000.{jpg,ppm}; the image
    // payload must be binary PPM (JPEG decode is outside the accelerated path); uploads each view (pmk_set_view -> K0 pyramid)
    void init(PmMvps& pmmvps, const vector<int>& images, const string prefix, const int nimages, const int nillums, const int maxLevel,
              const int size, const int alloc);
    int getWidth(const int index, const int level) const;
    int getHeight(const int index, const int level) const;
    Vector3f project(const int index, const Vector4f& coord, const int level) const;      // Camera::project on the device (pmk_probe)
    // <base>.pgm (P5) or <base>.pbm (P4) -> grey values as Image::readPGMImage / readPBMImage deliver them (image.cpp:881-999)
    static bool readMask(const string& base, vector<unsigned char>& grey, int& w, int& h);
    int getMask(const Vector4f& coord, const int level) const;                             // photoSet.cpp:223-233 (0 / -1)
    int getMask(const int index, const Vector4f& coord, const int level) const;            // photo.cpp:44-52 (-1 / 0 / 255)
    int image2index(const int image) const;
    void setDistances() {}

    vector<int> m_images;
    int m_nimages, m_nillums;
    string m_prefix;
    map<int, int> m_dict;

private:
    PmMvps* m_pmmvps = nullptr;
};

// ---- PatchManager (patch_manager.hpp:31-107) ----------------------------------------------------------------------------------------------
class PatchManager {
public:
    PatchManager(PmMvps& pmmvps);
    void init();
    void image2index(Patch& patch);
    void index2image(Patch& patch);
    void collectPatches(const int target = 0);          // fills m_ppatches from the device store, in the reference's collect order
    void addPatch(Ppatch& ppatch);                      // setGrids result + registration in the device grids
    void setGrids(Patch& patch);
    void computeNcc(Patch& patch) const;                // K1 on one patch
    // K1 on a batch in one call: Patch objects are marshalled into the byte-lean wire format of pmk_ncc_eval_packed (3-float coord / normal,
    // byte view ids, 31 B per patch at 6 views instead of 60) when every patch qualifies, else into the wide one; same scores bit for bit
    void computeNcc(vector<Ppatch>& ppatches) const;
    void readPatches();
    void readPatches(const int iter);
    void writePatches(const string prefix, bool bExportPLY, bool bExportPatch, bool bExportPSet);
    void writePly(const vector<Ppatch>& ppatches, const string filename);
    void writePly(const vector<Ppatch>& ppatches, const string filename, const vector<Vector3i>& colors);
    // ---- the rest of the reference's public surface (patch_manager.hpp:31-107), each a pass-through to the device store ----
    void removePatch(const Ppatch& ppatch);                               // :303-325 (by Patch::m_id, the collect-order index)
    void updateDepthMaps(Ppatch& ppatch);                                 // :191-221
    void setGridsImages(Patch& patch, vector<int>& images) const;         // :223-239
    void setVGrids(Patch& patch);                                         // :251-261
    void setVImagesVGrids(Ppatch& ppatch);                                // :263-265
    void setVImagesVGrids(Patch& patch);                                  // :267-301
    int isVisible0(const Patch& patch, const int image, int& ix, int& iy, const float strict);   // :327-333
    int isVisible(const Patch& patch, const int image, const int& ix, const int& iy, const float strict);   // :335-376
    void setScales(Patch& patch) const;                                   // :378-399
    void sortPatches(vector<Ppatch>& ppatches, const int ascend = 1) const;    // :406-433
    void findNeighbors(const Patch& patch, vector<Ppatch>& neighbors, const float scale = 1.0f, const int margin = 1, const int skipvis = 0);   // :671-728
    // m_pgrids / m_vpgrids / m_dpgrids live in HBM; syncGrids() brings a host copy in the reference's shape (cell -> patches of
    // m_ppatches, which it refreshes first), so that callers that walk the public grids keep working.  m_dpgrids holds null for m_MAXDEPTH.
    void syncGrids();
    // sizes of m_pgrids / m_vpgrids cells and m_dpgrids ids of one view (the grids themselves stay in HBM)
    vector<int> cellCounts(const int image, const int vgrid = 0) const;
    vector<int> depthMap(const int image) const;

    vector<int> m_gheights, m_gwidths;
    vector<Ppatch> m_ppatches;
    // Patch::m_id = index in m_ppatches (collect order); the device speaks store ids, which equal the index only right after a rebuild
    vector<int> m_storeIds;                                               // m_ppatches index -> store id (collectPatches)
    int storeId(const Patch& patch) const;                                // -1 when the patch is not one collectPatches delivered
    Ppatch byStoreId(const int id) const;                                 // null when the id is not in m_ppatches
    vector<vector<vector<Ppatch> > > m_pgrids, m_vpgrids;                // [image][cell] (patch_manager.hpp:89-96), filled by syncGrids()
    vector<vector<Ppatch> > m_dpgrids;                                    // [image][cell] (patch_manager.hpp:100-104)

protected:
    void readPatchFile(const string& name);
    PmMvps& m_pmmvps;
    int m_nimages;
    std::unordered_map<int, int> m_indexOfStoreId;                        // store id -> m_ppatches index
};

class DepthNormInit {
public:
    DepthNormInit(PmMvps& pmmvps) : m_pmmvps(pmmvps), m_nplys(0) {}
    void init(const string prefix, const int nfiles) { m_prefix = prefix; m_nplys = nfiles; }
    void createPatches();                               // depth_normal_init.cpp:29-33, isTest branch: readPatches()
protected:
    PmMvps& m_pmmvps;
    string m_prefix;
    int m_nplys;
};

class Propagate {
public:
    Propagate(PmMvps& pmmvps) : MAX_NUM_OF_PATCHES(0), MAX_NUM_OF_PROPAG(0), m_pmmvps(pmmvps), m_ecount(0), m_fcount0(0), m_fcount1(0), m_pcount(0) {}
    void init();
    void run(const int iter);                           // K4 (pmk_propagate)
    int MAX_NUM_OF_PATCHES, MAX_NUM_OF_PROPAG;
    unsigned long long m_seed = 0x9E3779B97F4A7C15ull; // PMR1 key (the reference's NLopt BOBYQA has no seed)
    unsigned long long m_stats[16] = {0};
protected:
    PmMvps& m_pmmvps;
    int m_ecount, m_fcount0, m_fcount1, m_pcount;
};

class Optim {
public:
    Optim(PmMvps& pmmvps) : m_pmmvps(pmmvps) {}
    void init() {}
    int preProcess(Patch& patch);                       // pmk_pre_process
    void refinePatch(Patch& patch, const int time);     // pmk_refine (PMR1)
    int postProcess(Patch& patch);                      // pmk_post_process (store-independent part)
    float computeINCC(const Vector4f& coord, const Vector4f& normal, const vector<int>& indexes, const int isRobust);
    unsigned long long m_seed = 0x9E3779B97F4A7C15ull, m_stream = 0;
protected:
    PmMvps& m_pmmvps;
};

class Filter {
public:
    Filter(PmMvps& pmmvps) : m_pmmvps(pmmvps) {}
    void init() {}
    void run();                                         // K5..K9 (pmk_filter)
    int m_counts[6] = {0, 0, 0, 0, 0, 0};
protected:
    PmMvps& m_pmmvps;
};

// ---- PmMvps (pmmvps.hpp:25-107) -----------------------------------------------------------------------------------------------------------
class PmMvps {
public:
    PmMvps();
    virtual ~PmMvps();

    void init(const Option& option);
    void run();
    int isNeighborRadius(const Patch& lhs, const Patch& rhs, const float hunit, const float neighborThreshold, const float radius) const;
    int isNeighbor(const Patch& lhs, const Patch& rhs, const float hunit, const float neighborThreshold) const;
    int isNeighbor(const Patch& lhs, const Patch& rhs, const float neighborThreshold) const;

    int m_nimages, m_nillums;
    vector<int> m_images;
    string m_prefix;
    int m_level, m_csize;
    float m_nccThreshold;
    int m_wsize, m_minImageNumThreshold;
    vector<vector<int> > m_visdata, m_visdata2;
    float m_quadThreshold;
    int m_tau, m_depth;
    float m_angleThreshold0, m_angleThreshold1;
    int m_countThreshold1;
    float m_neighborThreshold, m_neighborThreshold1, m_neighborThreshold2;
    float m_nccThresholdBefore, m_maxAngleThreshold;

    PhotoSet m_photoSet;
    DepthNormInit m_dnInit;
    PatchManager m_patchManager;
    Propagate m_propagate;
    Optim m_optim;
    Filter m_filter;

    // the device context behind every member above (one per GPU); m_device / m_sweepGroup are read by init()
    pmk_ctx* m_ctx;
    int m_device, m_sweepGroup;
    void syncDepth();                                   // pushes m_depth / thresholds poked by the caller to the device

protected:
    void updateThreshold();
};

#endif /* PMK_HOST_PMMVPS_HPP */
