// mvskit_b200/host/pmmvps.cpp -- see pmmvps.hpp.  Host glue only: file formats of the reference (option file, CONTOUR camera
// files, .patch text, ASCII PLY) and calls into the C ABI.  Compile with -ffp-contract=off (isNeighbor mirrors the reference's
// float operation order on the host).
#include "pmmvps.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>

using std::cerr;
using std::endl;
using std::ifstream;
using std::ofstream;

namespace {

[[noreturn]] void die(const string& what) {
    cerr << what << endl;
    exit(1);
}

void chk(int rc, const char* where) {
    if (rc != PMK_OK) die(string(where) + ": " + pmk_last_error());
}

}  // namespace

// ---- Option ----------------------------------------------------------------------------------------------------------------------------
Option::Option() {                                                  // defaults of option.cpp:19-33
    m_nimages = 0; m_nillums = 1;
    m_level = 1; m_csize = 2; m_wsize = 7;
    m_nccThreshold = 0.7f;
    m_minImageNum = 3;
    m_cpu = 4;
    m_setEdge = 0; m_useBound = 0; m_useVisData = 0;
    m_sequence = -1;
    m_flag = -10;
    m_maxAngleThreshold = 10.0f * M_PI / 180.0f;
    m_quadThreshold = 2.5f;
}

void Option::init(const string prefix, const string option) {       // option.cpp:35-145: whitespace-separated "name value" pairs
    m_prefix = prefix;
    m_option = option;
    ifstream in((prefix + option).c_str());
    string name;
    while (in >> name) {
        if (name[0] == '#') { string rest; std::getline(in, rest); continue; }
        if (name == "image") in >> m_nimages;
        else if (name == "illum") in >> m_nillums;
        else if (name == "level") in >> m_level;
        else if (name == "csize") in >> m_csize;
        else if (name == "threshold") in >> m_nccThreshold;
        else if (name == "wsize") in >> m_wsize;
        else if (name == "minImageNum") in >> m_minImageNum;
        else if (name == "CPU") in >> m_cpu;
        else if (name == "setEdge") in >> m_setEdge;
        else if (name == "useBound") in >> m_useBound;
        else if (name == "useVisData") in >> m_useVisData;
        else if (name == "sequence") in >> m_sequence;
        else if (name == "maxAngle") { in >> m_maxAngleThreshold; m_maxAngleThreshold *= M_PI / 180.0f; }
        else if (name == "quad") in >> m_quadThreshold;
        else if (name == "images") {
            in >> m_flag;
            if (m_flag == -1) {
                int first = 0, last = 0;
                in >> first >> last;
                for (int i = first; i < last; ++i) m_images.push_back(i);
            } else if (0 < m_flag) {
                for (int i = 0; i < m_flag; ++i) { int index = 0; in >> index; m_images.push_back(index); }
            } else die("flag is not valid: " + std::to_string(m_flag));
        } else die("Unrecognizable option: " + name);
    }
    if (m_flag == -10) die("m_flag not specified: " + std::to_string(m_flag));
    for (int i = 0; i < (int)m_images.size(); ++i) m_dict[m_images[i]] = i;
    initVisdata();
}

void Option::initVisdata() {                                        // option.cpp:147-166: without vis.dat every other view is a candidate
    if (m_useVisData != 0) return;
    const int n = (int)m_images.size();
    m_visdata2.assign(n, vector<int>());
    for (int y = 0; y < n; ++y)
        for (int x = 0; x < n; ++x)
            if (x != y) m_visdata2[y].push_back(x);
}

// ---- Patch -----------------------------------------------------------------------------------------------------------------------------
Patch::Patch() {
    m_ncc = -1.0f;
    m_nimages = 0; m_iter = 0; m_collected = 0; m_flag = 0; m_dflag = 0; m_fix = 0; m_id = -1;
    m_dscale = 0.0f; m_ascale = 0.0f; m_tmp = 0.0f;
}

float Patch::score2(const float threshold) const { return std::max(0.0f, m_ncc - threshold) * (int)m_images.size(); }

std::istream& operator>>(std::istream& istr, Vector4f& v) { return istr >> v(0) >> v(1) >> v(2) >> v(3); }
std::ostream& operator<<(std::ostream& ostr, const Vector4f& v) { return ostr << v(0) << " " << v(1) << " " << v(2) << " " << v(3); }

std::istream& operator>>(std::istream& istr, Patch& rhs) {          // the .patch record of patch.cpp:31-56
    string header;
    istr >> header >> rhs.m_coord >> rhs.m_normal >> rhs.m_ncc >> rhs.m_dscale >> rhs.m_ascale;
    if (header == "PATCHA") { int type; Vector4f dir; istr >> type >> dir; }
    int n = 0;
    istr >> n;
    rhs.m_images.resize(std::max(n, 0));
    for (int i = 0; i < n; ++i) istr >> rhs.m_images[i];
    istr >> n;
    rhs.m_vimages.resize(std::max(n, 0));
    for (int i = 0; i < n; ++i) istr >> rhs.m_vimages[i];
    return istr;
}

std::ostream& operator<<(std::ostream& ostr, const Patch& rhs) {    // patch.cpp:58-80
    ostr << "PATCHES" << endl << rhs.m_coord << endl << rhs.m_normal << endl
         << rhs.m_ncc << ' ' << rhs.m_dscale << ' ' << rhs.m_ascale << endl
         << (int)rhs.m_images.size() << endl;
    for (size_t i = 0; i < rhs.m_images.size(); ++i) ostr << rhs.m_images[i] << ' ';
    ostr << endl << (int)rhs.m_vimages.size() << endl;
    for (size_t i = 0; i < rhs.m_vimages.size(); ++i) ostr << rhs.m_vimages[i] << ' ';
    ostr << endl;
    return ostr;
}

// ---- PhotoSet --------------------------------------------------------------------------------------------------------------------------
namespace {

// binary PPM (P6, maxval 255) -> interleaved RGB
bool read_ppm(const string& name, vector<unsigned char>& rgb, int& w, int& h) {
    FILE* f = fopen(name.c_str(), "rb");
    if (!f) return false;
    char magic[3] = {0, 0, 0};
    int maxval = 0;
    auto token = [&](int& out) {
        int c = fgetc(f);
        while (c == '#' || isspace(c)) { if (c == '#') while (c != '\n' && c != EOF) c = fgetc(f); c = fgetc(f); }
        out = 0;
        while (c >= '0' && c <= '9') { out = out * 10 + (c - '0'); c = fgetc(f); }
    };
    if (fread(magic, 1, 2, f) != 2 || magic[0] != 'P' || magic[1] != '6') { fclose(f); return false; }
    token(w); token(h); token(maxval);
    if (w <= 0 || h <= 0 || maxval != 255) { fclose(f); return false; }
    rgb.resize((size_t)w * h * 3);
    const bool ok = fread(rgb.data(), 1, rgb.size(), f) == rgb.size();
    fclose(f);
    return ok;
}

// <prefix>mask/%08d.pgm (binary P5) or .pbm (binary P4), the two formats Image::alloc accepts (image.cpp:143-161; completeName
// prefers .pgm, :82-89).  Like readPBMImage (:881-943) the P4 bits are taken as ONE stream (no per-row padding), bit set = 0.
}  // namespace
bool PhotoSet::readMask(const string& base, vector<unsigned char>& grey, int& w, int& h) {
    auto token = [](FILE* f, int& out) {
        int c = fgetc(f);
        while (c == '#' || isspace(c)) { if (c == '#') while (c != '\n' && c != EOF) c = fgetc(f); c = fgetc(f); }
        out = 0;
        while (c >= '0' && c <= '9') { out = out * 10 + (c - '0'); c = fgetc(f); }
    };
    char magic[2] = {0, 0};
    if (FILE* f = fopen((base + ".pgm").c_str(), "rb")) {
        int maxval = 0;
        bool ok = fread(magic, 1, 2, f) == 2 && magic[0] == 'P' && magic[1] == '5';
        if (ok) { token(f, w); token(f, h); token(f, maxval); ok = w > 0 && h > 0; }
        if (ok) { grey.resize((size_t)w * h); ok = fread(grey.data(), 1, grey.size(), f) == grey.size(); }
        fclose(f);
        if (!ok) cerr << "Only accept binary pgm format" << base << ".pgm" << endl;
        return ok;
    }
    if (FILE* f = fopen((base + ".pbm").c_str(), "rb")) {
        bool ok = fread(magic, 1, 2, f) == 2 && magic[0] == 'P' && magic[1] == '4';
        if (ok) { token(f, w); token(f, h); ok = w > 0 && h > 0; }
        if (ok) {
            const size_t n = (size_t)w * h;
            vector<unsigned char> bits((n + 7) / 8);
            ok = fread(bits.data(), 1, bits.size(), f) == bits.size();
            grey.resize(n);
            for (size_t i = 0; ok && i < n; ++i) grey[i] = ((bits[i >> 3] >> (7 - (i & 7))) & 1) ? 0 : 255;
        }
        fclose(f);
        if (!ok) cerr << "Only accept binary pbm format: " << base << ".pbm" << endl;
        return ok;
    }
    return false;
}

void PhotoSet::init(PmMvps& pmmvps, const vector<int>& images, const string prefix, const int nimages, const int nillums, const int maxLevel,
                    const int size, const int alloc) {
    (void)maxLevel; (void)size; (void)alloc;
    m_pmmvps = &pmmvps;
    m_images = images; m_nimages = nimages; m_nillums = nillums; m_prefix = prefix;
    for (int i = 0; i < m_nimages; ++i) m_dict[images[i]] = i;
    cerr << "Reading images: " << std::flush;
    for (int i = 0; i < m_nimages; ++i) {
        char cname[1024], iname[1024];
        snprintf(cname, sizeof(cname), "%stxt/%08d.txt", prefix.c_str(), i);                  // photoSet.cpp:47
        ifstream cam(cname);
        string header;
        float P[12];
        if (!(cam >> header)) die(string("Camera file not found: ") + cname);
        if (header == "CONTOUR") {                                                             // camera.cpp:39-53, txtType 0: the raw 3 x 4
            for (int k = 0; k < 12; ++k) cam >> P[k];
        } else if (header == "CONTOUR2") {                                                     // txtType 2: K, Euler angles in degrees, translation
            float intr[6], extr[6];
            for (int k = 0; k < 6; ++k) cam >> intr[k];
            for (int k = 0; k < 6; ++k) cam >> extr[k];
            chk(pmk_contour2_to_projection(intr, extr, P), "pmk_contour2_to_projection");      // camera.cpp:116-131, quat2proj :241-261
        } else die(string("Unrecognizable txt format: ") + cname);                             // CONTOUR3 is "to be written" in the reference too (:133-135)
        vector<unsigned char> rgb;
        int w = 0, h = 0;
        bool ok = false;
        const char* ext[2] = {"ppm", "jpg"};                                                  // completeName prefers .ppm (image.cpp:61-65)
        for (int e = 0; e < 2 && !ok; ++e) {
            snprintf(iname, sizeof(iname), "%simage/%04d%04d.%s", prefix.c_str(), i, 0, ext[e]);
            ok = read_ppm(iname, rgb, w, h);
        }
        if (ok) chk(pmk_set_view(pmmvps.m_ctx, i, P, rgb.data(), w, h), "pmk_set_view");
        else {
            // a real JPEG: decoded on the device (nvJPEG) where the reference calls CImg::load_jpeg (image.cpp:834-837)
            snprintf(iname, sizeof(iname), "%simage/%04d%04d.jpg", prefix.c_str(), i, 0);
            FILE* f = fopen(iname, "rb");
            if (!f) die(string("Image not found: ") + iname);
            vector<unsigned char> bytes;
            unsigned char buf[65536];
            size_t got;
            while ((got = fread(buf, 1, sizeof(buf), f)) > 0) bytes.insert(bytes.end(), buf, buf + got);
            fclose(f);
            chk(pmk_set_view_jpeg(pmmvps.m_ctx, i, P, bytes.data(), bytes.size(), &w, &h), "pmk_set_view_jpeg");
        }
        char mname[1024];
        snprintf(mname, sizeof(mname), "%smask/%08d", prefix.c_str(), i);                    // photoSet.cpp:48
        vector<unsigned char> grey;
        int mw = 0, mh = 0;
        if (readMask(mname, grey, mw, mh)) {
            cerr << "Read mask: " << mname << endl;                                           // image.cpp:148
            chk(pmk_set_view_mask(pmmvps.m_ctx, i, grey.data(), mw, mh), "pmk_set_view_mask");
        }
        cerr << "*" << std::flush;
    }
    cerr << endl;
}

int PhotoSet::getWidth(const int index, const int level) const { int w = 0, h = 0; chk(pmk_get_level_dims(m_pmmvps->m_ctx, index, level, &w, &h), "getWidth"); return w; }
int PhotoSet::getHeight(const int index, const int level) const { int w = 0, h = 0; chk(pmk_get_level_dims(m_pmmvps->m_ctx, index, level, &w, &h), "getHeight"); return h; }

Vector3f PhotoSet::project(const int index, const Vector4f& coord, const int level) const {
    if (level != m_pmmvps->m_level) die("PhotoSet::project: only the working level is resident on the device");
    float c[4] = {coord(0), coord(1), coord(2), coord(3)}, out[3];
    chk(pmk_probe(m_pmmvps->m_ctx, 1, &index, c, nullptr, out, nullptr, nullptr, nullptr, nullptr, nullptr), "project");
    return Vector3f(out[0], out[1], out[2]);
}

// PhotoSet::getMask(coord, level) (photoSet.cpp:223-233) and getMask(index, coord, level) (:219-221), on the device (pmk_probe_mask)
int PhotoSet::getMask(const Vector4f& coord, const int level) const { return getMask(-1, coord, level); }
int PhotoSet::getMask(const int index, const Vector4f& coord, const int level) const {
    if (level != m_pmmvps->m_level) die("PhotoSet::getMask: only the working level is resident on the device");
    const float c[4] = {coord(0), coord(1), coord(2), coord(3)};
    int out = -1;
    chk(pmk_probe_mask(m_pmmvps->m_ctx, 1, index, c, &out), "getMask");
    return out;
}

int PhotoSet::image2index(const int image) const {
    map<int, int>::const_iterator it = m_dict.find(image);
    return it == m_dict.end() ? -1 : it->second;
}

// ---- PatchManager ----------------------------------------------------------------------------------------------------------------------
PatchManager::PatchManager(PmMvps& pmmvps) : m_pmmvps(pmmvps), m_nimages(0) {}

void PatchManager::init() {                                         // patch_manager.cpp:24-52
    m_nimages = m_pmmvps.m_nimages;
    m_gheights.assign(m_nimages, 0);
    m_gwidths.assign(m_nimages, 0);
    for (int i = 0; i < m_nimages; ++i) chk(pmk_get_grid_dims(m_pmmvps.m_ctx, i, &m_gwidths[i], &m_gheights[i]), "pmk_get_grid_dims");
    chk(pmk_store_clear(m_pmmvps.m_ctx), "pmk_store_clear");
    m_ppatches.clear();
}

void PatchManager::image2index(Patch& patch) {
    vector<int> out;
    for (size_t i = 0; i < patch.m_images.size(); ++i) {
        const int index = m_pmmvps.m_photoSet.image2index(patch.m_images[i]);
        if (index != -1) out.push_back(index);
    }
    patch.m_images.swap(out);
}

void PatchManager::index2image(Patch& patch) {
    for (size_t i = 0; i < patch.m_images.size(); ++i) patch.m_images[i] = m_pmmvps.m_photoSet.m_images[patch.m_images[i]];
    for (size_t i = 0; i < patch.m_vimages.size(); ++i) patch.m_vimages[i] = m_pmmvps.m_photoSet.m_images[patch.m_vimages[i]];
}

void PatchManager::collectPatches(const int target) {
    (void)target;                                                   // m_fix is never set on this path
    m_pmmvps.syncDepth();
    int n = 0;
    chk(pmk_store_count(m_pmmvps.m_ctx, &n), "pmk_store_count");
    const int maxv = m_nimages;
    vector<float> coord((size_t)n * 4), normal((size_t)n * 4), scal((size_t)n * 4);
    vector<int> images((size_t)n * maxv), nimg(n), grids((size_t)n * maxv * 2), vimages((size_t)n * maxv), nvimg(n), vgrids((size_t)n * maxv * 2);
    int got = 0;
    chk(pmk_store_get(m_pmmvps.m_ctx, n, maxv, coord.data(), normal.data(), scal.data(), images.data(), nimg.data(), grids.data(),
                      vimages.data(), nvimg.data(), vgrids.data(), &got), "pmk_store_get");
    m_ppatches.clear();
    m_ppatches.reserve(got);
    for (int i = 0; i < std::min(n, got); ++i) {
        Ppatch pp(new Patch());
        Patch& p = *pp;
        for (int k = 0; k < 4; ++k) { p.m_coord(k) = coord[4 * i + k]; p.m_normal(k) = normal[4 * i + k]; }
        p.m_ncc = scal[4 * i]; p.m_dscale = scal[4 * i + 1]; p.m_ascale = scal[4 * i + 2]; p.m_tmp = scal[4 * i + 3];
        p.m_nimages = nimg[i];
        p.m_id = i;
        for (int k = 0; k < nimg[i]; ++k) {
            p.m_images.push_back(images[(size_t)i * maxv + k]);
            p.m_grids.push_back(Vector2i(grids[((size_t)i * maxv + k) * 2], grids[((size_t)i * maxv + k) * 2 + 1]));
        }
        for (int k = 0; k < nvimg[i]; ++k) {
            p.m_vimages.push_back(vimages[(size_t)i * maxv + k]);
            p.m_vgrids.push_back(Vector2i(vgrids[((size_t)i * maxv + k) * 2], vgrids[((size_t)i * maxv + k) * 2 + 1]));
        }
        m_ppatches.push_back(pp);
    }
    m_storeIds.assign(m_ppatches.size(), -1);
    m_indexOfStoreId.clear();
    int nid = 0;
    if (!m_ppatches.empty()) chk(pmk_store_ids(m_pmmvps.m_ctx, (int)m_storeIds.size(), m_storeIds.data(), &nid), "pmk_store_ids");
    for (int i = 0; i < (int)m_storeIds.size(); ++i) m_indexOfStoreId[m_storeIds[i]] = i;
}

int PatchManager::storeId(const Patch& patch) const {
    return (patch.m_id >= 0 && patch.m_id < (int)m_storeIds.size()) ? m_storeIds[patch.m_id] : -1;
}

Ppatch PatchManager::byStoreId(const int id) const {
    const auto it = m_indexOfStoreId.find(id);
    return it == m_indexOfStoreId.end() ? Ppatch() : m_ppatches[it->second];
}

namespace {

int add_records(PmMvps& pm, const vector<Ppatch>& pps) {
    const int n = (int)pps.size(), V = pm.m_nimages;
    if (n == 0) return 0;
    vector<float> coord((size_t)n * 4), normal((size_t)n * 4), scal((size_t)n * 4);
    vector<int> images((size_t)n * V, 0), nimg(n);
    for (int i = 0; i < n; ++i) {
        const Patch& p = *pps[i];
        for (int k = 0; k < 4; ++k) { coord[4 * i + k] = p.m_coord(k); normal[4 * i + k] = p.m_normal(k); }
        scal[4 * i] = p.m_ncc; scal[4 * i + 1] = p.m_dscale; scal[4 * i + 2] = p.m_ascale; scal[4 * i + 3] = p.m_tmp;
        nimg[i] = std::min((int)p.m_images.size(), V);
        for (int k = 0; k < nimg[i]; ++k) images[(size_t)i * V + k] = p.m_images[k];
    }
    pm.syncDepth();
    return pmk_store_add(pm.m_ctx, n, coord.data(), normal.data(), scal.data(), images.data(), nimg.data(), V);
}

}  // namespace

void PatchManager::setGrids(Patch& patch) {                         // patch_manager.cpp:241-249, cell index on the device (bit-exact)
    patch.m_grids.clear();
    const int n = (int)patch.m_images.size();
    if (n == 0) return;
    vector<float> c((size_t)n * 4);
    vector<int> cell((size_t)n * 2);
    for (int i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) c[4 * i + k] = patch.m_coord(k);
    chk(pmk_probe(m_pmmvps.m_ctx, n, patch.m_images.data(), c.data(), nullptr, nullptr, nullptr, nullptr, nullptr, cell.data(), nullptr), "setGrids");
    for (int i = 0; i < n; ++i) patch.m_grids.push_back(Vector2i(cell[2 * i], cell[2 * i + 1]));
}

void PatchManager::addPatch(Ppatch& ppatch) {
    vector<Ppatch> one(1, ppatch);
    chk(add_records(m_pmmvps, one), "addPatch");
}

void PatchManager::computeNcc(Patch& patch) const {                 // patch_manager.cpp:401-404
    const int n = (int)patch.m_images.size();
    float c[4] = {patch.m_coord(0), patch.m_coord(1), patch.m_coord(2), patch.m_coord(3)};
    float m[4] = {patch.m_normal(0), patch.m_normal(1), patch.m_normal(2), patch.m_normal(3)};
    float incc = 2.0f, ncc = 0.0f;
    if (n > 0) chk(pmk_ncc_eval(m_pmmvps.m_ctx, 1, c, m, patch.m_images.data(), &n, n, &incc, &ncc, nullptr), "computeNcc");
    patch.m_ncc = ncc;
}

void PatchManager::computeNcc(vector<Ppatch>& ppatches) const {      // patch_manager.cpp:401-404 over a batch, one K1 launch
    const int n = (int)ppatches.size();
    if (n == 0) return;
    int stride = 1;
    bool lean = m_nimages <= 255;                                     // byte view ids, 255 = no view
    for (const Ppatch& pp : ppatches) {
        stride = std::max(stride, (int)pp->m_images.size());
        lean = lean && pp->m_coord(3) == 1.0f && pp->m_normal(3) == 0.0f && pp->m_images.size() <= 255;
        for (int v : pp->m_images) lean = lean && v >= 0 && v < 255;
    }
    stride = std::min(stride, 255);
    vector<float> incc(n, 2.0f), ncc(n, 0.0f);
    if (lean) {
        vector<float> c((size_t)n * 3), m((size_t)n * 3);
        vector<uint8_t> views((size_t)n * stride, 255), counts(n);
        for (int i = 0; i < n; ++i) {
            const Patch& p = *ppatches[i];
            for (int k = 0; k < 3; ++k) { c[3 * (size_t)i + k] = p.m_coord(k); m[3 * (size_t)i + k] = p.m_normal(k); }
            const int ni = std::min((int)p.m_images.size(), stride);
            counts[i] = (uint8_t)ni;
            for (int k = 0; k < ni; ++k) views[(size_t)i * stride + k] = (uint8_t)p.m_images[k];
        }
        chk(pmk_ncc_eval_packed(m_pmmvps.m_ctx, n, c.data(), m.data(), views.data(), counts.data(), stride, incc.data(), ncc.data(), nullptr), "computeNcc");
    } else {
        vector<float> c((size_t)n * 4), m((size_t)n * 4);
        vector<int> views((size_t)n * stride, -1), counts(n);
        for (int i = 0; i < n; ++i) {
            const Patch& p = *ppatches[i];
            for (int k = 0; k < 4; ++k) { c[4 * (size_t)i + k] = p.m_coord(k); m[4 * (size_t)i + k] = p.m_normal(k); }
            const int ni = std::min((int)p.m_images.size(), stride);
            counts[i] = ni;
            for (int k = 0; k < ni; ++k) views[(size_t)i * stride + k] = p.m_images[k];
        }
        chk(pmk_ncc_eval(m_pmmvps.m_ctx, n, c.data(), m.data(), views.data(), counts.data(), stride, incc.data(), ncc.data(), nullptr), "computeNcc");
    }
    for (int i = 0; i < n; ++i) ppatches[i]->m_ncc = ppatches[i]->m_images.empty() ? 0.0f : ncc[i];
}

void PatchManager::readPatchFile(const string& name) {              // patch_manager.cpp:435-497
    ifstream in(name.c_str());
    if (!in.is_open()) return;
    string header;
    int pnum = 0;
    in >> header >> pnum;
    vector<Ppatch> pps;
    for (int p = 0; p < pnum; ++p) {
        Ppatch pp(new Patch());
        in >> *pp;
        pp->m_fix = 0;
        pp->m_vimages.clear();
        image2index(*pp);
        if (pp->m_images.empty()) break;                            // the reference returns here, dropping the rest of the file
        pp->m_tmp = pp->score2(m_pmmvps.m_nccThreshold);
        pps.push_back(pp);
    }
    chk(add_records(m_pmmvps, pps), "readPatches");
}

void PatchManager::readPatches() {
    char buffer[1024];
    snprintf(buffer, sizeof(buffer), "%sply/%08d.patch", m_pmmvps.m_prefix.c_str(), 0);
    readPatchFile(buffer);
}

void PatchManager::readPatches(const int iter) {
    char buffer[1024];
    snprintf(buffer, sizeof(buffer), "%sply/%08d.patch", m_pmmvps.m_prefix.c_str(), iter);
    readPatchFile(buffer);
}

void PatchManager::writePatches(const string prefix, bool bExportPLY, bool bExportPatch, bool bExportPSet) {
    (void)bExportPSet;                                              // commented out in the reference (patch_manager.cpp:523-538)
    collectPatches(1);
    if (bExportPLY) writePly(m_ppatches, prefix + ".ply");
    if (bExportPatch) {
        ofstream out((prefix + ".patch").c_str());
        out << "PATCHES" << endl << (int)m_ppatches.size() << endl;
        for (size_t p = 0; p < m_ppatches.size(); ++p) {
            Patch patch = *m_ppatches[p];
            index2image(patch);
            out << patch << "\n";
        }
    }
}

namespace {
void ply_header(ofstream& out, int n) {                              // patch_manager.cpp:544-556
    out << "ply" << '\n' << "format ascii 1.0" << '\n' << "element vertex " << n << '\n'
        << "property float x" << '\n' << "property float y" << '\n' << "property float z" << '\n'
        << "property float nx" << '\n' << "property float ny" << '\n' << "property float nz" << '\n'
        << "property uchar diffuse_red" << '\n' << "property uchar diffuse_green" << '\n' << "property uchar diffuse_blue" << '\n'
        << "end_header" << '\n';
}
}  // namespace

void PatchManager::writePly(const vector<Ppatch>& patches, const string filename) {
    // colours of m_ppatches order come from the device (mean of Image::getColor over m_images); `patches` must be m_ppatches
    const int n = (int)patches.size();
    vector<unsigned char> rgb((size_t)std::max(n, 1) * 3, 0);
    if (n > 0) chk(pmk_store_colors(m_pmmvps.m_ctx, n, rgb.data()), "pmk_store_colors");
    vector<Vector3i> colors(n);
    for (int i = 0; i < n; ++i) colors[i] = Vector3i(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
    writePly(patches, filename, colors);
}

void PatchManager::writePly(const vector<Ppatch>& patches, const string filename, const vector<Vector3i>& colors) {
    ofstream out(filename.c_str());
    ply_header(out, (int)patches.size());
    for (size_t i = 0; i < patches.size(); ++i) {
        const Patch& p = *patches[i];
        out << p.m_coord(0) << ' ' << p.m_coord(1) << ' ' << p.m_coord(2) << ' ' << p.m_normal(0) << ' ' << p.m_normal(1) << ' ' << p.m_normal(2) << ' '
            << colors[i](0) << ' ' << colors[i](1) << ' ' << colors[i](2) << '\n';
    }
}

// ---- pass-throughs: the arithmetic and the grids are the device's, only the Patch objects are marshalled ---------------------------
namespace {
struct PatchRec {
    float c[4], m[4], s[4];
    vector<int> images;
    int n;
    explicit PatchRec(const Patch& p) : images(p.m_images), n((int)p.m_images.size()) {
        for (int k = 0; k < 4; ++k) { c[k] = p.m_coord(k); m[k] = p.m_normal(k); }
        s[0] = p.m_ncc; s[1] = p.m_dscale; s[2] = p.m_ascale; s[3] = p.m_tmp;
    }
};
}  // namespace

void PatchManager::removePatch(const Ppatch& ppatch) {
    const int id = storeId(*ppatch);                                    // m_id is the collect-order index (collectPatches)
    if (id < 0) { cerr << "removePatch: not a stored patch (collectPatches first)" << endl; return; }
    chk(pmk_store_remove(m_pmmvps.m_ctx, 1, &id), "removePatch");
}

void PatchManager::updateDepthMaps(Ppatch& ppatch) {
    const int id = storeId(*ppatch);
    if (id < 0) { cerr << "updateDepthMaps: not a stored patch (collectPatches first)" << endl; return; }
    chk(pmk_store_update_depth_maps(m_pmmvps.m_ctx, 1, &id), "updateDepthMaps");
}

void PatchManager::setGridsImages(Patch& patch, vector<int>& images) const {
    patch.m_images.clear();
    patch.m_grids.clear();
    const int n = (int)images.size();
    if (n == 0) return;
    vector<float> c((size_t)n * 4);
    vector<int> cell((size_t)n * 2), ok(n);
    for (int i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) c[4 * i + k] = patch.m_coord(k);
    chk(pmk_probe(m_pmmvps.m_ctx, n, images.data(), c.data(), nullptr, nullptr, nullptr, nullptr, nullptr, cell.data(), ok.data()), "setGridsImages");
    for (int i = 0; i < n; ++i)
        if (ok[i]) { patch.m_images.push_back(images[i]); patch.m_grids.push_back(Vector2i(cell[2 * i], cell[2 * i + 1])); }
}

void PatchManager::setVGrids(Patch& patch) {
    patch.m_vgrids.clear();
    const int n = (int)patch.m_vimages.size();
    if (n == 0) return;
    vector<float> c((size_t)n * 4);
    vector<int> cell((size_t)n * 2);
    for (int i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) c[4 * i + k] = patch.m_coord(k);
    chk(pmk_probe(m_pmmvps.m_ctx, n, patch.m_vimages.data(), c.data(), nullptr, nullptr, nullptr, nullptr, nullptr, cell.data(), nullptr), "setVGrids");
    for (int i = 0; i < n; ++i) patch.m_vgrids.push_back(Vector2i(cell[2 * i], cell[2 * i + 1]));
}

void PatchManager::setVImagesVGrids(Ppatch& ppatch) { setVImagesVGrids(*ppatch); }

void PatchManager::setVImagesVGrids(Patch& patch) {
    m_pmmvps.syncDepth();
    vector<char> seen(m_nimages, 0);
    for (size_t i = 0; i < patch.m_images.size(); ++i) seen[patch.m_images[i]] = 1;
    for (size_t i = 0; i < patch.m_vimages.size(); ++i) seen[patch.m_vimages[i]] = 1;
    vector<int> cand;
    for (int v = 0; v < m_nimages; ++v) if (!seen[v]) cand.push_back(v);
    const int n = (int)cand.size();
    if (n == 0) return;
    vector<float> c((size_t)n * 4), m((size_t)n * 4);
    for (int i = 0; i < n; ++i) for (int k = 0; k < 4; ++k) { c[4 * i + k] = patch.m_coord(k); m[4 * i + k] = patch.m_normal(k); }
    vector<int> vis(n), cell((size_t)n * 2);
    chk(pmk_probe_visible(m_pmmvps.m_ctx, n, c.data(), m.data(), cand.data(), nullptr, m_pmmvps.m_neighborThreshold, vis.data(), cell.data()), "setVImagesVGrids");
    for (int i = 0; i < n; ++i)
        if (vis[i]) { patch.m_vimages.push_back(cand[i]); patch.m_vgrids.push_back(Vector2i(cell[2 * i], cell[2 * i + 1])); }
}

int PatchManager::isVisible0(const Patch& patch, const int image, int& ix, int& iy, const float strict) {
    m_pmmvps.syncDepth();
    const PatchRec r(patch);
    int vis = 0, cell[2] = {0, 0};
    chk(pmk_probe_visible(m_pmmvps.m_ctx, 1, r.c, r.m, &image, nullptr, strict, &vis, cell), "isVisible0");
    ix = cell[0]; iy = cell[1];
    return vis;
}

int PatchManager::isVisible(const Patch& patch, const int image, const int& ix, const int& iy, const float strict) {
    m_pmmvps.syncDepth();
    const PatchRec r(patch);
    const int cell[2] = {ix, iy};
    int vis = 0;
    chk(pmk_probe_visible(m_pmmvps.m_ctx, 1, r.c, r.m, &image, cell, strict, &vis, nullptr), "isVisible");
    return vis;
}

void PatchManager::setScales(Patch& patch) const {
    const PatchRec r(patch);
    if (r.n < 1) return;
    float ds = 0.0f, as = 0.0f;
    chk(pmk_probe_scales(m_pmmvps.m_ctx, 1, r.c, r.images.data(), &r.n, r.n, &ds, &as), "setScales");
    patch.m_dscale = ds; patch.m_ascale = as;
}

void PatchManager::sortPatches(vector<Ppatch>& ppatches, const int ascend) const {
    const int npatches = (int)ppatches.size();
    if (npatches == 0) return;
    vector<Ppatch> unscored;                                                                              // :411-415, scored in one K1 launch
    for (int n = 0; n < npatches; ++n) if (ppatches[n]->m_ncc < 0.0f) unscored.push_back(ppatches[n]);
    computeNcc(unscored);
    if (npatches == 1) return;
    for (int i = 0; i < npatches; ++i)                                                                    // the reference's swap sort, :419-432
        for (int j = i + 1; j < npatches; ++j) {
            const bool swap = ascend ? (ppatches[i]->m_ncc > ppatches[j]->m_ncc) : (ppatches[i]->m_ncc < ppatches[j]->m_ncc);
            if (swap) ppatches[i].swap(ppatches[j]);
        }
}

void PatchManager::findNeighbors(const Patch& patch, vector<Ppatch>& neighbors, const float scale, const int margin, const int skipvis) {
    (void)skipvis;                                                       // unused in the reference as well (:671)
    m_pmmvps.syncDepth();
    const PatchRec r(patch);
    if (r.n < 1) return;
    int cap = 1024, count = 0;
    vector<int> ids(cap);
    chk(pmk_probe_neighbors(m_pmmvps.m_ctx, 1, r.c, r.m, r.s, r.images.data(), &r.n, r.n, scale, margin, cap, ids.data(), &count), "findNeighbors");
    if (count > cap) {
        cap = count; ids.resize(cap);
        chk(pmk_probe_neighbors(m_pmmvps.m_ctx, 1, r.c, r.m, r.s, r.images.data(), &r.n, r.n, scale, margin, cap, ids.data(), &count), "findNeighbors");
    }
    for (int k = 0; k < std::min(count, cap); ++k) {                      // store ids -> m_ppatches (collectPatches)
        const Ppatch pp = byStoreId(ids[k]);
        if (pp) neighbors.push_back(pp);
    }
}

void PatchManager::syncGrids() {
    collectPatches(0);
    m_pgrids.assign(m_nimages, vector<vector<Ppatch> >());
    m_vpgrids.assign(m_nimages, vector<vector<Ppatch> >());
    m_dpgrids.assign(m_nimages, vector<Ppatch>());
    for (int v = 0; v < m_nimages; ++v) {
        const int nc = m_gwidths[v] * m_gheights[v];
        for (int which = 0; which < 2; ++which) {
            vector<int> offs(nc + 1);
            int total = 0;
            chk(pmk_store_cell_ids(m_pmmvps.m_ctx, v, which, offs.data(), nullptr, 0, &total), "syncGrids");
            vector<int> ids(std::max(total, 1));
            if (total > 0) chk(pmk_store_cell_ids(m_pmmvps.m_ctx, v, which, offs.data(), ids.data(), total, &total), "syncGrids");
            vector<vector<Ppatch> >& g = which ? m_vpgrids[v] : m_pgrids[v];
            g.assign(nc, vector<Ppatch>());
            for (int c = 0; c < nc; ++c)
                for (int k = offs[c]; k < offs[c + 1]; ++k) {
                    const Ppatch pp = byStoreId(ids[k]);
                    if (pp) g[c].push_back(pp);
                }
        }
        const vector<int> dm = depthMap(v);
        m_dpgrids[v].assign(nc, Ppatch());
        for (int c = 0; c < nc; ++c) if (dm[c] >= 0) m_dpgrids[v][c] = byStoreId(dm[c]);
    }
}

vector<int> PatchManager::cellCounts(const int image, const int vgrid) const {
    vector<int> out((size_t)m_gwidths[image] * m_gheights[image]);
    chk(pmk_store_cell_counts(m_pmmvps.m_ctx, image, vgrid, out.data()), "pmk_store_cell_counts");
    return out;
}

vector<int> PatchManager::depthMap(const int image) const {
    vector<int> out((size_t)m_gwidths[image] * m_gheights[image]);
    chk(pmk_store_depth_map(m_pmmvps.m_ctx, image, out.data()), "pmk_store_depth_map");
    return out;
}

// ---- DepthNormInit / Propagate / Optim / Filter ------------------------------------------------------------------------------------------
void DepthNormInit::createPatches() { m_pmmvps.m_patchManager.readPatches(); }

void Propagate::init() {                                            // propagate.cpp:22-26
    MAX_NUM_OF_PROPAG = 2;
    MAX_NUM_OF_PATCHES = MAX_NUM_OF_PROPAG * m_pmmvps.m_csize * m_pmmvps.m_csize;
}

void Propagate::run(const int iter) {                               // propagate.cpp:28-64
    time_t start = time(NULL);
    cerr << "Expanding patches..." << std::flush;
    m_pmmvps.syncDepth();
    chk(pmk_propagate(m_pmmvps.m_ctx, iter, m_seed, (uint64_t*)m_stats), "pmk_propagate");
    m_ecount = (int)m_stats[1]; m_fcount0 = (int)m_stats[4]; m_fcount1 = (int)m_stats[5]; m_pcount = (int)(m_stats[6] + m_stats[7]);
    cerr << endl << "---- EXPANSION: " << (time(NULL) - start) << " secs ----" << endl;
    cerr << "total pass fail0 fail1 refinepatch: " << m_ecount << " " << m_pcount << " " << m_fcount0 << " " << m_fcount1 << " " << m_pcount + m_fcount1 << endl;
}

namespace {
struct One {
    float c[4], m[4];
    vector<int> images;
    int n;
    One(const Patch& p) : images(p.m_images), n((int)p.m_images.size()) {
        for (int k = 0; k < 4; ++k) { c[k] = p.m_coord(k); m[k] = p.m_normal(k); }
    }
};
}  // namespace

int Optim::preProcess(Patch& patch) {                               // optim.cpp:137-163
    One o(patch);
    const int V = m_pmmvps.m_nimages;
    if (o.n == 0) return -1;
    vector<int> out(V);
    int ret = -1, nout = 0;
    float ds = 0.0f, as = 0.0f;
    m_pmmvps.syncDepth();
    chk(pmk_pre_process(m_pmmvps.m_ctx, 1, o.c, o.m, o.images.data(), &o.n, o.n, V, &ret, out.data(), &nout, &ds, &as), "preProcess");
    patch.m_images.assign(out.begin(), out.begin() + nout);
    patch.m_dscale = ds; patch.m_ascale = as;
    return ret;
}

void Optim::refinePatch(Patch& patch, const int time) {             // optim.cpp:470-547, PMR1 in place of NLopt BOBYQA
    (void)time;
    One o(patch);
    if (o.n == 0) return;
    float ncc = 0.0f;
    const uint64_t stream = m_stream;
    m_pmmvps.syncDepth();
    chk(pmk_refine(m_pmmvps.m_ctx, 1, o.c, o.m, &patch.m_dscale, o.images.data(), &o.n, o.n, &stream, m_seed, &ncc, nullptr), "refinePatch");
    for (int k = 0; k < 4; ++k) { patch.m_coord(k) = o.c[k]; patch.m_normal(k) = o.m[k]; }
    patch.m_ncc = ncc;
}

int Optim::postProcess(Patch& patch) {                              // optim.cpp:260-290 (the store-reading tail runs inside pmk_propagate)
    One o(patch);
    const int V = m_pmmvps.m_nimages;
    if (o.n == 0) return -1;
    vector<int> out(V), grids((size_t)V * 2);
    int ret = -1, nout = 0;
    float tmp = 0.0f;
    m_pmmvps.syncDepth();
    chk(pmk_post_process(m_pmmvps.m_ctx, 1, o.c, o.m, &patch.m_ncc, o.images.data(), &o.n, o.n, V, &ret, out.data(), &nout, grids.data(), &tmp), "postProcess");
    if (ret == 0) {
        patch.m_images.assign(out.begin(), out.begin() + nout);
        patch.m_grids.clear();
        for (int i = 0; i < nout; ++i) patch.m_grids.push_back(Vector2i(grids[2 * i], grids[2 * i + 1]));
        patch.m_nimages = nout;
        patch.m_tmp = tmp;
    }
    return ret;
}

float Optim::computeINCC(const Vector4f& coord, const Vector4f& normal, const vector<int>& indexes, const int isRobust) {
    if (!isRobust) die("Optim::computeINCC: only the robust form is on the accelerated path");
    const int n = (int)indexes.size();
    if (n < 2) return 2.0f;
    float c[4] = {coord(0), coord(1), coord(2), coord(3)}, m[4] = {normal(0), normal(1), normal(2), normal(3)}, incc = 2.0f;
    chk(pmk_ncc_eval(m_pmmvps.m_ctx, 1, c, m, indexes.data(), &n, n, &incc, nullptr, nullptr), "computeINCC");
    return incc;
}

void Filter::run() {                                                // filter.cpp:25-49
    m_pmmvps.syncDepth();
    chk(pmk_filter(m_pmmvps.m_ctx, m_counts), "pmk_filter");
    cerr << "FilterOutside " << m_counts[0] << " -> " << m_counts[0] - m_counts[1] << ", Filter Exact -" << m_counts[2] << ", FilterNeighbor -" << m_counts[3]
         << ", FilterGroups -" << m_counts[4] << " => " << m_counts[5] << endl;
}

// ---- PmMvps ----------------------------------------------------------------------------------------------------------------------------
PmMvps::PmMvps() : m_dnInit(*this), m_patchManager(*this), m_propagate(*this), m_optim(*this), m_filter(*this), m_ctx(nullptr), m_device(0), m_sweepGroup(1) {
    m_nimages = m_nillums = 0; m_level = 1; m_csize = 2; m_nccThreshold = 0.7f; m_wsize = 7; m_minImageNumThreshold = 3;
    m_quadThreshold = 2.5f; m_tau = 0; m_depth = 0; m_angleThreshold0 = m_angleThreshold1 = 0.0f; m_countThreshold1 = 4;
    m_neighborThreshold = m_neighborThreshold1 = m_neighborThreshold2 = 0.0f; m_nccThresholdBefore = 0.0f; m_maxAngleThreshold = 0.0f;
}

PmMvps::~PmMvps() { if (m_ctx) pmk_destroy(m_ctx); }

void PmMvps::init(const Option& option) {                           // pmmvps.cpp:18-68
    m_images = option.m_images;
    m_nimages = option.m_nimages;
    m_nillums = option.m_nillums;
    m_prefix = option.m_prefix;
    m_level = option.m_level;
    m_csize = option.m_csize;
    m_nccThreshold = option.m_nccThreshold;
    m_wsize = option.m_wsize;
    m_minImageNumThreshold = option.m_minImageNum;
    m_visdata = option.m_visdata;
    m_visdata2 = option.m_visdata2;
    m_tau = std::min(option.m_minImageNum * 2, m_nimages);
    m_depth = 0;
    if (m_nillums != 1) die("multi-illumination input is outside the accelerated path (m_nillums must be 1)");
    if (m_ctx) { pmk_destroy(m_ctx); m_ctx = nullptr; }
    pmk_config cfg;
    pmk_default_config(&cfg);
    cfg.device = m_device;
    cfg.nviews = m_nimages; cfg.level = m_level; cfg.csize = m_csize; cfg.wsize = m_wsize; cfg.min_image_num = m_minImageNumThreshold;
    cfg.ncc_threshold = m_nccThreshold;
    cfg.max_angle_threshold = option.m_maxAngleThreshold;
    cfg.quad_threshold = option.m_quadThreshold;
    cfg.sweep_group = m_sweepGroup;
    chk(pmk_create(&cfg, &m_ctx), "pmk_create");
    m_photoSet.init(*this, m_images, m_prefix, m_nimages, m_nillums, m_level + 3, m_wsize, 1);
    m_photoSet.setDistances();
    m_patchManager.init();
    m_dnInit.init(m_prefix, m_nimages + 1);
    m_propagate.init();
    m_optim.init();
    m_filter.init();
    pmk_thresholds t;
    chk(pmk_get_thresholds(m_ctx, &t), "pmk_get_thresholds");      // the device side derived them exactly as pmmvps.cpp:54-67
    m_angleThreshold0 = t.angle_threshold0; m_angleThreshold1 = t.angle_threshold1;
    m_countThreshold1 = 4;
    m_neighborThreshold = t.neighbor_threshold; m_neighborThreshold1 = t.neighbor_threshold1; m_neighborThreshold2 = t.neighbor_threshold2;
    m_nccThresholdBefore = t.ncc_threshold_before;
    m_maxAngleThreshold = option.m_maxAngleThreshold;
    m_quadThreshold = option.m_quadThreshold;
}

void PmMvps::syncDepth() {
    chk(pmk_set_depth(m_ctx, m_depth), "pmk_set_depth");
    chk(pmk_set_ncc_thresholds(m_ctx, m_nccThreshold, m_nccThresholdBefore), "pmk_set_ncc_thresholds");
}

void PmMvps::updateThreshold() {                                    // pmmvps.cpp:70-74
    m_nccThreshold -= 0.05f;
    m_nccThresholdBefore -= 0.05f;
    m_countThreshold1 = 2;
}

void PmMvps::run() {                                                // pmmvps.cpp:76-114
    time_t start = time(NULL);
    m_dnInit.createPatches();
    ++m_depth;
    const int ITER = 3;
    for (int iter = 0; iter < ITER; ++iter) {
        cerr << "\n---------------------" << endl << "Iteration: " << iter << endl << "---------------------" << endl;
        m_propagate.run(iter);
        cerr << "\nWriting Iter " << iter << " to file..." << endl;
        m_patchManager.writePatches(m_prefix + "ply/refined_patches_before_refine_" + std::to_string(iter), true, false, false);
        m_filter.run();
        updateThreshold();
        ++m_depth;
        cerr << "\nWriting Iter " << iter << " to file..." << endl;
        m_patchManager.writePatches(m_prefix + "ply/refined_patches_" + std::to_string(iter), true, false, false);
    }
    cerr << "---- Total: " << (time(NULL) - start) << " secs ----" << endl;
}

namespace {
void pack10(const Patch& p, float* o) {
    for (int k = 0; k < 4; ++k) { o[k] = p.m_coord(k); o[4 + k] = p.m_normal(k); }
    o[8] = p.m_dscale;
    o[9] = (float)(p.m_images.empty() ? 0 : p.m_images[0]);
}
int neighbor_probe(const PmMvps& pm, const Patch& lhs, const Patch& rhs, const float* hunit, const float* radius, float thr) {
    float l[10], r[10];
    int out = 0;
    pack10(lhs, l); pack10(rhs, r);
    chk(pmk_probe_neighbor(pm.m_ctx, 1, l, r, hunit, radius, thr, &out), "isNeighbor");
    return out;
}
}  // namespace

// the three neighbour tests run on the device (the decision arithmetic of pmmvps.cpp:117-180 in the reference's operation order)
int PmMvps::isNeighbor(const Patch& lhs, const Patch& rhs, const float neighborThreshold) const {
    return neighbor_probe(*this, lhs, rhs, nullptr, nullptr, neighborThreshold);
}

int PmMvps::isNeighbor(const Patch& lhs, const Patch& rhs, const float hunit, const float neighborThreshold) const {
    return neighbor_probe(*this, lhs, rhs, &hunit, nullptr, neighborThreshold);
}

int PmMvps::isNeighborRadius(const Patch& lhs, const Patch& rhs, const float hunit, const float neighborThreshold, const float radius) const {
    return neighbor_probe(*this, lhs, rhs, &hunit, &radius, neighborThreshold);
}
