// mvskit_b200/host/main.cpp -- the reference's driver (test/test.cpp:155-161) on the B200 path:
//     pmmvps_b200 <prefix/> [option-file] [--device N] [--group G] [--filter-only ITER] [--selftest]
// `--filter-only ITER` is test/test_filter.cpp: m_depth = 1, readPatches(ITER), Filter::run.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <iostream>
#include <vector>

#include "pmmvps.hpp"

// --patch-io <in.patch> <out.patch>: parse a .patch file with the mirror's stream operators and write it back (no GPU needed);
// the CPU test suite compares the records with what the reference itself reads from the same file.
static int patch_io(const char* in_name, const char* out_name) {
    std::ifstream in(in_name);
    if (!in.is_open()) { std::cerr << "cannot open " << in_name << std::endl; return 1; }
    std::string header;
    int pnum = 0;
    in >> header >> pnum;
    std::ofstream out(out_name);
    out << "PATCHES" << std::endl << pnum << std::endl;
    for (int p = 0; p < pnum; ++p) {
        Patch patch;
        in >> patch;
        out << patch << "\n";
    }
    return in.fail() ? 1 : 0;
}

// --mask-io <prefix/mask/%08d> <out.pgm>: read a mask with the mirror's reader and write the grey values back as P5 (no GPU needed)
static int mask_io(const char* base, const char* out_name) {
    std::vector<unsigned char> grey;
    int w = 0, h = 0;
    if (!PhotoSet::readMask(base, grey, w, h)) { std::cerr << "no mask at " << base << std::endl; return 1; }
    std::ofstream out(out_name, std::ios::binary);
    out << "P5\n" << w << " " << h << "\n255\n";
    out.write((const char*)grey.data(), (std::streamsize)grey.size());
    return 0;
}

// --selftest: the PatchManager pass-throughs on the loaded seeds (needs a GPU): one-patch computeNcc (wide call) against the batched
// computeNcc (byte-lean call) bit for bit, sortPatches on unscored patches, setScales / isVisible0 / findNeighbors / removePatch (by collect
// index, on a store that is NOT in collect order) / updateDepthMaps / syncGrids for internal consistency.  Exit code 0 and "selftest ok ..." on success.
static int selftest(PmMvps& pmmvps) {
    PatchManager& pm = pmmvps.m_patchManager;
    pm.readPatches();
    pmmvps.m_depth = 1;
    pm.collectPatches();
    const int total = (int)pm.m_ppatches.size();
    const int n = std::min(total, 512);
    if (n < 8) { std::cerr << "selftest: too few seeds (" << total << ")" << std::endl; return 1; }
    std::vector<Ppatch> batch;
    std::vector<float> single(n);
    for (int i = 0; i < n; ++i) {
        Ppatch pp(new Patch(*pm.m_ppatches[i]));
        pp->m_ncc = -1.0f;
        pm.computeNcc(*pp);
        single[i] = pp->m_ncc;
        pp->m_ncc = -1.0f;
        batch.push_back(pp);
    }
    pm.computeNcc(batch);
    int bad = 0, scored = 0;
    for (int i = 0; i < n; ++i) {
        if (std::memcmp(&single[i], &batch[i]->m_ncc, sizeof(float)) != 0) ++bad;
        if (batch[i]->m_ncc > 0.0f) ++scored;
    }
    if (bad || scored < n / 2) { std::cerr << "selftest: batched computeNcc differs on " << bad << " of " << n << " patches, " << scored << " scored" << std::endl; return 1; }
    for (int i = 0; i < n; ++i) batch[i]->m_ncc = -1.0f;
    pm.sortPatches(batch, 0);
    for (int i = 1; i < n; ++i)
        if (batch[i - 1]->m_ncc < batch[i]->m_ncc) { std::cerr << "selftest: sortPatches order broken at " << i << std::endl; return 1; }
    // setScales reproduces the stored scales; every patch is visible in its own reference view at its own cell
    int vis = 0, scales = 0, withnb = 0;
    for (int i = 0; i < 64; ++i) {
        Patch p(*pm.m_ppatches[i]);
        p.m_dscale = p.m_ascale = 0.0f;
        pm.setScales(p);
        if (p.m_dscale > 0.0f && p.m_ascale > 0.0f) ++scales;
        int ix = -1, iy = -1;
        if (pm.isVisible0(p, p.m_images[0], ix, iy, 0.5f) && ix == p.m_grids[0](0) && iy == p.m_grids[0](1)) ++vis;
        std::vector<Ppatch> nb;
        pm.findNeighbors(p, nb, 4.0f, 2, 0);
        if (!nb.empty()) ++withnb;
    }
    if (vis < 60 || scales < 60) { std::cerr << "selftest: isVisible0 " << vis << " / 64, setScales " << scales << " / 64" << std::endl; return 1; }
    // removePatch takes exactly one patch out of the store and out of the grids
    pm.syncGrids();
    size_t reg0 = 0;
    for (const auto& view : pm.m_pgrids) for (const auto& cell : view) reg0 += cell.size();
    Ppatch victim = pm.m_ppatches[n / 2];
    const size_t nreg = victim->m_images.size();
    pm.removePatch(victim);
    pm.syncGrids();
    size_t reg1 = 0;
    for (const auto& view : pm.m_pgrids) for (const auto& cell : view) reg1 += cell.size();
    if ((int)pm.m_ppatches.size() != total - 1 || reg0 - reg1 != nreg) {
        std::cerr << "selftest: removePatch left " << pm.m_ppatches.size() << " of " << total << " patches, " << (reg0 - reg1) << " registrations gone, expected " << nreg << std::endl;
        return 1;
    }
    pm.updateDepthMaps(pm.m_ppatches[0]);
    std::cout << "selftest ok: " << n << " patches scored identically by the wide and the byte-lean call (" << scored << " valid), sorted; isVisible0 " << vis
              << " / 64, setScales " << scales << " / 64, findNeighbors non-empty for " << withnb << " / 64; removePatch took " << nreg << " registrations" << std::endl;
    return 0;
}

int main(int argc, char* argv[]) {
    if (argc == 4 && !strcmp(argv[1], "--patch-io")) return patch_io(argv[2], argv[3]);
    if (argc == 4 && !strcmp(argv[1], "--mask-io")) return mask_io(argv[2], argv[3]);
    if (argc < 2) {
        std::cerr << "usage: " << argv[0] << " <prefix/> [option] [--device N] [--group G] [--filter-only ITER]" << std::endl;
        return 2;
    }
    std::string prefix = argv[1], optname = "option";
    int device = 0, group = 1, filter_only = -1, self = 0;
    for (int i = 2; i < argc; ++i) {
        if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--group") && i + 1 < argc) group = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--filter-only") && i + 1 < argc) filter_only = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--selftest")) self = 1;
        else optname = argv[i];
    }
    if (prefix.empty() || prefix[prefix.size() - 1] != '/') prefix += "/";

    Option option;
    option.init(prefix, optname);

    PmMvps pmmvps;
    pmmvps.m_device = device;
    pmmvps.m_sweepGroup = group;
    pmmvps.init(option);
    if (self) return selftest(pmmvps);

    if (filter_only >= 0) {
        pmmvps.m_depth = 1;
        pmmvps.m_patchManager.readPatches(filter_only);
        pmmvps.m_filter.run();
        pmmvps.m_patchManager.writePatches(prefix + "ply/filtered", true, true, false);
    } else {
        pmmvps.run();
        pmmvps.m_patchManager.writePatches(prefix + "ply/final", false, true, false);
    }
    pmmvps.m_patchManager.collectPatches();
    std::cout << "patches " << pmmvps.m_patchManager.m_ppatches.size() << std::endl;
    return 0;
}
