// mvskit_b200/host/main.cpp -- the reference's driver (test/test.cpp:155-161) on the B200 path:
//     pmmvps_b200 <prefix/> [option-file] [--device N] [--group G] [--filter-only ITER]
// `--filter-only ITER` is test/test_filter.cpp: m_depth = 1, readPatches(ITER), Filter::run.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>

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

int main(int argc, char* argv[]) {
    if (argc == 4 && !strcmp(argv[1], "--patch-io")) return patch_io(argv[2], argv[3]);
    if (argc == 4 && !strcmp(argv[1], "--mask-io")) return mask_io(argv[2], argv[3]);
    if (argc < 2) {
        std::cerr << "usage: " << argv[0] << " <prefix/> [option] [--device N] [--group G] [--filter-only ITER]" << std::endl;
        return 2;
    }
    std::string prefix = argv[1], optname = "option";
    int device = 0, group = 1, filter_only = -1;
    for (int i = 2; i < argc; ++i) {
        if (!strcmp(argv[i], "--device") && i + 1 < argc) device = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--group") && i + 1 < argc) group = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--filter-only") && i + 1 < argc) filter_only = atoi(argv[++i]);
        else optname = argv[i];
    }
    if (prefix.empty() || prefix[prefix.size() - 1] != '/') prefix += "/";

    Option option;
    option.init(prefix, optname);

    PmMvps pmmvps;
    pmmvps.m_device = device;
    pmmvps.m_sweepGroup = group;
    pmmvps.init(option);

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
