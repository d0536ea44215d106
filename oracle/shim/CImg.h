// oracle/shim/CImg.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for CImg (absent from the reference tree and from this image).  The reference only
// uses CImg to decode "<prefix>image/%04d0000.jpg" (image/image.cpp:827-879).  The synthetic
// scenes store an exact binary PPM (P6) payload under that name, so decode parity is moot and
// the u8 pixels the reference sees are exactly the generator's.
#ifndef PM_ORACLE_CIMG_SHIM
#define PM_ORACLE_CIMG_SHIM

#include <cstdio>
#include <cstdlib>
#include <exception>
#include <vector>

namespace cimg_library {

struct CImgException : public std::exception {
    const char* what() const noexcept override { return "CImg shim: cannot read image"; }
};

namespace cimg {
inline const char* imagemagick_path(const char* p = 0) { return p; }
}

template <typename T>
class CImg {
public:
    CImg() : w_(0), h_(0), c_(0) {}
    bool is_empty() const { return d_.empty(); }
    int width() const { return w_; }
    int height() const { return h_; }
    int spectrum() const { return c_; }
    size_t size() const { return d_.size(); }
    // (x, y, z, c) accessor; storage here is interleaved, the caller re-packs anyway
    T operator()(int x, int y, int, int c) const { return d_[(static_cast<size_t>(y) * w_ + x) * c_ + c]; }
    CImg& load_jpeg(const char* file) { return load(file); }
    CImg& load(const char* file) {
        FILE* f = std::fopen(file, "rb");
        if (!f) throw CImgException();
        char magic[3] = {0, 0, 0};
        int maxv = 0;
        if (std::fscanf(f, "%2s", magic) != 1 || magic[0] != 'P' || (magic[1] != '6' && magic[1] != '5')) { std::fclose(f); throw CImgException(); }
        c_ = magic[1] == '6' ? 3 : 1;
        if (!read_int(f, w_) || !read_int(f, h_) || !read_int(f, maxv) || maxv != 255) { std::fclose(f); throw CImgException(); }
        std::fgetc(f);  // single whitespace after maxval
        d_.resize(static_cast<size_t>(w_) * h_ * c_);
        const size_t got = std::fread(d_.data(), 1, d_.size(), f);
        std::fclose(f);
        if (got != d_.size()) { d_.clear(); throw CImgException(); }
        return *this;
    }
private:
    static bool read_int(FILE* f, int& v) {
        int ch = std::fgetc(f);
        for (;;) {
            while (ch == ' ' || ch == '\n' || ch == '\r' || ch == '\t') ch = std::fgetc(f);
            if (ch == '#') { while (ch != '\n' && ch != EOF) ch = std::fgetc(f); continue; }
            break;
        }
        if (ch < '0' || ch > '9') return false;
        v = 0;
        while (ch >= '0' && ch <= '9') { v = v * 10 + (ch - '0'); ch = std::fgetc(f); }
        std::ungetc(ch, f);
        return true;
    }
    int w_, h_, c_;
    std::vector<T> d_;
};

}  // namespace cimg_library

#endif
