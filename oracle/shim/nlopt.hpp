// oracle/shim/nlopt.hpp -- TEST INFRASTRUCTURE ONLY.
//
// NLopt is a third-party dependency of the reference (pmmvps/optim.cpp:12, call sites
// :511-530: LN_BOBYQA, 3 variables, xtol_rel 1e-7, maxeval 500).  It is not vendored, no
// version is pinned anywhere in the reference, and it is absent from this image, so BOBYQA's
// trajectory cannot be reproduced: PARITY UNPINNED at this boundary.  Per BASELINE.json's
// north_star the refinement is re-defined as a seeded, counter-based random search over the
// same three encoded variables, with the same bounds and the same objective (Optim::cost_func,
// which IS reference source and is pinned by the oracle).  This shim is the CPU definition of
// that schedule ("PMR1"); warp_refine / k3_refine in mvskit_b200/csrc/pmk_cand.cuh is the CUDA one.
//
// PMR1 (batch-synchronous halving random search):
//   best = clamp(x0); fbest = f(best)
//   r = {4, 4, 4}                       (x[0]: pixels of parallax, x[1..2]: units of pi/48)
//   for level in 0..11:
//       for cand in 0..7:   u = philox4x32_10(key=seed, ctr={stream_lo, stream_hi, level, cand})
//                           x_c[i] = clamp(best[i] + r[i] * U(u[i])),  U(v) = ((v>>8)+0.5)*2^-23 - 1
//       c* = argmin_c f(x_c) (lowest index on ties); if f(x_c*) < fbest: best, fbest = x_c*, f
//       r *= 0.6
//   97 objective evaluations, always "XTOL_REACHED".
#ifndef PM_ORACLE_NLOPT_SHIM
#define PM_ORACLE_NLOPT_SHIM

#include <cstdint>
#include <stdexcept>
#include <vector>

namespace pmr1 {

static const int kLevels = 12;
static const int kCands = 8;
static const double kShrink = 0.6;
static const double kRange0 = 4.0;

struct State {
    uint64_t seed;
    uint64_t stream;
    // optional trace of every evaluated point, for teacher-forced parity tests
    std::vector<double>* trace;
};
inline State& state() { static State s = {0x9E3779B97F4A7C15ull, 0, 0}; return s; }

inline void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c[4]) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
        const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
        const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = static_cast<uint32_t>(p1);
        const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = static_cast<uint32_t>(p0);
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

inline double uniform_pm1(uint32_t v) { return (static_cast<double>(v >> 8) + 0.5) * (1.0 / 8388608.0) - 1.0; }

}  // namespace pmr1

namespace nlopt {

enum algorithm { LN_BOBYQA = 34 };
enum result { FAILURE = -1, SUCCESS = 1, STOPVAL_REACHED = 2, FTOL_REACHED = 3, XTOL_REACHED = 4, MAXEVAL_REACHED = 5 };

typedef double (*func)(unsigned n, const double* x, double* grad, void* f_data);

class opt {
public:
    opt(algorithm, unsigned n) : n_(n), f_(0), fd_(0), maxeval_(0) {}
    void set_min_objective(func f, void* f_data) { f_ = f; fd_ = f_data; }
    void set_xtol_rel(double) {}
    void set_maxeval(int m) { maxeval_ = m; }
    void set_lower_bounds(const std::vector<double>& lb) { lb_ = lb; }
    void set_upper_bounds(const std::vector<double>& ub) { ub_ = ub; }

    result optimize(std::vector<double>& x, double& minf) {
        if (!f_ || x.size() != n_ || n_ != 3) throw std::invalid_argument("pmr1: bad problem");
        pmr1::State& st = pmr1::state();
        std::vector<double> best(x);
        clamp(best);
        double fbest = eval(best, st);
        double r[3] = {pmr1::kRange0, pmr1::kRange0, pmr1::kRange0};
        std::vector<double> cand(3), win(3);
        for (int level = 0; level < pmr1::kLevels; ++level) {
            double fwin = 0.0;
            int have = 0;
            for (int c = 0; c < pmr1::kCands; ++c) {
                uint32_t ctr[4] = {static_cast<uint32_t>(st.stream), static_cast<uint32_t>(st.stream >> 32),
                                   static_cast<uint32_t>(level), static_cast<uint32_t>(c)};
                pmr1::philox4x32_10(static_cast<uint32_t>(st.seed), static_cast<uint32_t>(st.seed >> 32), ctr);
                for (int i = 0; i < 3; ++i) cand[i] = best[i] + r[i] * pmr1::uniform_pm1(ctr[i]);
                clamp(cand);
                const double fc = eval(cand, st);
                if (!have || fc < fwin) { fwin = fc; win = cand; have = 1; }
            }
            if (fwin < fbest) { fbest = fwin; best = win; }
            for (int i = 0; i < 3; ++i) r[i] *= pmr1::kShrink;
        }
        x = best;
        minf = fbest;
        return XTOL_REACHED;
    }

private:
    void clamp(std::vector<double>& v) const {
        for (unsigned i = 0; i < n_; ++i) {
            if (i < ub_.size() && v[i] > ub_[i]) v[i] = ub_[i];
            if (i < lb_.size() && v[i] < lb_[i]) v[i] = lb_[i];
        }
    }
    double eval(const std::vector<double>& v, pmr1::State& st) const {
        const double f = f_(n_, v.data(), 0, fd_);
        if (st.trace) { st.trace->push_back(v[0]); st.trace->push_back(v[1]); st.trace->push_back(v[2]); st.trace->push_back(f); }
        return f;
    }
    unsigned n_;
    func f_;
    void* fd_;
    int maxeval_;
    std::vector<double> lb_, ub_;
};

}  // namespace nlopt

#endif
