// CPU thread emulation of the v2 Fourier kernels (atlas_b200/csrc/fft2_core.cuh): the exact kernel bodies run with
// one OS thread per CUDA thread (std::barrier for __syncthreads, one barrier per half-warp for __syncwarp) and are
// compared with a naive O(n L) DFT.  Covers every block-level radix M1, the half-warp-local 256-point transforms,
// the filter-table construction, the cp.async staging of the next field and the field loop.
// Built and run by tests/test_fft_core_cpu.py (g++ -std=c++20 -pthread).
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include "../../atlas_b200/csrc/fft2_core.cuh"

namespace sptrans {
namespace emu {
static std::unique_ptr<std::barrier<>> g_block;
static std::vector<std::unique_ptr<std::barrier<>>> g_half;
void sync_block() { g_block->arrive_and_wait(); }
void sync_halfwarp(int tid) { g_half[tid >> 4]->arrive_and_wait(); }
static void run_block(int nt, const std::function<void(int)>& body) {
    g_block = std::make_unique<std::barrier<>>(nt);
    g_half.clear();
    for (int h = 0; h < nt / 16; ++h) g_half.push_back(std::make_unique<std::barrier<>>(16));
    std::vector<std::thread> th;
    th.reserve(nt);
    for (int t = 0; t < nt; ++t) th.emplace_back(body, t);
    for (auto& x : th) x.join();
}
}  // namespace emu
}  // namespace sptrans

using namespace sptrans;
using namespace sptrans::fft2;

static double urand() { return rand() / (double)RAND_MAX - 0.5; }

template <int R>
static double check_radix() {  // register butterflies against the DFT definition, both signs
    double worst = 0;
    for (int fwd = 0; fwd < 2; ++fwd) {
        double2 v[R], x[R];
        for (int j = 0; j < R; ++j) x[j] = v[j] = make_double2(urand(), urand());
        if (fwd) dftX<R, true>(v);
        else dftX<R, false>(v);
        for (int k = 0; k < R; ++k) {
            double re = 0, im = 0;
            for (int j = 0; j < R; ++j) {
                const double ang = (fwd ? -2 : 2) * M_PI * ((j * k) % R) / R;
                re += x[j].x * std::cos(ang) - x[j].y * std::sin(ang);
                im += x[j].x * std::sin(ang) + x[j].y * std::cos(ang);
            }
            worst = std::fmax(worst, std::hypot(v[k].x - re, v[k].y - im));
        }
    }
    return worst;
}

struct Case {
    int n, L, nf, F, has_s, mlimit, adjoint, nb_uv;
};

template <int M1, int NT>
static int run_case(const Case& c) {
    const int n = c.n, L = c.L, nf = c.nf, M = M1 * kM2;
    int m1 = 0;
    if (conv_length_v2(n + 2 * L, &m1) != M || m1 != M1) {
        std::printf("case n=%d L=%d does not select M1=%d (got %d)\n", n, L, M1, m1);
        return 1;
    }
    // geometry: one latitude pair (index 0), every m has a sym row and an asym row
    std::vector<int> nlat0(L + 2, 0);
    std::vector<long long> fb_rowoff(L + 2);
    for (int m = 0; m <= L + 1; ++m) fb_rowoff[m] = 2LL * m;
    PairMeta pm{};
    pm.n = n; pm.L = L; pm.M = M; pm.m1 = M1; pm.has_s = c.has_s; pm.F = c.F;
    pm.rowN = 0; pm.rowS = n;  // grid: [field][2 rows]
    pm.chirp_off = 0; pm.filt_off = 0; pm.tw_off = 0;
    const long long npts = 2LL * n;
    std::vector<double2> chirp(2 * L + 1 + n), W1(kM2), T(256), filt(M);
    for (int u = 0; u <= 2 * L; ++u) {
        const double ang = M_PI * (double)chirp_residue(u, 0, n) / n;
        chirp[u] = make_double2(std::cos(ang), std::sin(ang));
    }
    for (int i = 0; i < n; ++i) {
        const double ang = M_PI * (double)chirp_residue(i, -2LL * L, n) / n;
        chirp[2 * L + 1 + i] = make_double2(std::cos(ang), std::sin(ang));
    }
    for (int t = 0; t < kM2; ++t) W1[t] = make_double2(std::cos(-2 * M_PI * t / M), std::sin(-2 * M_PI * t / M));
    for (int q = 0; q < 16; ++q)
        for (int l = 0; l < 16; ++l) T[q * 16 + l] = make_double2(std::cos(-2 * M_PI * q * l / 256), std::sin(-2 * M_PI * q * l / 256));
    const size_t smem = (size_t)M + 256 + 4 * (L + 1) + n + (L + 1) + 16;  // double2 units: X, then the larger of the two staging areas
    std::vector<double2> X(smem);
    emu::run_block(NT, [&](int tid) { filter_table_body<M1, NT>(pm, tid, X.data(), W1.data(), T.data(), filt.data()); });

    std::vector<double> scale(1, 1.7), weights(1, 0.37);
    std::vector<double2> fb((size_t)2 * (L + 1) * nf);
    std::vector<double> gp((size_t)nf * npts, 0.);
    Fft2Args a{};
    a.meta = &pm; a.nf = nf; a.F = c.F; a.mlimit = c.mlimit; a.nb_uv = c.nb_uv; a.fb_rowoff = fb_rowoff.data(); a.nlat0 = nlat0.data();
    a.nleg = 1; a.twid = W1.data(); a.t256 = T.data(); a.chirp = chirp.data(); a.filt = filt.data();
    a.scale_lat = scale.data(); a.weights = weights.data(); a.fb = fb.data(); a.gp = gp.data(); a.npts = npts;
    a.adjoint = c.adjoint; a.gp_aligned16 = 1;
    std::vector<double> cs_(n), sn_(n);
    for (int r = 0; r < n; ++r) { cs_[r] = std::cos(2 * M_PI * r / n); sn_[r] = std::sin(2 * M_PI * r / n); }

    // ---- inverse ----
    for (auto& v : fb) v = make_double2(urand(), urand());
    for (int f0 = 0; f0 < nf; f0 += c.F)
        emu::run_block(NT, [&](int tid) { fourier2_inv_body<M1, NT>(a, 0, f0, std::min(c.F, nf - f0), tid, X.data()); });
    const int Lc = std::min(L, c.mlimit);
    double err_inv = 0, nrm_inv = 0;
    for (int f = 0; f < nf; ++f) {
        std::vector<double2> FN(Lc + 1), FS(Lc + 1);
        for (int m = 0; m <= Lc; ++m) {
            double2 s = fb[(size_t)(2 * m) * nf + f], as = fb[(size_t)(2 * m + 1) * nf + f];
            if (m == 0) s.y = as.y = 0;
            if (c.has_s) { FN[m] = cadd(s, as); FS[m] = csub(s, as); }
            else { FN[m] = csub(s, as); FS[m] = make_double2(0, 0); }
        }
        const double sc = f < c.nb_uv ? scale[0] : 1.0;
        for (int i = 0; i < n; ++i) {
            double xn = FN[0].x, xs = FS[0].x;
            for (int m = 1; m <= Lc; ++m) {
                const int r = (int)(((long long)m * i) % n);
                xn += 2 * (FN[m].x * cs_[r] - FN[m].y * sn_[r]);
                xs += 2 * (FS[m].x * cs_[r] - FS[m].y * sn_[r]);
            }
            err_inv = std::fmax(err_inv, std::fabs(gp[f * npts + i] - xn * sc));
            if (c.has_s) err_inv = std::fmax(err_inv, std::fabs(gp[f * npts + n + i] - xs * sc));
            nrm_inv = std::fmax(nrm_inv, std::fabs(xn * sc));
        }
    }
    // ---- direct ----
    for (auto& v : gp) v = urand();
    for (auto& v : fb) v = make_double2(1e30, 1e30);
    for (int f0 = 0; f0 < nf; f0 += c.F)
        emu::run_block(NT, [&](int tid) { fourier2_dir_body<M1, NT>(a, 0, f0, std::min(c.F, nf - f0), tid, X.data()); });
    double err_dir = 0, nrm_dir = 0;
    for (int f = 0; f < nf; ++f) {
        const double sc = f < c.nb_uv ? scale[0] : 1.0;
        for (int m = 0; m <= L; ++m) {
            double2 fn = make_double2(0, 0), fs = make_double2(0, 0);
            for (int i = 0; i < n; ++i) {
                const int r = (int)(((long long)m * i) % n);
                const double xn = gp[f * npts + i] * sc, xs = c.has_s ? gp[f * npts + n + i] * sc : 0.;
                fn.x += xn * cs_[r]; fn.y -= xn * sn_[r];
                fs.x += xs * cs_[r]; fs.y -= xs * sn_[r];
            }
            // adjoint of the inverse w.r.t. the spectral inner product that counts m > 0 twice: every m alike
            double w = c.adjoint ? 1.0 : weights[0] / n;
            double2 s, as;
            if (c.has_s) { s = make_double2((fn.x + fs.x) * w, (fn.y + fs.y) * w); as = make_double2((fn.x - fs.x) * w, (fn.y - fs.y) * w); }
            else { s = make_double2(fn.x * w, fn.y * w); as = s; }
            const double2 gs = fb[(size_t)(2 * m) * nf + f], ga = fb[(size_t)(2 * m + 1) * nf + f];
            err_dir = std::fmax(err_dir, std::hypot(gs.x - s.x, gs.y - s.y));
            err_dir = std::fmax(err_dir, std::hypot(ga.x - as.x, ga.y - as.y));
            nrm_dir = std::fmax(nrm_dir, std::hypot(s.x, s.y));
        }
    }
    const bool ok = err_inv < 2e-12 * nrm_inv && err_dir < 2e-12 * nrm_dir;
    std::printf("M1=%2d NT=%3d n=%4d L=%4d nf=%d F=%d has_s=%d mlimit=%d adj=%d: inv %.2e (max %.2e)  dir %.2e (max %.2e) %s\n",
                M1, NT, n, L, nf, c.F, c.has_s, c.mlimit, c.adjoint, err_inv, nrm_inv, err_dir, nrm_dir, ok ? "ok" : "FAIL");
    return ok ? 0 : 1;
}

int main() {
    int bad = 0;
    double e;
#define RADIX(R)                                                  \
    e = check_radix<R>();                                         \
    std::printf("radix %2d butterfly: max err %.2e\n", R, e);     \
    bad += e > 1e-13;
    RADIX(6) RADIX(8) RADIX(9) RADIX(10) RADIX(12) RADIX(15) RADIX(16) RADIX(18) RADIX(20) RADIX(24) RADIX(25) RADIX(27)
    RADIX(30) RADIX(32)
    // (n, L) chosen so that n + 2L falls in the window of each radix; n % 4 == 0 like the octahedral rows
    bad += run_case<8, 256>({1200, 400, 3, 2, 1, 400, 0, 1});
    bad += run_case<9, 256>({1372, 450, 2, 2, 1, 300, 0, 0});
    bad += run_case<10, 256>({1500, 500, 2, 1, 1, 500, 0, 0});
    bad += run_case<12, 128>({1800, 600, 3, 3, 0, 600, 0, 0});
    bad += run_case<15, 256>({2400, 700, 2, 2, 1, 700, 1, 0});
    bad += run_case<16, 256>({2500, 790, 1, 1, 1, 790, 0, 0});
    bad += run_case<18, 256>({2800, 880, 2, 2, 1, 880, 0, 2});
    bad += run_case<20, 256>({3000, 1040, 1, 1, 1, 1040, 0, 0});
    bad += run_case<24, 256>({3584, 1279, 2, 2, 1, 1278, 0, 0});
    bad += run_case<25, 256>({3840, 1279, 1, 1, 1, 1279, 0, 0});
    bad += run_case<27, 256>({4352, 1279, 1, 1, 1, 1279, 0, 0});
    bad += run_case<30, 256>({5120, 1279, 3, 3, 1, 1279, 0, 0});
    bad += run_case<32, 256>({5136, 1500, 1, 1, 1, 1500, 0, 0});
    bad += run_case<30, 128>({5000, 1279, 2, 2, 1, 1279, 0, 0});
    bad += run_case<24, 256>({5000, 500, 2, 2, 1, 500, 0, 0});  // long rows, low truncation: n > 0.7 M (late chirp loads)
    bad += run_case<16, 128>({3700, 190, 2, 2, 1, 190, 0, 0});
    std::printf(bad ? "FAILED (%d)\n" : "ALL OK\n", bad);
    return bad ? 1 : 0;
}
