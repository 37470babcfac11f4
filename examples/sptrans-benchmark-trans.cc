// sptrans-benchmark-trans -- the harness shape of ecmwf/atlas src/sandbox/benchmark_trans/atlas-benchmark-trans.cc
// (:250-288: constructor timed separately, niter timed invtrans calls with host buffers, per-iteration / min / max lines)
// over the C ABI of the B200 engine (include/sptrans_b200.h), so that the two programs' outputs can be laid side by side:
//
//     atlas-benchmark-trans   --grid O1280 --type local --nscalar 137 --niter 5
//     sptrans-benchmark-trans --grid O1280              --nscalar 137 --niter 5 [--nvordiv K] [--dirtrans]
//
// Build:  g++ -O2 -std=c++17 -I include examples/sptrans-benchmark-trans.cc -L atlas_b200 -lsptrans_b200 \
//             -Wl,-rpath,$PWD/atlas_b200 -o sptrans-benchmark-trans
// Without a CUDA device the plan constructor refuses (this engine has no CPU fallback) and the program exits with 2.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <random>
#include <string>
#include <vector>

#include "sptrans_b200.h"

namespace {
double seconds_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
int fail(const char* what) {
    std::fprintf(stderr, "%s: %s\n", what, sptrans_last_error());
    return 2;
}
}  // namespace

int main(int argc, char** argv) {
    std::string gridname = "O32";
    int truncation = -1, nscalar = 1, nvordiv = 0, niter = 5, device = 0;
    bool dirtrans = false;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "--grid") gridname = next();
        else if (a == "--truncation") truncation = std::atoi(next());
        else if (a == "--nscalar") nscalar = std::atoi(next());
        else if (a == "--nvordiv") nvordiv = std::atoi(next());
        else if (a == "--niter") niter = std::atoi(next());
        else if (a == "--device") device = std::atoi(next());
        else if (a == "--dirtrans") dirtrans = true;
        else {
            std::fprintf(stderr, "usage: %s [--grid O<N>|F<N>] [--truncation T] [--nscalar n] [--nvordiv n] [--niter n] "
                                 "[--device d] [--dirtrans]\n", argv[0]);
            return 1;
        }
    }
    if (gridname.size() < 2 || (gridname[0] != 'O' && gridname[0] != 'F')) {
        std::fprintf(stderr, "grid must be O<N> (octahedral) or F<N> (regular Gaussian)\n");
        return 1;
    }
    const int N = std::atoi(gridname.c_str() + 1);
    if (N < 1) return 1;
    if (truncation < 0) truncation = N - 1;  // cubic relation, atlas-benchmark-trans.cc:110-114
    const int nlat = 2 * N;
    std::vector<int> nx(nlat);
    std::vector<double> lat(nlat), w(nlat);
    if (sptrans_gaussian_latitudes(N, lat.data(), w.data())) return fail("gaussian latitudes");
    if (gridname[0] == 'O') {
        if (sptrans_octahedral_nx(N, nx.data())) return fail("octahedral grid");
    }
    else std::fill(nx.begin(), nx.end(), 4 * N);

    auto c0 = std::chrono::steady_clock::now();
    sptrans_plan* plan = nullptr;
    if (sptrans_plan_create(&plan, nlat, nx.data(), lat.data(), w.data(), truncation,
                            gridname[0] == 'F' ? SPTRANS_GRID_REGULAR : 0u, device))
        return fail("constructor");
    std::printf("type=b200                        constructor:   %g s\n", seconds_since(c0));

    const size_t nspec2 = sptrans_nb_spectral_coefficients(plan), npts = sptrans_nb_gridpoints(plan);
    std::mt19937_64 rng(20260925);
    std::normal_distribution<double> gauss;
    auto spectra = [&](int nf) {
        std::vector<double> sp(nspec2 * std::max(nf, 0));
        size_t c = 0;
        for (int m = 0; m <= truncation; ++m)
            for (int n = m; n <= truncation; ++n, ++c)
                for (int imag = 0; imag < 2; ++imag)
                    for (int f = 0; f < nf; ++f)
                        sp[(2 * c + imag) * nf + f] = (m == 0 && imag) ? 0. : gauss(rng) * std::pow(1. + n, -1.5);
        return sp;
    };
    std::vector<double> sp_scalar = spectra(nscalar), sp_vor = spectra(nvordiv), sp_div = spectra(nvordiv);
    std::vector<double> gp(npts * (static_cast<size_t>(nscalar) + 2 * nvordiv));

    auto run = [&](const char* what, auto&& call) -> int {
        double tmin = std::numeric_limits<double>::max(), tmax = 0.;
        for (int n = 0; n < niter; ++n) {
            auto t0 = std::chrono::steady_clock::now();
            if (call()) return fail(what);
            const double s = seconds_since(t0);
            std::printf("type=b200      %s[%03d]: %g s\n", what, n, s);
            tmin = std::min(tmin, s);
            tmax = std::max(tmax, s);
        }
        std::printf("type=b200      %s[min]: %g s\n", what, tmin);
        std::printf("type=b200      %s[max]: %g s\n", what, tmax);
        return 0;
    };
    int rc = run("invtrans", [&] {
        return sptrans_invtrans(plan, nscalar, sp_scalar.data(), nvordiv, sp_vor.data(), sp_div.data(), gp.data());
    });
    if (!rc && dirtrans && nscalar > 0) {
        std::vector<double> back(sp_scalar.size());
        const double* gp_scalar = gp.data() + npts * 2 * static_cast<size_t>(nvordiv);
        rc = run("dirtrans", [&] { return sptrans_dirtrans_scalar(plan, nscalar, gp_scalar, back.data()); });
        double err = 0.;
        for (size_t i = 0; i + 2 * static_cast<size_t>(nscalar) < back.size(); ++i)  // all but the m == T coefficient
            err = std::max(err, std::fabs(back[i] - sp_scalar[i]));
        std::printf("type=b200      round trip max |diff| (m < T): %g\n", err);
    }
    float ms[8];
    if (!rc && !sptrans_last_timings(plan, ms))
        std::printf("type=b200      last call on the device: pack %.3f  Legendre %.3f  Fourier %.3f  H2D %.3f  D2H %.3f ms\n", ms[0],
                    ms[1], ms[2], ms[3], ms[4]);
    sptrans_plan_destroy(plan);
    return rc;
}
