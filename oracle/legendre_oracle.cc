// =====================================================================================
// TEST INFRASTRUCTURE ONLY.  CPU restatement ("oracle") of the ecmwf/atlas TransLocal
// spectral transform.  Nothing in the product library (atlas_b200/csrc, include/) may
// include, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
//
// Parity status:
//   * Legendre polynomials (orc_legendre_*)  : PINNED bit-for-bit against the unmodified
//     reference source compiled in place (oracle/_ref/libref_legendre.so, see Makefile).
//   * inverse transform (orc_invtrans*)      : PINNED against the reference's own
//     closed-form harmonic tests (src/tests/trans/test_transgeneral.cc:80-374, tolerances
//     1e-13 scalar / 2e-6 wind) -- the reference cannot be built here (needs eckit/ecbuild).
//   * Gaussian latitudes                      : PINNED against the reference's 12-decimal
//     tables (tests/golden/gaussian_latitudes_N*.txt).
//   * dirtrans / invtrans_grad / uv->vordiv   : "parity unpinned" -- NotImplemented in
//     TransLocal (TransLocal.cc:848-857,1599-1685); defined here as the exact quadrature
//     adjoint of the inverse and validated by round trip / adjoint identity only.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/src/atlas).  Third-party arithmetic the reference delegates to
// (eckit::linalg gemm, FFTW/pocketfft c2r) is restated from its published semantics:
// plain column-major C=A*B, and the unnormalised backward c2r DFT
//   out[j] = X0 + sum_{k>=1} 2 Re(X_k exp(+2 pi i j k / n))       (linalg/fft/FFTW.cc:38-62).
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle_internal.h"


// This translation unit holds the parts whose results are pinned BIT-FOR-BIT against the
// reference (Legendre polynomials, Gaussian quadrature).  It is compiled with
// -ffp-contract=off and no -march flags (see Makefile), like a default build of the reference.
namespace orc {
constexpr double kRad2Deg = 180. * M_1_PI;

// -------------------------------------------------------------------------------------
// Gaussian latitudes and quadrature weights.
// grid/detail/spacing/gaussian/Latitudes.cc:94-166 (Newton + weight), :227-274 (driver).
// -------------------------------------------------------------------------------------
void gaussian_quadrature_npole_equator(int N, double* lats_deg, double* weights) {
    const size_t kdgl = 2 * static_cast<size_t>(N);
    std::vector<double> zzfn(N + 1);
    {
        std::vector<double> zfn(kdgl + 1);
        zfn[0] = 2.;  // IFS normalisation, Latitudes.cc:236
        for (size_t jn = 1; jn <= kdgl; ++jn) {
            zfn[jn] = 2.;
            for (size_t jgl = 1; jgl <= jn; ++jgl) zfn[jn] *= std::sqrt(1. - 0.25 / static_cast<double>(jgl * jgl));
            size_t iodd = jn % 2;
            for (size_t jgl = 2; jgl <= jn - iodd; jgl += 2) {
                zfn[jn - jgl] = zfn[jn - jgl + 2] * static_cast<double>((jgl - 1) * (2 * jn - jgl + 2)) /
                                static_cast<double>(jgl * (2 * jn - jgl + 1));
            }
        }
        size_t iodd = kdgl % 2;
        for (size_t jgl = iodd, ik = iodd; jgl <= kdgl; jgl += 2, ++ik) zzfn[ik] = zfn[jgl];
    }
    const double* pfn = zzfn.data();
    const size_t kn = kdgl;
    const size_t kodd = kn % 2;
    auto newton = [&](double x, double& xn, double& mod) {  // Latitudes.cc:94-137
        double zdlk = (kodd == 0) ? 0.5 * pfn[0] : 0.;
        double zdlldn = 0.;
        size_t ik = 1;
        for (size_t jn = 2 - kodd; jn <= kn; jn += 2, ++ik) {
            zdlk += pfn[ik] * std::cos(static_cast<double>(jn) * x);
            zdlldn -= pfn[ik] * static_cast<double>(jn) * std::sin(static_cast<double>(jn) * x);
        }
        mod = (zdlldn != 0) ? -zdlk / zdlldn : 0.;
        xn = x + mod;
    };
    auto weight = [&](double x) {  // Latitudes.cc:139-166
        double zdlldn = 0.;
        size_t ik = 1;
        for (size_t jn = 2 - kodd; jn <= kn; jn += 2, ++ik)
            zdlldn -= pfn[ik] * static_cast<double>(jn) * std::sin(static_cast<double>(jn) * x);
        return static_cast<double>(2 * kn + 1) / (zdlldn * zdlldn);
    };
    const double ztol = std::numeric_limits<double>::epsilon() * 1000.;
    for (int jgl = 0; jgl < N; ++jgl) {
        double z = (4. * (jgl + 1.) - 1.) * M_PI / (4. * 2. * N + 2.);  // Latitudes.cc:258
        double zx = z + 1. / (std::tan(z) * (8. * (2. * N) * (2. * N)));
        double zxn = zx, mod = 0, zw = 0;
        bool tol_reached = false;
        for (int it = 1; it <= 21; ++it) {  // Latitudes.cc:198-209
            newton(zx, zxn, mod);
            zx = zxn;
            if (tol_reached) {
                zw = weight(zx);
                break;
            }
            if (std::abs(mod) <= ztol) tol_reached = true;
        }
        lats_deg[jgl] = 90. - zxn * kRad2Deg;
        weights[jgl] = zw;
    }
}

// -------------------------------------------------------------------------------------
// Legendre polynomials, restated from trans/local/LegendrePolynomials.cc.
// Compiled WITHOUT fp contraction (see Makefile) so results are bit-identical to the
// reference source compiled in place (oracle/_ref); tests/test_oracle_legendre.py checks.
// -------------------------------------------------------------------------------------
inline size_t idx_zfn(int trc, int jn, int jk) { return static_cast<size_t>(jk) + static_cast<size_t>(trc + 1) * jn; }
inline size_t idx_mn(int trc, int jm, int jn) {
    return static_cast<size_t>(2 * trc + 3 - jm) * jm / 2 + jn - jm;
}

// LegendrePolynomials.cc:24-45
void legendre_zfn(int trc, double* zfn) {
    zfn[idx_zfn(trc, 0, 0)] = 2.;
    for (int jn = 1; jn <= trc; ++jn) {
        double zfnn = zfn[idx_zfn(trc, 0, 0)];
        for (int jgl = 1; jgl <= jn; ++jgl) zfnn *= std::sqrt(1. - 0.25 / (jgl * jgl));
        int iodd = jn % 2;
        zfn[idx_zfn(trc, jn, jn)] = zfnn;
        for (int jgl = 2; jgl <= jn - iodd; jgl += 2) {
            double num = ((jgl - 1.) * (2. * jn - jgl + 2.));
            double den = (jgl * (2. * jn - jgl + 1.));
            zfn[idx_zfn(trc, jn, jn - jgl)] = zfn[idx_zfn(trc, jn, jn - jgl + 2)] * num / den;
        }
    }
}

// LegendrePolynomials.cc:47-151 (one latitude, all m<=n<=trc)
void legendre_lat(int trc, double lat, double* legpol, double* zfn, std::vector<double>& vsin,
                  std::vector<double>& vcos) {
    double theta = (M_PI_2 - lat);
    double costh = std::cos(theta);
    volatile double sinth = std::sqrt(1. - costh * costh);  // :61 (as in ectrans)
    legpol[idx_mn(trc, 0, 0)] = 1.;
    vsin.resize(trc + 1);
    vcos.resize(trc + 1);
    for (int j = 1; j <= trc; j++) {
        vsin[j] = std::sin(j * theta);
        vcos[j] = std::cos(j * theta);
    }
    double inv_sinth = 0.;
    if (std::abs(sinth) <= std::sqrt(std::numeric_limits<double>::epsilon())) {  // :71-74
        costh = 1.;
        sinth = 0.;
    }
    else {
        inv_sinth = 1. / sinth;
    }
    // m = 0 and m = 1 columns from the cosine / sine series, :85-115
    for (int parity = 0; parity < 2; ++parity) {
        // the reference does even n first (jn=2,4,..) then odd n (1,3,..); columns are independent
        for (int jn = (parity == 0 ? 2 : 1); jn <= trc; jn += 2) {
            double zdlk, zdlldn = 0.0;
            if (parity == 0) {
                zdlk = 0.5 * zfn[idx_zfn(trc, jn, 0)];
            }
            else {
                zfn[idx_zfn(trc, jn, 0)] = 0.;
                zdlk = 0.;
            }
            double zdsq = 1. / std::sqrt(jn * (jn + 1.));
            for (int jk = (parity == 0 ? 2 : 1); jk <= jn; jk += 2) {
                zdlk = zdlk + zfn[idx_zfn(trc, jn, jk)] * vcos[jk];
                zdlldn = zdlldn + zdsq * zfn[idx_zfn(trc, jn, jk)] * jk * vsin[jk];
            }
            legpol[idx_mn(trc, 0, jn)] = zdlk;
            legpol[idx_mn(trc, 1, jn)] = zdlldn;
        }
    }
    // diagonal, :122-130, with underflow flush
    double flush = inv_sinth * std::numeric_limits<double>::min();
    for (int jn = 2; jn <= trc; ++jn) {
        double sq = std::sqrt((2. * jn + 1.) / (2. * jn));
        legpol[idx_mn(trc, jn, jn)] = legpol[idx_mn(trc, jn - 1, jn - 1)] * sinth * sq;
        if (std::abs(legpol[idx_mn(trc, jn, jn)]) < flush) legpol[idx_mn(trc, jn, jn)] = 0.0;
    }
    // four-point recurrence (Belousov eq. 17), :136-149
    for (int jn = 3; jn <= trc; ++jn) {
        for (int jm = 2; jm < jn; ++jm) {
            double cn = ((2. * jn + 1.) * (jn + jm - 3.) * (jn + jm - 1.));
            double cd = ((2. * jn - 3.) * (jn + jm - 2.) * (jn + jm));
            double dn = ((2. * jn + 1.) * (jn - jm + 1.) * (jn + jm - 1.));
            double dd = ((2. * jn - 1.) * (jn + jm - 2.) * (jn + jm));
            double en = ((2. * jn + 1.) * (jn - jm));
            double ed = ((2. * jn - 1.) * (jn + jm));
            legpol[idx_mn(trc, jm, jn)] = std::sqrt(cn / cd) * legpol[idx_mn(trc, jm - 2, jn - 2)] -
                                          std::sqrt(dn / dd) * legpol[idx_mn(trc, jm - 2, jn - 1)] * costh +
                                          std::sqrt(en / ed) * legpol[idx_mn(trc, jm, jn - 1)] * costh;
        }
    }
}

size_t num_n(int truncation, int m, bool symmetric) {  // TransLocal.cc:183-187
    int len = (truncation - m + (symmetric ? 2 : 1)) / 2;
    return static_cast<size_t>(len < 0 ? 0 : len);
}
size_t add_padding(size_t n) { return static_cast<size_t>(std::ceil(n / 8.)) * 8; }  // TransLocal.cc:236-238
size_t legendre_size(size_t truncation) { return (truncation + 2) * (truncation + 1) / 2; }  // :174-176

// LegendrePolynomials.cc:154-209: sym/asym tables, n descending, [k + K*jlat]
void legendre_tables(int trc, int nlats, const double* lats, double* leg_sym, double* leg_asym,
                     const size_t* start_sym, const size_t* start_asym, int nthreads) {
    std::vector<double> zfn0(static_cast<size_t>(trc + 1) * (trc + 1));
    legendre_zfn(trc, zfn0.data());
    // NB: legendre_lat zeroes zfn(jn,0) for odd jn (reference :101) -- idempotent, so a per-thread
    // copy gives the same values as the reference's shared array.
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        std::vector<double> zfn(zfn0);
        std::vector<double> legpol(legendre_size(trc));
        std::vector<double> vs, vc;
#pragma omp for schedule(dynamic, 1)
        for (int jlat = 0; jlat < nlats; ++jlat) {
            legendre_lat(trc, lats[jlat], legpol.data(), zfn.data(), vs, vc);
            for (int jm = 0; jm <= trc; jm++) {
                size_t is1 = num_n(trc, jm, true), ia1 = num_n(trc, jm, false);
                size_t is2 = 0, ia2 = 0;
                for (int jn = trc; jn >= jm; jn--) {
                    if ((jn - jm) % 2 == 0) leg_sym[start_sym[jm] + is1 * jlat + is2++] = legpol[idx_mn(trc, jm, jn)];
                    else leg_asym[start_asym[jm] + ia1 * jlat + ia2++] = legpol[idx_mn(trc, jm, jn)];
                }
            }
        }
    }
}


}  // namespace orc
