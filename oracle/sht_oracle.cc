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
//     tables (tests/golden/gaussian_latitudes_N*.npy).
//   * unstructured path (orc_invtrans_unstructured): PINNED against the structured path on a
//     regular grid's points and the same closed-form harmonics at scattered points (1e-13).
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
#include <immintrin.h>
#include <omp.h>
#endif
#include "oracle_internal.h"

namespace {

constexpr double kLatPole     = 89.9999999;   // trans/local/TransLocal.cc:49
constexpr double kEarthRadius = 6371229.;     // util/Earth.h:24
constexpr double kDeg2Rad     = M_PI / 180.;  // util/Constants.h (degreesToRadians)
constexpr double kRad2Deg     = 180. * M_1_PI;

using cplx = std::complex<double>;

// -------------------------------------------------------------------------------------
// Generic complex FFT (any length): mixed radix with direct small-prime butterflies,
// Bluestein for lengths that contain a prime factor > 13.  sign = +1 => exp(+2 pi i jk/n).
// This stands in for FFTW/pocketfft (third-party, absent); checked against the naive DFT
// in tests/test_oracle_fft.py.
// -------------------------------------------------------------------------------------
struct FFTPlan {
    int n = 0;
    int sign = +1;
    std::vector<int> factors;
    std::vector<cplx> tw;  // exp(sign 2 pi i k / n), k<n
    // Bluestein
    bool bluestein = false;
    int m = 0;
    std::vector<cplx> chirp;   // exp(sign i pi k^2/n)
    std::vector<cplx> bfilt;   // FFT of conj chirp filter (length m)
    FFTPlan* sub = nullptr;    // power-of-two plan, forward (+)
    FFTPlan* subinv = nullptr;
    ~FFTPlan() {
        delete sub;
        delete subinv;
    }
};

void fft_exec(const FFTPlan& p, cplx* data, cplx* work);

FFTPlan* fft_plan(int n, int sign) {
    FFTPlan* p = new FFTPlan;
    p->n = n;
    p->sign = sign;
    int r = n;
    for (int f : {4, 2, 3, 5, 7, 11, 13}) {
        while (r % f == 0) {
            p->factors.push_back(f);
            r /= f;
        }
    }
    if (r != 1) {
        // Bluestein: x_k = chirp_k * sum_j (a_j chirp_j) * conj(chirp)_{k-j}
        p->bluestein = true;
        p->factors.clear();
        int m = 1;
        while (m < 2 * n - 1) m <<= 1;
        p->m = m;
        p->chirp.resize(n);
        for (int k = 0; k < n; ++k) {
            long long k2 = (static_cast<long long>(k) * k) % (2LL * n);
            double ang = sign * M_PI * static_cast<double>(k2) / n;
            p->chirp[k] = cplx(std::cos(ang), std::sin(ang));
        }
        p->sub = fft_plan(m, +1);
        p->subinv = fft_plan(m, -1);
        p->bfilt.assign(m, cplx(0, 0));
        p->bfilt[0] = std::conj(p->chirp[0]);
        for (int k = 1; k < n; ++k) {
            p->bfilt[k] = std::conj(p->chirp[k]);
            p->bfilt[m - k] = std::conj(p->chirp[k]);
        }
        // conj(chirp) filter must be convolved: y_k = sum_j a'_j b_{k-j}, b_k = chirp_k^{-1}... see exec
        std::vector<cplx> w(m);
        fft_exec(*p->sub, p->bfilt.data(), w.data());
        return p;
    }
    p->tw.resize(n);
    for (int k = 0; k < n; ++k) {
        double ang = sign * 2. * M_PI * k / n;
        p->tw[k] = cplx(std::cos(ang), std::sin(ang));
    }
    return p;
}

// Stockham autosort, decimation in frequency.  Radix 4 and 2 are specialised (they carry the power-of-two
// transforms inside Bluestein, i.e. most of the CPU baseline's Fourier time); other radices use the
// generic O(f^2) butterfly.  Per-stage twiddles are read with a running index (no modulo in the loops).
void fft_exec(const FFTPlan& p, cplx* data, cplx* work) {
    const int n = p.n;
    if (n <= 1) return;
    if (p.bluestein) {
        const int m = p.m;
        static thread_local std::vector<cplx> a, w;  // per-thread work space (plans are shared between threads)
        a.assign(m, cplx(0, 0));
        w.resize(m);
        // With chirp_k = exp(s i pi k^2/n):  exp(s 2 pi i jk/n) = chirp_j chirp_k conj(chirp_{k-j})
        for (int j = 0; j < n; ++j) a[j] = data[j] * p.chirp[j];
        fft_exec(*p.sub, a.data(), w.data());
        for (int k = 0; k < m; ++k) a[k] *= p.bfilt[k];
        fft_exec(*p.subinv, a.data(), w.data());
        const double sc = 1. / m;
        for (int k = 0; k < n; ++k) data[k] = a[k] * p.chirp[k] * sc;
        return;
    }
    cplx* x = data;
    cplx* y = work;
    int l = n;  // remaining length
    int s = 1;  // stride
    const double sg = p.sign;
    for (int f : p.factors) {
        const int mm = l / f;
        const int tstep = n / l;  // twiddle exp(sign 2 pi i q b / l) = tw[q b tstep]
        if (f == 4) {
            for (int q = 0; q < mm; ++q) {
                const cplx w1 = p.tw[q * tstep], w2 = p.tw[2 * q * tstep], w3 = p.tw[3 * q * tstep];
                for (int r = 0; r < s; ++r) {
                    const cplx a0 = x[r + s * q], a1 = x[r + s * (q + mm)], a2 = x[r + s * (q + 2 * mm)],
                               a3 = x[r + s * (q + 3 * mm)];
                    const cplx s02 = a0 + a2, d02 = a0 - a2, s13 = a1 + a3;
                    const cplx d13 = cplx(-sg * (a1.imag() - a3.imag()), sg * (a1.real() - a3.real()));  // (sign i)(a1-a3)
                    y[r + s * (4 * q)] = s02 + s13;
                    y[r + s * (4 * q + 1)] = (d02 + d13) * w1;
                    y[r + s * (4 * q + 2)] = (s02 - s13) * w2;
                    y[r + s * (4 * q + 3)] = (d02 - d13) * w3;
                }
            }
        }
        else if (f == 2) {
            for (int q = 0; q < mm; ++q) {
                const cplx w1 = p.tw[q * tstep];
                for (int r = 0; r < s; ++r) {
                    const cplx a0 = x[r + s * q], a1 = x[r + s * (q + mm)];
                    y[r + s * (2 * q)] = a0 + a1;
                    y[r + s * (2 * q + 1)] = (a0 - a1) * w1;
                }
            }
        }
        else {
            const int fstep = n / f;  // exp(sign 2 pi i a b / f) = tw[(a b mod f) fstep]
            for (int q = 0; q < mm; ++q) {
                cplx t[16];
                for (int r = 0; r < s; ++r) {
                    for (int a = 0; a < f; ++a) t[a] = x[r + s * (q + mm * a)];
                    for (int b = 0; b < f; ++b) {
                        cplx acc = t[0];
                        int ab = 0;
                        for (int a = 1; a < f; ++a) {
                            ab += b;
                            if (ab >= f) ab -= f;
                            acc += t[a] * p.tw[ab * fstep];
                        }
                        y[r + s * (b + f * q)] = acc * p.tw[(q * b) * tstep];
                    }
                }
            }
        }
        std::swap(x, y);
        l = mm;
        s *= f;
    }
    if (x != data) std::memcpy(data, x, sizeof(cplx) * n);
}

// -------------------------------------------------------------------------------------
// c2r / r2c built on the complex FFT (semantics of linalg/fft/FFTW.cc:38-62, fftw c2r).
// -------------------------------------------------------------------------------------
struct RealFFT {
    int n = 0;
    FFTPlan* inv = nullptr;  // +, length n
    FFTPlan* fwd = nullptr;  // -, length n
    ~RealFFT() {
        delete inv;
        delete fwd;
    }
};

RealFFT* rfft_plan(int n) {
    RealFFT* r = new RealFFT;
    r->n = n;
    r->inv = fft_plan(n, +1);
    r->fwd = fft_plan(n, -1);
    return r;
}

// out[j] = Re sum_k c_k X_k e^{+2 pi i jk/n}, Hermitian extension of in[0..n/2]
void c2r_fft(const RealFFT& r, const cplx* in, double* out, cplx* buf /*2n*/) {
    const int n = r.n;
    cplx* z = buf;
    z[0] = cplx(in[0].real(), 0.);
    for (int k = 1; k <= n / 2; ++k) {
        z[k] = in[k];
        if (n - k != k) z[n - k] = std::conj(in[k]);
        else z[k] = cplx(in[k].real(), 0.);
    }
    fft_exec(*r.inv, z, buf + n);
    for (int j = 0; j < n; ++j) out[j] = z[j].real();
}

// Two real rows of the same length in one complex transform (CPU-baseline fast path): z = x_a + i x_b has the
// spectrum Z_k = A_k + i B_k, Z_{n-k} = conj(A_k) + i conj(B_k).
void c2r_pair_fft(const RealFFT& r, const cplx* ina, const cplx* inb, double* outa, double* outb, cplx* buf /*2n*/) {
    const int n = r.n;
    cplx* z = buf;
    const cplx I(0., 1.);
    z[0] = cplx(ina[0].real(), inb[0].real());
    for (int k = 1; k <= n / 2; ++k) {
        if (n - k != k) {
            z[k] = ina[k] + I * inb[k];
            z[n - k] = std::conj(ina[k]) + I * std::conj(inb[k]);
        }
        else {
            z[k] = cplx(ina[k].real(), inb[k].real());
        }
    }
    fft_exec(*r.inv, z, buf + n);
    for (int j = 0; j < n; ++j) {
        outa[j] = z[j].real();
        outb[j] = z[j].imag();
    }
}
void r2c_pair_fft(const RealFFT& r, const double* ina, const double* inb, cplx* outa, cplx* outb, cplx* buf /*2n*/) {
    const int n = r.n;
    for (int j = 0; j < n; ++j) buf[j] = cplx(ina[j], inb[j]);
    fft_exec(*r.fwd, buf, buf + n);
    outa[0] = cplx(buf[0].real(), 0.);
    outb[0] = cplx(buf[0].imag(), 0.);
    for (int k = 1; k <= n / 2; ++k) {
        const cplx zk = buf[k], zc = std::conj(buf[n - k]);
        outa[k] = 0.5 * (zk + zc);
        outb[k] = cplx(0., -0.5) * (zk - zc);
    }
}

// literal DFT, O(n^2): the "reference-as-written" semantic anchor for the FFT
void c2r_naive(int n, const cplx* in, double* out) {
    for (int j = 0; j < n; ++j) {
        double acc = in[0].real();
        for (int k = 1; k <= n / 2; ++k) {
            double ang = 2. * M_PI * (static_cast<long long>(j) * k % n) / n;
            double term = in[k].real() * std::cos(ang) - in[k].imag() * std::sin(ang);
            acc += (2 * k == n) ? in[k].real() * std::cos(ang) : 2. * term;
        }
        out[j] = acc;
    }
}

// forward real -> half complex, X_k = sum_j x_j e^{-2 pi i jk/n}, k = 0..n/2
void r2c_fft(const RealFFT& r, const double* in, cplx* out, cplx* buf /*2n*/) {
    const int n = r.n;
    for (int j = 0; j < n; ++j) buf[j] = cplx(in[j], 0.);
    fft_exec(*r.fwd, buf, buf + n);
    for (int k = 0; k <= n / 2; ++k) out[k] = buf[k];
}

using orc::add_padding;
using orc::gaussian_quadrature_npole_equator;
using orc::legendre_lat;
using orc::legendre_size;
using orc::legendre_tables;
using orc::legendre_zfn;
using orc::num_n;

// TransLocal.cc:272-300
int fourier_truncation(int truncation, int nx, int /*nxmax*/, int ndgl, double lat, bool fullgrid) {
    int trc = truncation;
    int trclin = ndgl - 1;
    int trcquad = ndgl * 2 / 3 - 1;
    if (truncation >= trclin || fullgrid) {
        trc = (nx - 1) / 2;
    }
    else if (truncation >= trcquad) {
        double weight = 3 * (trclin - truncation) / ndgl;  // integer division, as in the reference :287
        double sqcos = std::pow(std::cos(lat), 2);
        trc = static_cast<int>((nx - 1) / (2 + weight * sqcos));
    }
    else {
        double sqcos = std::pow(std::cos(lat), 2);
        trc = static_cast<int>((nx - 1) / (2 + sqcos) - 1);
    }
    return std::min(truncation, trc);
}

// -------------------------------------------------------------------------------------
// Plan = the state the TransLocal constructor builds (TransLocal.cc:322-770) for a
// *global* structured grid without projection (reduced or regular, with or without
// pole/equator rows).
// -------------------------------------------------------------------------------------
struct Plan {
    int nlat = 0;
    int T = 0;
    bool regular = false;
    std::vector<int> nx;
    std::vector<double> lat_deg;
    int nxmax = 0;
    size_t npts = 0;
    int nlatsNH = 0, nlatsSH = 0, nlatsLeg = 0, nlatsLegReduced = 0, nlatsLegDomain = 0;
    std::vector<int> nlat0;
    std::vector<size_t> sym_begin, asym_begin;
    std::vector<double> leg_sym, leg_asym;
    std::vector<double> weights;  // per-latitude quadrature weights for dirtrans (sum = 1)
    std::vector<RealFFT*> ffts;   // indexed by nx value (sparse)
    int nthreads = 1;
    ~Plan() {
        for (auto* f : ffts) delete f;
    }
    const RealFFT& fft(int n) const { return *ffts[n]; }
};

bool approx_equal(double a, double b) {  // eckit::types::is_approximately_equal default eps
    return std::abs(a - b) <= 1e-12 * std::max(1., std::max(std::abs(a), std::abs(b)));
}

Plan* plan_create(int nlat, const int* nx, const double* lat_deg, int truncation, int regular,
                  const double* weights, int nthreads) {
    Plan* p = new Plan;
    p->nlat = nlat;
    p->T = truncation;
    p->regular = regular != 0;
    p->nx.assign(nx, nx + nlat);
    p->lat_deg.assign(lat_deg, lat_deg + nlat);
    p->nthreads = nthreads > 0 ? nthreads : 1;
    if (weights) p->weights.assign(weights, weights + nlat);
    for (int j = 0; j < nlat; ++j) {
        p->nxmax = std::max(p->nxmax, nx[j]);
        p->npts += nx[j];
    }
    // hemisphere bookkeeping, TransLocal.cc:376-392
    int neqtr = 0;
    for (int j = 0; j < nlat; ++j) {
        double lat = lat_deg[j];
        if (approx_equal(lat, 0.)) neqtr++;
        else if (lat < 0) p->nlatsSH++;
        else p->nlatsNH++;
    }
    if (neqtr > 0) {
        p->nlatsNH++;
        p->nlatsSH++;
    }
    p->nlatsLegDomain = std::max(p->nlatsNH, p->nlatsSH);
    const int nlatsGlobal = nlat;
    p->nlatsLeg = (nlatsGlobal + 1) / 2;                 // :435
    const int jlatMinLeg = 0;                            // global domain, :448-457
    p->nlatsLegReduced = jlatMinLeg + p->nlatsLegDomain;  // :459
    // nlat0, :462-488
    p->nlat0.assign(truncation + 1, 0);
    {
        int nmen0 = -1;
        for (int jlat = 0; jlat < nlatsGlobal / 2; jlat++) {
            double lat = lat_deg[jlat] * kDeg2Rad;
            int nmen = fourier_truncation(truncation, nx[jlat], p->nxmax, nlatsGlobal, lat, p->regular);
            nmen = std::max(nmen0, nmen);
            int ndgluj = std::max(jlatMinLeg, jlat);
            for (int j = nmen0 + 1; j <= nmen; j++) p->nlat0[j] = ndgluj;
            nmen0 = nmen;
        }
        for (int j = nmen0 + 1; j <= truncation; j++) p->nlat0[j] = p->nlatsLeg;
    }
    // Legendre latitudes with pole clamp, :533-546
    std::vector<double> lats(p->nlatsLeg);
    for (int j = 0; j < p->nlatsLeg; ++j) {
        double lat = lat_deg[j];
        if (lat > kLatPole) lat = kLatPole;
        if (lat < -kLatPole) lat = -kLatPole;
        lats[j] = lat * kDeg2Rad;
    }
    // table offsets, :592-606
    p->sym_begin.assign(truncation + 3, 0);
    p->asym_begin.assign(truncation + 3, 0);
    size_t size_sym = 0, size_asym = 0;
    for (int jm = 0; jm <= truncation + 1; jm++) {
        size_sym += add_padding(num_n(truncation + 1, jm, true) * static_cast<size_t>(p->nlatsLeg));
        size_asym += add_padding(num_n(truncation + 1, jm, false) * static_cast<size_t>(p->nlatsLeg));
        p->sym_begin[jm + 1] = size_sym;
        p->asym_begin[jm + 1] = size_asym;
    }
    p->leg_sym.assign(size_sym, 0.);
    p->leg_asym.assign(size_asym, 0.);
    legendre_tables(truncation + 1, p->nlatsLeg, lats.data(), p->leg_sym.data(), p->leg_asym.data(),
                    p->sym_begin.data(), p->asym_begin.data(), p->nthreads);  // :634-637
    // FFT "plans", :651-686
    p->ffts.assign(p->nxmax + 1, nullptr);
    for (int j = 0; j < nlat; ++j)
        if (!p->ffts[nx[j]]) p->ffts[nx[j]] = rfft_plan(nx[j]);
    return p;
}

inline size_t pos_fourier(const Plan& p, int jfld, int imag, int jlat, int jm, int nb_fields, int nlats) {
    // TransLocal.h:177-180 (posMethod), widened to size_t
    return static_cast<size_t>(imag) +
           2 * (static_cast<size_t>(jm) + static_cast<size_t>(p.T + 1) * (jlat + static_cast<size_t>(nlats) * jfld));
}

// naive column-major GEMM C(MxN) = A(MxK) B(KxN): stands in for eckit::linalg gemm
// (linalg/dense/MatrixMultiply_EckitLinalg.cc:64-67)
void gemm_naive(const double* A, const double* B, double* C, size_t M, size_t K, size_t N) {
    for (size_t j = 0; j < N; ++j) {
        for (size_t i = 0; i < M; ++i) C[i + M * j] = 0.;
        for (size_t k = 0; k < K; ++k) {
            const double b = B[k + K * j];
            for (size_t i = 0; i < M; ++i) C[i + M * j] += A[i + M * k] * b;
        }
    }
}

// Register-blocked GEMM of the CPU baseline: C(M x N, column-major, ld M) = A(M x K, column-major, ld M) * B, where
// b(k, n) = B[k * sbk + n * sbn] (sbk = 1, sbn = K: the column-major B of the inverse transform; sbk = K', sbn = 1: a
// transposed table block in the direct transform).  Same maths as gemm_naive.  The image has no BLAS (eckit's "lapack" /
// "mkl" backends are what the reference would use on a production host), so this is a plain micro-kernel: 16 rows x 8
// columns of accumulators in AVX-512 registers (masked loads / stores for the last rows), 8 x 6 in AVX2, chosen at run
// time -- the library is built in one container and timed on another machine's host cores.

__attribute__((target("avx512f"))) static void gemm_kernel_avx512(const double* A, const double* B, double* C, size_t M, size_t K,
                                                                  size_t N, size_t sbk, size_t sbn, bool accumulate) {
    constexpr size_t NR = 8;
    for (size_t j = 0; j < N; j += NR) {
        const size_t nb = std::min(NR, N - j);
        for (size_t i0 = 0; i0 < M; i0 += 16) {
            const size_t mb = std::min<size_t>(16, M - i0);
            const __mmask8 m0 = static_cast<__mmask8>(mb >= 8 ? 0xff : (1u << mb) - 1u);
            const __mmask8 m1 = static_cast<__mmask8>(mb >= 16 ? 0xff : (mb > 8 ? (1u << (mb - 8)) - 1u : 0u));
            __m512d acc0[NR], acc1[NR];
            for (size_t c = 0; c < NR; ++c) acc0[c] = acc1[c] = _mm512_setzero_pd();
            const double* a = A + i0;
            if (nb == NR) {
                for (size_t k = 0; k < K; ++k, a += M) {
                    const __m512d a0 = _mm512_maskz_loadu_pd(m0, a), a1 = _mm512_maskz_loadu_pd(m1, a + 8);
                    const double* b = B + k * sbk + j * sbn;
#pragma GCC unroll 8
                    for (size_t c = 0; c < NR; ++c) {
                        const __m512d bv = _mm512_set1_pd(b[c * sbn]);
                        acc0[c] = _mm512_fmadd_pd(a0, bv, acc0[c]);
                        acc1[c] = _mm512_fmadd_pd(a1, bv, acc1[c]);
                    }
                }
            }
            else {
                for (size_t k = 0; k < K; ++k, a += M) {
                    const __m512d a0 = _mm512_maskz_loadu_pd(m0, a), a1 = _mm512_maskz_loadu_pd(m1, a + 8);
                    const double* b = B + k * sbk + j * sbn;
                    for (size_t c = 0; c < nb; ++c) {
                        const __m512d bv = _mm512_set1_pd(b[c * sbn]);
                        acc0[c] = _mm512_fmadd_pd(a0, bv, acc0[c]);
                        acc1[c] = _mm512_fmadd_pd(a1, bv, acc1[c]);
                    }
                }
            }
            for (size_t c = 0; c < nb; ++c) {
                double* cp = C + i0 + M * (j + c);
                if (accumulate) {
                    acc0[c] = _mm512_add_pd(acc0[c], _mm512_maskz_loadu_pd(m0, cp));
                    acc1[c] = _mm512_add_pd(acc1[c], _mm512_maskz_loadu_pd(m1, cp + 8));
                }
                _mm512_mask_storeu_pd(cp, m0, acc0[c]);
                _mm512_mask_storeu_pd(cp + 8, m1, acc1[c]);
            }
        }
    }
}

static void gemm_kernel_avx2(const double* A, const double* B, double* C, size_t M, size_t K, size_t N, size_t sbk, size_t sbn,
                             bool accumulate) {
    constexpr size_t NR = 6, MR = 8;
    for (size_t j = 0; j < N; j += NR) {
        const size_t nb = std::min(NR, N - j);
        size_t i0 = 0;
        for (; i0 + MR <= M; i0 += MR) {
            __m256d acc0[NR], acc1[NR];
            for (size_t c = 0; c < NR; ++c) acc0[c] = acc1[c] = _mm256_setzero_pd();
            const double* a = A + i0;
            for (size_t k = 0; k < K; ++k, a += M) {
                const __m256d a0 = _mm256_loadu_pd(a), a1 = _mm256_loadu_pd(a + 4);
                const double* b = B + k * sbk + j * sbn;
                for (size_t c = 0; c < nb; ++c) {
                    const __m256d bv = _mm256_set1_pd(b[c * sbn]);
                    acc0[c] = _mm256_fmadd_pd(a0, bv, acc0[c]);
                    acc1[c] = _mm256_fmadd_pd(a1, bv, acc1[c]);
                }
            }
            for (size_t c = 0; c < nb; ++c) {
                double* cp = C + i0 + M * (j + c);
                if (accumulate) {
                    acc0[c] = _mm256_add_pd(acc0[c], _mm256_loadu_pd(cp));
                    acc1[c] = _mm256_add_pd(acc1[c], _mm256_loadu_pd(cp + 4));
                }
                _mm256_storeu_pd(cp, acc0[c]);
                _mm256_storeu_pd(cp + 4, acc1[c]);
            }
        }
        for (; i0 < M; ++i0)   // last rows
            for (size_t c = 0; c < nb; ++c) {
                double sum = accumulate ? C[i0 + M * (j + c)] : 0.;
                for (size_t k = 0; k < K; ++k) sum += A[i0 + M * k] * B[k * sbk + (j + c) * sbn];
                C[i0 + M * (j + c)] = sum;
            }
    }
}

static void gemm_fast(const double* A, const double* B, double* C, size_t M, size_t K, size_t N, size_t sbk, size_t sbn,
                      bool accumulate = false) {
    static const bool have512 = __builtin_cpu_supports("avx512f");
    if (have512) gemm_kernel_avx512(A, B, C, M, K, N, sbk, sbn, accumulate);
    else gemm_kernel_avx2(A, B, C, M, K, N, sbk, sbn, accumulate);
}

void gemm_blocked(const double* A, const double* B, double* C, size_t M, size_t K, size_t N) {
    gemm_fast(A, B, C, M, K, N, 1, K);
}

// CPU-baseline variant of invtrans_legendre: same split / GEMM / merge, but threads own blocks of 8 consecutive
// zonal wavenumbers so that the merge writes 8 adjacent complex values per (field, latitude) instead of one
// value per cache line (the reference's posMethod layout makes m the fastest index).
void invtrans_legendre_blocked(const Plan& p, int truncation, int nlats, int nb_fields, const double* spectra,
                               double* scl_fourier) {
    const int T = p.T;
    constexpr int MB = 8;
    const int nblocks = (T + 1 + MB - 1) / MB;
#pragma omp parallel for schedule(dynamic, 1) num_threads(p.nthreads)
    for (int blk = 0; blk < nblocks; ++blk) {
        const int m0 = blk * MB, m1 = std::min(T + 1, m0 + MB);
        std::vector<std::vector<double>> cs(MB), ca(MB);
        int ncols_m[MB], nimag_m[MB];
        for (int jm = m0; jm < m1; ++jm) {
            const int b = jm - m0;
            const size_t size_sym = num_n(T + 1, jm, true), size_asym = num_n(T + 1, jm, false);
            const int n_imag = (jm ? 2 : 1);
            const int ncols = p.nlatsLegReduced - p.nlat0[jm];
            ncols_m[b] = ncols;
            nimag_m[b] = n_imag;
            if (ncols <= 0) continue;
            const size_t rows = static_cast<size_t>(nb_fields) * n_imag;
            std::vector<double> a_sym(rows * size_sym), a_asym(rows * size_asym);
            cs[b].assign(rows * ncols, 0.);
            ca[b].assign(rows * ncols, 0.);
            size_t is = 0, ia = 0;
            const size_t ioff = static_cast<size_t>(2 * truncation + 3 - jm) * jm / 2 * nb_fields * 2;
            for (int jn = T + 1; jn >= jm; jn--)
                for (int imag = 0; imag < n_imag; imag++)
                    for (int jfld = 0; jfld < nb_fields; jfld++) {
                        size_t idx = jfld + static_cast<size_t>(nb_fields) * (imag + 2 * (jn - jm));
                        double v = (jn <= truncation && jm < truncation) ? spectra[idx + ioff] : 0.;
                        if ((jn - jm) % 2 == 0) a_sym[is++] = v;
                        else a_asym[ia++] = v;
                    }
            gemm_blocked(a_sym.data(), p.leg_sym.data() + p.sym_begin[jm] + p.nlat0[jm] * size_sym, cs[b].data(), rows,
                         size_sym, ncols);
            if (size_asym > 0)
                gemm_blocked(a_asym.data(), p.leg_asym.data() + p.asym_begin[jm] + p.nlat0[jm] * size_asym, ca[b].data(),
                             rows, size_asym, ncols);
        }
        for (int jfld = 0; jfld < nb_fields; jfld++) {
            for (int jlat = 0; jlat < std::max(p.nlatsNH, p.nlatsSH); jlat++) {
                for (int jm = m0; jm < m1; ++jm) {
                    const int b = jm - m0;
                    const int ncols = ncols_m[b], n_imag = nimag_m[b];
                    for (int imag = 0; imag < n_imag; imag++) {
                        if (jlat < p.nlatsNH) {
                            const int col = ncols - p.nlatsNH + jlat;
                            double v = 0.;
                            if (ncols > 0 && col >= 0) {
                                const size_t idx = jfld + static_cast<size_t>(nb_fields) * (imag + n_imag * col);
                                v = cs[b][idx] + ca[b][idx];
                            }
                            scl_fourier[pos_fourier(p, jfld, imag, jlat, jm, nb_fields, nlats)] = v;
                        }
                    }
                }
                for (int jm = m0; jm < m1; ++jm) {  // southern rows second: the equator row ends up as sym - asym (:1061-1070)
                    const int b = jm - m0;
                    const int ncols = ncols_m[b], n_imag = nimag_m[b];
                    for (int imag = 0; imag < n_imag; imag++) {
                        if (jlat < p.nlatsSH) {
                            const int col = ncols - p.nlatsSH + jlat;
                            double v = 0.;
                            if (ncols > 0 && col >= 0) {
                                const size_t idx = jfld + static_cast<size_t>(nb_fields) * (imag + n_imag * col);
                                v = cs[b][idx] - ca[b][idx];
                            }
                            scl_fourier[pos_fourier(p, jfld, imag, nlats - jlat - 1, jm, nb_fields, nlats)] = v;
                        }
                    }
                }
            }
        }
    }
}

// TransLocal::invtrans_legendre, TransLocal.cc:939-1097.  `fast` only changes the GEMM loop
// nest and threads over m; the split / merge follow the reference literally.
void invtrans_legendre(const Plan& p, int truncation, int nlats, int nb_fields, const double* spectra,
                       double* scl_fourier, bool fast) {
    const int T = p.T;
    if (fast) {
        invtrans_legendre_blocked(p, truncation, nlats, nb_fields, spectra, scl_fourier);
        return;
    }
    for (int jm = 0; jm <= T; jm++) {
        const size_t size_sym = num_n(T + 1, jm, true);
        const size_t size_asym = num_n(T + 1, jm, false);
        const int n_imag = (jm ? 2 : 1);
        const int ncols = p.nlatsLegReduced - p.nlat0[jm];
        const long size_fourier = static_cast<long>(nb_fields) * n_imag * ncols;
        if (size_fourier > 0) {
            const size_t rows = static_cast<size_t>(nb_fields) * n_imag;
            std::vector<double> a_sym(rows * size_sym), a_asym(rows * size_asym);
            std::vector<double> c_sym(size_fourier), c_asym(size_fourier, 0.);
            {
                size_t is = 0, ia = 0;
                const size_t ioff = static_cast<size_t>(2 * truncation + 3 - jm) * jm / 2 * nb_fields * 2;  // :970
                for (int jn = T + 1; jn >= jm; jn--) {  // descending n, :978
                    for (int imag = 0; imag < n_imag; imag++) {
                        for (int jfld = 0; jfld < nb_fields; jfld++) {
                            size_t idx = jfld + static_cast<size_t>(nb_fields) * (imag + 2 * (jn - jm));
                            double v = (jn <= truncation && jm < truncation) ? spectra[idx + ioff] : 0.;  // :982
                            if ((jn - jm) % 2 == 0) a_sym[is++] = v;
                            else a_asym[ia++] = v;
                        }
                    }
                }
            }
            {
                const double* Bs = p.leg_sym.data() + p.sym_begin[jm] + p.nlat0[jm] * size_sym;  // :1008
                const double* Ba = p.leg_asym.data() + p.asym_begin[jm] + p.nlat0[jm] * size_asym;
                if (fast) {
                    gemm_blocked(a_sym.data(), Bs, c_sym.data(), rows, size_sym, ncols);
                    if (size_asym > 0) gemm_blocked(a_asym.data(), Ba, c_asym.data(), rows, size_asym, ncols);
                }
                else {
                    gemm_naive(a_sym.data(), Bs, c_sym.data(), rows, size_sym, ncols);
                    if (size_asym > 0) gemm_naive(a_asym.data(), Ba, c_asym.data(), rows, size_asym, ncols);
                }
            }
            auto posF = [&](int jfld, int imag, int jlat, int nlatsH) {  // :955-957
                return jfld + static_cast<size_t>(nb_fields) * (imag + n_imag * (ncols - nlatsH + jlat));
            };
            for (int jlat = 0; jlat < p.nlatsNH; jlat++) {  // :1034-1059
                const bool inside = ncols - p.nlatsNH + jlat >= 0;
                for (int imag = 0; imag < n_imag; imag++)
                    for (int jfld = 0; jfld < nb_fields; jfld++) {
                        double v = 0.;
                        if (inside) {
                            size_t idx = posF(jfld, imag, jlat, p.nlatsNH);
                            v = c_sym[idx] + c_asym[idx];
                        }
                        scl_fourier[pos_fourier(p, jfld, imag, jlat, jm, nb_fields, nlats)] = v;
                    }
            }
            for (int jlat = 0; jlat < p.nlatsSH; jlat++) {  // :1061-1079
                const int jslat = nlats - jlat - 1;
                const bool inside = ncols - p.nlatsSH + jlat >= 0;
                for (int imag = 0; imag < n_imag; imag++)
                    for (int jfld = 0; jfld < nb_fields; jfld++) {
                        double v = 0.;
                        if (inside) {
                            size_t idx = posF(jfld, imag, jlat, p.nlatsSH);
                            v = c_sym[idx] - c_asym[idx];
                        }
                        scl_fourier[pos_fourier(p, jfld, imag, jslat, jm, nb_fields, nlats)] = v;
                    }
            }
        }
        else {
            for (int jlat = 0; jlat < nlats; jlat++)
                for (int imag = 0; imag < n_imag; imag++)
                    for (int jfld = 0; jfld < nb_fields; jfld++)
                        scl_fourier[pos_fourier(p, jfld, imag, jlat, jm, nb_fields, nlats)] = 0.;
        }
    }
}

// TransLocal::invtrans_fourier_reduced (:1155-1196) and _regular with FFT (:1101-1137), global
// domain (jlonMin = 0).  Both reduce to: for fld, for lat: pack n/2+1 complex, c2r(nx_lat), copy.
void invtrans_fourier(const Plan& p, int nlats, int nb_fields, const double* scl_fourier, double* gp, bool fast,
                      bool naive_dft) {
    std::vector<size_t> row_off(nlats + 1, 0);
    for (int j = 0; j < nlats; ++j) row_off[j + 1] = row_off[j] + p.nx[j];
    const size_t npts = row_off[nlats];
    if (fast && !naive_dft) {
        // CPU-baseline variant: rows j and nlats-1-j have the same length on a global grid -> one complex FFT per pair
        const int npairs = (nlats + 1) / 2;
#pragma omp parallel num_threads(p.nthreads)
        {
            std::vector<cplx> ina(p.nxmax / 2 + 1), inb(p.nxmax / 2 + 1), buf(2 * static_cast<size_t>(p.nxmax));
#pragma omp for collapse(2) schedule(dynamic, 8)
            for (int jfld = 0; jfld < nb_fields; jfld++) {
                for (int jp = 0; jp < npairs; jp++) {
                    const int ja = jp, jb = nlats - 1 - jp;
                    const int n = p.nx[ja];
                    const int num_complex = n / 2 + 1;
                    auto pack = [&](int jlat, cplx* in) {
                        in[0] = cplx(scl_fourier[pos_fourier(p, jfld, 0, jlat, 0, nb_fields, nlats)], 0.);
                        const size_t base = pos_fourier(p, jfld, 0, jlat, 0, nb_fields, nlats);
                        const int top = std::min(num_complex - 1, p.T);
                        for (int jm = 1; jm <= top; jm++) in[jm] = cplx(scl_fourier[base + 2 * jm], scl_fourier[base + 2 * jm + 1]);
                        for (int jm = top + 1; jm < num_complex; jm++) in[jm] = cplx(0., 0.);
                    };
                    pack(ja, ina.data());
                    if (jb != ja && p.nx[jb] == n) {
                        pack(jb, inb.data());
                        c2r_pair_fft(p.fft(n), ina.data(), inb.data(), gp + npts * jfld + row_off[ja],
                                     gp + npts * jfld + row_off[jb], buf.data());
                    }
                    else {
                        c2r_fft(p.fft(n), ina.data(), gp + npts * jfld + row_off[ja], buf.data());
                        if (jb != ja) {
                            pack(jb, inb.data());
                            c2r_fft(p.fft(p.nx[jb]), inb.data(), gp + npts * jfld + row_off[jb], buf.data());
                        }
                    }
                }
            }
        }
        return;
    }
#pragma omp parallel num_threads(fast ? p.nthreads : 1)
    {
        std::vector<cplx> in(p.nxmax / 2 + 1), buf(2 * static_cast<size_t>(p.nxmax));
#pragma omp for collapse(2) schedule(static)
        for (int jfld = 0; jfld < nb_fields; jfld++) {
            for (int jlat = 0; jlat < nlats; jlat++) {
                const int n = p.nx[jlat];
                const int num_complex = n / 2 + 1;
                in[0] = cplx(scl_fourier[pos_fourier(p, jfld, 0, jlat, 0, nb_fields, nlats)], 0.);  // :1165
                for (int jm = 1; jm < num_complex; jm++) {
                    if (jm <= p.T)
                        in[jm] = cplx(scl_fourier[pos_fourier(p, jfld, 0, jlat, jm, nb_fields, nlats)],
                                      scl_fourier[pos_fourier(p, jfld, 1, jlat, jm, nb_fields, nlats)]);
                    else in[jm] = cplx(0., 0.);
                }
                double* out = gp + npts * jfld + row_off[jlat];
                if (naive_dft) c2r_naive(n, in.data(), out);
                else c2r_fft(p.fft(n), in.data(), out, buf.data());
            }
        }
    }
}

// TransLocal::invtrans_uv, :1409-1484 (structured branch)
void invtrans_uv(const Plan& p, int truncation, int nb_scalar_fields, int nb_vordiv_fields, const double* spectra,
                 double* gp, bool fast, bool naive_dft) {
    if (nb_scalar_fields <= 0) return;
    const int nb_fields = nb_scalar_fields;
    const int nlats = p.nlat;
    const size_t size_fourier_max = static_cast<size_t>(nb_fields) * 2 * nlats;
    std::vector<double> scl_fourier(size_fourier_max * (p.T + 1), 0.);  // zero fill, :1426-1428
    invtrans_legendre(p, truncation, nlats, nb_fields, spectra, scl_fourier.data(), fast);
    invtrans_fourier(p, nlats, nb_fields, scl_fourier.data(), gp, fast, naive_dft);
    if (nb_vordiv_fields > 0) {  // u,v = U,V / cos(lat), :1443-1469
        std::vector<double> coslatinvs(nlats);
        for (int j = 0; j < nlats; ++j) {
            double lat = p.lat_deg[j];
            if (lat > kLatPole) lat = kLatPole;
            if (lat < -kLatPole) lat = -kLatPole;
            coslatinvs[j] = 1. / std::cos(lat * kDeg2Rad);
        }
        size_t idx = 0;
        for (int jfld = 0; jfld < 2 * nb_vordiv_fields && jfld < nb_fields; jfld++)
            for (int jlat = 0; jlat < nlats; jlat++)
                for (int jlon = 0; jlon < p.nx[jlat]; jlon++) gp[idx++] *= coslatinvs[jlat];
    }
}

// extend_truncation, TransLocal.cc:1496-1519
void extend_truncation(int old_truncation, int nb_fields, const double* old_spectra, double* new_spectra) {
    const int new_truncation = old_truncation + 1;
    size_t k = 0, k_old = 0;
    for (int m = 0; m <= new_truncation; m++)
        for (int n = m; n <= new_truncation; n++)
            for (int imag = 0; imag < 2; imag++)
                for (int jfld = 0; jfld < nb_fields; jfld++) {
                    if (m == new_truncation || n == new_truncation) new_spectra[k++] = 0.;
                    else new_spectra[k++] = old_spectra[k_old++];
                }
}

// prfi1b + vd2uv, trans/local/VorDivToUVLocal.cc:31-184 (Temperton 1991 eq 2.12/2.13)
void vd2uv(int truncation, int nb_vordiv_fields, const double* vorticity_spectra, const double* divergence_spectra,
           double* U, double* V) {
    const int T = truncation;
    const int nf = nb_vordiv_fields;
    std::vector<double> repsnm(static_cast<size_t>(T + 1) * (T + 6) / 2);
    const int nlei1 = T + 4 + (T + 4 + 1) % 2;
    const double ra = kEarthRadius;
    std::vector<double> rlapin(T + 3);
    {
        size_t idx = 0;
        for (int jm = 0; jm <= T; ++jm)
            for (int jn = jm; jn <= T + 2; ++jn, ++idx)
                repsnm[idx] = std::sqrt((jn * jn - jm * jm) / (4. * jn * jn - 1.));  // int*int as in :81
        repsnm[0] = 0.;
        for (int jn = 1; jn <= T + 2; ++jn) rlapin[jn] = -ra * ra / (jn * (jn + 1.));
        rlapin[0] = 0.;
    }
    auto prfi1b = [&](int km, const double* rspec, double* pia) {  // :31-53
        int ilcm = T + 1 - km, ioff = (2 * T - km + 3) * km;
        for (int j = 1; j <= ilcm; j++) {
            size_t inm = ioff + (ilcm - j) * 2;
            for (int jfld = 0; jfld < nf; jfld++) {
                int ir = 2 * jfld, ii = ir + 1;
                pia[ir * nlei1 + j + 1] = rspec[inm * nf + jfld];
                pia[ii * nlei1 + j + 1] = rspec[(inm + 1) * nf + jfld];
            }
        }
        for (int jfld = 0; jfld < 2 * nf; jfld++) {
            pia[jfld * nlei1] = 0.;
            pia[jfld * nlei1 + 1] = 0.;
            pia[jfld * nlei1 + ilcm + 2] = 0.;
        }
    };
    std::vector<double> zepsnm(T + 6), zlapin(T + 6), zn(T + 6);
    std::vector<double> rvor(2 * static_cast<size_t>(nf) * nlei1), rdiv(rvor.size()), ru(rvor.size()), rv(rvor.size());
    for (int km = 0; km <= T; ++km) {
        for (int jn = km - 1; jn <= T + 2; ++jn) {  // reversed order "for accuracy", :98-116
            int ij = T + 3 - jn;
            if (jn >= 0) {
                zlapin[ij] = rlapin[jn];
                zepsnm[ij] = (jn < km) ? 0. : repsnm[jn + (2 * T - km + 5) * km / 2];
            }
            else {
                zlapin[ij] = 0.;
                zepsnm[ij] = 0.;
            }
            zn[ij] = jn;
        }
        zn[0] = T + 3;
        std::fill(rvor.begin(), rvor.end(), 0.);
        std::fill(rdiv.begin(), rdiv.end(), 0.);
        std::fill(ru.begin(), ru.end(), 0.);
        std::fill(rv.begin(), rv.end(), 0.);
        prfi1b(km, vorticity_spectra, rvor.data());
        prfi1b(km, divergence_spectra, rdiv.data());
        if (km == 0) {  // :133-142
            for (int jfld = 0; jfld < nf; ++jfld) {
                int ir = 2 * jfld * nlei1 - 1;
                for (int ji = 2; ji < T + 4 - km; ++ji) {
                    double psiM1 = zn[ji + 1] * zepsnm[ji] * zlapin[ji + 1];
                    double psiP1 = zn[ji - 2] * zepsnm[ji - 1] * zlapin[ji - 1];
                    ru[ir + ji] = +psiM1 * rvor[ir + ji + 1] - psiP1 * rvor[ir + ji - 1];
                    rv[ir + ji] = -psiM1 * rdiv[ir + ji + 1] + psiP1 * rdiv[ir + ji - 1];
                }
            }
        }
        else {  // :144-156
            for (int jfld = 0; jfld < nf; ++jfld) {
                int ir = 2 * jfld * nlei1 - 1, ii = ir + nlei1;
                for (int ji = 2; ji < T + 4 - km; ++ji) {
                    double chiIm = km * zlapin[ji];
                    double psiM1 = zn[ji + 1] * zepsnm[ji] * zlapin[ji + 1];
                    double psiP1 = zn[ji - 2] * zepsnm[ji - 1] * zlapin[ji - 1];
                    ru[ir + ji] = -chiIm * rdiv[ii + ji] + psiM1 * rvor[ir + ji + 1] - psiP1 * rvor[ir + ji - 1];
                    ru[ii + ji] = +chiIm * rdiv[ir + ji] + psiM1 * rvor[ii + ji + 1] - psiP1 * rvor[ii + ji - 1];
                    rv[ir + ji] = -chiIm * rvor[ii + ji] - psiM1 * rdiv[ir + ji + 1] + psiP1 * rdiv[ir + ji - 1];
                    rv[ii + ji] = +chiIm * rvor[ir + ji] - psiM1 * rdiv[ii + ji + 1] + psiP1 * rdiv[ii + ji - 1];
                }
            }
        }
        {  // :160-181
            int ilcm = T - km;
            int ioff = (2 * T - km + 3) * km;
            double za_r = 1. / kEarthRadius;
            for (int j = 0; j <= ilcm; ++j) {
                size_t inm = ioff + (ilcm - j) * 2;
                for (int jfld = 0; jfld < nf; ++jfld) {
                    int ir = 2 * jfld * nlei1, ii = ir + nlei1;
                    size_t idx = inm * nf + jfld;
                    U[idx] = ru[ir + j + 2] * za_r;
                    V[idx] = rv[ir + j + 2] * za_r;
                    idx += nf;
                    U[idx] = ru[ii + j + 2] * za_r;
                    V[idx] = rv[ii + j + 2] * za_r;
                }
            }
        }
    }
}

// TransLocal::invtrans(nscal, sp, nvd, vor, div, gp), :1523-1597
void invtrans_full(const Plan& p, int nb_scalar_fields, const double* scalar_spectra, int nb_vordiv_fields,
                   const double* vorticity_spectra, const double* divergence_spectra, double* gp, bool fast,
                   bool naive_dft) {
    const int T = p.T;
    if (nb_vordiv_fields > 0) {
        const size_t ext = 2 * legendre_size(T + 1);
        std::vector<double> vor_ext(ext * nb_vordiv_fields), div_ext(ext * nb_vordiv_fields);
        std::vector<double> U_ext(ext * nb_vordiv_fields), V_ext(ext * nb_vordiv_fields), scalar_ext;
        extend_truncation(T, nb_vordiv_fields, vorticity_spectra, vor_ext.data());
        extend_truncation(T, nb_vordiv_fields, divergence_spectra, div_ext.data());
        vd2uv(T + 1, nb_vordiv_fields, vor_ext.data(), div_ext.data(), U_ext.data(), V_ext.data());
        if (nb_scalar_fields > 0) {
            scalar_ext.resize(ext * nb_scalar_fields);
            extend_truncation(T, nb_scalar_fields, scalar_spectra, scalar_ext.data());
        }
        const int nb_all = 2 * nb_vordiv_fields + nb_scalar_fields;
        std::vector<double> all(ext * nb_all);
        size_t k = 0, i = 0, j = 0, l = 0;
        for (int m = 0; m <= T + 1; m++)
            for (int n = m; n <= T + 1; n++)
                for (int imag = 0; imag < 2; imag++) {
                    for (int f = 0; f < nb_vordiv_fields; f++) all[k++] = U_ext[i++];
                    for (int f = 0; f < nb_vordiv_fields; f++) all[k++] = V_ext[j++];
                    for (int f = 0; f < nb_scalar_fields; f++) all[k++] = scalar_ext[l++];
                }
        invtrans_uv(p, T + 1, nb_all, nb_vordiv_fields, all.data(), gp, fast, naive_dft);
    }
    else if (nb_scalar_fields > 0) {
        invtrans_uv(p, T, nb_scalar_fields, 0, scalar_spectra, gp, fast, naive_dft);
    }
}

// -------------------------------------------------------------------------------------
// Direct transform (NOT in TransLocal -- parity unpinned; semantics of TransIFS/ectrans,
// trans/ifs/TransIFS.cc:503-515, and of test_transgeneral.cc:1494-1585):
//   F_m(lat) = (1/nx) sum_j f_j exp(-i m lambda_j)
//   X_n^m    = sum_lat w_lat Pbar_n^m(mu_lat) F_m(lat),   sum_lat w_lat = 1
// using the same per-latitude zonal truncation (nlat0) as the inverse.  All (m<=T, n<=T)
// coefficients are produced (no m==T drop here).
// -------------------------------------------------------------------------------------
// `Tout` = truncation of the produced spectra (T, or T+1 for the wind path below); `rowscale` (per latitude
// row, may be null) multiplies the grid values before the transform.
void dirtrans_general(const Plan& p, int nb_fields, const double* gp, double* spectra, int Tout, const double* rowscale) {
    const int T = p.T;
    const int nlats = p.nlat;
    const int nf = nb_fields;
    std::vector<size_t> row_off(nlats + 1, 0);
    for (int j = 0; j < nlats; ++j) row_off[j + 1] = row_off[j] + p.nx[j];
    const size_t npts = row_off[nlats];
    // Fourier stage: four[fld][lat][m] complex; rows j and nlats-1-j share one complex FFT
    std::vector<cplx> four(static_cast<size_t>(nf) * nlats * (T + 1), cplx(0, 0));
    const int npairs = (nlats + 1) / 2;
#pragma omp parallel num_threads(p.nthreads)
    {
        std::vector<cplx> oa(p.nxmax / 2 + 1), ob(p.nxmax / 2 + 1), buf(2 * static_cast<size_t>(p.nxmax));
#pragma omp for collapse(2) schedule(dynamic, 8)
        for (int f = 0; f < nf; ++f)
            for (int jp = 0; jp < npairs; ++jp) {
                const int ja = jp, jb = nlats - 1 - jp;
                const int n = p.nx[ja];
                const int mmax = std::min(T, (n - 1) / 2);
                auto put = [&](int jlat, const cplx* out) {
                    cplx* dst = &four[(static_cast<size_t>(f) * nlats + jlat) * (T + 1)];
                    const double rs = (rowscale ? rowscale[jlat] : 1.) / static_cast<double>(n);
                    for (int m = 0; m <= mmax; ++m) dst[m] = out[m] * rs;
                };
                if (jb != ja && p.nx[jb] == n) {
                    r2c_pair_fft(p.fft(n), gp + npts * f + row_off[ja], gp + npts * f + row_off[jb], oa.data(), ob.data(),
                                 buf.data());
                    put(ja, oa.data());
                    put(jb, ob.data());
                }
                else {
                    r2c_fft(p.fft(n), gp + npts * f + row_off[ja], oa.data(), buf.data());
                    put(ja, oa.data());
                    if (jb != ja) {
                        r2c_fft(p.fft(p.nx[jb]), gp + npts * f + row_off[jb], ob.data(), buf.data());
                        const int mm2 = std::min(T, (p.nx[jb] - 1) / 2);
                        cplx* dst = &four[(static_cast<size_t>(f) * nlats + jb) * (T + 1)];
                        const double rs = (rowscale ? rowscale[jb] : 1.) / static_cast<double>(p.nx[jb]);
                        for (int m = 0; m <= mm2; ++m) dst[m] = ob[m] * rs;
                    }
                }
            }
    }
    const size_t nspec2 = 2 * legendre_size(Tout);
    std::fill(spectra, spectra + nspec2 * nf, 0.);
    constexpr int MB = 8;
    const int nblocks = (T + 1 + MB - 1) / MB;
    const int R = 2 * nf;  // rows: f + nf*imag
#pragma omp parallel for schedule(dynamic, 1) num_threads(p.nthreads)
    for (int blk = 0; blk < nblocks; ++blk) {
        const int m0 = blk * MB, m1 = std::min(T + 1, m0 + MB);
        const int nl0 = p.nlat0[m0];  // nlat0 is non-decreasing in m
        const int ncmax = p.nlatsLegReduced - nl0;
        if (ncmax <= 0) continue;
        // gather w * (F_N +- F_S) for the 8 wavenumbers of the block: gs/ga[b][i + R*col]
        std::vector<double> gs(static_cast<size_t>(MB) * R * ncmax, 0.), ga(static_cast<size_t>(MB) * R * ncmax, 0.);
        for (int jl = nl0; jl < p.nlatsLegReduced; ++jl) {
            const int js_row = nlats - 1 - jl;
            const bool has_n = jl < p.nlatsNH;
            const bool has_s = jl < p.nlatsSH && js_row != jl;
            const double w = p.weights.empty() ? 0. : p.weights[jl];
            for (int f = 0; f < nf; ++f) {
                const cplx* rn = &four[(static_cast<size_t>(f) * nlats + jl) * (T + 1)];
                const cplx* rsn = &four[(static_cast<size_t>(f) * nlats + js_row) * (T + 1)];
                for (int m = m0; m < m1; ++m) {
                    if (jl < p.nlat0[m]) continue;
                    const cplx fn = has_n ? rn[m] : cplx(0, 0), fs = has_s ? rsn[m] : cplx(0, 0);
                    const cplx a = (fn + fs) * w, c = (fn - fs) * w;
                    const size_t base = (static_cast<size_t>(m - m0) * ncmax + (jl - nl0)) * R;
                    gs[base + f] = a.real();
                    gs[base + nf + f] = a.imag();
                    ga[base + f] = c.real();
                    ga[base + nf + f] = c.imag();
                }
            }
        }
        for (int m = m0; m < m1; ++m) {
            const int ncols = p.nlatsLegReduced - p.nlat0[m];
            if (ncols <= 0) continue;
            const size_t Ks = num_n(T + 1, m, true), Ka = num_n(T + 1, m, false);
            const size_t ioff = static_cast<size_t>(2 * Tout + 3 - m) * m / 2 * nf * 2;
            const size_t skip = static_cast<size_t>(p.nlat0[m] - nl0) * R;
            const double* Gs = gs.data() + static_cast<size_t>(m - m0) * ncmax * R + skip;
            const double* Ga = ga.data() + static_cast<size_t>(m - m0) * ncmax * R + skip;
            for (int par = 0; par < 2; ++par) {
                const size_t K = par ? Ka : Ks;
                if (K == 0) continue;
                const double* P = (par ? p.leg_asym.data() + p.asym_begin[m] : p.leg_sym.data() + p.sym_begin[m]) +
                                  K * static_cast<size_t>(p.nlat0[m]);
                const double* G = par ? Ga : Gs;
                // X[i + R*k] = sum_col G[i + R*col] * P[k + K*col]   (tables hold n descending: k = 0 <-> highest n)
                std::vector<double> X(static_cast<size_t>(R) * K, 0.);
                gemm_fast(G, P, X.data(), R, ncols, K, K, 1);   // b(col, k) = P[k + K * col]
                const int ntop = par == 0 ? ((T + 1 - m) % 2 == 0 ? T + 1 : T) : ((T + 1 - m) % 2 == 1 ? T + 1 : T);
                for (size_t k = 0; k < K; ++k) {
                    const int n = ntop - 2 * static_cast<int>(k);
                    if (n > Tout || n < m) continue;
                    const size_t base = ioff + static_cast<size_t>(nf) * 2 * (n - m);
                    for (int f = 0; f < nf; ++f) {
                        spectra[base + f] = X[k * R + f];
                        if (m > 0) spectra[base + nf + f] = X[k * R + nf + f];
                    }
                }
            }
        }
    }
}

void dirtrans(const Plan& p, int nb_fields, const double* gp, double* spectra) {
    dirtrans_general(p, nb_fields, gp, spectra, p.T, nullptr);
}

// Direct transform of wind to vorticity / divergence (NOT in TransLocal -- parity unpinned; API of
// TransImpl::dirtrans(nb_fields, wind, vor, div), trans/detail/TransImpl.h:180-181; wind layout
// [u_1..u_k | v_1..v_k][npts] as produced by the inverse, TransLocal.cc:1561-1589).
// With Ut = u/(a cos(lat)), Vt = v/(a cos(lat)) transformed as scalars up to n = T+1 and
// eps(n,m) = sqrt((n^2-m^2)/(4n^2-1)):
//   zeta_n^m = i m Vt_n^m + (n+1) eps(n,m) Ut_{n-1}^m - n eps(n+1,m) Ut_{n+1}^m
//   D_n^m    = i m Ut_n^m - (n+1) eps(n,m) Vt_{n-1}^m + n eps(n+1,m) Vt_{n+1}^m
// (integration by parts of zeta = (1/(a cos^2)) [dV/dlambda - cos d(U)/dphi], U = u cos, V = v cos).
void dirtrans_wind(const Plan& p, int nb_fields, const double* wind, double* vor, double* div) {
    const int T = p.T, nf = nb_fields;
    std::vector<double> rs(p.nlat);
    for (int j = 0; j < p.nlat; ++j) {
        double lat = p.lat_deg[j];
        if (lat > kLatPole) lat = kLatPole;
        if (lat < -kLatPole) lat = -kLatPole;
        rs[j] = 1. / (kEarthRadius * std::cos(lat * kDeg2Rad));
    }
    const size_t next = 2 * legendre_size(T + 1);
    std::vector<double> uv(next * 2 * nf);
    dirtrans_general(p, 2 * nf, wind, uv.data(), T + 1, rs.data());  // fields: u_1..u_k, v_1..v_k
    auto eps = [](int n, int m) { return (n == 0 || n < m) ? 0. : std::sqrt((double(n) * n - double(m) * m) / (4. * n * n - 1.)); };
    auto ext = [&](int m, int n, int imag, int f) -> double {  // f in [0, 2nf)
        if (n < m || n > T + 1) return 0.;
        return uv[(2 * (static_cast<size_t>(2 * (T + 1) + 3 - m) * m / 2 + (n - m)) + imag) * (2 * nf) + f];
    };
    for (int m = 0; m <= T; ++m)
        for (int n = m; n <= T; ++n) {
            const size_t c = static_cast<size_t>(2 * T + 3 - m) * m / 2 + (n - m);
            const double em = (n + 1) * eps(n, m), ep = n * eps(n + 1, m);
            for (int f = 0; f < nf; ++f) {
                const int fu = f, fv = nf + f;
                const double ur = ext(m, n, 0, fu), ui = ext(m, n, 1, fu), vr = ext(m, n, 0, fv), vi = ext(m, n, 1, fv);
                // i m (a + i b) = (-m b) + i (m a)
                double zr = -m * vi + em * ext(m, n - 1, 0, fu) - ep * ext(m, n + 1, 0, fu);
                double zi = +m * vr + em * ext(m, n - 1, 1, fu) - ep * ext(m, n + 1, 1, fu);
                double dr = -m * ui - em * ext(m, n - 1, 0, fv) + ep * ext(m, n + 1, 0, fv);
                double di = +m * ur - em * ext(m, n - 1, 1, fv) + ep * ext(m, n + 1, 1, fv);
                if (m == 0) zi = di = 0.;
                vor[(2 * c) * nf + f] = zr;
                vor[(2 * c + 1) * nf + f] = zi;
                div[(2 * c) * nf + f] = dr;
                div[(2 * c + 1) * nf + f] = di;
            }
        }
}

// Gradient of scalar fields (NOT in TransLocal, TransLocal.cc:848-857; TransIFS semantics ifs/TransIFS.cc:2075-2142):
// out = [E-W_1..E-W_k | N-S_1..N-S_k][npts] with E-W = (1/(a cos)) d/dlambda, N-S = (1/a) d/dphi.
// grad f is the irrotational wind of the velocity potential chi = f, i.e. the reference's own
// vorticity/divergence inverse with vor = 0 and D = laplace(f): D_n^m = -n(n+1)/a^2 f_n^m.
void invtrans_grad(const Plan& p, int nb_fields, const double* spectra, double* grad, bool fast) {
    const int T = p.T, nf = nb_fields;
    const size_t nspec = 2 * legendre_size(T) * nf;
    std::vector<double> vor(nspec, 0.), div(nspec);
    size_t c = 0;
    for (int m = 0; m <= T; ++m)
        for (int n = m; n <= T; ++n, ++c)
            for (int imag = 0; imag < 2; ++imag)
                for (int f = 0; f < nf; ++f)
                    div[(2 * c + imag) * nf + f] = -(n * (n + 1.)) / (kEarthRadius * kEarthRadius) * spectra[(2 * c + imag) * nf + f];
    invtrans_full(p, 0, nullptr, nf, vor.data(), div.data(), grad, fast, false);
}

}  // namespace

// =====================================================================================
// C ABI for ctypes (tests/, bench.py cpu_baseline) -- test infrastructure only.
// =====================================================================================
// TransLocal::invtrans_unstructured (TransLocal.cc:1289-1392), the default path of an UnstructuredGrid: for every point
// the Legendre functions at the point's latitude (compute_legendre_polynomials_lat, :1322), one (2 nf x ns)(ns x 1)
// product per zonal wavenumber jm <= truncation (:1327-1336: `jm <= truncation`, i.e. the m == truncation column IS
// used here, unlike the structured path's `jm < truncation`, :982), the Fourier sum as a dot product with
// (1, 0, 2 cos(jm lon), -2 sin(jm lon), ...) (:1352-1362), and u, v = U, V / cos(lat) for the first 2 nb_vordiv
// fields without any pole clamp (:1375-1384).  `spectra` is [m][n][re/im][fld] at `truncation`; gp is [fld][point].
static void invtrans_unstructured(int truncation, int nb_fields, int nb_vordiv_fields, const double* spectra, int npts,
                                  const double* lon_deg, const double* lat_deg, double* gp, int nthreads) {
    std::vector<double> zfn(static_cast<size_t>(truncation + 1) * (truncation + 1));
    legendre_zfn(truncation, zfn.data());
    const double deg2rad = M_PI / 180.;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        std::vector<double> legendre(static_cast<size_t>(truncation + 2) * (truncation + 1) / 2), vs, vc;
        std::vector<double> scl(static_cast<size_t>(2) * nb_fields * (truncation + 1));
#pragma omp for schedule(dynamic, 1)
        for (int ip = 0; ip < npts; ++ip) {
            const double lon = lon_deg[ip] * deg2rad, lat = lat_deg[ip] * deg2rad;
            legendre_lat(truncation, lat, legendre.data(), zfn.data(), vs, vc);
            // Legendre transform: scl[jm][imag][fld] = sum_n spectra[m][n][imag][fld] * P_n^m      (:1327-1336)
            std::fill(scl.begin(), scl.end(), 0.);
            for (int jm = 0; jm <= truncation; ++jm) {
                const size_t noff = static_cast<size_t>(2 * truncation + 3 - jm) * jm / 2;
                const int ns = truncation - jm + 1;
                double* c = scl.data() + static_cast<size_t>(jm) * 2 * nb_fields;
                for (int k = 0; k < ns; ++k) {
                    const double pk = legendre[noff + k];
                    const double* a = spectra + (noff + k) * 2 * nb_fields;
                    for (int r = 0; r < 2 * nb_fields; ++r) c[r] += a[r] * pk;
                }
            }
            // Fourier transformation                                                                 (:1352-1372)
            for (int f = 0; f < nb_fields; ++f) {
                double acc = 1. * scl[f] + 0. * scl[nb_fields + f];
                for (int jm = 1; jm <= truncation; ++jm) {
                    const double* c = scl.data() + static_cast<size_t>(jm) * 2 * nb_fields;
                    acc += (+2. * std::cos(jm * lon)) * c[f] + (-2. * std::sin(jm * lon)) * c[nb_fields + f];
                }
                gp[static_cast<size_t>(f) * npts + ip] = acc;
            }
            if (nb_vordiv_fields > 0) {                                                            // (:1375-1384)
                const double coslat = std::cos(lat);
                for (int j = 0; j < 2 * nb_vordiv_fields && j < nb_fields; ++j) gp[static_cast<size_t>(j) * npts + ip] /= coslat;
            }
        }
    }
}

extern "C" {

int orc_max_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_gaussian_quadrature(int N, double* lats_deg_2N, double* weights_2N) {
    gaussian_quadrature_npole_equator(N, lats_deg_2N, weights_2N);
    int end = 2 * N - 1;
    for (int j = 0; j < N; ++j, --end) {  // Latitudes.cc:81-86
        lats_deg_2N[end] = -lats_deg_2N[j];
        weights_2N[end] = weights_2N[j];
    }
}

void orc_compute_zfn(int trc, double* zfn) { legendre_zfn(trc, zfn); }

void orc_legendre_lat(int trc, double lat_rad, double* legpol, double* zfn) {
    std::vector<double> vs, vc;
    legendre_lat(trc, lat_rad, legpol, zfn, vs, vc);
}

int orc_fourier_truncation(int truncation, int nx, int nxmax, int ndgl, double lat_rad, int fullgrid) {
    return fourier_truncation(truncation, nx, nxmax, ndgl, lat_rad, fullgrid != 0);
}

void* orc_plan_create(int nlat, const int* nx, const double* lat_deg, int truncation, int regular,
                      const double* weights_or_null, int nthreads) {
    return plan_create(nlat, nx, lat_deg, truncation, regular, weights_or_null, nthreads);
}
void orc_plan_destroy(void* plan) { delete static_cast<Plan*>(plan); }

int orc_plan_nlat0(void* plan, int* nlat0 /*T+1*/) {
    Plan* p = static_cast<Plan*>(plan);
    std::copy(p->nlat0.begin(), p->nlat0.end(), nlat0);
    return p->nlatsLeg;
}
size_t orc_plan_table_sizes(void* plan, size_t* sym_begin /*T+3*/, size_t* asym_begin /*T+3*/) {
    Plan* p = static_cast<Plan*>(plan);
    std::copy(p->sym_begin.begin(), p->sym_begin.end(), sym_begin);
    std::copy(p->asym_begin.begin(), p->asym_begin.end(), asym_begin);
    return p->leg_sym.size() + p->leg_asym.size();
}
const double* orc_plan_leg_sym(void* plan) { return static_cast<Plan*>(plan)->leg_sym.data(); }
const double* orc_plan_leg_asym(void* plan) { return static_cast<Plan*>(plan)->leg_asym.data(); }
size_t orc_plan_npts(void* plan) { return static_cast<Plan*>(plan)->npts; }

// mode: 0 = literal (single thread, naive GEMM, naive DFT)   "reference as written"
//       1 = literal GEMM + FFT
//       2 = fast (OpenMP over m / (fld,lat), blocked GEMM, FFT)   CPU baseline
void orc_invtrans(void* plan, int nb_scalar, const double* scalar_spectra, int nb_vordiv, const double* vor,
                  const double* div, double* gp, int mode) {
    invtrans_full(*static_cast<Plan*>(plan), nb_scalar, scalar_spectra, nb_vordiv, vor, div, gp, mode == 2, mode == 0);
}

// Legendre stage only: scl_fourier[imag + 2*(m + (T+1)*(lat + nlat*fld))]  (TransLocal.h:177-180)
void orc_invtrans_legendre(void* plan, int truncation, int nb_fields, const double* spectra, double* scl_fourier,
                           int fast) {
    Plan* p = static_cast<Plan*>(plan);
    std::fill(scl_fourier, scl_fourier + static_cast<size_t>(nb_fields) * 2 * p->nlat * (p->T + 1), 0.);
    invtrans_legendre(*p, truncation, p->nlat, nb_fields, spectra, scl_fourier, fast != 0);
}

void orc_vd2uv(int truncation, int nb_fields, const double* vor, const double* div, double* U, double* V) {
    vd2uv(truncation, nb_fields, vor, div, U, V);
}

void orc_extend_truncation(int old_truncation, int nb_fields, const double* old_sp, double* new_sp) {
    extend_truncation(old_truncation, nb_fields, old_sp, new_sp);
}

void orc_dirtrans(void* plan, int nb_fields, const double* gp, double* spectra) {
    dirtrans(*static_cast<Plan*>(plan), nb_fields, gp, spectra);
}

void orc_dirtrans_wind(void* plan, int nb_fields, const double* wind, double* vor, double* div) {
    dirtrans_wind(*static_cast<Plan*>(plan), nb_fields, wind, vor, div);
}

void orc_invtrans_grad(void* plan, int nb_fields, const double* spectra, double* grad) {
    invtrans_grad(*static_cast<Plan*>(plan), nb_fields, spectra, grad, true);
}

// FFT self-checks
// TransLocal on an UnstructuredGrid; `truncation` is that of the data (T for scalar calls, T+1 after extend_truncation +
// vd2uv for the wind path, TransLocal.cc:1556-1589)
void orc_invtrans_unstructured(int truncation, int nb_fields, int nb_vordiv_fields, const double* spectra, int npts,
                               const double* lon_deg, const double* lat_deg, double* gp, int nthreads) {
    invtrans_unstructured(truncation, nb_fields, nb_vordiv_fields, spectra, npts, lon_deg, lat_deg, gp, nthreads);
}

void orc_c2r(int n, const double* in_complex, double* out, int naive) {
    const cplx* in = reinterpret_cast<const cplx*>(in_complex);
    if (naive) {
        c2r_naive(n, in, out);
        return;
    }
    RealFFT* r = rfft_plan(n);
    std::vector<cplx> buf(2 * static_cast<size_t>(n));
    c2r_fft(*r, in, out, buf.data());
    delete r;
}
void orc_r2c(int n, const double* in, double* out_complex) {
    RealFFT* r = rfft_plan(n);
    std::vector<cplx> buf(2 * static_cast<size_t>(n));
    r2c_fft(*r, in, reinterpret_cast<cplx*>(out_complex), buf.data());
    delete r;
}

}  // extern "C"
