// Spectral-space operators of the wind path.
//
// vd2uv: (vorticity, divergence) -> (U, V) = (u, v) cos(lat), Temperton 1991 eq. 2.12/2.13.  Replaces
// vd2uv + prfi1b (ecmwf/atlas src/atlas/trans/local/VorDivToUVLocal.cc:31-184), which copies each zonal
// wavenumber into four freshly allocated work vectors per m; here every output coefficient is one thread:
//
//   U_n^m = [ i m lap(n) D_n^m + (n-1) eps(n,m) lap(n-1) zeta_{n-1}^m - (n+2) eps(n+1,m) lap(n+1) zeta_{n+1}^m ] / a
//   V_n^m = [ i m lap(n) zeta_n^m - (n-1) eps(n,m) lap(n-1) D_{n-1}^m + (n+2) eps(n+1,m) lap(n+1) D_{n+1}^m ] / a
//   lap(n) = -a^2 / (n (n+1)),  lap(0) = 0;   eps(n,m) = sqrt((n^2-m^2)/(4n^2-1)),  eps(0,0) = 0
//
// with coefficients outside m <= n <= T taken as zero, a = 6371229 m (util/Earth.h:24).
// The products are associated exactly as in the reference (:135-137, :146-155, :172-178).
//
// extend_truncation (TransLocal.cc:1496-1519) and the interleave of [U | V | scalars] into one field-major
// spectral array (:1561-1582) are fused into one gather kernel.
#include "plan.hpp"

namespace sptrans {

namespace {

constexpr double kEarthRadius = 6371229.;

__device__ __forceinline__ double lapin(int n) {
    return n > 0 ? -kEarthRadius * kEarthRadius / (n * (n + 1.)) : 0.;
}
__device__ __forceinline__ double epsnm(int n, int m) {
    if (n == 0 || n < m) return 0.;
    return sqrt((static_cast<double>(n) * n - static_cast<double>(m) * m) / (4. * n * n - 1.));
}

// in/out layout [m][n][re/im][fld] at truncation T; one thread per (coefficient, field)
__global__ void vd2uv_kernel(int T, int nf, const double* __restrict__ vor, const double* __restrict__ div,
                             double* __restrict__ U, double* __restrict__ V) {
    const long long ncoef = static_cast<long long>(T + 1) * (T + 2) / 2;
    const long long total = ncoef * nf;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(e % nf);
        const long long c = e / nf;
        // invert c = (2T+3-m) m / 2 + (n - m)
        int m = static_cast<int>(((2.0 * T + 3.0) - sqrt((2.0 * T + 3.0) * (2.0 * T + 3.0) - 8.0 * static_cast<double>(c))) * 0.5);
        while (static_cast<long long>(2 * T + 3 - m) * m / 2 > c) --m;
        while (static_cast<long long>(2 * T + 3 - (m + 1)) * (m + 1) / 2 <= c) ++m;
        const int n = m + static_cast<int>(c - static_cast<long long>(2 * T + 3 - m) * m / 2);
        const long long ire = (2 * c) * nf + f, iim = ire + nf;
        const double chiIm = m * lapin(n);
        const double psiM1 = (n - 1) * epsnm(n, m) * lapin(n - 1);
        const double psiP1 = (n + 2) * epsnm(n + 1, m) * lapin(n + 1);
        double vm_r = 0., vm_i = 0., dm_r = 0., dm_i = 0., vp_r = 0., vp_i = 0., dp_r = 0., dp_i = 0.;
        if (n - 1 >= m) {
            const long long j = (2 * (c - 1)) * nf + f;
            vm_r = vor[j]; vm_i = vor[j + nf]; dm_r = div[j]; dm_i = div[j + nf];
        }
        if (n + 1 <= T) {
            const long long j = (2 * (c + 1)) * nf + f;
            vp_r = vor[j]; vp_i = vor[j + nf]; dp_r = div[j]; dp_i = div[j + nf];
        }
        const double v_r = vor[ire], v_i = vor[iim], d_r = div[ire], d_i = div[iim];
        const double za_r = 1. / kEarthRadius;
        double ur, ui, vr, vi;
        if (m == 0) {
            ur = +psiM1 * vm_r - psiP1 * vp_r;
            vr = -psiM1 * dm_r + psiP1 * dp_r;
            ui = 0.;  // the reference leaves the (unused) imaginary part of m = 0 at zero (:133-142)
            vi = 0.;
        }
        else {
            ur = -chiIm * d_i + psiM1 * vm_r - psiP1 * vp_r;
            ui = +chiIm * d_r + psiM1 * vm_i - psiP1 * vp_i;
            vr = -chiIm * v_i - psiM1 * dm_r + psiP1 * dp_r;
            vi = +chiIm * v_r - psiM1 * dm_i + psiP1 * dp_i;
        }
        U[ire] = ur * za_r;
        U[iim] = ui * za_r;
        V[ire] = vr * za_r;
        V[iim] = vi * za_r;
    }
}

// Build the merged spectral array at truncation T+1 with fields [U_1..U_k | V_1..V_k | s_1..s_j]
// from vor/div/scalars given at truncation T:  zero padding of extend_truncation + vd2uv at T+1 +
// interleave, all in one pass (reference :1540-1582 does five passes over host vectors).
__global__ void merge_uv_scalar_kernel(int T, int nvd, int nsc, const double* __restrict__ vor,
                                       const double* __restrict__ div, const double* __restrict__ sc,
                                       double* __restrict__ all) {
    const int Te = T + 1;
    const int nall = 2 * nvd + nsc;
    const long long ncoef_e = static_cast<long long>(Te + 1) * (Te + 2) / 2;
    const long long total = ncoef_e * nall;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(e % nall);
        const long long c = e / nall;  // coefficient index at truncation Te
        int m = static_cast<int>(((2.0 * Te + 3.0) - sqrt((2.0 * Te + 3.0) * (2.0 * Te + 3.0) - 8.0 * static_cast<double>(c))) * 0.5);
        while (static_cast<long long>(2 * Te + 3 - m) * m / 2 > c) --m;
        while (static_cast<long long>(2 * Te + 3 - (m + 1)) * (m + 1) / 2 <= c) ++m;
        const int n = m + static_cast<int>(c - static_cast<long long>(2 * Te + 3 - m) * m / 2);
        // index of (m, n') at the data truncation T, valid if m <= T and n' <= T
        auto at = [&](const double* a, int nfld, int fld, int nn, int imag) -> double {
            if (m > T || nn > T || nn < m) return 0.;
            const long long ct = static_cast<long long>(2 * T + 3 - m) * m / 2 + (nn - m);
            return a[(2 * ct + imag) * nfld + fld];
        };
        double out_r, out_i;
        if (f >= 2 * nvd) {
            const int fs = f - 2 * nvd;
            out_r = at(sc, nsc, fs, n, 0);
            out_i = at(sc, nsc, fs, n, 1);
        }
        else {
            const bool isU = f < nvd;
            const int fv = isU ? f : f - nvd;
            const double chiIm = m * lapin(n);
            const double psiM1 = (n - 1) * epsnm(n, m) * lapin(n - 1);
            const double psiP1 = (n + 2) * epsnm(n + 1, m) * lapin(n + 1);
            // U uses (D_n, zeta_{n-1}, zeta_{n+1}); V uses (zeta_n, D_{n-1}, D_{n+1}) with opposite stencil signs
            const double* A = isU ? div : vor;   // the "chi" operand at n
            const double* Bq = isU ? vor : div;  // the stencil operand at n-1, n+1
            const double sgn = isU ? 1. : -1.;
            const double a_r = at(A, nvd, fv, n, 0), a_i = at(A, nvd, fv, n, 1);
            const double bm_r = at(Bq, nvd, fv, n - 1, 0), bm_i = at(Bq, nvd, fv, n - 1, 1);
            const double bp_r = at(Bq, nvd, fv, n + 1, 0), bp_i = at(Bq, nvd, fv, n + 1, 1);
            const double za_r = 1. / kEarthRadius;
            if (m == 0) {
                out_r = (sgn * psiM1 * bm_r - sgn * psiP1 * bp_r) * za_r;
                out_i = 0.;
            }
            else {
                out_r = (-chiIm * a_i + sgn * psiM1 * bm_r - sgn * psiP1 * bp_r) * za_r;
                out_i = (+chiIm * a_r + sgn * psiM1 * bm_i - sgn * psiP1 * bp_i) * za_r;
            }
        }
        all[(2 * c) * nall + f] = out_r;
        all[(2 * c + 1) * nall + f] = out_i;
    }
}

// Spectral part of invtrans_grad: from scalar coefficients X at truncation T build, at truncation T+1, the fields
//   [E-W_1..E-W_k | N-S_1..N-S_k],  E-W_n^m = i m X_n^m / a,
//   N-S_n^m = [ (n+2) eps(n+1,m) X_{n+1}^m - (n-1) eps(n,m) X_{n-1}^m ] / a          ((1-mu^2) dPbar/dmu recurrence)
// whose inverse transform, divided by cos(lat) in the Fourier store, is (1/(a cos)) df/dlambda and (1/a) df/dphi.
// This is vd2uv applied to the velocity potential chi = f (vor = 0), cf. VorDivToUVLocal.cc:133-157.
__global__ void grad_spectra_kernel(int T, int nf, const double* __restrict__ sp, double* __restrict__ all) {
    const int Te = T + 1;
    const int nall = 2 * nf;
    const long long ncoef_e = static_cast<long long>(Te + 1) * (Te + 2) / 2;
    const long long total = ncoef_e * nall;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(e % nall);
        const long long c = e / nall;
        int m = static_cast<int>(((2.0 * Te + 3.0) - sqrt((2.0 * Te + 3.0) * (2.0 * Te + 3.0) - 8.0 * static_cast<double>(c))) * 0.5);
        while (static_cast<long long>(2 * Te + 3 - m) * m / 2 > c) --m;
        while (static_cast<long long>(2 * Te + 3 - (m + 1)) * (m + 1) / 2 <= c) ++m;
        const int n = m + static_cast<int>(c - static_cast<long long>(2 * Te + 3 - m) * m / 2);
        const int fs = f < nf ? f : f - nf;
        auto at = [&](int nn, int imag) -> double {
            if (m > T || nn > T || nn < m) return 0.;
            const long long ct = static_cast<long long>(2 * T + 3 - m) * m / 2 + (nn - m);
            return sp[(2 * ct + imag) * nf + fs];
        };
        const double inv_a = 1. / kEarthRadius;
        double out_r, out_i;
        if (f < nf) {  // E-W: i m X
            out_r = -m * at(n, 1) * inv_a;
            out_i = m * at(n, 0) * inv_a;
        }
        else {
            const double cp = (n + 2) * epsnm(n + 1, m), cm = (n - 1) * epsnm(n, m);
            out_r = (cp * at(n + 1, 0) - cm * at(n - 1, 0)) * inv_a;
            out_i = (cp * at(n + 1, 1) - cm * at(n - 1, 1)) * inv_a;
        }
        if (m == 0) out_i = 0.;
        all[(2 * c) * nall + f] = out_r;
        all[(2 * c + 1) * nall + f] = out_i;
    }
}

// Spectral part of dirtrans(wind -> vor, div): input = packed scalar transforms (to n = T+1) of
// Ut = u/(a cos), Vt = v/(a cos) as produced by the direct Legendre kernel, layout [m][parity][k][2 fld + re/im]
// with fields [u_1..u_k | v_1..v_k]:
//   zeta_n^m = i m Vt_n^m + (n+1) eps(n,m) Ut_{n-1}^m - n eps(n+1,m) Ut_{n+1}^m
//   D_n^m    = i m Ut_n^m - (n+1) eps(n,m) Vt_{n-1}^m + n eps(n+1,m) Vt_{n+1}^m
__global__ void uv_to_vordiv_kernel(int T, int nf, const long long* __restrict__ sp_rowoff,
                                    const double* __restrict__ packed, double* __restrict__ vor,
                                    double* __restrict__ div) {
    const long long ncoef = static_cast<long long>(T + 1) * (T + 2) / 2;
    const long long total = ncoef * nf;
    const int ld = 4 * nf;  // 2 * (2 nf) columns
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(e % nf);
        const long long c = e / nf;
        int m = static_cast<int>(((2.0 * T + 3.0) - sqrt((2.0 * T + 3.0) * (2.0 * T + 3.0) - 8.0 * static_cast<double>(c))) * 0.5);
        while (static_cast<long long>(2 * T + 3 - m) * m / 2 > c) --m;
        while (static_cast<long long>(2 * T + 3 - (m + 1)) * (m + 1) / 2 <= c) ++m;
        const int n = m + static_cast<int>(c - static_cast<long long>(2 * T + 3 - m) * m / 2);
        auto ext = [&](int nn, int imag, int fld) -> double {
            if (nn < m || nn > T + 1) return 0.;
            const int p = (nn - m) & 1, k = (nn - m) >> 1;
            return packed[(sp_rowoff[2 * m + p] + k) * ld + 2 * fld + imag];
        };
        const int fu = f, fv = nf + f;
        const double em = (n + 1) * epsnm(n, m), ep = n * epsnm(n + 1, m);
        double zr = -m * ext(n, 1, fv) + em * ext(n - 1, 0, fu) - ep * ext(n + 1, 0, fu);
        double zi = +m * ext(n, 0, fv) + em * ext(n - 1, 1, fu) - ep * ext(n + 1, 1, fu);
        double dr = -m * ext(n, 1, fu) - em * ext(n - 1, 0, fv) + ep * ext(n + 1, 0, fv);
        double di = +m * ext(n, 0, fu) - em * ext(n - 1, 1, fv) + ep * ext(n + 1, 1, fv);
        if (m == 0) zi = di = 0.;
        vor[(2 * c) * nf + f] = zr;
        vor[(2 * c + 1) * nf + f] = zi;
        div[(2 * c) * nf + f] = dr;
        div[(2 * c + 1) * nf + f] = di;
    }
}


// Adjoint (transpose over the reals) of merge_uv_scalar_kernel: from the adjoint variables of the merged fields
// [U_1..U_k | V_1..V_k | s_1..s_j] at truncation T+1 -- read straight from the packed output of the direct Legendre
// kernel, layout [m][parity][k][2 fld + re/im] -- to the adjoint variables of vor, div and the scalars at truncation T.
// With c(n) = m lap(n), pm(n) = (n-1) eps(n,m) lap(n-1), pp(n) = (n+2) eps(n+1,m) lap(n+1) (the forward stencil above):
//   zeta^_r(n) = [ pm(n+1) U^_r(n+1) - pp(n-1) U^_r(n-1) + c(n) V^_i(n) ] / a
//   zeta^_i(n) = [ pm(n+1) U^_i(n+1) - pp(n-1) U^_i(n-1) - c(n) V^_r(n) ] / a
//   D^_r(n)    = [-pm(n+1) V^_r(n+1) + pp(n-1) V^_r(n-1) + c(n) U^_i(n) ] / a
//   D^_i(n)    = [-pm(n+1) V^_i(n+1) + pp(n-1) V^_i(n-1) - c(n) U^_r(n) ] / a
// and the imaginary parts at m = 0, which the forward operator ignores, get zero.
__global__ void merge_uv_scalar_adj_kernel(int T, int nvd, int nsc, const long long* __restrict__ sp_rowoff,
                                           const double* __restrict__ packed, double* __restrict__ vor,
                                           double* __restrict__ div, double* __restrict__ sc) {
    const long long ncoef = static_cast<long long>(T + 1) * (T + 2) / 2;
    const int nout = nvd + nsc;
    const long long total = ncoef * nout;
    const int ld = 2 * (2 * nvd + nsc);
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(e % nout);
        const long long c = e / nout;
        int m = static_cast<int>(((2.0 * T + 3.0) - sqrt((2.0 * T + 3.0) * (2.0 * T + 3.0) - 8.0 * static_cast<double>(c))) * 0.5);
        while (static_cast<long long>(2 * T + 3 - m) * m / 2 > c) --m;
        while (static_cast<long long>(2 * T + 3 - (m + 1)) * (m + 1) / 2 <= c) ++m;
        const int n = m + static_cast<int>(c - static_cast<long long>(2 * T + 3 - m) * m / 2);
        auto ext = [&](int nn, int imag, int fld) -> double {
            if (nn < m || nn > T + 1) return 0.;
            const int p = (nn - m) & 1, k = (nn - m) >> 1;
            return packed[(sp_rowoff[2 * m + p] + k) * ld + 2 * fld + imag];
        };
        if (f >= nvd) {
            const int fs = f - nvd, fm = 2 * nvd + fs;
            sc[(2 * c) * nsc + fs] = ext(n, 0, fm);
            sc[(2 * c + 1) * nsc + fs] = m == 0 ? 0. : ext(n, 1, fm);
            continue;
        }
        const int fu = f, fv = nvd + f;
        const double za_r = 1. / kEarthRadius;
        const double cn = m * lapin(n);
        const double pmn = n * epsnm(n + 1, m) * lapin(n);        // pm(n+1)
        const double ppn = (n + 1) * epsnm(n, m) * lapin(n);      // pp(n-1)
        double zr = (pmn * ext(n + 1, 0, fu) - ppn * ext(n - 1, 0, fu) + cn * ext(n, 1, fv)) * za_r;
        double zi = (pmn * ext(n + 1, 1, fu) - ppn * ext(n - 1, 1, fu) - cn * ext(n, 0, fv)) * za_r;
        double dr = (-pmn * ext(n + 1, 0, fv) + ppn * ext(n - 1, 0, fv) + cn * ext(n, 1, fu)) * za_r;
        double di = (-pmn * ext(n + 1, 1, fv) + ppn * ext(n - 1, 1, fv) - cn * ext(n, 0, fu)) * za_r;
        if (m == 0) zi = di = 0.;
        vor[(2 * c) * nvd + f] = zr;
        vor[(2 * c + 1) * nvd + f] = zi;
        div[(2 * c) * nvd + f] = dr;
        div[(2 * c + 1) * nvd + f] = di;
    }
}

// Adjoint of grad_spectra_kernel: packed adjoint variables of [E-W_1..E-W_k | N-S_1..N-S_k] at T+1 -> scalar adjoint at T.
// With cp(n) = (n+2) eps(n+1,m), cm(n) = (n-1) eps(n,m):
//   X^_r(n) = [ +m EW^_i(n) + cp(n-1) NS^_r(n-1) - cm(n+1) NS^_r(n+1) ] / a
//   X^_i(n) = [ -m EW^_r(n) + cp(n-1) NS^_i(n-1) - cm(n+1) NS^_i(n+1) ] / a      (zero at m = 0)
__global__ void grad_spectra_adj_kernel(int T, int nf, const long long* __restrict__ sp_rowoff,
                                        const double* __restrict__ packed, double* __restrict__ sp) {
    const long long ncoef = static_cast<long long>(T + 1) * (T + 2) / 2;
    const long long total = ncoef * nf;
    const int ld = 4 * nf;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(e % nf);
        const long long c = e / nf;
        int m = static_cast<int>(((2.0 * T + 3.0) - sqrt((2.0 * T + 3.0) * (2.0 * T + 3.0) - 8.0 * static_cast<double>(c))) * 0.5);
        while (static_cast<long long>(2 * T + 3 - m) * m / 2 > c) --m;
        while (static_cast<long long>(2 * T + 3 - (m + 1)) * (m + 1) / 2 <= c) ++m;
        const int n = m + static_cast<int>(c - static_cast<long long>(2 * T + 3 - m) * m / 2);
        auto ext = [&](int nn, int imag, int fld) -> double {
            if (nn < m || nn > T + 1) return 0.;
            const int p = (nn - m) & 1, k = (nn - m) >> 1;
            return packed[(sp_rowoff[2 * m + p] + k) * ld + 2 * fld + imag];
        };
        const int few = f, fns = nf + f;
        const double inv_a = 1. / kEarthRadius;
        const double cpm = (n + 1) * epsnm(n, m);      // cp(n-1)
        const double cmp = n * epsnm(n + 1, m);        // cm(n+1)
        double xr = (+m * ext(n, 1, few) + cpm * ext(n - 1, 0, fns) - cmp * ext(n + 1, 0, fns)) * inv_a;
        double xi = (-m * ext(n, 0, few) + cpm * ext(n - 1, 1, fns) - cmp * ext(n + 1, 1, fns)) * inv_a;
        if (m == 0) xi = 0.;
        sp[(2 * c) * nf + f] = xr;
        sp[(2 * c + 1) * nf + f] = xi;
    }
}

// Transpose of uv_to_vordiv_kernel (spectral part of dirtrans_wind2vordiv_adj, TransImpl.h:69-70): adjoint variables of
// (vor, div) at truncation T -> adjoint variables of the scaled wind transforms [Ut_1..Ut_k | Vt_1..Vt_k] at truncation
// T+1 in the [m][n][re/im][field] layout the pack kernel reads.  With em(n) = (n+1) eps(n,m), ep(n) = n eps(n+1,m):
//   Ut^_r(n) = em(n+1) z^_r(n+1) - ep(n-1) z^_r(n-1) + m d^_i(n)      Ut^_i(n) = em(n+1) z^_i(n+1) - ep(n-1) z^_i(n-1) - m d^_r(n)
//   Vt^_r(n) = -em(n+1) d^_r(n+1) + ep(n-1) d^_r(n-1) + m z^_i(n)     Vt^_i(n) = -em(n+1) d^_i(n+1) + ep(n-1) d^_i(n-1) - m z^_r(n)
// (z^, d^ vanish outside m <= n <= T; the imaginary parts at m = 0 do not enter, the forward operator sets them to zero).
__global__ void uv_to_vordiv_adj_kernel(int T, int nf, const double* __restrict__ vor, const double* __restrict__ div,
                                        double* __restrict__ all) {
    const int Te = T + 1;
    const int nall = 2 * nf;
    const long long ncoef_e = static_cast<long long>(Te + 1) * (Te + 2) / 2;
    const long long total = ncoef_e * nall;
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int f = static_cast<int>(e % nall);
        const long long c = e / nall;
        int m = static_cast<int>(((2.0 * Te + 3.0) - sqrt((2.0 * Te + 3.0) * (2.0 * Te + 3.0) - 8.0 * static_cast<double>(c))) * 0.5);
        while (static_cast<long long>(2 * Te + 3 - m) * m / 2 > c) --m;
        while (static_cast<long long>(2 * Te + 3 - (m + 1)) * (m + 1) / 2 <= c) ++m;
        const int n = m + static_cast<int>(c - static_cast<long long>(2 * Te + 3 - m) * m / 2);
        const bool isU = f < nf;
        const int fs = isU ? f : f - nf;
        auto at = [&](const double* a, int nn, int imag) -> double {
            if (m > T || nn > T || nn < m || (m == 0 && imag)) return 0.;
            const long long ct = static_cast<long long>(2 * T + 3 - m) * m / 2 + (nn - m);
            return a[(2 * ct + imag) * nf + fs];
        };
        const double emp = (n + 2) * epsnm(n + 1, m);  // em(n+1)
        const double epm = (n - 1) * epsnm(n, m);      // ep(n-1)
        const double* S = isU ? vor : div;             // stencil operand
        const double* X = isU ? div : vor;             // i m operand
        const double sg = isU ? 1. : -1.;
        double out_r = sg * (emp * at(S, n + 1, 0) - epm * at(S, n - 1, 0)) + m * at(X, n, 1);
        double out_i = sg * (emp * at(S, n + 1, 1) - epm * at(S, n - 1, 1)) - m * at(X, n, 0);
        if (m == 0) out_i = 0.;
        all[(2 * c) * nall + f] = out_r;
        all[(2 * c + 1) * nall + f] = out_i;
    }
}

}  // namespace

int launch_uv_to_vordiv_adj(cudaStream_t s, int T, int nf, const double* d_vor, const double* d_div, double* d_all,
                            uint64_t* launches) {
    const long long total = static_cast<long long>(T + 2) * (T + 3) / 2 * (2 * nf);
    if (total == 0) return SPTRANS_OK;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    uv_to_vordiv_adj_kernel<<<blocks, 256, 0, s>>>(T, nf, d_vor, d_div, d_all);
    if (launches) ++*launches;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int launch_merge_uv_scalar_adj(cudaStream_t s, int T, int nvd, int nsc, const long long* d_sp_rowoff, const double* d_packed,
                               double* d_vor, double* d_div, double* d_sc, uint64_t* launches) {
    const long long total = static_cast<long long>(T + 1) * (T + 2) / 2 * (nvd + nsc);
    if (total == 0) return SPTRANS_OK;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    merge_uv_scalar_adj_kernel<<<blocks, 256, 0, s>>>(T, nvd, nsc, d_sp_rowoff, d_packed, d_vor, d_div, d_sc);
    if (launches) ++*launches;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int launch_grad_spectra_adj(cudaStream_t s, int T, int nf, const long long* d_sp_rowoff, const double* d_packed, double* d_sp,
                            uint64_t* launches) {
    const long long total = static_cast<long long>(T + 1) * (T + 2) / 2 * nf;
    if (total == 0) return SPTRANS_OK;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    grad_spectra_adj_kernel<<<blocks, 256, 0, s>>>(T, nf, d_sp_rowoff, d_packed, d_sp);
    if (launches) ++*launches;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int launch_grad_spectra(cudaStream_t s, int T, int nf, const double* d_sp, double* d_all, uint64_t* launches) {
    const long long total = static_cast<long long>(T + 2) * (T + 3) / 2 * (2 * nf);
    if (total == 0) return SPTRANS_OK;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    grad_spectra_kernel<<<blocks, 256, 0, s>>>(T, nf, d_sp, d_all);
    if (launches) ++*launches;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int launch_uv_to_vordiv(cudaStream_t s, int T, int nf, const long long* d_sp_rowoff, const double* d_packed,
                        double* d_vor, double* d_div, uint64_t* launches) {
    const long long total = static_cast<long long>(T + 1) * (T + 2) / 2 * nf;
    if (total == 0) return SPTRANS_OK;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    uv_to_vordiv_kernel<<<blocks, 256, 0, s>>>(T, nf, d_sp_rowoff, d_packed, d_vor, d_div);
    if (launches) ++*launches;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int launch_vd2uv(cudaStream_t s, int T, int nf, const double* d_vor, const double* d_div, double* d_U, double* d_V,
                 uint64_t* launches) {
    const long long total = static_cast<long long>(T + 1) * (T + 2) / 2 * nf;
    if (total == 0) return SPTRANS_OK;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    vd2uv_kernel<<<blocks, 256, 0, s>>>(T, nf, d_vor, d_div, d_U, d_V);
    if (launches) ++*launches;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int launch_merge_uv_scalar(cudaStream_t s, int T, int nvd, int nsc, const double* d_vor, const double* d_div,
                           const double* d_sc, double* d_all, uint64_t* launches) {
    const long long total = static_cast<long long>(T + 2) * (T + 3) / 2 * (2 * nvd + nsc);
    if (total == 0) return SPTRANS_OK;
    const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
    merge_uv_scalar_kernel<<<blocks, 256, 0, s>>>(T, nvd, nsc, d_vor, d_div, d_sc, d_all);
    if (launches) ++*launches;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

}  // namespace sptrans
