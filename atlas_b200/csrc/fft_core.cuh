// Core of the shared-memory Fourier kernels: in-place power-of-two FFT passes and the chirp-z
// (Bluestein) index/phase algebra.  Everything here is __host__ __device__ so that the exact same
// index math is unit-tested on the CPU (tests/cpu/test_fft_core.cc) before it ever runs on a GPU.
//
// Why chirp-z: the octahedral grid has row lengths n = 20 + 4 j (ecmwf/atlas grid/detail/grid/Gaussian.cc:
// 127-134); at O1280 85 % of the rows (by points) contain a prime factor > 13, so a mixed-radix FFT alone
// cannot serve them.  A row pair (north, south) is one complex sequence z = x_N + i x_S with
//     z_i = sum_{m=-L..L} Z_m e^{+2 pi i m i / n},           L = zonal truncation at this latitude,
// evaluated as a length-M cyclic convolution, M = 2^a >= n + 2L:
//     z_i = C_i * sum_u (Z_{u-L} A_u) b_{i-u},   A_u = e^{i pi u^2/n},  b_k = e^{-i pi k^2/n},
//     C_i = e^{i pi (i^2 - 2 L i)/n}.
// The forward FFT is decimation-in-frequency (natural in, digit-reversed out), the filter spectrum is
// stored in the same digit-reversed order, and the inverse is decimation-in-time with the passes
// reversed -- no reordering pass anywhere.  The direct transform uses the conjugate tables.
#pragma once

#ifdef __CUDACC__
#define SPT_HD __host__ __device__ __forceinline__
#else
#define SPT_HD inline
struct double2 {
    double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
#endif

#include "fft_consts.h"

namespace sptrans {
namespace fftc {

SPT_HD double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SPT_HD double2 cmulc(double2 a, double2 b) {  // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
SPT_HD double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
SPT_HD double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
SPT_HD double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
// multiply by -i (FWD) or +i (INV):  FWD: (x,y)->(y,-x)   INV: (x,y)->(-y,x)
template <bool FWD>
SPT_HD double2 rot90(double2 a) {
    return FWD ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}

// shared-memory index padding: one slot per 8 elements keeps every pass (unit stride and the
// strided tail passes) free of bank conflicts for 16-byte elements
SPT_HD int pad(int i) { return i + (i >> 3); }
SPT_HD int padded_len(int M) { return M + (M >> 3); }

// radix-R DFT on registers, sign -1 if FWD (e^{-2 pi i jk/R}) else +1
template <bool FWD>
SPT_HD void dft2(double2* v) {
    double2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <bool FWD>
SPT_HD void dft4(double2* v) {
    double2 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
    double2 s13 = cadd(v[1], v[3]), d13 = rot90<FWD>(csub(v[1], v[3]));
    v[0] = cadd(s02, s13);
    v[2] = csub(s02, s13);
    v[1] = cadd(d02, d13);
    v[3] = csub(d02, d13);
}
template <bool FWD>
SPT_HD void dft8(double2* v) {
    const double h = 0.70710678118654752440;
    // stage 1: pairs (k, k+4)
    double2 a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
    double2 a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
    double2 a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
    double2 a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
    // twiddle the odd half by w8^k
    // w8 = e^{-+ i pi/4}: FWD (1-i)/sqrt2 ; INV (1+i)/sqrt2
    double2 t1 = FWD ? make_double2((b1.x + b1.y) * h, (b1.y - b1.x) * h) : make_double2((b1.x - b1.y) * h, (b1.y + b1.x) * h);
    double2 t2 = rot90<FWD>(b2);
    double2 t3 = FWD ? make_double2((b3.y - b3.x) * h, -(b3.x + b3.y) * h) : make_double2(-(b3.x + b3.y) * h, (b3.x - b3.y) * h);
    // two radix-4 DFTs: even outputs from a*, odd outputs from (b0,t1,t2,t3)
    double2 e[4] = {a0, a1, a2, a3};
    double2 o[4] = {b0, t1, t2, t3};
    dft4<FWD>(e);
    dft4<FWD>(o);
    v[0] = e[0];
    v[2] = e[1];
    v[4] = e[2];
    v[6] = e[3];
    v[1] = o[0];
    v[3] = o[1];
    v[5] = o[2];
    v[7] = o[3];
}
template <int R, bool FWD>
SPT_HD void dftR(double2* v) {
    if (R == 2) dft2<FWD>(v);
    else if (R == 4) dft4<FWD>(v);
    else dft8<FWD>(v);
}

// One butterfly of a decimation-in-frequency pass (natural -> digit-reversed), sign -1:
//   elements X[base + t + j*Nb/R], y = DFT_R(x), y_j *= w_Nb^{j t},  w_Nb = e^{-2 pi i/Nb}
// `tw` is the master table W[k] = e^{-2 pi i k / Wn}, Wn a multiple of Nb.
template <int R>
SPT_HD void dif_butterfly(double2* X, int q, int Nb, const double2* tw, int Wn) {
    const int span = Nb / R;
    const int blk = q / span, t = q - blk * span;
    const int base = blk * Nb + t;
    double2 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = X[pad(base + j * span)];
    dftR<R, true>(v);
    if (t != 0) {
        const double2 w1 = tw[t * (Wn / Nb)];
        double2 w = w1;
#pragma unroll
        for (int j = 1; j < R; ++j) {
            v[j] = cmul(v[j], w);
            if (j + 1 < R) w = cmul(w, w1);
        }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) X[pad(base + j * span)] = v[j];
}

// Inverse of dif_butterfly up to the factor R (decimation in time, sign +1): x_j = y_j * conj(w^{jt}),
// then DFT_R with sign +1.  If `filt` is non-null each loaded element is first multiplied by
// filt[index] (or its conjugate): this fuses the Bluestein pointwise product into the first inverse pass.
template <int R, bool CONJ_FILT>
SPT_HD void dit_butterfly(double2* X, int q, int Nb, const double2* tw, int Wn, const double2* filt) {
    const int span = Nb / R;
    const int blk = q / span, t = q - blk * span;
    const int base = blk * Nb + t;
    double2 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        v[j] = X[pad(base + j * span)];
        if (filt) v[j] = CONJ_FILT ? cmulc(v[j], filt[base + j * span]) : cmul(v[j], filt[base + j * span]);
    }
    if (t != 0) {
        const double2 w1 = tw[t * (Wn / Nb)];
        double2 w = w1;
#pragma unroll
        for (int j = 1; j < R; ++j) {
            v[j] = cmulc(v[j], w);
            if (j + 1 < R) w = cmul(w, w1);
        }
    }
    dftR<R, false>(v);
#pragma unroll
    for (int j = 0; j < R; ++j) X[pad(base + j * span)] = v[j];
}

// Pass schedule for M = 2^logM: radix 8 while possible, then one radix 4 or 2.
// DIF order: pass p has block length Nb_p = M / prod_{q<p} R_q.
struct Schedule {
    int npass;
    int radix[6];
    int nb[6];
};
SPT_HD Schedule make_schedule(int logM) {
    Schedule s;
    s.npass = 0;
    int rem = logM, Nb = 1 << logM;
    while (rem >= 3) {
        s.radix[s.npass] = 8;
        s.nb[s.npass] = Nb;
        ++s.npass;
        Nb >>= 3;
        rem -= 3;
    }
    if (rem == 2) {
        s.radix[s.npass] = 4;
        s.nb[s.npass] = Nb;
        ++s.npass;
    }
    else if (rem == 1) {
        s.radix[s.npass] = 2;
        s.nb[s.npass] = Nb;
        ++s.npass;
    }
    return s;
}


#if defined(__CUDA_ARCH__)
#define SPT_SYNC() __syncthreads()
#else
#define SPT_SYNC() ((void)0)
#endif

// Forward (DIF) transform of `nseq` sequences of length M stored back to back (each padded_len(M) slots).
// Cooperative over (tid, nthr); on the host nthr == 1 and SPT_SYNC is a no-op.
SPT_HD void fft_dif_all(double2* X, int nseq, int logM, const double2* tw, int Wn, int tid, int nthr) {
    const int M = 1 << logM;
    const Schedule s = make_schedule(logM);
    const int PL = padded_len(M);
    for (int p = 0; p < s.npass; ++p) {
        const int R = s.radix[p], Nb = s.nb[p];
        const int per = M / R;
        for (int w = tid; w < nseq * per; w += nthr) {
            const int sq = w / per, q = w - sq * per;
            double2* Xs = X + sq * PL;
            if (R == 8) dif_butterfly<8>(Xs, q, Nb, tw, Wn);
            else if (R == 4) dif_butterfly<4>(Xs, q, Nb, tw, Wn);
            else dif_butterfly<2>(Xs, q, Nb, tw, Wn);
        }
        SPT_SYNC();
    }
}

// Inverse (DIT) transform, unnormalised, with the pointwise filter product fused into its first pass.
template <bool CONJ_FILT>
SPT_HD void fft_dit_all(double2* X, int nseq, int logM, const double2* tw, int Wn, const double2* filt, int tid,
                        int nthr) {
    const int M = 1 << logM;
    const Schedule s = make_schedule(logM);
    const int PL = padded_len(M);
    for (int p = s.npass - 1; p >= 0; --p) {
        const int R = s.radix[p], Nb = s.nb[p];
        const int per = M / R;
        const double2* f = (p == s.npass - 1) ? filt : nullptr;
        for (int w = tid; w < nseq * per; w += nthr) {
            const int sq = w / per, q = w - sq * per;
            double2* Xs = X + sq * PL;
            if (R == 8) dit_butterfly<8, CONJ_FILT>(Xs, q, Nb, tw, Wn, f);
            else if (R == 4) dit_butterfly<4, CONJ_FILT>(Xs, q, Nb, tw, Wn, f);
            else dit_butterfly<2, CONJ_FILT>(Xs, q, Nb, tw, Wn, f);
        }
        SPT_SYNC();
    }
}


// =====================================================================================================
// Mixed-radix engine (M = 2^a 3^b 5^c): same in-place DIF / DIT scheme with radices {16, 9, 8, 5, 4, 3, 2}.
// Choosing M among all 5-smooth lengths instead of powers of two cuts the convolution length of the
// chirp-z by ~25 % on the octahedral grid (n + 2L is rarely just below a power of two).
// =====================================================================================================
template <bool FWD>
SPT_HD void dft3(double2* v) {
    const double c = 0.86602540378443864676;  // sin(pi/3)
    const double2 s = cadd(v[1], v[2]), d = csub(v[1], v[2]);
    const double2 m = make_double2(v[0].x - 0.5 * s.x, v[0].y - 0.5 * s.y);
    const double2 rd = rot90<FWD>(d);
    const double2 r = make_double2(c * rd.x, c * rd.y);
    v[0] = cadd(v[0], s);
    v[1] = cadd(m, r);
    v[2] = csub(m, r);
}
template <bool FWD>
SPT_HD void dft5(double2* v) {
    const double c1 = 0.30901699437494742410, c2 = -0.80901699437494742410;  // cos(2pi/5), cos(4pi/5)
    const double s1 = 0.95105651629515357212, s2 = 0.58778525229247312917;   // sin(2pi/5), sin(4pi/5)
    const double2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    const double2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    const double2 p1 = make_double2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    const double2 p2 = make_double2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    const double2 q1 = rot90<FWD>(make_double2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
    const double2 q2 = rot90<FWD>(make_double2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
    v[0] = make_double2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
    v[1] = cadd(p1, q1);
    v[4] = csub(p1, q1);
    v[2] = cadd(p2, q2);
    v[3] = csub(p2, q2);
}
// Shared-memory index swizzle of the mixed-radix engine (no padding): the three low index bits (= the 128-byte
// bank group of a 16-byte element) are XORed with bits 3..5 and 4..6.  Unit-stride octets stay conflict free,
// and so do the power-of-two strides 2, 4, 8, 16 of the last passes (span < 8); odd radices run first, where
// every span is a multiple of 8.  Requires M % 8 == 0.
SPT_HD int swz(int i) { return i ^ (((i >> 3) ^ (i >> 4)) & 7); }

// exact q / d for q < 65536 via a precomputed reciprocal (integer division is ~20 instructions on the GPU)
SPT_HD unsigned fastdiv_magic(unsigned d) { return static_cast<unsigned>((0x100000000ull + d - 1) / d); }
SPT_HD int fastdiv(int q, unsigned magic, int d) {
#ifdef __CUDA_ARCH__
    return d == 1 ? q : static_cast<int>(__umulhi(static_cast<unsigned>(q), magic));
#else
    return d == 1 ? q : static_cast<int>((static_cast<unsigned long long>(static_cast<unsigned>(q)) * magic) >> 32);
#endif
}

// multiply by e^{-+ 2 pi i k / N} given (cos, sin) of 2 pi k / N
template <bool FWD>
SPT_HD double2 twc(double2 a, double c, double s) {
    return FWD ? make_double2(a.x * c + a.y * s, a.y * c - a.x * s) : make_double2(a.x * c - a.y * s, a.y * c + a.x * s);
}
template <bool FWD>
SPT_HD void dft9(double2* v) {
    const double c1 = 0.76604444311897803520, s1 = 0.64278760968653932632;   // 2pi/9
    const double c2 = 0.17364817766693034885, s2 = 0.98480775301220805937;   // 4pi/9
    const double c4 = -0.93969262078590838405, s4 = 0.34202014332566873304;  // 8pi/9
    double2 y[3][3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
        double2 t[3] = {v[b], v[b + 3], v[b + 6]};
        dft3<FWD>(t);
        y[b][0] = t[0];
        y[b][1] = t[1];
        y[b][2] = t[2];
    }
    y[1][1] = twc<FWD>(y[1][1], c1, s1);
    y[1][2] = twc<FWD>(y[1][2], c2, s2);
    y[2][1] = twc<FWD>(y[2][1], c2, s2);
    y[2][2] = twc<FWD>(y[2][2], c4, s4);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double2 t[3] = {y[0][c], y[1][c], y[2][c]};
        dft3<FWD>(t);
        v[c] = t[0];
        v[c + 3] = t[1];
        v[c + 6] = t[2];
    }
}
template <bool FWD>
SPT_HD void dft16(double2* v) {
    const double h = 0.70710678118654752440;
    const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;  // 2pi/16
    double2 y[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        double2 t[4] = {v[b], v[b + 4], v[b + 8], v[b + 12]};
        dft4<FWD>(t);
        y[b][0] = t[0];
        y[b][1] = t[1];
        y[b][2] = t[2];
        y[b][3] = t[3];
    }
    // twiddles w16^{b c}
    y[1][1] = twc<FWD>(y[1][1], c1, s1);   // w^1
    y[1][2] = twc<FWD>(y[1][2], h, h);     // w^2
    y[1][3] = twc<FWD>(y[1][3], s1, c1);   // w^3
    y[2][1] = twc<FWD>(y[2][1], h, h);     // w^2
    y[2][2] = rot90<FWD>(y[2][2]);         // w^4
    y[2][3] = twc<FWD>(y[2][3], -h, h);    // w^6
    y[3][1] = twc<FWD>(y[3][1], s1, c1);   // w^3
    y[3][2] = twc<FWD>(y[3][2], -h, h);    // w^6
    y[3][3] = twc<FWD>(y[3][3], -c1, -s1); // w^9
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double2 t[4] = {y[0][c], y[1][c], y[2][c], y[3][c]};
        dft4<FWD>(t);
        v[c] = t[0];
        v[c + 4] = t[1];
        v[c + 8] = t[2];
        v[c + 12] = t[3];
    }
}
// Odd-prime radix P (7, 11, 13) for the direct mixed-radix transforms of rows whose length has these factors:
// y_0 = sum x_j;  y_k, y_{P-k} = x_0 + sum_j c_{jk} (x_j + x_{P-j})  -+  i sgn sum_j s_{jk} (x_j - x_{P-j}),  j, k = 1..(P-1)/2
// with c, s = cos, sin(2 pi j k / P): (P-1)^2 / 2 real multiply-adds per component instead of (P-1)^2 complex ones.
template <int P, bool FWD>
SPT_HD void dft_prime(double2* v) {
    constexpr int H = (P - 1) / 2;
    double2 a[H], b[H];
#pragma unroll
    for (int j = 1; j <= H; ++j) {
        a[j - 1] = cadd(v[j], v[P - j]);
        b[j - 1] = csub(v[j], v[P - j]);
    }
    double2 y0 = v[0];
#pragma unroll
    for (int j = 0; j < H; ++j) y0 = cadd(y0, a[j]);
    double2 out[P];
    out[0] = y0;
#pragma unroll
    for (int k = 1; k <= H; ++k) {
        double2 re = v[0], im = make_double2(0., 0.);
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            const double c = WConst<P>::c((j * k) % P), s = WConst<P>::s((j * k) % P);
            re.x += c * a[j - 1].x;
            re.y += c * a[j - 1].y;
            im.x += s * b[j - 1].x;
            im.y += s * b[j - 1].y;
        }
        // e^{-+ i theta}(x_j) + e^{+- i theta}(x_{P-j}) = cos (x_j + x_{P-j}) -+ i sin (x_j - x_{P-j}); -i (FWD) / +i (INV) times im
        const double2 r = rot90<FWD>(im);
        out[k] = cadd(re, r);
        out[P - k] = csub(re, r);
    }
#pragma unroll
    for (int k = 0; k < P; ++k) v[k] = out[k];
}
template <int R, bool FWD>
SPT_HD void dftN(double2* v) {
    if (R == 7 || R == 11 || R == 13) {
        dft_prime<(R == 7 || R == 11 || R == 13) ? R : 7, FWD>(v);
        return;
    }
    if (R == 2) dft2<FWD>(v);
    else if (R == 3) dft3<FWD>(v);
    else if (R == 4) dft4<FWD>(v);
    else if (R == 5) dft5<FWD>(v);
    else if (R == 8) dft8<FWD>(v);
    else if (R == 9) dft9<FWD>(v);
    else dft16<FWD>(v);
}

// powers w^1..w^{R-1} of one twiddle with a shallow dependency tree
template <int R>
SPT_HD void twiddle_powers(double2 w1, double2* w /*[R]*/) {
    w[1] = w1;
    if (R > 2) w[2] = cmul(w1, w1);
    if (R > 3) w[3] = cmul(w[2], w1);
    if (R > 4) w[4] = cmul(w[2], w[2]);
#pragma unroll
    for (int j = 5; j < R; ++j) w[j] = (j & 1) ? cmul(w[j - 1], w1) : cmul(w[j / 2], w[j / 2]);
}

// Two-level twiddle table of e^{-2 pi i k / M}: W[k] = Wa[k >> 6] * Wb[k & 63]  (Wa: M/64+1 entries, Wb: 64)
SPT_HD double2 twiddle2(const double2* Wa, const double2* Wb, int k) { return cmul(Wa[k >> 6], Wb[k & 63]); }

template <int R>
SPT_HD void dif_butterfly_g(double2* X, int q, int Nb, int span, unsigned span_magic, int S, const double2* Wa,
                            const double2* Wb) {
    const int blk = fastdiv(q, span_magic, span), t = q - blk * span;
    const int base = blk * Nb + t;
    double2 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = X[swz(base + j * span)];
    dftN<R, true>(v);
    if (t != 0) {
        double2 w[R];
        twiddle_powers<R>(twiddle2(Wa, Wb, t * S), w);
#pragma unroll
        for (int j = 1; j < R; ++j) v[j] = cmul(v[j], w[j]);
    }
#pragma unroll
    for (int j = 0; j < R; ++j) X[swz(base + j * span)] = v[j];
}

template <int R, bool CONJ_FILT>
SPT_HD void dit_butterfly_g(double2* X, int q, int Nb, int span, unsigned span_magic, int S, const double2* Wa,
                            const double2* Wb, const double2* filt) {
    const int blk = fastdiv(q, span_magic, span), t = q - blk * span;
    const int base = blk * Nb + t;
    double2 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        v[j] = X[swz(base + j * span)];
        if (filt) v[j] = CONJ_FILT ? cmulc(v[j], filt[base + j * span]) : cmul(v[j], filt[base + j * span]);
    }
    if (t != 0) {
        double2 w[R];
        twiddle_powers<R>(twiddle2(Wa, Wb, t * S), w);
#pragma unroll
        for (int j = 1; j < R; ++j) v[j] = cmulc(v[j], w[j]);
    }
    dftN<R, false>(v);
#pragma unroll
    for (int j = 0; j < R; ++j) X[swz(base + j * span)] = v[j];
}

// Large odd-prime radices (17, 19, 23) of the direct transforms.  Holding P inputs and P outputs would not fit the register
// budget, so the outputs are produced pair by pair (k, P-k) from the (P-1)/2 sums and differences x_j +- x_{P-j} and stored
// at once; the twiddles w^k are a running product, w^{P-k} = w^P conj(w^k) with w^P looked up as well.
template <int P>
SPT_HD void dif_butterfly_prime(double2* X, int q, int Nb, int span, unsigned span_magic, int S, const double2* Wa,
                                const double2* Wb) {
    constexpr int H = (P - 1) / 2;
    const int blk = fastdiv(q, span_magic, span), t = q - blk * span;
    const int base = blk * Nb + t;
    const double2 x0 = X[swz(base)];
    double2 a[H], b[H];
    double2 y0 = x0;
#pragma unroll
    for (int j = 1; j <= H; ++j) {
        const double2 u = X[swz(base + j * span)], w = X[swz(base + (P - j) * span)];
        a[j - 1] = cadd(u, w);
        b[j - 1] = csub(u, w);
        y0 = cadd(y0, a[j - 1]);
    }
    X[swz(base)] = y0;
    double2 w1 = make_double2(1., 0.), wP = make_double2(1., 0.), wk = make_double2(1., 0.);
    if (t != 0) {
        w1 = twiddle2(Wa, Wb, t * S);
        wP = twiddle2(Wa, Wb, t * S * P);
    }
#pragma unroll
    for (int k = 1; k <= H; ++k) {
        double2 re = x0, im = make_double2(0., 0.);
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            const double c = WConst<P>::c((j * k) % P), s = WConst<P>::s((j * k) % P);
            re.x += c * a[j - 1].x;
            re.y += c * a[j - 1].y;
            im.x += s * b[j - 1].x;
            im.y += s * b[j - 1].y;
        }
        const double2 r = rot90<true>(im);
        double2 yk = cadd(re, r), yq = csub(re, r);
        if (t != 0) {
            wk = cmul(wk, w1);
            yk = cmul(yk, wk);
            yq = cmul(yq, cmulc(wP, wk));
        }
        X[swz(base + k * span)] = yk;
        X[swz(base + (P - k) * span)] = yq;
    }
}
template <int P>
SPT_HD void dit_butterfly_prime(double2* X, int q, int Nb, int span, unsigned span_magic, int S, const double2* Wa,
                                const double2* Wb) {
    constexpr int H = (P - 1) / 2;
    const int blk = fastdiv(q, span_magic, span), t = q - blk * span;
    const int base = blk * Nb + t;
    const double2 x0 = X[swz(base)];
    double2 w1 = make_double2(1., 0.), wP = make_double2(1., 0.), wj = make_double2(1., 0.);
    if (t != 0) {
        w1 = twiddle2(Wa, Wb, t * S);
        wP = twiddle2(Wa, Wb, t * S * P);
    }
    double2 a[H], b[H];
    double2 y0 = x0;
#pragma unroll
    for (int j = 1; j <= H; ++j) {
        double2 u = X[swz(base + j * span)], w = X[swz(base + (P - j) * span)];
        if (t != 0) {
            wj = cmul(wj, w1);
            u = cmulc(u, wj);
            w = cmulc(w, cmulc(wP, wj));
        }
        a[j - 1] = cadd(u, w);
        b[j - 1] = csub(u, w);
        y0 = cadd(y0, a[j - 1]);
    }
    X[swz(base)] = y0;
#pragma unroll
    for (int k = 1; k <= H; ++k) {
        double2 re = x0, im = make_double2(0., 0.);
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            const double c = WConst<P>::c((j * k) % P), s = WConst<P>::s((j * k) % P);
            re.x += c * a[j - 1].x;
            re.y += c * a[j - 1].y;
            im.x += s * b[j - 1].x;
            im.y += s * b[j - 1].y;
        }
        const double2 r = rot90<false>(im);
        X[swz(base + k * span)] = cadd(re, r);
        X[swz(base + (P - k) * span)] = csub(re, r);
    }
}

struct ScheduleG {
    int npass;
    int radix[10];
    int nb[10];       // block length of the pass
    int stride[10];   // S = M / nb
    unsigned span_magic[10];  // reciprocal of span = nb / radix
    unsigned per_magic[10];   // reciprocal of M / radix (butterflies per sequence)
};
// radices: odd ones first (9s, 3, 5s: their spans stay multiples of 8), then 16s, then the remaining power of two
SPT_HD ScheduleG make_schedule_g(int M) {
    ScheduleG s;
    s.npass = 0;
    int a = 0, b = 0, c = 0, r = M;
    int p7 = 0, p11 = 0, p13 = 0, p17 = 0, p19 = 0, p23 = 0;   // (only the direct transforms of smooth row lengths have these; convolution lengths do not)
    while (r % 2 == 0) { r /= 2; ++a; }
    while (r % 3 == 0) { r /= 3; ++b; }
    while (r % 5 == 0) { r /= 5; ++c; }
    while (r % 7 == 0) { r /= 7; ++p7; }
    while (r % 11 == 0) { r /= 11; ++p11; }
    while (r % 13 == 0) { r /= 13; ++p13; }
    while (r % 17 == 0) { r /= 17; ++p17; }
    while (r % 19 == 0) { r /= 19; ++p19; }
    while (r % 23 == 0) { r /= 23; ++p23; }
    int Nb = M;
    auto push = [&](int R) {
        s.radix[s.npass] = R;
        s.nb[s.npass] = Nb;
        s.stride[s.npass] = M / Nb;
        s.span_magic[s.npass] = fastdiv_magic(static_cast<unsigned>(Nb / R));
        s.per_magic[s.npass] = fastdiv_magic(static_cast<unsigned>(M / R));
        ++s.npass;
        Nb /= R;
    };
    while (p23 >= 1) { push(23); --p23; }
    while (p19 >= 1) { push(19); --p19; }
    while (p17 >= 1) { push(17); --p17; }
    while (p13 >= 1) { push(13); --p13; }
    while (p11 >= 1) { push(11); --p11; }
    while (p7 >= 1) { push(7); --p7; }
    while (b >= 2) { push(9); b -= 2; }
    if (b == 1) push(3);
    while (c >= 1) { push(5); --c; }
    while (a >= 4) { push(16); a -= 4; }
    if (a == 3) push(8);
    else if (a == 2) push(4);
    else if (a == 1) push(2);
    return s;
}

// Where the forward (DIF) transform leaves frequency k: pass p splits k = k_p + R_p k' and sends residue k_p to the k_p-th
// sub-block of its block, so position = sum_p k_p * (M / (R_0 ... R_p)) -- the mixed-radix digit reversal.  The inverse (DIT)
// transform expects its input in this order.  Used by the direct (non chirp-z) transforms of smooth row lengths.
SPT_HD int dif_output_position(const ScheduleG& s, int M, int k) {
    int pos = 0, span = M;
    for (int p = 0; p < s.npass; ++p) {
        const int R = s.radix[p];
        span /= R;
        pos += (k % R) * span;
        k /= R;
    }
    return pos;
}
// lengths the engine can transform directly: no prime factor above 23
SPT_HD bool is_direct_length(int n) {
    if (n < 1) return false;
    const int primes[9] = {2, 3, 5, 7, 11, 13, 17, 19, 23};
    for (int i = 0; i < 9; ++i)
        while (n % primes[i] == 0) n /= primes[i];
    return n == 1;
}
// slots of one sequence in shared memory: the XOR swizzle permutes within groups of eight
SPT_HD int swz_len(int M) { return (M + 7) & ~7; }

template <int R>
SPT_HD void dif_pass_g(double2* X, int nseq, int M, int Nb, int S, unsigned span_magic, unsigned per_magic,
                       const double2* Wa, const double2* Wb, int tid, int nthr, int ss) {
    const int per = M / R, span = Nb / R;
    for (int w = tid; w < nseq * per; w += nthr) {
        const int sq = nseq == 1 ? 0 : fastdiv(w, per_magic, per), q = w - sq * per;
        if constexpr (R >= 17) dif_butterfly_prime<R>(X + sq * ss, q, Nb, span, span_magic, S, Wa, Wb);
        else dif_butterfly_g<R>(X + sq * ss, q, Nb, span, span_magic, S, Wa, Wb);
    }
}
template <int R, bool CONJ_FILT>
SPT_HD void dit_pass_g(double2* X, int nseq, int M, int Nb, int S, unsigned span_magic, unsigned per_magic,
                       const double2* Wa, const double2* Wb, const double2* filt, int tid, int nthr, int ss) {
    const int per = M / R, span = Nb / R;
    for (int w = tid; w < nseq * per; w += nthr) {
        const int sq = nseq == 1 ? 0 : fastdiv(w, per_magic, per), q = w - sq * per;
        if constexpr (R >= 17) dit_butterfly_prime<R>(X + sq * ss, q, Nb, span, span_magic, S, Wa, Wb);  // (never filtered)
        else dit_butterfly_g<R, CONJ_FILT>(X + sq * ss, q, Nb, span, span_magic, S, Wa, Wb, filt);
    }
}

// seq_stride: distance of consecutive sequences in X (0: M; the direct transforms of lengths that are not multiples of 8
// keep swz_len(M) slots per sequence)
template <bool BIGP = false>   // BIGP: the schedule may hold the radices 17, 19, 23 (direct transforms only)
SPT_HD void fft_dif_g(double2* X, int nseq, int M, const ScheduleG& s, const double2* Wa, const double2* Wb, int tid,
                      int nthr, int seq_stride = 0) {
    const int ss = seq_stride ? seq_stride : M;
    for (int p = 0; p < s.npass; ++p) {
        const int Nb = s.nb[p], S = s.stride[p];
        const unsigned sm_ = s.span_magic[p], pm_ = s.per_magic[p];
        switch (s.radix[p]) {
            case 16: dif_pass_g<16>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 23: if constexpr (BIGP) dif_pass_g<23>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 19: if constexpr (BIGP) dif_pass_g<19>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 17: if constexpr (BIGP) dif_pass_g<17>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 13: dif_pass_g<13>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 11: dif_pass_g<11>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 7: dif_pass_g<7>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 9: dif_pass_g<9>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 8: dif_pass_g<8>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 5: dif_pass_g<5>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 4: dif_pass_g<4>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 3: dif_pass_g<3>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            default: dif_pass_g<2>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
        }
        SPT_SYNC();
    }
}
template <bool CONJ_FILT, bool BIGP = false>
SPT_HD void fft_dit_g(double2* X, int nseq, int M, const ScheduleG& s, const double2* Wa, const double2* Wb,
                      const double2* filt, int tid, int nthr, int seq_stride = 0) {
    const int ss = seq_stride ? seq_stride : M;
    for (int p = s.npass - 1; p >= 0; --p) {
        const int Nb = s.nb[p], S = s.stride[p];
        const unsigned sm_ = s.span_magic[p], pm_ = s.per_magic[p];
        const double2* f = (p == s.npass - 1) ? filt : nullptr;
        switch (s.radix[p]) {
            case 16: dit_pass_g<16, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 23: if constexpr (BIGP) dit_pass_g<23, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 19: if constexpr (BIGP) dit_pass_g<19, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 17: if constexpr (BIGP) dit_pass_g<17, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 13: dit_pass_g<13, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 11: dit_pass_g<11, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 7: dit_pass_g<7, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 9: dit_pass_g<9, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 8: dit_pass_g<8, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 5: dit_pass_g<5, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 4: dit_pass_g<4, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 3: dit_pass_g<3, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            default: dit_pass_g<2, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
        }
        SPT_SYNC();
    }
}


// ---- fused boundary passes -------------------------------------------------------------------------------
// The first forward pass can take its inputs from a generator (the chirp-modulated spectrum / grid row) instead
// of reading X, and the last inverse pass can hand its outputs to a sink (chirp multiply + global store) instead
// of writing X: this removes the zero-fill, the separate load sweep and the separate store sweep of shared memory.
template <int R, class Gen>
SPT_HD void dif_first_pass_g(double2* X, int nseq, int M, unsigned per_magic, const double2* Wa, const double2* Wb,
                             int tid, int nthr, Gen gen) {
    const int per = M / R;  // first pass: one block of length M, span = per, twiddle stride 1
    for (int w = tid; w < nseq * per; w += nthr) {
        const int sq = nseq == 1 ? 0 : fastdiv(w, per_magic, per), t = w - sq * per;
        double2 v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) v[j] = gen(sq, t + j * per);
        dftN<R, true>(v);
        if (t != 0) {
            double2 tw[R];
            twiddle_powers<R>(twiddle2(Wa, Wb, t), tw);
#pragma unroll
            for (int j = 1; j < R; ++j) v[j] = cmul(v[j], tw[j]);
        }
        double2* Xs = X + sq * M;
#pragma unroll
        for (int j = 0; j < R; ++j) Xs[swz(t + j * per)] = v[j];
    }
}
template <int R, bool CONJ_FILT, class Sink>
SPT_HD void dit_last_pass_g(const double2* X, int nseq, int M, unsigned per_magic, const double2* Wa, const double2* Wb,
                            const double2* filt, int tid, int nthr, Sink sink) {
    const int per = M / R;
    for (int w = tid; w < nseq * per; w += nthr) {
        const int sq = nseq == 1 ? 0 : fastdiv(w, per_magic, per), t = w - sq * per;
        const double2* Xs = X + sq * M;
        double2 v[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            v[j] = Xs[swz(t + j * per)];
            if (filt) v[j] = CONJ_FILT ? cmulc(v[j], filt[t + j * per]) : cmul(v[j], filt[t + j * per]);
        }
        if (t != 0) {
            double2 tw[R];
            twiddle_powers<R>(twiddle2(Wa, Wb, t), tw);
#pragma unroll
            for (int j = 1; j < R; ++j) v[j] = cmulc(v[j], tw[j]);
        }
        dftN<R, false>(v);
#pragma unroll
        for (int j = 0; j < R; ++j) sink(sq, t + j * per, v[j]);
    }
}

template <class Gen>
SPT_HD void fft_dif_g_gen(double2* X, int nseq, int M, const ScheduleG& s, const double2* Wa, const double2* Wb, int tid,
                          int nthr, Gen gen) {
    const unsigned pm0 = s.per_magic[0];
    switch (s.radix[0]) {
        case 16: dif_first_pass_g<16>(X, nseq, M, pm0, Wa, Wb, tid, nthr, gen); break;
        case 13: dif_first_pass_g<13>(X, nseq, M, pm0, Wa, Wb, tid, nthr, gen); break;
        case 11: dif_first_pass_g<11>(X, nseq, M, pm0, Wa, Wb, tid, nthr, gen); break;
        case 7: dif_first_pass_g<7>(X, nseq, M, pm0, Wa, Wb, tid, nthr, gen); break;
        case 9: dif_first_pass_g<9>(X, nseq, M, pm0, Wa, Wb, tid, nthr, gen); break;
        case 8: dif_first_pass_g<8>(X, nseq, M, pm0, Wa, Wb, tid, nthr, gen); break;
        case 5: dif_first_pass_g<5>(X, nseq, M, pm0, Wa, Wb, tid, nthr, gen); break;
        case 4: dif_first_pass_g<4>(X, nseq, M, pm0, Wa, Wb, tid, nthr, gen); break;
        case 3: dif_first_pass_g<3>(X, nseq, M, pm0, Wa, Wb, tid, nthr, gen); break;
        default: dif_first_pass_g<2>(X, nseq, M, pm0, Wa, Wb, tid, nthr, gen); break;
    }
    SPT_SYNC();
}
// remaining forward passes (p >= 1)
SPT_HD void fft_dif_g_rest(double2* X, int nseq, int M, const ScheduleG& s, const double2* Wa, const double2* Wb, int tid,
                           int nthr) {
    const int ss = M;
    for (int p = 1; p < s.npass; ++p) {
        const int Nb = s.nb[p], S = s.stride[p];
        const unsigned sm_ = s.span_magic[p], pm_ = s.per_magic[p];
        switch (s.radix[p]) {
            case 16: dif_pass_g<16>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 13: dif_pass_g<13>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 11: dif_pass_g<11>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 7: dif_pass_g<7>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 9: dif_pass_g<9>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 8: dif_pass_g<8>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 5: dif_pass_g<5>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 4: dif_pass_g<4>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            case 3: dif_pass_g<3>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
            default: dif_pass_g<2>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, tid, nthr, ss); break;
        }
        SPT_SYNC();
    }
}
// inverse passes p = npass-1 .. 1 in place, then the last pass (p = 0) into the sink
template <bool CONJ_FILT, class Sink>
SPT_HD void fft_dit_g_sink(double2* X, int nseq, int M, const ScheduleG& s, const double2* Wa, const double2* Wb,
                           const double2* filt, int tid, int nthr, Sink sink) {
    const int ss = M;
    for (int p = s.npass - 1; p >= 1; --p) {
        const int Nb = s.nb[p], S = s.stride[p];
        const unsigned sm_ = s.span_magic[p], pm_ = s.per_magic[p];
        const double2* f = (p == s.npass - 1) ? filt : nullptr;
        switch (s.radix[p]) {
            case 16: dit_pass_g<16, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 13: dit_pass_g<13, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 11: dit_pass_g<11, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 7: dit_pass_g<7, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 9: dit_pass_g<9, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 8: dit_pass_g<8, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 5: dit_pass_g<5, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 4: dit_pass_g<4, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            case 3: dit_pass_g<3, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
            default: dit_pass_g<2, CONJ_FILT>(X, nseq, M, Nb, S, sm_, pm_, Wa, Wb, f, tid, nthr, ss); break;
        }
        SPT_SYNC();
    }
    const double2* f0 = (s.npass == 1) ? filt : nullptr;
    const unsigned pm0 = s.per_magic[0];
    switch (s.radix[0]) {
        case 16: dit_last_pass_g<16, CONJ_FILT>(X, nseq, M, pm0, Wa, Wb, f0, tid, nthr, sink); break;
        case 13: dit_last_pass_g<13, CONJ_FILT>(X, nseq, M, pm0, Wa, Wb, f0, tid, nthr, sink); break;
        case 11: dit_last_pass_g<11, CONJ_FILT>(X, nseq, M, pm0, Wa, Wb, f0, tid, nthr, sink); break;
        case 7: dit_last_pass_g<7, CONJ_FILT>(X, nseq, M, pm0, Wa, Wb, f0, tid, nthr, sink); break;
        case 9: dit_last_pass_g<9, CONJ_FILT>(X, nseq, M, pm0, Wa, Wb, f0, tid, nthr, sink); break;
        case 8: dit_last_pass_g<8, CONJ_FILT>(X, nseq, M, pm0, Wa, Wb, f0, tid, nthr, sink); break;
        case 5: dit_last_pass_g<5, CONJ_FILT>(X, nseq, M, pm0, Wa, Wb, f0, tid, nthr, sink); break;
        case 4: dit_last_pass_g<4, CONJ_FILT>(X, nseq, M, pm0, Wa, Wb, f0, tid, nthr, sink); break;
        case 3: dit_last_pass_g<3, CONJ_FILT>(X, nseq, M, pm0, Wa, Wb, f0, tid, nthr, sink); break;
        default: dit_last_pass_g<2, CONJ_FILT>(X, nseq, M, pm0, Wa, Wb, f0, tid, nthr, sink); break;
    }
}

// smallest 5-smooth multiple of 8 >= need; returns 0 if none <= limit
SPT_HD int conv_length_smooth(int need, int limit) {
    int best = 0;
    for (long long p5 = 1; p5 <= limit; p5 *= 5)
        for (long long p3 = p5; p3 <= limit; p3 *= 3)
            for (long long p2 = p3 * 8; p2 <= limit; p2 *= 2)
                if (p2 >= need && (best == 0 || p2 < best)) best = static_cast<int>(p2);
    return best;
}

// exact phase arithmetic for the chirps: returns r in [0, 2n) with r == (a*a + c*a) mod 2n
SPT_HD long long chirp_residue(long long a, long long c, long long n) {
    long long v = (a * a + c * a) % (2 * n);
    if (v < 0) v += 2 * n;
    return v;
}

// smallest power of two >= n + 2L (at least 8)
SPT_HD int conv_length(int n, int L, int* logM) {
    int M = 8, l = 3;
    while (M < n + 2 * L) {
        M <<= 1;
        ++l;
    }
    *logM = l;
    return M;
}

}  // namespace fftc
}  // namespace sptrans
