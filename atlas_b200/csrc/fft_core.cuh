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
