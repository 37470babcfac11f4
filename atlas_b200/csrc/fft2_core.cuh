// Register-tiled chirp-z engine ("v2") of the Fourier stage.
//
// The v1 kernels (fft_core.cuh) run ~5 radix passes per FFT through shared memory with a block barrier after each:
// measured 0.56 cycles per point and pass per SM whatever the radix (profiles/ncu_summary_r01.md) -- bound by
// shared-memory round trips and barriers, not by FP64 arithmetic.  v2 restricts the convolution length to
//     M = M1 * 256,   M1 in {8, 9, 10, 12, 15, 16, 18, 20, 24, 25, 27, 30, 32}
// and runs every FFT as ONE block-level radix-M1 pass (a whole butterfly in the registers of one thread, composite
// radices by Good-Thomas / Cooley-Tukey with compile-time twiddles) plus M1 independent 256-point transforms, each
// owned by a half-warp: radix-16 in registers, a 16x16 transpose through the half-warp's own 4 KB of shared memory
// (__syncwarp only), radix-16 again.  The forward transform leaves its spectrum in registers, where it is
// multiplied by the filter (stored in exactly that order, coalesced) and transformed back straight away.
// Per sequence: 4 shared-memory round trips and 3 block barriers instead of ~12 and ~12.
//
// Index algebra (forward, sign -1; the inverse runs the same steps backwards with conjugated twiddles):
//   i = t + 256 j,  k = j' + M1 k'        X^[j' + M1 k'] = sum_t w256^{t k'} [ w_M^{t j'} sum_j w_M1^{j j'} x[t + 256 j] ]
//   t = l + 16 a,   k' = a' + 16 l'       E^[a' + 16 l'] = sum_l w16^{l l'} [ w256^{l a'} sum_a w16^{a a'} e[l + 16 a] ]
//
// Everything is written once for the device and for the CPU thread emulation of tests/cpu/test_fft2_emul.cc
// (one OS thread per CUDA thread, barriers for __syncthreads / __syncwarp), which checks the complete kernels
// bodies against a naive DFT before they ever run on a GPU.
#pragma once

#include "fft_core.cuh"
#include "fft_consts.h"
#include "fourier_types.hpp"

#if defined(__CUDACC__)
#define SPT2_DEV __device__ __forceinline__
#define SPT2_SYNC_BLOCK() __syncthreads()
#define SPT2_SYNC_HALFWARP(tid) __syncwarp(0xFFFFu << ((tid) & 16))
#define SPT2_LDG(p) __ldg(p)
#else
#include <algorithm>
#include <cstring>
using std::min;
#define SPT2_DEV inline
namespace sptrans { namespace emu { void sync_block(); void sync_halfwarp(int tid); } }
#define SPT2_SYNC_BLOCK() ::sptrans::emu::sync_block()
#define SPT2_SYNC_HALFWARP(tid) ::sptrans::emu::sync_halfwarp(tid)
#define SPT2_LDG(p) (*(p))
#endif

namespace sptrans {
namespace fft2 {

using namespace fftc;

constexpr int kM2 = 256;  // length of the half-warp-local transforms

// ---- asynchronous global -> shared copies (raw staging of the next field's inputs) ------------------------------
SPT2_DEV void cp_async16(void* smem_dst, const void* gsrc) {
#if defined(__CUDA_ARCH__)
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
#else
    *static_cast<double2*>(smem_dst) = *static_cast<const double2*>(gsrc);
#endif
}
SPT2_DEV void cp_async_commit_wait_all() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}

// ---- composite register butterflies -----------------------------------------------------------------------------
// a * w_N^k, w_N = e^{-2 pi i/N} (FWD) or its conjugate; k is a compile-time constant after unrolling
template <int N, bool FWD>
SPT_HD double2 mulw(double2 a, int k) {
    k %= N;
    if (k == 0) return a;
    if ((4 * k) % N == 0) {
        const int q = 4 * k / N;
        if (q == 2) return make_double2(-a.x, -a.y);
        return q == 1 ? rot90<FWD>(a) : rot90<!FWD>(a);
    }
    return twc<FWD>(a, WConst<N>::c(k), WConst<N>::s(k));
}

SPT_HD constexpr int modinv(int a, int m) {  // a^{-1} mod m (a, m coprime, small)
    for (int x = 1; x < m; ++x)
        if ((a * x) % m == 1) return x;
    return 1;
}

template <int R, bool FWD>
SPT_HD void dftX(double2* v);

// Cooley-Tukey: j = B a + b, k = ka + A kb, internal twiddles w_R^{b ka}
template <int A, int B, bool FWD>
SPT_HD void dft_ct(double2* v) {
    constexpr int R = A * B;
    double2 y[R];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        double2 t[A];
#pragma unroll
        for (int a = 0; a < A; ++a) t[a] = v[B * a + b];
        dftX<A, FWD>(t);
#pragma unroll
        for (int ka = 0; ka < A; ++ka) y[ka * B + b] = mulw<R, FWD>(t[ka], b * ka);
    }
#pragma unroll
    for (int ka = 0; ka < A; ++ka) {
        double2 t[B];
#pragma unroll
        for (int b = 0; b < B; ++b) t[b] = y[ka * B + b];
        dftX<B, FWD>(t);
#pragma unroll
        for (int kb = 0; kb < B; ++kb) v[ka + A * kb] = t[kb];
    }
}

// Good-Thomas (A, B coprime): j = (B a + A b) mod R, k = (B B^-1 ka + A A^-1 kb) mod R, no internal twiddles
template <int A, int B, bool FWD>
SPT_HD void dft_pfa(double2* v) {
    constexpr int R = A * B;
    constexpr int Binv = modinv(B % A, A), Ainv = modinv(A % B, B);
    double2 y[R];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        double2 t[A];
#pragma unroll
        for (int a = 0; a < A; ++a) t[a] = v[(B * a + A * b) % R];
        dftX<A, FWD>(t);
#pragma unroll
        for (int ka = 0; ka < A; ++ka) y[ka * B + b] = t[ka];
    }
#pragma unroll
    for (int ka = 0; ka < A; ++ka) {
        double2 t[B];
#pragma unroll
        for (int b = 0; b < B; ++b) t[b] = y[ka * B + b];
        dftX<B, FWD>(t);
#pragma unroll
        for (int kb = 0; kb < B; ++kb) v[(B * Binv * ka + A * Ainv * kb) % R] = t[kb];
    }
}

template <int R, bool FWD>
SPT_HD void dftX(double2* v) {
    if constexpr (R == 2 || R == 3 || R == 4 || R == 5 || R == 8 || R == 9 || R == 16) dftN<R, FWD>(v);
    else if constexpr (R == 6) dft_pfa<2, 3, FWD>(v);
    else if constexpr (R == 10) dft_pfa<2, 5, FWD>(v);
    else if constexpr (R == 12) dft_pfa<3, 4, FWD>(v);
    else if constexpr (R == 15) dft_pfa<3, 5, FWD>(v);
    else if constexpr (R == 18) dft_pfa<2, 9, FWD>(v);
    else if constexpr (R == 20) dft_pfa<4, 5, FWD>(v);
    else if constexpr (R == 24) dft_pfa<3, 8, FWD>(v);
    else if constexpr (R == 25) dft_ct<5, 5, FWD>(v);
    else if constexpr (R == 27) dft_ct<3, 9, FWD>(v);
    else if constexpr (R == 30) dft_pfa<5, 6, FWD>(v);
    else if constexpr (R == 32) dft_ct<2, 16, FWD>(v);
    else static_assert(R == 2, "unsupported radix");
}

constexpr bool radix_supported(int R) {
    return R == 8 || R == 9 || R == 10 || R == 12 || R == 15 || R == 16 || R == 18 || R == 20 || R == 24 || R == 25 ||
           R == 27 || R == 30 || R == 32;
}
// smallest supported M = M1 * 256 >= need (0 if none)
SPT_HD int conv_length_v2(int need, int* m1_out) {
    const int radices[13] = {8, 9, 10, 12, 15, 16, 18, 20, 24, 25, 27, 30, 32};
    for (int r = 0; r < 13; ++r)
        if (radices[r] * kM2 >= need) {
            *m1_out = radices[r];
            return radices[r] * kM2;
        }
    *m1_out = 0;
    return 0;
}

// ---- block-level radix-M1 pass ----------------------------------------------------------------------------------
// forward: v[j] = x[t + 256 j] (natural) -> X[256 j' + t] = w_M^{t j'} DFT_M1(v)[j']
// powers of w1 = w_M^t in blocks of four (w^{4g} carried, w^{4g+1..3} = w^{4g} w^{1..3}): dependency chain M1/4 deep
struct TwPow {
    double2 w1, w2, w3, w4, wb;
    SPT_HD explicit TwPow(double2 w) : w1(w), w2(cmul(w, w)), w3(cmul(cmul(w, w), w)), w4(cmul(cmul(w, w), cmul(w, w))), wb(w) {}
    // w^j for the compile-time j of an unrolled loop running j = 1, 2, 3, ... in order
    SPT_HD double2 get(int j) {
        const int r = j & 3;
        if (j < 4) return r == 1 ? w1 : (r == 2 ? w2 : w3);
        if (r == 0) {
            wb = (j == 4) ? w4 : cmul(wb, w4);
            return wb;
        }
        return cmul(wb, r == 1 ? w1 : (r == 2 ? w2 : w3));
    }
};
template <int M1>
SPT2_DEV void passA_fwd_store(double2* v, double2* X, int t, double2 w1) {
    dftX<M1, true>(v);
    X[t] = v[0];
    TwPow tw(w1);
#pragma unroll
    for (int j = 1; j < M1; ++j) X[kM2 * j + t] = cmul(v[j], tw.get(j));
}
// inverse: v[j'] = conj(w_M^{t j'}) X[256 j' + t], then DFT_M1 with sign +1 -> v[j] = x[t + 256 j] * (scale M1)
template <int M1>
SPT2_DEV void passA_inv_load(double2* v, const double2* X, int t, double2 w1) {
    v[0] = X[t];
    TwPow tw(w1);
#pragma unroll
    for (int j = 1; j < M1; ++j) v[j] = cmulc(X[kM2 * j + t], tw.get(j));
    dftX<M1, false>(v);
}

// ---- half-warp-local 256-point transforms -------------------------------------------------------------------------
// `base`: the 256 contiguous elements of sub-transform s; l = lane within the half-warp.  The 16x16 transpose uses a
// rotation (row a', slot (l + a') & 15) so that both the row-wise and the column-wise accesses are conflict free.
// T256 = e^{-2 pi i q l/256} at [q*16 + l] (shared memory in the transform kernels).
// MODE 0: forward, multiply by filt (or its conjugate), inverse, back in place (natural order).  The 16 filter values
//         of this lane are requested first so that their L2 latency hides behind the forward transform.
// MODE 1: forward only, registers * scale written to out[q*16 + l]  (filter-table construction)
template <int MODE, bool CONJ_FILT>
SPT2_DEV void local256(double2* base, int l, int tid, const double2* __restrict__ T256,
                       const double2* __restrict__ filt, double2* __restrict__ out, double scale) {
    double2 e[16];
    double2 fl[MODE == 0 ? 16 : 1];
    if constexpr (MODE == 0) {
#pragma unroll
        for (int q = 0; q < 16; ++q) fl[q] = SPT2_LDG(filt + q * 16 + l);
    }
#pragma unroll
    for (int a = 0; a < 16; ++a) e[a] = base[l + 16 * a];
    dft16<true>(e);
#pragma unroll
    for (int q = 1; q < 16; ++q) e[q] = cmul(e[q], T256[q * 16 + l]);
    SPT2_SYNC_HALFWARP(tid);  // every lane has read its column before rows are overwritten
#pragma unroll
    for (int q = 0; q < 16; ++q) base[16 * q + ((l + q) & 15)] = e[q];
    SPT2_SYNC_HALFWARP(tid);
#pragma unroll
    for (int q = 0; q < 16; ++q) e[q] = base[16 * l + ((q + l) & 15)];
    dft16<true>(e);
    if constexpr (MODE == 1) {
#pragma unroll
        for (int q = 0; q < 16; ++q) out[q * 16 + l] = make_double2(e[q].x * scale, e[q].y * scale);
    }
    else {
#pragma unroll
        for (int q = 0; q < 16; ++q) e[q] = CONJ_FILT ? cmulc(e[q], fl[q]) : cmul(e[q], fl[q]);
        dft16<false>(e);
#pragma unroll
        for (int q = 1; q < 16; ++q) e[q] = cmulc(e[q], T256[q * 16 + l]);
#pragma unroll
        for (int q = 0; q < 16; ++q) base[16 * l + ((q + l) & 15)] = e[q];  // own row
        SPT2_SYNC_HALFWARP(tid);
#pragma unroll
        for (int q = 0; q < 16; ++q) e[q] = base[16 * q + ((l + q) & 15)];
        dft16<false>(e);
        SPT2_SYNC_HALFWARP(tid);  // all rows read before columns are written
#pragma unroll
        for (int a = 0; a < 16; ++a) base[l + 16 * a] = e[a];
    }
}

template <int M1, int NT, bool CONJ_FILT>
SPT2_DEV void local_phase(double2* X, int tid, const double2* __restrict__ T256, const double2* __restrict__ filt) {
    const int l = tid & 15;
    for (int s = tid >> 4; s < M1; s += NT / 16)
        local256<0, CONJ_FILT>(X + kM2 * s, l, tid, T256, filt + kM2 * s, nullptr, 1.0);
}

// ---- filter table of one (n, L) class ---------------------------------------------------------------------------------
// b_k = e^{-i pi k^2/n} for k in [-2L, n-1] (cyclic, zero elsewhere), transformed by the forward machinery and written
// in register order with the 1/M of the unnormalised inverse folded in.  One block of NT threads.
template <int M1, int NT>
SPT2_DEV void filter_table_body(const PairMeta& pm, int tid, double2* X, const double2* __restrict__ W1,
                                const double2* __restrict__ T256, double2* __restrict__ filt_out) {
    const int M = M1 * kM2, L = pm.L;
    const int n = pm.n;
    for (int t = tid; t < kM2; t += NT) {
        double2 v[M1];
#pragma unroll
        for (int j = 0; j < M1; ++j) {
            const int idx = t + kM2 * j;  // cyclic index: k = idx for idx < n, k = idx - M for idx >= M - 2L
            int k = idx;
            bool on = idx < n;
            if (idx >= M - 2 * L) {
                k = idx - M;
                on = true;
            }
            double2 val = make_double2(0., 0.);
            if (on) {
                double s, c;
#if defined(__CUDA_ARCH__)
                sincospi(-static_cast<double>(chirp_residue(k, 0, n)) / n, &s, &c);
#else
                const double ang = -3.14159265358979323846 * static_cast<double>(chirp_residue(k, 0, n)) / n;
                s = __builtin_sin(ang);
                c = __builtin_cos(ang);
#endif
                val = make_double2(c, s);
            }
            v[j] = val;
        }
        passA_fwd_store<M1>(v, X, t, SPT2_LDG(W1 + t));
    }
    SPT2_SYNC_BLOCK();
    const int l = tid & 15;
    for (int s = tid >> 4; s < M1; s += NT / 16)
        local256<1, false>(X + kM2 * s, l, tid, T256, nullptr, filt_out + kM2 * s, 1.0 / M);
}

// ---- transform kernels ------------------------------------------------------------------------------------------------
// Shared memory (double2 units): X[M] | T256[256] | S (staging area of the field being fetched).
// Loads from the L2-resident class tables (chirps, filter) are requested in batches ahead of the arithmetic that hides
// them: with 8 warps per SM all at the same phase, a load issued next to its use stalls the whole SM for an L2 round trip
// (first ncu capture: 43 % of the stall samples were long-scoreboard waits on exactly these loads).
template <int M1>
struct V2Counts {
    static constexpr int NA = (M1 + 1) / 2;        // 256 j <= 2L < M/2: inputs of the inverse / outputs of the direct kernel
    static constexpr int NJ = (7 * M1 + 9) / 10;   // chirp values C requested ahead of the butterfly (rows up to 0.7 M)
};

template <int NT>
SPT2_DEV void load_t256(double2* Tsm, const double2* __restrict__ t256, int tid) {
    for (int e = tid; e < 256; e += NT) Tsm[e] = SPT2_LDG(t256 + e);
}

// Row table of one latitude pair (shared memory, built once per block): rowtab[2 m + par] = row of (m, par, pair) in
// the exchange buffer.  The per-field staging / store loops then need no dependent global loads.
template <int NT>
SPT2_DEV void build_rowtab(const Fft2Args& a, int pair, int L, int tid, int* rowtab) {
    for (int e = tid; e < 2 * (L + 1); e += NT) {
        const int m = e >> 1, par = e & 1;
        const int n0 = SPT2_LDG(a.nlat0 + m);
        rowtab[e] = static_cast<int>(SPT2_LDG(a.fb_rowoff + m) + static_cast<long long>(par) * (a.nleg - n0) + (pair - n0));
    }
}
// inverse: raw sym / asym parts of field f, S[2 m + par], m <= Lc
template <int NT>
SPT2_DEV void stage_inv_inputs(const Fft2Args& a, int f, int Lc, int tid, double2* S, const int* rowtab) {
    const int cnt = 2 * (Lc + 1);
    for (int e = tid; e < cnt; e += NT) cp_async16(S + e, a.fb + static_cast<long long>(rowtab[e]) * a.nf + f);
}
// the same for the first field of a block, before the row table exists: four independent address chains per thread
template <int NT>
SPT2_DEV void stage_inv_inputs_first(const Fft2Args& a, int pair, int f, int Lc, int tid, double2* S) {
    const int cnt = 2 * (Lc + 1);
    for (int e0 = tid; e0 < cnt; e0 += 4 * NT) {
        const double2* src[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = min(e0 + u * NT, cnt - 1);
            const int m = e >> 1, par = e & 1;
            const int n0 = SPT2_LDG(a.nlat0 + m);
            const long long row = SPT2_LDG(a.fb_rowoff + m) + static_cast<long long>(par) * (a.nleg - n0) + (pair - n0);
            src[u] = a.fb + row * a.nf + f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (e0 + u * NT < cnt) cp_async16(S + e0 + u * NT, src[u]);
    }
}

template <int M1, int NT>
SPT2_DEV void fourier2_inv_body(const Fft2Args& a, int pair, int f0, int nfb, int tid, double2* X) {
    using CN = V2Counts<M1>;
    const PairMeta pm = a.meta[pair];
    const int n = pm.n, L = pm.L;
    const int Lc = min(L, a.mlimit);
    double2* Tsm = X + M1 * kM2;
    double2* S = Tsm + 256;
    int* rowtab = reinterpret_cast<int*>(S + 2 * (L + 1));
    const double2* __restrict__ A = a.chirp + pm.chirp_off;
    const double2* __restrict__ C = A + (2 * L + 1);
    const double2* __restrict__ W1 = a.twid + pm.tw_off;
    const double2* __restrict__ F2 = a.filt + pm.filt_off;
    stage_inv_inputs_first<NT>(a, pair, f0, Lc, tid, S);
    if (nfb > 1) build_rowtab<NT>(a, pair, L, tid, rowtab);  // first read after the block barriers of field f0
    load_t256<NT>(Tsm, a.t256, tid);
    for (int fi = 0; fi < nfb; ++fi) {
        const int f = f0 + fi;
        cp_async_commit_wait_all();
        SPT2_SYNC_BLOCK();  // S holds field f; X is free (previous field's outputs are in registers / stored)
        for (int t = tid; t < kM2; t += NT) {
            double2 ach[CN::NA];
#pragma unroll
            for (int j = 0; j < CN::NA; ++j)
                ach[j] = (kM2 * j <= 2 * L) ? SPT2_LDG(A + min(t + kM2 * j, 2 * L)) : make_double2(0., 0.);
            const double2 w1 = SPT2_LDG(W1 + t);
            double2 v[M1];
#pragma unroll
            for (int j = 0; j < M1; ++j) {
                double2 val = make_double2(0., 0.);
                if (j < CN::NA && kM2 * j <= 2 * L) {  // block-uniform: elements beyond 2L are zero padding
                    const int i = t + kM2 * j;
                    const int m = i - L, am = m < 0 ? -m : m;
                    const int ar = min(am, Lc);
                    double2 cs = S[2 * ar], ca = S[2 * ar + 1];
                    if (am == 0) cs.y = ca.y = 0.;  // only Re of m = 0 enters (reference :1165)
                    double2 FN, FS;
                    if (pm.has_s) {
                        FN = cadd(cs, ca);
                        FS = csub(cs, ca);
                    }
                    else {  // equator row: the reference's southern loop overwrites it with sym - asym (:1061-1070)
                        FN = csub(cs, ca);
                        FS = make_double2(0., 0.);
                    }
                    // Z_m = F_N + i F_S ;  Z_{-m} = conj(F_N) + i conj(F_S)
                    const double2 Z = m >= 0 ? make_double2(FN.x - FS.y, FN.y + FS.x)
                                             : make_double2(FN.x + FS.y, FS.x - FN.y);
                    const double2 za = cmul(Z, ach[j < CN::NA ? j : 0]);
                    if (am <= Lc) val = za;
                }
                v[j] = val;
            }
            passA_fwd_store<M1>(v, X, t, w1);
        }
        SPT2_SYNC_BLOCK();  // S consumed, X complete
        if (fi + 1 < nfb) stage_inv_inputs<NT>(a, f + 1, Lc, tid, S, rowtab);  // lands behind the transforms
        local_phase<M1, NT, false>(X, tid, Tsm, F2);
        SPT2_SYNC_BLOCK();
        const double sc = (f < a.nb_uv) ? a.scale_lat[pair] : 1.0;  // u,v = U,V / cos(lat)  (reference :1443-1469)
        double* __restrict__ gN = a.gp + f * a.npts + pm.rowN;
        double* __restrict__ gS = a.gp + f * a.npts + pm.rowS;
        for (int t = tid; t < kM2; t += NT) {
            double2 c[CN::NJ];
#pragma unroll
            for (int j = 0; j < CN::NJ; ++j)
                c[j] = (kM2 * j < n) ? SPT2_LDG(C + min(t + kM2 * j, n - 1)) : make_double2(0., 0.);
            const double2 w1 = SPT2_LDG(W1 + t);
            double2 v[M1];
            passA_inv_load<M1>(v, X, t, w1);
#pragma unroll
            for (int j = 0; j < M1; ++j) {
                const int i = t + kM2 * j;
                if (i < n) {
                    const double2 cc = j < CN::NJ ? c[j < CN::NJ ? j : 0] : SPT2_LDG(C + i);
                    const double2 z = cmul(v[j], cc);
                    gN[i] = z.x * sc;
                    if (pm.has_s) gS[i] = z.y * sc;
                }
            }
        }
    }
}

// ---- direct Fourier stage (grid -> spectral) ------------------------------------------------------------------------------
// Staging area: northern row (n doubles) and southern row (n doubles) of the field being fetched.
template <int NT>
SPT2_DEV void stage_dir_inputs(const Fft2Args& a, const PairMeta& pm, int f, int tid, double* S) {
    const int n = pm.n, q = n / 2;  // v2 classes have even n: rows are 16-byte multiples
    const double* gN = a.gp + f * a.npts + pm.rowN;
    const double* gS = a.gp + f * a.npts + pm.rowS;
    if (a.gp_aligned16) {
        for (int e = tid; e < q; e += NT) cp_async16(S + 2 * e, gN + 2 * e);
        if (pm.has_s)
            for (int e = tid; e < q; e += NT) cp_async16(S + n + 2 * e, gS + 2 * e);
    }
    else {  // grid buffer at an odd multiple of 8 bytes: plain loads
        for (int e = tid; e < n; e += NT) S[e] = gN[e];
        if (pm.has_s)
            for (int e = tid; e < n; e += NT) S[n + e] = gS[e];
    }
}

template <int M1, int NT>
SPT2_DEV void fourier2_dir_body(const Fft2Args& a, int pair, int f0, int nfb, int tid, double2* X) {
    using CN = V2Counts<M1>;
    const PairMeta pm = a.meta[pair];
    const int n = pm.n, L = pm.L;
    double2* Tsm = X + M1 * kM2;
    double* S = reinterpret_cast<double*>(Tsm + 256);
    int* rowtab = reinterpret_cast<int*>(S + 2 * n);
    const double2* __restrict__ A = a.chirp + pm.chirp_off;
    const double2* __restrict__ C = A + (2 * L + 1);
    const double2* __restrict__ W1 = a.twid + pm.tw_off;
    const double2* __restrict__ F2 = a.filt + pm.filt_off;
    // direct transform: quadrature weight, 1/n normalisation.  Adjoint of the inverse (invtrans_adj): no weight and no
    // normalisation; w.r.t. the ectrans / TransIFS spectral inner product (m > 0 counted twice) every m gets the same factor
    const double wq0 = a.adjoint ? 1.0 : a.weights[pair];
    const double inv_n = a.adjoint ? 1.0 : 1.0 / n;
    stage_dir_inputs<NT>(a, pm, f0, tid, S);
    load_t256<NT>(Tsm, a.t256, tid);
    build_rowtab<NT>(a, pair, L, tid, rowtab);  // first read after several block barriers
    for (int fi = 0; fi < nfb; ++fi) {
        const int f = f0 + fi;
        cp_async_commit_wait_all();
        SPT2_SYNC_BLOCK();  // S holds field f; the previous field's spectrum has been read out of X
        const double sc = (f < a.nb_uv) ? a.scale_lat[pair] : 1.0;  // wind components enter as u,v * scale(lat)
        for (int t = tid; t < kM2; t += NT) {
            double2 c[CN::NJ];
#pragma unroll
            for (int j = 0; j < CN::NJ; ++j)
                c[j] = (kM2 * j < n) ? SPT2_LDG(C + min(t + kM2 * j, n - 1)) : make_double2(0., 0.);
            const double2 w1 = SPT2_LDG(W1 + t);
            double2 v[M1];
#pragma unroll
            for (int j = 0; j < M1; ++j) {
                double2 val = make_double2(0., 0.);
                if (kM2 * j < n) {  // block-uniform
                    const int i = t + kM2 * j;
                    if (i < n) {
                        const double xn = S[i] * sc;
                        const double xs = pm.has_s ? S[n + i] * sc : 0.;
                        const double2 cc = j < CN::NJ ? c[j < CN::NJ ? j : 0] : SPT2_LDG(C + i);
                        val = cmulc(make_double2(xn, xs), cc);
                    }
                }
                v[j] = val;
            }
            passA_fwd_store<M1>(v, X, t, w1);
        }
        SPT2_SYNC_BLOCK();
        if (fi + 1 < nfb) stage_dir_inputs<NT>(a, pm, f + 1, tid, S);
        local_phase<M1, NT, true>(X, tid, Tsm, F2);
        SPT2_SYNC_BLOCK();
        for (int t = tid; t < kM2; t += NT) {
            const double2 w1 = SPT2_LDG(W1 + t);
            double2 v[M1];
            passA_inv_load<M1>(v, X, t, w1);
#pragma unroll
            for (int j = 0; j < CN::NA; ++j)
                if (kM2 * j <= 2 * L) X[kM2 * j + t] = v[j];  // same addresses this thread has just read: in place
        }
        SPT2_SYNC_BLOCK();
        for (int m0 = tid; m0 <= L; m0 += 4 * NT) {  // four zonal wavenumbers per thread and sweep, chirp loads batched
            double2 ap[4], am_[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = min(m0 + u * NT, L);
                ap[u] = SPT2_LDG(A + L + m);
                am_[u] = SPT2_LDG(A + L - m);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = m0 + u * NT;
                if (m <= L) {
                    const double wq = wq0;
                    double2 Gp = cmulc(X[L + m], ap[u]);
                    double2 Gm = cmulc(X[L - m], am_[u]);
                    Gp.x *= inv_n; Gp.y *= inv_n; Gm.x *= inv_n; Gm.y *= inv_n;
                    // F_N = (G_m + conj(G_-m))/2 ; F_S = (G_m - conj(G_-m))/(2i)
                    const double2 FN = make_double2(0.5 * (Gp.x + Gm.x), 0.5 * (Gp.y - Gm.y));
                    const double2 FS = make_double2(0.5 * (Gp.y + Gm.y), -0.5 * (Gp.x - Gm.x));
                    double2 s, as;
                    if (pm.has_s) {
                        s = make_double2((FN.x + FS.x) * wq, (FN.y + FS.y) * wq);
                        as = make_double2((FN.x - FS.x) * wq, (FN.y - FS.y) * wq);
                    }
                    else {
                        s = make_double2(FN.x * wq, FN.y * wq);
                        as = s;
                    }
                    a.fb[static_cast<long long>(rowtab[2 * m]) * a.nf + f] = s;
                    a.fb[static_cast<long long>(rowtab[2 * m + 1]) * a.nf + f] = as;
                }
            }
        }
    }
}

}  // namespace fft2
}  // namespace sptrans
