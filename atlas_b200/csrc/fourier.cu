// Fourier stage of the spectral transform on the GPU: batched shared-memory chirp-z FFTs over the ragged
// rows of a reduced Gaussian grid, one complex transform per (latitude pair, field).
//
// Replaces TransLocal::invtrans_fourier_reduced / _regular (ecmwf/atlas src/atlas/trans/local/TransLocal.cc:
// 1101-1196: per field and latitude pack nx/2+1 complex values, one FFTW c2r behind a mutex, copy out), the
// hemisphere merge of invtrans_legendre (:1034-1079, fused into the load: north = sym + asym, south =
// sym - asym) and the u,v = U,V / cos(lat) pass of invtrans_uv (:1443-1469, fused into the store).
// The direct kernel is the mirror image (r2c, split into sym/asym parts, quadrature weight applied).
//
// Layout of the Legendre<->Fourier exchange buffer (double2 = (re,im)):
//     fb[(fb_rowoff[m] + par*ncol(m) + (j - nlat0[m])) * nf + field],   ncol(m) = nleg - nlat0[m]
// Algorithm and index algebra: fft_core.cuh (unit-tested on the CPU).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include "fft2_core.cuh"
#include "fft_core.cuh"
#include "fourier_types.hpp"
#include "plan.hpp"

namespace sptrans {

namespace {

using namespace fftc;

constexpr int kMaxM = 13824;     // largest convolution length a single CTA can hold in shared memory (221 KB)
constexpr int kFftThreads = 256;

// the pass schedule (radices, reciprocals for index arithmetic) is computed once per class on the host: deriving it
// in every thread cost ~2 us per block in 64-bit divisions
__device__ __forceinline__ void load_schedule(const ScheduleG* __restrict__ g, ScheduleG* s, int tid) {
    const int nw = sizeof(ScheduleG) / 4;
    if (tid < nw) reinterpret_cast<int*>(s)[tid] = reinterpret_cast<const int*>(g)[tid];
}

__device__ __forceinline__ void load_twiddles(const double2* __restrict__ g, double2* s, int M, int tid, int nthr) {
    const int ntw = M / 64 + 1 + 64;
    for (int e = tid; e < ntw; e += nthr) s[e] = g[e];
}

// one block per distinct (n, L): chirps, two-level twiddles and the digit-reversed filter spectrum
__global__ void __launch_bounds__(kFftThreads)
chirp_tables_kernel(const PairMeta* __restrict__ cls, double2* __restrict__ chirp, double2* __restrict__ filt,
                    double2* __restrict__ twid) {
    extern __shared__ double2 X[];
    const PairMeta pm = cls[blockIdx.x];
    const int L = pm.L, M = pm.M;
    const int n = pm.mode ? pm.n / 2 : pm.n;  // length of the complex chirp-z problem
    const int tid = threadIdx.x, nthr = blockDim.x;
    double2* A = chirp + pm.chirp_off;
    double2* C = A + (2 * L + 1);
    if (pm.mode) {  // e^{2 pi i m / n_row}: recombines the even/odd sample transforms of a real row
        double2* Wr = C + n;
        for (int m = tid; m <= L; m += nthr) {
            double s, c;
            sincospi(2.0 * m / pm.n, &s, &c);
            Wr[m] = make_double2(c, s);
        }
    }
    double2* Wa = twid + pm.tw_off;
    double2* Wb = Wa + (M / 64 + 1);
    for (int u = tid; u <= 2 * L; u += nthr) {
        double s, c;
        sincospi(static_cast<double>(chirp_residue(u, 0, n)) / n, &s, &c);
        A[u] = make_double2(c, s);
    }
    for (int i = tid; i < n; i += nthr) {
        double s, c;
        sincospi(static_cast<double>(chirp_residue(i, -2LL * L, n)) / n, &s, &c);
        C[i] = make_double2(c, s);
    }
    for (int k = tid; k <= M / 64; k += nthr) {
        double s, c;
        sincospi(-2.0 * (64.0 * k) / M, &s, &c);
        Wa[k] = make_double2(c, s);
    }
    for (int k = tid; k < 64; k += nthr) {
        double s, c;
        sincospi(-2.0 * k / M, &s, &c);
        Wb[k] = make_double2(c, s);
    }
    const int PL = M;
    double2* sW = X + PL;
    for (int e = tid; e < PL; e += nthr) X[e] = make_double2(0., 0.);
    __syncthreads();
    load_twiddles(Wa, sW, M, tid, nthr);
    for (int e = tid; e < n + 2 * L; e += nthr) {
        const int k = e - 2 * L;  // k in [-2L, n-1]
        double s, c;
        sincospi(-static_cast<double>(chirp_residue(k, 0, n)) / n, &s, &c);
        const int idx = (k % M + M) % M;
        X[swz(idx)] = make_double2(c, s);
    }
    __syncthreads();
    const ScheduleG sc = make_schedule_g(M);
    fft_dif_g(X, 1, M, sc, sW, sW + (M / 64 + 1), tid, nthr);
    const double scl = 1.0 / M;
    for (int k = tid; k < M; k += nthr) {
        const double2 v = X[swz(k)];
        filt[pm.filt_off + k] = make_double2(v.x * scl, v.y * scl);
    }
}

__global__ void __launch_bounds__(2 * kFftThreads, 1)
fourier_inv_kernel(const PairMeta* __restrict__ meta, const int2* __restrict__ blocks, int nf,
                   int mlimit, int nb_uv, const double2* __restrict__ fb, const long long* __restrict__ fb_rowoff,
                   const int* __restrict__ nlat0, int nleg, const ScheduleG* __restrict__ scheds, const double2* __restrict__ twid,
                   const double2* __restrict__ chirp, const double2* __restrict__ filt,
                   const double* __restrict__ coslatinv, double* __restrict__ gp, long long npts) {
    extern __shared__ double2 X[];
    const int2 bd = blocks[blockIdx.x];
    const int pair = bd.x, f0 = bd.y & 0xffff;
    const PairMeta pm = meta[pair];
    const int nfb = bd.y >> 16;  // fields of this block (block descriptor: f0 | count << 16)
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n = pm.n, L = pm.L;
    const int M = pm.M, PL = M;
    const int Lc = min(L, mlimit);
    if (Lc < 0) {  // no zonal wavenumber resolved / requested at this latitude: rows are zero
        for (int w = tid; w < nfb * n; w += nthr) {
            const int fi = w / n, i = w - fi * n;
            gp[(f0 + fi) * npts + pm.rowN + i] = 0.;
            if (pm.has_s) gp[(f0 + fi) * npts + pm.rowS + i] = 0.;
        }
        return;
    }
    double2* sW = X + pm.F * PL;
    __shared__ ScheduleG sc;
    load_schedule(scheds + pm.sched, &sc, tid);
    for (int e = tid; e < nfb * PL; e += nthr) X[e] = make_double2(0., 0.);
    load_twiddles(twid + pm.tw_off, sW, M, tid, nthr);
    __syncthreads();
    const double2* A = chirp + pm.chirp_off;
    const double2* C = A + (2 * L + 1);
    for (int w = tid; w < nfb * (Lc + 1); w += nthr) {
        const int m = w / nfb, fi = w - m * nfb;
        const int n0 = nlat0[m];
        const long long is = (fb_rowoff[m] + (pair - n0)) * nf + f0 + fi;
        const long long ia = is + static_cast<long long>(nleg - n0) * nf;
        double2 cs = fb[is], ca = fb[ia];
        if (m == 0) cs.y = ca.y = 0.;  // only Re of m = 0 enters (reference :1165)
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
        const double2 Zp = make_double2(FN.x - FS.y, FN.y + FS.x);
        X[fi * PL + swz(L + m)] = cmul(Zp, A[L + m]);
        if (m > 0) {
            const double2 Zm = make_double2(FN.x + FS.y, FS.x - FN.y);
            X[fi * PL + swz(L - m)] = cmul(Zm, A[L - m]);
        }
    }
    __syncthreads();
    fft_dif_g(X, nfb, M, sc, sW, sW + (M / 64 + 1), tid, nthr);
    fft_dit_g<false>(X, nfb, M, sc, sW, sW + (M / 64 + 1), filt + pm.filt_off, tid, nthr);
    for (int w = tid; w < nfb * n; w += nthr) {
        const int fi = w / n, i = w - fi * n;
        const double2 z = cmul(X[fi * PL + swz(i)], C[i]);
        const int f = f0 + fi;
        double sn = 1., ss = 1.;
        if (f < nb_uv) {  // u,v = U,V / cos(lat)  (reference :1443-1469)
            sn = ss = coslatinv[pair];  // grid is symmetric about the equator
        }
        gp[f * npts + pm.rowN + i] = z.x * sn;
        if (pm.has_s) gp[f * npts + pm.rowS + i] = z.y * ss;
    }
}

// Rows r0 <= r < r1 (r = 2 m + parity, m <= L) of one latitude pair: local exchange buffer -> buffer of the rank that owns m.
// One warp per row, up to 5 x 16 B per lane in flight; the rows were written by other blocks, so they are read
// through L2 (ld.cg).
__device__ __forceinline__ void push_pair_rows(int pair, int r0, int r1, int nf, const double2* __restrict__ fb,
                                               const long long* __restrict__ fb_rowoff, const int* __restrict__ nlat0,
                                               int nleg, const int* __restrict__ owner, const PeerDst& dst, int me) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    constexpr int kU = 5;
    for (int r = r0 + warp; r < r1; r += nw) {
        const int m = r >> 1, par = r & 1;
        const int o = owner[m];
        if (o == me) continue;
        const int n0 = nlat0[m];
        const long long row = fb_rowoff[m] + static_cast<long long>(par) * (nleg - n0) + (pair - n0);
        const double2* src = fb + row * nf;
        double2* out = reinterpret_cast<double2*>(dst.base[o]) + row * nf;
        for (int i0 = lane; i0 < nf; i0 += 32 * kU) {
            double2 v[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (i0 + 32 * u < nf) v[u] = __ldcg(src + i0 + 32 * u);
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (i0 + 32 * u < nf) out[i0 + 32 * u] = v[u];
        }
    }
}

// Sharded plans: every row of the exchange buffer belongs to the rank that owns its zonal wavenumber, and the rows of a
// latitude pair are complete once all the field-group blocks of the pair are through.  The block that completes a pair puts it
// on a work queue, and EVERY block that finishes -- not just that one -- takes chunks of d_push_rows rows off the queue and ships
// them (nf double2 = whole rows: long coalesced NVLink stores) while the other SMs keep transforming.  With the completing
// block shipping all 2 (L + 1) rows alone (5.6 MB at the equator), the last pairs of a launch left one block per pair
// copying for ~0.4 ms after everybody else had finished.
// Queue (ints behind the nleg pair counters): [0] tail, [1] head, [8 .. 8 + nleg) pair + 1 once published,
// [8 + nleg .. 8 + 2 nleg) next chunk of the slot.  Zeroed by the host before the stage.
__device__ int d_push_rows = 128;  // rows per chunk (8-rank emulation on one B200: 16: 2.18, 32: 1.90, 64: 1.80, 128: 1.73, 256: 1.74 ms per rank;
                                   // SPTRANS_PUSH_ROWS=0 (measurement): the completing block ships the whole pair alone, 2.21 ms)
__device__ __forceinline__ void finish_pair_and_push(int pair, int nblk, int* __restrict__ pair_done, int nleg,
                                                     const PairMeta* __restrict__ meta, int nf, const double2* __restrict__ fb,
                                                     const long long* __restrict__ fb_rowoff, const int* __restrict__ nlat0,
                                                     const int* __restrict__ owner, const PeerDst& dst, int me) {
    __shared__ int s_pair, s_chunk, s_publisher;
    int* q = pair_done + nleg;
    const int tid = threadIdx.x;
    const int cfg_rows = d_push_rows;
    const int kPushRows = cfg_rows > 0 ? cfg_rows : (1 << 20);
    __threadfence();   // this block's rows are visible before its arrival is
    __syncthreads();
    if (tid == 0) {
        const int done = atomicAdd(pair_done + pair, 1);
        s_publisher = (done == nblk - 1);
        if (done == nblk - 1) {
            pair_done[pair] = 0;  // ready for the next call
            __threadfence();
            const int slot = atomicAdd(q, 1);
            atomicExch(q + 8 + slot, pair + 1);
        }
    }
    bool pushed = false;
    for (;;) {
        __syncthreads();
        if (cfg_rows <= 0 && !s_publisher) break;
        if (tid == 0) {
            int found = -1, chunk = 0;
            int s = atomicAdd(q + 1, 0);
            const int tail = atomicAdd(q, 0);
            while (s < tail) {
                int pp = atomicAdd(q + 8 + s, 0);
                while (pp == 0) pp = atomicAdd(q + 8 + s, 0);  // reserved, published within a few instructions by a running block
                const int nch = (2 * (meta[pp - 1].L + 1) + kPushRows - 1) / kPushRows;
                const int c = atomicAdd(q + 8 + nleg + s, 1);
                if (c < nch) {
                    found = pp - 1;
                    chunk = c;
                    break;
                }
                atomicMax(q + 1, s + 1);  // slot exhausted
                ++s;
            }
            s_pair = found;
            s_chunk = chunk;
        }
        __syncthreads();
        const int pp = s_pair;
        if (pp < 0) break;
        __threadfence();  // (pairs with the fences of the blocks that wrote the rows)
        const int r0 = s_chunk * kPushRows;
        push_pair_rows(pp, r0, min(r0 + kPushRows, 2 * (meta[pp].L + 1)), nf, fb, fb_rowoff, nlat0, nleg, owner, dst, me);
        pushed = true;
    }
    if (pushed) __threadfence_system();  // remote rows are visible to the peers before this kernel completes
}

__global__ void __launch_bounds__(2 * kFftThreads, 1)
fourier_dir_kernel(const PairMeta* __restrict__ meta, const int2* __restrict__ blocks, int nf,
                   int nb_uv, const double* __restrict__ gp, long long npts, const long long* __restrict__ fb_rowoff,
                   const int* __restrict__ nlat0, int nleg, const ScheduleG* __restrict__ scheds, const double2* __restrict__ twid,
                   const double2* __restrict__ chirp, const double2* __restrict__ filt,
                   const double* __restrict__ weights, const double* __restrict__ coslat, double2* __restrict__ fb,
                   int adjoint, const int* __restrict__ owner, const __grid_constant__ PeerDst dst, int me,
                   int* __restrict__ pair_done) {
    extern __shared__ double2 X[];
    const int2 bd = blocks[blockIdx.x];
    const int pair = bd.x, f0 = bd.y & 0xffff;
    const PairMeta pm = meta[pair];
    const int nfb = bd.y >> 16;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n = pm.n, L = pm.L;
    if (L < 0) return;
    const int M = pm.M, PL = M;
    double2* sW = X + pm.F * PL;
    __shared__ ScheduleG sc;
    load_schedule(scheds + pm.sched, &sc, tid);
    for (int e = tid; e < nfb * PL; e += nthr) X[e] = make_double2(0., 0.);
    load_twiddles(twid + pm.tw_off, sW, M, tid, nthr);
    __syncthreads();
    const double2* A = chirp + pm.chirp_off;
    const double2* C = A + (2 * L + 1);
    for (int w = tid; w < nfb * n; w += nthr) {
        const int fi = w / n, i = w - fi * n;
        const int f = f0 + fi;
        double xn = gp[f * npts + pm.rowN + i];
        double xs = pm.has_s ? gp[f * npts + pm.rowS + i] : 0.;
        if (f < nb_uv) {  // wind components enter the vor/div transform as u,v / (a cos(lat))
            xn *= coslat[pair];
            xs *= coslat[pair];
        }
        X[fi * PL + swz(i)] = cmulc(make_double2(xn, xs), C[i]);
    }
    __syncthreads();
    fft_dif_g(X, nfb, M, sc, sW, sW + (M / 64 + 1), tid, nthr);
    fft_dit_g<true>(X, nfb, M, sc, sW, sW + (M / 64 + 1), filt + pm.filt_off, tid, nthr);
    // direct transform: quadrature weight, 1/n normalisation.  Adjoint of the inverse (invtrans_adj): no weight and no
    // normalisation.  The adjoint is taken w.r.t. the spectral inner product of ectrans / TransIFS, which counts the
    // m > 0 coefficients twice (test_transgeneral.cc:1683-1686), so the factor 2 of the m > 0 harmonics in
    // f = sum_n X_n^0 P + 2 Re sum_{m>0} ... cancels and every m gets the same factor.
    const double wq0 = adjoint ? 1.0 : weights[pair];
    const double inv_n = adjoint ? 1.0 : 1.0 / n;
    for (int w = tid; w < nfb * (L + 1); w += nthr) {
        const int m = w / nfb, fi = w - m * nfb;
        const double wq = wq0;
        double2 Gp = cmulc(X[fi * PL + swz(L + m)], A[L + m]);
        double2 Gm = cmulc(X[fi * PL + swz(L - m)], A[L - m]);
        Gp.x *= inv_n; Gp.y *= inv_n; Gm.x *= inv_n; Gm.y *= inv_n;
        // F_N = (G_m + conj(G_-m))/2 ; F_S = (G_m - conj(G_-m))/(2i)
        const double2 FN = make_double2(0.5 * (Gp.x + Gm.x), 0.5 * (Gp.y - Gm.y));
        const double2 FS = make_double2(0.5 * (Gp.y + Gm.y), -0.5 * (Gp.x - Gm.x));
        double2 s, a;
        if (pm.has_s) {
            s = make_double2((FN.x + FS.x) * wq, (FN.y + FS.y) * wq);
            a = make_double2((FN.x - FS.x) * wq, (FN.y - FS.y) * wq);
        }
        else {
            s = make_double2(FN.x * wq, FN.y * wq);
            a = s;
        }
        const int n0 = nlat0[m];
        const long long is = (fb_rowoff[m] + (pair - n0)) * nf + f0 + fi;
        const long long ia = is + static_cast<long long>(nleg - n0) * nf;
        fb[is] = s;
        fb[ia] = a;
    }
    if (owner)  // sharded plan: completed pairs are shipped to the owners of their zonal wavenumbers
        finish_pair_and_push(pair, (nf + pm.F - 1) / pm.F, pair_done, nleg, meta, nf, fb, fb_rowoff, nlat0, owner, dst, me);
}

// ---- direct mode (mode 3): rows whose length has no prime factor above 23 need no chirp-z ----
// x_i + i y_i = sum_{|m| <= L} Z_m e^{2 pi i m i / n} is ONE unnormalised inverse DFT of length n with Z_m stored at
// frequency m mod n (2L < n: no overlap), instead of two transforms of length M >= n + 2L and three pointwise products.
// The transforms are the same shared-memory passes (fft_core.cuh), with the prime radices 7 .. 23 next to 2 .. 16; the
// digit-reversed slot of every frequency comes from a table built with the plan (dif_output_position).
__global__ void __launch_bounds__(kFftThreads)
direct_tables_kernel(const PairMeta* __restrict__ cls, double2* __restrict__ chirp, double2* __restrict__ twid) {
    const PairMeta pm = cls[blockIdx.x];
    const int M = pm.M, tid = threadIdx.x, nthr = blockDim.x;
    double2* Wa = twid + pm.tw_off;
    double2* Wb = Wa + (M / 64 + 1);
    for (int k = tid; k <= M / 64; k += nthr) {
        double s, c;
        sincospi(-2.0 * (64.0 * k) / M, &s, &c);
        Wa[k] = make_double2(c, s);
    }
    for (int k = tid; k < 64; k += nthr) {
        double s, c;
        sincospi(-2.0 * k / M, &s, &c);
        Wb[k] = make_double2(c, s);
    }
    __shared__ ScheduleG sc;
    if (tid == 0) sc = make_schedule_g(M);
    __syncthreads();
    int* pos = reinterpret_cast<int*>(chirp + pm.chirp_off);
    for (int k = tid; k < M; k += nthr) pos[k] = swz(dif_output_position(sc, M, k));
}

// Launch shapes of the direct kernels: (threads, resident blocks per SM the register budget is set for).  The passes are
// short (one or two butterflies per thread), so what matters is how many blocks in different phases share an SM: the global
// gathers of one block run under the butterflies of the others.
struct DirectArgs {
    const PairMeta* meta;
    int nf, mlimit, nb_uv;
    const long long* fb_rowoff;
    const int* nlat0;
    int nleg;
    const ScheduleG* scheds;
    const double2* twid;
    const double2* chirp;
    const double* scale_lat;   // inverse: 1/cos(lat); direct: wind scaling (per latitude pair)
    const double* weights;
    double2* fb;
    double* gp;
    long long npts;
    int adjoint;
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
fourier_inv_direct_kernel(const __grid_constant__ DirectArgs a, const int2* __restrict__ blocks) {
    extern __shared__ double2 X[];
    const int2 bd = blocks[blockIdx.x];
    const int pair = bd.x, f0 = bd.y & 0xffff, nfb = bd.y >> 16;
    const PairMeta pm = a.meta[pair];
    const int tid = threadIdx.x;
    constexpr int nthr = NT;
    const int n = pm.n, L = pm.L, PL = swz_len(n), nf = a.nf;
    const long long npts = a.npts;
    double* __restrict__ gp = a.gp;
    const int Lc = min(L, a.mlimit);
    if (Lc < 0) {
        for (int w = tid; w < nfb * n; w += nthr) {
            const int fi = w / n, i = w - fi * n;
            gp[(f0 + fi) * npts + pm.rowN + i] = 0.;
            if (pm.has_s) gp[(f0 + fi) * npts + pm.rowS + i] = 0.;
        }
        return;
    }
    double2* sW = X + pm.F * PL;
    __shared__ ScheduleG sc;
    load_schedule(a.scheds + pm.sched, &sc, tid);
    for (int e = tid; e < nfb * PL; e += nthr) X[e] = make_double2(0., 0.);
    load_twiddles(a.twid + pm.tw_off, sW, n, tid, nthr);
    __syncthreads();
    const int* __restrict__ pos = reinterpret_cast<const int*>(a.chirp + pm.chirp_off);
    const double2* __restrict__ fb = a.fb;
    const int tot = nfb * (Lc + 1);
    const unsigned nfb_magic = fastdiv_magic(static_cast<unsigned>(nfb));  // (tot < 65536)
    constexpr int kU = 4;  // independent gathers in flight per thread
    for (int w0 = tid; w0 < tot; w0 += kU * nthr) {
        double2 cs[kU], ca[kU];
        int pp[kU], pq[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int w = w0 + u * nthr;
            if (w < tot) {
                const int m = fastdiv(w, nfb_magic, nfb), fi = w - m * nfb;
                const int n0 = a.nlat0[m];
                const long long is = (a.fb_rowoff[m] + (pair - n0)) * nf + f0 + fi;
                cs[u] = fb[is];
                ca[u] = fb[is + static_cast<long long>(a.nleg - n0) * nf];
                pp[u] = pos[m];
                pq[u] = pos[m ? n - m : 0];
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int w = w0 + u * nthr;
            if (w < tot) {
                const int m = fastdiv(w, nfb_magic, nfb), fi = w - m * nfb;
                if (m == 0) cs[u].y = ca[u].y = 0.;
                double2 FN, FS;
                if (pm.has_s) {
                    FN = cadd(cs[u], ca[u]);
                    FS = csub(cs[u], ca[u]);
                }
                else {
                    FN = csub(cs[u], ca[u]);
                    FS = make_double2(0., 0.);
                }
                X[fi * PL + pp[u]] = make_double2(FN.x - FS.y, FN.y + FS.x);
                if (m > 0) X[fi * PL + pq[u]] = make_double2(FN.x + FS.y, FS.x - FN.y);
            }
        }
    }
    __syncthreads();
    fft_dit_g<false, true>(X, nfb, n, sc, sW, sW + (n / 64 + 1), nullptr, tid, nthr, PL);
    for (int fi = 0; fi < nfb; ++fi) {
        const int f = f0 + fi;
        const double sn = (f < a.nb_uv) ? a.scale_lat[pair] : 1.;
        double* __restrict__ gN = gp + f * npts + pm.rowN;
        double* __restrict__ gS = gp + f * npts + pm.rowS;
        for (int i = tid; i < n; i += nthr) {
            const double2 z = X[fi * PL + swz(i)];
            gN[i] = z.x * sn;
            if (pm.has_s) gS[i] = z.y * sn;
        }
    }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
fourier_dir_direct_kernel(const __grid_constant__ DirectArgs a, const int2* __restrict__ blocks, const int* __restrict__ owner,
                          const __grid_constant__ PeerDst dst, int me, int* __restrict__ pair_done) {
    extern __shared__ double2 X[];
    const int2 bd = blocks[blockIdx.x];
    const int pair = bd.x, f0 = bd.y & 0xffff, nfb = bd.y >> 16;
    const PairMeta pm = a.meta[pair];
    const int tid = threadIdx.x;
    constexpr int nthr = NT;
    const int n = pm.n, L = pm.L, PL = swz_len(n), nf = a.nf;
    if (L < 0) return;
    const long long npts = a.npts;
    double2* sW = X + pm.F * PL;
    __shared__ ScheduleG sc;
    // the two rows of every field go straight into the interleaved (north, south) sequence: 8-byte asynchronous copies, all in
    // flight at once
    for (int fi = 0; fi < nfb; ++fi) {
        const double* __restrict__ gN = a.gp + (f0 + fi) * npts + pm.rowN;
        const double* __restrict__ gS = a.gp + (f0 + fi) * npts + pm.rowS;
        double* Xf = reinterpret_cast<double*>(X + fi * PL);
        if (pm.has_s) {
            for (int i = tid; i < n; i += nthr) {
                double* d = Xf + 2 * swz(i);
                cp_async8(d, gN + i);
                cp_async8(d + 1, gS + i);
            }
        }
        else {
            for (int i = tid; i < n; i += nthr) X[fi * PL + swz(i)] = make_double2(gN[i], 0.);
        }
    }
    load_schedule(a.scheds + pm.sched, &sc, tid);
    load_twiddles(a.twid + pm.tw_off, sW, n, tid, nthr);
    cp_async_wait_all();
    __syncthreads();
    if (f0 < a.nb_uv) {  // wind components enter as u, v times the latitude's scaling
        const double sl = a.scale_lat[pair];
        const int nuv = min(nfb, a.nb_uv - f0);
        for (int fi = 0; fi < nuv; ++fi)
            for (int i = tid; i < n; i += nthr) {
                double2 v = X[fi * PL + swz(i)];
                v.x *= sl;
                v.y *= sl;
                X[fi * PL + swz(i)] = v;
            }
        __syncthreads();
    }
    fft_dif_g<true>(X, nfb, n, sc, sW, sW + (n / 64 + 1), tid, nthr, PL);
    const int* __restrict__ pos = reinterpret_cast<const int*>(a.chirp + pm.chirp_off);
    double2* __restrict__ fb = a.fb;
    // (weights and conventions: see fourier_dir_kernel)
    const double wq = a.adjoint ? 1.0 : a.weights[pair];
    const double hw = 0.5 * wq * (a.adjoint ? 1.0 : 1.0 / n);
    const int tot = nfb * (L + 1);
    const unsigned nfb_magic = fastdiv_magic(static_cast<unsigned>(nfb));
    constexpr int kU = 4;
    for (int w0 = tid; w0 < tot; w0 += kU * nthr) {
        long long is[kU], ia[kU];
        int pp[kU], pq[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int w = w0 + u * nthr;
            if (w < tot) {
                const int m = fastdiv(w, nfb_magic, nfb), fi = w - m * nfb;
                const int n0 = a.nlat0[m];
                is[u] = (a.fb_rowoff[m] + (pair - n0)) * nf + f0 + fi;
                ia[u] = is[u] + static_cast<long long>(a.nleg - n0) * nf;
                pp[u] = pos[m];
                pq[u] = pos[m ? n - m : 0];
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int w = w0 + u * nthr;
            if (w < tot) {
                const int m = fastdiv(w, nfb_magic, nfb), fi = w - m * nfb;
                const double2 Gp = X[fi * PL + pp[u]], Gm = X[fi * PL + pq[u]];
                // F_N = (G_m + conj(G_-m))/2 ; F_S = (G_m - conj(G_-m))/(2i), each times weight / n
                const double2 FN = make_double2(hw * (Gp.x + Gm.x), hw * (Gp.y - Gm.y));
                const double2 FS = make_double2(hw * (Gp.y + Gm.y), -hw * (Gp.x - Gm.x));
                double2 sy, as;
                if (pm.has_s) {
                    sy = cadd(FN, FS);
                    as = csub(FN, FS);
                }
                else sy = as = FN;
                fb[is[u]] = sy;
                fb[ia[u]] = as;
            }
        }
    }
    if (owner)  // sharded plan: see fourier_dir_kernel
        finish_pair_and_push(pair, (nf + pm.F - 1) / pm.F, pair_done, a.nleg, a.meta, nf, fb, a.fb_rowoff, a.nlat0, owner, dst, me);
}

// (threads, blocks per SM) shapes compiled for the direct kernels; SPTRANS_FFTD_SHAPE picks one for the groups that fit it
#define SPT_DIRECT_SHAPES(X) X(512, 1) X(256, 2) X(128, 4)

// ---- row mode (mode 1): rows too long for the packed north/south transform (n + 2L > kMaxM, e.g. O2560) ----
// A real row x_i, i < n, is transformed through z'_k = x_2k + i x_2k+1 (k < n/2):
//   z'_k = sum_{m=-L..L} G_m e^{2 pi i m k/(n/2)},   G_m = F_m (1 + i w^m),  F_-m = conj(F_m),  w = e^{2 pi i/n}
// i.e. the same chirp-z machinery with length n/2; the direct transform inverts the packing with
//   F_m = 1/2 [ (H_m + conj H_-m)/2 + conj(w^m) (H_m - conj H_-m)/(2i) ],  H = chirp-z spectrum of z'.
__global__ void __launch_bounds__(2 * kFftThreads, 1)
fourier_inv_rows_kernel(const PairMeta* __restrict__ meta, const int2* __restrict__ blocks, int nf, int mlimit,
                        int nb_uv, const double2* __restrict__ fb, const long long* __restrict__ fb_rowoff,
                        const int* __restrict__ nlat0, int nleg, const ScheduleG* __restrict__ scheds, const double2* __restrict__ twid,
                        const double2* __restrict__ chirp, const double2* __restrict__ filt,
                        const double* __restrict__ coslatinv, double* __restrict__ gp, long long npts) {
    extern __shared__ double2 X[];
    const int2 bd = blocks[blockIdx.x];
    const int pair = bd.x, f = bd.y & 0xffff;
    const PairMeta pm = meta[pair];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int L = pm.L, M = pm.M, nh = pm.n / 2;
    const int Lc = min(L, mlimit);
    double2* sW = X + M;
    __shared__ ScheduleG sc;
    load_schedule(scheds + pm.sched, &sc, tid);
    load_twiddles(twid + pm.tw_off, sW, M, tid, nthr);
    const double2* A = chirp + pm.chirp_off;
    const double2* C = A + (2 * L + 1);
    const double2* Wr = C + nh;
    const double scale = f < nb_uv ? coslatinv[pair] : 1.;
    for (int row = 0; row < (pm.has_s ? 2 : 1); ++row) {
        __syncthreads();
        for (int e = tid; e < M; e += nthr) X[e] = make_double2(0., 0.);
        __syncthreads();
        for (int m = tid; m <= Lc; m += nthr) {
            const int n0 = nlat0[m];
            const long long is = (fb_rowoff[m] + (pair - n0)) * nf + f;
            double2 cs = fb[is], ca = fb[is + static_cast<long long>(nleg - n0) * nf];
            if (m == 0) cs.y = ca.y = 0.;
            // northern row: sym + asym, southern: sym - asym; the single equator row: sym - asym (reference :1061-1070)
            const bool minus = (row == 1) || !pm.has_s;
            const double2 Fm = minus ? csub(cs, ca) : cadd(cs, ca);
            const double2 w = Wr[m];
            // G_m = F_m (1 + i w^m),  G_-m = conj(F_m) (1 + i conj(w^m))
            const double2 gp1 = make_double2(1. - w.y, w.x);   // 1 + i w
            const double2 gm1 = make_double2(1. + w.y, w.x);   // 1 + i conj(w)
            X[swz(L + m)] = cmul(cmul(Fm, gp1), A[L + m]);
            if (m > 0) X[swz(L - m)] = cmul(cmul(cconj(Fm), gm1), A[L - m]);
        }
        __syncthreads();
        fft_dif_g(X, 1, M, sc, sW, sW + (M / 64 + 1), tid, nthr);
        fft_dit_g<false>(X, 1, M, sc, sW, sW + (M / 64 + 1), filt + pm.filt_off, tid, nthr);
        double* out = gp + f * npts + (row == 0 ? pm.rowN : pm.rowS);
        for (int k = tid; k < nh; k += nthr) {
            const double2 z = cmul(X[swz(k)], C[k]);
            *reinterpret_cast<double2*>(out + 2 * k) = make_double2(z.x * scale, z.y * scale);
        }
    }
}

__global__ void __launch_bounds__(2 * kFftThreads, 1)
fourier_dir_rows_kernel(const PairMeta* __restrict__ meta, const int2* __restrict__ blocks, int nf, int nb_uv,
                        const double* __restrict__ gp, long long npts, const long long* __restrict__ fb_rowoff,
                        const int* __restrict__ nlat0, int nleg, const ScheduleG* __restrict__ scheds, const double2* __restrict__ twid,
                        const double2* __restrict__ chirp, const double2* __restrict__ filt,
                        const double* __restrict__ weights, const double* __restrict__ uvscale,
                        double2* __restrict__ fb, int adjoint) {
    extern __shared__ double2 X[];
    const int2 bd = blocks[blockIdx.x];
    const int pair = bd.x, f = bd.y & 0xffff;
    const PairMeta pm = meta[pair];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int L = pm.L, M = pm.M, nh = pm.n / 2;
    if (L < 0) return;
    double2* sW = X + M;
    __shared__ ScheduleG sc;
    load_schedule(scheds + pm.sched, &sc, tid);
    load_twiddles(twid + pm.tw_off, sW, M, tid, nthr);
    const double2* A = chirp + pm.chirp_off;
    const double2* C = A + (2 * L + 1);
    const double2* Wr = C + nh;
    const double wq0 = adjoint ? static_cast<double>(pm.n) : weights[pair];  // adjoint: undo the 1/n of F_m below
    const double scale = f < nb_uv ? uvscale[pair] : 1.;
    const double inv_nh = 1.0 / nh;
    for (int row = 0; row < (pm.has_s ? 2 : 1); ++row) {
        __syncthreads();
        for (int e = tid; e < M; e += nthr) X[e] = make_double2(0., 0.);
        __syncthreads();
        const double* in = gp + f * npts + (row == 0 ? pm.rowN : pm.rowS);
        for (int k = tid; k < nh; k += nthr) {
            const double2 x = *reinterpret_cast<const double2*>(in + 2 * k);
            X[swz(k)] = cmulc(make_double2(x.x * scale, x.y * scale), C[k]);
        }
        __syncthreads();
        fft_dif_g(X, 1, M, sc, sW, sW + (M / 64 + 1), tid, nthr);
        fft_dit_g<true>(X, 1, M, sc, sW, sW + (M / 64 + 1), filt + pm.filt_off, tid, nthr);
        for (int m = tid; m <= L; m += nthr) {
            double2 Hp = cmulc(X[swz(L + m)], A[L + m]);
            double2 Hm = cmulc(X[swz(L - m)], A[L - m]);
            Hp.x *= inv_nh; Hp.y *= inv_nh; Hm.x *= inv_nh; Hm.y *= inv_nh;
            // E/n' = (H_m + conj H_-m)/2 ; O/n' = (H_m - conj H_-m)/(2i) ; F_m = (E + conj(w^m) O)/n
            const double2 Ev = make_double2(0.5 * (Hp.x + Hm.x), 0.5 * (Hp.y - Hm.y));
            const double2 Ov = make_double2(0.5 * (Hp.y + Hm.y), -0.5 * (Hp.x - Hm.x));
            const double2 t = cmulc(Ov, Wr[m]);
            const double wq = wq0;  // adjoint: every m alike (spectral inner product counts m > 0 twice, see fourier_dir_kernel)
            const double2 Fm = make_double2(0.5 * (Ev.x + t.x) * wq, 0.5 * (Ev.y + t.y) * wq);
            const int n0 = nlat0[m];
            const long long is = (fb_rowoff[m] + (pair - n0)) * nf + f;
            const long long ia = is + static_cast<long long>(nleg - n0) * nf;
            if (row == 0) {        // northern row first: park w F_N in both slots
                fb[is] = Fm;
                fb[ia] = Fm;
            }
            else {                 // southern row (same thread, same m): sym = w (F_N + F_S), asym = w (F_N - F_S)
                const double2 fnw = fb[is];
                fb[is] = cadd(fnw, Fm);
                fb[ia] = csub(fnw, Fm);
            }
        }
    }
}

// ---- v2 kernels: register-tiled chirp-z (fft2_core.cuh) ---------------------------------------------------------------
// One block = one latitude pair x a run of fields (transformed one after the other, the next field's inputs staged by
// cp.async behind the current transforms).  256 threads and one block per SM for M1 >= 18 (a radix-M1 butterfly
// needs 4 M1 registers for its data alone); 128 threads and two blocks per SM for M1 <= 16.
__host__ __device__ constexpr int v2_threads(int M1) { return M1 <= 16 ? 128 : 256; }
__host__ __device__ constexpr int v2_min_blocks(int M1) { return M1 <= 16 ? 2 : 1; }

// per class: chirps A_u, C_i and W1[t] = e^{-2 pi i t/M}; block 0 also writes the shared T256 table
__global__ void __launch_bounds__(256)
chirp_tables2_kernel(const PairMeta* __restrict__ cls, double2* __restrict__ chirp, double2* __restrict__ twid,
                     double2* __restrict__ t256) {
    const PairMeta pm = cls[blockIdx.x];
    const int L = pm.L, n = pm.n, M = pm.M;
    const int tid = threadIdx.x, nthr = blockDim.x;
    double2* A = chirp + pm.chirp_off;
    double2* C = A + (2 * L + 1);
    for (int u = tid; u <= 2 * L; u += nthr) {
        double s, c;
        sincospi(static_cast<double>(chirp_residue(u, 0, n)) / n, &s, &c);
        A[u] = make_double2(c, s);
    }
    for (int i = tid; i < n; i += nthr) {
        double s, c;
        sincospi(static_cast<double>(chirp_residue(i, -2LL * L, n)) / n, &s, &c);
        C[i] = make_double2(c, s);
    }
    for (int t = tid; t < fft2::kM2; t += nthr) {
        double s, c;
        sincospi(-2.0 * t / M, &s, &c);
        twid[pm.tw_off + t] = make_double2(c, s);
    }
    if (blockIdx.x == 0)
        for (int e = tid; e < 256; e += nthr) {
            double s, c;
            sincospi(-2.0 * ((e >> 4) * (e & 15)) / 256.0, &s, &c);
            t256[e] = make_double2(c, s);
        }
}

// filter spectrum of one class, in the register order of the forward machinery (runs after chirp_tables2_kernel)
__global__ void __launch_bounds__(256, 1)
filter_tables2_kernel(const PairMeta* __restrict__ cls, const double2* __restrict__ twid,
                      const double2* __restrict__ t256, double2* __restrict__ filt) {
    extern __shared__ double2 X[];
    const PairMeta pm = cls[blockIdx.x];
    const double2* W1 = twid + pm.tw_off;
    double2* out = filt + pm.filt_off;
    const int tid = threadIdx.x;
    switch (pm.m1) {
#define SPT_CASE(R) case R: fft2::filter_table_body<R, 256>(pm, tid, X, W1, t256, out); break;
        SPT_CASE(8) SPT_CASE(9) SPT_CASE(10) SPT_CASE(12) SPT_CASE(15) SPT_CASE(16) SPT_CASE(18) SPT_CASE(20)
        SPT_CASE(24) SPT_CASE(25) SPT_CASE(27) SPT_CASE(30) SPT_CASE(32)
#undef SPT_CASE
        default: break;
    }
}

template <int M1>
__global__ void __launch_bounds__(v2_threads(M1), v2_min_blocks(M1))
fourier2_inv_kernel(const __grid_constant__ Fft2Args a, const int2* __restrict__ blocks) {
    extern __shared__ double2 X[];
    const int2 bd = blocks[blockIdx.x];
    fft2::fourier2_inv_body<M1, v2_threads(M1)>(a, bd.x, bd.y & 0xffff, bd.y >> 16, threadIdx.x, X);
}

template <int M1>
__global__ void __launch_bounds__(v2_threads(M1), v2_min_blocks(M1))
fourier2_dir_kernel(const __grid_constant__ Fft2Args a, const int2* __restrict__ blocks, const int* __restrict__ owner,
                    const __grid_constant__ PeerDst dst, int me, int* __restrict__ pair_done) {
    extern __shared__ double2 X[];
    const int2 bd = blocks[blockIdx.x];
    fft2::fourier2_dir_body<M1, v2_threads(M1)>(a, bd.x, bd.y & 0xffff, bd.y >> 16, threadIdx.x, X);
    if (owner)  // sharded plan: completed pairs are shipped to the owners of their zonal wavenumbers (see v1)
        finish_pair_and_push(bd.x, (a.nf + a.F - 1) / a.F, pair_done, a.nleg, a.meta, a.nf, a.fb, a.fb_rowoff, a.nlat0, owner, dst, me);
}

// Cost model used to pick the convolution length: every pass is one read+write sweep of shared memory;
// larger radices do more arithmetic per point.
double pass_cost(int R) {
    // measured (profiles/ncu_summary_r01.md): the kernels are issue/latency bound with a barrier per pass, so a pass
    // costs about the same whatever its radix; the radix only adds arithmetic
    switch (R) {
        case 16: return 1.32;
        case 9: return 1.25;
        case 8: return 1.24;
        case 5: return 1.19;
        case 4: return 1.16;
        case 3: return 1.13;
        default: return 1.08;
    }
}
// SPTRANS_FFT_MAXM (testing only) lowers the single-CTA limit so that small grids exercise the row-mode kernels
int max_conv_length() {
    static int v = [] {
        const char* e = std::getenv("SPTRANS_FFT_MAXM");
        const int x = e ? std::atoi(e) : 0;
        return (x >= 64 && x <= kMaxM) ? x : kMaxM;
    }();
    return v;
}
int choose_conv_length(int need) {
    int best = 0;
    double best_cost = 1e300;
    const long long lim = max_conv_length();
    for (long long p5 = 1; p5 <= lim; p5 *= 5)
        for (long long p3 = p5; p3 <= lim; p3 *= 3)
            for (long long p2 = p3 * 8; p2 <= lim; p2 *= 2) {  // multiples of 8 (index swizzle)
                if (p2 < need) continue;
                const ScheduleG s = fftc::make_schedule_g(static_cast<int>(p2));
                double c = 0.4;  // load/store/pointwise sweeps
                for (int p = 0; p < s.npass; ++p) c += 2.0 * pass_cost(s.radix[p]);  // forward + inverse
                c *= static_cast<double>(p2);
                if (c < best_cost) {
                    best_cost = c;
                    best = static_cast<int>(p2);
                }
            }
    return best;
}

int fields_per_block(int M) {
    // enough sequences per block to keep 256 threads busy, at most ~64 KB of shared memory
    int F = 4096 / M;
    if (F < 1) F = 1;
    if (F > 16) F = 16;
    return F;
}

size_t block_smem_bytes(int M, int F) {
    return (static_cast<size_t>(F) * M + (M / 64 + 1 + 64)) * sizeof(double2);
}
size_t direct_smem_bytes(int n, int F) {
    return (static_cast<size_t>(F) * fftc::swz_len(n) + (n / 64 + 1 + 64)) * sizeof(double2);
}

}  // namespace

struct FftGroups {
    // launch groups: v1 by shared-memory footprint (so that small rows run several blocks per SM), v2 by radix
    std::vector<size_t> smem;              // dynamic shared memory of the group (v2: inverse kernel)
    std::vector<size_t> smem_dir;          // v2: direct kernel (its staging area holds two grid rows)
    std::vector<std::vector<int>> pairs;   // pairs per group, costliest first
    std::vector<int> mode;                 // 0: packed north/south kernels, 1: row kernels, 2: v2 kernels, 3: direct kernels
    std::vector<int> m1;                   // v2: block-level radix of the group
    int nf = -1;                           // block lists below are built for this number of fields
    int nchunks = 0;                       // ... split into this many field chunks (host-pointer pipelines; 1 otherwise)
    int F2 = 1;                            // v2: fields per block for this nf
    std::vector<int2*> d_blocks;           // per group: [chunk][pair][field run] descriptors (pair, f0 | count << 16)
    std::vector<int> nblocks;
    std::vector<std::vector<int>> chunk_begin;  // per group: first block of every chunk (+ end)
    std::vector<int> chunk_field;          // [nchunks + 1] field boundaries of the chunks
};
// Launch groups of one Fourier stage are independent (disjoint latitude pairs); with SPTRANS_FFT_STREAMS = N > 1 (default 2:
// measured 12.2 -> 11.6 ms inverse, 12.7 -> 12.2 ms direct at TCo1279 L137; more streams add nothing) they are
// spread round-robin over N streams (fork / join with events around the stage), so that the blocks of the next group
// fill the SMs the tail of the previous group leaves idle (every group is a 20-30 wave launch at one block per SM).
struct StreamFan {
    std::vector<cudaStream_t> aux;
    std::vector<cudaEvent_t> done;
    cudaEvent_t fork = nullptr;
};
// Per-plan state of the Fourier stage, owned by the plan (Plan::fft): no process-global tables, so threads working on
// different plans never touch shared host state (the contract stays "one in-flight call per plan", like TransLocal).
struct FftState {
    FftGroups grp;
    std::vector<PairMeta> meta;
    double2* t256 = nullptr;
    StreamFan fan;
};
static FftState& fft_state(Plan& p) {
    if (!p.fft) p.fft = new FftState();
    return *static_cast<FftState*>(p.fft);
}

namespace {
int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}
// X | T256 | S | row table (int per exchange-buffer row of the pair)
size_t v2_smem_inv(int M, int L) {
    return (static_cast<size_t>(M) + 256 + 2 * (L + 1)) * sizeof(double2) + 2 * (L + 1) * sizeof(int);
}
size_t v2_smem_dir(int M, int n, int L) {
    return (static_cast<size_t>(M) + 256) * sizeof(double2) + 2 * static_cast<size_t>(n) * sizeof(double) + 2 * (L + 1) * sizeof(int);
}
constexpr size_t kSmemLimit = 226 * 1024;  // 227 KB opt-in maximum minus the kernels' few bytes of static shared memory

template <int M1>
int v2_set_attributes() {
    SPT_CUDA(cudaFuncSetAttribute(fourier2_inv_kernel<M1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    SPT_CUDA(cudaFuncSetAttribute(fourier2_dir_kernel<M1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemLimit));
    return SPTRANS_OK;
}
}  // namespace

#define SPT_V2_RADICES(X) X(8) X(9) X(10) X(12) X(15) X(16) X(18) X(20) X(24) X(25) X(27) X(30) X(32)

int build_fft_tables(Plan& p) {
    HostGeom& g = p.g;
    const int nleg = g.nleg;
    std::vector<PairMeta> meta(nleg);
    std::map<std::pair<int, int>, int> cls_index;
    std::vector<PairMeta> classes;
    long long chirp_total = 0, filt_total = 0, tw_total = 256;  // twiddle slot 0..255: the shared T256 table of v2
    // v2 (register-tiled) kernels: rows of even length (16-byte row starts for the cp.async staging) whose
    // convolution fits M1 * 256; everything else (short rows, rows beyond 8192) stays on the v1 kernels
    bool all_even = true;
    for (int j = 0; j < g.nlat; ++j) all_even = all_even && (g.nx[j] % 2 == 0);
    // (the kernels keep exchange-buffer row indices as 32-bit integers)
    const bool use_v2 = env_int("SPTRANS_FFT_V2", 1) != 0 && all_even && g.fb_rowoff.back() < 2147483647LL;
    const int v2_min_need = env_int("SPTRANS_FFT2_MIN", 1281);
    // direct transforms (no chirp-z) of the rows whose length has no prime factor above 23; SPTRANS_FFT_DIRECT=0 sends every row through
    // the chirp-z kernels, and so does the test knob SPTRANS_FFT_MAXM unless SPTRANS_FFT_DIRECT=1 is set with it
    const bool use_direct = env_int("SPTRANS_FFT_DIRECT", std::getenv("SPTRANS_FFT_MAXM") ? 0 : 1) != 0;
    for (int j = 0; j < nleg; ++j) {
        PairMeta pm{};
        pm.n = g.nx[j];
        pm.L = g.mmax[j];
        pm.rowN = g.gp_rowoff[j];  // (global offsets, or offsets into this rank's band with SPTRANS_SHARD_LOCAL_IO)
        pm.rowS = g.gp_rowoff[g.nlat - 1 - j];
        pm.has_s = (g.nlat - 1 - j != j) ? 1 : 0;
        if (pm.L >= 0 && 2 * pm.L >= pm.n) {
            set_error("sptrans_plan_create: zonal truncation at a latitude row reaches nx/2 (aliasing); unsupported");
            return SPTRANS_ERR_INVALID;
        }
        const int Luse = std::max(pm.L, 0);
        int M = 0;
        pm.mode = 0;
        pm.m1 = 0;
        if (use_direct && pm.L >= 0 && pm.n <= kMaxM && fftc::is_direct_length(pm.n)) {
            M = pm.n;
            pm.mode = 3;
        }
        if (M == 0 && use_v2 && pm.L >= 0 && pm.n + 2 * Luse >= v2_min_need) {
            int m1 = 0;
            const int M2 = fft2::conv_length_v2(pm.n + 2 * Luse, &m1);
            if (M2 && v2_smem_inv(M2, Luse) <= kSmemLimit && v2_smem_dir(M2, pm.n, Luse) <= kSmemLimit) {
                M = M2;
                pm.m1 = m1;
            }
        }
        if (M == 0) M = choose_conv_length(pm.n + 2 * Luse);
        if (M == 0 && pm.n % 2 == 0) {  // too long for the packed transform: one row at a time, length n/2
            M = choose_conv_length(pm.n / 2 + 2 * Luse);
            pm.mode = 1;
        }
        if (M == 0) {
            set_error("sptrans_plan_create: row length too large for the shared-memory Fourier kernels "
                      "(n/2 + 2*truncation must not exceed 13824)");
            return SPTRANS_ERR_NOT_IMPLEMENTED;
        }
        pm.M = M;
        pm.F = (pm.mode == 1 || pm.m1) ? 1 : fields_per_block(M);
        if (pm.mode == 3) pm.F = std::max(1, std::min(16, env_int("SPTRANS_FFTD_POINTS", 4096) / pm.n));
        auto key = std::make_pair(pm.n, pm.mode == 3 ? -3 : Luse);  // (the direct tables do not depend on the truncation)
        auto it = cls_index.find(key);
        if (it == cls_index.end()) {
            PairMeta c = pm;
            c.L = Luse;
            c.chirp_off = chirp_total;
            c.filt_off = filt_total;
            c.tw_off = tw_total;
            c.sched = static_cast<int>(classes.size());
            if (pm.mode == 3) chirp_total += (pm.n + 3) / 4;  // swizzled slot of every frequency, n ints
            else {
                chirp_total += 2LL * Luse + 1 + (pm.mode ? pm.n / 2 + Luse + 1 : pm.n);
                filt_total += M;
            }
            tw_total += pm.m1 ? fft2::kM2 : M / 64 + 1 + 64;
            cls_index[key] = static_cast<int>(classes.size());
            classes.push_back(c);
            it = cls_index.find(key);
        }
        pm.chirp_off = classes[it->second].chirp_off;
        pm.filt_off = classes[it->second].filt_off;
        pm.tw_off = classes[it->second].tw_off;
        pm.sched = classes[it->second].sched;
        meta[j] = pm;
    }
    SPT_CUDA(cudaMalloc(&p.d_chirp, std::max<long long>(chirp_total, 1) * sizeof(double2)));
    SPT_CUDA(cudaMalloc(&p.d_filt, std::max<long long>(filt_total, 1) * sizeof(double2)));
    SPT_CUDA(cudaMalloc(&p.d_twiddle, std::max<long long>(tw_total, 1) * sizeof(double2)));
    p.bytes_tables += (chirp_total + filt_total + tw_total) * sizeof(double2);
    fft_state(p).t256 = p.d_twiddle;
    {
        std::vector<ScheduleG> sch(classes.size());
        for (size_t c = 0; c < classes.size(); ++c)
            if (!classes[c].m1) sch[c] = fftc::make_schedule_g(classes[c].M);
        ScheduleG* d_s = nullptr;
        SPT_CUDA(cudaMalloc(&d_s, std::max<size_t>(sch.size(), 1) * sizeof(ScheduleG)));
        SPT_CUDA(cudaMemcpyAsync(d_s, sch.data(), sch.size() * sizeof(ScheduleG), cudaMemcpyHostToDevice, p.stream));
        SPT_CUDA(cudaStreamSynchronize(p.stream));
        p.d_fft_order = reinterpret_cast<int*>(d_s);  // owned by the plan (freed with it)
    }
    std::vector<PairMeta> cls1, cls2, cls3;
    for (const PairMeta& c : classes) (c.mode == 3 ? cls3 : c.m1 ? cls2 : cls1).push_back(c);
    PairMeta *d_cls1 = nullptr, *d_cls2 = nullptr, *d_cls3 = nullptr;
    const size_t smem_max = block_smem_bytes(kMaxM, 1);
    SPT_CUDA(cudaFuncSetAttribute(chirp_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    SPT_CUDA(cudaFuncSetAttribute(fourier_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    SPT_CUDA(cudaFuncSetAttribute(fourier_dir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    SPT_CUDA(cudaFuncSetAttribute(fourier_inv_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    SPT_CUDA(cudaFuncSetAttribute(fourier_dir_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
#define SPT_ATTR_D(NT, MB)                                                                                                       \
    SPT_CUDA(cudaFuncSetAttribute(fourier_inv_direct_kernel<NT, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max)); \
    SPT_CUDA(cudaFuncSetAttribute(fourier_dir_direct_kernel<NT, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    SPT_DIRECT_SHAPES(SPT_ATTR_D)
#undef SPT_ATTR_D
    if (!cls3.empty()) {
        SPT_CUDA(cudaMalloc(&d_cls3, cls3.size() * sizeof(PairMeta)));
        SPT_CUDA(cudaMemcpyAsync(d_cls3, cls3.data(), cls3.size() * sizeof(PairMeta), cudaMemcpyHostToDevice, p.stream));
        direct_tables_kernel<<<static_cast<int>(cls3.size()), kFftThreads, 0, p.stream>>>(d_cls3, p.d_chirp, p.d_twiddle);
        p.launches++;
        SPT_CUDA(cudaGetLastError());
    }
    if (!cls1.empty()) {
        SPT_CUDA(cudaMalloc(&d_cls1, cls1.size() * sizeof(PairMeta)));
        SPT_CUDA(cudaMemcpyAsync(d_cls1, cls1.data(), cls1.size() * sizeof(PairMeta), cudaMemcpyHostToDevice, p.stream));
        chirp_tables_kernel<<<static_cast<int>(cls1.size()), kFftThreads, smem_max, p.stream>>>(d_cls1, p.d_chirp, p.d_filt,
                                                                                             p.d_twiddle);
        p.launches++;
        SPT_CUDA(cudaGetLastError());
    }
    if (!cls2.empty()) {
        int rc = SPTRANS_OK;
#define SPT_ATTR(R) if (rc == SPTRANS_OK) rc = v2_set_attributes<R>();
        SPT_V2_RADICES(SPT_ATTR)
#undef SPT_ATTR
        if (rc) return rc;
        const size_t smem_f = static_cast<size_t>(32) * fft2::kM2 * sizeof(double2);
        SPT_CUDA(cudaFuncSetAttribute(filter_tables2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));
        SPT_CUDA(cudaMalloc(&d_cls2, cls2.size() * sizeof(PairMeta)));
        SPT_CUDA(cudaMemcpyAsync(d_cls2, cls2.data(), cls2.size() * sizeof(PairMeta), cudaMemcpyHostToDevice, p.stream));
        chirp_tables2_kernel<<<static_cast<int>(cls2.size()), 256, 0, p.stream>>>(d_cls2, p.d_chirp, p.d_twiddle, p.d_twiddle);
        SPT_CUDA(cudaGetLastError());
        filter_tables2_kernel<<<static_cast<int>(cls2.size()), 256, smem_f, p.stream>>>(d_cls2, p.d_twiddle, p.d_twiddle,
                                                                                       p.d_filt);
        SPT_CUDA(cudaGetLastError());
        p.launches += 2;
    }
    SPT_CUDA(cudaMalloc(&p.d_pair_meta, meta.size() * sizeof(PairMeta)));
    SPT_CUDA(cudaMemcpyAsync(p.d_pair_meta, meta.data(), meta.size() * sizeof(PairMeta), cudaMemcpyHostToDevice,
                             p.stream));
    // launch groups over this rank's latitude band
    FftGroups grp;
    const size_t buckets[] = {28 * 1024, 37 * 1024, 56 * 1024, 75 * 1024, 113 * 1024, smem_max};
    std::vector<std::vector<int>> by(6), by_direct(6);
    std::map<int, std::vector<int>> by_m1;
    std::vector<int> rowmode;
    for (int j = g.pair_begin; j < g.pair_end; ++j) {
        if (meta[j].m1) {
            by_m1[meta[j].m1].push_back(j);
            continue;
        }
        if (meta[j].mode == 1) {
            rowmode.push_back(j);
            continue;
        }
        const size_t need = meta[j].mode == 3 ? direct_smem_bytes(meta[j].n, meta[j].F) : block_smem_bytes(meta[j].M, meta[j].F);
        int b = 0;
        while (need > buckets[b]) ++b;
        (meta[j].mode == 3 ? by_direct : by)[b].push_back(j);
    }
    auto add_group = [&](std::vector<int> v, int mode, int m1) {
        std::stable_sort(v.begin(), v.end(), [&](int x, int y) {
            return meta[x].M != meta[y].M ? meta[x].M > meta[y].M : meta[x].n > meta[y].n;
        });
        size_t need = 0, need_dir = 0;
        for (int j : v) {
            if (mode == 2) {
                need = std::max(need, v2_smem_inv(meta[j].M, std::max(meta[j].L, 0)));
                need_dir = std::max(need_dir, v2_smem_dir(meta[j].M, meta[j].n, std::max(meta[j].L, 0)));
            }
            else if (mode == 3) need = std::max(need, direct_smem_bytes(meta[j].n, meta[j].F));
            else need = std::max(need, block_smem_bytes(meta[j].M, mode == 1 ? 1 : meta[j].F));
        }
        grp.smem.push_back(need);
        grp.smem_dir.push_back(mode == 2 ? need_dir : need);
        grp.pairs.push_back(v);
        grp.mode.push_back(mode);
        grp.m1.push_back(m1);
    };
    for (auto it = by_m1.rbegin(); it != by_m1.rend(); ++it) add_group(it->second, 2, it->first);
    if (!rowmode.empty()) add_group(rowmode, 1, 0);
    for (int b = 5; b >= 0; --b)
        if (!by_direct[b].empty()) add_group(by_direct[b], 3, 0);
    for (int b = 5; b >= 0; --b)
        if (!by[b].empty()) add_group(by[b], 0, 0);
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    cudaFree(d_cls1);
    cudaFree(d_cls2);
    cudaFree(d_cls3);
    fft_state(p).grp = grp;
    fft_state(p).meta = meta;
    return SPTRANS_OK;
}

// out[0..3]: grid points per field, out[4..7]: exchange-buffer rows (double2 per field) of the latitude pairs of this rank's
// band that run on the direct / register-tiled chirp-z / shared-memory-pass chirp-z / row-mode chirp-z kernels
void fourier_path_stats(Plan& p, long long* out) {
    for (int i = 0; i < 8; ++i) out[i] = 0;
    if (!p.fft) return;
    const std::vector<PairMeta>& meta = fft_state(p).meta;
    for (int j = p.g.pair_begin; j < p.g.pair_end && j < static_cast<int>(meta.size()); ++j) {
        const PairMeta& pm = meta[j];
        const int path = pm.mode == 3 ? 0 : pm.m1 ? 1 : pm.mode == 1 ? 3 : 2;
        out[path] += static_cast<long long>(pm.n) * (pm.has_s ? 2 : 1);
        out[4 + path] += 2LL * (pm.L + 1);
    }
}

void clone_fft_state(Plan& src, Plan& dst) {
    FftState& d = fft_state(dst);
    const FftState& s = fft_state(src);
    d.meta = s.meta;
    d.t256 = s.t256;
    d.grp = s.grp;
    d.grp.d_blocks.clear();   // block lists are per plan (built for the field count of its calls)
    d.grp.nblocks.clear();
    d.grp.chunk_begin.clear();
    d.grp.nf = -1;
    d.grp.nchunks = 0;
}

static void fan_free(Plan& p);
void free_fft_tables(Plan& p) {
    if (!p.fft) return;
    for (int2* d : fft_state(p).grp.d_blocks) cudaFree(d);
    fan_free(p);
    delete static_cast<FftState*>(p.fft);
    p.fft = nullptr;
}

// Block lists for `nf` fields split into `nchunks` contiguous field chunks.  Within a group the blocks are ordered
// [chunk][pair][field run], so that one chunk is a contiguous range of descriptors (launched on its own when the host
// pipelines overlap copies of one chunk with transforms of the next) and the whole group is still one launch otherwise.
static int ensure_block_lists(Plan& p, int nf, int nchunks = 1) {
    FftGroups& grp = fft_state(p).grp;
    if (nf > 0xffff) {
        set_error("fourier stage: more than 65535 fields in one call");
        return SPTRANS_ERR_INVALID;
    }
    nchunks = std::max(1, std::min(nchunks, nf));
    if (grp.nf == nf && grp.nchunks == nchunks) return SPTRANS_OK;
    for (int2* d : grp.d_blocks) cudaFree(d);
    grp.d_blocks.clear();
    grp.nblocks.clear();
    grp.chunk_begin.clear();
    const std::vector<PairMeta>& meta = fft_state(p).meta;
    // v2: fields per block, chosen so that the blocks of a pair carry (almost) equal numbers of fields
    const int ft = std::max(1, env_int("SPTRANS_FFT2_F", 6));
    const int nblk = (nf + ft - 1) / ft;
    grp.F2 = (nf + nblk - 1) / std::max(nblk, 1);
    grp.chunk_field.assign(nchunks + 1, nf);
    for (int c = 0; c < nchunks; ++c) grp.chunk_field[c] = static_cast<int>(static_cast<long long>(nf) * c / nchunks);
    for (size_t gi = 0; gi < grp.pairs.size(); ++gi) {
        std::vector<int2> blocks;
        std::vector<int> cb(nchunks + 1, 0);
        for (int c = 0; c < nchunks; ++c) {
            cb[c] = static_cast<int>(blocks.size());
            const int fb = grp.chunk_field[c], fe = grp.chunk_field[c + 1];
            for (int j : grp.pairs[gi]) {
                int step = grp.mode[gi] == 2 ? grp.F2 : meta[j].F;
                if (grp.mode[gi] == 2 && nchunks > 1) {  // equal runs within the chunk
                    const int nb = (fe - fb + ft - 1) / ft;
                    step = (fe - fb + nb - 1) / std::max(nb, 1);
                }
                for (int f0 = fb; f0 < fe; f0 += step) blocks.push_back(make_int2(j, f0 | (std::min(step, fe - f0) << 16)));
            }
        }
        cb[nchunks] = static_cast<int>(blocks.size());
        int2* d = nullptr;
        SPT_CUDA(cudaMalloc(&d, std::max<size_t>(blocks.size(), 1) * sizeof(int2)));
        SPT_CUDA(cudaMemcpyAsync(d, blocks.data(), blocks.size() * sizeof(int2), cudaMemcpyHostToDevice, p.stream));
        SPT_CUDA(cudaStreamSynchronize(p.stream));
        grp.d_blocks.push_back(d);
        grp.nblocks.push_back(static_cast<int>(blocks.size()));
        grp.chunk_begin.push_back(cb);
    }
    grp.nf = nf;
    grp.nchunks = nchunks;
    return SPTRANS_OK;
}

int fourier_set_chunks(Plan& p, int nf, int nchunks, std::vector<int>* field_bounds) {
    int rc = ensure_block_lists(p, nf, nchunks);
    if (rc) return rc;
    if (field_bounds) *field_bounds = fft_state(p).grp.chunk_field;
    return SPTRANS_OK;
}

static DirectArgs make_direct_args(Plan& p, int nf) {
    DirectArgs a{};
    a.meta = reinterpret_cast<const PairMeta*>(p.d_pair_meta);
    a.nf = nf;
    a.fb_rowoff = p.d_fb_rowoff;
    a.nlat0 = p.d_nlat0;
    a.nleg = p.g.nleg;
    a.scheds = reinterpret_cast<const ScheduleG*>(p.d_fft_order);
    a.twid = p.d_twiddle;
    a.chirp = p.d_chirp;
    a.npts = p.g.gp_stride;
    return a;
}
// launch shape of a direct group as threads * 16 + blocks per SM (SPTRANS_FFTD_SHAPE="<threads>x<blocks>" for experiments)
static int direct_shape(size_t smem) {
    if (smem > 113 * 1024) return 512 * 16 + 1;  // one block per SM
    int nt = 256, mb = 2;
    if (smem <= 56 * 1024) nt = 128, mb = 4;  // measured at TCo1279 L137: 11.06 / 11.32 ms per direction against 11.18 / 11.38
    if (const char* e = std::getenv("SPTRANS_FFTD_SHAPE")) std::sscanf(e, "%dx%d", &nt, &mb);
    return nt * 16 + mb;
}

static Fft2Args make_v2_args(Plan& p, const FftGroups& grp, int nf) {
    Fft2Args a{};
    a.meta = reinterpret_cast<const PairMeta*>(p.d_pair_meta);
    a.nf = nf;
    a.F = grp.F2;
    a.fb_rowoff = p.d_fb_rowoff;
    a.nlat0 = p.d_nlat0;
    a.nleg = p.g.nleg;
    a.twid = p.d_twiddle;
    a.t256 = fft_state(p).t256;
    a.chirp = p.d_chirp;
    a.filt = p.d_filt;
    a.npts = p.g.gp_stride;
    return a;
}



static StreamFan* fan_begin(Plan& p) {
    const int ns = std::max(1, std::min(8, env_int("SPTRANS_FFT_STREAMS", 2)));
    StreamFan& f = fft_state(p).fan;
    if (ns > 1 && f.fork && static_cast<int>(f.aux.size()) != ns - 1) fan_free(p);
    if (ns <= 1) return nullptr;
    if (!f.fork) {
        cudaEventCreateWithFlags(&f.fork, cudaEventDisableTiming);
        f.aux.resize(ns - 1);
        f.done.resize(ns - 1);
        for (int i = 0; i < ns - 1; ++i) {
            cudaStreamCreateWithFlags(&f.aux[i], cudaStreamNonBlocking);
            cudaEventCreateWithFlags(&f.done[i], cudaEventDisableTiming);
        }
    }
    cudaEventRecord(f.fork, p.stream);
    for (cudaStream_t a : f.aux) cudaStreamWaitEvent(a, f.fork, 0);
    return &f;
}
static cudaStream_t fan_stream(Plan& p, StreamFan* f, int k) {
    if (!f) return p.stream;
    const int ns = static_cast<int>(f->aux.size()) + 1;
    return (k % ns == 0) ? p.stream : f->aux[k % ns - 1];
}
static void fan_end(Plan& p, StreamFan* f) {
    if (!f) return;
    for (size_t i = 0; i < f->aux.size(); ++i) {
        cudaEventRecord(f->done[i], f->aux[i]);
        cudaStreamWaitEvent(p.stream, f->done[i], 0);
    }
}
static void fan_free(Plan& p) {
    if (!p.fft) return;
    StreamFan& f = fft_state(p).fan;
    for (cudaStream_t a : f.aux) cudaStreamDestroy(a);
    for (cudaEvent_t e : f.done) cudaEventDestroy(e);
    if (f.fork) cudaEventDestroy(f.fork);
    f = StreamFan{};
}

// Block range of launch group gi for `chunk` (-1: every chunk of the current lists, i.e. all fields)
static void chunk_range(const FftGroups& grp, size_t gi, int chunk, int* first, int* count) {
    if (chunk < 0) {
        *first = 0;
        *count = grp.nblocks[gi];
    }
    else {
        *first = grp.chunk_begin[gi][chunk];
        *count = grp.chunk_begin[gi][chunk + 1] - *first;
    }
}
static int lists_for(Plan& p, int nf, int chunk) {
    const FftGroups& grp = fft_state(p).grp;
    if (chunk >= 0) {
        if (grp.nf != nf || chunk >= grp.nchunks) {
            set_error("fourier stage: field chunk requested without matching block lists (fourier_set_chunks)");
            return SPTRANS_ERR_INVALID;
        }
        return SPTRANS_OK;
    }
    return ensure_block_lists(p, nf, grp.nf == nf ? grp.nchunks : 1);
}

int launch_fourier_inv(Plan& p, int nf, int mlimit, const double* d_fourier, double* d_gp, int nb_uv, const double* d_scale,
                       int chunk) {
    if (p.g.points) {
        set_error("this entry point is not available for point-set plans");
        return SPTRANS_ERR_NOT_IMPLEMENTED;
    }
    int rc = lists_for(p, nf, chunk);
    if (rc) return rc;
    const FftGroups& grp = fft_state(p).grp;
    const double* lat_scale = d_scale ? d_scale : p.d_coslatinv;  // per latitude pair, applied to fields < nb_uv
    StreamFan* fan = fan_begin(p);
    int launched = 0;
    auto launch_groups = [&]() -> int {
        for (size_t gi = 0; gi < grp.pairs.size(); ++gi) {
            int first = 0, nblk = 0;
            chunk_range(grp, gi, chunk, &first, &nblk);
            if (nblk == 0) continue;
            const int2* blk = grp.d_blocks[gi] + first;
            const cudaStream_t st = fan_stream(p, fan, launched++);
            if (grp.mode[gi] == 2) {
                Fft2Args a = make_v2_args(p, grp, nf);
                a.mlimit = mlimit;
                a.nb_uv = nb_uv;
                a.scale_lat = lat_scale;
                a.fb = const_cast<double2*>(reinterpret_cast<const double2*>(d_fourier));
                a.gp = d_gp;
                switch (grp.m1[gi]) {
#define SPT_LAUNCH(R)                                                                          \
    case R:                                                                                    \
        fourier2_inv_kernel<R><<<nblk, v2_threads(R), grp.smem[gi], st>>>(a, blk);             \
        break;
                    SPT_V2_RADICES(SPT_LAUNCH)
#undef SPT_LAUNCH
                    default: set_error("fourier: unsupported v2 radix"); return SPTRANS_ERR_INVALID;
                }
                p.launches++;
                SPT_CUDA(cudaGetLastError());
                continue;
            }
            const int threads = grp.smem[gi] > 113 * 1024 ? 2 * kFftThreads : kFftThreads;  // one block per SM: 16 warps
            if (grp.mode[gi] == 3) {
                DirectArgs a = make_direct_args(p, nf);
                a.mlimit = mlimit;
                a.nb_uv = nb_uv;
                a.scale_lat = lat_scale;
                a.fb = const_cast<double2*>(reinterpret_cast<const double2*>(d_fourier));
                a.gp = d_gp;
                const int shape = direct_shape(grp.smem[gi]);
                switch (shape) {
#define SPT_LAUNCH_D(NT, MB)                                                                    \
    case NT * 16 + MB:                                                                          \
        fourier_inv_direct_kernel<NT, MB><<<nblk, NT, grp.smem[gi], st>>>(a, blk);              \
        break;
                    SPT_DIRECT_SHAPES(SPT_LAUNCH_D)
#undef SPT_LAUNCH_D
                    default: set_error("fourier: unsupported shape of the direct kernels"); return SPTRANS_ERR_INVALID;
                }
                p.launches++;
                SPT_CUDA(cudaGetLastError());
                continue;
            }
            if (grp.mode[gi]) {
                fourier_inv_rows_kernel<<<nblk, threads, grp.smem[gi], st>>>(
                    reinterpret_cast<const PairMeta*>(p.d_pair_meta), blk, nf, mlimit, nb_uv,
                    reinterpret_cast<const double2*>(d_fourier), p.d_fb_rowoff, p.d_nlat0, p.g.nleg,
                    reinterpret_cast<const ScheduleG*>(p.d_fft_order), p.d_twiddle, p.d_chirp, p.d_filt, lat_scale, d_gp, p.g.gp_stride);
                p.launches++;
                SPT_CUDA(cudaGetLastError());
                continue;
            }
            fourier_inv_kernel<<<nblk, threads, grp.smem[gi], st>>>(
                reinterpret_cast<const PairMeta*>(p.d_pair_meta), blk, nf, mlimit, nb_uv,
                reinterpret_cast<const double2*>(d_fourier), p.d_fb_rowoff, p.d_nlat0, p.g.nleg,
                reinterpret_cast<const ScheduleG*>(p.d_fft_order), p.d_twiddle, p.d_chirp, p.d_filt, lat_scale, d_gp, p.g.gp_stride);
            p.launches++;
            SPT_CUDA(cudaGetLastError());
        }
        return SPTRANS_OK;
    };
    rc = launch_groups();
    fan_end(p, fan);  // the auxiliary streams are joined on every exit path
    return rc;
}

static int launch_fourier_dir_impl(Plan& p, int nf, const double* d_gp, double* d_fourier, int nb_uv, int adjoint,
                                   const int* d_owner, const PeerDst& dst, int chunk) {
    if (p.g.points) {  // like TransLocal: no direct / adjoint transform from scattered points
        set_error("direct and adjoint transforms are not available for point-set plans");
        return SPTRANS_ERR_NOT_IMPLEMENTED;
    }
    if (!p.d_weights && !adjoint) {
        set_error("dirtrans: plan was created without quadrature weights");
        return SPTRANS_ERR_INVALID;
    }
    // (the fused push of a sharded plan counts the blocks of a latitude pair: unchunked lists only)
    int rc = d_owner ? ensure_block_lists(p, nf, 1) : lists_for(p, nf, chunk);
    if (rc) return rc;
    const FftGroups& grp = fft_state(p).grp;
    // wind fields enter the vor/div transform as u,v / (a cos(lat)); the adjoint of the inverse wind transform
    // applies the inverse's own 1 / cos(lat)
    const double* uv_scale = adjoint ? p.d_coslatinv : p.d_uvscale;
    if (d_owner) {  // empty work queue of completed pairs (finish_pair_and_push); the launch groups of all streams share it
        SPT_CUDA(cudaMemsetAsync(p.d_pair_done + p.g.nleg, 0, (2 * static_cast<size_t>(p.g.nleg) + 8) * sizeof(int), p.stream));
        if (std::getenv("SPTRANS_PUSH_ROWS")) {  // measurement knob
            const int v = env_int("SPTRANS_PUSH_ROWS", 128);
            SPT_CUDA(cudaMemcpyToSymbolAsync(d_push_rows, &v, sizeof(int), 0, cudaMemcpyHostToDevice, p.stream));
        }
    }
    StreamFan* fan = fan_begin(p);
    int launched = 0;
    auto launch_groups = [&]() -> int {
        for (size_t gi = 0; gi < grp.pairs.size(); ++gi) {
            int first = 0, nblk = 0;
            chunk_range(grp, gi, d_owner ? -1 : chunk, &first, &nblk);
            if (nblk == 0) continue;
            const int2* blk = grp.d_blocks[gi] + first;
            const cudaStream_t st = fan_stream(p, fan, launched++);
            if (grp.mode[gi] == 2) {
                Fft2Args a = make_v2_args(p, grp, nf);
                a.nb_uv = nb_uv;
                a.scale_lat = uv_scale;
                a.weights = p.d_weights;
                a.fb = reinterpret_cast<double2*>(d_fourier);
                a.gp = const_cast<double*>(d_gp);
                a.adjoint = adjoint;
                a.gp_aligned16 = (reinterpret_cast<uintptr_t>(d_gp) % 16 == 0 && p.g.gp_stride % 2 == 0) ? 1 : 0;
                switch (grp.m1[gi]) {
#define SPT_LAUNCH(R)                                                                                             \
    case R:                                                                                                       \
        fourier2_dir_kernel<R><<<nblk, v2_threads(R), grp.smem_dir[gi], st>>>(a, blk, d_owner, dst, p.g.rank,     \
                                                                             p.d_pair_done);                      \
        break;
                    SPT_V2_RADICES(SPT_LAUNCH)
#undef SPT_LAUNCH
                    default: set_error("fourier: unsupported v2 radix"); return SPTRANS_ERR_INVALID;
                }
                p.launches++;
                SPT_CUDA(cudaGetLastError());
                continue;
            }
            const int threads = grp.smem[gi] > 113 * 1024 ? 2 * kFftThreads : kFftThreads;
            if (grp.mode[gi] == 3) {
                DirectArgs a = make_direct_args(p, nf);
                a.nb_uv = nb_uv;
                a.scale_lat = uv_scale;
                a.weights = p.d_weights;
                a.fb = reinterpret_cast<double2*>(d_fourier);
                a.gp = const_cast<double*>(d_gp);
                a.adjoint = adjoint;
                const int shape = direct_shape(grp.smem[gi]);
                switch (shape) {
#define SPT_LAUNCH_D(NT, MB)                                                                                            \
    case NT * 16 + MB:                                                                                                  \
        fourier_dir_direct_kernel<NT, MB><<<nblk, NT, grp.smem[gi], st>>>(a, blk, d_owner, dst, p.g.rank, p.d_pair_done); \
        break;
                    SPT_DIRECT_SHAPES(SPT_LAUNCH_D)
#undef SPT_LAUNCH_D
                    default: set_error("fourier: unsupported shape of the direct kernels"); return SPTRANS_ERR_INVALID;
                }
                p.launches++;
                SPT_CUDA(cudaGetLastError());
                continue;
            }
            if (grp.mode[gi]) {
                fourier_dir_rows_kernel<<<nblk, threads, grp.smem[gi], st>>>(
                    reinterpret_cast<const PairMeta*>(p.d_pair_meta), blk, nf, nb_uv, d_gp, p.g.gp_stride, p.d_fb_rowoff, p.d_nlat0,
                    p.g.nleg, reinterpret_cast<const ScheduleG*>(p.d_fft_order), p.d_twiddle, p.d_chirp, p.d_filt, p.d_weights,
                    uv_scale, reinterpret_cast<double2*>(d_fourier), adjoint);
                p.launches++;
                SPT_CUDA(cudaGetLastError());
                continue;
            }
            fourier_dir_kernel<<<nblk, threads, grp.smem[gi], st>>>(
                reinterpret_cast<const PairMeta*>(p.d_pair_meta), blk, nf, nb_uv, d_gp, p.g.gp_stride, p.d_fb_rowoff, p.d_nlat0,
                p.g.nleg, reinterpret_cast<const ScheduleG*>(p.d_fft_order), p.d_twiddle, p.d_chirp, p.d_filt, p.d_weights, uv_scale,
                reinterpret_cast<double2*>(d_fourier), adjoint, d_owner, dst, p.g.rank, p.d_pair_done);
            p.launches++;
            SPT_CUDA(cudaGetLastError());
        }
        return SPTRANS_OK;
    };
    rc = launch_groups();
    fan_end(p, fan);
    return rc;
}

int launch_fourier_dir(Plan& p, int nf, const double* d_gp, double* d_fourier, int nb_uv, int adjoint, int chunk) {
    return launch_fourier_dir_impl(p, nf, d_gp, d_fourier, nb_uv, adjoint, nullptr, PeerDst{}, chunk);
}

int launch_fourier_dir_peers(Plan& p, int nf, const double* d_gp, const PeerDst& dst, bool* fused) {
    int rc = ensure_block_lists(p, nf, 1);
    if (rc) return rc;
    const FftGroups& grp = fft_state(p).grp;
    bool rows = false;
    for (size_t gi = 0; gi < grp.pairs.size(); ++gi) rows = rows || (grp.mode[gi] == 1 && grp.nblocks[gi] > 0);
    *fused = !rows;
    if (rows) return launch_fourier_dir_impl(p, nf, d_gp, dst.base[p.g.rank], 0, 0, nullptr, dst, -1);
    return launch_fourier_dir_impl(p, nf, d_gp, dst.base[p.g.rank], 0, 0, p.d_owner, dst, -1);
}

}  // namespace sptrans
