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
#include <map>
#include <vector>

#include "fft_core.cuh"
#include "plan.hpp"

namespace sptrans {

struct PairMeta {
    long long chirp_off;  // A_u (2L+1 entries) then C_i (n entries), double2 units
    long long filt_off;   // Bhat, M entries, digit-reversed order, scaled by 1/M
    long long rowN;       // grid offset of the northern row
    long long rowS;       // grid offset of the southern row
    int n;                // row length
    int L;                // zonal truncation at this latitude (mmax[j]); -1: nothing resolved
    int logM;
    int has_s;            // 0 for the equator row of a grid with an odd number of latitudes
};

namespace {

using namespace fftc;

constexpr int kWn = 8192;  // master twiddle table length (largest supported M)

__global__ void twiddle_kernel(double2* W, int Wn) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < Wn) {
        double s, c;
        sincospi(-2.0 * k / Wn, &s, &c);
        W[k] = make_double2(c, s);
    }
}

// one block per distinct (n, L): chirps and the digit-reversed filter spectrum
__global__ void __launch_bounds__(256)
chirp_tables_kernel(const PairMeta* __restrict__ cls, const double2* __restrict__ W, int Wn,
                    double2* __restrict__ chirp, double2* __restrict__ filt) {
    extern __shared__ double2 X[];
    const PairMeta pm = cls[blockIdx.x];
    const int n = pm.n, L = pm.L, M = 1 << pm.logM;
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
    for (int e = tid; e < padded_len(M); e += nthr) X[e] = make_double2(0., 0.);
    __syncthreads();
    for (int e = tid; e < n + 2 * L; e += nthr) {
        const int k = e - 2 * L;  // k in [-2L, n-1]
        double s, c;
        sincospi(-static_cast<double>(chirp_residue(k, 0, n)) / n, &s, &c);
        const int idx = (k % M + M) % M;
        X[pad(idx)] = make_double2(c, s);
    }
    __syncthreads();
    fft_dif_all(X, 1, pm.logM, W, Wn, tid, nthr);
    const double sc = 1.0 / M;
    for (int k = tid; k < M; k += nthr) {
        const double2 v = X[pad(k)];
        filt[pm.filt_off + k] = make_double2(v.x * sc, v.y * sc);
    }
}

__global__ void __launch_bounds__(512)
fourier_inv_kernel(const PairMeta* __restrict__ meta, const int* __restrict__ group_pairs, int ngf, int F, int nf,
                   int mlimit, int nb_uv, const double2* __restrict__ fb, const long long* __restrict__ fb_rowoff,
                   const int* __restrict__ nlat0, int nleg, const double2* __restrict__ W,
                   const double2* __restrict__ chirp, const double2* __restrict__ filt,
                   const double* __restrict__ coslatinv, double* __restrict__ gp, long long npts) {
    extern __shared__ double2 X[];
    const int pair = group_pairs[blockIdx.x / ngf];
    const int f0 = (blockIdx.x % ngf) * F;
    const int nfb = min(F, nf - f0);
    const PairMeta pm = meta[pair];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n = pm.n, L = pm.L;
    const int M = 1 << pm.logM, PL = padded_len(M);
    const int Lc = min(L, mlimit);
    if (Lc < 0) {  // no zonal wavenumber resolved / requested at this latitude: rows are zero
        for (int w = tid; w < nfb * n; w += nthr) {
            const int fi = w / n, i = w - fi * n;
            gp[(f0 + fi) * npts + pm.rowN + i] = 0.;
            if (pm.has_s) gp[(f0 + fi) * npts + pm.rowS + i] = 0.;
        }
        return;
    }
    for (int e = tid; e < nfb * PL; e += nthr) X[e] = make_double2(0., 0.);
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
        X[fi * PL + pad(L + m)] = cmul(Zp, A[L + m]);
        if (m > 0) {
            const double2 Zm = make_double2(FN.x + FS.y, FS.x - FN.y);
            X[fi * PL + pad(L - m)] = cmul(Zm, A[L - m]);
        }
    }
    __syncthreads();
    fft_dif_all(X, nfb, pm.logM, W, kWn, tid, nthr);
    fft_dit_all<false>(X, nfb, pm.logM, W, kWn, filt + pm.filt_off, tid, nthr);
    for (int w = tid; w < nfb * n; w += nthr) {
        const int fi = w / n, i = w - fi * n;
        const double2 z = cmul(X[fi * PL + pad(i)], C[i]);
        const int f = f0 + fi;
        double sn = 1., ss = 1.;
        if (f < nb_uv) {  // u,v = U,V / cos(lat)  (reference :1443-1469)
            sn = ss = coslatinv[pair];  // grid is symmetric about the equator
        }
        gp[f * npts + pm.rowN + i] = z.x * sn;
        if (pm.has_s) gp[f * npts + pm.rowS + i] = z.y * ss;
    }
}

__global__ void __launch_bounds__(512)
fourier_dir_kernel(const PairMeta* __restrict__ meta, const int* __restrict__ group_pairs, int ngf, int F, int nf,
                   int nb_uv, const double* __restrict__ gp, long long npts, const long long* __restrict__ fb_rowoff,
                   const int* __restrict__ nlat0, int nleg, const double2* __restrict__ W,
                   const double2* __restrict__ chirp, const double2* __restrict__ filt,
                   const double* __restrict__ weights, const double* __restrict__ coslat, double2* __restrict__ fb) {
    extern __shared__ double2 X[];
    const int pair = group_pairs[blockIdx.x / ngf];
    const int f0 = (blockIdx.x % ngf) * F;
    const int nfb = min(F, nf - f0);
    const PairMeta pm = meta[pair];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int n = pm.n, L = pm.L;
    if (L < 0) return;
    const int M = 1 << pm.logM, PL = padded_len(M);
    for (int e = tid; e < nfb * PL; e += nthr) X[e] = make_double2(0., 0.);
    __syncthreads();
    const double2* A = chirp + pm.chirp_off;
    const double2* C = A + (2 * L + 1);
    for (int w = tid; w < nfb * n; w += nthr) {
        const int fi = w / n, i = w - fi * n;
        const int f = f0 + fi;
        double xn = gp[f * npts + pm.rowN + i];
        double xs = pm.has_s ? gp[f * npts + pm.rowS + i] : 0.;
        if (f < nb_uv) {  // wind components enter the transform as U,V = u,v * cos(lat)
            xn *= coslat[pair];
            xs *= coslat[pair];
        }
        X[fi * PL + pad(i)] = cmulc(make_double2(xn, xs), C[i]);
    }
    __syncthreads();
    fft_dif_all(X, nfb, pm.logM, W, kWn, tid, nthr);
    fft_dit_all<true>(X, nfb, pm.logM, W, kWn, filt + pm.filt_off, tid, nthr);
    const double wq = weights[pair];
    const double inv_n = 1.0 / n;
    for (int w = tid; w < nfb * (L + 1); w += nthr) {
        const int m = w / nfb, fi = w - m * nfb;
        double2 Gp = cmulc(X[fi * PL + pad(L + m)], A[L + m]);
        double2 Gm = cmulc(X[fi * PL + pad(L - m)], A[L - m]);
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
}

int fields_per_block(int logM) {
    // aim at ~64 KB of shared memory per block (3 blocks / SM) but never less than one sequence
    const int M = 1 << logM;
    int F = 4096 / M;
    if (F < 1) F = 1;
    if (F > 16) F = 16;
    return F;
}

}  // namespace

struct FftGroups {
    std::vector<int> logM;                 // distinct logM, descending
    std::vector<std::vector<int>> pairs;   // pairs per group, longest rows first
    std::vector<int*> d_pairs;
};
static std::map<Plan*, FftGroups> g_groups;

int build_fft_tables(Plan& p) {
    HostGeom& g = p.g;
    const int nleg = g.nleg;
    std::vector<PairMeta> meta(nleg);
    std::map<std::pair<int, int>, int> cls_index;
    std::vector<PairMeta> classes;
    long long chirp_total = 0, filt_total = 0;
    for (int j = 0; j < nleg; ++j) {
        PairMeta pm{};
        pm.n = g.nx[j];
        pm.L = g.mmax[j];
        pm.rowN = g.rowoff[j];
        pm.rowS = g.rowoff[g.nlat - 1 - j];
        pm.has_s = (g.nlat - 1 - j != j) ? 1 : 0;
        if (pm.L >= 0 && 2 * pm.L >= pm.n) {
            set_error("sptrans_plan_create: zonal truncation at a latitude row reaches nx/2 (aliasing); unsupported");
            return SPTRANS_ERR_INVALID;
        }
        const int Luse = std::max(pm.L, 0);
        int logM = 0;
        const int M = fftc::conv_length(pm.n, Luse, &logM);
        if (M > kWn) {
            set_error("sptrans_plan_create: row length + 2*truncation exceeds 8192 (grids beyond O1280 need the "
                      "two-level Fourier kernel, not available yet)");
            return SPTRANS_ERR_NOT_IMPLEMENTED;
        }
        pm.logM = logM;
        auto key = std::make_pair(pm.n, Luse);
        auto it = cls_index.find(key);
        if (it == cls_index.end()) {
            PairMeta c = pm;
            c.L = Luse;
            c.chirp_off = chirp_total;
            c.filt_off = filt_total;
            chirp_total += 2LL * Luse + 1 + pm.n;
            filt_total += M;
            cls_index[key] = static_cast<int>(classes.size());
            classes.push_back(c);
            it = cls_index.find(key);
        }
        pm.chirp_off = classes[it->second].chirp_off;
        pm.filt_off = classes[it->second].filt_off;
        meta[j] = pm;
    }
    double2* d_W = nullptr;
    SPT_CUDA(cudaMalloc(&d_W, kWn * sizeof(double2)));
    twiddle_kernel<<<(kWn + 255) / 256, 256, 0, p.stream>>>(d_W, kWn);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    SPT_CUDA(cudaMalloc(&p.d_chirp, std::max<long long>(chirp_total, 1) * sizeof(double2)));
    SPT_CUDA(cudaMalloc(&p.d_filt, std::max<long long>(filt_total, 1) * sizeof(double2)));
    p.bytes_tables += (chirp_total + filt_total + kWn) * sizeof(double2);
    PairMeta* d_cls = nullptr;
    SPT_CUDA(cudaMalloc(&d_cls, classes.size() * sizeof(PairMeta)));
    SPT_CUDA(cudaMemcpyAsync(d_cls, classes.data(), classes.size() * sizeof(PairMeta), cudaMemcpyHostToDevice, p.stream));
    const size_t smem_max = static_cast<size_t>(fftc::padded_len(kWn)) * sizeof(double2);
    SPT_CUDA(cudaFuncSetAttribute(chirp_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    SPT_CUDA(cudaFuncSetAttribute(fourier_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    SPT_CUDA(cudaFuncSetAttribute(fourier_dir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    chirp_tables_kernel<<<static_cast<int>(classes.size()), 256, smem_max, p.stream>>>(d_cls, d_W, kWn, p.d_chirp,
                                                                                      p.d_filt);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    SPT_CUDA(cudaMalloc(&p.d_pair_meta, meta.size() * sizeof(PairMeta)));
    SPT_CUDA(cudaMemcpyAsync(p.d_pair_meta, meta.data(), meta.size() * sizeof(PairMeta), cudaMemcpyHostToDevice,
                             p.stream));
    p.d_twiddle = d_W;
    // launch groups by logM over this rank's latitude band
    FftGroups grp;
    std::map<int, std::vector<int>, std::greater<int>> by;
    for (int j = g.pair_begin; j < g.pair_end; ++j) by[meta[j].logM].push_back(j);
    for (auto& kv : by) {
        std::vector<int> v = kv.second;
        std::stable_sort(v.begin(), v.end(), [&](int a, int b) { return meta[a].n > meta[b].n; });
        int* d = nullptr;
        SPT_CUDA(cudaMalloc(&d, v.size() * sizeof(int)));
        SPT_CUDA(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, p.stream));
        grp.logM.push_back(kv.first);
        grp.pairs.push_back(v);
        grp.d_pairs.push_back(d);
    }
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    cudaFree(d_cls);
    g_groups[&p] = grp;
    return SPTRANS_OK;
}

void free_fft_tables(Plan& p) {
    auto it = g_groups.find(&p);
    if (it != g_groups.end()) {
        for (int* d : it->second.d_pairs) cudaFree(d);
        g_groups.erase(it);
    }
}

int launch_fourier_inv(Plan& p, int nf, int mlimit, const double* d_fourier, double* d_gp, int nb_uv) {
    const FftGroups& grp = g_groups[&p];
    for (size_t gi = 0; gi < grp.logM.size(); ++gi) {
        const int logM = grp.logM[gi];
        const int F = fields_per_block(logM);
        const int ngf = (nf + F - 1) / F;
        const size_t smem = static_cast<size_t>(F) * fftc::padded_len(1 << logM) * sizeof(double2);
        const int threads = (logM >= 13) ? 512 : 256;
        const long long blocks = static_cast<long long>(grp.pairs[gi].size()) * ngf;
        fourier_inv_kernel<<<static_cast<unsigned>(blocks), threads, smem, p.stream>>>(
            reinterpret_cast<const PairMeta*>(p.d_pair_meta), grp.d_pairs[gi], ngf, F, nf, mlimit, nb_uv,
            reinterpret_cast<const double2*>(d_fourier), p.d_fb_rowoff, p.d_nlat0, p.g.nleg, p.d_twiddle, p.d_chirp,
            p.d_filt, p.d_coslatinv, d_gp, p.g.npts);
        p.launches++;
        SPT_CUDA(cudaGetLastError());
    }
    return SPTRANS_OK;
}

int launch_fourier_dir(Plan& p, int nf, const double* d_gp, double* d_fourier, int nb_uv) {
    if (!p.d_weights) {
        set_error("dirtrans: plan was created without quadrature weights");
        return SPTRANS_ERR_INVALID;
    }
    const FftGroups& grp = g_groups[&p];
    for (size_t gi = 0; gi < grp.logM.size(); ++gi) {
        const int logM = grp.logM[gi];
        const int F = fields_per_block(logM);
        const int ngf = (nf + F - 1) / F;
        const size_t smem = static_cast<size_t>(F) * fftc::padded_len(1 << logM) * sizeof(double2);
        const int threads = (logM >= 13) ? 512 : 256;
        const long long blocks = static_cast<long long>(grp.pairs[gi].size()) * ngf;
        fourier_dir_kernel<<<static_cast<unsigned>(blocks), threads, smem, p.stream>>>(
            reinterpret_cast<const PairMeta*>(p.d_pair_meta), grp.d_pairs[gi], ngf, F, nf, nb_uv, d_gp, p.g.npts,
            p.d_fb_rowoff, p.d_nlat0, p.g.nleg, p.d_twiddle, p.d_chirp, p.d_filt, p.d_weights, p.d_coslat,
            reinterpret_cast<double2*>(d_fourier));
        p.launches++;
        SPT_CUDA(cudaGetLastError());
    }
    return SPTRANS_OK;
}

}  // namespace sptrans
