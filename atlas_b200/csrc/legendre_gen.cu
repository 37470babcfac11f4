// Device generation of the normalised associated Legendre functions Pbar_n^m(mu_j) for all
// northern latitude rows, written straight into the GEMM-ready table layout
//     tab[(m,parity)][k][jj]   n = m + parity + 2k (ascending),  jj = j - nlat0[m]  (latitude fastest)
//
// Replaces compute_legendre_polynomials (ecmwf/atlas src/atlas/trans/local/LegendrePolynomials.cc:154-209,
// minutes of serial host time at T1279) with one kernel.  The four-point Belousov recurrence (:136-149)
// couples (m,n) only to rows n-1 and n-2, so all m advance together row by row; the m=0,1 columns and the
// diagonal are seeded from the host (host_setup.cc).  Every operation is an explicitly rounded IEEE
// double op in the reference's order (no fma contraction), so the table is BIT-IDENTICAL to the
// reference's output -- tests/test_gpu_legendre.py checks this against oracle/_ref.
#include <cstdio>
#include <vector>

#include "plan.hpp"

namespace sptrans {

namespace {

constexpr int kGenThreads = 256;

// kLatsPerBlock latitudes advance together (32-byte store runs); 4 normally, 2 when the three row buffers of a
// very high truncation (T > ~2400) would not fit in shared memory
template <bool kCacheLayout, int kLatsPerBlock>
__global__ void __launch_bounds__(kGenThreads)
legendre_gen_kernel(int trc,            // table truncation (T+1)
                    int Tm,             // highest m stored (T)
                    int nleg,           // northern rows
                    const double* __restrict__ xcos,   // [nleg]
                    const double* __restrict__ col0,   // [nleg][trc+1]
                    const double* __restrict__ col1,   // [nleg][trc+1]
                    const double* __restrict__ diag,   // [nleg][trc+1]
                    const int* __restrict__ nlat0,     // [T+2]
                    const long long* __restrict__ tab_off,  // [2(T+1)]
                    const int* __restrict__ tab_pitch,      // [T+1]
                    double* __restrict__ tab) {
    extern __shared__ double sm[];
    const int W = trc + 1;
    // three rotating rows of [W][kLatsPerBlock]
    double* rows[3] = {sm, sm + W * kLatsPerBlock, sm + 2 * W * kLatsPerBlock};
    const int j0 = blockIdx.x * kLatsPerBlock;
    double xl[kLatsPerBlock];
#pragma unroll
    for (int l = 0; l < kLatsPerBlock; ++l) xl[l] = (j0 + l < nleg) ? xcos[j0 + l] : 0.;

    auto store = [&](int m, int n, int l, double v) {
        if (m > Tm) return;
        const int j = j0 + l;
        if (j >= nleg) return;
        const int par = (n - m) & 1;
        const int k = (n - m) >> 1;
        if constexpr (kCacheLayout) {
            // reference cache layout (TransLocal.cc:592-647, LegendrePolynomials.cc:181-205): per (m,parity)
            // block element [kdesc + K*j], kdesc = 0 <-> highest n; tab_off = block begin, tab_pitch = K
            const int K = tab_pitch[2 * m + par];
            tab[tab_off[2 * m + par] + static_cast<long long>(K) * j + (K - 1 - k)] = v;
        }
        else {
            const int jj = j - nlat0[m];
            const long long off = tab_off[2 * m + par];
            if (jj < 0 || off < 0) return;  // latitude pruned for this m / block owned by another rank
            tab[off + static_cast<long long>(k) * tab_pitch[m] + jj] = v;
        }
    };

    // rows n = 0, 1, 2 come entirely from the seeds
    for (int n = 0; n <= trc; ++n) {
        double* cur = rows[n % 3];
        const double* p1 = rows[(n + 2) % 3];  // row n-1
        const double* p2 = rows[(n + 1) % 3];  // row n-2
        for (int idx = threadIdx.x; idx < (n + 1) * kLatsPerBlock; idx += kGenThreads) {
            const int m = idx / kLatsPerBlock;
            const int l = idx % kLatsPerBlock;
            const int j = j0 + l;
            double v = 0.;
            if (j < nleg) {
                if (m == 0) v = col0[static_cast<size_t>(j) * W + n];
                else if (m == 1) v = col1[static_cast<size_t>(j) * W + n];
                else if (m == n) v = diag[static_cast<size_t>(j) * W + n];
                else {
                    // Belousov eq. (17); operand order and roundings as in the reference :138-148
                    const double dn_ = n, dm_ = m;
                    const double cn = __dmul_rn(__dmul_rn(__dadd_rn(__dmul_rn(2., dn_), 1.), __dsub_rn(__dadd_rn(dn_, dm_), 3.)),
                                                __dsub_rn(__dadd_rn(dn_, dm_), 1.));
                    const double cd = __dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(2., dn_), 3.), __dsub_rn(__dadd_rn(dn_, dm_), 2.)),
                                                __dadd_rn(dn_, dm_));
                    const double dn = __dmul_rn(__dmul_rn(__dadd_rn(__dmul_rn(2., dn_), 1.), __dadd_rn(__dsub_rn(dn_, dm_), 1.)),
                                                __dsub_rn(__dadd_rn(dn_, dm_), 1.));
                    const double dd = __dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(2., dn_), 1.), __dsub_rn(__dadd_rn(dn_, dm_), 2.)),
                                                __dadd_rn(dn_, dm_));
                    const double en = __dmul_rn(__dadd_rn(__dmul_rn(2., dn_), 1.), __dsub_rn(dn_, dm_));
                    const double ed = __dmul_rn(__dsub_rn(__dmul_rn(2., dn_), 1.), __dadd_rn(dn_, dm_));
                    const double t1 = __dmul_rn(__dsqrt_rn(__ddiv_rn(cn, cd)), p2[(m - 2) * kLatsPerBlock + l]);
                    const double t2 = __dmul_rn(__dmul_rn(__dsqrt_rn(__ddiv_rn(dn, dd)), p1[(m - 2) * kLatsPerBlock + l]), xl[l]);
                    const double t3 = __dmul_rn(__dmul_rn(__dsqrt_rn(__ddiv_rn(en, ed)), p1[m * kLatsPerBlock + l]), xl[l]);
                    v = __dadd_rn(__dsub_rn(t1, t2), t3);
                }
            }
            cur[idx] = v;
            store(m, n, l, v);
        }
        __syncthreads();
    }
}

}  // namespace

int generate_legendre_table(Plan& p) {
    HostGeom& g = p.g;
    const int trc = g.T + 1;
    const int W = trc + 1;
    std::vector<double> lats(g.nleg), xcos, col0, col1, diag;
    for (int j = 0; j < g.nleg; ++j) {
        double lat = g.lat_deg[j];
        const double pole = g.points ? 90. : 89.9999999;  // latPole, trans/local/TransLocal.cc:49,:533-546 (grids only)
        if (lat > pole) lat = pole;
        if (lat < -pole) lat = -pole;
        lats[j] = lat * (M_PI / 180.);
    }
    legendre_seeds(trc, g.nleg, lats.data(), xcos, col0, col1, diag);

    const size_t tab_bytes = static_cast<size_t>(g.tab_size) * sizeof(double);
    SPT_CUDA(cudaMalloc(&p.d_tab, std::max<size_t>(tab_bytes, 16)));
    SPT_CUDA(cudaMemsetAsync(p.d_tab, 0, tab_bytes, p.stream));
    p.bytes_tables += tab_bytes;

    double *d_x = nullptr, *d_c0 = nullptr, *d_c1 = nullptr, *d_dg = nullptr;
    long long* d_off = nullptr;
    int* d_pitch = nullptr;
    const size_t colb = static_cast<size_t>(g.nleg) * W * sizeof(double);
    SPT_CUDA(cudaMalloc(&d_x, g.nleg * sizeof(double)));
    SPT_CUDA(cudaMalloc(&d_c0, colb));
    SPT_CUDA(cudaMalloc(&d_c1, colb));
    SPT_CUDA(cudaMalloc(&d_dg, colb));
    SPT_CUDA(cudaMalloc(&d_off, g.tab_off.size() * sizeof(long long)));
    SPT_CUDA(cudaMalloc(&d_pitch, g.tab_pitch.size() * sizeof(int)));
    SPT_CUDA(cudaMemcpyAsync(d_x, xcos.data(), g.nleg * sizeof(double), cudaMemcpyHostToDevice, p.stream));
    SPT_CUDA(cudaMemcpyAsync(d_c0, col0.data(), colb, cudaMemcpyHostToDevice, p.stream));
    SPT_CUDA(cudaMemcpyAsync(d_c1, col1.data(), colb, cudaMemcpyHostToDevice, p.stream));
    SPT_CUDA(cudaMemcpyAsync(d_dg, diag.data(), colb, cudaMemcpyHostToDevice, p.stream));
    SPT_CUDA(cudaMemcpyAsync(d_off, g.tab_off.data(), g.tab_off.size() * sizeof(long long), cudaMemcpyHostToDevice,
                             p.stream));
    SPT_CUDA(cudaMemcpyAsync(d_pitch, g.tab_pitch.data(), g.tab_pitch.size() * sizeof(int), cudaMemcpyHostToDevice,
                             p.stream));

    if (3ull * W * 4 * sizeof(double) <= 200 * 1024) {
        const size_t smem = 3ull * W * 4 * sizeof(double);
        SPT_CUDA(cudaFuncSetAttribute(legendre_gen_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        legendre_gen_kernel<false, 4><<<(g.nleg + 3) / 4, kGenThreads, smem, p.stream>>>(trc, g.T, g.nleg, d_x, d_c0, d_c1, d_dg,
                                                                                       p.d_nlat0, d_off, d_pitch, p.d_tab);
    }
    else {
        const size_t smem = 3ull * W * 2 * sizeof(double);
        if (smem > 226 * 1024) {
            set_error("sptrans_plan_create: truncation too high for the device Legendre generator");
            return SPTRANS_ERR_NOT_IMPLEMENTED;
        }
        SPT_CUDA(cudaFuncSetAttribute(legendre_gen_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        legendre_gen_kernel<false, 2><<<(g.nleg + 1) / 2, kGenThreads, smem, p.stream>>>(trc, g.T, g.nleg, d_x, d_c0, d_c1, d_dg,
                                                                                       p.d_nlat0, d_off, d_pitch, p.d_tab);
    }
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    cudaFree(d_x);
    cudaFree(d_c0);
    cudaFree(d_c1);
    cudaFree(d_dg);
    cudaFree(d_off);
    cudaFree(d_pitch);
    return SPTRANS_OK;
}

int export_legendre_cache(const Plan& p, double* h_out) {
    const HostGeom& g = p.g;
    const int T = g.T;
    const int trc = T + 1;
    const int W = trc + 1;
    // reference offsets (TransLocal.cc:592-606): blocks for m = 0..T+1 padded to 8 doubles, all nleg latitudes
    std::vector<long long> begin(2 * (T + 2), 0);
    std::vector<int> Ks(2 * (T + 2), 0);
    long long size_sym = 0, size_asym = 0;
    auto pad8 = [](long long n) { return (n + 7) / 8 * 8; };
    for (int m = 0; m <= T + 1; ++m) {
        Ks[2 * m] = num_n(trc, m, 0);
        Ks[2 * m + 1] = num_n(trc, m, 1);
        begin[2 * m] = size_sym;
        begin[2 * m + 1] = size_asym;  // relative to the asym base, fixed up below
        size_sym += pad8(static_cast<long long>(Ks[2 * m]) * g.nleg);
        size_asym += pad8(static_cast<long long>(Ks[2 * m + 1]) * g.nleg);
    }
    for (int m = 0; m <= T + 1; ++m) begin[2 * m + 1] += size_sym;
    const long long total = size_sym + size_asym;

    std::vector<double> lats(g.nleg), xcos, col0, col1, diag;
    for (int j = 0; j < g.nleg; ++j) {
        double lat = g.lat_deg[j];
        const double pole = g.points ? 90. : 89.9999999;
        if (lat > pole) lat = pole;
        if (lat < -pole) lat = -pole;
        lats[j] = lat * (M_PI / 180.);
    }
    legendre_seeds(trc, g.nleg, lats.data(), xcos, col0, col1, diag);

    double *d_out = nullptr, *d_x = nullptr, *d_c0 = nullptr, *d_c1 = nullptr, *d_dg = nullptr;
    long long* d_begin = nullptr;
    int* d_K = nullptr;
    const size_t colb = static_cast<size_t>(g.nleg) * W * sizeof(double);
    SPT_CUDA(cudaMalloc(&d_out, std::max<long long>(total, 1) * sizeof(double)));
    SPT_CUDA(cudaMemset(d_out, 0, total * sizeof(double)));
    SPT_CUDA(cudaMalloc(&d_x, g.nleg * sizeof(double)));
    SPT_CUDA(cudaMalloc(&d_c0, colb));
    SPT_CUDA(cudaMalloc(&d_c1, colb));
    SPT_CUDA(cudaMalloc(&d_dg, colb));
    SPT_CUDA(cudaMalloc(&d_begin, begin.size() * sizeof(long long)));
    SPT_CUDA(cudaMalloc(&d_K, Ks.size() * sizeof(int)));
    SPT_CUDA(cudaMemcpy(d_x, xcos.data(), g.nleg * sizeof(double), cudaMemcpyHostToDevice));
    SPT_CUDA(cudaMemcpy(d_c0, col0.data(), colb, cudaMemcpyHostToDevice));
    SPT_CUDA(cudaMemcpy(d_c1, col1.data(), colb, cudaMemcpyHostToDevice));
    SPT_CUDA(cudaMemcpy(d_dg, diag.data(), colb, cudaMemcpyHostToDevice));
    SPT_CUDA(cudaMemcpy(d_begin, begin.data(), begin.size() * sizeof(long long), cudaMemcpyHostToDevice));
    SPT_CUDA(cudaMemcpy(d_K, Ks.data(), Ks.size() * sizeof(int), cudaMemcpyHostToDevice));
    if (3ull * W * 4 * sizeof(double) <= 200 * 1024) {
        const size_t smem = 3ull * W * 4 * sizeof(double);
        SPT_CUDA(cudaFuncSetAttribute(legendre_gen_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        legendre_gen_kernel<true, 4><<<(g.nleg + 3) / 4, kGenThreads, smem, p.stream>>>(trc, trc, g.nleg, d_x, d_c0, d_c1, d_dg,
                                                                                      p.d_nlat0, d_begin, d_K, d_out);
    }
    else {
        const size_t smem = 3ull * W * 2 * sizeof(double);
        SPT_CUDA(cudaFuncSetAttribute(legendre_gen_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        legendre_gen_kernel<true, 2><<<(g.nleg + 1) / 2, kGenThreads, smem, p.stream>>>(trc, trc, g.nleg, d_x, d_c0, d_c1, d_dg,
                                                                                      p.d_nlat0, d_begin, d_K, d_out);
    }
    SPT_CUDA(cudaGetLastError());
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    SPT_CUDA(cudaMemcpy(h_out, d_out, total * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_out);
    cudaFree(d_x);
    cudaFree(d_c0);
    cudaFree(d_c1);
    cudaFree(d_dg);
    cudaFree(d_begin);
    cudaFree(d_K);
    return SPTRANS_OK;
}

namespace {
// reference cache block (m, parity): index k_ref + K * lat, k_ref = 0 <-> highest n   ->   table block [k][lat - nlat0[m]], n ascending
__global__ void import_cache_kernel(int nleg, const int* __restrict__ nlat0, const long long* __restrict__ ref_begin,
                                    const int* __restrict__ ref_K, const long long* __restrict__ tab_off,
                                    const int* __restrict__ tab_pitch, const double* __restrict__ blob,
                                    double* __restrict__ tab) {
    const int b = blockIdx.y;  // 2 m + parity
    const int m = b >> 1;
    const int K = ref_K[b];
    const int l0 = nlat0[m];
    const int ncol = nleg - l0;
    if (K <= 0 || ncol <= 0) return;
    const long long total = static_cast<long long>(K) * ncol;
    const double* src = blob + ref_begin[b];
    double* dst = tab + tab_off[b];
    const int pitch = tab_pitch[m];
    for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int lat = static_cast<int>(e / K), kr = static_cast<int>(e - static_cast<long long>(lat) * K);  // read order
        dst[static_cast<long long>(K - 1 - kr) * pitch + lat] = src[static_cast<long long>(K) * (l0 + lat) + kr];
    }
}
}  // namespace

// Replace the contents of the device Legendre table by the values of a reference-layout cache blob (what
// sptrans_export_legendre_cache writes, and what the reference's LegendreCacheCreatorLocal / TransLocal write:
// TransLocal.cc:592-647).  Entries the table does not hold (latitudes below nlat0[m], the m = T+1 blocks) are skipped.
int import_legendre_cache(Plan& p, const double* h_blob, size_t bytes) {
    const HostGeom& g = p.g;
    const int T = g.T, trc = T + 1;
    if (bytes != legendre_cache_doubles(g) * sizeof(double)) {
        set_error("sptrans_import_legendre_cache: blob size " + std::to_string(bytes) + " does not match this grid / truncation (" +
                  std::to_string(legendre_cache_doubles(g) * sizeof(double)) + " bytes expected)");
        return SPTRANS_ERR_INVALID;
    }
    if (g.nranks != 1) {
        set_error("sptrans_import_legendre_cache: not available for sharded plans");
        return SPTRANS_ERR_NOT_IMPLEMENTED;
    }
    std::vector<long long> begin(2 * (T + 1), 0);
    std::vector<int> Ks(2 * (T + 1), 0);
    long long size_sym = 0, size_asym = 0;
    auto pad8 = [](long long n) { return (n + 7) / 8 * 8; };
    for (int m = 0; m <= T + 1; ++m) {
        const int ks = num_n(trc, m, 0), ka = num_n(trc, m, 1);
        if (m <= T) {
            Ks[2 * m] = ks;
            Ks[2 * m + 1] = ka;
            begin[2 * m] = size_sym;
            begin[2 * m + 1] = size_asym;
        }
        size_sym += pad8(static_cast<long long>(ks) * g.nleg);
        size_asym += pad8(static_cast<long long>(ka) * g.nleg);
    }
    for (int m = 0; m <= T; ++m) begin[2 * m + 1] += size_sym;
    double* d_blob = nullptr;
    long long *d_begin = nullptr, *d_off = nullptr;
    int *d_K = nullptr, *d_pitch = nullptr;
    SPT_CUDA(cudaMalloc(&d_blob, std::max<size_t>(bytes, 8)));
    SPT_CUDA(cudaMalloc(&d_begin, begin.size() * sizeof(long long)));
    SPT_CUDA(cudaMalloc(&d_off, g.tab_off.size() * sizeof(long long)));
    SPT_CUDA(cudaMalloc(&d_K, Ks.size() * sizeof(int)));
    SPT_CUDA(cudaMalloc(&d_pitch, g.tab_pitch.size() * sizeof(int)));
    SPT_CUDA(cudaMemcpy(d_blob, h_blob, bytes, cudaMemcpyHostToDevice));
    SPT_CUDA(cudaMemcpy(d_begin, begin.data(), begin.size() * sizeof(long long), cudaMemcpyHostToDevice));
    SPT_CUDA(cudaMemcpy(d_off, g.tab_off.data(), g.tab_off.size() * sizeof(long long), cudaMemcpyHostToDevice));
    SPT_CUDA(cudaMemcpy(d_K, Ks.data(), Ks.size() * sizeof(int), cudaMemcpyHostToDevice));
    SPT_CUDA(cudaMemcpy(d_pitch, g.tab_pitch.data(), g.tab_pitch.size() * sizeof(int), cudaMemcpyHostToDevice));
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    dim3 grid(64, 2 * (T + 1));
    import_cache_kernel<<<grid, 256, 0, p.stream>>>(g.nleg, p.d_nlat0, d_begin, d_K, d_off, d_pitch, d_blob, p.d_tab);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    cudaFree(d_blob);
    cudaFree(d_begin);
    cudaFree(d_off);
    cudaFree(d_K);
    cudaFree(d_pitch);
    tc_free(p);  // tensor-core operand images are derived from this table: rebuilt on next use
    return p.d_tabT ? build_transposed_table(p) : SPTRANS_OK;  // keep an existing transpose in step
}

size_t legendre_cache_doubles(const HostGeom& g) {
    auto pad8 = [](long long n) { return (n + 7) / 8 * 8; };
    long long tot = 0;
    for (int m = 0; m <= g.T + 1; ++m)
        tot += pad8(static_cast<long long>(num_n(g.T + 1, m, 0)) * g.nleg) +
               pad8(static_cast<long long>(num_n(g.T + 1, m, 1)) * g.nleg);
    return static_cast<size_t>(tot);
}

}  // namespace sptrans
