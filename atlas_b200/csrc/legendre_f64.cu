// fp64 Legendre transforms as a ragged batched GEMM on the FP64 tensor path (DMMA, mma.sync m8n8k4.f64).
//
// Replaces TransLocal::invtrans_legendre (ecmwf/atlas src/atlas/trans/local/TransLocal.cc:939-1097: per-m
// split into sym/antisym, two eckit gemm calls, hemisphere merge) and, for the direct transform, the
// Legendre stage the reference only has through ectrans (ifs/TransIFS.cc:503-515).
//
// tcgen05 has no f64 kind (ptxas: "Unknown modifier .kind::f64"), so the fp64 configurations run on
// DMMA; measured peak on B200: 37.1 TFLOP/s for both DFMA and DMMA (profiles/microbench_f64_r01.txt),
// reached from shared memory only with >= 32x56 warp tiles -- hence the 32x72 warp tile below.
//
// Work decomposition: one tile = (m, parity, 128-row block, 144-column block) of
//     inverse:  C[lat][r] = sum_k  P[k][lat] * S[k][r]        (A = P^T, K-major)
//     direct :  X[k][r]   = sum_lat P[k][lat] * G[lat][r]     (A = P,   row-major)
// with r = 2*field + (re|im).  Tiles are sorted by cost and claimed through an atomic counter by a
// persistent grid of one CTA per SM (8 warps, 4 x 2, each 32 rows x 72 columns = 4 x 9 DMMA tiles).
// Operands are staged by 16-byte cp.async into a 4-stage shared-memory ring; row pitches are
// = 4 (mod 16) doubles so every fragment load is bank-conflict free.
//
// The hemisphere merge of the reference (:1034-1079) is NOT done here: the Fourier kernels read the
// symmetric / antisymmetric parts and form north = s + a, south = s - a on the fly (fourier.cu).
#include <algorithm>
#include <cstdio>
#include <vector>

#include "plan.hpp"

namespace sptrans {

namespace {

constexpr int kWarpsM = 4, kWarpsN = 2;
constexpr int kMI = 4;   // 8-row DMMA tiles per warp  (32 rows)
constexpr int kNJ = 9;   // 8-col DMMA tiles per warp  (72 cols)
static_assert(kWarpsM * kMI * 8 == kBM && kWarpsN * kNJ * 8 == kBN, "tile shape");

constexpr int kBPitch = kBN + 4;     // Bs[kBK][kBN+4]
constexpr int kBStage = kBK * kBPitch;
#ifndef SPT_BULK
#define SPT_BULK 1   // 1: TMA-fed, warp-decoupled kernel (legendre_bulk.cuh); 0: per-thread cp.async baseline below
#endif
#if !SPT_BULK
constexpr int kAInvPitch = kBM + 4;  // inverse: As[kBK][kBM+4]
constexpr int kADirPitch = kBK + 4;  // direct : As[kBM][kBK+4]
constexpr int kAInvStage = kBK * kAInvPitch;
constexpr int kADirStage = kBM * kADirPitch;
constexpr int kAStage = (kAInvStage > kADirStage ? kAInvStage : kADirStage);
constexpr size_t kLegSmemBytes = static_cast<size_t>(kStages) * (kAStage + kBStage) * sizeof(double) + 16;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    const int sz = valid ? 16 : 0;  // src-size 0 => destination is zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
#endif
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

#if !SPT_BULK
// kDirect == false : A tile is [kBK][kBM] (k rows, latitude contiguous)   -> C rows are latitudes
// kDirect == true  : A tile is [kBM][kBK] (table rows, latitude contiguous) -> C rows are table rows (n)
template <bool kDirect, bool kPeers>
__global__ void __launch_bounds__(kLegThreads, 1)
legendre_dmma_kernel(const LegTile* __restrict__ tiles, int ntiles, int* __restrict__ counter,
                     const double* __restrict__ tab, const double* __restrict__ B, double* __restrict__ C, int ldb,
                     const __grid_constant__ PeerDst dst) {
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + kStages * kAStage;
    __shared__ LegTile s_tl[2];   // current / next tile descriptor (fetched one tile ahead by thread 0)
    __shared__ int s_ti[2];

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    // warp -> (row block, column block): consecutive warps (= the four SM sub-partitions) cover the first two row
    // blocks, so that a partial tile with <= 64 valid rows still keeps every sub-partition's FP64 pipe busy
    const int wm = warp / kWarpsN, wn = warp % kWarpsN;

    auto load_stage = [&](const LegTile& tl, int kb, int st) {
        const double* Ag = tab + tl.a_off;
        const double* Bg = B + tl.b_off;
        double* as = As + st * kAStage;
        double* bs = Bs + st * kBStage;
        if (!kDirect) {
            // kBK rows x kBM latitudes, 2 doubles per chunk
            for (int c = tid; c < kBK * (kBM / 2); c += kLegThreads) {
                const int kk = c / (kBM / 2), cc = (c % (kBM / 2)) * 2;
                const bool ok = cc < tl.a_rows;  // a_rows: readable latitude columns (even)
                const double* src = Ag + static_cast<long long>(kb * kBK + kk) * tl.a_pitch + (ok ? cc : 0);
                cp_async16(as + kk * kAInvPitch + cc, src, ok);
            }
        }
        else {
            // kBM table rows x kBK latitudes
            for (int c = tid; c < kBM * (kBK / 2); c += kLegThreads) {
                const int rr = c / (kBK / 2), cc = (c % (kBK / 2)) * 2;
                const bool ok = rr < tl.a_rows;  // a_rows: readable table rows
                const double* src = Ag + static_cast<long long>(ok ? rr : 0) * tl.a_pitch + kb * kBK + cc;
                cp_async16(as + rr * kADirPitch + cc, src, ok);
            }
        }
        for (int c = tid; c < kBK * (kBN / 2); c += kLegThreads) {
            const int kk = c / (kBN / 2), cc = (c % (kBN / 2)) * 2;
            const int row = kb * kBK + kk;
            const bool ok = (cc < tl.n_valid) && (row < tl.b_rows);
            const double* src = Bg + static_cast<long long>(ok ? row : 0) * ldb + (ok ? cc : 0);
            cp_async16(bs + kk * kBPitch + cc, src, ok);
        }
    };
    // first kStages-1 stages of a tile (one commit group each, as the main loop expects)
    auto prologue = [&](const LegTile& tl) {
#pragma unroll
        for (int s = 0; s < kStages - 1; ++s) {
            if (s < tl.k_steps) load_stage(tl, s, s);
            cp_async_commit();
        }
    };

    if (tid == 0) {
        const int ti = atomicAdd(counter, 1);
        s_ti[0] = ti;
        if (ti < ntiles) s_tl[0] = tiles[ti];
    }
    __syncthreads();
    if (s_ti[0] >= ntiles) return;
    prologue(s_tl[0]);

    for (int buf = 0;; buf ^= 1) {
        const LegTile tl = s_tl[buf];
        // claim the next tile and fetch its descriptor while this one computes
        if (tid == 0) {
            const int ti = atomicAdd(counter, 1);
            s_ti[buf ^ 1] = ti;
            if (ti < ntiles) s_tl[buf ^ 1] = tiles[ti];
        }

        double acc[kMI][kNJ][2];
#pragma unroll
        for (int i = 0; i < kMI; ++i)
#pragma unroll
            for (int j = 0; j < kNJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.;

        const int ksteps = tl.k_steps;
        const int row_w = wm * (kMI * 8);   // warp's first row in the tile
        const int col_w = wn * (kNJ * 8);   // warp's first column
        const bool warp_active = (row_w < tl.m_valid) && (col_w < tl.n_valid);

        for (int kb = 0; kb < ksteps; ++kb) {
            cp_async_wait<kStages - 2>();
            __syncthreads();
            {
                const int nk = kb + kStages - 1;
                if (nk < ksteps) load_stage(tl, nk, nk % kStages);
                cp_async_commit();
            }
            if (warp_active) {
                const double* as = As + (kb % kStages) * kAStage;
                const double* bs = Bs + (kb % kStages) * kBStage;
                // fragments are double buffered: the shared-memory loads of k-step ks+1 are issued before the
                // DMMAs of step ks, so that the two warps of a sub-partition (which run in lock step between
                // barriers) never wait on LDS latency with an idle FP64 pipe
                double a[2][kMI], b[2][kNJ];
                auto load_frag = [&](int ks, int fb) {
#pragma unroll
                    for (int i = 0; i < kMI; ++i) {
                        if (!kDirect) a[fb][i] = as[(ks * 4 + t) * kAInvPitch + row_w + 8 * i + g];
                        else a[fb][i] = as[(row_w + 8 * i + g) * kADirPitch + ks * 4 + t];
                    }
#pragma unroll
                    for (int j = 0; j < kNJ; ++j) b[fb][j] = bs[(ks * 4 + t) * kBPitch + col_w + 8 * j + g];
                };
                load_frag(0, 0);
#pragma unroll
                for (int ks = 0; ks < kBK / 4; ++ks) {
                    if (ks + 1 < kBK / 4) load_frag(ks + 1, (ks + 1) & 1);
#pragma unroll
                    for (int i = 0; i < kMI; ++i) {
                        if (row_w + 8 * i < tl.m_valid) {
#pragma unroll
                            for (int j = 0; j < kNJ; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[ks & 1][i], b[ks & 1][j]);
                        }
                    }
                }
            }
        }
        cp_async_wait<0>();
        __syncthreads();  // every stage buffer is free again; the next descriptor written by thread 0 is visible
        const int next_ti = s_ti[buf ^ 1];
        // start streaming the next tile's operands before this tile's results are written out
        if (next_ti < ntiles) prologue(s_tl[buf ^ 1]);
        // epilogue: thread holds C[row_w+8i+g][col_w+8j+2t .. +1] = (re, im) of one field
        if (warp_active) {
#pragma unroll
            for (int i = 0; i < kMI; ++i) {
                const int row = row_w + 8 * i + g;
                if (row < tl.m_valid) {
                    double* Cg;
                    if (!kPeers) Cg = C + tl.c_off;
                    else {
                        // fused exchange: the row goes straight into the Fourier-side buffer of the rank whose
                        // latitude band contains it (NVLink store; same layout on every rank)
                        const int lat = tl.lat0 + row;
                        int d = 0;
                        while (d + 1 < dst.nranks && lat >= dst.band[d + 1]) ++d;
                        Cg = dst.base[d] + tl.c_off;
                    }
#pragma unroll
                    for (int j = 0; j < kNJ; ++j) {
                        const int col = col_w + 8 * j + 2 * t;
                        if (col < tl.n_valid) {
                            double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
                            *reinterpret_cast<double2*>(Cg + static_cast<long long>(row) * ldb + col) = v;
                        }
                    }
                }
            }
        }
        if (next_ti >= ntiles) break;
    }
    if (kPeers) __threadfence_system();  // remote rows are visible to the peers before this kernel completes
}

#else
#include "legendre_bulk.cuh"
#endif  // !SPT_BULK

// spectra [m][n][re/im][fld]  ->  packed [m][p][k][2 fld + re/im], rows k >= K_eff (and whole blocks with
// m >= trunc) are zero: this is the split + zero padding of TransLocal.cc:970-1003, including the
// `jn <= truncation && jm < truncation` rule (:982) that drops the m == truncation column.
constexpr int kPackRows = 16;  // rows of one (m, parity) block per CTA: keeps the m = 0 blocks from being the long pole

// flags & kPackKeepMT: the m == trunc column is kept (unstructured point sets, TransLocal.cc:1331; adjoint of the direct
// transform); flags & kPackDirAdj: operand of the adjoint of the DIRECT transform -- Im(m = 0) dropped.  (The transpose
// over the reals would halve m > 0 because the inverse Fourier kernel doubles them; the adjoint w.r.t. the ectrans /
// TransIFS spectral inner product, which counts m > 0 twice, multiplies them by 2 again: test_transgeneral.cc:1683-1703.)
__global__ void pack_spectra_kernel(int /*T*/, int nf, int trunc, const long long* __restrict__ sp_rowoff,
                                    const int* __restrict__ my_m, const double* __restrict__ spec,
                                    double* __restrict__ packed, int flags, const long long* __restrict__ spec_off) {
    const int m = my_m[blockIdx.x];
    const int p = blockIdx.y;
    const long long row0 = sp_rowoff[2 * m + p];
    const int rows = static_cast<int>(sp_rowoff[2 * m + p + 1] - row0);
    const int k0 = blockIdx.z * kPackRows;
    if (k0 >= rows) return;
    const int kn = min(kPackRows, rows - k0);
    // reference :970; spec_off: the spectral array holds only this rank's zonal wavenumbers (SPTRANS_SHARD_LOCAL_IO)
    const long long ioff = (spec_off ? spec_off[m] : static_cast<long long>(2 * trunc + 3 - m) * m / 2) * nf * 2;
    // a thread moves one (re, im) pair: two 8-byte loads along the fields (each stream coalesced), one 16-byte store
    const bool keep = (m < trunc || (flags & kPackKeepMT));
    const bool drop_im = (flags & kPackDirAdj) && m == 0;
    double2* __restrict__ out = reinterpret_cast<double2*>(packed);
    for (int e = threadIdx.x; e < kn * nf; e += blockDim.x) {
        const int k = k0 + e / nf;
        const int f = e % nf;
        const int n = m + p + 2 * k;
        double2 v = make_double2(0., 0.);
        if (n <= trunc && keep) {
            const double* src = spec + ioff + static_cast<long long>(nf) * (2 * (n - m)) + f;
            v.x = src[0];
            v.y = drop_im ? 0. : src[nf];
        }
        out[(row0 + k) * nf + f] = v;
    }
}

// packed [m][p][k][2 fld + re/im] -> spectra [m][n][re/im][fld] for all n <= T
__global__ void unpack_spectra_kernel(int T, int nf, const long long* __restrict__ sp_rowoff,
                                      const int* __restrict__ my_m, const double* __restrict__ packed,
                                      double* __restrict__ spec, int drop_mT, const long long* __restrict__ spec_off) {
    const int m = my_m[blockIdx.x];
    const int p = blockIdx.y;
    const long long row0 = sp_rowoff[2 * m + p];
    const int K = (T - m + (p ? 1 : 2)) / 2;  // n <= T
    const int k0 = blockIdx.z * kPackRows;
    if (k0 >= K) return;
    const int kn = min(kPackRows, K - k0);
    const long long ioff = (spec_off ? spec_off[m] : static_cast<long long>(2 * T + 3 - m) * m / 2) * nf * 2;
    const double2* __restrict__ in = reinterpret_cast<const double2*>(packed);
    const bool zero_all = drop_mT && m >= T;  // adjoint of the scalar inverse, which ignores the m == T column (:982)
    for (int e = threadIdx.x; e < kn * nf; e += blockDim.x) {
        const int k = k0 + e / nf;
        const int f = e % nf;
        const int n = m + p + 2 * k;
        double2 v = in[(row0 + k) * nf + f];
        if (m == 0) v.y = 0.;  // Im of the zonal-mean coefficients is identically zero
        if (zero_all) v = make_double2(0., 0.);
        double* dst = spec + ioff + static_cast<long long>(nf) * (2 * (n - m)) + f;
        dst[0] = v.x;
        dst[nf] = v.y;
    }
}

}  // namespace

// Tile lists for the current (nf, truncation-of-data).  Cheap (tens of thousands of entries); cached.
int build_tiles(Plan& p, int nf, int trunc, int dir_trunc) {
    if (p.tiles_nf == nf && p.tiles_trunc == trunc && p.tiles_dir_trunc == dir_trunc) return SPTRANS_OK;
    const HostGeom& g = p.g;
    const int ld = 2 * nf;
    struct Keyed {
        double cost;
        LegTile t;
    };
    std::vector<Keyed> inv, dir;
    for (int m : g.my_m) {
        const int ncol = g.nleg - g.nlat0[m];
        if (ncol <= 0) continue;
        const int pitch = g.tab_pitch[m];
        for (int par = 0; par < 2; ++par) {
            const int Ktab = g.tab_K[2 * m + par];
            const long long fb_row0 = g.fb_rowoff[m] + static_cast<long long>(par) * ncol;  // double2 rows of nf
            // ---- inverse: K = #n <= trunc of this parity (0 if m >= trunc: column dropped, :982)
            int Kinv = (m < trunc) ? std::min(Ktab, num_n(trunc, m, par)) : 0;
            if (Kinv > 0) {
                const int ksteps = round_up(Kinv, kBK) / kBK;
                for (int l0 = 0; l0 < ncol; l0 += kBM) {
                    // cropped plans: only the latitude pairs that hold rows of the crop are evaluated
                    if (g.cropped && (g.nlat0[m] + l0 + kBM <= g.pair_begin || g.nlat0[m] + l0 >= g.pair_end)) continue;
                    for (int n0 = 0; n0 < ld; n0 += kBN) {
                        LegTile t{};
                        t.a_off = g.tab_off[2 * m + par] + l0;
                        t.a_pitch = pitch;
                        t.a_rows = pitch - l0;
                        t.b_off = g.sp_rowoff[2 * m + par] * ld + n0;
                        t.b_rows = ksteps * kBK;
                        t.c_off = (fb_row0 + l0) * ld + n0;
                        t.lat0 = g.nlat0[m] + l0;
                        t.m_valid = std::min(kBM, ncol - l0);
                        t.n_valid = std::min(kBN, ld - n0);
                        t.k_steps = ksteps;
                        t.ks_last = (Kinv - (ksteps - 1) * kBK + 3) / 4;
                        inv.push_back({static_cast<double>(ksteps) * round_up(t.m_valid, 8), t});
                    }
                }
            }
            // ---- direct: rows = all n <= dir_trunc (T, or T+1 for the wind path) of this parity, contraction over latitudes
            const int Kdir = std::min(Ktab, num_n(dir_trunc, m, par));
            if (Kdir > 0) {
                const int ksteps = pitch / kBK;
                const int rows_tab = round_up(std::max(Ktab, 1), kBK);
                for (int r0 = 0; r0 < Kdir; r0 += kBM) {
                    for (int n0 = 0; n0 < ld; n0 += kBN) {
                        LegTile t{};
#if SPT_BULK
                        // A = transposed table block [pitch latitudes][rows_tab n-rows]: K-major like the inverse's
                        t.a_off = g.tab_off[2 * m + par] + r0;
                        t.a_pitch = rows_tab;
                        t.a_rows = rows_tab - r0;
#else
                        t.a_off = g.tab_off[2 * m + par] + static_cast<long long>(r0) * pitch;
                        t.a_pitch = pitch;
                        t.a_rows = rows_tab - r0;
#endif
                        t.b_off = fb_row0 * ld + n0;
                        t.b_rows = ncol;
                        t.lat0 = g.nlat0[m];   // latitude pair of B row 0 (selects the source rank when the rows are pulled from peers)
                        t.c_off = (g.sp_rowoff[2 * m + par] + r0) * ld + n0;
                        t.m_valid = std::min(kBM, Kdir - r0);
                        t.n_valid = std::min(kBN, ld - n0);
                        t.k_steps = ksteps;
                        t.ks_last = (ncol - (ksteps - 1) * kBK + 3) / 4;
                        dir.push_back({static_cast<double>(ksteps) * round_up(t.m_valid, 8), t});
                    }
                }
            }
        }
    }
    auto finish = [&](std::vector<Keyed>& v, LegTile*& d_ptr, int& count) -> int {
        std::stable_sort(v.begin(), v.end(), [](const Keyed& a, const Keyed& b) { return a.cost > b.cost; });
        std::vector<LegTile> flat(v.size());
        for (size_t i = 0; i < v.size(); ++i) flat[i] = v[i].t;
        if (d_ptr) cudaFree(d_ptr);
        d_ptr = nullptr;
        count = static_cast<int>(flat.size());
        SPT_CUDA(cudaMalloc(&d_ptr, std::max<size_t>(flat.size(), 1) * sizeof(LegTile)));
        SPT_CUDA(cudaMemcpyAsync(d_ptr, flat.data(), flat.size() * sizeof(LegTile), cudaMemcpyHostToDevice, p.stream));
        SPT_CUDA(cudaStreamSynchronize(p.stream));
        return SPTRANS_OK;
    };
    int rc = finish(inv, p.d_tiles_inv, p.n_tiles_inv);
    if (rc) return rc;
    rc = finish(dir, p.d_tiles_dir, p.n_tiles_dir);
    if (rc) return rc;
    p.tiles_nf = nf;
    p.tiles_trunc = trunc;
    p.tiles_dir_trunc = dir_trunc;
    return SPTRANS_OK;
}

int launch_pack_spectra(Plan& p, int nf, int trunc, const double* d_spec, double* d_packed, int flags) {
    const int nm = static_cast<int>(p.g.my_m.size());
    if (nm == 0) return SPTRANS_OK;
    dim3 grid(nm, 2, (round_up(p.g.T / 2 + 2, kBK) + kPackRows - 1) / kPackRows);
    if (p.d_spec_off && trunc != p.g.T) {
        set_error("SPTRANS_SHARD_LOCAL_IO: spectral arrays hold this rank's zonal wavenumbers at truncation T only");
        return SPTRANS_ERR_INVALID;
    }
    pack_spectra_kernel<<<grid, 256, 0, p.stream>>>(p.g.T, nf, trunc, p.d_sp_rowoff, p.d_my_m, d_spec, d_packed, flags,
                                                    p.d_spec_off);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int launch_unpack_spectra(Plan& p, int nf, const double* d_packed, double* d_spec, int drop_mT) {
    const int nm = static_cast<int>(p.g.my_m.size());
    if (nm == 0) return SPTRANS_OK;
    dim3 grid(nm, 2, (p.g.T / 2 + 1 + kPackRows - 1) / kPackRows);
    unpack_spectra_kernel<<<grid, 256, 0, p.stream>>>(p.g.T, nf, p.d_sp_rowoff, p.d_my_m, d_packed, d_spec, drop_mT,
                                                      p.d_spec_off);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

template <bool kDirect, bool kPeers>
static int launch_gemm(Plan& p, int nf, const LegTile* tiles, int ntiles, const double* B, double* C,
                       const PeerDst& dst) {
    if (ntiles == 0) return SPTRANS_OK;
#if SPT_BULK
    constexpr size_t smem_bytes = kBulkSmemBytes;
    if (kDirect && !p.d_tabT) {  // first direct transform of this plan: inverse-only users never pay for the second table
        int rc = build_transposed_table(p);
        if (rc) return rc;
    }
    const double* table = kDirect ? p.d_tabT : p.d_tab;  // the direct transform contracts over latitudes: P^T is K-major
#else
    constexpr size_t smem_bytes = kLegSmemBytes;
    const double* table = p.d_tab;
#endif
    auto kernel = legendre_dmma_kernel<kDirect, kPeers>;
    // the opt-in is per device (context), and one process may hold plans on several devices: remembered per ordinal
    static bool attr_set[4][64] = {};
    if (p.device >= 64 || !attr_set[2 * kDirect + kPeers][p.device]) {
        SPT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_bytes)));
        if (p.device < 64) attr_set[2 * kDirect + kPeers][p.device] = true;
    }
    SPT_CUDA(cudaMemsetAsync(p.d_tile_counter, 0, sizeof(int), p.stream));
    const int grid = std::min(ntiles, p.num_sms);
    kernel<<<grid, kLegThreads, smem_bytes, p.stream>>>(tiles, ntiles, p.d_tile_counter, table, B, C, 2 * nf, dst);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int launch_legendre_inv(Plan& p, int nf, const double* d_packed, double* d_fourier) {
    return launch_gemm<false, false>(p, nf, p.d_tiles_inv, p.n_tiles_inv, d_packed, d_fourier, PeerDst{});
}
int launch_legendre_inv_peers(Plan& p, int nf, const double* d_packed, const PeerDst& dst) {
    return launch_gemm<false, true>(p, nf, p.d_tiles_inv, p.n_tiles_inv, d_packed, nullptr, dst);
}
int launch_legendre_dir(Plan& p, int nf, const double* d_fourier, double* d_packed) {
    return launch_gemm<true, false>(p, nf, p.d_tiles_dir, p.n_tiles_dir, d_fourier, d_packed, PeerDst{});
}
int launch_legendre_dir_peers(Plan& p, int nf, const PeerDst& src, double* d_packed) {
    return launch_gemm<true, true>(p, nf, p.d_tiles_dir, p.n_tiles_dir, nullptr, d_packed, src);
}

namespace {
// one (m, parity) block: in [rows][pitch] -> out [pitch][rows]  (rows = n-rows padded to kBK, pitch = latitudes padded to 16)
__global__ void transpose_table_kernel(const long long* __restrict__ tab_off, const int* __restrict__ tab_K,
                                       const int* __restrict__ tab_pitch, const double* __restrict__ in,
                                       double* __restrict__ out) {
    __shared__ double tile[32][33];
    const int b = blockIdx.z;
    const long long off = tab_off[b];
    if (off < 0) return;  // block of another rank
    const int rows = (max(tab_K[b], 1) + kBK - 1) / kBK * kBK, pitch = tab_pitch[b >> 1];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    if (r0 >= rows || c0 >= pitch) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int r = ty; r < 32; r += 8)
        if (r0 + r < rows && c0 + tx < pitch) tile[r][tx] = in[off + static_cast<long long>(r0 + r) * pitch + c0 + tx];
    __syncthreads();
    for (int c = ty; c < 32; c += 8)
        if (c0 + c < pitch && r0 + tx < rows) out[off + static_cast<long long>(c0 + c) * rows + r0 + tx] = tile[tx][c];
}
}  // namespace

// P^T for the direct transform (same block offsets and sizes as the table itself); a no-op for the cp.async baseline
int build_transposed_table(Plan& p) {
#if SPT_BULK
    const HostGeom& g = p.g;
    const size_t tab_bytes = static_cast<size_t>(g.tab_size) * sizeof(double);
    if (!p.d_tabT) {
        SPT_CUDA(cudaMalloc(&p.d_tabT, std::max<size_t>(tab_bytes, 16)));
        p.bytes_tables += tab_bytes;
    }
    SPT_CUDA(cudaMemsetAsync(p.d_tabT, 0, tab_bytes, p.stream));
    long long* d_off = nullptr;
    int *d_K = nullptr, *d_pitch = nullptr;
    SPT_CUDA(cudaMalloc(&d_off, g.tab_off.size() * sizeof(long long)));
    SPT_CUDA(cudaMalloc(&d_K, g.tab_K.size() * sizeof(int)));
    SPT_CUDA(cudaMalloc(&d_pitch, g.tab_pitch.size() * sizeof(int)));
    SPT_CUDA(cudaMemcpyAsync(d_off, g.tab_off.data(), g.tab_off.size() * sizeof(long long), cudaMemcpyHostToDevice, p.stream));
    SPT_CUDA(cudaMemcpyAsync(d_K, g.tab_K.data(), g.tab_K.size() * sizeof(int), cudaMemcpyHostToDevice, p.stream));
    SPT_CUDA(cudaMemcpyAsync(d_pitch, g.tab_pitch.data(), g.tab_pitch.size() * sizeof(int), cudaMemcpyHostToDevice, p.stream));
    const int max_rows = round_up(g.T / 2 + 2, kBK), max_pitch = round_up(std::max(g.nleg, 1), 16);
    dim3 grid((max_pitch + 31) / 32, (max_rows + 31) / 32, 2 * (g.T + 1));
    transpose_table_kernel<<<grid, dim3(32, 8), 0, p.stream>>>(d_off, d_K, d_pitch, p.d_tab, p.d_tabT);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    cudaFree(d_off);
    cudaFree(d_K);
    cudaFree(d_pitch);
#endif
    return SPTRANS_OK;
}

}  // namespace sptrans
