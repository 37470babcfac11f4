// Mixed-precision Legendre transforms on the 5th-generation tensor cores (BASELINE config 4):
// tcgen05.mma kind::tf32 with fp32 accumulators in tensor memory, operands staged by the TMA engine
// (cp.async.bulk, 1-D bulk copies completing on mbarriers).
//
// Replaces, like legendre_f64.cu, TransLocal::invtrans_legendre (ecmwf/atlas src/atlas/trans/local/
// TransLocal.cc:939-1097) and the direct Legendre stage of ectrans, but at fp32-level accuracy: every fp64
// operand x is split as x = hi + lo with hi = tf32(x), lo = fp32(x - hi), and each product is formed as
// hi*hi' + lo*hi' + hi*lo'  ("3xTF32"), which restores fp32 product accuracy from the 10-bit tf32 mantissa.
// tcgen05 has no f64 kind, so the fp64 configurations stay on the DMMA kernel.
//
// Operand format.  tcgen05.mma reads both operands from shared memory through 64-bit descriptors.  Instead
// of TMA tensor maps (one descriptor per ragged (m, parity) block would be needed) the Legendre tables are
// stored in HBM *as ready-made shared-memory tile images*: K-major, SWIZZLE_NONE canonical layout, i.e.
// 8-row x 16-byte core matrices stored contiguously,
//      float index of element (row, k) in an R x 16 image = ((k/4) * (R/8) + row/8) * 32 + (row%8) * 4 + k%4
// so that one contiguous cp.async.bulk per operand and pipeline stage fills the stage buffer and the
// descriptors are (start, LBO = (R/8)*128 bytes between K-adjacent core matrices, SBO = 128 bytes between
// 8-row groups).  The descriptor semantics were established on hardware with scratch/tc/tc_probe.cu
// (profiles/tcgen05_probe_r01.txt).
//
// Kernel structure (one persistent CTA per SM, 8 warps):
//   warp 0 lane 0 : producer  -- waits `empty[s]`, arms `full[s]` with the stage byte count, issues 4 bulk copies
//   warp 1 lane 0 : MMA issuer -- waits `full[s]`, issues 2 k-steps x 3 products x n_inst MMAs, commits to `empty[s]`;
//                    after the last K chunk commits to `tmem_full`
//   warps 4..7    : epilogue  -- wait `tmem_full`, tcgen05.ld 32x32b (warp w owns TMEM lanes 32(w%4)..: a thread holds 16
//                    columns of ONE row), convert fp32 -> fp64, transpose the 32 x 16 block through the warp's own 4 KB of
//                    shared memory so that the global stores are 128-byte runs along the rows of C (a thread storing its
//                    own row made every store instruction touch 32 cache lines: the epilogue of a 128 x 288 tile then
//                    costs as many LSU cycles as the tile's MMAs take), arrive on `tmem_empty`
// Tiles are assigned round-robin from a cost-sorted list, so all roles walk the same sequence without
// communicating.
#include <algorithm>
#include <cstdio>
#include <vector>

#include "plan.hpp"

namespace sptrans {

namespace {

constexpr int kTcThreads = 256;
constexpr int kTcM = 128;        // tile rows = TMEM lanes
constexpr int kTcKC = 16;        // K chunk per pipeline stage (two K=8 MMA steps)
constexpr int kTcStages = 4;      // at most; fewer when the field count makes the stages larger (TcParams::n_stages)
constexpr int kTcEpiPitch = 17;   // doubles per row of an epilogue warp's 32 x 16 staging block (odd: conflict-free both ways)
constexpr size_t kTcEpiBytes = 4ull * 32 * kTcEpiPitch * sizeof(double);
constexpr int kTcMaxN = 512;     // TMEM columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // bounded spin: a protocol bug ends in a trap (reported as a CUDA error), never in a hung device
    for (long long it = 0; it < (1ll << 31); ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// K-major SWIZZLE_NONE descriptor: leading offset = byte distance of K-adjacent core matrices,
// stride offset = byte distance of adjacent 8-row groups; bit 46 = sm_100 descriptor version
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// One tile of the tensor-core batched GEMM C[128 x N] = A[128 x K] B[N x K]^T (fp32 images, K-major).
struct alignas(16) TcTile {
    long long a_img;   // float offset of the first A image (hi table; lo table uses the same offset)
    long long b_img;   // float offset of the first B image
    long long c_off;   // double offset of C(row 0, col 0)
    int n_chunks;      // K chunks of kTcKC
    int m_valid;       // rows of C to store
    int pad0, pad1;
};

struct TcParams {
    const TcTile* tiles;
    int ntiles;
    const float* a_hi;
    const float* a_lo;
    const float* b_hi;
    const float* b_lo;
    double* C;
    int ldc;          // doubles
    int n_cols;       // valid columns of C (2 nf)
    int n_tot;        // padded columns (multiple of 16) = rows of a B image
    int n_inst;       // MMA instructions per k-step along N
    int n_each;       // columns per MMA instruction (multiple of 16, <= 256)
    int n_stages;     // depth of the operand ring (<= kTcStages)
};

__global__ void __launch_bounds__(kTcThreads, 1) legendre_tc_kernel(TcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar_full[kTcStages], bar_empty[kTcStages], bar_tmem_full, bar_tmem_empty;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t a_bytes = kTcM * kTcKC * 4;               // one A image
    const uint32_t b_bytes = static_cast<uint32_t>(p.n_tot) * kTcKC * 4;  // one B image
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    uint8_t* stage_base = smem;

    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_tmem_full, 1);
        mbar_init(&bar_tmem_empty, 4);  // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {  // ===== producer =====
            int stage = 0;
            uint32_t phase = 0;
            for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x) {
                const TcTile tl = p.tiles[ti];
                for (int c = 0; c < tl.n_chunks; ++c) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    uint8_t* sb = stage_base + static_cast<size_t>(stage) * stage_bytes;
                    mbar_expect_tx(&bar_full[stage], stage_bytes);
                    const long long ao = tl.a_img + static_cast<long long>(c) * (kTcM * kTcKC);
                    const long long bo = tl.b_img + static_cast<long long>(c) * (static_cast<long long>(p.n_tot) * kTcKC);
                    bulk_g2s(sb, p.a_hi + ao, a_bytes, &bar_full[stage]);
                    bulk_g2s(sb + a_bytes, p.a_lo + ao, a_bytes, &bar_full[stage]);
                    bulk_g2s(sb + 2 * a_bytes, p.b_hi + bo, b_bytes, &bar_full[stage]);
                    bulk_g2s(sb + 2 * a_bytes + b_bytes, p.b_lo + bo, b_bytes, &bar_full[stage]);
                    if (++stage == p.n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    }
    else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            int stage = 0;
            uint32_t phase = 0, tphase = 0;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(p.n_each >> 3) << 17) |
                                   (static_cast<uint32_t>(kTcM >> 4) << 24);
            const uint32_t a_lbo = (kTcM / 8) * 128, b_lbo = static_cast<uint32_t>(p.n_tot / 8) * 128;
            for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x) {
                const TcTile tl = p.tiles[ti];
                mbar_wait(&bar_tmem_empty, tphase ^ 1);  // accumulator drained by the epilogue of the previous tile
                asm volatile("tcgen05.fence::after_thread_sync;");
                for (int c = 0; c < tl.n_chunks; ++c) {
                    mbar_wait(&bar_full[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t sb = smem_u32(stage_base + static_cast<size_t>(stage) * stage_bytes);
                    const uint32_t sa_hi = sb, sa_lo = sb + a_bytes, sb_hi = sb + 2 * a_bytes, sb_lo = sb + 2 * a_bytes + b_bytes;
#pragma unroll
                    for (int ks = 0; ks < kTcKC / 8; ++ks) {
                        for (int h = 0; h < p.n_inst; ++h) {
                            const uint32_t boff = static_cast<uint32_t>(h * (p.n_each / 8)) * 128 + ks * 2 * b_lbo;
                            const uint32_t aoff = ks * 2 * a_lbo;
                            const uint32_t d = tmem_base + static_cast<uint32_t>(h * p.n_each);
                            const uint32_t first = (c == 0 && ks == 0) ? 0u : 1u;
                            umma_tf32(d, umma_desc(sa_hi + aoff, a_lbo, 128), umma_desc(sb_hi + boff, b_lbo, 128), idesc, first);
                            umma_tf32(d, umma_desc(sa_lo + aoff, a_lbo, 128), umma_desc(sb_hi + boff, b_lbo, 128), idesc, 1u);
                            umma_tf32(d, umma_desc(sa_hi + aoff, a_lbo, 128), umma_desc(sb_lo + boff, b_lbo, 128), idesc, 1u);
                        }
                    }
                    umma_commit(&bar_empty[stage]);  // stage buffer reusable once these MMAs have read it
                    if (++stage == p.n_stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&bar_tmem_full);
                tphase ^= 1;
            }
        }
    }
    else if (warp >= 4) {  // ===== epilogue =====
        uint32_t tphase = 0;
        const int lane_grp = warp & 3;  // TMEM lanes 32*lane_grp .. +31
        double* stg = reinterpret_cast<double*>(smem + static_cast<size_t>(p.n_stages) * stage_bytes) + lane_grp * (32 * kTcEpiPitch);
        const int half = lane >> 4, cl = lane & 15;
        for (int ti = blockIdx.x; ti < p.ntiles; ti += gridDim.x) {
            const TcTile tl = p.tiles[ti];
            mbar_wait(&bar_tmem_full, tphase);
            asm volatile("tcgen05.fence::after_thread_sync;");
            double* cblk = p.C + tl.c_off + static_cast<long long>(lane_grp * 32) * p.ldc;
            const int rows_valid = min(32, tl.m_valid - lane_grp * 32);   // rows of this warp's block inside the tile
            for (int c0 = 0; c0 < p.n_tot; c0 += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                if (rows_valid > 0 && c0 < p.n_cols) {   // (warp-uniform)
#pragma unroll
                    for (int j = 0; j < 16; ++j) stg[lane * kTcEpiPitch + j] = static_cast<double>(__uint_as_float(r[j]));
                    __syncwarp();
                    const int col = c0 + cl;
                    if (col < p.n_cols) {
#pragma unroll 4
                        for (int rr = half; rr < rows_valid; rr += 2)   // two rows per instruction, 128 bytes each
                            cblk[static_cast<long long>(rr) * p.ldc + col] = stg[rr * kTcEpiPitch + cl];
                    }
                    __syncwarp();
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tmem_empty);
            tphase ^= 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
}

// ---- operand preparation -----------------------------------------------------------------------------------

__device__ __forceinline__ void split_tf32(double x, float& hi, float& lo) {
    const float xf = static_cast<float>(x);
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(xf));
    hi = __uint_as_float(h);
    lo = static_cast<float>(x - static_cast<double>(hi));
}
// float index of element (row, k) inside an R x 16 image
__device__ __forceinline__ int img_index(int R, int row, int k) { return ((k >> 2) * (R >> 3) + (row >> 3)) * 32 + (row & 7) * 4 + (k & 3); }

// Legendre table images.  kDirect == false: A(row = latitude, k = wavenumber index) = P[k][lat]
//                         kDirect == true : A(row = wavenumber index, k = latitude)  = P[k_row][lat]
// one block per (m, parity); images ordered [row tile][K chunk]
template <bool kDirect>
__global__ void build_a_images_kernel(int T, const int* __restrict__ my_m, const int* __restrict__ nlat0, int nleg,
                                      const long long* __restrict__ tab_off, const int* __restrict__ tab_pitch,
                                      const int* __restrict__ tab_K, const double* __restrict__ tab,
                                      const long long* __restrict__ img_off, float* __restrict__ hi,
                                      float* __restrict__ lo) {
    const int m = my_m[blockIdx.x], par = blockIdx.y;
    const int ncol = nleg - nlat0[m];
    const int K = tab_K[2 * m + par];
    if (ncol <= 0 || K <= 0) return;
    const int pitch = tab_pitch[m];
    const double* P = tab + tab_off[2 * m + par];
    const int rows = kDirect ? K : ncol, kdim = kDirect ? ncol : K;
    const int row_tiles = (rows + kTcM - 1) / kTcM, chunks = (kdim + kTcKC - 1) / kTcKC;
    const long long base = img_off[2 * m + par];
    const long long total = static_cast<long long>(row_tiles) * chunks * (kTcM * kTcKC);
    for (long long e = threadIdx.x; e < total; e += blockDim.x) {
        const int in_img = static_cast<int>(e % (kTcM * kTcKC));
        const long long img = e / (kTcM * kTcKC);
        const int rt = static_cast<int>(img / chunks), ch = static_cast<int>(img % chunks);
        // decode in_img -> (row, k) by inverting img_index: iterate source-major instead: choose (row,k) from e
        const int row_l = in_img / kTcKC, k_l = in_img % kTcKC;  // source-order coordinates
        const int row = rt * kTcM + row_l, kk = ch * kTcKC + k_l;
        double v = 0.;
        if (row < rows && kk < kdim) v = kDirect ? P[static_cast<long long>(row) * pitch + kk] : P[static_cast<long long>(kk) * pitch + row];
        float h, l;
        split_tf32(v, h, l);
        const long long dst = base + img * (kTcM * kTcKC) + img_index(kTcM, row_l, k_l);
        hi[dst] = h;
        lo[dst] = l;
    }
}

// float index of the four K-consecutive elements (row, 4 q .. 4 q + 3) of an R-row B image sequence: 16-byte aligned
__device__ __forceinline__ long long img_quad(int R, int row, int q) {
    return static_cast<long long>(q >> 2) * (R * kTcKC) + ((q & 3) * (R >> 3) + (row >> 3)) * 32 + (row & 7) * 4;
}
__device__ __forceinline__ void store_split4(const double v[4], float* __restrict__ hi, float* __restrict__ lo, long long dst) {
    float4 h, l;
    split_tf32(v[0], h.x, l.x);
    split_tf32(v[1], h.y, l.y);
    split_tf32(v[2], h.z, l.z);
    split_tf32(v[3], h.w, l.w);
    *reinterpret_cast<float4*>(hi + dst) = h;
    *reinterpret_cast<float4*>(lo + dst) = l;
}
constexpr int kTcPrepZ = 8;   // blocks per (m, parity) of the operand-image kernels

// spectra [m][n][re/im][fld] -> B images of the inverse: B(row = 2 fld + re/im, k = wavenumber index), zero rows /
// zero coefficients as in pack_spectra_kernel (the `jn <= truncation && jm < truncation` rule, TransLocal.cc:982).
// A thread converts four consecutive k of one row: its loads run along the fields (two 8-byte streams per warp, re and im),
// its two stores are 16 bytes each and consecutive rows are 16 bytes apart in the image.
__global__ void pack_spectra_tc_kernel(int T, int nf, int trunc, int n_tot, const int* __restrict__ my_m,
                                       const int* __restrict__ tab_K, const long long* __restrict__ bimg_off,
                                       const double* __restrict__ spec, float* __restrict__ hi, float* __restrict__ lo) {
    const int m = my_m[blockIdx.x], par = blockIdx.y;
    const int K = tab_K[2 * m + par];
    if (K <= 0) return;
    const int chunks = (K + kTcKC - 1) / kTcKC;
    const long long base = bimg_off[2 * m + par] * (static_cast<long long>(n_tot) * kTcKC);  // bimg_off counts 16-wide K chunks
    const long long ioff = static_cast<long long>(2 * trunc + 3 - m) * m / 2 * nf * 2;
    const int total = chunks * (kTcKC / 4) * n_tot;
    for (int e = blockIdx.z * blockDim.x + threadIdx.x; e < total; e += gridDim.z * blockDim.x) {
        const int q = e / n_tot, r = e - q * n_tot;
        double v[4] = {0., 0., 0., 0.};
        if (r < 2 * nf && m < trunc) {
            const int f = r >> 1, imag = r & 1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = 4 * q + i, n = m + par + 2 * k;
                if (n <= trunc && k < K) v[i] = spec[ioff + static_cast<long long>(nf) * (imag + 2 * (n - m)) + f];
            }
        }
        store_split4(v, hi, lo, base + img_quad(n_tot, r, q));
    }
}

// exchange buffer [lat][2 fld + re/im] (fp64) -> B images of the direct transform: B(row = r, k = latitude); four consecutive
// latitudes of one row per thread (loads contiguous along the row of the buffer, 16-byte stores)
__global__ void transpose_fourier_tc_kernel(int nf, int n_tot, const int* __restrict__ my_m, const int* __restrict__ nlat0,
                                            int nleg, const long long* __restrict__ fb_rowoff,
                                            const long long* __restrict__ bimg_off, const double* __restrict__ fb,
                                            float* __restrict__ hi, float* __restrict__ lo) {
    const int m = my_m[blockIdx.x], par = blockIdx.y;
    const int ncol = nleg - nlat0[m];
    if (ncol <= 0) return;
    const int chunks = (ncol + kTcKC - 1) / kTcKC;
    const long long base = bimg_off[2 * m + par] * (static_cast<long long>(n_tot) * kTcKC);
    const double* src = fb + (fb_rowoff[m] + static_cast<long long>(par) * ncol) * (2 * nf);
    const int total = chunks * (kTcKC / 4) * n_tot;
    for (int e = blockIdx.z * blockDim.x + threadIdx.x; e < total; e += gridDim.z * blockDim.x) {
        const int q = e / n_tot, r = e - q * n_tot;
        double v[4] = {0., 0., 0., 0.};
        if (r < 2 * nf) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int jj = 4 * q + i;
                if (jj < ncol) v[i] = src[static_cast<long long>(jj) * (2 * nf) + r];
            }
        }
        store_split4(v, hi, lo, base + img_quad(n_tot, r, q));
    }
}

}  // namespace

struct TcState {
    bool ready = false;
    // A images (Legendre tables), inverse and direct
    float *a_inv_hi = nullptr, *a_inv_lo = nullptr, *a_dir_hi = nullptr, *a_dir_lo = nullptr;
    std::vector<long long> a_inv_off, a_dir_off;   // [2(T+1)] float offsets of the first image of each block
    long long* d_a_inv_off = nullptr;
    long long* d_a_dir_off = nullptr;
    // B image row offsets (units of 16-row K chunks; multiply by n_tot*16 floats)
    std::vector<long long> b_inv_off, b_dir_off;   // [2(T+1)+1]
    long long* d_b_inv_off = nullptr;
    long long* d_b_dir_off = nullptr;
    int* d_tab_K = nullptr;
    long long* d_tab_off = nullptr;
    int* d_tab_pitch = nullptr;
    // per-nf state
    int nf = -1, trunc = -1, dir_trunc = -1;
    float *b_hi = nullptr, *b_lo = nullptr;
    size_t b_cap = 0;
    TcTile* d_tiles_inv = nullptr;
    TcTile* d_tiles_dir = nullptr;
    int n_tiles_inv = 0, n_tiles_dir = 0;
    size_t bytes = 0;
};

static TcState* tc_state(Plan& p) {
    if (!p.tc) p.tc = new TcState();
    return static_cast<TcState*>(p.tc);
}

void tc_free(Plan& p) {
    if (!p.tc) return;
    TcState* s = static_cast<TcState*>(p.tc);
    void* ptrs[] = {s->a_inv_hi, s->a_inv_lo, s->a_dir_hi, s->a_dir_lo, s->d_a_inv_off, s->d_a_dir_off, s->d_b_inv_off,
                    s->d_b_dir_off, s->d_tab_K, s->d_tab_off, s->d_tab_pitch, s->b_hi, s->b_lo, s->d_tiles_inv, s->d_tiles_dir};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    delete s;
    p.tc = nullptr;
}

template <class T>
static int up(T*& d, const std::vector<T>& h, cudaStream_t st) {
    SPT_CUDA(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
    SPT_CUDA(cudaMemcpyAsync(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    return SPTRANS_OK;
}

// Build the split-operand table images once per plan (device-side conversion of the fp64 table).
int tc_prepare_tables(Plan& p) {
    TcState* s = tc_state(p);
    if (s->ready) return SPTRANS_OK;
    const HostGeom& g = p.g;
    const int T = g.T;
    s->a_inv_off.assign(2 * (T + 1), 0);
    s->a_dir_off.assign(2 * (T + 1), 0);
    s->b_inv_off.assign(2 * (T + 1) + 1, 0);
    s->b_dir_off.assign(2 * (T + 1) + 1, 0);
    long long inv_tot = 0, dir_tot = 0;
    for (int m = 0; m <= T; ++m) {
        const int ncol = std::max(0, g.nleg - g.nlat0[m]);
        const bool mine = g.tab_off[2 * m] >= 0;
        for (int par = 0; par < 2; ++par) {
            const int K = g.tab_K[2 * m + par];
            const int i = 2 * m + par;
            s->a_inv_off[i] = inv_tot;
            s->a_dir_off[i] = dir_tot;
            const long long lat_tiles = (ncol + kTcM - 1) / kTcM, k_chunks = (K + kTcKC - 1) / kTcKC;
            const long long k_tiles = (K + kTcM - 1) / kTcM, lat_chunks = (ncol + kTcKC - 1) / kTcKC;
            if (mine && ncol > 0 && K > 0) {
                inv_tot += lat_tiles * k_chunks * (kTcM * kTcKC);
                dir_tot += k_tiles * lat_chunks * (kTcM * kTcKC);
            }
            s->b_inv_off[i + 1] = s->b_inv_off[i] + (mine ? k_chunks : 0);
            s->b_dir_off[i + 1] = s->b_dir_off[i] + ((mine && K > 0) ? lat_chunks : 0);
        }
    }
    SPT_CUDA(cudaMalloc(&s->a_inv_hi, std::max<long long>(inv_tot, 1) * sizeof(float)));
    SPT_CUDA(cudaMalloc(&s->a_inv_lo, std::max<long long>(inv_tot, 1) * sizeof(float)));
    SPT_CUDA(cudaMalloc(&s->a_dir_hi, std::max<long long>(dir_tot, 1) * sizeof(float)));
    SPT_CUDA(cudaMalloc(&s->a_dir_lo, std::max<long long>(dir_tot, 1) * sizeof(float)));
    s->bytes = static_cast<size_t>(inv_tot + dir_tot) * 2 * sizeof(float);
    p.bytes_tables += s->bytes;
    int rc;
    if ((rc = up(s->d_a_inv_off, s->a_inv_off, p.stream))) return rc;
    if ((rc = up(s->d_a_dir_off, s->a_dir_off, p.stream))) return rc;
    if ((rc = up(s->d_b_inv_off, s->b_inv_off, p.stream))) return rc;
    if ((rc = up(s->d_b_dir_off, s->b_dir_off, p.stream))) return rc;
    if ((rc = up(s->d_tab_K, g.tab_K, p.stream))) return rc;
    if ((rc = up(s->d_tab_off, g.tab_off, p.stream))) return rc;
    if ((rc = up(s->d_tab_pitch, g.tab_pitch, p.stream))) return rc;
    const int nm = static_cast<int>(g.my_m.size());
    if (nm > 0) {
        dim3 grid(nm, 2);
        build_a_images_kernel<false><<<grid, 256, 0, p.stream>>>(T, p.d_my_m, p.d_nlat0, g.nleg, s->d_tab_off, s->d_tab_pitch,
                                                                 s->d_tab_K, p.d_tab, s->d_a_inv_off, s->a_inv_hi, s->a_inv_lo);
        build_a_images_kernel<true><<<grid, 256, 0, p.stream>>>(T, p.d_my_m, p.d_nlat0, g.nleg, s->d_tab_off, s->d_tab_pitch,
                                                                s->d_tab_K, p.d_tab, s->d_a_dir_off, s->a_dir_hi, s->a_dir_lo);
        p.launches += 2;
        SPT_CUDA(cudaGetLastError());
    }
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    s->ready = true;
    return SPTRANS_OK;
}

static int n_total_cols(int nf) { return round_up(2 * nf, 16); }

int tc_build_tiles(Plan& p, int nf, int trunc, int dir_trunc) {
    TcState* s = tc_state(p);
    if (s->nf == nf && s->trunc == trunc && s->dir_trunc == dir_trunc) return SPTRANS_OK;
    const HostGeom& g = p.g;
    const int n_tot = n_total_cols(nf);
    if (n_tot > kTcMaxN) {
        set_error("tensor-core Legendre path: at most 256 fields per call (512 TMEM columns)");
        return SPTRANS_ERR_INVALID;
    }
    const int ld = 2 * nf;
    struct Keyed {
        long long cost;
        TcTile t;
    };
    std::vector<Keyed> inv, dir;
    for (int m : g.my_m) {
        const int ncol = g.nleg - g.nlat0[m];
        if (ncol <= 0) continue;
        for (int par = 0; par < 2; ++par) {
            const int Ktab = g.tab_K[2 * m + par];
            const int i = 2 * m + par;
            const long long fb_row0 = g.fb_rowoff[m] + static_cast<long long>(par) * ncol;
            const int k_chunks_tab = (Ktab + kTcKC - 1) / kTcKC;
            const int Kinv = (m < trunc) ? std::min(Ktab, num_n(trunc, m, par)) : 0;
            if (Kinv > 0) {
                const int chunks = (Kinv + kTcKC - 1) / kTcKC;
                for (int l0 = 0, lt = 0; l0 < ncol; l0 += kTcM, ++lt) {
                    TcTile t{};
                    t.a_img = s->a_inv_off[i] + static_cast<long long>(lt) * k_chunks_tab * (kTcM * kTcKC);
                    t.b_img = s->b_inv_off[i] * (static_cast<long long>(n_tot) * kTcKC);
                    t.c_off = (fb_row0 + l0) * ld;
                    t.n_chunks = chunks;
                    t.m_valid = std::min(kTcM, ncol - l0);
                    inv.push_back({static_cast<long long>(chunks), t});
                }
            }
            const int Kdir = std::min(Ktab, num_n(dir_trunc, m, par));
            if (Kdir > 0) {
                const int lat_chunks = (ncol + kTcKC - 1) / kTcKC;
                for (int r0 = 0, rt = 0; r0 < Kdir; r0 += kTcM, ++rt) {
                    TcTile t{};
                    t.a_img = s->a_dir_off[i] + static_cast<long long>(rt) * lat_chunks * (kTcM * kTcKC);
                    t.b_img = s->b_dir_off[i] * (static_cast<long long>(n_tot) * kTcKC);
                    t.c_off = (g.sp_rowoff[i] + r0) * ld;
                    t.n_chunks = lat_chunks;
                    t.m_valid = std::min(kTcM, Kdir - r0);
                    dir.push_back({static_cast<long long>(lat_chunks), t});
                }
            }
        }
    }
    auto finish = [&](std::vector<Keyed>& v, TcTile*& d_ptr, int& count) -> int {
        std::stable_sort(v.begin(), v.end(), [](const Keyed& a, const Keyed& b) { return a.cost > b.cost; });
        std::vector<TcTile> flat(v.size());
        for (size_t k = 0; k < v.size(); ++k) flat[k] = v[k].t;
        if (d_ptr) cudaFree(d_ptr);
        d_ptr = nullptr;
        count = static_cast<int>(flat.size());
        SPT_CUDA(cudaMalloc(&d_ptr, std::max<size_t>(flat.size(), 1) * sizeof(TcTile)));
        SPT_CUDA(cudaMemcpyAsync(d_ptr, flat.data(), flat.size() * sizeof(TcTile), cudaMemcpyHostToDevice, p.stream));
        SPT_CUDA(cudaStreamSynchronize(p.stream));
        return SPTRANS_OK;
    };
    int rc = finish(inv, s->d_tiles_inv, s->n_tiles_inv);
    if (rc) return rc;
    rc = finish(dir, s->d_tiles_dir, s->n_tiles_dir);
    if (rc) return rc;
    // B image buffers (inverse: K chunks of the spectra; direct: latitude chunks of the exchange buffer)
    const size_t need = static_cast<size_t>(std::max(s->b_inv_off.back(), s->b_dir_off.back())) * n_tot * kTcKC;
    if (need > s->b_cap) {
        if (s->b_hi) cudaFree(s->b_hi);
        if (s->b_lo) cudaFree(s->b_lo);
        s->b_hi = s->b_lo = nullptr;
        SPT_CUDA(cudaMalloc(&s->b_hi, std::max<size_t>(need, 4) * sizeof(float)));
        SPT_CUDA(cudaMalloc(&s->b_lo, std::max<size_t>(need, 4) * sizeof(float)));
        s->b_cap = need;
    }
    s->nf = nf;
    s->trunc = trunc;
    s->dir_trunc = dir_trunc;
    return SPTRANS_OK;
}

static int launch_tc_gemm(Plan& p, int nf, const TcTile* tiles, int ntiles, const float* a_hi, const float* a_lo,
                          double* C) {
    if (ntiles == 0) return SPTRANS_OK;
    TcState* s = tc_state(p);
    TcParams prm{};
    prm.tiles = tiles;
    prm.ntiles = ntiles;
    prm.a_hi = a_hi;
    prm.a_lo = a_lo;
    prm.b_hi = s->b_hi;
    prm.b_lo = s->b_lo;
    prm.C = C;
    prm.ldc = 2 * nf;
    prm.n_cols = 2 * nf;
    prm.n_tot = n_total_cols(nf);
    prm.n_inst = (prm.n_tot + 255) / 256;
    prm.n_each = round_up((prm.n_tot + prm.n_inst - 1) / prm.n_inst, 16);
    if (prm.n_each * prm.n_inst != prm.n_tot) {
        // make the instruction split exact by padding the image height (n_tot is already a multiple of 16)
        prm.n_inst = 1;
        while (prm.n_tot / prm.n_inst > 256 || prm.n_tot % (16 * prm.n_inst) != 0) ++prm.n_inst;
        prm.n_each = prm.n_tot / prm.n_inst;
    }
    const size_t stage_bytes = 2ull * kTcM * kTcKC * 4 + 2ull * prm.n_tot * kTcKC * 4;
    // operand ring + the epilogue warps' staging blocks; static shared memory (barriers) comes on top of the 226 KB
    const size_t budget = 226 * 1024 - 128 - kTcEpiBytes;
    prm.n_stages = static_cast<int>(std::min<size_t>(kTcStages, budget / stage_bytes));
    if (prm.n_stages < 2) {
        set_error("tensor-core Legendre path: stage buffers exceed shared memory for this field count");
        return SPTRANS_ERR_INVALID;
    }
    const size_t smem = stage_bytes * prm.n_stages + kTcEpiBytes + 128;
    // the opt-in is per device (context), and one process may hold plans on several devices: remembered per ordinal
    static size_t attr_smem[64] = {};
    if (p.device >= 64 || smem > attr_smem[p.device]) {
        SPT_CUDA(cudaFuncSetAttribute(legendre_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        if (p.device < 64) attr_smem[p.device] = smem;
    }
    const int grid = std::min(ntiles, p.num_sms);
    legendre_tc_kernel<<<grid, kTcThreads, smem, p.stream>>>(prm);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int launch_legendre_inv_tc(Plan& p, int nf, int trunc, const double* d_spec, double* d_fourier, cudaEvent_t after_pack) {
    TcState* s = tc_state(p);
    const int nm = static_cast<int>(p.g.my_m.size());
    if (nm == 0) return SPTRANS_OK;
    dim3 grid(nm, 2, kTcPrepZ);
    pack_spectra_tc_kernel<<<grid, 256, 0, p.stream>>>(p.g.T, nf, trunc, n_total_cols(nf), p.d_my_m, s->d_tab_K, s->d_b_inv_off,
                                                       d_spec, s->b_hi, s->b_lo);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    if (after_pack) cudaEventRecord(after_pack, p.stream);   // stage timing: operand images | tensor-core GEMM
    return launch_tc_gemm(p, nf, s->d_tiles_inv, s->n_tiles_inv, s->a_inv_hi, s->a_inv_lo, d_fourier);
}

int launch_legendre_dir_tc(Plan& p, int nf, const double* d_fourier, double* d_packed, cudaEvent_t after_pack) {
    TcState* s = tc_state(p);
    const int nm = static_cast<int>(p.g.my_m.size());
    if (nm == 0) return SPTRANS_OK;
    dim3 grid(nm, 2, kTcPrepZ);
    transpose_fourier_tc_kernel<<<grid, 256, 0, p.stream>>>(nf, n_total_cols(nf), p.d_my_m, p.d_nlat0, p.g.nleg, p.d_fb_rowoff,
                                                            s->d_b_dir_off, d_fourier, s->b_hi, s->b_lo);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    if (after_pack) cudaEventRecord(after_pack, p.stream);
    return launch_tc_gemm(p, nf, s->d_tiles_dir, s->n_tiles_dir, s->a_dir_hi, s->a_dir_lo, d_packed);
}

}  // namespace sptrans
