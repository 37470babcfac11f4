// fp64 Legendre GEMM kernel, TMA-fed and warp-decoupled variant (the default; legendre_f64.cu keeps the per-thread
// cp.async kernel behind -DSPT_BULK=0 as the baseline it was measured against).
//
// Same tiling as the baseline -- one CTA per SM, 8 warps as 4 x 2, warp tile 32 x 72 = 4 x 9 DMMA m8n8k4 tiles, a ring of
// kStages shared-memory stages of kBK contraction steps -- but
//   * operands are fetched by the TMA engine: per stage 16 + 16 one-dimensional bulk copies (cp.async.bulk.shared.global,
//     SASS UBLKCP), one per operand row (A: 128 doubles of the table, B: <= 144 doubles of spectra / Fourier rows),
//     completing on the stage's `full` mbarrier (expect-tx byte counts).  Every warp issues the copies of its own two A
//     rows and two B rows -- ~20 instructions per warp and stage instead of ~220 per thread with cp.async, and no single
//     producer warp that the others wait for (UBLKCP takes uniform operands: a warp issuing all 32 copies runs a 32-trip
//     serial loop and was measured to hold the block barrier up by ~900 cycles per stage);
//   * stages are handed back split-phase (SPT_SPLIT, default): a warp releases a stage with one arrival on its `empty`
//     mbarrier (count 8) and waits for everybody's release only one stage LATER, just before it refills its share, so
//     nobody waits unless a whole stage ahead.  Measured at TCo1279 L137 with 32-step stages: 15.98 / 15.83 ms against
//     17.0 / 16.86 ms with a __syncthreads() per stage (-DSPT_SPLIT=0).  (A single producer warp refilling for everybody
//     was slower than either: 19.8 ms.  So was dropping the one __syncthreads() per TILE as well -- operand ring and tile
//     descriptors running on across tile boundaries, epilogues not aligned: 15.72 / 16.34 ms against 15.37 / 15.45 ms.)
//   * both directions read K-major A tiles: the direct transform uses a transposed copy of the table (P^T, [lat][k] per
//     (m, parity) block, same block offsets), so its A rows are 1 KB copies as well (128-byte copies of the untransposed
//     table made the TMA issue rate the bottleneck: 35 ms instead of 18 ms).
//
// Rows / columns of a stage that lie outside the operand are never copied: the stale shared-memory values there only
// reach rows / columns of C that are not stored.  The one exception are latitude rows of B past the end of a direct
// tile's block, whose table entries are zero: they are cleared, because 0 x (stale NaN) would poison the sum.
// (included by legendre_f64.cu inside namespace sptrans::<anonymous>, after the tile constants and dmma884)
#pragma once
#ifndef SPT_SPLIT
#define SPT_SPLIT 1
#endif

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
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {  // non-blocking probe
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return;
    // bounded spin (~4 s of SM clock): a protocol bug ends in a trap (reported as a CUDA error), never in a hung device
    const long long t0 = clock64();
    while (!mbar_try(bar, parity))
        if (clock64() - t0 > 8000000000ll) asm volatile("trap;");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int kBulkAStage = kBK * (kBM + 4);  // As[kBK][kBM + 4] for both directions
constexpr int kBulkAPitch = kBM + 4;
constexpr size_t kBulkSmemBytes = static_cast<size_t>(kStages) * (kBulkAStage + kBStage) * sizeof(double) + 16;
static_assert(kBK % (kLegThreads / 32) == 0 && 2 * kBK / (kLegThreads / 32) <= 32, "rows of a stage are split evenly over the warps");

// kDirect == false : A = table P   (tile [kBK n-rows][kBM latitudes]),  B = packed spectra,  C = Fourier buffer
// kDirect == true  : A = table P^T (tile [kBK latitudes][kBM n-rows]),  B = Fourier buffer,  C = packed spectra
template <bool kDirect, bool kPeers>
__global__ void __launch_bounds__(kLegThreads, 1)
legendre_dmma_kernel(const LegTile* __restrict__ tiles, int ntiles, int* __restrict__ counter,
                     const double* __restrict__ tab, const double* __restrict__ B, double* __restrict__ C, int ldb,
                     const __grid_constant__ PeerDst dst) {
    extern __shared__ __align__(16) double smem[];
    double* As = smem;
    double* Bs = smem + kStages * kBulkAStage;
    __shared__ LegTile s_tl[2];   // current / next tile descriptor (fetched one tile ahead by thread 0)
    __shared__ int s_ti[2];
    __shared__ __align__(8) uint64_t s_full[kStages];   // the bulk copies of stage s have landed
#if SPT_SPLIT
    __shared__ __align__(8) uint64_t s_empty[kStages];  // all eight warps are done reading stage s
    uint32_t empty_phase = (1u << kStages) - 1u;        // the first wait on a fresh `empty` barrier passes (parity 1)
#endif

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    // warp -> (row block, column block): consecutive warps (= the four SM sub-partitions) cover the first two row
    // blocks, so that a partial tile with <= 64 valid rows still keeps every sub-partition's FP64 pipe busy
    const int wm = warp / kWarpsN, wn = warp % kWarpsN;
    uint32_t full_phase = 0;                        // bit s: parity of the completion of s_full[s] to wait for next

    // Every warp fetches its own share of a stage: rows [rpw * warp, rpw * (warp + 1)) of A and of B, one bulk copy per
    // lane (lanes 0..rpw-1: A rows, lanes rpw..2 rpw-1: B rows), after ONE arrive.expect_tx for the warp's bytes
    // (`full` counts the eight warps).  The stage must have been released by all warps (block barrier).
    constexpr int rpw = kBK / (kLegThreads / 32);
    auto load_stage = [&](const LegTile& tl, int kb, int st) {
        const double* Ag = tab + tl.a_off;
        const double* Bg = (kDirect && kPeers) ? nullptr : B + tl.b_off;
        double* as = As + st * kBulkAStage;
        double* bs = Bs + st * kBStage;
        uint64_t* bar = &s_full[st];
        const uint32_t a_bytes = static_cast<uint32_t>(min(kBM, tl.a_rows)) * 8u;  // readable part of an A row
        const uint32_t b_bytes = static_cast<uint32_t>(tl.n_valid) * 8u;
        const int row0 = rpw * warp;                                              // first row of this warp's share
        const int b_tot = kDirect ? max(0, min(kBK, tl.b_rows - kb * kBK)) : kBK;  // B rows inside the block
        const int b_n = max(0, min(rpw, b_tot - row0));                            // ... of this warp's share
        if (kDirect && lane >= rpw + b_n && lane < 2 * rpw)  // latitude past the block: table entries are zero, B must be finite
            for (int c = 0; c < kBN; c += 2)
                *reinterpret_cast<double2*>(bs + (row0 + lane - rpw) * kBPitch + c) = make_double2(0., 0.);
        __syncwarp();
        if (lane == 0) mbar_expect_tx(bar, rpw * a_bytes + static_cast<uint32_t>(b_n) * b_bytes);
        __syncwarp();
        if (lane < rpw)
            bulk_g2s(as + (row0 + lane) * kBulkAPitch, Ag + static_cast<long long>(kb * kBK + row0 + lane) * tl.a_pitch, a_bytes, bar);
        else if (lane < rpw + b_n) {
            const int r = kb * kBK + row0 + lane - rpw;   // row of B = latitude pair tl.lat0 + r (direct) / n-row (inverse)
            const double* src;
            if (kDirect && kPeers) {
                // fused exchange of the direct transform: the Fourier rows of a latitude pair live in the exchange buffer of
                // the rank that owns its band (same layout on every rank) and are PULLED from there by this TMA copy -- over
                // NVLink for the peers' bands -- instead of being pushed by the Fourier stage
                const int lat = tl.lat0 + r;
                int d = 0;
                while (d + 1 < dst.nranks && lat >= dst.band[d + 1]) ++d;
                src = dst.base[d] + tl.b_off + static_cast<long long>(r) * ldb;
            }
            else src = Bg + static_cast<long long>(r) * ldb;
            bulk_g2s(bs + (row0 + lane - rpw) * kBPitch, src, b_bytes, bar);
        }
    };
    // first kStages-1 stages of a tile; every stage of the ring is free at this point (tile boundary)
    auto prologue = [&](const LegTile& tl) {
#pragma unroll
        for (int s = 0; s < kStages - 1; ++s)
            if (s < tl.k_steps) {
#if SPT_SPLIT
                mbar_wait(&s_empty[s], (empty_phase >> s) & 1u);
                empty_phase ^= 1u << s;
#endif
                load_stage(tl, s, s);
            }
    };

    // thread 0: publish the claimed tile index and start copying its descriptor (4 x 16 bytes) into s_tl[slot]
    auto fetch_next = [&](int ti, int slot) {
        s_ti[slot] = ti;
        if (ti < ntiles) {
            const char* src = reinterpret_cast<const char*>(tiles + ti);
            const uint32_t dstaddr = smem_u32(&s_tl[slot]);
            static_assert(sizeof(LegTile) % 16 == 0, "descriptor is copied in 16-byte pieces");
#pragma unroll
            for (int c = 0; c < static_cast<int>(sizeof(LegTile)); c += 16)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(dstaddr + c), "l"(src + c) : "memory");
        }
    };

    if (tid == 0) {
        for (int st = 0; st < kStages; ++st) {
            mbar_init(&s_full[st], kLegThreads / 32);
#if SPT_SPLIT
            mbar_init(&s_empty[st], kLegThreads / 32);
#endif
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int ti = atomicAdd(counter, 1);
        s_ti[0] = ti;
        if (ti < ntiles) s_tl[0] = tiles[ti];
    }
    __syncthreads();
    if (s_ti[0] >= ntiles) return;
    prologue(s_tl[0]);

    for (int buf = 0;; buf ^= 1) {
        const LegTile tl = s_tl[buf];
        // Claim the next tile now, but do not wait for the atomic: its result is first touched after the first stage of
        // this tile, when thread 0 starts an asynchronous copy of the descriptor into shared memory (awaited at the end of
        // the tile).  Done synchronously, warp 0 sat out ~2 us of global-memory latency at the start of every tile and
        // the other seven warps waited for it at the first stage barrier.
        int claimed = 0;
        if (tid == 0) claimed = atomicAdd(counter, 1);

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
            const int st = kb % kStages;
            mbar_wait(&s_full[st], (full_phase >> st) & 1u);
            full_phase ^= 1u << st;
#if !SPT_SPLIT
            __syncthreads();  // everybody is done with stage kb-1: its buffer is refilled with block kb + kStages - 1
            {
                const int nk = kb + kStages - 1;
                if (nk < ksteps) load_stage(tl, nk, nk % kStages);
            }
#endif
            if (tid == 0 && kb == 1) fetch_next(claimed, buf ^ 1);
            if (warp_active) {
                const double* as = As + st * kBulkAStage;
                const double* bs = Bs + st * kBStage;
                // fragments are double buffered: the shared-memory loads of k-step ks+1 are issued before the DMMAs of step ks
                double a[2][kMI], b[2][kNJ];
                auto load_frag = [&](int ks, int fb) {
#pragma unroll
                    for (int i = 0; i < kMI; ++i) a[fb][i] = as[(ks * 4 + t) * kBulkAPitch + row_w + 8 * i + g];
#pragma unroll
                    for (int j = 0; j < kNJ; ++j) b[fb][j] = bs[(ks * 4 + t) * kBPitch + col_w + 8 * j + g];
                };
                // the last stage of a tile holds operand rows in its first ks_last k-steps only
                const int ks_n = (kb == ksteps - 1) ? tl.ks_last : kBK / 4;
                load_frag(0, 0);
#pragma unroll
                for (int ks = 0; ks < kBK / 4; ++ks) {
                    if (ks >= ks_n) break;
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
#if SPT_SPLIT
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[st]);
            const int nk = kb + kStages - 1;
            if (nk < ksteps) {   // stage released by this warp one step ago; refilled once everybody has released it
                const int ns = nk % kStages;
                mbar_wait(&s_empty[ns], (empty_phase >> ns) & 1u);
                empty_phase ^= 1u << ns;
                load_stage(tl, nk, ns);
            }
#endif
        }
        if (tid == 0) {
            if (ksteps < 2) fetch_next(claimed, buf ^ 1);
            asm volatile("cp.async.wait_all;\n" ::: "memory");
        }
        __syncthreads();  // every warp is through the tile: the ring is free; the next descriptor written by thread 0 is visible
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
                    if (!kPeers || kDirect) Cg = C + tl.c_off;
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
    if (kPeers && !kDirect) __threadfence_system();  // remote rows are visible to the peers before this kernel completes
}
