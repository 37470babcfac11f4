// Probe of tcgen05.mma kind::tf32 shared-memory descriptor semantics (SWIZZLE_NONE, K-major) on sm_100a.
// One CTA, M=128, N in {64,144}, K=32.  Tile images are built on the host in the canonical "core matrix"
// layout (8 rows x 16 bytes contiguous) with configurable strides; several descriptor interpretations are
// tried in one run and the one that reproduces the reference GEMM is reported.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // bounded spin: a protocol bug must end in a trap, never in a hung GPU
    for (int it = 0; it < 20000000; ++it) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int version_bit) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    if (version_bit) d |= (uint64_t)1 << 46;
    return d;  // swizzle none, base offset 0
}

struct Params {
    const float* A;   // tile image A (128 x 32)
    const float* B;   // tile image B (N x 32)
    float* D;         // 128 x N row-major
    int N;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // as passed to the descriptor fields (leading, stride)
    uint32_t a_kstep, b_kstep;            // byte advance of the start address per K=8 step
    int version_bit;
    int use_bulk;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_base;
    float* sA = reinterpret_cast<float*>(smem);
    float* sB = reinterpret_cast<float*>(smem + 128 * 32 * 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t a_bytes = 128 * 32 * 4, b_bytes = p.N * 32 * 4;
    if (tid == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = tmem_base;
    if (p.use_bulk) {
        if (tid == 0) {
            mbar_expect_tx(&bar_full, a_bytes + b_bytes);
            bulk_g2s(sA, p.A, a_bytes, &bar_full);
            bulk_g2s(sB, p.B, b_bytes, &bar_full);
        }
    } else {
        for (int i = tid; i < 128 * 32; i += 128) sA[i] = p.A[i];
        for (int i = tid; i < p.N * 32; i += 128) sB[i] = p.B[i];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    if (tid == 0) {
        if (p.use_bulk) mbar_wait(&bar_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((128u >> 4) << 24);
        for (int ks = 0; ks < 4; ++ks) {
            const uint64_t da = make_desc(smem_u32(sA) + ks * p.a_kstep, p.a_lbo, p.a_sbo, p.version_bit);
            const uint64_t db = make_desc(smem_u32(sB) + ks * p.b_kstep, p.b_lbo, p.b_sbo, p.version_bit);
            const uint32_t acc = ks > 0 ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "setp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tbase), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    // everyone waits for the MMAs
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    // epilogue: warp w reads lanes 32w..32w+31
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < p.N; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 16; ++j) p.D[row * p.N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase));
}

// image layout: element (row, k) of an R x 32 operand at byte offset
//   (k/4)*kc_stride + (row/8)*rg_stride + (row%8)*16 + (k%4)*4
static void build_image(const std::vector<float>& M, int R, std::vector<float>& img, uint32_t kc_stride, uint32_t rg_stride) {
    img.assign((size_t)R * 32, 0.f);
    for (int r = 0; r < R; ++r)
        for (int k = 0; k < 32; ++k) {
            size_t off = (size_t)(k / 4) * kc_stride + (size_t)(r / 8) * rg_stride + (r % 8) * 16 + (k % 4) * 4;
            img[off / 4] = M[(size_t)r * 32 + k];
        }
}

int main() {
    int fails = 0;
    for (int N : {64, 144}) {
        std::vector<float> A(128 * 32), B((size_t)N * 32), ref((size_t)128 * N);
        for (int i = 0; i < 128; ++i) for (int k = 0; k < 32; ++k) A[i * 32 + k] = ((i * 7 + k * 3) % 17 - 8) / 8.0f;
        for (int j = 0; j < N; ++j) for (int k = 0; k < 32; ++k) B[j * 32 + k] = ((j * 5 + k * 11) % 13 - 6) / 4.0f;
        for (int i = 0; i < 128; ++i) for (int j = 0; j < N; ++j) { double s = 0; for (int k = 0; k < 32; ++k) s += (double)A[i*32+k] * B[j*32+k]; ref[(size_t)i*N+j] = (float)s; }
        // two physical layouts: L1 = k-chunk major ([kc][rowgroup]), L2 = row-group major ([rowgroup][kc])
        for (int layout = 0; layout < 2; ++layout) {
            uint32_t a_kc, a_rg, b_kc, b_rg;
            if (layout == 0) { a_kc = 16 * 128; a_rg = 128; b_kc = (N / 8) * 128; b_rg = 128; }
            else             { a_kc = 128; a_rg = 8 * 128; b_kc = 128; b_rg = 8 * 128; }
            std::vector<float> imgA, imgB;
            build_image(A, 128, imgA, a_kc, a_rg);
            build_image(B, N, imgB, b_kc, b_rg);
            float *dA, *dB, *dD;
            CK(cudaMalloc(&dA, imgA.size() * 4)); CK(cudaMalloc(&dB, imgB.size() * 4)); CK(cudaMalloc(&dD, ref.size() * 4));
            CK(cudaMemcpy(dA, imgA.data(), imgA.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dB, imgB.data(), imgB.size() * 4, cudaMemcpyHostToDevice));
            for (int swap = 0; swap < 2; ++swap)
              for (int ver = 0; ver < 2; ++ver)
                for (int bulk = 0; bulk < 2; ++bulk) {
                    Params p{};
                    p.A = dA; p.B = dB; p.D = dD; p.N = N;
                    // swap==0: descriptor "leading" field = K-direction stride, "stride" field = row-group stride
                    p.a_lbo = swap ? a_rg : a_kc; p.a_sbo = swap ? a_kc : a_rg;
                    p.b_lbo = swap ? b_rg : b_kc; p.b_sbo = swap ? b_kc : b_rg;
                    p.a_kstep = 2 * a_kc; p.b_kstep = 2 * b_kc;  // K=8 per MMA = two 16-byte chunks
                    p.version_bit = ver; p.use_bulk = bulk;
                    CK(cudaMemset(dD, 0xFF, ref.size() * 4));
                    size_t smem = 128 * 32 * 4 + (size_t)N * 32 * 4 + 1024;
                    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    probe_kernel<<<1, 128, smem>>>(p);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("N=%d layout=%d swap=%d ver=%d bulk=%d: CUDA ERROR %s\n", N, layout, swap, ver, bulk, cudaGetErrorString(e)); return 2; }
                    std::vector<float> out(ref.size());
                    CK(cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost));
                    double err = 0; for (size_t i = 0; i < out.size(); ++i) { double d = fabs((double)out[i] - ref[i]); if (!(d == d)) d = 1e30; if (d > err) err = d; }
                    printf("N=%3d layout=%d swap=%d ver=%d bulk=%d : max err %.3e %s\n", N, layout, swap, ver, bulk, err, err < 1e-4 ? "MATCH" : "");
                    if (err < 1e-4) fails = fails; 
                }
            cudaFree(dA); cudaFree(dB); cudaFree(dD);
        }
    }
    return 0;
}
