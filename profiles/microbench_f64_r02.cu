// Microbenchmark: FP64 DFMA vs DMMA throughput on sm_100a (scratch, not product)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void k_dfma(double* out, int iters) {
    double a[16]; double b = 1.000000001, c = 1e-9;
    #pragma unroll
    for (int i=0;i<16;i++) a[i] = threadIdx.x + i;
    for (int it=0; it<iters; ++it) {
        #pragma unroll
        for (int i=0;i<16;i++) a[i] = fma(a[i], b, c);
    }
    double s=0; 
    #pragma unroll
    for (int i=0;i<16;i++) s+=a[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int NACC>
__global__ void k_dmma884(double* out, int iters) {
    double c0[NACC], c1[NACC];
    #pragma unroll
    for (int i=0;i<NACC;i++){c0[i]=0;c1[i]=0;}
    double a = 1.0 + threadIdx.x*1e-9, b = 1.0 - threadIdx.x*1e-9;
    for (int it=0; it<iters; ++it) {
        #pragma unroll
        for (int i=0;i<NACC;i++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s=0;
    #pragma unroll
    for (int i=0;i<NACC;i++) s+=c0[i]+c1[i];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

template<int NACC>
__global__ void k_dmma16816(double* out, int iters) {
    double c[NACC][4];
    #pragma unroll
    for (int i=0;i<NACC;i++){c[i][0]=c[i][1]=c[i][2]=c[i][3]=0;}
    double a0 = 1.0 + threadIdx.x*1e-9, b0 = 1.0 - threadIdx.x*1e-9;
    for (int it=0; it<iters; ++it) {
        #pragma unroll
        for (int i=0;i<NACC;i++) {
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                : "d"(a0),"d"(a0),"d"(a0),"d"(a0),"d"(a0),"d"(a0),"d"(a0),"d"(a0), "d"(b0),"d"(b0),"d"(b0),"d"(b0));
        }
    }
    double s=0;
    #pragma unroll
    for (int i=0;i<NACC;i++) s+=c[i][0]+c[i][1]+c[i][2]+c[i][3];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// smem-fed dmma: warp tile (8*MI) x (8*NJ), k4 steps from shared memory
template<int MI,int NJ>
__global__ void k_dmma_smem(double* out, int iters, int ksteps) {
    extern __shared__ double sm[];
    // A: [k][lat] tile 4 x (8*MI*warpsM) ; just use per-warp region, size ksteps*4*(8*MI + 8*NJ)
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < ksteps*4*(8*MI+8*NJ); i += blockDim.x) sm[i] = 1.0 + i*1e-9;
    __syncthreads();
    double c0[MI][NJ], c1[MI][NJ];
    #pragma unroll
    for (int i=0;i<MI;i++) for (int j=0;j<NJ;j++){c0[i][j]=0;c1[i][j]=0;}
    for (int it=0; it<iters; ++it) {
        for (int ks=0; ks<ksteps; ++ks) {
            const double* pa = sm + ks*4*(8*MI+8*NJ) + t*(8*MI+8*NJ+0);
            double a[MI], b[NJ];
            #pragma unroll
            for (int i=0;i<MI;i++) a[i] = pa[i*8+g];
            #pragma unroll
            for (int j=0;j<NJ;j++) b[j] = pa[8*MI + j*8+g];
            #pragma unroll
            for (int i=0;i<MI;i++)
              #pragma unroll
              for (int j=0;j<NJ;j++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                    : "+d"(c0[i][j]), "+d"(c1[i][j]) : "d"(a[i]), "d"(b[j]));
        }
    }
    double s=0;
    #pragma unroll
    for (int i=0;i<MI;i++) for (int j=0;j<NJ;j++) s+=c0[i][j]+c1[i][j];
    out[blockIdx.x*blockDim.x+threadIdx.x]=s + warp;
}

// a launch that fails (e.g. 16 warps x 32x56 accumulators do not fit the register file) must not print a rate
#define RATE(label, wps, fl)                                                                            \
    do {                                                                                                \
        cudaError_t le = cudaGetLastError();                                                            \
        if (le != cudaSuccess) printf("%s warps/SM=%2d  launch failed: %s\n", label, wps, cudaGetErrorString(le)); \
        else printf("%s warps/SM=%2d  %.2f TFLOP/s (%.3f ms)\n", label, wps, (fl) / ms * 1e-9, ms);     \
    } while (0)

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
    printf("dev %s sm=%d cc=%d.%d smem/blk optin=%zu clock=%d kHz\n", p.name, p.multiProcessorCount, p.major,p.minor, p.sharedMemPerBlockOptin, p.clockRate);
    double* out; CK(cudaMalloc(&out, 148*16*1024*sizeof(double)));
    cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    int nsm = p.multiProcessorCount;
    for (int wps : {4, 8, 16, 32}) {
        int iters = 20000;
        k_dfma<<<nsm, wps*32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dfma<<<nsm, wps*32>>>(out, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms,e0,e1);
        double fl = 2.0*16*iters*(double)nsm*wps*32;
        RATE("DFMA ", wps, fl);
    }
    for (int wps : {4, 8, 16, 32}) {
        int iters = 20000;
        k_dmma884<8><<<nsm, wps*32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dmma884<8><<<nsm, wps*32>>>(out, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms,e0,e1);
        double fl = 2.0*8*8*4*8*iters*(double)nsm*wps;
        RATE("DMMA884 nacc=8", wps, fl);
    }
    for (int wps : {4, 8, 16}) {
        int iters = 20000;
        k_dmma884<16><<<nsm, wps*32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dmma884<16><<<nsm, wps*32>>>(out, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms,e0,e1);
        double fl = 2.0*8*8*4*16*iters*(double)nsm*wps;
        RATE("DMMA884 nacc=16", wps, fl);
    }
    for (int wps : {4, 8, 16}) {
        int iters = 5000;
        k_dmma16816<8><<<nsm, wps*32>>>(out, 100); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_dmma16816<8><<<nsm, wps*32>>>(out, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms,e0,e1);
        double fl = 2.0*16*8*16*8*iters*(double)nsm*wps;
        RATE("DMMA16816 nacc=8", wps, fl);
    }
    {
        int ksteps = 16; int iters = 2000;
        for (int wps : {4, 8, 12, 16}) {
            size_t sh = ksteps*4*(8*4+8*7)*sizeof(double);
            k_dmma_smem<4,7><<<nsm, wps*32, sh>>>(out, 10, ksteps); cudaDeviceSynchronize();
            cudaEventRecord(e0); k_dmma_smem<4,7><<<nsm, wps*32, sh>>>(out, iters, ksteps); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms,e0,e1);
            double fl = 2.0*256*4*7*ksteps*(double)iters*nsm*wps;
            RATE("DMMA smem-fed 32x56", wps, fl);
        }
        for (int wps : {4, 8, 16}) {
            size_t sh = ksteps*4*(8*4+8*4)*sizeof(double);
            k_dmma_smem<4,4><<<nsm, wps*32, sh>>>(out, 10, ksteps); cudaDeviceSynchronize();
            cudaEventRecord(e0); k_dmma_smem<4,4><<<nsm, wps*32, sh>>>(out, iters, ksteps); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms,e0,e1);
            double fl = 2.0*256*4*4*ksteps*(double)iters*nsm*wps;
            RATE("DMMA smem-fed 32x32", wps, fl);
        }
        for (int wps : {4, 8, 16}) {
            size_t sh = ksteps*4*(8*2+8*4)*sizeof(double);
            k_dmma_smem<2,4><<<nsm, wps*32, sh>>>(out, 10, ksteps); cudaDeviceSynchronize();
            cudaEventRecord(e0); k_dmma_smem<2,4><<<nsm, wps*32, sh>>>(out, iters, ksteps); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms,e0,e1);
            double fl = 2.0*256*2*4*ksteps*(double)iters*nsm*wps;
            RATE("DMMA smem-fed 16x32", wps, fl);
        }
    }
    return 0;
}
