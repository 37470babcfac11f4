// Gather / scatter kernels of the one exchange step of the multi-GPU transform: the transposition between
// the zonal-wavenumber sharding of the Legendre stage and the latitude-band sharding of the Fourier stage.
// The reference has no counterpart (TransLocal refuses mpi::size() > 1, ecmwf/atlas
// src/atlas/trans/local/TransLocal.cc:338-340; ectrans does this transposition internally with MPI).
// The collective itself (one all-to-all over NVLink) is issued by the host through NCCL
// (atlas_b200/dist.py); these kernels only pack/unpack contiguous row runs, 16 bytes per thread.
#include "plan.hpp"

namespace sptrans {

namespace {

__global__ void exchange_copy_kernel(const ExSeg* __restrict__ segs, int nf, double2* __restrict__ fb,
                                     double2* __restrict__ buf, int gather) {
    const ExSeg s = segs[blockIdx.x];
    const long long n = static_cast<long long>(s.nrows) * nf;
    double2* a = fb + s.fb_row * nf;
    double2* b = buf + s.buf_row * nf;
    if (gather)
        for (long long i = threadIdx.x; i < n; i += blockDim.x) b[i] = a[i];
    else
        for (long long i = threadIdx.x; i < n; i += blockDim.x) a[i] = b[i];
}

}  // namespace

int launch_exchange_copy(Plan& p, int nf, const ExSeg* d_segs, int nseg, double* d_fourier, double* d_buf,
                         bool gather) {
    if (nseg == 0) return SPTRANS_OK;
    exchange_copy_kernel<<<nseg, 256, 0, p.stream>>>(d_segs, nf, reinterpret_cast<double2*>(d_fourier),
                                                     reinterpret_cast<double2*>(d_buf), gather ? 1 : 0);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

}  // namespace sptrans
