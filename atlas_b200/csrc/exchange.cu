// Gather / scatter kernels of the one exchange step of the multi-GPU transform: the transposition between
// the zonal-wavenumber sharding of the Legendre stage and the latitude-band sharding of the Fourier stage.
// The reference has no counterpart (TransLocal refuses mpi::size() > 1, ecmwf/atlas
// src/atlas/trans/local/TransLocal.cc:338-340; ectrans does this transposition internally with MPI).
// Two implementations:
//  * peer memory (default on an NVLink/NVSwitch node): the exchange buffers of all ranks are mapped into every
//    process; the inverse Legendre kernel stores its rows directly into the consumer's buffer
//    (legendre_f64.cu, kPeers), the direct transform pushes its rows with exchange_push_kernel, and
//    peer_barrier_kernel is the only synchronisation (flag exchange over NVLink, no host involvement);
//  * NCCL: pack -> all_to_all_single (issued by the host, atlas_b200/dist.py) -> unpack.
#include <cstdio>

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

// Direct transform: rows of this rank's latitude band go to the buffers of the ranks that own their zonal
// wavenumber.  Segments are contiguous runs of rows (nrows * nf double2); blockIdx.y splits long runs.
__global__ void exchange_push_kernel(const ExSeg* __restrict__ segs, int nf, const double2* __restrict__ fb,
                                     const __grid_constant__ PeerDst dst, int me) {
    const ExSeg s = segs[blockIdx.x];
    if (s.peer == me) return;  // already in place: producer and consumer buffer are the same
    const long long n = static_cast<long long>(s.nrows) * nf;
    const long long per = (n + gridDim.y - 1) / gridDim.y;
    const long long lo = per * blockIdx.y, hi = min(n, lo + per);
    const double2* src = fb + s.fb_row * nf;
    double2* out = reinterpret_cast<double2*>(dst.base[s.peer]) + s.fb_row * nf;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) out[i] = src[i];
    __threadfence_system();
}

struct PeerFlags {
    unsigned long long* flags[kMaxPeers];
};

// All-ranks barrier on the stream: rank `me` publishes `epoch` in slot `me` of every peer's flag array and waits
// until all slots of its own array have reached it.  Kernels launched earlier on this stream have completed (and
// fenced their peer stores) before this one starts, so passing the barrier means every producer is done.
// A peer that never arrives trips the ~10 s watchdog and the kernel traps instead of hanging the GPU.
__global__ void peer_barrier_kernel(const __grid_constant__ PeerFlags pf, int me, int nranks, unsigned long long epoch) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
    __threadfence_system();
    unsigned long long* remote = pf.flags[r] + me;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(epoch) : "memory");
    const unsigned long long* mine = pf.flags[me] + r;
    const long long t0 = clock64();
    for (;;) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        if (clock64() - t0 > 20000000000LL) {
            printf("sptrans: peer barrier timeout (rank %d waiting for rank %d, epoch %llu, seen %llu)\n", me, r, epoch, v);
            __trap();
        }
    }
}

}  // namespace

static double* peer_buffer(const PeerState& ps, int r, int parity) {
    return reinterpret_cast<double*>(static_cast<char*>(ps.peer_region[r]) + kPeerFlagBytes) + static_cast<size_t>(parity) * ps.buf_doubles;
}

PeerDst make_peer_dst(const Plan& p) {
    PeerDst d{};
    d.nranks = p.g.nranks;
    for (int r = 0; r < p.g.nranks; ++r) {
        d.base[r] = peer_buffer(p.peer, r, p.peer.parity);
        d.band[r] = p.g.band[r];
    }
    d.band[p.g.nranks] = p.g.band[p.g.nranks];
    return d;
}

int launch_exchange_push(Plan& p, int nf) {
    const int nseg = static_cast<int>(p.ex.band_side.size());
    if (nseg == 0) return SPTRANS_OK;
    const PeerDst dst = make_peer_dst(p);
    dim3 grid(nseg, 4);
    exchange_push_kernel<<<grid, 256, 0, p.stream>>>(p.d_ex_band, nf, reinterpret_cast<const double2*>(dst.base[p.g.rank]), dst,
                                                     p.g.rank);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int launch_peer_barrier(Plan& p) {
    PeerFlags pf{};
    for (int r = 0; r < p.g.nranks; ++r) pf.flags[r] = static_cast<unsigned long long*>(p.peer.peer_region[r]);
    p.peer.epoch++;
    peer_barrier_kernel<<<1, 32, 0, p.stream>>>(pf, p.g.rank, p.g.nranks, p.peer.epoch);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

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
