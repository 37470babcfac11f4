// C ABI, single-process multi-GPU plan: one host thread drives N sharded plans (one per device) whose exchange regions
// are mapped into each other by peer access -- the in-process counterpart of the one-process-per-GPU set-up of
// atlas_b200/dist.py, for hosts like the TransB200 adaptor that live in one process (SURVEY 8e "Process model": single
// process, N devices).  No reference equivalent: TransLocal refuses mpi::size() > 1 (trans/local/TransLocal.cc:338-340).
//
// The user hands GLOBAL arrays in the reference's layouts (host memory, or device memory reachable under UVA); every
// device receives only its share (the spectra of its zonal wavenumbers, the grid rows of its latitude band:
// SPTRANS_SHARD_LOCAL_IO), the sharded transforms run stream-ordered on all devices at once -- the device-side flag
// barrier of exchange.cu is their only synchronisation -- and every device writes its share of the result back.
// Everything a call needs is allocated BEFORE anything is enqueued (prepare): a cudaMalloc on a device whose stream
// waits at the barrier for a peer that has not been given its work yet would dead-lock the single host thread.
#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "plan.hpp"

using namespace sptrans;

struct sptrans_multi {
    std::vector<sptrans_plan*> plans;
    std::vector<int> devices;
    int nf = -1;
    bool direct_ready = false;
    std::vector<double*> d_sp, d_gp;
};

namespace {

size_t packed_doubles(const Plan& p, int nf) { return static_cast<size_t>(p.g.sp_rowoff.back()) * 2 * nf; }

int prepare(sptrans_multi& m, int nf, bool direct) {
    const int R = static_cast<int>(m.plans.size());
    if (m.nf == nf && (m.direct_ready || !direct)) return SPTRANS_OK;
    int rc;
    if (m.nf != nf) {
        std::vector<void*> regions(R);
        for (int r = 0; r < R; ++r) {
            if ((rc = sptrans_peer_alloc(m.plans[r], nf, nullptr))) return rc;
            if ((rc = sptrans_peer_region(m.plans[r], &regions[r], nullptr))) return rc;
        }
        for (int r = 0; r < R; ++r)
            if ((rc = sptrans_peer_attach_ptrs(m.plans[r], R, regions.data()))) return rc;
        for (int r = 0; r < R; ++r) {
            Plan& p = m.plans[r]->p;
            SPT_CUDA(cudaSetDevice(p.device));
            if ((rc = build_tiles(p, nf, p.g.T, p.g.T))) return rc;
            if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nf)))) return rc;
            if ((rc = fourier_set_chunks(p, nf, 1, nullptr))) return rc;
            if (m.d_sp[r]) cudaFree(m.d_sp[r]);
            if (m.d_gp[r]) cudaFree(m.d_gp[r]);
            m.d_sp[r] = m.d_gp[r] = nullptr;
            SPT_CUDA(cudaMalloc(&m.d_sp[r], std::max<size_t>(2 * static_cast<size_t>(p.g.spec_ncoef) * nf, 2) * sizeof(double)));
            SPT_CUDA(cudaMalloc(&m.d_gp[r], std::max<size_t>(static_cast<size_t>(p.g.gp_stride) * nf, 2) * sizeof(double)));
        }
        m.nf = nf;
        m.direct_ready = false;
    }
    if (direct && !m.direct_ready) {
        for (int r = 0; r < R; ++r) {
            Plan& p = m.plans[r]->p;
            SPT_CUDA(cudaSetDevice(p.device));
            if (!p.d_weights) {
                set_error("sptrans_multi_dirtrans_scalar: plan was created without quadrature weights");
                return SPTRANS_ERR_INVALID;
            }
            if (!p.d_tabT && (rc = build_transposed_table(p))) return rc;
        }
        m.direct_ready = true;
    }
    auto sync_all = [&]() -> int {
        for (int r = 0; r < R; ++r) {
            SPT_CUDA(cudaSetDevice(m.plans[r]->p.device));
            SPT_CUDA(cudaDeviceSynchronize());
        }
        return SPTRANS_OK;
    };
    if ((rc = sync_all())) return rc;
    // Warm-up, stage by stage with a host synchronisation after each: the first launch of a kernel loads its module
    // (CUDA loads lazily), which may synchronise the device -- harmless here, fatal once a barrier kernel is spinning on
    // it for a peer whose work this very host thread has not enqueued yet.
    for (int r = 0; r < R; ++r) {
        SPT_CUDA(cudaSetDevice(m.plans[r]->p.device));
        SPT_CUDA(cudaMemsetAsync(m.d_sp[r], 0, std::max<size_t>(2 * static_cast<size_t>(m.plans[r]->p.g.spec_ncoef) * nf, 2) * sizeof(double), m.plans[r]->p.stream));
        SPT_CUDA(cudaMemsetAsync(m.d_gp[r], 0, std::max<size_t>(static_cast<size_t>(m.plans[r]->p.g.gp_stride) * nf, 2) * sizeof(double), m.plans[r]->p.stream));
        if ((rc = sptrans_invtrans_legendre_peers(m.plans[r], nf, m.d_sp[r]))) return rc;
    }
    if ((rc = sync_all())) return rc;
    for (int r = 0; r < R; ++r)
        if ((rc = sptrans_peer_barrier(m.plans[r]))) return rc;   // every rank's barrier kernel is enqueued before the host waits
    if ((rc = sync_all())) return rc;
    for (int r = 0; r < R; ++r) {
        Plan& p = m.plans[r]->p;
        SPT_CUDA(cudaSetDevice(p.device));
        if ((rc = launch_fourier_inv(p, nf, p.g.T - 1, make_peer_dst(p).base[p.g.rank], m.d_gp[r], 0))) return rc;
    }
    if ((rc = sync_all())) return rc;
    for (int r = 0; r < R; ++r) sptrans_peer_advance(m.plans[r]);
    if (direct) {
        for (int r = 0; r < R; ++r)
            if ((rc = sptrans_dirtrans_fourier_peers(m.plans[r], nf, m.d_gp[r]))) return rc;
        if ((rc = sync_all())) return rc;
        for (int r = 0; r < R; ++r)
            if ((rc = sptrans_peer_barrier(m.plans[r]))) return rc;
        if ((rc = sync_all())) return rc;
        for (int r = 0; r < R; ++r) {
            Plan& p = m.plans[r]->p;
            SPT_CUDA(cudaSetDevice(p.device));
            if ((rc = launch_legendre_dir(p, nf, make_peer_dst(p).base[p.g.rank], p.d_packed))) return rc;
            if ((rc = launch_unpack_spectra(p, nf, p.d_packed, m.d_sp[r]))) return rc;
        }
        if ((rc = sync_all())) return rc;
        for (int r = 0; r < R; ++r) sptrans_peer_advance(m.plans[r]);
    }
    return SPTRANS_OK;
}

// global [m][n][re/im][field] <-> this rank's [my m][n][re/im][field]
int copy_spectra(Plan& p, int nf, double* d_local, double* global, bool to_device) {
    const int T = p.g.T;
    for (int m : p.g.my_m) {
        const size_t n = static_cast<size_t>(T - m + 1) * 2 * nf;
        double* g = global + static_cast<size_t>(2 * T + 3 - m) * m / 2 * 2 * nf;
        double* l = d_local + static_cast<size_t>(p.g.spec_off[m]) * 2 * nf;
        SPT_CUDA(cudaMemcpyAsync(to_device ? l : g, to_device ? g : l, n * sizeof(double), cudaMemcpyDefault, p.stream));
    }
    return SPTRANS_OK;
}
// global [field][point] <-> this rank's [field][rows of my band]
int copy_grid(Plan& p, int nf, double* d_local, double* global, bool to_device) {
    const HostGeom& g = p.g;
    auto span = [&](int j_first, int j_last) -> int {   // rows j_first..j_last are contiguous in both layouts
        if (j_last < j_first) return SPTRANS_OK;
        const size_t width = static_cast<size_t>(g.rowoff[j_last + 1] - g.rowoff[j_first]) * sizeof(double);
        double* gl = global + g.rowoff[j_first];
        double* lo = d_local + g.gp_rowoff[j_first];
        SPT_CUDA(cudaMemcpy2DAsync(to_device ? lo : gl, (to_device ? static_cast<size_t>(g.gp_stride) : static_cast<size_t>(g.npts)) * sizeof(double),
                                   to_device ? gl : lo, (to_device ? static_cast<size_t>(g.npts) : static_cast<size_t>(g.gp_stride)) * sizeof(double),
                                   width, nf, cudaMemcpyDefault, p.stream));
        return SPTRANS_OK;
    };
    if (g.pair_end <= g.pair_begin) return SPTRANS_OK;
    int rc = span(g.pair_begin, g.pair_end - 1);   // northern rows (and the equator row of a grid with an odd number of rows)
    if (rc) return rc;
    int s_first = g.nlat - g.pair_end, s_last = g.nlat - 1 - g.pair_begin;   // their mirrors, north -> south
    if (s_first <= g.pair_end - 1) s_first = g.pair_end;                      // the equator row is not mirrored
    return span(s_first, s_last);
}

}  // namespace

extern "C" {

int sptrans_multi_create(sptrans_multi** out, int nlat, const int* nx, const double* lat_deg, const double* weights, int truncation,
                         unsigned flags, int ndevices, const int* devices) {
    if (!out || ndevices < 1 || ndevices > kMaxPeers) {
        set_error("sptrans_multi_create: invalid arguments (1..8 devices: one NVLink / NVSwitch node)");
        return SPTRANS_ERR_INVALID;
    }
    *out = nullptr;
    const int ndev = sptrans_device_count();
    if (ndev <= 0) {
        set_error("sptrans_multi_create: no CUDA device visible (this engine has no CPU fallback)");
        return SPTRANS_ERR_CUDA;
    }
    sptrans_multi* m = new (std::nothrow) sptrans_multi();
    if (!m) {
        set_error("out of host memory");
        return SPTRANS_ERR_INVALID;
    }
    m->d_sp.assign(ndevices, nullptr);
    m->d_gp.assign(ndevices, nullptr);
    for (int r = 0; r < ndevices; ++r) m->devices.push_back(devices ? devices[r] : r);
    // peer access between every pair of distinct devices (NVLink / NVSwitch, or PCIe P2P)
    for (int a : m->devices)
        for (int b : m->devices) {
            if (a == b) continue;
            int ok = 0;
            if (a < 0 || a >= ndev || b < 0 || b >= ndev || cudaDeviceCanAccessPeer(&ok, a, b) != cudaSuccess || !ok) {
                set_error("sptrans_multi_create: devices cannot access each other's memory");
                delete m;
                return SPTRANS_ERR_CUDA;
            }
            cudaSetDevice(a);
            const cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                set_error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                delete m;
                return SPTRANS_ERR_CUDA;
            }
            cudaGetLastError();
        }
    for (int r = 0; r < ndevices; ++r) {
        sptrans_plan* p = nullptr;
        const int rc = sptrans_plan_create_sharded(&p, nlat, nx, lat_deg, weights, truncation, flags | SPTRANS_SHARD_LOCAL_IO,
                                                   m->devices[r], r, ndevices);
        if (rc) {
            sptrans_multi_destroy(m);
            return rc;
        }
        m->plans.push_back(p);
    }
    *out = m;
    return SPTRANS_OK;
}

int sptrans_multi_destroy(sptrans_multi* m) {
    if (!m) return SPTRANS_OK;
    for (size_t r = 0; r < m->plans.size(); ++r) {
        cudaSetDevice(m->plans[r]->p.device);
        cudaDeviceSynchronize();
        if (m->d_sp[r]) cudaFree(m->d_sp[r]);
        if (m->d_gp[r]) cudaFree(m->d_gp[r]);
    }
    for (sptrans_plan* p : m->plans) sptrans_plan_destroy(p);
    delete m;
    return SPTRANS_OK;
}

int sptrans_multi_size(const sptrans_multi* m) { return m ? static_cast<int>(m->plans.size()) : 0; }
sptrans_plan* sptrans_multi_plan(sptrans_multi* m, int rank) {
    return (m && rank >= 0 && rank < static_cast<int>(m->plans.size())) ? m->plans[rank] : nullptr;
}

static int finish_all(sptrans_multi* m) {
    for (sptrans_plan* pl : m->plans) {
        SPT_CUDA(cudaSetDevice(pl->p.device));
        SPT_CUDA(cudaStreamSynchronize(pl->p.stream));
        SPT_CUDA(cudaGetLastError());
    }
    return SPTRANS_OK;
}

int sptrans_multi_invtrans_scalar(sptrans_multi* m, int nf, const double* spectra, double* gp) {
    if (!m || nf < 0 || (nf > 0 && (!spectra || !gp))) {
        set_error("sptrans_multi_invtrans_scalar: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (nf == 0) return SPTRANS_OK;
    int rc = prepare(*m, nf, false);
    if (rc) return rc;
    const int R = static_cast<int>(m->plans.size());
    // Enqueue every device's upload + transform first.  Copies from / to PAGEABLE host memory are synchronous with the
    // host (a device-to-host copy returns only when its stream has drained), so the downloads are issued only after every
    // device has been given its work: a device waiting at the barrier for a peer this thread has not served yet would
    // otherwise block the thread for good.
    for (int r = 0; r < R; ++r) {
        Plan& p = m->plans[r]->p;
        SPT_CUDA(cudaSetDevice(p.device));
        if ((rc = copy_spectra(p, nf, m->d_sp[r], const_cast<double*>(spectra), true))) return rc;
        if ((rc = sptrans_invtrans_sharded(m->plans[r], nf, m->d_sp[r], m->d_gp[r]))) return rc;
    }
    for (int r = 0; r < R; ++r) {
        Plan& p = m->plans[r]->p;
        SPT_CUDA(cudaSetDevice(p.device));
        if ((rc = copy_grid(p, nf, m->d_gp[r], gp, false))) return rc;
    }
    return finish_all(m);
}

int sptrans_multi_dirtrans_scalar(sptrans_multi* m, int nf, const double* gp, double* spectra) {
    if (!m || nf < 0 || (nf > 0 && (!spectra || !gp))) {
        set_error("sptrans_multi_dirtrans_scalar: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (nf == 0) return SPTRANS_OK;
    int rc = prepare(*m, nf, true);
    if (rc) return rc;
    const int R = static_cast<int>(m->plans.size());
    for (int r = 0; r < R; ++r) {   // uploads + transforms of every device first, downloads afterwards (see the inverse)
        Plan& p = m->plans[r]->p;
        SPT_CUDA(cudaSetDevice(p.device));
        if ((rc = copy_grid(p, nf, m->d_gp[r], const_cast<double*>(gp), true))) return rc;
        if ((rc = sptrans_dirtrans_sharded(m->plans[r], nf, m->d_gp[r], m->d_sp[r]))) return rc;
    }
    for (int r = 0; r < R; ++r) {
        Plan& p = m->plans[r]->p;
        SPT_CUDA(cudaSetDevice(p.device));
        if ((rc = copy_spectra(p, nf, m->d_sp[r], spectra, false))) return rc;
    }
    return finish_all(m);
}

}  // extern "C"
