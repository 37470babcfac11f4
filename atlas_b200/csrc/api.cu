// C ABI of the B200 spectral-transform engine (include/sptrans_b200.h).
// Orchestration only: plan life cycle, workspaces, host<->device staging, stage sequencing and timing.
// There is no CPU path: without a CUDA device every entry point returns SPTRANS_ERR_CUDA.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <new>

#include "plan.hpp"

using namespace sptrans;

namespace sptrans {

bool is_device_pointer(const void* ptr) {
    if (!ptr) return false;
    cudaPointerAttributes at{};
    cudaError_t e = cudaPointerGetAttributes(&at, ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int ensure(double*& buf, size_t& cap, size_t need_doubles) {
    if (need_doubles <= cap && buf) return SPTRANS_OK;
    if (buf) cudaFree(buf);
    buf = nullptr;
    cap = 0;
    SPT_CUDA(cudaMalloc(&buf, std::max<size_t>(need_doubles, 2) * sizeof(double)));
    cap = need_doubles;
    return SPTRANS_OK;
}

}  // namespace sptrans

namespace {

template <class T>
int upload(T*& dptr, const std::vector<T>& h, cudaStream_t s) {
    SPT_CUDA(cudaMalloc(&dptr, std::max<size_t>(h.size(), 1) * sizeof(T)));
    SPT_CUDA(cudaMemcpyAsync(dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return SPTRANS_OK;
}

size_t packed_doubles(const Plan& p, int nf) { return static_cast<size_t>(p.g.sp_rowoff.back()) * 2 * nf; }
size_t fourier_doubles(const Plan& p, int nf) {
    // + one tile of slack rows so that masked tile loads never form addresses past the allocation
    return (static_cast<size_t>(p.g.fb_rowoff.back()) + kBM) * 2 * nf;
}
size_t spec_doubles(const Plan& p, int nf, int trunc) { return static_cast<size_t>(trunc + 1) * (trunc + 2) * nf; }

struct StageTimer {
    Plan& p;
    explicit StageTimer(Plan& pl): p(pl) {
        for (float& t : p.t_ms) t = 0.f;
    }
    void mark(int i) { cudaEventRecord(p.ev[i], p.stream); }
    void finish(int n_marks, const int* slot) {
        if (p.async) {  // sptrans_set_async: return after enqueueing; sptrans_last_timings reads the events lazily
            p.pending_marks = n_marks;
            for (int i = 0; i + 1 < n_marks; ++i) p.pending_slots[i] = slot[i];
            return;
        }
        cudaStreamSynchronize(p.stream);
        if (p.s_d2h) cudaStreamSynchronize(p.s_d2h);   // chunked copies of the grid fields back to the host
        p.d2h_pending = false;
        for (int i = 0; i + 1 < n_marks; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]);
            p.t_ms[slot[i]] += ms;
        }
    }
};

// ---- host-pointer pipelines ------------------------------------------------------------------------------------
// Fields are independent through the Fourier stage, and the grid-point arrays are [field][point]: the device<->host
// copy of one chunk of fields runs on a copy stream while the Fourier kernels work on the next chunk (the reference's
// contract -- results visible on return, TransLocal.cc:1523-1597 -- holds: the call synchronises at its end unless
// sptrans_set_async is on).  With an inverse and a direct transform in flight on two plans (sptrans_plan_clone) both
// directions of the PCIe link are busy at the same time.
constexpr int kMaxHostChunks = 16;
constexpr int kEvD2HDone = 32, kEvGpConsumed = 33;   // slots of Plan::ev_chunk beyond the per-chunk hand-over events
int host_chunks(int nf) {
    static int v = [] {
        const char* e = std::getenv("SPTRANS_HOST_CHUNKS");
        const int x = e ? std::atoi(e) : 6;
        return std::max(1, std::min(kMaxHostChunks, x));
    }();
    return std::max(1, std::min(v, nf));
}
int copy_streams(Plan& p) {
    if (p.s_h2d) return SPTRANS_OK;
    SPT_CUDA(cudaStreamCreateWithFlags(&p.s_h2d, cudaStreamNonBlocking));
    SPT_CUDA(cudaStreamCreateWithFlags(&p.s_d2h, cudaStreamNonBlocking));
    for (auto& e : p.ev_chunk) SPT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return SPTRANS_OK;
}
// Fourier-inverse stage of `nf` fields, chunk by chunk, every finished chunk copied to the host array on the D2H stream
int fourier_inv_to_host(Plan& p, int nf, int mlimit, const double* d_fourier, double* d_gp, int nb_uv, double* h_gp) {
    int rc = copy_streams(p);
    if (rc) return rc;
    std::vector<int> fb;
    if ((rc = fourier_set_chunks(p, nf, host_chunks(nf), &fb))) return rc;
    const int nc = static_cast<int>(fb.size()) - 1;
    const size_t stride = static_cast<size_t>(p.g.points ? p.g.npts : p.g.gp_stride);
    // the staging buffer may still be being copied out by the previous call on this plan (asynchronous mode): only the
    // Fourier kernels wait for that, the spectra upload and the Legendre stage of this call have already overlapped it
    if (p.d2h_pending) SPT_CUDA(cudaStreamWaitEvent(p.stream, p.ev_chunk[kEvD2HDone], 0));
    for (int c = 0; c < nc; ++c) {
        if ((rc = launch_fourier_inv(p, nf, mlimit, d_fourier, d_gp, nb_uv, nullptr, c))) return rc;
        SPT_CUDA(cudaEventRecord(p.ev_chunk[c], p.stream));
        SPT_CUDA(cudaStreamWaitEvent(p.s_d2h, p.ev_chunk[c], 0));
        SPT_CUDA(cudaMemcpyAsync(h_gp + fb[c] * stride, d_gp + fb[c] * stride, (fb[c + 1] - fb[c]) * stride * sizeof(double),
                                 cudaMemcpyDeviceToHost, p.s_d2h));
    }
    SPT_CUDA(cudaEventRecord(p.ev_chunk[kEvD2HDone], p.s_d2h));
    p.d2h_pending = true;   // the call is complete when this event is (StageTimer::finish / sptrans_synchronize wait for it)
    return SPTRANS_OK;
}
// Fourier-direct stage of `nf` fields read from a host array: chunk c + 1 crosses the bus while chunk c is transformed
int fourier_dir_from_host(Plan& p, int nf, const double* h_gp, double* d_gp, double* d_fourier, int nb_uv, int adjoint) {
    int rc = copy_streams(p);
    if (rc) return rc;
    std::vector<int> fb;
    if ((rc = fourier_set_chunks(p, nf, host_chunks(nf), &fb))) return rc;
    const int nc = static_cast<int>(fb.size()) - 1;
    const size_t stride = static_cast<size_t>(p.g.gp_stride);
    // the staging buffer may still be read by the Fourier kernels of the previous call on this plan (asynchronous mode)
    if (p.gp_in_use) SPT_CUDA(cudaStreamWaitEvent(p.s_h2d, p.ev_chunk[kEvGpConsumed], 0));
    for (int c = 0; c < nc; ++c) {
        SPT_CUDA(cudaMemcpyAsync(d_gp + fb[c] * stride, h_gp + fb[c] * stride, (fb[c + 1] - fb[c]) * stride * sizeof(double),
                                 cudaMemcpyHostToDevice, p.s_h2d));
        SPT_CUDA(cudaEventRecord(p.ev_chunk[c], p.s_h2d));
    }
    for (int c = 0; c < nc; ++c) {
        SPT_CUDA(cudaStreamWaitEvent(p.stream, p.ev_chunk[c], 0));
        if ((rc = launch_fourier_dir(p, nf, d_gp, d_fourier, nb_uv, adjoint, c))) return rc;
    }
    SPT_CUDA(cudaEventRecord(p.ev_chunk[kEvGpConsumed], p.stream));
    p.gp_in_use = true;
    return SPTRANS_OK;
}

void peer_release(Plan& p) {
    PeerState& ps = p.peer;
    for (int r = 0; r < kMaxPeers; ++r) {
        if (ps.ipc_opened[r] && ps.peer_region[r]) cudaIpcCloseMemHandle(ps.peer_region[r]);
        ps.ipc_opened[r] = false;
        ps.peer_region[r] = nullptr;
    }
    if (ps.region) cudaFree(ps.region);
    ps = PeerState{};
}

int check_plan(sptrans_plan* plan) {
    if (!plan) {
        set_error("null plan");
        return SPTRANS_ERR_INVALID;
    }
    SPT_CUDA(cudaSetDevice(plan->p.device));
    return SPTRANS_OK;
}

// inverse transform of `nf` fields whose spectra sit on the device at truncation `trunc` (T or T+1)
int run_inverse(Plan& p, int nf, int trunc, const double* d_spec, double* d_gp, int nb_uv, StageTimer& tm,
                int& marks, int* slots, double* h_gp = nullptr) {
    // Point sets sum EVERY zonal wavenumber up to the truncation of the data (TransLocal.cc:1331,:1347-1362), grids drop
    // the m == truncation column of a scalar transform (:982): the tiles then run to T+1 over zero rows
    const bool keep_mT = p.g.points && trunc == p.g.T;
    int rc = build_tiles(p, nf, keep_mT ? trunc + 1 : trunc, p.g.T);
    if (rc) return rc;
    rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nf));
    if (rc) return rc;
    rc = ensure(p.d_fourier, p.fourier_cap, fourier_doubles(p, nf));
    if (rc) return rc;
    if (p.precision == SPTRANS_PREC_TC_SPLIT) {
        if (p.g.points) {
            set_error("point-set plans run the fp64 Legendre kernel only");
            return SPTRANS_ERR_NOT_IMPLEMENTED;
        }
        if ((rc = tc_prepare_tables(p))) return rc;
        if ((rc = tc_build_tiles(p, nf, trunc, p.g.T))) return rc;
        tm.mark(marks);
        rc = launch_legendre_inv_tc(p, nf, trunc, d_spec, p.d_fourier, p.ev[marks + 1]);  // records the next mark after packing
        if (rc) return rc;
        slots[marks++] = 0;   // operand images (split-tf32 spectra)
        slots[marks++] = 1;   // tcgen05 GEMM
    }
    else {
        tm.mark(marks);
        rc = launch_pack_spectra(p, nf, trunc, d_spec, p.d_packed, keep_mT ? kPackKeepMT : 0);
        if (rc) return rc;
        slots[marks++] = 0;
        tm.mark(marks);
        rc = launch_legendre_inv(p, nf, p.d_packed, p.d_fourier);
        if (rc) return rc;
        slots[marks++] = 1;
    }
    tm.mark(marks);
    if (p.g.points) rc = launch_points_inv(p, nf, std::min(p.g.T, trunc), p.d_fourier, d_gp, nb_uv);
    else if (p.g.cropped) {   // global rows of the band into the work array, then the copy-out of the crop (:1180-1187)
        if ((rc = ensure(p.d_band, p.band_cap, static_cast<size_t>(p.g.gp_stride) * nf))) return rc;
        if ((rc = launch_fourier_inv(p, nf, std::min(p.g.T, trunc - 1), p.d_fourier, p.d_band, nb_uv))) return rc;
        rc = launch_crop_gather(p, nf, p.d_band, d_gp);
    }
    else if (h_gp) rc = fourier_inv_to_host(p, nf, std::min(p.g.T, trunc - 1), p.d_fourier, d_gp, nb_uv, h_gp);
    else rc = launch_fourier_inv(p, nf, std::min(p.g.T, trunc - 1), p.d_fourier, d_gp, nb_uv);
    if (rc) return rc;
    slots[marks++] = 2;
    tm.mark(marks);
    return SPTRANS_OK;
}

}  // namespace

extern "C" {

const char* sptrans_last_error(void) { return last_error_cstr(); }

int sptrans_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int sptrans_gaussian_latitudes(int N, double* lat_deg_2N, double* weights_2N) {
    if (N < 1 || !lat_deg_2N || !weights_2N) {
        set_error("sptrans_gaussian_latitudes: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    gaussian_quadrature(N, lat_deg_2N, weights_2N);
    return SPTRANS_OK;
}

int sptrans_octahedral_nx(int N, int* nx_2N) {
    if (N < 1 || !nx_2N) {
        set_error("sptrans_octahedral_nx: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    for (int j = 0; j < N; ++j) nx_2N[j] = nx_2N[2 * N - 1 - j] = 20 + 4 * j;
    return SPTRANS_OK;
}

int sptrans_fourier_truncation(int truncation, int nx, int nxmax, int ndgl, double lat_rad, int fullgrid) {
    return fourier_truncation(truncation, nx, nxmax, ndgl, lat_rad, fullgrid != 0);
}

namespace {
struct PointSet {   // sptrans_plan_create_points: per point, its row among the distinct |latitudes| etc. (see HostGeom)
    std::vector<int> row;
    std::vector<double> sign, lon, coslatinv;
};
struct Crop {       // sptrans_plan_create_cropped
    int jlat_min, nlat;
    const int *nx, *jlon_min;
};
}  // namespace

static int create_plan(sptrans_plan** out, int nlat, const int* nx, const double* lat_deg, const double* weights,
                       int truncation, unsigned flags, int device, int rank, int nranks, const PointSet* pts,
                       const Crop* crop = nullptr) {
    if (!out) {
        set_error("sptrans_plan_create: null output pointer");
        return SPTRANS_ERR_INVALID;
    }
    *out = nullptr;
    int ndev = sptrans_device_count();
    if (ndev <= 0) {
        set_error("sptrans_plan_create: no CUDA device visible (this engine has no CPU fallback)");
        return SPTRANS_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        set_error("sptrans_plan_create: bad device ordinal");
        return SPTRANS_ERR_INVALID;
    }
    sptrans_plan* sp = new (std::nothrow) sptrans_plan();
    if (!sp) {
        set_error("out of host memory");
        return SPTRANS_ERR_INVALID;
    }
    Plan& p = sp->p;
    p.device = device;
    p.flags = flags;
    int rc = build_geometry(p.g, nlat, nx, lat_deg, weights, truncation, (flags & SPTRANS_GRID_REGULAR) != 0, rank,
                            nranks);
    if (rc) {
        delete sp;
        return rc;
    }
    if (crop) {
        HostGeom& g = p.g;
        if (nranks != 1 || crop->nlat < 1 || crop->jlat_min < 0 || crop->jlat_min + crop->nlat > nlat || !crop->nx || !crop->jlon_min) {
            set_error("sptrans_plan_create_cropped: invalid arguments");
            delete sp;
            return SPTRANS_ERR_INVALID;
        }
        g.cropped = true;
        g.crop_jlat_min = crop->jlat_min;
        g.crop_nx.assign(crop->nx, crop->nx + crop->nlat);
        g.crop_jlon_min.assign(crop->jlon_min, crop->jlon_min + crop->nlat);
        int pb = g.nleg, pe = 0;
        g.crop_npts = 0;
        for (int r = 0; r < crop->nlat; ++r) {
            const int jg = crop->jlat_min + r;
            if (crop->nx[r] < 1 || crop->nx[r] > nx[jg] || crop->jlon_min[r] < 0 || crop->jlon_min[r] >= nx[jg]) {
                set_error("sptrans_plan_create_cropped: a crop row needs 1 <= nx <= nx(global row) and 0 <= jlon_min < nx(global row)");
                delete sp;
                return SPTRANS_ERR_INVALID;
            }
            const int pair = jg < g.nleg ? jg : nlat - 1 - jg;
            pb = std::min(pb, pair);
            pe = std::max(pe, pair + 1);
            g.crop_npts += crop->nx[r];
        }
        g.pair_begin = pb;   // the Fourier stage runs on the latitude pairs that hold the crop's rows only
        g.pair_end = pe;
    }
    set_io_layout(p.g, (flags & SPTRANS_SHARD_LOCAL_IO) != 0);
    auto fail = [&](int code) {
        sptrans_plan_destroy(sp);
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("cudaSetDevice failed");
        return fail(SPTRANS_ERR_CUDA);
    }
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) p.num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        return fail(SPTRANS_ERR_CUDA);
    }
    p.own_stream = true;
    for (auto& e : p.ev) cudaEventCreate(&e);
    HostGeom& g = p.g;
    if (pts) {
        g.points = true;
        g.pt_row = pts->row;
        g.pt_sign = pts->sign;
        g.pt_lon = pts->lon;
        g.pt_coslatinv = pts->coslatinv;
        g.npts = static_cast<long long>(pts->row.size());  // grid-point arrays of this plan are [field][point]
        if ((rc = upload(p.d_pt_row, g.pt_row, p.stream))) return fail(rc);
        if ((rc = upload(p.d_pt_sign, g.pt_sign, p.stream))) return fail(rc);
        if ((rc = upload(p.d_pt_lon, g.pt_lon, p.stream))) return fail(rc);
        if ((rc = upload(p.d_pt_coslatinv, g.pt_coslatinv, p.stream))) return fail(rc);
    }
    if ((rc = upload(p.d_nlat0, g.nlat0, p.stream))) return fail(rc);
    if ((rc = upload(p.d_fb_rowoff, g.fb_rowoff, p.stream))) return fail(rc);
    if ((rc = upload(p.d_sp_rowoff, g.sp_rowoff, p.stream))) return fail(rc);
    if ((rc = upload(p.d_rowoff, g.rowoff, p.stream))) return fail(rc);
    if ((rc = upload(p.d_nx, g.nx, p.stream))) return fail(rc);
    if ((rc = upload(p.d_my_m, g.my_m, p.stream))) return fail(rc);
    if ((rc = upload(p.d_owner, g.owner, p.stream))) return fail(rc);
    if (g.local_io && (rc = upload(p.d_spec_off, g.spec_off, p.stream))) return fail(rc);
    if ((rc = upload(p.d_pair_done, std::vector<int>(3 * std::max(g.nleg, 1) + 8, 0), p.stream))) return fail(rc);
    {
        std::vector<double> ci(g.nleg), c(g.nleg);
        for (int j = 0; j < g.nleg; ++j) {
            double lat = g.lat_deg[j];
            const double pole = 89.9999999;  // TransLocal.cc:49, :1449-1455
            if (lat > pole) lat = pole;
            if (lat < -pole) lat = -pole;
            c[j] = std::cos(lat * (M_PI / 180.));
            ci[j] = 1. / c[j];
        }
        if ((rc = upload(p.d_coslatinv, ci, p.stream))) return fail(rc);
        if ((rc = upload(p.d_coslat, c, p.stream))) return fail(rc);
        {
            std::vector<double> us(g.nleg);
            for (int j = 0; j < g.nleg; ++j) us[j] = 1. / (6371229. * c[j]);  // util/Earth.h:24
            if ((rc = upload(p.d_uvscale, us, p.stream))) return fail(rc);
        }
        if (!g.weights.empty()) {
            std::vector<double> w(g.weights.begin(), g.weights.begin() + g.nleg);
            if ((rc = upload(p.d_weights, w, p.stream))) return fail(rc);
            std::vector<double> wn(g.nleg);
            for (int j = 0; j < g.nleg; ++j) wn[j] = w[j] / g.nx[j];
            if ((rc = upload(p.d_dirscale, wn, p.stream))) return fail(rc);
            for (int j = 0; j < g.nleg; ++j) wn[j] = w[j] / (g.nx[j] * 6371229. * c[j]);
            if ((rc = upload(p.d_dirscale_uv, wn, p.stream))) return fail(rc);
        }
    }
    if (cudaMalloc(&p.d_tile_counter, sizeof(int)) != cudaSuccess) {
        set_error("cudaMalloc failed");
        return fail(SPTRANS_ERR_CUDA);
    }
    build_exchange(g, p.ex);
    if ((rc = upload(p.d_ex_m, p.ex.m_side, p.stream))) return fail(rc);
    if ((rc = upload(p.d_ex_band, p.ex.band_side, p.stream))) return fail(rc);
    if ((rc = generate_legendre_table(p))) return fail(rc);  // (its transpose is built by the first direct transform)
    // (point-set plans evaluate the Fourier sums point by point, points.cu: no FFT tables)
    if (!g.points && (rc = build_fft_tables(p))) return fail(rc);
    if (g.cropped) {
        if ((rc = upload_crop_rows(p))) return fail(rc);
        g.npts = g.crop_npts;   // grid-point arrays of this plan are [field][crop point]
    }
    if (cudaStreamSynchronize(p.stream) != cudaSuccess) {
        set_error(std::string("plan setup failed: ") + cudaGetErrorString(cudaGetLastError()));
        return fail(SPTRANS_ERR_CUDA);
    }
    *out = sp;
    return SPTRANS_OK;
}

int sptrans_plan_create_sharded(sptrans_plan** out, int nlat, const int* nx, const double* lat_deg,
                                const double* weights, int truncation, unsigned flags, int device, int rank,
                                int nranks) {
    return create_plan(out, nlat, nx, lat_deg, weights, truncation, flags, device, rank, nranks, nullptr);
}

int sptrans_plan_create(sptrans_plan** plan, int nlat, const int* nx, const double* lat_deg, const double* weights,
                        int truncation, unsigned flags, int device) {
    return create_plan(plan, nlat, nx, lat_deg, weights, truncation, flags, device, 0, 1, nullptr);
}

int sptrans_plan_create_cropped(sptrans_plan** plan, int nlat, const int* nx, const double* lat_deg, int truncation, unsigned flags,
                                int device, int jlat_min, int nlat_crop, const int* nx_crop, const int* jlon_min) {
    const Crop crop{jlat_min, nlat_crop, nx_crop, jlon_min};
    return create_plan(plan, nlat, nx, lat_deg, nullptr, truncation, flags, device, 0, 1, nullptr, &crop);
}

int sptrans_plan_create_points(sptrans_plan** plan, size_t npoints, const double* lon_deg, const double* lat_deg,
                               int truncation, int device) {
    if (!plan || npoints == 0 || !lon_deg || !lat_deg || truncation < 0) {
        set_error("sptrans_plan_create_points: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    // rows = distinct |latitude| values, north to south, mirrored about the equator (an equator row stays single)
    std::vector<double> a(npoints);
    for (size_t i = 0; i < npoints; ++i) {
        a[i] = std::fabs(lat_deg[i]);
        if (!(a[i] <= 90.)) {
            set_error("sptrans_plan_create_points: latitude outside [-90, 90]");
            return SPTRANS_ERR_INVALID;
        }
    }
    std::vector<double> north(a);
    std::sort(north.begin(), north.end(), std::greater<double>());
    north.erase(std::unique(north.begin(), north.end()), north.end());
    const bool equator = north.back() == 0.;
    if (equator && north.size() == 1) north.insert(north.begin(), 45.);  // the geometry needs one off-equator row pair
    const int nn = static_cast<int>(north.size());
    {
        // One Legendre table row per distinct |latitude|: nn * (T+2)^2 / 2 doubles (plus the Legendre<->Fourier buffer of
        // the calls).  TransLocal's unstructured path needs O(T^2) memory whatever the number of points (TransLocal.cc:
        // 1289-1392); a genuinely scattered point set (every latitude distinct) can therefore fit the reference and not
        // this plan -- fail early and say so instead of running out of device memory half way through.
        const double table_bytes = 8.0 * nn * (truncation + 2.0) * (truncation + 2.0) / 2.0;
        size_t free_b = 0, total_b = 0;
        if (sptrans_device_count() > 0 && device >= 0 && cudaSetDevice(device) == cudaSuccess &&
            cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && table_bytes > 0.7 * static_cast<double>(free_b)) {
            set_error("sptrans_plan_create_points: " + std::to_string(nn) + " distinct latitudes at truncation " +
                      std::to_string(truncation) + " need a Legendre table of " + std::to_string(static_cast<long long>(table_bytes / 1e9)) +
                      " GB, more than the device can hold; transform the points in batches of fewer distinct latitudes");
            return SPTRANS_ERR_INVALID;
        }
        cudaGetLastError();
    }
    std::vector<double> lat(north);
    for (int j = nn - 1 - (equator ? 1 : 0); j >= 0; --j) lat.push_back(-north[j]);
    const int nlat = static_cast<int>(lat.size());
    // no zonal truncation at any latitude (TransLocal.cc:1331-1362 sums every m): rows long enough for the linear rule
    std::vector<int> nx(nlat, 2 * truncation + 2);
    PointSet ps;
    ps.row.resize(npoints);
    ps.sign.resize(npoints);
    ps.lon.resize(npoints);
    ps.coslatinv.resize(npoints);
    for (size_t i = 0; i < npoints; ++i) {
        ps.row[i] = static_cast<int>(std::lower_bound(north.begin(), north.end(), a[i], std::greater<double>()) - north.begin());
        ps.sign[i] = lat_deg[i] < 0. ? -1. : 1.;
        ps.lon[i] = lon_deg[i] * (M_PI / 180.);                    // util::Constants::degreesToRadians()
        ps.coslatinv[i] = 1. / std::cos(lat_deg[i] * (M_PI / 180.));  // no pole clamp in this path (:1380-1384)
    }
    return create_plan(plan, nlat, nx.data(), lat.data(), nullptr, truncation, SPTRANS_GRID_REGULAR, device, 0, 1, &ps);
}

int sptrans_plan_destroy(sptrans_plan* sp) {
    if (!sp) return SPTRANS_OK;
    Plan& p = sp->p;
    if (p.clones > 0) {
        set_error("sptrans_plan_destroy: the plan still has clones that borrow its tables; destroy them first");
        return SPTRANS_ERR_INVALID;
    }
    cudaSetDevice(p.device);
    if (p.stream) cudaStreamSynchronize(p.stream);
    free_fft_tables(p);
    tc_free(p);
    peer_release(p);
    std::vector<void*> ptrs = {p.d_tiles_inv, p.d_tiles_dir, p.d_tile_counter, p.d_pair_done, p.d_packed, p.d_fourier, p.d_spec,
                               p.d_spec2, p.d_gp, p.d_rows, p.d_band, p.d_crop_rows};
    if (p.parent) p.parent->clones--;   // a clone borrows every table from its parent
    else
        ptrs.insert(ptrs.end(), {p.d_tab, p.d_tabT, p.d_nlat0, p.d_fb_rowoff, p.d_sp_rowoff, p.d_rowoff, p.d_nx, p.d_my_m, p.d_owner,
                                 p.d_weights, p.d_coslatinv, p.d_coslat, p.d_uvscale, p.d_dirscale, p.d_dirscale_uv, p.d_pair_meta, p.d_twiddle,
                                 p.d_chirp, p.d_filt, p.d_fft_order, p.d_ex_m, p.d_ex_band, p.d_pt_row, p.d_pt_sign, p.d_pt_lon,
                                 p.d_pt_coslatinv, p.d_spec_off});
    for (void* q : ptrs)
        if (q) cudaFree(q);
    if (p.h_pinned) cudaFreeHost(p.h_pinned);
    for (auto& e : p.ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : p.ev_chunk)
        if (e) cudaEventDestroy(e);
    if (p.s_mark) cudaStreamDestroy(p.s_mark);
    if (p.s_h2d) cudaStreamDestroy(p.s_h2d);
    if (p.s_d2h) cudaStreamDestroy(p.s_d2h);
    if (p.own_stream && p.stream) cudaStreamDestroy(p.stream);
    delete sp;
    return SPTRANS_OK;
}

int sptrans_truncation(const sptrans_plan* plan) { return plan ? plan->p.g.T : -1; }
size_t sptrans_nb_gridpoints(const sptrans_plan* plan) { return plan ? static_cast<size_t>(plan->p.g.npts) : 0; }
size_t sptrans_nb_spectral_coefficients(const sptrans_plan* plan) {
    return plan ? static_cast<size_t>(plan->p.g.T + 1) * (plan->p.g.T + 2) : 0;
}
int sptrans_get_nlat0(const sptrans_plan* plan, int* nlat0) {
    if (!plan || !nlat0) return SPTRANS_ERR_INVALID;
    std::copy(plan->p.g.nlat0.begin(), plan->p.g.nlat0.begin() + plan->p.g.T + 1, nlat0);
    return SPTRANS_OK;
}
size_t sptrans_device_bytes(const sptrans_plan* plan) {
    if (!plan) return 0;
    const Plan& p = plan->p;
    return p.bytes_tables + (p.packed_cap + p.fourier_cap + p.spec_cap + p.spec2_cap + p.gp_cap + p.rows_cap) * sizeof(double);
}

int sptrans_local_sizes(const sptrans_plan* plan, size_t* spec_doubles_per_field, size_t* gridpoints_per_field) {
    if (!plan) return SPTRANS_ERR_INVALID;
    if (spec_doubles_per_field) *spec_doubles_per_field = 2 * static_cast<size_t>(plan->p.g.spec_ncoef);
    if (gridpoints_per_field) *gridpoints_per_field = static_cast<size_t>(plan->p.g.points ? plan->p.g.npts : plan->p.g.gp_stride);
    return SPTRANS_OK;
}

size_t sptrans_legendre_cache_size(const sptrans_plan* plan) {
    return plan ? legendre_cache_doubles(plan->p.g) * sizeof(double) : 0;
}
int sptrans_export_legendre_cache(const sptrans_plan* plan, void* out) {
    if (!plan || !out) {
        set_error("sptrans_export_legendre_cache: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    SPT_CUDA(cudaSetDevice(plan->p.device));
    return export_legendre_cache(plan->p, static_cast<double*>(out));
}

int sptrans_import_legendre_cache(sptrans_plan* plan, const void* blob, size_t bytes) {
    if (!plan || !blob) {
        set_error("sptrans_import_legendre_cache: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    int rc = check_plan(plan);
    if (rc) return rc;
    return import_legendre_cache(plan->p, static_cast<const double*>(blob), bytes);
}

int sptrans_legendre_cache_uid(char* out, size_t out_len, const char* prefix, int truncation, int kind, int n_or_ny, double south,
                               double north, int nlat, const double* lat_deg, int flt) {
    if (!out || !prefix || out_len == 0 || (kind == SPTRANS_UID_OTHER && (nlat <= 0 || !lat_deg))) {
        set_error("sptrans_legendre_cache_uid: invalid arguments");
        return -1;
    }
    const std::string uid = legendre_cache_uid(prefix, truncation, kind, n_or_ny, south, north, nlat, lat_deg, flt != 0);
    if (uid.size() + 1 > out_len) {
        set_error("sptrans_legendre_cache_uid: output buffer too small");
        return -1;
    }
    std::memcpy(out, uid.c_str(), uid.size() + 1);
    return static_cast<int>(uid.size());
}

size_t sptrans_legendre_cache_estimate(int truncation) {
    return static_cast<size_t>(truncation) * truncation * truncation / 2 * sizeof(double);
}

int sptrans_set_precision(sptrans_plan* plan, int precision) {
    if (!plan || (precision != SPTRANS_PREC_FP64 && precision != SPTRANS_PREC_TC_SPLIT)) {
        set_error("sptrans_set_precision: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    int rc = check_plan(plan);
    if (rc) return rc;
    plan->p.precision = precision;
    if (precision == SPTRANS_PREC_TC_SPLIT) return tc_prepare_tables(plan->p);
    return SPTRANS_OK;
}

int sptrans_set_stream(sptrans_plan* plan, void* cuda_stream) {
    if (!plan) return SPTRANS_ERR_INVALID;
    Plan& p = plan->p;
    if (p.own_stream && p.stream) {
        cudaStreamSynchronize(p.stream);
        cudaStreamDestroy(p.stream);
    }
    p.stream = static_cast<cudaStream_t>(cuda_stream);
    p.own_stream = false;
    return SPTRANS_OK;
}

int sptrans_set_async(sptrans_plan* plan, int on) {
    if (!plan) return SPTRANS_ERR_INVALID;
    plan->p.async = on != 0;
    return SPTRANS_OK;
}

int sptrans_synchronize(sptrans_plan* plan) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    if (p.s_d2h) SPT_CUDA(cudaStreamSynchronize(p.s_d2h));
    if (p.s_h2d) SPT_CUDA(cudaStreamSynchronize(p.s_h2d));
    p.d2h_pending = false;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int sptrans_mark(sptrans_plan* plan, void** mark) {
    int rc = check_plan(plan);
    if (rc) return rc;
    if (!mark) {
        set_error("sptrans_mark: null output pointer");
        return SPTRANS_ERR_INVALID;
    }
    Plan& p = plan->p;
    cudaEvent_t ev = nullptr;
    SPT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (p.s_d2h && p.d2h_pending) {
        // the work enqueued so far ends on two streams (transforms, copies back to the host): join them on a side stream
        // so that neither has to wait for the other
        if (!p.s_mark) SPT_CUDA(cudaStreamCreateWithFlags(&p.s_mark, cudaStreamNonBlocking));
        cudaEvent_t tail = nullptr;
        SPT_CUDA(cudaEventCreateWithFlags(&tail, cudaEventDisableTiming));
        SPT_CUDA(cudaEventRecord(tail, p.stream));
        SPT_CUDA(cudaStreamWaitEvent(p.s_mark, tail, 0));
        SPT_CUDA(cudaStreamWaitEvent(p.s_mark, p.ev_chunk[kEvD2HDone], 0));
        SPT_CUDA(cudaEventRecord(ev, p.s_mark));
        cudaEventDestroy(tail);   // (released once the recorded work has completed)
    }
    else SPT_CUDA(cudaEventRecord(ev, p.stream));
    *mark = ev;
    return SPTRANS_OK;
}

int sptrans_wait_mark(sptrans_plan* plan, void* mark) {
    int rc = check_plan(plan);
    if (rc) return rc;
    if (!mark) {
        set_error("sptrans_wait_mark: null mark");
        return SPTRANS_ERR_INVALID;
    }
    Plan& p = plan->p;
    cudaEvent_t ev = static_cast<cudaEvent_t>(mark);
    SPT_CUDA(cudaStreamWaitEvent(p.stream, ev, 0));
    if ((rc = copy_streams(p))) return rc;
    SPT_CUDA(cudaStreamWaitEvent(p.s_h2d, ev, 0));   // a direct transform starts with uploads on the copy stream
    return SPTRANS_OK;
}

int sptrans_release_mark(void* mark) {
    if (mark) cudaEventDestroy(static_cast<cudaEvent_t>(mark));
    return SPTRANS_OK;
}

int sptrans_plan_clone(sptrans_plan* src, sptrans_plan** out) {
    if (!src || !out) {
        set_error("sptrans_plan_clone: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    *out = nullptr;
    int rc = check_plan(src);
    if (rc) return rc;
    Plan& s = src->p;
    if (s.parent || s.g.nranks != 1 || s.g.points) {
        set_error("sptrans_plan_clone: only unsharded grid plans that are not clones themselves can be cloned");
        return SPTRANS_ERR_INVALID;
    }
    if (!s.d_tabT && s.d_weights && (rc = build_transposed_table(s))) return rc;  // shared by the clones' direct transforms
    sptrans_plan* sp = new (std::nothrow) sptrans_plan();
    if (!sp) {
        set_error("out of host memory");
        return SPTRANS_ERR_INVALID;
    }
    Plan& p = sp->p;
    p.g = s.g;
    p.device = s.device;
    p.flags = s.flags;
    p.num_sms = s.num_sms;
    p.parent = &s;
    // borrowed (never freed by the clone): geometry arrays, Legendre tables, Fourier tables
    p.d_tab = s.d_tab; p.d_tabT = s.d_tabT; p.d_nlat0 = s.d_nlat0; p.d_fb_rowoff = s.d_fb_rowoff; p.d_rowoff = s.d_rowoff;
    p.d_nx = s.d_nx; p.d_weights = s.d_weights; p.d_coslatinv = s.d_coslatinv; p.d_coslat = s.d_coslat;
    p.d_uvscale = s.d_uvscale; p.d_dirscale = s.d_dirscale; p.d_dirscale_uv = s.d_dirscale_uv; p.d_sp_rowoff = s.d_sp_rowoff; p.d_my_m = s.d_my_m;
    p.d_owner = s.d_owner; p.d_spec_off = s.d_spec_off; p.d_pair_meta = s.d_pair_meta; p.d_twiddle = s.d_twiddle;
    p.d_chirp = s.d_chirp; p.d_filt = s.d_filt; p.d_fft_order = s.d_fft_order; p.d_ex_m = s.d_ex_m; p.d_ex_band = s.d_ex_band;
    p.ex = s.ex;
    clone_fft_state(s, p);
    auto fail = [&](int code) {
        sptrans_plan_destroy(sp);
        return code;
    };
    s.clones++;
    if (cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        return fail(SPTRANS_ERR_CUDA);
    }
    p.own_stream = true;
    for (auto& e : p.ev) cudaEventCreate(&e);
    if (cudaMalloc(&p.d_tile_counter, sizeof(int)) != cudaSuccess) {
        set_error("cudaMalloc failed");
        return fail(SPTRANS_ERR_CUDA);
    }
    if ((rc = upload(p.d_pair_done, std::vector<int>(3 * std::max(p.g.nleg, 1) + 8, 0), p.stream))) return fail(rc);
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    *out = sp;
    return SPTRANS_OK;
}

// ---------------------------------------------------------------------------------------------------
int sptrans_invtrans_scalar(sptrans_plan* plan, int nf, const double* spectra, double* gp) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf < 0 || (nf > 0 && (!spectra || !gp))) {
        set_error("sptrans_invtrans_scalar: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (nf == 0) return SPTRANS_OK;  // reference: `if (nb_scalar_fields > 0)` TransLocal.cc:1412
    if (p.g.nranks != 1) {
        set_error("whole-transform entry points need an unsharded plan; use the stage-level API");
        return SPTRANS_ERR_INVALID;
    }
    const int T = p.g.T;
    StageTimer tm(p);
    int marks = 0, slots[8];
    const double* d_spec = spectra;
    double* d_gp = gp;
    const bool spec_host = !is_device_pointer(spectra), gp_host = !is_device_pointer(gp);
    const size_t nspec = spec_doubles(p, nf, T), ngp = static_cast<size_t>(p.g.npts) * nf;
    if (spec_host) {
        if ((rc = ensure(p.d_spec, p.spec_cap, nspec))) return rc;
        tm.mark(marks);
        SPT_CUDA(cudaMemcpyAsync(p.d_spec, spectra, nspec * sizeof(double), cudaMemcpyHostToDevice, p.stream));
        slots[marks++] = 3;
        d_spec = p.d_spec;
    }
    if (gp_host) {
        if ((rc = ensure(p.d_gp, p.gp_cap, ngp))) return rc;
        d_gp = p.d_gp;
    }
    const bool pipelined = gp_host && !p.g.points && !p.g.cropped;   // D2H of finished field chunks behind the Fourier kernels
    if ((rc = run_inverse(p, nf, T, d_spec, d_gp, 0, tm, marks, slots, pipelined ? gp : nullptr))) return rc;
    if (gp_host && !pipelined) {
        SPT_CUDA(cudaMemcpyAsync(gp, d_gp, ngp * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
        slots[marks++] = 4;
        tm.mark(marks);
    }
    tm.finish(marks + 1, slots);
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int sptrans_invtrans(sptrans_plan* plan, int nsc, const double* scalar_spectra, int nvd, const double* vor,
                     const double* div, double* gp) {
    if (nvd <= 0) {
        // reference :1591-1596: scalars only, written at gp + 2*nb_gp*nb_vordiv_fields (= gp)
        return sptrans_invtrans_scalar(plan, nsc, scalar_spectra, gp);
    }
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nsc < 0 || !vor || !div || !gp || (nsc > 0 && !scalar_spectra)) {
        set_error("sptrans_invtrans: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (p.g.nranks != 1) {
        set_error("whole-transform entry points need an unsharded plan; use the stage-level API");
        return SPTRANS_ERR_INVALID;
    }
    const int T = p.g.T;
    const int nall = 2 * nvd + nsc;
    StageTimer tm(p);
    int marks = 0, slots[8];
    const size_t nvd_spec = spec_doubles(p, nvd, T), nsc_spec = spec_doubles(p, nsc, T);
    const size_t ngp = static_cast<size_t>(p.g.npts) * nall;
    // stage inputs: [vor | div | scalars] in d_spec2 when they live on the host
    const double *d_vor = vor, *d_div = div, *d_sc = scalar_spectra;
    const bool in_host = !is_device_pointer(vor);
    if (in_host) {
        if ((rc = ensure(p.d_spec2, p.spec2_cap, 2 * nvd_spec + nsc_spec))) return rc;
        tm.mark(marks);
        SPT_CUDA(cudaMemcpyAsync(p.d_spec2, vor, nvd_spec * sizeof(double), cudaMemcpyHostToDevice, p.stream));
        SPT_CUDA(cudaMemcpyAsync(p.d_spec2 + nvd_spec, div, nvd_spec * sizeof(double), cudaMemcpyHostToDevice, p.stream));
        if (nsc > 0)
            SPT_CUDA(cudaMemcpyAsync(p.d_spec2 + 2 * nvd_spec, scalar_spectra, nsc_spec * sizeof(double),
                                     cudaMemcpyHostToDevice, p.stream));
        slots[marks++] = 3;
        d_vor = p.d_spec2;
        d_div = p.d_spec2 + nvd_spec;
        d_sc = p.d_spec2 + 2 * nvd_spec;
    }
    if ((rc = ensure(p.d_spec, p.spec_cap, spec_doubles(p, nall, T + 1)))) return rc;
    tm.mark(marks);
    if ((rc = launch_merge_uv_scalar(p.stream, T, nvd, nsc, d_vor, d_div, d_sc, p.d_spec, &p.launches))) return rc;
    slots[marks++] = 0;
    double* d_gp = gp;
    const bool gp_host = !is_device_pointer(gp);
    if (gp_host) {
        if ((rc = ensure(p.d_gp, p.gp_cap, ngp))) return rc;
        d_gp = p.d_gp;
    }
    const bool pipelined = gp_host && !p.g.points && !p.g.cropped;
    if ((rc = run_inverse(p, nall, T + 1, p.d_spec, d_gp, 2 * nvd, tm, marks, slots, pipelined ? gp : nullptr))) return rc;
    if (gp_host && !pipelined) {
        SPT_CUDA(cudaMemcpyAsync(gp, d_gp, ngp * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
        slots[marks++] = 4;
        tm.mark(marks);
    }
    tm.finish(marks + 1, slots);
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int sptrans_invtrans_vordiv2wind(sptrans_plan* plan, int nvd, const double* vor, const double* div, double* gp) {
    return sptrans_invtrans(plan, 0, nullptr, nvd, vor, div, gp);  // reference :1488-1492
}

static int dirtrans_scalar_impl(sptrans_plan* plan, int nf, const double* gp, double* spectra, int adjoint) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf < 0 || (nf > 0 && (!spectra || !gp))) {
        set_error("sptrans_dirtrans_scalar: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (nf == 0) return SPTRANS_OK;
    if (p.g.points || p.g.cropped) {  // TransLocal: ATLAS_NOTIMPLEMENTED (TransLocal.cc:1599-1604, :1671-1676)
        set_error("direct and adjoint transforms are not available for point-set and cropped-grid plans");
        return SPTRANS_ERR_NOT_IMPLEMENTED;
    }
    if (p.g.nranks != 1) {
        set_error("whole-transform entry points need an unsharded plan; use the stage-level API");
        return SPTRANS_ERR_INVALID;
    }
    if (!p.d_weights && !adjoint) {
        set_error("sptrans_dirtrans_scalar: plan was created without quadrature weights");
        return SPTRANS_ERR_INVALID;
    }
    const int T = p.g.T;
    StageTimer tm(p);
    int marks = 0, slots[8];
    const bool spec_host = !is_device_pointer(spectra), gp_host = !is_device_pointer(gp);
    const size_t nspec = spec_doubles(p, nf, T), ngp = static_cast<size_t>(p.g.npts) * nf;
    const double* d_gp = gp;
    double* d_spec = spectra;
    if (gp_host) {
        if ((rc = ensure(p.d_gp, p.gp_cap, ngp))) return rc;
        d_gp = p.d_gp;
    }
    if (spec_host) {
        if ((rc = ensure(p.d_spec, p.spec_cap, nspec))) return rc;
        d_spec = p.d_spec;
    }
    if ((rc = build_tiles(p, nf, T, T))) return rc;
    if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nf)))) return rc;
    if ((rc = ensure(p.d_fourier, p.fourier_cap, fourier_doubles(p, nf)))) return rc;
    tm.mark(marks);
    if (gp_host) rc = fourier_dir_from_host(p, nf, gp, p.d_gp, p.d_fourier, 0, adjoint);  // H2D of chunk c+1 || transforms of chunk c
    else rc = launch_fourier_dir(p, nf, d_gp, p.d_fourier, 0, adjoint);
    if (rc) return rc;
    slots[marks++] = 2;
    tm.mark(marks);
    if (p.precision == SPTRANS_PREC_TC_SPLIT) {
        if ((rc = tc_prepare_tables(p))) return rc;
        if ((rc = tc_build_tiles(p, nf, T, T))) return rc;
        if ((rc = launch_legendre_dir_tc(p, nf, p.d_fourier, p.d_packed, p.ev[marks + 1]))) return rc;
        slots[marks++] = 0;   // operand images (split-tf32 Fourier rows), then the tcgen05 GEMM
    }
    else if ((rc = launch_legendre_dir(p, nf, p.d_fourier, p.d_packed))) return rc;
    slots[marks++] = 1;
    tm.mark(marks);
    if ((rc = launch_unpack_spectra(p, nf, p.d_packed, d_spec, adjoint))) return rc;
    slots[marks++] = 0;
    tm.mark(marks);
    if (spec_host) {
        SPT_CUDA(cudaMemcpyAsync(spectra, d_spec, nspec * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
        slots[marks++] = 4;
        tm.mark(marks);
    }
    tm.finish(marks + 1, slots);
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int sptrans_dirtrans_scalar(sptrans_plan* plan, int nf, const double* gp, double* spectra) {
    return dirtrans_scalar_impl(plan, nf, gp, spectra, 0);
}

int sptrans_invtrans_adj_scalar(sptrans_plan* plan, int nf, const double* gp, double* spectra) {
    return dirtrans_scalar_impl(plan, nf, gp, spectra, 1);
}

int sptrans_dirtrans_wind2vordiv(sptrans_plan* plan, int nf, const double* wind, double* vor, double* div) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf < 0 || (nf > 0 && (!wind || !vor || !div))) {
        set_error("sptrans_dirtrans_wind2vordiv: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (nf == 0) return SPTRANS_OK;
    if (p.g.points || p.g.cropped) {
        set_error("direct and adjoint transforms are not available for point-set and cropped-grid plans");
        return SPTRANS_ERR_NOT_IMPLEMENTED;
    }
    if (p.g.nranks != 1) {
        set_error("whole-transform entry points need an unsharded plan; use the stage-level API");
        return SPTRANS_ERR_INVALID;
    }
    if (!p.d_weights) {
        set_error("sptrans_dirtrans_wind2vordiv: plan was created without quadrature weights");
        return SPTRANS_ERR_INVALID;
    }
    const int T = p.g.T;
    const int nall = 2 * nf;
    StageTimer tm(p);
    int marks = 0, slots[8];
    const bool gp_host = !is_device_pointer(wind), sp_host = !is_device_pointer(vor);
    const size_t ngp = static_cast<size_t>(p.g.npts) * nall, nspec = spec_doubles(p, nf, T);
    const double* d_gp = wind;
    double *d_vor = vor, *d_div = div;
    if (gp_host) {
        if ((rc = ensure(p.d_gp, p.gp_cap, ngp))) return rc;
        d_gp = p.d_gp;
    }
    if (sp_host) {
        if ((rc = ensure(p.d_spec, p.spec_cap, 2 * nspec))) return rc;
        d_vor = p.d_spec;
        d_div = p.d_spec + nspec;
    }
    if ((rc = build_tiles(p, nall, T, T + 1))) return rc;
    if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nall)))) return rc;
    if ((rc = ensure(p.d_fourier, p.fourier_cap, fourier_doubles(p, nall)))) return rc;
    tm.mark(marks);
    if (gp_host) rc = fourier_dir_from_host(p, nall, wind, p.d_gp, p.d_fourier, nall, 0);
    else rc = launch_fourier_dir(p, nall, d_gp, p.d_fourier, nall);
    if (rc) return rc;
    slots[marks++] = 2;
    tm.mark(marks);
    if ((rc = launch_legendre_dir(p, nall, p.d_fourier, p.d_packed))) return rc;
    slots[marks++] = 1;
    tm.mark(marks);
    if ((rc = launch_uv_to_vordiv(p.stream, T, nf, p.d_sp_rowoff, p.d_packed, d_vor, d_div, &p.launches))) return rc;
    slots[marks++] = 0;
    tm.mark(marks);
    if (sp_host) {
        SPT_CUDA(cudaMemcpyAsync(vor, d_vor, nspec * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
        SPT_CUDA(cudaMemcpyAsync(div, d_div, nspec * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
        slots[marks++] = 4;
        tm.mark(marks);
    }
    tm.finish(marks + 1, slots);
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int sptrans_invtrans_grad(sptrans_plan* plan, int nf, const double* spectra, double* grad) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf < 0 || (nf > 0 && (!spectra || !grad))) {
        set_error("sptrans_invtrans_grad: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (nf == 0) return SPTRANS_OK;
    if (p.g.nranks != 1) {
        set_error("whole-transform entry points need an unsharded plan; use the stage-level API");
        return SPTRANS_ERR_INVALID;
    }
    const int T = p.g.T;
    const int nall = 2 * nf;
    StageTimer tm(p);
    int marks = 0, slots[8];
    const bool sp_host = !is_device_pointer(spectra), gp_host = !is_device_pointer(grad);
    const size_t nspec = spec_doubles(p, nf, T), ngp = static_cast<size_t>(p.g.npts) * nall;
    const double* d_sp = spectra;
    if (sp_host) {
        if ((rc = ensure(p.d_spec2, p.spec2_cap, nspec))) return rc;
        tm.mark(marks);
        SPT_CUDA(cudaMemcpyAsync(p.d_spec2, spectra, nspec * sizeof(double), cudaMemcpyHostToDevice, p.stream));
        slots[marks++] = 3;
        d_sp = p.d_spec2;
    }
    if ((rc = ensure(p.d_spec, p.spec_cap, spec_doubles(p, nall, T + 1)))) return rc;
    tm.mark(marks);
    if ((rc = launch_grad_spectra(p.stream, T, nf, d_sp, p.d_spec, &p.launches))) return rc;
    slots[marks++] = 0;
    double* d_gp = grad;
    if (gp_host) {
        if ((rc = ensure(p.d_gp, p.gp_cap, ngp))) return rc;
        d_gp = p.d_gp;
    }
    const bool pipelined = gp_host && !p.g.points && !p.g.cropped;
    if ((rc = run_inverse(p, nall, T + 1, p.d_spec, d_gp, nall, tm, marks, slots, pipelined ? grad : nullptr))) return rc;
    if (gp_host && !pipelined) {
        SPT_CUDA(cudaMemcpyAsync(grad, d_gp, ngp * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
        slots[marks++] = 4;
        tm.mark(marks);
    }
    tm.finish(marks + 1, slots);
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int sptrans_vordiv_to_uv(int truncation, int nf, const double* vor, const double* div, double* U, double* V,
                         int device) {
    if (truncation < 0 || nf < 0 || !vor || !div || !U || !V) {
        set_error("sptrans_vordiv_to_uv: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (sptrans_device_count() <= 0) {
        set_error("sptrans_vordiv_to_uv: no CUDA device visible (this engine has no CPU fallback)");
        return SPTRANS_ERR_CUDA;
    }
    SPT_CUDA(cudaSetDevice(device));
    const size_t n = static_cast<size_t>(truncation + 1) * (truncation + 2) * nf;
    if (n == 0) return SPTRANS_OK;
    const bool host = !is_device_pointer(vor);
    double* buf = nullptr;
    const double *d_vor = vor, *d_div = div;
    double *d_U = U, *d_V = V;
    if (host) {
        SPT_CUDA(cudaMalloc(&buf, 4 * n * sizeof(double)));
        SPT_CUDA(cudaMemcpy(buf, vor, n * sizeof(double), cudaMemcpyHostToDevice));
        SPT_CUDA(cudaMemcpy(buf + n, div, n * sizeof(double), cudaMemcpyHostToDevice));
        d_vor = buf;
        d_div = buf + n;
        d_U = buf + 2 * n;
        d_V = buf + 3 * n;
    }
    int rc = launch_vd2uv(nullptr, truncation, nf, d_vor, d_div, d_U, d_V, nullptr);
    if (rc) {
        if (buf) cudaFree(buf);
        return rc;
    }
    SPT_CUDA(cudaDeviceSynchronize());
    if (host) {
        SPT_CUDA(cudaMemcpy(U, d_U, n * sizeof(double), cudaMemcpyDeviceToHost));
        SPT_CUDA(cudaMemcpy(V, d_V, n * sizeof(double), cudaMemcpyDeviceToHost));
        cudaFree(buf);
    }
    return SPTRANS_OK;
}

// ---- stage-level API ---------------------------------------------------------------------------------
int sptrans_fourier_path_stats(sptrans_plan* plan, long long* out8) {
    if (!plan || !out8) {
        set_error("sptrans_fourier_path_stats: null argument");
        return SPTRANS_ERR_INVALID;
    }
    fourier_path_stats(plan->p, out8);
    return SPTRANS_OK;
}
size_t sptrans_fourier_elems_per_field(const sptrans_plan* plan) {
    return plan ? static_cast<size_t>(plan->p.g.fb_rowoff.back()) + kBM : 0;
}

int sptrans_invtrans_legendre(sptrans_plan* plan, int nf, int trunc, const double* d_spectra, double* d_fourier) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf <= 0 || !d_spectra || !d_fourier || (trunc != p.g.T && trunc != p.g.T + 1)) {
        set_error("sptrans_invtrans_legendre: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if ((rc = build_tiles(p, nf, trunc, p.g.T))) return rc;
    if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nf)))) return rc;
    if ((rc = launch_pack_spectra(p, nf, trunc, d_spectra, p.d_packed))) return rc;
    if ((rc = launch_legendre_inv(p, nf, p.d_packed, d_fourier))) return rc;
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    return SPTRANS_OK;
}

int sptrans_invtrans_fourier(sptrans_plan* plan, int nf, int mlimit, const double* d_fourier, double* d_gp,
                             int nb_uv) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf <= 0 || !d_fourier || !d_gp) {
        set_error("sptrans_invtrans_fourier: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if ((rc = launch_fourier_inv(p, nf, mlimit, d_fourier, d_gp, nb_uv))) return rc;
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    return SPTRANS_OK;
}

int sptrans_dirtrans_fourier(sptrans_plan* plan, int nf, const double* d_gp, double* d_fourier, int nb_uv) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf <= 0 || !d_fourier || !d_gp) {
        set_error("sptrans_dirtrans_fourier: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if ((rc = launch_fourier_dir(p, nf, d_gp, d_fourier, nb_uv))) return rc;
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    return SPTRANS_OK;
}

int sptrans_dirtrans_legendre(sptrans_plan* plan, int nf, const double* d_fourier, double* d_spectra) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf <= 0 || !d_fourier || !d_spectra) {
        set_error("sptrans_dirtrans_legendre: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if ((rc = build_tiles(p, nf, p.g.T, p.g.T))) return rc;
    if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nf)))) return rc;
    if ((rc = launch_legendre_dir(p, nf, d_fourier, p.d_packed))) return rc;
    if ((rc = launch_unpack_spectra(p, nf, p.d_packed, d_spectra))) return rc;
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    return SPTRANS_OK;
}

int sptrans_shard_layout(int nlat, const int* nx, const double* lat_deg, int truncation, unsigned flags, int rank,
                         int nranks, int* owner, int* band, long long* m_side_rows, long long* band_side_rows) {
    HostGeom g;
    int rc = build_geometry(g, nlat, nx, lat_deg, nullptr, truncation, (flags & SPTRANS_GRID_REGULAR) != 0, rank, nranks);
    if (rc) return rc;
    ExchangeLayout ex;
    build_exchange(g, ex);
    if (owner) std::copy(g.owner.begin(), g.owner.end(), owner);
    if (band) std::copy(g.band.begin(), g.band.end(), band);
    if (m_side_rows) std::copy(ex.m_side_rows.begin(), ex.m_side_rows.end(), m_side_rows);
    if (band_side_rows) std::copy(ex.band_side_rows.begin(), ex.band_side_rows.end(), band_side_rows);
    return SPTRANS_OK;
}

long long sptrans_shard_segments(int nlat, const int* nx, const double* lat_deg, int truncation, unsigned flags,
                                 int rank, int nranks, int side, long long* out) {
    HostGeom g;
    int rc = build_geometry(g, nlat, nx, lat_deg, nullptr, truncation, (flags & SPTRANS_GRID_REGULAR) != 0, rank, nranks);
    if (rc) return -1;
    ExchangeLayout ex;
    build_exchange(g, ex);
    const std::vector<ExSeg>& v = side == 0 ? ex.m_side : ex.band_side;
    if (out)
        for (size_t i = 0; i < v.size(); ++i) {
            out[3 * i] = v[i].fb_row;
            out[3 * i + 1] = v[i].buf_row;
            out[3 * i + 2] = v[i].nrows;
        }
    return static_cast<long long>(v.size());
}

int sptrans_exchange_rows(const sptrans_plan* plan, long long* m_side_rows, long long* band_side_rows) {
    if (!plan || !m_side_rows || !band_side_rows) return SPTRANS_ERR_INVALID;
    const ExchangeLayout& ex = plan->p.ex;
    std::copy(ex.m_side_rows.begin(), ex.m_side_rows.end(), m_side_rows);
    std::copy(ex.band_side_rows.begin(), ex.band_side_rows.end(), band_side_rows);
    return SPTRANS_OK;
}

int sptrans_exchange_pack(sptrans_plan* plan, int nf, int side, const double* d_fourier, double* d_buf) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf <= 0 || !d_fourier || !d_buf) {
        set_error("sptrans_exchange_pack: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    const bool ms = side == 0;
    rc = launch_exchange_copy(p, nf, ms ? p.d_ex_m : p.d_ex_band, static_cast<int>(ms ? p.ex.m_side.size() : p.ex.band_side.size()),
                              const_cast<double*>(d_fourier), d_buf, true);
    if (rc) return rc;
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    return SPTRANS_OK;
}

int sptrans_exchange_unpack(sptrans_plan* plan, int nf, int side, const double* d_buf, double* d_fourier) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf <= 0 || !d_fourier || !d_buf) {
        set_error("sptrans_exchange_unpack: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    const bool ms = side == 0;
    rc = launch_exchange_copy(p, nf, ms ? p.d_ex_m : p.d_ex_band, static_cast<int>(ms ? p.ex.m_side.size() : p.ex.band_side.size()),
                              d_fourier, const_cast<double*>(d_buf), false);
    if (rc) return rc;
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    return SPTRANS_OK;
}

// ---- peer-memory exchange of a sharded plan (NVLink / NVSwitch node, one process per GPU) ----------------------

static int peer_ready(Plan& p, int nf, const char* who) {
    const PeerState& ps = p.peer;
    if (!ps.region || ps.nranks != p.g.nranks || ps.nf != nf) {
        set_error(std::string(who) + ": peer buffers are not allocated/attached for this number of fields");
        return SPTRANS_ERR_INVALID;
    }
    for (int r = 0; r < ps.nranks; ++r)
        if (!ps.peer_region[r]) {
            set_error(std::string(who) + ": peer region of rank " + std::to_string(r) + " is not attached");
            return SPTRANS_ERR_INVALID;
        }
    if (p.precision != SPTRANS_PREC_FP64) {
        set_error(std::string(who) + ": the peer-memory exchange is implemented for the fp64 Legendre kernel");
        return SPTRANS_ERR_NOT_IMPLEMENTED;
    }
    return SPTRANS_OK;
}

int sptrans_peer_alloc(sptrans_plan* plan, int nf, unsigned char* ipc_handle_out) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf <= 0 || p.g.nranks > kMaxPeers) {
        set_error("sptrans_peer_alloc: invalid arguments (at most 8 ranks: one NVSwitch node)");
        return SPTRANS_ERR_INVALID;
    }
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    peer_release(p);
    PeerState& ps = p.peer;
    ps.nranks = p.g.nranks;
    ps.nf = nf;
    ps.buf_doubles = fourier_doubles(p, nf);
    const size_t bytes = kPeerFlagBytes + 2 * ps.buf_doubles * sizeof(double);
    SPT_CUDA(cudaMalloc(&ps.region, bytes));
    SPT_CUDA(cudaMemset(ps.region, 0, bytes));
    SPT_CUDA(cudaDeviceSynchronize());
    ps.peer_region[p.g.rank] = ps.region;
    if (ipc_handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == SPTRANS_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t h;
        SPT_CUDA(cudaIpcGetMemHandle(&h, ps.region));
        std::memcpy(ipc_handle_out, &h, sizeof(h));
    }
    return SPTRANS_OK;
}

int sptrans_peer_attach_ipc(sptrans_plan* plan, int nranks, const unsigned char* handles) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    PeerState& ps = p.peer;
    if (!ps.region || nranks != p.g.nranks || !handles) {
        set_error("sptrans_peer_attach_ipc: call sptrans_peer_alloc first; nranks must match the plan");
        return SPTRANS_ERR_INVALID;
    }
    for (int r = 0; r < nranks; ++r) {
        if (r == p.g.rank) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + static_cast<size_t>(r) * SPTRANS_IPC_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        SPT_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ps.peer_region[r] = ptr;
        ps.ipc_opened[r] = true;
    }
    return SPTRANS_OK;
}

int sptrans_peer_attach_ptrs(sptrans_plan* plan, int nranks, void* const* regions) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    PeerState& ps = p.peer;
    if (!ps.region || nranks != p.g.nranks || !regions) {
        set_error("sptrans_peer_attach_ptrs: call sptrans_peer_alloc first; nranks must match the plan");
        return SPTRANS_ERR_INVALID;
    }
    for (int r = 0; r < nranks; ++r)
        if (r != p.g.rank) ps.peer_region[r] = regions[r];
    return SPTRANS_OK;
}

int sptrans_peer_region(const sptrans_plan* plan, void** region, size_t* bytes) {
    if (!plan || !plan->p.peer.region) {
        set_error("sptrans_peer_region: no peer region allocated");
        return SPTRANS_ERR_INVALID;
    }
    if (region) *region = plan->p.peer.region;
    if (bytes) *bytes = kPeerFlagBytes + 2 * plan->p.peer.buf_doubles * sizeof(double);
    return SPTRANS_OK;
}

int sptrans_peer_buffer(const sptrans_plan* plan, double** d_fourier) {
    if (!plan || !plan->p.peer.region || !d_fourier) {
        set_error("sptrans_peer_buffer: no peer region allocated");
        return SPTRANS_ERR_INVALID;
    }
    *d_fourier = make_peer_dst(plan->p).base[plan->p.g.rank];
    return SPTRANS_OK;
}

int sptrans_peer_free(sptrans_plan* plan) {
    int rc = check_plan(plan);
    if (rc) return rc;
    SPT_CUDA(cudaStreamSynchronize(plan->p.stream));
    peer_release(plan->p);
    return SPTRANS_OK;
}

int sptrans_peer_barrier(sptrans_plan* plan) {
    int rc = check_plan(plan);
    if (rc) return rc;
    if ((rc = peer_ready(plan->p, plan->p.peer.nf, "sptrans_peer_barrier"))) return rc;
    return launch_peer_barrier(plan->p);
}

int sptrans_peer_advance(sptrans_plan* plan) {
    if (!plan) return SPTRANS_ERR_INVALID;
    plan->p.peer.parity ^= 1;
    return SPTRANS_OK;
}

// The four stream-ordered halves of the sharded transforms (no host synchronisation; the caller synchronises
// the stream).  sptrans_invtrans_sharded / sptrans_dirtrans_sharded chain them with the barrier.
int sptrans_invtrans_legendre_peers(sptrans_plan* plan, int nf, const double* d_spectra) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if ((rc = peer_ready(p, nf, "sptrans_invtrans_legendre_peers"))) return rc;
    if (!d_spectra) {
        set_error("sptrans_invtrans_legendre_peers: null spectra");
        return SPTRANS_ERR_INVALID;
    }
    if ((rc = build_tiles(p, nf, p.g.T, p.g.T))) return rc;
    if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nf)))) return rc;
    if ((rc = launch_pack_spectra(p, nf, p.g.T, d_spectra, p.d_packed))) return rc;
    return launch_legendre_inv_peers(p, nf, p.d_packed, make_peer_dst(p));
}

int sptrans_dirtrans_fourier_peers(sptrans_plan* plan, int nf, const double* d_gp) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if ((rc = peer_ready(p, nf, "sptrans_dirtrans_fourier_peers"))) return rc;
    if (!d_gp) {
        set_error("sptrans_dirtrans_fourier_peers: null grid-point array");
        return SPTRANS_ERR_INVALID;
    }
    bool fused = false;
    if ((rc = launch_fourier_dir_peers(p, nf, d_gp, make_peer_dst(p), &fused))) return rc;
    return fused ? SPTRANS_OK : launch_exchange_push(p, nf);
}

// Direct transform, exchange by PULL (SPTRANS_DIR_PULL=1; the default stays the push): the Fourier stage writes its rows into the
// local exchange buffer only, and after the barrier the Legendre GEMM fetches every operand row from the buffer of the rank
// that owns its latitude band.  Measured (profiles/emul8_timing_r02.txt, profiles/bench_tco1279_r02_n2_pull_vs_push.txt):
// the push of completed latitude pairs costs the Fourier stage 0.4-0.6 ms per rank at 8 ranks (an extra read + write of every
// row), which the pull does not have -- but on real NVLink the GEMM's operand copies then run at the link's latency: at N = 2
// the Fourier stage gains 0.95 ms and the Legendre stage loses 1.4-1.9 ms.  Kept as an option and as the measurement.
int sptrans_dirtrans_fourier_local(sptrans_plan* plan, int nf, const double* d_gp) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if ((rc = peer_ready(p, nf, "sptrans_dirtrans_fourier_local"))) return rc;
    if (!d_gp) {
        set_error("sptrans_dirtrans_fourier_local: null grid-point array");
        return SPTRANS_ERR_INVALID;
    }
    return launch_fourier_dir(p, nf, d_gp, make_peer_dst(p).base[p.g.rank], 0, 0);
}

int sptrans_dirtrans_legendre_pull(sptrans_plan* plan, int nf, double* d_spectra) {
    int rc = check_plan(plan);
    if (rc) return rc;
    Plan& p = plan->p;
    if ((rc = peer_ready(p, nf, "sptrans_dirtrans_legendre_pull"))) return rc;
    if (!d_spectra) {
        set_error("sptrans_dirtrans_legendre_pull: null spectra");
        return SPTRANS_ERR_INVALID;
    }
    if ((rc = build_tiles(p, nf, p.g.T, p.g.T))) return rc;
    if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nf)))) return rc;
    if ((rc = launch_legendre_dir_peers(p, nf, make_peer_dst(p), p.d_packed))) return rc;
    return launch_unpack_spectra(p, nf, p.d_packed, d_spectra);
}

static bool dir_exchange_push() {
    static const bool v = [] {
        const char* e = std::getenv("SPTRANS_DIR_PULL");
        return !(e && std::atoi(e) != 0);
    }();
    return v;
}

int sptrans_invtrans_sharded(sptrans_plan* plan, int nf, const double* d_spectra, double* d_gp) {
    if (!plan || !d_gp) {
        set_error("sptrans_invtrans_sharded: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    Plan& p = plan->p;
    int rc;
    cudaEventRecord(p.ev[0], p.stream);
    if ((rc = sptrans_invtrans_legendre_peers(plan, nf, d_spectra))) return rc;
    cudaEventRecord(p.ev[1], p.stream);
    if ((rc = launch_peer_barrier(p))) return rc;
    cudaEventRecord(p.ev[2], p.stream);
    double* local = make_peer_dst(p).base[p.g.rank];
    if ((rc = launch_fourier_inv(p, nf, p.g.T - 1, local, d_gp, 0))) return rc;
    cudaEventRecord(p.ev[3], p.stream);
    p.pending_marks = 4;
    const int slots[3] = {1, 5, 2};  // legendre (incl. pack), exchange wait, fourier
    std::memcpy(p.pending_slots, slots, sizeof(slots));
    p.peer.parity ^= 1;
    return SPTRANS_OK;
}

int sptrans_dirtrans_sharded(sptrans_plan* plan, int nf, const double* d_gp, double* d_spectra) {
    if (!plan || !d_spectra) {
        set_error("sptrans_dirtrans_sharded: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    Plan& p = plan->p;
    int rc;
    cudaEventRecord(p.ev[0], p.stream);
    const bool push = dir_exchange_push();   // default; SPTRANS_DIR_PULL=1: the Legendre GEMM pulls its rows instead
    if ((rc = push ? sptrans_dirtrans_fourier_peers(plan, nf, d_gp) : sptrans_dirtrans_fourier_local(plan, nf, d_gp))) return rc;
    cudaEventRecord(p.ev[1], p.stream);
    if ((rc = launch_peer_barrier(p))) return rc;
    cudaEventRecord(p.ev[2], p.stream);
    if (push) {
        double* local = make_peer_dst(p).base[p.g.rank];
        if ((rc = build_tiles(p, nf, p.g.T, p.g.T))) return rc;
        if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nf)))) return rc;
        if ((rc = launch_legendre_dir(p, nf, local, p.d_packed))) return rc;
        if ((rc = launch_unpack_spectra(p, nf, p.d_packed, d_spectra))) return rc;
    }
    else if ((rc = sptrans_dirtrans_legendre_pull(plan, nf, d_spectra))) return rc;
    cudaEventRecord(p.ev[3], p.stream);
    p.pending_marks = 4;
    const int slots[3] = {2, 5, 1};  // fourier (+ push), exchange wait, legendre (pulling its rows; incl. unpack)
    std::memcpy(p.pending_slots, slots, sizeof(slots));
    p.peer.parity ^= 1;
    return SPTRANS_OK;
}

int sptrans_last_timings(const sptrans_plan* plan, float out_ms[8]) {
    if (!plan || !out_ms) return SPTRANS_ERR_INVALID;
    if (plan->p.pending_marks > 0) {   // stream-ordered sharded call: read the events now
        Plan& p = const_cast<Plan&>(plan->p);
        cudaStreamSynchronize(p.stream);
        for (float& t : p.t_ms) t = 0.f;
        for (int i = 0; i + 1 < p.pending_marks; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]);
            p.t_ms[p.pending_slots[i]] += ms;
        }
        p.pending_marks = 0;
    }
    std::memcpy(out_ms, plan->p.t_ms, 8 * sizeof(float));
    return SPTRANS_OK;
}

uint64_t sptrans_kernel_launches(const sptrans_plan* plan) { return plan ? plan->p.launches : 0; }

}  // extern "C"
