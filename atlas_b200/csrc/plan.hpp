// Internal plan structures of the B200 spectral-transform engine (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "sptrans_b200.h"

namespace sptrans {

// ----- tile geometry of the fp64 DMMA Legendre kernels (see legendre_f64.cu) -----
constexpr int kBM = 128;      // rows per CTA tile (latitudes in the inverse, total wavenumbers in the direct)
constexpr int kBN = 144;      // columns per CTA tile (2*field + re/im)
// contraction steps per pipeline stage x stages: 32 x 3 (215 KB of shared memory) measured 3-5 % faster than 16 x 4
// with the TMA-fed kernel (half as many stage hand-overs); the cp.async baseline (-DSPT_BULK=0) needs 16 x 4
#ifndef SPT_BK
#define SPT_BK 32
#endif
#ifndef SPT_STAGES
#define SPT_STAGES 3
#endif
constexpr int kBK = SPT_BK;          // contraction step
constexpr int kStages = SPT_STAGES;  // depth of the shared-memory operand ring
constexpr int kLegThreads = 256;

// One CTA tile of a ragged batched GEMM  C[M x N] = A[M x K] * B[K x N].
struct alignas(16) LegTile {
    long long a_off;  // doubles, into the Legendre table: first element of the tile
    long long b_off;  // doubles, into B (packed spectra / Fourier buffer): row 0, column n0
    long long c_off;  // doubles, into C: (row m0, column n0)
    int a_pitch;      // row pitch of the table block (doubles)
    int m_valid;      // valid rows of this tile (<= kBM)
    int n_valid;      // valid columns of this tile (<= kBN), even
    int k_steps;      // number of kBK steps
    int a_rows;       // inverse: #lat columns readable from a_off (pitch - lat0); direct: #table rows readable
    int b_rows;       // rows of B that may be read (rest zero-filled)
    int lat0;         // inverse: latitude-pair index of tile row 0 (selects the destination rank of a sharded plan)
    int ks_last;      // 4-wide DMMA k-steps of the LAST stage that hold operand rows (the rest of the stage is padding)
};

struct HostGeom {
    int T = 0;        // truncation
    int nlat = 0;     // latitude rows
    int nleg = 0;     // (nlat+1)/2 : northern rows incl. equator   (nlatsLeg_, TransLocal.cc:435)
    bool regular = false;
    bool has_equator = false;
    std::vector<int> nx;            // [nlat]
    std::vector<long long> rowoff;  // [nlat+1]
    std::vector<double> lat_deg;    // [nlat]
    std::vector<double> weights;    // [nlat] or empty
    std::vector<int> nlat0;         // [T+2]  (entry T+1 == nleg)
    std::vector<int> mmax;          // [nleg] highest m with nlat0[m] <= j  (-1 if none)
    int nxmax = 0;
    long long npts = 0;
    // layout of the grid-point and spectral arrays the entry points read / write.  Default: the global arrays of the
    // reference ([field][all points], [all coefficients][field]).  With SPTRANS_SHARD_LOCAL_IO a sharded plan addresses
    // only its own share: grid arrays [field][rows of my latitude band: northern rows, then their southern mirrors],
    // spectral arrays [my zonal wavenumbers ascending][n][re/im][field].
    bool local_io = false;
    // cropped (regional) structured grids -- a nest of this global grid (TransLocal.cc:371-531): the transform runs on the
    // latitude pairs [pair_begin, pair_end) that hold the crop's rows, with the GLOBAL grid's zonal truncation and row
    // lengths, into a band-layout work array; crop row r then takes crop_nx[r] points of global row crop_jlat_min + r
    // starting at longitude index crop_jlon_min[r], wrapping around (the copy-out of TransLocal.cc:1180-1187)
    bool cropped = false;
    int crop_jlat_min = 0;
    std::vector<int> crop_nx, crop_jlon_min;
    long long crop_npts = 0;
    std::vector<long long> gp_rowoff;    // [nlat] offset of row j within one field of the grid array (-1: not held)
    long long gp_stride = 0;             // points per field in the grid array
    std::vector<long long> spec_off;     // [T+1] first complex coefficient of zonal wavenumber m at truncation T (-1: not held)
    long long spec_ncoef = 0;            // complex coefficients per field in the spectral array
    // Legendre table layout (n ascending, k = (n-m-p)/2), built for truncation T+1
    std::vector<long long> tab_off;  // [2*(T+1)] index 2*m+p : offset in doubles
    std::vector<int> tab_K;          // [2*(T+1)] rows with n <= T+1
    std::vector<int> tab_pitch;      // [T+1]     roundup(nleg - nlat0[m], 16)
    long long tab_size = 0;          // doubles
    // packed-spectra layout [m][p][Kpad][2 nf]: row prefix (multiply by 2*nf)
    std::vector<long long> sp_rowoff;  // [2*(T+1)+1]
    // Fourier buffer layout [m][p][jj][fld](re,im): row prefix (multiply by nf), in double2 units
    std::vector<long long> fb_rowoff;  // [T+2] : sum_{m'<m} 2*(nleg - nlat0[m'])
    // sharding
    int rank = 0, nranks = 1;
    std::vector<int> my_m;       // zonal wavenumbers owned by this rank (all if nranks==1)
    std::vector<int> owner;      // [T+1] rank that owns zonal wavenumber m
    std::vector<int> band;       // [nranks+1] latitude-pair band boundaries
    int pair_begin = 0, pair_end = 0;  // latitude pairs [begin,end) of this rank's Fourier band
    // point-set plans (sptrans_plan_create_points): the rows are the distinct |latitudes| of the points, mirrored;
    // the Fourier stage is a direct sum per point
    bool points = false;
    std::vector<int> pt_row;        // [npts] northern row index of the point's |latitude|
    std::vector<double> pt_sign;    // [npts] +1 north / equator, -1 south (antisymmetric part flips)
    std::vector<double> pt_lon;     // [npts] longitude in radians (degrees * pi/180, as the reference forms it)
    std::vector<double> pt_coslatinv;  // [npts] 1 / cos(lat), not clamped (TransLocal.cc:1380-1384)
};

inline int num_n(int truncation, int m, int parity) {  // #n in [m, truncation] with (n-m)%2 == parity
    int len = (truncation - m + (parity ? 1 : 2)) / 2;
    return len < 0 ? 0 : len;
}
inline int round_up(int x, int q) { return (x + q - 1) / q * q; }

// One contiguous run of rows of the Legendre<->Fourier exchange buffer (a row = nf double2) and where it
// sits in a packed send / receive buffer.
struct ExSeg {
    long long fb_row;
    long long buf_row;
    int nrows;
    int peer;   // rank at the other end of this run
};
// Exchange plan between the m-sharded Legendre stage and the latitude-band-sharded Fourier stage.
// m_side: segments of the rows this rank owns as m-owner, grouped by the band owner they travel to/from;
// band_side: segments of the rows this rank owns as band owner, grouped by m-owner.
struct ExchangeLayout {
    std::vector<ExSeg> m_side, band_side;
    std::vector<long long> m_side_rows, band_side_rows;  // [nranks] rows per peer
};
// Peer-memory exchange of a sharded plan (one GPU per process on an NVLink/NVSwitch node): every rank owns a region
//   [ flags: kPeerFlagBytes | exchange buffer 0 | exchange buffer 1 ]
// mapped into all peers (CUDA IPC or pointers supplied by the host).  Producers store rows straight into the
// consumer's buffer; exchanges alternate between the two buffers, so that one device-side barrier per
// exchange is enough (a rank can only overwrite buffer b after every peer has passed the barrier of the
// following exchange, i.e. has finished reading b).
constexpr int kMaxPeers = 8;
constexpr size_t kPeerFlagBytes = 4096;
struct PeerDst {            // kernel argument: where rows of the exchange buffer go
    double* base[kMaxPeers];
    int band[kMaxPeers + 1];
    int nranks;
};
struct PeerState {
    int nranks = 0;
    int nf = 0;
    size_t buf_doubles = 0;
    void* region = nullptr;                 // local allocation
    void* peer_region[kMaxPeers] = {};      // mapped regions, [me] == region
    bool ipc_opened[kMaxPeers] = {};
    int parity = 0;
    unsigned long long epoch = 0;
};
// per distinct row length: Bluestein / chirp-z tables on the device
struct FftLen {
    int n = 0;       // row length
    int L = 0;       // highest zonal wavenumber used at this length
    int M = 0;       // power-of-two convolution length >= n + 2L
    int logM = 0;
    long long chirp_off = 0;   // double2 units into d_chirp: A_u (2L+1) then C_i (n)
    long long filt_off = 0;    // double2 units into d_filt: Bhat (M), digit-reversed order
};

struct Plan {
    HostGeom g;
    int device = 0;
    unsigned flags = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int num_sms = 148;
    // device tables
    double* d_tab = nullptr;       // Legendre table
    double* d_tabT = nullptr;      // its transpose per (m, parity) block: K-major operand of the direct transform
    int* d_nlat0 = nullptr;        // [T+2]
    long long* d_fb_rowoff = nullptr;  // [T+2]
    long long* d_rowoff = nullptr;     // [nlat+1]
    int* d_nx = nullptr;               // [nlat]
    double* d_weights = nullptr;       // [nleg]
    double* d_coslatinv = nullptr;     // [nleg] 1/cos(lat), latitude clamped to +-89.9999999 (TransLocal.cc:1447-1458)
    double* d_coslat = nullptr;        // [nleg]
    double* d_uvscale = nullptr;       // [nleg] 1/(a cos(lat)): wind -> scaled wind of the direct vor/div transform
    double* d_dirscale = nullptr;      // [nleg] quadrature weight / nx: latitude factor of the adjoint of the direct transform
    double* d_dirscale_uv = nullptr;   // [nleg] weight / (nx a cos(lat)): the same for the adjoint of wind -> vor/div
    long long* d_sp_rowoff = nullptr;  // [2(T+1)+1]
    int* d_my_m = nullptr;             // [my_m.size()]
    long long* d_spec_off = nullptr;   // [T+1] HostGeom::spec_off (only with SPTRANS_SHARD_LOCAL_IO, else null)
    int* d_owner = nullptr;            // [T+1] rank that owns zonal wavenumber m
    int* d_pair_done = nullptr;        // [nleg] field-group blocks finished per latitude pair (sharded direct Fourier), then the
                                       // work queue of completed pairs whose rows are being shipped ([2 nleg + 8], fourier.cu)
    int* d_pt_row = nullptr;           // point-set plans: see HostGeom
    double* d_pt_sign = nullptr;
    double* d_pt_lon = nullptr;
    double* d_pt_coslatinv = nullptr;
    // FFT tables
    std::vector<FftLen> fft_len;       // distinct lengths
    std::vector<int> pair_len_idx;     // [nleg] -> index into fft_len
    void* d_pair_meta = nullptr;       // [nleg] PairMeta (fourier.cu)
    double2* d_twiddle = nullptr;      // master twiddle table e^{-2 pi i k/8192}
    double2* d_chirp = nullptr;
    double2* d_filt = nullptr;
    int* d_fft_order = nullptr;        // block schedule, longest rows first
    int fft_smem_max = 0;
    // tile lists (device) for the current nf; rebuilt when nf / truncation changes
    int tiles_nf = -1, tiles_trunc = -1, tiles_dir_trunc = -1;
    LegTile* d_tiles_inv = nullptr;
    int n_tiles_inv = 0;
    LegTile* d_tiles_dir = nullptr;
    int n_tiles_dir = 0;
    int* d_tile_counter = nullptr;
    // workspaces (grown on demand)
    double* d_packed = nullptr;   size_t packed_cap = 0;   // packed spectra [m][p][Kpad][2nf]
    double* d_fourier = nullptr;  size_t fourier_cap = 0;  // Fourier buffer
    double* d_spec = nullptr;     size_t spec_cap = 0;     // device copy of spectra (host-pointer mode)
    double* d_spec2 = nullptr;    size_t spec2_cap = 0;
    double* d_gp = nullptr;       size_t gp_cap = 0;       // device copy of grid fields (host-pointer mode)
    double* d_rows = nullptr;     size_t rows_cap = 0;     // row-layout image of a Field-layout grid buffer
    double* d_band = nullptr;     size_t band_cap = 0;     // cropped plans: rows of the latitude band the crop is cut from
    void* d_crop_rows = nullptr;                           // cropped plans: CropRow[crop rows] (fields.cu)
    void* h_pinned = nullptr;     size_t pinned_cap = 0;
    int precision = 0;           // SPTRANS_PREC_FP64 | SPTRANS_PREC_TC_SPLIT
    void* tc = nullptr;          // TcState (legendre_tc.cu)
    void* fft = nullptr;         // FftState (fourier.cu): launch groups, per-pair metadata, auxiliary streams
    // host-pointer pipelines: copies run on their own streams, field chunk by field chunk, next to the transforms
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr, s_mark = nullptr;
    cudaEvent_t ev_chunk[34] = {};   // hand-over events of the chunks; [32] last D2H copy done, [33] staged grid fields consumed
    bool d2h_pending = false;        // a chunked copy back to the host may still be running on s_d2h
    bool gp_in_use = false;          // ev_chunk[33] has been recorded
    bool async = false;              // sptrans_set_async: whole-transform calls return after enqueueing
    Plan* parent = nullptr;          // sptrans_plan_clone: tables are borrowed from this plan
    int clones = 0;
    ExchangeLayout ex;
    ExSeg* d_ex_m = nullptr;
    ExSeg* d_ex_band = nullptr;
    PeerState peer;
    size_t bytes_tables = 0;
    // stats
    uint64_t launches = 0;
    float t_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[8] = {};
    int pending_marks = 0;      // stream-ordered (sharded) calls: events recorded, elapsed times read lazily
    int pending_slots[8] = {};
};

// ---- error handling ----
void set_error(const std::string& msg);
const char* last_error_cstr();
#define SPT_CUDA(call)                                                                               \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            ::sptrans::set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                                 std::to_string(__LINE__) + ")");                                    \
            return SPTRANS_ERR_CUDA;                                                                 \
        }                                                                                            \
    } while (0)

void build_exchange(const HostGeom& g, ExchangeLayout& ex);

// ---- api.cu ----
bool is_device_pointer(const void* ptr);                          // cudaPointerGetAttributes: device or managed
int ensure(double*& buf, size_t& cap, size_t need_doubles);       // grow-only device workspace

// ---- host_setup.cc ----
int fourier_truncation(int truncation, int nx, int nxmax, int ndgl, double lat, bool fullgrid);
void gaussian_quadrature(int N, double* lat_deg_2N, double* weights_2N);
int build_geometry(HostGeom& g, int nlat, const int* nx, const double* lat_deg, const double* weights, int T,
                   bool regular, int rank, int nranks);
void set_io_layout(HostGeom& g, bool local_io);   // (band layout also for cropped plans)
std::string legendre_cache_uid(const char* prefix, int truncation, int kind, int n_or_ny, double south, double north, int nlat,
                               const double* lat_deg, bool flt);   // fills gp_rowoff / gp_stride / spec_off / spec_ncoef
// per-latitude seeds for the device Legendre recurrence: x=cos(theta), s=sin(theta), columns m=0,1 and the diagonal
void legendre_seeds(int trc, int nlats, const double* lats_rad, std::vector<double>& x, std::vector<double>& col0,
                    std::vector<double>& col1, std::vector<double>& diag);

// ---- legendre_gen.cu ----
int generate_legendre_table(Plan& p);
int export_legendre_cache(const Plan& p, double* h_out);
int import_legendre_cache(Plan& p, const double* h_blob, size_t bytes);
size_t legendre_cache_doubles(const HostGeom& g);

// ---- legendre_f64.cu ----
int build_tiles(Plan& p, int nf, int trunc, int dir_trunc);
// flags: kPackKeepMT keeps the m == trunc column (dropped by the scalar inverse, TransLocal.cc:982);
//        kPackDirAdj drops Im(m = 0) (operand of the adjoint of the direct transform)
constexpr int kPackKeepMT = 1, kPackDirAdj = 2;
int launch_pack_spectra(Plan& p, int nf, int trunc, const double* d_spec, double* d_packed, int flags = 0);
int launch_unpack_spectra(Plan& p, int nf, const double* d_packed, double* d_spec, int drop_mT = 0);
int launch_legendre_inv(Plan& p, int nf, const double* d_packed, double* d_fourier);
// same, every output row stored into the exchange buffer of the rank that owns its latitude band
int launch_legendre_inv_peers(Plan& p, int nf, const double* d_packed, const PeerDst& dst);
int launch_legendre_dir(Plan& p, int nf, const double* d_fourier, double* d_packed);
// same, every Fourier row pulled from the exchange buffer of the rank that owns its latitude band (TMA copies over NVLink)
int launch_legendre_dir_peers(Plan& p, int nf, const PeerDst& src, double* d_packed);
int build_transposed_table(Plan& p);   // lazily, by the first direct transform; rebuilt after a cache import

// ---- legendre_tc.cu (tcgen05 split-TF32 path) ----
int tc_prepare_tables(Plan& p);
int tc_build_tiles(Plan& p, int nf, int trunc, int dir_trunc);
// after_pack (optional): recorded between the operand-image kernel and the tensor-core GEMM (stage timings)
int launch_legendre_inv_tc(Plan& p, int nf, int trunc, const double* d_spec, double* d_fourier, cudaEvent_t after_pack = nullptr);
int launch_legendre_dir_tc(Plan& p, int nf, const double* d_fourier, double* d_packed, cudaEvent_t after_pack = nullptr);
void tc_free(Plan& p);

// ---- fourier.cu ----
int build_fft_tables(Plan& p);
void free_fft_tables(Plan& p);
void fourier_path_stats(Plan& p, long long* out8);  // grid points / exchange rows of this rank's band per Fourier path
void clone_fft_state(Plan& src, Plan& dst);   // launch groups / per-pair metadata of a plan that borrows src's tables
// fields < nb_uv are multiplied by d_scale[latitude pair] in the store (default: 1 / cos(lat), the wind scaling)
// chunk >= 0: only the fields of that chunk of the split set up by fourier_set_chunks (host-pointer pipelines)
int launch_fourier_inv(Plan& p, int nf, int mlimit, const double* d_fourier, double* d_gp, int nb_uv,
                       const double* d_scale = nullptr, int chunk = -1);
int launch_fourier_dir(Plan& p, int nf, const double* d_gp, double* d_fourier, int nb_uv, int adjoint = 0, int chunk = -1);
// split the fields of the next Fourier-stage launches into `nchunks` contiguous chunks; returns the nchunks + 1 field bounds
int fourier_set_chunks(Plan& p, int nf, int nchunks, std::vector<int>* field_bounds);
// sharded direct transform: every output row is stored into the exchange buffer of the rank that owns its zonal
// wavenumber (NVLink stores).  *fused = false if the plan has row-mode groups, which write the local buffer only
// (the caller then pushes the rows with launch_exchange_push)
int launch_fourier_dir_peers(Plan& p, int nf, const double* d_gp, const PeerDst& dst, bool* fused);

// ---- exchange.cu ----
int launch_exchange_copy(Plan& p, int nf, const ExSeg* d_segs, int nseg, double* d_fourier, double* d_buf, bool gather);
PeerDst make_peer_dst(const Plan& p);      // exchange buffers of the current parity on every rank
int launch_exchange_push(Plan& p, int nf);  // band-side rows of the local buffer -> owners of their zonal wavenumber
int launch_peer_barrier(Plan& p);

// ---- points.cu ----
// Fourier stage of a point-set plan: gp[f][ip] = sum_m c_m Re((S_m +- A_m)(lat of ip) e^{i m lon_ip}), m <= mlimit
int launch_points_inv(Plan& p, int nf, int mlimit, const double* d_fourier, double* d_gp, int nb_uv);

// ---- fields.cu ----
// atlas Field layout (node, level, component) <-> transform rows [component * nlev + level][node]
int launch_gp_repack(Plan& p, int nlev, int ncomp, const double* d_in, double* d_out, bool to_rows);
// cropped plans: band-layout rows (Fourier stage output) -> the crop's points [field][crop point]
int upload_crop_rows(Plan& p);
int launch_crop_gather(Plan& p, int nf, const double* d_band, double* d_gp);

// ---- vordiv.cu ----
int launch_grad_spectra(cudaStream_t s, int T, int nf, const double* d_sp, double* d_all, uint64_t* launches);
int launch_uv_to_vordiv(cudaStream_t s, int T, int nf, const long long* d_sp_rowoff, const double* d_packed,
                        double* d_vor, double* d_div, uint64_t* launches);
int launch_merge_uv_scalar(cudaStream_t s, int T, int nvd, int nsc, const double* d_vor, const double* d_div,
                           const double* d_sc, double* d_all, uint64_t* launches);
// adjoints of the two spectral operators above, reading the packed output of the direct Legendre kernel at T+1
int launch_merge_uv_scalar_adj(cudaStream_t s, int T, int nvd, int nsc, const long long* d_sp_rowoff, const double* d_packed,
                               double* d_vor, double* d_div, double* d_sc, uint64_t* launches);
int launch_grad_spectra_adj(cudaStream_t s, int T, int nf, const long long* d_sp_rowoff, const double* d_packed, double* d_sp,
                            uint64_t* launches);
// transpose of launch_uv_to_vordiv: (vor, div) adjoint variables at T -> [Ut | Vt] adjoint variables at T+1, [m][n][re/im][field]
int launch_uv_to_vordiv_adj(cudaStream_t s, int T, int nf, const double* d_vor, const double* d_div, double* d_all,
                            uint64_t* launches);
int launch_vd2uv(cudaStream_t s, int T, int nf, const double* d_vor, const double* d_div, double* d_U, double* d_V,
                 uint64_t* launches);

}  // namespace sptrans

// the opaque handle of the C ABI
struct sptrans_plan {
    sptrans::Plan p;
};
