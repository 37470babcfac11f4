/*
 * sptrans_b200.h -- C ABI of the B200-native spherical-harmonics transform engine.
 *
 * This is the drop-in boundary below atlas::trans::TransImpl: plain pointers and sizes, no
 * atlas / eckit / torch types.  Every entry point names the reference interface it replaces
 * (paths relative to ecmwf/atlas src/atlas/).  The adaptor class that registers this engine
 * as `type("b200")` with atlas's TransFactory is include/atlas_b200/TransB200.h; the
 * binding a maintainer would add on the reference side is shown in INTEGRATION.md.
 *
 * Conventions (identical to the reference):
 *   spectra : [m][n][re/im][field], field fastest; (T+1)(T+2) doubles per field
 *             (functionspace/Spectral.h:43-61, trans/local/TransLocal.cc:970-981)
 *   grid    : [field][point]; points ordered by latitude row north->south, longitudes
 *             lambda_i = 2 pi i / nx(row) from 0 (trans/local/TransLocal.cc:1160-1187)
 *   wind    : all u fields, then all v fields (then scalars)   (TransLocal.cc:1561-1589)
 * All data pointers may be device pointers (used in place) or host pointers (staged through
 * the plan's device workspace, the grid-point side in field chunks on copy streams next to the Fourier kernels); the engine detects which with
 * cudaPointerGetAttributes, as atlas itself does in parallel/detail/DevicePacker.hic:20-28.
 * There is NO CPU fallback: every call fails with SPTRANS_ERR_CUDA if no device is present.
 *
 * Thread-safety: like TransLocal (mutable scratch, trans/local/TransLocal.h:245) a plan
 * supports one in-flight call at a time.  Calls are blocking: results are visible on return.
 */
#ifndef SPTRANS_B200_H
#define SPTRANS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sptrans_plan sptrans_plan;

enum {
    SPTRANS_OK = 0,
    SPTRANS_ERR_INVALID = 1,        /* bad argument (eckit::BadParameter / ATLAS_ASSERT in the reference) */
    SPTRANS_ERR_CUDA = 2,           /* CUDA runtime / launch failure, or no device */
    SPTRANS_ERR_NOT_IMPLEMENTED = 3 /* ATLAS_NOTIMPLEMENTED in the reference */
};

/* flags for sptrans_plan_create */
enum {
    SPTRANS_GRID_REGULAR = 1u,  /* RegularGrid(grid): linear zonal truncation (TransLocal.cc:281-284) */
    SPTRANS_NO_FP64_TABLE = 2u, /* keep only the split-integer Legendre table (saves HBM; fp64 kernels unavailable) */
    /* sharded plans: the spectral / grid-point arrays of the sharded entry points hold ONLY this rank's share --
     * spectra [my zonal wavenumbers, ascending][n = m..T][re/im][field], grid fields [field][rows of my latitude
     * band: northern rows north->south, then their southern mirrors north->south] (stride: sptrans_local_sizes).
     * Without it every rank passes full-size arrays and touches only its share (36 GB per rank at TCo2559 L137). */
    SPTRANS_SHARD_LOCAL_IO = 4u
};

/* precision selector for sptrans_set_precision */
enum {
    SPTRANS_PREC_FP64 = 0,  /* DMMA fp64 tensor path (BASELINE configs 1-3,5) */
    SPTRANS_PREC_TC_SPLIT = 1 /* tcgen05 split-operand tensor-core path (BASELINE config 4) */
};

/* Last error message of the calling thread (errors in the reference are C++ exceptions,
 * runtime/Exception.h:23-76; the TransB200 adaptor rethrows this string). */
const char* sptrans_last_error(void);

/* Library / device probe: returns the number of visible CUDA devices (0 => every other call fails). */
int sptrans_device_count(void);

/* ---- grid helpers (replace the Grid getters TransLocal's constructor consumes) ------------------- */

/* Gaussian latitudes (degrees, north->south, 2N values) and quadrature weights (sum == 1).
 * Replaces grid/detail/spacing/gaussian/Latitudes.cc:227-274 (compute_gaussian_quadrature_npole_equator). */
int sptrans_gaussian_latitudes(int N, double* lat_deg_2N, double* weights_2N);

/* Octahedral reduced Gaussian grid O<N>: nx[j] = 20 + 4 j mirrored (grid/detail/grid/Gaussian.cc:127-134). */
int sptrans_octahedral_nx(int N, int* nx_2N);

/* Highest zonal wavenumber kept at a latitude (trans/local/TransLocal.cc:272-300 fourier_truncation). */
int sptrans_fourier_truncation(int truncation, int nx, int nxmax, int ndgl, double lat_rad, int fullgrid);

/* ---- plan = what TransLocal::TransLocal builds (trans/local/TransLocal.cc:322-770) ---------------- */

/* Create a transform plan for a GLOBAL structured grid given as rows.
 *   nlat          number of latitude rows (StructuredGrid::ny)
 *   nx[nlat]      points per row (StructuredGrid::nx(j))
 *   lat_deg[nlat] row latitudes in degrees, north->south, symmetric about the equator
 *   weights[nlat] quadrature weights (sum 1) for the direct transform, or NULL (=> dirtrans unavailable)
 *   truncation    spectral truncation T
 *   device        CUDA device ordinal
 * Replaces TransBuilderGrid<TransLocal>::make -> TransLocal ctor (trans/detail/TransFactory.h:114-119).
 * Legendre tables are generated ON THE DEVICE (bit-identical to trans/local/LegendrePolynomials.cc). */
int sptrans_plan_create(sptrans_plan** plan, int nlat, const int* nx, const double* lat_deg, const double* weights,
                        int truncation, unsigned flags, int device);

/* m-sharded plan for multi-GPU runs (one process per GPU): this rank owns the zonal wavenumbers
 * { m : owner(m) == rank } (work-balanced pairing) and the latitude pairs of band `rank`.
 * No reference equivalent: TransLocal throws for mpi::size()>1 (TransLocal.cc:338-340). */
int sptrans_plan_create_sharded(sptrans_plan** plan, int nlat, const int* nx, const double* lat_deg,
                                const double* weights, int truncation, unsigned flags, int device, int rank,
                                int nranks);

/* Trans(global_grid, domain, truncation) with a NON-global domain: the regional grid is a cropping of a global structured
 * grid (TransLocal.cc:371-531; anything else is refused there, :410-420).  Rows jlat_min .. jlat_min + nlat_crop - 1 of the
 * global grid are kept; crop row r holds nx_crop[r] consecutive longitudes of its global row starting at index jlon_min[r]
 * and wrapping around (jlonMin_, :501-531; copy-out :1180-1187).  The transform uses the GLOBAL grid's per-latitude zonal
 * truncation (nlat0_, :462-488) and FFTs of the global row lengths, exactly like the reference; grid-point arrays of the
 * plan are [field][crop point].  Inverse transforms only (scalar, vor/div -> wind, general, gradient), like TransLocal. */
int sptrans_plan_create_cropped(sptrans_plan** plan, int nlat, const int* nx, const double* lat_deg, int truncation,
                                unsigned flags, int device, int jlat_min, int nlat_crop, const int* nx_crop,
                                const int* jlon_min);

/* Trans(UnstructuredGrid, truncation): a plan for `npoints` arbitrary points (lon, lat in degrees), the counterpart of
 * TransLocal's unstructured path (TransLocal.cc:740-770 set-up, :1289-1392 invtrans_unstructured).  Grid-point arrays of
 * such a plan are [field][point]; every zonal wavenumber m <= T enters at every point (no reduced-grid truncation, and
 * unlike the structured path the m == T column of a scalar call is kept, :1331); u, v = U, V / cos(lat) without pole
 * clamp (:1380-1384).  Inverse transforms only (scalar, vor/div -> wind, general, gradient): like TransLocal there is no
 * direct transform from scattered points (SPTRANS_ERR_NOT_IMPLEMENTED).  The Legendre stage runs once per DISTINCT
 * |latitude| on the tensor pipe, so regional lon-lat boxes (TransLocal.cc:397-407: "no_nest" regular grids, which the
 * reference also evaluates without FFT) cost one table row per latitude, not per point. */
int sptrans_plan_create_points(sptrans_plan** plan, size_t npoints, const double* lon_deg, const double* lat_deg,
                               int truncation, int device);
int sptrans_plan_destroy(sptrans_plan* plan); /* atlas__Trans__delete, trans/detail/TransInterface.cc:86-89 */

/* inspectors: TransImpl::truncation / grid().size() / nb_spectral_coefficients (TransLocal.h:84-91) */
int sptrans_truncation(const sptrans_plan* plan);
size_t sptrans_nb_gridpoints(const sptrans_plan* plan);
size_t sptrans_nb_spectral_coefficients(const sptrans_plan* plan); /* (T+1)(T+2) */
/* first northern latitude index at which wavenumber m is resolved (TransLocal.cc:462-488); out[T+1] */
int sptrans_get_nlat0(const sptrans_plan* plan, int* nlat0);
/* per-field sizes of the arrays the (sharded) entry points address: doubles of the spectral array, points of the grid
 * array.  (T+1)(T+2) and nb_gridpoints unless the plan was created with SPTRANS_SHARD_LOCAL_IO. */
int sptrans_local_sizes(const sptrans_plan* plan, size_t* spec_doubles_per_field, size_t* gridpoints_per_field);
/* bytes of device memory held by the plan (tables + workspaces) */
size_t sptrans_device_bytes(const sptrans_plan* plan);

/* Export the Legendre tables in the reference's cache layout [all sym blocks][all asym blocks], blocks
 * padded to 8 doubles, n descending (TransLocal.cc:592-647, LegendrePolynomials.cc:181-205).
 * `out` is a HOST buffer of sptrans_legendre_cache_size() bytes. */
size_t sptrans_legendre_cache_size(const sptrans_plan* plan);
int sptrans_export_legendre_cache(const sptrans_plan* plan, void* out);
/* Load the Legendre tables from a blob in that same layout -- a cache file written by the reference
 * (LegendreCacheCreatorLocal::create, trans/local/LegendreCacheCreatorLocal.cc:137-146; TransLocal reads it at
 * TransLocal.cc:608-647 when Cache::legendre() is set) or by sptrans_export_legendre_cache.  `blob` is a HOST buffer of
 * exactly sptrans_legendre_cache_size() bytes (anything else is SPTRANS_ERR_INVALID, like the reference's size
 * assertion TransLocal.cc:612).  A plan always generates its own tables at creation (30 ms at T1279 on the device,
 * faster than reading the 8.4 GB file); importing replaces them, e.g. to reproduce a run of the reference bit for bit. */
int sptrans_import_legendre_cache(sptrans_plan* plan, const void* blob, size_t bytes);

/* LegendreCacheCreatorLocal::uid (trans/local/LegendreCacheCreatorLocal.cc:66-119): the string that names a Legendre
 * cache file.  `kind` is what the reference derives from the Grid type: a global Gaussian grid ("GaussianN<N>"), a global
 * regular lon-lat grid with / without pole rows ("L-ny<ny>" / "S-ny<ny>"), a regional regular grid with linear y-spacing
 * ("Regional-south<ymin>-north<ymax>-ny<ny>"), anything else ("grid-" + first 10 hex digits of eckit::MD5 over
 * lround(lat_j * 1e8) of the nlat rows).  The option hash is eckit::MD5("flt", bool)[0:10].  prefix "local" reproduces the
 * reference's strings exactly (the 66 fixtures of src/tests/trans/test_trans_localcache.cc:264-360 are a CPU test here);
 * this backend's creator uses the same prefix on purpose: its tables are bit-identical to TransLocal's, so cache files are
 * interchangeable between the two backends.  Returns the length written (without the terminator), or -1. */
enum { SPTRANS_UID_GAUSSIAN = 0, SPTRANS_UID_LONLAT = 1, SPTRANS_UID_SHIFTED_LONLAT = 2, SPTRANS_UID_REGIONAL = 3, SPTRANS_UID_OTHER = 4 };
int sptrans_legendre_cache_uid(char* out, size_t out_len, const char* prefix, int truncation, int kind, int n_or_ny,
                               double south, double north, int nlat, const double* lat_deg, int flt);
/* LegendreCacheCreatorLocal::estimate (:148-150): T^3 / 2 * 8 bytes */
size_t sptrans_legendre_cache_estimate(int truncation);

/* Select the arithmetic of the Legendre stage: SPTRANS_PREC_FP64 (default; DMMA, results match the fp64 oracle to
 * 1e-13) or SPTRANS_PREC_TC_SPLIT (tcgen05 kind::tf32 with split operands and fp32 accumulation in tensor memory:
 * fp32-level accuracy, BASELINE config 4).  The reference has one precision only (double, eckit gemm). */
int sptrans_set_precision(sptrans_plan* plan, int precision);

/* run all subsequent calls on this cudaStream_t (default: the plan's own stream) */
int sptrans_set_stream(sptrans_plan* plan, void* cuda_stream);

/* Asynchronous calls (the optional explicit-stream variant of SURVEY 8b "Threading"; the reference is blocking only).
 * With sptrans_set_async(plan, 1) the whole-transform entry points (invtrans*, dirtrans*, invtrans_grad) return as soon
 * as their work is enqueued on the plan's streams; host buffers must be page-locked for the copies to be asynchronous and
 * must not be touched until sptrans_synchronize(plan) returns.  Calls on one plan still execute in issue order.
 * sptrans_plan_clone gives a second plan on the same device that BORROWS the tables of `src` (Legendre tables, Fourier
 * tables, geometry) and owns only its streams and workspaces: one transform can be in flight on each, e.g. an inverse
 * and a direct transform whose host<->device copies then use both directions of the PCIe link at once.  Clones must be
 * destroyed before `src`.
 * Dependencies between asynchronous calls on DIFFERENT plans (e.g. the direct transform on the clone reads the grid
 * fields the inverse on the original is still writing to host memory) are expressed on the device, without host
 * synchronisation: sptrans_mark(a, &m) marks the end of everything enqueued on plan a so far, sptrans_wait_mark(b, m)
 * makes the calls issued on plan b from now on wait for it, sptrans_release_mark(m) frees the mark (at any time). */
int sptrans_set_async(sptrans_plan* plan, int on);
int sptrans_synchronize(sptrans_plan* plan);
int sptrans_mark(sptrans_plan* plan, void** mark);
int sptrans_wait_mark(sptrans_plan* plan, void* mark);
int sptrans_release_mark(void* mark);
int sptrans_plan_clone(sptrans_plan* src, sptrans_plan** clone);

/* ---- transforms ------------------------------------------------------------------------------------ */

/* TransImpl::invtrans(nb_scalar_fields, scalar_spectra, gp_fields)            trans/detail/TransImpl.h:136-137
 * = TransLocal::invtrans -> invtrans_uv(truncation_, ...)                      TransLocal.cc:931-934          */
int sptrans_invtrans_scalar(sptrans_plan* plan, int nb_fields, const double* scalar_spectra, double* gp_fields);

/* TransImpl::invtrans(nb_scalar, scalar_spectra, nb_vordiv, vor, div, gp)      trans/detail/TransImpl.h:125-127
 * = TransLocal::invtrans (extend_truncation, vd2uv, merged spectra at T+1)     TransLocal.cc:1523-1597
 * gp layout: [u_1..u_k | v_1..v_k | s_1..s_j][npts]                                                            */
int sptrans_invtrans(sptrans_plan* plan, int nb_scalar_fields, const double* scalar_spectra, int nb_vordiv_fields,
                     const double* vorticity_spectra, const double* divergence_spectra, double* gp_fields);

/* TransImpl::invtrans(nb_vordiv, vor, div, gp)                                 trans/detail/TransImpl.h:144-146 */
int sptrans_invtrans_vordiv2wind(sptrans_plan* plan, int nb_vordiv_fields, const double* vorticity_spectra,
                                 const double* divergence_spectra, double* gp_fields);

/* TransImpl::dirtrans(nb_fields, scalar_fields, scalar_spectra)                trans/detail/TransImpl.h:172-173
 * NotImplemented in TransLocal (TransLocal.cc:1671-1676); semantics of TransIFS (ifs/TransIFS.cc:503-515).  */
int sptrans_dirtrans_scalar(sptrans_plan* plan, int nb_fields, const double* gp_fields, double* scalar_spectra);


/* TransImpl::invtrans_grad(spfield, gradfield) in IFS-style pointers: grad layout
 * [E-W_1..E-W_k | N-S_1..N-S_k][npts], i.e. component 0 = (1/(a cos))d/dlambda, 1 = (1/a)d/dphi
 * (ifs/TransIFS.cc:2075-2142).  NotImplemented in TransLocal (TransLocal.cc:848-857).                          */

/* TransImpl::invtrans_adj(nb_scalar_fields, gp_fields, scalar_spectra)            trans/detail/TransImpl.h:155-157
 * adjoint of sptrans_invtrans_scalar in the ectrans / TransIFS convention: <invtrans x, y>_grid = <x, invtrans_adj y>_spec
 * with the spectral inner product that counts m > 0 coefficients twice (test_transgeneral.cc:1683-1686).
 * NotImplemented in TransLocal (TransLocal.cc:1599-1604); semantics of the adjoint tests test_transgeneral.cc:1591-1818. */
int sptrans_invtrans_adj_scalar(sptrans_plan* plan, int nb_fields, const double* gp_fields, double* scalar_spectra);

/* TransImpl::dirtrans(nb_fields, wind_fields, vorticity_spectra, divergence_spectra)   TransImpl.h:180-181
 * wind layout [u_1..u_k | v_1..v_k][npts] (what invtrans_vordiv2wind produces); NotImplemented in TransLocal
 * (TransLocal.cc:1680-1685).  Exact inverse of sptrans_invtrans_vordiv2wind up to quadrature error.          */
int sptrans_dirtrans_wind2vordiv(sptrans_plan* plan, int nb_fields, const double* wind_fields,
                                 double* vorticity_spectra, double* divergence_spectra);

/* TransImpl::invtrans_grad(spfield, gradfield) in IFS-style pointers: grad layout
 * [E-W_1..E-W_k | N-S_1..N-S_k][npts], i.e. component 0 = (1/(a cos))d/dlambda, 1 = (1/a)d/dphi
 * (ifs/TransIFS.cc:2075-2142).  NotImplemented in TransLocal (TransLocal.cc:848-857).                          */
int sptrans_invtrans_grad(sptrans_plan* plan, int nb_fields, const double* scalar_spectra, double* grad_fields);

/* ---- adjoints.  ATLAS_NOTIMPLEMENTED in TransLocal (TransLocal.cc:899-929, :1599-1667); semantics of the reference's
 * adjoint tests (src/tests/trans/test_transgeneral.cc:1591-1818), which TransIFS / ectrans pass: <A x, y>_grid =
 * <x, A* y>_spec with the Euclidean sum over grid points and a spectral inner product that counts every stored m > 0
 * coefficient twice (:1683-1686, :1790-1793).  I.e. invtrans_adj = C^-1 A^T, dirtrans_adj = B^T C, C = diag(1 | 2). --- */

/* TransImpl::invtrans_adj(nb_scalar_fields, gp_fields, nb_vordiv_fields, vor, div, scalar_spectra)  TransImpl.h:147-149
 * adjoint of sptrans_invtrans: gp layout [u_1..u_k | v_1..v_k | s_1..s_j][npts] in, spectra at truncation T out. */
int sptrans_invtrans_adj(sptrans_plan* plan, int nb_scalar_fields, const double* gp_fields, int nb_vordiv_fields,
                         double* vorticity_spectra, double* divergence_spectra, double* scalar_spectra);
/* TransImpl::invtrans_adj(nb_vordiv_fields, wind_fields, vor, div)                                  TransImpl.h:165-166
 * = atlas__Trans__invtrans_vordiv2wind_adj (trans/detail/TransInterface.h:66-67) */
int sptrans_invtrans_vordiv2wind_adj(sptrans_plan* plan, int nb_vordiv_fields, const double* wind_fields,
                                     double* vorticity_spectra, double* divergence_spectra);
/* TransImpl::invtrans_grad_adj(gradfield, spfield) in IFS-style pointers (TransImpl.h:93-97): adjoint of
 * sptrans_invtrans_grad, grad layout [E-W_1..E-W_k | N-S_1..N-S_k][npts] in. */
int sptrans_invtrans_grad_adj(sptrans_plan* plan, int nb_fields, const double* grad_fields, double* scalar_spectra);
/* TransImpl::dirtrans_adj(spfield, gpfield) in IFS-style pointers (TransImpl.h:63-67): adjoint of
 * sptrans_dirtrans_scalar -- spectra in, grid fields out (needs quadrature weights). */
int sptrans_dirtrans_adj_scalar(sptrans_plan* plan, int nb_fields, const double* scalar_spectra, double* gp_fields);
/* TransImpl::dirtrans_wind2vordiv_adj(spvor, spdiv, gpwind) in IFS-style pointers (TransImpl.h:69-70; the reference's
 * test_2level_adjoint_test_with_vortdiv, test_transgeneral.cc:1725-1818): adjoint of sptrans_dirtrans_wind2vordiv --
 * vor/div spectra at truncation T in, wind fields [u_1..u_k | v_1..v_k][npts] out (needs quadrature weights). */
int sptrans_dirtrans_wind2vordiv_adj(sptrans_plan* plan, int nb_fields, const double* vorticity_spectra,
                                     const double* divergence_spectra, double* wind_fields);

/* ---- atlas Field layouts (multi-level Fields).  Entry points behind TransImpl's Field overloads
 * (trans/detail/TransImpl.h:54-100; C bindings atlas__Trans__*_field, trans/detail/TransInterface.h:72-96).
 * A spectral Field (nspec2, nlev) IS the raw [coeff][field] layout.  A grid-point Field is (node, level) for scalars and
 * (node, level, 2) for wind / gradient, LAST index fastest (owned nodes 0..npts-1 in grid order,
 * functionspace/detail/StructuredColumns.h:104-137): the transpose of the [field][node] rows above, with
 * field = component * nlev + level exactly as TransIFS packs them on the host (trans/ifs/TransIFS.cc:610-667,
 * :1392-1437, :2113-2137).  Here the repack is one tiled transpose on the device; pointers may be host or device
 * (an atlas Field whose device copy is valid passes array().device_data(), array/Array.h:177-183). ----------------- */
int sptrans_invtrans_field(sptrans_plan* plan, int nlev, const double* spfield, double* gpfield);
int sptrans_dirtrans_field(sptrans_plan* plan, int nlev, const double* gpfield, double* spfield);
int sptrans_invtrans_adj_field(sptrans_plan* plan, int nlev, const double* gpfield, double* spfield);
int sptrans_invtrans_vordiv2wind_field(sptrans_plan* plan, int nlev, const double* spvor, const double* spdiv,
                                       double* gpwind /* (npts, nlev, 2): component 0 = u, 1 = v */);
int sptrans_dirtrans_wind2vordiv_field(sptrans_plan* plan, int nlev, const double* gpwind, double* spvor,
                                       double* spdiv);
int sptrans_invtrans_grad_field(sptrans_plan* plan, int nlev, const double* spfield,
                                double* gradfield /* (npts, nlev, 2): component 0 = E-W, 1 = N-S */);
int sptrans_dirtrans_adj_field(sptrans_plan* plan, int nlev, const double* spfield, double* gpfield);
int sptrans_invtrans_vordiv2wind_adj_field(sptrans_plan* plan, int nlev, const double* gpwind, double* spvor,
                                           double* spdiv);
int sptrans_invtrans_grad_adj_field(sptrans_plan* plan, int nlev, const double* gradfield, double* spfield);
int sptrans_dirtrans_wind2vordiv_adj_field(sptrans_plan* plan, int nlev, const double* spvor, const double* spdiv,
                                           double* gpwind);

/* VorDivToUV::execute(nb_coeff, nb_fields, vor, div, U, V)   trans/VorDivToUV.h:121-122,
 * = vd2uv, trans/local/VorDivToUVLocal.cc:62-184.  Plan-free: spectral space only.                            */
int sptrans_vordiv_to_uv(int truncation, int nb_fields, const double* vorticity, const double* divergence,
                         double* U, double* V, int device);

/* ---- stage-level entry points (multi-GPU pipelines, benchmarks, tests) ---------------------------- */

/* number of double2 (re,im) elements per field of the Legendre<->Fourier exchange buffer */
size_t sptrans_fourier_elems_per_field(const sptrans_plan* plan);
/* Which Fourier kernels serve the latitude rows of this plan (of this rank's band): out8[0..3] = grid points per field,
 * out8[4..7] = exchange-buffer rows (one (re,im) pair per field each) on 0: the direct mixed-radix kernels (row length
 * without prime factors above 23), 1: the register-tiled chirp-z kernels, 2: the shared-memory-pass chirp-z kernels,
 * 3: the row-mode chirp-z kernels (rows beyond the single-CTA limit).  The reference has one FFTW plan per row length
 * (TransLocal.cc:1122-1150); here the row length decides the algorithm.  Zero for point-set plans. */
int sptrans_fourier_path_stats(sptrans_plan* plan, long long* out8);
/* Legendre stage only: device spectra -> device Fourier buffer  (TransLocal::invtrans_legendre, :939-1097) */
int sptrans_invtrans_legendre(sptrans_plan* plan, int nb_fields, int truncation_of_data, const double* d_spectra,
                              double* d_fourier);
/* Fourier stage only: device Fourier buffer -> device grid     (TransLocal::invtrans_fourier_*, :1101-1196) */
int sptrans_invtrans_fourier(sptrans_plan* plan, int nb_fields, int mlimit, const double* d_fourier,
                             double* d_gp, int nb_uv_fields);
int sptrans_dirtrans_fourier(sptrans_plan* plan, int nb_fields, const double* d_gp, double* d_fourier,
                             int nb_uv_fields);
int sptrans_dirtrans_legendre(sptrans_plan* plan, int nb_fields, const double* d_fourier, double* d_spectra);

/* ---- multi-GPU: one exchange step between the two shardings (no reference equivalent; ectrans does the same
 * transposition internally over MPI).  "m side" = rows of the exchange buffer this rank owns as owner of its
 * zonal wavenumbers, grouped by destination latitude band; "band side" = rows it owns as owner of its latitude
 * band, grouped by the rank owning the zonal wavenumber.  Row = nb_fields (re,im) pairs.
 *   inverse: legendre -> pack(m side) -> all-to-all (send m_side_rows, recv band_side_rows) -> unpack(band side) -> fourier
 *   direct : fourier  -> pack(band side) -> all-to-all (send band_side_rows, recv m_side_rows) -> unpack(m side) -> legendre */
/* host-only layout query (no device needed): owner[T+1], band[nranks+1], rows per peer [nranks] each */
int sptrans_shard_layout(int nlat, const int* nx, const double* lat_deg, int truncation, unsigned flags, int rank,
                         int nranks, int* owner, int* band, long long* m_side_rows, long long* band_side_rows);
/* host-only: the copy segments of one side as triples (exchange-buffer row, packed-buffer row, number of rows);
 * returns the number of segments (call with out == NULL to size the buffer) */
long long sptrans_shard_segments(int nlat, const int* nx, const double* lat_deg, int truncation, unsigned flags,
                                 int rank, int nranks, int side, long long* out);
int sptrans_exchange_rows(const sptrans_plan* plan, long long* m_side_rows, long long* band_side_rows);
int sptrans_exchange_pack(sptrans_plan* plan, int nb_fields, int side, const double* d_fourier, double* d_buf);
int sptrans_exchange_unpack(sptrans_plan* plan, int nb_fields, int side, const double* d_buf, double* d_fourier);


/* ---- peer-memory exchange (default multi-GPU path on one NVLink / NVSwitch node, <= 8 ranks) ----------------
 * Every rank allocates one region [flags | exchange buffer 0 | exchange buffer 1] and maps the regions of all
 * peers (CUDA IPC: the 64-byte handles travel over whatever the host uses -- MPI in atlas, torch.distributed
 * here; or plain device pointers when the host already shares memory).  After that a sharded transform is ONE
 * stream-ordered call per rank, with no host synchronisation and no collective library:
 *   inverse: the Legendre GEMM stores every output row straight into the Fourier-side buffer of the rank that
 *            owns its latitude band (NVLink stores from the GEMM epilogue) -> device-side barrier -> Fourier;
 *   direct : Fourier -> rows pushed to the owners of their zonal wavenumber -> device-side barrier -> Legendre.
 * Calls are collective: every rank must issue the same sequence.  d_* are device pointers; the grid-point
 * array is full size, only the rows of this rank's latitude band are read/written; the spectral array is full
 * size, only this rank's zonal wavenumbers are read/written. */
#define SPTRANS_IPC_HANDLE_BYTES 64
int sptrans_peer_alloc(sptrans_plan* plan, int nb_fields, unsigned char* ipc_handle_out /* 64 bytes or NULL */);
int sptrans_peer_attach_ipc(sptrans_plan* plan, int nranks, const unsigned char* handles /* nranks x 64 bytes */);
int sptrans_peer_attach_ptrs(sptrans_plan* plan, int nranks, void* const* regions /* from sptrans_peer_region */);
int sptrans_peer_region(const sptrans_plan* plan, void** region, size_t* bytes);
int sptrans_peer_buffer(const sptrans_plan* plan, double** d_fourier);   /* local exchange buffer in use */
int sptrans_peer_free(sptrans_plan* plan);
int sptrans_invtrans_sharded(sptrans_plan* plan, int nb_fields, const double* d_spectra, double* d_gp);
int sptrans_dirtrans_sharded(sptrans_plan* plan, int nb_fields, const double* d_gp, double* d_spectra);
/* the halves of the two calls above, for hosts that sequence the stages themselves (tests emulate N ranks on
 * one GPU this way): produce into the peers' buffers / barrier kernel / flip to the other buffer pair */
int sptrans_invtrans_legendre_peers(sptrans_plan* plan, int nb_fields, const double* d_spectra);
int sptrans_dirtrans_fourier_peers(sptrans_plan* plan, int nb_fields, const double* d_gp);   /* Fourier stage + push (round 1) */
/* direct transform, exchange by pull (sptrans_dirtrans_sharded with SPTRANS_DIR_PULL=1; measured slower than the push on real
 * NVLink, see api.cu): Fourier stage into the local exchange buffer; after the barrier the Legendre GEMM fetches every
 * Fourier row from the buffer of the rank that owns its latitude band */
int sptrans_dirtrans_fourier_local(sptrans_plan* plan, int nb_fields, const double* d_gp);
int sptrans_dirtrans_legendre_pull(sptrans_plan* plan, int nb_fields, double* d_spectra);
int sptrans_peer_barrier(sptrans_plan* plan);
int sptrans_peer_advance(sptrans_plan* plan);

/* ---- single-process multi-GPU plan (SURVEY 8e "Process model": one process, N devices).  One host thread drives N
 * sharded plans, one per device, whose exchange regions are mapped into each other with cudaDeviceEnablePeerAccess;
 * the transforms take GLOBAL arrays in the reference's layouts (host memory, or device memory under UVA), every device
 * is given only its share (SPTRANS_SHARD_LOCAL_IO), the device-side barrier of the peer-memory exchange is the only
 * synchronisation between the devices, and the calls are blocking like the reference's.  This is what the TransB200
 * adaptor selects with the config key "gpus" (include/atlas_b200/TransB200.h); one-process-per-GPU hosts use the
 * sptrans_peer_* / sptrans_*_sharded entry points above instead.  `devices` may be NULL (ordinals 0..ndevices-1) and may
 * name the same device more than once (tests emulate N ranks on one GPU that way). ------------------------------ */
typedef struct sptrans_multi sptrans_multi;
int sptrans_multi_create(sptrans_multi** multi, int nlat, const int* nx, const double* lat_deg, const double* weights,
                         int truncation, unsigned flags, int ndevices, const int* devices);
int sptrans_multi_destroy(sptrans_multi* multi);
int sptrans_multi_size(const sptrans_multi* multi);
sptrans_plan* sptrans_multi_plan(sptrans_multi* multi, int rank);   /* the sharded plan of one device (inspectors, timings) */
int sptrans_multi_invtrans_scalar(sptrans_multi* multi, int nb_fields, const double* scalar_spectra, double* gp_fields);
int sptrans_multi_dirtrans_scalar(sptrans_multi* multi, int nb_fields, const double* gp_fields, double* scalar_spectra);

/* kernel time of the last call's stages in milliseconds (CUDA events on the plan's stream):
 * out[0]=pack/unpack+vd2uv, out[1]=Legendre GEMM, out[2]=Fourier, out[3]=H2D, out[4]=D2H,
 * out[5]=wait at the exchange barrier (sharded calls; these read their events lazily here),
 * out[6]=Field-layout repack (the *_field entry points) */
int sptrans_last_timings(const sptrans_plan* plan, float out_ms[8]);
/* number of kernels this library launched on behalf of the plan since creation */
uint64_t sptrans_kernel_launches(const sptrans_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* SPTRANS_B200_H */
