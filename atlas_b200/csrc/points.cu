// Transform to arbitrary points (TransLocal with an UnstructuredGrid, ecmwf/atlas src/atlas/trans/local/TransLocal.cc:
// invtrans_unstructured :1289-1392, the default path; invtrans_unstructured_precomp :1200-1285).
//
// The reference evaluates, for every point, all Legendre functions at the point's latitude, 2(T+1) GEMMs with one column,
// and the Fourier sum as a dot product with (1, 0, 2 cos(m lon), -2 sin(m lon), ...) (:1347-1362).  Here a point-set plan
// (sptrans_plan_create_points) keeps one table row per DISTINCT |latitude|, so the Legendre stage is the same tensor-pipe
// GEMM as for grids -- points that share a latitude (regional lon-lat boxes: every row) share its output -- and the
// Fourier sums of all points are this one kernel: a block per point, a thread per field,
//     gp[f][ip] = sum_{m <= mlimit} c_m Re( (S_m +- A_m)(|lat_ip|)[f] e^{i m lon_ip} ),   c_0 = 1, c_m = 2,
// reading the symmetric / antisymmetric parts S, A straight from the Legendre <-> Fourier exchange buffer (the southern
// hemisphere flips the antisymmetric part), Im of m = 0 ignored (:1351), u, v = U, V / cos(lat) for the wind rows (:1380).
#include "plan.hpp"

namespace sptrans {

namespace {

constexpr int kPtThreads = 128;

__global__ void __launch_bounds__(kPtThreads)
points_inv_kernel(int nf, int mlimit, int nleg, int nb_uv, long long npts, const int* __restrict__ nlat0,
                  const long long* __restrict__ fb_rowoff, const int* __restrict__ pt_row,
                  const double* __restrict__ pt_sign, const double* __restrict__ pt_lon,
                  const double* __restrict__ pt_coslatinv, const double2* __restrict__ fb, double* __restrict__ gp) {
    extern __shared__ double2 cs[];  // e^{i m lon}, m = 0..mlimit
    const long long ip = blockIdx.x;
    const double lon = pt_lon[ip];
    for (int m = threadIdx.x; m <= mlimit; m += kPtThreads) {
        double s, c;
        sincos(m * lon, &s, &c);   // the reference's own expression: std::cos(jm * lon), std::sin(jm * lon)
        cs[m] = make_double2(c, s);
    }
    __syncthreads();
    const int jj = pt_row[ip];
    const double sg = pt_sign[ip];
    for (int f = threadIdx.x; f < nf; f += kPtThreads) {
        double acc = 0.;
        for (int m = 0; m <= mlimit; ++m) {
            const int n0 = nlat0[m];
            if (jj < n0) continue;  // zonal wavenumber not carried at this row (never for point-set plans: nlat0 = 0)
            const long long rs = fb_rowoff[m] + (jj - n0);
            const long long ra = rs + (nleg - n0);
            const double2 S = fb[rs * nf + f], A = fb[ra * nf + f];
            const double xr = S.x + sg * A.x, xi = S.y + sg * A.y;
            if (m == 0) acc += xr;
            else acc += 2. * (xr * cs[m].x - xi * cs[m].y);
        }
        if (f < nb_uv) acc *= pt_coslatinv[ip];
        gp[static_cast<long long>(f) * npts + ip] = acc;
    }
}

}  // namespace

int launch_points_inv(Plan& p, int nf, int mlimit, const double* d_fourier, double* d_gp, int nb_uv) {
    const long long npts = p.g.npts;
    if (npts == 0 || nf == 0) return SPTRANS_OK;
    const size_t smem = static_cast<size_t>(mlimit + 1) * sizeof(double2);
    if (smem > 48 * 1024)
        SPT_CUDA(cudaFuncSetAttribute(points_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    points_inv_kernel<<<static_cast<unsigned>(npts), kPtThreads, smem, p.stream>>>(
        nf, mlimit, p.g.nleg, nb_uv, npts, p.d_nlat0, p.d_fb_rowoff, p.d_pt_row, p.d_pt_sign, p.d_pt_lon, p.d_pt_coslatinv,
        reinterpret_cast<const double2*>(d_fourier), d_gp);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

}  // namespace sptrans
