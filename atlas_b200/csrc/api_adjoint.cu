// C ABI, adjoint transforms (TransImpl::invtrans_adj / invtrans_grad_adj / dirtrans_adj / dirtrans_wind2vordiv_adj,
// trans/detail/TransImpl.h:63-100, :147-166).  All of them are ATLAS_NOTIMPLEMENTED in TransLocal
// (trans/local/TransLocal.cc:899-929, :1599-1667); the semantics are those of the reference's adjoint tests, which TransIFS /
// ectrans pass at 1e-12 (src/tests/trans/test_transgeneral.cc:1591-1818): <A x, y>_grid = <x, A* y>_spec, where the grid
// inner product is the Euclidean sum over grid points and the SPECTRAL inner product counts every stored coefficient with
// m > 0 twice (`adj_value += (m1 > 0 ? 2 * temp : temp)`, :1683-1686, :1790-1793).  With C = diag(1 for m = 0, 2 for m > 0):
//   invtrans_adj = C^-1 (inverse)^T,   dirtrans_adj = (direct)^T C,   dirtrans_wind2vordiv_adj = (wind2vordiv)^T C.
// The same kernels run backwards:
//   C^-1 (inverse)^T = unpack o Legendre-direct GEMM o Fourier-direct kernel without quadrature weight and 1/nx (the factor
//                 2 of the m > 0 harmonics cancels against C^-1), the wind rows scaled by the inverse's own 1 / cos(lat),
//                 then the transposed spectral stencils (vordiv.cu; they act within one m, so they commute with C);
//   (direct)^T C = Fourier-inverse kernel o Legendre-inverse GEMM o pack, rows scaled by weight / nx (the inverse Fourier
//                 kernel's factor 2 for m > 0 is exactly C).
#include <algorithm>

#include "plan.hpp"

using namespace sptrans;

namespace {

int adj_args_ok(sptrans_plan* plan, const char* who) {
    if (!plan) {
        set_error(std::string(who) + ": null plan");
        return SPTRANS_ERR_INVALID;
    }
    SPT_CUDA(cudaSetDevice(plan->p.device));
    if (plan->p.g.points || plan->p.g.cropped) {
        set_error(std::string(who) + ": direct and adjoint transforms are not available for point-set and cropped-grid plans");
        return SPTRANS_ERR_NOT_IMPLEMENTED;
    }
    if (plan->p.g.nranks != 1) {
        set_error(std::string(who) + ": whole-transform entry points need an unsharded plan; use the stage-level API");
        return SPTRANS_ERR_INVALID;
    }
    if (plan->p.precision != SPTRANS_PREC_FP64) {
        set_error(std::string(who) + ": adjoints are implemented for the fp64 Legendre kernel");
        return SPTRANS_ERR_NOT_IMPLEMENTED;
    }
    return SPTRANS_OK;
}

size_t spec_doubles(int nf, int trunc) { return static_cast<size_t>(trunc + 1) * (trunc + 2) * nf; }
size_t packed_doubles(const Plan& p, int nf) { return static_cast<size_t>(p.g.sp_rowoff.back()) * 2 * nf; }
size_t fourier_doubles(const Plan& p, int nf) { return (static_cast<size_t>(p.g.fb_rowoff.back()) + kBM) * 2 * nf; }

// grid fields [nall][npts] -> packed adjoint variables at truncation T+1 (first nb_uv rows are wind rows)
int gp_to_packed_adj(Plan& p, int nall, int nb_uv, const double* gp, cudaEvent_t* ev) {
    const int T = p.g.T;
    const size_t ngp = static_cast<size_t>(p.g.npts) * nall;
    int rc;
    const double* d_gp = gp;
    cudaEventRecord(ev[0], p.stream);
    if (!is_device_pointer(gp)) {
        if ((rc = ensure(p.d_gp, p.gp_cap, ngp))) return rc;
        SPT_CUDA(cudaMemcpyAsync(p.d_gp, gp, ngp * sizeof(double), cudaMemcpyHostToDevice, p.stream));
        d_gp = p.d_gp;
    }
    if ((rc = build_tiles(p, nall, T, T + 1))) return rc;
    if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nall)))) return rc;
    if ((rc = ensure(p.d_fourier, p.fourier_cap, fourier_doubles(p, nall)))) return rc;
    cudaEventRecord(ev[1], p.stream);
    if ((rc = launch_fourier_dir(p, nall, d_gp, p.d_fourier, nb_uv, /*adjoint=*/1))) return rc;
    cudaEventRecord(ev[2], p.stream);
    if ((rc = launch_legendre_dir(p, nall, p.d_fourier, p.d_packed))) return rc;
    cudaEventRecord(ev[3], p.stream);
    return SPTRANS_OK;
}

int finish_timings(Plan& p, bool d2h) {
    cudaEventRecord(p.ev[5], p.stream);
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    for (float& t : p.t_ms) t = 0.f;
    const int slot[5] = {3, 2, 1, 0, 4};  // H2D, Fourier, Legendre, spectral stencil, D2H
    for (int i = 0; i < 5; ++i) {
        float ms = 0.f;
        if (i == 4 && !d2h) break;
        cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]);
        p.t_ms[slot[i]] += ms;
    }
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

}  // namespace

extern "C" {

int sptrans_invtrans_adj(sptrans_plan* plan, int nsc, const double* gp, int nvd, double* vor, double* div,
                         double* scalar_spectra) {
    if (nvd <= 0) return sptrans_invtrans_adj_scalar(plan, nsc, gp, scalar_spectra);
    int rc = adj_args_ok(plan, "sptrans_invtrans_adj");
    if (rc) return rc;
    Plan& p = plan->p;
    if (nsc < 0 || !gp || !vor || !div || (nsc > 0 && !scalar_spectra)) {
        set_error("sptrans_invtrans_adj: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    const int T = p.g.T;
    const int nall = 2 * nvd + nsc;
    if ((rc = gp_to_packed_adj(p, nall, 2 * nvd, gp, p.ev))) return rc;
    const size_t nvd_spec = spec_doubles(nvd, T), nsc_spec = spec_doubles(nsc, T);
    const bool out_host = !is_device_pointer(vor);
    double *d_vor = vor, *d_div = div, *d_sc = scalar_spectra;
    if (out_host) {
        if ((rc = ensure(p.d_spec2, p.spec2_cap, 2 * nvd_spec + nsc_spec))) return rc;
        d_vor = p.d_spec2;
        d_div = p.d_spec2 + nvd_spec;
        d_sc = p.d_spec2 + 2 * nvd_spec;
    }
    if ((rc = launch_merge_uv_scalar_adj(p.stream, T, nvd, nsc, p.d_sp_rowoff, p.d_packed, d_vor, d_div, d_sc, &p.launches)))
        return rc;
    cudaEventRecord(p.ev[4], p.stream);
    if (out_host) {
        SPT_CUDA(cudaMemcpyAsync(vor, d_vor, nvd_spec * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
        SPT_CUDA(cudaMemcpyAsync(div, d_div, nvd_spec * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
        if (nsc > 0)
            SPT_CUDA(cudaMemcpyAsync(scalar_spectra, d_sc, nsc_spec * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
    }
    return finish_timings(p, out_host);
}

int sptrans_invtrans_vordiv2wind_adj(sptrans_plan* plan, int nvd, const double* wind, double* vor, double* div) {
    if (nvd == 0) return SPTRANS_OK;
    return sptrans_invtrans_adj(plan, 0, wind, nvd, vor, div, nullptr);
}

int sptrans_invtrans_grad_adj(sptrans_plan* plan, int nf, const double* grad, double* spectra) {
    int rc = adj_args_ok(plan, "sptrans_invtrans_grad_adj");
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf < 0 || (nf > 0 && (!grad || !spectra))) {
        set_error("sptrans_invtrans_grad_adj: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (nf == 0) return SPTRANS_OK;
    const int T = p.g.T;
    if ((rc = gp_to_packed_adj(p, 2 * nf, 2 * nf, grad, p.ev))) return rc;
    const size_t nspec = spec_doubles(nf, T);
    const bool out_host = !is_device_pointer(spectra);
    double* d_sp = spectra;
    if (out_host) {
        if ((rc = ensure(p.d_spec2, p.spec2_cap, nspec))) return rc;
        d_sp = p.d_spec2;
    }
    if ((rc = launch_grad_spectra_adj(p.stream, T, nf, p.d_sp_rowoff, p.d_packed, d_sp, &p.launches))) return rc;
    cudaEventRecord(p.ev[4], p.stream);
    if (out_host) SPT_CUDA(cudaMemcpyAsync(spectra, d_sp, nspec * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
    return finish_timings(p, out_host);
}

int sptrans_dirtrans_adj_scalar(sptrans_plan* plan, int nf, const double* spectra, double* gp) {
    int rc = adj_args_ok(plan, "sptrans_dirtrans_adj_scalar");
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf < 0 || (nf > 0 && (!spectra || !gp))) {
        set_error("sptrans_dirtrans_adj_scalar: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (nf == 0) return SPTRANS_OK;
    if (!p.d_dirscale) {
        set_error("sptrans_dirtrans_adj_scalar: plan was created without quadrature weights");
        return SPTRANS_ERR_INVALID;
    }
    const int T = p.g.T;
    const size_t nspec = spec_doubles(nf, T), ngp = static_cast<size_t>(p.g.npts) * nf;
    const double* d_spec = spectra;
    double* d_gp = gp;
    cudaEventRecord(p.ev[0], p.stream);
    if (!is_device_pointer(spectra)) {
        if ((rc = ensure(p.d_spec, p.spec_cap, nspec))) return rc;
        SPT_CUDA(cudaMemcpyAsync(p.d_spec, spectra, nspec * sizeof(double), cudaMemcpyHostToDevice, p.stream));
        d_spec = p.d_spec;
    }
    const bool gp_host = !is_device_pointer(gp);
    if (gp_host) {
        if ((rc = ensure(p.d_gp, p.gp_cap, ngp))) return rc;
        d_gp = p.d_gp;
    }
    // inverse tiles up to n = T+1 keep the m = T column (the data rows n = T+1 are zero)
    if ((rc = build_tiles(p, nf, T + 1, T))) return rc;
    if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nf)))) return rc;
    if ((rc = ensure(p.d_fourier, p.fourier_cap, fourier_doubles(p, nf)))) return rc;
    cudaEventRecord(p.ev[1], p.stream);
    if ((rc = launch_pack_spectra(p, nf, T, d_spec, p.d_packed, kPackKeepMT | kPackDirAdj))) return rc;
    cudaEventRecord(p.ev[2], p.stream);
    if ((rc = launch_legendre_inv(p, nf, p.d_packed, p.d_fourier))) return rc;
    cudaEventRecord(p.ev[3], p.stream);
    if ((rc = launch_fourier_inv(p, nf, T, p.d_fourier, d_gp, nf, p.d_dirscale))) return rc;
    cudaEventRecord(p.ev[4], p.stream);
    if (gp_host) SPT_CUDA(cudaMemcpyAsync(gp, d_gp, ngp * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
    cudaEventRecord(p.ev[5], p.stream);
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    for (float& t : p.t_ms) t = 0.f;
    const int slot[5] = {3, 0, 1, 2, 4};
    for (int i = 0; i < 5; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]);
        p.t_ms[slot[i]] += ms;
    }
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

int sptrans_dirtrans_wind2vordiv_adj(sptrans_plan* plan, int nf, const double* vor, const double* div, double* wind) {
    int rc = adj_args_ok(plan, "sptrans_dirtrans_wind2vordiv_adj");
    if (rc) return rc;
    Plan& p = plan->p;
    if (nf < 0 || (nf > 0 && (!vor || !div || !wind))) {
        set_error("sptrans_dirtrans_wind2vordiv_adj: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    if (nf == 0) return SPTRANS_OK;
    if (!p.d_dirscale_uv) {
        set_error("sptrans_dirtrans_wind2vordiv_adj: plan was created without quadrature weights");
        return SPTRANS_ERR_INVALID;
    }
    // wind2vordiv = S o B(T+1) o diag(1 / (a cos)), S the spectral stencil (uv_to_vordiv_kernel).  Its adjoint w.r.t. the
    // spectral inner product that counts m > 0 twice: diag(1 / (a cos)) B(T+1)^T C S^T -- S acts within one m, so C passes
    // through it -- i.e. the transposed stencil followed by the machinery of dirtrans_adj at truncation T+1.
    const int T = p.g.T, nall = 2 * nf;
    const size_t nspec = spec_doubles(nf, T), ngp = static_cast<size_t>(p.g.npts) * nall;
    const double *d_vor = vor, *d_div = div;
    double* d_gp = wind;
    cudaEventRecord(p.ev[0], p.stream);
    if (!is_device_pointer(vor)) {
        if ((rc = ensure(p.d_spec2, p.spec2_cap, 2 * nspec))) return rc;
        SPT_CUDA(cudaMemcpyAsync(p.d_spec2, vor, nspec * sizeof(double), cudaMemcpyHostToDevice, p.stream));
        SPT_CUDA(cudaMemcpyAsync(p.d_spec2 + nspec, div, nspec * sizeof(double), cudaMemcpyHostToDevice, p.stream));
        d_vor = p.d_spec2;
        d_div = p.d_spec2 + nspec;
    }
    const bool gp_host = !is_device_pointer(wind);
    if (gp_host) {
        if ((rc = ensure(p.d_gp, p.gp_cap, ngp))) return rc;
        d_gp = p.d_gp;
    }
    if ((rc = ensure(p.d_spec, p.spec_cap, spec_doubles(nall, T + 1)))) return rc;
    if ((rc = build_tiles(p, nall, T + 1, T))) return rc;
    if ((rc = ensure(p.d_packed, p.packed_cap, packed_doubles(p, nall)))) return rc;
    if ((rc = ensure(p.d_fourier, p.fourier_cap, fourier_doubles(p, nall)))) return rc;
    cudaEventRecord(p.ev[1], p.stream);
    if ((rc = launch_uv_to_vordiv_adj(p.stream, T, nf, d_vor, d_div, p.d_spec, &p.launches))) return rc;
    if ((rc = launch_pack_spectra(p, nall, T + 1, p.d_spec, p.d_packed, kPackDirAdj))) return rc;
    cudaEventRecord(p.ev[2], p.stream);
    if ((rc = launch_legendre_inv(p, nall, p.d_packed, p.d_fourier))) return rc;
    cudaEventRecord(p.ev[3], p.stream);
    if ((rc = launch_fourier_inv(p, nall, T, p.d_fourier, d_gp, nall, p.d_dirscale_uv))) return rc;
    cudaEventRecord(p.ev[4], p.stream);
    if (gp_host) SPT_CUDA(cudaMemcpyAsync(wind, d_gp, ngp * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
    cudaEventRecord(p.ev[5], p.stream);
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    for (float& t : p.t_ms) t = 0.f;
    const int slot[5] = {3, 0, 1, 2, 4};
    for (int i = 0; i < 5; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]);
        p.t_ms[slot[i]] += ms;
    }
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

}  // extern "C"
