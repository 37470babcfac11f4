// C ABI, atlas Field layouts: the entry points behind TransImpl's Field overloads (trans/detail/TransImpl.h:54-100,
// C bindings atlas__Trans__{invtrans,dirtrans,...}_field in trans/detail/TransInterface.h:74-100) for multi-level
// Fields.  A spectral Field (nspec2, nlev) is already the raw [coeff][field] layout; a grid-point Field is
// (node, level[, component]) with the last index fastest, i.e. the transpose of the IFS-style [field][node] rows the
// transform kernels work on.  TransIFS repacks on the host (trans/ifs/TransIFS.cc:610-667, :1392-1437, :2113-2137); here
// the repack is one tiled device transpose (fields.cu) between the user's buffer and the row image `Plan::d_rows`,
// and the transform itself is the same call chain as the raw-pointer entry points (api.cu), handed device pointers.
#include <cstring>

#include "plan.hpp"

using namespace sptrans;

namespace {

int field_args_ok(sptrans_plan* plan, int nlev, const void* a, const void* b, const char* who) {
    if (!plan) {
        set_error(std::string(who) + ": null plan");
        return SPTRANS_ERR_INVALID;
    }
    if (nlev < 0 || (nlev > 0 && (!a || !b))) {
        set_error(std::string(who) + ": invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    SPT_CUDA(cudaSetDevice(plan->p.device));
    return SPTRANS_OK;
}

// rows [ncomp * nlev][npts] in p.d_rows  ->  Field (npts, nlev, ncomp) at `user` (device: in place; host: via p.d_gp)
int rows_to_field(Plan& p, int nlev, int ncomp, double* user) {
    const size_t n = static_cast<size_t>(p.g.npts) * nlev * ncomp;
    int rc;
    float ms = 0.f;
    cudaEventRecord(p.ev[6], p.stream);
    if (is_device_pointer(user)) {
        if ((rc = launch_gp_repack(p, nlev, ncomp, p.d_rows, user, false))) return rc;
        cudaEventRecord(p.ev[7], p.stream);
    }
    else {
        if ((rc = ensure(p.d_gp, p.gp_cap, n))) return rc;
        if ((rc = launch_gp_repack(p, nlev, ncomp, p.d_rows, p.d_gp, false))) return rc;
        cudaEventRecord(p.ev[7], p.stream);
        SPT_CUDA(cudaMemcpyAsync(user, p.d_gp, n * sizeof(double), cudaMemcpyDeviceToHost, p.stream));
    }
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    cudaEventElapsedTime(&ms, p.ev[6], p.ev[7]);
    p.t_ms[6] = ms;
    return SPTRANS_OK;
}

// Field (npts, nlev, ncomp) at `user`  ->  rows in p.d_rows (stream-ordered; the transform that follows runs on p.stream)
int field_to_rows(Plan& p, int nlev, int ncomp, const double* user) {
    const size_t n = static_cast<size_t>(p.g.npts) * nlev * ncomp;
    int rc;
    if ((rc = ensure(p.d_rows, p.rows_cap, n))) return rc;
    const double* src = user;
    if (!is_device_pointer(user)) {
        if ((rc = ensure(p.d_gp, p.gp_cap, n))) return rc;
        SPT_CUDA(cudaMemcpyAsync(p.d_gp, user, n * sizeof(double), cudaMemcpyHostToDevice, p.stream));
        src = p.d_gp;
    }
    cudaEventRecord(p.ev[6], p.stream);
    if ((rc = launch_gp_repack(p, nlev, ncomp, src, p.d_rows, true))) return rc;
    cudaEventRecord(p.ev[7], p.stream);
    return SPTRANS_OK;
}

void note_repack_time(Plan& p) {  // after the transform's own synchronisation
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.ev[6], p.ev[7]) == cudaSuccess) p.t_ms[6] = ms;
    else cudaGetLastError();
}

}  // namespace

extern "C" {

int sptrans_invtrans_field(sptrans_plan* plan, int nlev, const double* spfield, double* gpfield) {
    int rc = field_args_ok(plan, nlev, spfield, gpfield, "sptrans_invtrans_field");
    if (rc || nlev == 0) return rc;
    Plan& p = plan->p;
    if ((rc = ensure(p.d_rows, p.rows_cap, static_cast<size_t>(p.g.npts) * nlev))) return rc;
    if ((rc = sptrans_invtrans_scalar(plan, nlev, spfield, p.d_rows))) return rc;
    return rows_to_field(p, nlev, 1, gpfield);
}

int sptrans_invtrans_vordiv2wind_field(sptrans_plan* plan, int nlev, const double* spvor, const double* spdiv,
                                       double* gpwind) {
    int rc = field_args_ok(plan, nlev, spvor, gpwind, "sptrans_invtrans_vordiv2wind_field");
    if (rc || nlev == 0) return rc;
    if (!spdiv) {
        set_error("sptrans_invtrans_vordiv2wind_field: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    Plan& p = plan->p;
    if ((rc = ensure(p.d_rows, p.rows_cap, static_cast<size_t>(p.g.npts) * nlev * 2))) return rc;
    if ((rc = sptrans_invtrans_vordiv2wind(plan, nlev, spvor, spdiv, p.d_rows))) return rc;
    return rows_to_field(p, nlev, 2, gpwind);
}

int sptrans_invtrans_grad_field(sptrans_plan* plan, int nlev, const double* spfield, double* gradfield) {
    int rc = field_args_ok(plan, nlev, spfield, gradfield, "sptrans_invtrans_grad_field");
    if (rc || nlev == 0) return rc;
    Plan& p = plan->p;
    if ((rc = ensure(p.d_rows, p.rows_cap, static_cast<size_t>(p.g.npts) * nlev * 2))) return rc;
    if ((rc = sptrans_invtrans_grad(plan, nlev, spfield, p.d_rows))) return rc;
    return rows_to_field(p, nlev, 2, gradfield);
}

int sptrans_dirtrans_field(sptrans_plan* plan, int nlev, const double* gpfield, double* spfield) {
    int rc = field_args_ok(plan, nlev, gpfield, spfield, "sptrans_dirtrans_field");
    if (rc || nlev == 0) return rc;
    Plan& p = plan->p;
    if ((rc = field_to_rows(p, nlev, 1, gpfield))) return rc;
    if ((rc = sptrans_dirtrans_scalar(plan, nlev, p.d_rows, spfield))) return rc;
    note_repack_time(p);
    return SPTRANS_OK;
}

int sptrans_invtrans_adj_field(sptrans_plan* plan, int nlev, const double* gpfield, double* spfield) {
    int rc = field_args_ok(plan, nlev, gpfield, spfield, "sptrans_invtrans_adj_field");
    if (rc || nlev == 0) return rc;
    Plan& p = plan->p;
    if ((rc = field_to_rows(p, nlev, 1, gpfield))) return rc;
    if ((rc = sptrans_invtrans_adj_scalar(plan, nlev, p.d_rows, spfield))) return rc;
    note_repack_time(p);
    return SPTRANS_OK;
}

int sptrans_dirtrans_wind2vordiv_field(sptrans_plan* plan, int nlev, const double* gpwind, double* spvor,
                                       double* spdiv) {
    int rc = field_args_ok(plan, nlev, gpwind, spvor, "sptrans_dirtrans_wind2vordiv_field");
    if (rc || nlev == 0) return rc;
    if (!spdiv) {
        set_error("sptrans_dirtrans_wind2vordiv_field: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    Plan& p = plan->p;
    if ((rc = field_to_rows(p, nlev, 2, gpwind))) return rc;
    if ((rc = sptrans_dirtrans_wind2vordiv(plan, nlev, p.d_rows, spvor, spdiv))) return rc;
    note_repack_time(p);
    return SPTRANS_OK;
}

int sptrans_dirtrans_adj_field(sptrans_plan* plan, int nlev, const double* spfield, double* gpfield) {
    int rc = field_args_ok(plan, nlev, spfield, gpfield, "sptrans_dirtrans_adj_field");
    if (rc || nlev == 0) return rc;
    Plan& p = plan->p;
    if ((rc = ensure(p.d_rows, p.rows_cap, static_cast<size_t>(p.g.npts) * nlev))) return rc;
    if ((rc = sptrans_dirtrans_adj_scalar(plan, nlev, spfield, p.d_rows))) return rc;
    return rows_to_field(p, nlev, 1, gpfield);
}

int sptrans_invtrans_vordiv2wind_adj_field(sptrans_plan* plan, int nlev, const double* gpwind, double* spvor,
                                           double* spdiv) {
    int rc = field_args_ok(plan, nlev, gpwind, spvor, "sptrans_invtrans_vordiv2wind_adj_field");
    if (rc || nlev == 0) return rc;
    if (!spdiv) {
        set_error("sptrans_invtrans_vordiv2wind_adj_field: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    Plan& p = plan->p;
    if ((rc = field_to_rows(p, nlev, 2, gpwind))) return rc;
    if ((rc = sptrans_invtrans_vordiv2wind_adj(plan, nlev, p.d_rows, spvor, spdiv))) return rc;
    note_repack_time(p);
    return SPTRANS_OK;
}

int sptrans_invtrans_grad_adj_field(sptrans_plan* plan, int nlev, const double* gradfield, double* spfield) {
    int rc = field_args_ok(plan, nlev, gradfield, spfield, "sptrans_invtrans_grad_adj_field");
    if (rc || nlev == 0) return rc;
    Plan& p = plan->p;
    if ((rc = field_to_rows(p, nlev, 2, gradfield))) return rc;
    if ((rc = sptrans_invtrans_grad_adj(plan, nlev, p.d_rows, spfield))) return rc;
    note_repack_time(p);
    return SPTRANS_OK;
}

int sptrans_dirtrans_wind2vordiv_adj_field(sptrans_plan* plan, int nlev, const double* spvor, const double* spdiv,
                                           double* gpwind) {
    int rc = field_args_ok(plan, nlev, spvor, gpwind, "sptrans_dirtrans_wind2vordiv_adj_field");
    if (rc || nlev == 0) return rc;
    if (!spdiv) {
        set_error("sptrans_dirtrans_wind2vordiv_adj_field: invalid arguments");
        return SPTRANS_ERR_INVALID;
    }
    Plan& p = plan->p;
    if ((rc = ensure(p.d_rows, p.rows_cap, static_cast<size_t>(p.g.npts) * nlev * 2))) return rc;
    if ((rc = sptrans_dirtrans_wind2vordiv_adj(plan, nlev, spvor, spdiv, p.d_rows))) return rc;
    return rows_to_field(p, nlev, 2, gpwind);
}

}  // extern "C"
