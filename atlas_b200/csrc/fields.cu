// Repacking between atlas Field storage and the IFS-style row buffers of the transform.
//
// A multi-level grid-point Field is (node, level[, component]) with the LAST index fastest (atlas row-major arrays;
// owned nodes are 0..npts-1 in grid order, functionspace/detail/StructuredColumns.h:104-137), the transform works on
// rows [field][node] with field = component * nlev + level -- exactly the packing loops of TransIFS
// (ecmwf/atlas src/atlas/trans/ifs/TransIFS.cc:610-667 gp, :1392-1437 wind, :2113-2137 gradient), which run on
// the host there.  Here it is one tiled transpose on the device (32 x 32 doubles through shared memory, both sides
// coalesced): 2 x 8 bytes of HBM traffic per value.
#include "plan.hpp"

namespace sptrans {

namespace {

constexpr int kTile = 32;

// to_rows:  rows[(c * nlev + l) * npts + p] = field[(p * nlev + l) * ncomp + c]
// !to_rows: the inverse assignment
template <bool TO_ROWS>
__global__ void __launch_bounds__(kTile * 8)
gp_repack_kernel(const double* __restrict__ in, double* __restrict__ out, long long npts, int nlev, int ncomp) {
    __shared__ double tile[kTile][kTile + 1];
    const int nf = nlev * ncomp;
    const long long p0 = static_cast<long long>(blockIdx.x) * kTile;
    const int q0 = blockIdx.y * kTile;  // index into the field-fastest order q = l * ncomp + c
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (TO_ROWS) {
        for (int r = ty; r < kTile; r += 8) {  // r: point within the tile, tx: q within the tile
            const long long p = p0 + r;
            const int q = q0 + tx;
            if (p < npts && q < nf) tile[r][tx] = in[p * nf + q];
        }
        __syncthreads();
        for (int r = ty; r < kTile; r += 8) {  // r: q within the tile, tx: point
            const long long p = p0 + tx;
            const int q = q0 + r;
            if (p < npts && q < nf) {
                const int l = q / ncomp, c = q - l * ncomp;
                out[static_cast<long long>(c * nlev + l) * npts + p] = tile[tx][r];
            }
        }
    }
    else {
        for (int r = ty; r < kTile; r += 8) {
            const long long p = p0 + tx;
            const int q = q0 + r;
            if (p < npts && q < nf) {
                const int l = q / ncomp, c = q - l * ncomp;
                tile[tx][r] = in[static_cast<long long>(c * nlev + l) * npts + p];
            }
        }
        __syncthreads();
        for (int r = ty; r < kTile; r += 8) {
            const long long p = p0 + r;
            const int q = q0 + tx;
            if (p < npts && q < nf) out[p * nf + q] = tile[r][tx];
        }
    }
}

}  // namespace

int launch_gp_repack(Plan& p, int nlev, int ncomp, const double* d_in, double* d_out, bool to_rows) {
    const long long npts = p.g.npts;
    const int nf = nlev * ncomp;
    if (npts == 0 || nf == 0) return SPTRANS_OK;
    dim3 grid(static_cast<unsigned>((npts + kTile - 1) / kTile), static_cast<unsigned>((nf + kTile - 1) / kTile));
    dim3 block(kTile, 8);
    if (to_rows) gp_repack_kernel<true><<<grid, block, 0, p.stream>>>(d_in, d_out, npts, nlev, ncomp);
    else gp_repack_kernel<false><<<grid, block, 0, p.stream>>>(d_in, d_out, npts, nlev, ncomp);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

}  // namespace sptrans
