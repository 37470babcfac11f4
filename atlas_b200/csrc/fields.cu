// Repacking between atlas Field storage and the IFS-style row buffers of the transform.
//
// A multi-level grid-point Field is (node, level[, component]) with the LAST index fastest (atlas row-major arrays;
// owned nodes are 0..npts-1 in grid order, functionspace/detail/StructuredColumns.h:104-137), the transform works on
// rows [field][node] with field = component * nlev + level -- exactly the packing loops of TransIFS
// (ecmwf/atlas src/atlas/trans/ifs/TransIFS.cc:610-667 gp, :1392-1437 wind, :2113-2137 gradient), which run on
// the host there.  Here it is one tiled transpose on the device (32 x 32 doubles through shared memory, both sides
// coalesced): 2 x 8 bytes of HBM traffic per value.
#include <algorithm>
#include <vector>

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

// ---- cropped (regional) structured grids ------------------------------------------------------------------------------
// The reference transforms the rows of the GLOBAL grid the regional grid is cut from and copies out, per row, the
// longitudes of the crop starting at jlonMin with wrap-around (TransLocal.cc:501-531, :1180-1187).  Here the Fourier
// kernels write the global rows of the latitude band into a work array and this kernel does the copy-out.
struct CropRow {
    long long src;   // offset of the global row within one field of the band work array
    long long dst;   // offset of the crop row within one field of the user's array
    int nxg;         // points of the global row
    int start;       // first longitude index of the crop (jlonMin)
    int n;           // points of the crop row
};

namespace {
__global__ void crop_gather_kernel(const CropRow* __restrict__ rows, int nf, long long band_stride, long long crop_npts,
                                   const double* __restrict__ band, double* __restrict__ gp) {
    const CropRow r = rows[blockIdx.x];
    for (int f = blockIdx.y; f < nf; f += gridDim.y) {
        const double* src = band + f * band_stride + r.src;
        double* dst = gp + f * crop_npts + r.dst;
        for (int i = threadIdx.x; i < r.n; i += blockDim.x) {
            int k = r.start + i;
            if (k >= r.nxg) k -= r.nxg;   // (start < nxg and n <= nxg: one wrap at most)
            dst[i] = src[k];
        }
    }
}
}  // namespace

int upload_crop_rows(Plan& p) {
    const HostGeom& g = p.g;
    std::vector<CropRow> rows(g.crop_nx.size());
    long long dst = 0;
    for (size_t r = 0; r < rows.size(); ++r) {
        const int jg = g.crop_jlat_min + static_cast<int>(r);
        rows[r].src = g.gp_rowoff[jg];
        rows[r].dst = dst;
        rows[r].nxg = g.nx[jg];
        rows[r].start = g.crop_jlon_min[r];
        rows[r].n = g.crop_nx[r];
        dst += g.crop_nx[r];
    }
    SPT_CUDA(cudaMalloc(&p.d_crop_rows, std::max<size_t>(rows.size(), 1) * sizeof(CropRow)));
    SPT_CUDA(cudaMemcpyAsync(p.d_crop_rows, rows.data(), rows.size() * sizeof(CropRow), cudaMemcpyHostToDevice, p.stream));
    SPT_CUDA(cudaStreamSynchronize(p.stream));
    return SPTRANS_OK;
}

int launch_crop_gather(Plan& p, int nf, const double* d_band, double* d_gp) {
    const int nrows = static_cast<int>(p.g.crop_nx.size());
    if (nrows == 0 || nf == 0) return SPTRANS_OK;
    dim3 grid(nrows, std::min(nf, 64));
    crop_gather_kernel<<<grid, 128, 0, p.stream>>>(static_cast<const CropRow*>(p.d_crop_rows), nf, p.g.gp_stride, p.g.crop_npts,
                                                   d_band, d_gp);
    p.launches++;
    SPT_CUDA(cudaGetLastError());
    return SPTRANS_OK;
}

}  // namespace sptrans
