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

// Slab variant (the default): a block moves P consecutive points x ALL fields.  On the Field side that is ONE contiguous run of
// P * nf doubles (points are adjacent, fields fastest), on the row side nf runs of P doubles; the 32 x 32 tiles above touch the
// Field side in 256-byte pieces 8 nf bytes apart and reached 2.9 TB/s (5.1 / 4.0 ms per call at TCo1279 L137, measured through
// the Field entry points; this kernel: 3.6 / 3.1 ms, profiles/field_layout_timing_r02.txt).  Shared memory holds the slab as [point][field] with an odd pitch, so that the row-side accesses (lanes
// along the points) are conflict free; the Field side walks it linearly (one reciprocal multiply per element for the pitch).
template <bool TO_ROWS>
__global__ void __launch_bounds__(256)
gp_repack_slab_kernel(const double* __restrict__ in, double* __restrict__ out, long long npts, int nlev, int ncomp, int P,
                      unsigned nf_magic) {
    extern __shared__ double slab[];
    const int nf = nlev * ncomp, pitch = nf | 1;
    const long long p0 = static_cast<long long>(blockIdx.x) * P;
    const int np = static_cast<int>(min(static_cast<long long>(P), npts - p0));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int total = np * nf;
    constexpr int kU = 8;               // global accesses in flight per thread (the kernel is pure data movement)
    const int nseg = (np + 31) >> 5;    // 32-point segments of a row
    const int nitem = nf * nseg;        // (row, segment) pairs, one warp access each
    // row side of item `it`: global offset and slab index of this lane's element (-1: beyond the slab)
    auto row_item = [&](int it, long long& g, int& sidx) {
        const int r = it / nseg, seg = it - r * nseg;
        const int c = r / nlev, l = r - c * nlev;
        const int pp = seg * 32 + lane;
        g = static_cast<long long>(r) * npts + p0 + pp;
        sidx = pp < np ? pp * pitch + l * ncomp + c : -1;
    };
    auto slab_index = [&](int i) {
        const int pp = static_cast<int>(__umulhi(static_cast<unsigned>(i), nf_magic));
        return pp * pitch + (i - pp * nf);
    };
    if (TO_ROWS) {
        // (measured: batching this direction like the other one made it slower -- 6.3 ms against 3.1 ms per call at TCo1279
        // L137 -- because a row's 512-byte run is then written by two warps at different times; plain loops it is)
        const double* src = in + p0 * nf;
        for (int i = tid; i < total; i += 256) slab[slab_index(i)] = src[i];
        __syncthreads();
        for (int r = warp; r < nf; r += 8) {
            const int c = r / nlev, l = r - c * nlev;
            const int q = l * ncomp + c;
            double* dst = out + static_cast<long long>(r) * npts + p0;
            for (int pp = lane; pp < np; pp += 32) dst[pp] = slab[pp * pitch + q];
        }
    }
    else {
        for (int it0 = warp; it0 < nitem; it0 += 8 * kU) {
            double v[kU];
            int sidx[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int it = it0 + 8 * u;
                sidx[u] = -1;
                if (it < nitem) {
                    long long g;
                    row_item(it, g, sidx[u]);
                    if (sidx[u] >= 0) v[u] = in[g];
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (sidx[u] >= 0) slab[sidx[u]] = v[u];
        }
        __syncthreads();
        double* dst = out + p0 * nf;
        for (int i0 = tid; i0 < total; i0 += 256 * kU) {
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (i0 + 256 * u < total) dst[i0 + 256 * u] = slab[slab_index(i0 + 256 * u)];
        }
    }
}

}  // namespace

int launch_gp_repack(Plan& p, int nlev, int ncomp, const double* d_in, double* d_out, bool to_rows) {
    const long long npts = p.g.npts;
    const int nf = nlev * ncomp;
    if (npts == 0 || nf == 0) return SPTRANS_OK;
    {
        // points per slab: ~72 KB of shared memory (three blocks per SM), multiples of 32 where the field count allows
        const int pitch = nf | 1;
        int P = 9216 / pitch;
        P = P >= 32 ? std::min(256, P / 32 * 32) : P / 8 * 8;
        if (nf >= 2 && P >= 8 && static_cast<long long>(P) * nf < 65536) {
            const size_t smem = static_cast<size_t>(P) * pitch * sizeof(double);
            static bool attr_set[64] = {};   // (per device, like the Legendre kernels)
            if (p.device >= 64 || !attr_set[p.device]) {
                SPT_CUDA(cudaFuncSetAttribute(gp_repack_slab_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
                SPT_CUDA(cudaFuncSetAttribute(gp_repack_slab_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
                if (p.device < 64) attr_set[p.device] = true;
            }
            const unsigned magic = static_cast<unsigned>((0x100000000ull + nf - 1) / nf);   // exact i / nf for i < 65536
            const unsigned nblk = static_cast<unsigned>((npts + P - 1) / P);
            if (to_rows) gp_repack_slab_kernel<true><<<nblk, 256, smem, p.stream>>>(d_in, d_out, npts, nlev, ncomp, P, magic);
            else gp_repack_slab_kernel<false><<<nblk, 256, smem, p.stream>>>(d_in, d_out, npts, nlev, ncomp, P, magic);
            p.launches++;
            SPT_CUDA(cudaGetLastError());
            return SPTRANS_OK;
        }
    }
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
