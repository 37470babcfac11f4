"""Multi-GPU spectral transform: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

Sharding: the Legendre stage is independent per zonal wavenumber m, the Fourier stage per latitude
pair, so rank r owns a cost-balanced set of m (and only those blocks of the Legendre table) and a band of
latitude pairs.  Between the two stages there is exactly one exchange -- the classic spectral-transform
transposition of the (my m) x (your latitudes) blocks of the Legendre<->Fourier buffer.  Two implementations:

* exchange="peer" (default): the exchange buffers of all ranks are mapped into every process (CUDA IPC over
  NVLink/NVSwitch; this module only ships the 64-byte handles through torch.distributed once).  The inverse
  Legendre GEMM stores its rows straight into the consumer's buffer from its epilogue, the direct transform
  pushes its rows with one copy kernel, and a device-side flag barrier is the only synchronisation: one
  stream-ordered library call per transform, no collective library on the data path.
* exchange="nccl": pack -> ONE all_to_all_single -> unpack (kept as the portable fallback and as the baseline
  the fused path is measured against).

The grid-point fields stay latitude-band distributed (atlas's StructuredColumns distribution) unless
`gather_grid()` is called, which is the single all-gather the north star mentions.

The reference has no counterpart: TransLocal throws for mpi::size() > 1 (trans/local/TransLocal.cc:338-340).
All arithmetic happens in the CUDA library; this module only sequences stages and the collective.
"""
import ctypes as C

import numpy as np

from . import _lib
from .grid import Grid, StructuredGrid
from .trans import Trans, _ptr


def shard_layout(grid, truncation, rank, nranks):
    """Host-only view of the sharding (no GPU needed): owner[m], band[], rows exchanged with every peer."""
    T = int(truncation)
    owner = np.zeros(T + 1, dtype=np.int32)
    band = np.zeros(nranks + 1, dtype=np.int32)
    ms = np.zeros(nranks, dtype=np.int64)
    bs = np.zeros(nranks, dtype=np.int64)
    nx, lat = grid.nx(), grid.y()
    flags = 1 if grid.regular else 0
    LL = C.POINTER(C.c_longlong)
    _lib.check(_lib.lib.sptrans_shard_layout(grid.ny(), nx.ctypes.data_as(_lib.c_int_p), lat.ctypes.data_as(_lib.c_double_p), T, flags,
                                             rank, nranks, owner.ctypes.data_as(_lib.c_int_p), band.ctypes.data_as(_lib.c_int_p),
                                             ms.ctypes.data_as(LL), bs.ctypes.data_as(LL)))
    return owner, band, ms, bs


def shard_segments(grid, truncation, rank, nranks, side):
    """Copy segments (exchange-buffer row, packed-buffer row, nrows) of one side; side 0 = m side, 1 = band side."""
    nx, lat = grid.nx(), grid.y()
    flags = 1 if grid.regular else 0
    args = (grid.ny(), nx.ctypes.data_as(_lib.c_int_p), lat.ctypes.data_as(_lib.c_double_p), int(truncation), flags, rank, nranks, side)
    n = _lib.lib.sptrans_shard_segments(*args, None)
    out = np.zeros((max(n, 0), 3), dtype=np.int64)
    if n > 0:
        _lib.lib.sptrans_shard_segments(*args, out.ctypes.data_as(C.POINTER(C.c_longlong)))
    return out


class ShardedTrans:
    """m-sharded / latitude-band-sharded transform over an initialised torch.distributed process group."""

    def __init__(self, grid, truncation, device, group=None, exchange="peer", local_io=False):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if isinstance(grid, str):
            grid = Grid(grid)
        self.grid, self.T = grid, int(truncation)
        # local_io (SPTRANS_SHARD_LOCAL_IO): the arrays handed to invtrans / dirtrans hold only this rank's share --
        # spectra of its zonal wavenumbers, grid rows of its latitude band -- instead of full-size arrays of which only
        # that share is touched (36 GB per rank at TCo2559 L137)
        self.local_io = bool(local_io)
        self.trans = Trans(grid, truncation, device=device, rank=self.rank, nranks=self.world, local_io=self.local_io)
        ms = np.zeros(self.world, dtype=np.int64)
        bs = np.zeros(self.world, dtype=np.int64)
        LL = C.POINTER(C.c_longlong)
        _lib.check(_lib.lib.sptrans_exchange_rows(self.trans._h, ms.ctypes.data_as(LL), bs.ctypes.data_as(LL)))
        self.m_rows, self.band_rows = ms, bs
        self.device = torch.device("cuda", device)
        self._nf = None
        self._owner, self._band, _, _ = shard_layout(grid, self.T, self.rank, self.world)
        self._rowoff = np.concatenate([[0], np.cumsum(grid.nx(), dtype=np.int64)])
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.exchange = exchange
        # run the library on torch's current stream: the NCCL collective is ordered against that stream, so the
        # pack -> all_to_all -> unpack chain needs no extra synchronisation
        with torch.cuda.device(self.device):
            self.trans.set_stream(torch.cuda.current_stream().cuda_stream)

    def _attach_peers(self, nf):
        """Allocate this rank's exchange region and map every peer's (collective)."""
        h = self.trans._h
        handle = C.create_string_buffer(64)
        _lib.check(_lib.lib.sptrans_peer_alloc(h, int(nf), handle))
        handles = [None] * self.world
        self.dist.all_gather_object(handles, handle.raw, group=self.group)
        _lib.check(_lib.lib.sptrans_peer_attach_ipc(h, self.world, b"".join(handles)))
        self.dist.barrier(group=self.group)   # nobody stores into a region before everybody has mapped it

    def _buffers(self, nf):
        if self._nf == nf:
            return
        if self.exchange == "peer":
            self._attach_peers(nf)
            self._nf = nf
            return
        t = self.torch
        self.fb = t.zeros(self.trans.fourier_elems_per_field() * 2 * nf, dtype=t.float64, device=self.device)
        self.buf_m = t.empty(max(int(self.m_rows.sum()), 1) * 2 * nf, dtype=t.float64, device=self.device)
        self.buf_b = t.empty(max(int(self.band_rows.sum()), 1) * 2 * nf, dtype=t.float64, device=self.device)
        self._nf = nf

    # ---- host-side views of this rank's share (NumPy; tests, benchmark set-up) ----
    def my_m(self):
        return [m for m in range(self.T + 1) if self._owner[m] == self.rank]

    def local_spectra(self, sp_global, nf):
        """[my zonal wavenumbers ascending][n][re/im][field] slab of a global [m][n][re/im][field] array."""
        T = self.T
        parts = [sp_global[(2 * T + 3 - m) * m // 2 * 2 * nf:((2 * T + 3 - m) * m // 2 + T - m + 1) * 2 * nf] for m in self.my_m()]
        return np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0)

    def scatter_local_spectra(self, sp_local, sp_global, nf):
        """inverse of local_spectra: writes this rank's zonal wavenumbers into a global array"""
        T, o = self.T, 0
        for m in self.my_m():
            n = (T - m + 1) * 2 * nf
            g0 = (2 * T + 3 - m) * m // 2 * 2 * nf
            sp_global[g0:g0 + n] = sp_local[o:o + n]
            o += n

    def local_grid_rows(self, gp_global, nf):
        """[field][rows of my band: northern rows, then their southern mirrors] from a global [field][point] array,
        padded to the plan's per-field stride (sptrans_local_sizes)."""
        _, stride = self.trans.local_sizes()
        g2 = np.asarray(gp_global).reshape(nf, self.grid.size())
        out = np.zeros((nf, stride))
        o = 0
        for a, b in self._band_spans(self.rank):
            out[:, o:o + b - a] = g2[:, a:b]
            o += b - a
        return out.reshape(-1)

    def scatter_local_grid_rows(self, gp_local, gp_global, nf):
        _, stride = self.trans.local_sizes()
        g2 = np.asarray(gp_global).reshape(nf, self.grid.size())
        l2 = np.asarray(gp_local).reshape(nf, stride)
        o = 0
        for a, b in self._band_spans(self.rank):
            g2[:, a:b] = l2[:, o:o + b - a]
            o += b - a

    def invtrans(self, nf, d_spec, d_gp):
        """d_spec: full-size spectral array (only this rank's zonal wavenumbers are read);
        d_gp: full-size grid array (only this rank's latitude rows are written).  With local_io both hold only this
        rank's share (see local_spectra / local_grid_rows)."""
        self._buffers(nf)
        tr, h = self.trans, self.trans._h
        if self.exchange == "peer":
            _lib.check(_lib.lib.sptrans_invtrans_sharded(h, int(nf), _ptr(d_spec), _ptr(d_gp)))
            return
        tr.invtrans_legendre(nf, self.T, d_spec, self.fb)
        _lib.check(_lib.lib.sptrans_exchange_pack(h, nf, 0, _ptr(self.fb), _ptr(self.buf_m)))
        k = 2 * nf
        self.dist.all_to_all_single(self.buf_b, self.buf_m, [int(r) * k for r in self.band_rows], [int(r) * k for r in self.m_rows],
                                    group=self.group)
        _lib.check(_lib.lib.sptrans_exchange_unpack(h, nf, 1, _ptr(self.buf_b), _ptr(self.fb)))
        tr.invtrans_fourier(nf, self.T - 1, self.fb, d_gp)

    def dirtrans(self, nf, d_gp, d_spec):
        """d_gp: grid array whose rows of this rank's band are valid; d_spec: spectral array, this rank's m written."""
        self._buffers(nf)
        tr, h = self.trans, self.trans._h
        if self.exchange == "peer":
            _lib.check(_lib.lib.sptrans_dirtrans_sharded(h, int(nf), _ptr(d_gp), _ptr(d_spec)))
            return
        tr.dirtrans_fourier(nf, d_gp, self.fb)
        _lib.check(_lib.lib.sptrans_exchange_pack(h, nf, 1, _ptr(self.fb), _ptr(self.buf_b)))
        k = 2 * nf
        self.dist.all_to_all_single(self.buf_m, self.buf_b, [int(r) * k for r in self.m_rows], [int(r) * k for r in self.band_rows],
                                    group=self.group)
        _lib.check(_lib.lib.sptrans_exchange_unpack(h, nf, 0, _ptr(self.buf_m), _ptr(self.fb)))
        tr.dirtrans_legendre(nf, self.fb, d_spec)

    def band_rows_slice(self):
        """Grid rows owned by this rank: (north rows [j0,j1), south rows [nlat-j1, nlat-j0))."""
        return int(self._band[self.rank]), int(self._band[self.rank + 1])

    def _band_spans(self, r):
        """Point ranges [a, b) of the grid rows owned by rank r: its northern rows and their southern mirrors."""
        nlat, ro = self.grid.ny(), self._rowoff
        j0, j1 = int(self._band[r]), int(self._band[r + 1])
        spans = [(int(ro[j0]), int(ro[j1])), (int(ro[nlat - j1]), int(ro[nlat - j0]))]
        if spans[0][1] > spans[1][0]:   # the band holds the equator row of a grid with an odd number of rows
            spans = [(spans[0][0], spans[1][1])]
        return [(a, b) for a, b in spans if b > a]

    def gather_grid(self, nf, d_gp, d_global=None):
        """Replicate the latitude-band-distributed grid fields on every rank: ONE all-gather (the "single NCCL allgather
        that reassembles the grid-point Field" of the north star).  Every rank contributes exactly the rows it
        owns, padded to the longest contribution, and every other row of the global array is overwritten with its owner's
        values -- no precondition on what the rows of the other bands held before the call.
        Full-size arrays: gathered in place in `d_gp`.  local_io: `d_gp` is this rank's band, `d_global` [field][point]
        receives all bands."""
        t = self.torch
        npts = self.grid.size()
        spans = [self._band_spans(r) for r in range(self.world)]
        lens = [sum(b - a for a, b in sp) for sp in spans]
        lmax = max(max(lens), 1)
        if self.local_io:
            _, stride = self.trans.local_sizes()
            if d_global is None:
                raise ValueError("gather_grid: local_io plans need the global output array")
            gp2d = d_global.view(nf, npts)
            mine = d_gp.view(nf, stride)[:, :lens[self.rank]]
        else:
            gp2d = d_gp.view(nf, npts)
            mine = None
        send = t.empty(nf, lmax, dtype=t.float64, device=self.device)
        if mine is not None:
            send[:, :lens[self.rank]].copy_(mine)
        else:
            o = 0
            for a, b in spans[self.rank]:
                send[:, o:o + b - a].copy_(gp2d[:, a:b])
                o += b - a
        recv = t.empty(self.world, nf, lmax, dtype=t.float64, device=self.device)
        self.dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
        for r in range(self.world):
            if r == self.rank and not self.local_io:
                continue
            o = 0
            for a, b in spans[r]:
                gp2d[:, a:b].copy_(recv[r, :, o:o + b - a])
                o += b - a
