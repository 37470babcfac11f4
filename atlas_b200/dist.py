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
import json
import os
import time

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

    def __init__(self, grid, truncation, device, group=None, exchange="peer"):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if isinstance(grid, str):
            grid = Grid(grid)
        self.grid, self.T = grid, int(truncation)
        self.trans = Trans(grid, truncation, device=device, rank=self.rank, nranks=self.world)
        ms = np.zeros(self.world, dtype=np.int64)
        bs = np.zeros(self.world, dtype=np.int64)
        LL = C.POINTER(C.c_longlong)
        _lib.check(_lib.lib.sptrans_exchange_rows(self.trans._h, ms.ctypes.data_as(LL), bs.ctypes.data_as(LL)))
        self.m_rows, self.band_rows = ms, bs
        self.device = torch.device("cuda", device)
        self._nf = None
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

    def invtrans(self, nf, d_spec, d_gp):
        """d_spec: full-size spectral array (only this rank's zonal wavenumbers are read);
        d_gp: full-size grid array (only this rank's latitude rows are written)."""
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
        owner, band, _, _ = shard_layout(self.grid, self.T, self.rank, self.world)
        return int(band[self.rank]), int(band[self.rank + 1])

    def gather_grid(self, nf, d_gp):
        """Replicate the latitude-band-distributed grid fields on every rank (one all-reduce of disjoint rows)."""
        self.dist.all_reduce(d_gp, group=self.group)


def bench_sharded(args, rank, world, local_rank, metric, unit, fp64_peak):
    """bench.py for N > 1 (strong scaling: the TCo1279 L137 job is fixed, split over N GPUs)."""
    import torch
    import torch.distributed as dist

    import bench as B
    import helpers as H

    gridname, T, nf = B.workload(args.workload)
    grid = Grid(gridname)
    st = ShardedTrans(grid, T, local_rank, exchange=args.exchange)
    npts = grid.size()
    d_sp = torch.from_numpy(H.synthetic_spectra(T, nf)).to(st.device)
    d_gp = torch.zeros(nf * npts, dtype=torch.float64, device=st.device)
    d_sp2 = torch.zeros_like(d_sp)
    for _ in range(args.warmup):
        st.invtrans(nf, d_sp, d_gp)
        st.dirtrans(nf, d_gp, d_sp2)
    torch.cuda.synchronize()
    dist.barrier()
    launches0 = st.trans.kernel_launches()
    sampler = B.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        st.invtrans(nf, d_sp, d_gp)
        st.dirtrans(nf, d_gp, d_sp2)
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=st.device, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = torch.tensor([st.trans.kernel_launches() - launches0], device=st.device, dtype=torch.float64)
    dist.all_reduce(launches)
    # stage times of the last inverse + direct pair on every rank (outside the timed region)
    stage = torch.zeros(6, device=st.device, dtype=torch.float64)
    if args.exchange == "peer":
        st.invtrans(nf, d_sp, d_gp)
        ti = st.trans.last_timings()
        st.dirtrans(nf, d_gp, d_sp2)
        td = st.trans.last_timings()
        stage = torch.tensor([ti["legendre"], ti["exchange_wait"], ti["fourier"], td["fourier"], td["exchange_wait"], td["legendre"]],
                             device=st.device, dtype=torch.float64)
    stage_all = [torch.zeros_like(stage) for _ in range(world)]
    dist.all_gather(stage_all, stage)
    # correctness of the sharded round trip: every rank owns some m; gather spectra and compare on rank 0
    owner, band, _, _ = shard_layout(grid, T, rank, world)
    if rank == 0:
        clocks = sampler.stop()
        ms_per_step = float(ms.item()) / args.steps
        out = {
            "metric": metric, "value": 1e3 / ms_per_step, "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload} L{nf} invtrans+dirtrans fp64 (grid {gridname}, T{T})",
                       "parallelism": f"zonal-wavenumber sharded Legendre x latitude-band sharded Fourier over {world} GPUs; "
                                      + ("exchange fused into the kernels over NVLink peer memory (Legendre epilogue stores / "
                                         "push kernel, device-side barrier)" if args.exchange == "peer"
                                         else "one NCCL all-to-all per direction") + "; grid fields stay band-distributed",
                       "l2": "inputs larger than L2"},
            "clocks": clocks, "gpu_launches": int(launches.item()),
            "e2e": None, "roofline": None, "cpu_baseline": None,
            "stage_ms_per_rank": {k: [round(float(x[i]), 3) for x in stage_all] for i, k in enumerate(
                ["inv_legendre", "inv_exchange_wait", "inv_fourier", "dir_fourier_push", "dir_exchange_wait", "dir_legendre"])},
        }
        print(json.dumps(out))
    dist.destroy_process_group()
