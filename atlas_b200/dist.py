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
    # (polled less often than in the single-GPU bench: NVML queries from rank 0's process contend with its kernel launches,
    # and every rank waits for rank 0 at the exchange barriers)
    sampler = B.ClockSampler(local_rank, period_s=0.05)
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
    # steady-state duration of the two calls on every rank (back to back, no host synchronisation), again untimed
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    dist.barrier()
    for e in evs:
        e[0].record()
        st.invtrans(nf, d_sp, d_gp)
        e[1].record()
        st.dirtrans(nf, d_gp, d_sp2)
        e[2].record()
    torch.cuda.synchronize()
    calls = torch.tensor([float(np.median([e[0].elapsed_time(e[1]) for e in evs])), float(np.median([e[1].elapsed_time(e[2]) for e in evs]))],
                         device=st.device, dtype=torch.float64)
    calls_all = [torch.zeros_like(calls) for _ in range(world)]
    dist.all_gather(calls_all, calls)
    owner, band, _, _ = shard_layout(grid, T, rank, world)
    # ---- end to end with pinned HOST buffers: every rank moves its own share (the spectra of its zonal wavenumbers,
    # the grid rows of its latitude band) over its own PCIe link, inside the timed region
    e2e = None
    if not args.no_e2e:
        rowoff = np.concatenate([[0], np.cumsum(grid.nx(), dtype=np.int64)])
        nlat = grid.ny()
        j0, j1 = int(band[rank]), int(band[rank + 1])
        spans = [(int(rowoff[j0]), int(rowoff[j1])), (int(rowoff[nlat - j1]), int(rowoff[nlat - j0]))]
        if spans[0][1] > spans[1][0]:   # band contains the equator row of an odd grid: one span
            spans = [(spans[0][0], spans[1][1])]
        my_m = [m for m in range(T + 1) if owner[m] == rank]
        chunks = [((2 * T + 3 - m) * m // 2 * nf * 2, (T - m + 1) * 2 * nf) for m in my_m]
        h_sp = torch.from_numpy(H.synthetic_spectra(T, nf)).pin_memory()
        h_sp2 = torch.zeros_like(h_sp).pin_memory()
        h_gp = [torch.zeros(nf, b - a, dtype=torch.float64).pin_memory() for a, b in spans]
        gp2d = d_gp.view(nf, npts)

        def e2e_step():
            for off, n in chunks:
                d_sp[off:off + n].copy_(h_sp[off:off + n], non_blocking=True)
            st.invtrans(nf, d_sp, d_gp)
            for (a, b), h in zip(spans, h_gp):
                h.copy_(gp2d[:, a:b], non_blocking=True)
            for (a, b), h in zip(spans, h_gp):
                gp2d[:, a:b].copy_(h, non_blocking=True)
            st.dirtrans(nf, d_gp, d_sp2)
            for off, n in chunks:
                h_sp2[off:off + n].copy_(d_sp2[off:off + n], non_blocking=True)

        for _ in range(2):
            e2e_step()
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            e2e_step()
        g1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ems = torch.tensor([g0.elapsed_time(g1)], device=st.device, dtype=torch.float64)
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        nbytes = torch.tensor([8.0 * (sum(n for _, n in chunks) + nf * sum(b - a for a, b in spans))], device=st.device, dtype=torch.float64)
        dist.all_reduce(nbytes)
        e2e_ms = float(ems.item()) / args.steps
        e2e = {"value": 1e3 / e2e_ms, "unit": unit, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(nbytes.item()),
               "d2h_bytes_per_step": int(nbytes.item()),
               "note": "bytes summed over ranks; every rank copies its own shard from/to pinned host memory"}
    if rank == 0:
        clocks = sampler.stop()
        ms_per_step = float(ms.item()) / args.steps
        roofline = None
        if args.exchange == "peer":
            nlat0 = st.trans.nlat0()
            nleg = (grid.ny() + 1) // 2
            fl = [B.legendre_flops(nlat0, T, nleg, nf, T, [m for m in range(T + 1) if owner[m] == 0]),
                  sum(2.0 * nf * (1 if m == 0 else 2) * (T - m + 1) * max(0, nleg - int(nlat0[m])) for m in range(T + 1) if owner[m] == 0)]
            leg = [float(stage_all[0][0]), float(stage_all[0][5])]
            ach = (fl[0] + fl[1]) / ((leg[0] + leg[1]) * 1e-3) / 1e12
            roofline = {"kernel": "legendre_dmma_kernel<inverse + peer stores | direct> on rank 0 (its share of the zonal wavenumbers; "
                                  "launch time includes the spectra pack / unpack kernel)",
                        "bound": "tensor", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                        "traffic": None, "flops_per_launch": {"inverse": fl[0], "direct": fl[1]},
                        "ms_per_launch": {"inverse": leg[0], "direct": leg[1]}, "share_of_step": (leg[0] + leg[1]) / ms_per_step,
                        "peak_source": "fp64 DMMA/DFMA microbenchmark on this pool's B200 (profiles/microbench_f64_r01.txt)"}
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
            "e2e": e2e, "roofline": roofline, "cpu_baseline": None,
            "stage_ms_per_rank": {k: [round(float(x[i]), 3) for x in stage_all] for i, k in enumerate(
                ["inv_legendre", "inv_exchange_wait", "inv_fourier", "dir_fourier_push", "dir_exchange_wait", "dir_legendre"])},
            "call_ms_per_rank": {"invtrans": [round(float(x[0]), 3) for x in calls_all], "dirtrans": [round(float(x[1]), 3) for x in calls_all]},
        }
    dist.destroy_process_group()
    return out if rank == 0 else None
