#!/usr/bin/env python
"""Benchmark of the spectral-transform hot path (BASELINE.json metric: TCo1279 L137 transforms/sec, inv+dir).

One "step" = one inverse + one direct transform of 137 scalar fields on the octahedral grid O1280 at
truncation 1279 (fp64), synthetic spectra (PCG64 seed 20260925, N(0,1)(1+n)^-1.5).

  value : steps/s with inputs and outputs resident in HBM (device pointers handed to the C ABI)
  e2e   : steps/s through the same C-ABI calls with pinned HOST buffers; the H2D copy of every input and
          the D2H copy of every output happen inside the timed region.  Headline e2e = asynchronous calls
          (sptrans_set_async) on a plan and its clone: the inverse of step i and the direct transform of the grid
          fields of step i-1 are in flight together, so both directions of the PCIe link are busy;
          e2e.blocking = the same two calls issued blocking, one after the other, on the same data
  roofline : the dominant kernel (fp64 DMMA Legendre GEMM): algorithmic flops / CUDA-event time of that kernel
  cpu_baseline : the CPU oracle (restatement of TransLocal; the reference itself cannot be built here) on the
          host cores, on a bounded sample, reported next to the GPU number

`--impl reference` times the CPU oracle only (all host threads) and prints the same JSON shape.
Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under torch.distributed.run (one rank/GPU).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

METRIC = "TCo1279 L137 transforms/sec (inv+dir)"  # BASELINE.json's metric; other --workload values rename it (metric_name)
UNIT = "transforms/s"
FP64_DMMA_PEAK_TFLOPS = 37.1  # measured on this pool's B200 (profiles/microbench_f64_r01.txt); no fp64 entry in MEASURED_PEAKS.json


CPU_NOTE = ("restatement of TransLocal (+ the quadrature-adjoint dirtrans), OpenMP over zonal wavenumbers and over (field, latitude); "
            "no BLAS in the image, so the Legendre GEMM is a hand-written AVX-512 / AVX2 micro-kernel (20-30 GFLOP/s per thread "
            "measured; an MKL/OpenBLAS eckit backend would be ~2x that).  At TCo1279 L137 the GEMM is ~15 % of the CPU time; the "
            "rest is what the reference does as well: allocating and zero-filling the 7.2 GB Fourier work array per call "
            "(TransLocal.cc:1426-1428), the split / merge loops (:970-1079) and 350 720 FFTs per direction (:1155-1196)")


def metric_name(workload_name):
    _, _, nf = workload(workload_name)
    return f"{workload_name} L{nf} transforms/sec (inv+dir)"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/ncu_traffic_r01.json:
    dram__bytes_read.sum + dram__bytes_write.sum, mean of the inverse and the direct launch), or None."""
    try:
        t = json.load(open(os.path.join(REPO, "profiles", "ncu_traffic_r01.json")))[kernel]
        return 0.5 * (float(t["inverse"]) + float(t["direct"]))
    except Exception:
        return None


def host_threads():
    """Threads for the CPU arm: every core this process may run on.  (Not omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would silently turn the reference arm into a single-thread run; the oracle's
    parallel regions take an explicit num_threads.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def workload(name):
    table = {
        "TCo2559": ("O2560", 2559, 137),
        "TCo1279": ("O1280", 1279, 137),
        "TCo399": ("O400", 399, 137),
        "TCo159": ("O160", 159, 137),
        "O32": ("O32", 31, 4),
    }
    return table[name]


def legendre_flops(nlat0, T, nleg, nf, trunc, ms=None):
    """Algorithmic (pruned) Legendre flops of one direction, SURVEY 8(d):
    sum_m 2 * (nf * nimag(m)) * (K_s + K_a) * (nleg - nlat0[m]) with K counted up to `trunc`
    (over the zonal wavenumbers `ms` of one rank, all by default)."""
    tot = 0.0
    for m in (range(T + 1) if ms is None else ms):
        if m >= trunc and trunc == T:  # scalar inverse drops m == T (TransLocal.cc:982)
            continue
        nimag = 1 if m == 0 else 2
        K = trunc - m + 1
        tot += 2.0 * nf * nimag * K * max(0, nleg - int(nlat0[m]))
    return tot


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line): NVML polled
    every 10 ms from a thread of this process (nvidia_ml_py), nvidia-smi -lms as the fallback."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index=0, period_s=0.01):
        self.index = index
        self.period_s = period_s
        self.sm, self.mx, self.bits = [], [], 0
        self.proc = None
        self.nvml = None
        self._stop = threading.Event()

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
            self._poll_once()
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_once(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            self.bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            self.bits |= int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))

    def _poll(self):
        while not self._stop.wait(self.period_s):
            try:
                self._poll_once()
            except Exception:
                break

    def _read(self):
        names = {3: 0x8, 4: 0x40, 5: 0x20, 6: 0x4}
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) >= 7 and r[0].replace(".", "").isdigit():
                self.sm.append(float(r[0]))
                self.mx.append(float(r[1]))
                for i, bit in names.items():
                    if r[i].lower().startswith("active"):
                        self.bits |= bit

    def stop(self):
        if self.nvml is None and not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        self._stop.set()
        if self.nvml is not None:
            try:
                self._poll_once()
            except Exception:
                pass
            self.thread.join(timeout=1)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        reasons = sorted(name for bit, name in self.REASONS.items() if self.bits & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": reasons, "samples": len(self.sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def cpu_sample_fields(args, T, npts, nf):
    """Fields of the CPU sample: all of them (nothing extrapolated) unless --cpu-fields asks for fewer or the host cannot
    hold the oracle's working set (Legendre tables + Fourier work array + grid fields + spectra)."""
    nfs = nf if args.cpu_fields <= 0 else min(nf, args.cpu_fields)
    try:
        import psutil

        avail = psutil.virtual_memory().available
        per_field = 16.0 * (T + 1) * 2 * ((T + 2) // 2 * 2) + 8.0 * npts * 2 + 8.0 * (T + 1) * (T + 2) * 2
        tables = 8.0 * (T + 2) ** 3 / 2 * 2.2
        while nfs > 1 and tables + per_field * nfs > 0.8 * avail:
            nfs = max(1, nfs // 2)
    except Exception:
        pass
    return nfs


def run_reference(args, rank, world):
    """CPU arm: the oracle port of TransLocal (+ the quadrature-adjoint dirtrans) on all host threads, on the SAME
    configuration as the GPU arm: every step is one inverse + one direct transform of all the workload's fields."""
    if rank != 0:
        return
    import helpers as H
    from oracle import pyoracle as po

    gridname, T, nf = workload(args.workload)
    N = int(gridname[1:])
    lat, w = po.gaussian_quadrature(N)
    nx = np.array([20 + 4 * j for j in range(N)] + [20 + 4 * j for j in range(N - 1, -1, -1)], dtype=np.int32)
    threads = host_threads()
    t0 = time.time()
    plan = po.OraclePlan(nx, lat, T, weights=w, nthreads=threads)
    setup_s = time.time() - t0
    nfs = cpu_sample_fields(args, T, int(nx.sum()), nf)
    sp = H.synthetic_spectra(T, nfs)
    times = []
    t_begin = time.time()
    for it in range(args.warmup + args.steps):
        t0 = time.time()
        gp = plan.invtrans(nfs, sp, mode=2)
        back = plan.dirtrans(nfs, gp)
        dt = time.time() - t0
        if it >= args.warmup:
            times.append(dt)
        # the whole arm has to end within a few minutes whatever K the driver asks for: once two timed steps are in
        # and the budget is spent, the remaining (identical) steps are not run; the count actually timed is reported
        if len(times) >= 2 and time.time() - t_begin > args.cpu_budget_s:
            break
    per_step = float(np.mean(times)) * nf / nfs  # (nfs == nf unless the host is too small: then linear in the fields)
    value = 1.0 / per_step
    sample = (f"all {nf} fields, full grid, inv+dir per step, nothing extrapolated" if nfs == nf else
              f"{nfs} of {nf} fields, full grid, inv+dir, time scaled x{nf}/{nfs}")
    out = {
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload} L{nf} invtrans+dirtrans fp64 (grid {gridname}, T{T})"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                         "sample": f"{sample}; {len(times)} timed steps of {args.steps} requested (min {min(times):.2f}s, "
                                   f"max {max(times):.2f}s); plan setup {setup_s:.1f}s untimed",
                         "note": CPU_NOTE},
        "steps_timed": len(times),
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json(out)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


_JSON_FD = None


def emit_json(obj):
    """The ONE JSON line of the contract, on the real stdout (libraries that chat on fd 1, e.g. NCCL's version
    banner, are diverted to stderr by main())."""
    line = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, line)


def mask_mT(sp, T, nf):
    """copy of a [m][n][re/im][field] array with the m == T column zeroed (the scalar inverse ignores it, TransLocal.cc:982)"""
    out = np.array(sp, copy=True)
    out.reshape(-1, 2, nf)[-1] = 0.0
    return out


def bench_e2e(args, trans, torch, H, T, nf, npts, nspec, sp_host):
    """End to end through the C ABI on pinned HOST buffers, copies inside the timed region.

    blocking : trans.invtrans(host spectra -> host grid fields); trans.dirtrans(the same host grid fields -> host spectra),
               each call blocking like the reference's (results visible on return).  Inside a call the grid-side copy runs
               in field chunks next to the Fourier kernels, but the two calls cannot overlap each other.
    pipelined: the same two calls issued asynchronously (sptrans_set_async) on a plan and its clone (sptrans_plan_clone:
               tables shared).  Step i transforms spectra -> grid fields[i % 3] and, in flight at the same time, grid
               fields[(i - 1) % 3] -> spectra: the direct transform consumes what the previous step's inverse produced,
               so H2D and D2H traffic cross the PCIe link together.  The dependencies between the two plans are device-side
               marks (sptrans_mark / sptrans_wait_mark); the host synchronises at the end of the timed region.
    The round trip is asserted: the spectra that come back equal the input with the m == T column masked."""
    nsteps = args.steps
    gp_h = [torch.empty(nf * npts, dtype=torch.float64).pin_memory() for _ in range(2)]
    sp2_host = torch.empty(nspec, dtype=torch.float64).pin_memory()
    sp_np, sp2_np = sp_host.numpy(), sp2_host.numpy()
    gp_np = [g.numpy() for g in gp_h]
    want = mask_mT(sp_np, T, nf)
    scale = float(np.abs(want).max())

    tol = 1e-11 if args.precision == "fp64" else 1e-4   # tf32x3 Legendre stage: fp32-level accuracy (2e-6 per transform)

    def check(tag):
        err = float(np.abs(sp2_np - want).max()) / scale
        if not err < tol:
            raise SystemExit(f"bench e2e ({tag}): round trip through host buffers differs from the input spectra "
                             f"(m == T column masked): rel max {err:.3e}")
        return err

    # blocking
    for _ in range(max(1, min(args.warmup, 2))):
        trans.invtrans(nf, sp_np, gp_np[0])
        trans.dirtrans(nf, gp_np[0], sp2_np)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.time()
    for _ in range(nsteps):
        trans.invtrans(nf, sp_np, gp_np[0])   # H2D spectra, transform, D2H grid fields
        trans.dirtrans(nf, gp_np[0], sp2_np)  # H2D grid fields, transform, D2H spectra
    blocking_ms = (time.time() - t0) * 1e3 / nsteps
    e1.record()
    torch.cuda.synchronize()
    err_blocking = check("blocking")
    # pipelined: inverse on the plan, direct on its clone, both asynchronous; dependencies between the two plans are marks
    # on the device (sptrans_mark / sptrans_wait_mark), the host synchronises once, at the end of the timed region
    other = trans.clone()
    trans.set_async(True)
    other.set_async(True)
    gp_h.append(torch.empty(nf * npts, dtype=torch.float64).pin_memory())
    gp_np.append(gp_h[2].numpy())

    def run(nsteps_):
        """step i: inverse spectra -> grid[i % 3] on the plan; direct grid[(i - 1) % 3] -> spectra on the clone (waits for the
        inverse of step i-1); the inverse of step i waits for the direct transform of step i-2, the last reader of its buffer"""
        marks_a, marks_b = {}, {}
        marks_a[-1] = trans.mark()                   # grid[(0 - 1) % 3] was primed (and synchronised) by the caller
        for i in range(nsteps_):
            if i - 2 in marks_b:
                trans.wait_mark(marks_b[i - 2])
            trans.invtrans(nf, sp_np, gp_np[i % 3])
            marks_a[i] = trans.mark()
            other.wait_mark(marks_a[i - 1])
            other.dirtrans(nf, gp_np[(i - 1) % 3], sp2_np)
            marks_b[i] = other.mark()
        trans.synchronize()
        other.synchronize()
        for m in list(marks_a.values()) + list(marks_b.values()):
            trans.release_mark(m)

    trans.invtrans(nf, sp_np, gp_np[2])   # untimed: the grid fields "step -1" would have produced
    trans.synchronize()
    run(2)
    sp2_host.zero_()
    trans.invtrans(nf, sp_np, gp_np[2])
    trans.synchronize()
    t0 = time.time()
    run(nsteps)                           # exactly nsteps inverse + nsteps direct transforms, every copy inside
    pipelined_ms = (time.time() - t0) * 1e3 / nsteps   # host wall clock, everything synchronised at both ends
    err_pipe = check("pipelined")
    trans.set_async(False)
    del other
    bytes_step = 8 * (nspec + nf * npts)
    return {"value": 1e3 / pipelined_ms, "unit": UNIT, "ms_per_step": pipelined_ms,
            "h2d_bytes_per_step": bytes_step, "d2h_bytes_per_step": bytes_step,
            "mode": "asynchronous C-ABI calls on a plan and its clone (inverse of step i and direct transform of step i-1's grid "
                    "fields in flight together, ordered by device-side marks), pinned host buffers, host synchronised at both ends "
                    "of the timed region (K inverse + K direct transforms)",
            "pcie_gbs_each_way": bytes_step / pipelined_ms / 1e6,
            "roundtrip_rel_max": err_pipe,
            "blocking": {"value": 1e3 / blocking_ms, "ms_per_step": blocking_ms, "roundtrip_rel_max": err_blocking,
                         "mode": "blocking invtrans then dirtrans on the same host arrays (reference call semantics)"}}


def rank_spectra(T, nf, ms, seed=20260925):
    """Synthetic spectra of the zonal wavenumbers `ms` only, in the shard-local layout [m in ms][n][re/im][field]
    (same law as helpers.synthetic_spectra -- N(0,1)(1+n)^-1.5, Im(m=0)=0 -- but one independent stream per m, so that a
    rank can generate its own share without materialising the global array: 7.2 GB at TCo2559 L137)."""
    parts = []
    for m in ms:
        rng = np.random.Generator(np.random.PCG64([seed, m]))
        a = rng.standard_normal((T - m + 1, 2, nf))
        a *= ((1.0 + np.arange(m, T + 1)) ** -1.5)[:, None, None]
        if m == 0:
            a[:, 1, :] = 0.0
        parts.append(a.reshape(-1))
    return np.ascontiguousarray(np.concatenate(parts)) if parts else np.zeros(0)


def bench_sharded(args, rank, world, local_rank):
    """bench.py for N > 1 (strong scaling: the job is fixed, split over N GPUs).  Every rank holds only its own share of
    the spectra (its zonal wavenumbers) and of the grid fields (its latitude band): SPTRANS_SHARD_LOCAL_IO."""
    import torch
    import torch.distributed as dist

    import atlas_b200
    import helpers as H
    from atlas_b200.dist import ShardedTrans

    inv_only = args.direction == "inv"
    gridname, T, nf = workload(args.workload)
    grid = atlas_b200.Grid(gridname)
    t0 = time.time()
    st = ShardedTrans(grid, T, local_rank, exchange=args.exchange, local_io=True)
    setup_s = time.time() - t0
    dev = st.device
    npts = grid.size()
    nspec_loc, stride = st.trans.local_sizes()
    my_m = st.my_m()
    big = args.workload == "TCo2559"
    sp_loc = rank_spectra(T, nf, my_m) if big else st.local_spectra(H.synthetic_spectra(T, nf), nf)
    assert sp_loc.size == nspec_loc * nf
    h_sp = torch.from_numpy(sp_loc).pin_memory()
    d_sp = h_sp.to(dev)
    d_gp = torch.zeros(nf * stride, dtype=torch.float64, device=dev)
    d_sp2 = torch.zeros_like(d_sp)

    def step():
        st.invtrans(nf, d_sp, d_gp)
        if not inv_only:
            st.dirtrans(nf, d_gp, d_sp2)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    launches0 = st.trans.kernel_launches()
    # (polled less often than in the single-GPU bench: NVML queries from rank 0's process contend with its kernel launches,
    # and every rank waits for rank 0 at the exchange barriers)
    sampler = ClockSampler(local_rank, period_s=0.05)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = torch.tensor([st.trans.kernel_launches() - launches0], device=dev, dtype=torch.float64)
    dist.all_reduce(launches)
    clocks = sampler.stop() if rank == 0 else None

    # ---- parity of the sharded result (outside the timed region) ---------------------------------------------------------
    parity = sharded_parity(args, st, torch, dist, H, grid, T, nf, d_sp, d_gp, d_sp2, rank, world, inv_only, big)

    # stage times of the last inverse + direct pair on every rank (outside the timed region)
    stage = torch.zeros(6, device=dev, dtype=torch.float64)
    if args.exchange == "peer":
        st.invtrans(nf, d_sp, d_gp)
        ti = st.trans.last_timings()
        td = {"legendre": 0.0, "exchange_wait": 0.0, "fourier": 0.0}
        if not inv_only:
            st.dirtrans(nf, d_gp, d_sp2)
            td = st.trans.last_timings()
        stage = torch.tensor([ti["legendre"], ti["exchange_wait"], ti["fourier"], td["fourier"], td["exchange_wait"], td["legendre"]],
                             device=dev, dtype=torch.float64)
    stage_all = [torch.zeros_like(stage) for _ in range(world)]
    dist.all_gather(stage_all, stage)
    # steady-state duration of the two calls on every rank (back to back, no host synchronisation), again untimed
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    dist.barrier()
    for e in evs:
        e[0].record()
        st.invtrans(nf, d_sp, d_gp)
        e[1].record()
        if not inv_only:
            st.dirtrans(nf, d_gp, d_sp2)
        e[2].record()
    torch.cuda.synchronize()
    calls = torch.tensor([float(np.median([e[0].elapsed_time(e[1]) for e in evs])), float(np.median([e[1].elapsed_time(e[2]) for e in evs]))],
                         device=dev, dtype=torch.float64)
    calls_all = [torch.zeros_like(calls) for _ in range(world)]
    dist.all_gather(calls_all, calls)

    # ---- optional: the all-gather that replicates the band-distributed grid fields (north star's "single allgather") ----
    gather = None
    if args.gather:
        d_glob = torch.empty(nf * npts, dtype=torch.float64, device=dev)
        st.gather_grid(nf, d_gp, d_glob)
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(3):
            st.gather_grid(nf, d_gp, d_glob)
        g1.record()
        torch.cuda.synchronize()
        gms = torch.tensor([g0.elapsed_time(g1) / 3], device=dev, dtype=torch.float64)
        dist.all_reduce(gms, op=dist.ReduceOp.MAX)
        gather = {"ms": float(gms.item()), "bytes_received_per_rank": 8.0 * nf * npts * (world - 1) / world,
                  "what": "all_gather_into_tensor of every rank's latitude band (padded) + reassembly into [field][point]"}
        del d_glob

    # ---- end to end with pinned HOST buffers: every rank moves its own share (the spectra of its zonal wavenumbers, the
    # grid rows of its latitude band) over its own PCIe link as ONE contiguous copy per array, inside the timed region.
    # Software-pipelined like the single-GPU e2e: the direct transform of step i consumes the grid fields that step i-1's
    # inverse produced (other host buffer), so its H2D copy crosses the bus while this step's D2H copy does.
    e2e = None
    if not args.no_e2e:
        h_sp2 = torch.zeros_like(h_sp).pin_memory()
        h_gp = [torch.zeros(nf * stride, dtype=torch.float64).pin_memory() for _ in range(2)]
        d_gp_in = torch.zeros_like(d_gp)
        main_s = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        ev_in, ev_inv, ev_used = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()

        def e2e_step(i):
            d_sp.copy_(h_sp, non_blocking=True)
            if not inv_only:
                s_in.wait_event(ev_used)              # previous direct transform has consumed d_gp_in
                s_in.wait_stream(s_out)               # ... and the previous step's grid fields have landed in host memory
                with torch.cuda.stream(s_in):
                    d_gp_in.copy_(h_gp[(i + 1) % 2], non_blocking=True)
                    ev_in.record(s_in)
            st.invtrans(nf, d_sp, d_gp)
            ev_inv.record(main_s)
            s_out.wait_event(ev_inv)
            with torch.cuda.stream(s_out):
                h_gp[i % 2].copy_(d_gp, non_blocking=True)
            if not inv_only:
                main_s.wait_event(ev_in)
                st.dirtrans(nf, d_gp_in, d_sp2)
                ev_used.record(main_s)
                h_sp2.copy_(d_sp2, non_blocking=True)
            main_s.wait_stream(s_out)

        ev_used.record(main_s)
        for i in range(2):
            e2e_step(i)
        torch.cuda.synchronize()
        dist.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(args.steps):
            e2e_step(i)
        g1.record()
        torch.cuda.synchronize()
        dist.barrier()
        ems = torch.tensor([g0.elapsed_time(g1)], device=dev, dtype=torch.float64)
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        h2d = 8.0 * (sp_loc.size + (0 if inv_only else nf * stride))
        d2h = 8.0 * (nf * stride + (0 if inv_only else sp_loc.size))
        nb = torch.tensor([h2d, d2h], device=dev, dtype=torch.float64)
        dist.all_reduce(nb)
        e2e_ms = float(ems.item()) / args.steps
        # the grid fields that came back through the host equal the device result
        back_err = float((h_gp[(args.steps - 1) % 2].to(dev) - d_gp).abs().max())
        e2e = {"value": 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(nb[0].item()),
               "d2h_bytes_per_step": int(nb[1].item()), "host_copy_max_abs_diff": back_err,
               "aggregate_host_gbs": {"h2d": float(nb[0].item()) / e2e_ms / 1e6, "d2h": float(nb[1].item()) / e2e_ms / 1e6,
                                      "note": "all ranks together; on this pool's boxes the sum saturates near 55-60 GB/s per direction "
                                              "whatever the number of GPUs (measured at N = 1, 2, 8), so the end-to-end figure stops scaling"},
               "note": "bytes summed over ranks; every rank copies its own shard from/to pinned host memory as one contiguous "
                       "copy per array; the direct transform of step i reads the grid fields step i-1's inverse produced"}
    out = None
    if rank == 0:
        ms_per_step = float(ms.item()) / args.steps
        roofline = None
        nlat0 = st.trans.nlat0()
        nleg = (grid.ny() + 1) // 2
        if args.exchange == "peer":
            own0 = [m for m in range(T + 1) if st._owner[m] == 0]
            fl = [legendre_flops(nlat0, T, nleg, nf, T, own0),
                  0.0 if inv_only else sum(2.0 * nf * (1 if m == 0 else 2) * (T - m + 1) * max(0, nleg - int(nlat0[m])) for m in own0)]
            leg = [float(stage_all[0][0]), float(stage_all[0][5])]
            ach = (fl[0] + fl[1]) / ((leg[0] + leg[1]) * 1e-3) / 1e12
            roofline = {"kernel": "legendre_dmma_kernel<inverse + peer stores | direct> on rank 0 (its share of the zonal wavenumbers; "
                                  "launch time includes the spectra pack / unpack kernel)",
                        "bound": "tensor", "achieved": ach, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_DMMA_PEAK_TFLOPS,
                        "traffic": None, "flops_per_launch": {"inverse": fl[0], "direct": fl[1]},
                        "ms_per_launch": {"inverse": leg[0], "direct": leg[1]}, "share_of_step": (leg[0] + leg[1]) / ms_per_step,
                        "peak_source": "fp64 DMMA/DFMA microbenchmark on this pool's B200 (profiles/microbench_f64_r01.txt)"}
        try:
            hbm = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0))
        except Exception:
            hbm = 6650.0
        # Fourier stage of rank 0: bytes of its band = its rows of the exchange buffer + its grid rows
        j0, j1 = st.band_rows_slice()
        mm = np.arange(T + 1)
        fb_rows0 = sum(2 * max(0, j1 - max(j0, int(nlat0[m]))) for m in mm)
        four_bytes0 = 16.0 * nf * fb_rows0 + 8.0 * nf * stride
        f_ms = [float(stage_all[0][2]), float(stage_all[0][3])]
        roofline_fourier = {"kernel": "direct mixed-radix / register-tiled chirp-z / row-mode Fourier kernels on rank 0 (its latitude band)",
                            "bound": "hbm", "paths": st.trans.fourier_paths(),
                            "achieved": (1 if inv_only else 2) * four_bytes0 / (max(f_ms[0] + f_ms[1], 1e-9) * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                            "bytes_per_launch_group": four_bytes0, "ms_per_launch_group": {"inverse": f_ms[0], "direct": f_ms[1]}}
        roofline_fourier["frac"] = roofline_fourier["achieved"] / hbm
        what = "invtrans" if inv_only else "invtrans+dirtrans"
        out = {
            "metric": metric_name(args.workload) if not inv_only else f"{args.workload} L{nf} invtrans/sec", "value": 1e3 / ms_per_step,
            "unit": "invtrans/s" if inv_only else UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload} L{nf} {what} fp64 (grid {gridname}, T{T})",
                       "parallelism": f"zonal-wavenumber sharded Legendre x latitude-band sharded Fourier over {world} GPUs; "
                                      + ("exchange fused into the kernels over NVLink peer memory (Legendre epilogue stores / "
                                         "push kernel, device-side barrier)" if args.exchange == "peer"
                                         else "one NCCL all-to-all per direction") + "; every rank holds only its share of the spectra "
                                      "and of the grid fields (SPTRANS_SHARD_LOCAL_IO); grid fields stay band-distributed",
                       "l2": "inputs larger than L2", "plan_setup_s": setup_s, "device_bytes_rank0": st.trans.device_bytes()},
            "clocks": clocks, "gpu_launches": int(launches.item()),
            "e2e": e2e, "roofline": roofline, "roofline_fourier": roofline_fourier, "cpu_baseline": None, "parity": parity,
            "gather_grid": gather,
            "stage_ms_per_rank": {k: [round(float(x[i]), 3) for x in stage_all] for i, k in enumerate(
                ["inv_legendre", "inv_exchange_wait", "inv_fourier", "dir_fourier_push", "dir_exchange_wait", "dir_legendre"])},
            "call_ms_per_rank": {"invtrans": [round(float(x[0]), 3) for x in calls_all], "dirtrans": [round(float(x[1]), 3) for x in calls_all]},
        }
    dist.barrier()
    dist.destroy_process_group()
    if parity["max_rel"] is None or not parity["max_rel"] < parity["tolerance"]:
        raise SystemExit(f"bench: sharded result failed its parity check: {parity}")
    return out


def sharded_parity(args, st, torch, dist, H, grid, T, nf, d_sp, d_gp, d_sp2, rank, world, inv_only, big):
    """Correctness of what the timed region computed, checked on every rank against an independent result:
    * TCo1279 and smaller: rank 0 runs the SINGLE-GPU path of the same library (held to the CPU oracle by the -m gpu tests)
      on three sampled fields of the same spectra; every rank compares its band of the grid fields and its zonal
      wavenumbers of the direct transform with it;
    * TCo2559 (the unsharded tables do not fit next to the shard): closed-form sectoral harmonics
      (src/tests/trans/test_transgeneral.cc:80-374: Pbar_m^m cos/sin(m lon)) placed in the sampled fields, compared on
      every rank's own rows."""
    import atlas_b200

    dev = st.device
    npts = grid.size()
    nspec_loc, stride = st.trans.local_sizes()
    fields = sorted({0, nf // 2, nf - 1})
    tol = 1e-12
    spans = st._band_spans(st.rank)
    if big:
        # unit sectoral coefficients (m, n = m, re or im) in the sampled fields, zero elsewhere in those fields
        sets = [[(0, 0), (T // 3, 0)], [(1, 0), (T - 1, 1)], [(17, 1), (2, 0)]][:len(fields)]
        sp_chk = d_sp.clone().view(-1, nf)
        sp_chk[:, fields] = 0.0
        off = 0
        offs = {}
        for m in st.my_m():
            offs[m] = off
            off += (T - m + 1) * 2
        for f, harmonics in zip(fields, sets):
            for m, imag in harmonics:
                if m in offs:
                    sp_chk[offs[m] + imag, f] = 1.0
        gp_chk = torch.zeros_like(d_gp)
        st.invtrans(nf, sp_chk.view(-1), gp_chk)
        torch.cuda.synchronize()
        got = gp_chk.view(nf, stride)[fields].cpu().numpy()
        del gp_chk, sp_chk
        lon, lat = H.grid_lonlat(grid.nx(), grid.y())
        from atlas_b200 import _lib

        err = 0.0
        for k, harmonics in enumerate(sets):
            want = np.zeros(npts)
            for m, imag in harmonics:
                mask = H.expected_zonal_mask(T, grid.nx(), grid.y(), grid.regular, m, _lib.lib.sptrans_fourier_truncation)
                want += H.analytic_harmonic(m, m, imag, lon, lat) * mask
            o = 0
            scale = max(np.abs(want).max(), 1.0)
            for a, b in spans:
                err = max(err, float(np.abs(got[k, o:o + b - a] - want[a:b]).max()) / scale)
                o += b - a
        e = torch.tensor([err], device=dev, dtype=torch.float64)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        return {"max_rel": float(e.item()), "tolerance": 1e-11, "fields": fields,
                "against": "closed-form sectoral harmonics (test_transgeneral.cc:80-374) on every rank's own rows"}
    # reference: the single-GPU path on rank 0
    nsel = len(fields)
    ref_gp = torch.empty(nsel * npts, dtype=torch.float64, device=dev)
    ref_sp = torch.empty((T + 1) * (T + 2) * nsel, dtype=torch.float64, device=dev)
    if rank == 0:
        sp_g = torch.from_numpy(np.ascontiguousarray(H.synthetic_spectra(T, nf).reshape(-1, nf)[:, fields]).reshape(-1)).to(dev)
        one = atlas_b200.Trans(grid, T, atlas_b200.option.type("b200"), device=dev.index)
        one.invtrans(nsel, sp_g, ref_gp)
        if not inv_only:
            one.dirtrans(nsel, ref_gp, ref_sp)
        torch.cuda.synchronize()
        del one
    dist.broadcast(ref_gp, 0)
    dist.broadcast(ref_sp, 0)
    torch.cuda.synchronize()
    got = d_gp.view(nf, stride)[fields]
    ref2 = ref_gp.view(nsel, npts)
    scale = float(ref2.abs().max())
    err_g, o = 0.0, 0
    for a, b in spans:
        err_g = max(err_g, float((got[:, o:o + b - a] - ref2[:, a:b]).abs().max()) / scale)
        o += b - a
    err_s = 0.0
    if not inv_only:
        gs = d_sp2.view(-1, nf)[:, fields]
        rs = ref_sp.view(-1, nsel)
        sscale = float(rs.abs().max())
        o = 0
        for m in st.my_m():
            n = (T - m + 1) * 2
            g0 = (2 * T + 3 - m) * m // 2 * 2
            err_s = max(err_s, float((gs[o:o + n] - rs[g0:g0 + n]).abs().max()) / sscale)
            o += n
    e = torch.tensor([err_g, err_s], device=dev, dtype=torch.float64)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    return {"max_rel": float(e.max().item()), "invtrans_rel_max": float(e[0].item()), "dirtrans_rel_max": float(e[1].item()),
            "tolerance": tol, "fields": fields,
            "against": "single-GPU path of the same library on rank 0 (itself held to the CPU oracle by tests/test_gpu_parity.py), "
                       "compared on every rank's own latitude band / zonal wavenumbers after the timed region"}


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="TCo1279", choices=["TCo2559", "TCo1279", "TCo399", "TCo159", "O32"])
    ap.add_argument("--cpu-fields", type=int, default=0, help="fields in the CPU sample (0 = all fields of the workload: nothing extrapolated)")
    ap.add_argument("--cpu-budget-s", type=float, default=300.0, help="reference arm: stop after >= 2 timed steps once this much time is spent")
    ap.add_argument("--direction", default="both", choices=["both", "inv"],
                    help="N > 1: 'inv' times the inverse transform only (BASELINE config 5: TCo2559 L137 invtrans)")
    ap.add_argument("--gather", action="store_true", help="N > 1: also time the all-gather that replicates the grid fields")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "tc"],
                    help="Legendre arithmetic: fp64 DMMA (headline) or tcgen05 split-TF32 (BASELINE config 4; fp32-level accuracy)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer-memory exchange fused into the kernels (default) or one NCCL all-to-all")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    import atlas_b200
    import helpers as H

    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        out = bench_sharded(args, rank, world, local_rank)
        if out is not None:
            emit_json(out)
        return

    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())  # a non-blocking stream of our own: no implicit joins with the copy streams
    gridname, T, nf = workload(args.workload)
    grid = atlas_b200.Grid(gridname)
    t0 = time.time()
    trans = atlas_b200.Trans(grid, T, atlas_b200.option.type("b200"), device=0)
    if args.precision == "tc":
        trans.set_precision("tc")
    setup_s = time.time() - t0
    npts = grid.size()
    nspec = (T + 1) * (T + 2) * nf

    sp_host = torch.from_numpy(H.synthetic_spectra(T, nf)).pin_memory()
    d_sp = sp_host.cuda()
    d_gp = torch.empty(nf * npts, dtype=torch.float64, device="cuda")
    d_sp2 = torch.empty_like(d_sp)

    def step_device():
        trans.invtrans(nf, d_sp, d_gp)
        trans.dirtrans(nf, d_gp, d_sp2)

    # ---- device-resident timing -------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    leg_ms, four_ms, pack_ms = [], [], []
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = trans.kernel_launches()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream_ptr = torch.cuda.current_stream().cuda_stream
    trans.set_stream(stream_ptr)  # run the library on torch's current stream so torch events bracket it
    ev0.record()
    for _ in range(args.steps):
        trans.invtrans(nf, d_sp, d_gp)
        ti = trans.last_timings()
        trans.dirtrans(nf, d_gp, d_sp2)
        td = trans.last_timings()
        leg_ms.append((ti["legendre"], td["legendre"]))
        four_ms.append((ti["fourier"], td["fourier"]))
        pack_ms.append((ti["pack"], td["pack"]))
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = trans.kernel_launches() - launches0
    total_ms = ev0.elapsed_time(ev1)
    ms_per_step = total_ms / args.steps
    value = 1e3 / ms_per_step

    # ---- roofline of the dominant kernel (Legendre DMMA GEMM, both directions) ----------------------------
    nlat0 = trans.nlat0()
    nleg = (grid.ny() + 1) // 2
    fl_inv = legendre_flops(nlat0, T, nleg, nf, T)
    fl_dir = legendre_flops(nlat0, T, nleg, nf, T + 1) if False else sum(
        2.0 * nf * (1 if m == 0 else 2) * (T - m + 1) * max(0, nleg - int(nlat0[m])) for m in range(T + 1))
    leg_inv = float(np.mean([a for a, _ in leg_ms]))
    leg_dir = float(np.mean([b for _, b in leg_ms]))
    achieved = (fl_inv + fl_dir) / ((leg_inv + leg_dir) * 1e-3) / 1e12
    if args.precision == "tc":
        # 3 tf32 MMAs per algorithmic product; peak: dense tf32 = half the measured bf16 rate
        try:
            tf32_peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["bf16_tflops"]) / 2.0
            src = "half of the measured bf16 cuBLAS rate in MEASURED_PEAKS.json (tf32 runs at half the bf16 rate)"
        except Exception:
            tf32_peak, src = 1590.0 / 2.0, "half of the fallback bf16 figure"
        # (stage timings of the tensor-core path: "legendre" is the tcgen05 kernel alone, the kernels that build the split-tf32
        # operand images are under "pack")
        roofline = {"kernel": "legendre_tc_kernel (tcgen05.mma kind::tf32, 3 split products per MAC, TMEM fp32 accumulators)",
                    "bound": "tensor", "achieved": 3.0 * achieved, "peak": tf32_peak, "unit": "TFLOP/s",
                    "frac": 3.0 * achieved / tf32_peak, "traffic": None, "peak_source": src,
                    "algorithmic_tflops": achieved, "flops_per_launch": {"inverse": fl_inv, "direct": fl_dir},
                    "ms_per_launch": {"inverse": leg_inv, "direct": leg_dir}, "share_of_step": (leg_inv + leg_dir) / ms_per_step}
    else:
      roofline = {
        "kernel": "legendre_dmma_kernel<inverse|direct> (fp64 mma.sync m8n8k4, operands staged by TMA bulk copies)", "bound": "tensor", "achieved": achieved,
        "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_DMMA_PEAK_TFLOPS,
        "traffic": ncu_traffic("legendre_dmma_kernel") if args.workload == "TCo1279" else None,
        "peak_source": "fp64 DMMA/DFMA microbenchmark on this pool's B200 (profiles/microbench_f64_r01.txt); "
                       "MEASURED_PEAKS.json holds only bf16 and HBM peaks, tcgen05 has no f64 kind",
        "flops_per_launch": {"inverse": fl_inv, "direct": fl_dir},
        "ms_per_launch": {"inverse": leg_inv, "direct": leg_dir},
        "share_of_step": (leg_inv + leg_dir) / ms_per_step,
    }
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "of measured"
    except Exception:
        hbm, hbm_src = 6650.0, "of fallback"
    # Fourier stage: algorithmic bytes = exchange buffer (read or written once) + grid fields (written or read once)
    fb_bytes = sum(2 * 16.0 * nf * max(0, nleg - int(nlat0[m])) for m in range(T + 1))
    four_bytes = fb_bytes + 8.0 * npts * nf
    f_inv = float(np.mean([a for a, _ in four_ms]))
    f_dir = float(np.mean([b for _, b in four_ms]))
    roofline_fourier = {"kernel": "fourier2_inv_kernel / fourier2_dir_kernel (register-tiled chirp-z) + fourier_*_direct_kernel (mixed radix, rows "
                                  "without prime factors above 23) + v1 chirp-z kernels on short rows, all in shared memory", "bound": "hbm",
                        "paths": trans.fourier_paths(),
                        "achieved": 2 * four_bytes / ((f_inv + f_dir) * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                        "frac": 2 * four_bytes / ((f_inv + f_dir) * 1e-3) / 1e9 / hbm, "traffic": None, "peak_note": hbm_src,
                        "ms_per_launch_group": {"inverse": f_inv, "direct": f_dir}, "share_of_step": (f_inv + f_dir) / ms_per_step}

    # ---- end to end through the C ABI with pinned host buffers ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = bench_e2e(args, trans, torch, H, T, nf, npts, nspec, sp_host)

    cpu_baseline = None
    if not args.no_cpu_baseline:
        from oracle import pyoracle as po

        threads = host_threads()
        t0 = time.time()
        plan = po.OraclePlan(grid.nx(), grid.y(), T, weights=grid.weights(), nthreads=threads)
        osetup = time.time() - t0
        nfs = cpu_sample_fields(args, T, npts, nf)
        sps = sp_host.numpy() if nfs == nf else H.synthetic_spectra(T, nfs)
        t0 = time.time()
        gps = plan.invtrans(nfs, sps, mode=2)
        t_inv = time.time() - t0
        t0 = time.time()
        plan.dirtrans(nfs, gps)
        t_dir = time.time() - t0
        per_step = (t_inv + t_dir) * nf / nfs
        sample = (f"all {nf} fields, full grid, 1 inv + 1 dir, nothing extrapolated" if nfs == nf else
                  f"{nfs} of {nf} fields, full grid, 1 inv + 1 dir, time scaled x{nf}/{nfs}")
        cpu_baseline = {"value": 1.0 / per_step, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
                        "sample": f"{sample} (inv {t_inv:.2f}s, dir {t_dir:.2f}s; plan setup {osetup:.1f}s untimed)",
                        "note": CPU_NOTE}
        if nfs == nf:   # the same inputs went through both: parity of the timed GPU path against the oracle, all fields
            err = H.rel_max(d_gp.cpu().numpy(), gps)
            cpu_baseline["gpu_vs_oracle_invtrans_rel_max"] = err
            if not err < 1e-12:
                raise SystemExit(f"bench: device-resident invtrans differs from the CPU oracle (rel max {err:.3e})")
        del gps

    out = {
        "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if args.precision == "fp64" else "tf32x3 (fp32-level Legendre stage, fp64 Fourier stage)",
        "data": "synthetic",
        "config": {"workload": f"{args.workload} L{nf} invtrans+dirtrans " + ("fp64" if args.precision == "fp64" else "fp32 mixed-precision tensor-core Legendre") + f" (grid {gridname}, T{T})",
                   "l2": "inputs larger than L2 (spectra %.2f GB, grid fields %.2f GB per step)" % (8e-9 * nspec, 8e-9 * nf * npts),
                   "plan_setup_s": setup_s, "device_bytes": trans.device_bytes()},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "roofline_fourier": roofline_fourier, "cpu_baseline": cpu_baseline,
        "stage_ms": {"pack_inv": float(np.mean([a for a, _ in pack_ms])), "unpack_dir": float(np.mean([b for _, b in pack_ms])),
                     "legendre_inv": leg_inv, "legendre_dir": leg_dir, "fourier_inv": f_inv, "fourier_dir": f_dir},
    }
    emit_json(out)


if __name__ == "__main__":
    main()
