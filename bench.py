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


def run_reference(args, rank, world):
    """CPU arm: the oracle port of TransLocal (+ the quadrature-adjoint dirtrans) on all host threads."""
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
    nfs = min(nf, args.cpu_fields)
    sp = H.synthetic_spectra(T, nfs)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.time()
        gp = plan.invtrans(nfs, sp, mode=2)
        back = plan.dirtrans(nfs, gp)
        dt = time.time() - t0
        if it >= args.warmup:
            times.append(dt)
    per_step = float(np.mean(times)) * nf / nfs  # work is exactly linear in the number of fields
    value = 1.0 / per_step
    out = {
        "impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload} L{nf} invtrans+dirtrans fp64 (grid {gridname}, T{T})"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{nfs} of {nf} fields, full grid, inv+dir, time scaled x{nf}/{nfs}; plan setup {setup_s:.1f}s untimed"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_json(out)


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

        from atlas_b200 import dist as spdist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        out = spdist.bench_sharded(args, rank, world, local_rank, metric_name(args.workload), UNIT, FP64_DMMA_PEAK_TFLOPS)
        if out is not None:
            emit_json(out)
        return

    torch.cuda.set_device(0)
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
    roofline_fourier = {"kernel": "fourier2_inv_kernel / fourier2_dir_kernel (+ v1 kernels on short rows): chirp-z in shared memory", "bound": "hbm",
                        "achieved": 2 * four_bytes / ((f_inv + f_dir) * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                        "frac": 2 * four_bytes / ((f_inv + f_dir) * 1e-3) / 1e9 / hbm, "traffic": None, "peak_note": hbm_src,
                        "ms_per_launch_group": {"inverse": f_inv, "direct": f_dir}, "share_of_step": (f_inv + f_dir) / ms_per_step}

    # ---- end to end through the C ABI with pinned host buffers ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        gp_host = torch.empty(nf * npts, dtype=torch.float64).pin_memory()
        sp2_host = torch.empty(nspec, dtype=torch.float64).pin_memory()
        sp_np, gp_np, sp2_np = sp_host.numpy(), gp_host.numpy(), sp2_host.numpy()
        for _ in range(max(1, min(args.warmup, 2))):
            trans.invtrans(nf, sp_np, gp_np)
            trans.dirtrans(nf, gp_np, sp2_np)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            trans.invtrans(nf, sp_np, gp_np)   # H2D spectra, transform, D2H grid fields
            trans.dirtrans(nf, gp_np, sp2_np)  # H2D grid fields, transform, D2H spectra
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1) / args.steps
        e2e = {"value": 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": 8 * (nspec + nf * npts), "d2h_bytes_per_step": 8 * (nf * npts + nspec)}
        # sanity: round trip reproduces the spectra except the m == T column the scalar inverse drops
        err = float((sp2_host - sp_host).abs().max())
        e2e["roundtrip_max_abs_diff"] = err

    cpu_baseline = None
    if not args.no_cpu_baseline:
        from oracle import pyoracle as po

        N = int(gridname[1:])
        threads = host_threads()
        t0 = time.time()
        plan = po.OraclePlan(grid.nx(), grid.y(), T, weights=grid.weights(), nthreads=threads)
        osetup = time.time() - t0
        nfs = min(nf, args.cpu_fields)
        sps = H.synthetic_spectra(T, nfs)
        t0 = time.time()
        gps = plan.invtrans(nfs, sps, mode=2)
        t_inv = time.time() - t0
        t0 = time.time()
        plan.dirtrans(nfs, gps)
        t_dir = time.time() - t0
        per_step = (t_inv + t_dir) * nf / nfs
        cpu_baseline = {"value": 1.0 / per_step, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{nfs} of {nf} fields, full grid, 1 inv + 1 dir, time scaled x{nf}/{nfs} "
                                  f"(inv {t_inv:.2f}s, dir {t_dir:.2f}s; plan setup {osetup:.1f}s untimed)"}

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
