#!/usr/bin/env python
"""bench.py -- multiscale basis solves/sec of the basis stage (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            the sm_100a path
  python bench.py --impl reference --gpus N ...            the reference's CPU path

A step = one complete pass of the basis stage (stencil assembly + every PCG solve + the
coarse element matrices, bases written to HBM) over the workload's coarse cells.  N=1 runs
the configuration the metric is quoted on (256x256 coarse x 64x64 fine, periodic eps=1/64);
N>1 partitions the SAME problem over the ranks in contiguous Morton ranges exactly as the
reference distributes coarse cells over MPI ranks (strong scaling, no data-path collective;
the per-cell (M, b) are all-gathered over NCCL in the end-to-end leg only).

One JSON line on rank 0; see DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (coarse refinement r, fine refinement l, coeff kind, params, seed)
    "target": (8, 6, 1, (1.0 / 64, 0.9999), 0),       # 256x256 coarse x 64x64 fine (north star)
    "cfg2": (5, 5, 1, (1.0 / 64, 0.9999), 0),         # 32x32 x 32x32
    "cfg3": (7, 6, 1, (1.0 / 64, 0.9999), 0),         # 128x128 x 64x64
    "cfg4": (8, 5, 2, (2.0 ** -11, 0.2, 1e4, 1.0), 1234),
    "cfg5": (6, 8, 1, (1.0 / 64, 0.9999), 0),         # 64x64 x 256x256 (streamed tier)
    "cfg1": (3, 7, 0, (), 0),                         # the reference's default run
    "cfg1x64": (6, 7, 0, (), 0),                      # 64x64 coarse x 128x128 fine (cluster tier at scale)
    "target-refcoef": (8, 6, 0, (), 0),
    # dim = 3: (r, l, kind, par, seed, dim)
    "3d-16x16": (4, 4, 0, (), 0, 3),                  # 16^3 coarse x 16^3 fine hexes, 32768 solves
    "3d-8x32": (3, 5, 0, (), 0, 3),                   # 8^3 coarse x 32^3 fine hexes, 4096 solves
    "3d-32x8": (5, 3, 0, (), 0, 3),                   # 32^3 coarse x 8^3 fine hexes, 262144 solves
}


def workload(name):
    w = WORKLOADS[name]
    return w if len(w) == 6 else w + (2,)


METRIC = "multiscale basis solves/sec"
UNIT = "solves/s"


def describe(name, n_cells):
    r, l, kind, par, seed, dim = workload(name)
    kinds = ["reference MatrixCoeff (k=57, PI_D typo)", "periodic eps=1/64", "random inclusions 1e4",
             "constant", "table"]
    return "%s: %s coarse x %s fine Q1, %s, f=2, tol 1e-12 abs, %d cells = %d solves" % (
        name, "x".join([str(1 << r)] * dim), "x".join([str(1 << l)] * dim), kinds[kind], n_cells,
        (1 << dim) * n_cells)


# ------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if bits & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_baseline(name, cells_per_core=None, budget_s=12.0):
    """The reference's CPU path for this stage, restated (oracle/: CSR + SSOR(1.6)-PCG, the
    reference's own algorithm; deal.II is not available so the reference binary cannot be
    built).  All host cores, contiguous Morton ranges per core, bounded sample."""
    from oracle import oracle as O
    O.build()
    r, l, kind, par, seed, dim = workload(name)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    total = (1 << r) ** dim
    c = O.coeff(kind, par, seed)
    run_cells = O.run_cells if dim == 2 else O.run_cells3
    corners_of = O.coarse_corners if dim == 2 else O.coarse_corners3
    # calibrate on 1 cell per core, then size the sample to ~budget_s of wall time
    probe = min(total, cores)
    t0 = time.perf_counter()
    run_cells(l, corners_of(r, list(range(probe))), c, n_threads=cores, keep_phi=False)
    per_cell = max(1e-4, (time.perf_counter() - t0))
    if cells_per_core is None:
        cells_per_core = int(max(1, min(256, budget_s / per_cell)))
    ncell = min(total, cores * cells_per_core)
    # spread the sample over the Morton curve so that it sees the whole coefficient range
    stride = max(1, total // ncell)
    cells = list(range(0, stride * ncell, stride))[:ncell]
    cor = corners_of(r, cells)
    t0 = time.perf_counter()
    res = run_cells(l, cor, c, n_threads=cores, keep_phi=False)
    dt = time.perf_counter() - t0
    assert res["failed"] == 0
    return {"value": (1 << dim) * ncell / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of %d coarse cells (every %d-th along the Morton curve), %d per core, "
                      "%.2f s wall, mean SSOR-PCG iterations %.1f"
                      % (ncell, total, stride, cells_per_core, dt, float(res["iters"].mean())),
            "seconds": dt, "solves": (1 << dim) * ncell}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(args.workload, budget_s=6.0)
        if i >= args.warmup:
            vals.append(last)
    tot_s = sum(v["seconds"] for v in vals)
    tot_n = sum(v["solves"] for v in vals)
    value = tot_n / tot_s
    r, l, kind, par, seed, dim = workload(args.workload)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / len(vals),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": describe(args.workload, (1 << r) ** dim),
                   "note": "reference binary unbuildable here (needs deal.II/MPI/Trilinos): timed the "
                           "oracle port of its SSOR-PCG local solver on all host cores; each step is a "
                           "bounded sample of the workload's coarse cells"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": "port",
                         "sample": last["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--cells", type=int, default=0, help="limit the number of coarse cells (debug)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--max-iter", type=int, default=5000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 3) if not args.cells else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import mpi_parallel_multiscale_diffusion_fem_b200 as pkg
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc

    # stdout carries exactly ONE JSON line: anything a library prints to file descriptor 1 from here on
    # (NCCL's "NCCL version ..." banner when the environment sets NCCL_DEBUG, ...) is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the basis stage has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    r, l, kind, par, seed, dim = workload(args.workload)
    nb = 1 << dim
    total_cells = (1 << r) ** dim
    if args.cells:
        total_cells = min(total_cells, args.cells)
    lo, hi = pkg.morton_partition(total_cells, rank, world)
    n_local = hi - lo
    N = ((1 << l) + 1) ** dim

    # host inputs/outputs in pinned memory (the e2e leg copies them every step)
    corners_np = pkg.coarse_corners(r, lo, hi) if dim == 2 else pkg.coarse_corners3(r, lo, hi)
    h_corners = torch.from_numpy(corners_np).pin_memory()
    h_M = torch.empty((n_local, nb, nb), dtype=torch.float64).pin_memory()
    h_b = torch.empty((n_local, nb), dtype=torch.float64).pin_memory()
    h_it = torch.empty((n_local, nb), dtype=torch.int32).pin_memory()

    sh = pkg.BasisShard(l, corners_np, coeff_desc(kind, par, seed), device_id=local_rank,
                        variant=args.variant, dim=dim)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        sh.run_async(1e-12, args.max_iter, sptr)

    # working sets that could survive in the 126 MB L2 from one step to the next (the reference's
    # default run: 64 cells) are flushed out between timed steps by writing a 512 MB buffer
    ws_bytes = n_local * (10 if dim == 2 else 23) * N * 8
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda") if ws_bytes < (512 << 20) else None

    # ---- kernel-resident leg: inputs already in HBM -----------------------------------
    for _ in range(args.warmup):
        step()
    sh.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    solve_ms, launches = [], 0
    ms_total = 0.0
    if flush_buf is None:
        ev0.record(stream)
    for _ in range(args.steps):
        if flush_buf is not None:
            flush_buf.zero_()            # untimed
            torch.cuda.synchronize()
        step()
        # per-step device-side stats need the events of this run: sync this rank's stream
        sh.sync()
        st = sh.run_stats()
        solve_ms.append(st["ms_solve"])
        launches += st["launches"]
        if flush_buf is not None:
            # the stage's own CUDA events, recorded on the stream the kernels run on
            # (msb_get_run_stats: first launch of the assembly -> end of the element matrices)
            ms_total += st["ms_total"]
    if flush_buf is None:
        ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    if flush_buf is None:
        ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    it_np, res_np = sh.iteration_counts()
    alg_bytes, mean_k = sh.algorithmic_bytes()
    stats = torch.tensor([alg_bytes, float(it_np.sum()), float(launches), float(np.mean(solve_ms))],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        smax = stats.clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        dist.all_reduce(smax, op=dist.ReduceOp.MAX)
        solve_ms_max = float(smax[3].item())
    else:
        solve_ms_max = float(stats[3].item())
    alg_bytes_all, iters_all, launches_all = float(stats[0]), float(stats[1]), int(stats[2])
    n_solves = nb * total_cells
    value = n_solves / (ms_step * 1e-3)

    # ---- end-to-end leg: host buffers in, host buffers out, through the C ABI -------------
    e2e = None
    if not args.no_e2e:
        from mpi_parallel_multiscale_diffusion_fem_b200 import parallel

        def e2e_step():
            sh.set_cells_ptr(h_corners.data_ptr())                 # H2D corners; BasisQ1 data on device
            sh.run_async(1e-12, args.max_iter, sptr)
            sh.sync()
            sh.element_matrices_into(h_M.data_ptr(), h_b.data_ptr())   # D2H
            sh.iteration_counts_into(h_it.data_ptr())
            if world > 1:
                # the reference's compress(add) exchange (ms.tpp:253-254): every rank obtains the
                # per-cell coarse contributions of all ranks over NVLink (NCCL all_gather)
                parallel.gather_coarse_contributions(h_M, h_b, total_cells, device="cuda")

        e2e_step()
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        t0.record(stream)
        for _ in range(args.steps):
            e2e_step()                   # (no L2 flush here: every step starts from host buffers)
        t1.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - w0) * 1e3
        # the D2H copies run on the library's stream: take the larger of device and wall time
        tt = torch.tensor([max(t0.elapsed_time(t1), wall_ms)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item()) / args.steps
        e2e = {"value": n_solves / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h_corners.numel() * 8),
               "d2h_bytes_per_step": int(h_M.numel() * 8 + h_b.numel() * 8 + h_it.numel() * 4),
               "ms_per_step": e2e_ms,
               "api": "msb_set_cells + msb_run_async + msb_sync + msb_get_element_matrices + "
                      "msb_get_iteration_counts (pinned host buffers; bases stay device-resident "
                      "as in the reference, where they live inside the basis objects)"}

    # ---- checks that make the number meaningful -----------------------------------------
    M, b = sh.element_matrices()
    H = 1.0 / (1 << r)
    ok = bool(np.all(res_np <= 1e-12) and np.abs(M.sum(axis=2)).max() < 1e-8 * np.abs(M).max()
              and np.abs(b.sum(axis=1) - 2 * H ** dim).max() < 1e-9 * H ** dim)
    if not ok:
        raise SystemExit("bench.py: results failed the invariants (zero row sums / load / residual)")

    if rank == 0:
        peaks = {}
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        if os.path.exists(ppath):
            peaks = json.load(open(ppath))
            peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        # dominant kernel: the PCG solve kernel; algorithmic bytes N(96k+16) per solve
        #  of the solves ONE launch processes / its CUDA-event duration
        achieved = (alg_bytes_all / world) / (solve_ms_max * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get(args.workload, {}).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": describe(args.workload, total_cells),
                       "partition": "contiguous Morton ranges over %d rank(s) (p4est rule)" % world,
                       "l2": ("working set (stencil + bases = %.1f GB per GPU) far larger than L2; "
                              "no flush needed" % (ws_bytes / 1e9)) if flush_buf is None else
                             ("working set %.0f MB per GPU: L2 flushed between timed steps (untimed "
                              "write of a 512 MB buffer), each step timed by the stage's own CUDA events on its stream"
                              % (ws_bytes / 1e6)),
                       "mean_pcg_iterations": iters_all / n_solves,
                       "preconditioner": ("multilevel diagonal scaling (BPX), exact Galerkin diagonals"
                                          + (", exact solve of the 7x7 coarse level" if (dim == 2 and l == 5) or (dim == 2 and l == 6 and args.variant == 0) else "")
                                          if args.variant < 100 else "Jacobi (symmetric diagonal scaling)"),
                       "variant": args.variant},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches_all,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": (("solve_bpx_tm_kernel" if (l == 6 and args.variant in (0, 5, 7)) else "solve_bpx_kernel") if args.variant < 100 else "solve_smem_kernel") if sh.run_stats()["tier"] == 1 else (("solve_cluster_kernel" if (l == 7 and args.variant == 0) or (args.variant in (3, 4) and 5 <= l <= 7) else "stream_k*") if dim == 2 else "d3::k2_kernel + siblings"),
                         "kernel_ms_per_launch": solve_ms_max,
                         "peak_source": peak_src,
                         "note": "achieved = ALGORITHMIC streaming bytes N(96k+16)/solve (SURVEY 8d) / "
                                 "solve-kernel time; vectors live on chip, so real DRAM traffic "
                                 "(`traffic`, ncu) is far smaller and frac may exceed 1"},
        }
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(args.workload)
            cb.pop("seconds"), cb.pop("solves")
            line["cpu_baseline"] = cb
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    sh.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
