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
PER_CONFIG = ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5", "3d-16x16"]   # BASELINE.json configs 1-5 + the 3D path


def kernel_name(dim, l, variant, tier):
    if dim == 3:
        return "d3::k2m_kernel + siblings (HBM-streamed)"
    if tier == 1:
        if variant >= 100:
            return "solve_smem_kernel"
        if l in (5, 6) and (variant == 0 or 10 <= variant <= 12):
            return "solve_fused_kernel (assembly + 4 solves + element matrices in one launch)"
        return "solve_bpx_tm_kernel" if (l == 6 and variant in (5, 7, 9)) else "solve_bpx_kernel"
    if (l == 7 and variant in (0, 8)) or (variant in (3, 4) and 5 <= l <= 7):
        return "solve_cluster_kernel"
    return "stream_k* (HBM-streamed)"


def binding_resource(dim, l, variant, tier):
    """What actually bounds the dominant kernel (ncu evidence in profiles/): the HBM model of SURVEY 8(d)
    only BINDS the tiers whose vectors stream through HBM."""
    name = kernel_name(dim, l, variant, tier)
    if "HBM-streamed" in name:
        return "hbm"
    return "shared-memory crossbar (LSU wavefronts) + dependency latency; HBM is not binding (vectors stay on chip)"


class Ctx:
    pass


def measure(cx, name, steps, warmup, variant=0, cells=0, max_iter=5000, e2e=True, with_bases=False, tier=0):
    """One workload on this rank's Morton range: kernel-resident leg, optional end-to-end legs, invariants.
    Returns a dict (identical on every rank for the reduced quantities)."""
    torch, dist, pkg = cx.torch, cx.dist, cx.pkg
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc
    world, rank, local_rank = cx.world, cx.rank, cx.local_rank
    r, l, kind, par, seed, dim = workload(name)
    nb = 1 << dim
    total_cells = (1 << r) ** dim
    if cells:
        total_cells = min(total_cells, cells)
    lo, hi = pkg.morton_partition(total_cells, rank, world)
    n_local = hi - lo
    N = ((1 << l) + 1) ** dim
    dev = torch.device("cuda", local_rank)

    corners_np = pkg.coarse_corners(r, lo, hi) if dim == 2 else pkg.coarse_corners3(r, lo, hi)
    h_corners = torch.from_numpy(corners_np).pin_memory()
    h_M = torch.empty((total_cells, nb, nb), dtype=torch.float64, pin_memory=True)
    h_b = torch.empty((total_cells, nb), dtype=torch.float64, pin_memory=True)
    h_it = torch.empty((n_local, nb), dtype=torch.int32, pin_memory=True)

    sh = pkg.BasisShard(l, corners_np, coeff_desc(kind, par, seed), device_id=local_rank, variant=variant, dim=dim,
                        tier=tier)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        sh.run_async(1e-12, max_iter, sptr)

    # working sets that could survive in the 126 MB L2 from one step to the next (the reference's
    # default run: 64 cells) are flushed out between timed steps by writing a 512 MB buffer
    ws_bytes = n_local * (10 if dim == 2 else 23) * N * 8
    flush = ws_bytes < (512 << 20)
    if flush and cx.flush_buf is None:
        cx.flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    # ---- kernel-resident leg: inputs already in HBM -----------------------------------
    for _ in range(warmup):
        step()
    sh.sync()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    solve_ms, launches = [], 0
    ms_total = 0.0
    if not flush:
        ev0.record(stream)
    for _ in range(steps):
        if flush:
            cx.flush_buf.zero_()            # untimed
            torch.cuda.synchronize()
        step()
        # per-step device-side stats need the events of this run: sync this rank's stream
        sh.sync()
        st = sh.run_stats()
        solve_ms.append(st["ms_solve"])
        launches += st["launches"]
        if flush:
            # the stage's own CUDA events, recorded on the stream the kernels run on
            # (msb_get_run_stats: first launch of the assembly -> end of the element matrices)
            ms_total += st["ms_total"]
    if not flush:
        ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    if not flush:
        ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / steps
    it_np, res_np = sh.iteration_counts()
    alg_bytes, mean_k = sh.algorithmic_bytes()
    stats = torch.tensor([alg_bytes, float(it_np.sum()), float(launches), float(np.mean(solve_ms))],
                         dtype=torch.float64, device=dev)
    if world > 1:
        smax = stats.clone()
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        dist.all_reduce(smax, op=dist.ReduceOp.MAX)
        solve_ms_max = float(smax[3].item())
    else:
        solve_ms_max = float(stats[3].item())
    n_solves = nb * total_cells
    out = {
        "name": name, "dim": dim, "r": r, "l": l, "N": N, "total_cells": total_cells, "n_local": n_local,
        "n_solves": n_solves, "ms_step": ms_step, "value": n_solves / (ms_step * 1e-3),
        "alg_bytes_all": float(stats[0]), "iters_all": float(stats[1]), "launches_all": int(stats[2]),
        "solve_ms_max": solve_ms_max, "clocks": clocks, "tier": sh.run_stats()["tier"], "ws_bytes": ws_bytes,
        "flush": flush, "steps": steps, "warmup": warmup, "e2e": None, "e2e_with_bases": None,
    }

    # ---- end-to-end leg: host buffers in, host buffers out, through the C ABI -------------
    if e2e:
        from mpi_parallel_multiscale_diffusion_fem_b200 import parallel
        d_M = d_b = None
        if world > 1:
            d_M = torch.empty((total_cells, nb, nb), dtype=torch.float64, device=dev)
            d_b = torch.empty((total_cells, nb), dtype=torch.float64, device=dev)
        h_phi = None
        if with_bases:
            h_phi = torch.empty((n_local, nb, N), dtype=torch.float64, pin_memory=True)

        def e2e_step(bases):
            sh.set_cells_ptr(h_corners.data_ptr())                 # H2D corners; BasisQ1 data on device
            if bases:
                # the stage with the 2^dim solution_vectors of every cell brought to the host, where the reference
                # keeps them (basis.hpp:216): msb_run_with_bases pipelines the reordering + D2H of chunk k behind
                # the solves of chunk k+1
                sh.run_with_bases(1e-12, max_iter, out_addr=h_phi.data_ptr())
            else:
                sh.run_async(1e-12, max_iter, sptr)
                sh.sync()
            if world > 1:
                # the reference's compress(add) exchange (ms.tpp:253-254): every rank obtains the per-cell
                # coarse contributions of all ranks over NVLink -- NCCL all_gather straight out of the
                # library's device buffers (msb_get_device_results), then ONE device->host copy
                parallel.gather_coarse_contributions_device(sh, total_cells, dev, d_M, d_b)
                h_M.copy_(d_M, non_blocking=True)
                h_b.copy_(d_b, non_blocking=True)
                sh.iteration_counts_into(h_it.data_ptr())
                torch.cuda.current_stream().synchronize()
            else:
                sh.element_matrices_into(h_M.data_ptr(), h_b.data_ptr())   # D2H
                sh.iteration_counts_into(h_it.data_ptr())

        def timed(bases, k):
            e2e_step(bases)
            barrier()
            t0 = torch.cuda.Event(enable_timing=True)
            t1 = torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            t0.record(stream)
            for _ in range(k):
                e2e_step(bases)          # (no L2 flush here: every step starts from host buffers)
            t1.record(stream)
            barrier()
            wall_ms = (time.perf_counter() - w0) * 1e3
            # the copies run on the library's streams: take the larger of device and wall time
            tt = torch.tensor([max(t0.elapsed_time(t1), wall_ms)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item()) / k

        e2e_ms = timed(False, steps)
        d2h = int((h_M.numel() + h_b.numel()) * 8 + h_it.numel() * 4)
        out["e2e"] = {
            "value": n_solves / (e2e_ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": int(h_corners.numel() * 8), "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
            "api": "msb_set_cells + msb_run_async + msb_sync + msb_get_element_matrices + "
                   "msb_get_iteration_counts (pinned host buffers; bases stay device-resident); at N>1 the "
                   "NCCL all_gather reads msb_get_device_results and the gathered (M, b) go to the host once"}
        if with_bases:
            kb = max(1, min(steps, 3))
            wb_ms = timed(True, kb)
            out["e2e_with_bases"] = {
                "value": n_solves / (wb_ms * 1e-3), "unit": UNIT, "ms_per_step": wb_ms, "steps": kb,
                "h2d_bytes_per_step": int(h_corners.numel() * 8),
                "d2h_bytes_per_step": d2h + int(h_phi.numel() * 8),
                "api": "msb_set_cells + msb_run_with_bases (the stage with the bases of every local cell delivered into "
                       "pinned host memory in deal.II DoF order, reordering + D2H of chunk k pipelined behind the "
                       "solves of chunk k+1) + msb_get_element_matrices + msb_get_iteration_counts: what a caller "
                       "pays who keeps the bases host-side like the reference (basis.hpp:216, ms.tpp:386-393)"}
            # spot check of the bulk copy: partition of unity on the last cell
            pu = float((h_phi[-1].sum(dim=0) - 1.0).abs().max())
            if not pu < 1e-9:
                raise SystemExit("bench.py: bulk bases failed the partition-of-unity check (%g)" % pu)
            del h_phi

    # ---- checks that make the number meaningful -----------------------------------------
    M, b = sh.element_matrices()
    H = 1.0 / (1 << r)
    ok = bool(np.all(res_np <= 1e-12) and np.abs(M.sum(axis=2)).max() < 1e-8 * np.abs(M).max()
              and np.abs(b.sum(axis=1) - 2 * H ** dim).max() < 1e-9 * H ** dim)
    if not ok:
        raise SystemExit("bench.py: %s failed the invariants (zero row sums / load / residual)" % name)
    sh.close()
    del sh
    return out


def load_profile_facts(name, build_id):
    """profiles/traffic.json: per-cell DRAM bytes and pipe utilisations of the dominant kernel from the committed
    ncu --set full page -- quoted ONLY when it was captured on the binary that is running (msb_build_id)."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, "profiles/traffic.json missing"
    tj = json.load(open(tpath))
    ent = tj.get("workloads", {}).get(name)
    if ent is None:
        return None, "no ncu capture of this workload"
    if tj.get("build_id") != build_id:
        return None, ("ncu capture is of build %s, running build %s: not quoted" % (tj.get("build_id"), build_id))
    return ent, None


def roofline_of(m, variant, build_id, peaks):
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if "hbm_gbs" in peaks:
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    world_share = m["alg_bytes_all"] * m["n_local"] / max(1, m["total_cells"])   # rank 0's launch
    achieved = world_share / (m["solve_ms_max"] * 1e-3) / 1e9
    ent, why = load_profile_facts(m["name"], build_id)
    bind = binding_resource(m["dim"], m["l"], variant, m["tier"])
    rf = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
          "traffic": None if ent is None else ent["dram_bytes_per_cell"] * m["n_local"],
          "kernel": kernel_name(m["dim"], m["l"], variant, m["tier"]),
          "kernel_ms_per_launch": m["solve_ms_max"], "peak_source": peak_src,
          "binding_resource": bind,
          "note": "achieved = ALGORITHMIC streaming bytes N(96k+16)/solve (SURVEY 8d) of the solves ONE launch of "
                  "this rank processes / solve-kernel time: a MODEL throughput.  For the on-chip tiers the vectors "
                  "never leave the SM, real DRAM traffic (`traffic`, ncu, per launch of this rank) is ~1 % of HBM "
                  "peak and frac > 1 says nothing about kernel quality -- `secondary` (ncu pipe utilisations "
                  "against the measured on-chip peaks) does"}
    if ent is None:
        rf["traffic_note"] = why
    else:
        rf["traffic_source"] = ent.get("source")
        rf["secondary"] = {k: ent[k] for k in ("smem_wavefront_frac", "fp64_pipe_frac", "issue_slot_frac",
                                                "bank_conflict_wavefront_frac") if k in ent}
    ppath = os.path.join(ROOT, "profiles", "onchip_peaks.json")
    if os.path.exists(ppath):
        op = json.load(open(ppath))
        rf["onchip_peaks"] = {"fp64_tflops": op.get("fp64", {}).get("tflops"),
                              "smem_bytes_per_clk_per_sm": op.get("smem", {}).get("lds128", {}).get("bytes_per_clk_per_sm"),
                              "source": "scripts/probes/onchip_peaks.cu on this pool's B200 (profiles/onchip_peaks.json)"}
    return rf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="target", choices=sorted(WORKLOADS))
    ap.add_argument("--cells", type=int, default=0, help="limit the number of coarse cells (debug)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--tier", type=int, default=0, help="msb_tier: 0 auto, 1 shared memory, 2 streamed / cluster")
    ap.add_argument("--max-iter", type=int, default=5000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-per-config", action="store_true",
                    help="skip the per_config block (BASELINE configs 1-5 + 3D after the headline leg)")
    ap.add_argument("--no-bases", action="store_true", help="skip the e2e_with_bases leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = max(args.warmup, 3) if not args.cells else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import mpi_parallel_multiscale_diffusion_fem_b200 as pkg
    from mpi_parallel_multiscale_diffusion_fem_b200 import binding

    # stdout carries exactly ONE JSON line: anything a library prints to file descriptor 1 from here on
    # (NCCL's "NCCL version ..." banner when the environment sets NCCL_DEBUG, ...) is sent to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    cx = Ctx()
    cx.torch, cx.dist, cx.pkg = torch, dist, pkg
    cx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = rank = int(os.environ.get("RANK", "0"))
    cx.local_rank = local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cx.flush_buf = None
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the basis stage has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    build_id = binding.build_id()

    m = measure(cx, args.workload, args.steps, args.warmup, variant=args.variant, cells=args.cells,
                max_iter=args.max_iter, e2e=not args.no_e2e,
                with_bases=not (args.no_e2e or args.no_bases), tier=args.tier)

    # ---- the other BASELINE configurations, each at the rank count of this run -------------
    per_config = None
    if not args.no_per_config and not args.cells and args.workload == "target" and args.variant == 0:
        per_config = {}
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) \
            if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        for name in PER_CONFIG:
            k = 2 if name == "cfg5" else 3
            pm = measure(cx, name, k, 3, e2e=False)
            rf = roofline_of(pm, 0, build_id, peaks)
            per_config[name] = {
                "workload": describe(name, pm["total_cells"]), "value": pm["value"], "unit": UNIT,
                "ms_per_step": pm["ms_step"], "steps": k, "warmup": 3,
                "mean_pcg_iterations": pm["iters_all"] / pm["n_solves"],
                "model_frac": rf["frac"], "model_gbs": rf["achieved"], "kernel": rf["kernel"],
                "binding_resource": rf["binding_resource"], "kernel_ms_per_launch": pm["solve_ms_max"],
                "gpu_launches": pm["launches_all"], "clocks": pm["clocks"],
                "l2": "flushed between steps" if pm["flush"] else "working set >> L2"}

    if rank == 0:
        r, l, kind, par, seed, dim = workload(args.workload)
        peaks = {}
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(ppath):
            peaks = json.load(open(ppath))
        rf = roofline_of(m, args.variant, build_id, peaks)
        n_solves = m["n_solves"]
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["ms_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": describe(args.workload, m["total_cells"]),
                       "partition": "contiguous Morton ranges over %d rank(s) (p4est rule)" % world,
                       "l2": ("working set (stencil + bases = %.1f GB per GPU) far larger than L2; "
                              "no flush needed" % (m["ws_bytes"] / 1e9)) if not m["flush"] else
                             ("working set %.0f MB per GPU: L2 flushed between timed steps (untimed "
                              "write of a 512 MB buffer), each step timed by the stage's own CUDA events on its stream"
                              % (m["ws_bytes"] / 1e6)),
                       "mean_pcg_iterations": m["iters_all"] / n_solves,
                       "preconditioner": ("multilevel diagonal scaling (BPX), exact Galerkin diagonals (the reciprocal "
                                          "diagonals of the coarse levels are STORED as float; all arithmetic f64)"
                                          + (", exact solve of the 7x7 coarse level" if (dim == 2 and l == 5) or (dim == 2 and l == 6 and args.variant in (0, 9)) else "")
                                          if args.variant < 100 else "Jacobi (symmetric diagonal scaling)"),
                       "initial_guess": ("coarse Q1 shape function (fused stage)" if "solve_fused" in kernel_name(dim, l, args.variant, m["tier"])
                                         else "zero"),
                       "variant": args.variant, "build_id": build_id},
            "clocks": m["clocks"],
            "e2e": m["e2e"],
            "e2e_with_bases": m["e2e_with_bases"],
            "gpu_launches": m["launches_all"],
            "roofline": rf,
        }
        if per_config is not None:
            line["per_config"] = per_config
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(args.workload)
            cb.pop("seconds"), cb.pop("solves")
            line["cpu_baseline"] = cb
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
