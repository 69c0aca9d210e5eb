"""Per-stage cycle split of solve_bpx_kernel (profiling build, MSB_STAGE_TIMERS).
usage: MSB_LIBRARY=.../libmsfem_basis_prof.so python scripts/stage_timers.py [workload] [cells]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MSB_LIBRARY", os.path.join(
    ROOT, "mpi_parallel_multiscale_diffusion_fem_b200", "libmsfem_basis_prof.so"))
import mpi_parallel_multiscale_diffusion_fem_b200 as pkg  # noqa: E402
from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc  # noqa: E402
from bench import WORKLOADS  # noqa: E402

# (fused kernel, l = 6 variant 0: stage 0 also holds the on-chip stencil assembly, stage 1 is unused, stage 9 is the
#  fused fine prolongation + direction update, stage 10 the epilogue with the element-matrix sums)
NAMES = {0: "prologue (scale + Galerkin)", 1: "rhs init + z0", 2: "stencil q=Ap", 3: "reduce p.q",
         4: "r update + stage u (+barrier)", 5: "restrict 0->1->2 (wide levels)",
         6: "direct restrict to 7x7/3x3/1x1", 7: "direct interpolate to 15x15", 11: "prolong 2->1 (wide)",
         8: "fine prolong + reduce r.z,|r|", 9: "p update (+barrier)", 10: "epilogue (write phi)"}
ORDER = [0, 1, 2, 3, 4, 5, 6, 7, 11, 8, 9, 10]


CL_NAMES = {0: "prologue (coefficients to SMEM)", 1: "rhs init + r.r", 2: "restrict level 1 + barrier",
            3: "restrict level 2 + barrier", 4: "restrict level 3, gather + barrier",
            5: "levels >= 3 down / up (redundant)", 6: "levels 2, 1 up", 7: "fine z + r.z all-reduce",
            8: "p update (slab + halo rows)", 9: "stencil q = K p + p.q all-reduce",
            10: "x, r update, stage r + r.r all-reduce", 11: "epilogue (write phi) + barrier"}


def cluster_main(wl, cells, variant):
    """solve_cluster_kernel: cycles of thread 0 of every CTA, summed over the CTAs of all clusters."""
    r, l, kind, par, seed = WORKLOADS[wl]
    lib = pkg.load_library()
    out = (C.c_ulonglong * 16)()
    with pkg.BasisShard(l, pkg.coarse_corners(r, 0, cells), coeff_desc(kind, par, seed), variant=variant,
                        tier=pkg.TIER_STREAMED) as sh:
        sh.run(1e-12, 5000)
        lib.msb_debug_stage_cycles_cl(out, 1)
        sh.run(1e-12, 5000)
        lib.msb_debug_stage_cycles_cl(out, 1)
        it, _ = sh.iteration_counts()
        st = sh.run_stats()
    cyc = np.array(out[:12], dtype=np.float64)
    ctas = cells * ((1 << l) // 16)
    # a pass iterates until all its bases have converged (variant 4: two passes of two bases)
    pass_its = sum((max(row[0], row[1]) + max(row[2], row[3])) if variant == 4 else max(row) for row in it)
    print("workload %s cells %d l=%d: solve %.3f ms, mean k %.1f, cycles per CTA %.0f" %
          (wl, cells, l, st["ms_solve"], it.mean(), cyc.sum() / ctas))
    for idx in range(12):
        print("  %-42s %5.1f%%   %8.0f cycles/pass-iteration" %
              (CL_NAMES[idx], 100 * cyc[idx] / cyc.sum(), cyc[idx] / ctas / (pass_its / cells)))


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "target"
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
    variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    r, l, kind, par, seed = WORKLOADS[wl]
    if (l == 7 and variant == 0) or variant in (3, 4):
        return cluster_main(wl, cells, variant)
    lib = pkg.load_library()
    out = (C.c_ulonglong * 16)()
    with pkg.BasisShard(l, pkg.coarse_corners(r, 0, cells), coeff_desc(kind, par, seed), variant=variant) as sh:
        sh.run(1e-12, 5000)
        fn = lib.msb_debug_stage_cycles_fu if (l == 6 and variant == 0) else (
            lib.msb_debug_stage_cycles_tm if (l == 6 and variant in (5, 7, 9)) else lib.msb_debug_stage_cycles)
        fn(out, 1)
        sh.run(1e-12, 5000)
        fn(out, 1)
        it, _ = sh.iteration_counts()
        st = sh.run_stats()
    fused = l == 6 and variant == 0
    cyc = np.array(out[:16 if fused else 12], dtype=np.float64)
    tot = cyc.sum()
    n_iter = it.sum() / (4.0 / max(1, 4 // (4 if l <= 5 and variant == 0 else 1)))  # per solve-group iterations
    print("workload %s cells %d l=%d variant %d: solve kernel %.3f ms, mean k %.1f" %
          (wl, cells, l, variant, st["ms_solve"], it.mean()))
    nrhs = 2 if (l == 6 and variant in (0, 5, 7, 9)) or (l == 5 and variant == 0) else (4 if l <= 5 and variant == 3 else 1)
    group_its = it.sum() / nrhs
    print("cycles per CTA: %.0f  (per pass-iteration of %d bases: %.0f)" % (tot / cells, nrhs, tot / group_its))
    names = dict(NAMES)
    order = list(ORDER)
    if fused:
        names.update({12: "prologue: sine tables + node stencils", 13: "prologue: Galerkin level 1, sqrt(d), scaling",
                      14: "prologue: Galerkin levels 2..5", 0: "prologue: exact 7x7 inverse -> TMEM",
                      1: "Dirichlet table + rhs init (per pass)", 8: "reduce r.z, |r| (+ level-1 barrier)",
                      9: "fine prolong + x, p update (+barrier)", 10: "epilogue (phi, M, b)"})
        names.update({1: "pass init: x0 = g, halo, coarse vectors (once per pass)",
                      6: "7x7 level: restriction, exact solve, block barrier",
                      7: "15x15 level: scaling + interpolation from 7x7"})
        order = [12, 13, 14, 0, 1, 2, 3, 4, 5, 6, 7, 11, 8, 9, 10]
    for idx in order:
        nm, c = names[idx], cyc[idx]
        per_it = c / group_its
        print("  %-52s %5.1f%%   %8.0f cycles/iteration" % (nm, 100 * c / tot, per_it))


if __name__ == "__main__":
    main()
