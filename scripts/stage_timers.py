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

NAMES = ["prologue (scale + Galerkin)", "rhs init + z0", "stencil q=Ap", "reduce p.q",
         "r update + stage u (+barrier)", "restrict wide levels", "warp-0 coarse chain",
         "prolong wide levels", "fine prolong + reduce r.z,|r|", "p update (+barrier)",
         "epilogue (write phi)"]


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "target"
    cells = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
    variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    r, l, kind, par, seed = WORKLOADS[wl]
    lib = pkg.load_library()
    out = (C.c_ulonglong * 16)()
    with pkg.BasisShard(l, pkg.coarse_corners(r, 0, cells), coeff_desc(kind, par, seed), variant=variant) as sh:
        sh.run(1e-12, 5000)
        lib.msb_debug_stage_cycles(out, 1)
        sh.run(1e-12, 5000)
        lib.msb_debug_stage_cycles(out, 1)
        it, _ = sh.iteration_counts()
        st = sh.run_stats()
    cyc = np.array(out[:11], dtype=np.float64)
    tot = cyc.sum()
    n_iter = it.sum() / (4.0 / max(1, 4 // (4 if l <= 5 and variant == 0 else 1)))  # per solve-group iterations
    print("workload %s cells %d l=%d variant %d: solve kernel %.3f ms, mean k %.1f" %
          (wl, cells, l, variant, st["ms_solve"], it.mean()))
    print("cycles per CTA: %.0f  (per group-iteration: %.0f)" % (tot / cells, tot / it[:, 0].sum() if l <= 5 else tot / it.sum()))
    for nm, c in zip(NAMES, cyc):
        per_it = c / (it[:, 0].sum() if l <= 5 else it.sum())
        print("  %-34s %5.1f%%   %8.0f cycles/iteration" % (nm, 100 * c / tot, per_it))


if __name__ == "__main__":
    main()
