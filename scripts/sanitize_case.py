"""Small run of every default kernel for compute-sanitizer (racecheck / memcheck / synccheck)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi_parallel_multiscale_diffusion_fem_b200 as pkg
from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc

cases = [(6, 0, 2), (6, 6, 1), (5, 0, 2), (5, 3, 1), (4, 0, 2), (3, 0, 2), (7, 0, 1), (7, 4, 1), (6, 100, 1)]
if len(sys.argv) > 1:   # e.g. "5 4 6:6" = all l=5, l=4 cases and (l=6, variant 6)
    cases = [c for c in cases if str(c[0]) in sys.argv[1:] or "%d:%d" % (c[0], c[1]) in sys.argv[1:]]
for l, variant, cells in cases:
    with pkg.BasisShard(l, pkg.coarse_corners(4, 7, 7 + cells), coeff_desc(pkg.COEFF_REFERENCE), variant=variant) as sh:
        sh.run(1e-12, 40 if l < 7 else 6, allow_no_convergence=True)
        M, b = sh.element_matrices()
        it, res = sh.iteration_counts()
        print("l=%d variant=%d cells=%d iters=%s max|rowsum M|=%.2e" % (l, variant, cells, it[0].tolist(), np.abs(M.sum(axis=2)).max()), flush=True)
# dim = 3: (l, variant, cells); 300 cells switch the fused coarse+fine kernel on
cases3 = [(3, 0, 300), (3, 1, 2), (4, 0, 2), (2, 3, 2)]
if len(sys.argv) > 1:
    cases3 = [c for c in cases3 if "3d" in sys.argv[1:]]
for l, variant, cells in cases3:
    with pkg.BasisShard(l, pkg.coarse_corners3(3, 5, 5 + cells), coeff_desc(pkg.COEFF_REFERENCE), variant=variant,
                        dim=3) as sh:
        sh.run(1e-12, 6, allow_no_convergence=True)
        M, b = sh.element_matrices()
        it, res = sh.iteration_counts()
        print("3d l=%d variant=%d cells=%d iters=%s" % (l, variant, cells, it[0].tolist()), flush=True)
