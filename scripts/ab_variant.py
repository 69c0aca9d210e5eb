"""A/B of a kernel variant against the default on a slice of a workload: iterations, bases, timing.
usage: python scripts/ab_variant.py <workload> <variant> [cells]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi_parallel_multiscale_diffusion_fem_b200 as pkg  # noqa: E402
from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc  # noqa: E402
from bench import workload  # noqa: E402

wl, variant = sys.argv[1], int(sys.argv[2])
cells = int(sys.argv[3]) if len(sys.argv) > 3 else 1184
r, l, kind, par, seed, dim = workload(wl)
cor = pkg.coarse_corners(r, 0, cells) if dim == 2 else pkg.coarse_corners3(r, 0, cells)
out = {}
for v in (0, variant):
    with pkg.BasisShard(l, cor, coeff_desc(kind, par, seed), variant=v, dim=dim) as sh:
        for _ in range(3):
            sh.run(1e-12, 5000)
        it, res = sh.iteration_counts()
        out[v] = (it, sh.element_matrices()[0], [sh.basis(c, 1) for c in (0, cells - 1)], sh.run_stats())
        print("variant %d: solve %.3f ms total %.3f ms mean k %.2f max res %.2e" % (
            v, out[v][3]["ms_solve"], out[v][3]["ms_total"], it.mean(), res.max()))
a, b = out[0], out[variant]
rel = lambda x, y: np.linalg.norm(x - y) / np.linalg.norm(y)
print("M rel diff %.2e, basis rel diff %.2e %.2e" % (rel(b[1], a[1]), rel(b[2][0], a[2][0]), rel(b[2][1], a[2][1])))
