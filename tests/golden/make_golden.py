"""Generates tests/golden/oracle_golden.json from the CPU oracle.

The reference holds no numerical fixture (test/test_dummy.cc only prints the MPI rank
count) and cannot be built here (deal.II absent), so these vectors pin the ORACLE against
itself over time and give the GPU tests a committed target; the only numbers that come
from outside the oracle are the survey-time cross-check values in survey_crosscheck.json
(SURVEY.md Appendix B).  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402


def morton(ix, iy, r):
    m = 0
    for bit in range(r):
        m |= ((ix >> bit) & 1) << (2 * bit) | ((iy >> bit) & 1) << (2 * bit + 1)
    return m


CASES = [
    # name, r, l, (ix, iy), coeff kind, params, seed
    ("default_cell00", 3, 7, (0, 0), O.COEFF_REFERENCE, (), 0),
    ("default_cell35", 3, 7, (3, 5), O.COEFF_REFERENCE, (), 0),
    ("cfg2_cell_5_7", 5, 5, (5, 7), O.COEFF_PERIODIC, (1.0 / 64, 0.9999), 0),
    ("cfg3_cell_100_17", 7, 6, (100, 17), O.COEFF_PERIODIC, (1.0 / 64, 0.9999), 0),
    ("cfg4_cell_200_31", 8, 5, (200, 31), O.COEFF_INCLUSIONS, (2.0 ** -11, 0.2, 1e4, 1.0), 1234),
    ("target_cell_77_200", 8, 6, (77, 200), O.COEFF_PERIODIC, (1.0 / 64, 0.9999), 0),
    ("target_refcoef_cell_9_3", 8, 6, (9, 3), O.COEFF_REFERENCE, (), 0),
    ("constant_3p5", 4, 4, (3, 9), O.COEFF_CONSTANT, (3.5,), 0),
]


def main():
    out = {}
    for name, r, l, (ix, iy), kind, par, seed in CASES:
        m = morton(ix, iy, r)
        cor = O.coarse_corners(r, [m])
        c = O.coeff(kind, par, seed)
        res = O.run_cells(l, cor, c)
        resj = O.run_cells(l, cor, c, precond=O.PRECOND_JACOBI, max_iter=5000)
        n = 1 << l
        d = O.dof_map(l)
        phi = res["phi"][0]
        probes = [(n // 2, n // 2), (1, 1), (n - 1, 1), (n // 3, 2 * n // 3), (n // 4, n // 8)]
        out[name] = dict(
            r=r, l=l, cell=[ix, iy], morton=m, kind=kind, par=list(par), seed=seed,
            M=res["M"][0].tolist(), b=res["b"][0].tolist(),
            iters_ssor=res["iters"][0].tolist(), iters_jacobi=resj["iters"][0].tolist(),
            phi_norms=np.linalg.norm(phi, axis=1).tolist(),
            probes=[[jx, jy] for jx, jy in probes],
            phi_probes=[[float(phi[i][d[jy, jx]]) for i in range(4)] for jx, jy in probes],
            phi_weighted_sum=[float(np.dot(phi[i], np.cos(np.arange(phi.shape[1])))) for i in range(4)],
        )
        print(name, res["iters"][0], resj["iters"][0])
    with open(os.path.join(os.path.dirname(__file__), "oracle_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
