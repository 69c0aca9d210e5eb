"""Pins the CPU oracle (oracle/msfem_oracle.c) -- no GPU needed.

The reference has no numerical test and cannot be built here, so the oracle is pinned
against (1) the exact invariants of SURVEY.md Appendix B, (2) the survey-time
cross-check values, (3) the committed goldens it generated itself
(tests/golden/make_golden.py), so any later change of the oracle is caught.
"""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _morton(ix, iy, r):
    m = 0
    for bit in range(r):
        m |= ((ix >> bit) & 1) << (2 * bit) | ((iy >> bit) & 1) << (2 * bit + 1)
    return m


def test_dof_map_matches_survey_spot_values(oracle):
    sv = json.load(open(os.path.join(GOLD, "survey_crosscheck.json")))
    d = oracle.dof_map(7)
    n = 128
    assert [int(d[0, 0]), int(d[0, n]), int(d[n, 0]), int(d[n, n])] == sv["dof_corners_l7"]
    head = oracle.dof_map(3)[:3, :5]
    assert head.tolist() == sv["dof_map_head"]


@pytest.mark.parametrize("l", [1, 2, 3, 5, 6])
def test_dof_map_is_a_permutation_and_counts(oracle, l):
    n = 1 << l
    d = oracle.dof_map(l)
    assert sorted(d.ravel().tolist()) == list(range((n + 1) ** 2))
    bd = oracle.boundary_dofs(l)
    assert bd.size == 4 * n and np.all(np.diff(bd.astype(np.int64)) > 0)
    rowptr, col, val, F = oracle.assemble(l, oracle.coarse_corners(2, [5])[0],
                                          oracle.coeff(oracle.COEFF_CONSTANT, (1.0,)))
    assert rowptr[-1] == (3 * n + 1) ** 2
    # diagonal first in every row, then ascending columns
    for r in (0, 1, (n + 1) ** 2 // 2, (n + 1) ** 2 - 1):
        cols = col[rowptr[r]:rowptr[r + 1]]
        assert cols[0] == r and np.all(np.diff(cols[1:].astype(np.int64)) > 0)


def test_dof_map_first_touch_reference_walk(oracle):
    """Independent pure-Python first-touch walk over recursively refined cells."""
    l = 4
    n = 1 << l
    order = []

    def rec(x0, y0, size):
        if size == 1:
            order.append((x0, y0))
            return
        h = size // 2
        for c in range(4):  # deal.II child order: x is the low bit
            rec(x0 + (c & 1) * h, y0 + (c >> 1) * h, h)

    rec(0, 0, n)
    dof, nxt = {}, 0
    for ix, iy in order:
        for v in range(4):
            key = (ix + (v & 1), iy + (v >> 1))
            if key not in dof:
                dof[key] = nxt
                nxt += 1
    d = oracle.dof_map(l)
    for (jx, jy), k in dof.items():
        assert d[jy, jx] == k


def test_reference_coefficient_formula(oracle):
    """matrix_coeff.tpp:66-91 with the PI_D typo (coefficients.h:21)."""
    PI_D = 3.14592653509793218403
    c = oracle.coeff(oracle.COEFF_REFERENCE)
    for x, y in [(0.1, 0.7), (0.33, 0.01), (0.9991, 0.5)]:
        a = 1.0 - 0.9999 * (0.5 * np.sin(2 * PI_D * 57 * x) + 0.5 * np.sin(2 * PI_D * 57 * y))
        A = oracle.coeff_eval(c, x, y)
        assert abs(A[0, 0] - a) <= 4e-16 * abs(a) and abs(A[1, 1] - a) <= 4e-16 * abs(a)
        assert abs(A[0, 1]) <= 1e-15 and abs(A[1, 0]) <= 1e-15
        assert 1e-4 - 1e-12 <= a <= 1.9999 + 1e-12


def test_basis_q1_is_the_bilinear_interpolant(oracle):
    cor = oracle.coarse_corners(8, [_morton(200, 31, 8)])[0]
    coef = oracle.basis_q1_coeffs(cor)
    for i in range(4):
        for v in range(4):
            x, y = cor[v]
            val = coef[0, i] + coef[1, i] * x + coef[2, i] * y + coef[3, i] * x * y
            assert abs(val - (1.0 if i == v else 0.0)) < 1e-9


def test_survey_crosscheck_default_cell(oracle):
    sv = json.load(open(os.path.join(GOLD, "survey_crosscheck.json")))
    res = oracle.run_cells(7, oracle.coarse_corners(3, [0]), oracle.coeff(oracle.COEFF_REFERENCE))
    M, b, phi = res["M"][0], res["b"][0], res["phi"][0]
    assert np.abs(M - np.array(sv["M"])).max() < 1e-9
    assert np.abs(b - np.array(sv["b"])).max() < 1e-12
    d = oracle.dof_map(7)
    assert abs(phi[0][d[64, 64]] - sv["phi0_centre"]) < 1e-10
    assert abs(phi[3][d[64, 64]] - sv["phi3_centre"]) < 1e-10
    assert np.abs(np.linalg.norm(phi, axis=1) - np.array(sv["phi_norms"])).max() < 1e-8
    lo, hi = sv["iters_ssor_cell00"]
    assert all(lo <= k <= hi for k in res["iters"][0])


def test_invariants_partition_of_unity_rowsums_load(oracle):
    """SURVEY Appendix B invariants 1-3 and 6."""
    r, l = 3, 5
    H = 1.0 / (1 << r)
    cells = [_morton(2, 2, r), _morton(1, 6, r)]
    res = oracle.run_cells(l, oracle.coarse_corners(r, cells), oracle.coeff(oracle.COEFF_REFERENCE))
    for k in range(2):
        phi, M, b = res["phi"][k], res["M"][k], res["b"][k]
        assert np.abs(phi.sum(axis=0) - 1.0).max() < 5e-12
        assert np.abs(M - M.T).max() < 1e-12
        assert np.abs(M.sum(axis=1)).max() < 1e-12
        assert abs(b.sum() - 2.0 * H * H) < 1e-14
        ev = np.linalg.eigvalsh(0.5 * (M + M.T))
        assert ev[0] > -1e-12 and ev[1] > 1e-6
    # invariant 6: a(x,y) = a(y,x) on a diagonal coarse cell
    M, b = res["M"][0], res["b"][0]
    assert abs(M[1, 1] - M[2, 2]) < 1e-11 and abs(M[0, 1] - M[0, 2]) < 1e-11 and abs(b[1] - b[2]) < 1e-14


def test_constant_coefficient_closed_form(oracle):
    """SURVEY Appendix B invariant 4."""
    a0, r, l = 3.5, 4, 4
    H = 1.0 / (1 << r)
    cor = oracle.coarse_corners(r, [_morton(3, 9, r)])
    res = oracle.run_cells(l, cor, oracle.coeff(oracle.COEFF_CONSTANT, (a0,)))
    Mref = a0 * np.array([[2 / 3, -1 / 6, -1 / 6, -1 / 3], [-1 / 6, 2 / 3, -1 / 3, -1 / 6],
                          [-1 / 6, -1 / 3, 2 / 3, -1 / 6], [-1 / 3, -1 / 6, -1 / 6, 2 / 3]])
    assert np.abs(res["M"][0] - Mref).max() < 1e-11
    assert np.abs(res["b"][0] - H * H / 2).max() < 1e-15
    # phi_i is the coarse bilinear function itself
    n = 1 << l
    d = oracle.dof_map(l)
    jy, jx = np.meshgrid(np.arange(n + 1), np.arange(n + 1), indexing="ij")
    s, t = jx / n, jy / n
    exact = [(1 - s) * (1 - t), s * (1 - t), (1 - s) * t, s * t]
    for i in range(4):
        got = res["phi"][0][i][d]
        assert np.abs(got - exact[i]).max() < 1e-10


def test_condensed_system_solution_solves_interior_equations(oracle):
    """K phi_i = 0 on interior rows (what condense + PCG + distribute deliver)."""
    r, l = 5, 5
    cor = oracle.coarse_corners(r, [_morton(5, 7, r)])
    c = oracle.coeff(oracle.COEFF_PERIODIC, (1.0 / 64, 0.9999))
    res = oracle.run_cells(l, cor, c)
    rowptr, col, val, F = oracle.assemble(l, cor[0], c)
    import scipy.sparse as sp
    N = oracle.n_dofs(l)
    K = sp.csr_matrix((val, col.astype(np.int64), rowptr.astype(np.int64)), shape=(N, N))
    bd = oracle.boundary_dofs(l)
    interior = np.setdiff1d(np.arange(N), bd)
    for i in range(4):
        rvec = K @ res["phi"][0][i]
        assert np.linalg.norm(rvec[interior]) <= 1.0001e-12
        # constraint values are the BasisQ1 data
        assert np.abs(res["phi"][0][i][bd] - oracle.constraint_values(l, cor[0], i)).max() == 0.0
    # M = Phi^T K Phi, b = Phi^T F
    Phi = res["phi"][0]
    assert np.abs(Phi @ (K @ Phi.T) - res["M"][0]).max() < 1e-13
    assert np.abs(Phi @ F - res["b"][0]).max() < 1e-16
    assert abs(F.sum() - 2.0 / (1 << r) ** 2) < 1e-16


def test_ssor_and_jacobi_reach_the_same_bases(oracle):
    r, l = 8, 5
    cor = oracle.coarse_corners(r, [_morton(200, 31, r)])
    c = oracle.coeff(oracle.COEFF_INCLUSIONS, (2.0 ** -11, 0.2, 1e4, 1.0), 1234)
    a = oracle.run_cells(l, cor, c)
    b = oracle.run_cells(l, cor, c, precond=oracle.PRECOND_JACOBI, max_iter=5000)
    rel = np.linalg.norm(a["phi"] - b["phi"]) / np.linalg.norm(a["phi"])
    assert rel < 1e-10
    assert a["failed"] == 0 and b["failed"] == 0


def test_no_convergence_is_reported(oracle):
    res = oracle.run_cells(5, oracle.coarse_corners(3, [0]), oracle.coeff(oracle.COEFF_REFERENCE),
                           max_iter=5)
    assert res["failed"] == 1 and np.all(res["iters"] == 5)


def test_inclusion_coefficient_statistics(oracle):
    c = oracle.coeff(oracle.COEFF_INCLUSIONS, (2.0 ** -11, 0.2, 1e4, 1.0), 1234)
    bs = 2.0 ** -11
    rng = np.random.default_rng(0)
    pts = rng.integers(0, 2048, size=(4000, 2))
    vals = np.array([oracle.coeff_eval(c, (ix + 0.3) * bs, (iy + 0.6) * bs)[0, 0] for ix, iy in pts])
    assert set(np.unique(vals)) <= {1.0, 1e4}
    frac = np.mean(vals == 1e4)
    assert 0.17 < frac < 0.23
    # constant inside one block
    a = oracle.coeff_eval(c, 5.1 * bs, 7.2 * bs)[0, 0]
    assert a == oracle.coeff_eval(c, 5.9 * bs, 7.01 * bs)[0, 0]


def test_oracle_goldens_are_stable(oracle):
    gold = json.load(open(os.path.join(GOLD, "oracle_golden.json")))
    for name, g in gold.items():
        if g["l"] > 6:
            continue  # the n=128 cases are covered by the survey cross-check test
        cor = oracle.coarse_corners(g["r"], [g["morton"]])
        c = oracle.coeff(g["kind"], g["par"], g["seed"])
        res = oracle.run_cells(g["l"], cor, c)
        assert np.abs(res["M"][0] - np.array(g["M"])).max() < 1e-12, name
        assert np.abs(res["b"][0] - np.array(g["b"])).max() < 1e-15, name
        assert res["iters"][0].tolist() == g["iters_ssor"], name
        d = oracle.dof_map(g["l"])
        for (jx, jy), vals in zip(g["probes"], g["phi_probes"]):
            got = [res["phi"][0][i][d[jy, jx]] for i in range(4)]
            assert np.abs(np.array(got) - np.array(vals)).max() < 1e-12, name


def test_global_solution_is_the_weighted_sum(oracle):
    res = oracle.run_cells(4, oracle.coarse_corners(2, [3]), oracle.coeff(oracle.COEFF_REFERENCE))
    w = np.array([0.3, -1.2, 2.5, 0.7])
    got = oracle.global_solution(res["phi"][0], w)
    assert np.abs(got - w @ res["phi"][0]).max() < 1e-14


def test_openmp_batch_equals_serial(oracle):
    cor = oracle.coarse_corners(3)[:12]
    c = oracle.coeff(oracle.COEFF_REFERENCE)
    a = oracle.run_cells(4, cor, c, n_threads=1)
    b = oracle.run_cells(4, cor, c, n_threads=4)
    assert np.array_equal(a["M"], b["M"]) and np.array_equal(a["iters"], b["iters"])
    assert np.array_equal(a["phi"], b["phi"])


# ------------------------------------------------------------------------------ dim = 3
def test_dof_map3_is_first_touch_permutation(oracle):
    for l in (1, 2, 3):
        d = oracle.dof_map3(l)
        N = oracle.n_dofs3(l)
        assert sorted(d.ravel().tolist()) == list(range(N))
        # the first fine cell holds DoFs 0..7 in deal.II's lexicographic vertex order
        assert d[0:2, 0:2, 0:2].ravel().tolist() == list(range(8))
    # l = 1: second cell (child 1, +x) adds its four new vertices in vertex order
    d = oracle.dof_map3(1)
    assert [d[0, 0, 2], d[0, 1, 2], d[1, 0, 2], d[1, 1, 2]] == [8, 9, 10, 11]
    assert oracle.boundary_dofs3(2).size == 5 ** 3 - 3 ** 3


def test_matrix_coeff3_is_rotated_isotropic(oracle):
    """MatrixCoeff<3> (matrix_coeff.tpp:28-41, :66-91): R (a I) R^T = a I up to rounding."""
    c = oracle.coeff(oracle.COEFF_REFERENCE)
    A = oracle.coeff_eval3(c, 0.3, 0.2, 0.9)
    a2 = oracle.coeff_eval(c, 0.3, 0.2)[0, 0]
    assert np.allclose(A, a2 * np.eye(3), atol=1e-15)


def test_basis_q1_3d_is_nodal(oracle):
    cor = oracle.coarse_corners3(1, [5])[0]
    coef = oracle.basis_q1_coeffs3(cor)
    pm = np.array([[1, x, y, z, x * y, y * z, x * z, x * y * z] for x, y, z in cor])
    assert np.allclose(pm @ coef, np.eye(8), atol=1e-12)


def test_run_cells3_invariants(oracle):
    """SURVEY Appendix B carried to dim 3; a = const reproduces the trilinear element matrix."""
    import scipy.sparse as sp
    l = 3
    cor = oracle.coarse_corners3(1, [5])
    N = oracle.n_dofs3(l)
    for kind, par in ((oracle.COEFF_CONSTANT, (1.0,)), (oracle.COEFF_REFERENCE, ())):
        c = oracle.coeff(kind, par)
        rp, col, val, F = oracle.assemble3(l, cor[0], c)
        K = sp.csr_matrix((val, col.astype(np.int64), rp.astype(np.int64)), shape=(N, N))
        assert abs(K - K.T).max() < 1e-15
        assert np.abs(K @ np.ones(N)).max() < 1e-14
        assert abs(F.sum() - 2.0 * 0.125) < 1e-14
        r = oracle.run_cells3(l, cor, c)
        assert r["failed"] == 0 and (r["res"] <= 1e-12).all()
        assert np.abs(r["phi"][0].sum(axis=0) - 1).max() < 1e-10
        assert np.abs(r["M"][0].sum(axis=1)).max() < 1e-12
        assert abs(r["b"][0].sum() - 0.25) < 1e-12
        if kind == oracle.COEFF_CONSTANT:
            h = 0.5
            assert np.allclose(np.diag(r["M"][0]), h / 3, atol=1e-12)
            assert np.isclose(r["M"][0][0, 7], -h / 12, atol=1e-12)
            assert np.isclose(r["M"][0][0, 1], 0.0, atol=1e-12)
