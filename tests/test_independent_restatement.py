"""The C oracle against a SECOND, independent restatement of the reference's local-basis stage
(tests/golden/independent_restatement.py: pure numpy / scipy, own DoF numbering and assembly, sparse DIRECT
solve, no code shared with oracle/msfem_oracle.c) -- no GPU needed.

What this buys: the reference pins nothing numerically and cannot be built here (no deal.II), so every parity
claim rests on restatements.  Two restatements written independently from the reference's source lines and from
SURVEY.md Appendix A -- one in C with SSOR-PCG, one in numpy with a direct solver -- must agree on every quantity
the north star names: DoF maps and constraint index sets bit for bit, bases / M / b to solver accuracy.  The
committed vectors (independent_golden.json) were produced by running the numpy restatement; the first test
re-runs it on the small cases so the file cannot drift from its generator.
"""
import json
import os
import sys

import numpy as np
import pytest

GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD_DIR)
CASES = json.load(open(os.path.join(GOLD_DIR, "independent_golden.json")))
TOL = 1e-8  # north-star tolerance (observed agreement: 1e-10 .. 1e-13)


def _checksum(a):
    return int(sum((i + 1) * int(d) for i, d in enumerate(np.asarray(a).ravel())) % (1 << 61))


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


@pytest.mark.parametrize("name", ["ref_l3_cell_2_5", "constant_l3_f35", "periodic_l4", "ref_l4_skewed_quad"])
def test_committed_vectors_are_what_the_generator_produces(name):
    import independent_restatement as IR
    l, spec, kind, par, seed, f = IR.CASES[name]
    res = IR.run_cell(l, IR.corners_of(spec), kind, par, seed, f)
    g = CASES[name]
    assert np.allclose(res["M"], np.array(g["M"]), rtol=1e-12, atol=1e-14)
    assert np.allclose(res["b"], np.array(g["b"]), rtol=1e-12, atol=1e-18)
    assert _checksum(res["dof"]) == g["dof_checksum"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_the_independent_restatement(oracle, name):
    g = CASES[name]
    l, n = g["l"], 1 << g["l"]
    cor = np.array(g["corners"], dtype=np.float64)
    # integers: bit-exact
    d = oracle.dof_map(l)
    assert [int(d[0, 0]), int(d[0, n]), int(d[n, 0]), int(d[n, n])] == g["dof_corners"]
    assert _checksum(d) == g["dof_checksum"]
    bd = oracle.boundary_dofs(l)
    assert bd.size == g["n_boundary"] and _checksum(bd) == g["boundary_dofs_checksum"]
    # constraint values (BasisQ1 data): both sides invert the same 4x4 point matrix, ~eps/H^2 apart
    for ib in range(4):
        vals = oracle.constraint_values(l, cor, ib)
        assert np.abs(vals[:6] - np.array(g["constraint_values_head"][ib])).max() < 1e-9
    # floating point: bases, M, b
    co = oracle.coeff(g["kind"], g["par"], g["seed"])
    ref = oracle.run_cells(l, cor[None], co, rhs_value=g["f"])
    assert ref["failed"] == 0
    assert _rel(ref["M"][0], np.array(g["M"])) < TOL
    assert _rel(ref["b"][0], np.array(g["b"])) < TOL
    for (jx, jy), vals in zip(g["probes"], g["phi_probes"]):
        got = np.array([ref["phi"][0][ib][d[jy, jx]] for ib in range(4)])
        assert np.abs(got - np.array(vals)).max() < TOL
    for ib in range(4):
        assert abs(np.linalg.norm(ref["phi"][0][ib]) - g["phi_norms"][ib]) < TOL * g["phi_norms"][ib]
