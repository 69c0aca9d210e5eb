"""dim = 3 parity of the sm_100a basis stage (through the C ABI) with the CPU oracle's
restatement of DiffusionProblemBasis<3> (diffusion_problem_basis.inst.cc:15-16).

Same bars as in 2D: DoF maps and constraint index sets bit-exact; bases within 1e-8 relative
L2; M (8x8), b (8) within 1e-8 relative; both sides converged to ||r||_2 <= 1e-12 absolute.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL_PHI = 1e-8
TOL_MB = 1e-8


def _coeffs(msb, oracle, kind, par=()):
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc
    return coeff_desc(kind, par, 0), oracle.coeff(kind, par, 0)


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def _skewed_hex():
    # a general straight-edged hexahedron (trilinear, non-affine mapping)
    c = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 0],
                  [0, 0, 1], [1, 0, 1], [0, 1, 1], [1, 1, 1]], dtype=np.float64)
    d = np.array([[0, 0, 0], [.1, -.05, .02], [-.04, .08, 0], [.12, .1, -.06],
                  [.03, .02, .07], [-.05, .04, .1], [.02, -.06, -.03], [.2, .15, .1]])
    return ((c + d) * 0.25 + 0.3)[None]


@pytest.mark.parametrize("l", [1, 2, 3, 4, 5])
def test_dof_map_bit_exact_3d(msb, oracle, l):
    cd, _ = _coeffs(msb, oracle, msb.COEFF_CONSTANT, (1.0,))
    with msb.BasisShard(l, oracle.coarse_corners3(0), cd, dim=3) as sh:
        assert np.array_equal(sh.dof_map(), oracle.dof_map3(l))


@pytest.mark.parametrize("l", [2, 4])
def test_constraint_sets_3d(msb, oracle, l):
    cd, _ = _coeffs(msb, oracle, msb.COEFF_REFERENCE)
    cor = oracle.coarse_corners3(2, [13, 62])
    with msb.BasisShard(l, cor, cd, dim=3) as sh:
        for cell in range(2):
            for ib in (0, 3, 5, 7):
                dofs, vals = sh.constraints(cell, ib)
                assert np.array_equal(dofs, oracle.boundary_dofs3(l))
                assert np.abs(vals - oracle.constraint_values3(l, cor[cell], ib)).max() < 1e-9


@pytest.mark.parametrize("l,kind,par,skew", [(3, 0, (), False), (4, 0, (), False), (3, 3, (2.5,), False),
                                             (3, 0, (), True)])
def test_matrix_free_operator_matches_csr_vmult_3d(msb, oracle, l, kind, par, skew):
    import scipy.sparse as sp
    cd, co = _coeffs(msb, oracle, kind, par)
    cor = _skewed_hex() if skew else oracle.coarse_corners3(2, [37])
    rowptr, col, val, F = oracle.assemble3(l, cor[0], co)
    N = oracle.n_dofs3(l)
    K = sp.csr_matrix((val, col.astype(np.int64), rowptr.astype(np.int64)), shape=(N, N))
    rng = np.random.default_rng(l)
    with msb.BasisShard(l, cor, cd, dim=3) as sh:
        for _ in range(2):
            x = rng.standard_normal(N)
            assert _rel(sh.apply_operator(0, x), K @ x) < 5e-14
        assert np.abs(sh.load_vector(0) - F).max() < 1e-15 * max(1.0, np.abs(F).max() / 1e-6)


@pytest.mark.parametrize("l,r,cells,kind,par", [
    (1, 1, [2, 4], 0, ()), (2, 1, [0, 7], 0, ()), (3, 2, [5, 21, 63], 0, ()), (4, 2, [9, 40], 0, ()), (3, 1, [3], 3, (1.0,)),
    (5, 3, [100], 0, ())])
def test_bases_and_element_matrices_match_oracle_3d(msb, oracle, l, r, cells, kind, par):
    cd, co = _coeffs(msb, oracle, kind, par)
    cor = oracle.coarse_corners3(r, cells)
    ref = oracle.run_cells3(l, cor, co, n_threads=4)
    assert ref["failed"] == 0
    with msb.BasisShard(l, cor, cd, dim=3) as sh:
        sh.run(1e-12, 1000)
        M, b = sh.element_matrices()
        it, res = sh.iteration_counts()
        assert M.shape == (len(cells), 8, 8) and it.shape == (len(cells), 8)
        assert (res <= 1e-12).all() and (it >= 0).all()   # (0: the initial guess x_0 = g already solves the system)
        for c in range(len(cells)):
            for ib in range(8):
                assert _rel(sh.basis(c, ib), ref["phi"][c, ib]) < TOL_PHI
            assert _rel(M[c], ref["M"][c]) < TOL_MB
            assert _rel(b[c], ref["b"][c]) < TOL_MB


def test_skewed_hex_matches_oracle(msb, oracle):
    cd, co = _coeffs(msb, oracle, msb.COEFF_REFERENCE)
    cor = _skewed_hex()
    ref = oracle.run_cells3(3, cor, co)
    with msb.BasisShard(3, cor, cd, dim=3) as sh:
        sh.run(1e-12, 1000)
        M, b = sh.element_matrices()
        for ib in range(8):
            assert _rel(sh.basis(0, ib), ref["phi"][0, ib]) < TOL_PHI
        assert _rel(M[0], ref["M"][0]) < TOL_MB and _rel(b[0], ref["b"][0]) < TOL_MB


def test_invariants_3d(msb, oracle):
    """SURVEY Appendix B in 3D: partition of unity, zero row sums, sum b = f |K|, symmetry."""
    cd, _ = _coeffs(msb, oracle, msb.COEFF_REFERENCE)
    r, l = 2, 4
    cor = oracle.coarse_corners3(r)
    with msb.BasisShard(l, cor, cd, dim=3) as sh:
        sh.run(1e-12, 1000)
        M, b = sh.element_matrices()
        assert np.abs(M.sum(axis=2)).max() < 1e-9
        assert np.abs(M - M.transpose(0, 2, 1)).max() < 1e-9
        assert np.abs(b.sum(axis=1) - 2.0 * (0.25 ** 3)).max() < 1e-11
        s = sum(sh.basis(17, ib) for ib in range(8))
        assert np.abs(s - 1.0).max() < 1e-9
        # multilevel preconditioner keeps the iteration count small
        it, _ = sh.iteration_counts()
        assert it.max() < 60
        w = np.random.default_rng(0).standard_normal((sh.n_cells, 8))
        sh.set_global_weights(w)
        ref = sum(w[17, ib] * sh.basis(17, ib) for ib in range(8))
        assert np.abs(sh.global_solution(17) - ref).max() < 1e-13


def test_constant_coefficient_reproduces_trilinear_basis(msb, oracle):
    """a = const on a brick: the multiscale bases are the trilinear Q1 functions themselves."""
    cd, _ = _coeffs(msb, oracle, msb.COEFF_CONSTANT, (1.0,))
    cor = oracle.coarse_corners3(1, [6])
    l = 3
    with msb.BasisShard(l, cor, cd, dim=3) as sh:
        sh.run(1e-12, 1000)
        M, _ = sh.element_matrices()
        h = 0.5
        assert np.allclose(np.diag(M[0]), h / 3, atol=1e-10)
        dm = sh.dof_map()
        n = 1 << l
        g = np.arange(n + 1) / n
        Z, Y, X = np.meshgrid(g, g, g, indexing="ij")
        phi7 = np.empty(sh.N)
        phi7[dm.ravel()] = (X * Y * Z).ravel()
        assert np.abs(sh.basis(0, 7) - phi7).max() < 1e-10


def test_unsupported_3d_requests_fail_loudly(msb, oracle):
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import MsbError, coeff_desc
    cor = oracle.coarse_corners3(0)
    with pytest.raises(MsbError):
        msb.BasisShard(3, cor, coeff_desc(msb.COEFF_PERIODIC, (0.1, 0.9)), dim=3)
    with pytest.raises(MsbError):
        msb.BasisShard(7, cor, coeff_desc(msb.COEFF_CONSTANT, (1.0,)), dim=3)
    with pytest.raises(MsbError):
        msb.BasisShard(3, cor, coeff_desc(msb.COEFF_CONSTANT, (1.0,)), dim=3, tier=msb.TIER_SMEM)


def test_no_convergence_is_reported_3d(msb, oracle):
    """max_iter too small: MSB_ERR_NO_CONVERGENCE + the failing (cell, basis), as the reference's
    SolverControl::NoConvergence (basis.tpp:303)."""
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import MsbError, coeff_desc
    cor = oracle.coarse_corners3(1, [0, 1])
    with msb.BasisShard(4, cor, coeff_desc(msb.COEFF_REFERENCE), dim=3) as sh:
        with pytest.raises(MsbError) as ei:
            sh.run(1e-12, 3)
        assert ei.value.code == -5
        cell, ib, res = sh.failure()
        assert (cell, ib) == (0, 0) and res > 1e-12
        it, _ = sh.iteration_counts()
        assert (it == 3).all()
        sh.run(1e-12, 1000)          # the handle stays usable
        assert sh.failure()[0] == -1


def test_handle_reuse_with_set_cells_3d(msb, oracle):
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc
    a, b = oracle.coarse_corners3(2, [3, 9]), oracle.coarse_corners3(2, [40, 41])
    cd = coeff_desc(msb.COEFF_REFERENCE)
    with msb.BasisShard(3, a, cd, dim=3) as sh, msb.BasisShard(3, b, cd, dim=3) as sh2:
        sh.run()
        Ma, _ = sh.element_matrices()
        sh.set_cells(b)
        sh.run()
        Mb, _ = sh.element_matrices()
        sh2.run()
        Mb2, _ = sh2.element_matrices()
        assert np.array_equal(Mb, Mb2)           # deterministic, independent of handle history
        assert not np.array_equal(Ma, Mb)


def test_more_than_65535_cells_3d(msb, oracle):
    """gridDim.y slices: a shard larger than 65 535 coarse cells runs in several launches."""
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc
    r = 6                                   # 64^3 = 262 144 coarse hexes; take 70 000 of them
    lo, hi = 100000, 170000
    cor = msb.coarse_corners3(r, lo, hi)
    with msb.BasisShard(1, cor, coeff_desc(msb.COEFF_REFERENCE), dim=3) as sh:
        sh.run(1e-12, 100)
        M, b = sh.element_matrices()
        it, res = sh.iteration_counts()
        assert (res <= 1e-12).all() and (it >= 0).all()   # (0: a single unknown whose initial guess x_0 = g is already exact)
        assert np.abs(M.sum(axis=2)).max() < 1e-12
        assert np.abs(b.sum(axis=1) - 2.0 / 64 ** 3).max() < 1e-17
        pick = [0, 65534, 65535, 65536, 69999]
        ref = oracle.run_cells3(1, cor[pick], oracle.coeff(oracle.COEFF_REFERENCE), keep_phi=False)
        for k, c in enumerate(pick):
            assert _rel(M[c], ref["M"][k]) < TOL_MB


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6])
def test_alternative_3d_kernels_agree_with_the_default(msb, oracle, variant):
    """variant 1: untiled K2, 2: K2 at 3 CTAs/SM, 3: general trilinear assembly instead of the brick
    tables, 4: one launch per coarse level instead of the fused kernel, 5: separate fine-level
    kernel, 6: plane-marching K2 instead of the whole-cell K2 of small meshes.  Same algorithm, so
    the same iteration counts and bases to rounding."""
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc
    cor = msb.coarse_corners3(3, 0, 320)          # >= 296 cells: the fused fine step is active
    cd = coeff_desc(msb.COEFF_REFERENCE)
    with msb.BasisShard(3, cor, cd, dim=3) as a, msb.BasisShard(3, cor, cd, dim=3, variant=variant) as b:
        a.run()
        b.run()
        ia, _ = a.iteration_counts()
        ib, _ = b.iteration_counts()
        assert np.abs(ia - ib).max() <= 1
        Ma, ba = a.element_matrices()
        Mb, bb = b.element_matrices()
        assert _rel(Ma, Mb) < 1e-10 and _rel(ba, bb) < 1e-10
        for c in (0, 319):
            assert _rel(a.basis(c, 5), b.basis(c, 5)) < 1e-10


@pytest.mark.parametrize("variant", [1, 7])
def test_k2_tilings_agree_at_16_cubed(msb, oracle, variant):
    """n = 16: the default K2 sweeps a whole z-column per CTA (one CTA per cell); variant 7 marches two z-chunks of 8
    planes (halo planes inside the cell), variant 1 is the untiled kernel.  Same sums up to their order."""
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc
    cor = msb.coarse_corners3(2, 5, 5 + 12)
    cd = coeff_desc(msb.COEFF_REFERENCE)
    with msb.BasisShard(4, cor, cd, dim=3) as a, msb.BasisShard(4, cor, cd, dim=3, variant=variant) as b:
        a.run()
        b.run()
        ia, ra = a.iteration_counts()
        ib, _ = b.iteration_counts()
        assert (ra <= 1e-12).all() and np.abs(ia - ib).max() <= 1
        Ma, ba = a.element_matrices()
        Mb, bb = b.element_matrices()
        assert _rel(Ma, Mb) < 1e-10 and _rel(ba, bb) < 1e-10
        for c in (0, 11):
            assert _rel(a.basis(c, 3), b.basis(c, 3)) < 1e-10


def test_largest_accepted_3d_local_mesh_l6(msb, oracle):
    """msb_create accepts n_refine_local = 6 for dim 3 (64^3 fine hexahedra, 274 625 DoFs per solve): DoF map
    bit-exact, the 8 bases / M / b of one coarse cell against the oracle, partition of unity."""
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc
    cor = msb.coarse_corners3(1, 6, 7)
    co = oracle.coeff(oracle.COEFF_REFERENCE)
    ref = oracle.run_cells3(6, cor, co)
    with msb.BasisShard(6, cor, coeff_desc(msb.COEFF_REFERENCE), dim=3) as sh:
        assert np.array_equal(sh.dof_map(), oracle.dof_map3(6))
        sh.run(1e-12, 5000)
        it, res = sh.iteration_counts()
        assert np.all(res <= 1e-12) and np.all(it > 0)
        phis = sh.bases()[0]
        for ib in range(8):
            e = np.linalg.norm(phis[ib] - ref["phi"][0][ib]) / np.linalg.norm(ref["phi"][0][ib])
            assert e < 1e-8, (ib, e)
        assert np.abs(phis.sum(axis=0) - 1.0).max() < 1e-9
        M, b = sh.element_matrices()
        assert np.abs(M[0] - ref["M"][0]).max() < 1e-8 * np.abs(ref["M"][0]).max()
        assert np.abs(b[0] - ref["b"][0]).max() < 1e-8 * np.abs(ref["b"][0]).max()


def _table3(oracle, cor, l, fn):
    """[C][n^3][8][9]: fn(x, y, z) -> 3x3 tensor at the fine quadrature points of axis-aligned bricks."""
    n = 1 << l
    g = [0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)]
    tab = np.empty((cor.shape[0], n ** 3, 8, 9))
    for c in range(cor.shape[0]):
        x0 = cor[c, 0]
        h = (cor[c, 7] - cor[c, 0]) / n
        for iz in range(n):
            for iy in range(n):
                for ix in range(n):
                    for q in range(8):
                        p = x0 + (np.array([ix, iy, iz]) + np.array([g[q & 1], g[(q >> 1) & 1], g[q >> 2]])) * h
                        tab[c, (iz * n + iy) * n + ix, q] = np.asarray(fn(*p)).ravel()
    return tab


def test_table_coefficient_3d(msb, oracle):
    """MSB_COEFF_TABLE for dim 3 (any host TensorFunction<2,3>, basis.tpp:202-203; round 1 rejected it):
    (a) the table of MatrixCoeff<3> reproduces the analytic kind; (b) a genuinely anisotropic, position-dependent
    SPD tensor against the oracle's tabulated 3D assembly; (c) an unsymmetric tensor is refused."""
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc
    l = 3
    cor = msb.coarse_corners3(1, 2, 4)
    co = oracle.coeff(oracle.COEFF_REFERENCE)
    tab = _table3(oracle, cor, l, lambda x, y, z: oracle.coeff_eval3(co, x, y, z))
    with msb.BasisShard(l, cor, coeff_desc(msb.COEFF_TABLE), table=tab, dim=3) as a, \
            msb.BasisShard(l, cor, coeff_desc(msb.COEFF_REFERENCE), dim=3) as b:
        a.run(1e-12, 5000)
        b.run(1e-12, 5000)
        Ma, ba = a.element_matrices()
        Mb, bb = b.element_matrices()
        assert _rel(Ma, Mb) < 1e-11 and _rel(ba, bb) < 1e-11

    def aniso(x, y, z):
        q = np.array([[1.0, 0.2 * np.sin(5 * x), 0.0], [0.0, 1.0, 0.3 * z], [0.0, 0.0, 1.0]])
        d = np.diag([1.0 + 0.5 * np.cos(7 * y), 2.0 + x, 0.3 + z * z])
        return q.T @ d @ q          # SPD, full, position dependent

    tab2 = _table3(oracle, cor, l, aniso)
    ref = oracle.run_cells3(l, cor, oracle.coeff(oracle.COEFF_CONSTANT, (1.0,)), table=tab2, n_threads=2)
    assert ref["failed"] == 0
    with msb.BasisShard(l, cor, coeff_desc(msb.COEFF_TABLE), table=tab2, dim=3) as sh:
        sh.run(1e-12, 5000)
        it, res = sh.iteration_counts()
        assert np.all(res <= 1e-12)
        M, b = sh.element_matrices()
        phis = sh.bases()
        for c in range(2):
            assert _rel(phis[c], ref["phi"][c]) < TOL_PHI
            assert _rel(M[c], ref["M"][c]) < TOL_MB and _rel(b[c], ref["b"][c]) < TOL_MB
            assert np.abs(phis[c].sum(axis=0) - 1.0).max() < 1e-9
    bad = tab2.copy()
    bad[1, 100, 3, 1] += 0.01       # a01 != a10
    with pytest.raises(msb.MsbError) as e:
        msb.BasisShard(l, cor, coeff_desc(msb.COEFF_TABLE), table=bad, dim=3)
    assert e.value.code == -1 and "not symmetric" in str(e.value)
