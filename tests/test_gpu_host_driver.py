"""End-to-end: the C++ driver (host/src/main.cxx, the reference's main.cxx shape) with the
basis stage on the GPU against an independent restatement of the coarse MsFEM problem
(ms.tpp:106-296) fed with the ORACLE's element matrices.  north_star: the final coarse
solution matches within 1e-8 relative L2."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PI_D = 3.14592653509793218403


def _exe(name):
    """Path of a host program; built on demand (the C++ mirror links the CUDA library's C ABI only)."""
    exe = os.path.join(ROOT, "host", "_build", name)
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    return exe


def _coarse_reference(oracle, r, l, co):
    """Coarse system from oracle (M, b): first-touch coarse DoFs, Dirichlet on x=0 / y=0
    ((x-.5)^2+(y-.5)^2, dirichlet_bc.tpp:22), Neumann cos(2 PI_D x) cos(2 PI_D y) on x=1 / y=1
    (neumann_bc.tpp:22) with 2-point Gauss on faces, direct solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    nc = 1 << r
    H = 1.0 / nc
    cor = oracle.coarse_corners(r)
    res = oracle.run_cells(l, cor, co, n_threads=os.cpu_count() or 1, keep_phi=False)
    assert res["failed"] == 0
    np_ = nc + 1
    dof = -np.ones((np_, np_), dtype=np.int64)
    nxt = 0
    cells = []
    for m in range(nc * nc):
        ix = sum(((m >> (2 * b)) & 1) << b for b in range(r))
        iy = sum(((m >> (2 * b + 1)) & 1) << b for b in range(r))
        ld = []
        for v in range(4):
            jx, jy = ix + (v & 1), iy + (v >> 1)
            if dof[jy, jx] < 0:
                dof[jy, jx] = nxt
                nxt += 1
            ld.append(dof[jy, jx])
        cells.append((ix, iy, ld))
    n = nxt
    K = sp.lil_matrix((n, n))
    f = np.zeros(n)
    g = [0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)]
    neu = lambda x, y: np.cos(2 * PI_D * x) * np.cos(2 * PI_D * y)
    for m, (ix, iy, ld) in enumerate(cells):
        Me, be = res["M"][m], res["b"][m].copy()
        if ix == nc - 1:  # face x = 1: vertices 1, 3
            for q in range(2):
                val = neu(1.0, (iy + g[q]) * H) * 0.5 * H
                be[1] += val * (1 - g[q])
                be[3] += val * g[q]
        if iy == nc - 1:  # face y = 1: vertices 2, 3
            for q in range(2):
                val = neu((ix + g[q]) * H, 1.0) * 0.5 * H
                be[2] += val * (1 - g[q])
                be[3] += val * g[q]
        for i in range(4):
            f[ld[i]] += be[i]
            for j in range(4):
                K[ld[i], ld[j]] += Me[i, j]
    K = K.tocsr()
    fixed = np.zeros(n, dtype=bool)
    gval = np.zeros(n)
    for jy in range(np_):
        for jx in range(np_):
            if jx == 0 or jy == 0:
                fixed[dof[jy, jx]] = True
                gval[dof[jy, jx]] = (jx * H - 0.5) ** 2 + (jy * H - 0.5) ** 2
    free = ~fixed
    u = gval.copy()
    u[free] = spla.spsolve(K[free][:, free].tocsc(), f[free] - K[free][:, fixed] @ gval[fixed])
    return u


@pytest.mark.parametrize("r,l,coeff,kind,par,seed", [
    (3, 7, "reference", 0, (), 0),                       # the reference's default run
    (4, 5, "periodic", 1, (1.0 / 64, 0.9999), 0),
    (4, 6, "inclusions", 2, (2.0 ** -11, 0.2, 1e4, 1.0), 1234),
])
def test_cpp_driver_coarse_solution(oracle, tmp_path, r, l, coeff, kind, par, seed):
    exe = _exe("msfem_main")
    dump = str(tmp_path / "coarse.txt")
    out = subprocess.run([exe, "--n-refine", str(r), "--n-refine-local", str(l), "--coeff", coeff,
                          "--dump", dump, "--output"], cwd=str(tmp_path), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "basis initialization and computation" in out.stdout
    lines = open(dump).read().split("\n")
    n = int(lines[0])
    u = np.array([float(v) for v in lines[1:1 + n]])
    ref = _coarse_reference(oracle, r, l, oracle.coeff(kind, par, seed))
    assert u.size == ref.size
    rel = np.linalg.norm(u - ref) / np.linalg.norm(ref)
    assert rel < 1e-8, rel
    # output files: one VTU per coarse cell named like the reference (basis.tpp:380-389) + records
    names = os.listdir(str(tmp_path))
    assert "solution-ms_fine-2d.pvtu" in names
    assert sum(nm.startswith("solution-ms_fine-2d.00000.cell-0_%d:" % r) for nm in names) == 4 ** r


def test_cpp_driver_reports_no_convergence_like_the_reference(tmp_path):
    """A failed local solve surfaces as an exception caught in main -> exit code 1
    (main.cxx:57-81)."""
    exe = _exe("msfem_main")
    out = subprocess.run([exe, "--coeff", "nonsense"], cwd=str(tmp_path), capture_output=True, text=True)
    assert out.returncode == 1 and "Exception on processing" in out.stderr


def test_cpp_driver_on_two_gpus_matches_one_gpu(tmp_path):
    """One handle and one host thread per device over contiguous Morton ranges (the ownership
    the reference gets from mpirun -n P): the coarse solution must not depend on P."""
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "mpi_parallel_multiscale_diffusion_fem_b200", "libmsfem_basis.so"))
    if lib.msb_device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = _exe("msfem_main")
    sols = []
    for gpus in (1, 2):
        dump = str(tmp_path / ("coarse%d.txt" % gpus))
        out = subprocess.run([exe, "--n-refine", "4", "--n-refine-local", "5", "--gpus", str(gpus),
                              "--dump", dump], cwd=str(tmp_path), capture_output=True, text=True)
        assert out.returncode == 0, out.stderr + out.stdout
        lines = open(dump).read().split("\n")
        n = int(lines[0])
        sols.append(np.array([float(v) for v in lines[1:1 + n]]))
    assert np.array_equal(sols[0], sols[1])     # same kernels, same cells: bitwise identical


def test_cpp_basis3d_matches_oracle(oracle, tmp_path):
    """The dim = 3 basis stage through the C++ mirror (DiffusionProblemBasis<3>::run_all):
    element matrices and right-hand sides of all 8 coarse hexes against the oracle."""
    exe = _exe("msfem_basis3d")
    dump = tmp_path / "mb3d.txt"
    out = subprocess.run([exe, "--n-refine", "1", "--n-refine-local", "3", "--dump", str(dump)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "8 coarse cells, 64 local solves" in out.stdout
    ref = oracle.run_cells3(3, oracle.coarse_corners3(1), oracle.coeff(oracle.COEFF_REFERENCE), n_threads=4,
                            keep_phi=False)
    rows = [ln.split() for ln in open(dump)]
    assert [r[0] for r in rows] == ["0_1:%d" % k for k in range(8)]      # CellId order = Morton
    for k, r in enumerate(rows):
        v = np.array(r[1:], dtype=np.float64)
        M, b = v[:64].reshape(8, 8), v[64:]
        assert np.linalg.norm(M - ref["M"][k]) / np.linalg.norm(ref["M"][k]) < 1e-8
        assert np.linalg.norm(b - ref["b"][k]) / np.linalg.norm(ref["b"][k]) < 1e-8


def test_cpp_truth_run_and_msfem_error(oracle, tmp_path):
    """the standard-FEM "truth" run (DiffusionProblem<2>, main.cxx:29-35) through the
    C++ mirror.  Checked against an independent direct solve of the oracle's CSR of the same fine
    problem, and the MsFEM reconstruction must be far closer to the fine solution than the coarse
    standard FEM is (the point of the method)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    exe = _exe("msfem_main")
    r, l = 2, 4                                   # 4x4 coarse cells x 16x16 fine cells = 64x64 mesh
    dump = str(tmp_path / "truth.txt")
    out = subprocess.run([exe, "--n-refine", str(r), "--n-refine-local", str(l), "--truth", "--output",
                          "--dump", dump], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr + out.stdout
    vals = dict(ln.split() for ln in open(dump).read().split("\n") if ln and ln[0].isalpha())
    ms_err, coarse_err = float(vals["ms_vs_fine_rel_l2"]), float(vals["coarse_vs_fine_rel_l2"])
    assert ms_err < 0.05, ms_err
    assert ms_err < 0.5 * coarse_err, (ms_err, coarse_err)
    assert os.path.exists(tmp_path / ("solution-std_2d_refinements-%d.0000.vtu" % (r + l)))
    assert os.path.exists(tmp_path / ("solution-std_2d_refinements-%d.pvtu" % r))

    # independent fine solve: the oracle's CSR of the unit square refined r+l times
    L = r + l
    n = 1 << L
    np_, h = n + 1, 1.0 / n
    co = oracle.coeff(oracle.COEFF_REFERENCE)
    unit = np.array([[0, 0], [1, 0], [0, 1], [1, 1]], dtype=np.float64)
    rowptr, col, val, F = oracle.assemble(L, unit, co)
    N = np_ * np_
    K = sp.csr_matrix((val, col.astype(np.int64), rowptr.astype(np.int64)), shape=(N, N))
    dof = oracle.dof_map(L)
    g = [0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)]
    b = F.copy()
    for k in range(n):
        for q in range(2):
            s = (k + g[q]) * h
            vx = np.cos(2 * PI_D * 1.0) * np.cos(2 * PI_D * s) * 0.5 * h
            b[dof[k, n]] += vx * (1 - g[q])
            b[dof[k + 1, n]] += vx * g[q]
            vy = np.cos(2 * PI_D * s) * np.cos(2 * PI_D * 1.0) * 0.5 * h
            b[dof[n, k]] += vy * (1 - g[q])
            b[dof[n, k + 1]] += vy * g[q]
    u = np.zeros(N)
    fixed = np.zeros(N, dtype=bool)
    for j in range(np_):
        for (jx, jy) in ((0, j), (j, 0)):
            d = dof[jy, jx]
            fixed[d] = True
            u[d] = (jx * h - 0.5) ** 2 + (jy * h - 0.5) ** 2
    free = ~fixed
    u[free] = spla.spsolve(K[free][:, free].tocsc(), b[free] - K[free][:, fixed] @ u[fixed])
    # the driver's fine solution, from its VTU (lexicographic point data)
    txt = open(tmp_path / ("solution-std_2d_refinements-%d.0000.vtu" % L)).read()
    data = txt.split('Name="u" format="ascii">')[1].split("</DataArray>")[0].split()
    u_cpp = np.array(data, dtype=np.float64).reshape(np_, np_)
    u_ref = u[dof.ravel()].reshape(np_, np_)
    assert np.linalg.norm(u_cpp - u_ref) / np.linalg.norm(u_ref) < 1e-8


def _coarse_reference_3d(oracle, r, l, co):
    """The dim = 3 coarse problem (ms.tpp:106-296 with dim = 3) fed with the ORACLE's 8x8 element
    matrices: first-touch coarse DoFs on the 3D Morton-ordered hexes, Dirichlet data on x=0 / y=0 / z=0,
    Neumann data on x=1 / y=1 / z=1 with 2x2 Gauss points and bilinear face shape values, direct solve."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    nc = 1 << r
    H = 1.0 / nc
    res = oracle.run_cells3(l, oracle.coarse_corners3(r), co, n_threads=os.cpu_count() or 1, keep_phi=False)
    assert res["failed"] == 0
    np_ = nc + 1
    dof = -np.ones((np_, np_, np_), dtype=np.int64)
    nxt, cells = 0, []
    for m in range(nc ** 3):
        idx = [sum(((m >> (3 * b + a)) & 1) << b for b in range(r)) for a in range(3)]
        ld = []
        for v in range(8):
            j = [idx[a] + ((v >> a) & 1) for a in range(3)]
            if dof[j[2], j[1], j[0]] < 0:
                dof[j[2], j[1], j[0]] = nxt
                nxt += 1
            ld.append(dof[j[2], j[1], j[0]])
        cells.append((idx, ld))
    n = nxt
    K = sp.lil_matrix((n, n))
    f = np.zeros(n)
    g = [0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)]
    neu = lambda x, y: np.cos(2 * PI_D * x) * np.cos(2 * PI_D * y)
    for m, (idx, ld) in enumerate(cells):
        be = res["b"][m].copy()
        for axis in range(3):
            if idx[axis] != nc - 1:
                continue
            others = [b for b in range(3) if b != axis]
            for q0 in range(2):
                for q1 in range(2):
                    t = [0.0, 0.0, 0.0]
                    t[axis] = 1.0
                    t[others[0]], t[others[1]] = g[q0], g[q1]
                    x = [(idx[b] + t[b]) * H for b in range(3)]
                    val = neu(x[0], x[1]) * (0.5 * H) ** 2
                    for v in range(8):
                        if not (v >> axis) & 1:
                            continue
                        shape = 1.0
                        for b in others:
                            shape *= t[b] if (v >> b) & 1 else 1.0 - t[b]
                        be[v] += val * shape
        for i in range(8):
            f[ld[i]] += be[i]
            for j in range(8):
                K[ld[i], ld[j]] += res["M"][m][i, j]
    u = np.zeros(n)
    fixed = np.zeros(n, dtype=bool)
    for jz in range(np_):
        for jy in range(np_):
            for jx in range(np_):
                if jx == 0 or jy == 0 or jz == 0:
                    d = dof[jz, jy, jx]
                    fixed[d] = True
                    u[d] = (jx * H - 0.5) ** 2 + (jy * H - 0.5) ** 2
    K = K.tocsr()
    free = ~fixed
    u[free] = spla.spsolve(K[free][:, free].tocsc(), f[free] - K[free][:, fixed] @ u[fixed])
    return u


@pytest.mark.parametrize("r,l", [(1, 3), (2, 2)])
def test_cpp_driver_3d_coarse_solution(oracle, tmp_path, r, l):
    """main.cxx:42-55 (the 3D block) through the C++ mirror: final coarse solution within 1e-8."""
    exe = _exe("msfem_main")
    dump = str(tmp_path / "coarse3d.txt")
    out = subprocess.run([exe, "--dim", "3", "--n-refine", str(r), "--n-refine-local", str(l), "--dump", dump,
                          "--output"], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr + out.stdout
    lines = open(dump).read().split("\n")
    n = int(lines[0])
    u = np.array([float(x) for x in lines[1:1 + n]])
    ref = _coarse_reference_3d(oracle, r, l, oracle.coeff(oracle.COEFF_REFERENCE))
    assert n == ref.size == ((1 << r) + 1) ** 3
    assert np.linalg.norm(u - ref) / np.linalg.norm(ref) < 1e-8
    assert os.path.exists(tmp_path / ("solution-ms_coarse-3d_refinements-%d.0000.vtu" % r))
    assert os.path.exists(tmp_path / "solution-ms_fine-3d.pvtu")
    assert len([f for f in os.listdir(tmp_path) if f.startswith("solution-ms_fine-3d.")]) == 8 ** r + 1


def test_cpp_truth_run_3d(tmp_path):
    """main.cxx:44-50: the 3D standard problems (DiffusionProblem<3>, 27-point operator through the C ABI)
    next to the 3D multiscale problem; MsFEM must beat the coarse standard FEM against the fine one."""
    exe = _exe("msfem_main")
    dump = str(tmp_path / "truth3d.txt")
    out = subprocess.run([exe, "--dim", "3", "--n-refine", "2", "--n-refine-local", "3", "--truth", "--dump", dump],
                         cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "Solving >> STANDARD << problem in 3D." in out.stdout
    vals = dict(ln.split() for ln in open(dump).read().split("\n") if ln and ln[0].isalpha())
    ms_err, coarse_err = float(vals["ms_vs_fine_rel_l2"]), float(vals["coarse_vs_fine_rel_l2"])
    assert ms_err < 0.05, ms_err
    assert ms_err < coarse_err, (ms_err, coarse_err)
