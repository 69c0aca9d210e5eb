"""ctypes loader for the CPU ORACLE (oracle/msfem_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / reference arm -- never from the product package.
Parity unpinned against the reference binary (deal.II is unavailable here);
pinned against SURVEY.md Appendix B invariants in tests/test_oracle.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmsfem_oracle.so")

COEFF_REFERENCE, COEFF_PERIODIC, COEFF_INCLUSIONS, COEFF_CONSTANT, COEFF_TABLE = range(5)
PRECOND_SSOR, PRECOND_JACOBI = 0, 1


class OrcCoeff(C.Structure):
    _fields_ = [("kind", C.c_int32), ("seed", C.c_int32), ("par", C.c_double * 6)]


def build(force=False):
    src = os.path.join(_HERE, "msfem_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.orc_basis_q1_value.restype = C.c_double
    return _lib


def coeff(kind, par=(), seed=0):
    c = OrcCoeff()
    c.kind, c.seed = kind, seed
    for i, v in enumerate(par):
        c.par[i] = v
    return c


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def n_dofs(l):
    return ((1 << l) + 1) ** 2


def dof_map(l):
    out = np.empty(n_dofs(l), dtype=np.uint32)
    lib().orc_dof_map(C.c_int(l), _p(out, C.c_uint32))
    return out.reshape((1 << l) + 1, (1 << l) + 1)


def boundary_dofs(l):
    out = np.empty(4 << l, dtype=np.uint32)
    k = lib().orc_boundary_dofs(C.c_int(l), _p(out, C.c_uint32))
    return out[:k]


def basis_q1_coeffs(corners):
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(8)
    out = np.empty(16, dtype=np.float64)
    lib().orc_basis_q1_coeffs(_p(corners, C.c_double), _p(out, C.c_double))
    return out.reshape(4, 4)


def coeff_eval(c, x, y):
    out = np.empty(4, dtype=np.float64)
    lib().orc_coeff_eval(C.byref(c), C.c_double(x), C.c_double(y), _p(out, C.c_double))
    return out.reshape(2, 2)


def constraint_values(l, corners, ib):
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(8)
    out = np.empty(4 << l, dtype=np.float64)
    lib().orc_constraint_values(C.c_int(l), _p(corners, C.c_double), C.c_int(ib), _p(out, C.c_double))
    return out


def assemble(l, corners, c, rhs_value=2.0, table=None):
    """Returns (rowptr, col, val, F) of the unconstrained fine stiffness matrix."""
    n = 1 << l
    N = n_dofs(l)
    nnz = (3 * n + 1) ** 2
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(8)
    rowptr = np.empty(N + 1, dtype=np.uint64)
    col = np.empty(nnz, dtype=np.uint32)
    val = np.empty(nnz, dtype=np.float64)
    F = np.empty(N, dtype=np.float64)
    tab = None if table is None else np.ascontiguousarray(table, dtype=np.float64)
    lib().orc_assemble(C.c_int(l), _p(corners, C.c_double), C.byref(c),
                       None if tab is None else _p(tab, C.c_double), C.c_double(rhs_value),
                       _p(rowptr, C.c_uint64), _p(col, C.c_uint32), _p(val, C.c_double),
                       _p(F, C.c_double))
    return rowptr, col, val, F


def run_cells(l, corners, c, rhs_value=2.0, tol=1e-12, max_iter=1000, precond=PRECOND_SSOR,
              omega=1.6, n_threads=1, keep_phi=True, table=None):
    """The reference's hot loop (ms.tpp:81-87) on the CPU.  corners: [C,4,2]."""
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 4, 2)
    nc = corners.shape[0]
    N = n_dofs(l)
    phi = np.empty((nc, 4, N), dtype=np.float64) if keep_phi else None
    M = np.empty((nc, 4, 4), dtype=np.float64)
    b = np.empty((nc, 4), dtype=np.float64)
    iters = np.empty((nc, 4), dtype=np.int32)
    res = np.empty((nc, 4), dtype=np.float64)
    tab = None if table is None else np.ascontiguousarray(table, dtype=np.float64)
    failed = lib().orc_run_cells(
        C.c_int(l), C.c_int(nc), _p(corners, C.c_double), C.byref(c),
        None if tab is None else _p(tab, C.c_double), C.c_double(rhs_value), C.c_double(tol),
        C.c_int(max_iter), C.c_int(precond), C.c_double(omega), C.c_int(n_threads),
        None if phi is None else _p(phi, C.c_double), _p(M, C.c_double), _p(b, C.c_double),
        _p(iters, C.c_int32), _p(res, C.c_double))
    return dict(phi=phi, M=M, b=b, iters=iters, res=res, failed=failed)


def global_solution(phi, w):
    phi = np.ascontiguousarray(phi, dtype=np.float64)
    w = np.ascontiguousarray(w, dtype=np.float64)
    out = np.empty(phi.shape[1], dtype=np.float64)
    lib().orc_global_solution(C.c_int(phi.shape[1]), _p(phi, C.c_double), _p(w, C.c_double),
                              _p(out, C.c_double))
    return out


def coarse_corners(r, cells=None):
    """Corner points [C,4,2] of the 2^r x 2^r coarse mesh on [0,1]^2 in Morton
    (CellId / p4est) order, deal.II vertex order."""
    nc = 1 << r
    H = 1.0 / nc
    m = np.arange(nc * nc, dtype=np.uint64) if cells is None else np.asarray(cells, dtype=np.uint64)

    def compact(v):
        v = v & np.uint64(0x5555555555555555)
        v = (v | (v >> np.uint64(1))) & np.uint64(0x3333333333333333)
        v = (v | (v >> np.uint64(2))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v >> np.uint64(4))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v >> np.uint64(8))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v >> np.uint64(16))) & np.uint64(0x00000000FFFFFFFF)
        return v

    ix = compact(m).astype(np.float64)
    iy = compact(m >> np.uint64(1)).astype(np.float64)
    out = np.empty((m.size, 4, 2), dtype=np.float64)
    for v in range(4):
        out[:, v, 0] = (ix + (v & 1)) * H
        out[:, v, 1] = (iy + (v >> 1)) * H
    return out


# ------------------------------------------------------------------------ 3D --

def n_dofs3(l):
    return ((1 << l) + 1) ** 3


def n_boundary3(l):
    n = 1 << l
    return (n + 1) ** 3 - (n - 1) ** 3


def dof_map3(l):
    np_ = (1 << l) + 1
    out = np.empty(n_dofs3(l), dtype=np.uint32)
    lib().orc3_dof_map(C.c_int(l), _p(out, C.c_uint32))
    return out.reshape(np_, np_, np_)


def boundary_dofs3(l):
    out = np.empty(n_boundary3(l), dtype=np.uint32)
    k = lib().orc3_boundary_dofs(C.c_int(l), _p(out, C.c_uint32))
    return out[:k]


def basis_q1_coeffs3(corners):
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(24)
    out = np.empty(64, dtype=np.float64)
    lib().orc3_basis_q1_coeffs(_p(corners, C.c_double), _p(out, C.c_double))
    return out.reshape(8, 8)


def coeff_eval3(c, x, y, z):
    out = np.empty(9, dtype=np.float64)
    lib().orc3_coeff_eval(C.byref(c), C.c_double(x), C.c_double(y), C.c_double(z), _p(out, C.c_double))
    return out.reshape(3, 3)


def constraint_values3(l, corners, ib):
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(24)
    out = np.empty(n_boundary3(l), dtype=np.float64)
    lib().orc3_constraint_values(C.c_int(l), _p(corners, C.c_double), C.c_int(ib), _p(out, C.c_double))
    return out


def assemble3(l, corners, c, rhs_value=2.0):
    N = n_dofs3(l)
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(24)
    rowptr = np.empty(N + 1, dtype=np.uint64)
    col = np.empty(27 * N, dtype=np.uint32)
    val = np.empty(27 * N, dtype=np.float64)
    F = np.empty(N, dtype=np.float64)
    f = lib().orc3_assemble
    f.restype = C.c_uint64
    nnz = f(C.c_int(l), _p(corners, C.c_double), C.byref(c), C.c_double(rhs_value),
            _p(rowptr, C.c_uint64), _p(col, C.c_uint32), _p(val, C.c_double), _p(F, C.c_double))
    return rowptr, col[:nnz], val[:nnz], F


def run_cells3(l, corners, c, rhs_value=2.0, tol=1e-12, max_iter=1000, precond=PRECOND_SSOR,
               omega=1.6, n_threads=1, keep_phi=True, table=None):
    """DiffusionProblemBasis<3>::run() over the given coarse hexes.  corners: [C,8,3]; table: tabulated tensors
    [C][n^3][8][9] (a user TensorFunction<2,3>::value_list at the fine quadrature points) or None."""
    corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, 8, 3)
    nc = corners.shape[0]
    N = n_dofs3(l)
    phi = np.empty((nc, 8, N), dtype=np.float64) if keep_phi else None
    M = np.empty((nc, 8, 8), dtype=np.float64)
    b = np.empty((nc, 8), dtype=np.float64)
    iters = np.empty((nc, 8), dtype=np.int32)
    res = np.empty((nc, 8), dtype=np.float64)
    tab = None if table is None else np.ascontiguousarray(table, dtype=np.float64)
    failed = lib().orc3_run_cells_table(
        C.c_int(l), C.c_int(nc), _p(corners, C.c_double), C.byref(c),
        None if tab is None else _p(tab, C.c_double), C.c_double(rhs_value),
        C.c_double(tol), C.c_int(max_iter), C.c_int(precond), C.c_double(omega), C.c_int(n_threads),
        None if phi is None else _p(phi, C.c_double), _p(M, C.c_double), _p(b, C.c_double),
        _p(iters, C.c_int32), _p(res, C.c_double))
    return dict(phi=phi, M=M, b=b, iters=iters, res=res, failed=failed)


def coarse_corners3(r, cells=None):
    """Corner points [C,8,3] of the (2^r)^3 coarse mesh on [0,1]^3 in 3D Morton order."""
    nc = 1 << r
    H = 1.0 / nc
    m = np.arange(nc ** 3, dtype=np.uint64) if cells is None else np.asarray(cells, dtype=np.uint64)
    idx = np.zeros((3, m.size), dtype=np.uint64)
    for bit in range(r):
        for a in range(3):
            idx[a] |= ((m >> np.uint64(3 * bit + a)) & np.uint64(1)) << np.uint64(bit)
    out = np.empty((m.size, 8, 3), dtype=np.float64)
    for v in range(8):
        out[:, v, 0] = (idx[0] + (v & 1)) * H
        out[:, v, 1] = (idx[1] + ((v >> 1) & 1)) * H
        out[:, v, 2] = (idx[2] + (v >> 2)) * H
    return out
