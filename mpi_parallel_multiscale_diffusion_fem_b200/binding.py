"""ctypes binding of the C ABI in include/msfem_basis.h (libmsfem_basis.so).

No CPU fallback: if the CUDA library is missing, or there is no sm_100 device, every
compute entry raises.  Nothing here imports or calls oracle/.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

COEFF_REFERENCE, COEFF_PERIODIC, COEFF_INCLUSIONS, COEFF_CONSTANT, COEFF_TABLE = range(5)
TIER_AUTO, TIER_SMEM, TIER_STREAMED = range(3)
ABI_VERSION = 1

MSB_OK = 0
MSB_ERR_NO_CONVERGENCE = -5

# every symbol include/msfem_basis.h declares
EXPORTED_SYMBOLS = [
    "msb_create", "msb_set_cells", "msb_run", "msb_run_async", "msb_sync", "msb_get_failure",
    "msb_get_element_matrices", "msb_get_iteration_counts", "msb_get_basis", "msb_get_dof_map",
    "msb_get_constraints", "msb_apply_operator", "msb_get_load_vector", "msb_set_global_weights",
    "msb_get_global_solution", "msb_get_run_stats", "msb_get_algorithmic_bytes", "msb_destroy",
    "msb_last_error", "msb_device_count", "msb_version",
    "msb_get_bases", "msb_get_global_solutions", "msb_get_device_results", "msb_build_id",
    "msb_run_with_bases",
]


class MsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("msfem_basis error %d: %s" % (code, msg))
        self.code = code


class CoeffDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("seed", C.c_int32), ("par", C.c_double * 6)]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("dim", C.c_int32), ("n_refine_local", C.c_int32),
                ("n_cells", C.c_int32), ("device_id", C.c_int32), ("tier", C.c_int32),
                ("variant", C.c_int32), ("reserved", C.c_int32), ("rhs_value", C.c_double),
                ("coeff", CoeffDesc)]


def lib_path():
    # MSB_LIBRARY selects another build of the SAME library (e.g. the stage-timer profiling build)
    return os.environ.get("MSB_LIBRARY") or os.path.join(_HERE, "libmsfem_basis.so")


_lib = None


def load_library():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise MsbError(-3, "CUDA library %s is not built (run __graft_entry__.build()); "
                               "there is no CPU fallback for the basis stage" % p)
        lib = C.CDLL(p)
        lib.msb_last_error.restype = C.c_char_p
        lib.msb_version.restype = C.c_char_p
        lib.msb_build_id.restype = C.c_char_p
        lib.msb_run.argtypes = [C.c_void_p, C.c_double, C.c_int32]
        lib.msb_run_async.argtypes = [C.c_void_p, C.c_double, C.c_int32, C.c_void_p]
        for name in ("msb_sync", "msb_destroy"):
            getattr(lib, name).argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def build_id():
    """Identity of the kernel sources the loaded library was compiled from (msb_build_id)."""
    return load_library().msb_build_id().decode()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def coeff_desc(kind, par=(), seed=0):
    d = CoeffDesc()
    d.kind, d.seed = int(kind), int(seed)
    for i, v in enumerate(par):
        d.par[i] = float(v)
    return d


class BasisShard:
    """The coarse cells one GPU owns; replaces the std::map<CellId, DiffusionProblemBasis<dim>>
    of diffusion_problem_ms.hpp:230 together with the loop ms.tpp:81-87 that runs it."""

    def __init__(self, n_refine_local, corners, coeff, rhs_value=2.0, device_id=0,
                 tier=TIER_AUTO, variant=0, table=None, dim=2):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.dim = int(dim)
        self.nb = 1 << self.dim  # bases per coarse cell (GeometryInfo<dim>::vertices_per_cell)
        corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, self.nb, self.dim)
        self.n_cells = corners.shape[0]
        self.l = int(n_refine_local)
        self.n = 1 << self.l
        self.N = (self.n + 1) ** self.dim
        self.n_boundary = self.N - (self.n - 1) ** self.dim
        cfg = Config()
        cfg.abi_version, cfg.dim, cfg.n_refine_local = ABI_VERSION, self.dim, self.l
        cfg.n_cells, cfg.device_id, cfg.tier, cfg.variant = self.n_cells, device_id, tier, variant
        cfg.rhs_value = rhs_value
        cfg.coeff = coeff
        tab = None
        if table is not None:
            tab = np.ascontiguousarray(table, dtype=np.float64)
            assert tab.size == self.n_cells * self.n ** self.dim * self.nb * self.dim * self.dim
        self._check(self._lib.msb_create(C.byref(cfg), _dp(corners),
                                         None if tab is None else _dp(tab), C.byref(self._h)))

    # -- plumbing ------------------------------------------------------------------------
    def _check(self, rc, allow=()):
        if rc != MSB_OK and rc not in allow:
            raise MsbError(rc, self._lib.msb_last_error().decode())
        return rc

    def close(self):
        if self._h:
            self._lib.msb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_cells(self, corners, table=None):
        corners = np.ascontiguousarray(corners, dtype=np.float64).reshape(-1, self.nb, self.dim)
        assert corners.shape[0] == self.n_cells
        tab = None if table is None else np.ascontiguousarray(table, dtype=np.float64)
        self._check(self._lib.msb_set_cells(self._h, _dp(corners), None if tab is None else _dp(tab)))

    # -- the stage -----------------------------------------------------------------------
    def run(self, tol=1e-12, max_iter=1000, allow_no_convergence=False):
        allow = (MSB_ERR_NO_CONVERGENCE,) if allow_no_convergence else ()
        return self._check(self._lib.msb_run(self._h, tol, max_iter), allow)

    def run_async(self, tol=1e-12, max_iter=1000, stream=None):
        self._check(self._lib.msb_run_async(self._h, tol, max_iter, C.c_void_p(stream or 0)))

    def run_with_bases(self, tol=1e-12, max_iter=1000, out=None, out_addr=None):
        """msb_run_with_bases: the stage with the device->host copy of the bases pipelined behind the solves."""
        if out_addr is None:
            if out is None:
                out = np.empty((self.n_cells, self.nb, self.N), dtype=np.float64)
            out_addr = out.ctypes.data
        self._lib.msb_run_with_bases.argtypes = [C.c_void_p, C.c_double, C.c_int32, C.c_void_p]
        self._check(self._lib.msb_run_with_bases(self._h, tol, max_iter, C.c_void_p(out_addr)))
        return out

    def sync(self, allow_no_convergence=False):
        allow = (MSB_ERR_NO_CONVERGENCE,) if allow_no_convergence else ()
        return self._check(self._lib.msb_sync(self._h), allow)

    # -- raw-pointer variants (caller-owned, e.g. pinned, host buffers) ---------------------
    def set_cells_ptr(self, corners_addr):
        self._check(self._lib.msb_set_cells(self._h, C.c_void_p(corners_addr), None))

    def element_matrices_into(self, M_addr, b_addr):
        self._check(self._lib.msb_get_element_matrices(self._h, C.c_void_p(M_addr), C.c_void_p(b_addr)))

    def iteration_counts_into(self, it_addr, res_addr=0):
        self._check(self._lib.msb_get_iteration_counts(self._h, C.c_void_p(it_addr),
                                                       C.c_void_p(res_addr or None)))

    # -- accessors -----------------------------------------------------------------------
    def element_matrices(self):
        M = np.empty((self.n_cells, self.nb, self.nb), dtype=np.float64)
        b = np.empty((self.n_cells, self.nb), dtype=np.float64)
        self._check(self._lib.msb_get_element_matrices(self._h, _dp(M), _dp(b)))
        return M, b

    def iteration_counts(self):
        it = np.empty((self.n_cells, self.nb), dtype=np.int32)
        res = np.empty((self.n_cells, self.nb), dtype=np.float64)
        self._check(self._lib.msb_get_iteration_counts(
            self._h, it.ctypes.data_as(C.POINTER(C.c_int32)), _dp(res)))
        return it, res

    def failure(self):
        cell, ib, res = C.c_int32(), C.c_int32(), C.c_double()
        self._check(self._lib.msb_get_failure(self._h, C.byref(cell), C.byref(ib), C.byref(res)))
        return cell.value, ib.value, res.value

    def basis(self, cell, index_basis):
        out = np.empty(self.N, dtype=np.float64)
        self._check(self._lib.msb_get_basis(self._h, C.c_int32(cell), C.c_int32(index_basis), _dp(out)))
        return out

    def bases(self, cell0=0, n_cells=None, out=None):
        """msb_get_bases: [n_cells, 2^dim, N] in the deal.II DoF order, one call."""
        n_cells = self.n_cells - cell0 if n_cells is None else n_cells
        if out is None:
            out = np.empty((n_cells, self.nb, self.N), dtype=np.float64)
        self._check(self._lib.msb_get_bases(self._h, C.c_int32(cell0), C.c_int32(n_cells), _dp(out)))
        return out

    def bases_into(self, cell0, n_cells, out_addr):
        self._check(self._lib.msb_get_bases(self._h, C.c_int32(cell0), C.c_int32(n_cells), C.c_void_p(out_addr)))

    def global_solutions(self, cell0=0, n_cells=None):
        n_cells = self.n_cells - cell0 if n_cells is None else n_cells
        out = np.empty((n_cells, self.N), dtype=np.float64)
        self._check(self._lib.msb_get_global_solutions(self._h, C.c_int32(cell0), C.c_int32(n_cells), _dp(out)))
        return out

    def device_results(self):
        """Device addresses of (M, b, iteration counts) of the last run (msb_get_device_results)."""
        m, b, it = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.msb_get_device_results(self._h, C.byref(m), C.byref(b), C.byref(it)))
        return m.value, b.value, it.value

    def dof_map(self):
        out = np.empty(self.N, dtype=np.uint32)
        self._check(self._lib.msb_get_dof_map(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32))))
        return out.reshape((self.n + 1,) * self.dim)

    def constraints(self, cell, index_basis):
        dofs = np.empty(self.n_boundary, dtype=np.uint32)
        vals = np.empty(self.n_boundary, dtype=np.float64)
        self._check(self._lib.msb_get_constraints(
            self._h, C.c_int32(cell), C.c_int32(index_basis),
            dofs.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(vals)))
        return dofs, vals

    def apply_operator(self, cell, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.N, dtype=np.float64)
        self._check(self._lib.msb_apply_operator(self._h, C.c_int32(cell), _dp(x), _dp(y)))
        return y

    def load_vector(self, cell):
        F = np.empty(self.N, dtype=np.float64)
        self._check(self._lib.msb_get_load_vector(self._h, C.c_int32(cell), _dp(F)))
        return F

    def set_global_weights(self, w):
        w = np.ascontiguousarray(w, dtype=np.float64).reshape(self.n_cells, self.nb)
        self._check(self._lib.msb_set_global_weights(self._h, _dp(w)))

    def global_solution(self, cell):
        out = np.empty(self.N, dtype=np.float64)
        self._check(self._lib.msb_get_global_solution(self._h, C.c_int32(cell), _dp(out)))
        return out

    def run_stats(self):
        t, ts, nl, tier = C.c_float(), C.c_float(), C.c_int32(), C.c_int32()
        self._check(self._lib.msb_get_run_stats(self._h, C.byref(t), C.byref(ts), C.byref(nl),
                                                C.byref(tier)))
        return dict(ms_total=t.value, ms_solve=ts.value, launches=nl.value, tier=tier.value)

    def algorithmic_bytes(self):
        b, k = C.c_double(), C.c_double()
        self._check(self._lib.msb_get_algorithmic_bytes(self._h, C.byref(b), C.byref(k)))
        return b.value, k.value
