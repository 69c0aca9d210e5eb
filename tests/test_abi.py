"""The C-ABI library loads and exports every symbol include/msfem_basis.h declares.
No compute call is made: these tests run without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "msfem_basis.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(msb_[a-z_]+)\s*\(", txt)))


def test_header_symbols_are_exported(msb):
    lib = msb.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), "libmsfem_basis.so does not export " + name
    assert sorted(msb.EXPORTED_SYMBOLS) == declared


def test_version_and_error_string(msb):
    lib = msb.load_library()
    assert b"sm_100a" in lib.msb_version()
    assert isinstance(lib.msb_last_error(), bytes)


def test_argument_validation_happens_before_any_device_work(msb):
    from mpi_parallel_multiscale_diffusion_fem_b200.binding import Config, coeff_desc
    lib = msb.load_library()
    corners = msb.coarse_corners(1)
    h = C.c_void_p()

    def make(**kw):
        cfg = Config()
        cfg.abi_version, cfg.dim, cfg.n_refine_local, cfg.n_cells = 1, 2, 5, 4
        cfg.rhs_value = 2.0
        cfg.coeff = coeff_desc(msb.COEFF_REFERENCE)
        for k, v in kw.items():
            setattr(cfg, k, v)
        return cfg

    ptr = corners.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.msb_create(C.byref(make(abi_version=99)), ptr, None, C.byref(h)) == -1
    assert lib.msb_create(C.byref(make(dim=4)), ptr, None, C.byref(h)) == -2      # only dim 2 and 3
    assert lib.msb_create(C.byref(make(dim=3, n_refine_local=7)), ptr, None, C.byref(h)) == -2
    assert b"dim=3" in lib.msb_last_error()
    assert lib.msb_create(C.byref(make(n_refine_local=12)), ptr, None, C.byref(h)) == -2
    assert lib.msb_create(C.byref(make(n_cells=0)), ptr, None, C.byref(h)) == -1
    assert lib.msb_create(C.byref(make(coeff=coeff_desc(msb.COEFF_TABLE))), ptr, None, C.byref(h)) == -1
    assert lib.msb_create(None, ptr, None, C.byref(h)) == -1
    assert not h.value


def test_no_cpu_fallback_without_a_device(msb):
    lib = msb.load_library()
    if lib.msb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(msb.MsbError) as e:
        msb.BasisShard(5, msb.coarse_corners(1), msb.binding.coeff_desc(msb.COEFF_REFERENCE))
    assert e.value.code == -3 and "no CPU path" in str(e.value)


def test_product_package_never_touches_the_oracle():
    """The product (package, CUDA sources, C ABI header, C++ host mirror) never imports, links or
    calls anything under oracle/."""
    bad = re.compile(r"(import\s+oracle|from\s+oracle|msfem_oracle|orc_[a-z_]+\s*\(|oracle/_build)")
    roots = [os.path.join(ROOT, "mpi_parallel_multiscale_diffusion_fem_b200"),
             os.path.join(ROOT, "host"), os.path.join(ROOT, "include")]
    n = 0
    for root in roots:
        for dirpath, dirs, files in os.walk(root):
            dirs[:] = [d for d in dirs if d not in ("_build", "__pycache__")]
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ".cxx", "Makefile")):
                    src = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert not bad.search(src), f
                    n += 1
    assert n >= 12


def test_coarse_mesh_helpers(msb):
    c = msb.coarse_corners(3)
    assert c.shape == (64, 4, 2)
    # Morton order, x the low bit: cell 1 is to the right of cell 0, cell 2 above it
    assert np.allclose(c[1, 0], [0.125, 0.0]) and np.allclose(c[2, 0], [0.0, 0.125])
    assert np.allclose(c[63, 3], [1.0, 1.0])
    assert msb.cell_id_string(3, 5) == "0_3:011"
    for world in (1, 2, 3, 4, 8):
        ranges = [msb.morton_partition(64, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == 64
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))


def test_dealii_branch_of_the_host_mirror_compiles():
    """MSFEM_WITH_DEALII: the mirror's constructor with the reference's exact signature
    (diffusion_problem_basis.hpp:79-83) and the reference's construction loop (ms.tpp:54-66) written against it,
    compiled against API-compatible stubs of the deal.II headers (tests/fake_dealii; deal.II itself is absent)."""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("needs g++")
    p = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-DMSFEM_WITH_DEALII",
                        "-I" + os.path.join(ROOT, "tests", "fake_dealii"), "-I" + os.path.join(ROOT, "host", "include"),
                        "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "fake_dealii", "check_reference_ctor.cxx")],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
