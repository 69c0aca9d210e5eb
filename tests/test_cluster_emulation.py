"""Host-side check of the cluster / DSMEM kernel's index and synchronisation logic (no GPU).

scripts/emu/cluster_emu.cpp compiles msb_solve_cluster.cu UNCHANGED for the host: every CUDA
thread becomes an OS thread, __syncthreads / warp shuffles / st.async + mbarrier complete_tx /
tensor memory are emulated, shared and tensor memory start NaN-poisoned.  The emulated kernel
must reproduce a plain full-array multilevel PCG of the same condensed systems (the reference's
per-basis sequence, diffusion_problem_basis.tpp:450-465): identical iteration counts, bases to
1e-12, partition of unity, residual below the stopping tolerance.  This is a development / test
tool: nothing in the library, bench.py or the GPU tests routes through it.
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.isdir(CUDA_INC):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path_factory.mktemp("emu") / "cluster_emu")
    subprocess.check_call(
        ["g++", "-O1", "-std=c++20", "-pthread", "-DMSB_EMU", "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"),
         "-I" + os.path.join(ROOT, "mpi_parallel_multiscale_diffusion_fem_b200", "csrc"),
         os.path.join(ROOT, "scripts", "emu", "cluster_emu.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("l,tmem", [(5, 1), (5, 0), (6, 1), (6, 0)])
def test_emulated_cluster_kernel_equals_plain_pcg(emu, l, tmem):
    """l = 5 / 6: clusters of 2 / 4 CTAs (edge slabs only / edge and interior slabs); tmem = 1: four bases
    per pass with coefficients and x in (emulated) tensor memory, 0: two passes of two bases."""
    p = subprocess.run([emu, str(l), str(tmem)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout + p.stderr
    rows = re.findall(r"plain PCG basis (\d): iters (\d+) \(kernel (\d+)\)\s+rel diff ([0-9.e+-]+)", p.stdout)
    assert len(rows) == 4
    for _, it_plain, it_kernel, diff in rows:
        assert it_plain == it_kernel and 20 <= int(it_kernel) <= 40
        assert float(diff) < 1e-12
    m = re.search(r"partition of unity defect ([0-9.e+-]+), worst residual ([0-9.e+-]+)", p.stdout)
    assert float(m.group(1)) < 1e-10 and float(m.group(2)) <= 1e-12


@pytest.fixture(scope="module")
def emu_tsan(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.isdir(CUDA_INC):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path_factory.mktemp("emu_tsan") / "cluster_emu_tsan")
    p = subprocess.run(
        ["g++", "-O1", "-g", "-fsanitize=thread", "-std=c++20", "-pthread", "-DMSB_EMU", "-I" + CUDA_INC,
         "-I" + os.path.join(ROOT, "include"),
         "-I" + os.path.join(ROOT, "mpi_parallel_multiscale_diffusion_fem_b200", "csrc"),
         os.path.join(ROOT, "scripts", "emu", "cluster_emu.cpp"), "-o", exe], capture_output=True, text=True)
    if p.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + p.stderr[-200:])
    return exe


@pytest.mark.parametrize("tmem", [1, 0])
def test_emulated_cluster_kernel_is_race_free_under_thread_sanitizer(emu_tsan, tmem):
    """Every cross-CTA exchange of the kernel (st.async pushes into a peer's shared memory, mbarrier waits,
    buffer reuse from one iteration to the next) under ThreadSanitizer: a write that is not ordered after
    the owner's last read of the old content -- the hazard table of DESIGN.md 3.4 -- would be reported as
    a data race.  Cluster of 2 CTAs x 128 threads (l = 5), both kernel flavours."""
    p = subprocess.run([emu_tsan, "5", str(tmem)], capture_output=True, text=True, timeout=900)
    out = p.stdout + p.stderr
    if "unexpected memory mapping" in out or "FATAL: ThreadSanitizer" in out:
        pytest.skip("ThreadSanitizer cannot run in this container")
    assert "WARNING: ThreadSanitizer" not in out, out[-3000:]
    assert p.returncode == 0, out[-2000:]


def test_emulated_banded_inverse_of_the_7x7_level(tmp_path):
    """bpx::exact7_build (banded L D L^T + 49 parallel substitutions, used by the n = 32 kernel for the exact
    solve of its 7x7 coarse level) run by 128 emulated threads: A * A^-1 = I for a 9-point operator with
    coefficient jumps of 1, 1e4 and 1e6."""
    if shutil.which("g++") is None or not os.path.isdir(CUDA_INC):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path / "exact7_emu")
    subprocess.check_call(
        ["g++", "-O1", "-std=c++20", "-pthread", "-DMSB_EMU", "-I" + CUDA_INC, "-I" + os.path.join(ROOT, "include"),
         "-I" + os.path.join(ROOT, "mpi_parallel_multiscale_diffusion_fem_b200", "csrc"),
         os.path.join(ROOT, "scripts", "emu", "exact7_emu.cpp"), "-o", exe])
    for contrast in ("1", "1e4", "1e6"):
        p = subprocess.run([exe, contrast], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stdout + p.stderr


_FUSED_EXE = {}


@pytest.mark.parametrize("nl,kind,flavour", [(6, 1, 0), (6, 0, 1), (6, 2, 2), (5, 1, 0), (5, 2, 0), (5, 3, 0)])
def test_emulated_fused_stage_equals_a_plain_host_solve(tmp_path_factory, nl, kind, flavour):
    """scripts/emu/fused_emu.cpp compiles msb_solve_fused.cu UNCHANGED for the host (512 OS threads for the 64 x 64
    instantiation, 128 for the 32 x 32 one, shared / tensor memory NaN-poisoned) and compares the whole stage of one
    coarse cell -- on-chip assembly, the initial guess x_0 = g, 2 x 2 multilevel PCG solves, the element-matrix
    epilogue -- with a plain host implementation of the reference's sequence (basis.tpp:159-285, 450-465):
    coefficient kinds reference / periodic / inclusions / constant, the three residual flavours."""
    if shutil.which("g++") is None or not os.path.isdir(CUDA_INC):
        pytest.skip("needs g++ and the CUDA headers")
    if nl not in _FUSED_EXE:   # one build per instantiation
        exe = str(tmp_path_factory.mktemp("emu_fused") / ("fused_emu%d" % nl))
        subprocess.check_call(
            ["g++", "-O1", "-std=c++20", "-pthread", "-DMSB_EMU", "-DEMU_NL=%d" % nl, "-I" + CUDA_INC,
             "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "mpi_parallel_multiscale_diffusion_fem_b200", "csrc"),
             os.path.join(ROOT, "scripts", "emu", "fused_emu.cpp"), "-o", exe])
        _FUSED_EXE[nl] = exe
    exe = _FUSED_EXE[nl]
    p = subprocess.run([exe, str(kind), "500", str(flavour)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.rstrip().endswith("OK"), p.stdout + p.stderr
    its = [int(m) for m in re.findall(r"basis \d: iters (\d+)", p.stdout)]
    assert len(its) == 4 and (max(its) == 0 if kind == 3 else 10 <= max(its) <= 45)   # constant coefficient: x_0 = g is exact
