"""scripts/dealii_dump: the diff that pins the oracle against the real reference must be mechanical the day a
deal.II machine appears.  No deal.II here, so this test writes a dump in exactly the
format the patched reference produces -- from the independent numpy restatement -- and runs the diff script on it:
parser, support-point -> vertex mapping and every comparison are exercised; a corrupted dump must be reported."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _write_dump(path, l, corners, res, cell_id, corrupt=False):
    n = 1 << l
    dof = res["dof"]
    N = (n + 1) ** 2
    c = np.array(corners, dtype=np.float64)
    pos = np.zeros((N, 2))
    for jy in range(n + 1):
        for jx in range(n + 1):
            s, t = jx / n, jy / n
            pos[dof[jy, jx]] = (1 - s) * (1 - t) * c[0] + s * (1 - t) * c[1] + (1 - s) * t * c[2] + s * t * c[3]
    with open(path, "w") as out:
        out.write("cell %s dim 2 n_refine_local %d n_dofs %d\n" % (cell_id, l, N))
        out.write("corners " + " ".join(repr(float(v)) for v in c.ravel()) + "\n")
        for i in range(N):
            phi = res["phi"][:, i].copy()
            if corrupt and i == N // 2:
                phi[1] += 1e-5
            out.write("dof %d %r %r F %r %s\n" % (i, float(pos[i, 0]), float(pos[i, 1]), float(res["F"][i]),
                                                 " ".join(repr(float(v)) for v in phi)))
        for k in range(4):
            for d, v in zip(res["boundary_dofs"], res["constraint_values"][k]):
                out.write("constraint %d %d %r\n" % (k, d, float(v)))
        for i in range(4):
            for j in range(4):
                out.write("M %d %d %r\n" % (i, j, float(res["M"][i, j])))
        for i in range(4):
            out.write("b %d %r\n" % (i, float(res["b"][i])))


def test_diff_script_on_a_dump_in_the_reference_format(tmp_path):
    import independent_restatement as IR
    l = 4
    corners = IR.corners_of((3, 2, 5))
    res = IR.run_cell(l, corners, 0, (), 0)
    good = tmp_path / "good"
    bad = tmp_path / "bad"
    good.mkdir()
    bad.mkdir()
    _write_dump(str(good / "basis_dump.cell-0_3:213.txt"), l, corners, res, "0_3:213")
    _write_dump(str(bad / "basis_dump.cell-0_3:213.txt"), l, corners, res, "0_3:213", corrupt=True)
    script = os.path.join(ROOT, "scripts", "dealii_dump", "diff_against_oracle.py")
    p = subprocess.run([sys.executable, script, str(good), "--independent"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "dof map EXACT, constraint sets EXACT" in p.stdout and "1 of 1 dumped cells match" in p.stdout
    p = subprocess.run([sys.executable, script, str(bad)], capture_output=True, text=True)
    assert p.returncode == 1 and "FAIL" in p.stdout


def test_patch_applies_to_the_reference_source(tmp_path):
    """The committed patch must apply cleanly to the reference's diffusion_problem_basis.tpp (when the reference tree
    is present: it is not on the GPU boxes)."""
    import shutil
    import pytest
    ref = "/root/reference/include/base/diffusion_problem_basis.tpp"
    if not os.path.exists(ref) or shutil.which("patch") is None:
        pytest.skip("reference tree or patch(1) not available")
    d = tmp_path / "include" / "base"
    d.mkdir(parents=True)
    shutil.copy(ref, str(d / "diffusion_problem_basis.tpp"))
    p = subprocess.run(["patch", "-p1", "--dry-run", "-i",
                        os.path.join(ROOT, "scripts", "dealii_dump", "basis_dump.patch")],
                       cwd=str(tmp_path), capture_output=True, text=True)
    assert p.returncode == 0 and "diffusion_problem_basis.tpp" in p.stdout, p.stdout + p.stderr
    assert os.path.getsize(os.path.join(ROOT, "scripts", "dealii_dump", "basis_dump.patch")) > 2000
