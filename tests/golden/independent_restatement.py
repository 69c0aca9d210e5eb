#!/usr/bin/env python
"""A SECOND, independent restatement of the reference's local-basis stage in pure numpy / scipy.

It shares no code with oracle/msfem_oracle.c (no import, no call, different algorithms where the
mathematics allows): own first-touch DoF numbering, own element assembly straight from the integrand
of assemble_system, a sparse DIRECT solve of the condensed system instead of SSOR-PCG, dense
M = Phi^T K Phi.  Both the C oracle and the CUDA path must reproduce its vectors
(tests/test_independent_restatement.py, tests/test_gpu_parity.py::test_independent_restatement_on_gpu),
so a wrong reading of the reference would have to be made twice, independently, to go unnoticed.

Reference lines restated (read, never copied):
  local mesh / DoFs   /root/reference/include/base/diffusion_problem_basis.tpp:90-99, 102-117
                      (general_cell + refine_global -> cells in Morton order, FE_Q(1) first-touch numbering:
                       SURVEY.md Appendix A.1-A.2 for the deal.II part)
  constraints         basis.tpp:119-135; Coefficients::BasisQ1<2>, include/coefficients/basis_q1.tpp:26-47, 116-133
  assemble_system     basis.tpp:159-242: QGauss<2>(2); cell_matrix(i,j) += grad_i * A(q) * grad_j * JxW,
                      cell_rhs(i) += phi_i * f(q) * JxW; f = 2 (right_hand_side.tpp:27-40)
  coefficient         Coefficients::MatrixCoeff<2>, matrix_coeff.tpp:17-25, 66-91; constants matrix_coeff.hpp:45-48;
                      PI_D (sic) coefficients.h:21
  condensed solve     basis.tpp:450-465 + 293-317 (the converged solution does not depend on the solver)
  element matrix      basis.tpp:245-285

Usage:  python tests/golden/independent_restatement.py          (re)writes independent_golden.json
This is test infrastructure: nothing in the product path imports it.
"""
import json
import math
import os

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

PI_D = 3.14592653509793218403  # coefficients.h:21 -- not pi; reproduced verbatim
M64 = (1 << 64) - 1


# ----------------------------------------------------------------------------- coefficients
def _mix64(z):
    z &= M64
    z ^= z >> 33
    z = (z * 0xff51afd7ed558ccd) & M64
    z ^= z >> 33
    z = (z * 0xc4ceb9fe1a85ec53) & M64
    z ^= z >> 33
    return z


def coefficient(kind, par, seed):
    """Returns A(x, y) -> 2x2 array.  kinds as in include/msfem_basis.h / BASELINE.md section 4."""
    if kind == 0:  # MatrixCoeff<2>
        al = PI_D / 3
        rot = np.array([[math.cos(al), math.sin(al)], [-math.sin(al), math.cos(al)]])

        def A(x, y):
            a = 1.0 * (1.0 - 0.9999 * (0.5 * math.sin(2 * PI_D * 57 * x) + 0.5 * math.sin(2 * PI_D * 57 * y)))
            return rot @ (a * np.eye(2)) @ rot.T
        return A
    if kind == 1:  # periodic, true pi
        eps, scale = par

        def A(x, y):
            a = 1.0 - scale * (0.5 * math.sin(2 * math.pi * x / eps) + 0.5 * math.sin(2 * math.pi * y / eps))
            return a * np.eye(2)
        return A
    if kind == 2:  # random inclusions: counter-based hash of the block index
        block, prob, a_in, a_bg = par

        def A(x, y):
            bx, by = int(math.floor(x / block)), int(math.floor(y / block))
            h = (bx * 0x9E3779B97F4A7C15) & M64
            h ^= _mix64((by + 0xC2B2AE3D27D4EB4F * (seed & 0xffffffff)) & M64)
            h = _mix64(h)
            inside = (h >> 11) * (1.0 / 9007199254740992.0) < prob
            return (a_in if inside else a_bg) * np.eye(2)
        return A
    if kind == 3:
        return lambda x, y: par[0] * np.eye(2)
    raise ValueError(kind)


# ----------------------------------------------------------------------------- mesh and DoFs
def morton_cells(l):
    """Fine cells (ix, iy) in the order refine_global leaves them: children c = ix_bit + 2 iy_bit, recursively."""
    cells = [(0, 0)]
    for _ in range(l):
        cells = [(2 * ix + (c & 1), 2 * iy + (c >> 1)) for (ix, iy) in cells for c in range(4)]
    return cells


def dof_numbering(l):
    """FE_Q(1) on a serial Triangulation: vertices numbered in the order cells (Morton) and their local
    vertices (x fastest) first touch them.  Returns dof[jy, jx]."""
    n = 1 << l
    dof = -np.ones((n + 1, n + 1), dtype=np.int64)
    nxt = 0
    for ix, iy in morton_cells(l):
        for v in range(4):
            jx, jy = ix + (v & 1), iy + (v >> 1)
            if dof[jy, jx] < 0:
                dof[jy, jx] = nxt
                nxt += 1
    assert nxt == (n + 1) ** 2
    return dof


def vertex_position(corners, n, jx, jy):
    """Bilinear image of the uniform grid on the coarse cell (deal.II vertex order v0 v1 / v2 v3)."""
    s, t = jx / n, jy / n
    c = np.asarray(corners, dtype=np.float64).reshape(4, 2)
    return (1 - s) * (1 - t) * c[0] + s * (1 - t) * c[1] + (1 - s) * t * c[2] + s * t * c[3]


# ----------------------------------------------------------------------------- assembly
def assemble(l, corners, A, f=2.0):
    """K (N x N, unconstrained) and F in the DoF numbering of dof_numbering(l)."""
    n = 1 << l
    dof = dof_numbering(l)
    N = (n + 1) ** 2
    g = [0.5 - 0.5 / math.sqrt(3.0), 0.5 + 0.5 / math.sqrt(3.0)]
    rows, cols, vals = [], [], []
    F = np.zeros(N)
    for iy in range(n):
        for ix in range(n):
            P = np.array([vertex_position(corners, n, ix + (v & 1), iy + (v >> 1)) for v in range(4)])
            ids = [dof[iy + (v >> 1), ix + (v & 1)] for v in range(4)]
            Ke = np.zeros((4, 4))
            Fe = np.zeros(4)
            for qy in range(2):
                for qx in range(2):
                    xi, eta = g[qx], g[qy]
                    shp = np.array([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta])
                    dref = np.array([[-(1 - eta), -(1 - xi)], [(1 - eta), -xi], [-eta, (1 - xi)], [eta, xi]])
                    J = P.T @ dref                      # d(x,y)/d(xi,eta)
                    detJ = np.linalg.det(J)
                    grad = dref @ np.linalg.inv(J)      # rows: physical gradients of the shape functions
                    xq = shp @ P
                    Aq = A(xq[0], xq[1])
                    JxW = detJ * 0.25
                    Ke += grad @ Aq @ grad.T * JxW
                    Fe += shp * f * JxW
            for i in range(4):
                F[ids[i]] += Fe[i]
                for j in range(4):
                    rows.append(ids[i])
                    cols.append(ids[j])
                    vals.append(Ke[i, j])
    K = sp.csr_matrix((vals, (rows, cols)), shape=(N, N))
    return K, F, dof


def basis_q1_coefficients(corners):
    c = np.asarray(corners, dtype=np.float64).reshape(4, 2)
    pm = np.array([[1.0, x, y, x * y] for x, y in c])
    return np.linalg.inv(pm)      # columns = monomial coefficients of the 4 coarse shape functions


def run_cell(l, corners, kind, par, seed, f=2.0):
    n = 1 << l
    A = coefficient(kind, par, seed)
    K, F, dof = assemble(l, corners, A, f)
    N = (n + 1) ** 2
    cm = basis_q1_coefficients(corners)
    bnd = sorted(int(dof[jy, jx]) for jy in range(n + 1) for jx in range(n + 1)
                 if jx in (0, n) or jy in (0, n))
    pos = np.zeros((N, 2))
    for jy in range(n + 1):
        for jx in range(n + 1):
            pos[dof[jy, jx]] = vertex_position(corners, n, jx, jy)
    inner = np.setdiff1d(np.arange(N), np.array(bnd))
    Kcsc = K.tocsc()
    KII = Kcsc[inner][:, inner]
    KIB = Kcsc[inner][:, bnd]
    lu = spla.splu(KII.tocsc())
    phi = np.zeros((4, N))
    gvals = np.zeros((4, len(bnd)))
    for ib in range(4):
        gb = cm[0, ib] + cm[1, ib] * pos[bnd, 0] + cm[2, ib] * pos[bnd, 1] + cm[3, ib] * pos[bnd, 0] * pos[bnd, 1]
        gvals[ib] = gb
        phi[ib, bnd] = gb
        phi[ib, inner] = lu.solve(-(KIB @ gb))
    M = phi @ (K @ phi.T)
    b = phi @ F
    return dict(dof=dof, boundary_dofs=bnd, constraint_values=gvals, phi=phi, M=M, b=b, K=K, F=F)


CASES = {
    # name: (l, corners or (r, ix, iy), kind, par, seed, f)
    "ref_l3_cell_2_5": (3, (3, 2, 5), 0, (), 0, 2.0),
    "ref_l5_cell_0_0": (5, (3, 0, 0), 0, (), 0, 2.0),
    "periodic_l4": (4, (5, 17, 9), 1, (1.0 / 64, 0.9999), 0, 2.0),
    "periodic_l6_target_cell": (6, (8, 37, 101), 1, (1.0 / 64, 0.9999), 0, 2.0),
    "inclusions_l5": (5, (8, 200, 13), 2, (2.0 ** -11, 0.2, 1e4, 1.0), 1234, 2.0),
    "constant_l3_f35": (3, (2, 1, 2), 3, (2.5,), 0, 3.5),
    "ref_l4_rectangle": (4, [[0.50, 0.25], [0.75, 0.25], [0.50, 0.375], [0.75, 0.375]], 0, (), 0, 2.0),
    "ref_l4_skewed_quad": (4, [[0.10, 0.20], [0.35, 0.22], [0.12, 0.41], [0.38, 0.47]], 0, (), 0, 3.5),
}


def corners_of(spec):
    if isinstance(spec, tuple):
        r, ix, iy = spec
        H = 1.0 / (1 << r)
        return [[ix * H, iy * H], [(ix + 1) * H, iy * H], [ix * H, (iy + 1) * H], [(ix + 1) * H, (iy + 1) * H]]
    return spec


def main():
    out = {}
    for name, (l, spec, kind, par, seed, f) in CASES.items():
        cor = corners_of(spec)
        res = run_cell(l, cor, kind, par, seed, f)
        n = 1 << l
        dof = res["dof"]
        probes = [(n // 2, n // 2), (1, 1), (n - 1, 2), (n // 3, n - 1), (n, n // 2), (0, 0)]
        out[name] = {
            "l": l, "corners": cor, "kind": kind, "par": list(par), "seed": seed, "f": f,
            "dof_corners": [int(dof[0, 0]), int(dof[0, n]), int(dof[n, 0]), int(dof[n, n])],
            "dof_checksum": int(sum((i + 1) * int(d) for i, d in enumerate(dof.ravel())) % (1 << 61)),
            "boundary_dofs_checksum": int(sum((i + 1) * d for i, d in enumerate(res["boundary_dofs"])) % (1 << 61)),
            "n_boundary": len(res["boundary_dofs"]),
            "constraint_values_head": res["constraint_values"][:, :6].tolist(),
            "M": res["M"].tolist(), "b": res["b"].tolist(),
            "probes": probes,
            "phi_probes": [[float(res["phi"][ib, dof[jy, jx]]) for ib in range(4)] for (jx, jy) in probes],
            "phi_norms": [float(np.linalg.norm(res["phi"][ib])) for ib in range(4)],
        }
        print(name, "M00 %.15g  |phi0| %.12g" % (res["M"][0, 0], out[name]["phi_norms"][0]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "independent_golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
