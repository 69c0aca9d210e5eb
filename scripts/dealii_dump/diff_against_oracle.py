#!/usr/bin/env python
"""Diff a deal.II dump of the reference's basis objects against this repository's restatements.

  python scripts/dealii_dump/diff_against_oracle.py DUMP_DIR [--gpu] [--independent]

DUMP_DIR holds the basis_dump.cell-<CellId>.txt files the patched reference wrote (README.md in this
directory).  For every dumped coarse cell this script compares, against oracle/msfem_oracle.c (and with --gpu
against the CUDA path through the C ABI, with --independent against tests/golden/independent_restatement.py):

  DoF map               bit for bit   (vertex position -> DoF index, basis.tpp:106)
  constraint index sets bit for bit   (basis.tpp:119-135), values to 1e-9
  load vector F         1e-12 rel     (basis.tpp:217-221)
  bases phi_i           1e-8 rel l2   (basis.tpp:293-317)
  M, b                  1e-8 rel      (basis.tpp:245-285)

Exit code 0 = the restatements match the reference on every dumped cell: the day this has run green on a
deal.II machine, `parity` of the oracle is pinned.  Nothing here is on the product path.
"""
import argparse
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def parse_dump(path):
    head = None
    corners = None
    dof_rows, cons, M, b = [], {}, {}, {}
    for line in open(path):
        w = line.split()
        if not w:
            continue
        if w[0] == "cell":
            head = {"id": w[1], "dim": int(w[3]), "l": int(w[5]), "n_dofs": int(w[7])}
        elif w[0] == "corners":
            corners = np.array([float(v) for v in w[1:]])
        elif w[0] == "dof":
            dim = head["dim"]
            i = int(w[1])
            xyz = [float(v) for v in w[2:2 + dim]]
            assert w[2 + dim] == "F"
            F = float(w[3 + dim])
            phi = [float(v) for v in w[4 + dim:]]
            dof_rows.append((i, xyz, F, phi))
        elif w[0] == "constraint":
            cons.setdefault(int(w[1]), []).append((int(w[2]), float(w[3])))
        elif w[0] == "M":
            M[(int(w[1]), int(w[2]))] = float(w[3])
        elif w[0] == "b":
            b[int(w[1])] = float(w[2])
    dim, nb = head["dim"], 1 << head["dim"]
    N = head["n_dofs"]
    pos = np.zeros((N, dim))
    F = np.zeros(N)
    phi = np.zeros((nb, N))
    for i, xyz, f, p in dof_rows:
        pos[i], F[i], phi[:, i] = xyz, f, p
    Mm = np.array([[M[(i, j)] for j in range(nb)] for i in range(nb)])
    bb = np.array([b[i] for i in range(nb)])
    return head, corners.reshape(nb, dim), pos, F, phi, cons, Mm, bb


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dump_dir")
    ap.add_argument("--gpu", action="store_true", help="also compare the CUDA path (needs a B200)")
    ap.add_argument("--independent", action="store_true", help="also compare the numpy/scipy restatement (2D, slow for l >= 6)")
    ap.add_argument("--coeff-kind", type=int, default=0, help="msb_coeff_kind the reference ran with (0 = MatrixCoeff)")
    args = ap.parse_args()
    from oracle import oracle as O
    O.build()
    files = sorted(glob.glob(os.path.join(args.dump_dir, "basis_dump.cell-*.txt")))
    if not files:
        raise SystemExit("no basis_dump.cell-*.txt in " + args.dump_dir)
    bad = 0
    for path in files:
        head, corners, pos, F, phi, cons, M, b = parse_dump(path)
        dim, l, nb = head["dim"], head["l"], 1 << head["dim"]
        n = 1 << l
        # ---- DoF map: vertex (jx, jy[, jz]) of every DoF from its support point.  The coarse cells of the
        # reference are axis-aligned (refined hyper_cube, ms.tpp:97-99); general cells would need the inverse
        # of the multilinear map here.
        lo, hi = corners.min(axis=0), corners.max(axis=0)
        idx = np.rint((pos - lo) / (hi - lo) * n).astype(np.int64)
        dmap = (O.dof_map(l) if dim == 2 else O.dof_map3(l))
        want = dmap[idx[:, 1], idx[:, 0]] if dim == 2 else dmap[idx[:, 2], idx[:, 1], idx[:, 0]]
        ok_map = bool(np.array_equal(want, np.arange(head["n_dofs"])))
        # ---- constraint sets
        bd = O.boundary_dofs(l) if dim == 2 else O.boundary_dofs3(l)
        ok_cons, worst_val = True, 0.0
        for k in range(nb):
            got = sorted(cons.get(k, []))
            ok_cons &= [g[0] for g in got] == [int(v) for v in bd]
            vals = O.constraint_values(l, corners, k) if dim == 2 else O.constraint_values3(l, corners, k)
            worst_val = max(worst_val, float(np.abs(np.array([g[1] for g in got]) - vals).max()))
        # ---- floating point against the oracle
        co = O.coeff(args.coeff_kind)
        run = O.run_cells if dim == 2 else O.run_cells3
        ref = run(l, corners[None], co)
        asm = O.assemble(l, corners, co) if dim == 2 else O.assemble3(l, corners, co)
        e_F = rel(asm[3], F)
        e_phi = max(rel(ref["phi"][0][k], phi[k]) for k in range(nb))
        e_M, e_b = rel(ref["M"][0], M), rel(ref["b"][0], b)
        line = ("%s: dof map %s, constraint sets %s (values %.1e), F %.1e, phi %.1e, M %.1e, b %.1e"
                % (head["id"], "EXACT" if ok_map else "DIFFERS", "EXACT" if ok_cons else "DIFFER", worst_val, e_F,
                   e_phi, e_M, e_b))
        good = ok_map and ok_cons and worst_val < 1e-9 and e_F < 1e-12 and e_phi < 1e-8 and e_M < 1e-8 and e_b < 1e-8
        if args.independent and dim == 2:
            sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
            import independent_restatement as IR
            res = IR.run_cell(l, corners.tolist(), args.coeff_kind, (), 0)
            e2 = max(rel(res["phi"][k], phi[k]) for k in range(nb))
            line += " | numpy restatement: phi %.1e, M %.1e" % (e2, rel(res["M"], M))
            good &= e2 < 1e-8 and rel(res["M"], M) < 1e-8
        if args.gpu:
            import mpi_parallel_multiscale_diffusion_fem_b200 as pkg
            from mpi_parallel_multiscale_diffusion_fem_b200.binding import coeff_desc
            with pkg.BasisShard(l, corners[None], coeff_desc(args.coeff_kind), dim=dim) as sh:
                sh.run(1e-12, 5000)
                Mg, bg = sh.element_matrices()
                pg = sh.bases()[0]
                ok_gmap = bool(np.array_equal(sh.dof_map(), dmap))
            e3 = max(rel(pg[k], phi[k]) for k in range(nb))
            line += " | CUDA: dof map %s, phi %.1e, M %.1e, b %.1e" % ("EXACT" if ok_gmap else "DIFFERS", e3,
                                                                       rel(Mg[0], M), rel(bg[0], b))
            good &= ok_gmap and e3 < 1e-8 and rel(Mg[0], M) < 1e-8 and rel(bg[0], b) < 1e-8
        print(("OK   " if good else "FAIL ") + line)
        bad += not good
    print("%d of %d dumped cells match" % (len(files) - bad, len(files)))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
