#!/usr/bin/env python
"""Freeze outputs of the LIVE reference into tests/golden/*.npz.

TEST INFRASTRUCTURE.  Run in the development container only (needs
/root/reference); the fixtures it writes are committed so that the GPU box,
where the reference tree does not exist, can still check parity against the
reference's own numbers.

    python oracle/make_golden.py

Fixtures
  tables.npz        GLL nodes / barycentric / quadrature weights, D1, E for
                    orders 1..10 from sem.basis_functions.LagrangeGaussLobatto;
                    sem.quadratures.GaussLobatto(n) for n = 1..12; hierarchical
                    node orders; _subface_slice outputs.
  case_*.npz        one Poisson pipeline per (mesh, order, manager, rcm):
                    L2G map, permuted nodes, hierarchical DOF ids, fe.invJ,
                    fe.detJxW, fe.x_phys, u, A u, b, diag(A), on_ebc, Dirichlet
                    values and the reference's solution (Schur path for
                    DOFManagerSC, full assembled spsolve for DOFManager).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import live_reference as lr  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")

CASES = [
    # name, kind, nx, ny, p, sc, rcm
    ("S448_sc", "S", 4, 4, 8, True, False),
    ("S448_sc_rcm", "S", 4, 4, 8, True, True),
    ("C448_sc_rcm", "C", 4, 4, 8, True, True),
    ("C534_dm", "C", 5, 3, 4, False, False),
    ("C3310_dm_rcm", "C", 3, 3, 10, False, True),
    ("S324_dm", "S", 3, 2, 4, False, False),
    ("S324_dm_rcm", "S", 3, 2, 4, False, True),
    ("S324_sc", "S", 3, 2, 4, True, False),
    ("S324_sc_rcm", "S", 3, 2, 4, True, True),
    ("C552_sc_rcm", "C", 5, 5, 2, True, True),
    ("S888_sc_rcm", "S", 8, 8, 8, True, True),     # config-1 substitute
    ("C888_sc_rcm", "C", 8, 8, 8, True, True),
]


def tables():
    lr.install_shims()
    from sem.basis_functions import LagrangeGaussLobatto
    from sem.geometry import Quadrilateral
    from sem.mapping import _subface_slice
    from sem.quadratures import GaussLobatto
    out = {}
    for p in range(1, 11):
        b = LagrangeGaussLobatto(p)
        out["nodes_%d" % p] = b.nodes
        out["bary_%d" % p] = b.bary_wts
        out["quad_%d" % p] = b.quad_rule.weights
        out["D_%d" % p] = b.D1
        out["E_%d" % p] = b._interp_eq_mat
    for n in range(1, 13):
        g = GaussLobatto(n)
        out["gl_x_%d" % n] = g.abscissa
        out["gl_w_%d" % n] = g.weights
    for N in (2, 3, 5, 9, 11):
        out["hier_%d" % N] = Quadrilateral(N, N).hierarchical_node_order
    arr = np.arange(2 * 4 * 5).reshape(2, 4, 5)
    for f in range(4):
        out["face_%d" % f] = np.ascontiguousarray(_subface_slice(f, arr, 2))
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    tables()
    for name, kind, nx, ny, p, sc, rcm in CASES:
        r = lr.run_case(kind, nx, ny, p, sc, rcm)
        keep = {k: r[k] for k in ("l2g", "hier", "nodes", "invJ", "JxW", "x_phys", "u", "Au", "b",
                                  "diag", "on_ebc", "ebc_vals", "solution")}
        keep["meta"] = np.array([nx, ny, p, int(sc), int(rcm), ord(kind)])
        np.savez_compressed(os.path.join(OUT, "case_%s.npz" % name), **keep)
        print(name, r["l2g"].shape, "||Au|| = %.15g" % np.linalg.norm(r["Au"]),
              "||sol|| = %.15g" % np.linalg.norm(r["solution"]))


if __name__ == "__main__":
    main()
