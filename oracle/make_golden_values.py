#!/usr/bin/env python
"""Freeze the LIVE reference's field evaluation into tests/golden/values_*.npz
(SURVEY.md 8(f) row 4): ``DOFManager.values_at_nodes`` (sem/discrete.py:235-258,
GLL coefficients -> values at the equispaced mesh nodes through
``TensorProduct.interpolate_on_grid_eq``, sem/basis_functions.py:539-569) and
``DOFManager.interpolate`` at a few physical points (sem/discrete.py:221-233).

TEST INFRASTRUCTURE; development container only (needs /root/reference).

    python oracle/make_golden_values.py
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
    ("C534_dm", "C", 5, 3, 4, False, False),
    ("C448_sc_rcm", "C", 4, 4, 8, True, True),
    ("C3310_dm_rcm", "C", 3, 3, 10, False, True),
    ("S324_sc", "S", 3, 2, 4, True, False),
]
POINTS = np.array([[0.1, -0.3], [-0.85, 0.9], [0.55, 0.55], [-0.999, -0.999]])


def main():
    lr.install_shims()
    for name, kind, nx, ny, p, sc, rcm in CASES:
        mesh = lr.build_mesh(kind, nx, ny, p)
        mngr = lr.make_manager(mesh, p, sc, rcm)
        x, y = mesh.nodes
        # two fields (leading axis), deterministic, not in the polynomial space
        coeffs = np.stack([np.sin(3 * x) * np.cos(2 * y), np.exp(0.5 * x - 0.25 * y)])
        values = mngr.values_at_nodes(coeffs)
        mesh._compute_cell_centroids()
        pts = np.array([mngr.interpolate(coeffs, pt) for pt in POINTS])
        np.savez_compressed(os.path.join(OUT, "values_%s.npz" % name), coeffs=coeffs,
                            values=values, points=POINTS, point_values=pts,
                            meta=np.array([nx, ny, p, int(sc), int(rcm), ord(kind)]))
        print(name, values.shape, "||values|| = %.15g" % np.linalg.norm(values), pts.ravel()[:3])


if __name__ == "__main__":
    main()
