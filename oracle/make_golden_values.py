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


def many_points(seed, n=96):
    """Seeded points for the batched point location: uniform in the domain, plus points on
    the domain boundary and its corners (element-boundary hits inside the mesh are covered
    by the uniform ones landing within 1e-8 of an edge only by luck, so a few points are
    snapped onto coordinate lines that are element edges of the straight meshes)."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-1.0, 1.0, size=(n, 2))
    pts[:8, 0] = [-1.0, 1.0, -1.0, 1.0, -1.0, 1.0, 0.0, 0.0]
    pts[:4, 1] = [-1.0, -1.0, 1.0, 1.0]
    pts[8:12, 0] = 0.0
    return pts


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
        # batched location: cell number, parametric coordinates and values, point by point
        # through find_elem_containing_point / Mapping.inv / interpolate of the live reference
        from sem.rootfind import SolverFailure
        from sem.discrete import OutsideDomain
        first_node = {int(mesh.get_cell(i).node_ind_lexicographic[0, 0]) * 1000003
                      + int(mesh.get_cell(i).node_ind_lexicographic[-1, -1]): i
                      for i in range(mesh.n_cells)}
        kept, cells, xpar, vals = [], [], [], []
        for pt in many_points(len(name)):
            try:
                fe, xp = mngr.find_elem_containing_point(pt)
            except (SolverFailure, OutsideDomain):
                # the reference gives up when Newton does not converge in a cell that does
                # not contain the point (sem/rootfind.py:52-53 is not caught at
                # sem/discrete.py:274-278), and rejects boundary points whose parametric
                # coordinate rounds to 1 + 2e-16; such points have no reference answer
                continue
            kept.append(pt)
            cells.append(first_node[int(fe.node_ind[0, 0]) * 1000003 + int(fe.node_ind[-1, -1])])
            xpar.append(np.array(xp, dtype=float))
            vals.append(fe.interpolate(coeffs[..., fe.node_ind], xp))
        mp = np.array(kept)
        np.savez_compressed(os.path.join(OUT, "values_%s.npz" % name), coeffs=coeffs,
                            values=values, points=POINTS, point_values=pts,
                            many_points=mp, many_cells=np.array(cells),
                            many_xparam=np.array(xpar), many_values=np.array(vals),
                            meta=np.array([nx, ny, p, int(sc), int(rcm), ord(kind)]))
        print(name, values.shape, "||values|| = %.15g" % np.linalg.norm(values), pts.ravel()[:3])


if __name__ == "__main__":
    main()
