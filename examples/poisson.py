#!/usr/bin/env python
"""Poisson's equation on a plate -- the reference's example on the B200 engine.

Mirrors the flow of the reference's examples/poisson.py (mesh -> essential /
natural boundary conditions -> operators -> solve) through the drop-in API of
this package; the per-element einsum loop (examples/poisson.py:145-200) and
the sparse direct solve (:203-259, sem/discrete.py:502-528) are replaced by
the matrix-free device operator and the device-resident Jacobi-PCG.

    -lap(u) = 1 on [-1, 1]^2,  u = 0.2 ((x + 1) + (y + 1)) on "ebc" (left + bottom),
    du/dn = 0 on "nbc" (right + top)              (examples/poisson.py:125-143, :200)

    python examples/poisson.py [--n 64] [--order 8] [--msh file.msh] [--kind S|C]
                               [--solver matrix-free|condensed|condensed-two-level]

``--solver condensed`` follows the reference example literally: static
condensation of the element interiors (DOFManagerSC, sem/discrete.py:404-528),
solve on the element-exterior DOFs, interior back-substitution -- on the device;
``condensed-two-level`` adds the vertex coarse space to the Jacobi preconditioner.

With ``--msh`` the mesh is read from a Gmsh 2.2 binary file with physical
names "ebc", "nbc" (lines) and a surface (the reference's examples/meshes/
square.geo); otherwise a structured n x n mesh is built (and, with
``--write-msh``, written out and read back through the importer).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from spectralelementmethod_b200 import discrete, grid_importers, meshgen  # noqa: E402
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS  # noqa: E402


def run(n=16, order=8, msh=None, kind="S", write_msh=None, rtol=1e-12, quiet=False,
        solver="matrix-free"):
    import torch
    t0 = time.perf_counter()
    if msh is None and write_msh is not None:
        msh = meshgen.write_gmsh22_binary(write_msh, n, n, order, kind)
    if msh is not None:
        mesh = grid_importers.load_msh(msh, 2)
    else:
        mesh = meshgen.structured_quad_mesh(n, n, order, kind)
    b1 = LagrangeGaussLobatto(order)
    mngr = discrete.DOFManagerSC(mesh, 1, TensorProductQS(b1, b1), rcm_order=False)

    # essential boundary values, boundary element by boundary element, like the reference
    soln = np.zeros(mngr.ndof)
    on_ebc = np.zeros(mngr.ndof, dtype=bool)
    for _fe, bfe in mngr.boundary_elements("ebc", x_phys=True):
        loc = bfe.node_ind
        x, y = bfe.x_phys
        soln[loc] = 0.2 * ((x + 1) + (y + 1))
        on_ebc[loc] = True
    t_setup = time.perf_counter() - t0

    t0 = time.perf_counter()
    extra = {}
    if solver in ("condensed", "condensed-two-level"):
        op = mngr.condensed_poisson_operator(dirichlet=on_ebc)
        if solver == "condensed-two-level":
            extra["preconditioner"] = "two-level"
    elif solver == "matrix-free":
        op = mngr.poisson_operator(dirichlet=on_ebc)
    else:
        raise ValueError("solver must be 'matrix-free', 'condensed' or 'condensed-two-level'")
    u, info = op.solve(f=1.0, dirichlet_values=torch.from_numpy(soln).to(op.dev), rtol=rtol,
                       **extra)
    torch.cuda.synchronize()
    t_solve = time.perf_counter() - t0
    u = u.cpu().numpy()
    if not quiet:
        print("mesh: %d cells of order %d, %d DOF (%d on the essential boundary)"
              % (mesh.n_cells, order, mngr.ndof, int(on_ebc.sum())))
        print("host set-up %.2f s; %s operator + PCG %.2f s: %d iterations, relative residual %.2e"
              % (t_setup, solver, t_solve, info.iterations, info.rel_residual))
        print("u: min %.6f  max %.6f  mean %.6f" % (u.min(), u.max(), u.mean()))
    return mngr, on_ebc, soln, u, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=64, help="cells per side of the structured mesh")
    ap.add_argument("--order", type=int, default=8)
    ap.add_argument("--kind", default="S", choices=["S", "C"], help="straight or curved cells")
    ap.add_argument("--msh", default=None, help="Gmsh 2.2 binary mesh to read instead")
    ap.add_argument("--write-msh", default=None,
                    help="write the structured mesh to this .msh file and read it back")
    ap.add_argument("--solver", default="matrix-free", choices=["matrix-free", "condensed", "condensed-two-level"])
    args = ap.parse_args()
    run(args.n, args.order, args.msh, args.kind, args.write_msh, solver=args.solver)


if __name__ == "__main__":
    main()
