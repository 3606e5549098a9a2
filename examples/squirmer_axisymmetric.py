#!/usr/bin/env python
"""Axisymmetric flow past a sphere / a squirmer -- the reference's example on the B200 engine.

Mirrors examples/squirmer-axisymmetric.py of the reference (stream function - vorticity
formulation, two DOFs per node): `FixedSphere` and `Squirmer` with the same boundary
conditions (:119-161), initial guess (:109-117) and Newton iteration (:389-457).  The dense
per-element operators (:163-257), local Jacobians (:259-297) and the Schur-complement /
SuperLU step (:299-387) are replaced by the matrix-free device Jacobian
(csrc/semk_stokes.cu) and restarted GMRES with a nodal 2x2 block-Jacobi preconditioner.

    python examples/squirmer_axisymmetric.py [--nr 4] [--nt 6] [--order 8] [--r-out 100]
                                             [--re 0.0] [--beta 1.0] [--speed 1.0] [--fixed]
                                             [--preconditioner poisson|block-jacobi]

``--preconditioner poisson`` (default): flexible GMRES with the block-triangular
preconditioner built from the weighted, statically condensed Poisson operator (mesh-
independent: 10 M DOF in ~15 s per linear solve); ``block-jacobi``: the nodal 2 x 2 blocks
(small meshes only).

No .msh ships with the reference and gmsh is not available, so the mesh is the structured
annulus sector of meshgen.annulus_sector_mesh (the transfinite mesh of
examples/meshes/donut.geo: unit sphere, shell at r_out, symmetry axis).
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from spectralelementmethod_b200 import discrete, meshgen, stokes  # noqa: E402
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS  # noqa: E402


def run(nr=4, nt=6, order=8, r_out=100.0, n_rey=0.0, beta=1.0, speed=1.0, fixed=False,
        quiet=False, tol=1e-6, restart=400, preconditioner="poisson"):
    mesh = meshgen.annulus_sector_mesh(nr, nt, order, r_out)
    b1 = LagrangeGaussLobatto(order)
    # (rcm_order=False: the scalar manager of the Poisson preconditioner shares the numbering)
    dm = discrete.DOFManagerSC(mesh, 2, TensorProductQS(b1, b1), rcm_order=False)
    slip = stokes.zero_slip_vel if fixed else stokes.squirmer_vslip_profile(beta)
    bc = stokes.squirmer_boundary_data(dm, 1.0 if fixed else speed, slip)
    op = dm.axisymmetric_stokes_operator(n_rey=n_rey, essential=bc.essential)
    t0 = time.perf_counter()
    poisson = preconditioner == "poisson"
    state, hist = op.newton_solve(op.from_host(bc.state0), bc.cint, tol=tol, restart=restart,
                                  gmres_rtol=1e-10 if poisson else 1e-12,
                                  gmres_maxiter=(3 if poisson else 20) * restart,
                                  verbose=not quiet, precondition="poisson" if poisson else True)
    el = time.perf_counter() - t0
    soln = state.cpu().numpy()
    if not quiet:
        print(" => %d DOF, %d Newton iterations, %d GMRES iterations, %.2f s"
              % (op.n_dof, len(hist), sum(i.iterations for _, i in hist), el))
        print("    |sfn| = %.12g, |vort| = %.12g" % (np.linalg.norm(soln[0::2]),
                                                     np.linalg.norm(soln[1::2])))
    return soln, hist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nr", type=int, default=4)
    ap.add_argument("--nt", type=int, default=6)
    ap.add_argument("--order", type=int, default=8)
    ap.add_argument("--r-out", type=float, default=100.0)
    ap.add_argument("--re", type=float, default=0.0)
    ap.add_argument("--beta", type=float, default=1.0)
    ap.add_argument("--speed", type=float, default=1.0)
    ap.add_argument("--fixed", action="store_true", help="fixed sphere in uniform flow (no slip)")
    ap.add_argument("--preconditioner", default="poisson", choices=["poisson", "block-jacobi"])
    ap.add_argument("--swim", action="store_true",
                    help="find the swimming speed (zero net force) by the secant iteration of "
                         "the reference's calc_speed; its docstring quotes 0.92571156681483957 "
                         "for Re = 1, beta = 1 on meshes/donut.msh (15 x 9 elements, r_out = 100)")
    a = ap.parse_args()
    if a.swim:
        mesh = meshgen.annulus_sector_mesh(a.nr, a.nt, a.order, a.r_out)
        b1 = LagrangeGaussLobatto(a.order)
        dm = discrete.DOFManagerSC(mesh, 2, TensorProductQS(b1, b1), rcm_order=False)
        t0 = time.perf_counter()
        speed, _, hist = stokes.squirmer_speed(dm, a.re, a.beta, verbose=True, restart=400,
                                               gmres_rtol=1e-10, gmres_maxiter=1200)
        print(" => swimming speed %.12g at Re = %g, beta = %g (%d flow solves, %.1f s)"
              % (speed, a.re, a.beta, len(hist), time.perf_counter() - t0))
        return
    run(a.nr, a.nt, a.order, a.r_out, a.re, a.beta, a.speed, a.fixed,
        preconditioner=a.preconditioner)


if __name__ == "__main__":
    main()
