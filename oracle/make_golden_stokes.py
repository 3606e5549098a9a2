#!/usr/bin/env python
"""Freeze what the reference's OWN squirmer example computes into tests/golden/stokes_*.npz
(SURVEY.md 8(f) row 3 / BASELINE config 4).

The class of examples/squirmer-axisymmetric.py is executed live (oracle/live_squirmer.py:
name aliases only, the arithmetic is the example's) on small structured annulus-sector
meshes standing in for examples/meshes/donut.geo (no .msh ships, gmsh is absent):
`pre_assembly` (:163-257: BCs + the local operators E2e, Lve, Ae, Me), `compute_local_system`
(:259-297) at a perturbed state, and the Newton loop on the Schur complement (`solve`,
:389-457) to convergence.

TEST INFRASTRUCTURE; development container only (needs /root/reference).

    python oracle/make_golden_stokes.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle import live_squirmer as ls  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")

CASES = [
    # name, kind, nr, nt, p, r_out, n_rey, kwargs of pre_assembly, elements whose dense operators are kept
    ("fixed_344_re0", "FixedSphere", 3, 4, 4, 10.0, 0.0, {}, None),
    ("fixed_225_re1", "FixedSphere", 2, 2, 5, 6.0, 1.0, {}, None),
    ("squirmer_238_re0", "Squirmer", 2, 3, 8, 100.0, 0.0, dict(speed=0.8, beta=1.5), (0, 5)),
    ("squirmer_334_re05", "Squirmer", 3, 3, 4, 8.0, 0.5, dict(speed=1.0, beta=-2.0), (0, 4, 8)),
]


def main():
    for name, kind, nr, nt, p, r_out, n_rey, kw, keep in CASES:
        mod, mesh, prob = ls.make_problem(nr, nt, p, r_out=r_out, kind=kind)
        dm = prob.dof_mngr
        prob.set_initial_guess()
        prob.pre_assembly(n_rey, **kw)
        fes = list(dm.finite_elements(x_phys=True, Jacobian=True))
        l2g = np.array([fe.node_ind for fe in fes], dtype=np.uint32)
        n_elem = len(fes)
        keep_idx = np.arange(n_elem) if keep is None else np.array(keep)
        N = p + 1
        pp, qq = np.ogrid[0:N, 0:N]
        E2e = np.array([prob.operators[e][0] for e in keep_idx])
        Lve = np.array([prob.operators[e][1] for e in keep_idx])
        Me = np.array([prob.operators[e][3].to_array()[pp, qq, pp, qq] for e in keep_idx])
        state0 = prob.soln_vec.copy()              # potential-flow guess + essential BC values
        # local systems at a perturbed state (exercises the advection terms when n_rey != 0)
        rng = np.random.default_rng(7)
        pert = state0 + 0.05 * rng.standard_normal(state0.size) * (1.0 + np.abs(state0))
        prob.soln_vec[:] = pert
        sys_l = [prob.compute_local_system(fes[e], prob.operators[e]) for e in keep_idx]
        jac = np.array([s[0] for s in sys_l])
        rhs = np.array([s[1] for s in sys_l])
        # all elements: residual only (small), for the assembled-residual check
        rhs_all = np.array([prob.compute_local_system(fe, o)[1]
                            for fe, o in zip(fes, prob.operators)])
        # Newton loop from the unperturbed guess, recording every increment norm
        prob.soln_vec[:] = state0
        prob.solve(it_max=20, tol=1e-10)
        out = dict(
            kind=kind, nr=nr, nt=nt, p=p, r_out=r_out, n_rey=n_rey,
            speed=kw.get("speed", 1.0), beta=kw.get("beta", np.nan),
            nodes=mesh.nodes.copy(), l2g=l2g,
            ndof_exterior=dm.ndof_exterior,
            dof_mask=prob.dof_mask.copy(), cint=prob.cint.copy(),
            state0=state0, perturbed=pert, keep=keep_idx,
            E2e=E2e, Lve=Lve, Me=Me, jac=jac, rhs=rhs, rhs_all=rhs_all,
            solution=prob.soln_vec.copy())
        path = os.path.join(OUT, "stokes_%s.npz" % name)
        np.savez_compressed(path, **out)
        print(name, "ndof", dm.ndof, "file %.0f KB" % (os.path.getsize(path) / 1024.0),
              "|soln|", np.linalg.norm(prob.soln_vec))


if __name__ == "__main__":
    main()
