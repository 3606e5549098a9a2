"""Config 4 (BASELINE.json): axisymmetric Stokes system scaled to ~1e7 DOF, apply throughput
of the two-field matrix-free Jacobian and a bounded GMRES run.  python tests/stokes_bench.py
[nr nt p n_rey reps gmres_iters]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from spectralelementmethod_b200 import discrete, meshgen, stokes  # noqa: E402
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS  # noqa: E402


def main():
    a = sys.argv[1:]
    nr, nt, p = (int(a[0]), int(a[1]), int(a[2])) if len(a) >= 3 else (224, 352, 8)
    n_rey = float(a[3]) if len(a) > 3 else 0.0
    reps = int(a[4]) if len(a) > 4 else 50
    gm = int(a[5]) if len(a) > 5 else 0
    pe = int(a[6]) if len(a) > 6 else None
    t0 = time.perf_counter()
    mesh = meshgen.annulus_sector_mesh(nr, nt, p, 100.0)
    b1 = LagrangeGaussLobatto(p)
    prec = a[7] if len(a) > 7 else "block-jacobi"
    mgr = discrete.DOFManagerSC if prec == "poisson" else discrete.DOFManager
    dm = mgr(mesh, 2, TensorProductQS(b1, b1), rcm_order=False)
    op = dm.axisymmetric_stokes_operator(n_rey=n_rey, elems_per_patch=pe)
    torch.cuda.synchronize()
    setup = time.perf_counter() - t0
    x = torch.from_numpy(np.sin(np.arange(op.n_dof) * 1e-3)).cuda()
    if op.advection:
        op.linearize(x)
    y = op.new_vector()
    for _ in range(5):
        op.apply_unmasked(x, out=y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        op.apply_unmasked(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    peak = 6538.6
    try:
        peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_GBps", peak)
    except Exception:
        pass
    alg = op.algorithmic_bytes_per_apply
    out = {"workload": "annulus sector %dx%d elements p=%d, 2 DOF/node" % (nr, nt, p),
           "n_rey": n_rey, "dof": op.n_dof, "elems_per_patch": op.elems_per_patch,
           "smem_bytes_per_cta": op.smem_bytes, "n_fac": op.n_fac, "setup_seconds": setup,
           "ms_per_apply": ms, "gdof_per_s": op.n_dof / ms / 1e6,
           "algorithmic_bytes": alg, "achieved_GBps": alg / ms / 1e6,
           "frac_of_hbm_peak": alg / ms / 1e6 / peak, "checksum": float(y.sum().item())}
    if gm:
        bc = stokes.squirmer_boundary_data(dm, 1.0, stokes.squirmer_vslip_profile(1.0),
                                           x_phys=op.x_phys.cpu().numpy().reshape(op.n_elem, 2, p + 1, p + 1))
        op.set_essential(bc.essential)
        s0 = op.from_host(bc.state0)
        rhs = op.from_host(bc.cint) - op.residual(s0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if prec == "poisson" and len(a) > 8:
            op._poisson_prec = stokes.PoissonBlockPreconditioner(
                op, rtol=float(a[8]), reaction_term=(len(a) <= 9 or a[9] != "K"))
        d, info = op.solve_gmres(rhs, rtol=1e-8, restart=min(gm, 300), maxiter=gm,
                                 precondition="poisson" if prec == "poisson" else True)
        torch.cuda.synchronize()
        out["gmres"] = {"seconds": time.perf_counter() - t0, "iterations": info.iterations,
                        "restarts": info.restarts, "converged": info.converged,
                        "rel_residual": info.rel_residual,
                        "true_rel_residual": info.true_rel_residual, "preconditioner": prec}
        pp = getattr(op, "_poisson_prec", None)
        if pp is not None:
            out["gmres"]["poisson_solves"] = pp.solves
            out["gmres"]["poisson_outer_iterations"] = pp.inner_outer_iterations
    if len(a) > 10 and a[10] == "newton":
        # the whole nonlinear solve (Newton + flexible GMRES + Poisson block preconditioner)
        bc = stokes.squirmer_boundary_data(dm, 1.0, stokes.squirmer_vslip_profile(1.0),
                                           x_phys=op.x_phys.cpu().numpy().reshape(op.n_elem, 2, p + 1, p + 1))
        op.set_essential(bc.essential)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        state, hist = op.newton_solve(op.from_host(bc.state0), bc.cint, it_max=12, tol=1e-6,
                                      gmres_rtol=1e-8, restart=300, gmres_maxiter=900,
                                      precondition="poisson", verbose=True)
        torch.cuda.synchronize()
        out["newton"] = {"seconds": time.perf_counter() - t0, "n_rey": n_rey,
                         "steps": [(du, i.iterations, i.true_rel_residual) for du, i in hist]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
