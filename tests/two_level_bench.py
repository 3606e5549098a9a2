"""Time to solution of the two-level preconditioned condensed PCG at config 2 (GPU box):

    python tests/two_level_bench.py [nx] [order] [inner_rtol] [two-level|three-level]

("three-level" = the experimental aggregation level under the vertex coarse space; it is
compared with the two-level solution in the same run.)
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from spectralelementmethod_b200 import discrete, meshgen  # noqa: E402
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS  # noqa: E402


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    p = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    inner_rtol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-2
    which = sys.argv[4] if len(sys.argv) > 4 else "two-level"
    max_tiles = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
    skip_two = len(sys.argv) > 6 and sys.argv[6] == "skip-two-level"
    mesh = meshgen.structured_quad_mesh(nx, nx, p, "S")
    b1 = LagrangeGaussLobatto(p)
    mngr = discrete.DOFManagerSC(mesh, 1, TensorProductQS(b1, b1), rcm_order=False)
    sc = mngr.condensed_poisson_operator(dirichlet=mngr.boundary_node_mask("ebc"))
    b = sc.lift(sc.rhs(1.0), None)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sc._build_coarse()
    torch.cuda.synchronize()
    t_coarse = time.perf_counter() - t0
    for rep in range(0 if skip_two else 2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, info = sc.solve_pcg(b, rtol=1e-12, preconditioner="two-level", inner_rtol=inner_rtol)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        print("two-level nx=%d p=%d n_ext=%d: coarse build %.2f s; solve %.3f s, %d outer its, "
              "%d inner its, converged %s, rel residual %.2e"
              % (nx, p, sc.n_ext, t_coarse, el, info.iterations, sc.last_inner_iterations,
                 info.converged, info.rel_residual), flush=True)
    if not skip_two:
        u = sc.backsolve(x, 1.0)
        print("checksum %.15g, true residual %.2e" % (float(u.sum()), info.true_rel_residual),
              flush=True)
    if which == "three-level":
        t0 = time.perf_counter()
        sc._build_top(max_tiles)
        torch.cuda.synchronize()
        t_top = time.perf_counter() - t0
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            x3, info3 = sc.solve_pcg(b, rtol=1e-12, preconditioner="three-level",
                                     inner_rtol=inner_rtol, max_tiles=max_tiles)
            torch.cuda.synchronize()
            el = time.perf_counter() - t0
            print("three-level (%d aggregates): top build %.2f s; solve %.3f s, %d outer its, "
                  "%d inner its, converged %s, rel residual %.2e, true residual %.2e%s"
                  % (sc._top[2], t_top, el, info3.iterations, sc.last_inner_iterations,
                     info3.converged, info3.rel_residual, info3.true_rel_residual,
                     "" if skip_two else ", diff to two-level %.2e"
                     % float((x3 - x).norm() / x.norm())), flush=True)


if __name__ == "__main__":
    main()
