"""Profiling helper (GPU box): one three-level solve at a given size, for an ncu launch list.

    python tests/ml_profile.py [nx] [order] [max_tiles]
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
    max_tiles = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    mesh = meshgen.structured_quad_mesh(nx, nx, p, "S")
    b1 = LagrangeGaussLobatto(p)
    mngr = discrete.DOFManagerSC(mesh, 1, TensorProductQS(b1, b1), rcm_order=False)
    sc = mngr.condensed_poisson_operator(dirichlet=mngr.boundary_node_mask("ebc"))
    b = sc.lift(sc.rhs(1.0), None)
    sc._build_top(max_tiles)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, info = sc.solve_pcg(b, rtol=1e-12, preconditioner="three-level", max_tiles=max_tiles)
        torch.cuda.synchronize()
        print("solve %.4f s, %d outer, %d inner" % (time.perf_counter() - t0, info.iterations,
                                                    info.inner_iterations), flush=True)


if __name__ == "__main__":
    main()
