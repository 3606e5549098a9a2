"""Timing of the condensed (static-condensation) apply alone, for A/B runs and ncu
captures on the GPU box:

    python tests/condensed_bench.py [nx] [order] [steps]
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
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    t0 = time.perf_counter()
    mesh = meshgen.structured_quad_mesh(nx, nx, p, "S")
    b1 = LagrangeGaussLobatto(p)
    mngr = discrete.DOFManagerSC(mesh, 1, TensorProductQS(b1, b1), rcm_order=False)
    sc = mngr.condensed_poisson_operator(dirichlet=mngr.boundary_node_mask("ebc"))
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0
    u = torch.randn(sc.n_ext, dtype=torch.float64, device="cuda",
                    generator=torch.Generator(device="cuda").manual_seed(0))
    out = torch.empty_like(u)
    dot = torch.zeros(1, dtype=torch.float64, device="cuda")
    for _ in range(5):
        sc.apply(u, out=out, dot_out=dot)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        sc.apply(u, out=out, dot_out=dot)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    alg = sc.algorithmic_bytes_per_apply
    print("nx=%d p=%d n_ext=%d setup %.1fs  apply %.4f ms  %.0f GB/s algorithmic  checksum %.17g"
          % (nx, p, sc.n_ext, t_setup, ms, alg / ms / 1e6, float(out.double().sum())), flush=True)


if __name__ == "__main__":
    main()
