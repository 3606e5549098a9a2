"""Config-2 condensed operator (DOFManagerSC numbering, 1024 x 1024 elements, p = 8): a few
applies, for ncu captures of sc_matvec_kernel / sc_node_kernel."""
import sys

import torch

sys.path.insert(0, ".")
from spectralelementmethod_b200 import discrete, meshgen  # noqa: E402
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
mesh = meshgen.structured_quad_mesh(nx, nx, 8, "S")
b1 = LagrangeGaussLobatto(8)
mngr = discrete.DOFManagerSC(mesh, 1, TensorProductQS(b1, b1), rcm_order=False)
sc = mngr.condensed_poisson_operator(dirichlet=mngr.boundary_node_mask("ebc"))
u = torch.randn(sc.n_ext, dtype=torch.float64, device="cuda",
                generator=torch.Generator(device="cuda").manual_seed(0))
out = torch.empty_like(u)
for _ in range(reps):
    sc.apply(u, out=out)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(reps):
    sc.apply(u, out=out)
ev1.record()
torch.cuda.synchronize()
print("condensed apply: %.4f ms, checksum %.12e" % (ev0.elapsed_time(ev1) / reps, float(out.sum())))
