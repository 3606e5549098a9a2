"""Where does one Poisson solve of the Stokes preconditioner spend its time?  (GPU box)
python tests/stokes_prec_profile.py [nr nt p]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from spectralelementmethod_b200 import discrete, meshgen, stokes  # noqa: E402
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS  # noqa: E402

a = sys.argv[1:]
nr, nt, p = (int(a[0]), int(a[1]), int(a[2])) if len(a) >= 3 else (224, 352, 8)
mesh = meshgen.annulus_sector_mesh(nr, nt, p, 100.0)
b1 = LagrangeGaussLobatto(p)
dm = discrete.DOFManagerSC(mesh, 2, TensorProductQS(b1, b1), rcm_order=False)
op = dm.axisymmetric_stokes_operator()
bc = stokes.squirmer_boundary_data(dm, 1.0, stokes.squirmer_vslip_profile(1.0),
                                   x_phys=op.x_phys.cpu().numpy().reshape(op.n_elem, 2, p + 1, p + 1))
op.set_essential(bc.essential)
pre = stokes.PoissonBlockPreconditioner(op)
sc = pre.sc
r = torch.randn(op.n_nodes, dtype=torch.float64, device=op.dev) * pre.free_s
pre.poisson_solve(r)


def timed(fn, reps=5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


f = r * pre.inv_mass_int
t_rhs, g = timed(lambda: sc.rhs(f))
g = g + r[:pre.n_ext]
t_lift, b = timed(lambda: sc.lift(g, None))
t_pcg, (x, info) = timed(lambda: sc.solve_pcg(b, rtol=pre.rtol, preconditioner="three-level"))
t_back, u = timed(lambda: sc.backsolve(x, f))
t_all, _ = timed(lambda: pre.poisson_solve(r))
v = torch.randn(op.n_dof, dtype=torch.float64, device=op.dev)
w = op.new_vector()
t_prec, _ = timed(lambda: pre(v, w), 3)
print("n_ext %d, elements %d: rhs %.2f ms, lift %.2f ms, pcg %.2f ms (%d outer / %d inner), "
      "backsolve %.2f ms, whole solve %.2f ms, whole preconditioner %.2f ms"
      % (pre.n_ext, sc.n_elem, t_rhs, t_lift, t_pcg, info.iterations, info.inner_iterations,
         t_back, t_all, t_prec))
