"""PCIe ceiling of the end-to-end apply (bench.py `e2e`): pinned 537 MB buffers (config 2,
one nodal vector) copied host->device, device->host and both at once on two streams, next to
the staged host apply (semk_poisson_apply_host_staged_f64) at several stage counts."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from spectralelementmethod_b200 import discrete, meshgen  # noqa: E402
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS  # noqa: E402

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
p = 8
n = (nx * p + 1) ** 2
h_u = torch.empty(n, dtype=torch.float64).pin_memory()
h_y = torch.empty(n, dtype=torch.float64).pin_memory()
h_u.normal_()
d_u = torch.empty(n, dtype=torch.float64, device="cuda")
d_y = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {"bytes_per_vector": 8 * n}


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_u.copy_(h_u, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_y.copy_(d_y, non_blocking=True)


def both():
    h2d()
    d2h()


for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    t = timed(fn)
    res[name] = {"ms": 1e3 * t, "GBps_per_direction": 8 * n / t / 1e9}
    print(name, res[name], flush=True)

mesh = meshgen.structured_quad_mesh(nx, nx, p, "S")
b1 = LagrangeGaussLobatto(p)
mngr = discrete.DOFManager(mesh, 1, TensorProductQS(b1, b1), rcm_order=False)
op = mngr.poisson_operator(dirichlet=mngr.boundary_node_mask("ebc"))
for stages in (1, 4, 8, 16, 32, 64):
    t = timed(lambda: op.apply_host(h_u, h_y, scratch=(d_u, d_y), stages=stages), reps=4)
    res["staged_%d" % stages] = {"ms": 1e3 * t, "gdof_per_s": n / t / 1e9}
    print("stages", stages, res["staged_%d" % stages], flush=True)
json.dump(res, open("gpurun_out/r02_pcie_probe.json", "w"), indent=1)
