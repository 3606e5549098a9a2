"""Staged host apply with the result written by the apply kernels STRAIGHT into pinned host
memory (zero-copy stores over PCIe; no device-to-host copies): does the write-out fused with
the download beat the copy engine?  Config 2, 16 / 32 stages."""
import ctypes as C
import sys
import time

import torch

sys.path.insert(0, ".")
from spectralelementmethod_b200 import _lib, device, discrete, meshgen  # noqa: E402
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS  # noqa: E402

nx, p = 1024, 8
mesh = meshgen.structured_quad_mesh(nx, nx, p, "S")
b1 = LagrangeGaussLobatto(p)
mngr = discrete.DOFManager(mesh, 1, TensorProductQS(b1, b1), rcm_order=False)
op = mngr.poisson_operator(dirichlet=mngr.boundary_node_mask("ebc"))
lib = _lib.load()
n = op.n_nodes
h_u = torch.empty(n, dtype=torch.float64).pin_memory().normal_()
h_y = torch.empty(n, dtype=torch.float64).pin_memory()
d_u, d_y = op.new_vector(), op.new_vector()
up = torch.cuda.Stream()
flags = int(op._masked_flags)


def run(S, timing=False):
    arr, ns = op.stage_table(S)
    comp = torch.cuda.current_stream()
    ev = lambda: torch.cuda.Event(enable_timing=timing)  # noqa: E731
    t0 = ev(); t0.record(comp)
    up.wait_event(t0)
    marks = []
    pb = cb = rb = ub = 0
    for i in range(ns):
        s = arr[i]
        with torch.cuda.stream(up):
            a = ev(); a.record(up)
            if s.u_need > ub:
                d_u[ub:s.u_need].copy_(h_u[ub:s.u_need], non_blocking=True)
            b = ev(); b.record(up)
        comp.wait_event(b)
        c = ev(); c.record(comp)
        _lib.check(lib.semk_poisson_apply_range_f64(
            C.byref(op._op), device.ptr(d_u), C.c_void_p(h_y.data_ptr()), flags, pb, s.patch_end, cb,
            s.chunk_end, rb, s.rec_end, C.c_void_p(comp.cuda_stream)))
        d = ev(); d.record(comp)
        marks.append((a, b, c, d))
        pb, cb, rb, ub = s.patch_end, s.chunk_end, s.rec_end, s.u_need
    return t0, marks


ref = None
for S in (16, 32, 8):
    run(S)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    for _ in range(4):
        run(S)
    torch.cuda.synchronize()
    print("zero-copy y, %2d stages: %.3f ms per apply" % (S, (time.perf_counter() - w0) / 4 * 1e3), flush=True)
    if ref is None:
        ref = op.apply(d_u).cpu()
    print("   matches device apply:", torch.equal(h_y, ref), flush=True)
t0, marks = run(16, True)
torch.cuda.synchronize()
print("stage   up_start up_end | comp_start comp_end   (ms after start)")
for i, m in enumerate(marks):
    print("%5d  %8.3f %8.3f | %8.3f %8.3f" % ((i,) + tuple(t0.elapsed_time(x) for x in m)))
w0 = time.perf_counter()
for _ in range(4):
    op.apply_host(h_u, h_y, (d_u, d_y), stages=16)
print("copy-engine pipeline (C driver), 16 stages: %.3f ms per apply" % ((time.perf_counter() - w0) / 4 * 1e3))
# whole apply in one launch with y in host memory (u already on the device)
torch.cuda.synchronize()
w0 = time.perf_counter()
for _ in range(3):
    _lib.check(lib.semk_poisson_apply_f64(C.byref(op._op), device.ptr(d_u), C.c_void_p(h_y.data_ptr()),
                                          flags, None, device.stream_ptr()))
torch.cuda.synchronize()
print("one full apply writing y to host memory: %.3f ms" % ((time.perf_counter() - w0) / 3 * 1e3))
