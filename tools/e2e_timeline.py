"""Timeline of the staged host apply (semk_poisson_apply_host_staged_f64), rebuilt in Python from
apply_range + three streams + timing events: when does every upload / compute / download of
the 16 stages start and end?  Config 2 (1024 x 1024 elements, p = 8)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from spectralelementmethod_b200 import discrete, meshgen  # noqa: E402
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS  # noqa: E402

nx, p, S = 1024, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 16
mesh = meshgen.structured_quad_mesh(nx, nx, p, "S")
b1 = LagrangeGaussLobatto(p)
mngr = discrete.DOFManager(mesh, 1, TensorProductQS(b1, b1), rcm_order=False)
op = mngr.poisson_operator(dirichlet=mngr.boundary_node_mask("ebc"))
n = op.n_nodes
h_u = torch.empty(n, dtype=torch.float64).pin_memory().normal_()
h_y = torch.empty(n, dtype=torch.float64).pin_memory()
d_u, d_y = op.new_vector(), op.new_vector()
arr, ns = op.stage_table(S)
st = [(arr[i].patch_end, arr[i].chunk_end, arr[i].rec_end, arr[i].u_need, arr[i].y_final) for i in range(ns)]
up, down = torch.cuda.Stream(), torch.cuda.Stream()


def run(comp, timing):
    ev = lambda: torch.cuda.Event(enable_timing=timing)  # noqa: E731
    t0 = ev()
    t0.record(comp)
    up.wait_event(t0)
    down.wait_event(t0)
    marks = []
    pb = cb = rb = ub = yb = 0
    for (pe, ce, re, un, yf) in st:
        with torch.cuda.stream(up):
            a = ev(); a.record(up)
            if un > ub:
                d_u[ub:un].copy_(h_u[ub:un], non_blocking=True)
            b = ev(); b.record(up)
        with torch.cuda.stream(comp):
            comp.wait_event(b)
            c = ev(); c.record(comp)
            op.apply_range(d_u, d_y, pb, pe, cb, ce, rb, re)
            d = ev(); d.record(comp)
        with torch.cuda.stream(down):
            down.wait_event(d)
            e = ev(); e.record(down)
            if yf > yb:
                h_y[yb:yf].copy_(d_y[yb:yf], non_blocking=True)
            f = ev(); f.record(down)
        marks.append((a, b, c, d, e, f))
        pb, cb, rb, ub, yb = pe, ce, re, un, yf
    comp.wait_event(marks[-1][5])
    return t0, marks


for name, comp in (("default stream", torch.cuda.default_stream()), ("side stream", torch.cuda.Stream())):
    with torch.cuda.stream(comp):
        run(comp, False)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(3):
            run(comp, False)
        torch.cuda.synchronize()
        print("%s: %.3f ms per apply (python-driven pipeline, no timing events)"
              % (name, (time.perf_counter() - w0) / 3 * 1e3), flush=True)
        t0, marks = run(comp, True)
        torch.cuda.synchronize()
    print("stage   up_start up_end | comp_start comp_end | down_start down_end   (ms after start)")
    for i, m in enumerate(marks):
        print("%5d  %8.3f %8.3f | %8.3f %8.3f | %8.3f %8.3f" % ((i,) + tuple(t0.elapsed_time(x) for x in m)))
# does a smaller persistent grid (less HBM pressure while the copy engines run) help?
for cap in (0, 296, 148, 74, 37):
    op._op.max_ctas = cap
    for S2 in (16, 32):
        t = time.perf_counter()
        for _ in range(4):
            op.apply_host(h_u, h_y, (d_u, d_y), stages=S2)
        print("max_ctas %3d stages %2d: %.3f ms per apply (C driver)" % (cap, S2, (time.perf_counter() - t) / 4 * 1e3), flush=True)
op._op.max_ctas = 0
ref = op.apply(d_u)
print("matches device apply:", torch.equal(h_y.cuda(), ref))
t = time.perf_counter()
for _ in range(3):
    op.apply_host(h_u, h_y, (d_u, d_y), stages=S)
print("C driver: %.3f ms per apply" % ((time.perf_counter() - t) / 3 * 1e3))
