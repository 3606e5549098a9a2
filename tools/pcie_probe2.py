"""Where do the 13.1 ms of one staged host apply go?  (PCIe probe, part 2.)  537 MB each way:
(a) 16 + 16 chunked copies on two streams, no dependencies; (b) the same with the staged
apply's dependency pattern (download i waits for upload i) but no kernels; (c) as (b) with a
50 us dummy kernel between; (d) 8 back-to-back repetitions of (b) (steady state)."""
import json
import sys
import time

import torch

n = (1024 * 8 + 1) ** 2
S = 16
h_u = torch.empty(n, dtype=torch.float64).pin_memory()
h_y = torch.empty(n, dtype=torch.float64).pin_memory()
h_u.normal_()
d_u = torch.empty(n, dtype=torch.float64, device="cuda")
d_y = torch.empty(n, dtype=torch.float64, device="cuda")
up, down, comp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
cuts = [(i * n) // S for i in range(S + 1)]
res = {}


def timed(fn, reps=4):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def chunks_free():
    for i in range(S):
        a, b = cuts[i], cuts[i + 1]
        with torch.cuda.stream(up):
            d_u[a:b].copy_(h_u[a:b], non_blocking=True)
        with torch.cuda.stream(down):
            h_y[a:b].copy_(d_y[a:b], non_blocking=True)


def chunks_dep(kernel=False, reps=1):
    for _ in range(reps):
        for i in range(S):
            a, b = cuts[i], cuts[i + 1]
            with torch.cuda.stream(up):
                d_u[a:b].copy_(h_u[a:b], non_blocking=True)
                e = torch.cuda.Event()
                e.record(up)
            with torch.cuda.stream(comp):
                comp.wait_event(e)
                if kernel:
                    d_y[a:b].copy_(d_u[a:b], non_blocking=True)
                e2 = torch.cuda.Event()
                e2.record(comp)
            with torch.cuda.stream(down):
                down.wait_event(e2)
                h_y[a:b].copy_(d_y[a:b], non_blocking=True)


for name, fn, div in (("chunked_free", chunks_free, 1), ("chunked_dep", chunks_dep, 1),
                      ("chunked_dep_kernel", lambda: chunks_dep(True), 1),
                      ("chunked_dep_x8", lambda: chunks_dep(False, 8), 8),
                      ("chunked_dep_kernel_x8", lambda: chunks_dep(True, 8), 8)):
    t = timed(fn) / div
    res[name] = {"ms_per_537MB_each_way": 1e3 * t, "GBps_per_direction": 8 * n / t / 1e9}
    print(name, res[name], flush=True)
json.dump(res, open("gpurun_out/r02_pcie_probe2.json", "w"), indent=1)
