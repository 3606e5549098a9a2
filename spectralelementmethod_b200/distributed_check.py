"""Self-check of the partitioned (multi-GPU) path against the single-GPU operator.

Run under torchrun on N GPUs (tests/multigpu_check.py prints the result; bench.py calls
it before timing and puts the numbers into its JSON line as ``parity``).  Every rank
builds its strip; every rank ALSO builds the small global problem on its own GPU and
checks, on the rows it holds (gathered by global id):

  * apply / fused dot / diagonal / RHS / lifted RHS / Jacobi-PCG solution of the
    distributed operator, for the NVLink peer-memory exchange and for NCCL p2p, and that
    the two exchanges give bit-identical results;
  * the peer-memory all-reduce (csrc/semk_ml.cu) against NCCL, bit-identical on all ranks;
  * the statically condensed path (distributed Jacobi-PCG) and the native multilevel
    driver (two- and three-level) against the global solve: solution, true residual and
    outer iteration counts against the single-GPU multilevel solve of the global problem.

Raises AssertionError on any mismatch.  No oracle / CPU code is involved: the reference
here is this package's own single-GPU path, which tests/ pins against the oracle."""
import numpy as np
import torch
import torch.distributed as dist

from . import discrete, meshgen
from .basis_functions import LagrangeGaussLobatto, TensorProductQS
from .distributed import (DistributedCondensedPoisson, DistributedPoisson, PeerComm,
                          StripPartition)

__all__ = ["run"]


def run(rank, world, dev, nxl=24, ny=20, p=8, kind="C"):
    """Returns a dict of the measured margins (see the module docstring)."""
    bounds = (-1.0, -1.0 + 2.0 * world, -1.0, 1.0)
    part = StripPartition(rank, world, nxl, ny, p, bounds=bounds)
    gid = torch.from_numpy(part.global_ids()).to(dev)

    # global problem on every rank's own GPU (small), as the reference
    gx = meshgen.lattice_coordinates(kind, nxl * world, ny, p, bounds)
    gmesh = meshgen.structured_quad_mesh(nxl * world, ny, p, kind, bounds, nodes=gx)
    b1 = LagrangeGaussLobatto(p)
    gm = discrete.DOFManager(gmesh, 1, TensorProductQS(b1, b1), rcm_order=False)
    gon = gm.boundary_node_mask("ebc")
    gop = gm.poisson_operator(dirichlet=gon)

    g = torch.Generator(device=dev).manual_seed(7)
    ug = torch.randn(gop.n_nodes, dtype=torch.float64, device=dev, generator=g)
    want = gop.apply(ug)
    dot_ref = torch.zeros(1, dtype=torch.float64, device=dev)
    gop.apply(ug, dot_out=dot_ref)
    bg = gop.lift(gop.rhs(1.0), None)
    xg, info = gop.solve_pcg(bg, rtol=1e-12, check_every=10)
    results = {}
    for exchange in ("peer", "nccl"):
        dp = DistributedPoisson(part, p, kind, exchange=exchange)
        assert np.array_equal(dp.on_ebc, gon[part.global_ids()])
        results[exchange] = check_one(dp, gop, gid, ug, want, dot_ref, bg, xg, dev)
        if dp.halo is not None:
            dp.halo.check()
        dp.close()
    # the two exchange paths add the same two numbers: bit-identical results
    assert torch.equal(results["peer"][0], results["nccl"][0])
    # the peer-memory all-reduce against NCCL: same sums, bit-identical on every rank
    comm = PeerComm(rank, world, 4104, None, dev)
    for n in (1, 2, 4097, 4104):
        for rep in range(3):                    # both parities of the double buffer
            v = torch.randn(n, dtype=torch.float64, device=dev, generator=g) * (rank + 1)
            ref = v.clone()
            dist.all_reduce(ref)
            got = comm.allreduce(v.clone())
            assert float((got - ref).abs().max()) <= 1e-13 * float(ref.abs().max() + 1), n
            gathered = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(gathered, got)
            assert all(torch.equal(gathered[0], q) for q in gathered)
    comm.check()
    comm.close()
    # statically condensed path: distributed PCG on the exterior DOFs + local back-solve;
    # single-GPU multilevel solves of the GLOBAL problem as the reference for the counts
    gmesh2 = meshgen.structured_quad_mesh(nxl * world, ny, p, kind, bounds, nodes=gx.copy())
    gsc_m = discrete.DOFManagerSC(gmesh2, 1, TensorProductQS(b1, b1), rcm_order=False)
    gsc = gsc_m.condensed_poisson_operator(dirichlet=gsc_m.boundary_node_mask("ebc"))
    single = {}
    for pre in ("two-level", "three-level"):
        _, inf = gsc.solve(1.0, None, rtol=1e-12, preconditioner=pre, max_tiles=12)
        single[pre] = inf
    sc_res = {}
    for exchange in ("peer", "nccl"):
        dc = DistributedCondensedPoisson(part, p, kind, exchange=exchange)
        cg = torch.from_numpy(dc.global_ids()).to(dev)
        # host-driven loop on both exchanges: the two exchanges add the same two numbers, so
        # the iterates are bit-identical
        xs, it_sc, rel_sc, ok_sc = dc.solve(1.0, None, rtol=1e-12, check_every=10, native=False)
        serr_sc = float((xs - xg[cg]).norm() / xg.norm())
        assert ok_sc and serr_sc < 1e-9, (ok_sc, serr_sc)
        sc_res[exchange] = (xs, it_sc, serr_sc)
        if exchange == "peer":
            # the native driver (peer all-reduces): same recurrence, same solution
            xn, it_n, rel_n, ok_n = dc.solve(1.0, None, rtol=1e-12, check_every=10)
            assert ok_n and it_sc - 10 <= it_n <= it_sc + 3, (it_n, it_sc)   # (the host loop counts to its next poll)
            assert float((xn - xs).norm() / xs.norm()) < 1e-9
        if exchange == "peer":
            # native multilevel driver on the partition (semk_sc_mlpcg_solve_f64): halo
            # exchanges and all-reduces are peer-memory kernels issued by the driver
            for pre in ("two-level", "three-level"):
                for rep in range(2):
                    x2, it2, rel2, ok2 = dc.solve(1.0, None, rtol=1e-12, preconditioner=pre,
                                                  max_tiles=12)
                    inf = dc.last_info
                    serr2 = float((x2 - xg[cg]).norm() / xg.norm())
                    assert ok2 and serr2 < 1e-9 and it2 < it_sc, (pre, ok2, serr2, it2, it_sc)
                    assert inf.true_rel_residual < 1e-10, inf.true_rel_residual
                    assert abs(it2 - single[pre].iterations) <= 3, (pre, it2, single[pre].iterations)
                sc_res[pre] = (it2, inf.inner_iterations, inf.true_rel_residual, serr2)
            assert sc_res["three-level"][1] < sc_res["two-level"][1]
        dc.close()
    assert torch.equal(sc_res["peer"][0], sc_res["nccl"][0])
    err, it, serr = results["peer"][1:]
    out = {"world": world, "problem": "%dx%d elements per rank, p=%d, kind %s" % (nxl, ny, p, kind),
           "apply_rel_err": err, "pcg_iterations": it, "pcg_iterations_single_gpu": info.iterations,
           "pcg_solution_rel_diff": serr, "peer_equals_nccl_bitwise": True,
           "peer_allreduce_equals_nccl": True,
           "condensed_pcg_iterations": sc_res["peer"][1],
           "condensed_solution_rel_diff": sc_res["peer"][2]}
    for pre in ("two-level", "three-level"):
        r = sc_res[pre]
        out[pre] = {"outer_iterations": r[0], "inner_iterations": r[1],
                    "outer_iterations_single_gpu": single[pre].iterations,
                    "inner_iterations_single_gpu": single[pre].inner_iterations,
                    "true_rel_residual": r[2], "solution_rel_diff": r[3]}
    return out


def check_one(dp, gop, gid, ug, want, dot_ref, bg, xg, dev):
    u = ug[gid].contiguous()
    dot = torch.zeros(1, dtype=torch.float64, device=dev)
    y = dp.apply(u, dot_out=dot)
    dist.all_reduce(dot)
    err = float((y - want[gid]).norm() / want.norm())
    assert err < 1e-13, err
    assert abs(float(dot) - float(dot_ref)) < 1e-11 * abs(float(dot_ref)), (float(dot), float(dot_ref))
    assert float((dp.diagonal() - gop.diagonal()[gid]).abs().max()) < 1e-11
    assert float((dp.rhs(1.0) - gop.rhs(1.0)[gid]).abs().max()) < 1e-14

    # the host-buffer path: staged local apply + device-side exchange + second download of the
    # exchanged columns; bitwise equal to the same two steps on device-resident vectors
    uh = torch.empty(u.numel(), dtype=torch.float64).pin_memory()
    yh = torch.empty(u.numel(), dtype=torch.float64).pin_memory()
    uh.copy_(u)
    dp.apply_host(uh, yh, stages=3)
    y_dev = dp.dop.finish(dp.host_operator().apply(u), u)
    assert torch.equal(yh.to(dev), y_dev)
    assert float((y_dev - want[gid]).norm() / want.norm()) < 1e-13

    b = dp.lift(dp.rhs(1.0), None)
    assert float((b - bg[gid]).abs().max()) < 1e-13
    x, it, rel, ok = dp.solve_pcg(b, rtol=1e-12, check_every=10)
    serr = float((x - xg[gid]).norm() / xg.norm())
    assert ok and serr < 1e-9, (ok, serr)
    if dp.halo is not None:
        # the native driver (peer all-reduces) took the solve above; the host-driven loop
        # with NCCL all-reduces must agree (same recurrence, reductions in another order)
        xh, ith, relh, okh = dp.solve_pcg(b, rtol=1e-12, check_every=10, native=False)
        assert okh and ith - 10 <= it <= ith + 3, (ith, it)   # (the host loop counts to its next poll)
        assert float((xh - x).norm() / x.norm()) < 1e-9
    return y, err, it, serr

