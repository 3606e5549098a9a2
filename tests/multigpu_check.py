"""Multi-GPU parity check, launched by hand under torchrun on the GPU box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multigpu_check.py

Every rank builds its strip; rank 0 additionally builds the *global* mesh on
its own GPU and checks that the distributed apply / RHS / diagonal / PCG
solution equal the single-GPU results on the rows it holds (gathered by global
id).  Not collected by pytest (needs N GPUs)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from spectralelementmethod_b200 import discrete, meshgen  # noqa: E402
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS  # noqa: E402
from spectralelementmethod_b200.distributed import (DistributedCondensedPoisson, DistributedPoisson,  # noqa: E402
                                                    StripPartition)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    nxl, ny, p, kind = 24, 20, 8, "C"
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
            dp.halo.close()
    # the two exchange paths add the same two numbers: bit-identical results
    assert torch.equal(results["peer"][0], results["nccl"][0])
    # statically condensed path: distributed PCG on the exterior DOFs + local back-solve
    sc_res = {}
    for exchange in ("peer", "nccl"):
        dc = DistributedCondensedPoisson(part, p, kind, exchange=exchange)
        cg = torch.from_numpy(dc.global_ids()).to(dev)
        xs, it_sc, rel_sc, ok_sc = dc.solve(1.0, None, rtol=1e-12, check_every=10)
        serr_sc = float((xs - xg[cg]).norm() / xg.norm())
        assert ok_sc and serr_sc < 1e-9, (ok_sc, serr_sc)
        sc_res[exchange] = (xs, it_sc, serr_sc)
        # two-level preconditioner on the partition (device wiring of
        # distributed.distributed_two_level_pcg)
        x2, it2, rel2, ok2 = dc.solve(1.0, None, rtol=1e-12, preconditioner="two-level")
        serr2 = float((x2 - xg[cg]).norm() / xg.norm())
        assert ok2 and serr2 < 1e-9 and it2 < it_sc, (ok2, serr2, it2, it_sc)
        sc_res[exchange] += (it2, dc.last_inner_iterations)
        if getattr(dc, "_halo_c", None) is not None:
            dc._halo_c.check()
            dc._halo_c.close()
        if dc.halo is not None:
            dc.halo.check()
            dc.halo.close()
    assert torch.equal(sc_res["peer"][0], sc_res["nccl"][0])
    if rank == 0:
        err, it, serr = results["peer"][1:]
        print("multigpu_check ok: world=%d apply err %.2e, PCG %d its (single GPU %d), "
              "solution diff %.2e; peer == nccl bitwise" % (world, err, it, info.iterations, serr))
        print("multigpu_check condensed ok: PCG %d its, solution diff %.2e vs the single-GPU "
              "uncondensed solve; peer == nccl bitwise" % (sc_res["peer"][1], sc_res["peer"][2]))
        print("multigpu_check two-level ok: %d outer / %d inner its"
              % (sc_res["peer"][3], sc_res["peer"][4]))
    dist.barrier()
    dist.destroy_process_group()


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

    b = dp.lift(dp.rhs(1.0), None)
    assert float((b - bg[gid]).abs().max()) < 1e-13
    x, it, rel, ok = dp.solve_pcg(b, rtol=1e-12, check_every=10)
    serr = float((x - xg[gid]).norm() / xg.norm())
    assert ok and serr < 1e-9, (ok, serr)
    return y, err, it, serr


if __name__ == "__main__":
    main()
