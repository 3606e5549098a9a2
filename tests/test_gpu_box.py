"""kernel_variant 2 (csrc/semk_box.cu): the apply kernel with the arithmetic gather for
regularly numbered structured meshes.  It evaluates the same sums in the same order as the
table-driven kernel, so every result must be BIT-IDENTICAL to it; the table-driven kernel is
in turn checked against the oracle (tests/test_gpu_parity.py), and a few cases repeat the
oracle comparison here (T1: the reference's own invJ / detJxW, <= 1e-12)."""
import numpy as np
import pytest
import torch

import sem_oracle as so
from conftest import build_package_case, rel_l2
from spectralelementmethod_b200 import _lib

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def host(t):
    return t.detach().cpu().numpy()


CASES = [
    # kind, nx, ny, p, elems_per_patch          tiles        note
    ("S", 8, 24, 8, 16),                      # 4 x 3       exact tiles, interior patches
    ("C", 7, 19, 8, 16),                      # ragged in both directions (empty slots)
    ("C", 6, 24, 4, 16),
    ("C", 3, 25, 6, 8),                       # 1 x 8 tiles, ragged
    ("S", 4, 16, 12, 8),
    ("C", 5, 17, 3, 8),
    ("C", 8, 32, 2, 16),                      # n1 = 3: one interior node row per tile
    ("C", 6, 24, 5, 16),                      # odd order: node rows of a tile start at odd ids
    ("C", 12, 24, 4, 32),                     # 4 x 8 tiles (the automatic patch size for p <= 4)
    ("C", 9, 19, 3, 32),                      # ... ragged
]


def _pair(mngr, pe, **kw):
    box = mngr.poisson_operator(mode="box", elems_per_patch=pe, **kw)
    col = mngr.poisson_operator(mode="column", elems_per_patch=pe, **kw)
    assert col.kernel_variant == 0
    return box, col


@pytest.mark.parametrize("kind,nx,ny,p,pe", CASES)
def test_box_kernel_is_bit_identical_to_table_kernel(kind, nx, ny, p, pe):
    mesh, mngr = build_package_case(kind, nx, ny, p, False, False)
    rng = np.random.default_rng(11)
    n = mngr.ndof
    u = dev(rng.standard_normal(n))
    # (a) no Dirichlet nodes: every full or ragged tile is a box
    box, col = _pair(mngr, pe)
    assert box.kernel_variant == 2 and box.box_ld == ny * p + 1
    assert box.n_box_patches == box.n_patch
    y = box.apply_unmasked(u)
    assert torch.equal(y, col.apply_unmasked(u))
    assert torch.equal(y, box.apply_unmasked(u))            # deterministic
    # (b) essential boundary on two sides + one interior Dirichlet node: the patches that
    # hold a Dirichlet node fall back to the tables inside the same kernel
    on = np.zeros(n, dtype=bool)
    ids = np.arange(n).reshape(nx * p + 1, ny * p + 1)
    on[ids[0, :]] = True
    on[ids[:, 0]] = True
    on[ids[nx * p - 1, ny * p - 1]] = True                # inside the last tile
    box, col = _pair(mngr, pe, dirichlet=on)
    assert box.kernel_variant == 2 and 0 < box.n_box_patches < box.n_patch
    d0 = torch.zeros(1, dtype=torch.float64, device="cuda")
    d1 = torch.zeros(1, dtype=torch.float64, device="cuda")
    yb = box.apply(u, dot_out=d0)
    yc = col.apply(u, dot_out=d1)
    assert torch.equal(yb, yc)
    assert float(d0) == float(d1)
    for flags in (0, _lib.MASK_IN, _lib.MASK_OUT, _lib.MASK_IN | _lib.MASK_OUT):
        assert torch.equal(box.apply(u, flags=flags), col.apply(u, flags=flags))
    # the staged host apply runs sub-ranges of the patch sequence through the same kernel
    uh = torch.empty(n, dtype=torch.float64).pin_memory()
    uh.copy_(u.cpu())
    yh = torch.empty(n, dtype=torch.float64).pin_memory()
    box.apply_host(uh, yh, stages=3)
    assert torch.equal(yh, yc.cpu())
    # solves through the native PCG driver agree bit for bit as well
    b = box.lift(box.rhs(1.0), None)
    x0, i0 = box.solve_pcg(b, rtol=1e-10, maxiter=20000)
    x1, i1 = col.solve_pcg(b, rtol=1e-10, maxiter=20000)
    assert i0.converged and i0.iterations == i1.iterations and torch.equal(x0, x1)


@pytest.mark.parametrize("kind,nx,ny,p,pe", [("C", 6, 17, 8, 16), ("C", 3, 9, 4, 8), ("S", 2, 8, 10, 8)])
def test_box_kernel_vs_oracle(kind, nx, ny, p, pe):
    mesh, mngr = build_package_case(kind, nx, ny, p, False, False)
    r = so.run_case(kind, nx, ny, p, False, False, solve=False)
    rng = np.random.default_rng(12)
    u = rng.standard_normal(mngr.ndof)
    ref = so.apply_dense_batched(r["L"], r["l2g"], u)
    op = mngr.poisson_operator(geometric_factors=(r["invJ"], r["JxW"]), mode="box",
                               elems_per_patch=pe)
    assert op.kernel_variant == 2
    assert rel_l2(host(op.apply_unmasked(dev(u))), ref) < 1e-12
    # device geometry (tier T2)
    op2 = mngr.poisson_operator(mode="box", elems_per_patch=pe)
    assert rel_l2(host(op2.apply_unmasked(dev(u))), ref) < 1e-11


def test_box_mode_falls_back_when_the_numbering_is_not_a_lattice():
    # RCM / static-condensation numberings: no tile is a box -> the table-driven kernel
    for sc, rcm in ((False, True), (True, False)):
        mesh, mngr = build_package_case("C", 4, 16, 4, sc, rcm)
        op = mngr.poisson_operator(mode="box", elems_per_patch=16)
        assert op.kernel_variant == 0 and op.box_ld == 0
    # a user-supplied element order never qualifies
    mesh, mngr = build_package_case("C", 4, 16, 4, False, False)
    op = mngr.poisson_operator(mode="box", elems_per_patch=16, elem_order=np.arange(64)[::-1].copy())
    assert op.kernel_variant == 0
    # 4-element patches have no box instantiation
    op = mngr.poisson_operator(mode="box", elems_per_patch=4)
    assert op.kernel_variant == 0
    # the automatic choice on a lattice is the box kernel
    assert mngr.poisson_operator(elems_per_patch=16).kernel_variant == 2


def test_box_kernel_at_size_512x512_p8():
    """Properties at a size where the persistent grid wraps many times: equality with the
    table-driven kernel, symmetry, constants in the null space."""
    mesh, mngr = build_package_case("C", 512, 512, 8, False, False)
    n = mngr.ndof
    box = mngr.poisson_operator(mode="box")
    col = mngr.poisson_operator(mode="column")
    assert box.kernel_variant == 2 and box.elems_per_patch == 16
    g = torch.Generator(device="cuda").manual_seed(3)
    u = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    v = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    Au = box.apply_unmasked(u)
    assert torch.equal(Au, col.apply_unmasked(u))
    a, b = float(torch.dot(v, Au)), float(torch.dot(u, box.apply_unmasked(v)))
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))
    one = torch.ones(n, dtype=torch.float64, device="cuda")
    assert float(box.apply_unmasked(one).abs().max()) <= 1e-9 * float(box.diagonal(masked=False).max())


def test_batched_host_apply_equals_device_apply():
    """semk_poisson_apply_host_batch_f64: several applies on pinned host buffers in one call,
    two device scratch sets used alternately -- every result bit-identical to the
    device-resident apply; more applies than scratch sets exercises the reuse ordering."""
    mesh, mngr = build_package_case("C", 12, 40, 6, False, False)
    n = mngr.ndof
    on = np.zeros(n, dtype=bool)
    on[: 40 * 6 + 1] = True
    op = mngr.poisson_operator(dirichlet=on)
    rng = np.random.default_rng(5)
    us = [torch.from_numpy(rng.standard_normal(n)).pin_memory() for _ in range(5)]
    ys = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(5)]
    for stages in (1, 4):
        for y in ys:
            y.fill_(float("nan"))
        op.apply_host_many(us, ys, stages=stages)
        for u, y in zip(us, ys):
            assert torch.equal(y, op.apply(u.cuda()).cpu())
    # numpy buffers and an empty batch are accepted too
    un, yn = rng.standard_normal(n), np.empty(n)
    op.apply_host_many([un], [yn], stages=3)
    assert np.array_equal(yn, op.apply(dev(un)).cpu().numpy())
    assert op.apply_host_many([], []) == []
    with pytest.raises(ValueError):
        op.apply_host_many(us[:2], ys[:1])
