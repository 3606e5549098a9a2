"""Parity of the device static-condensation path (csrc/semk_sc.cu, SURVEY.md
8(f) row 1) with the reference's DOFManagerSC formulation
(sem/discrete.py:404-528).

Anchors: the golden solutions frozen from the live reference's
``DOFManagerSC.solve`` (tests/golden/case_*_sc*.npz), the CPU oracle's local
Schur complements / condensed system built from the reference's own invJ and
detJxW (tier T1) and from the device geometry (tier T2), and size-independent
properties at a larger size.  Tolerance: 1e-12 relative L2 (BASELINE.json).
"""
import numpy as np
import pytest
import torch

import sem_oracle as so
from conftest import build_package_case, golden_case_names, load_case, rel_l2
from spectralelementmethod_b200 import _lib, discrete, meshgen
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS

pytestmark = pytest.mark.gpu

TOL = 1e-12
SC_CASES = [n for n in golden_case_names() if "_sc" in n]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def oracle_condensed(g):
    """The condensed system of a golden case from the reference's own invJ / detJxW
    (oracle/sem_oracle.py:condensed_system, pinned in tests/test_oracle_golden.py)."""
    return so.condensed_system(int(g["p"]), g["invJ"], g["JxW"], g["l2g"])


def condensed_operator(name, tier):
    g = load_case(name)
    mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"])
    kw = {}
    if tier == "T1":
        kw["geometric_factors"] = (g["invJ"], g["JxW"])
    sc = mngr.condensed_poisson_operator(dirichlet=g["on_ebc"], **kw)
    return g, mngr, sc


@pytest.mark.parametrize("name", SC_CASES)
@pytest.mark.parametrize("tier", ["T1", "T2"])
def test_local_schur_and_condensed_system_vs_oracle(name, tier):
    g, mngr, sc = condensed_operator(name, tier)
    ref = oracle_condensed(g)
    assert sc.n_ext == ref["n_ext"] == mngr.ndof_exterior
    assert np.array_equal(sc.l2g_ext_host.astype(np.int64), ref["ids"])
    S = sc.local_schur()
    assert S.shape == ref["S"].shape
    assert rel_l2(S, ref["S"]) < TOL
    rng = np.random.default_rng(0)
    u = rng.standard_normal(sc.n_ext)
    dot = torch.zeros(1, dtype=torch.float64, device="cuda")
    y = host(sc.apply(dev(u), flags=0, dot_out=dot))
    want = ref["Sg"] @ u
    assert rel_l2(y, want) < TOL
    assert abs(float(dot.item()) - u @ want) <= 1e-11 * abs(u @ want) + 1e-11
    # Dirichlet elimination: Shat = M S M + (I - M)
    on = g["on_ebc"][:sc.n_ext]
    uf = np.where(on, 0.0, u)
    want_m = np.where(on, u, ref["Sg"] @ uf)
    assert rel_l2(host(sc.apply(dev(u))), want_m) < TOL
    assert rel_l2(host(sc.diagonal(masked=False)), ref["Sg"].diagonal()) < TOL
    assert rel_l2(host(sc.rhs(1.0)), ref["grhs"]) < TOL
    # bit-reproducible
    y2 = host(sc.apply(dev(u), flags=0))
    assert np.array_equal(y, y2)


@pytest.mark.parametrize("name", SC_CASES)
@pytest.mark.parametrize("tier", ["T1", "T2"])
def test_condensed_solve_vs_reference_golden(name, tier):
    g, mngr, sc = condensed_operator(name, tier)
    u, info = sc.solve(1.0, g["ebc_vals"], rtol=1e-13)
    assert info.converged, info
    assert u.numel() == mngr.ndof
    assert rel_l2(host(u), g["solution"]) < TOL


def test_condensed_solve_matches_uncondensed_pcg_with_nodal_load():
    mesh, mngr = build_package_case("C", 12, 10, 6, True, True)
    x, y = mesh.nodes
    on = mngr.boundary_node_mask("ebc")
    vals = np.where(on, 0.2 * ((x + 1) + (y + 1)), 0.0)
    f = 1.0 + np.sin(2 * x) * np.cos(3 * y)
    sc = mngr.condensed_poisson_operator(dirichlet=on)
    u_sc, info_sc = sc.solve(f, vals, rtol=1e-13)
    full = mngr.poisson_operator(dirichlet=on)
    u_full, info_full = full.solve(dev(f), vals, rtol=1e-13)
    assert info_sc.converged and info_full.converged
    assert rel_l2(host(u_sc), host(u_full)) < 1e-11
    assert info_sc.iterations < info_full.iterations
    # the condensed solution satisfies the FULL system: true residual of Ahat u = bhat
    b = full.lift(full.rhs(dev(f)), vals)
    r = b - full.apply(u_sc)
    assert float(r.norm() / b.norm()) < 1e-11


def test_condensed_properties_at_size():
    # 96 x 96 elements of order 8: 590 k DOF, 147 k exterior DOF
    mesh, mngr = build_package_case("C", 96, 96, 8, True, False)
    on = mngr.boundary_node_mask("ebc")
    sc = mngr.condensed_poisson_operator(dirichlet=on)
    n = sc.n_ext
    rng = np.random.default_rng(1)
    u, v = dev(rng.standard_normal(n)), dev(rng.standard_normal(n))
    Su, Sv = sc.apply_unmasked(u), sc.apply_unmasked(v)
    assert abs(float(v @ Su - u @ Sv)) <= 1e-11 * float(Su.norm() * v.norm())      # symmetric
    ones = sc.new_vector(1.0)
    assert float(sc.apply_unmasked(ones).abs().max()) <= 1e-10 * float(Su.abs().max())  # S 1 = 0
    lin = sc.apply_unmasked(2.0 * u - 3.0 * v)
    assert float((lin - (2.0 * Su - 3.0 * Sv)).norm() / lin.norm()) < 1e-13
    assert torch.equal(sc.apply_unmasked(u), Su)                                   # deterministic
    x_nodes, y_nodes = mesh.nodes
    vals = np.where(on, 0.2 * ((x_nodes + 1) + (y_nodes + 1)), 0.0)
    sol, info = sc.solve(1.0, vals, rtol=1e-12)
    assert info.converged
    full = mngr.poisson_operator(dirichlet=on)
    b = full.lift(full.rhs(1.0), vals)
    r = b - full.apply(sol)
    assert float(r.norm() / b.norm()) < 1e-10


def test_condensed_operator_errors():
    mesh, mngr = build_package_case("S", 3, 2, 4, False, False)
    with pytest.raises(ValueError):                       # needs DOFManagerSC
        from spectralelementmethod_b200.condensed import CondensedPoissonOperator
        CondensedPoissonOperator(mngr)
    mesh, mngr = build_package_case("S", 3, 2, 4, True, False)
    bad = np.zeros(mngr.ndof, dtype=bool)
    bad[-1] = True                                        # an element-interior node
    with pytest.raises(ValueError):
        mngr.condensed_poisson_operator(dirichlet=bad)
    mesh = meshgen.structured_quad_mesh(2, 2, 12)
    b1 = LagrangeGaussLobatto(12)
    mngr = discrete.DOFManagerSC(mesh, 1, TensorProductQS(b1, b1), rcm_order=False)
    with pytest.raises(NotImplementedError):              # the reference's range: order <= 10
        mngr.condensed_poisson_operator()
    sc = build_package_case("S", 3, 2, 4, True, False)[1].condensed_poisson_operator()
    with pytest.raises(ValueError):
        sc.apply(torch.zeros(3, dtype=torch.float64, device="cuda"))
    lib = _lib.load()
    assert lib.semk_sc_apply_f64(None, None, None, 0, None, None) == _lib.ERR_INVALID


@pytest.mark.parametrize("n_cells,p,rings,rcm", [(5, 3, 1, False), (3, 5, 1, True), (7, 4, 2, True),
                                                 (6, 8, 3, True), (5, 10, 2, False)])
def test_condensed_on_unstructured_meshes_vs_oracle(n_cells, p, rings, rcm):
    """Irregular vertex valence (3, 5, 6, 7 cells around a node): the node -> entries
    lists have any length.  Local Schur complements, condensed apply and the full
    solve against the oracle's Schur path (sem/discrete.py:404-528) on the same mesh."""
    mesh = meshgen.pinwheel_mesh(n_cells, p, rings=rings)
    b1 = LagrangeGaussLobatto(p)
    mngr = discrete.DOFManagerSC(mesh, 1, TensorProductQS(b1, b1), rcm_order=rcm)
    on = mngr.boundary_node_mask("ebc")
    l2g = mngr.node_map_array()
    geo = so.geometry(so.Basis(p), mesh.nodes, l2g)
    ref = so.condensed_system(p, geo["invJ"], geo["JxW"], l2g)
    sc = mngr.condensed_poisson_operator(dirichlet=on)
    assert sc.n_ext == ref["n_ext"] == mngr.ndof_exterior
    assert rel_l2(sc.local_schur(), ref["S"]) < 1e-11
    u = np.random.default_rng(3).standard_normal(sc.n_ext)
    assert rel_l2(host(sc.apply_unmasked(dev(u))), ref["Sg"] @ u) < 1e-11
    assert rel_l2(host(sc.rhs(1.0)), ref["grhs"]) < 1e-11
    x, y = mesh.nodes
    vals = np.where(on, 0.3 * x - 0.2 * y + 0.1, 0.0)
    L = so.local_stiffness(so.Basis(p), geo["invJ"], geo["JxW"])
    want = so.solve_schur(L, geo["JxW"], l2g, sc.n_ext, on, vals)
    sol, info = sc.solve(1.0, vals, rtol=1e-13)
    assert info.converged and rel_l2(host(sol), want) < 1e-10


def test_condensed_weighted_stiffness_vs_oracle():
    """rho-weighted twin (examples/squirmer-axisymmetric.py:194-213) through the
    condensed path: weight w = 2 + x at the GLL points."""
    g = load_case("C448_sc_rcm")
    mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"])
    w = 2.0 + g["x_phys"][:, 0]
    ref = so.condensed_system(int(g["p"]), g["invJ"], g["JxW"] * w, g["l2g"])
    # (the oracle's load would be weighted too; only the operator is compared)
    sc = mngr.condensed_poisson_operator(weight=lambda x, y: 2.0 + x)
    u = np.random.default_rng(4).standard_normal(sc.n_ext)
    assert rel_l2(host(sc.apply(dev(u))), ref["Sg"] @ u) < TOL
    sc2 = mngr.condensed_poisson_operator(weight=w)
    assert rel_l2(sc2.local_schur(), ref["S"]) < TOL


@pytest.mark.parametrize("name", SC_CASES[:4])
def test_device_built_integer_tables_equal_the_host_builders(name):
    """condensed_tables_device / coarse_tables_device (stable sorts on the GPU) give the
    very tables of the NumPy builders (T0 tier: integer tables bit-exact, weights equal)."""
    from spectralelementmethod_b200.condensed import coarse_tables, condensed_tables
    g, mngr, sc = condensed_operator(name, "T2")
    N = int(g["p"]) + 1
    ext_loc = sc.ext_loc_host
    l2g = mngr.node_map_array().reshape(-1, N * N)
    l2g_ext, nptr, npos = condensed_tables(l2g, ext_loc, sc.n_ext)
    u32 = lambda t: host(t).view(np.uint32)        # noqa: E731
    assert np.array_equal(u32(sc._t["l2g_ext"]), l2g_ext)
    assert np.array_equal(u32(sc._t["node_ptr"]), nptr)
    assert np.array_equal(u32(sc._t["node_pos"]), npos)
    b1 = LagrangeGaussLobatto(int(g["p"]))
    ct = coarse_tables(l2g_ext, nptr, npos, sc.dirichlet_host, np.asarray(b1.nodes))
    cs, t, n_v = sc._build_coarse()
    assert n_v == ct["n_v"]
    for key in ("vert_c", "vptr", "vpos", "pv", "rptr", "ridx"):
        assert np.array_equal(u32(t[key]).reshape(ct[key].shape), ct[key]), key
    assert np.array_equal(host(t["pw"]), ct["pw"]) and np.array_equal(host(t["rw"]), ct["rw"])
    assert np.array_equal(host(t["phi"]), ct["phi"])
    assert np.array_equal(t["dirichlet_c_host"], ct["dirichlet_c"])


@pytest.mark.parametrize("name", SC_CASES[:3])
def test_stored_interior_operator_backsolve_equals_the_refactorising_one(name):
    """W_e = A_ii^-1 A_ie kept by the Schur pass + the interior solution of the load kept by
    rhs(): the streaming back-substitution gives the refactorising kernel's result, for a
    scalar and for a nodal load, and falls back when the load differs."""
    g = load_case(name)
    mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"])
    stored = mngr.condensed_poisson_operator(dirichlet=g["on_ebc"], store_interior=True)
    plain = mngr.condensed_poisson_operator(dirichlet=g["on_ebc"], store_interior=False)
    assert stored._W is not None and plain._W is None
    rng = np.random.default_rng(2)
    fn = dev(rng.standard_normal(stored.n_nodes))
    for f in (1.0, fn):
        b = stored.lift(stored.rhs(f), g["ebc_vals"])
        assert stored._c is not None
        x, info = stored.solve_pcg(b, rtol=1e-13)
        fast = stored.backsolve(x, f)
        slow = plain.backsolve(x, f)
        assert rel_l2(host(fast), host(slow)) < 1e-12
    # a different load than the cached one: the refactorising path is taken, same answer
    other = stored.backsolve(x, 2.0)
    assert rel_l2(host(other), host(plain.backsolve(x, 2.0))) < 1e-13
    u, info = stored.solve(1.0, g["ebc_vals"], rtol=1e-13)
    assert rel_l2(host(u), g["solution"]) < TOL


def test_stored_interior_inverse_load_equals_element_pass():
    """store_interior_inverse: the condensed load and the back-substitution from the stored
    W / A_ii^-1 (streaming kernels) against the refactorising element pass, for a scalar and a
    nodal load, with a weight and a reaction term."""
    mesh, mngr = build_package_case("C", 9, 7, 6, True, True)
    on = mngr.boundary_node_mask("ebc")
    rng = np.random.default_rng(8)
    NN = 49
    reaction = rng.uniform(0.0, 2.0, size=(mesh.n_cells, NN))
    kw = dict(dirichlet=on, weight=lambda x, y: 1.5 + x, reaction=reaction)
    a = mngr.condensed_poisson_operator(store_interior=True, store_interior_inverse=True, **kw)
    b = mngr.condensed_poisson_operator(store_interior=False, **kw)
    assert a._Ainv is not None and b._W is None
    f = dev(rng.standard_normal(mesh.n_nodes))
    for load in (1.0, f):
        ga, gb = a.rhs(load), b.rhs(load)
        assert rel_l2(host(ga), host(gb)) < 1e-12
        x = dev(rng.standard_normal(a.n_ext))
        assert rel_l2(host(a.backsolve(x, load)), host(b.backsolve(x, load))) < 1e-11
    ua, ia = a.solve(f, None, rtol=1e-13)
    ub, ib = b.solve(f, None, rtol=1e-13)
    assert ia.converged and ib.converged and rel_l2(host(ua), host(ub)) < 1e-10
