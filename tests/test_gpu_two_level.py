"""Two-level preconditioner of the condensed system on the device (Jacobi + vertex
coarse space; csrc/semk_sc.cu coarse kernels, semk_sc_pcg2_solve_f64) against the
golden solutions of the live reference, the NumPy emulation on the same host tables
(tests/test_two_level_host.py) and the Jacobi-PCG path."""
import numpy as np
import pytest
import torch

import sem_oracle as so
from conftest import build_package_case, golden_case_names, load_case, rel_l2
from spectralelementmethod_b200 import discrete, meshgen
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS
from test_two_level_host import emulate

pytestmark = pytest.mark.gpu

SC_CASES = [n for n in golden_case_names() if "_sc" in n]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("name", SC_CASES)
def test_two_level_solve_vs_reference_golden(name):
    g = load_case(name)
    mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"])
    sc = mngr.condensed_poisson_operator(dirichlet=g["on_ebc"])
    u, info = sc.solve(1.0, g["ebc_vals"], rtol=1e-13, preconditioner="two-level")
    assert info.converged, info
    assert rel_l2(host(u), g["solution"]) < 1e-12
    uj, info_j = sc.solve(1.0, g["ebc_vals"], rtol=1e-13)
    assert info.iterations <= info_j.iterations
    assert sc.last_inner_iterations > 0


def test_coarse_operator_and_iteration_count_match_the_emulation():
    g = load_case("C888_sc_rcm")
    p = int(g["p"])
    mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], p, g["sc"], g["rcm"])
    sc = mngr.condensed_poisson_operator(dirichlet=g["on_ebc"],
                                         geometric_factors=(g["invJ"], g["JxW"]))
    c = so.condensed_system(p, g["invJ"], g["JxW"], g["l2g"])
    xj, itj, x2, it2, ct, Ace, _ = emulate(p, c, g["on_ebc"], g["ebc_vals"], so.Basis(p).nodes)
    cs, t, n_v = sc._build_coarse()
    assert n_v == ct["n_v"]
    assert rel_l2(host(t["Ace"]).reshape(-1, 4, 4), Ace) < 1e-12
    xc = np.random.default_rng(2).standard_normal(n_v)
    vc = ct["vert_c"].astype(np.int64)
    yl = np.einsum("eac,ec->ea", Ace, xc[vc]).ravel()
    want = np.add.reduceat(yl[ct["vpos"].astype(np.int64)], ct["vptr"].astype(np.int64)[:-1])
    want = np.where(ct["dirichlet_c"], xc, want)
    dot = torch.zeros(1, dtype=torch.float64, device="cuda")
    got = host(sc.coarse_apply(dev(xc), dot_out=dot))
    assert rel_l2(got, want) < 1e-12
    assert abs(float(dot.item()) - xc @ want) < 1e-10 * abs(xc @ want)
    b = sc.lift(sc.rhs(1.0), g["ebc_vals"])
    x, info = sc.solve_pcg(b, rtol=1e-12, preconditioner="two-level")
    assert info.converged and abs(info.iterations - it2) <= 3
    assert rel_l2(host(x), x2) < 1e-10
    with pytest.raises(ValueError):
        sc.solve_pcg(b, preconditioner="multigrid")


def test_two_level_at_size_and_on_an_unstructured_mesh():
    mesh, mngr = build_package_case("C", 96, 96, 8, True, False)
    on = mngr.boundary_node_mask("ebc")
    x, y = mesh.nodes
    vals = np.where(on, 0.2 * ((x + 1) + (y + 1)), 0.0)
    sc = mngr.condensed_poisson_operator(dirichlet=on)
    u2, info2 = sc.solve(1.0, vals, rtol=1e-12, preconditioner="two-level")
    uj, infoj = sc.solve(1.0, vals, rtol=1e-12)
    assert info2.converged and infoj.converged
    assert info2.iterations < 60 and info2.iterations * 10 < infoj.iterations
    assert rel_l2(host(u2), host(uj)) < 1e-9
    full = mngr.poisson_operator(dirichlet=on)
    b = full.lift(full.rhs(1.0), vals)
    assert float((b - full.apply(u2)).norm() / b.norm()) < 1e-10
    # irregular vertex valence
    pm = meshgen.pinwheel_mesh(7, 4, rings=3)
    b1 = LagrangeGaussLobatto(4)
    pmngr = discrete.DOFManagerSC(pm, 1, TensorProductQS(b1, b1), rcm_order=True)
    pon = pmngr.boundary_node_mask("ebc")
    px, py = pm.nodes
    pvals = np.where(pon, 0.3 * px - 0.2 * py + 0.1, 0.0)
    psc = pmngr.condensed_poisson_operator(dirichlet=pon)
    a2, i2 = psc.solve(1.0, pvals, rtol=1e-13, preconditioner="two-level")
    aj, ij = psc.solve(1.0, pvals, rtol=1e-13)
    assert i2.converged and rel_l2(host(a2), host(aj)) < 1e-10
