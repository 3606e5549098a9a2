"""Parity of the device axisymmetric Stokes / Navier-Stokes path (csrc/semk_stokes.cu,
spectralelementmethod_b200/stokes.py) with the reference's example
(examples/squirmer-axisymmetric.py).

Anchors: golden vectors frozen from the example's OWN class run live
(tests/golden/stokes_*.npz, oracle/make_golden_stokes.py) and the CPU oracle's restatement
of its operators (oracle/sem_oracle.py: stokes_*, pinned in tests/test_oracle_stokes.py),
fed the oracle's geometry (tier T1) or the device geometry (tier T2); size-independent
properties at a larger size.  Tolerance: 1e-12 relative L2 on the apply / residual;
the Newton-GMRES solutions against the reference's direct solves to 1e-8.
"""
import glob
import os

import numpy as np
import pytest
import torch

import sem_oracle as so
from conftest import GOLDEN, rel_l2
from spectralelementmethod_b200 import discrete, meshgen, stokes
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS

pytestmark = pytest.mark.gpu

TOL = 1e-12
CASES = sorted(os.path.basename(f)[len("stokes_"):-4]
               for f in glob.glob(os.path.join(GOLDEN, "stokes_*.npz")))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def load(name):
    with np.load(os.path.join(GOLDEN, "stokes_%s.npz" % name)) as z:
        return {k: z[k] for k in z.files}


def manager(nr, nt, p, r_out):
    mesh = meshgen.annulus_sector_mesh(nr, nt, p, r_out)
    b1 = LagrangeGaussLobatto(p)
    return mesh, discrete.DOFManagerSC(mesh, 2, TensorProductQS(b1, b1))


def slip_of(g):
    if str(g["kind"]) == "Squirmer":
        return float(g["speed"]), stokes.squirmer_vslip_profile(float(g["beta"]))
    return 1.0, stokes.zero_slip_vel


def oracle_system(g, l2g, nodes, state):
    basis = so.Basis(int(g["p"]))
    geo = so.geometry(basis, nodes, l2g)
    ops = so.stokes_local_operators(basis, geo["x_phys"], geo["invJ"], geo["JxW"])
    jac, rhs = so.stokes_local_system(ops, float(g["n_rey"]), state[0::2][l2g], state[1::2][l2g])
    return geo, jac, rhs


def assemble_local(rhs, l2g):
    E, nn = l2g.shape[0], l2g.shape[1] * l2g.shape[2]
    gid = (2 * l2g.reshape(E, nn).astype(np.int64)[:, :, None] + np.arange(2)).reshape(E, 2 * nn)
    out = np.zeros(2 * (int(l2g.max()) + 1))
    np.add.at(out, gid.ravel(), np.where(np.isfinite(rhs), rhs, 0.0).ravel())
    return out


@pytest.mark.parametrize("name", CASES)
def test_mesh_tables_and_boundary_data_vs_reference(name):
    g = load(name)
    mesh, dm = manager(int(g["nr"]), int(g["nt"]), int(g["p"]), float(g["r_out"]))
    assert np.array_equal(mesh.nodes, g["nodes"])
    assert np.array_equal(mesh.node_map_array().reshape(g["l2g"].shape), g["l2g"])
    assert dm.ndof == 2 * mesh.n_nodes and dm.ndof_exterior == int(g["ndof_exterior"])
    speed, slip = slip_of(g)
    bc = stokes.squirmer_boundary_data(dm, speed, slip)
    n_ext = int(g["ndof_exterior"])
    assert np.array_equal(~bc.essential[:n_ext], g["dof_mask"])     # bit-exact mask
    assert not bc.essential[n_ext:].any()
    assert rel_l2(bc.state0, g["state0"]) < TOL
    assert np.abs(bc.cint[:n_ext] - g["cint"]).max() <= TOL * max(np.abs(g["cint"]).max(), 1.0)
    assert not bc.cint[n_ext:].any()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tier", ["T1", "T2"])
def test_jacobian_apply_and_residual_vs_oracle(name, tier):
    g = load(name)
    mesh, dm = manager(int(g["nr"]), int(g["nt"]), int(g["p"]), float(g["r_out"]))
    l2g = g["l2g"]
    state = g["perturbed"]
    geo, jac, rhs = oracle_system(g, l2g, g["nodes"], state)
    kw = {}
    if tier == "T1":
        kw["geometric_factors"] = (geo["invJ"], geo["JxW"], geo["x_phys"])
    op = dm.axisymmetric_stokes_operator(n_rey=float(g["n_rey"]), **kw)
    J = so.stokes_global_jacobian(jac, l2g)
    tol = TOL if tier == "T1" else 20 * TOL
    # residual: the golden local residuals of the example, assembled
    res = host(op.residual(dev(state)))
    ref = -assemble_local(g["rhs_all"], l2g)
    ess = ~np.isfinite(assemble_local_mask(g["rhs_all"], l2g))
    assert rel_l2(res[~ess], ref[~ess]) < tol
    # Jacobian about the same state, random direction, columns / rows on the axis left out
    # (JxW / rho is infinite there in the reference and eliminated by the BCs)
    rng = np.random.default_rng(3)
    u = rng.standard_normal(J.shape[0])
    axis = np.repeat(g["nodes"][0] == 0.0, 2)
    u[axis] = 0.0
    y = host(op.apply_unmasked(dev(u)))
    yr = J @ u
    assert rel_l2(y[~axis], yr[~axis]) < tol
    # linearity and bitwise determinism
    v = rng.standard_normal(J.shape[0])
    a = host(op.apply_unmasked(dev(2.0 * u - 0.5 * v)))
    b = 2.0 * y - 0.5 * host(op.apply_unmasked(dev(v)))
    assert rel_l2(a, b) < 1e-13
    assert np.array_equal(host(op.apply_unmasked(dev(u))), y)


def assemble_local_mask(rhs, l2g):
    """NaN at the DOFs that receive a non-finite local entry."""
    E, nn = l2g.shape[0], l2g.shape[1] * l2g.shape[2]
    gid = (2 * l2g.reshape(E, nn).astype(np.int64)[:, :, None] + np.arange(2)).reshape(E, 2 * nn)
    out = np.zeros(2 * (int(l2g.max()) + 1))
    bad = ~np.isfinite(rhs)
    out[gid[bad]] = np.nan
    return out


@pytest.mark.parametrize("name", CASES)
def test_masked_matrix_vs_oracle(name):
    g = load(name)
    if g["l2g"].shape[0] * int(g["p"]) ** 2 > 200:
        pytest.skip("column-by-column matrix only on the smallest cases")
    mesh, dm = manager(int(g["nr"]), int(g["nt"]), int(g["p"]), float(g["r_out"]))
    geo, jac, rhs = oracle_system(g, g["l2g"], g["nodes"], g["state0"])
    n_ext = int(g["ndof_exterior"])
    essential = np.zeros(2 * mesh.n_nodes, dtype=bool)
    essential[:n_ext] = ~g["dof_mask"]
    op = dm.axisymmetric_stokes_operator(n_rey=float(g["n_rey"]), essential=essential)
    op.linearize(dev(g["state0"]))
    A = op.to_scipy_csr(masked=True).toarray()
    J = so.stokes_global_jacobian(jac, g["l2g"]).toarray()
    free = ~essential
    assert np.abs(A[np.ix_(free, free)] - J[np.ix_(free, free)]).max() <= 1e-11 * np.abs(J).max()
    assert np.array_equal(A[essential][:, essential], np.eye(int(essential.sum())))
    assert not A[np.ix_(free, essential)].any() and not A[np.ix_(essential, free)].any()


@pytest.mark.parametrize("name", CASES)
def test_newton_gmres_reproduces_reference_solution(name):
    g = load(name)
    mesh, dm = manager(int(g["nr"]), int(g["nt"]), int(g["p"]), float(g["r_out"]))
    speed, slip = slip_of(g)
    bc = stokes.squirmer_boundary_data(dm, speed, slip)
    op = dm.axisymmetric_stokes_operator(n_rey=float(g["n_rey"]), essential=bc.essential)
    state, hist = op.newton_solve(dev(bc.state0), bc.cint, it_max=20, tol=1e-10,
                                  gmres_rtol=1e-13, restart=400, gmres_maxiter=4000)
    assert all(info.true_rel_residual < 1e-9 for _, info in hist)
    assert rel_l2(host(state), g["solution"]) < 1e-8
    if float(g["n_rey"]) == 0.0:
        assert len(hist) <= 3           # linear problem: one step + convergence checks


def test_block_jacobi_is_the_nodal_block_inverse():
    g = load("fixed_344_re0")
    mesh, dm = manager(int(g["nr"]), int(g["nt"]), int(g["p"]), float(g["r_out"]))
    n_ext = int(g["ndof_exterior"])
    essential = np.zeros(2 * mesh.n_nodes, dtype=bool)
    essential[:n_ext] = ~g["dof_mask"]
    op = dm.axisymmetric_stokes_operator(essential=essential)
    A = op.to_scipy_csr(masked=True).toarray()
    binv = host(op.block_jacobi())
    for nd in range(0, mesh.n_nodes, 7):
        blk = A[2 * nd:2 * nd + 2, 2 * nd:2 * nd + 2]
        assert np.abs(binv[nd].reshape(2, 2) @ blk - np.eye(2)).max() < 1e-10


def test_larger_mesh_apply_vs_oracle():
    """12 x 16 elements of order 8 on the graded annulus (r_out = 100): apply against the
    oracle's assembled Jacobian at a size where nothing is a corner case."""
    nr, nt, p, r_out = 12, 16, 8, 100.0
    mesh, dm = manager(nr, nt, p, r_out)
    l2g = mesh.node_map_array().reshape(-1, p + 1, p + 1)
    g = dict(p=p, n_rey=0.0)
    state = np.zeros(2 * mesh.n_nodes)
    geo, jac, rhs = oracle_system(g, l2g, mesh.nodes, state)
    J = so.stokes_global_jacobian(jac, l2g)
    op = dm.axisymmetric_stokes_operator()
    rng = np.random.default_rng(5)
    u = rng.standard_normal(J.shape[0])
    axis = np.repeat(mesh.nodes[0] == 0.0, 2)
    u[axis] = 0.0
    y = host(op.apply_unmasked(dev(u)))
    yr = J @ u
    assert rel_l2(y[~axis], yr[~axis]) < 1e-11


def test_gmres_true_residual_and_restarts():
    """Restarted GMRES on a 3 x 4, p = 6 squirmer problem: the TRUE residual of the returned
    iterate meets the tolerance, with and without restarts, and agrees with a host sparse
    direct solve of the masked matrix.  (The nodal block-Jacobi preconditioner is not
    mesh-independent: iteration counts grow quickly with the mesh, see DESIGN.md.)"""
    from scipy.sparse.linalg import spsolve
    mesh, dm = manager(3, 4, 6, 20.0)
    bc = stokes.squirmer_boundary_data(dm, 1.0, stokes.squirmer_vslip_profile(1.0))
    op = dm.axisymmetric_stokes_operator(essential=bc.essential)
    rhs = dev(bc.cint) - op.residual(dev(bc.state0))
    x, info = op.solve_gmres(rhs, rtol=1e-10, restart=2000, maxiter=2000)
    assert info.restarts <= 3 and info.true_rel_residual < 1e-9, info
    x2, info2 = op.solve_gmres(rhs, rtol=1e-6, restart=150, maxiter=6000)
    assert info2.true_rel_residual < 1e-5 and info2.restarts > 1, info2
    A = op.to_scipy_csr(masked=True).tocsc()
    b = host(rhs).copy()
    b[bc.essential] = 0.0
    xd = spsolve(A, b)
    assert rel_l2(host(x), xd) < 1e-7


@pytest.mark.parametrize("name", ["fixed_344_re0", "squirmer_238_re0", "fixed_225_re1",
                                  "squirmer_334_re05"])
def test_poisson_block_preconditioner_reproduces_reference_solution(name):
    """GMRES with the block-triangular preconditioner built from the weighted, statically
    condensed Poisson operator (multilevel PCG inside): same solution as the reference's
    direct solve, in a number of iterations that does not grow like the block-Jacobi one."""
    g = load(name)
    nr, nt, p = int(g["nr"]), int(g["nt"]), int(g["p"])
    mesh = meshgen.annulus_sector_mesh(nr, nt, p, float(g["r_out"]))
    b1 = LagrangeGaussLobatto(p)
    dm = discrete.DOFManagerSC(mesh, 2, TensorProductQS(b1, b1), rcm_order=False)
    speed, slip = slip_of(g)
    bc = stokes.squirmer_boundary_data(dm, speed, slip)
    op = dm.axisymmetric_stokes_operator(n_rey=float(g["n_rey"]), essential=bc.essential)
    state, hist = op.newton_solve(dev(bc.state0), bc.cint, it_max=10, tol=1e-9, gmres_rtol=1e-12,
                                  restart=300, gmres_maxiter=300, precondition="poisson")
    assert all(info.true_rel_residual < 1e-8 for _, info in hist)
    assert hist[0][1].iterations <= 200, hist[0][1]
    # the golden solution lives in the RCM numbering of the reference's default manager:
    # match the nodes through their coordinates
    mine = np.lexsort((mesh.nodes[1], mesh.nodes[0]))
    ref = np.lexsort((g["nodes"][1], g["nodes"][0]))
    assert np.array_equal(mesh.nodes[:, mine], g["nodes"][:, ref])
    sol = host(state)
    for comp in (0, 1):
        assert rel_l2(sol[comp::2][mine], g["solution"][comp::2][ref]) < 1e-7


def test_preconditioned_solve_vs_oracle_direct_solve_medium_mesh():
    """8 x 12 elements of order 8 on the graded annulus (r_out = 100, 12.6 k DOF): one Newton
    step at Re = 0 by flexible GMRES + the Poisson block preconditioner against the oracle's
    restatement of the example's Schur-complement / SuperLU step on the same data."""
    nr, nt, p, r_out = 8, 12, 8, 100.0
    mesh = meshgen.annulus_sector_mesh(nr, nt, p, r_out)
    b1 = LagrangeGaussLobatto(p)
    dm = discrete.DOFManagerSC(mesh, 2, TensorProductQS(b1, b1), rcm_order=False)
    bc = stokes.squirmer_boundary_data(dm, 0.9, stokes.squirmer_vslip_profile(-1.0))
    op = dm.axisymmetric_stokes_operator(essential=bc.essential)
    rhs = dev(bc.cint) - op.residual(dev(bc.state0))
    d, info = op.solve_gmres(rhs, rtol=1e-11, restart=300, maxiter=600, precondition="poisson")
    assert info.true_rel_residual < 1e-9 and info.iterations < 200, info
    l2g = mesh.node_map_array().reshape(-1, p + 1, p + 1)
    g = dict(p=p, n_rey=0.0)
    geo, jac, lrhs = oracle_system(g, l2g, mesh.nodes, bc.state0)
    n_ext = dm.ndof_exterior // 2
    dref = so.stokes_newton_step(jac, lrhs, l2g, n_ext, ~bc.essential[:2 * n_ext],
                                 bc.cint[:2 * n_ext])
    got = host(d)
    for comp in (0, 1):
        # (without the row equilibration of solve_gmres the same residual tolerance leaves
        # 1e-3 here: the rows of the system span eight orders of magnitude)
        assert rel_l2(got[comp::2], dref[comp::2]) < 1e-5


def test_swimming_speed_matches_the_number_in_the_reference_docstring():
    """End-to-end known answer: the docstring of the example's `calc_speed`
    (examples/squirmer-axisymmetric.py:661-668) quotes a swimming speed of
    0.92571156681483957 for Re = 1, beta = 1 on meshes/donut.msh -- 15 x 9 transfinite
    elements of order 8 between r = 1 and r = 100 with progression 1.35
    (examples/meshes/donut.geo).  The same problem on the structured stand-in mesh (geometric
    grading 100^(1/15) = 1.359) through the device operators, the Newton / flexible-GMRES
    solver and the force functional reproduces it to 1e-7; the Stokes limit gives 1."""
    mesh = meshgen.annulus_sector_mesh(15, 9, 8, 100.0)
    b1 = LagrangeGaussLobatto(8)
    dm = discrete.DOFManagerSC(mesh, 2, TensorProductQS(b1, b1), rcm_order=False)
    speed, sol, hist = stokes.squirmer_speed(dm, 1.0, 1.0, restart=400, gmres_rtol=1e-10,
                                             gmres_maxiter=1200)
    assert abs(speed - 0.92571156681483957) < 1e-6, (speed, hist)
    assert abs(hist[-1][1]) < 1e-5           # force-free
    speed0, _, _ = stokes.squirmer_speed(dm, 0.0, 1.0, restart=400, gmres_rtol=1e-10,
                                         gmres_maxiter=1200)
    assert abs(speed0 - 1.0) < 1e-5, speed0  # U = 2/3 B1 = 1 for a Stokes squirmer
