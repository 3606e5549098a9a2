"""The CPU oracle's restatement of the axisymmetric Stokes / Navier-Stokes recipe
(oracle/sem_oracle.py: stokes_*) against golden vectors frozen from the reference's OWN
example class run live (oracle/make_golden_stokes.py, examples/squirmer-axisymmetric.py).
Mesh tables bit-exact; operators, local systems and Newton solutions within 1e-12."""
import glob
import os

import numpy as np
import pytest

import sem_oracle as so
from conftest import GOLDEN as GOLDEN_DIR, rel_l2

TOL = 1e-12
STOKES = sorted(os.path.basename(f)[len("stokes_"):-4]
                for f in glob.glob(os.path.join(GOLDEN_DIR, "stokes_*.npz")))


def load(name):
    with np.load(os.path.join(GOLDEN_DIR, "stokes_%s.npz" % name)) as z:
        return {k: z[k] for k in z.files}


def finite_rel(a, b):
    """relative max difference over the finite entries; the non-finite patterns (JxW/rho
    on the axis of symmetry) must coincide."""
    m = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), m)
    return np.abs(a[m] - b[m]).max() / np.abs(b[m]).max()


def oracle_setup(g):
    nr, nt, p = int(g["nr"]), int(g["nt"]), int(g["p"])
    nodes = so.annulus_nodes(nr, nt, p, float(g["r_out"]))
    l2g = so.mesh_l2g(nr, nt, p)
    nodes, l2g, n_ext = so.static_condensation(nodes, l2g)
    nodes, l2g = so.rcm_exterior(nodes, l2g, n_ext)
    basis = so.Basis(p)
    geo = so.geometry(basis, nodes, l2g)
    ops = so.stokes_local_operators(basis, geo["x_phys"], geo["invJ"], geo["JxW"])
    return nodes, l2g, n_ext, ops


def test_golden_files_present():
    assert len(STOKES) >= 4


@pytest.mark.parametrize("name", STOKES)
def test_stokes_operators_and_local_systems(name):
    g = load(name)
    nodes, l2g, n_ext, ops = oracle_setup(g)
    assert np.array_equal(l2g, g["l2g"]) and l2g.dtype == np.uint32
    assert np.array_equal(nodes, g["nodes"])
    assert 2 * n_ext == int(g["ndof_exterior"])
    keep = g["keep"]
    assert finite_rel(ops["E2e"][keep], g["E2e"]) < TOL
    assert finite_rel(ops["Lve"][keep], g["Lve"]) < TOL
    assert finite_rel(ops["Me"][keep], g["Me"]) < TOL
    sfn, vort = g["perturbed"][0::2], g["perturbed"][1::2]
    jac, rhs = so.stokes_local_system(ops, float(g["n_rey"]), sfn[l2g], vort[l2g])
    assert finite_rel(jac[keep], g["jac"]) < TOL
    assert finite_rel(rhs[keep], g["rhs"]) < TOL
    assert finite_rel(rhs, g["rhs_all"]) < TOL


@pytest.mark.parametrize("name", STOKES)
def test_stokes_newton_reproduces_reference_solution(name):
    g = load(name)
    nodes, l2g, n_ext, ops = oracle_setup(g)
    x = g["state0"].copy()
    for it in range(20):
        jac, rhs = so.stokes_local_system(ops, float(g["n_rey"]), x[0::2][l2g], x[1::2][l2g])
        d = so.stokes_newton_step(jac, rhs, l2g, n_ext, g["dof_mask"], g["cint"])
        x += d
        if np.linalg.norm(d[1::2]) < 1e-10:
            break
    assert it < 19
    assert rel_l2(x, g["solution"]) < 1e-11


@pytest.mark.parametrize("name", STOKES)
def test_global_jacobian_matches_schur_path(name):
    """The assembled uncondensed Jacobian (what the device apply is compared with) gives
    the same Newton increment as the example's Schur-complement path."""
    from scipy.sparse.linalg import spsolve
    g = load(name)
    nodes, l2g, n_ext, ops = oracle_setup(g)
    x = g["state0"]
    jac, rhs = so.stokes_local_system(ops, float(g["n_rey"]), x[0::2][l2g], x[1::2][l2g])
    d_ref = so.stokes_newton_step(jac, rhs, l2g, n_ext, g["dof_mask"], g["cint"])
    J = so.stokes_global_jacobian(jac, l2g)
    n = J.shape[0]
    unk = np.ones(n, dtype=bool)
    unk[:2 * n_ext] = g["dof_mask"]
    b = np.zeros(n)
    b[:2 * n_ext] = g["cint"]
    E, nn = l2g.shape[0], l2g.shape[1] * l2g.shape[2]
    gid = (2 * l2g.reshape(E, nn).astype(np.int64)[:, :, None] + np.arange(2)).reshape(E, 2 * nn)
    np.add.at(b, gid.ravel(), np.where(np.isfinite(rhs), rhs, 0.0).ravel())
    d = np.zeros(n)
    d[unk] = spsolve(J[unk][:, unk].tocsc(), b[unk])
    assert rel_l2(d, d_ref) < 1e-9
