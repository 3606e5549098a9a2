"""The CPU oracle (oracle/sem_oracle.py) against golden vectors frozen from the
live reference (oracle/make_golden.py).  Integer tables bit-exact; FP64 tables
bit-exact; operator / solution quantities within 1e-12 relative L2."""
import numpy as np
import pytest

import sem_oracle as so
from conftest import golden_case_names, load_case, rel_l2

TOL = 1e-12


def test_tables_bit_exact(golden_tables):
    for p in range(1, 11):
        b = so.Basis(p)
        assert np.array_equal(b.nodes, golden_tables["nodes_%d" % p])
        assert np.array_equal(b.bary, golden_tables["bary_%d" % p])
        assert np.array_equal(b.w, golden_tables["quad_%d" % p])
        assert np.array_equal(b.D, golden_tables["D_%d" % p])
        assert np.array_equal(b.E, golden_tables["E_%d" % p])


def test_hier_and_faces(golden_tables):
    for N in (2, 3, 5, 9, 11):
        assert np.array_equal(so.hier_order(N), golden_tables["hier_%d" % N])
    arr = np.arange(2 * 4 * 5).reshape(2, 4, 5)
    for f in range(4):
        assert np.array_equal(so.face_nodes(arr, f), golden_tables["face_%d" % f])


@pytest.mark.parametrize("name", golden_case_names())
def test_case(name):
    g = load_case(name)
    big = g["l2g"].shape[0] > 32
    r = so.run_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"],
                    python_loop_apply=not big)
    # T0: integer tables and the permuted coordinates, bit-exact
    assert np.array_equal(r["l2g"], g["l2g"])
    assert r["l2g"].dtype == np.uint32
    assert np.array_equal(r["nodes"], g["nodes"])
    assert np.array_equal(r["on_ebc"], g["on_ebc"])
    # geometry and operator quantities
    assert rel_l2(r["x_phys"], g["x_phys"]) < TOL
    assert rel_l2(r["invJ"], g["invJ"]) < TOL
    assert rel_l2(r["JxW"], g["JxW"]) < TOL
    assert rel_l2(r["Au"], g["Au"]) < TOL
    assert rel_l2(r["b"], g["b"]) < TOL
    assert rel_l2(r["diag"], g["diag"]) < TOL
    assert rel_l2(r["ebc_vals"], g["ebc_vals"]) < TOL
    assert rel_l2(r["solution"], g["solution"]) < TOL


def test_python_loop_and_batched_apply_agree():
    g = load_case("C534_dm")
    b = so.Basis(g["p"])
    L = so.local_stiffness(b, g["invJ"], g["JxW"])
    y1 = so.apply_dense_local(L, g["l2g"], g["u"])
    y2 = so.apply_dense_batched(L, g["l2g"], g["u"])
    assert rel_l2(y1, y2) < 1e-14
    assert rel_l2(y1, g["Au"]) < TOL


def test_pcg_on_oracle_matrix_matches_direct_solution():
    g = load_case("S448_sc")
    b = so.Basis(g["p"])
    L = so.local_stiffness(b, g["invJ"], g["JxW"])
    n = g["nodes"].shape[1]
    A = so.assemble_csr(L, g["l2g"], n)
    on, vals = g["on_ebc"], g["ebc_vals"]
    free = ~on
    Af = A[free][:, free].tocsr()
    rhs = g["b"][free] - A[free][:, on] @ vals[on]
    x, it = so.pcg_jacobi(Af, rhs, np.zeros(rhs.size), 1e-13, 5000)
    sol = vals.copy()
    sol[free] = x
    assert rel_l2(sol, g["solution"]) < TOL
    assert 100 < it < 400      # 218 iterations at n=4


@pytest.mark.parametrize("name", [n for n in golden_case_names() if "_sc" in n])
def test_condensed_system_reproduces_reference_solution(name):
    """oracle.condensed_system (checker of the device static-condensation path):
    solving its assembled Schur system and back-substituting gives the
    reference's DOFManagerSC.solve result (sem/discrete.py:502-528)."""
    from scipy.sparse.linalg import spsolve
    g = load_case(name)
    c = so.condensed_system(int(g["p"]), g["invJ"], g["JxW"], g["l2g"])
    n_ext = c["n_ext"]
    assert np.allclose(c["S"], np.swapaxes(c["S"], 1, 2), rtol=0, atol=1e-11)
    on = g["on_ebc"][:n_ext]
    assert not g["on_ebc"][n_ext:].any()
    free = ~on
    sol = g["ebc_vals"].copy()
    ext = sol[:n_ext]
    A1 = c["Sg"][free]
    ext[free] = spsolve(A1[:, free].tocsc(), c["grhs"][free] - A1[:, on] @ ext[on])
    inner = np.linalg.solve(c["Aii"], (c["f_int"] - np.einsum("eij,ej->ei", c["Aie"],
                                                              sol[c["ids"]]))[..., None])
    sol[c["int_ids"]] = inner[..., 0]
    assert rel_l2(sol, g["solution"]) < TOL


def test_c_openmp_restatement_matches_python_loop():
    """oracle/sem_oracle_c.c (the all-threads CPU baseline) == the NumPy loop."""
    if so.c_lib() is None:
        pytest.skip("oracle C library not built")
    for name in ("C534_dm", "C888_sc_rcm"):
        g = load_case(name)
        L = so.local_stiffness(so.Basis(g["p"]), g["invJ"], g["JxW"])
        y = so.apply_dense_c(L, g["l2g"], g["u"])
        assert rel_l2(y, g["Au"]) < TOL
        assert rel_l2(y, so.apply_dense_batched(L, g["l2g"], g["u"])) < 1e-13
    assert so.c_threads() >= 1


# permutation-invariant known answers taken from the live reference
# (unmasked operator on u = sin(3x)cos(2y); Poisson solve with the BCs of config 1)
APPENDIX_B_OPERATOR = [
    # kind, nx, ny, p, sc, rcm, |A u|_2, max|A u|, sum b, sum diag, min diag, |u|_2
    ("S", 4, 4, 8, True, True, 4.871362215690856, 0.6613923998438448, 4.0, 7581.257142858502,
     0.6759259259257385, 14.8506520866344),
    ("S", 4, 4, 8, False, False, 4.871362215690856, 0.6613923998438448, 4.0, 7581.257142858502,
     0.6759259259257385, 14.8506520866344),
    ("C", 4, 4, 8, True, False, 4.951536253550776, 0.6754227080246211, 4.0, 7831.9289921098,
     0.6711123657391915, 14.81396698424245),
    ("C", 5, 3, 4, False, False, 4.273803547745486, 1.0875759595339827, 4.0, 1716.7979340362976,
     0.7770480519335237, 7.276314819816275),
    ("C", 3, 3, 10, False, False, 5.863015326611709, 0.9877851353604654, 4.0, 7305.133785417742,
     0.6685660077924067, 13.89541775044856),
]
APPENDIX_B_SOLVE = [
    # kind, nx, ny, p, rcm, n EBC nodes, |u|_2, sum u, max u
    ("S", 4, 4, 8, False, 65, 28.98108681850342, 857.0220183132233, 1.4384900658533568),
    ("S", 4, 4, 8, True, 65, 28.98108681850335, 857.0220183132216, 1.4384900658533508),
    ("C", 4, 4, 8, False, 65, 28.874703751175932, 853.1756357787481, 1.4384900658551993),
    ("C", 5, 3, 4, False, 33, 14.395512215342288, 211.2447623761652, 1.438491628298089),
]


@pytest.mark.parametrize("row", APPENDIX_B_OPERATOR)
def test_survey_known_answers_operator(row):
    kind, nx, ny, p, sc, rcm, nAu, mAu, sb, sd, md, nu = row
    r = so.run_case(kind, nx, ny, p, sc, rcm, solve=False)
    rel = lambda a, b: abs(a - b) / abs(b)          # noqa: E731
    assert rel(np.linalg.norm(r["Au"]), nAu) < 1e-12
    assert rel(np.abs(r["Au"]).max(), mAu) < 1e-11
    assert rel(r["b"].sum(), sb) < 1e-13
    assert rel(r["diag"].sum(), sd) < 1e-12
    assert rel(r["diag"].min(), md) < 1e-11
    assert rel(np.linalg.norm(r["u"]), nu) < 1e-13


@pytest.mark.parametrize("row", APPENDIX_B_SOLVE)
def test_survey_known_answers_solve(row):
    kind, nx, ny, p, rcm, n_ebc, nu, su, mu = row
    r = so.run_case(kind, nx, ny, p, True, rcm)
    assert int(r["on_ebc"].sum()) == n_ebc
    u = r["solution"]
    assert abs(np.linalg.norm(u) - nu) / nu < 1e-12
    assert abs(u.sum() - su) / su < 1e-12
    assert abs(u.max() - mu) / mu < 1e-11
