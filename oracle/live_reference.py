"""Drive the UNMODIFIED reference (/root/reference/sem) through its current API.

TEST / BASELINE INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py to freeze golden
vectors under tests/golden/, by tests that are skipped when no reference tree is
reachable, and by bench.py's cpu_baseline / --impl reference legs.  No product module
imports this file.  The tree is /root/reference in the development container; on the GPU
box it is oracle/_ref, an offline `pip install --target` of the unmodified reference made
by oracle/install_ref.sh (git-ignored, not gpurun-ignored: it travels with the snapshot).

The reference has bit-rotted against numpy 2 / scipy 1.18 / python 3.12 and
needs h5py (absent).  Five monkey-patch shims (no source edits) make it run:
  1. np.bool / np.float aliases        (sem/discrete.py:151,371; sem/sp_array.py:39)
  2. scipy.special.comb positional arg (sem/geometry.py:149)
  3. fake ``h5py`` serving the GLL table (sem/basis_functions.py:364-370)
  4. open(..., 'rU') for load_msh       (sem/grid_importers.py:56) [not needed here]
  5. matplotlib is never imported.
"""
import os
import sys
import types
import warnings

import numpy as np

def _find_root():
    env = os.environ.get("SEM_REFERENCE_ROOT")
    if env:
        return env
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in ("/root/reference", os.path.join(here, "_ref")):
        if os.path.isdir(os.path.join(cand, "sem")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()

_installed = False


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "sem"))


def install_shims(max_order=16):
    """Install the shims and put the reference on sys.path.  Idempotent."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, ".."))
    from spectralelementmethod_b200 import gll_tables

    if not hasattr(np, "bool"):
        np.bool = bool
    if not hasattr(np, "float"):
        np.float = float

    import scipy.special as sf
    _comb = sf.comb

    def comb(n, k, exact=False, **kw):
        return _comb(n, k, exact=exact, **kw)
    sf.comb = comb

    class _Group(dict):
        attrs = {"max_order": max_order}

    class _File(object):
        def __init__(self, path, mode="r"):
            grp = _Group()
            for p in range(1, max_order + 1):
                grp[str(p)] = np.array(gll_tables.half_table(p), dtype=np.float64)
            self._root = {"GaussLegendreLobatto": grp}

        def __enter__(self):
            return self._root

        def __exit__(self, *a):
            return False

    h5 = types.ModuleType("h5py")
    h5.File = _File
    sys.modules["h5py"] = h5

    sys.path.insert(0, REF_ROOT)
    warnings.filterwarnings("ignore", category=RuntimeWarning)
    _installed = True


# --------------------------------------------------------------------------
# Synthetic meshes (pattern of tests/test_discrete.py:22-38)
# --------------------------------------------------------------------------
def node_coordinates(kind, nx, ny, p):
    NX, NY = nx * p + 1, ny * p + 1
    X, Y = np.meshgrid(np.linspace(-1, 1, NX), np.linspace(-1, 1, NY), indexing="ij")
    if kind == "C":
        s = 0.08 * np.sin(np.pi * X) * np.sin(np.pi * Y)
        X = X + s
        Y = Y + s
    elif kind != "S":
        raise ValueError(kind)
    return np.vstack([X.ravel(), Y.ravel()])


def build_mesh(kind, nx, ny, p):
    """Reference Mesh with 'ebc' = left+bottom, 'nbc' = right+top."""
    install_shims()
    from sem.discrete import Mesh
    from sem.geometry import Quadrilateral
    NX, NY = nx * p + 1, ny * p + 1
    gid = np.arange(NX * NY).reshape(NX, NY)
    mesh = Mesh(2)
    mesh.set_nodes(node_coordinates(kind, nx, ny, p))
    g = mesh.add_geometry(Quadrilateral(p + 1, p + 1))
    r = mesh.new_region("interior")
    ebc = mesh.new_boundary("ebc")
    nbc = mesh.new_boundary("nbc")
    c = 0
    for ex in range(nx):
        for ey in range(ny):
            mesh.add_cell(gid[ex * p:ex * p + p + 1, ey * p:ey * p + p + 1].copy(), g, r)
            if ex == 0:
                mesh.add_boundary_cell(c, ebc, 1, 0)
            if ey == 0:
                mesh.add_boundary_cell(c, ebc, 1, 2)
            if ex == nx - 1:
                mesh.add_boundary_cell(c, nbc, 1, 1)
            if ey == ny - 1:
                mesh.add_boundary_cell(c, nbc, 1, 3)
            c += 1
    return mesh


def make_manager(mesh, p, sc, rcm):
    install_shims()
    from sem.basis_functions import LagrangeGaussLobatto, TensorProductQS
    from sem.discrete import DOFManager, DOFManagerSC
    b1 = LagrangeGaussLobatto(p)
    basis = TensorProductQS(b1, b1)
    cls = DOFManagerSC if sc else DOFManager
    return cls(mesh, 1, basis, rcm_order=rcm)


def local_stiffness(fe):
    """examples/poisson.py:166-193 through the current API."""
    D = fe.basis.get_D1_matrices()
    invJ = fe.invJ
    JxW = fe.detJxW
    N = D[0].shape[0]
    g0 = np.einsum("mp,imn->imnp", D[0], invJ[0])
    g1 = np.einsum("nq,imn->imnq", D[1], invJ[1])
    L = np.zeros((N, N, N, N))
    p, q, r = np.ogrid[0:N, 0:N, 0:N]
    L[p, q, r, q] += np.einsum("mn,imnp,imnr->pnr", JxW, g0, g0)
    L += np.einsum("mn,imnp,imns->pnms", JxW, g0, g1)
    L += np.einsum("mn,imnq,imnr->mqrn", JxW, g1, g0)
    L[p, q, p, r] += np.einsum("mn,imnq,imns->mqs", JxW, g1, g1)
    return L, JxW


def dirichlet_data(mngr, n_dof):
    """examples/poisson.py:125-143 via the current boundary_elements API.
    Returns (on_ebc bool[n_dof], values float64[n_dof])."""
    on_ebc = np.zeros(n_dof, dtype=bool)
    vals = np.zeros(n_dof)
    for fe, bfe in mngr.boundary_elements("ebc", x_phys=True):
        loc = bfe.node_ind
        x, y = bfe.x_phys
        vals[loc] = 0.2 * ((x + 1) + (y + 1))
        on_ebc[loc] = True
    return on_ebc, vals


def run_case(kind, nx, ny, p, sc, rcm, solve=True):
    """Everything the parity tests need from one live-reference run."""
    from scipy import sparse
    mesh = build_mesh(kind, nx, ny, p)
    mngr = make_manager(mesh, p, sc, rcm)
    n = mngr.ndof
    N = p + 1
    E = mesh.n_cells
    l2g = np.empty((E, N, N), dtype=np.uint32)
    hier = np.empty((E, N * N), dtype=np.uint32)
    invJ = np.empty((E, 2, 2, N, N))
    JxW = np.empty((E, N, N))
    xph = np.empty((E, 2, N, N))
    rows, cols, data = [], [], []
    b = np.zeros(n)
    diag = np.zeros(n)
    local_systems = []
    for e, fe in enumerate(mngr.finite_elements(x_phys=True, Jacobian=True)):
        l2g[e] = fe.node_ind
        hier[e] = fe.global_dof_ind_hier
        invJ[e] = fe.invJ
        xph[e] = fe.x_phys
        L, w = local_stiffness(fe)
        JxW[e] = w
        idx = fe.node_ind.ravel().astype(np.int64)
        r, c = np.meshgrid(idx, idx, indexing="ij")
        rows.append(r.ravel()); cols.append(c.ravel()); data.append(L.reshape(-1))
        np.add.at(b, idx, w.ravel())
        np.add.at(diag, idx, np.einsum("pqpq->pq", L).ravel())
        local_systems.append((L.reshape(N * N, N * N), w.reshape(N * N).copy()))
    A = sparse.coo_matrix((np.concatenate(data), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(n, n)).tocsr()
    x, y = mesh.nodes
    u = np.sin(3 * x) * np.cos(2 * y)
    out = dict(l2g=l2g, hier=hier, nodes=mesh.nodes.copy(), invJ=invJ, JxW=JxW, x_phys=xph,
               u=u, Au=A @ u, b=b, diag=diag, A=A)
    on_ebc, vals = dirichlet_data(mngr, n)
    out["on_ebc"] = on_ebc
    out["ebc_vals"] = vals
    if solve:
        if sc:
            lsys = [mngr.reorder_local_system_hier(fe, ls)
                    for fe, ls in zip(mngr.finite_elements(), local_systems)]
            gsys = mngr.init_global_linear_system()
            mngr.assemble_global_sc_system(gsys, lsys)
            sol = vals.copy()
            mngr.solve(gsys, lsys, sol, on_ebc[:mngr.ndof_exterior])
        else:
            from scipy.sparse.linalg import spsolve
            free = ~on_ebc
            sol = vals.copy()
            rhs = b[free] - A[free][:, on_ebc] @ vals[on_ebc]
            sol[free] = spsolve(A[free][:, free].tocsc(), rhs)
        out["solution"] = sol
    return out


# --------------------------------------------------------------------------
# Timing of the UNMODIFIED reference (bench.py cpu_baseline, kind "reference")
# --------------------------------------------------------------------------
def time_apply(kind, nx, ny, p, target_seconds=5.0):
    """The reference's own operator apply (examples/squirmer-axisymmetric.py:268-295 with
    the dense 4-index local stiffness of examples/poisson.py:181-193): a Python loop over
    elements, `y[idx] += einsum('pqrs,rs', Lse, u[idx])`.  Returns (applies, seconds, ndof)."""
    import time
    mesh = build_mesh(kind, nx, ny, p)
    mngr = make_manager(mesh, p, False, False)
    ops = []
    for fe in mngr.finite_elements(x_phys=True, Jacobian=True):
        L, _ = local_stiffness(fe)
        ops.append((fe.node_ind.astype(np.int64), L))
    x, y = mesh.nodes
    u = np.sin(3 * x) * np.cos(2 * y)

    def apply():
        out = np.zeros_like(u)
        for idx, L in ops:
            out[idx] += np.einsum("pqrs,rs", L, u[idx])
        return out
    apply()
    reps, t0 = 0, time.perf_counter()
    while True:
        apply()
        reps += 1
        el = time.perf_counter() - t0
        if el >= target_seconds:
            break
    return reps, el, int(mngr.ndof)


def time_pipeline(kind, nx, ny, p):
    """The reference's whole solver pipeline, once: DOFManagerSC numbering, FiniteElement
    geometry + local stiffness per element, hierarchical reorder, Schur assembly, spsolve,
    interior back-substitution (sem/discrete.py:283-528).  Returns a dict of seconds."""
    import time
    t0 = time.perf_counter()
    mesh = build_mesh(kind, nx, ny, p)
    mngr = make_manager(mesh, p, True, False)
    t1 = time.perf_counter()
    local_systems = []
    for fe in mngr.finite_elements(x_phys=True, Jacobian=True):
        L, w = local_stiffness(fe)
        N = w.shape[0]
        local_systems.append((L.reshape(N * N, N * N), w.reshape(N * N).copy()))
    t2 = time.perf_counter()
    on_ebc, vals = dirichlet_data(mngr, mngr.ndof)
    lsys = [mngr.reorder_local_system_hier(fe, ls)
            for fe, ls in zip(mngr.finite_elements(), local_systems)]
    gsys = mngr.init_global_linear_system()
    mngr.assemble_global_sc_system(gsys, lsys)
    sol = vals.copy()
    mngr.solve(gsys, lsys, sol, on_ebc[:mngr.ndof_exterior])
    t3 = time.perf_counter()
    return {"seconds": t3 - t0, "numbering_seconds": t1 - t0, "operator_seconds": t2 - t1,
            "solve_seconds": t3 - t2, "dof": int(mngr.ndof), "checksum": float(sol.sum())}
