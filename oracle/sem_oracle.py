"""CPU oracle: a NumPy restatement of the reference's algorithm for the hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; it
is the checker, never the thing shipped or measured as the engine.  The product
package (spectralelementmethod_b200/) never imports it and has no CPU fallback.

Parity pinning: the reference's own tests hold no golden vectors for this path
(both test files are stale and only check analytic values
at rtol 1e-2).  The oracle is therefore pinned against outputs of the
*reference itself run in the development container* (oracle/live_reference.py
drives the unmodified /root/reference/sem through five compatibility shims);
oracle/make_golden.py froze those outputs into tests/golden/*.npz and
tests/test_oracle_golden.py checks this restatement against them.

Every function cites the reference lines it follows (paths relative to the
reference tree).  Where it is cheap the restatement is batched over elements,
but it keeps the reference's formulas (dense 4-index local stiffness from the
four einsums, dense local apply, LU solve with the equispaced matrix, ...).
The only data shared with the product is the GLL table
(spectralelementmethod_b200/gll_tables.py = bytes of sem/data/basis-data.hdf5).
"""
import os
import sys

import numpy as np
import scipy.linalg as sla
from scipy import sparse
from scipy.sparse import csgraph
from scipy.sparse.linalg import spsolve

_HERE = os.path.dirname(os.path.abspath(__file__))


def _tables():
    sys.path.insert(0, os.path.join(_HERE, ".."))
    try:
        from spectralelementmethod_b200 import gll_tables
    finally:
        sys.path.pop(0)
    return gll_tables


# --------------------------------------------------------------------------
# 1-D tables
# --------------------------------------------------------------------------
def gll(order):
    """nodes, barycentric weights, quadrature weights on [-1, 1], mirrored from
    the stored non-negative half (sem/basis_functions.py:372-388)."""
    half = np.array(_tables().half_table(order))
    n = order + 1
    nodes, bary, quad = np.zeros(n), np.zeros(n), np.zeros(n)
    m = n // 2
    nodes[m:], bary[m:], quad[m:] = half[0], half[1], half[2]
    if n % 2 == 1:
        nodes[:m] = -half[0, -1:0:-1]
        bary[:m] = half[1, -1:0:-1]
        quad[:m] = half[2, -1:0:-1]
    else:
        nodes[:m] = -half[0, -1::-1]
        bary[:m] = -half[1, -1::-1]
        quad[:m] = half[2, -1::-1]
    return nodes, bary, quad


def diff_matrix(nodes, bary):
    """D[i,j] = (w_j/w_i)/(x_i-x_j), D[i,i] = -sum_j D[i,j]
    (sem/basis_functions.py:213-217)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        D = bary[None, :] / bary[:, None]
        D /= nodes[:, None] - nodes[None, :]
    np.fill_diagonal(D, 0.0)
    np.fill_diagonal(D, -D.sum(axis=1))
    return D


def lagrange_eval(nodes, bary, x):
    """B[i,j] = l_j(x_i), barycentric form (sem/basis_functions.py:226-255)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        kern = bary / (x[..., None] - nodes)
        s = kern.sum(axis=-1)
        s.shape += (1,)
        out = kern / s
    out[np.isnan(out)] = 1.0
    return out


def interp_eq_matrix(nodes, bary):
    """GLL coefficients -> values at equispaced points
    (sem/basis_functions.py:221-223)."""
    return lagrange_eval(nodes, bary, np.linspace(-1, 1, nodes.size))


class Basis(object):
    """The tables one order needs."""

    def __init__(self, order):
        self.order = order
        self.N = order + 1
        self.nodes, self.bary, self.w = gll(order)
        self.D = diff_matrix(self.nodes, self.bary)
        self.E = interp_eq_matrix(self.nodes, self.bary)
        self.E_lu = sla.lu_factor(self.E)


def hier_order(N):
    """Hierarchical local node order of an N x N quadrilateral: 4 vertices,
    edges xi0=-1, xi0=+1, xi1=-1, xi1=+1 (open), interior
    (sem/geometry.py:151-212)."""
    lin = np.arange(N * N).reshape(N, N)
    parts = [[lin[0, 0]], [lin[0, -1]], [lin[-1, 0]], [lin[-1, -1]],
             lin[0, 1:-1], lin[-1, 1:-1], lin[1:-1, 0], lin[1:-1, -1],
             lin[1:-1, 1:-1].ravel()]
    return np.concatenate([np.asarray(p).ravel() for p in parts]).astype(np.uint32)


# --------------------------------------------------------------------------
# synthetic meshes (tests/test_discrete.py:22-38 pattern)
# --------------------------------------------------------------------------
def mesh_nodes(kind, nx, ny, p):
    NX, NY = nx * p + 1, ny * p + 1
    X, Y = np.meshgrid(np.linspace(-1, 1, NX), np.linspace(-1, 1, NY), indexing="ij")
    if kind == "C":
        s = 0.08 * np.sin(np.pi * X) * np.sin(np.pi * Y)
        X = X + s
        Y = Y + s
    return np.vstack([X.ravel(), Y.ravel()])


def mesh_l2g(nx, ny, p):
    """uint32[E, N, N]; cell (ex, ey) at index ex*ny+ey; node id = i*NY + j."""
    NX, NY = nx * p + 1, ny * p + 1
    gid = np.arange(NX * NY).reshape(NX, NY)
    return np.stack([gid[ex * p:ex * p + p + 1, ey * p:ey * p + p + 1]
                     for ex in range(nx) for ey in range(ny)]).astype(np.uint32)


def mesh_boundary_faces(nx, ny):
    """{'ebc': [(cell, face)...], 'nbc': [...]}: ebc = left (face 0) + bottom
    (face 2), nbc = right (1) + top (3); per cell in the order 0, 2, 1, 3."""
    out = {"ebc": [], "nbc": []}
    for ex in range(nx):
        for ey in range(ny):
            c = ex * ny + ey
            if ex == 0:
                out["ebc"].append((c, 0))
            if ey == 0:
                out["ebc"].append((c, 2))
            if ex == nx - 1:
                out["nbc"].append((c, 1))
            if ey == ny - 1:
                out["nbc"].append((c, 3))
    return out


def face_nodes(cell_map, face):
    """Counter-clockwise node ids of a face of a 2-D cell map
    (sem/mapping.py:19-76 specialised to ndim=2)."""
    if face == 0:
        return cell_map[..., 0, ::-1]
    if face == 1:
        return cell_map[..., -1, :]
    if face == 2:
        return cell_map[..., :, 0]
    return cell_map[..., ::-1, -1]


def permute_nodes(nodes, l2g, perm):
    """Mesh._permute_nodes (sem/discrete.py:1115-1127): new node k = old
    perm[k]; returns the new (nodes, l2g)."""
    nodes = nodes.copy()
    nodes[:, :perm.size] = nodes[:, perm]
    inv = np.zeros_like(perm)
    inv[perm] = np.arange(perm.size)
    return nodes, inv[l2g].astype(np.uint32)


def static_condensation(nodes, l2g):
    """Exterior-first ordering (sem/discrete.py:314-359).  Returns
    (nodes, l2g, n_exterior)."""
    E, N = l2g.shape[0], l2g.shape[1]
    h = hier_order(N)
    n_ext = N * N - (N - 2) ** 2
    flat = l2g.reshape(E, -1)
    ext = np.unique(flat[:, h[:n_ext]].astype(int))
    itr = np.sort(flat[:, h[n_ext:]].astype(int).ravel())
    perm = np.concatenate((ext, itr))
    assert perm.size == nodes.shape[1]
    nodes, l2g = permute_nodes(nodes, l2g, perm)
    return nodes, l2g, ext.size


def _graph(idsets, size):
    rows = np.concatenate([np.repeat(ids, ids.shape[1], axis=1).ravel() for ids in idsets])
    cols = np.concatenate([np.tile(ids, (1, ids.shape[1])).ravel() for ids in idsets])
    g = sparse.coo_matrix((np.ones(rows.size, dtype=bool), (rows, cols)), (size, size))
    return g.tocsr()


def rcm_all(nodes, l2g):
    """DOFManager._reorder_nodes_rcm (sem/discrete.py:142-178)."""
    E = l2g.shape[0]
    g = _graph([l2g.reshape(E, -1)], nodes.shape[1])
    perm = csgraph.reverse_cuthill_mckee(g, True)
    return permute_nodes(nodes, l2g, perm)


def rcm_exterior(nodes, l2g, n_ext):
    """DOFManagerSC._reorder_nodes_rcm (sem/discrete.py:361-402): RCM over the
    exterior nodes, interior identity."""
    E, N = l2g.shape[0], l2g.shape[1]
    h = hier_order(N)
    k = N * N - (N - 2) ** 2
    g = _graph([l2g.reshape(E, -1)[:, h[:k]]], n_ext)
    n = nodes.shape[1]
    perm = np.empty(n, np.uint32)
    perm[:n_ext] = csgraph.reverse_cuthill_mckee(g, True)
    perm[n_ext:] = np.arange(n_ext, n)
    return permute_nodes(nodes, l2g, perm)


def build_case(kind, nx, ny, p, sc, rcm):
    """nodes, l2g after the manager's renumbering (DOFManager /
    DOFManagerSC with rcm_order)."""
    nodes, l2g = mesh_nodes(kind, nx, ny, p), mesh_l2g(nx, ny, p)
    if sc:
        nodes, l2g, n_ext = static_condensation(nodes, l2g)
        if rcm:
            nodes, l2g = rcm_exterior(nodes, l2g, n_ext)
    elif rcm:
        nodes, l2g = rcm_all(nodes, l2g)
    return nodes, l2g


# --------------------------------------------------------------------------
# geometry and the local operator
# --------------------------------------------------------------------------
def geometry(basis, nodes, l2g):
    """x_phys [E,2,N,N], J, invJ [E,2,2,N,N], detJ, JxW [E,N,N].

    x_phys: LU solves with E along axis 0 then axis 1
    (sem/mapping.py:98-103, sem/basis_functions.py:599-624);
    J[i,a] = d x_i/d xi_a (sem/mapping.py:113, sem/basis_functions.py:626-650);
    det/inv closed form with inv *= 1/det (sem/linalg.py:105-115);
    JxW = (detJ*w_m)*w_n (sem/quadratures.py:268-275)."""
    E, N = l2g.shape[0], basis.N
    X = nodes[:, l2g]                                   # [2, E, N, N]
    X = np.moveaxis(X, 1, 0).copy()                     # [E, 2, N, N]
    # axis 0 solve
    a = np.moveaxis(X, 2, 0).reshape(N, -1)
    a = sla.lu_solve(basis.E_lu, a).reshape(N, E, 2, N)
    a = np.moveaxis(a, 0, 2)                            # [E, 2, N, N]
    # axis 1 solve
    b = np.moveaxis(a, 3, 0).reshape(N, -1)
    b = sla.lu_solve(basis.E_lu, b).reshape(N, E, 2, N)
    xph = np.ascontiguousarray(np.moveaxis(b, 0, 3))    # [E, 2, N, N]
    d0 = np.einsum("mr,eirn->eimn", basis.D, xph)       # d/dxi0
    d1 = np.einsum("ns,eims->eimn", basis.D, xph)       # d/dxi1
    J = np.stack([d0, d1], axis=2)                      # [E, i, a, N, N]
    det = J[:, 0, 0] * J[:, 1, 1] - J[:, 0, 1] * J[:, 1, 0]
    assert np.all(det > 0)                              # sem/mapping.py:117
    inv = np.empty_like(J)
    inv[:, 0, 0] = J[:, 1, 1]
    inv[:, 0, 1] = -J[:, 0, 1]
    inv[:, 1, 0] = -J[:, 1, 0]
    inv[:, 1, 1] = J[:, 0, 0]
    inv *= (1 / det)[:, None, None]
    JxW = det.copy()
    JxW *= basis.w[:, None]
    JxW *= basis.w[None, :]
    return dict(x_phys=xph, J=J, invJ=inv, detJ=det, JxW=JxW)


def local_stiffness(basis, invJ, JxW):
    """Dense local stiffness L[E,p,q,r,s] by the reference's four einsums
    (examples/poisson.py:166-193), batched over elements."""
    D, N = basis.D, basis.N
    g0 = np.einsum("mp,eimn->eimnp", D, invJ[:, 0])     # gradh_xi0
    g1 = np.einsum("nq,eimn->eimnq", D, invJ[:, 1])     # gradh_xi1
    E = invJ.shape[0]
    L = np.zeros((E, N, N, N, N))
    p, q, r = np.ogrid[0:N, 0:N, 0:N]
    L[:, p, q, r, q] += np.einsum("emn,eimnp,eimnr->epnr", JxW, g0, g0)
    L += np.einsum("emn,eimnp,eimns->epnms", JxW, g0, g1)
    L += np.einsum("emn,eimnq,eimnr->emqrn", JxW, g1, g0)
    L[:, p, q, p, r] += np.einsum("emn,eimnq,eimns->emqs", JxW, g1, g1)
    return L


def apply_dense_local(L, l2g, u):
    """y[L2G] += einsum('pqrs,rs', L_e, u[L2G]) element by element
    (examples/squirmer-axisymmetric.py:268-270,284-295 + the scatter of
    sem/discrete.py:499)."""
    y = np.zeros_like(u)
    for e in range(l2g.shape[0]):
        idx = l2g[e]
        y[idx] += np.einsum("pqrs,rs", L[e], u[idx])
    return y


def apply_dense_batched(L, l2g, u):
    """Same contraction, batched (used where the Python loop is too slow)."""
    yl = np.einsum("epqrs,ers->epq", L, u[l2g])
    y = np.zeros_like(u)
    np.add.at(y, l2g.ravel(), yl.ravel())
    return y


def assemble_csr(L, l2g, n):
    """COO of all local blocks -> CSR with duplicates summed
    (sem/discrete.py:491-499,507 applied to the full local matrices)."""
    E, N = l2g.shape[0], l2g.shape[1]
    ids = l2g.reshape(E, -1).astype(np.int64)
    nn = N * N
    rows = np.repeat(ids, nn, axis=1).ravel()
    cols = np.tile(ids, (1, nn)).ravel()
    return sparse.coo_matrix((L.reshape(-1), (rows, cols)), shape=(n, n)).tocsr()


def assemble_vector(loc, l2g, n):
    out = np.zeros(n)
    np.add.at(out, l2g.ravel(), loc.ravel())
    return out


def local_diagonal(L):
    return np.einsum("epqpq->epq", L)


def dirichlet_data(nodes_unused, l2g, x_phys, faces):
    """on_ebc mask and u = 0.2((x+1)+(y+1)) at the GLL points of the 'ebc'
    faces (examples/poisson.py:125-143 via boundary_elements,
    sem/discrete.py:211-219,702-705)."""
    n = int(l2g.max()) + 1
    on = np.zeros(n, dtype=bool)
    vals = np.zeros(n)
    for c, f in sorted(faces["ebc"], key=lambda cf: cf[0]):
        loc = face_nodes(l2g[c], f)
        x = face_nodes(x_phys[c, 0], f)
        y = face_nodes(x_phys[c, 1], f)
        vals[loc] = 0.2 * ((x + 1) + (y + 1))
        on[loc] = True
    return on, vals


def solve_direct(A, b, on_ebc, vals):
    """Essential-BC elimination + sparse direct solve on the full assembled
    matrix (the operations of sem/discrete.py:505-511)."""
    free = ~on_ebc
    sol = vals.copy()
    Af = A[free]
    rhs = b[free] - Af[:, on_ebc] @ vals[on_ebc]
    sol[free] = spsolve(Af[:, free].tocsc(), rhs)
    return sol


def solve_schur(L, JxW, l2g, n_ext, on_ebc, vals):
    """The reference's actual solver: hierarchical reorder, local Schur
    complements, COO assembly over exterior DOFs, eliminate essential BCs,
    spsolve, interior back-substitution (sem/discrete.py:404-528).  Requires
    the exterior-first numbering (DOFManagerSC)."""
    E, N = l2g.shape[0], l2g.shape[1]
    nn = N * N
    h = hier_order(N).astype(np.int64)
    ne = nn - (N - 2) ** 2
    Lm = L.reshape(E, nn, nn)[:, h][:, :, h]            # reorder_local_system_hier
    rh = JxW.reshape(E, nn)[:, h]
    gid = l2g.reshape(E, nn).astype(np.int64)[:, h]
    Aee, Aei = Lm[:, :ne, :ne], Lm[:, :ne, ne:]
    Aie, Aii = Lm[:, ne:, :ne], Lm[:, ne:, ne:]
    X = np.swapaxes(np.linalg.solve(np.swapaxes(Aii, 1, 2), np.swapaxes(Aei, 1, 2)), 1, 2)
    S = Aee - X @ Aie
    g = rh[:, :ne] - np.einsum("eij,ej->ei", X, rh[:, ne:])
    ids = gid[:, :ne]
    rows = np.repeat(ids, ne, axis=1).ravel()
    cols = np.tile(ids, (1, ne)).ravel()
    Sg = sparse.coo_matrix((S.reshape(-1), (rows, cols)), shape=(n_ext, n_ext)).tocsr()
    grhs = np.zeros(n_ext)
    np.add.at(grhs, ids.ravel(), g.ravel())
    sol = vals.copy()
    ebc = on_ebc[:n_ext]
    free = ~ebc
    ext = sol[:n_ext]
    A1 = Sg[free]
    r1 = grhs[free] - A1[:, ebc] @ ext[ebc]
    ext[free] = spsolve(A1[:, free].tocsc(), r1)
    inner = np.linalg.solve(Aii, (rh[:, ne:] - np.einsum("eij,ej->ei", Aie, sol[ids]))[..., None])
    sol[gid[:, ne:]] = inner[..., 0]
    return sol


def condensed_system(p, invJ, JxW, l2g):
    """Local Schur complements, condensed local loads and the assembled condensed
    system over the element-exterior DOFs: compute_local_sc_system /
    assemble_global_sc_system (sem/discrete.py:438-500) with the Poisson recipe
    of examples/poisson.py:166-200.  Requires the exterior-first numbering."""
    N, nn, ne = p + 1, (p + 1) ** 2, 4 * p
    L = local_stiffness(Basis(p), invJ, JxW)
    E = L.shape[0]
    h = hier_order(N).astype(np.int64)
    Lm = L.reshape(E, nn, nn)[:, h][:, :, h]               # reorder_local_system_hier
    rh = JxW.reshape(E, nn)[:, h]
    ids = l2g.reshape(E, nn).astype(np.int64)[:, h][:, :ne]
    Aee, Aei, Aie, Aii = Lm[:, :ne, :ne], Lm[:, :ne, ne:], Lm[:, ne:, :ne], Lm[:, ne:, ne:]
    X = np.swapaxes(np.linalg.solve(np.swapaxes(Aii, 1, 2), np.swapaxes(Aei, 1, 2)), 1, 2)
    S = Aee - X @ Aie
    gl = rh[:, :ne] - np.einsum("eij,ej->ei", X, rh[:, ne:])
    n_ext = int(ids.max()) + 1
    rows = np.repeat(ids, ne, axis=1).ravel()
    cols = np.tile(ids, (1, ne)).ravel()
    Sg = sparse.coo_matrix((S.reshape(-1), (rows, cols)), shape=(n_ext, n_ext)).tocsr()
    grhs = np.zeros(n_ext)
    np.add.at(grhs, ids.ravel(), gl.ravel())
    return dict(S=S, g_loc=gl, ids=ids, Sg=Sg, grhs=grhs, n_ext=n_ext, Aii=Aii, Aie=Aie,
                f_int=rh[:, ne:], int_ids=l2g.reshape(E, nn).astype(np.int64)[:, h][:, ne:])


def pcg_jacobi(A, b, x0, rtol, maxiter):
    """Plain Jacobi-PCG on an assembled SPD matrix (protocol of SURVEY.md
    8(d): stop on the recursive residual ||r|| <= rtol ||b||).  Checker for
    the device solver's iteration counts; not part of the reference."""
    dinv = 1.0 / A.diagonal()
    x = x0.copy()
    r = b - A @ x
    z = dinv * r
    p = z.copy()
    rz = r @ z
    bb = b @ b
    it = 0
    while it < maxiter and r @ r > rtol * rtol * bb:
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        z = dinv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
    return x, it


def run_case(kind, nx, ny, p, sc, rcm, solve=True, python_loop_apply=False):
    """The oracle's counterpart of oracle/live_reference.run_case."""
    basis = Basis(p)
    nodes, l2g = build_case(kind, nx, ny, p, sc, rcm)
    n = nodes.shape[1]
    geo = geometry(basis, nodes, l2g)
    L = local_stiffness(basis, geo["invJ"], geo["JxW"])
    x, y = nodes
    u = np.sin(3 * x) * np.cos(2 * y)
    Au = apply_dense_local(L, l2g, u) if python_loop_apply else apply_dense_batched(L, l2g, u)
    out = dict(l2g=l2g, nodes=nodes, invJ=geo["invJ"], JxW=geo["JxW"], x_phys=geo["x_phys"],
               u=u, Au=Au, b=assemble_vector(geo["JxW"], l2g, n),
               diag=assemble_vector(local_diagonal(L), l2g, n), L=L)
    on, vals = dirichlet_data(nodes, l2g, geo["x_phys"], mesh_boundary_faces(nx, ny))
    out["on_ebc"], out["ebc_vals"] = on, vals
    if solve:
        if sc:
            n_ext = int(np.unique(l2g.reshape(l2g.shape[0], -1)[
                :, hier_order(basis.N)[:basis.N ** 2 - (basis.N - 2) ** 2]]).size)
            out["solution"] = solve_schur(L, geo["JxW"], l2g, n_ext, on, vals)
        else:
            A = assemble_csr(L, l2g, n)
            out["solution"] = solve_direct(A, out["b"], on, vals)
    return out


# --------------------------------------------------------------------------
# C/OpenMP restatement of the dense local apply (all host threads)
# --------------------------------------------------------------------------
_C_LIB = None


def c_lib():
    """ctypes handle of oracle/_build/libsem_oracle_c.so (built by
    oracle/build_c.sh / __graft_entry__.build()), or None if not built."""
    global _C_LIB
    if _C_LIB is None:
        import ctypes
        path = os.path.join(_HERE, "_build", "libsem_oracle_c.so")
        if not os.path.exists(path):
            return None
        os.environ.setdefault("OMP_PROC_BIND", "true")   # unpinned threads scale negatively
        lib = ctypes.CDLL(path)
        lib.sem_oracle_c_threads.restype = ctypes.c_int
        lib.sem_oracle_c_set_threads.restype = None
        lib.sem_oracle_c_set_threads.argtypes = [ctypes.c_int]
        lib.sem_oracle_c_apply_dense.restype = None
        lib.sem_oracle_c_apply_dense.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int64] + \
            [ctypes.c_void_p] * 4
        _C_LIB = lib
    return _C_LIB


def apply_dense_c(L, l2g, u):
    """y[L2G] += L_e . u[L2G] over all elements with OpenMP (same contraction
    as apply_dense_local; examples/squirmer-axisymmetric.py:268-295)."""
    lib = c_lib()
    if lib is None:
        raise RuntimeError("oracle C library not built (run oracle/build_c.sh)")
    E, N = l2g.shape[0], l2g.shape[1]
    NN = N * N
    Lc = np.ascontiguousarray(L, dtype=np.float64).reshape(E, NN, NN)
    idx = np.ascontiguousarray(l2g, dtype=np.uint32).reshape(E, NN)
    uc = np.ascontiguousarray(u, dtype=np.float64)
    y = np.empty_like(uc)
    lib.sem_oracle_c_apply_dense(E, NN, uc.size, Lc.ctypes.data, idx.ctypes.data, uc.ctypes.data,
                                 y.ctypes.data)
    return y


def c_threads():
    lib = c_lib()
    return int(lib.sem_oracle_c_threads()) if lib is not None else 1


def c_set_threads(n=None):
    """Use ``n`` OpenMP threads (default: every host core) whatever OMP_NUM_THREADS says
    (torchrun exports OMP_NUM_THREADS=1 to its workers).  Returns the count in effect."""
    lib = c_lib()
    if lib is None:
        return 1
    lib.sem_oracle_c_set_threads(int(n or os.cpu_count() or 1))
    return c_threads()


# --------------------------------------------------------------------------
# Axisymmetric Stokes / Navier-Stokes in stream function - vorticity form
# (examples/squirmer-axisymmetric.py).  Two DOFs per node,
# DOF id = 2*node + comp (sem/discrete.py:561-576): comp 0 = stream function,
# comp 1 = vorticity.  Pinned by tests/golden/stokes_*.npz, which hold what the
# example's OWN class computes when run live (oracle/live_squirmer.py,
# oracle/make_golden_stokes.py).
# --------------------------------------------------------------------------
def annulus_nodes(nr, nt, p, r_out):
    """Synthetic stand-in for examples/meshes/donut.geo (no .msh ships): meridional
    half plane (rho, z), r = r_out**s in [1, r_out], theta from pi down to 0 so that
    detJ > 0; equispaced in the parametric coordinates of every element."""
    NR, NT = nr * p + 1, nt * p + 1
    r = r_out ** np.linspace(0.0, 1.0, NR)
    th = np.linspace(np.pi, 0.0, NT)
    sin = np.sin(th)
    sin[0] = 0.0
    sin[-1] = 0.0
    return np.vstack([np.outer(r, sin).ravel(), np.outer(r, np.cos(th)).ravel()])


def annulus_boundary_faces(nr, nt):
    """{name: [(cell, face)]} in the registration order of the mesh builder."""
    out = {"sphere": [], "shell": [], "symaxis": []}
    c = 0
    for ex in range(nr):
        for ey in range(nt):
            if ex == 0:
                out["sphere"].append((c, 0))
            if ex == nr - 1:
                out["shell"].append((c, 1))
            if ey == 0:
                out["symaxis"].append((c, 2))
            if ey == nt - 1:
                out["symaxis"].append((c, 3))
            c += 1
    return out


def stokes_local_operators(basis, x_phys, invJ, JxW):
    """Dense local operators of examples/squirmer-axisymmetric.py:177-254, batched over
    elements: E2e, Lve [E,p,q,r,s], the mass diagonal Me [E,m,n] and the four diagonals
    of the advection operator Ae WITHOUT the Reynolds-number factor (:228-249)."""
    D, N = basis.D, basis.N
    rho = x_phys[:, 0]
    g0 = np.einsum("mp,eimn->eimnp", D, invJ[:, 0])     # gradh_xi0 (:188)
    g1 = np.einsum("nq,eimn->eimnq", D, invJ[:, 1])     # gradh_xi1 (:190)
    E = invJ.shape[0]
    rJ = rho * JxW                                      # rho_JxW (:194)
    E2 = np.zeros((E, N, N, N, N))
    p, q, r = np.ogrid[0:N, 0:N, 0:N]
    E2[:, p, q, r, q] += np.einsum("emn,eimnp,eimnr->epnr", rJ, g0, g0)      # :198-199
    E2 += np.einsum("emn,eimnp,eimns->epnms", rJ, g0, g1)                    # :200-201
    E2 += np.einsum("emn,eimnq,eimnr->emqrn", rJ, g1, g0)                    # :204-205
    E2[:, p, q, p, r] += np.einsum("emn,eimnq,eimns->emqs", rJ, g1, g1)      # :206-207
    Lv = E2.copy()                                                           # :210
    pp, qq = np.ogrid[0:N, 0:N]
    with np.errstate(divide="ignore", invalid="ignore"):
        jr = JxW / rho                                                       # :211 (inf on the axis)
    Lv[:, pp, qq, pp, qq] += jr
    E2[:, p, q, r, q] += 2 * np.einsum("emn,emnr->emnr", JxW, g0[:, 0])     # :221
    E2[:, p, q, p, r] += 2 * np.einsum("emn,emns->emns", JxW, g1[:, 0])     # :222
    adv = dict(
        a1=(np.einsum("emn,emnr,emnu->emnru", JxW, g0[:, 0], g1[:, 1]) -
            np.einsum("emn,emnr,emnu->emnru", JxW, g0[:, 1], g1[:, 0])),    # :227-231, axes [0,1,2,1,0,3]
        a2=(np.einsum("emn,emns,emnt->emnst", JxW, g1[:, 0], g0[:, 1]) -
            np.einsum("emn,emns,emnt->emnst", JxW, g1[:, 1], g0[:, 0])),    # :233-238, axes [0,1,0,2,3,1]
        a3=np.einsum("emn,emnr->emnr", jr, g0[:, 1]),                       # :240-242, axes [0,1,2,1,0,1]
        a4=np.einsum("emn,emns->emns", jr, g1[:, 1]))                       # :244-247, axes [0,1,0,2,0,1]
    Me = rJ * rho                                                           # :252
    return dict(E2e=E2, Lve=Lv, Me=Me, adv=adv)


def stokes_local_system(ops, n_rey, sfn, vort):
    """(jac_l [E,2nn,2nn], -res_l [E,2nn]) of compute_local_system
    (examples/squirmer-axisymmetric.py:259-297); sfn, vort [E,N,N] local values."""
    E2, Lv, Me, adv = ops["E2e"], ops["Lve"], ops["Me"], ops["adv"]
    E, N = Me.shape[0], Me.shape[1]
    nn = N * N
    old_err = np.seterr(invalid="ignore")      # 0 * inf on the axis of symmetry, as in the reference
    a1, a2, a3, a4 = (n_rey * adv[k] for k in ("a1", "a2", "a3", "a4"))
    m, n, k = np.ogrid[0:N, 0:N, 0:N]
    # Ae.dot_dense(vort, [4, 5]) -> [m,n,r,s] (:276).  Entries are PLACED on their
    # Kronecker diagonals like KroneckerArray.to_array (sem/sp_array.py:104-113), never
    # multiplied by a zero: JxW/rho is infinite on the axis of symmetry.
    Aw = np.zeros((E, N, N, N, N))
    Aw[:, m, n, k, n] += np.einsum("emnru,emu->emnr", a1, vort)
    Aw[:, m, n, m, k] += np.einsum("emnst,etn->emns", a2, vort)
    Aw[:, m, n, k, n] += a3 * vort[:, :, :, None]
    Aw[:, m, n, m, k] += a4 * vort[:, :, :, None]
    # Ae.dot_dense(sfn, [2, 3]) -> [m,n,t,u] (:281)
    As = np.zeros((E, N, N, N, N))
    As[:, m, n, m, k] += np.einsum("emnru,ern->emnu", a1, sfn)
    As[:, m, n, k, n] += np.einsum("emnst,ems->emnt", a2, sfn)
    mm, nn2 = np.ogrid[0:N, 0:N]
    As[:, mm, nn2, mm, nn2] += (np.einsum("emnr,ern->emn", a3, sfn) +
                                np.einsum("emns,ems->emn", a4, sfn))
    jac = np.zeros((E, 2 * nn, 2 * nn))
    res = np.zeros((E, 2 * nn))
    Md = np.zeros((E, N, N, N, N))
    pp, qq = np.ogrid[0:N, 0:N]
    Md[:, pp, qq, pp, qq] = Me
    jac[:, 0::2, 0::2] = Aw.reshape(E, nn, nn)
    jac[:, 0::2, 1::2] = (As + Lv).reshape(E, nn, nn)
    with np.errstate(invalid="ignore"):
        res[:, 0::2] = (np.einsum("emnrs,ers->emn", Aw, sfn) +
                        np.einsum("epqrs,ers->epq", Lv, vort)).reshape(E, nn)
    jac[:, 1::2, 0::2] = E2.reshape(E, nn, nn)
    jac[:, 1::2, 1::2] = -Md.reshape(E, nn, nn)
    res[:, 1::2] = (np.einsum("epqrs,ers->epq", E2, sfn) - Me * vort).reshape(E, nn)
    np.seterr(**old_err)
    return jac, -res


def hier_dof_order(N, dpn):
    """Local DOF order with the element-exterior DOFs first (sem/discrete.py:611-625):
    node order of hier_order, DOFs of a node adjacent."""
    h = hier_order(N).astype(np.int64)
    return (h[:, None] * dpn + np.arange(dpn)[None, :]).ravel()


def stokes_newton_step(jac, rhs, l2g, n_ext_nodes, dof_mask, cint):
    """One pass of the example's solve loop (:420-431): hierarchical reorder (:300-306),
    local Schur complements by a transposed dense solve (:318-321), COO assembly over the
    exterior DOFs with the natural-BC contour integrals as the initial RHS (:336-358),
    elimination of essential rows / columns against a ZERO increment (:362-370), spsolve,
    interior back-substitution (:372-386).  Returns the increment dsoln [2 n_nodes]."""
    E, N = l2g.shape[0], l2g.shape[1]
    nn = N * N
    hd = hier_dof_order(N, 2)
    ne = 2 * (nn - (N - 2) ** 2)
    n_ext = 2 * n_ext_nodes
    A = jac[:, hd][:, :, hd]
    b = rhs[:, hd]
    gid = (2 * l2g.reshape(E, nn).astype(np.int64)[:, :, None] + np.arange(2)).reshape(E, 2 * nn)[:, hd]
    Aee, Aei, Aie, Aii = A[:, :ne, :ne], A[:, :ne, ne:], A[:, ne:, :ne], A[:, ne:, ne:]
    with np.errstate(invalid="ignore"):
        X = np.swapaxes(np.linalg.solve(np.swapaxes(Aii, 1, 2), np.swapaxes(Aei, 1, 2)), 1, 2)
        S = Aee - X @ Aie
        g = b[:, :ne] - np.einsum("eij,ej->ei", X, b[:, ne:])
    ids = gid[:, :ne]
    rows = np.repeat(ids, ne, axis=1).ravel()
    cols = np.tile(ids, (1, ne)).ravel()
    Sg = sparse.coo_matrix((S.reshape(-1), (rows, cols)), shape=(n_ext, n_ext)).tocsr()
    grhs = cint.copy()
    np.add.at(grhs, ids.ravel(), g.ravel())
    n_tot = 2 * (int(l2g.max()) + 1)
    d = np.zeros(n_tot)
    unk = dof_mask
    A1 = Sg[unk]
    r1 = grhs[unk] - A1[:, ~unk] @ d[:n_ext][~unk]
    d[:n_ext][unk] = spsolve(A1[:, unk].tocsc(), r1)
    inner = np.linalg.solve(Aii, (b[:, ne:] - np.einsum("eij,ej->ei", Aie, d[ids]))[..., None])
    d[gid[:, ne:]] = inner[..., 0]
    return d


def stokes_global_jacobian(jac, l2g):
    """Assembled (uncondensed, no BCs) Jacobian over all 2*n_nodes DOFs: what the
    matrix-free device apply is compared with."""
    E, N = l2g.shape[0], l2g.shape[1]
    nn = N * N
    gid = (2 * l2g.reshape(E, nn).astype(np.int64)[:, :, None] + np.arange(2)).reshape(E, 2 * nn)
    rows = np.repeat(gid, 2 * nn, axis=1).ravel()
    cols = np.tile(gid, (1, 2 * nn)).ravel()
    J = np.where(np.isfinite(jac), jac, 0.0)   # JxW/rho on the axis: rows / columns eliminated by the BCs
    n = 2 * (int(l2g.max()) + 1)
    return sparse.coo_matrix((J.reshape(-1), (rows, cols)), shape=(n, n)).tocsr()
