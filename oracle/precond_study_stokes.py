#!/usr/bin/env python
"""Study (CPU, oracle matrices): a block-triangular preconditioner for the axisymmetric
Stokes stream function - vorticity system built from ONE weighted Poisson operator.

    python oracle/precond_study_stokes.py

System after eliminating the essential DOFs (rows wte at psi-free nodes I, rows wdef at
omega-free nodes I + Gamma, Gamma = sphere):   L om = a,   E psi - M om = b.
Preconditioner: om_G = -b_G / M_G;  om_I = Khat^-1 (a - L_IG om_G);  psi = Khat^-1 (b_I + M om_I)
with Khat = rho-weighted stiffness, Dirichlet on the whole boundary (exact inverse here).
TEST / DESIGN INFRASTRUCTURE, not product code."""
import sys
import os
import time

import numpy as np
from scipy import sparse
from scipy.sparse.linalg import splu, gmres, LinearOperator

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sem_oracle as so  # noqa: E402


def build(nr, nt, p, r_out, variant):
    nodes = so.annulus_nodes(nr, nt, p, r_out)
    l2g = so.mesh_l2g(nr, nt, p)
    basis = so.Basis(p)
    geo = so.geometry(basis, nodes, l2g)
    ops = so.stokes_local_operators(basis, geo["x_phys"], geo["invJ"], geo["JxW"])
    E_, N = l2g.shape[0], p + 1
    z = np.zeros((E_, N, N))
    jac, _ = so.stokes_local_system(ops, 0.0, z, z)
    J = so.stokes_global_jacobian(jac, l2g).tocsr()
    n = nodes.shape[1]
    NR, NT = nr * p + 1, nt * p + 1
    gid = np.arange(n).reshape(NR, NT)
    sphere, shell = gid[0, :], gid[-1, :]
    axis = np.concatenate([gid[:, 0], gid[:, -1]])
    ess_s = np.zeros(n, bool)
    ess_w = np.zeros(n, bool)
    ess_s[sphere] = ess_s[shell] = ess_s[axis] = True
    ess_w[shell] = ess_w[axis] = True
    ess = np.empty(2 * n, bool)
    ess[0::2], ess[1::2] = ess_s, ess_w
    free = ~ess
    A = J[free][:, free].tocsc()
    # scalar weighted stiffness (rho JxW grad.grad) assembled from Lve minus the 1/rho term
    nn = N * N
    Kloc = ops["E2e"].copy()
    # remove the first-derivative part: rebuild K from Lve - diag(JxW/rho)
    Kloc = ops["Lve"].copy()
    pp, qq = np.ogrid[0:N, 0:N]
    with np.errstate(divide="ignore", invalid="ignore"):
        jr = geo["JxW"] / geo["x_phys"][:, 0]
    if variant == "K":
        Kloc[:, pp, qq, pp, qq] -= jr
    Kloc = np.where(np.isfinite(Kloc), Kloc, 0.0)
    ids = l2g.reshape(E_, nn).astype(np.int64)
    rows = np.repeat(ids, nn, axis=1).ravel()
    cols = np.tile(ids, (1, nn)).ravel()
    K = sparse.coo_matrix((Kloc.reshape(-1), (rows, cols)), shape=(n, n)).tocsr()
    return J, A, K, ess_s, ess_w, free, n


def run(nr, nt, p, r_out=100.0, variant="K"):
    J, A, K, ess_s, ess_w, free, n = build(nr, nt, p, r_out, variant)
    I = ~ess_s                       # psi-free = interior nodes
    G = (~ess_w) & ess_s             # sphere nodes: omega free, psi essential
    Kii = splu(K[I][:, I].tocsc())
    # index maps into the reduced vector (free DOFs in global DOF order)
    pos = -np.ones(2 * n, dtype=np.int64)
    pos[free] = np.arange(free.sum())
    nodes_I, nodes_G = np.flatnonzero(I), np.flatnonzero(G)
    p_psi, p_omI, p_omG = pos[2 * nodes_I], pos[2 * nodes_I + 1], pos[2 * nodes_G + 1]
    Jc = J.tocsr()
    M_G = -Jc[2 * nodes_G + 1][:, 2 * nodes_G + 1].diagonal()
    M_I = -Jc[2 * nodes_I + 1][:, 2 * nodes_I + 1].diagonal()
    L_IG = Jc[2 * nodes_I][:, 2 * nodes_G + 1]

    def prec(r):
        z = np.zeros_like(r)
        a, bI, bG = r[p_psi], r[p_omI], r[p_omG]      # rows: wte at I (dof 2k), wdef at I, wdef at G
        omG = -bG / M_G
        omI = Kii.solve(a - L_IG @ omG)
        psi = Kii.solve(bI + M_I * omI)
        z[p_psi], z[p_omI], z[p_omG] = psi, omI, omG
        return z

    rng = np.random.default_rng(0)
    b = rng.standard_normal(A.shape[0])
    res = []
    t0 = time.perf_counter()
    x, info = gmres(A, b, M=LinearOperator(A.shape, prec), rtol=1e-8, restart=400, maxiter=400,
                    callback=lambda rk: res.append(rk), callback_type="pr_norm")
    true = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    print("%2dx%2d p=%d r_out=%g variant=%s: %6d DOF, |Gamma|=%4d, GMRES its %4d, info %d, true res %.1e  (%.1f s)"
          % (nr, nt, p, r_out, variant, A.shape[0], G.sum(), len(res), info, true,
             time.perf_counter() - t0), flush=True)


if __name__ == "__main__":
    for variant in ("K", "L"):
        for (nr, nt, p) in ((3, 4, 4), (6, 8, 4), (12, 16, 4), (6, 8, 8), (24, 32, 4)):
            run(nr, nt, p, 100.0, variant)


def fgmres(A, b, prec, rtol, maxiter):
    """Flexible GMRES without restart (right preconditioning, modified Gram-Schmidt)."""
    n = b.size
    V = [b / np.linalg.norm(b)]
    Z = []
    H = np.zeros((maxiter + 1, maxiter))
    beta = np.linalg.norm(b)
    for j in range(maxiter):
        Z.append(prec(V[j]))
        w = A @ Z[j]
        for i in range(j + 1):
            H[i, j] = V[i] @ w
            w -= H[i, j] * V[i]
        H[j + 1, j] = np.linalg.norm(w)
        V.append(w / H[j + 1, j])
        e1 = np.zeros(j + 2)
        e1[0] = beta
        y, *_ = np.linalg.lstsq(H[:j + 2, :j + 1], e1, rcond=None)
        r = np.linalg.norm(H[:j + 2, :j + 1] @ y - e1) / beta
        if r < rtol:
            break
    x = sum(yi * zi for yi, zi in zip(y, Z))
    return x, j + 1


def run_inexact(nr, nt, p, inner_rtol, variant="K", r_out=100.0):
    from scipy.sparse.linalg import cg
    J, A, K, ess_s, ess_w, free, n = build(nr, nt, p, r_out, variant)
    I = ~ess_s
    G = (~ess_w) & ess_s
    Kii = K[I][:, I].tocsr()
    dinv = 1.0 / Kii.diagonal()
    pos = -np.ones(2 * n, dtype=np.int64)
    pos[free] = np.arange(free.sum())
    nodes_I, nodes_G = np.flatnonzero(I), np.flatnonzero(G)
    p_psi, p_omI, p_omG = pos[2 * nodes_I], pos[2 * nodes_I + 1], pos[2 * nodes_G + 1]
    Jc = J.tocsr()
    M_G = -Jc[2 * nodes_G + 1][:, 2 * nodes_G + 1].diagonal()
    M_I = -Jc[2 * nodes_I + 1][:, 2 * nodes_I + 1].diagonal()
    L_IG = Jc[2 * nodes_I][:, 2 * nodes_G + 1]
    inner = [0]

    def solve(rhs):
        def cb(_):
            inner[0] += 1
        x, _ = cg(Kii, rhs, rtol=inner_rtol, maxiter=20000, M=sparse.diags(dinv), callback=cb)
        return x

    def prec(r):
        z = np.zeros_like(r)
        a, bI, bG = r[p_psi], r[p_omI], r[p_omG]
        omG = -bG / M_G
        omI = solve(a - L_IG @ omG)
        psi = solve(bI + M_I * omI)
        z[p_psi], z[p_omI], z[p_omG] = psi, omI, omG
        return z

    rng = np.random.default_rng(0)
    b = rng.standard_normal(A.shape[0])
    x, its = fgmres(A, b, prec, 1e-8, 400)
    true = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    print("inexact %2dx%2d p=%d variant=%s inner rtol %.0e: FGMRES its %d, true res %.1e, inner CG its %d"
          % (nr, nt, p, variant, inner_rtol, its, true, inner[0]), flush=True)
