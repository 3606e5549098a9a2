#!/usr/bin/env python
"""Design study behind the two-level preconditioner (TEST INFRASTRUCTURE / evidence, not
product code): iteration counts of PCG on the oracle's assembled condensed system with
(a) Jacobi on the uncondensed operator, (b) Jacobi on the condensed operator, (c) Jacobi +
a vertex coarse space with an exact coarse solve, and with inexact inner solves.

    python oracle/precond_study.py

Output on the development container (p = 8, straight cells):
    n=8   full Jacobi 434   condensed Jacobi 116  condensed Jacobi+vertex coarse 32
    n=16  full Jacobi 869   condensed Jacobi 223  condensed Jacobi+vertex coarse 32
    n=32  full Jacobi 1682  condensed Jacobi 424  condensed Jacobi+vertex coarse 32
    n=64  full Jacobi 3382  condensed Jacobi 850  condensed Jacobi+vertex coarse 31
    curved n=64 inner rtol 1e-01: outer 41 (exact coarse: 37), inner its mean 45
    curved n=64 inner rtol 1e-02: outer 37 (exact coarse: 37), inner its mean 67
    curved n=64 inner rtol 1e-04: outer 37 (exact coarse: 37), inner its mean 135
"""
import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.abspath(__file__)))
import sem_oracle as so
from scipy import sparse
from scipy.sparse.linalg import splu

def pcg(A, b, apply_M, rtol=1e-12, maxiter=100000):
    x = np.zeros_like(b); r = b.copy(); z = apply_M(r); p = z.copy(); rz = r@z; bb = b@b; it=0
    while it < maxiter and r@r > rtol*rtol*bb:
        Ap = A@p; alpha = rz/(p@Ap); x += alpha*p; r -= alpha*Ap; z = apply_M(r); rzn = r@z; p = z + (rzn/rz)*p; rz = rzn; it += 1
    return x, it

p = 8
for n in (8, 16, 32, 64):
    basis = so.Basis(p)
    nodes, l2g = so.build_case("S", n, n, p, True, False)
    geo = so.geometry(basis, nodes, l2g)
    c = so.condensed_system(p, geo["invJ"], geo["JxW"], l2g)
    on, vals = so.dirichlet_data(nodes, l2g, geo["x_phys"], so.mesh_boundary_faces(n, n))
    n_ext = c["n_ext"]; free = ~on[:n_ext]
    S = c["Sg"][free][:, free].tocsr(); b = c["grhs"][free]
    # full system for comparison
    L = so.local_stiffness(basis, geo["invJ"], geo["JxW"])
    A = so.assemble_csr(L, l2g, nodes.shape[1]); fr = ~on
    Af = A[fr][:, fr].tocsr(); bf = so.assemble_vector(geo["JxW"], l2g, nodes.shape[1])[fr]
    d = 1.0/S.diagonal(); dA = 1.0/Af.diagonal()
    _, it_full = pcg(Af, bf, lambda r: dA*r)
    _, it_j = pcg(S, b, lambda r: d*r)
    # two-level: vertex coarse space (bilinear Q1 interpolation from element vertices to exterior nodes) + Jacobi, additive
    N = p+1
    ids = c["ids"]  # [E, 4p] hierarchical exterior: 4 vertices first
    # Q1 shape functions at the exterior nodes' parametric (equispaced) positions
    h = so.hier_order(N)[:4*p].astype(int); m = h//N; nn = h%N
    xi = m/(N-1.0); eta = nn/(N-1.0)
    # vertex order in hier: (0,0),(0,N-1),(N-1,0),(N-1,N-1)
    phi = np.stack([(1-xi)*(1-eta), (1-xi)*eta, xi*(1-eta), xi*eta], axis=1)  # [4p,4]
    vert = ids[:, :4]                    # global ids of element vertices
    vids = np.unique(vert); vmap = -np.ones(n_ext, int); vmap[vids] = np.arange(vids.size)
    rows = np.repeat(ids, 4, axis=1).ravel(); cols = vmap[np.tile(vert, (1, 1))]  # placeholder
    E = ids.shape[0]
    rows = np.repeat(ids[:, :, None], 4, axis=2).ravel()
    cols = np.repeat(vmap[vert][:, None, :], 4*p, axis=1).ravel()
    data = np.tile(phi[None], (E,1,1)).ravel()
    P = sparse.coo_matrix((data,(rows,cols)),shape=(n_ext, vids.size)).tocsr()
    # duplicates summed: divide each row by multiplicity (node shared by k elements gets k copies)
    mult = np.bincount(ids.ravel(), minlength=n_ext).astype(float)
    P = sparse.diags(1.0/mult) @ P
    Pf = P[free]
    keepc = np.asarray(Pf.sum(axis=0)).ravel() > 0
    # drop coarse dofs that sit on Dirichlet vertices
    onv = on[:n_ext][vids]
    Pf = Pf[:, ~onv]
    Ac = (Pf.T @ S @ Pf).tocsc(); lu = splu(Ac)
    M2 = lambda r: d*r + Pf @ lu.solve(Pf.T @ r)
    _, it_2 = pcg(S, b, M2)
    print("n=%d  full Jacobi %d  condensed Jacobi %d  condensed Jacobi+vertex coarse %d   (coarse dofs %d)" % (n, it_full, it_j, it_2, Ac.shape[0]), flush=True)

print("inexact coarse solves (inner Jacobi-PCG on the coarse matrix)")
for n in (32, 64):
    basis = so.Basis(p)
    nodes, l2g = so.build_case("C", n, n, p, True, False)
    geo = so.geometry(basis, nodes, l2g)
    c = so.condensed_system(p, geo["invJ"], geo["JxW"], l2g)
    on, vals = so.dirichlet_data(nodes, l2g, geo["x_phys"], so.mesh_boundary_faces(n, n))
    n_ext = c["n_ext"]; free = ~on[:n_ext]
    S = c["Sg"][free][:, free].tocsr(); b = c["grhs"][free]; d = 1.0/S.diagonal()
    ids = c["ids"]; E = ids.shape[0]
    h = so.hier_order(N)[:4*p].astype(int); m = h//N; nn = h%N
    xi = m/(N-1.0); eta = nn/(N-1.0)
    phi = np.stack([(1-xi)*(1-eta), (1-xi)*eta, xi*(1-eta), xi*eta], axis=1)
    vert = ids[:, :4]; vids = np.unique(vert); vmap = -np.ones(n_ext, int); vmap[vids] = np.arange(vids.size)
    rows = np.repeat(ids[:, :, None], 4, axis=2).ravel()
    cols = np.repeat(vmap[vert][:, None, :], 4*p, axis=1).ravel()
    data = np.tile(phi[None], (E,1,1)).ravel()
    P = sparse.coo_matrix((data,(rows,cols)),shape=(n_ext, vids.size)).tocsr()
    mult = np.bincount(ids.ravel(), minlength=n_ext).astype(float)
    P = sparse.diags(1.0/mult) @ P
    Pf = P[free][:, ~on[:n_ext][vids]]
    Ac = (Pf.T @ S @ Pf).tocsr(); dc = 1.0/Ac.diagonal(); lu = splu(Ac.tocsc())
    _, it_exact = pcg(S, b, lambda r: d*r + Pf @ lu.solve(Pf.T @ r))
    for tol in (1e-1, 1e-2, 1e-4):
        inner = []
        def M(r):
            rc = Pf.T @ r
            xc, itc = pcg(Ac, rc, lambda q: dc*q, rtol=tol)
            inner.append(itc)
            return d*r + Pf @ xc
        _, it_o = pcg(S, b, M)
        print("curved n=%d inner rtol %.0e: outer %d (exact coarse: %d), inner its mean %.0f" % (n, tol, it_o, it_exact, np.mean(inner)), flush=True)
