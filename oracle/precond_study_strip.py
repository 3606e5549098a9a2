"""Study (CPU, NumPy/SciPy): inner preconditioners for the vertex coarse operator on a long
strip partition.  The coarse operator of the condensed SEM system with linear edge
interpolation is spectrally a Q1 Laplacian on the element-vertex grid, which is what this
script uses.  Counts PCG iterations to a relative residual of 1e-2 (the inner tolerance of
the multilevel driver) for

  (a) Jacobi + piecewise-constant k x k tiles, exact dense inverse        (one GPU today)
  (b) the same with the larger tiles a replicated 4096-aggregate top level forces at 8 GPUs
  (c) Jacobi + RANK-LOCAL exact inverse of the k x k tile operator (block diagonal over the
      ranks) + a global dense inverse over super-tiles of s x s tiles
  (d) (c) without the super-tile level

    python oracle/precond_study_strip.py [elements per rank side] [ranks]
"""
import sys

import numpy as np
from scipy import sparse
from scipy.sparse.linalg import splu


def q1_laplacian(nx, ny):
    """Q1 stiffness on an (nx+1) x (ny+1) vertex grid of unit squares, Dirichlet on x = 0 and
    y = 0 (identity rows), vertex id = i * (ny + 1) + j."""
    ke = np.array([[4, -1, -1, -2], [-1, 4, -2, -1], [-1, -2, 4, -1], [-2, -1, -1, 4]]) / 6.0
    NY = ny + 1
    ex, ey = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    v = np.stack([ex * NY + ey, ex * NY + ey + 1, (ex + 1) * NY + ey, (ex + 1) * NY + ey + 1],
                 axis=-1).reshape(-1, 4)
    rows = np.repeat(v, 4, axis=1).ravel()
    cols = np.tile(v, (1, 4)).ravel()
    vals = np.tile(ke.ravel(), len(v))
    n = (nx + 1) * NY
    A = sparse.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    I, J = np.divmod(np.arange(n), NY)
    free = (I > 0) & (J > 0)
    M = sparse.diags(free.astype(float))
    return (M @ A @ M + sparse.diags((~free).astype(float))).tocsr(), free, I, J


def pcg(A, b, M, rtol, maxiter=2000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz, bb, it = r @ z, b @ b, 0
    while it < maxiter and r @ r > rtol * rtol * bb:
        Ap = A @ p
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        z = M(r)
        rzn = r @ z
        p = z + (rzn / rz) * p
        rz = rzn
        it += 1
    return x, it


def aggregation(tile, free, n_agg):
    ids = np.flatnonzero(free)
    return sparse.coo_matrix((np.ones(ids.size), (ids, tile[ids])),
                             shape=(free.size, n_agg)).tocsr()


def main():
    nxl = int(sys.argv[1]) if len(sys.argv) > 1 else 221
    ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    ny = nxl
    nx = nxl * ranks
    A, free, I, J = q1_laplacian(nx, ny)
    d = A.diagonal()
    rng = np.random.default_rng(0)
    b = np.where(free, rng.standard_normal(A.shape[0]), 0.0)

    def tiles(k):
        tx, ty = -(-nx // k), -(-ny // k)
        return (np.maximum(I - 1, 0) // k) * ty + np.maximum(J - 1, 0) // k, tx * ty, (tx, ty)

    def two_level(k):
        t, n_agg, _ = tiles(k)
        P = aggregation(t, free, n_agg)
        A3 = (P.T @ A @ P).tocsc()
        A3 = A3 + sparse.diags((A3.diagonal() == 0).astype(float))
        lu = splu(A3.tocsc())
        return lambda r: r / d + P @ lu.solve(P.T @ r), n_agg

    def block_local(k, s, with_super=True):
        t, n_agg, (tx, ty) = tiles(k)
        P = aggregation(t, free, n_agg)
        A3 = (P.T @ A @ P).tocsr()
        A3 = A3 + sparse.diags((A3.diagonal() == 0).astype(float))
        # rank of a tile: the rank that owns its first element column
        ti = np.arange(n_agg) // ty
        rank_of = np.minimum(ti * k // nxl, ranks - 1)
        blocks = []
        for r in range(ranks):
            sel = np.flatnonzero(rank_of == r)
            blocks.append((sel, splu(A3[sel][:, sel].tocsc())))
        # super-tiles: s x s tiles
        sx, sy = -(-tx // s), -(-ty // s)
        st = (ti // s) * sy + (np.arange(n_agg) % ty) // s
        P3 = sparse.coo_matrix((np.ones(n_agg), (np.arange(n_agg), st)), shape=(n_agg, sx * sy)).tocsr()
        A4 = (P3.T @ A3 @ P3).tocsc()
        lu4 = splu(A4)

        def M(r):
            r3 = P.T @ r
            y3 = np.zeros(n_agg)
            for sel, lu in blocks:
                y3[sel] = lu.solve(r3[sel])
            if with_super:
                y3 += P3 @ lu4.solve(P3.T @ r3)
            return r / d + P @ y3
        return M, n_agg, sx * sy

    print("strip %d x %d elements (%d ranks), %d vertices" % (nx, ny, ranks, A.shape[0]))
    _, itj = pcg(A, b, lambda r: r / d, 1e-2)
    print("Jacobi only: %d iterations" % itj)
    for k in (8, 16, 24, 40):
        M, n_agg = two_level(k)
        _, it = pcg(A, b, M, 1e-2)
        print("(a/b) global tiles k=%2d (%6d aggregates, dense %7.1f MB): %3d iterations"
              % (k, n_agg, n_agg * n_agg * 8 / 1e6, it))
    for k, s in ((8, 4), (16, 2), (16, 4), (16, 8), (24, 4)):
        M, n_agg, n_super = block_local(k, s, True)
        _, it = pcg(A, b, M, 1e-2)
        M2, _, _ = block_local(k, s, False)
        _, it2 = pcg(A, b, M2, 1e-2)
        print("(c) rank-local tiles k=%2d (%5d per rank, dense %6.1f MB per rank) + super-tiles "
              "s=%d (%5d global): %3d iterations; (d) without super-tiles: %3d"
              % (k, n_agg // ranks, (n_agg / ranks) ** 2 * 8 / 1e6, s, n_super, it, it2))


if __name__ == "__main__":
    main()
