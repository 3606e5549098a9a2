"""Host tables of the two-level preconditioner (condensed.coarse_tables) and a NumPy
emulation of the device algorithm on the same tables: Jacobi + vertex coarse space on
the condensed system (csrc/semk_sc.cu coarse kernels, semk_sc_pcg2_solve_f64)."""
import numpy as np
import pytest

import sem_oracle as so
from conftest import load_case, rel_l2
from spectralelementmethod_b200.condensed import (aggregate_tables, coarse_tables, condensed_tables,
                                                   element_tiles, top_level_inverse)


def _pcg(apply, b, M, rtol, x0=None, maxiter=100000):
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - apply(x)
    z = M(r)
    p = z.copy()
    rz, bb, it = r @ z, b @ b, 0
    while it < maxiter and r @ r > rtol * rtol * bb:
        Ap = apply(p)
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        z = M(r)
        rzn = r @ z
        p = z + (rzn / rz) * p
        rz = rzn
        it += 1
    return x, it


def emulate(p, c, on, vals, gll, tile=None, extra=None):
    """The device algorithm in NumPy on the product's host tables; returns
    (x_jacobi, its_jacobi, x_two_level, its_two_level, tables, Ace)."""
    N, NE = p + 1, 4 * p
    n_ext = c["n_ext"]
    ids, S_e = c["ids"], c["S"]
    E = ids.shape[0]
    l2g_ext, nptr, npos = condensed_tables(
        np.pad(ids, ((0, 0), (0, N * N - NE))).astype(np.uint32), np.arange(NE), n_ext)
    assert np.array_equal(l2g_ext.astype(np.int64), ids)
    nptr = nptr.astype(np.int64)
    D = on[:n_ext]
    ct = coarse_tables(l2g_ext, nptr, npos, D, gll)

    def fine(u):
        yl = np.einsum("ekj,ej->ek", S_e, np.where(D, 0.0, u)[ids]).ravel()
        return np.where(D, u, np.add.reduceat(yl[npos], nptr[:-1]))
    vc = ct["vert_c"].astype(np.int64)
    Dc = ct["dirichlet_c"]
    Phi = ct["phi"][None] * (~D)[ids][:, :, None] * (~Dc)[vc][:, None, :]
    Ace = np.einsum("eka,ekj,ejc->eac", Phi, S_e, Phi)
    vptr, vpos = ct["vptr"].astype(np.int64), ct["vpos"].astype(np.int64)

    def coarse(xc):
        yl = np.einsum("eac,ec->ea", Ace, xc[vc]).ravel()
        return np.where(Dc, xc, np.add.reduceat(yl[vpos], vptr[:-1]))
    dc = np.where(Dc, 1.0, np.add.reduceat(np.einsum("eaa->ea", Ace).ravel()[vpos], vptr[:-1]))
    pv, pw = ct["pv"].astype(np.int64), ct["pw"]
    rptr, ridx, rw = ct["rptr"].astype(np.int64), ct["ridx"].astype(np.int64), ct["rw"]

    def restrict(r):
        out = np.zeros(ct["n_v"])
        nz = rptr[1:] > rptr[:-1]
        out[nz] = np.add.reduceat(rw * r[ridx], rptr[:-1][nz])
        return out

    def prolong(xc):
        return pw[:, 0] * xc[pv[:, 0]] + pw[:, 1] * xc[pv[:, 1]]
    sdiag = np.add.reduceat(np.einsum("ekk->ek", S_e).ravel()[npos], nptr[:-1])
    dinv = 1.0 / np.where(D, 1.0, sdiag)

    inner2 = []

    def two_level(r):
        xc, itc = _pcg(coarse, restrict(r), lambda q: q / dc, 1e-2)
        inner2.append(itc)
        return dinv * r + prolong(xc)
    gv = np.where(D, vals[:n_ext], 0.0)
    yl = np.einsum("ekj,ej->ek", S_e, gv[ids]).ravel()
    b = c["grhs"] - np.where(D, 0.0, np.add.reduceat(yl[npos], nptr[:-1]))
    b[D] = gv[D]
    x0 = np.where(D, b, 0.0)
    xj, itj = _pcg(fine, b, lambda r: dinv * r, 1e-12, x0)
    x2, it2 = _pcg(fine, b, two_level, 1e-12, x0)
    if tile is not None:
        # third level: Jacobi + aggregation with a dense inverse inside the coarse solve
        at = aggregate_tables(ct["vert_c"], ct["vptr"], ct["vpos"], Dc, tile)
        A3inv = top_level_inverse(Ace, ct["vert_c"], at["agg"], at["n_agg"])
        agg = at["agg"].astype(np.int64)
        valid = agg != 0xFFFFFFFF
        aptr, aidx = at["aptr"].astype(np.int64), at["aidx"].astype(np.int64)
        inner3 = []

        def coarse_precond(q):
            y3 = A3inv @ np.add.reduceat(q[aidx], aptr[:-1])
            z = q / dc
            z[valid] += y3[agg[valid]]
            return z

        def three_level(r):
            xc, itc = _pcg(coarse, restrict(r), coarse_precond, 1e-2)
            inner3.append(itc)
            return dinv * r + prolong(xc)
        x3, it3 = _pcg(fine, b, three_level, 1e-12, x0)
        extra.update(x3=x3, it3=it3, inner3=sum(inner3), inner2=sum(inner2), at=at, A3inv=A3inv)
    return xj, itj, x2, it2, ct, Ace, (restrict, prolong)


def test_coarse_tables_structure_and_transpose():
    g = load_case("C448_sc_rcm")
    p = int(g["p"])
    c = so.condensed_system(p, g["invJ"], g["JxW"], g["l2g"])
    xj, itj, x2, it2, ct, Ace, (restrict, prolong) = emulate(
        p, c, g["on_ebc"], g["ebc_vals"], so.Basis(p).nodes)
    n_ext, n_v = c["n_ext"], ct["n_v"]
    assert n_v == (g["nx"] + 1) * (g["ny"] + 1)
    assert ct["phi"].shape == (4 * p, 4) and np.allclose(ct["phi"].sum(axis=1), 1.0)
    assert ct["pv"].shape == (n_ext, 2) and ct["pw"].shape == (n_ext, 2)
    D = g["on_ebc"][:n_ext]
    assert not ct["pw"][D].any()                          # Dirichlet rows are empty
    assert not ct["pw"][ct["dirichlet_c"][ct["pv"].astype(int)]].any()
    free_rows = ~D
    full = ct["pw"].sum(axis=1)[free_rows]
    assert (full <= 1.0 + 1e-15).all() and (full > 0).any()
    rng = np.random.default_rng(0)
    r, xc = rng.standard_normal(n_ext), rng.standard_normal(n_v)
    assert abs(restrict(r) @ xc - r @ prolong(xc)) < 1e-12 * np.linalg.norm(r) * np.linalg.norm(xc)
    assert np.allclose(Ace, np.swapaxes(Ace, 1, 2), rtol=0, atol=1e-12)
    # both preconditioners reach the reference's solution; the coarse space cuts the iterations
    assert rel_l2(x2, g["solution"][:n_ext]) < 1e-11 and rel_l2(xj, g["solution"][:n_ext]) < 1e-11
    assert it2 < itj


@pytest.mark.parametrize("n,limit", [(8, 40), (16, 40)])
def test_two_level_iterations_do_not_grow_with_the_mesh(n, limit):
    p = 4
    basis = so.Basis(p)
    nodes, l2g = so.build_case("C", n, n, p, True, False)
    geo = so.geometry(basis, nodes, l2g)
    c = so.condensed_system(p, geo["invJ"], geo["JxW"], l2g)
    on, vals = so.dirichlet_data(nodes, l2g, geo["x_phys"], so.mesh_boundary_faces(n, n))
    xj, itj, x2, it2, ct, Ace, _ = emulate(p, c, on, vals, basis.nodes)
    assert it2 <= limit and it2 < itj
    assert rel_l2(x2, xj) < 1e-10


def test_prolongation_reproduces_linear_functions_on_straight_edges():
    p, n = 5, 4
    basis = so.Basis(p)
    nodes, l2g = so.build_case("S", n, n, p, True, False)
    geo = so.geometry(basis, nodes, l2g)
    c = so.condensed_system(p, geo["invJ"], geo["JxW"], l2g)
    N, NE = p + 1, 4 * p
    ids = c["ids"]
    l2g_ext, nptr, npos = condensed_tables(
        np.pad(ids, ((0, 0), (0, N * N - NE))).astype(np.uint32), np.arange(NE), c["n_ext"])
    ct = coarse_tables(l2g_ext, nptr, npos, None, basis.nodes)
    # physical GLL coordinates of the exterior nodes
    h = so.hier_order(N)[:NE].astype(int)
    xg = np.zeros((2, c["n_ext"]))
    xg[:, ids.ravel()] = np.moveaxis(geo["x_phys"].reshape(ids.shape[0], 2, N * N)[:, :, h], 1, 0
                                     ).reshape(2, -1)
    lin = 0.3 + 1.7 * xg[0] - 0.6 * xg[1]
    vids = np.unique(ids[:, :4])
    pv, pw = ct["pv"].astype(int), ct["pw"]
    got = pw[:, 0] * lin[vids][pv[:, 0]] + pw[:, 1] * lin[vids][pv[:, 1]]
    assert np.allclose(got, lin, rtol=0, atol=1e-13)


@pytest.mark.parametrize("n_cells,rings", [(3, 2), (7, 3)])
def test_two_level_emulation_on_irregular_vertex_valence(n_cells, rings):
    """Pinwheel meshes (3 / 7 cells around the centre vertex): the edge-based prolongation
    and the vertex -> entries lists cope with any valence."""
    from spectralelementmethod_b200 import discrete, meshgen
    from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS
    p = 4
    mesh = meshgen.pinwheel_mesh(n_cells, p, rings=rings)
    b1 = LagrangeGaussLobatto(p)
    mngr = discrete.DOFManagerSC(mesh, 1, TensorProductQS(b1, b1), rcm_order=True)
    on = mngr.boundary_node_mask("ebc")
    l2g = mngr.node_map_array()
    geo = so.geometry(so.Basis(p), mesh.nodes, l2g)
    c = so.condensed_system(p, geo["invJ"], geo["JxW"], l2g)
    x, y = mesh.nodes
    vals = np.where(on, 0.3 * x - 0.2 * y + 0.1, 0.0)
    xj, itj, x2, it2, ct, Ace, (restrict, prolong) = emulate(p, c, on, vals, so.Basis(p).nodes)
    assert ct["n_v"] == np.unique(c["ids"][:, :4]).size
    assert n_cells in np.diff(ct["vptr"].astype(np.int64))           # the centre vertex
    assert rel_l2(x2, xj) < 1e-10 and it2 <= itj


def test_three_level_emulation_cuts_the_inner_iterations():
    """Aggregation level under the vertex coarse space (condensed.element_tiles /
    aggregate_tables / top_level_inverse): same outer count and solution, far fewer inner
    iterations.  NumPy emulation of semk_sc_pcg3_solve_f64 on the product's host tables."""
    p, n = 4, 24
    basis = so.Basis(p)
    nodes, l2g = so.build_case("C", n, n, p, True, False)
    geo = so.geometry(basis, nodes, l2g)
    c = so.condensed_system(p, geo["invJ"], geo["JxW"], l2g)
    on, vals = so.dirichlet_data(nodes, l2g, geo["x_phys"], so.mesh_boundary_faces(n, n))

    class M(object):                      # what element_tiles reads from a mesh
        _structured_shape = (n, n)
    tile = element_tiles(M(), n * n, max_tiles=36)           # 4 x 4 element tiles
    assert np.unique(tile).size == 36
    extra = {}
    xj, itj, x2, it2, ct, Ace, _ = emulate(p, c, on, vals, basis.nodes, tile=tile, extra=extra)
    at = extra["at"]
    free = ~ct["dirichlet_c"]
    assert (at["agg"][~free] == 0xFFFFFFFF).all() and (at["agg"][free] < at["n_agg"]).all()
    assert at["aptr"][-1] == free.sum() and np.array_equal(np.sort(at["aidx"]), np.flatnonzero(free))
    assert np.allclose(extra["A3inv"], extra["A3inv"].T)
    assert abs(extra["it3"] - it2) <= 2 and extra["inner3"] * 2 < extra["inner2"]
    assert rel_l2(extra["x3"], x2) < 1e-10
