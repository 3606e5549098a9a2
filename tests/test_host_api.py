"""Host-side mirror of the reference API (tables, quadratures, bases, geometry,
meshes, DOF managers): tier T0 -- bit-exact against golden vectors frozen from
the live reference -- plus the analytic checks the reference's own (stale)
tests intended (tests/test_basis.py:52-163, tests/test_discrete.py:19-41)."""
import numpy as np
import pytest

from conftest import build_package_case, golden_case_names, load_case
from spectralelementmethod_b200 import basis_functions as bf
from spectralelementmethod_b200 import discrete, geometry, meshgen, quadratures
from spectralelementmethod_b200.mapping import _subface_slice


def test_gll_tables_bit_exact(golden_tables):
    for p in range(1, 11):
        b = bf.LagrangeGaussLobatto(p)
        assert np.array_equal(b.nodes, golden_tables["nodes_%d" % p])
        assert np.array_equal(b.bary_wts, golden_tables["bary_%d" % p])
        assert np.array_equal(b.quad_rule.weights, golden_tables["quad_%d" % p])
        assert np.array_equal(b.D1, golden_tables["D_%d" % p])
        assert np.array_equal(b.interp_eq_mat, golden_tables["E_%d" % p])


def test_gauss_lobatto_generator_bit_exact(golden_tables):
    for n in range(1, 13):
        g = quadratures.GaussLobatto(n)
        assert np.array_equal(g.abscissa, golden_tables["gl_x_%d" % n])
        assert np.array_equal(g.weights, golden_tables["gl_w_%d" % n])
        assert g.deg == 2 * n - 3
    with pytest.raises(ValueError):
        quadratures.GaussLobatto(0)
    with pytest.raises(ValueError):
        quadratures.GaussLobatto(2.5)


def test_order_limits():
    with pytest.raises(ValueError):
        bf.LagrangeGaussLobatto(0)
    with pytest.raises(NotImplementedError):
        bf.LagrangeGaussLobatto(11, allow_extended=False)   # the reference's max_order
    with pytest.raises(NotImplementedError):
        bf.LagrangeGaussLobatto(17)
    b = bf.LagrangeGaussLobatto(16)                           # fixture table
    assert abs(b.quad_rule.weights.sum() - 2.0) < 1e-14
    assert np.abs(b.D1 @ np.ones(17)).max() < 1e-11


@pytest.mark.parametrize("p", [1, 2, 3, 4, 7, 8, 10, 13, 16])
def test_basis_analytic_properties(p):
    b = bf.LagrangeGaussLobatto(p)
    N = p + 1
    # Kronecker delta at the nodes (tests/test_basis.py:52-58)
    assert np.array_equal(b(b.nodes), np.eye(N))
    # derivative of a constant vanishes; of x is one
    assert np.abs(b.D1 @ np.ones(N)).max() < 2e-12 * N
    assert np.abs(b.D1 @ b.nodes - 1.0).max() < 1e-11
    # weights integrate 1 and x+1 (tests/test_basis.py:99-105)
    assert abs(b.quad_rule.weights.sum() - 2.0) < 1e-13
    assert abs(b.integrate(b.nodes + 1.0) - 2.0) < 1e-13
    # GLL exactness up to degree 2p-1
    for k in range(0, 2 * p):
        exact = 0.0 if k % 2 else 2.0 / (k + 1)
        assert abs(b.quad_rule(b.nodes ** k) - exact) < 1e-12
    # E^{-1} really inverts the equispaced interpolation matrix
    assert np.abs(b.interp_eq_inv @ b.interp_eq_mat - np.eye(N)).max() < 1e-9 * max(1, p - 9) ** 4
    # interpolation of sin(pi x) (tests/test_basis.py:60-97)
    if p >= 8:
        x = np.linspace(-1, 1, 50)
        f = np.sin(np.pi * b.nodes)
        assert np.allclose(b.interpolate(f, x), np.sin(np.pi * x), rtol=1e-2, atol=1e-4)
        df = b.deriv(f)
        assert np.allclose(b.interpolate(df, x), np.pi * np.cos(np.pi * x), rtol=1e-2, atol=2e-3)


def test_interpolate_hits_nodes_and_scalars():
    b = bf.LagrangeGaussLobatto(5)
    f = np.arange(6.0)
    assert b.interpolate(f, np.float64(b.nodes[2])) == 2.0
    out = b.interpolate(np.stack([f, 2 * f]), b.nodes)
    assert np.allclose(out, np.stack([f, 2 * f]).T)


def test_tensor_product_basis():
    b5, b6 = bf.LagrangeGaussLobatto(5), bf.LagrangeGaussLobatto(6)
    tp = bf.TensorProductQS(b5, b6)
    assert tp.ndim == 2 and tp.coeff_shape == (6, 7) and tp.n_coeffs == 42 and tp.n_subbases == 2
    assert tp.get_subbasis(0) is b5 and tp.get_subbasis(1) is b6
    assert [d.shape for d in tp.get_D1_matrices()] == [(6, 6), (7, 7)]
    X, Y = tp.nodegrid()
    # Kronecker delta (tests/test_basis.py:121-128)
    V = tp((X.ravel(), Y.ravel()))
    assert np.allclose(V.reshape(42, 42), np.eye(42))
    # interpolation at random points of (xy, x+y) (tests/test_basis.py:130-139)
    rng = np.random.default_rng(0)
    xs, ys = rng.uniform(-1, 1, 50), rng.uniform(-1, 1, 50)
    coeffs = np.stack([X * Y, X + Y])
    vals = tp.interpolate(coeffs, (xs, ys))
    assert np.allclose(vals, np.stack([xs * ys, xs + ys]).T.T if vals.shape[0] == 2
                       else np.stack([xs * ys, xs + ys]).T, atol=1e-12)
    # grid interpolation 50 x 49 (tests/test_basis.py:141-147)
    gx, gy = np.linspace(-1, 1, 50), np.linspace(-1, 1, 49)
    G = tp.interpolate_on_grid(X * Y, (gx, gy))
    assert G.shape == (50, 49) and np.allclose(G, np.outer(gx, gy), atol=1e-12)
    # coefficient fitting from the equispaced grid and back (tests/test_basis.py:149-157)
    ex, ey = np.meshgrid(np.linspace(-1, 1, 6), np.linspace(-1, 1, 7), indexing="ij")
    c = tp.compute_coeffs_grid_eq(np.stack([ex * ey, ex + ey]))
    assert np.allclose(c, coeffs, atol=1e-12)
    assert np.allclose(tp.interpolate_on_grid_eq(c), np.stack([ex * ey, ex + ey]), atol=1e-12)
    c2 = tp.compute_coeffs_grid(np.stack([ex * ey, ex + ey]),
                                (np.linspace(-1, 1, 6), np.linspace(-1, 1, 7)))
    assert np.allclose(c2, coeffs, atol=1e-12)
    # gradient / integrate (tests/test_basis.py:159-163)
    g = tp.gradient(coeffs)
    assert g.shape == (2, 2, 6, 7)
    assert np.allclose(g[0, 0], Y, atol=1e-12) and np.allclose(g[1, 0], X, atol=1e-12)
    assert np.allclose(g[0, 1], 1.0, atol=1e-12) and np.allclose(g[1, 1], 1.0, atol=1e-12)
    assert abs(tp.integrate(X * X + Y) - 4.0 / 3.0) < 1e-13
    w = tp.quad_rule.xweight(np.ones((6, 7)))
    assert np.allclose(w, np.outer(b5.quad_rule.weights, b6.quad_rule.weights))
    with pytest.raises(ValueError):
        bf.TensorProductQS(b5, object())
    with pytest.raises(ValueError):
        tp((xs,))


def test_quadrature_classes():
    q = quadratures.Quadrature1D(np.array([-1.0, 0.0, 1.0]), np.array([1 / 3, 4 / 3, 1 / 3]))
    assert q.ndim == 1 and q.n_points == 3
    assert abs(q(lambda x: x ** 2) - 2 / 3) < 1e-15
    assert abs(q(np.array([1.0, 0.0, 1.0])) - 2 / 3) < 1e-15
    assert np.allclose(q.integrate(np.ones((3, 2, 2))), 2 * np.ones((2, 2)))
    assert np.allclose(q.xweight(np.ones(3)), q.weights)
    t = quadratures.TensorQuadratureRule(q, q)
    assert t.ndim == 2 and t.n_points == 9 and t.shape == (3, 3) and t.n_subquads == 2
    assert abs(t.integrate(np.ones((3, 3))) - 4.0) < 1e-15
    assert np.allclose(t.xweight(np.ones((3, 3))), np.outer(q.weights, q.weights))
    assert repr(q) == "Quadrature1D(n=3)"


def test_geometry_tables(golden_tables):
    for N in (2, 3, 5, 9, 11):
        q = geometry.Quadrilateral(N, N)
        assert np.array_equal(q.hierarchical_node_order, golden_tables["hier_%d" % N])
        assert q.hierarchical_node_order.dtype == np.uint32
        assert q.n_nodes == N * N and q.n_interior_nodes == (N - 2) ** 2
        assert q.n_sub_geometries() == 4 and q.n_sub_geometries(0) == 4
        assert isinstance(q.sub_geometry(0), geometry.Line)
    with pytest.raises(ValueError):
        geometry.Quadrilateral(3, 3).n_sub_geometries(3)
    arr = np.arange(2 * 4 * 5).reshape(2, 4, 5)
    for f in range(4):
        assert np.array_equal(_subface_slice(f, arr, 2), golden_tables["face_%d" % f])
    # faces are views, traversed counter-clockwise (SURVEY appendix B fixture)
    a9 = np.arange(9).reshape(3, 3)
    assert [_subface_slice(f, a9, 2).tolist() for f in range(4)] == \
        [[2, 1, 0], [6, 7, 8], [0, 3, 6], [8, 5, 2]]
    assert np.shares_memory(_subface_slice(1, a9, 2), a9)


@pytest.mark.parametrize("name", golden_case_names())
def test_l2g_maps_masks_and_hier_dofs_bit_exact(name):
    g = load_case(name)
    mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"])
    maps = mngr.node_map_array()
    assert maps.dtype == np.uint32
    assert np.array_equal(maps, g["l2g"])
    assert np.array_equal(mesh.nodes, g["nodes"])
    assert np.array_equal(mngr.boundary_node_mask("ebc"), g["on_ebc"])
    hier = np.stack([fe.global_dof_ind_hier for fe in mngr.finite_elements()])
    assert hier.dtype == np.uint32 and np.array_equal(hier, g["hier"])
    assert mngr.ndof == g["nodes"].shape[1]
    if g["sc"]:
        N = g["p"] + 1
        n_ext = np.unique(g["hier"][:, :N * N - (N - 2) ** 2]).size
        assert mngr.ndof_exterior == n_ext and mngr.ndof_interior == mngr.ndof - n_ext


def test_per_cell_api_equals_bulk_builder():
    """Mesh built cell by cell like tests/test_discrete.py:22-38 == bulk builder."""
    nx, ny, p = 3, 2, 4
    bulk = meshgen.structured_quad_mesh(nx, ny, p, "C")
    mesh = discrete.Mesh(2)
    mesh.set_nodes(meshgen.lattice_coordinates("C", nx, ny, p))
    gid = mesh.add_geometry(geometry.Quadrilateral(p + 1, p + 1))
    rid = mesh.new_region("interior")
    ebc, nbc = mesh.new_boundary("ebc"), mesh.new_boundary("nbc")
    NY = ny * p + 1
    ids = np.arange((nx * p + 1) * NY).reshape(nx * p + 1, NY)
    c = 0
    for ex in range(nx):
        for ey in range(ny):
            mesh.add_cell(ids[ex * p:ex * p + p + 1, ey * p:ey * p + p + 1], gid, rid)
            if ex == 0:
                mesh.add_boundary_cell(c, ebc, 1, 0)
            if ey == 0:
                mesh.add_boundary_cell(c, ebc, 1, 2)
            if ex == nx - 1:
                mesh.add_boundary_cell(c, nbc, 1, 1)
            if ey == ny - 1:
                mesh.add_boundary_cell(c, nbc, 1, 3)
            c += 1
    assert mesh.n_cells == bulk.n_cells == 6 and mesh.n_nodes == bulk.n_nodes
    assert np.array_equal(mesh.node_map_array(), bulk.node_map_array())
    assert np.array_equal(mesh.boundary_node_ind("ebc"), bulk.boundary_node_ind("ebc"))
    assert mesh.n_boundary_cells == bulk.n_boundary_cells
    cell = mesh.get_cell(3)
    assert cell.region_name == "interior" and cell.n_nodes == 25
    assert np.array_equal(cell.vertex_node_ind, cell.node_ind_lexicographic[[0, 0, -1, -1], [0, -1, 0, -1]])
    assert [c.node_ind_lexicographic[0, 0] for c in mesh.cells_on_boundary("nbc")] == \
        [c.node_ind_lexicographic[0, 0] for c in bulk.cells_on_boundary("nbc")]
    assert mesh.cells_are_neighbors(mesh.get_cell(0), mesh.get_cell(1)) == 3
    assert mesh.cells_are_neighbors(mesh.get_cell(0), mesh.get_cell(2)) == 1
    assert mesh.cells_are_neighbors(mesh.get_cell(0), mesh.get_cell(5)) == -1
    with pytest.raises(ValueError):
        mesh.set_nodes(np.zeros((3, 4)))


def test_permute_nodes_mutates_in_place_like_the_reference():
    coords = meshgen.lattice_coordinates("S", 2, 2, 2)
    mesh = meshgen.structured_quad_mesh(2, 2, 2, nodes=coords)
    view = mesh.get_cell(0).node_ind_lexicographic
    before = coords[:, view].copy()
    b1 = bf.LagrangeGaussLobatto(2)
    discrete.DOFManagerSC(mesh, 1, bf.TensorProductQS(b1, b1), rcm_order=True)
    assert mesh.nodes is coords                      # caller's array is permuted too
    assert np.array_equal(coords[:, view], before)   # cell views follow the renumbering
    assert mesh.condensed


def test_flag_validation_and_boundary_iteration_without_geometry():
    mesh, mngr = build_package_case("S", 2, 2, 3, True, False)
    with pytest.raises(ValueError):
        list(mngr.finite_elements(bogus=True))
    pairs = list(mngr.boundary_elements("ebc"))
    assert len(pairs) == 4                # 2 left faces + 2 bottom faces (corner cell has both)
    parent, sub = pairs[0]
    assert sub.n_nodes == 4 and sub.node_ind.shape == (4,)
    assert np.array_equal(sub.parent_dofs(), [3, 2, 1, 0])
    assert mngr.ndof_per_node == 1 and mngr.mesh is mesh


def test_schur_helpers_on_a_synthetic_local_system():
    """reorder_local_system_hier / compute_local_sc_system algebra
    (sem/discrete.py:428-476) on a random SPD local matrix."""
    mesh, mngr = build_package_case("S", 1, 1, 3, True, False)
    fe = next(mngr.finite_elements())
    rng = np.random.default_rng(1)
    M = rng.standard_normal((16, 16))
    A = M @ M.T + 16 * np.eye(16)
    rhs = rng.standard_normal(16)
    Ah, bh = mngr.reorder_local_system_hier(fe, (A, rhs))
    h = fe.loc_dof_ind_hier
    assert np.array_equal(Ah, A[np.ix_(h, h)]) and np.array_equal(bh, rhs[h])
    S, g = mngr.compute_local_sc_system(fe, (Ah, bh))
    ne = fe.ndof_exterior
    x = np.linalg.solve(Ah, bh)
    assert np.allclose(S @ x[:ne], g, atol=1e-10)
    gsys = mngr.init_global_linear_system()
    assert gsys[0].data.size == ne * ne and gsys[1].size == mngr.ndof_exterior
    mngr.assemble_global_sc_system(gsys, [(Ah, bh)])
    sol = np.zeros(16)
    mngr.solve(gsys, [(Ah, bh)], sol, np.zeros(ne, dtype=bool))
    assert np.allclose(sol[fe.global_dof_ind_hier], x, atol=1e-10)


def test_condensed_tables_host():
    """Tables of the device static-condensation path (condensed.condensed_tables):
    exterior L2G in hierarchical order and the node -> entries lists."""
    from spectralelementmethod_b200.condensed import condensed_tables
    g = load_case("C448_sc_rcm")
    mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"])
    N = g["p"] + 1
    geo = mesh.get_geometries()[0]
    ext_loc = np.asarray(geo.exterior_node_ind)
    assert np.array_equal(ext_loc, geo.hierarchical_node_order[:4 * g["p"]])
    l2g = mesh.node_map_array().reshape(-1, N * N)
    l2g_ext, ptr, pos = condensed_tables(l2g, ext_loc, mngr.ndof_exterior)
    # the live reference's fe.global_dof_ind_hier (golden), exterior part
    assert np.array_equal(l2g_ext, g["hier"][:, :4 * g["p"]])
    for e, fe in enumerate(mngr.finite_elements()):
        assert np.array_equal(l2g_ext[e], fe.global_dof_ind_hier[:fe.ndof_exterior])
    assert ptr[0] == 0 and ptr[-1] == l2g_ext.size and pos.size == l2g_ext.size
    flat = l2g_ext.ravel()
    for node in (0, 1, mngr.ndof_exterior // 2, mngr.ndof_exterior - 1):
        mine = pos[ptr[node]:ptr[node + 1]]
        assert (flat[mine] == node).all() and (np.diff(mine.astype(np.int64)) > 0).all()
        assert mine.size == (flat == node).sum()
    # scatter-add through the table == np.add.at
    loc = np.random.default_rng(0).standard_normal(flat.size)
    want = np.zeros(mngr.ndof_exterior)
    np.add.at(want, flat, loc)
    got = np.add.reduceat(loc[pos], ptr[:-1].astype(np.int64))
    assert np.allclose(got, want, rtol=0, atol=1e-13)


def _values_cases():
    import glob
    import os
    from conftest import GOLDEN
    return sorted(os.path.basename(p)[7:-4] for p in glob.glob(os.path.join(GOLDEN, "values_*.npz")))


@pytest.mark.parametrize("name", _values_cases())
def test_values_at_nodes_host_vs_reference(name):
    """DOFManager.values_at_nodes (sem/discrete.py:235-258) against the live
    reference's output (oracle/make_golden_values.py), two stacked fields."""
    import os
    from conftest import GOLDEN, rel_l2
    d = dict(np.load(os.path.join(GOLDEN, "values_%s.npz" % name)))
    nx, ny, p, sc, rcm, kind = d["meta"].tolist()
    mesh, mngr = build_package_case(chr(kind), nx, ny, p, bool(sc), bool(rcm))
    v = mngr.values_at_nodes(d["coeffs"])
    assert v.shape == d["values"].shape
    assert rel_l2(v, d["values"]) < 1e-14


def test_condensed_tables_reject_a_numbering_that_is_not_exterior_first():
    from spectralelementmethod_b200.condensed import condensed_tables
    mesh, mngr = build_package_case("S", 3, 2, 4, False, False)      # plain DOFManager: lexicographic
    N = 5
    geo = mesh.get_geometries()[0]
    l2g = mesh.node_map_array().reshape(-1, N * N)
    n_ext_true = mesh.n_nodes - 6 * 9
    with pytest.raises(AssertionError):
        condensed_tables(l2g, np.asarray(geo.exterior_node_ind), n_ext_true)


def test_multithreaded_host_helpers_are_bit_exact():
    """csrc/semk_hostnum.cpp (the path large meshes take) against the whole-array NumPy
    expressions (the path the golden cases take): same node maps, same permuted coordinates,
    same exterior / interior counts; structured node maps against the closed form."""
    import numpy as np
    from spectralelementmethod_b200 import discrete, meshgen
    from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS
    for kind, nx, ny, p in (("C", 90, 70, 4), ("S", 40, 44, 8)):
        ex, ey = np.divmod(np.arange(nx * ny, dtype=np.int64), ny)
        m = np.arange(p + 1, dtype=np.int64)
        closed = ((ex * p * (ny * p + 1) + ey * p)[:, None, None] + m[None, :, None] * (ny * p + 1)
                  + m[None, None, :]).astype(np.uint32)
        assert np.array_equal(meshgen.structured_node_maps(nx, ny, p), closed)
        b1 = LagrangeGaussLobatto(p)
        basis = TensorProductQS(b1, b1)
        fast_mesh = meshgen.structured_quad_mesh(nx, ny, p, kind)
        assert fast_mesh.n_nodes >= (1 << 16)
        calls = []
        orig = discrete.DOFManagerSC._fast_static_condensation

        def spy(self):
            ok = orig(self)
            calls.append(ok)
            return ok
        discrete.DOFManagerSC._fast_static_condensation = spy
        try:
            fast = discrete.DOFManagerSC(fast_mesh, 1, basis, rcm_order=False)
            discrete.DOFManagerSC._fast_static_condensation = lambda self: False
            slow_mesh = meshgen.structured_quad_mesh(nx, ny, p, kind)
            slow = discrete.DOFManagerSC(slow_mesh, 1, basis, rcm_order=False)
        finally:
            discrete.DOFManagerSC._fast_static_condensation = orig
        assert calls == [True]                      # the helper really ran
        assert np.array_equal(fast_mesh.node_map_array(), slow_mesh.node_map_array())
        assert np.array_equal(fast_mesh.nodes, slow_mesh.nodes)
        assert fast.ndof_exterior == slow.ndof_exterior and fast.ndof_interior == slow.ndof_interior
        # with RCM on top (exterior nodes only) the two paths still agree
        f2 = discrete.DOFManagerSC(meshgen.structured_quad_mesh(nx, ny, p, kind), 1, basis)
        discrete.DOFManagerSC._fast_static_condensation = lambda self: False
        try:
            s2 = discrete.DOFManagerSC(meshgen.structured_quad_mesh(nx, ny, p, kind), 1, basis)
        finally:
            discrete.DOFManagerSC._fast_static_condensation = orig
        assert np.array_equal(f2.mesh.node_map_array(), s2.mesh.node_map_array())


def test_lattice_coordinates_equal_the_meshgrid_recipe_bit_for_bit():
    """meshgen.lattice_coordinates writes the lattice in place (no meshgrid / vstack temporaries,
    the curved displacement from 1-D sines); the values must be those of the plain recipe of
    SURVEY appendix B: X, Y = meshgrid(...); s = 0.08 sin(pi X) sin(pi Y); X + s, Y + s."""
    for kind in "SC":
        for nx, ny, p, b in ((5, 7, 3, (-1.0, 1.0, -1.0, 1.0)), (16, 9, 8, (-1.0, 3.0, -1.0, 1.0)),
                             (4, 6, 10, (0.0, 2.5, -0.3, 0.9))):
            X, Y = np.meshgrid(np.linspace(b[0], b[1], nx * p + 1), np.linspace(b[2], b[3], ny * p + 1),
                               indexing="ij")
            if kind == "C":
                s = 0.08 * np.sin(np.pi * X) * np.sin(np.pi * Y)
                X, Y = X + s, Y + s
            want = np.vstack([X.ravel(), Y.ravel()])
            got = meshgen.lattice_coordinates(kind, nx, ny, p, b)
            assert got.shape == want.shape and np.array_equal(got, want)
    with pytest.raises(ValueError):
        meshgen.lattice_coordinates("X", 2, 2, 2)
