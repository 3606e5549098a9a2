"""Gmsh 2.2 importer vs the reference loader's frozen output (tests/golden/msh_*.npz,
made by oracle/make_golden_msh.py from the unmodified sem.grid_importers.load_msh)."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT
from spectralelementmethod_b200 import discrete, grid_importers as gi, meshgen
from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS

GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "msh_*.npz")))


def _load(tmp_path, z):
    path = tmp_path / "m.msh"
    path.write_bytes(z["msh_bytes"].tobytes())
    return gi.load_msh(str(path), 2)


def _boundary_rows(mesh):
    rows = []
    for cell in sorted(mesh._boundary_map):
        for bnd_id, lst in mesh._boundary_map[cell].items():
            for bd in lst:
                rows.append((cell, bnd_id, bd.index, bd.ndim))
    return np.array(rows, dtype=np.int64).reshape(-1, 4)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[4:-4] for p in GOLDEN])
def test_load_msh_matches_reference(tmp_path, path):
    z = np.load(path)
    mesh = _load(tmp_path, z)
    # bit-exact: coordinates, L2G maps (Gmsh -> lexicographic conversion), region ids
    assert np.array_equal(mesh.nodes, z["nodes"])
    maps = np.stack([c.node_ind_lexicographic for c in mesh.cells])
    assert maps.dtype == np.uint32 and np.array_equal(maps, z["node_maps"])
    assert np.array_equal([c.region_id for c in mesh.cells], z["region_ids"])
    assert list(mesh._region_names) == z["region_names"].tolist()
    assert list(mesh._boundary_names) == z["boundary_names"].tolist()
    # adjacency found by face hashing == the reference's centroid search
    adj = np.array([[-1 if v is None else v for v in mesh.get_cell(i)._adj_map]
                    for i in range(mesh.n_cells)])
    assert np.array_equal(adj, z["adjacency"])
    for i in range(mesh.n_cells):
        for face in range(4):
            nb = mesh.get_cell(i).neighbor(face)
            assert (nb is None) == (z["adjacency"][i, face] < 0)
    # boundary faces, including the per-cell registration order (nearest boundary
    # element first; exact ties may be ordered either way by the reference's argsort)
    got, want = _boundary_rows(mesh), z["boundary"]
    assert sorted(map(tuple, got)) == sorted(map(tuple, want))
    if not np.array_equal(got, want):
        for cell in np.unique(want[:, 0]):
            g, w = got[got[:, 0] == cell], want[want[:, 0] == cell]
            assert sorted(map(tuple, g)) == sorted(map(tuple, w))


@pytest.mark.parametrize("path", GOLDEN[:3], ids=[os.path.basename(p)[4:-4] for p in GOLDEN[:3]])
def test_loaded_mesh_drives_the_dof_manager(tmp_path, path):
    """The imported mesh is a full citizen: a DOF manager on it yields the same masks as
    one on the directly built mesh, up to the file's node permutation."""
    z = np.load(path)
    nx, ny, p, seed = (int(v) for v in z["meta"])
    mesh = _load(tmp_path, z)
    b1 = LagrangeGaussLobatto(p)
    mngr = discrete.DOFManager(mesh, 1, TensorProductQS(b1, b1), rcm_order=False)
    on_ebc = mngr.boundary_node_mask("ebc")
    x, y = mesh.nodes
    lo_x, lo_y = x.min(), y.min()
    assert np.array_equal(on_ebc, (x == lo_x) | (y == lo_y))
    on_nbc = mngr.boundary_node_mask("nbc")
    assert np.array_equal(on_nbc, (x == x.max()) | (y == y.max()))


def test_gmsh_ordering_tables():
    # a line: the two ends, then the interior in order
    assert gi.gmsh_to_lexicographic((5,)).tolist() == [0, 2, 3, 4, 1]
    # bilinear quad: counter-clockwise from (0, 0); lexicographic = [i, j] row-major
    assert gi.gmsh_to_lexicographic((2, 2)).tolist() == [0, 3, 1, 2]
    # biquadratic: corners 0-3, edges 4-7 (south, east, north, west), centre 8
    assert gi.gmsh_to_lexicographic((3, 3)).reshape(3, 3).tolist() == [[0, 7, 3], [4, 8, 6], [1, 5, 2]]
    # every table is a permutation, and the interior of order p is the table of order p - 2
    for n in range(2, 12):
        t = gi.gmsh_to_lexicographic((n, n)).reshape(n, n)
        assert sorted(t.ravel().tolist()) == list(range(n * n))
        if n >= 4:
            inner = gi.gmsh_to_lexicographic((n - 2, n - 2)).reshape(n - 2, n - 2)
            assert np.array_equal(t[1:-1, 1:-1] - (4 * n - 4), inner)


def test_single_cell_and_error_paths(tmp_path):
    # one cell: no neighbours, four boundary faces (the reference's search fails here)
    path = meshgen.write_gmsh22_binary(str(tmp_path / "one.msh"), 1, 1, 4)
    mesh = gi.load_msh(path, 2)
    assert mesh.n_cells == 1 and mesh.get_cell(0)._adj_map == [None] * 4
    assert sorted(r[2] for r in _boundary_rows(mesh)) == [0, 1, 2, 3]
    raw = open(path, "rb").read()
    bad = tmp_path / "bad.msh"
    bad.write_bytes(raw.replace(b"2.2 1 8", b"4.1 1 8", 1))
    with pytest.raises(gi.FileFormatError):
        gi.load_msh(str(bad), 2)
    bad.write_bytes(raw.replace(b"2.2 1 8\n\x01\x00\x00\x00\n", b"2.2 0 8\n", 1))
    with pytest.raises(NotImplementedError):
        gi.load_msh(str(bad), 2)
    bad.write_bytes(raw.replace(b"$MeshFormat", b"$Meshformat", 1))
    with pytest.raises(gi.FileFormatError):
        gi.load_msh(str(bad), 2)
    bad.write_bytes(raw.replace(b"$EndNodes", b"$EndNodez", 1))
    with pytest.raises(gi.FileFormatError):
        gi.load_msh(str(bad), 2)
    bad.write_bytes(raw[:len(raw) // 2])
    with pytest.raises((gi.FileFormatError, ValueError)):
        gi.load_msh(str(bad), 2)


def test_large_file_is_fast(tmp_path):
    """256 x 256 cells of order 4: the face-hash neighbour search and the blockwise
    reordering finish in seconds (the reference's element loop + all-pairs centroid
    search needs O(E^2) distance evaluations: minutes to hours at this size)."""
    import time
    nx = ny = 256
    path = meshgen.write_gmsh22_binary(str(tmp_path / "big.msh"), nx, ny, 4, "C", shuffle_seed=5)
    t0 = time.perf_counter()
    mesh = gi.load_msh(path, 2)
    el = time.perf_counter() - t0
    assert mesh.n_cells == nx * ny and el < 30.0
    adj = mesh._adj_array
    assert (adj >= 0).sum() == 2 * (nx * (ny - 1) + ny * (nx - 1))
    assert len(_boundary_rows(mesh)) == 2 * (nx + ny)


def test_construct_geometry_table_matches_the_reference_keys():
    """sem/grid_importers.py:19-42: Gmsh element type -> geometry constructor."""
    from spectralelementmethod_b200 import geometry as geo
    table = gi.construct_geometry
    assert sorted(table) == sorted([1, 8, 26, 27, 28, 62, 63, 64, 65, 66,
                                    3, 10, 36, 37, 38, 47, 48, 49, 50, 51])
    line = table[27]()
    quad = table[49]()
    assert isinstance(line, geo.Line) and line.shape == (5,)
    assert isinstance(quad, geo.Quadrilateral) and quad.shape == (9, 9)


@pytest.mark.parametrize("per_header", [1, 3])
def test_any_element_blocking_gives_one_homogeneous_mesh(tmp_path, per_header):
    """One $Elements header per element (or per few elements) must load into the same,
    homogeneous mesh as whole-block files: node_map_array / boundary masks / the DOF
    managers work on it (the reference keeps one array per cell and accepts any blocking,
    sem/discrete.py:1031-1048)."""
    nx, ny, p = 5, 4, 3
    whole = gi.load_msh(meshgen.write_gmsh22_binary(str(tmp_path / "a.msh"), nx, ny, p, "C",
                                                     shuffle_seed=4), 2)
    split = gi.load_msh(meshgen.write_gmsh22_binary(str(tmp_path / "b.msh"), nx, ny, p, "C",
                                                     shuffle_seed=4,
                                                     elements_per_header=per_header), 2)
    assert split._is_homogeneous() and split.n_cells == nx * ny
    assert np.array_equal(split.node_map_array(), whole.node_map_array())
    assert np.array_equal(split.nodes, whole.nodes)
    assert np.array_equal(_boundary_rows(split), _boundary_rows(whole))
    b1 = LagrangeGaussLobatto(p)
    m1 = discrete.DOFManagerSC(split, 1, TensorProductQS(b1, b1), rcm_order=True)
    m2 = discrete.DOFManagerSC(whole, 1, TensorProductQS(b1, b1), rcm_order=True)
    assert np.array_equal(m1.node_map_array(), m2.node_map_array())
    assert np.array_equal(m1.boundary_node_mask("ebc"), m2.boundary_node_mask("ebc"))


def test_blocks_of_one_geometry_merge_whatever_the_call_pattern():
    from spectralelementmethod_b200.geometry import Quadrilateral
    nx, ny, p = 3, 3, 2
    maps = meshgen.structured_node_maps(nx, ny, p).reshape(nx * ny, p + 1, p + 1)
    mesh = discrete.Mesh(2)
    mesh.set_nodes(meshgen.lattice_coordinates("S", nx, ny, p, (-1.0, 1.0, -1.0, 1.0)))
    g = mesh.add_geometry(Quadrilateral(p + 1, p + 1))
    r = mesh.new_region("interior")
    mesh.add_cells(maps[:4], g, r)                     # two bulk blocks ...
    mesh.add_cells(maps[4:6], g, r)
    first = mesh.get_cell(0).node_ind_lexicographic.copy()
    for k in range(6, 9):                              # ... and add_cell after a get_cell
        mesh.add_cell(maps[k], g, r)
    assert mesh._is_homogeneous() and mesh.n_cells == 9
    assert np.array_equal(mesh.node_map_array(), maps)
    assert np.array_equal(mesh.get_cell(0).node_ind_lexicographic, first)
    assert np.array_equal(mesh.get_cell(7).node_ind_lexicographic, maps[7])
    assert [c.region_id for c in mesh.cells] == [r] * 9
