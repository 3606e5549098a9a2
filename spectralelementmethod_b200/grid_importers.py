"""Gmsh 2.2 binary mesh importer (drop-in for ``sem.grid_importers``).

Mirrors the reference's ``load_msh(file_path, ndim)`` (sem/grid_importers.py:45-69):
same sections, same error types, same resulting ``Mesh`` -- regions and
boundaries from ``$PhysicalNames``, node coordinates, cells with their node
ids converted from Gmsh's recursive corner/edge/interior ordering to the
lexicographic order the rest of the package uses, boundary faces and the
cell-to-cell adjacency.

Two things are done differently:
  * the ordering conversion is one fancy-index per element BLOCK (the
    reference converts element by element in Python,
    sem/grid_importers.py:199-209 and :273-332);
  * neighbours and boundary faces come from a hash of the faces' vertex pairs,
    O(E log E), instead of the reference's all-pairs centroid search
    (sem/grid_importers.py:221-270, O(E^2) distances).  Boundary faces of a
    cell are registered in the reference's order: nearest boundary-cell
    centroid first.
"""
import numpy as np

from . import discrete
from . import geometry as geo

__all__ = ["FileFormatError", "load_msh", "parse_format", "gmsh_to_lexicographic",
           "GMSH_LINE_TYPES", "GMSH_QUAD_TYPES", "construct_geometry"]


class FileFormatError(Exception):
    """Raised when a mesh file is malformed (sem/grid_importers.py:9-12)."""


# Gmsh element type -> points per direction (sem/grid_importers.py:19-42)
GMSH_LINE_TYPES = {1: 2, 8: 3, 26: 4, 27: 5, 28: 6, 62: 7, 63: 8, 64: 9, 65: 10, 66: 11}
GMSH_QUAD_TYPES = {3: 2, 10: 3, 36: 4, 37: 5, 38: 6, 47: 7, 48: 8, 49: 9, 50: 10, 51: 11}


def _make_geometry(elem_type):
    if elem_type in GMSH_LINE_TYPES:
        return geo.Line(GMSH_LINE_TYPES[elem_type])
    if elem_type in GMSH_QUAD_TYPES:
        n = GMSH_QUAD_TYPES[elem_type]
        return geo.Quadrilateral(n, n)
    raise KeyError(elem_type)


def _geometry_factory(elem_type):
    return lambda: _make_geometry(elem_type)


# The reference's public table (sem/grid_importers.py:19-42): Gmsh element type ->
# zero-argument constructor of the matching geometry object.
construct_geometry = {t: _geometry_factory(t) for t in list(GMSH_LINE_TYPES) + list(GMSH_QUAD_TYPES)}


_ORDER_CACHE = {}


def gmsh_to_lexicographic(shape):
    """``idx`` with ``lex.ravel() = gmsh[idx]``: position in Gmsh's node list of
    every node of the lexicographic ``shape`` grid (first index = first
    parametric direction).  Gmsh numbers the 4 corners counter-clockwise from
    (0, 0), then the edges in the same sense (interior edge nodes following the
    sense of travel), then the interior grid recursively in the same way; a
    line is end, end, interior (sem/grid_importers.py:273-332)."""
    shape = tuple(int(s) for s in shape)
    if shape in _ORDER_CACHE:
        return _ORDER_CACHE[shape]
    if len(shape) == 0:
        idx = np.zeros(1, dtype=np.intp)
    elif len(shape) > 2:
        raise NotImplementedError("Can only take 2 arguments for now...")
    else:
        M, N = (shape[0], 1) if len(shape) == 1 else shape
        visit = []                      # (i, j) in Gmsh order
        lo_i, hi_i, lo_j, hi_j = 0, M - 1, 0, N - 1
        while lo_i < hi_i and lo_j < hi_j:
            visit += [(lo_i, lo_j), (hi_i, lo_j), (hi_i, hi_j), (lo_i, hi_j)]
            visit += [(i, lo_j) for i in range(lo_i + 1, hi_i)]             # south, i up
            visit += [(hi_i, j) for j in range(lo_j + 1, hi_j)]             # east, j up
            visit += [(i, hi_j) for i in range(hi_i - 1, lo_i, -1)]         # north, i down
            visit += [(lo_i, j) for j in range(hi_j - 1, lo_j, -1)]         # west, j down
            lo_i, hi_i, lo_j, hi_j = lo_i + 1, hi_i - 1, lo_j + 1, hi_j - 1
        if lo_i == hi_i and lo_j == hi_j:
            visit.append((lo_i, lo_j))                                      # centre node
        elif lo_j == hi_j and lo_i < hi_i:                                  # line along i
            visit += [(lo_i, lo_j), (hi_i, lo_j)] + [(i, lo_j) for i in range(lo_i + 1, hi_i)]
        elif lo_i == hi_i and lo_j < hi_j:                                  # line along j
            visit += [(lo_i, lo_j), (lo_i, hi_j)] + [(lo_i, j) for j in range(lo_j + 1, hi_j)]
        idx = np.empty(M * N, dtype=np.intp)
        for k, (i, j) in enumerate(visit):
            idx[i * N + j] = k
        assert len(visit) == M * N
    _ORDER_CACHE[shape] = idx
    return idx


def parse_format(f):
    """Check the ``$MeshFormat`` section; returns (is_binary, position after
    it) (sem/grid_importers.py:72-102)."""
    if not f.readline().startswith(b"$MeshFormat"):
        raise FileFormatError("Expected 'MeshFormat' data")
    fields = f.readline().split()
    if len(fields) != 3:
        raise FileFormatError("Unable to recognize file format")
    version, is_binary, data_size = fields
    if version != b"2.2":
        raise FileFormatError("Expected Gmsh file format version 2.2, but"
                              "got {} instead".format(version.decode("utf-8")))
    if is_binary not in (b"0", b"1"):
        raise FileFormatError("Unable to recognize file format")
    is_binary = bool(int(is_binary))
    if data_size != b"8":
        raise FileFormatError("Expected a data size of 8, but got"
                              "{} instead".format(data_size.decode("utf-8")))
    if is_binary:
        f.readline()                    # the binary "1" (endianness probe) and its newline
    if not f.readline().startswith(b"$EndMeshFormat"):
        raise FileFormatError("Malformed mesh format specification")
    return is_binary, f.tell()


def _parse_physical_names(f, mesh):
    """-> (region_id_map, boundary_id_map, boundary names) keyed by the 0-based
    physical id (sem/grid_importers.py:105-133)."""
    if not f.readline().startswith(b"$PhysicalNames"):
        raise FileFormatError("Expected 'PhysicalNames' data")
    n_names = int(f.readline().rstrip())
    region_id_map, boundary_id_map = {}, {}
    for i in range(n_names):
        fields = f.readline().split()
        ndim = int(fields[0])
        phys_id = int(fields[1]) - 1
        assert phys_id == i             # consecutive numbering, like the reference
        name = fields[2].strip(b'"').decode("utf-8")
        if ndim == mesh.ndim:
            region_id_map[phys_id] = mesh.new_region(name)
        elif ndim < mesh.ndim:
            boundary_id_map[phys_id] = mesh.new_boundary(name)
    if not f.readline().startswith(b"$EndPhysicalNames"):
        raise FileFormatError("Wrong number of physical names specifed")
    return region_id_map, boundary_id_map


def _parse_nodes_bin(f, mesh):
    if not f.readline().startswith(b"$Nodes"):
        raise FileFormatError("Expected 'Nodes' data")
    n_nodes = int(f.readline().rstrip())
    dt = np.dtype([("index", "<i4"), ("coord", "<f8", (3,))])
    raw = f.read(dt.itemsize * n_nodes)
    if len(raw) != dt.itemsize * n_nodes:
        raise FileFormatError("Expected end of 'Nodes' data")
    nodes_in = np.frombuffer(raw, dtype=dt)
    f.readline()
    if not f.readline().startswith(b"$EndNodes"):
        raise FileFormatError("Expected end of 'Nodes' data")
    assert np.all(nodes_in["index"] == np.arange(1, n_nodes + 1))
    mesh.set_nodes(np.ascontiguousarray(nodes_in["coord"][:, :mesh.ndim].T))


def _parse_elements_bin(f, mesh, region_id_map, boundary_id_map):
    """Cells go to the mesh block by block; lower-dimensional elements are
    returned as (vertex ids [B, 2**d], all node ids list, boundary ids [B])."""
    if not f.readline().startswith(b"$Elements"):
        raise FileFormatError("Expected 'Elements' data")
    n_elems = int(f.readline().rstrip())
    n_read = 0
    geo_ids = {}
    bnd_verts, bnd_ids, bnd_centroids = [], [], []
    while n_read < n_elems:
        header = np.frombuffer(f.read(12), dtype="<i4")
        if header.size != 3:
            raise FileFormatError("Expected 'Elements' data")
        elem_type, n_follow, n_tags = (int(v) for v in header)
        try:
            geometry = _make_geometry(elem_type)
        except KeyError:
            raise KeyError(elem_type)   # unsupported element type, like the reference
        is_cell = geometry.ndim == mesh.ndim
        if is_cell and elem_type not in geo_ids:
            geo_ids[elem_type] = mesh.add_geometry(geometry)
        n_nodes = geometry.n_nodes
        dt = np.dtype([("index", "<u4"), ("tags", "<u4", (n_tags,)), ("node_ix", "<u4", (n_nodes,))])
        raw = f.read(dt.itemsize * n_follow)
        if len(raw) != dt.itemsize * n_follow:
            raise FileFormatError("Expected 'Elements' data")
        data = np.frombuffer(raw, dtype=dt)
        assert np.all(data["index"] == np.arange(n_read + 1, n_read + n_follow + 1))
        # 1-based -> 0-based, Gmsh order -> lexicographic: one gather for the block
        lex = (data["node_ix"].astype(np.uint32) - np.uint32(1))[:, gmsh_to_lexicographic(geometry.shape)]
        phys = data["tags"][:, 0].astype(np.int64) - 1
        if is_cell:
            regions = np.array([region_id_map[int(q)] for q in phys], dtype=np.int32)
            mesh.add_cells(lex.reshape((-1,) + tuple(geometry.shape)), geo_ids[elem_type], regions)
        elif geometry.ndim < mesh.ndim:
            bnd_verts.append(lex[:, geometry.vertex_node_ind])
            bnd_ids.append(np.array([boundary_id_map[int(q)] for q in phys], dtype=np.int64))
            bnd_centroids.append(mesh.nodes[:, lex].mean(axis=2).T)
        n_read += n_follow
    f.readline()
    if not f.readline().startswith(b"$EndElements"):
        raise FileFormatError("Expected 'Elements' data")
    if bnd_verts:
        if len({v.shape[1] for v in bnd_verts}) != 1:
            raise NotImplementedError("boundary elements of mixed dimension")
        return np.concatenate(bnd_verts), np.concatenate(bnd_ids), np.concatenate(bnd_centroids)
    return (np.zeros((0, 2), dtype=np.uint32), np.zeros(0, dtype=np.int64),
            np.zeros((0, mesh.ndim)))


def find_cell_neighbors(mesh, bnd_verts, bnd_ids, bnd_centroids):
    """Fill the cell adjacency and register the boundary faces
    (replaces sem/grid_importers.py:221-270).  A face is identified by the
    sorted pair of its end-vertex ids."""
    if mesh.ndim != 2:
        raise NotImplementedError("neighbour search is implemented for 2-D meshes")
    blocks = mesh._blocks_flushed()
    E = mesh.n_cells
    n_faces = 4
    n = np.int64(mesh.n_nodes)
    keys = np.empty((E, n_faces), dtype=np.int64)
    centroids = np.empty((E, mesh.ndim))
    for blk, start in zip(blocks, mesh._block_start):
        g = mesh._geometries[blk.geometry_id]
        flat = blk.node_maps.reshape(blk.n_cells, -1)
        v = flat[:, g.vertex_node_ind].astype(np.int64)
        for side, mask in enumerate(g.corner_verts):
            a, b = v[:, np.flatnonzero(mask)].T
            keys[start:start + blk.n_cells, side] = np.minimum(a, b) * n + np.maximum(a, b)
        centroids[start:start + blk.n_cells] = mesh.nodes[:, flat].mean(axis=2).T
    # interior faces: the same key twice
    flat_keys = keys.ravel()
    order = np.argsort(flat_keys, kind="stable")
    sk = flat_keys[order]
    same = np.flatnonzero(sk[1:] == sk[:-1])
    adj = np.full((E, n_faces), -1, dtype=np.int64)
    f0, f1 = order[same], order[same + 1]
    adj.ravel()[f0] = f1 // n_faces
    adj.ravel()[f1] = f0 // n_faces
    mesh._adj_array = adj
    # boundary faces: a lower-dimensional element whose vertices are a face's vertices
    if bnd_verts.shape[0]:
        bv = bnd_verts.astype(np.int64)
        bkey = np.minimum(bv[:, 0], bv[:, 1]) * n + np.maximum(bv[:, 0], bv[:, 1])
        pos = np.searchsorted(sk, bkey)
        pos = np.minimum(pos, sk.size - 1)
        hit = sk[pos] == bkey
        cand = []                       # (cell, distance, boundary element, side)
        for b in np.flatnonzero(hit).tolist():
            p = pos[b]
            while p < sk.size and sk[p] == bkey[b]:     # (an interface line may touch 2 cells)
                face = int(order[p])
                cell, side = divmod(face, n_faces)
                d = float(np.linalg.norm(bnd_centroids[b] - centroids[cell]))
                cand.append((cell, d, b, side))
                p += 1
        # the reference visits cells in ascending order and, per cell, the boundary
        # elements nearest-centroid first
        cand.sort()
        for cell, _, b, side in cand:
            mesh.add_boundary_cell(cell, int(bnd_ids[b]), mesh.ndim - 1, side)


def load_msh(file_path, ndim):
    """Read a Gmsh 2.2 *binary* ``.msh`` file into a ``discrete.Mesh``
    (sem/grid_importers.py:45-69)."""
    with open(file_path, "rb") as f:
        is_binary, _ = parse_format(f)
        if not is_binary:
            raise NotImplementedError(
                "Reading ASCII *.msh files is not yet supported. Save the "
                "mesh in binary format and try again.")
        mesh = discrete.Mesh(ndim)
        region_id_map, boundary_id_map = _parse_physical_names(f, mesh)
        _parse_nodes_bin(f, mesh)
        bnd = _parse_elements_bin(f, mesh, region_id_map, boundary_id_map)
    find_cell_neighbors(mesh, *bnd)
    return mesh
