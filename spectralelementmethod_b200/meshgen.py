"""Synthetic quad meshes, built vectorised (additive; no reference equivalent).

The reference only gets meshes from Gmsh files (sem/grid_importers.py) or by
hand, cell by cell (tests/test_discrete.py:22-38).  ``structured_quad_mesh``
produces exactly what that hand-built recipe yields -- same node numbering
(node id = i*NY + j on the (nx*p+1) x (ny*p+1) lattice, equispaced high-order
nodes per cell), same cell order (ex outer, ey inner), same boundary faces --
but with whole-array operations so that the 1024 x 1024, p = 8 configuration
(85 M map entries) takes seconds.

Boundaries follow examples/meshes/square.geo: "ebc" (essential) = left +
bottom, "nbc" (natural) = right + top, region "interior".
"""
import numpy as np

from .discrete import Mesh
from .geometry import Quadrilateral

__all__ = ["lattice_coordinates", "structured_node_maps", "structured_quad_mesh",
           "annulus_coordinates", "annulus_sector_mesh"]


def lattice_coordinates(kind, nx, ny, p, bounds=(-1.0, 1.0, -1.0, 1.0)):
    """``float64[2, NX*NY]`` node coordinates.  kind 'S': tensor lattice;
    'C': the same lattice displaced by s = 0.08 sin(pi X) sin(pi Y) in both
    coordinates (curved elements, boundary fixed)."""
    x0, x1, y0, y1 = bounds
    x = np.linspace(x0, x1, nx * p + 1)
    y = np.linspace(y0, y1, ny * p + 1)
    if kind not in ("S", "C"):
        raise ValueError("kind must be 'S' (straight) or 'C' (curved)")
    # written straight into the result (no meshgrid / vstack temporaries: 1 GB each at
    # 1024 x 1024 elements, p = 8); same values bit for bit
    out = np.empty((2, x.size * y.size))
    X = out[0].reshape(x.size, y.size)
    Y = out[1].reshape(x.size, y.size)
    X[:] = x[:, None]
    Y[:] = y[None, :]
    if kind == "C":
        # s = (0.08 sin(pi X)) sin(pi Y) is separable: the sines are taken on the 1-D grids
        s = (0.08 * np.sin(np.pi * x))[:, None] * np.sin(np.pi * y)[None, :]
        X += s
        Y += s
    return out


def structured_node_maps(nx, ny, p, node_offset=0):
    """``uint32[nx*ny, p+1, p+1]`` lexicographic node ids of every cell."""
    NY = ny * p + 1
    if nx * ny >= 4096:
        # large meshes: the multi-threaded host helper (csrc/semk_hostnum.cpp), same integers
        try:
            from . import _lib
            out = np.empty((nx * ny, p + 1, p + 1), dtype=np.uint32)
            if _lib.load().semk_host_structured_maps(nx, ny, p, int(node_offset), out.ctypes.data,
                                                     _lib.host_threads()) == 0:
                return out
        except (ImportError, OSError):
            pass
    ex, ey = np.divmod(np.arange(nx * ny, dtype=np.int64), ny)
    base = ex * p * NY + ey * p
    m = np.arange(p + 1, dtype=np.int64)
    maps = base[:, None, None] + m[None, :, None] * NY + m[None, None, :] + node_offset
    return maps.astype(np.uint32)


def structured_quad_mesh(nx, ny, p, kind="S", bounds=(-1.0, 1.0, -1.0, 1.0), nodes=None):
    """Mesh of nx x ny quadrilaterals of order p on a rectangle."""
    mesh = Mesh(2)
    mesh.set_nodes(lattice_coordinates(kind, nx, ny, p, bounds) if nodes is None else nodes)
    g = mesh.add_geometry(Quadrilateral(p + 1, p + 1))
    r = mesh.new_region("interior")
    ebc = mesh.new_boundary("ebc")
    nbc = mesh.new_boundary("nbc")
    mesh.add_cells(structured_node_maps(nx, ny, p), g, r)
    cell = np.arange(nx * ny).reshape(nx, ny)
    # insertion order per cell mirrors the hand-built recipe: face 0, 2, 1, 3
    for c in range(nx * ny) if nx * ny <= 4096 else ():
        ex, ey = divmod(c, ny)
        if ex == 0:
            mesh.add_boundary_cell(c, ebc, 1, 0)
        if ey == 0:
            mesh.add_boundary_cell(c, ebc, 1, 2)
        if ex == nx - 1:
            mesh.add_boundary_cell(c, nbc, 1, 1)
        if ey == ny - 1:
            mesh.add_boundary_cell(c, nbc, 1, 3)
    if nx * ny > 4096:
        # same per-cell lists, built from the four sides (only corner cells
        # carry two faces of one boundary: face 0 before 2, face 1 before 3)
        mesh.add_boundary_cells(cell[0, :], ebc, 1, 0)
        mesh.add_boundary_cells(cell[:, 0], ebc, 1, 2)
        mesh.add_boundary_cells(cell[-1, :], nbc, 1, 1)
        mesh.add_boundary_cells(cell[:, -1], nbc, 1, 3)
    mesh._structured_shape = (nx, ny)
    return mesh


def annulus_coordinates(nr, nt, p, r_out=100.0):
    """``float64[2, NR*NT]`` meridional coordinates (rho, z) of a spherical-annulus
    sector: r = r_out**s in [1, r_out] (geometric grading, as the ``Progression`` of
    examples/meshes/donut.geo), polar angle from pi (the -z axis) down to 0 (the +z axis)
    so that detJ = r > 0 with xi0 radial and xi1 along the angle; equispaced in the
    parametric coordinates of every cell; rho is exactly 0 on the axis of symmetry."""
    r = float(r_out) ** np.linspace(0.0, 1.0, nr * p + 1)
    th = np.linspace(np.pi, 0.0, nt * p + 1)
    sin = np.sin(th)
    sin[0] = 0.0
    sin[-1] = 0.0
    return np.vstack([np.outer(r, sin).ravel(), np.outer(r, np.cos(th)).ravel()])


def annulus_sector_mesh(nr, nt, p, r_out=100.0):
    """nr x nt quadrilaterals of order p between the unit sphere and the shell r = r_out
    in the meridional half plane: the synthetic stand-in for examples/meshes/donut.geo
    (physical names "sphere", "shell", "symaxis", region "interior"; BASELINE config 4)."""
    mesh = Mesh(2)
    mesh.set_nodes(annulus_coordinates(nr, nt, p, r_out))
    g = mesh.add_geometry(Quadrilateral(p + 1, p + 1))
    reg = mesh.new_region("interior")
    sphere = mesh.new_boundary("sphere")
    shell = mesh.new_boundary("shell")
    axis = mesh.new_boundary("symaxis")
    mesh.add_cells(structured_node_maps(nr, nt, p), g, reg)
    cell = np.arange(nr * nt).reshape(nr, nt)
    if nr * nt <= 4096:
        # per-cell registration order of a hand-built mesh: faces 0, 1, 2, 3
        for c in range(nr * nt):
            ex, ey = divmod(c, nt)
            if ex == 0:
                mesh.add_boundary_cell(c, sphere, 1, 0)
            if ex == nr - 1:
                mesh.add_boundary_cell(c, shell, 1, 1)
            if ey == 0:
                mesh.add_boundary_cell(c, axis, 1, 2)
            if ey == nt - 1:
                mesh.add_boundary_cell(c, axis, 1, 3)
    else:
        mesh.add_boundary_cells(cell[0, :], sphere, 1, 0)
        mesh.add_boundary_cells(cell[-1, :], shell, 1, 1)
        mesh.add_boundary_cells(cell[:, 0], axis, 1, 2)
        mesh.add_boundary_cells(cell[:, -1], axis, 1, 3)
    mesh._structured_shape = (nr, nt)
    return mesh


def write_gmsh22_binary(path, nx, ny, p, kind="S", bounds=(-1.0, 1.0, -1.0, 1.0), shuffle_seed=None,
                        elements_per_header=None):
    """Write the structured mesh of ``structured_quad_mesh`` as a Gmsh 2.2
    *binary* ``.msh`` file (the format sem/grid_importers.py:45-218 reads):
    physical names 1 "ebc" and 2 "nbc" (lines: left + bottom / right + top)
    and 3 "interior" (quadrilaterals); high-order line and quadrilateral
    elements in Gmsh's node ordering.  ``shuffle_seed`` permutes the element
    order inside each block and the node numbering (a mesh generator gives no
    ordering guarantees).  ``elements_per_header``: split every element block into
    headers of at most that many elements (1 = one header per element, the blocking
    Gmsh's own MSH2 binary writer is reported to use).  Test / example helper; p <= 10."""
    from .grid_importers import GMSH_LINE_TYPES, GMSH_QUAD_TYPES, gmsh_to_lexicographic
    n1 = p + 1
    line_type = {v: k for k, v in GMSH_LINE_TYPES.items()}[n1]
    quad_type = {v: k for k, v in GMSH_QUAD_TYPES.items()}[n1]
    xy = lattice_coordinates(kind, nx, ny, p, bounds)
    n_nodes = xy.shape[1]
    maps = structured_node_maps(nx, ny, p).reshape(nx * ny, n1, n1).astype(np.int64)
    cell = np.arange(nx * ny).reshape(nx, ny)
    # boundary lines as (node ids along the face in the face's lexicographic order, phys id)
    lines, phys = [], []
    for cells, take, tag in ((cell[0, :], lambda m: m[:, 0, :], 1), (cell[:, 0], lambda m: m[:, :, 0], 1),
                             (cell[-1, :], lambda m: m[:, -1, :], 2), (cell[:, -1], lambda m: m[:, :, -1], 2)):
        lines.append(take(maps[cells]))
        phys.append(np.full(len(cells), tag))
    lines, phys = np.concatenate(lines), np.concatenate(phys)
    perm = np.arange(n_nodes)
    if shuffle_seed is not None:
        rng = np.random.default_rng(shuffle_seed)
        perm = rng.permutation(n_nodes)                 # new id of old node k
        order = rng.permutation(len(lines))
        lines, phys = lines[order], phys[order]
        maps = maps[rng.permutation(len(maps))]
    inv_line = np.argsort(gmsh_to_lexicographic((n1,)))           # gmsh position -> lexicographic
    inv_quad = np.argsort(gmsh_to_lexicographic((n1, n1)))
    coords = np.zeros((n_nodes, 3))
    coords[perm, :2] = xy.T
    with open(path, "wb") as f:
        f.write(b"$MeshFormat\n2.2 1 8\n" + np.array([1], dtype="<i4").tobytes() + b"\n$EndMeshFormat\n")
        f.write(b'$PhysicalNames\n3\n1 1 "ebc"\n1 2 "nbc"\n2 3 "interior"\n$EndPhysicalNames\n')
        f.write(b"$Nodes\n%d\n" % n_nodes)
        rec = np.empty(n_nodes, dtype=[("index", "<i4"), ("coord", "<f8", (3,))])
        rec["index"] = np.arange(1, n_nodes + 1)
        rec["coord"] = coords
        f.write(rec.tobytes() + b"\n$EndNodes\n")
        f.write(b"$Elements\n%d\n" % (len(lines) + len(maps)))
        first = 1
        for etype, node_ix, tag in ((line_type, perm[lines][:, inv_line], phys),
                                    (quad_type, perm[maps.reshape(len(maps), -1)][:, inv_quad],
                                     np.full(len(maps), 3))):
            chunk = len(node_ix) if not elements_per_header else int(elements_per_header)
            for c0 in range(0, len(node_ix), max(chunk, 1)):
                part, ptag = node_ix[c0:c0 + chunk], tag[c0:c0 + chunk]
                f.write(np.array([etype, len(part), 2], dtype="<i4").tobytes())
                rec = np.empty(len(part), dtype=[("index", "<u4"), ("tags", "<u4", (2,)),
                                                 ("node_ix", "<u4", (part.shape[1],))])
                rec["index"] = np.arange(first, first + len(part))
                rec["tags"][:, 0] = ptag
                rec["tags"][:, 1] = ptag
                rec["node_ix"] = part + 1
                f.write(rec.tobytes())
                first += len(part)
        f.write(b"\n$EndElements\n")
    return path


def quad_mesh_from_vertices(vertices, quads, p, boundary=None):
    """Unstructured mesh of straight-sided quadrilaterals of order ``p``.

    vertices : float[V, 2]; quads : int[E, 4] vertex ids in the order
    (u0, u1) = (0,0), (0,1), (1,0), (1,1) (the lexicographic corner order of
    ``Quadrilateral``).  High-order nodes are the equispaced bilinear lattice
    of every cell; nodes on a shared edge get one id (vertices first, then edge
    interiors, then cell interiors).  ``boundary``: name of a boundary to which
    every unshared face is assigned (default: none).  Test / example helper for
    meshes with irregular vertices (3, 5, 6 ... cells around a point)."""
    vertices = np.asarray(vertices, dtype=np.float64)
    quads = np.asarray(quads, dtype=np.int64).reshape(-1, 4)
    n1 = p + 1
    V, E = len(vertices), len(quads)
    s = np.linspace(0.0, 1.0, n1)
    edge_ids = {}
    coords = [vertices]
    n_nodes = V
    maps = np.empty((E, n1, n1), dtype=np.int64)
    # (corner a, corner b, index expression of the edge from a to b)
    edges = [(0, 1, lambda k: (0, k)), (2, 3, lambda k: (p, k)),
             (0, 2, lambda k: (k, 0)), (1, 3, lambda k: (k, p))]
    face_count = {}
    for e, q in enumerate(quads):
        x00, x01, x10, x11 = vertices[q]
        maps[e, 0, 0], maps[e, 0, p], maps[e, p, 0], maps[e, p, p] = q
        for a, b, at in edges:
            va, vb = int(q[a]), int(q[b])
            key = (min(va, vb), max(va, vb))
            face_count[key] = face_count.get(key, 0) + 1
            if p > 1:
                if key not in edge_ids:
                    edge_ids[key] = n_nodes
                    lo, hi = vertices[key[0]], vertices[key[1]]
                    coords.append(lo + s[1:-1, None] * (hi - lo))
                    n_nodes += p - 1
                first = edge_ids[key]
                for k in range(1, p):
                    kk = k if va == key[0] else p - k      # position counted from the lower vertex
                    maps[e][at(k)] = first + kk - 1
        if p > 1:
            a, b = np.meshgrid(s[1:-1], s[1:-1], indexing="ij")
            xy = ((1 - a)[..., None] * ((1 - b)[..., None] * x00 + b[..., None] * x01)
                  + a[..., None] * ((1 - b)[..., None] * x10 + b[..., None] * x11))
            maps[e, 1:-1, 1:-1] = n_nodes + np.arange((p - 1) ** 2).reshape(p - 1, p - 1)
            coords.append(xy.reshape(-1, 2))
            n_nodes += (p - 1) ** 2
    mesh = Mesh(2)
    mesh.set_nodes(np.ascontiguousarray(np.concatenate(coords).T))
    g = mesh.add_geometry(Quadrilateral(n1, n1))
    r = mesh.new_region("interior")
    mesh.add_cells(maps.astype(np.uint32), g, r)
    if boundary is not None:
        bid = mesh.new_boundary(boundary)
        for e, q in enumerate(quads):
            for face, (a, b) in enumerate(((0, 1), (2, 3), (0, 2), (1, 3))):
                key = (min(int(q[a]), int(q[b])), max(int(q[a]), int(q[b])))
                if face_count[key] == 1:
                    mesh.add_boundary_cell(e, bid, 1, face)
    return mesh


def pinwheel_mesh(n_cells, p, boundary="ebc", rings=1):
    """``n_cells`` quadrilaterals around one interior vertex (valence
    ``n_cells``: 3, 5, 6 ... are irregular), optionally surrounded by further
    rings of cells.  Cell k of ring 0 spans the centre, ray k, ray k+1 and an
    outer corner between them."""
    ang = 2.0 * np.pi * np.arange(n_cells) / n_cells
    half = np.pi / n_cells
    verts = [np.zeros(2)]
    ray = lambda rad: [rad * np.array([np.cos(a), np.sin(a)]) for a in ang]          # noqa: E731
    mid = lambda rad: [rad * np.array([np.cos(a + half), np.sin(a + half)]) for a in ang]   # noqa: E731
    quads = []
    verts += ray(1.0) + mid(1.5)
    A = lambda k: 1 + k % n_cells                       # noqa: E731
    B = lambda k: 1 + n_cells + k % n_cells             # noqa: E731
    for k in range(n_cells):
        quads.append((0, A(k + 1), A(k), B(k)))         # (0,0) centre, (0,1) ray k+1, (1,0) ray k
    base_a, base_b = A, B
    for ring in range(1, rings):
        off = len(verts)
        verts += ray(1.0 + ring) + mid(1.5 + ring)
        A2 = lambda k, o=off: o + k % n_cells           # noqa: E731
        B2 = lambda k, o=off: o + n_cells + k % n_cells  # noqa: E731
        for k in range(n_cells):
            # two cells per sector: over the ray and over the corner
            quads.append((base_a(k), base_b(k), A2(k), B2(k)))
            quads.append((base_b(k), base_a(k + 1), B2(k), A2(k + 1)))
        base_a, base_b = A2, B2
    return quad_mesh_from_vertices(np.array(verts), np.array(quads), p, boundary)
