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

__all__ = ["lattice_coordinates", "structured_node_maps", "structured_quad_mesh"]


def lattice_coordinates(kind, nx, ny, p, bounds=(-1.0, 1.0, -1.0, 1.0)):
    """``float64[2, NX*NY]`` node coordinates.  kind 'S': tensor lattice;
    'C': the same lattice displaced by s = 0.08 sin(pi X) sin(pi Y) in both
    coordinates (curved elements, boundary fixed)."""
    x0, x1, y0, y1 = bounds
    X, Y = np.meshgrid(np.linspace(x0, x1, nx * p + 1), np.linspace(y0, y1, ny * p + 1),
                       indexing="ij")
    if kind == "C":
        s = 0.08 * np.sin(np.pi * X) * np.sin(np.pi * Y)
        X = X + s
        Y = Y + s
    elif kind != "S":
        raise ValueError("kind must be 'S' (straight) or 'C' (curved)")
    return np.vstack([X.ravel(), Y.ravel()])


def structured_node_maps(nx, ny, p, node_offset=0):
    """``uint32[nx*ny, p+1, p+1]`` lexicographic node ids of every cell."""
    NY = ny * p + 1
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
