"""Reference-cell topology -- host mirror of the reference's ``sem.geometry``.

An ``NCube(*shape)`` describes a tensor-product cell with ``shape[d]`` nodes
along axis ``d`` (equispaced in parametric space, the Gmsh convention,
sem/geometry.py:96-98).  What the hot path needs from it is the *hierarchical
node order* (vertices, then the open edges/faces by increasing dimension, then
the interior; sem/geometry.py:151-212) that defines the exterior-first
numbering of ``DOFManagerSC``; it is an integer table and is reproduced
bit-exactly (tier T0).  Face numbering of a quadrilateral: 0: xi0=-1,
1: xi0=+1, 2: xi1=-1, 3: xi1=+1 (sem/geometry.py:245-255).
"""
import itertools
from math import comb as _comb

import numpy as np

__all__ = ["Geometry", "NCube", "Line", "Quadrilateral"]


class Geometry(object):
    pass


class NCube(Geometry):
    def __init__(self, *shape):
        if not all(isinstance(s, (int, np.integer)) and s > 0 for s in shape):
            raise AssertionError("shape entries must be positive integers")
        self._shape = tuple(int(s) for s in shape)
        self._n_nodes = int(np.prod(self._shape))
        inner = 1
        for s in self._shape:
            inner *= s - 2
        if inner < 0:
            raise AssertionError("cell too small to have an interior count")
        self._n_interior_nodes = inner
        self._n_exterior_nodes = self._n_nodes - inner
        self._node_locations = np.meshgrid(
            *(np.linspace(-1, 1, s) for s in self._shape), indexing="ij", sparse=True)
        self._hier_node_order = self._compute_hierarchical_node_ordering()
        self._sub_geo_data = [self.sub_geometry_ix_exps(d) for d in range(self.ndim + 1)]
        self._sub_geo_class = NCube

    # -- sizes -----------------------------------------------------------------
    @property
    def ndim(self):
        return len(self._shape)

    @property
    def shape(self):
        return self._shape

    @property
    def n_nodes(self):
        return self._n_nodes

    @property
    def n_exterior_nodes(self):
        return self._n_exterior_nodes

    @property
    def n_interior_nodes(self):
        return self._n_interior_nodes

    @property
    def nodes(self):
        return self._node_locations

    # -- local index tables ----------------------------------------------------
    @property
    def hierarchical_node_order(self):
        return self._hier_node_order

    @property
    def vertex_node_ind(self):
        return self._hier_node_order[:2 ** self.ndim]

    @property
    def exterior_node_ind(self):
        return self._hier_node_order[:self._n_exterior_nodes]

    @property
    def interior_node_ind(self):
        return self._hier_node_order[self._n_exterior_nodes:]

    def n_sub_geometries(self, dim=-1):
        """Number of ``dim``-dimensional faces of the cell: 2^(n-dim) C(n,dim)
        (sem/geometry.py:130-149)."""
        n = self.ndim
        if dim < 0:
            dim = n + dim
        if dim > n:
            raise ValueError("No {}D sub-geometry in a {}D parent geometry".format(dim, n))
        if dim < 0:
            raise ValueError("Dimension of sub-elements must be > 0")
        return 2 ** (n - dim) * _comb(n, dim)

    def sub_geometry_ix_exps(self, dim=None, inclusive=True):
        """Index expressions selecting the nodes of every ``dim``-dimensional
        face: list of (shape, index-tuple).  Faces are enumerated by the
        combination of fixed axes (ascending), then by low/high end of each
        fixed axis (sem/geometry.py:151-195)."""
        n = self.ndim
        if dim is None:
            dim = n - 1
        if dim > n:
            raise ValueError("No {}D sub-geometry on a {}D parent geometry".format(dim, n))
        if dim < 0:
            raise ValueError("Dimension of sub-elements must be > 0")
        lo, trim = (0, 0) if inclusive else (1, 1)
        out = []
        for fixed in itertools.combinations(range(n), n - dim):
            ends = [(0, self._shape[a] - 1) for a in fixed]
            for where in itertools.product(*ends):
                pos = dict(zip(fixed, where))
                idx, shp = [], []
                for d in range(n):
                    if d in pos:
                        idx.append(pos[d])
                    else:
                        idx.append(slice(lo, self._shape[d] - trim))
                        shp.append(self._shape[d] - 2 * trim)
                out.append((tuple(shp), tuple(idx)))
        return out

    def _compute_hierarchical_node_ordering(self):
        lin = np.arange(self._n_nodes).reshape(self._shape)
        pieces = []
        for d in range(self.ndim + 1):
            for _, ix in self.sub_geometry_ix_exps(d, False):
                pieces.append(np.asarray(lin[ix]).ravel())
        order = np.concatenate(pieces).astype(np.uint32)
        if order.size != self._n_nodes:
            raise AssertionError("hierarchical order does not cover the cell")
        return order

    def sub_geometry(self, axis):
        shape = self._shape[axis + 1:] + self._shape[:axis]
        return self._sub_geo_class(*shape)


class Line(NCube):
    # +-->u0   (0)--*--(1)
    corner_verts = [np.array([True, False]), np.array([False, True])]

    def __init__(self, shape_u):
        NCube.__init__(self, shape_u)
        self._sub_geo_class = None

    @property
    def ndim(self):
        return 1

    def sub_geometry(self):
        raise NotImplementedError("The sub-geometry of a line is a single "
                                  "point, which is all not useful.")


class Quadrilateral(NCube):
    #        1--(3)--3
    #        |       |
    # u1    (0)  *  (1)
    # |      |       |
    # +--u0  0--(2)--2
    corner_verts = [np.array([1, 1, 0, 0], dtype=bool),
                    np.array([0, 0, 1, 1], dtype=bool),
                    np.array([1, 0, 1, 0], dtype=bool),
                    np.array([0, 1, 0, 1], dtype=bool)]

    def __init__(self, shape_u, shape_v):
        NCube.__init__(self, shape_u, shape_v)
        self._sub_geo_class = Line

    @property
    def ndim(self):
        return 2
