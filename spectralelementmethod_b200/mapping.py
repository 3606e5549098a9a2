"""Element mappings -- host mirror of the reference's ``sem.mapping``.

``Mapping(basis, cell, compute_flags)`` keeps the reference's constructor and
properties (``x_phys``, ``J``, ``invJ``, ``detJ``; sem/mapping.py:79-178) but
the numbers come from the CUDA geometry kernel (csrc/semk_geom.cu): either
handed in by the DOF manager, which evaluates whole batches of elements at
once, or computed on the device for this single cell.  There is no NumPy
fallback for the geometry; without a GPU the properties raise.

``_subface_slice`` is pure index manipulation (views), tier T0.
"""
import numpy as np

from . import rootfind

__all__ = ["OutsideDomain", "Mapping", "SubMapping", "_subface_slice"]


class OutsideDomain(Exception):
    """A physical point lies outside the parametric domain of an element."""


def _subface_slice(face, arr, ndim):
    """View of the values of ``arr`` (trailing ``ndim`` axes = element axes)
    on face number ``face`` (reference: sem/mapping.py:19-76).

    Faces 2a / 2a+1 are the low / high end of axis ``a``.  In 2-D the edges are
    returned oriented counter-clockwise around the element (faces 0 and 3 are
    reversed); in higher dimensions a low face has its remaining axes reversed
    so that its orientation is outward.
    """
    if ndim <= 1:
        raise AssertionError("sub-faces need a parent of dimension > 1")
    if face >= 2 * ndim:
        raise AssertionError("face number out of range")
    rank = arr.ndim - ndim
    axis = rank + face // 2
    high = bool(face % 2)
    # element axes rolled so that the face-normal axis comes first
    order = list(range(rank)) + list(range(axis, arr.ndim)) + list(range(rank, axis))
    rolled = arr.transpose(order)
    lead = (slice(None),) * rank
    if ndim == 2:
        flip = slice(None, None, -1) if face in (0, 3) else slice(None)
        return rolled[lead + ((-1 if high else 0), flip)]
    if high:
        return rolled[lead + (-1,)]
    low = rolled[lead + (0,)]
    return low.transpose(list(range(rank)) + list(range(arr.ndim - 2, rank - 1, -1)))


class Mapping(object):
    """Parametric <-> physical map of one cell (isoparametric)."""

    def __init__(self, basis, cell, compute_flags, _precomputed=None):
        self._basis = basis
        self._cell = cell
        self._cmpflags = compute_flags
        need_x = compute_flags.get("x_phys", False)
        need_j = compute_flags.get("Jacobian", False)
        if need_j and basis.ndim != 2:
            raise NotImplementedError("Only supporting 2D elements right now")
        if need_x or need_j:
            geo = _precomputed if _precomputed is not None else self._device_geometry(need_j)
            self._x_phys = geo["x_phys"]
            if need_j:
                self._J, self._invJ, self._detJ = geo["J"], geo["invJ"], geo["detJ"]

    def _device_geometry(self, jacobian):
        from . import device
        l2g = np.ascontiguousarray(self._cell.node_ind_lexicographic, dtype=np.uint32)[None]
        out = device.element_geometry(self._basis, self._cell._mesh.nodes, l2g, jacobian=jacobian)
        return {k: v[0] for k, v in out.items()}

    @property
    def ndim(self):
        return self._basis.ndim

    @property
    def x_phys(self):
        return self._x_phys

    @property
    def J(self):
        return self._J

    @property
    def invJ(self):
        return self._invJ

    @property
    def detJ(self):
        return self._detJ

    def __call__(self, x_param):
        """Physical coordinates of parametric point(s) (sem/mapping.py:141-144)."""
        return self._basis.interpolate(self.x_phys, x_param).swapaxes(-1, 0)

    def inv(self, x_phys, x_param_guess=None):
        """Parametric coordinates of a physical point by Newton iteration
        (sem/mapping.py:146-178); raises OutsideDomain if it lands outside
        [-1, 1]^ndim."""
        target = np.array(x_phys).reshape(self.ndim)
        guess = np.zeros_like(target) if x_param_guess is None else x_param_guess
        xi = rootfind.newton(lambda s: self(s) - target, guess,
                             lambda s: self._basis.interpolate(self.J, s),
                             it_max=8, tol=1e-8)
        if (xi >= -1.0).all() and (xi <= 1.0).all():
            return xi
        raise OutsideDomain("Given physical point is not in the parametric "
                            "domain of the finite element.")

    def get_submapping(self, face):
        return SubMapping(self, face)


class SubMapping(Mapping):
    """Mapping restricted to one face of a parent mapping
    (sem/mapping.py:184-272)."""

    def __init__(self, parent_mapping, face):
        self._face = face
        self._parent_mapping = parent_mapping
        self._basis = parent_mapping._basis.get_subbasis(face // 2)
        self._cell = parent_mapping._cell.sub_cell(face)
        self._cmpflags = dict(parent_mapping._cmpflags)
        if parent_mapping._cmpflags.get("Jacobian", False):
            self._normal_vec = self._compute_normal_vec()
            self._cmpflags["normal"] = True

    def _from_parent(self, name):
        pm = self._parent_mapping
        return _subface_slice(self._face, getattr(pm, name), pm.ndim)

    @property
    def x_phys(self):
        return self._from_parent("x_phys")

    @property
    def J(self):
        return self._from_parent("J")

    @property
    def invJ(self):
        return self._from_parent("invJ")

    @property
    def detJ(self):
        return self._from_parent("detJ")

    def _tangents(self):
        """Un-normalised tangent vector(s) of the face, counter-clockwise in
        2-D (sem/mapping.py:228-254)."""
        pm, face = self._parent_mapping, self._face
        Jf = self._from_parent("J")
        a, high, nd = face // 2, bool(face % 2), pm.ndim
        if high:
            cols = list(range(a + 1, nd)) + list(range(0, a))
        else:
            cols = list(range(a - 1, -1, -1)) + list(range(nd - 1, a, -1))
        if self.ndim == 1:
            tan = Jf[:, cols[0]].copy()
            if face in (0, 3):
                tan *= -1
            return tan
        return Jf[:, cols]

    def _compute_normal_vec(self):
        tan = self._tangents()
        if self.ndim != 1:
            # the reference discards np.cross for 2-D faces and then fails
            # (sem/mapping.py:204-205); only edges of 2-D cells are supported.
            raise NotImplementedError("only 1D sub-elements are supported.")
        nvec = np.roll(tan, 1, axis=0)   # (t_x, t_y) -> (t_y, -t_x): outward normal * dS
        nvec[1] *= -1
        return nvec

    @property
    def n_dS(self):
        return self._normal_vec

    @property
    def dS(self):
        return np.linalg.norm(self._normal_vec, axis=0)

    @property
    def unit_normal(self):
        return self._normal_vec / self.dS

    def inv(self, x_phys):
        raise NotImplementedError("Cannot compute the parametric coordinates"
                                  "of a SubMapping from physical coordinates.")
