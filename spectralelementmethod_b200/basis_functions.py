"""Finite-element basis sets -- host mirror of the reference's
``sem.basis_functions``.

Same public names and call signatures as ``sem/basis_functions.py`` so that
code written against the reference (``LagrangeGaussLobatto(p)``,
``TensorProductQS(b, b)``, ``basis.get_D1_matrices()``, ``basis.gradient`` ...)
runs unchanged.  Everything here is small host-side table work; the per-element
tensor contractions of the hot path are done on the GPU (csrc/), which reads
the tables produced here (``D1``, ``interp_eq_inv``, quadrature weights).

T0 parity (bit-exact): nodes / barycentric weights / quadrature
weights come from ``gll_tables`` (bytes of sem/data/basis-data.hdf5 for order
<= 10); ``D1`` follows the reference's operation order
(sem/basis_functions.py:213-217): ratio of barycentric weights, divide by the
node differences, zero the diagonal, then the negative row sum.
"""
import itertools

import numpy as np
import scipy.linalg as _sla

from . import gll_tables, quadratures

__all__ = ["BarycentricLagrange", "LagrangeGaussLobatto", "TensorProduct",
           "NodalTensorProduct", "TensorProductQS"]


# --------------------------------------------------------------------------
# mix-ins (names follow sem/basis_functions.py:17-182)
# --------------------------------------------------------------------------
class _Basis(object):
    def get_coeff_rank(self, coeffs):
        return coeffs.ndim - self.ndim

    def interpolate(self, coeffs, x):
        if x.ndim != 1 or coeffs.shape[-1] != self.n_coeffs:
            raise AssertionError("bad shapes for interpolate")
        return np.einsum("mr,...r->...m", self(x), coeffs)


class _Nodal(object):
    @property
    def nodes(self):
        return self._nodes

    @property
    def n_nodes(self):
        return self._nodes.size

    @property
    def n_coeffs(self):
        return self._nodes.size


class _QuadSupported(object):
    """Nodal basis whose nodes carry quadrature weights."""

    def __init__(self, quad_wts):
        self._quad_rule = quadratures.Quadrature1D(self._nodes, quad_wts)

    @property
    def quad_rule(self):
        return self._quad_rule

    def integrate(self, coeffs):
        return self._quad_rule.integrate(coeffs)


class _Basis1D(_Basis):
    ndim = 1

    @property
    def coeff_shape(self):
        return (self.n_coeffs,)

    @property
    def D1(self):
        """First-derivative matrix: c' = D1 @ c."""
        return self._D1

    def get_D1_matrix(self, dim=0):
        # The reference reads a nonexistent attribute here
        # (sem/basis_functions.py:102-110); returning D1 is the evident intent.
        return self._D1

    def get_D1_matrices(self):
        return [self._D1]

    def deriv(self, coeffs):
        if coeffs.shape[-1] != self.n_coeffs:
            raise AssertionError("last axis must hold the coefficients")
        return np.einsum("mr,...r->...m", self._D1, coeffs)

    def gradient(self, coeffs):
        return self.deriv(coeffs)


class _BasisND(_Basis):
    @property
    def ndim(self):
        return self._ndim

    @property
    def D1(self):
        # (reference quirk kept: _BasisND.D1 reads self._D1,
        #  sem/basis_functions.py:138-144; TensorProduct only sets _D1_mats)
        return self._D1_mats

    def get_D1_matrix(self, dim):
        return self._D1_mats[dim]

    def get_D1_matrices(self):
        return list(self._D1_mats)


# --------------------------------------------------------------------------
# 1-D Lagrange bases
# --------------------------------------------------------------------------
class BarycentricLagrange(_Basis1D, _Nodal):
    """Lagrange basis on given nodes with given barycentric weights
    (reference: sem/basis_functions.py:185-341)."""

    def __init__(self, nodes, bary_wts):
        self._nodes = nodes
        self._bary_wts = bary_wts

        # D[i,j] = (w_j / w_i) / (x_i - x_j);  D[i,i] = -sum_{j != i} D[i,j]
        D = bary_wts[None, :] / bary_wts[:, None]
        with np.errstate(divide="ignore", invalid="ignore"):
            D /= nodes[:, None] - nodes[None, :]
        np.fill_diagonal(D, 0.0)
        np.fill_diagonal(D, -D.sum(axis=1))
        self._D1 = D

        # values of every basis function on the equispaced grid of the same
        # size (the mesh-node convention, sem/geometry.py:96-98) and its LU
        self._interp_eq_mat = self(np.linspace(-1, 1, self.n_nodes))
        self._interp_eq_mat_lu = _sla.lu_factor(self._interp_eq_mat)
        self._interp_eq_inv = None

    @property
    def deg(self):
        return self._nodes.size - 1

    @property
    def bary_wts(self):
        return self._bary_wts

    @property
    def interp_eq_mat(self):
        """E[i,j] = l_j(x_eq_i): GLL coefficients -> equispaced-node values."""
        return self._interp_eq_mat

    @property
    def interp_eq_inv(self):
        """E^{-1} (equispaced values -> GLL coefficients), Newton-refined in
        extended precision so it is the correctly rounded inverse of the
        float64 matrix ``interp_eq_mat`` (additive; feeds the CUDA geometry
        kernel, which applies E^{-1} explicitly instead of an LU solve)."""
        if self._interp_eq_inv is None:
            E = self._interp_eq_mat.astype(np.longdouble)
            X = np.linalg.inv(self._interp_eq_mat).astype(np.longdouble)
            eye = np.eye(E.shape[0], dtype=np.longdouble)
            for _ in range(4):
                X = X + X @ (eye - E @ X)
            self._interp_eq_inv = np.ascontiguousarray(X.astype(np.float64))
        return self._interp_eq_inv

    def __call__(self, x):
        """B[..., j] = l_j(x[...]) by the barycentric formula; a point that
        hits a node exactly yields the Kronecker row
        (sem/basis_functions.py:226-255)."""
        with np.errstate(divide="ignore", invalid="ignore"):
            kern = self._bary_wts / (x[..., None] - self._nodes)
            tot = kern.sum(axis=-1)
            tot.shape += (1,)
            out = kern / tot
        out[np.isnan(out)] = 1.0
        return out

    def interpolate(self, f, x, broadcast=False):
        """Evaluate the interpolant of nodal values ``f`` (last axis = nodes)
        at points ``x``; with ``broadcast`` the leading axes of ``f`` pair up
        with the axes of ``x`` (sem/basis_functions.py:260-341)."""
        x = np.asarray(x)
        with np.errstate(divide="ignore", invalid="ignore"):
            kern = self._bary_wts / (x[..., None] - self._nodes)
            ksum = kern.sum(axis=-1)[...]
            if broadcast:
                n_free = f.ndim - 1 - x.ndim
                free = list(range(n_free))
                out = np.einsum(kern, [Ellipsis, n_free],
                                f, [Ellipsis] + free + [n_free],
                                [Ellipsis] + free)[...]
            else:
                n_free = f.ndim - 1
                out = np.inner(kern, f)[...]
            ksum.shape += (1,) * n_free
            out /= ksum

        hit = np.nonzero(np.isinf(kern))
        pts, which = hit[:-1], hit[-1]
        if which.size > 0:
            if x.ndim == 0:
                out[...] = f[..., which[0]]
            elif broadcast:
                out[pts] = f[pts + (Ellipsis, which)]
            else:
                out[pts] = np.rollaxis(f[..., which], -1)
        return out[()]

    def __repr__(self):
        return "%s(deg=%d)" % (type(self).__name__, self.deg)


class LagrangeGaussLobatto(BarycentricLagrange, _QuadSupported):
    """Lagrange basis through the Gauss-Legendre-Lobatto points
    (reference: sem/basis_functions.py:344-393).

    The reference reads half-tables from HDF5 and supports order <= 10
    (NotImplementedError above).  We embed the same bytes and additionally
    ship a fixture for orders 11..16 -- pass
    ``allow_extended=False`` to get the reference's strict behaviour.
    """

    def __init__(self, order, allow_extended=True):
        if order < 1:
            raise ValueError("Must specify an order of 1 or greater.")
        max_order = gll_tables.MAX_ORDER if allow_extended else gll_tables.REFERENCE_MAX_ORDER
        if order > max_order:
            raise NotImplementedError(
                "Basis only available up to order {}.".format(max_order))
        half = np.array(gll_tables.half_table(int(order)), dtype=np.float64)

        n = order + 1
        m = n // 2
        full = np.zeros((3, n))
        full[:, m:] = half
        if n % 2 == 1:      # odd number of nodes: centre node is its own mirror
            mirror = half[:, -1:0:-1]
            sign = 1.0
        else:               # even: barycentric weights alternate through 0
            mirror = half[:, ::-1]
            sign = -1.0
        full[0, :m] = -mirror[0]
        full[1, :m] = sign * mirror[1]
        full[2, :m] = mirror[2]
        self._n_coeffs = n
        BarycentricLagrange.__init__(self, full[0].copy(), full[1].copy())
        _QuadSupported.__init__(self, full[2].copy())


# --------------------------------------------------------------------------
# tensor-product bases
# --------------------------------------------------------------------------
class TensorProduct(_BasisND):
    """Tensor product of lower-dimensional bases
    (reference: sem/basis_functions.py:396-659)."""

    def __init__(self, *subbases):
        if len(subbases) < 1:
            raise ValueError("Tensor product basis must comprise at "
                             "least two lower dimensional bases.")
        self._subbases = subbases
        self._ndim = sum(b.ndim for b in subbases)
        self._coeff_shape = tuple(itertools.chain.from_iterable(
            b.coeff_shape for b in subbases))
        self._subbasis_dims = []
        self._D1_mats = []
        self._n_coeffs = 1
        first = 0
        for b in subbases:
            self._n_coeffs *= b.n_coeffs
            if isinstance(b, _Basis1D):
                self._subbasis_dims.append(first)
                self._D1_mats.append(b.D1)
                first += 1
            else:
                self._subbasis_dims.append(slice(first, first + b.ndim))
                self._D1_mats.extend(b._D1_mats)
                first += b.ndim

    @property
    def coeff_shape(self):
        return self._coeff_shape

    @property
    def n_coeffs(self):
        return self._n_coeffs

    @property
    def n_subbases(self):
        return len(self._subbases)

    def get_subbasis(self, dim):
        """Basis of the face normal to ``dim`` (sem/basis_functions.py:450-472)."""
        if self.ndim == 2:
            return self._subbases[dim]
        rolled = self._subbases[dim + 1:] + self._subbases[:dim]
        return type(self)(*rolled)

    def iter_subbases(self, reverse=False):
        pairs = zip(self._subbasis_dims, self._subbases)
        return reversed(list(pairs)) if reverse else pairs

    def __call__(self, x):
        if len(x) != self.ndim:
            raise ValueError("Cannot evaluate {}-dimensional basis at "
                             "a {}-dimensional set of points"
                             .format(self.ndim, len(x)))
        args = []
        for i, (dim, b) in enumerate(self.iter_subbases()):
            args += [b(x[dim]), [Ellipsis, i]]
        args.append([Ellipsis] + list(range(self.n_subbases)))
        return np.einsum(*args)

    # -- interpolation -------------------------------------------------------
    def _check_coeffs(self, coeffs):
        if coeffs.shape[-self.n_subbases:] != self.coeff_shape:
            raise AssertionError("coefficient array has the wrong trailing shape")

    def interpolate(self, coeffs, x):
        self._check_coeffs(coeffs)
        out = coeffs
        for dim, b in self.iter_subbases(reverse=True):
            out = b.interpolate(out, x[dim], broadcast=(dim < self.ndim - 1))
        return out

    def interpolate_on_grid(self, coeffs, x):
        if len(x) != self.ndim:
            raise AssertionError("need one coordinate vector per dimension")
        self._check_coeffs(coeffs)
        out = coeffs
        for dim, b in self.iter_subbases(reverse=True):
            out = b.interpolate(out, x[dim])
        return out

    def _apply_along_axes(self, arr, op):
        """Apply ``op(dim, basis, matrix[n_dim, -1]) -> matrix`` along every
        tensor axis of the trailing ``ndim`` axes of ``arr``."""
        nd = self.n_subbases
        lead = arr.shape[:-nd]
        out = arr
        for dim, b in self.iter_subbases():
            ax = out.ndim - nd + dim
            moved = np.moveaxis(out, ax, 0)
            shp = moved.shape
            res = op(dim, b, np.ascontiguousarray(moved).reshape(shp[0], -1))
            out = np.moveaxis(res.reshape(shp), 0, ax)
        return np.ascontiguousarray(out).reshape(lead + self.coeff_shape)

    def interpolate_on_grid_eq(self, coeffs):
        """GLL coefficients -> values on the equispaced grid of the same
        shape (sem/basis_functions.py:539-569)."""
        self._check_coeffs(coeffs)
        return self._apply_along_axes(
            coeffs, lambda dim, b, m: np.dot(b._interp_eq_mat, m))

    def compute_coeffs_grid(self, values, x):
        """Coefficients from values on the tensor grid ``x``
        (sem/basis_functions.py:571-597)."""
        if len(x) != self.ndim or tuple(len(xd) for xd in x) != self.coeff_shape:
            raise AssertionError("grid shape must equal the coefficient shape")
        self._check_coeffs(values)
        return self._apply_along_axes(
            values, lambda dim, b, m: _sla.solve(b(np.asarray(x[dim])), m))

    def compute_coeffs_grid_eq(self, values):
        """Coefficients from values on the equispaced grid, by LU solves along
        each axis (sem/basis_functions.py:599-624)."""
        self._check_coeffs(values)
        return self._apply_along_axes(
            values, lambda dim, b, m: _sla.lu_solve(b._interp_eq_mat_lu, m))

    # -- differentiation -------------------------------------------------------
    def deriv(self, coeffs, dim):
        """d/d(xi_dim) of the coefficient array (sem/basis_functions.py:626-639)."""
        self._check_coeffs(coeffs)
        nd = self.n_subbases
        out_s = [Ellipsis] + list(range(nd))
        in_s = [nd if d == dim else d for d in out_s]
        return np.einsum(self._D1_mats[dim], [dim, nd], coeffs, in_s, out_s)

    def gradient(self, coeffs):
        """Stack of derivatives along every direction; result
        ``[ndim, ...rank, *coeff_shape]`` (sem/basis_functions.py:641-650)."""
        self._check_coeffs(coeffs)
        lead = coeffs.shape[:-self.n_subbases]
        g = np.empty((self.ndim,) + lead + self.coeff_shape)
        for i in range(self.ndim):
            g[i] = self.deriv(coeffs, i)
        return g

    def __repr__(self):
        return "%s(%s)" % (type(self).__name__,
                           ", ".join(repr(b) for b in self._subbases))

    def __str__(self):
        rows = ["[dim %d]: %s" % (i, b) for i, b in enumerate(self._subbases)]
        return "<%dD Basis> with basis functions:\n%s" % (self.ndim, "\n".join(rows))


class NodalTensorProduct(TensorProduct):
    def __init__(self, *subbases):
        self.check_subbases(subbases)
        TensorProduct.__init__(self, *subbases)

    @property
    def nodes(self):
        return tuple(b.nodes for b in self._subbases)

    def nodegrid(self, sparse=False):
        return np.meshgrid(*self.nodes, indexing="ij", sparse=sparse)

    def check_subbases(self, subbases):
        if not all(isinstance(b, _Nodal) for b in subbases):
            raise ValueError("All subbases must be nodal.")


class TensorProductQS(NodalTensorProduct, _QuadSupported):
    """Nodal tensor-product basis with the tensor quadrature rule on its
    nodes (reference: sem/basis_functions.py:683-697)."""

    def __init__(self, *subbases):
        if not all(isinstance(b, _QuadSupported) for b in subbases):
            raise ValueError("All subbases must be supported by a quadrature "
                             "rule.")
        TensorProduct.__init__(self, *subbases)
        self._quad_rule = quadratures.TensorQuadratureRule(
            *(b._quad_rule for b in self._subbases))
