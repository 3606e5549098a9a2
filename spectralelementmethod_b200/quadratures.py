"""Quadrature rules -- host-side mirror of the reference's ``sem.quadratures``.

Drop-in for ``sem/quadratures.py`` (class and method names, argument meaning,
return layouts).  These are tiny host tables (N <= 17 doubles); the hot path
never evaluates them per element -- the CUDA geometry kernel consumes the 1-D
weight vectors directly (csrc/semk_geom.cuh).

Bit-exactness contract (tier T0): ``GaussLobatto(n)``
must reproduce ``sem.quadratures.GaussLobatto`` (sem/quadratures.py:148-193)
bit for bit, which pins the sequence of floating-point operations (companion
roots, one Newton sweep, symmetrisation, rescale) but not how they are spelled.
"""
import numpy as np
from numpy.polynomial import legendre as _leg

__all__ = ["Quadrature1D", "GaussLobatto", "TensorQuadratureRule"]


class Quadrature1D(object):
    """An n-point rule on [-1, 1]  (reference: sem/quadratures.py:14-118)."""

    ndim = 1

    def __init__(self, abscissa, weights):
        self._abscissa = abscissa
        self._weights = weights

    # -- accessors ---------------------------------------------------------
    @property
    def n_points(self):
        return len(self._abscissa)

    @property
    def abscissa(self):
        return self._abscissa

    @property
    def weights(self):
        return self._weights

    def get_abscissa(self):
        return self._abscissa

    def get_weights(self):
        return self._weights

    # -- operations --------------------------------------------------------
    def __call__(self, f):
        """Integrate values-at-abscissa, or a callable evaluated there
        (sem/quadratures.py:59-83: values first, callable on TypeError)."""
        w = self._weights
        try:
            return np.dot(w, f)
        except TypeError:
            return np.dot(w, f(self._abscissa))

    def integrate(self, values):
        """Weighted sum over the leading axis (sem/quadratures.py:98-109)."""
        w = self._weights
        if values.shape[0] != w.size:
            raise AssertionError("leading axis must match the number of points")
        tail = values.shape[1:]
        return np.dot(w, values.reshape(w.size, -1)).reshape(tail)

    def xweight(self, f_vals):
        """Values times weights, not summed (sem/quadratures.py:111-115)."""
        return f_vals * self._weights

    def __repr__(self):
        return "%s(n=%d)" % (type(self).__name__, self.n_points)


class GaussLobatto(Quadrature1D):
    """n-point Gauss-Legendre-Lobatto rule (sem/quadratures.py:121-200).

    Interior abscissa are the roots of P'_{n-1}; w_i ~ 1 / P_{n-1}(x_i)^2,
    rescaled to sum to 2.  Exact for polynomials of degree <= 2n-3.
    """

    def __init__(self, n):
        if int(n) != n or int(n) < 1:
            raise ValueError("n must be a positive integer")
        n = int(n)
        Pn = _leg.Legendre.basis(n - 1)
        dPn = Pn.deriv()
        d2Pn = dPn.deriv()

        x = np.zeros(n)
        x[0], x[-1] = -1.0, 1.0
        inner = dPn.roots()                      # companion-matrix eigenvalues
        inner = inner - dPn(inner) / d2Pn(inner)  # one Newton polish
        x[1:-1] = inner

        w = np.ones(n)
        w[1:-1] /= Pn(x[1:-1]) ** 2

        # enforce the +-symmetry exactly, then normalise
        x[1:-1] = (x[1:-1] - x[-2:0:-1]) / 2.0
        w[1:-1] = (w[1:-1] + w[-2:0:-1]) / 2.0
        w *= 2.0 / w.sum()
        Quadrature1D.__init__(self, x, w)

    @property
    def deg(self):
        """Highest polynomial degree integrated exactly."""
        return 2 * len(self._abscissa) - 3


class TensorQuadratureRule(object):
    """Tensor product of 1-D rules (sem/quadratures.py:203-275)."""

    def __init__(self, *quad_rules):
        self._abscissa = [r.abscissa for r in quad_rules]
        self._weights = [r.weights for r in quad_rules]
        self._ndim = sum(r.ndim for r in quad_rules)
        self._n_points = 1
        for r in quad_rules:
            self._n_points *= r.abscissa.size

    @property
    def ndim(self):
        return self._ndim

    @property
    def n_points(self):
        return self._n_points

    @property
    def n_subquads(self):
        return len(self._weights)

    @property
    def shape(self):
        return tuple(len(a) for a in self._abscissa)

    @property
    def abscissa(self):
        return list(self._abscissa)

    @property
    def weights(self):
        return list(self._weights)

    def get_abscissa(self, sparse=False):
        return np.meshgrid(*self._abscissa, indexing="ij", sparse=sparse)

    def get_weights(self, sparse=False):
        """Sparse: list of broadcastable 1-D weight grids.  Dense: the
        reference returns ``np.prod`` of the dense grids, i.e. a scalar
        (sem/quadratures.py:246-252, a latent bug we keep for drop-in parity)."""
        grids = np.meshgrid(*self._weights, indexing="ij", sparse=sparse)
        return grids if sparse else np.prod(grids)

    def integrate(self, f_vals):
        out = f_vals
        for w in reversed(self._weights):
            out = np.inner(out, w)
        return out

    def __call__(self, f):
        try:
            return self.integrate(f)
        except TypeError:
            return self.integrate(f(self._abscissa))

    def xweight(self, f_vals):
        """f * w0[:,None] * w1[None,:] ... applied axis 0 first, in place on a
        copy (sem/quadratures.py:268-275; the order fixes the rounding)."""
        out = f_vals.copy()
        for w in self.get_weights(sparse=True):
            out *= w
        return out
