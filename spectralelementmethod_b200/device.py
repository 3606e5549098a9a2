"""Thin torch <-> libsemk glue: device tensors in, C-ABI kernel launches out.

PyTorch is only the carrier of device memory and streams here; every number is
produced by the hand-written sm_100a kernels in csrc/.  All functions raise if
no CUDA device is present -- there is no CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib

__all__ = ["stream_ptr", "as_i32_bits", "basis_tables", "element_geometry", "ptr"]


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or host) address of a tensor / numpy array, or NULL for None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    return C.c_void_p(t.data_ptr())


def as_i32_bits(a, device="cuda"):
    """uint32 numpy array -> int32 torch tensor with the same bits (torch has
    no arithmetic on uint32; the kernels reinterpret)."""
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return torch.from_numpy(a.view(np.int32)).to(device)


def _f64(a, device="cuda"):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)


class BasisTables(object):
    """Device copies of the 1-D tables the kernels need for one tensor basis."""

    def __init__(self, basis):
        subs = [b for _, b in basis.iter_subbases()]
        if basis.ndim != 2 or len(subs) != 2:
            raise NotImplementedError("Only supporting 2D elements right now")
        b0, b1 = subs
        if b0.n_coeffs != b1.n_coeffs or not (np.array_equal(b0.D1, b1.D1)
                                              and np.array_equal(b0.nodes, b1.nodes)):
            raise NotImplementedError("the engine needs the same 1-D basis in both directions")
        self.n1 = int(b0.n_coeffs)
        if self.n1 < 2 or self.n1 > _lib.MAX_N1:
            raise NotImplementedError("Basis only available up to order %d." % (_lib.MAX_N1 - 1))
        self.D_host = np.ascontiguousarray(b0.D1, dtype=np.float64)
        self.Einv_host = np.ascontiguousarray(b0.interp_eq_inv, dtype=np.float64)
        self.w_host = np.ascontiguousarray(b0.quad_rule.weights, dtype=np.float64)
        self._dev = None

    def dev(self):
        if self._dev is None:
            self._dev = (_f64(self.D_host), _f64(self.Einv_host), _f64(self.w_host))
        return self._dev


def basis_tables(basis):
    tab = getattr(basis, "_semk_tables", None)
    if tab is None:
        tab = BasisTables(basis)
        basis._semk_tables = tab
    return tab


def geom_factors(tab, nodes_dev, l2g_dev, n_elem, elem_of_slot=None, G=None, g_patch_stride=0,
                 elems_per_patch=1, JxW=None, x_phys=None, J=None, invJ=None, detJ=None,
                 check=True):
    """Launch K1 (csrc/semk_geom.cu).  Raises AssertionError on a non-positive
    Jacobian, like the reference's ``assert np.all(det_jacobian > 0)``
    (sem/mapping.py:117)."""
    lib = _lib.load()
    D, Einv, w = tab.dev()
    bad = torch.zeros(1, dtype=torch.int32, device=nodes_dev.device)
    _lib.check(lib.semk_geom_factors_f64(
        tab.n1, int(n_elem), ptr(nodes_dev[0]), ptr(nodes_dev[1]), ptr(l2g_dev), ptr(Einv),
        ptr(D), ptr(w), ptr(elem_of_slot), ptr(G), int(g_patch_stride), int(elems_per_patch),
        ptr(JxW), ptr(x_phys), ptr(J), ptr(invJ), ptr(detJ), ptr(bad), stream_ptr()))
    if check and int(bad.item()) != 0:
        raise AssertionError("non-positive Jacobian determinant in the mesh")


def element_geometry(basis, nodes, l2g, jacobian=True):
    """Geometry of a batch of elements, returned as host arrays in the
    reference's layouts: x_phys [E,2,N,N]; J, invJ [E,2,2,N,N]; detJ [E,N,N]."""
    _lib.require_device()
    tab = basis_tables(basis)
    N = tab.n1
    l2g = np.ascontiguousarray(l2g, dtype=np.uint32).reshape(-1, N * N)
    E = l2g.shape[0]
    nodes_dev = _f64(nodes)
    if nodes_dev.shape[0] != 2:
        raise NotImplementedError("Only supporting 2D elements right now")
    l2g_dev = as_i32_bits(l2g)
    kw = dict(dtype=torch.float64, device="cuda")
    out = {"x_phys": torch.empty((E, 2, N, N), **kw)}
    if jacobian:
        out["J"] = torch.empty((E, 2, 2, N, N), **kw)
        out["invJ"] = torch.empty((E, 2, 2, N, N), **kw)
        out["detJ"] = torch.empty((E, N, N), **kw)
    geom_factors(tab, nodes_dev, l2g_dev, E, x_phys=out["x_phys"], J=out.get("J"),
                 invJ=out.get("invJ"), detJ=out.get("detJ"), check=jacobian)
    return {k: v.cpu().numpy() for k, v in out.items()}
