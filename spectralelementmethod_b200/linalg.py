"""Small dense helpers -- host mirror of the hot part of ``sem.linalg``.

Only ``det_inv_2x2`` (sem/linalg.py:105-115) is on the path; the legacy
``sp_schur_solve`` (sem/linalg.py:9-102) depends on element attributes that no
longer exist in the reference and is out of scope.  On the
device the same closed form lives in csrc/semk_geom.cuh.
"""
import numpy as np

__all__ = ["det_inv_2x2"]


def det_inv_2x2(mat):
    """Determinant and inverse of a stack of 2x2 matrices ``mat[2, 2, ...]``.

    The inverse is the adjugate scaled by the reciprocal of the determinant
    (reciprocal first, then multiply -- the order matters in the last bit and
    matches sem/linalg.py:113-114).
    """
    a, b, c, d = mat[0, 0], mat[0, 1], mat[1, 0], mat[1, 1]
    det = a * d - b * c
    inv = np.stack([np.stack([d, -b]), np.stack([-c, a])]).astype(mat.dtype, copy=False)
    inv *= 1 / det
    return det, inv
