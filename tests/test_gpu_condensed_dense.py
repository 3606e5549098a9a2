"""The literal DOFManagerSC.solve inputs on the device (csrc/semk_sc.cu, dense mode):
the caller's own hierarchically ordered local systems are condensed, solved and
back-substituted by ``DOFManagerSC.solve_device`` (sem/discrete.py:478-528).
Anchors: the golden solutions of the live reference's ``solve`` and the host
drop-in (NumPy/SciPy mirror of the same reference methods) on a non-Poisson
system."""
import numpy as np
import pytest

import sem_oracle as so
from conftest import build_package_case, load_case, rel_l2
from spectralelementmethod_b200 import _lib

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _local_systems(mngr, g, shift=0.0):
    p = int(g["p"])
    nn = (p + 1) ** 2
    L = so.local_stiffness(so.Basis(p), g["invJ"], g["JxW"]).reshape(-1, nn, nn)
    jxw = g["JxW"].reshape(-1, nn)
    out = []
    for e, fe in enumerate(mngr.finite_elements()):
        lmat = L[e] + shift * np.diag(jxw[e])          # shift > 0: a Helmholtz-like operator
        out.append(mngr.reorder_local_system_hier(fe, (lmat, jxw[e] * (1.0 + shift))))
    return out


@pytest.mark.parametrize("name", ["S324_sc", "C448_sc_rcm", "C552_sc_rcm", "S888_sc_rcm"])
def test_solve_device_vs_reference_golden(name):
    g = load_case(name)
    mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"])
    systems = _local_systems(mngr, g)
    dof_vec = g["ebc_vals"].copy()
    info = mngr.solve_device(systems, dof_vec, g["on_ebc"][:mngr.ndof_exterior], rtol=1e-13)
    assert info.converged
    assert rel_l2(dof_vec, g["solution"]) < TOL


def test_solve_device_matches_host_drop_in_on_a_helmholtz_system():
    g = load_case("C448_sc_rcm")
    mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"])
    systems = _local_systems(mngr, g, shift=3.0)
    on = g["on_ebc"][:mngr.ndof_exterior]
    host_vec = g["ebc_vals"].copy()
    gsys = mngr.init_global_linear_system()
    mngr.assemble_global_sc_system(gsys, systems)
    mngr.solve(gsys, systems, host_vec, on)
    dev_vec = g["ebc_vals"].copy()
    mngr.solve_device(systems, dev_vec, on, rtol=1e-13)
    assert rel_l2(dev_vec, host_vec) < TOL
    assert rel_l2(dev_vec, g["solution"]) > 1e-3          # it is a different problem


def test_solve_device_errors():
    g = load_case("S324_sc")
    mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"])
    systems = _local_systems(mngr, g)
    on = g["on_ebc"][:mngr.ndof_exterior]
    skew = [(a + np.triu(np.ones_like(a), 1), b) for a, b in systems]
    with pytest.raises(NotImplementedError):              # not symmetric
        mngr.solve_device(skew, g["ebc_vals"].copy(), on)
    neg = [(-a, b) for a, b in systems]
    with pytest.raises(AssertionError):                   # interior block not positive definite
        mngr.solve_device(neg, g["ebc_vals"].copy(), on)
    with pytest.raises(ValueError):
        mngr.solve_device(systems, g["ebc_vals"].copy(), g["on_ebc"])   # mask of the wrong length
    lib = _lib.load()
    assert lib.semk_sc_element_dense_f64(9, 0, None, None, None, 1, None, 0, None, None, None,
                                         None, None) == _lib.ERR_INVALID
