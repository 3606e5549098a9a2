import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
ORACLE_DIR = os.path.join(ROOT, "oracle")
if ORACLE_DIR not in sys.path:
    sys.path.insert(0, ORACLE_DIR)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # The shared libraries are build artefacts (git-ignored): build them once if absent.
    lib = os.path.join(ROOT, "spectralelementmethod_b200", "csrc", "libsemk.so")
    olib = os.path.join(ROOT, "oracle", "_build", "libsem_oracle_c.so")
    if not (os.path.exists(lib) and os.path.exists(olib)):
        import __graft_entry__
        __graft_entry__.build()


def golden_case_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "case_*.npz")))


def load_case(name):
    d = dict(np.load(os.path.join(GOLDEN, "case_%s.npz" % name)))
    nx, ny, p, sc, rcm, kind = d.pop("meta").tolist()
    d.update(nx=nx, ny=ny, p=p, sc=bool(sc), rcm=bool(rcm), kind=chr(kind))
    return d


@pytest.fixture(scope="session")
def golden_tables():
    return dict(np.load(os.path.join(GOLDEN, "tables.npz")))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def build_package_case(kind, nx, ny, p, sc, rcm):
    """Mesh + DOF manager of the product package for a synthetic case."""
    from spectralelementmethod_b200 import discrete, meshgen
    from spectralelementmethod_b200.basis_functions import LagrangeGaussLobatto, TensorProductQS
    mesh = meshgen.structured_quad_mesh(nx, ny, p, kind)
    b1 = LagrangeGaussLobatto(p)
    basis = TensorProductQS(b1, b1)
    cls = discrete.DOFManagerSC if sc else discrete.DOFManager
    return mesh, cls(mesh, 1, basis, rcm_order=rcm)
