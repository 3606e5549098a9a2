"""Run the reference's OWN squirmer example class (examples/squirmer-axisymmetric.py) live.

TEST INFRASTRUCTURE ONLY (development container: needs /root/reference/examples, which
does not travel).  oracle/make_golden_stokes.py uses it to freeze golden vectors for the
axisymmetric Stokes / Navier-Stokes row.  No product module imports
this file.

The example is older than the sem package next to it.  On top of the five shims of
oracle/live_reference.py it needs these NAME aliases (no source edits; the arithmetic that
runs is the example's own `pre_assembly`, `compute_local_system`, `compute_loc_sc_sys`,
`assemble_global_sc_sys`, `_solve_boundary_unks`, `_solve_interior_unks`, `solve`):
  a. basis_funcs.LagrangeAtGaussLobatto  -> sem.basis_functions.LagrangeGaussLobatto
     (the example hard-codes order 8, examples/squirmer-axisymmetric.py:91; the alias takes
     the order from ORDER below so that small golden cases are possible)
  b. basis_funcs.TensorProductSupported  -> sem.basis_functions.TensorProductQS
  c. basis.get_diff_matrices()           -> get_D1_matrices()      (:185)
  d. bnd_fe.normal()                     -> -SubFiniteElement.n_dS (:135, "non-normalized
     unit vector": the normal times the arc-length factor, pointing from the body INTO the
     fluid.  The method no longer exists; the orientation is fixed by the example's own
     numbers: with this sign the force-free speed of a Stokes squirmer comes out as +1 (the
     secant guesses of calc_speed are 0.99 / 1.01, :630-744) and Re = 1, beta = 1 gives the
     0.9257 quoted in its docstring (:667); with +n_dS both come out mirrored (-1, and the
     pusher's 1.088).  Only enters the natural-BC contour integral `cint`.)
  e. python 2 names: itertools.izip -> zip, xrange -> range       (:345,:366,:427-431)
  f. np.ogrid[[slice, ...]] list index -> tuple index (numpy 2)   (:201,:216,:224 and
     sem/sp_array.py:107)
  g. Static_COO_Matrix.tocsr()           -> tocoo().tocsr()        (:368; the class in
     sem/discrete.py:26-41 only has tocoo, which is what DOFManagerSC.solve itself uses)
"""
import itertools
import os
import types

import numpy as np

from . import live_reference as live

EXAMPLE = os.path.join(live.REF_ROOT, "examples", "squirmer-axisymmetric.py")
ORDER = 8


def available():
    return live.available() and os.path.isfile(EXAMPLE)


class _OGrid(object):
    """np.ogrid that also accepts a list of slices (numpy 1 behaviour)."""

    def __getitem__(self, key):
        if isinstance(key, list):
            key = tuple(key)
        return np.ogrid[key]


class _NumpyProxy(types.ModuleType):
    def __init__(self):
        types.ModuleType.__init__(self, "numpy_proxy")
        self.ogrid = _OGrid()

    def __getattr__(self, name):
        return getattr(np, name)


def load_example():
    """exec the example's source into a fresh module and return it."""
    live.install_shims()
    import sem.basis_functions as bf
    import sem.discrete as sd
    import sem.sp_array as spa

    if not hasattr(bf, "LagrangeAtGaussLobatto"):
        bf.LagrangeAtGaussLobatto = lambda order: bf.LagrangeGaussLobatto(ORDER)
        bf.TensorProductSupported = bf.TensorProductQS
        bf.TensorProduct.get_diff_matrices = bf.TensorProduct.get_D1_matrices
        sd.SubFiniteElement.normal = lambda self: -self.n_dS
        itertools.izip = zip
        sd.Static_COO_Matrix.tocsr = lambda self: self.tocoo().tocsr()
        spa.np = _NumpyProxy()
    mod = types.ModuleType("squirmer_example")
    mod.__dict__["xrange"] = range
    src = open(EXAMPLE).read()
    code = compile(src, EXAMPLE, "exec")
    mod.__dict__["__name__"] = "squirmer_example"
    exec(code, mod.__dict__)
    mod.np = _NumpyProxy()
    return mod


# --------------------------------------------------------------------------
# Structured annulus-sector mesh mimicking examples/meshes/donut.geo: meridional half
# plane (rho, z), r in [1, r_out], polar angle theta in [0, pi] from the +z axis;
# boundaries "sphere" (r = 1), "shell" (r = r_out), "symaxis" (rho = 0).
# --------------------------------------------------------------------------
def annulus_nodes(nr, nt, p, r_out):
    """Node coordinates [2, (nr p + 1)(nt p + 1)]: equispaced in the parametric
    coordinates (s, theta) of every element, r = r_out**s (geometric grading).  xi0 runs
    radially outwards, xi1 from theta = pi down to 0 so that detJ = +r > 0."""
    NR, NT = nr * p + 1, nt * p + 1
    s = np.linspace(0.0, 1.0, NR)
    th = np.linspace(np.pi, 0.0, NT)
    r = r_out ** s
    sin = np.sin(th)
    sin[0] = 0.0
    sin[-1] = 0.0          # rho is exactly zero on the axis of symmetry
    R, S = np.meshgrid(r, sin, indexing="ij")
    _, Cz = np.meshgrid(r, np.cos(th), indexing="ij")
    return np.vstack([(R * S).ravel(), (R * Cz).ravel()])


def build_annulus_mesh(nr, nt, p, r_out=100.0):
    live.install_shims()
    from sem.discrete import Mesh
    from sem.geometry import Quadrilateral
    NR, NT = nr * p + 1, nt * p + 1
    gid = np.arange(NR * NT).reshape(NR, NT)
    mesh = Mesh(2)
    mesh.set_nodes(annulus_nodes(nr, nt, p, r_out))
    g = mesh.add_geometry(Quadrilateral(p + 1, p + 1))
    reg = mesh.new_region("interior")
    sphere = mesh.new_boundary("sphere")
    shell = mesh.new_boundary("shell")
    axis = mesh.new_boundary("symaxis")
    c = 0
    for ex in range(nr):
        for ey in range(nt):
            mesh.add_cell(gid[ex * p:ex * p + p + 1, ey * p:ey * p + p + 1].copy(), g, reg)
            if ex == 0:
                mesh.add_boundary_cell(c, sphere, 1, 0)
            if ex == nr - 1:
                mesh.add_boundary_cell(c, shell, 1, 1)
            if ey == 0:
                mesh.add_boundary_cell(c, axis, 1, 2)
            if ey == nt - 1:
                mesh.add_boundary_cell(c, axis, 1, 3)
            c += 1
    return mesh


def make_problem(nr, nt, p, r_out=100.0, kind="FixedSphere"):
    global ORDER
    ORDER = p
    mod = load_example()
    mesh = build_annulus_mesh(nr, nt, p, r_out)
    prob = getattr(mod, kind)(mesh)
    return mod, mesh, prob
