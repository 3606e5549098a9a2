"""Freeze golden vectors for the Gmsh importer.

TEST INFRASTRUCTURE ONLY; runs in the development container, where the
reference tree exists.  For a few synthetic Gmsh 2.2 binary files (written by
spectralelementmethod_b200.meshgen.write_gmsh22_binary, with shuffled node and
element numbering) the UNMODIFIED reference loader ``sem.grid_importers.load_msh``
(sem/grid_importers.py:45-69; shim 4 of oracle/live_reference.py strips the
``U`` of its ``open(..., 'rbU')``) produces the mesh; this script stores the
file bytes next to what the reference made of them:

    tests/golden/msh_<name>.npz : msh_bytes (uint8), nodes, node_maps [E, n1, n1],
        region_ids, adjacency [E, 4] (-1 = none), boundary rows
        (cell, boundary id, face) in the reference's per-cell insertion order,
        names.

    python oracle/make_golden_msh.py
"""
import builtins
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import live_reference as lr  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")

# name, nx, ny, p, kind, shuffle seed
CASES = [
    ("q1_3x2", 3, 2, 1, "S", None),
    ("q2_4x3_shuffled", 4, 3, 2, "C", 7),
    ("q3_5x4_shuffled", 5, 4, 3, "C", 11),
    ("q5_3x3_shuffled", 3, 3, 5, "C", 1),
    ("q8_2x2_shuffled", 2, 2, 8, "S", 3),
    ("q10_2x3", 2, 3, 10, "C", None),
]
# (a single-cell mesh makes the reference's neighbour search index an empty array,
#  sem/grid_importers.py:253 -- no golden vector for it; the package handles it)


def reference_load(path):
    lr.install_shims()
    import sem.grid_importers as ref_gi

    def open_no_u(file, mode="r", *a, **kw):
        return builtins.open(file, mode.replace("U", ""), *a, **kw)
    ref_gi.open = open_no_u              # module-level name shadows the builtin
    return ref_gi.load_msh(path, 2)


def snapshot(mesh):
    maps = np.stack([c.node_ind_lexicographic for c in mesh.cells])
    regions = np.array([c.region_id for c in mesh.cells], dtype=np.int64)
    adj = np.array([[-1 if v is None else v for v in mesh._adj_map[i]] for i in range(mesh.n_cells)],
                   dtype=np.int64)
    rows = []
    for cell in sorted(mesh._boundary_map):
        for bnd_id, lst in mesh._boundary_map[cell].items():
            for bd in lst:
                rows.append((cell, bnd_id, bd.index, bd.ndim))
    return dict(nodes=np.ascontiguousarray(mesh.nodes), node_maps=maps, region_ids=regions,
                adjacency=adj, boundary=np.array(rows, dtype=np.int64).reshape(-1, 4),
                region_names=np.array(mesh._region_names), boundary_names=np.array(mesh._boundary_names))


def main():
    from spectralelementmethod_b200 import meshgen
    os.makedirs(OUT, exist_ok=True)
    for name, nx, ny, p, kind, seed in CASES:
        with tempfile.TemporaryDirectory() as tmp:
            path = meshgen.write_gmsh22_binary(os.path.join(tmp, "m.msh"), nx, ny, p, kind,
                                               shuffle_seed=seed)
            raw = np.fromfile(path, dtype=np.uint8)
            snap = snapshot(reference_load(path))
        snap["msh_bytes"] = raw
        snap["meta"] = np.array([nx, ny, p, -1 if seed is None else seed])
        np.savez_compressed(os.path.join(OUT, "msh_%s.npz" % name), **snap)
        print(name, snap["node_maps"].shape, "boundary faces:", len(snap["boundary"]),
              "file bytes:", raw.size)


if __name__ == "__main__":
    main()
