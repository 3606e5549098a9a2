"""The C-ABI shared library: loads, exports every symbol include/semk.h
declares, and its host-side plan builder (pure C++, no GPU) produces
consistent tables.  No compute kernels are launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from spectralelementmethod_b200 import _lib, meshgen, operators

HEADER = os.path.join(ROOT, "include", "semk.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(semk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libsemk.so does not export %s" % n
    # ... and the ctypes binding covers exactly the declared API
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_channel():
    lib = _lib.load()
    assert lib.semk_version() == 100
    out = ctypes.c_void_p()
    rc = lib.semk_hostplan_create(1, 1, 1, None, None, 0, 16, None, ctypes.byref(out))
    assert rc == _lib.ERR_UNSUPPORTED and b"n1" in lib.semk_last_error()
    with pytest.raises(NotImplementedError):
        _lib.check(rc)
    assert lib.semk_device_available() in (0, 1)


def test_struct_layout_matches_header(tmp_path):
    # compile the header with gcc and compare sizeof / offsetof of every field with ctypes
    fields = [f[0] for f in _lib.semk_op._fields_]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "semk.h"', 'int main(void) {',
           '  printf("%zu %zu\\n", sizeof(struct semk_op), sizeof(struct semk_pcg_info));']
    src += ['  printf("%%zu\\n", offsetof(struct semk_op, %s));' % f for f in fields]
    sc_fields = [f[0] for f in _lib.semk_sc_op._fields_]
    src += ['  printf("%zu\\n", sizeof(struct semk_sc_op));']
    src += ['  printf("%%zu\\n", offsetof(struct semk_sc_op, %s));' % f for f in sc_fields]
    co_fields = [f[0] for f in _lib.semk_sc_coarse._fields_]
    src += ['  printf("%zu\\n", sizeof(struct semk_sc_coarse));']
    src += ['  printf("%%zu\\n", offsetof(struct semk_sc_coarse, %s));' % f for f in co_fields]
    top_fields = [f[0] for f in _lib.semk_sc_top._fields_]
    src += ['  printf("%zu\\n", sizeof(struct semk_sc_top));']
    src += ['  printf("%%zu\\n", offsetof(struct semk_sc_top, %s));' % f for f in top_fields]
    st_fields = [f[0] for f in _lib.semk_stokes_op._fields_]
    src += ['  printf("%zu\\n", sizeof(struct semk_stokes_op));']
    src += ['  printf("%%zu\\n", offsetof(struct semk_stokes_op, %s));' % f for f in st_fields]
    src += ['  return 0; }']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    import subprocess
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == ctypes.sizeof(_lib.semk_op)
    assert int(out[1]) == ctypes.sizeof(_lib.semk_pcg_info) == 24
    for f, off in zip(fields, out[2:]):
        assert getattr(_lib.semk_op, f).offset == int(off), f
    rest = out[2 + len(fields):]
    assert int(rest[0]) == ctypes.sizeof(_lib.semk_sc_op)
    for f, off in zip(sc_fields, rest[1:]):
        assert getattr(_lib.semk_sc_op, f).offset == int(off), f
    rest = rest[1 + len(sc_fields):]
    assert int(rest[0]) == ctypes.sizeof(_lib.semk_sc_coarse)
    for f, off in zip(co_fields, rest[1:]):
        assert getattr(_lib.semk_sc_coarse, f).offset == int(off), f
    rest = rest[1 + len(co_fields):]
    assert int(rest[0]) == ctypes.sizeof(_lib.semk_sc_top)
    for f, off in zip(top_fields, rest[1:]):
        assert getattr(_lib.semk_sc_top, f).offset == int(off), f
    rest = rest[1 + len(top_fields):]
    assert int(rest[0]) == ctypes.sizeof(_lib.semk_stokes_op)
    for f, off in zip(st_fields, rest[1:]):
        assert getattr(_lib.semk_stokes_op, f).offset == int(off), f


def _plan(nx, ny, p, pe, order=None, dirichlet=None):
    N = p + 1
    l2g = meshgen.structured_node_maps(nx, ny, p)
    n_nodes = (nx * p + 1) * (ny * p + 1)
    sc, ar = _lib.hostplan(N, l2g, n_nodes, order, pe, dirichlet)
    return l2g.reshape(-1, N * N), n_nodes, sc, ar


def check_plan(l2g, n_nodes, sc, ar, pe, dirichlet=None):
    NN = l2g.shape[1]
    E = l2g.shape[0]
    n_patch = sc[_lib.PS_N_PATCH]
    eos = ar[_lib.PA_ELEM_OF_SLOT]
    n_order = eos.size                      # engine slots: elements + empty (-1) padding slots
    assert n_patch == -(-n_order // pe) and sc[_lib.PS_N_SLOT_ELEMS] == n_patch * pe
    ptr = ar[_lib.PA_PATCH_NODE_PTR]
    pnode = ar[_lib.PA_PNODE]
    npriv = ar[_lib.PA_PATCH_NPRIV]
    base = ar[_lib.PA_PATCH_SLOT_BASE]
    ES = sc[_lib.PS_ELOC_STRIDE]
    n1 = int(round(NN ** 0.5))
    assert ES % 8 == 0 and NN * pe <= ES < NN * pe + 8
    # per patch a table [m][le][t]; bring it to [slot][m*n1 + t]
    eloc = ar[_lib.PA_ELOC].reshape(-1, ES)[:, :NN * pe].reshape(-1, n1, pe, n1)
    eloc = eloc.transpose(0, 2, 1, 3).reshape(-1, NN)
    nnodes = ar[_lib.PA_PATCH_NNODES]
    assert np.all(ptr % 4 == 0)
    assert sorted(eos[eos >= 0].tolist()) == list(range(E)) and np.all(eos >= -1)
    pad = pnode == 0xFFFFFFFF
    ids = pnode & _lib.NODE_ID_MASK
    shared_flag = ((pnode & _lib.NODE_SHARED) != 0) & ~pad
    dir_flag = ((pnode & _lib.NODE_DIRICHLET) != 0) & ~pad
    if dirichlet is not None:
        assert np.array_equal(dir_flag[~pad], dirichlet[ids[~pad]].astype(bool))
    else:
        assert not dir_flag.any()
    touched = np.zeros(n_nodes, dtype=int)
    written = np.zeros(n_nodes, dtype=int)      # how many patches write the node to y
    slots_seen = 0
    for p in range(n_patch):
        a, b = ptr[p], ptr[p] + nnodes[p]
        assert np.all(pnode[b:ptr[p + 1]] == 0xFFFFFFFF) and ptr[p + 1] - b < 4
        loc_ids = ids[a:b]
        assert len(set(loc_ids.tolist())) == b - a
        # classes: [private | shared], each ascending
        c1 = npriv[p]
        assert 0 <= c1 <= b - a
        assert not shared_flag[a:a + c1].any() and shared_flag[a + c1:b].all()
        for lo, hi in ((0, c1), (c1, b - a)):
            assert np.all(np.diff(loc_ids[lo:hi]) > 0)
        written[loc_ids[:c1]] += 1
        assert base[p] == slots_seen
        slots_seen += (b - a) - c1
        touched[loc_ids] += 1
        s0, s1 = p * pe, min((p + 1) * pe, n_order)
        live = np.flatnonzero(eos[s0:s1] >= 0) + s0
        assert live.size >= 1
        # eloc reproduces the L2G rows of the patch's elements (empty slots: index 0)
        assert np.array_equal(loc_ids[eloc[live].astype(int)], l2g[eos[live]])
        holes = np.setdiff1d(np.arange(s0, s1), live)
        assert np.all(eloc[holes] == 0)
        assert (b - a) <= sc[_lib.PS_MAX_PATCH_NODES]
    assert slots_seen == sc[_lib.PS_N_SLOTS]
    # private <=> touched by exactly one patch
    is_shared_node = np.zeros(n_nodes, dtype=bool)
    is_shared_node[ids[shared_flag]] = True
    assert sc[_lib.PS_N_PNODE] == pnode.size
    # uniform-stride device blocks mirror the compact tables; identical blocks are stored
    # once and the per-patch header says which ones a patch uses
    PS, ELS = sc[_lib.PS_PN_STRIDE], sc[_lib.PS_EL_STRIDE]
    assert PS % 4 == 0 and PS >= sc[_lib.PS_MAX_PATCH_NODES] and ELS % 8 == 0
    npu, neu = sc[_lib.PS_N_PN_UNIQUE], sc[_lib.PS_N_EL_UNIQUE]
    pnblk = ar[_lib.PA_PNBLK].reshape(npu, PS)
    elblk = ar[_lib.PA_ELBLK].reshape(neu, ELS)
    hdr = ar[_lib.PA_PATCH_HDR].reshape(n_patch, 8)
    niu = sc[_lib.PS_N_INV_UNIQUE]
    invblk = ar[_lib.PA_INVBLK].reshape(niu, -1)
    assert len(set(map(bytes, invblk))) == niu
    assert len(set(map(bytes, pnblk))) == npu and len(set(map(bytes, elblk))) == neu
    for p in range(n_patch):
        h = hdr[p]
        assert h[0] == nnodes[p] and h[1] == npriv[p] and h[2] == base[p]
        assert h[3] == 0
        assert h[5] < npu and h[6] < neu
        full = pnode[ptr[p]:ptr[p] + nnodes[p]]
        assert h[4] == (full & _lib.NODE_ID_MASK).min()
        blk = pnblk[h[5]]
        rel = (full & _lib.NODE_ID_MASK) - h[4]
        assert np.array_equal(blk[:nnodes[p]], rel | (full & ~np.uint32(_lib.NODE_ID_MASK)))
        assert np.all(blk[nnodes[p]:] == 0xFFFFFFFF)
        eb = elblk[h[6]]
        assert np.array_equal(eb[:NN * pe], ar[_lib.PA_ELOC].reshape(-1, ES)[p, :NN * pe])
        # inverse table: node k <- scratch positions (m*RS + le*n1 + t) of its contributions,
        # ascending element slot, 0xffff padded
        W, RS = sc[_lib.PS_INV_WIDTH], ((n1 * pe - 1 + 15) & ~15) + 1
        assert W % 4 == 0 and W >= 4 and sc[_lib.PS_INV_STRIDE] == PS * W and h[7] < niu
        ib = invblk[h[7]].reshape(PS, W)
        tab = eb[:NN * pe].reshape(n1, pe, n1)
        want_inv = [[] for _ in range(nnodes[p])]
        for le in range(min(pe, n_order - p * pe)):
            if eos[p * pe + le] < 0:
                continue
            for m in range(n1):
                for t in range(n1):
                    want_inv[tab[m, le, t]].append(m * RS + le * n1 + t)
        for k in range(nnodes[p]):
            c = len(want_inv[k])
            assert 1 <= c <= W and ib[k, :c].tolist() == want_inv[k] and np.all(ib[k, c:] == 0xFFFF)
        assert np.all(ib[nnodes[p]:] == 0xFFFF)
    # every touched node is written by exactly one patch or is a shared (slot) node
    is_shared_node = np.zeros(n_nodes, dtype=bool)
    is_shared_node[ids[shared_flag]] = True
    assert sc[_lib.PS_N_PNODE] == pnode.size
    assert np.array_equal(written == 1, (touched >= 1) & ~is_shared_node)
    assert not (written > 1).any()
    assert np.array_equal(is_shared_node, touched > 1)
    # CSR of interface slots: every slot exactly once, grouped under its node
    sn = ar[_lib.PA_SHARED_NODE]
    sp = ar[_lib.PA_SHARED_PTR]
    ss = ar[_lib.PA_SHARED_SLOT]
    assert sc[_lib.PS_N_SHARED] == sn.size == is_shared_node.sum()
    assert np.all(np.diff((sn & _lib.NODE_ID_MASK).astype(np.int64)) > 0)
    assert sp[0] == 0 and sp[-1] == ss.size == slots_seen
    # interface tables: affine chunks + per-node records together cover every shared node
    # exactly once with its slots in ascending patch order
    ext = ar[_lib.PA_SHARED_EXT]
    nch, nrec = sc[_lib.PS_N_SHARED_CHUNK], sc[_lib.PS_N_SHARED_REC]
    chunk = ar[_lib.PA_SHARED_CHUNK].reshape(-1, 8)[:nch].astype(np.int64)
    rec = ar[_lib.PA_SHARED_REC].reshape(-1, 8)[:nrec]
    want = {int(sn[i] & _lib.NODE_ID_MASK): (ss[sp[i]:sp[i + 1]].tolist(),
                                             bool(sn[i] & _lib.NODE_DIRICHLET))
            for i in range(sn.size)}
    seen = {}
    for node0, dn, a0, da, b0, db, ln, mask in chunk.tolist():
        assert 1 <= ln <= 32
        dn, da, db = ((v - (1 << 32) if v >= (1 << 31) else v) for v in (dn, da, db))
        for k in range(ln):
            g = node0 + k * dn
            assert g not in seen
            seen[g] = ([a0 + k * da, b0 + k * db], bool((mask >> k) & 1))
    for r in rec.tolist():
        g = r[0] & _lib.NODE_ID_MASK
        cnt = r[1]
        assert cnt >= 3
        lst = r[2:2 + cnt] if cnt <= 6 else r[2:7] + ext[r[7]:r[7] + cnt - 5].tolist()
        assert g not in seen
        seen[g] = (lst, bool(r[0] & _lib.NODE_DIRICHLET))
    assert seen == want
    # both interface tables are sorted by the highest patch involved (staged host apply)
    slot_patch = np.searchsorted(base, np.arange(slots_seen), side="right") - 1
    cmax, rmax = ar[_lib.PA_CHUNK_MAXPATCH][:nch], ar[_lib.PA_REC_MAXPATCH][:nrec]
    assert cmax.size == nch and rmax.size == nrec
    assert np.all(np.diff(cmax) >= 0) and np.all(np.diff(rmax) >= 0)
    for c, (node0, dn, a0, da, b0, db, ln, mask) in zip(cmax.tolist(), chunk.tolist()):
        dn, da, db = ((v - (1 << 32) if v >= (1 << 31) else v) for v in (dn, da, db))
        assert c == slot_patch[b0] == slot_patch[b0 + (ln - 1) * db] > slot_patch[a0]
    for c, r in zip(rmax.tolist(), rec.tolist()):
        cnt = r[1]
        lst = r[2:2 + cnt] if cnt <= 6 else r[2:7] + ext[r[7]:r[7] + cnt - 5].tolist()
        assert c == slot_patch[lst[-1]]
    pmax = ar[_lib.PA_PATCH_MAXNODE]
    for p in range(n_patch):
        assert pmax[p] == ids[ptr[p]:ptr[p] + nnodes[p]].max()
    for lst, _ in want.values():
        assert lst == sorted(lst) and len(lst) >= 2


@pytest.mark.parametrize("nx,ny,p,pe", [(8, 8, 8, 16), (5, 3, 4, 16), (7, 5, 2, 8), (3, 3, 10, 4),
                                        (1, 1, 3, 16), (6, 6, 1, 4)])
def test_hostplan_structured(nx, ny, p, pe):
    from spectralelementmethod_b200.discrete import Mesh
    mesh = meshgen.structured_quad_mesh(nx, ny, p)
    order = operators.default_element_order(mesh, pe)
    rng = np.random.default_rng(0)
    n_nodes = (nx * p + 1) * (ny * p + 1)
    dirichlet = (rng.uniform(size=n_nodes) < 0.2).astype(np.uint8)
    l2g, n_nodes, sc, ar = _plan(nx, ny, p, pe, order, dirichlet)
    check_plan(l2g, n_nodes, sc, ar, pe, dirichlet)
    if (nx, ny, p, pe) == (8, 8, 8, 16):          # 2x8 tiles: a 4x1 arrangement of patches
        assert sc[_lib.PS_N_PATCH] == 4
        assert sc[_lib.PS_MAX_PATCH_NODES] == 17 * 65
        assert sc[_lib.PS_N_SHARED] == 3 * 65


def test_hostplan_deduplicates_table_blocks():
    """Regular numbering: interior / edge / corner patches share their relative node list
    and index table, so the device pools hold 9 blocks however large the mesh."""
    nx, ny, p, pe = 12, 40, 4, 16
    mesh = meshgen.structured_quad_mesh(nx, ny, p)
    order = operators.default_element_order(mesh, pe)
    l2g, n_nodes, sc, ar = _plan(nx, ny, p, pe, order, None)
    assert sc[_lib.PS_N_PATCH] == 30
    assert sc[_lib.PS_N_PN_UNIQUE] == 9 and sc[_lib.PS_N_EL_UNIQUE] <= 9
    check_plan(l2g, n_nodes, sc, ar, pe)


def test_hostplan_tables_do_not_depend_on_the_thread_count():
    """semk_hostplan_create_mt: the per-patch passes run on several threads (contiguous patch
    ranges, thread-local block pools merged in patch order); every table must be byte-identical
    to the single-threaded build -- structured with ragged tiles, scrambled numbering with a
    random element order, irregular vertices, and more threads than patches."""
    rng = np.random.default_rng(8)
    cases = []
    for nx, ny, p, pe in ((23, 41, 4, 16), (7, 19, 8, 16), (13, 26, 3, 32), (2, 3, 5, 16)):
        mesh = meshgen.structured_quad_mesh(nx, ny, p)
        l2g = mesh.node_map_array().reshape(nx * ny, -1)
        d = (rng.uniform(size=mesh.n_nodes) < 0.1).astype(np.uint8)
        cases.append((p + 1, l2g, mesh.n_nodes, operators.default_element_order(mesh, pe), pe, d))
    nx, ny, p = 11, 9, 3
    l2g = meshgen.structured_node_maps(nx, ny, p).reshape(nx * ny, -1)
    n_nodes = (nx * p + 1) * (ny * p + 1)
    perm = rng.permutation(n_nodes).astype(np.uint32)
    cases.append((p + 1, perm[l2g], n_nodes, rng.permutation(nx * ny), 8, None))
    mesh = meshgen.pinwheel_mesh(7, 4, rings=3)
    cases.append((5, mesh.node_map_array().reshape(mesh.n_cells, -1), mesh.n_nodes, None, 8, None))
    for n1, l2g, n_nodes, order, pe, d in cases:
        sc1, ar1 = _lib.hostplan(n1, l2g, n_nodes, order, pe, d, threads=1)
        for t in (2, 3, 8):
            sc, ar = _lib.hostplan(n1, l2g, n_nodes, order, pe, d, threads=t)
            assert sc == sc1
            for k in ar1:
                assert np.array_equal(ar[k], ar1[k]), (k, t)


def test_hostplan_scrambled_numbering_and_order():
    """Arbitrary node numbering (as after RCM) and a random element order."""
    nx, ny, p, pe = 6, 5, 3, 8
    l2g = meshgen.structured_node_maps(nx, ny, p)
    n_nodes = (nx * p + 1) * (ny * p + 1)
    rng = np.random.default_rng(3)
    perm = rng.permutation(n_nodes).astype(np.uint32)
    l2g = perm[l2g]
    order = rng.permutation(nx * ny)
    sc, ar = _lib.hostplan(p + 1, l2g, n_nodes, order, pe, None)
    check_plan(l2g.reshape(nx * ny, -1), n_nodes, sc, ar, pe)
    assert np.array_equal(ar[_lib.PA_ELEM_OF_SLOT], order)


@pytest.mark.parametrize("n_cells,p,rings,pe", [(5, 3, 1, 16), (6, 2, 2, 8), (3, 4, 2, 4), (7, 3, 3, 16)])
def test_hostplan_irregular_vertices(n_cells, p, rings, pe):
    """3 / 5 / 6 / 7 cells around a vertex: inverse tables wider than 4 entries."""
    mesh = meshgen.pinwheel_mesh(n_cells, p, rings=rings)
    l2g = mesh.node_map_array().reshape(mesh.n_cells, -1)
    sc, ar = _lib.hostplan(p + 1, l2g, mesh.n_nodes, None, pe, None)
    check_plan(l2g, mesh.n_nodes, sc, ar, pe)
    if n_cells >= 5 and pe >= n_cells:
        assert sc[_lib.PS_INV_WIDTH] == 8
    if n_cells == 3:
        assert sc[_lib.PS_INV_WIDTH] == 4


def test_hostplan_ragged_tiles_are_padded():
    """A mesh that is not a whole number of 2x8 tiles: the default order pads the ragged
    tiles with empty slots, so every patch stays one compact tile."""
    nx, ny, p, pe = 5, 20, 3, 16
    mesh = meshgen.structured_quad_mesh(nx, ny, p)
    order = operators.default_element_order(mesh, pe)
    assert order.size == 3 * 3 * 16 and (order == -1).sum() == 3 * 3 * 16 - nx * ny
    l2g, n_nodes, sc, ar = _plan(nx, ny, p, pe, order, None)
    check_plan(l2g, n_nodes, sc, ar, pe)
    assert sc[_lib.PS_N_PATCH] == 9
    assert sc[_lib.PS_MAX_PATCH_NODES] == (2 * p + 1) * (8 * p + 1)     # never more than a full tile
    # holes inside a patch, and a patch of holes only is rejected
    bad = np.concatenate([np.arange(nx * ny), np.full(16, -1)])
    with pytest.raises(ValueError):
        _lib.hostplan(p + 1, l2g, n_nodes, bad, pe, None)
    with pytest.raises(ValueError):
        _lib.hostplan(p + 1, l2g, n_nodes, np.arange(nx * ny - 1), pe, None)   # an element missing


def test_hostplan_rejects_bad_input():
    l2g = meshgen.structured_node_maps(2, 2, 2)
    with pytest.raises(ValueError):
        _lib.hostplan(3, l2g, 25, np.array([0, 1, 1, 2]), 4, None)      # not a permutation
    with pytest.raises(ValueError):
        _lib.hostplan(3, l2g, 20, None, 4, None)                          # id out of range
    with pytest.raises(ValueError):
        _lib.hostplan(3, l2g, 25, None, 4, np.zeros(7, dtype=np.uint8))   # mask length
    with pytest.raises(NotImplementedError):
        _lib.hostplan(18, np.zeros((1, 324), dtype=np.uint32), 400, None, 4, None)


def test_patch_size_choice_and_smem_budget():
    for n1 in range(2, 18):
        pe = operators.choose_elems_per_patch(n1)
        bx, by = operators._TILES[pe]
        p = n1 - 1
        smem = operators.patch_smem_bytes(n1, pe, (bx * p + 1) * (by * p + 1))
        assert smem <= 227 * 1024
        assert operators.g_patch_stride_of(n1, pe) % 2 == 0 and operators.eloc_patch_stride_of(n1, pe) % 8 == 0
    assert operators.choose_elems_per_patch(9) == 16


def test_default_element_order_tiles():
    mesh = meshgen.structured_quad_mesh(8, 8, 1)
    order = operators.default_element_order(mesh, 16)
    first = sorted(order[:16].tolist())
    assert first == sorted(ex * 8 + ey for ex in range(2) for ey in range(8))
    o44 = operators.default_element_order(mesh, 16, tile=(4, 4))
    assert sorted(o44[:16].tolist()) == sorted(ex * 8 + ey for ex in range(4) for ey in range(4))
    mesh._structured_shape = None
    mo = operators.default_element_order(mesh, 16)
    assert sorted(mo.tolist()) == list(range(64))
    # a Morton run of 16 cells on an 8x8 grid is a 4x4 block
    ex, ey = np.divmod(mo[:16], 8)
    assert ex.max() - ex.min() == 3 and ey.max() - ey.min() == 3


def test_product_fails_loudly_without_gpu():
    lib = _lib.load()
    if lib.semk_device_available():
        pytest.skip("a GPU is present")
    from conftest import build_package_case
    mesh, mngr = build_package_case("S", 2, 2, 3, False, False)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mngr.poisson_operator()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        next(mngr.finite_elements(x_phys=True, Jacobian=True))
