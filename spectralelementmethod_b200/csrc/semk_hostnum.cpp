// semk_hostnum.cpp -- host-side (no GPU) integer-table helpers, multi-threaded.
//
// What it replaces: the reference renumbers the mesh for static condensation in Python,
// `DOFManagerSC._do_static_condensation` (sem/discrete.py:314-359: np.unique of the
// element-exterior ids, then the sorted interior ids) followed by `Mesh._permute_nodes`
// (sem/discrete.py:1115-1127: permute the coordinate columns, rewrite every node map through
// the inverse permutation).  The host mirror does the same with whole-array NumPy calls
// (discrete.py); at config 2 (85 M map entries, 67 M nodes) those single-threaded random
// gathers / scatters take 3.5 s -- more than the whole device solve.  Here the same tables
// are produced by OpenMP loops: identical integers (bit-exact, tests/test_host_api.py).
#include <omp.h>
#include <stdint.h>

#include <cstring>
#include <vector>

#include "../../include/semk.h"

namespace {
int pick_threads(int requested) {
  int n = requested > 0 ? requested : omp_get_num_procs();
  if (n > 64) n = 64;
  return n < 1 ? 1 : n;
}
}  // namespace

// node maps of a structured nx x ny mesh of order p: map[c][m][n] = base(c) + m*NY + n
extern "C" int semk_host_structured_maps(int64_t nx, int64_t ny, int32_t p, int64_t node_offset,
                                         uint32_t *out, int32_t n_threads) {
  if (nx < 1 || ny < 1 || p < 1 || !out) return SEMK_ERR_INVALID;
  const int64_t NY = ny * p + 1, n1 = p + 1;
  if ((nx * p + 1) * NY + node_offset > 4294967295LL) return SEMK_ERR_UNSUPPORTED;
#pragma omp parallel for num_threads(pick_threads(n_threads)) schedule(static)
  for (int64_t c = 0; c < nx * ny; ++c) {
    const int64_t ex = c / ny, ey = c - ex * ny;
    const int64_t base = ex * p * NY + ey * p + node_offset;
    uint32_t *o = out + c * n1 * n1;
    for (int64_t m = 0; m < n1; ++m)
      for (int64_t n = 0; n < n1; ++n) o[m * n1 + n] = (uint32_t)(base + m * NY + n);
  }
  return SEMK_OK;
}

// Exterior-first renumbering of a homogeneous mesh, in place.
//   maps      [n_cells][nn] uint32 node ids (rewritten through the inverse permutation)
//   ext_idx / int_idx : positions inside a cell of its exterior / interior nodes
//   nodes     ndim rows of `node_stride` doubles (first n_nodes permuted in place), or NULL
//   order_out [n_nodes]: new node k is old node order_out[k]
// Returns SEMK_ERR_UNSUPPORTED when the mesh is not one the fast path covers (a node that is
// exterior in one cell and interior in another, repeated interior ids, nodes in no cell): the
// caller then takes the NumPy path, which reproduces the reference's behaviour for those.
extern "C" int semk_host_sc_numbering(int64_t n_nodes, int64_t n_cells, int32_t nn, uint32_t *maps,
                                      const int32_t *ext_idx, int32_t n_ext_idx,
                                      const int32_t *int_idx, int32_t n_int_idx, double *nodes,
                                      int32_t ndim, int64_t node_stride, int64_t *n_ext_out,
                                      int64_t *n_int_out, uint32_t *order_out, int32_t n_threads) {
  if (n_nodes < 1 || n_cells < 1 || nn < 1 || !maps || !ext_idx || !int_idx || !order_out ||
      n_nodes > 4294967295LL)
    return SEMK_ERR_INVALID;
  const int T = pick_threads(n_threads);
  std::vector<uint8_t> mext((size_t)n_nodes, 0), mint((size_t)n_nodes, 0);
  bool bad = false;
#pragma omp parallel for num_threads(T) schedule(static) reduction(|| : bad)
  for (int64_t c = 0; c < n_cells; ++c) {
    const uint32_t *m = maps + c * nn;
    for (int k = 0; k < n_ext_idx; ++k) {
      if (m[ext_idx[k]] >= (uint64_t)n_nodes) { bad = true; continue; }
      mext[m[ext_idx[k]]] = 1;      // concurrent writes of the same value: benign
    }
    for (int k = 0; k < n_int_idx; ++k) {
      if (m[int_idx[k]] >= (uint64_t)n_nodes) { bad = true; continue; }
      mint[m[int_idx[k]]] = 1;
    }
  }
  if (bad) return SEMK_ERR_INVALID;
  // block-wise exclusive prefix sums of the two marks
  const int64_t B = 1 << 16;
  const int64_t nb = (n_nodes + B - 1) / B;
  std::vector<int64_t> ce((size_t)nb + 1, 0), ci((size_t)nb + 1, 0);
  bool mixed = false;
#pragma omp parallel for num_threads(T) schedule(static) reduction(|| : mixed)
  for (int64_t b = 0; b < nb; ++b) {
    int64_t e = 0, i = 0;
    const int64_t hi = (b + 1) * B < n_nodes ? (b + 1) * B : n_nodes;
    for (int64_t k = b * B; k < hi; ++k) {
      e += mext[k];
      i += mint[k];
      if (mext[k] == mint[k]) mixed = true;   // both (exterior and interior) or neither
    }
    ce[b + 1] = e;
    ci[b + 1] = i;
  }
  if (mixed) return SEMK_ERR_UNSUPPORTED;
  for (int64_t b = 0; b < nb; ++b) {
    ce[b + 1] += ce[b];
    ci[b + 1] += ci[b];
  }
  const int64_t n_ext = ce[nb], n_int = ci[nb];
  if (n_int != n_cells * (int64_t)n_int_idx) return SEMK_ERR_UNSUPPORTED;  // repeated interior ids
  std::vector<uint32_t> inv((size_t)n_nodes);
#pragma omp parallel for num_threads(T) schedule(static)
  for (int64_t b = 0; b < nb; ++b) {
    int64_t e = ce[b], i = n_ext + ci[b];
    const int64_t hi = (b + 1) * B < n_nodes ? (b + 1) * B : n_nodes;
    for (int64_t k = b * B; k < hi; ++k) {
      const int64_t id = mext[k] ? e++ : i++;
      inv[k] = (uint32_t)id;
      order_out[id] = (uint32_t)k;
    }
  }
#pragma omp parallel for num_threads(T) schedule(static)
  for (int64_t j = 0; j < n_cells * (int64_t)nn; ++j) maps[j] = inv[maps[j]];
  if (nodes) {
    std::vector<double> tmp((size_t)n_nodes);
    for (int d = 0; d < ndim; ++d) {
      double *row = nodes + (int64_t)d * node_stride;
#pragma omp parallel for num_threads(T) schedule(static)
      for (int64_t k = 0; k < n_nodes; ++k) tmp[k] = row[order_out[k]];
#pragma omp parallel for num_threads(T) schedule(static)
      for (int64_t k = 0; k < n_nodes; ++k) row[k] = tmp[k];
    }
  }
  *n_ext_out = n_ext;
  *n_int_out = n_int;
  return SEMK_OK;
}
