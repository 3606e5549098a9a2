// semk_field.cu -- field evaluation / output.
//
// DOFManager.values_at_nodes (sem/discrete.py:235-258): for every element, the GLL
// coefficients gathered through the L2G map are resampled on the equispaced grid of
// the element (TensorProduct.interpolate_on_grid_eq, sem/basis_functions.py:539-569:
// the 1-D matrix E = `_interp_eq_mat` applied along the last axis, then along the
// first) and written back through the L2G map.  Where elements share a node the
// reference's loop lets the LAST element win; `winner` marks, per element-local
// entry, whether that element is the last one containing the node, so the result
// is the reference's bit pattern regardless of scheduling.  HBM-bound (one gather
// and one scatter per element-local node), any order 1..16.
#include "semk_common.cuh"

namespace {

constexpr int kFieldThreads = 256;

__global__ void __launch_bounds__(kFieldThreads)
    values_at_nodes_kernel(int n1, int pe, int64_t n_elem, const uint32_t *__restrict__ l2g,
                           const uint8_t *__restrict__ winner, const double *__restrict__ Emat,
                           const double *__restrict__ coeffs, double *__restrict__ values) {
  extern __shared__ __align__(16) double field_smem[];
  const int NN = n1 * n1;
  double *sE = field_smem;       // [n1][n1]
  double *sU = sE + NN;          // [pe][n1][n1] coefficients
  double *sT = sU + pe * NN;     // [pe][n1][n1] after the pass along the last axis
  const int tid = threadIdx.x;
  for (int i = tid; i < NN; i += kFieldThreads) sE[i] = Emat[i];
  const int64_t n_groups = (n_elem + pe - 1) / pe;
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    __syncthreads();  // sE is loaded / the previous group is done with sU, sT
    const int64_t e0 = grp * pe;
    const int ne = (int)((n_elem - e0) < (int64_t)pe ? (n_elem - e0) : (int64_t)pe);
    const int tot = ne * NN;
    const uint32_t *ids = l2g + e0 * NN;
    for (int idx = tid; idx < tot; idx += kFieldThreads) sU[idx] = coeffs[ids[idx]];
    __syncthreads();
    for (int idx = tid; idx < tot; idx += kFieldThreads) {
      const int le = idx / NN, k = idx - le * NN;
      const int m = k / n1, n = k - m * n1;
      const double *urow = sU + le * NN + m * n1;
      const double *erow = sE + n * n1;
      double acc = 0.0;
      for (int s = 0; s < n1; ++s) acc = fma(erow[s], urow[s], acc);
      sT[idx] = acc;
    }
    __syncthreads();
    for (int idx = tid; idx < tot; idx += kFieldThreads) {
      if (!winner[e0 * NN + idx]) continue;
      const int le = idx / NN, k = idx - le * NN;
      const int m = k / n1, n = k - m * n1;
      const double *tcol = sT + le * NN + n;
      const double *erow = sE + m * n1;
      double acc = 0.0;
      for (int r = 0; r < n1; ++r) acc = fma(erow[r], tcol[r * n1], acc);
      values[ids[idx]] = acc;
    }
  }
}

}  // namespace

extern "C" int semk_values_at_nodes_f64(int n1, int64_t n_elem, const uint32_t *l2g,
                                        const uint8_t *winner, const double *Emat,
                                        const double *coeffs, double *values, void *stream) {
  SEMK_REQUIRE(n_elem > 0 && l2g && winner && Emat && coeffs && values && coeffs != values,
               "semk_values_at_nodes_f64: bad argument");
  if (n1 < 2 || n1 > SEMK_MAX_N1) {
    semk_set_error("semk_values_at_nodes_f64: n1 outside [2, 17]");
    return SEMK_ERR_UNSUPPORTED;
  }
  const int NN = n1 * n1;
  int pe = kFieldThreads / NN;
  if (pe < 1) pe = 1;
  const size_t smem = sizeof(double) * (size_t)(NN + 2 * pe * NN);
  const int64_t groups = (n_elem + pe - 1) / pe;
  const unsigned grid = (unsigned)(groups < 148 * 8 ? groups : 148 * 8);
  values_at_nodes_kernel<<<grid, kFieldThreads, smem, semk_stream(stream)>>>(
      n1, pe, n_elem, l2g, winner, Emat, coeffs, values);
  SEMK_LAUNCH_CHECK("values_at_nodes_kernel");
  return SEMK_OK;
}
