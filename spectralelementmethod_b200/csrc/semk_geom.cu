// semk_geom.cu -- K1: per-element geometric factors (set-up, runs once per mesh).
//
// Reference being replaced, per element (all Python/NumPy/LAPACK there):
//   x_phys  = compute_coeffs_grid_eq(mesh.nodes[:, L2G])     sem/mapping.py:98-103,
//             two lu_solve sweeps with the equispaced->GLL matrix E
//                                                   sem/basis_functions.py:599-624
//   J[i,a]  = d x_i / d xi_a = gradient(x_phys).swapaxes(0,1)   sem/mapping.py:113
//   detJ, invJ = det_inv_2x2(J)  (adjugate * (1/det))        sem/linalg.py:105-115
//   JxW     = (detJ * w_m) * w_n                          sem/quadratures.py:268-275
//   G_ab    = JxW * sum_i invJ[a,i] invJ[b,i]  (implicit in examples/poisson.py:166-193)
//
// Differences by design (SURVEY.md section 0 fact 5 / hard part 7): E^{-1} is
// applied as an explicit, correctly rounded matrix instead of an LU solve, and
// an element-local origin is subtracted before the contractions so that the
// cancellation in J = D x does not grow with 1/h.  One CTA per element slot,
// one thread per node; this kernel is not on the per-iteration path.
#include "semk_common.cuh"

namespace {

template <int N>
__global__ void __launch_bounds__(((N * N + 31) / 32) * 32)
    geom_kernel(int64_t n_elem, const double *__restrict__ nodes_x,
                const double *__restrict__ nodes_y, const uint32_t *__restrict__ l2g,
                const double *__restrict__ Einv, const double *__restrict__ D,
                const double *__restrict__ w, const int64_t *__restrict__ elem_of_slot,
                double *__restrict__ G, int64_t g_patch_stride, int pe,
                double *__restrict__ JxW,
                double *__restrict__ x_phys, double *__restrict__ Jout,
                double *__restrict__ invJout, double *__restrict__ detJout,
                int32_t *__restrict__ bad_flag) {
  constexpr int NN = N * N;
  __shared__ double sE[NN], sD[NN], sw[N];
  __shared__ double sA[2][NN], sB[2][NN];
  const int k = threadIdx.x;
  const bool on = k < NN;
  const int m = on ? k / N : 0, n = on ? k % N : 0;
  if (on) {
    sE[k] = Einv[k];
    sD[k] = D[k];
  }
  if (k < N) sw[k] = w[k];

  for (int64_t slot = blockIdx.x; slot < n_elem; slot += gridDim.x) {
    const int64_t e = elem_of_slot ? elem_of_slot[slot] : slot;
    if (e < 0) continue;  // empty slot of the engine order (uniform over the CTA): G stays 0
    const uint32_t *row = l2g + e * NN;
    const uint32_t g0 = row[0];
    const double ox = nodes_x[g0], oy = nodes_y[g0];
    __syncthreads();  // tables loaded / previous iteration done with sA, sB
    if (on) {
      const uint32_t g = row[k];
      sA[0][k] = __dsub_rn(nodes_x[g], ox);
      sA[1][k] = __dsub_rn(nodes_y[g], oy);
    }
    __syncthreads();
    // equispaced values -> GLL coefficients: axis 0, then axis 1
    double t0 = 0.0, t1 = 0.0;
    if (on) {
#pragma unroll
      for (int r = 0; r < N; ++r) {
        const double c = sE[m * N + r];
        t0 = fma(c, sA[0][r * N + n], t0);
        t1 = fma(c, sA[1][r * N + n], t1);
      }
      sB[0][k] = t0;
      sB[1][k] = t1;
    }
    __syncthreads();
    double xg = 0.0, yg = 0.0;
    if (on) {
#pragma unroll
      for (int s = 0; s < N; ++s) {
        const double c = sE[n * N + s];
        xg = fma(c, sB[0][m * N + s], xg);
        yg = fma(c, sB[1][m * N + s], yg);
      }
      sA[0][k] = xg;  // all reads of sA happened before the previous barrier
      sA[1][k] = yg;
    }
    __syncthreads();
    if (on) {
      double j00 = 0.0, j01 = 0.0, j10 = 0.0, j11 = 0.0;
#pragma unroll
      for (int r = 0; r < N; ++r) {
        const double d0 = sD[m * N + r];  // d/dxi0 acts on the first index
        const double d1 = sD[n * N + r];  // d/dxi1 acts on the second index
        j00 = fma(d0, sA[0][r * N + n], j00);
        j10 = fma(d0, sA[1][r * N + n], j10);
        j01 = fma(d1, sA[0][m * N + r], j01);
        j11 = fma(d1, sA[1][m * N + r], j11);
      }
      const double det = __dsub_rn(__dmul_rn(j00, j11), __dmul_rn(j01, j10));
      if (!(det > 0.0)) atomicExch(bad_flag, 1);
      const double rdet = __drcp_rn(det);
      const double i00 = __dmul_rn(j11, rdet), i01 = __dmul_rn(-j01, rdet);
      const double i10 = __dmul_rn(-j10, rdet), i11 = __dmul_rn(j00, rdet);
      const double jw = __dmul_rn(__dmul_rn(det, sw[m]), sw[n]);
      if (G) {
        // patch-interleaved layout: ((c*N + m)*PE + le)*N + n inside the patch block
        const int64_t patch = slot / pe;
        const int lp = (int)(slot - patch * pe);
        const int64_t grow = (int64_t)N * pe;
        double *gs = G + patch * g_patch_stride + (int64_t)m * grow + lp * N + n;
        gs[0] = jw * (i00 * i00 + i01 * i01);
        gs[N * grow] = jw * (i00 * i10 + i01 * i11);
        gs[2 * N * grow] = jw * (i10 * i10 + i11 * i11);
      }
      if (JxW) JxW[e * NN + k] = jw;
      if (detJout) detJout[e * NN + k] = det;
      if (x_phys) {
        x_phys[(e * 2 + 0) * NN + k] = xg + ox;
        x_phys[(e * 2 + 1) * NN + k] = yg + oy;
      }
      if (Jout) {
        double *jp = Jout + e * 4 * NN;
        jp[k] = j00;
        jp[NN + k] = j01;
        jp[2 * NN + k] = j10;
        jp[3 * NN + k] = j11;
      }
      if (invJout) {
        double *ip = invJout + e * 4 * NN;
        ip[k] = i00;
        ip[NN + k] = i01;
        ip[2 * NN + k] = i10;
        ip[3 * NN + k] = i11;
      }
    }
  }
}

// G from the reference's own invJ / detJxW arrays (parity tier T1).
__global__ void gfactors_from_invj_kernel(int n1, int64_t n_elem,
                                          const double *__restrict__ invJ,
                                          const double *__restrict__ JxW,
                                          const int64_t *__restrict__ elem_of_slot,
                                          double *__restrict__ G, int64_t g_patch_stride,
                                          int pe) {
  const int NN = n1 * n1;
  const int64_t total = n_elem * NN;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t slot = i / NN;
    const int k = (int)(i - slot * NN);
    const int64_t e = elem_of_slot ? elem_of_slot[slot] : slot;
    if (e < 0) continue;
    const double *ip = invJ + e * 4 * NN;
    const double i00 = ip[k], i01 = ip[NN + k], i10 = ip[2 * NN + k], i11 = ip[3 * NN + k];
    const double jw = JxW[e * NN + k];
    const int m = k / n1, n = k - m * n1;
    const int64_t patch = slot / pe;
    const int lp = (int)(slot - patch * pe);
    const int64_t row = (int64_t)n1 * pe;
    double *gs = G + patch * g_patch_stride + (int64_t)m * row + lp * n1 + n;
    gs[0] = jw * (i00 * i00 + i01 * i01);
    gs[n1 * row] = jw * (i00 * i10 + i01 * i11);
    gs[2 * n1 * row] = jw * (i10 * i10 + i11 * i11);
  }
}

// G <- weight * G, node by node (weight in the reference's element-local layout).
__global__ void scale_gfactors_kernel(int n1, int64_t n_elem, const double *__restrict__ weight,
                                      const int64_t *__restrict__ elem_of_slot,
                                      double *__restrict__ G, int64_t g_patch_stride, int pe) {
  const int NN = n1 * n1;
  const int64_t total = n_elem * NN;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t slot = i / NN;
    const int k = (int)(i - slot * NN);
    const int64_t e = elem_of_slot ? elem_of_slot[slot] : slot;
    if (e < 0) continue;
    const double w = weight[e * NN + k];
    const int m = k / n1, n = k - m * n1;
    const int64_t patch = slot / pe;
    const int lp = (int)(slot - patch * pe);
    const int64_t row = (int64_t)n1 * pe;
    double *gs = G + patch * g_patch_stride + (int64_t)m * row + lp * n1 + n;
    gs[0] *= w;
    gs[n1 * row] *= w;
    gs[2 * n1 * row] *= w;
  }
}

}  // namespace

extern "C" int semk_geom_factors_f64(int n1, int64_t n_elem, const double *nodes_x,
                                     const double *nodes_y, const uint32_t *l2g,
                                     const double *Einv, const double *D, const double *w,
                                     const int64_t *elem_of_slot, double *G,
                                     int64_t g_patch_stride, int elems_per_patch, double *JxW,
                                     double *x_phys, double *J, double *invJ, double *detJ,
                                     int32_t *bad_flag, void *stream) {
  SEMK_REQUIRE(n_elem >= 0, "semk_geom_factors_f64: negative n_elem");
  SEMK_REQUIRE(nodes_x && nodes_y && l2g && Einv && D && w && bad_flag,
               "semk_geom_factors_f64: null input pointer");
  SEMK_REQUIRE(!G || (elems_per_patch >= 1 &&
                      g_patch_stride >= (int64_t)3 * n1 * n1 * elems_per_patch),
               "semk_geom_factors_f64: g_patch_stride too small");
  if (n_elem == 0) return SEMK_OK;
  const unsigned grid = (unsigned)(n_elem < (1 << 20) ? n_elem : (1 << 20));
#define SEMK_CALL(NV)                                                                      \
  geom_kernel<NV><<<grid, ((NV * NV + 31) / 32) * 32, 0, semk_stream(stream)>>>(           \
      n_elem, nodes_x, nodes_y, l2g, Einv, D, w, elem_of_slot, G, g_patch_stride,           \
      elems_per_patch > 0 ? elems_per_patch : 1, JxW, x_phys, J, invJ, detJ, bad_flag)
  SEMK_DISPATCH_N1(n1, SEMK_CALL)
#undef SEMK_CALL
  SEMK_LAUNCH_CHECK("geom_kernel");
  return SEMK_OK;
}

extern "C" int semk_gfactors_from_invj_f64(int n1, int64_t n_elem, const double *invJ,
                                           const double *JxW, const int64_t *elem_of_slot,
                                           double *G, int64_t g_patch_stride,
                                           int elems_per_patch, void *stream) {
  SEMK_REQUIRE(n1 >= 2 && n1 <= SEMK_MAX_N1, "semk_gfactors_from_invj_f64: bad n1");
  SEMK_REQUIRE(invJ && JxW && G, "semk_gfactors_from_invj_f64: null pointer");
  SEMK_REQUIRE(elems_per_patch >= 1 && g_patch_stride >= (int64_t)3 * n1 * n1 * elems_per_patch,
               "semk_gfactors_from_invj_f64: g_patch_stride too small");
  if (n_elem <= 0) return SEMK_OK;
  const int NN = n1 * n1;
  const int64_t total = n_elem * NN;
  const int block = 256;
  const int64_t want = (total + block - 1) / block;
  const unsigned grid = (unsigned)(want < 148 * 32 ? want : 148 * 32);
  gfactors_from_invj_kernel<<<grid, block, 0, semk_stream(stream)>>>(
      n1, n_elem, invJ, JxW, elem_of_slot, G, g_patch_stride, elems_per_patch);
  SEMK_LAUNCH_CHECK("gfactors_from_invj_kernel");
  return SEMK_OK;
}

extern "C" int semk_scale_gfactors_f64(int n1, int64_t n_elem, const double *weight,
                                       const int64_t *elem_of_slot, double *G,
                                       int64_t g_patch_stride, int elems_per_patch,
                                       void *stream) {
  SEMK_REQUIRE(n1 >= 2 && n1 <= SEMK_MAX_N1, "semk_scale_gfactors_f64: bad n1");
  SEMK_REQUIRE(weight && G, "semk_scale_gfactors_f64: null pointer");
  SEMK_REQUIRE(elems_per_patch >= 1 && g_patch_stride >= (int64_t)3 * n1 * n1 * elems_per_patch,
               "semk_scale_gfactors_f64: g_patch_stride too small");
  if (n_elem <= 0) return SEMK_OK;
  const int64_t total = n_elem * n1 * n1;
  const int block = 256;
  const int64_t want = (total + block - 1) / block;
  const unsigned grid = (unsigned)(want < 148 * 32 ? want : 148 * 32);
  scale_gfactors_kernel<<<grid, block, 0, semk_stream(stream)>>>(n1, n_elem, weight, elem_of_slot, G,
                                                                g_patch_stride, elems_per_patch);
  SEMK_LAUNCH_CHECK("scale_gfactors_kernel");
  return SEMK_OK;
}
