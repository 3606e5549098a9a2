// semk_locate.cu -- batched point location and field evaluation on the device.
//
// Reference being replaced, one point at a time on the host:
//   DOFManager.find_elem_containing_point   sem/discrete.py:263-280
//       sort ALL cells by centroid distance, try them in that order
//   Mapping.inv                             sem/mapping.py:146-178
//       Newton on x(xi) - x_phys from xi = 0, at most 8 steps, |dx| <= 1e-8,
//       (rootfind.newton, sem/rootfind.py:22-53); inside iff -1 <= xi <= 1
//   DOFManager.interpolate                  sem/discrete.py:221-233
//       fe.interpolate(coeffs[fe.node_ind], xi): tensor barycentric Lagrange
//
// Here: one thread per point.  Candidate elements come from a uniform bin grid over the
// mesh (host-built CSR: every element is listed in all bins its bounding box touches), and
// are tried in ascending centroid distance -- the reference's order restricted to the
// elements that can contain the point, so the element found is the reference's.  The
// mapping and its Jacobian are evaluated from the element's GLL-point coordinates:
//   x(xi)      = sum_mn X[m][n] l_m(xi0) l_n(xi1)
//   dx/dxi0    = sum_mn X[m][n] l'_m(xi0) l_n(xi1),   l'_m(xi) = sum_k l_k(xi) D[k][m]
// (the reference interpolates the nodal Jacobian, which is the same polynomial).
#include "semk_common.cuh"

namespace {

constexpr int kMaxN = SEMK_MAX_N1;
constexpr int kLocThreads = 128;
constexpr double kInsideSlack = 1e-10;

// l_j(x) for the N nodes by the barycentric formula; an exact node hit gives the unit vector
__device__ __forceinline__ void lagrange_all(int N, const double *__restrict__ nodes,
                                             const double *__restrict__ bw, double x,
                                             double *__restrict__ L) {
  int hit = -1;
  double sum = 0.0;
  for (int j = 0; j < N; ++j) {
    const double d = x - nodes[j];
    if (d == 0.0) hit = j;
    const double t = bw[j] / d;
    L[j] = t;
    sum += t;
  }
  if (hit >= 0) {
    for (int j = 0; j < N; ++j) L[j] = (j == hit) ? 1.0 : 0.0;
  } else {
    const double inv = 1.0 / sum;
    for (int j = 0; j < N; ++j) L[j] *= inv;
  }
}

// dL[m] = sum_k L[k] D[k][m]
__device__ __forceinline__ void lagrange_deriv(int N, const double *__restrict__ D,
                                               const double *__restrict__ L,
                                               double *__restrict__ dL) {
  for (int m = 0; m < N; ++m) {
    double s = 0.0;
    for (int k = 0; k < N; ++k) s = fma(L[k], D[k * N + m], s);
    dL[m] = s;
  }
}

__global__ void __launch_bounds__(kLocThreads)
    locate_kernel(int N, const double *__restrict__ x_phys, const double *__restrict__ centroids,
                  const double *__restrict__ nodes_g, const double *__restrict__ bw_g,
                  const double *__restrict__ D_g, double bx0, double by0, double inv_hx,
                  double inv_hy, int gx, int gy, const uint32_t *__restrict__ bin_ptr,
                  const uint32_t *__restrict__ bin_elems, int64_t n_points,
                  const double *__restrict__ px, const double *__restrict__ py, int it_max,
                  double tol, int64_t *__restrict__ elem_out, double *__restrict__ xi_out) {
  __shared__ double nodes[kMaxN], bw[kMaxN], D[kMaxN * kMaxN];
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    nodes[i] = nodes_g[i];
    bw[i] = bw_g[i];
  }
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) D[i] = D_g[i];
  __syncthreads();
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= n_points) return;
  const double X = px[q], Y = py[q];
  const int NN = N * N;
  int64_t found = -1;
  double fx0 = 0.0, fx1 = 0.0;
  const int bi = (int)floor((X - bx0) * inv_hx), bj = (int)floor((Y - by0) * inv_hy);
  if (bi >= 0 && bi < gx && bj >= 0 && bj < gy) {
    const uint32_t c0 = bin_ptr[bi * gy + bj], c1 = bin_ptr[bi * gy + bj + 1];
    // candidates in ascending (centroid distance, element id): repeated selection
    double last_d = -1.0;
    uint32_t last_e = 0;
    for (uint32_t tried = c0; tried < c1 && found < 0; ++tried) {
      double best_d = 0.0;
      uint32_t best_e = 0xffffffffu;
      for (uint32_t c = c0; c < c1; ++c) {
        const uint32_t e = bin_elems[c];
        const double dx = X - centroids[2 * (int64_t)e], dy = Y - centroids[2 * (int64_t)e + 1];
        const double d = dx * dx + dy * dy;
        const bool after = (d > last_d) || (d == last_d && last_d >= 0.0 && e > last_e);
        const bool first = last_d < 0.0;
        if (!(first || after)) continue;
        if (best_e == 0xffffffffu || d < best_d || (d == best_d && e < best_e)) {
          best_d = d;
          best_e = e;
        }
      }
      if (best_e == 0xffffffffu) break;
      last_d = best_d;
      last_e = best_e;
      // Newton inverse map in element best_e
      const double *Xe = x_phys + (int64_t)best_e * 2 * NN;
      double xi0 = 0.0, xi1 = 0.0;
      bool conv = false;
      for (int it = 0; it < it_max && !conv; ++it) {
        double L0[kMaxN], L1[kMaxN], d0[kMaxN], d1[kMaxN];
        lagrange_all(N, nodes, bw, xi0, L0);
        lagrange_all(N, nodes, bw, xi1, L1);
        lagrange_deriv(N, D, L0, d0);
        lagrange_deriv(N, D, L1, d1);
        double f0 = -X, f1 = -Y, j00 = 0.0, j01 = 0.0, j10 = 0.0, j11 = 0.0;
        for (int m = 0; m < N; ++m) {
          double ax = 0.0, ay = 0.0, bxs = 0.0, bys = 0.0;   // sums over n with L1 / d1
          for (int n = 0; n < N; ++n) {
            const double vx = Xe[m * N + n], vy = Xe[NN + m * N + n];
            ax = fma(vx, L1[n], ax);
            ay = fma(vy, L1[n], ay);
            bxs = fma(vx, d1[n], bxs);
            bys = fma(vy, d1[n], bys);
          }
          f0 = fma(L0[m], ax, f0);
          f1 = fma(L0[m], ay, f1);
          j00 = fma(d0[m], ax, j00);
          j10 = fma(d0[m], ay, j10);
          j01 = fma(L0[m], bxs, j01);
          j11 = fma(L0[m], bys, j11);
        }
        const double det = j00 * j11 - j01 * j10;
        const double dx0 = (-f0 * j11 + f1 * j01) / det;
        const double dx1 = (f0 * j10 - f1 * j00) / det;
        xi0 += dx0;
        xi1 += dx1;
        if (!(xi0 == xi0) || !(xi1 == xi1)) break;              // singular Jacobian: give up
        conv = sqrt(dx0 * dx0 + dx1 * dx1) <= tol;
      }
      // inside test of sem/mapping.py:175-177, with a rounding allowance: a point ON the mesh
      // boundary may come out as 1 + 2e-16 (the reference then raises OutsideDomain)
      if (conv && fabs(xi0) <= 1.0 + kInsideSlack && fabs(xi1) <= 1.0 + kInsideSlack) {
        found = (int64_t)best_e;
        fx0 = fmin(fmax(xi0, -1.0), 1.0);
        fx1 = fmin(fmax(xi1, -1.0), 1.0);
      }
    }
  }
  elem_out[q] = found;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  xi_out[q] = found >= 0 ? fx0 : nan;
  xi_out[n_points + q] = found >= 0 ? fx1 : nan;
}

__global__ void __launch_bounds__(kLocThreads)
    interp_points_kernel(int N, int64_t n_points, const int64_t *__restrict__ elem,
                         const double *__restrict__ xi, const uint32_t *__restrict__ l2g,
                         const double *__restrict__ nodes_g, const double *__restrict__ bw_g,
                         const double *__restrict__ coeffs, double *__restrict__ values) {
  __shared__ double nodes[kMaxN], bw[kMaxN];
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    nodes[i] = nodes_g[i];
    bw[i] = bw_g[i];
  }
  __syncthreads();
  const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (q >= n_points) return;
  const int64_t e = elem[q];
  if (e < 0) {
    values[q] = __longlong_as_double(0x7ff8000000000000LL);
    return;
  }
  double L0[kMaxN], L1[kMaxN];
  lagrange_all(N, nodes, bw, xi[q], L0);
  lagrange_all(N, nodes, bw, xi[n_points + q], L1);
  const uint32_t *ids = l2g + e * (int64_t)(N * N);
  double acc = 0.0;
  for (int m = 0; m < N; ++m) {
    double row = 0.0;
    for (int n = 0; n < N; ++n) row = fma(coeffs[ids[m * N + n]], L1[n], row);
    acc = fma(L0[m], row, acc);
  }
  values[q] = acc;
}

}  // namespace

extern "C" int semk_locate_points_f64(int n1, int64_t n_elem, const double *x_phys,
                                      const double *centroids, const double *gll_nodes,
                                      const double *bary_wts, const double *D, double bin_x0,
                                      double bin_y0, double bin_hx, double bin_hy, int bins_x,
                                      int bins_y, const uint32_t *bin_ptr,
                                      const uint32_t *bin_elems, int64_t n_points,
                                      const double *points, int it_max, double tol,
                                      int64_t *elem_out, double *xi_out, void *stream) {
  SEMK_REQUIRE(n1 >= 2 && n1 <= SEMK_MAX_N1, "semk_locate_points_f64: n1 outside [2, 17]");
  SEMK_REQUIRE(n_elem > 0 && x_phys && centroids && gll_nodes && bary_wts && D && bin_ptr &&
                   bin_elems && bins_x > 0 && bins_y > 0 && bin_hx > 0.0 && bin_hy > 0.0,
               "semk_locate_points_f64: bad argument");
  SEMK_REQUIRE(n_points >= 0 && it_max >= 1 && tol > 0.0, "semk_locate_points_f64: bad control");
  if (n_points == 0) return SEMK_OK;
  SEMK_REQUIRE(points && elem_out && xi_out, "semk_locate_points_f64: null pointer");
  const unsigned grid = (unsigned)((n_points + kLocThreads - 1) / kLocThreads);
  locate_kernel<<<grid, kLocThreads, 0, semk_stream(stream)>>>(
      n1, x_phys, centroids, gll_nodes, bary_wts, D, bin_x0, bin_y0, 1.0 / bin_hx, 1.0 / bin_hy,
      bins_x, bins_y, bin_ptr, bin_elems, n_points, points, points + n_points, it_max, tol,
      elem_out, xi_out);
  SEMK_LAUNCH_CHECK("locate_kernel");
  return SEMK_OK;
}

extern "C" int semk_interpolate_points_f64(int n1, int64_t n_points, const int64_t *elem,
                                           const double *xi, const uint32_t *l2g,
                                           const double *gll_nodes, const double *bary_wts,
                                           const double *coeffs, double *values, void *stream) {
  SEMK_REQUIRE(n1 >= 2 && n1 <= SEMK_MAX_N1, "semk_interpolate_points_f64: n1 outside [2, 17]");
  SEMK_REQUIRE(n_points >= 0, "semk_interpolate_points_f64: bad size");
  if (n_points == 0) return SEMK_OK;
  SEMK_REQUIRE(elem && xi && l2g && gll_nodes && bary_wts && coeffs && values,
               "semk_interpolate_points_f64: null pointer");
  const unsigned grid = (unsigned)((n_points + kLocThreads - 1) / kLocThreads);
  interp_points_kernel<<<grid, kLocThreads, 0, semk_stream(stream)>>>(
      n1, n_points, elem, xi, l2g, gll_nodes, bary_wts, coeffs, values);
  SEMK_LAUNCH_CHECK("interp_points_kernel");
  return SEMK_OK;
}
