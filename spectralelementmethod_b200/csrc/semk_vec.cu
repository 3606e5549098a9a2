// semk_vec.cu -- K4: fused vector kernels of Jacobi-PCG and the native driver.
//
// The reference solves the assembled (Schur-complement) system with SuperLU
// (`sparse.linalg.spsolve`, sem/discrete.py:511) plus per-element interior
// back-substitution (:513-524).  The engine replaces the solve on this path
// by a device-resident Jacobi-preconditioned CG on
//     Ahat = M A M + (I - M),   M = diag(free-DOF mask)
// (the Dirichlet elimination of sem/discrete.py:505-510 written as an SPD
// operator on the full vector).  All scalars stay on the device:
//     sc[0]=rz  sc[1]=pAp  sc[2]=rz_new  sc[3]=rr  sc[4]=bb
//     sc[5]=iterations done  sc[6]=converged flag  sc[7]=breakdown flag
// Kernels are HBM-bound streams; each vector is touched once per kernel and
// the dot products are fused into the pass that produces their operands.
// Reductions are deterministic: per-CTA partials + the last CTA to finish
// sums them in a fixed order (no floating-point atomics).
#include "semk_common.cuh"

namespace {

constexpr int kVecThreads = 256;
constexpr int kVecMaxBlocks = 148 * 8;

inline int vec_blocks(int64_t n) {
  const int64_t want = (n + kVecThreads * 4 - 1) / (kVecThreads * 4);
  return (int)(want < 1 ? 1 : (want < kVecMaxBlocks ? want : kVecMaxBlocks));
}

// partials layout: [kVecMaxBlocks][4] doubles, then one 64-bit arrival counter
__device__ __forceinline__ unsigned long long *counter_of(double *partials) {
  return reinterpret_cast<unsigned long long *>(partials + 4 * kVecMaxBlocks);
}

// Publish this CTA's NV partial sums; returns true (to all threads) in the CTA
// that arrives last, with `tot[0..NV)` holding the fixed-order totals in thread 0.
template <int NV>
__device__ __forceinline__ bool finish_reduction(double (&v)[NV], double *partials,
                                                 double (&tot)[NV]) {
  __shared__ double red[32];
  __shared__ bool is_last;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const double s = semk_block_sum(v[j], red);
    if (threadIdx.x == 0) partials[4 * blockIdx.x + j] = s;
  }
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long t = atomicAdd(counter_of(partials), 1ull);
    is_last = (t == (unsigned long long)gridDim.x - 1ull);
  }
  __syncthreads();
  if (!is_last) return false;
  __threadfence();
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    double s = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x)
      s += __ldcg(partials + 4 * b + j);
    tot[j] = semk_block_sum(s, red);
  }
  if (threadIdx.x == 0) *counter_of(partials) = 0ull;
  return true;
}

__global__ void __launch_bounds__(kVecThreads)
    pcg_init_kernel(int64_t n, int64_t n_dot, const double *__restrict__ b,
                    const double *__restrict__ Ax, const double *__restrict__ dinv,
                    const uint8_t *__restrict__ dirichlet, double *__restrict__ r,
                    double *__restrict__ p, double *__restrict__ sc,
                    double *__restrict__ partials) {
  double acc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool fixed = dirichlet && dirichlet[i];
    const double bi = fixed ? 0.0 : b[i];
    const double ri = fixed ? 0.0 : bi - Ax[i];
    const double zi = dinv[i] * ri;
    r[i] = ri;
    p[i] = zi;
    if (i < n_dot) {
      acc[0] = fma(ri, zi, acc[0]);
      acc[1] = fma(ri, ri, acc[1]);
      acc[2] = fma(bi, bi, acc[2]);
    }
  }
  double tot[3];
  if (finish_reduction<3>(acc, partials, tot) && threadIdx.x == 0) {
    sc[0] = tot[0];
    sc[1] = 0.0;
    sc[2] = tot[0];
    sc[3] = tot[1];
    sc[4] = tot[2];
    sc[5] = 0.0;
    sc[6] = 0.0;
    sc[7] = 0.0;
  }
}

// VEC: all vector pointers are 16-byte aligned -> 128-bit loads/stores, two pairs
// per thread and iteration in flight (these kernels are pure HBM streams).
template <bool VEC>
__global__ void __launch_bounds__(kVecThreads)
    pcg_update_xr_kernel(int64_t n, int64_t n_dot, const double *__restrict__ p,
                         const double *__restrict__ Ap, const double *__restrict__ dinv,
                         double *__restrict__ x, double *__restrict__ r, double *__restrict__ sc,
                         double *__restrict__ partials, double tol2) {
  const volatile double *vsc = sc;
  if (vsc[6] != 0.0 || vsc[7] != 0.0) return;  // converged / broken down: freeze
  const double rz = vsc[0], pAp = vsc[1];
  const bool ok = (pAp > 0.0) && (rz == rz);
  const double alpha = ok ? rz / pAp : 0.0;
  double acc[2] = {0.0, 0.0};
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  auto one = [&](int64_t i) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, Ap[i], r[i]);
    r[i] = ri;
    if (i < n_dot) {
      acc[0] = fma(ri * dinv[i], ri, acc[0]);
      acc[1] = fma(ri, ri, acc[1]);
    }
  };
  if (VEC) {
    const int64_t npair = n >> 1;
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    const double2 *Ap2 = reinterpret_cast<const double2 *>(Ap);
    const double2 *d2 = reinterpret_cast<const double2 *>(dinv);
    double2 *x2 = reinterpret_cast<double2 *>(x);
    double2 *r2 = reinterpret_cast<double2 *>(r);
    for (int64_t j0 = tid; j0 < npair; j0 += 2 * nthreads) {
      const int64_t j1 = j0 + nthreads;
      const bool two = j1 < npair;
      const double2 pa = p2[j0], aa = Ap2[j0], xa = x2[j0], ra = r2[j0], da = d2[j0];
      double2 pb = pa, ab = aa, xb = xa, rb = ra, db = da;
      if (two) {
        pb = p2[j1];
        ab = Ap2[j1];
        xb = x2[j1];
        rb = r2[j1];
        db = d2[j1];
      }
      double2 xo, ro;
      xo.x = fma(alpha, pa.x, xa.x);
      xo.y = fma(alpha, pa.y, xa.y);
      ro.x = fma(-alpha, aa.x, ra.x);
      ro.y = fma(-alpha, aa.y, ra.y);
      x2[j0] = xo;
      r2[j0] = ro;
      if (2 * j0 < n_dot) {
        acc[0] = fma(ro.x * da.x, ro.x, acc[0]);
        acc[1] = fma(ro.x, ro.x, acc[1]);
      }
      if (2 * j0 + 1 < n_dot) {
        acc[0] = fma(ro.y * da.y, ro.y, acc[0]);
        acc[1] = fma(ro.y, ro.y, acc[1]);
      }
      if (two) {
        xo.x = fma(alpha, pb.x, xb.x);
        xo.y = fma(alpha, pb.y, xb.y);
        ro.x = fma(-alpha, ab.x, rb.x);
        ro.y = fma(-alpha, ab.y, rb.y);
        x2[j1] = xo;
        r2[j1] = ro;
        if (2 * j1 < n_dot) {
          acc[0] = fma(ro.x * db.x, ro.x, acc[0]);
          acc[1] = fma(ro.x, ro.x, acc[1]);
        }
        if (2 * j1 + 1 < n_dot) {
          acc[0] = fma(ro.y * db.y, ro.y, acc[0]);
          acc[1] = fma(ro.y, ro.y, acc[1]);
        }
      }
    }
    if ((n & 1) && tid == 0) one(n - 1);
  } else {
    for (int64_t i = tid; i < n; i += nthreads) one(i);
  }
  double tot[2];
  if (finish_reduction<2>(acc, partials, tot) && threadIdx.x == 0) {
    sc[2] = tot[0];
    sc[3] = tot[1];
    sc[5] += 1.0;
    if (!ok) sc[7] = 1.0;
    if (tot[1] <= tol2 * sc[4]) sc[6] = 1.0;
  }
}

template <bool VEC>
__global__ void __launch_bounds__(kVecThreads)
    pcg_update_p_kernel(int64_t n, const double *__restrict__ r, const double *__restrict__ dinv,
                        double *__restrict__ p, double *__restrict__ sc,
                        double *__restrict__ partials) {
  const volatile double *vsc = sc;
  if (vsc[6] != 0.0 || vsc[7] != 0.0) return;
  const double rz_new = vsc[2], rz = vsc[0];
  const double beta = (rz != 0.0) ? rz_new / rz : 0.0;
  __shared__ bool is_last;
  __syncthreads();  // every thread of this CTA has read the scalars
  if (threadIdx.x == 0) {
    // every CTA has read sc[0] before it arrives; the last arrival rotates rz
    __threadfence();
    const unsigned long long t = atomicAdd(counter_of(partials), 1ull);
    is_last = (t == (unsigned long long)gridDim.x - 1ull);
  }
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {
    const int64_t npair = n >> 1;
    const double2 *r2 = reinterpret_cast<const double2 *>(r);
    const double2 *d2 = reinterpret_cast<const double2 *>(dinv);
    double2 *p2 = reinterpret_cast<double2 *>(p);
    for (int64_t j0 = tid; j0 < npair; j0 += 2 * nthreads) {
      const int64_t j1 = j0 + nthreads;
      const bool two = j1 < npair;
      const double2 ra = r2[j0], da = d2[j0], pa = p2[j0];
      double2 rb = ra, db = da, pb = pa;
      if (two) {
        rb = r2[j1];
        db = d2[j1];
        pb = p2[j1];
      }
      double2 o;
      o.x = fma(beta, pa.x, da.x * ra.x);
      o.y = fma(beta, pa.y, da.y * ra.y);
      p2[j0] = o;
      if (two) {
        o.x = fma(beta, pb.x, db.x * rb.x);
        o.y = fma(beta, pb.y, db.y * rb.y);
        p2[j1] = o;
      }
    }
    if ((n & 1) && tid == 0) p[n - 1] = fma(beta, p[n - 1], dinv[n - 1] * r[n - 1]);
  } else {
    for (int64_t i = tid; i < n; i += nthreads) p[i] = fma(beta, p[i], dinv[i] * r[i]);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    sc[0] = rz_new;
    *counter_of(partials) = 0ull;
  }
}

__global__ void __launch_bounds__(kVecThreads)
    dot_kernel(int64_t n, const double *__restrict__ a, const double *__restrict__ b,
               double *__restrict__ out, double *__restrict__ partials) {
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    acc[0] = fma(a[i], b[i], acc[0]);
  double tot[1];
  if (finish_reduction<1>(acc, partials, tot) && threadIdx.x == 0) out[0] = tot[0];
}

}  // namespace

extern "C" int64_t semk_vec_partials_len(int64_t n) {
  (void)n;
  return 4 * kVecMaxBlocks + 2;
}

extern "C" int semk_pcg_init_f64(int64_t n, int64_t n_dot, const double *b, const double *Ax,
                                 const double *dinv, const uint8_t *dirichlet, double *r,
                                 double *p, double *sc, double *partials, void *stream) {
  SEMK_REQUIRE(n > 0 && n_dot >= 0 && n_dot <= n, "semk_pcg_init_f64: bad sizes");
  SEMK_REQUIRE(b && Ax && dinv && r && p && sc && partials, "semk_pcg_init_f64: null pointer");
  pcg_init_kernel<<<vec_blocks(n), kVecThreads, 0, semk_stream(stream)>>>(
      n, n_dot, b, Ax, dinv, dirichlet, r, p, sc, partials);
  SEMK_LAUNCH_CHECK("pcg_init_kernel");
  return SEMK_OK;
}

static inline bool aligned16(const void *a, const void *b, const void *c, const void *d,
                             const void *e) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
           reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(d) |
           reinterpret_cast<uintptr_t>(e)) &
          15u) == 0;
}

static int update_xr(int64_t n, int64_t n_dot, const double *p, const double *Ap,
                     const double *dinv, double *x, double *r, double *sc, double *partials,
                     double tol2, cudaStream_t st) {
  if (aligned16(p, Ap, dinv, x, r))
    pcg_update_xr_kernel<true><<<vec_blocks(n), kVecThreads, 0, st>>>(n, n_dot, p, Ap, dinv, x, r,
                                                                     sc, partials, tol2);
  else
    pcg_update_xr_kernel<false><<<vec_blocks(n), kVecThreads, 0, st>>>(n, n_dot, p, Ap, dinv, x,
                                                                      r, sc, partials, tol2);
  SEMK_LAUNCH_CHECK("pcg_update_xr_kernel");
  return SEMK_OK;
}

extern "C" int semk_pcg_update_xr_f64(int64_t n, int64_t n_dot, const double *p, const double *Ap,
                                      const double *dinv, double *x, double *r, double *sc,
                                      double *partials, void *stream) {
  SEMK_REQUIRE(n > 0 && n_dot >= 0 && n_dot <= n, "semk_pcg_update_xr_f64: bad sizes");
  SEMK_REQUIRE(p && Ap && dinv && x && r && sc && partials,
               "semk_pcg_update_xr_f64: null pointer");
  // tol2 < 0: the device-side convergence freeze is disabled (caller decides)
  return update_xr(n, n_dot, p, Ap, dinv, x, r, sc, partials, -1.0, semk_stream(stream));
}

extern "C" int semk_pcg_update_p_f64(int64_t n, const double *r, const double *dinv, double *p,
                                      double *sc, double *partials, void *stream) {
  SEMK_REQUIRE(n > 0, "semk_pcg_update_p_f64: bad size");
  SEMK_REQUIRE(r && dinv && p && sc && partials, "semk_pcg_update_p_f64: null pointer");
  if (aligned16(r, dinv, p, p, p))
    pcg_update_p_kernel<true><<<vec_blocks(n), kVecThreads, 0, semk_stream(stream)>>>(
        n, r, dinv, p, sc, partials);
  else
    pcg_update_p_kernel<false><<<vec_blocks(n), kVecThreads, 0, semk_stream(stream)>>>(
        n, r, dinv, p, sc, partials);
  SEMK_LAUNCH_CHECK("pcg_update_p_kernel");
  return SEMK_OK;
}

extern "C" int semk_dot_f64(int64_t n, const double *a, const double *b, double *out,
                            double *partials, void *stream) {
  SEMK_REQUIRE(n > 0 && a && b && out && partials, "semk_dot_f64: bad argument");
  dot_kernel<<<vec_blocks(n), kVecThreads, 0, semk_stream(stream)>>>(n, a, b, out, partials);
  SEMK_LAUNCH_CHECK("dot_kernel");
  return SEMK_OK;
}

extern "C" int semk_pcg_solve_f64(const semk_op *op, const double *b, double *x,
                                  const double *dinv, double *work, double *sc,
                                  double *vec_partials, double rtol, int maxiter, int check_every,
                                  semk_pcg_info *info, void *stream) {
  SEMK_REQUIRE(op && b && x && dinv && work && sc && vec_partials && info,
               "semk_pcg_solve_f64: null pointer");
  SEMK_REQUIRE(maxiter >= 0 && check_every >= 1 && rtol >= 0.0, "semk_pcg_solve_f64: bad control");
  cudaStream_t st = semk_stream(stream);
  const int64_t n = op->n_nodes;
  const int64_t n_pad = (n + 31) & ~(int64_t)31;  // keeps the sub-vectors 16-byte aligned
  double *r = work, *p = work + n_pad, *Ap = work + 2 * n_pad;
  const int flags = SEMK_MASK_IN | SEMK_MASK_OUT | SEMK_DIRICHLET_IDENTITY;
  const double tol2 = rtol * rtol;
  static double *h_sc = nullptr;  // pinned landing zone for the 64-byte scalar block
  if (!h_sc) SEMK_CUDA_CHECK(cudaMallocHost(&h_sc, 8 * sizeof(double)));

  int rc = semk_poisson_apply_f64(op, x, Ap, flags, nullptr, st);
  if (rc != SEMK_OK) return rc;
  rc = semk_pcg_init_f64(n, n, b, Ap, dinv, op->dirichlet, r, p, sc, vec_partials, st);
  if (rc != SEMK_OK) return rc;

  auto poll = [&]() -> int {
    SEMK_CUDA_CHECK(cudaMemcpyAsync(h_sc, sc, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    SEMK_CUDA_CHECK(cudaStreamSynchronize(st));
    return SEMK_OK;
  };
  rc = poll();
  if (rc != SEMK_OK) return rc;
  info->bnorm = sqrt(h_sc[4]);
  info->iterations = 0;
  info->status = 0;
  if (h_sc[4] == 0.0 || h_sc[3] <= tol2 * h_sc[4]) {
    info->rel_residual = h_sc[4] > 0.0 ? sqrt(h_sc[3] / h_sc[4]) : 0.0;
    return SEMK_OK;
  }

  auto one_iteration = [&]() -> int {
    int e = semk_poisson_apply_f64(op, p, Ap, flags, sc + 1, st);
    if (e != SEMK_OK) return e;
    e = update_xr(n, n, p, Ap, dinv, x, r, sc, vec_partials, tol2, st);
    if (e != SEMK_OK) return e;
    return semk_pcg_update_p_f64(n, r, dinv, p, sc, vec_partials, st);
  };

  // Capture `check_every` iterations once and replay: removes per-launch host
  // cost from the loop.  Falls back to eager launches if capture is refused.
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool use_graph = false;
  if (check_every > 1 && st != nullptr) {
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int e = SEMK_OK;
      for (int k = 0; k < check_every && e == SEMK_OK; ++k) e = one_iteration();
      cudaError_t ce = cudaStreamEndCapture(st, &graph);
      if (e == SEMK_OK && ce == cudaSuccess && graph &&
          cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess)
        use_graph = true;
      else
        (void)cudaGetLastError();
    } else {
      (void)cudaGetLastError();
    }
  }

  int status = 1;  // maxiter unless proven otherwise
  int launched = 0;
  rc = SEMK_OK;
  while (launched < maxiter) {
    if (use_graph) {
      if (cudaGraphLaunch(exec, st) != cudaSuccess) {
        semk_set_error("semk_pcg_solve_f64: cudaGraphLaunch failed");
        rc = SEMK_ERR_CUDA;
        break;
      }
      launched += check_every;
    } else {
      for (int k = 0; k < check_every && launched < maxiter && rc == SEMK_OK; ++k, ++launched)
        rc = one_iteration();
      if (rc != SEMK_OK) break;
    }
    rc = poll();
    if (rc != SEMK_OK) break;
    if (h_sc[7] != 0.0) {
      status = SEMK_ERR_BREAKDOWN;
      break;
    }
    if (h_sc[6] != 0.0) {
      status = 0;
      break;
    }
  }
  if (exec) cudaGraphExecDestroy(exec);
  if (graph) cudaGraphDestroy(graph);
  if (rc != SEMK_OK) return rc;
  info->iterations = (int32_t)h_sc[5];
  info->status = status;
  info->rel_residual = sqrt(h_sc[3] / h_sc[4]);
  if (status == SEMK_ERR_BREAKDOWN) {
    semk_set_error("semk_pcg_solve_f64: breakdown (p.Ap <= 0 or non-finite)");
    return SEMK_ERR_BREAKDOWN;
  }
  return SEMK_OK;
}
