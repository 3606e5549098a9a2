// semk_ml.cu -- multilevel-preconditioned CG on the condensed system, one driver for one
// GPU and for a strip partition, and the small-vector all-reduce over NVLink peer memory
// it needs between ranks.
//
// What it replaces: the reference solves the condensed (Schur-complement) system over the
// element-exterior DOFs directly, `spsolve` at sem/discrete.py:502-511, then back-
// substitutes the interiors (:513-524).  Here the same system is solved iteratively:
//     outer : (flexible) CG on Shat x = b
//     M^-1 r = dinv r + P xc,   xc ~= Ac^-1 P^T r        vertex coarse space, inner PCG
//     inner preconditioner = dinv_c (+ P2 A3inv P2^T, aggregation with a dense inverse)
// Design (B200-first): every inner iteration is a handful of small kernels whose scalars
// (alpha, beta, convergence / breakdown flags) never leave the device.  The scalar logic
// lives in single-CTA "scalar step" kernels which also carry the cross-rank all-reduce:
// push to every peer's region over NVLink, publish an epoch, wait, sum in rank order.
// Vector kernels only READ scalars and write per-CTA partial sums, so there is no
// read/update race on the scalar block and the host only polls it every few iterations.
#include "semk_common.cuh"

#include <cstdlib>

namespace {

constexpr int kT = 256;                 // vector kernels
constexpr int kMaxBlocks = kSemkRedMaxBlocks;
constexpr int kStepThreads = 1024;      // scalar-step / all-reduce kernel (one CTA)
constexpr long long kSpinLimit = 4000000000LL;  // ~2 s of SM clock

inline int blocks_for(int64_t n) {
  const int64_t want = (n + kT * 4 - 1) / (kT * 4);
  return (int)(want < 1 ? 1 : (want < kMaxBlocks ? want : kMaxBlocks));
}

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- scalar block of one PCG level (device doubles) ------------------------------------
enum {
  S_RZ = 0,      // r.z of the current direction
  S_PAP = 1,     // p.Ap (this rank's share until the ALPHA step has run)
  S_RZN = 2,     // new r.z          } contiguous: reduced together
  S_ZAP = 3,     // z.Ap (flexible)  }
  S_RR = 4,      // r.r
  S_BB = 5,      // b.b
  S_ALPHA = 6,
  S_BETA = 7,
  S_ITER = 8,
  S_CONV = 9,    // 1 = converged: x is final, later kernels of the chunk do nothing
  S_BREAK = 10,  // 1 = p.Ap <= 0 or non-finite
  S_TOL2 = 11,
  S_LEN = 16
};

enum {
  OP_NONE = 0,
  OP_ALPHA = 1,
  OP_RR = 2,
  OP_RR_FIRST = 3,
  OP_BETA = 4,
  OP_BETA_FIRST = 5,
  OP_RR2 = 6,            // inner level: buf ends with { r.r, r.(dinv r) }
  OP_RR2_FIRST = 7,
  OP_BETA_TOP = 8,       // inner level: r.z = r.(dinv r) + y3.r3, then beta
  OP_BETA_TOP_FIRST = 9,
  OP_RR2_KEEP = 10,      // as OP_RR2 for an initial residual: b.b and the count are kept
  // no aggregation level between the two: OP_RR2* immediately followed by OP_BETA_TOP*
  OP_RR2_BETA = 11,
  OP_RR2_BETA_FIRST = 12,
  OP_RR2_BETA_KEEP = 13
};

struct CommDev {
  int rank, world;
  long long capacity;
  char *regions[SEMK_COMM_MAX_WORLD];
  int *status;
};

__device__ __forceinline__ unsigned long long *region_epoch(char *r) {
  return reinterpret_cast<unsigned long long *>(r);
}
__device__ __forceinline__ unsigned long long *region_flags(char *r) {
  return reinterpret_cast<unsigned long long *>(r + 64);
}
__device__ __forceinline__ double *region_recv(char *r, int parity, int src, long long cap,
                                               int world) {
  return reinterpret_cast<double *>(r + 256) + ((long long)parity * world + src) * cap;
}

// All-reduce buf[0, n) over the ranks (one CTA; all threads call).  Returns false after a
// time-out (status word raised); buf is then left unreduced.
__device__ bool comm_allreduce(const CommDev &c, double *buf, int n) {
  if (c.world <= 1) return true;
  __shared__ int timed_out;
  char *mine = c.regions[c.rank];
  const unsigned long long epoch = *region_epoch(mine) + 1ull;   // same value for all threads
  const int par = (int)(epoch & 1ull);
  __syncthreads();                                               // everyone has read the epoch
  for (int w = 0; w < c.world; ++w) {
    double *dst = region_recv(c.regions[w], par, c.rank, c.capacity, c.world);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  if ((int)threadIdx.x < c.world) {
    st_release_sys_u64(region_flags(c.regions[threadIdx.x]) + c.rank, epoch);
    const unsigned long long *f = region_flags(mine) + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(f) < epoch) {
      if (clock64() - t0 > kSpinLimit) {
        timed_out = 1;
        break;
      }
      __nanosleep(32);
    }
  }
  __syncthreads();
  if (timed_out) {
    if (threadIdx.x == 0) {
      atomicExch(c.status, 1);
      *region_epoch(mine) = epoch;
    }
    return false;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int w = 0; w < c.world; ++w)
      s += __ldcg(region_recv(mine, par, w, c.capacity, c.world) + i);   // fixed rank order
    buf[i] = s;
  }
  if (threadIdx.x == 0) *region_epoch(mine) = epoch;
  __syncthreads();
  return true;
}

// One scalar step: all-reduce `buf[0, n)` in place, then update the scalar block.
//   OP_ALPHA      buf = s + S_PAP (1): alpha = rz / pAp, breakdown test
//   OP_RR[_FIRST] buf ends with r.r at buf[n - 1]: s[S_RR] = it; FIRST also sets b.b = r.r
//                 (inner solves start from x = 0, so r = b); convergence test
//                 (and ++iterations unless FIRST)
//   OP_BETA[_FIRST] buf = s + S_RZN (2): beta (flexible or standard), rz <- rz_new
//   OP_RR2[_FIRST] inner level, buf = [r3 (na) | r.r | r.(dinv r)]: as OP_RR, and the Jacobi
//                 part of r.z is parked in s[S_RZN]
//   OP_BETA_TOP[_FIRST] inner level, no reduction: r.z = s[S_RZN] + y3 . r3 -- the aggregation
//                 correction z += P2 y3 contributes (P2 y3).r = y3.(P2^T r) = y3.r3, both
//                 replicated vectors, so z never has to be formed for the dot -- then beta
__global__ void __launch_bounds__(kStepThreads)
    ml_step_kernel(CommDev c, double *__restrict__ buf, int n, int op, int flexible,
                   double *__restrict__ s, const double *__restrict__ y3,
                   const double *__restrict__ r3v, int na) {
  const bool ok = n > 0 ? comm_allreduce(c, buf, n) : true;
  double top_dot = 0.0;
  if (op == OP_BETA_TOP || op == OP_BETA_TOP_FIRST) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < na; i += blockDim.x) acc = fma(y3[i], r3v[i], acc);
    top_dot = semk_block_sum(acc, red);
  }
  if (threadIdx.x != 0) return;
  if (!ok) {
    s[S_BREAK] = 2.0;
    return;
  }
  // converged in the previous iteration: its direction kernel has applied the last x
  // update, from now on everything is frozen (inner / partitioned Jacobi loops)
  if (op == OP_ALPHA && s[S_CONV] == 1.0) s[S_CONV] = 2.0;
  if (s[S_CONV] != 0.0 || s[S_BREAK] != 0.0) return;   // frozen
  switch (op) {
    case OP_ALPHA: {
      const double pAp = s[S_PAP], rz = s[S_RZ];
      if (!(pAp > 0.0) || !(rz == rz)) {
        s[S_BREAK] = 1.0;
        s[S_ALPHA] = 0.0;
      } else {
        s[S_ALPHA] = rz / pAp;
      }
      break;
    }
    case OP_RR:
    case OP_RR_FIRST: {
      const double rr = buf[n - 1];
      s[S_RR] = rr;
      if (op == OP_RR_FIRST)
        s[S_BB] = rr;
      else
        s[S_ITER] += 1.0;   // one more completed update of x
      if (rr <= s[S_TOL2] * s[S_BB]) s[S_CONV] = 1.0;
      break;
    }
    case OP_RR2:
    case OP_RR2_FIRST:
    case OP_RR2_KEEP:
    case OP_RR2_BETA:
    case OP_RR2_BETA_FIRST:
    case OP_RR2_BETA_KEEP: {
      const double rr = buf[n - 2];
      const double rzn = buf[n - 1];
      s[S_RR] = rr;
      s[S_RZN] = rzn;
      if (op == OP_RR2_FIRST || op == OP_RR2_BETA_FIRST)
        s[S_BB] = rr;
      else if (op == OP_RR2 || op == OP_RR2_BETA)
        s[S_ITER] += 1.0;
      if (rr <= s[S_TOL2] * s[S_BB]) {
        s[S_CONV] = 1.0;
      } else if (op >= OP_RR2_BETA) {
        const double rz = s[S_RZ];
        s[S_BETA] = (op != OP_RR2_BETA || rz == 0.0) ? 0.0 : rzn / rz;
        s[S_RZ] = rzn;
      }
      break;
    }
    case OP_BETA_TOP:
    case OP_BETA_TOP_FIRST: {
      const double rzn = s[S_RZN] + top_dot;
      const double rz = s[S_RZ];
      s[S_BETA] = (op == OP_BETA_TOP_FIRST || rz == 0.0) ? 0.0 : rzn / rz;
      s[S_RZ] = rzn;
      break;
    }
    case OP_BETA:
    case OP_BETA_FIRST: {
      const double rzn = s[S_RZN];
      if (op == OP_BETA_FIRST) {
        s[S_BETA] = 0.0;
      } else {
        const double rz = s[S_RZ];
        s[S_BETA] = (rz != 0.0) ? (flexible ? -s[S_ALPHA] * s[S_ZAP] / rz : rzn / rz) : 0.0;
      }
      s[S_RZ] = rzn;
      break;
    }
    default:
      break;
  }
}

__global__ void __launch_bounds__(kStepThreads)
    allreduce_kernel(CommDev c, double *__restrict__ buf, int n) {
  (void)comm_allreduce(c, buf, n);
}

// x += alpha p ; r -= alpha Ap ; z = dinv r ; rr_out = sum_{i < n_dot} r_i^2 (this rank)
__global__ void __launch_bounds__(kT)
    ml_update_kernel(int64_t n, int64_t n_dot, const double *__restrict__ p,
                     const double *__restrict__ Ap, const double *__restrict__ dinv,
                     double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
                     const double *__restrict__ s, double *__restrict__ rr_out,
                     double *__restrict__ partials) {
  const bool frozen = s[S_CONV] != 0.0 || s[S_BREAK] != 0.0;
  const double alpha = s[S_ALPHA];
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double ri = r[i];
    if (!frozen) {
      x[i] = fma(alpha, p[i], x[i]);
      ri = fma(-alpha, Ap[i], ri);
      r[i] = ri;
    }
    z[i] = dinv[i] * ri;
    if (i < n_dot) acc[0] = fma(ri, ri, acc[0]);
  }
  double tot[1];
  if (semk_finish_reduction<1>(acc, partials, tot) && threadIdx.x == 0) rr_out[0] = tot[0];
}

// z = dinv r ; rr_out = sum_{i < n_dot} r_i^2   (first step of a solve: no direction yet)
__global__ void __launch_bounds__(kT)
    ml_first_kernel(int64_t n, int64_t n_dot, const double *__restrict__ dinv,
                    const double *__restrict__ r, double *__restrict__ z,
                    double *__restrict__ rr_out, double *__restrict__ partials) {
  double acc[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double ri = r[i];
    z[i] = dinv[i] * ri;
    if (i < n_dot) acc[0] = fma(ri, ri, acc[0]);
  }
  double tot[1];
  if (semk_finish_reduction<1>(acc, partials, tot) && threadIdx.x == 0) rr_out[0] = tot[0];
}

// z += P xc (at most two vertices per fine node) ; out2 = { r.z, z.Ap } over the owned prefix
__global__ void __launch_bounds__(kT)
    ml_prolong_dot_kernel(int64_t n, int64_t n_dot, const uint32_t *__restrict__ pv,
                          const double *__restrict__ pw, const double *__restrict__ xc,
                          const double *__restrict__ r, const double *__restrict__ Ap,
                          double *__restrict__ z, double *__restrict__ out2,
                          double *__restrict__ partials) {
  double acc[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double wa = pw[2 * i], wb = pw[2 * i + 1];
    double zi = z[i];
    if (wa != 0.0) zi = fma(wa, xc[pv[2 * i]], zi);
    if (wb != 0.0) zi = fma(wb, xc[pv[2 * i + 1]], zi);
    z[i] = zi;
    if (i < n_dot) {
      acc[0] = fma(r[i], zi, acc[0]);
      if (Ap) acc[1] = fma(zi, Ap[i], acc[1]);
    }
  }
  double tot[2];
  if (semk_finish_reduction<2>(acc, partials, tot) && threadIdx.x == 0) {
    out2[0] = tot[0];
    out2[1] = tot[1];
  }
}

// ---- inner (vertex) level: the preconditioned residual z = dinv r + P2 y3 is never stored ----
// r -= alpha Ap (unless first / frozen) ; out2 = { r.r, r.(dinv r) } (owned).  x += alpha p
// rides with the direction update below (that kernel reads p anyway): 4 + 6 vector passes.
// VEC: every pointer is 16-byte aligned -> 128-bit accesses, two pairs in flight per thread
// (these are pure HBM streams on the fine level of the partitioned Jacobi-PCG).
template <bool VEC>
__global__ void __launch_bounds__(kT)
    ml_inner_update_kernel(int64_t n, int64_t n_dot, int first, const double *__restrict__ Ap,
                           const double *__restrict__ dinv, double *__restrict__ r,
                           const double *__restrict__ s, double *__restrict__ out2,
                           double *__restrict__ partials) {
  const bool move = !first && s[S_CONV] == 0.0 && s[S_BREAK] == 0.0;
  const double alpha = s[S_ALPHA];
  double acc[2] = {0.0, 0.0};
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  auto one = [&](int64_t i) {
    double ri = r[i];
    if (move) {
      ri = fma(-alpha, Ap[i], ri);
      r[i] = ri;
    }
    if (i < n_dot) {
      acc[0] = fma(ri, ri, acc[0]);
      acc[1] = fma(ri * dinv[i], ri, acc[1]);
    }
  };
  if (VEC) {
    const int64_t npair = n >> 1;
    const double2 *Ap2 = reinterpret_cast<const double2 *>(Ap);
    const double2 *d2 = reinterpret_cast<const double2 *>(dinv);
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 zero2 = make_double2(0.0, 0.0);
    for (int64_t j0 = tid; j0 < npair; j0 += 2 * nthreads) {
      const int64_t jj[2] = {j0, j0 + nthreads};
      const bool on[2] = {true, jj[1] < npair};
      double2 av[2], rv[2], dv[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        rv[q] = on[q] ? r2[jj[q]] : zero2;
        dv[q] = on[q] ? d2[jj[q]] : zero2;
        av[q] = (on[q] && move) ? Ap2[jj[q]] : zero2;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (!on[q]) continue;
        double2 ro = rv[q];
        if (move) {
          ro.x = fma(-alpha, av[q].x, rv[q].x);
          ro.y = fma(-alpha, av[q].y, rv[q].y);
          r2[jj[q]] = ro;
        }
        if (2 * jj[q] < n_dot) {
          acc[0] = fma(ro.x, ro.x, acc[0]);
          acc[1] = fma(ro.x * dv[q].x, ro.x, acc[1]);
        }
        if (2 * jj[q] + 1 < n_dot) {
          acc[0] = fma(ro.y, ro.y, acc[0]);
          acc[1] = fma(ro.y * dv[q].y, ro.y, acc[1]);
        }
      }
    }
    if ((n & 1) && tid == 0) one(n - 1);
  } else {
    for (int64_t i = tid; i < n; i += nthreads) one(i);
  }
  double tot[2];
  if (semk_finish_reduction<2>(acc, partials, tot) && threadIdx.x == 0) {
    out2[0] = tot[0];
    out2[1] = tot[1];
  }
}

// x += alpha p_old (unless first) ; p = dinv r + y3[agg] + beta p  (agg == NULL: no
// aggregation level).  On the iteration that converged (S_CONV == 1, set by the step kernel
// in between) only the x update is done -- x then is the converged iterate; the next ALPHA
// step moves the state on to 2 = frozen, and this kernel does nothing any more.
template <bool VEC>
__global__ void __launch_bounds__(kT)
    ml_inner_direction_kernel(int64_t n, int first, const double *__restrict__ dinv,
                              const double *__restrict__ r, const uint32_t *__restrict__ agg,
                              const double *__restrict__ y3, double *__restrict__ p,
                              double *__restrict__ x, const double *__restrict__ s) {
  const double state = s[S_CONV];
  if (s[S_BREAK] != 0.0 || state == 2.0) return;
  const bool move_p = state == 0.0;
  const bool move_x = !first;
  const double beta = s[S_BETA], alpha = s[S_ALPHA];
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  auto top = [&](int64_t i) -> double {
    if (!agg) return 0.0;
    const uint32_t a = agg[i];
    return a != 0xffffffffu ? y3[a] : 0.0;
  };
  auto one = [&](int64_t i) {
    const double pi = p[i];
    if (move_x) x[i] = fma(alpha, pi, x[i]);
    if (move_p) p[i] = fma(beta, pi, fma(dinv[i], r[i], top(i)));
  };
  if (VEC) {
    const int64_t npair = n >> 1;
    const double2 *d2 = reinterpret_cast<const double2 *>(dinv);
    const double2 *r2 = reinterpret_cast<const double2 *>(r);
    double2 *p2 = reinterpret_cast<double2 *>(p);
    double2 *x2 = reinterpret_cast<double2 *>(x);
    const double2 zero2 = make_double2(0.0, 0.0);
    for (int64_t j0 = tid; j0 < npair; j0 += 2 * nthreads) {
      const int64_t jj[2] = {j0, j0 + nthreads};
      const bool on[2] = {true, jj[1] < npair};
      double2 dv[2], rv[2], pv[2], xv[2];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        pv[q] = on[q] ? p2[jj[q]] : zero2;
        dv[q] = (on[q] && move_p) ? d2[jj[q]] : zero2;
        rv[q] = (on[q] && move_p) ? r2[jj[q]] : zero2;
        xv[q] = (on[q] && move_x) ? x2[jj[q]] : zero2;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (!on[q]) continue;
        if (move_x) {
          double2 xo;
          xo.x = fma(alpha, pv[q].x, xv[q].x);
          xo.y = fma(alpha, pv[q].y, xv[q].y);
          x2[jj[q]] = xo;
        }
        if (move_p) {
          double2 o;
          o.x = fma(beta, pv[q].x, fma(dv[q].x, rv[q].x, top(2 * jj[q])));
          o.y = fma(beta, pv[q].y, fma(dv[q].y, rv[q].y, top(2 * jj[q] + 1)));
          p2[jj[q]] = o;
        }
      }
    }
    if ((n & 1) && tid == 0) one(n - 1);
  } else {
    for (int64_t i = tid; i < n; i += nthreads) one(i);
  }
}

inline bool ml_aligned16(const void *a, const void *b, const void *c, const void *d,
                         const void *e) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
           reinterpret_cast<uintptr_t>(c) | reinterpret_cast<uintptr_t>(d) |
           reinterpret_cast<uintptr_t>(e)) &
          15u) == 0;
}

// y = Ac x with Ac in ELL format (column major: entry k of row v at [k * n + v]) ;
// dot_out = x . y over all local rows.  One thread per row: coalesced matrix reads.
__global__ void __launch_bounds__(kT)
    ml_ell_spmv_kernel(int64_t n, int width, const uint32_t *__restrict__ cols,
                       const double *__restrict__ vals, const double *__restrict__ x,
                       double *__restrict__ y, double *__restrict__ dot_out,
                       double *__restrict__ partials) {
  double acc[1] = {0.0};
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n;
       v += (int64_t)gridDim.x * blockDim.x) {
    double sum = 0.0;
    for (int k = 0; k < width; ++k)
      sum = fma(vals[(int64_t)k * n + v], x[cols[(int64_t)k * n + v]], sum);
    y[v] = sum;
    acc[0] = fma(x[v], sum, acc[0]);
  }
  if (dot_out) {
    double tot[1];
    if (semk_finish_reduction<1>(acc, partials, tot) && threadIdx.x == 0) dot_out[0] = tot[0];
  }
}

// Assemble the vertex coarse operator from the element matrices Ace into ELL rows: one
// thread per vertex walks its element entries in ascending order (fixed summation order).
// Row layout: diagonal first, then columns in order of first appearance, padded with
// (v, 0).  Dirichlet vertices get the identity row.
__global__ void __launch_bounds__(kT)
    ml_ell_build_kernel(int64_t n_v, int width, const uint32_t *__restrict__ vptr,
                        const uint32_t *__restrict__ vpos, const uint32_t *__restrict__ vert_c,
                        const double *__restrict__ Ace, const uint8_t *__restrict__ dirichlet_c,
                        uint32_t *__restrict__ cols, double *__restrict__ vals,
                        int *__restrict__ overflow) {
  const int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (v >= n_v) return;
  uint32_t lc[32];
  double lv[32];
  int cnt = 1;
  lc[0] = (uint32_t)v;
  lv[0] = 0.0;
  for (uint32_t q = vptr[v]; q < vptr[v + 1]; ++q) {
    const uint32_t ent = vpos[q];
    const uint32_t e = ent >> 2, i = ent & 3u;
    for (int j = 0; j < 4; ++j) {
      const uint32_t cidx = vert_c[e * 4 + j];
      const double a = Ace[(int64_t)e * 16 + i * 4 + j];
      int k = 0;
      while (k < cnt && lc[k] != cidx) ++k;
      if (k == cnt) {
        if (cnt >= width) {
          atomicExch(overflow, 1);
          continue;
        }
        lc[cnt] = cidx;
        lv[cnt] = 0.0;
        ++cnt;
      }
      lv[k] += a;
    }
  }
  if (dirichlet_c && dirichlet_c[v]) lv[0] = 1.0;
  for (int k = 0; k < width; ++k) {
    cols[(int64_t)k * n_v + v] = k < cnt ? lc[k] : (uint32_t)v;
    vals[(int64_t)k * n_v + v] = k < cnt ? lv[k] : 0.0;
  }
}

// y = A x, A dense [n][n] row major in FP32 (the top-level inverse only enters the
// preconditioner; FP64 accumulation): one warp per row
__global__ void __launch_bounds__(kT)
    ml_dense_matvec_f32_kernel(int64_t n, const float *__restrict__ A,
                               const double *__restrict__ x, double *__restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const float *row = A + i * n;
    double s0 = 0.0, s1 = 0.0;
    if ((n & 3) == 0) {
      const float4 *row4 = reinterpret_cast<const float4 *>(row);
      const double2 *x2 = reinterpret_cast<const double2 *>(x);
      for (int64_t j = lane; j < (n >> 2); j += 32) {
        const float4 a = row4[j];
        const double2 b0 = x2[2 * j], b1 = x2[2 * j + 1];
        s0 = fma((double)a.x, b0.x, s0);
        s1 = fma((double)a.y, b0.y, s1);
        s0 = fma((double)a.z, b1.x, s0);
        s1 = fma((double)a.w, b1.y, s1);
      }
    } else {
      for (int64_t j = lane; j < n; j += 32) s0 = fma((double)row[j], x[j], s0);
    }
    double sum = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if (lane == 0) y[i] = sum;
  }
}

// p = z + beta p  (frozen: nothing)
__global__ void __launch_bounds__(kT)
    ml_direction_kernel(int64_t n, const double *__restrict__ z, double *__restrict__ p,
                        const double *__restrict__ s) {
  if (s[S_CONV] != 0.0 || s[S_BREAK] != 0.0) return;
  const double beta = s[S_BETA];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = fma(beta, p[i], z[i]);
}

// rc = P^T r over the owned fine nodes: one thread per coarse row, fixed order
__global__ void __launch_bounds__(kT)
    ml_restrict_kernel(int64_t n_v, const uint32_t *__restrict__ rptr,
                       const uint32_t *__restrict__ ridx, const double *__restrict__ rw,
                       const double *__restrict__ r, int64_t n_owned, double *__restrict__ rc) {
  for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_v;
       v += (int64_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (uint32_t q = rptr[v]; q < rptr[v + 1]; ++q) {
      const uint32_t g = ridx[q];
      if ((int64_t)g < n_owned) s = fma(rw[q], r[g], s);
    }
    rc[v] = s;
  }
}

// r3[a] = sum of q over the (owned) vertices of aggregate a: one warp per aggregate
__global__ void __launch_bounds__(kT)
    ml_agg_restrict_kernel(int64_t n_agg, const uint32_t *__restrict__ aptr,
                           const uint32_t *__restrict__ aidx, const double *__restrict__ q,
                           double *__restrict__ r3) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t a = warp; a < n_agg; a += nwarps) {
    double s = 0.0;
    for (uint32_t k = aptr[a] + lane; k < aptr[a + 1]; k += 32) s += q[aidx[k]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) r3[a] = s;
  }
}

// y = A x, A dense [n][n] row major: one warp per row, 128-bit loads when n is even
__global__ void __launch_bounds__(kT)
    ml_dense_matvec_kernel(int64_t n, const double *__restrict__ A, const double *__restrict__ x,
                           double *__restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < n; i += nwarps) {
    const double *row = A + i * n;
    double s0 = 0.0, s1 = 0.0;
    if ((n & 1) == 0) {
      const double2 *row2 = reinterpret_cast<const double2 *>(row);
      const double2 *x2 = reinterpret_cast<const double2 *>(x);
      for (int64_t j = lane; j < (n >> 1); j += 32) {
        const double2 a = row2[j], b = x2[j];
        s0 = fma(a.x, b.x, s0);
        s1 = fma(a.y, b.y, s1);
      }
    } else {
      for (int64_t j = lane; j < n; j += 32) s0 = fma(row[j], x[j], s0);
    }
    double s = s0 + s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) y[i] = s;
  }
}

__global__ void __launch_bounds__(kT)
    ml_resid_kernel(int64_t n, const double *__restrict__ b, const double *__restrict__ Ax,
                    const uint8_t *__restrict__ dirichlet, double *__restrict__ r,
                    int64_t n_dot, double *__restrict__ out2, double *__restrict__ partials) {
  // r = b - Ax (0 on Dirichlet rows) ; out2 = { r.r, b.b } over the owned prefix (masked b)
  double acc[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool fixed = dirichlet && dirichlet[i];
    const double bi = fixed ? 0.0 : b[i];
    const double ri = fixed ? 0.0 : bi - Ax[i];
    r[i] = ri;
    if (i < n_dot) {
      acc[0] = fma(ri, ri, acc[0]);
      acc[1] = fma(bi, bi, acc[1]);
    }
  }
  double tot[2];
  if (semk_finish_reduction<2>(acc, partials, tot) && threadIdx.x == 0) {
    out2[0] = tot[0];
    out2[1] = tot[1];
  }
}

// A3 = P2^T Ac P2 from the element coarse matrices, one thread per row: the thread walks
// the (local) vertices of its aggregate, their element entries and the four columns of
// each -- a fixed order and no atomics, so the result is bit-reproducible.
__global__ void __launch_bounds__(kT)
    ml_top_assemble_kernel(int64_t n_agg, const uint32_t *__restrict__ agg,
                           const uint32_t *__restrict__ aptr, const uint32_t *__restrict__ aidx,
                           const uint32_t *__restrict__ vptr, const uint32_t *__restrict__ vpos,
                           const uint32_t *__restrict__ vert_c, const double *__restrict__ Ace,
                           double *__restrict__ A3) {
  const int64_t a = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (a >= n_agg) return;
  double *row = A3 + a * n_agg;
  for (uint32_t k = aptr[a]; k < aptr[a + 1]; ++k) {
    const uint32_t v = aidx[k];
    for (uint32_t q = vptr[v]; q < vptr[v + 1]; ++q) {
      const uint32_t ent = vpos[q];            // e * 4 + i
      const uint32_t e = ent >> 2, i = ent & 3u;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t col = agg[vert_c[e * 4 + j]];
        if (col != 0xffffffffu) row[col] += Ace[(int64_t)e * 16 + i * 4 + j];
      }
    }
  }
}

CommDev make_comm(const semk_comm *c) {
  CommDev d{};
  d.world = 1;
  if (c) {
    d.rank = c->rank;
    d.world = c->world;
    d.capacity = c->capacity;
    for (int w = 0; w < SEMK_COMM_MAX_WORLD; ++w) d.regions[w] = static_cast<char *>(c->regions[w]);
    d.status = c->status;
  }
  return d;
}

int check_comm(const semk_comm *c, const char *who) {
  if (!c) return SEMK_OK;
  if (c->world < 1 || c->world > SEMK_COMM_MAX_WORLD || c->rank < 0 || c->rank >= c->world ||
      c->capacity < 1 || !c->status) {
    semk_set_error(std::string(who) + ": inconsistent semk_comm");
    return SEMK_ERR_INVALID;
  }
  for (int w = 0; w < c->world; ++w)
    if (!c->regions[w]) {
      semk_set_error(std::string(who) + ": semk_comm with an unmapped region");
      return SEMK_ERR_INVALID;
    }
  return SEMK_OK;
}

}  // namespace

extern "C" int64_t semk_comm_region_bytes(int32_t world, int64_t capacity) {
  if (world < 1 || world > SEMK_COMM_MAX_WORLD || capacity < 1) return -1;
  return 256 + 2 * (int64_t)world * capacity * (int64_t)sizeof(double);
}

extern "C" int semk_comm_allreduce_f64(const semk_comm *comm, double *buf, int64_t n,
                                       void *stream) {
  SEMK_REQUIRE(comm && buf && n > 0, "semk_comm_allreduce_f64: bad argument");
  int rc = check_comm(comm, "semk_comm_allreduce_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(n <= comm->capacity, "semk_comm_allreduce_f64: n exceeds the region capacity");
  if (comm->world == 1) return SEMK_OK;
  allreduce_kernel<<<1, kStepThreads, 0, semk_stream(stream)>>>(make_comm(comm), buf, (int)n);
  SEMK_LAUNCH_CHECK("allreduce_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_coarse_ell_build_f64(const semk_sc_coarse *cs, int width, uint32_t *cols,
                                           double *vals, int32_t *overflow, void *stream) {
  SEMK_REQUIRE(cs && cs->n_v > 0 && cs->Ace && cs->vert_c && cs->vptr && cs->vpos && cols && vals &&
                   overflow,
               "semk_sc_coarse_ell_build_f64: bad argument");
  SEMK_REQUIRE(width >= 1 && width <= 32, "semk_sc_coarse_ell_build_f64: width outside [1, 32]");
  const unsigned grid = (unsigned)((cs->n_v + kT - 1) / kT);
  ml_ell_build_kernel<<<grid, kT, 0, semk_stream(stream)>>>(
      cs->n_v, width, cs->vptr, cs->vpos, cs->vert_c, cs->Ace, cs->dirichlet_c, cols, vals,
      overflow);
  SEMK_LAUNCH_CHECK("ml_ell_build_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_coarse_ell_apply_f64(const semk_sc_coarse *cs, const double *x, double *y,
                                           double *dot_out, void *stream) {
  SEMK_REQUIRE(cs && cs->n_v > 0 && cs->ell_width > 0 && cs->ell_cols && cs->ell_vals && x && y &&
                   x != y,
               "semk_sc_coarse_ell_apply_f64: bad argument");
  SEMK_REQUIRE(!dot_out || cs->partials, "semk_sc_coarse_ell_apply_f64: dot_out needs partials");
  ml_ell_spmv_kernel<<<blocks_for(cs->n_v), kT, 0, semk_stream(stream)>>>(
      cs->n_v, (int)cs->ell_width, cs->ell_cols, cs->ell_vals, x, y, dot_out, cs->partials);
  SEMK_LAUNCH_CHECK("ml_ell_spmv_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_top_assemble_f64(const semk_sc_coarse *cs, int64_t n_agg,
                                       const uint32_t *agg, const uint32_t *aptr_all,
                                       const uint32_t *aidx_all, double *A3, void *stream) {
  SEMK_REQUIRE(cs && cs->Ace && cs->vert_c && cs->vptr && cs->vpos && n_agg > 0 && agg &&
                   aptr_all && aidx_all && A3,
               "semk_sc_top_assemble_f64: bad argument");
  cudaStream_t st = semk_stream(stream);
  SEMK_CUDA_CHECK(cudaMemsetAsync(A3, 0, sizeof(double) * (size_t)n_agg * (size_t)n_agg, st));
  const unsigned grid = (unsigned)((n_agg + 63) / 64);
  ml_top_assemble_kernel<<<grid, 64, 0, st>>>(n_agg, agg, aptr_all, aidx_all, cs->vptr, cs->vpos,
                                              cs->vert_c, cs->Ace, A3);
  SEMK_LAUNCH_CHECK("ml_top_assemble_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_mlpcg_solve_f64(const semk_sc_op *op, const semk_sc_coarse *cs,
                                       const semk_sc_top *top, const semk_ml_dist *dist,
                                       const double *b, double *x, const double *dinv,
                                       const double *dinv_c, double *work, double *work_c,
                                       double *sc, double *vec_partials,
                                       const semk_ml_opts *opts, semk_ml_info *info,
                                       void *stream) {
  SEMK_REQUIRE(op && cs && b && x && dinv && dinv_c && work && work_c && sc && vec_partials &&
                   opts && info,
               "semk_sc_mlpcg_solve_f64: null pointer");
  SEMK_REQUIRE(opts->maxiter >= 0 && opts->inner_maxiter >= 1 && opts->rtol >= 0.0 &&
                   opts->inner_rtol > 0.0 && (opts->levels == 2 || opts->levels == 3) &&
                   opts->inner_chunk >= 1,
               "semk_sc_mlpcg_solve_f64: bad control");
  SEMK_REQUIRE(cs->pv && cs->pw && cs->rptr && cs->ridx && cs->rw,
               "semk_sc_mlpcg_solve_f64: missing transfer tables");
  const bool three = opts->levels == 3;
  if (three)
    SEMK_REQUIRE(top && top->n_agg > 0 && top->agg && top->aptr && top->aidx &&
                     (top->A3inv || top->A3inv_f32),
                 "semk_sc_mlpcg_solve_f64: levels = 3 needs a consistent semk_sc_top");
  const semk_comm *comm = dist ? dist->comm : nullptr;
  int rcode = check_comm(comm, "semk_sc_mlpcg_solve_f64");
  if (rcode != SEMK_OK) return rcode;
  const bool multi = comm && comm->world > 1;
  semk_halo *halo_f = multi ? dist->halo_f : nullptr;
  semk_halo *halo_c = multi ? dist->halo_c : nullptr;
  if (multi) SEMK_REQUIRE(halo_f && halo_c, "semk_sc_mlpcg_solve_f64: partition without halos");
  cudaStream_t st = semk_stream(stream);
  // On one GPU the inner iterations are replayed from a CUDA graph (eight small kernels per
  // iteration, a few microseconds each: launch gaps are a third of an inner solve on small
  // coarse levels).  Stream capture needs a non-default stream, so the whole solve moves to a
  // private one that is ordered after the caller's stream and joined with it at the end.
  // (Multi-GPU: the halo epochs are kernel arguments that change per launch -- no graph.)
  struct Private {
    cudaStream_t s = nullptr;
    cudaEvent_t in = nullptr, out = nullptr;
  };
  thread_local Private priv;
  const bool use_graph = !multi && opts->inner_chunk >= 1 && !getenv("SEMK_NO_GRAPH");
  cudaStream_t caller_stream = st;
  if (use_graph) {
    if (!priv.s) {
      SEMK_CUDA_CHECK(cudaStreamCreateWithFlags(&priv.s, cudaStreamNonBlocking));
      SEMK_CUDA_CHECK(cudaEventCreateWithFlags(&priv.in, cudaEventDisableTiming));
      SEMK_CUDA_CHECK(cudaEventCreateWithFlags(&priv.out, cudaEventDisableTiming));
    }
    SEMK_CUDA_CHECK(cudaEventRecord(priv.in, caller_stream));
    SEMK_CUDA_CHECK(cudaStreamWaitEvent(priv.s, priv.in, 0));
    st = priv.s;
  }
  struct Join {     // whatever way the function returns, the caller's stream waits for the solve
    bool on;
    cudaStream_t from, to;
    cudaEvent_t ev;
    ~Join() {
      if (on) {
        cudaEventRecord(ev, from);
        cudaStreamWaitEvent(to, ev, 0);
      }
    }
  } join{use_graph, st, caller_stream, priv.out};
  struct GraphHolder {
    cudaGraph_t g = nullptr;
    cudaGraphExec_t x = nullptr;
    ~GraphHolder() {
      if (x) cudaGraphExecDestroy(x);
      if (g) cudaGraphDestroy(g);
    }
  } inner_graph;
  const int64_t n = op->n_ext, nv = cs->n_v, n_elem = op->n_elem;
  const int64_t na = three ? top->n_agg : 0;
  const int64_t n_dot = multi ? dist->n_owned_f : n, nv_dot = multi ? dist->n_owned_c : nv;
  SEMK_REQUIRE(n_dot >= 0 && n_dot <= n && nv_dot >= 0 && nv_dot <= nv,
               "semk_sc_mlpcg_solve_f64: bad owned prefix");
  if (multi)
    SEMK_REQUIRE(comm->capacity >= na + 8, "semk_sc_mlpcg_solve_f64: comm capacity < n_agg + 8");
  const int64_t n_pad = (n + 31) & ~(int64_t)31, nv_pad = (nv + 31) & ~(int64_t)31;
  const int64_t na_pad = (na + 31) & ~(int64_t)31;
  double *r = work, *p = work + n_pad, *Ap = work + 2 * n_pad, *z = work + 3 * n_pad;
  double *rc = work_c, *xc = work_c + nv_pad;          // rc doubles as the inner residual
  double *pi = work_c + 2 * nv_pad, *Api = work_c + 3 * nv_pad;
  double *r3 = work_c + 6 * nv_pad, *y3 = r3 + na_pad + 32;   // r3[na..na+2) = {r.r, r.(dinv r)}
  double *so = sc, *si = sc + S_LEN;                   // scalar blocks: outer, inner
  double *scratch2 = sc + 2 * S_LEN;                   // {r.r, b.b} of the initial residual
  const int flags = SEMK_MASK_IN | SEMK_MASK_OUT | SEMK_DIRICHLET_IDENTITY;
  const CommDev cd = make_comm(multi ? comm : nullptr);
  const dim3 g(blocks_for(n)), gc(blocks_for(nv)), blk(kT);
  const int64_t want_a = (na * 32 + kT - 1) / kT;
  const dim3 ga((unsigned)(want_a < 1 ? 1 : (want_a < kMaxBlocks ? want_a : kMaxBlocks)));
  // pinned landing zone for the scalar blocks: per call, so concurrent solves on other
  // streams / threads do not share it
  double *h = nullptr;
  SEMK_CUDA_CHECK(cudaMallocHost(&h, 3 * S_LEN * sizeof(double)));
  struct Unpin {
    double *host;
    ~Unpin() { cudaFreeHost(host); }
  } unpin{h};

  auto fetch = [&]() -> int {
    SEMK_CUDA_CHECK(cudaMemcpyAsync(h, sc, 3 * S_LEN * sizeof(double), cudaMemcpyDeviceToHost, st));
    SEMK_CUDA_CHECK(cudaStreamSynchronize(st));
    return SEMK_OK;
  };
  auto step = [&](double *buf, int nred, int opcode, double *s) -> int {
    // the inner preconditioner is a fixed linear operator: standard beta there
    const int flex = (s == so) ? opts->flexible : 0;
    ml_step_kernel<<<1, kStepThreads, 0, st>>>(cd, buf, nred, opcode, flex, s, nullptr, nullptr,
                                               0);
    SEMK_LAUNCH_CHECK("ml_step_kernel");
    return SEMK_OK;
  };
  auto exchange = [&](semk_halo *hl, int64_t n_local, double *y, const double *u,
                      const uint8_t *dir, double *dot) -> int {
    if (!hl) return SEMK_OK;
    hl->epoch += 1;
    return semk_halo_exchange_f64(hl->n_col, n_local, y, u, dir, hl->mine, hl->left, hl->right,
                                  hl->epoch, dot, hl->status, st);
  };
  auto fine_apply = [&](const double *in, double *out, double *dot) -> int {
    int e = semk_sc_apply_f64(op, in, out, flags, dot, st);
    if (e != SEMK_OK) return e;
    return exchange(halo_f, n, out, in, op->dirichlet, dot);
  };
  const bool ell = cs->ell_width > 0 && cs->ell_cols && cs->ell_vals;
  auto coarse_apply = [&](const double *in, double *out, double *dot) -> int {
    if (ell) {
      ml_ell_spmv_kernel<<<gc, blk, 0, st>>>(nv, (int)cs->ell_width, cs->ell_cols, cs->ell_vals,
                                             in, out, dot, cs->partials);
      SEMK_LAUNCH_CHECK("ml_ell_spmv_kernel");
    } else {
      int e = semk_sc_coarse_apply_f64(n_elem, cs, in, out, flags, dot, st);
      if (e != SEMK_OK) return e;
    }
    return exchange(halo_c, nv, out, in, cs->dirichlet_c, dot);
  };
  // One inner iteration (first = the set-up step of a solve: no direction yet).  z = dinv r
  // + P2 y3 is never stored: its dot with r is r.(dinv r) + y3.r3, and the new direction is
  // formed straight from r, dinv and y3.  Two all-reduces: p.Ap and [r3 | r.r | r.(dinv r)].
  auto inner_step = [&](bool first) -> int {
    int e;
    if (!first) {
      if ((e = coarse_apply(pi, Api, si + S_PAP)) != SEMK_OK) return e;
      if ((e = step(si + S_PAP, 1, OP_ALPHA, si)) != SEMK_OK) return e;
    }
    if (ml_aligned16(pi, Api, dinv_c, xc, rc))
      ml_inner_update_kernel<true><<<gc, blk, 0, st>>>(nv, nv_dot, first ? 1 : 0, Api, dinv_c, rc,
                                                       si, r3 + na, vec_partials);
    else
      ml_inner_update_kernel<false><<<gc, blk, 0, st>>>(nv, nv_dot, first ? 1 : 0, Api, dinv_c, rc,
                                                        si, r3 + na, vec_partials);
    SEMK_LAUNCH_CHECK("ml_inner_update_kernel");
    if (three) {
      ml_agg_restrict_kernel<<<ga, blk, 0, st>>>(na, top->aptr, top->aidx, rc, r3);
      SEMK_LAUNCH_CHECK("ml_agg_restrict_kernel");
    }
    if (three) {
      if ((e = step(r3, (int)na + 2, first ? OP_RR2_FIRST : OP_RR2, si)) != SEMK_OK) return e;
      if (top->A3inv_f32)
        ml_dense_matvec_f32_kernel<<<ga, blk, 0, st>>>(na, top->A3inv_f32, r3, y3);
      else
        ml_dense_matvec_kernel<<<ga, blk, 0, st>>>(na, top->A3inv, r3, y3);
      SEMK_LAUNCH_CHECK("ml_dense_matvec_kernel");
      ml_step_kernel<<<1, kStepThreads, 0, st>>>(cd, nullptr, 0,
                                                 first ? OP_BETA_TOP_FIRST : OP_BETA_TOP, 0, si,
                                                 y3, r3, (int)na);
      SEMK_LAUNCH_CHECK("ml_step_kernel");
    } else {
      if ((e = step(r3, 2, first ? OP_RR2_BETA_FIRST : OP_RR2_BETA, si)) != SEMK_OK) return e;
    }
    if (ml_aligned16(dinv_c, rc, pi, xc, pi))
      ml_inner_direction_kernel<true><<<gc, blk, 0, st>>>(
          nv, first ? 1 : 0, dinv_c, rc, three ? top->agg : nullptr, y3, pi, xc, si);
    else
      ml_inner_direction_kernel<false><<<gc, blk, 0, st>>>(
          nv, first ? 1 : 0, dinv_c, rc, three ? top->agg : nullptr, y3, pi, xc, si);
    SEMK_LAUNCH_CHECK("ml_inner_direction_kernel");
    return SEMK_OK;
  };
  auto inner_iteration = [&]() -> int { return inner_step(false); };
  int64_t inner_sum = 0;
  int inner_solves = 0, predicted = 8;
  // xc ~= Ac^-1 rc (rc is overwritten by the inner residual)
  auto inner_solve = [&]() -> int {
    int e;
    SEMK_CUDA_CHECK(cudaMemsetAsync(xc, 0, sizeof(double) * nv, st));
    SEMK_CUDA_CHECK(cudaMemsetAsync(si, 0, sizeof(double) * S_TOL2, st));   // keeps S_TOL2
    SEMK_CUDA_CHECK(cudaMemsetAsync(pi, 0, sizeof(double) * nv, st));
    if ((e = inner_step(true)) != SEMK_OK) return e;
    int launched = 0;
    int chunk = predicted > 2 ? predicted - 1 : 1;
    while (launched < opts->inner_maxiter) {
      if (chunk > opts->inner_maxiter - launched) chunk = opts->inner_maxiter - launched;
      if (use_graph) {
        // one graph = inner_chunk iterations; iterations past convergence are no-ops (the
        // kernels read S_CONV), so rounding the chunk up to whole graphs is harmless
        const int per = opts->inner_chunk;
        if (!inner_graph.x) {
          SEMK_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
          int ce = SEMK_OK;
          for (int k = 0; k < per && ce == SEMK_OK; ++k) ce = inner_iteration();
          cudaError_t ee = cudaStreamEndCapture(st, &inner_graph.g);
          if (ce != SEMK_OK) return ce;
          SEMK_CUDA_CHECK(ee);
          SEMK_CUDA_CHECK(cudaGraphInstantiate(&inner_graph.x, inner_graph.g, 0));
        }
        const int reps = (chunk + per - 1) / per;
        for (int k = 0; k < reps; ++k) SEMK_CUDA_CHECK(cudaGraphLaunch(inner_graph.x, st));
        chunk = reps * per;
      } else {
        for (int k = 0; k < chunk; ++k)
          if ((e = inner_iteration()) != SEMK_OK) return e;
      }
      launched += chunk;
      if ((e = fetch()) != SEMK_OK) return e;
      const double *hi = h + S_LEN;
      if (hi[S_BREAK] != 0.0) {
        semk_set_error(hi[S_BREAK] == 2.0
                           ? "semk_sc_mlpcg_solve_f64: a rank did not arrive at an all-reduce"
                           : "semk_sc_mlpcg_solve_f64: inner breakdown (p.Ap <= 0 or non-finite)");
        return hi[S_BREAK] == 2.0 ? SEMK_ERR_CUDA : SEMK_ERR_BREAKDOWN;
      }
      if (hi[S_CONV] != 0.0) break;
      chunk = opts->inner_chunk;
    }
    const int its = (int)h[S_LEN + S_ITER];
    predicted = its;
    inner_sum += its;
    ++inner_solves;
    return SEMK_OK;
  };
  // z = dinv r is done by the update kernel; add P Ac^-1 P^T r and leave {r.z, z.Ap}
  auto precondition = [&](bool first) -> int {
    ml_restrict_kernel<<<gc, blk, 0, st>>>(nv, cs->rptr, cs->ridx, cs->rw, r, n_dot, rc);
    SEMK_LAUNCH_CHECK("ml_restrict_kernel");
    int e = exchange(halo_c, nv, rc, nullptr, nullptr, nullptr);
    if (e != SEMK_OK) return e;
    if ((e = inner_solve()) != SEMK_OK) return e;
    ml_prolong_dot_kernel<<<g, blk, 0, st>>>(n, n_dot, cs->pv, cs->pw, xc, r,
                                             first ? nullptr : Ap, z, so + S_RZN, vec_partials);
    SEMK_LAUNCH_CHECK("ml_prolong_dot_kernel");
    if ((e = step(so + S_RZN, 2, first ? OP_BETA_FIRST : OP_BETA, so)) != SEMK_OK) return e;
    ml_direction_kernel<<<g, blk, 0, st>>>(n, z, p, so);
    SEMK_LAUNCH_CHECK("ml_direction_kernel");
    return SEMK_OK;
  };

  // ---- outer loop ---------------------------------------------------------------------
  SEMK_CUDA_CHECK(cudaMemsetAsync(sc, 0, sizeof(double) * 3 * S_LEN, st));
  const double tol2 = opts->rtol * opts->rtol;
  h[0] = tol2;
  h[1] = opts->inner_rtol * opts->inner_rtol;
  SEMK_CUDA_CHECK(cudaMemcpyAsync(so + S_TOL2, h, sizeof(double), cudaMemcpyHostToDevice, st));
  SEMK_CUDA_CHECK(cudaMemcpyAsync(si + S_TOL2, h + 1, sizeof(double), cudaMemcpyHostToDevice, st));
  SEMK_CUDA_CHECK(cudaMemsetAsync(p, 0, sizeof(double) * n, st));
  if ((rcode = fine_apply(x, Ap, nullptr)) != SEMK_OK) return rcode;
  ml_resid_kernel<<<g, blk, 0, st>>>(n, b, Ap, op->dirichlet, r, n_dot, scratch2, vec_partials);
  SEMK_LAUNCH_CHECK("ml_resid_kernel");
  if (multi && (rcode = semk_comm_allreduce_f64(comm, scratch2, 2, st)) != SEMK_OK) return rcode;
  if ((rcode = fetch()) != SEMK_OK) return rcode;
  double rr = h[2 * S_LEN], bb = h[2 * S_LEN + 1];
  info->bnorm = sqrt(bb);
  info->iterations = 0;
  info->status = 0;
  info->inner_iterations = 0;
  info->inner_solves = 0;
  info->true_rel_residual = bb > 0.0 ? sqrt(rr / bb) : 0.0;
  info->rel_residual = info->true_rel_residual;
  if (bb == 0.0 || rr <= tol2 * bb) return SEMK_OK;
  h[2] = bb;   // (the fetch above has completed: h is free again until the next one)
  SEMK_CUDA_CHECK(cudaMemcpyAsync(so + S_BB, h + 2, sizeof(double), cudaMemcpyHostToDevice, st));
  ml_first_kernel<<<g, blk, 0, st>>>(n, n_dot, dinv, r, z, so + S_RR, vec_partials);
  SEMK_LAUNCH_CHECK("ml_first_kernel");
  if ((rcode = step(so + S_RR, 1, OP_RR, so)) != SEMK_OK) return rcode;
  if ((rcode = precondition(true)) != SEMK_OK) return rcode;
  int status = 1, it = 0;
  while (it < opts->maxiter) {
    if ((rcode = fine_apply(p, Ap, so + S_PAP)) != SEMK_OK) return rcode;
    if ((rcode = step(so + S_PAP, 1, OP_ALPHA, so)) != SEMK_OK) return rcode;
    ml_update_kernel<<<g, blk, 0, st>>>(n, n_dot, p, Ap, dinv, x, r, z, so, so + S_RR,
                                        vec_partials);
    SEMK_LAUNCH_CHECK("ml_update_kernel");
    if ((rcode = step(so + S_RR, 1, OP_RR, so)) != SEMK_OK) return rcode;
    if ((rcode = fetch()) != SEMK_OK) return rcode;
    ++it;
    rr = h[S_RR];
    if (h[S_BREAK] != 0.0) {
      status = SEMK_ERR_BREAKDOWN;
      break;
    }
    if (h[S_CONV] != 0.0) {
      status = 0;
      break;
    }
    if ((rcode = precondition(false)) != SEMK_OK) return rcode;
  }
  // true residual of the returned iterate
  if ((rcode = fine_apply(x, Ap, nullptr)) != SEMK_OK) return rcode;
  ml_resid_kernel<<<g, blk, 0, st>>>(n, b, Ap, op->dirichlet, z, n_dot, scratch2, vec_partials);
  SEMK_LAUNCH_CHECK("ml_resid_kernel");
  if (multi && (rcode = semk_comm_allreduce_f64(comm, scratch2, 2, st)) != SEMK_OK) return rcode;
  if ((rcode = fetch()) != SEMK_OK) return rcode;
  info->iterations = it;
  info->status = status;
  info->rel_residual = sqrt(rr / bb);
  info->true_rel_residual = sqrt(h[2 * S_LEN] / bb);
  info->inner_iterations = inner_sum;
  info->inner_solves = inner_solves;
  if (status == SEMK_ERR_BREAKDOWN) {
    semk_set_error(h[S_BREAK] == 2.0
                       ? "semk_sc_mlpcg_solve_f64: a rank did not arrive at an all-reduce"
                       : "semk_sc_mlpcg_solve_f64: breakdown (p.Ap <= 0 or non-finite)");
    return h[S_BREAK] == 2.0 ? SEMK_ERR_CUDA : SEMK_ERR_BREAKDOWN;
  }
  return SEMK_OK;
}

// ---------------------------------------------------------------------------------------
// Jacobi-PCG on a strip partition, native: the loop of semk_pcg_solve_f64 with the
// interface exchange after every apply and the two dot-product reductions of an iteration
// carried by peer-memory all-reduce kernels -- no NCCL call and no host code between the
// kernels; the host polls 128 bytes every check_every iterations.
// ---------------------------------------------------------------------------------------
extern "C" int semk_pcg_dist_solve_f64(const semk_op *op, const semk_sc_op *sc_op, semk_halo *halo,
                                       const semk_comm *comm, int64_t n_owned, const double *b,
                                       double *x, const double *dinv, double *work, double *sc,
                                       double *vec_partials, double rtol, int maxiter,
                                       int check_every, semk_pcg_info *info, void *stream) {
  SEMK_REQUIRE((op != nullptr) != (sc_op != nullptr),
               "semk_pcg_dist_solve_f64: exactly one of op / sc_op must be given");
  SEMK_REQUIRE(b && x && dinv && work && sc && vec_partials && info,
               "semk_pcg_dist_solve_f64: null pointer");
  SEMK_REQUIRE(maxiter >= 0 && check_every >= 1 && rtol >= 0.0,
               "semk_pcg_dist_solve_f64: bad control");
  int rcode = check_comm(comm, "semk_pcg_dist_solve_f64");
  if (rcode != SEMK_OK) return rcode;
  const bool multi = comm && comm->world > 1;
  if (multi) SEMK_REQUIRE(halo, "semk_pcg_dist_solve_f64: partition without a halo");
  cudaStream_t st = semk_stream(stream);
  const int64_t n = op ? op->n_nodes : sc_op->n_ext;
  const uint8_t *dirichlet = op ? op->dirichlet : sc_op->dirichlet;
  const int64_t n_dot = multi ? n_owned : n;
  SEMK_REQUIRE(n_dot >= 0 && n_dot <= n, "semk_pcg_dist_solve_f64: bad owned prefix");
  const int64_t n_pad = (n + 31) & ~(int64_t)31;
  double *r = work, *p = work + n_pad, *Ap = work + 2 * n_pad;
  double *s0 = sc, *pair = sc + S_LEN;       // scalar block; {r.r, r.(dinv r)} / {r.r, b.b}
  const int flags = SEMK_MASK_IN | SEMK_MASK_OUT | SEMK_DIRICHLET_IDENTITY;
  const CommDev cd = make_comm(multi ? comm : nullptr);
  const dim3 g(blocks_for(n)), blk(kT);
  double *h = nullptr;
  SEMK_CUDA_CHECK(cudaMallocHost(&h, 2 * S_LEN * sizeof(double)));
  struct Unpin {
    double *host;
    ~Unpin() { cudaFreeHost(host); }
  } unpin{h};
  auto fetch = [&]() -> int {
    SEMK_CUDA_CHECK(cudaMemcpyAsync(h, sc, 2 * S_LEN * sizeof(double), cudaMemcpyDeviceToHost, st));
    SEMK_CUDA_CHECK(cudaStreamSynchronize(st));
    return SEMK_OK;
  };
  auto step = [&](double *buf, int nred, int opcode) -> int {
    ml_step_kernel<<<1, kStepThreads, 0, st>>>(cd, buf, nred, opcode, 0, s0, nullptr, nullptr, 0);
    SEMK_LAUNCH_CHECK("ml_step_kernel");
    return SEMK_OK;
  };
  auto apply = [&](const double *in, double *out, double *dot) -> int {
    int e = op ? semk_poisson_apply_f64(op, in, out, flags, dot, st)
               : semk_sc_apply_f64(sc_op, in, out, flags, dot, st);
    if (e != SEMK_OK || !multi) return e;
    halo->epoch += 1;
    return semk_halo_exchange_f64(halo->n_col, n, out, in, dirichlet, halo->mine, halo->left,
                                  halo->right, halo->epoch, dot, halo->status, st);
  };
  auto update = [&](int first) -> int {
    if (ml_aligned16(p, Ap, dinv, x, r))
      ml_inner_update_kernel<true><<<g, blk, 0, st>>>(n, n_dot, first, Ap, dinv, r, s0, pair,
                                                      vec_partials);
    else
      ml_inner_update_kernel<false><<<g, blk, 0, st>>>(n, n_dot, first, Ap, dinv, r, s0, pair,
                                                       vec_partials);
    SEMK_LAUNCH_CHECK("ml_inner_update_kernel");
    return SEMK_OK;
  };
  auto direction = [&](int first) -> int {
    if (ml_aligned16(dinv, r, p, x, p))
      ml_inner_direction_kernel<true><<<g, blk, 0, st>>>(n, first, dinv, r, nullptr, nullptr, p, x,
                                                         s0);
    else
      ml_inner_direction_kernel<false><<<g, blk, 0, st>>>(n, first, dinv, r, nullptr, nullptr, p,
                                                          x, s0);
    SEMK_LAUNCH_CHECK("ml_inner_direction_kernel");
    return SEMK_OK;
  };

  SEMK_CUDA_CHECK(cudaMemsetAsync(sc, 0, sizeof(double) * 2 * S_LEN, st));
  SEMK_CUDA_CHECK(cudaMemsetAsync(p, 0, sizeof(double) * n, st));
  if ((rcode = apply(x, Ap, nullptr)) != SEMK_OK) return rcode;
  ml_resid_kernel<<<g, blk, 0, st>>>(n, b, Ap, dirichlet, r, n_dot, pair, vec_partials);
  SEMK_LAUNCH_CHECK("ml_resid_kernel");
  if (multi && (rcode = semk_comm_allreduce_f64(comm, pair, 2, st)) != SEMK_OK) return rcode;
  if ((rcode = fetch()) != SEMK_OK) return rcode;
  const double rr0 = h[S_LEN], bb = h[S_LEN + 1], tol2 = rtol * rtol;
  info->bnorm = sqrt(bb);
  info->iterations = 0;
  info->status = 0;
  info->rel_residual = bb > 0.0 ? sqrt(rr0 / bb) : 0.0;
  if (bb == 0.0 || rr0 <= tol2 * bb) return SEMK_OK;
  h[0] = bb;
  h[1] = tol2;
  SEMK_CUDA_CHECK(cudaMemcpyAsync(s0 + S_BB, h, sizeof(double), cudaMemcpyHostToDevice, st));
  SEMK_CUDA_CHECK(cudaMemcpyAsync(s0 + S_TOL2, h + 1, sizeof(double), cudaMemcpyHostToDevice, st));
  if ((rcode = update(1)) != SEMK_OK) return rcode;
  if ((rcode = step(pair, 2, OP_RR2_BETA_KEEP)) != SEMK_OK) return rcode;
  if ((rcode = direction(1)) != SEMK_OK) return rcode;
  int status = 1, launched = 0;
  while (launched < maxiter) {
    const int chunk = check_every < maxiter - launched ? check_every : maxiter - launched;
    for (int k = 0; k < chunk; ++k) {
      if ((rcode = apply(p, Ap, s0 + S_PAP)) != SEMK_OK) return rcode;
      if ((rcode = step(s0 + S_PAP, 1, OP_ALPHA)) != SEMK_OK) return rcode;
      if ((rcode = update(0)) != SEMK_OK) return rcode;
      if ((rcode = step(pair, 2, OP_RR2_BETA)) != SEMK_OK) return rcode;
      if ((rcode = direction(0)) != SEMK_OK) return rcode;
    }
    launched += chunk;
    if ((rcode = fetch()) != SEMK_OK) return rcode;
    if (h[S_BREAK] != 0.0) {
      status = SEMK_ERR_BREAKDOWN;
      break;
    }
    if (h[S_CONV] != 0.0) {
      status = 0;
      break;
    }
  }
  info->iterations = (int32_t)h[S_ITER];
  info->status = status;
  info->rel_residual = sqrt(h[S_RR] / bb);
  if (status == SEMK_ERR_BREAKDOWN) {
    semk_set_error(h[S_BREAK] == 2.0
                       ? "semk_pcg_dist_solve_f64: a rank did not arrive at an all-reduce"
                       : "semk_pcg_dist_solve_f64: breakdown (p.Ap <= 0 or non-finite)");
    return h[S_BREAK] == 2.0 ? SEMK_ERR_CUDA : SEMK_ERR_BREAKDOWN;
  }
  return SEMK_OK;
}
