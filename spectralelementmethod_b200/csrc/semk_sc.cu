// semk_sc.cu -- static condensation on the device.
//
// The reference's solver formulation (DOFManagerSC, sem/discrete.py:283-528):
// eliminate the (p-1)^2 interior DOFs of every element,
//     S_e = A_ee - A_ei A_ii^{-1} A_ie,   g_e = f_e - A_ei A_ii^{-1} f_i
// (compute_local_sc_system, :438-476), assemble and solve the condensed system
// over the element-exterior DOFs (:478-511) and recover the interiors element
// by element (:513-524).  Here:
//
//   sc_element_kernel<N>  one CTA per element.  Rebuilds the 4-index local
//       stiffness of examples/poisson.py:166-193 entry by entry from the
//       geometric factors (closed form below), Cholesky-factorises the interior
//       block in shared memory (A_ii = L L^T), Z = L^{-1} [A_ie | f_i], and
//       writes  S_e = A_ee - Z^T Z (packed lower triangle),
//               g_e = f_e - Z^T (L^{-1} f_i),
//               u_i = L^{-T} (L^{-1} f_i - Z u_e)      (back-substitution).
//       Set-up / once-per-solve work; the factorisation is recomputed instead
//       of stored (19 KB per element at p = 8 against 4 KB for S_e).
//   sc_matvec_kernel<N>   the condensed apply, HBM-bound: persistent CTAs stream the
//       packed S_e blocks through a double-buffered TMA pipeline (evict-first),
//       gather the 4p exterior values through the L2G map one group ahead and
//       form y_loc = S_e u_e, one thread per row, fixed summation order.
//   sc_node_kernel        every exterior node sums its entries of y_loc in the
//       fixed order of the node -> entries table (the reference's
//       `grhs[inds_ext] += ...`, :499, without atomics), applies the Dirichlet
//       elimination (:505-510) and takes the fused u.y.
//
// Local stiffness entry (p,q),(r,s) of  y = D^T (G00 ur + G01 us) + (G01 ur + G11 us) D,
// ur = D u, us = u D^T:
//   A = [q==s] sum_m D[m][p] G00[m][q] D[m][r]  +  D[r][p] G01[r][q] D[q][s]
//     + D[p][r] G01[p][s] D[s][q]               +  [p==r] sum_n G11[p][n] D[n][s] D[n][q]
#include <math.h>

#include "semk_common.cuh"

namespace {

constexpr int kScThreads = 128;
constexpr int kScTX = 16, kScTY = 8;  // trailing-update thread tile of the factorisation

template <int N>
struct ScCfg {
  static constexpr int NN = N * N;
  static constexpr int NE = 4 * (N - 1);         // exterior nodes of an element
  static constexpr int NI = (N - 2) * (N - 2);   // interior nodes
  static constexpr int NS = NE * (NE + 1) / 2;   // packed S_e (always even)
  static constexpr int LDI = NI | 1;             // odd row strides: conflict-free columns
  static constexpr int LDZ = NE + 1;             // NE columns of A_ie + the load column
  static constexpr int kDoubles =
      3 * NN + NN + NN + NI * LDI + NI * LDZ + NE * NE + NI + NE + NE + NI + NI * LDI;
  static constexpr size_t kSmem = sizeof(double) * kDoubles + sizeof(int) * (NE + 4);
  static constexpr int EPB = kScThreads / NE;    // elements per CTA step of the matvec
};

struct ScElemArgs {
  int64_t n_elem;
  const int64_t *slot_of_elem;
  const double *G;
  int64_t g_patch_stride;
  int pe;
  const double *D;
  const int32_t *ext_loc;
  const uint32_t *l2g;
  const double *JxW;
  const double *f_nodal;
  double f_scale;
  int mode;
  double *S_out;
  int64_t s_stride;
  double *sdiag_loc;
  double *g_loc;
  double *u;
  int32_t *bad_flag;
  // dense mode (semk_sc_element_dense_f64): the caller's own local systems, hierarchical
  // local order (exterior DOFs first), instead of the Poisson recipe on G
  double *W_out;               // [n_elem][NE][NI] A_ii^{-1} A_ie, transposed (mode STORE), or nullptr
  double *c_out;               // [n_elem][NI] A_ii^{-1} f_i (with the load), or nullptr
  const double *A_dense;       // [n_elem][NN][NN] or nullptr
  const double *f_dense;       // [n_elem][NN]
  const uint32_t *l2g_hier;    // [n_elem][NN] global ids in hierarchical local order
  const double *react;         // [n_elem][NN] nodal reaction term added to the local diagonal, or nullptr
  double *Ainv_out;            // [n_elem][NI][NI] A_ii^{-1} (mode STORE_INV), or nullptr
};

template <int N>
__global__ void __launch_bounds__(kScThreads) sc_element_kernel(ScElemArgs a) {
  using C = ScCfg<N>;
  constexpr int NN = C::NN, NE = C::NE, NI = C::NI, NS = C::NS, LDI = C::LDI, LDZ = C::LDZ;
  constexpr int M = N - 2;
  extern __shared__ __align__(128) double sc_smem[];
  double *sG = sc_smem;            // [3][N][N] geometric factors G00, G01, G11
  double *sD = sG + 3 * NN;        // [N][N]
  double *sR = sD + NN;            // [N][N] reaction (nodal diagonal) term, zeros without one
  double *sA = sR + NN;            // [NI][LDI] interior block -> Cholesky factor (lower)
  double *sZ = sA + NI * LDI;      // [NI][LDZ] A_ie | f_i  ->  L^{-1} of both
  double *sE = sZ + NI * LDZ;      // [NE][NE] A_ee (lower part used)
  double *sInv = sE + NE * NE;     // [NI] 1 / L[i][i]
  double *sFe = sInv + NI;         // [NE] exterior part of the load
  double *sUe = sFe + NE;          // [NE] exterior values (back-substitution)
  double *sT = sUe + NE;           // [NI] back-substitution vector
  double *sX = sT + NI;            // [NI][LDI] L^{-1} (mode STORE_INV only)
  int *sExt = reinterpret_cast<int *>(sX + NI * LDI);  // [NE] lexicographic index of exterior k
  int *sBad = sExt + NE;
  const int tid = threadIdx.x;
  const int tx = tid % kScTX, ty = tid / kScTX;
  const int NP = N * a.pe;
  const bool need_f = (a.mode & (SEMK_SC_RHS | SEMK_SC_BACKSOLVE)) != 0;

  auto entry = [&](int p, int q, int r, int s) -> double {
    double v = sD[r * N + p] * sG[NN + r * N + q] * sD[q * N + s] +
               sD[p * N + r] * sG[NN + p * N + s] * sD[s * N + q];
    if (q == s) {
#pragma unroll
      for (int m = 0; m < N; ++m) v = fma(sD[m * N + p] * sD[m * N + r], sG[m * N + q], v);
    }
    if (p == r) {
#pragma unroll
      for (int n = 0; n < N; ++n)
        v = fma(sD[n * N + s] * sD[n * N + q], sG[2 * NN + p * N + n], v);
      if (q == s) v += sR[p * N + q];
    }
    return v;
  };

  for (int64_t e = blockIdx.x; e < a.n_elem; e += gridDim.x) {
    __syncthreads();  // the previous element is completely done with shared memory
    if (!a.A_dense) {
      const int64_t slot = a.slot_of_elem ? a.slot_of_elem[e] : e;
      const int64_t patch = slot / a.pe;
      const int lp = (int)(slot - patch * a.pe);
      const double *g = a.G + patch * a.g_patch_stride + lp * N;
      for (int i = tid; i < 3 * NN; i += kScThreads) {
        const int c = i / NN, k = i - c * NN;
        const int m = k / N, n = k - m * N;
        sG[i] = g[(c * N + m) * NP + n];
      }
      for (int i = tid; i < NN; i += kScThreads) sD[i] = a.D[i];
      for (int i = tid; i < NN; i += kScThreads) sR[i] = a.react ? a.react[e * NN + i] : 0.0;
      for (int i = tid; i < NE; i += kScThreads) sExt[i] = a.ext_loc[i];
    }
    if (tid == 0) *sBad = 0;
    __syncthreads();
    // ---- local stiffness blocks --------------------------------------------------
    if (a.A_dense) {
      // the caller's local matrix, hierarchical order: rows / columns [0, NE) exterior,
      // [NE, NN) interior (reorder_local_system_hier, sem/discrete.py:428-436)
      const double *A = a.A_dense + e * (int64_t)NN * NN;
      for (int idx = tid; idx < NI * NI; idx += kScThreads) {
        const int i = idx / NI, j = idx - i * NI;
        if (j <= i) sA[i * LDI + j] = A[(int64_t)(NE + i) * NN + NE + j];
      }
      for (int idx = tid; idx < NI * NE; idx += kScThreads) {
        const int i = idx / NE, k = idx - i * NE;
        sZ[i * LDZ + k] = A[(int64_t)(NE + i) * NN + k];
      }
      for (int idx = tid; idx < NE * NE; idx += kScThreads) {
        const int k = idx / NE, l = idx - k * NE;
        if (l <= k) sE[idx] = A[(int64_t)k * NN + l];
      }
      if (need_f) {
        const double *fh = a.f_dense + e * (int64_t)NN;
        const uint32_t *row = a.l2g_hier + e * (int64_t)NN;
        for (int i = tid; i < NI; i += kScThreads) sZ[i * LDZ + NE] = fh[NE + i];
        for (int k = tid; k < NE; k += kScThreads) {
          sFe[k] = fh[k];
          if (a.mode & SEMK_SC_BACKSOLVE) sUe[k] = a.u[row[k]];
        }
      }
    } else {
    for (int idx = tid; idx < NI * NI; idx += kScThreads) {
      const int i = idx / NI, j = idx - i * NI;
      if (j <= i) sA[i * LDI + j] = entry(1 + i / M, 1 + i % M, 1 + j / M, 1 + j % M);
    }
    for (int idx = tid; idx < NI * NE; idx += kScThreads) {
      const int i = idx / NE, k = idx - i * NE;
      const int x = sExt[k];
      sZ[i * LDZ + k] = entry(1 + i / M, 1 + i % M, x / N, x % N);
    }
    for (int idx = tid; idx < NE * NE; idx += kScThreads) {
      const int k = idx / NE, l = idx - k * NE;
      if (l <= k) {
        const int x = sExt[k], z = sExt[l];
        sE[idx] = entry(x / N, x % N, z / N, z % N);
      }
    }
    }
    if (need_f && !a.A_dense) {
      const double *jw = a.JxW + e * NN;
      const uint32_t *row = a.l2g + e * NN;
      for (int i = tid; i < NI; i += kScThreads) {
        const int k = (1 + i / M) * N + 1 + i % M;
        const double f = a.f_nodal ? a.f_nodal[row[k]] : 1.0;
        sZ[i * LDZ + NE] = a.f_scale * jw[k] * f;
      }
      for (int k = tid; k < NE; k += kScThreads) {
        const int x = sExt[k];
        const double f = a.f_nodal ? a.f_nodal[row[x]] : 1.0;
        sFe[k] = a.f_scale * jw[x] * f;
        if (a.mode & SEMK_SC_BACKSOLVE) sUe[k] = a.u[row[x]];
      }
    }
    __syncthreads();
    // ---- A_ii = L L^T, right-looking, in place (lower triangle) --------------------
    for (int k = 0; k < NI; ++k) {
      const double d = sA[k * LDI + k];  // nobody writes (k,k) during step k
      const double inv = 1.0 / sqrt(d);
      if (tid == 0) {
        sInv[k] = inv;
        if (!(d > 0.0)) *sBad = 1;
      }
      for (int i = k + 1 + tid; i < NI; i += kScThreads) sA[i * LDI + k] *= inv;
      __syncthreads();
      for (int i = k + 1 + ty; i < NI; i += kScTY) {
        const double lik = sA[i * LDI + k];
        for (int j = k + 1 + tx; j <= i; j += kScTX)
          sA[i * LDI + j] = fma(-lik, sA[j * LDI + k], sA[i * LDI + j]);
      }
      __syncthreads();
    }
    // ---- Z = L^{-1} [A_ie | f_i]: one thread per column, forward substitution -------
    if (tid < NE + (need_f ? 1 : 0)) {
      for (int i = 0; i < NI; ++i) {
        double acc = sZ[i * LDZ + tid];
        for (int j = 0; j < i; ++j) acc = fma(-sA[i * LDI + j], sZ[j * LDZ + tid], acc);
        sZ[i * LDZ + tid] = acc * sInv[i];
      }
    }
    __syncthreads();
    if (tid == 0 && *sBad) *a.bad_flag = 1;
    // ---- S_e = A_ee - Z^T Z (packed lower triangle, row major) ----------------------
    if (a.mode & SEMK_SC_SCHUR) {
      double *So = a.S_out + e * a.s_stride;
      for (int q = tid; q < NS; q += kScThreads) {
        int k = (int)((sqrt(8.0 * q + 1.0) - 1.0) * 0.5);
        while ((k + 1) * (k + 2) / 2 <= q) ++k;
        while (k * (k + 1) / 2 > q) --k;
        const int l = q - k * (k + 1) / 2;
        double acc = sE[k * NE + l];
        for (int i = 0; i < NI; ++i) acc = fma(-sZ[i * LDZ + k], sZ[i * LDZ + l], acc);
        So[q] = acc;
        if (k == l && a.sdiag_loc) a.sdiag_loc[e * NE + k] = acc;
      }
    }
    // ---- g_e = f_e - Z^T (L^{-1} f_i) ------------------------------------------------
    if (a.mode & SEMK_SC_RHS) {
      for (int k = tid; k < NE; k += kScThreads) {
        double acc = sFe[k];
        for (int i = 0; i < NI; ++i) acc = fma(-sZ[i * LDZ + k], sZ[i * LDZ + NE], acc);
        a.g_loc[e * NE + k] = acc;
      }
    }
    // ---- u_i = L^{-T} (L^{-1} f_i - Z u_e) -------------------------------------------
    if (a.mode & SEMK_SC_BACKSOLVE) {
      for (int i = tid; i < NI; i += kScThreads) {
        double acc = sZ[i * LDZ + NE];
        for (int k = 0; k < NE; ++k) acc = fma(-sZ[i * LDZ + k], sUe[k], acc);
        sT[i] = acc;
      }
      __syncthreads();
      if (tid < 32) {
        for (int i = NI - 1; i >= 0; --i) {
          const double xi = sT[i] * sInv[i];
          __syncwarp();
          if (tid == 0) sT[i] = xi;
          for (int j = tid; j < i; j += 32) sT[j] = fma(-sA[i * LDI + j], xi, sT[j]);
          __syncwarp();
        }
      }
      __syncthreads();
      if (a.A_dense) {
        const uint32_t *row = a.l2g_hier + e * (int64_t)NN;
        for (int i = tid; i < NI; i += kScThreads) a.u[row[NE + i]] = sT[i];
      } else {
        const uint32_t *row = a.l2g + e * NN;
        for (int i = tid; i < NI; i += kScThreads) a.u[row[(1 + i / M) * N + 1 + i % M]] = sT[i];
      }
    }
    // ---- keep A_ii^{-1} = L^{-T} L^{-1} (symmetric, dense): with W it turns the condensed
    // load of ANY later right-hand side into a streaming pass (sc_load_stored_kernel) --------
    if ((a.mode & SEMK_SC_STORE_INV) && a.Ainv_out) {
      __syncthreads();
      if (tid < NI) {   // column tid of L^{-1}: forward substitution of the unit vector
        for (int i = 0; i < tid; ++i) sX[i * LDI + tid] = 0.0;
        sX[tid * LDI + tid] = sInv[tid];
        for (int i = tid + 1; i < NI; ++i) {
          double acc = 0.0;
          for (int j = tid; j < i; ++j) acc = fma(-sA[i * LDI + j], sX[j * LDI + tid], acc);
          sX[i * LDI + tid] = acc * sInv[i];
        }
      }
      __syncthreads();
      double *Ao = a.Ainv_out + e * (int64_t)NI * NI;
      for (int idx = tid; idx < NI * NI; idx += kScThreads) {
        const int i = idx / NI, j = idx - i * NI;
        const int k0 = i > j ? i : j;
        double acc = 0.0;
        for (int k = k0; k < NI; ++k) acc = fma(sX[k * LDI + i], sX[k * LDI + j], acc);
        Ao[idx] = acc;
      }
    }
    // ---- keep the interior solution operator: W = A_ii^{-1} A_ie = L^{-T} Z and
    // c = A_ii^{-1} f_i = L^{-T} (L^{-1} f_i), one thread per column, backward substitution
    // in place; the back-substitution u_i = c - W u_e then is one streaming pass
    // (sc_backsolve_stored_kernel) instead of a refactorisation per element ---------------
    const bool store_w = (a.mode & SEMK_SC_STORE) && a.W_out;
    const bool store_c = need_f && a.c_out;
    if (store_w || store_c) {
      __syncthreads();  // every reader of Z is done
      const bool mine = (store_w && tid < NE) || (store_c && tid == NE);
      if (mine) {
        for (int i = NI - 1; i >= 0; --i) {
          double acc = sZ[i * LDZ + tid];
          for (int j = i + 1; j < NI; ++j) acc = fma(-sA[j * LDI + i], sZ[j * LDZ + tid], acc);
          sZ[i * LDZ + tid] = acc * sInv[i];
        }
      }
      __syncthreads();
      if (store_w) {
        double *Wo = a.W_out + e * (int64_t)NE * NI;
        for (int idx = tid; idx < NE * NI; idx += kScThreads) {
          const int k = idx / NI, i = idx - k * NI;
          Wo[idx] = sZ[i * LDZ + k];
        }
      }
      if (store_c) {
        double *co = a.c_out + e * (int64_t)NI;
        for (int i = tid; i < NI; i += kScThreads) co[i] = sZ[i * LDZ + NE];
      }
    }
  }
}

// Back-substitution from the stored interior operator: u_i = c - W u_e for every element,
// one thread per interior row; W is stored transposed ([NE][NI]), so the NI threads of an
// element read consecutive doubles.  A pure stream over W (12.5 KB per element at p = 8).
template <int N>
__global__ void __launch_bounds__(256)
    sc_backsolve_stored_kernel(int64_t n_elem, const double *__restrict__ W,
                               const double *__restrict__ c, const uint32_t *__restrict__ l2g,
                               const int32_t *__restrict__ ext_loc, double *__restrict__ u) {
  using C = ScCfg<N>;
  constexpr int NN = C::NN, NE = C::NE, NI = C::NI, M = N - 2;
  constexpr int GT = ((NI + 31) / 32) * 32;     // threads per element
  constexpr int GPB = 256 / GT;                 // elements per CTA step
  __shared__ double sUe[GPB][NE];
  const int tid = threadIdx.x;
  const int grp = tid / GT, i = tid - grp * GT;
  const int64_t n_steps = (n_elem + GPB - 1) / GPB;
  for (int64_t step = blockIdx.x; step < n_steps; step += gridDim.x) {
    const int64_t e = step * GPB + grp;
    const bool on = grp < GPB && e < n_elem;
    __syncthreads();
    if (on) {
      const uint32_t *row = l2g + e * NN;
      for (int k = i; k < NE; k += GT) sUe[grp][k] = u[row[ext_loc[k]]];
    }
    __syncthreads();
    if (on && i < NI) {
      const double *Wt = W + e * (int64_t)NE * NI + i;
      double acc = c ? c[e * (int64_t)NI + i] : 0.0;
#pragma unroll 8
      for (int k = 0; k < NE; ++k) acc = fma(-__ldcs(Wt + (int64_t)k * NI), sUe[grp][k], acc);
      u[l2g[e * NN + (1 + i / M) * N + 1 + i % M]] = acc;
    }
  }
}

// Condensed load from the stored interior operators: for every element
//   c = A_ii^{-1} f_i,   g = f_e - W^T f_i   (A_ei A_ii^{-1} = W^T by symmetry),
// f = f_scale * JxW * f_nodal[l2g] as in sc_element_kernel.  A pure stream over A_ii^{-1}
// (19 KB) and W (12.5 KB per element at p = 8) instead of a refactorisation per element.
template <int N>
__global__ void __launch_bounds__(256)
    sc_load_stored_kernel(int64_t n_elem, const double *__restrict__ W,
                          const double *__restrict__ Ainv, const uint32_t *__restrict__ l2g,
                          const int32_t *__restrict__ ext_loc, const double *__restrict__ JxW,
                          const double *__restrict__ f_nodal, double f_scale,
                          double *__restrict__ g_loc, double *__restrict__ c_out) {
  using C = ScCfg<N>;
  constexpr int NN = C::NN, NE = C::NE, NI = C::NI, M = N - 2;
  constexpr int GT = ((NI + 31) / 32) * 32;
  constexpr int GPB = 256 / GT;
  __shared__ double sF[GPB][NN];
  const int tid = threadIdx.x;
  const int grp = tid / GT, i = tid - grp * GT;
  const int64_t n_steps = (n_elem + GPB - 1) / GPB;
  for (int64_t step = blockIdx.x; step < n_steps; step += gridDim.x) {
    const int64_t e = step * GPB + grp;
    const bool on = grp < GPB && e < n_elem;
    __syncthreads();
    if (on) {
      const uint32_t *row = l2g + e * NN;
      const double *jw = JxW + e * NN;
      for (int k = i; k < NN; k += GT)
        sF[grp][k] = f_scale * jw[k] * (f_nodal ? f_nodal[row[k]] : 1.0);
    }
    __syncthreads();
    if (on && i < NI) {
      // A_ii^{-1} is symmetric: row i read as column i, consecutive across the threads
      const double *Ac = Ainv + e * (int64_t)NI * NI + i;
      double acc = 0.0;
#pragma unroll 7
      for (int j = 0; j < NI; ++j)
        acc = fma(__ldcs(Ac + (int64_t)j * NI), sF[grp][(1 + j / M) * N + 1 + j % M], acc);
      c_out[e * (int64_t)NI + i] = acc;
    }
    if (on && i < NE) {
      const double *Wr = W + e * (int64_t)NE * NI + (int64_t)i * NI;
      double acc = sF[grp][ext_loc[i]];
      for (int j = 0; j < NI; ++j)
        acc = fma(-Wr[j], sF[grp][(1 + j / M) * N + 1 + j % M], acc);
      g_loc[e * (int64_t)NE + i] = acc;
    }
  }
}

// ---- condensed apply, element part: y_loc = S_e u[l2g_ext] ------------------------------
// Persistent and double-buffered: while group i (EPB elements) is being multiplied,
// the TMA engine (cp.async.bulk, evict-first: S is streamed exactly once per apply)
// fills the other stage with the S blocks of the CTA's next group and every thread has
// its next u value in flight (index load, then value load); one barrier per group.
// One thread per row of S_e; entry (r, c), r >= c, of the packed block sits at
// r (r + 1) / 2 + c, so a half-warp reading one column hits 16 distinct banks
// (triangular numbers are a permutation mod 16).
template <int N>
__global__ void __launch_bounds__(kScThreads)
    sc_matvec_kernel(int64_t n_elem, const double *__restrict__ S,
                         const uint32_t *__restrict__ l2g_ext, const double *__restrict__ u,
                         const uint8_t *__restrict__ mask_in, double *__restrict__ y_loc) {
  using C = ScCfg<N>;
  constexpr int NE = C::NE, NS = C::NS, EPB = C::EPB;
  constexpr int STG = EPB * NS, UST = EPB * NE;
  extern __shared__ __align__(128) double sc_smem[];
  double *sS = sc_smem;                                            // [2][EPB][NS]
  double *sU = sS + 2 * STG;                                       // [2][EPB][NE]
  uint64_t *mbar = reinterpret_cast<uint64_t *>(sU + 2 * UST);     // [2]
  const int tid = threadIdx.x;
  const int le = tid / NE, k = tid - le * NE;
  const bool lane_on = le < EPB;
  const int64_t n_groups = (n_elem + EPB - 1) / EPB;
  const uint64_t policy = semk_policy_evict_first();
  if (tid == 0) {
    semk_mbar_init(&mbar[0], 1);
    semk_mbar_init(&mbar[1], 1);
    semk_fence_mbar_init();
  }
  __syncthreads();
  auto count_of = [&](int64_t grp) -> int {
    const int64_t left = n_elem - grp * EPB;
    return (int)(left < (int64_t)EPB ? left : (int64_t)EPB);
  };
  auto issue = [&](int64_t grp, int s) {  // thread 0
    const uint32_t bytes = (uint32_t)(count_of(grp) * NS * sizeof(double));
    semk_mbar_expect_tx(&mbar[s], bytes);
    semk_bulk_g2s_hint(sS + s * STG, S + grp * EPB * NS, bytes, &mbar[s], policy);
  };
  auto gather = [&](int64_t grp) -> double {
    double v = 0.0;
    if (lane_on && le < count_of(grp)) {
      const uint32_t g = l2g_ext[(grp * EPB + le) * NE + k];
      v = u[g];
      if (mask_in && mask_in[g]) v = 0.0;
    }
    return v;
  };
  int64_t grp = blockIdx.x;
  if (grp >= n_groups) return;
  if (tid == 0) issue(grp, 0);
  double unext = gather(grp);
  if (lane_on) sU[tid] = unext;
  __syncthreads();
  for (int it = 0;; ++it) {
    const int s = it & 1;
    const int64_t nxt = grp + gridDim.x;
    const bool more = nxt < n_groups;
    if (more) {
      if (tid == 0) issue(nxt, s ^ 1);  // stage s^1 was last read before the barrier below
      unext = gather(nxt);
    }
    semk_mbar_wait(&mbar[s], (uint32_t)((it >> 1) & 1));
    if (lane_on && le < count_of(grp)) {
      const double *sm = sS + s * STG + le * NS;
      const double *uu = sU + s * UST + le * NE;
      const int kk = k * (k + 1) / 2;
      double acc = 0.0;
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        const int idx = (j <= k) ? kk + j : j * (j + 1) / 2 + k;
        acc = fma(sm[idx], uu[j], acc);
      }
      y_loc[(grp * EPB + le) * NE + k] = acc;
    }
    if (!more) break;
    if (lane_on) sU[(s ^ 1) * UST + tid] = unext;
    __syncthreads();
    grp = nxt;
  }
}

template <int N>
constexpr size_t sc_matvec_smem() {
  return sizeof(double) * (2 * (size_t)ScCfg<N>::EPB * (ScCfg<N>::NS + ScCfg<N>::NE)) + 16;
}

// ---- vertex coarse space of the two-level preconditioner ----------------------------------
// Ace = Phi_e^T S_e Phi_e, Phi_e = diag(free nodes) phi diag(free vertices).
template <int N>
__global__ void __launch_bounds__(kScThreads)
    sc_coarse_elem_kernel(int64_t n_elem, const double *__restrict__ S,
                          const uint32_t *__restrict__ l2g_ext,
                          const uint8_t *__restrict__ dirichlet, const double *__restrict__ phi,
                          double *__restrict__ Ace) {
  using C = ScCfg<N>;
  constexpr int NE = C::NE, NS = C::NS, EPB = C::EPB;
  extern __shared__ __align__(128) double sc_smem[];
  double *sS = sc_smem;                 // [EPB][NS]
  double *sP0 = sS + EPB * NS;          // [NE][4] phi
  double *sPhi = sP0 + NE * 4;          // [EPB][NE][4] masked Phi_e
  double *sT = sPhi + EPB * NE * 4;     // [EPB][NE][4] S_e Phi_e
  const int tid = threadIdx.x;
  const int le = tid / NE, k = tid - le * NE;
  for (int i = tid; i < NE * 4; i += kScThreads) sP0[i] = phi[i];
  const int64_t n_groups = (n_elem + EPB - 1) / EPB;
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    __syncthreads();
    const int64_t e0 = grp * EPB;
    const int ne = (int)((n_elem - e0) < (int64_t)EPB ? (n_elem - e0) : (int64_t)EPB);
    const double2 *src = reinterpret_cast<const double2 *>(S + e0 * NS);
    double2 *dst = reinterpret_cast<double2 *>(sS);
    for (int q = tid; q < ne * (NS / 2); q += kScThreads) dst[q] = src[q];
    const bool on = le < ne;
    if (on) {
      const uint32_t *ids = l2g_ext + (e0 + le) * NE;
      const double fn = (dirichlet && dirichlet[ids[k]]) ? 0.0 : 1.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double fv = (dirichlet && dirichlet[ids[c]]) ? 0.0 : 1.0;  // vertices: k = 0..3
        sPhi[(le * NE + k) * 4 + c] = sP0[k * 4 + c] * fn * fv;
      }
    }
    __syncthreads();
    if (on) {
      const double *s = sS + le * NS;
      const double *ph = sPhi + le * NE * 4;
      const int kk = k * (k + 1) / 2;
      double t[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < NE; ++j) {
        const double sv = s[(j <= k) ? kk + j : j * (j + 1) / 2 + k];
#pragma unroll
        for (int c = 0; c < 4; ++c) t[c] = fma(sv, ph[j * 4 + c], t[c]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) sT[(le * NE + k) * 4 + c] = t[c];
    }
    __syncthreads();
    for (int q = tid; q < ne * 16; q += kScThreads) {
      const int l2 = q >> 4, a = (q >> 2) & 3, c = q & 3;
      const double *ph = sPhi + l2 * NE * 4;
      const double *tt = sT + l2 * NE * 4;
      double acc = 0.0;
      for (int j = 0; j < NE; ++j) acc = fma(ph[j * 4 + a], tt[j * 4 + c], acc);
      Ace[(e0 + l2) * 16 + a * 4 + c] = acc;
    }
  }
}

template <int N>
constexpr size_t sc_coarse_elem_smem() {
  return sizeof(double) * ((size_t)ScCfg<N>::EPB * ScCfg<N>::NS + (size_t)ScCfg<N>::NE * 4 +
                           2 * (size_t)ScCfg<N>::EPB * ScCfg<N>::NE * 4);
}

// element part of the coarse apply: y_loc_c[e][a] = sum_c Ace[e][a][c] x[vert_c[e][c]]
__global__ void __launch_bounds__(256)
    sc_coarse_matvec_kernel(int64_t n_elem, const double *__restrict__ Ace,
                            const uint32_t *__restrict__ vert_c, const double *__restrict__ x,
                            double *__restrict__ y_loc_c) {
  const int64_t total = n_elem * 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i >> 2;
    const double *A = Ace + e * 16 + (i & 3) * 4;
    const uint32_t *v = vert_c + e * 4;
    double acc = A[0] * x[v[0]];
    acc = fma(A[1], x[v[1]], acc);
    acc = fma(A[2], x[v[2]], acc);
    acc = fma(A[3], x[v[3]], acc);
    y_loc_c[i] = acc;
  }
}

// ---- condensed apply / assembly, node part ------------------------------------------------
constexpr int kNodeThreads = 256;

__global__ void __launch_bounds__(kNodeThreads)
    sc_node_kernel(int64_t n_ext, const uint32_t *__restrict__ node_ptr,
                   const uint32_t *__restrict__ node_pos, const double *__restrict__ loc,
                   const uint8_t *__restrict__ dirichlet, const double *__restrict__ u,
                   double *__restrict__ y, int flags, double fill, double *__restrict__ partials,
                   double *__restrict__ dot_out) {
  double acc[1] = {0.0};
  const bool mask_out = dirichlet && (flags & SEMK_MASK_OUT);
  const bool identity = (flags & SEMK_DIRICHLET_IDENTITY) && u;
  for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < n_ext;
       g += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t q0 = node_ptr[g], q1 = node_ptr[g + 1];
    double s = 0.0;
    for (uint32_t q = q0; q < q1; ++q) s += loc[node_pos[q]];
    if (mask_out && dirichlet[g]) s = identity ? u[g] : fill;
    y[g] = s;
    if (dot_out) acc[0] = fma(u[g], s, acc[0]);
  }
  if (dot_out) {
    double tot[1];
    if (semk_finish_reduction<1>(acc, partials, tot) && threadIdx.x == 0) dot_out[0] = tot[0];
  }
}

inline unsigned node_blocks(int64_t n) {
  const int64_t want = (n + kNodeThreads - 1) / kNodeThreads;
  return (unsigned)(want < 1 ? 1 : (want < kSemkRedMaxBlocks ? want : kSemkRedMaxBlocks));
}

int check_sc_op(const semk_sc_op *op, const char *who) {
  if (!op) {
    semk_set_error(std::string(who) + ": null operator");
    return SEMK_ERR_INVALID;
  }
  if (op->n1 < 3 || op->n1 > 11) {
    semk_set_error(std::string(who) + ": static condensation supports orders 2..10");
    return SEMK_ERR_UNSUPPORTED;
  }
  const int ne = 4 * (op->n1 - 1);
  if (op->n_ext_loc != ne || op->s_stride != (int64_t)ne * (ne + 1) / 2 || op->n_elem <= 0 ||
      op->n_ext <= 0 || !op->S || !op->l2g_ext || !op->y_loc || !op->node_ptr ||
      !op->node_pos) {
    semk_set_error(std::string(who) + ": inconsistent semk_sc_op");
    return SEMK_ERR_INVALID;
  }
  if ((reinterpret_cast<uintptr_t>(op->S) & 15u) != 0) {
    semk_set_error(std::string(who) + ": S must be 16-byte aligned");
    return SEMK_ERR_INVALID;
  }
  return SEMK_OK;
}

}  // namespace

#define SEMK_DISPATCH_SC(n1, CALL)                                                   \
  switch (n1) {                                                                      \
    case 3: CALL(3); break;                                                          \
    case 4: CALL(4); break;                                                          \
    case 5: CALL(5); break;                                                          \
    case 6: CALL(6); break;                                                          \
    case 7: CALL(7); break;                                                          \
    case 8: CALL(8); break;                                                          \
    case 9: CALL(9); break;                                                          \
    case 10: CALL(10); break;                                                        \
    case 11: CALL(11); break;                                                        \
    default:                                                                         \
      semk_set_error("static condensation supports orders 2..10 (n1 in [3, 11])");   \
      return SEMK_ERR_UNSUPPORTED;                                                   \
  }

static int launch_sc_element(int n1, const ScElemArgs &a, void *stream);

extern "C" int semk_sc_element_f64(int n1, int64_t n_elem, const int64_t *slot_of_elem,
                                   const double *G, int64_t g_patch_stride, int elems_per_patch,
                                   const double *D, const int32_t *ext_loc, const uint32_t *l2g,
                                   const double *JxW, const double *f_nodal, double f_scale,
                                   int mode, double *S_out, int64_t s_stride, double *sdiag_loc,
                                   double *g_loc, double *u, double *W_out, double *c_out,
                                   int32_t *bad_flag, void *stream) {
  return semk_sc_element_react_f64(n1, n_elem, slot_of_elem, G, g_patch_stride, elems_per_patch,
                                   D, ext_loc, l2g, JxW, f_nodal, f_scale, mode, S_out, s_stride,
                                   sdiag_loc, g_loc, u, W_out, c_out, nullptr, nullptr, bad_flag,
                                   stream);
}

extern "C" int semk_sc_element_react_f64(int n1, int64_t n_elem, const int64_t *slot_of_elem,
                                         const double *G, int64_t g_patch_stride,
                                         int elems_per_patch, const double *D,
                                         const int32_t *ext_loc, const uint32_t *l2g,
                                         const double *JxW, const double *f_nodal,
                                         double f_scale, int mode, double *S_out,
                                         int64_t s_stride, double *sdiag_loc, double *g_loc,
                                         double *u, double *W_out, double *c_out,
                                         const double *react, double *Ainv_out,
                                         int32_t *bad_flag, void *stream) {
  SEMK_REQUIRE(n_elem > 0 && G && D && ext_loc && bad_flag &&
                   elems_per_patch > 0 && g_patch_stride > 0,
               "semk_sc_element_f64: bad argument");
  SEMK_REQUIRE(mode != 0 && (mode & ~31) == 0, "semk_sc_element_f64: bad mode");
  SEMK_REQUIRE(!(mode & SEMK_SC_STORE_INV) || Ainv_out,
               "semk_sc_element_f64: STORE_INV needs Ainv_out");
  SEMK_REQUIRE(!(mode & SEMK_SC_STORE) || W_out, "semk_sc_element_f64: STORE needs W_out");
  SEMK_REQUIRE(!c_out || (mode & (SEMK_SC_RHS | SEMK_SC_BACKSOLVE)),
               "semk_sc_element_f64: c_out needs a load (RHS or BACKSOLVE mode)");
  if (n1 < 3 || n1 > 11) {
    semk_set_error("semk_sc_element_f64: static condensation supports orders 2..10");
    return SEMK_ERR_UNSUPPORTED;
  }
  const int ne = 4 * (n1 - 1);
  if (mode & SEMK_SC_SCHUR)
    SEMK_REQUIRE(S_out && s_stride == (int64_t)ne * (ne + 1) / 2,
                 "semk_sc_element_f64: SCHUR needs S_out and the packed stride");
  if (mode & SEMK_SC_RHS)
    SEMK_REQUIRE(g_loc && JxW && l2g, "semk_sc_element_f64: RHS needs g_loc, JxW, l2g");
  if (mode & SEMK_SC_BACKSOLVE)
    SEMK_REQUIRE(u && JxW && l2g, "semk_sc_element_f64: BACKSOLVE needs u, JxW, l2g");
  ScElemArgs a;
  a.n_elem = n_elem;
  a.slot_of_elem = slot_of_elem;
  a.G = G;
  a.g_patch_stride = g_patch_stride;
  a.pe = elems_per_patch;
  a.D = D;
  a.ext_loc = ext_loc;
  a.l2g = l2g;
  a.JxW = JxW;
  a.f_nodal = f_nodal;
  a.f_scale = f_scale;
  a.mode = mode;
  a.S_out = S_out;
  a.s_stride = s_stride;
  a.sdiag_loc = sdiag_loc;
  a.g_loc = g_loc;
  a.u = u;
  a.bad_flag = bad_flag;
  a.W_out = W_out;
  a.c_out = c_out;
  a.A_dense = nullptr;
  a.f_dense = nullptr;
  a.l2g_hier = nullptr;
  a.react = react;
  a.Ainv_out = Ainv_out;
  return launch_sc_element(n1, a, stream);
}

extern "C" int semk_sc_element_dense_f64(int n1, int64_t n_elem, const double *A_hier,
                                         const double *f_hier, const uint32_t *l2g_hier, int mode,
                                         double *S_out, int64_t s_stride, double *sdiag_loc,
                                         double *g_loc, double *u, int32_t *bad_flag,
                                         void *stream) {
  SEMK_REQUIRE(n_elem > 0 && A_hier && bad_flag, "semk_sc_element_dense_f64: bad argument");
  SEMK_REQUIRE(mode != 0 && (mode & ~7) == 0, "semk_sc_element_dense_f64: bad mode");
  if (n1 < 3 || n1 > 11) {
    semk_set_error("semk_sc_element_dense_f64: static condensation supports orders 2..10");
    return SEMK_ERR_UNSUPPORTED;
  }
  const int ne = 4 * (n1 - 1);
  if (mode & SEMK_SC_SCHUR)
    SEMK_REQUIRE(S_out && s_stride == (int64_t)ne * (ne + 1) / 2,
                 "semk_sc_element_dense_f64: SCHUR needs S_out and the packed stride");
  if (mode & SEMK_SC_RHS)
    SEMK_REQUIRE(g_loc && f_hier, "semk_sc_element_dense_f64: RHS needs g_loc and f_hier");
  if (mode & SEMK_SC_BACKSOLVE)
    SEMK_REQUIRE(u && f_hier && l2g_hier,
                 "semk_sc_element_dense_f64: BACKSOLVE needs u, f_hier, l2g_hier");
  ScElemArgs a;
  a.n_elem = n_elem;
  a.slot_of_elem = nullptr;
  a.G = nullptr;
  a.g_patch_stride = 0;
  a.pe = 1;
  a.D = nullptr;
  a.ext_loc = nullptr;
  a.l2g = nullptr;
  a.JxW = nullptr;
  a.f_nodal = nullptr;
  a.f_scale = 1.0;
  a.mode = mode;
  a.S_out = S_out;
  a.s_stride = s_stride;
  a.sdiag_loc = sdiag_loc;
  a.g_loc = g_loc;
  a.u = u;
  a.bad_flag = bad_flag;
  a.W_out = nullptr;
  a.c_out = nullptr;
  a.A_dense = A_hier;
  a.f_dense = f_hier;
  a.l2g_hier = l2g_hier;
  a.react = nullptr;
  a.Ainv_out = nullptr;
  return launch_sc_element(n1, a, stream);
}

static int launch_sc_element(int n1, const ScElemArgs &a, void *stream) {
  const int64_t n_elem = a.n_elem;
  cudaStream_t st = semk_stream(stream);
  const unsigned grid = (unsigned)(n_elem < 148 * 16 ? n_elem : 148 * 16);
#define SEMK_CALL(NV)                                                                      \
  do {                                                                                     \
    SEMK_CUDA_CHECK(cudaFuncSetAttribute(sc_element_kernel<NV>,                       \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                         (int)ScCfg<NV>::kSmem));                          \
    sc_element_kernel<NV><<<grid, kScThreads, ScCfg<NV>::kSmem, st>>>(a);             \
  } while (0)
  SEMK_DISPATCH_SC(n1, SEMK_CALL)
#undef SEMK_CALL
  SEMK_LAUNCH_CHECK("sc_element_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_backsolve_stored_f64(int n1, int64_t n_elem, const double *W,
                                           const double *c, const uint32_t *l2g,
                                           const int32_t *ext_loc, double *u, void *stream) {
  SEMK_REQUIRE(n_elem > 0 && W && l2g && ext_loc && u, "semk_sc_backsolve_stored_f64: bad argument");
  cudaStream_t st = semk_stream(stream);
#define SEMK_CALL(NV)                                                                      \
  do {                                                                                     \
    constexpr int GT = ((ScCfg<NV>::NI + 31) / 32) * 32, GPB = 256 / GT;                   \
    const int64_t steps = (n_elem + GPB - 1) / GPB;                                        \
    const unsigned grid = (unsigned)(steps < 148 * 16 ? steps : 148 * 16);                 \
    sc_backsolve_stored_kernel<NV><<<grid, 256, 0, st>>>(n_elem, W, c, l2g, ext_loc, u);   \
  } while (0)
  SEMK_DISPATCH_SC(n1, SEMK_CALL)
#undef SEMK_CALL
  SEMK_LAUNCH_CHECK("sc_backsolve_stored_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_load_stored_f64(int n1, int64_t n_elem, const double *W, const double *Ainv,
                                       const uint32_t *l2g, const int32_t *ext_loc,
                                       const double *JxW, const double *f_nodal, double f_scale,
                                       double *g_loc, double *c_out, void *stream) {
  SEMK_REQUIRE(n_elem > 0 && W && Ainv && l2g && ext_loc && JxW && g_loc && c_out,
               "semk_sc_load_stored_f64: bad argument");
  cudaStream_t st = semk_stream(stream);
#define SEMK_CALL(NV)                                                                      \
  do {                                                                                     \
    constexpr int GT = ((ScCfg<NV>::NI + 31) / 32) * 32, GPB = 256 / GT;                   \
    const int64_t steps = (n_elem + GPB - 1) / GPB;                                        \
    const unsigned grid = (unsigned)(steps < 148 * 16 ? steps : 148 * 16);                 \
    sc_load_stored_kernel<NV><<<grid, 256, 0, st>>>(n_elem, W, Ainv, l2g, ext_loc, JxW,    \
                                                    f_nodal, f_scale, g_loc, c_out);       \
  } while (0)
  SEMK_DISPATCH_SC(n1, SEMK_CALL)
#undef SEMK_CALL
  SEMK_LAUNCH_CHECK("sc_load_stored_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_apply_f64(const semk_sc_op *op, const double *u, double *y, int flags,
                                 double *dot_out, void *stream) {
  int rc = check_sc_op(op, "semk_sc_apply_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(u && y && u != y, "semk_sc_apply_f64: null / aliased vectors");
  SEMK_REQUIRE(!dot_out || op->partials, "semk_sc_apply_f64: dot_out needs op->partials");
  cudaStream_t st = semk_stream(stream);
  const uint8_t *mask_in = (op->dirichlet && (flags & SEMK_MASK_IN)) ? op->dirichlet : nullptr;
  int n_sm = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess)
      (void)cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
#define SEMK_CALL(NV)                                                                      \
  do {                                                                                     \
    const int64_t groups = (op->n_elem + ScCfg<NV>::EPB - 1) / ScCfg<NV>::EPB;             \
    {                                                                                      \
      int per_sm = 0;                                                                      \
      SEMK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(                       \
          &per_sm, sc_matvec_kernel<NV>, kScThreads, sc_matvec_smem<NV>()));       \
      if (per_sm < 1) per_sm = 1;                                                          \
      const int64_t cap = (int64_t)n_sm * per_sm;                                          \
      const unsigned grid = (unsigned)(groups < cap ? groups : cap);                       \
      sc_matvec_kernel<NV><<<grid, kScThreads, sc_matvec_smem<NV>(), st>>>(         \
          op->n_elem, op->S, op->l2g_ext, u, mask_in, op->y_loc);                          \
    }                                                                                      \
  } while (0)
  SEMK_DISPATCH_SC(op->n1, SEMK_CALL)
#undef SEMK_CALL
  SEMK_LAUNCH_CHECK("sc_matvec_kernel");
  sc_node_kernel<<<node_blocks(op->n_ext), kNodeThreads, 0, st>>>(
      op->n_ext, op->node_ptr, op->node_pos, op->y_loc, op->dirichlet, u, y, flags, 0.0,
      op->partials, dot_out);
  SEMK_LAUNCH_CHECK("sc_node_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_assemble_f64(const semk_sc_op *op, const double *loc, double *out,
                                    int flags, double fill_dirichlet, void *stream) {
  int rc = check_sc_op(op, "semk_sc_assemble_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(loc && out, "semk_sc_assemble_f64: null pointer");
  sc_node_kernel<<<node_blocks(op->n_ext), kNodeThreads, 0, semk_stream(stream)>>>(
      op->n_ext, op->node_ptr, op->node_pos, loc, op->dirichlet, nullptr, out,
      flags & SEMK_MASK_OUT, fill_dirichlet, nullptr, nullptr);
  SEMK_LAUNCH_CHECK("sc_node_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_coarse_elem_f64(const semk_sc_op *op, const double *phi, double *Ace_out,
                                       void *stream) {
  int rc = check_sc_op(op, "semk_sc_coarse_elem_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(phi && Ace_out, "semk_sc_coarse_elem_f64: null pointer");
  cudaStream_t st = semk_stream(stream);
#define SEMK_CALL(NV)                                                                      \
  do {                                                                                     \
    const int64_t groups = (op->n_elem + ScCfg<NV>::EPB - 1) / ScCfg<NV>::EPB;             \
    const unsigned grid = (unsigned)(groups < 148 * 8 ? groups : 148 * 8);                 \
    sc_coarse_elem_kernel<NV><<<grid, kScThreads, sc_coarse_elem_smem<NV>(), st>>>(        \
        op->n_elem, op->S, op->l2g_ext, op->dirichlet, phi, Ace_out);                      \
  } while (0)
  SEMK_DISPATCH_SC(op->n1, SEMK_CALL)
#undef SEMK_CALL
  SEMK_LAUNCH_CHECK("sc_coarse_elem_kernel");
  return SEMK_OK;
}

static int check_coarse(int64_t n_elem, const semk_sc_coarse *cs, const char *who) {
  if (!cs || n_elem <= 0 || cs->n_v <= 0 || !cs->Ace || !cs->vert_c || !cs->y_loc_c || !cs->vptr ||
      !cs->vpos) {
    semk_set_error(std::string(who) + ": inconsistent semk_sc_coarse");
    return SEMK_ERR_INVALID;
  }
  return SEMK_OK;
}

extern "C" int semk_sc_coarse_apply_f64(int64_t n_elem, const semk_sc_coarse *cs, const double *x,
                                        double *y, int flags, double *dot_out, void *stream) {
  int rc = check_coarse(n_elem, cs, "semk_sc_coarse_apply_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(x && y && x != y, "semk_sc_coarse_apply_f64: null / aliased vectors");
  SEMK_REQUIRE(!dot_out || cs->partials, "semk_sc_coarse_apply_f64: dot_out needs partials");
  cudaStream_t st = semk_stream(stream);
  const int64_t want = (n_elem * 4 + 255) / 256;
  const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
  sc_coarse_matvec_kernel<<<grid, 256, 0, st>>>(n_elem, cs->Ace, cs->vert_c, x, cs->y_loc_c);
  SEMK_LAUNCH_CHECK("sc_coarse_matvec_kernel");
  sc_node_kernel<<<node_blocks(cs->n_v), kNodeThreads, 0, st>>>(
      cs->n_v, cs->vptr, cs->vpos, cs->y_loc_c, cs->dirichlet_c, x, y, flags, 0.0, cs->partials,
      dot_out);
  SEMK_LAUNCH_CHECK("sc_node_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_coarse_assemble_f64(int64_t n_elem, const semk_sc_coarse *cs,
                                           const double *loc, double *out, void *stream) {
  int rc = check_coarse(n_elem, cs, "semk_sc_coarse_assemble_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(loc && out, "semk_sc_coarse_assemble_f64: null pointer");
  sc_node_kernel<<<node_blocks(cs->n_v), kNodeThreads, 0, semk_stream(stream)>>>(
      cs->n_v, cs->vptr, cs->vpos, loc, nullptr, nullptr, out, 0, 0.0, nullptr, nullptr);
  SEMK_LAUNCH_CHECK("sc_node_kernel");
  return SEMK_OK;
}
