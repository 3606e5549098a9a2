// semk_stokes.cu -- axisymmetric Stokes / linearised Navier-Stokes in stream function -
// vorticity form, matrix-free, two DOFs per node (BASELINE config 4).
//
// Reference being replaced (examples/squirmer-axisymmetric.py, per element, in Python):
//   pre_assembly (:163-257): dense 4-index local operators E2e, Lve [N,N,N,N] from four
//     O(N^5) einsums with rho*JxW, the Kronecker-sparse advection operator Ae and the
//     diagonal mass operator Me;
//   compute_local_system (:259-297): local Jacobian blocks
//       jac[0::2,0::2] = Ae.vort          jac[0::2,1::2] = Ae.sfn + Lve
//       jac[1::2,0::2] = E2e              jac[1::2,1::2] = -Me
//     and the residual res[0::2] = (Ae.vort).sfn + Lve vort, res[1::2] = E2e sfn - Me vort;
//   DOF interleaving dof = 2*node + comp (sem/discrete.py:561-576), scatter-add
//     (:353, sem/discrete.py:499).
//
// None of the dense operators is formed.  With a_i = invJ[0][i], b_i = invJ[1][i] (rows of
// d xi / d x), psi_0 = D psi, psi_1 = psi D^T (parametric derivatives) the local Jacobian
// applied to (psi, omega) is, node by node,
//   row 0 (vorticity transport):  K omega + dL omega
//                                 + e0 psi_0 + e1 psi_1 + f0 omega_0 + f1 omega_1 + f2 omega
//   row 1 (vorticity definition): K psi + c0 psi_0 + c1 psi_1 - dM omega
// K = rho-weighted stiffness (factors G00, G01, G11 = rho JxW sum_i a_i a_i / a_i b_i /
// b_i b_i, :194-207), c0 = 2 JxW a_0, c1 = 2 JxW b_0 (:221-222), dL = JxW / rho (:211; 0 on
// the axis of symmetry, where both fields are essential), dM = rho^2 JxW (:252), and the
// five advection coefficients of the linearisation about a state (Psi, Omega) (:227-249):
//   e0 = Re (q Omega_1 + s0 Omega)   e1 = Re (-q Omega_0 + s1 Omega)
//   f0 = -Re q Psi_1                 f1 = Re q Psi_0          f2 = Re (s0 Psi_0 + s1 Psi_1)
// with q = JxW (a_0 b_1 - a_1 b_0) and s0 = dL a_1, s1 = dL b_1.  The advection term is
// bilinear, so the nonlinear residual is the same kernel with the e/f terms halved and the
// state as the input.
//
// Kernel layout: the persistent patch kernel of semk_apply.cu (same plan tables, TMA-staged
// factor block, gather-style assembly without atomics, interface slots) carrying two
// fields: nodal values are gathered as 16-byte (psi, omega) pairs, the element operator runs
// once per field through the same transpose scratch, results are assembled from two result
// buffers with one pass over the inverse table and written as 16-byte pairs.
#include "semk_elem.cuh"

namespace {

constexpr int kFacStokes = 7;    // G00 G01 G11 c0 c1 dL dM
constexpr int kFacAdv = 12;      // + e0 e1 f0 f1 f2

struct StokesSmem {
  size_t hdr, fs, stage0, stage_bytes, pn_off, el_off, inv, a, b, r0, r1, total;
};
__host__ __device__ inline StokesSmem stokes_smem_layout(int N, int PE, int64_t f_patch_stride,
                                                         int64_t pn_patch_stride,
                                                         int64_t eloc_patch_stride,
                                                         int64_t inv_patch_stride) {
  StokesSmem L;
  const size_t scratch = 8 * (size_t)N * scratch_row_stride(N, PE);
  size_t o = 32;  // mbarriers: tables[2], factors, inverse table
  L.hdr = o;
  o += 64;
  L.fs = o;
  o += sizeof(double) * (size_t)f_patch_stride;
  L.stage0 = o;
  L.pn_off = 0;
  size_t st = 4 * (size_t)pn_patch_stride;
  L.el_off = st;
  st += 2 * (size_t)eloc_patch_stride;
  st = (st + 15) & ~(size_t)15;
  L.stage_bytes = st;
  o += 2 * st;
  L.inv = o;
  o += 2 * (size_t)inv_patch_stride;
  o = (o + 15) & ~(size_t)15;
  L.a = o;
  o += scratch;
  o = (o + 15) & ~(size_t)15;
  L.b = o;
  o += scratch;
  o = (o + 15) & ~(size_t)15;
  L.r0 = o;
  o += scratch;
  o = (o + 15) & ~(size_t)15;
  L.r1 = o;
  o += scratch > 256 ? scratch : 256;
  L.total = o;
  return L;
}

// One field through the element operator.  On entry col[m] = field[m][t]; on exit
// R[m][tid] (+)= (K field)[m][t] + extra[m], where `point(m, ur_m, us_m)` returns the
// collocated extra of node (m, t) from the two parametric derivatives (and may add to the
// other field's result buffer).  Four barriers; every thread of the CTA must call it.
template <int N, int RS, class Point, class Hook0, class Hook1>
__device__ __forceinline__ void stokes_field_pass(const DMatEO &dm, int le, int t, bool active,
                                                  const double (&col)[N], double *__restrict__ A,
                                                  double *__restrict__ B, double *__restrict__ R,
                                                  bool accumulate, const double *__restrict__ f,
                                                  int f_row, uint64_t *f_ready, uint32_t f_parity,
                                                  Point point, Hook0 after_first_barrier,
                                                  Hook1 factors_consumed) {
  const int tidp = le * N + t;
  double ur[N], tmp[N], us[N];
  if (active) {
#pragma unroll
    for (int m = 0; m < N; ++m) A[m * RS + tidp] = col[m];
  }
  __syncthreads();
  after_first_barrier();
  if (active) {
    mat_D<N>(dm, col, ur);
#pragma unroll
    for (int s = 0; s < N; ++s) tmp[s] = A[t * RS + le * N + s];
    mat_D<N>(dm, tmp, us);
#pragma unroll
    for (int n = 0; n < N; ++n) B[t * RS + le * N + n] = us[n];
  }
  __syncthreads();
  if (f_ready) semk_mbar_wait(f_ready, f_parity);
  double ycol[N], w1[N], ex[N];
  if (active) {
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const double usc = B[m * RS + tidp];
      const double g00 = f[m * f_row], g01 = f[(N + m) * f_row], g11 = f[(2 * N + m) * f_row];
      tmp[m] = g00 * ur[m] + g01 * usc;
      w1[m] = g01 * ur[m] + g11 * usc;
      ex[m] = point(m, ur[m], usc);
    }
    mat_Dt<N>(dm, tmp, ycol);
#pragma unroll
    for (int m = 0; m < N; ++m) A[m * RS + tidp] = w1[m];
  }
  __syncthreads();
  factors_consumed();
  if (active) {
#pragma unroll
    for (int n = 0; n < N; ++n) tmp[n] = A[t * RS + le * N + n];
    mat_Dt<N>(dm, tmp, us);
#pragma unroll
    for (int q = 0; q < N; ++q) B[t * RS + le * N + q] = us[q];
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const double v = ycol[m] + ex[m] + B[m * RS + tidp];
      if (accumulate)
        R[m * RS + tidp] += v;
      else
        R[m * RS + tidp] = v;
    }
  }
}

struct Nop {
  __device__ __forceinline__ void operator()() const {}
};

constexpr int kStokesGatherBatch = 4;

__host__ __device__ constexpr int stokes_threads(int N, int PE) { return ((N * PE + 31) / 32) * 32; }

// resident CTAs per SM the register budget is set for (<= 224 registers at 96 threads)
__host__ __device__ constexpr int stokes_min_blocks(int N, int PE) {
  return stokes_threads(N, PE) <= 96 && N <= 9 ? 3 : (stokes_threads(N, PE) <= 128 ? 2 : 1);
}

template <int N, int PE, bool ADV>
__global__ void __launch_bounds__(stokes_threads(N, PE), stokes_min_blocks(N, PE))
    stokes_patch_kernel(semk_op op, DMatEO dm, const double2 *__restrict__ u,
                        double2 *__restrict__ y, double adv_scale) {
  constexpr int NP = N * PE;
  constexpr int kThreads = stokes_threads(N, PE);
  constexpr int RS = scratch_row_stride(N, PE);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const StokesSmem L = stokes_smem_layout(N, PE, op.g_patch_stride, op.pn_patch_stride,
                                          op.eloc_patch_stride, op.inv_patch_stride);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);
  double *Fs = reinterpret_cast<double *>(smem_raw + L.fs);
  const uint16_t *inv_s = reinterpret_cast<const uint16_t *>(smem_raw + L.inv);
  double *As = reinterpret_cast<double *>(smem_raw + L.a);
  double *Bs = reinterpret_cast<double *>(smem_raw + L.b);
  double *R0 = reinterpret_cast<double *>(smem_raw + L.r0);
  double *R1 = reinterpret_cast<double *>(smem_raw + L.r1);
  double2 *slots = reinterpret_cast<double2 *>(op.slot_buf);

  const int tid = threadIdx.x;
  const int le = tid / N, t = tid - le * N;
  const uint32_t pn_bytes = 4u * (uint32_t)op.pn_patch_stride;
  const uint32_t el_bytes = 2u * (uint32_t)op.eloc_patch_stride;
  const uint32_t f_bytes = (uint32_t)(op.g_patch_stride * sizeof(double));
  const uint32_t inv_bytes = 2u * (uint32_t)op.inv_patch_stride;
  const int inv_w4 = (int)(op.inv_width >> 2);

  auto stage_ptr = [&](int s) { return smem_raw + L.stage0 + (size_t)s * L.stage_bytes; };
  uint32_t *hdr_ring = reinterpret_cast<uint32_t *>(smem_raw + L.hdr);
  auto issue_tables = [&](uint32_t pi, uint32_t ei, int s, int64_t patch_after, int slot_after) {
    unsigned char *base = stage_ptr(s);
    const bool more = patch_after >= 0;
    semk_mbar_expect_tx(&mbar[s], pn_bytes + el_bytes + (more ? 32u : 0u));
    semk_bulk_g2s(base + L.pn_off, op.pnode + (int64_t)pi * op.pn_patch_stride, pn_bytes, &mbar[s]);
    semk_bulk_g2s(base + L.el_off, op.eloc + (int64_t)ei * op.eloc_patch_stride, el_bytes, &mbar[s]);
    if (more) semk_bulk_g2s(hdr_ring + 8 * slot_after, op.patch_hdr + 8 * patch_after, 32u, &mbar[s]);
  };
  auto issue_inv = [&](uint32_t block) {
    semk_mbar_expect_tx(&mbar[3], inv_bytes);
    semk_bulk_g2s(smem_raw + L.inv, op.inv + (int64_t)block * op.inv_patch_stride, inv_bytes,
                  &mbar[3]);
  };
  auto issue_f = [&](int64_t patch) {
    semk_mbar_expect_tx(&mbar[2], f_bytes);
    semk_bulk_g2s(Fs, op.G + patch * op.g_patch_stride, f_bytes, &mbar[2]);
  };

  const int64_t step = (int64_t)gridDim.x;
  const int64_t p_first = (int64_t)blockIdx.x;
  const int64_t p_end = op.n_patch;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) semk_mbar_init(&mbar[i], 1);
    semk_fence_mbar_init();
    if (p_first < p_end) {
      const uint32_t pi = op.patch_hdr[8 * p_first + 5], ei = op.patch_hdr[8 * p_first + 6];
      const int64_t p2 = p_first + step;
      unsigned char *base = stage_ptr(0);
      semk_mbar_expect_tx(&mbar[0], pn_bytes + el_bytes + 32u + (p2 < p_end ? 32u : 0u));
      semk_bulk_g2s(base + L.pn_off, op.pnode + (int64_t)pi * op.pn_patch_stride, pn_bytes, &mbar[0]);
      semk_bulk_g2s(base + L.el_off, op.eloc + (int64_t)ei * op.eloc_patch_stride, el_bytes, &mbar[0]);
      semk_bulk_g2s(hdr_ring, op.patch_hdr + 8 * p_first, 32u, &mbar[0]);
      if (p2 < p_end) semk_bulk_g2s(hdr_ring + 8, op.patch_hdr + 8 * p2, 32u, &mbar[0]);
      issue_f(p_first);
    }
  }
  __syncthreads();

  // (psi, omega) columns of the patch about to be processed, gathered as 16-byte pairs and
  // carried in registers: the gather for patch i+1 is issued after the element operator of
  // patch i and lands during its assembly and write-out.
  double pcol[N], wcol[N];
  auto gather_columns = [&](int s_tab, int64_t patch_of) {
    const unsigned char *sbn = stage_ptr(s_tab);
    const uint32_t *pnb = reinterpret_cast<const uint32_t *>(sbn + L.pn_off);
    const uint32_t id0 = hdr_ring[8 * s_tab + 4];
    const uint16_t *elb = reinterpret_cast<const uint16_t *>(sbn + L.el_off);
    const bool act = (le < PE) && (patch_of * PE + le < op.n_elem);
    if (act) {
      uint32_t pn[N];
#pragma unroll
      for (int m = 0; m < N; ++m) pn[m] = pnb[elb[m * NP + tid]];
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const double2 v = u[id0 + (pn[m] & SEMK_NODE_ID_MASK)];
        pcol[m] = v.x;
        wcol[m] = v.y;
      }
    }
  };
  if (p_first < p_end) {
    semk_mbar_wait(&mbar[0], 0);
    gather_columns(0, p_first);
  }
  __syncthreads();

  int it = 0;
  for (int64_t patch = p_first; patch < p_end; patch += step, ++it) {
    const int s = it & 1;
    const uint32_t par = (uint32_t)((it >> 1) & 1);
    const int64_t next = patch + step;
    const bool has_next = next < p_end;
    unsigned char *sb = stage_ptr(s);
    const uint32_t *pn_s = reinterpret_cast<const uint32_t *>(sb + L.pn_off);
    semk_mbar_wait(&mbar[s], par);
    const uint32_t *hdr = hdr_ring + 8 * s;
    const int npn = (int)hdr[0];
    const int npriv = (int)hdr[1];
    const int slot_base = (int)hdr[2];
    const uint32_t id0 = hdr[4];
    const uint32_t inv_block = hdr[7];
    const int64_t slot0 = patch * PE;
    const bool active = (le < PE) && (slot0 + le < op.n_elem);
    const double *f = Fs + tid;

    auto refill_tables = [&]() {
      if (tid == 0) issue_inv(inv_block);
      if (tid == 0 && has_next) {
        const uint32_t *hn = hdr_ring + 8 * (s ^ 1);
        const int64_t after = next + step;
        issue_tables(hn[5], hn[6], s ^ 1, after < p_end ? after : -1, s);
      }
    };
    auto refill_f = [&]() {
      if (tid == 0 && has_next) issue_f(next);
    };
    const int tidp = le * N + t;
    // ---- field 0 (stream function) -> row 1, and its advection share of row 0 ----------
    auto point_psi = [&](int m, double ur, double usc) {
      if (ADV)
        R0[m * RS + tidp] = adv_scale * (f[(7 * N + m) * NP] * ur + f[(8 * N + m) * NP] * usc);
      return f[(3 * N + m) * NP] * ur + f[(4 * N + m) * NP] * usc - f[(6 * N + m) * NP] * wcol[m];
    };
    stokes_field_pass<N, RS>(dm, le, t, active, pcol, As, Bs, R1, false, f, NP, &mbar[2],
                             (uint32_t)(it & 1), point_psi, refill_tables, Nop());
    // ---- field 1 (vorticity) -> row 0 -----------------------------------------------------
    auto point_om = [&](int m, double ur, double usc) {
      double e = f[(5 * N + m) * NP] * wcol[m];
      if (ADV)
        e += adv_scale * (f[(9 * N + m) * NP] * ur + f[(10 * N + m) * NP] * usc +
                          f[(11 * N + m) * NP] * wcol[m]);
      return e;
    };
    stokes_field_pass<N, RS>(dm, le, t, active, wcol, As, Bs, R0, ADV, f, NP, nullptr, 0u, point_om,
                             Nop(), refill_f);
    if (has_next) {
      semk_mbar_wait(&mbar[s ^ 1], (uint32_t)(((it + 1) >> 1) & 1));
      gather_columns(s ^ 1, next);
    }
    // ---- assemble by gathering from the two result buffers, fixed order ----------------
    semk_mbar_wait(&mbar[3], (uint32_t)(it & 1));
    __syncthreads();
    for (int k0 = tid; k0 < npn; k0 += kStokesGatherBatch * kThreads) {
      uint32_t pnv[kStokesGatherBatch];
      uint2 ev[kStokesGatherBatch];
#pragma unroll
      for (int j = 0; j < kStokesGatherBatch; ++j) {
        const int k = k0 + j * kThreads;
        const bool in = k < npn;
        pnv[j] = in ? pn_s[k] : 0xffffffffu;
        ev[j] = in ? reinterpret_cast<const uint2 *>(inv_s)[(size_t)k * inv_w4]
                   : make_uint2(0xffffffffu, 0xffffffffu);
      }
#pragma unroll
      for (int j = 0; j < kStokesGatherBatch; ++j) {
        const int k = k0 + j * kThreads;
        if (k >= npn) continue;
        const uint32_t q[4] = {ev[j].x & 0xffffu, ev[j].x >> 16, ev[j].y & 0xffffu, ev[j].y >> 16};
        double v0 = 0.0, v1 = 0.0;
        bool first = true;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (q[i] != 0xffffu) {
            v0 = first ? R0[q[i]] : v0 + R0[q[i]];
            v1 = first ? R1[q[i]] : v1 + R1[q[i]];
            first = false;
          }
        for (int w = 1; w < inv_w4; ++w) {
          const uint2 e = reinterpret_cast<const uint2 *>(inv_s)[(size_t)k * inv_w4 + w];
          const uint32_t r[4] = {e.x & 0xffffu, e.x >> 16, e.y & 0xffffu, e.y >> 16};
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (r[i] != 0xffffu) {
              v0 += R0[r[i]];
              v1 += R1[r[i]];
            }
        }
        const uint32_t pn = pnv[j];
        if (k < npriv)
          y[id0 + (pn & SEMK_NODE_ID_MASK)] = make_double2(v0, v1);
        else
          slots[slot_base + (k - npriv)] = make_double2(v0, v1);
      }
    }
    // (the next patch's first operator barrier orders this write-out before any overwrite)
  }
}

// Interface reduction for two fields: the tables of semk_apply.cu's shared_nodes_kernel
// (affine chunks of two-patch nodes, per-node records for 3+ patches) with 16-byte slots.
__global__ void __launch_bounds__(256)
    stokes_shared_nodes_kernel(semk_op op, double2 *__restrict__ y, int chunk_blocks) {
  const double2 *slots = reinterpret_cast<const double2 *>(op.slot_buf);
  if ((int)blockIdx.x < chunk_blocks) {
    const int lane = threadIdx.x & 31;
    const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= op.n_shared_chunk) return;
    const uint4 c0 = reinterpret_cast<const uint4 *>(op.shared_chunk)[2 * c];
    const uint4 c1 = reinterpret_cast<const uint4 *>(op.shared_chunk)[2 * c + 1];
    if (lane < (int)c1.z) {
      const double2 a = slots[c0.z + (uint32_t)lane * c0.w];
      const double2 b = slots[c1.x + (uint32_t)lane * c1.y];
      y[c0.x + (uint32_t)lane * c0.y] = make_double2(a.x + b.x, a.y + b.y);
    }
  } else {
    const int64_t nb = (int64_t)gridDim.x - chunk_blocks;
    const int64_t stride = nb * blockDim.x;
    for (int64_t i = ((int64_t)blockIdx.x - chunk_blocks) * blockDim.x + threadIdx.x;
         i < op.n_shared; i += stride) {
      const uint4 r0 = reinterpret_cast<const uint4 *>(op.shared_rec)[2 * i];
      const uint4 r1 = reinterpret_cast<const uint4 *>(op.shared_rec)[2 * i + 1];
      const uint32_t cnt = r0.y;
      const uint32_t sl[6] = {r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      double2 v = slots[sl[0]];
      for (uint32_t j = 1; j < 6; ++j)
        if (j < cnt && !(cnt > 6 && j == 5)) {
          const double2 a = slots[sl[j]];
          v.x += a.x;
          v.y += a.y;
        }
      if (cnt > 6) {
        const uint32_t *ext = op.shared_ext + r1.w;
        for (uint32_t j = 5; j < cnt; ++j) {
          const double2 a = slots[ext[j - 5]];
          v.x += a.x;
          v.y += a.y;
        }
      }
      y[r0.x & SEMK_NODE_ID_MASK] = v;
    }
  }
}

// Factor block from the geometry in the reference's layouts.
__global__ void stokes_factors_kernel(int n1, int pe, int64_t n_slot_elems,
                                      const double *__restrict__ invJ,
                                      const double *__restrict__ JxW,
                                      const double *__restrict__ x_phys,
                                      const int64_t *__restrict__ elem_of_slot,
                                      double *__restrict__ F, int64_t f_patch_stride) {
  const int NN = n1 * n1, NP = n1 * pe;
  const int64_t total = n_slot_elems * NN;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t slot = i / NN;
    const int k = (int)(i - slot * NN);
    const int m = k / n1, t = k - m * n1;
    const int64_t e = elem_of_slot ? elem_of_slot[slot] : slot;
    const int64_t patch = slot / pe;
    const int lp = (int)(slot - patch * pe);
    double *f = F + patch * f_patch_stride + lp * n1 + t;
    double v[kFacStokes] = {0, 0, 0, 0, 0, 0, 0};
    if (e >= 0) {
      const double *ij = invJ + e * 4 * NN + k;
      const double a0 = ij[0], a1 = ij[NN], b0 = ij[2 * NN], b1 = ij[3 * NN];
      const double jw = JxW[e * NN + k], rho = x_phys[e * 2 * NN + k];
      const double rj = rho * jw;
      v[0] = rj * (a0 * a0 + a1 * a1);
      v[1] = rj * (a0 * b0 + a1 * b1);
      v[2] = rj * (b0 * b0 + b1 * b1);
      v[3] = 2.0 * (jw * a0);
      v[4] = 2.0 * (jw * b0);
      v[5] = rho > 0.0 ? jw / rho : 0.0;
      v[6] = rj * rho;
    }
#pragma unroll
    for (int c = 0; c < kFacStokes; ++c) f[(c * n1 + m) * NP] = v[c];
  }
}

// Advection coefficients of the linearisation about `state` (one CTA per element slot,
// one thread per node; derivatives through shared memory).
__global__ void stokes_linearize_kernel(int n1, int pe, const double *__restrict__ D,
                                        const double *__restrict__ invJ,
                                        const double *__restrict__ JxW,
                                        const double *__restrict__ x_phys,
                                        const uint32_t *__restrict__ l2g,
                                        const int64_t *__restrict__ elem_of_slot,
                                        const double2 *__restrict__ state, double n_rey,
                                        double *__restrict__ F, int64_t f_patch_stride) {
  extern __shared__ double sh[];
  const int NN = n1 * n1, NP = n1 * pe;
  double *P = sh, *W = sh + NN, *Ds = sh + 2 * NN;
  const int64_t slot = blockIdx.x;
  const int64_t e = elem_of_slot ? elem_of_slot[slot] : slot;
  const int k = threadIdx.x;
  const int64_t patch = slot / pe;
  const int lp = (int)(slot - patch * pe);
  if (k < NN) Ds[k] = D[k];
  if (e >= 0 && k < NN) {
    const double2 v = state[l2g[e * NN + k]];
    P[k] = v.x;
    W[k] = v.y;
  }
  __syncthreads();
  if (k >= NN) return;
  const int m = k / n1, t = k - m * n1;
  double *f = F + patch * f_patch_stride + lp * n1 + t;
  double v[5] = {0, 0, 0, 0, 0};
  if (e >= 0) {
    double p0 = 0, p1 = 0, w0 = 0, w1 = 0;
    for (int r = 0; r < n1; ++r) {
      p0 = fma(Ds[m * n1 + r], P[r * n1 + t], p0);
      w0 = fma(Ds[m * n1 + r], W[r * n1 + t], w0);
      p1 = fma(Ds[t * n1 + r], P[m * n1 + r], p1);
      w1 = fma(Ds[t * n1 + r], W[m * n1 + r], w1);
    }
    const double *ij = invJ + e * 4 * NN + k;
    const double a0 = ij[0], a1 = ij[NN], b0 = ij[2 * NN], b1 = ij[3 * NN];
    const double jw = JxW[e * NN + k], rho = x_phys[e * 2 * NN + k];
    const double q = jw * (a0 * b1 - a1 * b0);
    const double dl = rho > 0.0 ? jw / rho : 0.0;
    const double s0 = dl * a1, s1 = dl * b1;
    const double om = W[k];
    v[0] = n_rey * (q * w1 + s0 * om);
    v[1] = n_rey * (-q * w0 + s1 * om);
    v[2] = -n_rey * q * p1;
    v[3] = n_rey * q * p0;
    v[4] = n_rey * (s0 * p0 + s1 * p1);
  }
#pragma unroll
  for (int c = 0; c < 5; ++c) f[((kFacStokes + c) * n1 + m) * NP] = v[c];
}

// Element-local diagonals of the three blocks that have one: K + dL (Lve), K + c0 D_pp +
// c1 D_qq (E2e) and the advection block d row0 / d psi (ADV), in engine slot order.
__global__ void stokes_local_diag_kernel(int n1, int pe, int n_fac, int64_t n_slot_elems,
                                         const double *__restrict__ F, int64_t f_patch_stride,
                                         const double *__restrict__ D, double *__restrict__ locL,
                                         double *__restrict__ locE, double *__restrict__ locM,
                                         double *__restrict__ locA) {
  const int NN = n1 * n1, NP = n1 * pe;
  const int64_t total = n_slot_elems * NN;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t slot = i / NN;
    const int k = (int)(i - slot * NN);
    const int p = k / n1, q = k - p * n1;
    const int64_t patch = slot / pe;
    const int lp = (int)(slot - patch * pe);
    const double *g = F + patch * f_patch_stride + lp * n1;
    const double dpp = D[p * n1 + p], dqq = D[q * n1 + q];
    double acc = 2.0 * g[(n1 + p) * NP + q] * dpp * dqq;
    for (int m = 0; m < n1; ++m) {
      const double d0 = D[m * n1 + p], d1 = D[m * n1 + q];
      acc = fma(g[m * NP + q], d0 * d0, acc);
      acc = fma(g[(2 * n1 + p) * NP + m], d1 * d1, acc);
    }
    locL[i] = acc + g[(5 * n1 + p) * NP + q];
    locE[i] = acc + g[(3 * n1 + p) * NP + q] * dpp + g[(4 * n1 + p) * NP + q] * dqq;
    locM[i] = g[(6 * n1 + p) * NP + q];
    if (locA)
      locA[i] = n_fac >= kFacAdv
                    ? g[(7 * n1 + p) * NP + q] * dpp + g[(8 * n1 + p) * NP + q] * dqq
                    : 0.0;
  }
}

__global__ void scatter_fix_kernel(int64_t n, const int64_t *__restrict__ idx,
                                   const double *__restrict__ src, double *__restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = idx[i];
    y[d] = src ? src[d] : 0.0;
  }
}

// ---- GMRES building blocks ----------------------------------------------------------
// h[j] = V_j . w for j < k (one pass over w and the k basis vectors), then ||w||^2 in
// h[k]; fixed-order two-stage reduction -> bit-reproducible.
constexpr int kDotThreads = 256;
constexpr int kDotMaxBlocks = 148 * 4;
constexpr int kDotChunk = 8;  // basis vectors per pass over w

__global__ void __launch_bounds__(kDotThreads)
    multi_dot_kernel(int64_t n, int k, const double *__restrict__ V, int64_t ldv,
                     const double *__restrict__ w, double *__restrict__ partials) {
  __shared__ double red[32];
  for (int j0 = 0; j0 < k; j0 += kDotChunk) {
    double acc[kDotChunk];
#pragma unroll
    for (int j = 0; j < kDotChunk; ++j) acc[j] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
      const double wi = w[i];
#pragma unroll
      for (int j = 0; j < kDotChunk; ++j)
        if (j0 + j < k) acc[j] = fma(V[(int64_t)(j0 + j) * ldv + i], wi, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < kDotChunk; ++j) {
      if (j0 + j >= k) break;
      const double s = semk_block_sum(acc[j], red);
      if (threadIdx.x == 0) partials[(int64_t)(j0 + j) * gridDim.x + blockIdx.x] = s;
    }
  }
}
__global__ void __launch_bounds__(256)
    multi_dot_finish_kernel(int k, int nblocks, const double *__restrict__ partials,
                            double *__restrict__ out) {
  __shared__ double red[32];
  for (int j = blockIdx.x; j < k; j += gridDim.x) {
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) s += partials[(int64_t)j * nblocks + b];
    s = semk_block_sum(s, red);
    if (threadIdx.x == 0) out[j] = s;
  }
}
// w -= sum_j h[j] V_j (h on the device), optionally followed by nothing else
__global__ void __launch_bounds__(256)
    multi_axpy_kernel(int64_t n, int k, const double *__restrict__ V, int64_t ldv,
                      const double *__restrict__ h, double sign, double *__restrict__ w) {
  extern __shared__ double hs[];
  for (int j = threadIdx.x; j < k; j += blockDim.x) hs[j] = sign * h[j];
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double acc = w[i];
    for (int j = 0; j < k; ++j) acc = fma(hs[j], V[(int64_t)j * ldv + i], acc);
    w[i] = acc;
  }
}
// out = alpha * a (+ b)
__global__ void __launch_bounds__(256)
    scale_add_kernel(int64_t n, double alpha, const double *__restrict__ a,
                     const double *__restrict__ b, double *__restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = b ? fma(alpha, a[i], b[i]) : alpha * a[i];
}
// z_n = B_n r_n with the node's 2x2 block B_n = binv[n][0..3] (row major)
__global__ void __launch_bounds__(256)
    block2_apply_kernel(int64_t n_nodes, const double *__restrict__ binv,
                        const double2 *__restrict__ r, double2 *__restrict__ z) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double2 b0 = reinterpret_cast<const double2 *>(binv)[2 * i];
    const double2 b1 = reinterpret_cast<const double2 *>(binv)[2 * i + 1];
    const double2 v = r[i];
    z[i] = make_double2(b0.x * v.x + b0.y * v.y, b1.x * v.x + b1.y * v.y);
  }
}

// ---- glue of the Poisson block preconditioner (stokes.PoissonBlockPreconditioner) ---------
// t = (0, nimg * src_omega): the boundary vorticities om_G = -b_G / M_G as a (psi, omega) vector
__global__ void __launch_bounds__(256)
    prec_gamma_kernel(int64_t n_nodes, const double2 *__restrict__ src,
                      const double *__restrict__ nimg, double2 *__restrict__ t) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes;
       i += (int64_t)gridDim.x * blockDim.x)
    t[i] = make_double2(0.0, src[i].y * nimg[i]);
}
// which = 0: r = src_psi - y_psi           (rows wte minus L om_G)
// which = 1: r = src_omega + coef * om     (rows wdef plus M om_I)
// r is zeroed outside the free nodes; f = r * inv_mass_int is the nodal load whose element
// load JxW f reproduces r on the element interiors (zero on exterior nodes)
__global__ void __launch_bounds__(256)
    prec_rhs_kernel(int64_t n_nodes, int which, const double2 *__restrict__ src,
                    const double2 *__restrict__ y, const double *__restrict__ coef,
                    const double *__restrict__ om, const uint8_t *__restrict__ free_mask,
                    const double *__restrict__ inv_mass_int, double *__restrict__ r,
                    double *__restrict__ f) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes;
       i += (int64_t)gridDim.x * blockDim.x) {
    double v = which == 0 ? src[i].x - y[i].x : fma(coef[i], om[i], src[i].y);
    if (!free_mask[i]) v = 0.0;
    r[i] = v;
    f[i] = v * inv_mass_int[i];
  }
}
// dst = (free ? psi : 0, free ? om_I : om_G)
__global__ void __launch_bounds__(256)
    prec_out_kernel(int64_t n_nodes, const uint8_t *__restrict__ free_mask,
                    const double *__restrict__ psi, const double *__restrict__ om_i,
                    const double2 *__restrict__ t, double2 *__restrict__ dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_nodes;
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool fr = free_mask[i] != 0;
    dst[i] = make_double2(fr ? psi[i] : 0.0, fr ? om_i[i] : t[i].y);
  }
}
// b = Dirichlet ? 0 : g + r  (condensed right-hand side of a residual with zero boundary data)
__global__ void __launch_bounds__(256)
    cond_rhs_finish_kernel(int64_t n, const double *__restrict__ g, const double *__restrict__ r,
                           const uint8_t *__restrict__ dirichlet, double *__restrict__ b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    b[i] = (dirichlet && dirichlet[i]) ? 0.0 : g[i] + r[i];
}

inline int grid_for(int64_t n, int threads, int cap) {
  const int64_t want = (n + threads - 1) / threads;
  return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

template <int N, int PE, bool ADV>
int run_stokes(const semk_op &op, const DMatEO &dm, const double *u, double *y, double adv_scale,
               cudaStream_t st, int *grid_out) {
  const size_t smem = stokes_smem_layout(N, PE, op.g_patch_stride, op.pn_patch_stride,
                                         op.eloc_patch_stride, op.inv_patch_stride)
                          .total;
  if (smem > 227 * 1024) {
    semk_set_error("stokes kernel: shared memory request exceeds 227 KB");
    return SEMK_ERR_UNSUPPORTED;
  }
  static size_t configured = 0;
  static int per_sm = 0, sms = 0;
  auto kern = stokes_patch_kernel<N, PE, ADV>;
  if (smem > configured || per_sm == 0) {
    SEMK_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SEMK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern,
                                                                  stokes_threads(N, PE), smem));
    int dev = 0;
    SEMK_CUDA_CHECK(cudaGetDevice(&dev));
    SEMK_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    configured = smem;
  }
  if (per_sm < 1) {
    semk_set_error("stokes kernel: does not fit on an SM");
    return SEMK_ERR_UNSUPPORTED;
  }
  const int64_t resident = (int64_t)per_sm * sms;
  const int64_t want = (op.max_ctas > 0 && op.max_ctas < resident) ? op.max_ctas : resident;
  const unsigned grid = (unsigned)(op.n_patch < want ? op.n_patch : want);
  if (grid_out) *grid_out = (int)grid;
  if (grid == 0) return SEMK_OK;
  kern<<<grid, stokes_threads(N, PE), smem, st>>>(op, dm, reinterpret_cast<const double2 *>(u),
                                                 reinterpret_cast<double2 *>(y), adv_scale);
  SEMK_LAUNCH_CHECK("stokes_patch_kernel");
  return SEMK_OK;
}

int check_stokes(const semk_stokes_op *sop, const char *who) {
  if (!sop) {
    semk_set_error(std::string(who) + ": null operator");
    return SEMK_ERR_INVALID;
  }
  const semk_op &op = sop->plan;
  if (op.n1 < 2 || op.n1 > SEMK_MAX_N1) {
    semk_set_error(std::string(who) + ": n1 outside [2, 17]");
    return SEMK_ERR_UNSUPPORTED;
  }
  if (op.elems_per_patch != 4 && op.elems_per_patch != 8) {
    semk_set_error(std::string(who) + ": elems_per_patch must be 4 or 8");
    return SEMK_ERR_UNSUPPORTED;
  }
  if (sop->n_fac != kFacStokes && sop->n_fac != kFacAdv) {
    semk_set_error(std::string(who) + ": n_fac must be 7 (Stokes) or 12 (with advection)");
    return SEMK_ERR_INVALID;
  }
  const int64_t nnp = (int64_t)op.n1 * op.n1 * op.elems_per_patch;
  if (!op.G || !op.pnode || !op.eloc || !op.patch_hdr || !op.inv || (op.g_patch_stride & 1) ||
      op.g_patch_stride < sop->n_fac * nnp || (op.n_slots > 0 && !op.slot_buf) || !op.D_host) {
    semk_set_error(std::string(who) + ": operator tables incomplete");
    return SEMK_ERR_INVALID;
  }
  return SEMK_OK;
}

}  // namespace

extern "C" int64_t semk_stokes_smem_bytes(int n1, int elems_per_patch, int64_t f_patch_stride,
                                          int64_t pn_patch_stride, int64_t eloc_patch_stride,
                                          int64_t inv_patch_stride) {
  return (int64_t)stokes_smem_layout(n1, elems_per_patch, f_patch_stride, pn_patch_stride,
                                     eloc_patch_stride, inv_patch_stride)
      .total;
}

extern "C" int semk_stokes_factors_f64(int n1, int64_t n_slot_elems, const double *invJ,
                                       const double *JxW, const double *x_phys,
                                       const int64_t *elem_of_slot, double *F,
                                       int64_t f_patch_stride, int elems_per_patch, void *stream) {
  SEMK_REQUIRE(n1 >= 2 && n1 <= SEMK_MAX_N1 && invJ && JxW && x_phys && F,
               "semk_stokes_factors_f64: bad arguments");
  const int64_t total = n_slot_elems * n1 * n1;
  if (total == 0) return SEMK_OK;
  stokes_factors_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, semk_stream(stream)>>>(
      n1, elems_per_patch, n_slot_elems, invJ, JxW, x_phys, elem_of_slot, F, f_patch_stride);
  SEMK_LAUNCH_CHECK("stokes_factors_kernel");
  return SEMK_OK;
}

extern "C" int semk_stokes_linearize_f64(int n1, int64_t n_slot_elems, const double *D,
                                         const double *invJ, const double *JxW,
                                         const double *x_phys, const uint32_t *l2g,
                                         const int64_t *elem_of_slot, const double *state,
                                         double n_rey, double *F, int64_t f_patch_stride,
                                         int elems_per_patch, void *stream) {
  SEMK_REQUIRE(n1 >= 2 && n1 <= SEMK_MAX_N1 && D && invJ && JxW && x_phys && l2g && state && F,
               "semk_stokes_linearize_f64: bad arguments");
  if (n_slot_elems == 0) return SEMK_OK;
  const int NN = n1 * n1;
  const int threads = ((NN + 31) / 32) * 32;
  stokes_linearize_kernel<<<(unsigned)n_slot_elems, threads, 3 * NN * sizeof(double),
                            semk_stream(stream)>>>(
      n1, elems_per_patch, D, invJ, JxW, x_phys, l2g, elem_of_slot,
      reinterpret_cast<const double2 *>(state), n_rey, F, f_patch_stride);
  SEMK_LAUNCH_CHECK("stokes_linearize_kernel");
  return SEMK_OK;
}

extern "C" int semk_stokes_local_diag_f64(const semk_stokes_op *sop, const double *D_dev,
                                          int64_t n_slot_elems, double *locL, double *locE,
                                          double *locM, double *locA, void *stream) {
  int rc = check_stokes(sop, "semk_stokes_local_diag_f64");
  if (rc != SEMK_OK) return rc;
  const semk_op &op = sop->plan;
  const int64_t total = n_slot_elems * op.n1 * op.n1;
  if (total == 0) return SEMK_OK;
  stokes_local_diag_kernel<<<grid_for(total, 256, 148 * 16), 256, 0, semk_stream(stream)>>>(
      op.n1, op.elems_per_patch, sop->n_fac, n_slot_elems, op.G, op.g_patch_stride, D_dev, locL,
      locE, locM, locA);
  SEMK_LAUNCH_CHECK("stokes_local_diag_kernel");
  return SEMK_OK;
}

extern "C" int semk_stokes_apply_f64(const semk_stokes_op *sop, const double *u, double *y,
                                     int zero_essential_rows, double adv_scale, void *stream) {
  int rc = check_stokes(sop, "semk_stokes_apply_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(u && y && u != y, "semk_stokes_apply_f64: u, y must be distinct device buffers");
  SEMK_REQUIRE(((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0,
               "semk_stokes_apply_f64: u, y must be 16-byte aligned");
  const semk_op &op = sop->plan;
  cudaStream_t st = semk_stream(stream);
  DMatEO dm;
  if (!make_dmat_eo(op.n1, op.D_host, &dm)) {
    semk_set_error("semk_stokes_apply_f64: D is not centro-antisymmetric");
    return SEMK_ERR_INVALID;
  }
  const bool adv = sop->n_fac == kFacAdv;
#define SEMK_CALL(NV)                                                                          \
  do {                                                                                         \
    if (op.elems_per_patch == 8)                                                               \
      rc = adv ? run_stokes<NV, 8, true>(op, dm, u, y, adv_scale, st, nullptr)                 \
               : run_stokes<NV, 8, false>(op, dm, u, y, adv_scale, st, nullptr);               \
    else                                                                                       \
      rc = adv ? run_stokes<NV, 4, true>(op, dm, u, y, adv_scale, st, nullptr)                 \
               : run_stokes<NV, 4, false>(op, dm, u, y, adv_scale, st, nullptr);               \
    if (rc != SEMK_OK) return rc;                                                              \
  } while (0)
  SEMK_DISPATCH_N1(op.n1, SEMK_CALL)
#undef SEMK_CALL
  const int chunk_blocks = (int)((op.n_shared_chunk + 7) / 8);
  const int64_t want = (op.n_shared + 255) / 256;
  const int rec_blocks = (int)(want < 148 * 64 ? want : 148 * 64);
  if (chunk_blocks + rec_blocks > 0) {
    stokes_shared_nodes_kernel<<<chunk_blocks + rec_blocks, 256, 0, st>>>(
        op, reinterpret_cast<double2 *>(y), chunk_blocks);
    SEMK_LAUNCH_CHECK("stokes_shared_nodes_kernel");
  }
  if (zero_essential_rows && sop->n_ess > 0) {
    scatter_fix_kernel<<<grid_for(sop->n_ess, 256, 148 * 4), 256, 0, st>>>(sop->n_ess, sop->ess_dof,
                                                                          nullptr, y);
    SEMK_LAUNCH_CHECK("scatter_fix_kernel");
  }
  return SEMK_OK;
}

extern "C" int semk_scatter_fix_f64(int64_t n, const int64_t *idx, const double *src, double *y,
                                    void *stream) {
  if (n <= 0) return SEMK_OK;
  SEMK_REQUIRE(idx && y, "semk_scatter_fix_f64: null pointer");
  scatter_fix_kernel<<<grid_for(n, 256, 148 * 4), 256, 0, semk_stream(stream)>>>(n, idx, src, y);
  SEMK_LAUNCH_CHECK("scatter_fix_kernel");
  return SEMK_OK;
}

extern "C" int64_t semk_multi_dot_partials_len(int k) { return (int64_t)(k + 1) * kDotMaxBlocks; }

extern "C" int semk_multi_dot_f64(int64_t n, int k, const double *V, int64_t ldv, const double *w,
                                  double *out, double *partials, void *stream) {
  SEMK_REQUIRE(n > 0 && k >= 1 && V && w && out && partials, "semk_multi_dot_f64: bad arguments");
  cudaStream_t st = semk_stream(stream);
  const int blocks = grid_for(n, kDotThreads * 4, kDotMaxBlocks);
  multi_dot_kernel<<<blocks, kDotThreads, 0, st>>>(n, k, V, ldv, w, partials);
  SEMK_LAUNCH_CHECK("multi_dot_kernel");
  multi_dot_finish_kernel<<<k < 64 ? k : 64, 256, 0, st>>>(k, blocks, partials, out);
  SEMK_LAUNCH_CHECK("multi_dot_finish_kernel");
  return SEMK_OK;
}

extern "C" int semk_multi_axpy_f64(int64_t n, int k, const double *V, int64_t ldv, const double *h,
                                   double sign, double *w, void *stream) {
  SEMK_REQUIRE(n > 0 && k >= 1 && k <= 4096 && V && h && w, "semk_multi_axpy_f64: bad arguments");
  multi_axpy_kernel<<<grid_for(n, 256, 148 * 8), 256, k * sizeof(double), semk_stream(stream)>>>(
      n, k, V, ldv, h, sign, w);
  SEMK_LAUNCH_CHECK("multi_axpy_kernel");
  return SEMK_OK;
}

extern "C" int semk_vec_scale_add_f64(int64_t n, double alpha, const double *a, const double *b,
                                      double *out, void *stream) {
  SEMK_REQUIRE(n > 0 && a && out, "semk_vec_scale_add_f64: bad arguments");
  scale_add_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, semk_stream(stream)>>>(n, alpha, a, b, out);
  SEMK_LAUNCH_CHECK("scale_add_kernel");
  return SEMK_OK;
}

extern "C" int semk_block2_apply_f64(int64_t n_nodes, const double *binv, const double *r,
                                     double *z, void *stream) {
  SEMK_REQUIRE(n_nodes > 0 && binv && r && z, "semk_block2_apply_f64: bad arguments");
  block2_apply_kernel<<<grid_for(n_nodes, 256, 148 * 8), 256, 0, semk_stream(stream)>>>(
      n_nodes, binv, reinterpret_cast<const double2 *>(r), reinterpret_cast<double2 *>(z));
  SEMK_LAUNCH_CHECK("block2_apply_kernel");
  return SEMK_OK;
}

extern "C" int semk_stokes_prec_gamma_f64(int64_t n_nodes, const double *src, const double *nimg,
                                          double *t, void *stream) {
  SEMK_REQUIRE(n_nodes > 0 && src && nimg && t, "semk_stokes_prec_gamma_f64: bad arguments");
  prec_gamma_kernel<<<grid_for(n_nodes, 256, 148 * 8), 256, 0, semk_stream(stream)>>>(
      n_nodes, reinterpret_cast<const double2 *>(src), nimg, reinterpret_cast<double2 *>(t));
  SEMK_LAUNCH_CHECK("prec_gamma_kernel");
  return SEMK_OK;
}

extern "C" int semk_stokes_prec_rhs_f64(int64_t n_nodes, int which, const double *src,
                                        const double *y, const double *coef, const double *om,
                                        const uint8_t *free_mask, const double *inv_mass_int,
                                        double *r, double *f, void *stream) {
  SEMK_REQUIRE(n_nodes > 0 && src && free_mask && inv_mass_int && r && f &&
                   (which == 0 ? y != nullptr : (which == 1 && coef && om)),
               "semk_stokes_prec_rhs_f64: bad arguments");
  prec_rhs_kernel<<<grid_for(n_nodes, 256, 148 * 8), 256, 0, semk_stream(stream)>>>(
      n_nodes, which, reinterpret_cast<const double2 *>(src),
      reinterpret_cast<const double2 *>(y), coef, om, free_mask, inv_mass_int, r, f);
  SEMK_LAUNCH_CHECK("prec_rhs_kernel");
  return SEMK_OK;
}

extern "C" int semk_stokes_prec_out_f64(int64_t n_nodes, const uint8_t *free_mask,
                                        const double *psi, const double *om_i, const double *t,
                                        double *dst, void *stream) {
  SEMK_REQUIRE(n_nodes > 0 && free_mask && psi && om_i && t && dst,
               "semk_stokes_prec_out_f64: bad arguments");
  prec_out_kernel<<<grid_for(n_nodes, 256, 148 * 8), 256, 0, semk_stream(stream)>>>(
      n_nodes, free_mask, psi, om_i, reinterpret_cast<const double2 *>(t),
      reinterpret_cast<double2 *>(dst));
  SEMK_LAUNCH_CHECK("prec_out_kernel");
  return SEMK_OK;
}

extern "C" int semk_sc_rhs_finish_f64(int64_t n, const double *g, const double *r,
                                      const uint8_t *dirichlet, double *b, void *stream) {
  SEMK_REQUIRE(n > 0 && g && r && b, "semk_sc_rhs_finish_f64: bad arguments");
  cond_rhs_finish_kernel<<<grid_for(n, 256, 148 * 8), 256, 0, semk_stream(stream)>>>(n, g, r,
                                                                                    dirichlet, b);
  SEMK_LAUNCH_CHECK("cond_rhs_finish_kernel");
  return SEMK_OK;
}
