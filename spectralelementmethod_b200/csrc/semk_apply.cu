// semk_apply.cu -- K2/K3/K5: matrix-free Poisson apply, assembly, diagonal.
//
// Reference being replaced (per element, in a Python loop):
//   Lse[p,q,r,s] from four O(N^5) einsums            examples/poisson.py:181-193
//   u_loc = u[L2G];  y_loc = einsum('pqrs,rs', L, u_loc)
//                                    examples/squirmer-axisymmetric.py:268-295
//   scatter-add through the L2G map                  sem/discrete.py:491-499
//   Dirichlet row/column elimination                 sem/discrete.py:505-510
//
// Here the dense Lse is never formed.  With the symmetric geometric factors
// G (semk_geom.cu) the local operator is
//   ur = D u,  us = u D^T,  w0 = G00.ur + G01.us,  w1 = G01.ur + G11.us,
//   y  = D^T w0 + w1 D.
//
// Kernel layout (patch kernel, the production path):
//   * one CTA per patch of PE element slots, N threads per element; thread t
//     of an element owns column t (nodes [m][t], m = 0..N-1) in registers.
//   * contractions along the in-thread axis are N*N DFMAs whose D operand is
//     a compile-time-indexed __constant__ entry (folded into the DFMA as a
//     constant-bank operand: no shared-memory traffic, no registers);
//     the other axis is reached by transposing through shared memory.
//   * the patch's geometric factors (59% of all HBM traffic) arrive by one
//     TMA bulk copy (cp.async.bulk -> SASS UBLKCP) signalled on an mbarrier,
//     issued before the nodal gather so the two overlap.
//   * nodal values are gathered once per patch into shared memory through the
//     patch node table (coalesced for locality-preserving numberings), with
//     the Dirichlet mask carried in the table's flag bits.
//   * assembly inside the patch is a fixed-order gather (no atomics); private nodes are
//     stored straight to y, shared nodes go to interface slots that the tiny
//     `shared_nodes_kernel` sums in a fixed order => bit-reproducible.
#include "semk_common.cuh"

#include <cmath>
#include <cstring>

#include "semk_elem.cuh"

namespace {

// The element-local operator for one column-owning thread.
//   tidp = le*N + t identifies (element-in-CTA, column);  ucol[m] = u[m][t] on
//   entry, ycol[m] = y[m][t] on exit.
//   A, B: CTA-wide scratch, N rows of RS doubles.
//   g: this thread's column base inside the patch's G block; the factor
//      (c, m) of this column sits at g[(c*N + m) * g_row].
//   g_ready: mbarrier guarding a TMA-staged G (nullptr when g is global).
// Contains four __syncthreads(); every thread of the CTA must call it.
struct NoHook {
  __device__ __forceinline__ void operator()() const {}
};

template <int N, int RS, class DM, class Hook = NoHook, class Hook0 = NoHook>
__device__ __forceinline__ void local_poisson(const DM &dm, int le, int t, bool active,
                                              const double (&ucol)[N], double (&ycol)[N],
                                              double *__restrict__ A, double *__restrict__ B,
                                              const double *__restrict__ g, int g_row,
                                              uint64_t *g_ready, uint32_t g_parity = 0,
                                              Hook g_consumed = Hook(),
                                              Hook0 after_first_barrier = Hook0()) {
  const int tidp = le * N + t;
  double ur[N], tmp[N], us[N];
  if (active) {
#pragma unroll
    for (int m = 0; m < N; ++m) A[m * RS + tidp] = ucol[m];
  }
  __syncthreads();
  after_first_barrier();  // every thread has left the previous patch's write-out
  if (active) {
    mat_D<N>(dm, ucol, ur);  // ur[m][t] = sum_r D[m][r] u[r][t]
#pragma unroll
    for (int s = 0; s < N; ++s) tmp[s] = A[t * RS + le * N + s];  // row t of u
    mat_D<N>(dm, tmp, us);                                        // us[t][n] = sum_s D[n][s] u[t][s]
#pragma unroll
    for (int n = 0; n < N; ++n) B[t * RS + le * N + n] = us[n];
  }
  __syncthreads();
  // G staged by TMA: every thread observes the mbarrier phase itself (acquire)
  if (g_ready) semk_mbar_wait(g_ready, g_parity);
  double w1[N];
  if (active) {
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const double usc = B[m * RS + tidp];  // us[m][t]
      const double g00 = g[m * g_row], g01 = g[(N + m) * g_row], g11 = g[(2 * N + m) * g_row];
      tmp[m] = g00 * ur[m] + g01 * usc;  // w0[m][t]
      w1[m] = g01 * ur[m] + g11 * usc;   // w1[m][t]
    }
    mat_Dt<N>(dm, tmp, ycol);  // y0[p][t] = sum_m D[m][p] w0[m][t]
#pragma unroll
    for (int m = 0; m < N; ++m) A[m * RS + tidp] = w1[m];
  }
  __syncthreads();
  g_consumed();  // every thread has read its G entries: the G buffer may be refilled
  if (active) {
#pragma unroll
    for (int n = 0; n < N; ++n) tmp[n] = A[t * RS + le * N + n];  // row t of w1
    mat_Dt<N>(dm, tmp, us);                                       // y1[t][q] = sum_n w1[t][n] D[n][q]
#pragma unroll
    for (int q = 0; q < N; ++q) B[t * RS + le * N + q] = us[q];
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int m = 0; m < N; ++m) ycol[m] += B[m * RS + tidp];
  }
}

template <int N, int PE>
struct PatchCfg {
  static constexpr int kThreads = ((N * PE + 31) / 32) * 32;
  static constexpr int kRS = scratch_row_stride(N, PE);
};

// Resident CTAs per SM the kernel is compiled for (register budget): what the
// shared-memory footprint of the standard tiles (2x8, 1x8, 1x4 elements) allows.
__host__ __device__ constexpr int patch_min_blocks(int N, int PE) {
  const int bx = PE == 32 ? 4 : (PE == 16 ? 2 : 1), by = PE == 4 ? 4 : 8, p = N - 1;
  const long long mpn = (long long)(bx * p + 1) * (by * p + 1);
  const long long mpn4 = (mpn + 3) & ~3LL;
  const long long nn = (long long)N * N;
  const long long g = 8 * ((3 * nn * PE + 1) & ~1LL);
  const long long tab = 2 * ((4 * mpn4 + 2 * ((nn * PE + 7) & ~7LL) + 15) & ~15LL) + 64;
  const long long scr = 8LL * N * scratch_row_stride(N, PE);
  const long long ua = scr;
  const long long total = 32 + g + tab + 8 * mpn4 + ua + (scr > 256 ? scr : 256) + 1024;
  const long long by_smem = 233472 / total;
  const int threads = ((N * PE + 31) / 32) * 32;
  const long long by_regs = 65536 / ((long long)threads * 96);  // assume <= 96 registers/thread
  long long r = by_smem < by_regs ? by_smem : by_regs;
  return r < 1 ? 1 : (r > 8 ? 8 : (int)r);
}

constexpr int kGatherBatch = 8;

// Persistent, software-pipelined patch kernel.
//
// Each CTA loops over patches blockIdx.x, blockIdx.x + gridDim.x, ...  While it
// works on patch i, the TMA engine is already filling the other pipeline stage
// with patch i+1's geometric factors, node block and index block (issued one
// whole patch-time ahead: DRAM latency is off the critical path), and the
// nodal values of patch i+1 are gathered into registers right after the
// element operator of patch i, so that the loads are in flight during the
// assembly and the write-out.  CTAs never exit between patches,
// so no SM slot idles on CTA launch / retire (~2.4 k cycles each on this chip).
//
// MODE_APPLY:    y = A u (masked per flags), optional dot partials.
// MODE_ASSEMBLE: y = assembly of the element-local field `loc` (slot order).
template <int N, int PE, int MODE>
__global__ void __launch_bounds__(PatchCfg<N, PE>::kThreads, patch_min_blocks(N, PE))
    patch_kernel(semk_op op, DMatEO dm, const double *__restrict__ u,
                 const double *__restrict__ loc, double *__restrict__ y, int flags,
                 double fill_dirichlet, double *__restrict__ dot_partials, int64_t patch_begin,
                 int64_t patch_end) {
  constexpr int NN = N * N;
  constexpr int NP = N * PE;
  constexpr int kThreads = PatchCfg<N, PE>::kThreads;
  constexpr int RS = PatchCfg<N, PE>::kRS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PatchSmem L = patch_smem_layout(N, PE, MODE, op.g_patch_stride, op.pn_patch_stride,
                                        op.eloc_patch_stride, op.inv_patch_stride);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);  // [0..1]: tables, [2]: G, [3]: inv
  double *Gs = reinterpret_cast<double *>(smem_raw + L.gs);
  const uint16_t *inv_s = reinterpret_cast<const uint16_t *>(smem_raw + L.inv);
  double *As = reinterpret_cast<double *>(smem_raw + L.ua);
  double *Bs = reinterpret_cast<double *>(smem_raw + L.bs);
  double *red = reinterpret_cast<double *>(smem_raw + L.red);

  const int tid = threadIdx.x;
  const int le = tid / N, t = tid - le * N;
  const uint32_t pn_bytes = 4u * (uint32_t)op.pn_patch_stride;
  const uint32_t el_bytes = 2u * (uint32_t)op.eloc_patch_stride;
  const uint32_t g_bytes = (uint32_t)(op.g_patch_stride * sizeof(double));
  const uint32_t inv_bytes = 2u * (uint32_t)op.inv_patch_stride;
  const int inv_w4 = (int)(op.inv_width >> 2);  // 8-byte groups of inverse entries per node

  auto stage_ptr = [&](int s) { return smem_raw + L.stage0 + (size_t)s * L.stage_bytes; };
  // Patch headers travel two patches ahead through a 2-entry ring: the header of patch
  // i+1 names the (deduplicated) table blocks the copy for patch i+1 must fetch, so it
  // has to be in shared memory when that copy is issued, during patch i.
  uint32_t *hdr_ring = reinterpret_cast<uint32_t *>(smem_raw + L.hdr);
  // one thread: tables of `patch` (blocks pi, ei) into stage s, plus -- if it exists --
  // the header of the patch after it (`patch_after`) into ring entry `slot_after`
  auto issue_tables = [&](uint32_t pi, uint32_t ei, int s, int64_t patch_after, int slot_after) {
    unsigned char *base = stage_ptr(s);
    const bool more = patch_after >= 0;
    semk_mbar_expect_tx(&mbar[s], pn_bytes + el_bytes + (more ? 32u : 0u));
    semk_bulk_g2s(base + L.pn_off, op.pnode + (int64_t)pi * op.pn_patch_stride, pn_bytes, &mbar[s]);
    semk_bulk_g2s(base + L.el_off, op.eloc + (int64_t)ei * op.eloc_patch_stride, el_bytes, &mbar[s]);
    if (more) semk_bulk_g2s(hdr_ring + 8 * slot_after, op.patch_hdr + 8 * patch_after, 32u, &mbar[s]);
  };
  // The inverse table is single-buffered: patch i's table is fetched right after the
  // first barrier of patch i (when every thread has left patch i-1's write-out, its last
  // reader) and is first needed five barriers later, in patch i's own write-out.
  auto issue_inv = [&](uint32_t block) {  // one thread
    semk_mbar_expect_tx(&mbar[3], inv_bytes);
    semk_bulk_g2s(smem_raw + L.inv, op.inv + (int64_t)block * op.inv_patch_stride, inv_bytes,
                  &mbar[3]);
  };
  auto issue_g = [&](int64_t patch) {  // one thread
    semk_mbar_expect_tx(&mbar[2], g_bytes);
    semk_bulk_g2s(Gs, op.G + patch * op.g_patch_stride, g_bytes, &mbar[2]);
  };

  // this CTA's patch sequence: round-robin over the grid, so that at any time the
  // resident CTAs work on a compact window of the mesh (DRAM page / L2 locality)
  const int64_t step = (int64_t)gridDim.x;
  const int64_t p_first = patch_begin + (int64_t)blockIdx.x;
  const int64_t p_end = patch_end;  // (a sub-range of the patches: staged host apply)
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) semk_mbar_init(&mbar[i], 1);
    semk_fence_mbar_init();
    if (p_first < p_end) {
      // the first header is read directly (block indices needed right now) and also
      // copied into ring entry 0; the second patch's header goes to entry 1
      const uint32_t pi = op.patch_hdr[8 * p_first + 5], ei = op.patch_hdr[8 * p_first + 6];
      const int64_t p2 = p_first + step;
      unsigned char *base = stage_ptr(0);
      semk_mbar_expect_tx(&mbar[0], pn_bytes + el_bytes + 32u + (p2 < p_end ? 32u : 0u));
      semk_bulk_g2s(base + L.pn_off, op.pnode + (int64_t)pi * op.pn_patch_stride, pn_bytes, &mbar[0]);
      semk_bulk_g2s(base + L.el_off, op.eloc + (int64_t)ei * op.eloc_patch_stride, el_bytes, &mbar[0]);
      semk_bulk_g2s(hdr_ring, op.patch_hdr + 8 * p_first, 32u, &mbar[0]);
      if (p2 < p_end) semk_bulk_g2s(hdr_ring + 8, op.patch_hdr + 8 * p2, 32u, &mbar[0]);
      if (MODE == MODE_APPLY) issue_g(p_first);
    }
  }
  __syncthreads();  // mbarrier initialisation visible to every waiter

  // This thread's column of nodal values for the patch about to be processed,
  // gathered straight from global memory through the staged tables (entries of
  // nodes shared by several elements of the patch hit L1) and carried in
  // registers across the loop: the gather for patch i+1 is issued right after
  // the element operator of patch i, so its latency hides behind the assembly
  // and write-out of patch i.
  double ucol[N];
  uint32_t ucol_dir = 0;  // bit m set: entry m of the staged column is a Dirichlet node
  auto gather_column = [&](int s_tab, int64_t patch_of) {
    const unsigned char *sbn = stage_ptr(s_tab);
    const uint32_t *pnb = reinterpret_cast<const uint32_t *>(sbn + L.pn_off);
    const uint32_t id0 = hdr_ring[8 * s_tab + 4];  // ring entry = iteration parity = stage
    const uint16_t *elb = reinterpret_cast<const uint16_t *>(sbn + L.el_off);
    const bool act = (le < PE) && (patch_of * PE + le < op.n_elem);
    if (act) {
      uint32_t pn[N];
#pragma unroll
      for (int m = 0; m < N; ++m) pn[m] = pnb[elb[m * NP + tid]];
      // loads only: nothing below may depend on the values until the next patch
      // starts, or the warp would stall here instead of overlapping the latency
      ucol_dir = 0;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        ucol[m] = u[id0 + (pn[m] & SEMK_NODE_ID_MASK)];
        ucol_dir |= (pn[m] >> 31) << m;
      }
    }
  };
  if (MODE == MODE_APPLY && p_first < p_end) {
    semk_mbar_wait(&mbar[0], 0);
    gather_column(0, p_first);
  }
  __syncthreads();

  double dot = 0.0;
  const bool want_dot = (MODE == MODE_APPLY) && (dot_partials != nullptr);
  int it = 0;
  for (int64_t patch = p_first; patch < p_end; patch += step, ++it) {
    const int s = it & 1;
    const uint32_t par = (uint32_t)((it >> 1) & 1);
    const int64_t next = patch + step;
    const bool has_next = next < p_end;
    unsigned char *sb = stage_ptr(s);
    const uint32_t *pn_s = reinterpret_cast<const uint32_t *>(sb + L.pn_off);
    semk_mbar_wait(&mbar[s], par);
    const uint32_t *hdr = hdr_ring + 8 * s;  // read before this patch's first barrier
    const int npn = (int)hdr[0];
    const int npriv = (int)hdr[1];
    const int slot_base = (int)hdr[2];
    const uint32_t id0 = hdr[4];
    const uint32_t inv_block = hdr[7];
    const int64_t slot0 = patch * PE;
    const bool active = (le < PE) && (slot0 + le < op.n_elem);


    double ycol[N];
    if (MODE == MODE_APPLY) {
      if ((flags & SEMK_MASK_IN) && ucol_dir) {
#pragma unroll
        for (int m = 0; m < N; ++m)
          if ((ucol_dir >> m) & 1u) ucol[m] = 0.0;
      }
      // The single G buffer is refilled for the next patch as soon as every thread
      // has consumed this patch's factors (hook runs right after that barrier):
      // the copy then has the rest of this patch and the start of the next to land.
      auto refill_g = [&]() {
        if (tid == 0 && has_next) issue_g(next);
      };
      // Table stage s^1 (patch i-1's tables) is free once every thread has passed the
      // first barrier of this patch's operator: no end-of-patch barrier is needed.
      auto refill_tables = [&]() {
        if (tid == 0) issue_inv(inv_block);
        if (tid == 0 && has_next) {
          const uint32_t *hn = hdr_ring + 8 * (s ^ 1);  // header of the next patch
          const int64_t after = next + step;
          issue_tables(hn[5], hn[6], s ^ 1, after < p_end ? after : -1, s);
        }
      };
      local_poisson<N, RS>(dm, le, t, active, ucol, ycol, As, Bs, Gs + tid, NP, &mbar[2],
                           (uint32_t)(it & 1), refill_g, refill_tables);
      // u . y taken element by element: sum_e (Q_e u) . y_e = u . (sum_e Q_e^T y_e), so the
      // write-out never looks at u again.  Dirichlet rows are left out here (under MASK_IN
      // their entries of u are zero anyway) and come back as identity rows below.
      if (want_dot && active) {
#pragma unroll
        for (int m = 0; m < N; ++m) {
          const bool drop = (flags & SEMK_MASK_OUT) && ((ucol_dir >> m) & 1u);
          dot = fma(drop ? 0.0 : ucol[m], ycol[m], dot);
        }
      }
      // next patch: its tables landed long ago; start its gather now
      if (has_next) {
        semk_mbar_wait(&mbar[s ^ 1], (uint32_t)(((it + 1) >> 1) & 1));
        gather_column(s ^ 1, next);
      }
    } else {
      __syncthreads();  // previous patch's write-out done: its table stage may be refilled
      if (tid == 0) issue_inv(inv_block);
      if (tid == 0 && has_next) {
        const uint32_t *hn = hdr_ring + 8 * (s ^ 1);
        const int64_t after = next + step;
        issue_tables(hn[5], hn[6], s ^ 1, after < p_end ? after : -1, s);
      }
      if (active) {
        const double *lr = loc + (slot0 + le) * NN;
#pragma unroll
        for (int m = 0; m < N; ++m) ycol[m] = lr[m * N + t];
      }
    }

    // ---- assemble by GATHERING (no atomics, no colour phases): every thread leaves its
    // column of element results in scratch B (its own entries, the ones it just read),
    // then each patch node sums its <= inv_width contributions in a fixed order (ascending
    // element slot) straight from B and is stored: private nodes -> y, shared nodes ->
    // interface slots ------------------------------------------------------------------
    if (active) {
#pragma unroll
      for (int m = 0; m < N; ++m) Bs[m * RS + le * N + t] = ycol[m];
    }
    semk_mbar_wait(&mbar[3], (uint32_t)(it & 1));  // this patch's inverse table has landed
    __syncthreads();
    for (int k0 = tid; k0 < npn; k0 += kGatherBatch * kThreads) {
      uint32_t pnv[kGatherBatch];
      uint2 ev[kGatherBatch];
#pragma unroll
      for (int j = 0; j < kGatherBatch; ++j) {
        const int k = k0 + j * kThreads;
        const bool in = k < npn;
        pnv[j] = in ? pn_s[k] : 0xffffffffu;
        ev[j] = in ? reinterpret_cast<const uint2 *>(inv_s)[(size_t)k * inv_w4]
                   : make_uint2(0xffffffffu, 0xffffffffu);
      }
      double vv[kGatherBatch];
#pragma unroll
      for (int j = 0; j < kGatherBatch; ++j) {
        // (every listed node has at least one contribution)
        const uint32_t e0 = ev[j].x & 0xffffu, e1 = ev[j].x >> 16;
        const uint32_t e2 = ev[j].y & 0xffffu, e3 = ev[j].y >> 16;
        double v = (e0 != 0xffffu) ? Bs[e0] : 0.0;
        if (e1 != 0xffffu) v += Bs[e1];
        if (e2 != 0xffffu) v += Bs[e2];
        if (e3 != 0xffffu) v += Bs[e3];
        vv[j] = v;
      }
      if (inv_w4 > 1) {  // irregular meshes: more than four elements of the patch at a node
#pragma unroll
        for (int j = 0; j < kGatherBatch; ++j) {
          const int k = k0 + j * kThreads;
          if (k >= npn) continue;
          for (int w = 1; w < inv_w4; ++w) {
            const uint2 e = reinterpret_cast<const uint2 *>(inv_s)[(size_t)k * inv_w4 + w];
            const uint32_t q[4] = {e.x & 0xffffu, e.x >> 16, e.y & 0xffffu, e.y >> 16};
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (q[i] != 0xffffu) vv[j] += Bs[q[i]];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < kGatherBatch; ++j) {
        const int k = k0 + j * kThreads;
        if (k >= npn) continue;
        const uint32_t pn = pnv[j];
        double v = vv[j];
        if (k < npriv) {
          const uint32_t g = id0 + (pn & SEMK_NODE_ID_MASK);
          if ((pn & SEMK_NODE_DIRICHLET) && (flags & SEMK_MASK_OUT)) {
            if (MODE == MODE_APPLY) {
              v = 0.0;
              if (flags & SEMK_DIRICHLET_IDENTITY) {
                v = u[g];
                dot = fma(v, v, dot);  // the identity rows' share of u . y
              }
            } else {
              v = fill_dirichlet;
            }
          }
          y[g] = v;
        } else {
          op.slot_buf[slot_base + (k - npriv)] = v;
        }
      }
    }
    // no barrier here: the next patch's first operator barrier orders this write-out
    // (reads of stage s, of the inverse table and of B) before anything that overwrites them
  }
  if (want_dot) __syncthreads();  // the reduction scratch aliases B: let the write-out finish
  if (want_dot) {
    const double sres = semk_block_sum(dot, red);
    if (tid == 0) dot_partials[blockIdx.x] = sres;
  }
}

// Interface reduction: sum the partial sums of every shared node in a fixed
// order (ascending patch) and write the node's final value.
//   CTAs [0, chunk_blocks): one warp per affine chunk of <= 32 two-patch nodes --
//     node ids and both slot runs are arithmetic progressions, so a whole patch
//     edge is reduced with coalesced accesses and 32 bytes of table per chunk;
//   remaining CTAs: per-node records (nodes touched by 3+ patches), slot list inline.
// sub-range of the interface tables (both sorted by the highest patch involved, so the
// entries that are complete after patches [0, P) form a prefix): staged host apply
struct InterfaceRange {
  int64_t chunk_begin, chunk_end, rec_begin, rec_end;
};

constexpr int kChunkWarps = 8;   // warps per CTA of 256 threads
#ifndef SEMK_CHUNK_UNROLL
#define SEMK_CHUNK_UNROLL 4
#endif
constexpr int kChunkUnroll = SEMK_CHUNK_UNROLL;  // chunks per warp

// Writes the final value of a shared node; returns its identity-row share of u . y
// (the rest of u . y is taken element by element in the patch kernel).
template <int MODE>
__device__ __forceinline__ double finish_shared_node(const semk_op &op, uint32_t g, bool dir,
                                                     double v, const double *__restrict__ u,
                                                     double *__restrict__ y, int flags,
                                                     double fill_dirichlet, bool want_dot) {
  double d = 0.0;
  if (dir && (flags & SEMK_MASK_OUT)) {
    if (MODE == MODE_APPLY) {
      v = 0.0;
      if (flags & SEMK_DIRICHLET_IDENTITY) {
        v = u[g];
        d = v * v;
      }
    } else {
      v = fill_dirichlet;
    }
  }
  y[g] = v;
  return want_dot ? d : 0.0;
}

template <int MODE>
__global__ void __launch_bounds__(256)
    shared_nodes_kernel(semk_op op, const double *__restrict__ u, double *__restrict__ y,
                        int flags, double fill_dirichlet, double *__restrict__ dot_partials,
                        int64_t partial_offset, int chunk_blocks, InterfaceRange rng) {
  __shared__ double red[32];
  double dot = 0.0;
  const bool want_dot = (MODE == MODE_APPLY) && (dot_partials != nullptr);
  if ((int)blockIdx.x < chunk_blocks) {
    const int lane = threadIdx.x & 31;
    // each warp reduces kChunkUnroll chunks: all table loads first, then all slot
    // loads, then the stores (three dependent memory levels, kept wide)
    const int64_t cbase =
        rng.chunk_begin + ((int64_t)blockIdx.x * kChunkWarps + (threadIdx.x >> 5)) * kChunkUnroll;
    uint4 c0[kChunkUnroll], c1[kChunkUnroll];
    double va[kChunkUnroll], vb[kChunkUnroll];
#pragma unroll
    for (int k = 0; k < kChunkUnroll; ++k) {
      const int64_t c = cbase + k;
      const bool on = c < rng.chunk_end;
      // c0 = {node0, dn, a0, da}, c1 = {b0, db, len, Dirichlet mask}
      c0[k] = on ? reinterpret_cast<const uint4 *>(op.shared_chunk)[2 * c] : make_uint4(0, 0, 0, 0);
      c1[k] = on ? reinterpret_cast<const uint4 *>(op.shared_chunk)[2 * c + 1]
                 : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < kChunkUnroll; ++k) {
      const bool on = lane < (int)c1[k].z;
      va[k] = on ? op.slot_buf[c0[k].z + (uint32_t)lane * c0[k].w] : 0.0;
      vb[k] = on ? op.slot_buf[c1[k].x + (uint32_t)lane * c1[k].y] : 0.0;
    }
#pragma unroll
    for (int k = 0; k < kChunkUnroll; ++k) {
      if (lane < (int)c1[k].z) {
        const uint32_t g = c0[k].x + (uint32_t)lane * c0[k].y;
        dot += finish_shared_node<MODE>(op, g, ((c1[k].w >> lane) & 1u) != 0, va[k] + vb[k], u, y,
                                        flags, fill_dirichlet, want_dot);
      }
    }
  } else {
    // per-node records (nodes touched by 3+ patches): one node per thread, the whole
    // slot list inline in a 32-byte record, so two dependent memory levels
    const int64_t nb = (int64_t)gridDim.x - chunk_blocks;
    const int64_t stride = nb * blockDim.x;
    for (int64_t i = rng.rec_begin + ((int64_t)blockIdx.x - chunk_blocks) * blockDim.x + threadIdx.x;
         i < rng.rec_end; i += stride) {
      const uint4 r0 = reinterpret_cast<const uint4 *>(op.shared_rec)[2 * i];
      const uint4 r1 = reinterpret_cast<const uint4 *>(op.shared_rec)[2 * i + 1];
      const uint32_t cnt = r0.y;
      const uint32_t sl[6] = {r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
      double val[6];
#pragma unroll
      for (int j = 0; j < 6; ++j)
        val[j] = ((uint32_t)j < cnt && !(cnt > 6 && j == 5)) ? op.slot_buf[sl[j]] : 0.0;
      double v = val[0];  // ascending patch order: deterministic
#pragma unroll
      for (int j = 1; j < 6; ++j)
        if ((uint32_t)j < cnt && !(cnt > 6 && j == 5)) v += val[j];
      if (cnt > 6) {
        const uint32_t *ext = op.shared_ext + r1.w;
        for (uint32_t j = 5; j < cnt; ++j) v += op.slot_buf[ext[j - 5]];
      }
      dot += finish_shared_node<MODE>(op, r0.x & SEMK_NODE_ID_MASK,
                                      (r0.x & SEMK_NODE_DIRICHLET) != 0, v, u, y, flags,
                                      fill_dirichlet, want_dot);
    }
  }
  if (want_dot) {
    const double s = semk_block_sum(dot, red);
    if (threadIdx.x == 0) dot_partials[partial_offset + blockIdx.x] = s;
  }
}

// Final, fixed-order sum of the per-CTA partials (single CTA).
__global__ void __launch_bounds__(1024)
    reduce_partials_kernel(const double *__restrict__ partials, int64_t n,
                           double *__restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
  s = semk_block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s;
}

constexpr int kSharedBlocks = 148 * 64;  // upper bound on the per-node part of the interface kernel

inline void interface_blocks(const InterfaceRange &r, int *chunk_blocks, int *rec_blocks) {
  const int64_t per_block = (int64_t)kChunkWarps * kChunkUnroll;
  *chunk_blocks = (int)((r.chunk_end - r.chunk_begin + per_block - 1) / per_block);
  const int64_t want = (r.rec_end - r.rec_begin + 255) / 256;
  *rec_blocks = (int)(want < kSharedBlocks ? want : kSharedBlocks);
}

template <int MODE>
int launch_interface(const semk_op &op, const double *u, double *y, int flags, double fill,
                     double *partials, int64_t partial_offset, cudaStream_t st, int *blocks_out,
                     const InterfaceRange *range = nullptr) {
  const InterfaceRange all{0, op.n_shared_chunk, 0, op.n_shared};
  const InterfaceRange r = range ? *range : all;
  int chunk_blocks = 0, rec_blocks = 0;
  interface_blocks(r, &chunk_blocks, &rec_blocks);
  if (blocks_out) *blocks_out = chunk_blocks + rec_blocks;
  if (chunk_blocks + rec_blocks == 0) return SEMK_OK;
  shared_nodes_kernel<MODE><<<chunk_blocks + rec_blocks, 256, 0, st>>>(
      op, u, y, flags, fill, partials, partial_offset, chunk_blocks, r);
  SEMK_LAUNCH_CHECK("shared_nodes_kernel");
  return SEMK_OK;
}

// ---- simple atomic-scatter kernel (independent cross-check) -------------------
// Groups PEA consecutive element slots per CTA; reads the L2G map and the
// patch-interleaved G straight from global memory.
template <int N, int PEA>
__global__ void __launch_bounds__(PatchCfg<N, PEA>::kThreads)
    atomic_kernel(DMat dm, int64_t n_elem, const uint32_t *__restrict__ l2g,
                  const int64_t *__restrict__ elem_of_slot, const double *__restrict__ G,
                  int64_t g_patch_stride, int pe_plan, const uint8_t *__restrict__ dirichlet,
                  const double *__restrict__ u, double *__restrict__ y, int flags) {
  constexpr int NN = N * N;
  constexpr int RS = PatchCfg<N, PEA>::kRS;
  __shared__ double As[N * RS], Bs[N * RS];
  const int tid = threadIdx.x;
  const int le = tid / N, t = tid - le * N;
  const int64_t slot = (int64_t)blockIdx.x * PEA + le;
  bool active = (le < PEA) && (slot < n_elem);
  int64_t e = -1;
  if (active) e = elem_of_slot ? elem_of_slot[slot] : slot;
  active = active && e >= 0;  // (a negative entry is an empty slot of the engine order)
  double ucol[N], ycol[N];
  uint32_t gid[N];
  const double *g = G;
  if (active) {
    const uint32_t *row = l2g + e * NN;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      gid[m] = row[m * N + t];
      double v = u[gid[m]];
      if (dirichlet && (flags & SEMK_MASK_IN) && dirichlet[gid[m]]) v = 0.0;
      ucol[m] = v;
    }
    const int64_t patch = slot / pe_plan;
    const int lp = (int)(slot - patch * pe_plan);
    g = G + patch * g_patch_stride + lp * N + t;
  }
  local_poisson<N, RS>(dm, le < PEA ? le : 0, t, active, ucol, ycol, As, Bs, g, N * pe_plan,
                       nullptr);
  if (active) {
#pragma unroll
    for (int m = 0; m < N; ++m) atomicAdd(y + gid[m], ycol[m]);
  }
}

__global__ void dirichlet_fix_kernel(int64_t n, const uint8_t *__restrict__ dirichlet,
                                     const double *__restrict__ u, double *__restrict__ y,
                                     int flags) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    if (dirichlet[i]) y[i] = (flags & SEMK_DIRICHLET_IDENTITY) ? u[i] : 0.0;
}

// ---- K5: element-local diagonal ------------------------------------------------
// diag[p][q] = sum_m G00[m][q] D[m][p]^2 + 2 G01[p][q] D[p][p] D[q][q]
//            + sum_n G11[p][n] D[n][q]^2
__global__ void local_diag_kernel(int n1, int pe, int64_t n_slot_elems,
                                  const double *__restrict__ G, int64_t g_patch_stride,
                                  const double *__restrict__ D, double *__restrict__ loc) {
  const int NN = n1 * n1;
  const int NP = n1 * pe;
  const int64_t total = n_slot_elems * NN;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t slot = i / NN;
    const int k = (int)(i - slot * NN);
    const int p = k / n1, q = k - p * n1;
    const int64_t patch = slot / pe;
    const int lp = (int)(slot - patch * pe);
    const double *g = G + patch * g_patch_stride + lp * n1;  // + (c*n1 + m)*NP + n
    double acc = 2.0 * g[(n1 + p) * NP + q] * D[p * n1 + p] * D[q * n1 + q];
    for (int m = 0; m < n1; ++m) {
      const double d0 = D[m * n1 + p], d1 = D[m * n1 + q];
      acc = fma(g[m * NP + q], d0 * d0, acc);
      acc = fma(g[(2 * n1 + p) * NP + m], d1 * d1, acc);
    }
    loc[i] = acc;
  }
}

__global__ void weighted_local_kernel(int NN, int64_t n_elem, const double *__restrict__ JxW,
                                      const uint32_t *__restrict__ l2g,
                                      const int64_t *__restrict__ elem_of_slot,
                                      const double *__restrict__ f, double *__restrict__ loc) {
  const int64_t total = n_elem * NN;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t slot = i / NN;
    const int k = (int)(i - slot * NN);
    const int64_t e = elem_of_slot ? elem_of_slot[slot] : slot;
    double v = 0.0;  // empty slot
    if (e >= 0) {
      v = JxW[e * NN + k];
      if (f) v *= f[l2g[e * NN + k]];
    }
    loc[i] = v;
  }
}

template <int PE, int MODE>
struct PatchLaunch {
  // CTAs per SM of this instantiation for the given dynamic shared memory
  template <int N>
  static int occupancy(size_t smem, int *per_sm, int *sms) {
    static size_t configured = 0;  // per instantiation; one device per process
    static int cached_per_sm = 0, cached_sms = 0;
    auto kern = patch_kernel<N, PE, MODE>;
    if (smem > configured || cached_per_sm == 0) {
      SEMK_CUDA_CHECK(
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SEMK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
          &cached_per_sm, kern, PatchCfg<N, PE>::kThreads, smem));
      int dev = 0;
      SEMK_CUDA_CHECK(cudaGetDevice(&dev));
      SEMK_CUDA_CHECK(cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev));
      configured = smem;
    }
    *per_sm = cached_per_sm;
    *sms = cached_sms;
    return SEMK_OK;
  }

  template <int N>
  static int run(const semk_op &op, const DMatEO &dm, const double *u, const double *loc,
                 double *y, int flags, double fill, double *partials, cudaStream_t st,
                 int *grid_out, int64_t pb, int64_t pe) {
    const size_t smem = patch_smem_layout(N, PE, MODE, op.g_patch_stride, op.pn_patch_stride,
                                          op.eloc_patch_stride, op.inv_patch_stride)
                            .total;
    if (smem > 227 * 1024) {
      semk_set_error("patch kernel: shared memory request exceeds 227 KB");
      return SEMK_ERR_UNSUPPORTED;
    }
    int per_sm = 0, sms = 0;
    int rc = occupancy<N>(smem, &per_sm, &sms);
    if (rc != SEMK_OK) return rc;
    if (per_sm < 1) {
      semk_set_error("patch kernel: does not fit on an SM");
      return SEMK_ERR_UNSUPPORTED;
    }
    // persistent grid: every CTA stays resident and loops over its patches, round-robin
    const int64_t resident = (int64_t)per_sm * sms;
    const int64_t np = pe - pb;
    // (op.max_ctas: the caller may cap the grid, e.g. to keep it from being a multiple of
    // the number of patches per tile column -- see semk.h)
    const int64_t want = (op.max_ctas > 0 && op.max_ctas < resident) ? op.max_ctas : resident;
    const unsigned grid = (unsigned)(np < want ? np : want);
    if (grid_out) *grid_out = (int)grid;
    if (grid == 0) return SEMK_OK;
    patch_kernel<N, PE, MODE><<<grid, PatchCfg<N, PE>::kThreads, smem, st>>>(
        op, dm, u, loc, y, flags, fill, partials, pb, pe);
    SEMK_LAUNCH_CHECK("patch_kernel");
    return SEMK_OK;
  }
};

// elements per patch supported by the compiled kernels
inline bool pe_supported(int pe) { return pe == 4 || pe == 8 || pe == 16 || pe == 32; }

// 32-element patches (4 x 8 tiles) exist for the low orders only (n1 <= kBigPatchMaxN1):
// small elements amortise the per-patch barriers and tables better over a larger patch and
// have a smaller share of interface nodes.
constexpr int kBigPatchMaxN1 = 7;
template <int NV, int MODE, bool OK = (NV <= kBigPatchMaxN1)>
struct BigPatch {
  static int run(const semk_op &op, const DMatEO &dm, const double *u, const double *loc,
                 double *y, int flags, double fill, double *partials, cudaStream_t st,
                 int *grid_out, int64_t pb, int64_t pe) {
    return PatchLaunch<32, MODE>::template run<NV>(op, dm, u, loc, y, flags, fill, partials, st,
                                                   grid_out, pb, pe);
  }
  static int occupancy(size_t smem, int *per_sm, int *sms) {
    return PatchLaunch<32, MODE>::template occupancy<NV>(smem, per_sm, sms);
  }
};
template <int NV, int MODE>
struct BigPatch<NV, MODE, false> {
  static int run(const semk_op &, const DMatEO &, const double *, const double *, double *, int,
                 double, double *, cudaStream_t, int *, int64_t, int64_t) {
    semk_set_error("32-element patches are compiled for n1 <= 7 only");
    return SEMK_ERR_UNSUPPORTED;
  }
  static int occupancy(size_t, int *, int *) {
    semk_set_error("32-element patches are compiled for n1 <= 7 only");
    return SEMK_ERR_UNSUPPORTED;
  }
};

template <int MODE>
int launch_patch(const semk_op &op, const DMatEO &dm, const double *u, const double *loc,
                 double *y, int flags, double fill, double *partials, cudaStream_t st,
                 int *grid_out, int64_t pb = 0, int64_t pe = -1) {
  if (pe < 0) pe = op.n_patch;
  if (MODE == MODE_APPLY && op.kernel_variant == 1)
    return semk_ho_launch(op, u, y, flags, partials, st, grid_out, pb, pe);
#define SEMK_CALL(NV)                                                                     \
  do {                                                                                    \
    int rc;                                                                               \
    if (op.elems_per_patch == 32)                                                         \
      rc = BigPatch<NV, MODE>::run(op, dm, u, loc, y, flags, fill, partials, st, grid_out, pb, pe); \
    else if (op.elems_per_patch == 16)                                                    \
      rc = PatchLaunch<16, MODE>::template run<NV>(op, dm, u, loc, y, flags, fill, partials, st, \
                                                   grid_out, pb, pe); \
    else if (op.elems_per_patch == 8)                                                     \
      rc = PatchLaunch<8, MODE>::template run<NV>(op, dm, u, loc, y, flags, fill, partials, st,   \
                                                  grid_out, pb, pe);  \
    else                                                                                  \
      rc = PatchLaunch<4, MODE>::template run<NV>(op, dm, u, loc, y, flags, fill, partials, st,   \
                                                  grid_out, pb, pe);  \
    if (rc != SEMK_OK) return rc;                                                         \
  } while (0)
  SEMK_DISPATCH_N1(op.n1, SEMK_CALL)
#undef SEMK_CALL
  return SEMK_OK;
}

int check_op(const semk_op *op, const char *who) {
  if (!op) {
    semk_set_error(std::string(who) + ": null operator");
    return SEMK_ERR_INVALID;
  }
  if (op->n1 < 2 || op->n1 > SEMK_MAX_N1) {
    semk_set_error(std::string(who) + ": n1 outside [2, 17]");
    return SEMK_ERR_UNSUPPORTED;
  }
  if (!pe_supported(op->elems_per_patch)) {
    semk_set_error(std::string(who) + ": elems_per_patch must be 4, 8, 16 or 32");
    return SEMK_ERR_UNSUPPORTED;
  }
  const int64_t nnp = (int64_t)op->n1 * op->n1 * op->elems_per_patch;
  if (!op->pnode || !op->eloc || !op->patch_hdr || !op->inv || (op->pn_patch_stride & 3) != 0 ||
      op->inv_width < 4 || (op->inv_width & 3) != 0 ||
      op->inv_patch_stride != op->pn_patch_stride * op->inv_width ||
      op->pn_patch_stride < op->max_patch_nodes ||
      (reinterpret_cast<uintptr_t>(op->patch_hdr) & 15u) != 0 ||
      (op->eloc_patch_stride & 7) != 0 ||
      op->eloc_patch_stride < nnp ||
      (op->n_slots > 0 && !op->slot_buf) ||
      (op->n_shared > 0 && (!op->shared_rec || !op->shared_ext)) ||
      (op->n_shared_chunk > 0 && !op->shared_chunk)) {
    semk_set_error(std::string(who) + ": operator tables incomplete");
    return SEMK_ERR_INVALID;
  }
  if ((op->g_patch_stride & 1) != 0 || op->g_patch_stride < 3 * nnp) {
    semk_set_error(std::string(who) +
                   ": g_patch_stride must be even (16-byte TMA granularity) and >= 3*NN*PE");
    return SEMK_ERR_INVALID;
  }
  return SEMK_OK;
}

}  // namespace

extern "C" int64_t semk_partials_len(int64_t n_patch, int64_t n_shared) {
  // persistent grid (<= n_patch) + chunk CTAs (<= n_shared / 2 / 8 + 1) + record CTAs
  return n_patch + n_shared / (2 * kChunkWarps) + kSharedBlocks + 16;
}

extern "C" int64_t semk_resident_ctas(int n1, int elems_per_patch, int64_t g_patch_stride,
                                      int64_t pn_patch_stride, int64_t eloc_patch_stride,
                                      int64_t inv_patch_stride) {
  if (!pe_supported(elems_per_patch)) return -1;
  const size_t smem = patch_smem_layout(n1, elems_per_patch, MODE_APPLY, g_patch_stride,
                                        pn_patch_stride, eloc_patch_stride, inv_patch_stride)
                          .total;
  if (smem > 227 * 1024) return -1;
  int per_sm = 0, sms = 0;
  auto run = [&]() -> int {
#define SEMK_CALL(NV)                                                                  \
  do {                                                                                 \
    int rc;                                                                            \
    if (elems_per_patch == 32)                                                         \
      rc = BigPatch<NV, MODE_APPLY>::occupancy(smem, &per_sm, &sms);                   \
    else if (elems_per_patch == 16)                                                    \
      rc = PatchLaunch<16, MODE_APPLY>::template occupancy<NV>(smem, &per_sm, &sms);   \
    else if (elems_per_patch == 8)                                                     \
      rc = PatchLaunch<8, MODE_APPLY>::template occupancy<NV>(smem, &per_sm, &sms);    \
    else                                                                               \
      rc = PatchLaunch<4, MODE_APPLY>::template occupancy<NV>(smem, &per_sm, &sms);    \
    if (rc != SEMK_OK) return rc;                                                      \
  } while (0)
    SEMK_DISPATCH_N1(n1, SEMK_CALL)
#undef SEMK_CALL
    return SEMK_OK;
  };
  if (run() != SEMK_OK) return -1;
  return (int64_t)per_sm * sms;
}

extern "C" int64_t semk_resident_ctas_variant(int kernel_variant, int n1, int elems_per_patch,
                                              int64_t g_patch_stride, int64_t pn_patch_stride,
                                              int64_t eloc_patch_stride,
                                              int64_t inv_patch_stride) {
  if (kernel_variant == 0)
    return semk_resident_ctas(n1, elems_per_patch, g_patch_stride, pn_patch_stride,
                              eloc_patch_stride, inv_patch_stride);
  if (kernel_variant != 1 || !pe_supported(elems_per_patch) || elems_per_patch == 32) return -1;
  const size_t smem = patch_smem_layout(n1, elems_per_patch, MODE_APPLY, g_patch_stride,
                                        pn_patch_stride, eloc_patch_stride, inv_patch_stride)
                          .total;
  if (smem > 227 * 1024) return -1;
  return semk_ho_resident(n1, elems_per_patch, smem);
}

extern "C" int64_t semk_patch_smem_bytes(int n1, int elems_per_patch, int64_t g_patch_stride,
                                         int64_t pn_patch_stride, int64_t eloc_patch_stride,
                                         int64_t inv_patch_stride) {
  return (int64_t)patch_smem_layout(n1, elems_per_patch, MODE_APPLY, g_patch_stride,
                                    pn_patch_stride, eloc_patch_stride, inv_patch_stride)
      .total;
}

extern "C" int semk_poisson_apply_f64(const semk_op *op, const double *u, double *y, int flags,
                                      double *dot_out, void *stream) {
  int rc = check_op(op, "semk_poisson_apply_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(u && y && u != y, "semk_poisson_apply_f64: u and y must be distinct buffers");
  SEMK_REQUIRE(op->G && op->D_host, "semk_poisson_apply_f64: missing G or D");
  SEMK_REQUIRE(!dot_out || op->partials, "semk_poisson_apply_f64: dot_out needs op->partials");
  cudaStream_t st = semk_stream(stream);
  DMatEO dm;
  if (!make_dmat_eo(op->n1, op->D_host, &dm)) {
    semk_set_error(
        "semk_poisson_apply_f64: differentiation matrix is not centro-antisymmetric "
        "(the engine needs a basis whose nodes are symmetric about 0)");
    return SEMK_ERR_UNSUPPORTED;
  }
  double *partials = dot_out ? op->partials : nullptr;
  int grid = 0;
  rc = launch_patch<MODE_APPLY>(*op, dm, u, nullptr, y, flags, 0.0, partials, st, &grid);
  if (rc != SEMK_OK) return rc;
  int shared_blocks = 0;
  rc = launch_interface<MODE_APPLY>(*op, u, y, flags, 0.0, partials, grid, st, &shared_blocks);
  if (rc != SEMK_OK) return rc;
  if (dot_out) {
    reduce_partials_kernel<<<1, 1024, 0, st>>>(partials, grid + shared_blocks, dot_out);
    SEMK_LAUNCH_CHECK("reduce_partials_kernel");
  }
  return SEMK_OK;
}

// A sub-range of the apply: patches [patch_begin, patch_end) and the interface entries
// [chunk_begin, chunk_end) / [rec_begin, rec_end) (both tables are sorted by the highest
// patch involved, so "everything patches [0, P) complete" is a prefix of each).
extern "C" int semk_poisson_apply_range_f64(const semk_op *op, const double *u, double *y,
                                            int flags, int64_t patch_begin, int64_t patch_end,
                                            int64_t chunk_begin, int64_t chunk_end,
                                            int64_t rec_begin, int64_t rec_end, void *stream) {
  int rc = check_op(op, "semk_poisson_apply_range_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(u && y && u != y && op->G && op->D_host,
               "semk_poisson_apply_range_f64: bad arguments");
  SEMK_REQUIRE(0 <= patch_begin && patch_begin <= patch_end && patch_end <= op->n_patch &&
                   0 <= chunk_begin && chunk_begin <= chunk_end &&
                   chunk_end <= op->n_shared_chunk && 0 <= rec_begin && rec_begin <= rec_end &&
                   rec_end <= op->n_shared,
               "semk_poisson_apply_range_f64: range outside the operator");
  cudaStream_t st = semk_stream(stream);
  DMatEO dm;
  if (!make_dmat_eo(op->n1, op->D_host, &dm)) {
    semk_set_error("semk_poisson_apply_range_f64: D is not centro-antisymmetric");
    return SEMK_ERR_UNSUPPORTED;
  }
  if (patch_end > patch_begin) {
    rc = launch_patch<MODE_APPLY>(*op, dm, u, nullptr, y, flags, 0.0, nullptr, st, nullptr,
                                  patch_begin, patch_end);
    if (rc != SEMK_OK) return rc;
  }
  const InterfaceRange r{chunk_begin, chunk_end, rec_begin, rec_end};
  return launch_interface<MODE_APPLY>(*op, u, y, flags, 0.0, nullptr, 0, st, nullptr, &r);
}

extern "C" int semk_assemble_f64(const semk_op *op, const double *loc, double *out, int flags,
                                 double fill_dirichlet, void *stream) {
  int rc = check_op(op, "semk_assemble_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(loc && out, "semk_assemble_f64: null pointer");
  cudaStream_t st = semk_stream(stream);
  DMatEO dm;
  std::memset(&dm, 0, sizeof(dm));
  rc = launch_patch<MODE_ASSEMBLE>(*op, dm, nullptr, loc, out, flags, fill_dirichlet, nullptr, st,
                                   nullptr);
  if (rc != SEMK_OK) return rc;
  return launch_interface<MODE_ASSEMBLE>(*op, nullptr, out, flags, fill_dirichlet, nullptr, 0, st,
                                         nullptr);
}

extern "C" int semk_poisson_local_diag_f64(const semk_op *op, const double *D_dev, double *loc,
                                           void *stream) {
  int rc = check_op(op, "semk_poisson_local_diag_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(loc && op->G && D_dev, "semk_poisson_local_diag_f64: null pointer");
  cudaStream_t st = semk_stream(stream);
  const int64_t n_slot_elems = op->n_patch * op->elems_per_patch;
  const int64_t total = n_slot_elems * op->n1 * op->n1;
  const int64_t want = (total + 255) / 256;
  const unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
  local_diag_kernel<<<grid, 256, 0, st>>>(op->n1, op->elems_per_patch, n_slot_elems, op->G,
                                          op->g_patch_stride, D_dev, loc);
  SEMK_LAUNCH_CHECK("local_diag_kernel");
  return SEMK_OK;
}

extern "C" int semk_weighted_local_f64(int n1, int64_t n_elem, int64_t n_slot_elems,
                                       const double *JxW, const uint32_t *l2g,
                                       const int64_t *elem_of_slot, const double *f, double *loc,
                                       void *stream) {
  SEMK_REQUIRE(n1 >= 2 && n1 <= SEMK_MAX_N1, "semk_weighted_local_f64: bad n1");
  SEMK_REQUIRE(JxW && loc && (!f || l2g), "semk_weighted_local_f64: null pointer");
  SEMK_REQUIRE(n_slot_elems >= n_elem, "semk_weighted_local_f64: n_slot_elems < n_elem");
  cudaStream_t st = semk_stream(stream);
  const int NN = n1 * n1;
  if (n_slot_elems > n_elem)
    SEMK_CUDA_CHECK(cudaMemsetAsync(loc + n_elem * NN, 0,
                                    sizeof(double) * (n_slot_elems - n_elem) * NN, st));
  if (n_elem <= 0) return SEMK_OK;
  const int64_t total = n_elem * NN;
  const int64_t want = (total + 255) / 256;
  const unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
  weighted_local_kernel<<<grid, 256, 0, st>>>(NN, n_elem, JxW, l2g, elem_of_slot, f, loc);
  SEMK_LAUNCH_CHECK("weighted_local_kernel");
  return SEMK_OK;
}

extern "C" int semk_poisson_apply_atomic_f64(int n1, int64_t n_elem, int64_t n_nodes,
                                             const uint32_t *l2g, const int64_t *elem_of_slot,
                                             const double *G, int64_t g_patch_stride,
                                             int elems_per_patch, const double *D_host,
                                             const uint8_t *dirichlet, const double *u, double *y,
                                             int flags, void *stream) {
  SEMK_REQUIRE(l2g && G && D_host && u && y && u != y,
               "semk_poisson_apply_atomic_f64: null or aliased pointer");
  SEMK_REQUIRE(n_elem > 0 && n_nodes > 0 && elems_per_patch > 0,
               "semk_poisson_apply_atomic_f64: empty problem");
  SEMK_REQUIRE(n1 >= 2 && n1 <= SEMK_MAX_N1, "semk_poisson_apply_atomic_f64: bad n1");
  cudaStream_t st = semk_stream(stream);
  const DMat dm = make_dmat(n1, D_host);
  SEMK_CUDA_CHECK(cudaMemsetAsync(y, 0, sizeof(double) * n_nodes, st));
  constexpr int PEA = 8;
  const unsigned grid = (unsigned)((n_elem + PEA - 1) / PEA);
#define SEMK_CALL(NV)                                                                       \
  atomic_kernel<NV, PEA><<<grid, PatchCfg<NV, PEA>::kThreads, 0, st>>>(                     \
      dm, n_elem, l2g, elem_of_slot, G, g_patch_stride, elems_per_patch, dirichlet, u, y, flags)
  SEMK_DISPATCH_N1(n1, SEMK_CALL)
#undef SEMK_CALL
  SEMK_LAUNCH_CHECK("atomic_kernel");
  if (dirichlet && (flags & SEMK_MASK_OUT)) {
    const int64_t want = (n_nodes + 255) / 256;
    const unsigned g2 = (unsigned)(want < 148 * 16 ? want : 148 * 16);
    dirichlet_fix_kernel<<<g2, 256, 0, st>>>(n_nodes, dirichlet, u, y, flags);
    SEMK_LAUNCH_CHECK("dirichlet_fix_kernel");
  }
  return SEMK_OK;
}

// Copy streams / events of the staged host apply, one set per device, created on
// first use (never destroyed: they live as long as the process).
namespace {
constexpr int kMaxStages = 64;
struct HostPipe {
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t start = nullptr, up[kMaxStages] = {}, done[kMaxStages] = {}, finish = nullptr;
  bool ready = false;
};
HostPipe *host_pipe() {
  static HostPipe pipes[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  HostPipe &p = pipes[dev];
  if (!p.ready) {
    if (cudaStreamCreateWithFlags(&p.h2d, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithFlags(&p.d2h, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    auto mk = [](cudaEvent_t *e) {
      return cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
    };
    if (!mk(&p.start) || !mk(&p.finish)) return nullptr;
    for (int i = 0; i < kMaxStages; ++i)
      if (!mk(&p.up[i]) || !mk(&p.done[i])) return nullptr;
    p.ready = true;
  }
  return &p;
}
}  // namespace

extern "C" int semk_poisson_apply_host_staged_f64(const semk_op *op, const semk_stage *stages,
                                                  int n_stages, const double *u_host,
                                                  double *y_host, double *d_u, double *d_y,
                                                  int flags, void *stream) {
  int rc = check_op(op, "semk_poisson_apply_host_staged_f64");
  if (rc != SEMK_OK) return rc;
  SEMK_REQUIRE(stages && n_stages >= 1 && n_stages <= kMaxStages,
               "semk_poisson_apply_host_staged_f64: 1..64 stages");
  SEMK_REQUIRE(u_host && y_host && d_u && d_y && d_u != d_y,
               "semk_poisson_apply_host_staged_f64: null / aliased buffers");
  SEMK_REQUIRE(op->G && op->D_host, "semk_poisson_apply_host_staged_f64: missing G or D");
  {
    int64_t pp = 0, pc = 0, pr = 0, pu = 0, py = 0;
    for (int i = 0; i < n_stages; ++i) {
      const semk_stage &s = stages[i];
      SEMK_REQUIRE(s.patch_end >= pp && s.chunk_end >= pc && s.rec_end >= pr && s.u_need >= pu &&
                       s.y_final >= py && s.u_need <= op->n_nodes && s.y_final <= op->n_nodes,
                   "semk_poisson_apply_host_staged_f64: stage table not monotone");
      pp = s.patch_end, pc = s.chunk_end, pr = s.rec_end, pu = s.u_need, py = s.y_final;
    }
    SEMK_REQUIRE(pp == op->n_patch && pc == op->n_shared_chunk && pr == op->n_shared &&
                     pu == op->n_nodes && py == op->n_nodes,
                 "semk_poisson_apply_host_staged_f64: last stage must complete the operator");
  }
  DMatEO dm;
  if (!make_dmat_eo(op->n1, op->D_host, &dm)) {
    semk_set_error("semk_poisson_apply_host_staged_f64: D is not centro-antisymmetric");
    return SEMK_ERR_UNSUPPORTED;
  }
  HostPipe *P = host_pipe();
  SEMK_REQUIRE(P, "semk_poisson_apply_host_staged_f64: could not create copy streams");
  cudaStream_t st = semk_stream(stream);
  // the copy streams start after whatever the caller queued on `stream`
  SEMK_CUDA_CHECK(cudaEventRecord(P->start, st));
  SEMK_CUDA_CHECK(cudaStreamWaitEvent(P->h2d, P->start, 0));
  SEMK_CUDA_CHECK(cudaStreamWaitEvent(P->d2h, P->start, 0));
  int64_t pb = 0, cb = 0, rb = 0, ub = 0, yb = 0;
  for (int i = 0; i < n_stages; ++i) {
    const semk_stage &s = stages[i];
    // upload the part of u this stage's patches read ...
    if (s.u_need > ub)
      SEMK_CUDA_CHECK(cudaMemcpyAsync(d_u + ub, u_host + ub, sizeof(double) * (size_t)(s.u_need - ub),
                                      cudaMemcpyHostToDevice, P->h2d));
    SEMK_CUDA_CHECK(cudaEventRecord(P->up[i], P->h2d));
    SEMK_CUDA_CHECK(cudaStreamWaitEvent(st, P->up[i], 0));
    // ... run its patches and the interface entries they complete ...
    rc = launch_patch<MODE_APPLY>(*op, dm, d_u, nullptr, d_y, flags, 0.0, nullptr, st, nullptr, pb,
                                  s.patch_end);
    if (rc != SEMK_OK) return rc;
    const InterfaceRange r{cb, s.chunk_end, rb, s.rec_end};
    rc = launch_interface<MODE_APPLY>(*op, d_u, d_y, flags, 0.0, nullptr, 0, st, nullptr, &r);
    if (rc != SEMK_OK) return rc;
    SEMK_CUDA_CHECK(cudaEventRecord(P->done[i], st));
    // ... and download the part of y that is final now
    SEMK_CUDA_CHECK(cudaStreamWaitEvent(P->d2h, P->done[i], 0));
    if (s.y_final > yb)
      SEMK_CUDA_CHECK(cudaMemcpyAsync(y_host + yb, d_y + yb, sizeof(double) * (size_t)(s.y_final - yb),
                                      cudaMemcpyDeviceToHost, P->d2h));
    pb = s.patch_end, cb = s.chunk_end, rb = s.rec_end, ub = s.u_need, yb = s.y_final;
  }
  SEMK_CUDA_CHECK(cudaEventRecord(P->finish, P->d2h));
  SEMK_CUDA_CHECK(cudaStreamWaitEvent(st, P->finish, 0));
  SEMK_CUDA_CHECK(cudaStreamSynchronize(st));
  return SEMK_OK;
}

extern "C" int semk_poisson_apply_host_f64(const semk_op *op, const double *u_host,
                                           double *y_host, double *d_u, double *d_y, int flags,
                                           void *stream) {
  SEMK_REQUIRE(op && u_host && y_host && d_u && d_y, "semk_poisson_apply_host_f64: null pointer");
  cudaStream_t st = semk_stream(stream);
  const size_t bytes = sizeof(double) * (size_t)op->n_nodes;
  SEMK_CUDA_CHECK(cudaMemcpyAsync(d_u, u_host, bytes, cudaMemcpyHostToDevice, st));
  int rc = semk_poisson_apply_f64(op, d_u, d_y, flags, nullptr, st);
  if (rc != SEMK_OK) return rc;
  SEMK_CUDA_CHECK(cudaMemcpyAsync(y_host, d_y, bytes, cudaMemcpyDeviceToHost, st));
  SEMK_CUDA_CHECK(cudaStreamSynchronize(st));
  return SEMK_OK;
}
