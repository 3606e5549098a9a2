// semk_ho.cu -- the Poisson apply for high orders: column / row thread PAIRS.
//
// Same operator, same plan tables, same persistent / TMA-staged / gather-assembled
// structure as patch_kernel (semk_apply.cu); what changes is the thread mapping of the
// element operator.  patch_kernel gives every element column to ONE thread, which then does
// all four 1-D contractions of that column / row in turn: at p >= 10 that is > 160 registers
// per thread and, with N*PE threads per CTA, 9-12 warps per SM -- the kernel is bound by
// instruction latency, not by DRAM or the FP64 pipe (profiles/r02_ncu_base_p16_summary.txt:
// 8.4 warps/SM, issue active 46 %, FP64 pipe 29 %, DRAM 34 %).
//
// Here every column t of an element is owned by a PAIR of adjacent lanes:
//   lane "col" holds the column u[.][t], does the two contractions along the column
//     (ur = D u, y0 = D^T w0) and the collocated products with the geometric factors;
//   lane "row" reads row t from the transpose scratch and does the two contractions along
//     the row (us = u D^T, y1 = w1 D).
// Both lanes run the SAME even-odd contraction code on their own vector at the same time
// (no divergence in the DFMA stream, the D operand stays a uniform constant-bank operand),
// threads per CTA double and registers per thread drop, so twice the warps are resident.
// Reference lines replaced: as semk_apply.cu (examples/poisson.py:166-193 local stiffness,
// examples/squirmer-axisymmetric.py:268-295 local apply, sem/discrete.py:491-510).
#include "semk_elem.cuh"

namespace {

__host__ __device__ constexpr int ho_threads(int N, int PE) { return ((2 * N * PE + 31) / 32) * 32; }

// resident CTAs per SM the register budget is set for
__host__ __device__ constexpr int ho_min_blocks(int N, int PE) {
  const long long nn = (long long)N * N, p = N - 1;
  const int bx = PE == 16 ? 2 : 1, by = PE == 4 ? 4 : 8;
  const long long mpn4 = (((long long)(bx * p + 1) * (by * p + 1)) + 3) & ~3LL;
  const long long g = 8 * ((3 * nn * PE + 1) & ~1LL);
  const long long tab = 2 * ((4 * mpn4 + 2 * ((nn * PE + 7) & ~7LL) + 15) & ~15LL) + 64;
  const long long scr = 8LL * N * scratch_row_stride(N, PE);
  const long long total = 32 + g + tab + 8 * mpn4 + 2 * scr + 1024;
  const long long by_smem = 233472 / total;
  const long long by_regs = 65536 / ((long long)ho_threads(N, PE) * 104);
  const long long r = by_smem < by_regs ? by_smem : by_regs;
  return r < 1 ? 1 : (r > 8 ? 8 : (int)r);
}

constexpr int kHoGatherBatch = 4;

// The differentiation matrices as a kernel parameter, placed 8 bytes PAST a 16-byte boundary of
// the parameter block on purpose: at p = 16 the kernel runs 8 % faster with the even-odd
// tables there than on the boundary itself (0.986 vs 1.068 ms per apply, same box, both
// twice: profiles/r02_box_ab.txt, call 66; the column kernels do not care).  The alignas
// makes the position independent of sizeof(semk_op).
struct alignas(16) HoDMat {
  double pad;
  DMatEO dm;
};

template <int N, int PE, bool DOT>
__global__ void __launch_bounds__(ho_threads(N, PE), ho_min_blocks(N, PE))
    ho_patch_kernel(semk_op op, HoDMat hd, const double *__restrict__ u, double *__restrict__ y,
                    int flags, double *__restrict__ dot_partials, int64_t patch_begin,
                    int64_t patch_end) {
  constexpr int NP = N * PE;
  constexpr int kThreads = ho_threads(N, PE);
  constexpr int RS = scratch_row_stride(N, PE);
  constexpr int H = (N + 1) / 2;  // rows gathered by the column lane; the row lane takes N - H
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const PatchSmem L = patch_smem_layout(N, PE, MODE_APPLY, op.g_patch_stride, op.pn_patch_stride,
                                        op.eloc_patch_stride, op.inv_patch_stride);
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);
  double *Gs = reinterpret_cast<double *>(smem_raw + L.gs);
  const uint16_t *inv_s = reinterpret_cast<const uint16_t *>(smem_raw + L.inv);
  double *As = reinterpret_cast<double *>(smem_raw + L.ua);
  double *Bs = reinterpret_cast<double *>(smem_raw + L.bs);
  double *red = reinterpret_cast<double *>(smem_raw + L.red);

  const int tid = threadIdx.x;
  const int pair = tid >> 1;
  const bool is_row = (tid & 1) != 0;
  const int le = pair / N, t = pair - le * N;
  const int tidp = le * N + t;
  const uint32_t pn_bytes = 4u * (uint32_t)op.pn_patch_stride;
  const uint32_t el_bytes = 2u * (uint32_t)op.eloc_patch_stride;
  const uint32_t g_bytes = (uint32_t)(op.g_patch_stride * sizeof(double));
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
  auto issue_g = [&](int64_t patch) {
    semk_mbar_expect_tx(&mbar[2], g_bytes);
    semk_bulk_g2s(Gs, op.G + patch * op.g_patch_stride, g_bytes, &mbar[2]);
  };

  const int64_t step = (int64_t)gridDim.x;
  const int64_t p_first = patch_begin + (int64_t)blockIdx.x;
  const int64_t p_end = patch_end;
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
      issue_g(p_first);
    }
  }
  __syncthreads();

  // Each lane of a pair gathers half of the column (col lane: rows [0, H), row lane: rows
  // [H, N)) for the patch about to be processed and carries it in registers.
  const int m0 = is_row ? H : 0;
  const int mcount = is_row ? (N - H) : H;
  double uh[H];
  uint32_t uh_dir = 0;  // bit m: node (m, t) is a Dirichlet node (this lane's rows only)
  auto gather_half = [&](int s_tab, int64_t patch_of) {
    const unsigned char *sbn = stage_ptr(s_tab);
    const uint32_t *pnb = reinterpret_cast<const uint32_t *>(sbn + L.pn_off);
    const uint32_t id0 = hdr_ring[8 * s_tab + 4];
    const uint16_t *elb = reinterpret_cast<const uint16_t *>(sbn + L.el_off);
    const bool act = (le < PE) && (patch_of * PE + le < op.n_elem);
    if (act) {
      uint32_t pn[H];
#pragma unroll
      for (int k = 0; k < H; ++k)
        pn[k] = (k < mcount) ? pnb[elb[(m0 + k) * NP + tidp]] : 0u;
      uh_dir = 0;
#pragma unroll
      for (int k = 0; k < H; ++k) {
        if (k < mcount) {
          uh[k] = u[id0 + (pn[k] & SEMK_NODE_ID_MASK)];
          uh_dir |= (pn[k] >> 31) << (m0 + k);
        }
      }
    }
  };
  if (p_first < p_end) {
    semk_mbar_wait(&mbar[0], 0);
    gather_half(0, p_first);
  }
  __syncthreads();

  double dot = 0.0;
  constexpr bool want_dot = DOT;
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
    const uint32_t col_dir = DOT ? (uh_dir | __shfl_xor_sync(0xffffffffu, uh_dir, 1)) : 0u;

    // ---- element operator -----------------------------------------------------------
    if (active) {
#pragma unroll
      for (int k = 0; k < H; ++k)
        if (k < mcount) {
          const bool z = (flags & SEMK_MASK_IN) && ((uh_dir >> (m0 + k)) & 1u);
          As[(m0 + k) * RS + tidp] = z ? 0.0 : uh[k];
        }
    }
    __syncthreads();  // #1: u of the patch in scratch A
    if (tid == 0) issue_inv(inv_block);
    if (tid == 0 && has_next) {
      const uint32_t *hn = hdr_ring + 8 * (s ^ 1);
      const int64_t after = next + step;
      issue_tables(hn[5], hn[6], s ^ 1, after < p_end ? after : -1, s);
    }
    double v[N], o[N];
    if (active) {
      // col lane: column t (stride RS);  row lane: row t (contiguous)
      const double *src = is_row ? (As + t * RS + le * N) : (As + tidp);
      const int stride = is_row ? 1 : RS;
#pragma unroll
      for (int k = 0; k < N; ++k) v[k] = src[k * stride];
      mat_D<N>(hd.dm, v, o);
      if (is_row) {
#pragma unroll
        for (int n = 0; n < N; ++n) Bs[t * RS + le * N + n] = o[n];  // us[t][n]
      }
      // (col lane: o = ur, the derivative along the column)
    }
    __syncthreads();  // #2: us in scratch B
    semk_mbar_wait(&mbar[2], (uint32_t)(it & 1));
    if (active && !is_row) {
      const double *g = Gs + tidp;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const double usc = Bs[m * RS + tidp];
        const double g00 = g[m * NP], g01 = g[(N + m) * NP], g11 = g[(2 * N + m) * NP];
        const double urm = o[m];
        o[m] = g00 * urm + g01 * usc;                    // w0[m][t]
        As[m * RS + tidp] = g01 * urm + g11 * usc;       // w1[m][t]
      }
    }
    __syncthreads();  // #3: w1 in scratch A; every thread is done with G
    if (tid == 0 && has_next) issue_g(next);
    double ycol[N];
    if (active) {
      if (is_row) {
#pragma unroll
        for (int n = 0; n < N; ++n) o[n] = As[t * RS + le * N + n];  // row t of w1
      }
      mat_Dt<N>(hd.dm, o, ycol);  // col lane: y0[.][t] = D^T w0;  row lane: y1[t][.] = w1 D
      if (is_row) {
#pragma unroll
        for (int q = 0; q < N; ++q) Bs[t * RS + le * N + q] = ycol[q];
      }
    }
    __syncthreads();  // #4: y1 in scratch B
    if (active && !is_row) {
#pragma unroll
      for (int m = 0; m < N; ++m) {
        ycol[m] += Bs[m * RS + tidp];
        Bs[m * RS + tidp] = ycol[m];  // element result, where the write-out gathers from
      }
      if (DOT) {
#pragma unroll
        for (int m = 0; m < N; ++m) {
          const bool dir = (col_dir >> m) & 1u;
          const bool zero_in = (flags & SEMK_MASK_IN) && dir;
          const bool drop = (flags & SEMK_MASK_OUT) && dir;
          dot = fma((zero_in || drop) ? 0.0 : v[m], ycol[m], dot);
        }
      }
    }
    if (has_next) {
      semk_mbar_wait(&mbar[s ^ 1], (uint32_t)(((it + 1) >> 1) & 1));
      gather_half(s ^ 1, next);
    }
    // ---- assemble by gathering, fixed order (as patch_kernel) -----------------------------
    semk_mbar_wait(&mbar[3], (uint32_t)(it & 1));
    __syncthreads();  // #5
    for (int k0 = tid; k0 < npn; k0 += kHoGatherBatch * kThreads) {
      uint32_t pnv[kHoGatherBatch];
      uint2 ev[kHoGatherBatch];
#pragma unroll
      for (int j = 0; j < kHoGatherBatch; ++j) {
        const int k = k0 + j * kThreads;
        const bool in = k < npn;
        pnv[j] = in ? pn_s[k] : 0xffffffffu;
        ev[j] = in ? reinterpret_cast<const uint2 *>(inv_s)[(size_t)k * inv_w4]
                   : make_uint2(0xffffffffu, 0xffffffffu);
      }
#pragma unroll
      for (int j = 0; j < kHoGatherBatch; ++j) {
        const int k = k0 + j * kThreads;
        if (k >= npn) continue;
        const uint32_t e0 = ev[j].x & 0xffffu, e1 = ev[j].x >> 16;
        const uint32_t e2 = ev[j].y & 0xffffu, e3 = ev[j].y >> 16;
        double val = (e0 != 0xffffu) ? Bs[e0] : 0.0;
        if (e1 != 0xffffu) val += Bs[e1];
        if (e2 != 0xffffu) val += Bs[e2];
        if (e3 != 0xffffu) val += Bs[e3];
        for (int w = 1; w < inv_w4; ++w) {
          const uint2 e = reinterpret_cast<const uint2 *>(inv_s)[(size_t)k * inv_w4 + w];
          const uint32_t q[4] = {e.x & 0xffffu, e.x >> 16, e.y & 0xffffu, e.y >> 16};
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (q[i] != 0xffffu) val += Bs[q[i]];
        }
        const uint32_t pn = pnv[j];
        if (k < npriv) {
          const uint32_t g = id0 + (pn & SEMK_NODE_ID_MASK);
          if ((pn & SEMK_NODE_DIRICHLET) && (flags & SEMK_MASK_OUT)) {
            val = 0.0;
            if (flags & SEMK_DIRICHLET_IDENTITY) {
              val = u[g];
              if (DOT) dot = fma(val, val, dot);
            }
          }
          y[g] = val;
        } else {
          op.slot_buf[slot_base + (k - npriv)] = val;
        }
      }
    }
  }
  if (DOT) {
    __syncthreads();
    const double sres = semk_block_sum(dot, red);
    if (tid == 0) dot_partials[blockIdx.x] = sres;
  }
}

template <int N, int PE>
struct HoLaunch {
  static int occupancy(size_t smem, int *per_sm, int *sms) {
    static size_t configured = 0;
    static int cached_per_sm = 0, cached_sms = 0;
    auto kern = ho_patch_kernel<N, PE, true>;
    if (smem > configured || cached_per_sm == 0) {
      SEMK_CUDA_CHECK(
          cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SEMK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cached_per_sm, kern,
                                                                    ho_threads(N, PE), smem));
      int dev = 0;
      SEMK_CUDA_CHECK(cudaGetDevice(&dev));
      SEMK_CUDA_CHECK(cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev));
      configured = smem;
    }
    *per_sm = cached_per_sm;
    *sms = cached_sms;
    return SEMK_OK;
  }
  static int run(const semk_op &op, const DMatEO &dm, const double *u, double *y, int flags,
                 double *partials, cudaStream_t st, int *grid_out, int64_t pb, int64_t pe) {
    const size_t smem = patch_smem_layout(N, PE, MODE_APPLY, op.g_patch_stride,
                                          op.pn_patch_stride, op.eloc_patch_stride,
                                          op.inv_patch_stride)
                            .total;
    if (smem > 227 * 1024) {
      semk_set_error("pair kernel: shared memory request exceeds 227 KB");
      return SEMK_ERR_UNSUPPORTED;
    }
    int per_sm = 0, sms = 0;
    int rc = occupancy(smem, &per_sm, &sms);
    if (rc != SEMK_OK) return rc;
    if (per_sm < 1) {
      semk_set_error("pair kernel: does not fit on an SM");
      return SEMK_ERR_UNSUPPORTED;
    }
    const int64_t resident = (int64_t)per_sm * sms;
    const int64_t np = pe - pb;
    const int64_t want = (op.max_ctas > 0 && op.max_ctas < resident) ? op.max_ctas : resident;
    const unsigned grid = (unsigned)(np < want ? np : want);
    if (grid_out) *grid_out = (int)grid;
    if (grid == 0) return SEMK_OK;
    HoDMat hd;
    hd.pad = 0.0;
    hd.dm = dm;
    if (partials) {
      ho_patch_kernel<N, PE, true><<<grid, ho_threads(N, PE), smem, st>>>(op, hd, u, y, flags,
                                                                          partials, pb, pe);
    } else {
      static bool configured_nodot = false;
      if (!configured_nodot) {
        SEMK_CUDA_CHECK(cudaFuncSetAttribute(ho_patch_kernel<N, PE, false>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024));
        configured_nodot = true;
      }
      ho_patch_kernel<N, PE, false><<<grid, ho_threads(N, PE), smem, st>>>(op, hd, u, y, flags,
                                                                           nullptr, pb, pe);
    }
    SEMK_LAUNCH_CHECK("ho_patch_kernel");
    return SEMK_OK;
  }
};

}  // namespace

// orders the pair kernel is compiled for (n1 = p + 1)
static constexpr int kHoMinN1 = 9;

int semk_ho_launch(const semk_op &op, const double *u, double *y, int flags, double *partials,
                   cudaStream_t st, int *grid_out, int64_t pb, int64_t pe) {
  DMatEO dm;
  if (!make_dmat_eo(op.n1, op.D_host, &dm)) {
    semk_set_error("pair kernel: differentiation matrix is not centro-antisymmetric");
    return SEMK_ERR_UNSUPPORTED;
  }
  if (op.n1 < kHoMinN1 || (op.elems_per_patch != 4 && op.elems_per_patch != 8 &&
                           op.elems_per_patch != 16)) {
    semk_set_error("pair kernel: compiled for n1 >= 9 and 4, 8 or 16 elements per patch");
    return SEMK_ERR_UNSUPPORTED;
  }
  if (pe < 0) pe = op.n_patch;
  int rc = SEMK_ERR_UNSUPPORTED;
#define SEMK_HO(NV)                                                                           \
  case NV:                                                                                    \
    rc = op.elems_per_patch == 16                                                             \
             ? HoLaunch<NV, 16>::run(op, dm, u, y, flags, partials, st, grid_out, pb, pe)     \
             : (op.elems_per_patch == 8                                                       \
                    ? HoLaunch<NV, 8>::run(op, dm, u, y, flags, partials, st, grid_out, pb, pe) \
                    : HoLaunch<NV, 4>::run(op, dm, u, y, flags, partials, st, grid_out, pb, pe)); \
    break;
  switch (op.n1) {
    SEMK_HO(9) SEMK_HO(10) SEMK_HO(11) SEMK_HO(12) SEMK_HO(13) SEMK_HO(14) SEMK_HO(15) SEMK_HO(16)
    SEMK_HO(17)
    default: break;
  }
#undef SEMK_HO
  return rc;
}

int64_t semk_ho_resident(int n1, int elems_per_patch, size_t smem) {
  int per_sm = 0, sms = 0, rc = SEMK_ERR_UNSUPPORTED;
#define SEMK_HO(NV)                                                                       \
  case NV:                                                                                \
    rc = elems_per_patch == 16 ? HoLaunch<NV, 16>::occupancy(smem, &per_sm, &sms)         \
                               : (elems_per_patch == 8                                    \
                                      ? HoLaunch<NV, 8>::occupancy(smem, &per_sm, &sms)   \
                                      : HoLaunch<NV, 4>::occupancy(smem, &per_sm, &sms)); \
    break;
  switch (n1) {
    SEMK_HO(9) SEMK_HO(10) SEMK_HO(11) SEMK_HO(12) SEMK_HO(13) SEMK_HO(14) SEMK_HO(15) SEMK_HO(16)
    SEMK_HO(17)
    default: break;
  }
#undef SEMK_HO
  if (rc != SEMK_OK) return -1;
  return (int64_t)per_sm * sms;
}
