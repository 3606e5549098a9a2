// semk_common.cuh -- shared helpers for the libsemk translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/semk.h"

void semk_set_error(const std::string &msg);

#define SEMK_CUDA_CHECK(expr)                                                        \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      semk_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));            \
      return SEMK_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define SEMK_LAUNCH_CHECK(name)                                                      \
  do {                                                                               \
    cudaError_t _e = cudaGetLastError();                                             \
    if (_e != cudaSuccess) {                                                         \
      semk_set_error(std::string(name) + " launch: " + cudaGetErrorString(_e));      \
      return SEMK_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

#define SEMK_REQUIRE(cond, msg)                                                      \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      semk_set_error(msg);                                                           \
      return SEMK_ERR_INVALID;                                                       \
    }                                                                                \
  } while (0)

// Expand `CALL(N)` for the run-time value n1 in [2, 17].
#define SEMK_DISPATCH_N1(n1, CALL)                                                   \
  switch (n1) {                                                                      \
    case 2: CALL(2); break;                                                          \
    case 3: CALL(3); break;                                                          \
    case 4: CALL(4); break;                                                          \
    case 5: CALL(5); break;                                                          \
    case 6: CALL(6); break;                                                          \
    case 7: CALL(7); break;                                                          \
    case 8: CALL(8); break;                                                          \
    case 9: CALL(9); break;                                                          \
    case 10: CALL(10); break;                                                        \
    case 11: CALL(11); break;                                                        \
    case 12: CALL(12); break;                                                        \
    case 13: CALL(13); break;                                                        \
    case 14: CALL(14); break;                                                        \
    case 15: CALL(15); break;                                                        \
    case 16: CALL(16); break;                                                        \
    case 17: CALL(17); break;                                                        \
    default:                                                                         \
      semk_set_error("n1 outside [2, 17]");                                          \
      return SEMK_ERR_UNSUPPORTED;                                                   \
  }

static inline cudaStream_t semk_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

#ifdef __CUDACC__
// ---- async-proxy (TMA bulk copy) + mbarrier PTX wrappers (sm_90+/sm_100a) ----
__device__ __forceinline__ uint32_t semk_smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void semk_mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(semk_smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void semk_fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void semk_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(semk_smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool semk_mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(semk_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void semk_mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!semk_mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP);
// completion is signalled on `bar` as `bytes` transaction bytes.
__device__ __forceinline__ void semk_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                              uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(semk_smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(semk_smem_u32(bar))
      : "memory");
}
// The same with an L2 eviction-priority hint (data that is streamed exactly once).
__device__ __forceinline__ uint64_t semk_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void semk_bulk_g2s_hint(void *smem_dst, const void *gmem_src,
                                                   uint32_t bytes, uint64_t *bar,
                                                   uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1], %2, [%3], %4;" ::"r"(semk_smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(semk_smem_u32(bar)), "l"(policy)
      : "memory");
}

// L2 prefetch of one 128-byte line (SASS: CCTL.E.PF2) / of a contiguous block
// through the TMA engine (SASS: UBLKPF.L2; 16-byte aligned, size % 16 == 0).
__device__ __forceinline__ void semk_prefetch_l2(const void *p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void semk_bulk_prefetch_l2(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Deterministic block reduction of one double (blockDim.x a multiple of 32, <= 1024).
// Result valid in thread 0.
__device__ __forceinline__ double semk_block_sum(double v, double *smem_scratch /*[32]*/) {
  const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) smem_scratch[wid] = v;
  __syncthreads();
  const unsigned nw = (blockDim.x + 31u) >> 5;
  double t = 0.0;
  if (wid == 0) {
    t = (lane < nw) ? smem_scratch[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
  }
  __syncthreads();
  return t;
}
// ---- deterministic grid-wide reduction shared by the vector kernels (semk_vec.cu) and
// the condensed operator (semk_sc.cu).  partials layout: [kSemkRedMaxBlocks][4] doubles,
// then one 64-bit arrival counter (zeroed once by the caller; every reduction leaves it
// at zero).  Grids that use it have at most kSemkRedMaxBlocks CTAs.
constexpr int kSemkRedMaxBlocks = 148 * 8;
__device__ __forceinline__ unsigned long long *semk_red_counter(double *partials) {
  return reinterpret_cast<unsigned long long *>(partials + 4 * kSemkRedMaxBlocks);
}
// Publish this CTA's NV partial sums; returns true (to all threads) in the CTA
// that arrives last, with `tot[0..NV)` holding the fixed-order totals in thread 0.
template <int NV>
__device__ __forceinline__ bool semk_finish_reduction(double (&v)[NV], double *partials,
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
    const unsigned long long t = atomicAdd(semk_red_counter(partials), 1ull);
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
  if (threadIdx.x == 0) *semk_red_counter(partials) = 0ull;
  return true;
}
#endif
