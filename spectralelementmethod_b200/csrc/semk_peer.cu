// semk_peer.cu -- interface exchange of the strip partition over peer memory.
//
// The reference has no multi-process path (its scatter-add `grhs[inds] += ...`,
// sem/discrete.py:499, is serial); SURVEY.md 8(e) defines the exchange: after the
// local apply every rank holds PARTIAL sums on the node columns it shares with
// its left / right neighbour, and the two partial sums must be added.
//
// Instead of NCCL send/recv + separate add / fix-up kernels, ONE kernel per
// apply does the whole step over NVLink peer memory:
//   push : store my boundary column straight into the neighbour's receive
//          buffer (peer pointer obtained through CUDA IPC), then publish the
//          exchange epoch in the neighbour's flag word (system-scope release);
//   wait : spin (system-scope acquire) on my own flag until the neighbour's
//          column of this epoch has landed;
//   add  : y += received column, re-impose the Dirichlet identity rows on the
//          shared column and take the doubly counted u^2 out of the fused u.y.
// Two CTAs (left side, right side); a push never waits, so two neighbouring
// ranks cannot deadlock, and the receive buffers are double-buffered by epoch
// parity (a neighbour can be at most one exchange ahead).  A spin that exceeds
// ~2 s gives up and raises a status word instead of hanging the GPU.
#include "semk_common.cuh"

#include <cstring>

namespace {

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int kHaloThreads = 1024;
thread_local bool g_light_exchange = false;   // set around the call by semk_poisson_apply_halo_f64
constexpr long long kSpinLimit = 4000000000LL;  // ~2 s of SM clock

struct HaloSide {
  double *y;                       // my boundary column inside y (local)
  const double *u;                 // same column of u (Dirichlet identity rows)
  const unsigned char *dirichlet;  // same column of the mask, or nullptr
  double *peer_recv;               // neighbour's receive buffer for this epoch parity (peer)
  unsigned long long *peer_flag;   // neighbour's flag word for this epoch parity (peer)
  const double *my_recv;           // my receive buffer for this epoch parity (local)
  const unsigned long long *my_flag;
  int active;                      // 0: no neighbour on this side
  int subtract_dup;                // 1: this rank does not own the column (take u^2 out of dot)
};

// LIGHT = the variant that runs on a side stream NEXT TO the persistent apply kernel: 128
// threads and <= 32 registers, so that it fits into what three resident apply CTAs leave
// free on an SM (4096 registers, ~5 KB of shared memory) and does not have to wait for one
// of them to retire.
constexpr int kHaloLightThreads = 128;

template <bool LIGHT>
__global__ void __launch_bounds__(LIGHT ? kHaloLightThreads : kHaloThreads, LIGHT ? 16 : 1)
    halo_exchange_kernel(HaloSide left, HaloSide right, int64_t n, unsigned long long epoch,
                         double *__restrict__ dot_inout, int *__restrict__ status) {
  const HaloSide &S = blockIdx.x == 0 ? left : right;
  if (!S.active) return;
  __shared__ double red[32];
  __shared__ int timed_out;
  // ---- push my partial sums into the neighbour's memory -------------------------
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) S.peer_recv[i] = S.y[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    st_release_sys(S.peer_flag, epoch);
    // ---- wait for the neighbour's column of the same epoch ------------------------
    timed_out = 0;
    const long long t0 = clock64();
    while (ld_acquire_sys(S.my_flag) < epoch) {
      if (clock64() - t0 > kSpinLimit) {
        timed_out = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (timed_out) {
    if (threadIdx.x == 0) atomicExch(status, 1);
    return;
  }
  // ---- add, re-impose identity rows, fix the fused dot ----------------------------
  double dup = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    double v = S.y[i] + __ldcg(S.my_recv + i);  // written by the peer: bypass L1
    if (S.dirichlet && S.dirichlet[i]) {
      v = S.u[i];
      if (S.subtract_dup) dup = fma(v, v, dup);
    }
    S.y[i] = v;
  }
  if (dot_inout && S.subtract_dup) {
    const double s = semk_block_sum(dup, red);
    if (threadIdx.x == 0) dot_inout[0] -= s;
  }
}

}  // namespace

extern "C" int semk_peer_alloc(int64_t bytes, void **dev_ptr, unsigned char *handle_out) {
  SEMK_REQUIRE(bytes > 0 && dev_ptr && handle_out, "semk_peer_alloc: bad arguments");
  void *p = nullptr;
  SEMK_CUDA_CHECK(cudaMalloc(&p, (size_t)bytes));
  cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    semk_set_error(std::string("semk_peer_alloc: ") + cudaGetErrorString(e));
    (void)cudaGetLastError();
    return SEMK_ERR_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == SEMK_PEER_HANDLE_BYTES, "IPC handle size");
  std::memcpy(handle_out, &h, sizeof(h));
  *dev_ptr = p;
  return SEMK_OK;
}

extern "C" int semk_peer_open(const unsigned char *handle, void **dev_ptr) {
  SEMK_REQUIRE(handle && dev_ptr, "semk_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void *p = nullptr;
  SEMK_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return SEMK_OK;
}

extern "C" int semk_peer_close(void *dev_ptr) {
  if (dev_ptr) SEMK_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
  return SEMK_OK;
}

extern "C" int semk_peer_free(void *dev_ptr) {
  if (dev_ptr) SEMK_CUDA_CHECK(cudaFree(dev_ptr));
  return SEMK_OK;
}

extern "C" int semk_halo_exchange_f64(int64_t n_col, int64_t n_local, double *y, const double *u,
                                      const uint8_t *dirichlet, void *my_region,
                                      void *left_region, void *right_region, uint64_t epoch,
                                      double *dot_inout, int *status, void *stream) {
  SEMK_REQUIRE(n_col > 0 && n_local >= n_col && y && my_region && status && epoch > 0,
               "semk_halo_exchange_f64: bad arguments");
  SEMK_REQUIRE(!dirichlet || u, "semk_halo_exchange_f64: the Dirichlet fix-up needs u");
  // region layout (see semk_halo_region_bytes): 4 flag words in the first 256 bytes
  //   flag[0..1]: written by the LEFT neighbour (parity 0, 1); flag[2..3]: by the RIGHT one
  // then recv_left[2][n_col], recv_right[2][n_col]
  const int par = (int)(epoch & 1u);
  auto flags = [](void *r) { return reinterpret_cast<unsigned long long *>(r); };
  auto bufs = [](void *r) { return reinterpret_cast<double *>(static_cast<char *>(r) + 256); };
  HaloSide L{}, R{};
  if (left_region) {
    L.active = 1;
    L.y = y;
    L.u = u;
    L.dirichlet = dirichlet;
    // I am the left neighbour's RIGHT side
    L.peer_recv = bufs(left_region) + (2 + par) * n_col;
    L.peer_flag = flags(left_region) + 2 + par;
    L.my_recv = bufs(my_region) + (0 + par) * n_col;
    L.my_flag = flags(my_region) + 0 + par;
    L.subtract_dup = 0;  // the right neighbour of a shared column owns it: that is me
  }
  if (right_region) {
    const int64_t off = n_local - n_col;
    R.active = 1;
    R.y = y + off;
    R.u = u ? u + off : nullptr;
    R.dirichlet = dirichlet ? dirichlet + off : nullptr;
    R.peer_recv = bufs(right_region) + (0 + par) * n_col;
    R.peer_flag = flags(right_region) + 0 + par;
    R.my_recv = bufs(my_region) + (2 + par) * n_col;
    R.my_flag = flags(my_region) + 2 + par;
    R.subtract_dup = 1;
  }
  if (!L.active && !R.active) return SEMK_OK;
  if (g_light_exchange)
    halo_exchange_kernel<true><<<2, kHaloLightThreads, 0, semk_stream(stream)>>>(
        L, R, n_col, (unsigned long long)epoch, dot_inout, status);
  else
    halo_exchange_kernel<false><<<2, kHaloThreads, 0, semk_stream(stream)>>>(
        L, R, n_col, (unsigned long long)epoch, dot_inout, status);
  SEMK_LAUNCH_CHECK("halo_exchange_kernel");
  return SEMK_OK;
}

// y = A u on a strip partition with the interface exchange OVERLAPPED with the bulk of the
// local apply.  The plan's patch sequence starts with the two boundary tile columns
// (patches [0, bnd_patch_end); their interface entries are the prefixes [0, bnd_chunk_end)
// / [0, bnd_rec_end) of the interface tables): they run on a high-priority side stream,
// followed there by the exchange kernel -- push, publish, wait for the neighbour, add --
// while the caller's stream works through the remaining patches, and the two streams join.
// A neighbour may now be late by almost a whole apply before anybody waits for it.
// halo->epoch is advanced.  No fused dot (use semk_poisson_apply_f64 +
// semk_halo_exchange_f64 when u.y is needed).
extern "C" int semk_poisson_apply_halo_f64(const semk_op *op, const double *u, double *y,
                                           int flags, int64_t bnd_patch_end,
                                           int64_t bnd_chunk_end, int64_t bnd_rec_end,
                                           semk_halo *halo, const uint8_t *dirichlet,
                                           void *stream) {
  SEMK_REQUIRE(op && halo && u && y, "semk_poisson_apply_halo_f64: null pointer");
  struct Side {
    cudaStream_t s = nullptr;
    cudaEvent_t a = nullptr, x = nullptr;
  };
  thread_local Side side;
  if (!side.s) {
    int lo = 0, hi = 0;
    SEMK_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    SEMK_CUDA_CHECK(cudaStreamCreateWithPriority(&side.s, cudaStreamNonBlocking, hi));
    SEMK_CUDA_CHECK(cudaEventCreateWithFlags(&side.a, cudaEventDisableTiming));
    SEMK_CUDA_CHECK(cudaEventCreateWithFlags(&side.x, cudaEventDisableTiming));
  }
  cudaStream_t st = semk_stream(stream);
  // the side stream starts after whatever the caller queued on `stream` (u ready, y free)
  SEMK_CUDA_CHECK(cudaEventRecord(side.a, st));
  SEMK_CUDA_CHECK(cudaStreamWaitEvent(side.s, side.a, 0));
  // boundary tile columns + the interface entries they complete, then the exchange, all on
  // the high-priority side stream; submitted FIRST so that their CTAs are placed before the
  // persistent kernel of the interior patches fills the remaining slots
  int rc = semk_poisson_apply_range_f64(op, u, y, flags, 0, bnd_patch_end, 0, bnd_chunk_end, 0,
                                        bnd_rec_end, side.s);
  if (rc != SEMK_OK) return rc;
  halo->epoch += 1;
  g_light_exchange = true;
  rc = semk_halo_exchange_f64(halo->n_col, op->n_nodes, y, u, dirichlet, halo->mine, halo->left,
                              halo->right, halo->epoch, nullptr, halo->status, side.s);
  g_light_exchange = false;
  if (rc != SEMK_OK) return rc;
  SEMK_CUDA_CHECK(cudaEventRecord(side.x, side.s));
  // interior patches on the caller's stream, concurrently (disjoint patches, slots and
  // result nodes)
  rc = semk_poisson_apply_range_f64(op, u, y, flags, bnd_patch_end, op->n_patch, bnd_chunk_end,
                                    op->n_shared_chunk, bnd_rec_end, op->n_shared, stream);
  if (rc != SEMK_OK) return rc;
  SEMK_CUDA_CHECK(cudaStreamWaitEvent(st, side.x, 0));
  return SEMK_OK;
}

extern "C" int64_t semk_halo_region_bytes(int64_t n_col) {
  return n_col > 0 ? 256 + 4 * n_col * (int64_t)sizeof(double) : -1;
}
