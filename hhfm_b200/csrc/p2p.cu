// Data-parallel exchange over NVLink peer memory, fused into the optimizer.
//
// Every rank exports a copy of its gradient arena in a cudaMalloc'd buffer that the other ranks of the box map through
// CUDA IPC (double buffered: a buffer is only rewritten after the barrier that proves every peer has finished reading it).
// A step is: local scatter kernels -> copy into the export buffer -> hhfm_p2p_barrier (flags written into every peer with
// system-scope release stores, each rank spins on its own copy) -> hhfm_opt_*_dense_l2_p2p, which reads element i of the
// gradient from EVERY rank's export buffer (fixed rank order, so all replicas compute bit-identical sums and stay
// bit-identical), applies the TF1 optimizer update to the local replica and clears the local arena slice.
// There is no separate all-reduce pass: the reduction traffic (G x n x 4 bytes over NVLink per rank) is the
// optimizer's gradient read.  NCCL stays in use for broadcast / all-gather plumbing only.
#include <string.h>

#include "common.cuh"

namespace hhfm {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
  const float* g[kMaxPeers];
};

__device__ __forceinline__ float4 ld_peer4(const float4* p) { return __ldcv(p); }   // never a stale cached line
__device__ __forceinline__ float ld_peer(const float* p) { return __ldcv(p); }

struct OptP2 {
  float lr, lamda, b1, b2, eps;   // adagrad: lr, lamda; adam: lr_t, b1, b2, eps; momentum: lr, b1 = momentum; sgd: lr
};

// same expressions as opt.cu::opt_elem, so a 1-rank p2p step is bit-identical to the local kernel
template <int KIND>
__device__ __forceinline__ void opt_elem2(float& w, float& a, float& b, float g, const OptP2& p) {
  if (KIND == HHFM_OPT_ADAGRAD) {
    a = a + g * g;
    w = w - p.lr * g / sqrtf(a);
  } else if (KIND == HHFM_OPT_ADAM) {
    a = p.b1 * a + (1.f - p.b1) * g;
    b = p.b2 * b + (1.f - p.b2) * (g * g);
    w = w - p.lr * a / (sqrtf(b) + p.eps);
  } else if (KIND == HHFM_OPT_MOMENTUM) {
    a = a * p.b1 + g;
    w = w - p.lr * a;
  } else {
    w = w - p.lr * g;
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) opt_dense_p2p_kernel(float* __restrict__ w, float* __restrict__ s1, float* __restrict__ s2,
                                                            const PeerPtrs peers, int n_ranks, float* __restrict__ g_zero,
                                                            int64_t n, OptP2 p, float* sq_partials) {
  __shared__ float scratch[32];
  const int64_t n4 = n >> 2;
  float sq = 0.f;
  float4* w4 = reinterpret_cast<float4*>(w);
  float4* a4 = reinterpret_cast<float4*>(s1);
  float4* b4 = reinterpret_cast<float4*>(s2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    // the remote loads of four ranks are in flight together (one NVLink round trip per group); the sum is still formed in
    // rank order, so every replica computes the same bits
    float4 gv = f4_zero();
    for (int r0 = 0; r0 < n_ranks; r0 += 4) {
      float4 t[4];
#pragma unroll
      for (int q = 0; q < 4; q++)
        t[q] = (r0 + q < n_ranks) ? ld_peer4(reinterpret_cast<const float4*>(peers.g[r0 + q]) + i) : f4_zero();
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (r0 + q < n_ranks) gv = f4_add(gv, t[q]);
    }
    float4 wv = w4[i];
    float4 av = (KIND != HHFM_OPT_SGD) ? a4[i] : f4_zero();
    float4 bv = (KIND == HHFM_OPT_ADAM) ? b4[i] : f4_zero();
    sq += f4_dot(wv, wv);
    gv = f4_fma(wv, p.lamda, gv);
    opt_elem2<KIND>(wv.x, av.x, bv.x, gv.x, p);
    opt_elem2<KIND>(wv.y, av.y, bv.y, gv.y, p);
    opt_elem2<KIND>(wv.z, av.z, bv.z, gv.z, p);
    opt_elem2<KIND>(wv.w, av.w, bv.w, gv.w, p);
    w4[i] = wv;
    if (KIND != HHFM_OPT_SGD) a4[i] = av;
    if (KIND == HHFM_OPT_ADAM) b4[i] = bv;
    if (g_zero) reinterpret_cast<float4*>(g_zero)[i] = f4_zero();
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float gv = ld_peer(peers.g[0] + i);
    for (int r = 1; r < n_ranks; r++) gv += ld_peer(peers.g[r] + i);
    float wv = w[i];
    float av = (KIND != HHFM_OPT_SGD) ? s1[i] : 0.f;
    float bv = (KIND == HHFM_OPT_ADAM) ? s2[i] : 0.f;
    sq += wv * wv;
    gv = fmaf(wv, p.lamda, gv);
    opt_elem2<KIND>(wv, av, bv, gv, p);
    w[i] = wv;
    if (KIND != HHFM_OPT_SGD) s1[i] = av;
    if (KIND == HHFM_OPT_ADAM) s2[i] = bv;
    if (g_zero) g_zero[i] = 0.f;
  }
  if (sq_partials != nullptr) {
    const float b = block_sum(sq, scratch);
    write_partial(sq_partials, b);
  }
}

struct FlagPtrs {
  int32_t* f[kMaxPeers];
};

// One CTA.  Thread r publishes this rank's epoch into rank r's flag array, then waits for rank r's epoch in the local
// array.  Bounded spin: a lost peer must surface as a launch error, not hang the GPU.
__global__ void p2p_barrier_kernel(const FlagPtrs flags, int rank, int n_ranks, int32_t epoch, int* err) {
  const int r = threadIdx.x;
  if (r >= n_ranks) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flags.f[r] + rank), "r"(epoch) : "memory");
  const int32_t* mine = flags.f[rank] + r;
  const long long t0 = clock64();
  for (;;) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (v - epoch >= 0) break;
    if (clock64() - t0 > 20000000000LL) {     // ~10 s
      if (err) *err = 1;
      __trap();
    }
  }
  __threadfence_system();
}

// n_per_rank values are read from every rank.  The callers reduce their own loss partials locally first and publish ONE
// float per rank: reading all 2048 partial slots of 8 ranks from a single warp was 512 dependent remote loads per lane
// (~0.3 ms per step at 8 GPUs).
__global__ void loss_finalize_p2p_kernel(const PeerPtrs lp, int n_ranks, int n_per_rank, const float* __restrict__ sp,
                                         float half_lamda, float* __restrict__ out) {
  float a = 0.f, b = 0.f;
  for (int r = 0; r < n_ranks; r++)
    for (int i = threadIdx.x; i < n_per_rank; i += 32) a += ld_peer(lp.g[r] + i);
  if (sp)
    for (int i = threadIdx.x; i < kPartials; i += 32) b += sp[i];
  a = warp_sum(a);
  b = warp_sum(b);
  if (threadIdx.x == 0) out[0] = a + half_lamda * b;
}

static int dense_grid2(int64_t n) {
  int64_t need = ((n >> 2) + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > kPartials) cap = kPartials;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int hhfm_p2p_alloc(int64_t bytes, void** dev_ptr, void* handle64) {
  HHFM_REQUIRE(bytes > 0 && dev_ptr && handle64, "p2p_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    set_error("p2p_alloc: %s", cudaGetErrorString(e));
    if (p) cudaFree(p);
    cudaGetLastError();
    return HHFM_ERR_LAUNCH;
  }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return HHFM_OK;
}

extern "C" int hhfm_p2p_open(const void* handle64, void** dev_ptr) {
  HHFM_REQUIRE(handle64 && dev_ptr, "p2p_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    set_error("p2p_open: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return HHFM_ERR_LAUNCH;
  }
  *dev_ptr = p;
  return HHFM_OK;
}

extern "C" int hhfm_p2p_close(void* dev_ptr) {
  if (dev_ptr && cudaIpcCloseMemHandle(dev_ptr) != cudaSuccess) cudaGetLastError();
  return HHFM_OK;
}

extern "C" int hhfm_p2p_free(void* dev_ptr) {
  if (dev_ptr && cudaFree(dev_ptr) != cudaSuccess) cudaGetLastError();
  return HHFM_OK;
}

extern "C" int hhfm_p2p_barrier(const int64_t* flag_ptrs_host, int32_t rank, int32_t n_ranks, int32_t epoch,
                                hhfm_stream_t stream) {
  HHFM_REQUIRE(flag_ptrs_host && n_ranks >= 1 && n_ranks <= kMaxPeers && rank >= 0 && rank < n_ranks, "p2p_barrier: bad arguments");
  FlagPtrs f{};
  for (int r = 0; r < n_ranks; r++) f.f[r] = reinterpret_cast<int32_t*>(flag_ptrs_host[r]);
  p2p_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, n_ranks, epoch, nullptr);
  return check_launch("p2p_barrier_kernel");
}

extern "C" int hhfm_opt_dense_l2_p2p(int32_t kind, float* w, float* s1, float* s2, const int64_t* grad_ptrs_host,
                                     int32_t n_ranks, float* g_zero, int64_t n, float lr, float lamda, float beta1,
                                     float beta2, float eps, float* sq_partials, hhfm_stream_t stream) {
  HHFM_REQUIRE(w && grad_ptrs_host && n > 0 && n_ranks >= 1 && n_ranks <= kMaxPeers, "opt_dense_l2_p2p: bad arguments");
  HHFM_REQUIRE(kind == HHFM_OPT_SGD || s1, "opt_dense_l2_p2p: optimizer state is NULL");
  HHFM_REQUIRE(kind != HHFM_OPT_ADAM || s2, "opt_dense_l2_p2p: adam needs two state buffers");
  PeerPtrs pp{};
  uintptr_t al = (uintptr_t)w | (uintptr_t)s1 | (uintptr_t)s2 | (uintptr_t)g_zero;
  for (int r = 0; r < n_ranks; r++) {
    pp.g[r] = reinterpret_cast<const float*>(grad_ptrs_host[r]);
    al |= (uintptr_t)grad_ptrs_host[r];
  }
  HHFM_REQUIRE((al & 15) == 0 || n < 4, "opt_dense_l2_p2p: buffers must be 16-byte aligned");
  OptP2 p{lr, lamda, beta1, beta2, eps};
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = dense_grid2(n);
  switch (kind) {
    case HHFM_OPT_ADAGRAD: opt_dense_p2p_kernel<HHFM_OPT_ADAGRAD><<<grid, 256, 0, st>>>(w, s1, s2, pp, n_ranks, g_zero, n, p, sq_partials); break;
    case HHFM_OPT_ADAM: opt_dense_p2p_kernel<HHFM_OPT_ADAM><<<grid, 256, 0, st>>>(w, s1, s2, pp, n_ranks, g_zero, n, p, sq_partials); break;
    case HHFM_OPT_MOMENTUM: opt_dense_p2p_kernel<HHFM_OPT_MOMENTUM><<<grid, 256, 0, st>>>(w, s1, s2, pp, n_ranks, g_zero, n, p, sq_partials); break;
    case HHFM_OPT_SGD: opt_dense_p2p_kernel<HHFM_OPT_SGD><<<grid, 256, 0, st>>>(w, s1, s2, pp, n_ranks, g_zero, n, p, sq_partials); break;
    default: set_error("opt_dense_l2_p2p: unknown optimizer kind %d", kind); return HHFM_ERR_BAD_ARG;
  }
  return check_launch("opt_dense_p2p_kernel");
}

extern "C" int hhfm_loss_finalize_p2p(const int64_t* partial_ptrs_host, int32_t n_ranks, int32_t n_per_rank,
                                      const float* sq_partials, float half_lamda, float* loss_out, hhfm_stream_t stream) {
  HHFM_REQUIRE(partial_ptrs_host && loss_out && n_ranks >= 1 && n_ranks <= kMaxPeers && n_per_rank >= 1 && n_per_rank <= kPartials,
               "loss_finalize_p2p: bad arguments");
  PeerPtrs pp{};
  for (int r = 0; r < n_ranks; r++) pp.g[r] = reinterpret_cast<const float*>(partial_ptrs_host[r]);
  loss_finalize_p2p_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp, n_ranks, n_per_rank, sq_partials, half_lamda, loss_out);
  return check_launch("loss_finalize_p2p_kernel");
}
