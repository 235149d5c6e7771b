// Data-parallel training step: hot-replica fold + cross-GPU gradient all-reduce + TF1 optimizer + loss
// reduction as ONE persistent cooperative kernel, `dp_step_kernel`.
//
// Every rank owns an exchange buffer X in symmetric memory (same size on every rank, mapped into every peer, and -- on an
// NVSwitch box -- bound to one multicast object; the allocation and the address exchange are torch.distributed plumbing,
// hhfm_b200/dist.py).  One step, after the rank's scatter kernels have filled its private gradient arena:
//
//   A  fold the hot-row replicas into the arena (fixed replica order), move the arena into X and clear it for the next
//      step, reduce the rank's loss partials into the loss slot of X
//   -- cross-GPU barrier (system-scope release / acquire flags in symmetric memory)
//   B  rank r owns slice r of X: sum element i over all ranks -- `multimem.ld_reduce` (the NVSwitch adds the ranks' copies
//      in flight) or, without a multicast mapping, peer loads in rank order -- and write the sum back into EVERY rank's X
//      (`multimem.st` / peer stores).  All replicas therefore consume the same bits and stay bit-identical.
//   -- cross-GPU barrier
//   C  g = X (local), g_eff = g + lamda*w, optimizer update of the local replica (same expressions as opt.cu), sum of w^2
//      for the regulariser, loss_out = all-rank loss + sum_seg 0.5*lamda_seg*|w_seg|^2.
//
// With n_ranks == 1 the barriers and phase B drop out and the kernel is the single-GPU tail of a step (fold + optimizer
// + loss in one launch instead of three).  The barrier spins are bounded: on a timeout the kernel raises a sticky error
// flag in device memory (read back by the host together with the loss) and every later step returns immediately -- the
// CUDA context stays usable (no __trap).
#include <cooperative_groups.h>
#include <string.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace hhfm {

constexpr int kMaxPeers = 16;
constexpr int kMaxSegs = 4;
constexpr int kDpThreads = 512;

struct DpSeg {
  float* w;
  float* s1;
  float* s2;
  int64_t off;   // gradient of w[0, n) = X[off, off + n)
  int64_t n;
  float lamda;
};

struct OptP2 {
  float lr, lamda, b1, b2, eps;   // adagrad: lr; adam: lr_t, b1, b2, eps; momentum: lr, b1 = momentum; sgd: lr
};

struct DpArgs {
  DpSeg seg[kMaxSegs];
  int n_seg;
  float* arena;          // rank-private gradient arena, n_g floats used (16-byte aligned, readable up to roundup4(n_g))
  int64_t n_g;
  float* ghot;           // hot-row replicas [n_rep, n_hot, K] (+ bias [n_rep, n_hot]) or NULL
  float* ghot_bias;
  int n_rep, n_hot, K;
  const int32_t* hot_rows;
  int64_t bias_off;      // offset of the feature_bias gradient inside the arena (used with ghot_bias)
  const float* loss_partials;
  float* x;              // this rank's exchange buffer: roundup4(n_g) gradient floats + one float4 whose .x is the loss
  float* x_mc;           // multicast alias of X or NULL
  float* x_peer[kMaxPeers];
  int32_t* flag_peer[kMaxPeers];   // flag arrays (>= kMaxPeers int32 each) of every rank, own one included
  int rank, n_ranks;
  int32_t* state;        // device: [0] step counter, [1] sticky error
  OptP2 p;
  float* sq_ws;          // [gridDim.x] per-CTA regulariser partials
  float* loss_out;
  long long timeout_cycles;
};

// same expressions as opt.cu::opt_elem, so the fused step is bit-identical to the separate optimizer kernel
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

__device__ __forceinline__ float4 ld_cv4(const float4* p) { return __ldcv(p); }   // never a stale L1 line

__device__ __forceinline__ float4 multimem_ld_reduce4(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}

__device__ __forceinline__ void multimem_st4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Executed by one warp of CTA 0 after a grid-wide sync: lane r publishes `value` into rank r's flag array and waits for
// rank r's value in the local array.  Returns false on a timeout.
__device__ __forceinline__ bool cross_rank_barrier(const DpArgs& a, int32_t value) {
  const int r = threadIdx.x;
  bool ok = true;
  if (r < a.n_ranks) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(a.flag_peer[r] + a.rank), "r"(value) : "memory");
    const int32_t* mine = a.flag_peer[a.rank] + r;
    const long long t0 = clock64();
    for (;;) {
      int32_t v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if (v - value >= 0) break;
      if (clock64() - t0 > a.timeout_cycles) {
        ok = false;
        break;
      }
      __nanosleep(40);
    }
    __threadfence_system();
  }
  return ok;
}

template <int KIND>
__global__ void __launch_bounds__(kDpThreads) dp_step_kernel(const DpArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ float scratch[32];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  const int32_t step = __ldcv(a.state);
  if (__ldcv(a.state + 1) != 0) return;      // sticky error of an earlier step: uniform over the grid (written after the first sync only)

  // ---- A1: hot-row replicas -> arena (replica order fixed: reproducible), replicas cleared ----
  if (a.ghot != nullptr && a.n_hot > 0) {
    const int kv = a.K >> 2;
    const int64_t total = (int64_t)a.n_hot * kv;
    for (int64_t i = tid; i < total; i += nth) {
      const int s = (int)(i / kv), c = (int)(i % kv);
      float4 acc = f4_zero();
      for (int r0 = 0; r0 < a.n_rep; r0 += 8) {
        float4 t[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
          float4* p = reinterpret_cast<float4*>(a.ghot + ((size_t)(r0 + q) * a.n_hot + s) * a.K) + c;
          t[q] = (r0 + q < a.n_rep) ? *p : f4_zero();
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
          if (r0 + q < a.n_rep) {
            acc = f4_add(acc, t[q]);
            *(reinterpret_cast<float4*>(a.ghot + ((size_t)(r0 + q) * a.n_hot + s) * a.K) + c) = f4_zero();
          }
        }
      }
      float4* d = reinterpret_cast<float4*>(a.arena + (size_t)a.hot_rows[s] * a.K) + c;
      *d = f4_add(*d, acc);
    }
    if (a.ghot_bias != nullptr) {
      for (int64_t s = tid; s < a.n_hot; s += nth) {
        float acc = 0.f;
        for (int r = 0; r < a.n_rep; r++) {
          float* p = a.ghot_bias + (size_t)r * a.n_hot + s;
          acc += *p;
          *p = 0.f;
        }
        a.arena[a.bias_off + a.hot_rows[s]] += acc;
      }
    }
    grid.sync();
  }

  // ---- A2: arena -> X (arena cleared), loss partials -> loss slot ----
  const int64_t n_g4 = (a.n_g + 3) >> 2;
  {
    float4* src = reinterpret_cast<float4*>(a.arena);
    float4* dst = reinterpret_cast<float4*>(a.x);
    for (int64_t i = tid; i < n_g4; i += nth) {
      dst[i] = src[i];
      src[i] = f4_zero();
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
      float l = 0.f;
      for (int i = threadIdx.x; i < kPartials; i += 32) l += a.loss_partials[i];
      l = warp_sum(l);
      if (threadIdx.x == 0) dst[n_g4] = make_float4(l, 0.f, 0.f, 0.f);
    }
  }
  const int64_t total4 = n_g4 + 1;

  if (a.n_ranks > 1) {
    __threadfence_system();
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x < 32) {
      if (!cross_rank_barrier(a, 2 * step + 1)) a.state[1] = 1;
    }
    grid.sync();

    // ---- B: all-reduce of slice `rank`, result written into every rank's X ----
    const int64_t lo = total4 * a.rank / a.n_ranks, hi = total4 * (a.rank + 1) / a.n_ranks;
    if (a.x_mc != nullptr) {
      for (int64_t i = lo + tid; i < hi; i += nth) {
        const float4 v = multimem_ld_reduce4(a.x_mc + 4 * i);
        multimem_st4(a.x_mc + 4 * i, v);
      }
    } else {
      for (int64_t i = lo + tid; i < hi; i += nth) {
        float4 v = f4_zero();
        for (int r0 = 0; r0 < a.n_ranks; r0 += 4) {      // four remote loads in flight, summed in rank order
          float4 t[4];
#pragma unroll
          for (int q = 0; q < 4; q++)
            t[q] = (r0 + q < a.n_ranks) ? ld_cv4(reinterpret_cast<const float4*>(a.x_peer[r0 + q]) + i) : f4_zero();
#pragma unroll
          for (int q = 0; q < 4; q++)
            if (r0 + q < a.n_ranks) v = f4_add(v, t[q]);
        }
        for (int r = 0; r < a.n_ranks; r++) reinterpret_cast<float4*>(a.x_peer[r])[i] = v;
      }
    }
    __threadfence_system();
    grid.sync();
    if (blockIdx.x == 0 && threadIdx.x < 32) {
      if (!cross_rank_barrier(a, 2 * step + 2)) a.state[1] = 1;
    }
    grid.sync();
  } else {
    grid.sync();
  }

  // ---- C: optimizer over the segments, regulariser partials ----
  float reg = 0.f;
  for (int sgi = 0; sgi < a.n_seg; sgi++) {
    const DpSeg sg = a.seg[sgi];
    OptP2 p = a.p;
    p.lamda = sg.lamda;
    float sq = 0.f;
    const bool vec = ((sg.off & 3) == 0);
    const int64_t n4 = vec ? (sg.n >> 2) : 0;
    float4* w4 = reinterpret_cast<float4*>(sg.w);
    float4* a4 = reinterpret_cast<float4*>(sg.s1);
    float4* b4 = reinterpret_cast<float4*>(sg.s2);
    const float4* g4 = reinterpret_cast<const float4*>(a.x + sg.off);
    for (int64_t i = tid; i < n4; i += nth) {
      float4 gv = ld_cv4(g4 + i);
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
    }
    for (int64_t i = (n4 << 2) + tid; i < sg.n; i += nth) {
      float gv = __ldcv(a.x + sg.off + i);
      float wv = sg.w[i];
      float av = (KIND != HHFM_OPT_SGD) ? sg.s1[i] : 0.f;
      float bv = (KIND == HHFM_OPT_ADAM) ? sg.s2[i] : 0.f;
      sq += wv * wv;
      gv = fmaf(wv, p.lamda, gv);
      opt_elem2<KIND>(wv, av, bv, gv, p);
      sg.w[i] = wv;
      if (KIND != HHFM_OPT_SGD) sg.s1[i] = av;
      if (KIND == HHFM_OPT_ADAM) sg.s2[i] = bv;
    }
    if (sg.lamda > 0.f) reg += 0.5f * sg.lamda * sq;
  }
  {
    const float b = block_sum(reg, scratch);
    if (threadIdx.x == 0) a.sq_ws[blockIdx.x] = b;
  }
  grid.sync();
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    float r = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) r += __ldcv(a.sq_ws + i);
    r = warp_sum(r);
    if (threadIdx.x == 0) {
      a.loss_out[0] = __ldcv(a.x + 4 * n_g4) + r;
      a.state[0] = step + 1;
    }
  }
}

template <int KIND>
static int launch_dp(const DpArgs& a, cudaStream_t st) {
  static int grid_cached = 0;
  if (grid_cached == 0) {
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_step_kernel<KIND>, kDpThreads, 0);
    if (e != cudaSuccess || per_sm < 1) {
      set_error("dp_step: occupancy query failed: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return HHFM_ERR_LAUNCH;
    }
    grid_cached = sm_count();      // one CTA per SM: every CTA is resident, grid.sync() is legal
  }
  void* params[] = {(void*)&a};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)dp_step_kernel<KIND>, dim3(grid_cached), dim3(kDpThreads), params, 0, st);
  if (e != cudaSuccess) {
    set_error("dp_step_kernel: cooperative launch failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return HHFM_ERR_LAUNCH;
  }
  return check_launch("dp_step_kernel");
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int64_t hhfm_dp_exchange_floats(int64_t n_grad) {
  if (n_grad < 1) return -1;
  return ((n_grad + 3) / 4 + 1) * 4;
}

extern "C" int hhfm_dp_step(int32_t kind, const hhfm_dp_segment* segs, int32_t n_segs, float* arena, int64_t n_grad,
                            float* ghot, float* ghot_bias, int32_t n_rep, int32_t n_hot, int64_t K, const int32_t* hot_rows,
                            int64_t bias_off, const float* loss_partials, float* x_local, float* x_multicast,
                            const int64_t* x_peers_host, const int64_t* flag_peers_host, int32_t rank, int32_t n_ranks,
                            int32_t* state, float lr, float beta1, float beta2, float eps, float* reg_workspace,
                            float* loss_out, double timeout_s, hhfm_stream_t stream) {
  HHFM_REQUIRE(segs && n_segs >= 1 && n_segs <= kMaxSegs, "dp_step: 1..4 segments");
  HHFM_REQUIRE(arena && n_grad > 0 && loss_partials && x_local && state && reg_workspace && loss_out, "dp_step: NULL argument");
  HHFM_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxPeers && rank >= 0 && rank < n_ranks, "dp_step: bad rank / n_ranks");
  HHFM_REQUIRE(n_ranks == 1 || (x_peers_host && flag_peers_host), "dp_step: peer tables required for n_ranks > 1");
  HHFM_REQUIRE((((uintptr_t)arena | (uintptr_t)x_local | (uintptr_t)x_multicast) & 15) == 0, "dp_step: buffers must be 16-byte aligned");
  HHFM_REQUIRE(ghot == nullptr || (n_rep >= 1 && n_hot >= 1 && K > 0 && K % 4 == 0 && hot_rows), "dp_step: bad hot-row plan");
  DpArgs a;
  memset(&a, 0, sizeof(a));
  a.n_seg = n_segs;
  for (int i = 0; i < n_segs; i++) {
    const hhfm_dp_segment& s = segs[i];
    HHFM_REQUIRE(s.w && s.n > 0 && s.offset >= 0 && s.offset + s.n <= n_grad, "dp_step: segment outside the gradient range");
    HHFM_REQUIRE(kind == HHFM_OPT_SGD || s.s1, "dp_step: optimizer state is NULL");
    HHFM_REQUIRE(kind != HHFM_OPT_ADAM || s.s2, "dp_step: adam needs two state buffers");
    HHFM_REQUIRE((s.offset & 3) != 0 || s.n < 4 || (((uintptr_t)s.w | (uintptr_t)s.s1 | (uintptr_t)s.s2) & 15) == 0,
                 "dp_step: segment buffers must be 16-byte aligned");
    a.seg[i] = DpSeg{s.w, s.s1, s.s2, s.offset, s.n, s.lamda};
  }
  a.arena = arena;
  a.n_g = n_grad;
  a.ghot = (ghot && n_hot > 0) ? ghot : nullptr;
  a.ghot_bias = a.ghot ? ghot_bias : nullptr;
  a.n_rep = n_rep;
  a.n_hot = n_hot;
  a.K = (int)K;
  a.hot_rows = hot_rows;
  a.bias_off = bias_off;
  HHFM_REQUIRE(a.ghot_bias == nullptr || bias_off >= 0, "dp_step: bias_off required with ghot_bias");
  a.loss_partials = loss_partials;
  a.x = x_local;
  a.x_mc = (n_ranks > 1) ? x_multicast : nullptr;
  for (int r = 0; r < n_ranks && n_ranks > 1; r++) {
    a.x_peer[r] = reinterpret_cast<float*>(x_peers_host[r]);
    a.flag_peer[r] = reinterpret_cast<int32_t*>(flag_peers_host[r]);
    HHFM_REQUIRE(a.x_peer[r] && a.flag_peer[r] && ((uintptr_t)a.x_peer[r] & 15) == 0, "dp_step: bad peer address");
  }
  a.rank = rank;
  a.n_ranks = n_ranks;
  a.state = state;
  a.p = OptP2{lr, 0.f, beta1, beta2, eps};
  a.sq_ws = reg_workspace;
  a.loss_out = loss_out;
  if (timeout_s <= 0) timeout_s = 120.0;
  a.timeout_cycles = (long long)(timeout_s * 1.9e9);
  cudaStream_t st = (cudaStream_t)stream;
  switch (kind) {
    case HHFM_OPT_ADAGRAD: return launch_dp<HHFM_OPT_ADAGRAD>(a, st);
    case HHFM_OPT_ADAM: return launch_dp<HHFM_OPT_ADAM>(a, st);
    case HHFM_OPT_MOMENTUM: return launch_dp<HHFM_OPT_MOMENTUM>(a, st);
    case HHFM_OPT_SGD: return launch_dp<HHFM_OPT_SGD>(a, st);
    default: set_error("dp_step: unknown optimizer kind %d", kind); return HHFM_ERR_BAD_ARG;
  }
}
