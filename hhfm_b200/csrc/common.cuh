// Shared device/host helpers for libhhfm_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/hhfm_sm100.h"

namespace hhfm {

constexpr int kPartials = 2048;   // slots of every loss/sq partial buffer (>= any grid we launch)
constexpr int kBlock = 256;       // threads per CTA of the gather/scatter kernels (8 warps)

// ---- error plumbing (thread-local, no exceptions across the ABI) ------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);
int sm_count();

#define HHFM_REQUIRE(cond, ...)                    \
  do {                                             \
    if (!(cond)) {                                 \
      hhfm::set_error(__VA_ARGS__);                \
      return HHFM_ERR_BAD_ARG;                     \
    }                                              \
  } while (0)

// ---- device primitives ------------------------------------------------------------------------------
// Vector reduction to global memory: one REDG.E.ADD.F32x4 per 16 bytes (sm_90+), no return value.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_scale(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 f4_fma(float4 a, float s, float4 c) {
  return make_float4(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w));
}
__device__ __forceinline__ float f4_hsum(float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float f4_dot(float4 a, float4 b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// Sum over an aligned power-of-two group of LPS lanes (xor butterfly; every lane gets the result).
// Must be called by all 32 lanes of the warp.
template <int LPS>
__device__ __forceinline__ float group_sum(float x) {
#pragma unroll
  for (int o = LPS / 2; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

__device__ __forceinline__ float warp_sum(float x) { return group_sum<32>(x); }

// CTA-wide sum; result valid in thread 0.  `scratch` holds >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float x, float* scratch) {
  x = warp_sum(x);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[w] = x;
  __syncthreads();
  float r = 0.f;
  if (w == 0) {
    r = (lane < (int)(blockDim.x + 31) / 32) ? scratch[lane] : 0.f;
    r = warp_sum(r);
  }
  return r;
}

// Each CTA owns partials[blockIdx.x]; CTA 0 also clears the unused tail so the finalize kernel can sum
// all kPartials slots in a fixed order.
__device__ __forceinline__ void write_partial(float* partials, float block_value) {
  if (partials == nullptr) return;
  if (threadIdx.x == 0) partials[blockIdx.x] = block_value;
  if (blockIdx.x == 0)
    for (int i = gridDim.x + threadIdx.x; i < kPartials; i += blockDim.x) partials[i] = 0.f;
}

// Touched-row tracking for the sparse (*_rows) optimizers: first toucher in this step appends the row.
__device__ __forceinline__ void touch_row(int32_t* stamp_arr, int32_t stamp, int32_t* list, int32_t* count, int row) {
  if (stamp_arr == nullptr) return;
  if (__ldcv(stamp_arr + row) != stamp) {
    int old = atomicExch(stamp_arr + row, stamp);
    if (old != stamp) list[atomicAdd(count, 1)] = row;
  }
}

// Counter-based generator shared by the device sampler (sampler.cu) and the dropout masks (fm.cu): every draw is a pure
// function of (seed, counter), restated bit-exactly by the oracle.
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// tf.nn.dropout(x, keep) = x / keep * floor(keep + u), u uniform in [0,1): element (sample s, column k) is kept iff
// u24(seed, s, k) * 2^-24 < keep.  Returns the 0/1 mask value.
__device__ __forceinline__ float dropout_keep01(uint64_t seed, int64_t s, int k, float keep) {
  const uint64_t h = splitmix64(seed ^ splitmix64((uint64_t)s * 0x100000001B3ull + (uint64_t)k));
  return ((float)(h >> 40) * (1.0f / 16777216.0f) < keep) ? 1.f : 0.f;
}

// Embedding row fragment held by one lane: VPL float4 chunks, chunk c = lg + i*LPS of the K/4 in a row.
template <int LPS, int VPL>
struct Frag {
  float4 v[VPL];
};

// A row of K floats viewed by the LPS lanes of a group.  K may be smaller than 4*LPS*VPL (e.g. K = 100): the
// surplus lanes read zeros and write nothing.  When K == 4*LPS*VPL the bound test folds away (kExact below).
template <int LPS, int VPL>
__device__ __forceinline__ void frag_load(Frag<LPS, VPL>& r, const float* __restrict__ V, int row, int K, int lg) {
  const float4* p = reinterpret_cast<const float4*>(V) + (size_t)row * (K >> 2);
  const int kv = K >> 2;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int c = lg + i * LPS;
    r.v[i] = (c < kv) ? ldg4(p + c) : f4_zero();
  }
}

template <int LPS, int VPL>
__device__ __forceinline__ void frag_red_ptr(float* __restrict__ p, int K, int lg, const Frag<LPS, VPL>& r) {
  const int kv = K >> 2;
#pragma unroll
  for (int i = 0; i < VPL; i++) {
    const int c = lg + i * LPS;
    if (c < kv) red_add_v4(p + 4 * c, r.v[i]);
  }
}

template <int LPS, int VPL>
__device__ __forceinline__ void frag_red(float* __restrict__ G, int row, int K, int lg, const Frag<LPS, VPL>& r) {
  frag_red_ptr<LPS, VPL>(G + (size_t)row * K, K, lg, r);
}

// Two-level sort-free scatter target.  Rows flagged hot (hot_slot[row] >= 0) accumulate into one of n_rep
// replicas of a small [n_rep, n_hot, K] buffer chosen per lane-group, so the reductions of a row that thousands
// of samples share spread over n_rep addresses / L2 slices instead of serialising on one; hhfm_hot_fold sums the
// replicas back into the [M, K] gradient before the optimizer.  Cold rows go straight to the gradient table.
struct HotPlan {
  const int32_t* slot;   // [M] slot id or -1, NULL = no hot rows
  float* ghot;           // [n_rep, n_hot, K]
  float* ghot_bias;      // [n_rep, n_hot] (FM feature_bias gradient) or NULL
  int n_rep, n_hot;
};

template <int LPS, int VPL>
__device__ __forceinline__ void scatter_row(float* __restrict__ G, const HotPlan& hp, int rep, int row, int K, int lg,
                                            const Frag<LPS, VPL>& r) {
  float* p = G + (size_t)row * K;
  if (hp.slot != nullptr) {
    const int s = __ldg(hp.slot + row);
    if (s >= 0) p = hp.ghot + ((size_t)rep * hp.n_hot + s) * K;
  }
  frag_red_ptr<LPS, VPL>(p, K, lg, r);
}

__device__ __forceinline__ void scatter_bias(float* __restrict__ gb, const HotPlan& hp, int rep, int row, float v) {
  float* p = gb + row;
  if (hp.slot != nullptr && hp.ghot_bias != nullptr) {
    const int s = __ldg(hp.slot + row);
    if (s >= 0) p = hp.ghot_bias + (size_t)rep * hp.n_hot + s;
  }
  atomicAdd(p, v);
}

template <int LPS, int VPL>
__device__ __forceinline__ void frag_zero(Frag<LPS, VPL>& r) {
#pragma unroll
  for (int i = 0; i < VPL; i++) r.v[i] = f4_zero();
}

template <int LPS, int VPL>
__device__ __forceinline__ float frag_dot(const Frag<LPS, VPL>& a, const Frag<LPS, VPL>& b) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; i++) s += f4_dot(a.v[i], b.v[i]);
  return s;
}

// Dispatch on the factor size: K/4 float4 per row spread over LPS lanes x VPL chunks.
#define HHFM_DISPATCH_K(K, CALL)                                   \
  do {                                                             \
    const int kv__ = (int)((K) >> 2);                              \
    if (kv__ <= 2) { CALL(2, 1); }                                 \
    else if (kv__ <= 4) { CALL(4, 1); }                            \
    else if (kv__ <= 8) { CALL(8, 1); }                            \
    else if (kv__ <= 16) { CALL(16, 1); }                          \
    else if (kv__ <= 32) { CALL(32, 1); }                          \
    else if (kv__ <= 64) { CALL(32, 2); }                          \
    else { CALL(32, 4); }                                          \
  } while (0)

inline int grid_for(int64_t groups_needed, int groups_per_block, int blocks_per_sm) {
  int64_t need = (groups_needed + groups_per_block - 1) / groups_per_block;
  int64_t cap = (int64_t)sm_count() * blocks_per_sm;
  if (cap > kPartials) cap = kPartials;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

}  // namespace hhfm
