// Wide&Deep (WDMF.py:51-126): the WIDE half of tf.contrib.learn.DNNLinearCombinedClassifier and its FTRL optimizer.
//
// The reference builds, per input column, a hashed sparse column (10^5 buckets) and, per pair of columns, a crossed column
// (10^4 buckets); the linear model is  wide(x) = b + sum_f w_lin[bucket_f(x_f)] + sum_{i<j} w_cross[(i,j)][bucket_ij(x_i, x_j)].
// TensorFlow's bucket functions are fingerprints of the id STRINGS (not restatable from the reference tree), so
// this restatement documents its own:
//   * single columns: one table keyed by the loader's global feature id (collision-free; a token shared by two columns
//     shares its weight, where TF's per-column tables would not);
//   * crosses: splitmix64((x_i << 32) | x_j) mod n_cross_buckets  (oracle: wd_cross_bucket, bit-exact).
// The deep half is the DeepFM tower without the FM terms (dfm.cu: hhfm_wd_deep_*).
#include <algorithm>

#include "common.cuh"

namespace hhfm {

__host__ __device__ inline uint64_t wd_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ int wd_cross_bucket(int xi, int xj, int n_buckets) {
  return (int)(wd_splitmix64(((uint64_t)(uint32_t)xi << 32) | (uint64_t)(uint32_t)xj) % (uint64_t)n_buckets);
}

constexpr int kWdMaxF = 32;

// one thread per sample; the tables (M + P * n_buckets floats) live in L2
__global__ void __launch_bounds__(256) wd_wide_fwd_kernel(const int32_t* __restrict__ idx, int64_t B, int F, const float* __restrict__ wl,
                                                          const float* __restrict__ wc, const float* __restrict__ bw, int nb,
                                                          float* __restrict__ out) {
  const float b = __ldg(bw);
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < B; s += (int64_t)gridDim.x * blockDim.x) {
    int x[kWdMaxF];
    float acc = 0.f;
#pragma unroll 4
    for (int f = 0; f < F; f++) { x[f] = __ldg(idx + s * F + f); acc += __ldg(wl + x[f]); }
    int p = 0;
    for (int i = 0; i < F; i++)
      for (int j = i + 1; j < F; j++, p++) acc += __ldg(wc + (size_t)p * nb + wd_cross_bucket(x[i], x[j], nb));
    out[s] = acc + b;
  }
}

__global__ void __launch_bounds__(256) wd_wide_bwd_kernel(const int32_t* __restrict__ idx, int64_t B, int F, const float* __restrict__ g,
                                                          int nb, float* __restrict__ gwl, float* __restrict__ gwc,
                                                          float* __restrict__ gbw) {
  __shared__ float scratch[32];
  float gb = 0.f;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < B; s += (int64_t)gridDim.x * blockDim.x) {
    int x[kWdMaxF];
    const float gs = __ldg(g + s);
    gb += gs;
#pragma unroll 4
    for (int f = 0; f < F; f++) { x[f] = __ldg(idx + s * F + f); atomicAdd(gwl + x[f], gs); }
    int p = 0;
    for (int i = 0; i < F; i++)
      for (int j = i + 1; j < F; j++, p++) atomicAdd(gwc + (size_t)p * nb + wd_cross_bucket(x[i], x[j], nb), gs);
  }
  gb = block_sum(gb, scratch);
  if (threadIdx.x == 0 && gb != 0.f) atomicAdd(gbw, gb);
}

// TF1 ApplyFtrl with learning_rate_power = -0.5 (tf.train.FtrlOptimizer defaults; the linear half of
// DNNLinearCombinedClassifier):  n' = n + g^2;  z += g - (sqrt(n') - sqrt(n)) / lr * w;
// w = |z| > l1 ? (sign(z) l1 - z) / (sqrt(n') / lr + 2 l2) : 0;  n = n'.
__global__ void __launch_bounds__(256) ftrl_dense_kernel(float* __restrict__ w, float* __restrict__ accum, float* __restrict__ linear,
                                                         float* __restrict__ g, int64_t n, float lr, float l1, float l2, int zero_grad) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    if (zero_grad) g[i] = 0.f;
    const float a0 = accum[i], a1 = a0 + gi * gi;
    const float r0 = sqrtf(a0), r1 = sqrtf(a1);
    const float z = linear[i] + gi - (r1 - r0) / lr * w[i];
    const float quad = r1 / lr + 2.f * l2;
    const float sgn = z > 0.f ? 1.f : (z < 0.f ? -1.f : 0.f);
    w[i] = fabsf(z) > l1 ? (sgn * l1 - z) / quad : 0.f;
    accum[i] = a1;
    linear[i] = z;
  }
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int hhfm_wd_wide_fwd(const int32_t* idx, int64_t B, int64_t F, const float* w_lin, const float* w_cross,
                                const float* b_wide, int64_t M, int32_t n_cross_buckets, float* out, hhfm_stream_t stream) {
  HHFM_REQUIRE(idx && w_lin && w_cross && b_wide && out, "wd_wide_fwd: NULL argument");
  HHFM_REQUIRE(B >= 0 && F >= 1 && F <= kWdMaxF && M > 0 && n_cross_buckets >= 1, "wd_wide_fwd: bad sizes");
  if (B == 0) return HHFM_OK;
  const int grid = (int)std::min<int64_t>((B + 255) / 256, (int64_t)sm_count() * 8);
  wd_wide_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx, B, (int)F, w_lin, w_cross, b_wide, n_cross_buckets, out);
  return check_launch("wd_wide_fwd_kernel");
}

extern "C" int hhfm_wd_wide_bwd(const int32_t* idx, int64_t B, int64_t F, const float* gsample, int64_t M,
                                int32_t n_cross_buckets, float* g_lin, float* g_cross, float* g_b, hhfm_stream_t stream) {
  HHFM_REQUIRE(idx && gsample && g_lin && g_cross && g_b, "wd_wide_bwd: NULL argument");
  HHFM_REQUIRE(B >= 0 && F >= 1 && F <= kWdMaxF && M > 0 && n_cross_buckets >= 1, "wd_wide_bwd: bad sizes");
  if (B == 0) return HHFM_OK;
  const int grid = (int)std::min<int64_t>((B + 255) / 256, (int64_t)sm_count() * 8);
  wd_wide_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(idx, B, (int)F, gsample, n_cross_buckets, g_lin, g_cross, g_b);
  return check_launch("wd_wide_bwd_kernel");
}

extern "C" int hhfm_opt_ftrl_dense(float* w, float* accum, float* linear, float* g, int64_t n, float lr, float l1, float l2,
                                   int32_t zero_grad, hhfm_stream_t stream) {
  HHFM_REQUIRE(w && accum && linear && g && n >= 0 && lr > 0.f && l1 >= 0.f && l2 >= 0.f, "opt_ftrl_dense: bad argument");
  if (n == 0) return HHFM_OK;
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 8);
  ftrl_dense_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, accum, linear, g, n, lr, l1, l2, zero_grad);
  return check_launch("ftrl_dense_kernel");
}
