// negative sampling and evaluate_AUC on the device.
//
// Reference: Train.sample_negative (FM.py:284-294) draws `num` uniform items per row and re-draws while the item is in
// positive_feedback[key(row)]; Train.evaluate_AUC (FM.py:296-324) scores 50 such negatives per positive and reports the
// fraction with pos > neg.  The reference draws from numpy's global Mersenne stream inside Python loops; here every draw
// is a pure function of (seed, row, column, attempt) -- a counter-based splitmix64 hash -- so the sample is reproducible,
// order-independent and restated bit-exactly by the oracle (oracle/hhfm_oracle.py::sample_negative_hashed).  It is
// statistically the reference's sampler (uniform + rejection), not its random stream; the host sampler in trainer.py
// keeps the stream-compatible variant.
#include "common.cuh"

namespace hhfm {

// item = n_user + floor(u32 * n_item / 2^32), u32 = top half of the hash of (seed, cell, attempt)
__device__ __forceinline__ int draw_item(uint64_t seed, uint64_t cell, uint32_t attempt, int n_user, int n_item) {
  const uint64_t h = splitmix64(seed ^ splitmix64(cell * 0x100000001B3ull + attempt));
  return n_user + (int)(((h >> 32) * (uint64_t)n_item) >> 32);
}

__device__ __forceinline__ bool code_present(const int64_t* __restrict__ codes, int64_t n, int64_t code) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = __ldg(codes + mid);
    if (v < code) lo = mid + 1;
    else hi = mid;
  }
  return lo < n && __ldg(codes + lo) == code;
}

__global__ void __launch_bounds__(256) sample_negatives_kernel(const int32_t* __restrict__ key_id, int64_t n, int num,
                                                               int n_user, int n_item, const int64_t* __restrict__ codes,
                                                               int64_t n_codes, int64_t span, uint64_t seed,
                                                               int32_t* __restrict__ out, int64_t out_stride, int64_t out_col0) {
  const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n * num) return;
  const int64_t r = cell / num;
  const int j = (int)(cell - r * num);
  const int64_t kid = key_id ? (int64_t)__ldg(key_id + r) : -1;
  int item = 0;
  for (uint32_t attempt = 0;; attempt++) {
    item = draw_item(seed, (uint64_t)cell, attempt, n_user, n_item);
    // a key that never trained has no positives (defaultdict(set) in the reference); 4096 attempts bound the loop on a
    // degenerate key that holds the whole catalog (the reference would spin forever there)
    if (kid < 0 || attempt >= 4096u || !code_present(codes, n_codes, kid * span + item)) break;
  }
  out[r * out_stride + out_col0 + j] = item;
}

// out[(r*num + j), :F] = rows[r, :F] with column 1 (the item) replaced by items[r, j]; columns [F, out_stride) = -1
// (FM.py:303-305; the padding keeps pair-ranking records 16-byte aligned)
__global__ void __launch_bounds__(256) expand_rows_kernel(const int32_t* __restrict__ rows, int64_t n, int F, int64_t stride,
                                                          const int32_t* __restrict__ items, int num, int32_t* __restrict__ out,
                                                          int out_stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * num * out_stride) return;
  const int f = (int)(i % out_stride);
  const int64_t cell = i / out_stride;
  const int64_t r = cell / num;
  out[i] = (f >= F) ? -1 : (f == 1) ? __ldg(items + cell) : __ldg(rows + r * stride + f);
}

// wins += #{(r, j): pos[r] > neg[r*num + j]}      (FM.py:321-323)
__global__ void __launch_bounds__(256) auc_count_kernel(const float* __restrict__ pos, const float* __restrict__ neg, int64_t n,
                                                        int num, unsigned long long* __restrict__ wins) {
  __shared__ unsigned long long s_w[8];
  unsigned long long w = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n * num; i += (int64_t)gridDim.x * blockDim.x)
    w += (__ldg(pos + i / num) > __ldg(neg + i)) ? 1ull : 0ull;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int i = 0; i < 8; i++) t += s_w[i];
    if (t) atomicAdd(wins, t);
  }
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int hhfm_sample_negatives(const int32_t* key_id, int64_t n, int32_t num, int32_t n_user, int32_t n_item,
                                     const int64_t* pf_codes, int64_t n_codes, int64_t span, uint64_t seed, int32_t* out,
                                     int64_t out_stride, int64_t out_col0, hhfm_stream_t stream) {
  HHFM_REQUIRE(out && n >= 0 && num >= 1 && n_item >= 1 && n_user >= 0, "sample_negatives: bad arguments");
  HHFM_REQUIRE(out_stride >= out_col0 + num && out_col0 >= 0, "sample_negatives: out_stride too small");
  HHFM_REQUIRE(n_codes == 0 || pf_codes, "sample_negatives: pf_codes is NULL");
  if (n == 0) return HHFM_OK;
  const int64_t cells = n * num;
  sample_negatives_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      key_id, n, num, n_user, n_item, pf_codes, n_codes, span, seed, out, out_stride, out_col0);
  return check_launch("sample_negatives_kernel");
}

extern "C" int hhfm_expand_rows(const int32_t* rows, int64_t n, int32_t F, int64_t stride, const int32_t* items, int32_t num,
                                int32_t* out, int32_t out_stride, hhfm_stream_t stream) {
  HHFM_REQUIRE(rows && items && out && F >= 2 && stride >= F && out_stride >= F && num >= 1 && n >= 0, "expand_rows: bad arguments");
  if (n == 0) return HHFM_OK;
  const int64_t total = n * num * out_stride;
  expand_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows, n, F, stride, items, num, out,
                                                                                      out_stride);
  return check_launch("expand_rows_kernel");
}

extern "C" int hhfm_auc_count(const float* pos, const float* neg, int64_t n, int32_t num, uint64_t* wins, hhfm_stream_t stream) {
  HHFM_REQUIRE(pos && neg && wins && num >= 1 && n >= 0, "auc_count: bad arguments");
  if (n == 0) return HHFM_OK;
  int64_t blocks = (n * num + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  auc_count_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pos, neg, n, num, reinterpret_cast<unsigned long long*>(wins));
  return check_launch("auc_count_kernel");
}
