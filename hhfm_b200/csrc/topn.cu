// K6/K7: full-catalog top-N, exact path, and the evaluate_TopK walk.
// Reference: FM.topk (FM.py:172-185), BPR.topk (BPR.py:131-136), MF.topk (MF.py:144-149), OUR.topk
// (OurModel7.py:229-295), Train.evaluate_TopK (FM.py:325-359).
//
// Exactness contract: scores are computed in the oracle's canonical fp32 order (k ascending, multiply and add
// rounded separately: __fmul_rn/__fadd_rn are never contracted into FMA), so they are bit-identical to
// oracle/hhfm_oracle.py and the (score desc, id asc) selection reproduces tf.nn.top_k index lists exactly.
#include "common.cuh"

namespace hhfm {

// ---------------------------------------------------------------------------------------------------
// query builder: one warp per context row, lanes stride over k; fields are combined in column order.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pool_seq(const float* __restrict__ V, const int32_t* __restrict__ ids, int n, int mode,
                                          int K, int k) {
  float acc = __ldg(V + (size_t)__ldg(ids) * K + k);
  for (int j = 1; j < n; j++) {
    const float x = __ldg(V + (size_t)__ldg(ids + j) * K + k);
    acc = (mode == HHFM_POOL_MAX) ? fmaxf(acc, x) : __fadd_rn(acc, x);
  }
  if (mode == HHFM_POOL_MEAN) acc = __fdiv_rn(acc, (float)n);
  return acc;
}

__global__ void __launch_bounds__(256) build_query_kernel(int kind, const int32_t* __restrict__ A, int64_t C, int stride,
                                                          int n_ctx, int n_time, int pc, int pt, int pf,
                                                          const float* __restrict__ V, int K, float* __restrict__ Q,
                                                          float* __restrict__ Fc) {
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  const int32_t* row = A + c * stride;
  const int lane = threadIdx.x & 31;
  for (int k = lane; k < K; k += 32) {
    const float u = __ldg(V + (size_t)__ldg(row) * K + k);
    float q = u;
    if (kind == HHFM_QUERY_FM) {                      // FM.py:176-177
      float f = 0.f;
      if (n_ctx > 0) f = pool_seq(V, row + 2, n_ctx, HHFM_POOL_SUM, K, k);
      Fc[c * K + k] = f;
      q = __fadd_rn(u, f);
    } else if (kind == HHFM_QUERY_HHFM) {             // OurModel7.py:243-292
      const int num = 1 + (n_ctx > 0) + (n_time > 0);
      float cc = 0.f, tt = 0.f;
      if (n_ctx > 0) cc = pool_seq(V, row + 2, n_ctx, pc, K, k);
      if (n_time > 0) tt = pool_seq(V, row + 2 + n_ctx, n_time, pt, K, k);
      if (num > 1) {
        if (pf == HHFM_POOL_MAX) {
          if (n_ctx > 0) q = fmaxf(q, cc);
          if (n_time > 0) q = fmaxf(q, tt);
        } else {
          if (n_ctx > 0) q = __fadd_rn(q, cc);
          if (n_time > 0) q = __fadd_rn(q, tt);
          if (pf == HHFM_POOL_MEAN) q = __fdiv_rn(q, (float)num);
        }
      }
    }
    Q[c * K + k] = q;
  }
}

// ---------------------------------------------------------------------------------------------------
// exact scorer: CTA tile = 32 contexts x 64 items, K consumed in chunks of 64 through shared memory;
// every output keeps one accumulator that walks k in ascending order across the chunks.
// ---------------------------------------------------------------------------------------------------
constexpr int kTC = 32, kTN = 64, kKC = 64;

template <bool FM>
__global__ void __launch_bounds__(256) score_exact_kernel(const float* __restrict__ Q, const float* __restrict__ Fc,
                                                          int64_t C, const float* __restrict__ items,
                                                          const float* __restrict__ item_bias, int64_t N, int K,
                                                          float* __restrict__ scores, int64_t score_stride) {
  __shared__ float s_it[kTN][kKC + 1];
  __shared__ float s_q[kTC][kKC + 1];
  __shared__ float s_f[FM ? kTC : 1][kKC + 1];
  const int64_t n0 = (int64_t)blockIdx.x * kTN, c0 = (int64_t)blockIdx.y * kTC;
  const int t = threadIdx.x, it = t & (kTN - 1), cg = t >> 6;   // cg in 0..3 ; contexts cg + 4*r
  float acc[kTC / 4];
#pragma unroll
  for (int r = 0; r < kTC / 4; r++) acc[r] = 0.f;

  for (int k0 = 0; k0 < K; k0 += kKC) {
    const int kc = min(kKC, K - k0);
    __syncthreads();
    for (int i = t; i < kTN * kc; i += 256) {
      const int r = i / kc, k = i % kc;
      s_it[r][k] = (n0 + r < N) ? __ldg(items + (n0 + r) * K + k0 + k) : 0.f;
    }
    for (int i = t; i < kTC * kc; i += 256) {
      const int r = i / kc, k = i % kc;
      const bool ok = (c0 + r) < C;
      s_q[r][k] = ok ? __ldg(Q + (c0 + r) * K + k0 + k) : 0.f;
      if (FM) s_f[r][k] = ok ? __ldg(Fc + (c0 + r) * K + k0 + k) : 0.f;
    }
    __syncthreads();
    for (int k = 0; k < kc; k++) {
      const float v = s_it[it][k];
#pragma unroll
      for (int r = 0; r < kTC / 4; r++) {
        const int c = cg + 4 * r;
        const float b = FM ? __fadd_rn(v, s_f[c][k]) : v;       // FM.py:178 ItemWithFeature
        const float p = __fmul_rn(s_q[c][k], b);                // FM.py:180
        acc[r] = (k0 + k == 0) ? p : __fadd_rn(acc[r], p);      // FM.py:183 reduce_sum, k ascending
      }
    }
  }
  const int64_t n = n0 + it;
  if (n < N) {
    const float bias = (FM && item_bias) ? __ldg(item_bias + n) : 0.f;
#pragma unroll
    for (int r = 0; r < kTC / 4; r++) {
      const int64_t c = c0 + cg + 4 * r;
      if (c < C) scores[c * score_stride + n] = (FM && item_bias) ? __fadd_rn(bias, acc[r]) : acc[r];   // FM.py:185
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// top-tp select per row under (score desc, id asc): 64-bit radix select + bitonic sort of the winners.
// key = orderable(score) << 32 | (0xFFFFFFFF - id); all keys of a row are distinct, so exactly
// min(tp, n) keys are >= the tp-th largest.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long make_key(float f, int id) {
  f = f + 0.0f;                                      // -0.0 -> +0.0: tf.nn.top_k compares values, not bits
  unsigned u = __float_as_uint(f);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)id);
}
__device__ __forceinline__ float key_score(unsigned long long k) {
  unsigned u = (unsigned)(k >> 32);
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  return __uint_as_float(u);
}
__device__ __forceinline__ int key_id(unsigned long long k) { return (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)); }

constexpr int kSelectStage = 4096;     // rows with at most this many entries keep their keys in shared memory

__global__ void __launch_bounds__(256) select_kernel(const float* __restrict__ scores, const int32_t* __restrict__ ids,
                                                     const int32_t* __restrict__ counts, int64_t row_stride, int64_t n_max,
                                                     int tp, int tp_pow2, int staged, int id_offset,
                                                     float* __restrict__ out_scores, int32_t* __restrict__ out_ids) {
  extern __shared__ unsigned long long s_keys[];     // tp_pow2 winners (+ the row's keys when staged)
  __shared__ unsigned s_hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_krem, s_cnt, s_done;
  const int64_t c = blockIdx.x;
  const float* sc = scores + c * row_stride;
  const int32_t* idp = ids ? ids + c * row_stride : nullptr;
  int64_t n = counts ? (int64_t)counts[c] : n_max;
  if (n > n_max) n = n_max;
  const int want = (int)(n < tp ? n : tp);
  const int t = threadIdx.x;
  unsigned long long* s_all = s_keys + tp_pow2;
  if (staged)
    for (int64_t i = t; i < n; i += 256) s_all[i] = make_key(sc[i], idp ? idp[i] : (int)i);
  auto key_at = [&](int64_t i) { return staged ? s_all[i] : make_key(sc[i], idp ? idp[i] : (int)i); };

  unsigned long long thresh = 0ull;
  if (n > tp) {
    if (t == 0) { s_prefix = 0ull; s_krem = tp; s_done = 0; }
    for (int pass = 0; pass < 8; pass++) {
      const int shift = 56 - 8 * pass;
      s_hist[t] = 0u;
      __syncthreads();
      const unsigned long long prefix = s_prefix;
      for (int64_t i = t; i < n; i += 256) {
        const unsigned long long k = key_at(i);
        if (pass == 0 || (k >> (shift + 8)) == prefix) atomicAdd(&s_hist[(unsigned)(k >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (t < 32) {
        // bin holding the krem-th largest key: lane t owns bins 8t..8t+7, suffix sums across lanes by shuffles
        unsigned h[8], T = 0;
#pragma unroll
        for (int q = 0; q < 8; q++) { h[q] = s_hist[t * 8 + q]; T += h[q]; }
        unsigned S = T;                               // becomes the sum over lanes >= t
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned v = __shfl_down_sync(0xffffffffu, S, o);
          if (t + o < 32) S += v;
        }
        const unsigned krem = (unsigned)s_krem;
        const unsigned bal = __ballot_sync(0xffffffffu, S >= krem);   // true for lanes 0..target
        const int target = 31 - __clz((int)bal);
        if (t == target) {
          unsigned cum = S - T;
          int bsel = 0;
          bool done = false;
#pragma unroll
          for (int q = 7; q >= 0; q--) {
            if (!done) {
              if (q == 0 || cum + h[q] >= krem) { bsel = q; done = true; }
              else cum += h[q];
            }
          }
          s_krem = (int)(krem - cum);
          s_prefix = (prefix << 8) | (unsigned long long)(t * 8 + bsel);
          // every key of the chosen bin is among the winners: the lowest key this prefix can have is a valid threshold and
          // the remaining digits need not be looked at (typical after 2-3 of the 8 passes on a few hundred candidates)
          if (shift > 0 && h[bsel] == krem - cum) {
            s_prefix <<= shift;
            s_done = 1;
          }
        }
      }
      __syncthreads();
      if (s_done) break;
    }
    thresh = s_prefix;                               // the tp-th best key, or a lower bound that no losing key reaches
  }
  if (t == 0) s_cnt = 0;
  for (int i = t; i < tp_pow2; i += 256) s_keys[i] = 0ull;   // 0 sorts last
  __syncthreads();
  for (int64_t i = t; i < n; i += 256) {
    const unsigned long long k = key_at(i);
    if (k >= thresh) {
      const int slot = atomicAdd(&s_cnt, 1);
      if (slot < tp_pow2) s_keys[slot] = k;
    }
  }
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= tp_pow2; size <<= 1) {
    for (int strd = size >> 1; strd > 0; strd >>= 1) {
      for (int i = t; i < tp_pow2; i += 256) {
        const int j = i ^ strd;
        if (j > i) {
          const unsigned long long a = s_keys[i], b = s_keys[j];
          const bool desc = ((i & size) == 0);
          if (desc ? (a < b) : (a > b)) { s_keys[i] = b; s_keys[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = t; i < tp; i += 256) {
    if (i < want) {
      const unsigned long long k = s_keys[i];
      out_ids[c * tp + i] = key_id(k) + id_offset;
      if (out_scores) out_scores[c * tp + i] = key_score(k);
    } else {
      out_ids[c * tp + i] = -1;
      if (out_scores) out_scores[c * tp + i] = -INFINITY;
    }
  }
}

// evaluate_TopK walk, FM.py:336-357 (one thread per evaluated row)
__global__ void metrics_walk_kernel(const int32_t* __restrict__ pred, const int32_t* __restrict__ target,
                                    const uint8_t* __restrict__ in_pf, int64_t C, int tp, int TopK,
                                    int32_t* __restrict__ rank_code) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int item = target[c];
  const bool pf = in_pf[c] != 0;
  int n = 0, code = -2;
  for (int t = 0; t < tp; t++) {
    const int it = pred[c * tp + t];
    if (n > TopK - 1) { code = -1; break; }             // FM.py:342-347
    else if (it == item) { code = n; break; }           // :348-353
    else if (pf) continue;                              // :354-355 (tests the TARGET item: loop-invariant)
    else n = n + 1;                                     // :357
  }
  rank_code[c] = code;
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int hhfm_topn_build_query(int32_t kind, const int32_t* A, int64_t C, int64_t stride, int32_t n_ctx,
                                     int32_t n_time, int32_t pool_ctx, int32_t pool_time, int32_t pool_stack,
                                     const float* V, int64_t M, int64_t K, float* Q, float* Fc, hhfm_stream_t stream) {
  HHFM_REQUIRE(A && V && Q, "topn_build_query: NULL argument");
  HHFM_REQUIRE(kind >= 0 && kind <= 2, "topn_build_query: bad kind %d", kind);
  HHFM_REQUIRE(kind != HHFM_QUERY_FM || Fc, "topn_build_query: FM needs the Fc output");
  HHFM_REQUIRE(C >= 0 && K > 0 && M > 0 && stride >= 2 + n_ctx + n_time, "topn_build_query: bad sizes");
  if (C == 0) return HHFM_OK;
  const int grid = (int)((C + 7) / 8);
  build_query_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kind, A, C, (int)stride, n_ctx, n_time, pool_ctx, pool_time,
                                                            pool_stack, V, (int)K, Q, Fc);
  return check_launch("build_query_kernel");
}

extern "C" int hhfm_topn_score_exact(int32_t kind, const float* Q, const float* Fc, int64_t C, const float* items,
                                     const float* item_bias, int64_t N, int64_t K, float* scores, int64_t score_stride,
                                     hhfm_stream_t stream) {
  HHFM_REQUIRE(Q && items && scores, "topn_score_exact: NULL argument");
  HHFM_REQUIRE(kind != HHFM_QUERY_FM || Fc, "topn_score_exact: FM needs Fc");
  HHFM_REQUIRE(C >= 0 && N >= 0 && K > 0 && score_stride >= N, "topn_score_exact: bad sizes");
  if (C == 0 || N == 0) return HHFM_OK;
  dim3 grid((unsigned)((N + kTN - 1) / kTN), (unsigned)((C + kTC - 1) / kTC));
  HHFM_REQUIRE(grid.y <= 65535, "topn_score_exact: too many contexts per call (%lld); chunk the rows", (long long)C);
  if (kind == HHFM_QUERY_FM)
    score_exact_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(Q, Fc, C, items, item_bias, N, (int)K, scores, score_stride);
  else
    score_exact_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(Q, nullptr, C, items, nullptr, N, (int)K, scores, score_stride);
  return check_launch("score_exact_kernel");
}

extern "C" int hhfm_topn_select(const float* scores, const int32_t* ids, const int32_t* counts, int64_t C,
                                int64_t row_stride, int64_t n, int32_t tp, int32_t id_offset, float* out_scores,
                                int32_t* out_ids, hhfm_stream_t stream) {
  HHFM_REQUIRE(scores && out_ids, "topn_select: NULL argument");
  HHFM_REQUIRE(tp >= 1 && tp <= 1024, "topn_select: tp=%d out of range [1,1024]", tp);
  HHFM_REQUIRE(C >= 0 && n >= 0 && row_stride >= n, "topn_select: bad sizes");
  if (C == 0) return HHFM_OK;
  int p2 = 1;
  while (p2 < tp) p2 <<= 1;
  const int staged = n <= kSelectStage ? 1 : 0;
  const size_t smem = ((size_t)p2 + (staged ? (size_t)n : 0)) * sizeof(unsigned long long);
  select_kernel<<<(unsigned)C, 256, smem, (cudaStream_t)stream>>>(scores, ids, counts, row_stride, n, tp, p2, staged,
                                                                  id_offset, out_scores, out_ids);
  return check_launch("select_kernel");
}

extern "C" int hhfm_metrics_walk(const int32_t* pred, const int32_t* target, const uint8_t* target_in_pf, int64_t C,
                                 int32_t tp, int32_t TopK, int32_t* rank_code, hhfm_stream_t stream) {
  HHFM_REQUIRE(pred && target && target_in_pf && rank_code, "metrics_walk: NULL argument");
  HHFM_REQUIRE(tp >= 1 && TopK >= 1, "metrics_walk: bad tp/TopK");
  if (C == 0) return HHFM_OK;
  metrics_walk_kernel<<<(unsigned)((C + 127) / 128), 128, 0, (cudaStream_t)stream>>>(pred, target, target_in_pf, C, tp,
                                                                                      TopK, rank_code);
  return check_launch("metrics_walk_kernel");
}
