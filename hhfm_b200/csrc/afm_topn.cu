// AFM full-catalog scorer (AFM.py:209-246): scores [C, N] of every item for every context row, with the
// item-independent part of the attention network hoisted out of the item loop.
//
// A context row has F fields; field `item_col` is replaced by each of the N items in turn.  Of the F(F-1)/2 pairs of
// AFM.py:105-112 only the F-1 pairs that contain the item field depend on the item:
//   * context pairs (i, j != item_col): logits s_p, products P_p and the projections t_p = P_p . w_pred are per CONTEXT.
//     afm_ctx_kernel reduces them to three numbers per context: m = max s_p, D = sum exp(s_p - m), T = sum exp(s_p - m) t_p.
//   * item pairs (item, f): P = E_n * E_f, so Z = P W + b = E_n (diag(E_f) W) + b: for a fixed (context, field) the N item
//     rows multiply ONE K x A matrix W_f = diag(E_f) W.  afm_topn_score_kernel builds W_f in shared memory, keeps 4 item
//     rows x A/2 logit columns per lane in registers (lane = (column half, item mod 16), the layout of afm2_kernel in
//     afm.cu), reads the tile's 512 item rows from a shared-memory copy staged once per CTA, and folds each item pair into a
//     running (max, numerator, denominator) per item.
//   out[c, n] = (T' + sum_f e_f t_f) / (D' + sum_f e_f) + sum of biases + b0,  e_f = exp(s_f - max), primes rescaled to the
//   common max: the softmax-weighted sum of AFM.py:125-139 with the terms grouped differently (fp32 rounding differs from
//   the op-by-op order at the 1e-7 level; the parity tests hold it to 1e-5).
// Work per (context, item): (F-1) K A FMAs instead of F(F-1)/2 K A: 5x fewer at F = 10.
#include <stdlib.h>

#include "common.cuh"

namespace hhfm {

constexpr int kAfmTopnMaxF = 16;

struct AfmTopnArgs {
  const int32_t* rows;      // [C, row_stride] context rows (field item_col ignored)
  int64_t row_stride;
  int C, F, item_col;
  const float *V, *bias, *b0, *W, *batt, *pvec, *wpred;
  int64_t item_base, N;     // item n has table row item_base + n
  float* stats;             // [C, 4]: m, D, T, unused
  float* scores;            // [C, N]
};

// ---- one warp per context: the context-only pairs ----
template <int KD, int NW>
__global__ void __launch_bounds__(NW * 32) afm_ctx_kernel(const AfmTopnArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int TD = (KD + 31) / 32;
  float* sW = smem;                                  // [KD][KD]
  float* sEall = sW + KD * KD;                       // [NW][F][KD]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < KD * KD; i += blockDim.x) sW[i] = __ldg(a.W + i);
  __syncthreads();
  const int c = blockIdx.x * NW + warp;
  if (c >= a.C) return;
  float* sE = sEall + (size_t)warp * a.F * KD;
  const int32_t* rec = a.rows + (int64_t)c * a.row_stride;
  for (int f = 0; f < a.F; f++) {
    if (f == a.item_col) continue;
    const int id = __ldg(rec + f);
    for (int k = lane; k < KD; k += 32) sE[f * KD + k] = __ldg(a.V + (size_t)id * KD + k);
  }
  __syncwarp();
  float m = -INFINITY, D = 0.f, T = 0.f;
  for (int i = 0; i < a.F; i++) {
    if (i == a.item_col) continue;
    for (int j = i + 1; j < a.F; j++) {
      if (j == a.item_col) continue;
      const float* ei = sE + i * KD; const float* ej = sE + j * KD;
      float z[TD], tp = 0.f;
#pragma unroll
      for (int t = 0; t < TD; t++) { const int aa = lane + 32 * t; z[t] = (aa < KD) ? __ldg(a.batt + aa) : 0.f; }
      for (int k = 0; k < KD; k++) {
        const float pk = ei[k] * ej[k];
#pragma unroll
        for (int t = 0; t < TD; t++) { const int aa = lane + 32 * t; if (aa < KD) z[t] = fmaf(pk, sW[k * KD + aa], z[t]); }
      }
      float sp = 0.f;
#pragma unroll
      for (int t = 0; t < TD; t++) {
        const int aa = lane + 32 * t;
        if (aa < KD) { sp = fmaf(fmaxf(z[t], 0.f), __ldg(a.pvec + aa), sp); tp = fmaf(ei[aa] * ej[aa], __ldg(a.wpred + aa), tp); }
      }
      const float s = warp_sum(sp);
      const float tt = warp_sum(tp);
      const float mn = fmaxf(m, s);
      const float sc = (m == -INFINITY) ? 0.f : expf(m - mn);
      const float e = expf(s - mn);
      D = fmaf(D, sc, e);
      T = fmaf(T, sc, e * tt);
      m = mn;
    }
  }
  if (lane == 0) { a.stats[4 * c] = m; a.stats[4 * c + 1] = D; a.stats[4 * c + 2] = T; a.stats[4 * c + 3] = 0.f; }
}

// ---- CTA = (512-item tile, context): the item pairs ----
template <int KD>
__global__ void __launch_bounds__(256, 1) afm_topn_score_kernel(const AfmTopnArgs a) {
  constexpr int NC = KD / 2, NR = 4, TI = 8 * 16 * NR;
  __shared__ __align__(16) float sWf[2][KD * KD];
  __shared__ __align__(16) float stf[2][KD];
  __shared__ __align__(16) float sEc[kAfmTopnMaxF][KD];
  __shared__ float sb[KD], sp[KD], sbias[kAfmTopnMaxF];
  const int c = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = (lane >> 4) * NC, l = lane & 15;
  const int32_t* rec = a.rows + (int64_t)c * a.row_stride;
  for (int i = threadIdx.x; i < a.F * KD; i += blockDim.x) {
    const int f = i / KD, k = i % KD;
    if (f != a.item_col) sEc[f][k] = __ldg(a.V + (size_t)__ldg(rec + f) * KD + k);
  }
  for (int i = threadIdx.x; i < KD; i += blockDim.x) { sb[i] = __ldg(a.batt + i); sp[i] = __ldg(a.pvec + i); }
  if (threadIdx.x < a.F) sbias[threadIdx.x] = (a.bias && (int)threadIdx.x != a.item_col) ? __ldg(a.bias + __ldg(rec + threadIdx.x)) : 0.f;

  // the tile's item rows, staged once for all F-1 fields (row stride KD+4: conflict-free LDS.128 with one row per lane)
  extern __shared__ __align__(16) float sX[];
  constexpr int XS = KD + 4;
  for (int i = threadIdx.x; i < TI * (KD / 4); i += blockDim.x) {
    const int item = i / (KD / 4), c4 = i % (KD / 4);
    int64_t q = (int64_t)blockIdx.x * TI + item;
    q = q < a.N ? q : a.N - 1;
    *reinterpret_cast<float4*>(sX + item * XS + 4 * c4) = __ldg(reinterpret_cast<const float4*>(a.V + (size_t)(a.item_base + q) * KD) + c4);
  }
  int64_t n[NR];
  const float* xrow[NR];
  float m[NR], num[NR], den[NR];
  const float m0 = __ldg(a.stats + 4 * c), d0 = __ldg(a.stats + 4 * c + 1), t0 = __ldg(a.stats + 4 * c + 2);
#pragma unroll
  for (int r = 0; r < NR; r++) {
    const int item = warp * (16 * NR) + l + 16 * r;
    n[r] = (int64_t)blockIdx.x * TI + item;
    xrow[r] = sX + item * XS;
    m[r] = m0; num[r] = t0; den[r] = d0;
  }
  __syncthreads();

  int it = 0;
  for (int f = 0; f < a.F; f++) {
    if (f == a.item_col) continue;
    float* Wf = sWf[it & 1];
    float* tf = stf[it & 1];
    it++;
    // W_f = diag(E_f) W, t_f = E_f * w_pred
    for (int i = threadIdx.x; i < KD * KD / 4; i += blockDim.x) {
      const float e = sEc[f][(4 * i) / KD];
      const float4 w = __ldg(reinterpret_cast<const float4*>(a.W) + i);
      reinterpret_cast<float4*>(Wf)[i] = make_float4(e * w.x, e * w.y, e * w.z, e * w.w);
    }
    for (int i = threadIdx.x; i < KD; i += blockDim.x) tf[i] = sEc[f][i] * __ldg(a.wpred + i);
    __syncthreads();     // one barrier per field: the other buffer's readers finished before the previous barrier

    float acc[NR][NC], t[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
      t[r] = 0.f;
#pragma unroll
      for (int cc = 0; cc < NC; cc++) acc[r][cc] = sb[c0 + cc];
    }
#pragma unroll 1
    for (int k = 0; k < KD; k += 4) {
      float x[NR][4];
      const float4 tv = *reinterpret_cast<const float4*>(tf + k);
#pragma unroll
      for (int r = 0; r < NR; r++) {
        const float4 x4 = *reinterpret_cast<const float4*>(xrow[r] + k);
        x[r][0] = x4.x; x[r][1] = x4.y; x[r][2] = x4.z; x[r][3] = x4.w;
        t[r] = fmaf(x4.x, tv.x, t[r]); t[r] = fmaf(x4.y, tv.y, t[r]); t[r] = fmaf(x4.z, tv.z, t[r]); t[r] = fmaf(x4.w, tv.w, t[r]);
      }
#pragma unroll
      for (int kk = 0; kk < 4; kk++) {
        const float4* w4 = reinterpret_cast<const float4*>(Wf + (k + kk) * KD + c0);
#pragma unroll
        for (int c4 = 0; c4 < NC / 4; c4++) {
          const float4 w = w4[c4];
#pragma unroll
          for (int r = 0; r < NR; r++) {
            acc[r][4 * c4] = fmaf(x[r][kk], w.x, acc[r][4 * c4]);
            acc[r][4 * c4 + 1] = fmaf(x[r][kk], w.y, acc[r][4 * c4 + 1]);
            acc[r][4 * c4 + 2] = fmaf(x[r][kk], w.z, acc[r][4 * c4 + 2]);
            acc[r][4 * c4 + 3] = fmaf(x[r][kk], w.w, acc[r][4 * c4 + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < NR; r++) {
      float s = 0.f;
#pragma unroll
      for (int cc = 0; cc < NC; cc++) s = fmaf(fmaxf(acc[r][cc], 0.f), sp[c0 + cc], s);
      s += __shfl_xor_sync(0xffffffffu, s, 16);
      const float mn = fmaxf(m[r], s);
      const float sc = (m[r] == -INFINITY) ? 0.f : expf(m[r] - mn);
      const float e = expf(s - mn);
      den[r] = fmaf(den[r], sc, e);
      num[r] = fmaf(num[r], sc, e * t[r]);
      m[r] = mn;
    }
  }

  if (lane < 16) {
    const float b0 = a.b0 ? __ldg(a.b0) : 0.f;
#pragma unroll
    for (int r = 0; r < NR; r++) {
      if (n[r] < a.N) {
        float fb = 0.f;
        for (int f = 0; f < a.F; f++) fb += (f == a.item_col) ? (a.bias ? __ldg(a.bias + a.item_base + n[r]) : 0.f) : sbias[f];
        a.scores[(int64_t)c * a.N + n[r]] = (num[r] / den[r] + fb) + b0;        // AFM.py:141-142
      }
    }
  }
}

template <int KD>
static int launch_afm_topn(const AfmTopnArgs& a, cudaStream_t st) {
  constexpr int NW = 4;
  const size_t smem = ((size_t)KD * KD + (size_t)NW * a.F * KD) * sizeof(float);
  auto ctx = afm_ctx_kernel<KD, NW>;
  if (cudaFuncSetAttribute(ctx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("afm_ctx_kernel: cannot reserve %zu bytes of shared memory", smem);
    return HHFM_ERR_LAUNCH;
  }
  ctx<<<(a.C + NW - 1) / NW, NW * 32, smem, st>>>(a);
  int rc = check_launch("afm_ctx_kernel");
  if (rc) return rc;
  constexpr int TI = 512;
  dim3 grid((unsigned)((a.N + TI - 1) / TI), (unsigned)a.C);
  const size_t xs = (size_t)TI * (KD + 4) * sizeof(float);
  auto score = afm_topn_score_kernel<KD>;
  if (cudaFuncSetAttribute(score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)xs) != cudaSuccess) {
    set_error("afm_topn_score_kernel: cannot reserve %zu bytes of shared memory", xs);
    return HHFM_ERR_LAUNCH;
  }
  score<<<grid, 256, xs, st>>>(a);
  return check_launch("afm_topn_score_kernel");
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int hhfm_afm_topn_supported(int64_t F, int64_t K, int64_t A) {
  return (K == A && (K == 16 || K == 32 || K == 64) && F >= 2 && F <= kAfmTopnMaxF) ? 1 : 0;
}

extern "C" int hhfm_afm_topn_scores(const int32_t* rows, int64_t row_stride, int64_t C, int64_t F, int32_t item_col,
                                    const float* V, const float* bias, const float* b0, const float* W, const float* batt,
                                    const float* pvec, const float* wpred, int64_t M, int64_t K, int64_t A, int64_t item_base,
                                    int64_t N, float* stats, float* scores, hhfm_stream_t stream) {
  HHFM_REQUIRE(rows && V && W && batt && pvec && wpred && stats && scores, "afm_topn_scores: NULL argument");
  HHFM_REQUIRE(hhfm_afm_topn_supported(F, K, A), "afm_topn_scores: shape F=%lld K=%lld A=%lld not covered (K == A in {16,32,64}, 2 <= F <= 16)",
               (long long)F, (long long)K, (long long)A);
  HHFM_REQUIRE(item_col >= 0 && item_col < F && row_stride >= F, "afm_topn_scores: bad item_col / row_stride");
  HHFM_REQUIRE(C >= 0 && C <= 65535 && N >= 1 && item_base >= 0 && item_base + N <= M, "afm_topn_scores: bad C / N / item_base");
  if (C == 0) return HHFM_OK;
  AfmTopnArgs a{rows, row_stride, (int)C, (int)F, (int)item_col, V, bias, b0, W, batt, pvec, wpred, item_base, N, stats, scores};
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 64) return launch_afm_topn<64>(a, st);
  if (K == 32) return launch_afm_topn<32>(a, st);
  return launch_afm_topn<16>(a, st);
}
