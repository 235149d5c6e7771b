// K2 on the tensor cores: the three matrix products of Attentional FM (AFM.py:103-148 and their gradients) as 3xTF32 split
// GEMMs on tcgen05 (dfm_tc.cu), the rest as small warp-per-sample kernels.  Per chunk of n samples (R = n*P pair rows):
//   afm_pairs_kernel      P_p = E_i * E_j                                     -> Pm [R, K]
//   split pass            Pm -> lo part, k-blocked transposes
//   GEMM                  H = relu(Pm W + b)                                   -> H [R, A]          (AFM.py:117-123)
//   afm_attn_kernel       s_p = H_p . p, a = softmax_p(s), afm = sum a_p P_p, out, loss; backward through the softmax:
//                         dZ_p = ds_p p (H_p > 0) overwrites H; a, g kept      (AFM.py:123-146)
//   split pass            dZ -> lo part, k-blocked transposes
//   GEMM                  dP = dZ W^T                                          -> dP [R, K]
//   GEMM (split-K)        dW += Pm^T dZ
//   afm_embed_bwd_kernel  dP_p += a_p g w_pred; dE_i += dP_p*E_j, dE_j += dP_p*E_i; sort-free scatter (hot-row replicas)
// The chunk (2048 samples, ~0.2 GB of intermediates) keeps most of the intermediate traffic in the 126 MB L2.
// The fused fp32 CUDA-core kernel (afm.cu) stays as the path for shapes this one does not cover and as a cross-check.
#include <stdlib.h>

#include "common.cuh"
#include "dfm_tc.cuh"

namespace hhfm {

constexpr int kAfmTcMaxF = 16;
constexpr int kAfmTcMaxP = kAfmTcMaxF * (kAfmTcMaxF - 1) / 2;

struct AfmTcArgs {
  const int32_t* idx;      // [n, F] (chunk)
  int64_t n;
  int F, K, A, P;
  const float* V;
  const float* bias;
  const float* b0;
  const float* pvec;
  const float* wpred;
  const float* labels;     // chunk
  float* out;              // chunk or NULL
  float* Pm;               // [n*P, K]
  float* H;                // [n*P, A]: relu(Z) on entry, dZ on exit (TRAIN)
  const float* dP;         // [n*P, K]
  float* avec;             // [n*P]
  float* gvec;             // [n]
  float* gV;
  float* gbias;
  float* gb0;
  float* gbatt;
  float* gp;
  float* gwpred;
  float* loss_partials;
  int accumulate_loss;     // 0: write the partial slots (first chunk), 1: add to them
  int32_t* touch_stamp;
  int32_t stamp;
  int32_t* touched_rows;
  int32_t* touched_count;
  HotPlan hot;
};

__device__ __forceinline__ void pair_table(int F, unsigned char* pi, unsigned char* pj) {
  if (threadIdx.x == 0) {
    int p = 0;
    for (int i = 0; i < F; i++)
      for (int j = i + 1; j < F; j++) { pi[p] = (unsigned char)i; pj[p] = (unsigned char)j; p++; }
  }
  __syncthreads();
}

// one warp per sample: gather E [F, K] into shared memory, write the P pairwise products as rows of Pm
__global__ void __launch_bounds__(256) afm_pairs_kernel(const AfmTcArgs a) {
  extern __shared__ __align__(16) float sm[];
  __shared__ unsigned char sPI[kAfmTcMaxP], sPJ[kAfmTcMaxP];
  pair_table(a.F, sPI, sPJ);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int F = a.F, K = a.K, P = a.P, kv = K >> 2;
  float* sE = sm + (size_t)warp * F * K;
  for (int64_t s = (int64_t)blockIdx.x * nw + warp; s < a.n; s += (int64_t)gridDim.x * nw) {
    for (int f = 0; f < F; f++) {
      const int id = __ldg(a.idx + s * F + f);
      const float4* src = reinterpret_cast<const float4*>(a.V + (size_t)id * K);
      for (int c = lane; c < kv; c += 32) reinterpret_cast<float4*>(sE + f * K)[c] = __ldg(src + c);
    }
    __syncwarp();
    float4* dst = reinterpret_cast<float4*>(a.Pm + (size_t)s * P * K);
    for (int i = lane; i < P * kv; i += 32) {
      const int p = i / kv, c = i - p * kv;
      const float4 x = reinterpret_cast<const float4*>(sE + sPI[p] * K)[c];
      const float4 y = reinterpret_cast<const float4*>(sE + sPJ[p] * K)[c];
      dst[i] = f4_mul(x, y);
    }
    __syncwarp();
  }
}

constexpr int kAT = 4;     // K, A <= 128: up to 4 values per lane

template <bool TRAIN>
__global__ void __launch_bounds__(256) afm_attn_kernel(const AfmTcArgs a) {
  __shared__ float sS[8][kAfmTcMaxP];
  __shared__ float sD[8][kAfmTcMaxP];
  __shared__ float scratch[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int F = a.F, K = a.K, A = a.A, P = a.P;
  const float b0 = a.b0 ? __ldg(a.b0) : 0.f;
  float pv[kAT], wp[kAT], gp_acc[kAT], gb_acc[kAT], gwp_acc[kAT];
#pragma unroll
  for (int t = 0; t < kAT; t++) {
    const int i = lane + 32 * t;
    pv[t] = (i < A) ? __ldg(a.pvec + i) : 0.f;
    wp[t] = (i < K) ? __ldg(a.wpred + i) : 0.f;
    gp_acc[t] = 0.f; gb_acc[t] = 0.f; gwp_acc[t] = 0.f;
  }
  float loss_acc = 0.f, g0_acc = 0.f;
  const int64_t warp_g = (int64_t)blockIdx.x * nw + warp;
  const int rep = a.hot.slot ? (int)(warp_g % a.hot.n_rep) : 0;
  float* S = sS[warp];
  float* D = sD[warp];
  for (int64_t s = warp_g; s < a.n; s += (int64_t)gridDim.x * nw) {
    const int my_id = (lane < F) ? __ldg(a.idx + s * F + lane) : 0;
    const float bsum = warp_sum((lane < F && a.bias) ? __ldg(a.bias + my_id) : 0.f);
    float* Hs = a.H + (size_t)s * P * A;
    const float* Ps = a.Pm + (size_t)s * P * K;
    // attention logits s_p = relu(Z_p) . p
    for (int p = 0; p < P; p++) {
      float part = 0.f;
#pragma unroll
      for (int t = 0; t < kAT; t++) {
        const int i = lane + 32 * t;
        if (i < A) part = fmaf(Hs[p * A + i], pv[t], part);
      }
      part = warp_sum(part);
      if (lane == 0) S[p] = part;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int p = lane; p < P; p += 32) mx = fmaxf(mx, S[p]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float den = 0.f;
    for (int p = lane; p < P; p += 32) { const float e = expf(S[p] - mx); S[p] = e; den += e; }
    den = warp_sum(den);
    __syncwarp();
    for (int p = lane; p < P; p += 32) S[p] = S[p] / den;            // a_p
    __syncwarp();
    float afm[kAT];
#pragma unroll
    for (int t = 0; t < kAT; t++) afm[t] = 0.f;
    for (int p = 0; p < P; p++) {
      const float ap = S[p];
#pragma unroll
      for (int t = 0; t < kAT; t++) {
        const int k = lane + 32 * t;
        if (k < K) afm[t] = fmaf(ap, __ldg(Ps + p * K + k), afm[t]);
      }
    }
    float part = 0.f;
#pragma unroll
    for (int t = 0; t < kAT; t++) part = fmaf(afm[t], wp[t], part);
    const float out = (warp_sum(part) + bsum) + b0;                    // AFM.py:142 add_n
    if (a.out && lane == 0) a.out[s] = out;
    if (!TRAIN) continue;

    const float g = out - __ldg(a.labels + s);
    if (lane == 0) { loss_acc += 0.5f * g * g; g0_acc += g; a.gvec[s] = g; }
    float dafm[kAT];
#pragma unroll
    for (int t = 0; t < kAT; t++) { dafm[t] = g * wp[t]; gwp_acc[t] = fmaf(g, afm[t], gwp_acc[t]); }
    float dot = 0.f;
    for (int p = 0; p < P; p++) {
      float v = 0.f;
#pragma unroll
      for (int t = 0; t < kAT; t++) {
        const int k = lane + 32 * t;
        if (k < K) v = fmaf(__ldg(Ps + p * K + k), dafm[t], v);
      }
      v = warp_sum(v);
      if (lane == 0) D[p] = v;
      dot = fmaf(S[p], v, dot);
    }
    __syncwarp();
    for (int p = lane; p < P; p += 32) {
      D[p] = S[p] * (D[p] - dot);                                      // d s_p
      a.avec[s * P + p] = S[p];
    }
    __syncwarp();
    for (int p = 0; p < P; p++) {
      const float ds = D[p];
#pragma unroll
      for (int t = 0; t < kAT; t++) {
        const int i = lane + 32 * t;
        if (i < A) {
          const float h = Hs[p * A + i];
          const float dz = (h > 0.f) ? ds * pv[t] : 0.f;
          gp_acc[t] = fmaf(ds, h, gp_acc[t]);
          gb_acc[t] += dz;
          Hs[p * A + i] = dz;
        }
      }
    }
    if (lane < F) {
      if (a.gbias) scatter_bias(a.gbias, a.hot, rep, my_id, g);
      touch_row(a.touch_stamp, a.stamp, a.touched_rows, a.touched_count, my_id);
    }
    __syncwarp();
  }
  if (!TRAIN) return;
#pragma unroll
  for (int t = 0; t < kAT; t++) {
    const int i = lane + 32 * t;
    if (i < A) { atomicAdd(a.gbatt + i, gb_acc[t]); atomicAdd(a.gp + i, gp_acc[t]); }
    if (i < K) atomicAdd(a.gwpred + i, gwp_acc[t]);
  }
  const float bl = block_sum(loss_acc, scratch);
  if (a.accumulate_loss) {
    if (threadIdx.x == 0) a.loss_partials[blockIdx.x] += bl;            // same grid as the first chunk: slot owned by this CTA
  } else {
    write_partial(a.loss_partials, bl);
  }
  const float bg = block_sum(g0_acc, scratch);
  if (threadIdx.x == 0 && a.gb0 && bg != 0.f) atomicAdd(a.gb0, bg);
}

// dP_p (GEMM) + a_p g w_pred -> dE rows -> scatter
__global__ void __launch_bounds__(256) afm_embed_bwd_kernel(const AfmTcArgs a) {
  extern __shared__ __align__(16) float sm[];
  __shared__ unsigned char sPI[kAfmTcMaxP], sPJ[kAfmTcMaxP];
  pair_table(a.F, sPI, sPJ);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int F = a.F, K = a.K, P = a.P, kv = K >> 2;
  float* sE = sm + (size_t)warp * 2 * F * K;
  float* sDE = sE + (size_t)F * K;
  float wp[kAT];
#pragma unroll
  for (int t = 0; t < kAT; t++) { const int k = lane + 32 * t; wp[t] = (k < K) ? __ldg(a.wpred + k) : 0.f; }
  const int64_t warp_g = (int64_t)blockIdx.x * nw + warp;
  const int rep = a.hot.slot ? (int)(warp_g % a.hot.n_rep) : 0;
  for (int64_t s = warp_g; s < a.n; s += (int64_t)gridDim.x * nw) {
    for (int f = 0; f < F; f++) {
      const int id = __ldg(a.idx + s * F + f);
      const float4* src = reinterpret_cast<const float4*>(a.V + (size_t)id * K);
      for (int c = lane; c < kv; c += 32) {
        reinterpret_cast<float4*>(sE + f * K)[c] = __ldg(src + c);
        reinterpret_cast<float4*>(sDE + f * K)[c] = f4_zero();
      }
    }
    __syncwarp();
    const float g = __ldg(a.gvec + s);
    const float* dPs = a.dP + (size_t)s * P * K;
    for (int p = 0; p < P; p++) {
      const float ag = __ldg(a.avec + s * P + p) * g;
      const int fi = sPI[p], fj = sPJ[p];
#pragma unroll
      for (int t = 0; t < kAT; t++) {
        const int k = lane + 32 * t;
        if (k < K) {                         // the warp is the only writer of sDE and lane k owns column k
          const float dp = fmaf(ag, wp[t], __ldg(dPs + p * K + k));
          sDE[fi * K + k] = fmaf(dp, sE[fj * K + k], sDE[fi * K + k]);
          sDE[fj * K + k] = fmaf(dp, sE[fi * K + k], sDE[fj * K + k]);
        }
      }
    }
    __syncwarp();
    for (int f = 0; f < F; f++) {
      const int id = __ldg(a.idx + s * F + f);
      float* dst = a.gV + (size_t)id * K;
      if (a.hot.slot != nullptr) {
        const int hs = __ldg(a.hot.slot + id);
        if (hs >= 0) dst = a.hot.ghot + ((size_t)rep * a.hot.n_hot + hs) * K;
      }
      for (int c = lane; c < kv; c += 32) red_add_v4(dst + 4 * c, reinterpret_cast<const float4*>(sDE + f * K)[c]);
    }
    __syncwarp();
  }
}

struct AfmTcLayout {
  int64_t chunk, R;
  int64_t pm, plo, pt, ptlo, h, hlo, ht, htlo, dp, avec, gvec, wp, wplo, wt, wtlo, scratch, scratch_floats, total;
};

static void afm_tc_layout(int64_t B, int F, int K, int A, AfmTcLayout& t) {
  auto r4 = [](int64_t x) { return (x + 3) / 4 * 4; };
  const int P = F * (F - 1) / 2;
  t.chunk = B < 2048 ? B : 2048;
  t.R = t.chunk * P;
  const int64_t Rp = (t.R + 31) / 32 * 32;
  int64_t o = 0;
  t.pm = o; o += t.R * K;
  t.plo = o; o += t.R * K;
  t.pt = o; o += Rp * K;
  t.ptlo = o; o += Rp * K;
  t.h = o; o += t.R * A;
  t.hlo = o; o += t.R * A;
  t.ht = o; o += Rp * A;
  t.htlo = o; o += Rp * A;
  t.dp = o; o += t.R * K;
  t.avec = o; o += r4(t.R);
  t.gvec = o; o += r4(t.chunk);
  t.wp = o; o += (int64_t)K * A;
  t.wplo = o; o += (int64_t)K * A;
  t.wt = o; o += (int64_t)A * K;
  t.wtlo = o; o += (int64_t)A * K;
  t.scratch = r4(o);
  t.scratch_floats = tf_splitk_scratch_floats();
  t.total = t.scratch + t.scratch_floats;
}

static int afm_tc_run(const int32_t* idx, int64_t B, int F, const float* V, const float* bias, const float* b0, const float* W,
                      const float* batt, const float* pvec, const float* wpred, int K, int A, const float* labels, float* out,
                      float* gV, float* gbias, float* gb0, float* gW, float* gbatt, float* gp, float* gwpred, float* loss_partials,
                      int32_t* touch_stamp, int32_t stamp, int32_t* touched_rows, int32_t* touched_count, HotPlan hot, float* ws,
                      bool train, cudaStream_t st) {
  AfmTcLayout t;
  afm_tc_layout(B, F, K, A, t);
  const int P = F * (F - 1) / 2;
  int rc;
  if ((rc = tf_prep_weight(W, K, A, ws + t.wp, ws + t.wplo, A, ws + t.wt, ws + t.wtlo, K, st))) return rc;
  // one grid for every chunk so that a CTA owns the same loss-partial slot in all of them
  const int64_t first = B < t.chunk ? B : t.chunk;
  int attn_grid = (int)((first + 7) / 8);
  if (attn_grid > 2 * sm_count()) attn_grid = 2 * sm_count();
  if (attn_grid > kPartials) attn_grid = kPartials;
  for (int64_t c0 = 0; c0 < B; c0 += t.chunk) {
    const int64_t n = (B - c0 < t.chunk) ? B - c0 : t.chunk;
    const int64_t R = n * P;
    AfmTcArgs a{};
    a.idx = idx + c0 * F; a.n = n; a.F = F; a.K = K; a.A = A; a.P = P;
    a.V = V; a.bias = bias; a.b0 = b0; a.pvec = pvec; a.wpred = wpred;
    a.labels = labels ? labels + c0 : nullptr; a.out = out ? out + c0 : nullptr;
    a.Pm = ws + t.pm; a.H = ws + t.h; a.dP = ws + t.dp; a.avec = ws + t.avec; a.gvec = ws + t.gvec;
    a.gV = gV; a.gbias = gbias; a.gb0 = gb0; a.gbatt = gbatt; a.gp = gp; a.gwpred = gwpred;
    a.loss_partials = loss_partials; a.accumulate_loss = c0 > 0 ? 1 : 0;
    a.touch_stamp = touch_stamp; a.stamp = stamp; a.touched_rows = touched_rows; a.touched_count = touched_count; a.hot = hot;
    const int grid_s = (int)((n + 7) / 8 < 4 * (int64_t)sm_count() ? (n + 7) / 8 : 4 * (int64_t)sm_count());
    const size_t smem_pairs = (size_t)8 * F * K * sizeof(float), smem_bwd = 2 * smem_pairs;
    if (smem_bwd > 48 * 1024) {
      cudaFuncSetAttribute(afm_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pairs);
      cudaFuncSetAttribute(afm_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bwd);
    }
    afm_pairs_kernel<<<grid_s, 256, smem_pairs, st>>>(a);
    if ((rc = check_launch("afm_pairs_kernel"))) return rc;
    if ((rc = tf_split_transpose(ws + t.pm, R, K, K, ws + t.plo, train ? ws + t.pt : nullptr, train ? ws + t.ptlo : nullptr, st))) return rc;
    TfGemm g{};
    g.A = ws + t.pm; g.A_lo = ws + t.plo; g.lda = K; g.B = ws + t.wt; g.B_lo = ws + t.wtlo; g.ldb = K;
    g.M = (int)R; g.N = A; g.K = K; g.epi = TF_EPI_BIAS_RELU; g.C = ws + t.h; g.ldc = A; g.bias = batt;
    if ((rc = tf_gemm(g, st))) return rc;
    if (!train) {
      afm_attn_kernel<false><<<attn_grid, 256, 0, st>>>(a);
      if ((rc = check_launch("afm_attn_kernel"))) return rc;
      continue;
    }
    afm_attn_kernel<true><<<attn_grid, 256, 0, st>>>(a);
    if ((rc = check_launch("afm_attn_kernel"))) return rc;
    if ((rc = tf_split_transpose(ws + t.h, R, A, A, ws + t.hlo, ws + t.ht, ws + t.htlo, st))) return rc;
    TfGemm d{};     // dP = dZ W^T: A = dZ [R, A], B = W [K, A] (N = K rows, contraction over A)
    d.A = ws + t.h; d.A_lo = ws + t.hlo; d.lda = A; d.B = ws + t.wp; d.B_lo = ws + t.wplo; d.ldb = A;
    d.M = (int)R; d.N = K; d.K = A; d.epi = TF_EPI_STORE; d.C = ws + t.dp; d.ldc = K;
    if ((rc = tf_gemm(d, st))) return rc;
    TfGemm w{};     // dW += Pm^T dZ: k-blocked panels over the R pair rows
    w.A = ws + t.pt; w.A_lo = ws + t.ptlo; w.B = ws + t.ht; w.B_lo = ws + t.htlo; w.k_blocked = 1;
    w.M = K; w.N = A; w.K = (int)R; w.epi = TF_EPI_ATOMIC; w.C = gW; w.ldc = A;
    w.scratch = ws + t.scratch; w.scratch_floats = t.scratch_floats;
    if ((rc = tf_gemm(w, st))) return rc;
    afm_embed_bwd_kernel<<<grid_s, 256, smem_bwd, st>>>(a);
    if ((rc = check_launch("afm_embed_bwd_kernel"))) return rc;
  }
  return HHFM_OK;
}

static int afm_tc_check(const int32_t* idx, int64_t B, int64_t F, const float* V, const float* W, const float* batt,
                        const float* pvec, const float* wpred, int64_t M, int64_t K, int64_t A, const void* ws) {
  HHFM_REQUIRE(idx && V && W && batt && pvec && wpred && ws, "afm_tc: NULL argument");
  HHFM_REQUIRE(B >= 0 && M > 0 && B * F * (F - 1) / 2 < (1ll << 31), "afm_tc: bad sizes");
  HHFM_REQUIRE(F >= 2 && F <= kAfmTcMaxF, "afm_tc: F=%lld out of range [2,%d]", (long long)F, kAfmTcMaxF);
  HHFM_REQUIRE(K % 4 == 0 && K >= 4 && K <= 128 && A % 4 == 0 && A >= 4 && A <= 128, "afm_tc: K=%lld A=%lld unsupported (multiples of 4, <= 128)",
               (long long)K, (long long)A);
  HHFM_REQUIRE((((uintptr_t)V | (uintptr_t)ws) & 15) == 0, "afm_tc: V and workspace must be 16-byte aligned");
  return HHFM_OK;
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int64_t hhfm_workspace_bytes_afm(int64_t B, int64_t F, int64_t K, int64_t A) {
  if (B < 1 || F < 2 || F > kAfmTcMaxF || K < 4 || K > 128 || A < 4 || A > 128) return -1;
  AfmTcLayout t;
  afm_tc_layout(B, (int)F, (int)K, (int)A, t);
  return t.total * (int64_t)sizeof(float);
}

extern "C" int hhfm_afm_fwd_tc(const int32_t* idx, int64_t B, int64_t F, const float* V, const float* bias, const float* b0,
                               const float* W, const float* batt, const float* pvec, const float* wpred, int64_t M, int64_t K,
                               int64_t A, float* out, float* workspace, hhfm_stream_t stream) {
  int rc = afm_tc_check(idx, B, F, V, W, batt, pvec, wpred, M, K, A, workspace);
  if (rc) return rc;
  HHFM_REQUIRE(out != nullptr, "afm_fwd_tc: out is NULL");
  if (B == 0) return HHFM_OK;
  return afm_tc_run(idx, B, (int)F, V, bias, b0, W, batt, pvec, wpred, (int)K, (int)A, nullptr, out, nullptr, nullptr, nullptr,
                    nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr, HotPlan{}, workspace, false,
                    (cudaStream_t)stream);
}

extern "C" int hhfm_afm_fwd_bwd_sqloss_tc(const int32_t* idx, int64_t B, int64_t F, const float* V, const float* bias,
                                          const float* b0, const float* W, const float* batt, const float* pvec,
                                          const float* wpred, int64_t M, int64_t K, int64_t A, const float* labels,
                                          float* out, float* gV, float* gbias, float* gb0, float* gW, float* gbatt, float* gp,
                                          float* gwpred, float* loss_partials, int32_t* touch_stamp, int32_t stamp,
                                          int32_t* touched_rows, int32_t* touched_count, const int32_t* hot_slot, float* ghot,
                                          float* ghot_bias, int32_t n_rep, int32_t n_hot, float* workspace,
                                          hhfm_stream_t stream) {
  int rc = afm_tc_check(idx, B, F, V, W, batt, pvec, wpred, M, K, A, workspace);
  if (rc) return rc;
  HHFM_REQUIRE(B > 0 && labels && gV && gW && gbatt && gp && gwpred && loss_partials, "afm_fwd_bwd_sqloss_tc: NULL argument");
  HHFM_REQUIRE(!touch_stamp || (touched_rows && touched_count), "afm_fwd_bwd_sqloss_tc: touch_stamp needs touched_rows/count");
  HHFM_REQUIRE(!hot_slot || (ghot && n_rep >= 1 && n_hot >= 1), "afm_fwd_bwd_sqloss_tc: hot_slot needs ghot, n_rep, n_hot");
  return afm_tc_run(idx, B, (int)F, V, bias, b0, W, batt, pvec, wpred, (int)K, (int)A, labels, out, gV, gbias, gb0, gW, gbatt, gp,
                    gwpred, loss_partials, touch_stamp, stamp, touched_rows, touched_count,
                    HotPlan{hot_slot, ghot, ghot_bias, n_rep, n_hot}, workspace, true, (cudaStream_t)stream);
}
