// K2: Attentional FM (Newcode/AFM.py:103-148), fp32 SIMT, forward and fused forward + squared loss + backward.
//
//   E = V[x] [F,K];  P_p = E_i * E_j for the pairs i<j in lexicographic order (AFM.py:105-112), never materialised in HBM
//   Z_p = P_p W + b (W [K,A]);  s_p = relu(Z_p) . p;  a = softmax_p(s)  (AFM.py:117-125)
//   afm = sum_p a_p P_p;  out = afm . w_pred + sum_f bias[x_f] + b0      (AFM.py:130-142)
//   loss = 0.5 sum (y - out)^2  (+ lamda_attention/2 ||W||^2 is added by the caller through the dense-L2 optimizer)
//
// One warp owns one sample; a CTA of NW warps works on NW samples at a time.  W lives in shared memory with a padded
// row (A+1) so that both the forward (lanes over a) and the backward (lanes over k) read it without bank conflicts.
// The weight gradient dW = sum P_p^T dZ_p is accumulated in REGISTERS: warp w owns the k-slice [w K/NW, (w+1) K/NW)
// and, after a CTA barrier, sweeps the (E, dZ) of all NW samples of the tile from shared memory.  Embedding gradients
// go through the same sort-free vector reductions (and hot-row replicas) as the FM / HHFM kernels.
// This kernel is compute-bound (about 1.1 MFLOP per sample at F=10, K=A=64); a tcgen05 version with split-precision
// operands is the planned next step.
#include "common.cuh"

namespace hhfm {

constexpr int kAfmMaxF = 16;
constexpr int kAfmMaxP = kAfmMaxF * (kAfmMaxF - 1) / 2;

struct AfmArgs {
  const int32_t* idx;     // [B, F]
  int64_t B;
  int F, K, A, P;
  const float* V;
  const float* bias;
  const float* b0;
  const float* W;         // [K, A]
  const float* batt;      // [A]
  const float* pvec;      // [A]
  const float* wpred;     // [K]
  const float* labels;
  float* out;
  float* gV;
  float* gbias;
  float* gb0;
  float* gW;
  float* gbatt;
  float* gp;
  float* gwpred;
  float* loss_partials;
  int32_t* touch_stamp;
  int32_t stamp;
  int32_t* touched_rows;
  int32_t* touched_count;
  HotPlan hot;
};

__host__ __device__ inline size_t afm_per_warp_floats(int F, int K, int A, int P) {
  return 2 * (size_t)F * K + (size_t)P * A + (size_t)((2 * P + 3) / 4 * 4);
}

__host__ __device__ inline size_t afm_smem_floats(int NW, int F, int K, int A, int P) {
  // sW[K][A+1] + sp[A] + sb[A] + swp[K] + NW * (sE[F][K] + sDE[F][K] + sZ[P][A] + sS[P] + sD[P] (padded)) + valid flags
  return (size_t)K * (A + 1) + 2 * (size_t)A + K + (size_t)NW * afm_per_warp_floats(F, K, A, P) + 32;
}

constexpr int PB = 8;       // pairs per register-blocked sweep of the two matrix products

template <int TK, int TA, int NW, bool TRAIN>
__global__ void __launch_bounds__(NW * 32) afm_kernel(const AfmArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float scratch[32];
  const int K = a.K, A = a.A, F = a.F, P = a.P;
  const int AS = A + 1;
  float* sW = smem;                       // [K][A+1]
  float* sp = sW + (size_t)K * AS;        // [A]
  float* sb = sp + A;                     // [A]
  float* swp = sb + A;                    // [K]
  float* warp_base = swp + K;
  const size_t per_warp = afm_per_warp_floats(F, K, A, P);
  int* sValid = reinterpret_cast<int*>(warp_base + (size_t)NW * per_warp);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sE = warp_base + warp * per_warp;   // [F][K]
  float* sDE = sE + (size_t)F * K;           // [F][K]  d E accumulated over the pairs
  float* sZ = sDE + (size_t)F * K;           // [P][A]  Z, later dZ
  float* sS = sZ + (size_t)P * A;            // [P]     s, later a
  float* sD = sS + P;                        // [P]     d a, later d s

  for (int i = threadIdx.x; i < K * A; i += blockDim.x) sW[(i / A) * AS + (i % A)] = __ldg(a.W + i);
  for (int i = threadIdx.x; i < A; i += blockDim.x) { sp[i] = __ldg(a.pvec + i); sb[i] = __ldg(a.batt + i); }
  for (int i = threadIdx.x; i < K; i += blockDim.x) swp[i] = __ldg(a.wpred + i);
  __syncthreads();

  const float b0 = a.b0 ? __ldg(a.b0) : 0.f;
  const int rep = a.hot.slot ? (int)(((int64_t)blockIdx.x * NW + warp) % a.hot.n_rep) : 0;
  // pair table (i,j) packed as i*16+j, computed once per warp into registers-of-smem: reuse scratch-free small array
  __shared__ unsigned char sPairI[kAfmMaxP], sPairJ[kAfmMaxP];
  if (threadIdx.x == 0) {
    int p = 0;
    for (int i = 0; i < F; i++)
      for (int j = i + 1; j < F; j++) { sPairI[p] = (unsigned char)i; sPairJ[p] = (unsigned char)j; p++; }
  }
  __syncthreads();

  // persistent accumulators
  constexpr int KSMAX = (32 * TK) / NW;              // k-slice rows per warp (K <= 32*TK)
  float dWacc[TRAIN ? KSMAX : 1][TA];
  float gbatt_acc[TA], gp_acc[TA], gwp_acc[TK];
  float loss_acc = 0.f, g0_acc = 0.f;
  if (TRAIN) {
#pragma unroll
    for (int r = 0; r < KSMAX; r++)
#pragma unroll
      for (int t = 0; t < TA; t++) dWacc[r][t] = 0.f;
  }
#pragma unroll
  for (int t = 0; t < TA; t++) { gbatt_acc[t] = 0.f; gp_acc[t] = 0.f; }
#pragma unroll
  for (int t = 0; t < TK; t++) gwp_acc[t] = 0.f;
  const int KS = K / NW;                             // host guarantees K % NW == 0

  for (int64_t s0 = (int64_t)blockIdx.x * NW; s0 < a.B; s0 += (int64_t)gridDim.x * NW) {
    const int64_t s = s0 + warp;
    const bool valid = s < a.B;
    if (lane == 0) sValid[warp] = valid ? 1 : 0;
    if (valid) {
      const int32_t* rec = a.idx + s * F;
      // ---- gather E (coalesced float4 rows) and the bias sum ----
      float bsum = 0.f;
      for (int f = 0; f < F; f++) {
        const int id = __ldg(rec + f);
        const float4* src = reinterpret_cast<const float4*>(a.V + (size_t)id * K);
        for (int c = lane; c < (K >> 2); c += 32) reinterpret_cast<float4*>(sE + f * K)[c] = __ldg(src + c);
        if (a.bias) bsum += __ldg(a.bias + id);
      }
      __syncwarp();
      // ---- attention logits: Z_p = P_p W + b, s_p = relu(Z_p) . p   (4 pairs per sweep over k) ----
      for (int p0 = 0; p0 < P; p0 += PB) {
        float acc[PB][TA];
#pragma unroll
        for (int q = 0; q < PB; q++)
#pragma unroll
          for (int t = 0; t < TA; t++) acc[q][t] = (lane + 32 * t < A) ? sb[lane + 32 * t] : 0.f;
        const float* ei[PB]; const float* ej[PB];
#pragma unroll
        for (int q = 0; q < PB; q++) {
          const int p = min(p0 + q, P - 1);
          ei[q] = sE + sPairI[p] * K; ej[q] = sE + sPairJ[p] * K;
        }
        for (int k = 0; k < K; k += 4) {
          float pk[PB][4];
#pragma unroll
          for (int q = 0; q < PB; q++) {
            const float4 x = *reinterpret_cast<const float4*>(ei[q] + k);
            const float4 y = *reinterpret_cast<const float4*>(ej[q] + k);
            pk[q][0] = x.x * y.x; pk[q][1] = x.y * y.y; pk[q][2] = x.z * y.z; pk[q][3] = x.w * y.w;
          }
#pragma unroll
          for (int kk = 0; kk < 4; kk++) {
#pragma unroll
            for (int t = 0; t < TA; t++) {
              const int aa = lane + 32 * t;
              const float w = (aa < A) ? sW[(k + kk) * AS + aa] : 0.f;
#pragma unroll
              for (int q = 0; q < PB; q++) acc[q][t] = fmaf(pk[q][kk], w, acc[q][t]);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < PB; q++) {
          const int p = p0 + q;
          if (p < P) {
            float part = 0.f;
#pragma unroll
            for (int t = 0; t < TA; t++) {
              const int aa = lane + 32 * t;
              if (aa < A) {
                sZ[p * A + aa] = acc[q][t];
                part = fmaf(fmaxf(acc[q][t], 0.f), sp[aa], part);
              }
            }
            part = warp_sum(part);
            if (lane == 0) sS[p] = part;
          }
        }
      }
      __syncwarp();
      // ---- softmax over pairs (AFM.py:125) ----
      float mx = -INFINITY;
      for (int p = lane; p < P; p += 32) mx = fmaxf(mx, sS[p]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float den = 0.f;
      for (int p = lane; p < P; p += 32) { const float e = expf(sS[p] - mx); sS[p] = e; den += e; }
      den = warp_sum(den);
      __syncwarp();
      for (int p = lane; p < P; p += 32) sS[p] = sS[p] / den;
      __syncwarp();
      // ---- afm = sum_p a_p P_p (lanes over k), out ----
      float afm[TK];
#pragma unroll
      for (int t = 0; t < TK; t++) afm[t] = 0.f;
      for (int p = 0; p < P; p++) {
        const float ap = sS[p];
        const float* xi = sE + sPairI[p] * K; const float* xj = sE + sPairJ[p] * K;
#pragma unroll
        for (int t = 0; t < TK; t++) {
          const int k = lane + 32 * t;
          if (k < K) afm[t] = fmaf(ap, xi[k] * xj[k], afm[t]);
        }
      }
      float part = 0.f;
#pragma unroll
      for (int t = 0; t < TK; t++) { const int k = lane + 32 * t; if (k < K) part = fmaf(afm[t], swp[k], part); }
      const float bil = warp_sum(part);
      const float out = (bil + bsum) + b0;                       // AFM.py:142 add_n
      if (a.out && lane == 0) a.out[s] = out;

      if (TRAIN) {
        const float y = __ldg(a.labels + s);
        const float diff = y - out;
        const float g = -diff;
        if (lane == 0) { loss_acc += 0.5f * diff * diff; g0_acc += g; }
        // d afm = g w_pred ; d w_pred += g afm
        float dafm[TK];
#pragma unroll
        for (int t = 0; t < TK; t++) {
          const int k = lane + 32 * t;
          dafm[t] = (k < K) ? g * swp[k] : 0.f;
          gwp_acc[t] = fmaf(g, afm[t], gwp_acc[t]);
        }
        // d a_p = P_p . d afm ; softmax backward
        float dot = 0.f;
        for (int p = 0; p < P; p++) {
          const float* xi = sE + sPairI[p] * K; const float* xj = sE + sPairJ[p] * K;
          float v = 0.f;
#pragma unroll
          for (int t = 0; t < TK; t++) { const int k = lane + 32 * t; if (k < K) v = fmaf(xi[k] * xj[k], dafm[t], v); }
          v = warp_sum(v);
          if (lane == 0) sD[p] = v;
          dot = fmaf(sS[p], v, dot);
        }
        __syncwarp();
        for (int p = lane; p < P; p += 32) sD[p] = sS[p] * (sD[p] - dot);       // d s_p
        __syncwarp();
        // d Z_p = d s_p p (Z_p > 0) (overwrites Z) ; d p += d s_p relu(Z_p) ; d b += d Z_p
        for (int p = 0; p < P; p++) {
          const float ds = sD[p];
#pragma unroll
          for (int t = 0; t < TA; t++) {
            const int aa = lane + 32 * t;
            if (aa < A) {
              const float z = sZ[p * A + aa];
              const float dz = (z > 0.f) ? ds * sp[aa] : 0.f;
              gp_acc[t] = fmaf(ds, fmaxf(z, 0.f), gp_acc[t]);
              gbatt_acc[t] += dz;
              sZ[p * A + aa] = dz;
            }
          }
        }
        __syncwarp();
        // d P_p = a_p d afm + W d Z_p (lanes over k, 4 pairs per sweep over a) ; d E_i += dP*E_j ; d E_j += dP*E_i
        for (int i = lane; i < F * K; i += 32) sDE[i] = 0.f;
        __syncwarp();
        for (int p0 = 0; p0 < P; p0 += PB) {
          float dp[PB][TK];
#pragma unroll
          for (int q = 0; q < PB; q++) {
            const float ap = (p0 + q < P) ? sS[p0 + q] : 0.f;
#pragma unroll
            for (int t = 0; t < TK; t++) dp[q][t] = ap * dafm[t];
          }
          for (int a0 = 0; a0 < A; a0 += 4) {
            float dz[PB][4];
#pragma unroll
            for (int q = 0; q < PB; q++) {
              const float4 v = *reinterpret_cast<const float4*>(sZ + min(p0 + q, P - 1) * A + a0);
              dz[q][0] = v.x; dz[q][1] = v.y; dz[q][2] = v.z; dz[q][3] = v.w;
            }
#pragma unroll
            for (int aa = 0; aa < 4; aa++) {
#pragma unroll
              for (int t = 0; t < TK; t++) {
                const int k = lane + 32 * t;
                const float w = (k < K) ? sW[k * AS + a0 + aa] : 0.f;
#pragma unroll
                for (int q = 0; q < PB; q++) dp[q][t] = fmaf(w, dz[q][aa], dp[q][t]);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < PB; q++) {
            const int p = p0 + q;
            if (p < P) {
              const int fi = sPairI[p], fj = sPairJ[p];
#pragma unroll
              for (int t = 0; t < TK; t++) {
                const int k = lane + 32 * t;
                if (k < K) {                 // the warp is the only writer of sDE and lane k owns column k
                  const float xi = sE[fi * K + k], xj = sE[fj * K + k];
                  sDE[fi * K + k] = fmaf(dp[q][t], xj, sDE[fi * K + k]);
                  sDE[fj * K + k] = fmaf(dp[q][t], xi, sDE[fj * K + k]);
                }
              }
            }
          }
        }
        // scatter d E rows (vector reductions), d bias
        __syncwarp();
        for (int f = 0; f < F; f++) {
          const int id = __ldg(rec + f);
          float* dst = a.gV + (size_t)id * K;
          if (a.hot.slot != nullptr) {
            const int hs = __ldg(a.hot.slot + id);
            if (hs >= 0) dst = a.hot.ghot + ((size_t)rep * a.hot.n_hot + hs) * K;
          }
          for (int c = lane; c < (K >> 2); c += 32) red_add_v4(dst + 4 * c, reinterpret_cast<const float4*>(sDE + f * K)[c]);
          if (lane == 0) {
            if (a.gbias) scatter_bias(a.gbias, a.hot, rep, id, g);
            touch_row(a.touch_stamp, a.stamp, a.touched_rows, a.touched_count, id);
          }
        }
      }
    }
    if (TRAIN) {
      __syncthreads();
      // ---- d W: warp w owns k in [w*KS, (w+1)*KS); sweep (E, dZ) of every valid sample of this tile ----
      for (int ws = 0; ws < NW; ws++) {
        if (!sValid[ws]) continue;
        const float* oE = warp_base + ws * per_warp;
        const float* oZ = oE + 2 * (size_t)F * K;
        for (int p = 0; p < P; p++) {
          float dz[TA];
#pragma unroll
          for (int t = 0; t < TA; t++) { const int aa = lane + 32 * t; dz[t] = (aa < A) ? oZ[p * A + aa] : 0.f; }
          const float* xi = oE + sPairI[p] * K + warp * KS; const float* xj = oE + sPairJ[p] * K + warp * KS;
#pragma unroll
          for (int r = 0; r < KSMAX; r += 4) {
            if (r < KS) {
              const float4 x = *reinterpret_cast<const float4*>(xi + r);
              const float4 y = *reinterpret_cast<const float4*>(xj + r);
              const float pk[4] = {x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w};
#pragma unroll
              for (int rr = 0; rr < 4; rr++)
#pragma unroll
                for (int t = 0; t < TA; t++) dWacc[r + rr][t] = fmaf(pk[rr], dz[t], dWacc[r + rr][t]);
            }
          }
        }
      }
      __syncthreads();
    }
  }

  if (TRAIN) {
    // flush the register accumulators
#pragma unroll
    for (int r = 0; r < KSMAX; r++) {
      if (r < KS) {
#pragma unroll
        for (int t = 0; t < TA; t++) {
          const int aa = lane + 32 * t;
          if (aa < A && dWacc[r][t] != 0.f) atomicAdd(a.gW + (size_t)(warp * KS + r) * A + aa, dWacc[r][t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < TA; t++) {
      const int aa = lane + 32 * t;
      if (aa < A) { atomicAdd(a.gbatt + aa, gbatt_acc[t]); atomicAdd(a.gp + aa, gp_acc[t]); }
    }
#pragma unroll
    for (int t = 0; t < TK; t++) { const int k = lane + 32 * t; if (k < K) atomicAdd(a.gwpred + k, gwp_acc[t]); }
    const float bl = block_sum(loss_acc, scratch);
    write_partial(a.loss_partials, bl);
    const float bg = block_sum(g0_acc, scratch);
    if (threadIdx.x == 0 && a.gb0 && bg != 0.f) atomicAdd(a.gb0, bg);
  }
}

template <int TK, int TA, int NW, bool TRAIN>
static int launch_afm(const AfmArgs& a, cudaStream_t st) {
  const size_t smem = afm_smem_floats(NW, a.F, a.K, a.A, a.P) * sizeof(float);
  if (smem > 220 * 1024) return 1;
  auto kern = afm_kernel<TK, TA, NW, TRAIN>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("afm_kernel: cannot reserve %zu bytes of shared memory", smem);
    return HHFM_ERR_LAUNCH;
  }
  int occ = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NW * 32, smem);
  if (occ < 1) occ = 1;
  int64_t need = (a.B + NW - 1) / NW;
  int64_t cap = (int64_t)sm_count() * occ;
  if (cap > kPartials) cap = kPartials;
  const int grid = (int)(need < cap ? need : cap);
  kern<<<grid, NW * 32, smem, st>>>(a);
  return check_launch("afm_kernel");
}

template <bool TRAIN>
static int dispatch_afm(const AfmArgs& a, cudaStream_t st) {
  // largest warp count whose shared-memory footprint fits; K % NW == 0 and K/NW % 4 == 0 are needed by the dW sweep
  const int tk = (a.K + 31) / 32, ta = (a.A + 31) / 32;
  int rc = 1;
#define TRY(TKV, TAV, NWV)                                                      \
  if (rc == 1 && tk == TKV && ta == TAV && a.K % (NWV * 4) == 0) rc = launch_afm<TKV, TAV, NWV, TRAIN>(a, st);
  TRY(1, 1, 4) TRY(1, 1, 2) TRY(2, 2, 4) TRY(2, 2, 2) TRY(4, 4, 4) TRY(4, 4, 2) TRY(4, 4, 1)
#undef TRY
  if (rc == 1) {
    set_error("afm: configuration F=%d K=%d A=%d does not fit (need K,A in {<=32,<=64,<=128} with equal tiers, shared memory <= 220 KB)",
              a.F, a.K, a.A);
    return HHFM_ERR_UNSUPPORTED;
  }
  return rc;
}

static int check_afm(const int32_t* idx, int64_t B, int64_t F, const float* V, const float* W, const float* batt,
                     const float* pvec, const float* wpred, int64_t M, int64_t K, int64_t A) {
  HHFM_REQUIRE(idx && V && W && batt && pvec && wpred, "afm: NULL argument");
  HHFM_REQUIRE(B >= 0 && M > 0, "afm: bad sizes");
  HHFM_REQUIRE(F >= 2 && F <= kAfmMaxF, "afm: F=%lld out of range [2,%d]", (long long)F, kAfmMaxF);
  HHFM_REQUIRE(K % 4 == 0 && K >= 4 && K <= 128 && A % 4 == 0 && A >= 4 && A <= 128, "afm: K=%lld A=%lld unsupported (multiples of 4, <= 128)",
               (long long)K, (long long)A);
  HHFM_REQUIRE(((uintptr_t)V & 15) == 0, "afm: V must be 16-byte aligned");
  return HHFM_OK;
}

static AfmArgs make_afm(const int32_t* idx, int64_t B, int64_t F, const float* V, const float* bias, const float* b0,
                        const float* W, const float* batt, const float* pvec, const float* wpred, int64_t K, int64_t A) {
  AfmArgs a{};
  a.idx = idx; a.B = B; a.F = (int)F; a.K = (int)K; a.A = (int)A; a.P = (int)(F * (F - 1) / 2);
  a.V = V; a.bias = bias; a.b0 = b0; a.W = W; a.batt = batt; a.pvec = pvec; a.wpred = wpred;
  return a;
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int hhfm_afm_fwd(const int32_t* idx, int64_t B, int64_t F, const float* V, const float* bias, const float* b0,
                            const float* W, const float* batt, const float* pvec, const float* wpred, int64_t M, int64_t K,
                            int64_t A, float* out, hhfm_stream_t stream) {
  int rc = check_afm(idx, B, F, V, W, batt, pvec, wpred, M, K, A);
  if (rc) return rc;
  HHFM_REQUIRE(out != nullptr, "afm_fwd: out is NULL");
  if (B == 0) return HHFM_OK;
  AfmArgs a = make_afm(idx, B, F, V, bias, b0, W, batt, pvec, wpred, K, A);
  a.out = out;
  return dispatch_afm<false>(a, (cudaStream_t)stream);
}

extern "C" int hhfm_afm_fwd_bwd_sqloss(const int32_t* idx, int64_t B, int64_t F, const float* V, const float* bias,
                                       const float* b0, const float* W, const float* batt, const float* pvec,
                                       const float* wpred, int64_t M, int64_t K, int64_t A, const float* labels, float* out,
                                       float* gV, float* gbias, float* gb0, float* gW, float* gbatt, float* gp, float* gwpred,
                                       float* loss_partials, int32_t* touch_stamp, int32_t stamp, int32_t* touched_rows,
                                       int32_t* touched_count, const int32_t* hot_slot, float* ghot, float* ghot_bias,
                                       int32_t n_rep, int32_t n_hot, hhfm_stream_t stream) {
  int rc = check_afm(idx, B, F, V, W, batt, pvec, wpred, M, K, A);
  if (rc) return rc;
  HHFM_REQUIRE(B > 0 && labels && gV && gW && gbatt && gp && gwpred && loss_partials, "afm_fwd_bwd_sqloss: NULL argument");
  HHFM_REQUIRE(!touch_stamp || (touched_rows && touched_count), "afm_fwd_bwd_sqloss: touch_stamp needs touched_rows/count");
  HHFM_REQUIRE(!hot_slot || (ghot && n_rep >= 1 && n_hot >= 1), "afm_fwd_bwd_sqloss: hot_slot needs ghot, n_rep, n_hot");
  AfmArgs a = make_afm(idx, B, F, V, bias, b0, W, batt, pvec, wpred, K, A);
  a.labels = labels; a.out = out; a.gV = gV; a.gbias = gbias; a.gb0 = gb0; a.gW = gW; a.gbatt = gbatt; a.gp = gp;
  a.gwpred = gwpred; a.loss_partials = loss_partials; a.touch_stamp = touch_stamp; a.stamp = stamp;
  a.touched_rows = touched_rows; a.touched_count = touched_count;
  a.hot = HotPlan{hot_slot, ghot, ghot_bias, n_rep, n_hot};
  return dispatch_afm<true>(a, (cudaStream_t)stream);
}
