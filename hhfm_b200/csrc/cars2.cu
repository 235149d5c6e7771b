// K10: CARS2 (Newcode/CARS2.py:66-187), the context-aware baseline of the reference's comparison (main.py:50-63).
//
//   u = UI[user], i = UI[item], n = sum_j UI[neg_j], c = Context[fea]
//   pik_p = sum_{d,c} u_d W[d,p,c] c_c ;  qjk_q(x) = sum_{d,c} x_d Z[d,q,c] c_c
//   PositiveFeadback = u.i + pik.A + qjk(i).B                              (CARS2.py:104-106)
//   loss = -sum log sigmoid(Pos - Neg) + lamda/2 (|UI|^2 + |Context|^2 + |W|^2 + |Z|^2 + |A|^2 + |B|^2)   (:113-123)
// The per-sample work collapses onto two small matrices computed once per call:
//   T[d,c] = sum_q B_q Z[d,q,c],  S[d,c] = sum_p A_p W[d,p,c]
//   PositiveFeadback = u.i + u^T S c + i^T T c ;  Pos - Neg = u.delta + delta^T T c,  delta = i - n   (pik.A cancels)
// so one warp per sample does a [D x Dc] matrix-vector product from shared memory, and the gradient of T is the GEMM
// Delta^T C (rows g*delta and c per sample), expanded afterwards into dZ = B (x) dT and dB_q = <Z[:,q,:], dT>.
// W and A only receive their regulariser gradients.  Full-catalog scoring: score(c, item) = item . (u + T c) + const.
#include "common.cuh"
#include "dfm_tc.cuh"

namespace hhfm {

int sgemm_tn_splitk(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N, int Kd, float* C, int64_t ldc,
                    cudaStream_t st);     // dfm.cu: C[M,N] += A[Kd,M]^T B[Kd,N]

constexpr int kCarsTD = 4;      // D <= 128
constexpr int kCarsTC = 4;      // Dc <= 128
constexpr int kCarsMaxNeg = 16;

struct CarsArgs {
  const int32_t* rec;      // [B, stride]: user, item, fea, neg...
  int64_t B, stride;
  int n_neg, D, Dc;
  const float* UI;
  const float* Ctx;
  const float* T;          // [D, Dc]
  const float* S;          // [D, Dc]
  float* out;              // SCORE: PositiveFeadback [B];  QUERY: Q [B, D] = u + T c
  float* gUI;
  float* gCtx;
  float* Delta;            // [B, D]   g * delta
  float* Cm;               // [B, Dc]  c
  float* loss_partials;
};

enum { CARS_SCORE = 0, CARS_TRAIN = 1, CARS_QUERY = 2 };

// T[d,c] = sum_q B_q Z[d,q,c] ; S[d,c] = sum_p A_p W[d,p,c]
__global__ void cars2_prep_kernel(const float* __restrict__ W, const float* __restrict__ Z, const float* __restrict__ A,
                                  const float* __restrict__ Bv, int D, int Dp, int Dq, int Dc, float* __restrict__ T,
                                  float* __restrict__ S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * Dc) return;
  const int d = i / Dc, c = i % Dc;
  float t = 0.f, s = 0.f;
  for (int q = 0; q < Dq; q++) t = fmaf(__ldg(Bv + q), __ldg(Z + ((size_t)d * Dq + q) * Dc + c), t);
  for (int p = 0; p < Dp; p++) s = fmaf(__ldg(A + p), __ldg(W + ((size_t)d * Dp + p) * Dc + c), s);
  T[i] = t;
  S[i] = s;
}

template <int MODE>
__global__ void __launch_bounds__(256) cars2_kernel(const CarsArgs a) {
  extern __shared__ float sm[];
  __shared__ float scratch[32];
  const int D = a.D, Dc = a.Dc, ldt = Dc + 1;
  float* sT = sm;                                   // [D][Dc+1]
  float* sS = sT + (size_t)D * ldt;                 // [D][Dc+1] (SCORE only)
  const int nmat = (MODE == CARS_SCORE) ? 2 : 1;
  float* warp_base = sm + (size_t)nmat * D * ldt;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float* sc = warp_base + (size_t)warp * (D + Dc);  // [Dc] context vector
  float* sd = sc + Dc;                              // [D]  delta
  for (int i = threadIdx.x; i < D * Dc; i += blockDim.x) {
    sT[(i / Dc) * ldt + (i % Dc)] = __ldg(a.T + i);
    if (MODE == CARS_SCORE) sS[(i / Dc) * ldt + (i % Dc)] = __ldg(a.S + i);
  }
  __syncthreads();
  float loss_acc = 0.f;
  for (int64_t s = (int64_t)blockIdx.x * nw + warp; s < a.B; s += (int64_t)gridDim.x * nw) {
    const int32_t* r = a.rec + s * a.stride;
    const int user = __ldg(r), item = __ldg(r + 1), fea = __ldg(r + 2);
    float u[kCarsTD], ip[kCarsTD], in[kCarsTD];
#pragma unroll
    for (int t = 0; t < kCarsTD; t++) {
      const int d = lane + 32 * t;
      u[t] = (d < D) ? __ldg(a.UI + (size_t)user * D + d) : 0.f;
      ip[t] = (d < D && MODE != CARS_QUERY) ? __ldg(a.UI + (size_t)item * D + d) : 0.f;
      in[t] = 0.f;
    }
    if (MODE == CARS_TRAIN)
      for (int j = 0; j < a.n_neg; j++) {
        const int neg = __ldg(r + 3 + j);
#pragma unroll
        for (int t = 0; t < kCarsTD; t++) {
          const int d = lane + 32 * t;
          if (d < D) in[t] += __ldg(a.UI + (size_t)neg * D + d);       // CARS2.py:90 reduce_sum over the negatives
        }
      }
    for (int c = lane; c < Dc; c += 32) sc[c] = __ldg(a.Ctx + (size_t)fea * Dc + c);
    __syncwarp();
    float Tc[kCarsTD], Sc[kCarsTD];
#pragma unroll
    for (int t = 0; t < kCarsTD; t++) {
      const int d = lane + 32 * t;
      float x = 0.f, y = 0.f;
      if (d < D)
        for (int c = 0; c < Dc; c++) {
          x = fmaf(sT[d * ldt + c], sc[c], x);
          if (MODE == CARS_SCORE) y = fmaf(sS[d * ldt + c], sc[c], y);
        }
      Tc[t] = x; Sc[t] = y;
    }
    if (MODE == CARS_QUERY) {
#pragma unroll
      for (int t = 0; t < kCarsTD; t++) { const int d = lane + 32 * t; if (d < D) a.out[s * D + d] = u[t] + Tc[t]; }
      __syncwarp();
      continue;
    }
    if (MODE == CARS_SCORE) {
      float part = 0.f;
#pragma unroll
      for (int t = 0; t < kCarsTD; t++) part += u[t] * ip[t] + u[t] * Sc[t] + ip[t] * Tc[t];
      part = warp_sum(part);
      if (lane == 0) a.out[s] = part;
      __syncwarp();
      continue;
    }
    // ---- TRAIN ----
    float dl[kCarsTD], part = 0.f;
#pragma unroll
    for (int t = 0; t < kCarsTD; t++) { dl[t] = ip[t] - in[t]; part += u[t] * dl[t] + dl[t] * Tc[t]; }
    const float x = warp_sum(part);
    const float sg = 1.f / (1.f + expf(-x));          // tf.sigmoid
    if (lane == 0) loss_acc += -logf(sg);             // CARS2.py:113
    const float g = sg - 1.f;
#pragma unroll
    for (int t = 0; t < kCarsTD; t++) {
      const int d = lane + 32 * t;
      if (d < D) {
        const float dd = g * (u[t] + Tc[t]);          // d delta
        atomicAdd(a.gUI + (size_t)user * D + d, g * dl[t]);
        atomicAdd(a.gUI + (size_t)item * D + d, dd);
        for (int j = 0; j < a.n_neg; j++) atomicAdd(a.gUI + (size_t)__ldg(r + 3 + j) * D + d, -dd);
        sd[d] = dl[t];
        a.Delta[s * D + d] = g * dl[t];
      }
    }
    __syncwarp();
    for (int c = lane; c < Dc; c += 32) {
      float dc = 0.f;
      for (int d = 0; d < D; d++) dc = fmaf(sT[d * ldt + c], sd[d], dc);
      atomicAdd(a.gCtx + (size_t)fea * Dc + c, g * dc);
      a.Cm[s * Dc + c] = sc[c];
    }
    __syncwarp();
  }
  if (MODE == CARS_TRAIN) {
    const float bl = block_sum(loss_acc, scratch);
    write_partial(a.loss_partials, bl);
  }
}

// one CTA per q: gZ[d,q,c] += B_q dT[d,c] ; gB_q += sum_{d,c} Z[d,q,c] dT[d,c]
__global__ void __launch_bounds__(256) cars2_expand_kernel(const float* __restrict__ Z, const float* __restrict__ Bv,
                                                           const float* __restrict__ dT, int D, int Dq, int Dc,
                                                           float* __restrict__ gZ, float* __restrict__ gB) {
  __shared__ float scratch[32];
  const int q = blockIdx.x;
  const float bq = __ldg(Bv + q);
  float acc = 0.f;
  for (int i = threadIdx.x; i < D * Dc; i += blockDim.x) {
    const int d = i / Dc, c = i % Dc;
    const size_t z = ((size_t)d * Dq + q) * Dc + c;
    const float t = __ldg(dT + i);
    gZ[z] += bq * t;
    acc = fmaf(__ldg(Z + z), t, acc);
  }
  const float tot = block_sum(acc, scratch);
  if (threadIdx.x == 0) gB[q] += tot;
}

static size_t cars_smem(int mode, int D, int Dc, int nw) {
  return ((size_t)(mode == CARS_SCORE ? 2 : 1) * D * (Dc + 1) + (size_t)nw * (D + Dc)) * sizeof(float);
}

template <int MODE>
static int launch_cars(const CarsArgs& a, cudaStream_t st) {
  const size_t smem = cars_smem(MODE, a.D, a.Dc, 8);
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(cars2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    set_error("cars2_kernel: cannot reserve %zu bytes of shared memory", smem);
    return HHFM_ERR_LAUNCH;
  }
  const int grid = grid_for(a.B, 8, 2);
  cars2_kernel<MODE><<<grid, 256, smem, st>>>(a);
  return check_launch("cars2_kernel");
}

static int cars_check(const int32_t* rec, int64_t B, int64_t stride, int n_neg, const float* params, int64_t n_ui, int64_t M,
                      int64_t D, int64_t Dp, int64_t Dq, int64_t Dc, const void* ws) {
  HHFM_REQUIRE(rec && params && ws, "cars2: NULL argument");
  HHFM_REQUIRE(B >= 0 && stride >= 3 + n_neg && n_neg >= 0 && n_neg <= kCarsMaxNeg, "cars2: bad record shape");
  HHFM_REQUIRE(n_ui > 0 && M > 0 && D >= 1 && D <= 32 * kCarsTD && Dc >= 1 && Dc <= 32 * kCarsTC && Dp >= 1 && Dq >= 1,
               "cars2: D <= %d, Dc <= %d required", 32 * kCarsTD, 32 * kCarsTC);
  return HHFM_OK;
}

struct CarsLayout {
  int64_t ui, ctx, w, z, a, b, total;      // parameter block
};
static CarsLayout cars_layout(int64_t n_ui, int64_t M, int64_t D, int64_t Dp, int64_t Dq, int64_t Dc) {
  CarsLayout l;
  int64_t o = 0;
  l.ui = o; o += n_ui * D;
  l.ctx = o; o += M * Dc;
  l.w = o; o += D * Dp * Dc;
  l.z = o; o += D * Dq * Dc;
  l.a = o; o += Dp;
  l.b = o; o += Dq;
  l.total = o;
  return l;
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int64_t hhfm_cars2_param_count(int64_t n_ui, int64_t M, int64_t D, int64_t Dp, int64_t Dq, int64_t Dc) {
  return cars_layout(n_ui, M, D, Dp, Dq, Dc).total;
}

extern "C" int64_t hhfm_workspace_bytes_cars2(int64_t B, int64_t D, int64_t Dc) {
  return (3 * D * Dc + B * (D + Dc) + 16) * (int64_t)sizeof(float);      // T, S, dT, Delta, Cm
}

// mode 0: out[B] = PositiveFeadback (records [user, item, fea]); mode 2: out[B, D] = u + T c (records [user, *, fea])
extern "C" int hhfm_cars2_fwd(const int32_t* rec, int64_t B, int64_t stride, const float* params, int64_t n_ui, int64_t M,
                              int64_t D, int64_t Dp, int64_t Dq, int64_t Dc, int32_t mode, float* out, float* workspace,
                              hhfm_stream_t stream) {
  int rc = cars_check(rec, B, stride, 0, params, n_ui, M, D, Dp, Dq, Dc, workspace);
  if (rc) return rc;
  HHFM_REQUIRE(out && (mode == CARS_SCORE || mode == CARS_QUERY), "cars2_fwd: out is NULL or bad mode");
  if (B == 0) return HHFM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const CarsLayout l = cars_layout(n_ui, M, D, Dp, Dq, Dc);
  float* T = workspace;
  float* S = workspace + D * Dc;
  cars2_prep_kernel<<<(unsigned)((D * Dc + 255) / 256), 256, 0, st>>>(params + l.w, params + l.z, params + l.a, params + l.b, (int)D,
                                                                      (int)Dp, (int)Dq, (int)Dc, T, S);
  if ((rc = check_launch("cars2_prep_kernel"))) return rc;
  CarsArgs a{};
  a.rec = rec; a.B = B; a.stride = stride; a.n_neg = 0; a.D = (int)D; a.Dc = (int)Dc;
  a.UI = params + l.ui; a.Ctx = params + l.ctx; a.T = T; a.S = S; a.out = out;
  return mode == CARS_SCORE ? launch_cars<CARS_SCORE>(a, st) : launch_cars<CARS_QUERY>(a, st);
}

// records [user, item, fea, neg_0 .. neg_{n_neg-1}, pad]; gradients ACCUMULATE into gparams (same layout as params);
// the lamda term is applied by the caller through hhfm_opt_*_dense_l2 over the whole block.
extern "C" int hhfm_cars2_fwd_bwd(const int32_t* rec, int64_t B, int64_t stride, int32_t n_neg, const float* params,
                                  int64_t n_ui, int64_t M, int64_t D, int64_t Dp, int64_t Dq, int64_t Dc, float* gparams,
                                  float* loss_partials, float* workspace, hhfm_stream_t stream) {
  int rc = cars_check(rec, B, stride, n_neg, params, n_ui, M, D, Dp, Dq, Dc, workspace);
  if (rc) return rc;
  HHFM_REQUIRE(B > 0 && n_neg >= 1 && gparams && loss_partials, "cars2_fwd_bwd: NULL argument or empty batch");
  cudaStream_t st = (cudaStream_t)stream;
  const CarsLayout l = cars_layout(n_ui, M, D, Dp, Dq, Dc);
  float* T = workspace;
  float* S = T + D * Dc;
  float* dT = S + D * Dc;
  float* Delta = dT + D * Dc;
  float* Cm = Delta + B * D;
  cars2_prep_kernel<<<(unsigned)((D * Dc + 255) / 256), 256, 0, st>>>(params + l.w, params + l.z, params + l.a, params + l.b, (int)D,
                                                                      (int)Dp, (int)Dq, (int)Dc, T, S);
  if ((rc = check_launch("cars2_prep_kernel"))) return rc;
  cudaMemsetAsync(dT, 0, (size_t)D * Dc * sizeof(float), st);
  CarsArgs a{};
  a.rec = rec; a.B = B; a.stride = stride; a.n_neg = n_neg; a.D = (int)D; a.Dc = (int)Dc;
  a.UI = params + l.ui; a.Ctx = params + l.ctx; a.T = T; a.S = S;
  a.gUI = gparams + l.ui; a.gCtx = gparams + l.ctx; a.Delta = Delta; a.Cm = Cm; a.loss_partials = loss_partials;
  if ((rc = launch_cars<CARS_TRAIN>(a, st))) return rc;
  if ((rc = sgemm_tn_splitk(Delta, D, Cm, Dc, (int)D, (int)Dc, (int)B, dT, Dc, st))) return rc;
  cars2_expand_kernel<<<(unsigned)Dq, 256, 0, st>>>(params + l.z, params + l.b, dT, (int)D, (int)Dq, (int)Dc, gparams + l.z,
                                                   gparams + l.b);
  return check_launch("cars2_expand_kernel");
}
