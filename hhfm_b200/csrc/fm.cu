// K1: FM / MF gather + second-order interaction, fused with the squared loss and the backward scatter.
// Reference graph: Newcode/FM.py:99-126 (forward + loss), TF autodiff for the backward; MF.py:81-98.
//
// Mapping: one sample is owned by a group of LPS = K/4 lanes (an aligned sub-warp), each lane holding one
// float4 of every embedding row -> every gathered row is one fully coalesced K*4-byte read, the reduction
// over fields is per-lane register work, and only the final sum over k needs log2(LPS) shuffles.
// The backward re-reads the rows (L1/L2 hits) and issues one REDG.E.ADD.F32x4 per 16 bytes.
#include "common.cuh"

namespace hhfm {

struct FmArgs {
  const int32_t* row_ptr;
  const int32_t* col;
  const float* val;
  int64_t B;
  int F;
  const float* V;
  const float* bias;
  const float* b0;
  int K;
  int interaction;
  const float* labels;
  const float* gout;
  float* out;
  float* gV;
  float* gbias;
  float* gb0;
  float* loss_partials;
  int32_t* touch_stamp;
  int32_t stamp;
  int32_t* touched_rows;
  int32_t* touched_count;
  int groups_active;
  HotPlan hot;
};

enum { FM_FWD = 0, FM_TRAIN = 1, FM_BWD = 2 };

template <int LPS, int VPL, int MODE>
__global__ void __launch_bounds__(kBlock) fm_kernel(const FmArgs a) {
  __shared__ float scratch[32];
  using F4 = Frag<LPS, VPL>;
  const int lane = threadIdx.x & 31, lg = lane % LPS, grp = lane / LPS;
  const int64_t warp_g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int ga = a.groups_active;
  const int K = a.K;
  const float b0 = a.b0 ? __ldg(a.b0) : 0.f;
  float loss_acc = 0.f, g0_acc = 0.f;
  const int rep = a.hot.slot ? (int)((warp_g * (32 / LPS) + grp) % a.hot.n_rep) : 0;

  for (int64_t s0 = warp_g * ga; s0 < a.B; s0 += n_warps * ga) {
    const int64_t s = s0 + grp;
    const bool valid = (grp < ga) && (s < a.B);
    int64_t beg = 0, end = 0;
    if (valid) {
      if (a.row_ptr) {
        beg = __ldg(a.row_ptr + s);
        end = __ldg(a.row_ptr + s + 1);
      } else {
        beg = s * a.F;
        end = beg + a.F;
      }
    }
    F4 S, Q, e0, e1;
    frag_zero(S);
    frag_zero(Q);
    frag_zero(e0);
    frag_zero(e1);
    float bsum = 0.f, part = 0.f;

    if (a.interaction == 0) {
      // ---- FM.py:99-109: S = sum_f e_f, Q = sum_f e_f^2 (fields in order, 4 gathers in flight) ----
      for (int64_t j = beg; j < end; j += 4) {
        int id[4];
        float vv[4];
        F4 e[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const bool ok = (j + u) < end;
          id[u] = ok ? __ldg(a.col + j + u) : -1;
          vv[u] = (ok && a.val) ? __ldg(a.val + j + u) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (id[u] >= 0) frag_load(e[u], a.V, id[u], K, lg);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (id[u] >= 0) {
            if (a.bias) bsum += vv[u] * __ldg(a.bias + id[u]);
#pragma unroll
            for (int i = 0; i < VPL; i++) {
              float4 x = a.val ? f4_scale(e[u].v[i], vv[u]) : e[u].v[i];
              S.v[i] = f4_add(S.v[i], x);
              Q.v[i] = f4_add(Q.v[i], f4_mul(x, x));
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < VPL; i++) {
        float4 t = f4_sub(f4_mul(S.v[i], S.v[i]), Q.v[i]);
        part += 0.5f * f4_hsum(t);
      }
    } else {
      // ---- MF.py:81-92: out = sum_k V[x0]*V[x1]; the bias term is not added (:92) ----
      if (valid) {
        frag_load(e0, a.V, __ldg(a.col + beg), K, lg);
        frag_load(e1, a.V, __ldg(a.col + beg + 1), K, lg);
      }
      part = frag_dot(e0, e1);
    }
    const float bil = group_sum<LPS>(part);
    const float out = (bil + bsum) + b0;   // FM.py:120 add_n([Bilinear, Feature_bias, Bias])

    if (MODE == FM_FWD) {
      if (valid && lg == 0) a.out[s] = out;
    }
    float g = 0.f;
    if (MODE == FM_TRAIN) {
      const float y = valid ? __ldg(a.labels + s) : out;
      const float diff = y - out;
      g = -diff;                            // d(0.5*diff^2)/d out
      if (valid && lg == 0) {
        loss_acc += 0.5f * diff * diff;     // FM.py:124 tf.nn.l2_loss
        if (a.out) a.out[s] = out;
      }
    } else if (MODE == FM_BWD) {
      g = valid ? __ldg(a.gout + s) : 0.f;
    }
    if (MODE == FM_FWD || !valid) {
      // nothing to scatter (forward-only launch, or a padding group of the last warp)
    } else {
    if (lg == 0) g0_acc += g;

    // ---- backward: gV[x_f] += g*val_f*(S - e_f); gbias[x_f] += g*val_f ----
    if (a.interaction == 0) {
      for (int64_t j = beg; j < end; j += 4) {
        int id[4];
        float vv[4];
        F4 e[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const bool ok = (j + u) < end;
          id[u] = ok ? __ldg(a.col + j + u) : -1;
          vv[u] = (ok && a.val) ? __ldg(a.val + j + u) : 1.f;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (id[u] >= 0) frag_load(e[u], a.V, id[u], K, lg);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (id[u] >= 0) {
            const float gv = g * vv[u];
            F4 d;
#pragma unroll
            for (int i = 0; i < VPL; i++) {
              float4 x = a.val ? f4_scale(e[u].v[i], vv[u]) : e[u].v[i];
              d.v[i] = f4_scale(f4_sub(S.v[i], x), gv);
            }
            scatter_row<LPS, VPL>(a.gV, a.hot, rep, id[u], K, lg, d);
            if (lg == 0) {
              if (a.gbias) scatter_bias(a.gbias, a.hot, rep, id[u], gv);
              touch_row(a.touch_stamp, a.stamp, a.touched_rows, a.touched_count, id[u]);
            }
          }
        }
      }
    } else {
      const int x0 = __ldg(a.col + beg), x1 = __ldg(a.col + beg + 1);
      F4 d0, d1;
#pragma unroll
      for (int i = 0; i < VPL; i++) {
        d0.v[i] = f4_scale(e1.v[i], g);
        d1.v[i] = f4_scale(e0.v[i], g);
      }
      scatter_row<LPS, VPL>(a.gV, a.hot, rep, x0, K, lg, d0);
      scatter_row<LPS, VPL>(a.gV, a.hot, rep, x1, K, lg, d1);
      if (lg == 0) {
        touch_row(a.touch_stamp, a.stamp, a.touched_rows, a.touched_count, x0);
        touch_row(a.touch_stamp, a.stamp, a.touched_rows, a.touched_count, x1);
      }
    }
    }  // valid && MODE != FM_FWD
  }

  if (MODE == FM_TRAIN) {
    const float bl = block_sum(loss_acc, scratch);
    write_partial(a.loss_partials, bl);
  }
  if (MODE != FM_FWD && a.gb0 != nullptr) {
    const float bg = block_sum(g0_acc, scratch);
    if (threadIdx.x == 0 && bg != 0.f) atomicAdd(a.gb0, bg);
  }
}

template <int LPS, int VPL, int MODE>
static int launch_fm(const FmArgs& a, int deterministic, cudaStream_t st) {
  static int occ = 0;
  if (occ == 0) {
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fm_kernel<LPS, VPL, MODE>, kBlock, 0);
    if (occ < 1) occ = 1;
  }
  FmArgs b = a;
  constexpr int G = 32 / LPS;
  if (deterministic) {
    b.groups_active = 1;
    fm_kernel<LPS, VPL, MODE><<<1, 32, 0, st>>>(b);
  } else {
    b.groups_active = G;
    const int grid = grid_for(a.B, (kBlock / 32) * G, occ);
    fm_kernel<LPS, VPL, MODE><<<grid, kBlock, 0, st>>>(b);
  }
  return check_launch("fm_kernel");
}

template <int MODE>
static int dispatch_fm(const FmArgs& a, int deterministic, cudaStream_t st) {
#define CALL(L, V) return launch_fm<L, V, MODE>(a, deterministic, st)
  HHFM_DISPATCH_K(a.K, CALL);
#undef CALL
  return HHFM_ERR_UNSUPPORTED;
}

static int check_common(int64_t B, int64_t F, const void* col, const void* V, int64_t M, int64_t K, int interaction,
                        const void* row_ptr) {
  HHFM_REQUIRE(B >= 0 && M > 0, "fm: bad sizes B=%lld M=%lld", (long long)B, (long long)M);
  HHFM_REQUIRE(col != nullptr && V != nullptr, "fm: col and V must not be NULL");
  HHFM_REQUIRE(K > 0 && K % 4 == 0 && K <= 512, "fm: K=%lld unsupported (need K %% 4 == 0, K <= 512)", (long long)K);
  HHFM_REQUIRE(row_ptr != nullptr || F > 0, "fm: fixed-width batch needs F > 0");
  HHFM_REQUIRE(interaction == 0 || interaction == 1, "fm: interaction must be 0 (FM) or 1 (MF)");
  HHFM_REQUIRE(interaction == 0 || (row_ptr == nullptr && F >= 2), "fm: MF interaction needs fixed width F >= 2");
  HHFM_REQUIRE(((uintptr_t)V & 15) == 0, "fm: V must be 16-byte aligned");
  return HHFM_OK;
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int hhfm_fm_fwd(const int32_t* row_ptr, const int32_t* col, const float* val, int64_t B, int64_t F,
                           const float* V, const float* bias, const float* b0, int64_t M, int64_t K,
                           int32_t interaction, float* out, hhfm_stream_t stream) {
  int rc = check_common(B, F, col, V, M, K, interaction, row_ptr);
  if (rc) return rc;
  HHFM_REQUIRE(out != nullptr, "fm_fwd: out is NULL");
  if (B == 0) return HHFM_OK;
  FmArgs a{};
  a.row_ptr = row_ptr; a.col = col; a.val = val; a.B = B; a.F = (int)F; a.V = V; a.bias = bias; a.b0 = b0;
  a.K = (int)K; a.interaction = interaction; a.out = out;
  return dispatch_fm<FM_FWD>(a, 0, (cudaStream_t)stream);
}

extern "C" int hhfm_fm_fwd_bwd_sqloss(const int32_t* row_ptr, const int32_t* col, const float* val, int64_t B,
                                      int64_t F, const float* V, const float* bias, const float* b0, int64_t M,
                                      int64_t K, int32_t interaction, const float* labels, float* out, float* gV,
                                      float* gbias, float* gb0, float* loss_partials, int32_t* touch_stamp,
                                      int32_t stamp, int32_t* touched_rows, int32_t* touched_count,
                                      const int32_t* hot_slot, float* ghot, float* ghot_bias, int32_t n_rep,
                                      int32_t n_hot, int32_t deterministic, hhfm_stream_t stream) {
  int rc = check_common(B, F, col, V, M, K, interaction, row_ptr);
  if (rc) return rc;
  HHFM_REQUIRE(!hot_slot || (ghot && n_rep >= 1 && n_hot >= 1), "fm_fwd_bwd_sqloss: hot_slot needs ghot, n_rep, n_hot");
  HHFM_REQUIRE(labels && gV && loss_partials, "fm_fwd_bwd_sqloss: labels, gV and loss_partials are required");
  HHFM_REQUIRE(B > 0, "fm_fwd_bwd_sqloss: empty batch");
  HHFM_REQUIRE(!touch_stamp || (touched_rows && touched_count), "fm_fwd_bwd_sqloss: touch_stamp needs touched_rows/count");
  HHFM_REQUIRE(((uintptr_t)gV & 15) == 0, "fm: gV must be 16-byte aligned");
  FmArgs a{};
  a.row_ptr = row_ptr; a.col = col; a.val = val; a.B = B; a.F = (int)F; a.V = V; a.bias = bias; a.b0 = b0;
  a.K = (int)K; a.interaction = interaction; a.labels = labels; a.out = out; a.gV = gV; a.gbias = gbias; a.gb0 = gb0;
  a.loss_partials = loss_partials; a.touch_stamp = touch_stamp; a.stamp = stamp; a.touched_rows = touched_rows;
  a.touched_count = touched_count;
  a.hot = HotPlan{hot_slot, ghot, ghot_bias, n_rep, n_hot};
  return dispatch_fm<FM_TRAIN>(a, deterministic, (cudaStream_t)stream);
}

extern "C" int hhfm_fm_bwd(const int32_t* row_ptr, const int32_t* col, const float* val, int64_t B, int64_t F,
                           const float* V, int64_t M, int64_t K, int32_t interaction, const float* gout, float* gV,
                           float* gbias, float* gb0, int32_t deterministic, hhfm_stream_t stream) {
  int rc = check_common(B, F, col, V, M, K, interaction, row_ptr);
  if (rc) return rc;
  HHFM_REQUIRE(gout && gV, "fm_bwd: gout and gV are required");
  if (B == 0) return HHFM_OK;
  FmArgs a{};
  a.row_ptr = row_ptr; a.col = col; a.val = val; a.B = B; a.F = (int)F; a.V = V; a.K = (int)K;
  a.interaction = interaction; a.gout = gout; a.gV = gV; a.gbias = gbias; a.gb0 = gb0;
  return dispatch_fm<FM_BWD>(a, deterministic, (cudaStream_t)stream);
}
