// DeepFM tower on the 5th-generation tensor cores with fp32-grade accuracy: 3xTF32 split GEMM on tcgen05.
//
// tcgen05 has no fp32-input MMA kind and the parity bar is 1e-5 relative, so every fp32 operand x is used as
//   hi = the tf32 the hardware sees (kind::tf32 ignores the low 13 mantissa bits of the fp32 word in shared memory),
//   lo = x - hi  (exact in fp32, materialised once per tensor by split_transpose_kernel),
// and  A.B ~= A_lo.B_hi + A_hi.B_lo + A_hi.B_hi  with fp32 accumulation in TMEM (the dropped lo.lo term and the tf32
// rounding of lo are ~2^-21 relative).  Three MMAs per k-step at tf32 rate (~1.1 PFLOP/s dense) is still ~5x the fp32
// CUDA-core peak (74 TFLOP/s).
//
// One kernel covers all nine GEMMs of a training step in "TN" form, C[M,N] = A[M,K] . B[N,K]^T, both operands K-major,
// staged by TMA with 128-byte swizzle (32 fp32 per row):
//   forward   H_{i+1} = relu(H_i . W_i + b_i)      A = H_i [B, d_i],            B = W_i^T [d_{i+1}, d_i]   EPI_BIAS_RELU
//   d input   dZ_i = (dZ_{i+1} . W_i^T) * relu'(H_i) A = dZ_{i+1} [B, d_{i+1}],  B = W_i [d_i, d_{i+1}]     EPI_MASK / EPI_SCATTER
//   d weight  dW_i = H_i^T . dZ_{i+1}               A = H_i^T [d_i, B],          B = dZ_{i+1}^T [d_{i+1}, B] EPI_ATOMIC (split-K)
// Transposed and lo copies of the activations come from split_transpose_kernel (one pass per activation tensor).
// Roles per CTA (320 threads): warp 4 = TMA producer, warp 5 = single-thread MMA issuer, warps 0-3 and 6-9 = epilogue
// (two per TMEM lane quarter, alternating 32-column chunks); 2 smem stages of {A_hi, A_lo, B_hi, B_lo}, 2 TMEM accumulators.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "dfm_tc.cuh"
#include "tc_common.cuh"

namespace hhfm {

constexpr int kTfM = 128;           // rows per CTA tile (UMMA M)
constexpr int kTfKC = 32;           // fp32 per 128-byte swizzle row
constexpr int kTfEpiWarps = 16;     // 4 per TMEM lane quarter
constexpr int kTfThreads = (2 + kTfEpiWarps) * 32;   // warps 0-3, 6-17: epilogue; warp 4: TMA; warp 5: MMA
constexpr int kTfMaxStages = 4;
constexpr int kTfABytes = kTfM * 128;            // one A tile (hi or lo)
constexpr int kTfSmemBudget = 216 * 1024;        // operand stages (the rest of the 227 KB: barriers + alignment slack)
constexpr int kTfSmemBytes = kTfSmemBudget + 1024 + 256;

struct TfKernelArgs {
  int M, N, K;
  int n_mtiles, n_ntiles, bn, splits, chunks_total, chunks_per_split, n_units;
  int drain_every;       // k-chunks accumulated in TMEM before the partial sum is added into fp32 registers
  int n_stages, stage_bytes;   // smem pipeline depth / bytes per stage for this bn
  int a_rpc, b_rpc;            // > 0: operands are k-blocked panels [k-chunk][rows][32]; rows per chunk panel of A / B
  int epi;
  float* C;
  int64_t ldc;
  const float* bias;
  const float* proj;
  const float* mask;
  int64_t ldmask;
  const int32_t* idx;
  int F, Kemb;
  HotPlan hot;
  int* err;
};

__global__ void __launch_bounds__(kTfThreads, 1) tf32x3_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                    const __grid_constant__ CUtensorMap tmAlo,
                                                                    const __grid_constant__ CUtensorMap tmB,
                                                                    const __grid_constant__ CUtensorMap tmBlo,
                                                                    const TfKernelArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTfSmemBudget);
  uint64_t* full = bars;                      // [stages] TMA -> MMA
  uint64_t* empty = bars + kTfMaxStages;      // [stages] MMA -> TMA
  uint64_t* t_full = empty + kTfMaxStages;    // [2] accumulator ready
  uint64_t* t_empty = t_full + 2;             // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 512;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmAlo); tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmBlo);
    for (int i = 0; i < kTfMaxStages; i++) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; i++) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, kTfEpiWarps * 32); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t b_bytes = (uint32_t)a.bn * 128u;
  const uint32_t stage_tx = 2u * kTfABytes + 2u * b_bytes;

  // unit u -> (n tile, m tile, k split), n tile fastest: the CTAs working at the same time share the A tile (activations,
  // streamed from HBM once) while the B tiles (weights) are small and stay in L2
  auto decode = [&](int u, int& mt, int& nt, int& sp) {
    nt = u % a.n_ntiles;
    const int r = u / a.n_ntiles;
    mt = r % a.n_mtiles;
    sp = r / a.n_mtiles;
  };

  if (warp == 4) {
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        int mt, nt, sp; decode(u, mt, nt, sp);
        const int c0 = sp * a.chunks_per_split, c1 = min(a.chunks_total, c0 + a.chunks_per_split);
        for (int c = c0; c < c1; c++) {
          mbar_wait(empty + st, ph ^ 1, a.err);
          mbar_expect_tx(full + st, stage_tx);
          uint8_t* sb = smem + st * a.stage_bytes;
          // row-major operands: tile = (columns of k-chunk c, rows of the tile); k-blocked panels: tile = rows of panel c
          const int ax = a.a_rpc ? 0 : c * kTfKC, ay = a.a_rpc ? c * a.a_rpc + mt * kTfM : mt * kTfM;
          const int bx = a.b_rpc ? 0 : c * kTfKC, by = a.b_rpc ? c * a.b_rpc + nt * a.bn : nt * a.bn;
          tma_load_2d(sb, &tmA, ax, ay, full + st);
          tma_load_2d(sb + kTfABytes, &tmAlo, ax, ay, full + st);
          tma_load_2d(sb + 2 * kTfABytes, &tmB, bx, by, full + st);
          tma_load_2d(sb + 2 * kTfABytes + b_bytes, &tmBlo, bx, by, full + st);
          if (++st == a.n_stages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(kTfM, a.bn);
      int st = 0, acc = 0; uint32_t ph = 0, tph = 0;
      for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        int mt, nt, sp; decode(u, mt, nt, sp);
        const int c0 = sp * a.chunks_per_split, c1 = min(a.chunks_total, c0 + a.chunks_per_split);
        for (int c = c0; c < c1; c++) {
          const bool first = ((c - c0) % a.drain_every) == 0;       // first k-chunk of a TMEM partial sum
          if (first) {
            mbar_wait(t_empty + acc, tph ^ 1, a.err);
            tc_fence_after();
          }
          const uint32_t d = tmem_base + (uint32_t)(acc * 256);
          mbar_wait(full + st, ph, a.err);
          tc_fence_after();
          const uint32_t sb = smem_u32(smem + st * a.stage_bytes);
          const uint32_t ah = sb, al = sb + kTfABytes, bh = sb + 2 * kTfABytes, bl = bh + b_bytes;
#pragma unroll
          for (int k4 = 0; k4 < 4; k4++) {          // UMMA_K = 8 tf32 = 32 bytes inside the 128-byte swizzle row
            const uint32_t o = k4 * 32;
            umma_tf32(d, make_sdesc(al + o), make_sdesc(bh + o), idesc, (!first || k4 != 0) ? 1u : 0u);
            umma_tf32(d, make_sdesc(ah + o), make_sdesc(bl + o), idesc, 1u);
            umma_tf32(d, make_sdesc(ah + o), make_sdesc(bh + o), idesc, 1u);
          }
          umma_commit(empty + st);
          if (++st == a.n_stages) { st = 0; ph ^= 1; }
          if (((c - c0) % a.drain_every) == a.drain_every - 1 || c == c1 - 1) {
            umma_commit(t_full + acc);              // partial sum complete: hand it to the epilogue warps
            if (++acc == 2) { acc = 0; tph ^= 1; }
          }
        }
      }
    }
  } else {
    // ---- epilogue: thread = (TMEM lane quarter, lane) owns output row m = mt*128 + quarter*32 + lane; the two warps of
    // a quarter take alternating 32-column chunks ----
    // epilogue warps are 0-3 and 6-17: warp & 3 is the TMEM lane quarter a warp may touch, `sub` its turn among the
    // four warps of the quarter (16-column chunk c belongs to the warp with c % 4 == sub)
    const int quarter = warp & 3;
    const int ew = warp < 4 ? warp : warp - 2;          // 0..15
    const int sub = ew >> 2;
    int acc = 0; uint32_t tph = 0;
    const int n_chunks = a.bn / 16;                     // 16-column chunks (bn is a multiple of 16)
    const int rep = a.hot.slot ? (int)(blockIdx.x % a.hot.n_rep) : 0;
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
      int mt, nt, sp; decode(u, mt, nt, sp);
      const int c0k = sp * a.chunks_per_split;
      const bool has_work = c0k < a.chunks_total;
      const int m = mt * kTfM + quarter * 32 + lane;
      // Two-level accumulation.  The tensor core adds into its fp32 accumulator with truncation, a bias that grows with
      // the number of accumulation steps (measured: 8e-3 absolute after K = 4096 on N(0,1) data).  So TMEM only holds the
      // partial sum of `drain_every` k-chunks; the partial sums are added here in registers with round-to-nearest.
      const int c1k = min(a.chunks_total, c0k + a.chunks_per_split);
      const int n_drains = has_work ? (c1k - c0k + a.drain_every - 1) / a.drain_every : 0;
      float accr[4][16];
      for (int dr = 0; dr < n_drains; dr++) {
        mbar_wait(t_full + acc, tph, a.err);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * 256);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = sub + 4 * j;
          if (c < n_chunks) {
            uint32_t r[16];
            tmem_ld16(taddr + c * 16, r);
            tmem_ld_wait_for16(r);
#pragma unroll
            for (int i = 0; i < 16; i++) accr[j][i] = (dr == 0) ? __uint_as_float(r[i]) : accr[j][i] + __uint_as_float(r[i]);
          }
        }
        tc_fence_before();
        mbar_arrive(t_empty + acc);
        if (++acc == 2) { acc = 0; tph ^= 1; }
      }
      if (a.epi == TF_EPI_SCATTER) {
        // d(H_0)[m, n] belongs to embedding row idx[m, n / Kemb], element n % Kemb.  A 16-column chunk lies inside one
        // field (Kemb % 16 == 0 is required by the host), so the id and hot-slot lookups of all four chunks are issued
        // together before the reductions (they were a chain of 32 dependent global loads per unit).
        float* dstp[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int n = nt * a.bn + (sub + 4 * j) * 16;
          dstp[j] = nullptr;
          if (sub + 4 * j < n_chunks && m < a.M && has_work && n < a.N) {
            const int f = n / a.Kemb;
            const int row = __ldg(a.idx + (int64_t)m * a.F + f);
            const int hs = a.hot.slot ? __ldg(a.hot.slot + row) : -1;
            dstp[j] = (hs >= 0 ? a.hot.ghot + ((size_t)rep * a.hot.n_hot + hs) * a.Kemb : a.C + (int64_t)row * a.Kemb) + (n - f * a.Kemb);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (dstp[j] == nullptr) continue;
#pragma unroll
          for (int i = 0; i < 16; i += 4) red_add_v4(dstp[j] + i, make_float4(accr[j][i], accr[j][i + 1], accr[j][i + 2], accr[j][i + 3]));
        }
        continue;
      }
      if (a.epi == TF_EPI_BIAS_RELU_PROJ) {
        // this thread's share of relu(row + bias) . proj over its 16-column chunks; one plain store per (row, n-tile, group)
        float p = 0.f;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int c = sub + 4 * j;
          if (c >= n_chunks) continue;
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const int n = nt * a.bn + c * 16 + i;
            if (n < a.N) p = fmaf(fmaxf(accr[j][i] + __ldg(a.bias + n), 0.f), __ldg(a.proj + n), p);
          }
        }
        if (m < a.M && has_work) a.C[(int64_t)m * a.ldc + nt * 4 + sub] = p;
        continue;
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int c = sub + 4 * j;
        if (c >= n_chunks) continue;
        const float (&r)[16] = accr[j];
        const int nb = nt * a.bn + c * 16;
        if (m >= a.M || !has_work) continue;
        float* crow = a.C + (int64_t)m * a.ldc;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const int n = nb + i;
          if (n >= a.N || n >= (nt + 1) * a.bn) continue;
          float v[4] = {r[i], r[i + 1], r[i + 2], r[i + 3]};
          const int nv = min(4, a.N - n);
          if (a.epi == TF_EPI_BIAS_RELU) {
#pragma unroll
            for (int q = 0; q < 4; q++) v[q] = (q < nv) ? fmaxf(v[q] + __ldg(a.bias + n + q), 0.f) : 0.f;
          }
          if (a.epi == TF_EPI_ATOMIC) {
            // split-K: this unit's partial tile goes to its own [128][bn] slot of the scratch buffer (plain stores);
            // reduce_partials_kernel adds the splits into C.  (One atomic per element and split made every weight-gradient
            // GEMM cost ~0.35 ms regardless of its size: millions of reductions on a few thousand addresses.)
            float* slot = a.C + ((int64_t)u * kTfM + (quarter * 32 + lane)) * a.bn + (c * 16 + i);
            *reinterpret_cast<float4*>(slot) = make_float4(v[0], v[1], v[2], v[3]);
          } else if (nv == 4) {
            *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
              if (q < nv) crow[n + q] = v[q];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kTmemCols);
}

// X [rows, ld] (cols valid) -> X_lo [rows, ld] and the K-BLOCKED transposes XT / XT_lo: panel p = rows [32p, 32p+32) of X
// stored as [cols][32] (element (r, c) at ((r/32)*cols + c)*32 + r%32), i.e. exactly the 128-byte-row tile a TMA box wants
// for the weight-gradient GEMM (K = the sample index), contiguous per k-chunk.  A plain [cols, rows] transpose made that
// GEMM fetch 128 bytes from each of ~600 rows half a megabyte apart per k-chunk (6x slower per chunk than the row-major
// GEMMs).  Rows beyond `rows` inside the last panel are written as zeros.  32x32 tiles through shared memory so both the
// reads and the writes are coalesced.  XT / XT_lo may be NULL.
// With `mask` (the forward activation H of the same shape): v = X * (mask > 0), i.e. the relu backward, and the masked value is
// also written to Xout (which may alias mask): the d-input GEMM stores the unmasked product and this pass finishes it
// with coalesced reads (a mask lookup in the GEMM epilogue was a chain of dependent global loads per tile).
// With `gidx` (ids [rows, gF]) X is the embedding table instead: v = X[gidx[r, c / gK], c % gK] (the flattened embeddings
// of layer 0), written to Xout as the hi operand.
__global__ void split_transpose_kernel(const float* X, const float* mask, float* Xout, int64_t rows, int cols, int64_t ld,
                                       float* Xlo, float* __restrict__ XT, float* __restrict__ XTlo,
                                       const int32_t* __restrict__ gidx, int gF, int gK) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int64_t r = r0 + ty + 8 * i;
    const int c = c0 + tx;
    float v = 0.f;
    if (r < rows && c < cols) {
      if (gidx) {
        const int f = c / gK;
        v = __ldg(X + (int64_t)__ldg(gidx + r * gF + f) * gK + (c - f * gK));
        Xout[r * ld + c] = v;
      } else {
        v = X[r * ld + c];
      }
      if (mask) {
        v = mask[r * ld + c] > 0.f ? v : 0.f;
        Xout[r * ld + c] = v;
      }
      if (Xlo) Xlo[r * ld + c] = tf32_lo(v);
    }
    tile[ty + 8 * i][tx] = v;
  }
  if (XT == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int c = c0 + ty + 8 * i;
    if (c < cols) {
      const float v = tile[tx][ty + 8 * i];                    // zero for rows >= `rows`
      const int64_t o = ((int64_t)blockIdx.x * cols + c) * 32 + tx;
      XT[o] = v;
      if (XTlo) XTlo[o] = tf32_lo(v);
    }
  }
}

// The same pass with 16-byte accesses: a CTA owns 32 rows x 128 columns; thread (row = t / 8, g = t % 8) reads the float4s
// g, g + 8, g + 16, g + 24 of its row (8 threads = 128 contiguous bytes), writes Xout / Xlo the same way, parks the values in a
// padded shared tile, and thread (column = t / 8 + 32 j, q = t % 8) then writes samples 4q .. 4q+3 of that column as one float4
// of the k-blocked transposed panels.  Requires ld % 4 == 0, 16-byte aligned blocks and (gather form) gK % 4 == 0.
__global__ void __launch_bounds__(256) split_transpose_vec_kernel(const float* X, const float* mask, float* Xout, int64_t rows, int cols,
                                                                  int64_t ld, float* Xlo, float* __restrict__ XT,
                                                                  float* __restrict__ XTlo, const int32_t* __restrict__ gidx, int gF,
                                                                  int gK) {
  __shared__ float tile[32][129];
  const int t = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 128;
  {
    const int rl = t >> 3, g = t & 7;
    const int64_t r = r0 + rl;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int cl = (g + 8 * j) * 4;
      const int c = c0 + cl;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < rows && c < cols) {            // cols <= ld and ld % 4 == 0: the float4 stays inside the row's padded block
        if (gidx) {
          const int f = c / gK;
          v = __ldg(reinterpret_cast<const float4*>(X + (int64_t)__ldg(gidx + r * gF + f) * gK + (c - f * gK)));
          *reinterpret_cast<float4*>(Xout + r * ld + c) = v;
        } else {
          v = *reinterpret_cast<const float4*>(X + r * ld + c);
        }
        if (mask) {
          const float4 m = *reinterpret_cast<const float4*>(mask + r * ld + c);
          v = make_float4(m.x > 0.f ? v.x : 0.f, m.y > 0.f ? v.y : 0.f, m.z > 0.f ? v.z : 0.f, m.w > 0.f ? v.w : 0.f);
          *reinterpret_cast<float4*>(Xout + r * ld + c) = v;
        }
        if (Xlo) *reinterpret_cast<float4*>(Xlo + r * ld + c) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
      }
      tile[rl][cl] = v.x; tile[rl][cl + 1] = v.y; tile[rl][cl + 2] = v.z; tile[rl][cl + 3] = v.w;
    }
  }
  if (XT == nullptr) return;
  __syncthreads();
  const int q = t & 7;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int cl = (t >> 3) + 32 * j;
    const int c = c0 + cl;
    if (c < cols) {
      const float4 v = make_float4(tile[4 * q][cl], tile[4 * q + 1][cl], tile[4 * q + 2][cl], tile[4 * q + 3][cl]);   // zero for rows >= `rows`
      const int64_t o = ((int64_t)blockIdx.x * cols + c) * 32 + 4 * q;
      *reinterpret_cast<float4*>(XT + o) = v;
      if (XTlo) *reinterpret_cast<float4*>(XTlo + o) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
    }
  }
}

static bool split_vec_ok(const void* a, const void* b, const void* c, const void* d, const void* e, const void* f, int64_t ld) {
  const char* env = getenv("HHFM_DFM_SPLIT_VEC");           // 0 = the 32 x 32 scalar tiles (A/B measurements)
  if (env && env[0] == '0') return false;
  const uintptr_t all = (uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d | (uintptr_t)e | (uintptr_t)f;
  return (ld & 3) == 0 && (all & 15) == 0;
}

// X0[b, f*K + k] = V[idx[b, f], k]: the flattened embeddings as a dense [B, F*K] matrix (the TMA operand of layer 0)
__global__ void __launch_bounds__(256) gather_x0_kernel(const int32_t* __restrict__ idx, int64_t B, int F, int K,
                                                        const float* __restrict__ V, float* __restrict__ X0, int64_t ld) {
  const int kv = K >> 2;
  const int64_t total = B * F * kv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % kv);
    const int64_t bf = i / kv;
    const int f = (int)(bf % F);
    const int64_t b = bf / F;
    const int row = __ldg(idx + b * F + f);
    reinterpret_cast<float4*>(X0 + b * ld + (int64_t)f * K)[c] = __ldg(reinterpret_cast<const float4*>(V + (int64_t)row * K) + c);
  }
}

// column sums of X [rows, ld] (cols valid) accumulated into out[cols] (bias gradients)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int64_t rows, int cols, int64_t ld,
                                                     float* __restrict__ out) {
  __shared__ float s[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.y * 32 + tx;
  float acc = 0.f;
  if (c < cols)
    for (int64_t r = (int64_t)blockIdx.x * 8 + ty; r < rows; r += (int64_t)gridDim.x * 8) acc += X[r * ld + c];
  s[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += s[i][tx];
    atomicAdd(out + c, t);
  }
}

// W [rows, cols] row-major (ld = cols) -> Wp [rows, ldp] (+lo) and WT [cols, ldtp] (+lo), zero padded
__global__ void __launch_bounds__(256) prep_weight_kernel(const float* __restrict__ W, int rows, int cols, float* __restrict__ Wp,
                                                          float* __restrict__ Wplo, int ldp, float* __restrict__ WT,
                                                          float* __restrict__ WTlo, int ldtp) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float v = W[i], lo = tf32_lo(v);
    Wp[(int64_t)r * ldp + c] = v; Wplo[(int64_t)r * ldp + c] = lo;
    WT[(int64_t)c * ldtp + r] = v; WTlo[(int64_t)c * ldtp + r] = lo;
  }
}

// C[m, n] += sum over splits of the partial tiles written by the split-K epilogue (unit u = nt + n_ntiles*(mt + n_mtiles*sp))
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ scratch, int n_mtiles, int n_ntiles,
                                                              int splits, int bn, int M, int N, float* __restrict__ C, int64_t ldc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * N) return;
  const int m = (int)(i / N), n = (int)(i % N);
  const int mt = m / kTfM, nt = n / bn;
  float acc = 0.f;
  for (int sp = 0; sp < splits; sp++) {
    const int64_t u = nt + (int64_t)n_ntiles * (mt + (int64_t)n_mtiles * sp);
    acc += scratch[(u * kTfM + (m - mt * kTfM)) * bn + (n - nt * bn)];
  }
  C[(int64_t)m * ldc + n] += acc;
}

static int make_tmap_f32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return HHFM_ERR_LAUNCH;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)kTfKC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(f32 %lld x %lld, ld %lld) failed (%d)", (long long)rows, (long long)cols, (long long)ld, (int)r);
    return HHFM_ERR_LAUNCH;
  }
  return HHFM_OK;
}

int64_t tf_splitk_scratch_floats() { return (int64_t)(sm_count() + 8) * kTfM * 256; }

int tf_pick_bn(int N) {
  // N-tile width: a multiple of 16, at most 256 (four 16-column chunks per epilogue thread), least padded work first
  if (N <= 256) return (N + 15) / 16 * 16;
  int best = 256, best_cost = 1 << 30;
  for (int bn = 256; bn >= 64; bn -= 16) {
    const int cost = (N + bn - 1) / bn * bn;
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

int tf_proj_partials(int N) { return 4 * ((N + tf_pick_bn(N) - 1) / tf_pick_bn(N)); }

int tf_gemm(const TfGemm& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return HHFM_OK;
  HHFM_REQUIRE(g.k_blocked || ((g.lda % 4) == 0 && (g.ldb % 4) == 0), "tf_gemm: operand leading dimensions must be multiples of 4 floats");
  HHFM_REQUIRE((((uintptr_t)g.A | (uintptr_t)g.A_lo | (uintptr_t)g.B | (uintptr_t)g.B_lo) & 15) == 0, "tf_gemm: operands must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(tf32x3_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTfSmemBytes) != cudaSuccess) {
      set_error("tf32x3_gemm_kernel: cannot reserve %d bytes of shared memory", kTfSmemBytes);
      return HHFM_ERR_LAUNCH;
    }
    attr_set = true;
  }
  TfKernelArgs a{};
  a.M = g.M; a.N = g.N; a.K = g.K;
  a.bn = tf_pick_bn(g.N);
  a.n_mtiles = (g.M + kTfM - 1) / kTfM;
  a.n_ntiles = (g.N + a.bn - 1) / a.bn;
  a.chunks_total = (g.K + kTfKC - 1) / kTfKC;
  a.splits = 1;
  if (g.epi == TF_EPI_ATOMIC) {
    const int tiles = a.n_mtiles * a.n_ntiles;
    int splits = sm_count() / tiles;                       // one wave: every split ends in tile-sized atomics, keep them few
    const int max_splits = (a.chunks_total + 15) / 16;     // at least 16 k-chunks (512 samples) per split
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    a.splits = splits;
  }
  a.chunks_per_split = (a.chunks_total + a.splits - 1) / a.splits;
  a.splits = (a.chunks_total + a.chunks_per_split - 1) / a.chunks_per_split;
  a.n_units = a.n_mtiles * a.n_ntiles * a.splits;
  {
    // k-chunks (32 k-elements = 12 MMAs each) per TMEM partial sum: 1 gives < 4e-6 relative error; larger values trade
    // accuracy (the truncation bias grows with the number of accumulation steps) for less epilogue work
    const char* e = getenv("HHFM_TF_DRAIN");
    a.drain_every = e ? atoi(e) : 1;
    if (a.drain_every < 1) a.drain_every = 1;
  }
  a.stage_bytes = 2 * kTfABytes + 2 * a.bn * 128;
  a.n_stages = kTfSmemBudget / a.stage_bytes;
  if (a.n_stages > kTfMaxStages) a.n_stages = kTfMaxStages;
  a.epi = g.epi; a.C = g.C; a.ldc = g.ldc; a.bias = g.bias; a.proj = g.proj; a.mask = g.mask; a.ldmask = g.ldmask;
  HHFM_REQUIRE(g.epi != TF_EPI_BIAS_RELU_PROJ || (g.proj && g.bias && g.ldc >= 4 * a.n_ntiles),
               "tf_gemm: the projection epilogue needs bias, proj and ldc >= 4 * n-tiles");
  if (g.epi == TF_EPI_ATOMIC) {
    HHFM_REQUIRE(g.scratch && ((uintptr_t)g.scratch & 15) == 0, "tf_gemm: split-K needs a 16-byte aligned scratch buffer");
    HHFM_REQUIRE((int64_t)a.n_units * kTfM * a.bn <= g.scratch_floats, "tf_gemm: split-K scratch too small");
    a.C = g.scratch;
  }
  a.idx = g.idx; a.F = g.F; a.Kemb = g.Kemb; a.hot = g.hot; a.err = nullptr;
  HHFM_REQUIRE(g.epi != TF_EPI_MASK, "tf_gemm: the relu mask is applied by tf_split_transpose, not by the GEMM epilogue");
  HHFM_REQUIRE(g.epi != TF_EPI_SCATTER || g.Kemb % 16 == 0, "tf_gemm: the scatter epilogue needs an embedding size that is a multiple of 16");
  HHFM_REQUIRE(g.epi == TF_EPI_SCATTER || g.epi == TF_EPI_ATOMIC || g.epi == TF_EPI_BIAS_RELU_PROJ ||
                   ((g.ldc % 4) == 0 && ((uintptr_t)g.C & 15) == 0),
               "tf_gemm: C must be 16-byte aligned with ldc %% 4 == 0");
  CUtensorMap tA, tAl, tB, tBl;
  int rc;
  if (g.k_blocked) {
    // operands are panels [ceil(K/32)][rows][32]: a 2-D tensor of 128-byte rows; a tile that runs past its panel reads the
    // next panel's rows, which only feeds output rows / columns beyond M / N (discarded by the epilogue)
    const int64_t panels = (g.K + kTfKC - 1) / kTfKC;
    a.a_rpc = g.M; a.b_rpc = g.N;
    if ((rc = make_tmap_f32(&tA, g.A, panels * g.M, kTfKC, kTfKC, kTfM))) return rc;
    if ((rc = make_tmap_f32(&tAl, g.A_lo, panels * g.M, kTfKC, kTfKC, kTfM))) return rc;
    if ((rc = make_tmap_f32(&tB, g.B, panels * g.N, kTfKC, kTfKC, a.bn))) return rc;
    if ((rc = make_tmap_f32(&tBl, g.B_lo, panels * g.N, kTfKC, kTfKC, a.bn))) return rc;
  } else {
    if ((rc = make_tmap_f32(&tA, g.A, g.M, g.K, g.lda, kTfM))) return rc;
    if ((rc = make_tmap_f32(&tAl, g.A_lo, g.M, g.K, g.lda, kTfM))) return rc;
    if ((rc = make_tmap_f32(&tB, g.B, g.N, g.K, g.ldb, a.bn))) return rc;
    if ((rc = make_tmap_f32(&tBl, g.B_lo, g.N, g.K, g.ldb, a.bn))) return rc;
  }
  const int grid = a.n_units < sm_count() ? a.n_units : sm_count();
  tf32x3_gemm_kernel<<<grid, kTfThreads, kTfSmemBytes, st>>>(tA, tAl, tB, tBl, a);
  if ((rc = check_launch("tf32x3_gemm_kernel"))) return rc;
  if (g.epi == TF_EPI_ATOMIC) {
    const int64_t n = (int64_t)g.M * g.N;
    reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g.scratch, a.n_mtiles, a.n_ntiles, a.splits, a.bn, g.M, g.N,
                                                                       g.C, g.ldc);
    rc = check_launch("reduce_partials_kernel");
  }
  return rc;
}

// inference form of the split: no transposes, no mask -- a float4 sweep over the padded [rows, ld] block (1.43 -> ~0.45 ms per
// 2.1 M x 152 activations; the tiled kernel above is built around the transposes the backward needs)
__global__ void __launch_bounds__(256) split_lo_vec_kernel(const float4* __restrict__ X, float4* __restrict__ Xlo, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = X[i];
    Xlo[i] = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  }
}

int tf_split_transpose(const float* X, int64_t rows, int cols, int64_t ld, float* Xlo, float* XT, float* XTlo, cudaStream_t st,
                       const float* mask, float* Xout) {
  if (rows <= 0 || cols <= 0) return HHFM_OK;
  if (XT == nullptr && XTlo == nullptr && mask == nullptr && Xlo != nullptr && (ld & 3) == 0 &&
      (((uintptr_t)X | (uintptr_t)Xlo) & 15) == 0) {
    const int64_t n4 = rows * (ld >> 2);
    const int64_t need = (n4 + 255) / 256, cap = (int64_t)sm_count() * 16;
    split_lo_vec_kernel<<<(unsigned)(need < cap ? need : cap), 256, 0, st>>>(reinterpret_cast<const float4*>(X),
                                                                            reinterpret_cast<float4*>(Xlo), n4);
    return check_launch("split_lo_vec_kernel");
  }
  if (split_vec_ok(X, mask, Xout, Xlo, XT, XTlo, ld)) {
    dim3 gridv((unsigned)((rows + 31) / 32), (unsigned)((cols + 127) / 128));
    split_transpose_vec_kernel<<<gridv, 256, 0, st>>>(X, mask, Xout, rows, cols, ld, Xlo, XT, XTlo, nullptr, 0, 0);
    return check_launch("split_transpose_vec_kernel");
  }
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  split_transpose_kernel<<<grid, 256, 0, st>>>(X, mask, Xout, rows, cols, ld, Xlo, XT, XTlo, nullptr, 0, 0);
  return check_launch("split_transpose_kernel");
}

// X0 = flattened embeddings V[idx] [B, F*K] with its lo part and (optionally) the k-blocked transposes, in one pass
int tf_gather_split_transpose(const int32_t* idx, int64_t B, int F, int K, const float* V, float* X0, int64_t ld, float* Xlo,
                              float* XT, float* XTlo, cudaStream_t st) {
  if ((K & 3) == 0 && split_vec_ok(V, nullptr, X0, Xlo, XT, XTlo, ld)) {
    dim3 gridv((unsigned)((B + 31) / 32), (unsigned)((F * K + 127) / 128));
    split_transpose_vec_kernel<<<gridv, 256, 0, st>>>(V, nullptr, X0, B, F * K, ld, Xlo, XT, XTlo, idx, F, K);
    return check_launch("split_transpose_vec_kernel(gather)");
  }
  dim3 grid((unsigned)((B + 31) / 32), (unsigned)((F * K + 31) / 32));
  split_transpose_kernel<<<grid, 256, 0, st>>>(V, nullptr, X0, B, F * K, ld, Xlo, XT, XTlo, idx, F, K);
  return check_launch("split_transpose_kernel(gather)");
}

int tf_gather_x0(const int32_t* idx, int64_t B, int F, int K, const float* V, float* X0, int64_t ld, cudaStream_t st) {
  int64_t blocks = (B * F * (K >> 2) + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  gather_x0_kernel<<<(unsigned)blocks, 256, 0, st>>>(idx, B, F, K, V, X0, ld);
  return check_launch("gather_x0_kernel");
}

int tf_colsum(const float* X, int64_t rows, int cols, int64_t ld, float* out, cudaStream_t st) {
  int64_t bx = (rows + 8 * 64 - 1) / (8 * 64);
  if (bx > 2 * sm_count()) bx = 2 * sm_count();
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)((cols + 31) / 32));
  colsum_kernel<<<grid, 256, 0, st>>>(X, rows, cols, ld, out);
  return check_launch("colsum_kernel");
}

int tf_prep_weight(const float* W, int rows, int cols, float* Wp, float* Wplo, int ldp, float* WT, float* WTlo, int ldtp,
                   cudaStream_t st) {
  const int64_t n = (int64_t)rows * cols;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  prep_weight_kernel<<<(unsigned)blocks, 256, 0, st>>>(W, rows, cols, Wp, Wplo, ldp, WT, WTlo, ldtp);
  return check_launch("prep_weight_kernel");
}

}  // namespace hhfm

using namespace hhfm;

// C[M,N] = A[M,K] . B[N,K]^T with fp32-grade accuracy on the tensor cores (3xTF32 split, see the file header).
// workspace: (M*lda + N*ldb) floats for the lo parts.  lda, ldb, ldc multiples of 4; all pointers 16-byte aligned.
extern "C" int hhfm_gemm_tn_tf32x3(const float* A, int64_t lda, const float* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                                   float* C, int64_t ldc, float* workspace, hhfm_stream_t stream) {
  HHFM_REQUIRE(A && B && C && workspace, "gemm_tn_tf32x3: NULL argument");
  HHFM_REQUIRE(M > 0 && N > 0 && K > 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm_tn_tf32x3: bad sizes");
  HHFM_REQUIRE(lda >= K && ldb >= K && ldc >= N, "gemm_tn_tf32x3: leading dimensions too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* Alo = workspace;
  float* Blo = workspace + M * lda;
  int rc;
  if ((rc = tf_split_transpose(A, M, (int)K, lda, Alo, nullptr, nullptr, st))) return rc;
  if ((rc = tf_split_transpose(B, N, (int)K, ldb, Blo, nullptr, nullptr, st))) return rc;
  TfGemm g{};
  g.A = A; g.A_lo = Alo; g.B = B; g.B_lo = Blo; g.M = (int)M; g.N = (int)N; g.K = (int)K; g.lda = lda; g.ldb = ldb;
  g.epi = TF_EPI_STORE; g.C = C; g.ldc = ldc;
  return tf_gemm(g, st);
}
