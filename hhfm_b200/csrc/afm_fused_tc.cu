// K2 on the 5th-generation tensor cores: the AFM training pass (Newcode/AFM.py:103-148 + autodiff) as ONE kernel.
//
// The three matrix products of a sample -- Z = P W (attention logits), dP = dZ W^T, dW += P^T dZ, 2*P*K*A flops each --
// run on tcgen05 as 3xTF32 split products (fp32-grade accuracy, see dfm_tc.cu); everything around them (pair products,
// relu, softmax over the pairs, the weighted sum, the chain rule, the embedding scatter) runs in the same CTA between
// the MMA phases, so the [pairs, K] tensors never leave shared memory / TMEM.
//
// Tile = 2 samples = 128 operand rows: sample slot ss owns rows [64 ss, 64 ss + P), the rows P..63 of a slot stay zero.
//   P operand   [128 rows][K = 64]   written by the threads as (x, lo(x)) in the 128-byte-swizzled K-major tile layout
//                                    (two 32-column tiles per part); the SAME buffer is the MN-major operand of dW.
//   GEMM 1      Z[128, 64]   = P . W          A = P (K-major),  B = W^T tiles (K-major)        -> TMEM columns [0, 64)
//   epilogue 1  thread = (row, column half): Z row from TMEM, relu, logits, softmax over the slot's pairs, afm, out,
//               loss, d a, d s, dZ = ds * p * relu'  -> dZ operand tiles (x, lo) ; column sums -> d p, d b
//   GEMM 2      dP[128, 64]  = dZ . W^T       A = dZ (K-major), B = W tiles (K-major)          -> TMEM columns [64, 128)
//   GEMM 3      D[128, 64]  += [P ; P_lo]^T . (dZ + dZ_lo)   both operands MN-major views of the tiles above, reduction over
//               the 128 tile rows; rows 0..63 of D hold P^T dZ, rows 64..127 hold P_lo^T dZ      -> TMEM columns [128, 192)
//   epilogue 2  dP row from TMEM (+ a_p * d afm) -> shared memory; thread = (slot, field, float4): dE_f = sum_j dP_(f,j) * E_j,
//               one vector reduction per 16 bytes into the gradient table (hot-row replicas as in K1 / K3)
// dW stays in TMEM for kDrain tiles, then the partial sum is added into a shared-memory accumulator (two-level
// accumulation: the tensor core adds into its accumulator with truncation, see dfm_tc.cu); the CTA adds its accumulator
// into the global gradient once at the end.
//
// Shapes covered: K == A == 64, F <= 11 (P <= 55 pairs).  Everything else stays on the fp32 SIMT kernels (afm.cu).
#include <stdlib.h>

#include "afm.cuh"
#include "staged.cuh"
#include "tc_common.cuh"

namespace hhfm {

namespace aft {

constexpr int KD = 64;                 // K == A
constexpr int kRows = 128;             // operand rows per tile (2 sample slots x 64)
constexpr int kSlot = 64;              // rows per sample slot
constexpr int kThreads = 256;          // 8 warps: warp & 3 = TMEM lane quarter (rows), warp >> 2 = column half
constexpr int kMaxF = 11;
constexpr int kDrain = 2;              // tiles between two drains of the dW accumulator
constexpr int kEP = KD + 4;            // padded row of the staged embeddings (floats)
constexpr uint32_t kTile = kRows * 128;          // bytes of one [128 rows][32 fp32] operand tile
constexpr uint32_t kWTile = KD * 128;            // bytes of one [64 rows][32 fp32] weight tile
constexpr uint32_t kTmemCols = 256;

// shared-memory map (bytes); operand tiles are 1024-byte aligned
constexpr uint32_t oP = 0;                              // [x kb0][x kb1][lo kb0][lo kb1]
constexpr uint32_t oZ = oP + 4 * kTile;                 // dZ tiles, same order; reused for the dP rows after GEMM 2/3
constexpr uint32_t oWt = oZ + 4 * kTile;                // W^T tiles (rows = a, contiguous k): [x kb0][x kb1][lo kb0][lo kb1]
constexpr uint32_t oWn = oWt + 4 * kWTile;              // W tiles (rows = k, contiguous a)
constexpr int kDWP = KD + 1;           // padded row of the dW accumulator (bank spread for the row-per-lane atomics)
constexpr uint32_t oDW = oWn + 4 * kWTile;              // float dW accumulator [64][kDWP]
constexpr uint32_t oE = oDW + KD * kDWP * 4;             // float E[2][kMaxF][kEP]
constexpr uint32_t oMisc = oE + 2 * kMaxF * kEP * 4;
constexpr uint32_t kSmemBytes = oMisc + 8192 + 1024;    // + alignment slack

struct Misc {
  float batt[KD], pvec[KD], wpred[KD];
  float gbatt[KD], gp[KD];
  float s_part[2][kRows];      // logits by column half
  float att[kRows];            // softmax weights a_p
  float da_part[2][kRows];
  float ds[kRows];
  float afm[2][KD], dafm[2][KD];
  float red[2][4];             // per-slot partial sums of out
  float g[2], bsum[2];
  int ids[2][kMaxF + 1];
  unsigned char pi[kSlot], pj[kSlot];
  unsigned char pidx[kMaxF][kMaxF + 1];
  uint64_t bar;
  uint32_t tmem;
};
static_assert(sizeof(Misc) <= 8192, "Misc does not fit its shared-memory slot");

// byte offset of 16-byte chunk c (0..7) of row r inside a 128-byte-swizzled tile
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

// MN-major operand descriptor (SWIZZLE_128B): 32 fp32 of the M/N index are contiguous (one tile row), the next block of
// 32 starts LBO bytes further (the next tile), 8 reduction rows form one 1024-byte atom, the next 8 start SBO bytes further.
__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32 instruction descriptor with both operands MN-major (transpose bits 15 / 16)
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N) { return make_idesc_tf32(M, N) | (1u << 15) | (1u << 16); }

__device__ __forceinline__ float4 lo4(float4 v) { return make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w)); }

// lane l ends with the sum over the warp's lanes of v[l] (reduce-scatter butterfly, 31 shuffles)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; i++) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

__global__ void __launch_bounds__(kThreads, 1) afm_fused_tc_kernel(const AfmArgs a, const int64_t n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float scratch[32];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  Misc& mi = *reinterpret_cast<Misc*>(smem + oMisc);
  float* dWs = reinterpret_cast<float*>(smem + oDW);
  float* Es = reinterpret_cast<float*>(smem + oE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = a.F, P = a.P;
  const int quarter = warp & 3, half = warp >> 2;
  const int row = quarter * 32 + lane;          // operand row / TMEM lane of this thread
  const int ss = row >> 6, pp = row & 63;       // sample slot, pair index
  const bool prow = pp < P;

  // ---- one-time setup: zero the operand tiles (pad rows stay zero), W tiles, small vectors, pair tables, TMEM ----
  for (uint32_t i = tid; i < (8 * kTile) / 16; i += kThreads) reinterpret_cast<float4*>(smem + oP)[i] = f4_zero();
  for (int i = tid; i < KD * kDWP; i += kThreads) dWs[i] = 0.f;
  for (int i = tid; i < KD * (KD / 4); i += kThreads) {
    const int k = i / (KD / 4), c = i % (KD / 4);           // W row k, float4 chunk c of the A columns
    const float4 w = __ldg(reinterpret_cast<const float4*>(a.W) + i);
    // W tiles (rows = k, contiguous a): B operand of dP = dZ . W^T
    *reinterpret_cast<float4*>(smem + oWn + (c >> 3) * kWTile + swz(k, c & 7)) = w;
    *reinterpret_cast<float4*>(smem + oWn + (2 + (c >> 3)) * kWTile + swz(k, c & 7)) = lo4(w);
    // W^T tiles (rows = a, contiguous k): B operand of Z = P . W
    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int an = 4 * c + q;
      const uint32_t off = swz(an, (k & 31) >> 2) + (uint32_t)(k & 3) * 4u;
      *reinterpret_cast<float*>(smem + oWt + (k >> 5) * kWTile + off) = wv[q];
      *reinterpret_cast<float*>(smem + oWt + (2 + (k >> 5)) * kWTile + off) = tf32_lo(wv[q]);
    }
  }
  if (tid < KD) {
    mi.batt[tid] = __ldg(a.batt + tid); mi.pvec[tid] = __ldg(a.pvec + tid); mi.wpred[tid] = __ldg(a.wpred + tid);
    mi.gbatt[tid] = 0.f; mi.gp[tid] = 0.f;
  }
  if (tid == 0) {
    int p = 0;
    for (int i = 0; i < F; i++)
      for (int j = i + 1; j < F; j++) {                      // AFM.py:105-112: pairs i < j in lexicographic order
        mi.pi[p] = (unsigned char)i; mi.pj[p] = (unsigned char)j;
        mi.pidx[i][j] = (unsigned char)p; mi.pidx[j][i] = (unsigned char)p;
        p++;
      }
    mbar_init(&mi.bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&mi.tmem, kTmemCols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = mi.tmem;
  const uint32_t t_lane = tmem + ((uint32_t)(quarter * 32) << 16);
  const uint32_t sP = smem_u32(smem + oP), sZ = smem_u32(smem + oZ), sWt = smem_u32(smem + oWt), sWn = smem_u32(smem + oWn);
  const float b0 = a.b0 ? __ldg(a.b0) : 0.f;
  const int rep = a.hot.slot ? (int)(blockIdx.x % a.hot.n_rep) : 0;
  uint32_t bar_ph = 0;
  float loss_acc = 0.f, gb0_acc = 0.f, gwp_acc = 0.f;
  int since_drain = 0;

  // add the TMEM dW partial sum into the shared accumulator: dW[k][a] = D[k][a] + D[64 + k][a]
  auto drain_dw = [&]() {
    uint32_t r[32];
    tmem_ld32(t_lane + 128 + half * 32, r);
    tmem_ld_wait_for(r);
    float* dst = dWs + (row & 63) * kDWP + half * 32;
#pragma unroll
    for (int i = 0; i < 32; i++) atomicAdd(dst + i, __uint_as_float(r[i]));
  };

  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    // ---- phase 0: ids, bias, embedding rows of the two samples ----
    if (tid < 2 * F) {
      const int s2 = tid / F, f = tid % F;
      const int64_t s = 2 * t + s2;
      mi.ids[s2][f] = (s < a.B) ? __ldg(a.idx + s * F + f) : -1;
    }
    __syncthreads();
    for (int i = tid; i < 2 * F * (KD / 4); i += kThreads) {
      const int s2 = i / (F * (KD / 4)), rem = i % (F * (KD / 4)), f = rem / (KD / 4), c = rem % (KD / 4);
      const int id = mi.ids[s2][f];
      const float4 v = id >= 0 ? __ldg(reinterpret_cast<const float4*>(a.V + (size_t)id * KD) + c) : f4_zero();
      *reinterpret_cast<float4*>(Es + (s2 * kMaxF + f) * kEP + 4 * c) = v;
    }
    if (tid < 2) {
      float bs = 0.f;
      if (a.bias)
        for (int f = 0; f < F; f++) { const int id = mi.ids[tid][f]; if (id >= 0) bs += __ldg(a.bias + id); }
      mi.bsum[tid] = bs;
    }
    __syncthreads();

    // ---- phase 1: pair products -> P operand tiles (x and lo(x)) ----
    for (int i = tid; i < 2 * P * (KD / 4); i += kThreads) {
      const int s2 = i / (P * (KD / 4)), rem = i % (P * (KD / 4)), p = rem / (KD / 4), c = rem % (KD / 4);
      const float4 ei = *reinterpret_cast<const float4*>(Es + (s2 * kMaxF + mi.pi[p]) * kEP + 4 * c);
      const float4 ej = *reinterpret_cast<const float4*>(Es + (s2 * kMaxF + mi.pj[p]) * kEP + 4 * c);
      const float4 v = f4_mul(ei, ej);
      const uint32_t off = (uint32_t)(c >> 3) * kTile + swz(s2 * kSlot + p, c & 7);
      *reinterpret_cast<float4*>(smem + oP + off) = v;
      *reinterpret_cast<float4*>(smem + oP + 2 * kTile + off) = lo4(v);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_tf32(kRows, KD);
#pragma unroll
      for (int kb = 0; kb < 2; kb++)
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {
          const uint32_t o = k4 * 32;
          const uint32_t ax = sP + kb * kTile + o, al = sP + (2 + kb) * kTile + o;
          const uint32_t bx = sWt + kb * kWTile + o, bl = sWt + (2 + kb) * kWTile + o;
          umma_tf32(tmem, make_sdesc(al), make_sdesc(bx), idesc, (kb | k4) ? 1u : 0u);
          umma_tf32(tmem, make_sdesc(ax), make_sdesc(bl), idesc, 1u);
          umma_tf32(tmem, make_sdesc(ax), make_sdesc(bx), idesc, 1u);
        }
      umma_commit(&mi.bar);
    }
    mbar_wait(&mi.bar, bar_ph, nullptr);
    bar_ph ^= 1;
    tc_fence_after();

    // ---- epilogue 1: logits, softmax, afm, out, loss ----
    float z[32];
    {
      uint32_t r[32];
      tmem_ld32(t_lane + half * 32, r);
      tmem_ld_wait_for(r);
      float sp = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i++) {
        z[i] = __uint_as_float(r[i]) + mi.batt[half * 32 + i];                   // AFM.py:117-123
        sp = fmaf(fmaxf(z[i], 0.f), mi.pvec[half * 32 + i], sp);
      }
      mi.s_part[half][row] = sp;
    }
    __syncthreads();
    float att = 0.f;
    {
      const float* s0 = mi.s_part[0] + ss * kSlot;
      const float* s1 = mi.s_part[1] + ss * kSlot;
      float mx = -INFINITY;
      for (int p = 0; p < P; p++) mx = fmaxf(mx, s0[p] + s1[p]);
      float den = 0.f;
      for (int p = 0; p < P; p++) den += expf((s0[p] + s1[p]) - mx);            // AFM.py:125 softmax over the pairs
      if (prow) att = expf((s0[pp] + s1[pp]) - mx) / den;
      if (half == 0) mi.att[row] = att;
    }
    __syncthreads();
    if (tid < 2 * KD) {
      const int s2 = tid >> 6, k = tid & 63;
      const float* e = Es + s2 * kMaxF * kEP + k;
      float acc = 0.f;
      for (int p = 0; p < P; p++) acc = fmaf(mi.att[s2 * kSlot + p], e[mi.pi[p] * kEP] * e[mi.pj[p] * kEP], acc);   // AFM.py:130
      mi.afm[s2][k] = acc;
      const float part = warp_sum(acc * mi.wpred[k]);                            // AFM.py:138-139
      if (lane == 0) mi.red[s2][warp & 1] = part;
    }
    __syncthreads();
    if (tid < 2) {
      const int64_t s = 2 * t + tid;
      float g = 0.f;
      if (s < a.B) {
        const float out = ((mi.red[tid][0] + mi.red[tid][1]) + mi.bsum[tid]) + b0;   // AFM.py:142
        const float diff = __ldg(a.labels + s) - out;
        g = -diff;
        loss_acc += 0.5f * diff * diff;                                          // AFM.py:146 tf.nn.l2_loss
        gb0_acc += g;
        if (a.out) a.out[s] = out;
      }
      mi.g[tid] = g;
    }
    __syncthreads();
    if (tid < 2 * KD) {
      const int s2 = tid >> 6, k = tid & 63;
      const float g = mi.g[s2];
      mi.dafm[s2][k] = g * mi.wpred[k];
      gwp_acc = fmaf(g, mi.afm[s2][k], gwp_acc);
    }
    __syncthreads();
    {
      // d a_p = d afm . P_p over this thread's 32 columns
      const float* ei = Es + (ss * kMaxF + mi.pi[prow ? pp : 0]) * kEP + half * 32;
      const float* ej = Es + (ss * kMaxF + mi.pj[prow ? pp : 0]) * kEP + half * 32;
      const float* df = mi.dafm[ss] + half * 32;
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 x = *reinterpret_cast<const float4*>(ei + i), y = *reinterpret_cast<const float4*>(ej + i);
        const float4 d = *reinterpret_cast<const float4*>(df + i);
        acc = fmaf(d.x, x.x * y.x, acc); acc = fmaf(d.y, x.y * y.y, acc); acc = fmaf(d.z, x.z * y.z, acc); acc = fmaf(d.w, x.w * y.w, acc);
      }
      mi.da_part[half][row] = prow ? acc : 0.f;
    }
    __syncthreads();
    {
      const float* d0 = mi.da_part[0] + ss * kSlot;
      const float* d1 = mi.da_part[1] + ss * kSlot;
      const float* at = mi.att + ss * kSlot;
      float sd = 0.f;
      for (int p = 0; p < P; p++) sd = fmaf(at[p], d0[p] + d1[p], sd);
      const float ds = att * ((d0[pp] + d1[pp]) - sd);                           // softmax backward
      // dZ = ds * p * relu'(Z + b); d p += ds * relu(Z + b); d b += dZ (column sums over the tile rows)
      float dz[32], hp[32];
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const bool on = z[i] > 0.f;
        dz[i] = on ? ds * mi.pvec[half * 32 + i] : 0.f;
        hp[i] = on ? ds * z[i] : 0.f;
      }
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const float4 v = make_float4(dz[4 * c], dz[4 * c + 1], dz[4 * c + 2], dz[4 * c + 3]);
        const uint32_t off = (uint32_t)half * kTile + swz(row, c);
        *reinterpret_cast<float4*>(smem + oZ + off) = v;
        *reinterpret_cast<float4*>(smem + oZ + 2 * kTile + off) = lo4(v);
      }
      const float cb = warp_colsum32(dz, lane);
      const float cp = warp_colsum32(hp, lane);
      atomicAdd(&mi.gbatt[half * 32 + lane], cb);
      atomicAdd(&mi.gp[half * 32 + lane], cp);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc_tf32(kRows, KD);
#pragma unroll
      for (int ab = 0; ab < 2; ab++)
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {
          const uint32_t o = k4 * 32;
          const uint32_t ax = sZ + ab * kTile + o, al = sZ + (2 + ab) * kTile + o;
          const uint32_t bx = sWn + ab * kWTile + o, bl = sWn + (2 + ab) * kWTile + o;
          umma_tf32(tmem + 64, make_sdesc(al), make_sdesc(bx), idesc, (ab | k4) ? 1u : 0u);
          umma_tf32(tmem + 64, make_sdesc(ax), make_sdesc(bl), idesc, 1u);
          umma_tf32(tmem + 64, make_sdesc(ax), make_sdesc(bx), idesc, 1u);
        }
      const uint32_t idesc_mn = make_idesc_tf32_mn(kRows, KD);
#pragma unroll
      for (int ks = 0; ks < kRows / 8; ks++) {               // 8 tile rows (one swizzle atom) per MMA
        const uint32_t o = ks * 1024;
        const uint64_t ad = make_sdesc_mn(sP + o, kTile);     // M blocks: x kb0, x kb1, lo kb0, lo kb1
        umma_tf32(tmem + 128, ad, make_sdesc_mn(sZ + o, kTile), idesc_mn, (since_drain | ks) ? 1u : 0u);
        umma_tf32(tmem + 128, ad, make_sdesc_mn(sZ + 2 * kTile + o, kTile), idesc_mn, 1u);
      }
      umma_commit(&mi.bar);
    }
    mbar_wait(&mi.bar, bar_ph, nullptr);
    bar_ph ^= 1;
    tc_fence_after();
    since_drain++;

    // ---- epilogue 2: dP rows (+ the direct path a_p * d afm) -> shared memory (the dZ tiles are free now) ----
    {
      uint32_t r[32];
      tmem_ld32(t_lane + 64 + half * 32, r);
      tmem_ld_wait_for(r);
      const float* df = mi.dafm[ss] + half * 32;
      float* dst = reinterpret_cast<float*>(smem + oZ) + row * KD;
#pragma unroll
      for (int c = 0; c < 8; c++) {
        const float4 v = make_float4(fmaf(att, df[4 * c], __uint_as_float(r[4 * c])), fmaf(att, df[4 * c + 1], __uint_as_float(r[4 * c + 1])),
                                     fmaf(att, df[4 * c + 2], __uint_as_float(r[4 * c + 2])), fmaf(att, df[4 * c + 3], __uint_as_float(r[4 * c + 3])));
        *reinterpret_cast<float4*>(dst + 4 * ((half * 8 + c) ^ (row & 15))) = v;
      }
      if (since_drain == kDrain) drain_dw();
    }
    if (since_drain == kDrain) since_drain = 0;
    tc_fence_before();
    __syncthreads();
    // dE_f = sum_{j != f} dP_(f,j) * E_j ; one vector reduction per 16 bytes of the gradient row
    for (int i = tid; i < 2 * F * (KD / 4); i += kThreads) {
      const int s2 = i / (F * (KD / 4)), rem = i % (F * (KD / 4)), f = rem / (KD / 4), c = rem % (KD / 4);
      const int id = mi.ids[s2][f];
      if (id < 0) continue;
      float4 acc = f4_zero();
      for (int j = 0; j < F; j++) {
        if (j == f) continue;
        const int r2 = s2 * kSlot + mi.pidx[f][j];
        const float4 dp = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(smem + oZ) + r2 * KD + 4 * (c ^ (r2 & 15)));
        const float4 e = *reinterpret_cast<const float4*>(Es + (s2 * kMaxF + j) * kEP + 4 * c);
        acc.x = fmaf(dp.x, e.x, acc.x); acc.y = fmaf(dp.y, e.y, acc.y); acc.z = fmaf(dp.z, e.z, acc.z); acc.w = fmaf(dp.w, e.w, acc.w);
      }
      float* dst = a.gV + (size_t)id * KD;
      int hs = -1;
      if (a.hot.slot != nullptr) {
        hs = __ldg(a.hot.slot + id);
        if (hs >= 0) dst = a.hot.ghot + ((size_t)rep * a.hot.n_hot + hs) * KD;
      }
      red_add_v4(dst + 4 * c, acc);
      if (c == 0) {
        if (a.gbias) {
          float* pb = (hs >= 0 && a.hot.ghot_bias != nullptr) ? a.hot.ghot_bias + (size_t)rep * a.hot.n_hot + hs : a.gbias + id;
          atomicAdd(pb, mi.g[s2]);
        }
        if (a.touch_stamp) a.touch_stamp[id] = a.stamp;      // compacted into the touched-row list afterwards
      }
    }
    __syncthreads();        // Es / ids / the dP rows are free for the next tile; the P tiles were released by the commit wait
    // the dZ tile region was used as plain storage: its pad rows must read as zero again for the next GEMM 2 / 3
    for (int i = tid; i < 2 * (kSlot - P) * 8; i += kThreads) {
      const int s2 = i / ((kSlot - P) * 8), rem = i % ((kSlot - P) * 8), r2 = s2 * kSlot + P + rem / 8, c = rem % 8;
      *reinterpret_cast<float4*>(smem + oZ + swz(r2, c)) = f4_zero();
      *reinterpret_cast<float4*>(smem + oZ + kTile + swz(r2, c)) = f4_zero();
    }
  }

  if (since_drain != 0) drain_dw();
  tc_fence_before();
  __syncthreads();
  // ---- flush the CTA's accumulators ----
  for (int i = tid; i < KD * KD; i += kThreads) {
    const float v = dWs[(i >> 6) * kDWP + (i & 63)];
    if (v != 0.f) atomicAdd(a.gW + i, v);
  }
  if (tid < KD) {
    atomicAdd(a.gbatt + tid, mi.gbatt[tid]);
    atomicAdd(a.gp + tid, mi.gp[tid]);
  }
  if (tid < 2 * KD) atomicAdd(a.gwpred + (tid & 63), gwp_acc);
  {
    const float bl = block_sum(loss_acc, scratch);
    write_partial(a.loss_partials, bl);
    const float bg = block_sum(gb0_acc, scratch);
    if (tid == 0 && a.gb0 != nullptr && bg != 0.f) atomicAdd(a.gb0, bg);
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace aft

int dispatch_afm_fused_tc(const AfmArgs& a, int64_t M, cudaStream_t st) {
  if (a.K != aft::KD || a.A != aft::KD || a.F > aft::kMaxF || a.F < 2 || a.P > aft::kSlot) return 1;
  const char* env = getenv("HHFM_AFM_TC");               // opt-in while the dW product is being reworked
  if (!env || env[0] != '1') return 1;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(aft::afm_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)aft::kSmemBytes) != cudaSuccess) {
      cudaGetLastError();
      return 1;
    }
    configured = true;
  }
  const int64_t n_tiles = (a.B + 1) / 2;
  int grid = sm_count();
  if (grid > kPartials) grid = kPartials;
  if ((int64_t)grid > n_tiles) grid = (int)n_tiles;
  aft::afm_fused_tc_kernel<<<grid, aft::kThreads, aft::kSmemBytes, st>>>(a, n_tiles);
  int rc = check_launch("afm_fused_tc_kernel");
  if (rc != HHFM_OK) return rc;
  if (a.touch_stamp != nullptr) rc = launch_touched_compact(a.touch_stamp, a.stamp, M, a.touched_rows, a.touched_count, st);
  return rc;
}

}  // namespace hhfm
