// K6 tensor-core path: full-catalog scoring as a bf16 GEMM on tcgen05 (accumulators in TMEM, operands staged by TMA)
// used as a FILTER, followed by exact fp32 rescoring of the few survivors -- index lists stay bit-identical to the
// exact path / the oracle (FM.py:172-185, BPR.py:131-136, MF.py:144-149, OurModel7.py:229-295).
//
//   1. prep      Q [C,K] fp32 -> bf16 A [C,Kp], per-row ||q||, (FM) R_c = q.Fc and ||Fc||;  items -> bf16 B [N,Kp]
//                (FM: the item bias rides in an extra k-chunk as hi+lo bf16 against two 1.0 columns of A), max ||v||.
//   2. GEMM #1   tc_score_kernel<.., EMIT=false>: persistent, warp-specialised (TMA producer / single-thread tcgen05.mma
//                issuer / 4 epilogue warps).  CTA tile 128 contexts x BN items, all of Kp per stage.  The epilogue reads
//                the fp32 accumulator with tcgen05.ld (thread t owns context row t) and keeps only group maxima:
//                one per 32 items (small catalogs) or one per BN-item tile -> gmax [C, n_groups].
//   3. threshold tau_c = tp-th largest group maximum of row c (select_kernel).  The tp group maxima are tp distinct
//                items with approximate score >= tau_c, so the exact tp-th best score s* >= tau_c - E_c.
//   4. GEMM #2   tc_score_kernel<.., EMIT=true>: the same GEMM again (it is the cheapest stage), the epilogue now
//                compares against tau_c - 2 E_c and appends the ids of the few survivors (~2 tp per row).
//   5. rescore   survivors are rescored exactly in the canonical fp32 order (512 B per candidate instead of whole
//                groups), then select_kernel sorts them (score desc, id asc).
// E_c bounds |approx - exact| for row c: bf16 rounding (rel. 2^-8 per operand) gives 2^-7 ||q|| ||v|| by Cauchy-Schwarz;
// we use 2^-7 ||q_c|| max_n||v_n|| plus the bias split and fp32 rounding terms.  Every true top-tp item n has
// approx_n >= s* - E >= tau - 2E, so it is emitted: the candidate set contains the exact answer, ties included.
// A row whose candidates overflow the buffer is flagged and redone by the exact path (host side).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace hhfm {

constexpr int kGroup = 32;          // items per group maximum (= one tcgen05.ld.32x32b.x32)
constexpr int kBM = 128;            // contexts per CTA tile (UMMA M)
constexpr int kKC = 64;             // bf16 elements per 128-byte swizzle row
constexpr int kTcSubs = 4;           // epilogue warps per TMEM lane quarter (each owns BN/4 accumulator columns)
constexpr int kTcThreads = (2 + 4 * kTcSubs) * 32;   // warps 0-3 and 6-17: epilogue, warp 4: TMA, warp 5: MMA

// ---------------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_max_pos(float* addr, float v) {   // v >= 0
  atomicMax(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// items fp32 [N,K] (+bias [N]) -> bf16 [N,Kp]; stats[0] = max ||v_n||, stats[1] = max |b_n|.  One warp per item.
__global__ void __launch_bounds__(256) tc_prep_items_kernel(const float* __restrict__ items, const float* __restrict__ bias,
                                                            int64_t N, int K, int Kp, __nv_bfloat16* __restrict__ out,
                                                            float* __restrict__ stats) {
  const int64_t n = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float sq = 0.f;
  for (int k = lane; k < Kp; k += 32) {
    float v = 0.f;
    if (k < K) {
      v = __ldg(items + n * K + k);
      sq += v * v;
    } else if (bias != nullptr && k == K) {
      v = __bfloat162float(__float2bfloat16_rn(__ldg(bias + n)));
    } else if (bias != nullptr && k == K + 1) {
      const float b = __ldg(bias + n);
      v = b - __bfloat162float(__float2bfloat16_rn(b));
    }
    out[n * Kp + k] = __float2bfloat16_rn(v);
  }
  sq = warp_sum(sq);
  if (lane == 0) {
    atomic_max_pos(stats + 0, sqrtf(sq) * 1.0000005f);
    if (bias != nullptr) atomic_max_pos(stats + 1, fabsf(__ldg(bias + n)));
  }
}

// Q fp32 [C,K] (+Fc for FM) -> bf16 [C,Kp]; qinfo[c] = {||q||, R = q.Fc, ||Fc||, 0}.  One warp per context.
__global__ void __launch_bounds__(256) tc_prep_queries_kernel(const float* __restrict__ Q, const float* __restrict__ Fc,
                                                              int64_t C, int K, int Kp, int fm, __nv_bfloat16* __restrict__ out,
                                                              float4* __restrict__ qinfo) {
  const int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (c >= C) return;
  float sq = 0.f, r = 0.f, sf = 0.f;
  for (int k = lane; k < Kp; k += 32) {
    float v = 0.f;
    if (k < K) {
      v = __ldg(Q + c * K + k);
      sq += v * v;
      if (fm) {
        const float f = __ldg(Fc + c * K + k);
        r = fmaf(v, f, r);
        sf += f * f;
      }
    } else if (fm && (k == K || k == K + 1)) {
      v = 1.f;
    }
    out[c * Kp + k] = __float2bfloat16_rn(v);
  }
  sq = warp_sum(sq); r = warp_sum(r); sf = warp_sum(sf);
  if (lane == 0) qinfo[c] = make_float4(sqrtf(sq) * 1.0000005f, r, sqrtf(sf) * 1.0000005f, 0.f);
}

// E_c: bound on |approximate - exact| for row c (see header).
__device__ __forceinline__ float row_error_bound(float4 qi, float vmax, float bmax, int fm, int K) {
  float e = 0.0078125f * qi.x * vmax;                                // 2^-7 ||q|| max||v||
  e += 1.1920929e-07f * (float)(K + 8) * qi.x * (vmax + qi.z);       // fp32 accumulation / canonical-order rounding
  if (fm) e += 1.5258789e-05f * bmax + 4.7683716e-07f * (fabsf(qi.y) + bmax);   // bias hi+lo split, R_c rounding
  return e * 1.01f + 1e-30f;
}

// ---------------------------------------------------------------------------------------------------
// the GEMM: gmax[c, g] = max over items n in group g of (A[c,:] . B[n,:])
// ---------------------------------------------------------------------------------------------------
struct TcArgs {
  int n_row_blocks, n_tiles, tiles_per_unit, splits, n_units;
  int64_t C, N;
  int64_t gmax_stride;    // floats per gmax row
  float* gmax;
  int fine;               // 1: one maximum per 32 items, 0: one per BN-item tile
  int tile_stride;        // max pass on a SAMPLE of the catalog: unit tile t covers item tile t * tile_stride
  const float* thr_emit;  // EMIT: per-row threshold in approximate-score space
  int32_t* seg_ids;       // EMIT: [C, splits, cap_u] survivor ids; segment (row, split) is private to ONE thread
  int32_t* seg_cnt;       // EMIT: [C, splits] survivors found (may exceed cap_u -> overflow)
  int cap_u;
  int* err;
};

template <int NKC, int BN, int STAGES>
struct TcSmem {
  static constexpr int kABytes = NKC * kBM * 128;
  static constexpr int kBStage = NKC * BN * 128;
  static constexpr int kBytes = kABytes + STAGES * kBStage + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int NKC, int BN, int STAGES, bool EMIT>
__global__ void __launch_bounds__(kTcThreads, 1) tc_score_kernel(const __grid_constant__ CUtensorMap tmA,
                                                          const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + TcSmem<NKC, BN, STAGES>::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + STAGES * TcSmem<NKC, BN, STAGES>::kBStage);
  uint64_t* full = bars;                    // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;          // [STAGES] MMA -> TMA
  uint64_t* a_full = bars + 2 * STAGES;     // A tile landed
  uint64_t* a_empty = a_full + 1;           // MMA finished with the A tile
  uint64_t* t_full = a_empty + 1;           // [2] accumulator ready
  uint64_t* t_empty = t_full + 2;           // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 2 * BN;    // 512 (BN=256) or 256 (BN=128): power of two >= 32

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; i++) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(a_full, 1); mbar_init(a_empty, 1);
    for (int i = 0; i < 2; i++) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 4 * kTcSubs * 32); }
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ================= TMA producer (one lane) =================
    if (lane == 0) {
      int st = 0; uint32_t ph = 0, aph = 0;
      for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        const int rb = u / a.splits, sp = u % a.splits;
        const int t0 = sp * a.tiles_per_unit, t1 = min(a.n_tiles, t0 + a.tiles_per_unit);
        mbar_wait(a_empty, aph ^ 1, a.err);
        mbar_expect_tx(a_full, TcSmem<NKC, BN, STAGES>::kABytes);
#pragma unroll
        for (int kc = 0; kc < NKC; kc++) tma_load_2d(sA + kc * kBM * 128, &tmA, kc * kKC, rb * kBM, a_full);
        aph ^= 1;
        for (int t = t0; t < t1; t++) {
          mbar_wait(empty + st, ph ^ 1, a.err);
          mbar_expect_tx(full + st, TcSmem<NKC, BN, STAGES>::kBStage);
#pragma unroll
          for (int kc = 0; kc < NKC; kc++)
            tma_load_2d(sB + st * TcSmem<NKC, BN, STAGES>::kBStage + kc * BN * 128, &tmB, kc * kKC, t * a.tile_stride * BN, full + st);
          if (++st == STAGES) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ================= MMA issuer (one lane) =================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBM, BN);
      int st = 0, acc = 0; uint32_t ph = 0, aph = 0, tph = 0;
      for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        const int sp = u % a.splits;
        const int t0 = sp * a.tiles_per_unit, t1 = min(a.n_tiles, t0 + a.tiles_per_unit);
        mbar_wait(a_full, aph, a.err);
        aph ^= 1;
        for (int t = t0; t < t1; t++) {
          mbar_wait(t_empty + acc, tph ^ 1, a.err);
          mbar_wait(full + st, ph, a.err);
          tc_fence_after();
          const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + st * TcSmem<NKC, BN, STAGES>::kBStage);
#pragma unroll
          for (int kc = 0; kc < NKC; kc++) {
#pragma unroll
            for (int k4 = 0; k4 < 4; k4++) {       // UMMA_K = 16 bf16 = 32 bytes inside the 128-byte swizzle row
              const uint64_t ad = make_sdesc(a0 + kc * kBM * 128 + k4 * 32);
              const uint64_t bd = make_sdesc(b0 + kc * BN * 128 + k4 * 32);
              umma_bf16(tmem_base + acc * BN, ad, bd, idesc, (kc | k4) != 0);
            }
          }
          umma_commit(empty + st);        // smem stage reusable once these MMAs retire
          umma_commit(t_full + acc);      // accumulator complete
          if (++st == STAGES) { st = 0; ph ^= 1; }
          if (++acc == 2) { acc = 0; tph ^= 1; }
        }
        umma_commit(a_empty);             // A tile reusable
      }
    }
  } else {
    // ================= epilogue: 16 warps.  A warp may only touch the TMEM lane quarter (warp % 4), so the four warps of a
    // quarter share its 32 context rows and split the BN accumulator columns in four: thread (quarter, lane, sub) owns
    // context row quarter*32+lane for BN/4 item columns.  History: 4 warps -> tensor pipe 42 % (emission) / 69 % (max pass)
    // busy waiting for the epilogue; 8 warps -> emission 1.21 -> 0.95 ms; 16 warps with 16-column TMEM loads (two register
    // buffers of 16, no spills at 576 threads). =================
    int acc = 0; uint32_t tph = 0;
    const int quarter = warp & 3;
    const int sub = (warp < 4 ? warp : warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    constexpr int kChunksTile = BN / kGroup;
    constexpr int kCols = BN / kTcSubs;              // columns per thread and tile (64 or 32)
    constexpr int kChunks = kCols / kGroup;          // 32-item groups per thread and tile (2 or 1)
    constexpr int kC16 = kCols / 16;                 // 16-column TMEM loads per thread and tile (4 or 2)
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
      const int rb = u / a.splits, sp = u % a.splits;
      const int t0 = sp * a.tiles_per_unit, t1 = min(a.n_tiles, t0 + a.tiles_per_unit);
      const int64_t row = (int64_t)rb * kBM + row_in_tile;
      const float thr = (EMIT && row < a.C) ? __ldg(a.thr_emit + row) : INFINITY;
      int32_t* seg = EMIT ? a.seg_ids + (((int64_t)row * a.splits + sp) * kTcSubs + sub) * a.cap_u : nullptr;
      int n_emit = 0;
      for (int t = t0; t < t1; t++) {
        mbar_wait(t_full + acc, tph, a.err);
        tc_fence_after();
        float gm[kChunks];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + sub * kCols);
        const int64_t n_base = (int64_t)t * a.tile_stride * BN + sub * kCols;
        if (!EMIT) {
#pragma unroll
          for (int c = 0; c < kChunks; c++) {
            uint32_t r[32];
            tmem_ld32(taddr + c * kGroup, r);
            tmem_ld_wait();
            float m = -INFINITY;
            if (n_base + (c + 1) * kGroup <= a.N) {
#pragma unroll
              for (int i = 0; i < 32; i++) m = fmaxf(m, __uint_as_float(r[i]));
            } else {                            // last tile: items beyond N are TMA zero fill, not scores
#pragma unroll
              for (int i = 0; i < 32; i++)
                if (n_base + c * kGroup + i < a.N) m = fmaxf(m, __uint_as_float(r[i]));
            }
            gm[c] = m;
          }
        } else {
          // Survivor emission.  Two register buffers: the tcgen05.ld of chunk c+1 is in flight while chunk c is scanned
          // (with one buffer every chunk paid the full TMEM load latency).  The scan is a max tree over two 8-item
          // sub-groups; only a sub-group holding a survivor builds a bit mask.  The loop is unrolled by 2 only, so the
          // epilogue stays inside the instruction cache (an unrolled 8 x 32 emission body is ~100 KB of SASS and ran 10x
          // slower).
          auto scan = [&](const uint32_t (&r)[16], int c) {
            float mq[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
              float m = __uint_as_float(r[8 * q]);
#pragma unroll
              for (int i = 1; i < 8; i++) m = fmaxf(m, __uint_as_float(r[8 * q + i]));
              mq[q] = m;
            }
            if (fmaxf(mq[0], mq[1]) >= thr) {   // rare: a few hundred survivors per row
#pragma unroll
              for (int q = 0; q < 2; q++) {
                if (mq[q] >= thr) {
                  unsigned mask = 0u;
#pragma unroll
                  for (int i = 0; i < 8; i++) mask |= (__uint_as_float(r[8 * q + i]) >= thr ? 1u : 0u) << i;
                  const int nb = (int)n_base + c * 16 + 8 * q;
                  while (mask) {                // no atomics: the segment belongs to this thread
                    const int i = __ffs((int)mask) - 1;
                    mask &= mask - 1;
                    if (nb + i < a.N) {
                      if (n_emit < a.cap_u) seg[n_emit] = nb + i;
                      n_emit++;
                    }
                  }
                }
              }
            }
          };
          uint32_t ra[16], rb[16];
          tmem_ld16(taddr, ra);
#pragma unroll 1
          for (int c = 0; c < kC16; c += 2) {
            tmem_ld_wait_for16(ra);
            tmem_ld16(taddr + (c + 1) * 16, rb);
            scan(ra, c);
            tmem_ld_wait_for16(rb);
            if (c + 2 < kC16) tmem_ld16(taddr + (c + 2) * 16, ra);
            scan(rb, c + 1);
          }
        }
        tc_fence_before();
        mbar_arrive(t_empty + acc);
        if (!EMIT && row < a.C) {
          if (a.fine) {
            float* dst = a.gmax + row * a.gmax_stride + (int64_t)t * kChunksTile + sub * kChunks;
            if (kChunks == 2) *reinterpret_cast<float2*>(dst) = make_float2(gm[0], gm[kChunks - 1]);
            else dst[0] = gm[0];
          } else {                                   // one maximum per quarter tile
            float m = gm[0];
#pragma unroll
            for (int c = 1; c < kChunks; c++) m = fmaxf(m, gm[c]);
            a.gmax[row * a.gmax_stride + kTcSubs * t + sub] = m;
          }
        }
        if (++acc == 2) { acc = 0; tph ^= 1; }
      }
      if (EMIT && row < a.C) a.seg_cnt[(row * a.splits + sp) * kTcSubs + sub] = n_emit;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------
// thresholds: thr_emit[c] = tau_c - 2 E_c (approximate-score space)
// ---------------------------------------------------------------------------------------------------
__global__ void tc_threshold_kernel(int64_t C, int fm, int K, const float* __restrict__ tau, int tau_stride,
                                    const float4* __restrict__ qinfo, const float* __restrict__ stats, int sampled,
                                    float* __restrict__ thr_emit, float* __restrict__ thr_verify) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float E = row_error_bound(qinfo[c], stats[0], stats[1], fm, K);
  if (sampled) {
    // tau is a rank statistic of a SAMPLE of the catalog: a heuristic cut, proven per row after rescoring:
    // an item that was not emitted has approx < thr_emit, hence exact < thr_emit + E = thr_verify
    thr_emit[c] = tau[c * tau_stride];
    thr_verify[c] = tau[c * tau_stride] + E;
  } else {
    thr_emit[c] = tau[c * tau_stride] - 2.f * E;      // guaranteed: tp distinct items have approx >= tau
    thr_verify[c] = -INFINITY;
  }
}

// Sampled mode cut: any value whose rank among the n sampled group maxima is close to j works (the cut is a heuristic
// that is proven per row afterwards), so instead of an exact radix select this finds the lower edge of the bucket that
// holds the j-th largest value in a 256-bucket histogram over [min, max], refined (at most twice more) while that
// bucket holds more than max(4, j/8) values.  The result has rank >= j and < j + max(4, j/8) (or the refinement depth
// ran out on heavily tied data, which only lowers the cut).  One CTA per row, the row staged in shared memory.
constexpr int kCutThreads = 128;
__global__ void __launch_bounds__(kCutThreads) tc_cut_kernel(const float* __restrict__ gmax, int64_t gmax_stride, int n, int j,
                                                             int fm, int K, const float4* __restrict__ qinfo,
                                                             const float* __restrict__ stats, float* __restrict__ thr_emit,
                                                             float* __restrict__ thr_verify) {
  extern __shared__ float s_val[];                 // n values
  __shared__ unsigned s_hist[256];
  __shared__ float s_red[2][kCutThreads / 32];
  __shared__ float s_lo, s_width;
  __shared__ int s_need, s_done;
  const int64_t c = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  const float* row = gmax + c * gmax_stride;
  float mx = -INFINITY, mn = INFINITY;
  for (int i = t; i < n; i += kCutThreads) {
    const float v = row[i];
    s_val[i] = v;
    if (v > -INFINITY) { mx = fmaxf(mx, v); mn = fminf(mn, v); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if (lane == 0) { s_red[0][w] = mx; s_red[1][w] = mn; }
  __syncthreads();
  if (t == 0) {
    float a = s_red[0][0], b = s_red[1][0];
    for (int i = 1; i < kCutThreads / 32; i++) { a = fmaxf(a, s_red[0][i]); b = fminf(b, s_red[1][i]); }
    s_lo = b;
    s_width = (a - b) * (1.0f / 256.0f);
    s_need = j;
    s_done = !(a > b);                               // all equal (or empty): the cut is that value
  }
  __syncthreads();
  for (int level = 0; level < 3 && !s_done; level++) {
    for (int i = t; i < 256; i += kCutThreads) s_hist[i] = 0u;
    __syncthreads();
    const float lo = s_lo, width = s_width, inv = 1.0f / width;
    for (int i = t; i < n; i += kCutThreads) {
      const float v = s_val[i];
      if (v >= lo) {
        const float f = (v - lo) * inv;
        // values above the current range were counted by the level before (s_need already excludes them)
        if (level == 0 || f < 256.0f) atomicAdd(&s_hist[min(255, (int)f)], 1u);
      }
    }
    __syncthreads();
    if (t < 32) {
      unsigned h[8], T = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) { h[q] = s_hist[t * 8 + q]; T += h[q]; }
      unsigned S = T;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_down_sync(0xffffffffu, S, o);
        if (t + o < 32) S += v;
      }
      const unsigned need = (unsigned)s_need;
      const unsigned bal = __ballot_sync(0xffffffffu, S >= need);
      const int target = bal ? 31 - __clz((int)bal) : 0;      // fewer than `need` values in range: take the lowest bucket
      if (t == target) {
        unsigned cum = S - T;
        int bsel = 0;
        bool done = false;
#pragma unroll
        for (int q = 7; q >= 0; q--) {
          if (!done) {
            if (q == 0 || cum + h[q] >= need) { bsel = q; done = true; }
            else cum += h[q];
          }
        }
        const int b = t * 8 + bsel;
        const unsigned in_bucket = h[0] * (bsel == 0) + h[1] * (bsel == 1) + h[2] * (bsel == 2) + h[3] * (bsel == 3) +
                                   h[4] * (bsel == 4) + h[5] * (bsel == 5) + h[6] * (bsel == 6) + h[7] * (bsel == 7);
        const float new_lo = lo + (float)b * width;
        s_lo = new_lo;
        s_width = width * (1.0f / 256.0f);
        s_need = (int)(need > cum ? need - cum : 1);
        const unsigned limit = (unsigned)max(4, j / 8);
        if (in_bucket <= limit || !(new_lo + width * (1.0f / 256.0f) > new_lo)) s_done = 1;
      }
    }
    __syncthreads();
  }
  if (t == 0) {
    const float tau = s_lo;
    const float E = row_error_bound(qinfo[c], stats[0], stats[1], fm, K);
    thr_emit[c] = tau;
    thr_verify[c] = tau + E;
  }
}

// Sampled mode: the candidate set of row c is proven complete iff it holds >= tp items and the tp-th best exact score
// is >= thr_verify[c] (every item that was filtered out scores strictly below that).  Otherwise the row is flagged and
// redone by the exact path, like a candidate-buffer overflow.
__global__ void tc_verify_kernel(int64_t C, int tp, const float* __restrict__ out_scores, const int32_t* __restrict__ cand_cnt,
                                 const float* __restrict__ thr_verify, int32_t* __restrict__ overflow) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float tv = thr_verify[c];
  if (tv == -INFINITY) return;
  if (cand_cnt[c] < tp || !(out_scores[c * tp + (tp - 1)] >= tv)) overflow[c] = 1;
}

// ---------------------------------------------------------------------------------------------------
// exact rescoring of the survivors.  One CTA per context row; a warp takes 32 candidates at a time, stages their item
// rows 32 k at a time with coalesced 128-byte reads, and lane r walks row r in ascending k (canonical order).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tc_rescore_kernel(int kind, const float* __restrict__ Q, const float* __restrict__ Fc,
                                                         const float* __restrict__ items, const float* __restrict__ item_bias,
                                                         int64_t N, int K, int splits, int cap_u, const int32_t* __restrict__ seg_ids,
                                                         const int32_t* __restrict__ seg_cnt, int cap, float* __restrict__ cand_scores,
                                                         int32_t* __restrict__ cand_ids, int32_t* __restrict__ cand_cnt,
                                                         int32_t* __restrict__ overflow) {
  __shared__ float s_tile[8][32][33];
  __shared__ int s_off[512];                // prefix of per-split survivor counts (splits <= 2*SMs+1 <= 511)
  __shared__ int s_over;
  extern __shared__ float s_q[];            // q[K] (+ Fc[K] for FM)
  const int64_t c = blockIdx.x;
  const int fm = kind == HHFM_QUERY_FM;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    s_q[k] = Q[c * K + k];
    if (fm) s_q[K + k] = Fc[c * K + k];
  }
  if (threadIdx.x == 0) {
    int tot = 0, over = 0;
    for (int sp = 0; sp < splits; sp++) {
      int n = seg_cnt[c * splits + sp];
      if (n > cap_u) { over = 1; n = cap_u; }
      s_off[sp] = tot;
      tot += n;
    }
    s_off[splits] = tot;
    s_over = over | (tot > cap ? 1 : 0);
  }
  __syncthreads();
  const int total = s_off[splits];
  const int cnt = total < cap ? total : cap;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float (*tile)[33] = s_tile[warp];
  for (int i0 = warp * 32; i0 < cnt; i0 += nw * 32) {
    const int i = i0 + lane;
    int my_id = -1;
    if (i < cnt) {
      int lo = 0, hi = splits;              // largest sp with s_off[sp] <= i
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= i) lo = mid; else hi = mid;
      }
      my_id = seg_ids[(c * splits + lo) * cap_u + (i - s_off[lo])];
    }
    float acc = 0.f;
    for (int kc = 0; kc < K; kc += 32) {
      const int kk = kc + lane;
#pragma unroll
      for (int r = 0; r < 32; r++) {
        const int id = __shfl_sync(0xffffffffu, my_id, r);
        tile[r][lane] = (id >= 0 && kk < K) ? __ldg(items + (int64_t)id * K + kk) : 0.f;
      }
      __syncwarp();
      const int kn = min(32, K - kc);
      for (int j = 0; j < kn; j++) {
        const float v = tile[lane][j];
        const float x = fm ? __fadd_rn(v, s_q[K + kc + j]) : v;
        const float p = __fmul_rn(s_q[kc + j], x);
        acc = (kc + j == 0) ? p : __fadd_rn(acc, p);
      }
      __syncwarp();
    }
    if (i < cnt) {
      cand_scores[c * cap + i] = (fm && item_bias) ? __fadd_rn(__ldg(item_bias + my_id), acc) : acc;
      cand_ids[c * cap + i] = my_id;
    }
  }
  if (threadIdx.x == 0) {
    overflow[c] = s_over;
    cand_cnt[c] = cnt;
  }
}

// ---------------------------------------------------------------------------------------------------
// exact rescoring, second version: dense candidate lists + a flat work list, then a persistent kernel in which every
// warp is an independent pipeline over work items (row, batch of 32 candidates): the 32 item rows and the query row(s)
// of the NEXT item are in flight as bulk copies (cp.async.bulk, one mbarrier per stage) while the current item is
// scored from shared memory, lane r walking candidate r in ascending k (canonical order, bit-identical to the oracle).
// The first version took 32 coalesced loads per 32-k chunk with the warp stalled on each group (ncu: long scoreboard
// 29 stall cycles per issue, 8.5 % of DRAM peak).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_compact_kernel(int nseg, int cap_u, int cap, const int32_t* __restrict__ seg_ids,
                                                         const int32_t* __restrict__ seg_cnt, int32_t* __restrict__ cand_ids,
                                                         int32_t* __restrict__ cand_cnt, int32_t* __restrict__ overflow,
                                                         int32_t* __restrict__ work, int32_t* __restrict__ n_work) {
  __shared__ int s_off[512];
  __shared__ int s_warp_tot[4];
  __shared__ int s_over;
  const int64_t c = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  if (t == 0) s_over = 0;
  __syncthreads();
  // exclusive prefix of the clamped segment counts: 4 entries per thread, warp scan, then the 4 warp totals
  int v[4], sum = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int sgi = t * 4 + q;
    int n = sgi < nseg ? seg_cnt[c * nseg + sgi] : 0;
    if (n > cap_u) { n = cap_u; s_over = 1; }
    v[q] = n;
    sum += n;
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += x;
  }
  if (lane == 31) s_warp_tot[w] = incl;
  __syncthreads();
  int base = incl - sum;
  for (int i = 0; i < w; i++) base += s_warp_tot[i];
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int sgi = t * 4 + q;
    if (sgi < nseg) {
      const int32_t* src = seg_ids + (c * nseg + sgi) * cap_u;
      for (int k = 0; k < v[q]; k++)
        if (base + k < cap) cand_ids[c * cap + base + k] = src[k];
    }
    base += v[q];
  }
  if (t == 127) {
    const int total = base;
    const int cnt = total < cap ? total : cap;
    cand_cnt[c] = cnt;
    overflow[c] = (s_over || total > cap) ? 1 : 0;
    const int nb = (cnt + 31) / 32;
    const int w0 = atomicAdd(n_work, nb);
    for (int b = 0; b < nb; b++) work[w0 + b] = (int32_t)((c << 6) | b);
  }
}

struct RescoreArgs {
  int kind;
  const float* Q;
  const float* Fc;
  const float* items;
  const float* item_bias;
  int K, cap;
  const int32_t* cand_ids;
  const int32_t* cand_cnt;
  const int32_t* work;
  const int32_t* n_work;
  float* cand_scores;
};

__host__ __device__ inline size_t rescore_stage_bytes(int K, int fm) { return (size_t)(32 + 1 + (fm ? 1 : 0)) * (K + 4) * 4; }

__global__ void __launch_bounds__(256, 1) tc_rescore_staged_kernel(const RescoreArgs a, int warps) {
  extern __shared__ __align__(128) unsigned char rs_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp >= warps) return;
  const int K = a.K, kv = K >> 2, fm = a.kind == HHFM_QUERY_FM, ld = K + 4;
  const size_t stage_b = rescore_stage_bytes(K, fm);
  unsigned char* base = rs_smem + (size_t)warp * (2 * stage_b + 128);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * stage_b);
  if (lane == 0) {
    mbar_init(bars + 0, 1);
    mbar_init(bars + 1, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncwarp();
  const int n_work = __ldg(a.n_work);
  const int tw = gridDim.x * warps;
  const int gw = blockIdx.x * warps + warp;
  const uint32_t row_bytes = (uint32_t)K * 4u;

  // software pipeline (registers): the work word two items ahead, the candidate id one item ahead
  auto load_work = [&](int i) { return i < n_work ? __ldg(a.work + i) : -1; };
  auto load_id = [&](int wk) {
    if (wk < 0) return -1;
    const int64_t row = wk >> 6;
    const int i = (wk & 63) * 32 + lane;
    return i < __ldg(a.cand_cnt + row) ? __ldg(a.cand_ids + row * a.cap + i) : -1;
  };
  auto issue = [&](int wk, int id, int st) {          // bulk copies of item `wk` into stage st
    float* sb = reinterpret_cast<float*>(base + (size_t)st * stage_b);
    const int64_t row = wk >> 6;
    const unsigned valid = __ballot_sync(0xffffffffu, id >= 0);
    if (lane == 0) mbar_expect_tx(bars + st, row_bytes * (uint32_t)(__popc(valid) + 1 + fm));
    __syncwarp();
    if (id >= 0) bulk_copy_1d(sb + (size_t)(1 + fm + lane) * ld, a.items + (int64_t)id * K, row_bytes, bars + st);
    if (lane == 0) bulk_copy_1d(sb, a.Q + row * K, row_bytes, bars + st);
    if (fm && lane == 1) bulk_copy_1d(sb + ld, a.Fc + row * K, row_bytes, bars + st);
  };

  int wk_cur = load_work(gw);
  int id_cur = load_id(wk_cur);
  int wk_next = load_work(gw + tw);
  if (wk_cur >= 0) issue(wk_cur, id_cur, 0);
  int id_next = load_id(wk_next);
  uint32_t phbits = 0u;      // bit st = parity of the next completion of stage st
  int st = 0;
  for (int i = gw; i < n_work; i += tw) {
    const int wk_next2 = load_work(i + 2 * tw);
    if (wk_next >= 0) issue(wk_next, id_next, st ^ 1);
    const int id_next2 = load_id(wk_next2);
    // ---- score item wk_cur from stage st ----
    mbar_wait(bars + st, (phbits >> st) & 1u, nullptr);
    phbits ^= 1u << st;
    const float* sb = reinterpret_cast<const float*>(base + (size_t)st * stage_b);
    const float4* q4 = reinterpret_cast<const float4*>(sb);
    const float4* f4 = reinterpret_cast<const float4*>(sb + ld);
    const float4* r4 = reinterpret_cast<const float4*>(sb + (size_t)(1 + fm + lane) * ld);
    if (id_cur >= 0) {
      float acc = 0.f;
      for (int k4 = 0; k4 < kv; k4++) {
        const float4 v = r4[k4], q = q4[k4];
        float4 x = v;
        if (fm) {
          const float4 f = f4[k4];
          x = make_float4(__fadd_rn(v.x, f.x), __fadd_rn(v.y, f.y), __fadd_rn(v.z, f.z), __fadd_rn(v.w, f.w));
        }
        const float p0 = __fmul_rn(q.x, x.x);
        acc = (k4 == 0) ? p0 : __fadd_rn(acc, p0);
        acc = __fadd_rn(acc, __fmul_rn(q.y, x.y));
        acc = __fadd_rn(acc, __fmul_rn(q.z, x.z));
        acc = __fadd_rn(acc, __fmul_rn(q.w, x.w));
      }
      const int64_t row = wk_cur >> 6;
      const int ci = (wk_cur & 63) * 32 + lane;
      a.cand_scores[row * a.cap + ci] = (fm && a.item_bias) ? __fadd_rn(__ldg(a.item_bias + id_cur), acc) : acc;
    }
    __syncwarp();                                   // the stage is re-armed two iterations from now
    wk_cur = wk_next; id_cur = id_next;
    wk_next = wk_next2; id_next = id_next2;
    st ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int round_up(int64_t x, int64_t m) { return (int)((x + m - 1) / m * m); }

struct TcPlan {
  int Kp, nkc, bn, stages;
  bool ok;
};

static TcPlan tc_plan(int kind, int64_t K) {
  TcPlan p{};
  const int64_t kk = K + (kind == HHFM_QUERY_FM ? 2 : 0);
  p.Kp = round_up(kk, kKC);
  p.nkc = p.Kp / kKC;
  p.ok = true;
  if (p.nkc == 1) { p.bn = 256; p.stages = 4; }
  else if (p.nkc == 2) { p.bn = 256; p.stages = 2; }
  else if (p.nkc == 3) { p.bn = 128; p.stages = 3; }
  else if (p.nkc == 4) { p.bn = 128; p.stages = 2; }
  else p.ok = false;
  return p;
}

static int make_tmap(CUtensorMap* map, const void* base, int64_t rows, int Kp, int box_rows) {
  EncodeTiledFn cuTensorMapEncodeTiled = encode_tiled_fn();
  if (cuTensorMapEncodeTiled == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return HHFM_ERR_LAUNCH;
  }
  cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
  cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return HHFM_ERR_LAUNCH;
  }
  return HHFM_OK;
}

template <int NKC, int BN, int STAGES, bool EMIT>
static int launch_tc(const CUtensorMap& tA, const CUtensorMap& tB, const TcArgs& a, cudaStream_t st) {
  constexpr int smem = TcSmem<NKC, BN, STAGES>::kBytes;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(tc_score_kernel<NKC, BN, STAGES, EMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      set_error("tc_score_kernel: cannot reserve %d bytes of shared memory", smem);
      return HHFM_ERR_LAUNCH;
    }
    attr_set = true;
  }
  const int grid = a.n_units < sm_count() ? a.n_units : sm_count();
  tc_score_kernel<NKC, BN, STAGES, EMIT><<<grid, kTcThreads, smem, st>>>(tA, tB, a);
  return check_launch("tc_score_kernel");
}

struct TcLayout {   // carve-up of the caller's workspace
  size_t off_A, off_qinfo, off_gmax, off_tau, off_thr, off_thrv, off_seg, off_segcnt, off_cs, off_ci, off_cc, off_work, off_nwork, off_err, total;
  int64_t gmax_stride;
  int n_groups, cap, cap_u, fine;
  int n_row_blocks, n_tiles, tiles_per_unit, splits, n_units;      // emission pass (and the max pass when not sampled)
  int sample_stride, rank_j, s_tiles, s_tiles_per_unit, s_splits, s_units;   // max pass
};

static void tc_decompose(int n_row_blocks, int n_tiles, int* splits_out, int* tiles_per_unit_out, int* units_out) {
  // unit = (row block of 128 contexts, contiguous range of item tiles), units are dealt round-robin to one persistent CTA
  // per SM.  ncu showed the first version (2 units per SM) idle 31 % of the time in a ragged last wave (320 units on
  // 148 SMs), so the split count is now searched between ~4 and ~12 units per SM for the smallest makespan
  // ceil(units / SMs) * tiles_per_unit.
  const int sms = sm_count();
  int lo = (4 * sms + n_row_blocks - 1) / n_row_blocks, hi = (12 * sms + n_row_blocks - 1) / n_row_blocks;
  if (lo < 1) lo = 1;
  if (hi > n_tiles) hi = n_tiles;
  if (hi > 112) hi = 112;                 // the candidate compaction handles at most 512 = kTcSubs * splits segments per row
  if (lo > hi) lo = hi;
  int64_t best = -1;
  int best_splits = lo, best_tpu = (n_tiles + lo - 1) / lo;
  for (int sp = lo; sp <= hi; sp++) {
    const int tpu = (n_tiles + sp - 1) / sp;
    const int real = (n_tiles + tpu - 1) / tpu;
    const int64_t units = (int64_t)n_row_blocks * real;
    const int64_t span = (units + sms - 1) / sms * tpu;
    if (best < 0 || span < best) { best = span; best_splits = real; best_tpu = tpu; }
  }
  *tiles_per_unit_out = best_tpu;
  *splits_out = best_splits;
  *units_out = n_row_blocks * best_splits;
}

static int tc_sample_stride_env() {
  const char* e = getenv("HHFM_TOPN_SAMPLE");       // 1 = never sample (two full passes), unset = 8 for large catalogs
  if (e == nullptr) return 8;
  const int v = atoi(e);
  return v < 1 ? 1 : (v > 64 ? 64 : v);
}

static TcLayout tc_layout(int64_t C, int64_t N, int Kp, int tp, int bn) {
  TcLayout L{};
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const int64_t n_tiles = (N + bn - 1) / bn;
  L.n_row_blocks = (int)((C + kBM - 1) / kBM);
  L.n_tiles = (int)n_tiles;
  tc_decompose(L.n_row_blocks, L.n_tiles, &L.splits, &L.tiles_per_unit, &L.n_units);
  // Max pass.  Large catalogs: only every `sample_stride`-th item tile is scored and the cut is the rank_j-th largest
  // 32-item group maximum of that sample, chosen so that ~3.5 tp items of the full catalog pass it (proven per row
  // afterwards, tc_verify_kernel).  Small catalogs: every tile, tau = tp-th largest group maximum (guaranteed cut).
  int s = tc_sample_stride_env();
  // expected survivors of the cut = x * tp, x = 3.5 (HHFM_TOPN_CUT_X10 = 10 x for A/B runs).  Measured at C = 16 384, N = 10^6,
  // tp = 100: x = 3.5 -> no row fails its proof, 4.7 ms; x = 2.5 -> 4 rows fail (the proof needs tp candidates above the cut
  // PLUS the bf16 error bound, which eats most of the margin) and their exact redo costs 2 ms; x = 2.0 -> 126 rows.
  static const int x10 = [] { const char* e = getenv("HHFM_TOPN_CUT_X10"); const int v = e ? atoi(e) : 35; return v < 15 ? 15 : (v > 80 ? 80 : v); }();
  int j = (int)((x10 * (int64_t)tp + 10 * s - 1) / (10 * s)) + 4;
  if (s > 1 && (n_tiles / s) * (bn / kGroup) < 8 * (int64_t)j) s = 1;
  if (j > 1024) s = 1;
  L.sample_stride = s;
  L.rank_j = s > 1 ? j : tp;
  L.s_tiles = (int)((n_tiles + s - 1) / s);
  L.fine = (s > 1) || n_tiles < 4 * (int64_t)tp;      // one maximum per 32 items for a tight tau
  if (L.fine) {
    L.n_groups = s > 1 ? L.s_tiles * (bn / kGroup) : (int)((N + kGroup - 1) / kGroup);
    L.gmax_stride = (int64_t)L.s_tiles * (bn / kGroup);
  } else {
    L.n_groups = kTcSubs * (int)n_tiles;              // one maximum per quarter tile
    L.gmax_stride = (int64_t)kTcSubs * n_tiles;
  }
  tc_decompose(L.n_row_blocks, L.s_tiles, &L.s_splits, &L.s_tiles_per_unit, &L.s_units);
  L.cap = (s > 1 ? 8 : 4) * tp + 256;                 // survivors per row (expected ~2 tp, ~3.5 tp when sampled)
  L.cap_u = (4 * L.cap) / (kTcSubs * L.splits) + 32;  // per (row, split, column quarter) segment: 4x the even share + slack
  if (L.cap_u > L.cap) L.cap_u = L.cap;
  size_t o = 0;
  L.off_A = o; o = align(o + (size_t)C * Kp * 2);
  L.off_qinfo = o; o = align(o + (size_t)C * 16);
  L.off_gmax = o; o = align(o + (size_t)C * L.gmax_stride * 4);
  L.off_tau = o; o = align(o + (size_t)C * L.rank_j * 4 * 2);      // select writes [C,rank_j] scores + ids
  L.off_thr = o; o = align(o + (size_t)C * 4);
  L.off_thrv = o; o = align(o + (size_t)C * 4);
  L.off_seg = o; o = align(o + (size_t)C * kTcSubs * L.splits * L.cap_u * 4);
  L.off_segcnt = o; o = align(o + (size_t)C * kTcSubs * L.splits * 4);
  L.off_cs = o; o = align(o + (size_t)C * L.cap * 4);
  L.off_ci = o; o = align(o + (size_t)C * L.cap * 4);
  L.off_cc = o; o = align(o + (size_t)C * 4);
  L.off_work = o; o = align(o + (size_t)C * ((L.cap + 31) / 32) * 4);
  L.off_nwork = o; o = align(o + 256);
  L.off_err = o; o = align(o + 256);
  L.total = o;
  return L;
}

template <bool EMIT>
static int run_gemm(const TcPlan& p, const CUtensorMap& tA, const CUtensorMap& tB, const TcArgs& a, cudaStream_t st) {
  if (p.nkc == 1) return launch_tc<1, 256, 4, EMIT>(tA, tB, a, st);
  if (p.nkc == 2) return launch_tc<2, 256, 2, EMIT>(tA, tB, a, st);
  if (p.nkc == 3) return launch_tc<3, 128, 3, EMIT>(tA, tB, a, st);
  return launch_tc<4, 128, 2, EMIT>(tA, tB, a, st);
}

static TcArgs make_tc_args(const TcPlan& p, const TcLayout& L, int64_t C, int64_t N, uint8_t* ws, bool max_pass) {
  TcArgs a{};
  a.C = C; a.N = N;
  a.n_row_blocks = L.n_row_blocks;
  if (max_pass) {
    a.n_tiles = L.s_tiles; a.tiles_per_unit = L.s_tiles_per_unit; a.splits = L.s_splits; a.n_units = L.s_units;
    a.tile_stride = L.sample_stride;
  } else {
    a.n_tiles = L.n_tiles; a.tiles_per_unit = L.tiles_per_unit; a.splits = L.splits; a.n_units = L.n_units;
    a.tile_stride = 1;
  }
  a.gmax = reinterpret_cast<float*>(ws + L.off_gmax);
  a.gmax_stride = L.gmax_stride;
  a.fine = L.fine;
  a.thr_emit = reinterpret_cast<const float*>(ws + L.off_thr);
  a.seg_ids = reinterpret_cast<int32_t*>(ws + L.off_seg);
  a.seg_cnt = reinterpret_cast<int32_t*>(ws + L.off_segcnt);
  a.cap_u = L.cap_u;
  a.err = reinterpret_cast<int*>(ws + L.off_err);
  return a;
}

}  // namespace hhfm

using namespace hhfm;

extern "C" int hhfm_topn_select(const float* scores, const int32_t* ids, const int32_t* counts, int64_t C,
                                int64_t row_stride, int64_t n, int32_t tp, int32_t id_offset, float* out_scores,
                                int32_t* out_ids, hhfm_stream_t stream);

extern "C" int hhfm_topn_tc_supported(int32_t kind, int64_t N, int64_t K, int32_t tp) {
  if (kind < 0 || kind > 2 || K <= 0 || tp < 1 || tp > 1024) return 0;
  if (!tc_plan(kind, K).ok) return 0;
  return ((N + kGroup - 1) / kGroup) >= tp ? 1 : 0;     // need at least tp group maxima for the threshold (fine mode)
}

extern "C" int64_t hhfm_topn_tc_item_operand_bytes(int32_t kind, int64_t N, int64_t K) {
  TcPlan p = tc_plan(kind, K);
  return p.ok ? (int64_t)N * p.Kp * 2 : 0;
}

extern "C" int64_t hhfm_workspace_bytes_topn(int32_t kind, int64_t C, int64_t N, int64_t K, int32_t tp) {
  TcPlan p = tc_plan(kind, K);
  if (!p.ok) return 0;
  return (int64_t)tc_layout(C, N, p.Kp, tp, p.bn).total;
}

extern "C" int hhfm_topn_tc_prepare_items(int32_t kind, const float* items, const float* item_bias, int64_t N, int64_t K,
                                          void* item_operand, float* stats, hhfm_stream_t stream) {
  HHFM_REQUIRE(items && item_operand && stats && N > 0, "topn_tc_prepare_items: NULL argument");
  TcPlan p = tc_plan(kind, K);
  HHFM_REQUIRE(p.ok, "topn_tc_prepare_items: K=%lld unsupported by the tensor-core path", (long long)K);
  HHFM_REQUIRE(((uintptr_t)item_operand & 127) == 0, "topn_tc_prepare_items: item_operand must be 128-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(stats, 0, 2 * sizeof(float), st);
  tc_prep_items_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(items, kind == HHFM_QUERY_FM ? item_bias : nullptr, N, (int)K,
                                                              p.Kp, reinterpret_cast<__nv_bfloat16*>(item_operand), stats);
  return check_launch("tc_prep_items_kernel");
}

// Stage 1-3: filter.  Leaves gmax / tau / qinfo in the workspace for hhfm_topn_rescore_merge.
extern "C" int hhfm_topn_score(int32_t kind, const float* Q, const float* Fc, int64_t C, const void* item_operand,
                               int64_t N, int64_t K, int32_t tp, const float* stats, void* workspace,
                               int64_t workspace_bytes, hhfm_stream_t stream) {
  HHFM_REQUIRE(Q && item_operand && workspace && stats, "topn_score: NULL argument");
  HHFM_REQUIRE(kind != HHFM_QUERY_FM || Fc, "topn_score: FM needs Fc");
  HHFM_REQUIRE(hhfm_topn_tc_supported(kind, N, K, tp), "topn_score: configuration not supported by the tensor-core path");
  HHFM_REQUIRE(C > 0, "topn_score: C must be > 0");
  TcPlan p = tc_plan(kind, K);
  TcLayout L = tc_layout(C, N, p.Kp, tp, p.bn);
  HHFM_REQUIRE(workspace_bytes >= (int64_t)L.total, "topn_score: workspace too small (%lld < %lld)", (long long)workspace_bytes,
               (long long)L.total);
  HHFM_REQUIRE(((uintptr_t)workspace & 255) == 0, "topn_score: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(ws + L.off_A);
  float4* qinfo = reinterpret_cast<float4*>(ws + L.off_qinfo);
  float* tau_sc = reinterpret_cast<float*>(ws + L.off_tau);
  const int fm = kind == HHFM_QUERY_FM;
  cudaMemsetAsync(ws + L.off_err, 0, sizeof(int), st);
  tc_prep_queries_kernel<<<(unsigned)((C + 7) / 8), 256, 0, st>>>(Q, Fc, C, (int)K, p.Kp, fm, A, qinfo);
  int rc = check_launch("tc_prep_queries_kernel");
  if (rc) return rc;
  CUtensorMap tA, tB;
  if ((rc = make_tmap(&tA, A, C, p.Kp, kBM))) return rc;
  if ((rc = make_tmap(&tB, item_operand, N, p.Kp, p.bn))) return rc;
  // max pass (sampled for large catalogs) -> cut -> emission pass over the whole catalog
  const int32_t rj = L.rank_j;
  int32_t* tau_idj = reinterpret_cast<int32_t*>(ws + L.off_tau + (size_t)C * rj * 4);
  TcArgs a1 = make_tc_args(p, L, C, N, ws, true);
  if ((rc = run_gemm<false>(p, tA, tB, a1, st))) return rc;
  if (L.sample_stride > 1 && (size_t)L.n_groups * 4 <= 96 * 1024) {
    const size_t smem = (size_t)L.n_groups * sizeof(float);
    if (smem > 48 * 1024) cudaFuncSetAttribute(tc_cut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_cut_kernel<<<(unsigned)C, kCutThreads, smem, st>>>(a1.gmax, L.gmax_stride, L.n_groups, rj, fm, (int)K, qinfo, stats,
                                                         reinterpret_cast<float*>(ws + L.off_thr),
                                                         reinterpret_cast<float*>(ws + L.off_thrv));
    if ((rc = check_launch("tc_cut_kernel"))) return rc;
  } else {
    if ((rc = hhfm_topn_select(a1.gmax, nullptr, nullptr, C, L.gmax_stride, L.n_groups, rj, 0, tau_sc, tau_idj, stream))) return rc;
    tc_threshold_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(C, fm, (int)K, tau_sc + (rj - 1), rj, qinfo, stats,
                                                                    L.sample_stride > 1 ? 1 : 0,
                                                                    reinterpret_cast<float*>(ws + L.off_thr),
                                                                    reinterpret_cast<float*>(ws + L.off_thrv));
    if ((rc = check_launch("tc_threshold_kernel"))) return rc;
  }
  TcArgs a2 = make_tc_args(p, L, C, N, ws, false);
  return run_gemm<true>(p, tA, tB, a2, st);
}

// Stage 5: exact rescoring of the survivors + final (score desc, id asc) selection.
extern "C" int hhfm_topn_rescore_merge(int32_t kind, const float* Q, const float* Fc, int64_t C, const float* items,
                                       const float* item_bias, int64_t N, int64_t K, int32_t tp, int32_t id_offset,
                                       void* workspace, int64_t workspace_bytes, float* out_scores, int32_t* out_ids,
                                       int32_t* overflow, hhfm_stream_t stream) {
  HHFM_REQUIRE(Q && items && workspace && out_ids && overflow, "topn_rescore_merge: NULL argument");
  TcPlan p = tc_plan(kind, K);
  HHFM_REQUIRE(p.ok, "topn_rescore_merge: unsupported K");
  TcLayout L = tc_layout(C, N, p.Kp, tp, p.bn);
  HHFM_REQUIRE(workspace_bytes >= (int64_t)L.total, "topn_rescore_merge: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* cs = reinterpret_cast<float*>(ws + L.off_cs);
  int32_t* ci = reinterpret_cast<int32_t*>(ws + L.off_ci);
  int32_t* cc = reinterpret_cast<int32_t*>(ws + L.off_cc);
  const int fm = kind == HHFM_QUERY_FM;
  HHFM_REQUIRE(kTcSubs * L.splits <= 511, "topn_rescore_merge: too many item splits");
  int rc;
  const size_t stage_b = rescore_stage_bytes((int)K, fm);
  if (K % 4 == 0 && (((uintptr_t)items | (uintptr_t)Q | (uintptr_t)Fc) & 15) == 0 && 2 * stage_b + 128 <= (size_t)200 * 1024 &&
      L.cap <= 2048) {
    int32_t* n_work = reinterpret_cast<int32_t*>(ws + L.off_nwork);
    cudaMemsetAsync(n_work, 0, sizeof(int32_t), st);
    tc_compact_kernel<<<(unsigned)C, 128, 0, st>>>(kTcSubs * L.splits, L.cap_u, L.cap, reinterpret_cast<const int32_t*>(ws + L.off_seg),
                                                  reinterpret_cast<const int32_t*>(ws + L.off_segcnt), ci, cc, overflow,
                                                  reinterpret_cast<int32_t*>(ws + L.off_work), n_work);
    if ((rc = check_launch("tc_compact_kernel"))) return rc;
    int warps = (int)(((size_t)200 * 1024) / (2 * stage_b + 128));
    if (warps > 8) warps = 8;
    const size_t smem = (size_t)warps * (2 * stage_b + 128);
    if (cudaFuncSetAttribute(tc_rescore_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("tc_rescore_staged_kernel: cannot reserve %d bytes of shared memory", (int)smem);
      return HHFM_ERR_LAUNCH;
    }
    RescoreArgs ra{kind, Q, Fc, items, fm ? item_bias : nullptr, (int)K, L.cap, ci, cc,
                   reinterpret_cast<const int32_t*>(ws + L.off_work), n_work, cs};
    tc_rescore_staged_kernel<<<sm_count(), warps * 32, smem, st>>>(ra, warps);
    rc = check_launch("tc_rescore_staged_kernel");
  } else {
    const size_t smem = (size_t)K * (fm ? 2 : 1) * sizeof(float);
    tc_rescore_kernel<<<(unsigned)C, 256, smem, st>>>(kind, Q, Fc, items, fm ? item_bias : nullptr, N, (int)K, kTcSubs * L.splits, L.cap_u,
                                                      reinterpret_cast<const int32_t*>(ws + L.off_seg),
                                                      reinterpret_cast<const int32_t*>(ws + L.off_segcnt), L.cap, cs, ci, cc,
                                                      overflow);
    rc = check_launch("tc_rescore_kernel");
  }
  if (rc) return rc;
  if (L.sample_stride > 1) HHFM_REQUIRE(out_scores != nullptr, "topn_rescore_merge: out_scores is required (sampled cut verification)");
  rc = hhfm_topn_select(cs, ci, cc, C, L.cap, L.cap, tp, id_offset, out_scores, out_ids, stream);
  if (rc || L.sample_stride == 1) return rc;
  tc_verify_kernel<<<(unsigned)((C + 255) / 256), 256, 0, st>>>(C, tp, out_scores, cc,
                                                               reinterpret_cast<const float*>(ws + L.off_thrv), overflow);
  return check_launch("tc_verify_kernel");
}
