// Argument block shared by the AFM kernels (afm.cu: fp32 SIMT; afm_fused_tc.cu: fused tcgen05 training pass).
#pragma once
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

// afm_fused_tc.cu: returns 1 when the shape is not covered (the caller falls back to the SIMT kernels), else an HHFM_* code.
int dispatch_afm_fused_tc(const AfmArgs& a, int64_t M, cudaStream_t st);

}  // namespace hhfm
