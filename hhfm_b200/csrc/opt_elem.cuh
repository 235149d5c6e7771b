// The element update of the TF-1.x optimizers (K5), shared by the optimizer kernels (opt.cu) and the in-place path of the staged
// scatter kernels (single-touch rows, fm.cu / pairrank.cu) so that both produce the same bits.
#pragma once
#include "common.cuh"

namespace hhfm {

struct OptP {
  float lr, lamda, p1, p2, p3;   // adam: p1=beta1 p2=beta2 p3=eps ; momentum: p1=mu
};

template <int KIND>
__device__ __forceinline__ void opt_elem(float& w, float& s1, float& s2, float g, const OptP& p) {
  if (KIND == HHFM_OPT_ADAGRAD) {
    s1 = fmaf(g, g, s1);                                // explicit contractions: the same bits as p2p.cu::opt_elem2
    w = w - p.lr * g / sqrtf(s1);
  } else if (KIND == HHFM_OPT_ADAM) {
    s1 = fmaf(p.p1, s1, (1.f - p.p1) * g);
    s2 = fmaf(p.p2, s2, (1.f - p.p2) * (g * g));
    w = w - p.lr * s1 / (sqrtf(s2) + p.p3);
  } else if (KIND == HHFM_OPT_MOMENTUM) {
    s1 = fmaf(s1, p.p1, g);
    w = fmaf(-p.lr, s1, w);
  } else {
    w = fmaf(-p.lr, g, w);
  }
}

template <int KIND>
__device__ __forceinline__ void opt_vec4(float4& w, float4& s1, float4& s2, float4 g, const OptP& p) {
  opt_elem<KIND>(w.x, s1.x, s2.x, g.x, p);
  opt_elem<KIND>(w.y, s1.y, s2.y, g.y, p);
  opt_elem<KIND>(w.z, s1.z, s2.z, g.z, p);
  opt_elem<KIND>(w.w, s1.w, s2.w, g.w, p);
}

// Rows that exactly ONE sample of the step references (hhfm_count_refs): the staged scatter kernels apply the optimizer step
// of such a row on the spot -- gradient from registers, the row's weights from the gather that is already in shared memory,
// one read of its accumulator -- instead of a read-modify-write of its line in the gradient arena followed by the rows
// optimizer's six row transfers.  The row is not stamped, so the touched-row optimizer never sees it.  Adagrad and SGD only
// (an untouched row does not move under either).  Nobody else reads the row in this step, so the in-place write races with
// nothing.
struct SingleTouch {
  const int32_t* ref_count;   // [M] references per row in this step; nullptr = off
  float* V;                   // the embedding table, writable
  float* acc;                 // Adagrad accumulator [M, K]; unused for SGD
  float* bias;                // FM: feature_bias [M], writable; nullptr = the model has none
  float* bias_acc;
  float lr;
  int kind;                   // HHFM_OPT_ADAGRAD | HHFM_OPT_SGD
};

constexpr int kSingleTouchCap = 3;    // at most this many rows of a sample go the in-place way (their accumulators sit in registers)

// w (the gathered row chunk) and g -> updated chunk stored to the table, accumulator chunk `a` (already loaded) stored back
__device__ __forceinline__ void single_touch_apply(const SingleTouch& s, size_t off, float4 w, float4 a, float4 g) {
  const OptP p{s.lr, 0.f, 0.f, 0.f, 0.f};
  float4 z = f4_zero();
  if (s.kind == HHFM_OPT_ADAGRAD) {
    opt_vec4<HHFM_OPT_ADAGRAD>(w, a, z, g, p);
    *reinterpret_cast<float4*>(s.acc + off) = a;
  } else {
    opt_vec4<HHFM_OPT_SGD>(w, a, z, g, p);
  }
  *reinterpret_cast<float4*>(s.V + off) = w;
}

// The in-place steps of a sample are DEFERRED by one sample: the kernel issues the accumulator loads when it meets the rows and
// applies the update at the start of the warp's next sample, when the loads have long landed -- the consumer never waits on DRAM.
// One float4 chunk per lane (K <= 128).  All members are registers of the warp's lanes; n is warp-uniform.
struct PendingRows {
  int n;
  int id[kSingleTouchCap];
  float4 w[kSingleTouchCap], g[kSingleTouchCap], a[kSingleTouchCap];
};

__device__ __forceinline__ void pending_flush(const SingleTouch& s, PendingRows& pr, int K, int kv, int lane) {
#pragma unroll
  for (int q = 0; q < kSingleTouchCap; q++)
    if (q < pr.n && lane < kv) single_touch_apply(s, (size_t)pr.id[q] * K + 4 * lane, pr.w[q], pr.a[q], pr.g[q]);
  pr.n = 0;
}

// host side (fm.cu): validates a plan coming through the C-ABI and converts it
int single_touch_from_abi(const void* abi_plan, const float* V, int64_t K, SingleTouch* out);

}  // namespace hhfm
