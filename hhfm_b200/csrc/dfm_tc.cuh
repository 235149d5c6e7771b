// Internal interface of the 3xTF32 tensor-core GEMM used by the DeepFM tower (dfm_tc.cu); see that file for the scheme.
#pragma once
#include "common.cuh"

namespace hhfm {

enum { TF_EPI_STORE = 0, TF_EPI_BIAS_RELU = 1, TF_EPI_MASK = 2, TF_EPI_ATOMIC = 3, TF_EPI_SCATTER = 4, TF_EPI_BIAS_RELU_PROJ = 5 };

// C[M,N] (epilogue) = A[M,K] . B[N,K]^T, all row-major with the K index contiguous; *_lo = x - tf32(x) of the same shape.
struct TfGemm {
  const float* A;
  const float* A_lo;
  const float* B;
  const float* B_lo;
  int M, N, K;
  int64_t lda, ldb;
  int k_blocked;         // 1: A and B are k-blocked panels [ceil(K/32)][M or N][32] (tf_split_transpose's XT layout)
  int epi;
  float* C;              // EPI_SCATTER: the embedding gradient table
  int64_t ldc;
  const float* bias;     // EPI_BIAS_RELU, EPI_BIAS_RELU_PROJ
  const float* proj;     // EPI_BIAS_RELU_PROJ [N]: instead of relu(A.B^T + bias) itself, each thread stores its share of
                         // relu(...) . proj: C[m, nt*4 + g] for n-tile nt and column group g (ldc >= tf_proj_partials(N));
                         // the consumer adds the partials in index order (deterministic, no atomics)
  const float* mask;     // EPI_MASK (may alias C)
  int64_t ldmask;
  const int32_t* idx;    // EPI_SCATTER: [M, F] ids; output column n belongs to row idx[m, n / Kemb], element n % Kemb
  int F, Kemb;
  HotPlan hot;           // EPI_SCATTER: optional hot-row replicas
  float* scratch;        // EPI_ATOMIC (split-K): partial tiles, tf_splitk_scratch_floats() floats
  int64_t scratch_floats;
};

int tf_gemm(const TfGemm& g, cudaStream_t st);
int64_t tf_splitk_scratch_floats();
int tf_proj_partials(int N);      // partial sums per output row written by EPI_BIAS_RELU_PROJ
// mask != NULL: v = X * (mask > 0) is what gets split / transposed, and it is also written to Xout (may alias mask)
int tf_split_transpose(const float* X, int64_t rows, int cols, int64_t ld, float* Xlo, float* XT, float* XTlo, cudaStream_t st,
                       const float* mask = nullptr, float* Xout = nullptr);
int tf_gather_split_transpose(const int32_t* idx, int64_t B, int F, int K, const float* V, float* X0, int64_t ld, float* Xlo,
                              float* XT, float* XTlo, cudaStream_t st);
int tf_gather_x0(const int32_t* idx, int64_t B, int F, int K, const float* V, float* X0, int64_t ld, cudaStream_t st);
int tf_colsum(const float* X, int64_t rows, int cols, int64_t ld, float* out, cudaStream_t st);
int tf_prep_weight(const float* W, int rows, int cols, float* Wp, float* Wplo, int ldp, float* WT, float* WTlo, int ldtp,
                   cudaStream_t st);

}  // namespace hhfm
