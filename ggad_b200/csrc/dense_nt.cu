// tcgen05 fp32-accurate GEMM (dense_cutlass.cuh), operand layouts A ColumnMajor / B RowMajor.
#include "dense_cutlass.cuh"

namespace ggad {
int fast_f32_nt(int M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, float alpha,
                float beta, int relu, void* ws, size_t ws_bytes, size_t* ws_needed, cudaStream_t st) {
  using LA = cutlass::layout::ColumnMajor;
  using LB = cutlass::layout::RowMajor;
  // the weight-gradient layout: always through the stream-K scheduler (it degrades to plain tiles when there are many)
  if (relu) return run_fast_f32<LA, LB, true, true>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, ws, ws_bytes, ws_needed, st);
  return run_fast_f32<LA, LB, false, true>(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, ws, ws_bytes, ws_needed, st);
}
}  // namespace ggad
