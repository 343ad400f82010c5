// K6 host side: ggad_dense_matmul -- the dense projections of the GGAD path.
//   large, 16-byte-aligned problems -> tcgen05 fp32-accurate GEMM (dense_cutlass.cuh; UTCHMMA / LDTM / UTMALDG)
//   everything else (the h/4 -> 1 score layer, 200 x 20 x 64 mini-batch blocks, odd leading dimensions)
//                                    -> a plain register-tiled SIMT FFMA kernel (exact fp32)
// Roofline: tensor-bound for the full-batch layers (C1 layer 1: 7535 x 745 x 300 = 3.4 GFLOP), latency-bound
// for the mini-batch blocks.
#include "common.cuh"

namespace ggad {

int fast_f32_tn(int, int, int, const float*, int64_t, const float*, int64_t, float*, int64_t, float, float, int, void*, size_t,
                size_t*, cudaStream_t);
int fast_f32_nn(int, int, int, const float*, int64_t, const float*, int64_t, float*, int64_t, float, float, int, void*, size_t,
                size_t*, cudaStream_t);
int fast_f32_nt(int, int, int, const float*, int64_t, const float*, int64_t, float*, int64_t, float, float, int, void*, size_t,
                size_t*, cudaStream_t);

// C[m,n] = act(alpha * sum_k A(m,k) B(k,n) + beta * C[m,n]) with arbitrary element strides:
// A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn].  64 x 64 tile, 16-deep k slices, 4 x 4 per thread.
constexpr int kBM = 64, kBN = 64, kBK = 16;
__global__ void __launch_bounds__(256) dense_simt_kernel(int M, int N, int K, const float* __restrict__ A, int64_t sam, int64_t sak,
                                                         const float* __restrict__ B, int64_t sbk, int64_t sbn,
                                                         float* __restrict__ C, int64_t ldc, float alpha, float beta, int relu) {
  __shared__ float sA[kBK][kBM + 4];
  __shared__ float sB[kBK][kBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kBK) {
    for (int i = tid; i < kBM * kBK; i += 256) {
      // consecutive threads walk the contiguous direction of each operand
      const int mm = (sak == 1) ? i / kBK : i % kBM, kk = (sak == 1) ? i % kBK : i / kBM;
      const int gm = m0 + mm, gk = k0 + kk;
      sA[kk][mm] = (gm < M && gk < K) ? __ldg(A + gm * sam + gk * sak) : 0.f;
    }
    for (int i = tid; i < kBN * kBK; i += 256) {
      const int nn = (sbk == 1) ? i / kBK : i % kBN, kk = (sbk == 1) ? i % kBK : i / kBN;
      const int gn = n0 + nn, gk = k0 + kk;
      sB[kk][nn] = (gn < N && gk < K) ? __ldg(B + gk * sbk + gn * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sA[kk][tm]);
      const float4 b4 = *reinterpret_cast<const float4*>(&sB[kk][tn]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + tm + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tn + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      if (beta != 0.f) v = fmaf(beta, C[gm * ldc + gn], v);
      if (relu) v = fmaxf(v, 0.f);
      C[gm * ldc + gn] = v;
    }
  }
}

int dense_matmul_impl(int trans_a, int trans_b, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                      int64_t ldb, float* C, int64_t ldc, float alpha, float beta, int relu, int path, cudaStream_t st) {
  GGAD_REQUIRE(M >= 0 && N >= 0 && K >= 0 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), GGAD_ERR_INVALID,
               "dense_matmul: bad sizes");
  if (M == 0 || N == 0) return GGAD_OK;
  GGAD_REQUIRE(A && B && C, GGAD_ERR_INVALID, "dense_matmul: null pointer");
  GGAD_REQUIRE(lda >= (trans_a ? M : K) && ldb >= (trans_b ? K : N) && ldc >= N, GGAD_ERR_INVALID,
               "dense_matmul: leading dimension smaller than the row length");
  GGAD_REQUIRE(path >= 0 && path <= 2, GGAD_ERR_INVALID, "dense_matmul: path must be 0 (auto), 1 (SIMT) or 2 (tensor core)");
  // tensor-core path: TMA needs 16-byte aligned bases and row pitches; it supports A[M,K] K-major with either B
  // layout, or A stored transposed ([K,M], M contiguous) with B[K,N] N-contiguous (the weight-gradient GEMM)
  // (pitches AND the extents of the contiguous dimensions: N for C and an N-contiguous B, K for K-contiguous
  // operands, M for a transposed A)
  const bool aligned = aligned16(A) && aligned16(B) && aligned16(C) && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 &&
                       N % 4 == 0 && (trans_a ? M % 4 == 0 : K % 4 == 0) && (trans_b ? K % 4 == 0 : true);
  const bool layout_ok = !(trans_a && trans_b);
  const bool big = M >= 64 && N >= 16 && K >= 8 && M * N * K >= (1ll << 18);
  GGAD_REQUIRE(path != 2 || (aligned && layout_ok && K > 0), GGAD_ERR_UNSUPPORTED,
               "dense_matmul: the tensor-core path needs 16-byte aligned operands and not both operands transposed");
  if (K > 0 && aligned && layout_ok && (path == 2 || (path == 0 && big))) {
    auto fn = trans_a ? fast_f32_nt : (trans_b ? fast_f32_tn : fast_f32_nn);
    size_t need = 0;
    void* ws = nullptr;
    int rc = fn(int(M), int(N), int(K), A, lda, B, ldb, C, ldc, alpha, beta, relu, nullptr, 0, &need, st);
    if (rc == -2) {  // the kernel wants a workspace: take it from the library's stream-ordered pool
      GGAD_CUDA_OK(temp_alloc(&ws, need, st));
      rc = fn(int(M), int(N), int(K), A, lda, B, ldb, C, ldc, alpha, beta, relu, ws, need, &need, st);
      GGAD_CUDA_OK(cudaFreeAsync(ws, st));
    }
    GGAD_REQUIRE(rc == 0 || (rc == -1 && path == 0), GGAD_ERR_CUDA, "dense_matmul: tcgen05 GEMM failed at stage %d (%s)", -rc,
                 cudaGetErrorString(cudaGetLastError()));
    if (rc == 0) {
      count_launch(1);
      return GGAD_OK;
    }
    // auto mode and the tensor-core kernel declined the problem (can_implement): the SIMT kernel takes it
  }
  // SIMT: element strides of A(m,k) and B(k,n)
  const int64_t sam = trans_a ? 1 : lda, sak = trans_a ? lda : 1;
  const int64_t sbk = trans_b ? 1 : ldb, sbn = trans_b ? ldb : 1;
  dim3 grid((unsigned)((N + kBN - 1) / kBN), (unsigned)((M + kBM - 1) / kBM));
  GGAD_REQUIRE(grid.y <= 65535, GGAD_ERR_UNSUPPORTED, "dense_matmul: M too large for the SIMT path");
  dense_simt_kernel<<<grid, 256, 0, st>>>(int(M), int(N), int(K), A, sam, sak, B, sbk, sbn, C, ldc, alpha, beta, relu);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

}  // namespace ggad
