// K6 host side: ggad_dense_matmul -- the dense projections of the GGAD path.
//   large, 16-byte-aligned problems -> tcgen05 fp32-accurate GEMM (dense_cutlass.cuh; UTCHMMA / LDTM / UTMALDG)
//   everything else (the h/4 -> 1 score layer, 200 x 20 x 64 mini-batch blocks, odd leading dimensions)
//                                    -> a plain register-tiled SIMT FFMA kernel (exact fp32)
// Roofline: tensor-bound for the full-batch layers (C1 layer 1: 7535 x 745 x 300 = 3.4 GFLOP), latency-bound
// for the mini-batch blocks.
#include <stdlib.h>

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
// Split-K (weight gradients: K = number of nodes, M x N = a few tiles): blockIdx.z owns the k range
// [z * k_chunk, (z+1) * k_chunk) and writes its partial tile to part[z][M][N]; splitk_reduce_kernel sums the slices
// in a fixed order (deterministic) and applies alpha / beta / ReLU.
__global__ void __launch_bounds__(256) dense_simt_kernel(int M, int N, int K, const float* __restrict__ A, int64_t sam, int64_t sak,
                                                         const float* __restrict__ B, int64_t sbk, int64_t sbn,
                                                         float* __restrict__ C, int64_t ldc, float alpha, float beta, int relu,
                                                         int k_chunk, float* __restrict__ part) {
  __shared__ float sA[kBK][kBM + 4];
  __shared__ float sB[kBK][kBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * kBM, n0 = blockIdx.x * kBN;
  const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
  float acc[4][4] = {};
  const int k_lo = part ? blockIdx.z * k_chunk : 0;
  const int k_hi = part ? ((k_lo + k_chunk < K) ? k_lo + k_chunk : K) : K;
  if (part) {
    C = part + int64_t(blockIdx.z) * M * N;
    ldc = N;
    alpha = 1.f;
    beta = 0.f;
    relu = 0;
  }
  for (int k0 = k_lo; k0 < k_hi; k0 += kBK) {
    for (int i = tid; i < kBM * kBK; i += 256) {
      // consecutive threads walk the contiguous direction of each operand
      const int mm = (sak == 1) ? i / kBK : i % kBM, kk = (sak == 1) ? i % kBK : i / kBM;
      const int gm = m0 + mm, gk = k0 + kk;
      sA[kk][mm] = (gm < M && gk < k_hi) ? __ldg(A + gm * sam + gk * sak) : 0.f;
    }
    for (int i = tid; i < kBN * kBK; i += 256) {
      const int nn = (sbk == 1) ? i / kBK : i % kBN, kk = (sbk == 1) ? i % kBK : i / kBN;
      const int gn = n0 + nn, gk = k0 + kk;
      sB[kk][nn] = (gn < N && gk < k_hi) ? __ldg(B + gk * sbk + gn * sbn) : 0.f;
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

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int M, int N, float* __restrict__ C, int64_t ldc,
                                     float alpha, float beta, int relu) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= int64_t(M) * N) return;
  float v = 0.f;
  for (int z = 0; z < splits; ++z) v += part[int64_t(z) * M * N + i];
  const int m = int(i / N), n = int(i % N);
  v *= alpha;
  if (beta != 0.f) v = fmaf(beta, C[m * ldc + n], v);
  if (relu) v = fmaxf(v, 0.f);
  C[m * ldc + n] = v;
}

int sm_count_cached();

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
  // few output tiles and a long k loop (dW = dY^T X: K = number of nodes) would leave most SMs idle: those go to the
  // split-K SIMT path below
  const int64_t tc_tiles = ((M + 127) / 128) * ((N + 127) / 128);
  // (measured, profiles/r02h_dense_projections.txt: the stream-K A^T B kernel wins from 300 x 300 outputs up --
  // C3 layer 2: 0.22 ms vs 0.65 ms SIMT / 0.29 ms cuBLAS -- and loses below, e.g. 64 x 20 x 8000: 0.20 vs 0.07 ms)
  // (... and again for very long k loops: 64 x 20 x 3.7 M runs 1.73 ms stream-K vs 4.16 ms SIMT split-K)
  const bool streamk_ok = trans_a && ((M >= 128 && N >= 128) || K >= (1ll << 18)) && !getenv("GGAD_DENSE_NO_STREAMK");
  const bool skinny = tc_tiles < 48 && K >= 4096 && !streamk_ok;
  const bool big = M >= 64 && N >= 8 && K >= 8 && M * N * K >= (1ll << 18) && !skinny && (!trans_a || streamk_ok || tc_tiles >= 48);
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
  const int64_t tiles = int64_t(grid.x) * grid.y;
  const int sms = sm_count_cached();
  int splits = 1;
  if (sms > 0 && tiles < 2 * sms && K >= 512) {
    splits = int((3ll * sms + tiles - 1) / tiles);                 // ~3 CTAs per SM in total
    if (splits > K / 128) splits = int(K / 128);
    if (splits > 256) splits = 256;
  }
  if (splits > 1) {
    int k_chunk = int((K + splits - 1) / splits);
    k_chunk = (k_chunk + kBK - 1) / kBK * kBK;
    splits = int((K + k_chunk - 1) / k_chunk);
    float* part = nullptr;
    GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&part), size_t(splits) * M * N * 4, st));
    grid.z = splits;
    dense_simt_kernel<<<grid, 256, 0, st>>>(int(M), int(N), int(K), A, sam, sak, B, sbk, sbn, C, ldc, alpha, beta, relu, k_chunk, part);
    GGAD_CUDA_OK(cudaGetLastError());
    splitk_reduce_kernel<<<(unsigned)((M * N + 255) / 256), 256, 0, st>>>(part, splits, int(M), int(N), C, ldc, alpha, beta, relu);
    GGAD_CUDA_OK(cudaGetLastError());
    GGAD_CUDA_OK(cudaFreeAsync(part, st));
    count_launch(2);
    return GGAD_OK;
  }
  dense_simt_kernel<<<grid, 256, 0, st>>>(int(M), int(N), int(K), A, sam, sak, B, sbk, sbn, C, ldc, alpha, beta, relu, 0, nullptr);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

}  // namespace ggad
