// K7 on the device: the adjacency preprocessing of program A (utils.py:47-54 normalize_adj, run.py:98-101):
//   deg   = A.sum(1)                       fp64 row sums                       -> ggad_csr_row_sum_f64
//   A_hat = (A D)^T D + I,  D = diag(deg^-1/2)    fp64 products, ONE rounding to fp32   \  ggad_csr_scale_add_identity
//   R     = A + I                                                                        /  (on A^T with D, on A without)
// The products are taken in scipy's order -- (a * d[r]) * d[c] on the transposed CSR, then + 1.0 on the diagonal,
// then the cast -- and IEEE fp64 multiply/add are exact-rounded on both sides, so the fp32 values are bit-identical
// to the reference's dense adj / raw_adj (tested).  D itself comes from a host-built table indexed by the (integer)
// degree, because pow() is the one operation whose last bit differs between libm and CUDA.
#include <cub/cub.cuh>

#include "common.cuh"

namespace ggad {

__global__ void row_sum_f64_kernel(const int64_t* __restrict__ rowptr, const float* __restrict__ val, int64_t n_rows,
                                   double* __restrict__ out) {
  const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;  // one warp per row
  const int lane = threadIdx.x & 31;
  if (w >= n_rows) return;
  const int64_t s = rowptr[w], e = rowptr[w + 1];
  double acc = 0.0;
  if (val) {
    for (int64_t t = s + lane; t < e; t += 32) acc += double(val[t]);
  } else {
    acc = (lane == 0) ? double(e - s) : 0.0;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) out[w] = acc;
}

// position of the first column >= r in row r (columns sorted), and whether it IS r
__device__ __forceinline__ int64_t diag_pos(const int32_t* __restrict__ col, int64_t s, int64_t e, int64_t r, bool& has) {
  int64_t lo = s, hi = e;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (int64_t(col[mid]) < r) lo = mid + 1;
    else hi = mid;
  }
  has = lo < e && int64_t(col[lo]) == r;
  return lo;
}

__global__ void identity_count_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n,
                                      int64_t* __restrict__ out_len) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r > n) return;
  if (r == n) {
    out_len[r] = 0;
    return;
  }
  bool has;
  diag_pos(col, rowptr[r], rowptr[r + 1], r, has);
  out_len[r] = rowptr[r + 1] - rowptr[r] + (has ? 0 : 1);
}

__global__ void identity_fill_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                     const float* __restrict__ val, const double* __restrict__ scale, int64_t n,
                                     const int64_t* __restrict__ out_rowptr, int32_t* __restrict__ out_col,
                                     float* __restrict__ out_val) {
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;  // one warp per row
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  const int64_t s = rowptr[r], e = rowptr[r + 1], o = out_rowptr[r];
  bool has;
  const int64_t p = diag_pos(col, s, e, r, has);
  const double sr = scale ? scale[r] : 1.0;
  for (int64_t t = s + lane; t < e; t += 32) {
    const int32_t c = col[t];
    double v = val ? double(val[t]) : 1.0;
    if (scale) v = __dmul_rn(__dmul_rn(v, sr), scale[c]);   // (a * d[r]) * d[c]: scipy's order, no FMA contraction
    if (has && t == p) v = __dadd_rn(v, 1.0);               // + I on an existing diagonal entry
    const int64_t dst = o + (t - s) + ((!has && t >= p) ? 1 : 0);
    out_col[dst] = c;
    out_val[dst] = float(v);                                // the single rounding to fp32 (run.py:106-109)
  }
  if (!has && lane == 0) {
    out_col[o + (p - s)] = int32_t(r);
    out_val[o + (p - s)] = 1.0f;
  }
}

static inline unsigned blocks_for(int64_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

int csr_row_sum_f64_impl(const int64_t* rowptr, const float* val, int64_t n_rows, double* out, cudaStream_t st) {
  GGAD_REQUIRE(rowptr && out && n_rows >= 0, GGAD_ERR_INVALID, "csr_row_sum_f64: bad arguments");
  if (n_rows == 0) return GGAD_OK;
  row_sum_f64_kernel<<<blocks_for(n_rows * 32), 256, 0, st>>>(rowptr, val, n_rows, out);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

int csr_add_identity_rowptr_impl(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t* out_rowptr, int64_t* nnz_host,
                                 cudaStream_t st) {
  GGAD_REQUIRE(rowptr && out_rowptr && n >= 0 && (col || n == 0), GGAD_ERR_INVALID, "csr_add_identity_rowptr: bad arguments");
  int64_t* len = nullptr;
  GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&len), size_t(n + 1) * 8, st));
  identity_count_kernel<<<blocks_for(n + 1), 256, 0, st>>>(rowptr, col, n, len);
  GGAD_CUDA_OK(cudaGetLastError());
  size_t tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, len, out_rowptr, n + 1, st);
  void* tmp = nullptr;
  GGAD_CUDA_OK(temp_alloc(&tmp, tmp_bytes, st));
  GGAD_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, len, out_rowptr, n + 1, st));
  count_launch(2);
  GGAD_CUDA_OK(cudaFreeAsync(tmp, st));
  GGAD_CUDA_OK(cudaFreeAsync(len, st));
  if (nnz_host) {
    GGAD_CUDA_OK(cudaMemcpyAsync(nnz_host, out_rowptr + n, 8, cudaMemcpyDeviceToHost, st));
    GGAD_CUDA_OK(cudaStreamSynchronize(st));
  }
  return GGAD_OK;
}

int csr_scale_add_identity_impl(const int64_t* rowptr, const int32_t* col, const float* val, const double* scale, int64_t n,
                                const int64_t* out_rowptr, int32_t* out_col, float* out_val, cudaStream_t st) {
  GGAD_REQUIRE(rowptr && out_rowptr && out_col && out_val && n >= 0, GGAD_ERR_INVALID, "csr_scale_add_identity: bad arguments");
  if (n == 0) return GGAD_OK;
  identity_fill_kernel<<<blocks_for(n * 32), 256, 0, st>>>(rowptr, col, val, scale, n, out_rowptr, out_col, out_val);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

}  // namespace ggad
