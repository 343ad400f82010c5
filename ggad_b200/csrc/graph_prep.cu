// K7: index / degree work on the device (CSR build, transpose, row extraction, histograms),
// the synthetic R-MAT generator, and the two small dense row kernels of the affinity path.
// Integer outputs are bit-exact with the CPU restatement (sorting is by exact integer keys).
// cub::DeviceRadixSort (shipped with the CUDA toolkit) does the one-off sorts; nothing here
// is on the per-step hot path.
#include <cub/cub.cuh>

#include "common.cuh"

namespace ggad {

// ---------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------
__global__ void rowptr_from_sorted_keys(const uint64_t* __restrict__ keys, int64_t n, int64_t n_rows,
                                        int64_t* __restrict__ rowptr) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  const uint64_t target = uint64_t(r) << 32;  // first key with row >= r
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1;
    else hi = mid;
  }
  rowptr[r] = lo;
}

__global__ void low32_of_keys(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = int32_t(uint32_t(keys[i] & 0xffffffffull));
}

__global__ void transpose_keys(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_rows,
                               int64_t nnz, uint64_t* __restrict__ keys, int64_t* __restrict__ idx) {
  const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  int64_t lo = 0, hi = n_rows;  // row r with rowptr[r] <= e < rowptr[r+1]
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (rowptr[mid + 1] <= e) lo = mid + 1;
    else hi = mid;
  }
  keys[e] = (uint64_t(uint32_t(col[e])) << 32) | uint64_t(uint32_t(lo));
  if (idx) idx[e] = e;
}

__global__ void gather_vals(const float* __restrict__ val, const int64_t* __restrict__ perm, int64_t n,
                            float* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = val[perm[i]];
}

__global__ void extract_rows_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                    const float* __restrict__ val, const int32_t* __restrict__ rows, int64_t n_sel,
                                    const int64_t* __restrict__ sub_rowptr, int32_t* __restrict__ sub_col,
                                    float* __restrict__ sub_val) {
  const int64_t w = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;  // one warp per selected row
  const int lane = threadIdx.x & 31;
  if (w >= n_sel) return;
  const int64_t r = rows[w];
  const int64_t s = rowptr[r], n = rowptr[r + 1] - s, o = sub_rowptr[w];
  for (int64_t t = lane; t < n; t += 32) {
    sub_col[o + t] = col[s + t];
    if (val) sub_val[o + t] = val[s + t];
  }
}

__global__ void col_hist_kernel(const int32_t* __restrict__ col, int64_t nnz, int32_t* __restrict__ counts) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < nnz; e += stride) atomicAdd(counts + col[e], 1);
}

// ---- mini-batch frontier (src/graphsage.py:305-311,335-341 as array ops on the device) ----
// degree of block row i = |N(nodes[i])| (+1 for the node itself when add_self and it is not already a neighbor)
__global__ void block_degree_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col, int64_t n_nodes,
                                    const int32_t* __restrict__ nodes, int64_t n_batch, int add_self,
                                    int64_t* __restrict__ deg) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i > n_batch) return;
  if (i == n_batch) {
    deg[i] = 0;  // scan sentinel
    return;
  }
  const int64_t v = nodes[i];
  int64_t d = 0;
  bool has_self = false;
  if (v >= 0 && v < n_nodes) {
    const int64_t s = rowptr[v], e = rowptr[v + 1];
    d = e - s;
    if (add_self) {  // neighbor lists are sorted: binary search for v
      int64_t lo = s, hi = e;
      while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (col[mid] < v) lo = mid + 1;
        else hi = mid;
      }
      has_self = lo < e && col[lo] == v;
    }
  }
  deg[i] = d + ((add_self && !has_self) ? 1 : 0);
}

// one 128-thread CTA per block row (a frontier holds hub rows with 10^5 neighbors: one warp per row made the launch
// as long as its longest row): copy the neighbor slice, append the node itself if it was not a neighbor
constexpr int kFillThreads = 128;
__global__ void __launch_bounds__(kFillThreads) block_fill_kernel(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                                  int64_t n_nodes, const int32_t* __restrict__ nodes, int64_t n_batch,
                                                                  int add_self, const int64_t* __restrict__ block_rowptr,
                                                                  int32_t* __restrict__ out) {
  const int64_t w = blockIdx.x;
  const int lane = threadIdx.x;
  if (w >= n_batch) return;
  const int64_t v = nodes[w];
  const int64_t o = block_rowptr[w], total = block_rowptr[w + 1] - o;
  int64_t s = 0, n = 0;
  if (v >= 0 && v < n_nodes) {
    s = rowptr[v];
    n = rowptr[v + 1] - s;
  }
  for (int64_t t = lane; t < n; t += kFillThreads) out[o + t] = col[s + t];
  if (add_self && total > n && lane == 0) out[o + n] = int32_t(v);
}

// block column (global id) -> position in the sorted frontier, plus the exact batch-local column degree
__global__ void block_remap_kernel(const int32_t* __restrict__ cols, int64_t nnz, const int32_t* __restrict__ uniq,
                                   int64_t n_unique, int32_t* __restrict__ local, int32_t* __restrict__ cdeg) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < nnz; e += stride) {
    const int32_t c = cols[e];
    int64_t lo = 0, hi = n_unique;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (uniq[mid] < c) lo = mid + 1;
      else hi = mid;
    }
    local[e] = int32_t(lo);
    atomicAdd(cdeg + lo, 1);
  }
}

// ---- dense per-row helpers of the affinity path (warp per row) ----
__global__ void row_inv_norm_kernel(const float* __restrict__ x, int64_t ldx, int64_t n_rows, int d,
                                    float* __restrict__ inv_norm, float* __restrict__ sumsq) {
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const float4* p = reinterpret_cast<const float4*>(x + r * ldx);
  float ss = 0.f;
  for (int c = lane; c < (d >> 2); c += 32) {
    const float4 v = __ldg(p + c);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if (lane == 0) {
    if (sumsq) sumsq[r] = ss;
    const float nrm = sqrtf(ss);
    inv_norm[r] = nrm > 0.f ? 1.f / nrm : 0.f;  // pow(norm,-1) with inf -> 0 (run.py:177-179)
  }
}

__global__ void normalize_backward_kernel(const float* __restrict__ e, int64_t lde, const float* __restrict__ inv_norm,
                                          float* __restrict__ g, int64_t ldg, int64_t n_rows, int d) {
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rows) return;
  const float inv = __ldg(inv_norm + r);
  const float4* pe = reinterpret_cast<const float4*>(e + r * lde);
  float4* pg = reinterpret_cast<float4*>(g + r * ldg);
  float dt = 0.f;
  for (int c = lane; c < (d >> 2); c += 32) {
    const float4 a = __ldg(pe + c), b = pg[c];
    dt += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) dt += __shfl_xor_sync(0xffffffffu, dt, off);
  const float s = dt * inv * inv;  // <e^, g> * inv  with e^ = e * inv
  for (int c = lane; c < (d >> 2); c += 32) {
    const float4 a = __ldg(pe + c);
    float4 b = pg[c];
    b.x = (b.x - a.x * s) * inv;
    b.y = (b.y - a.y * s) * inv;
    b.z = (b.z - a.z * s) * inv;
    b.w = (b.w - a.w * s) * inv;
    pg[c] = b;
  }
}

// ---------------------------------------------------------------------------
// R-MAT generator
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9e3779b97f4a7c15ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return z ^ (z >> 31);
}

struct RmatParams {
  int64_t n_edges, n_local;
  int32_t shard_bits, shard, local_bits;
  uint64_t seed;
  uint32_t ta, tab, tabc;  // 16-bit thresholds for (a), (a+b), (a+b+c)
  uint32_t t0, t1;         // P(col bit = 0 | row bit = 0), P(col bit = 0 | row bit = 1)
  int64_t filter_lo, filter_hi;
};

__global__ void rmat_kernel(RmatParams p, uint64_t* __restrict__ keys, unsigned long long* __restrict__ counter) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < p.n_edges; e += stride) {
    uint64_t st = p.seed * 0xd1342543de82ef95ull + uint64_t(p.shard) * 0x2545f4914f6cdd1dull + uint64_t(e);
    splitmix64(st);
    // source shard: row bits fixed by `shard`, column bits drawn from the conditional
    int64_t src_shard = 0;
    uint64_t bits = splitmix64(st);
    for (int l = p.shard_bits - 1; l >= 0; --l) {
      const uint32_t u = uint32_t(bits & 0xffffu);
      bits >>= 16;
      const int rbit = (p.shard >> l) & 1;
      const int cbit = u < (rbit ? p.t1 : p.t0) ? 0 : 1;
      src_shard = (src_shard << 1) | cbit;
    }
    int64_t dl = 0, sl = 0;
    for (int attempt = 0; attempt < 32; ++attempt) {
      dl = 0;
      sl = 0;
      int have = 0;
      for (int l = 0; l < p.local_bits; ++l) {
        if (have == 0) {
          bits = splitmix64(st);
          have = 4;
        }
        const uint32_t u = uint32_t(bits & 0xffffu);
        bits >>= 16;
        --have;
        const int q = u < p.ta ? 0 : (u < p.tab ? 1 : (u < p.tabc ? 2 : 3));  // 0:a 1:b 2:c 3:d
        dl = (dl << 1) | (q >> 1);
        sl = (sl << 1) | (q & 1);
      }
      if (dl < p.n_local && sl < p.n_local) break;
      if (attempt == 31) {
        dl %= p.n_local;
        sl %= p.n_local;
      }
    }
    const int64_t dst = int64_t(p.shard) * p.n_local + dl;
    const int64_t src = src_shard * p.n_local + sl;
    if (p.filter_lo < p.filter_hi) {
      if (src >= p.filter_lo && src < p.filter_hi) {
        const unsigned long long pos = atomicAdd(counter, 1ull);
        keys[pos] = (uint64_t(src) << 32) | uint64_t(dst);
      }
    } else {
      keys[e] = (uint64_t(dst) << 32) | uint64_t(src);
    }
  }
}

// ---------------------------------------------------------------------------
// host implementations
// ---------------------------------------------------------------------------
static inline unsigned blocks_for(int64_t n, int threads = 256) { return (unsigned)((n + threads - 1) / threads); }

static int bits_for(int64_t n) {
  int b = 0;
  while ((int64_t(1) << b) < n) ++b;
  return b < 1 ? 1 : b;
}

int coo_keys_to_csr_impl(uint64_t* keys, int64_t n, int64_t n_rows, int64_t* rowptr, int32_t* col, cudaStream_t st) {
  GGAD_REQUIRE(n >= 0 && n_rows >= 0 && rowptr && (n == 0 || (keys && col)), GGAD_ERR_INVALID, "coo_keys_to_csr: bad arguments");
  if (n > 0) {
    uint64_t* alt = nullptr;
    GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&alt), size_t(n) * 8, st));
    cub::DoubleBuffer<uint64_t> db(keys, alt);
    size_t tmp_bytes = 0;
    const int end_bit = 32 + bits_for(n_rows);
    GGAD_CUDA_OK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, db, n, 0, end_bit, st));
    void* tmp = nullptr;
    GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&tmp), tmp_bytes, st));
    GGAD_CUDA_OK(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, db, n, 0, end_bit, st));
    if (db.Current() != keys) GGAD_CUDA_OK(cudaMemcpyAsync(keys, db.Current(), size_t(n) * 8, cudaMemcpyDeviceToDevice, st));
    GGAD_CUDA_OK(cudaFreeAsync(tmp, st));
    GGAD_CUDA_OK(cudaFreeAsync(alt, st));
    low32_of_keys<<<blocks_for(n), 256, 0, st>>>(keys, n, col);
    GGAD_CUDA_OK(cudaGetLastError());
    count_launch(1);
  }
  rowptr_from_sorted_keys<<<blocks_for(n_rows + 1), 256, 0, st>>>(keys, n, n_rows, rowptr);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

int csr_transpose_impl(const int64_t* rowptr, const int32_t* col, const float* val, int64_t n_rows, int64_t n_cols,
                       int64_t nnz, int64_t* rowptrT, int32_t* colT, float* valT, int64_t* perm, cudaStream_t st) {
  GGAD_REQUIRE(rowptr && rowptrT && n_rows >= 0 && n_cols >= 0 && nnz >= 0, GGAD_ERR_INVALID, "csr_transpose: bad arguments");
  GGAD_REQUIRE(nnz == 0 || (col && colT), GGAD_ERR_INVALID, "csr_transpose: col/colT required");
  GGAD_REQUIRE(!val || valT, GGAD_ERR_INVALID, "csr_transpose: valT required when val is given");
  uint64_t *k0 = nullptr, *k1 = nullptr;
  int64_t *i0 = nullptr, *i1 = nullptr;
  const bool need_idx = (val != nullptr) || (perm != nullptr);
  if (nnz > 0) {
    GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&k0), size_t(nnz) * 8, st));
    GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&k1), size_t(nnz) * 8, st));
    if (need_idx) {
      GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&i0), size_t(nnz) * 8, st));
      GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&i1), size_t(nnz) * 8, st));
    }
    transpose_keys<<<blocks_for(nnz), 256, 0, st>>>(rowptr, col, n_rows, nnz, k0, i0);
    GGAD_CUDA_OK(cudaGetLastError());
    count_launch(1);
    cub::DoubleBuffer<uint64_t> dk(k0, k1);
    size_t tmp_bytes = 0;
    void* tmp = nullptr;
    // keys are (col, row) with rows already ascending inside each source row; a stable sort on the
    // col bits alone would do, but sorting the full key keeps the result independent of input order.
    const int end_bit = 32 + bits_for(n_cols);
    if (need_idx) {
      cub::DoubleBuffer<int64_t> di(i0, i1);
      GGAD_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, di, nnz, 0, end_bit, st));
      GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&tmp), tmp_bytes, st));
      GGAD_CUDA_OK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, di, nnz, 0, end_bit, st));
      if (val) {
        gather_vals<<<blocks_for(nnz), 256, 0, st>>>(val, di.Current(), nnz, valT);
        GGAD_CUDA_OK(cudaGetLastError());
        count_launch(1);
      }
      if (perm) GGAD_CUDA_OK(cudaMemcpyAsync(perm, di.Current(), size_t(nnz) * 8, cudaMemcpyDeviceToDevice, st));
    } else {
      GGAD_CUDA_OK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, dk, nnz, 0, end_bit, st));
      GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&tmp), tmp_bytes, st));
      GGAD_CUDA_OK(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, dk, nnz, 0, end_bit, st));
    }
    low32_of_keys<<<blocks_for(nnz), 256, 0, st>>>(dk.Current(), nnz, colT);
    GGAD_CUDA_OK(cudaGetLastError());
    rowptr_from_sorted_keys<<<blocks_for(n_cols + 1), 256, 0, st>>>(dk.Current(), nnz, n_cols, rowptrT);
    GGAD_CUDA_OK(cudaGetLastError());
    count_launch(2);
    GGAD_CUDA_OK(cudaFreeAsync(tmp, st));
    GGAD_CUDA_OK(cudaFreeAsync(k0, st));
    GGAD_CUDA_OK(cudaFreeAsync(k1, st));
    if (need_idx) {
      GGAD_CUDA_OK(cudaFreeAsync(i0, st));
      GGAD_CUDA_OK(cudaFreeAsync(i1, st));
    }
  } else {
    GGAD_CUDA_OK(cudaMemsetAsync(rowptrT, 0, size_t(n_cols + 1) * 8, st));
  }
  return GGAD_OK;
}

int csr_extract_rows_impl(const int64_t* rowptr, const int32_t* col, const float* val, const int32_t* rows, int64_t n_sel,
                          const int64_t* sub_rowptr, int32_t* sub_col, float* sub_val, cudaStream_t st) {
  GGAD_REQUIRE(rowptr && rows && sub_rowptr && n_sel >= 0, GGAD_ERR_INVALID, "csr_extract_rows: bad arguments");
  GGAD_REQUIRE(!val || sub_val, GGAD_ERR_INVALID, "csr_extract_rows: sub_val required when val is given");
  if (n_sel == 0) return GGAD_OK;
  extract_rows_kernel<<<blocks_for(n_sel * 32), 256, 0, st>>>(rowptr, col, val, rows, n_sel, sub_rowptr, sub_col, sub_val);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

int col_histogram_impl(const int32_t* col, int64_t nnz, int32_t* counts, int64_t n_cols, cudaStream_t st) {
  GGAD_REQUIRE(counts && n_cols >= 0 && nnz >= 0 && (nnz == 0 || col), GGAD_ERR_INVALID, "col_histogram: bad arguments");
  GGAD_CUDA_OK(cudaMemsetAsync(counts, 0, size_t(n_cols) * 4, st));
  if (nnz == 0) return GGAD_OK;
  unsigned blocks = blocks_for(nnz);
  if (blocks > 148u * 32u) blocks = 148u * 32u;
  col_hist_kernel<<<blocks, 256, 0, st>>>(col, nnz, counts);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}


// val[e] = 1 / sqrt(counts[col[e]]): the batch-local column normalisation of a hop block as a per-edge value
__global__ void edge_inv_sqrt_count_kernel(const int32_t* __restrict__ col, int64_t nnz, const int32_t* __restrict__ counts,
                                           float* __restrict__ val) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < nnz; e += stride)
    val[e] = __fdiv_rn(1.f, __fsqrt_rn(float(__ldg(counts + col[e]))));
}

int block_col_weights_impl(const int32_t* col, int64_t nnz, int32_t* counts, int64_t n_nodes, float* val, cudaStream_t st) {
  GGAD_REQUIRE(counts && val && n_nodes >= 0 && nnz >= 0 && (nnz == 0 || col), GGAD_ERR_INVALID, "block_col_weights: bad arguments");
  int rc = col_histogram_impl(col, nnz, counts, n_nodes, st);
  if (rc != GGAD_OK || nnz == 0) return rc;
  unsigned blocks = blocks_for(nnz);
  if (blocks > 148u * 32u) blocks = 148u * 32u;
  edge_inv_sqrt_count_kernel<<<blocks, 256, 0, st>>>(col, nnz, counts, val);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

int block_rowptr_impl(const int64_t* rowptr, const int32_t* col, int64_t n_nodes, const int32_t* nodes, int64_t n_batch,
                      int add_self, int64_t* block_rowptr, int64_t* nnz_host, cudaStream_t st) {
  GGAD_REQUIRE(rowptr && (nodes || n_batch == 0) && block_rowptr && nnz_host && n_batch >= 0 && n_nodes >= 0,
               GGAD_ERR_INVALID, "block_rowptr: bad arguments");
  int64_t* deg = nullptr;
  GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&deg), size_t(n_batch + 1) * 8, st));
  block_degree_kernel<<<blocks_for(n_batch + 1), 256, 0, st>>>(rowptr, col, n_nodes, nodes, n_batch, add_self, deg);
  GGAD_CUDA_OK(cudaGetLastError());
  size_t tmp_bytes = 0;
  GGAD_CUDA_OK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, deg, block_rowptr, n_batch + 1, st));
  void* tmp = nullptr;
  GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&tmp), tmp_bytes ? tmp_bytes : 8, st));
  GGAD_CUDA_OK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, deg, block_rowptr, n_batch + 1, st));
  count_launch(2);
  GGAD_CUDA_OK(cudaMemcpyAsync(nnz_host, block_rowptr + n_batch, 8, cudaMemcpyDeviceToHost, st));
  GGAD_CUDA_OK(cudaFreeAsync(tmp, st));
  GGAD_CUDA_OK(cudaFreeAsync(deg, st));
  GGAD_CUDA_OK(cudaStreamSynchronize(st));
  return GGAD_OK;
}

int block_fill_impl(const int64_t* rowptr, const int32_t* col, int64_t n_nodes, const int32_t* nodes, int64_t n_batch,
                    int add_self, const int64_t* block_rowptr, int32_t* block_col, cudaStream_t st) {
  GGAD_REQUIRE(rowptr && (nodes || n_batch == 0) && block_rowptr && n_batch >= 0, GGAD_ERR_INVALID, "block_fill: bad arguments");
  if (n_batch == 0) return GGAD_OK;
  GGAD_REQUIRE(n_batch < (int64_t(1) << 31), GGAD_ERR_UNSUPPORTED, "block_fill: too many block rows");
  block_fill_kernel<<<(unsigned)n_batch, kFillThreads, 0, st>>>(rowptr, col, n_nodes, nodes, n_batch, add_self, block_rowptr,
                                                                block_col);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

int unique_sorted_impl(const int32_t* keys, int64_t n, int64_t key_bound, int32_t* uniq, int64_t* n_unique_host,
                       cudaStream_t st) {
  GGAD_REQUIRE(n >= 0 && n_unique_host && (n == 0 || (keys && uniq)), GGAD_ERR_INVALID, "unique_sorted: bad arguments");
  if (n == 0) {
    *n_unique_host = 0;
    return GGAD_OK;
  }
  uint32_t *k0 = nullptr, *k1 = nullptr;
  int64_t* d_count = nullptr;
  GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&k0), size_t(n) * 4, st));
  GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&k1), size_t(n) * 4, st));
  GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&d_count), 8, st));
  GGAD_CUDA_OK(cudaMemcpyAsync(k0, keys, size_t(n) * 4, cudaMemcpyDeviceToDevice, st));
  cub::DoubleBuffer<uint32_t> db(k0, k1);
  size_t sort_bytes = 0, sel_bytes = 0;
  const int end_bit = bits_for(key_bound > 1 ? key_bound : 2);
  GGAD_CUDA_OK(cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, db, n, 0, end_bit, st));
  GGAD_CUDA_OK(cub::DeviceSelect::Unique(nullptr, sel_bytes, k0, reinterpret_cast<uint32_t*>(uniq), d_count, n, st));
  void* tmp = nullptr;
  const size_t tmp_bytes = sort_bytes > sel_bytes ? sort_bytes : sel_bytes;
  GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&tmp), tmp_bytes ? tmp_bytes : 8, st));
  size_t b = sort_bytes;
  GGAD_CUDA_OK(cub::DeviceRadixSort::SortKeys(tmp, b, db, n, 0, end_bit, st));
  b = sel_bytes;
  GGAD_CUDA_OK(cub::DeviceSelect::Unique(tmp, b, db.Current(), reinterpret_cast<uint32_t*>(uniq), d_count, n, st));
  GGAD_CUDA_OK(cudaMemcpyAsync(n_unique_host, d_count, 8, cudaMemcpyDeviceToHost, st));
  GGAD_CUDA_OK(cudaFreeAsync(tmp, st));
  GGAD_CUDA_OK(cudaFreeAsync(k0, st));
  GGAD_CUDA_OK(cudaFreeAsync(k1, st));
  GGAD_CUDA_OK(cudaFreeAsync(d_count, st));
  GGAD_CUDA_OK(cudaStreamSynchronize(st));
  return GGAD_OK;
}

int block_remap_impl(const int32_t* cols, int64_t nnz, const int32_t* uniq, int64_t n_unique, int32_t* local,
                     int32_t* cdeg, cudaStream_t st) {
  GGAD_REQUIRE(nnz >= 0 && n_unique >= 0 && cdeg && (nnz == 0 || (cols && uniq && local)), GGAD_ERR_INVALID,
               "block_remap: bad arguments");
  GGAD_CUDA_OK(cudaMemsetAsync(cdeg, 0, size_t(n_unique) * 4, st));
  if (nnz == 0) return GGAD_OK;
  unsigned blocks = blocks_for(nnz);
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  block_remap_kernel<<<blocks, 256, 0, st>>>(cols, nnz, uniq, n_unique, local, cdeg);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

int row_inv_norm_impl(const float* x, int64_t ldx, int64_t n_rows, int32_t d, float* inv_norm, float* sumsq, cudaStream_t st) {
  GGAD_REQUIRE(x && inv_norm && n_rows >= 0 && d > 0 && d % 4 == 0, GGAD_ERR_INVALID, "row_inv_norm: bad arguments");
  GGAD_REQUIRE(ldx % 4 == 0 && ldx >= d && aligned16(x), GGAD_ERR_ALIGN, "row_inv_norm: x must be 16-byte aligned, ldx multiple of 4");
  if (n_rows == 0) return GGAD_OK;
  row_inv_norm_kernel<<<blocks_for(n_rows * 32), 256, 0, st>>>(x, ldx, n_rows, d, inv_norm, sumsq);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

int normalize_backward_impl(const float* e, int64_t lde, const float* inv_norm, float* g, int64_t ldg, int64_t n_rows,
                            int32_t d, cudaStream_t st) {
  GGAD_REQUIRE(e && inv_norm && g && n_rows >= 0 && d > 0 && d % 4 == 0, GGAD_ERR_INVALID, "normalize_backward: bad arguments");
  GGAD_REQUIRE(lde % 4 == 0 && ldg % 4 == 0 && lde >= d && ldg >= d && aligned16(e) && aligned16(g), GGAD_ERR_ALIGN,
               "normalize_backward: e/g must be 16-byte aligned with leading dimensions multiple of 4");
  if (n_rows == 0) return GGAD_OK;
  normalize_backward_kernel<<<blocks_for(n_rows * 32), 256, 0, st>>>(e, lde, inv_norm, g, ldg, n_rows, d);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

int rmat_keys_impl(uint64_t* keys, int64_t n_edges, int64_t n_local, int32_t n_shards, int32_t shard, uint64_t seed,
                   float a, float b, float c, int64_t filter_lo, int64_t filter_hi, int64_t* n_out_host, cudaStream_t st) {
  GGAD_REQUIRE(keys && n_edges >= 0 && n_local > 0, GGAD_ERR_INVALID, "rmat_keys: bad arguments");
  GGAD_REQUIRE(n_shards >= 1 && (n_shards & (n_shards - 1)) == 0 && shard >= 0 && shard < n_shards, GGAD_ERR_INVALID,
               "rmat_keys: n_shards must be a power of two and 0 <= shard < n_shards");
  GGAD_REQUIRE(int64_t(n_shards) * n_local < (int64_t(1) << 31), GGAD_ERR_UNSUPPORTED, "rmat_keys: node ids must fit int32");
  const float dd = 1.f - a - b - c;
  GGAD_REQUIRE(a > 0 && b >= 0 && c >= 0 && dd >= 0, GGAD_ERR_INVALID, "rmat_keys: bad (a,b,c)");
  RmatParams p;
  p.n_edges = n_edges; p.n_local = n_local; p.shard = shard; p.seed = seed;
  p.shard_bits = 0;
  while ((1 << p.shard_bits) < n_shards) ++p.shard_bits;
  p.local_bits = bits_for(n_local);
  p.ta = uint32_t(a * 65536.f); p.tab = uint32_t((a + b) * 65536.f); p.tabc = uint32_t((a + b + c) * 65536.f);
  p.t0 = uint32_t(a / (a + b) * 65536.f);
  p.t1 = (c + dd) > 0 ? uint32_t(c / (c + dd) * 65536.f) : 65536u;
  p.filter_lo = filter_lo; p.filter_hi = filter_hi;
  const bool filtered = filter_lo < filter_hi;
  unsigned long long* counter = nullptr;
  if (filtered) {
    GGAD_REQUIRE(n_out_host, GGAD_ERR_INVALID, "rmat_keys: n_out_host required with a filter");
    GGAD_CUDA_OK(temp_alloc(reinterpret_cast<void**>(&counter), 8, st));
    GGAD_CUDA_OK(cudaMemsetAsync(counter, 0, 8, st));
  }
  if (n_edges > 0) {
    unsigned blocks = blocks_for(n_edges);
    if (blocks > 148u * 64u) blocks = 148u * 64u;
    rmat_kernel<<<blocks, 256, 0, st>>>(p, keys, counter);
    GGAD_CUDA_OK(cudaGetLastError());
    count_launch(1);
  }
  if (filtered) {
    unsigned long long h = 0;
    GGAD_CUDA_OK(cudaMemcpyAsync(&h, counter, 8, cudaMemcpyDeviceToHost, st));
    GGAD_CUDA_OK(cudaStreamSynchronize(st));
    GGAD_CUDA_OK(cudaFreeAsync(counter, st));
    *n_out_host = int64_t(h);
  } else if (n_out_host) {
    *n_out_host = n_edges;
  }
  return GGAD_OK;
}

}  // namespace ggad
