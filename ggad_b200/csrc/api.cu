// extern "C" surface of libggad_b200.so (see include/ggad_b200.h).  Error strings are
// thread-local; nothing here throws across the ABI.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace ggad {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static cudaMemPool_t g_pools[64] = {};
static std::mutex g_pool_mu;

cudaError_t temp_alloc(void** p, size_t bytes, cudaStream_t st) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev >= 64) return cudaMallocAsync(p, bytes, st);
  cudaMemPool_t pool;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pools[dev]) {
      cudaMemPoolProps props = {};
      props.allocType = cudaMemAllocationTypePinned;
      props.location.type = cudaMemLocationTypeDevice;
      props.location.id = dev;
      e = cudaMemPoolCreate(&g_pools[dev], &props);
      if (e != cudaSuccess) {
        g_pools[dev] = nullptr;
        return e;
      }
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(g_pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool = g_pools[dev];
  }
  return cudaMallocFromPoolAsync(p, bytes ? bytes : 8, pool, st);
}

int sm_count_cached() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  if (dev < 64 && cached[dev] > 0) return cached[dev];
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  if (dev < 64) cached[dev] = sms;
  return sms;
}

// implementations (gather_reduce.cu / graph_prep.cu)
int gather_reduce_impl(const ggad_gather_desc_t* d, cudaStream_t st);
int plan_build_impl(const int64_t*, int64_t, int64_t, int32_t*, int64_t*, cudaStream_t, int64_t n_tiles_padded = 0);
int coo_keys_to_csr_impl(uint64_t*, int64_t, int64_t, int64_t*, int32_t*, cudaStream_t);
int csr_transpose_impl(const int64_t*, const int32_t*, const float*, int64_t, int64_t, int64_t, int64_t*, int32_t*, float*,
                       int64_t*, cudaStream_t);
int csr_extract_rows_impl(const int64_t*, const int32_t*, const float*, const int32_t*, int64_t, const int64_t*, int32_t*,
                          float*, cudaStream_t);
int col_histogram_impl(const int32_t*, int64_t, int32_t*, int64_t, cudaStream_t);
int row_inv_norm_impl(const float*, int64_t, int64_t, int32_t, float*, float*, cudaStream_t);
int normalize_backward_impl(const float*, int64_t, const float*, float*, int64_t, int64_t, int32_t, cudaStream_t);
int rmat_keys_impl(uint64_t*, int64_t, int64_t, int32_t, int32_t, uint64_t, float, float, float, int64_t, int64_t, int64_t*,
                   cudaStream_t);

int halo_push_impl(const float*, int64_t, int64_t, int32_t, const uint32_t*, float* const*, int32_t, cudaStream_t);
int halo_chase_impl(const ggad_chase_desc_t*, cudaStream_t);
int minibatch_tail_fwd_impl(const ggad_tail_desc_t*, cudaStream_t);
int minibatch_tail_bwd_impl(const ggad_tail_desc_t*, cudaStream_t);
int block_col_weights_impl(const int32_t*, int64_t, int32_t*, int64_t, float*, cudaStream_t);
int csr_row_sum_f64_impl(const int64_t*, const float*, int64_t, double*, cudaStream_t);
int csr_add_identity_rowptr_impl(const int64_t*, const int32_t*, int64_t, int64_t*, int64_t*, cudaStream_t);
int csr_scale_add_identity_impl(const int64_t*, const int32_t*, const float*, const double*, int64_t, const int64_t*, int32_t*,
                                float*, cudaStream_t);
int dense_matmul_impl(int, int, int64_t, int64_t, int64_t, const float*, int64_t, const float*, int64_t, float*, int64_t, float,
                      float, int, int, cudaStream_t);

int block_rowptr_impl(const int64_t*, const int32_t*, int64_t, const int32_t*, int64_t, int, int64_t*, int64_t*, cudaStream_t);
int block_fill_impl(const int64_t*, const int32_t*, int64_t, const int32_t*, int64_t, int, const int64_t*, int32_t*,
                    cudaStream_t);
int unique_sorted_impl(const int32_t*, int64_t, int64_t, int32_t*, int64_t*, cudaStream_t);
int block_remap_impl(const int32_t*, int64_t, const int32_t*, int64_t, int32_t*, int32_t*, cudaStream_t);

// deterministic single-block sum of n floats into a double
__global__ void sum_to_double_kernel(const float* __restrict__ v, int64_t n, double scale, double* __restrict__ out) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += double(v[i]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = blockDim.x / 2; off > 0; off >>= 1) {
    if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0] * scale;
}

static void fill_desc(ggad_gather_desc_t& g, const ggad_resident_csr_t* m, const float* x, float* y, int32_t d, float* ws) {
  memset(&g, 0, sizeof(g));
  g.rowptr = m->rowptr; g.col = m->col; g.val = m->val; g.n_rows = m->n_rows; g.nnz = m->nnz;
  g.x = x; g.ldx = d; g.row_scale = m->row_scale; g.col_scale = m->col_scale; g.d = d;
  g.y = y; g.ldy = d;
  g.tile_row = m->tile_row; g.tile_edge = m->tile_edge; g.n_tiles = m->n_tiles;
  g.ws = m->tile_row ? ws : nullptr;
}

int spmm_fwd_bwd_host_impl(const ggad_resident_csr_t* a, const ggad_resident_csr_t* at, const float* x_host, float* y_host,
                           float* dx_host, double* loss_host, int32_t d, float* dev_x, float* dev_y, float* dev_dx,
                           float* dev_ws, cudaStream_t st, bool sync) {
  GGAD_REQUIRE(a && at && x_host && dx_host && dev_x && dev_y && dev_dx && dev_ws, GGAD_ERR_INVALID, "spmm_fwd_bwd_host: null pointer");
  GGAD_REQUIRE(a->n_rows == at->n_cols && a->n_cols == at->n_rows && a->nnz == at->nnz, GGAD_ERR_INVALID,
               "spmm_fwd_bwd_host: at is not the transpose shape of a");
  const size_t xin = size_t(a->n_cols) * d * 4, yout = size_t(a->n_rows) * d * 4;
  GGAD_CUDA_OK(cudaMemcpyAsync(dev_x, x_host, xin, cudaMemcpyHostToDevice, st));
  ggad_gather_desc_t g;
  fill_desc(g, a, dev_x, dev_y, d, dev_ws);
  // |y_r|^2 per row lands in the first n_rows floats after the tile workspace of `a`
  float* sumsq = dev_ws + 2 * (a->n_tiles > at->n_tiles ? a->n_tiles : at->n_tiles) * int64_t(d);
  g.sumsq = sumsq;
  int rc = gather_reduce_impl(&g, st);
  if (rc != GGAD_OK) return rc;
  double* dev_loss = reinterpret_cast<double*>(sumsq + ((a->n_rows + 3) & ~int64_t(3)));
  sum_to_double_kernel<<<1, 1024, 0, st>>>(sumsq, a->n_rows, 0.5, dev_loss);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  if (y_host) GGAD_CUDA_OK(cudaMemcpyAsync(y_host, dev_y, yout, cudaMemcpyDeviceToHost, st));
  fill_desc(g, at, dev_y, dev_dx, d, dev_ws);
  rc = gather_reduce_impl(&g, st);
  if (rc != GGAD_OK) return rc;
  GGAD_CUDA_OK(cudaMemcpyAsync(dx_host, dev_dx, xin, cudaMemcpyDeviceToHost, st));
  if (!sync) {  // enqueue only: loss_host must be pinned, the caller synchronises the stream
    if (loss_host) GGAD_CUDA_OK(cudaMemcpyAsync(loss_host, dev_loss, 8, cudaMemcpyDeviceToHost, st));
    return GGAD_OK;
  }
  double h = 0.0;
  GGAD_CUDA_OK(cudaMemcpyAsync(&h, dev_loss, 8, cudaMemcpyDeviceToHost, st));
  GGAD_CUDA_OK(cudaStreamSynchronize(st));
  if (loss_host) *loss_host = h;
  return GGAD_OK;
}

}  // namespace ggad

using namespace ggad;

extern "C" {

GGAD_API int ggad_version(void) { return 100; }
GGAD_API const char* ggad_last_error(void) { return g_err; }
GGAD_API int64_t ggad_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

GGAD_API int ggad_device_info(int* sm_count, int64_t* l2_bytes, int* cc_major, int* cc_minor, int64_t* hbm_bytes) {
  int dev = 0;
  GGAD_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  GGAD_CUDA_OK(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (l2_bytes) *l2_bytes = p.l2CacheSize;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (hbm_bytes) *hbm_bytes = (int64_t)p.totalGlobalMem;
  return GGAD_OK;
}

GGAD_API int64_t ggad_plan_num_tiles(int64_t n_rows, int64_t nnz) {
  if (n_rows < 0 || nnz < 0) return 0;
  return (n_rows + nnz + GGAD_TILE_ITEMS - 1) / GGAD_TILE_ITEMS;
}

GGAD_API int ggad_plan_build(const int64_t* rowptr, int64_t n_rows, int64_t nnz, int32_t* tile_row, int64_t* tile_edge,
                    ggad_stream_t stream) {
  return plan_build_impl(rowptr, n_rows, nnz, tile_row, tile_edge, (cudaStream_t)stream);
}

GGAD_API int ggad_plan_build_padded(const int64_t* rowptr, int64_t n_rows, int64_t nnz, int64_t n_tiles, int32_t* tile_row,
                                    int64_t* tile_edge, ggad_stream_t stream) {
  return plan_build_impl(rowptr, n_rows, nnz, tile_row, tile_edge, (cudaStream_t)stream, n_tiles);
}

GGAD_API int ggad_gather_reduce(const ggad_gather_desc_t* desc, ggad_stream_t stream) {
  return gather_reduce_impl(desc, (cudaStream_t)stream);
}

GGAD_API int ggad_trim_workspace(void) {
  int dev = 0;
  GGAD_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_pool_mu);
  if (dev < 64 && g_pools[dev]) {
    GGAD_CUDA_OK(cudaDeviceSynchronize());
    GGAD_CUDA_OK(cudaMemPoolTrimTo(g_pools[dev], 0));
  }
  return GGAD_OK;
}

GGAD_API int ggad_reserve_workspace(int64_t bytes, ggad_stream_t stream) {
  if (bytes <= 0) return GGAD_OK;
  void* p = nullptr;
  GGAD_CUDA_OK(temp_alloc(&p, size_t(bytes), (cudaStream_t)stream));
  GGAD_CUDA_OK(cudaFreeAsync(p, (cudaStream_t)stream));
  return GGAD_OK;
}

GGAD_API int ggad_halo_push(const float* y, int64_t ldy, int64_t n_rows, int32_t d, const uint32_t* peer_need,
                            float* const* y_peer_host, int32_t n_peer, ggad_stream_t stream) {
  return halo_push_impl(y, ldy, n_rows, d, peer_need, y_peer_host, n_peer, (cudaStream_t)stream);
}

GGAD_API int ggad_halo_chase(const ggad_chase_desc_t* desc, ggad_stream_t stream) {
  return halo_chase_impl(desc, (cudaStream_t)stream);
}

GGAD_API int ggad_dense_matmul(int32_t trans_a, int32_t trans_b, int64_t m, int64_t n, int64_t k, const float* a, int64_t lda,
                               const float* b, int64_t ldb, float* c, int64_t ldc, float alpha, float beta, int32_t relu,
                               int32_t path, ggad_stream_t stream) {
  return dense_matmul_impl(trans_a, trans_b, m, n, k, a, lda, b, ldb, c, ldc, alpha, beta, relu, path, (cudaStream_t)stream);
}

GGAD_API int ggad_csr_row_sum_f64(const int64_t* rowptr, const float* val, int64_t n_rows, double* deg, ggad_stream_t stream) {
  return csr_row_sum_f64_impl(rowptr, val, n_rows, deg, (cudaStream_t)stream);
}

GGAD_API int ggad_csr_add_identity_rowptr(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t* out_rowptr,
                                          int64_t* nnz_host, ggad_stream_t stream) {
  return csr_add_identity_rowptr_impl(rowptr, col, n, out_rowptr, nnz_host, (cudaStream_t)stream);
}

GGAD_API int ggad_csr_scale_add_identity(const int64_t* rowptr, const int32_t* col, const float* val, const double* scale,
                                         int64_t n, const int64_t* out_rowptr, int32_t* out_col, float* out_val,
                                         ggad_stream_t stream) {
  return csr_scale_add_identity_impl(rowptr, col, val, scale, n, out_rowptr, out_col, out_val, (cudaStream_t)stream);
}

GGAD_API int ggad_normalize_backward(const float* e, int64_t lde, const float* inv_norm, float* g, int64_t ldg, int64_t n_rows,
                            int32_t d, ggad_stream_t stream) {
  return normalize_backward_impl(e, lde, inv_norm, g, ldg, n_rows, d, (cudaStream_t)stream);
}

GGAD_API int ggad_row_inv_norm(const float* x, int64_t ldx, int64_t n_rows, int32_t d, float* inv_norm, float* sumsq,
                      ggad_stream_t stream) {
  return row_inv_norm_impl(x, ldx, n_rows, d, inv_norm, sumsq, (cudaStream_t)stream);
}

GGAD_API int ggad_coo_keys_to_csr(uint64_t* keys, int64_t n, int64_t n_rows, int64_t* rowptr, int32_t* col, ggad_stream_t stream) {
  return coo_keys_to_csr_impl(keys, n, n_rows, rowptr, col, (cudaStream_t)stream);
}

GGAD_API int ggad_csr_transpose(const int64_t* rowptr, const int32_t* col, const float* val, int64_t n_rows, int64_t n_cols,
                       int64_t nnz, int64_t* rowptrT, int32_t* colT, float* valT, int64_t* perm, ggad_stream_t stream) {
  return csr_transpose_impl(rowptr, col, val, n_rows, n_cols, nnz, rowptrT, colT, valT, perm, (cudaStream_t)stream);
}

GGAD_API int ggad_csr_extract_rows(const int64_t* rowptr, const int32_t* col, const float* val, const int32_t* rows, int64_t n_sel,
                          const int64_t* sub_rowptr, int32_t* sub_col, float* sub_val, ggad_stream_t stream) {
  return csr_extract_rows_impl(rowptr, col, val, rows, n_sel, sub_rowptr, sub_col, sub_val, (cudaStream_t)stream);
}

GGAD_API int ggad_col_histogram(const int32_t* col, int64_t nnz, int32_t* counts, int64_t n_cols, ggad_stream_t stream) {
  return col_histogram_impl(col, nnz, counts, n_cols, (cudaStream_t)stream);
}

GGAD_API int ggad_block_rowptr(const int64_t* adj_rowptr, const int32_t* adj_col, int64_t n_nodes, const int32_t* nodes,
                               int64_t n_batch, int32_t add_self, int64_t* block_rowptr, int64_t* nnz_host,
                               ggad_stream_t stream) {
  return block_rowptr_impl(adj_rowptr, adj_col, n_nodes, nodes, n_batch, add_self, block_rowptr, nnz_host,
                           (cudaStream_t)stream);
}

GGAD_API int ggad_block_fill(const int64_t* adj_rowptr, const int32_t* adj_col, int64_t n_nodes, const int32_t* nodes,
                             int64_t n_batch, int32_t add_self, const int64_t* block_rowptr, int32_t* block_col,
                             ggad_stream_t stream) {
  return block_fill_impl(adj_rowptr, adj_col, n_nodes, nodes, n_batch, add_self, block_rowptr, block_col,
                         (cudaStream_t)stream);
}

GGAD_API int ggad_unique_sorted(const int32_t* keys, int64_t n, int64_t key_bound, int32_t* uniq, int64_t* n_unique_host,
                                ggad_stream_t stream) {
  return unique_sorted_impl(keys, n, key_bound, uniq, n_unique_host, (cudaStream_t)stream);
}

GGAD_API int ggad_block_remap(const int32_t* cols, int64_t nnz, const int32_t* uniq, int64_t n_unique, int32_t* local,
                              int32_t* cdeg, ggad_stream_t stream) {
  return block_remap_impl(cols, nnz, uniq, n_unique, local, cdeg, (cudaStream_t)stream);
}

GGAD_API int ggad_block_col_weights(const int32_t* block_col, int64_t nnz, int32_t* counts, int64_t n_nodes, float* val,
                                    ggad_stream_t stream) {
  return block_col_weights_impl(block_col, nnz, counts, n_nodes, val, (cudaStream_t)stream);
}

GGAD_API int ggad_minibatch_tail_fwd(const ggad_tail_desc_t* desc, ggad_stream_t stream) {
  return minibatch_tail_fwd_impl(desc, (cudaStream_t)stream);
}
GGAD_API int ggad_minibatch_tail_bwd(const ggad_tail_desc_t* desc, ggad_stream_t stream) {
  return minibatch_tail_bwd_impl(desc, (cudaStream_t)stream);
}

GGAD_API int ggad_rmat_keys(uint64_t* keys, int64_t n_edges, int64_t n_local, int32_t n_shards, int32_t shard, uint64_t seed, float a,
                   float b, float c, int64_t filter_lo, int64_t filter_hi, int64_t* n_out_host, ggad_stream_t stream) {
  return rmat_keys_impl(keys, n_edges, n_local, n_shards, shard, seed, a, b, c, filter_lo, filter_hi, n_out_host,
                        (cudaStream_t)stream);
}

GGAD_API int ggad_spmm_fwd_bwd_host(const ggad_resident_csr_t* a, const ggad_resident_csr_t* at, const float* x_host, float* y_host,
                           float* dx_host, double* loss_host, int32_t d, float* dev_x, float* dev_y, float* dev_dx,
                           float* dev_ws, ggad_stream_t stream) {
  return spmm_fwd_bwd_host_impl(a, at, x_host, y_host, dx_host, loss_host, d, dev_x, dev_y, dev_dx, dev_ws,
                                (cudaStream_t)stream, true);
}

GGAD_API int ggad_spmm_fwd_bwd_host_enqueue(const ggad_resident_csr_t* a, const ggad_resident_csr_t* at, const float* x_host,
                                            float* y_host, float* dx_host, double* loss_host, int32_t d, float* dev_x,
                                            float* dev_y, float* dev_dx, float* dev_ws, ggad_stream_t stream) {
  return spmm_fwd_bwd_host_impl(a, at, x_host, y_host, dx_host, loss_host, d, dev_x, dev_y, dev_dx, dev_ws,
                                (cudaStream_t)stream, false);
}

}  // extern "C"
