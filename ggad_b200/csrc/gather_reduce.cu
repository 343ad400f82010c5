// Host side of ggad_gather_reduce / ggad_plan_build: validation and dispatch on the row width.
// The kernels live in gather_kernels.cuh and are instantiated per width class in gather_inst_*.cu.
#include "gather_kernels.cuh"

namespace ggad {

int launch_g4c1(const GatherArgs& a, cudaStream_t st, int sm_count);
int launch_g8c1(const GatherArgs& a, cudaStream_t st, int sm_count);
int launch_g16c1(const GatherArgs& a, cudaStream_t st, int sm_count);
int launch_g32c1(const GatherArgs& a, cudaStream_t st, int sm_count);
int launch_g32c2(const GatherArgs& a, cudaStream_t st, int sm_count);
int launch_g32c3(const GatherArgs& a, cudaStream_t st, int sm_count);
int launch_g32c4(const GatherArgs& a, cudaStream_t st, int sm_count);
int launch_g32c6(const GatherArgs& a, cudaStream_t st, int sm_count);

static int dispatch_width(const GatherArgs& a, cudaStream_t st, int sm_count) {
  const int V = a.d >> 2;
  if (V <= 4) return launch_g4c1(a, st, sm_count);
  if (V <= 8) return launch_g8c1(a, st, sm_count);
  if (V <= 16) return launch_g16c1(a, st, sm_count);
  if (V <= 32) return launch_g32c1(a, st, sm_count);
  if (V <= 64) return launch_g32c2(a, st, sm_count);
  if (V <= 96) return launch_g32c3(a, st, sm_count);
  if (V <= 128) return launch_g32c4(a, st, sm_count);
  return launch_g32c6(a, st, sm_count);
}

// ---------------------------------------------------------------------------
// merge-path plan
// ---------------------------------------------------------------------------
__global__ void plan_kernel(const int64_t* __restrict__ rowptr, int64_t n_rows, int64_t nnz, int64_t n_tiles,
                            int32_t* __restrict__ tile_row, int64_t* __restrict__ tile_edge) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t > n_tiles) return;
  int64_t diag = t * kTile;
  if (diag > n_rows + nnz) diag = n_rows + nnz;
  int64_t lo = diag > nnz ? diag - nnz : 0;
  int64_t hi = diag < n_rows ? diag : n_rows;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(rowptr + mid + 1) <= diag - mid - 1) lo = mid + 1;
    else hi = mid;
  }
  tile_row[t] = (int32_t)lo;
  tile_edge[t] = diag - lo;
}

int sm_count_cached();  // api.cu

int gather_reduce_impl(const ggad_gather_desc_t* d, cudaStream_t st) {
  GGAD_REQUIRE(d != nullptr, GGAD_ERR_INVALID, "gather_reduce: null descriptor");
  GGAD_REQUIRE(d->n_rows >= 0 && d->nnz >= 0, GGAD_ERR_INVALID, "gather_reduce: negative size");
  GGAD_REQUIRE(d->n_rows < (int64_t(1) << 31), GGAD_ERR_UNSUPPORTED, "gather_reduce: n_rows must fit int32");
  if (d->n_rows == 0) return GGAD_OK;
  GGAD_REQUIRE(d->rowptr && (d->col || d->nnz == 0) && d->x, GGAD_ERR_INVALID, "gather_reduce: rowptr/col/x required");
  GGAD_REQUIRE(d->d > 0 && d->d % 4 == 0, GGAD_ERR_INVALID, "gather_reduce: width d=%d must be a positive multiple of 4", d->d);
  GGAD_REQUIRE(d->d <= GGAD_MAX_WIDTH, GGAD_ERR_UNSUPPORTED, "gather_reduce: width d=%d > %d", d->d, GGAD_MAX_WIDTH);
  GGAD_REQUIRE(d->ldx % 4 == 0 && d->ldx >= d->d, GGAD_ERR_ALIGN, "gather_reduce: ldx=%lld must be >= d and a multiple of 4", (long long)d->ldx);
  GGAD_REQUIRE(aligned16(d->x), GGAD_ERR_ALIGN, "gather_reduce: x not 16-byte aligned");
  GGAD_REQUIRE(d->y || d->z || d->dot_out || d->sumsq || d->y_multicast, GGAD_ERR_INVALID, "gather_reduce: no output requested");
  if (d->y || d->z || d->y_multicast) {
    GGAD_REQUIRE(d->ldy % 4 == 0 && d->ldy >= d->d, GGAD_ERR_ALIGN, "gather_reduce: ldy=%lld must be >= d and a multiple of 4", (long long)d->ldy);
    GGAD_REQUIRE(aligned16(d->y) && aligned16(d->z), GGAD_ERR_ALIGN, "gather_reduce: y/z not 16-byte aligned");
  }
  GGAD_REQUIRE(!d->bias || aligned16(d->bias), GGAD_ERR_ALIGN, "gather_reduce: bias not 16-byte aligned");
  if (d->dot_out) {
    GGAD_REQUIRE(d->dot_mat && d->lddot % 4 == 0 && d->lddot >= d->d && aligned16(d->dot_mat), GGAD_ERR_ALIGN,
                 "gather_reduce: dot_mat must be 16-byte aligned with lddot >= d, multiple of 4");
  }
  const bool any_plan = d->tile_row || d->tile_edge || d->ws;
  if (any_plan) {
    GGAD_REQUIRE(d->tile_row && d->tile_edge && d->ws, GGAD_ERR_INVALID, "gather_reduce: plan needs tile_row, tile_edge and ws");
    GGAD_REQUIRE(d->n_tiles == ggad_plan_num_tiles(d->n_rows, d->nnz), GGAD_ERR_INVALID, "gather_reduce: n_tiles does not match the CSR");
    GGAD_REQUIRE(aligned16(d->ws) && aligned16(d->col) && (!d->val || aligned16(d->val)), GGAD_ERR_ALIGN,
                 "gather_reduce: ws/col/val must be 16-byte aligned for the tiled kernel");
  }
  GatherArgs a;
  a.rowptr = d->rowptr; a.col = d->col; a.val = d->val; a.n_rows = d->n_rows; a.nnz = d->nnz;
  a.x = d->x; a.ldx = d->ldx; a.xmap = d->xmap; a.col_scale = d->col_scale; a.row_scale = d->row_scale;
  a.d = d->d; a.relu = d->relu; a.bias = d->bias; a.prelu_slope = d->prelu_slope;
  a.y = d->y; a.z = d->z; a.ldy = d->ldy; a.sumsq = d->sumsq;
  a.dot_mat = d->dot_mat; a.lddot = d->lddot; a.dot_rows = d->dot_rows; a.dot_scale = d->dot_scale; a.dot_out = d->dot_out;
  a.tile_row = d->tile_row; a.tile_edge = d->tile_edge; a.n_tiles = d->n_tiles; a.ws = d->ws;
  GGAD_REQUIRE(d->n_peer >= 0 && d->n_peer <= 7, GGAD_ERR_INVALID, "gather_reduce: n_peer must be in [0, 7]");
  a.n_peer = d->n_peer;
  for (int p = 0; p < 7; ++p) {
    a.y_peer[p] = p < d->n_peer ? d->y_peer[p] : nullptr;
    GGAD_REQUIRE(p >= d->n_peer || (a.y_peer[p] && aligned16(a.y_peer[p])), GGAD_ERR_ALIGN, "gather_reduce: y_peer[%d] null or unaligned", p);
  }
  a.y_mc = d->y_multicast;
  a.peer_need = d->peer_need;
  a.tile_done = d->tile_done;
  a.tile_epoch = d->tile_epoch;
  a.mc_min = d->mc_min_peers;
  a.n_x_rows = d->n_x_rows;
  GGAD_REQUIRE(!a.tile_done || (d->tile_row && d->y && !a.y_mc), GGAD_ERR_INVALID,
               "gather_reduce: tile_done (chase mode) needs the merge-path plan and y, and excludes y_multicast");
  GGAD_REQUIRE((d->n_peer == 0 && !a.y_mc) || d->y, GGAD_ERR_INVALID, "gather_reduce: the fused exchange needs the local y");
  GGAD_REQUIRE(!a.peer_need || (d->n_peer > 0 && (!a.y_mc || d->mc_min_peers > 0)), GGAD_ERR_INVALID,
               "gather_reduce: peer_need needs y_peer[]; with y_multicast it needs mc_min_peers > 0 (hybrid exchange)");
  GGAD_REQUIRE(d->mc_min_peers >= 0 && (d->mc_min_peers == 0 || (a.y_mc && d->n_peer > 0)), GGAD_ERR_INVALID,
               "gather_reduce: mc_min_peers needs both y_multicast and y_peer[]");
  GGAD_REQUIRE(aligned16(a.y_mc), GGAD_ERR_ALIGN, "gather_reduce: y_multicast not 16-byte aligned");
  const int sms = sm_count_cached();
  if (sms <= 0) return GGAD_ERR_CUDA;
  return dispatch_width(a, st, sms);
}

int plan_build_impl(const int64_t* rowptr, int64_t n_rows, int64_t nnz, int32_t* tile_row, int64_t* tile_edge,
                    cudaStream_t st, int64_t n_tiles_padded) {
  GGAD_REQUIRE(rowptr && tile_row && tile_edge, GGAD_ERR_INVALID, "plan_build: null pointer");
  GGAD_REQUIRE(n_rows >= 0 && nnz >= 0 && n_rows < (int64_t(1) << 31), GGAD_ERR_INVALID, "plan_build: bad sizes");
  // a padded plan has more tiles than the CSR needs: the diagonal clamps at n_rows + nnz, so the extra tiles are empty
  GGAD_REQUIRE(n_tiles_padded == 0 || n_tiles_padded >= ggad_plan_num_tiles(n_rows, nnz), GGAD_ERR_INVALID,
               "plan_build: padded tile count smaller than the CSR needs");
  const int64_t n_tiles = n_tiles_padded ? n_tiles_padded : ggad_plan_num_tiles(n_rows, nnz);
  const int64_t threads = n_tiles + 1;
  plan_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(rowptr, n_rows, nnz, n_tiles, tile_row, tile_edge);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

}  // namespace ggad
