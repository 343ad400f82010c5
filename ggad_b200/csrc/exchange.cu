// Halo exchange of a locally produced row block over NVLink peer memory (stand-alone kernel).
//
// The gather kernel pushes the rows it finishes itself (gather_kernels.cuh: push_rows); this kernel does the
// same for a matrix that was NOT produced by a gather launch -- e.g. a rank's shard of the layer-0 features
// that just arrived from the host -- so that every layer input reaches exactly the ranks whose CSR shard
// gathers it, without a collective.  One thread moves one 16-byte chunk; four independent chunks per thread
// are in flight.  Bound: NVLink egress (bytes = sum over rows of popcount(need) * d * 4).
#include "common.cuh"

namespace ggad {

struct PushArgs {
  const float* y;
  int64_t ldy, n_rows;
  int32_t d, n_peer;
  const uint32_t* need;
  float* peer[7];
};

__global__ void __launch_bounds__(256) halo_push_kernel(const __grid_constant__ PushArgs a) {
  const int V = a.d >> 2;
  const int64_t total = a.n_rows * V;
  const uint32_t all = (1u << a.n_peer) - 1u;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  constexpr int U = 4;
  for (int64_t i0 = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i0 < total; i0 += U * stride) {
    float4 v[U];
    uint32_t nd[U];
    int64_t off[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      nd[u] = 0u;
      if (i < total) {
        const int64_t r = i / V;
        nd[u] = (a.need ? __ldg(a.need + r) : 0xffffffffu) & all;
        off[u] = r * a.ldy + (i - r * V) * 4;
        if (nd[u]) v[u] = __ldcs(reinterpret_cast<const float4*>(a.y + off[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      for (int p = 0; p < a.n_peer; ++p)
        if ((nd[u] >> p) & 1u) stg_peer_f4(reinterpret_cast<float4*>(a.peer[p] + off[u]), v[u]);
  }
}

int sm_count_cached();

int halo_push_impl(const float* y, int64_t ldy, int64_t n_rows, int32_t d, const uint32_t* need, float* const* peers,
                   int32_t n_peer, cudaStream_t st) {
  GGAD_REQUIRE(n_rows >= 0 && d > 0 && d % 4 == 0 && ldy % 4 == 0 && ldy >= d, GGAD_ERR_INVALID, "halo_push: bad shape");
  GGAD_REQUIRE(n_peer >= 0 && n_peer <= 7, GGAD_ERR_INVALID, "halo_push: n_peer must be in [0, 7]");
  if (n_rows == 0 || n_peer == 0) return GGAD_OK;
  GGAD_REQUIRE(y && peers && aligned16(y), GGAD_ERR_ALIGN, "halo_push: y null or not 16-byte aligned");
  PushArgs a;
  a.y = y; a.ldy = ldy; a.n_rows = n_rows; a.d = d; a.n_peer = n_peer; a.need = need;
  for (int p = 0; p < 7; ++p) {
    a.peer[p] = p < n_peer ? peers[p] : nullptr;
    GGAD_REQUIRE(p >= n_peer || (a.peer[p] && aligned16(a.peer[p])), GGAD_ERR_ALIGN, "halo_push: peer[%d] null or unaligned", p);
  }
  const int sms = sm_count_cached();
  if (sms <= 0) return GGAD_ERR_CUDA;
  const int64_t total = n_rows * (d >> 2);
  int64_t blocks = (total + 256 * 4 - 1) / (256 * 4);
  const int64_t cap = int64_t(sms) * 8;  // grid-stride over 8 CTAs per SM
  if (blocks > cap) blocks = cap;
  halo_push_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

// ---------------------------------------------------------------------------
// chase mode: the exchange as its own persistent kernel, fed by the gather kernel's tile-done flags
// ---------------------------------------------------------------------------
struct ChaseArgs {
  const float* y;
  int64_t ldy, n_rows, n_tiles;
  int32_t d, n_peer, epoch, mc_min;
  const int64_t* rowptr;
  const int32_t* tile_row;
  const int64_t* tile_edge;
  const int32_t* tile_done;
  const uint32_t* need;
  float* peer[7];
  float* y_mc;
};

// A CTA takes batches of kChaseTiles consecutive tiles (batch b, b + gridDim, ...): its first threads wait for the
// batch's flags in parallel, the masks of the rows those tiles finished are staged in shared memory and the needed
// rows compacted into a list (warp ballots), then the 256 threads stream the listed rows L2 -> registers -> peers,
// kChaseU rows per lane group in flight.  Per batch: three dependent L2 round trips + the copy itself, so a few dozen
// CTAs keep up with the gather kernel.  Bound: NVLink egress.
constexpr int kChaseTiles = 8;
constexpr int kChaseRows = 2048;  // rows staged per round (a batch with more row ends takes several rounds)
constexpr int kChaseU = 4;

__global__ void __launch_bounds__(256, 4) halo_chase_kernel(const __grid_constant__ ChaseArgs a) {
  __shared__ uint32_t s_need[kChaseRows];
  __shared__ uint16_t s_list[kChaseRows];
  __shared__ int s_cnt;
  const int tid = threadIdx.x, lane = tid & 31;
  const int V = a.d >> 2;
  // lanes per row: the smallest power of two >= V (capped at 32; wider rows loop over their chunks)
  int L = 1;
  while (L < V && L < 32) L <<= 1;
  const int rows_per_step = 256 / L;
  const int sub = tid / L, sl = tid % L;
  const uint32_t all = (1u << a.n_peer) - 1u;
  for (int64_t kb = int64_t(blockIdx.x) * kChaseTiles; kb < a.n_tiles; kb += int64_t(gridDim.x) * kChaseTiles) {
    const int64_t ke = (kb + kChaseTiles < a.n_tiles) ? kb + kChaseTiles : a.n_tiles;
    if (tid < int(ke - kb)) {
      uint32_t spins = 0;
      while (ld_acquire_gpu(a.tile_done + kb + tid) != a.epoch) {
        __nanosleep(100);
        if (++spins > (1u << 24)) asm volatile("trap;");  // seconds: the producer launch is missing -- fail, don't hang
      }
    }
    __syncthreads();
    const int64_t rbase = __ldg(a.tile_row + kb), rend = __ldg(a.tile_row + ke);
    for (int64_t rr = rbase; rr < rend; rr += kChaseRows) {
      const int nrow = int((rend - rr < kChaseRows) ? rend - rr : kChaseRows);
      if (tid == 0) s_cnt = 0;
      for (int j = tid; j < nrow; j += 256) s_need[j] = (a.need ? __ldg(a.need + rr + j) : 0xffffffffu) & all;
      __syncthreads();
      // a tile's first row that began in an earlier tile is finished (and pushed) by the fix-up kernel, not here
      if (tid < int(ke - kb)) {
        const int64_t r = __ldg(a.tile_row + kb + tid);
        if (r >= rr && r < rr + nrow && r < a.n_rows && __ldg(a.rowptr + r) < __ldg(a.tile_edge + kb + tid)) s_need[r - rr] = 0u;
      }
      __syncthreads();
      for (int j0 = 0; j0 < nrow; j0 += 256) {   // compact the needed rows (order is irrelevant)
        const int j = j0 + tid;
        const bool f = j < nrow && s_need[j] != 0u;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_cnt, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (f) s_list[base + __popc(bal & ((1u << lane) - 1u))] = uint16_t(j);
      }
      __syncthreads();
      const int cnt = s_cnt;
      for (int t0 = sub; t0 < cnt; t0 += rows_per_step * kChaseU) {
        for (int c0 = sl; c0 < V; c0 += L) {      // one trip unless the row is wider than 32 chunks
          float4 v[kChaseU];
          uint32_t nd[kChaseU];
          int64_t off[kChaseU];
#pragma unroll
          for (int u = 0; u < kChaseU; ++u) {
            const int t = t0 + u * rows_per_step;
            nd[u] = 0u;
            if (t < cnt) {
              const int j = s_list[t];
              nd[u] = s_need[j];
              off[u] = (rr + j) * a.ldy + int64_t(c0) * 4;
              // L2-coherent load: written by another SM moments ago, made visible by the acquire above
              v[u] = __ldcg(reinterpret_cast<const float4*>(a.y + off[u]));
            }
          }
#pragma unroll
          for (int u = 0; u < kChaseU; ++u) {
            if (!nd[u]) continue;
            if (a.y_mc && __popc(nd[u]) >= a.mc_min) {
              asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.y_mc + off[u]), "f"(v[u].x),
                           "f"(v[u].y), "f"(v[u].z), "f"(v[u].w)
                           : "memory");
            } else {
              for (int p = 0; p < a.n_peer; ++p)
                if ((nd[u] >> p) & 1u) stg_peer_f4(reinterpret_cast<float4*>(a.peer[p] + off[u]), v[u]);
            }
          }
        }
      }
      __syncthreads();   // s_need / s_list are reused by the next round / batch
    }
  }
}

int halo_chase_impl(const ggad_chase_desc_t* d, cudaStream_t st) {
  GGAD_REQUIRE(d != nullptr, GGAD_ERR_INVALID, "halo_chase: null descriptor");
  GGAD_REQUIRE(d->n_rows >= 0 && d->n_tiles >= 0 && d->d > 0 && d->d % 4 == 0 && d->ldy % 4 == 0 && d->ldy >= d->d,
               GGAD_ERR_INVALID, "halo_chase: bad shape");
  GGAD_REQUIRE(d->n_peer >= 0 && d->n_peer <= 7, GGAD_ERR_INVALID, "halo_chase: n_peer must be in [0, 7]");
  if (d->n_rows == 0 || d->n_tiles == 0 || (d->n_peer == 0 && !d->y_multicast)) return GGAD_OK;
  GGAD_REQUIRE(d->y && d->rowptr && d->tile_row && d->tile_edge && d->tile_done, GGAD_ERR_INVALID, "halo_chase: null pointer");
  GGAD_REQUIRE(aligned16(d->y) && aligned16(d->y_multicast), GGAD_ERR_ALIGN, "halo_chase: y / y_multicast not 16-byte aligned");
  ChaseArgs a;
  a.y = d->y; a.ldy = d->ldy; a.n_rows = d->n_rows; a.n_tiles = d->n_tiles; a.d = d->d; a.n_peer = d->n_peer;
  a.epoch = d->tile_epoch; a.mc_min = d->mc_min_peers > 0 ? d->mc_min_peers : 0x7fffffff;
  a.rowptr = d->rowptr; a.tile_row = d->tile_row; a.tile_edge = d->tile_edge; a.tile_done = d->tile_done;
  a.need = d->peer_need; a.y_mc = d->y_multicast;
  for (int p = 0; p < 7; ++p) {
    a.peer[p] = p < d->n_peer ? d->y_peer[p] : nullptr;
    GGAD_REQUIRE(p >= d->n_peer || (a.peer[p] && aligned16(a.peer[p])), GGAD_ERR_ALIGN, "halo_chase: y_peer[%d] null or unaligned", p);
  }
  int ctas = d->n_ctas > 0 ? d->n_ctas : 48;
  if (int64_t(ctas) * kChaseTiles > d->n_tiles) ctas = int((d->n_tiles + kChaseTiles - 1) / kChaseTiles);
  halo_chase_kernel<<<(unsigned)ctas, 256, 0, st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

}  // namespace ggad
