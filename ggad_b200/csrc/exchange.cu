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

}  // namespace ggad
