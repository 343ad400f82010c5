// Feature rows staged through shared memory by the TMA engine -- the A/B variant of gather_tiled_kernel's
// register-direct 128-bit gathers.  Same merge-path tile, same second-level split over lane
// groups, same summation order (bit-identical results); only the way a neighbour row reaches the lane group differs:
//
//   KIND 1  cp.async.bulk.tensor.2d ... tile::gather4: one instruction fetches the four rows named by four column
//           indices into a [4][d] shared-memory tile (tensor map over x, box {d, 1});
//   KIND 2  cp.async.bulk (1-D): one 4*d-byte bulk copy per row;
//   KIND 3  gather4 again, but the two lane groups of a warp run in lock step: one mbarrier per warp and stage, eight
//           rows per group and stage, lane 0 of the warp issues for both groups (no divergent half-warps).
//
// Every lane group owns a ring of S stages of four rows and one mbarrier per stage; lane 0 of the group issues the
// copies S batches ahead, the group waits on the stage's barrier, adds the rows from shared memory (LDS.128) and hands
// the stage back.  Rows in flight per SM = CTAs/SM x 16 groups x 4 S, without a register per byte in flight.
// Selected at run time by GGAD_TMA_ROWS=1|2|3 (GGAD_TMA_STAGES=1..5; 1..3 for kind 3); GGAD_TMA_ROWS=4 mixes the
// two request paths inside every CTA (gather_mix_kernel below) for the plain launch at d = 64; off by default.
#include <cuda.h>
#include <string.h>

#include "gather_kernels.cuh"

namespace ggad {

namespace {

constexpr int kG = 16;                 // lanes per group (d = 64: one float4 per lane)
constexpr int kNgrp = kThreads / kG;   // 16 groups per CTA
constexpr int kRowBytes = 256;
constexpr int kStageBytes = 4 * kRowBytes;

__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int c0, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}

template <bool HASVAL, int S, int KIND>
struct TmaSmem {
  using L = TileSmem<kG, 1, HASVAL, false>;
  static constexpr int kRowsPerStage = (KIND == 3) ? 8 : 4;           // per lane group
  static constexpr int kIdx = L::kRows;                               // end of the index staging area
  static constexpr int kBars = (kIdx + 15) & ~15;                     // uint64[kNgrp][S] (KIND 3 uses one per warp and stage)
  static constexpr int kRing = (kBars + kNgrp * S * 8 + 127) & ~127;  // [kNgrp][S][rows][256 B], 128-byte aligned
  static constexpr int kBytes = kRing + kNgrp * S * kRowsPerStage * kRowBytes;
};

template <int MODE, int S, int KIND>
__global__ void __launch_bounds__(kThreads) gather_tma_kernel(const __grid_constant__ GatherArgs a,
                                                              const __grid_constant__ CUtensorMap xmap) {
  using T = TmaSmem<MODE != 0, S, KIND>;
  using L = typename T::L;
  constexpr int G = kG, NGRP = kNgrp;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
  int32_t* s_col = reinterpret_cast<int32_t*>(smem + L::kCol);
  float* s_val = reinterpret_cast<float*>(smem + L::kVal);
  int32_t* s_rend = reinterpret_cast<int32_t*>(smem + L::kRend);
  int32_t* s_ci = reinterpret_cast<int32_t*>(smem + L::kCi);
  int32_t* s_cj = reinterpret_cast<int32_t*>(smem + L::kCj);
  int32_t* s_flag = reinterpret_cast<int32_t*>(smem + L::kFlag);
  float4* s_part = reinterpret_cast<float4*>(smem + L::kPart);
  uint64_t* s_rbar = reinterpret_cast<uint64_t*>(smem + T::kBars);
  unsigned char* s_ring = smem + T::kRing;

  const int tid = threadIdx.x;
  const int64_t k = blockIdx.x;
  const int64_t r0 = __ldg(a.tile_row + k), r1 = __ldg(a.tile_row + k + 1);
  const int64_t e0 = __ldg(a.tile_edge + k), e1 = __ldg(a.tile_edge + k + 1);
  const int nr = int(r1 - r0), ne = int(e1 - e0);
  constexpr bool has_val = MODE != 0;

  // ---- stage the CSR slice exactly as gather_tiled_kernel does ----
  const int64_t e0a = e0 & ~int64_t(3);
  const int lead = int(e0 - e0a);
  const int cnt = ne + lead;
  int nb = (cnt + 3) & ~3;
  if (e0a + nb > a.nnz) nb = cnt & ~3;
  if (tid == 0) mbar_init(s_bar, 1);
  if (tid < NGRP * S) mbar_init(s_rbar + tid, 1);
  fence_mbar_init();
  __syncthreads();
  if (tid == 0 && nb > 0) {
    mbar_arrive_expect_tx(s_bar, uint32_t(nb) * 4u * (has_val ? 2u : 1u));
    tma_bulk_g2s(s_col, a.col + e0a, uint32_t(nb) * 4u, s_bar);
    if (has_val) tma_bulk_g2s(s_val, a.val + e0a, uint32_t(nb) * 4u, s_bar);
  }
  for (int t = nb + tid; t < cnt; t += kThreads) {
    s_col[t] = __ldg(a.col + e0a + t);
    if (has_val) s_val[t] = __ldg(a.val + e0a + t);
  }
  for (int j = tid; j <= nr; j += kThreads) {
    const int64_t rr = r0 + j + 1;
    int64_t v = (rr <= a.n_rows) ? (__ldg(a.rowptr + rr) - e0) : int64_t(kBig);
    s_rend[j] = v > kBig ? kBig : int(v);
  }
  const int rstart0 = (r0 < a.n_rows) ? int(__ldg(a.rowptr + r0) - e0) : 0;
  __syncthreads();

  const int items = nr + ne;
  const int ipg = (items + NGRP - 1) / NGRP;
  if (tid <= NGRP) {
    int diag = tid * ipg;
    if (diag > items) diag = items;
    int lo = diag > ne ? diag - ne : 0;
    int hi = diag < nr ? diag : nr;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_rend[mid] <= diag - mid - 1) lo = mid + 1;
      else hi = mid;
    }
    s_ci[tid] = lo;
    s_cj[tid] = diag - lo;
  }
  __syncthreads();
  if (nb > 0) mbar_wait(s_bar, 0);

  const int g = tid / G, gl = tid % G;
  const unsigned gmask = group_mask<G>();
  float4* my_part = s_part + (g * 2) * G;
  uint64_t* my_bar = s_rbar + g * S;
  unsigned char* my_ring = s_ring + g * S * kStageBytes;
  const EpiRegs ep{nullptr, 0.f, 0};

  if constexpr (KIND == 3) {
    // Warp-converged form: the two lane groups of a warp run in lock step on one mbarrier per stage; a stage holds
    // eight rows per group (two gather4 copies each), issued by lane 0 of the warp for both groups.
    constexpr int UB = 8;
    constexpr int kHalf = UB * kRowBytes;                      // one group's part of a stage
    const int warp = tid >> 5, lane = tid & 31;
    const int i1 = s_ci[g], j1 = s_cj[g], i2 = s_ci[g + 1], j2 = s_cj[g + 1];
    const int oj1 = s_cj[g ^ 1], oj2 = s_cj[(g ^ 1) + 1];      // the other group of this warp
    const int nbat = (j2 - j1 + UB - 1) / UB, onbat = (oj2 - oj1 + UB - 1) / UB;
    const int nit = nbat > onbat ? nbat : onbat;               // warp-uniform trip count
    uint64_t* wbar = s_rbar + warp * S;
    unsigned char* wring = s_ring + warp * S * 2 * kHalf;
    int row = i1;
    int cur_end = s_rend[row];
    bool head_pending = ((i1 == 0) ? rstart0 : s_rend[i1 - 1]) < j1;
    int flag = 0;
    float4 acc[1] = {f4_zero()};
    const int32_t* __restrict__ sc = s_col + lead;
    const float* __restrict__ sv = s_val + lead;

    auto flush = [&]() {
      if (head_pending) {
        my_part[gl] = acc[0];
        flag |= 1;
        head_pending = false;
      } else {
        finish_row<G, 1, 0, false>(a, r0 + row, acc, gl, gmask, ep);
      }
      acc[0] = f4_zero();
      ++row;
      cur_end = s_rend[row];
    };
    // lane 0 (group 2 * warp): its own range is (j1, j2), the partner's (oj1, oj2)
    auto issue = [&](int it) {
      const int st = it % S;
      unsigned char* dst = wring + st * 2 * kHalf;
      uint32_t bytes = 0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = (h ? oj1 : j1) + UB * it, end = h ? oj2 : j2;
        if (e < end) bytes += (end - e > 4) ? 2u * kStageBytes : uint32_t(kStageBytes);
      }
      mbar_arrive_expect_tx(wbar + st, bytes);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int e = (h ? oj1 : j1) + UB * it, end = h ? oj2 : j2;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int eq = e + 4 * q;
          if (eq < end) {
            const int c0 = sc[eq];
            const int c1 = (eq + 1 < end) ? sc[eq + 1] : c0;
            const int c2 = (eq + 2 < end) ? sc[eq + 2] : c0;
            const int c3 = (eq + 3 < end) ? sc[eq + 3] : c0;
            tma_gather4(dst + h * kHalf + q * kStageBytes, &xmap, 0, c0, c1, c2, c3, wbar + st);
          }
        }
      }
    };
    if (lane == 0)
      for (int it = 0; it < S && it < nit; ++it) issue(it);

    for (int it = 0; it < nit; ++it) {
      const int st = it % S;
      mbar_wait(wbar + st, uint32_t(it / S) & 1u);
      if (it < nbat) {
        const float4* rp = reinterpret_cast<const float4*>(wring + st * 2 * kHalf + (g & 1) * kHalf) + gl;
        const int e = j1 + UB * it;
        float4 xv[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) xv[u] = rp[u * (kRowBytes / 16)];
        if (e + UB <= cur_end && e + UB <= j2) {
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            if (MODE == 0) f4_add(acc[0], xv[u]);
            else f4_fma(acc[0], sv[e + u], xv[u]);
          }
        } else {
#pragma unroll
          for (int u = 0; u < UB; ++u) {
            if (e + u < j2) {
              while (e + u >= cur_end) flush();
              if (MODE == 0) f4_add(acc[0], xv[u]);
              else f4_fma(acc[0], sv[e + u], xv[u]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0 && it + S < nit) issue(it + S);
    }
    while (row < i2) flush();
    const int rs_tail = (i2 == 0) ? rstart0 : s_rend[i2 - 1];
    if (j2 > (rs_tail > j1 ? rs_tail : j1)) {
      my_part[G + gl] = acc[0];
      flag |= 2;
    }
    if (gl == 0) s_flag[g] = flag;
  } else {
    const int i1 = s_ci[g], j1 = s_cj[g], i2 = s_ci[g + 1], j2 = s_cj[g + 1];
    int row = i1;
    int cur_end = s_rend[row];
    bool head_pending = ((i1 == 0) ? rstart0 : s_rend[i1 - 1]) < j1;
    int flag = 0;
    float4 acc[1] = {f4_zero()};
    const int32_t* __restrict__ sc = s_col + lead;
    const float* __restrict__ sv = s_val + lead;
    const int nbat = (j2 - j1 + 3) >> 2;

    auto flush = [&]() {
      if (head_pending) {
        my_part[gl] = acc[0];
        flag |= 1;
        head_pending = false;
      } else {
        finish_row<G, 1, 0, false>(a, r0 + row, acc, gl, gmask, ep);
      }
      acc[0] = f4_zero();
      ++row;
      cur_end = s_rend[row];
    };
    // lane 0 of the group: fetch the four rows of batch `it` into its stage (the last batch repeats its last column)
    auto issue = [&](int it) {
      const int e = j1 + 4 * it;
      const int st = it % S;
      const int c0 = sc[e];
      const int c1 = (e + 1 < j2) ? sc[e + 1] : c0;
      const int c2 = (e + 2 < j2) ? sc[e + 2] : c0;
      const int c3 = (e + 3 < j2) ? sc[e + 3] : c0;
      unsigned char* dst = my_ring + st * kStageBytes;
      mbar_arrive_expect_tx(my_bar + st, kStageBytes);
      if constexpr (KIND == 1) {
        tma_gather4(dst, &xmap, 0, c0, c1, c2, c3, my_bar + st);
      } else {
        const char* xb = reinterpret_cast<const char*>(a.x);
        const uint64_t pitch = uint64_t(a.ldx) * 4u;
        tma_bulk_g2s(dst, xb + uint64_t(uint32_t(c0)) * pitch, kRowBytes, my_bar + st);
        tma_bulk_g2s(dst + kRowBytes, xb + uint64_t(uint32_t(c1)) * pitch, kRowBytes, my_bar + st);
        tma_bulk_g2s(dst + 2 * kRowBytes, xb + uint64_t(uint32_t(c2)) * pitch, kRowBytes, my_bar + st);
        tma_bulk_g2s(dst + 3 * kRowBytes, xb + uint64_t(uint32_t(c3)) * pitch, kRowBytes, my_bar + st);
      }
    };
    if (gl == 0)
      for (int it = 0; it < S && it < nbat; ++it) issue(it);

    for (int it = 0; it < nbat; ++it) {
      const int st = it % S;
      mbar_wait(my_bar + st, uint32_t(it / S) & 1u);
      const float4* rp = reinterpret_cast<const float4*>(my_ring + st * kStageBytes) + gl;
      float4 xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xv[u] = rp[u * (kRowBytes / 16)];
      const int e = j1 + 4 * it;
      if (e + 4 <= cur_end && e + 4 <= j2) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (MODE == 0) f4_add(acc[0], xv[u]);
          else f4_fma(acc[0], sv[e + u], xv[u]);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (e + u < j2) {
            while (e + u >= cur_end) flush();
            if (MODE == 0) f4_add(acc[0], xv[u]);
            else f4_fma(acc[0], sv[e + u], xv[u]);
          }
        }
      }
      // every lane has consumed its registers' worth of the stage (the adds above depend on the loads)
      __syncwarp(gmask);
      if (gl == 0 && it + S < nbat) issue(it + S);
    }
    while (row < i2) flush();
    const int rs_tail = (i2 == 0) ? rstart0 : s_rend[i2 - 1];
    if (j2 > (rs_tail > j1 ? rs_tail : j1)) {
      my_part[G + gl] = acc[0];
      flag |= 2;
    }
    if (gl == 0) s_flag[g] = flag;
  }
  __syncthreads();

  // ---- rows cut by group boundaries, in group order (as in gather_tiled_kernel) ----
  if (g == 0) {
    const int V = a.d >> 2;
    float4 chain[1] = {f4_zero()};
    float* ws_head = a.ws + (2 * k) * int64_t(a.d);
    float* ws_tail = ws_head + a.d;
    for (int q = 0; q < NGRP; ++q) {
      const int f = s_flag[q];
      const float4* part = s_part + (q * 2) * G;
      if (f & 1) {
        f4_add(chain[0], part[gl]);
        const int row = s_ci[q];
        if (row == 0 && rstart0 < 0) {
          if (gl < V) reinterpret_cast<float4*>(ws_head)[gl] = chain[0];
        } else {
          finish_row<G, 1, 0, false>(a, r0 + row, chain, gl, gmask, ep);
        }
        chain[0] = f4_zero();
      }
      if (f & 2) f4_add(chain[0], part[G + gl]);
    }
    if (gl < V) reinterpret_cast<float4*>(ws_tail)[gl] = chain[0];
  }
}

// ---------------------------------------------------------------------------
// Mixed request paths (GGAD_TMA_ROWS=4): in every CTA the first `nt` warps fetch their rows through the TMA engine
// (warp-converged gather4 ring, as KIND 3 above) while the other warps keep the register-direct 128-bit gathers of
// gather_tiled_kernel.  The question it answers: are the bytes in flight per SM bounded per request path (then the two
// windows add up) or further out?  The second-level split gives a TMA lane group `wq`/16 of the items a register-direct
// group gets, so both kinds of warp finish together; group boundaries -- hence the summation order of the few rows cut
// by them -- differ from gather_tiled_kernel, so the result equals the oracle within rounding, not bit for bit.
// ---------------------------------------------------------------------------
template <bool HASVAL>
struct MixSmem {
  using L = TileSmem<kG, 1, HASVAL, false>;
  static constexpr int kBars = (L::kRows + 15) & ~15;              // uint64[8 warps][kMaxStages]
  static constexpr int kMaxStages = 4;
  static constexpr int kRing = (kBars + 8 * kMaxStages * 8 + 127) & ~127;
  static int bytes(int nt, int stages) { return kRing + nt * stages * 16 * kRowBytes; }
};

template <int MODE>
__global__ void __launch_bounds__(kThreads, 4) gather_mix_kernel(const __grid_constant__ GatherArgs a,
                                                                 const __grid_constant__ CUtensorMap xmap, int nt, int S,
                                                                 int wq) {
  using T = MixSmem<MODE != 0>;
  using L = typename T::L;
  constexpr int G = kG, NGRP = kNgrp, UB = 8, kHalf = UB * kRowBytes;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
  int32_t* s_col = reinterpret_cast<int32_t*>(smem + L::kCol);
  float* s_val = reinterpret_cast<float*>(smem + L::kVal);
  int32_t* s_rend = reinterpret_cast<int32_t*>(smem + L::kRend);
  int32_t* s_ci = reinterpret_cast<int32_t*>(smem + L::kCi);
  int32_t* s_cj = reinterpret_cast<int32_t*>(smem + L::kCj);
  int32_t* s_flag = reinterpret_cast<int32_t*>(smem + L::kFlag);
  float4* s_part = reinterpret_cast<float4*>(smem + L::kPart);
  uint64_t* s_rbar = reinterpret_cast<uint64_t*>(smem + T::kBars);
  unsigned char* s_ring = smem + T::kRing;

  const int tid = threadIdx.x;
  const int64_t k = blockIdx.x;
  const int64_t r0 = __ldg(a.tile_row + k), r1 = __ldg(a.tile_row + k + 1);
  const int64_t e0 = __ldg(a.tile_edge + k), e1 = __ldg(a.tile_edge + k + 1);
  const int nr = int(r1 - r0), ne = int(e1 - e0);
  constexpr bool has_val = MODE != 0;

  const int64_t e0a = e0 & ~int64_t(3);
  const int lead = int(e0 - e0a);
  const int cnt = ne + lead;
  int nb = (cnt + 3) & ~3;
  if (e0a + nb > a.nnz) nb = cnt & ~3;
  if (tid == 0) mbar_init(s_bar, 1);
  if (tid < 8 * T::kMaxStages) mbar_init(s_rbar + tid, 1);
  fence_mbar_init();
  __syncthreads();
  if (tid == 0 && nb > 0) {
    mbar_arrive_expect_tx(s_bar, uint32_t(nb) * 4u * (has_val ? 2u : 1u));
    tma_bulk_g2s(s_col, a.col + e0a, uint32_t(nb) * 4u, s_bar);
    if (has_val) tma_bulk_g2s(s_val, a.val + e0a, uint32_t(nb) * 4u, s_bar);
  }
  for (int t = nb + tid; t < cnt; t += kThreads) {
    s_col[t] = __ldg(a.col + e0a + t);
    if (has_val) s_val[t] = __ldg(a.val + e0a + t);
  }
  for (int j = tid; j <= nr; j += kThreads) {
    const int64_t rr = r0 + j + 1;
    int64_t v = (rr <= a.n_rows) ? (__ldg(a.rowptr + rr) - e0) : int64_t(kBig);
    s_rend[j] = v > kBig ? kBig : int(v);
  }
  const int rstart0 = (r0 < a.n_rows) ? int(__ldg(a.rowptr + r0) - e0) : 0;
  __syncthreads();

  // weighted second-level merge path: groups 0 .. 2 nt - 1 (the TMA warps) weigh wq, the others 16
  const int items = nr + ne;
  if (tid <= NGRP) {
    const int ntg = 2 * nt;
    const int total = ntg * wq + (NGRP - ntg) * 16;
    const int cum = tid <= ntg ? tid * wq : ntg * wq + (tid - ntg) * 16;
    int diag = (tid == NGRP) ? items : int(int64_t(items) * cum / total);
    int lo = diag > ne ? diag - ne : 0;
    int hi = diag < nr ? diag : nr;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_rend[mid] <= diag - mid - 1) lo = mid + 1;
      else hi = mid;
    }
    s_ci[tid] = lo;
    s_cj[tid] = diag - lo;
  }
  __syncthreads();
  if (nb > 0) mbar_wait(s_bar, 0);

  const int g = tid / G, gl = tid % G;
  const int warp = tid >> 5, lane = tid & 31;
  const unsigned gmask = group_mask<G>();
  float4* my_part = s_part + (g * 2) * G;
  const EpiRegs ep{nullptr, 0.f, 0};
  {
    const int i1 = s_ci[g], j1 = s_cj[g], i2 = s_ci[g + 1], j2 = s_cj[g + 1];
    int row = i1;
    int cur_end = s_rend[row];
    bool head_pending = ((i1 == 0) ? rstart0 : s_rend[i1 - 1]) < j1;
    int flag = 0;
    float4 acc[1] = {f4_zero()};
    const int32_t* __restrict__ sc = s_col + lead;
    const float* __restrict__ sv = s_val + lead;
    auto flush = [&]() {
      if (head_pending) {
        my_part[gl] = acc[0];
        flag |= 1;
        head_pending = false;
      } else {
        finish_row<G, 1, 0, false>(a, r0 + row, acc, gl, gmask, ep);
      }
      acc[0] = f4_zero();
      ++row;
      cur_end = s_rend[row];
    };
    auto consume = [&](int e, const float4 (&xv)[UB]) {     // eight rows starting at edge e (the tail may be shorter)
      if (e + UB <= cur_end && e + UB <= j2) {
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          if (MODE == 0) f4_add(acc[0], xv[u]);
          else f4_fma(acc[0], sv[e + u], xv[u]);
        }
      } else {
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          if (e + u < j2) {
            while (e + u >= cur_end) flush();
            if (MODE == 0) f4_add(acc[0], xv[u]);
            else f4_fma(acc[0], sv[e + u], xv[u]);
          }
        }
      }
    };

    if (warp < nt) {
      // ---- TMA warps: both lane groups in lock step, one mbarrier per stage ----
      const int oj1 = s_cj[g ^ 1], oj2 = s_cj[(g ^ 1) + 1];
      const int nbat = (j2 - j1 + UB - 1) / UB, onbat = (oj2 - oj1 + UB - 1) / UB;
      const int nit = nbat > onbat ? nbat : onbat;
      uint64_t* wbar = s_rbar + warp * T::kMaxStages;
      unsigned char* wring = s_ring + warp * S * 2 * kHalf;
      auto issue = [&](int it, int st) {
        unsigned char* dst = wring + st * 2 * kHalf;
        uint32_t bytes = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = (h ? oj1 : j1) + UB * it, end = h ? oj2 : j2;
          if (e < end) bytes += (end - e > 4) ? 2u * kStageBytes : uint32_t(kStageBytes);
        }
        mbar_arrive_expect_tx(wbar + st, bytes);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int e = (h ? oj1 : j1) + UB * it, end = h ? oj2 : j2;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int eq = e + 4 * q;
            if (eq < end) {
              const int c0 = sc[eq];
              const int c1 = (eq + 1 < end) ? sc[eq + 1] : c0;
              const int c2 = (eq + 2 < end) ? sc[eq + 2] : c0;
              const int c3 = (eq + 3 < end) ? sc[eq + 3] : c0;
              tma_gather4(dst + h * kHalf + q * kStageBytes, &xmap, 0, c0, c1, c2, c3, wbar + st);
            }
          }
        }
      };
      if (lane == 0)
        for (int it = 0; it < S && it < nit; ++it) issue(it, it);
      int st = 0;
      uint32_t ph = 0;
      for (int it = 0; it < nit; ++it) {
        mbar_wait(wbar + st, ph);
        if (it < nbat) {
          const float4* rp = reinterpret_cast<const float4*>(wring + st * 2 * kHalf + (g & 1) * kHalf) + gl;
          float4 xv[UB];
#pragma unroll
          for (int u = 0; u < UB; ++u) xv[u] = rp[u * (kRowBytes / 16)];
          consume(j1 + UB * it, xv);
        }
        __syncwarp();
        if (lane == 0 && it + S < nit) issue(it + S, st);
        if (++st == S) {
          st = 0;
          ph ^= 1u;
        }
      }
    } else {
      // ---- register-direct warps: eight independent 128-bit gathers per lane in flight ----
      const char* lane_base = reinterpret_cast<const char*>(a.x) + uint32_t(gl) * 16u;
      asm volatile("" : "+l"(lane_base));
      const uint32_t row_bytes = uint32_t(a.ldx) * 4u;
      for (int e = j1; e < j2; e += UB) {
        float4 xv[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int c = (e + u < j2) ? sc[e + u] : sc[e];
          xv[u] = ldg_f4(reinterpret_cast<const float4*>(lane_base + uint64_t(uint32_t(c)) * row_bytes));
        }
        consume(e, xv);
      }
    }
    while (row < i2) flush();
    const int rs_tail = (i2 == 0) ? rstart0 : s_rend[i2 - 1];
    if (j2 > (rs_tail > j1 ? rs_tail : j1)) {
      my_part[G + gl] = acc[0];
      flag |= 2;
    }
    if (gl == 0) s_flag[g] = flag;
  }
  __syncthreads();

  if (g == 0) {
    const int V = a.d >> 2;
    float4 chain[1] = {f4_zero()};
    float* ws_head = a.ws + (2 * k) * int64_t(a.d);
    float* ws_tail = ws_head + a.d;
    for (int q = 0; q < NGRP; ++q) {
      const int f = s_flag[q];
      const float4* part = s_part + (q * 2) * G;
      if (f & 1) {
        f4_add(chain[0], part[gl]);
        const int row = s_ci[q];
        if (row == 0 && rstart0 < 0) {
          if (gl < V) reinterpret_cast<float4*>(ws_head)[gl] = chain[0];
        } else {
          finish_row<G, 1, 0, false>(a, r0 + row, chain, gl, gmask, ep);
        }
        chain[0] = f4_zero();
      }
      if (f & 2) f4_add(chain[0], part[G + gl]);
    }
    if (gl < V) reinterpret_cast<float4*>(ws_tail)[gl] = chain[0];
  }
}

template <int MODE>
int launch_mix(const GatherArgs& a, const CUtensorMap& map, cudaStream_t st, int nt, int stages, int wq) {
  using T = MixSmem<MODE != 0>;
  const int bytes = T::bytes(nt, stages);
  // the ring size is a run-time choice here, so the attribute is set on every launch (an A/B tool, not a hot path)
  GGAD_CUDA_OK(cudaFuncSetAttribute(gather_mix_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  gather_mix_kernel<MODE><<<(unsigned)a.n_tiles, kThreads, bytes, st>>>(a, map, nt, stages, wq);
  GGAD_CUDA_OK(cudaGetLastError());
  const int64_t fix_blocks = (a.n_tiles * kG + kThreads - 1) / kThreads;
  tile_fixup_kernel<kG, 1, false, false><<<(unsigned)fix_blocks, kThreads, 0, st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return GGAD_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

template <int MODE, int S, int KIND>
int launch_tma(const GatherArgs& a, const CUtensorMap& map, cudaStream_t st) {
  using T = TmaSmem<MODE != 0, S, KIND>;
  static bool attr_done[64] = {};
  int dev = 0;
  GGAD_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    GGAD_CUDA_OK(cudaFuncSetAttribute(gather_tma_kernel<MODE, S, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::kBytes));
    if (dev < 64) attr_done[dev] = true;
  }
  gather_tma_kernel<MODE, S, KIND><<<(unsigned)a.n_tiles, kThreads, T::kBytes, st>>>(a, map);
  GGAD_CUDA_OK(cudaGetLastError());
  const int64_t fix_blocks = (a.n_tiles * kG + kThreads - 1) / kThreads;
  tile_fixup_kernel<kG, 1, false, false><<<(unsigned)fix_blocks, kThreads, 0, st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return GGAD_OK;
}

template <int MODE, int KIND>
int launch_tma_stages(const GatherArgs& a, const CUtensorMap& map, cudaStream_t st, int stages) {
  switch (stages) {
    case 1: return launch_tma<MODE, 1, KIND>(a, map, st);
    case 2: return launch_tma<MODE, 2, KIND>(a, map, st);
    case 3: return launch_tma<MODE, 3, KIND>(a, map, st);
    case 5: return launch_tma<MODE, 5, KIND>(a, map, st);
    default: return launch_tma<MODE, 4, KIND>(a, map, st);
  }
}

}  // namespace

// Returns GGAD_OK after launching, an error code, or -1 when the variant is not selected / does not apply
// (the caller then runs gather_tiled_kernel).
int try_launch_tma_rows(const GatherArgs& a, cudaStream_t st) {
  // read per launch (an A/B knob, not a tuned default): tests and tools flip it inside one process
  const char* ek = getenv("GGAD_TMA_ROWS");
  const int kind = ek ? atoi(ek) : 0;
  if (kind < 1 || kind > 4) return -1;
  const char* es = getenv("GGAD_TMA_STAGES");
  const int stages = es ? atoi(es) : 4;
  if (a.d != 64 || a.xmap || a.col_scale || !a.tile_row || !a.y || a.n_x_rows <= 0) return -1;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if (kind != 2) {
    EncodeTiledFn enc = encode_fn();
    GGAD_REQUIRE(enc != nullptr, GGAD_ERR_CUDA, "gather_reduce: cuTensorMapEncodeTiled not available");
    const cuuint64_t gdim[2] = {cuuint64_t(a.d), cuuint64_t(a.n_x_rows)};
    const cuuint64_t gstride[1] = {cuuint64_t(a.ldx) * 4u};
    const cuuint32_t box[2] = {cuuint32_t(a.d), 1u};
    const cuuint32_t estr[2] = {1u, 1u};
    const char* ep = getenv("GGAD_TMA_L2_PROMOTION");   // 0 none, 1 64 B, 2 128 B, 3 256 B (a row is 256 B)
    const int promo = ep ? atoi(ep) : 3;
    const CUtensorMapL2promotion l2p = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                       : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    const CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(a.x), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GGAD_REQUIRE(r == CUDA_SUCCESS, GGAD_ERR_CUDA, "gather_reduce: cuTensorMapEncodeTiled failed (%d)", int(r));
    if (kind == 4) {
      const char* en = getenv("GGAD_TMA_WARPS");    // warps per CTA on the TMA path (0..8), the rest gather from registers
      const char* ew = getenv("GGAD_TMA_WEIGHT");   // items of a TMA lane group in 1/16 of a register-direct group's
      int nt = en ? atoi(en) : 2, wq = ew ? atoi(ew) : 10, sg = stages;
      nt = nt < 0 ? 0 : (nt > 8 ? 8 : nt);
      wq = wq < 1 ? 1 : (wq > 64 ? 64 : wq);
      sg = sg < 1 ? 1 : (sg > 4 ? 4 : sg);
      return a.val ? launch_mix<1>(a, map, st, nt, sg, wq) : launch_mix<0>(a, map, st, nt, sg, wq);
    }
    if (kind == 3) {
      const int s3 = stages < 1 ? 1 : (stages > 3 ? 3 : stages);
      if (a.val) return s3 == 1 ? launch_tma<1, 1, 3>(a, map, st) : (s3 == 2 ? launch_tma<1, 2, 3>(a, map, st) : launch_tma<1, 3, 3>(a, map, st));
      return s3 == 1 ? launch_tma<0, 1, 3>(a, map, st) : (s3 == 2 ? launch_tma<0, 2, 3>(a, map, st) : launch_tma<0, 3, 3>(a, map, st));
    }
    return a.val ? launch_tma_stages<1, 1>(a, map, st, stages) : launch_tma_stages<0, 1>(a, map, st, stages);
  }
  return a.val ? launch_tma_stages<1, 2>(a, map, st, stages) : launch_tma_stages<0, 2>(a, map, st, stages);
}

}  // namespace ggad
