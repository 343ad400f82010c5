// CSR neighbor gather-reduce with fused per-row epilogue (K1/K2/K3/K5 of DESIGN.md).
//
// Two kernels share one descriptor (ggad_gather_desc_t, include/ggad_b200.h):
//
//  * gather_tiled_kernel  -- the hot kernel.  Merge-path decomposition: a CTA owns a tile of
//    GGAD_TILE_ITEMS consecutive (row-end | edge) items of the CSR, so work per CTA is constant
//    whatever the degree distribution (power-law hubs are split, empty rows cost one item).
//    The tile's col/val slice is staged into shared memory with ONE TMA bulk copy
//    (cp.async.bulk + mbarrier complete_tx); row ends are staged as int32 offsets.  The tile is
//    then split evenly (second-level merge path) over lane groups of G lanes; a group walks its
//    items in order, keeping U neighbor rows (128-bit ld.global.nc per lane) in flight, and
//    finishes rows that lie entirely inside its range directly from registers (scale, bias,
//    PReLU/ReLU, |y|^2, dot epilogue).  Rows cut by a group boundary are combined through
//    shared memory in group order; rows cut by a tile boundary go through a small global
//    workspace and tile_fixup_kernel -- fixed summation order, no float atomics.
//
//  * gather_rows_kernel   -- one lane group per row; used when no plan is supplied.
//
// Roofline: HBM-bound (<= 0.5 flop/B).  Algorithmic bytes per launch
//   nnz*(4 + 4*[val]) + (n_rows+1)*8 + n_x_rows*d*4 + n_rows*d*4      
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace ggad {

struct GatherArgs {
  const int64_t* rowptr;
  const int32_t* col;
  const float* val;
  int64_t n_rows, nnz;
  const float* x;
  int64_t ldx;
  const int32_t* xmap;
  const float* col_scale;
  const float* row_scale;
  int32_t d, relu;
  const float* bias;
  const float* prelu_slope;
  float* y;
  float* z;
  int64_t ldy;
  float* sumsq;
  const float* dot_mat;
  int64_t lddot;
  const int32_t* dot_rows;
  const float* dot_scale;
  float* dot_out;
  const int32_t* tile_row;
  const int64_t* tile_edge;
  int64_t n_tiles;
  float* ws;
  float* y_peer[7];
  int n_peer;
  float* y_mc;
  const uint32_t* peer_need;
  int32_t* tile_done;   // chase mode: tile k publishes tile_done[k] = tile_epoch instead of pushing its rows
  int32_t tile_epoch;
  int32_t mc_min;       // hybrid exchange: rows needed by >= mc_min peers take the multicast address (0: y_mc takes all)
  int32_t n_x_rows;     // rows of x (0: not given); the TMA row-staging variant bounds its tensor map with it
};

// gather_tma.cu: rows staged through shared memory by TMA (A/B variant, GGAD_TMA_ROWS); -1 = not selected / not applicable
int try_launch_tma_rows(const GatherArgs& a, cudaStream_t st);

constexpr int kThreads = 256;
constexpr int kTile = GGAD_TILE_ITEMS;
constexpr int kBig = 0x3fffffff;

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if (G == 32) return 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  return ((1u << G) - 1u) << ((lane / G) * G);
}

// Store one finished 16-byte chunk of y: local copy, NVLink P2P copies into the peers' replicated
// matrices, or a single NVSwitch multicast store (multimem.st) that lands on every rank.
template <bool PEER>
__device__ __forceinline__ void store_y(const GatherArgs& a, int64_t r, int ch, const float4& v) {
  const int64_t off = r * a.ldy + int64_t(ch) * 4;
  if (!PEER) {  // single-GPU launches carry none of the exchange code (it costs 8-20 % through code size)
    if (a.y) stg_cs_f4(reinterpret_cast<float4*>(a.y + off), v);
    return;
  }
  if (a.y) stg_cs_f4(reinterpret_cast<float4*>(a.y + off), v);
  // halo exchange: only the peers whose next pass gathers this row receive it; a row that many peers need goes ONCE
  // through the NVSwitch multicast address instead (hybrid: mc_min > 0), every row does when only y_mc is given
  const uint32_t need = (a.peer_need ? __ldg(a.peer_need + r) : 0xffffffffu) & ((1u << a.n_peer) - 1u);
  if (a.y_mc && (a.mc_min <= 0 || __popc(need) >= a.mc_min)) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.y_mc + off), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
    return;
  }
  for (int p = 0; p < a.n_peer; ++p)
    if ((need >> p) & 1u) stg_peer_f4(reinterpret_cast<float4*>(a.y_peer[p] + off), v);
}

// Per-row epilogue; executed convergently by the G lanes of one group.
// Full per-row epilogue (bias / activation / z / sumsq / dot).  It is large, so kernels that use it keep the
// number of inlined copies at two (rolled boundary path below); a non-inlined call was measured slower
// (register spills around the call sites), 17 inlined copies 2x slower (instruction cache).
template <int G, int CH, bool PEER>
__device__ __forceinline__ void finish_row_full(const GatherArgs& a, int64_t r, const float4* acc, int gl, unsigned gmask) {
  const int V = a.d >> 2;
  const float rs = a.row_scale ? __ldg(a.row_scale + r) : 1.f;
  const bool want_dot = a.dot_out != nullptr;
  const bool want_ss = a.sumsq != nullptr;
  const float4* dm = nullptr;
  if (want_dot) {
    const int64_t dr = a.dot_rows ? (int64_t)__ldg(a.dot_rows + r) : r;
    dm = reinterpret_cast<const float4*>(a.dot_mat + dr * a.lddot);
  }
  const bool prelu = a.prelu_slope != nullptr;
  const float slope = prelu ? __ldg(a.prelu_slope) : 0.f;
  float ss = 0.f, dt = 0.f;
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    const int ch = gl + G * j;
    if (ch < V) {
      float4 v = acc[j];
      v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
      if (a.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias) + ch);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      if (a.z) stg_cs_f4(reinterpret_cast<float4*>(a.z + r * a.ldy) + ch, v);
      if (prelu) {
        v.x = v.x >= 0.f ? v.x : slope * v.x;
        v.y = v.y >= 0.f ? v.y : slope * v.y;
        v.z = v.z >= 0.f ? v.z : slope * v.z;
        v.w = v.w >= 0.f ? v.w : slope * v.w;
      } else if (a.relu) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      store_y<PEER>(a, r, ch, v);
      if (want_ss) ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      if (want_dot) {
        const float4 m = __ldg(dm + ch);
        dt += v.x * m.x + v.y * m.y + v.z * m.z + v.w * m.w;
      }
    }
  }
  if (want_ss || want_dot) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
      ss += __shfl_xor_sync(gmask, ss, off);
      dt += __shfl_xor_sync(gmask, dt, off);
    }
    if (gl == 0) {
      if (want_ss) a.sumsq[r] = ss;
      if (want_dot) a.dot_out[r] = dt * (a.dot_scale ? __ldg(a.dot_scale + r) : 1.f);
    }
  }
}


// Epilogue kinds (template parameter EPI of the tiled kernel):
//   0 plain   y = row_scale * acc
//   1 light   elementwise only: + bias, optional pre-activation store z, PReLU / ReLU -- the GCN-layer launch of
//             model.py:29-35.  Small enough to be inlined at every row end like the plain one (bias comes from shared
//             memory, the slope from a register), so the launch keeps the unrolled gather loop.
//   2 full    additionally |y|^2 and the dot epilogue (group shuffles); rolled boundary walk, three inlined copies.
//   (a fourth kind -- plain loop, then every lane group re-reads its rows from L2 and applies the full epilogue in one
//   rolled loop -- was measured 26 % slower than kind 1 on the GCN-layer launch and removed: profiles/r02a_variants_S64_deferred.txt)
struct EpiRegs {
  const float* s_bias;  // shared memory, d floats (zeros when there is no bias)
  float slope;
  int act;              // 0 none, 1 ReLU, 2 PReLU
};

template <int G, int CH, bool PEER>
__device__ __forceinline__ void finish_row_light(const GatherArgs& a, int64_t r, const float4 (&acc)[CH], int gl,
                                                 const EpiRegs& ep) {
  const int V = a.d >> 2;
  const float rs = a.row_scale ? __ldg(a.row_scale + r) : 1.f;
#pragma unroll
  for (int j = 0; j < CH; ++j) {
    const int ch = gl + G * j;
    if (ch < V) {
      float4 v = acc[j];
      const float4 b = reinterpret_cast<const float4*>(ep.s_bias)[ch];
      // multiply, round, add: the same two roundings as the full epilogue and the reference (A x rounded, + bias)
      v.x = __fadd_rn(__fmul_rn(v.x, rs), b.x); v.y = __fadd_rn(__fmul_rn(v.y, rs), b.y);
      v.z = __fadd_rn(__fmul_rn(v.z, rs), b.z); v.w = __fadd_rn(__fmul_rn(v.w, rs), b.w);
      if (a.z) stg_cs_f4(reinterpret_cast<float4*>(a.z + r * a.ldy) + ch, v);
      if (ep.act == 2) {
        v.x = v.x >= 0.f ? v.x : ep.slope * v.x;
        v.y = v.y >= 0.f ? v.y : ep.slope * v.y;
        v.z = v.z >= 0.f ? v.z : ep.slope * v.z;
        v.w = v.w >= 0.f ? v.w : ep.slope * v.w;
      } else if (ep.act == 1) {
        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
      }
      store_y<PEER>(a, r, ch, v);
    }
  }
}

template <int G, int CH, int EPI = 2, bool PEER = false>
__device__ __forceinline__ void finish_row(const GatherArgs& a, int64_t r, const float4 (&acc)[CH], int gl,
                                           unsigned gmask, const EpiRegs& ep = EpiRegs{nullptr, 0.f, 0}) {
  const int V = a.d >> 2;
  if (EPI == 1) {
    finish_row_light<G, CH, PEER>(a, r, acc, gl, ep);
    return;
  }
  const float rs = a.row_scale ? __ldg(a.row_scale + r) : 1.f;
  if (EPI == 0) {  // plain SpMM: y = row_scale * acc, nothing else requested
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int ch = gl + G * j;
      if (ch < V) {
        float4 v = acc[j];
        v.x *= rs; v.y *= rs; v.z *= rs; v.w *= rs;
        store_y<PEER>(a, r, ch, v);
      }
    }
    return;
  }
  finish_row_full<G, CH, PEER>(a, r, acc, gl, gmask);
}

// Resolve one edge: column -> (row of x or -1 to skip, weight).
template <bool GEN>
__device__ __forceinline__ void resolve_edge(const GatherArgs& a, int& c, float& w) {
  if (GEN) {
    if (a.col_scale) w *= __ldg(a.col_scale + c);
    if (a.xmap) c = __ldg(a.xmap + c);
    if (w == 0.f) c = -1;  // zero-scaled columns are skipped without touching x
  }
}

// ---------------------------------------------------------------------------
// group-per-row kernel (no plan)
// ---------------------------------------------------------------------------
template <int G, int CH, bool GEN>
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(const __grid_constant__ GatherArgs a) {
  constexpr int U = (CH == 1) ? 4 : 2;
  const int V = a.d >> 2;
  const int gl = threadIdx.x % G;
  const unsigned gmask = group_mask<G>();
  const int64_t gid = (int64_t(blockIdx.x) * kThreads + threadIdx.x) / G;
  const int64_t ngroups = int64_t(gridDim.x) * kThreads / G;
  const float4* __restrict__ x4 = reinterpret_cast<const float4*>(a.x);
  const int64_t ldx4 = a.ldx >> 2;
  const bool has_val = a.val != nullptr;
  for (int64_t r = gid; r < a.n_rows; r += ngroups) {
    const int64_t e0 = __ldg(a.rowptr + r), e1 = __ldg(a.rowptr + r + 1);
    float4 acc[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) acc[j] = f4_zero();
    for (int64_t e = e0; e < e1; e += U) {
      int c[U];
      float w[U];
      float4 xv[U][CH];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = e + u < e1;
        c[u] = ok ? __ldg(a.col + e + u) : -1;
        w[u] = ok ? (has_val ? __ldg(a.val + e + u) : 1.f) : 0.f;
        if (ok) resolve_edge<GEN>(a, c[u], w[u]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const int ch = gl + G * j;
          xv[u][j] = (c[u] >= 0 && ch < V) ? ldg_f4(x4 + int64_t(c[u]) * ldx4 + ch) : f4_zero();
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < CH; ++j) f4_fma(acc[j], w[u], xv[u][j]);
    }
    finish_row<G, CH>(a, r, acc, gl, gmask);
  }
}

// Rows [ra, rb) of the local y (finished and visible to this CTA) -> peers' replicas / multicast address.
// peer_need (halo exchange) selects per row which peers receive it.  The masks are staged in shared memory
// first (one coalesced load), then every thread moves two independent 16-byte chunks per iteration so the
// L2 re-reads overlap; rows nobody needs cost one shared-memory read.
template <bool PEER>
__device__ __forceinline__ void push_rows(const GatherArgs& a, int64_t ra, int64_t rb, int tid, int nthreads,
                                          uint32_t* s_need) {
  if (!PEER || a.y == nullptr || rb <= ra) return;
  const int V = a.d >> 2;
  const uint32_t all = (1u << a.n_peer) - 1u;
  const int nrow = int(rb - ra);
  const bool mc_all = a.y_mc != nullptr && a.mc_min <= 0;       // pure multicast: every row, once
  const bool mc_some = a.y_mc != nullptr && a.mc_min > 0;       // hybrid: rows that >= mc_min peers need
  for (int j = tid; j < nrow; j += nthreads)
    s_need[j] = mc_all ? 1u : ((a.peer_need ? __ldg(a.peer_need + ra + j) : 0xffffffffu) & all);
  __syncthreads();
  const int total = nrow * V;
  auto send = [&](int64_t off, uint32_t need, const float4& v) {
    if (mc_all || (mc_some && __popc(need) >= a.mc_min)) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.y_mc + off), "f"(v.x), "f"(v.y),
                   "f"(v.z), "f"(v.w)
                   : "memory");
    } else {
      for (int p = 0; p < a.n_peer; ++p)
        if ((need >> p) & 1u) stg_peer_f4(reinterpret_cast<float4*>(a.y_peer[p] + off), v);
    }
  };
  for (int i0 = tid; i0 < total; i0 += 2 * nthreads) {
    const int i1 = i0 + nthreads;
    const int j0 = i0 / V, j1 = i1 / V;
    const uint32_t n0 = s_need[j0];
    const uint32_t n1 = (i1 < total) ? s_need[j1] : 0u;
    const int64_t off0 = (ra + j0) * a.ldy + int64_t(i0 - j0 * V) * 4;
    const int64_t off1 = (ra + j1) * a.ldy + int64_t(i1 - j1 * V) * 4;
    float4 v0 = f4_zero(), v1 = f4_zero();
    // L2 (coherent) loads: the rows were written by this CTA a moment ago, not through the read-only path
    if (n0) v0 = __ldcg(reinterpret_cast<const float4*>(a.y + off0));
    if (n1) v1 = __ldcg(reinterpret_cast<const float4*>(a.y + off1));
    if (n0) send(off0, n0, v0);
    if (n1) send(off1, n1, v1);
  }
}

// EXPERIMENTAL (-DGGAD_PUSH_PER_GROUP=1, off in the shipped build): every lane group pushes the rows it finished
// on its own right after its gather loop -- no CTA-wide barrier, the tail of one group overlaps the gathers of the
// others.  A lane re-reads exactly the chunks it stored itself (same thread, same address: program order), two rows
// in flight.  Rows combined across groups are pushed from registers by group 0 (two epilogue copies only).
#ifndef GGAD_PUSH_PER_GROUP
#define GGAD_PUSH_PER_GROUP 0
#endif
// A/B knob: 1 = the light epilogue (EPI 1) also uses the rolled boundary walk of the full one (3 inlined copies,
// ~30 KB of code) instead of the plain kernel's unrolled walk (17 copies, ~49 KB)
#ifndef GGAD_LIGHT_ROLLED
#define GGAD_LIGHT_ROLLED 0
#endif
// Hybrid staging (A/B knob, round 2): besides the U neighbor rows a lane group keeps in flight in REGISTERS, GGAD_SMEM_ROWS
// more rows per batch are fetched with cp.async (LDGSTS, no destination registers) into a per-thread slot of shared
// memory and consumed after the register part -- more bytes in flight per SM than the 64-register budget allows
// (the kernel is latency-bound at 46 % occupancy, profiles/r02a_ncu_*).  Same summation order, bit-identical results.
#ifndef GGAD_SMEM_ROWS
#define GGAD_SMEM_ROWS 0
#endif
template <int G, int CH>
__device__ __forceinline__ void push_group_rows(const GatherArgs& a, int64_t ra, int64_t rb, int gl) {
  if (a.y == nullptr) return;
  const int V = a.d >> 2;
  const uint32_t all = (1u << a.n_peer) - 1u;
  const bool mc = a.y_mc != nullptr;
  auto need_of = [&](int64_t r) -> uint32_t {
    return mc ? 1u : ((a.peer_need ? __ldg(a.peer_need + r) : 0xffffffffu) & all);
  };
  auto send = [&](int64_t off, uint32_t need, const float4& v) {
    if (mc) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.y_mc + off), "f"(v.x), "f"(v.y),
                   "f"(v.z), "f"(v.w)
                   : "memory");
    } else {
      for (int p = 0; p < a.n_peer; ++p)
        if ((need >> p) & 1u) stg_peer_f4(reinterpret_cast<float4*>(a.y_peer[p] + off), v);
    }
  };
  for (int64_t r = ra; r < rb; r += 2) {
    const uint32_t n0 = need_of(r);
    const uint32_t n1 = (r + 1 < rb) ? need_of(r + 1) : 0u;
    float4 v0[CH], v1[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int ch = gl + G * j;
      v0[j] = (n0 && ch < V) ? __ldcg(reinterpret_cast<const float4*>(a.y + r * a.ldy) + ch) : f4_zero();
      v1[j] = (n1 && ch < V) ? __ldcg(reinterpret_cast<const float4*>(a.y + (r + 1) * a.ldy) + ch) : f4_zero();
    }
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int ch = gl + G * j;
      if (n0 && ch < V) send(r * a.ldy + int64_t(ch) * 4, n0, v0[j]);
      if (n1 && ch < V) send((r + 1) * a.ldy + int64_t(ch) * 4, n1, v1[j]);
    }
  }
}

// ---------------------------------------------------------------------------
// merge-path tiled kernel
// ---------------------------------------------------------------------------
template <int G, int CH, bool HASVAL = true, bool BIAS = false>
struct TileSmem {
  static constexpr int NGRP = kThreads / G;
  static constexpr int kBar = 0;                                   // uint64 mbarrier (+pad)
  static constexpr int kCol = 16;                                  // int32[kTile + 8]
  static constexpr int kVal = kCol + (kTile + 8) * 4;              // float[kTile + 8] (absent when unweighted)
  static constexpr int kRend = kVal + (HASVAL ? (kTile + 8) * 4 : 0);  // int32[kTile + 8]
  static constexpr int kCi = kRend + (kTile + 8) * 4;              // int32[NGRP + 1] (+pad)
  static constexpr int kCj = kCi + ((NGRP + 1 + 3) / 4) * 16;      // int32[NGRP + 1] (+pad)
  static constexpr int kFlag = kCj + ((NGRP + 1 + 3) / 4) * 16;    // int32[NGRP] (+pad)
  static constexpr int kPart = kFlag + ((NGRP + 3) / 4) * 16;      // float4[NGRP][2][G*CH]
  static constexpr int kBias = kPart + NGRP * 2 * G * CH * 16;     // float4[G*CH] bias row (light epilogue only)
  static constexpr int kRows = kBias + (BIAS ? G * CH * 16 : 0);   // float4[GGAD_SMEM_ROWS][kThreads] cp.async slots
  static constexpr int kBytes = kRows + ((CH == 1) ? GGAD_SMEM_ROWS * kThreads * 16 : 0);
};

// MODE 0: unweighted (val == NULL), 1: per-edge val, 2: general (col_scale / xmap, optional val).
// EPI false: plain y = row_scale * acc; true: bias / activation / z / sumsq / dot epilogue.
// PEER: the epilogue also stores every finished row to the peers' replicas / the multicast address.
template <int G, int CH, int MODE, int EPI, bool PEER>
__global__ void __launch_bounds__(kThreads, (CH == 1) ? 4 : ((CH == 2) ? 3 : 2)) gather_tiled_kernel(const __grid_constant__ GatherArgs a) {
  using L = TileSmem<G, CH, MODE != 0, EPI == 1>;
  constexpr int NGRP = L::NGRP;
  constexpr int U = (CH == 1) ? 8 : (CH == 2 ? 4 : 2);
  extern __shared__ __align__(16) unsigned char smem[];
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + L::kBar);
  int32_t* s_col = reinterpret_cast<int32_t*>(smem + L::kCol);
  float* s_val = reinterpret_cast<float*>(smem + L::kVal);
  int32_t* s_rend = reinterpret_cast<int32_t*>(smem + L::kRend);
  int32_t* s_ci = reinterpret_cast<int32_t*>(smem + L::kCi);
  int32_t* s_cj = reinterpret_cast<int32_t*>(smem + L::kCj);
  int32_t* s_flag = reinterpret_cast<int32_t*>(smem + L::kFlag);
  float4* s_part = reinterpret_cast<float4*>(smem + L::kPart);

  const int tid = threadIdx.x;
  const int64_t k = blockIdx.x;
  const int64_t r0 = __ldg(a.tile_row + k), r1 = __ldg(a.tile_row + k + 1);
  const int64_t e0 = __ldg(a.tile_edge + k), e1 = __ldg(a.tile_edge + k + 1);
  const int nr = int(r1 - r0);  // rows whose end falls inside the tile
  const int ne = int(e1 - e0);  // edges inside the tile
  const bool has_val = (MODE != 0) && a.val != nullptr;

  // ---- stage the CSR slice: TMA bulk copy for col/val, plain loads for the row ends ----
  const int64_t e0a = e0 & ~int64_t(3);  // 16-byte aligned start
  const int lead = int(e0 - e0a);
  const int cnt = ne + lead;
  int nb = (cnt + 3) & ~3;
  if (e0a + nb > a.nnz) nb = cnt & ~3;  // never read past the arrays; the ragged tail is loaded below
  if (tid == 0) {
    mbar_init(s_bar, 1);
    fence_mbar_init();
  }
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
  // start of row r0 relative to the tile (<= 0; < 0 means the row began in an earlier tile)
  const int rstart0 = (r0 < a.n_rows) ? int(__ldg(a.rowptr + r0) - e0) : 0;
  EpiRegs ep{nullptr, 0.f, 0};
  if constexpr (EPI == 1) {
    float* s_bias = reinterpret_cast<float*>(smem + L::kBias);
    for (int t = tid; t < G * CH * 4; t += kThreads) s_bias[t] = (a.bias && t < a.d) ? __ldg(a.bias + t) : 0.f;
    ep.s_bias = s_bias;
    ep.act = a.prelu_slope ? 2 : (a.relu ? 1 : 0);
    ep.slope = a.prelu_slope ? __ldg(a.prelu_slope) : 0.f;
  }
  __syncthreads();

  // ---- second-level merge path: split (nr + ne) items evenly over the NGRP groups ----
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
  const int V = a.d >> 2;
  float4* my_part = s_part + (g * 2) * (G * CH);

  {
    const int i1 = s_ci[g], j1 = s_cj[g], i2 = s_ci[g + 1], j2 = s_cj[g + 1];
    int row = i1;
    int cur_end = s_rend[row];
    bool head_pending = ((i1 == 0) ? rstart0 : s_rend[i1 - 1]) < j1;  // first row began before this group
    const bool head0 = head_pending;
    (void)head0;
    int flag = 0;
    float4 acc[CH];
    // Byte offset of this lane's 16-byte chunk(s) inside a row.  Lanes beyond the row width (V not a
    // multiple of G) re-read chunk 0: their sums are never stored, and no predicate is needed in the loop.
    const char* lane_base[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      acc[j] = f4_zero();
      lane_base[j] = reinterpret_cast<const char*>(a.x) + ((gl + G * j < V) ? uint32_t(gl + G * j) * 16u : 0u);
      asm volatile("" : "+l"(lane_base[j]));  // keep it one opaque 64-bit register pair (IMAD.WIDE addend)
    }
    const uint32_t row_bytes = uint32_t(a.ldx) * 4u;
    // one IMAD.WIDE.U32 per gathered chunk: (x + lane offset) + col * row_bytes
    auto load_chunk = [&](int c, int j) {
      return ldg_f4(reinterpret_cast<const float4*>(lane_base[j] + uint64_t(uint32_t(c)) * row_bytes));
    };
    const int32_t* __restrict__ sc = s_col + lead;
    const float* __restrict__ sv = s_val + lead;

    auto flush = [&]() {
      if (head_pending) {
#pragma unroll
        for (int j = 0; j < CH; ++j) my_part[gl + G * j] = acc[j];
        flag |= 1;
        head_pending = false;
      } else {
        finish_row<G, CH, EPI, false>(a, r0 + row, acc, gl, gmask, ep);
      }
#pragma unroll
      for (int j = 0; j < CH; ++j) acc[j] = f4_zero();
      ++row;
      cur_end = s_rend[row];
    };
    auto accumulate = [&](float w, const float4 (&v)[CH]) {
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        if (MODE == 0) f4_add(acc[j], v[j]);
        else f4_fma(acc[j], w, v[j]);
      }
    };

    if constexpr (EPI == 2 || (EPI == 1 && GGAD_LIGHT_ROLLED)) {
      // Large epilogue: a batch that crosses row ends (and the ragged last batch) walks its rows in a rolled
      // loop, so only three inlined copies of the epilogue exist in the kernel.
      auto walk_rows = [&](int e, int nb, const float (&w)[U], const float4 (&xv)[U][CH]) {
        int u0 = 0;
        while (u0 < nb) {
          if (e + u0 >= cur_end) {
            flush();
            continue;
          }
          const int lim = (cur_end - e < nb) ? (cur_end - e) : nb;
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (u >= u0 && u < lim) accumulate(w[u], xv[u]);
          u0 = lim;
        }
      };
      int e = j1;
      for (; e + U <= j2; e += U) {
        int c[U];
        float w[U];
        float4 xv[U][CH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          c[u] = sc[e + u];
          w[u] = (MODE == 0) ? 1.f : ((MODE == 1 || has_val) ? sv[e + u] : 1.f);
          if (MODE == 2) resolve_edge<true>(a, c[u], w[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < CH; ++j) xv[u][j] = (MODE < 2 || c[u] >= 0) ? load_chunk(c[u], j) : f4_zero();
        if (e + U <= cur_end) {
#pragma unroll
          for (int u = 0; u < U; ++u) accumulate(w[u], xv[u]);
        } else {
          walk_rows(e, U, w, xv);
        }
      }
      if (e < j2) {
        const int nb = j2 - e;
        int c[U];
        float w[U];
        float4 xv[U][CH];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool ok = u < nb;
          c[u] = ok ? sc[e + u] : -1;
          w[u] = ok ? ((MODE == 0) ? 1.f : ((MODE == 1 || has_val) ? sv[e + u] : 1.f)) : 0.f;
          if (MODE == 2 && ok) resolve_edge<true>(a, c[u], w[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < CH; ++j) xv[u][j] = (c[u] >= 0) ? load_chunk(c[u], j) : f4_zero();
        walk_rows(e, nb, w, xv);
      }
    } else {
      int e = j1;
#if GGAD_SMEM_ROWS > 0
      if constexpr (CH == 1 && MODE < 2) {
        constexpr int US = GGAD_SMEM_ROWS, UB = U + US;
        float4* my_rows = reinterpret_cast<float4*>(smem + L::kRows) + tid;     // slot s at my_rows[s * kThreads]
        const uint32_t my_rows_s = smem_u32(my_rows);
        for (; e + UB <= j2; e += UB) {
          int c[U];
          float w[U], ws_[US];
          float4 xv[U][CH];
          // shared-memory part first, so that all UB rows are in flight together
#pragma unroll
          for (int q = 0; q < US; ++q) {
            const int cq = sc[e + U + q];
            ws_[q] = (MODE == 0) ? 1.f : sv[e + U + q];
            const char* src = lane_base[0] + uint64_t(uint32_t(cq)) * row_bytes;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(my_rows_s + uint32_t(q) * (kThreads * 16u)), "l"(src) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll
          for (int u = 0; u < U; ++u) {
            c[u] = sc[e + u];
            w[u] = (MODE == 0) ? 1.f : sv[e + u];
          }
#pragma unroll
          for (int u = 0; u < U; ++u) xv[u][0] = load_chunk(c[u], 0);
          const bool inside = e + UB <= cur_end;
          if (inside) {
#pragma unroll
            for (int u = 0; u < U; ++u) accumulate(w[u], xv[u]);
          } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
              while (e + u >= cur_end) flush();
              accumulate(w[u], xv[u]);
            }
          }
          asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
          for (int q = 0; q < US; ++q) {
            float4 xs[CH];
            xs[0] = my_rows[q * kThreads];
            if (!inside) {
              while (e + U + q >= cur_end) flush();
            }
            accumulate(ws_[q], xs);
          }
        }
      }
#endif
      // full batches of U edges: U independent 128-bit gathers per lane in flight
      for (; e + U <= j2; e += U) {
        int c[U];
        float w[U];
        float4 xv[U][CH];
  #pragma unroll
        for (int u = 0; u < U; ++u) {
          c[u] = sc[e + u];
          w[u] = (MODE == 0) ? 1.f : ((MODE == 1 || has_val) ? sv[e + u] : 1.f);
          if (MODE == 2) resolve_edge<true>(a, c[u], w[u]);
        }
  #pragma unroll
        for (int u = 0; u < U; ++u)
  #pragma unroll
          for (int j = 0; j < CH; ++j) xv[u][j] = (MODE < 2 || c[u] >= 0) ? load_chunk(c[u], j) : f4_zero();
        if (e + U <= cur_end) {  // whole batch inside the current row: no boundary checks
  #pragma unroll
          for (int u = 0; u < U; ++u) accumulate(w[u], xv[u]);
        } else {
  #pragma unroll
          for (int u = 0; u < U; ++u) {
            while (e + u >= cur_end) flush();
            accumulate(w[u], xv[u]);
          }
        }
      }
      if (e < j2) {  // ragged tail (< U edges)
        int c[U];
        float w[U];
        float4 xv[U][CH];
  #pragma unroll
        for (int u = 0; u < U; ++u) {
          const bool ok = e + u < j2;
          c[u] = ok ? sc[e + u] : -1;
          w[u] = ok ? ((MODE == 0) ? 1.f : ((MODE == 1 || has_val) ? sv[e + u] : 1.f)) : 0.f;
          if (MODE == 2 && ok) resolve_edge<true>(a, c[u], w[u]);
        }
  #pragma unroll
        for (int u = 0; u < U; ++u)
  #pragma unroll
          for (int j = 0; j < CH; ++j) xv[u][j] = (c[u] >= 0) ? load_chunk(c[u], j) : f4_zero();
  #pragma unroll
        for (int u = 0; u < U; ++u) {
          if (e + u < j2) {
            while (e + u >= cur_end) flush();
            accumulate(w[u], xv[u]);
          }
        }
      }
    }
    while (row < i2) flush();
    // edges of row i2 consumed by this group without reaching its end -> tail partial
    const int rs_tail = (i2 == 0) ? rstart0 : s_rend[i2 - 1];
    if (j2 > (rs_tail > j1 ? rs_tail : j1)) {
#pragma unroll
      for (int j = 0; j < CH; ++j) my_part[G * CH + gl + G * j] = acc[j];
      flag |= 2;
    }
    if (gl == 0) s_flag[g] = flag;
#if GGAD_PUSH_PER_GROUP
    if constexpr (PEER) push_group_rows<G, CH>(a, r0 + i1 + (head0 ? 1 : 0), r0 + i2, gl);
#endif
  }
  __syncthreads();

  // ---- combine rows cut by group boundaries, in group order (group 0 does it) ----
  if (g == 0) {
    float4 chain[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) chain[j] = f4_zero();
    float* ws_head = a.ws + (2 * k) * int64_t(a.d);
    float* ws_tail = ws_head + a.d;
    for (int q = 0; q < NGRP; ++q) {
      const int f = s_flag[q];
      const float4* part = s_part + (q * 2) * (G * CH);
      if (f & 1) {
#pragma unroll
        for (int j = 0; j < CH; ++j) f4_add(chain[j], part[gl + G * j]);
        const int row = s_ci[q];
        if (row == 0 && rstart0 < 0) {  // began in an earlier tile: tile_fixup_kernel finishes it
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (gl + G * j < V) reinterpret_cast<float4*>(ws_head)[gl + G * j] = chain[j];
        } else {
          finish_row<G, CH, EPI, (GGAD_PUSH_PER_GROUP != 0) && PEER>(a, r0 + row, chain, gl, gmask, ep);
        }
#pragma unroll
        for (int j = 0; j < CH; ++j) chain[j] = f4_zero();
      }
      if (f & 2) {
#pragma unroll
        for (int j = 0; j < CH; ++j) f4_add(chain[j], part[G * CH + gl + G * j]);
      }
    }
    // whatever is left belongs to row r1, which ends in a later tile (zeros if nothing)
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (gl + G * j < V) reinterpret_cast<float4*>(ws_tail)[gl + G * j] = chain[j];
  }

  // ---- fused exchange: push the rows this tile finished to the peers that gather them next ----
  // Kept out of the gather loop on purpose: peer stores compiled into the (17x inlined) row epilogue cost
  // 25 % of the whole launch through code size even when no row is sent.  The rows were just written to
  // the local y by this CTA, so the re-read hits L2; chunks are contiguous per row -> 256 B NVLink writes.
#if !GGAD_PUSH_PER_GROUP
  if constexpr (PEER) {
    __syncthreads();
    if (a.tile_done) {
      // chase mode: the exchange runs in ggad_halo_chase on a few SMs of its own; this CTA only publishes "rows
      // [r0 (+1), r1) are final" -- a release store ordered after every thread's y stores by the barrier above
      if (tid == 0) st_release_gpu(a.tile_done + k, a.tile_epoch);
    } else {
      push_rows<PEER>(a, r0 + ((rstart0 < 0) ? 1 : 0), r1, tid, kThreads, reinterpret_cast<uint32_t*>(s_rend));
    }
  }
#endif
}

// Finish rows that were cut by tile boundaries: one lane group per tile whose first row began earlier.
template <int G, int CH, bool EPI, bool PEER>
__global__ void __launch_bounds__(kThreads) tile_fixup_kernel(const __grid_constant__ GatherArgs a) {  // EPI: any epilogue (full code)
  const int64_t k = (int64_t(blockIdx.x) * kThreads + threadIdx.x) / G;
  if (k >= a.n_tiles) return;
  const int gl = threadIdx.x % G;
  const unsigned gmask = group_mask<G>();
  const int V = a.d >> 2;
  const int64_t r0 = __ldg(a.tile_row + k), r1 = __ldg(a.tile_row + k + 1);
  if (r1 == r0 || r0 >= a.n_rows) return;                       // no row ends in this tile
  if (__ldg(a.rowptr + r0) >= __ldg(a.tile_edge + k)) return;   // row r0 starts inside this tile
  int64_t ks = k - 1;                                           // tiles ks..k-1 hold tail partials of row r0
  while (ks > 0 && __ldg(a.tile_row + ks) == r0) --ks;
  float4 acc[CH];
#pragma unroll
  for (int j = 0; j < CH; ++j) acc[j] = f4_zero();
  for (int64_t q = ks; q < k; ++q) {
    const float4* t = reinterpret_cast<const float4*>(a.ws + (2 * q + 1) * int64_t(a.d));
#pragma unroll
    for (int j = 0; j < CH; ++j)
      if (gl + G * j < V) f4_add(acc[j], t[gl + G * j]);
  }
  const float4* h = reinterpret_cast<const float4*>(a.ws + (2 * k) * int64_t(a.d));
#pragma unroll
  for (int j = 0; j < CH; ++j)
    if (gl + G * j < V) f4_add(acc[j], h[gl + G * j]);
  finish_row<G, CH, EPI ? 2 : 0, PEER>(a, r0, acc, gl, gmask);
}

// ---------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------
template <int G, int CH, int MODE, int EPI, bool PEER>
static int launch_tiled(const GatherArgs& a, cudaStream_t st) {
  using L = TileSmem<G, CH, MODE != 0, EPI == 1>;
  // the attribute is per device (context), so the "already set" flag is too -- per instantiation and device
  static bool attr_done[64] = {};
  int dev = 0;
  GGAD_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 64 || !attr_done[dev]) {
    GGAD_CUDA_OK(cudaFuncSetAttribute(gather_tiled_kernel<G, CH, MODE, EPI, PEER>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, L::kBytes));
    if (dev < 64) attr_done[dev] = true;
  }
  gather_tiled_kernel<G, CH, MODE, EPI, PEER><<<(unsigned)a.n_tiles, kThreads, L::kBytes, st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  const int64_t fix_blocks = (a.n_tiles * G + kThreads - 1) / kThreads;
  tile_fixup_kernel<G, CH, EPI != 0, PEER><<<(unsigned)fix_blocks, kThreads, 0, st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return GGAD_OK;
}

template <int G, int CH, int EPI, bool PEER>
static int launch_mode(const GatherArgs& a, cudaStream_t st) {
  if (a.xmap || a.col_scale) return launch_tiled<G, CH, 2, EPI, PEER>(a, st);
  if (a.val) return launch_tiled<G, CH, 1, EPI, PEER>(a, st);
  return launch_tiled<G, CH, 0, EPI, PEER>(a, st);
}

template <int G, int CH>
int launch_variant(const GatherArgs& a, cudaStream_t st, int sm_count) {
  if (a.n_rows == 0) return GGAD_OK;
  const bool gen = a.xmap || a.col_scale;
  const bool peer = a.n_peer > 0 || a.y_mc || a.tile_done;
  if (a.tile_row) {
    // epilogue kind: 2 = reductions (|y|^2, dot) or no y at all; 1 = elementwise only (bias / activation / z); 0 = plain
    static const bool force_full = getenv("GGAD_FORCE_FULL_EPI") != nullptr;   // A/B knob (profiling only)
    const int epi = (a.sumsq || a.dot_out || (!a.y && !a.y_mc) || (force_full && (a.bias || a.prelu_slope || a.relu || a.z)))
                        ? 2 : ((a.bias || a.prelu_slope || a.relu || a.z) ? 1 : 0);
    if (peer) {
      if (epi == 2) return launch_mode<G, CH, 2, true>(a, st);
      return epi == 1 ? launch_mode<G, CH, 1, true>(a, st) : launch_mode<G, CH, 0, true>(a, st);
    }
    if (epi == 2) return launch_mode<G, CH, 2, false>(a, st);
    if constexpr (G == 16 && CH == 1) {
      if (epi == 0) {
        const int r = try_launch_tma_rows(a, st);
        if (r >= 0) return r;
      }
    }
    return epi == 1 ? launch_mode<G, CH, 1, false>(a, st) : launch_mode<G, CH, 0, false>(a, st);
  }
  GGAD_REQUIRE(!peer, GGAD_ERR_UNSUPPORTED, "gather_reduce: peer / multicast stores need the merge-path plan");
  const int64_t gpb = kThreads / G;
  int64_t blocks = (a.n_rows + gpb - 1) / gpb;
  const int64_t cap = int64_t(sm_count) * 64;
  if (blocks > cap) blocks = cap;
  if (gen) gather_rows_kernel<G, CH, true><<<(unsigned)blocks, kThreads, 0, st>>>(a);
  else gather_rows_kernel<G, CH, false><<<(unsigned)blocks, kThreads, 0, st>>>(a);
  GGAD_CUDA_OK(cudaGetLastError());
  count_launch(1);
  return GGAD_OK;
}

}  // namespace ggad
