// Shared helpers for the ggad_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ggad_b200.h"

namespace ggad {

// thread-local error string + global launch counter (api.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define GGAD_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::ggad::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return GGAD_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

#define GGAD_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      ::ggad::set_error(__VA_ARGS__);  \
      return (code);                   \
    }                                  \
  } while (0)

// Stream-ordered temporary from the library's own memory pool (api.cu).  The pool keeps what it has allocated
// across synchronisations (release threshold = max), so a per-batch caller (the mini-batch frontier) does not pay
// the driver for physical memory on every call; ggad_trim_workspace() gives it back.  Free with cudaFreeAsync.
cudaError_t temp_alloc(void** p, size_t bytes, cudaStream_t st);

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D TMA bulk copy (global -> shared), cache-hinted ld/st
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  const uint32_t addr = smem_u32(bar);
  do {
    if (++spins > (1u << 26)) asm volatile("trap;");  // a lost TMA completion must not hang the GPU
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared; completion is signalled on `bar` (complete_tx::bytes).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }
// streaming (evict-first) 128-bit store for outputs that are not re-read by this kernel
__device__ __forceinline__ void stg_cs_f4(float4* p, const float4& v) { __stcs(p, v); }

// 128-bit store into a peer GPU's memory over NVLink (A/B knob: GGAD_PEER_ST = 0 plain, 1 streaming, 2 write-through)
#ifndef GGAD_PEER_ST
#define GGAD_PEER_ST 1
#endif
__device__ __forceinline__ void stg_peer_f4(float4* p, const float4& v) {
#if GGAD_PEER_ST == 0
  *p = v;
#elif GGAD_PEER_ST == 2
  __stwt(p, v);
#else
  __stcs(p, v);
#endif
}

// release / acquire flag accesses at GPU scope (tile-done flags between the gather kernel and ggad_halo_chase)
__device__ __forceinline__ void st_release_gpu(int32_t* p, int32_t v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_gpu(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4& a, float w, const float4& x) {
  a.x = fmaf(w, x.x, a.x);
  a.y = fmaf(w, x.y, a.y);
  a.z = fmaf(w, x.z, a.z);
  a.w = fmaf(w, x.w, a.w);
}
__device__ __forceinline__ void f4_add(float4& a, const float4& x) {
  a.x += x.x;
  a.y += x.y;
  a.z += x.z;
  a.w += x.w;
}

}  // namespace ggad
