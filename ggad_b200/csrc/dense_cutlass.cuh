// K6: the dense projections around the aggregation (X W^T of model.py:27 / fc1..fc4 of model.py:176-180 /
// W . agg^T of src/graphsage.py:412,419,430 and their backward GEMMs) on the 5th-generation tensor cores.
//
// fp32 in, fp32 out, fp32 accuracy: tcgen05 has no fp32 MMA kind and plain TF32 (10-bit mantissa) misses the 1e-4
// parity tolerance at K >= 64, so the operands are split on the fly into three bf16 terms each and the products are
// accumulated as bf16 x bf16 -> fp32 MMAs in tensor memory ("9xBF16", 5 significant bands): ~2^-22 relative error.
// The kernel is assembled from the CUTLASS 4.x sm100 building blocks (vendored header tree, header-only): TMA loads
// (UTMALDG) of fp32 tiles, a transform warp group that writes the bf16 terms to TMEM / shared memory (STTM),
// tcgen05.mma with TMEM accumulators (UTCHMMA), tcgen05.ld epilogue (LDTM) and a TMA store (UTMASTG).
// One translation unit per operand layout (dense_tn.cu, dense_nn.cu, dense_nt.cu) so they compile in parallel.
#pragma once
#include <cuda_runtime.h>

#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/dispatch_policy.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#include "cutlass/epilogue/fusion/operations.hpp"
#include "cutlass/epilogue/thread/activation.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/kernel/tile_scheduler.hpp"
#include "cutlass/kernel_hardware_info.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/util/packed_stride.hpp"

namespace ggad {

// C[M,N] = act(alpha * A[M,K] * B[K,N] + beta * C).  LayoutA / LayoutB are the CUTLASS tags of the GEMM operands:
// A RowMajor = K contiguous, B ColumnMajor = K contiguous (a row-major [N,K] weight), B RowMajor = N contiguous.
// STREAMK: the stream-K tile scheduler splits the k loop of the few output tiles of a weight-gradient GEMM
// (dW = dY^T X: K = number of nodes, M x N = 300 x 300) over all SMs, with a deterministic fix-up through a workspace.
template <class LayoutA, class LayoutB, bool RELU, bool STREAMK = false>
struct FastF32Gemm {
  using ArchTag = cutlass::arch::Sm100;
  using OpClass = cutlass::arch::OpClassTensorOp;
  using TileShape = cute::Shape<cute::_128, cute::_128, cute::_16>;
  using ClusterShape = cute::Shape<cute::_1, cute::_1, cute::_1>;
  static constexpr int kAlign = 4;  // floats: 16-byte rows for TMA
  using Fusion = std::conditional_t<RELU, cutlass::epilogue::fusion::LinCombEltAct<cutlass::epilogue::thread::ReLu, float, float>,
                                    cutlass::epilogue::fusion::LinearCombination<float, float>>;
  using CollectiveEpilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
      ArchTag, OpClass, TileShape, ClusterShape, cutlass::epilogue::collective::EpilogueTileAuto, float, float, float,
      cutlass::layout::RowMajor, kAlign, float, cutlass::layout::RowMajor, kAlign,
      cutlass::epilogue::collective::EpilogueScheduleAuto, Fusion>::CollectiveOp;
  using CollectiveMainloop = typename cutlass::gemm::collective::CollectiveBuilder<
      ArchTag, OpClass, float, LayoutA, kAlign, float, LayoutB, kAlign, float, TileShape, ClusterShape,
      cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename CollectiveEpilogue::SharedStorage))>,
      cutlass::gemm::KernelTmaWarpSpecialized1SmFastFP32Sm100>::CollectiveOp;
  using Scheduler = std::conditional_t<STREAMK, cutlass::gemm::StreamKScheduler, void>;
  using GemmKernel = cutlass::gemm::kernel::GemmUniversal<cute::Shape<int, int, int, int>, CollectiveMainloop, CollectiveEpilogue, Scheduler>;
  using Gemm = cutlass::gemm::device::GemmUniversalAdapter<GemmKernel>;
};

// returns 0, or a negative stage code (1 can_implement, 2 workspace, 3 initialize, 4 run) for the caller to report
template <class LayoutA, class LayoutB, bool RELU, bool STREAMK = false>
int run_fast_f32(int M, int N, int K, const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                 float alpha, float beta, void* ws, size_t ws_bytes, size_t* ws_needed, cudaStream_t st) {
  using G = typename FastF32Gemm<LayoutA, LayoutB, RELU, STREAMK>::Gemm;
  using StrideA = typename G::GemmKernel::StrideA;
  using StrideB = typename G::GemmKernel::StrideB;
  using StrideC = typename G::GemmKernel::StrideC;
  // packed strides first (sets the static unit mode), then the real leading dimensions
  StrideA sa = cutlass::make_cute_packed_stride(StrideA{}, cute::make_shape(M, K, 1));
  StrideB sb = cutlass::make_cute_packed_stride(StrideB{}, cute::make_shape(N, K, 1));
  StrideC sc = cutlass::make_cute_packed_stride(StrideC{}, cute::make_shape(M, N, 1));
  if constexpr (std::is_same_v<LayoutA, cutlass::layout::RowMajor>) cute::get<0>(sa) = lda; else cute::get<1>(sa) = lda;
  if constexpr (std::is_same_v<LayoutB, cutlass::layout::ColumnMajor>) cute::get<0>(sb) = ldb; else cute::get<1>(sb) = ldb;
  cute::get<0>(sc) = ldc;
  typename G::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm, {M, N, K, 1}, {A, sa, B, sb}, {{}, C, sc, C, sc}};
  args.epilogue.thread.alpha = alpha;
  args.epilogue.thread.beta = beta;
  int dev = 0;
  cudaGetDevice(&dev);
  args.hw_info.device_id = dev;
  args.hw_info.sm_count = cutlass::KernelHardwareInfo::query_device_multiprocessor_count(dev);
  G gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) return -1;
  const size_t need = G::get_workspace_size(args);
  if (ws_needed) *ws_needed = need;
  if (need > ws_bytes) return -2;
  if (gemm.initialize(args, ws, st) != cutlass::Status::kSuccess) return -3;
  if (gemm.run(st) != cutlass::Status::kSuccess) return -4;
  return 0;
}

}  // namespace ggad
