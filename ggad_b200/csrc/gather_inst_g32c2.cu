// Explicit instantiation of the gather-reduce kernels for lane groups of 32 lanes x 2 chunk(s)
// (one translation unit per width class so the variants compile in parallel).
#include "gather_kernels.cuh"

namespace ggad {
int launch_g32c2(const GatherArgs& a, cudaStream_t st, int sm_count) { return launch_variant<32, 2>(a, st, sm_count); }
}  // namespace ggad
