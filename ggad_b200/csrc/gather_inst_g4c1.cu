// Explicit instantiation of the gather-reduce kernels for lane groups of 4 lanes x 1 chunk(s)
// (one translation unit per width class so the variants compile in parallel).
#include "gather_kernels.cuh"

namespace ggad {
int launch_g4c1(const GatherArgs& a, cudaStream_t st, int sm_count) { return launch_variant<4, 1>(a, st, sm_count); }
}  // namespace ggad
