// Fused step kernel, default build: FMA contraction enabled.
#include "fpv_step_kernel.cuh"
namespace taco {
void launch_fpv_step_fast(const StepParams& p, cudaStream_t stream) { launch_any(p, stream); }
}  // namespace taco
