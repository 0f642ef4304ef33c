// Fused step kernel, default build: FMA contraction enabled.
#define TACO_VARIANT fast
#include "fpv_step_kernel.cuh"
namespace taco {
void launch_fpv_step_fast(const StepParams& p, cudaStream_t stream) { fast::launch_any(p, stream); }
}  // namespace taco
