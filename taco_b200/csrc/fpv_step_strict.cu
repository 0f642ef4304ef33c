// Fused step kernel, strict build: this translation unit is compiled with -fmad=false so every
// float32 multiply and add rounds separately, like the reference's eager torch ops (TACO_F_STRICT_FP).
#define TACO_VARIANT strict
#include "fpv_step_kernel.cuh"
namespace taco {
void launch_fpv_step_strict(const StepParams& p, cudaStream_t stream) { strict::launch_any(p, stream); }
}  // namespace taco
