// Actor MLP inference behind the C ABI (include/taco_b200.h, taco_actor_*).
// Placeholder until the CUDA-core / tcgen05 kernels land: every entry point reports TACO_E_INVALID.
#include <string>
#include "../../include/taco_b200.h"

extern "C" {
int taco_actor_create(int, const int32_t*, int32_t, TacoActor** out) { if (out) *out = nullptr; return TACO_E_INVALID; }
int taco_actor_destroy(TacoActor*) { return TACO_OK; }
int taco_actor_load(TacoActor*, const float* const*, const float* const*, float, void*) { return TACO_E_INVALID; }
int taco_actor_forward(TacoActor*, const float*, float*, int32_t, int32_t, void*) { return TACO_E_INVALID; }
}
