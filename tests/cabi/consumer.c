/* A plain-C consumer of libtaco_b200.so: what a non-Python integrator of the reference would link (the role gymtorch.cpp plays
 * upstream, python/isaacgym/_bindings/src/gymtorch/gymtorch.cpp:33-158).  No CUDA headers, no torch: the host-buffer entry point
 * takes ordinary malloc'ed memory.  Steps a flip env with a fixed action pattern and prints a summary line that
 * tests/test_cabi.py compares with the Python binding's result for the same configuration.
 *
 *   gcc -std=c99 -I include tests/cabi/consumer.c -L taco_b200/lib -ltaco_b200 -Wl,-rpath,$PWD/taco_b200/lib -o consumer
 *   ./consumer [num_envs] [steps]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "taco_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != TACO_OK) {                                                        \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, taco_last_error());  \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1000;
    const int steps = argc > 2 ? atoi(argv[2]) : 12;
    if (taco_abi_version() != TACO_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 1; }
    TacoCfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.abi_version = TACO_ABI_VERSION;
    cfg.num_envs = n; cfg.env_offset = 0; cfg.num_envs_global = n;
    cfg.task_mode = TACO_TASK_FLIP; cfg.len_obs = 1; cfg.len_states = 5;
    cfg.max_episode_length = 1000; cfg.control_freq_inv = 10; cfg.substeps = 2; cfg.delay_time = 20;
    cfg.flags = TACO_F_RANDOM_COPTER_POS | TACO_F_RANDOM_COPTER_QUAT | TACO_F_RANDOM_COPTER_VEL | TACO_F_RANDOM_TARGET_POS |
                TACO_F_RANDOM_TARGET_YAW | TACO_F_BATTERY_CONSUMPTION | TACO_F_ROTOR_RESPONSE | TACO_F_RANDOM_COMMAND | TACO_F_STRICT_FP;
    cfg.dt = 0.001f; cfg.rotor_response_time = 0.017f; cfg.difficulty = 1.0f; cfg.clip_actions = 1.0f; cfg.seed = 5;
    TacoEnv* env = NULL;
    CHECK(taco_env_create(&cfg, 0, &env));
    float* act = (float*)malloc((size_t)n * 4 * sizeof(float));
    float* rew = (float*)malloc((size_t)n * sizeof(float));
    int64_t* reset = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    uint8_t* tout = (uint8_t*)malloc((size_t)n);
    double sum_rew = 0.0; long long n_reset = 0, n_tout = 0;
    for (int t = 0; t < steps; ++t) {
        for (int i = 0; i < n; ++i) {                 /* a deterministic action pattern any binding can reproduce exactly */
            act[4 * i + 0] = (float)((i * 7 + t * 3) % 17) / 16.0f - 0.5f;
            act[4 * i + 1] = (float)((i * 5 + t) % 13) / 24.0f - 0.25f;
            act[4 * i + 2] = (float)((i * 3 + t * 2) % 11) / 20.0f - 0.25f;
            act[4 * i + 3] = (float)((i + t * 5) % 7) / 12.0f - 0.25f;
        }
        CHECK(taco_env_step_host(env, act, rew, reset, tout, NULL));      /* pageable memory: the copy pipeline; NULL = default stream */
        for (int i = 0; i < n; ++i) { sum_rew += rew[i]; n_reset += reset[i] != 0; n_tout += tout[i] != 0; }
    }
    double stats[TACO_NUM_STATS];
    CHECK(taco_env_stats(env, NULL, stats, NULL));
    printf("consumer n=%d steps=%d sum_rew=%.9e n_reset=%lld n_tout=%lld stats_sum_rew=%.9e stats_n_done=%.0f stats_env_steps=%.0f\n", n, steps,
           sum_rew, n_reset, n_tout, stats[0], stats[1], stats[7]);
    /* error path: status code + message, no exception across the boundary */
    if (taco_env_step_host(env, NULL, rew, reset, tout, NULL) != TACO_E_INVALID || strlen(taco_last_error()) == 0) { fprintf(stderr, "error path broken\n"); return 1; }
    CHECK(taco_env_destroy(env));
    free(act); free(rew); free(reset); free(tout);
    return 0;
}
