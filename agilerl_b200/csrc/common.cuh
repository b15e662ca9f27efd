// common.cuh — shared helpers for libb2rl (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b2rl.h"

namespace b2rl {

void set_error(const char *fmt, ...);

#define B2RL_CHECK_ARG(cond, ...)                       \
    do {                                                \
        if (!(cond)) {                                  \
            ::b2rl::set_error(__VA_ARGS__);             \
            return B2RL_EINVAL;                         \
        }                                               \
    } while (0)

#define B2RL_CUDA(expr)                                                                  \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            ::b2rl::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                              __FILE__, __LINE__);                                       \
            return B2RL_ECUDA;                                                           \
        }                                                                                \
    } while (0)

#define B2RL_LAUNCH_CHECK() B2RL_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// activations (torch semantics: ELU alpha=1, exact-erf GELU)
__device__ __forceinline__ float act_fwd(int act, float x) {
    switch (act) {
        case B2RL_ACT_RELU: return x > 0.f ? x : 0.f;
        case B2RL_ACT_ELU: return x > 0.f ? x : expm1f(x);
        case B2RL_ACT_GELU: return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
        case B2RL_ACT_TANH: return tanhf(x);
        default: return x;
    }
}
// derivative given pre-activation x and post-activation y
__device__ __forceinline__ float act_bwd(int act, float x, float y) {
    switch (act) {
        case B2RL_ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case B2RL_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
        case B2RL_ACT_GELU: {
            const float kA = 0.70710678118654752440f, kB = 0.39894228040143267794f;  // 1/sqrt(2), 1/sqrt(2pi)
            return 0.5f * (1.f + erff(x * kA)) + x * kB * expf(-0.5f * x * x);
        }
        case B2RL_ACT_TANH: return 1.f - y * y;
        default: return 1.f;
    }
}

}  // namespace b2rl
