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

extern unsigned long long g_conv_path[3];  // forward convolutions per path (b2rl_conv_path_count)
extern unsigned long long g_launches;   // kernels launched by this library (bench.py's gpu_launches)
#define B2RL_LAUNCH_CHECK()                 \
    do {                                    \
        ++::b2rl::g_launches;               \
        B2RL_CUDA(cudaGetLastError());      \
    } while (0)

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

// Philox4x32-10: one 128-bit block per (counter, key).
__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t c_lo, uint64_t c_hi, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)c_lo, c1 = (uint32_t)(c_lo >> 32), c2 = (uint32_t)c_hi, c3 = (uint32_t)(c_hi >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}


// standard normal from one Philox block (Box-Muller on two 24-bit uniforms)
__device__ __forceinline__ float philox_normal(uint64_t seed, uint64_t ctr, uint64_t stream) {
    uint32_t r[4];
    philox4x32_10(seed, ctr, stream, r);
    const float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = (float)(r[1] >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

}  // namespace b2rl
