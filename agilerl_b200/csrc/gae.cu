// gae.cu — PPO's return / advantage recurrence on device.
//
// Replaces RolloutBuffer.compute_returns_and_advantages (agilerl/components/rollout_buffer.py:413-481): a host
// NumPy loop `for t in reversed(range(T))` over [num_envs] rows, bracketed by a device->host copy of rewards /
// dones / values and a host->device copy of the results.  One thread per environment walks its T steps backwards;
// a launch is T*E*(3 reads + 2 writes) * 4 B of traffic and T dependent steps of latency.
//
// The reference's dtype walk is reproduced literally (oracle/gae.py; NumPy >= 2 promotion rules):
//   * rewards, values are float32, dones bool; `1.0 - dones.astype(float)` and `last_value.astype(float)` are
//     float64; `gamma * values[t+1]` is a float32 product (Python scalar x float32 array), widened afterwards;
//   * delta and the carried last_gae_lambda are float64, each advantages[t] is rounded to float32 when stored;
//   * returns = advantages(f32) + values(f32) in float32.  Monte-Carlo mode carries float64 likewise.
// No FMA contraction (explicit _rn intrinsics): results are bit-identical to the NumPy loop.
#include "common.cuh"

namespace b2rl {

__global__ void gae_scan_kernel(const float *__restrict__ rewards, const uint8_t *__restrict__ dones,
                                const float *__restrict__ values, const float *__restrict__ last_value,
                                const float *__restrict__ last_done, int64_t T, int64_t E, double gamma,
                                double gamma_lambda, int use_gae, float *__restrict__ advantages,
                                float *__restrict__ returns) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    const float gamma_f = (float)gamma;
    if (use_gae) {
        double last = 0.0;                                            // last_gae_lambda
        for (int64_t t = T - 1; t >= 0; --t) {
            double nnt, gv;
            if (t == T - 1) {
                nnt = __dsub_rn(1.0, (double)last_done[e]);
                gv = __dmul_rn(gamma, (double)last_value[e]);         // float64 bootstrap value
            } else {
                nnt = __dsub_rn(1.0, dones[(t + 1) * E + e] ? 1.0 : 0.0);
                gv = (double)__fmul_rn(gamma_f, values[(t + 1) * E + e]);   // float32 product, then widened
            }
            const double delta = __dsub_rn(__dadd_rn((double)rewards[t * E + e], __dmul_rn(gv, nnt)), (double)values[t * E + e]);
            last = __dadd_rn(delta, __dmul_rn(__dmul_rn(gamma_lambda, nnt), last));
            const float a = (float)last;
            advantages[t * E + e] = a;
            returns[t * E + e] = __fadd_rn(a, values[t * E + e]);
        }
    } else {
        double last = __dmul_rn((double)last_value[e], __dsub_rn(1.0, (double)last_done[e]));   // last_returns
        for (int64_t t = T - 1; t >= 0; --t) {
            const double keep = __dsub_rn(1.0, dones[t * E + e] ? 1.0 : 0.0);
            last = __dadd_rn((double)rewards[t * E + e], __dmul_rn(__dmul_rn(gamma, last), keep));
            const float r = (float)last;
            returns[t * E + e] = r;
            advantages[t * E + e] = __fsub_rn(r, values[t * E + e]);
        }
    }
}

}  // namespace b2rl

using namespace b2rl;

extern "C" int b2rl_gae_scan(const float *rewards, const uint8_t *dones, const float *values, const float *last_value,
                             const float *last_done, int64_t T, int64_t E, double gamma, double gae_lambda, int use_gae,
                             float *advantages, float *returns, void *stream) {
    B2RL_CHECK_ARG(rewards && dones && values && last_value && last_done && advantages && returns, "NULL buffer");
    B2RL_CHECK_ARG(T >= 0 && E >= 1, "bad rollout shape");
    if (T == 0) return B2RL_OK;
    const int threads = 128;
    gae_scan_kernel<<<(int)((E + threads - 1) / threads), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        rewards, dones, values, last_value, last_done, T, E, gamma, gamma * gae_lambda, use_gae, advantages, returns);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}
