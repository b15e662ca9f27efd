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

// One environment's reverse walk (the body of the NumPy loop, rollout_buffer.py:441-481).
__device__ __forceinline__ void gae_scan_env(const float *__restrict__ rewards, const uint8_t *__restrict__ dones,
                                             const float *__restrict__ values, const double *__restrict__ last_value,
                                             const float *__restrict__ last_done, int64_t T, int64_t E, double gamma,
                                             double gamma_lambda, int use_gae, float *__restrict__ advantages,
                                             float *__restrict__ returns, int64_t e) {
    const float gamma_f = (float)gamma;
    if (use_gae) {
        double last = 0.0;                                            // last_gae_lambda
        for (int64_t t = T - 1; t >= 0; --t) {
            double nnt, gv;
            if (t == T - 1) {
                nnt = __dsub_rn(1.0, (double)last_done[e]);
                gv = __dmul_rn(gamma, last_value[e]);                 // float64 bootstrap value (last_value.astype(float))
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
        double last = __dmul_rn(last_value[e], __dsub_rn(1.0, (double)last_done[e]));   // last_returns
        for (int64_t t = T - 1; t >= 0; --t) {
            const double keep = __dsub_rn(1.0, dones[t * E + e] ? 1.0 : 0.0);
            last = __dadd_rn((double)rewards[t * E + e], __dmul_rn(__dmul_rn(gamma, last), keep));
            const float r = (float)last;
            returns[t * E + e] = r;
            advantages[t * E + e] = __fsub_rn(r, values[t * E + e]);
        }
    }
}

__global__ void gae_scan_kernel(const float *__restrict__ rewards, const uint8_t *__restrict__ dones,
                                const float *__restrict__ values, const double *__restrict__ last_value,
                                const float *__restrict__ last_done, int64_t T, int64_t E, double gamma,
                                double gamma_lambda, int use_gae, float *__restrict__ advantages,
                                float *__restrict__ returns) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    gae_scan_env(rewards, dones, values, last_value, last_done, T, E, gamma, gamma_lambda, use_gae, advantages, returns, e);
}

// Fixed-order fp64 block sum: thread-strided partials, then a shared-memory tree (deterministic run to run).
__device__ __forceinline__ double block_sum_f64(double v, double *red) {
    red[threadIdx.x] = v;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] = __dadd_rn(red[threadIdx.x], red[threadIdx.x + s]);
        __syncthreads();
    }
    const double out = red[0];
    __syncthreads();
    return out;
}

// PPO's global advantage normalisation (ppo.py:831-834 / :935-944): (a - a.mean()) / (a.std() + 1e-8), std with
// Bessel's correction like torch.std.  Bit policy: the two reductions are carried in float64 in a fixed order
// (mean = f32(sum64 / n); var = sum64((a - mean64)^2) / (n - 1)), i.e. the correctly rounded statistics; torch's
// float32 cascade sums differ from them by a few ulp, so parity with the reference is stated as 1e-6 relative,
// not bit equality.  The elementwise part is float32 in the reference's operation order.
__device__ __forceinline__ void normalize_block(const float *__restrict__ a, int64_t n, float *__restrict__ out,
                                                double *red) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s = __dadd_rn(s, (double)a[i]);
    const double mean64 = __ddiv_rn(block_sum_f64(s, red), (double)n);
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double d = __dsub_rn((double)a[i], mean64);
        q = __dadd_rn(q, __dmul_rn(d, d));
    }
    const double var = __ddiv_rn(block_sum_f64(q, red), (double)(n > 1 ? n - 1 : 1));
    const float mean = (float)mean64;
    const float denom = __fadd_rn((float)__dsqrt_rn(var), 1e-8f);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) out[i] = __fdiv_rn(__fsub_rn(a[i], mean), denom);
}

__global__ void advantage_normalize_kernel(const float *__restrict__ a, int64_t n, float *__restrict__ out) {
    extern __shared__ double red[];
    normalize_block(a, n, out, red);
}

// The whole post-rollout step of one agent in ONE launch (BASELINE config 4: 256 environments): thread e walks
// environment e backwards, the CTA then normalises the T x E advantages it has just written.
__global__ void gae_fused_kernel(const float *__restrict__ rewards, const uint8_t *__restrict__ dones,
                                 const float *__restrict__ values, const double *__restrict__ last_value,
                                 const float *__restrict__ last_done, int64_t T, int64_t E, double gamma,
                                 double gamma_lambda, int use_gae, float *__restrict__ advantages,
                                 float *__restrict__ returns, float *__restrict__ adv_norm) {
    extern __shared__ double red[];
    if ((int64_t)threadIdx.x < E)
        gae_scan_env(rewards, dones, values, last_value, last_done, T, E, gamma, gamma_lambda, use_gae, advantages, returns,
                     threadIdx.x);
    __syncthreads();                      // the block's own global writes are visible to the block after the barrier
    normalize_block(advantages, T * E, adv_norm, red);
}

}  // namespace b2rl

using namespace b2rl;

extern "C" int b2rl_gae_scan(const float *rewards, const uint8_t *dones, const float *values, const double *last_value,
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

extern "C" int b2rl_advantage_normalize(const float *advantages, int64_t n, float *out, void *stream) {
    B2RL_CHECK_ARG(advantages && out && n >= 1, "bad arguments");
    advantage_normalize_kernel<<<1, 1024, 1024 * sizeof(double), static_cast<cudaStream_t>(stream)>>>(advantages, n, out);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

extern "C" int b2rl_gae_scan_normalize(const float *rewards, const uint8_t *dones, const float *values,
                                       const double *last_value, const float *last_done, int64_t T, int64_t E, double gamma,
                                       double gae_lambda, int use_gae, float *advantages, float *returns, float *adv_norm,
                                       void *stream) {
    B2RL_CHECK_ARG(rewards && dones && values && last_value && last_done && advantages && returns && adv_norm,
                   "NULL buffer");
    B2RL_CHECK_ARG(T >= 1 && E >= 1, "bad rollout shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (E <= 1024) {
        int threads = 128;
        while (threads < E) threads <<= 1;
        gae_fused_kernel<<<1, threads, threads * sizeof(double), s>>>(rewards, dones, values, last_value, last_done, T, E,
                                                                      gamma, gamma * gae_lambda, use_gae, advantages,
                                                                      returns, adv_norm);
        B2RL_LAUNCH_CHECK();
        return B2RL_OK;
    }
    int rc = b2rl_gae_scan(rewards, dones, values, last_value, last_done, T, E, gamma, gae_lambda, use_gae, advantages,
                           returns, stream);
    if (rc != B2RL_OK) return rc;
    return b2rl_advantage_normalize(advantages, T * E, adv_norm, stream);
}
