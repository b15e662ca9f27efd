// nn.cu — RainbowQNetwork / QNetwork forward + backward and the learn() tail on sm_100a.
//
// Replaces, for the off-policy path:
//   EvolvableCNN / EvolvableMLP forward            agilerl/modules/cnn.py:552-580, mlp.py:188-203
//   NoisyLinear forward / reset_noise              agilerl/modules/custom_components.py:89-131
//   DuelingDistributionalMLP.forward               agilerl/networks/custom_modules.py:127-162
//   RainbowDQN._dqn_loss / learn / soft_update     agilerl/algorithms/dqn_rainbow.py:284-501
//   DQN.update / learn                             agilerl/algorithms/dqn.py:274-358
//   torch.autograd backward, clip_grad_norm_, torch.optim.Adam (all ATen in the reference)
//
// Layout in HBM: one flat fp32 parameter buffer per network (encoder | head mu | head sigma),
// one flat epsilon buffer, NCHW activations, [rows, features] for linear layers.  Dense
// contractions go through the separable-index GEMM engine (gemm.cuh); everything row-wise
// (LayerNorm, dueling/softmax/projection/loss, optimiser) is a warp- or CTA-per-row kernel with
// shuffle reductions.
#include <math.h>
#include <string.h>

#include <type_traits>

#include "conv_tc.cuh"
#include "conv_i8.cuh"
#include "conv_st.cuh"
#include "conv_wst.cuh"
#include "conv_dst.cuh"
#include "conv_wi8.cuh"
#include "head_fused.cuh"
#include "gemm.cuh"
#include <cstdlib>

namespace b2rl {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
static inline int64_t layer_out_elems(const b2rl_layer &l) {
    return l.kind == B2RL_LAYER_CONV ? (int64_t)l.out_c * l.out_h * l.out_w : (int64_t)l.out_c;
}
static inline int64_t layer_in_elems(const b2rl_layer &l) {
    return l.kind == B2RL_LAYER_CONV ? (int64_t)l.in_c * l.in_h * l.in_w : (int64_t)l.in_c;
}
static inline int64_t layer_w_elems(const b2rl_layer &l) {
    return l.kind == B2RL_LAYER_CONV ? (int64_t)l.out_c * l.in_c * l.ksize * l.ksize : (int64_t)l.out_c * l.in_c;
}
static inline bool needs_pre(const b2rl_layer &l) { return l.act == B2RL_ACT_GELU; }

struct LayerBuf {
    float *z = nullptr;      // pre-LayerNorm linear output (ln layers)
    float *pre = nullptr;    // pre-activation (only kept for GELU)
    float *a = nullptr;      // layer output
    float *stats = nullptr;  // mean, rstd per row (ln layers)
    float *g = nullptr;      // gradient w.r.t. the layer output (backward rows only)
};
struct PassBufs {
    LayerBuf enc[B2RL_MAX_ENC], val[B2RL_MAX_HEAD], adv[B2RL_MAX_HEAD];
    int64_t rows = 0;
};

struct Bump {
    char *base;
    size_t off = 0;
    explicit Bump(void *b) : base(static_cast<char *>(b)) {}
    template <typename T> T *take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

static void carve_layers(Bump &b, const b2rl_layer *layers, int n, LayerBuf *out, int64_t rows, int64_t grad_rows) {
    for (int i = 0; i < n; ++i) {
        const b2rl_layer &l = layers[i];
        const int64_t oe = layer_out_elems(l);
        out[i].a = b.take<float>(rows * oe);
        if (l.ln != B2RL_LN_NONE) {
            out[i].z = b.take<float>(rows * oe);
            out[i].stats = b.take<float>(rows * 2);
        }
        if (needs_pre(l)) out[i].pre = b.take<float>(rows * oe);
        if (grad_rows > 0) out[i].g = b.take<float>(grad_rows * oe);
    }
}
static void carve_pass(Bump &b, const b2rl_net_desc &net, PassBufs &pb, int64_t rows, int64_t grad_rows) {
    pb.rows = rows;
    carve_layers(b, net.enc, net.n_enc, pb.enc, rows, grad_rows);
    carve_layers(b, net.val, net.n_val, pb.val, rows, grad_rows);
    carve_layers(b, net.adv, net.n_adv, pb.adv, rows, grad_rows);
}

static size_t gemm_need(int64_t M, int64_t N, int64_t K, bool wgrad_t = false) {
    if (M <= 0 || N <= 0 || K <= 0 || M > INT32_MAX) return 0;
    return plan_gemm((int)M, (int)N, (int)K, use_big_tile(M, wgrad_t), sm_count()).partial_floats;
}
// exact bound of the split-K scratch over every GEMM a pass can launch
static size_t max_partial_floats(const b2rl_net_desc &net, int64_t rows, int64_t brows) {
    size_t m = 0;
    auto upd = [&](const b2rl_layer &l) {
        size_t n;
        if (l.kind == B2RL_LAYER_CONV) {
            const int64_t P = (int64_t)l.out_h * l.out_w, Kc = (int64_t)l.in_c * l.ksize * l.ksize;
            n = gemm_need(rows * P, l.out_c, Kc); if (n > m) m = n;
            n = gemm_need(rows / 2 * P, l.out_c, Kc); if (n > m) m = n;     // first layer runs per chunk
            if (brows) { n = gemm_need(Kc + 1, l.out_c, brows * P, true); if (n > m) m = n; }
        } else {
            n = gemm_need(rows, l.out_c, l.in_c); if (n > m) m = n;
            n = gemm_need(rows / 2, l.out_c, l.in_c); if (n > m) m = n;
            if (brows) {
                n = gemm_need(l.out_c, l.in_c + 1, brows); if (n > m) m = n;
                n = gemm_need(brows, l.in_c, l.out_c); if (n > m) m = n;
            }
        }
    };
    for (int i = 0; i < net.n_enc; ++i) upd(net.enc[i]);
    for (int i = 0; i < net.n_val; ++i) upd(net.val[i]);
    for (int i = 0; i < net.n_adv; ++i) upd(net.adv[i]);
    for (int i = 0; i < net.n_enc; ++i)
        if (net.enc[i].kind == B2RL_LAYER_CONV) {
            size_t n = conv_tc_wsplit_floats(net.enc[i]);
            if (n > m) m = n;
            if (brows) { n = conv_wgrad_tc_partial_floats(net.enc[i], brows, sm_count()); if (n > m) m = n; }
            if (brows) { n = conv_wgrad_st_partial_floats(net.enc[i], sm_count()); if (n > m) m = n; }
            if (brows && i == 0) {
                n = conv_wi8_scratch_bytes((net.enc[i].out_c + 15) / 16 * 16, net.enc[i].in_c * net.enc[i].ksize * net.enc[i].ksize,
                                           net.enc[i].out_c, sm_count()) / sizeof(float) + 64;
                if (n > m) m = n;
            }
            if (brows && i > 0) { n = conv_dgrad_tc_scratch_floats(net.enc[i]); if (n > m) m = n; }
            if (brows && i > 0) { n = conv_dst_scratch_floats(net.enc[i]); if (n > m) m = n; }
        }
    return m;
}

constexpr int kNormBlocks = 296;

struct LearnWS {
    float *weff_actor, *weff_target, *gweff;
    PassBufs online, target;
    int32_t *a_star;
    float *q_online, *q_target;     // [rows, A] expected values (DQN / debugging)
    float *tdist;                   // [B, N] target distribution of a*
    float *proj;                    // [B, N]
    float *cmat;                    // [N, N] (driver-shape mode)
    float *partial;
    size_t partial_floats;
    float *partial_tg;              // scratch of the target forward (runs concurrently with the online one)
    size_t partial_tg_floats;
    float *partial_bw;              // scratch of the weight-gradient side stream
    size_t partial_bw_floats;
    float *norm_partials;           // kNormBlocks
    float *wsum;                    // 1
    unsigned int *ticket;           // last-CTA counter of the fused loss tail
    // parameter-only preparation of a step (b2rl_rainbow_prep): digit planes of the first layer for both networks,
    // pre-split weights of the later convolutions per network — kept apart from the split-K scratch so that they can be
    // produced on a side stream while the sampler runs
    void *prep0;
    size_t prep0_bytes;
    float *prep_split[2][B2RL_MAX_ENC];
    size_t prep_split_floats[B2RL_MAX_ENC];
    float *prep_dsplit[B2RL_MAX_ENC];       // input-gradient weight layout of the online network's convolutions (backward)
    size_t prep_dsplit_floats[B2RL_MAX_ENC];
    size_t bytes;
};

static void carve_learn(const b2rl_net_desc &net, int64_t B, bool two_sided_online, void *base, LearnWS &ws) {
    Bump b(base);
    ws.weff_actor = b.take<float>(net.n_params);
    ws.weff_target = b.take<float>(net.n_params);
    ws.gweff = b.take<float>(net.n_params);
    const int64_t online_rows = two_sided_online ? 2 * B : B;
    carve_pass(b, net, ws.online, online_rows, B);
    carve_pass(b, net, ws.target, B, 0);
    ws.a_star = b.take<int32_t>(B);
    ws.q_online = b.take<float>(online_rows * net.n_actions);
    ws.q_target = b.take<float>(B * net.n_actions);
    ws.tdist = b.take<float>(B * net.n_atoms);
    ws.proj = b.take<float>(B * net.n_atoms);
    ws.cmat = b.take<float>((int64_t)net.n_atoms * net.n_atoms);
    ws.partial_floats = max_partial_floats(net, online_rows, B);
    ws.partial = b.take<float>(ws.partial_floats);
    ws.partial_tg_floats = max_partial_floats(net, B, 0);
    ws.partial_tg = b.take<float>(ws.partial_tg_floats);
    ws.partial_bw_floats = ws.partial_floats;
    ws.partial_bw = b.take<float>(ws.partial_bw_floats);
    ws.norm_partials = b.take<float>(kNormBlocks);
    ws.wsum = b.take<float>(4);
    ws.ticket = b.take<unsigned int>(4);
    ws.prep0 = nullptr; ws.prep0_bytes = 0;
    for (int i = 0; i < B2RL_MAX_ENC; ++i) {
        ws.prep_split[0][i] = ws.prep_split[1][i] = nullptr; ws.prep_split_floats[i] = 0;
        ws.prep_dsplit[i] = nullptr; ws.prep_dsplit_floats[i] = 0;
    }
    if (net.n_enc >= 1 && net.enc[0].kind == B2RL_LAYER_CONV) {
        const b2rl_layer &l0 = net.enc[0];
        const int K = l0.in_c * l0.ksize * l0.ksize;
        ws.prep0_bytes = conv_i8_scratch_bytes((l0.out_c + 15) / 16 * 16, (K + 31) / 32 * 32, 2);
        ws.prep0 = b.take<unsigned char>(ws.prep0_bytes);
    }
    for (int i = 1; i < net.n_enc; ++i)
        if (net.enc[i].kind == B2RL_LAYER_CONV) {
            ws.prep_split_floats[i] = conv_tc_wsplit_floats(net.enc[i]);
            ws.prep_split[0][i] = b.take<float>(ws.prep_split_floats[i]);
            ws.prep_split[1][i] = b.take<float>(ws.prep_split_floats[i]);
            ws.prep_dsplit_floats[i] = conv_dst_scratch_floats(net.enc[i]);
            ws.prep_dsplit[i] = b.take<float>(ws.prep_dsplit_floats[i]);
        }
    ws.bytes = b.off + 256;
}

// side stream + events for the fork/join inside rainbow_loss (one set per process; the callers of the
// loss are ordered among themselves, so one side stream keeps every dependency a stream order)
struct ForkJoin {
    cudaStream_t side = nullptr;       // target-network forward (highest priority)
    cudaStream_t side_bw = nullptr;    // weight gradients of the backward (default priority)
    cudaEvent_t fork = nullptr, join = nullptr, bw = nullptr;
    cudaStream_t prep = nullptr;       // parameter-only preparation of a step, concurrent with the sampler
    cudaEvent_t prep_fork = nullptr, prep_done = nullptr;
};
static int fork_join(ForkJoin **out) {
    static ForkJoin fj;
    if (!fj.side) {
        // the forward chain is the population's critical path (the next sampler waits on its priorities):
        // give its side stream the highest priority so its CTAs are placed ahead of overlapped backward work
        int lo = 0, hi = 0;
        B2RL_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        B2RL_CUDA(cudaStreamCreateWithPriority(&fj.side, cudaStreamNonBlocking, hi));
        B2RL_CUDA(cudaEventCreateWithFlags(&fj.fork, cudaEventDisableTiming));
        B2RL_CUDA(cudaEventCreateWithFlags(&fj.join, cudaEventDisableTiming));
        B2RL_CUDA(cudaStreamCreateWithFlags(&fj.side_bw, cudaStreamNonBlocking));
        B2RL_CUDA(cudaEventCreateWithFlags(&fj.bw, cudaEventDisableTiming));
        B2RL_CUDA(cudaStreamCreateWithPriority(&fj.prep, cudaStreamNonBlocking, hi));
        B2RL_CUDA(cudaEventCreateWithFlags(&fj.prep_fork, cudaEventDisableTiming));
        B2RL_CUDA(cudaEventCreateWithFlags(&fj.prep_done, cudaEventDisableTiming));
    }
    *out = &fj;
    return B2RL_OK;
}

struct FwdWS {
    float *weff;
    PassBufs pass;
    float *partial;
    size_t partial_floats;
    size_t bytes;
};
static void carve_fwd(const b2rl_net_desc &net, int64_t rows, void *base, FwdWS &ws) {
    Bump b(base);
    ws.weff = b.take<float>(net.n_params);
    carve_pass(b, net, ws.pass, rows, 0);
    ws.partial_floats = max_partial_floats(net, rows, 0);
    ws.partial = b.take<float>(ws.partial_floats);
    ws.bytes = b.off + 256;
}

static int validate_net(const b2rl_net_desc &net) {
    B2RL_CHECK_ARG(net.n_enc >= 1 && net.n_enc <= B2RL_MAX_ENC, "n_enc out of range");
    B2RL_CHECK_ARG(net.n_val >= 1 && net.n_val <= B2RL_MAX_HEAD, "n_val out of range");
    B2RL_CHECK_ARG(net.n_adv >= 0 && net.n_adv <= B2RL_MAX_HEAD, "n_adv out of range");
    B2RL_CHECK_ARG((net.kind == B2RL_NET_RAINBOW) == (net.n_adv > 0), "rainbow nets need an advantage head");
    B2RL_CHECK_ARG(net.n_actions >= 1 && net.n_atoms >= 1, "bad action/atom count");
    return B2RL_OK;
}

// ------------------------------------------------------------------------------------------
// NoisyLinear: W_eff = mu + sigma * eps   (custom_components.py:97-99; mul then add, no FMA)
// ------------------------------------------------------------------------------------------
struct Seg { float *dst; const float *mu; const float *sigma; const float *eps; int64_t n; };
struct SegTable { Seg s[4 * (B2RL_MAX_ENC + 2 * B2RL_MAX_HEAD)]; int n; };    // room for two networks

__global__ void compose_kernel(SegTable t) {
    const Seg sg = t.s[blockIdx.y];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < sg.n; i += (int64_t)gridDim.x * blockDim.x)
        sg.dst[i] = __fadd_rn(sg.mu[i], __fmul_rn(sg.sigma[i], sg.eps[i]));
}
// grad_sigma = grad_Weff * eps  (autograd of mu + sigma*eps); grad_mu is grad_Weff itself.
__global__ void noisy_grad_kernel(SegTable t, int accumulate) {
    const Seg sg = t.s[blockIdx.y];   // dst = grad_sigma, mu = grad_Weff, eps = eps
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < sg.n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = sg.eps ? sg.mu[i] * sg.eps[i] : sg.mu[i];
        sg.dst[i] = accumulate ? sg.dst[i] + v : v;
    }
}

template <typename F>
static void for_each_layer(const b2rl_net_desc &net, F &&f) {
    for (int i = 0; i < net.n_enc; ++i) f(net.enc[i]);
    for (int i = 0; i < net.n_val; ++i) f(net.val[i]);
    for (int i = 0; i < net.n_adv; ++i) f(net.adv[i]);
}

// W_eff of every noisy layer of up to two parameter sets (online + target network) in one launch
static int compose_weights(const b2rl_net_desc &net, const float *params, const float *eps, float *weff,
                           cudaStream_t s, const float *params2 = nullptr, const float *eps2 = nullptr,
                           float *weff2 = nullptr) {
    SegTable t;
    t.n = 0;
    int64_t maxn = 0;
    auto add = [&](const float *p, const float *e, float *w) {
        for_each_layer(net, [&](const b2rl_layer &l) {
            if (!l.noisy) return;
            const int64_t nw = layer_w_elems(l);
            t.s[t.n++] = Seg{w + l.w_off, p + l.w_off, p + l.ws_off, e + l.we_off, nw};
            t.s[t.n++] = Seg{w + l.b_off, p + l.b_off, p + l.bs_off, e + l.be_off, (int64_t)l.out_c};
            if (nw > maxn) maxn = nw;
        });
    };
    add(params, eps, weff);
    if (params2) add(params2, eps2, weff2);
    if (t.n == 0) return B2RL_OK;
    int bx = (int)((maxn + 255) / 256);
    if (bx > 64) bx = 64;
    compose_kernel<<<dim3(bx, t.n), 256, 0, s>>>(t);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

// ------------------------------------------------------------------------------------------
// LayerNorm (+activation) forward / backward — one warp per row.
// ------------------------------------------------------------------------------------------
__global__ void ln_fwd_kernel(const float *__restrict__ z, const float *__restrict__ w, const float *__restrict__ b,
                              int act, float *__restrict__ a, float *__restrict__ pre, float *__restrict__ stats,
                              int64_t rows, int n) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float *zr = z + row * n;
    float s = 0.f;
    for (int c = lane; c < n; c += 32) s += zr[c];
    const float mean = warp_sum(s) / (float)n;
    float v = 0.f;
    for (int c = lane; c < n; c += 32) { const float d = zr[c] - mean; v += d * d; }
    const float var = warp_sum(v) / (float)n;          // biased variance, eps = 1e-5 (nn.LayerNorm)
    const float rstd = 1.0f / sqrtf(var + 1e-5f);
    if (lane == 0) { stats[row * 2] = mean; stats[row * 2 + 1] = rstd; }
    for (int c = lane; c < n; c += 32) {
        float y = (zr[c] - mean) * rstd;
        if (w) y = y * w[c] + b[c];
        if (pre) pre[row * n + c] = y;
        a[row * n + c] = act_fwd(act, y);
    }
}

// column sums for the LayerNorm affine gradients: dw[c] = sum_r gy*xhat, db[c] = sum_r gy.
// 32 columns x 32 row-groups per CTA, fixed-order smem reduction over the row groups.
__global__ void ln_affine_grad_kernel(const float *__restrict__ g, const float *__restrict__ z,
                                      const float *__restrict__ stats, const float *__restrict__ a,
                                      const float *__restrict__ pre, int act, float *__restrict__ dw,
                                      float *__restrict__ db, int64_t rows, int n, int accumulate) {
    __shared__ float sw[32][33], sb[32][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float w = 0.f, b = 0.f;
    if (c < n) {
        for (int64_t r = ry; r < rows; r += 32) {
            const int64_t o = r * n + c;
            const float gy = g[o] * act_bwd(act, pre ? pre[o] : 0.f, a[o]);
            const float xhat = (z[o] - stats[r * 2]) * stats[r * 2 + 1];
            w += gy * xhat;
            b += gy;
        }
    }
    sw[ry][cx] = w; sb[ry][cx] = b;
    __syncthreads();
    if (ry == 0 && c < n) {
        float tw = 0.f, tb = 0.f;
        for (int i = 0; i < 32; ++i) { tw += sw[i][cx]; tb += sb[i][cx]; }
        dw[c] = accumulate ? dw[c] + tw : tw;
        db[c] = accumulate ? db[c] + tb : tb;
    }
}

// g (dL/da) -> dL/dz in place.
__global__ void ln_bwd_kernel(float *__restrict__ g, const float *__restrict__ z, const float *__restrict__ stats,
                              const float *__restrict__ w, const float *__restrict__ a,
                              const float *__restrict__ pre, int act, int64_t rows, int n) {
    const int lane = threadIdx.x & 31;
    const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float mean = stats[row * 2], rstd = stats[row * 2 + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane; c < n; c += 32) {
        const int64_t o = row * n + c;
        const float gy = g[o] * act_bwd(act, pre ? pre[o] : 0.f, a[o]);
        const float gx = w ? gy * w[c] : gy;
        const float xhat = (z[o] - mean) * rstd;
        s1 += gx;
        s2 += gx * xhat;
    }
    s1 = warp_sum(s1) / (float)n;
    s2 = warp_sum(s2) / (float)n;
    for (int c = lane; c < n; c += 32) {
        const int64_t o = row * n + c;
        const float gy = g[o] * act_bwd(act, pre ? pre[o] : 0.f, a[o]);
        const float gx = w ? gy * w[c] : gy;
        const float xhat = (z[o] - mean) * rstd;
        g[o] = rstd * (gx - s1 - xhat * s2);
    }
}

__global__ void act_bwd_kernel(float *__restrict__ g, const float *__restrict__ a, const float *__restrict__ pre,
                               int act, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        g[i] *= act_bwd(act, pre ? pre[i] : 0.f, a[i]);
}

// ------------------------------------------------------------------------------------------
// one layer forward
// ------------------------------------------------------------------------------------------
struct ObsChunk {          // first-layer input rows [row0, row0+rows) come from this source
    const void *ptr;
    const int64_t *gather; // nullable: ring row of each batch row
    int64_t rows;
};

struct Scratch { float *partial; size_t floats; };

// B2RL_DISABLE_TC=1 forces every convolution onto the fp32 FFMA engine (A/B testing, parity triage)
static bool tc_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("B2RL_DISABLE_TC");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

// run f(std::integral_constant<int, ELEM>) for the element kind of an operand that may read
// observations (only first-layer operands are ever not plain fp32)
template <typename F>
static int dispatch_elem(int kind, F &&f) {
    switch (kind) {
        case EL_U8: return f(std::integral_constant<int, EL_U8>{});
        case EL_F32_NORM: return f(std::integral_constant<int, EL_F32_NORM>{});
        default: return f(std::integral_constant<int, EL_F32>{});
    }
}
using OpW = OpTraits<EL_F32, MAP_STRIDE, MAP_STRIDE, true, false>;     // weights [out, K], k contiguous

static int layer_forward(const b2rl_net_desc &net, const b2rl_layer &l, const float *W, const float *bias,
                         const float *params, const float *x_prev, const ObsChunk *chunks, int n_chunks,
                         int64_t rows, const LayerBuf &lb, const Scratch &sc, cudaStream_t s, bool reuse_split = false,
                         float *presplit = nullptr, size_t presplit_floats = 0) {   // presplit: this layer's weights already split there
    const int64_t oe = layer_out_elems(l);
    const bool first = x_prev == nullptr;
    int n_runs = first ? n_chunks : 1;
    int64_t row0 = 0;
    // uint8 frames with integer bounds: integer tensor path, every chunk in ONE launch (conv_i8.cuh)
    if (first && tc_enabled() && n_chunks >= 1 && n_chunks <= kI8MaxSegs &&
        conv_i8_ok(l, net.obs_u8 != 0, net.normalize != 0, net.obs_low, net.obs_high, chunks[0].ptr, 1) &&
        (n_chunks == 1 || reinterpret_cast<uintptr_t>(chunks[1].ptr) % 4 == 0)) {
        ConvI8Job jobs[kI8MaxSegs];
        int64_t r0 = 0;
        for (int i = 0; i < n_chunks; ++i) {
            jobs[i] = ConvI8Job{chunks[i].ptr, chunks[i].gather, chunks[i].rows, 1, {lb.a + r0 * oe, nullptr}};
            r0 += chunks[i].rows;
        }
        const float *Ws[2] = {W, W}, *bs[2] = {bias, bias};
        const int rc8 = launch_conv_fwd_i8(l, net.normalize != 0, net.obs_low, net.obs_high, Ws, bs, 1, jobs, n_chunks,
                                           sc.partial, sc.floats * sizeof(float), s, reuse_split);
        if (rc8 != 1) return rc8;
    }
    for (int run = 0; run < n_runs; ++run) {
        const int64_t r = first ? chunks[run].rows : rows;
        Operand A, Bm;
        Epilogue epi;
        if (first) {
            A.ptr = chunks[run].ptr;
            A.u8 = net.obs_u8;
            A.normalize = net.normalize;
            A.low = net.obs_low;
            A.high = net.obs_high;
        } else {
            A.ptr = x_prev;
        }
        const int64_t *gather = first ? chunks[run].gather : nullptr;
        int M, N, K;
        float *out_ptr = (l.ln != B2RL_LN_NONE ? lb.z : lb.a) + row0 * oe;
        if (l.kind == B2RL_LAYER_CONV) {
            const int P = l.out_h * l.out_w, KK = l.ksize * l.ksize;
            M = (int)(r * P); N = l.out_c; K = l.in_c * KK;
            A.row = map_pixel(P, l.out_w, (int64_t)l.in_c * l.in_h * l.in_w, l.stride * l.in_w, l.stride, gather);
            A.red = map_kernel(l.ksize, l.in_h * l.in_w, l.in_w);
            Bm.ptr = W; Bm.row = map_stride(K); Bm.red = map_stride(1);
            epi.om = map_pixel(P, l.out_w, (int64_t)l.out_c * P, l.out_w, 1);
            epi.on = map_stride(P);
        } else {
            M = (int)r; N = l.out_c; K = l.in_c;
            A.row = map_stride(K, gather); A.red = map_stride(1);
            Bm.ptr = W; Bm.row = map_stride(K); Bm.red = map_stride(1);
            epi.om = map_stride(N); epi.on = map_stride(1);
        }
        epi.out = out_ptr;
        epi.bias = bias;
        if (l.ln == B2RL_LN_NONE) {
            epi.act = l.act;
            epi.pre_out = lb.pre ? lb.pre + row0 * oe : nullptr;
        }
        int rc = 1;
        // fp32 activations of an earlier layer: receptive fields staged in shared memory by the TMA unit (conv_st.cuh)
        const bool pre_ok = presplit != nullptr && !first;
        float *wsp = pre_ok ? presplit : sc.partial;
        const size_t wsp_cap = pre_ok ? presplit_floats : sc.floats;
        if (l.kind == B2RL_LAYER_CONV && tc_enabled() && l.ln == B2RL_LN_NONE && !first && lb.pre == nullptr)
            rc = launch_conv_fwd_st(l, x_prev, W, bias, out_ptr, r, wsp, wsp_cap, s, reuse_split || pre_ok);
        if (rc == 1 && l.kind == B2RL_LAYER_CONV && tc_enabled() && l.ln == B2RL_LN_NONE)
            rc = launch_conv_fwd_tc(l, A, W, bias, out_ptr, lb.pre ? lb.pre + row0 * oe : nullptr, r, wsp, wsp_cap,
                                    s, run > 0 || reuse_split || pre_ok);   // tcgen05 3xTF32 (pre-split weights live in the split-K scratch;
                                                   // the second observation chunk reuses the first one's split)
        if (rc == 1 && l.kind == B2RL_LAYER_CONV)
            rc = dispatch_elem(A.elem_kind(), [&](auto ek) {
                return launch_igemm<OpTraits<decltype(ek)::value, MAP_PIXEL, MAP_KERNEL, true, false>, OpW,
                                    EpiTraits<EPI_STORE, MAP_PIXEL, MAP_STRIDE>>(A, Bm, epi, M, N, K, sc.partial,
                                                                                   sc.floats, s);
            });
        else if (rc == 1)
            rc = dispatch_elem(A.elem_kind(), [&](auto ek) {
                return launch_igemm<OpTraits<decltype(ek)::value, MAP_STRIDE, MAP_STRIDE, true, false>, OpW,
                                    EpiTraits<EPI_STORE, MAP_STRIDE, MAP_STRIDE>>(A, Bm, epi, M, N, K, sc.partial,
                                                                                    sc.floats, s);
            });
        if (rc != B2RL_OK) return rc;
        row0 += r;
    }
    if (l.ln != B2RL_LN_NONE) {
        B2RL_CHECK_ARG(l.kind == B2RL_LAYER_LINEAR, "LayerNorm is only supported after linear layers");
        const float *lw = l.ln == B2RL_LN_AFFINE ? params + l.lnw_off : nullptr;
        const float *lbias = l.ln == B2RL_LN_AFFINE ? params + l.lnb_off : nullptr;
        const int warps = 4;
        ln_fwd_kernel<<<(int)((rows + warps - 1) / warps), warps * 32, 0, s>>>(lb.z, lw, lbias, l.act, lb.a, lb.pre,
                                                                               lb.stats, rows, l.out_c);
        B2RL_LAUNCH_CHECK();
    }
    return B2RL_OK;
}

static inline const float *eff_w(const b2rl_layer &l, const float *params, const float *weff, bool use_noise) {
    return (l.noisy && use_noise) ? weff + l.w_off : params + l.w_off;
}
static inline const float *eff_b(const b2rl_layer &l, const float *params, const float *weff, bool use_noise) {
    return (l.noisy && use_noise) ? weff + l.b_off : params + l.b_off;
}

static int forward_pass(const b2rl_net_desc &net, const float *params, const float *weff, bool use_noise,
                        const ObsChunk *chunks, int n_chunks, int64_t rows, PassBufs &pb, const Scratch &sc,
                        cudaStream_t s, bool first_done = false,      // first_done: pb.enc[0].a already holds layer 0's output
                        float *const *presplit = nullptr, const size_t *presplit_floats = nullptr) {
    const float *x = nullptr;
    for (int i = 0; i < net.n_enc; ++i) {
        if (i == 0 && first_done) { x = pb.enc[0].a; continue; }
        const b2rl_layer &l = net.enc[i];
        int rc = layer_forward(net, l, eff_w(l, params, weff, use_noise), eff_b(l, params, weff, use_noise), params, x,
                               chunks, n_chunks, rows, pb.enc[i], sc, s, false, presplit ? presplit[i] : nullptr,
                               presplit ? presplit_floats[i] : 0);
        if (rc != B2RL_OK) return rc;
        x = pb.enc[i].a;
    }
    const float *latent = x;
    // ---- head: one fused launch for both chains when the tile fits in shared memory ----------------
    {
        HeadDesc hd;
        hd.n_val = net.n_val; hd.n_adv = net.n_adv;
        hd.latent = net.val[0].in_c;
        int maxdim = hd.latent;
        bool ok = tc_enabled() && (net.n_val + net.n_adv) <= kHeadMaxLayers;
        auto fill = [&](const b2rl_layer &l, const LayerBuf &lb, HeadLayer &h) {
            if (l.kind != B2RL_LAYER_LINEAR) ok = false;
            h.w = eff_w(l, params, weff, use_noise); h.b = eff_b(l, params, weff, use_noise);
            h.lnw = l.ln == B2RL_LN_AFFINE ? params + l.lnw_off : nullptr;
            h.lnb = l.ln == B2RL_LN_AFFINE ? params + l.lnb_off : nullptr;
            h.a = lb.a; h.z = lb.z; h.pre = lb.pre; h.stats = lb.stats;
            h.in = l.in_c; h.out = l.out_c; h.ln = l.ln; h.act = l.act;
            if (l.in_c > maxdim) maxdim = l.in_c;
            if (l.out_c > maxdim) maxdim = l.out_c;
        };
        for (int i = 0; i < net.n_val && ok; ++i) fill(net.val[i], pb.val[i], hd.l[i]);
        for (int i = 0; i < net.n_adv && ok; ++i) fill(net.adv[i], pb.adv[i], hd.l[net.n_val + i]);
        hd.maxdim = maxdim;
        const size_t smem = sizeof(float) * ((size_t)2 * kHeadRows * head_pitch(maxdim) + (kHeadThreads / 32) * kHeadRows);
        if (ok && smem <= 160 * 1024) {
            static bool attr_set = false;
            if (!attr_set) {
                B2RL_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                attr_set = true;
            }
            head_fwd_kernel<<<dim3((unsigned)((rows + kHeadRows - 1) / kHeadRows), (net.n_val > 0 && net.n_adv > 0) ? 2 : 1),
                              kHeadThreads, smem, s>>>(hd, latent, rows);
            B2RL_LAUNCH_CHECK();
            return B2RL_OK;
        }
    }
    const float *xv = latent;
    for (int i = 0; i < net.n_val; ++i) {
        const b2rl_layer &l = net.val[i];
        int rc = layer_forward(net, l, eff_w(l, params, weff, use_noise), eff_b(l, params, weff, use_noise), params, xv,
                               nullptr, 0, rows, pb.val[i], sc, s);
        if (rc != B2RL_OK) return rc;
        xv = pb.val[i].a;
    }
    const float *xa = latent;
    for (int i = 0; i < net.n_adv; ++i) {
        const b2rl_layer &l = net.adv[i];
        int rc = layer_forward(net, l, eff_w(l, params, weff, use_noise), eff_b(l, params, weff, use_noise), params, xa,
                               nullptr, 0, rows, pb.adv[i], sc, s);
        if (rc != B2RL_OK) return rc;
        xa = pb.adv[i].a;
    }
    return B2RL_OK;
}

// ------------------------------------------------------------------------------------------
// Rainbow head: dueling combine, softmax, clamp, expectation / argmax / projection / loss.
// One CTA per batch row; dynamic smem holds x[A*N].
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce_sum(float v, float *red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    return t;
}

// x[a][n] = v[n] + adv[a][n] - mean_a adv[.][n]    (custom_modules.py:149-153)
__device__ __forceinline__ void dueling_to_smem(const float *__restrict__ v, const float *__restrict__ adv, int A,
                                                int N, float *x) {
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float m = 0.f;
        for (int a = 0; a < A; ++a) m += adv[a * N + n];
        m = m / (float)A;
        const float vn = v[n];
        for (int a = 0; a < A; ++a) x[a * N + n] = vn + adv[a * N + n] - m;
    }
    __syncthreads();
}

// softmax over atoms of action row `a` (in place in smem), clamp(min=1e-3); one warp per action.
__device__ __forceinline__ float softmax_clamp_row(float *xr, int N, const float *__restrict__ support) {
    const int lane = threadIdx.x & 31;
    float m = -INFINITY;
    for (int n = lane; n < N; n += 32) m = fmaxf(m, xr[n]);
    m = warp_max(m);
    float s = 0.f;
    for (int n = lane; n < N; n += 32) { const float e = expf(xr[n] - m); xr[n] = e; s += e; }
    s = warp_sum(s);
    float q = 0.f;
    for (int n = lane; n < N; n += 32) {
        float p = xr[n] / s;
        p = fmaxf(p, 1e-3f);                       // clamp(min=1e-3), no renormalisation (quirk Q8)
        xr[n] = p;
        if (support) q += p * support[n];
    }
    return warp_sum(q);
}

// mode 0: write q[row][a]; argmax -> a_star (int32) and/or argmax64
__global__ void rainbow_q_kernel(const float *__restrict__ v, const float *__restrict__ adv,
                                 const float *__restrict__ support, int A, int N, float *__restrict__ q_out,
                                 int32_t *__restrict__ a_star, int64_t *__restrict__ argmax64) {
    extern __shared__ float sm[];
    float *x = sm;            // A*N
    float *q = sm + A * N;    // A
    const int64_t row = blockIdx.x;
    dueling_to_smem(v + row * N, adv + row * (int64_t)A * N, A, N, x);
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int a = warp; a < A; a += nwarps) {
        const float qa = softmax_clamp_row(x + a * N, N, support);
        if (lane == 0) q[a] = qa;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = 0;
        float bv = q[0];
        for (int a = 1; a < A; ++a)
            if (q[a] > bv) { bv = q[a]; best = a; }    // first maximum, like torch.argmax
        if (a_star) a_star[row] = best;
        if (argmax64) argmax64[row] = best;
    }
    if (q_out)
        for (int a = threadIdx.x; a < A; a += blockDim.x) q_out[row * A + a] = q[a];
}

// RainbowQNetwork.forward(q=False, log=...) (custom_modules.py:153-160): per-atom distributions [rows, A, N] —
// softmax + clamp(min=1e-3), or log_softmax when log_mode (no clamp, like the reference).
__global__ void rainbow_dist_kernel(const float *__restrict__ v, const float *__restrict__ adv, int A, int N,
                                    int log_mode, float *__restrict__ out) {
    extern __shared__ float sm[];
    float *x = sm;            // A*N
    const int64_t row = blockIdx.x;
    dueling_to_smem(v + row * N, adv + row * (int64_t)A * N, A, N, x);
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int a = warp; a < A; a += nwarps) {
        float *xr = x + a * N;
        if (!log_mode) {
            softmax_clamp_row(xr, N, nullptr);
        } else {
            float m = -INFINITY;
            for (int n = lane; n < N; n += 32) m = fmaxf(m, xr[n]);
            m = warp_max(m);
            float sum = 0.f;
            for (int n = lane; n < N; n += 32) sum += expf(xr[n] - m);
            sum = warp_sum(sum);
            const float lse = logf(sum);
            for (int n = lane; n < N; n += 32) xr[n] = xr[n] - m - lse;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < A * N; k += blockDim.x) out[row * (int64_t)A * N + k] = x[k];
}

struct ProjCfg { float gamma, v_min, v_max, delta_z; };

// Target distribution of a* and (canonical shapes) the C51 projection, sequential per row in the
// reference's index_add_ order so the fp32 sums are bit-faithful (dqn_rainbow.py:318-360).
// When v_on / adv_on are given (the online network's outputs on next_obs) the kernel first selects
// a* = argmax_a E[Z_online(next_obs, a)] itself (dqn_rainbow.py:315; same code as rainbow_q_kernel)
// and records it in a_star; otherwise it reads a_star.
// returns the shared-memory row holding the projected distribution of this row (valid when proj != NULL)
__device__ __forceinline__ float *rainbow_target_row(float *sm, const float *__restrict__ v, const float *__restrict__ adv,
                                      int32_t *__restrict__ a_star, const float *__restrict__ reward,
                                      const float *__restrict__ done, const float *__restrict__ support, ProjCfg pc,
                                      int A, int N, float *__restrict__ tdist, float *__restrict__ proj,
                                      const float *__restrict__ v_on, const float *__restrict__ adv_on) {
    float *x = sm;                 // A*N
    float *pr = sm + A * N;        // N   projected
    float *wl = pr + N;            // N
    float *wu = wl + N;            // N
    int *Li = reinterpret_cast<int *>(wu + N);   // N
    int *Ui = Li + N;                            // N
    const int64_t row = blockIdx.x;
    __shared__ int s_as;
    if (v_on != nullptr) {                 // host guarantees A <= 3*N (q[] borrows the projection scratch)
        float *q = pr;
        dueling_to_smem(v_on + row * N, adv_on + row * (int64_t)A * N, A, N, x);
        const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
        for (int a = warp; a < A; a += nwarps) {
            const float qa = softmax_clamp_row(x + a * N, N, support);
            if (lane == 0) q[a] = qa;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int best = 0;
            float bv = q[0];
            for (int a = 1; a < A; ++a)
                if (q[a] > bv) { bv = q[a]; best = a; }    // first maximum, like torch.argmax
            s_as = best;
            a_star[row] = best;
        }
        __syncthreads();
    } else {
        if (threadIdx.x == 0) s_as = a_star[row];
        __syncthreads();
    }
    const int as = s_as;
    dueling_to_smem(v + row * N, adv + row * (int64_t)A * N, A, N, x);
    float *p = x + as * N;
    if ((threadIdx.x >> 5) == 0) softmax_clamp_row(p, N, nullptr);
    __syncthreads();
    if (tdist)
        for (int n = threadIdx.x; n < N; n += blockDim.x) tdist[row * N + n] = p[n];
    if (!proj) return pr;
    const float r = reward[row], d = done[row];
    const float g = __fmul_rn(__fsub_rn(1.0f, d), pc.gamma);              // (1 - dones) * gamma
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float tz = __fadd_rn(r, __fmul_rn(g, support[n]));                  // rewards + (..) * support
        tz = fminf(fmaxf(tz, pc.v_min), pc.v_max);                          // clamp
        const float b = __fdiv_rn(__fsub_rn(tz, pc.v_min), pc.delta_z);
        int L = (int)floorf(b), U = (int)ceilf(b);
        if (U > 0 && U == L) L -= 1;                                        // quirk Q9 order
        if ((N - 1) > L && U == L) U += 1;
        wl[n] = __fsub_rn((float)U, b);
        wu[n] = __fsub_rn(b, (float)L);
        // delta_z is a rounded float: b can exceed N-1 by an ulp (U == N).  The reference's index_add_ would then
        // raise; here the stray mass stays in the last atom instead of landing in the neighbouring scratch row
        Li[n] = L < 0 ? 0 : (L > N - 1 ? N - 1 : L);
        Ui[n] = U < 0 ? 0 : (U > N - 1 ? N - 1 : U);
        pr[n] = 0.f;
    }
    __syncthreads();
    // index_add_ order kept per destination bin, bins in parallel: bin m receives its l-contributions in ascending n,
    // then its u-contributions in ascending n — the same fp32 additions in the same order as the reference's two
    // sequential index_add_ calls (dqn_rainbow.py:350-360), without one thread walking all 2N updates
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
        float acc = 0.f;
        for (int n = 0; n < N; ++n)
            if (Li[n] == m) acc = __fadd_rn(acc, __fmul_rn(p[n], wl[n]));
        for (int n = 0; n < N; ++n)
            if (Ui[n] == m) acc = __fadd_rn(acc, __fmul_rn(p[n], wu[n]));
        pr[m] = acc;
    }
    __syncthreads();
    for (int n = threadIdx.x; n < N; n += blockDim.x) proj[row * N + n] = pr[n];
    return pr;
}

__global__ void rainbow_target_kernel(const float *__restrict__ v, const float *__restrict__ adv,
                                      int32_t *__restrict__ a_star, const float *__restrict__ reward,
                                      const float *__restrict__ done, const float *__restrict__ support, ProjCfg pc,
                                      int A, int N, float *__restrict__ tdist, float *__restrict__ proj,
                                      const float *__restrict__ v_on, const float *__restrict__ adv_on) {
    extern __shared__ float sm[];
    rainbow_target_row(sm, v, adv, a_star, reward, done, support, pc, A, N, tdist, proj, v_on, adv_on);
}

// Driver-shape (quirk Q2) projection: reward/done arrive [B,1,1], the reference's broadcasting
// makes every row's projected target the SUM over all B Bellman shifts applied to that row's
// target distribution:  proj[j][m] = sum_n tdist[j][n] * C[n][m],
// C[n][m] = sum_i ( wl[i][n] [L[i][n]==m] + wu[i][n] [U[i][n]==m] ).
__global__ void q2_cmat_kernel(const float *__restrict__ reward, const float *__restrict__ done,
                               const float *__restrict__ support, ProjCfg pc, int64_t B, int N,
                               float *__restrict__ cmat) {
    // one CTA per source atom n; accumulate over samples in order
    const int n = blockIdx.x;
    extern __shared__ float row[];   // N
    for (int m = threadIdx.x; m < N; m += blockDim.x) row[m] = 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int64_t i = 0; i < B; ++i) {
            const float g = __fmul_rn(__fsub_rn(1.0f, done[i]), pc.gamma);
            float tz = __fadd_rn(reward[i], __fmul_rn(g, support[n]));
            tz = fminf(fmaxf(tz, pc.v_min), pc.v_max);
            const float b = __fdiv_rn(__fsub_rn(tz, pc.v_min), pc.delta_z);
            int L = (int)floorf(b), U = (int)ceilf(b);
            if (U > 0 && U == L) L -= 1;
            if ((N - 1) > L && U == L) U += 1;
            row[L] += __fsub_rn((float)U, b);
            row[U] += __fsub_rn(b, (float)L);
        }
    }
    __syncthreads();
    for (int m = threadIdx.x; m < N; m += blockDim.x) cmat[n * N + m] = row[m];
}
__global__ void q2_project_kernel(const float *__restrict__ tdist, const float *__restrict__ cmat, int N,
                                  float *__restrict__ proj) {
    const int64_t j = blockIdx.x;
    for (int m = threadIdx.x; m < N; m += blockDim.x) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s += tdist[j * N + n] * cmat[n * N + m];
        proj[j * N + m] = s;
    }
}

// Cross-entropy of the taken action's log-softmax against the projected target, plus dL/dlogits.
// loss_i = -sum_n proj[i][n] * log_softmax(x[a_i])[n]                      (dqn_rainbow.py:362-367)
// d loss_i / d x[a_i][n] = softmax[n] * sum_m proj[i][m] - proj[i][n]
// dueling backward: dv[n] = dx[n];  dadv[a][n] = [a==a_i] dx[n] - dx[n]/A
// projrow: this row's projected distribution (global or shared memory)
__device__ __forceinline__ void rainbow_loss_row(float *sm, const float *__restrict__ v, const float *__restrict__ adv,
                                    const float *__restrict__ action, const float *projrow,
                                    const float *__restrict__ weights, int weights_mode, int64_t B, int A, int N,
                                    int accumulate, float *__restrict__ loss_elem, float *__restrict__ dv,
                                    float *__restrict__ dadv) {
    float *x = sm;              // N (action row only)
    float *red = sm + N;        // 32
    const int64_t row = blockIdx.x;
    const float *vr = v + row * N;
    const float *ar = adv + row * (int64_t)A * N;
    const int ai = (int)(long long)action[row];            // actions.squeeze().long()
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float m = 0.f;
        for (int a = 0; a < A; ++a) m += ar[a * N + n];
        m = m / (float)A;
        x[n] = vr[n] + ar[ai * N + n] - m;
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, x[n]);
    // block max via shuffles + smem
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
    float s = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) s += expf(x[n] - mx);
    s = block_reduce_sum(s, red);
    const float lse = logf(s);
    float l = 0.f, ps = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float lp = x[n] - mx - lse;
        const float pj = projrow[n];
        l -= pj * lp;
        ps += pj;
    }
    l = block_reduce_sum(l, red);
    ps = block_reduce_sum(ps, red);
    // per-row gradient scale of the scalar loss
    float gs;
    if (weights_mode == 1) gs = weights[row] / (float)B;                    // mean(l * w)
    else if (weights_mode == 2) {                                            // mean(l) * mean(w)  (Q1)
        float ws = 0.f;
        for (int64_t i = threadIdx.x; i < B; i += blockDim.x) ws += weights[i];
        ws = block_reduce_sum(ws, red);
        gs = (ws / (float)B) / (float)B;
    } else gs = 1.0f / (float)B;
    if (threadIdx.x == 0) loss_elem[row] = accumulate ? loss_elem[row] + l : l;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const float sm_n = expf(x[n] - mx) / s;
        const float dx = gs * (sm_n * ps - projrow[n]);
        dv[row * N + n] = dx;
        const float sh = dx / (float)A;
        for (int a = 0; a < A; ++a) dadv[row * (int64_t)A * N + a * N + n] = (a == ai ? dx : 0.f) - sh;
    }
}

__global__ void rainbow_loss_kernel(const float *__restrict__ v, const float *__restrict__ adv,
                                    const float *__restrict__ action, const float *__restrict__ proj,
                                    const float *__restrict__ weights, int weights_mode, int64_t B, int A, int N,
                                    int accumulate, float *__restrict__ loss_elem, float *__restrict__ dv,
                                    float *__restrict__ dadv) {
    extern __shared__ float sm[];
    rainbow_loss_row(sm, v, adv, action, proj + (int64_t)blockIdx.x * N, weights, weights_mode, B, A, N, accumulate, loss_elem,
                     dv, dadv);
}

// The loss tail of _dqn_loss for canonical shapes in ONE launch (dqn_rainbow.py:315-367, :434, :487-488): CTA = batch
// row — a* selection + target distribution + C51 projection, then the cross-entropy of the online row against the
// projection still sitting in shared memory (+ dL/dlogits), then the LAST CTA to finish (ticket counter) reduces the
// per-sample losses to the scalar loss in a fixed order and writes the priorities.
__global__ void rainbow_tail_kernel(const float *__restrict__ v_tg, const float *__restrict__ adv_tg,
                                    int32_t *__restrict__ a_star, const float *__restrict__ reward,
                                    const float *__restrict__ done, const float *__restrict__ support, ProjCfg pc, int A,
                                    int N, float *__restrict__ proj, const float *__restrict__ v_nx,
                                    const float *__restrict__ adv_nx, const float *__restrict__ v_ob,
                                    const float *__restrict__ adv_ob, const float *__restrict__ action,
                                    const float *__restrict__ weights, int weights_mode, int64_t B, int accumulate,
                                    float *__restrict__ loss_elem, float *__restrict__ dv, float *__restrict__ dadv,
                                    float prior_eps, float *__restrict__ loss_scalar, float *__restrict__ priorities,
                                    unsigned int *__restrict__ ticket) {
    extern __shared__ float sm[];
    float *pr = rainbow_target_row(sm, v_tg, adv_tg, a_star, reward, done, support, pc, A, N, nullptr, proj, v_nx, adv_nx);
    __syncthreads();
    // the loss part reuses the dueling scratch x[0 .. N + 32) — pr (behind the A*N floats) stays intact
    rainbow_loss_row(sm, v_ob, adv_ob, action, pr, weights, weights_mode, B, A, N, accumulate, loss_elem, dv, dadv);
    __shared__ unsigned int s_ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    float *red = sm;
    float a = 0.f, w = 0.f;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
        const float l = __ldcg(loss_elem + i);
        if (weights_mode == 1) a += l * weights[i];
        else a += l;
        if (weights_mode == 2) w += weights[i];
        if (priorities) priorities[i] = l + prior_eps;
    }
    a = block_reduce_sum(a, red);
    w = block_reduce_sum(w, red);
    if (threadIdx.x == 0) {
        float L = a / (float)B;
        if (weights_mode == 2) L *= w / (float)B;
        *loss_scalar = L;
        *ticket = 0u;
    }
}

// scalar loss + priorities (single CTA): dqn_rainbow.py:434 / :487-488
__global__ void rainbow_scalar_kernel(const float *__restrict__ loss_elem, const float *__restrict__ weights,
                                      int weights_mode, int64_t B, float prior_eps, float *__restrict__ loss_scalar,
                                      float *__restrict__ priorities) {
    __shared__ float red[32];
    float a = 0.f, w = 0.f;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
        const float l = loss_elem[i];
        if (weights_mode == 1) a += l * weights[i];
        else a += l;
        if (weights_mode == 2) w += weights[i];
        if (priorities) priorities[i] = l + prior_eps;
    }
    a = block_reduce_sum(a, red);
    w = block_reduce_sum(w, red);
    if (threadIdx.x == 0) {
        float L = a / (float)B;
        if (weights_mode == 2) L *= w / (float)B;
        *loss_scalar = L;
    }
}

// ------------------------------------------------------------------------------------------
// DQN loss (dqn.py:290-314): y = r + gamma*q_t*(1-d); loss = mean((q[a]-y)^2); single CTA.
// ------------------------------------------------------------------------------------------
__global__ void dqn_loss_kernel(const float *__restrict__ q_eval, const float *__restrict__ q_next_online,
                                const float *__restrict__ q_next_target, const float *__restrict__ action,
                                const float *__restrict__ reward, const float *__restrict__ done, float gamma,
                                int double_dqn, int64_t B, int A, float *__restrict__ dq,
                                float *__restrict__ loss_elem, float *__restrict__ loss_scalar) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
        const float *qt = q_next_target + i * A;
        float qsel;
        if (double_dqn) {
            const float *qo = q_next_online + i * A;
            int best = 0;
            for (int a = 1; a < A; ++a) if (qo[a] > qo[best]) best = a;
            qsel = qt[best];
        } else {
            qsel = qt[0];
            for (int a = 1; a < A; ++a) qsel = fmaxf(qsel, qt[a]);
        }
        const float y = __fadd_rn(reward[i], __fmul_rn(__fmul_rn(gamma, qsel), __fsub_rn(1.0f, done[i])));
        const int ai = (int)(long long)action[i];
        const float diff = q_eval[i * A + ai] - y;
        acc += diff * diff;
        if (loss_elem) loss_elem[i] = diff * diff;
        for (int a = 0; a < A; ++a) dq[i * A + a] = (a == ai) ? 2.0f * diff / (float)B : 0.f;
    }
    acc = block_reduce_sum(acc, red);
    if (threadIdx.x == 0) *loss_scalar = acc / (float)B;
}

// copy / argmax of plain Q outputs
__global__ void q_argmax_kernel(const float *__restrict__ q, int A, int64_t rows, float *__restrict__ q_out,
                                int64_t *__restrict__ argmax64) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= rows) return;
    int best = 0;
    for (int a = 0; a < A; ++a) {
        if (q_out) q_out[i * A + a] = q[i * A + a];
        if (q[i * A + a] > q[i * A + best]) best = a;
    }
    if (argmax64) argmax64[i] = best;
}

// ------------------------------------------------------------------------------------------
// backward of one layer
// ------------------------------------------------------------------------------------------
// Weight gradients do not feed the dL/dx chain: when a side stream is given they are enqueued there
// (own scratch) right after the layer's dL/dz exists, and the caller joins before the optimiser.
struct WgradSide {
    cudaStream_t s;
    Scratch sc;
    cudaEvent_t ev;
};
static int fork_to(const WgradSide *side, cudaStream_t from) {
    B2RL_CUDA(cudaEventRecord(side->ev, from));
    B2RL_CUDA(cudaStreamWaitEvent(side->s, side->ev, 0));
    return B2RL_OK;
}

// activations whose derivative is a function of the layer's OUTPUT alone (no pre-activation buffer needed)
static inline bool act_needs_output_only(int act) { return act == B2RL_ACT_RELU || act == B2RL_ACT_ELU || act == B2RL_ACT_TANH; }

static int layer_backward(const b2rl_net_desc &net, const b2rl_layer &l, const float *W, const float *params,
                          const float *x_in,            // input activations of this layer (backward rows)
                          const ObsChunk *obs,          // first layer: where the observations come from
                          const LayerBuf &lb, int64_t row_off,  // forward buffers + offset of the backward rows
                          float *g_out,                 // dL/d(layer output) for the B backward rows
                          float *g_in, bool accumulate_gin, float *grads, float *gweff, int accumulate_grads,
                          int64_t B, const Scratch &sc, cudaStream_t s, const WgradSide *side = nullptr,
                          bool skip_own_act = false,     // the layer above (or the head) already applied this layer's activation backward
                          int below_act = B2RL_ACT_NONE,   // fold the activation backward of the layer below (output x_in) into g_in
                          const float *dsplit = nullptr) {  // input-gradient weights already laid out (b2rl_rainbow_prep)
    const int64_t oe = layer_out_elems(l);
    const float *a = lb.a + row_off * oe;
    const float *pre = lb.pre ? lb.pre + row_off * oe : nullptr;
    // first convolution over uint8 frames: the integer weight-gradient path also applies the activation backward
    // (its channel-maximum pass reads the same planes), one launch less on the chain
    bool act_in_wgrad = false;
    Operand obs_op;
    if (obs && l.kind == B2RL_LAYER_CONV && l.ln == B2RL_LN_NONE && tc_enabled() && !(side && g_in) && !skip_own_act) {
        obs_op.ptr = obs->ptr; obs_op.u8 = net.obs_u8; obs_op.normalize = net.normalize; obs_op.low = net.obs_low; obs_op.high = net.obs_high;
        act_in_wgrad = conv_wgrad_i8_ok(l, obs_op, B, sc.partial, sc.floats * sizeof(float));
    }
    // (1) through activation (+LayerNorm)
    if (act_in_wgrad || skip_own_act) {
        // nothing here
    } else if (l.ln != B2RL_LN_NONE) {
        const float *z = lb.z + row_off * oe;
        const float *st = lb.stats + row_off * 2;
        if (l.ln == B2RL_LN_AFFINE) {
            ln_affine_grad_kernel<<<(l.out_c + 31) / 32, 1024, 0, s>>>(g_out, z, st, a, pre, l.act, grads + l.lnw_off,
                                                                        grads + l.lnb_off, B, l.out_c, accumulate_grads);
            B2RL_LAUNCH_CHECK();
        }
        const float *lw = l.ln == B2RL_LN_AFFINE ? params + l.lnw_off : nullptr;
        ln_bwd_kernel<<<(int)((B + 3) / 4), 128, 0, s>>>(g_out, z, st, lw, a, pre, l.act, B, l.out_c);
        B2RL_LAUNCH_CHECK();
    } else if (l.act != B2RL_ACT_NONE) {
        const int64_t n = B * oe;
        int blocks = (int)((n + 255) / 256);
        if (blocks > sm_count() * 8) blocks = sm_count() * 8;
        act_bwd_kernel<<<blocks, 256, 0, s>>>(g_out, a, pre, l.act, n);
        B2RL_LAUNCH_CHECK();
    }
    // (2) weight + bias gradient:  dW[o][k] = sum_r G[r][o] * X[r][k],  db[o] = sum_r G[r][o]
    float *gw = (l.noisy ? gweff : grads) + l.w_off;
    float *gb = (l.noisy ? gweff : grads) + l.b_off;
    const int acc_w = l.noisy ? 0 : accumulate_grads;   // noisy: accumulate happens in noisy_grad
    cudaStream_t sw = s;
    const Scratch *scw = &sc;
    if (side && g_in) {                                   // overlap this layer's dW with its dX and everything below
        int rcf = fork_to(side, s);
        if (rcf != B2RL_OK) return rcf;
        sw = side->s;
        scw = &side->sc;
    }
    {
        Operand A, Bm;
        Epilogue epi;
        epi.kind = EPI_WGRAD;
        epi.out = gw; epi.db = gb; epi.accumulate = acc_w;
        int M, N, K;
        int rc;
        if (l.kind == B2RL_LAYER_CONV) {
            const int P = l.out_h * l.out_w, KK = l.ksize * l.ksize;
            const int Kc = l.in_c * KK;
            // dW^T[tap][co] = sum_pix im2col[tap][pix] * dOut[co][pix]: the im2col operand is the
            // 128-row A tile (taps are hoisted per thread, one pixel decode per k-tile), dOut the
            // 32-wide B tile; the extra row tap == Kc reads 1.0 and yields the bias gradient.
            M = Kc + 1; N = l.out_c; K = (int)(B * P);
            epi.kind = EPI_WGRAD_T;
            epi.wcols = Kc;
            if (obs) {
                A.ptr = obs->ptr; A.u8 = net.obs_u8; A.normalize = net.normalize; A.low = net.obs_low; A.high = net.obs_high;
            } else A.ptr = x_in;
            A.row = map_kernel(l.ksize, l.in_h * l.in_w, l.in_w);
            A.red = map_pixel(P, l.out_w, (int64_t)l.in_c * l.in_h * l.in_w, l.stride * l.in_w, l.stride,
                              obs ? obs->gather : nullptr);
            A.ones_row = Kc;
            Bm.ptr = g_out; Bm.row = map_stride(P); Bm.red = map_pixel(P, l.out_w, (int64_t)l.out_c * P, l.out_w, 1);
            rc = 1;
            if (tc_enabled() && obs)   // uint8 frames: integer tensor path, frame bytes as the MN-major operand (conv_wi8.cuh)
                rc = launch_conv_wgrad_i8(l, A, g_out, gw, gb, acc_w, B, scw->partial, scw->floats * sizeof(float), sw,
                                          (act_in_wgrad && l.act != B2RL_ACT_NONE) ? a : nullptr, pre, l.act);
            if (rc == 1 && act_in_wgrad) { set_error("integer weight-gradient path refused a layer it had accepted"); return B2RL_ECUDA; }
            if (rc == 1 && tc_enabled())   // staged operands (conv_wst.cuh)
                rc = launch_conv_wgrad_st(l, A, g_out, gw, gb, acc_w, B, scw->partial, scw->floats, sw);
            if (rc == 1 && tc_enabled())   // tcgen05 3xTF32, gathered operands, split over pixels
                rc = launch_conv_wgrad_tc(l, A, g_out, gw, gb, acc_w, B, scw->partial, scw->floats, sw);
            if (rc == 1) rc = dispatch_elem(A.elem_kind(), [&](auto ek) {
                return launch_igemm<OpTraits<decltype(ek)::value, MAP_KERNEL, MAP_PIXEL, true, true>,
                                    OpTraits<EL_F32, MAP_STRIDE, MAP_PIXEL, true, false>,
                                    EpiTraits<EPI_WGRAD_T, MAP_STRIDE, MAP_STRIDE>>(A, Bm, epi, M, N, K, scw->partial,
                                                                                      scw->floats, sw);
            });
        } else {
            M = l.out_c; N = l.in_c + 1; K = (int)B;
            epi.wcols = l.in_c;
            A.ptr = g_out; A.row = map_stride(1); A.red = map_stride(l.out_c);
            if (obs) {
                Bm.ptr = obs->ptr; Bm.u8 = net.obs_u8; Bm.normalize = net.normalize; Bm.low = net.obs_low; Bm.high = net.obs_high;
            } else Bm.ptr = x_in;
            Bm.row = map_stride(1);
            Bm.red = map_stride(l.in_c, obs ? obs->gather : nullptr);
            Bm.ones_row = l.in_c;
            rc = dispatch_elem(Bm.elem_kind(), [&](auto ek) {
                return launch_igemm<OpTraits<EL_F32, MAP_STRIDE, MAP_STRIDE, false, false>,
                                    OpTraits<decltype(ek)::value, MAP_STRIDE, MAP_STRIDE, false, true>,
                                    EpiTraits<EPI_WGRAD, MAP_STRIDE, MAP_STRIDE>>(A, Bm, epi, M, N, K, scw->partial,
                                                                                    scw->floats, sw);
            });
        }
        if (rc != B2RL_OK) return rc;
    }
    // (3) input gradient
    if (g_in) {
        Operand A, Bm;
        Epilogue epi;
        int rc;
        if (l.kind == B2RL_LAYER_CONV) {
            const int P = l.out_h * l.out_w, KK = l.ksize * l.ksize;
            const int Kc = l.in_c * KK;
            if (!accumulate_gin && tc_enabled()) {     // parity-class convolutions on tcgen05 (no atomics)
                const int rc_st = launch_conv_dgrad_st(l, g_out, W, g_in, B, sc.partial, sc.floats, s, dsplit);
                if (rc_st != 1) return rc_st;
                const int rc_tc = launch_conv_dgrad_tc(l, g_out, W, g_in, B, sc.partial, sc.floats, s);
                if (rc_tc != 1) return rc_tc;
            }
            if (!accumulate_gin) B2RL_CUDA(cudaMemsetAsync(g_in, 0, sizeof(float) * B * layer_in_elems(l), s));
            A.ptr = g_out; A.row = map_pixel(P, l.out_w, (int64_t)l.out_c * P, l.out_w, 1); A.red = map_stride(P);
            Bm.ptr = W; Bm.row = map_stride(1); Bm.red = map_stride(Kc);
            epi.kind = EPI_ATOMIC; epi.out = g_in;
            epi.om = map_pixel(P, l.out_w, (int64_t)l.in_c * l.in_h * l.in_w, l.stride * l.in_w, l.stride);
            epi.on = map_kernel(l.ksize, l.in_h * l.in_w, l.in_w);
            rc = launch_igemm<OpTraits<EL_F32, MAP_PIXEL, MAP_STRIDE, false, false>,
                              OpTraits<EL_F32, MAP_STRIDE, MAP_STRIDE, false, false>,
                              EpiTraits<EPI_ATOMIC, MAP_PIXEL, MAP_KERNEL>>(A, Bm, epi, (int)(B * P), Kc, l.out_c, nullptr,
                                                                           0, s);
        } else {
            A.ptr = g_out; A.row = map_stride(l.out_c); A.red = map_stride(1);
            Bm.ptr = W; Bm.row = map_stride(1); Bm.red = map_stride(l.in_c);
            epi.out = g_in; epi.om = map_stride(l.in_c); epi.on = map_stride(1); epi.accumulate = accumulate_gin ? 1 : 0;
            if (below_act != B2RL_ACT_NONE && !accumulate_gin && x_in) { epi.mask_a = x_in; epi.mask_act = below_act; }
            rc = launch_igemm<OpTraits<EL_F32, MAP_STRIDE, MAP_STRIDE, true, false>,
                              OpTraits<EL_F32, MAP_STRIDE, MAP_STRIDE, false, false>,
                              EpiTraits<EPI_STORE, MAP_STRIDE, MAP_STRIDE>>(A, Bm, epi, (int)B, l.in_c, l.out_c, sc.partial,
                                                                           sc.floats, s);
        }
        if (rc != B2RL_OK) return rc;
    }
    return B2RL_OK;
}

static int backward_pass(const b2rl_net_desc &net, const float *params, const float *weff, bool use_noise,
                         const float *eps, PassBufs &pb, int64_t row_off, int64_t B, const ObsChunk &obs, float *grads,
                         float *gweff, int accumulate, const Scratch &sc, cudaStream_t s,
                         const WgradSide *side = nullptr, float *const *dsplit = nullptr) {
    const LayerBuf &lat = pb.enc[net.n_enc - 1];
    const float *latent = lat.a + row_off * layer_out_elems(net.enc[net.n_enc - 1]);
    float *g_latent = lat.g;
    // heads: two fused launches (dZ/dX chain, then every dW/db) when the tile fits in shared memory
    bool head_done = false;
    bool top_act_fused = false;
    const b2rl_layer &ltop = net.enc[net.n_enc - 1];
    const bool top_fusable = net.n_enc >= 2 && ltop.ln == B2RL_LN_NONE && act_needs_output_only(ltop.act);
    if (tc_enabled() && B <= 2048 && net.n_val + net.n_adv <= kHeadMaxLayers) {
        HeadBwdDesc hd;
        hd.n_val = net.n_val; hd.n_adv = net.n_adv; hd.latent = net.val[0].in_c;
        hd.g_latent = g_latent; hd.accumulate = accumulate; hd.n_ln = 0;
        hd.latent_a = nullptr; hd.latent_act = B2RL_ACT_NONE;
        const int n_tiles = (int)((B + kHeadRows - 1) / kHeadRows);
        int maxdim = hd.latent, ctas = 0;
        size_t part_used = 0;
        bool ok = true;
        auto fill = [&](const b2rl_layer *layers, LayerBuf *bufs, int i, int slot) {
            const b2rl_layer &l = layers[i];
            const LayerBuf &lb = bufs[i];
            HeadBwdLayer &h = hd.l[slot];
            if (l.kind != B2RL_LAYER_LINEAR) { ok = false; return; }
            const int64_t oe = l.out_c;
            h.w = eff_w(l, params, weff, use_noise);
            h.lnw = l.ln == B2RL_LN_AFFINE ? params + l.lnw_off : nullptr;
            h.a = lb.a + row_off * oe;
            h.z = lb.z ? lb.z + row_off * oe : nullptr;
            h.pre = lb.pre ? lb.pre + row_off * oe : nullptr;
            h.stats = lb.stats ? lb.stats + row_off * 2 : nullptr;
            h.x = i == 0 ? latent : bufs[i - 1].a + row_off * layers[i - 1].out_c;
            h.g = lb.g;
            h.dw = (l.noisy ? gweff : grads) + l.w_off;
            h.db = (l.noisy ? gweff : grads) + l.b_off;
            h.dlnw = h.dlnb = h.lnpart = nullptr;
            if (l.ln == B2RL_LN_AFFINE) {
                h.dlnw = grads + l.lnw_off; h.dlnb = grads + l.lnb_off;
                h.lnpart = (side ? side->sc.partial : sc.partial) + part_used;
                part_used += (size_t)n_tiles * 2 * l.out_c;
                hd.ln_layer[hd.n_ln++] = slot;
            }
            h.in = l.in_c; h.out = l.out_c; h.ln = l.ln; h.act = l.act;
            h.acc_w = l.noisy ? 0 : accumulate;   // noisy: the accumulate happens in noisy_grad
            if (l.in_c > maxdim) maxdim = l.in_c;
            if (l.out_c > maxdim) maxdim = l.out_c;
            hd.wg_start[slot] = ctas;
            ctas += (l.out_c + kHeadThreads / 32 - 1) / (kHeadThreads / 32);
        };
        for (int i = 0; i < net.n_val && ok; ++i) fill(net.val, pb.val, i, i);
        for (int i = 0; i < net.n_adv && ok; ++i) fill(net.adv, pb.adv, i, net.n_val + i);
        hd.wg_start[net.n_val + net.n_adv] = ctas;
        hd.maxdim = maxdim;
        const size_t smem = sizeof(float) * ((size_t)(3 + kHeadThreads / 32) * kHeadRows * head_pitch(maxdim) +
                                             (size_t)kHeadRows * hd.latent + (kHeadThreads / 32) * kHeadRows);
        if (ok && smem <= 160 * 1024 && part_used <= (side ? side->sc.floats : sc.floats) && maxdim <= kHeadWgMaxIn) {
            static bool attr_set = false;
            if (!attr_set) {
                B2RL_CUDA(cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
                B2RL_CUDA(cudaFuncSetAttribute(head_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)(kHeadWgSmemFloats * sizeof(float))));
                attr_set = true;
            }
            if (top_fusable) { hd.latent_a = latent; hd.latent_act = ltop.act; top_act_fused = true; }
            head_bwd_kernel<<<n_tiles, kHeadThreads, smem, s>>>(hd, B);
            B2RL_LAUNCH_CHECK();
            cudaStream_t sw = s;
            if (side) {
                int rcf = fork_to(side, s);
                if (rcf != B2RL_OK) return rcf;
                sw = side->s;
            }
            head_wgrad_kernel<<<ctas + hd.n_ln, kHeadThreads, kHeadWgSmemFloats * sizeof(float), sw>>>(hd, B, n_tiles);
            B2RL_LAUNCH_CHECK();
            head_done = true;
        }
    }
    for (int head = 0; head < 2 && !head_done; ++head) {
        const b2rl_layer *layers = head == 0 ? net.val : net.adv;
        LayerBuf *bufs = head == 0 ? pb.val : pb.adv;
        const int n = head == 0 ? net.n_val : net.n_adv;
        for (int i = n - 1; i >= 0; --i) {
            const b2rl_layer &l = layers[i];
            const float *x_in = i == 0 ? latent : bufs[i - 1].a + row_off * layer_out_elems(layers[i - 1]);
            float *g_in = i == 0 ? g_latent : bufs[i - 1].g;
            const bool acc_gin = (i == 0 && head == 1);
            int rc = layer_backward(net, l, eff_w(l, params, weff, use_noise), params, x_in, nullptr, bufs[i], row_off,
                                    bufs[i].g, g_in, acc_gin, grads, gweff, accumulate, B, sc, s, nullptr);
            if (rc != B2RL_OK) return rc;
        }
    }
    // encoder
    bool skip_act = top_act_fused;          // the fused head backward already went through the top layer's activation
    for (int i = net.n_enc - 1; i >= 0; --i) {
        const b2rl_layer &l = net.enc[i];
        const float *x_in = i == 0 ? nullptr : pb.enc[i - 1].a + row_off * layer_out_elems(net.enc[i - 1]);
        float *g_in = i == 0 ? nullptr : pb.enc[i - 1].g;
        // a linear layer's input gradient can carry the activation backward of the layer below it (not the first layer:
        // the integer weight-gradient path applies that one itself)
        int below = B2RL_ACT_NONE;
        if (i >= 2 && l.kind == B2RL_LAYER_LINEAR && net.enc[i - 1].ln == B2RL_LN_NONE && act_needs_output_only(net.enc[i - 1].act))
            below = net.enc[i - 1].act;
        int rc = layer_backward(net, l, eff_w(l, params, weff, use_noise), params, x_in, i == 0 ? &obs : nullptr,
                                pb.enc[i], row_off, pb.enc[i].g, g_in, false, grads, gweff, accumulate, B, sc, s, side,
                                skip_act, below, (dsplit && conv_dgrad_st_shape_ok(l)) ? dsplit[i] : nullptr);
        skip_act = below != B2RL_ACT_NONE;
        if (rc != B2RL_OK) return rc;
    }
    if (side) {                                           // every weight gradient must have landed before the tail
        B2RL_CUDA(cudaEventRecord(side->ev, side->s));
        B2RL_CUDA(cudaStreamWaitEvent(s, side->ev, 0));
    }
    // noisy layers: grad_mu = grad_Weff, grad_sigma = grad_Weff * eps (autograd of mu + sigma*eps)
    SegTable tm, tsg;
    tm.n = tsg.n = 0;
    int64_t maxn = 0;
    for_each_layer(net, [&](const b2rl_layer &l) {
        if (!l.noisy) return;
        const int64_t nw = layer_w_elems(l);
        tsg.s[tsg.n++] = Seg{grads + l.ws_off, gweff + l.w_off, nullptr, use_noise ? eps + l.we_off : nullptr, nw};
        tsg.s[tsg.n++] = Seg{grads + l.bs_off, gweff + l.b_off, nullptr, use_noise ? eps + l.be_off : nullptr,
                             (int64_t)l.out_c};
        tm.s[tm.n++] = Seg{grads + l.w_off, gweff + l.w_off, nullptr, nullptr, nw};
        tm.s[tm.n++] = Seg{grads + l.b_off, gweff + l.b_off, nullptr, nullptr, (int64_t)l.out_c};
        if (nw > maxn) maxn = nw;
    });
    if (tsg.n > 0) {
        int bx = (int)((maxn + 255) / 256);
        if (bx > 64) bx = 64;
        if (use_noise) {            // sigma and mu gradients of every noisy layer in ONE launch (segments are independent)
            for (int i = 0; i < tm.n; ++i) tsg.s[tsg.n++] = tm.s[i];
            noisy_grad_kernel<<<dim3(bx, tsg.n), 256, 0, s>>>(tsg, accumulate);
            B2RL_LAUNCH_CHECK();
        } else {
            if (!accumulate)        // eval-mode noisy layers: W = mu, sigma gets no gradient
                for (int i = 0; i < tsg.n; ++i) B2RL_CUDA(cudaMemsetAsync(tsg.s[i].dst, 0, sizeof(float) * tsg.s[i].n, s));
            noisy_grad_kernel<<<dim3(bx, tm.n), 256, 0, s>>>(tm, accumulate);
            B2RL_LAUNCH_CHECK();
        }
    }
    return B2RL_OK;
}

}  // namespace b2rl

namespace b2rl {

// ------------------------------------------------------------------------------------------
// learn() tail: clip_grad_norm_ + Adam + Polyak  (dqn_rainbow.py:473-483, optimizer_wrapper.py)
// ------------------------------------------------------------------------------------------
struct AdamCfg {
    int clip;
    float max_norm;
    float w1;            // 1 - beta1
    float beta2, w2;     // beta2, 1 - beta2
    float neg_step;      // -(lr / bias_correction1)
    float bc2_sqrt;      // sqrt(bias_correction2)
    float eps;
    float tau, one_minus_tau;
};

__global__ void sqnorm_kernel(const float *__restrict__ g, int64_t n, float *__restrict__ partials) {
    __shared__ float red[32];
    float s = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        s += g[i] * g[i];
    s = block_reduce_sum(s, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void adam_polyak_kernel(float *__restrict__ p, float *__restrict__ g, float *__restrict__ m,
                                   float *__restrict__ v, float *__restrict__ tgt, int64_t n,
                                   const float *__restrict__ partials, int nparts, AdamCfg c,
                                   const b2rl_step_state *__restrict__ state) {
    if (state) {        // graph-replayed step: this step's scalars, same double arithmetic as the host path
        c.neg_step = (float)(-__ddiv_rn(state->lr, state->bias_correction1));
        c.bc2_sqrt = (float)__dsqrt_rn(state->bias_correction2);
    }
    __shared__ float coef_s;
    if (threadIdx.x < 32) {                                 // every block combines the partial norms in the same order
        float coef = 1.f;
        if (c.clip) {
            float tot = 0.f;
            for (int i = threadIdx.x; i < nparts; i += 32) tot += partials[i];
            tot = warp_sum(tot);
            const float total_norm = sqrtf(tot);
            coef = c.max_norm / (total_norm + 1e-6f);      // clip_grad_norm_: clamp(coef, max=1)
            coef = coef > 1.f ? 1.f : coef;
        }
        if (threadIdx.x == 0) coef_s = coef;
    }
    __syncthreads();
    const float coef = coef_s;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i];
        if (c.clip) { gi = gi * coef; g[i] = gi; }
        float mi = m[i], vi = v[i];
        mi = fmaf(c.w1, gi - mi, mi);                        // exp_avg.lerp_(grad, 1-beta1)
        vi = vi * c.beta2 + c.w2 * gi * gi;                  // mul_(beta2).addcmul_(g, g, 1-beta2)
        m[i] = mi; v[i] = vi;
        const float denom = sqrtf(vi) / c.bc2_sqrt + c.eps;
        const float pi = p[i] + (c.neg_step * mi) / denom;   // addcdiv_(exp_avg, denom, value=-step_size)
        p[i] = pi;
        if (tgt) tgt[i] = __fadd_rn(__fmul_rn(c.tau, pi), __fmul_rn(c.one_minus_tau, tgt[i]));   // soft_update
    }
}

// ------------------------------------------------------------------------------------------
// NoisyLinear.reset_noise (custom_components.py:116-131)
// ------------------------------------------------------------------------------------------
struct NoiseSeg { float *eps_w; float *eps_b; int in, out; int64_t z_off; int which; };
struct NoiseTable { NoiseSeg s[2 * (B2RL_MAX_ENC + 2 * B2RL_MAX_HEAD)]; int n; };      // room for two networks in one launch

__device__ __forceinline__ float scale_noise(float x) {    // x.sign() * x.abs().sqrt()
    const float r = __fsqrt_rn(fabsf(x));
    return x > 0.f ? r : (x < 0.f ? -r : 0.f);
}

template <bool kPhilox>
__global__ void noise_reset_kernel(NoiseTable t, const float *__restrict__ normals, uint64_t seed, uint64_t offset,
                                   const b2rl_step_state *__restrict__ state = nullptr, int which = 0) {
    const NoiseSeg sg = t.s[blockIdx.y];
    if (state) offset = state->noise_offset[which >= 0 ? which : sg.which];     // which < 0: per segment (two networks, one launch)
    const int64_t n = (int64_t)sg.in * sg.out;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int o = (int)(e / sg.in), i = (int)(e - (int64_t)o * sg.in);
        float zi, zo;
        if (kPhilox) {
            zi = philox_normal(seed, offset + (uint64_t)(sg.z_off + i), 0x4E4F4953ull);
            zo = philox_normal(seed, offset + (uint64_t)(sg.z_off + sg.in + o), 0x4E4F4953ull);
        } else {
            zi = normals[sg.z_off + i];
            zo = normals[sg.z_off + sg.in + o];
        }
        const float fo = scale_noise(zo);
        sg.eps_w[e] = __fmul_rn(fo, scale_noise(zi));       // epsilon_out.ger(epsilon_in)
        if (i == 0) sg.eps_b[o] = fo;
    }
}

static int64_t build_noise_table(const b2rl_net_desc &net, float *eps, NoiseTable &t, bool append = false, int which = 0) {
    if (!append) t.n = 0;
    int64_t z = 0;
    for_each_layer(net, [&](const b2rl_layer &l) {
        if (!l.noisy) return;
        t.s[t.n++] = NoiseSeg{eps ? eps + l.we_off : nullptr, eps ? eps + l.be_off : nullptr, l.in_c, l.out_c, z, which};
        z += l.in_c + l.out_c;
    });
    return z;
}

static int noise_reset(const b2rl_net_desc &net, float *eps, const float *normals, uint64_t seed, uint64_t offset,
                       cudaStream_t s, const b2rl_step_state *state = nullptr, int which = 0) {
    NoiseTable t;
    build_noise_table(net, eps, t);
    if (t.n == 0) return B2RL_OK;
    int64_t maxn = 0;
    for (int i = 0; i < t.n; ++i) {
        const int64_t n = (int64_t)t.s[i].in * t.s[i].out;
        if (n > maxn) maxn = n;
    }
    int bx = (int)((maxn + 255) / 256);
    if (bx > 64) bx = 64;
    if (normals) noise_reset_kernel<false><<<dim3(bx, t.n), 256, 0, s>>>(t, normals, 0, 0);
    else noise_reset_kernel<true><<<dim3(bx, t.n), 256, 0, s>>>(t, nullptr, seed, offset, state, which);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

static bool has_noisy(const b2rl_net_desc &net) {
    bool any = false;
    for_each_layer(net, [&](const b2rl_layer &l) { any |= l.noisy != 0; });
    return any;
}

static int head_smem_check(size_t bytes) {
    B2RL_CHECK_ARG(bytes <= 48 * 1024, "n_actions * n_atoms too large for the head kernels (%zu B smem)", bytes);
    return B2RL_OK;
}

static int optim_step(const b2rl_net_desc &net, const b2rl_learn_cfg &cfg, const b2rl_learn_bufs &bufs, LearnWS &ws,
                      cudaStream_t s) {
    const int64_t n = net.n_params;
    if (cfg.clip) {
        sqnorm_kernel<<<kNormBlocks, 256, 0, s>>>(bufs.grads, n, ws.norm_partials);
        B2RL_LAUNCH_CHECK();
    }
    AdamCfg c;
    c.clip = cfg.clip;
    c.max_norm = (float)cfg.max_grad_norm;
    c.w1 = (float)(1.0 - cfg.beta1);
    c.beta2 = (float)cfg.beta2;
    c.w2 = (float)(1.0 - cfg.beta2);
    c.neg_step = (float)(-(cfg.lr / cfg.bias_correction1));
    c.bc2_sqrt = (float)sqrt(cfg.bias_correction2);
    c.eps = (float)cfg.adam_eps;
    c.tau = (float)cfg.tau;
    c.one_minus_tau = (float)(1.0 - cfg.tau);
    int blocks = (int)((n + 255) / 256);
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    adam_polyak_kernel<<<blocks, 256, 0, s>>>(bufs.actor_params, bufs.grads, bufs.exp_avg, bufs.exp_avg_sq,
                                              bufs.target_params, n, ws.norm_partials, kNormBlocks, c, bufs.step_state);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

// Everything of a loss pass that depends on the parameters only — W = mu + sigma * eps of the noisy layers, the int8 digit
// planes of the first layer (both networks), the tf32 hi/lo split of the later convolutions — on a library-owned side stream
// forked from `s`, so that it runs under the sampler that precedes the loss; rainbow_loss(prepped) joins it.
static bool prep_first_i8(const b2rl_net_desc &net, const b2rl_learn_bufs &bufs) {
    const b2rl_layer &l0 = net.enc[0];
    return tc_enabled() && !l0.noisy &&
           conv_i8_ok(l0, net.obs_u8 != 0, net.normalize != 0, net.obs_low, net.obs_high, bufs.next_obs, 2) &&
           reinterpret_cast<uintptr_t>(bufs.obs) % 4 == 0;
}
static int rainbow_prep(const b2rl_net_desc &net, const b2rl_learn_cfg &cfg, const b2rl_learn_bufs &bufs, LearnWS &ws,
                        cudaStream_t s) {
    ForkJoin *fj = nullptr;
    int rc;
    if ((rc = fork_join(&fj)) != B2RL_OK) return rc;
    B2RL_CUDA(cudaEventRecord(fj->prep_fork, s));
    B2RL_CUDA(cudaStreamWaitEvent(fj->prep, fj->prep_fork, 0));
    cudaStream_t sp = fj->prep;
    if (cfg.use_noise) {
        if ((rc = compose_weights(net, bufs.actor_params, bufs.actor_eps, ws.weff_actor, sp, bufs.target_params,
                                  bufs.target_eps, ws.weff_target)) != B2RL_OK)
            return rc;
    }
    if (prep_first_i8(net, bufs) && ws.prep0) {
        const b2rl_layer &l0 = net.enc[0];
        const float *Ws[2] = {bufs.actor_params + l0.w_off, bufs.target_params + l0.w_off};
        if ((rc = launch_weight_digits(l0, net.normalize != 0, net.obs_low, net.obs_high, Ws, 2, ws.prep0, ws.prep0_bytes, sp)) != B2RL_OK)
            return rc;
    }
    for (int i = 1; i < net.n_enc; ++i) {
        const b2rl_layer &l = net.enc[i];
        if (l.kind != B2RL_LAYER_CONV || l.noisy || !tc_enabled() || ws.prep_split[0][i] == nullptr) continue;
        for (int which = 0; which < 2; ++which) {
            const float *W = (which == 0 ? bufs.actor_params : bufs.target_params) + l.w_off;
            if ((rc = launch_weight_split(l, W, ws.prep_split[which][i], ws.prep_split_floats[i], sp)) != B2RL_OK) return rc;
        }
        // the backward of this step (same parameters: the optimiser runs after it) takes its input-gradient weights from here
        if (ws.prep_dsplit[i] && conv_dgrad_st_shape_ok(l)) {
            rc = launch_dgrad_st_weight_split(l, bufs.actor_params + l.w_off, ws.prep_dsplit[i], ws.prep_dsplit_floats[i], sp);
            if (rc != B2RL_OK && rc != 1) return rc;
        }
    }
    B2RL_CUDA(cudaMemsetAsync(ws.ticket, 0, sizeof(unsigned int), sp));      // the loss tail's last-CTA counter
    B2RL_CUDA(cudaEventRecord(fj->prep_done, sp));
    return B2RL_OK;
}

static int rainbow_loss(const b2rl_net_desc &net, const b2rl_learn_cfg &cfg, const b2rl_learn_bufs &bufs, LearnWS &ws,
                        cudaStream_t s) {
    const int64_t B = cfg.batch;
    const int A = net.n_actions, N = net.n_atoms;
    const bool noise = cfg.use_noise != 0;
    const bool prepped = (cfg.reserved_ & 1) != 0;          // b2rl_rainbow_prep was enqueued for this pass
    int rc;
    if (prepped) {
        ForkJoin *fjp = nullptr;
        if ((rc = fork_join(&fjp)) != B2RL_OK) return rc;
        B2RL_CUDA(cudaStreamWaitEvent(s, fjp->prep_done, 0));
    }
    if (noise && !prepped) {
        if ((rc = compose_weights(net, bufs.actor_params, bufs.actor_eps, ws.weff_actor, s, bufs.target_params,
                                  bufs.target_eps, ws.weff_target)) != B2RL_OK)
            return rc;
    }
    Scratch sc{ws.partial, ws.partial_floats};
    // target network on next_obs (forward #2) on a side stream, concurrently with the online pass
    static const bool fork_env = !(getenv("B2RL_NO_FORK") && getenv("B2RL_NO_FORK")[0] == '1');
    const bool fork_on = fork_env && (cfg.side_streams & 1);
    // First layer over uint8 frames: the online and the target network both read next_obs (dqn_rainbow.py:306-318), so
    // one launch gathers every frame once — next_obs against [W_online | W_target] (256 accumulator columns), obs
    // against W_online — before the target chain forks off (conv_i8.cuh).
    bool shared_first = false;
    {
        const b2rl_layer &l0 = net.enc[0];
        if (tc_enabled() && !l0.noisy &&
            conv_i8_ok(l0, net.obs_u8 != 0, net.normalize != 0, net.obs_low, net.obs_high, bufs.next_obs, 2) &&
            reinterpret_cast<uintptr_t>(bufs.obs) % 4 == 0) {
            const int64_t oe = layer_out_elems(l0);
            ConvI8Job jobs[2] = {
                ConvI8Job{bufs.next_obs, bufs.row_idx, B, 2, {ws.online.enc[0].a, ws.target.enc[0].a}},
                ConvI8Job{bufs.obs, bufs.row_idx, B, 1, {ws.online.enc[0].a + B * oe, nullptr}}};
            const float *Ws[2] = {bufs.actor_params + l0.w_off, bufs.target_params + l0.w_off};
            const float *bs[2] = {bufs.actor_params + l0.b_off, bufs.target_params + l0.b_off};
            if (prepped && ws.prep0)
                rc = launch_conv_fwd_i8(l0, net.normalize != 0, net.obs_low, net.obs_high, Ws, bs, 2, jobs, 2, ws.prep0,
                                        ws.prep0_bytes, s, true);
            else
                rc = launch_conv_fwd_i8(l0, net.normalize != 0, net.obs_low, net.obs_high, Ws, bs, 2, jobs, 2, ws.partial,
                                        ws.partial_floats * sizeof(float), s);
            if (rc == B2RL_OK) shared_first = true;
            else if (rc != 1) return rc;
        }
    }
    ForkJoin *fj = nullptr;
    cudaStream_t st = s;
    if (fork_on) {
        if ((rc = fork_join(&fj)) != B2RL_OK) return rc;
        B2RL_CUDA(cudaEventRecord(fj->fork, s));
        B2RL_CUDA(cudaStreamWaitEvent(fj->side, fj->fork, 0));
        st = fj->side;
    }
    ObsChunk tg_chunk{bufs.next_obs, bufs.row_idx, B};
    Scratch sc_tg{ws.partial_tg, ws.partial_tg_floats};
    if ((rc = forward_pass(net, bufs.target_params, ws.weff_target, noise, &tg_chunk, 1, B, ws.target, sc_tg, st,
                           shared_first, prepped ? ws.prep_split[1] : nullptr, ws.prep_split_floats)) != B2RL_OK)
        return rc;
    if (fork_on) B2RL_CUDA(cudaEventRecord(fj->join, fj->side));
    // online network on [next_obs ; obs]  (forwards #1 and #3 of _dqn_loss share weights and noise)
    ObsChunk on_chunks[2] = {{bufs.next_obs, bufs.row_idx, B}, {bufs.obs, bufs.row_idx, B}};
    if ((rc = forward_pass(net, bufs.actor_params, ws.weff_actor, noise, on_chunks, 2, 2 * B, ws.online, sc, s,
                           shared_first, prepped ? ws.prep_split[0] : nullptr, ws.prep_split_floats)) != B2RL_OK)
        return rc;
    if (fork_on) B2RL_CUDA(cudaStreamWaitEvent(s, fj->join, 0));
    const float *v_on = ws.online.val[net.n_val - 1].a, *adv_on = ws.online.adv[net.n_adv - 1].a;
    const float *v_tg = ws.target.val[net.n_val - 1].a, *adv_tg = ws.target.adv[net.n_adv - 1].a;
    const size_t sm_q = sizeof(float) * ((size_t)A * N + A);
    const size_t sm_t = sizeof(float) * ((size_t)A * N + 3 * (size_t)N) + sizeof(int) * 2 * (size_t)N;
    if ((rc = head_smem_check(sm_t)) != B2RL_OK) return rc;
    // next_actions = actor(next_obs).argmax(1) (dqn_rainbow.py:315): selected inside the target kernel
    const bool fuse_q = A <= 3 * N;
    if (!fuse_q) {
        rainbow_q_kernel<<<(int)B, 128, sm_q, s>>>(v_on, adv_on, bufs.support, A, N, nullptr, ws.a_star, nullptr);
        B2RL_LAUNCH_CHECK();
    }
    const float *fq_v = fuse_q ? v_on : nullptr, *fq_a = fuse_q ? adv_on : nullptr;
    ProjCfg pc{(float)cfg.gamma, (float)cfg.v_min, (float)cfg.v_max, (float)cfg.delta_z};
    static const bool fuse_tail_env = !(getenv("B2RL_NO_TAIL_FUSE") && getenv("B2RL_NO_TAIL_FUSE")[0] == '1');
    if (!cfg.driver_shapes && fuse_q && fuse_tail_env && A * N >= N + 32) {
        if (!prepped) B2RL_CUDA(cudaMemsetAsync(ws.ticket, 0, sizeof(unsigned int), s));
        rainbow_tail_kernel<<<(int)B, 128, sm_t, s>>>(
            v_tg, adv_tg, ws.a_star, bufs.reward, bufs.done, bufs.support, pc, A, N, ws.proj, v_on, adv_on, v_on + B * N,
            adv_on + B * (int64_t)A * N, bufs.action, bufs.weights, cfg.weights_mode, B, cfg.accumulate, bufs.loss_elem,
            ws.online.val[net.n_val - 1].g, ws.online.adv[net.n_adv - 1].g, (float)cfg.prior_eps, bufs.loss_scalar,
            bufs.priorities, ws.ticket);
        B2RL_LAUNCH_CHECK();
    } else {
    if (!cfg.driver_shapes) {
        rainbow_target_kernel<<<(int)B, 128, sm_t, s>>>(v_tg, adv_tg, ws.a_star, bufs.reward, bufs.done, bufs.support,
                                                        pc, A, N, nullptr, ws.proj, fq_v, fq_a);
        B2RL_LAUNCH_CHECK();
    } else {
        rainbow_target_kernel<<<(int)B, 128, sm_t, s>>>(v_tg, adv_tg, ws.a_star, bufs.reward, bufs.done, bufs.support,
                                                        pc, A, N, ws.tdist, nullptr, fq_v, fq_a);
        B2RL_LAUNCH_CHECK();
        q2_cmat_kernel<<<N, 64, sizeof(float) * N, s>>>(bufs.reward, bufs.done, bufs.support, pc, B, N, ws.cmat);
        B2RL_LAUNCH_CHECK();
        q2_project_kernel<<<(int)B, 64, 0, s>>>(ws.tdist, ws.cmat, N, ws.proj);
        B2RL_LAUNCH_CHECK();
    }
    // log-softmax of the taken action on obs rows [B, 2B) + loss + dL/dlogits
    rainbow_loss_kernel<<<(int)B, 128, sizeof(float) * ((size_t)N + 32), s>>>(
        v_on + B * N, adv_on + B * (int64_t)A * N, bufs.action, ws.proj, bufs.weights, cfg.weights_mode, B, A, N,
        cfg.accumulate, bufs.loss_elem, ws.online.val[net.n_val - 1].g, ws.online.adv[net.n_adv - 1].g);
    B2RL_LAUNCH_CHECK();
    rainbow_scalar_kernel<<<1, 256, 0, s>>>(bufs.loss_elem, bufs.weights, cfg.weights_mode, B, (float)cfg.prior_eps,
                                            bufs.loss_scalar, bufs.priorities);
    B2RL_LAUNCH_CHECK();
    }
    if (bufs.proj_dist)
        B2RL_CUDA(cudaMemcpyAsync(bufs.proj_dist, ws.proj, sizeof(float) * B * N, cudaMemcpyDeviceToDevice, s));
    return B2RL_OK;
}

static int check_learn_args(const b2rl_net_desc *net, const b2rl_learn_cfg *cfg, const b2rl_learn_bufs *bufs,
                            LearnWS &ws) {
    B2RL_CHECK_ARG(net && cfg && bufs, "NULL descriptor");
    int rc = validate_net(*net);
    if (rc != B2RL_OK) return rc;
    B2RL_CHECK_ARG(cfg->batch >= 1, "Batch size must be greater than or equal to one.");
    B2RL_CHECK_ARG(bufs->actor_params && bufs->target_params && bufs->grads && bufs->exp_avg && bufs->exp_avg_sq,
                   "NULL parameter/optimizer buffer");
    B2RL_CHECK_ARG(bufs->obs && bufs->next_obs && bufs->action && bufs->reward && bufs->done, "NULL batch buffer");
    B2RL_CHECK_ARG(cfg->weights_mode == 0 || bufs->weights, "PER weights missing");
    carve_learn(*net, cfg->batch, true, bufs->workspace, ws);
    B2RL_CHECK_ARG(bufs->workspace && bufs->workspace_bytes >= ws.bytes, "workspace too small: need %zu bytes, got %zu",
                   ws.bytes, bufs->workspace_bytes);
    return B2RL_OK;
}

}  // namespace b2rl

using namespace b2rl;

extern "C" {

int b2rl_net_workspace_bytes(const b2rl_net_desc *net_host, int64_t rows, int with_backward, size_t *out_host) {
    B2RL_CHECK_ARG(net_host && out_host, "NULL argument");
    int rc = validate_net(*net_host);
    if (rc != B2RL_OK) return rc;
    if (with_backward) {
        LearnWS ws;
        carve_learn(*net_host, rows, true, nullptr, ws);
        *out_host = ws.bytes;
    } else {
        FwdWS ws;
        carve_fwd(*net_host, rows, nullptr, ws);
        *out_host = ws.bytes;
    }
    return B2RL_OK;
}

int b2rl_noise_count(const b2rl_net_desc *net_host, int64_t *out_host) {
    B2RL_CHECK_ARG(net_host && out_host, "NULL argument");
    NoiseTable t;
    *out_host = build_noise_table(*net_host, nullptr, t);
    return B2RL_OK;
}

int b2rl_noise_reset_from_normals(const b2rl_net_desc *net_host, float *eps, const float *normals, void *stream) {
    B2RL_CHECK_ARG(net_host && normals, "NULL argument");
    B2RL_CHECK_ARG(eps || !has_noisy(*net_host), "eps buffer is NULL");
    return noise_reset(*net_host, eps, normals, 0, 0, as_stream(stream));
}

int b2rl_noise_reset_philox(const b2rl_net_desc *net_host, float *eps, uint64_t seed, uint64_t offset, void *stream) {
    B2RL_CHECK_ARG(net_host, "NULL argument");
    B2RL_CHECK_ARG(eps || !has_noisy(*net_host), "eps buffer is NULL");
    return noise_reset(*net_host, eps, nullptr, seed, offset, as_stream(stream));
}

// mode 0: expected Q-values (+ argmax); 1: per-atom distributions; 2: per-atom log-probabilities
static int net_forward(const b2rl_net_desc *net_host, const float *params, const float *eps, int use_noise,
                       const float *support, const void *obs, const int64_t *row_idx, int64_t rows, int mode,
                       float *out, int64_t *argmax_out, void *workspace, size_t workspace_bytes, void *stream) {
    B2RL_CHECK_ARG(net_host && params && obs, "NULL argument");
    int rc = validate_net(*net_host);
    if (rc != B2RL_OK) return rc;
    if (rows <= 0) return B2RL_OK;
    const b2rl_net_desc &net = *net_host;
    B2RL_CHECK_ARG(mode == 0 || net.kind == B2RL_NET_RAINBOW, "per-atom distributions exist for rainbow networks only");
    FwdWS ws;
    carve_fwd(net, rows, workspace, ws);
    B2RL_CHECK_ARG(workspace && workspace_bytes >= ws.bytes, "workspace too small: need %zu bytes, got %zu", ws.bytes,
                   workspace_bytes);
    cudaStream_t s = as_stream(stream);
    const bool noise = use_noise && has_noisy(net);
    if (noise) {
        B2RL_CHECK_ARG(eps, "eps buffer is NULL");
        if ((rc = compose_weights(net, params, eps, ws.weff, s)) != B2RL_OK) return rc;
    }
    Scratch sc{ws.partial, ws.partial_floats};
    ObsChunk chunk{obs, row_idx, rows};
    if ((rc = forward_pass(net, params, ws.weff, noise, &chunk, 1, rows, ws.pass, sc, s)) != B2RL_OK) return rc;
    const int A = net.n_actions, N = net.n_atoms;
    if (net.kind == B2RL_NET_RAINBOW) {
        const size_t sm_q = sizeof(float) * ((size_t)A * N + A);
        if ((rc = head_smem_check(sm_q)) != B2RL_OK) return rc;
        const float *v = ws.pass.val[net.n_val - 1].a, *adv = ws.pass.adv[net.n_adv - 1].a;
        if (mode == 0) {
            B2RL_CHECK_ARG(support, "support is NULL");
            rainbow_q_kernel<<<(int)rows, 128, sm_q, s>>>(v, adv, support, A, N, out, nullptr, argmax_out);
        } else {
            B2RL_CHECK_ARG(out, "output is NULL");
            rainbow_dist_kernel<<<(int)rows, 128, sm_q, s>>>(v, adv, A, N, mode == 2, out);
        }
        B2RL_LAUNCH_CHECK();
        return B2RL_OK;
    }
    q_argmax_kernel<<<(int)((rows + 127) / 128), 128, 0, s>>>(ws.pass.val[net.n_val - 1].a, A, rows, out, argmax_out);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_noise_reset_state(const b2rl_net_desc *net_host, float *eps, uint64_t seed, const b2rl_step_state *state,
                           int which, void *stream) {
    B2RL_CHECK_ARG(net_host && eps && state && (which == 0 || which == 1), "bad arguments");
    return noise_reset(*net_host, eps, nullptr, seed, 0, as_stream(stream), state, which);
}

int b2rl_noise_reset_state_pair(const b2rl_net_desc *net_host, float *eps_actor, float *eps_target, uint64_t seed,
                                const b2rl_step_state *state, void *stream) {
    B2RL_CHECK_ARG(net_host && eps_actor && eps_target && state, "bad arguments");
    NoiseTable t;
    build_noise_table(*net_host, eps_actor, t, false, 0);
    build_noise_table(*net_host, eps_target, t, true, 1);
    if (t.n == 0) return B2RL_OK;
    int64_t maxn = 0;
    for (int i = 0; i < t.n; ++i) {
        const int64_t n = (int64_t)t.s[i].in * t.s[i].out;
        if (n > maxn) maxn = n;
    }
    int bx = (int)((maxn + 255) / 256);
    if (bx > 64) bx = 64;
    noise_reset_kernel<true><<<dim3(bx, t.n), 256, 0, as_stream(stream)>>>(t, nullptr, seed, 0, state, -1);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_net_forward_q(const b2rl_net_desc *net_host, const float *params, const float *eps, int use_noise,
                       const float *support, const void *obs, const int64_t *row_idx, int64_t rows, float *q_out,
                       int64_t *argmax_out, void *workspace, size_t workspace_bytes, void *stream) {
    return net_forward(net_host, params, eps, use_noise, support, obs, row_idx, rows, 0, q_out, argmax_out, workspace,
                       workspace_bytes, stream);
}

int b2rl_net_forward_dist(const b2rl_net_desc *net_host, const float *params, const float *eps, int use_noise,
                          const void *obs, const int64_t *row_idx, int64_t rows, int log_probs, float *dist_out,
                          void *workspace, size_t workspace_bytes, void *stream) {
    return net_forward(net_host, params, eps, use_noise, nullptr, obs, row_idx, rows, log_probs ? 2 : 1, dist_out, nullptr,
                       workspace, workspace_bytes, stream);
}


int b2rl_rainbow_loss(const b2rl_net_desc *net_host, const b2rl_learn_cfg *cfg_host, const b2rl_learn_bufs *bufs_host,
                      void *stream) {
    LearnWS ws;
    int rc = check_learn_args(net_host, cfg_host, bufs_host, ws);
    if (rc != B2RL_OK) return rc;
    B2RL_CHECK_ARG(net_host->kind == B2RL_NET_RAINBOW, "not a rainbow network");
    B2RL_CHECK_ARG(bufs_host->support && bufs_host->loss_elem && bufs_host->loss_scalar, "NULL rainbow buffer");
    return rainbow_loss(*net_host, *cfg_host, *bufs_host, ws, as_stream(stream));
}

int b2rl_rainbow_prep(const b2rl_net_desc *net_host, const b2rl_learn_cfg *cfg_host, const b2rl_learn_bufs *bufs_host,
                      void *stream) {
    LearnWS ws;
    int rc = check_learn_args(net_host, cfg_host, bufs_host, ws);
    if (rc != B2RL_OK) return rc;
    B2RL_CHECK_ARG(net_host->kind == B2RL_NET_RAINBOW, "not a rainbow network");
    return rainbow_prep(*net_host, *cfg_host, *bufs_host, ws, as_stream(stream));
}

int b2rl_rainbow_backward(const b2rl_net_desc *net_host, const b2rl_learn_cfg *cfg_host,
                          const b2rl_learn_bufs *bufs_host, void *stream) {
    LearnWS ws;
    int rc = check_learn_args(net_host, cfg_host, bufs_host, ws);
    if (rc != B2RL_OK) return rc;
    Scratch sc{ws.partial, ws.partial_floats};
    ObsChunk obs{bufs_host->obs, bufs_host->row_idx, cfg_host->batch};
    static const bool fork_env = !(getenv("B2RL_NO_FORK") && getenv("B2RL_NO_FORK")[0] == '1');
    const bool fork_on = fork_env && (cfg_host->side_streams & 2);
    WgradSide side_storage;
    const WgradSide *side = nullptr;
    if (fork_on) {
        ForkJoin *fj = nullptr;
        if ((rc = fork_join(&fj)) != B2RL_OK) return rc;
        side_storage = WgradSide{fj->side_bw, Scratch{ws.partial_bw, ws.partial_bw_floats}, fj->bw};
        side = &side_storage;
    }
    // cfg.reserved_ bit 0: this step's b2rl_rainbow_prep laid out the input-gradient weights as well
    return backward_pass(*net_host, bufs_host->actor_params, ws.weff_actor, cfg_host->use_noise != 0,
                         bufs_host->actor_eps, ws.online, cfg_host->batch, cfg_host->batch, obs, bufs_host->grads,
                         ws.gweff, cfg_host->accumulate, sc, as_stream(stream), side,
                         (cfg_host->reserved_ & 1) ? ws.prep_dsplit : nullptr);
}

int b2rl_optim_step(const b2rl_net_desc *net_host, const b2rl_learn_cfg *cfg_host, const b2rl_learn_bufs *bufs_host,
                    void *stream) {
    LearnWS ws;
    int rc = check_learn_args(net_host, cfg_host, bufs_host, ws);
    if (rc != B2RL_OK) return rc;
    return optim_step(*net_host, *cfg_host, *bufs_host, ws, as_stream(stream));
}

int b2rl_rainbow_learn(const b2rl_net_desc *net_host, const b2rl_learn_cfg *cfg_host, const b2rl_learn_bufs *bufs_host,
                       void *stream) {
    int rc = b2rl_rainbow_loss(net_host, cfg_host, bufs_host, stream);
    if (rc != B2RL_OK) return rc;
    if ((rc = b2rl_rainbow_backward(net_host, cfg_host, bufs_host, stream)) != B2RL_OK) return rc;
    return b2rl_optim_step(net_host, cfg_host, bufs_host, stream);
}

int b2rl_dqn_learn(const b2rl_net_desc *net_host, const b2rl_learn_cfg *cfg_host, const b2rl_learn_bufs *bufs_host,
                   void *stream) {
    LearnWS ws;
    int rc = check_learn_args(net_host, cfg_host, bufs_host, ws);
    if (rc != B2RL_OK) return rc;
    const b2rl_net_desc &net = *net_host;
    const b2rl_learn_cfg &cfg = *cfg_host;
    const b2rl_learn_bufs &bufs = *bufs_host;
    B2RL_CHECK_ARG(net.kind == B2RL_NET_Q, "not a Q network");
    B2RL_CHECK_ARG(bufs.loss_scalar, "NULL loss buffer");
    cudaStream_t s = as_stream(stream);
    const int64_t B = cfg.batch;
    const int A = net.n_actions;
    Scratch sc{ws.partial, ws.partial_floats};
    // q-networks run in eval mode inside DQN.update (plain Linear layers; no noise)
    int64_t row_off = 0;
    if (cfg.double_dqn) {
        ObsChunk ch[2] = {{bufs.next_obs, bufs.row_idx, B}, {bufs.obs, bufs.row_idx, B}};
        if ((rc = forward_pass(net, bufs.actor_params, ws.weff_actor, false, ch, 2, 2 * B, ws.online, sc, s)) != B2RL_OK)
            return rc;
        row_off = B;
    } else {
        ObsChunk ch{bufs.obs, bufs.row_idx, B};
        if ((rc = forward_pass(net, bufs.actor_params, ws.weff_actor, false, &ch, 1, B, ws.online, sc, s)) != B2RL_OK)
            return rc;
    }
    ObsChunk tg{bufs.next_obs, bufs.row_idx, B};
    if ((rc = forward_pass(net, bufs.target_params, ws.weff_target, false, &tg, 1, B, ws.target, sc, s)) != B2RL_OK)
        return rc;
    const float *q_on = ws.online.val[net.n_val - 1].a;
    dqn_loss_kernel<<<1, 256, 0, s>>>(q_on + row_off * A, q_on, ws.target.val[net.n_val - 1].a, bufs.action, bufs.reward,
                                      bufs.done, (float)cfg.gamma, cfg.double_dqn, B, A,
                                      ws.online.val[net.n_val - 1].g, bufs.loss_elem, bufs.loss_scalar);
    B2RL_LAUNCH_CHECK();
    ObsChunk obs{bufs.obs, bufs.row_idx, B};
    if ((rc = backward_pass(net, bufs.actor_params, ws.weff_actor, false, bufs.actor_eps, ws.online, row_off, B, obs,
                            bufs.grads, ws.gweff, 0, sc, s)) != B2RL_OK)
        return rc;
    return optim_step(net, cfg, bufs, ws, s);
}


int b2rl_debug_read(long long *out_host, int n) {
    long long *buf = tc_debug_buffer();
    B2RL_CHECK_ARG(out_host && n >= 1 && n <= 128, "bad debug read");
    B2RL_CHECK_ARG(buf != nullptr, "diagnostics are off (set B2RL_TC_DBG=<cta>)");
    B2RL_CUDA(cudaDeviceSynchronize());
    B2RL_CUDA(cudaMemcpy(out_host, buf, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    return B2RL_OK;
}

int b2rl_encoder_layer_forward(const b2rl_net_desc *net_host, int layer, const float *params, const void *input,
                               const int64_t *row_idx, int64_t rows, float *out, void *workspace,
                               size_t workspace_bytes, int reuse_split, void *stream) {
    B2RL_CHECK_ARG(net_host && params && input && out, "NULL argument");
    B2RL_CHECK_ARG(layer >= 0 && layer < net_host->n_enc, "layer out of range");
    const b2rl_layer &l = net_host->enc[layer];
    B2RL_CHECK_ARG(l.ln == B2RL_LN_NONE && !l.noisy, "profiling hook handles plain conv/linear layers");
    LayerBuf lb;
    lb.a = out;
    Scratch sc{static_cast<float *>(workspace), workspace_bytes / sizeof(float)};
    ObsChunk ch{input, row_idx, rows};
    return layer_forward(*net_host, l, params + l.w_off, params + l.b_off, params,
                         layer == 0 ? nullptr : static_cast<const float *>(input), &ch, 1, rows, lb, sc,
                         as_stream(stream), reuse_split != 0);
}

int b2rl_conv_staged_paths(int mask) {
    const int prev = st_mask();
    if (mask >= 0) st_mask() = mask & 7;
    return prev;
}

int b2rl_encoder_layer_wgrad(const b2rl_net_desc *net_host, int layer, const void *input, const int64_t *row_idx,
                             int64_t rows, const float *g_out, float *grads, void *workspace, size_t workspace_bytes,
                             void *stream) {
    B2RL_CHECK_ARG(net_host && input && g_out && grads, "NULL argument");
    B2RL_CHECK_ARG(layer >= 0 && layer < net_host->n_enc, "layer out of range");
    b2rl_layer l = net_host->enc[layer];
    B2RL_CHECK_ARG(l.ln == B2RL_LN_NONE && !l.noisy, "profiling hook handles plain conv/linear layers");
    l.act = B2RL_ACT_NONE;                                 // g_out is the gradient at the layer's pre-activation output
    LayerBuf lb;
    Scratch sc{static_cast<float *>(workspace), workspace_bytes / sizeof(float)};
    ObsChunk ch{input, row_idx, rows};
    return layer_backward(*net_host, l, nullptr, nullptr, layer == 0 ? nullptr : static_cast<const float *>(input),
                          layer == 0 ? &ch : nullptr, lb, 0, const_cast<float *>(g_out), nullptr, false, grads, nullptr, 0, rows,
                          sc, as_stream(stream));
}

int b2rl_encoder_layer_dgrad(const b2rl_net_desc *net_host, int layer, const float *params, const float *g_out, int64_t rows,
                             float *g_in, void *workspace, size_t workspace_bytes, void *stream) {
    B2RL_CHECK_ARG(net_host && params && g_out && g_in, "NULL argument");
    B2RL_CHECK_ARG(layer >= 1 && layer < net_host->n_enc, "layer out of range (the first layer has no input gradient)");
    b2rl_layer l = net_host->enc[layer];
    B2RL_CHECK_ARG(l.kind == B2RL_LAYER_CONV && l.ln == B2RL_LN_NONE && !l.noisy, "profiling hook handles plain conv layers");
    Scratch sc{static_cast<float *>(workspace), workspace_bytes / sizeof(float)};
    cudaStream_t s = as_stream(stream);
    int rc = launch_conv_dgrad_st(l, g_out, params + l.w_off, g_in, rows, sc.partial, sc.floats, s);
    if (rc == 1) rc = launch_conv_dgrad_tc(l, g_out, params + l.w_off, g_in, rows, sc.partial, sc.floats, s);
    if (rc == 1) { set_error("layer outside the tensor-core input-gradient kernels"); return B2RL_EUNSUPPORTED; }
    return rc;
}

}  // extern "C"

#include "ddpg.cuh"
#include "maddpg.cuh"
