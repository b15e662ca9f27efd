// ddpg.cuh — DDPG / TD3 learn() on the HBM replay (BASELINE configs[2]); included by nn.cu.
//
// Replaces agilerl/algorithms/ddpg.py:422-500 and td3.py:459-551 with the networks of
// networks/actors.py:78-210 (DeterministicActor: MLP encoder -> LayerNorm MLP head, Tanh) and
// networks/q_networks.py:302-443 (ContinuousQNetwork: MLP encoder -> cat(latent, action) -> LayerNorm MLP head -> 1).
//
// Every network here is a chain of tiny (Linear -> LayerNorm -> activation) layers over 17..64-wide rows: the
// whole step is launch- and latency-bound, so each chain runs as ONE fused launch forward (head_fwd_kernel: a
// 4-row tile walks every layer with activations in shared memory) and two backward (head_bwd_kernel: dZ / dX
// chain; head_wgrad_kernel: every dW / db / LayerNorm-affine gradient) — the kernels of the Rainbow head, driven
// with the chain's own layer table.  What is specific to this learner are five elementwise kernels: the
// cat(latent, action), the target action (in-place policy noise on the batch's action tensor, clip, add, clamp —
// the reference's quirk order), the TD target + MSE + dL/dq seeds, the -mean(q) actor loss, and a column slice.
// A learn call is 23 launches for TD3 on a policy step (reference: several hundred ATen launches).
#pragma once

namespace b2rl {

struct Chain {                     // one MLP chain of a network description
    const b2rl_layer *layers;
    int n;
    int in_features;
};
static inline Chain enc_chain(const b2rl_net_desc &d) { return Chain{d.enc, d.n_enc, d.enc[0].in_c}; }
static inline Chain val_chain(const b2rl_net_desc &d) { return Chain{d.val, d.n_val, d.val[0].in_c}; }

static int chain_check(const Chain &c) {
    B2RL_CHECK_ARG(c.n >= 1 && c.n <= kHeadMaxLayers, "chain depth outside the fused kernels' layer table");
    for (int i = 0; i < c.n; ++i) {
        B2RL_CHECK_ARG(c.layers[i].kind == B2RL_LAYER_LINEAR && !c.layers[i].noisy, "DDPG/TD3 chains are plain linear layers");
        B2RL_CHECK_ARG(c.layers[i].in_c <= kHeadWgMaxIn && c.layers[i].out_c <= kHeadWgMaxIn, "layer wider than the fused kernels");
    }
    return B2RL_OK;
}
static int chain_maxdim(const Chain &c) {
    int m = c.in_features;
    for (int i = 0; i < c.n; ++i) { m = c.layers[i].in_c > m ? c.layers[i].in_c : m; m = c.layers[i].out_c > m ? c.layers[i].out_c : m; }
    return m;
}
static int head_kernels_ready() {
    static bool attr_set = false;
    if (!attr_set) {
        B2RL_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        B2RL_CUDA(cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        B2RL_CUDA(cudaFuncSetAttribute(head_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kHeadWgSmemFloats * sizeof(float))));
        attr_set = true;
    }
    return B2RL_OK;
}

// y = chain(x): one launch.  bufs[i].a (and z / stats / pre where the layer has them) receive every layer's output.
static int chain_forward(const Chain &c, const float *params, const float *x, int64_t rows, const LayerBuf *bufs,
                         cudaStream_t s) {
    int rc = head_kernels_ready();
    if (rc != B2RL_OK) return rc;
    HeadDesc hd;
    hd.n_val = c.n; hd.n_adv = 0; hd.latent = c.in_features; hd.maxdim = chain_maxdim(c);
    for (int i = 0; i < c.n; ++i) {
        const b2rl_layer &l = c.layers[i];
        HeadLayer &h = hd.l[i];
        h.w = params + l.w_off; h.b = params + l.b_off;
        h.lnw = l.ln == B2RL_LN_AFFINE ? params + l.lnw_off : nullptr;
        h.lnb = l.ln == B2RL_LN_AFFINE ? params + l.lnb_off : nullptr;
        h.a = bufs[i].a; h.z = bufs[i].z; h.pre = bufs[i].pre; h.stats = bufs[i].stats;
        h.in = l.in_c; h.out = l.out_c; h.ln = l.ln; h.act = l.act;
    }
    const size_t smem = sizeof(float) * ((size_t)2 * kHeadRows * head_pitch(hd.maxdim) + (kHeadThreads / 32) * kHeadRows);
    B2RL_CHECK_ARG(smem <= 160 * 1024, "chain too wide for the fused forward kernel");
    head_fwd_kernel<<<(int)((rows + kHeadRows - 1) / kHeadRows), kHeadThreads, smem, s>>>(hd, x, rows);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

// Backward of one chain for B rows.  bufs[n-1].g holds dL/d(chain output) on entry; g_in (nullable scratch is NOT
// allowed: [B, in_features]) receives dL/dx.  grads == NULL: only the dX chain runs (gradient w.r.t. the input
// through frozen weights — the actor loss through critic_1).  lnpart: scratch for LayerNorm-affine partials.
static int chain_backward(const Chain &c, const float *params, const float *x, int64_t B, const LayerBuf *bufs, float *g_in,
                          float *grads, float *lnpart, size_t lnpart_floats, cudaStream_t s) {
    int rc = head_kernels_ready();
    if (rc != B2RL_OK) return rc;
    HeadBwdDesc hd;
    hd.n_val = c.n; hd.n_adv = 0; hd.latent = c.in_features; hd.g_latent = g_in; hd.latent_a = nullptr; hd.latent_act = B2RL_ACT_NONE; hd.accumulate = 0; hd.n_ln = 0;
    hd.maxdim = chain_maxdim(c);
    const int n_tiles = (int)((B + kHeadRows - 1) / kHeadRows);
    int ctas = 0;
    size_t part_used = 0;
    for (int i = 0; i < c.n; ++i) {
        const b2rl_layer &l = c.layers[i];
        HeadBwdLayer &h = hd.l[i];
        h.w = params + l.w_off;
        h.lnw = l.ln == B2RL_LN_AFFINE ? params + l.lnw_off : nullptr;
        h.a = bufs[i].a; h.z = bufs[i].z; h.pre = bufs[i].pre; h.stats = bufs[i].stats;
        h.x = i == 0 ? x : bufs[i - 1].a;
        h.g = bufs[i].g;
        h.dw = grads ? grads + l.w_off : nullptr; h.db = grads ? grads + l.b_off : nullptr;
        h.dlnw = h.dlnb = h.lnpart = nullptr;
        if (grads && l.ln == B2RL_LN_AFFINE) {
            h.dlnw = grads + l.lnw_off; h.dlnb = grads + l.lnb_off;
            h.lnpart = lnpart + part_used;
            part_used += (size_t)n_tiles * 2 * l.out_c;
            hd.ln_layer[hd.n_ln++] = i;
        }
        h.in = l.in_c; h.out = l.out_c; h.ln = l.ln; h.act = l.act; h.acc_w = 0;
        hd.wg_start[i] = ctas;
        ctas += (l.out_c + kHeadThreads / 32 - 1) / (kHeadThreads / 32);
    }
    hd.wg_start[c.n] = ctas;
    B2RL_CHECK_ARG(part_used <= lnpart_floats, "LayerNorm scratch too small");
    const size_t smem = sizeof(float) * ((size_t)(3 + kHeadThreads / 32) * kHeadRows * head_pitch(hd.maxdim) +
                                         (size_t)kHeadRows * hd.latent + (kHeadThreads / 32) * kHeadRows);
    B2RL_CHECK_ARG(smem <= 160 * 1024, "chain too wide for the fused backward kernel");
    head_bwd_kernel<<<n_tiles, kHeadThreads, smem, s>>>(hd, B);
    B2RL_LAUNCH_CHECK();
    if (grads) {
        head_wgrad_kernel<<<ctas + hd.n_ln, kHeadThreads, kHeadWgSmemFloats * sizeof(float), s>>>(hd, B, n_tiles);
        B2RL_LAUNCH_CHECK();
    }
    return B2RL_OK;
}

// ---- the learner's own elementwise kernels ---------------------------------------------------------------------
// cat[r] = [latent[r] | act[r]]   (q_networks.py:424-425); act == NULL leaves the action columns as they are
__global__ void ddpg_concat_kernel(const float *__restrict__ latent, int L, const float *__restrict__ act, int A, int64_t B,
                                   float *__restrict__ out) {
    const int W = L + A;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < B * W; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / W;
        const int c = (int)(e - r * W);
        if (c < L) out[e] = latent[r * L + c];
        else if (act) out[e] = act[r * A + (c - L)];
    }
}
// dst[r][0:n] = src[r][col0 : col0+n]
__global__ void ddpg_slice_kernel(const float *__restrict__ src, int ld, int col0, int n, int64_t B, float *__restrict__ dst) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < B * n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / n;
        dst[e] = src[r * ld + col0 + (int)(e - r * n)];
    }
}
// td3.py:495-503 / ddpg.py:452-462: noise = actions.data.normal_(0, policy_noise)  — IN PLACE on the batch's action
// tensor (kept: the caller's tensor holds the noise afterwards) — clamp(+-noise_clip); next = clamp(actor_target(next_obs)
// + noise, low, high), written straight into the action columns of the target critics' cat buffer.
__global__ void ddpg_target_action_kernel(const float *__restrict__ a_t, float *__restrict__ action_inout,
                                          const float *__restrict__ injected, uint64_t seed, uint64_t offset,
                                          float policy_noise, float noise_clip, const float *__restrict__ low,
                                          const float *__restrict__ high, int A, int L, int64_t B, float *__restrict__ cat,
                                          float *__restrict__ cat2) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < B * A; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / A;
        const int c = (int)(e - r * A);
        float nz = injected ? injected[e] : __fmul_rn(philox_normal(seed, offset + (uint64_t)e, 0x54443341ull /* "TD3A" */), policy_noise);
        action_inout[e] = nz;
        nz = fminf(fmaxf(nz, -noise_clip), noise_clip);
        const float v = fminf(fmaxf(__fadd_rn(a_t[e], nz), low[c]), high[c]);
        cat[r * (L + A) + L + c] = v;
        if (cat2) cat2[r * (L + A) + L + c] = v;          // both target critics see the same next action
    }
}
// y = r + (1-d) * gamma * min(q1', q2');  loss = mse(q1, y) (+ mse(q2, y));  seeds dL/dq_i = 2 (q_i - y) / B.
// One CTA, fixed-order sums (B is a few hundred).
__global__ void ddpg_td_loss_kernel(const float *__restrict__ q1, const float *__restrict__ q2, const float *__restrict__ qn1,
                                    const float *__restrict__ qn2, const float *__restrict__ reward,
                                    const float *__restrict__ done, float gamma, int64_t B, float *__restrict__ g1,
                                    float *__restrict__ g2, float *__restrict__ loss) {
    __shared__ float red[32];
    float s1 = 0.f, s2 = 0.f;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
        const float qn = qn2 ? fminf(qn1[i], qn2[i]) : qn1[i];
        const float y = __fadd_rn(reward[i], __fmul_rn(__fmul_rn(__fsub_rn(1.0f, done[i]), gamma), qn));
        const float d1 = q1[i] - y;
        s1 += d1 * d1;
        g1[i] = 2.0f * d1 / (float)B;
        if (q2) {
            const float d2 = q2[i] - y;
            s2 += d2 * d2;
            g2[i] = 2.0f * d2 / (float)B;
        }
    }
    s1 = block_reduce_sum(s1, red);
    s2 = block_reduce_sum(s2, red);
    if (threadIdx.x == 0) *loss = s1 / (float)B + (q2 ? s2 / (float)B : 0.f);
}
// actor_loss = -mean(q); seed dL/dq = -1/B
__global__ void ddpg_actor_loss_kernel(const float *__restrict__ q, int64_t B, float *__restrict__ g, float *__restrict__ loss) {
    __shared__ float red[32];
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) { s += q[i]; g[i] = -1.0f / (float)B; }
    s = block_reduce_sum(s, red);
    if (threadIdx.x == 0) *loss = -(s / (float)B);
}

struct DdpgWS {
    // critic passes: [0,1] critics on (obs, action) with gradients; [2,3] target critics on (next_obs, a'); [4] critic_1 on
    // (obs, actor(obs)) for the actor loss
    LayerBuf c_enc[5][B2RL_MAX_ENC], c_head[5][B2RL_MAX_HEAD];
    float *cat[5];                   // [B, latent + act]
    float *g_cat[5];                 // dL/dcat of passes 0, 1, 4
    LayerBuf a_enc[2][B2RL_MAX_ENC], a_head[2][B2RL_MAX_HEAD];   // [0] actor(obs) with gradients, [1] actor_target(next_obs)
    float *g_lat_actor;              // dL/dlatent of the actor head
    float *g_obs;                    // dL/dobs scratch (unused result of the first chains)
    float *g_act;                    // [B, act] dQ/da
    float *lnpart;
    size_t lnpart_floats;
    float *norm_partials;
    size_t bytes;
};
static void carve_ddpg(const b2rl_net_desc &actor, const b2rl_net_desc &critic, int64_t B, void *base, DdpgWS &ws) {
    Bump b(base);
    const int L = critic.enc[critic.n_enc - 1].out_c, A = critic.val[0].in_c - L;
    for (int p = 0; p < 5; ++p) {
        const bool grad = p == 0 || p == 1 || p == 4;
        carve_layers(b, critic.enc, critic.n_enc, ws.c_enc[p], B, grad ? B : 0);
        carve_layers(b, critic.val, critic.n_val, ws.c_head[p], B, grad ? B : 0);
        ws.cat[p] = b.take<float>(B * (L + A));
        ws.g_cat[p] = grad ? b.take<float>(B * (L + A)) : nullptr;
    }
    for (int p = 0; p < 2; ++p) {
        carve_layers(b, actor.enc, actor.n_enc, ws.a_enc[p], B, p == 0 ? B : 0);
        carve_layers(b, actor.val, actor.n_val, ws.a_head[p], B, p == 0 ? B : 0);
    }
    ws.g_lat_actor = b.take<float>(B * actor.val[0].in_c);
    int obs_w = critic.enc[0].in_c > actor.enc[0].in_c ? critic.enc[0].in_c : actor.enc[0].in_c;
    ws.g_obs = b.take<float>(B * obs_w);
    ws.g_act = b.take<float>(B * A);
    const int64_t tiles = (B + kHeadRows - 1) / kHeadRows;
    ws.lnpart_floats = (size_t)tiles * 2 * kHeadWgMaxIn * 4;
    ws.lnpart = b.take<float>(ws.lnpart_floats);
    ws.norm_partials = b.take<float>(kNormBlocks);
    ws.bytes = b.off + 256;
}

// torch.optim.Adam (no clipping) on one flat parameter buffer; tgt != NULL also applies the Polyak update
static int ddpg_adam(float *p, float *g, float *m, float *v, float *tgt, int64_t n, double lr, double beta1, double beta2,
                     double eps, double bc1, double bc2, double tau, const float *partials, cudaStream_t s) {
    AdamCfg c;
    c.clip = 0; c.max_norm = 0.f;
    c.w1 = (float)(1.0 - beta1); c.beta2 = (float)beta2; c.w2 = (float)(1.0 - beta2);
    c.neg_step = (float)(-(lr / bc1)); c.bc2_sqrt = (float)sqrt(bc2); c.eps = (float)eps;
    c.tau = (float)tau; c.one_minus_tau = (float)(1.0 - tau);
    int blocks = (int)((n + 255) / 256);
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    adam_polyak_kernel<<<blocks, 256, 0, s>>>(p, g, m, v, tgt, n, partials, kNormBlocks, c, nullptr);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}
__global__ void ddpg_polyak_kernel(const float *__restrict__ p, float *__restrict__ tgt, int64_t n, float tau, float omt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        tgt[i] = __fadd_rn(__fmul_rn(tau, p[i]), __fmul_rn(omt, tgt[i]));
}

static inline int ew_blocks(int64_t n) { int b = (int)((n + 255) / 256); return b < 1 ? 1 : (b > 1184 ? 1184 : b); }

// critic forward: enc(obs) -> cat(latent, action) -> head.  act_in_cat: the action columns of ws.cat[p] were already
// written by ddpg_target_action_kernel (target pass); else they are copied from `action`.
static int critic_forward(const b2rl_net_desc &critic, const float *params, const float *obs, const float *action,
                          int64_t B, DdpgWS &ws, int p, cudaStream_t s) {
    const Chain ce = enc_chain(critic), ch = val_chain(critic);
    const int L = critic.enc[critic.n_enc - 1].out_c, A = ch.in_features - L;
    int rc = chain_forward(ce, params, obs, B, ws.c_enc[p], s);
    if (rc != B2RL_OK) return rc;
    const float *lat = ws.c_enc[p][critic.n_enc - 1].a;
    ddpg_concat_kernel<<<ew_blocks(B * (L + A)), 256, 0, s>>>(lat, L, action, A, B, ws.cat[p]);   // action == NULL: in place
    B2RL_LAUNCH_CHECK();
    return chain_forward(ch, params, ws.cat[p], B, ws.c_head[p], s);
}

}  // namespace b2rl

using namespace b2rl;

extern "C" {

int b2rl_ddpg_workspace_bytes(const b2rl_net_desc *actor_host, const b2rl_net_desc *critic_host, int64_t batch, size_t *out_host) {
    B2RL_CHECK_ARG(actor_host && critic_host && out_host && batch >= 1, "bad arguments");
    DdpgWS ws;
    carve_ddpg(*actor_host, *critic_host, batch, nullptr, ws);
    *out_host = ws.bytes;
    return B2RL_OK;
}

struct ActorWS { LayerBuf enc[B2RL_MAX_ENC], head[B2RL_MAX_HEAD]; size_t bytes; };
static void carve_actor(const b2rl_net_desc &actor, int64_t rows, void *base, ActorWS &ws) {
    Bump b(base);
    carve_layers(b, actor.enc, actor.n_enc, ws.enc, rows, 0);
    carve_layers(b, actor.val, actor.n_val, ws.head, rows, 0);
    ws.bytes = b.off + 256;
}

int b2rl_actor_workspace_bytes(const b2rl_net_desc *actor_host, int64_t rows, size_t *out_host) {
    B2RL_CHECK_ARG(actor_host && out_host && rows >= 1, "bad arguments");
    ActorWS ws;
    carve_actor(*actor_host, rows, nullptr, ws);
    *out_host = ws.bytes;
    return B2RL_OK;
}

// DeterministicActor.forward (actors.py:188-210): out [rows, act_dim] = tanh-head(encoder(obs)).
int b2rl_actor_forward(const b2rl_net_desc *actor_host, const float *params, const float *obs, int64_t rows, float *out,
                       void *workspace, size_t workspace_bytes, void *stream) {
    B2RL_CHECK_ARG(actor_host && params && obs && out, "NULL argument");
    if (rows <= 0) return B2RL_OK;
    const b2rl_net_desc &actor = *actor_host;
    ActorWS ws;
    carve_actor(actor, rows, workspace, ws);
    B2RL_CHECK_ARG(workspace && workspace_bytes >= ws.bytes, "workspace too small: need %zu bytes, got %zu", ws.bytes, workspace_bytes);
    cudaStream_t s = as_stream(stream);
    int rc;
    if ((rc = chain_check(enc_chain(actor))) != B2RL_OK || (rc = chain_check(val_chain(actor))) != B2RL_OK) return rc;
    if ((rc = chain_forward(enc_chain(actor), params, obs, rows, ws.enc, s)) != B2RL_OK) return rc;
    if ((rc = chain_forward(val_chain(actor), params, ws.enc[actor.n_enc - 1].a, rows, ws.head, s)) != B2RL_OK) return rc;
    const int A = actor.val[actor.n_val - 1].out_c;
    B2RL_CUDA(cudaMemcpyAsync(out, ws.head[actor.n_val - 1].a, sizeof(float) * rows * A, cudaMemcpyDeviceToDevice, s));
    return B2RL_OK;
}

int b2rl_ddpg_learn(const b2rl_net_desc *actor_host, const b2rl_net_desc *critic_host, const b2rl_ddpg_cfg *cfg_host,
                    const b2rl_ddpg_bufs *bufs_host, void *stream) {
    B2RL_CHECK_ARG(actor_host && critic_host && cfg_host && bufs_host, "NULL descriptor");
    const b2rl_net_desc &actor = *actor_host, &critic = *critic_host;
    const b2rl_ddpg_cfg &cfg = *cfg_host;
    const b2rl_ddpg_bufs &bf = *bufs_host;
    const int64_t B = cfg.batch;
    const int n_c = cfg.twin ? 2 : 1;
    B2RL_CHECK_ARG(B >= 1, "Batch size must be greater than or equal to one.");
    B2RL_CHECK_ARG(bf.obs && bf.next_obs && bf.action && bf.reward && bf.done && bf.action_low && bf.action_high, "NULL batch buffer");
    B2RL_CHECK_ARG(bf.actor && bf.actor_target && bf.critic[0] && bf.critic_target[0] && bf.critic_loss, "NULL network buffer");
    int rc;
    const Chain ae = enc_chain(actor), ah = val_chain(actor), ce = enc_chain(critic), ch = val_chain(critic);
    if ((rc = chain_check(ae)) != B2RL_OK || (rc = chain_check(ah)) != B2RL_OK || (rc = chain_check(ce)) != B2RL_OK ||
        (rc = chain_check(ch)) != B2RL_OK)
        return rc;
    const int L = critic.enc[critic.n_enc - 1].out_c, A = ch.in_features - L;
    B2RL_CHECK_ARG(A >= 1 && actor.val[actor.n_val - 1].out_c == A && ah.in_features == actor.enc[actor.n_enc - 1].out_c,
                   "actor / critic shapes do not fit together");
    B2RL_CHECK_ARG(critic.val[critic.n_val - 1].out_c == 1, "critic head must end in one value");
    DdpgWS ws;
    carve_ddpg(actor, critic, B, bf.workspace, ws);
    B2RL_CHECK_ARG(bf.workspace && bf.workspace_bytes >= ws.bytes, "workspace too small: need %zu bytes, got %zu", ws.bytes,
                   bf.workspace_bytes);
    cudaStream_t s = as_stream(stream);

    // (1) Q(obs, action) of every critic — BEFORE the batch's action tensor is overwritten with noise
    for (int i = 0; i < n_c; ++i)
        if ((rc = critic_forward(critic, bf.critic[i], bf.obs, bf.action, B, ws, i, s)) != B2RL_OK) return rc;
    // (2) target action, target Q
    if ((rc = chain_forward(ae, bf.actor_target, bf.next_obs, B, ws.a_enc[1], s)) != B2RL_OK) return rc;
    if ((rc = chain_forward(ah, bf.actor_target, ws.a_enc[1][actor.n_enc - 1].a, B, ws.a_head[1], s)) != B2RL_OK) return rc;
    ddpg_target_action_kernel<<<ew_blocks(B * A), 256, 0, s>>>(ws.a_head[1][actor.n_val - 1].a, bf.action, bf.noise, cfg.noise_seed,
                                                              cfg.noise_offset, (float)cfg.policy_noise, (float)cfg.noise_clip,
                                                              bf.action_low, bf.action_high, A, L, B, ws.cat[2],
                                                              n_c > 1 ? ws.cat[3] : nullptr);
    B2RL_LAUNCH_CHECK();
    for (int i = 0; i < n_c; ++i)
        if ((rc = critic_forward(critic, bf.critic_target[i], bf.next_obs, nullptr, B, ws, 2 + i, s)) != B2RL_OK) return rc;
    // (3) TD target, MSE, dL/dq seeds
    const float *q1 = ws.c_head[0][critic.n_val - 1].a, *q2 = n_c > 1 ? ws.c_head[1][critic.n_val - 1].a : nullptr;
    const float *qn1 = ws.c_head[2][critic.n_val - 1].a, *qn2 = n_c > 1 ? ws.c_head[3][critic.n_val - 1].a : nullptr;
    ddpg_td_loss_kernel<<<1, 512, 0, s>>>(q1, q2, qn1, qn2, bf.reward, bf.done, (float)cfg.gamma, B,
                                          ws.c_head[0][critic.n_val - 1].g, n_c > 1 ? ws.c_head[1][critic.n_val - 1].g : nullptr,
                                          bf.critic_loss);
    B2RL_LAUNCH_CHECK();
    // (4) critic backward + Adam (+ Polyak on policy steps: the critics do not change again before their soft update)
    for (int i = 0; i < n_c; ++i) {
        if ((rc = chain_backward(ch, bf.critic[i], ws.cat[i], B, ws.c_head[i], ws.g_cat[i], bf.critic_grads[i], ws.lnpart,
                                 ws.lnpart_floats, s)) != B2RL_OK)
            return rc;
        ddpg_slice_kernel<<<ew_blocks(B * L), 256, 0, s>>>(ws.g_cat[i], L + A, 0, L, B, ws.c_enc[i][critic.n_enc - 1].g);
        B2RL_LAUNCH_CHECK();
        if ((rc = chain_backward(ce, bf.critic[i], bf.obs, B, ws.c_enc[i], ws.g_obs, bf.critic_grads[i], ws.lnpart,
                                 ws.lnpart_floats, s)) != B2RL_OK)
            return rc;
        if ((rc = ddpg_adam(bf.critic[i], bf.critic_grads[i], bf.critic_m[i], bf.critic_v[i],
                            cfg.policy_update ? bf.critic_target[i] : nullptr, critic.n_params, cfg.lr_critic, cfg.beta1,
                            cfg.beta2, cfg.adam_eps, cfg.bc1_critic, cfg.bc2_critic, cfg.tau, ws.norm_partials, s)) != B2RL_OK)
            return rc;
    }
    if (!cfg.policy_update) return B2RL_OK;
    // (5) actor step through the UPDATED critic_1: -mean Q(obs, actor(obs))
    B2RL_CHECK_ARG(bf.actor_grads && bf.actor_m && bf.actor_v && bf.actor_loss, "NULL actor optimiser buffer");
    if ((rc = chain_forward(ae, bf.actor, bf.obs, B, ws.a_enc[0], s)) != B2RL_OK) return rc;
    if ((rc = chain_forward(ah, bf.actor, ws.a_enc[0][actor.n_enc - 1].a, B, ws.a_head[0], s)) != B2RL_OK) return rc;
    if ((rc = critic_forward(critic, bf.critic[0], bf.obs, ws.a_head[0][actor.n_val - 1].a, B, ws, 4, s)) != B2RL_OK) return rc;
    ddpg_actor_loss_kernel<<<1, 512, 0, s>>>(ws.c_head[4][critic.n_val - 1].a, B, ws.c_head[4][critic.n_val - 1].g, bf.actor_loss);
    B2RL_LAUNCH_CHECK();
    if ((rc = chain_backward(ch, bf.critic[0], ws.cat[4], B, ws.c_head[4], ws.g_cat[4], nullptr, ws.lnpart, ws.lnpart_floats,
                             s)) != B2RL_OK)
        return rc;
    ddpg_slice_kernel<<<ew_blocks(B * A), 256, 0, s>>>(ws.g_cat[4], L + A, L, A, B, ws.a_head[0][actor.n_val - 1].g);
    B2RL_LAUNCH_CHECK();
    // dL/dlatent of the head lands directly in the encoder's output-gradient buffer
    if ((rc = chain_backward(ah, bf.actor, ws.a_enc[0][actor.n_enc - 1].a, B, ws.a_head[0], ws.a_enc[0][actor.n_enc - 1].g,
                             bf.actor_grads, ws.lnpart, ws.lnpart_floats, s)) != B2RL_OK)
        return rc;
    if ((rc = chain_backward(ae, bf.actor, bf.obs, B, ws.a_enc[0], ws.g_obs, bf.actor_grads, ws.lnpart, ws.lnpart_floats, s)) != B2RL_OK)
        return rc;
    return ddpg_adam(bf.actor, bf.actor_grads, bf.actor_m, bf.actor_v, bf.actor_target, actor.n_params, cfg.lr_actor, cfg.beta1,
                     cfg.beta2, cfg.adam_eps, cfg.bc1_actor, cfg.bc2_actor, cfg.tau, ws.norm_partials, s);
}

}  // extern "C"
