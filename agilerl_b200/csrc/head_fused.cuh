// head_fused.cuh — the whole MLP head (value chain + advantage chain) forward in ONE launch.
//
// The head of RainbowQNetwork / QNetwork is a handful of tiny (Noisy)Linear -> LayerNorm -> act
// layers (agilerl/networks/custom_modules.py:127-162, utils/evolvable_networks.py:527-644):
// 27k MAC per row for the north-star net.  As separate GEMM launches each of them costs a kernel's
// ramp-up/tail (5-10 us) for microseconds of math; here one CTA takes a tile of 4 rows through
// every layer of both chains, activations staying in shared memory, weights streamed from L2
// (they are a few hundred KB at most and shared by all CTAs).  It writes exactly the buffers the
// unfused path writes (z, stats, pre, a per layer), so backward and the loss kernels are unchanged.
#pragma once
#include "common.cuh"

namespace b2rl {

constexpr int kHeadRows = 4;                   // rows per CTA (more CTAs, shorter per-thread chains)
constexpr int kHeadLanes = 64;                 // output lanes per row
constexpr int kHeadThreads = 256;
constexpr int kHeadMaxLayers = 2 * B2RL_MAX_HEAD;

struct HeadLayer {
    const float *w, *b;         // effective weights [out, in], bias [out]
    const float *lnw, *lnb;     // LayerNorm affine (nullable)
    float *a, *z, *pre, *stats; // outputs (z/stats only when ln, pre only for GELU)
    int in, out, ln, act;
};
struct HeadDesc {
    HeadLayer l[kHeadMaxLayers];
    int n_val, n_adv;           // layers l[0..n_val) = value chain, l[n_val..n_val+n_adv) = advantage chain
    int latent;
    int maxdim;                 // largest feature width incl. latent (smem row pitch = maxdim + 1)
};

// shared-memory row pitch: rows land 8 banks apart, 16-byte aligned
__host__ __device__ inline int head_pitch(int maxdim) { return ((maxdim + 31) & ~31) + 8; }

// sum over the 8 threads of a warp that share (tid & 3), then over the CTA's 8 warps in fixed order
__device__ __forceinline__ float head_row_sum(float v, float *red, int tid) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    __syncthreads();                                   // red may still be read from the previous call
    if ((tid & 31) < kHeadRows) red[(tid >> 5) * kHeadRows + (tid & 3)] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kHeadThreads / 32; ++w) t += red[w * kHeadRows + (tid & 3)];
    return t;
}

__global__ void __launch_bounds__(kHeadThreads) head_fwd_kernel(const HeadDesc hd, const float *__restrict__ latent,
                                                                int64_t rows) {
    extern __shared__ __align__(16) float hsm[];
    const int pitch = head_pitch(hd.maxdim);
    float *xb = hsm;                                  // [rows][pitch] layer input
    float *yb = hsm + kHeadRows * pitch;              // [rows][pitch] layer output (pre-LN)
    float *red = yb + kHeadRows * pitch;              // [warps][rows]
    const int tid = threadIdx.x;
    const int r = tid & (kHeadRows - 1), og = tid / kHeadRows;   // row within the tile / k-slice, output lane
    const int64_t row0 = (int64_t)blockIdx.x * kHeadRows;
    const int64_t row = row0 + r;
    const bool rok = row < rows;

    // gridDim.y == 2: the value chain and the advantage chain of a row tile run in different CTAs (the chains are
    // independent and each is a sequence of dependent weight-streaming rounds: half the serial length per CTA)
    for (int chain = blockIdx.y; chain < 2; chain += gridDim.y) {
        const int l0 = chain == 0 ? 0 : hd.n_val, l1 = chain == 0 ? hd.n_val : hd.n_val + hd.n_adv;
        if (l0 == l1) continue;
        __syncthreads();
        for (int i = og; i < hd.latent; i += kHeadLanes) xb[r * pitch + i] = rok ? latent[row * hd.latent + i] : 0.f;
        __syncthreads();
        for (int li = l0; li < l1; ++li) {
            const HeadLayer L = hd.l[li];
            // ---- linear: y[q][o] = b[o] + sum_i x[q][i] * W[o][i]
            // thread (og, ks = r): output lane og, quarter ks of the reduction (16-byte interleave, so the
            // four ks-threads of an output read one contiguous 64 B piece of its weight row); all rows of
            // the tile in registers share each weight load; xor-shuffle over the ks bits finishes the sum.
            for (int o0 = 0; o0 < L.out; o0 += kHeadLanes) {
                const int o = o0 + og;
                float acc[kHeadRows];
#pragma unroll
                for (int q = 0; q < kHeadRows; ++q) acc[q] = 0.f;
                if (o < L.out) {
                    const float *wr = L.w + (int64_t)o * L.in;
                    if ((L.in & 15) == 0 && (reinterpret_cast<uintptr_t>(L.w) & 15) == 0) {
#pragma unroll 4
                        for (int i = r * 4; i < L.in; i += 16) {
                            const float4 w = __ldg(reinterpret_cast<const float4 *>(wr + i));
#pragma unroll
                            for (int q = 0; q < kHeadRows; ++q) {
                                const float4 x = *reinterpret_cast<const float4 *>(xb + q * pitch + i);
                                acc[q] = fmaf(x.x, w.x, acc[q]);
                                acc[q] = fmaf(x.y, w.y, acc[q]);
                                acc[q] = fmaf(x.z, w.z, acc[q]);
                                acc[q] = fmaf(x.w, w.w, acc[q]);
                            }
                        }
                    } else {
                        for (int i = r; i < L.in; i += kHeadRows) {
                            const float w = __ldg(wr + i);
#pragma unroll
                            for (int q = 0; q < kHeadRows; ++q) acc[q] = fmaf(xb[q * pitch + i], w, acc[q]);
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < kHeadRows; ++q) {
                    acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 1);
                    acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], 2);
                }
                if (o < L.out) {
                    float v = acc[0];
#pragma unroll
                    for (int q = 1; q < kHeadRows; ++q) v = (r == q) ? acc[q] : v;
                    yb[r * pitch + o] = v + (L.b ? L.b[o] : 0.f);
                }
            }
            __syncthreads();
            float mean = 0.f, rstd = 1.f;
            if (L.ln != B2RL_LN_NONE) {
                // two-pass row statistics, fixed-order combine
                float sacc = 0.f;
                for (int o = og; o < L.out; o += kHeadLanes) sacc += yb[r * pitch + o];
                mean = head_row_sum(sacc, red, tid) / (float)L.out;
                float v = 0.f;
                for (int o = og; o < L.out; o += kHeadLanes) { const float d = yb[r * pitch + o] - mean; v += d * d; }
                rstd = 1.0f / sqrtf(head_row_sum(v, red, tid) / (float)L.out + 1e-5f);
                if (og == 0 && rok && L.stats) { L.stats[row * 2] = mean; L.stats[row * 2 + 1] = rstd; }
            }
            // ---- normalise / activate, write global + next layer's input
            for (int o = og; o < L.out; o += kHeadLanes) {
                float y = yb[r * pitch + o];
                if (L.ln != B2RL_LN_NONE) {
                    if (rok && L.z) L.z[row * L.out + o] = y;
                    y = (y - mean) * rstd;
                    if (L.lnw) y = y * L.lnw[o] + L.lnb[o];
                }
                if (rok && L.pre) L.pre[row * L.out + o] = y;
                y = act_fwd(L.act, y);
                if (rok) L.a[row * L.out + o] = y;
                xb[r * pitch + o] = y;
            }
            __syncthreads();
        }
    }
}

}  // namespace b2rl

namespace b2rl {

// ------------------------------------------------------------------------------------------------
// Backward of the whole head in two launches.
//   head_bwd_kernel   : per 4-row tile, both chains, last layer -> first: through activation and
//                       LayerNorm (dL/da -> dL/dz, stored in place for the weight-gradient pass),
//                       per-tile LayerNorm-affine partial sums, dX = dZ * W down to dL/dlatent.
//   head_wgrad_kernel : one warp per output neuron: dW[o][:] = sum_r dz[r][o] * x[r][:], db[o],
//                       plus the fixed-order reduction of the LayerNorm partials.
// ------------------------------------------------------------------------------------------------
struct HeadBwdLayer {
    const float *w;             // effective weights [out, in]
    const float *lnw;           // LayerNorm weight (nullable)
    const float *a, *z, *pre, *stats;   // forward buffers of the backward rows
    const float *x;             // layer input of the backward rows ([B, in]); latent for the first layer
    float *g;                   // in (chain's last layer): dL/d(output);  out: dL/dz      [B, out]
    float *dw, *db;             // weight / bias gradient destination
    float *dlnw, *dlnb;         // LayerNorm affine gradient destination (nullable)
    float *lnpart;              // [tiles][2][out] partial sums (nullable)
    int in, out, ln, act;
    int acc_w;                  // add into dw/db instead of overwriting
};
struct HeadBwdDesc {
    HeadBwdLayer l[kHeadMaxLayers];
    int wg_start[kHeadMaxLayers + 1];   // first head_wgrad CTA of each layer (8 outputs per CTA)
    int ln_layer[kHeadMaxLayers];       // layers with LayerNorm-affine gradients (one reduce CTA each)
    int n_ln;
    int n_val, n_adv, latent, maxdim;
    float *g_latent;            // out: dL/dlatent [B, latent]
    const float *latent_a;      // nullable: output of the layer that produced the latent; its activation backward is
    int latent_act;             //           folded into the g_latent store (g *= act'(latent_a))
    int accumulate;             // add into dlnw/dlnb instead of overwriting
};

__global__ void __launch_bounds__(kHeadThreads) head_bwd_kernel(const HeadBwdDesc hd, int64_t B) {
    extern __shared__ __align__(16) float hsm[];
    const int pitch = head_pitch(hd.maxdim);
    float *gb = hsm;                                  // [rows][pitch] dL/dz of the current layer
    float *nb = gb + kHeadRows * pitch;               // [rows][pitch] dL/d(input) being built / xhat
    float *tb = nb + kHeadRows * pitch;               // [rows][pitch] dL/dy before the LayerNorm weight
    float *lat = tb + kHeadRows * pitch;              // [rows][latent] accumulated dL/dlatent
    float *red = lat + kHeadRows * hd.latent;         // [warps][rows] row-sum scratch
    float *wp = red + (kHeadThreads / 32) * kHeadRows;   // [warps][rows][pitch] per-warp dX partials
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = tid & (kHeadRows - 1), og = tid / kHeadRows;
    const int64_t row = (int64_t)blockIdx.x * kHeadRows + r;
    const bool rok = row < B;
    for (int i = tid; i < kHeadRows * hd.latent; i += kHeadThreads) lat[i] = 0.f;

    for (int chain = 0; chain < 2; ++chain) {
        const int l0 = chain == 0 ? 0 : hd.n_val, l1 = chain == 0 ? hd.n_val : hd.n_val + hd.n_adv;
        for (int li = l1 - 1; li >= l0; --li) {
            const HeadBwdLayer L = hd.l[li];
            __syncthreads();
            // ---- dL/d(output): global for the chain's last layer, else what the layer above produced
            if (li == l1 - 1) {
                for (int o = og; o < L.out; o += kHeadLanes) gb[r * pitch + o] = rok ? L.g[row * L.out + o] : 0.f;
            } else {
                for (int o = og; o < L.out; o += kHeadLanes) gb[r * pitch + o] = nb[r * pitch + o];
            }
            __syncthreads();
            // ---- through activation (+ LayerNorm): gb <- dL/dz
            if (L.ln != B2RL_LN_NONE) {
                const float mean = rok ? L.stats[row * 2] : 0.f, rstd = rok ? L.stats[row * 2 + 1] : 0.f;
                float s1 = 0.f, s2 = 0.f;
                for (int o = og; o < L.out; o += kHeadLanes) {
                    float gy = 0.f, xhat = 0.f;
                    if (rok) {
                        const int64_t gi = row * L.out + o;
                        gy = gb[r * pitch + o] * act_bwd(L.act, L.pre ? L.pre[gi] : 0.f, L.a[gi]);
                        xhat = (L.z[gi] - mean) * rstd;
                    }
                    const float gx = L.lnw ? gy * L.lnw[o] : gy;
                    s1 += gx; s2 += gx * xhat;
                    gb[r * pitch + o] = gx;
                    nb[r * pitch + o] = xhat;          // nb was consumed above: reuse it
                    tb[r * pitch + o] = gy;
                }
                const float t1 = head_row_sum(s1, red, tid) / (float)L.out;
                const float t2 = head_row_sum(s2, red, tid) / (float)L.out;
                // LayerNorm-affine partials of this tile (rows past B hold zeros)
                if (L.lnpart && r == 0) {
                    for (int o = og; o < L.out; o += kHeadLanes) {
                        float pw = 0.f, pb = 0.f;
#pragma unroll
                        for (int q = 0; q < kHeadRows; ++q) {
                            pw += tb[q * pitch + o] * nb[q * pitch + o];
                            pb += tb[q * pitch + o];
                        }
                        L.lnpart[((int64_t)blockIdx.x * 2 + 0) * L.out + o] = pw;
                        L.lnpart[((int64_t)blockIdx.x * 2 + 1) * L.out + o] = pb;
                    }
                }
                for (int o = og; o < L.out; o += kHeadLanes) {
                    const float dz = rstd * (gb[r * pitch + o] - t1 - nb[r * pitch + o] * t2);
                    gb[r * pitch + o] = dz;
                    if (rok) L.g[row * L.out + o] = dz;
                }
            } else {
                const bool store = L.act != B2RL_ACT_NONE || li != l1 - 1;
                for (int o = og; o < L.out; o += kHeadLanes) {
                    float dz = gb[r * pitch + o];
                    if (rok) {
                        const int64_t gi = row * L.out + o;
                        if (L.act != B2RL_ACT_NONE) dz *= act_bwd(L.act, L.pre ? L.pre[gi] : 0.f, L.a[gi]);
                        if (store) L.g[gi] = dz;
                    }
                    gb[r * pitch + o] = dz;
                }
            }
            __syncthreads();
            // ---- dX[q][i] = sum_o dz[q][o] * W[o][i]: lanes over i (coalesced rows of W), each warp takes
            // every 8th output; per-warp partials meet in shared memory and are added in warp order
            for (int i0 = 0; i0 < L.in; i0 += 32) {
                const int i = i0 + lane;
                float acc[kHeadRows];
#pragma unroll
                for (int q = 0; q < kHeadRows; ++q) acc[q] = 0.f;
                if (i < L.in) {
#pragma unroll 4
                    for (int o = warp; o < L.out; o += kHeadThreads / 32) {
                        const float w = __ldg(L.w + (int64_t)o * L.in + i);
#pragma unroll
                        for (int q = 0; q < kHeadRows; ++q) acc[q] = fmaf(gb[q * pitch + o], w, acc[q]);
                    }
#pragma unroll
                    for (int q = 0; q < kHeadRows; ++q) wp[(warp * kHeadRows + q) * pitch + i] = acc[q];
                }
            }
            __syncthreads();
            for (int i = og; i < L.in; i += kHeadLanes) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < kHeadThreads / 32; ++w) t += wp[(w * kHeadRows + r) * pitch + i];
                if (li == l0) lat[r * hd.latent + i] += t;
                else nb[r * pitch + i] = t;
            }
        }
    }
    __syncthreads();
    for (int i = og; i < hd.latent; i += kHeadLanes)
        if (rok) {
            float v = lat[r * hd.latent + i];
            if (hd.latent_a) v *= act_bwd(hd.latent_act, 0.f, hd.latent_a[row * hd.latent + i]);
            hd.g_latent[row * hd.latent + i] = v;
        }
}

constexpr int kHeadWgMaxIn = 512;                    // widest layer input the fused weight-gradient kernel takes
constexpr int kHeadWgSmemFloats = 16384;              // 64 KB chunk of the layer input staged per pass

// One CTA = 8 consecutive output neurons of one layer (one per warp).  The layer input x [B, in] and
// the CTA's 8 columns of dZ are staged through shared memory in chunks of rows (coalesced, every load
// in flight at once); each warp then walks the chunk's rows in order with its lanes striding the
// input dimension: dW[o][i] = sum_r dz[r][o] * x[r][i], db[o] = sum_r dz[r][o]  (fixed order).
__global__ void __launch_bounds__(kHeadThreads) head_wgrad_kernel(const HeadBwdDesc hd, int64_t B, int n_tiles) {
    extern __shared__ __align__(16) float wsm[];
    const int nl = hd.n_val + hd.n_adv;
    const int t = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarps = kHeadThreads / 32;
    if (t < hd.wg_start[nl]) {
        int li = 0;
        while (t >= hd.wg_start[li + 1]) ++li;
        const HeadBwdLayer L = hd.l[li];
        const int o0 = (t - hd.wg_start[li]) * kWarps;
        const int o = o0 + warp;
        const int rows_per = max(1, min((int)B, (kHeadWgSmemFloats - 0) / (L.in + kWarps)));
        float *xs = wsm;                                   // [rows_per][in]
        float *dzs = wsm + (size_t)rows_per * L.in;        // [rows_per][8]
        float acc[kHeadWgMaxIn / 32];
#pragma unroll
        for (int c = 0; c < kHeadWgMaxIn / 32; ++c) acc[c] = 0.f;
        float accb = 0.f;
        const bool vec = (L.in & 3) == 0 && (reinterpret_cast<uintptr_t>(L.x) & 15) == 0;
        for (int64_t r0 = 0; r0 < B; r0 += rows_per) {
            const int rc = (int)min((int64_t)rows_per, B - r0);
            __syncthreads();
            if (vec) {
                const float4 *src = reinterpret_cast<const float4 *>(L.x + r0 * L.in);
                float4 *dst = reinterpret_cast<float4 *>(xs);
                for (int e = threadIdx.x; e < rc * L.in / 4; e += kHeadThreads) dst[e] = __ldg(src + e);
            } else {
                for (int e = threadIdx.x; e < rc * L.in; e += kHeadThreads) xs[e] = __ldg(L.x + r0 * L.in + e);
            }
            for (int e = threadIdx.x; e < rc * kWarps; e += kHeadThreads) {
                const int r = e / kWarps, w = e - r * kWarps;
                dzs[e] = (o0 + w < L.out) ? __ldg(L.g + (r0 + r) * L.out + o0 + w) : 0.f;
            }
            __syncthreads();
            if (o < L.out) {
#pragma unroll 4
                for (int r = 0; r < rc; ++r) {
                    const float dz = dzs[r * kWarps + warp];
                    accb += dz;
                    const float *xr = xs + r * L.in + lane;
#pragma unroll
                    for (int c = 0; c < kHeadWgMaxIn / 32; ++c)
                        if (c * 32 < L.in && c * 32 + lane < L.in) acc[c] = fmaf(dz, xr[c * 32], acc[c]);
                }
            }
        }
        if (o < L.out) {
#pragma unroll
            for (int c = 0; c < kHeadWgMaxIn / 32; ++c) {
                const int i = c * 32 + lane;
                if (i < L.in) {
                    float *pw = L.dw + (int64_t)o * L.in + i;
                    *pw = L.acc_w ? *pw + acc[c] : acc[c];
                }
            }
            if (lane == 0) L.db[o] = L.acc_w ? L.db[o] + accb : accb;
        }
    } else {
        // LayerNorm affine gradients: fixed-order sum of the per-tile partials
        const HeadBwdLayer L = hd.l[hd.ln_layer[t - hd.wg_start[nl]]];
        for (int c = threadIdx.x; c < L.out; c += kHeadThreads) {
            float sw = 0.f, sb = 0.f;
            for (int k = 0; k < n_tiles; ++k) {
                sw += L.lnpart[((int64_t)k * 2 + 0) * L.out + c];
                sb += L.lnpart[((int64_t)k * 2 + 1) * L.out + c];
            }
            L.dlnw[c] = hd.accumulate ? L.dlnw[c] + sw : sw;
            L.dlnb[c] = hd.accumulate ? L.dlnb[c] + sb : sb;
        }
    }
}

}  // namespace b2rl
