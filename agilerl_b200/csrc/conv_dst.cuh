// conv_dst.cuh — convolution input gradient for kernel = 2 x stride (k4 s2), operands staged in shared memory.
//
// For stride s the input pixels split into s*s parity classes; each class is a stride-1 convolution of G = dL/dout with
// a 2x2 kernel and padding 1 (conv_tc.cuh explains the decomposition).  The im2col operand of those convolutions is THE
// SAME for every class — only the weights differ — so here the classes become extra output columns of one product:
//
//   D[position (b, yy, xx)][cls * Cin + ci] = sum_{co, a, b'} G[b, co, yy-1+a, xx-1+b'] * W[co, ci, py + s(1-a), px + s(1-b')]
//   dX[b, ci, s*yy + py, s*xx + px] = D[...][(py*s + px) * Cin + ci]
//
// i.e. M = rows * ceil(H/s) * ceil(W/s) positions, N = s*s*Cin (128 for the north-star layer), K = 4*Cout: a quarter of
// the im2col work of the per-class kernel, one launch instead of s*s grid slices, and the epilogue owns all four pixels of
// a 2x2 output block, so it writes 8-byte pairs that are contiguous across a warp.
//
// Same persistent structure as conv_st.cuh: whole images of G arrive by cp.async.bulk (double buffered across tiles),
// producer warps read 2x2 windows with bounds masks (the zero padding), split tf32 hi/lo, K-major stores; the weights
// ([hi rows | lo rows] x K, 128 KB) stay resident; two TMEM accumulators of 2*N columns; four epilogue warps.
#pragma once
#include "conv_st.cuh"
#include "conv_i8.cuh"

namespace b2rl {

constexpr int kDsMaxSeg = 4;
constexpr int kDsEpiWarps = 8;                                   // the epilogue moves 4x the columns of the forward kernel's
constexpr int kDsThreads = (kStProdWarps + 2 + kDsEpiWarps) * 32;

struct ConvDstParams {
    const float *g;              // [rows, Cout, OH, OW]
    const float *w_hl;           // dgrad_st_weight_split_kernel layout
    float *dx;                   // [rows, Cin, IH, IW]
    int M;                       // rows * Yc * Xc positions
    int Cout, OH, OW, Cin, IH, IW;
    int Yc, Xc;                  // class grid: ceil(IH/2), ceil(IW/2)
    int ncp, NT, k_pad;          // Cin padded to 16, 4*ncp, 4*Cout
    int R, rows_p, n_prod, n_tiles, nseg_max;
    uint32_t img_bytes;          // Cout*OH*OW*4
};

// combined class weights, chunk-major (one k-block = 4 chunks of 4 taps, k = (co, a, b')):
//   [chunk k/4][part hi, lo][n = cls*ncp + ci][k%4]      value = W[co][ci][py + 2(1-a)][px + 2(1-b')], cls = py*2 + px
__global__ void dgrad_st_weight_split_kernel(const float *__restrict__ w, int Cout, int Cin, int ncp, float *__restrict__ w_hl) {
    const int NT = 4 * ncp, K = 4 * Cout, total = NT * K;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = e / NT, n = e - k * NT;
        const int co = k >> 2, a = (k >> 1) & 1, b = k & 1;
        const int cls = n / ncp, ci = n - cls * ncp;
        const int py = cls >> 1, px = cls & 1;
        float v = 0.f;
        if (ci < Cin) v = w[(((int64_t)co * Cin + ci) * 4 + (py + 2 * (1 - a))) * 4 + (px + 2 * (1 - b))];
        const float hi = tc::tf32_rn(v);
        const int64_t o = (int64_t)(k >> 2) * 2 * NT * 4 + (n >> 3) * 32 + (n & 7) * 4 + (k & 3);
        w_hl[o] = hi;
        w_hl[o + (int64_t)NT * 4] = v - hi;
    }
}

static inline size_t conv_dst_scratch_floats(const b2rl_layer &l) {
    const int ncp = (l.in_c + 15) / 16 * 16;
    return (size_t)2 * 4 * ncp * 4 * l.out_c;
}

static inline size_t conv_dst_smem_bytes(int NT, int k_pad, int rows_p, int nseg_max, uint32_t img_bytes) {
    return (size_t)kStStages * 2 * rows_p * kStBK * 4 + (size_t)2 * NT * k_pad * 4 + (size_t)2 * nseg_max * img_bytes +
           8 * (2 * kStStages + 4 + kStMaxBGroups + 4) + 16 + 128;
}

__global__ void __launch_bounds__(kDsThreads, 1) conv_dgrad_st_kernel(const ConvDstParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = p.k_pad / kStBK, NBG = (KB + 3) >> 2;
    const int Pc = p.Yc * p.Xc, Pg = p.OH * p.OW;
    const uint32_t a_part = (uint32_t)p.rows_p * (kStBK * 4), a_stage = 2 * a_part;
    const uint32_t b_kb = (uint32_t)p.NT * (2 * kStBK * 4);
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t a_s = sbase, b_s = a_s + kStStages * a_stage, slab_s = b_s + (uint32_t)KB * b_kb;
    const uint32_t slab_buf = (uint32_t)p.nseg_max * p.img_bytes;
    const uint32_t bars_a = (slab_s + 2u * slab_buf + 15u) & ~15u;
    uint8_t *gen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    uint64_t *full_a = reinterpret_cast<uint64_t *>(gen + (bars_a - sbase));
    uint64_t *empty_a = full_a + kStStages;
    uint64_t *acc_full = empty_a + kStStages;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *b_full = acc_empty + 2;                  // [NBG]
    uint64_t *slab_full = b_full + kStMaxBGroups;      // [2]
    uint64_t *slab_empty = slab_full + 2;              // [2]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(slab_empty + 2);
    const uint32_t lbo_a = (uint32_t)p.rows_p * 16, lbo_b = (uint32_t)p.NT * 2 * 16;
    const int my_tiles = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < kStStages; ++s) {
            tc::mbar_init(&full_a[s], (uint32_t)p.n_prod);
            tc::mbar_init(&empty_a[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], kDsEpiWarps);
            tc::mbar_init(&slab_full[s], 1);
            tc::mbar_init(&slab_empty[s], (uint32_t)p.n_prod);
        }
        for (int g = 0; g < NBG; ++g) tc::mbar_init(&b_full[g], 1);
        tc::fence_barrier_init();
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 4 * p.NT) tmem_cols <<= 1;              // two accumulators of [hi | lo] columns
    if (warp == kStProdWarps) tc::tmem_alloc(tmem_ptr, tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    auto tile_range = [&](int i, int &m0, int &m1) {
        const int t = (int)blockIdx.x + i * (int)gridDim.x;
        m0 = t * p.R;
        m1 = (m0 + p.R < p.M) ? m0 + p.R : p.M;
    };

    if (warp == kStProdWarps + 1) {
        // ================================ copy warp ================================
        if (lane == 0) {
            for (int i = 0; i < my_tiles; ++i) {
                int m0, m1;
                tile_range(i, m0, m1);
                const int b_first = m0 / Pc, b_last = (m1 - 1) / Pc;
                const int buf = i & 1;
                if (i >= 2) tc::mbar_wait(&slab_empty[buf], (uint32_t)(((i >> 1) - 1) & 1));
                const uint32_t bytes = (uint32_t)(b_last - b_first + 1) * p.img_bytes;
                tc::mbar_expect_tx(&slab_full[buf], bytes);
                tc::bulk_g2s(slab_s + (uint32_t)buf * slab_buf, p.g + (int64_t)b_first * (p.img_bytes / 4), bytes, &slab_full[buf]);
                if (i == 0)
                    for (int g = 0; g < NBG; ++g) {
                        const int nkb = KB - 4 * g < 4 ? KB - 4 * g : 4;
                        const uint32_t wb = (uint32_t)nkb * b_kb;
                        tc::mbar_expect_tx(&b_full[g], wb);
                        tc::bulk_g2s(b_s + (uint32_t)g * 4u * b_kb, p.w_hl + (size_t)g * 4 * (b_kb / 4), wb, &b_full[g]);
                    }
            }
        }
        __syncwarp();
    } else if (warp == kStProdWarps) {
        // ================================ MMA warp ================================
        const uint32_t idesc2 = tc::make_idesc_tf32(kTcBM, 2 * p.NT);
        const uint32_t idesc1 = tc::make_idesc_tf32(kTcBM, p.NT);
        const uint64_t da_step = (uint64_t)((2 * lbo_a) >> 4), db_step = (uint64_t)((2 * lbo_b) >> 4);
        int stage = 0;
        uint32_t sph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int acc = i & 1;
            tc::mbar_wait(&acc_empty[acc], (uint32_t)(((i >> 1) & 1) ^ 1));
            const uint32_t d_addr = tmem_d + (uint32_t)(acc * 2 * p.NT);
            for (int kb = 0; kb < KB; ++kb) {
                if (i == 0 && (kb & 3) == 0) tc::mbar_wait(&b_full[kb >> 2], 0);
                tc::mbar_wait(&full_a[stage], sph);
                tc::tc_fence_after();
                const uint32_t a_addr = a_s + (uint32_t)stage * a_stage;
                const uint64_t dah0 = tc::make_desc(a_addr, lbo_a, 128), dal0 = tc::make_desc(a_addr + a_part, lbo_a, 128);
                const uint64_t db0 = tc::make_desc(b_s + (uint32_t)kb * b_kb, lbo_b, 128);
                if (tc::elect_one()) {
#pragma unroll
                    for (int j = 0; j < kStBK / 8; ++j) {
                        tc::mma_tf32(d_addr, dah0 + j * da_step, db0 + j * db_step, idesc2, (kb | j) ? 1u : 0u);
                        tc::mma_tf32(d_addr, dal0 + j * da_step, db0 + j * db_step, idesc1, 1u);
                    }
                    tc::mma_commit(&empty_a[stage]);
                    if (kb == KB - 1) tc::mma_commit(&acc_full[acc]);
                }
                __syncwarp();
                if (++stage == kStStages) { stage = 0; sph ^= 1u; }
            }
        }
    } else if (warp < kStProdWarps) {
        // ================================ producer warps ================================
        if (warp < p.n_prod) {
            const int nrb = p.rows_p >> 5;
            const int rb = warp % nrb, q = warp / nrb;
            const int r = rb * 32 + lane;
            const uint32_t dst_off = (uint32_t)q * lbo_a + (uint32_t)(r >> 3) * 128 + (uint32_t)(r & 7) * 16;
            const uint32_t ow4 = (uint32_t)p.OW * 4u, pg4 = (uint32_t)Pg * 4u;
            int stage = 0;
            uint32_t sph = 1;
            for (int i = 0; i < my_tiles; ++i) {
                int m0, m1;
                tile_range(i, m0, m1);
                const int buf = i & 1;
                const int m = (m0 + r < m1) ? m0 + r : m1 - 1;
                const int b = m / Pc, pos = m - b * Pc;
                const int yy = pos / p.Xc, xx = pos - yy * p.Xc;
                // window rows yy-1, yy and columns xx-1, xx of G (zero outside the plane)
                const bool y0 = yy >= 1 && yy - 1 < p.OH, y1 = yy < p.OH, x0 = xx >= 1 && xx - 1 < p.OW, x1 = xx < p.OW;
                const bool m00 = y0 && x0, m01 = y0 && x1, m10 = y1 && x0, m11 = y1 && x1;
                const uint32_t base = slab_s + (uint32_t)buf * slab_buf + (uint32_t)(b - m0 / Pc) * p.img_bytes +
                                      (uint32_t)(((yy - 1) * p.OW + (xx - 1)) * 4);      // wraps below the plane only where masked
                tc::mbar_wait(&slab_full[buf], (uint32_t)((i >> 1) & 1));
                for (int kb = 0; kb < KB; ++kb) {
                    const uint32_t addr = base + (uint32_t)(kb * 4 + q) * pg4;
                    float v[4];
                    v[0] = m00 ? __uint_as_float(tc::lds32(addr)) : 0.f;
                    v[1] = m01 ? __uint_as_float(tc::lds32(addr + 4u)) : 0.f;
                    v[2] = m10 ? __uint_as_float(tc::lds32(addr + ow4)) : 0.f;
                    v[3] = m11 ? __uint_as_float(tc::lds32(addr + ow4 + 4u)) : 0.f;
                    float hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { hi[j] = tc::tf32_rn_fast(v[j]); lo[j] = v[j] - hi[j]; }
                    tc::mbar_wait(&empty_a[stage], sph);
                    const uint32_t dst = a_s + (uint32_t)stage * a_stage + dst_off;
                    tc::sts128(dst, hi[0], hi[1], hi[2], hi[3]);
                    tc::sts128(dst + a_part, lo[0], lo[1], lo[2], lo[3]);
                    tc::fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tc::mbar_arrive(&full_a[stage]);
                        if (kb == KB - 1) tc::mbar_arrive(&slab_empty[buf]);
                    }
                    if (++stage == kStStages) { stage = 0; sph ^= 1u; }
                }
            }
        }
    } else {
        // ================================ epilogue warps ================================
        const int q = warp & 3, chalf = (warp - (kStProdWarps + 2)) >> 2;       // TMEM lane quarter, half of the channel groups
        const int64_t plane = (int64_t)p.IH * p.IW;
        for (int i = 0; i < my_tiles; ++i) {
            int m0, m1;
            tile_range(i, m0, m1);
            const int em = m0 + q * 32 + lane;
            const bool e_ok = em < m1;
            int b = 0, yy = 0, xx = 0;
            if (e_ok) { b = em / Pc; const int pos = em - b * Pc; yy = pos / p.Xc; xx = pos - yy * p.Xc; }
            const int acc = i & 1;
            tc::mbar_wait(&acc_full[acc], (uint32_t)((i >> 1) & 1));
            tc::tc_fence_after();
            const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * p.NT);
            const bool warp_has_rows = m0 + q * 32 < m1;
            const bool pair_ok = 2 * xx + 1 < p.IW;
            float *o_base = p.dx + (int64_t)b * p.Cin * plane + (int64_t)(2 * yy) * p.IW + 2 * xx;
            for (int c0 = chalf * 8; c0 < p.ncp && warp_has_rows; c0 += 16) {
                float v[4][8];
#pragma unroll
                for (int cls = 0; cls < 4; ++cls) {
                    uint32_t rh[8], rl[8];
                    tc::tmem_ld8(lane_addr + (uint32_t)(cls * p.ncp + c0), rh);
                    tc::tmem_ld8(lane_addr + (uint32_t)(p.NT + cls * p.ncp + c0), rl);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[cls][j] = __uint_as_float(rh[j]) + __uint_as_float(rl[j]);
                }
                if (c0 + 16 >= p.ncp) {
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
                }
                if (e_ok) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (c0 + j >= p.Cin) break;
                        float *o = o_base + (int64_t)(c0 + j) * plane;
#pragma unroll
                        for (int py = 0; py < 2; ++py) {
                            if (2 * yy + py >= p.IH) break;
                            if (pair_ok) *reinterpret_cast<float2 *>(o + py * p.IW) = make_float2(v[py * 2][j], v[py * 2 + 1][j]);
                            else o[py * p.IW] = v[py * 2][j];
                        }
                    }
                }
            }
            if (!warp_has_rows) {
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kStProdWarps) tc::tmem_dealloc(tmem_d, tmem_cols);
}

// dL/d(layer input) [rows, Cin, in_h, in_w] (overwritten).  returns B2RL_OK, or 1 when the layer is outside this kernel.
// the combined class weights of a layer this kernel can take, into `scratch` (conv_dst_scratch_floats): what
// launch_conv_dgrad_st does itself unless it is handed the result (presplit)
static bool conv_dgrad_st_shape_ok(const b2rl_layer &l) {
    if (!st_enabled('d') || l.ksize != 4 || l.stride != 2 || l.in_w % 2 != 0) return false;
    const int ncp = (l.in_c + 15) / 16 * 16, k_pad = 4 * l.out_c;
    return 4 * ncp <= 128 && k_pad % kStBK == 0 && k_pad / kStBK <= 4 * kStMaxBGroups;
}
static int launch_dgrad_st_weight_split(const b2rl_layer &l, const float *W, float *scratch, size_t scratch_cap, cudaStream_t s) {
    if (!conv_dgrad_st_shape_ok(l) || scratch == nullptr || conv_dst_scratch_floats(l) > scratch_cap ||
        reinterpret_cast<uintptr_t>(scratch) % 16 != 0)
        return 1;
    const int ncp = (l.in_c + 15) / 16 * 16, total = 4 * ncp * 4 * l.out_c;
    dgrad_st_weight_split_kernel<<<(total + 255) / 256, 256, 0, s>>>(W, l.out_c, l.in_c, ncp, scratch);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

static int launch_conv_dgrad_st(const b2rl_layer &l, const float *g, const float *W, float *g_in, int64_t rows, float *scratch,
                                size_t scratch_cap, cudaStream_t s, const float *presplit = nullptr) {
    if (!st_enabled('d')) return 1;
    if (l.ksize != 4 || l.stride != 2 || l.in_w % 2 != 0) return 1;
    const int ncp = (l.in_c + 15) / 16 * 16, NT = 4 * ncp, k_pad = 4 * l.out_c;
    const int Yc = (l.in_h + 1) / 2, Xc = (l.in_w + 1) / 2, Pc = Yc * Xc, Pg = l.out_h * l.out_w;
    if (NT > 128 || k_pad % kStBK != 0 || k_pad / kStBK > 4 * kStMaxBGroups) return 1;
    if (((int64_t)l.out_c * Pg) % 4 != 0 || reinterpret_cast<uintptr_t>(g) % 16 != 0 || reinterpret_cast<uintptr_t>(g_in) % 8 != 0) return 1;
    if (rows * (int64_t)Pc > INT32_MAX || rows < 1) return 1;
    if (!presplit && (scratch == nullptr || conv_dst_scratch_floats(l) > scratch_cap || reinterpret_cast<uintptr_t>(scratch) % 16 != 0))
        return 1;
    const int M = (int)(rows * Pc), sms = sm_count();
    const uint32_t img_bytes = (uint32_t)l.out_c * Pg * 4u;
    int R = 0, rows_p = 0, n_tiles = 0, nseg_max = 0;
    size_t smem = 0;
    for (int w = (M + sms * kTcBM - 1) / (sms * kTcBM); w <= 64; ++w) {
        R = (M + sms * w - 1) / (sms * w);
        if (R < 1) R = 1;
        n_tiles = (M + R - 1) / R;
        rows_p = (R + 31) / 32 * 32;
        nseg_max = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int m0 = t * R, m1 = (m0 + R < M) ? m0 + R : M;
            const int ns = (m1 - 1) / Pc - m0 / Pc + 1;
            if (ns > nseg_max) nseg_max = ns;
        }
        smem = conv_dst_smem_bytes(NT, k_pad, rows_p, nseg_max, img_bytes);
        if (nseg_max <= kDsMaxSeg && smem <= (size_t)kStSmemMax) break;
        if (R <= 8) return 1;
        R = 0;
    }
    if (R == 0) return 1;
    if (!presplit) {
        const int total = NT * k_pad;
        dgrad_st_weight_split_kernel<<<(total + 255) / 256, 256, 0, s>>>(W, l.out_c, l.in_c, ncp, scratch);
        B2RL_LAUNCH_CHECK();
    }
    ConvDstParams p;
    p.g = g; p.w_hl = presplit ? presplit : scratch; p.dx = g_in; p.M = M;
    p.Cout = l.out_c; p.OH = l.out_h; p.OW = l.out_w; p.Cin = l.in_c; p.IH = l.in_h; p.IW = l.in_w;
    p.Yc = Yc; p.Xc = Xc; p.ncp = ncp; p.NT = NT; p.k_pad = k_pad;
    p.R = R; p.rows_p = rows_p; p.n_prod = (rows_p / 32) * (kStBK / 4); p.n_tiles = n_tiles; p.nseg_max = nseg_max;
    p.img_bytes = img_bytes;
    static bool attr_set = false;
    if (!attr_set) {
        B2RL_CUDA(cudaFuncSetAttribute(conv_dgrad_st_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStSmemMax));
        attr_set = true;
    }
    const int grid = n_tiles < sms ? n_tiles : sms;
    conv_dgrad_st_kernel<<<grid, kDsThreads, smem, s>>>(p);
    B2RL_LAUNCH_CHECK();
    ++g_conv_path[2];
    return B2RL_OK;
}

}  // namespace b2rl
