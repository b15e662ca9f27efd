// conv_st.cuh — forward convolution over fp32 activations (every conv layer after the first) with the input STAGED in
// shared memory by the TMA unit instead of gathered tap by tap from L2.
//
// Why: the im2col gather of conv_tc.cuh issues one 8-byte load per two taps and re-reads every input element
// (k/s)^2 times; at B = 256 it runs at ~1.7 TB/s of L1 requests with the SMs waiting on the scoreboard (ncu: 4 stall
// cycles per issue on long_scoreboard, tensor pipe 10 % busy).  Here a tile's receptive field — for each image the tile
// touches, the input rows [oy_lo*s, oy_hi*s + k) of every channel, which are CONTIGUOUS in NCHW — arrives by
// cp.async.bulk (one copy per image x channel, all of them in flight at once), is read back with 8-byte LDS by the
// producer warps (30-cycle latency, no re-read from L2), split into tf32 hi/lo and stored in the UMMA K-major layout.
//
// Persistent, warp-specialised, one CTA per SM:
//   producer warps (<= 16) : warp = (block of 32 tile rows, one 4-tap chunk of the 16-tap k-block): 2 LDS.64, split, 2 STS.128
//   MMA warp               : per k-block 2 k-steps x (A_hi x [W_hi; W_lo], A_lo x W_hi) into one of two TMEM accumulators
//   copy warp              : lane = input channel: bulk copies of that channel's rows for the NEXT tile as soon as the
//                            producers have released the channel (per-channel full / empty barriers);
//                            the pre-split weights (2 * n_pad * K floats) are copied ONCE per CTA and stay resident
//   epilogue warps (4)     : TMEM -> hi + lo columns -> bias, activation -> coalesced NCHW stores.
// Tiles are R <= 128 consecutive output pixels with R chosen so that the tile count is a multiple of the CTA count.
#pragma once
#include "conv_tc.cuh"

namespace b2rl {

constexpr int kStProdWarps = 16, kStEpiWarps = 4;
constexpr int kStThreads = (kStProdWarps + 2 + kStEpiWarps) * 32;
constexpr int kStBK = 16;          // taps per k-block (4 chunks of 4)
constexpr int kStStages = 3;       // A stages
constexpr int kStMaxSeg = 8;       // images one tile may touch
constexpr int kStMaxCin = 64;
constexpr int kStMaxBGroups = 32;  // weight copies of 4 k-blocks each: K <= 2048
constexpr int kStSmemMax = 227 * 1024;

struct ConvStParams {
    const float *x;              // [rows, Cin, H, W] fp32
    const float *w_hl;           // weight_split_kernel layout
    const float *bias;
    float *out;                  // [rows, N, P]
    int M, N, n_pad, k_pad;
    int P, OW, S, Cin, H, W;
    int R, rows_p, n_prod, n_tiles;
    int rtot_max;                // input rows per channel the slab holds
    int act;
    int dbg;                     // timing experiments (B2RL_ST_DBG): 1 = producers do not wait for the slab, 2 = no slab copies at all,
                                 // 4 = producers skip the proxy fence (results are garbage in every case)
};

// rows of image `bimg` a tile [m0, m1) of output pixels needs: first output row and number of input rows
__host__ __device__ __forceinline__ void st_segment(int bimg, int m0, int m1, int P, int OW, int S, int KS, int &oy_lo, int &n_in) {
    const int p_lo = m0 - bimg * P > 0 ? m0 - bimg * P : 0;
    const int p_hi = m1 - bimg * P < P ? m1 - bimg * P : P;
    oy_lo = p_lo / OW;
    n_in = ((p_hi - 1) / OW - oy_lo) * S + KS;
}

static inline size_t conv_st_smem_bytes(int n_pad, int k_pad, int rows_p, int Cin, int rtot_max, int W) {
    return (size_t)kStStages * 2 * rows_p * kStBK * 4 + (size_t)2 * n_pad * k_pad * 4 + (size_t)Cin * rtot_max * W * 4 +
           (size_t)n_pad * 4 + 16 + 8 * (2 * kStStages + 4 + kStMaxBGroups + 2 * kStMaxCin) + 16 + 128;
}

namespace tc {
__device__ __forceinline__ void lds64(uint32_t addr, float &a, float &b) {
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(addr));
}
}  // namespace tc

template <int KS>
__global__ void __launch_bounds__(kStThreads, 1) conv_fwd_st_kernel(const ConvStParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int KK = KS * KS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = p.k_pad / kStBK, NBG = (KB + 3) >> 2;
    const uint32_t a_part = (uint32_t)p.rows_p * (kStBK * 4), a_stage = 2 * a_part;   // hi rows, then lo rows
    const uint32_t b_kb = (uint32_t)p.n_pad * (2 * kStBK * 4);                        // [W_hi; W_lo] of one k-block
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t a_s = sbase, b_s = a_s + kStStages * a_stage, slab_s = b_s + (uint32_t)KB * b_kb;
    const uint32_t row_bytes = (uint32_t)p.W * 4, chan_stride = (uint32_t)p.rtot_max * row_bytes;
    const uint32_t bias_a = slab_s + (uint32_t)p.Cin * chan_stride;
    const uint32_t bars_a = (bias_a + (uint32_t)p.n_pad * 4 + 15u) & ~15u;
    uint8_t *gen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    uint64_t *full_a = reinterpret_cast<uint64_t *>(gen + (bars_a - sbase));   // [stages] im2col stage written
    uint64_t *empty_a = full_a + kStStages;                                     // [stages] stage consumed (MMA commit)
    uint64_t *acc_full = empty_a + kStStages;                                   // [2] accumulator complete
    uint64_t *acc_empty = acc_full + 2;                                         // [2] accumulator drained
    uint64_t *b_full = acc_empty + 2;                                           // [NBG] weight group landed
    uint64_t *slab_full = b_full + kStMaxBGroups;                               // [NBG] rows of a 64-tap channel group landed
    uint64_t *slab_empty = slab_full + kStMaxCin;                               // [NBG] group released by the producers
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(slab_empty + kStMaxCin);
    const uint32_t lbo_a = (uint32_t)p.rows_p * 16, lbo_b = (uint32_t)p.n_pad * 2 * 16;
    const int my_tiles = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < kStStages; ++s) {
            tc::mbar_init(&full_a[s], (uint32_t)p.n_prod);
            tc::mbar_init(&empty_a[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], kStEpiWarps);
        }
        for (int g = 0; g < NBG; ++g) tc::mbar_init(&b_full[g], 1);
        for (int g = 0; g < NBG; ++g) {
            const int c_lo = (g * 4 * kStBK) / KK, c_end = ((g + 1) * 4 * kStBK + KK - 1) / KK;      // channels of the group
            tc::mbar_init(&slab_full[g], (uint32_t)((c_end < p.Cin ? c_end : p.Cin) - c_lo));          // one arrival per channel copy lane
            tc::mbar_init(&slab_empty[g], (uint32_t)p.n_prod);
        }
        tc::fence_barrier_init();
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 4 * p.n_pad) tmem_cols <<= 1;           // two accumulators of [A.W_hi | A.W_lo]
    if (warp == kStProdWarps) tc::tmem_alloc(tmem_ptr, tmem_cols);
    for (int n = tid; n < p.n_pad; n += kStThreads) {
        const float v = (n < p.N && p.bias) ? p.bias[n] : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_a + 4u * n), "f"(v) : "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    if (warp == kStProdWarps + 1) {
        // ================================ copy warp ================================
        auto weights = [&](int g) {
            const int nkb = KB - 4 * g < 4 ? KB - 4 * g : 4;
            const uint32_t bytes = (uint32_t)nkb * b_kb;
            tc::mbar_expect_tx(&b_full[g], bytes);
            tc::bulk_g2s(b_s + (uint32_t)g * 4u * b_kb, p.w_hl + (size_t)g * 4 * (b_kb / 4), bytes, &b_full[g]);
        };
        if (lane == 0 && my_tiles > 0) weights(0);
        __syncwarp();
        for (int i = 0; i < my_tiles; ++i) {
            const int t = (int)blockIdx.x + i * (int)gridDim.x;
            const int m0 = t * p.R, m1 = (m0 + p.R < p.M) ? m0 + p.R : p.M;
            const int b_first = m0 / p.P, b_last = (m1 - 1) / p.P;
            int rows_tot = 0;
            for (int b = b_first; b <= b_last; ++b) {
                int oy_lo, n_in;
                st_segment(b, m0, m1, p.P, p.OW, p.S, KS, oy_lo, n_in);
                rows_tot += n_in;
            }
            for (int c = lane; c < p.Cin && !(p.dbg & 2); c += 32) {
                const int grp = (c * KK) / (4 * kStBK);                   // 64-tap group this channel belongs to
                if (i > 0) tc::mbar_wait(&slab_empty[grp], (uint32_t)((i - 1) & 1));
                tc::mbar_expect_tx(&slab_full[grp], (uint32_t)rows_tot * row_bytes);
                uint32_t dst = slab_s + (uint32_t)c * chan_stride;
                for (int b = b_first; b <= b_last; ++b) {
                    int oy_lo, n_in;
                    st_segment(b, m0, m1, p.P, p.OW, p.S, KS, oy_lo, n_in);
                    const float *src = p.x + (((int64_t)b * p.Cin + c) * p.H + (int64_t)oy_lo * p.S) * p.W;
                    tc::bulk_g2s(dst, src, (uint32_t)n_in * row_bytes, &slab_full[grp]);
                    dst += (uint32_t)n_in * row_bytes;
                }
            }
            if (i == 0) {
                if (lane == 0)
                    for (int g = 1; g < NBG; ++g) weights(g);
            }
            __syncwarp();
        }
    } else if (warp == kStProdWarps) {
        // ================================ MMA warp ================================
        const uint32_t idesc2 = tc::make_idesc_tf32(kTcBM, 2 * p.n_pad);      // A_hi x [W_hi; W_lo]
        const uint32_t idesc1 = tc::make_idesc_tf32(kTcBM, p.n_pad);          // A_lo x W_hi
        const uint64_t da_step = (uint64_t)((2 * lbo_a) >> 4), db_step = (uint64_t)((2 * lbo_b) >> 4);
        int stage = 0;
        uint32_t sph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int acc = i & 1;
            tc::mbar_wait(&acc_empty[acc], (uint32_t)(((i >> 1) & 1) ^ 1));     // passes at first use
            const uint32_t d_addr = tmem_d + (uint32_t)(acc * 2 * p.n_pad);
            for (int kb = 0; kb < KB; ++kb) {
                if (i == 0 && (kb & 3) == 0) tc::mbar_wait(&b_full[kb >> 2], 0);
                if (!(p.dbg & 8)) tc::mbar_wait(&full_a[stage], sph);
                tc::tc_fence_after();
                const uint32_t a_addr = a_s + (uint32_t)stage * a_stage;
                const uint64_t dah0 = tc::make_desc(a_addr, lbo_a, 128), dal0 = tc::make_desc(a_addr + a_part, lbo_a, 128);
                const uint64_t db0 = tc::make_desc(b_s + (uint32_t)kb * b_kb, lbo_b, 128);
                if (tc::elect_one()) {
#pragma unroll
                    for (int j = 0; j < kStBK / 8; ++j) {
                        if (p.dbg & 16) break;
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
            const int rb = warp % nrb, q = warp / nrb;                     // block of 32 rows, chunk of the k-block
            const int r = rb * 32 + lane;
            const uint32_t dst_off = (uint32_t)q * lbo_a + (uint32_t)(r >> 3) * 128 + (uint32_t)(r & 7) * 16;
            int stage = 0;
            uint32_t sph = 1;                                              // empty barriers pass at first use
            for (int i = 0; i < my_tiles; ++i) {
                const int t = (int)blockIdx.x + i * (int)gridDim.x;
                const int m0 = t * p.R, m1 = (m0 + p.R < p.M) ? m0 + p.R : p.M;
                const int m = (m0 + r < m1) ? m0 + r : m1 - 1;            // rows beyond the tile repeat its last row
                const int b = m / p.P, pix = m - b * p.P;
                const int oy = pix / p.OW, ox = pix - oy * p.OW;
                int rowbase = 0, oy_lo = 0, n_in = 0;
                for (int bb = m0 / p.P; bb <= b; ++bb) {
                    rowbase += n_in;
                    st_segment(bb, m0, m1, p.P, p.OW, p.S, KS, oy_lo, n_in);
                }
                const uint32_t src_row = slab_s + (uint32_t)(rowbase + (oy - oy_lo) * p.S) * row_bytes + (uint32_t)(ox * p.S) * 4u;
                const uint32_t tph = (uint32_t)(i & 1);
                for (int g = 0; g < NBG; ++g) {
                    if (!(p.dbg & 3)) tc::mbar_wait(&slab_full[g], tph);   // one poll per 64 taps
                    const int nk = KB - 4 * g < 4 ? KB - 4 * g : 4;
                    // all loads of the group first: their latency overlaps the previous k-block's fence
                    float v[4][4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (kk >= nk) break;
                        const int kb = 4 * g + kk;
                        int c, ky, kx0;
                        if (KS == 4) { c = kb; ky = q; kx0 = 0; }
                        else { c = kb >> 2; ky = ((kb & 3) << 1) + (q >> 1); kx0 = (q & 1) * 4; }
                        const uint32_t addr = src_row + (uint32_t)c * chan_stride + (uint32_t)ky * row_bytes + (uint32_t)kx0 * 4u;
                        tc::lds64(addr, v[kk][0], v[kk][1]);
                        tc::lds64(addr + 8u, v[kk][2], v[kk][3]);
                    }
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (kk >= nk) break;
                        float hi[4], lo[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) { hi[j] = tc::tf32_rn_fast(v[kk][j]); lo[j] = v[kk][j] - hi[j]; }
                        if (!(p.dbg & 8) && !tc::mbar_test(&empty_a[stage], sph)) tc::mbar_wait(&empty_a[stage], sph);
                        const uint32_t dst = a_s + (uint32_t)stage * a_stage + dst_off;
                        tc::sts128(dst, hi[0], hi[1], hi[2], hi[3]);
                        tc::sts128(dst + a_part, lo[0], lo[1], lo[2], lo[3]);
                        if (!(p.dbg & 4)) tc::fence_async_smem();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&full_a[stage]);
                        if (++stage == kStStages) { stage = 0; sph ^= 1u; }
                    }
                    if (lane == 0) tc::mbar_arrive(&slab_empty[g]);
                }
            }
        }
    } else {
        // ================================ epilogue warps ================================
        const int q = warp & 3;                                            // TMEM lanes 32*q .. 32*q+31
        const bool relu = p.act == B2RL_ACT_RELU, ident = p.act == B2RL_ACT_NONE;
        const int64_t oP = p.P;
        for (int i = 0; i < my_tiles; ++i) {
            const int t = (int)blockIdx.x + i * (int)gridDim.x;
            const int m0 = t * p.R, m1 = (m0 + p.R < p.M) ? m0 + p.R : p.M;
            const int em = m0 + q * 32 + lane;
            const bool e_ok = em < m1;
            int b_img = 0, pix = 0;
            if (e_ok) { b_img = em / p.P; pix = em - b_img * p.P; }
            const int acc = i & 1;
            tc::mbar_wait(&acc_full[acc], (uint32_t)((i >> 1) & 1));
            tc::tc_fence_after();
            const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * p.n_pad);
            const bool warp_has_rows = m0 + q * 32 < m1;
            for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
                if (!warp_has_rows) break;
                uint32_t rh[16], rl[16];
                tc::tmem_ld16(lane_addr + (uint32_t)c0, rh);
                tc::tmem_ld16(lane_addr + (uint32_t)(p.n_pad + c0), rl);
                if (c0 + 16 >= p.n_pad) {                                  // every TMEM read of this tile is in registers
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&acc_empty[acc]);
                }
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    v[j] = (__uint_as_float(rh[j]) + __uint_as_float(rl[j])) + __uint_as_float(tc::lds32(bias_a + 4u * (c0 + j)));
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if (!ident) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = act_fwd_slow(p.act, v[j]);
                }
                if (e_ok) {
                    float *o = p.out + ((int64_t)b_img * p.N + c0) * oP + pix;
                    const int nv = p.N - c0 < 16 ? p.N - c0 : 16;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < nv) o[j * oP] = v[j];
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

// B2RL_ST selects which layers' passes take the staged kernels: a string of the letters f (forward), w (weight gradient),
// d (input gradient); default "d" — measured on B200 at the benchmark shapes the staged input gradient is 2.3x faster than
// the per-class gather kernel, while the staged forward / weight-gradient kernels (16-tap stages: one mbarrier round trip
// and one proxy fence per 4 taps of a row) are still slower than the gather kernels they would replace.
static int &st_mask() {
    static int mask = -1;
    if (mask < 0) {
        const char *e = getenv("B2RL_ST");
        if (!e) e = "d";
        mask = 0;
        for (; *e; ++e) mask |= (*e == 'f') ? 1 : (*e == 'w') ? 2 : (*e == 'd') ? 4 : 0;
    }
    return mask;
}
static bool st_enabled(char which) { return (st_mask() & (which == 'f' ? 1 : which == 'w' ? 2 : 4)) != 0; }

// returns B2RL_OK, or 1 when the layer is outside what this kernel handles (caller falls back to conv_tc.cuh's gather kernel).
// wsplit as in launch_conv_fwd_tc (same split layout, so the two paths may share it).
static int launch_conv_fwd_st(const b2rl_layer &l, const float *x, const float *W, const float *bias, float *out, int64_t rows,
                              float *wsplit, size_t wsplit_cap, cudaStream_t s, bool reuse_split = false) {
    if (!st_enabled('f')) return 1;
    const int KS = l.ksize, KK = KS * KS, K = l.in_c * KK, P = l.out_h * l.out_w;
    const int n_pad = (l.out_c + 15) / 16 * 16, k_pad = (K + kTcBK - 1) / kTcBK * kTcBK;
    if (!(KS == 4 || KS == 8) || K != k_pad || n_pad > 64 || l.in_c > kStMaxCin || k_pad / kStBK > 4 * kStMaxBGroups) return 1;
    if (l.stride % 2 != 0 || l.in_w % 4 != 0 || (l.in_h * l.in_w) % 4 != 0 || reinterpret_cast<uintptr_t>(x) % 16 != 0) return 1;
    if (rows * (int64_t)P > INT32_MAX || rows < 1) return 1;
    if (wsplit == nullptr || conv_tc_wsplit_floats(l) > wsplit_cap || reinterpret_cast<uintptr_t>(wsplit) % 16 != 0) return 1;
    const int M = (int)(rows * P), sms = sm_count();
    // tile rows: the smallest wave count whose tiles fit, tile count a multiple of the CTA count
    int R = 0, rows_p = 0, rtot_max = 0, n_tiles = 0;
    size_t smem = 0;
    for (int w = (M + sms * kTcBM - 1) / (sms * kTcBM); w <= 64; ++w) {
        R = (M + sms * w - 1) / (sms * w);
        if (R < 1) R = 1;
        n_tiles = (M + R - 1) / R;
        rows_p = (R + 31) / 32 * 32;
        rtot_max = 0;
        bool ok = true;
        for (int t = 0; t < n_tiles && ok; ++t) {
            const int m0 = t * R, m1 = (m0 + R < M) ? m0 + R : M;
            const int b_first = m0 / P, b_last = (m1 - 1) / P;
            if (b_last - b_first + 1 > kStMaxSeg) ok = false;
            int tot = 0;
            for (int b = b_first; b <= b_last && ok; ++b) {
                int oy_lo, n_in;
                st_segment(b, m0, m1, P, l.out_w, l.stride, KS, oy_lo, n_in);
                tot += n_in;
            }
            if (tot > rtot_max) rtot_max = tot;
        }
        smem = conv_st_smem_bytes(n_pad, k_pad, rows_p, l.in_c, rtot_max, l.in_w);
        if (ok && smem <= (size_t)kStSmemMax) break;
        if (R <= 8) return 1;
        R = 0;
    }
    if (R == 0) return 1;
    float *w_hl = wsplit;
    uint32_t *koff = reinterpret_cast<uint32_t *>(wsplit + (size_t)2 * n_pad * k_pad);
    if (!reuse_split) {
        const int total = n_pad * k_pad;
        weight_split_kernel<<<(total + 255) / 256, 256, 0, s>>>(W, l.out_c, K, n_pad, k_pad, w_hl, nullptr, KK, KS, l.in_h * l.in_w,
                                                                l.in_w, koff);
        B2RL_LAUNCH_CHECK();
    }
    ConvStParams p;
    p.x = x; p.w_hl = w_hl; p.bias = bias; p.out = out;
    p.M = M; p.N = l.out_c; p.n_pad = n_pad; p.k_pad = k_pad;
    p.P = P; p.OW = l.out_w; p.S = l.stride; p.Cin = l.in_c; p.H = l.in_h; p.W = l.in_w;
    p.R = R; p.rows_p = rows_p; p.n_prod = (rows_p / 32) * (kStBK / 4); p.n_tiles = n_tiles;
    p.rtot_max = rtot_max; p.act = l.act;
    { const char *e = getenv("B2RL_ST_DBG"); p.dbg = e ? atoi(e) : 0; }
    const int grid = n_tiles < sms ? n_tiles : sms;
    auto launch = [&](auto kern) -> int {
        static bool attr_set[2] = {false, false};
        const int slot = KS == 4 ? 0 : 1;
        if (!attr_set[slot]) {
            B2RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStSmemMax));
            attr_set[slot] = true;
        }
        kern<<<grid, kStThreads, smem, s>>>(p);
        B2RL_LAUNCH_CHECK();
        ++g_conv_path[2];
        return B2RL_OK;
    };
    if (KS == 4) return launch(conv_fwd_st_kernel<4>);
    return launch(conv_fwd_st_kernel<8>);
}

}  // namespace b2rl
