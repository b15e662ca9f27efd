// conv_wst.cuh — convolution weight gradient with both operands STAGED in shared memory by the TMA unit.
//
//   dW[co][tap] = sum_pix im2col(x)[pix][tap] * G[pix][co]        (+ db[co] = sum_pix G[pix][co])
//
// UMMA view: D[M = 128 taps][N = Cout] += A[taps][K = pixels] * B[Cout][pixels], both operands K-major (pixels contiguous).
// (tcgen05 kind::tf32 does NOT take MN-major operands on this part — tools/probes/umma_mn_probe*.cu: a_major = 1 or
// b_major = 1 returns zeros for tf32 while the same descriptors work for bf16 — so the im2col tile has to be transposed.)
// The transpose costs nothing extra here because the input already sits in shared memory: a producer thread owns
// (one tap, 4 consecutive pixels), reads its 4 values with LDS.32 at the 4 pixels' slab addresses (a per-tile table) plus
// the tap's offset, and writes ONE 16-byte chunk of the K-major tile
//   (tap m, pixel k) at (k/4)*LBO + (m/8)*128 + (m%8)*16 + (k%4)*4,   LBO = 2048  (128 taps x 32 pixels per stage).
// Lanes = 8 consecutive taps x 4 pixel chunks: the stores of a quarter warp are 128 contiguous bytes, and the loads hit
// 32 different banks when the 16 pixels share an image row (conv_tc.cuh's wgrad kernel scatters sixteen 4-byte stores
// per thread and k-block and gathers from L2).
//
// Persistent, one CTA per SM, tiles of R <= 128 consecutive output pixels (as in conv_st.cuh):
//   copy warp      : lane = input channel: bulk copies of the rows a tile needs (fp32 activations, or uint8 frames of the
//                    sampled ring rows); lane = output channel: bulk copies of the tile's run of G (16-byte aligned supersets)
//   producer warps : 16 warps, per (128-tap tile, 32-pixel block): 8 taps x 32 pixels each (two items of 8 taps x 16 pixels),
//                    tf32 hi/lo split (frames: exact integers, one part)
//   G warps (4)    : G runs -> K-major hi/lo tiles of every pixel block of the tile (double buffered across tiles),
//                    bias-gradient sums; after the last tile TMEM -> partial[cta][tap][co]
//   MMA warp       : per stage 4 pixel groups x (A_hi.G_hi, A_lo.G_hi, A_hi.G_lo) into the tap tile's TMEM accumulator,
//                    accumulating over ALL tiles of the CTA: one partial per CTA, reduced in fixed order by
//                    splitk_reduce_kernel (deterministic, no atomics).
#pragma once
#include "conv_st.cuh"

namespace b2rl {

constexpr int kWsProdWarps = 16, kWsGWarps = 4;
constexpr int kWsThreads = (kWsProdWarps + 2 + kWsGWarps) * 32;
constexpr int kWsStages = 2;
constexpr int kWsMaxSeg = 4;
constexpr uint32_t kWsLbo = kTcBM * 16;                 // K-major im2col^T stage: 8 pixel chunks x (128 taps x 16 B)
constexpr uint32_t kWsAPart = 8 * kWsLbo;               // 16 KB: 128 taps x 32 pixels

struct ConvWstParams {
    const void *x;               // fp32 activations [rows, Cin, H, W] or uint8 frames (ring), NCHW
    const int64_t *gather;       // ring row of every batch row (frames) or NULL
    const float *g;              // dL/d(layer output) [rows, N, P]
    float *partial;              // [ctas][Mtaps + 1][N]
    int64_t in_bstride;          // Cin*H*W elements
    int M, N, n_pad, Mtaps;      // pixels, Cout, padded Cout, Cin*k*k
    int P, OW, S, Cin, H, W;
    int R, n_tiles;
    uint32_t chan_bytes;         // slab bytes per input channel
    uint32_t gslot;              // floats per (segment, channel) run slot of the G slab
    int nseg_max;
    float low, scale;            // frames: value = byte - low, partial sums scaled by 1/(high-low)
};

struct ConvWstSmem {
    uint32_t a, gt, gslab, xslab, bsum, ptab, bars, total;
};
__host__ __device__ __forceinline__ ConvWstSmem conv_wst_carve(bool exact, int n_pad, int rows_p, int Cin, uint32_t chan_bytes, int nseg_max,
                                         int N, uint32_t gslot) {
    ConvWstSmem c;
    uint32_t o = 0;
    c.a = o; o += kWsStages * (exact ? 1u : 2u) * kWsAPart;
    c.gt = o; o += 2u * (uint32_t)(rows_p / 32) * 2u * (uint32_t)n_pad * 32u * 4u;
    c.gslab = o; o += (uint32_t)nseg_max * (uint32_t)N * gslot * 4u;
    c.xslab = o; o += (uint32_t)Cin * chan_bytes;
    c.bsum = o; o += kWsGWarps * 32u * 4u;
    c.ptab = o; o += (uint32_t)rows_p * 4u;
    c.bars = o; o += 8u * (2 * kWsStages + 2 + 2 + 2 + 1 + 2 * kStMaxCin) + 16u;
    c.total = o + 128u;
    return c;
}

// image `bimg` of tile [m0, m1): pixel range, first output row, input rows
__host__ __device__ __forceinline__ void wst_segment(int bimg, int m0, int m1, int P, int OW, int S, int KS, int &p_lo, int &p_hi,
                                                     int &oy_lo, int &n_in) {
    p_lo = m0 - bimg * P > 0 ? m0 - bimg * P : 0;
    p_hi = m1 - bimg * P < P ? m1 - bimg * P : P;
    oy_lo = p_lo / OW;
    n_in = ((p_hi - 1) / OW - oy_lo) * S + KS;
}

template <int KS, bool U8>
__global__ void __launch_bounds__(kWsThreads, 1) conv_wgrad_st_kernel(const ConvWstParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int KK = KS * KS;
    constexpr bool EXACT = U8;
    constexpr uint32_t ESZ = U8 ? 1u : 4u;
    constexpr uint32_t a_stage = (EXACT ? 1u : 2u) * kWsAPart;
    constexpr int kChanWarps = KK / 8;                                              // warps (8 taps each) that read one channel per tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows_p = (p.R + 31) / 32 * 32, npb_max = rows_p / 32;
    const int MT = (p.Mtaps + kTcBM - 1) / kTcBM;
    const ConvWstSmem cv = conv_wst_carve(EXACT, p.n_pad, rows_p, p.Cin, p.chan_bytes, p.nseg_max, p.N, p.gslot);
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t a_s = sbase + cv.a, gt_s = sbase + cv.gt, gslab_s = sbase + cv.gslab, xslab_s = sbase + cv.xslab;
    const uint32_t bsum_s = sbase + cv.bsum, ptab_s = sbase + cv.ptab;
    uint8_t *gen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    uint64_t *full_a = reinterpret_cast<uint64_t *>(gen + cv.bars);      // [stages] (16 producer warps)
    uint64_t *empty_a = full_a + kWsStages;                               // [stages] (MMA commit)
    uint64_t *gt_full = empty_a + kWsStages;                              // [2] G tiles of a tile built (4 G warps)
    uint64_t *gt_empty = gt_full + 2;                                     // [2] consumed (MMA commit)
    uint64_t *gs_full = gt_empty + 2;                                     // [1] G slab landed; [1] G slab released (4 G warps)
    uint64_t *gs_empty = gs_full + 1;
    uint64_t *acc_done = gs_empty + 1;
    uint64_t *slab_full = acc_done + 1;                                   // [Cin]
    uint64_t *slab_empty = slab_full + kStMaxCin;                         // [Cin]
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(slab_empty + kStMaxCin);
    const uint32_t gt_part = (uint32_t)p.n_pad * 32u * 4u;                // hi (or lo) tile of one pixel block
    const uint32_t gt_buf = (uint32_t)npb_max * 2u * gt_part;             // all pixel blocks of one tile
    const uint32_t lbo_b = (uint32_t)p.n_pad * 16u;
    const uint32_t row_bytes = (uint32_t)p.W * ESZ;
    const int my_tiles = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < kWsStages; ++s) {
            tc::mbar_init(&full_a[s], kWsProdWarps);
            tc::mbar_init(&empty_a[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&gt_full[s], kWsGWarps);
            tc::mbar_init(&gt_empty[s], 1);
        }
        tc::mbar_init(gs_full, 1);
        tc::mbar_init(gs_empty, kWsGWarps);
        tc::mbar_init(acc_done, 1);
        for (int c = 0; c < p.Cin; ++c) {
            tc::mbar_init(&slab_full[c], 1);
            tc::mbar_init(&slab_empty[c], kChanWarps);
        }
        tc::fence_barrier_init();
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < MT * p.n_pad) tmem_cols <<= 1;
    if (warp == kWsProdWarps) tc::tmem_alloc(tmem_ptr, tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    auto tile_range = [&](int i, int &m0, int &m1) {
        const int t = (int)blockIdx.x + i * (int)gridDim.x;
        m0 = t * p.R;
        m1 = (m0 + p.R < p.M) ? m0 + p.R : p.M;
    };

    if (warp == kWsProdWarps + 1) {
        // ================================ copy warp ================================
        for (int i = 0; i < my_tiles; ++i) {
            int m0, m1;
            tile_range(i, m0, m1);
            const int b_first = m0 / p.P, b_last = (m1 - 1) / p.P;
            // ---- G runs: lane = output channel
            if (i > 0) tc::mbar_wait(gs_empty, (uint32_t)((i - 1) & 1));
            {
                uint32_t bytes = 0;
                for (int co = lane; co < p.N; co += 32)
                    for (int b = b_first; b <= b_last; ++b) {
                        int p_lo, p_hi, oy_lo, n_in;
                        wst_segment(b, m0, m1, p.P, p.OW, p.S, KS, p_lo, p_hi, oy_lo, n_in);
                        const int64_t e0 = ((int64_t)b * p.N + co) * p.P + p_lo;
                        bytes += (uint32_t)((((e0 + (p_hi - p_lo) + 3) & ~(int64_t)3) - (e0 & ~(int64_t)3)) * 4);
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
                if (lane == 0) tc::mbar_expect_tx(gs_full, bytes);
                __syncwarp();
                for (int co = lane; co < p.N; co += 32)
                    for (int b = b_first; b <= b_last; ++b) {
                        int p_lo, p_hi, oy_lo, n_in;
                        wst_segment(b, m0, m1, p.P, p.OW, p.S, KS, p_lo, p_hi, oy_lo, n_in);
                        const int64_t e0 = ((int64_t)b * p.N + co) * p.P + p_lo;
                        const int64_t a0 = e0 & ~(int64_t)3, a1 = (e0 + (p_hi - p_lo) + 3) & ~(int64_t)3;
                        const uint32_t dst = gslab_s + (uint32_t)(((b - b_first) * p.N + co) * (int)p.gslot) * 4u;
                        tc::bulk_g2s(dst, p.g + a0, (uint32_t)(a1 - a0) * 4u, gs_full);
                    }
            }
            // ---- input rows: lane = input channel
            for (int c = lane; c < p.Cin; c += 32) {
                if (i > 0) tc::mbar_wait(&slab_empty[c], (uint32_t)((i - 1) & 1));
                uint32_t bytes = 0;
                for (int b = b_first; b <= b_last; ++b) {
                    int p_lo, p_hi, oy_lo, n_in;
                    wst_segment(b, m0, m1, p.P, p.OW, p.S, KS, p_lo, p_hi, oy_lo, n_in);
                    const uint32_t st = (uint32_t)(oy_lo * p.S) * row_bytes, en = st + (uint32_t)n_in * row_bytes;
                    bytes += ((en + 15u) & ~15u) - (st & ~15u);
                }
                tc::mbar_expect_tx(&slab_full[c], bytes);
                uint32_t dst = xslab_s + (uint32_t)c * p.chan_bytes;
                for (int b = b_first; b <= b_last; ++b) {
                    int p_lo, p_hi, oy_lo, n_in;
                    wst_segment(b, m0, m1, p.P, p.OW, p.S, KS, p_lo, p_hi, oy_lo, n_in);
                    const uint32_t st = (uint32_t)(oy_lo * p.S) * row_bytes, en = st + (uint32_t)n_in * row_bytes;
                    const uint32_t a0 = st & ~15u, a1 = (en + 15u) & ~15u;
                    const int64_t bb = p.gather ? __ldg(p.gather + b) : (int64_t)b;
                    const uint8_t *src = static_cast<const uint8_t *>(p.x) + (bb * p.in_bstride + (int64_t)c * p.H * p.W) * ESZ + a0;
                    tc::bulk_g2s(dst, src, a1 - a0, &slab_full[c]);
                    dst += a1 - a0;
                }
            }
            __syncwarp();
        }
    } else if (warp == kWsProdWarps) {
        // ================================ MMA warp ================================
        const uint32_t idesc = tc::make_idesc_tf32(kTcBM, p.n_pad);
        const uint64_t da_step = (uint64_t)((2 * kWsLbo) >> 4), db_step = (uint64_t)((2 * lbo_b) >> 4);
        int stage = 0;
        uint32_t sph = 0;
        for (int i = 0; i < my_tiles; ++i) {
            int m0, m1;
            tile_range(i, m0, m1);
            const int npb = (m1 - m0 + 31) / 32;
            const int buf = i & 1;
            tc::mbar_wait(&gt_full[buf], (uint32_t)((i >> 1) & 1));
            for (int tt = 0; tt < MT; ++tt) {
                const uint32_t d_addr = tmem_d + (uint32_t)(tt * p.n_pad);
                for (int pb = 0; pb < npb; ++pb) {
                    tc::mbar_wait(&full_a[stage], sph);
                    tc::tc_fence_after();
                    const uint32_t a_addr = a_s + (uint32_t)stage * a_stage;
                    const uint32_t g_addr = gt_s + (uint32_t)buf * gt_buf + (uint32_t)pb * 2u * gt_part;
                    const uint64_t dah0 = tc::make_desc(a_addr, kWsLbo, 128), dal0 = tc::make_desc(a_addr + kWsAPart, kWsLbo, 128);
                    const uint64_t dbh0 = tc::make_desc(g_addr, lbo_b, 128), dbl0 = tc::make_desc(g_addr + gt_part, lbo_b, 128);
                    if (tc::elect_one()) {
#pragma unroll
                        for (int pg = 0; pg < 4; ++pg) {                       // 8 pixels per MMA
                            const uint64_t ah = dah0 + pg * da_step, al = dal0 + pg * da_step;
                            tc::mma_tf32(d_addr, ah, dbh0 + pg * db_step, idesc, (i | pb | pg) ? 1u : 0u);
                            if (!EXACT) tc::mma_tf32(d_addr, al, dbh0 + pg * db_step, idesc, 1u);
                            tc::mma_tf32(d_addr, ah, dbl0 + pg * db_step, idesc, 1u);
                        }
                        tc::mma_commit(&empty_a[stage]);
                    }
                    __syncwarp();
                    if (++stage == kWsStages) { stage = 0; sph ^= 1u; }
                }
            }
            if (tc::elect_one()) {
                tc::mma_commit(&gt_empty[buf]);
                if (i == my_tiles - 1) tc::mma_commit(acc_done);
            }
            __syncwarp();
        }
    } else if (warp < kWsProdWarps) {
        // ================================ producer warps ================================
        int stage = 0;
        uint32_t sph = 1;
        const float exact_bias = 8388608.f + p.low;
        const int pt = tid;                                            // 0..511: pixel-table slot this thread fills
        const int t8 = lane & 7, qq = lane >> 3;                       // tap within the warp's octet, pixel chunk within a quad
        for (int i = 0; i < my_tiles; ++i) {
            int m0, m1;
            tile_range(i, m0, m1);
            const int npb = (m1 - m0 + 31) / 32;
            const uint32_t tph = (uint32_t)(i & 1);
            const int b_first = m0 / p.P;
            // ---- slab address of the first element of every pixel's receptive field (pixels past the tile repeat its last one)
            asm volatile("bar.sync 2, %0;" ::"r"(kWsProdWarps * 32) : "memory");      // the previous tile's table is no longer read
            if (pt < npb * 32) {
                const int m = (m0 + pt < m1) ? m0 + pt : m1 - 1;
                const int b = m / p.P, pix = m - b * p.P;
                const int oy = pix / p.OW, ox = pix - oy * p.OW;
                uint32_t seg_off = 0;
                int oy_lo = 0;
                for (int bb = b_first; bb <= b; ++bb) {
                    int p_lo, p_hi, n_in;
                    wst_segment(bb, m0, m1, p.P, p.OW, p.S, KS, p_lo, p_hi, oy_lo, n_in);
                    const uint32_t st = (uint32_t)(oy_lo * p.S) * row_bytes, en = st + (uint32_t)n_in * row_bytes;
                    if (bb < b) seg_off += ((en + 15u) & ~15u) - (st & ~15u);
                    else seg_off += st & 15u;
                }
                const uint32_t src = xslab_s + seg_off + (uint32_t)((oy - oy_lo) * p.S) * row_bytes + (uint32_t)(ox * p.S) * ESZ;
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(ptab_s + 4u * pt), "r"(src) : "memory");
            }
            asm volatile("bar.sync 2, %0;" ::"r"(kWsProdWarps * 32) : "memory");
            for (int tt = 0; tt < MT; ++tt) {
                const int tap = tt * kTcBM + warp * 8 + t8;                // this lane's tap; the warp's 8 taps share a channel
                const bool ok = tt * kTcBM + warp * 8 < p.Mtaps;           // warp-uniform (Mtaps is a multiple of 16)
                int c;
                uint32_t off;
                if (KS == 4) { c = tap >> 4; off = (uint32_t)((tap >> 2) & 3) * row_bytes + (uint32_t)(tap & 3) * ESZ; }
                else { c = tap >> 6; off = (uint32_t)((tap >> 3) & 7) * row_bytes + (uint32_t)(tap & 7) * ESZ; }
                off += (uint32_t)c * p.chan_bytes;
                if (ok) tc::mbar_wait(&slab_full[c], tph);
                const uint32_t dst_t = (uint32_t)warp * 128u + (uint32_t)t8 * 16u;
                for (int pb = 0; pb < npb; ++pb) {
                    float v[2][4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {                          // pixel chunks qq and 4 + qq of the block
                        uint32_t a0, a1, a2, a3;
                        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3)
                                     : "r"(ptab_s + 4u * (uint32_t)(pb * 32 + (h * 4 + qq) * 4)));
                        if (!ok) { v[h][0] = v[h][1] = v[h][2] = v[h][3] = 0.f; continue; }
                        if (U8) {
                            uint32_t b0, b1, b2, b3;
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b0) : "r"(a0 + off));
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b1) : "r"(a1 + off));
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b2) : "r"(a2 + off));
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b3) : "r"(a3 + off));
                            v[h][0] = __uint_as_float(b0 | 0x4B000000u) - exact_bias;      // (2^23 + byte) - (2^23 + low): exact
                            v[h][1] = __uint_as_float(b1 | 0x4B000000u) - exact_bias;
                            v[h][2] = __uint_as_float(b2 | 0x4B000000u) - exact_bias;
                            v[h][3] = __uint_as_float(b3 | 0x4B000000u) - exact_bias;
                        } else {
                            v[h][0] = __uint_as_float(tc::lds32(a0 + off));
                            v[h][1] = __uint_as_float(tc::lds32(a1 + off));
                            v[h][2] = __uint_as_float(tc::lds32(a2 + off));
                            v[h][3] = __uint_as_float(tc::lds32(a3 + off));
                        }
                    }
                    tc::mbar_wait(&empty_a[stage], sph);
                    const uint32_t dst = a_s + (uint32_t)stage * a_stage + dst_t + (uint32_t)qq * kWsLbo;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t d = dst + (uint32_t)(h * 4) * kWsLbo;
                        if (EXACT) {
                            tc::sts128(d, v[h][0], v[h][1], v[h][2], v[h][3]);
                        } else {
                            float hi[4], lo[4];
#pragma unroll
                            for (int j = 0; j < 4; ++j) { hi[j] = tc::tf32_rn_fast(v[h][j]); lo[j] = v[h][j] - hi[j]; }
                            tc::sts128(d, hi[0], hi[1], hi[2], hi[3]);
                            tc::sts128(d + kWsAPart, lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                    tc::fence_async_smem();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(&full_a[stage]);
                    if (++stage == kWsStages) { stage = 0; sph ^= 1u; }
                }
                if (lane == 0 && ok) tc::mbar_arrive(&slab_empty[c]);
            }
        }
    } else {
        // ================================ G warps ================================
        const int gt = tid - (kWsProdWarps + 2) * 32;                  // 0..127
        const int co = gt % p.n_pad, chunk0 = gt / p.n_pad, chunk_step = (kWsGWarps * 32) / p.n_pad;
        float bsum = 0.f;
        for (int i = 0; i < my_tiles; ++i) {
            int m0, m1;
            tile_range(i, m0, m1);
            const int npb = (m1 - m0 + 31) / 32;
            const int buf = i & 1;
            const int b_first = m0 / p.P;
            tc::mbar_wait(gs_full, (uint32_t)(i & 1));
            tc::mbar_wait(&gt_empty[buf], (uint32_t)(((i >> 1) & 1) ^ 1));      // passes at first use
            for (int chunk = chunk0; chunk < npb * 8; chunk += chunk_step) {     // 4 consecutive pixels of channel co
                const int pb = chunk >> 3, q = chunk & 7;
                float v[4];
                int m = m0 + chunk * 4;
                int b = m / p.P, pp = m - b * p.P;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    v[j] = 0.f;
                    if (co < p.N && m + j < m1) {
                        const int p_lo = (b == b_first) ? m0 - b_first * p.P : 0;
                        const int64_t e0 = ((int64_t)b * p.N + co) * p.P + p_lo;
                        const uint32_t a = gslab_s + ((uint32_t)(((b - b_first) * p.N + co) * (int)p.gslot) + (uint32_t)(e0 & 3) + (uint32_t)(pp - p_lo)) * 4u;
                        v[j] = __uint_as_float(tc::lds32(a));
                    }
                    if (++pp == p.P) { pp = 0; ++b; }
                }
                float h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { h[j] = tc::tf32_rn_fast(v[j]); l[j] = v[j] - h[j]; bsum += v[j]; }
                const uint32_t o = gt_s + (uint32_t)buf * gt_buf + (uint32_t)pb * 2u * gt_part + (uint32_t)q * lbo_b +
                                   (uint32_t)(co >> 3) * 128u + (uint32_t)(co & 7) * 16u;
                tc::sts128(o, h[0], h[1], h[2], h[3]);
                tc::sts128(o + gt_part, l[0], l[1], l[2], l[3]);
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(&gt_full[buf]);
                tc::mbar_arrive(gs_empty);
            }
        }
        // ---- bias gradient of this CTA: fixed-order sum of the threads that share a channel
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bsum_s + 4u * gt), "f"(bsum) : "memory");
        asm volatile("bar.sync 1, %0;" ::"r"(kWsGWarps * 32) : "memory");
        float *part = p.partial + (int64_t)blockIdx.x * (p.Mtaps + 1) * p.N;
        if (gt < p.N) {
            float t = 0.f;
            for (int j = gt; j < kWsGWarps * 32; j += p.n_pad) t += __uint_as_float(tc::lds32(bsum_s + 4u * j));
            part[(int64_t)p.Mtaps * p.N + gt] = t;
        }
        // ---- epilogue: TMEM lanes 32*(warp%4).. = taps of the tile
        if (my_tiles > 0) tc::mbar_wait(acc_done, 0);
        tc::tc_fence_after();
        const int q = warp & 3;
        for (int tt = 0; tt < MT; ++tt) {
            const int tap = tt * kTcBM + q * 32 + lane;
            for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
                uint32_t r[16];
                if (my_tiles > 0) tc::tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(tt * p.n_pad + c0), r);
                if (tap < p.Mtaps) {
                    float *o = part + (int64_t)tap * p.N + c0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < p.N) o[j] = my_tiles > 0 ? __uint_as_float(r[j]) * p.scale : 0.f;
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kWsProdWarps) tc::tmem_dealloc(tmem_d, tmem_cols);
}

static inline size_t conv_wgrad_st_partial_floats(const b2rl_layer &l, int sms) {
    return (size_t)sms * ((size_t)l.in_c * l.ksize * l.ksize + 1) * l.out_c;
}

// returns B2RL_OK, or 1 when the layer is outside what this kernel handles (caller: conv_tc.cuh's wgrad kernel)
static int launch_conv_wgrad_st(const b2rl_layer &l, const Operand &X, const float *g, float *dw, float *db, int accumulate,
                                int64_t rows, float *partial, size_t partial_cap, cudaStream_t s) {
    if (!st_enabled('w')) return 1;
    const int KS = l.ksize, KK = KS * KS, Kc = l.in_c * KK, P = l.out_h * l.out_w;
    const int n_pad = (l.out_c + 15) / 16 * 16, sms = sm_count();
    const bool u8 = X.u8;
    if (!(KS == 4 || KS == 8) || n_pad > 64 || (kWsGWarps * 32) % n_pad != 0 || l.out_c % 4 != 0 || l.in_c > kStMaxCin) return 1;
    if (X.elem_kind() == EL_F32_NORM) return 1;
    if (u8 && X.normalize && !(X.low == floorf(X.low) && fabsf(X.low) <= 1024.f && X.high > X.low)) return 1;
    if (u8 ? (l.stride % 4 != 0 || l.in_w % 4 != 0 || (l.in_h * l.in_w) % 16 != 0)
           : (l.stride % 2 != 0 || l.in_w % 4 != 0 || (l.in_h * l.in_w) % 4 != 0)) return 1;
    if (!u8 && X.red.gather != nullptr) return 1;
    if (reinterpret_cast<uintptr_t>(X.ptr) % 16 != 0 || reinterpret_cast<uintptr_t>(g) % 16 != 0) return 1;
    if (rows * (int64_t)P > INT32_MAX || rows < 1 || partial == nullptr) return 1;
    const int MT = (Kc + kTcBM - 1) / kTcBM;
    if (MT * n_pad > 512) return 1;
    const int M = (int)(rows * P);
    const uint32_t esz = u8 ? 1u : 4u, row_bytes = (uint32_t)l.in_w * esz;
    int R = 0, n_tiles = 0, nseg_max = 0;
    uint32_t chan_bytes = 0, gslot = 0;
    ConvWstSmem cv{};
    for (int w = (M + sms * kTcBM - 1) / (sms * kTcBM); w <= 64; ++w) {
        R = (M + sms * w - 1) / (sms * w);
        if (R < 1) R = 1;
        n_tiles = (M + R - 1) / R;
        chan_bytes = 0; nseg_max = 0;
        for (int t = 0; t < n_tiles; ++t) {
            const int m0 = t * R, m1 = (m0 + R < M) ? m0 + R : M;
            const int b_first = m0 / P, b_last = (m1 - 1) / P;
            if (b_last - b_first + 1 > nseg_max) nseg_max = b_last - b_first + 1;
            uint32_t tot = 0;
            for (int b = b_first; b <= b_last; ++b) {
                int p_lo, p_hi, oy_lo, n_in;
                wst_segment(b, m0, m1, P, l.out_w, l.stride, KS, p_lo, p_hi, oy_lo, n_in);
                const uint32_t st = (uint32_t)(oy_lo * l.stride) * row_bytes, en = st + (uint32_t)n_in * row_bytes;
                tot += ((en + 15u) & ~15u) - (st & ~15u);
            }
            if (tot > chan_bytes) chan_bytes = tot;
        }
        gslot = (uint32_t)((R + 6 + 3) & ~3);
        cv = conv_wst_carve(u8, n_pad, (R + 31) / 32 * 32, l.in_c, chan_bytes, nseg_max, l.out_c, gslot);
        if (nseg_max <= kWsMaxSeg && cv.total <= (uint32_t)kStSmemMax) break;
        if (R <= 8) return 1;
        R = 0;
    }
    if (R == 0) return 1;
    const int grid = n_tiles < sms ? n_tiles : sms;
    if ((size_t)grid * (Kc + 1) * l.out_c > partial_cap) return 1;
    ConvWstParams p;
    p.x = X.ptr; p.gather = X.red.gather; p.g = g; p.partial = partial;
    p.in_bstride = (int64_t)l.in_c * l.in_h * l.in_w;
    p.M = M; p.N = l.out_c; p.n_pad = n_pad; p.Mtaps = Kc;
    p.P = P; p.OW = l.out_w; p.S = l.stride; p.Cin = l.in_c; p.H = l.in_h; p.W = l.in_w;
    p.R = R; p.n_tiles = n_tiles; p.chan_bytes = chan_bytes; p.gslot = gslot; p.nseg_max = nseg_max;
    p.low = (u8 && X.normalize) ? X.low : 0.f;
    p.scale = (u8 && X.normalize) ? 1.0f / (X.high - X.low) : 1.0f;
    auto launch = [&](auto kern, int slot) -> int {
        static bool attr_set[4] = {false, false, false, false};
        if (!attr_set[slot]) {
            B2RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStSmemMax));
            attr_set[slot] = true;
        }
        kern<<<grid, kWsThreads, cv.total, s>>>(p);
        B2RL_LAUNCH_CHECK();
        ++g_conv_path[2];
        return B2RL_OK;
    };
    int rc;
    if (u8) rc = KS == 4 ? launch(conv_wgrad_st_kernel<4, true>, 0) : launch(conv_wgrad_st_kernel<8, true>, 1);
    else rc = KS == 4 ? launch(conv_wgrad_st_kernel<4, false>, 2) : launch(conv_wgrad_st_kernel<8, false>, 3);
    if (rc != B2RL_OK) return rc;
    Epilogue epi;
    epi.kind = EPI_WGRAD_T; epi.out = dw; epi.db = db; epi.wcols = Kc; epi.accumulate = accumulate;
    launch_splitk_reduce<EpiTraits<EPI_WGRAD_T, MAP_STRIDE, MAP_STRIDE>>(partial, grid, Kc + 1, l.out_c, epi, s);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

}  // namespace b2rl
