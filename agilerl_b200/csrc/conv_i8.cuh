// conv_i8.cuh — first-layer convolution over uint8 frames on the INTEGER tensor path (tcgen05.mma kind::i8).
//
// Replaces, for uint8 observations with integer bounds (Box(0, 255, uint8) image spaces), the first conv2d of
// EvolvableCNN.forward (agilerl/modules/cnn.py:552-580) together with RLAlgorithm.preprocess_observation's
// (x - low) / (high - low) (agilerl/utils/algo_utils.py:1131-1180).
//
// Why integers.  The tf32 kernel of conv_tc.cuh spends its time turning bytes into floats: per tap a PRMT, an FADD
// and a quarter of a 16-byte shared-memory store (8.7 M warp instructions for B = 256, issue-bound at 10 % of the HBM
// roof).  The tensor cores of sm_100a multiply u8 x s8 natively, so the im2col operand can be the frame bytes
// themselves: a thread's work per 16 taps is four 4-byte loads and ONE 16-byte store, no arithmetic at all.
// The weights carry the precision instead: per output channel c, w[c, :] / 2^e_c (2^e_c > max|w[c, :]|) is rounded
// to a 31-bit fixed-point integer q and written as four balanced base-256 digits d0..d3 in [-128, 127]
// (q = d0 2^24 + d1 2^16 + d2 2^8 + d3), so
//     sum_k (x_k - low) w[c,k]  =  2^(e_c - 30) * ( D0 2^24 + D1 2^16 + D2 2^8 + D3 ),   Dj = sum_k x_k dj[c,k]  (int32, EXACT)
// — ONE MMA with N = 4 * Cout_pad multiplies a k-step of 32 taps with all four digit planes; the four int32
// accumulator columns of a channel are recombined in fp32 in the epilogue, scaled by 2^(e_c-30) / (high - low),
// biased and activated.  The only rounding is the 2^-30 quantisation of w relative to its channel maximum
// (<= 2^-29 max|w_c| per weight) and the fp32 recombination (<= 2^-23 relative): tighter than the fp32 dot
// product of the reference; parity gate 1e-5 on losses / 2e-5 max|g| on gradients (tests/test_learn_gpu.py).
//
// Shared-memory operand layout (K-major, no swizzle, 8-bit elements: a core matrix is 8 rows x 16 taps = 128 B):
//     element (row r, tap k) at  (k/16)*LBO + (r/8)*128 + (r%8)*16 + (k%16),   LBO = rows*16
// The whole K extent of both operands is resident (K = 256 for the north-star layer: A 32 KB, B 32 KB per net),
// so there is no pipeline ring: gather everything, one fence, k_pad/32 MMAs, one commit.  Latency is hidden by
// 2-3 co-resident CTAs per SM, not by stages.
//
// Two weight sets ("nets" = 2) share one gather: the online and the target network both read next_obs in
// _dqn_loss (dqn_rainbow.py:306-318), so their first layers run as ONE MMA stream with N = 8 * Cout_pad = 256
// accumulator columns over the same im2col tile.
#pragma once
#include "conv_tc.cuh"

namespace b2rl {

constexpr int kI8MaxSegs = 2;

struct ConvI8Seg {
    const uint8_t *x;          // frames (ring base when gather != NULL)
    const int64_t *gather;     // optional ring rows per batch row
    float *out[2];             // NCHW outputs, one per weight set used by this segment
    int M;                     // pixels = rows * P
    int nets;                  // 1: first weight set only, 2: both
    int cta0;                  // first CTA of this segment
};

struct ConvI8Params {
    ConvI8Seg seg[kI8MaxSegs];
    int n_seg;
    const int8_t *wd;          // digit planes of both weight sets in the smem tile layout: rows = net*4*n_pad + digit*n_pad + n
    const float *scale;        // [nets_total * n_pad]   2^(e_c - 30) / (high - low)
    const float *bias[2];      // [Cout] per weight set
    int64_t in_bstride;        // Cin*H*W
    int N, n_pad, k_pad;       // Cout, Cout rounded up to 16, Cin*k*k rounded up to 32
    int nets_total;            // weight sets resident in wd (1 or 2)
    int P, OW, sy, sx;         // output pixels per image, output width, in-row stride (s*W), in-col stride (s)
    int act;
    int Cin, HW, W;            // input channels, plane size, row pitch (the gather walks kernel rows with running pointers)
    // staged mode (slab != 0): the rows of every channel a tile needs are bulk-copied into shared memory by a copy warp
    // (16-byte aligned supersets, three tiles in flight) and the gather warps read them with LDS
    int slab, KS, S;
    int nseg_max;              // images a tile may touch
    uint32_t chan_bytes;       // slab bytes per channel (all segments of a tile)
};

constexpr int kI8Slabs = 3;
static inline size_t conv_i8_smem_bytes(int n_pad, int k_pad, int nets_total, int Cin = 0, uint32_t chan_bytes = 0) {
    return 2 * (size_t)kTcBM * k_pad + (size_t)nets_total * 4 * n_pad * k_pad + 2 * (size_t)nets_total * n_pad * 4 + 256 + 1024 +
           (size_t)kI8Slabs * Cin * chan_bytes;
}
// image `bimg` of tile [m0, m1): first output row it needs and the byte range [st, en) of every channel plane
__host__ __device__ __forceinline__ void i8_segment(int bimg, int m0, int m1, int P, int OW, int S, int KS, int W, int &oy_lo,
                                                    uint32_t &st, uint32_t &en) {
    const int p_lo = m0 - bimg * P > 0 ? m0 - bimg * P : 0;
    const int p_hi = m1 - bimg * P < P ? m1 - bimg * P : P;
    oy_lo = p_lo / OW;
    const int n_in = ((p_hi - 1) / OW - oy_lo) * S + KS;
    st = (uint32_t)(oy_lo * S * W);
    en = st + (uint32_t)(n_in * W);
}
// scratch the launcher needs (bytes): digit planes + scales + low-correction + tap-group offsets
static inline size_t conv_i8_scratch_bytes(int n_pad, int k_pad, int nets_total) {
    return (size_t)nets_total * 4 * n_pad * k_pad + (size_t)nets_total * n_pad * 8 + (size_t)k_pad + 256;
}

// One CTA per (output channel, weight set): channel maximum -> power-of-two scale -> four balanced int8 digits per
// weight, stored where the conv kernel's bulk copy expects them.  Channel n of set `net` is row net*4*n_pad + d*n_pad + n.
__global__ void weight_digits_kernel(const float *__restrict__ w0, const float *__restrict__ w1, int N, int K, int n_pad,
                                     int k_pad, int nets_total, double inv_range, float low, int8_t *__restrict__ wd,
                                     float *__restrict__ scale, float *__restrict__ lowcorr) {
    const int n = blockIdx.x, net = blockIdx.y;
    const float *w = net == 0 ? w0 : w1;
    __shared__ float red[32];
    __shared__ double redd[32];
    float mx = 0.f;
    if (n < N)
        for (int k = threadIdx.x; k < K; k += blockDim.x) mx = fmaxf(mx, fabsf(w[(int64_t)n * K + k]));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
    int e = 0;
    if (mx > 0.f && mx < INFINITY) frexpf(mx, &e);                    // mx = m 2^e, m in [0.5, 1)  ->  2^e > mx
    const double up = ldexp(1.0, 30 - e);                            // exact power of two
    const int rows_total = nets_total * 4 * n_pad;
    const uint32_t lbo = (uint32_t)rows_total * 16;
    double qsum = 0.0;
    for (int k = threadIdx.x; k < k_pad; k += blockDim.x) {
        long long q = 0;
        if (n < N && k < K && mx > 0.f && mx < INFINITY) q = llrint((double)w[(int64_t)n * K + k] * up);   // |q| <= 2^30
        qsum += (double)q;
        int d[4];
#pragma unroll
        for (int i = 3; i >= 1; --i) {                               // balanced digits, least significant first
            const int lowb = (int)(((q + 128) & 255) - 128);
            d[i] = lowb;
            q = (q - lowb) >> 8;
        }
        d[0] = (int)q;                                               // |d0| <= 65
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = net * 4 * n_pad + i * n_pad + n;
            wd[(size_t)(k >> 4) * lbo + (size_t)(r >> 3) * 128 + (r & 7) * 16 + (k & 15)] = (int8_t)d[i];
        }
    }
    // low * sum_k w_q[c,k]: the integer accumulators see raw bytes x, the layer wants (x - low)
    for (int o = 16; o > 0; o >>= 1) qsum += __shfl_xor_sync(0xffffffffu, qsum, o);
    if ((threadIdx.x & 31) == 0) redd[threadIdx.x >> 5] = qsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += redd[i];
        const float sc = (mx > 0.f && mx < INFINITY) ? (float)(ldexp(1.0, e - 30) * inv_range) : 0.f;
        scale[net * n_pad + n] = sc;
        lowcorr[net * n_pad + n] = (float)(-(double)low * t * ldexp(1.0, e - 30) * inv_range);
    }
}

namespace tc {
// cute::UMMA::InstrDescriptor for kind::i8: c_format[4,6) = 2 (S32) | a_format[7,10) = 0 (U8) | b_format[10,13) = 1 (S8)
// | K-major A/B | n_dim[17,23) = N>>3 | m_dim[24,29) = M>>4
__device__ __forceinline__ uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
}  // namespace tc

// Persistent, warp-specialised kernel: one CTA per SM walks its tiles (128 output pixels each) through a three-stage
// pipeline whose stages run concurrently on different warps —
//   gather warps (8)   : frame bytes -> A stage s (two stages), running pointers instead of an offset table:
//                        per 16 taps four 4-byte loads, two pointer bumps, one 16-byte store;
//   MMA warp (1)       : k_pad/32 tcgen05.mma kind::i8 into TMEM accumulator s (two accumulators), commits free the
//                        A stage and hand the accumulator over;
//   epilogue warps (8) : TMEM -> digit recombination -> scale, bias, activation -> coalesced NCHW stores.
// The digit planes of every weight set are copied into shared memory ONCE per CTA.  KS = kernel size (4 or 8, so
// a kernel row is one or two aligned 4-byte groups), CPT = 16-tap chunks per gather thread (k_pad / 32).
constexpr int kI8GatherWarps = 8, kI8EpiWarps = 8;
constexpr int kI8ThreadsP = (kI8GatherWarps + 1 + kI8EpiWarps + 1) * 32;      // + the slab copy warp (idle when slab == 0)

template <int KS, int CPT>
__global__ void __launch_bounds__(kI8ThreadsP, 1) conv_fwd_i8_kernel(const ConvI8Params p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows_b = p.nets_total * 4 * p.n_pad;                       // rows of the resident B tile
    const int acc_cols = rows_b;                                         // columns of one accumulator buffer
    const uint32_t a_bytes = (uint32_t)kTcBM * p.k_pad, b_bytes = (uint32_t)rows_b * p.k_pad;
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t a_s = sbase, b_s = sbase + 2 * a_bytes;
    const uint32_t scale_a = b_s + b_bytes;                             // nets_total*n_pad floats (16-byte aligned)
    const uint32_t bias_a = scale_a + (uint32_t)p.nets_total * p.n_pad * 4;
    const uint32_t bars_a = (bias_a + (uint32_t)p.nets_total * p.n_pad * 4 + 15u) & ~15u;
    uint8_t *gen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    uint64_t *full_a = reinterpret_cast<uint64_t *>(gen + (bars_a - sbase));   // [2] im2col stage written (8 warp arrivals)
    uint64_t *empty_a = full_a + 2;                                             // [2] stage consumed (MMA commit)
    uint64_t *acc_full = full_a + 4;                                            // [2] accumulator complete (MMA commit)
    uint64_t *acc_empty = full_a + 6;                                           // [2] accumulator drained (8 warp arrivals)
    uint64_t *full_b = full_a + 8;                                              // digit planes landed (tx count)
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(full_a + 9);
    uint64_t *slab_full = full_a + 10;                                          // [kI8Slabs] rows of a tile landed (tx count)
    uint64_t *slab_empty = slab_full + kI8Slabs;                                // [kI8Slabs] released by the 8 gather warps
    const uint32_t slab_s = (bars_a + 8u * (10 + 2 * kI8Slabs) + 15u) & ~15u;
    const uint32_t slab_buf = (uint32_t)p.Cin * p.chan_bytes;
    const uint32_t lbo_a = kTcBM * 16, lbo_b = (uint32_t)rows_b * 16;
    // tiles of this CTA: t = blockIdx.x + i * gridDim.x over [segment 0 tiles | segment 1 tiles]
    const int tiles0 = (p.seg[0].M + kTcBM - 1) / kTcBM;
    const int tiles1 = p.n_seg > 1 ? (p.seg[1].M + kTcBM - 1) / kTcBM : 0;
    const int n_tiles = tiles0 + tiles1;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&full_a[s], kI8GatherWarps);
            tc::mbar_init(&empty_a[s], 1);
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], kI8EpiWarps);
        }
        tc::mbar_init(full_b, 1);
        for (int s = 0; s < kI8Slabs; ++s) {
            tc::mbar_init(&slab_full[s], 1);
            tc::mbar_init(&slab_empty[s], kI8GatherWarps);
        }
        tc::fence_barrier_init();
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < 2 * acc_cols) tmem_cols <<= 1;
    if (warp == kI8GatherWarps) tc::tmem_alloc(tmem_ptr, tmem_cols);
    for (int n = tid; n < p.nets_total * p.n_pad; n += kI8ThreadsP) {
        const int net = n / p.n_pad, c = n - net * p.n_pad;
        const float *bp = p.bias[net];
        // bias + the low-bound correction of this channel (the scale table holds [scale | lowcorr])
        const float bv = ((c < p.N && bp) ? bp[c] : 0.f) + __ldg(p.scale + p.nets_total * p.n_pad + n);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(scale_a + 4u * n), "f"(__ldg(p.scale + n)) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_a + 4u * n), "f"(bv) : "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    if (warp == kI8GatherWarps) {
        // ================================ MMA warp ================================
        if (tc::elect_one()) {
            tc::mbar_expect_tx(full_b, b_bytes);
            tc::bulk_g2s(b_s, p.wd, b_bytes, full_b);
        }
        __syncwarp();
        tc::mbar_wait(full_b, 0);
        const uint64_t db0 = tc::make_desc(b_s, lbo_b, 128);
        const uint64_t da_step = (uint64_t)((2 * lbo_a) >> 4), db_step = (uint64_t)((2 * lbo_b) >> 4);
        for (int i = 0; i < my_tiles; ++i) {
            const int t = (int)blockIdx.x + i * (int)gridDim.x;
            const int nets = t < tiles0 ? p.seg[0].nets : p.seg[1].nets;
            const int s = i & 1;
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            tc::mbar_wait(&acc_empty[s], ph ^ 1u);                   // epilogue has drained accumulator s (passes at first use)
            tc::mbar_wait(&full_a[s], ph);
            tc::tc_fence_after();
            const uint32_t idesc = tc::make_idesc_i8(kTcBM, nets * 4 * p.n_pad);   // first weight set = leading rows of every column
            const uint64_t da0 = tc::make_desc(a_s + (uint32_t)s * a_bytes, lbo_a, 128);
            const uint32_t d_addr = tmem_d + (uint32_t)(s * acc_cols);
            if (tc::elect_one()) {
#pragma unroll
                for (int j = 0; j < CPT; ++j)                         // one MMA k-step = 32 taps = 2 core-matrix columns
                    tc::mma_i8(d_addr, da0 + j * da_step, db0 + j * db_step, idesc, j ? 1u : 0u);
                tc::mma_commit(&empty_a[s]);
                tc::mma_commit(&acc_full[s]);
            }
            __syncwarp();
        }
    } else if (warp < kI8GatherWarps) {
        // ================================ gather warps ================================
        // thread = (im2col row, half of the taps): channels [half*Cin/2, (half+1)*Cin/2), every kernel row of them.
        // The loads of tile i+1 are issued BEFORE tile i's registers are stored: two tiles of DRAM latency in flight.
        const int row = tid & (kTcBM - 1), half = tid >> 7;
        const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16 + (uint32_t)(half * CPT) * lbo_a;
        const int64_t half_off = (int64_t)half * (p.Cin / 2) * p.HW;
        const int64_t step_row = p.W, step_chan = (int64_t)p.HW - (int64_t)(KS - 1) * p.W;
        auto issue = [&](int i, uint32_t (&raw)[CPT][4]) {
            const int t = (int)blockIdx.x + i * (int)gridDim.x;
            const ConvI8Seg &sg = t < tiles0 ? p.seg[0] : p.seg[1];
            const int m0 = (t < tiles0 ? t : t - tiles0) * kTcBM;
            const int m = m0 + row;
            if (p.slab) {
                // ---- staged: this row's receptive field inside slab i % kI8Slabs
                const int m1 = m0 + kTcBM < sg.M ? m0 + kTcBM : sg.M;
                const int mm = m < m1 ? m : m1 - 1;                  // rows beyond the tile repeat its last one
                const int b = mm / p.P, pix = mm - b * p.P;
                const int oy = pix / p.OW, ox = pix - oy * p.OW;
                uint32_t off = 0;
                int oy_lo = 0;
                for (int bb = m0 / p.P; bb <= b; ++bb) {
                    uint32_t st, en;
                    i8_segment(bb, m0, m1, p.P, p.OW, p.S, p.KS, p.W, oy_lo, st, en);
                    if (bb < b) off += ((en + 15u) & ~15u) - (st & ~15u);
                    else off += st & 15u;
                }
                const int buf = i % kI8Slabs;
                uint32_t ptr = slab_s + (uint32_t)buf * slab_buf + (uint32_t)(half * (p.Cin / 2)) * p.chan_bytes + off +
                               (uint32_t)((oy - oy_lo) * p.S * p.W + ox * p.S);
                const uint32_t s_row = (uint32_t)p.W, s_chan = p.chan_bytes - (uint32_t)((KS - 1) * p.W);
                tc::mbar_wait(&slab_full[buf], (uint32_t)((i / kI8Slabs) & 1));
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
#pragma unroll
                    for (int rr = 0; rr < 16 / KS; ++rr) {
                        const int r = c * (16 / KS) + rr;
#pragma unroll
                        for (int gx = 0; gx < KS / 4; ++gx) raw[c][rr * (KS / 4) + gx] = tc::lds32(ptr + 4u * gx);
                        ptr += (r % KS == KS - 1) ? s_chan : s_row;
                    }
                }
                return;
            }
            int64_t rowbase = sg.gather ? __ldg(sg.gather) * p.in_bstride : 0;     // rows beyond M read a valid address
            if (m < sg.M) {
                const int b = m / p.P, pix = m - b * p.P;
                const int oy = pix / p.OW, ox = pix - oy * p.OW;
                const int64_t bb = sg.gather ? __ldg(sg.gather + b) : (int64_t)b;
                rowbase = bb * p.in_bstride + (int64_t)(oy * p.sy + ox * p.sx);
            }
            const uint8_t *ptr = sg.x + rowbase + half_off;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
#pragma unroll
                for (int rr = 0; rr < 16 / KS; ++rr) {               // kernel rows inside this 16-tap chunk
                    const int r = c * (16 / KS) + rr;                // compile-time row counter of this thread
#pragma unroll
                    for (int gx = 0; gx < KS / 4; ++gx)
                        raw[c][rr * (KS / 4) + gx] = __ldg(reinterpret_cast<const uint32_t *>(ptr + 4 * gx));
                    ptr += (r % KS == KS - 1) ? step_chan : step_row;
                }
            }
        };
        auto store = [&](int i, uint32_t (&raw)[CPT][4]) {
            const int s = i & 1;
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            tc::mbar_wait(&empty_a[s], ph ^ 1u);                     // MMAs that read stage s two tiles ago have retired
            const uint32_t dst = a_s + (uint32_t)s * a_bytes + row_off;
#pragma unroll
            for (int c = 0; c < CPT; ++c) tc::sts128u(dst + (uint32_t)c * lbo_a, raw[c][0], raw[c][1], raw[c][2], raw[c][3]);
            tc::fence_async_smem();              // generic-proxy smem writes -> visible to the async (tensor) proxy
            __syncwarp();
            if (lane == 0) {
                tc::mbar_arrive(&full_a[s]);
                if (p.slab) tc::mbar_arrive(&slab_empty[i % kI8Slabs]);      // this tile's slab has been read into registers
            }
        };
        uint32_t ra[CPT][4], rb[CPT][4];
        if (my_tiles > 0) issue(0, ra);
        for (int i = 0; i < my_tiles; i += 2) {
            if (i + 1 < my_tiles) issue(i + 1, rb);
            store(i, ra);
            if (i + 2 < my_tiles) issue(i + 2, ra);
            if (i + 1 < my_tiles) store(i + 1, rb);
        }
    } else if (warp == kI8GatherWarps + 1 + kI8EpiWarps) {
        // ================================ slab copy warp ================================
        // lane = (image of the tile, channel): one bulk copy of the channel rows the tile touches
        if (p.slab) {
            for (int i = 0; i < my_tiles; ++i) {
                const int t = (int)blockIdx.x + i * (int)gridDim.x;
                const ConvI8Seg &sg = t < tiles0 ? p.seg[0] : p.seg[1];
                const int m0 = (t < tiles0 ? t : t - tiles0) * kTcBM;
                const int m1 = m0 + kTcBM < sg.M ? m0 + kTcBM : sg.M;
                const int b_first = m0 / p.P, nseg = (m1 - 1) / p.P - b_first + 1;
                const int buf = i % kI8Slabs;
                if (i >= kI8Slabs) tc::mbar_wait(&slab_empty[buf], (uint32_t)((i / kI8Slabs - 1) & 1));
                const int sgi = lane / p.Cin, c = lane - sgi * p.Cin;
                uint32_t bytes = 0, dst_off = 0, a0 = 0;
                if (sgi < nseg) {
                    for (int bb = b_first; bb <= b_first + sgi; ++bb) {
                        int oy_lo;
                        uint32_t st, en;
                        i8_segment(bb, m0, m1, p.P, p.OW, p.S, p.KS, p.W, oy_lo, st, en);
                        a0 = st & ~15u;
                        bytes = ((en + 15u) & ~15u) - a0;
                        if (bb < b_first + sgi) dst_off += bytes;
                    }
                }
                uint32_t total = bytes;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
                if (lane == 0) tc::mbar_expect_tx(&slab_full[buf], total);
                __syncwarp();
                if (sgi < nseg) {
                    const int64_t bb = sg.gather ? __ldg(sg.gather + b_first + sgi) : (int64_t)(b_first + sgi);
                    tc::bulk_g2s(slab_s + (uint32_t)buf * slab_buf + (uint32_t)c * p.chan_bytes + dst_off,
                                 sg.x + bb * p.in_bstride + (int64_t)c * p.HW + a0, bytes, &slab_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================ epilogue warps ================================
        // warp owns TMEM lanes 32*(warp%4)..+31 (its 32 pixel rows) and every second group of 8 channels
        const int q = warp & 3, sub = (warp - (kI8GatherWarps + 1)) >> 2;
        const bool relu = p.act == B2RL_ACT_RELU, ident = p.act == B2RL_ACT_NONE;
        const bool int_combine = p.k_pad <= 256;                      // (D2 << 8) + D3 stays inside int32 up to 256 taps
        const int64_t oP = p.P;
        const int blocks = p.n_pad / 16;
        const bool full_n = (p.N % 16) == 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int t = (int)blockIdx.x + i * (int)gridDim.x;
            const ConvI8Seg &sg = t < tiles0 ? p.seg[0] : p.seg[1];
            const int em = (t < tiles0 ? t : t - tiles0) * kTcBM + q * 32 + lane;
            const bool e_ok = em < sg.M;
            int b_img = 0, pix = 0;
            if (e_ok) { b_img = em / p.P; pix = em - b_img * p.P; }
            const int64_t o_row = ((int64_t)b_img * p.N + sub * 8) * oP + pix;       // channel sub*8 of this pixel
            const int s = i & 1;
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            tc::mbar_wait(&acc_full[s], ph);
            tc::tc_fence_after();
            const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * acc_cols + sub * 8);
            const int nets = sg.nets;
            for (int net = 0; net < nets; ++net) {
                float *o_net = sg.out[net] + o_row;
                uint32_t col = lane_addr + (uint32_t)(net * 4 * p.n_pad);
                uint32_t so = scale_a + 4u * (uint32_t)(net * p.n_pad + sub * 8), bo = bias_a + 4u * (uint32_t)(net * p.n_pad + sub * 8);
                for (int blk = 0; blk < blocks; ++blk, col += 16u, so += 64u, bo += 64u, o_net += 16 * oP) {
                    uint32_t d0[8], d1[8], d2[8], d3[8];
                    tc::tmem_ld8(col, d0);
                    tc::tmem_ld8(col + (uint32_t)p.n_pad, d1);
                    tc::tmem_ld8(col + 2u * (uint32_t)p.n_pad, d2);
                    tc::tmem_ld8(col + 3u * (uint32_t)p.n_pad, d3);
                    tc::tmem_ld_wait();
                    if (net == nets - 1 && blk == blocks - 1) {       // every TMEM read of this tile has landed in registers
                        tc::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) tc::mbar_arrive(&acc_empty[s]);
                    }
                    float sc[8], bi[8];
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sc[0]), "=f"(sc[1]), "=f"(sc[2]), "=f"(sc[3]) : "r"(so));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(sc[4]), "=f"(sc[5]), "=f"(sc[6]), "=f"(sc[7]) : "r"(so + 16u));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bi[0]), "=f"(bi[1]), "=f"(bi[2]), "=f"(bi[3]) : "r"(bo));
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(bi[4]), "=f"(bi[5]), "=f"(bi[6]), "=f"(bi[7]) : "r"(bo + 16u));
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float tsum;
                        if (int_combine) {        // exact integer pairs, two conversions instead of four
                            const int hi = ((int)d0[j] << 8) + (int)d1[j], lo = ((int)d2[j] << 8) + (int)d3[j];
                            tsum = fmaf((float)hi, 65536.f, (float)lo);
                        } else {                  // smallest digit plane first
                            tsum = fmaf((float)(int)d2[j], 256.f, (float)(int)d3[j]);
                            tsum = fmaf((float)(int)d1[j], 65536.f, tsum);
                            tsum = fmaf((float)(int)d0[j], 16777216.f, tsum);
                        }
                        v[j] = fmaf(tsum, sc[j], bi[j]);
                    }
                    if (relu) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (!ident) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = act_fwd_slow(p.act, v[j]);
                    }
                    if (e_ok) {
                        float *o = o_net;
                        if (full_n) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) { *o = v[j]; o += oP; }
                        } else {
                            const int c0 = blk * 16 + sub * 8;
#pragma unroll
                            for (int j = 0; j < 8; ++j) { if (c0 + j < p.N) *o = v[j]; o += oP; }
                        }
                    }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == kI8GatherWarps) tc::tmem_dealloc(tmem_d, tmem_cols);
}

static bool i8_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("B2RL_DISABLE_I8");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

// Can this layer run on the integer path?  uint8 observations, integer bounds (so x - low is an integer and the
// normalisation is one multiply of the accumulator), 4-tap groups contiguous and 4-byte aligned, whole-K tiles in smem.
static bool conv_i8_ok(const b2rl_layer &l, bool obs_u8, bool normalize, float low, float high, const void *x,
                       int nets_total) {
    if (!i8_enabled() || !obs_u8 || l.kind != B2RL_LAYER_CONV || l.ln != B2RL_LN_NONE) return false;
    if (l.act == B2RL_ACT_GELU) return false;                         // no pre-activation copy on this path
    if (normalize && !(low == floorf(low) && high == floorf(high) && high > low && fabsf(low) <= 1024.f)) return false;
    if (!(l.ksize == 4 || l.ksize == 8) || l.in_c % 2 != 0) return false;   // kernel rows = aligned 4-byte groups; two tap halves
    const int K = l.in_c * l.ksize * l.ksize, n_pad = (l.out_c + 15) / 16 * 16;
    if (K % 32 != 0 || nets_total * 4 * n_pad > 256) return false;   // two accumulators of nets*4*n_pad columns in TMEM
    const int cpt = K / 32;
    if (!(cpt == 1 || cpt == 2 || cpt == 4 || cpt == 8 || cpt == 16)) return false;
    if (conv_i8_smem_bytes(n_pad, K, nets_total) > 200 * 1024) return false;
    // frames: 4 consecutive taps of a kernel row are 4 contiguous, 4-byte aligned bytes
    return l.stride % 4 == 0 && l.in_w % 4 == 0 && (l.in_h * l.in_w) % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 4 == 0;
}

struct ConvI8Job {                 // one segment as the host describes it
    const void *x;
    const int64_t *gather;
    int64_t rows;
    int nets;
    float *out[2];
};

// the digit planes + scales of `nets_total` weight sets into `scratch` (layout of launch_conv_fwd_i8, which may then be
// called with reuse_digits = true on the same scratch)
static int launch_weight_digits(const b2rl_layer &l, bool normalize, float low, float high, const float *const W[2], int nets_total,
                                void *scratch, size_t scratch_bytes, cudaStream_t s) {
    const int K = l.in_c * l.ksize * l.ksize;
    const int n_pad = (l.out_c + 15) / 16 * 16, k_pad = (K + 31) / 32 * 32;
    if (scratch == nullptr || conv_i8_scratch_bytes(n_pad, k_pad, nets_total) > scratch_bytes ||
        reinterpret_cast<uintptr_t>(scratch) % 16 != 0)
        return B2RL_EINVAL;
    int8_t *wd = static_cast<int8_t *>(scratch);
    float *scale = reinterpret_cast<float *>(wd + (size_t)nets_total * 4 * n_pad * k_pad);
    const double inv_range = normalize ? 1.0 / ((double)high - (double)low) : 1.0;
    weight_digits_kernel<<<dim3(n_pad, nets_total), 128, 0, s>>>(W[0], nets_total > 1 ? W[1] : W[0], l.out_c, K, n_pad, k_pad, nets_total,
                                                                inv_range, normalize ? low : 0.f, wd, scale,
                                                                scale + (size_t)nets_total * n_pad);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

// scratch: conv_i8_scratch_bytes(...) bytes, 16-byte aligned.  W[net], bias[net]: fp32 parameters of each weight set.
static int launch_conv_fwd_i8(const b2rl_layer &l, bool normalize, float low, float high, const float *const W[2],
                              const float *const bias[2], int nets_total, const ConvI8Job *jobs, int n_jobs, void *scratch,
                              size_t scratch_bytes, cudaStream_t s, bool reuse_digits = false) {
    const int KK = l.ksize * l.ksize, K = l.in_c * KK, P = l.out_h * l.out_w;
    const int n_pad = (l.out_c + 15) / 16 * 16, k_pad = (K + 31) / 32 * 32;
    if (scratch == nullptr || conv_i8_scratch_bytes(n_pad, k_pad, nets_total) > scratch_bytes) return 1;
    if (reinterpret_cast<uintptr_t>(scratch) % 16 != 0 || n_jobs < 1 || n_jobs > kI8MaxSegs) return 1;
    int8_t *wd = static_cast<int8_t *>(scratch);
    float *scale = reinterpret_cast<float *>(wd + (size_t)nets_total * 4 * n_pad * k_pad);   // [scale | lowcorr]
    const double inv_range = normalize ? 1.0 / ((double)high - (double)low) : 1.0;
    const float lo = normalize ? low : 0.f;
    if (!reuse_digits) {
        weight_digits_kernel<<<dim3(n_pad, nets_total), 128, 0, s>>>(W[0], nets_total > 1 ? W[1] : W[0], l.out_c, K, n_pad, k_pad,
                                                                    nets_total, inv_range, lo, wd, scale,
                                                                    scale + (size_t)nets_total * n_pad);
        B2RL_LAUNCH_CHECK();
    }
    ConvI8Params p;
    memset(&p, 0, sizeof(p));
    int tiles = 0;
    for (int i = 0; i < n_jobs; ++i) {
        if (jobs[i].rows * (int64_t)P > INT32_MAX) return 1;
        ConvI8Seg &sg = p.seg[i];
        sg.x = static_cast<const uint8_t *>(jobs[i].x); sg.gather = jobs[i].gather;
        sg.out[0] = jobs[i].out[0]; sg.out[1] = jobs[i].out[1];
        sg.M = (int)(jobs[i].rows * P); sg.nets = jobs[i].nets; sg.cta0 = tiles;
        tiles += (sg.M + kTcBM - 1) / kTcBM;
    }
    p.n_seg = n_jobs;
    p.wd = wd; p.scale = scale; p.bias[0] = bias[0]; p.bias[1] = nets_total > 1 ? bias[1] : bias[0];
    p.in_bstride = (int64_t)l.in_c * l.in_h * l.in_w;
    p.N = l.out_c; p.n_pad = n_pad; p.k_pad = k_pad; p.nets_total = nets_total;
    p.P = P; p.OW = l.out_w; p.sy = l.stride * l.in_w; p.sx = l.stride; p.act = l.act;
    p.Cin = l.in_c; p.HW = l.in_h * l.in_w; p.W = l.in_w;
    // staged input, OFF by default (B2RL_I8_SLAB=1 turns it on): measured on B200 at B = 256 the bulk-copied slabs are
    // slower than the register gather (16.3 us vs 14.6 us per launch) — the kernel is bound by its per-tile hand-offs,
    // not by the gather's DRAM latency.  Needs planes and images 16-byte aligned, one copy lane per (image, channel).
    p.slab = 0; p.KS = l.ksize; p.S = l.stride; p.nseg_max = 0; p.chan_bytes = 0;
    {
        static int want = -1;
        if (want < 0) { const char *e = getenv("B2RL_I8_SLAB"); want = (e && e[0] == '1') ? 1 : 0; }
        bool ok = want == 1 && (l.in_h * l.in_w) % 16 == 0;
        int nseg_max = 0;
        uint32_t chan_bytes = 0;
        for (int i = 0; i < n_jobs && ok; ++i) {
            if (reinterpret_cast<uintptr_t>(jobs[i].x) % 16 != 0) ok = false;
            const int M = p.seg[i].M;
            for (int m0 = 0; m0 < M; m0 += kTcBM) {
                const int m1 = m0 + kTcBM < M ? m0 + kTcBM : M;
                const int b_first = m0 / P, b_last = (m1 - 1) / P;
                if (b_last - b_first + 1 > nseg_max) nseg_max = b_last - b_first + 1;
                uint32_t tot = 0;
                for (int b = b_first; b <= b_last; ++b) {
                    int oy_lo;
                    uint32_t st, en;
                    i8_segment(b, m0, m1, P, l.out_w, l.stride, l.ksize, l.in_w, oy_lo, st, en);
                    tot += ((en + 15u) & ~15u) - (st & ~15u);
                }
                if (tot > chan_bytes) chan_bytes = tot;
            }
        }
        if (ok && nseg_max * l.in_c <= 32 && conv_i8_smem_bytes(n_pad, k_pad, nets_total, l.in_c, chan_bytes) <= 200 * 1024) {
            p.slab = 1; p.nseg_max = nseg_max; p.chan_bytes = chan_bytes;
        }
    }
    const size_t smem = conv_i8_smem_bytes(n_pad, k_pad, nets_total, p.slab ? l.in_c : 0, p.chan_bytes);
    const int grid = tiles < sm_count() ? tiles : sm_count();          // persistent: one CTA per SM walks its tiles
    if (grid <= 0) return B2RL_OK;
    auto launch = [&](auto kern) -> int {
        { const int rca = ensure_big_smem(kern); if (rca != B2RL_OK) return rca; }
        kern<<<grid, kI8ThreadsP, smem, s>>>(p);
        B2RL_LAUNCH_CHECK();
        ++g_conv_path[1];
        return B2RL_OK;
    };
    const int cpt = k_pad / 32;
#define B2RL_I8_CASE(KS_, CPT_) if (l.ksize == KS_ && cpt == CPT_) return launch(conv_fwd_i8_kernel<KS_, CPT_>)
    B2RL_I8_CASE(8, 2); B2RL_I8_CASE(8, 4); B2RL_I8_CASE(8, 8); B2RL_I8_CASE(8, 16);
    B2RL_I8_CASE(4, 1); B2RL_I8_CASE(4, 2); B2RL_I8_CASE(4, 4); B2RL_I8_CASE(4, 8); B2RL_I8_CASE(4, 16);
#undef B2RL_I8_CASE
    return 1;
}

}  // namespace b2rl
