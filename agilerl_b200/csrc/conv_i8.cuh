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

constexpr int kI8Threads = 256 + 32;   // 8 gather/epilogue warps + the MMA warp
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
    const uint32_t *koff4;     // [k_pad / 4] input offset of every group of 4 consecutive taps
    int64_t in_bstride;        // Cin*H*W
    int N, n_pad, k_pad;       // Cout, Cout rounded up to 16, Cin*k*k rounded up to 32
    int nets_total;            // weight sets resident in wd (1 or 2)
    int P, OW, sy, sx;         // output pixels per image, output width, in-row stride (s*W), in-col stride (s)
    int act;
    float low;                 // integer lower bound: D is computed on raw bytes, low * sum_k w is folded into the bias term
};

static inline size_t conv_i8_smem_bytes(int n_pad, int k_pad, int nets_total) {
    return (size_t)kTcBM * k_pad + (size_t)nets_total * 4 * n_pad * k_pad + (size_t)k_pad + 2 * (size_t)nets_total * n_pad * 4 +
           256 + 1024;
}
// scratch the launcher needs (bytes): digit planes + scales + low-correction + tap-group offsets
static inline size_t conv_i8_scratch_bytes(int n_pad, int k_pad, int nets_total) {
    return (size_t)nets_total * 4 * n_pad * k_pad + (size_t)nets_total * n_pad * 8 + (size_t)k_pad + 256;
}

// One CTA per (output channel, weight set): channel maximum -> power-of-two scale -> four balanced int8 digits per
// weight, stored where the conv kernel's bulk copy expects them.  Channel n of set `net` is row net*4*n_pad + d*n_pad + n.
__global__ void weight_digits_kernel(const float *__restrict__ w0, const float *__restrict__ w1, int N, int K, int n_pad,
                                     int k_pad, int nets_total, double inv_range, float low, int8_t *__restrict__ wd,
                                     float *__restrict__ scale, float *__restrict__ lowcorr, int KK, int KS, int HW, int W,
                                     uint32_t *__restrict__ koff4) {
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
    if (n == 0 && net == 0 && koff4)
        for (int g = threadIdx.x; g < k_pad / 4; g += blockDim.x) {
            const int k = g * 4;
            uint32_t off = 0;                                         // padded taps read offset 0 against zero digits
            if (k < K) {
                const int ci = k / KK, rem = k - ci * KK;
                const int ky = rem / KS, kx = rem - ky * KS;
                off = (uint32_t)(ci * HW + ky * W + kx);
            }
            koff4[g] = off;
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

// CPT = 16-tap chunks per gather thread (k_pad / 16 / 2): compile-time so the raw words stay in registers.
template <int CPT>
__global__ void __launch_bounds__(kI8Threads, (CPT <= 8 ? 3 : 2)) conv_fwd_i8_kernel(const ConvI8Params p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5;
    // which segment (input rows x weight sets) this CTA belongs to
    int si = 0;
    if (p.n_seg > 1 && (int)blockIdx.x >= p.seg[1].cta0) si = 1;
    const ConvI8Seg &sg = p.seg[si];
    const int nets = sg.nets;
    const int cta = (int)blockIdx.x - sg.cta0;
    const int rows_b = p.nets_total * 4 * p.n_pad;                       // rows of the resident B tile
    const uint32_t a_bytes = (uint32_t)kTcBM * p.k_pad, b_bytes = (uint32_t)rows_b * p.k_pad;
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t a_s = sbase, b_s = sbase + a_bytes;
    const uint32_t koff_a = b_s + b_bytes;                              // k_pad/4 words
    const uint32_t scale_a = koff_a + (uint32_t)p.k_pad;                // nets_total*n_pad floats
    const uint32_t bias_a = scale_a + (uint32_t)p.nets_total * p.n_pad * 4;
    const uint32_t bars_a = (bias_a + (uint32_t)p.nets_total * p.n_pad * 4 + 15u) & ~15u;
    uint8_t *gen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    uint64_t *full_a = reinterpret_cast<uint64_t *>(gen + (bars_a - sbase));   // im2col tile written (8 warp arrivals)
    uint64_t *full_b = full_a + 1;                                              // digit planes landed (tx count)
    uint64_t *mma_done = full_a + 2;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(full_a + 3);
    const uint32_t lbo_a = kTcBM * 16, lbo_b = (uint32_t)rows_b * 16;

    if (tid == 0) {
        tc::mbar_init(full_a, 8);
        tc::mbar_init(full_b, 1);
        tc::mbar_init(mma_done, 1);
        tc::fence_barrier_init();
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < nets * 4 * p.n_pad) tmem_cols <<= 1;
    if (warp == 8) tc::tmem_alloc(tmem_ptr, tmem_cols);
    for (int g = tid; g < p.k_pad / 4; g += kI8Threads)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(koff_a + 4u * g), "r"(__ldg(p.koff4 + g)) : "memory");
    for (int n = tid; n < p.nets_total * p.n_pad; n += kI8Threads) {
        const int net = n / p.n_pad, c = n - net * p.n_pad;
        const float *bp = p.bias[net];
        // bias + the low-bound correction of this channel (scale table holds [scale | lowcorr])
        const float bv = ((c < p.N && bp) ? bp[c] : 0.f) + __ldg(p.scale + p.nets_total * p.n_pad + n);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(scale_a + 4u * n), "f"(__ldg(p.scale + n)) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_a + 4u * n), "f"(bv) : "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    if (warp == 8) {
        // ---- MMA warp: one bulk copy brings every digit plane, then k_pad/32 MMAs over the resident tiles
        if (tc::elect_one()) {
            if (nets == p.nets_total) {
                tc::mbar_expect_tx(full_b, b_bytes);
                tc::bulk_g2s(b_s, p.wd, b_bytes, full_b);
            } else {        // first weight set only: its rows are the leading part of every 16-tap column of the tile
                const uint32_t part = (uint32_t)nets * 4 * p.n_pad * 16;
                const int cols = p.k_pad / 16;
                tc::mbar_expect_tx(full_b, part * cols);
                for (int c = 0; c < cols; ++c) tc::bulk_g2s(b_s + c * lbo_b, p.wd + (size_t)c * lbo_b, part, full_b);
            }
        }
        __syncwarp();
        tc::mbar_wait(full_b, 0);
        tc::mbar_wait(full_a, 0);
        tc::tc_fence_after();
        const uint32_t idesc = tc::make_idesc_i8(kTcBM, nets * 4 * p.n_pad);
        const uint64_t da0 = tc::make_desc(a_s, lbo_a, 128), db0 = tc::make_desc(b_s, lbo_b, 128);
        const uint64_t da_step = (uint64_t)((2 * lbo_a) >> 4), db_step = (uint64_t)((2 * lbo_b) >> 4);
        if (tc::elect_one()) {
            const int steps = p.k_pad / 32;                          // one MMA k-step = 32 taps = 2 core-matrix columns
            for (int j = 0; j < steps; ++j) tc::mma_i8(tmem_d, da0 + j * da_step, db0 + j * db_step, idesc, j ? 1u : 0u);
            tc::mma_commit(mma_done);
        }
        __syncwarp();
    } else {
        // ---- gather warps: thread = (im2col row, half of the taps); the raw frame bytes ARE the operand
        const int row = tid & (kTcBM - 1), half = tid >> 7;
        const int m = cta * kTcBM + row;
        const bool row_ok = m < sg.M;
        int64_t rowbase = sg.gather ? sg.gather[0] * p.in_bstride : 0;     // rows beyond M read a valid address
        if (row_ok) {
            const int b = m / p.P, pix = m - b * p.P;
            const int oy = pix / p.OW, ox = pix - oy * p.OW;
            const int64_t bb = sg.gather ? sg.gather[b] : (int64_t)b;
            rowbase = bb * p.in_bstride + (int64_t)(oy * p.sy + ox * p.sx);
        }
        const uint8_t *rp = sg.x + rowbase;
        uint32_t raw[CPT][4];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const uint32_t g0 = (uint32_t)(half * CPT + c) * 4;
            uint32_t off[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) off[j] = tc::lds32(koff_a + 4u * (g0 + j));
#pragma unroll
            for (int j = 0; j < 4; ++j) raw[c][j] = __ldg(reinterpret_cast<const uint32_t *>(rp + off[j]));
        }
        const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
#pragma unroll
        for (int c = 0; c < CPT; ++c)
            tc::sts128u(a_s + (uint32_t)(half * CPT + c) * lbo_a + row_off, raw[c][0], raw[c][1], raw[c][2], raw[c][3]);
        tc::fence_async_smem();                  // generic-proxy smem writes -> visible to the async (tensor) proxy
        __syncwarp();
        if ((tid & 31) == 0) tc::mbar_arrive(full_a);

        // ---- epilogue: warp w owns TMEM lanes 32*(w%4)..+31 (its rows) and channel groups of parity w/4
        tc::mbar_wait(mma_done, 0);
        tc::tc_fence_after();
        const int q = warp & 3;
        const int er = q * 32 + (tid & 31);
        const int em = cta * kTcBM + er;
        const bool e_ok = em < sg.M;
        int b_img = 0, pix = 0;
        if (e_ok) { b_img = em / p.P; pix = em - b_img * p.P; }
        const bool relu = p.act == B2RL_ACT_RELU, ident = p.act == B2RL_ACT_NONE;
        const int oP = p.P;
        for (int net = 0; net < nets; ++net) {
            float *outp = sg.out[net];
            for (int c0 = (warp >> 2) * 8; c0 < p.n_pad; c0 += 16) {      // 8 channels x 4 digit planes per pass
                uint32_t d0[8], d1[8], d2[8], d3[8];
                const uint32_t col = (uint32_t)(net * 4 * p.n_pad + c0);
                const uint32_t lane_addr = tmem_d + ((uint32_t)(q * 32) << 16);
                tc::tmem_ld8(lane_addr + col, d0);
                tc::tmem_ld8(lane_addr + col + (uint32_t)p.n_pad, d1);
                tc::tmem_ld8(lane_addr + col + 2u * (uint32_t)p.n_pad, d2);
                tc::tmem_ld8(lane_addr + col + 3u * (uint32_t)p.n_pad, d3);
                tc::tmem_ld_wait();
                if (e_ok) {
                    const int nv = min(8, p.N - c0);
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // smallest digit plane first: every partial sum is exact or rounded once at its own magnitude
                        float t = fmaf((float)(int)d2[j], 256.f, (float)(int)d3[j]);
                        t = fmaf((float)(int)d1[j], 65536.f, t);
                        t = fmaf((float)(int)d0[j], 16777216.f, t);
                        const float sc = __uint_as_float(tc::lds32(scale_a + 4u * (net * p.n_pad + c0 + j)));
                        v[j] = fmaf(t, sc, __uint_as_float(tc::lds32(bias_a + 4u * (net * p.n_pad + c0 + j))));
                    }
                    if (relu) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                    } else if (!ident) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = act_fwd_slow(p.act, v[j]);
                    }
                    float *o = outp + ((int64_t)b_img * p.N + c0) * oP + pix;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j < nv) o[j * oP] = v[j];
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 8) tc::tmem_dealloc(tmem_d, tmem_cols);
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
    const int K = l.in_c * l.ksize * l.ksize, k_pad = (K + 31) / 32 * 32, n_pad = (l.out_c + 15) / 16 * 16;
    if (nets_total * 4 * n_pad > 256 || (k_pad / 16) % 2 != 0) return false;
    const int cpt = k_pad / 32;
    if (!(cpt == 1 || cpt == 2 || cpt == 4 || cpt == 8 || cpt == 16)) return false;
    if (conv_i8_smem_bytes(n_pad, k_pad, nets_total) > 100 * 1024) return false;
    const bool vec = l.ksize % 4 == 0 && l.stride % 4 == 0 && l.in_w % 4 == 0 && (l.in_h * l.in_w) % 4 == 0 &&
                     reinterpret_cast<uintptr_t>(x) % 4 == 0 && K % 4 == 0;
    return vec;
}

struct ConvI8Job {                 // one segment as the host describes it
    const void *x;
    const int64_t *gather;
    int64_t rows;
    int nets;
    float *out[2];
};

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
    uint32_t *koff4 = reinterpret_cast<uint32_t *>(scale + 2 * (size_t)nets_total * n_pad);
    const double inv_range = normalize ? 1.0 / ((double)high - (double)low) : 1.0;
    const float lo = normalize ? low : 0.f;
    if (!reuse_digits) {
        weight_digits_kernel<<<dim3(n_pad, nets_total), 128, 0, s>>>(W[0], nets_total > 1 ? W[1] : W[0], l.out_c, K, n_pad, k_pad,
                                                                    nets_total, inv_range, lo, wd, scale,
                                                                    scale + (size_t)nets_total * n_pad, KK, l.ksize,
                                                                    l.in_h * l.in_w, l.in_w, koff4);
        B2RL_LAUNCH_CHECK();
    }
    ConvI8Params p;
    memset(&p, 0, sizeof(p));
    int cta = 0;
    for (int i = 0; i < n_jobs; ++i) {
        if (jobs[i].rows * (int64_t)P > INT32_MAX) return 1;
        ConvI8Seg &sg = p.seg[i];
        sg.x = static_cast<const uint8_t *>(jobs[i].x); sg.gather = jobs[i].gather;
        sg.out[0] = jobs[i].out[0]; sg.out[1] = jobs[i].out[1];
        sg.M = (int)(jobs[i].rows * P); sg.nets = jobs[i].nets; sg.cta0 = cta;
        cta += (sg.M + kTcBM - 1) / kTcBM;
    }
    p.n_seg = n_jobs;
    p.wd = wd; p.scale = scale; p.bias[0] = bias[0]; p.bias[1] = nets_total > 1 ? bias[1] : bias[0]; p.koff4 = koff4;
    p.in_bstride = (int64_t)l.in_c * l.in_h * l.in_w;
    p.N = l.out_c; p.n_pad = n_pad; p.k_pad = k_pad; p.nets_total = nets_total;
    p.P = P; p.OW = l.out_w; p.sy = l.stride * l.in_w; p.sx = l.stride; p.act = l.act; p.low = lo;
    const size_t smem = conv_i8_smem_bytes(n_pad, k_pad, nets_total);
    auto launch = [&](auto kern) -> int {
        { const int rca = ensure_big_smem(kern); if (rca != B2RL_OK) return rca; }
        kern<<<cta, kI8Threads, smem, s>>>(p);
        B2RL_LAUNCH_CHECK();
        return B2RL_OK;
    };
    switch (k_pad / 32) {
        case 1: return launch(conv_fwd_i8_kernel<1>);
        case 2: return launch(conv_fwd_i8_kernel<2>);
        case 4: return launch(conv_fwd_i8_kernel<4>);
        case 8: return launch(conv_fwd_i8_kernel<8>);
        case 16: return launch(conv_fwd_i8_kernel<16>);
        default: return 1;
    }
}

}  // namespace b2rl
