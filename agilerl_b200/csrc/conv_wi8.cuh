// conv_wi8.cuh — weight gradient of the FIRST convolution (uint8 frames) on the integer tensor path.
//
//   dW[co][tap] = sum_pix ((x[pix][tap] - low) / (high - low)) * G[pix][co]
//               = inv_range * ( sum_pix x * G  -  low * sum_pix G )                                  (+ db[co] = sum_pix G)
//
// The frame bytes are exact integers, so — as in the forward pass (conv_i8.cuh) — the im2col operand goes to the tensor
// core AS BYTES, with no conversion: the K-major [pixels x taps] tile the forward gather builds is, read as an MN-major
// operand, exactly the [taps x pixels] tile this product needs (kind::i8 takes MN-major A; tools/probes/umma_mn_probe2.cu).
// The precision is carried by G: per output channel, G / 2^E (2^E > max|G_c|, from chan_absmax_zero_kernel) is rounded to a
// 38-bit fixed-point integer q, shifted by 2^39 so that it is non-negative, and written as its five BYTES (rows
// d * n_pad + co of the B tile, u8 x u8 MMA): extracting them is byte permutes, no carry arithmetic.  A sixth group of
// rows holds a row of ones, so the same MMA also produces sum_pix x per tap, which removes the shift again:
// sum x q = sum x (q + 2^39) - 2^39 * sum x, exactly, in int64 (pixels beyond the tensor cancel the same way).
// ONE tcgen05.mma kind::i8 with N = 5 * n_pad + 16 multiplies 32 pixels of a 128-tap tile with all planes into exact
// int32 accumulators, which stay in TMEM across all pixel tiles of the persistent CTA (<= 768 pixels: sums < 2^26).
// The epilogue recombines the digits in int64 and writes one [Cout][taps] int64 partial per CTA (integer sums: exact and
// order-independent, so the result is deterministic); wgrad_i8_finish_kernel adds the partials and scales them into dW / db.
// (A first version added into one shared accumulator with 64-bit atomics: 1.2 M atomics on 8 192 addresses were half of
// the kernel's time.)
// Error: the 2^-38 quantisation of G relative to its channel maximum, times sum|x| — below fp32 round-off of the result.
//
// Persistent, warp-specialised like conv_fwd_i8_kernel: 8 gather warps (frame bytes -> A stage, identical code path),
// 8 quantiser warps (G -> digit planes; after the last tile: TMEM -> int64 atomics), one MMA warp; two stages.
#pragma once
#include "conv_i8.cuh"

namespace b2rl {

constexpr int kWi8Digits = 5;
constexpr int kWi8Bits = 37;                 // |q| < 2^37; q + 2^39 has five bytes; int64 sums over 2^25 pixel-bytes stay below 2^63
constexpr int kWi8OnesRows = 16;             // extra B rows: row 0 = ones (sum of the frame bytes per tap), 15 rows of zero padding
constexpr int kWi8Slots = 64;                // per-channel partial maxima (one block each, no atomics, no initialisation)
constexpr int kWi8Threads = (kI8GatherWarps + 1 + kI8EpiWarps) * 32;

// scratch layout (bytes from a 256-byte aligned base): [maxbuf: n_pad*kWi8Slots u32][accb: n_pad int64][part: ctas*N*Kc int64]
static inline size_t conv_wi8_scratch_bytes(int n_pad, int Kc, int N, int ctas) {
    return (size_t)n_pad * kWi8Slots * 4 + (size_t)n_pad * 8 + (size_t)ctas * Kc * N * 8 + 64;
}

// block (co, slot): max |G| over the images b = slot, slot + kWi8Slots, ... of channel co -> maxbuf[co][slot];
// the blocks also zero the int64 accumulators the weight-gradient kernel adds into, and — when `a` is given — apply the
// layer's activation backward to G in place first (the act_bwd_kernel launch this replaces reads the same planes).
__global__ void chan_absmax_zero_kernel(float *__restrict__ g, const float *__restrict__ a, const float *__restrict__ pre, int act,
                                        int64_t rows, int N, int P, uint32_t *__restrict__ maxbuf, long long *__restrict__ acc,
                                        int64_t acc_n) {
    __shared__ float red[8];
    const int co = blockIdx.x, slot = blockIdx.y;
    const int64_t nb = (int64_t)gridDim.x * gridDim.y, bid = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    for (int64_t i = bid * blockDim.x + threadIdx.x; i < acc_n; i += nb * blockDim.x) acc[i] = 0;
    float mx = 0.f;
    const bool vec = (P & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 && (!a || (reinterpret_cast<uintptr_t>(a) & 15) == 0) && !pre;
    for (int64_t b = slot; b < rows; b += kWi8Slots) {
        const int64_t base = (b * N + co) * P;
        if (vec) {                                   // planes are multiples of 16 bytes: four elements per access
            float4 *g4 = reinterpret_cast<float4 *>(g + base);
            const float4 *a4 = a ? reinterpret_cast<const float4 *>(a + base) : nullptr;
            for (int i = threadIdx.x; i < P / 4; i += blockDim.x) {
                float4 v = g4[i];
                if (a4) {
                    const float4 y = a4[i];
                    v.x *= act_bwd(act, 0.f, y.x); v.y *= act_bwd(act, 0.f, y.y);
                    v.z *= act_bwd(act, 0.f, y.z); v.w *= act_bwd(act, 0.f, y.w);
                    g4[i] = v;
                }
                mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
            }
            continue;
        }
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            float v = g[base + i];
            if (a) {
                v *= act_bwd(act, pre ? pre[base + i] : 0.f, a[base + i]);
                g[base + i] = v;
            }
            mx = fmaxf(mx, fabsf(v));
        }
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
        maxbuf[co * kWi8Slots + slot] = __float_as_uint(mx);
    }
}

// exponent E with 2^E > m (m finite, >= 0); E = INT_MIN/2 marks an all-zero (or non-finite) channel
__host__ __device__ __forceinline__ int wi8_exponent(float m) {
    if (!(m > 0.f) || !(m < INFINITY)) return -100000;
    int e;
    frexpf(m, &e);
    return e;
}

struct ConvWi8Params {
    const uint8_t *x;            // frames (ring base)
    const int64_t *gather;       // ring row per batch row, or NULL
    const float *g;              // [rows, N, P]
    const uint32_t *maxbuf;      // [n_pad][kWi8Slots] float bits
    long long *part;             // [ctas][N][Kc]
    long long *accb;             // [n_pad]
    int64_t in_bstride;
    int M, N, n_pad, Kc, k_pad;
    int P, OW, sy, sx;
    int Cin, HW, W;
};

static inline size_t conv_wi8_smem_bytes(int n_pad, int k_pad) {
    return 2 * (size_t)kTcBM * k_pad + 2 * (size_t)(kWi8Digits * n_pad + kWi8OnesRows) * kTcBM + (size_t)n_pad * 8 + 256 + 1024;
}

template <int KS, int CPT>
__global__ void __launch_bounds__(kWi8Threads, 1) conv_wgrad_i8_kernel(const ConvWi8Params p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows_b = kWi8Digits * p.n_pad + kWi8OnesRows;                // rows of the digit tile = columns of an accumulator
    const int MT = (p.Kc + kTcBM - 1) / kTcBM;
    const uint32_t a_bytes = (uint32_t)kTcBM * p.k_pad, b_bytes = (uint32_t)rows_b * kTcBM;
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t a_s = sbase, b_s = sbase + 2 * a_bytes;
    const uint32_t up_a = b_s + 2 * b_bytes;                               // n_pad doubles: 2^(37 - E_co), 0 for an all-zero channel
    const uint32_t bars_a = (up_a + (uint32_t)p.n_pad * 8 + 15u) & ~15u;
    uint8_t *gen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    double *up_tab = reinterpret_cast<double *>(gen + (up_a - sbase));
    uint64_t *full_a = reinterpret_cast<uint64_t *>(gen + (bars_a - sbase));   // [2] frame bytes written (8 gather warps)
    uint64_t *full_g = full_a + 2;                                              // [2] digit planes written (8 quantiser warps)
    uint64_t *empty = full_a + 4;                                               // [2] stage consumed (MMA commit)
    uint64_t *acc_done = full_a + 6;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(full_a + 7);
    const uint32_t lbo_b = (uint32_t)rows_b * 16;
    const int n_tiles = (p.M + kTcBM - 1) / kTcBM;
    const int my_tiles = ((int)blockIdx.x < n_tiles) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&full_a[s], kI8GatherWarps);
            tc::mbar_init(&full_g[s], kI8EpiWarps);
            tc::mbar_init(&empty[s], 1);
        }
        tc::mbar_init(acc_done, 1);
        tc::fence_barrier_init();
    }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < MT * rows_b) tmem_cols <<= 1;
    if (warp == kI8GatherWarps) tc::tmem_alloc(tmem_ptr, tmem_cols);
    for (int c = tid; c < p.n_pad; c += kWi8Threads) {
        float m = 0.f;
        if (c < p.N)
            for (int sl = 0; sl < kWi8Slots; ++sl) m = fmaxf(m, __uint_as_float(__ldg(p.maxbuf + c * kWi8Slots + sl)));
        const int e = wi8_exponent(m);
        up_tab[c] = e > -100000 ? ldexp(1.0, kWi8Bits - e) : 0.0;
    }
    // the ones rows of both B stages (the quantiser warps never touch them): row kWi8Digits * n_pad = 1, the rest 0
    for (int e = tid; e < 2 * 8 * kWi8OnesRows * 4; e += kWi8Threads) {          // [stage][pixel chunk][row][4 words]
        const int w = e & 3, row = (e >> 2) % kWi8OnesRows, chunk = (e >> 2) / kWi8OnesRows % 8, stg = (e >> 2) / (kWi8OnesRows * 8);
        const int n = kWi8Digits * p.n_pad + row;
        const uint32_t a = b_s + (uint32_t)stg * b_bytes + (uint32_t)chunk * lbo_b + (uint32_t)(n >> 3) * 128u + (uint32_t)(n & 7) * 16u + 4u * w;
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(row == 0 ? 0x01010101u : 0u) : "memory");
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    if (warp == kI8GatherWarps) {
        // ================================ MMA warp ================================
        // A: MN-major u8, (tap m, pixel k) at (m/16)*SBO + (k/8)*LBO + (k%8)*16 + m%16 with SBO = 2048, LBO = 128 — the
        // forward kernel's K-major [pixel][tap] image.  B: K-major s8 digit planes.
        // c = S32, a = U8, b = U8 (the shifted digits are unsigned bytes), A MN-major
        const uint32_t idesc = (2u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | ((uint32_t)(rows_b >> 3) << 17) | ((uint32_t)(kTcBM >> 4) << 24);
        for (int i = 0; i < my_tiles; ++i) {
            const int s = i & 1;
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            tc::mbar_wait(&full_a[s], ph);
            tc::mbar_wait(&full_g[s], ph);
            tc::tc_fence_after();
            const uint32_t a_addr = a_s + (uint32_t)s * a_bytes, g_addr = b_s + (uint32_t)s * b_bytes;
            if (tc::elect_one()) {
                for (int tt = 0; tt < MT; ++tt) {
                    const uint32_t d_addr = tmem_d + (uint32_t)(tt * rows_b);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {                       // 32 pixels per MMA
                        const uint64_t da = tc::make_desc(a_addr + (uint32_t)tt * 8u * (kTcBM * 16) + (uint32_t)ks * 512u, 128, kTcBM * 16);
                        const uint64_t db = tc::make_desc(g_addr + (uint32_t)ks * 2u * lbo_b, lbo_b, 128);
                        tc::mma_i8(d_addr, da, db, idesc, (i | ks) ? 1u : 0u);
                    }
                }
                tc::mma_commit(&empty[s]);
                if (i == my_tiles - 1) tc::mma_commit(acc_done);
            }
            __syncwarp();
        }
    } else if (warp < kI8GatherWarps) {
        // ================================ gather warps (conv_fwd_i8_kernel's register path) ================================
        const int row = tid & (kTcBM - 1), half = tid >> 7;
        const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16 + (uint32_t)(half * CPT) * (kTcBM * 16);
        const int64_t half_off = (int64_t)half * (p.Cin / 2) * p.HW;
        const int64_t step_row = p.W, step_chan = (int64_t)p.HW - (int64_t)(KS - 1) * p.W;
        auto issue = [&](int i, uint32_t (&raw)[CPT][4]) {
            const int t = (int)blockIdx.x + i * (int)gridDim.x;
            const int m = t * kTcBM + row;
            int64_t rowbase = p.gather ? __ldg(p.gather) * p.in_bstride : 0;      // pixels beyond M read a valid address (their G is 0)
            if (m < p.M) {
                const int b = m / p.P, pix = m - b * p.P;
                const int oy = pix / p.OW, ox = pix - oy * p.OW;
                const int64_t bb = p.gather ? __ldg(p.gather + b) : (int64_t)b;
                rowbase = bb * p.in_bstride + (int64_t)(oy * p.sy + ox * p.sx);
            }
            const uint8_t *ptr = p.x + rowbase + half_off;
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
#pragma unroll
                for (int rr = 0; rr < 16 / KS; ++rr) {
                    const int r = c * (16 / KS) + rr;
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
            tc::mbar_wait(&empty[s], ph ^ 1u);
            const uint32_t dst = a_s + (uint32_t)s * a_bytes + row_off;
#pragma unroll
            for (int c = 0; c < CPT; ++c) tc::sts128u(dst + (uint32_t)c * (kTcBM * 16), raw[c][0], raw[c][1], raw[c][2], raw[c][3]);
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full_a[s]);
        };
        uint32_t ra[CPT][4], rb[CPT][4];
        if (my_tiles > 0) issue(0, ra);
        for (int i = 0; i < my_tiles; i += 2) {
            if (i + 1 < my_tiles) issue(i + 1, rb);
            store(i, ra);
            if (i + 2 < my_tiles) issue(i + 2, ra);
            if (i + 1 < my_tiles) store(i + 1, rb);
        }
    } else {
        // ================================ quantiser warps ================================
        // warp = 16-pixel chunk of the tile, lane = output channel (+32 per pass): 16 values of G -> five int8 digits each
        const int gw = warp - (kI8GatherWarps + 1);                         // 0..7
        const bool vec_ok = (p.P % 4) == 0 && (reinterpret_cast<uintptr_t>(p.g) % 16) == 0;
        double bsum[2] = {0.0, 0.0};                                        // sum of q of this thread's channels: integers below 2^53
        for (int i = 0; i < my_tiles; ++i) {
            const int t = (int)blockIdx.x + i * (int)gridDim.x;
            const int m0 = t * kTcBM + gw * 16;
            const int s = i & 1;
            const uint32_t ph = (uint32_t)((i >> 1) & 1);
            int pass = 0;
            for (int co = lane; co < p.n_pad; co += 32, ++pass) {
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = 0.f;
                if (co < p.N && m0 < p.M) {
                    const int b = m0 / p.P, pix = m0 - b * p.P;
                    if (vec_ok && (pix & 3) == 0 && pix + 16 <= p.P && m0 + 16 <= p.M) {
                        const float4 *q4 = reinterpret_cast<const float4 *>(p.g + ((int64_t)b * p.N + co) * p.P + pix);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 f = __ldg(q4 + j);
                            v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
                        }
                    } else {
                        int bb = b, pp = pix;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            if (m0 + j < p.M) v[j] = __ldg(p.g + ((int64_t)bb * p.N + co) * p.P + pp);
                            if (++pp == p.P) { pp = 0; ++bb; }
                        }
                    }
                }
                const double up = up_tab[co];
                // q' = rint(v * up) + 2^39 in [2^39 - 2^37, 2^39 + 2^37]: A = q' >> 16 (24 bits), t = q' & 65535; the five digit
                // bytes of a pixel are bytes 2, 1, 0 of A and bytes 1, 0 of t — gathered four pixels at a time with byte permutes
                uint32_t Aw[16], tw[16];
                double qs = 0.0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const double qr = rint((double)v[j] * up);
                    qs += qr;
                    const double qd = qr + 549755813888.0;
                    const int A = __double2int_rd(qd * 1.52587890625e-05);
                    Aw[j] = (uint32_t)A;
                    tw[j] = (uint32_t)__double2int_rn(qd - (double)A * 65536.0);
                }
                uint32_t dw[kWi8Digits][4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const uint32_t a0 = Aw[4 * w], a1 = Aw[4 * w + 1], a2 = Aw[4 * w + 2], a3 = Aw[4 * w + 3];
                    const uint32_t t0 = tw[4 * w], t1 = tw[4 * w + 1], t2 = tw[4 * w + 2], t3 = tw[4 * w + 3];
                    // byte k of four registers -> one word: two permutes pick the byte of each pair, a third joins the pairs
                    dw[0][w] = __byte_perm(__byte_perm(a0, a1, 0x0062), __byte_perm(a2, a3, 0x0062), 0x5410);
                    dw[1][w] = __byte_perm(__byte_perm(a0, a1, 0x0051), __byte_perm(a2, a3, 0x0051), 0x5410);
                    dw[2][w] = __byte_perm(__byte_perm(a0, a1, 0x0040), __byte_perm(a2, a3, 0x0040), 0x5410);
                    dw[3][w] = __byte_perm(__byte_perm(t0, t1, 0x0051), __byte_perm(t2, t3, 0x0051), 0x5410);
                    dw[4][w] = __byte_perm(__byte_perm(t0, t1, 0x0040), __byte_perm(t2, t3, 0x0040), 0x5410);
                }
                bsum[pass & 1] += qs;
                if (pass == 0) tc::mbar_wait(&empty[s], ph ^ 1u);               // the MMAs that read stage s two tiles ago have retired
                const uint32_t dst = b_s + (uint32_t)s * b_bytes + (uint32_t)gw * lbo_b + (uint32_t)(co >> 3) * 128 + (uint32_t)(co & 7) * 16;
#pragma unroll
                for (int d = 0; d < kWi8Digits; ++d)
                    tc::sts128u(dst + (uint32_t)(d * p.n_pad / 8) * 128u, dw[d][0], dw[d][1], dw[d][2], dw[d][3]);
            }
            tc::fence_async_smem();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full_g[s]);
        }
        {
            int pass = 0;
            for (int co = lane; co < p.N; co += 32, ++pass)
                if (bsum[pass & 1] != 0.0)
                    atomicAdd(reinterpret_cast<unsigned long long *>(p.accb + co), (unsigned long long)__double2ll_rn(bsum[pass & 1]));
        }
        // ---- epilogue: TMEM lanes 32*(warp%4).. = taps; digit columns -> int64 partial of this CTA
        if (my_tiles > 0) {
            tc::mbar_wait(acc_done, 0);
            tc::tc_fence_after();
            const int q = warp & 3, half = gw >> 2;
            for (int tt = half; tt < MT; tt += 2) {
                const int tap = tt * kTcBM + q * 32 + lane;
                uint32_t ones[8];
                tc::tmem_ld8(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(tt * rows_b + kWi8Digits * p.n_pad), ones);
                tc::tmem_ld_wait();
                const long long shift = (long long)(int)ones[0] << 39;          // 2^39 * sum_pix x of this tap
                for (int c0 = 0; c0 < p.n_pad; c0 += 8) {
                    uint32_t r[kWi8Digits][8];
#pragma unroll
                    for (int d = 0; d < kWi8Digits; ++d)
                        tc::tmem_ld8(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(tt * rows_b + d * p.n_pad + c0), r[d]);
                    tc::tmem_ld_wait();
                    if (tap < p.Kc) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (c0 + j >= p.N) break;
                            long long tsum = 0;
#pragma unroll
                            for (int d = 0; d < kWi8Digits; ++d) tsum = tsum * 256 + (long long)(int)r[d][j];
                            p.part[((int64_t)blockIdx.x * p.N + c0 + j) * p.Kc + tap] = tsum - shift;   // lanes = consecutive taps: coalesced
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

// dW[co][tap] (+)= inv_range * 2^(E - 37) * (sum_cta part[cta][co][tap] - low * accb[co]),   db[co] (+)= 2^(E - 37) * accb[co]
__global__ void wgrad_i8_finish_kernel(const long long *__restrict__ part, int ctas, const long long *__restrict__ accb,
                                       const uint32_t *__restrict__ maxbuf, int N, int Kc, double inv_range, double low,
                                       float *__restrict__ dw, float *__restrict__ db, int accumulate) {
    const int co = blockIdx.x;
    __shared__ double sc_s;
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int sl = 0; sl < kWi8Slots; ++sl) m = fmaxf(m, __uint_as_float(maxbuf[co * kWi8Slots + sl]));
        const int e = wi8_exponent(m);
        sc_s = e > -100000 ? ldexp(1.0, e - kWi8Bits) : 0.0;
    }
    __syncthreads();
    const double sc = sc_s;
    const double sg = (double)accb[co];
    // block = 128 taps x 8 groups of CTAs: 8 x fewer dependent loads per thread, then a fixed-order sum of the groups
    __shared__ long long red[8][128];
    const int tx = threadIdx.x & 127, grp = threadIdx.x >> 7;
    const int tap = blockIdx.y * 128 + tx;
    long long t = 0;
    if (tap < Kc) {
        const long long *q = part + (int64_t)co * Kc + tap;
#pragma unroll 4
        for (int c = grp; c < ctas; c += 8) t += q[(int64_t)c * N * Kc];
    }
    red[grp][tx] = t;
    __syncthreads();
    if (grp == 0 && tap < Kc) {
#pragma unroll
        for (int k = 1; k < 8; ++k) t += red[k][tx];
        const double v = ((double)t - low * sg) * sc * inv_range;
        float *o = dw + (int64_t)co * Kc + tap;
        *o = accumulate ? *o + (float)v : (float)v;
    }
    if (blockIdx.y == 0 && threadIdx.x == 0 && db) {
        const float v = (float)(sg * sc);
        db[co] = accumulate ? db[co] + v : v;
    }
}

// can launch_conv_wgrad_i8 take this layer?  (pure host check: the caller may then leave the activation backward to it)
static bool conv_wgrad_i8_ok(const b2rl_layer &l, const Operand &X, int64_t rows, const void *scratch, size_t scratch_bytes) {
    static int want = -1;
    if (want < 0) { const char *e = getenv("B2RL_WGRAD_I8"); want = (e && e[0] == '0') ? 0 : 1; }
    if (!want || !X.u8) return false;
    if (!conv_i8_ok(l, true, X.normalize != 0, X.low, X.high, X.ptr, 1)) return false;
    const int KK = l.ksize * l.ksize, Kc = l.in_c * KK, P = l.out_h * l.out_w;
    const int n_pad = (l.out_c + 15) / 16 * 16, k_pad = (Kc + 31) / 32 * 32;
    const int MT = (Kc + kTcBM - 1) / kTcBM;
    const int rows_b = kWi8Digits * n_pad + kWi8OnesRows;
    if (rows_b > 256 || MT * rows_b > 512 || n_pad > 64 || k_pad != Kc) return false;
    const int cpt = k_pad / 32;
    if (!(l.ksize == 8 ? (cpt == 2 || cpt == 4 || cpt == 8) : (cpt == 1 || cpt == 2 || cpt == 4 || cpt == 8))) return false;
    // int64 totals: pixels * 255 * 2^37 must stay below 2^63
    if (rows < 1 || rows * (int64_t)P > INT32_MAX || rows * (int64_t)P * 255 >= ((int64_t)1 << 26)) return false;
    if (conv_wi8_smem_bytes(n_pad, k_pad) > 200 * 1024) return false;
    const int n_tiles = (int)((rows * P + kTcBM - 1) / kTcBM);
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    if (scratch == nullptr || conv_wi8_scratch_bytes(n_pad, Kc, l.out_c, grid) > scratch_bytes || reinterpret_cast<uintptr_t>(scratch) % 16 != 0)
        return false;
    // a CTA's int32 accumulators see tiles_per_cta * 128 pixels of x * digit <= 255 * 255; its int64 partial
    // tiles_per_cta * 128 * 255 * 2^40
    const int64_t tiles_per_cta = (n_tiles + grid - 1) / grid;
    return tiles_per_cta * kTcBM * 255 * 255 < ((int64_t)1 << 31) && tiles_per_cta * kTcBM * 255 < ((int64_t)1 << 22);
}

// g is dL/d(layer output); with act_a != NULL it is the gradient BEFORE the layer's activation backward, which the first
// kernel then applies in place (g becomes what act_bwd_kernel would have left).  Call only when conv_wgrad_i8_ok.
static int launch_conv_wgrad_i8(const b2rl_layer &l, const Operand &X, float *g, float *dw, float *db, int accumulate,
                                int64_t rows, void *scratch, size_t scratch_bytes, cudaStream_t s, const float *act_a = nullptr,
                                const float *act_pre = nullptr, int act = 0) {
    if (!conv_wgrad_i8_ok(l, X, rows, scratch, scratch_bytes)) return 1;
    const int KK = l.ksize * l.ksize, Kc = l.in_c * KK, P = l.out_h * l.out_w;
    const int n_pad = (l.out_c + 15) / 16 * 16, k_pad = (Kc + 31) / 32 * 32;
    const size_t smem = conv_wi8_smem_bytes(n_pad, k_pad);
    const int M = (int)(rows * P);
    const int n_tiles = (M + kTcBM - 1) / kTcBM;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    uint32_t *maxbuf = static_cast<uint32_t *>(scratch);
    long long *accb = reinterpret_cast<long long *>(maxbuf + (size_t)n_pad * kWi8Slots);
    long long *part = accb + n_pad;
    chan_absmax_zero_kernel<<<dim3(l.out_c, kWi8Slots), 128, 0, s>>>(g, act_a, act_pre, act, rows, l.out_c, P, maxbuf, accb,
                                                                     (int64_t)n_pad);
    B2RL_LAUNCH_CHECK();
    ConvWi8Params p;
    p.x = static_cast<const uint8_t *>(X.ptr); p.gather = X.red.gather; p.g = g; p.maxbuf = maxbuf; p.part = part; p.accb = accb;
    p.in_bstride = (int64_t)l.in_c * l.in_h * l.in_w;
    p.M = M; p.N = l.out_c; p.n_pad = n_pad; p.Kc = Kc; p.k_pad = k_pad;
    p.P = P; p.OW = l.out_w; p.sy = l.stride * l.in_w; p.sx = l.stride;
    p.Cin = l.in_c; p.HW = l.in_h * l.in_w; p.W = l.in_w;
    auto launch = [&](auto kern) -> int {
        { const int rca = ensure_big_smem(kern); if (rca != B2RL_OK) return rca; }
        kern<<<grid, kWi8Threads, smem, s>>>(p);
        B2RL_LAUNCH_CHECK();
        return B2RL_OK;
    };
    const int cpt = k_pad / 32;
    int rc = 1;
#define B2RL_WI8_CASE(KS_, CPT_) if (l.ksize == KS_ && cpt == CPT_) rc = launch(conv_wgrad_i8_kernel<KS_, CPT_>)
    B2RL_WI8_CASE(8, 2); B2RL_WI8_CASE(8, 4); B2RL_WI8_CASE(8, 8);
    B2RL_WI8_CASE(4, 1); B2RL_WI8_CASE(4, 2); B2RL_WI8_CASE(4, 4); B2RL_WI8_CASE(4, 8);
#undef B2RL_WI8_CASE
    if (rc != B2RL_OK) return rc;
    const double inv_range = X.normalize ? 1.0 / ((double)X.high - (double)X.low) : 1.0;
    wgrad_i8_finish_kernel<<<dim3(l.out_c, (Kc + 127) / 128), 1024, 0, s>>>(part, grid, accb, maxbuf, l.out_c, Kc, inv_range,
                                                                           X.normalize ? (double)X.low : 0.0, dw, db, accumulate);
    B2RL_LAUNCH_CHECK();
    ++g_conv_path[1];
    return B2RL_OK;
}

}  // namespace b2rl
