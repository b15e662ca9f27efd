// conv_tc.cuh — convolution forward as an implicit GEMM on the 5th-gen tensor cores (tcgen05).
//
//   out[b, co, oy, ox] = act( bias[co] + sum_k  im2col(x)[pixel, k] * W[co, k] ),   k = (ci, ky, kx)
//
// D (128 pixels x Cout, fp32) lives in TMEM; A (im2col tile) and B (weight tile) are built in
// shared memory by the CTA's own threads — the A operand is a gather (uint8 frames straight from
// the replay ring through the sampled row indices, exact (x-low)/(high-low) through a LUT), so
// there is no TMA descriptor that could stage it — in the canonical K-major no-swizzle UMMA
// layout, and one elected thread issues tcgen05.mma.kind::tf32.
//
// fp32-equivalent accuracy (the 1e-5 parity gate) comes from the 3xTF32 split:
//   x = hi + lo, hi = tf32_rn(x), lo = x - hi (exact);   x*w ~= hi_x*hi_w + lo_x*hi_w + hi_x*lo_w
// three MMAs per k-step accumulating into the same fp32 TMEM tile (dropped term ~2^-22 relative).
//
// smem tile layout (per operand, per stage; T = 4 tf32 per 16 B):
//   element (row r, k) at  (k/4)*LBO + (r/8)*128 + (r%8)*16 + (k%4)*4      LBO = rows*16, SBO = 128
// i.e. core matrices (8 rows x 16 B) are contiguous 128 B, stacked along M/N first, then along K —
// exactly "((8,n),2):((1,SBO),LBO)" of the UMMA K-major INTERLEAVE descriptor.
#pragma once
#include "gemm.cuh"

namespace b2rl {

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// K-major, no swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate, K-major A/B:
//   c_format[4,6)=1 (F32) | a_format[7,10)=2 (TF32) | b_format[10,13)=2 | n_dim[17,23)=N>>3 | m_dim[24,29)=M>>4
__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

}  // namespace tc

constexpr int kTcBM = 128;      // pixels per CTA (UMMA M)
constexpr int kTcBK = 32;       // k per stage (4 MMA k-steps of 8)
constexpr int kTcStages = 2;

struct ConvTcParams {
    const void *x;              // input: uint8 frames / fp32 activations (NCHW)
    const float *w;             // [Cout, K] row-major (K = Cin*k*k)
    const float *bias;          // [Cout]
    float *out;                 // NCHW
    float *pre_out;             // optional pre-activation copy
    const int64_t *gather;      // optional ring rows per batch row
    int64_t in_bstride;         // Cin*H*W
    int M, N, K;                // pixels (rows*P), Cout, Cin*k*k
    int n_pad, k_pad;           // N rounded up to 16, K rounded up to 32
    int P, OW, sy, sx;          // output pixels per image, output width, in-row stride (s*W), in-col stride (s)
    int KK, KS, HW, W;          // kernel taps per channel, kernel size, input H*W, input W
    int act;
    int normalize;
    float low, high;
};

static inline size_t conv_tc_smem_bytes(int n_pad, int k_pad) {
    const size_t a = (size_t)kTcBM * kTcBK * 4, b = (size_t)n_pad * kTcBK * 4;
    return kTcStages * 2 * (a + b) + (size_t)k_pad * 4 + 256 * 4 + (size_t)n_pad * 4 + 64 + 1024;
}

template <int ELEM>
__global__ void __launch_bounds__(128) conv_fwd_tc_kernel(const ConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t a_bytes = kTcBM * kTcBK * 4, b_bytes = (uint32_t)p.n_pad * kTcBK * 4;
    // carve: [stage][A_hi, A_lo, B_hi, B_lo] | koff | lut | bias | barriers | tmem ptr
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t stage_bytes = 2 * (a_bytes + b_bytes);
    auto a_hi = [&](int s) { return base + (size_t)s * stage_bytes; };
    auto a_lo = [&](int s) { return base + (size_t)s * stage_bytes + a_bytes; };
    auto b_hi = [&](int s) { return base + (size_t)s * stage_bytes + 2 * a_bytes; };
    auto b_lo = [&](int s) { return base + (size_t)s * stage_bytes + 2 * a_bytes + b_bytes; };
    int *koff = reinterpret_cast<int *>(base + (size_t)kTcStages * stage_bytes);
    float *lut = reinterpret_cast<float *>(koff + p.k_pad);
    float *sbias = lut + 256;
    uint64_t *bars = reinterpret_cast<uint64_t *>((reinterpret_cast<uintptr_t>(sbias + p.n_pad) + 15) & ~uintptr_t(15));
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + kTcStages);

    // ---- one-time setup ----------------------------------------------------------------------
    for (int k = tid; k < p.k_pad; k += 128) {
        int off = -1;
        if (k < p.K) {
            const int ci = k / p.KK, rem = k - ci * p.KK;
            const int ky = rem / p.KS, kx = rem - ky * p.KS;
            off = ci * p.HW + ky * p.W + kx;
        }
        koff[k] = off;
    }
    if (ELEM == EL_U8)
        for (int i = tid; i < 256; i += 128) lut[i] = p.normalize ? __fdiv_rn((float)i - p.low, p.high - p.low) : (float)i;
    for (int n = tid; n < p.n_pad; n += 128) sbias[n] = (n < p.N && p.bias) ? p.bias[n] : 0.f;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.n_pad) tmem_cols <<= 1;
    if (warp == 0) tc::tmem_alloc(tmem_ptr, tmem_cols);
    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) tc::mbar_init(&bars[s], 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    // this thread's im2col row (output pixel)
    const int m = blockIdx.x * kTcBM + tid;
    const bool row_ok = m < p.M;
    int64_t rowbase = 0;
    if (row_ok) {
        const int b = m / p.P, pix = m - b * p.P;
        const int oy = pix / p.OW, ox = pix - oy * p.OW;
        const int64_t bb = p.gather ? p.gather[b] : (int64_t)b;
        rowbase = bb * p.in_bstride + (int64_t)(oy * p.sy + ox * p.sx);
    }
    const uint32_t idesc = tc::make_idesc_tf32(kTcBM, p.n_pad);
    const uint32_t lbo_a = kTcBM * 16, lbo_b = (uint32_t)p.n_pad * 16;
    const int KB = p.k_pad / kTcBK;
    const uint32_t row_off = (uint32_t)(tid >> 3) * 128 + (uint32_t)(tid & 7) * 16;   // (r/8)*128 + (r%8)*16

    for (int kb = 0; kb < KB; ++kb) {
        const int s = kb & 1;
        if (kb >= kTcStages) tc::mbar_wait(&bars[s], (uint32_t)((kb / kTcStages - 1) & 1));   // MMAs of this stage's last use done
        // ---- A tile: 32 gathered values of this pixel row -> hi/lo -> 8 x 16 B stores each
        uint32_t raw[kTcBK];
#pragma unroll
        for (int j = 0; j < kTcBK; ++j) {
            const int off = koff[kb * kTcBK + j];
            raw[j] = 0u;
            if (row_ok && off >= 0) {
                if (ELEM == EL_U8) raw[j] = (uint32_t)__ldg(static_cast<const uint8_t *>(p.x) + rowbase + off);
                else raw[j] = __float_as_uint(__ldg(static_cast<const float *>(p.x) + rowbase + off));
            }
        }
        uint8_t *ah = a_hi(s) + row_off, *al = a_lo(s) + row_off;
#pragma unroll
        for (int c = 0; c < kTcBK / 4; ++c) {
            float hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kk = c * 4 + j;
                float v;
                if (ELEM == EL_U8) v = (row_ok && koff[kb * kTcBK + kk] >= 0) ? lut[raw[kk]] : 0.f;
                else if (ELEM == EL_F32_NORM)
                    v = (row_ok && koff[kb * kTcBK + kk] >= 0) ? __fdiv_rn(__uint_as_float(raw[kk]) - p.low, p.high - p.low) : 0.f;
                else v = __uint_as_float(raw[kk]);
                hi[j] = tc::tf32_rn(v);
                lo[j] = v - hi[j];
            }
            *reinterpret_cast<float4 *>(ah + (size_t)c * lbo_a) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4 *>(al + (size_t)c * lbo_a) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        // ---- B tile: n_pad x 32 weights, 16 B (4 k) per item
        for (int it = tid; it < p.n_pad * (kTcBK / 4); it += 128) {
            const int n = it % p.n_pad, c = it / p.n_pad;
            float hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = kb * kTcBK + c * 4 + j;
                const float v = (n < p.N && k < p.K) ? __ldg(p.w + (int64_t)n * p.K + k) : 0.f;
                hi[j] = tc::tf32_rn(v);
                lo[j] = v - hi[j];
            }
            const uint32_t o = (uint32_t)c * lbo_b + (uint32_t)(n >> 3) * 128 + (uint32_t)(n & 7) * 16;
            *reinterpret_cast<float4 *>(b_hi(s) + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4 *>(b_lo(s) + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        tc::fence_async_smem();          // generic-proxy smem writes -> visible to the async (tensor) proxy
        __syncthreads();
        if (tid == 0) {
            tc::tc_fence_after();
            const uint32_t ah_addr = tc::smem_u32(a_hi(s)), al_addr = tc::smem_u32(a_lo(s));
            const uint32_t bh_addr = tc::smem_u32(b_hi(s)), bl_addr = tc::smem_u32(b_lo(s));
#pragma unroll
            for (int j = 0; j < kTcBK / 8; ++j) {      // one MMA k-step = 8 tf32 = 2 core-matrix columns
                const uint64_t dah = tc::make_desc(ah_addr + 2 * j * lbo_a, lbo_a, 128);
                const uint64_t dal = tc::make_desc(al_addr + 2 * j * lbo_a, lbo_a, 128);
                const uint64_t dbh = tc::make_desc(bh_addr + 2 * j * lbo_b, lbo_b, 128);
                const uint64_t dbl = tc::make_desc(bl_addr + 2 * j * lbo_b, lbo_b, 128);
                tc::mma_tf32(tmem_d, dah, dbh, idesc, (kb | j) ? 1u : 0u);
                tc::mma_tf32(tmem_d, dal, dbh, idesc, 1u);
                tc::mma_tf32(tmem_d, dah, dbl, idesc, 1u);
            }
            tc::mma_commit(&bars[s]);
        }
    }
    // ---- wait for every outstanding MMA group ----------------------------------------------------
    for (int s = 0; s < kTcStages; ++s) {
        const int uses = (KB - s + kTcStages - 1) / kTcStages;      // k-blocks that used stage s
        if (uses > 0) tc::mbar_wait(&bars[s], (uint32_t)((uses - 1) & 1));
    }
    tc::tc_fence_after();

    // ---- epilogue: TMEM -> registers -> bias + activation -> NCHW ---------------------------------
    int b_img = 0, pix = 0;
    if (row_ok) { b_img = m / p.P; pix = m - b_img * p.P; }
    for (int c0 = 0; c0 < p.n_pad; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = c0 + j;
                if (n < p.N) {
                    float v = __uint_as_float(r[j]) + sbias[n];
                    const int64_t o = ((int64_t)b_img * p.N + n) * p.P + pix;
                    if (p.pre_out) p.pre_out[o] = v;
                    p.out[o] = act_fwd(p.act, v);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, tmem_cols);
    (void)lane;
}

// returns B2RL_OK, or 1 when the shape is outside what the tensor-core kernel handles (caller falls
// back to the FFMA engine)
static int launch_conv_fwd_tc(const b2rl_layer &l, const Operand &A, const float *W, const float *bias, float *out,
                              float *pre_out, int64_t rows, cudaStream_t s) {
    const int KK = l.ksize * l.ksize, K = l.in_c * KK, P = l.out_h * l.out_w;
    const int n_pad = (l.out_c + 15) / 16 * 16, k_pad = (K + kTcBK - 1) / kTcBK * kTcBK;
    if (n_pad > 256 || k_pad > 8192 || rows * (int64_t)P > INT32_MAX) return 1;
    const size_t smem = conv_tc_smem_bytes(n_pad, k_pad);
    if (smem > 200 * 1024) return 1;
    ConvTcParams p;
    p.x = A.ptr; p.w = W; p.bias = bias; p.out = out; p.pre_out = pre_out; p.gather = A.row.gather;
    p.in_bstride = (int64_t)l.in_c * l.in_h * l.in_w;
    p.M = (int)(rows * P); p.N = l.out_c; p.K = K; p.n_pad = n_pad; p.k_pad = k_pad;
    p.P = P; p.OW = l.out_w; p.sy = l.stride * l.in_w; p.sx = l.stride;
    p.KK = KK; p.KS = l.ksize; p.HW = l.in_h * l.in_w; p.W = l.in_w;
    p.act = l.act; p.normalize = A.normalize; p.low = A.low; p.high = A.high;
    const int grid = (p.M + kTcBM - 1) / kTcBM;
    auto launch = [&](auto kern) -> int {
        B2RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 128, smem, s>>>(p);
        B2RL_LAUNCH_CHECK();
        return B2RL_OK;
    };
    switch (A.elem_kind()) {
        case EL_U8: return launch(conv_fwd_tc_kernel<EL_U8>);
        case EL_F32_NORM: return launch(conv_fwd_tc_kernel<EL_F32_NORM>);
        default: return launch(conv_fwd_tc_kernel<EL_F32>);
    }
}

}  // namespace b2rl
