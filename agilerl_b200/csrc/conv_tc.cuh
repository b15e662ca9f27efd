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
    // try_wait suspends the thread in hardware up to the hint before it returns false: few polls,
    // no issue slots burnt while the tensor pipe or the producers are busy
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680u)
        : "memory");
}
// one lane of a converged warp (the CUTLASS elect_one_sync idiom): code under `if (elect_one())` keeps
// its operands in uniform registers, which is what the tcgen05 / bulk-copy instructions take
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared by the TMA unit; completion is signalled on the mbarrier's tx count
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// non-blocking probe of an mbarrier phase (true: the phase with this parity has completed)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// round-to-nearest (ties away) to tf32 on the bit pattern: two integer instructions instead of the ten the compiler emits for
// cvt.rna.tf32.f32; identical for every finite value whose rounding does not overflow the exponent
__device__ __forceinline__ float tf32_rn_fast(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

}  // namespace tc

constexpr int kTcBM = 128;      // pixels per CTA (UMMA M)
constexpr int kTcBK = 32;       // k per stage (4 MMA k-steps of 8)
constexpr int kTcStages = 2;    // operand stages of the weight-gradient kernel (both tiles thread-built)
constexpr int kTcBStages = 4;   // forward: B-operand ring (bulk-copied by the TMA unit, two k-blocks ahead)
constexpr int kTcThreads = 256; // producers: two threads per im2col row (16 k each per stage)
constexpr int kTcFwdThreads = kTcThreads + 64;   // forward kernel: + MMA warp + weight-copy warp

struct ConvTcParams {
    const void *x;              // input: uint8 frames / fp32 activations (NCHW)
    const float *w_hi, *w_lo;   // weights pre-split into tf32 hi/lo, already in the smem tile layout
    const uint32_t *koff;       // [k_pad] input offset of every tap
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
    // PAD variant (input gradient as s*s stride-1 convolutions over dL/dout, one per parity class = blockIdx.y)
    int pad, IH;                // implicit zero padding, input plane height (its width is W)
    int cls_s;                  // classes per dimension (the forward stride); output pixel = (yy*s + py, xx*s + px)
    int out_H, out_W;           // plane of the scattered output
    int64_t w_class_stride;     // floats between the pre-split weights of consecutive classes
    long long *dbg;             // diagnostics (B2RL_TC_DBG=<cta>): clock64 stamps of one CTA's roles, else NULL
    int dbg_cta;
};

static inline size_t conv_tc_smem_bytes(int n_pad, int k_pad, int a_parts = 2, int a_stages = 2, int b_stages = kTcBStages) {
    const size_t a = (size_t)kTcBM * kTcBK * 4, b = (size_t)n_pad * kTcBK * 4;
    return a_stages * a_parts * a + b_stages * 2 * b + (size_t)k_pad * 4 + 256 * 4 + (size_t)n_pad * 4 + 256 + 1024;
}

// Split W [N, K] into tf32 hi / lo and store it in the order the conv kernel's B tiles use: one
// [2*n_pad rows x 32 k] tile per k-block whose rows 0..n_pad-1 are the hi parts and n_pad..2*n_pad-1 the
// lo parts, so that ONE tcgen05.mma with N = 2*n_pad multiplies an im2col k-step with both:
//   [k-block][chunk c = (k%32)/4][part (hi, lo)][n/8][n%8][k%4]   (2 * n_pad * 32 floats per k-block, zero padded)
// (w_hl is the whole array; the unused w_lo argument is kept for the call sites)
__global__ void weight_split_kernel(const float *__restrict__ w, int N, int K, int n_pad, int k_pad,
                                    float *__restrict__ w_hi, float *__restrict__ w_lo, int KK, int KS, int HW, int W,
                                    uint32_t *__restrict__ koff) {
    const int total = n_pad * k_pad;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = e / n_pad, n = e - k * n_pad;
        const float v = (n < N && k < K) ? w[(int64_t)n * K + k] : 0.f;
        const float hi = tc::tf32_rn(v);
        const int kb = k / kTcBK, kin = k - kb * kTcBK;
        const int64_t o = (int64_t)kb * 2 * n_pad * kTcBK + (int64_t)(kin >> 2) * 2 * n_pad * 4 + (n >> 3) * 32 + (n & 7) * 4 + (kin & 3);
        w_hi[o] = hi;
        w_hi[o + (int64_t)n_pad * 4] = v - hi;
        if (n == 0 && koff) {      // input offset of tap k = (ci, ky, kx); padded taps read offset 0 against zero weights
            uint32_t off = 0;
            if (k < K) {
                const int ci = k / KK, rem = k - ci * KK;
                const int ky = rem / KS, kx = rem - ky * KS;
                off = (uint32_t)(ci * HW + ky * W + kx);
            }
            koff[k] = off;
        }
    }
}

__device__ __noinline__ float act_fwd_slow(int act, float v) { return act_fwd(act, v); }

// opt a kernel into large dynamic shared memory once (the launchers cap their request at kTcSmemCap):
// cudaFuncSetAttribute costs microseconds of host time per call, and there are ~10 conv launches per step
constexpr int kTcSmemCap = 200 * 1024;
template <typename Kern>
static int ensure_big_smem(Kern kern) {
    // keyed by the function's address: every instantiation shares the pointer TYPE, not the pointer
    static const void *seen[64];
    static int n_seen = 0;
    const void *key = reinterpret_cast<const void *>(kern);
    for (int i = 0; i < n_seen; ++i)
        if (seen[i] == key) return B2RL_OK;
    B2RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemCap));
    if (n_seen < 64) seen[n_seen++] = key;
    return B2RL_OK;
}

// Forward convolution.  Latency structure (the kernel is bound by the im2col gather, not by the
// tensor pipe): a thread keeps the raw taps of the next D k-blocks in registers (for packed uint8
// frames D = 8 covers a whole 8x8x4 receptive field: every DRAM round trip of a row is in flight at
// once), the weight tiles arrive by cp.async.bulk two k-blocks ahead into a 4-deep ring signalled
// by mbarrier transaction counts, and tcgen05.commit releases A stage / B stage pairs.
//
// EXACT_A: uint8 observations with integer bounds.  (x - low) is a small integer, exactly
// representable in tf32, so the im2col operand needs no lo part: 2 MMAs per k-step instead of 3,
// no LUT, half the A-tile bytes; the 1/(high - low) of the normalisation is applied to the fp32
// accumulator in the epilogue (one true division per output).  vs. the reference's
// fl((x-low)/(high-low)) * w summed in fp32 the difference is <= 2^-23 relative per term.
// VEC: 4 consecutive taps are contiguous and aligned (4 packed bytes / two 8-byte fp32 loads).
// SB: slots of the weight-tile ring; tiles are requested SB-2 k-blocks ahead (a bulk copy takes on the
// order of a microsecond to land — more than one k-block of work)
template <int ELEM, bool EXACT_A, bool VEC, int D, bool PAD = false, int SA = 2, int SB = kTcBStages>
__global__ void __launch_bounds__(kTcFwdThreads) conv_fwd_tc_kernel(const ConvTcParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int RAWN = (ELEM == EL_U8 && VEC) ? 4 : 16;      // raw words per thread per k-block
    constexpr int CH = kTcBK / 4 / 2;                          // 4-tap chunks per thread per k-block (4)
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row = tid & (kTcBM - 1), half = tid >> 7;       // im2col row, which 16 of the 32 k
    const uint32_t a_bytes = kTcBM * kTcBK * 4, b_bytes = (uint32_t)p.n_pad * kTcBK * 4;
    // carve (32-bit shared-space addresses so that the compiler emits LDS/STS):
    //   A stages [hi (, lo)] | B stages [hi, lo] | koff | lut | bias | barriers | tmem ptr
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t a_stage = (EXACT_A ? 1 : 2) * a_bytes;
    auto a_hi = [&](int s) { return sbase + (uint32_t)s * a_stage; };
    auto a_lo = [&](int s) { return sbase + (uint32_t)s * a_stage + a_bytes; };          // unused when EXACT_A
    const uint32_t bbase = sbase + SA * a_stage;
    auto b_hi = [&](int s) { return bbase + (uint32_t)s * 2 * b_bytes; };
    auto b_lo = [&](int s) { return bbase + (uint32_t)s * 2 * b_bytes + b_bytes; };
    const uint32_t koff_a = bbase + SB * 2 * b_bytes;
    const uint32_t lut_a = koff_a + (uint32_t)p.k_pad * 4;
    const uint32_t bias_a = lut_a + 256 * 4;
    const uint32_t bars_a = (bias_a + (uint32_t)p.n_pad * 4 + 15u) & ~15u;
    const uint32_t tptr_a = bars_a + 8 * (2 * SA + 2 * SB);
    uint8_t *gen = smem_raw + (sbase - tc::smem_u32(smem_raw));        // generic alias of sbase (barriers only)
    uint64_t *mma_bar = reinterpret_cast<uint64_t *>(gen + (bars_a - sbase));   // [SA] MMA group done
    uint64_t *full_a = mma_bar + SA;                                     // [SA] im2col tile written
    uint64_t *full_b = full_a + SA;                                      // [SB] weight tile landed
    uint64_t *empty_b = full_b + SB;                                     // [SB] weight tile consumed (MMAs retired)
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(gen + (tptr_a - sbase));
    const int KB = p.k_pad / kTcBK;
    // diagnostics: time stamps of producer warp 0 (slots 0..) and of the MMA lane (slots 64..) of one CTA
    const bool dbg_on = p.dbg != nullptr && (int)blockIdx.x == p.dbg_cta && blockIdx.y == 0;
    auto stamp = [&](int slot) { if (dbg_on) p.dbg[slot] = clock64(); };
    if (tid == 0) stamp(0);

    // ---- one-time setup ----------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < SA; ++s) {
            tc::mbar_init(&mma_bar[s], 1);
            tc::mbar_init(&full_a[s], kTcThreads / 32);      // one elected arrive per producer warp
        }
        for (int s = 0; s < SB; ++s) {
            tc::mbar_init(&full_b[s], 1);
            tc::mbar_init(&empty_b[s], 1);
        }
        tc::fence_barrier_init();
    }
    // MMA group j (k-block j) has retired: its A stage and its B ring slot may be overwritten
    auto wait_mma = [&](int j) { tc::mbar_wait(&mma_bar[j % SA], (uint32_t)((j / SA) & 1)); };
    auto issue_b = [&](int kb) {       // elected thread: weight tile of k-block kb -> its ring slot
        const int sb = kb & (SB - 1);
        tc::mbar_expect_tx(&full_b[sb], 2 * b_bytes);
        const int64_t wo = (PAD ? (int64_t)blockIdx.y * p.w_class_stride : 0) + (int64_t)kb * 2 * p.n_pad * kTcBK;
        tc::bulk_g2s(b_hi(sb), p.w_hi + wo, 2 * b_bytes, &full_b[sb]);          // hi|lo tile, one copy
    };
    for (int k = tid; k < p.k_pad; k += kTcFwdThreads)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(koff_a + 4u * k), "r"(__ldg(p.koff + k)) : "memory");
    if (ELEM == EL_U8 && !EXACT_A)
        for (int i = tid; i < 256; i += kTcFwdThreads) {
            const float v = p.normalize ? __fdiv_rn((float)i - p.low, p.high - p.low) : (float)i;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(lut_a + 4u * i), "f"(v) : "memory");
        }
    for (int n = tid; n < p.n_pad; n += kTcFwdThreads) {
        const float v = (n < p.N && p.bias) ? p.bias[n] : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_a + 4u * n), "f"(v) : "memory");
    }
    uint32_t tmem_cols = 32;                                   // D = [A.W_hi | A.W_lo]: 2*n_pad columns
    while ((int)tmem_cols < 2 * p.n_pad) tmem_cols <<= 1;
    if (warp == 0) tc::tmem_alloc(tmem_ptr, tmem_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (tid == 0) stamp(1);
    const uint32_t tmem_d = *tmem_ptr;
    const uint32_t idesc2 = tc::make_idesc_tf32(kTcBM, 2 * p.n_pad);      // A_hi x [W_hi; W_lo]
    const uint32_t idesc1 = tc::make_idesc_tf32(kTcBM, p.n_pad);          // A_lo x W_hi
    const uint32_t lbo_a = kTcBM * 16, lbo_b = (uint32_t)p.n_pad * 2 * 16;   // a 4-tap chunk holds hi rows, then lo rows

    if (warp == kTcThreads / 32 + 1) {
        // ---- weight-copy warp: keeps the TMA ring SB-2 k-blocks ahead, so that requesting a tile (an
        // expect_tx and two bulk copies, a few hundred cycles of issue latency) is off the MMA lane's path
        if (tc::elect_one()) {
            for (int j = 0; j < SB - 2 && j < KB; ++j) issue_b(j);
        }
        __syncwarp();
        for (int kb = 0; kb + SB - 2 < KB; ++kb) {
            // ring slot (kb+SB-2)%SB was last read by MMA group kb-2.  Its own "consumed" barrier, not the
            // A-stage barrier: the next completion on this slot needs the copy issued right here, so this
            // warp can never fall a whole phase behind the barrier it polls.
            if (kb >= 2) tc::mbar_wait(&empty_b[(kb - 2) % SB], (uint32_t)(((kb - 2) / SB) & 1));
            if (tc::elect_one()) issue_b(kb + SB - 2);
            __syncwarp();
        }
    }
    if (warp == kTcThreads / 32) {
        // ---- MMA warp.  The whole warp runs the loop converged (waits, descriptor arithmetic stay in the
        // uniform datapath); one elected lane issues the copies, the MMAs and the commit.
        const uint64_t da_step = (uint64_t)((2 * lbo_a) >> 4), db_step = (uint64_t)((2 * lbo_b) >> 4);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % SA, sb = kb & (SB - 1);
            tc::mbar_wait(&full_b[sb], (uint32_t)((kb / SB) & 1));
            if ((tid & 31) == 0 && kb < 16) stamp(64 + 3 * kb);
            tc::mbar_wait(&full_a[s], (uint32_t)((kb / SA) & 1));
            if ((tid & 31) == 0 && kb < 16) stamp(65 + 3 * kb);
            tc::tc_fence_after();
            // descriptors of consecutive k-steps differ only in the 14-bit start-address field
            const uint64_t dah0 = tc::make_desc(a_hi(s), lbo_a, 128), dal0 = tc::make_desc(a_lo(s), lbo_a, 128);
            const uint64_t db0 = tc::make_desc(b_hi(sb), lbo_b, 128);
            if (tc::elect_one()) {
#pragma unroll
                for (int j = 0; j < kTcBK / 8; ++j) {      // one MMA k-step = 8 tf32 = 2 core-matrix columns
                    // columns [0, n_pad) += A_hi.W_hi (+ A_lo.W_hi), columns [n_pad, 2 n_pad) += A_hi.W_lo
                    tc::mma_tf32(tmem_d, dah0 + j * da_step, db0 + j * db_step, idesc2, (kb | j) ? 1u : 0u);
                    if (!EXACT_A) tc::mma_tf32(tmem_d, dal0 + j * da_step, db0 + j * db_step, idesc1, 1u);
                }
                tc::mma_commit(&mma_bar[s]);
                tc::mma_commit(&empty_b[sb]);
            }
            if ((tid & 31) == 0 && kb < 16) stamp(66 + 3 * kb);
            __syncwarp();
        }
    }

    if (warp < kTcThreads / 32) {
    // ---- producer warps: this thread's im2col row (output pixel); rows beyond M read a safe address
    const int m = blockIdx.x * kTcBM + row;
    const bool row_ok = m < p.M;
    int64_t rowbase = p.gather ? p.gather[0] * p.in_bstride : 0;
    uint32_t mask_y = 0, mask_x = 0;                  // PAD: which kernel rows / columns fall inside the plane
    if (row_ok) {
        const int b = m / p.P, pix = m - b * p.P;
        const int oy = pix / p.OW, ox = pix - oy * p.OW;
        const int64_t bb = p.gather ? p.gather[b] : (int64_t)b;
        rowbase = bb * p.in_bstride + (int64_t)((oy * p.sy + ox * p.sx) - (PAD ? p.pad * (p.W + 1) : 0));
        if (PAD) {
            for (int a = 0; a < p.KS; ++a) {
                mask_y |= (uint32_t)((unsigned)(oy - p.pad + a) < (unsigned)p.IH) << a;
                mask_x |= (uint32_t)((unsigned)(ox - p.pad + a) < (unsigned)p.W) << a;
            }
        }
    }
    const uint8_t *row_u8 = static_cast<const uint8_t *>(p.x) + rowbase;
    const float *row_f32 = static_cast<const float *>(p.x) + rowbase;
    const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;   // (r/8)*128 + (r%8)*16
    const float exact_bias = 8388608.f + p.low;       // EXACT_A: (2^23 + byte) - (2^23 + low), both exact

    uint32_t raw[D][RAWN];
    auto gather = [&](int kb, uint32_t (&dst)[RAWN]) {
        const int k0 = kb * kTcBK + half * (CH * 4);
        if (PAD) {                              // fp32, per-tap bounds: entry = offset | ky << 24 | kx << 28
#pragma unroll
            for (int j = 0; j < CH * 4; ++j) {
                const uint32_t e = tc::lds32(koff_a + 4u * (k0 + j));
                const bool ok = ((mask_y >> ((e >> 24) & 15u)) & (mask_x >> (e >> 28)) & 1u) != 0;
                dst[j] = ok ? __float_as_uint(__ldg(row_f32 + (e & 0xFFFFFFu))) : 0u;
            }
        } else if (ELEM != EL_U8 && VEC) {      // fp32 activations: 4 taps = 2 aligned 8-byte loads
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                const uint32_t off = tc::lds32(koff_a + 4u * (k0 + c * 4));
                const float2 *q2 = reinterpret_cast<const float2 *>(row_f32 + off);
                const float2 a = __ldg(q2), b = __ldg(q2 + 1);
                dst[c * 4 + 0] = __float_as_uint(a.x); dst[c * 4 + 1] = __float_as_uint(a.y);
                dst[c * 4 + 2] = __float_as_uint(b.x); dst[c * 4 + 3] = __float_as_uint(b.y);
            }
        } else if (ELEM == EL_U8 && VEC) {
            // offsets first, loads second: the shared-memory reads pipeline instead of each one gating its load
#pragma unroll
            for (int c = 0; c < CH; ++c) dst[c] = tc::lds32(koff_a + 4u * (k0 + c * 4));
#pragma unroll
            for (int c = 0; c < CH; ++c)
                dst[c] = __ldg(reinterpret_cast<const uint32_t *>(row_u8 + dst[c]));   // 4 packed taps
        } else {
#pragma unroll
            for (int j = 0; j < CH * 4; ++j) {
                const uint32_t off = tc::lds32(koff_a + 4u * (k0 + j));
                if (ELEM == EL_U8) dst[j] = (uint32_t)__ldg(row_u8 + off);
                else dst[j] = __float_as_uint(__ldg(row_f32 + off));
            }
        }
    };
    if (KB >= D) {
#pragma unroll
        for (int d = 0; d < D; ++d) gather(d, raw[d]);
    } else {
#pragma unroll
        for (int d = 0; d < D; ++d)
            if (d < KB) gather(d, raw[d]);
    }
    if (tid == 0) stamp(2);

    auto step = [&](int kb, uint32_t (&cur)[RAWN]) {
        const int s = kb % SA;
        if (kb >= SA) wait_mma(kb - SA);     // A stage s is free again
        if (tid == 0 && kb < 16) stamp(3 + 3 * kb);
        // ---- convert + hi/lo split + 16-byte smem stores of the gathered taps
        const uint32_t ah = a_hi(s) + row_off + (uint32_t)(half * CH) * lbo_a, al = a_lo(s) + row_off + (uint32_t)(half * CH) * lbo_a;
        const int k0 = kb * kTcBK + half * (CH * 4);
        // padded taps meet zero weights; only non-finite fp32 garbage could leak through them
        const bool tail = ELEM != EL_U8 && kb == KB - 1 && p.K != p.k_pad;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            float hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kk = c * 4 + j;
                float v;
                if (ELEM == EL_U8) {
                    if (EXACT_A) {
                        const uint32_t w = VEC ? __byte_perm(cur[c], 0x4B000000u, 0x7650u + j) : (cur[kk] | 0x4B000000u);
                        v = __uint_as_float(w) - exact_bias;                       // exact in tf32
                    } else {
                        const uint32_t byte = VEC ? ((cur[c] >> (8 * j)) & 0xFFu) : cur[kk];
                        v = __uint_as_float(tc::lds32(lut_a + 4u * byte));
                    }
                } else if (ELEM == EL_F32_NORM) v = __fdiv_rn(__uint_as_float(cur[kk]) - p.low, p.high - p.low);
                else v = __uint_as_float(cur[kk]);
                if (tail) v = (k0 + kk < p.K) ? v : 0.f;
                if (EXACT_A) { hi[j] = v; lo[j] = 0.f; }
                else { hi[j] = tc::tf32_rn_fast(v); lo[j] = v - hi[j]; }
            }
            tc::sts128(ah + (uint32_t)c * lbo_a, hi[0], hi[1], hi[2], hi[3]);
            if (!EXACT_A) tc::sts128(al + (uint32_t)c * lbo_a, lo[0], lo[1], lo[2], lo[3]);
        }
        if (tid == 0 && kb < 16) stamp(4 + 3 * kb);
        if (kb + D < KB) gather(kb + D, cur);   // refill this register slot D k-blocks ahead
        tc::fence_async_smem();                  // generic-proxy smem writes -> visible to the async (tensor) proxy
        __syncwarp();
        if ((tid & 31) == 0) tc::mbar_arrive(&full_a[s]);
        if (tid == 0 && kb < 16) stamp(5 + 3 * kb);
    };
    for (int kb0 = 0; kb0 < KB; kb0 += D) {
#pragma unroll
        for (int d = 0; d < D; ++d)
            if (kb0 + d < KB) step(kb0 + d, raw[d]);
    }
    // ---- the last commit covers every earlier MMA of the issuing thread
    if (KB > 0) wait_mma(KB - 1);
    if (tid == 0) stamp(60);
    tc::tc_fence_after();

    // ---- epilogue: warp w owns TMEM lanes 32*(w%4)..+31 (its rows) and the 16-column groups of parity w/4
    {
        const int q = warp & 3;
        const int er = q * 32 + (tid & 31);                  // tile row of this lane
        const int em = blockIdx.x * kTcBM + er;
        int b_img = 0, pix = 0;
        bool e_ok = em < p.M;
        if (e_ok) { b_img = em / p.P; pix = em - b_img * p.P; }
        int64_t out_P = p.P;
        if (PAD) {                                           // scatter into the class's pixels of the full plane
            const int yy = pix / p.OW, xx = pix - yy * p.OW;
            const int y = yy * p.cls_s + (int)blockIdx.y / p.cls_s, x = xx * p.cls_s + (int)blockIdx.y % p.cls_s;
            e_ok = e_ok && y < p.out_H && x < p.out_W;
            pix = y * p.out_W + x;
            out_P = (int64_t)p.out_H * p.out_W;
        }
        const bool relu = p.act == B2RL_ACT_RELU, ident = p.act == B2RL_ACT_NONE;
        const float scale = (EXACT_A && p.normalize) ? 1.0f / (p.high - p.low) : 1.0f;
        const int oP = (int)out_P;
        for (int c0 = (warp >> 2) * 16; c0 < p.n_pad; c0 += 32) {
            uint32_t r[16], r2[16];
            tc::tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            tc::tmem_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.n_pad + c0), r2);
#pragma unroll
            for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            if (e_ok) {
                const int nv = min(16, p.N - c0);
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float acc = EXACT_A ? __uint_as_float(r[j]) * scale : __uint_as_float(r[j]);
                    v[j] = acc + __uint_as_float(tc::lds32(bias_a + 4u * (c0 + j)));
                }
                const int64_t o0 = ((int64_t)b_img * p.N + c0) * out_P + pix;
                if (p.pre_out) {
                    float *po = p.pre_out + o0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (j < nv) po[j * oP] = v[j];
                }
                if (relu) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
                } else if (!ident) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = act_fwd_slow(p.act, v[j]);
                }
                float *o = p.out + o0;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < nv) o[j * oP] = v[j];
            }
        }
    }
    if (tid == 0) stamp(61);
    }   // producer warps
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, tmem_cols);
}


// B2RL_TC_DBG=<cta index>: the forward kernel records clock64 stamps of that CTA (b2rl_debug_read)
static int tc_debug_cta() {
    static int v = -2;
    if (v == -2) v = getenv("B2RL_TC_DBG") ? atoi(getenv("B2RL_TC_DBG")) : -1;
    return v;
}
static long long *tc_debug_buffer() {
    static long long *buf = nullptr;
    if (tc_debug_cta() < 0) return nullptr;
    if (!buf) {
        if (cudaMalloc(&buf, 128 * sizeof(long long)) != cudaSuccess) return nullptr;
        cudaMemset(buf, 0, 128 * sizeof(long long));
    }
    return buf;
}

static inline size_t conv_tc_wsplit_floats(const b2rl_layer &l) {
    const int K = l.in_c * l.ksize * l.ksize;
    const int n_pad = (l.out_c + 15) / 16 * 16, k_pad = (K + kTcBK - 1) / kTcBK * kTcBK;
    return (size_t)2 * n_pad * k_pad + k_pad + 4 * n_pad + 64;      // hi, lo, tap offsets (+ room for the int8 digit path's scales)
}

// tf32 hi/lo split (+ tap offsets) of one convolution's weights into `wsplit` — what launch_conv_fwd_tc / _st produce
// themselves unless called with reuse_split
static int launch_weight_split(const b2rl_layer &l, const float *W, float *wsplit, size_t wsplit_cap, cudaStream_t s) {
    const int KK = l.ksize * l.ksize, K = l.in_c * KK;
    const int n_pad = (l.out_c + 15) / 16 * 16, k_pad = (K + kTcBK - 1) / kTcBK * kTcBK;
    if (wsplit == nullptr || conv_tc_wsplit_floats(l) > wsplit_cap || reinterpret_cast<uintptr_t>(wsplit) % 16 != 0) return B2RL_EINVAL;
    uint32_t *koff = reinterpret_cast<uint32_t *>(wsplit + (size_t)2 * n_pad * k_pad);
    const int total = n_pad * k_pad;
    weight_split_kernel<<<(total + 255) / 256, 256, 0, s>>>(W, l.out_c, K, n_pad, k_pad, wsplit, nullptr, KK, l.ksize, l.in_h * l.in_w,
                                                            l.in_w, koff);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

// returns B2RL_OK, or 1 when the shape is outside what the tensor-core kernel handles (caller falls
// back to the FFMA engine).  wsplit: scratch for the pre-split weights (conv_tc_wsplit_floats).
static int launch_conv_fwd_tc(const b2rl_layer &l, const Operand &A, const float *W, const float *bias, float *out,
                              float *pre_out, int64_t rows, float *wsplit, size_t wsplit_cap, cudaStream_t s,
                              bool reuse_split = false) {   // reuse_split: wsplit still holds this layer's split weights
    const int KK = l.ksize * l.ksize, K = l.in_c * KK, P = l.out_h * l.out_w;
    const int n_pad = (l.out_c + 15) / 16 * 16, k_pad = (K + kTcBK - 1) / kTcBK * kTcBK;
    if (n_pad > 128 || k_pad > 8192 || rows * (int64_t)P > INT32_MAX) return 1;    // one MMA spans 2*n_pad <= 256 columns
    const bool exact = A.u8 && (!A.normalize || (A.low == floorf(A.low) && fabsf(A.low) <= 1024.f));
    const size_t smem = conv_tc_smem_bytes(n_pad, k_pad, exact ? 1 : 2, 2, kTcBStages);
    if (smem > 200 * 1024 || wsplit == nullptr || conv_tc_wsplit_floats(l) > wsplit_cap) return 1;
    if (reinterpret_cast<uintptr_t>(wsplit) % 16 != 0) return 1;          // bulk copies need 16-byte aligned sources
    float *w_hi = wsplit, *w_lo = nullptr;                 // one interleaved hi|lo array
    uint32_t *koff = reinterpret_cast<uint32_t *>(wsplit + (size_t)2 * n_pad * k_pad);
    if (!reuse_split) {
        const int total = n_pad * k_pad;
        weight_split_kernel<<<(total + 255) / 256, 256, 0, s>>>(W, l.out_c, K, n_pad, k_pad, w_hi, w_lo, KK, l.ksize, l.in_h * l.in_w,
                                                                l.in_w, koff);
        B2RL_LAUNCH_CHECK();
    }
    ConvTcParams p;
    p.x = A.ptr; p.w_hi = w_hi; p.w_lo = w_lo; p.koff = koff; p.bias = bias; p.out = out; p.pre_out = pre_out; p.gather = A.row.gather;
    p.in_bstride = (int64_t)l.in_c * l.in_h * l.in_w;
    p.M = (int)(rows * P); p.N = l.out_c; p.K = K; p.n_pad = n_pad; p.k_pad = k_pad;
    p.P = P; p.OW = l.out_w; p.sy = l.stride * l.in_w; p.sx = l.stride;
    p.KK = KK; p.KS = l.ksize; p.HW = l.in_h * l.in_w; p.W = l.in_w;
    p.act = l.act; p.normalize = A.normalize; p.low = A.normalize ? A.low : 0.f; p.high = A.normalize ? A.high : 1.f;
    p.pad = 0; p.IH = l.in_h; p.cls_s = 1; p.out_H = l.out_h; p.out_W = l.out_w; p.w_class_stride = 0;
    p.dbg = tc_debug_buffer(); p.dbg_cta = tc_debug_cta();
    // 4 consecutive taps are 4 contiguous, 4-byte aligned bytes when the kernel width, the column
    // stride, the row pitch and the plane / image sizes are all multiples of 4
    bool vec;
    if (A.u8)
        vec = l.ksize % 4 == 0 && l.stride % 4 == 0 && l.in_w % 4 == 0 && (l.in_h * l.in_w) % 4 == 0 &&
              reinterpret_cast<uintptr_t>(A.ptr) % 4 == 0;
    else   // fp32: every 4-tap chunk starts on an even float (8-byte aligned) -> two 8-byte loads
        vec = l.ksize % 4 == 0 && l.stride % 2 == 0 && l.in_w % 2 == 0 && (l.in_h * l.in_w) % 2 == 0 &&
              p.in_bstride % 2 == 0 && reinterpret_cast<uintptr_t>(A.ptr) % 8 == 0;
    const int grid = (p.M + kTcBM - 1) / kTcBM;
    auto launch = [&](auto kern) -> int {
        { const int rca = ensure_big_smem(kern); if (rca != B2RL_OK) return rca; }
        kern<<<grid, kTcFwdThreads, smem, s>>>(p);
        B2RL_LAUNCH_CHECK();
        ++g_conv_path[0];
        return B2RL_OK;
    };
    switch (A.elem_kind()) {
        case EL_U8:
            if (exact) return vec ? launch(conv_fwd_tc_kernel<EL_U8, true, true, 8>) : launch(conv_fwd_tc_kernel<EL_U8, true, false, 2>);
            return vec ? launch(conv_fwd_tc_kernel<EL_U8, false, true, 8>) : launch(conv_fwd_tc_kernel<EL_U8, false, false, 2>);
        case EL_F32_NORM:
            return vec ? launch(conv_fwd_tc_kernel<EL_F32_NORM, false, true, 2>) : launch(conv_fwd_tc_kernel<EL_F32_NORM, false, false, 2>);
        default:
            return vec ? launch(conv_fwd_tc_kernel<EL_F32, false, true, 4>) : launch(conv_fwd_tc_kernel<EL_F32, false, false, 2>);
    }
}


// ------------------------------------------------------------------------------------------------
// Input gradient on tcgen05.  For stride s the input pixels split into s*s parity classes
// (y mod s, x mod s); the pixels of one class receive contributions only from the kernel taps
// ky = py + s*a, kx = px + s*b, so each class is a dense stride-1 convolution of dL/dout with a
// T x T kernel (T = ceil(k/s)) and zero padding T-1:
//   dX[b, ci, s*yy+py, s*xx+px] = sum_{co,a',b'} dOut[b, co, yy-(T-1)+a', xx-(T-1)+b'] * W[co, ci, py+s*(T-1-a'), px+s*(T-1-b')]
// — the forward kernel with a bounds-checked gather (PAD) and a strided output scatter; gridDim.y
// walks the classes.  No atomics, no memset, deterministic.
// ------------------------------------------------------------------------------------------------
__global__ void dgrad_weight_split_kernel(const float *__restrict__ w, int Cout, int Cin, int KS, int S, int T, int n_pad,
                                          int k_pad, int in_plane, int in_w, float *__restrict__ w_hi,
                                          float *__restrict__ w_lo, uint32_t *__restrict__ koff) {
    const int cls = blockIdx.y, py = cls / S, px = cls % S;
    const int K = Cout * T * T, total = n_pad * k_pad;
    float *hl_c = w_hi + (int64_t)cls * 2 * total;          // [2*n_pad x k_pad] hi|lo tiles of this class
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = e / n_pad, n = e - k * n_pad;          // k = (co, a', b'), n = ci
        float v = 0.f;
        int co = 0, a = 0, b = 0;
        if (k < K) {
            co = k / (T * T);
            const int rem = k - co * T * T;
            a = rem / T; b = rem - a * T;
            const int ky = py + S * (T - 1 - a), kx = px + S * (T - 1 - b);
            if (n < Cin && ky < KS && kx < KS) v = w[(((int64_t)co * Cin + n) * KS + ky) * KS + kx];
        }
        const float hi = tc::tf32_rn(v);
        const int kb = k / kTcBK, kin = k - kb * kTcBK;
        const int64_t o = (int64_t)kb * 2 * n_pad * kTcBK + (int64_t)(kin >> 2) * 2 * n_pad * 4 + (n >> 3) * 32 + (n & 7) * 4 + (kin & 3);
        hl_c[o] = hi;
        hl_c[o + (int64_t)n_pad * 4] = v - hi;
        if (n == 0 && cls == 0)
            koff[k] = k < K ? ((uint32_t)(co * in_plane + a * in_w + b) | ((uint32_t)a << 24) | ((uint32_t)b << 28)) : 0u;
    }
}

static inline size_t conv_dgrad_tc_scratch_floats(const b2rl_layer &l) {
    const int T = (l.ksize + l.stride - 1) / l.stride;
    const int K = l.out_c * T * T;
    const int n_pad = (l.in_c + 15) / 16 * 16, k_pad = (K + kTcBK - 1) / kTcBK * kTcBK;
    return (size_t)l.stride * l.stride * 2 * n_pad * k_pad + k_pad;
}

// dL/d(layer input) [rows, Cin, in_h, in_w] (overwritten) from g = dL/d(layer output).
// returns B2RL_OK, or 1 when the shape is outside what this path handles.
static int launch_conv_dgrad_tc(const b2rl_layer &l, const float *g, const float *W, float *g_in, int64_t rows,
                                float *scratch, size_t scratch_cap, cudaStream_t s) {
    const int S = l.stride, T = (l.ksize + S - 1) / S;
    const int K = l.out_c * T * T;
    const int n_pad = (l.in_c + 15) / 16 * 16, k_pad = (K + kTcBK - 1) / kTcBK * kTcBK;
    const int Yc = (l.in_h + S - 1) / S, Xc = (l.in_w + S - 1) / S;
    const int64_t M = rows * Yc * Xc;
    if (n_pad > 128 || k_pad > 8192 || M > INT32_MAX || T > 15 || S * S > 64) return 1;
    if ((int64_t)l.out_c * l.out_h * l.out_w >= (1 << 24)) return 1;        // offset field of the tap table
    const size_t smem = conv_tc_smem_bytes(n_pad, k_pad, 2);
    if (smem > 200 * 1024 || scratch == nullptr || conv_dgrad_tc_scratch_floats(l) > scratch_cap) return 1;
    if (reinterpret_cast<uintptr_t>(scratch) % 16 != 0) return 1;
    const size_t cls_floats = (size_t)n_pad * k_pad;
    float *w_hi = scratch, *w_lo = nullptr;                // per class: one interleaved hi|lo array of 2*cls_floats
    uint32_t *koff = reinterpret_cast<uint32_t *>(scratch + (size_t)2 * S * S * cls_floats);
    dgrad_weight_split_kernel<<<dim3((unsigned)((cls_floats + 255) / 256), S * S), 256, 0, s>>>(
        W, l.out_c, l.in_c, l.ksize, S, T, n_pad, k_pad, l.out_h * l.out_w, l.out_w, w_hi, w_lo, koff);
    B2RL_LAUNCH_CHECK();
    ConvTcParams p;
    p.x = g; p.w_hi = w_hi; p.w_lo = w_lo; p.koff = koff; p.bias = nullptr; p.out = g_in; p.pre_out = nullptr; p.gather = nullptr;
    p.in_bstride = (int64_t)l.out_c * l.out_h * l.out_w;
    p.M = (int)M; p.N = l.in_c; p.K = K; p.n_pad = n_pad; p.k_pad = k_pad;
    p.P = Yc * Xc; p.OW = Xc; p.sy = l.out_w; p.sx = 1;
    p.KK = T * T; p.KS = T; p.HW = l.out_h * l.out_w; p.W = l.out_w;
    p.act = B2RL_ACT_NONE; p.normalize = 0; p.low = 0.f; p.high = 1.f;
    p.pad = T - 1; p.IH = l.out_h; p.cls_s = S; p.out_H = l.in_h; p.out_W = l.in_w;
    p.w_class_stride = (int64_t)2 * cls_floats;
    p.dbg = nullptr; p.dbg_cta = -1;
    auto kern = conv_fwd_tc_kernel<EL_F32, false, false, 2, true>;
    { const int rca = ensure_big_smem(kern); if (rca != B2RL_OK) return rca; }
    kern<<<dim3((unsigned)((p.M + kTcBM - 1) / kTcBM), S * S), kTcFwdThreads, smem, s>>>(p);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

// ------------------------------------------------------------------------------------------------
// Weight gradient on tcgen05:   dW[co][tap] = sum_pix  im2col(x)[pix][tap] * G[pix][co]
//
// UMMA view: D[M = 128 taps][N = Cout] += A[M][K = pixels] * B[N][K], both operands K-major
// (pixels contiguous).  A thread owns one pixel (coalesced, packed gather exactly as in the forward
// kernel) and scatters its 16 taps into the K-major tile with 4-byte stores; the K-chunk pitch is
// padded by 16 B (LBO = 128*16 + 16) so that the 32 pixel-lanes of a warp hit 32 different banks.
// (An MN-major im2col operand would allow 16-byte stores instead; not yet validated — round 2.)
// Pixels are split over gridDim.y CTAs; the fp32 partials are reduced in fixed order by
// splitk_reduce_kernel.
//
//   A (K-major, no swizzle):  (tap m, pixel k) at (k/4)*LBO_a + (m/8)*128 + (m%8)*16 + (k%4)*4,  LBO_a = 2064
//   B (K-major, no swizzle):  (co n,  pixel k) at (k/4)*LBO_b + (n/8)*128 + (n%8)*16 + (k%4)*4
// ------------------------------------------------------------------------------------------------
struct ConvWgradTcParams {
    const void *x;              // layer input: uint8 frames / fp32 activations (NCHW)
    const float *g;             // dL/d(layer output), NCHW [rows, N, P]
    float *partial;             // [splits][Mtaps (+1)][N]
    const int64_t *gather;      // optional ring rows per batch row
    int64_t in_bstride;
    int Mtaps, N, Kpix;         // taps (Cin*k*k), Cout, rows*P
    int n_pad;
    int P, OW, sy, sx;
    int KK, KS, HW, W;
    int pix_per_cta;            // multiple of 32
    int normalize;
    int with_bias;              // partial has Mtaps+1 rows; the last one carries the bias gradient (N <= 32)
    float low, high;
    float inv_ow, inv_p;        // 1 / OW, 1 / P (pixel decode without integer division; Kpix < 2^22)
};

constexpr uint32_t kWgLboA = kTcBM * 16 + 16;        // padded K-chunk pitch of the wgrad A tile
constexpr uint32_t kWgABytes = (kTcBK / 4) * kWgLboA;  // 8 chunks

static inline size_t conv_wgrad_tc_smem_bytes(int n_pad, int a_parts) {
    return (size_t)kTcStages * ((size_t)a_parts * kWgABytes + 2 * (size_t)n_pad * kTcBK * 4) + kTcBM * 4 + 256 * 4 + 64 + 1024 + 256;
}

// Same thread roles as the forward kernel: 8 producer warps build both operand tiles, a ninth warp's
// elected lane issues the MMAs.  EXACT_A (uint8 frames, integer low): the im2col operand is the exact
// integer x - low (one tf32 part, 2 MMAs per k-step) and 1/(high-low) scales the fp32 partial sums.
template <int ELEM, bool EXACT_A, bool VEC, int D>
__global__ void __launch_bounds__(kTcThreads + 32) conv_wgrad_tc_kernel(const ConvWgradTcParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int RAWN = (ELEM == EL_U8 && VEC) ? 4 : 16;
    constexpr int CH = 4;                                   // 4-tap chunks per thread (16 taps)
    const int tid = threadIdx.x, warp = tid >> 5;
    const int pl = tid & 31, tg = (tid >> 5) & 7;            // pixel lane within the 32-pixel block, tap group (16 taps)
    const uint32_t a_bytes = kWgABytes, b_bytes = (uint32_t)p.n_pad * kTcBK * 4;
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t stage_bytes = (EXACT_A ? 1 : 2) * a_bytes + 2 * b_bytes;
    auto a_hi = [&](int s) { return sbase + (uint32_t)s * stage_bytes; };
    auto a_lo = [&](int s) { return sbase + (uint32_t)s * stage_bytes + a_bytes; };     // unused when EXACT_A
    auto b_hi = [&](int s) { return sbase + (uint32_t)s * stage_bytes + (EXACT_A ? 1 : 2) * a_bytes; };
    auto b_lo = [&](int s) { return sbase + (uint32_t)s * stage_bytes + (EXACT_A ? 1 : 2) * a_bytes + b_bytes; };
    const uint32_t toff_a = sbase + kTcStages * stage_bytes;              // tap offsets of this CTA's 128 taps
    const uint32_t lut_a = toff_a + kTcBM * 4;
    const uint32_t bars_a = (lut_a + 256 * 4 + 15u) & ~15u;
    const uint32_t tptr_a = bars_a + 8 * 2 * kTcStages;
    uint8_t *gen = smem_raw + (sbase - tc::smem_u32(smem_raw));
    uint64_t *mma_bar = reinterpret_cast<uint64_t *>(gen + (bars_a - sbase));   // [kTcStages] MMA group done
    uint64_t *full = mma_bar + kTcStages;                                        // [kTcStages] both tiles written
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(gen + (tptr_a - sbase));

    const int tap0 = blockIdx.x * kTcBM;
    if (tid < kTcBM) {
        const int k = tap0 + tid;
        uint32_t off = 0;                    // tile rows past Mtaps compute garbage that is never stored
        if (k < p.Mtaps) {
            const int ci = k / p.KK, rem = k - ci * p.KK;
            const int ky = rem / p.KS, kx = rem - ky * p.KS;
            off = (uint32_t)(ci * p.HW + ky * p.W + kx);
        }
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(toff_a + 4u * tid), "r"(off) : "memory");
    }
    if (ELEM == EL_U8 && !EXACT_A)
        for (int i = tid; i < 256; i += kTcThreads + 32) {
            const float v = p.normalize ? __fdiv_rn((float)i - p.low, p.high - p.low) : (float)i;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(lut_a + 4u * i), "f"(v) : "memory");
        }
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < p.n_pad) tmem_cols <<= 1;
    if (warp == 0) tc::tmem_alloc(tmem_ptr, tmem_cols);
    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            tc::mbar_init(&mma_bar[s], 1);
            tc::mbar_init(&full[s], kTcThreads / 32);
        }
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;

    const int pix0 = blockIdx.y * p.pix_per_cta;
    const int pix1 = min(p.Kpix, pix0 + p.pix_per_cta);
    const int KB = (pix1 - pix0 + kTcBK - 1) / kTcBK;
    const uint32_t idesc = tc::make_idesc_tf32(kTcBM, p.n_pad);
    const uint32_t lbo_b = (uint32_t)p.n_pad * 16;

    if (warp == kTcThreads / 32) {
        // ---- MMA warp (converged loop, one elected lane issues) ------------------------------------------
        const uint64_t da_step = (uint64_t)((2 * kWgLboA) >> 4), db_step = (uint64_t)((2 * lbo_b) >> 4);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb & (kTcStages - 1);
            tc::mbar_wait(&full[s], (uint32_t)((kb / kTcStages) & 1));
            tc::tc_fence_after();
            const uint64_t dah0 = tc::make_desc(a_hi(s), kWgLboA, 128), dal0 = tc::make_desc(a_lo(s), kWgLboA, 128);
            const uint64_t dbh0 = tc::make_desc(b_hi(s), lbo_b, 128), dbl0 = tc::make_desc(b_lo(s), lbo_b, 128);
            if (tc::elect_one()) {
#pragma unroll
                for (int j = 0; j < kTcBK / 8; ++j) {                   // 8 pixels per MMA
                    tc::mma_tf32(tmem_d, dah0 + j * da_step, dbh0 + j * db_step, idesc, (kb | j) ? 1u : 0u);
                    if (!EXACT_A) tc::mma_tf32(tmem_d, dal0 + j * da_step, dbh0 + j * db_step, idesc, 1u);
                    tc::mma_tf32(tmem_d, dah0 + j * da_step, dbl0 + j * db_step, idesc, 1u);
                }
                tc::mma_commit(&mma_bar[s]);
            }
            __syncwarp();
        }
    } else {
        // ---- producer warps ---------------------------------------------------------------------------
        const int64_t safe_base = p.gather ? p.gather[0] * p.in_bstride : 0;
        const float exact_bias = 8388608.f + p.low;
        const int n = (tid >> 3) & 31, q = tid & 7;             // B: channel n (+32 per pass), pixels q*4..q*4+3
        uint32_t raw[D][RAWN];
        float gval[D][4];
        float bsum = 0.f;                                       // bias gradient of channel n over this CTA's pixels
        auto gather = [&](int kb, uint32_t (&dst)[RAWN], float (&gv)[4]) {
            // ---- A: this thread's pixel, taps [tg*16, tg*16+16); pixels past pix1 meet zero G values
            const int pix = min(pix0 + kb * kTcBK + pl, pix1 - 1);
            int64_t base;
            {
                int b = __float2int_rz(((float)pix + 0.5f) * p.inv_p), pp = pix - b * p.P;
                if (pp < 0) { --b; pp += p.P; } else if (pp >= p.P) { ++b; pp -= p.P; }      // reciprocal off by one
                const int oy = __float2int_rz(((float)pp + 0.5f) * p.inv_ow), ox = pp - oy * p.OW;
                const int64_t bb = p.gather ? p.gather[b] : (int64_t)b;
                base = bb * p.in_bstride + (int64_t)(oy * p.sy + ox * p.sx);
            }
            if (ELEM == EL_U8 && VEC) {
                const uint8_t *rp = static_cast<const uint8_t *>(p.x) + base;
#pragma unroll
                for (int c = 0; c < CH; ++c)
                    dst[c] = __ldg(reinterpret_cast<const uint32_t *>(rp + tc::lds32(toff_a + 4u * (tg * 16 + c * 4))));
            } else {
#pragma unroll
                for (int j = 0; j < CH * 4; ++j) {
                    const uint32_t off = tc::lds32(toff_a + 4u * (tg * 16 + j));
                    if (ELEM == EL_U8) dst[j] = (uint32_t)__ldg(static_cast<const uint8_t *>(p.x) + base + off);
                    else dst[j] = __float_as_uint(__ldg(static_cast<const float *>(p.x) + base + off));
                }
            }
            // ---- B: 4 consecutive pixels of channel n (only the first 32 channels are prefetched)
            const int px0 = pix0 + kb * kTcBK + q * 4;
            int b = __float2int_rz(((float)px0 + 0.5f) * p.inv_p), pp = px0 - b * p.P;
            if (pp < 0) { --b; pp += p.P; } else if (pp >= p.P) { ++b; pp -= p.P; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                gv[j] = (n < p.N && px0 + j < pix1) ? __ldg(p.g + ((int64_t)b * p.N + n) * p.P + pp) : 0.f;
                if (++pp == p.P) { pp = 0; ++b; }
            }
        };
#pragma unroll
        for (int d = 0; d < D; ++d)
            if (d < KB) gather(d, raw[d], gval[d]);

        auto step = [&](int kb, uint32_t (&cur)[RAWN], float (&gv)[4]) {
            const int s = kb & (kTcStages - 1);
            if (kb >= kTcStages) tc::mbar_wait(&mma_bar[s], (uint32_t)((kb / kTcStages - 1) & 1));
            // ---- A tile stores (transposed scatter): (k/4)*LBO_a + (m/8)*128 + (m%8)*16 + (k%4)*4, k = pl
            const uint32_t a_off = (uint32_t)(pl >> 2) * kWgLboA + (uint32_t)(pl & 3) * 4 + (uint32_t)(tg * 2) * 128;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int kk = c * 4 + j;
                    float v;
                    if (ELEM == EL_U8) {
                        if (EXACT_A) {
                            const uint32_t w = VEC ? __byte_perm(cur[c], 0x4B000000u, 0x7650u + j) : (cur[kk] | 0x4B000000u);
                            v = __uint_as_float(w) - exact_bias;
                        } else {
                            const uint32_t byte = VEC ? ((cur[c] >> (8 * j)) & 0xFFu) : cur[kk];
                            v = __uint_as_float(tc::lds32(lut_a + 4u * byte));
                        }
                    } else if (ELEM == EL_F32_NORM) v = __fdiv_rn(__uint_as_float(cur[kk]) - p.low, p.high - p.low);
                    else v = __uint_as_float(cur[kk]);
                    const uint32_t o = a_off + (uint32_t)(kk >> 3) * 128 + (uint32_t)(kk & 7) * 16;
                    if (EXACT_A) {
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a_hi(s) + o), "f"(v) : "memory");
                    } else {
                        const float hi = tc::tf32_rn_fast(v), lo = v - hi;
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a_hi(s) + o), "f"(hi) : "memory");
                        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a_lo(s) + o), "f"(lo) : "memory");
                    }
                }
            }
            // ---- B tile stores: (k/4)*LBO_b + (n/8)*128 + (n%8)*16
            {
                float hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { hi[j] = tc::tf32_rn_fast(gv[j]); lo[j] = gv[j] - hi[j]; bsum += gv[j]; }
                const uint32_t o = (uint32_t)q * lbo_b + (uint32_t)(n >> 3) * 128 + (uint32_t)(n & 7) * 16;
                if (n < p.n_pad) {          // n_pad may be 16: rows beyond it belong to the next K chunk
                    tc::sts128(b_hi(s) + o, hi[0], hi[1], hi[2], hi[3]);
                    tc::sts128(b_lo(s) + o, lo[0], lo[1], lo[2], lo[3]);
                }
                for (int n2 = n + 32; n2 < p.n_pad; n2 += 32) {        // wider layers: remaining channels
                    const int px0 = pix0 + kb * kTcBK + q * 4;
                    int b = px0 / p.P, pp = px0 - b * p.P;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float v = (n2 < p.N && px0 + j < pix1) ? __ldg(p.g + ((int64_t)b * p.N + n2) * p.P + pp) : 0.f;
                        if (++pp == p.P) { pp = 0; ++b; }
                        hi[j] = tc::tf32_rn_fast(v); lo[j] = v - hi[j];
                    }
                    const uint32_t o2 = (uint32_t)q * lbo_b + (uint32_t)(n2 >> 3) * 128 + (uint32_t)(n2 & 7) * 16;
                    tc::sts128(b_hi(s) + o2, hi[0], hi[1], hi[2], hi[3]);
                    tc::sts128(b_lo(s) + o2, lo[0], lo[1], lo[2], lo[3]);
                }
            }
            if (kb + D < KB) gather(kb + D, cur, gv);
            tc::fence_async_smem();
            __syncwarp();
            if ((tid & 31) == 0) tc::mbar_arrive(&full[s]);
        };
        for (int kb0 = 0; kb0 < KB; kb0 += D) {
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (kb0 + d < KB) step(kb0 + d, raw[d], gval[d]);
        }
        for (int s = 0; s < kTcStages; ++s) {
            const int uses = (KB - s + kTcStages - 1) / kTcStages;
            if (uses > 0) tc::mbar_wait(&mma_bar[s], (uint32_t)((uses - 1) & 1));
        }
        tc::tc_fence_after();
        // ---- bias gradient partial (row Mtaps of the partial matrix): sum the 8 pixel-chunk lanes of a channel
        if (p.with_bias && blockIdx.x == 0) {
            bsum += __shfl_xor_sync(0xffffffffu, bsum, 1);
            bsum += __shfl_xor_sync(0xffffffffu, bsum, 2);
            bsum += __shfl_xor_sync(0xffffffffu, bsum, 4);
            if (q == 0 && n < p.N) p.partial[((int64_t)blockIdx.y * (p.Mtaps + 1) + p.Mtaps) * p.N + n] = bsum;
        }
        // ---- epilogue: partial[z][tap][co]; warp w reads TMEM lanes 32*(w%4).., 16-column groups of parity w/4
        {
            const int qd = warp & 3;
            const int tap = tap0 + qd * 32 + (tid & 31);
            const float scale = (EXACT_A && p.normalize) ? 1.0f / (p.high - p.low) : 1.0f;
            for (int c0 = (warp >> 2) * 16; c0 < p.n_pad; c0 += 32) {
                uint32_t r[16];
                if (KB > 0) tc::tmem_ld16(tmem_d + ((uint32_t)(qd * 32) << 16) + (uint32_t)c0, r);
                if (tap < p.Mtaps) {
                    float *o = p.partial + ((int64_t)blockIdx.y * (p.Mtaps + p.with_bias) + tap) * p.N + c0;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < p.N) o[j] = KB > 0 ? __uint_as_float(r[j]) * scale : 0.f;
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, tmem_cols);
}

// db[co] = sum over rows and pixels of G[b][co][pix]   (one CTA per channel, fixed-order reduction)
__global__ void conv_bias_grad_kernel(const float *__restrict__ g, int64_t rows, int N, int P, float *__restrict__ db,
                                      int accumulate) {
    __shared__ float red[32];
    const int n = blockIdx.x;
    float s = 0.f;
    const int64_t total = rows * P;
    for (int64_t e = threadIdx.x; e < total; e += blockDim.x) {
        const int64_t b = e / P, pp = e - b * P;
        s += g[(b * N + n) * P + pp];
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        db[n] = accumulate ? db[n] + t : t;
    }
}

static inline size_t conv_wgrad_tc_partial_floats(const b2rl_layer &l, int64_t rows, int sms) {
    const int Kc = l.in_c * l.ksize * l.ksize;
    const int64_t Kpix = rows * l.out_h * l.out_w;
    const int mt = (Kc + kTcBM - 1) / kTcBM;
    int64_t splits = (2 * (int64_t)sms + mt - 1) / mt;
    const int64_t max_splits = (Kpix + 4 * kTcBK - 1) / (4 * kTcBK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    return (size_t)splits * (Kc + 1) * l.out_c;
}

// returns B2RL_OK, or 1 for shapes the tensor-core kernel does not take.
static int launch_conv_wgrad_tc(const b2rl_layer &l, const Operand &X, const float *g, float *dw, float *db,
                                int accumulate, int64_t rows, float *partial, size_t partial_cap, cudaStream_t s) {
    const int KK = l.ksize * l.ksize, Kc = l.in_c * KK, P = l.out_h * l.out_w;
    const int n_pad = (l.out_c + 15) / 16 * 16;
    const int64_t Kpix = rows * P;
    if (n_pad > 256 || Kpix >= (1 << 22) || partial == nullptr) return 1;     // float-reciprocal pixel decode is exact below 2^22
    const int mt = (Kc + kTcBM - 1) / kTcBM;
    int64_t splits = (2 * (int64_t)sm_count() + mt - 1) / mt;
    const int64_t max_splits = (Kpix + 4 * kTcBK - 1) / (4 * kTcBK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int pix_per_cta = (int)(((Kpix + splits - 1) / splits + kTcBK - 1) / kTcBK * kTcBK);
    splits = (Kpix + pix_per_cta - 1) / pix_per_cta;
    const int with_bias = l.out_c <= 32 ? 1 : 0;
    if ((size_t)splits * (Kc + 1) * l.out_c > partial_cap) return 1;
    const bool exact = X.u8 && (!X.normalize || (X.low == floorf(X.low) && fabsf(X.low) <= 1024.f));
    const size_t smem = conv_wgrad_tc_smem_bytes(n_pad, exact ? 1 : 2);
    if (smem > 200 * 1024) return 1;
    ConvWgradTcParams p;
    p.x = X.ptr; p.g = g; p.partial = partial; p.gather = X.red.gather;
    p.in_bstride = (int64_t)l.in_c * l.in_h * l.in_w;
    p.Mtaps = Kc; p.N = l.out_c; p.Kpix = (int)Kpix; p.n_pad = n_pad;
    p.P = P; p.OW = l.out_w; p.sy = l.stride * l.in_w; p.sx = l.stride;
    p.KK = KK; p.KS = l.ksize; p.HW = l.in_h * l.in_w; p.W = l.in_w;
    p.pix_per_cta = pix_per_cta;
    p.with_bias = with_bias;
    p.normalize = X.normalize; p.low = X.normalize ? X.low : 0.f; p.high = X.normalize ? X.high : 1.f;
    p.inv_ow = 1.0f / (float)l.out_w; p.inv_p = 1.0f / (float)P;
    const bool vec = X.u8 && l.ksize % 4 == 0 && l.stride % 4 == 0 && l.in_w % 4 == 0 && (l.in_h * l.in_w) % 4 == 0 &&
                     reinterpret_cast<uintptr_t>(X.ptr) % 4 == 0;
    dim3 grid(mt, (unsigned)splits);
    auto launch = [&](auto kern) -> int {
        { const int rca = ensure_big_smem(kern); if (rca != B2RL_OK) return rca; }
        kern<<<grid, kTcThreads + 32, smem, s>>>(p);
        B2RL_LAUNCH_CHECK();
        return B2RL_OK;
    };
    int rc;
    switch (X.elem_kind()) {
        case EL_U8:
            if (exact) rc = vec ? launch(conv_wgrad_tc_kernel<EL_U8, true, true, 4>) : launch(conv_wgrad_tc_kernel<EL_U8, true, false, 2>);
            else rc = vec ? launch(conv_wgrad_tc_kernel<EL_U8, false, true, 4>) : launch(conv_wgrad_tc_kernel<EL_U8, false, false, 2>);
            break;
        case EL_F32_NORM: rc = launch(conv_wgrad_tc_kernel<EL_F32_NORM, false, false, 2>); break;
        default: rc = launch(conv_wgrad_tc_kernel<EL_F32, false, false, 2>); break;
    }
    if (rc != B2RL_OK) return rc;
    // fixed-order reduction of the pixel splits, transposed into dW[co][tap] (+ db from the extra row)
    Epilogue epi;
    epi.kind = EPI_WGRAD_T; epi.out = dw; epi.db = with_bias ? db : nullptr; epi.wcols = Kc; epi.accumulate = accumulate;
    launch_splitk_reduce<EpiTraits<EPI_WGRAD_T, MAP_STRIDE, MAP_STRIDE>>(partial, (int)splits, Kc + with_bias, l.out_c, epi, s);
    B2RL_LAUNCH_CHECK();
    if (!with_bias) {
        conv_bias_grad_kernel<<<l.out_c, 1024, 0, s>>>(g, rows, l.out_c, P, db, accumulate);
        B2RL_LAUNCH_CHECK();
    }
    return B2RL_OK;
}

}  // namespace b2rl
