// conv_tc_persist.cuh — persistent variant of conv_fwd_tc_kernel.
//
// The one-tile-per-CTA kernel pays, for 128 output pixels, a CTA launch, a TMEM allocation, the
// barrier / table setup, one fully exposed gather latency (every warp of the CTA waits for its first
// taps at the same time) and an epilogue during which its producers idle.  Here a CTA stays resident
// and walks tiles  blockIdx.x, blockIdx.x + gridDim.x, ...:
//   * 8 producer warps run one continuous stream of k-blocks across tile boundaries: the register
//     ring that prefetches D k-blocks ahead simply runs into the next tile, so a gather latency is
//     exposed once per CTA, not once per tile;
//   * the MMA warp alternates between two TMEM accumulators;
//   * 4 epilogue warps drain accumulator i (TMEM -> bias/activation -> NCHW) while the producers and
//     the tensor core already work on tile i+1.
// Operand layouts, the 3xTF32 / exact-integer arithmetic, the weight-tile TMA ring and the PAD
// (input-gradient) mode are exactly those of conv_fwd_tc_kernel, so results are bit-identical to it.
#pragma once
// (included from the middle of conv_tc.cuh, after conv_fwd_tc_kernel and before its launchers)

namespace b2rl {

constexpr int kPsProducerThreads = kTcThreads;                      // 8 warps
constexpr int kPsMmaWarp = kPsProducerThreads / 32;                // warp 8
constexpr int kPsEpiWarp0 = kPsMmaWarp + 1;                        // warps 9..12
constexpr int kPsThreads = kPsProducerThreads + 32 + 128;

static inline size_t conv_tc_persist_smem_bytes(int n_pad, int k_pad, int a_parts, int a_stages) {
    return conv_tc_smem_bytes(n_pad, k_pad, a_parts, a_stages) + 64;      // + accumulator barriers
}

template <int ELEM, bool EXACT_A, bool VEC, int D, bool PAD, int SA>
__global__ void __launch_bounds__(kPsThreads) conv_fwd_tc_persist_kernel(const ConvTcParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    constexpr int RAWN = (ELEM == EL_U8 && VEC) ? 4 : 16;      // raw words per thread per k-block
    constexpr int CH = kTcBK / 4 / 2;                          // 4-tap chunks per thread per k-block (4)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t a_bytes = kTcBM * kTcBK * 4, b_bytes = (uint32_t)p.n_pad * kTcBK * 4;
    const uint32_t sbase = (tc::smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t a_stage = (EXACT_A ? 1 : 2) * a_bytes;
    auto a_hi = [&](int s) { return sbase + (uint32_t)s * a_stage; };
    auto a_lo = [&](int s) { return sbase + (uint32_t)s * a_stage + a_bytes; };          // unused when EXACT_A
    const uint32_t bbase = sbase + SA * a_stage;
    auto b_hi = [&](int s) { return bbase + (uint32_t)s * 2 * b_bytes; };
    auto b_lo = [&](int s) { return bbase + (uint32_t)s * 2 * b_bytes + b_bytes; };
    const uint32_t koff_a = bbase + kTcBStages * 2 * b_bytes;
    const uint32_t lut_a = koff_a + (uint32_t)p.k_pad * 4;
    const uint32_t bias_a = lut_a + 256 * 4;
    const uint32_t bars_a = (bias_a + (uint32_t)p.n_pad * 4 + 15u) & ~15u;
    const uint32_t tptr_a = bars_a + 8 * (2 * SA + kTcBStages + 4);
    uint8_t *gen = smem_raw + (sbase - tc::smem_u32(smem_raw));        // generic alias of sbase (barriers only)
    uint64_t *mma_bar = reinterpret_cast<uint64_t *>(gen + (bars_a - sbase));   // [SA] MMA group done -> A stage free
    uint64_t *full_a = mma_bar + SA;                                     // [SA] im2col tile written
    uint64_t *full_b = full_a + SA;                                      // [kTcBStages] weight tile landed
    uint64_t *acc_full = full_b + kTcBStages;                            // [2] accumulator complete
    uint64_t *acc_empty = acc_full + 2;                                  // [2] accumulator drained
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(gen + (tptr_a - sbase));

    const int KB = p.k_pad / kTcBK;
    const int mtiles = (p.M + kTcBM - 1) / kTcBM;
    const int ntiles = mtiles * (PAD ? p.cls_s * p.cls_s : 1);
    const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total = my_tiles * KB;                          // k-blocks this CTA produces / consumes

    // ---- one-time setup ----------------------------------------------------------------------
    if (tid == 0) {
        for (int s = 0; s < SA; ++s) {
            tc::mbar_init(&mma_bar[s], 1);
            tc::mbar_init(&full_a[s], kPsProducerThreads / 32);      // one elected arrive per producer warp
        }
        for (int s = 0; s < kTcBStages; ++s) tc::mbar_init(&full_b[s], 1);
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&acc_full[s], 1);
            tc::mbar_init(&acc_empty[s], 4);                         // one elected arrive per epilogue warp
        }
        tc::fence_barrier_init();
    }
    for (int k = tid; k < p.k_pad; k += kPsThreads)
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(koff_a + 4u * k), "r"(__ldg(p.koff + k)) : "memory");
    if (ELEM == EL_U8 && !EXACT_A)
        for (int i = tid; i < 256; i += kPsThreads) {
            const float v = p.normalize ? __fdiv_rn((float)i - p.low, p.high - p.low) : (float)i;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(lut_a + 4u * i), "f"(v) : "memory");
        }
    for (int n = tid; n < p.n_pad; n += kPsThreads) {
        const float v = (n < p.N && p.bias) ? p.bias[n] : 0.f;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(bias_a + 4u * n), "f"(v) : "memory");
    }
    uint32_t acc_cols = 32;                                    // columns of one accumulator: [A.W_hi | A.W_lo]
    while ((int)acc_cols < 2 * p.n_pad) acc_cols <<= 1;
    if (warp == 0) tc::tmem_alloc(tmem_ptr, 2 * acc_cols);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_d = *tmem_ptr;
    const uint32_t idesc2 = tc::make_idesc_tf32(kTcBM, 2 * p.n_pad), idesc1 = tc::make_idesc_tf32(kTcBM, p.n_pad);
    const uint32_t lbo_a = kTcBM * 16, lbo_b = (uint32_t)p.n_pad * 2 * 16;
    // k-block g used A stage g % SA for the (g / SA)-th time
    auto wait_mma = [&](int g) { tc::mbar_wait(&mma_bar[g % SA], (uint32_t)((g / SA) & 1)); };

    if (warp == kPsMmaWarp) {
        // ================= MMA warp (converged loop, one elected lane issues) ==============================
        int bg = 0, b_it = 0, b_kb = 0;                    // cursor of the weight-tile copies
        auto issue_b_next = [&]() {
            const int sb = bg & (kTcBStages - 1);
            int64_t wo = (int64_t)b_kb * 2 * p.n_pad * kTcBK;
            if (PAD) wo += (int64_t)(((int)blockIdx.x + b_it * (int)gridDim.x) / mtiles) * p.w_class_stride;
            if (tc::elect_one()) {
                tc::mbar_expect_tx(&full_b[sb], 2 * b_bytes);
                tc::bulk_g2s(b_hi(sb), p.w_hi + wo, 2 * b_bytes, &full_b[sb]);      // hi|lo tile, one copy
            }
            __syncwarp();
            ++bg;
            if (++b_kb == KB) { b_kb = 0; ++b_it; }
        };
        if (total > 0) issue_b_next();
        if (total > 1) issue_b_next();
        const uint64_t da_step = (uint64_t)((2 * lbo_a) >> 4), db_step = (uint64_t)((2 * lbo_b) >> 4);
        int g = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it & 1;
            if (it >= 2) tc::mbar_wait(&acc_empty[buf], (uint32_t)(((it >> 1) - 1) & 1));
            tc::tc_fence_after();
            const uint32_t tacc = tmem_d + (uint32_t)buf * acc_cols;
            for (int kb = 0; kb < KB; ++kb, ++g) {
                const int s = g % SA, sb = g & (kTcBStages - 1);
                if (g + 2 < total) {
                    if (g >= 2) wait_mma(g - 2);          // ring slot (g+2)%4 was last read by MMA group g-2
                    issue_b_next();
                }
                tc::mbar_wait(&full_b[sb], (uint32_t)((g / kTcBStages) & 1));
                tc::mbar_wait(&full_a[s], (uint32_t)((g / SA) & 1));
                tc::tc_fence_after();
                const uint64_t dah0 = tc::make_desc(a_hi(s), lbo_a, 128), dal0 = tc::make_desc(a_lo(s), lbo_a, 128);
                const uint64_t db0 = tc::make_desc(b_hi(sb), lbo_b, 128);
                if (tc::elect_one()) {
#pragma unroll
                    for (int j = 0; j < kTcBK / 8; ++j) {      // one MMA k-step = 8 tf32 = 2 core-matrix columns
                        tc::mma_tf32(tacc, dah0 + j * da_step, db0 + j * db_step, idesc2, (kb | j) ? 1u : 0u);
                        if (!EXACT_A) tc::mma_tf32(tacc, dal0 + j * da_step, db0 + j * db_step, idesc1, 1u);
                    }
                    tc::mma_commit(&mma_bar[s]);
                }
                __syncwarp();
            }
            if (tc::elect_one()) tc::mma_commit(&acc_full[buf]);
            __syncwarp();
        }
    } else if (warp < kPsMmaWarp) {
        // ================= producer warps ===============================================================
        const int row = tid & (kTcBM - 1), half = tid >> 7;   // im2col row, which 16 of the 32 k
        const uint32_t row_off = (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
        const float exact_bias = 8388608.f + p.low;
        // gather cursor: runs D k-blocks ahead of the convert, across tile boundaries
        int g_it = 0, g_kb = 0;
        const uint8_t *row_u8 = static_cast<const uint8_t *>(p.x);
        const float *row_f32 = static_cast<const float *>(p.x);
        uint32_t mask_y = 0, mask_x = 0;
        auto decode = [&](int it) {
            const int T = (int)blockIdx.x + it * (int)gridDim.x;
            const int mt = PAD ? T % mtiles : T;
            const int m = mt * kTcBM + row;
            int64_t rowbase = p.gather ? p.gather[0] * p.in_bstride : 0;      // rows beyond M read a safe address
            mask_y = mask_x = 0;
            if (m < p.M) {
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
            row_u8 = static_cast<const uint8_t *>(p.x) + rowbase;
            row_f32 = static_cast<const float *>(p.x) + rowbase;
        };
        auto gather_next = [&](uint32_t (&dst)[RAWN]) {
            const int k0 = g_kb * kTcBK + half * (CH * 4);
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
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const uint32_t off = tc::lds32(koff_a + 4u * (k0 + c * 4));
                    dst[c] = __ldg(reinterpret_cast<const uint32_t *>(row_u8 + off));   // 4 packed taps
                }
            } else {
#pragma unroll
                for (int j = 0; j < CH * 4; ++j) {
                    const uint32_t off = tc::lds32(koff_a + 4u * (k0 + j));
                    if (ELEM == EL_U8) dst[j] = (uint32_t)__ldg(row_u8 + off);
                    else dst[j] = __float_as_uint(__ldg(row_f32 + off));
                }
            }
            if (++g_kb == KB) {
                g_kb = 0;
                if (++g_it < my_tiles) decode(g_it);
            }
        };
        uint32_t raw[D][RAWN];
        if (my_tiles > 0) decode(0);
#pragma unroll
        for (int d = 0; d < D; ++d)
            if (d < total) gather_next(raw[d]);

        int c_kb = 0;                                          // k-block within its tile of the convert cursor
        auto step = [&](int g, uint32_t (&cur)[RAWN]) {
            const int s = g % SA;
            if (g >= SA) wait_mma(g - SA);                     // A stage s is free again
            const uint32_t ah = a_hi(s) + row_off + (uint32_t)(half * CH) * lbo_a, al = a_lo(s) + row_off + (uint32_t)(half * CH) * lbo_a;
            const int k0 = c_kb * kTcBK + half * (CH * 4);
            // padded taps meet zero weights; only non-finite fp32 garbage could leak through them
            const bool tail = ELEM != EL_U8 && c_kb == KB - 1 && p.K != p.k_pad;
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
                    else { hi[j] = tc::tf32_rn(v); lo[j] = v - hi[j]; }
                }
                tc::sts128(ah + (uint32_t)c * lbo_a, hi[0], hi[1], hi[2], hi[3]);
                if (!EXACT_A) tc::sts128(al + (uint32_t)c * lbo_a, lo[0], lo[1], lo[2], lo[3]);
            }
            if (g + D < total) gather_next(cur);     // refill this register slot D k-blocks ahead (maybe next tile)
            tc::fence_async_smem();                  // generic-proxy smem writes -> visible to the async (tensor) proxy
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&full_a[s]);
            if (++c_kb == KB) c_kb = 0;
        };
        for (int g0 = 0; g0 < total; g0 += D) {
#pragma unroll
            for (int d = 0; d < D; ++d)
                if (g0 + d < total) step(g0 + d, raw[d]);
        }
    } else {
        // ================= epilogue warps ===============================================================
        const int q = warp & 3;                                // TMEM lane quarter this warp may read
        const bool relu = p.act == B2RL_ACT_RELU, ident = p.act == B2RL_ACT_NONE;
        const float scale = (EXACT_A && p.normalize) ? 1.0f / (p.high - p.low) : 1.0f;
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it & 1;
            const int T = (int)blockIdx.x + it * (int)gridDim.x;
            const int cls = PAD ? T / mtiles : 0;
            const int mt = PAD ? T - cls * mtiles : T;
            const int em = mt * kTcBM + q * 32 + lane;
            bool e_ok = em < p.M;
            int b_img = 0, pix = 0;
            if (e_ok) { b_img = em / p.P; pix = em - b_img * p.P; }
            int64_t out_P = p.P;
            if (PAD) {                                           // scatter into the class's pixels of the full plane
                const int yy = pix / p.OW, xx = pix - yy * p.OW;
                const int y = yy * p.cls_s + cls / p.cls_s, x = xx * p.cls_s + cls % p.cls_s;
                e_ok = e_ok && y < p.out_H && x < p.out_W;
                pix = y * p.out_W + x;
                out_P = (int64_t)p.out_H * p.out_W;
            }
            const int oP = (int)out_P;
            tc::mbar_wait(&acc_full[buf], (uint32_t)((it >> 1) & 1));
            tc::tc_fence_after();
            for (int c0 = 0; c0 < p.n_pad; c0 += 16) {
                uint32_t r[16], r2[16];
                tc::tmem_ld16(tmem_d + (uint32_t)buf * acc_cols + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
                tc::tmem_ld16(tmem_d + (uint32_t)buf * acc_cols + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.n_pad + c0), r2);
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
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_d, 2 * acc_cols);
}

// grid = resident CTAs (occupancy x SM count), capped by the number of tiles
template <typename Kern>
static int launch_conv_persist(Kern kern, const ConvTcParams &p, int classes, size_t smem, cudaStream_t s) {
    B2RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B2RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    // resident CTAs per SM from the kernel's own footprint (registers, shared memory incl. the 1 KB the
    // system reserves per CTA, threads)
    cudaFuncAttributes fa;
    B2RL_CUDA(cudaFuncGetAttributes(&fa, kern));
    int dev = 0, smem_sm = 0, regs_sm = 0;
    B2RL_CUDA(cudaGetDevice(&dev));
    B2RL_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    B2RL_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
    const int regs_cta = ((fa.numRegs * 32 + 255) / 256 * 256) * (kPsThreads / 32);
    int per_sm = regs_sm / (regs_cta > 0 ? regs_cta : 1);
    const int by_smem = (int)(smem_sm / (smem + fa.sharedSizeBytes + 1024));
    if (by_smem < per_sm) per_sm = by_smem;
    if (2048 / kPsThreads < per_sm) per_sm = 2048 / kPsThreads;
    if (per_sm < 1) return 1;
    const int ntiles = (p.M + kTcBM - 1) / kTcBM * classes;
    int grid = per_sm * sm_count();
    if (grid > ntiles) grid = ntiles;
    kern<<<grid, kPsThreads, smem, s>>>(p);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

}  // namespace b2rl
