// gemm.cuh — one fp32 "separable-index" GEMM engine for every dense contraction on the path.
//
//   C[outM(m) + outN(n)] = epi( sum_r  A[aRow(m) + aRed(r)] * B[bRow(n) + bRed(r)] )
//
// Every operand address is a SUM of a row term and a reduction term, each produced by a tiny
// index map (dense stride, im2col output-pixel, im2col kernel-tap).  That single form covers
//   conv forward   (A = im2col(frames), B = W)          conv wgrad (+bias grad as a ones column)
//   conv dgrad     (col2im scatter epilogue)            linear forward / dX / dW
// so there is exactly one inner loop to optimise.  The map kinds, element types and epilogue are
// COMPILE-TIME traits (round-1 profile: with run-time kind switches the kernel was 22.8k SASS
// instructions and stalled on instruction fetch, profiles/r1_conv1_igemm_before.txt).
//
// fp32 FFMA on CUDA cores with fp32 accumulate: the reference's losses are pinned at 1e-5 against
// an fp32 CPU run, which single-pass TF32/BF16 tensor-core MMA does not hold; the
// tcgen05 3xTF32 variant is the planned upgrade for the two conv contractions and slots in
// behind the same launch wrappers.
//
// Tiling: BM x BN x BK smem tiles, TM x TN register micro-tiles, register-prefetch double
// buffering (raw loads are issued before the FMA block, converted after it), optional split-K
// with a deterministic second-stage reduction.
#pragma once
#include "common.cuh"

namespace b2rl {

enum { MAP_STRIDE = 0, MAP_PIXEL = 1, MAP_KERNEL = 2 };

struct IndexMap {
    int kind = MAP_STRIDE;              // informational (the kernels take the kind as a template arg)
    int64_t stride = 1;                 // STRIDE : idx * stride
    int P = 1, OW = 1;                  // PIXEL  : b = idx / P, pix = idx % P, oy = pix / OW, ox = pix % OW
    int64_t bstride = 0;                //          -> b * bstride + oy * sy + ox * sx
    int sy = 0, sx = 0;
    int KK = 1, KS = 1, HW = 0, W = 0;  // KERNEL : ci = idx / KK, rem = idx % KK, ky = rem / KS, kx = rem % KS
                                        //          -> ci * HW + ky * W + kx
    const int64_t *gather = nullptr;    // optional row gather (replay-ring rows): b -> gather[b]
};

template <int KIND>
__device__ __forceinline__ int64_t map_off(const IndexMap &mp, int idx) {
    if constexpr (KIND == MAP_STRIDE) {
        const int64_t i = mp.gather ? mp.gather[idx] : (int64_t)idx;
        return i * mp.stride;
    } else if constexpr (KIND == MAP_PIXEL) {
        const int b = idx / mp.P, pix = idx - b * mp.P;
        const int oy = pix / mp.OW, ox = pix - oy * mp.OW;
        const int64_t bb = mp.gather ? mp.gather[b] : (int64_t)b;
        return bb * mp.bstride + (int64_t)(oy * mp.sy + ox * mp.sx);
    } else {
        const int ci = idx / mp.KK, rem = idx - ci * mp.KK;
        const int ky = rem / mp.KS, kx = rem - ky * mp.KS;
        return (int64_t)(ci * mp.HW + ky * mp.W + kx);
    }
}

static inline IndexMap map_stride(int64_t stride, const int64_t *gather = nullptr) {
    IndexMap m; m.kind = MAP_STRIDE; m.stride = stride; m.gather = gather; return m;
}
static inline IndexMap map_pixel(int P, int OW, int64_t bstride, int sy, int sx, const int64_t *gather = nullptr) {
    IndexMap m; m.kind = MAP_PIXEL; m.P = P; m.OW = OW; m.bstride = bstride; m.sy = sy; m.sx = sx; m.gather = gather;
    return m;
}
static inline IndexMap map_kernel(int ksize, int HW, int W) {
    IndexMap m; m.kind = MAP_KERNEL; m.KK = ksize * ksize; m.KS = ksize; m.HW = HW; m.W = W; return m;
}

// element kinds
enum { EL_F32 = 0, EL_U8 = 1, EL_F32_NORM = 2 };

struct Operand {
    const void *ptr = nullptr;
    int u8 = 0;                  // host-side description; the kernel gets the kind as a template arg
    int normalize = 0;           // (x - low) / (high - low), true fp32 division (quirk Q11)
    float low = 0.f, high = 1.f;
    IndexMap row, red;
    int ones_row = -1;           // row index whose elements read as 1.0 (bias-grad column of wgrad)
    int64_t base = 0;
    int elem_kind() const { return u8 ? EL_U8 : (normalize ? EL_F32_NORM : EL_F32); }
};

// compile-time description of one operand
template <int ELEM_, int ROWK_, int REDK_, bool RED_FAST_, bool ONES_>
struct OpTraits {
    static constexpr int ELEM = ELEM_, ROWK = ROWK_, REDK = REDK_;
    static constexpr bool RED_FAST = RED_FAST_, ONES = ONES_;
};

template <int ELEM>
__device__ __forceinline__ uint32_t load_raw(const Operand &op, int64_t off) {
    if constexpr (ELEM == EL_U8) return (uint32_t)__ldg(static_cast<const uint8_t *>(op.ptr) + op.base + off);
    else return __float_as_uint(__ldg(static_cast<const float *>(op.ptr) + op.base + off));
}
// lut: 256-entry table for uint8 inputs: (x - low) / (high - low) (exact fp32 division), or x.
template <int ELEM>
__device__ __forceinline__ float convert_raw(const Operand &op, uint32_t raw, const float *lut) {
    if constexpr (ELEM == EL_U8) return lut[raw];
    else if constexpr (ELEM == EL_F32_NORM) return __fdiv_rn(__uint_as_float(raw) - op.low, op.high - op.low);
    else return __uint_as_float(raw);
}

enum { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_WGRAD = 2, EPI_WGRAD_T = 3 };   // _T: rows are weight columns

struct Epilogue {
    int kind = EPI_STORE;
    float *out = nullptr;
    IndexMap om, on;             // out[om(m) + on(n)]
    const float *bias = nullptr; // per-n bias
    int act = B2RL_ACT_NONE;
    float *pre_out = nullptr;    // optional copy of the pre-activation value (GELU backward)
    int accumulate = 0;          // out += v instead of out = v
    float *db = nullptr;         // EPI_WGRAD: column n == wcols goes to db[m]
    int wcols = 0;
    // EPI_STORE of an input gradient: fold the activation backward of the layer BELOW into the store
    // (out = v * act'(mask_a[o]); mask_a = that layer's output, same layout as out) — one launch less
    const float *mask_a = nullptr;
    int mask_act = B2RL_ACT_NONE;
};

template <int KIND_, int OMK_, int ONK_>
struct EpiTraits {
    static constexpr int KIND = KIND_, OMK = OMK_, ONK = ONK_;
};

template <class TE>
__device__ __forceinline__ void epi_apply(const Epilogue &e, int m, int n, int64_t om_off, float v) {
    if constexpr (TE::KIND == EPI_WGRAD_T) {      // m = weight column (or the bias row), n = output channel
        if (m == e.wcols) {
            if (e.db) e.db[n] = e.accumulate ? e.db[n] + v : v;
        } else {
            float *p = e.out + (int64_t)n * e.wcols + m;
            *p = e.accumulate ? *p + v : v;
        }
    } else if constexpr (TE::KIND == EPI_WGRAD) {
        if (n == e.wcols) {
            if (e.db) e.db[m] = e.accumulate ? e.db[m] + v : v;
        } else {
            float *p = e.out + (int64_t)m * e.wcols + n;
            *p = e.accumulate ? *p + v : v;
        }
    } else {
        const int64_t o = om_off + map_off<TE::ONK>(e.on, n);
        if constexpr (TE::KIND == EPI_ATOMIC) {
            atomicAdd(e.out + o, v);
        } else {
            if (e.bias) v += e.bias[n];
            if (e.pre_out) e.pre_out[o] = v;
            v = act_fwd(e.act, v);
            if (e.mask_a) v *= act_bwd(e.mask_act, 0.f, e.mask_a[o]);
            e.out[o] = e.accumulate ? e.out[o] + v : v;
        }
    }
}

template <int BM, int BN, int BK, int TM, int TN, class TA, class TB, class TE>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
igemm_kernel(const Operand A, const Operand Bop, const Epilogue epi, int M, int N, int K, int k_chunk,
             float *__restrict__ partial) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int EA = BM * BK / NT, EB = BN * BK / NT;
    static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");
    static_assert(NT % BK == 0, "red-fast mapping needs a fixed k lane per thread");
    static_assert(EA <= 32 && EB <= 32, "validity masks are 32-bit");
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    __shared__ float lutA[TA::ELEM == EL_U8 ? 256 : 1], lutB[TB::ELEM == EL_U8 ? 256 : 1];
    const int tid = threadIdx.x;
    if constexpr (TA::ELEM == EL_U8)
        for (int i = tid; i < 256; i += NT) lutA[i] = A.normalize ? __fdiv_rn((float)i - A.low, A.high - A.low) : (float)i;
    if constexpr (TB::ELEM == EL_U8)
        for (int i = tid; i < 256; i += NT)
            lutB[i] = Bop.normalize ? __fdiv_rn((float)i - Bop.low, Bop.high - Bop.low) : (float)i;

    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kz = blockIdx.z;
    const int k_begin = kz * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    // per-thread element coordinates inside a tile (compile-time after unrolling)
    int64_t a_row[EA], b_row[EB];
    uint32_t a_okm = 0, b_okm = 0, a_onem = 0, b_onem = 0;
    auto a_kk = [&](int i) { const int e = tid + i * NT; return TA::RED_FAST ? e % BK : e / BM; };
    auto a_mm = [&](int i) { const int e = tid + i * NT; return TA::RED_FAST ? e / BK : e % BM; };
    auto b_kk = [&](int i) { const int e = tid + i * NT; return TB::RED_FAST ? e % BK : e / BN; };
    auto b_nn = [&](int i) { const int e = tid + i * NT; return TB::RED_FAST ? e / BK : e % BN; };
#pragma unroll
    for (int i = 0; i < EA; ++i) {
        const int m = m0 + a_mm(i);
        const bool ok = m < M;
        a_okm |= (uint32_t)ok << i;
        if constexpr (TA::ONES) a_onem |= (uint32_t)(m == A.ones_row) << i;
        a_row[i] = ok ? map_off<TA::ROWK>(A.row, m) : 0;
    }
#pragma unroll
    for (int i = 0; i < EB; ++i) {
        const int n = n0 + b_nn(i);
        const bool ok = n < N;
        b_okm |= (uint32_t)ok << i;
        if constexpr (TB::ONES) b_onem |= (uint32_t)(n == Bop.ones_row) << i;
        b_row[i] = ok ? map_off<TB::ROWK>(Bop.row, n) : 0;
    }

    uint32_t ra[EA], rb[EB];
    uint32_t a_valid = 0, b_valid = 0;      // bit i: element i is inside the problem (else 0)
    auto load_tile = [&](int k0) {
        if constexpr (TA::RED_FAST) {       // every element of this thread shares one k: decode it once
            const int k = k0 + a_kk(0);
            const bool kok = k < k_end;
            const int64_t ro = kok ? map_off<TA::REDK>(A.red, k) : 0;
            a_valid = kok ? a_okm : 0u;
#pragma unroll
            for (int i = 0; i < EA; ++i)
                ra[i] = ((a_valid & ~a_onem) >> i) & 1u ? load_raw<TA::ELEM>(A, a_row[i] + ro) : 0u;
        } else {
            a_valid = 0;
#pragma unroll
            for (int i = 0; i < EA; ++i) {
                const int k = k0 + a_kk(i);
                const bool in = ((a_okm >> i) & 1u) && k < k_end;
                a_valid |= (uint32_t)in << i;
                ra[i] = (in && !((a_onem >> i) & 1u)) ? load_raw<TA::ELEM>(A, a_row[i] + map_off<TA::REDK>(A.red, k)) : 0u;
            }
        }
        if constexpr (TB::RED_FAST) {
            const int k = k0 + b_kk(0);
            const bool kok = k < k_end;
            const int64_t ro = kok ? map_off<TB::REDK>(Bop.red, k) : 0;
            b_valid = kok ? b_okm : 0u;
#pragma unroll
            for (int i = 0; i < EB; ++i)
                rb[i] = ((b_valid & ~b_onem) >> i) & 1u ? load_raw<TB::ELEM>(Bop, b_row[i] + ro) : 0u;
        } else {
            b_valid = 0;
#pragma unroll
            for (int i = 0; i < EB; ++i) {
                const int k = k0 + b_kk(i);
                const bool in = ((b_okm >> i) & 1u) && k < k_end;
                b_valid |= (uint32_t)in << i;
                rb[i] = (in && !((b_onem >> i) & 1u)) ? load_raw<TB::ELEM>(Bop, b_row[i] + map_off<TB::REDK>(Bop.red, k))
                                                      : 0u;
            }
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < EA; ++i) {
            float v = convert_raw<TA::ELEM>(A, ra[i], lutA);
            if constexpr (TA::ONES) v = ((a_onem >> i) & 1u) ? 1.f : v;
            v = ((a_valid >> i) & 1u) ? v : 0.f;
            As[buf][a_kk(i)][a_mm(i)] = v;
        }
#pragma unroll
        for (int i = 0; i < EB; ++i) {
            float v = convert_raw<TB::ELEM>(Bop, rb[i], lutB);
            if constexpr (TB::ONES) v = ((b_onem >> i) & 1u) ? 1.f : v;
            v = ((b_valid >> i) & 1u) ? v : 0.f;
            Bs[buf][b_kk(i)][b_nn(i)] = v;
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    int buf = 0;
    __syncthreads();            // LUTs ready
    if (k_begin < k_end) {
        load_tile(k_begin);
        store_tile(0);
    }
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool has_next = k0 + BK < k_end;
        if (has_next) load_tile(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
            if constexpr (TM % 4 == 0) {
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * TM + i]);
                    a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][ty * TM + i];
            }
            if constexpr (TN % 4 == 0) {
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    const float4 t = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * TN + j]);
                    b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // epilogue: rolled over rows to keep the code small (runs once)
#pragma unroll 1
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= M) break;
        int64_t om_off = 0;
        if constexpr (TE::KIND != EPI_WGRAD && TE::KIND != EPI_WGRAD_T) om_off = partial ? 0 : map_off<TE::OMK>(epi.om, m);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= N) continue;
            float v = 0.f;
#pragma unroll
            for (int ii = 0; ii < TM; ++ii) v = (ii == i) ? acc[ii][j] : v;     // register select (no local mem)
            if (partial) partial[((int64_t)kz * M + m) * N + n] = v;
            else epi_apply<TE>(epi, m, n, om_off, v);
        }
    }
}

// Second stage of split-K: fixed-order sum over the kz partials, then the real epilogue.
// 32 outputs x 8 z-slices per CTA: slice j sums z = j, j+8, ... ; the 8 slice sums are then added in
// slice order by one thread per output -> deterministic, and 8x more loads in flight per output.
template <class TE>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float *__restrict__ partial, int splits, int M, int N,
                                                            const Epilogue epi) {
    __shared__ float sl[8][33];
    const int64_t total = (int64_t)M * N;
    const int ox = threadIdx.x & 31, zs = threadIdx.x >> 5;
    for (int64_t e0 = (int64_t)blockIdx.x * 32; e0 < total; e0 += (int64_t)gridDim.x * 32) {
        const int64_t e = e0 + ox;
        float v = 0.f;
        if (e < total)
            for (int z = zs; z < splits; z += 8) v += partial[(int64_t)z * total + e];
        sl[zs][ox] = v;
        __syncthreads();
        if (zs == 0 && e < total) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) t += sl[j][ox];
            const int m = (int)(e / N), n = (int)(e % N);
            int64_t om_off = 0;
            if constexpr (TE::KIND != EPI_WGRAD && TE::KIND != EPI_WGRAD_T) om_off = map_off<TE::OMK>(epi.om, m);
            epi_apply<TE>(epi, m, n, om_off, t);
        }
        __syncthreads();
    }
}

template <class TE>
static void launch_splitk_reduce(const float *partial, int splits, int M, int N, const Epilogue &epi, cudaStream_t s) {
    const int64_t total = (int64_t)M * N;
    int64_t blocks = (total + 31) / 32;
    if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
    splitk_reduce_kernel<TE><<<(int)blocks, 256, 0, s>>>(partial, splits, M, N, epi);
}

// 128-row tiles only pay off when there are many row tiles; the FFMA conv-wgrad fallback (M = taps)
// asks for them explicitly through TE (EPI_WGRAD_T)
static inline bool use_big_tile(int64_t M, bool wgrad_t = false) { return M >= 4096 || (wgrad_t && M >= 192); }

struct GemmPlan {
    int splits = 1;
    size_t partial_floats = 0;
};

// Decide split-K so that small-M/N problems still fill the machine.
static inline GemmPlan plan_gemm(int M, int N, int K, bool big_tile, int sms) {
    const int BM = big_tile ? 128 : 32, BN = 32, BK = 16;
    const int64_t ctas = (int64_t)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    GemmPlan p;
    if (ctas < sms && K >= 4 * BK) {
        int64_t s = (2 * (int64_t)sms + ctas - 1) / ctas;
        const int64_t max_s = K / (2 * BK);
        if (s > max_s) s = max_s;
        if (s > 64) s = 64;
        if (s > 1) {
            p.splits = (int)s;
            p.partial_floats = (size_t)s * M * N;
        }
    }
    return p;
}

// Launch wrapper.  `partial` must hold plan.partial_floats floats when plan.splits > 1.
template <class TA, class TB, class TE>
static int launch_igemm(const Operand &A, const Operand &B, const Epilogue &epi, int M, int N, int K,
                        float *partial, size_t partial_cap_floats, cudaStream_t s) {
    if (M <= 0 || N <= 0) return B2RL_OK;
    const bool big = use_big_tile(M, TE::KIND == EPI_WGRAD_T);
    GemmPlan p = plan_gemm(M, N, K, big, sm_count());
    if (p.splits > 1 && TE::KIND != EPI_ATOMIC && (partial == nullptr || p.partial_floats > partial_cap_floats))
        p.splits = 1;
    constexpr int BK = 16;
    int k_chunk = K;
    if (p.splits > 1) {
        k_chunk = ((K + p.splits - 1) / p.splits + BK - 1) / BK * BK;
        p.splits = (K + k_chunk - 1) / k_chunk;
    }
    float *part = (p.splits > 1 && TE::KIND != EPI_ATOMIC) ? partial : nullptr;   // atomics compose with split-K
    if (big) {
        dim3 grid((M + 127) / 128, (N + 31) / 32, p.splits);
        igemm_kernel<128, 32, BK, 8, 4, TA, TB, TE><<<grid, 128, 0, s>>>(A, B, epi, M, N, K, k_chunk, part);
    } else {
        dim3 grid((M + 31) / 32, (N + 31) / 32, p.splits);
        igemm_kernel<32, 32, BK, 2, 4, TA, TB, TE><<<grid, 128, 0, s>>>(A, B, epi, M, N, K, k_chunk, part);
    }
    B2RL_LAUNCH_CHECK();
    if (part) {
        launch_splitk_reduce<TE>(part, p.splits, M, N, epi, s);
        B2RL_LAUNCH_CHECK();
    }
    return B2RL_OK;
}

}  // namespace b2rl
