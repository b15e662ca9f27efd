// gemm.cuh — one fp32 "separable-index" GEMM engine for every dense contraction on the path.
//
//   C[outM(m) + outN(n)] = epi( sum_r  A[aRow(m) + aRed(r)] * B[bRow(n) + bRed(r)] )
//
// Every operand address is a SUM of a row term and a reduction term, each produced by a tiny
// index map (dense stride, im2col output-pixel, im2col kernel-tap).  That single form covers
//   conv forward   (A = im2col(frames), B = W)          conv wgrad (+bias grad as a ones column)
//   conv dgrad     (col2im scatter epilogue)            linear forward / dX / dW
// so there is exactly one inner loop to optimise.  fp32 FFMA on CUDA cores with fp32 accumulate:
// the reference's losses are pinned at 1e-5 against an fp32 CPU run, which single-pass TF32/BF16
// tensor-core MMA does not hold; the tcgen05 3xTF32 variant is the planned upgrade for
// the two conv contractions and slots in behind the same launch wrapper.
//
// Tiling: BM x BN x BK smem tiles, TM x TN register micro-tiles, register-prefetch double
// buffering, optional split-K with a deterministic second-stage reduction.
#pragma once
#include "common.cuh"

namespace b2rl {

enum { MAP_STRIDE = 0, MAP_PIXEL = 1, MAP_KERNEL = 2 };

struct IndexMap {
    int kind = MAP_STRIDE;
    int64_t stride = 1;                 // STRIDE : idx * stride
    int P = 1, OW = 1;                  // PIXEL  : b = idx / P, pix = idx % P, oy = pix / OW, ox = pix % OW
    int64_t bstride = 0;                //          -> b * bstride + oy * sy + ox * sx
    int sy = 0, sx = 0;
    int KK = 1, KS = 1, HW = 0, W = 0;  // KERNEL : ci = idx / KK, rem = idx % KK, ky = rem / KS, kx = rem % KS
                                        //          -> ci * HW + ky * W + kx
    const int64_t *gather = nullptr;    // optional row gather (replay-ring rows): b -> gather[b]

    __device__ __forceinline__ int64_t off(int idx) const {
        if (kind == MAP_STRIDE) {
            const int64_t i = gather ? gather[idx] : (int64_t)idx;
            return i * stride;
        }
        if (kind == MAP_PIXEL) {
            const int b = idx / P, pix = idx - b * P;
            const int oy = pix / OW, ox = pix - oy * OW;
            const int64_t bb = gather ? gather[b] : (int64_t)b;
            return bb * bstride + (int64_t)oy * sy + (int64_t)ox * sx;
        }
        const int ci = idx / KK, rem = idx - ci * KK;
        const int ky = rem / KS, kx = rem - ky * KS;
        return (int64_t)ci * HW + (int64_t)ky * W + kx;
    }
};

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

struct Operand {
    const void *ptr = nullptr;
    int u8 = 0;                  // elements are uint8 observations
    int normalize = 0;           // (x - low) / (high - low), true fp32 division (quirk Q11)
    float low = 0.f, high = 1.f;
    IndexMap row, red;
    int ones_row = -1;           // row index whose elements read as 1.0 (bias-grad column of wgrad)
    int64_t base = 0;            // element offset added to every address

    __device__ __forceinline__ float fetch(int64_t off) const {
        if (u8) {
            const float v = (float)__ldg(static_cast<const uint8_t *>(ptr) + base + off);
            return normalize ? __fdiv_rn(v - low, high - low) : v;
        }
        const float v = __ldg(static_cast<const float *>(ptr) + base + off);
        return normalize ? __fdiv_rn(v - low, high - low) : v;
    }
};

enum { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_WGRAD = 2 };

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

    __device__ __forceinline__ void apply(int m, int n, float v) const {
        if (kind == EPI_WGRAD) {
            if (n == wcols) {
                if (db) db[m] = accumulate ? db[m] + v : v;
            } else {
                float *p = out + (int64_t)m * wcols + n;
                *p = accumulate ? *p + v : v;
            }
            return;
        }
        const int64_t o = om.off(m) + on.off(n);
        if (kind == EPI_ATOMIC) {
            atomicAdd(out + o, v);
            return;
        }
        if (bias) v += bias[n];
        if (pre_out) pre_out[o] = v;
        v = act_fwd(act, v);
        out[o] = accumulate ? out[o] + v : v;
    }
};

template <int BM, int BN, int BK, int TM, int TN, bool A_RED_FAST, bool B_RED_FAST>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
igemm_kernel(Operand A, Operand Bop, Epilogue epi, int M, int N, int K, int k_chunk, float *__restrict__ partial) {
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int EA = BM * BK / NT, EB = BN * BK / NT;
    static_assert(BM * BK % NT == 0 && BN * BK % NT == 0, "tile/threads mismatch");
    static_assert(TM % 4 == 0 || TM == 2 || TM == 1, "TM");
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int kz = blockIdx.z;
    const int k_begin = kz * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    // per-thread element coordinates inside a tile
    int a_mm[EA], a_kk[EA], b_nn[EB], b_kk[EB];
    int64_t a_row[EA], b_row[EB];
    bool a_ok[EA], b_ok[EB], a_one[EA], b_one[EB];
#pragma unroll
    for (int i = 0; i < EA; ++i) {
        const int e = tid + i * NT;
        if (A_RED_FAST) { a_kk[i] = e % BK; a_mm[i] = e / BK; }
        else            { a_mm[i] = e % BM; a_kk[i] = e / BM; }
        const int m = m0 + a_mm[i];
        a_ok[i] = m < M;
        a_one[i] = (m == A.ones_row);
        a_row[i] = a_ok[i] ? A.row.off(m) : 0;
    }
#pragma unroll
    for (int i = 0; i < EB; ++i) {
        const int e = tid + i * NT;
        if (B_RED_FAST) { b_kk[i] = e % BK; b_nn[i] = e / BK; }
        else            { b_nn[i] = e % BN; b_kk[i] = e / BN; }
        const int n = n0 + b_nn[i];
        b_ok[i] = n < N;
        b_one[i] = (n == Bop.ones_row);
        b_row[i] = b_ok[i] ? Bop.row.off(n) : 0;
    }

    float ra[EA], rb[EB];
    static_assert(NT % BK == 0, "red-fast mapping needs a fixed k lane per thread");
    auto load_tile = [&](int k0) {
        if (A_RED_FAST) {   // every element of this thread shares one k: decode it once
            const int k = k0 + a_kk[0];
            const bool kok = k < k_end;
            const int64_t ro = kok ? A.red.off(k) : 0;
#pragma unroll
            for (int i = 0; i < EA; ++i)
                ra[i] = (a_ok[i] && kok) ? (a_one[i] ? 1.f : A.fetch(a_row[i] + ro)) : 0.f;
        } else {
#pragma unroll
            for (int i = 0; i < EA; ++i) {
                const int k = k0 + a_kk[i];
                float v = 0.f;
                if (a_ok[i] && k < k_end) v = a_one[i] ? 1.f : A.fetch(a_row[i] + A.red.off(k));
                ra[i] = v;
            }
        }
        if (B_RED_FAST) {
            const int k = k0 + b_kk[0];
            const bool kok = k < k_end;
            const int64_t ro = kok ? Bop.red.off(k) : 0;
#pragma unroll
            for (int i = 0; i < EB; ++i)
                rb[i] = (b_ok[i] && kok) ? (b_one[i] ? 1.f : Bop.fetch(b_row[i] + ro)) : 0.f;
        } else {
#pragma unroll
            for (int i = 0; i < EB; ++i) {
                const int k = k0 + b_kk[i];
                float v = 0.f;
                if (b_ok[i] && k < k_end) v = b_one[i] ? 1.f : Bop.fetch(b_row[i] + Bop.red.off(k));
                rb[i] = v;
            }
        }
    };
    auto store_tile = [&](int buf) {
#pragma unroll
        for (int i = 0; i < EA; ++i) As[buf][a_kk[i]][a_mm[i]] = ra[i];
#pragma unroll
        for (int i = 0; i < EB; ++i) Bs[buf][b_kk[i]][b_nn[i]] = rb[i];
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    int buf = 0;
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
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[buf][kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= N) continue;
            if (partial) partial[((int64_t)kz * M + m) * N + n] = acc[i][j];
            else epi.apply(m, n, acc[i][j]);
        }
    }
}

// Second stage of split-K: fixed-order sum over the kz partials, then the real epilogue.
__global__ void splitk_reduce_kernel(const float *__restrict__ partial, int splits, int M, int N, Epilogue epi) {
    const int64_t total = (int64_t)M * N;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        float v = 0.f;
        for (int z = 0; z < splits; ++z) v += partial[(int64_t)z * total + e];
        epi.apply((int)(e / N), (int)(e % N), v);
    }
}

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
template <bool A_RED_FAST, bool B_RED_FAST>
static int launch_igemm(const Operand &A, const Operand &B, const Epilogue &epi, int M, int N, int K,
                        float *partial, size_t partial_cap_floats, cudaStream_t s) {
    if (M <= 0 || N <= 0) return B2RL_OK;
    const bool big = (int64_t)M >= 4096;
    GemmPlan p = plan_gemm(M, N, K, big, sm_count());
    if (p.splits > 1 && (partial == nullptr || p.partial_floats > partial_cap_floats)) { p.splits = 1; }
    if (epi.kind == EPI_ATOMIC) p.splits = p.splits;  // atomics compose with split-K directly
    constexpr int BK = 16;
    int k_chunk = K;
    if (p.splits > 1) {
        k_chunk = ((K + p.splits - 1) / p.splits + BK - 1) / BK * BK;
        p.splits = (K + k_chunk - 1) / k_chunk;
    }
    float *part = (p.splits > 1 && epi.kind != EPI_ATOMIC) ? partial : nullptr;
    if (big) {
        dim3 grid((M + 127) / 128, (N + 31) / 32, p.splits);
        igemm_kernel<128, 32, BK, 8, 4, A_RED_FAST, B_RED_FAST><<<grid, 128, 0, s>>>(A, B, epi, M, N, K, k_chunk, part);
    } else {
        dim3 grid((M + 31) / 32, (N + 31) / 32, p.splits);
        igemm_kernel<32, 32, BK, 2, 4, A_RED_FAST, B_RED_FAST><<<grid, 128, 0, s>>>(A, B, epi, M, N, K, k_chunk, part);
    }
    B2RL_LAUNCH_CHECK();
    if (part) {
        const int64_t total = (int64_t)M * N;
        int blocks = (int)((total + 255) / 256);
        if (blocks > sm_count() * 8) blocks = sm_count() * 8;
        splitk_reduce_kernel<<<blocks, 256, 0, s>>>(part, p.splits, M, N, epi);
        B2RL_LAUNCH_CHECK();
    }
    return B2RL_OK;
}

}  // namespace b2rl
