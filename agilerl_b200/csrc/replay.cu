// replay.cu — HBM-resident ring storage: wrap-split write, row gather, n-step fold.
//
// Replaces the TensorDict slice-assign / fancy-index of agilerl/components/replay_buffer.py
// (:100-107 write, :126/:204/:345 gather) and the Python loop of _get_n_step_info (:236-256).
// Pure byte movement: 128-bit coalesced loads/stores when rows are 16-byte multiples (uint8
// 4x84x84 frames are 28224 B = 1764 x 16 B), scalar path otherwise.
#include <math.h>

#include "common.cuh"

namespace b2rl {

template <typename V>
__global__ void ring_write_kernel(V *__restrict__ storage, const V *__restrict__ src, int64_t row_v,
                                  int64_t start, int64_t n, int64_t max_size) {
    const int64_t total = n * row_v;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / row_v, c = e - r * row_v;
        const int64_t dst_row = (start + r) % max_size;
        storage[dst_row * row_v + c] = src[e];
    }
}

// One CTA per row (grid-stride over rows) when rows are big; flat element loop when small.
template <typename V>
__global__ void gather_rows_kernel(V *__restrict__ dst, const V *__restrict__ storage,
                                   const int64_t *__restrict__ idx, int64_t row_v, int64_t n) {
    if (row_v >= 256) {
        for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
            const V *s = storage + idx[r] * row_v;
            V *d = dst + r * row_v;
            for (int64_t c = threadIdx.x; c < row_v; c += blockDim.x) d[c] = __ldg(s + c);
        }
    } else {
        const int64_t total = n * row_v;
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
             e += (int64_t)gridDim.x * blockDim.x) {
            const int64_t r = e / row_v, c = e - r * row_v;
            dst[e] = __ldg(storage + idx[r] * row_v + c);
        }
    }
}

constexpr int kMaxNStep = 16;
constexpr int kMaxFields = 8;
struct StepPtrs {
    const float *reward[kMaxNStep];
    const float *done[kMaxNStep];
    float scale[kMaxNStep];   // (float)(gamma ** k)
};

__global__ void nstep_fold_kernel(StepPtrs p, int n_step, int64_t num_envs, float *__restrict__ reward_out,
                                  int32_t *__restrict__ last_out) {
    // replay_buffer.py:236-256 — quirk Q3: reward of step k is added BEFORE its done is examined,
    // the loop stops after a step where ANY env is done, step 0's own done never stops it.
    for (int64_t e = threadIdx.x; e < num_envs; e += blockDim.x) reward_out[e] = p.reward[0][e];
    int last = 0;
    for (int k = 1; k < n_step; ++k) {
        int any = 0;
        for (int64_t e = threadIdx.x; e < num_envs; e += blockDim.x) {
            reward_out[e] = __fadd_rn(reward_out[e], __fmul_rn(p.reward[k][e], p.scale[k]));
            any |= (p.done[k][e] != 0.f);
        }
        last = k;
        if (__syncthreads_or(any)) break;
    }
    if (threadIdx.x == 0) *last_out = last;
}

struct SrcPtrs { const void *p[kMaxNStep]; };
__global__ void select_copy_kernel(uint8_t *__restrict__ dst, SrcPtrs srcs, const int32_t *__restrict__ which,
                                   int64_t bytes) {
    const uint8_t *s = static_cast<const uint8_t *>(srcs.p[*which]);
    const bool vec = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(dst) | (uintptr_t)bytes) & 15) == 0;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    if (vec) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (int64_t i = tid; i < bytes / 16; i += nt) d4[i] = s4[i];
    } else {
        for (int64_t i = tid; i < bytes; i += nt) dst[i] = s[i];
    }
}

// ---- n-step roll + ring write in ONE launch (MultiStepReplayBuffer.add, replay_buffer.py:173-194 + :206-258 + :72-112) ----
// The window's n per-env batches stay where ingest put them; every field of the n-step transition goes straight into
// its ring rows: obs / action (and any other key) from the oldest step, next_obs / done from the step the fold
// stopped at, reward folded like nstep_fold_kernel (fp32, gamma**k rounded to f32, quirk Q3).  blockIdx.y = field.
struct IngestField { uint8_t *ring; const uint8_t *src[kMaxNStep]; int64_t row_bytes; int role; };   // 0 first, 1 last, 2 reward
struct IngestArgs {
    IngestField f[kMaxFields];
    const float *reward[kMaxNStep];
    const float *done[kMaxNStep];
    float scale[kMaxNStep];
    int n_step;
    int64_t num_envs, cursor, max_size;
};
__global__ void nstep_ingest_kernel(IngestArgs a) {
    int last = 0;
    for (int k = 1; k < a.n_step; ++k) {           // every CTA derives the stopping step itself (a few hundred flags)
        int any = 0;
        for (int64_t e = threadIdx.x; e < a.num_envs; e += blockDim.x) any |= (a.done[k][e] != 0.f);
        last = k;
        if (__syncthreads_or(any)) break;
    }
    const IngestField f = a.f[blockIdx.y];
    if (f.role == 2) {
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < a.num_envs; e += (int64_t)gridDim.x * blockDim.x) {
            float r = a.reward[0][e];
            for (int k = 1; k <= last; ++k) r = __fadd_rn(r, __fmul_rn(a.reward[k][e], a.scale[k]));
            *reinterpret_cast<float *>(f.ring + ((a.cursor + e) % a.max_size) * f.row_bytes) = r;
        }
        return;
    }
    const uint8_t *src = f.src[f.role == 1 ? last : 0];
    const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(f.ring) | (uintptr_t)f.row_bytes) & 15) == 0;
    if (vec) {
        const int64_t rv = f.row_bytes / 16, total = a.num_envs * rv;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t e = i / rv, c = i - e * rv;
            reinterpret_cast<uint4 *>(f.ring + ((a.cursor + e) % a.max_size) * f.row_bytes)[c] =
                __ldg(reinterpret_cast<const uint4 *>(src + e * f.row_bytes) + c);
        }
    } else {
        const int64_t total = a.num_envs * f.row_bytes;
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int64_t e = i / f.row_bytes, c = i - e * f.row_bytes;
            f.ring[((a.cursor + e) % a.max_size) * f.row_bytes + c] = src[e * f.row_bytes + c];
        }
    }
}

// ---- all fields of a transition in one launch (blockIdx.y = field) -------------------------------
struct MultiField { void *dst; const void *src; int64_t row_bytes; int vec; };   // vec: 16 / 4 / 1 bytes per element
struct MultiFields { MultiField f[kMaxFields]; };

template <typename V>
__device__ __forceinline__ void gather_field(V *__restrict__ dst, const V *__restrict__ storage,
                                             const int64_t *__restrict__ idx, int64_t row_v, int64_t n) {
    if (row_v >= 256) {
        for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
            const V *s = storage + idx[r] * row_v;
            V *d = dst + r * row_v;
            for (int64_t c = threadIdx.x; c < row_v; c += blockDim.x) d[c] = __ldg(s + c);
        }
    } else {
        const int64_t total = n * row_v;
        for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
             e += (int64_t)gridDim.x * blockDim.x) {
            const int64_t r = e / row_v, c = e - r * row_v;
            dst[e] = __ldg(storage + idx[r] * row_v + c);
        }
    }
}
__global__ void gather_rows_multi_kernel(MultiFields mf, const int64_t *__restrict__ idx, int64_t n) {
    const MultiField f = mf.f[blockIdx.y];
    if (f.vec == 16) gather_field(static_cast<uint4 *>(f.dst), static_cast<const uint4 *>(f.src), idx, f.row_bytes / 16, n);
    else if (f.vec == 4) gather_field(static_cast<uint32_t *>(f.dst), static_cast<const uint32_t *>(f.src), idx, f.row_bytes / 4, n);
    else gather_field(static_cast<uint8_t *>(f.dst), static_cast<const uint8_t *>(f.src), idx, f.row_bytes, n);
}

template <typename V>
__device__ __forceinline__ void ring_write_field(V *__restrict__ storage, const V *__restrict__ src, int64_t row_v,
                                                 int64_t start, int64_t n, int64_t max_size) {
    const int64_t total = n * row_v;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / row_v, c = e - r * row_v;
        storage[((start + r) % max_size) * row_v + c] = src[e];
    }
}
__global__ void ring_write_multi_kernel(MultiFields mf, int64_t start, int64_t n, int64_t max_size) {
    const MultiField f = mf.f[blockIdx.y];     // dst = ring storage, src = the n new rows
    if (f.vec == 16) ring_write_field(static_cast<uint4 *>(f.dst), static_cast<const uint4 *>(f.src), f.row_bytes / 16, start, n, max_size);
    else if (f.vec == 4) ring_write_field(static_cast<uint32_t *>(f.dst), static_cast<const uint32_t *>(f.src), f.row_bytes / 4, start, n, max_size);
    else ring_write_field(static_cast<uint8_t *>(f.dst), static_cast<const uint8_t *>(f.src), f.row_bytes, start, n, max_size);
}

static int fill_fields(int n_fields, void *const *dst, const void *const *src, const int64_t *row_bytes, MultiFields &mf,
                       int64_t &max_row_bytes) {
    B2RL_CHECK_ARG(n_fields >= 1 && n_fields <= kMaxFields, "n_fields must be in [1, %d]", kMaxFields);
    B2RL_CHECK_ARG(dst && src && row_bytes, "NULL field table");
    max_row_bytes = 0;
    for (int i = 0; i < n_fields; ++i) {
        B2RL_CHECK_ARG(dst[i] && src[i] && row_bytes[i] > 0, "bad field %d", i);
        const uintptr_t al = reinterpret_cast<uintptr_t>(dst[i]) | reinterpret_cast<uintptr_t>(src[i]) | (uintptr_t)row_bytes[i];
        mf.f[i] = MultiField{dst[i], src[i], row_bytes[i], (al & 15) == 0 ? 16 : ((al & 3) == 0 ? 4 : 1)};
        if (row_bytes[i] > max_row_bytes) max_row_bytes = row_bytes[i];
    }
    return B2RL_OK;
}

template <typename F>
static int dispatch_width(const void *a, const void *b, int64_t row_bytes, F &&f) {
    const uintptr_t al = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | (uintptr_t)row_bytes;
    if ((al & 15) == 0) return f(uint4{}, row_bytes / 16);
    if ((al & 3) == 0) return f(uint32_t{}, row_bytes / 4);
    return f(uint8_t{}, row_bytes);
}

}  // namespace b2rl

using namespace b2rl;

namespace b2rl {
// B DISTINCT indices uniform over [0, N) — what randperm(N)[:B] yields as a set (ReplayBuffer.sample,
// replay_buffer.py:114-131, quirk Q12) — without generating the permutation: every slot draws from its own Philox
// counter stream and redraws while its value is already owned by another slot.  Ownership is decided by
// (round, slot) priority through atomicMin, so the result does not depend on thread timing.
__global__ void sample_distinct_kernel(uint64_t seed, uint64_t offset, int64_t N, int B, int tbits, int64_t *__restrict__ out) {
    extern __shared__ unsigned long long tab[];          // [T] keys (index + 1; 0 = empty)
    const int T = 1 << tbits;
    unsigned int *owner = reinterpret_cast<unsigned int *>(tab + T);   // [T] (round << 16 | slot) of the winning claim
    for (int i = threadIdx.x; i < T; i += blockDim.x) { tab[i] = 0ull; owner[i] = 0xFFFFFFFFu; }
    __syncthreads();
    const int i = threadIdx.x;
    bool done = i >= B;
    __shared__ int pending;
    for (unsigned int round = 0; round < 4096; ++round) {
        if (threadIdx.x == 0) pending = 0;
        __syncthreads();
        int slot = -1;
        long long idx = -1;
        if (!done) {
            uint32_t r[4];
            philox4x32_10(seed, offset + (uint64_t)i + (uint64_t)round * (uint64_t)B, 0x554E4946ull /* "UNIF" */, r);
            const unsigned long long u = ((unsigned long long)r[0] << 32) | r[1];
            idx = (long long)__umul64hi(u, (unsigned long long)N);           // floor(u * N / 2^64)
            const unsigned long long key = (unsigned long long)idx + 1ull;
            unsigned int h = (unsigned int)((key * 0x9E3779B97F4A7C15ull) >> (64 - tbits));
            while (true) {
                const unsigned long long prev = atomicCAS(&tab[h], 0ull, key);
                if (prev == 0ull || prev == key) break;
                h = (h + 1) & (T - 1);
            }
            slot = (int)h;
            atomicMin(&owner[slot], (round << 16) | (unsigned int)i);
        }
        __syncthreads();
        if (!done) {
            if (owner[slot] == ((round << 16) | (unsigned int)i)) { out[i] = idx; done = true; }
            else atomicAdd(&pending, 1);
        }
        __syncthreads();
        if (pending == 0) break;
        __syncthreads();
    }
}
// Dense case (B > N / 2, hence N < 2048): a full random permutation — every index gets a 64-bit Philox key, a bitonic
// sort orders (key, index) pairs in shared memory, the first B indices are the sample.
__global__ void sample_shuffle_kernel(uint64_t seed, uint64_t offset, int N, int B, int P, int64_t *__restrict__ out) {
    extern __shared__ unsigned long long keys[];         // [P] keys, then [P] payloads
    unsigned long long *val = keys + P;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        unsigned long long k = ~0ull;
        if (i < N) {
            uint32_t r[4];
            philox4x32_10(seed, offset + (uint64_t)i, 0x53484646ull /* "SHFF" */, r);
            k = (((unsigned long long)r[0] << 32) | r[1]) >> 1;         // < 2^63: padding keys sort last
        }
        keys[i] = k; val[i] = (unsigned long long)i;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < P; i += blockDim.x) {
                const int j = i ^ stride;
                if (j > i) {
                    const bool up = (i & size) == 0;
                    const unsigned long long ki = keys[i], kj = keys[j];
                    // ties broken by index so the order is total and deterministic
                    const bool gt = ki > kj || (ki == kj && val[i] > val[j]);
                    if (gt == up) { keys[i] = kj; keys[j] = ki; const unsigned long long t = val[i]; val[i] = val[j]; val[j] = t; }
                }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < B; i += blockDim.x) out[i] = (int64_t)val[i];
}

}  // namespace b2rl

extern "C" {

int b2rl_ring_write(void *storage, const void *src, int64_t row_bytes, int64_t start, int64_t n,
                    int64_t max_size, void *stream) {
    B2RL_CHECK_ARG(storage && src, "NULL buffer");
    B2RL_CHECK_ARG(row_bytes > 0 && max_size > 0, "row_bytes and max_size must be positive");
    B2RL_CHECK_ARG(start >= 0 && start < max_size, "cursor out of range");
    B2RL_CHECK_ARG(n >= 0 && n <= max_size, "cannot add more rows than max_size in one call");
    if (n == 0) return B2RL_OK;
    cudaStream_t s = as_stream(stream);
    return dispatch_width(storage, src, row_bytes, [&](auto tag, int64_t row_v) -> int {
        using V = decltype(tag);
        const int64_t total = n * row_v;
        int blocks = (int)((total + 255) / 256);
        const int cap_blocks = sm_count() * 8;
        if (blocks > cap_blocks) blocks = cap_blocks;
        ring_write_kernel<V><<<blocks, 256, 0, s>>>(static_cast<V *>(storage), static_cast<const V *>(src), row_v,
                                                    start, n, max_size);
        B2RL_LAUNCH_CHECK();
        return B2RL_OK;
    });
}

int b2rl_gather_rows(void *dst, const void *storage, const int64_t *idx, int64_t row_bytes, int64_t n,
                     void *stream) {
    B2RL_CHECK_ARG(dst && storage && idx, "NULL buffer");
    B2RL_CHECK_ARG(row_bytes > 0, "row_bytes must be positive");
    if (n <= 0) return B2RL_OK;
    cudaStream_t s = as_stream(stream);
    return dispatch_width(dst, storage, row_bytes, [&](auto tag, int64_t row_v) -> int {
        using V = decltype(tag);
        int blocks;
        if (row_v >= 256) blocks = (int)(n < (int64_t)sm_count() * 8 ? n : (int64_t)sm_count() * 8);
        else {
            const int64_t total = n * row_v;
            blocks = (int)((total + 255) / 256);
            if (blocks > sm_count() * 8) blocks = sm_count() * 8;
        }
        gather_rows_kernel<V><<<blocks, 256, 0, s>>>(static_cast<V *>(dst), static_cast<const V *>(storage), idx,
                                                     row_v, n);
        B2RL_LAUNCH_CHECK();
        return B2RL_OK;
    });
}

int b2rl_ring_write_multi(int n_fields, void *const *storage, const void *const *src, const int64_t *row_bytes,
                          int64_t start, int64_t n, int64_t max_size, void *stream) {
    B2RL_CHECK_ARG(max_size > 0 && start >= 0 && start < max_size, "cursor out of range");
    B2RL_CHECK_ARG(n >= 0 && n <= max_size, "cannot add more rows than max_size in one call");
    MultiFields mf;
    int64_t max_row = 0;
    int rc = fill_fields(n_fields, storage, src, row_bytes, mf, max_row);
    if (rc != B2RL_OK) return rc;
    if (n == 0) return B2RL_OK;
    int64_t blocks = (n * max_row / 16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    ring_write_multi_kernel<<<dim3((unsigned)blocks, n_fields), 256, 0, as_stream(stream)>>>(mf, start, n, max_size);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_gather_rows_multi(int n_fields, void *const *dst, const void *const *storage, const int64_t *row_bytes,
                           const int64_t *idx, int64_t n, void *stream) {
    B2RL_CHECK_ARG(idx, "NULL index buffer");
    MultiFields mf;
    int64_t max_row = 0;
    int rc = fill_fields(n_fields, dst, storage, row_bytes, mf, max_row);
    if (rc != B2RL_OK) return rc;
    if (n <= 0) return B2RL_OK;
    int64_t blocks = max_row >= 256 * 16 ? n : (n * max_row / 16 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
    gather_rows_multi_kernel<<<dim3((unsigned)blocks, n_fields), 256, 0, as_stream(stream)>>>(mf, idx, n);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_nstep_fold(const float *const *reward_steps, const float *const *done_steps, int n_step,
                    int64_t num_envs, double gamma, float *reward_out, int32_t *last_step_out, void *stream) {
    B2RL_CHECK_ARG(n_step >= 1 && n_step <= kMaxNStep, "n_step must be in [1, %d]", kMaxNStep);
    B2RL_CHECK_ARG(reward_steps && done_steps && reward_out && last_step_out, "NULL buffer");
    StepPtrs p;
    double g = 1.0;
    for (int k = 0; k < n_step; ++k) {
        p.reward[k] = reward_steps[k];
        p.done[k] = done_steps[k];
        // gamma ** k as CPython computes it (pow), then rounded to f32 by the tensor*scalar multiply
        p.scale[k] = (float)(k == 0 ? 1.0 : pow(gamma, (double)k));
        (void)g;
    }
    nstep_fold_kernel<<<1, 256, 0, as_stream(stream)>>>(p, n_step, num_envs, reward_out, last_step_out);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_select_copy(void *dst, const void *const *srcs_host, int n_srcs, const int32_t *which, int64_t bytes,
                     void *stream) {
    B2RL_CHECK_ARG(n_srcs >= 1 && n_srcs <= kMaxNStep, "n_srcs must be in [1, %d]", kMaxNStep);
    B2RL_CHECK_ARG(dst && srcs_host && which, "NULL buffer");
    if (bytes <= 0) return B2RL_OK;
    SrcPtrs sp;
    for (int i = 0; i < kMaxNStep; ++i) sp.p[i] = srcs_host[i < n_srcs ? i : 0];
    int blocks = (int)((bytes / 16 + 255) / 256);
    if (blocks < 1) blocks = 1;
    if (blocks > sm_count() * 4) blocks = sm_count() * 4;
    select_copy_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<uint8_t *>(dst), sp, which, bytes);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_nstep_ingest(int n_fields, void *const *ring, const void *const *src, const int64_t *row_bytes, const int32_t *role,
                      const float *const *reward_steps, const float *const *done_steps, int n_step, int64_t num_envs,
                      double gamma, int64_t cursor, int64_t max_size, void *stream) {
    B2RL_CHECK_ARG(n_fields >= 1 && n_fields <= b2rl::kMaxFields && n_step >= 1 && n_step <= b2rl::kMaxNStep, "bad field / step count");
    B2RL_CHECK_ARG(ring && src && row_bytes && role && reward_steps && done_steps && num_envs >= 1 && num_envs <= max_size,
                   "bad arguments");
    b2rl::IngestArgs a;
    int64_t max_bytes = 0;
    for (int i = 0; i < n_fields; ++i) {
        a.f[i].ring = static_cast<uint8_t *>(ring[i]);
        a.f[i].row_bytes = row_bytes[i];
        a.f[i].role = role[i];
        for (int k = 0; k < n_step; ++k) a.f[i].src[k] = static_cast<const uint8_t *>(src[i * n_step + k]);
        if (role[i] == 2) B2RL_CHECK_ARG(row_bytes[i] == 4, "the reward field must be one float32 per row");
        if (row_bytes[i] * num_envs > max_bytes) max_bytes = row_bytes[i] * num_envs;
    }
    double g = 1.0;
    for (int k = 0; k < n_step; ++k) {
        a.reward[k] = reward_steps[k]; a.done[k] = done_steps[k];
        a.scale[k] = (float)g;                      // gamma ** k as a Python double, rounded where torch rounds it
        g = pow(gamma, (double)(k + 1));
    }
    a.n_step = n_step; a.num_envs = num_envs; a.cursor = cursor; a.max_size = max_size;
    int bx = (int)((max_bytes / 16 + 255) / 256);
    bx = bx < 1 ? 1 : (bx > 296 ? 296 : bx);
    b2rl::nstep_ingest_kernel<<<dim3(bx, n_fields), 256, 0, b2rl::as_stream(stream)>>>(a);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_sample_uniform_distinct(uint64_t seed, uint64_t offset, int64_t N, int64_t B, int64_t *out_idx, void *stream) {
    B2RL_CHECK_ARG(out_idx && N >= 1 && B >= 1 && B <= N && B <= 1024, "need 1 <= B <= min(N, 1024)");
    if (2 * B > N) {      // dense draw (N <= 2047): rejection would crawl; sort all N indices by random keys instead
        int P = 1;
        while (P < N) P <<= 1;
        b2rl::sample_shuffle_kernel<<<1, 1024, (size_t)P * 2 * sizeof(unsigned long long), b2rl::as_stream(stream)>>>(
            seed, offset, (int)N, (int)B, P, out_idx);
        B2RL_LAUNCH_CHECK();
        return B2RL_OK;
    }
    int tbits = 4;
    while ((1 << tbits) < 2 * B) ++tbits;       // load factor <= 0.5; 24 KB of shared memory at B = 1024
    const size_t smem = (size_t)(1 << tbits) * (sizeof(unsigned long long) + sizeof(unsigned int));
    int threads = 32;
    while (threads < B) threads <<= 1;
    b2rl::sample_distinct_kernel<<<1, threads, smem, b2rl::as_stream(stream)>>>(seed, offset, N, (int)B, tbits, out_idx);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

/* HOST arithmetic of PrioritizedReplayBuffer.update_priorities (replay_buffer.py:411-428, :311-329): for every
 * float32 priority p: q = max((double)p, floor); out[i] = q ** alpha with glibc pow — the function CPython's float
 * `**` calls, so the leaves are bit-identical to the reference's — and the running maximum of q. */
int b2rl_host_priority_pow(const float *priority_host, int64_t n, double alpha, double floor_, double *out_host,
                           double *max_host) {
    B2RL_CHECK_ARG(n >= 0 && (n == 0 || (priority_host && out_host)), "bad arguments");
    double mx = max_host ? *max_host : 0.0;
    for (int64_t i = 0; i < n; ++i) {
        const double p = (double)priority_host[i];
        const double q = p > floor_ ? p : floor_;
        out_host[i] = pow(q, alpha);
        if (q > mx) mx = q;
    }
    if (max_host) *max_host = mx;
    return B2RL_OK;
}

}  // extern "C"
