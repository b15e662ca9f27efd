// tree.cu — fp64 sum/min priority trees resident in HBM (K2 / K1a of SURVEY §2).
//
// Replaces agilerl/components/segment_tree.py (list-backed Python trees) and the PER loops of
// agilerl/components/replay_buffer.py:306-309, 357-428.  Everything here is IEEE fp64 with
// explicit round-to-nearest intrinsics (no FMA contraction), so that given identical leaves the
// trees, the stratified upper bounds and the retrieved indices are bit-identical to the
// reference's CPython-double arithmetic.
#include <cstdarg>
#include <math.h>

#include "common.cuh"
#include <cstdlib>

namespace b2rl {

unsigned long long g_launches = 0;
unsigned long long g_conv_path[3] = {0, 0, 0};
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *last_error() { return g_err; }

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

__device__ __forceinline__ double dmin_py(double a, double b) { return b < a ? b : a; }  // Python min(a,b)

__global__ void tree_fill_kernel(double *st, double *mt, int64_t n2) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        st[i] = 0.0;
        mt[i] = inf;
    }
}

// One CTA walks the <= kMaxItems touched leaves up to the root, level by level.
// Sequential semantics of n __setitem__ calls: leaf <- last writer; every ancestor of a touched
// leaf is a pure function of its two children, so recomputing bottom-up after all leaves are
// placed yields the same bits as the reference's one-at-a-time updates.
constexpr int kTreeThreads = 1024;
constexpr int64_t kTreeChunk = 4096;

template <bool kRange, bool kFromPriority>
__global__ void __launch_bounds__(kTreeThreads)
tree_set_kernel(double *__restrict__ st, double *__restrict__ mt, int64_t cap,
                const int64_t *__restrict__ idx, const double *__restrict__ pa,
                const float *__restrict__ pri, int64_t n, int64_t tree_ptr, int64_t max_size,
                double range_val, double alpha, double floor_, double *max_priority) {
    const int tid = threadIdx.x;
    auto index_of = [&](int64_t i) -> int64_t {
        if (kRange) return (tree_ptr + i) % max_size;
        return idx[i];
    };
    double local_max = 0.0;
    for (int64_t i = tid; i < n; i += kTreeThreads) {
        const int64_t id = index_of(i);
        bool last = true;
        if (!kRange) {
            for (int64_t j = i + 1; j < n; ++j)
                if (idx[j] == id) { last = false; break; }
        } else {
            last = (i + max_size >= n);  // a later lap over the ring overwrites this slot
        }
        double v;
        if (kRange) {
            // max_priority given: range_val is the HOST maximum of raw priorities; the leaf is pow(max(host, device), alpha)
            v = max_priority ? pow(fmax(range_val, *max_priority), alpha) : range_val;
        } else if (kFromPriority) {
            double p = (double)pri[i];
            p = p < floor_ ? floor_ : p;          // max(priority, 1e-5), replay_buffer.py:425
            local_max = p > local_max ? p : local_max;
            v = pow(p, alpha);
        } else v = pa[i];
        if (last) {
            if (st) st[cap + id] = v;
            if (mt) mt[cap + id] = v;
        }
    }
    if (kFromPriority && max_priority != nullptr) {
        // block max -> *max_priority = max(*max_priority, max_i p_i)  (replay_buffer.py:329)
        __shared__ double smax[kTreeThreads / 32];
        for (int o = 16; o > 0; o >>= 1) {
            double other = __shfl_xor_sync(0xffffffffu, local_max, o);
            local_max = other > local_max ? other : local_max;
        }
        if ((tid & 31) == 0) smax[tid >> 5] = local_max;
        __syncthreads();
        if (tid == 0) {
            double m = *max_priority;
            for (int w = 0; w < kTreeThreads / 32; ++w) m = smax[w] > m ? smax[w] : m;
            *max_priority = m;
        }
    }
    __syncthreads();
    for (int shift = 1; (cap >> shift) >= 1; ++shift) {
        for (int64_t i = tid; i < n; i += kTreeThreads) {
            const int64_t node = (cap + index_of(i)) >> shift;
            if (st) st[node] = __dadd_rn(st[2 * node], st[2 * node + 1]);
            if (mt) mt[node] = dmin_py(mt[2 * node], mt[2 * node + 1]);
        }
        __syncthreads();
    }
}

// Fast path for a learn batch (n <= kFastMax leaves, both trees, cap <= 2^24): no global-memory round
// trip and no hashing inside the level loop.
//   1. last-writer-wins: a small hash set of leaf ids keeps, per id, the highest batch position.
//   2. every surviving thread registers the nodes of its leaf-to-root path (node ids are unique across
//      levels) in one shared-memory hash map node -> owning thread (first registrant wins; all threads
//      through a node hold identical values), then looks up, for every level, the thread that owns its
//      SIBLING (or none: then the sibling is untouched by this batch and its value was prefetched from
//      global memory up front, all levels at once).  Two barriers for all 17+ levels.
//   3. the level loop is then a thread-to-thread exchange through a double-buffered value array:
//      read the partner's value, add / min with the own one (registers), publish, one barrier.
// Same arithmetic as tree_set_kernel (left + right, Python min(left, right)), same last-writer rule.
constexpr int kFastMax = 512;
constexpr int kFastLevels = 24;
constexpr unsigned long long kFastEmpty64 = ~0ull;
constexpr int kFastDense = 2048;              // nodes below this id (the top 11 levels) use a direct-index owner table

__device__ __forceinline__ uint32_t fast_hash(uint32_t key, int bits) { return (key * 2654435761u) >> (32 - bits); }

template <bool kFromPriority>
__global__ void __launch_bounds__(kFastMax)
tree_set_fast_kernel(double *__restrict__ st, double *__restrict__ mt, int64_t cap, int levels,
                     const int64_t *__restrict__ idx, const double *__restrict__ pa, const float *__restrict__ pri,
                     int n, int tbits, double alpha, double floor_, double *max_priority) {
    extern __shared__ __align__(16) unsigned char tsm[];
    const int T = 1 << tbits;                     // node map slots
    // [T] u64 (node << 32 | owner) | [2][kFastMax] sum | [2][kFastMax] min | [kFastDense] int
    unsigned long long *nmap = reinterpret_cast<unsigned long long *>(tsm);
    double *vs = reinterpret_cast<double *>(tsm + (size_t)T * sizeof(unsigned long long));
    double *vm = vs + 2 * kFastMax;
    int *dense = reinterpret_cast<int *>(vm + 2 * kFastMax);      // [kFastDense] owner of the top-level nodes
    const int tid = threadIdx.x;
    for (int i = tid; i < T; i += blockDim.x) nmap[i] = kFastEmpty64;
    for (int i = tid; i < kFastDense; i += blockDim.x) dense[i] = -1;

    const bool have = tid < n;
    const uint32_t id = have ? (uint32_t)idx[tid] : 0u;
    double v = 0.0, local_max = 0.0;
    if (have) {
        if (kFromPriority) {
            double p = (double)pri[tid];
            p = p < floor_ ? floor_ : p;          // max(priority, 1e-5), replay_buffer.py:425
            local_max = p;
            v = pow(p, alpha);
        } else v = pa[tid];
    }
    if (kFromPriority && max_priority != nullptr) {
        // block max -> *max_priority = max(*max_priority, max_i p_i)  (replay_buffer.py:329)
        __shared__ double smax[kFastMax / 32];
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, local_max, o);
            local_max = other > local_max ? other : local_max;
        }
        if ((tid & 31) == 0) smax[tid >> 5] = local_max;
        __syncthreads();
        if (tid == 0) {
            double m = *max_priority;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = smax[w] > m ? smax[w] : m;
            *max_priority = m;
        }
    }
    __syncthreads();
    // ---- 1. last writer of every leaf: the map entry of the LEAF node keeps the highest position
    //         (key in the high word, so atomicMax on the packed word orders by position within a key)
    const uint32_t leaf = (uint32_t)cap + id;
    int lslot = 0;
    if (have) {
        lslot = (int)fast_hash(leaf, tbits);
        const unsigned long long mine = ((unsigned long long)leaf << 32) | (unsigned)tid;
        while (true) {
            const unsigned long long cur = atomicCAS(&nmap[lslot], kFastEmpty64, mine);
            if (cur == kFastEmpty64) break;
            if ((uint32_t)(cur >> 32) == leaf) { atomicMax(&nmap[lslot], mine); break; }
            lslot = (lslot + 1) & (T - 1);
        }
    }
    __syncthreads();
    const bool active = have && (uint32_t)(nmap[lslot] & 0xffffffffull) == (uint32_t)tid;
    // ---- prefetch the sibling of every node on this leaf's path (values before this batch)
    double sib_s[kFastLevels], sib_m[kFastLevels];
#pragma unroll
    for (int l = 0; l < kFastLevels; ++l) {
        sib_s[l] = 0.0; sib_m[l] = 0.0;
        if (active && l < levels) {
            const uint32_t sb = (leaf >> l) ^ 1u;
            sib_s[l] = st[sb];
            sib_m[l] = mt[sb];
        }
    }
    // ---- 2a. register the inner nodes of the path (the leaf entry already names this thread)
    if (active) {
#pragma unroll 4
        for (int l = 1; l < levels; ++l) {
            const uint32_t node = leaf >> l;
            if (node < (uint32_t)kFastDense) {           // any registrant may own it: plain store, no contention
                dense[node] = tid;
                continue;
            }
            const unsigned long long mine = ((unsigned long long)node << 32) | (unsigned)tid;
            int slot = (int)fast_hash(node, tbits);
            while (true) {
                const unsigned long long cur = atomicCAS(&nmap[slot], kFastEmpty64, mine);
                if (cur == kFastEmpty64 || (uint32_t)(cur >> 32) == node) break;
                slot = (slot + 1) & (T - 1);
            }
        }
    }
    __syncthreads();
    // ---- 2b. owner of the sibling at every level (-1: untouched by this batch)
    int partner[kFastLevels];
#pragma unroll
    for (int l = 0; l < kFastLevels; ++l) {
        partner[l] = -1;
        if (active && l < levels) {
            const uint32_t sb = (leaf >> l) ^ 1u;
            if (l > 0 && sb < (uint32_t)kFastDense) {
                partner[l] = dense[sb];
                continue;
            }
            int slot = (int)fast_hash(sb, tbits);
            while (true) {
                const unsigned long long cur = nmap[slot];
                if (cur == kFastEmpty64) break;
                if ((uint32_t)(cur >> 32) == sb) { partner[l] = (int)(cur & 0xffffffffull); break; }
                slot = (slot + 1) & (T - 1);
            }
        }
    }
    // ---- 3. level loop: thread-to-thread exchange
    double my_s = v, my_m = v;
    uint32_t node = leaf;
    if (active) {
        st[node] = v;
        mt[node] = v;
        vs[tid] = v;
        vm[tid] = v;
    }
    __syncthreads();
#pragma unroll
    for (int l = 0; l < kFastLevels; ++l) {
        if (l < levels) {
            const int rd = (l & 1) * kFastMax, wr = ((l + 1) & 1) * kFastMax;
            if (active) {
                double o_s = sib_s[l], o_m = sib_m[l];
                if (partner[l] >= 0) { o_s = vs[rd + partner[l]]; o_m = vm[rd + partner[l]]; }
                const bool left = (node & 1u) == 0;
                const double ns = left ? __dadd_rn(my_s, o_s) : __dadd_rn(o_s, my_s);
                const double nm = left ? dmin_py(my_m, o_m) : dmin_py(o_m, my_m);
                node >>= 1;
                st[node] = ns;
                mt[node] = nm;
                my_s = ns; my_m = nm;
                vs[wr + tid] = ns;
                vm[wr + tid] = nm;
            }
            __syncthreads();
        }
    }
}

static bool tree_fast_ok(const double *st, const double *mt, int64_t cap, int64_t n) {
    static const bool off = getenv("B2RL_TREE_SLOW") && getenv("B2RL_TREE_SLOW")[0] == '1';
    return !off && st && mt && n >= 1 && n <= kFastMax && cap >= 2 && cap <= ((int64_t)1 << kFastLevels);
}
template <bool kFromPriority>
static int launch_tree_fast(double *st, double *mt, int64_t cap, const int64_t *idx, const double *pa, const float *pri,
                            int64_t n, double alpha, double floor_, double *max_priority, cudaStream_t s) {
    int levels = 0;
    while (((int64_t)1 << levels) < cap) ++levels;
    // distinct path nodes <= sum over levels of min(n, nodes of the level): keep the map at most ~40 % full
    int64_t distinct = 0;
    for (int l = 0; l <= levels; ++l) distinct += n < ((int64_t)1 << l) ? n : ((int64_t)1 << l);
    int tbits = 8;
    while (((int64_t)1 << tbits) < distinct * 5 / 2) ++tbits;
    const int T = 1 << tbits;
    const size_t smem = (size_t)T * sizeof(unsigned long long) + (size_t)4 * kFastMax * sizeof(double) +
                        (size_t)kFastDense * sizeof(int);
    if (smem > 200 * 1024) return 1;
    const int threads = (int)((n + 31) / 32 * 32);
    auto kern = tree_set_fast_kernel<kFromPriority>;
    static bool big_smem = false;              // per instantiation; the launcher never asks for more than 200 KB
    if (!big_smem) {
        B2RL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        big_smem = true;
    }
    kern<<<1, threads, smem, s>>>(st, mt, cap, levels, idx, pa, pri, (int)n, tbits, alpha, floor_, max_priority);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

// Bulk rebuild of one level (used when a range update touches more than kTreeChunk leaves).
__global__ void tree_level_kernel(double *st, double *mt, int64_t first, int64_t count) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t node = first + i;
        st[node] = __dadd_rn(st[2 * node], st[2 * node + 1]);
        mt[node] = dmin_py(mt[2 * node], mt[2 * node + 1]);
    }
}
__global__ void tree_range_leaves_kernel(double *st, double *mt, int64_t cap, int64_t tree_ptr,
                                         int64_t n, int64_t max_size, double v, const double *max_dev = nullptr,
                                         double alpha = 0.0) {
    if (max_dev) v = pow(fmax(v, *max_dev), alpha);
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t id = (tree_ptr + i) % max_size;
        st[cap + id] = v;
        mt[cap + id] = v;
    }
}

__device__ __forceinline__ int64_t retrieve_dev(const double *__restrict__ st, int64_t cap, double ub) {
    int64_t idx = 1;
    while (idx < cap) {  // segment_tree.py:148-155 — strict '>' tie rule (quirk Q6)
        const int64_t left = 2 * idx;
        const double l = __ldg(st + left);
        if (l > ub) idx = left;
        else { ub = __dsub_rn(ub, l); idx = left + 1; }
    }
    return idx - cap;
}

__global__ void tree_retrieve_kernel(const double *st, int64_t cap, const double *ub, int64_t n, int64_t *out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = retrieve_dev(st, cap, ub[i]);
}

template <bool kPhilox>
__global__ void per_sample_kernel(const double *__restrict__ st, const double *__restrict__ mt, int64_t cap,
                                  const float *__restrict__ uniforms, uint64_t seed, uint64_t offset,
                                  int64_t B, double beta, int64_t size, int64_t *__restrict__ out_idx,
                                  float *__restrict__ out_w, const float *__restrict__ action_ring,
                                  const float *__restrict__ reward_ring, const float *__restrict__ done_ring,
                                  float *__restrict__ out_action, float *__restrict__ out_reward,
                                  float *__restrict__ out_done, const b2rl_step_state *__restrict__ state = nullptr) {
    if (state) { beta = state->beta; size = state->size; offset = state->sample_offset; }   // graph-replayed step
    // the top levels of the sum tree are walked by every sample: stage them once per CTA (one coalesced
    // load instead of ten dependent L2 round trips per thread)
    constexpr int kTop = 1024;
    __shared__ double top[kTop];
    const int ntop = (int)(2 * cap < kTop ? 2 * cap : kTop);
    for (int k = threadIdx.x; k < ntop; k += blockDim.x) top[k] = st[k];
    __syncthreads();
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= B) return;
    const double total = top[1];                                  // sum_tree.sum()
    const double segment = __ddiv_rn(total, (double)B);           // replay_buffer.py:369
    const double a = __dmul_rn(segment, (double)i);
    const double b = __dmul_rn(segment, (double)(i + 1));
    float u;
    if (kPhilox) {
        uint32_t r[4];
        philox4x32_10(seed, offset + (uint64_t)i, 0x50455253ull /* "PERS" */, r);
        u = (float)(r[0] >> 8) * (1.0f / 16777216.0f);            // [0,1) with 24 bits like torch.rand
    } else {
        u = uniforms[i];
    }
    const double ub = __dadd_rn(__dmul_rn((double)u, __dsub_rn(b, a)), a);   // :377
    int64_t idx;
    {
        double u2 = ub;
        int64_t node = 1;
        while (node < cap) {  // segment_tree.py:148-155 — strict '>' tie rule (quirk Q6); same arithmetic as retrieve_dev
            const int64_t left = 2 * node;
            const double l = left < ntop ? top[left] : __ldg(st + left);
            if (l > u2) node = left;
            else { u2 = __dsub_rn(u2, l); node = left + 1; }
        }
        idx = node - cap;
    }
    out_idx[i] = idx;
    if (out_w != nullptr) {                                        // :383-409
        const double p_min = __ddiv_rn(mt[1], total);
        const double max_w = pow(__dmul_rn(p_min, (double)size), -beta);
        const double p = __ddiv_rn(st[cap + idx], total);
        const double w = pow(__dmul_rn(p, (double)size), -beta);
        out_w[i] = (float)__ddiv_rn(w, max_w);
    }
    // fused gather of the small n-step fields of the sampled slot (frames are NOT copied: the conv
    // loader reads them straight from the ring through out_idx)
    if (out_action) out_action[i] = __ldg(action_ring + idx);
    if (out_reward) out_reward[i] = __ldg(reward_ring + idx);
    if (out_done) out_done[i] = __ldg(done_ring + idx);
}

static bool is_pow2(int64_t v) { return v > 0 && (v & (v - 1)) == 0; }

// what per_sample_kernel<true> / noise_reset_kernel<true> draw, written out (parity tests feed them to the oracle)
__global__ void philox_dump_kernel(uint64_t seed, uint64_t offset, int64_t n, int normal, float *__restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (normal) {
        out[i] = philox_normal(seed, offset + (uint64_t)i, 0x4E4F4953ull /* "NOIS" */);
    } else {
        uint32_t r[4];
        philox4x32_10(seed, offset + (uint64_t)i, 0x50455253ull /* "PERS" */, r);
        out[i] = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
    }
}

}  // namespace b2rl

using namespace b2rl;

extern "C" {

int b2rl_version(void) { return 100; }
unsigned long long b2rl_launch_count(void) { return b2rl::g_launches; }
unsigned long long b2rl_conv_path_count(int path) { return (path >= 0 && path < 3) ? b2rl::g_conv_path[path] : 0ull; }
const char *b2rl_last_error(void) { return b2rl::last_error(); }

int b2rl_device_sm_count(int device, int *out_host) {
    B2RL_CHECK_ARG(out_host != nullptr, "out_host is NULL");
    B2RL_CUDA(cudaDeviceGetAttribute(out_host, cudaDevAttrMultiProcessorCount, device));
    return B2RL_OK;
}

int b2rl_philox_uniforms(uint64_t seed, uint64_t offset, int64_t n, float *out, void *stream) {
    B2RL_CHECK_ARG(n >= 0 && (out || n == 0), "bad arguments");
    if (n == 0) return B2RL_OK;
    philox_dump_kernel<<<(int)((n + 255) / 256), 256, 0, as_stream(stream)>>>(seed, offset, n, 0, out);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}
int b2rl_philox_normals(uint64_t seed, uint64_t offset, int64_t n, float *out, void *stream) {
    B2RL_CHECK_ARG(n >= 0 && (out || n == 0), "bad arguments");
    if (n == 0) return B2RL_OK;
    philox_dump_kernel<<<(int)((n + 255) / 256), 256, 0, as_stream(stream)>>>(seed, offset, n, 1, out);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_tree_init(double *sum_tree, double *min_tree, int64_t cap, void *stream) {
    B2RL_CHECK_ARG(is_pow2(cap), "capacity must be positive and a power of 2.");
    B2RL_CHECK_ARG(sum_tree && min_tree, "tree pointer is NULL");
    const int64_t n2 = 2 * cap;
    const int blocks = (int)((n2 + 255) / 256 < 4096 ? (n2 + 255) / 256 : 4096);
    tree_fill_kernel<<<blocks, 256, 0, as_stream(stream)>>>(sum_tree, min_tree, n2);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_tree_set(double *sum_tree, double *min_tree, int64_t cap, const int64_t *idx,
                  const double *p_alpha, int64_t n, void *stream) {
    B2RL_CHECK_ARG(is_pow2(cap), "capacity must be positive and a power of 2.");
    B2RL_CHECK_ARG(n >= 0, "n must be >= 0");
    B2RL_CHECK_ARG(sum_tree || min_tree, "both trees are NULL");
    if (tree_fast_ok(sum_tree, min_tree, cap, n)) {
        const int rcf = launch_tree_fast<false>(sum_tree, min_tree, cap, idx, p_alpha, nullptr, n, 0.0, 0.0, nullptr,
                                                as_stream(stream));
        if (rcf != 1) return rcf;
    }
    for (int64_t off = 0; off < n; off += kTreeChunk) {   // chunks keep last-writer-wins order
        const int64_t m = n - off < kTreeChunk ? n - off : kTreeChunk;
        tree_set_kernel<false, false><<<1, kTreeThreads, 0, as_stream(stream)>>>(
            sum_tree, min_tree, cap, idx + off, p_alpha + off, nullptr, m, 0, 0, 0.0, 0.0, 0.0, nullptr);
        B2RL_LAUNCH_CHECK();
    }
    return B2RL_OK;
}

int b2rl_tree_set_from_priorities(double *sum_tree, double *min_tree, int64_t cap, const int64_t *idx,
                                  const float *priority, int64_t n, double alpha, double floor_,
                                  double *max_priority, void *stream) {
    B2RL_CHECK_ARG(is_pow2(cap), "capacity must be positive and a power of 2.");
    if (tree_fast_ok(sum_tree, min_tree, cap, n)) {
        const int rcf = launch_tree_fast<true>(sum_tree, min_tree, cap, idx, nullptr, priority, n, alpha, floor_,
                                               max_priority, as_stream(stream));
        if (rcf != 1) return rcf;
    }
    for (int64_t off = 0; off < n; off += kTreeChunk) {
        const int64_t m = n - off < kTreeChunk ? n - off : kTreeChunk;
        tree_set_kernel<false, true><<<1, kTreeThreads, 0, as_stream(stream)>>>(
            sum_tree, min_tree, cap, idx + off, nullptr, priority + off, m, 0, 0, 0.0, alpha, floor_, max_priority);
        B2RL_LAUNCH_CHECK();
    }
    return B2RL_OK;
}

static int tree_set_range_impl(double *sum_tree, double *min_tree, int64_t cap, int64_t tree_ptr, int64_t n,
                               int64_t max_size, double value, const double *max_dev, double alpha, void *stream) {
    B2RL_CHECK_ARG(is_pow2(cap), "capacity must be positive and a power of 2.");
    B2RL_CHECK_ARG(max_size > 0 && max_size <= cap, "max_size must be in (0, cap]");
    B2RL_CHECK_ARG(tree_ptr >= 0 && tree_ptr < max_size, "tree_ptr out of range");
    if (n <= 0) return B2RL_OK;
    cudaStream_t s = as_stream(stream);
    if (n <= kTreeChunk) {
        tree_set_kernel<true, false><<<1, kTreeThreads, 0, s>>>(sum_tree, min_tree, cap, nullptr, nullptr, nullptr,
                                                                 n, tree_ptr, max_size, value, alpha, 0.0,
                                                                 const_cast<double *>(max_dev));
        B2RL_LAUNCH_CHECK();
        return B2RL_OK;
    }
    const int64_t m = n < max_size ? n : max_size;   // more than a lap rewrites every slot
    tree_range_leaves_kernel<<<(int)((m + 255) / 256 < 2048 ? (m + 255) / 256 : 2048), 256, 0, s>>>(
        sum_tree, min_tree, cap, tree_ptr, m, max_size, value, max_dev, alpha);
    B2RL_LAUNCH_CHECK();
    for (int64_t count = cap / 2; count >= 1; count /= 2) {
        const int blocks = (int)((count + 255) / 256 < 2048 ? (count + 255) / 256 : 2048);
        tree_level_kernel<<<blocks, 256, 0, s>>>(sum_tree, min_tree, count, count);
        B2RL_LAUNCH_CHECK();
    }
    return B2RL_OK;
}

int b2rl_tree_set_range(double *sum_tree, double *min_tree, int64_t cap, int64_t tree_ptr, int64_t n,
                        int64_t max_size, double p_alpha, void *stream) {
    return tree_set_range_impl(sum_tree, min_tree, cap, tree_ptr, n, max_size, p_alpha, nullptr, 0.0, stream);
}

int b2rl_tree_set_range_devmax(double *sum_tree, double *min_tree, int64_t cap, int64_t tree_ptr, int64_t n,
                               int64_t max_size, double host_max, const double *max_priority_dev, double alpha,
                               void *stream) {
    B2RL_CHECK_ARG(max_priority_dev != nullptr, "max_priority_dev is NULL");
    return tree_set_range_impl(sum_tree, min_tree, cap, tree_ptr, n, max_size, host_max, max_priority_dev, alpha, stream);
}

int b2rl_tree_retrieve(const double *sum_tree, int64_t cap, const double *upperbound, int64_t n,
                       int64_t *out_idx, void *stream) {
    B2RL_CHECK_ARG(is_pow2(cap), "capacity must be positive and a power of 2.");
    if (n <= 0) return B2RL_OK;
    tree_retrieve_kernel<<<(int)((n + 127) / 128), 128, 0, as_stream(stream)>>>(sum_tree, cap, upperbound, n, out_idx);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_per_sample(const double *sum_tree, const double *min_tree, int64_t cap, const float *uniforms,
                    int64_t B, double beta, int64_t size, int64_t *out_idx, float *out_w, void *stream) {
    B2RL_CHECK_ARG(is_pow2(cap), "capacity must be positive and a power of 2.");
    B2RL_CHECK_ARG(B > 0 && uniforms && out_idx, "bad sample arguments");
    per_sample_kernel<false><<<(int)((B + 63) / 64), 64, 0, as_stream(stream)>>>(
        sum_tree, min_tree, cap, uniforms, 0, 0, B, beta, size, out_idx, out_w, nullptr, nullptr, nullptr, nullptr, nullptr,
        nullptr);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_per_sample_philox(const double *sum_tree, const double *min_tree, int64_t cap, uint64_t seed,
                           uint64_t offset, int64_t B, double beta, int64_t size, int64_t *out_idx,
                           float *out_w, void *stream) {
    B2RL_CHECK_ARG(is_pow2(cap), "capacity must be positive and a power of 2.");
    B2RL_CHECK_ARG(B > 0 && out_idx, "bad sample arguments");
    per_sample_kernel<true><<<(int)((B + 63) / 64), 64, 0, as_stream(stream)>>>(
        sum_tree, min_tree, cap, nullptr, seed, offset, B, beta, size, out_idx, out_w, nullptr, nullptr, nullptr, nullptr,
        nullptr, nullptr);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}


int b2rl_per_sample_fused(const double *sum_tree, const double *min_tree, int64_t cap, const float *uniforms,
                          uint64_t seed, uint64_t offset, int64_t B, double beta, int64_t size,
                          const float *action_ring, const float *reward_ring, const float *done_ring,
                          int64_t *out_idx, float *out_w, float *out_action, float *out_reward, float *out_done,
                          void *stream) {
    B2RL_CHECK_ARG(is_pow2(cap), "capacity must be positive and a power of 2.");
    B2RL_CHECK_ARG(B > 0 && out_idx, "bad sample arguments");
    B2RL_CHECK_ARG(action_ring && reward_ring && done_ring && out_action && out_reward && out_done, "NULL ring field");
    if (uniforms)
        per_sample_kernel<false><<<(int)((B + 63) / 64), 64, 0, as_stream(stream)>>>(
            sum_tree, min_tree, cap, uniforms, 0, 0, B, beta, size, out_idx, out_w, action_ring, reward_ring, done_ring,
            out_action, out_reward, out_done);
    else
        per_sample_kernel<true><<<(int)((B + 63) / 64), 64, 0, as_stream(stream)>>>(
            sum_tree, min_tree, cap, nullptr, seed, offset, B, beta, size, out_idx, out_w, action_ring, reward_ring,
            done_ring, out_action, out_reward, out_done);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

int b2rl_per_sample_fused_state(const double *sum_tree, const double *min_tree, int64_t cap, uint64_t seed,
                                const b2rl_step_state *state, int64_t B, const float *action_ring,
                                const float *reward_ring, const float *done_ring, int64_t *out_idx, float *out_w,
                                float *out_action, float *out_reward, float *out_done, void *stream) {
    B2RL_CHECK_ARG(is_pow2(cap), "capacity must be positive and a power of 2.");
    B2RL_CHECK_ARG(B > 0 && out_idx && state, "bad sample arguments");
    B2RL_CHECK_ARG(action_ring && reward_ring && done_ring && out_action && out_reward && out_done, "NULL ring field");
    per_sample_kernel<true><<<(int)((B + 63) / 64), 64, 0, as_stream(stream)>>>(
        sum_tree, min_tree, cap, nullptr, seed, 0, B, 0.0, 0, out_idx, out_w, action_ring, reward_ring, done_ring,
        out_action, out_reward, out_done, state);
    B2RL_LAUNCH_CHECK();
    return B2RL_OK;
}

}  // extern "C"
