// host_rng.cu — HOST helper (no device work): the first B entries of torch.randperm(n) drawn from torch's CPU generator.
//
// ReplayBuffer.sample (agilerl/components/replay_buffer.py:126) draws `torch.randperm(self.size)[:batch_size]`: a full
// Fisher-Yates shuffle of the whole buffer (8 MB of indices for BASELINE configs[2]'s 1 M transitions, ~8 ms on the host)
// of which the first B = 512 entries are used.  torch's CPU shuffle for n < UINT32_MAX / 20 is
//     r[i] = i;  for i in 0 .. n-2:  z = generator->random() % (n - i);  swap(r[i], r[z + i]);
// with generator->random() the next 32-bit output of the generator's mt19937.  Entry i is final after iteration i, so the
// prefix needs only the first B draws and the (at most 2B) positions they touch; the other n-1-B draws only advance the
// generator.  This function does exactly that on the generator's serialised state (torch.get_rng_state(): seed u64,
// left i32, seeded i32, next u64, state u64[624], then the normal-sample caches, which random() does not touch): same
// indices, same generator state afterwards, without the 8 MB permutation.  The Python side checks this function against
// torch.randperm once per process (indices AND the generator state) and keeps torch.randperm if they ever differ.
#include <vector>

#include "common.cuh"

namespace {

struct Mt {                 // at::mt19937 over the serialised state
    int32_t *left;          // outputs remaining before the next twist, plus one (at::mt19937::operator(): --left == 0 -> twist)
    uint64_t *next;         // index of the next output in state[]
    uint64_t *state;        // 624 words, 32 significant bits each

    inline void twist() {
        const uint32_t UP = 0x80000000u, LOW = 0x7fffffffu, MAG = 0x9908b0dfu;
        auto mix = [&](uint64_t u, uint64_t v) -> uint32_t {
            const uint32_t y = ((uint32_t)u & UP) | ((uint32_t)v & LOW);
            return (y >> 1) ^ ((v & 1u) ? MAG : 0u);
        };
        uint64_t *p = state;
        for (int j = 0; j < 624 - 397; ++j) p[j] = p[j + 397] ^ mix(p[j], p[j + 1]);
        for (int j = 624 - 397; j < 623; ++j) p[j] = p[j + 397 - 624] ^ mix(p[j], p[j + 1]);
        p[623] = p[396] ^ mix(p[623], p[0]);
        *left = 624;
        *next = 0;
    }
    inline uint32_t draw() {
        if (--(*left) == 0) twist();
        uint32_t y = (uint32_t)state[(*next)++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    inline void skip(int64_t k) {          // k outputs drawn and discarded
        while (k > 0) {
            if (*left <= 1) {              // the next draw twists first
                --(*left);
                twist();
                ++(*next);
                --k;
                continue;
            }
            const int64_t take = k < (int64_t)(*left - 1) ? k : (int64_t)(*left - 1);
            *left -= (int32_t)take;
            *next += (uint64_t)take;
            k -= take;
        }
    }
};

struct SparsePerm {         // r[] of the shuffle, identity except where a swap wrote (open addressing, power-of-two table)
    std::vector<int64_t> key, val;
    size_t mask;
    explicit SparsePerm(int64_t touched) {
        size_t cap = 16;
        while ((int64_t)cap < 4 * touched) cap <<= 1;
        key.assign(cap, -1);
        val.assign(cap, 0);
        mask = cap - 1;
    }
    inline size_t slot(int64_t k) const {
        size_t h = ((uint64_t)k * 0x9E3779B97F4A7C15ull) >> 17;
        for (h &= mask; key[h] != -1 && key[h] != k; h = (h + 1) & mask) {}
        return h;
    }
    inline int64_t get(int64_t k) const {
        const size_t h = slot(k);
        return key[h] == k ? val[h] : k;
    }
    inline void set(int64_t k, int64_t v) {
        const size_t h = slot(k);
        key[h] = k;
        val[h] = v;
    }
};

}  // namespace

extern "C" int b2rl_host_randperm_prefix(uint8_t *rng_state_host, int64_t state_bytes, int64_t n, int64_t B, int64_t *out_host) {
    B2RL_CHECK_ARG(rng_state_host && state_bytes >= 24 + 624 * 8, "generator state too short (%lld bytes)", (long long)state_bytes);
    B2RL_CHECK_ARG(n >= 0 && B >= 0 && B <= n && (B == 0 || out_host), "need 0 <= B <= n");
    B2RL_CHECK_ARG(n < (int64_t)(0xffffffffu / 20), "n too large for torch's 32-bit shuffle");
    Mt g{reinterpret_cast<int32_t *>(rng_state_host + 8), reinterpret_cast<uint64_t *>(rng_state_host + 16),
         reinterpret_cast<uint64_t *>(rng_state_host + 24)};
    B2RL_CHECK_ARG(*g.left >= 1 && *g.left <= 624 && *g.next <= 624, "unexpected generator state (left %d, next %llu)",
                   (int)*g.left, (unsigned long long)*g.next);
    if (n == 0) return B2RL_OK;
    const int64_t iters = n - 1;                           // the shuffle's loop count = its number of draws
    const int64_t head = B < iters ? B : iters;
    SparsePerm r(2 * head + 2);
    for (int64_t i = 0; i < head; ++i) {
        const int64_t z = (int64_t)(g.draw() % (uint64_t)(n - i));
        const int64_t ri = r.get(i), rj = r.get(i + z);
        out_host[i] = rj;                                  // r[i] after swap(r[i], r[i + z]); never touched again
        r.set(i + z, ri);
    }
    for (int64_t i = head; i < B; ++i) out_host[i] = r.get(i);      // only i = n-1 (B == n): the entry no iteration finalises
    g.skip(iters - head);
    return B2RL_OK;
}
