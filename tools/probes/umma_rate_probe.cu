// Probe: issue rate of tcgen05.mma kind::tf32 (M = 128) from shared memory, per operand layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_rate_probe umma_rate_probe.cu && ./umma_rate_probe
// One CTA issues `iters` rounds of 8 dependent-accumulate MMAs (K = 8 each, consecutive k-steps of one 64-tap stage) and
// reports clock64 cycles per MMA.  Operand contents are irrelevant (zero-filled); addresses and strides are the ones the
// convolution kernels use.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
    uint64_t d = layout << 61;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
template <int KIND>   // 0 tf32, 1 f16 (bf16), 2 i8
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else if (KIND == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: tf32 no-swizzle K-major, A rows_p = 96 (LBO 1536), B N rows (LBO N*16)   [what conv_st / conv_sg use]
// mode 1: tf32 no-swizzle, A rows 128 (LBO 2048)                                   [conv_tc]
// mode 2: tf32 SWIZZLE_128B K-major: rows of 128 B (32 tf32), SBO = 1024, k-step advances the start address by 32 B
// mode 3: bf16 no-swizzle K-major (K = 16 per MMA), LBO = 2048
// mode 4: i8 no-swizzle K-major (K = 32 per MMA), LBO = 2048                        [conv_i8]
template <int KIND>
__global__ void probe(int mode, int N, int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem + (base - smem_u32(smem)))[i] = 0;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_ptr;
    if (tid == 0) {
        const uint32_t fmt = KIND == 0 ? 2u : KIND == 1 ? 1u : 0u;
        const uint32_t cfmt = KIND == 2 ? 2u : 1u;
        const uint32_t idesc = (cfmt << 4) | (fmt << 7) | ((KIND == 2 ? 1u : fmt) << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_addr = base, b_addr = base + 96 * 1024;
        uint64_t da, db, da_step, db_step;
        if (mode == 0) { da = make_desc(a_addr, 1536, 128, 0); db = make_desc(b_addr, N * 16, 128, 0); da_step = (2 * 1536) >> 4; db_step = (2 * N * 16) >> 4; }
        else if (mode == 1 || mode == 3 || mode == 4) { da = make_desc(a_addr, 2048, 128, 0); db = make_desc(b_addr, N * 16, 128, 0); da_step = (2 * 2048) >> 4; db_step = (2 * N * 16) >> 4; }
        else { da = make_desc(a_addr, 16, 1024, 2); db = make_desc(b_addr, 16, 1024, 2); da_step = 32 >> 4; db_step = 32 >> 4; }
        const int ksteps = mode == 2 ? 4 : 8;       // SW128: 4 k-steps of 32 B inside one 128-byte row
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it)
            for (int j = 0; j < ksteps; ++j) mma<KIND>(tmem_d, da + j * da_step, db + j * db_step, idesc, (it | j) ? 1u : 0u);
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DN;\n\tbra W;\n\tDN:\n\t}\n" ::"r"(smem_u32(&bar)) : "memory");
        const long long t1 = clock64();
        out[0] = t1 - t0;
        out[1] = (long long)iters * ksteps;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_d) : "memory");
}

template <int KIND>
void run(const char *name, int mode, int N, long long *dout) {
    cudaFuncSetAttribute(probe<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    long long h[2] = {0, 0};
    for (int rep = 0; rep < 2; ++rep) {
        probe<KIND><<<1, 128, 170 * 1024>>>(mode, N, 256, dout);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
        if (rep == 1) printf("%-44s N=%3d: %s  %.1f cycles per MMA (%lld MMAs)\n", name, N, cudaGetErrorString(e), (double)h[0] / (double)h[1], h[1]);
    }
}

int main() {
    long long *dout;
    cudaMalloc(&dout, 16);
    for (int N : {32, 64, 128, 256}) {
        run<0>("tf32 no-swizzle, A pitch 1536 (conv_sg)", 0, N, dout);
        run<0>("tf32 no-swizzle, A pitch 2048 (conv_tc)", 1, N, dout);
        run<0>("tf32 SWIZZLE_128B", 2, N, dout);
        run<1>("bf16 no-swizzle", 3, N, dout);
        run<2>("i8   no-swizzle (conv_i8)", 4, N, dout);
    }
    return 0;
}
