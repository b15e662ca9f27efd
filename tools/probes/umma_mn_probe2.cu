// Probe 2: does MN-major work at all?  (a) kind::f16 with bf16 operands, A MN-major no-swizzle; (b) kind::tf32 with B MN-major.
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF); d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16; d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32; d |= (uint64_t)1 << 46;
    return d;
}
// mode 0: bf16, A MN-major (m/8)*SBO + (k/8)*LBO + (k%8)*16 + (m%8)*2, B K-major; K = 16 per MMA, 2 MMAs (k = 32)
// mode 1: bf16, A K-major reference
// mode 2: tf32, A K-major, B MN-major: (n/4)*SBO + (k/8)*LBO + (k%8)*16 + (n%4)*4
__global__ void probe(const float *A, const float *B, float *D, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    uint8_t *gen = smem + (base - smem_u32(smem));
    uint8_t *As = gen, *Bs = gen + 16384;
    for (int e = tid; e < 128 * 32; e += blockDim.x) {
        const int m = e / 32, k = e % 32;
        if (mode == 0) *reinterpret_cast<__nv_bfloat16 *>(As + (m / 8) * 512 + (k / 8) * 128 + (k % 8) * 16 + (m % 8) * 2) = __float2bfloat16(A[e]);
        else if (mode == 1) *reinterpret_cast<__nv_bfloat16 *>(As + (k / 8) * 2048 + (m / 8) * 128 + (m % 8) * 16 + (k % 8) * 2) = __float2bfloat16(A[e]);
        else if (mode == 2) *reinterpret_cast<float *>(As + (k / 4) * 2048 + (m / 8) * 128 + (m % 8) * 16 + (k % 4) * 4) = A[e];
        else As[(m / 16) * 512 + (k / 8) * 128 + (k % 8) * 16 + (m % 16)] = (uint8_t)(int)(A[e] + 6.f);       // u8 0..12
    }
    for (int e = tid; e < 32 * 32; e += blockDim.x) {
        const int n = e / 32, k = e % 32;
        if (mode < 2) *reinterpret_cast<__nv_bfloat16 *>(Bs + (k / 8) * 512 + (n / 8) * 128 + (n % 8) * 16 + (k % 8) * 2) = __float2bfloat16(B[e]);
        else if (mode == 2) *reinterpret_cast<float *>(Bs + (n / 4) * 512 + (k / 8) * 128 + (k % 8) * 16 + (n % 4) * 4) = B[e];
        else Bs[(k / 16) * 512 + (n / 8) * 128 + (n % 8) * 16 + (k % 16)] = (uint8_t)(int8_t)(int)B[e];
    }
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_ptr;
    if (tid == 0) {
        if (mode < 2) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((mode == 0 ? 1u : 0u) << 15) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
            for (int ks = 0; ks < 2; ++ks) {        // K = 16 per MMA
                const uint64_t da = mode == 0 ? make_desc(base + ks * 2 * 128, 128, 512) : make_desc(base + ks * 2 * 2048, 2048, 128);
                const uint64_t db = make_desc(base + 16384 + ks * 2 * 512, 512, 128);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)ks) : "memory");
            }
        } else if (mode == 3) {
            const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | ((32u >> 3) << 17) | ((128u >> 4) << 24);   // S32, U8 x S8, A MN-major
            const uint64_t da = make_desc(base, 128, 512);       // one MMA: K = 32 = 4 k-groups of 8 at LBO = 128; 8 m-chunks of 16 at SBO = 512
            const uint64_t db = make_desc(base + 16384, 512, 128);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        } else {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
            for (int ks = 0; ks < 4; ++ks) {        // K = 8 per MMA
                const uint64_t da = make_desc(base + ks * 2 * 2048, 2048, 128);
                const uint64_t db = make_desc(base + 16384 + ks * 128, 128, 512);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)ks) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DN;\n\tbra W;\n\tDN:\n\t}\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(tmem_d + ((uint32_t)(warp * 32) << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 32 + j] = mode == 3 ? (float)(int)r[j] : __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_d) : "memory");
}
int main() {
    static float hA[128 * 32], hB[32 * 32], hD[128 * 32], ref[128 * 32];
    for (int i = 0; i < 128 * 32; ++i) hA[i] = (float)((i * 7 + i / 32) % 13 - 6);
    for (int i = 0; i < 32 * 32; ++i) hB[i] = (float)((i * 5 + i / 32) % 11 - 5);
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) { float s = 0; for (int k = 0; k < 32; ++k) s += hA[m * 32 + k] * hB[n * 32 + k]; ref[m * 32 + n] = s; }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mode = 0; mode < 4; ++mode) {
        cudaMemset(dD, 0xFF, sizeof(hD));
        probe<<<1, 128, 40 * 1024>>>(dA, dB, dD, mode);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
        float err = 0;
        if (mode == 3) for (int m = 0; m < 128; ++m) for (int n = 0; n < 32; ++n) { float s2 = 0; for (int k = 0; k < 32; ++k) s2 += (hA[m * 32 + k] + 6.f) * hB[n * 32 + k]; ref[m * 32 + n] = s2; }
        for (int i = 0; i < 128 * 32; ++i) { float d = hD[i] - ref[i]; if (d < 0) d = -d; if (d != d) d = 1e30f; if (d > err) err = d; }
        printf("mode %d (%s): %s max err %g, D[0][0..3] = %g %g %g %g (ref %g %g %g %g)\n", mode,
               mode == 0 ? "bf16 A MN-major" : mode == 1 ? "bf16 A K-major" : mode == 2 ? "tf32 B MN-major" : "i8: u8 A MN-major x s8 B K-major", cudaGetErrorString(e), err, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
    }
    return 0;
}
