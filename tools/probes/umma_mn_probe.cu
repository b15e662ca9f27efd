// Probe: tcgen05.mma kind::tf32 with an MN-major A operand in the no-swizzle canonical layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_mn_probe umma_mn_probe.cu && ./umma_mn_probe
// Fills A[m][k] (m < 128, k < 32) and B[n][k] (n < 32) with small integers, runs 4 MMAs (K = 8 each) under several
// descriptor encodings and prints the max error of D against the host product for each.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout = 0) {
    uint64_t d = layout << 61;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

// variant: 0 = (LBO = k-group stride, SBO = m-chunk stride), 1 = swapped
__global__ void probe(const float *A, const float *B, float *D, int variant, uint32_t a_major_bit) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_ptr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    uint8_t *gen = smem + (base - smem_u32(smem));
    const uint32_t SBO = 512, LBO = 128;                     // A: 32 m-chunks x (4 k-groups x 128 B)
    float *As = reinterpret_cast<float *>(gen);               // 16 KB
    float *Bs = reinterpret_cast<float *>(gen + 16384);       // B K-major: (k/4)*LBOb + (n/8)*128 + (n%8)*16 + (k%4)*4, LBOb = 32*16
    for (int e = tid; e < 128 * 32; e += blockDim.x) {
        const int m = e / 32, k = e % 32;
        if (variant < 2) As[((m / 4) * SBO + (k / 8) * LBO + (k % 8) * 16 + (m % 4) * 4) / 4] = A[m * 32 + k];
        else             // SWIZZLE_128B MN-major: atoms of 32 m x 8 k (1 KB), 16-byte chunks XORed with the k row
            As[((m / 32) * 4096 + (k / 8) * 1024 + (k % 8) * 128 + ((((m % 32) / 4) ^ (k % 8)) * 16) + (m % 4) * 4) / 4] = A[m * 32 + k];
    }
    for (int e = tid; e < 32 * 32; e += blockDim.x) {
        const int n = e / 32, k = e % 32;
        Bs[((k / 4) * 512 + (n / 8) * 128 + (n % 8) * 16 + (k % 4) * 4) / 4] = B[n * 32 + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (a_major_bit << 15) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        for (int pg = 0; pg < 4; ++pg) {
            uint64_t da;
            if (variant == 0) da = make_desc(base + pg * LBO, LBO, SBO);
            else if (variant == 1) da = make_desc(base + pg * LBO, SBO, LBO);
            else if (variant == 2) da = make_desc(base + pg * 1024, 4096, 1024, 2);
            else da = make_desc(base + pg * 1024, 1024, 4096, 2);
            const uint64_t db = make_desc(base + 16384 + pg * 2 * 512, 512, 128);
            mma_tf32(tmem_d, da, db, idesc, pg ? 1u : 0u);
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
        for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem_d) : "memory");
}

int main() {
    float hA[128 * 32], hB[32 * 32], hD[128 * 32], ref[128 * 32];
    for (int i = 0; i < 128 * 32; ++i) hA[i] = (float)((i * 7 + i / 32) % 13 - 6);
    for (int i = 0; i < 32 * 32; ++i) hB[i] = (float)((i * 5 + i / 32) % 11 - 5);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 32; ++n) {
            float s = 0;
            for (int k = 0; k < 32; ++k) s += hA[m * 32 + k] * hB[n * 32 + k];
            ref[m * 32 + n] = s;
        }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, sizeof(hD));
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int variant = 0; variant < 4; ++variant)
        for (uint32_t bit = 0; bit < 2; ++bit) {
            cudaMemset(dD, 0xFF, sizeof(hD));
            probe<<<1, 128, 40 * 1024>>>(dA, dB, dD, variant, bit);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(hD, dD, sizeof(hD), cudaMemcpyDeviceToHost);
            float err = 0, mx = 0;
            for (int i = 0; i < 128 * 32; ++i) { float d = hD[i] - ref[i]; if (d < 0) d = -d; if (d > err || d != d) err = d != d ? 1e30f : d; if (hD[i] > mx) mx = hD[i]; }
            printf("variant %d (LBO/SBO %s) a_major=%u: %s max err %g, max |D| %g, D[0][0..3] = %g %g %g %g (ref %g %g %g %g)\n", variant,
                   variant == 0 ? "as canonical" : variant == 1 ? "swapped" : variant == 2 ? "SW128 lbo=atomMN sbo=kgroup" : "SW128 swapped", bit, cudaGetErrorString(e), err, mx, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1],
                   ref[2], ref[3]);
        }
    return 0;
}
