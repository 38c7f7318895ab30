// mma_bench.cu -- cost model of the evaluator's tcgen05.mma stream on one SM (timing experiment, not product code).
//   tools/_bin/mma_bench
// One CTA per SM (or two) issues REPS x CHAIN kind::f16 MMAs of M = 128, N in {64,128,192,256}, K = 16 from K-major,
// no-swizzle operands in shared memory laid out like evalnet.cu's (A: [k chunk][160 slots][16 B], B: [k chunk][rows][16 B]),
// and reports cycles per MMA for: same operands every time / A walking the image, B walking a ring of stages.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ULL << 46);
}
__device__ __forceinline__ uint32_t idesc_f16(int m, int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24); }
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 2) k_bench(int mode, int reps, long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    const int warp = threadIdx.x >> 5;
    constexpr int CG_STRIDE = 160 * 16, IMG = 8 * CG_STRIDE;                 // one 64-channel fp16 image
    constexpr int STAGE = 2 * 2 * N * 16;                                     // one k-step of B: 2 chunks x (hi|lo would be 2N rows) -> here N rows x 2 parts
    const uint32_t s_base = smem_u32(smem);
    for (int i = threadIdx.x; i < (2 * IMG + 4 * STAGE) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // halfs = 1.0
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = idesc_f16(128, N);
        const uint64_t a0 = smem_desc(s_base, CG_STRIDE, 128);
        const uint32_t b_base = s_base + 2 * IMG;
        const uint32_t b_lbo = 2 * N * 16;                                   // [chunk][hi rows N | lo rows N][16 B]
        const uint64_t b0 = smem_desc(b_base, b_lbo, 128);
        const long long t0 = clock64();
        uint32_t phase = 0;
        for (int r = 0; r < reps; ++r) {
            // one "layer" of the evaluator: 3 dy x 4 k-steps x 3 MMAs
            for (int dy = 0; dy < 3; ++dy)
                for (int ks = 0; ks < 4; ++ks) {
                    uint64_t a = a0, al = a0 + (IMG >> 4), b = b0;
                    if (mode >= 1) { a += (uint32_t)(16 * dy) + (uint32_t)ks * ((2 * CG_STRIDE) >> 4); al += (uint32_t)(16 * dy) + (uint32_t)ks * ((2 * CG_STRIDE) >> 4); }
                    if (mode >= 2) b += (uint32_t)((dy * 4 + ks) & 3) * (STAGE >> 4);
                    const uint32_t lo = (N * 16) >> 4;
                    mma(tmem, a, b, idesc, (dy | ks) ? 1u : 0u);
                    mma(tmem, a, b + lo, idesc, 1u);
                    mma(tmem, al, b, idesc, 1u);
                }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            if (mode == 3 || r == reps - 1) {                                // mode 3: wait for every layer like the real kernel
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            }
            phase ^= 1u;
        }
        out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
}

template <int N> void run(int ctas_per_sm, int sms) {
    const int reps = 200;
    long long* d; cudaMalloc(&d, 1024 * sizeof(long long));
    const int smem = 2 * 8 * 160 * 16 + 4 * (2 * 2 * N * 16) + 1024;
    cudaFuncSetAttribute(k_bench<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int mode = 0; mode < 4; ++mode) {
        k_bench<N><<<sms * ctas_per_sm, 128, smem>>>(mode, reps, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d mode %d: %s\n", N, mode, cudaGetErrorString(e)); return; }
        long long h[1024]; cudaMemcpy(h, d, sizeof(long long) * sms * ctas_per_sm, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < sms * ctas_per_sm; ++i) avg += h[i];
        avg /= sms * ctas_per_sm;
        const double per_mma = avg / (reps * 36.0), ideal = 128.0 * N * 16 / 4096.0;
        printf("N=%3d ctas/SM=%d mode %d (%s): %.1f cycles per MMA per CTA (%.1f per SM-MMA; math at nominal peak %.0f)\n", N, ctas_per_sm, mode,
               mode == 0 ? "same operands" : mode == 1 ? "A walks" : mode == 2 ? "A and B walk" : "A and B walk, wait per layer", per_mma, per_mma / ctas_per_sm, ideal);
    }
    cudaFree(d);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    for (int c = 1; c <= 2; ++c) { run<64>(c, p.multiProcessorCount); run<128>(c, p.multiProcessorCount); run<192>(c, p.multiProcessorCount); run<256>(c, p.multiProcessorCount); }
    return 0;
}
