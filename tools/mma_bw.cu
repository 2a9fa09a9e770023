// mma_bw.cu -- microbenchmark: tcgen05.mma issue/execute throughput per SM for the small-K shapes of the PFN kernel
// (sm_100a).  One CTA per SM; one thread issues `reps` rounds of a fixed MMA sequence and commits; the CTA measures
// clock64 from the first issue to the completion of the last commit.  Operand contents are irrelevant (zeros).
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bw tools/mma_bw.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool kTf32>
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    if constexpr (kTf32)
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
template <bool kTf32>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc) {
    if constexpr (kTf32)
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t make_idesc(bool tf32, int m, int n) {
    const uint32_t fmt = tf32 ? 2u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// mode bits: 1 = main MMAs (N = nmain), 2 = G MMAs (N = 16), 4 = A of the main MMAs from TMEM
template <bool kTf32>
__global__ void __launch_bounds__(128, 1) mma_kernel(long long* cycles, int reps, int mode, int nmain, int tiles, int commit_each) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int RB = kTf32 ? 128 : 64;
    unsigned char* sA = base;                      // 3 tiles x 128 rows x RB
    unsigned char* sB = sA + 3 * 128 * 128;        // 256 rows x RB
    unsigned char* sG = sB + 256 * 128;            // 16 rows
    for (int i = threadIdx.x; i < (3 * 128 * 128 + 256 * 128 + 16 * 128) / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        uint32_t leader;
        asm volatile("{.reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p;}" : "=r"(leader));
        const uint32_t sbo = 8 * RB;
        const uint32_t layout = kTf32 ? 2u : 4u;
        const uint32_t desc_hi = (sbo >> 4) | (1u << 14) | (layout << 29);
        const uint32_t a_lo = (smem_u32(sA) >> 4) | (1u << 16), b_lo = (smem_u32(sB) >> 4) | (1u << 16), g_lo = (smem_u32(sG) >> 4) | (1u << 16);
        const uint32_t idesc_main = make_idesc(kTf32, 128, nmain), idesc_g = make_idesc(kTf32, 128, 16);
        const int ksteps = kTf32 ? 4 : 2;
        const int stride = nmain + 16;  // TMEM columns per tile
        const uint32_t a_tmem = tb + 448;  // A operand columns (content irrelevant)
        uint32_t phase = 0;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (mode & 8) {  // k-outer, tile-inner: consecutive MMAs never touch the same accumulator
                if (leader) {
                    for (int k = 0; k < ksteps; ++k)
                        for (int m = 0; m < tiles; ++m) {
                            const uint32_t d_main = tb + (uint32_t)((m * stride) % 432);
                            if (mode & 4)
                                mma_ts<kTf32>(d_main, a_tmem + (uint32_t)(k * 8), ((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)(k * 2)), idesc_main, k > 0);
                            else
                                mma_ss<kTf32>(d_main, ((uint64_t)desc_hi << 32) | (a_lo + (uint32_t)((m * 128 * RB + k * 32) >> 4)),
                                              ((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)(k * 2)), idesc_main, k > 0);
                        }
                }
                __syncwarp();
                continue;
            }
            for (int m = 0; m < tiles; ++m) {
                const uint32_t d_main = tb + (uint32_t)((m * stride) % 432), d_g = d_main + nmain;
                if (leader) {
                if (mode & 2)
                    for (int k = 0; k < ksteps; ++k)
                        mma_ss<kTf32>(d_g, ((uint64_t)desc_hi << 32) | (a_lo + (uint32_t)((m * 128 * RB + k * 32) >> 4)),
                                      ((uint64_t)desc_hi << 32) | (g_lo + (uint32_t)(k * 2)), idesc_g, k > 0);
                if (mode & 1)
                    for (int k = 0; k < ksteps; ++k) {
                        if (mode & 4)
                            mma_ts<kTf32>(d_main, a_tmem + (uint32_t)(k * 8), ((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)(k * 2)), idesc_main, k > 0);
                        else
                            mma_ss<kTf32>(d_main, ((uint64_t)desc_hi << 32) | (a_lo + (uint32_t)((m * 128 * RB + k * 32) >> 4)),
                                          ((uint64_t)desc_hi << 32) | (b_lo + (uint32_t)(k * 2)), idesc_main, k > 0);
                    }
                }
                __syncwarp();
                if (commit_each) {
                    if (leader) tc_commit(&bar);
                    __syncwarp();
                    while (!mbar_try_wait(&bar, phase & 1)) {}
                    ++phase;
                }
            }
        }
        if (!commit_each) {
            if (leader) tc_commit(&bar);
            __syncwarp();
            while (!mbar_try_wait(&bar, 0)) {}
        }
        t1 = clock64();
        if (leader) cycles[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}

template <bool kTf32>
void run(const char* name, int mode, int nmain, int tiles, int commit_each, long long* cyc) {
    const int reps = 200;
    const size_t smem = 1024 + 3 * 128 * 128 + 256 * 128 + 16 * 128;
    cudaFuncSetAttribute(mma_kernel<kTf32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    mma_kernel<kTf32><<<148, 128, smem>>>(cyc, reps, mode, nmain, tiles, commit_each);
    cudaDeviceSynchronize();
    mma_kernel<kTf32><<<148, 128, smem>>>(cyc, reps, mode, nmain, tiles, commit_each);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    printf("%-5s %-28s N=%3d tiles=%d commit_each=%d  cycles/tile=%8.1f  (%s)\n", kTf32 ? "tf32" : "bf16", name, nmain, tiles,
           commit_each, avg / (reps * tiles), cudaGetErrorString(e));
}

int main() {
    long long* cyc;
    cudaMalloc(&cyc, 148 * 8);
    for (int ce = 0; ce < 2; ++ce) {
        run<true>("main SS", 1, 128, 3, ce, cyc);
        if (!ce) run<true>("main SS interleaved", 9, 128, 3, ce, cyc);
        if (!ce) run<true>("main TS interleaved", 13, 128, 3, ce, cyc);
        if (!ce) run<false>("main SS interleaved", 9, 128, 3, ce, cyc);
        if (!ce) run<false>("main TS interleaved", 13, 128, 3, ce, cyc);
        if (!ce) run<false>("main SS interleaved N=64", 9, 64, 3, ce, cyc);
        run<true>("G SS (N=16)", 2, 128, 3, ce, cyc);
        run<true>("G + main SS", 3, 128, 3, ce, cyc);
        run<true>("main A-in-TMEM", 5, 128, 3, ce, cyc);
        run<true>("main SS", 1, 64, 3, ce, cyc);
        run<true>("main A-in-TMEM", 5, 64, 3, ce, cyc);
        run<true>("main SS", 1, 256, 1, ce, cyc);
        run<true>("main A-in-TMEM", 5, 256, 1, ce, cyc);
        run<false>("main SS", 1, 128, 3, ce, cyc);
        run<false>("G SS (N=16)", 2, 128, 3, ce, cyc);
        run<false>("G + main SS", 3, 128, 3, ce, cyc);
        run<false>("main A-in-TMEM", 5, 128, 3, ce, cyc);
        run<false>("main SS", 1, 64, 3, ce, cyc);
        run<false>("main A-in-TMEM", 5, 64, 3, ce, cyc);
        run<false>("main SS", 1, 256, 1, ce, cyc);
        run<false>("main A-in-TMEM", 5, 256, 1, ce, cyc);
    }
    return 0;
}
