// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) for N in {64,128,256}, A from shared memory (SS)
// or tensor memory (TS), issued back-to-back by one elected thread; one CTA per SM.  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_rate tools/umma_rate.cu && /tmp/umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
    return (uint64_t)((a >> 4) & 0x3FFFu) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, bool TS, int NACC, int COMMITS, int BG>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long *out, const uint4 *gsrc) {
    __shared__ volatile int stop_flag;
    __shared__ uint64_t bar3;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2[2];
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        stop_flag = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base_s;
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        if (elect_one()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t dcol = tm + ((k & (NACC - 1)) * N);   // NACC independent accumulators
                    const uint64_t b = desc_sw128(base + 32768 + (k & 3) * 32);
                    if (TS) mma_ts(dcol, tm + 256 + (k & 3) * 8 + (k >> 2) * 32, b, idesc, 1u);
                    else mma_ss(dcol, desc_sw128(base + (k >> 2) * 16384 + (k & 3) * 32), b, idesc, 1u);
                }
#pragma unroll
                for (int c = 0; c < COMMITS; ++c)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[c])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            uint32_t done;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            } while (!done);
            t1 = clock64();
            out[blockIdx.x] = t1 - t0;
            stop_flag = 1;
        }
        __syncwarp();
    } else if (BG == 1) {
        // background: tcgen05.st into TMEM columns 384.. (not used by the MMAs) from warps 1..3
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) v[i] = i * 0x40004000u;
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + 384;
        while (!stop_flag) {
            asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]),"r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            for (volatile int d = 0; d < 8; ++d) {}
        }
    } else if (BG == 2 && warp == 1) {
        // background: 16 KB bulk copies global -> smem (region after the operands), ~ one per 512 cycles
        if (elect_one()) {
            uint32_t ph = 0;
            while (!stop_flag) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar3)), "r"(16384) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(base + 65536), "l"(gsrc + (blockIdx.x * 1024)), "r"(16384), "r"(smem_u32(&bar3)) : "memory");
                uint32_t done;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar3)), "r"(ph) : "memory");
                } while (!done);
                ph ^= 1;
            }
        }
        __syncwarp();
    } else if (BG == 3) {
        // background: generic st.shared.v4 traffic from warps 1..3 into the region after the operands
        uint4 *dst = reinterpret_cast<uint4 *>(smem + (base - smem_u32(smem)) + 65536) + threadIdx.x;
        uint4 v = make_uint4(1, 2, 3, 4);
        while (!stop_flag) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i * 128] = v;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

template <int N, bool TS, int NACC, int COMMITS = 0, int BG = 0>
void run(const char *name, int grid) {
    long long *out;
    cudaMalloc(&out, sizeof(long long) * 148);
    const int iters = 256, smem = 65536 + 32768 + 1024;
    uint4 *gsrc; cudaMalloc(&gsrc, 148 * 16384); cudaMemset(gsrc, 0, 148 * 16384);
    cudaFuncSetAttribute(rate_kernel<N, TS, NACC, COMMITS, BG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<N, TS, NACC, COMMITS, BG><<<grid, 128, smem>>>(iters, out, gsrc);
    rate_kernel<N, TS, NACC, COMMITS, BG><<<grid, 128, smem>>>(iters, out, gsrc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1ll << 60;
    for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
    printf("%-28s grid %3d: %7.1f .. %7.1f cycles per MMA  (%s)\n", name, grid, (double)mn / (iters * 8), (double)mx / (iters * 8), cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    for (int grid : {148}) {
        run<256, false, 2>("SS N=256 2 acc", grid);
        run<256, false, 1>("SS N=256 1 acc", grid);
        run<128, false, 2>("SS N=128 2 acc", grid);
        run<64, false, 2>("SS N=64  2 acc", grid);
        run<128, true, 2>("TS N=128 2 acc", grid);
        run<128, true, 1>("TS N=128 1 acc", grid);
        run<64, true, 2>("TS N=64  2 acc", grid);
        run<256, true, 1>("TS N=256 1 acc", grid);
        run<128, true, 2, 1>("TS N=128 1 commit / 8 MMA", grid);
        run<128, true, 2, 2>("TS N=128 2 commits / 8 MMA", grid);
        run<256, false, 2, 1>("SS N=256 1 commit / 8 MMA", grid);
        run<128, true, 2, 0, 1>("TS N=128 + STTM traffic", grid);
        run<128, true, 2, 0, 2>("TS N=128 + TMA 16KB copies", grid);
        run<128, true, 2, 0, 3>("TS N=128 + st.shared traffic", grid);
        run<256, false, 2, 0, 2>("SS N=256 + TMA 16KB copies", grid);
        run<256, false, 2, 0, 3>("SS N=256 + st.shared traffic", grid);
    }
    return 0;
}
