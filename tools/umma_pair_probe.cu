// Probe for the CTA-pair form of the int8 path: tcgen05.mma.cta_group::2.kind::i8, M = 256 (128
// rows per CTA, A operand in EACH CTA's tensor memory), N = 256 with the K-major SWIZZLE_64B B tile split between the two
// CTAs' shared memories (rows [0,128) in the leader, [128,256) in the peer), commit multicast to both CTAs.
//   1. correctness of one 256 x 256 x 64 product against a CPU product;
//   2. cycles per instruction (expected 128 for M256 N256 K32 if the pair runs at the single-CTA MAC rate per SM).
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_pair_probe tools/umma_pair_probe.cu && /tmp/umma_pair_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t addr) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ void mma_i8_ts_pair(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void commit_pair(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

static size_t image_offset(int n, int k) {   // SWIZZLE_64B, 512-byte atoms (n = row inside this CTA's half)
    const int c16 = k >> 4, b = k & 15, r = n & 7, grp = n >> 3;
    return (size_t)grp * 512 + r * 64 + ((c16 ^ ((r >> 1) & 3)) * 16) + b;
}

// iters == 0: correctness (one product, D written out); iters > 0: issue-rate loop
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
pair_kernel(const uint32_t *a_words /*[256][16]*/, const uint4 *b_image /*[2][8 KB]*/, int32_t *d_out /*[256][256]*/, int iters,
            long long *cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    uint8_t *bgen = smem + (base - smem_u32(smem));
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_rank();
    const int pair = blockIdx.x >> 1;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // the same warp of BOTH CTAs allocates
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 8192 / 16; i += 128) reinterpret_cast<uint4 *>(bgen)[i] = b_image[rank * 512 + i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base_s;
    uint32_t a[16];
    for (int i = 0; i < 16; ++i) a[i] = a_words[(rank * 128 + threadIdx.x) * 16 + i];
    const uint32_t t_lane = tm + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(t_lane + 256), "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(a[3]),"r"(a[4]),"r"(a[5]),"r"(a[6]),"r"(a[7]),
                    "r"(a[8]),"r"(a[9]),"r"(a[10]),"r"(a[11]),"r"(a[12]),"r"(a[13]),"r"(a[14]),"r"(a[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync();   // both CTAs: barrier initialised, TMEM allocated, A stored, B in shared memory
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // D int32 | A, B signed int8 | N = 256 | M = 256 (pair)
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((256u >> 4) << 24);
    if (rank == 0 && warp == 0) {
        if (elect_one()) {
            const long long t0 = clock64();
            const int n = iters > 0 ? iters : 1;
            for (int it = 0; it < n; ++it)
                for (int k = 0; k < 2; ++k) mma_i8_ts_pair(tm, tm + 256 + k * 8, desc_sw64(base + k * 32), idesc, (k > 0 || it > 0) ? 1u : 0u);
            commit_pair(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            if (cycles) cycles[pair] = clock64() - t0;
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);   // both CTAs: the commit is multicast
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (iters == 0) {
        for (int c0 = 0; c0 < 256; c0 += 16) {
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),
                           "=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15])
                         : "r"(t_lane + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 16; ++i) d_out[(rank * 128 + threadIdx.x) * 256 + c0 + i] = (int32_t)v[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync();   // nobody frees tensor memory / exits while the peer may still be using the pair's state
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

int main() {
    static int8_t A[256][64], B[256][64];
    static int32_t ref[256][256], got[256][256];
    srand(99);
    for (int r = 0; r < 256; ++r)
        for (int k = 0; k < 64; ++k) A[r][k] = (rand() % 100 < 15) ? (int8_t)(1 << ((k / 4 + r) % 7)) : 0;
    for (int n = 0; n < 256; ++n)
        for (int k = 0; k < 64; ++k) B[n][k] = (int8_t)(rand() % 256 - 128);
    for (int r = 0; r < 256; ++r)
        for (int n = 0; n < 256; ++n) {
            int32_t s = 0;
            for (int k = 0; k < 64; ++k) s += (int32_t)A[r][k] * (int32_t)B[n][k];
            ref[r][n] = s;
        }
    uint8_t *img = (uint8_t *)calloc(2 * 8192, 1);
    for (int n = 0; n < 256; ++n)
        for (int k = 0; k < 64; ++k) img[(n / 128) * 8192 + image_offset(n % 128, k)] = (uint8_t)B[n][k];
    uint32_t *da; uint4 *db; int32_t *dd; long long *dc;
    cudaMalloc(&da, sizeof(A)); cudaMalloc(&db, 2 * 8192); cudaMalloc(&dd, sizeof(ref)); cudaMalloc(&dc, 74 * sizeof(long long));
    cudaMemcpy(da, A, sizeof(A), cudaMemcpyHostToDevice);
    cudaMemcpy(db, img, 2 * 8192, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xFF, sizeof(ref));
    const int smem = 8192 + 2048;
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pair_kernel<<<2, 128, smem>>>(da, db, dd, 0, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(got, dd, sizeof(ref), cudaMemcpyDeviceToHost);
    long bad = 0; int fr = -1, fn = -1;
    for (int r = 0; r < 256; ++r)
        for (int n = 0; n < 256; ++n)
            if (got[r][n] != ref[r][n]) { if (!bad) { fr = r; fn = n; } ++bad; }
    printf("pair check M=256 N=256 K=64: %s, %ld / %d mismatches", cudaGetErrorString(e), bad, 256 * 256);
    if (bad) printf("  first at (%d,%d): got %d want %d", fr, fn, got[fr][fn], ref[fr][fn]);
    printf("\n");
    if (e != cudaSuccess) return 1;
    for (int rep = 0; rep < 2; ++rep) {
        const int iters = 2048;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        pair_kernel<<<148, 128, smem>>>(da, db, dd, iters, dc);
        cudaEventRecord(e1);
        e = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        long long h[74];
        cudaMemcpy(h, dc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0, mn = 1ll << 60;
        for (int i = 0; i < 74; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
        const double macs = 74.0 * iters * 2 * 256.0 * 256 * 32;
        printf("pair rate: %.1f .. %.1f cycles per M256 N256 K32 MMA, %.1f us, %.0f dense TOP/s (%s)\n", (double)mn / (iters * 2),
               (double)mx / (iters * 2), ms * 1e3, 2.0 * macs / (ms * 1e-3) * 1e-12, cudaGetErrorString(e));
    }
    return 0;
}
