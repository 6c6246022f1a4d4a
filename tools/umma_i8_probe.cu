// Probe for the int8 tensor-core path (tcgen05.mma kind::i8, A from TENSOR MEMORY, B from shared memory):
//   1. correctness of one M128 x N x K64 product for several shared-memory layouts of the K-major int8 B tile
//      (no swizzle / SWIZZLE_64B / SWIZZLE_128B with half-used rows) against a CPU product;
//   2. cycles per instruction of kind::i8 (K=32) next to kind::f16 (K=16), one CTA per SM, one elected issuer;
//   3. completion latency of tcgen05.st (x16 / x32) with and without a stream of MMAs in flight (not yet run: added at
//      the end of round 1 to test whether the TMEM store latency is what paces the A producers).
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_i8_probe tools/umma_i8_probe.cu && /tmp/umma_i8_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred px;\n\telect.sync _|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
// shared-memory matrix descriptor: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | layout << 61
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma_i8_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

// layouts: 0 = no swizzle (8 x 16 B core matrices, LBO = 128 between K neighbours, SBO = 512 between 8-row groups)
//          1 = SWIZZLE_64B  (rows of 64 B, 8-row atoms of 512 B, 16-byte chunk ^= (row >> 1) & 3)
//          2 = SWIZZLE_128B (rows of 128 B of which the first 64 hold data, atoms of 1024 B, chunk ^= row & 7)
static size_t image_offset(int layout, int n, int k) {
    const int c16 = k >> 4, b = k & 15, r = n & 7, grp = n >> 3;
    if (layout == 0) return (size_t)(grp * 4 + c16) * 128 + r * 16 + b;
    if (layout == 1) return (size_t)grp * 512 + r * 64 + ((c16 ^ ((r >> 1) & 3)) * 16) + b;
    return (size_t)grp * 1024 + r * 128 + ((c16 ^ r) * 16) + b;
}
static size_t image_bytes(int layout, int N) { return (size_t)(N / 8) * (layout == 2 ? 1024 : 512); }

template <int N>
__global__ void __launch_bounds__(128, 1) probe_kernel(int layout, uint32_t a_signed, const uint32_t *a_words /*[128][16]*/,
                                                       const uint4 *b_image, int b_bytes, int32_t *d_out /*[128][N]*/) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    uint8_t *bgen = smem + (base - smem_u32(smem));
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < b_bytes / 16; i += 128) reinterpret_cast<uint4 *>(bgen)[i] = b_image[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base_s;
    // A: thread r owns TMEM lane r; 16 words = 64 int8 K elements at columns 256..271
    uint32_t a[16];
    for (int i = 0; i < 16; ++i) a[i] = a_words[threadIdx.x * 16 + i];
    const uint32_t t_lane = tm + ((uint32_t)(warp * 32) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 :: "r"(t_lane + 256), "r"(a[0]),"r"(a[1]),"r"(a[2]),"r"(a[3]),"r"(a[4]),"r"(a[5]),"r"(a[6]),"r"(a[7]),
                    "r"(a[8]),"r"(a[9]),"r"(a[10]),"r"(a[11]),"r"(a[12]),"r"(a[13]),"r"(a[14]),"r"(a[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // c_format S32 (2) | a_format | b_format signed (1) | N >> 3 | M >> 4
    const uint32_t idesc = (2u << 4) | (a_signed << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    if (warp == 0) {
        if (elect_one()) {
            for (int k = 0; k < 2; ++k) {
                uint64_t bd;
                if (layout == 0) bd = make_desc(base + k * 256, 128, 512, 0);
                else if (layout == 1) bd = make_desc(base + k * 32, 0, 512, 4);
                else bd = make_desc(base + k * 32, 0, 1024, 2);
                mma_i8_ts(tm, tm + 256 + k * 8, bd, idesc, k > 0 ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(&bar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),
                       "=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15])
                     : "r"(t_lane + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; ++i) d_out[threadIdx.x * N + c0 + i] = (int32_t)v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

template <int N>
static void check(int layout, uint32_t a_signed) {
    static int8_t A[128][64], B[N][64];
    srand(1234 + layout);
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 64; ++k) A[r][k] = (rand() % 100 < 15) ? (int8_t)(1 << ((k / 4 + r) % 7)) : 0;   // weighted bits
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < 64; ++k) B[n][k] = (int8_t)(rand() % 256 - 128);
    static int32_t ref[128][N];
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n) {
            int32_t s = 0;
            for (int k = 0; k < 64; ++k) s += (int32_t)A[r][k] * (int32_t)B[n][k];
            ref[r][n] = s;
        }
    const size_t ib = image_bytes(layout, N);
    uint8_t *img = (uint8_t *)calloc(ib, 1);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < 64; ++k) img[image_offset(layout, n, k)] = (uint8_t)B[n][k];
    uint32_t *da; uint4 *db; int32_t *dd;
    cudaMalloc(&da, sizeof(A)); cudaMalloc(&db, ib); cudaMalloc(&dd, sizeof(ref));
    cudaMemcpy(da, A, sizeof(A), cudaMemcpyHostToDevice);   // little-endian: word c of row r = K elements 4c..4c+3
    cudaMemcpy(db, img, ib, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xFF, sizeof(ref));
    const int smem = (int)ib + 2048;
    cudaFuncSetAttribute(probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<N><<<1, 128, smem>>>(layout, a_signed, da, db, (int)ib, dd);
    cudaError_t e = cudaDeviceSynchronize();
    static int32_t got[128][N];
    cudaMemcpy(got, dd, sizeof(ref), cudaMemcpyDeviceToHost);
    long bad = 0; int fr = -1, fn = -1;
    for (int r = 0; r < 128; ++r)
        for (int n = 0; n < N; ++n)
            if (got[r][n] != ref[r][n]) { if (!bad) { fr = r; fn = n; } ++bad; }
    printf("check N=%3d layout=%d a_signed=%u: %s, %ld / %d mismatches", N, layout, a_signed, cudaGetErrorString(e), bad, 128 * N);
    if (bad) printf("  first at (%d,%d): got %d want %d", fr, fn, got[fr][fn], ref[fr][fn]);
    printf("\n");
    cudaFree(da); cudaFree(db); cudaFree(dd); free(img);
}

// ---- issue rate ----------------------------------------------------------------------------------------------------
template <int N, bool I8, int LAYOUT>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 32768 / 16; i += 128) reinterpret_cast<uint4 *>(smem + (base - smem_u32(smem)))[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base_s;
    constexpr uint32_t idesc = I8 ? ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24))
                                  : ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24));
    if (warp == 0) {
        if (elect_one()) {
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t dcol = tm + ((k >> 2) & 1) * (N <= 192 ? N : 0);   // two accumulators when they fit
                    uint64_t b;
                    if (!I8) b = make_desc(base + (k & 3) * 32, 0, 1024, 2);
                    else if (LAYOUT == 0) b = make_desc(base + (k & 1) * 256, 128, 512, 0);
                    else if (LAYOUT == 1) b = make_desc(base + (k & 1) * 32, 0, 512, 4);
                    else b = make_desc(base + (k & 1) * 32, 0, 1024, 2);
                    if (I8) mma_i8_ts(dcol, tm + 400 + (k & 1) * 8 + ((k >> 1) & 3) * 16, b, idesc, 1u);
                    else mma_f16_ts(dcol, tm + 400 + (k & 3) * 8 + (k >> 2) * 32, b, idesc, 1u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            mbar_wait(smem_u32(&bar), 0);
            out[blockIdx.x] = clock64() - t0;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

template <int N, bool I8, int LAYOUT>
static void rate(const char *name) {
    long long *out;
    cudaMalloc(&out, sizeof(long long) * 148);
    const int iters = 512, smem = 32768 + 2048, grid = 148;
    cudaFuncSetAttribute(rate_kernel<N, I8, LAYOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<N, I8, LAYOUT><<<grid, 128, smem>>>(iters, out);
    cudaEventRecord(e0);
    rate_kernel<N, I8, LAYOUT><<<grid, 128, smem>>>(iters, out);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148];
    cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    long long mx = 0, mn = 1ll << 60;
    for (int i = 0; i < grid; ++i) { mx = h[i] > mx ? h[i] : mx; mn = h[i] < mn ? h[i] : mn; }
    const double macs = (double)grid * iters * 8 * 128.0 * N * (I8 ? 32 : 16);
    printf("%-34s %7.1f .. %7.1f cycles per MMA, %.1f us, %.0f dense T(FL)OP/s, ~%.0f MHz  (%s)\n", name, (double)mn / (iters * 8),
           (double)mx / (iters * 8), ms * 1e3, 2.0 * macs / (ms * 1e-3) * 1e-12, (double)mx / (ms * 1e-3) * 1e-6, cudaGetErrorString(e));
    cudaFree(out);
}

// ---- tcgen05.st completion latency, with and without a stream of MMAs in flight ------------------------------------------
// Warps 1..3 store 16 (or 32) columns into their TMEM lanes and time `tcgen05.st` -> `tcgen05.wait::st`; warp 0's elected
// lane keeps the tensor pipe busy with kind::i8 MMAs (A from TMEM, B from shared memory) when `with_mma` is set.  The A
// producers of bm_mma_kernel wait for exactly this latency once per unit (deferred by one unit).
template <int COLS>
__global__ void __launch_bounds__(128, 1) sttm_latency_kernel(int iters, int with_mma, long long *out /*[grid][3]*/) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ int done_warps;   // the timing warps count themselves out; the MMA lane stops when all three are done
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        done_warps = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 16384 / 16; i += 128) reinterpret_cast<uint4 *>(smem + (base - smem_u32(smem)))[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base_s;
    constexpr uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    if (warp == 0) {
        if (with_mma && elect_one()) {
            while (*(volatile int *)&done_warps < 3) {
#pragma unroll
                for (int k = 0; k < 8; ++k)   // A at columns 256.., accumulators 0..255: disjoint from the stores at 384..
                    mma_i8_ts(tm + (k & 1) * 128, tm + 256 + (k & 1) * 8, make_desc(base + (k & 1) * 32, 0, 512, 4), idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            mbar_wait(smem_u32(&bar), 0);
        }
        __syncwarp();
    } else {
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) v[i] = i * 0x01010101u + lane;
        const uint32_t taddr = tm + ((uint32_t)(warp * 32) << 16) + 384;
        long long sum = 0, mx = 0;
        for (int i = 0; i < iters; ++i) {
            const long long t0 = clock64();
            if (COLS == 16)
                asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                             :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),
                                "r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]) : "memory");
            else
                asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
                             :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]),
                                "r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]) : "memory");
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            const long long dt = clock64() - t0;
            sum += dt;
            mx = dt > mx ? dt : mx;
            for (volatile int d = 0; d < 16; ++d) {}
        }
        if (lane == 0) {
            out[(blockIdx.x * 3 + (warp - 1)) * 2] = sum / iters;
            out[(blockIdx.x * 3 + (warp - 1)) * 2 + 1] = mx;
            atomicAdd(&done_warps, 1);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
    }
}

template <int COLS>
static void sttm_latency(int with_mma) {
    const int grid = 148, iters = 2000, smem = 16384 + 2048;
    long long *out;
    cudaMalloc(&out, sizeof(long long) * grid * 6);
    cudaFuncSetAttribute(sttm_latency_kernel<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    sttm_latency_kernel<COLS><<<grid, 128, smem>>>(iters, with_mma, out);
    cudaError_t e = cudaDeviceSynchronize();
    static long long h[148 * 6];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; long long mx = 0;
    for (int i = 0; i < grid * 3; ++i) { avg += (double)h[2 * i]; mx = h[2 * i + 1] > mx ? h[2 * i + 1] : mx; }
    printf("tcgen05.st x%d -> wait::st: avg %.0f cycles, max %lld, %s MMAs in flight (%s)\n", COLS, avg / (grid * 3), mx,
           with_mma ? "with" : "without", cudaGetErrorString(e));
    cudaFree(out);
}

int main() {
    for (int layout = 0; layout < 3; ++layout) {
        check<128>(layout, 1);
        check<128>(layout, 0);
    }
    check<192>(0, 1); check<192>(1, 1);
    check<64>(0, 1); check<64>(1, 1);
    check<96>(0, 1); check<96>(1, 1);
    rate<128, false, 2>("f16 TS N=128 K=16 (sw128)");
    rate<128, true, 0>("i8  TS N=128 K=32 (no swizzle)");
    rate<128, true, 1>("i8  TS N=128 K=32 (sw64)");
    rate<128, true, 2>("i8  TS N=128 K=32 (sw128)");
    rate<192, true, 0>("i8  TS N=192 K=32 (no swizzle)");
    rate<192, true, 1>("i8  TS N=192 K=32 (sw64)");
    rate<256, true, 1>("i8  TS N=256 K=32 (sw64)");
    rate<256, false, 2>("f16 TS N=256 K=16 (sw128)");
    sttm_latency<16>(0); sttm_latency<16>(1);
    sttm_latency<32>(0); sttm_latency<32>(1);
    return 0;
}
