// Dense ends of the forward pass on the 5th-gen tensor cores (VERDICT r1 "f2"):
//   h2_dense_tc_f32 :  Y[:, off : off + c] = act(op(A) . op(W) + b)      fp32 in, fp32 out
// replaces keras Dense (h2gcn/models/H2GCN.py:244-257: logits = dropout(final) . W_out, [N, 7p] x [7p, C]) and
// SparseDense on DENSE features (h2gcn/models/_layers.py:45-52: relu(X . W0), [N, F] x [F, p]); with the transpose flags
// also the two classifier-side contractions of the training step (dW_out = final^T . dlogits, dfinal = dlogits . W_out^T).
//
// Arithmetic: 3xTF32 — every fp32 operand is split on the fly into big = its top 19 bits (exactly a TF32 number) and
// small = x - big (exact in fp32; the tensor core truncates it to TF32 again), and
//     D += A_big.B_big + A_small.B_big + A_big.B_small        (`tcgen05.mma.kind::tf32`, fp32 accumulators in TMEM)
// drops only terms of relative size 2^-21.  What remains is the tensor core's round-toward-zero accumulation: one
// truncation per MMA instruction (K = 8), i.e. a bias of ~K/8 * 2^-25 of the result (the cross terms go to a second
// accumulator so that they do not add truncations to the large sum): measured 2e-6 of max-abs at K = 448, 5e-6 at K = 1433
// against an fp64 product, bar 1e-5 (north-star tolerance 1e-4).  The fp32 SIMT kernel h2_dense_f32 stays as the
// reference-order parity mode.
//
// Per CTA (288 threads): 128 rows of the output (UMMA M = 128), N = c padded to a multiple of 16 (tiles of <= 128 columns).
//   warps 0..7 : loaders, two groups of 4 on alternate K blocks, each with its NEXT block's global loads already in flight
//                in registers (4 K blocks of loads outstanding per CTA: with one group and no prefetch the kernel paid one
//                DRAM round trip per K block, 42 us for the Cora classifier under ncu) — — 128-bit coalesced global loads of a [128 x 32] fp32 slab of A (a K block = one 128-byte
//                swizzle row), split, two `st.shared.v4` into the K-major SWIZZLE_128B images of A_big / A_small; the W
//                slab likewise (scalar stores: it is tiny and arrives transposed); `fence.proxy.async` + mbarrier arrive.
//                Afterwards the epilogue: `tcgen05.ld` (lane = row), + bias, ReLU, stores into the concat slot.
//   warp 8     : allocates TMEM; one elected lane issues 4 K steps x 3 MMAs per K block and commits the stage.
// The op is HBM / latency bound (Cora classifier: 4.9 MB in, 76 KB out), so the loaders ARE the pipeline: 3-4 stages.
#include "bm_common.cuh"

namespace h2 {

constexpr int kDtRows = 128;        // UMMA M
constexpr int kDtKBlock = 32;       // fp32 per K block = 128 bytes = one SWIZZLE_128B row
constexpr int kDtGroup = 128;       // loader threads per K block (one group fills one stage)
constexpr int kDtGroups = 2;        // loader groups: group g takes the K blocks kb = g (mod 2)
constexpr int kDtLoaders = kDtGroup * kDtGroups;
constexpr int kDtThreads = kDtLoaders + 32;

struct DenseTcParams {
    const float *A;      // trans_a == 0: [m, k] row-major (lda);  trans_a == 1: [k, m] row-major (lda)
    const float *W;      // trans_w == 0: [k, n] row-major (ldw);  trans_w == 1: [n, k] row-major (ldw)
    const float *bias;   // [n] or nullptr
    float *Y;            // [m, *] row-major (ldy), column offset applied by the host
    int64_t lda, ldw, ldy;
    int32_t m, n, k;
    int32_t trans_a, trans_w, relu;
};

__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// byte offset of fp32 element (row r, k) inside a K-major SWIZZLE_128B tile (rows of 128 bytes, 8-row atoms of 1 KB)
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) << 4) | ((k & 3) << 2)));
}

__device__ __forceinline__ void split_tf32(float x, float &big, float &small) {
    big = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);   // sign, exponent, 10 mantissa bits: exactly a TF32 value
    small = x - big;                                           // exact
}

template <int NT>     // accumulator columns of one pass: 16, 32, 64 or 128
__global__ void __launch_bounds__(kDtThreads, 1) dense_tc_kernel(const __grid_constant__ DenseTcParams p) {
    constexpr int kStages = NT <= 64 ? 4 : 3;
    constexpr uint32_t kABytes = kDtRows * 128, kBBytes = NT * 128;
    constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;
    // two accumulators: [0, NT) the big x big products, [NT, 2 NT) the two cross terms.  The tensor core truncates (RZ)
    // once per instruction; with one accumulator that is 3 truncations per K step of a ~|result|-sized sum (measured
    // 1.6e-5 of max-abs at K = 1433), with the cross terms apart 1 — their own sum is 2^-10 of it and its truncation is noise
    constexpr uint32_t kTmemCols = 2 * NT < 32 ? 32 : 2 * NT;
    // D fp32 | A, B TF32 | K-major both | N | M = 128
    constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | ((128u >> 4) << 24);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    __shared__ uint64_t s_bar[2 * kStages + 1];
    __shared__ uint32_t s_tmem_base;
    const uint32_t bar_full = smem_u32(&s_bar[0]), bar_empty = bar_full + 8 * kStages, bar_acc = bar_full + 16 * kStages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kDtRows, n0 = blockIdx.y * NT;
    const int n_kb = (p.k + kDtKBlock - 1) / kDtKBlock;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, kDtGroup);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kDtLoaders / 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == kDtLoaders / 32) {
        // ===== MMA issuer =====
        if (elect_one()) {
            for (int kb = 0; kb < n_kb; ++kb) {
                const uint32_t s = kb % kStages;
                mbar_wait(bar_full + 8 * s, (kb / kStages) & 1);
                tc_fence_after();
                const uint32_t a_big = smem_base + s * kStageBytes, a_small = a_big + kABytes;
                const uint32_t b_big = a_small + kABytes, b_small = b_big + kBBytes;
#pragma unroll
                for (int kk = 0; kk < kDtKBlock / 8; ++kk) {   // K = 8 fp32 = 32 bytes per instruction
                    const uint32_t o = kk * 32;
                    umma_tf32_ss(tmem_base + NT, umma_desc_sw128(a_small + o), umma_desc_sw128(b_big + o), kIdesc, (kb | kk) ? 1u : 0u);
                    umma_tf32_ss(tmem_base + NT, umma_desc_sw128(a_big + o), umma_desc_sw128(b_small + o), kIdesc, 1u);
                    umma_tf32_ss(tmem_base, umma_desc_sw128(a_big + o), umma_desc_sw128(b_big + o), kIdesc, (kb | kk) ? 1u : 0u);
                }
                umma_commit(bar_empty + 8 * s);
            }
            umma_commit(bar_acc);
        }
        __syncwarp();
    } else {
        // ===== loaders =====
        const int grp = threadIdx.x / kDtGroup, t = threadIdx.x % kDtGroup;
        // the [128 x 32] fp32 slab of K block kb as 8 float4 per thread (8 lanes cover the 128 bytes of a row: coalesced)
        auto load_a = [&](int kb, float4 (&va)[8]) {
            const int k0 = kb * kDtKBlock;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = t + i * kDtGroup, r = idx >> 3, c4 = (idx & 7) * 4;
                const int64_t gr = m0 + r;
                va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (gr < p.m && kb < n_kb) {
                    const float *src = p.A + gr * p.lda + k0 + c4;
                    if (k0 + c4 + 4 <= p.k && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) va[i] = *reinterpret_cast<const float4 *>(src);
                    else {
                        if (k0 + c4 + 0 < p.k) va[i].x = src[0];
                        if (k0 + c4 + 1 < p.k) va[i].y = src[1];
                        if (k0 + c4 + 2 < p.k) va[i].z = src[2];
                        if (k0 + c4 + 3 < p.k) va[i].w = src[3];
                    }
                }
            }
        };
        float4 va[2][8];
        if (!p.trans_a) load_a(grp, va[0]);
        int buf = 0;
        for (int kb = grp; kb < n_kb; kb += kDtGroups, buf ^= 1) {
            const uint32_t s = kb % kStages;
            const int k0 = kb * kDtKBlock;
            uint8_t *a_big = smem_gen + s * kStageBytes, *a_small = a_big + kABytes;
            uint8_t *b_big = a_small + kABytes, *b_small = b_big + kBBytes;
            if (!p.trans_a) {                    // this group's NEXT block: its loads fly while the current one is split
                if (buf == 0) load_a(kb + kDtGroups, va[1]); else load_a(kb + kDtGroups, va[0]);
            }
            mbar_wait(bar_empty + 8 * s, ((kb / kStages) & 1) ^ 1);
            if (!p.trans_a) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int idx = t + i * kDtGroup, r = idx >> 3, c4 = (idx & 7) * 4;
                    const float4 v = buf == 0 ? va[0][i] : va[1][i];
                    float4 b, sm;
                    split_tf32(v.x, b.x, sm.x); split_tf32(v.y, b.y, sm.y);
                    split_tf32(v.z, b.z, sm.z); split_tf32(v.w, b.w, sm.w);
                    const uint32_t off = sw128_off(r, c4);
                    *reinterpret_cast<float4 *>(a_big + off) = b;
                    *reinterpret_cast<float4 *>(a_small + off) = sm;
                }
            } else {
                // A given as [k, m]: element (row r, k) = A[(k0 + k) * lda + m0 + r]; consecutive threads read consecutive m
                for (int idx = t; idx < kDtKBlock * kDtRows; idx += kDtGroup) {
                    const int kq = idx / kDtRows, r = idx % kDtRows;
                    float v = 0.f;
                    if (k0 + kq < p.k && m0 + r < p.m) v = p.A[(int64_t)(k0 + kq) * p.lda + m0 + r];
                    float b, sm;
                    split_tf32(v, b, sm);
                    const uint32_t off = sw128_off(r, kq);
                    *reinterpret_cast<float *>(a_big + off) = b;
                    *reinterpret_cast<float *>(a_small + off) = sm;
                }
            }
            // W slab: B[n][k] (K-major) = W[k0 + k][n0 + n] (or W[n0 + n][k0 + k] when it is given transposed)
            for (int idx = t; idx < kDtKBlock * NT; idx += kDtGroup) {
                int kq, nn;
                if (!p.trans_w) { kq = idx / NT; nn = idx % NT; } else { nn = idx / kDtKBlock; kq = idx % kDtKBlock; }
                float v = 0.f;
                if (k0 + kq < p.k && n0 + nn < p.n)
                    v = p.trans_w ? __ldg(p.W + (int64_t)(n0 + nn) * p.ldw + k0 + kq) : __ldg(p.W + (int64_t)(k0 + kq) * p.ldw + n0 + nn);
                float b, sm;
                split_tf32(v, b, sm);
                const uint32_t off = sw128_off(nn, kq);
                *reinterpret_cast<float *>(b_big + off) = b;
                *reinterpret_cast<float *>(b_small + off) = sm;
            }
            fence_proxy_async_smem();        // generic-proxy stores -> visible to the tensor core's async proxy
            mbar_arrive(bar_full + 8 * s);
        }
        // ===== epilogue: lane = row =====
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int quarter = warp & 3, chalf = warp >> 2;     // TMEM lane quarter of this warp, and which 16-column blocks it takes
        const int64_t gr = m0 + quarter * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int c0 = 16 * chalf; c0 < NT; c0 += 16 * kDtGroups) {
            uint32_t acc[16], acc2[16];
            cuda::ptx::tcgen05_ld_32x32b(acc, t_lane + c0);
            cuda::ptx::tcgen05_ld_32x32b(acc2, t_lane + NT + c0);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (gr < p.m) {
                float *dst = p.Y + gr * p.ldy + n0 + c0;
                float o[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float v = __uint_as_float(acc[e]) + __uint_as_float(acc2[e]);
                    if (p.bias && n0 + c0 + e < p.n) v += __ldg(p.bias + n0 + c0 + e);
                    o[e] = p.relu ? fmaxf(v, 0.f) : v;
                }
                if (n0 + c0 + 16 <= p.n && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4 *>(dst + e) = make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (n0 + c0 + e < p.n) dst[e] = o[e];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kDtLoaders / 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

template <int NT>
static int dense_tc_launch(const DenseTcParams &p, cudaStream_t st) {
    constexpr int kStages = NT <= 64 ? 4 : 3;
    constexpr size_t smem = (size_t)kStages * (2 * kDtRows * 128 + 2 * NT * 128) + 1024;
    auto kern = dense_tc_kernel<NT>;
    static bool done[64] = {};
    int dev = 0;
    H2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !done[dev]) {
        H2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    dim3 grid((unsigned)((p.m + kDtRows - 1) / kDtRows), (unsigned)((p.n + NT - 1) / NT));
    kern<<<grid, kDtThreads, smem, st>>>(p);
    H2_LAUNCHED("dense_tc_kernel");
    return H2_OK;
}

}  // namespace h2

using namespace h2;

extern "C" int h2_dense_tc_f32(int32_t m, int32_t k, int32_t n, const float *A, int64_t lda, int32_t trans_a, const float *W,
                               int64_t ldw, int32_t trans_w, const float *bias, int32_t relu, float *Y, int64_t ldy,
                               int64_t out_col_off, h2_stream_t s) {
    H2_REQUIRE(m >= 0 && k >= 1 && n >= 1, H2_ERR_INVALID, "h2_dense_tc_f32: m=%d k=%d n=%d", m, k, n);
    if (m == 0) return H2_OK;
    H2_REQUIRE(A && W && Y && out_col_off >= 0 && ldy >= out_col_off + n && lda >= (trans_a ? m : k) && ldw >= (trans_w ? k : n),
               H2_ERR_INVALID, "h2_dense_tc_f32: null pointer or leading dimension too small");
    DenseTcParams p;
    p.A = A; p.W = W; p.bias = bias; p.Y = Y + out_col_off;
    p.lda = lda; p.ldw = ldw; p.ldy = ldy;
    p.m = m; p.n = n; p.k = k; p.trans_a = trans_a ? 1 : 0; p.trans_w = trans_w ? 1 : 0; p.relu = relu ? 1 : 0;
    cudaStream_t st = (cudaStream_t)s;
    if (n <= 16) return dense_tc_launch<16>(p, st);
    if (n <= 32) return dense_tc_launch<32>(p, st);
    if (n <= 64) return dense_tc_launch<64>(p, st);
    return dense_tc_launch<128>(p, st);
}
