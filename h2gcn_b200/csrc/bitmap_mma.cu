// Dense-ish hop adjacency on the 5th-gen tensor cores.
//
// The exact-2-hop pattern P2 of a graph with average degree ~40 is 10-15 % dense at |V| = 10 k: a CSR gather has to
// pull nnz x 512 B through L2 (7.6 GB per round, profiles/r01a) although X itself is only 5 MB.  P2 is BINARY, so it
// is exactly representable in bf16; with the symmetric normalisation factored out,
//     Y = diag(dinv) . P . (diag(dinv) . X),
// the product P . X' is a dense contraction over a 0/1 matrix that `tcgen05.mma` evaluates exactly, given X' split
// into bf16 pieces (hi + lo [+ lo2]: 16 [24] significand bits; the fp32 accumulators live in TMEM).
//
// Format ("tile bitmap"): P is cut into units of 256 rows x 64 columns; only non-empty units are stored, each as
// 256 x uint64 (2 KB, 1 bit per entry instead of a 4-byte column index).  Units are ordered by (row tile, column
// chunk).  A persistent grid of <= 148 CTAs gets equal contiguous ranges of units ("stream-K"); a range that does
// not cover a whole row tile writes an fp32 partial tile that a fix-up kernel adds in a fixed order (deterministic).
//
// Per CTA (320 threads, 1 CTA / SM, all 512 TMEM columns: 2 x 128 accumulator columns + 4 A stages of 64):
//   warps 0..7  : A producers — thread r expands the 64 bits of row r into 64 bf16 (0.0 / 2.0: a single set bit per
//                 element, two ALU ops per 32-bit word; the factor 2 is folded into the epilogue scale) and writes
//                 them with ONE `tcgen05.st.32x32b.x32` into its own TMEM lane (the wait::st + arrive of a unit is
//                 deferred to just before the warp's next store).  Afterwards the same warps are the epilogue:
//                 `tcgen05.ld`, hi + lo, x dinv_row, a per-warp shared-memory stage, TMEM released, coalesced stores.
//   warp 8      : TMA producer — one elected lane; per unit two `cp.async.bulk` (1-D TMA, UBLKCP): the pre-packed,
//                 pre-swizzled X' tile [S*DG rows x 64 k] (bf16, K-major SWIZZLE_128B image) and the unit's 2 KB
//                 bitmap, both completing on the stage's mbarrier (8-stage ring).
//   warp 9      : allocates TMEM; ONE elected lane (`elect.sync`) runs the whole issue loop: per unit 2 x 4
//                 `tcgen05.mma.cta_group::1.kind::f16` (M=128, N=S*DG, K=16; A from TENSOR MEMORY, B from shared
//                 memory), `tcgen05.commit` to free the stages / publish the accumulator.  It is the LAST warp
//                 because the issue arbiter favours the highest warp id over the polling producer warps.
// Work items are (column group, unit) pairs, group-major, cut into <= 148 equal contiguous ranges (stream-K); a range
// that does not cover a whole (row tile, group) writes an fp32 partial tile and `bm_fixup_kernel` adds the partials
// of a tile in ascending slot order (deterministic).  Measurements behind these choices: profiles/README.md.
//
// Arithmetics (`splits`, include/h2gcn_b200.h).  The kernel in THIS file (bm_mma_kernel, single CTA, kind::f16) serves the
// 2 / 3 bf16-piece arithmetics.  The DEFAULT is int8 (H2_SPLITS_I8X3 / I8X2): X' as block-fixed-point — one step for the
// matrix, an exponent 2^t (t in 0..6) per ROW of X' carried by the 0/1 operand as bytes 0 / 2^t, 3 / 2 balanced base-256
// digits as the int8 B operand — on `tcgen05.mma.cta_group::2.kind::i8` (K = 32, twice the bf16 MAC rate, EXACT int32
// accumulation; the epilogue converts and scales once): bm_pair_kernel in bm_pair.cu.  This file also holds what both
// share: the tile-bitmap format and its construction, the stream-K schedules, and the operand packing — bm_pack_kernel
// (bf16 pieces) and the cooperative bm_pack_i8_kernel (B tiles of [S*DG x 64] bytes, SWIZZLE_64B, + 128 bytes of per-word
// {rotate, mask} constants; bitmaps in bit order 1 so that an operand word is one rotate + one mask).
#include "bm_common.cuh"

namespace h2 {

// ------------------------------------------------------------------------------------------------------------------
// format construction
// ------------------------------------------------------------------------------------------------------------------
__global__ void bm_flag_kernel(int32_t n_rows, int32_t n_chunks, const int64_t *__restrict__ rowptr,
                               const int32_t *__restrict__ col, int64_t *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    const int64_t base = (row / kTileRows) * n_chunks;
    for (int64_t k = s + lane; k < e; k += 32) flags[base + col[k] / kChunkCols] = 1;
}

__global__ void bm_fill_kernel(int32_t n_rows, int32_t n_chunks, const int64_t *__restrict__ rowptr,
                               const int32_t *__restrict__ col, const int64_t *__restrict__ unit_index,
                               int32_t *__restrict__ unit_chunk, unsigned long long *__restrict__ bits, int bit_order) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    const int64_t base = (row / kTileRows) * n_chunks;
    const int r = (int)(row % kTileRows);
    for (int64_t k = s + lane; k < e; k += 32) {
        const int c = col[k];
        const int chunk = c / kChunkCols;
        const int64_t u = unit_index[base + chunk];
        atomicOr(&bits[u * kTileRows + r], 1ull << bm_bit_pos(c % kChunkCols, bit_order));
        unit_chunk[u] = chunk;  // same value from every writer
    }
}

__global__ void bm_tile_ptr_kernel(int32_t n_tiles, int32_t n_chunks, const int64_t *__restrict__ unit_index,
                                   int64_t *__restrict__ tile_ptr) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= n_tiles) tile_ptr[t] = unit_index[(int64_t)t * n_chunks];
}

// ------------------------------------------------------------------------------------------------------------------
// X' packing: fp32 X (row-major, ld) -> per 64-row chunk a [S*DG x 64] bf16 K-major tile, 128-byte swizzled, i.e.
// byte-for-byte the shared-memory image the UMMA B descriptor expects.  n = s*DG + f, k = j % 64.
// ------------------------------------------------------------------------------------------------------------------
// Row source of the round input: rows [bound[q], bound[q+1]) live at ptr[q] (row-major, leading dimension ld).  With one
// part this is a plain matrix; with several parts the pointers are the ranks' row shards mapped over NVLink (peer /
// symmetric memory), i.e. the hop-boundary all-gather is fused into this kernel's loads.
struct PackSrc {
    const float *ptr[8];
    int32_t bound[9];
    int32_t n_parts;
    int32_t bf16;      // rows hold bf16 (ld counts bf16 elements); int8 pack only, single part
    int64_t ld;
};

__device__ __forceinline__ const float *pack_src_row(const PackSrc &src, int j) {
    int q = 0;
#pragma unroll
    for (int t = 1; t < 8; ++t) q += (t < src.n_parts && j >= src.bound[t]) ? 1 : 0;
    if (src.bf16) return reinterpret_cast<const float *>(reinterpret_cast<const uint16_t *>(src.ptr[q]) + (int64_t)(j - src.bound[q]) * src.ld);
    return src.ptr[q] + (int64_t)(j - src.bound[q]) * src.ld;
}
// 4 consecutive features starting at feature f of a source row (fp32: 16 bytes, bf16: 8 bytes widened)
__device__ __forceinline__ float4 pack_src_load4(const PackSrc &src, const float *row, int f) {
    if (!src.bf16) return *reinterpret_cast<const float4 *>(row + f);
    const uint2 u = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint16_t *>(row) + f);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xFFFF0000u));
}

template <int DG>
__global__ void __launch_bounds__(256) bm_pack_kernel(int32_t n_cols, int32_t d, int32_t n_groups, int32_t splits,
                                                      const __grid_constant__ PackSrc src,
                                                      const float *__restrict__ dinv, uint4 *__restrict__ out,
                                                      float *__restrict__ xfull, int64_t ld_full) {
    __shared__ float s_x[kChunkCols][DG + 1];
    const int chunk = blockIdx.x, g = blockIdx.y;
    const int j0 = chunk * kChunkCols, f0 = g * DG;
    // coalesced 128-bit loads of the [64 x DG] slab, scaled by dinv[j]
    constexpr int V = DG / 4;
    for (int idx = threadIdx.x; idx < kChunkCols * V; idx += blockDim.x) {
        const int k = idx / V, f4 = (idx % V) * 4;
        const int j = j0 + k;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j < n_cols && f0 + f4 < d) {   // d % 4 == 0: a float4 is either fully inside or fully outside
            v = *reinterpret_cast<const float4 *>(pack_src_row(src, j) + f0 + f4);   // plain load: may be peer memory
            if (xfull) *reinterpret_cast<float4 *>(xfull + (int64_t)j * ld_full + f0 + f4) = v;   // gathered fp32 copy
            const float sc = dinv ? __ldg(dinv + j) : 1.f;
            v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
        }
        s_x[k][f4] = v.x; s_x[k][f4 + 1] = v.y; s_x[k][f4 + 2] = v.z; s_x[k][f4 + 3] = v.w;
    }
    __syncthreads();
    const int n_rows_b = splits * DG;
    uint4 *tile = out + ((int64_t)chunk * n_groups + g) * (n_rows_b * 8);  // 8 x 16 B per n-row
    // thread -> (feature f, 16-byte k-chunk c16); all `splits` pieces of an element are produced together
    for (int idx = threadIdx.x; idx < DG * 8; idx += blockDim.x) {
        const int f = idx >> 3, c16 = idx & 7;
        uint32_t w[3][4];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float v = s_x[c16 * 8 + e][f];
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                if (t < splits) {
                    const __nv_bfloat16 b = __float2bfloat16_rn(v);
                    v -= __bfloat162float(b);
                    const uint32_t h = (uint32_t)__bfloat16_as_ushort(b);
                    if (e & 1) w[t][e >> 1] |= h << 16; else w[t][e >> 1] = h;
                }
            }
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            if (t < splits) {
                const int n = t * DG + f;
                const int dst = (n >> 3) * 64 + (n & 7) * 8 + (c16 ^ (n & 7));  // in 16-byte units, 1024-byte atoms
                tile[dst] = make_uint4(w[t][0], w[t][1], w[t][2], w[t][3]);
            }
        }
    }
}

// fp32 gather of the row shards into one matrix (rounds whose hops are all CSR)
__global__ void gather_rows_kernel(int32_t n_cols, int32_t d4, const __grid_constant__ PackSrc src, float *__restrict__ xfull,
                                   int64_t ld_full) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_cols * d4) return;
    const int j = (int)(idx / d4), c = (int)(idx % d4);
    reinterpret_cast<float4 *>(xfull + (int64_t)j * ld_full)[c] = reinterpret_cast<const float4 *>(pack_src_row(src, j))[c];
}

// the gathered copy of a row-sharded input keeps the row type of the shards (fp32, or bf16 with ld_full counting bf16
// elements): element offset `off`, 4 consecutive features
__device__ __forceinline__ float4 xfull_load4(const float *xfull, int64_t off, int bf16) {
    if (!bf16) return *reinterpret_cast<const float4 *>(xfull + off);
    const uint2 u = *reinterpret_cast<const uint2 *>(reinterpret_cast<const uint16_t *>(xfull) + off);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xFFFF0000u));
}
__device__ __forceinline__ void xfull_store4(float *xfull, int64_t off, float4 v, int bf16) {
    if (!bf16) { *reinterpret_cast<float4 *>(xfull + off) = v; return; }
    // the values came from bf16 rows: the upper halves are the original bits (exact)
    const uint32_t lo = (__float_as_uint(v.x) >> 16) | (__float_as_uint(v.y) & 0xFFFF0000u);
    const uint32_t hi = (__float_as_uint(v.z) >> 16) | (__float_as_uint(v.w) & 0xFFFF0000u);
    *reinterpret_cast<uint2 *>(reinterpret_cast<uint16_t *>(xfull) + off) = make_uint2(lo, hi);
}

// ------------------------------------------------------------------------------------------------------------------
// int8 operand (kind::i8): X' = diag(dinv) X as block-fixed-point.
//   x'[j][c] ~= step * 2^t(j) * q[j][c],   q an integer of S balanced base-256 digits (|q| <= i8_range(S)),
//   step = 2^(e_max - 5) / i8_range(S) for the whole matrix (e_max = exponent of max |x'|), t = exponent of ROW j in 0..6
//   (rows more than 64x below the maximum keep t = 0 and lose one bit per factor 2 below that: with 3 digits a row
//   10^4 below the maximum still carries 16 bits).  Round 1 used one exponent per 4 rows: a small row sharing a group
//   with a large one lost up to 6 more bits, visible in a ROW-wise error metric (tests: row_wise_relative_error).
// The 0/1 pattern operand carries 2^t (the MMA kernels expand bit -> byte 0 / 2^t), the digits are the int8 B operand,
// the int32 accumulation is exact, and the epilogue applies step * dinv_row once.
//
// ONE cooperative launch (bm_pack_i8_kernel; round 1 used two kernels, 8.5 + 9.3 us): phase 1 loads a [64 rows x 128
// features] slab per work item (128-bit loads, possibly from PEER memory: the hop-boundary all-gather of a row-sharded
// round, which also leaves the gathered fp32 copy), scales it by dinv and keeps it in shared memory, and reduces max |x'|
// per row and per CTA (no atomics on values: deterministic); a grid barrier (arrive / depart counters in the
// buffer header, self-resetting) makes the global maximum known; phase 2 derives the block exponents, quantises the slab
// still sitting in shared memory and writes, per 64-row chunk and column group, the K-major SWIZZLE_64B image of the
// [S*DG x 64] int8 B tile followed by the 16 {rotate, mask} pairs the A producers need for that chunk.  X is read once.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int f32_exponent(float x) { return (int)((__float_as_uint(x) >> 23) & 0xFFu) - 127; }

constexpr int kPackSlab = 128;     // features per work item of bm_pack_i8_kernel (measured r02: 64-feature items make the
                                   // kernel 11.4 instead of 13.6 us, -2 us on the cold round, nothing on the pipelined rate; not adopted)
constexpr int kPackThreads = 256;

struct I8Header {                  // first kI8HeaderBytes of the packed operand
    float step;                    // quantisation step of X' (read by the MMA kernels' epilogues)
    uint32_t pad[3];
    uint32_t arrive, depart;       // grid barrier of the pack kernel; zero between launches
};

template <int DG, int S>
__global__ void __launch_bounds__(kPackThreads) bm_pack_i8_kernel(int32_t n_cols, int32_t d, int32_t n_groups, int32_t n_slabs,
                                                                  const __grid_constant__ PackSrc src, const float *__restrict__ dinv,
                                                                  float *rowmax, float *blockmax, uint8_t *__restrict__ xpack,
                                                                  float *__restrict__ xfull, int64_t ld_full) {
    constexpr int NB = S * DG;
    constexpr int kTileBytes = NB * 64 + kI8ConstBytes;
    constexpr int V = kPackSlab / 4;
    __shared__ float s_x[kChunkCols][kPackSlab + 1];
    __shared__ float s_red[8];
    __shared__ int s_t[kChunkCols];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = (n_cols + kChunkCols - 1) / kChunkCols;
    const int n_items = n_chunks * n_slabs;
    I8Header *hdr = reinterpret_cast<I8Header *>(xpack);

    auto load_slab = [&](int chunk, int slab, bool first_pass) {
        const int j0 = chunk * kChunkCols, f0 = slab * kPackSlab;
        constexpr int kPer = kChunkCols * V / kPackThreads;    // 8 float4 per thread: ALL loads are issued before the first use
        float4 v[kPer];
        float sc[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int idx = threadIdx.x + i * kPackThreads;
            const int k = idx / V, f4 = (idx % V) * 4;
            const int j = j0 + k;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            sc[i] = 0.f;
            if (j < n_cols && f0 + f4 < d) {   // d % 4 == 0: a float4 is either fully inside or fully outside
                if (first_pass || !xfull) v[i] = pack_src_load4(src, pack_src_row(src, j), f0 + f4);   // plain load: may be peer memory
                else v[i] = xfull_load4(xfull, (int64_t)j * ld_full + f0 + f4, src.bf16);                // second pass: the local copy
                sc[i] = dinv ? __ldg(dinv + j) : 1.f;
            }
        }
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            const int idx = threadIdx.x + i * kPackThreads;
            const int k = idx / V, f4 = (idx % V) * 4;
            const int j = j0 + k;
            if (xfull && first_pass && j < n_cols && f0 + f4 < d)
                xfull_store4(xfull, (int64_t)j * ld_full + f0 + f4, v[i], src.bf16);          // gathered copy, in the row type
            s_x[k][f4] = v[i].x * sc[i]; s_x[k][f4 + 1] = v[i].y * sc[i]; s_x[k][f4 + 2] = v[i].z * sc[i]; s_x[k][f4 + 3] = v[i].w * sc[i];
        }
    };

    // ---- phase 1: maxima ----
    float cta_max = 0.f;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int chunk = item / n_slabs, slab = item % n_slabs;
        if (item != (int)blockIdx.x) __syncthreads();          // the previous item's slab has been reduced
        load_slab(chunk, slab, true);
        __syncthreads();
#pragma unroll 1
        for (int ri = 0; ri < kChunkCols / 8; ++ri) {          // warp w: rows 8w .. 8w + 7 of the chunk
            const int k = 8 * warp + ri;
            float m = 0.f;
#pragma unroll
            for (int c = 0; c < kPackSlab; c += 32) {
                // non-finite inputs must not poison the one global step: NaN drops out of fmaxf, +-Inf is skipped here
                // and saturates below (documented divergence: the fp32 CSR path propagates them like the reference)
                const float a = fabsf(s_x[k][c + lane]);
                m = fmaxf(m, a <= 3.4e38f ? a : 0.f);
            }
            m = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m)));   // non-negative floats order like their bits
            const int j = chunk * kChunkCols + k;
            if (lane == 0 && j < n_cols) rowmax[(int64_t)slab * n_cols + j] = m;
            cta_max = fmaxf(cta_max, m);
        }
    }
    if (lane == 0) s_red[warp] = cta_max;
    __syncthreads();
    if (threadIdx.x == 0) {
        float b = s_red[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) b = fmaxf(b, s_red[w]);
        blockmax[blockIdx.x] = b;
        // ---- grid barrier (cooperative launch: every CTA is resident) ----
        __threadfence();
        atomicAdd(&hdr->arrive, 1u);
        while (*reinterpret_cast<volatile uint32_t *>(&hdr->arrive) < gridDim.x) __nanosleep(32);
        __threadfence();
    }
    __syncthreads();

    // ---- phase 2: global maximum, block exponents, digits ----
    float gm = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kPackThreads) gm = fmaxf(gm, __ldcg(blockmax + i));
    gm = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(gm)));
    __syncthreads();                                           // s_red is reused
    if (lane == 0) s_red[warp] = gm;
    __syncthreads();
    gm = s_red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) gm = fmaxf(gm, s_red[w]);
    const int eg = gm > 0.f ? max(f32_exponent(gm), -96) : -96;
    if (blockIdx.x == 0 && threadIdx.x == 0) hdr->step = ldexpf(1.f, eg - 5) / (float)i8_range(S);
    const float mult = ldexpf((float)i8_range(S), 5 - eg);
    constexpr int kTilesPerSlab = kPackSlab / DG;              // column groups (B tiles) one slab feeds
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int chunk = item / n_slabs, slab = item % n_slabs;
        if (n_items > (int)gridDim.x) {                        // several items per CTA: the slab is no longer in shared memory
            __syncthreads();
            load_slab(chunk, slab, false);
        }
        if (threadIdx.x < kChunkCols) {                        // block exponent of every row of the chunk
            const int j = chunk * kChunkCols + threadIdx.x;
            float mg = 0.f;
            if (j < n_cols)
                for (int sl = 0; sl < n_slabs; ++sl) mg = fmaxf(mg, __ldcg(rowmax + (int64_t)sl * n_cols + j));
            s_t[threadIdx.x] = mg > 0.f ? min(max(f32_exponent(mg) - eg + kI8Levels, 0), kI8Levels) : 0;
        }
        __syncthreads();
        if (threadIdx.x < 16) {
            // word j of an A row = columns 4j..4j+3 = bits 8b + (j % 8) of half j / 8 (bm_bit_pos order 1).  The A
            // producers rotate right by j % 8 (bit b -> bit 0 of byte b), spread every bit over its byte (x 0xFF) and keep
            // bit t(4j + b) of byte b: the mask below.  Every B tile of the chunk carries a copy.
            const int j = threadIdx.x;
            const uint32_t mask = (1u << s_t[4 * j]) | (1u << (8 + s_t[4 * j + 1])) | (1u << (16 + s_t[4 * j + 2])) | (1u << (24 + s_t[4 * j + 3]));
            for (int tl = 0; tl < kTilesPerSlab; ++tl) {
                const int g = slab * kTilesPerSlab + tl;
                if (g < n_groups)
                    *reinterpret_cast<uint2 *>(xpack + kI8HeaderBytes + ((int64_t)chunk * n_groups + g) * kTileBytes + NB * 64 + j * 8) =
                        make_uint2((uint32_t)(j & 7), mask);
            }
        }
        // thread = (32-feature slice, feature fl, 16 consecutive k): S x 16 bytes per slice visit
        const int fl = threadIdx.x & 31, c16 = (threadIdx.x >> 5) & 3, half = threadIdx.x >> 7;
#pragma unroll 1
        for (int sl = half; sl < kPackSlab / 32; sl += 2) {
            const int f = sl * 32 + fl;                         // feature inside the slab
            const int g = (slab * kPackSlab + f) / DG;          // column group (B tile)
            if (g >= n_groups) continue;
            uint8_t *tile = xpack + kI8HeaderBytes + ((int64_t)chunk * n_groups + g) * kTileBytes;
            const int fg = (slab * kPackSlab + f) % DG;         // feature inside the group
            uint32_t w[S][4];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int k = c16 * 16 + e;
                const float sc = __int_as_float((127 - s_t[k]) << 23);   // 2^-t of row k
                float tq = s_x[k][f] * (mult * sc);
                tq = tq != tq ? 0.f : fminf(fmaxf(tq, -(float)i8_range(S)), (float)i8_range(S));   // NaN -> 0, +-Inf saturates
                int q = __float2int_rn(tq);
#pragma unroll
                for (int t = S - 1; t >= 0; --t) {        // piece 0 = most significant digit
                    const int dig = ((q + 128) & 255) - 128;
                    q = (q - dig) >> 8;
                    const uint32_t b = (uint32_t)dig & 0xFFu;
                    if (e & 3) w[t][e >> 2] |= b << (8 * (e & 3)); else w[t][e >> 2] = b;
                }
            }
#pragma unroll
            for (int t = 0; t < S; ++t) {
                const int n = t * DG + fg;
                const int off = (n >> 3) * 512 + (n & 7) * 64 + ((c16 ^ ((n >> 1) & 3)) << 4);   // SWIZZLE_64B, 512-byte atoms
                *reinterpret_cast<uint4 *>(tile + off) = make_uint4(w[t][0], w[t][1], w[t][2], w[t][3]);
            }
        }
    }
    // ---- depart: the last CTA re-arms the barrier for the next launch ----
    if (threadIdx.x == 0) {
        if (atomicAdd(&hdr->depart, 1u) == gridDim.x - 1) {
            hdr->arrive = 0;
            hdr->depart = 0;
            __threadfence();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------------------------------
#ifdef H2_BM_TRACE
__device__ long long g_bm_trace[148 * 32];
__device__ long long g_bm_trace2[8 * 64];   // CTA 0 only, per unit (first 64): see tools/dbg_run.py
#define BM_TRACE(slot) do { if (slot < 26) g_bm_trace[blockIdx.x * 32 + (slot)] = clock64(); } while (0)
#define BM_T2(row, unit) do { if (blockIdx.x == 0 && (unit) < 64u) g_bm_trace2[(row) * 64 + (unit)] = clock64(); } while (0)
// accumulated wait cycles: slots 26 MMA thread on full_a, 27/28/29 producer warp 0 on full_b / empty_a / wait::st,
// 30 TMA thread on empty_b, 31 units issued
// (kept in registers, stored once when the role ends)
#define BM_ACC_DECL() long long bm_acc__[3] = {0, 0, 0}
#define BM_WAIT_BEGIN() const long long bm_w0__ = clock64()
#define BM_WAIT_END(k) bm_acc__[k] += clock64() - bm_w0__
#define BM_COUNT(k) bm_acc__[k] += 1
#define BM_ACC_STORE(k, slot) g_bm_trace[blockIdx.x * 32 + (slot)] = bm_acc__[k]
#else
#define BM_TRACE(slot) do { } while (0)
#define BM_T2(row, unit) do { } while (0)
#define BM_ACC_DECL() do { } while (0)
#define BM_WAIT_BEGIN() do { } while (0)
#define BM_WAIT_END(k) do { } while (0)
#define BM_COUNT(k) do { } while (0)
#define BM_ACC_STORE(k, slot) do { } while (0)
#endif
struct BmParams {
    const int32_t *unit_chunk;
    const unsigned long long *bits;
    const BmSegment *seg;
    const int32_t *cta_seg_ptr;  // [n_ctas + 1]
    const uint4 *xpack;          // [n_chunks][n_groups] B tiles: [S*DG x 64] bf16, or int8 + 128 constant bytes
    const float *xstep;          // int8 path: quantisation step of X' (device scalar written by the pack kernel)
    const float *dinv_row;       // [n_rows] (local rows) or nullptr
    float *Y;                    // + out_col_off applied by the host
    float *partial;              // [n_partial_slots][256][DG]
    int64_t ldy;
    int32_t n_rows, d, n_groups, splits;
};

// bf16 pieces: kind::f16, K = 16, A stage = 2 x 32 TMEM columns, B tile rows of 128 bytes, SWIZZLE_128B.  (Round 1 also
// instantiated this kernel for the int8 digits; since r02 every int8 round runs on bm_pair_kernel, bm_pair.cu, and the
// int8 branches were removed here.)
template <int DG, int S>
struct BmCfg {
    static constexpr int NB = S * DG;                                             // UMMA N
    static constexpr uint32_t kBBytes = NB * 128;                                 // bytes per B tile in global memory
    static constexpr uint32_t kBStride = (kBBytes + 1023u) & ~1023u;              // shared-memory stage stride
    static constexpr uint32_t kAccCols = 2 * NB;                                  // two 128-row halves
    static constexpr uint32_t kAHalfCols = 32;                                    // TMEM columns of one half of an A stage
    static constexpr int kAStg = kAStages;
    static constexpr size_t kSmem = (size_t)kBStages * (kBStride + kTileRows * 8) + 8 * 32 * (DG + 4) * 4 + 1024;
    static_assert(kAccCols + kAStg * 2 * kAHalfCols <= 512 && NB % 16 == 0 && NB >= 16 && NB <= 256, "UMMA N / TMEM budget");
    static_assert(kAStg <= kAStagesMax && 8 % kAStg == 0, "A stage ring");
};

template <int DG, int S>
__global__ void __launch_bounds__(bm_threads(false), 1) bm_mma_kernel(const __grid_constant__ BmParams p) {
    using Cfg = BmCfg<DG, S>;
    constexpr int kProducerSets = bm_producer_sets(false);
    constexpr int NB = Cfg::NB;
    constexpr uint32_t kBBytes = Cfg::kBBytes, kBStride = Cfg::kBStride;
    constexpr uint32_t kAccCols = Cfg::kAccCols;
    constexpr uint32_t kACol0 = kAccCols;           // A stages follow the accumulators
    constexpr uint32_t kAHalfCols = Cfg::kAHalfCols, kAStageCols = 2 * kAHalfCols;
    constexpr int kAStg = Cfg::kAStg;
    constexpr uint32_t kTmemCols = 512;
    // instruction descriptor: D format (fp32 = 1 / int32 = 2) | A, B format (bf16 = 1 / signed int8 = 1) | N >> 3 | M >> 4
    constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NB >> 3) << 17) | ((128u >> 4) << 24);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // Shared-window addresses are made opaque to the compiler: left alone it REMATERIALISES them inside the unit loops
    // from S2R SR_CgaCtaId (a long-scoreboard read, measured as 30 % of the A producers' time) instead of keeping them
    // in a register.
    uint32_t smem_raw_u32 = smem_u32(smem_raw);
    asm volatile("" : "+r"(smem_raw_u32));
    const uint32_t smem_base = (smem_raw_u32 + 1023u) & ~1023u;   // B stages, 1024-byte aligned (swizzle atoms)
    const uint32_t bits_base = smem_base + kBStages * kBStride;          // + kBStages x 2 KB unit bitmaps
    const unsigned long long *bits_gen = reinterpret_cast<const unsigned long long *>(smem_raw + (bits_base - smem_raw_u32));
    constexpr int kStageStride = DG + 4;   // floats; +4 keeps 16-byte alignment and spreads rows over the banks
    float *stage_gen = reinterpret_cast<float *>(smem_raw + (bits_base - smem_raw_u32) + kBStages * kTileRows * 8);
    __shared__ uint64_t s_bar[2 * kAStg + 2 * kBStages + 2];
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_chunk[32];
    uint32_t bar0 = smem_u32(&s_bar[0]);
    asm volatile("" : "+r"(bar0));
    const uint32_t bar_full_a = bar0;
    const uint32_t bar_empty_a = bar0 + 8 * kAStg;
    const uint32_t bar_full_b = bar0 + 8 * (2 * kAStg);
    const uint32_t bar_empty_b = bar0 + 8 * (2 * kAStg + kBStages);
    const uint32_t bar_acc_full = bar0 + 8 * (2 * kAStg + 2 * kBStages);
    const uint32_t bar_acc_empty = bar0 + 8 * (2 * kAStg + 2 * kBStages + 1);

    // The issue arbiter favours the highest warp id of a sub-partition: the MMA issuer must not queue behind the
    // (busy-polling) producer warps, so it is the LAST warp of the CTA.
    constexpr int kTmaWarp = 8 * kProducerSets, kMmaWarp = 8 * kProducerSets + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int seg_begin = p.cta_seg_ptr[blockIdx.x], seg_end = p.cta_seg_ptr[blockIdx.x + 1];
    const int n_work = seg_end - seg_begin;

    if (threadIdx.x == 0) {
        BM_TRACE(0);
        for (int s = 0; s < kAStg; ++s) {
            mbar_init(bar_full_a + 8 * s, 8);             // one arrive per A-producer warp of the unit
            mbar_init(bar_empty_a + 8 * s, 1);   // tcgen05.commit
        }
        for (int s = 0; s < kBStages; ++s) {
            mbar_init(bar_full_b + 8 * s, 1);    // arrive.expect_tx by the TMA thread
            mbar_init(bar_empty_b + 8 * s, 1);   // tcgen05.commit
        }
        mbar_init(bar_acc_full, 1);
        mbar_init(bar_acc_empty, 8 * kProducerSets);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == kTmaWarp) {
        // ===== TMA producer (B tiles + unit bitmaps) =====
        // The column chunk of every unit comes from global memory: the WHOLE warp fetches 32 indices at a time (one
        // coalesced load, the next block prefetched while this one is issued) and parks them in shared memory; one
        // elected lane then issues the copies of those 32 units back to back.  (A single lane loading one index per
        // unit exposes the load latency whenever it does not have to wait for a free stage: measured ~380 cycles per
        // unit, more than the 256 cycles of int8 MMA work.)
        uint32_t it = 0;
        BM_ACC_DECL();
        for (int w = 0; w < n_work; ++w) {
            const BmSegment sg = p.seg[seg_begin + w];
            const int g = sg.group;
            int nxt = sg.unit_begin + lane < sg.unit_end ? p.unit_chunk[sg.unit_begin + lane] : 0;
            for (int u0 = sg.unit_begin; u0 < sg.unit_end; u0 += 32) {
                s_chunk[lane] = nxt;
                __syncwarp();
                if (u0 + 32 + lane < sg.unit_end) nxt = p.unit_chunk[u0 + 32 + lane];
                const int cnt = min(32, sg.unit_end - u0);
                if (elect_one()) {
                    for (int k = 0; k < cnt; ++k) {
                        const int chunk = s_chunk[k];
                        const uint32_t st = (it + k) % kBStages, ph = ((it + k) / kBStages) & 1;
                        { BM_WAIT_BEGIN(); mbar_wait(bar_empty_b + 8 * st, ph ^ 1); BM_WAIT_END(0); }
                        BM_T2(0, it + k);
                        mbar_arrive_expect_tx(bar_full_b + 8 * st, kBBytes + kTileRows * 8);
                        const uint4 *src = p.xpack + ((int64_t)chunk * p.n_groups + g) * (kBBytes / 16);
                        bulk_copy_g2s(smem_base + st * kBStride, src, kBBytes, bar_full_b + 8 * st);
                        bulk_copy_g2s(bits_base + st * (kTileRows * 8), p.bits + (int64_t)(u0 + k) * kTileRows, kTileRows * 8,
                                      bar_full_b + 8 * st);
                    }
                }
                it += cnt;
                __syncwarp();
            }
        }
        if (lane == 0) BM_ACC_STORE(0, 30);
    } else if (warp == kMmaWarp) {
        // ===== MMA issuer: ONE elected lane runs the whole loop (no per-unit reconvergence).  A single thread executes
        // ~one dependent instruction per 5-10 cycles, so the per-unit instruction count is what bounds the issue rate:
        // the stage index is made a compile-time constant (8-way unrolled dispatch on `it & 7`) so every descriptor,
        // TMEM address and barrier address is loop-invariant-base + immediate. =====
        if (elect_one()) {
            uint32_t it = 0, acc_it = 0;
            BM_ACC_DECL();
            auto issue_unit = [&](auto stage_c, uint32_t acc_first) {
                constexpr uint32_t ST = decltype(stage_c)::value;      // it & 7
                constexpr uint32_t sa = ST % kAStg, sb = ST % kBStages;
                static_assert(8 % kAStg == 0 && 8 % kBStages == 0, "stage rings must divide the unroll factor");
                // full_a implies full_b: the A producers read the unit's bitmap out of the same B stage
                { BM_WAIT_BEGIN(); mbar_wait(bar_full_a + 8 * sa, (it / kAStg) & 1); BM_WAIT_END(0); BM_COUNT(1); }
                BM_T2(5, it);
                tc_fence_after();
                const uint32_t b0 = smem_base + sb * kBStride;
                const uint32_t a0 = tmem_base + kACol0 + sa * kAStageCols;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int k = 0; k < kChunkCols / 16; ++k)   // 32 operand bytes (8 TMEM columns) per step
                        umma_bf16_ts(tmem_base + half * NB, a0 + half * kAHalfCols + k * 8, umma_desc_sw128(b0 + k * 32), kIdesc,
                                     k > 0 ? 1u : acc_first);
                }
                umma_commit(bar_empty_a + 8 * sa);   // both arrive once the MMAs above have consumed their operands
                umma_commit(bar_empty_b + 8 * sb);
                BM_T2(6, it);
                ++it;
            };
            for (int w = 0; w < n_work; ++w) {
                const BmSegment sg = p.seg[seg_begin + w];
                for (int once = 0; once < 1; ++once, ++acc_it) {
                    BM_TRACE(2 + 6 * w);
                    mbar_wait(bar_acc_empty, (acc_it & 1) ^ 1);   // epilogue of the previous accumulator has drained TMEM
                    tc_fence_after();
                    BM_TRACE(3 + 6 * w);
                    int left = sg.unit_end - sg.unit_begin;
                    uint32_t acc = 0;   // the first unit of a segment overwrites the accumulator
                    while (left > 0) {
                        switch (it & 7u) {   // falls through: consecutive units use consecutive stages
                            case 0: issue_unit(std::integral_constant<uint32_t, 0>{}, acc); acc = 1; if (--left == 0) break;
                            case 1: issue_unit(std::integral_constant<uint32_t, 1>{}, acc); acc = 1; if (--left == 0) break;
                            case 2: issue_unit(std::integral_constant<uint32_t, 2>{}, acc); acc = 1; if (--left == 0) break;
                            case 3: issue_unit(std::integral_constant<uint32_t, 3>{}, acc); acc = 1; if (--left == 0) break;
                            case 4: issue_unit(std::integral_constant<uint32_t, 4>{}, acc); acc = 1; if (--left == 0) break;
                            case 5: issue_unit(std::integral_constant<uint32_t, 5>{}, acc); acc = 1; if (--left == 0) break;
                            case 6: issue_unit(std::integral_constant<uint32_t, 6>{}, acc); acc = 1; if (--left == 0) break;
                            default: issue_unit(std::integral_constant<uint32_t, 7>{}, acc); acc = 1; --left;
                        }
                    }
                    umma_commit(bar_acc_full);
                    BM_TRACE(4 + 6 * w);
                }
            }
            BM_ACC_STORE(0, 26);
            BM_ACC_STORE(1, 31);
        }
        __syncwarp();
    } else {
        // ===== A producers, then epilogue =====
        // Two sets of 8 warps take alternate units: a warp's per-unit chain (barrier wake-up, LDS, bit expansion,
        // tcgen05.st, wait::st, arrive) is ~1000 cycles of latency against 512 cycles of MMA work per unit, and the
        // wait::st + arrive of a unit is deferred until just before the warp's next tcgen05.st.
        const int pw = warp;                              // 0..15
        const int set = pw >> 3;
        const int half = (pw >> 2) & 1, quarter = warp & 3;   // TMEM lane quarter this warp may touch
        const int r = half * 128 + quarter * 32 + lane;   // row inside the 256-row tile
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        uint32_t it = 0, acc_it = 0;
        BM_ACC_DECL();
        for (int w = 0; w < n_work; ++w) {
            const BmSegment sg = p.seg[seg_begin + w];
            const int g = sg.group;
            for (int once = 0; once < 1; ++once, ++acc_it) {
                bool pending = false;
                uint32_t pending_sa = 0;
                // sets of 8 warps (half x quarter) take alternate units
                constexpr int kStep = kProducerSets;
                const int mine = set;
                const int skip = (int)((uint32_t)(mine + kStep - (int)(it % kStep)) % kStep);
                const uint32_t it_end = it + (uint32_t)(sg.unit_end - sg.unit_begin);
                it += skip;
                // one unit: expand the bitmap row(s) into `a`, publish the previous unit, wait for a free A stage, store
                auto produce = [&](int u, uint32_t (&a)[32]) {
                    const uint32_t sb = it % kBStages, pb = (it / kBStages) & 1;
                    // the unit's bitmap rides in the B stage (TMA-prefetched)
                    { BM_WAIT_BEGIN(); mbar_wait(bar_full_b + 8 * sb, pb); BM_WAIT_END(0); }
                    if ((warp & 3) == 0 && lane == 0) BM_T2(1, it);
                    {
                        const unsigned long long bits = bits_gen[sb * kTileRows + r];
                        // element k of the row -> bf16 2.0 (0x4000) or 0: word j = (bit 2j) << 14 | (bit 2j+1) << 30
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint32_t byte = (uint32_t)(bits >> (8 * c)) & 0xFFu;
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                a[4 * c + i] = (byte * ((1u << (14 - 2 * i)) | (1u << (29 - 2 * i)))) & 0x40004000u;
                        }
                    }
                    const uint32_t sa = it % kAStg, pa = (it / kAStg) & 1;
                    if (pending) {   // publish the previous unit BEFORE waiting for a free stage (it may be the same stage)
                        BM_WAIT_BEGIN();
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_full_a + 8 * pending_sa);
                        BM_WAIT_END(2);
                    }
                    { BM_WAIT_BEGIN(); mbar_wait(bar_empty_a + 8 * sa, pa ^ 1); BM_WAIT_END(1); }
                    if ((warp & 3) == 0 && lane == 0) BM_T2(3, it);
                    tc_fence_after();
                    cuda::ptx::tcgen05_st_32x32b(t_lane + kACol0 + sa * kAStageCols + half * kAHalfCols, a);
                    if ((warp & 3) == 0 && lane == 0) BM_T2(4, it);
                    pending = true;
                    pending_sa = sa;
                };
                {
                    // (measured dead end: two alternating register arrays, so that a tcgen05.st never has its source
                    // registers overwritten by the next unit, change nothing)
                    uint32_t a[32];
                    for (int u = sg.unit_begin + skip; u < sg.unit_end; u += kStep, it += kStep) produce(u, a);
                }
                it = it_end;
                if (pending) {
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full_a + 8 * pending_sa);
                }
                // ---- epilogue for (segment, group): set s takes the 32-column blocks s, s+2, ... ----
                if (warp == 0 && lane == 0) BM_TRACE(5 + 6 * w);
                mbar_wait(bar_acc_full, acc_it & 1);
                tc_fence_after();
                if (warp == 0 && lane == 0) BM_TRACE(6 + 6 * w);
                const int64_t grow = (int64_t)sg.tile * kTileRows + r;
                const bool row_ok = grow < p.n_rows;
                // A holds 2.0, not 1.0
                const float scale = 0.5f * ((row_ok && p.dinv_row) ? p.dinv_row[grow] : 1.f);
                const int valid_cols = sg.partial_slot < 0 ? min(DG, p.d - g * DG) : DG;
                const uint32_t t_row = t_lane + half * NB;
                // TMEM -> registers (lane = row) -> per-warp shared-memory stage; the accumulator is released as soon as
                // it has been read, and the stage is then written out with fully coalesced 128-bit stores (a row of the
                // tile is contiguous across lanes) — per-lane row stores cost one 16-byte sector write per lane.
                float *stage = stage_gen + (warp & 7) * (32 * kStageStride);
#pragma unroll 1
                for (int c0 = set * 32; c0 < DG; c0 += 32 * kProducerSets) {
                    uint32_t acc[S][32];
#pragma unroll
                    for (int s = 0; s < S; ++s) cuda::ptx::tcgen05_ld_32x32b(acc[s], t_row + s * DG + c0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int q = 0; q < 32; q += 4) {
                        float4 o;
                        float *po = &o.x;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float v = __uint_as_float(acc[S - 1][q + e]);
#pragma unroll
                            for (int s = S - 2; s >= 0; --s) v += __uint_as_float(acc[s][q + e]);  // small pieces first
                            po[e] = v * scale;
                        }
                        *reinterpret_cast<float4 *>(stage + lane * kStageStride + c0 + q) = o;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (warp == 0 && lane == 0) BM_TRACE(7 + 6 * w);
                if (lane == 0) mbar_arrive(bar_acc_empty);   // TMEM drained: the next segment's MMAs may start
                {
                    constexpr int kLanesPerRow = DG / 4, kRowsPerInstr = 32 / kLanesPerRow;
                    const int rr = lane / kLanesPerRow, c = (lane % kLanesPerRow) * 4;
                    const int row0 = half * 128 + quarter * 32;
                    const bool col_ok = c + 4 <= valid_cols && (kProducerSets == 1 || (c / 32) % kProducerSets == set);
#pragma unroll
                    for (int j = 0; j < 32; j += kRowsPerInstr) {
                        const int row = j + rr;
                        const float4 v = *reinterpret_cast<const float4 *>(stage + row * kStageStride + c);
                        const int64_t gr = (int64_t)sg.tile * kTileRows + row0 + row;
                        float *drow = sg.partial_slot < 0 ? p.Y + gr * p.ldy + (int64_t)g * DG
                                                           : p.partial + ((int64_t)sg.partial_slot * kTileRows + row0 + row) * DG;
                        if (col_ok && (sg.partial_slot >= 0 || gr < p.n_rows)) *reinterpret_cast<float4 *>(drow + c) = v;
                    }
                }
                __syncwarp();   // the stage is reused by the next epilogue
            }
        }
        if (warp == 0 && lane == 0) { BM_ACC_STORE(0, 27); BM_ACC_STORE(1, 28); BM_ACC_STORE(2, 29); }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) BM_TRACE(1);
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// fix-up: Y[tile rows, group columns] = sum over the partial slots of that (tile, group), ascending slot order
// (deterministic).  One thread per output float4; grid = (256 * DG/4 / 256, n_fix).
template <int DG>
__global__ void __launch_bounds__(256) bm_fixup_kernel(const BmFix *__restrict__ fix, int32_t n_rows, int32_t d,
                                                       const float *__restrict__ partial, float *__restrict__ Y,
                                                       int64_t ldy) {
    const BmFix f = fix[blockIdx.y];
    constexpr int V = DG / 4;
    const int idx = blockIdx.x * 256 + threadIdx.x;
    const int r = idx / V, c4 = idx % V;
    const int64_t grow = (int64_t)f.tile * kTileRows + r;
    if (r >= kTileRows || grow >= n_rows || f.group * DG + c4 * 4 + 4 > d) return;
    const float4 *src = reinterpret_cast<const float4 *>(partial + ((int64_t)f.slot_begin * kTileRows + r) * DG + c4 * 4);
    constexpr int64_t kSlotStride = (int64_t)kTileRows * DG / 4;   // in float4
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    const int n = f.slot_end - f.slot_begin;
    int s = 0;
    for (; s + 4 <= n; s += 4) {   // 4 loads in flight, added in slot order
        const float4 t0 = __ldcs(src + (s + 0) * kSlotStride), t1 = __ldcs(src + (s + 1) * kSlotStride);
        const float4 t2 = __ldcs(src + (s + 2) * kSlotStride), t3 = __ldcs(src + (s + 3) * kSlotStride);
        a.x += t0.x; a.y += t0.y; a.z += t0.z; a.w += t0.w;
        a.x += t1.x; a.y += t1.y; a.z += t1.z; a.w += t1.w;
        a.x += t2.x; a.y += t2.y; a.z += t2.z; a.w += t2.w;
        a.x += t3.x; a.y += t3.y; a.z += t3.z; a.w += t3.w;
    }
    for (; s < n; ++s) {
        const float4 t = __ldcs(src + s * kSlotStride);
        a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    *reinterpret_cast<float4 *>(Y + grow * ldy + (int64_t)f.group * DG + c4 * 4) = a;
}

// rows of tiles that own no unit at all must still be written as zeros (TF zero-initialises its output)
__global__ void bm_zero_tiles_kernel(const int32_t *__restrict__ tiles, int32_t n_rows, int32_t d, float *__restrict__ Y,
                                     int64_t ldy) {
    const int tile = tiles[blockIdx.x];
    for (int idx = threadIdx.x; idx < kTileRows * (d / 4); idx += blockDim.x) {
        const int r = idx / (d / 4), c4 = idx % (d / 4);
        const int64_t grow = (int64_t)tile * kTileRows + r;
        if (grow < n_rows) *reinterpret_cast<float4 *>(Y + grow * ldy + c4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

static size_t align_up_sz(size_t x, size_t a) { return (x + a - 1) / a * a; }
// column-group width: S*DG is the UMMA N (<= 256), two accumulators of S*DG columns must fit the 512 TMEM columns
// column groups are rounded up to a power of two (1, 2, 4, 8): one schedule per count; padding groups compute zeros
static int groups_for(int d, int dg) { int g = (d + dg - 1) / dg, p2 = 1; while (p2 < g) p2 <<= 1; return p2; }
// int8: at least the two FH-feature halves of one pair-kernel group (a zero tile for d <= 32)
static int groups_for_splits(int d, int dg, int splits) { const int g = groups_for(d, dg); return splits_i8(splits) ? std::max(2, g) : g; }
// features per packed B tile: 32 for 3 bf16 pieces and narrow rounds; int8 rounds use 32 up to d = 64 (the two halves of
// the pair kernel's 64-feature group) and 64 above
// (H2_BM_FH32_UPTO = widest int8 round that still uses 32-feature tiles: measurement knob, default 64)
static int i8_fh32_upto() { static const int v = [] { const char *e = getenv("H2_BM_FH32_UPTO"); return e ? atoi(e) : 64; }(); return v; }
static int dg_for(int d, int splits) { return (d <= 32 || splits == 3 || (splits_i8(splits) && d <= i8_fh32_upto())) ? 32 : 64; }
static size_t i8_tile_bytes(int dg, int splits) { return (size_t)splits_pieces(splits) * dg * 64 + kI8ConstBytes; }
struct I8Layout {   // int8 xpack buffer: header | tiles | rowmax [n_slabs][n_cols] | blockmax [kPackMaxCtas]
    int64_t n_chunks, n_groups, n_slabs;
    size_t off_gmax4, off_blockmax, total;
};
constexpr int kPackMaxCtas = kNumSms * 8;
static I8Layout i8_layout(int32_t n_cols, int32_t d, int32_t splits) {
    I8Layout L;
    const int dg = dg_for(d, splits);
    L.n_chunks = ((int64_t)n_cols + kChunkCols - 1) / kChunkCols;
    L.n_groups = groups_for_splits(d, dg, splits);
    L.n_slabs = ((int64_t)d + kPackSlab - 1) / kPackSlab;
    L.off_gmax4 = align_up_sz(kI8HeaderBytes + (size_t)(L.n_chunks * L.n_groups) * i8_tile_bytes(dg, splits), 256);
    L.off_blockmax = L.off_gmax4 + align_up_sz((size_t)(L.n_slabs * n_cols) * 4, 256);
    L.total = L.off_blockmax + align_up_sz((size_t)kPackMaxCtas * 4, 256) + 256;
    return L;
}

// ---- CTA-pair int8 kernel (bm_pair.cu) ----
struct BmPairSeg { int32_t tile, unit_begin, unit_end, group, n_slots, slot, fix, slot_begin; };
struct BmPairFix { int32_t tile, group, slot_begin, n_slots; };
struct BmPairParams {
    const int32_t *unit_chunk; const unsigned long long *bits; const BmPairSeg *seg; const int32_t *cta_seg_ptr;
    const uint8_t *xpack; const float *xstep; const float *dinv_row; float *Y; float *partial; const BmPairFix *fix;
    uint32_t *sync; int32_t *status; int32_t n_fix; int64_t ldy; int32_t n_rows, d, n_groups_fh, y_bf16, safe_handover;
};
void pair_schedule(const std::vector<int64_t> &tp, int64_t n_units, int ng, int n_pairs_max, std::vector<BmPairSeg> &segs,
                   std::vector<int32_t> &pair_ptr, std::vector<BmPairFix> &fixes, int *n_slots_out);
int pair_launch(int S, int fh, int n_pairs, const BmPairParams &p, cudaStream_t st);
#ifdef H2_BM_PAIR_SEG_UNITS
constexpr int kPairMaxSegUnitsHost = H2_BM_PAIR_SEG_UNITS;
#else
constexpr int kPairMaxSegUnitsHost = 2048;
#endif
static size_t pair_sched_bytes(int64_t nt, int64_t n_units, int ng) {
    const size_t max_segs = (size_t)(kNumSms / 2 + ng * (nt + 1) + ng * (n_units / kPairMaxSegUnitsHost + 1) + 2);
    return align_up_sz(max_segs * sizeof(BmPairSeg), 256) + align_up_sz((size_t)(kNumSms / 2 + 2) * 4, 256) +
           align_up_sz(max_segs * 8, 256) + align_up_sz(max_segs * sizeof(BmPairFix), 256);   // segments, pair_ptr, arrival counters, fix list
}
// every int8 round runs on the pair kernel: FH = 32 (d <= 64; a round of <= 32 features computes a zero-padded second
// half) or FH = 64 features per CTA half, 1 / 2 / 4 column groups of 2*FH features
static bool pair_applies(int32_t splits, int d) { return splits_i8(splits) && d > 0; }
static int pair_fh(int d) { return d <= i8_fh32_upto() ? 32 : 64; }
static int pair_groups(int d) { return std::max(2, groups_for(d, pair_fh(d))) / 2; }

}  // namespace h2

using namespace h2;

extern "C" size_t h2_bm_host_bytes(void) { return sizeof(BmHost); }

// widest d one h2_bm_pack_x_f32 / h2_bm_spmm_f32 pair covers (8 column groups); wider rounds go in column slices
extern "C" int32_t h2_bm_max_width(int32_t splits) { return splits_valid(splits) ? 8 * dg_for(512, splits) : 0; }

extern "C" size_t h2_bm_index_bytes(int32_t n_rows, int32_t n_cols) {
    const int64_t nt = ((int64_t)n_rows + kTileRows - 1) / kTileRows, nc = ((int64_t)n_cols + kChunkCols - 1) / kChunkCols;
    // flags [nt*nc] + unit_index [nt*nc + 1] + tile_ptr [nt + 1] (all int64) + scan workspace
    return (size_t)(2 * nt * nc + nt + 4) * 8 + h2_scan_workspace_bytes(nt * nc) + 512;
}

// Phase 1 (SYNCHRONISES): marks the non-empty 256x64 units, numbers them, returns their count.
extern "C" int h2_bm_count(int32_t n_rows, int32_t n_cols, const int64_t *rowptr, const int32_t *col, void *index_ws,
                           size_t index_ws_bytes, int64_t *n_units_host, h2_stream_t s) {
    cudaStream_t st = (cudaStream_t)s;
    H2_REQUIRE(n_rows > 0 && n_cols > 0 && rowptr && index_ws && n_units_host, H2_ERR_INVALID, "h2_bm_count: bad argument");
    H2_REQUIRE(index_ws_bytes >= h2_bm_index_bytes(n_rows, n_cols), H2_ERR_WORKSPACE, "h2_bm_count: workspace too small");
    const int64_t nt = ((int64_t)n_rows + kTileRows - 1) / kTileRows, nc = ((int64_t)n_cols + kChunkCols - 1) / kChunkCols;
    H2_REQUIRE(nt * nc < (1ll << 28), H2_ERR_UNSUPPORTED, "h2_bm_count: %lld x %lld units is too many for the bitmap format",
               (long long)nt, (long long)nc);
    int64_t *flags = (int64_t *)index_ws;
    int64_t *unit_index = flags + nt * nc;
    void *scan_ws = (void *)(unit_index + nt * nc + 1 + nt + 1 + 2);
    H2_CUDA(cudaMemsetAsync(flags, 0, (size_t)nt * nc * 8, st));
    bm_flag_kernel<<<(unsigned)(((int64_t)n_rows * 32 + 255) / 256), 256, 0, st>>>(n_rows, (int)nc, rowptr, col, flags);
    H2_LAUNCHED("bm_flag_kernel");
    int rc = h2_exclusive_scan_i64(nt * nc, flags, unit_index, scan_ws, h2_scan_workspace_bytes(nt * nc), s);
    if (rc != H2_OK) return rc;
    H2_CUDA(cudaMemcpyAsync(n_units_host, unit_index + nt * nc, 8, cudaMemcpyDeviceToHost, st));
    H2_CUDA(cudaStreamSynchronize(st));
    return H2_OK;
}

static size_t sched_bytes(int64_t nt, int ng) {
    return align_up_sz((size_t)(kNumSms + ng * (nt + 1) + 2) * sizeof(BmSegment), 256) + align_up_sz((size_t)(kNumSms + 2) * 4, 256) +
           align_up_sz((size_t)(ng * (nt + 1)) * sizeof(BmFix), 256);
}

extern "C" size_t h2_bm_plan_dev_bytes(int32_t n_rows, int32_t n_cols, int64_t n_units) {
    const int64_t nt = ((int64_t)n_rows + kTileRows - 1) / kTileRows;
    size_t b = align_up_sz((size_t)n_units * 4, 256) + align_up_sz((size_t)n_units * kTileRows * 8, 256) +
               align_up_sz((size_t)(nt + 1) * 4, 256) + 1024;
    for (int k = 0; k < kNumScheds; ++k) b += sched_bytes(nt, 1 << k);
    for (int k = 0; k < kNumPairScheds; ++k) b += pair_sched_bytes(nt, n_units, 1 << k);
    return b + 256;   // + status word
}

// Phase 2 (SYNCHRONISES): fills the bitmaps and builds the stream-K schedule (segments, partial slots, fix-ups).
extern "C" int h2_bm_fill(int32_t n_rows, int32_t n_cols, const int64_t *rowptr, const int32_t *col, void *index_ws,
                          int64_t n_units, void *bm_host, void *bm_dev, size_t bm_dev_bytes, h2_stream_t s) {
    return h2_bm_fill_order(n_rows, n_cols, rowptr, col, index_ws, n_units, bm_host, bm_dev, bm_dev_bytes, 0, s);
}

extern "C" int h2_bm_fill_order(int32_t n_rows, int32_t n_cols, const int64_t *rowptr, const int32_t *col, void *index_ws,
                                int64_t n_units, void *bm_host, void *bm_dev, size_t bm_dev_bytes, int32_t bit_order,
                                h2_stream_t s) {
    cudaStream_t st = (cudaStream_t)s;
    H2_REQUIRE(n_rows > 0 && n_cols > 0 && rowptr && index_ws && bm_host && bm_dev && n_units >= 0 &&
               (bit_order == 0 || bit_order == 1), H2_ERR_INVALID, "h2_bm_fill: bad argument");
    H2_REQUIRE(bm_dev_bytes >= h2_bm_plan_dev_bytes(n_rows, n_cols, n_units), H2_ERR_WORKSPACE, "h2_bm_fill: plan buffer too small");
    H2_REQUIRE(n_units < 0x7fffffffLL, H2_ERR_UNSUPPORTED, "h2_bm_fill: too many units");
    const int64_t nt = ((int64_t)n_rows + kTileRows - 1) / kTileRows, nc = ((int64_t)n_cols + kChunkCols - 1) / kChunkCols;
    int64_t *flags = (int64_t *)index_ws;
    int64_t *unit_index = flags + nt * nc;
    int64_t *tile_ptr_dev = unit_index + nt * nc + 1;
    BmHost *h = (BmHost *)bm_host;
    memset(h, 0, sizeof(*h));
    h->magic = kBmMagic;
    h->n_rows = n_rows; h->n_cols = n_cols; h->n_tiles = (int)nt; h->n_chunks = (int)nc; h->n_units = n_units;
    h->bit_order = bit_order;
    size_t off = 0;
    h->off_unit_chunk = off; off += align_up_sz((size_t)n_units * 4, 256);
    h->off_bits = off;       off += align_up_sz((size_t)n_units * kTileRows * 8, 256);
    h->off_empty_tiles = off; off += align_up_sz((size_t)(nt + 1) * 4, 256);
    char *base = (char *)bm_dev;
    H2_CUDA(cudaMemsetAsync(base + h->off_bits, 0, (size_t)n_units * kTileRows * 8, st));
    if (n_units > 0) {
        bm_fill_kernel<<<(unsigned)(((int64_t)n_rows * 32 + 255) / 256), 256, 0, st>>>(
            n_rows, (int)nc, rowptr, col, unit_index, (int32_t *)(base + h->off_unit_chunk),
            (unsigned long long *)(base + h->off_bits), bit_order);
        H2_LAUNCHED("bm_fill_kernel");
    }
    bm_tile_ptr_kernel<<<(unsigned)((nt + 1 + 255) / 256), 256, 0, st>>>((int)nt, (int)nc, unit_index, tile_ptr_dev);
    H2_LAUNCHED("bm_tile_ptr_kernel");
    std::vector<int64_t> tp(nt + 1);
    H2_CUDA(cudaMemcpyAsync(tp.data(), tile_ptr_dev, (size_t)(nt + 1) * 8, cudaMemcpyDeviceToHost, st));
    int64_t nnz = 0;
    H2_CUDA(cudaMemcpyAsync(&nnz, rowptr + n_rows, 8, cudaMemcpyDeviceToHost, st));
    H2_CUDA(cudaStreamSynchronize(st));
    h->nnz = nnz;
    // ---- stream-K schedules on the host: for ng column groups the work items (group, unit) are linearised group-major
    // and cut into G equal contiguous ranges; a range is split at (group, tile) boundaries into segments -------------
    std::vector<int32_t> empty_tiles;
    for (int64_t q = 0; q < nt; ++q)
        if (tp[q + 1] == tp[q]) empty_tiles.push_back((int32_t)q);
    h->n_empty_tiles = (int64_t)empty_tiles.size();
    if (!empty_tiles.empty()) H2_CUDA(cudaMemcpyAsync(base + h->off_empty_tiles, empty_tiles.data(), empty_tiles.size() * 4, cudaMemcpyHostToDevice, st));
    std::vector<BmSegment> segs[kNumScheds];
    std::vector<int32_t> cta_ptr[kNumScheds];
    std::vector<BmFix> fixes[kNumScheds];
    for (int k = 0; k < kNumScheds; ++k) {
        const int ng = 1 << k;
        const int64_t total = n_units * ng;
        const int G = (int)std::min<int64_t>(kNumSms, total);
        BmSched &sc = h->sched[k];
        cta_ptr[k].assign(G + 1, 0);
        int n_slots = 0;
        // Every (group, tile) item a CTA enters costs an epilogue (TMEM drain + write-out + pipeline refill, measured
        // ~3 k cycles ~ 6 units of 512 cycles): ranges are cut at equal COST = units + kEpilogueUnits per item entered,
        // so CTAs whose range crosses a tile boundary get fewer units.
        constexpr int64_t kEpilogueUnits = 6;
        std::vector<int64_t> item_start;
        for (int64_t grp = 0; grp < ng; ++grp)
            for (int64_t t = 0; t < nt; ++t)
                if (tp[t + 1] > tp[t]) item_start.push_back(grp * n_units + tp[t]);
        auto cost_at = [&](int64_t q) {   // cost of linear units [0, q)
            return q + kEpilogueUnits * (int64_t)(std::lower_bound(item_start.begin(), item_start.end(), q) - item_start.begin());
        };
        const int64_t cost_total = cost_at(total);
        auto bound_for = [&](int c) {
            if (c <= 0) return (int64_t)0;
            if (c >= G) return total;
            const int64_t target = cost_total * c / G;
            int64_t lo = 0, hi = total;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (cost_at(mid) < target) lo = mid + 1; else hi = mid;
            }
            return lo;
        };
        for (int c = 0; c < G; ++c) {
            const int64_t q0 = bound_for(c), q1 = bound_for(c + 1);
            cta_ptr[k][c] = (int)segs[k].size();
            int64_t q = q0;
            while (q < q1) {
                const int64_t grp = q / n_units, u = q % n_units;
                const int64_t t = (int64_t)(std::upper_bound(tp.begin(), tp.end(), u) - tp.begin()) - 1;   // tile of unit u
                const int64_t e = std::min<int64_t>(std::min<int64_t>(tp[t + 1], n_units), u + (q1 - q));
                const bool whole = (u == tp[t] && e == tp[t + 1]);
                segs[k].push_back(BmSegment{(int32_t)t, (int32_t)u, (int32_t)e, whole ? -1 : n_slots, (int32_t)grp, 0});
                if (!whole) {
                    if (fixes[k].empty() || fixes[k].back().tile != (int32_t)t || fixes[k].back().group != (int32_t)grp)
                        fixes[k].push_back(BmFix{(int32_t)t, n_slots, n_slots, (int32_t)grp});
                    fixes[k].back().slot_end = ++n_slots;
                }
                q += e - u;
            }
        }
        cta_ptr[k][G] = (int)segs[k].size();
        H2_REQUIRE(segs[k].size() <= (size_t)(kNumSms + ng * (nt + 1) + 2) && fixes[k].size() <= (size_t)(ng * (nt + 1)),
                   H2_ERR_UNSUPPORTED, "h2_bm_fill: schedule table overflow");
        sc.n_ctas = G; sc.n_partial_slots = n_slots; sc.n_fix = (int)fixes[k].size();
        sc.off_seg = off;         off += align_up_sz((size_t)(kNumSms + ng * (nt + 1) + 2) * sizeof(BmSegment), 256);
        sc.off_cta_seg_ptr = off; off += align_up_sz((size_t)(kNumSms + 2) * 4, 256);
        sc.off_fix = off;         off += align_up_sz((size_t)(ng * (nt + 1)) * sizeof(BmFix), 256);
        if (!segs[k].empty()) H2_CUDA(cudaMemcpyAsync(base + sc.off_seg, segs[k].data(), segs[k].size() * sizeof(BmSegment), cudaMemcpyHostToDevice, st));
        H2_CUDA(cudaMemcpyAsync(base + sc.off_cta_seg_ptr, cta_ptr[k].data(), cta_ptr[k].size() * 4, cudaMemcpyHostToDevice, st));
        if (!fixes[k].empty()) H2_CUDA(cudaMemcpyAsync(base + sc.off_fix, fixes[k].data(), fixes[k].size() * sizeof(BmFix), cudaMemcpyHostToDevice, st));
    }
    // ---- schedules of the CTA-pair kernel (bm_pair.cu): one range per PAIR of CTAs, roles instead of a fix-up pass ----
    std::vector<BmPairSeg> psegs[kNumPairScheds];
    std::vector<int32_t> pptr[kNumPairScheds];
    std::vector<BmPairFix> pfix[kNumPairScheds];
    const int max_pairs = kNumSms / 2;
    for (int k = 0; k < kNumPairScheds; ++k) {
        const int ng = 1 << k;
        int n_slots = 0;
        pair_schedule(tp, n_units, ng, max_pairs, psegs[k], pptr[k], pfix[k], &n_slots);
        const size_t max_segs = (size_t)(kNumSms / 2 + ng * (nt + 1) + ng * (n_units / kPairMaxSegUnitsHost + 1) + 2);
        H2_REQUIRE(psegs[k].size() <= max_segs && pfix[k].size() <= max_segs, H2_ERR_UNSUPPORTED, "h2_bm_fill: pair schedule table overflow");
        BmSched &sc = h->sched_pair[k];
        sc.n_ctas = (int)pptr[k].size() - 1; sc.n_partial_slots = n_slots; sc.n_fix = (int)pfix[k].size();
        sc.off_seg = off;         off += align_up_sz(max_segs * sizeof(BmPairSeg), 256);
        sc.off_cta_seg_ptr = off; off += align_up_sz((size_t)(kNumSms / 2 + 2) * 4, 256);
        const size_t cnt_bytes = align_up_sz(max_segs * 8, 256);
        sc.off_fix = off;         off += cnt_bytes + align_up_sz(max_segs * sizeof(BmPairFix), 256);   // arrival counters (zero between launches), then the fix list
        if (!psegs[k].empty()) H2_CUDA(cudaMemcpyAsync(base + sc.off_seg, psegs[k].data(), psegs[k].size() * sizeof(BmPairSeg), cudaMemcpyHostToDevice, st));
        H2_CUDA(cudaMemcpyAsync(base + sc.off_cta_seg_ptr, pptr[k].data(), pptr[k].size() * 4, cudaMemcpyHostToDevice, st));
        H2_CUDA(cudaMemsetAsync(base + sc.off_fix, 0, cnt_bytes, st));
        if (!pfix[k].empty()) H2_CUDA(cudaMemcpyAsync(base + sc.off_fix + cnt_bytes, pfix[k].data(), pfix[k].size() * sizeof(BmPairFix), cudaMemcpyHostToDevice, st));
    }
    h->off_status = off; off += 256;
    H2_CUDA(cudaMemsetAsync(base + h->off_status, 0, 256, st));
    H2_REQUIRE(off <= bm_dev_bytes, H2_ERR_WORKSPACE, "h2_bm_fill: plan buffer too small (%zu > %zu)", off, bm_dev_bytes);
    H2_CUDA(cudaStreamSynchronize(st));   // the std::vectors go out of scope
    return H2_OK;
}

static int sched_index(int n_groups) { return n_groups <= 1 ? 0 : (n_groups == 2 ? 1 : (n_groups <= 4 ? 2 : 3)); }

extern "C" size_t h2_bm_xpack_bytes(int32_t n_cols, int32_t d, int32_t splits) {
    if (!splits_valid(splits) || n_cols <= 0 || d <= 0) return 0;
    if (splits_i8(splits)) return i8_layout(n_cols, d, splits).total;
    const int dg = dg_for(d, splits);
    const int64_t nc = ((int64_t)n_cols + kChunkCols - 1) / kChunkCols, ng = groups_for(d, dg);
    return (size_t)(nc * ng * splits * dg * 128) + 256;
}

extern "C" size_t h2_bm_partial_bytes(const void *bm_host, int32_t d, int32_t splits) {
    const BmHost *h = (const BmHost *)bm_host;
    if (!h || h->magic != kBmMagic || !splits_valid(splits)) return 0;
    if (pair_applies(splits, d)) {
        const int ng = pair_groups(d);
        if (ng > 4) return 0;
        return (size_t)h->sched_pair[sched_index(ng)].n_partial_slots * kTileRows * (2 * pair_fh(d)) * 4 + 256;
    }
    const int dg = dg_for(d, splits);
    const int ng = groups_for(d, dg);
    if (ng > 8) return 0;
    return (size_t)h->sched[sched_index(ng)].n_partial_slots * kTileRows * dg * 4 + 256;
}

namespace h2 {
static int fill_src(PackSrc &src, int32_t n_cols, int32_t n_parts, const float *const *ptrs, const int64_t *bounds, int64_t ld) {
    H2_REQUIRE(n_parts >= 1 && n_parts <= 8 && ptrs && bounds && bounds[0] == 0 && bounds[n_parts] == n_cols, H2_ERR_INVALID,
               "row shards: n_parts=%d (1..8), bounds must run from 0 to n_cols=%d", n_parts, n_cols);
    memset(&src, 0, sizeof(src));
    src.n_parts = n_parts; src.ld = ld;
    for (int q = 0; q < n_parts; ++q) {
        H2_REQUIRE((ptrs[q] || bounds[q + 1] == bounds[q]) && bounds[q + 1] >= bounds[q] && aligned16(ptrs[q]), H2_ERR_INVALID,
                   "row shards: part %d null / unordered / misaligned", q);
        src.ptr[q] = ptrs[q];
        src.bound[q] = (int32_t)bounds[q];
    }
    for (int q = n_parts; q <= 8; ++q) src.bound[q] = n_cols;
    return H2_OK;
}

// pack from row shards (n_parts pointers + bounds); xfull != nullptr also writes the gathered fp32 matrix
int bm_pack_parts(int32_t n_cols, int32_t d, int32_t splits, int32_t n_parts, const float *const *ptrs, const int64_t *bounds,
                  int64_t ld, const float *dinv_col, void *xpack, size_t xpack_bytes, float *xfull, int64_t ld_full,
                  h2_stream_t s, bool zero_header, bool x_bf16) {
    H2_REQUIRE(n_cols > 0 && d > 0 && d % 4 == 0 && splits_valid(splits) && xpack && ld >= d && ld % 4 == 0,
               H2_ERR_INVALID, "bm_pack: bad argument (d=%d splits=%d)", d, splits);
    H2_REQUIRE(!x_bf16 || (splits_i8(splits) && d % 8 == 0 && ld % 8 == 0 && (!xfull || ld_full % 8 == 0)), H2_ERR_UNSUPPORTED,
               "bm_pack: bf16 rows need the int8 digits and d, ld multiples of 8");
    H2_REQUIRE(xpack_bytes >= h2_bm_xpack_bytes(n_cols, d, splits) && aligned16(xpack), H2_ERR_WORKSPACE,
               "bm_pack: xpack buffer too small / misaligned");
    H2_REQUIRE(!xfull || (ld_full >= d && ld_full % 4 == 0 && aligned16(xfull)), H2_ERR_ALIGN, "bm_pack: xfull alignment");
    PackSrc src;
    int rc = fill_src(src, n_cols, n_parts, ptrs, bounds, ld);
    if (rc != H2_OK) return rc;
    src.bf16 = x_bf16 ? 1 : 0;
    const int dg = dg_for(d, splits);
    dim3 grid((unsigned)((n_cols + kChunkCols - 1) / kChunkCols), (unsigned)groups_for_splits(d, dg, splits));
    cudaStream_t st = (cudaStream_t)s;
    if (splits_i8(splits)) {
        // ONE cooperative launch: maxima (+ the gathered fp32 copy), grid barrier, quantisation from shared memory
        const I8Layout L = i8_layout(n_cols, d, splits);
        H2_REQUIRE(grid.y <= 8, H2_ERR_UNSUPPORTED, "bm_pack: d=%d needs %u column groups (max 8)", d, grid.y);
        uint8_t *xp = (uint8_t *)xpack;
        float *gmax4 = (float *)(xp + L.off_gmax4), *blockmax = (float *)(xp + L.off_blockmax);
        if (zero_header) H2_CUDA(cudaMemsetAsync(xp, 0, kI8HeaderBytes, st));   // barrier counters of a buffer seen for the first time
        const int S = splits_pieces(splits);
        const void *kern = dg == 32 ? (S == 2 ? (const void *)bm_pack_i8_kernel<32, 2> : (const void *)bm_pack_i8_kernel<32, 3>)
                                    : (S == 2 ? (const void *)bm_pack_i8_kernel<64, 2> : (const void *)bm_pack_i8_kernel<64, 3>);
        static int max_ctas[4][64] = {};    // co-resident CTAs per instantiation and device (the grid barrier needs them all resident)
        int dev = 0;
        H2_CUDA(cudaGetDevice(&dev));
        int &cap = max_ctas[(dg == 32 ? 0 : 2) + (S == 2 ? 0 : 1)][dev & 63];
        if (!cap) {
            int per_sm = 0, sms = 0;
            H2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPackThreads, 0));
            H2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            cap = std::max(1, std::min(per_sm, 4) * sms);
        }
        const int64_t n_items = L.n_chunks * L.n_slabs;
        const int g = (int)std::min<int64_t>(std::min<int64_t>(n_items, cap), kPackMaxCtas);
        int32_t a_ncols = n_cols, a_d = d, a_ng = (int32_t)grid.y, a_ns = (int32_t)L.n_slabs;
        int64_t a_ldf = ld_full;
        void *args[] = {&a_ncols, &a_d, &a_ng, &a_ns, &src, (void *)&dinv_col, &gmax4, &blockmax, &xp, &xfull, &a_ldf};
        H2_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)g), dim3(kPackThreads), args, 0, st));
        H2_LAUNCHED("bm_pack_i8_kernel");
        return H2_OK;
    }
    if (dg == 32) bm_pack_kernel<32><<<grid, 256, 0, st>>>(n_cols, d, grid.y, splits, src, dinv_col, (uint4 *)xpack, xfull, ld_full);
    else bm_pack_kernel<64><<<grid, 256, 0, st>>>(n_cols, d, grid.y, splits, src, dinv_col, (uint4 *)xpack, xfull, ld_full);
    H2_LAUNCHED("bm_pack_kernel");
    return H2_OK;
}

int gather_rows(int32_t n_cols, int32_t d, int32_t n_parts, const float *const *ptrs, const int64_t *bounds, int64_t ld,
                float *xfull, int64_t ld_full, h2_stream_t s, bool x_bf16) {
    if (x_bf16) {   // bf16 rows: a pure copy, moved as d / 2 fp32 words (d, ld, ld_full multiples of 8)
        H2_REQUIRE(d % 8 == 0 && ld % 8 == 0 && ld_full % 8 == 0, H2_ERR_ALIGN, "gather_rows: bf16 rows need d, ld multiples of 8");
        d /= 2; ld /= 2; ld_full /= 2;
    }
    H2_REQUIRE(n_cols >= 0 && d > 0 && d % 4 == 0 && xfull && ld % 4 == 0 && ld_full % 4 == 0 && ld >= d && ld_full >= d &&
               aligned16(xfull), H2_ERR_INVALID, "gather_rows: bad argument");
    if (n_cols == 0) return H2_OK;
    PackSrc src;
    int rc = fill_src(src, n_cols, n_parts, ptrs, bounds, ld);
    if (rc != H2_OK) return rc;
    const int64_t total = (int64_t)n_cols * (d / 4);
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)s>>>(n_cols, d / 4, src, xfull, ld_full);
    H2_LAUNCHED("gather_rows_kernel");
    return H2_OK;
}
}  // namespace h2

// X' = diag(dinv_col) X packed into bf16 pieces — shared by every bitmap hop of a round that uses the same dinv_col.
extern "C" int h2_bm_pack_x_f32(int32_t n_cols, int32_t d, int32_t splits, const float *X, int64_t ldx,
                                const float *dinv_col, void *xpack, size_t xpack_bytes, h2_stream_t s) {
    H2_REQUIRE(X && aligned16(X), H2_ERR_INVALID, "h2_bm_pack_x_f32: null / misaligned X");
    const int64_t bounds[2] = {0, n_cols};
    return bm_pack_parts(n_cols, d, splits, 1, &X, bounds, ldx, dinv_col, xpack, xpack_bytes, nullptr, 0, s, true, false);
}

// the same on a buffer whose header is known to be armed (h2_graph_bind_workspace zeroes it once): no memset per round
extern "C" int h2_bm_pack_x_f32_armed(int32_t n_cols, int32_t d, int32_t splits, const float *X, int64_t ldx,
                                      const float *dinv_col, void *xpack, size_t xpack_bytes, h2_stream_t s) {
    H2_REQUIRE(X && aligned16(X), H2_ERR_INVALID, "h2_bm_pack_x_f32: null / misaligned X");
    const int64_t bounds[2] = {0, n_cols};
    return bm_pack_parts(n_cols, d, splits, 1, &X, bounds, ldx, dinv_col, xpack, xpack_bytes, nullptr, 0, s, false, false);
}

// opt-in shared memory size: set once per kernel instantiation and device, not on every launch
template <typename K>
static int bm_smem_once(K kern, size_t smem, bool *done) {
    int dev = 0;
    H2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !done[dev]) {
        H2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    return H2_OK;
}

template <int DG, int S>
static int bm_launch(const BmSched &sc, const char *base, const BmParams &p, cudaStream_t st) {
    constexpr size_t smem = BmCfg<DG, S>::kSmem;
    auto kern = bm_mma_kernel<DG, S>;
    static bool attr_done[64] = {};
    { int rc = bm_smem_once(kern, smem, attr_done); if (rc != H2_OK) return rc; }
    kern<<<sc.n_ctas, bm_threads(false), smem, st>>>(p);
    H2_LAUNCHED("bm_mma_kernel");
    if (sc.n_fix > 0) {
        bm_fixup_kernel<DG><<<dim3((kTileRows * DG / 4 + 255) / 256, sc.n_fix), 256, 0, st>>>((const BmFix *)(base + sc.off_fix), p.n_rows, p.d, p.partial, p.Y, p.ldy);
        H2_LAUNCHED("bm_fixup_kernel");
    }
    return H2_OK;
}

// Y[:, off:off+d] = diag(dinv_row) . P . X'  for the bitmap-format pattern P (X' from h2_bm_pack_x_f32).
static int bm_spmm_impl(const void *bm_host, const void *bm_dev, int32_t d, int32_t splits, const void *xpack,
                        const float *dinv_row, float *Y, int64_t ldy, int64_t out_col_off, void *partial_ws,
                        size_t partial_bytes, h2_stream_t s, bool y_bf16);

extern "C" int h2_bm_spmm_f32(const void *bm_host, const void *bm_dev, int32_t d, int32_t splits, const void *xpack,
                              const float *dinv_row, float *Y, int64_t ldy, int64_t out_col_off, void *partial_ws,
                              size_t partial_bytes, h2_stream_t s) {
    return bm_spmm_impl(bm_host, bm_dev, d, splits, xpack, dinv_row, Y, ldy, out_col_off, partial_ws, partial_bytes, s, false);
}

namespace h2 {
// bf16 output rows (ldy / out_col_off count bf16 elements); int8 digits only
int bm_spmm_bf16_out(const void *bm_host, const void *bm_dev, int32_t d, int32_t splits, const void *xpack, const float *dinv_row,
                     void *Y, int64_t ldy, int64_t out_col_off, void *partial_ws, size_t partial_bytes, h2_stream_t s) {
    H2_REQUIRE(splits_i8(splits) && d % 8 == 0 && ldy % 8 == 0 && out_col_off % 8 == 0, H2_ERR_UNSUPPORTED,
               "bf16 output needs the int8 digits and d, ldy, offsets multiples of 8");
    return bm_spmm_impl(bm_host, bm_dev, d, splits, xpack, dinv_row, (float *)Y, ldy, out_col_off, partial_ws, partial_bytes, s, true);
}
}  // namespace h2

static int bm_spmm_impl(const void *bm_host, const void *bm_dev, int32_t d, int32_t splits, const void *xpack,
                        const float *dinv_row, float *Y, int64_t ldy, int64_t out_col_off, void *partial_ws,
                        size_t partial_bytes, h2_stream_t s, bool y_bf16) {
    cudaStream_t st = (cudaStream_t)s;
    const BmHost *h = (const BmHost *)bm_host;
    H2_REQUIRE(h && h->magic == kBmMagic && bm_dev && xpack && Y, H2_ERR_INVALID, "h2_bm_spmm_f32: bad plan / null argument");
    H2_REQUIRE(d > 0 && d % 4 == 0 && ldy % 4 == 0 && out_col_off % 4 == 0 && out_col_off >= 0 && out_col_off + d <= ldy &&
               aligned16(Y) && aligned16(xpack), H2_ERR_ALIGN, "h2_bm_spmm_f32: d=%d ldy=%lld off=%lld alignment", d,
               (long long)ldy, (long long)out_col_off);
    H2_REQUIRE(splits_valid(splits), H2_ERR_INVALID, "h2_bm_spmm_f32: splits must be 2, 3, H2_SPLITS_I8X2 or H2_SPLITS_I8X3");
    const bool i8 = splits_i8(splits);
    H2_REQUIRE(h->bit_order == (i8 ? 1 : 0), H2_ERR_INVALID, "h2_bm_spmm_f32: plan was filled with bit order %d, splits=%d needs %d "
               "(h2_bm_fill_order)", h->bit_order, splits, i8 ? 1 : 0);
    const char *base = (const char *)bm_dev;
    if (h->n_empty_tiles > 0) {
        // bf16 rows: the same kernel zeroes d/2 "floats" per row of a buffer whose float stride is ldy/2
        if (y_bf16) bm_zero_tiles_kernel<<<(unsigned)h->n_empty_tiles, 256, 0, st>>>((const int32_t *)(base + h->off_empty_tiles), h->n_rows, d / 2,
                                                                                 (float *)((uint16_t *)Y + out_col_off), ldy / 2);
        else bm_zero_tiles_kernel<<<(unsigned)h->n_empty_tiles, 256, 0, st>>>((const int32_t *)(base + h->off_empty_tiles), h->n_rows, d, Y + out_col_off, ldy);
        H2_LAUNCHED("bm_zero_tiles_kernel");
    }
    if (pair_applies(splits, d)) {
        // CTA-pair kernel: split tiles are finished in the kernel (no fix-up launch), any number of columns
        const int fh = pair_fh(d), ng = pair_groups(d);
        H2_REQUIRE(ng <= 4, H2_ERR_UNSUPPORTED, "h2_bm_spmm_f32: d=%d needs %d column groups of %d (max 4): split the columns", d, ng, 2 * fh);
        const BmSched &sp = h->sched_pair[sched_index(ng)];
        H2_REQUIRE(sp.n_partial_slots == 0 || (partial_ws && partial_bytes >= h2_bm_partial_bytes(bm_host, d, splits) && aligned16(partial_ws)),
                   H2_ERR_WORKSPACE, "h2_bm_spmm_f32: partial workspace too small");
        if (sp.n_ctas == 0) return H2_OK;
        BmPairParams pp;
        pp.unit_chunk = (const int32_t *)(base + h->off_unit_chunk);
        pp.bits = (const unsigned long long *)(base + h->off_bits);
        pp.seg = (const BmPairSeg *)(base + sp.off_seg);
        pp.cta_seg_ptr = (const int32_t *)(base + sp.off_cta_seg_ptr);
        pp.xpack = (const uint8_t *)xpack + kI8HeaderBytes;
        pp.xstep = (const float *)xpack;
        pp.dinv_row = dinv_row;
        pp.Y = y_bf16 ? (float *)((uint16_t *)Y + out_col_off) : Y + out_col_off;
        pp.y_bf16 = y_bf16 ? 1 : 0;
        pp.partial = (float *)partial_ws;
        pp.sync = (uint32_t *)(const_cast<char *>(base) + sp.off_fix);        // the plan buffer holds the arrival counters
        {
            const int64_t nt_ = h->n_tiles;
            const size_t max_segs = (size_t)(kNumSms / 2 + ng * (nt_ + 1) + ng * (h->n_units / kPairMaxSegUnitsHost + 1) + 2);
            pp.fix = (const BmPairFix *)(base + sp.off_fix + align_up_sz(max_segs * 8, 256));
        }
        pp.n_fix = sp.n_fix;
        pp.status = (int32_t *)(const_cast<char *>(base) + h->off_status);
        pp.ldy = ldy;
        pp.n_rows = h->n_rows; pp.d = d; pp.n_groups_fh = std::max(2, groups_for(d, fh));
        {
            // operands of one launch: packed X' tiles + bitmaps.  Beyond ~half of the 126 MB L2 they stream from DRAM, the
            // tensor pipe starves and the cross-CTA hand-over needs its cluster-scope release (bm_pair.cu, produce()).
            // H2_BM_PAIR_SAFE = 0 / 1 forces the choice (measurements, tests).
            static const int forced = [] { const char *e = getenv("H2_BM_PAIR_SAFE"); return e ? atoi(e) : -1; }();
            const double operand_bytes = (double)h2_bm_xpack_bytes(h->n_cols, d, splits) + (double)h->n_units * kTileRows * 8;
            pp.safe_handover = forced >= 0 ? forced : (operand_bytes > 48.0 * 1024 * 1024 ? 1 : 0);
        }
        return pair_launch(splits_pieces(splits), fh, sp.n_ctas, pp, st);
    }
    H2_REQUIRE(!y_bf16, H2_ERR_UNSUPPORTED, "h2_bm_spmm: bf16 rows are implemented for the int8 digits only");
    const int dg = dg_for(d, splits);
    const int n_groups = groups_for(d, dg);
    H2_REQUIRE(n_groups <= 8, H2_ERR_UNSUPPORTED, "h2_bm_spmm_f32: d=%d needs %d column groups (max 8): split the columns", d, n_groups);
    const BmSched &sc = h->sched[sched_index(n_groups)];
    H2_REQUIRE(sc.n_partial_slots == 0 || (partial_ws && partial_bytes >= h2_bm_partial_bytes(bm_host, d, splits) && aligned16(partial_ws)),
               H2_ERR_WORKSPACE, "h2_bm_spmm_f32: partial workspace too small");
    BmParams p;
    p.unit_chunk = (const int32_t *)(base + h->off_unit_chunk);
    p.bits = (const unsigned long long *)(base + h->off_bits);
    p.seg = (const BmSegment *)(base + sc.off_seg);
    p.cta_seg_ptr = (const int32_t *)(base + sc.off_cta_seg_ptr);
    p.xpack = (const uint4 *)((const char *)xpack + (i8 ? kI8HeaderBytes : 0));
    p.xstep = (const float *)xpack;
    p.dinv_row = dinv_row;
    p.Y = Y + out_col_off;
    p.partial = (float *)partial_ws;
    p.ldy = ldy;
    p.n_rows = h->n_rows; p.d = d; p.n_groups = n_groups; p.splits = splits;
    if (sc.n_ctas == 0) return H2_OK;
    if (splits == 2) return dg == 32 ? bm_launch<32, 2>(sc, base, p, st) : bm_launch<64, 2>(sc, base, p, st);
    return bm_launch<32, 3>(sc, base, p, st);
}

#ifdef H2_BM_TRACE
extern "C" int h2_debug_read(long long *host) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(host, h2::g_bm_trace, sizeof(long long) * 148 * 32);
}
extern "C" int h2_debug_read2(long long *host) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(host, h2::g_bm_trace2, sizeof(long long) * 8 * 64);
}
#endif
