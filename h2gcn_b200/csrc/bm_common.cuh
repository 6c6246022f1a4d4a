// Shared definitions of the tile-bitmap / tcgen05 path (bitmap_mma.cu: format, pack, single-CTA kernel; bm_pair.cu: the
// CTA-pair int8 kernel): format constants, plan structures, PTX helpers.
#pragma once
#include <cuda/ptx>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace h2 {

constexpr int kTileRows = 256;   // two UMMA M=128 accumulators share every B tile
constexpr int kChunkCols = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int kAStages = 4;    // A tiles live in TMEM: 4 x (2 halves x 32 columns) next to 2 x 128 accumulator columns
constexpr int kBStages = 8;    // B tiles (<= 16 KB) + unit bitmaps (2 KB) in shared memory
// single-CTA (bf16-piece) kernel: ONE set of 8 A-producer / epilogue warps (keeps the register footprint small enough for a
// CSR-gather CTA of the same round to share the SM) + TMA warp + MMA warp (last)
__host__ __device__ constexpr int bm_producer_sets(bool) { return 1; }
__host__ __device__ constexpr int bm_threads(bool) { return (8 * 1 + 2) * 32; }
constexpr uint32_t kBmMagic = 0x48324234u;  // "H2B4"
constexpr int kAStagesMax = 8;

// `splits` codes (include/h2gcn_b200.h): 2 / 3 = bf16 pieces; H2_SPLITS_I8X2 / H2_SPLITS_I8X3 = int8 digits with
// per-4-row block exponents (kind::i8, exact int32 accumulation)
__host__ __device__ constexpr bool splits_valid(int s) { return s == 2 || s == 3 || s == H2_SPLITS_I8X2 || s == H2_SPLITS_I8X3; }
__host__ __device__ constexpr bool splits_i8(int s) { return s == H2_SPLITS_I8X2 || s == H2_SPLITS_I8X3; }
__host__ __device__ constexpr int splits_pieces(int s) { return s == H2_SPLITS_I8X2 ? 2 : (s == H2_SPLITS_I8X3 ? 3 : s); }
// largest magnitude S balanced base-256 digits in [-128, 127] can carry on both signs: 127 * (256^S - 1) / 255
__host__ __device__ constexpr int i8_range(int S) { return S == 2 ? 32639 : 8355711; }
constexpr int kI8Levels = 6;          // row exponents t in 0..6: the A operand carries 2^t (<= 64) instead of 1
constexpr int kI8ConstBytes = 128;    // per B tile: 16 x {rotate amount, byte mask} for the A producers
constexpr int kI8HeaderBytes = 256;   // xpack header: fp32 quantisation step
constexpr int kAbsmaxRows = 32;       // rows per CTA of bm_absmax_kernel (8 warps x 4 rows)

// Bit position of column c (0..63) of a unit row.  Order 0: natural.  Order 1 (int8 path): the four columns of an
// operand word sit 8 bits apart, so that word j = 4 bytes {0, 2^t} comes out of ONE rotate + ONE mask:
//   bit(c) = 32 * (c / 32) + 8 * (c % 4) + (c % 32) / 4
__host__ __device__ constexpr int bm_bit_pos(int c, int order) {
    return order == 0 ? c : 32 * (c / 32) + 8 * (c % 4) + (c % 32) / 4;
}

struct BmSegment {       // one contiguous run of units inside one (row tile, column group), handled by one CTA
    int32_t tile;
    int32_t unit_begin;  // global unit index
    int32_t unit_end;
    int32_t partial_slot;  // -1: covers the whole tile -> write Y directly; else index into the partial workspace
    int32_t group;         // column group (DG features) this segment computes
    int32_t pad;
};

struct BmFix {           // one (row tile, column group) whose result is the ordered sum of partial slots
    int32_t tile;
    int32_t slot_begin;
    int32_t slot_end;
    int32_t group;
};

constexpr int kNumScheds = 4;   // stream-K schedules for 1, 2, 4, 8 column groups (work items are group-major)
constexpr int kNumPairScheds = 3;
struct BmSched {
    int32_t n_ctas, n_partial_slots, n_fix, pad;
    int64_t off_seg, off_cta_seg_ptr, off_fix;
};

struct BmHost {          // host header (caller's bm_host buffer)
    uint32_t magic;
    int32_t n_rows, n_cols, n_tiles, n_chunks;
    int32_t bit_order;   // bm_bit_pos order of the stored bitmaps: 0 (bf16 kernel) / 1 (int8 kernel)
    int64_t n_units;
    int64_t nnz;
    // offsets (bytes) into the device plan buffer
    int64_t off_unit_chunk, off_bits, off_empty_tiles, n_empty_tiles, off_status;
    BmSched sched[kNumScheds];
    BmSched sched_pair[kNumPairScheds];   // CTA-pair kernel (bm_pair.cu): <= 74 pairs, 1 / 2 / 4 column groups; n_fix = split items,
                                          // off_fix = their arrival counters
};

// ------------------------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a converged warp, chosen by `elect.sync`: unlike `lane == 0` the compiler knows the region is
// single-lane and emits straight-line uniform-datapath code for the UTCHMMA / UBLKCP / UTCBAR instructions in it.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred px;\n\t"
        "elect.sync _|px, 0xffffffff;\n\t"
        "@px mov.s32 %0, 1;\n\t"
        "}" : "+r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// one attempt, no loop: lets a caller start the (slow, ~200-cycle) phase check of the NEXT unit ahead of time
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory (128 lanes x 8 columns of packed bf16 pairs per K=16 step), B from shared memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_64B operand tile: rows of 64 bytes, 8-row atoms of 512 bytes (verified with tools/umma_i8_probe.cu)
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// K-major, SWIZZLE_128B operand tile: rows of 128 bytes, 8-row atoms of 1024 bytes (SBO), version 1 (sm_100).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}


// One 32-bit word of the int8 A operand: columns 4j..4j+3 of a bitmap row as bytes 0 / 2^t(column).  The four bits sit 8
// apart (bm_bit_pos order 1): rotate bit b to bit 0 of byte b, keep them, spread each over its byte (x 0xFF: no carries
// between bytes holding 0 / 1) and keep bit t(b) of byte b — `mask` = sum_b 2^(8 b + t(b)) comes from the pack kernel.
__device__ __forceinline__ uint32_t i8_expand_word(uint32_t x, uint32_t rot, uint32_t mask) {
    return ((__funnelshift_r(x, x, rot) & 0x01010101u) * 0xFFu) & mask;
}

// ---- cluster / cta_group::2 helpers (bm_pair.cu) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// Hand-overs BETWEEN the two CTAs of a pair (producer / epilogue warps of either CTA -> the leader's MMA thread).
// Formally they need release / acquire at CLUSTER scope: with the default (.cta) scope nothing orders the peer CTA's
// TMA-landed B half and tcgen05.st A rows before the MMA the leader issues on its behalf.  Measured (r02, profiles/README.md):
// with plain arrives a round is always right while the operands sit in L2 (the MMA thread trails the arrivals by hundreds of
// cycles), but on wide row shards whose operands stream from DRAM the MMA is issued the moment the last arrival lands, and
// ~1 launch in 3 had ONE stale 128-row half of one tile (n_cols = 262144; a 1 us sleep before the arrive hides it, a
// tcgen05.ld read-back of the staged rows does not, acquire.cluster on the waiting side alone does not; release.cluster on
// the arriving side fixes it: 0 of 60 launches wrong).  The release costs a ~2.5 k-cycle fence:
//   * accumulator hand-over (once per segment and warp): always release.cluster (mbar_arrive_remote);
//   * operand hand-over (once per unit): fence_release_cluster() + relaxed arrives in batches, only when the launch is
//     flagged safe_handover (operands larger than L2: the starved tensor pipe hides the fence) — bm_pair.cu, produce().
// H2_BM_PAIR_SCOPE_MODE (measurement knob): 0 = .cta on both sides (the old behaviour), 1 = release.cluster arrive only,
// 2 = acquire.cluster wait only, 3 = both.
#ifndef H2_BM_PAIR_SCOPE_MODE
#define H2_BM_PAIR_SCOPE_MODE 3
#endif
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
#if H2_BM_PAIR_SCOPE_MODE == 0 || H2_BM_PAIR_SCOPE_MODE == 2
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#else
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
#endif
}
// the two halves of a release.cluster arrive, for callers that publish several barriers behind one fence
__device__ __forceinline__ void fence_release_cluster() {
#if H2_BM_PAIR_SCOPE_MODE == 1 || H2_BM_PAIR_SCOPE_MODE == 3
    asm volatile("fence.acq_rel.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {   // a barrier of THIS CTA: release at CTA scope
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_cta(uint32_t cluster_addr) {   // default semantics: release at CTA scope
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // arrivals come from the peer CTA too
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
#if H2_BM_PAIR_SCOPE_MODE == 0 || H2_BM_PAIR_SCOPE_MODE == 1
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
#else
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
#endif
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {   // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_i8_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}


}  // namespace h2
