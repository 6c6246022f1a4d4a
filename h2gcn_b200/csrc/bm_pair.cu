// CTA-pair int8 kernel of the tile-bitmap path: Y = diag(dinv_row) . P . X'  on `tcgen05.mma.cta_group::2.kind::i8`.
//
// Replaces GCNLayer.sparse_dense_matmul (h2gcn/models/_layers.py:62-76) for dense-ish normalised BINARY hop patterns,
// like bm_mma_kernel (bitmap_mma.cu), for every int8 round wider than 32 features.  What differs from the single-CTA
// kernel, and why (measurements: profiles/README.md r01f / r02):
//
//  * A pair of CTAs (cluster of 2) owns a 256-row tile, 128 rows each.  One MMA covers M = 256 and ALL digits of a
//    2*FH-feature column group: N = 2*S*FH = 384 for 3 digits x 128 features (issued as N = 256 + N = 128), so every
//    bitmap row is expanded ONCE per unit (the single-CTA kernel expands it once per 64-feature group) and a unit carries
//    384 cycles of tensor-pipe work per CTA instead of 192: the A producers have twice the time per hand-over.
//  * The B tile of a unit is split between the two CTAs' shared memories (each loads the [S*FH x 64] int8 tile of ITS
//    FH-feature half — byte for byte the tile bm_pack_i8_kernel writes), so the L2 -> SM operand traffic per row is that
//    of a 512-row tile.
//  * Split tiles are finished IN the kernel: stream-K ranges cut a (tile, group) item into segments; every segment of
//    a split item converts its int32 accumulators to ONE fp32 value per feature.  The segment the schedule expects to
//    end last (always the last one of its pair) is the item's DESIGNATED finisher: it drains all its accumulator blocks
//    (the second one into the B ring, free by then), waits — bounded — for the arrival counter of its CTA rank, adds the
//    other segments' slots to its staged values in slot order inside its own write-out and stores Y.  The other segments
//    park their values in a partial slot and bump the counter; if the finisher's patience runs out it does the same and
//    whoever arrives LAST adds the slots in ascending order.  Either way the sum is deterministic, nobody waits for
//    long, there is no fix-up launch and no co-residency assumption.  (Measured alternatives, profiles/README.md r02:
//    fixed sender / receiver roles without a time bound: 77 k cycles against 46-49 k; a grid barrier and an even split of
//    the rows: 10-16 k idle cycles per CTA; last-arriver only: the sum sat on the critical path of the pair that ended
//    last, 60 k against 55 k cycles for the slowest CTA.)
//  * The ranges are cut at equal COST (units + calibrated epilogue costs), not equal unit counts (pair_schedule), and the
//    producers stage the first units of the NEXT segment before they drain the accumulators of the current one.
//  * Segments are also cut every 2048 units (2^17 columns): |sum| <= 2^17 * 64 * 128 = 2^30 keeps the int32 accumulators
//    exact for any number of columns (the single-CTA kernel refuses n_cols > 2^17).
//  * Hand-overs between the two CTAs (bm_common.cuh): the accumulator hand-over of the peer is relayed by its idle warp 9
//    with a cluster-scope release; the operand hand-over carries a release fence per 4 units in launches whose operands
//    stream from DRAM (template SAFE, chosen per launch from the operand bytes) and is a plain arrive otherwise.
//
// Roles per CTA (10 warps): warps 0..7 = A producers (two groups of 4 on alternate units; thread = bitmap row: 16
// rotate+mask words -> `tcgen05.st` -> wait::st -> remote arrive on the leader's barrier), then epilogue; warp 8 = TMA
// (`cp.async.bulk`: B half + expansion constants + 1 KB of bitmap per unit); warp 9 = TMEM allocation and, in the
// leader CTA, the ONE thread that issues every MMA / commit for both CTAs — in the peer CTA, the relay of the
// accumulator hand-over.
#include <stdint.h>
#include <stdlib.h>

#include "bm_common.cuh"

namespace h2 {

constexpr int kPairGroups = 2;                              // producer groups of 4 warps per CTA
constexpr int kPairThreads = (4 * kPairGroups + 2) * 32;    // 320
// register cap of the persistent pair CTA: what it leaves of the SM's 64 K registers is what the CSR-gather CTAs of the same
// round and the pack CTAs of the next one can use while it runs.  Measured (r02, pipelined round at the north-star point,
// tools/pipeline_parts.py): 144 registers 48.1 us, 128: 46.1 us, 112: 46.3, 104: 46.6, 96 (spills): 47.9; no spills down to
// 112.  (Gather CTAs of 4 instead of 8 warps: 46.0-47.0 us, not adopted.)
#ifndef H2_BM_PAIR_MAXREGS
#define H2_BM_PAIR_MAXREGS 128
#endif
constexpr int kPairMaxRegs = H2_BM_PAIR_MAXREGS;
#ifdef H2_BM_PAIR_SEG_UNITS
constexpr int kPairMaxSegUnits = H2_BM_PAIR_SEG_UNITS;      // test builds: force many accumulator cuts on small graphs
#else
constexpr int kPairMaxSegUnits = 2048;                      // int32 accumulators: 2^17 columns of at most 64 * 128
#endif

struct BmPairSeg {        // one contiguous run of units inside one (row tile, column group), handled by one CTA pair
    int32_t tile, unit_begin, unit_end, group;
    int32_t n_slots;      // 0: the segment covers its whole item and writes Y; else the item is split into |n_slots| segments
                          // (negative: this segment is the designated finisher of the item, see the epilogue)
    int32_t slot;         // split item: this segment's partial slot (slots of an item are consecutive, unit order)
    int32_t fix;          // split item: index of the item (arrival counters)
    int32_t slot_begin;   // split item: first slot of the item
};

struct BmPairFix {        // one split (tile, group) item: Y rows = scale * sum of slots [slot_begin, slot_begin + n_slots)
    int32_t tile, group, slot_begin, n_slots;
};

template <int S, int FH, bool SAFE = false>
struct PairCfg {
    static constexpr int NBH = S * FH;                      // B rows in ONE CTA's shared memory per unit
    static constexpr int N = 2 * NBH;                       // accumulator columns: 128 / 192 / 256 / 384
    static constexpr bool kTwoInstr = N > 256;              // N = 384 -> 256 (digits 0, 1) + 128 (digit 2)
    static constexpr uint32_t kBBytes = NBH * 64 + kI8ConstBytes;
    static constexpr uint32_t kBStride = (kBBytes + 1023u) & ~1023u;
    static constexpr uint32_t kBitsBytes = 128 * 8;         // this CTA's 128 bitmap rows of a unit
    static constexpr uint32_t kACol0 = N <= 256 ? 256 : 384;
    static constexpr int kAStg = (512 - (int)kACol0) / 16 >= 16 ? 16 : 8;   // A stage = 16 TMEM columns (128 rows x 64 int8)
    // B ring depth: what fits in ~176 KB next to the epilogue stage — the rest of the SM's 227 KB stays free so that the
    // 34 KB CTAs of the NEXT round's pack kernel and the gather CTAs can share the SM (with a 224 KB ring the pipelined
    // rounds of bench.py ran 62 us apart instead of ~57: nothing else fitted next to the MMA CTA; the depth itself
    // made no difference between 9 and 13 stages, profiles/README.md r02b)
    // SAFE launches (operands streaming from DRAM, release fence on the hand-over) are latency-bound on the B ring — every
    // stage is held ~2.5 k cycles longer — and have little to gain from co-resident CTAs: they take the whole 227 KB
    // (13 stages for 3 digits x 64 features instead of 9).
#ifndef H2_BM_PAIR_SMEM_KB
#define H2_BM_PAIR_SMEM_KB 176
#endif
    static constexpr int kBudgetKB = SAFE ? 225 : H2_BM_PAIR_SMEM_KB;
    static constexpr int kBStgFit = (int)((kBudgetKB * 1024 - 1024 - 8 * 32 * 36 * 4) / (kBStride + 1024));
    static constexpr int kBStg = kBStgFit > (SAFE ? 16 : 12) ? (SAFE ? 16 : 12) : kBStgFit;
    static constexpr int kStageStride = 36;                 // floats per staged row: 32 + 4 (16-byte aligned, bank spread)
    static constexpr size_t kSmem = (size_t)kBStg * (kBStride + kBitsBytes) + 8 * 32 * kStageStride * 4 + 1024;
    static_assert(kACol0 + kAStg * 16 <= 512 && N % 16 == 0, "TMEM budget / UMMA N");
    static_assert(kBStg >= kPairGroups + 2 && kAStg >= kPairGroups + 2, "rings too small for the producer groups");
};

struct BmPairParams {
    const int32_t *unit_chunk;
    const unsigned long long *bits;
    const BmPairSeg *seg;
    const int32_t *cta_seg_ptr;  // [n_pairs + 1]
    const uint8_t *xpack;        // [n_chunks][n_groups_fh] B tiles: [S*FH x 64] int8 (SWIZZLE_64B) + 128 constant bytes
    const float *xstep;
    const float *dinv_row;
    float *Y;                    // + out_col_off applied by the host
    float *partial;              // [n_slots][256][2*FH]
    const BmPairFix *fix;        // [n_fix] split items
    uint32_t *sync;              // [n_fix][2] arrival counters (item, CTA rank), zero between launches
    int32_t *status;             // unused (kept for the ABI of the plan buffer)
    int32_t n_fix;
    int64_t ldy;
    int32_t n_rows, d, n_groups_fh;
    int32_t y_bf16;              // Y rows hold bf16 (ldy counts bf16 elements): one rounding at the store
    int32_t safe_handover;       // 1: cluster-scope release fence before the producers' arrives (operands streaming from DRAM)
};

// 4 consecutive output features at element offset `off` of the Y buffer (fp32 rows or bf16 rows)
__device__ __forceinline__ void pair_store4(float *Y, int64_t off, float4 v, int bf16) {
    if (!bf16) { *reinterpret_cast<float4 *>(Y + off) = v; return; }
    uint32_t lo, hi;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(v.y), "f"(v.x));
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v.w), "f"(v.z));
    *reinterpret_cast<uint2 *>(reinterpret_cast<uint16_t *>(Y) + off) = make_uint2(lo, hi);
}

#ifdef H2_BM_TRACE
// per CTA: 0 start, 1 end, 2 MMA thread waits full_a, 3 MMA thread waits acc_empty, 4 / 5 producer warp 0 waits full_b /
// empty_a, 6 producer warp 0 expand + store (no waits), 7 TMA thread waits empty_b, 8 epilogue (warp 0), 9 receiver spin,
// 10 units, 11 segments   (tools/dbg_pair.py)
__device__ long long g_pair_trace[148 * 16];
__device__ long long g_pair_units[8 * 96];   // CTA 0, per unit: 0 TMA issued, 1 B landed (prod), 2 expanded, 3 A stage free, 4 stored+arrived, 5 MMA saw full_a, 6 MMA issued
#define PT_UNIT(row, u) do { if (blockIdx.x == 0 && (u) < 96u) g_pair_units[(row) * 96 + (u)] = clock64(); } while (0)
#define PT_DECL() long long pt_acc__[4] = {0, 0, 0, 0}
#define PT_BEGIN() const long long pt_t0__ = clock64()
#define PT_END(k) pt_acc__[k] += clock64() - pt_t0__
#define PT_STORE(k, slot) g_pair_trace[blockIdx.x * 16 + (slot)] = pt_acc__[k]
#define PT_STAMP(slot) g_pair_trace[blockIdx.x * 16 + (slot)] = clock64()
#else
#define PT_DECL() do { } while (0)
#define PT_BEGIN() do { } while (0)
#define PT_END(k) do { } while (0)
#define PT_STORE(k, slot) do { } while (0)
#define PT_STAMP(slot) do { } while (0)
#define PT_UNIT(row, u) do { } while (0)
#endif

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int32_t *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(int32_t *p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// SAFE: the producers' operand hand-over carries a cluster-scope release fence (p.safe_handover launches); the plain
// instantiation keeps the per-unit arrive of the L2-resident regime free of the batching bookkeeping
template <int S, int FH, bool SAFE>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(kPairMaxRegs) bm_pair_kernel(const __grid_constant__ BmPairParams p) {
    using Cfg = PairCfg<S, FH, SAFE>;
    constexpr int NBH = Cfg::NBH, N = Cfg::N;
    constexpr uint32_t kBBytes = Cfg::kBBytes, kBStride = Cfg::kBStride, kBitsBytes = Cfg::kBitsBytes;
    constexpr uint32_t kACol0 = Cfg::kACol0, kTmemCols = 512;
    constexpr int kAStg = Cfg::kAStg, kBStg = Cfg::kBStg, kStageStride = Cfg::kStageStride;
#ifdef H2_BM_PAIR_AHEAD
    constexpr int kAhead = H2_BM_PAIR_AHEAD;
#else
    constexpr int kAhead = (kAStg < kBStg ? kAStg : kBStg) - 2;
#endif
    static_assert(kAhead >= 0 && kAhead < kAStg && kAhead < kBStg, "run-ahead must stay inside both rings");
#ifndef H2_BM_PAIR_SAFE_BATCH
#define H2_BM_PAIR_SAFE_BATCH 4
#endif
    constexpr int kSafeBatch = SAFE ? H2_BM_PAIR_SAFE_BATCH : 1;   // units per release fence (1 without the fence: lowest hand-over latency)
    // a batch holds A stages of units u, u + G, ..., u + G (batch - 1): each must have been freed by a unit OLDER than u
    static_assert(kPairGroups * (kSafeBatch - 1) < kAStg && kPairGroups * (kSafeBatch - 1) < kBStg, "arrive batch must stay inside the A / B rings");
    // D int32 | A, B signed int8 | N | M = 256 (128 rows per CTA)
    constexpr uint32_t kN1 = Cfg::kTwoInstr ? 256 : N, kN2 = N - kN1;
    constexpr uint32_t kIdesc1 = (2u << 4) | (1u << 7) | (1u << 10) | ((kN1 >> 3) << 17) | ((256u >> 4) << 24);
    constexpr uint32_t kIdesc2 = (2u << 4) | (1u << 7) | (1u << 10) | (((kN2 ? kN2 : 16u) >> 3) << 17) | ((256u >> 4) << 24);

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint32_t smem_raw_u32 = smem_u32(smem_raw);
    asm volatile("" : "+r"(smem_raw_u32));   // keep the window address in a register (bitmap_mma.cu, same reason)
    const uint32_t smem_base = (smem_raw_u32 + 1023u) & ~1023u;
    const uint32_t bits_base = smem_base + kBStg * kBStride;
    const unsigned long long *bits_gen = reinterpret_cast<const unsigned long long *>(smem_raw + (bits_base - smem_raw_u32));
    float *stage_gen = reinterpret_cast<float *>(smem_raw + (bits_base - smem_raw_u32) + kBStg * kBitsBytes);
    __shared__ uint64_t s_bar[2 * kAStg + 2 * kBStg + 3];
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_chunk[32];
    __shared__ bool s_last;
    __shared__ bool s_fast[2];
    uint32_t bar0 = smem_u32(&s_bar[0]);
    asm volatile("" : "+r"(bar0));
    const uint32_t bar_full_a = bar0;                                  // leader only: 8 producer warps (4 of each CTA)
    const uint32_t bar_empty_a = bar0 + 8 * kAStg;                     // multicast commit
    const uint32_t bar_full_b = bar0 + 8 * (2 * kAStg);                // local TMA
    const uint32_t bar_empty_b = bar0 + 8 * (2 * kAStg + kBStg);       // multicast commit
    const uint32_t bar_acc_full = bar0 + 8 * (2 * kAStg + 2 * kBStg);  // multicast commit
    const uint32_t bar_acc_empty = bar_acc_full + 8;                   // leader only: its 8 epilogue warps + the peer's relay
    const uint32_t bar_acc_drained = bar_acc_full + 16;                // peer only: its 8 epilogue warps (local)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int seg_begin = p.cta_seg_ptr[pair], seg_end = p.cta_seg_ptr[pair + 1];
    const int n_work = seg_end - seg_begin;
    constexpr int kTmaWarp = 4 * kPairGroups, kMmaWarp = 4 * kPairGroups + 1;
    // the first segment and its first 32 chunk indices are fetched BEHIND the set-up below (barrier init, TMEM allocation,
    // cluster sync): two dependent global loads that otherwise delay the first TMA copy by ~2 k cycles
    BmPairSeg seg0 = BmPairSeg{0, 0, 0, 0, 0, 0, 0, 0};
    if (n_work > 0) seg0 = p.seg[seg_begin];
    int nxt0 = 0;
    if (warp == kTmaWarp && seg0.unit_begin + lane < seg0.unit_end) nxt0 = p.unit_chunk[seg0.unit_begin + lane];

    if (threadIdx.x == 0) {
        PT_STAMP(0);
        for (int s = 0; s < kAStg; ++s) {
            mbar_init(bar_full_a + 8 * s, 8);
            mbar_init(bar_empty_a + 8 * s, 1);
        }
        for (int s = 0; s < kBStg; ++s) {
            mbar_init(bar_full_b + 8 * s, 1);
            mbar_init(bar_empty_b + 8 * s, 1);
        }
        mbar_init(bar_acc_full, 1);
        mbar_init(bar_acc_empty, 4 * kPairGroups + 1);
        mbar_init(bar_acc_drained, 4 * kPairGroups);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {   // the same warp of both CTAs
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                     "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = s_tmem_base;

    if (warp == kTmaWarp) {
        // ===== TMA producer: this CTA's half of the B tile (+ constants) and of the unit's bitmap =====
        uint32_t it = 0;
        PT_DECL();
        for (int w = 0; w < n_work; ++w) {
            const BmPairSeg sg = w == 0 ? seg0 : p.seg[seg_begin + w];
            const int gfh = 2 * sg.group + (int)rank;        // FH-feature group whose tile is this CTA's half
            int nxt = w == 0 ? nxt0 : (sg.unit_begin + lane < sg.unit_end ? p.unit_chunk[sg.unit_begin + lane] : 0);
            for (int u0 = sg.unit_begin; u0 < sg.unit_end; u0 += 32) {
                s_chunk[lane] = nxt;
                __syncwarp();
                if (u0 + 32 + lane < sg.unit_end) nxt = p.unit_chunk[u0 + 32 + lane];
                const int cnt = min(32, sg.unit_end - u0);
                if (elect_one()) {
                    for (int k = 0; k < cnt; ++k) {
                        const int chunk = s_chunk[k];
                        const uint32_t st = (it + k) % kBStg, ph = ((it + k) / kBStg) & 1;
                        { PT_BEGIN(); mbar_wait(bar_empty_b + 8 * st, ph ^ 1); PT_END(0); }
                        PT_UNIT(0, it + k);
                        mbar_arrive_expect_tx(bar_full_b + 8 * st, kBBytes + kBitsBytes);
                        const uint8_t *src = p.xpack + ((int64_t)chunk * p.n_groups_fh + gfh) * kBBytes;
                        bulk_copy_g2s(smem_base + st * kBStride, src, kBBytes, bar_full_b + 8 * st);
                        bulk_copy_g2s(bits_base + st * kBitsBytes, p.bits + (int64_t)(u0 + k) * kTileRows + rank * 128, kBitsBytes,
                                      bar_full_b + 8 * st);
                    }
                }
                it += cnt;
                __syncwarp();
            }
        }
        if (elect_one()) PT_STORE(0, 7);
    } else if (warp == kMmaWarp) {
        // ===== MMA issuer: the leader's elected lane, for both CTAs =====
        if (rank == 0 && elect_one()) {
            uint32_t it = 0, acc_it = 0;
            PT_DECL();
            for (int w = 0; w < n_work; ++w) {
                const BmPairSeg sg = p.seg[seg_begin + w];
                { PT_BEGIN(); mbar_wait_cluster(bar_acc_empty, (acc_it & 1) ^ 1); PT_END(1); }   // both epilogues have drained the accumulators
                tc_fence_after();
                uint32_t acc = 0;
                for (int u = sg.unit_begin; u < sg.unit_end; ++u, ++it) {
                    const uint32_t sa = it % kAStg, sb = it % kBStg;
                    { PT_BEGIN(); mbar_wait_cluster(bar_full_a + 8 * sa, (it / kAStg) & 1); PT_END(0); }   // both CTAs: A stored (and B landed)
                    PT_UNIT(5, it);
                    tc_fence_after();
                    const uint32_t a0 = tmem_base + kACol0 + sa * 16;
                    const uint32_t b0 = smem_base + sb * kBStride;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        umma_i8_ts_pair(tmem_base, a0 + k * 8, umma_desc_sw64(b0 + k * 32), kIdesc1, k > 0 ? 1u : acc);
                        if constexpr (Cfg::kTwoInstr)   // digit 2: B rows [128, 192) of each CTA -> accumulator columns [256, 384)
                            umma_i8_ts_pair(tmem_base + 256, a0 + k * 8, umma_desc_sw64(b0 + 128 * 64 + k * 32), kIdesc2, k > 0 ? 1u : acc);
                    }
                    acc = 1;
                    umma_commit_pair(bar_empty_a + 8 * sa);
                    umma_commit_pair(bar_empty_b + 8 * sb);
                    PT_UNIT(6, it);
                }
                umma_commit_pair(bar_acc_full);
                ++acc_it;
            }
            PT_STORE(0, 2); PT_STORE(1, 3);
#ifdef H2_BM_TRACE
            g_pair_trace[blockIdx.x * 16 + 10] = it; g_pair_trace[blockIdx.x * 16 + 11] = n_work;
#endif
        } else if (rank != 0 && elect_one()) {
            // ===== peer CTA: relay of the accumulator hand-over =====
            // The peer's epilogue warps arrive on a LOCAL barrier (CTA scope is all they need); this otherwise idle thread
            // forwards every completed phase to the leader's barrier with the cluster-scope release the hand-over between
            // CTAs asks for (bm_common.cuh).  The ~2.5 k-cycle fence of that release used to sit in every epilogue warp
            // between its accumulator drain and its write-out, i.e. on the critical path at the end of every pair.
            const uint32_t acc_empty_leader = mapa_u32(bar_acc_empty, 0);
            for (int w = 0; w < n_work; ++w) {
                mbar_wait(bar_acc_drained, w & 1);
                mbar_arrive_remote(acc_empty_leader);
            }
        }
        __syncwarp();
    } else {
        // ===== A producers (4-warp groups on alternate units, one bitmap row per thread), then epilogue =====
        const int grp = warp >> 2, quarter = warp & 3;
        const int r = quarter * 32 + lane;                       // row inside this CTA's 128-row half
        const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t full_a_leader = mapa_u32(bar_full_a, 0);
        const uint32_t acc_done_local = rank == 0 ? bar_acc_empty : bar_acc_drained;   // see the relay in the peer's warp 9
        float *stage = stage_gen + warp * (32 * kStageStride);
        const float xstep = __ldg(p.xstep);
        uint32_t it = 0, acc_it = 0;   // units / accumulator phases before this segment
        PT_DECL();
        // this group's units hnd in [from, to) of a segment whose first unit is the it0-th unit of this CTA pair
        auto produce = [&](uint32_t it0, int from, int to) {
            uint32_t pend_first = 0;   // A stage of the batch's first unit; the others follow every kPairGroups stages
            int n_pend = 0;
            constexpr int arrive_batch = kSafeBatch;
            const int skip = (int)((uint32_t)(grp + kPairGroups - (int)(it0 % kPairGroups)) % kPairGroups);
            const int first = from + (int)((uint32_t)(skip + kPairGroups - from % kPairGroups) % kPairGroups);
            for (int hnd = first; hnd < to; hnd += kPairGroups) {
                const uint32_t iu = it0 + (uint32_t)hnd;
                const uint32_t sb = iu % kBStg, pb = (iu / kBStg) & 1;
                { PT_BEGIN(); mbar_wait(bar_full_b + 8 * sb, pb); PT_END(0); }
#ifdef H2_BM_TRACE
                const long long pt_x0 = clock64();
                if (quarter == 0 && lane == 0) PT_UNIT(1, iu);
#endif
                // word j = columns 4j..4j+3 as bytes 0 / 2^t(column) (i8_expand_word); {rotate, mask} per word ride behind
                // the B tile
                const uint4 *cst = reinterpret_cast<const uint4 *>(smem_raw + (smem_base - smem_raw_u32) + sb * kBStride + NBH * 64);
                const unsigned long long b0 = bits_gen[sb * 128 + r];
                const uint32_t x0 = (uint32_t)b0, x1 = (uint32_t)(b0 >> 32);
                uint32_t a[16];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint4 c = cst[q];
                    const uint32_t x = q < 4 ? x0 : x1;
                    a[2 * q] = i8_expand_word(x, c.x, c.y);
                    a[2 * q + 1] = i8_expand_word(x, c.z, c.w);
                }
                const uint32_t sa = iu % kAStg, pa = (iu / kAStg) & 1;
#ifdef H2_BM_TRACE
                pt_acc__[2] += clock64() - pt_x0;
                if (quarter == 0 && lane == 0) PT_UNIT(2, iu);
#endif
                { PT_BEGIN(); mbar_wait(bar_empty_a + 8 * sa, pa ^ 1); PT_END(1); }
#ifdef H2_BM_TRACE
                const long long pt_x1 = clock64();
                if (quarter == 0 && lane == 0) PT_UNIT(3, iu);
#endif
                tc_fence_after();
                cuda::ptx::tcgen05_st_32x32b(t_lane + kACol0 + sa * 16, a);
                // Hand-over to the leader's MMA thread.  p.safe_handover: ONE cluster-scope release fence per batch of
                // kSafeBatch units of this group, then a plain arrive per staged unit — what the memory model asks for when a
                // thread of the PEER CTA publishes operands (its TMA-landed B half, its tcgen05.st A rows) to the MMA thread of
                // the leader.  The fence costs ~2.5 k cycles: per unit it made the producers the bottleneck (38 -> 59 us per
                // tensor hop at the north-star point, 52 us in batches of 4), so rounds whose operands sit in L2 keep the
                // plain arrive (see the header: there the tensor pipe trails the arrivals by hundreds of cycles and the
                // hazard has never been observed; r02 stress tests), and shards whose operands stream from DRAM — where the
                // MMA is issued the moment the last arrival lands, and 1 launch in 3 had one stale 128-row half — take the fence,
                // which the starved tensor pipe hides.  Stages of a batch: units u, u + 2, ... < u + kAStg.
                if (n_pend++ == 0) pend_first = sa;
                if (n_pend == arrive_batch || hnd + kPairGroups >= to) {
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");   // ~22 cycles (tools/umma_i8_probe.cu part 3)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (SAFE) {
                            fence_release_cluster();
#pragma unroll
                            for (int q = 0; q < kSafeBatch; ++q)
                                if (q < n_pend) mbar_arrive_remote_relaxed(full_a_leader + 8 * ((pend_first + q * kPairGroups) % kAStg));
                        } else {
                            mbar_arrive_remote_cta(full_a_leader + 8 * sa);
                        }
                    }
                    n_pend = 0;
                }
#ifdef H2_BM_TRACE
                pt_acc__[2] += clock64() - pt_x1;
                if (quarter == 0 && lane == 0) PT_UNIT(4, iu);
#endif
            }
        };
        // Run-ahead: before the epilogue of segment w the producers already stage the first kAhead units of segment
        // w + 1 (their A stages become free as the MMA thread works through the LAST units of w, so nothing here waits
        // for the epilogue), then drain the accumulators while the MMA thread still has ~2 units of w to go.  The tensor
        // pipe restarts on w + 1 the moment the accumulators are free instead of after the write-out + one production
        // round trip (r02c trace: ~5 k idle cycles per segment boundary before).  kAhead < ring depths: unit k of
        // w + 1 needs the MMA of unit (k - depth) to have retired, always a unit of w or older.
        int pre = 0;   // units of the current segment staged during the previous segment
        for (int w = 0; w < n_work; ++w, ++acc_it) {
            const BmPairSeg sg = p.seg[seg_begin + w];
            const int n_units = sg.unit_end - sg.unit_begin;
            produce(it, pre, n_units);
            it += (uint32_t)n_units;
            pre = 0;
            if (kAhead > 0 && w + 1 < n_work) {
                const BmPairSeg nx = p.seg[seg_begin + w + 1];
                pre = min(kAhead, nx.unit_end - nx.unit_begin);
                produce(it, 0, pre);
            }
#ifdef H2_BM_TRACE
            const long long pt_e0 = clock64();
#endif

            // ---- epilogue: warp = (TMEM lane quarter, FH-feature half `sub` of the accumulator columns) ----
            const int sub = grp;                                    // kPairGroups == 2: one half each
            const int gfh = 2 * sg.group + sub;
            const int f_base = gfh * FH;                            // first feature of this warp's columns
            const int64_t grow = (int64_t)sg.tile * kTileRows + rank * 128 + r;
            // row scales of the 32 rows of this warp, for the coalesced write-out below (lane = row here)
            float *s_scale = stage + 32;                            // staged rows have 4 spare floats each: row r's scale at [r * 36 + 32]
            mbar_wait(bar_acc_full, acc_it & 1);
            tc_fence_after();
            // Split item, DESIGNATED finisher (n_slots < 0: the segment the schedule expects to end last, always the last
            // one of its pair): after staging its first block it waits — bounded — until the other segments have
            // published their slots, then the write-out below adds them to the staged values in slot order and stores Y:
            // no slot store, no fence, no counter round trip, no separate pass over the slots (r02d trace: that pass
            // cost ~9 k cycles at the very end of the pair that arrived last).  If the patience runs out the segment
            // publishes its slot like the others and whoever arrives last sums them (below): nobody waits for long,
            // no co-residency assumption.
            constexpr int kFastMaxSlots = 4;
            constexpr long long kFinisherPatience = 24000;          // cycles
            const int n_slots = sg.n_slots < 0 ? -sg.n_slots : sg.n_slots;
            const bool designated = sg.n_slots < 0 && n_slots <= kFastMaxSlots && w + 1 == n_work;
            bool fast = false;
            const int own_pos = sg.slot - sg.slot_begin;
            constexpr int kRowsPerInstr = 4;                        // 8 lanes x float4 = one 32-float row segment
            const int rr = lane >> 3, cc = (lane & 7) * 4;
            const int row0 = (int)rank * 128 + quarter * 32;        // first row (inside the 256-row tile) of this warp
            const float scale_row = xstep * ((grow < p.n_rows && p.dinv_row) ? p.dinv_row[grow] : 1.f);

            // accumulator block c0 (32 features of this warp's half) -> fp32 -> stage buffer `stg` (lane = row)
            auto drain = [&](int c0, float *stg) {
                __syncwarp();                   // the stage is free (previous block / segment written out)
                if (c0 == 0) s_scale[lane * kStageStride] = scale_row;
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {   // 16 accumulator columns per digit at a time: 48 live registers, not 96
                    uint32_t acc[S][16];
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        // accumulator columns: [CTA 0's B rows | CTA 1's B rows] per instruction; N = 384: digits 0, 1 in the
                        // first 256 columns (128 per CTA), digit 2 in the last 128 (64 per CTA)
                        const uint32_t col = (Cfg::kTwoInstr ? (s == 2 ? 256u + 64u * sub + c0 : 128u * sub + 64u * s + c0)
                                                             : (uint32_t)(NBH * sub + FH * s + c0)) + 16u * hb;
                        cuda::ptx::tcgen05_ld_32x32b(acc[s], t_lane + col);
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (hb == 1 && c0 + 32 >= FH) {   // last columns read: this warp is done with the accumulators
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_local(acc_done_local);
                    }
#pragma unroll
                    for (int q = 0; q < 16; q += 4) {
                        float4 o;
                        float *po = &o.x;
#pragma unroll
                        for (int e = 0; e < 4; ++e) {   // digits, most significant first: sum_s 256^(S-1-s) * acc_s
                            float v = (float)(int32_t)acc[S - 1][q + e];
                            float wgt = 256.f;
#pragma unroll
                            for (int s = S - 2; s >= 0; --s, wgt *= 256.f) v = fmaf((float)(int32_t)acc[s][q + e], wgt, v);
                            po[e] = v;
                        }
                        *reinterpret_cast<float4 *>(stg + lane * kStageStride + 16 * hb + q) = o;
                    }
                }
                __syncwarp();
            };
            // coalesced write-out of a staged block: an instruction covers 4 rows x 32 floats.  fast: + the other slots, -> Y;
            // split item: -> this segment's partial slot; whole item: -> Y
            auto write_out = [&](int c0, const float *stg) {
                const int fcol = sub * FH + c0 + cc;                               // column inside the 2*FH-wide group
                const bool col_ok = f_base + c0 + cc + 4 <= p.d;
                if (fast) {
                    // the (at most 3) other slots in ascending order around the own one; with one or two of them all 8 row
                    // instructions of the block are loaded in ONE pass (every load in flight before the first sum: one L2
                    // round trip per block instead of two), with three in two passes
                    const float *pb = p.partial + ((int64_t)sg.slot_begin * kTileRows + row0 + rr) * (2 * FH) + fcol;
                    auto finish_rows = [&](auto n_others_c, auto rows_c, int jh) {
                        constexpr int NO = decltype(n_others_c)::value, NR = decltype(rows_c)::value;
                        float4 t[NO][NR];
#pragma unroll
                        for (int o = 0; o < NO; ++o) {
                            const int k = o < own_pos ? o : o + 1;     // slot index of the o-th other slot
#pragma unroll
                            for (int i = 0; i < NR; ++i)
                                t[o][i] = k < n_slots ? __ldcg(reinterpret_cast<const float4 *>(pb + ((int64_t)k * kTileRows + jh + i * kRowsPerInstr) * (2 * FH)))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int i = 0; i < NR; ++i) {
                            const int row = jh + i * kRowsPerInstr + rr;
                            const float4 own = *reinterpret_cast<const float4 *>(stg + row * kStageStride + cc);
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                            for (int k = 0; k <= NO; ++k) {            // ascending slot order, the own slot at its place
                                if (k < n_slots) {   // compile-time indices only (a run-time index would put t[] in local memory)
                                    const float4 lo = t[k < NO ? k : NO - 1][i], hi = t[k > 0 ? k - 1 : 0][i];
                                    const float4 a = k == own_pos ? own : (k < own_pos ? lo : hi);
                                    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
                                }
                            }
                            const float sc = s_scale[row * kStageStride];
                            v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
                            const int64_t gr = (int64_t)sg.tile * kTileRows + row0 + row;
                            if (col_ok && gr < p.n_rows) pair_store4(p.Y, gr * p.ldy + (int64_t)sg.group * (2 * FH) + fcol, v, p.y_bf16);
                        }
                    };
                    if (n_slots <= 3) {
                        finish_rows(std::integral_constant<int, 2>{}, std::integral_constant<int, 8>{}, 0);
                    } else {
                        finish_rows(std::integral_constant<int, 3>{}, std::integral_constant<int, 4>{}, 0);
                        finish_rows(std::integral_constant<int, 3>{}, std::integral_constant<int, 4>{}, 4 * kRowsPerInstr);
                    }
                    return;
                }
#pragma unroll
                for (int j = 0; j < 32; j += kRowsPerInstr) {
                    const int row = j + rr;
                    float4 v = *reinterpret_cast<const float4 *>(stg + row * kStageStride + cc);
                    const int64_t gr = (int64_t)sg.tile * kTileRows + row0 + row;
                    if (sg.n_slots) {
                        float *dst = p.partial + ((int64_t)sg.slot * kTileRows + row0 + row) * (2 * FH) + fcol;
                        __stcg(reinterpret_cast<float4 *>(dst), v);
                    } else {
                        const float sc = s_scale[row * kStageStride];
                        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
                        if (col_ok && gr < p.n_rows) pair_store4(p.Y, gr * p.ldy + (int64_t)sg.group * (2 * FH) + fcol, v, p.y_bf16);
                    }
                }
            };
            if (designated) {
                // the pair's LAST segment: every MMA has completed and every TMA copy has been consumed, so the B ring is
                // free — its first bytes stage the second block, and ALL accumulators are drained before the wait
                float *stage2 = reinterpret_cast<float *>(smem_raw + (smem_base - smem_raw_u32)) + warp * (32 * kStageStride);
                static_assert(8 * 32 * Cfg::kStageStride * 4 <= Cfg::kBStg * Cfg::kBStride && FH <= 64, "second stage inside the B ring");
#pragma unroll 1
                for (int c0 = 0; c0 < FH; c0 += 32) drain(c0, c0 == 0 ? stage : stage2);
                if (threadIdx.x == 0) {
                    const int32_t *cnt = reinterpret_cast<const int32_t *>(p.sync + 2 * sg.fix + rank);
                    const long long t0 = clock64();
                    bool f = ld_acquire_gpu(cnt) == n_slots - 1;
                    while (!f && clock64() - t0 < kFinisherPatience) {
                        __nanosleep(64);
                        f = ld_acquire_gpu(cnt) == n_slots - 1;
                    }
                    if (f) p.sync[2 * sg.fix + rank] = 0;       // nobody else arrives any more: re-armed for the next launch
                    s_fast[acc_it & 1] = f;
                }
                named_bar_sync(1, 32 * 4 * kPairGroups);
                fast = s_fast[acc_it & 1];
#pragma unroll 1
                for (int c0 = 0; c0 < FH; c0 += 32) write_out(c0, c0 == 0 ? stage : stage2);
            } else {
#pragma unroll 1
                for (int c0 = 0; c0 < FH; c0 += 32) {
                    drain(c0, stage);
                    write_out(c0, stage);
                }
            }
            if (sg.n_slots && !fast) {
                // ---- split item: publish the slot; the LAST CTA (of this rank) to arrive adds all slots and writes Y ----
                constexpr int kEpiThreads = 32 * 4 * kPairGroups;
                __threadfence();                               // this thread's slot stores, before the counter
                named_bar_sync(1, kEpiThreads);
                if (threadIdx.x == 0) {
                    uint32_t *cnt = p.sync + 2 * sg.fix + rank;
                    const uint32_t prev = atomicAdd(cnt, 1u);
                    s_last = prev + 1 == (uint32_t)n_slots;
                    if (s_last) { *cnt = 0; __threadfence(); }  // re-armed for the next launch; acquire side of the hand-over
                }
                named_bar_sync(1, kEpiThreads);
                if (s_last) {
                    constexpr int kLanes = 2 * FH / 4;         // float4 lanes per row: 32 (FH = 64) / 16 (FH = 32)
                    constexpr int kRowsPerWarp = 32 / kLanes;
                    const int sub_row = lane / kLanes, c = (lane % kLanes) * 4;
                    const bool c_ok = sg.group * (2 * FH) + c + 4 <= p.d;
                    // kU rows per warp and pass, kS slots at a time: every load of a pass is in flight before the first
                    // addition (r02 trace: one slot after the other cost n_slots L2 round trips per pass, ~12 k cycles on
                    // the critical path of the CTA that finishes last); the additions keep the ascending slot order.
                    constexpr int kU = 4, kS = 4;
                    const int r_lo = (int)rank * 128;
                    const float *pbase = p.partial + (int64_t)sg.slot_begin * kTileRows * (2 * FH) + c;
                    for (int r0 = r_lo + (warp * kU) * kRowsPerWarp + sub_row; r0 < r_lo + 128; r0 += 8 * kU * kRowsPerWarp) {
                        float4 v[kU];
#pragma unroll
                        for (int i = 0; i < kU; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int k0 = 0; k0 < n_slots; k0 += kS) {
                            float4 t[kS][kU];
#pragma unroll
                            for (int k = 0; k < kS; ++k) {
#pragma unroll
                                for (int i = 0; i < kU; ++i) {
                                    const int rr = r0 + i * kRowsPerWarp;
                                    t[k][i] = k0 + k < n_slots
                                                  ? __ldcg(reinterpret_cast<const float4 *>(pbase + ((int64_t)(k0 + k) * kTileRows + rr) * (2 * FH)))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                                }
                            }
#pragma unroll
                            for (int k = 0; k < kS; ++k) {
#pragma unroll
                                for (int i = 0; i < kU; ++i) { v[i].x += t[k][i].x; v[i].y += t[k][i].y; v[i].z += t[k][i].z; v[i].w += t[k][i].w; }
                            }
                        }
#pragma unroll
                        for (int i = 0; i < kU; ++i) {
                            const int64_t gr = (int64_t)sg.tile * kTileRows + r0 + i * kRowsPerWarp;
                            if (gr < p.n_rows && c_ok) {
                                const float sc = xstep * (p.dinv_row ? p.dinv_row[gr] : 1.f);
                                pair_store4(p.Y, gr * p.ldy + (int64_t)sg.group * (2 * FH) + c,
                                            make_float4(v[i].x * sc, v[i].y * sc, v[i].z * sc, v[i].w * sc), p.y_bf16);
                            }
                        }
                    }
                }
                named_bar_sync(1, kEpiThreads);                // s_last is reused by the next segment
            }
#ifdef H2_BM_TRACE
            pt_acc__[3] += clock64() - pt_e0;
#endif
        }
        if (warp == 0 && lane == 0) { PT_STORE(0, 4); PT_STORE(1, 5); PT_STORE(2, 6); PT_STORE(3, 8); }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) PT_STAMP(1);
    cluster_sync_all();   // no CTA frees tensor memory or exits while its peer can still signal it
    if (warp == kMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- host: stream-K schedule with roles ------------------------------------------------------------------------------
// Work items (group, unit), group-major, are cut into `n_pairs` contiguous ranges; a range is split at (group, tile)
// boundaries and every kPairMaxSegUnits units.  Items covered by several segments get consecutive partial slots (unit
// order = summation order) and an arrival counter.
//
// The ranges are cut at equal COST, not equal unit counts.  Costs in units of one MMA step, calibrated on the per-pair
// traces of the north-star round (i8x3, d = 128: a unit is ~384 cycles; profiles/README.md r02c): a pair with one
// segment spends ~11 k cycles outside its MMAs (start-up 3.4 k, last epilogue 7.7 k), every further segment ~8 k more
// (accumulator drain with the tensor pipe idle, slot store, fence, counter, refill): kSender ~ 24 per split segment,
// kWhole ~ 10 for a segment that writes Y itself; kFix = extra charge for the segment holding an item's first units
// (0 since the designated finisher adds the slots inside its own write-out):
//   cost(range) = units + sum over its segments of (whole ? kWhole : kSender) + kFix per split item whose head it holds
// The smallest T for which a greedy left-to-right cut (every pair takes units while its cost stays <= T) needs at most
// n_pairs ranges is found by scanning T upwards.  H2_PAIR_COSTS="whole,sender,fix" overrides the constants (measurement
// knob; measured: 38.0-38.1 us per pipelined tensor hop for 13,21,0 / 10,24,0 / 13,27,0, 39.1 for 8,12,6).
void pair_schedule(const std::vector<int64_t> &tp, int64_t n_units, int ng, int n_pairs_max, std::vector<BmPairSeg> &segs,
                   std::vector<int32_t> &pair_ptr, std::vector<BmPairFix> &fixes, int *n_slots_out) {
    const int64_t nt = (int64_t)tp.size() - 1, total = n_units * ng;
    const int G = (int)std::min<int64_t>(n_pairs_max, total);
    segs.clear();
    fixes.clear();
    pair_ptr.assign(G + 1, 0);
    *n_slots_out = 0;
    if (G <= 0) return;
    int64_t kWhole = 10, kSender = 24, kFix = 0;
    if (const char *e = getenv("H2_PAIR_COSTS")) {
        long a = 0, b = 0, c = 0;
        if (sscanf(e, "%ld,%ld,%ld", &a, &b, &c) == 3 && a >= 0 && b >= 0 && c >= 0) { kWhole = a; kSender = b; kFix = c; }
    }
    // end (in linear units) of the segment that starts at q: the item's end or the accumulator cut
    auto seg_limit = [&](int64_t q, int64_t *item_begin, int64_t *item_end) {
        const int64_t grp = q / n_units, u = q % n_units;
        const int64_t t = (int64_t)(std::upper_bound(tp.begin(), tp.end(), u) - tp.begin()) - 1;   // tile of unit u
        *item_begin = grp * n_units + tp[t];
        *item_end = grp * n_units + tp[t + 1];
        return std::min<int64_t>(*item_end, q + kPairMaxSegUnits);
    };
    // greedy cut for a cost bound T: fills `bounds` (range ends) and returns the number of ranges used
    auto cut = [&](int64_t T, std::vector<int64_t> *bounds) {
        int used = 0;
        int64_t q = 0;
        while (q < total) {
            int64_t cost = 0;
            const int64_t q_start = q;
            while (q < total) {
                int64_t ib, ie;
                const int64_t lim = seg_limit(q, &ib, &ie);
                // entering a segment at q: if it runs to the end of its item AND starts at the item's begin it is whole
                const bool at_begin = q == ib;
                const int64_t fixed_whole = kWhole, fixed_split = kSender + (at_begin ? kFix : 0);
                // take the whole remainder of the item if it fits as a whole / closing segment
                const bool closes_item = lim == ie;
                const int64_t full_cost = (lim - q) + ((at_begin && closes_item) ? fixed_whole : fixed_split);
                if (cost + full_cost <= T) { cost += full_cost; q = lim; continue; }
                // partial: as many units as fit next to the split epilogue (at least 1 if the range is still empty).
                // A head or a tail shorter than the epilogue it costs is not worth a slot: the range rather ends at
                // the item boundary, or leaves the next pair a tail of at least kMinPiece units.
                int64_t room = T - cost - fixed_split;
                const int64_t kMinPiece = 4;
                if (q != q_start && room < (at_begin ? kMinPiece : 1)) break;
                if (room < 1) room = 1;
                int64_t take = std::min<int64_t>(room, lim - q);
                if (closes_item && (lim - q) - take < kMinPiece && (lim - q) - take > 0)
                    take = std::max<int64_t>(1, (lim - q) - kMinPiece);
                q += take;
                break;
            }
            ++used;
            if (bounds) bounds->push_back(q);
        }
        return used;
    };
    // smallest feasible T, scanning up from the perfect-balance bound in steps of 0.2 % (the greedy cut with its
    // minimum-piece rules is not monotone in T, so no bisection)
    int64_t lo = std::max<int64_t>(1, (total + kWhole * std::min<int64_t>(nt * ng, G)) / G);
    const int64_t step = std::max<int64_t>(1, lo / 512);
    while (cut(lo, nullptr) > G) lo += step;
    std::vector<int64_t> bounds;
    const int used = cut(lo, &bounds);
    for (int c = 0; c < G; ++c) {
        const int64_t q0 = c == 0 ? 0 : (c - 1 < used ? bounds[c - 1] : total), q1 = c < used ? bounds[c] : total;
        pair_ptr[c] = (int)segs.size();
        int64_t q = q0;
        while (q < q1) {
            int64_t ib, ie;
            const int64_t e = std::min<int64_t>(seg_limit(q, &ib, &ie), q1);
            const int64_t grp = q / n_units, u = q % n_units;
            const int64_t t = (int64_t)(std::upper_bound(tp.begin(), tp.end(), u) - tp.begin()) - 1;
            segs.push_back(BmPairSeg{(int32_t)t, (int32_t)u, (int32_t)(u + (e - q)), (int32_t)grp, 0, 0, 0, 0});
            q = e;
        }
    }
    pair_ptr[G] = (int)segs.size();
    // split items: segments of one item are consecutive in `segs` (unit order); they get consecutive slots + a counter
    int n_slots = 0;
    for (size_t i = 0; i < segs.size();) {
        size_t j = i + 1;
        while (j < segs.size() && segs[j].tile == segs[i].tile && segs[j].group == segs[i].group) ++j;
        if (j - i > 1) {
            fixes.push_back(BmPairFix{segs[i].tile, segs[i].group, n_slots, (int32_t)(j - i)});
            for (size_t k = i; k < j; ++k) {
                segs[k].n_slots = (int32_t)(j - i);
                segs[k].slot_begin = n_slots;
                segs[k].slot = n_slots + (int32_t)(k - i);
                segs[k].fix = (int32_t)fixes.size() - 1;
            }
            n_slots += (int32_t)(j - i);
        }
        i = j;
    }
    // designated finisher of every split item: the segment expected to END last (estimated finish time inside its pair),
    // provided it is the last segment of its pair (a finisher may wait for its mates; nothing may queue behind it)
    std::vector<int64_t> finish(segs.size(), 0);
    std::vector<char> last_of_pair(segs.size(), 0);
    for (int c = 0; c < G; ++c) {
        int64_t t = 0;
        for (int i = pair_ptr[c]; i < pair_ptr[c + 1]; ++i) {
            t += (segs[i].unit_end - segs[i].unit_begin) + (segs[i].n_slots ? kSender : kWhole);
            finish[i] = t;
        }
        if (pair_ptr[c + 1] > pair_ptr[c]) last_of_pair[pair_ptr[c + 1] - 1] = 1;
    }
    if (!getenv("H2_PAIR_NO_FINISHER"))
        for (size_t i = 0; i < segs.size();) {
            size_t j = i + 1;
            while (j < segs.size() && segs[j].tile == segs[i].tile && segs[j].group == segs[i].group) ++j;
            if (j - i > 1) {
                size_t best = i;
                for (size_t k = i + 1; k < j; ++k)
                    if (finish[k] >= finish[best]) best = k;
                if (last_of_pair[best]) segs[best].n_slots = -segs[best].n_slots;
            }
            i = j;
        }
    *n_slots_out = n_slots;
}

}  // namespace h2

// Host-only self-check of the pair schedule (no GPU needed: the CPU test suite calls it through ctypes).  `tile_units[t]` =
// non-empty units of row tile t.  Builds the schedule for `ng` column groups on `n_pairs` CTA pairs and verifies its
// invariants; returns 0 and fills stats[0..7] = {segments, split items, slots, longest segment, max units of a pair,
// min units of a pair, designated finishers, pairs used}, or a negative code naming the violated invariant.
extern "C" int h2_debug_pair_schedule_check(int32_t n_tiles, const int64_t *tile_units, int32_t ng, int32_t n_pairs, int64_t *stats) {
    using namespace h2;
    if (n_tiles < 0 || !tile_units || ng < 1 || n_pairs < 1 || !stats) return -100;
    std::vector<int64_t> tp((size_t)n_tiles + 1, 0);
    for (int t = 0; t < n_tiles; ++t) { if (tile_units[t] < 0) return -100; tp[t + 1] = tp[t] + tile_units[t]; }
    const int64_t n_units = tp[n_tiles];
    std::vector<BmPairSeg> segs;
    std::vector<int32_t> ptr;
    std::vector<BmPairFix> fixes;
    int n_slots = 0;
    pair_schedule(tp, n_units, ng, n_pairs, segs, ptr, fixes, &n_slots);
    if (n_units * ng == 0) return segs.empty() ? 0 : -1;
    const int G = (int)ptr.size() - 1;
    if (G < 1 || G > n_pairs || ptr[0] != 0 || ptr[G] != (int)segs.size()) return -2;
    // every (group, unit) exactly once, in group-major unit order across the pairs
    int64_t q = 0, longest = 0, max_pair = 0, min_pair = INT64_MAX, finishers = 0, used = 0;
    for (int c = 0; c < G; ++c) {
        if (ptr[c + 1] < ptr[c]) return -3;
        int64_t units = 0;
        for (int i = ptr[c]; i < ptr[c + 1]; ++i) {
            const BmPairSeg &sg = segs[i];
            const int64_t len = (int64_t)sg.unit_end - sg.unit_begin;
            if (len <= 0 || len > kPairMaxSegUnits) return -4;                                  // int32 accumulators
            if (sg.tile < 0 || sg.tile >= n_tiles || sg.group < 0 || sg.group >= ng) return -5;
            if (sg.unit_begin < tp[sg.tile] || sg.unit_end > tp[sg.tile + 1]) return -6;       // inside ONE row tile
            if ((int64_t)sg.group * n_units + sg.unit_begin != q) return -7;                    // contiguous, nothing twice
            q += len;
            units += len;
            longest = std::max(longest, len);
            const int ns = sg.n_slots < 0 ? -sg.n_slots : sg.n_slots;
            const bool whole = sg.unit_begin == tp[sg.tile] && sg.unit_end == tp[sg.tile + 1];
            if (whole != (ns == 0)) return -8;                                                  // split items have slots
            if (ns && (ns < 2 || sg.slot < sg.slot_begin || sg.slot >= sg.slot_begin + ns || sg.slot_begin < 0 ||
                       sg.slot_begin + ns > n_slots || sg.fix < 0 || sg.fix >= (int)fixes.size())) return -9;
            if (sg.n_slots < 0) { ++finishers; if (i != ptr[c + 1] - 1) return -10; }           // a finisher ends its pair
        }
        if (units) { ++used; max_pair = std::max(max_pair, units); min_pair = std::min(min_pair, units); }
    }
    if (q != n_units * ng) return -11;
    // per split item: consecutive slots in unit order, one counter, at most one finisher
    for (size_t f = 0; f < fixes.size(); ++f) {
        int seen = 0, fin = 0;
        int64_t next_unit = -1;
        for (const BmPairSeg &sg : segs) {
            if (sg.n_slots == 0 || sg.fix != (int)f) continue;
            if (sg.tile != fixes[f].tile || sg.group != fixes[f].group || sg.slot != fixes[f].slot_begin + seen) return -12;
            if (next_unit >= 0 && sg.unit_begin != next_unit) return -13;
            next_unit = sg.unit_end;
            ++seen;
            fin += sg.n_slots < 0;
        }
        if (seen != fixes[f].n_slots || fin > 1) return -14;
    }
    stats[0] = (int64_t)segs.size(); stats[1] = (int64_t)fixes.size(); stats[2] = n_slots; stats[3] = longest;
    stats[4] = max_pair; stats[5] = used ? min_pair : 0; stats[6] = finishers; stats[7] = used;
    return 0;
}

namespace h2 {

size_t pair_partial_bytes(int n_slots, int fh) { return (size_t)n_slots * kTileRows * (2 * fh) * 4; }

template <int S, int FH, bool SAFE>
static int pair_launch_ts(int n_pairs, const BmPairParams &p, cudaStream_t st) {
    using Cfg = PairCfg<S, FH, SAFE>;
    auto kern = bm_pair_kernel<S, FH, SAFE>;
    static bool attr_done[64] = {};
    int dev = 0;
    H2_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        H2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    kern<<<2 * n_pairs, kPairThreads, Cfg::kSmem, st>>>(p);   // __cluster_dims__(2, 1, 1)
    H2_LAUNCHED("bm_pair_kernel");
    return H2_OK;
}

template <int S, int FH>
static int pair_launch_t(int n_pairs, const BmPairParams &p, cudaStream_t st) {
    return p.safe_handover ? pair_launch_ts<S, FH, true>(n_pairs, p, st) : pair_launch_ts<S, FH, false>(n_pairs, p, st);
}

int pair_launch(int S, int fh, int n_pairs, const BmPairParams &p, cudaStream_t st) {
    if (S == 2) return fh == 32 ? pair_launch_t<2, 32>(n_pairs, p, st) : pair_launch_t<2, 64>(n_pairs, p, st);
    return fh == 32 ? pair_launch_t<3, 32>(n_pairs, p, st) : pair_launch_t<3, 64>(n_pairs, p, st);
}

}  // namespace h2

#ifdef H2_BM_TRACE
extern "C" int h2_debug_read_pair(long long *host) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(host, h2::g_pair_trace, sizeof(long long) * 148 * 16);
}
extern "C" int h2_debug_read_pair_units(long long *host) {
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(host, h2::g_pair_units, sizeof(long long) * 8 * 96);
}
#endif
