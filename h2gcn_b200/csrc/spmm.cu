// Fused multi-hop CSR x dense aggregation round (SURVEY.md §8a rows a6+a7+a8) — the gather path.
//
// One launch computes, for every hop h and every local vertex i,
//     Y[i, off_h : off_h + d] = sum_k val_h[k] * X[col_h[k], 0:d]          (k over row i of \bar{A}_h)
// i.e. GCNLayer.call + Flatten of the reference (h2gcn/models/_layers.py:62-81, H2GCN.py:271-272) with the hop
// outputs written directly at their concat offsets (ConcatLayer, _layers.py:90-96, becomes a no-op).
//
// Schedule: the hops are stacked into n_hops*n_rows "virtual rows".  The plan orders them by stored-entry count,
// descending (longest-processing-time first).  Virtual rows longer than `cta_threshold` get a whole CTA: each of
// its 8 warps reduces a contiguous slice of the row and the 8 partial rows are added in a FIXED order in shared
// memory.  The remaining virtual rows get one warp each.  No atomics anywhere, so the result is bit-reproducible
// run to run and independent of how rows are sharded over GPUs.
//
// Inner loop: a warp loads 32 (col, val) pairs coalesced, then walks them with warp shuffles; every X row is
// fetched with 128-bit read-only loads (`ld.global.nc.v4`), 8 rows in flight per warp, and the next batch's pairs are
// loaded before the current batch's gathers are issued.  When d/4 < 32 the warp is split into 32/LPR groups that take
// different nonzeros and are combined with a butterfly at the end.  The row type (fp32 / bf16) is a template parameter
// and full batches are unpredicated: 8 issued instructions per stored entry (r02; 48 before, issue-bound), which puts the
// kernel on the L2 -> SM roofline for long rows (82.6 % `lts__throughput` on the all-CSR round, profiles/README.md r02).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace h2 {

struct HopDev {
    const int64_t *rowptr;
    const int32_t *col;
    const float *val;
    const float *dinv;
    const float *dinv_row;
    int64_t out_off;
    int64_t in_off;
};

struct __align__(16) RowDesc {   // one schedule slot: where the virtual row's entries are, which row it is
    int64_t begin;
    int32_t len;
    int32_t vrow;
};

struct RoundParams {
    HopDev hop[H2_MAX_HOPS];
    const RowDesc *perm;   // [n_vrows] sorted by len descending; nullptr: natural order (plan-less callers)
    const float *X;     // fp32 rows, or bf16 rows when x_bf16 (then ldx and the in offsets count bf16 elements)
    float *Y;           // fp32 rows, or bf16 rows when y_bf16
    const float *bias;  // optional epilogue (sparse_dense): + bias[d], relu
    int64_t ldx, ldy;
    int64_t n_vrows;
    int64_t n_cta_rows;
    int32_t n_rows;
    int32_t n_hops;
    int32_t d4;  // d / 4
    int32_t relu;
    int32_t x_bf16, y_bf16;   // BASELINE config 5: bf16 features in / out, fp32 accumulation
};

// 4 consecutive features of a row: one 16-byte fp32 load, or one 8-byte load of 4 bf16 widened to fp32.  The row type is
// a COMPILE-TIME parameter: as a run-time flag inside the gather loop it cost a branch region per load (r02: 48 issued
// instructions per stored entry, the kernel was issue-bound at 71 % — profiles/README.md r02c).
template <bool XB>
__device__ __forceinline__ float4 load_x4(const char *p) {
    if constexpr (!XB) {
        return __ldg(reinterpret_cast<const float4 *>(p));
    } else {
        const uint2 u = __ldg(reinterpret_cast<const uint2 *>(p));
        return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                           __uint_as_float(u.y & 0xFFFF0000u));
    }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void store_y4(float *row, int j, float4 v, int bf16) {
    if (!bf16) reinterpret_cast<float4 *>(row)[j] = v;
    else reinterpret_cast<uint2 *>(row)[j] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// ---- plan kernels ----------------------------------------------------------------------------------------------
struct HopPtrs {
    const int64_t *rowptr[H2_MAX_HOPS];
};

__global__ void plan_keys_kernel(HopPtrs hp, int32_t n_rows, int64_t n_vrows, uint32_t *keys, int32_t *vrow) {
    int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (v >= n_vrows) return;
    int h = (int)(v / n_rows);
    int i = (int)(v - (int64_t)h * n_rows);
    int64_t len = hp.rowptr[h][i + 1] - hp.rowptr[h][i];
    keys[v] = len > 0xffffffffLL ? 0xffffffffu : (uint32_t)len;
    vrow[v] = (int32_t)v;
}

// keys sorted descending: count of keys > thr = first index with key <= thr; also total and max.
__global__ void plan_desc_kernel(HopPtrs hp, int32_t n_rows, int64_t n_vrows, const int32_t *__restrict__ vrow_sorted,
                                 RowDesc *__restrict__ out) {
    int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k >= n_vrows) return;
    const int v = vrow_sorted[k];
    const int h = v / n_rows, i = v - h * n_rows;
    const int64_t b = hp.rowptr[h][i], e = hp.rowptr[h][i + 1];
    RowDesc d;
    d.begin = b;
    d.len = (int32_t)min(e - b, (int64_t)0x7fffffff);
    d.vrow = v;
    out[k] = d;
}

__global__ void plan_counts_kernel(const uint32_t *keys_sorted, int64_t n_vrows, uint32_t thr, HopPtrs hp,
                                   int32_t n_rows, int32_t n_hops, PlanCounts *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int64_t lo = 0, hi = n_vrows;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (keys_sorted[mid] > thr) lo = mid + 1; else hi = mid;
    }
    out->n_cta_rows = lo;
    int64_t tot = 0;
    for (int h = 0; h < n_hops; ++h) tot += hp.rowptr[h][n_rows] - hp.rowptr[h][0];
    out->total_nnz = tot;
    out->max_row_nnz = n_vrows > 0 ? (int64_t)keys_sorted[0] : 0;
}

// ---- the fused round -------------------------------------------------------------------------------------------
// One batch step of accumulate_segment: G*U stored entries, U per lane group, all their X-row loads issued before the
// first FMA.  FULL: the batch holds 32 entries and nothing is predicated.  Otherwise the entries past `cnt` re-load the
// batch's last valid row (a row this output row reads anyway — no stray address, no select on the loaded values) and
// only their FMAs are predicated off.  Lanes past the row width load column 0 instead (qoff) and never store.
template <int LPR, int NV, int U, bool XB, bool FULL>
__device__ __forceinline__ void gather_step(int t, int cnt, int c, float v, const char *xlane, int row_bytes, const int (&qoff)[NV],
                                            int grp, float4 (&acc)[NV]) {
    constexpr int G = 32 / LPR;
    float4 x[U][NV];
    float w[U];
    bool on[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int idx = t + u * G + grp;            // < 32: G * U <= 32 and t is a multiple of G * U
        on[u] = FULL || idx < cnt;
        const int src = FULL ? idx : min(idx, cnt - 1);
        const int cc = __shfl_sync(0xffffffffu, c, src);
        w[u] = __shfl_sync(0xffffffffu, v, src);
        const char *xr = xlane + (int64_t)cc * (int64_t)row_bytes;   // one IMAD.WIDE
#pragma unroll
        for (int q = 0; q < NV; ++q) x[u][q] = load_x4<XB>(xr + qoff[q]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (on[u]) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                acc[q].x = fmaf(w[u], x[u][q].x, acc[q].x);
                acc[q].y = fmaf(w[u], x[u][q].y, acc[q].y);
                acc[q].z = fmaf(w[u], x[u][q].z, acc[q].z);
                acc[q].w = fmaf(w[u], x[u][q].w, acc[q].w);
            }
        }
    }
}

// entries [s, e) of one CSR row: 32 (col, val) pairs per coalesced load, broadcast by shuffles; per entry the warp issues
// 2 SHFL + 1 IMAD.WIDE + NV loads + 4 NV FFMA.  `xrow0` = row 0 of X (this hop's column slice), `row_bytes` the row
// stride in bytes (< 2^31).
template <int LPR, int NV, bool XB>
__device__ __forceinline__ void accumulate_segment(const int32_t *__restrict__ col, const float *__restrict__ val,
                                                   const float *__restrict__ dinv, int64_t s, int64_t e,
                                                   const char *__restrict__ xrow0, int row_bytes, int d4, int lane,
                                                   float4 (&acc)[NV]) {
    constexpr int G = 32 / LPR;             // nonzeros processed side by side
    constexpr int U0 = (NV >= 4) ? 2 : ((NV == 2) ? 4 : 8);
    constexpr int U = G * U0 > 32 ? 32 / G : U0;   // rows in flight per group
    constexpr int kVec = XB ? 8 : 16;       // bytes of 4 features
    const int grp = lane / LPR;
    const int lig = lane % LPR;
    const char *xlane = xrow0 + (lig < d4 ? lig * kVec : 0);
    asm volatile("" : "+l"(xlane));         // keep the lane's base in a register pair (otherwise re-derived per load)
    int qoff[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) qoff[q] = lig + q * LPR < d4 ? q * LPR * kVec : 0;
    // 32-bit loop state: the row's entries as (pointer, count) — a row holds < 2^31 entries (RowDesc.len)
    col += s;
    if (val) val += s;
    const int n = (int)(e - s);
    // the (col, val) pairs of batch b + 1 are loaded BEFORE the gathers of batch b are issued: their latency (two dependent
    // loads in factored mode) hides behind the batch instead of opening every batch with an exposed round trip
    auto load_cv = [&](int base, int &c, float &v) {
        const int k = base + lane;
        c = 0;
        v = 0.f;
        if (k < n) {
            c = __ldg(col + k);
            v = val ? __ldg(val + k) : __ldg(dinv + c);
        }
    };
    int c;
    float v;
    load_cv(0, c, v);
    for (int base = 0; base < n; base += 32) {
        int cn = 0;
        float vn = 0.f;
        if (base + 32 < n) load_cv(base + 32, cn, vn);
        if (n - base >= 32) {
#pragma unroll
            for (int t = 0; t < 32; t += G * U) gather_step<LPR, NV, U, XB, true>(t, 32, c, v, xlane, row_bytes, qoff, grp, acc);
        } else {
            const int cnt = n - base;
            for (int t = 0; t < cnt; t += G * U) gather_step<LPR, NV, U, XB, false>(t, cnt, c, v, xlane, row_bytes, qoff, grp, acc);
        }
        c = cn;
        v = vn;
    }
}

template <int LPR, int NV>
__device__ __forceinline__ void reduce_groups(float4 (&acc)[NV]) {
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, off);
            acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, off);
            acc[q].z += __shfl_xor_sync(0xffffffffu, acc[q].z, off);
            acc[q].w += __shfl_xor_sync(0xffffffffu, acc[q].w, off);
        }
    }
}

__device__ __forceinline__ float4 epilogue(float4 a, float scale, const float *bias, int j, int relu) {
    a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
    if (bias) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + j);
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    if (relu) {
        a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f);
    }
    return a;
}

template <int LPR, int NV, bool XB>
__global__ void __launch_bounds__(kCtaThreads, (NV == 1 ? 1024 : (NV == 2 ? 768 : 512)) / kCtaThreads)
fused_hops_gather_kernel(const __grid_constant__ RoundParams p) {
    extern __shared__ float4 s_part[];  // [kWarpsPerCta][d4] partial rows of a CTA-row
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t b = blockIdx.x;
    const bool cta_row = b < p.n_cta_rows;
    int64_t slot = cta_row ? b : p.n_cta_rows + (b - p.n_cta_rows) * kWarpsPerCta + warp;
    if (slot >= p.n_vrows) return;  // whole warp (only in the last warp-row CTA)
    int64_t v, s, e;
    if (p.perm) {   // one 16-byte load instead of the perm -> rowptr chain
        const int4 raw = __ldg(reinterpret_cast<const int4 *>(p.perm + slot));
        s = ((int64_t)(uint32_t)raw.y << 32) | (uint32_t)raw.x;
        e = s + raw.z;
        v = raw.w;
    } else {
        v = slot;
        s = 0; e = 0;
    }
    const int h = (int)(v / p.n_rows);
    const int i = (int)(v - (int64_t)h * p.n_rows);
    const HopDev &hop = p.hop[h];
    if (!p.perm) { s = __ldg(hop.rowptr + i); e = __ldg(hop.rowptr + i + 1); }
    if (cta_row) {
        int64_t seg = (e - s + kWarpsPerCta - 1) / kWarpsPerCta;
        seg = (seg + 31) & ~(int64_t)31;
        s = min(e, s + warp * seg);
        e = min(e, s + seg);
    }
    float4 acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    constexpr int kElt = XB ? 2 : 4;   // bytes per feature of an X row; ldx and the in offsets count elements of the row type
    const char *xrow0 = reinterpret_cast<const char *>(p.X) + hop.in_off * kElt;
    accumulate_segment<LPR, NV, XB>(hop.col, hop.val, hop.dinv, s, e, xrow0, (int)(p.ldx * kElt), p.d4, lane, acc);
    reduce_groups<LPR, NV>(acc);

    const float scale = hop.val ? 1.f : __ldg(hop.dinv_row + i);  // factored mode: dinv_i * sum_j dinv_j x_j
    float *yrow = p.y_bf16 ? reinterpret_cast<float *>(reinterpret_cast<uint16_t *>(p.Y) + (int64_t)i * p.ldy + hop.out_off)
                           : p.Y + (int64_t)i * p.ldy + hop.out_off;
    const int lig = lane % LPR;
    if (!cta_row) {
        if (lane < LPR) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const int j = lig + q * LPR;
                if (j < p.d4) store_y4(yrow, j, epilogue(acc[q], scale, p.bias, j, p.relu), p.y_bf16);
            }
        }
        return;
    }
    if (lane < LPR) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int j = lig + q * LPR;
            if (j < p.d4) s_part[warp * p.d4 + j] = acc[q];
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < p.d4; j += kCtaThreads) {
        float4 a = s_part[j];
#pragma unroll
        for (int w = 1; w < kWarpsPerCta; ++w) {  // fixed order => deterministic
            const float4 t = s_part[w * p.d4 + j];
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        store_y4(yrow, j, epilogue(a, scale, p.bias, j, p.relu), p.y_bf16);
    }
}

template <int LPR, int NV, bool XB>
static int launch_gather_t(const RoundParams &p, cudaStream_t st) {
    const int64_t warp_rows = p.n_vrows - p.n_cta_rows;
    const int64_t grid = p.n_cta_rows + (warp_rows + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid == 0) return H2_OK;
    H2_REQUIRE(grid < 0x7fffffffLL, H2_ERR_UNSUPPORTED, "fused round: grid of %lld CTAs exceeds 2^31", (long long)grid);
    // the partial rows are only needed by CTA-rows; without them the CTA takes no shared memory and fits next to the
    // persistent tensor-core CTA of the same round
    const size_t smem = p.n_cta_rows ? (size_t)kWarpsPerCta * p.d4 * sizeof(float4) : 0;
    fused_hops_gather_kernel<LPR, NV, XB><<<(unsigned)grid, kCtaThreads, smem, st>>>(p);
    H2_LAUNCHED("fused_hops_gather_kernel");
    return H2_OK;
}

template <int LPR, int NV>
static int launch_gather(const RoundParams &p, cudaStream_t st) {
    H2_REQUIRE(p.ldx * 4 < (1ll << 31), H2_ERR_UNSUPPORTED, "fused round: ldx=%lld is too wide (row stride must stay below 2 GiB)", (long long)p.ldx);
    return p.x_bf16 ? launch_gather_t<LPR, NV, true>(p, st) : launch_gather_t<LPR, NV, false>(p, st);
}

int run_gather_round(const RoundParams &p, cudaStream_t st) {
    const int d4 = p.d4;
    if (d4 <= 4) return launch_gather<4, 1>(p, st);
    if (d4 <= 8) return launch_gather<8, 1>(p, st);
    if (d4 <= 16) return launch_gather<16, 1>(p, st);
    if (d4 <= 32) return launch_gather<32, 1>(p, st);
    if (d4 <= 64) return launch_gather<32, 2>(p, st);
    if (d4 <= 128) return launch_gather<32, 4>(p, st);
    if (d4 <= 256) return launch_gather<32, 8>(p, st);
    set_error("fused round: d = %d > 1024 is not covered (split the columns)", d4 * 4);
    return H2_ERR_UNSUPPORTED;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t cub_sort_bytes(int64_t n) {
    // upper bound for cub::DeviceRadixSort with DoubleBuffer on n (u32, i32) pairs: histograms + spine only.
    return align_up((size_t)(n / 64 + 1) * 16, 256) + (1u << 20);
}

}  // namespace h2

using namespace h2;

extern "C" size_t h2_plan_host_bytes(void) { return sizeof(PlanHost); }

extern "C" size_t h2_plan_dev_bytes(int32_t n_rows, int32_t n_hops) {
    const size_t n = (size_t)(n_rows > 0 ? n_rows : 0) * (size_t)(n_hops > 0 ? n_hops : 0);
    return align_up(n * sizeof(RowDesc), 256) + 256;
}

extern "C" size_t h2_plan_workspace_bytes(int32_t n_rows, int32_t n_hops) {
    const size_t n = (size_t)(n_rows > 0 ? n_rows : 0) * (size_t)(n_hops > 0 ? n_hops : 0);
    // keys x2, vrow x2, counts, cub temp
    return 4 * align_up(n * 4, 256) + 256 + cub_sort_bytes((int64_t)n);
}

extern "C" int h2_plan_build(int32_t n_rows, int32_t n_hops, const h2_hop_t *hops, void *plan_host, void *plan_dev,
                             void *ws, size_t ws_bytes, h2_stream_t s) {
    cudaStream_t st = (cudaStream_t)s;
    H2_REQUIRE(n_rows >= 0 && n_hops >= 1 && n_hops <= H2_MAX_HOPS, H2_ERR_INVALID,
               "h2_plan_build: n_rows=%d n_hops=%d (1..%d hops)", n_rows, n_hops, H2_MAX_HOPS);
    H2_REQUIRE(hops && plan_host && (plan_dev || n_rows == 0), H2_ERR_INVALID, "h2_plan_build: null argument");
    const int64_t n = (int64_t)n_rows * n_hops;
    PlanHost *ph = (PlanHost *)plan_host;
    ph->magic = kPlanMagic;
    ph->n_rows = n_rows;
    ph->n_hops = n_hops;
    ph->cta_threshold = 256;
    ph->n_vrows = n;
    ph->n_cta_rows = 0;
    ph->total_nnz = 0;
    ph->max_row_nnz = 0;
    if (n == 0) return H2_OK;
    H2_REQUIRE(ws && ws_bytes >= h2_plan_workspace_bytes(n_rows, n_hops), H2_ERR_WORKSPACE,
               "h2_plan_build: workspace %zu < %zu bytes", ws_bytes, h2_plan_workspace_bytes(n_rows, n_hops));
    HopPtrs hp;
    for (int h = 0; h < n_hops; ++h) {
        H2_REQUIRE(hops[h].rowptr, H2_ERR_INVALID, "h2_plan_build: hop %d has no rowptr", h);
        hp.rowptr[h] = hops[h].rowptr;
    }
    char *w = (char *)ws;
    const size_t seg = align_up((size_t)n * 4, 256);
    uint32_t *keys_a = (uint32_t *)w;
    uint32_t *keys_b = (uint32_t *)(w + seg);
    int32_t *vrow_a = (int32_t *)(w + 2 * seg);
    int32_t *vrow_b = (int32_t *)(w + 3 * seg);
    PlanCounts *cnt = (PlanCounts *)(w + 4 * seg);
    void *cub_tmp = w + 4 * seg + 256;
    size_t cub_have = ws_bytes - (4 * seg + 256);
    RowDesc *perm = (RowDesc *)plan_dev;

    const int tpb = 256;
    plan_keys_kernel<<<(unsigned)((n + tpb - 1) / tpb), tpb, 0, st>>>(hp, n_rows, n, keys_a, vrow_a);
    H2_LAUNCHED("plan_keys_kernel");
    cub::DoubleBuffer<uint32_t> dk(keys_a, keys_b);
    cub::DoubleBuffer<int32_t> dv(vrow_a, vrow_b);
    size_t need = 0;
    H2_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, need, dk, dv, (int)n, 0, 32, st));
    H2_REQUIRE(need <= cub_have, H2_ERR_WORKSPACE, "h2_plan_build: sort needs %zu bytes, have %zu", need, cub_have);
    H2_CUDA(cub::DeviceRadixSort::SortPairsDescending(cub_tmp, need, dk, dv, (int)n, 0, 32, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    plan_desc_kernel<<<(unsigned)((n + tpb - 1) / tpb), tpb, 0, st>>>(hp, n_rows, n, dv.Current(), perm);
    H2_LAUNCHED("plan_desc_kernel");
    plan_counts_kernel<<<1, 32, 0, st>>>(dk.Current(), n, (uint32_t)ph->cta_threshold, hp, n_rows, n_hops, cnt);
    H2_LAUNCHED("plan_counts_kernel");
    PlanCounts hc;
    H2_CUDA(cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, st));
    H2_CUDA(cudaStreamSynchronize(st));
    ph->n_cta_rows = hc.n_cta_rows;
    ph->total_nnz = hc.total_nnz;
    ph->max_row_nnz = hc.max_row_nnz;
    return H2_OK;
}

namespace h2 {
int fill_round_params(RoundParams &p, const PlanHost *ph, const void *plan_dev, int32_t n_rows, int32_t n_hops,
                      const h2_hop_t *hops, int32_t d, const float *X, int64_t ldx, float *Y, int64_t ldy) {
    H2_REQUIRE(ph && ph->magic == kPlanMagic, H2_ERR_INVALID, "fused round: plan header is not initialised");
    H2_REQUIRE(ph->n_rows == n_rows && ph->n_hops == n_hops, H2_ERR_INVALID,
               "fused round: plan was built for n_rows=%d n_hops=%d, called with %d/%d", ph->n_rows, ph->n_hops,
               n_rows, n_hops);
    H2_REQUIRE(d > 0 && hops && (n_rows == 0 || (X && Y && plan_dev)), H2_ERR_INVALID, "fused round: null/empty argument");
    H2_REQUIRE(d % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && aligned16(X) && aligned16(Y), H2_ERR_ALIGN,
               "fused round: d=%d ldx=%lld ldy=%lld X=%p Y=%p must be 4-element / 16-byte aligned", d, (long long)ldx,
               (long long)ldy, (const void *)X, (void *)Y);
    H2_REQUIRE(ldx >= d, H2_ERR_INVALID, "fused round: ldx=%lld < d=%d", (long long)ldx, d);
    for (int h = 0; h < n_hops; ++h) {
        H2_REQUIRE(hops[h].rowptr && (hops[h].col || ph->total_nnz == 0), H2_ERR_INVALID, "fused round: hop %d null CSR", h);
        H2_REQUIRE(hops[h].val || (hops[h].dinv && hops[h].dinv_row), H2_ERR_INVALID,
                   "fused round: hop %d has neither val nor dinv/dinv_row", h);
        H2_REQUIRE(hops[h].out_col_off % 4 == 0 && hops[h].out_col_off >= 0 && hops[h].out_col_off + d <= ldy,
                   H2_ERR_ALIGN, "fused round: hop %d out_col_off=%lld (d=%d, ldy=%lld)", h,
                   (long long)hops[h].out_col_off, d, (long long)ldy);
        H2_REQUIRE(hops[h].in_col_off % 4 == 0 && hops[h].in_col_off >= 0 && hops[h].in_col_off + d <= ldx, H2_ERR_ALIGN,
                   "fused round: hop %d in_col_off=%lld (d=%d, ldx=%lld)", h, (long long)hops[h].in_col_off, d, (long long)ldx);
        p.hop[h] = HopDev{hops[h].rowptr, hops[h].col, hops[h].val, hops[h].dinv, hops[h].dinv_row, hops[h].out_col_off,
                          hops[h].in_col_off};
    }
    p.perm = (const RowDesc *)plan_dev;
    p.X = X; p.Y = Y; p.bias = nullptr;
    p.ldx = ldx; p.ldy = ldy;
    p.n_vrows = ph->n_vrows; p.n_cta_rows = ph->n_cta_rows;
    p.n_rows = n_rows; p.n_hops = n_hops; p.d4 = d / 4; p.relu = 0;
    p.x_bf16 = 0; p.y_bf16 = 0;
    return H2_OK;
}
}  // namespace h2

extern "C" int h2_fused_hops_spmm_f32(const void *plan_host, const void *plan_dev, int32_t n_rows, int32_t n_hops,
                                      const h2_hop_t *hops, int32_t d, const float *X, int64_t ldx, float *Y,
                                      int64_t ldy, h2_stream_t s) {
    H2_REQUIRE(n_hops >= 1 && n_hops <= H2_MAX_HOPS && n_rows >= 0, H2_ERR_INVALID, "fused round: n_rows=%d n_hops=%d",
               n_rows, n_hops);
    RoundParams p;
    int rc = fill_round_params(p, (const PlanHost *)plan_host, plan_dev, n_rows, n_hops, hops, d, X, ldx, Y, ldy);
    if (rc != H2_OK) return rc;
    if (n_rows == 0) return H2_OK;
    return run_gather_round(p, (cudaStream_t)s);
}

// same round with bf16 feature rows in and / or out (fp32 accumulation); ldx / ldy / offsets count elements of the row type
extern "C" int h2_fused_hops_spmm_ex(const void *plan_host, const void *plan_dev, int32_t n_rows, int32_t n_hops,
                                     const h2_hop_t *hops, int32_t d, const void *X, int64_t ldx, int32_t x_dtype, void *Y,
                                     int64_t ldy, int32_t y_dtype, h2_stream_t s) {
    H2_REQUIRE(n_hops >= 1 && n_hops <= H2_MAX_HOPS && n_rows >= 0, H2_ERR_INVALID, "fused round: n_rows=%d n_hops=%d",
               n_rows, n_hops);
    H2_REQUIRE((x_dtype == H2_F32 || x_dtype == H2_BF16) && (y_dtype == H2_F32 || y_dtype == H2_BF16), H2_ERR_INVALID,
               "fused round: dtype codes are H2_F32 / H2_BF16");
    const bool any16 = x_dtype == H2_BF16 || y_dtype == H2_BF16;
    H2_REQUIRE(!any16 || d % 8 == 0, H2_ERR_ALIGN, "fused round: bf16 rows need d %% 8 == 0 (d=%d)", d);
    H2_REQUIRE((x_dtype != H2_BF16 || ldx % 8 == 0) && (y_dtype != H2_BF16 || ldy % 8 == 0), H2_ERR_ALIGN,
               "fused round: bf16 leading dimensions must be multiples of 8");
    RoundParams p;
    int rc = fill_round_params(p, (const PlanHost *)plan_host, plan_dev, n_rows, n_hops, hops, d, (const float *)X, ldx, (float *)Y, ldy);
    if (rc != H2_OK) return rc;
    for (int h = 0; h < n_hops; ++h)
        H2_REQUIRE((x_dtype != H2_BF16 || hops[h].in_col_off % 8 == 0) && (y_dtype != H2_BF16 || hops[h].out_col_off % 8 == 0), H2_ERR_ALIGN,
                   "fused round: bf16 column offsets must be multiples of 8");
    p.x_bf16 = x_dtype == H2_BF16; p.y_bf16 = y_dtype == H2_BF16;
    if (n_rows == 0) return H2_OK;
    return run_gather_round(p, (cudaStream_t)s);
}

// ---- a5: SparseDense (+bias, +ReLU) -------------------------------------------------------------------------------
namespace h2 {
// scalar fallback for unit counts that are not a multiple of 4 (e.g. a logistic-regression "MO" first layer):
// one warp per row, lane = output column.
__global__ void sparse_dense_scalar_kernel(int32_t n_rows, const int64_t *__restrict__ rowptr,
                                           const int32_t *__restrict__ col, const float *__restrict__ val,
                                           const float *__restrict__ W, int32_t p, const float *__restrict__ bias,
                                           int relu, float *__restrict__ Y, int64_t ldy) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    for (int c0 = 0; c0 < p; c0 += 32) {
        const int c = c0 + lane;
        float acc = 0.f;
        for (int64_t k = s; k < e; ++k)
            if (c < p) acc = fmaf(__ldg(val + k), __ldg(W + (int64_t)__ldg(col + k) * p + c), acc);
        if (c < p) {
            if (bias) acc += bias[c];
            if (relu) acc = fmaxf(acc, 0.f);
            Y[row * ldy + c] = acc;
        }
    }
}
}  // namespace h2

extern "C" int h2_sparse_dense_f32(int32_t n_rows, const int64_t *rowptr, const int32_t *col, const float *val,
                                   const float *W, int32_t p, const float *bias, int32_t relu, float *Y, int64_t ldy,
                                   int64_t out_col_off, h2_stream_t s) {
    H2_REQUIRE(n_rows >= 0 && p >= 1 && out_col_off >= 0, H2_ERR_INVALID, "h2_sparse_dense_f32: n_rows=%d p=%d", n_rows, p);
    if (n_rows == 0) return H2_OK;
    H2_REQUIRE(rowptr && W && Y && val && ldy >= out_col_off + p, H2_ERR_INVALID, "h2_sparse_dense_f32: bad argument");
    const bool vec = p % 4 == 0 && p <= 1024 && ldy % 4 == 0 && out_col_off % 4 == 0 && aligned16(W) && aligned16(Y) &&
                     (!bias || aligned16(bias));
    if (!vec) {
        sparse_dense_scalar_kernel<<<(unsigned)(((int64_t)n_rows * 32 + 255) / 256), 256, 0, (cudaStream_t)s>>>(
            n_rows, rowptr, col, val, W, p, bias, relu, Y + out_col_off, ldy);
        H2_LAUNCHED("sparse_dense_scalar_kernel");
        return H2_OK;
    }
    RoundParams rp;
    rp.hop[0] = HopDev{rowptr, col, val, nullptr, nullptr, out_col_off, 0};
    rp.perm = nullptr;
    rp.X = W; rp.Y = Y; rp.bias = bias;
    rp.ldx = p; rp.ldy = ldy;
    rp.n_vrows = n_rows; rp.n_cta_rows = 0;
    rp.n_rows = n_rows; rp.n_hops = 1; rp.d4 = p / 4; rp.relu = relu;
    rp.x_bf16 = 0; rp.y_bf16 = 0;
    return run_gather_round(rp, (cudaStream_t)s);
}
