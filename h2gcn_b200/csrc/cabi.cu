// Library-level pieces of the C-ABI: error state, launch counter, and the host-buffer graph handle used by the
// end-to-end entry points (include/h2gcn_b200.h, "end-to-end entry points with HOST buffers").
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace h2 {

std::atomic<int64_t> g_launches{0};
static thread_local char t_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

}  // namespace h2

using namespace h2;

extern "C" int h2_abi_version(void) { return H2_ABI_VERSION; }
extern "C" const char *h2_last_error(void) { return t_err; }
extern "C" int64_t h2_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

static bool splits_is_i8(int32_t s);
// ---- host-buffer graph handle ----------------------------------------------------------------------------------
struct h2_graph {
    int32_t n_rows = 0, n_cols = 0, n_hops = 0, d_max = 0, splits = 2, row_begin = 0;
    std::vector<void *> owned;  // every cudaMalloc'ed pointer, freed in destroy
    h2_hop_t hops[H2_MAX_HOPS];
    const float *dinv[H2_MAX_HOPS] = {};
    // CSR subset
    int n_csr = 0;
    int csr_idx[H2_MAX_HOPS];
    std::vector<char> plan_host;
    void *plan_dev = nullptr;
    // bitmap subset
    int n_bm = 0;
    int bm_idx[H2_MAX_HOPS];
    std::vector<char> bm_host[H2_MAX_HOPS];
    void *bm_dev[H2_MAX_HOPS] = {};
    void *xpack = nullptr, *partial = nullptr;   // scratch of the tensor-core hops: a caller workspace (bound) or owned
    size_t xpack_bytes = 0, partial_bytes = 0;
    void *ws_owned = nullptr;                    // h2_graph_reserve: the library's own allocation behind xpack / partial
    float *x_dev = nullptr, *y_dev = nullptr;
    bool own_xy = false;
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_done = nullptr;               // end of the previous round on this handle (the scratch is per handle)
};

static int dev_alloc(h2_graph *g, void **p, size_t bytes) {
    H2_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    g->owned.push_back(*p);
    return H2_OK;
}

extern "C" int h2_graph_destroy(h2_graph_t *g) {
    if (!g) return H2_OK;
    for (void *p : g->owned) cudaFree(p);
    if (g->ws_owned) cudaFree(g->ws_owned);
    if (g->x_dev) cudaFree(g->x_dev);
    if (g->y_dev) cudaFree(g->y_dev);
    if (g->side) cudaStreamDestroy(g->side);
    if (g->ev_fork) cudaEventDestroy(g->ev_fork);
    if (g->ev_join) cudaEventDestroy(g->ev_join);
    if (g->ev_done) cudaEventDestroy(g->ev_done);
    delete g;
    return H2_OK;
}

// common tail of the two constructors: formats, schedules, scratch buffers
static int graph_finish(h2_graph *g, const bool *want_bitmap) {
    int rc = H2_OK;
    const int n_rows = g->n_rows, n_cols = g->n_cols;
    for (int h = 0; h < g->n_hops; ++h) {
        if (want_bitmap[h]) g->bm_idx[g->n_bm++] = h; else g->csr_idx[g->n_csr++] = h;
    }
    if (g->n_csr) {
        h2_hop_t sub[H2_MAX_HOPS];
        for (int k = 0; k < g->n_csr; ++k) sub[k] = g->hops[g->csr_idx[k]];
        g->plan_host.resize(h2_plan_host_bytes());
        void *ws = nullptr;
        const size_t ws_bytes = h2_plan_workspace_bytes(n_rows, g->n_csr);
        if ((rc = dev_alloc(g, &g->plan_dev, h2_plan_dev_bytes(n_rows, g->n_csr)))) return rc;
        H2_CUDA(cudaMalloc(&ws, ws_bytes ? ws_bytes : 16));
        rc = h2_plan_build(n_rows, g->n_csr, sub, g->plan_host.data(), g->plan_dev, ws, ws_bytes, nullptr);
        cudaFree(ws);
        if (rc) return rc;
    }
    for (int k = 0; k < g->n_bm; ++k) {
        const int h = g->bm_idx[k];
        void *iws = nullptr;
        const size_t iws_bytes = h2_bm_index_bytes(n_rows, n_cols);
        H2_CUDA(cudaMalloc(&iws, iws_bytes));
        int64_t n_units = 0;
        rc = h2_bm_count(n_rows, n_cols, g->hops[h].rowptr, g->hops[h].col, iws, iws_bytes, &n_units, nullptr);
        if (rc == H2_OK) {
            const size_t pb = h2_bm_plan_dev_bytes(n_rows, n_cols, n_units);
            rc = dev_alloc(g, &g->bm_dev[h], pb);
            g->bm_host[h].resize(h2_bm_host_bytes());
            if (rc == H2_OK)
                rc = h2_bm_fill_order(n_rows, n_cols, g->hops[h].rowptr, g->hops[h].col, iws, n_units, g->bm_host[h].data(),
                                      g->bm_dev[h], pb, splits_is_i8(g->splits) ? 1 : 0, nullptr);
        }
        cudaFree(iws);
        if (rc) return rc;
    }
    {
        cudaError_t e = cudaEventCreateWithFlags(&g->ev_done, cudaEventDisableTiming);
        if (e != cudaSuccess) return cuda_fail(e, "h2_graph: event");
    }
    if (g->n_bm && g->n_csr) {
        // The tensor-core hops run on an internal HIGH-priority stream and the CSR hops on the caller's stream: the
        // persistent MMA CTAs (1 per SM) are placed first and the gather CTAs fill the remaining register space.
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        cudaError_t e = cudaStreamCreateWithPriority(&g->side, cudaStreamNonBlocking, prio_hi);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&g->ev_join, cudaEventDisableTiming);
        if (e != cudaSuccess) return cuda_fail(e, "h2_graph: side stream");
    }
    return H2_OK;
}

// ---- scratch of the tensor-core hops (packed operand + stream-K partial tiles) ---------------------------------------
// The round entry points never allocate and never synchronise: the scratch for widths <= d_max is either a caller
// workspace (h2_graph_workspace_bytes + h2_graph_bind_workspace) or the library's own allocation made by the explicit,
// synchronising h2_graph_reserve (host-buffer handles reserve at h2_graph_create).
extern "C" int32_t h2_bm_max_width(int32_t splits);
extern "C" int h2_bm_pack_x_f32_armed(int32_t n_cols, int32_t d, int32_t splits, const float *X, int64_t ldx,
                                      const float *dinv_col, void *xpack, size_t xpack_bytes, h2_stream_t s);
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static void graph_scratch_sizes(const h2_graph *g, int32_t d, size_t *xpack_bytes, size_t *partial_bytes) {
    *xpack_bytes = *partial_bytes = 0;
    if (!g->n_bm || d <= 0) return;
    const int32_t dw = std::min(d, h2_bm_max_width(g->splits));     // wider rounds are computed in column slices
    *xpack_bytes = h2_bm_xpack_bytes(g->n_cols, dw, g->splits);
    size_t pbytes = 0;
    // partial-slot counts differ per column-group count: take the max over every width that changes the schedule
    for (int dd = 4; ; dd *= 2) {
        const int32_t w = std::min(dd, dw);
        for (int k = 0; k < g->n_bm; ++k) pbytes = std::max(pbytes, h2_bm_partial_bytes(g->bm_host[g->bm_idx[k]].data(), w, g->splits));
        if (dd >= dw) break;
    }
    *partial_bytes = pbytes;
}

extern "C" size_t h2_graph_workspace_bytes(const h2_graph_t *g, int32_t d_max) {
    if (!g) return 0;
    size_t xb, pb;
    graph_scratch_sizes(g, d_max, &xb, &pb);
    return xb || pb ? align256(xb) + align256(pb) + 256 : 0;
}

extern "C" int h2_graph_bind_workspace(h2_graph_t *g, int32_t d_max, void *ws, size_t ws_bytes) {
    H2_REQUIRE(g && d_max >= 4 && d_max % 4 == 0, H2_ERR_INVALID, "h2_graph_bind_workspace: bad argument (d_max=%d)", d_max);
    const size_t need = h2_graph_workspace_bytes(g, d_max);
    H2_REQUIRE(!need || (ws && ws_bytes >= need && aligned16(ws)), H2_ERR_WORKSPACE,
               "h2_graph_bind_workspace: %zu bytes needed for d_max=%d, got %zu", need, d_max, ws_bytes);
    size_t xb, pb;
    graph_scratch_sizes(g, d_max, &xb, &pb);
    char *base = (char *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    if (need && xb) H2_CUDA(cudaMemset(base, 0, 256));   // arms the pack kernel's grid-barrier counters (set-up call: may block)
    g->xpack = need ? base : nullptr;
    g->xpack_bytes = xb;
    g->partial = need ? base + align256(xb) : nullptr;
    g->partial_bytes = pb;
    g->d_max = d_max;
    return H2_OK;
}

// SYNCHRONISES (the old scratch may still be in use) and allocates: call it outside the hot loop.
extern "C" int h2_graph_reserve(h2_graph_t *g, int32_t d_max) {
    H2_REQUIRE(g && d_max >= 4 && d_max % 4 == 0, H2_ERR_INVALID, "h2_graph_reserve: bad argument (d_max=%d)", d_max);
    if (d_max <= g->d_max && (g->xpack || !g->n_bm) && (!g->own_xy || g->x_dev)) return H2_OK;
    H2_CUDA(cudaDeviceSynchronize());
    if (g->ws_owned) { cudaFree(g->ws_owned); g->ws_owned = nullptr; g->xpack = g->partial = nullptr; }
    const size_t need = h2_graph_workspace_bytes(g, d_max);
    if (need) H2_CUDA(cudaMalloc(&g->ws_owned, need));
    int rc = h2_graph_bind_workspace(g, d_max, g->ws_owned, need);
    if (rc != H2_OK) return rc;
    if (g->own_xy) {
        if (g->x_dev) { cudaFree(g->x_dev); g->x_dev = nullptr; }
        if (g->y_dev) { cudaFree(g->y_dev); g->y_dev = nullptr; }
        H2_CUDA(cudaMalloc((void **)&g->x_dev, (size_t)g->n_cols * d_max * 4 + 16));
        H2_CUDA(cudaMalloc((void **)&g->y_dev, (size_t)g->n_rows * g->n_hops * d_max * 4 + 16));
    }
    return H2_OK;
}
static int graph_reserve(h2_graph *g, int32_t d) { return h2_graph_reserve(g, d); }

static bool splits_ok(int32_t s) { return s == 2 || s == 3 || s == H2_SPLITS_I8X2 || s == H2_SPLITS_I8X3; }
static bool splits_is_i8(int32_t s) { return s == H2_SPLITS_I8X2 || s == H2_SPLITS_I8X3; }

static bool pick_bitmap(int32_t mode, int32_t splits, bool has_dinv, int64_t nnz, int32_t n_rows, int32_t n_cols) {
    const double density = (n_rows && n_cols) ? (double)nnz / ((double)n_rows * n_cols) : 0.0;
    // measured crossover on B200 (d = 128): a 256x64 unit costs ~6.9 ns on the tensor cores, a CSR entry ~41 ps of
    // gather => the bitmap wins above ~170 entries per unit, i.e. ~1 % density
    return has_dinv && nnz > 0 && mode != 1 && (mode == 2 || (density >= 0.01 && n_rows >= 128));
}

extern "C" int h2_graph_create(int32_t n_rows, int32_t n_cols, int32_t n_hops, const int64_t *const *rowptr_host,
                               const int32_t *const *col_host, const float *const *val_host,
                               const float *const *dinv_host, int32_t row_begin, int32_t d_max, int32_t mode,
                               int32_t splits, h2_graph_t **out) {
    H2_REQUIRE(out && n_rows >= 0 && n_cols >= 0 && n_hops >= 1 && n_hops <= H2_MAX_HOPS && d_max >= 4 && d_max % 4 == 0,
               H2_ERR_INVALID, "h2_graph_create: n_rows=%d n_cols=%d n_hops=%d d_max=%d", n_rows, n_cols, n_hops, d_max);
    H2_REQUIRE(rowptr_host && col_host && val_host && mode >= 0 && mode <= 2 && splits_ok(splits) &&
               row_begin >= 0, H2_ERR_INVALID, "h2_graph_create: null argument / bad mode");
    h2_graph *g = new h2_graph();
    g->n_rows = n_rows; g->n_cols = n_cols; g->n_hops = n_hops; g->d_max = 0; g->splits = splits; g->row_begin = row_begin;
    g->own_xy = true;
    int rc = H2_OK;
    auto fail = [&](int code) { h2_graph_destroy(g); return code; };
    bool want_bitmap[H2_MAX_HOPS];
    for (int h = 0; h < n_hops; ++h) {
        if (!rowptr_host[h]) { set_error("h2_graph_create: hop %d has no rowptr", h); return fail(H2_ERR_INVALID); }
        const int64_t nnz = rowptr_host[h][n_rows] - rowptr_host[h][0];
        void *rp = nullptr, *c = nullptr, *v = nullptr, *dv = nullptr;
        if ((rc = dev_alloc(g, &rp, (size_t)(n_rows + 1) * 8))) return fail(rc);
        if ((rc = dev_alloc(g, &c, (size_t)nnz * 4))) return fail(rc);
        cudaError_t e = cudaMemcpy(rp, rowptr_host[h], (size_t)(n_rows + 1) * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && nnz) e = cudaMemcpy(c, col_host[h], (size_t)nnz * 4, cudaMemcpyHostToDevice);
        const bool has_dinv = dinv_host && dinv_host[h];
        if (e == cudaSuccess && has_dinv) {
            if ((rc = dev_alloc(g, &dv, (size_t)n_cols * 4))) return fail(rc);
            e = cudaMemcpy(dv, dinv_host[h], (size_t)n_cols * 4, cudaMemcpyHostToDevice);
        }
        g->dinv[h] = (const float *)dv;
        want_bitmap[h] = pick_bitmap(mode, splits, has_dinv, nnz, n_rows, n_cols);
        if (!want_bitmap[h]) {
            if (!val_host[h] && nnz) { set_error("h2_graph_create: hop %d needs explicit values", h); return fail(H2_ERR_INVALID); }
            if ((rc = dev_alloc(g, &v, (size_t)nnz * 4))) return fail(rc);
            if (e == cudaSuccess && nnz) e = cudaMemcpy(v, val_host[h], (size_t)nnz * 4, cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) return fail(cuda_fail(e, "h2_graph_create: upload"));
        g->hops[h] = h2_hop_t{(const int64_t *)rp, (const int32_t *)c, (const float *)v, nullptr, nullptr, 0, 0};
    }
    if ((rc = graph_finish(g, want_bitmap))) return fail(rc);
    if ((rc = graph_reserve(g, d_max))) return fail(rc);
    *out = g;
    return H2_OK;
}

// Same handle over CSR arrays that ALREADY live on the device (no copies; the caller keeps them alive).  hops[h].val
// may be NULL when hops[h].dinv / dinv_row are given (factored CSR); a hop with dinv can take the tensor-core format.
extern "C" int h2_graph_create_device(int32_t n_rows, int32_t n_cols, int32_t n_hops, const h2_hop_t *hops,
                                      const int64_t *nnz_host, int32_t row_begin, int32_t mode, int32_t splits,
                                      h2_graph_t **out) {
    H2_REQUIRE(out && hops && nnz_host && n_rows >= 0 && n_cols >= 0 && n_hops >= 1 && n_hops <= H2_MAX_HOPS && mode >= 0 &&
               mode <= 2 && splits_ok(splits) && row_begin >= 0, H2_ERR_INVALID,
               "h2_graph_create_device: bad argument (n_hops=%d mode=%d splits=%d)", n_hops, mode, splits);
    h2_graph *g = new h2_graph();
    g->n_rows = n_rows; g->n_cols = n_cols; g->n_hops = n_hops; g->d_max = 0; g->splits = splits; g->row_begin = row_begin;
    g->own_xy = false;
    bool want_bitmap[H2_MAX_HOPS];
    for (int h = 0; h < n_hops; ++h) {
        if (!hops[h].rowptr || (!hops[h].val && !hops[h].dinv)) {
            set_error("h2_graph_create_device: hop %d needs rowptr and (val or dinv)", h);
            h2_graph_destroy(g);
            return H2_ERR_INVALID;
        }
        g->hops[h] = hops[h];
        g->dinv[h] = hops[h].dinv;
        want_bitmap[h] = pick_bitmap(mode, splits, hops[h].dinv != nullptr, nnz_host[h], n_rows, n_cols);
        if (!want_bitmap[h] && hops[h].val) { g->hops[h].dinv = nullptr; g->hops[h].dinv_row = nullptr; }
        if (!want_bitmap[h] && !hops[h].val) g->hops[h].dinv_row = hops[h].dinv + row_begin;
    }
    int rc = graph_finish(g, want_bitmap);
    if (rc) { h2_graph_destroy(g); return rc; }
    *out = g;
    return H2_OK;
}

// which format each hop got: fmt_out[h] = 0 (CSR) / 1 (tile bitmap)
extern "C" int h2_graph_formats(const h2_graph_t *g, int32_t *fmt_out) {
    H2_REQUIRE(g && fmt_out, H2_ERR_INVALID, "h2_graph_formats: null argument");
    for (int h = 0; h < g->n_hops; ++h) fmt_out[h] = 0;
    for (int k = 0; k < g->n_bm; ++k) fmt_out[g->bm_idx[k]] = 1;
    return H2_OK;
}

// y_host != nullptr: every hop's column block is copied back to the host buffer (same layout as Y) as soon as that hop
// is done, so the write-back of the CSR hops overlaps the tensor-core hops.
namespace h2 {
int bm_pack_parts(int32_t n_cols, int32_t d, int32_t splits, int32_t n_parts, const float *const *ptrs, const int64_t *bounds,
                  int64_t ld, const float *dinv_col, void *xpack, size_t xpack_bytes, float *xfull, int64_t ld_full,
                  h2_stream_t s, bool zero_header, bool x_bf16);
int bm_spmm_bf16_out(const void *bm_host, const void *bm_dev, int32_t d, int32_t splits, const void *xpack, const float *dinv_row,
                     void *Y, int64_t ldy, int64_t out_col_off, void *partial_ws, size_t partial_bytes, h2_stream_t s);
int gather_rows(int32_t n_cols, int32_t d, int32_t n_parts, const float *const *ptrs, const int64_t *bounds, int64_t ld,
                float *xfull, int64_t ld_full, h2_stream_t s, bool x_bf16);
}

struct RoundParts {   // the round input as row shards (device / peer pointers) + the scratch for its gathered fp32 copy
    int32_t n_parts;
    const float *const *ptrs;
    const int64_t *bounds;
    int64_t ld;
    float *xfull;
    int64_t ld_full;
};

static int graph_round_impl(h2_graph_t *g, int32_t d, const float *X, int64_t ldx, float *Y, int64_t ldy,
                            const int64_t *offsets, float *y_host, h2_stream_t s, const int64_t *x_offsets = nullptr,
                            const RoundParts *parts = nullptr, int32_t x_dtype = H2_F32, int32_t y_dtype = H2_F32) {
    cudaStream_t st = (cudaStream_t)s;
    const bool xb = x_dtype == H2_BF16, yb = y_dtype == H2_BF16;   // bf16 rows: ldx / ldy / offsets count bf16 elements
    H2_REQUIRE(!(xb || yb) || (!x_offsets && !y_host && d % 8 == 0), H2_ERR_UNSUPPORTED,
               "h2_graph_round_ex: bf16 rows are supported for forward device rounds (whole or row-sharded input) with d %% 8 == 0");
    H2_REQUIRE(g && (X || parts) && Y && offsets && d >= 4 && d % 4 == 0, H2_ERR_INVALID, "h2_graph_round: bad argument (d=%d)", d);
    int rc = H2_OK;
    H2_REQUIRE(!g->n_bm || (d <= g->d_max && g->xpack), H2_ERR_WORKSPACE,
               "h2_graph_round: width d=%d exceeds the reserved scratch (d_max=%d): call h2_graph_bind_workspace / "
               "h2_graph_reserve first (the round entry points do not allocate)", d, g->d_max);
    // The scratch (packed operand, partial tiles) is per handle: a round on another stream waits for the previous one.
    // (Not while the stream is being captured into a CUDA graph: a graph is ordered by construction and may not wait on
    // an event recorded outside the capture.)
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    H2_CUDA(cudaStreamIsCapturing(st, &cap));
    const bool guard = cap == cudaStreamCaptureStatusNone;
    if (guard) H2_CUDA(cudaStreamWaitEvent(st, g->ev_done, 0));
    const bool two = g->n_csr && g->n_bm;
    const int32_t max_w = g->n_bm ? h2_bm_max_width(g->splits) : d;   // widest column slice one tensor-core launch covers
    int first_bm = 0;
    if (parts) {
        // Row-sharded input: the all-gather is fused into the first consumer.  With tensor-core hops the pack kernel
        // reads the peers' shards directly and also leaves the gathered fp32 copy the CSR hops need; it runs BEFORE
        // the fork so that both streams see its output.  Without tensor-core hops a plain gather kernel does it.
        H2_REQUIRE(!x_offsets, H2_ERR_INVALID, "h2_graph_round: row shards and per-hop input offsets cannot be combined");
        H2_REQUIRE(!g->n_csr || parts->xfull, H2_ERR_INVALID, "h2_graph_round_parts: CSR hops need the x_full scratch");
        H2_REQUIRE(!(g->n_bm > 1 || d > max_w) || parts->xfull, H2_ERR_INVALID, "h2_graph_round_parts: several tensor hops / column slices need x_full");
        if (g->n_bm && d <= max_w) {
            const int h = g->bm_idx[0];
            rc = bm_pack_parts(g->n_cols, d, g->splits, parts->n_parts, parts->ptrs, parts->bounds, parts->ld, g->dinv[h],
                               g->xpack, g->xpack_bytes, g->n_csr || g->n_bm > 1 ? parts->xfull : nullptr, parts->ld_full, s, false, xb);
            if (rc != H2_OK) return rc;
            first_bm = 1;
        } else {
            rc = gather_rows(g->n_cols, d, parts->n_parts, parts->ptrs, parts->bounds, parts->ld, parts->xfull,
                             parts->ld_full, s, xb);
            if (rc != H2_OK) return rc;
        }
        X = parts->xfull;
        ldx = parts->ld_full;
    }
    cudaStream_t bm_stream = st;
    if (two) {   // fork: tensor-core hops on the high-priority stream, CSR hops stay on the caller's stream
        H2_CUDA(cudaEventRecord(g->ev_fork, st));
        H2_CUDA(cudaStreamWaitEvent(g->side, g->ev_fork, 0));
        bm_stream = g->side;
    }
    for (int k = 0; k < g->n_bm; ++k) {
        const int h = g->bm_idx[k];
        // rounds wider than one launch covers (8 column groups) are computed in column slices of the same buffers
        for (int32_t c0 = 0; c0 < d; c0 += max_w) {
            const int32_t w = std::min(max_w, d - c0);
            if (!(k == 0 && first_bm)) {
                if (xb) {
                    const float *xs = (const float *)((const uint16_t *)X + c0);
                    const int64_t b2[2] = {0, g->n_cols};
                    rc = bm_pack_parts(g->n_cols, w, g->splits, 1, &xs, b2, ldx, g->dinv[h], g->xpack, g->xpack_bytes, nullptr, 0,
                                       (h2_stream_t)bm_stream, false, true);
                } else {
                    rc = h2_bm_pack_x_f32_armed(g->n_cols, w, g->splits, X + (x_offsets ? x_offsets[h] : 0) + c0, ldx, g->dinv[h], g->xpack,
                                                g->xpack_bytes, (h2_stream_t)bm_stream);
                }
                if (rc != H2_OK) return rc;
            }
            if (yb) rc = bm_spmm_bf16_out(g->bm_host[h].data(), g->bm_dev[h], w, g->splits, g->xpack, g->dinv[h] + g->row_begin, Y, ldy,
                                          offsets[h] + c0, g->partial, g->partial_bytes, (h2_stream_t)bm_stream);
            else rc = h2_bm_spmm_f32(g->bm_host[h].data(), g->bm_dev[h], w, g->splits, g->xpack, g->dinv[h] + g->row_begin, Y, ldy,
                                     offsets[h] + c0, g->partial, g->partial_bytes, (h2_stream_t)bm_stream);
            if (rc != H2_OK) return rc;
        }
    }
    if (g->n_csr) {
        h2_hop_t sub[H2_MAX_HOPS];
        for (int k = 0; k < g->n_csr; ++k) {
            sub[k] = g->hops[g->csr_idx[k]];
            sub[k].out_col_off = offsets[g->csr_idx[k]];
            sub[k].in_col_off = x_offsets ? x_offsets[g->csr_idx[k]] : 0;
        }
        if (xb || yb) rc = h2_fused_hops_spmm_ex(g->plan_host.data(), g->plan_dev, g->n_rows, g->n_csr, sub, d, X, ldx, x_dtype, Y, ldy, y_dtype, s);
        else rc = h2_fused_hops_spmm_f32(g->plan_host.data(), g->plan_dev, g->n_rows, g->n_csr, sub, d, X, ldx, Y, ldy, s);
        if (rc != H2_OK) return rc;
        if (y_host)
            for (int k = 0; k < g->n_csr; ++k) {
                const int64_t off = offsets[g->csr_idx[k]];
                H2_CUDA(cudaMemcpy2DAsync(y_host + off, (size_t)ldy * 4, Y + off, (size_t)ldy * 4, (size_t)d * 4,
                                          (size_t)g->n_rows, cudaMemcpyDeviceToHost, st));
            }
    }
    if (two) {
        H2_CUDA(cudaEventRecord(g->ev_join, g->side));
        H2_CUDA(cudaStreamWaitEvent(st, g->ev_join, 0));
    }
    if (y_host)
        for (int k = 0; k < g->n_bm; ++k) {
            const int64_t off = offsets[g->bm_idx[k]];
            H2_CUDA(cudaMemcpy2DAsync(y_host + off, (size_t)ldy * 4, Y + off, (size_t)ldy * 4, (size_t)d * 4,
                                      (size_t)g->n_rows, cudaMemcpyDeviceToHost, st));
        }
    if (guard) H2_CUDA(cudaEventRecord(g->ev_done, st));
    return H2_OK;
}

extern "C" int h2_graph_round_parts(h2_graph_t *g, int32_t d, int32_t n_parts, const float *const *part_ptrs_host,
                                    const int64_t *bounds_host, int64_t ld_part, float *x_full, int64_t ld_full, float *Y,
                                    int64_t ldy, const int64_t *y_offsets_host, h2_stream_t s) {
    RoundParts parts{n_parts, part_ptrs_host, bounds_host, ld_part, x_full, ld_full};
    return graph_round_impl(g, d, nullptr, 0, Y, ldy, y_offsets_host, nullptr, s, nullptr, &parts);
}

extern "C" int h2_graph_round_parts_ex(h2_graph_t *g, int32_t d, int32_t n_parts, const void *const *part_ptrs_host,
                                       const int64_t *bounds_host, int64_t ld_part, int32_t x_dtype, void *x_full, int64_t ld_full,
                                       void *Y, int64_t ldy, int32_t y_dtype, const int64_t *y_offsets_host, h2_stream_t s) {
    H2_REQUIRE((x_dtype == H2_F32 || x_dtype == H2_BF16) && (y_dtype == H2_F32 || y_dtype == H2_BF16), H2_ERR_INVALID,
               "h2_graph_round_parts_ex: dtype codes are H2_F32 / H2_BF16");
    H2_REQUIRE(!g || !g->n_bm || splits_is_i8(g->splits) || (x_dtype == H2_F32 && y_dtype == H2_F32), H2_ERR_UNSUPPORTED,
               "h2_graph_round_parts_ex: bf16 rows on the tensor-core hops need the int8 digits");
    RoundParts parts{n_parts, (const float *const *)part_ptrs_host, bounds_host, ld_part, (float *)x_full, ld_full};
    return graph_round_impl(g, d, nullptr, 0, (float *)Y, ldy, y_offsets_host, nullptr, s, nullptr, &parts, x_dtype, y_dtype);
}

extern "C" int h2_graph_round(h2_graph_t *g, int32_t d, const float *X, int64_t ldx, float *Y, int64_t ldy,
                              const int64_t *offsets, h2_stream_t s) {
    return graph_round_impl(g, d, X, ldx, Y, ldy, offsets, nullptr, s);
}

extern "C" int h2_graph_round_ex(h2_graph_t *g, int32_t d, const void *X, int64_t ldx, int32_t x_dtype, void *Y, int64_t ldy,
                                 int32_t y_dtype, const int64_t *offsets, h2_stream_t s) {
    H2_REQUIRE((x_dtype == H2_F32 || x_dtype == H2_BF16) && (y_dtype == H2_F32 || y_dtype == H2_BF16), H2_ERR_INVALID,
               "h2_graph_round_ex: dtype codes are H2_F32 / H2_BF16");
    H2_REQUIRE(!g || !g->n_bm || splits_is_i8(g->splits) || (x_dtype == H2_F32 && y_dtype == H2_F32), H2_ERR_UNSUPPORTED,
               "h2_graph_round_ex: bf16 rows on the tensor-core hops need the int8 digits");
    return graph_round_impl(g, d, (const float *)X, ldx, (float *)Y, ldy, offsets, nullptr, s, nullptr, nullptr, x_dtype, y_dtype);
}

extern "C" int h2_graph_round_multi(h2_graph_t *g, int32_t d, const float *X, int64_t ldx, const int64_t *x_offsets, float *Y,
                                    int64_t ldy, const int64_t *y_offsets, h2_stream_t s) {
    H2_REQUIRE(x_offsets, H2_ERR_INVALID, "h2_graph_round_multi: null x_offsets");
    if (g) for (int h = 0; h < g->n_hops; ++h)
        H2_REQUIRE(x_offsets[h] % 4 == 0 && x_offsets[h] >= 0 && x_offsets[h] + d <= ldx, H2_ERR_ALIGN,
                   "h2_graph_round_multi: x_offsets[%d]=%lld", h, (long long)x_offsets[h]);
    return graph_round_impl(g, d, X, ldx, Y, ldy, y_offsets, nullptr, s, x_offsets);
}

extern "C" int h2_graph_round_host(h2_graph_t *g, int32_t d, const float *x_host, float *y_host, h2_stream_t s) {
    cudaStream_t st = (cudaStream_t)s;
    H2_REQUIRE(g && x_host && y_host && d >= 4 && d % 4 == 0 && g->own_xy, H2_ERR_INVALID,
               "h2_graph_round_host: bad argument (d=%d) or handle created over device arrays", d);
    H2_REQUIRE(d <= g->d_max, H2_ERR_WORKSPACE, "h2_graph_round_host: d=%d exceeds the d_max=%d the handle was created with "
               "(h2_graph_reserve)", d, g->d_max);
    int64_t offsets[H2_MAX_HOPS];
    for (int h = 0; h < g->n_hops; ++h) offsets[h] = (int64_t)h * d;  // GCNLayer + Flatten layout: [N, H*d]
    const int64_t ldy = (int64_t)g->n_hops * d;
    H2_CUDA(cudaMemcpyAsync(g->x_dev, x_host, (size_t)g->n_cols * d * 4, cudaMemcpyHostToDevice, st));
    static const bool contiguous = getenv("H2_E2E_CONTIGUOUS_D2H") != nullptr;   // measurement switch
    // Measurement switch H2_E2E_ZEROCOPY=1: with a pinned (device-mapped) result buffer the kernels store Y straight into it
    // over PCIe instead of the per-hop copies below.  Measured r02 (north-star point): 0.359 ms per call against 0.330 ms
    // with the copies — the epilogues stall on the posted writes while they hold the SMs — so the copies stay the default.
    static const bool zerocopy = getenv("H2_E2E_ZEROCOPY") && atoi(getenv("H2_E2E_ZEROCOPY")) != 0;
    if (zerocopy && !contiguous) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, y_host) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer &&
            aligned16(attr.devicePointer)) {
            int rc0 = graph_round_impl(g, d, g->x_dev, d, (float *)attr.devicePointer, ldy, offsets, nullptr, s);
            if (rc0 != H2_OK) return rc0;
            H2_CUDA(cudaStreamSynchronize(st));
            return H2_OK;
        }
        cudaGetLastError();   // pageable memory: not an error, take the staged path
    }
    int rc = graph_round_impl(g, d, g->x_dev, d, g->y_dev, ldy, offsets, contiguous ? nullptr : y_host, s);
    if (rc != H2_OK) return rc;
    if (contiguous) H2_CUDA(cudaMemcpyAsync(y_host, g->y_dev, (size_t)g->n_rows * ldy * 4, cudaMemcpyDeviceToHost, st));
    H2_CUDA(cudaStreamSynchronize(st));
    return H2_OK;
}
