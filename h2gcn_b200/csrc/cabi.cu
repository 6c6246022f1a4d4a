// Library-level pieces of the C-ABI: error state, launch counter, and the host-buffer graph handle used by the
// end-to-end entry points (include/h2gcn_b200.h, "end-to-end entry points with HOST buffers").
#include <stdarg.h>
#include <string.h>

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

// ---- host-buffer graph handle ----------------------------------------------------------------------------------
struct h2_graph {
    int32_t n_rows = 0, n_cols = 0, n_hops = 0, d_max = 0;
    std::vector<void *> owned;  // every cudaMalloc'ed pointer, freed in destroy
    h2_hop_t hops[H2_MAX_HOPS];
    std::vector<char> plan_host;
    void *plan_dev = nullptr;
    float *x_dev = nullptr, *y_dev = nullptr;
};

static int dev_alloc(h2_graph *g, void **p, size_t bytes) {
    H2_CUDA(cudaMalloc(p, bytes ? bytes : 16));
    g->owned.push_back(*p);
    return H2_OK;
}

extern "C" int h2_graph_destroy(h2_graph_t *g) {
    if (!g) return H2_OK;
    for (void *p : g->owned) cudaFree(p);
    delete g;
    return H2_OK;
}

extern "C" int h2_graph_create(int32_t n_rows, int32_t n_cols, int32_t n_hops, const int64_t *const *rowptr_host,
                               const int32_t *const *col_host, const float *const *val_host, int32_t d_max,
                               h2_graph_t **out) {
    H2_REQUIRE(out && n_rows >= 0 && n_cols >= 0 && n_hops >= 1 && n_hops <= H2_MAX_HOPS && d_max >= 4 && d_max % 4 == 0,
               H2_ERR_INVALID, "h2_graph_create: n_rows=%d n_cols=%d n_hops=%d d_max=%d", n_rows, n_cols, n_hops, d_max);
    H2_REQUIRE(rowptr_host && col_host && val_host, H2_ERR_INVALID, "h2_graph_create: null argument");
    h2_graph *g = new h2_graph();
    g->n_rows = n_rows; g->n_cols = n_cols; g->n_hops = n_hops; g->d_max = d_max;
    int rc = H2_OK;
    auto fail = [&](int code) { h2_graph_destroy(g); return code; };
    for (int h = 0; h < n_hops; ++h) {
        if (!rowptr_host[h]) { set_error("h2_graph_create: hop %d has no rowptr", h); return fail(H2_ERR_INVALID); }
        const int64_t nnz = rowptr_host[h][n_rows] - rowptr_host[h][0];
        void *rp = nullptr, *c = nullptr, *v = nullptr;
        if ((rc = dev_alloc(g, &rp, (size_t)(n_rows + 1) * 8))) return fail(rc);
        if ((rc = dev_alloc(g, &c, (size_t)nnz * 4))) return fail(rc);
        if ((rc = dev_alloc(g, &v, (size_t)nnz * 4))) return fail(rc);
        cudaError_t e = cudaMemcpy(rp, rowptr_host[h], (size_t)(n_rows + 1) * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && nnz) e = cudaMemcpy(c, col_host[h], (size_t)nnz * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess && nnz) e = cudaMemcpy(v, val_host[h], (size_t)nnz * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return fail(cuda_fail(e, "h2_graph_create: upload"));
        g->hops[h] = h2_hop_t{(const int64_t *)rp, (const int32_t *)c, (const float *)v, nullptr, nullptr, 0};
    }
    g->plan_host.resize(h2_plan_host_bytes());
    void *ws = nullptr;
    const size_t ws_bytes = h2_plan_workspace_bytes(n_rows, n_hops);
    if ((rc = dev_alloc(g, &g->plan_dev, h2_plan_dev_bytes(n_rows, n_hops)))) return fail(rc);
    if ((rc = dev_alloc(g, &ws, ws_bytes))) return fail(rc);
    if ((rc = h2_plan_build(n_rows, n_hops, g->hops, g->plan_host.data(), g->plan_dev, ws, ws_bytes, nullptr)))
        return fail(rc);
    if ((rc = dev_alloc(g, (void **)&g->x_dev, (size_t)n_cols * d_max * 4))) return fail(rc);
    if ((rc = dev_alloc(g, (void **)&g->y_dev, (size_t)n_rows * n_hops * d_max * 4))) return fail(rc);
    *out = g;
    return H2_OK;
}

extern "C" int h2_graph_round_host(h2_graph_t *g, int32_t d, const float *x_host, float *y_host, h2_stream_t s) {
    cudaStream_t st = (cudaStream_t)s;
    H2_REQUIRE(g && x_host && y_host && d >= 4 && d % 4 == 0 && d <= g->d_max, H2_ERR_INVALID,
               "h2_graph_round_host: bad argument (d=%d, d_max=%d)", d, g ? g->d_max : -1);
    h2_hop_t hops[H2_MAX_HOPS];
    for (int h = 0; h < g->n_hops; ++h) {
        hops[h] = g->hops[h];
        hops[h].out_col_off = (int64_t)h * d;  // GCNLayer + Flatten layout: [N, H*d]
    }
    const int64_t ldy = (int64_t)g->n_hops * d;
    H2_CUDA(cudaMemcpyAsync(g->x_dev, x_host, (size_t)g->n_cols * d * 4, cudaMemcpyHostToDevice, st));
    int rc = h2_fused_hops_spmm_f32(g->plan_host.data(), g->plan_dev, g->n_rows, g->n_hops, hops, d, g->x_dev, d,
                                    g->y_dev, ldy, s);
    if (rc != H2_OK) return rc;
    H2_CUDA(cudaMemcpyAsync(y_host, g->y_dev, (size_t)g->n_rows * ldy * 4, cudaMemcpyDeviceToHost, st));
    H2_CUDA(cudaStreamSynchronize(st));
    return H2_OK;
}
