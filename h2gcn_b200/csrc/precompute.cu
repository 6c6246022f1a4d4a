// Adjacency-power precompute on the GPU: removeEye, exact-distance-2 pattern,
// symmetric / random-walk normalisation.  Integer work is bit-exact against the reference; the fp32 adjacency
// values are produced in fp64 and rounded once, like scipy-fp64 -> astype(float32) in the reference
// (h2gcn/datasets/_dataset.py:109-124, 132-158, 528-535).
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace h2 {

// ---- removeEye ---------------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void remove_eye_kernel(int32_t n, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                  const float *__restrict__ val_in, int64_t *__restrict__ rowcount,
                                  const int64_t *__restrict__ rowptr_out, int32_t *__restrict__ col_out,
                                  float *__restrict__ val_out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    int64_t kept = 0;
    for (int64_t base = s; base < e; base += 32) {
        const int64_t k = base + lane;
        const int c = k < e ? col[k] : -1;
        const bool keep = (k < e) && (c != (int)row);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (FILL && keep) {
            const int64_t dst = rowptr_out[row] + kept + __popc(m & ((1u << lane) - 1));
            col_out[dst] = c;
            if (val_in && val_out) val_out[dst] = val_in[k];
        }
        kept += __popc(m);
    }
    if (!FILL && lane == 0) rowcount[row] = kept;
}

// ---- exact-distance-2 pattern ------------------------------------------------------------------------------------
// One CTA per output row and column range: a shared-memory bitmap collects the union of the neighbours' neighbour
// lists, then row i itself and its 1-hop neighbours are cleared (the `mt - prev_mt` of nhoodSplit, _dataset.py:157).
// Walking the bitmap in word order yields ascending columns, i.e. tf.sparse.reorder order, for free.
constexpr int kHop2Threads = 256;
constexpr int kHop2MaxWords = 48 * 1024;  // 192 KB bitmap = 1.5 M columns per pass

template <bool FILL>
__global__ void __launch_bounds__(kHop2Threads) hop2_kernel(int32_t n, const int64_t *__restrict__ rowptr,
                                                           const int32_t *__restrict__ col, int32_t row_begin,
                                                           int32_t row_end, int words_per_pass,
                                                           int64_t *__restrict__ rowcount2,
                                                           const int64_t *__restrict__ rowptr2,
                                                           int32_t *__restrict__ col2) {
    extern __shared__ uint32_t s_bits[];  // [words_per_pass]
    __shared__ int s_scan[kHop2Threads];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kWarps = kHop2Threads / 32;
    const int wpt = (words_per_pass + kHop2Threads - 1) / kHop2Threads;  // consecutive words per thread

    for (int64_t row = row_begin + (int64_t)blockIdx.x; row < row_end; row += gridDim.x) {
        const int i = (int)row;
        const int64_t s = rowptr[i], e = rowptr[i + 1];
        int64_t written = 0;
        for (int64_t c0 = 0; c0 < n; c0 += (int64_t)words_per_pass * 32) {
            const int64_t c1 = min((int64_t)n, c0 + (int64_t)words_per_pass * 32);
            for (int w = tid; w < words_per_pass; w += kHop2Threads) s_bits[w] = 0;
            __syncthreads();
            for (int64_t k = s + warp; k < e; k += kWarps) {
                const int j = col[k];
                const int64_t js = rowptr[j], je = rowptr[j + 1];
                for (int64_t q = js + lane; q < je; q += 32) {
                    const int t = col[q];
                    if (t >= c0 && t < c1) atomicOr(&s_bits[(t - c0) >> 5], 1u << (t & 31));
                }
            }
            __syncthreads();
            // distance 0 and distance 1 are not distance 2
            if (tid == 0 && i >= c0 && i < c1) atomicAnd(&s_bits[(i - c0) >> 5], ~(1u << (i & 31)));
            for (int64_t k = s + tid; k < e; k += kHop2Threads) {
                const int t = col[k];
                if (t >= c0 && t < c1) atomicAnd(&s_bits[(t - c0) >> 5], ~(1u << (t & 31)));
            }
            __syncthreads();
            const int w0 = tid * wpt, w1 = min(words_per_pass, w0 + wpt);
            int mine = 0;
            for (int w = w0; w < w1; ++w) mine += __popc(s_bits[w]);
            // block exclusive scan of `mine`
            s_scan[tid] = mine;
            __syncthreads();
            for (int off = 1; off < kHop2Threads; off <<= 1) {
                const int add = tid >= off ? s_scan[tid - off] : 0;
                __syncthreads();
                s_scan[tid] += add;
                __syncthreads();
            }
            const int incl = s_scan[tid];
            const int total = s_scan[kHop2Threads - 1];
            if (FILL) {
                int32_t *dst = col2 + rowptr2[i - row_begin] + written + (incl - mine);
                for (int w = w0; w < w1; ++w) {
                    uint32_t bits = s_bits[w];
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        *dst++ = (int32_t)(c0 + (int64_t)w * 32 + b);
                    }
                }
            }
            written += total;
            __syncthreads();
        }
        if (!FILL && tid == 0) rowcount2[i - row_begin] = written;
    }
}

// ---- normalisation -----------------------------------------------------------------------------------------------
__global__ void dinv_kernel(int32_t n_cols, const int64_t *__restrict__ rowptr, const int64_t *__restrict__ deg_all,
                            double *__restrict__ dinv64, float *__restrict__ dinv32) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= n_cols) return;
    const int64_t deg = deg_all ? deg_all[j] : rowptr[j + 1] - rowptr[j];
    // np.power(deg, -0.5) with inf -> 0 (_dataset.py:115-116): the zero-degree mask
    const double r = deg > 0 ? 1.0 / sqrt((double)deg) : 0.0;
    if (dinv64) dinv64[j] = r;
    if (dinv32) dinv32[j] = (float)r;
}

__global__ void sym_vals_kernel(int32_t n_rows, int32_t row_begin, const int64_t *__restrict__ rowptr,
                                const int32_t *__restrict__ col, const double *__restrict__ dinv64,
                                float *__restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const double di = dinv64[row_begin + row] * 1.0;  // (DInvSqrt @ adj) first: dinv_i * a_ij, a_ij == 1
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    for (int64_t k = s + lane; k < e; k += 32) val[k] = (float)(di * dinv64[col[k]]);
}

__global__ void rw_vals_kernel(int32_t n_rows, const int64_t *__restrict__ rowptr, float *__restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    const float v = e > s ? (float)(1.0 / (double)(e - s)) : 0.f;
    for (int64_t k = s + lane; k < e; k += 32) val[k] = v;
}

__global__ void validate_kernel(int32_t n_rows, int32_t n_cols, const int64_t *__restrict__ rowptr,
                                const int32_t *__restrict__ col, int *__restrict__ bad) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int64_t s = rowptr[row], e = rowptr[row + 1];
    if (e < s) { atomicExch(bad, 1); return; }
    for (int64_t k = s + lane; k < e; k += 32) {
        const int c = col[k];
        if (c < 0 || c >= n_cols || (k > s && col[k - 1] >= c)) atomicExch(bad, 1);
    }
}

static unsigned warp_row_grid(int64_t rows, int tpb) { return (unsigned)((rows * 32 + tpb - 1) / tpb); }

}  // namespace h2

using namespace h2;

extern "C" int h2_remove_eye_count(int32_t n, const int64_t *rowptr, const int32_t *col, int64_t *rowcount_out,
                                   h2_stream_t s) {
    H2_REQUIRE(n >= 0 && (n == 0 || (rowptr && rowcount_out)), H2_ERR_INVALID, "h2_remove_eye_count: bad argument");
    if (n == 0) return H2_OK;
    remove_eye_kernel<false><<<warp_row_grid(n, 256), 256, 0, (cudaStream_t)s>>>(n, rowptr, col, nullptr, rowcount_out, nullptr, nullptr, nullptr);
    H2_LAUNCHED("remove_eye_kernel<count>");
    return H2_OK;
}

extern "C" int h2_remove_eye_fill(int32_t n, const int64_t *rowptr, const int32_t *col, const float *val_in,
                                  const int64_t *rowptr_out, int32_t *col_out, float *val_out, h2_stream_t s) {
    H2_REQUIRE(n >= 0 && (n == 0 || (rowptr && rowptr_out)), H2_ERR_INVALID, "h2_remove_eye_fill: bad argument");
    if (n == 0) return H2_OK;
    remove_eye_kernel<true><<<warp_row_grid(n, 256), 256, 0, (cudaStream_t)s>>>(n, rowptr, col, val_in, nullptr, rowptr_out, col_out, val_out);
    H2_LAUNCHED("remove_eye_kernel<fill>");
    return H2_OK;
}

extern "C" size_t h2_scan_workspace_bytes(int64_t n) {
    // cub::DeviceScan temp storage is O(n / tile) descriptors; bound it generously without touching the device.
    return (size_t)((n > 0 ? n : 0) / 128 + 1) * 16 + (1u << 16);
}

extern "C" int h2_exclusive_scan_i64(int64_t n, const int64_t *counts, int64_t *rowptr, void *ws, size_t ws_bytes,
                                     h2_stream_t s) {
    cudaStream_t st = (cudaStream_t)s;
    H2_REQUIRE(n >= 0 && rowptr && (n == 0 || counts), H2_ERR_INVALID, "h2_exclusive_scan_i64: bad argument");
    H2_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int64_t), st));
    if (n == 0) return H2_OK;
    H2_REQUIRE(n < 0x7fffffffLL, H2_ERR_UNSUPPORTED, "h2_exclusive_scan_i64: n=%lld too large", (long long)n);
    size_t need = 0;
    H2_CUDA(cub::DeviceScan::InclusiveSum(nullptr, need, counts, rowptr + 1, (int)n, st));
    H2_REQUIRE(ws && need <= ws_bytes, H2_ERR_WORKSPACE, "h2_exclusive_scan_i64: need %zu bytes, have %zu", need, ws_bytes);
    H2_CUDA(cub::DeviceScan::InclusiveSum(ws, need, counts, rowptr + 1, (int)n, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return H2_OK;
}

template <bool FILL>
static int hop2_launch(int32_t n, const int64_t *rowptr, const int32_t *col, int32_t row_begin, int32_t row_end,
                       int64_t *rowcount2, const int64_t *rowptr2, int32_t *col2, cudaStream_t st) {
    H2_REQUIRE(n >= 0 && row_begin >= 0 && row_end >= row_begin && row_end <= n, H2_ERR_INVALID,
               "h2_hop2: n=%d rows [%d,%d)", n, row_begin, row_end);
    if (row_end == row_begin) return H2_OK;
    H2_REQUIRE(rowptr && (FILL ? rowptr2 != nullptr : rowcount2 != nullptr), H2_ERR_INVALID, "h2_hop2: null argument");
    const int64_t words_all = ((int64_t)n + 31) / 32;
    const int words = (int)(words_all < kHop2MaxWords ? words_all : kHop2MaxWords);
    const size_t smem = (size_t)words * 4;
    auto kern = hop2_kernel<FILL>;
    H2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kHop2MaxWords * 4)));
    const int64_t rows = (int64_t)row_end - row_begin;
    // persistent-ish: at most 16 CTAs per SM worth of blocks, each striding over rows
    const int64_t cap = (int64_t)kNumSms * (smem > 96 * 1024 ? 1 : (smem > 24 * 1024 ? 2 : 8));
    const unsigned grid = (unsigned)(rows < cap ? rows : cap);
    kern<<<grid, kHop2Threads, smem, st>>>(n, rowptr, col, row_begin, row_end, words, rowcount2, rowptr2, col2);
    H2_LAUNCHED(FILL ? "hop2_kernel<fill>" : "hop2_kernel<count>");
    return H2_OK;
}

extern "C" int h2_hop2_count(int32_t n, const int64_t *rowptr, const int32_t *col, int32_t row_begin, int32_t row_end,
                             int64_t *rowcount2, h2_stream_t s) {
    return hop2_launch<false>(n, rowptr, col, row_begin, row_end, rowcount2, nullptr, nullptr, (cudaStream_t)s);
}

extern "C" int h2_hop2_fill(int32_t n, const int64_t *rowptr, const int32_t *col, int32_t row_begin, int32_t row_end,
                            const int64_t *rowptr2, int32_t *col2, h2_stream_t s) {
    return hop2_launch<true>(n, rowptr, col, row_begin, row_end, nullptr, rowptr2, col2, (cudaStream_t)s);
}

extern "C" int h2_sym_normalize(int32_t n_rows, int32_t n_cols, int32_t row_begin, const int64_t *rowptr,
                                const int32_t *col, const int64_t *deg_all, double *dinv64, float *dinv32, float *val,
                                h2_stream_t s) {
    cudaStream_t st = (cudaStream_t)s;
    H2_REQUIRE(n_rows >= 0 && n_cols >= 0 && row_begin >= 0 && row_begin + (int64_t)n_rows <= n_cols, H2_ERR_INVALID,
               "h2_sym_normalize: n_rows=%d n_cols=%d row_begin=%d", n_rows, n_cols, row_begin);
    H2_REQUIRE(deg_all || n_rows == n_cols, H2_ERR_INVALID,
               "h2_sym_normalize: deg_all is required when the rows are a shard (n_rows != n_cols)");
    H2_REQUIRE(!val || dinv64, H2_ERR_INVALID, "h2_sym_normalize: val needs the dinv64 [n_cols] buffer");
    if (n_cols == 0) return H2_OK;
    H2_REQUIRE(rowptr || deg_all, H2_ERR_INVALID, "h2_sym_normalize: null rowptr");
    if (dinv64 || dinv32) {
        dinv_kernel<<<(unsigned)((n_cols + 255) / 256), 256, 0, st>>>(n_cols, rowptr, deg_all, dinv64, dinv32);
        H2_LAUNCHED("dinv_kernel");
    }
    if (val && n_rows > 0) {
        H2_REQUIRE(rowptr && col, H2_ERR_INVALID, "h2_sym_normalize: null CSR");
        sym_vals_kernel<<<warp_row_grid(n_rows, 256), 256, 0, st>>>(n_rows, row_begin, rowptr, col, dinv64, val);
        H2_LAUNCHED("sym_vals_kernel");
    }
    return H2_OK;
}

extern "C" int h2_rw_normalize(int32_t n_rows, const int64_t *rowptr, float *val, h2_stream_t s) {
    H2_REQUIRE(n_rows >= 0 && (n_rows == 0 || (rowptr && val)), H2_ERR_INVALID, "h2_rw_normalize: bad argument");
    if (n_rows == 0) return H2_OK;
    rw_vals_kernel<<<warp_row_grid(n_rows, 256), 256, 0, (cudaStream_t)s>>>(n_rows, rowptr, val);
    H2_LAUNCHED("rw_vals_kernel");
    return H2_OK;
}

extern "C" int h2_validate_csr(int32_t n_rows, int32_t n_cols, const int64_t *rowptr, const int32_t *col,
                               int32_t *flag_dev, h2_stream_t s) {
    cudaStream_t st = (cudaStream_t)s;
    H2_REQUIRE(n_rows >= 0 && n_cols >= 0 && (n_rows == 0 || (rowptr && flag_dev)), H2_ERR_INVALID,
               "h2_validate_csr: bad argument");
    if (n_rows == 0) return H2_OK;
    H2_CUDA(cudaMemsetAsync(flag_dev, 0, sizeof(int), st));
    validate_kernel<<<warp_row_grid(n_rows, 256), 256, 0, st>>>(n_rows, n_cols, rowptr, col, flag_dev);
    H2_LAUNCHED("validate_kernel");
    int hbad = 0;
    H2_CUDA(cudaMemcpyAsync(&hbad, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
    H2_CUDA(cudaStreamSynchronize(st));
    H2_REQUIRE(!hbad, H2_ERR_INDEX, "h2_validate_csr: column index out of range or rows not strictly ascending");
    return H2_OK;
}
