// Dense ends of the forward pass, fp32 SIMT "parity mode":
//   h2_sparse_dense_f32 : SparseDense.call (+ReLU)  (h2gcn/models/_layers.py:45-52, H2GCN.py:269-270)
//   h2_dense_f32        : keras Dense               (H2GCN.py:244-249)
//   h2_relu_slice_f32   : stand-alone ReLU / slice copy for layer strings the planner cannot fuse
// All of them write straight into a column slot of the concat buffer (ld + column offset).
#include "common.cuh"

namespace h2 {

struct HopDev;
struct RoundParams;
int run_gather_round(const RoundParams &p, cudaStream_t st);

// Y[r, off + c] = act(sum_k X[r, k] W[k, c] + b[c]).  One warp per (row, 32-column tile): lane = output column, the
// warp walks k reading X[r, k] as a broadcast and W[k, c0 + lane] coalesced.  k ascending => same summation order
// as the oracle's sequential-k loop (bit-identical up to FMA contraction).
constexpr int kDenseRowsPerCta = 8;

__global__ void __launch_bounds__(kDenseRowsPerCta * 32) dense_kernel(int32_t n_rows, int32_t k_dim, int32_t c_dim,
                                                                      const float *__restrict__ X, int64_t ldx,
                                                                      const float *__restrict__ W,
                                                                      const float *__restrict__ bias, int relu,
                                                                      float *__restrict__ Y, int64_t ldy) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = blockIdx.x * (int64_t)kDenseRowsPerCta + warp;
    const int c = blockIdx.y * 32 + lane;
    if (r >= n_rows) return;
    const float *xr = X + r * ldx;
    float acc = 0.f;
    for (int k0 = 0; k0 < k_dim; k0 += 32) {
        const int kk = k0 + lane;
        const float xv = kk < k_dim ? xr[kk] : 0.f;
        const int lim = min(32, k_dim - k0);
        for (int t = 0; t < lim; ++t) {
            const float x = __shfl_sync(0xffffffffu, xv, t);
            if (c < c_dim) acc = fmaf(x, __ldg(W + (int64_t)(k0 + t) * c_dim + c), acc);
        }
    }
    if (c < c_dim) {
        if (bias) acc += bias[c];
        if (relu) acc = fmaxf(acc, 0.f);
        Y[r * ldy + c] = acc;
    }
}

__global__ void relu_slice_kernel(int32_t n_rows, int32_t d, const float *__restrict__ X, int64_t ldx,
                                  float *__restrict__ Y, int64_t ldy, int relu) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_rows * d) return;
    const int64_t r = idx / d;
    const int c = (int)(idx - r * d);
    float v = X[r * ldx + c];
    if (relu) v = fmaxf(v, 0.f);
    Y[r * ldy + c] = v;
}

__global__ void sum_slices_kernel(int32_t n_rows, int32_t d, int32_t n_slices, const float *__restrict__ T, int64_t ldt,
                                  float *__restrict__ G, int64_t ldg, int accumulate, const float *__restrict__ mask_src,
                                  int64_t ld_mask) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_rows * d) return;
    const int64_t r = idx / d;
    const int c = (int)(idx - r * d);
    float a = accumulate ? G[r * ldg + c] : 0.f;
    for (int s = 0; s < n_slices; ++s) a += T[r * ldt + (int64_t)s * d + c];   // fixed order
    if (mask_src && !(mask_src[r * ld_mask + c] > 0.f)) a = 0.f;
    G[r * ldg + c] = a;
}

}  // namespace h2

using namespace h2;

extern "C" int h2_dense_f32(int32_t n_rows, int32_t k, int32_t c, const float *X, int64_t ldx, const float *W,
                            const float *bias, int32_t relu, float *Y, int64_t ldy, int64_t out_col_off,
                            h2_stream_t s) {
    H2_REQUIRE(n_rows >= 0 && k >= 0 && c >= 1, H2_ERR_INVALID, "h2_dense_f32: n_rows=%d k=%d c=%d", n_rows, k, c);
    if (n_rows == 0) return H2_OK;
    H2_REQUIRE(X && W && Y && ldx >= k && ldy >= out_col_off + c && out_col_off >= 0, H2_ERR_INVALID,
               "h2_dense_f32: null pointer or leading dimension too small");
    dim3 grid((unsigned)((n_rows + kDenseRowsPerCta - 1) / kDenseRowsPerCta), (unsigned)((c + 31) / 32));
    dense_kernel<<<grid, kDenseRowsPerCta * 32, 0, (cudaStream_t)s>>>(n_rows, k, c, X, ldx, W, bias, relu,
                                                                    Y + out_col_off, ldy);
    H2_LAUNCHED("dense_kernel");
    return H2_OK;
}

extern "C" int h2_relu_slice_f32(int32_t n_rows, int32_t d, const float *X, int64_t ldx, float *Y, int64_t ldy,
                                 int32_t relu, h2_stream_t s) {
    H2_REQUIRE(n_rows >= 0 && d >= 0, H2_ERR_INVALID, "h2_relu_slice_f32: n_rows=%d d=%d", n_rows, d);
    const int64_t total = (int64_t)n_rows * d;
    if (total == 0) return H2_OK;
    H2_REQUIRE(X && Y && ldx >= d && ldy >= d, H2_ERR_INVALID, "h2_relu_slice_f32: bad argument");
    relu_slice_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)s>>>(n_rows, d, X, ldx, Y, ldy, relu);
    H2_LAUNCHED("relu_slice_kernel");
    return H2_OK;
}

extern "C" int h2_sum_slices_f32(int32_t n_rows, int32_t d, int32_t n_slices, const float *T, int64_t ldt, float *G,
                                 int64_t ldg, int32_t accumulate, const float *mask_src, int64_t ld_mask, h2_stream_t s) {
    H2_REQUIRE(n_rows >= 0 && d >= 0 && n_slices >= 0, H2_ERR_INVALID, "h2_sum_slices_f32: bad sizes");
    const int64_t total = (int64_t)n_rows * d;
    if (total == 0) return H2_OK;
    H2_REQUIRE(G && (T || n_slices == 0) && ldg >= d && ldt >= (int64_t)n_slices * d, H2_ERR_INVALID, "h2_sum_slices_f32: bad argument");
    sum_slices_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)s>>>(n_rows, d, n_slices, T, ldt, G, ldg, accumulate,
                                                                                   mask_src, ld_mask);
    H2_LAUNCHED("sum_slices_kernel");
    return H2_OK;
}
