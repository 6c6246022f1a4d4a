/* CPU oracle (C part) for the H2GCN aggregation hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this library.
 * Parity status: pinned through oracle/h2gcn_oracle.py (which is checked against golden vectors produced by the
 * reference's own Python, tests/golden/make_golden.py); tests/test_oracle_golden.py checks these C functions against
 * the Python oracle bit for bit.
 *
 * Restates (paths relative to /root/reference/):
 *   - tf.sparse.sparse_dense_matmul as called from h2gcn/models/_layers.py:47,74,76 — TensorFlow's CPU functor
 *     (TF >= 2.0, third-party, not vendored): zero-initialised output, one pass over the nonzeros in stored
 *     (row-major sorted, _dataset.py:535) order, out[row,:] += val * b[col,:] in fp32, single thread.
 *   - GCNLayer.call + Flatten (_layers.py:78-81, H2GCN.py:271-272): hop h lands in columns [h*d, (h+1)*d).
 *   - TransformSPAdj.nhoodSplit for nhood=2 (_dataset.py:138-158): pattern of (A+I)^2 minus pattern of (A+I).
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -fopenmp -fno-fast-math; no FMA contraction so that the
 * arithmetic is mul-then-add like a non-FMA Eigen build).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* COO, single thread, stored order: the literal restatement of the TF CPU functor. */
void oracle_spmm_coo_f32(int64_t nnz, const int64_t *rows, const int64_t *cols, const float *vals,
                         const float *b, int64_t ldb, int64_t d, float *out, int64_t ldo, int64_t n_rows) {
    for (int64_t r = 0; r < n_rows; ++r) memset(out + r * ldo, 0, (size_t)d * sizeof(float));
    for (int64_t k = 0; k < nnz; ++k) {
        const float v = vals[k];
        const float *src = b + cols[k] * ldb;
        float *dst = out + rows[k] * ldo;
        for (int64_t j = 0; j < d; ++j) dst[j] += v * src[j];
    }
}

/* CSR form of the same sum: per output row the nonzeros are visited in the same ascending-column order, so the
 * result is bit-identical to the COO loop; rows are independent, which lets OpenMP use every host core
 * (threads <= 0 -> all).  This is the "all host threads it can use" CPU baseline. */
void oracle_spmm_csr_f32(int64_t n_rows, const int32_t *rowptr, const int32_t *col, const float *vals,
                         const float *b, int64_t ldb, int64_t d, float *out, int64_t ldo, int threads) {
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
#endif
    for (int64_t r = 0; r < n_rows; ++r) {
        float *dst = out + r * ldo;
        for (int64_t j = 0; j < d; ++j) dst[j] = 0.0f;
        for (int32_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
            const float v = vals[k];
            const float *src = b + (int64_t)col[k] * ldb;
            for (int64_t j = 0; j < d; ++j) dst[j] += v * src[j];
        }
    }
}

/* One fused round: Y[:, off1:off1+d] = A1 X, Y[:, off2:off2+d] = A2 X (GCNLayer + Flatten). */
void oracle_fused_round_f32(int64_t n, const int32_t *rp1, const int32_t *c1, const float *v1,
                            const int32_t *rp2, const int32_t *c2, const float *v2,
                            const float *x, int64_t ldx, int64_t d, float *y, int64_t ldy,
                            int64_t off1, int64_t off2, int threads) {
    oracle_spmm_csr_f32(n, rp1, c1, v1, x, ldx, d, y + off1, ldy, threads);
    oracle_spmm_csr_f32(n, rp2, c2, v2, x, ldx, d, y + off2, ldy, threads);
}

/* Exact-distance-2 pattern.  adj: CSR without diagonal, sorted columns.  Two calls: col2 == NULL counts
 * (fills rowptr2[0..n]), otherwise fills sorted columns.  Marker array per thread (dense, O(n)). */
int64_t oracle_hop2_csr(int32_t n, const int32_t *rowptr, const int32_t *col, int64_t *rowptr2, int32_t *col2,
                        int threads) {
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#else
    threads = 1;
#endif
    if (!col2) rowptr2[0] = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(threads)
#endif
    {
        int32_t *mark = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        int32_t *buf = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
        for (int32_t i = 0; i < n; ++i) mark[i] = -1;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 64)
#endif
        for (int32_t i = 0; i < n; ++i) {
            int32_t cnt = 0;
            mark[i] = i;                                              /* distance 0 */
            for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) mark[col[k]] = i;   /* distance 1 */
            for (int32_t k = rowptr[i]; k < rowptr[i + 1]; ++k) {
                const int32_t j = col[k];
                for (int32_t q = rowptr[j]; q < rowptr[j + 1]; ++q) {
                    const int32_t t = col[q];
                    if (mark[t] != i) { mark[t] = i; buf[cnt++] = t; }
                }
            }
            if (!col2) {
                rowptr2[i + 1] = cnt;
            } else {
                /* ascending order: insertion sort is fine for short rows, qsort-free radix not needed here */
                int32_t *dst = col2 + rowptr2[i];
                for (int32_t a = 0; a < cnt; ++a) dst[a] = buf[a];
                /* simple heap-less sort: shell sort */
                for (int32_t gap = cnt / 2; gap > 0; gap /= 2)
                    for (int32_t a = gap; a < cnt; ++a) {
                        int32_t t = dst[a], p = a;
                        while (p >= gap && dst[p - gap] > t) { dst[p] = dst[p - gap]; p -= gap; }
                        dst[p] = t;
                    }
            }
        }
        free(mark);
        free(buf);
    }
    if (!col2) {
        for (int32_t i = 0; i < n; ++i) rowptr2[i + 1] += rowptr2[i];
        return rowptr2[n];
    }
    return rowptr2[n];
}
