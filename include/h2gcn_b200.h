/* h2gcn_b200 — C-ABI of the B200-native H2GCN aggregation hot path.
 *
 * The reference (GemsLab/H2GCN) has no FFI layer: the path is Python calling TensorFlow / scipy.  Every entry
 * point below replaces one call site of the reference (paths relative to /root/reference/);
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add at each of them.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.  All pointers are DEVICE pointers unless the name ends in
 *    `_host`.  The caller owns every buffer; the library never allocates device memory behind the caller's back
 *    (workspace size queries + caller-provided workspaces).  The one exception is the `h2_graph_*` handle used by
 *    the host-buffer (end-to-end) entry points, which owns its device copies explicitly and is freed explicitly.
 *  - every function returns an int status (H2_OK == 0); nothing throws or aborts.  h2_last_error() returns a
 *    thread-local message for the last non-zero status.
 *  - every device function takes a CUDA stream (`void*` == cudaStream_t), enqueues and returns.  Only functions whose
 *    doc says "SYNCHRONISES" wait for the stream (they return a size the host needs).
 *  - CSR: rowptr is int64 [n+1] (a shard's hop-2 pattern can exceed 2^31 entries), col is int32, val is fp32.
 *    Column indices inside a row are ascending (tf.sparse.reorder order, h2gcn/datasets/_dataset.py:535).
 *  - dense matrices are row-major with an explicit leading dimension in ELEMENTS (so a kernel can read and write
 *    column slices of the zero-copy concat buffer).
 */
#ifndef H2GCN_B200_H
#define H2GCN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H2_ABI_VERSION 1

enum {
    H2_OK = 0,
    H2_ERR_INVALID = 1,     /* bad argument (null pointer, negative size, too many hops, ...)        -> ValueError  */
    H2_ERR_ALIGN = 2,       /* pointer / leading dimension / offset not aligned as documented          -> ValueError  */
    H2_ERR_WORKSPACE = 3,   /* caller-provided workspace too small                                      -> ValueError  */
    H2_ERR_CUDA = 4,        /* a CUDA runtime call failed (message has cudaGetErrorString)              -> RuntimeError */
    H2_ERR_UNSUPPORTED = 5, /* shape outside what the kernels cover                                     -> ValueError  */
    H2_ERR_INDEX = 6        /* h2_validate_csr found an out-of-range / unsorted column                  -> ValueError  */
};

#define H2_MAX_HOPS 8

/* element type of the dense feature matrices of a round (BASELINE config 5: bf16 features, fp32 accumulation) */
#define H2_F32 0
#define H2_BF16 1

typedef void *h2_stream_t; /* cudaStream_t */
typedef struct h2_graph h2_graph_t;

/* One normalised hop adjacency \bar{A}_h as the reference hands it to GCNLayer (a tf.SparseTensor, row-major sorted;
 * h2gcn/datasets/_dataset.py:528-535), in CSR form, plus where its product lands in the output row:
 * out[i, out_col_off : out_col_off + d] = sum_k val[k] * X[col[k], :]   (tf.stack(axis=-2) + Flatten column order,
 * h2gcn/models/_layers.py:78-81, h2gcn/models/H2GCN.py:271-272). */
typedef struct {
    const int64_t *rowptr; /* [n+1] */
    const int32_t *col;    /* [nnz] */
    const float *val;      /* [nnz] explicit fp32 values (reference semantics) or NULL => factored: val = dinv[i]*dinv[j] */
    const float *dinv;     /* [n_cols] column scale, only read when val == NULL */
    const float *dinv_row; /* [n_rows] row scale of the LOCAL rows (= dinv + row_begin), only read when val == NULL */
    int64_t out_col_off;   /* column offset (elements) inside the output row */
    int64_t in_col_off;    /* column offset (elements) inside the input row: 0 in the forward pass; in the backward pass hop h
                              reads ITS slice of the gradient (dX = sum_h A_h^T dY_h, A_h symmetric) */
} h2_hop_t;

/* ---- library / errors -------------------------------------------------------------------------------------- */
int h2_abi_version(void);
const char *h2_last_error(void);
/* number of kernels of THIS library launched by the calling process since load (bench.py's gpu_launches). */
int64_t h2_launch_count(void);

/* ---- a1..a4: adjacency-power precompute ---------------------------------------------------------------------
 * replaces TransformSPAdj.removeEye / nhoodSplit / normalize + sparse2Tensor,
 * h2gcn/datasets/_dataset.py:132-136, 138-158, 109-124, 528-535 as called from getTensors :559-576. */

/* removeEye: drops the diagonal entries of a CSR matrix; two calls: count (writes rowcount_out[n]), then fill
 * (val_in / val_out optional: NULL for a pure pattern). */
int h2_remove_eye_count(int32_t n, const int64_t *rowptr, const int32_t *col, int64_t *rowcount_out, h2_stream_t s);
int h2_remove_eye_fill(int32_t n, const int64_t *rowptr, const int32_t *col, const float *val_in,
                       const int64_t *rowptr_out, int32_t *col_out, float *val_out, h2_stream_t s);

/* exclusive prefix sum of int64 counts[n] into rowptr[n+1] (rowptr[0]=0).  ws from h2_scan_workspace_bytes. */
size_t h2_scan_workspace_bytes(int64_t n);
int h2_exclusive_scan_i64(int64_t n, const int64_t *counts, int64_t *rowptr, void *ws, size_t ws_bytes, h2_stream_t s);

/* nhoodSplit for nhood = 2: pattern of vertices at distance EXACTLY 2, bin((A+I)^2) - bin(A+I).
 * `rowptr/col` = A without self loops, rows sorted, LOCAL rows [row_begin, row_end) of a graph with n vertices are
 * produced (row sharding, SURVEY.md §8e: the full A is replicated, output rows are local).
 * count: rowcount2[row_end-row_begin];  fill: col2 ascending inside each row (tf.sparse.reorder order for free). */
int h2_hop2_count(int32_t n, const int64_t *rowptr, const int32_t *col, int32_t row_begin, int32_t row_end,
                  int64_t *rowcount2, h2_stream_t s);
int h2_hop2_fill(int32_t n, const int64_t *rowptr, const int32_t *col, int32_t row_begin, int32_t row_end,
                 const int64_t *rowptr2, int32_t *col2, h2_stream_t s);

/* normalize(., SYM_NORMALIZED) for a binary pattern: deg = row count (global degree vector supplied by the caller
 * when rows are sharded: `deg_all` [n_cols] int64, or NULL => computed from rowptr, requires n_rows == n_cols);
 * dinv64[i] = deg^-1/2 with inf -> 0 (the degree mask, _dataset.py:115-116); val[k] = fp32((dinv_i * 1.0) * dinv_j)
 * evaluated in fp64 (left-associated like `DInvSqrt @ adj @ DInvSqrt`, :117-118, then X.data.astype(float32), :535).
 * Any of dinv64 / dinv32 / val may be NULL to skip that output.  row_begin = global index of local row 0. */
int h2_sym_normalize(int32_t n_rows, int32_t n_cols, int32_t row_begin, const int64_t *rowptr, const int32_t *col,
                     const int64_t *deg_all, double *dinv64, float *dinv32, float *val, h2_stream_t s);
/* RW_NORMALIZED (_dataset.py:119-123): val = fp32(1/deg_i), inf -> 0. */
int h2_rw_normalize(int32_t n_rows, const int64_t *rowptr, float *val, h2_stream_t s);

/* debug-mode validation: columns in [0, n_cols), strictly ascending inside each row, rowptr monotone.
 * `flag_dev` is a caller-owned device int32 scratch word.  SYNCHRONISES.  Returns H2_ERR_INDEX on a violation
 * (TF raises InvalidArgumentError for the same on CPU). */
int h2_validate_csr(int32_t n_rows, int32_t n_cols, const int64_t *rowptr, const int32_t *col, int32_t *flag_dev,
                    h2_stream_t s);

/* ---- a6+a7+a8: the fused aggregation round ------------------------------------------------------------------
 * replaces GCNLayer.sparse_dense_matmul / GCNLayer.call (h2gcn/models/_layers.py:62-81) + Flatten (H2GCN.py:271-272)
 * + the copies of ConcatLayer.call (_layers.py:90-96) by writing every hop straight into its column slot.
 *
 * The schedule ("plan") depends only on the sparsity structure and is built once per graph, like the reference's
 * once-per-process tensors (H2GCN.py:54).  h2_plan_build SYNCHRONISES (it returns the launch geometry in the plan
 * header, which lives on the host inside the caller's `plan_host` buffer of h2_plan_host_bytes() bytes); the device
 * part of the plan lives in the caller's `plan_dev` buffer of h2_plan_dev_bytes(n_rows, n_hops) bytes. */
size_t h2_plan_host_bytes(void);
size_t h2_plan_dev_bytes(int32_t n_rows, int32_t n_hops);
size_t h2_plan_workspace_bytes(int32_t n_rows, int32_t n_hops);
int h2_plan_build(int32_t n_rows, int32_t n_hops, const h2_hop_t *hops_host, void *plan_host, void *plan_dev,
                  void *ws, size_t ws_bytes, h2_stream_t s);

/* Y[i, hop.out_col_off : +d] = \bar{A}_hop[i, :] . X   for every hop, one launch.
 * X: [n_cols, d] fp32 with leading dimension ldx; Y: [n_rows, *] fp32 with leading dimension ldy.  X and Y may be
 * disjoint column slices of the same buffer.  Requires d % 4 == 0, ldx % 4 == 0, ldy % 4 == 0, out_col_off % 4 == 0
 * and 16-byte aligned X / Y (H2_ERR_ALIGN otherwise).  The hops must be the ones the plan was built from. */
int h2_fused_hops_spmm_f32(const void *plan_host, const void *plan_dev, int32_t n_rows, int32_t n_hops,
                           const h2_hop_t *hops_host, int32_t d, const float *X, int64_t ldx, float *Y, int64_t ldy,
                           h2_stream_t s);

/* Same round with bf16 feature rows in and / or out (x_dtype / y_dtype = H2_F32 | H2_BF16), fp32 accumulation, one rounding
 * to bf16 at the store.  Leading dimensions and column offsets count ELEMENTS of the row type; bf16 rows need d, ld and
 * offsets to be multiples of 8 (16-byte rows). */
int h2_fused_hops_spmm_ex(const void *plan_host, const void *plan_dev, int32_t n_rows, int32_t n_hops,
                          const h2_hop_t *hops_host, int32_t d, const void *X, int64_t ldx, int32_t x_dtype, void *Y,
                          int64_t ldy, int32_t y_dtype, h2_stream_t s);

/* ---- a6 on the tensor cores: dense-ish BINARY hop patterns in "tile bitmap" format ------------------------------
 *   For a hop whose normalised values factor as dinv_row[i] * dinv_col[j] over a 0/1 pattern
 * (every SYM/RW-normalised nhoodSplit ring), Y = diag(dinv_row) . P . (diag(dinv_col) . X) is evaluated with
 * tcgen05.mma: P as exact bf16 0/1 tiles expanded on the fly from 1 bit per entry, X' = diag(dinv_col) X split into
 * `splits` (2 or 3) bf16 pieces, fp32 accumulation in TMEM.  Same call site as h2_fused_hops_spmm_f32
 * (GCNLayer.sparse_dense_matmul, h2gcn/models/_layers.py:62-76); results agree with the fp32 CSR path to ~4e-6
 * relative (splits = 2) — inside the 1e-4 north-star tolerance, not bit-identical.
 *
 * `splits` selects the arithmetic of the tensor-core path:
 *   2, 3               X' as 2 / 3 bf16 pieces (16 / 24 significand bits), kind::f16, fp32 accumulation;
 *   H2_SPLITS_I8X2/3   X' as 2 / 3 balanced base-256 int8 digits of a block-fixed-point number (one step for the whole
 *                      matrix, a power-of-two block exponent 2^t, t in 0..6, per 4 consecutive rows of X' carried by
 *                      the 0/1 operand as 0/2^t), kind::i8 at twice the bf16 rate, EXACT int32 accumulation, one fp32
 *                      rounding in the epilogue.  Measured against the fp32 oracle: ~4e-5 of max-abs (I8X2), ~2e-7 (I8X3).
 *                      The bitmaps of an int8 plan use bit order 1 (h2_bm_fill_order); n_cols <= 2^17 (int32 accumulators).
 *
 * Build (once per graph, two phases like hop2): h2_bm_count SYNCHRONISES and returns the number of non-empty
 * 256x64 units; the caller allocates h2_bm_plan_dev_bytes(); h2_bm_fill SYNCHRONISES (it builds the stream-K
 * schedule on the host).  `index_ws` (h2_bm_index_bytes) must stay untouched between the two calls. */
#define H2_SPLITS_I8X2 18
#define H2_SPLITS_I8X3 19
size_t h2_bm_host_bytes(void);
size_t h2_bm_index_bytes(int32_t n_rows, int32_t n_cols);
int h2_bm_count(int32_t n_rows, int32_t n_cols, const int64_t *rowptr, const int32_t *col, void *index_ws,
                size_t index_ws_bytes, int64_t *n_units_host, h2_stream_t s);
size_t h2_bm_plan_dev_bytes(int32_t n_rows, int32_t n_cols, int64_t n_units);
int h2_bm_fill(int32_t n_rows, int32_t n_cols, const int64_t *rowptr, const int32_t *col, void *index_ws,
               int64_t n_units, void *bm_host, void *bm_dev, size_t bm_dev_bytes, h2_stream_t s);
/* same with an explicit bitmap bit order: 0 = natural (bf16 `splits`), 1 = int8 `splits` */
int h2_bm_fill_order(int32_t n_rows, int32_t n_cols, const int64_t *rowptr, const int32_t *col, void *index_ws,
                     int64_t n_units, void *bm_host, void *bm_dev, size_t bm_dev_bytes, int32_t bit_order, h2_stream_t s);
/* widest d one pack + spmm pair covers (8 column groups: 512 for the int8 digits and 2 bf16 pieces, 256 for 3 bf16
 * pieces); the h2_graph_round* entry points compute wider rounds in column slices of the same buffers. */
int32_t h2_bm_max_width(int32_t splits);
/* per round: pack X' once per distinct dinv_col (NULL = no column scaling), then one h2_bm_spmm_f32 per hop.
 * Non-finite inputs: the int8 operand has ONE step for the whole matrix, so NaN is packed as 0 and +-Inf saturates to the
 * largest finite magnitude instead of poisoning every output (the fp32 CSR path propagates them like the reference). */
size_t h2_bm_xpack_bytes(int32_t n_cols, int32_t d, int32_t splits);
size_t h2_bm_partial_bytes(const void *bm_host, int32_t d, int32_t splits);
int h2_bm_pack_x_f32(int32_t n_cols, int32_t d, int32_t splits, const float *X, int64_t ldx, const float *dinv_col,
                     void *xpack, size_t xpack_bytes, h2_stream_t s);
/* same, for a buffer whose first 256 bytes were zeroed once and which only ever went through these pack calls (the int8
 * pack kernel keeps its grid-barrier counters there and re-arms them itself): no memset per call */
int h2_bm_pack_x_f32_armed(int32_t n_cols, int32_t d, int32_t splits, const float *X, int64_t ldx, const float *dinv_col,
                           void *xpack, size_t xpack_bytes, h2_stream_t s);
int h2_bm_spmm_f32(const void *bm_host, const void *bm_dev, int32_t d, int32_t splits, const void *xpack,
                   const float *dinv_row, float *Y, int64_t ldy, int64_t out_col_off, void *partial_ws,
                   size_t partial_bytes, h2_stream_t s);

/* ---- a5 / a10: dense ends writing into the concat buffer ----------------------------------------------------
 * replaces SparseDense.call (+ReLU) (_layers.py:45-52, H2GCN.py:269-270): Y[:, off:off+p] = act(Xs . W + b),
 * Xs CSR [n, F] fp32 (row-major sorted COO in the reference), W [F, p] row-major. */
int h2_sparse_dense_f32(int32_t n_rows, const int64_t *rowptr, const int32_t *col, const float *val, const float *W,
                        int32_t p, const float *bias, int32_t relu, float *Y, int64_t ldy, int64_t out_col_off,
                        h2_stream_t s);
/* replaces keras Dense (H2GCN.py:244-249): Y[n, c] = act(X[n, k] . W[k, c] + b), fp32 SIMT (parity mode). */
int h2_dense_f32(int32_t n_rows, int32_t k, int32_t c, const float *X, int64_t ldx, const float *W, const float *bias,
                 int32_t relu, float *Y, int64_t ldy, int64_t out_col_off, h2_stream_t s);
/* The same contraction on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split: fp32-equivalent, ~3e-7 of max-abs):
 * Y[m, off:off+n] = act(op(A) . op(W) + b) with A [m, k] (trans_a == 0) or [k, m] (trans_a != 0), W [k, n] (trans_w == 0)
 * or [n, k] (trans_w != 0), all fp32 row-major with leading dimensions in elements.  Default path of the classifier
 * (H2GCN.py:244-257), of SparseDense on dense features (_layers.py:45-52) and of the classifier-side contractions of the
 * training step (dW = final^T . dlogits: trans_a; dfinal = dlogits . W^T: trans_w).  No alignment requirements. */
int h2_dense_tc_f32(int32_t m, int32_t k, int32_t n, const float *A, int64_t lda, int32_t trans_a, const float *W, int64_t ldw,
                    int32_t trans_w, const float *bias, int32_t relu, float *Y, int64_t ldy, int64_t out_col_off, h2_stream_t s);
/* ReLU / copy of a column slice (layers the planner could not fuse away). */
int h2_relu_slice_f32(int32_t n_rows, int32_t d, const float *X, int64_t ldx, float *Y, int64_t ldy, int32_t relu,
                      h2_stream_t s);

/* ---- end-to-end entry points with HOST buffers (bench.py `e2e`, the call a CPU-side caller makes) ------------
 * The graph (all hops, in the format chosen per hop, + schedules) is uploaded once (h2_graph_create, like getTensors
 * runs once per process); every h2_graph_round_host call copies X host->device, runs the fused round and copies
 * Y [n_rows, n_hops*d] device->host on `s`, then SYNCHRONISES.  x_host / y_host should be pinned for full PCIe speed.
 * dinv_host[h]: fp32 [n_cols] scale vector when hop h is a normalised BINARY pattern (val = dinv[i]*dinv[j]); it
 * enables the tensor-core format for that hop (NULL entry / NULL array: CSR only).  row_begin: global index of local
 * row 0.  mode: 0 = auto (bitmap for density >= 1 %), 1 = CSR everywhere, 2 = bitmap wherever dinv is given. */
int h2_graph_create(int32_t n_rows, int32_t n_cols, int32_t n_hops, const int64_t *const *rowptr_host,
                    const int32_t *const *col_host, const float *const *val_host, const float *const *dinv_host,
                    int32_t row_begin, int32_t d_max, int32_t mode, int32_t splits, h2_graph_t **out);
/* Same handle over CSR arrays that ALREADY live on the device (no copies; the caller keeps them alive): this is what
 * the Python GCNLayer / HopPlan uses, so that one fused round is ONE C call.  hops[h].val may be NULL when hops[h].dinv
 * is given (factored CSR); hops with dinv may take the tensor-core format.  nnz_host[h] = stored entries of hop h. */
int h2_graph_create_device(int32_t n_rows, int32_t n_cols, int32_t n_hops, const h2_hop_t *hops, const int64_t *nnz_host,
                           int32_t row_begin, int32_t mode, int32_t splits, h2_graph_t **out);
/* Scratch of the tensor-core hops (packed operand, stream-K partial tiles) for rounds of width <= d_max.  The round entry
 * points below NEVER allocate or synchronise: before the first round either bind a caller-owned workspace of
 * h2_graph_workspace_bytes(g, d_max) bytes (16-byte aligned, alive while rounds are in flight), or let the library
 * allocate with h2_graph_reserve (SYNCHRONISES; h2_graph_create does this for its d_max).  A round wider than the reserved
 * d_max returns H2_ERR_WORKSPACE.  The scratch is per handle: rounds on the same handle are serialised by an event the
 * handle records at the end of every round (a round issued on another stream waits for the previous one), so a handle
 * may be used from several streams; independent rounds that should OVERLAP need one handle each. */
size_t h2_graph_workspace_bytes(const h2_graph_t *g, int32_t d_max);
int h2_graph_bind_workspace(h2_graph_t *g, int32_t d_max, void *ws, size_t ws_bytes);
int h2_graph_reserve(h2_graph_t *g, int32_t d_max);
/* fmt_out[h] = 0 (CSR) / 1 (tile bitmap, tensor cores) */
int h2_graph_formats(const h2_graph_t *g, int32_t *fmt_out);
int h2_graph_round_host(h2_graph_t *g, int32_t d, const float *x_host, float *y_host, h2_stream_t s);
/* same round on DEVICE buffers (X [n_cols, d] ld=ldx; Y: hop h at column offsets[h]); enqueues, does not synchronise. */
int h2_graph_round(h2_graph_t *g, int32_t d, const float *X, int64_t ldx, float *Y, int64_t ldy,
                   const int64_t *offsets_host, h2_stream_t s);
/* the round with bf16 feature rows (BASELINE config 5): X and / or Y hold bf16 (H2_BF16), accumulation is fp32 (CSR hops)
 * or exact int32 over the int8 digits of X' (tensor-core hops: the bf16 input is exactly representable in the digits'
 * 16 / 24 bits up to the row scaling, so the only rounding of the round is the final fp32 -> bf16 store).  int8 `splits`
 * only; d, ldx, ldy and the offsets are multiples of 8 elements. */
int h2_graph_round_ex(h2_graph_t *g, int32_t d, const void *X, int64_t ldx, int32_t x_dtype, void *Y, int64_t ldy,
                      int32_t y_dtype, const int64_t *offsets_host, h2_stream_t s);
/* backward-style round: hop h reads X[:, x_offsets[h] : +d] (its own slice) and writes Y[:, y_offsets[h] : +d]. */
int h2_graph_round_multi(h2_graph_t *g, int32_t d, const float *X, int64_t ldx, const int64_t *x_offsets_host, float *Y,
                         int64_t ldy, const int64_t *y_offsets_host, h2_stream_t s);
/* Multi-GPU round with the hop-boundary all-gather FUSED into the first consumer: the round input is given
 * as n_parts (<= 8) row shards — part q holds rows [bounds[q], bounds[q+1]) at part_ptrs[q] (row-major, ld_part) — which
 * may be PEER pointers (NVLink P2P / symmetric memory).  The pack kernel of the first tensor-core hop (or a plain gather
 * kernel when every hop is CSR) loads the shards directly over NVLink and also writes the gathered fp32 copy
 * x_full [n_cols, d] (caller scratch) that CSR hops read.  The caller provides the cross-rank barriers around the call. */
int h2_graph_round_parts(h2_graph_t *g, int32_t d, int32_t n_parts, const float *const *part_ptrs_host,
                         const int64_t *bounds_host, int64_t ld_part, float *x_full, int64_t ld_full, float *Y, int64_t ldy,
                         const int64_t *y_offsets_host, h2_stream_t s);
/* the same with bf16 feature rows (BASELINE config 5 on several GPUs): the shards and the gathered copy x_full hold
 * x_dtype rows (ld_part / ld_full count elements of that type), Y holds y_dtype rows; d, the leading dimensions and the
 * offsets are multiples of 8 when a bf16 type is involved. */
int h2_graph_round_parts_ex(h2_graph_t *g, int32_t d, int32_t n_parts, const void *const *part_ptrs_host,
                            const int64_t *bounds_host, int64_t ld_part, int32_t x_dtype, void *x_full, int64_t ld_full,
                            void *Y, int64_t ldy, int32_t y_dtype, const int64_t *y_offsets_host, h2_stream_t s);
/* G[:, 0:d] (+)= sum_s T[:, s*d : (s+1)*d], optionally masked by (mask_src > 0) (ReLU gradient).  accumulate != 0: add to G. */
int h2_sum_slices_f32(int32_t n_rows, int32_t d, int32_t n_slices, const float *T, int64_t ldt, float *G, int64_t ldg,
                      int32_t accumulate, const float *mask_src, int64_t ld_mask, h2_stream_t s);
int h2_graph_destroy(h2_graph_t *g);

#ifdef __cplusplus
}
#endif
#endif /* H2GCN_B200_H */
