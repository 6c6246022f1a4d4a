"""CPU oracle for the H2GCN aggregation hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import this
module.  The product (`h2gcn_b200/`) never does; it fails loudly when its CUDA library is missing.

Parity status: PINNED against the reference's own Python executed in this container.  `tests/golden/make_golden.py`
imports `/root/reference/h2gcn/{datasets/_dataset.py,models/*.py}` unmodified (TensorFlow replaced by a numpy shim
for the third-party ops) and commits its outputs; `tests/test_oracle_golden.py` checks every function below against
those vectors (indices and fp32 adjacency values bit-exact, activations to 1e-6).  What remains an assumption is the
inside of TensorFlow's CPU `SparseTensorDenseMatMul` (TF >= 2.0, "tested on 2.2", reference README.md:34; not
vendored): restated here as zero-init + in-order `out[r,:] += v * b[c,:]` in fp32.

Each function cites the reference lines it follows (paths relative to /root/reference/).
"""
import re

import numpy as np
import scipy.sparse as sp


# ----------------------------------------------------------------------------------------------------------
# adjacency precompute  (h2gcn/datasets/_dataset.py)
# ----------------------------------------------------------------------------------------------------------
def remove_eye(adj):
    """_dataset.py:132-136 (TransformSPAdj.removeEye): zero the diagonal, drop the zeros."""
    a = sp.csr_matrix(adj).tocoo()
    keep = a.row != a.col
    out = sp.csr_matrix((a.data[keep], (a.row[keep], a.col[keep])), shape=a.shape)
    out.sum_duplicates()
    out.sort_indices()
    return out


def nhood_split(adj, nhood):
    """_dataset.py:138-158 (TransformSPAdj.nhoodSplit).

    Returns [I, ring_1, ..., ring_k]: ring_h = pattern((A+I)^h) - pattern((A+I)^(h-1)), float64 ones.
    Stops early (short list!) when a power adds no entry (:151-153)."""
    n = adj.shape[0]
    assert adj.ndim == 2 and adj.shape[0] == adj.shape[1]
    step = (sp.csr_matrix(adj) + sp.identity(n, format="csr")).astype(np.float64)
    reach = sp.identity(n, dtype=np.float64, format="csr")
    rings = [reach]
    total = 0
    for _ in range(int(nhood)):
        grown = reach @ step
        grown.data[:] = (grown.data > 0)
        grown.eliminate_zeros()
        new_total = int(grown.nnz)  # == mt.sum() since every stored value is exactly 1.0
        if new_total == total:
            break
        total = new_total
        ring = (grown - reach).tocsr()
        ring.eliminate_zeros()
        ring.sort_indices()
        rings.append(ring)
        reach = grown
    return rings


def sym_normalize(mat):
    """_dataset.py:114-118 (NType.SYM_NORMALIZED): D^-1/2 M D^-1/2, zero-degree rows masked to 0, all fp64.
    Value order of operations: (dinv_i * m_ij) * dinv_j (left-assoc `DInvSqrt @ adj @ DInvSqrt`)."""
    m = sp.csr_matrix(mat).astype(np.float64)
    deg = np.asarray(m.sum(axis=1)).ravel()
    with np.errstate(divide="ignore"):
        dinv = np.power(deg, -0.5)
    dinv[np.isinf(dinv)] = 0.0
    out = m.copy()
    rows = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
    out.data = (dinv[rows] * m.data) * dinv[m.indices]
    return out, deg, dinv


def rw_normalize(mat):
    """_dataset.py:119-123 (NType.RW_NORMALIZED): D^-1 M."""
    m = sp.csr_matrix(mat).astype(np.float64)
    deg = np.asarray(m.sum(axis=1)).ravel()
    with np.errstate(divide="ignore"):
        dinv = np.power(deg, -1.0)
    dinv[np.isinf(dinv)] = 0.0
    out = m.copy()
    rows = np.repeat(np.arange(m.shape[0]), np.diff(m.indptr))
    out.data = dinv[rows] * m.data
    return out, deg, dinv


def row_normalize_features(features):
    """_dataset.py:502-509: X <- diag(1/rowsum) X with inf -> 0, in the dtype of X (fp32 for Planetoid).
    The row sums are taken with the CONTAINER'S OWN `.sum(1)` exactly like the reference: for the LIL matrices the
    Planetoid loader builds (:241) scipy evaluates it as a sequential fp32 mat-vec, while CSR uses a pairwise
    reduction — the two differ in the last ulp, so pass the same container type the reference holds."""
    rowsum = np.asarray(features.sum(1)).ravel()
    with np.errstate(divide="ignore"):
        inv = np.power(rowsum, -1)
    inv[np.isinf(inv)] = 0.0
    return (sp.diags(inv) @ features).tocsr()


def to_coo_sorted(spmat, dtype=np.float32):
    """_dataset.py:528-535 (sparse2Tensor): COO, values cast to fp32, tf.sparse.reorder => row-major, col ascending.
    Explicit zeros that scipy keeps stored are kept (TF keeps them too)."""
    c = sp.coo_matrix(spmat)
    order = np.lexsort((c.col, c.row))
    return c.row[order].astype(np.int64), c.col[order].astype(np.int64), c.data[order].astype(dtype)


def parse_hops(spec):
    """_dataset.py:560-562: ["1","2"] -> [[1],[2]];  "1,2" merges hops."""
    return [[int(x) for x in str(e).split(",")] for e in spec]


def adj_norm_hops(adj_no_eye, spec=("1", "2"), norm="sym"):
    """_dataset.py:559-576 (getTensors, getAdjNormHops branch): list of sorted-COO (rows, cols, vals_fp32)."""
    hops = parse_hops(spec)
    rings = nhood_split(adj_no_eye, max(max(h) for h in hops))
    out = []
    for sel in hops:
        merged = rings[sel[0]]          # IndexError if nhood_split stopped early, like the reference (:571)
        for i in sel[1:]:
            merged = merged + rings[i]
        normed = sym_normalize(merged)[0] if norm == "sym" else rw_normalize(merged)[0]
        out.append(to_coo_sorted(normed))
    return out


def coo_to_csr(rows, cols, n):
    """Helper (no reference counterpart): rowptr of a row-major sorted COO."""
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, np.asarray(rows) + 1, 1)
    return np.cumsum(rowptr), np.asarray(cols)


# ----------------------------------------------------------------------------------------------------------
# aggregation  (h2gcn/models/_layers.py)
# ----------------------------------------------------------------------------------------------------------
def spmm_coo(rows, cols, vals, b, n_rows, dtype=np.float32):
    """tf.sparse.sparse_dense_matmul as called at _layers.py:47,74,76.  Sequential, in stored order, in `dtype`.
    (scipy's csr @ dense uses the same per-row ascending order; kept explicit here so the order is visible.)"""
    b = np.asarray(b, dtype=dtype)
    vals = np.asarray(vals, dtype=dtype)
    out = np.zeros((n_rows, b.shape[1]), dtype=dtype)
    step = max(1, (1 << 22) // max(1, b.shape[1]))
    for s in range(0, len(vals), step):
        e = min(len(vals), s + step)
        np.add.at(out, rows[s:e], vals[s:e, None] * b[cols[s:e]])
    return out


def gcn_layer(adjhops, x, hops=None, dtype=np.float32):
    """_layers.py:78-81 (GCNLayer.call): stack of per-hop products on axis -2 => [N, H, d]."""
    n = x.shape[0]
    ys = [spmm_coo(r, c, v, x, n, dtype) for i, (r, c, v) in enumerate(adjhops) if hops is None or i in hops]
    return np.stack(ys, axis=-2)


def fused_round(adjhops, x, dtype=np.float32):
    """GCNLayer + Flatten (H2GCN.py:271-272): [N, H*d], column = hop*d + j."""
    y = gcn_layer(adjhops, x, None, dtype)
    return y.reshape(y.shape[0], -1)


def sparse_dense(feat_coo, n_rows, kernel, bias=None, dtype=np.float32):
    """_layers.py:45-52 (SparseDense.call) without activation."""
    out = spmm_coo(feat_coo[0], feat_coo[1], feat_coo[2], kernel, n_rows, dtype)
    if bias is not None:
        out = out + bias
    return out


def dense(x, kernel, bias=None):
    """keras Dense (H2GCN.py:244-249): x @ kernel (+ bias); fp32 sequential-k accumulation."""
    out = np.zeros((x.shape[0], kernel.shape[1]), dtype=np.float32)
    for j in range(kernel.shape[0]):
        out += x[:, j:j + 1] * kernel[j:j + 1, :]
    if bias is not None:
        out = out + bias
    return out


# ----------------------------------------------------------------------------------------------------------
# model forward  (h2gcn/models/H2GCN.py:210-346), inference only
# ----------------------------------------------------------------------------------------------------------
def forward(layer_setups, weights, feat_coo, n_rows, adjhops, sparse_input=True, return_activations=False):
    """H2GCN.__init__ (:226-292) + H2GCN.call (:307-341) folded into one interpreter.

    layer_setups: output of parse_network_setup; weights: list of arrays in layer order (kernel[, bias]).
    Dropout is the identity (training=False)."""
    x = feat_coo
    tagged = {}
    acts = []
    wi = 0
    sparse = sparse_input
    for ltype, conf in layer_setups:
        conf = dict(conf)
        tag = conf.pop("tag", None)
        if ltype == "F":
            k = weights[wi]; wi += 1
            b = None
            if conf["use_bias"]:
                b = weights[wi]; wi += 1
            if sparse:
                x = sparse_dense(x, n_rows, k, b)
                sparse = False
            else:
                x = dense(x, k, b)
        elif ltype == "D":
            pass
        elif ltype == "R":
            x = np.maximum(x, 0)
        elif ltype == "G":
            x = gcn_layer(adjhops, x, conf.get("hops"))
        elif ltype == "V":
            x = x.reshape(x.shape[0], -1)
        elif ltype == "C":
            sel = [v for name, v in tagged.items() if name in conf["tags"]]   # dict order = tagging order (_layers.py:91)
            x = np.concatenate(([x] if conf.get("addInputs", True) else []) + sel, axis=-1)
        elif ltype == "I":
            r, c, v = x
            d = np.zeros((n_rows, int(c.max()) + 1 if len(c) else 0), dtype=np.float32)
            d[r, c] = v
            x = d
            sparse = False
        elif ltype == "S":
            src = tagged[conf["loadTag"]] if conf.get("loadTag") else x
            x = src[:, conf["sliceObj"]]
        else:
            raise ValueError(f"Unsupported layer type {ltype} specified in this model.")
        acts.append(x)
        if tag:
            tagged[tag] = x
    return (x, acts) if return_activations else x


# ----------------------------------------------------------------------------------------------------------
# training step of the H2GCN-K family (test infrastructure for SURVEY.md §8f rank 1)
# ----------------------------------------------------------------------------------------------------------
def loss_and_grads(n_rounds, relu, W0, W_out, feat_csr, hops_csr, labels, mask, l2=0.0, drop_mask=None):
    """fp64 restatement of one training forward/backward for 'M<p>[-R]-T1-(G-V-T<k>)*K-C..-D-MO' models.

    Forward: r0 = act(X W0); r_k = [A1 r_{k-1} | A2 r_{k-1}]; final = [r_K | r0 | r1 | ... | r_{K-1}] (the concat order
    of _layers.py:90-96 as laid out by H2GCN.py:273-278); logits = (final * drop_mask) W_out.
    Loss (H2GCN.py:363-367, _metrics.py:8-16): sum_i m_i CE(logits_i, y_i) with m = mask / sum(mask), plus
    l2 * (||W0||^2 + ||W_out||^2) (keras.regularizers.l2).  TensorFlow differentiates this with autograd; the
    gradients below are the analytic ones.  hops_csr: scipy CSR matrices; returns (loss, dW0, dW_out)."""
    X = sp.csr_matrix(feat_csr).astype(np.float64)
    A = [sp.csr_matrix(a).astype(np.float64) for a in hops_csr]
    W0 = np.asarray(W0, dtype=np.float64)
    W_out = np.asarray(W_out, dtype=np.float64)
    z0 = X @ W0
    r = [np.maximum(z0, 0) if relu else z0]
    for _ in range(n_rounds):
        r.append(np.concatenate([a @ r[-1] for a in A], axis=1))
    order = [n_rounds] + list(range(n_rounds)) if n_rounds else [0]
    final = np.concatenate([r[k] for k in order], axis=1)
    dm = np.ones_like(final) if drop_mask is None else np.asarray(drop_mask, dtype=np.float64)
    logits = (final * dm) @ W_out
    y = np.asarray(labels, dtype=np.float64)
    m = np.asarray(mask, dtype=np.float64)
    m = m / m.sum()
    zmax = logits.max(1, keepdims=True)
    logp = logits - zmax - np.log(np.exp(logits - zmax).sum(1, keepdims=True))
    loss = float((-(y * logp).sum(1) * m).sum() + l2 * ((W0 ** 2).sum() + (W_out ** 2).sum()))
    dlogits = (np.exp(logp) * y.sum(1, keepdims=True) - y) * m[:, None]
    dW_out = (final * dm).T @ dlogits + 2 * l2 * W_out
    dfinal = (dlogits @ W_out.T) * dm
    widths = [r[k].shape[1] for k in order]
    offs = np.concatenate([[0], np.cumsum(widths)])
    g = {k: dfinal[:, offs[i]:offs[i + 1]].copy() for i, k in enumerate(order)}
    for k in range(n_rounds, 0, -1):
        d = r[k - 1].shape[1]
        g[k - 1] = g[k - 1] + sum(a.T @ g[k][:, h * d:(h + 1) * d] for h, a in enumerate(A))
    g0 = g[0] * (z0 > 0) if relu else g[0]
    dW0 = X.T @ g0 + 2 * l2 * W0
    return loss, np.asarray(dW0), np.asarray(dW_out)
