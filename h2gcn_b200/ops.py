"""Thin Python wrappers over the C-ABI (include/h2gcn_b200.h).  torch tensors are device buffers only.

Everything here runs on the GPU through libh2gcn_b200.so; there is no CPU path (see _cabi.lib()).
"""
import ctypes

import torch

from . import _cabi
from ._cabi import HopDesc, check, lib, ptr, require_cuda, stream_ptr


def _i64(t):
    return t if t.dtype == torch.int64 else t.to(torch.int64)


# ------------------------------------------------------------------------------------------------------------------
# device CSR container (what the reference carries around as a tf.SparseTensor, _dataset.py:528-535)
# ------------------------------------------------------------------------------------------------------------------
class SparseTensor:
    """Row-major-sorted sparse matrix on the device: CSR (rowptr int64, col int32, values fp32).

    Mirrors the attributes the reference reads from tf.SparseTensor (`indices`, `values`, `dense_shape`, `shape`);
    `indices` ([nnz, 2] int64, row-major order = tf.sparse.reorder order) is materialised lazily.
    `row_begin` is the global index of local row 0 when the rows are a shard (multi-GPU), else 0.
    """

    def __init__(self, rowptr, col, values, dense_shape, row_begin=0, dinv=None):
        require_cuda(rowptr, col, values)
        self.rowptr = _i64(rowptr).contiguous()
        self.col = col.to(torch.int32).contiguous()
        self.values = None if values is None else values.to(torch.float32).contiguous()
        self.dense_shape = tuple(int(x) for x in dense_shape)
        self.row_begin = int(row_begin)
        self.dinv = dinv  # fp32 [n_cols] when the values factor as dinv[i]*dinv[j] (sym-normalised binary pattern)
        self._indices = None

    @property
    def shape(self):
        return self.dense_shape

    @property
    def n_rows(self):
        return self.rowptr.numel() - 1

    @property
    def nnz(self):
        return int(self.col.numel())

    @property
    def device(self):
        return self.rowptr.device

    @property
    def indices(self):
        if self._indices is None:
            counts = self.rowptr[1:] - self.rowptr[:-1]
            rows = torch.repeat_interleave(torch.arange(self.n_rows, device=self.device) + self.row_begin, counts)
            self._indices = torch.stack([rows, self.col.to(torch.int64)], dim=1)
        return self._indices

    @classmethod
    def from_scipy(cls, m, device="cuda", dtype=torch.float32):
        """Host scipy matrix -> canonical device CSR (sorted, duplicates summed) — sparse2Tensor, _dataset.py:528-535."""
        import numpy as np
        import scipy.sparse as sp
        m = sp.csr_matrix(m)
        m.sum_duplicates()
        m.sort_indices()
        dev = torch.device(device)
        return cls(torch.from_numpy(m.indptr.astype(np.int64)).to(dev), torch.from_numpy(m.indices.astype(np.int32)).to(dev),
                   torch.from_numpy(m.data.astype(np.float32)).to(dev), m.shape)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.values.cpu().numpy(), self.col.cpu().numpy(), self.rowptr.cpu().numpy()),
                             shape=(self.n_rows, self.dense_shape[1]))


# ------------------------------------------------------------------------------------------------------------------
# precompute
# ------------------------------------------------------------------------------------------------------------------
def exclusive_scan(counts, stream=None):
    """int64 counts[n] -> rowptr[n+1]."""
    require_cuda(counts)
    n = counts.numel()
    out = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
    ws_bytes = lib().h2_scan_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=counts.device)
    check(lib().h2_exclusive_scan_i64(n, ptr(counts), ptr(out), ptr(ws), ws_bytes, stream_ptr(stream)))
    return out


def remove_eye(rowptr, col, val=None, stream=None):
    """TransformSPAdj.removeEye (_dataset.py:132-136) on a sorted CSR: drop diagonal entries.
    Returns (rowptr, col) or (rowptr, col, val) when values are given."""
    require_cuda(rowptr, col, val)
    n = rowptr.numel() - 1
    rowptr = _i64(rowptr).contiguous()
    col = col.to(torch.int32).contiguous()
    counts = torch.empty(n, dtype=torch.int64, device=rowptr.device)
    s = stream_ptr(stream)
    check(lib().h2_remove_eye_count(n, ptr(rowptr), ptr(col), ptr(counts), s))
    rp = exclusive_scan(counts, stream)
    nnz = int(rp[-1].item()) if n else 0
    out = torch.empty(nnz, dtype=torch.int32, device=rowptr.device)
    vout = None if val is None else torch.empty(nnz, dtype=torch.float32, device=rowptr.device)
    if val is not None:
        val = val.to(torch.float32).contiguous()
    check(lib().h2_remove_eye_fill(n, ptr(rowptr), ptr(col), ptr(val), ptr(rp), ptr(out), ptr(vout), s))
    return (rp, out) if val is None else (rp, out, vout)


def hop2_pattern(rowptr, col, row_begin=0, row_end=None, stream=None):
    """Exact-distance-2 pattern (nhoodSplit(adj, 2)[2], _dataset.py:138-158) of a self-loop-free sorted CSR.
    Returns (rowptr2 int64 [rows+1], col2 int32) for rows [row_begin, row_end); one host sync (count -> alloc -> fill)."""
    require_cuda(rowptr, col)
    n = rowptr.numel() - 1
    row_end = n if row_end is None else row_end
    rowptr = _i64(rowptr).contiguous()
    col = col.to(torch.int32).contiguous()
    rows = row_end - row_begin
    counts = torch.empty(rows, dtype=torch.int64, device=rowptr.device)
    s = stream_ptr(stream)
    check(lib().h2_hop2_count(n, ptr(rowptr), ptr(col), row_begin, row_end, ptr(counts), s))
    rp2 = exclusive_scan(counts, stream)
    nnz2 = int(rp2[-1].item()) if rows else 0
    col2 = torch.empty(nnz2, dtype=torch.int32, device=rowptr.device)
    check(lib().h2_hop2_fill(n, ptr(rowptr), ptr(col), row_begin, row_end, ptr(rp2), ptr(col2), s))
    return rp2, col2


def sym_normalize(rowptr, col, n_cols=None, row_begin=0, deg_all=None, stream=None, want_val=True):
    """normalize(., SYM_NORMALIZED) (_dataset.py:114-118) for a binary pattern.
    Returns (val fp32 [nnz] or None when want_val is False, dinv64 [n_cols], dinv32 [n_cols])."""
    require_cuda(rowptr, col)
    n_rows = rowptr.numel() - 1
    n_cols = n_rows if n_cols is None else n_cols
    dev = rowptr.device
    rowptr = _i64(rowptr).contiguous()
    col = col.to(torch.int32).contiguous()
    val = torch.empty(col.numel(), dtype=torch.float32, device=dev) if want_val else None
    d64 = torch.empty(n_cols, dtype=torch.float64, device=dev)
    d32 = torch.empty(n_cols, dtype=torch.float32, device=dev)
    check(lib().h2_sym_normalize(n_rows, n_cols, row_begin, ptr(rowptr), ptr(col), ptr(deg_all), ptr(d64), ptr(d32),
                                 ptr(val), stream_ptr(stream)))
    return val, d64, d32


def rw_normalize(rowptr, stream=None):
    """normalize(., RW_NORMALIZED) (_dataset.py:119-123) for a binary pattern: val = 1/deg_i."""
    require_cuda(rowptr)
    rowptr = _i64(rowptr).contiguous()
    n = rowptr.numel() - 1
    nnz = int(rowptr[-1].item()) if n else 0
    val = torch.empty(nnz, dtype=torch.float32, device=rowptr.device)
    check(lib().h2_rw_normalize(n, ptr(rowptr), ptr(val), stream_ptr(stream)))
    return val


def validate_csr(sp, stream=None):
    flag = torch.zeros(1, dtype=torch.int32, device=sp.device)
    check(lib().h2_validate_csr(sp.n_rows, sp.dense_shape[1], ptr(sp.rowptr), ptr(sp.col), ptr(flag), stream_ptr(stream)))


# ------------------------------------------------------------------------------------------------------------------
# fused aggregation round
# ------------------------------------------------------------------------------------------------------------------
class HopPlan:
    """One fused aggregation round over a fixed list of hop adjacencies: a thin owner of an `h2_graph_t` handle created
    over the hops' DEVICE arrays (no copies).  Built once per graph; `run` is a single C call.

    Every hop is stored in ONE of two formats, chosen by the library from its density when the plan is built:
      * CSR  -> `fused_hops_gather_kernel`, all CSR hops of the round in one launch (explicit fp32 values, or
                `factored=True`: index-only CSR + dinv);
      * tile bitmap -> tcgen05 kernel (csrc/bitmap_mma.cu), for normalised BINARY patterns (SparseTensor.dinv set) of
                density >= 1 %; CSR hops and tensor-core hops of one round overlap on two streams.
    mode: "auto" (by density), "csr" (reference-exact fp32 arithmetic everywhere), "tensor" (bitmap wherever possible).
    """
    MODES = {"auto": 0, "csr": 1, "tensor": 2}

    def __init__(self, hops, factored=False, mode="auto", splits=None, stream=None):
        if not 1 <= len(hops) <= _cabi.MAX_HOPS:
            raise ValueError(f"between 1 and {_cabi.MAX_HOPS} hops per fused round, got {len(hops)}")
        if mode not in self.MODES:
            raise ValueError(f"unknown mode {mode}")
        self.hops = list(hops)
        self.n_rows = hops[0].n_rows
        self.n_cols = hops[0].dense_shape[1]
        for h in hops:
            if h.n_rows != self.n_rows or h.dense_shape[1] != self.n_cols:
                raise ValueError("all hop adjacencies of a round must have the same shape")
            if factored and h.dinv is None:
                raise ValueError("factored mode needs SparseTensor.dinv on every hop")
            if h.row_begin != hops[0].row_begin:
                raise ValueError("all hops of a round must be the same row shard")
        splits = _cabi.splits_code(splits)
        self.factored, self.splits = factored, splits
        self.nnz = sum(h.nnz for h in hops)
        H = len(hops)
        desc = (HopDesc * H)()
        for k, h in enumerate(hops):
            desc[k].rowptr, desc[k].col = ptr(h.rowptr), ptr(h.col)
            desc[k].val = None if factored else ptr(h.values)
            desc[k].dinv = None if (mode == "csr" and not factored) else ptr(h.dinv)
            desc[k].dinv_row = None
            desc[k].out_col_off = 0
        nnz = (ctypes.c_int64 * H)(*[h.nnz for h in hops])
        self._h = ctypes.c_void_p()
        with torch.cuda.device(hops[0].device):
            check(lib().h2_graph_create_device(self.n_rows, self.n_cols, H, desc, nnz, hops[0].row_begin, self.MODES[mode],
                                               splits, ctypes.byref(self._h)))
        self._ws, self._ws_d = [], 0      # caller-owned scratch of the tensor-core hops (torch allocations, bound below)
        fmt = (ctypes.c_int32 * H)()
        check(lib().h2_graph_formats(self._h, fmt))
        self.tensor_idx = [k for k in range(H) if fmt[k] == 1]
        self.csr_idx = [k for k in range(H) if fmt[k] == 0]
        self.kernel_name = " + ".join(
            (["%s (tcgen05 tile-bitmap, %s) x%d" % ("bm_pair_kernel" if splits in (_cabi.H2_SPLITS_I8X2, _cabi.H2_SPLITS_I8X3) else "bm_mma_kernel", _cabi.SPLITS_NAME[splits], len(self.tensor_idx))] if self.tensor_idx else []) +
            (["fused_hops_gather_kernel (CSR gather) over %d hop(s)" % len(self.csr_idx)] if self.csr_idx else []))

    def reserve(self, d):
        """Scratch for rounds of width <= d: allocated here by the CALLER (torch's allocator) and bound to the handle
        (h2_graph_workspace_bytes / h2_graph_bind_workspace) — the round entry points never allocate or synchronise.
        Earlier, narrower workspaces stay alive until the plan is closed (rounds using them may still be in flight)."""
        if d <= self._ws_d:
            return
        d = (d + 3) // 4 * 4
        nbytes = int(lib().h2_graph_workspace_bytes(self._h, d))
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=self.hops[0].device)
        check(lib().h2_graph_bind_workspace(self._h, d, ptr(ws), nbytes))
        self._ws.append(ws)
        self._ws_d = d

    def run(self, x, out, offsets, d=None, stream=None):
        """out[:, offsets[h] : offsets[h]+d] = hops[h] @ x[:, :d]   (x, out may be column slices of one buffer)."""
        require_cuda(x, out)
        dt = {torch.float32: _cabi.H2_F32, torch.bfloat16: _cabi.H2_BF16}
        if x.dtype not in dt or out.dtype not in dt:
            raise ValueError("fused round: fp32 or bf16 feature matrices (accumulation is fp32 / exact int32 either way)")
        if x.stride(-1) != 1 or out.stride(-1) != 1:
            raise ValueError("x / out must be row-major (unit column stride)")
        d = x.shape[1] if d is None else d
        if x.shape[0] != self.n_cols or out.shape[0] != self.n_rows:
            raise ValueError(f"shape mismatch: x {tuple(x.shape)}, out {tuple(out.shape)}, adjacency "
                             f"[{self.n_rows}, {self.n_cols}]")
        if len(offsets) != len(self.hops):
            raise ValueError("one output column offset per hop")
        bf = x.dtype == torch.bfloat16 or out.dtype == torch.bfloat16
        q = 8 if bf else 4                      # rows are moved 16 bytes at a time
        if d % q or x.stride(0) % q or out.stride(0) % q or any(o % q or o < 0 or o + d > out.stride(0) for o in offsets) \
                or x.data_ptr() % 16 or out.data_ptr() % 16:
            raise ValueError(f"fused round: d={d}, leading dimensions and column offsets must be multiples of {q} "
                             "and the buffers 16-byte aligned")
        offs = (ctypes.c_int64 * len(offsets))(*offsets)
        self.reserve(d)
        if bf:    # BASELINE config 5: bf16 features in / out
            check(lib().h2_graph_round_ex(self._h, d, ptr(x), x.stride(0), dt[x.dtype], ptr(out), out.stride(0), dt[out.dtype],
                                          offs, stream_ptr(stream)))
        else:
            check(lib().h2_graph_round(self._h, d, ptr(x), x.stride(0), ptr(out), out.stride(0), offs, stream_ptr(stream)))
        return out

    def run_multi(self, x, x_offsets, out, offsets, d, stream=None):
        """Backward-style round: out[:, offsets[h] : +d] = hops[h] @ x[:, x_offsets[h] : +d] — every hop reads ITS OWN
        column slice of `x`.  With the symmetric hop adjacencies of undirected graphs this is the transposed product
        of the backward pass (dX = sum_h A_h^T dY_h; sum the slices with `sum_slices`)."""
        require_cuda(x, out)
        if x.dtype != torch.float32 or out.dtype != torch.float32 or x.stride(-1) != 1 or out.stride(-1) != 1:
            raise ValueError("fused round computes in fp32 on row-major buffers")
        if x.shape[0] != self.n_cols or out.shape[0] != self.n_rows or len(offsets) != len(self.hops) or \
                len(x_offsets) != len(self.hops):
            raise ValueError("shape mismatch / one offset per hop")
        xo = (ctypes.c_int64 * len(x_offsets))(*x_offsets)
        yo = (ctypes.c_int64 * len(offsets))(*offsets)
        self.reserve(d)
        check(lib().h2_graph_round_multi(self._h, d, ptr(x), x.stride(0), xo, ptr(out), out.stride(0), yo, stream_ptr(stream)))
        return out

    def run_parts(self, part_ptrs, bounds, ld_part, x_full, out, offsets, d, stream=None):
        """Round whose input is given as row shards (device pointers, possibly PEER memory of other ranks): the
        hop-boundary all-gather happens inside the first kernel of the round (h2_graph_round_parts).  `x_full`
        [n_cols, d] is scratch for the gathered copy and has the row type of the shards (fp32 or bf16: ld_part counts
        elements of that type); `out` may be fp32 or bf16.  The caller brackets the call with cross-rank barriers."""
        require_cuda(x_full, out)
        dt = {torch.float32: _cabi.H2_F32, torch.bfloat16: _cabi.H2_BF16}
        if x_full.dtype not in dt or out.dtype not in dt:
            raise ValueError("fused round: fp32 or bf16 feature matrices")
        P = len(part_ptrs)
        ptrs = (ctypes.c_void_p * P)(*part_ptrs)
        bnd = (ctypes.c_int64 * (P + 1))(*[int(b) for b in bounds])
        yo = (ctypes.c_int64 * len(offsets))(*offsets)
        self.reserve(d)
        if x_full.dtype == torch.bfloat16 or out.dtype == torch.bfloat16:
            if d % 8 or ld_part % 8 or x_full.stride(0) % 8 or out.stride(0) % 8 or any(o % 8 for o in offsets):
                raise ValueError("fused round with bf16 rows: d, leading dimensions and column offsets must be multiples of 8")
            check(lib().h2_graph_round_parts_ex(self._h, d, P, ptrs, bnd, ld_part, dt[x_full.dtype], ptr(x_full), x_full.stride(0),
                                                ptr(out), out.stride(0), dt[out.dtype], yo, stream_ptr(stream)))
        else:
            check(lib().h2_graph_round_parts(self._h, d, P, ptrs, bnd, ld_part, ptr(x_full), x_full.stride(0), ptr(out),
                                             out.stride(0), yo, stream_ptr(stream)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().h2_graph_destroy(self._h)
            self._h = ctypes.c_void_p()
            self._ws = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sum_slices(t, d, n_slices, g, accumulate=False, mask_src=None, stream=None):
    """g[:, :d] (+)= sum_s t[:, s*d:(s+1)*d], optionally zeroed where mask_src <= 0 (ReLU gradient)."""
    require_cuda(t, g, mask_src)
    check(lib().h2_sum_slices_f32(g.shape[0], d, n_slices, ptr(t), t.stride(0), ptr(g), g.stride(0), int(accumulate),
                                  ptr(mask_src), 0 if mask_src is None else mask_src.stride(0), stream_ptr(stream)))
    return g


def sparse_dense(feat, weight, bias=None, relu=False, out=None, out_col_off=0, stream=None):
    """SparseDense.call (_layers.py:45-52) [+ReLU]: out[:, off:off+p] = act(feat @ weight + bias)."""
    require_cuda(weight, bias, out)
    n, p = feat.n_rows, weight.shape[1]
    if feat.dense_shape[1] != weight.shape[0]:
        raise ValueError(f"feature dim {feat.dense_shape[1]} != kernel rows {weight.shape[0]}")
    weight = weight.contiguous()
    if out is None:
        out = torch.empty(n, p, dtype=torch.float32, device=weight.device)
    check(lib().h2_sparse_dense_f32(n, ptr(feat.rowptr), ptr(feat.col), ptr(feat.values), ptr(weight), p, ptr(bias),
                                    int(relu), ptr(out), out.stride(0), out_col_off, stream_ptr(stream)))
    return out


def dense(x, weight, bias=None, relu=False, out=None, out_col_off=0, stream=None, mode="tc"):
    """keras Dense (H2GCN.py:244-249): out[:, off:off+c] = act(x @ weight + bias), fp32 in / out.
    mode="tc" (default): tcgen05 3xTF32 kernel (fp32-equivalent); mode="simt": the fp32 SIMT kernel with the oracle's
    sequential-k summation order (parity mode)."""
    return matmul(x, weight, bias=bias, relu=relu, out=out, out_col_off=out_col_off, stream=stream, mode=mode)


def matmul(a, w, trans_a=False, trans_w=False, bias=None, relu=False, out=None, out_col_off=0, stream=None, mode="tc"):
    """out[:, off:off+n] = act(op(a) @ op(w) + bias): op = transpose when the flag is set (a: [k, m], w: [n, k]).  `a` and
    `w` may be column slices of wider row-major buffers.  The transposed forms are the classifier-side contractions of
    the training step (dW = final^T dlogits, dfinal = dlogits W^T)."""
    require_cuda(a, w, bias, out)
    if a.dtype != torch.float32 or w.dtype != torch.float32:
        raise ValueError("dense contraction computes in fp32")
    if a.stride(-1) != 1:
        a = a.contiguous()
    if w.stride(-1) != 1:
        w = w.contiguous()
    m, k = (a.shape[1], a.shape[0]) if trans_a else (a.shape[0], a.shape[1])
    n, kw = (w.shape[0], w.shape[1]) if trans_w else (w.shape[1], w.shape[0])
    if kw != k:
        raise ValueError(f"input dim {k} != kernel rows {kw}")
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    if mode == "simt":
        if trans_a or trans_w:
            raise ValueError("the SIMT parity kernel has no transposed forms")
        w = w.contiguous()
        check(lib().h2_dense_f32(m, k, n, ptr(a), a.stride(0), ptr(w), ptr(bias), int(relu), ptr(out), out.stride(0),
                                 out_col_off, stream_ptr(stream)))
    elif mode == "tc":
        check(lib().h2_dense_tc_f32(m, k, n, ptr(a), a.stride(0), int(trans_a), ptr(w), w.stride(0), int(trans_w), ptr(bias),
                                    int(relu), ptr(out), out.stride(0), out_col_off, stream_ptr(stream)))
    else:
        raise ValueError(f"unknown mode {mode}")
    return out


def relu_slice(x, out=None, relu=True, stream=None):
    require_cuda(x, out)
    n, d = x.shape
    if out is None:
        out = torch.empty(n, d, dtype=torch.float32, device=x.device)
    check(lib().h2_relu_slice_f32(n, d, ptr(x), x.stride(0), ptr(out), out.stride(0), int(relu), stream_ptr(stream)))
    return out


# ------------------------------------------------------------------------------------------------------------------
# host-buffer (end-to-end) graph handle
# ------------------------------------------------------------------------------------------------------------------
class HostGraph:
    """h2_graph_*: adjacency resident on the device, X in / Y out through HOST buffers every call (bench.py e2e)."""
    MODES = {"auto": 0, "csr": 1, "tensor": 2}

    def __init__(self, hops_host, n_rows, n_cols, d_max, dinv_host=None, row_begin=0, mode="auto", splits=None):
        """hops_host: list of (rowptr int64, col int32, val fp32) numpy arrays; dinv_host: optional list of fp32
        [n_cols] scale vectors (normalised binary patterns) enabling the tensor-core format."""
        import numpy as np
        self._keep = [(np.ascontiguousarray(r, dtype=np.int64), np.ascontiguousarray(c, dtype=np.int32),
                       np.ascontiguousarray(v, dtype=np.float32)) for r, c, v in hops_host]
        H = len(self._keep)
        self._dinv = [None if dv is None else np.ascontiguousarray(dv, dtype=np.float32)
                      for dv in (dinv_host if dinv_host is not None else [None] * H)]
        arr = lambda k: (ctypes.c_void_p * H)(*[a[k].ctypes.data for a in self._keep])
        dv = (ctypes.c_void_p * H)(*[None if a is None else a.ctypes.data for a in self._dinv])
        self._h = ctypes.c_void_p()
        self.n_rows, self.n_cols, self.n_hops = n_rows, n_cols, H
        check(lib().h2_graph_create(n_rows, n_cols, H, arr(0), arr(1), arr(2), dv, row_begin, d_max, self.MODES[mode],
                                    _cabi.splits_code(splits), ctypes.byref(self._h)))

    def round(self, x_host, y_host, stream=None):
        """x_host [n_cols, d] / y_host [n_rows, H*d]: (pinned) host torch tensors or numpy arrays, fp32, contiguous."""
        d = x_host.shape[1]
        xp = x_host.data_ptr() if hasattr(x_host, "data_ptr") else x_host.ctypes.data
        yp = y_host.data_ptr() if hasattr(y_host, "data_ptr") else y_host.ctypes.data
        check(lib().h2_graph_round_host(self._h, d, xp, yp, stream_ptr(stream)))
        return y_host

    def round_device(self, x, y, offsets, d=None, stream=None):
        """Same round on device tensors (no copies, no sync)."""
        require_cuda(x, y)
        d = x.shape[1] if d is None else d
        offs = (ctypes.c_int64 * self.n_hops)(*offsets)
        check(lib().h2_graph_round(self._h, d, ptr(x), x.stride(0), ptr(y), y.stride(0), offs, stream_ptr(stream)))
        return y

    def close(self):
        if self._h:
            lib().h2_graph_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
