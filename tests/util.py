"""Shared test helpers (CPU side).  The oracle is imported ONLY here and in the tests (oracle/ header)."""
import os

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
TINY = ["path4", "star5", "tri_tail", "isolated", "selfloops", "rand40"]
PLANETOID = ["cora", "citeseer"]


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def all_golden_names():
    return ["tiny_" + t for t in TINY] + ["planetoid_" + p for p in PLANETOID]


def raw_adj(z):
    n = len(z["adj_indptr"]) - 1
    return sp.csr_matrix((z["adj_data"], z["adj_indices"], z["adj_indptr"]), shape=(n, n))


def raw_feat(z):
    return sp.csr_matrix((z["feat_data"], z["feat_indices"], z["feat_indptr"]), shape=tuple(z["feat_shape"]))


def golden_hops(z):
    """[(rows, cols, vals)] of the reference's adj_hops tensors."""
    out, h = [], 0
    while f"hop{h}_rows" in z.files:
        out.append((z[f"hop{h}_rows"].astype(np.int64), z[f"hop{h}_cols"].astype(np.int64), z[f"hop{h}_vals"]))
        h += 1
    return out


def coo_to_csr(rows, cols, n):
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, np.asarray(rows, dtype=np.int64) + 1, 1)
    return np.cumsum(rowptr), np.asarray(cols, dtype=np.int32)


def setups_in(z):
    return sorted({k.split("/")[0] for k in z.files if k.endswith("/setup")})


def weights_of(z, setup):
    ws, i = [], 0
    while f"{setup}/W{i}" in z.files:
        ws.append(z[str(z[f"{setup}/W{i}"])])
        i += 1
    return ws


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    scale = max(np.abs(ref).max(), 1e-30) if ref.size else 1.0
    return float(np.abs(got - ref).max() / scale) if ref.size else 0.0


def row_rel_err(got, ref, floor=0.0):
    """Row-wise relative error: max over rows of ||got_row - ref_row||_inf / ||ref_row||_inf.  Unlike `rel_err` (which
    divides by the max-abs of the WHOLE reference tensor) rows with small norms count as much as the largest row, so an
    operand format that spends its bits on the globally largest rows shows up here.  Rows whose reference norm is
    <= `floor` (exact zeros: zero-degree rows) are compared absolutely against `floor`-scaled tolerance by the caller."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    rn = np.abs(ref).max(axis=1)
    en = np.abs(got - ref).max(axis=1)
    live = rn > floor
    if not live.any():
        return 0.0
    return float((en[live] / rn[live]).max())


def i8_block_quantize(xs, pieces):
    """numpy model of the int8 operand format of the tensor-core path (csrc/bitmap_mma.cu, bm_pack_i8_kernel): `xs` =
    fp32 diag(dinv) X.  Returns (dequantised fp64 matrix, step, exponents t per row).  Not part of the
    reference — it lets a test check the integer pipeline EXACTLY (the int32 accumulation has no rounding)."""
    xs = np.asarray(xs, dtype=np.float32)
    n = xs.shape[0]
    R = {2: 32639, 3: 8355711}[pieces]
    expo = lambda v: int(np.frexp(np.float32(v))[1]) - 1          # floor(log2 v) for normal fp32 v > 0
    gm = float(np.abs(xs).max()) if xs.size else 0.0
    eg = max(expo(gm), -96) if gm > 0 else -96
    mg = np.abs(xs).max(axis=1) if xs.size else np.zeros(n, dtype=np.float32)       # one exponent per ROW of X'
    t = np.array([min(max(expo(m) - eg + 6, 0), 6) if m > 0 else 0 for m in mg], dtype=np.int64)
    trow = t
    mult = np.float32(np.ldexp(np.float32(R), 5 - eg))
    scale = (mult * np.ldexp(np.float32(1), -trow).astype(np.float32)).astype(np.float32)
    q = np.rint((xs * scale[:, None]).astype(np.float32)).astype(np.int64)
    assert np.abs(q).max(initial=0) <= R
    step = np.float32(np.ldexp(np.float32(1), eg - 5) / np.float32(R))
    return q.astype(np.float64) * np.ldexp(1.0, trow)[:, None] * float(step), float(step), t
