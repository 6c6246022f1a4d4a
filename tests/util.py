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
