"""npz dataset plug-in — the on-disk format of the reference's `npz-datasets/` (DeepRobust-style
sparse-graph npz: adj_data/adj_indices/adj_indptr/adj_shape, attr_* , labels [, idx_train/idx_val/idx_test]) with the
adjacency conventions of `npz-datasets/dataset.py:28-55`: symmetrised (A + A^T), binarised, zero diagonal, float32.
BASELINE config 2 names "Cora (npz)"; the archives are download-only, so the tests write their own files."""
import numpy as np
import scipy.sparse as sp

from ._dataset import GraphData


def load_npz(filename):
    with np.load(filename, allow_pickle=True) as z:
        z = dict(z)
    adj = sp.csr_matrix((z["adj_data"], z["adj_indices"], z["adj_indptr"]), shape=tuple(z["adj_shape"]))
    if "attr_data" in z:
        feats = sp.csr_matrix((z["attr_data"], z["attr_indices"], z["attr_indptr"]), shape=tuple(z["attr_shape"]))
    elif "attr_matrix" in z:
        feats = sp.csr_matrix(z["attr_matrix"])
    else:
        feats = sp.identity(adj.shape[0], dtype=np.float32, format="csr")
    labels = z.get("labels")
    splits = {k: z[k] for k in ("idx_train", "idx_val", "idx_test") if k in z}
    return adj, feats.astype(np.float32), labels, splits


def canonical_adjacency(adj):
    """dataset.py:30-55: adj + adj.T, entries > 1 -> 1, diagonal 0, float32 CSR without explicit zeros."""
    adj = sp.csr_matrix(adj)
    adj = (adj + adj.T).tolil()
    adj[adj > 1] = 1
    adj.setdiag(0)
    adj = adj.astype("float32").tocsr()
    adj.eliminate_zeros()
    adj.sort_indices()
    assert abs(adj - adj.T).sum() == 0, "Input graph is not symmetric"
    return adj


class NpzData(GraphData):
    def __init__(self, filename, device="cuda"):
        adj, feats, labels, splits = load_npz(filename)
        adj = canonical_adjacency(adj)
        n = adj.shape[0]
        if labels is None:
            onehot = np.zeros((n, 1))
        else:
            labels = np.asarray(labels).astype(np.int64)
            onehot = np.eye(int(labels.max()) + 1)[labels]
        super().__init__(adj, feats.tolil(), onehot, device=device)
        for name, key in (("train", "idx_train"), ("val", "idx_val"), ("test", "idx_test")):
            if key in splits:
                m = np.zeros(n, dtype=bool)
                m[splits[key]] = True
                self._dense_data[f"{name}_mask"] = m
                y = np.zeros_like(onehot)
                y[m] = onehot[m]
                self._dense_data[f"y_{name}"] = y


def add_subparser_args(parser):
    sub = parser.add_argument_group("npz Format Data Arguments (datasets/npz.py)")
    sub.add_argument("--dataset", type=str, required=True, help="path of the .npz file")
    parser.function_hooks["argparse"].appendleft(argparse_callback)


def argparse_callback(args):
    args.objects["dataset"] = NpzData(args.dataset)
    print(f"===> Dataset loaded: {args.dataset}")
