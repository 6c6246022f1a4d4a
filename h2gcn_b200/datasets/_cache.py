"""On-disk cache of the adjacency-power precompute: the hop patterns of a graph are stored as
one .npz blob keyed by a hash of the adjacency, so a second process skips nhoodSplit.  Values are NOT stored — they are
a cheap function of the pattern (h2_sym_normalize) and are recomputed bit-exactly on load."""
import hashlib
import os

import numpy as np
import torch

from ..ops import SparseTensor

FORMAT_VERSION = 1


def graph_key(rowptr, col, spec, norm):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(rowptr, dtype=np.int64).tobytes())
    h.update(np.ascontiguousarray(col, dtype=np.int32).tobytes())
    h.update(repr((list(spec), str(norm), FORMAT_VERSION)).encode())
    return h.hexdigest()[:32]


def save_hops(cache_dir, key, hops):
    os.makedirs(cache_dir, exist_ok=True)
    blob = {"version": np.array(FORMAT_VERSION), "n_hops": np.array(len(hops))}
    for k, h in enumerate(hops):
        blob[f"rowptr{k}"] = h.rowptr.cpu().numpy()
        blob[f"col{k}"] = h.col.cpu().numpy()
        blob[f"shape{k}"] = np.array(h.dense_shape, dtype=np.int64)
    tmp = os.path.join(cache_dir, f".{key}.{os.getpid()}.npz")
    np.savez(tmp, **blob)
    os.replace(tmp, os.path.join(cache_dir, key + ".npz"))


def load_hops(cache_dir, key, device):
    path = os.path.join(cache_dir, key + ".npz")
    if not os.path.exists(path):
        return None
    with np.load(path) as z:
        if int(z["version"]) != FORMAT_VERSION:
            return None
        out = []
        for k in range(int(z["n_hops"])):
            col = torch.from_numpy(z[f"col{k}"]).to(device)
            out.append(SparseTensor(torch.from_numpy(z[f"rowptr{k}"]).to(device), col,
                                    torch.ones(col.numel(), dtype=torch.float32, device=device), tuple(z[f"shape{k}"])))
        return out
