"""Data container + adjacency-power precompute — mirrors the hot-path part of `h2gcn/datasets/_dataset.py`.

Kept names: `TransformSPAdj.{NType, normalize, addEye, removeEye, nhoodSplit}`, `PlanetoidData(dataset_str,
dataset_path, val_size)` with `.adj_remove_eye()`, `.row_normalize_features()`, `.getTensors(getAdjNormHops=...)`,
`.sparse2Tensor`.  File loading and label/mask bookkeeping are host Python (numpy/scipy, no networkx); every step
that the scope table marks as hot path (removeEye, nhoodSplit, normalize, the COO/CSR
canonicalisation) runs on the GPU through the C-ABI and returns device tensors.
"""
import pickle as pkl
import sys
import warnings
from argparse import Namespace
from enum import Enum
from itertools import chain

import numpy as np
import scipy.sparse as sp
import torch

from .. import ops
from ..ops import SparseTensor


def _pattern(sp_or_scipy, device):
    """-> (rowptr int64, col int32) device tensors of a canonical (sorted, deduplicated) CSR pattern."""
    if isinstance(sp_or_scipy, SparseTensor):
        return sp_or_scipy.rowptr, sp_or_scipy.col
    m = sp.csr_matrix(sp_or_scipy)
    m.sum_duplicates()
    m.sort_indices()
    return (torch.from_numpy(m.indptr.astype(np.int64)).to(device), torch.from_numpy(m.indices.astype(np.int32)).to(device))


def _ones(n, device):
    return torch.ones(n, dtype=torch.float32, device=device)


def _merge_patterns(mats, n_rows, n_cols):
    """Sum of binary patterns with disjoint supports (`sum([adjSplits[i] for i in elem])`, _dataset.py:571-572) ->
    union pattern in canonical order.  Index plumbing only (sort of 64-bit keys)."""
    idx = torch.cat([m.indices for m in mats], dim=0)
    key = idx[:, 0] * n_cols + idx[:, 1]
    key, counts = torch.unique(key, sorted=True, return_counts=True)
    rows = torch.div(key, n_cols, rounding_mode="floor")
    col = (key - rows * n_cols).to(torch.int32)
    rowptr = torch.zeros(n_rows + 1, dtype=torch.int64, device=key.device)
    rowptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n_rows), 0)
    return SparseTensor(rowptr, col, counts.to(torch.float32), (n_rows, n_cols))


class TransformSPAdj:
    """GPU versions of the reference's sparse adjacency transforms (_dataset.py:101-158)."""

    class NType(Enum):
        ORDINARY = 0
        SYM_NORMALIZED = 1
        RW_NORMALIZED = 2
        CHEBY = 3

    device = "cuda"

    @classmethod
    def normalize(cls, adj, Ntype):
        """_dataset.py:109-124.  `adj`: SparseTensor (binary pattern, values ignored unless ORDINARY) or scipy matrix.
        Returns a SparseTensor whose values are fp32(fp64 product), with `.dinv` set for SYM."""
        if not isinstance(adj, SparseTensor):
            adj = SparseTensor.from_scipy(adj, cls.device)
        if Ntype == cls.NType.ORDINARY:
            return adj
        if adj.nnz and not bool((adj.values == 1).all()):
            raise ValueError("normalize: only binary patterns are supported on the GPU path (nhoodSplit output)")
        n_rows, n_cols = adj.dense_shape
        if Ntype == cls.NType.SYM_NORMALIZED:
            val, _, d32 = ops.sym_normalize(adj.rowptr, adj.col, n_cols=n_cols)
            return SparseTensor(adj.rowptr, adj.col, val, adj.dense_shape, dinv=d32)
        if Ntype == cls.NType.RW_NORMALIZED:
            return SparseTensor(adj.rowptr, adj.col, ops.rw_normalize(adj.rowptr), adj.dense_shape)
        raise NotImplementedError("CHEBY normalisation is not used by any H2GCN config (SURVEY.md §2 #4: out of scope)")

    @staticmethod
    def addEye(adj):
        adj = sp.csr_matrix(adj).tolil(copy=True)
        adj.setdiag(1)
        return adj.tocsr()

    @classmethod
    def removeEye(cls, adj):
        """_dataset.py:132-136.  scipy in -> scipy out (the container keeps host matrices like the reference), the
        diagonal is removed by the GPU kernel; SparseTensor in -> SparseTensor out."""
        if isinstance(adj, SparseTensor):
            rp, col, val = ops.remove_eye(adj.rowptr, adj.col, adj.values)
            return SparseTensor(rp, col, val, adj.dense_shape)
        m = sp.csr_matrix(adj)
        m.sum_duplicates()
        m.sort_indices()
        dev = torch.device(cls.device)
        rp, col, val = ops.remove_eye(torch.from_numpy(m.indptr.astype(np.int64)).to(dev),
                                      torch.from_numpy(m.indices.astype(np.int32)).to(dev),
                                      torch.from_numpy(m.data.astype(np.float32)).to(dev))
        return sp.csr_matrix((val.cpu().numpy().astype(m.dtype), col.cpu().numpy(), rp.cpu().numpy()), shape=m.shape)

    @classmethod
    def nhoodSplit(cls, adj, nhood):
        """_dataset.py:138-158 for nhood <= 2: [I, P1, P2] as device SparseTensors with values 1.0.
        P1 = pattern(A) (A must be free of self loops, as after adj_remove_eye), P2 = vertices at distance exactly 2.
        Like the reference the list is SHORT when a power adds nothing (:151-153)."""
        dev = torch.device(cls.device)
        rowptr, col = _pattern(adj, dev)
        n = rowptr.numel() - 1
        assert (adj.dense_shape if isinstance(adj, SparseTensor) else adj.shape) == (n, n)
        if nhood > 2:
            raise NotImplementedError("nhoodSplit: nhood > 2 is not used by any H2GCN config (SURVEY.md §8a a2)")
        eye = SparseTensor(torch.arange(n + 1, dtype=torch.int64, device=dev),
                           torch.arange(n, dtype=torch.int32, device=dev), _ones(n, dev), (n, n))
        out = [eye]
        if nhood >= 1 and n > 0:  # the first power always grows (0 -> n + nnz entries), even for an edgeless graph
            out.append(SparseTensor(rowptr, col, _ones(col.numel(), dev), (n, n)))
            if nhood >= 2 and col.numel() > 0:
                rp2, col2 = ops.hop2_pattern(rowptr, col)
                if col2.numel() > 0:
                    out.append(SparseTensor(rp2, col2, _ones(col2.numel(), dev), (n, n)))
        return out


class GraphData:
    """In-memory graph container with the reference's preprocessing API (used for synthetic graphs and as the base of
    PlanetoidData).  adj / features are host scipy matrices like in the reference; tensors live on `device`."""

    def __init__(self, adj, features, labels_onehot=None, device="cuda"):
        self._sparse_data = dict(sparse_adj=sp.csr_matrix(adj), features=features)
        self._dense_data = dict()
        n = adj.shape[0]
        if labels_onehot is None:
            labels_onehot = np.zeros((n, 1))
        self._dense_data["y_all"] = labels_onehot
        for k in ("train_mask", "val_mask", "test_mask", "wild_mask"):
            self._dense_data[k] = np.zeros(n, dtype=bool)
        for k in ("y_train", "y_val", "y_test", "y_wild"):
            self._dense_data[k] = np.zeros_like(labels_onehot)
        self.device = device
        self.preprocessedAdj = None
        self.preprocessedFeature = None

    # ---- reference-style attribute access --------------------------------------------------------------------
    @property
    def sparse_adj(self):
        return self._sparse_data["sparse_adj"]

    @sparse_adj.setter
    def sparse_adj(self, v):
        self._sparse_data["sparse_adj"] = v

    @property
    def features(self):
        return self._sparse_data["features"]

    @features.setter
    def features(self, v):
        self._sparse_data["features"] = v

    def __getattr__(self, name):  # only called when normal lookup fails: y_train, train_mask, ...
        dd = self.__dict__.get("_dense_data", {})
        if name in dd:
            return dd[name]
        raise AttributeError(name)

    @property
    def labels(self):
        idx, labels = np.where(self.y_all)
        labels = labels.astype(np.int32)
        if len(idx) != self.num_samples:  # Citeseer: unlabeled vertices -> -1 (_dataset.py:343-350)
            part = labels
            labels = np.zeros(self.num_samples) - 1
            labels[idx] = part
        return labels

    @property
    def num_labels(self):
        return self.y_all.shape[1]

    @property
    def num_samples(self):
        return self.features.shape[0]

    @property
    def feature_dim(self):
        return self.features.shape[1]

    # ---- preprocessing (H2GCN.preprocessing_data, H2GCN.py:46-54) ---------------------------------------------
    def adj_add_eye(self):
        self.sparse_adj = TransformSPAdj.addEye(self.sparse_adj)
        self.preprocessedAdj = True

    def adj_remove_eye(self):
        TransformSPAdj.device = self.device
        self.sparse_adj = TransformSPAdj.removeEye(self.sparse_adj)
        self.preprocessedAdj = True

    def get_eye(self):
        return sp.identity(self.num_samples, dtype=self.sparse_adj.dtype)

    def row_normalize_features(self):
        """_dataset.py:502-509 (host, O(nnz(X)), once per process): X <- diag(1/rowsum) X, inf -> 0, dtype of X."""
        feats = self.features  # container type kept: scipy's LIL and CSR row sums differ in the last fp32 ulp
        with np.errstate(divide="ignore"):
            inv = np.power(np.asarray(feats.sum(1)).ravel(), -1)
        inv[np.isinf(inv)] = 0.
        self.features = sp.diags(inv) @ feats
        self.preprocessedFeature = True

    @classmethod
    def sparse2Tensor(cls, spmat, dtype=np.float32, device="cuda"):
        """_dataset.py:528-535: canonical (row-major, ascending column) fp32 sparse tensor on the device."""
        if isinstance(spmat, list):
            return [cls.sparse2Tensor(x, dtype, device) for x in spmat]
        if isinstance(spmat, SparseTensor):
            return spmat
        return SparseTensor.from_scipy(spmat, device)

    def getTensors(self, getDenseAdj=False, getAdjHops=None, getAdjNormHops=None,
                   normType=TransformSPAdj.NType.SYM_NORMALIZED, dtype=np.float32, cache_dir=None):
        """_dataset.py:537-584.  `adj_hops` is a list of device SparseTensors in the order of `getAdjNormHops`.
        cache_dir: optional on-disk cache of the (merged, un-normalised) hop patterns keyed by a hash of the adjacency
        (datasets/_cache.py); the normalisation is recomputed on load, so cached and fresh results are bit-identical."""
        dev = torch.device(self.device)
        TransformSPAdj.device = self.device
        tensors = Namespace()
        for key, value in self._sparse_data.items():
            setattr(tensors, key, self.sparse2Tensor(value, dtype, dev))
        if getDenseAdj:
            raise NotImplementedError("dense adjacency tensors are not part of the H2GCN path (SURVEY.md §2 #4)")
        tensors.adj = tensors.sparse_adj
        if getAdjHops:
            # _dataset.py:551-558: un-normalised merged hop patterns.  The reference densifies them into one
            # [N, H, N] constant; the only models that ask for them (setups without a G layer: the MLP configs
            # M64-R-MO, M64-D-MO, ...) never read the tensor, so they stay sparse device tensors here (same hops,
            # same order, values 1.0) instead of an O(N^2) array.
            hops = [[int(x) for x in str(elem).split(",")] for elem in getAdjHops]
            splits = TransformSPAdj.nhoodSplit(tensors.adj, max(chain(*hops)))
            n = self.num_samples
            tensors.adj_hops = [splits[e[0]] if len(e) == 1 else _merge_patterns([splits[i] for i in e], n, n) for e in hops]
        if getAdjNormHops:
            hops = [[int(x) for x in str(elem).split(",")] for elem in getAdjNormHops]
            hop_max = max(chain(*hops))
            if normType == TransformSPAdj.NType.CHEBY:
                raise NotImplementedError("CHEBY hops are not used by any H2GCN config")
            n = self.num_samples
            merged = key = None
            if cache_dir is not None:
                from . import _cache
                key = _cache.graph_key(tensors.adj.rowptr.cpu().numpy(), tensors.adj.col.cpu().numpy(), getAdjNormHops, "pattern")
                merged = _cache.load_hops(cache_dir, key, dev)
            if merged is None:
                splits = TransformSPAdj.nhoodSplit(tensors.adj, hop_max)
                merged = [splits[e[0]] if len(e) == 1 else _merge_patterns([splits[i] for i in e], n, n) for e in hops]
                if cache_dir is not None:
                    _cache.save_hops(cache_dir, key, merged)
            tensors.adj_hops = [TransformSPAdj.normalize(x, normType) for x in merged]
        for key, value in self._dense_data.items():
            setattr(tensors, key, torch.as_tensor(np.asarray(value), dtype=torch.float32, device=dev))
        tensors.labels = torch.as_tensor(self.labels, device=dev)
        tensors.preprocessedAdj = self.preprocessedAdj
        tensors.preprocessedFeature = self.preprocessedFeature
        return tensors


class PlanetoidData(GraphData):
    """Planetoid `ind.<name>.{x,y,tx,ty,allx,ally,graph,test.index}` loader (_dataset.py:195-334)."""

    def __init__(self, dataset_str, dataset_path, val_size=None, device="cuda"):
        self.dataset_str = dataset_str
        self.dataset_path = dataset_path
        self.device = device
        self._sparse_data, self._dense_data = dict(), dict()
        self.preprocessedAdj = None
        self.preprocessedFeature = None
        self.load_data(dataset_str, dataset_path, val_size=val_size)

    @staticmethod
    def parse_index_file(filename):
        return [int(line.strip()) for line in open(filename)]

    @staticmethod
    def sample_mask(idx, l):
        mask = np.zeros(l, dtype=bool)
        mask[idx] = True
        return mask

    @staticmethod
    def graphDict2Adj(graph):
        """Undirected, unweighted adjacency of a {vertex: [neighbours]} dict, vertices 0..len-1; a self loop gives a
        diagonal 1 (what nx.adjacency_matrix(nx.from_dict_of_lists(graph), nodelist=range(n)) yields, :184-186)."""
        n = len(graph)
        src = np.fromiter(chain.from_iterable([u] * len(vs) for u, vs in graph.items()), dtype=np.int64)
        dst = np.fromiter(chain.from_iterable(graph.values()), dtype=np.int64)
        rows = np.concatenate([src, dst])
        cols = np.concatenate([dst, src])
        m = sp.csr_matrix((np.ones(len(rows), dtype=np.int64), (rows, cols)), shape=(n, n))
        m.sum_duplicates()
        m.data[:] = 1
        m.sort_indices()
        return m

    def load_data(self, dataset_str, dataset_path="data", save_plot=None, val_size=None):
        objects = []
        for name in ['x', 'y', 'tx', 'ty', 'allx', 'ally', 'graph']:
            with open("{}/{}.{}".format(dataset_path, dataset_str, name), 'rb') as f:
                objects.append(pkl.load(f, encoding='latin1') if sys.version_info > (3, 0) else pkl.load(f))
        x, y, tx, ty, allx, ally, graph = tuple(objects)
        test_idx_reorder = self.parse_index_file("{}/{}.test.index".format(dataset_path, dataset_str))
        test_idx_range = np.sort(test_idx_reorder)
        full = range(min(test_idx_reorder), max(test_idx_reorder) + 1)
        if len(full) != len(test_idx_range):  # citeseer: isolated test vertices become zero rows (:226-239)
            tx_ext = sp.lil_matrix((len(full), x.shape[1]))
            tx_ext[test_idx_range - min(test_idx_range), :] = tx
            tx = tx_ext
            ty_ext = np.zeros((len(full), y.shape[1]))
            ty_ext[test_idx_range - min(test_idx_range), :] = ty
            ty = ty_ext
            self.non_valid_samples = set(full) - set(test_idx_range)
        else:
            self.non_valid_samples = set()
        features = sp.vstack((allx, tx)).tolil()
        features[test_idx_reorder, :] = features[test_idx_range, :]
        adj = self.graphDict2Adj(graph).astype(np.float32)
        labels = np.vstack((ally, ty))
        labels[test_idx_reorder, :] = labels[test_idx_range, :]
        self.non_valid_samples = self.non_valid_samples.union(set(list(np.where(labels.sum(1) == 0)[0])))

        n = labels.shape[0]
        train_mask = self.sample_mask(range(len(y)), n)
        test_mask = self.sample_mask(test_idx_range.tolist(), n)
        val_mask = ~(train_mask | test_mask)
        if val_size is not None:
            if np.sum(val_mask) > val_size:
                val_mask = self.sample_mask(range(len(y), len(y) + val_size), n)
            else:
                print(f"Val set size set to {np.sum(val_mask)} due to insufficient samples.")
        wild_mask = ~(train_mask | val_mask | test_mask)
        for n_i in self.non_valid_samples:
            for m, what in ((train_mask, "training"), (test_mask, "test"), (val_mask, "val")):
                if m[n_i]:
                    warnings.warn(f"Non valid samples detected in {what} set")
                    m[n_i] = False
                    break
            wild_mask[n_i] = False
        self._sparse_data["sparse_adj"] = adj
        self._sparse_data["features"] = features
        self._dense_data["y_all"] = labels
        for name, m in (("train", train_mask), ("val", val_mask), ("test", test_mask), ("wild", wild_mask)):
            yy = np.zeros(labels.shape)
            yy[m, :] = labels[m, :]
            self._dense_data[f"{name}_mask"] = m
            self._dense_data[f"y_{name}"] = yy
        return adj, features
