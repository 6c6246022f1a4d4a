"""Planetoid loader + CLI surface.  The loader is host code (CPU test, against the reference's own loader output stored
in the golden files, only where /root/reference's data files exist); the CLI test runs the whole argv -> hooks ->
preprocessing -> forward chain on the GPU with a synthetic Planetoid-format dataset written to a temp directory."""
import os
import pickle
from collections import defaultdict

import numpy as np
import pytest
import scipy.sparse as sp

from tests import util

REF_DATA = "/root/reference/baselines/gcn/gcn/data"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_DATA, "ind.cora.graph")), reason="reference data files not mounted")
@pytest.mark.parametrize("name", ["cora", "citeseer"])
def test_planetoid_loader_matches_reference_loader(name):
    from h2gcn_b200.datasets._dataset import PlanetoidData
    z = util.load_golden("planetoid_" + name)
    d = PlanetoidData("ind." + name, REF_DATA, val_size=500, device="cpu")
    adj = sp.csr_matrix(d.sparse_adj)
    adj.sort_indices()
    assert np.array_equal(adj.indptr, z["adj_indptr"]) and np.array_equal(adj.indices, z["adj_indices"])
    assert np.array_equal(adj.data, z["adj_data"])
    feat = sp.csr_matrix(d.features)
    feat.sort_indices()
    assert np.array_equal(feat.indptr, z["feat_indptr"]) and np.array_equal(feat.indices, z["feat_indices"])
    assert np.array_equal(feat.data.astype(np.float32), z["feat_data"])
    assert d.num_labels == int(z["num_labels"])
    with np.errstate(divide="ignore"):
        d.row_normalize_features()
    f = sp.coo_matrix(d.features)
    order = np.lexsort((f.col, f.row))
    assert np.array_equal(f.data[order].astype(np.float32), z["featn_vals"])


def write_planetoid(path, name, n=300, n_feat=40, n_class=4, seed=0):
    """A tiny dataset in the Planetoid on-disk format (ind.<name>.{x,y,tx,ty,allx,ally,graph,test.index})."""
    from h2gcn_b200.utils import synth
    rng = np.random.default_rng(seed)
    adj = synth.uniform_graph(n, 900, seed=seed)
    graph = defaultdict(list)
    for i in range(n):
        graph[i] = list(adj.indices[adj.indptr[i]:adj.indptr[i + 1]])
    feats = sp.random(n, n_feat, density=0.2, random_state=np.random.RandomState(seed), dtype=np.float32).tocsr()
    labels = np.eye(n_class)[rng.integers(0, n_class, size=n)]
    n_train, n_test = 40, 60
    test_idx = np.arange(n - n_test, n)
    objs = {"x": feats[:n_train], "y": labels[:n_train], "allx": feats[:n - n_test], "ally": labels[:n - n_test],
            "tx": feats[test_idx], "ty": labels[test_idx], "graph": graph}
    for k, v in objs.items():
        with open(os.path.join(path, f"ind.{name}.{k}"), "wb") as f:
            pickle.dump(v, f)
    with open(os.path.join(path, f"ind.{name}.test.index"), "w") as f:
        f.write("\n".join(str(i) for i in rng.permutation(test_idx)))
    return adj, feats, labels


@pytest.mark.gpu
def test_cli_forward_on_synthetic_planetoid(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from h2gcn_b200 import run_experiments
    from oracle import h2gcn_oracle as O
    adj, feats, labels = write_planetoid(str(tmp_path), "syn")
    args, logits = run_experiments.main(["H2GCN", "planetoid", "--dataset", "ind.syn", "--dataset_path", str(tmp_path),
                                         "--network_setup", "M64-R-T1-G-V-T2-G-V-C1-C2-D0.5-MO", "--adj_nhood", "1", "2"])
    assert logits.shape == (300, 4)
    for key in ("tensors", "model", "train_step", "test_step", "predict_step", "embed_step", "dataset"):
        assert key in args.objects
    # train_step (H2GCN.py:66-74) runs on the same kernels: the loss goes down on the synthetic labels
    tensors = args.objects["tensors"]
    tensors["train_mask"] = torch.ones_like(tensors["train_mask"])
    tensors["y_train"] = tensors["y_all"]
    losses = [args.objects["train_step"](**tensors)["train_loss"] for _ in range(60)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5]) - 0.02, losses[::10]
    # same forward through the oracle with the model's weights
    model, t = args.objects["model"], args.objects["tensors"]
    logits = args.objects["predict_step"](**t)              # with the trained weights
    weights = [w.cpu().numpy() for w in model.trainable_variables]
    fi = t["features"].indices.cpu().numpy()
    hops = [(h.indices[:, 0].cpu().numpy(), h.indices[:, 1].cpu().numpy(), h.values.cpu().numpy()) for h in t["adj_hops"]]
    from h2gcn_b200.models import parse_network_setup
    conf = parse_network_setup("M64-R-T1-G-V-T2-G-V-C1-C2-D0.5-MO", 4, _dense_units=64, _dropout_rate=0.5)
    ref = O.forward(conf, weights, (fi[:, 0], fi[:, 1], t["features"].values.cpu().numpy()), 300, hops)
    assert util.rel_err(logits.cpu().numpy(), ref) <= 1e-4


def test_npz_adjacency_conventions(tmp_path):
    """npz loader (reference npz-datasets/dataset.py:28-55): symmetrise, binarise, zero the diagonal."""
    from h2gcn_b200.datasets.npz import NpzData, canonical_adjacency
    a = sp.csr_matrix(np.array([[1, 1, 0, 0], [0, 0, 2, 0], [0, 1, 0, 0], [0, 0, 0, 0]], dtype=np.float64))
    c = canonical_adjacency(a).toarray()
    assert (c == np.array([[0, 1, 0, 0], [1, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 0]])).all()
    feats = sp.csr_matrix(np.arange(12, dtype=np.float32).reshape(4, 3))
    fn = str(tmp_path / "toy.npz")
    np.savez(fn, adj_data=a.data, adj_indices=a.indices, adj_indptr=a.indptr, adj_shape=a.shape,
             attr_data=feats.data, attr_indices=feats.indices, attr_indptr=feats.indptr, attr_shape=feats.shape,
             labels=np.array([0, 1, 1, 2]), idx_train=np.array([0, 1]), idx_val=np.array([2]), idx_test=np.array([3]))
    d = NpzData(fn, device="cpu")
    assert d.num_labels == 3 and d.num_samples == 4 and d.feature_dim == 3
    assert (sp.csr_matrix(d.sparse_adj).toarray() == c).all()
    assert list(d.train_mask) == [True, True, False, False] and d.y_test[3, 2] == 1


@pytest.mark.gpu
def test_precompute_cache_round_trip(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from h2gcn_b200.datasets._dataset import GraphData
    z = util.load_golden("planetoid_citeseer")
    def tensors():
        data = GraphData(util.raw_adj(z), util.raw_feat(z).tolil(), device="cuda")
        data.adj_remove_eye()
        return data.getTensors(getAdjNormHops=["1", "2"], cache_dir=str(tmp_path))
    first = tensors()
    assert len(os.listdir(tmp_path)) == 1
    second = tensors()                       # served from the cache
    for a, b, (gr, gc, gv) in zip(first.adj_hops, second.adj_hops, util.golden_hops(z)):
        assert torch.equal(a.rowptr, b.rowptr) and torch.equal(a.col, b.col) and torch.equal(a.values, b.values)
        assert np.array_equal(b.values.cpu().numpy(), gv) and np.array_equal(b.indices[:, 1].cpu().numpy(), gc)
