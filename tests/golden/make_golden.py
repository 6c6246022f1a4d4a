#!/usr/bin/env python
"""Generate the golden vectors in this directory by EXECUTING THE REFERENCE'S OWN PYTHON, unmodified.

Run here (container with /root/reference mounted):  python tests/golden/make_golden.py

What runs from /root/reference (imported, never copied):
  h2gcn/datasets/_dataset.py : PlanetoidData.load_data, row_normalize_features, adj_remove_eye, getTensors
                               (-> TransformSPAdj.nhoodSplit / normalize / sparse2Tensor)
  h2gcn/models/__init__.py   : parse_network_setup
  h2gcn/models/_layers.py    : SparseDense, GCNLayer, ConcatLayer
  h2gcn/models/H2GCN.py      : H2GCN.__init__, H2GCN.call (with saveActivations to capture every layer output)

What is shimmed (third-party, absent from the image): `tensorflow` -> tests/golden/tf_shim.py (numpy),
`scipy.sparse.linalg.eigen.arpack` (module path removed in scipy>=1.8; only `eigsh` is imported from it and the
hot path never calls it) and `np.bool` (alias removed in numpy>=1.24).

Outputs (committed): one .npz per graph with raw inputs, the reference's preprocessed tensors (bit-exact
targets for indices / fp32 values) and, per network_setup, the weights plus sampled rows + fp64 checksums of
every layer activation.  `digests.json` holds sha256 digests for cases too big to commit (Pubmed).
"""
import hashlib
import json
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, HERE)

# --- shims for removed third-party names (see module docstring) -----------------------------------------
np.bool = bool
_eig = types.ModuleType("scipy.sparse.linalg.eigen")
_arp = types.ModuleType("scipy.sparse.linalg.eigen.arpack")
from scipy.sparse.linalg import eigsh  # noqa: E402

_arp.eigsh = eigsh
sys.modules["scipy.sparse.linalg.eigen"] = _eig
sys.modules["scipy.sparse.linalg.eigen.arpack"] = _arp
import tf_shim  # noqa: E402

tf = tf_shim.install()
sys.path.insert(0, os.path.join(REF, "h2gcn"))
import datasets._dataset as ref_ds  # noqa: E402  (reference code)
import models as ref_models  # noqa: E402  (reference code)
import models.H2GCN as ref_h2gcn  # noqa: E402  (reference code)

SETUPS = {
    "h2gcn2": "M64-R-T1-G-V-T2-G-V-C1-C2-D0.5-MO",   # default, H2GCN.py:11-13
    "h2gcn1": "M64-R-T1-G-V-C1-D0.5-MO",             # README.md:127
    "h2gcn2_norelu": "M64-T1-G-V-T2-G-V-C1-C2-MO",   # configs/real-cora_full/h2gcn.json
    "h2gcn1_norelu_nodrop": "M64-T1-G-V-C1-MO",
    "mlp": "M64-R-D0.5-MO",                          # configs/syn-cora/mlp.json
    "h2gcn2_hop2only": "M16-R-T1-G1-V-T2-G0_1-V-C1-C2-MO",  # G<i_j> hop subsets, models/__init__.py:86-93
}
ROW_SAMPLE = 41  # keep every 41st row of each activation


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_dataset(adj, features, labels_onehot):
    """Build a reference PlanetoidData around in-memory matrices (bypassing only the file loader)."""
    d = ref_ds.PlanetoidData.__new__(ref_ds.PlanetoidData)
    d._sparse_data = dict()
    d._dense_data = dict()
    n = adj.shape[0]
    d._sparse_data["sparse_adj"] = adj
    d._sparse_data["features"] = features
    m = np.zeros(n, dtype=bool)
    d._dense_data["y_all"] = labels_onehot
    for k in ("train_mask", "val_mask", "test_mask", "wild_mask"):
        d._dense_data[k] = m
    for k in ("y_train", "y_val", "y_test", "y_wild"):
        d._dense_data[k] = np.zeros_like(labels_onehot)
    return d


def run_reference(dataset, hops_spec, setups, normalize_features=True, seed=1234):
    """Mirror of H2GCN.preprocessing_data (H2GCN.py:46-54) + model forward, all reference code."""
    out = {}
    raw_adj = sp.csr_matrix(dataset.sparse_adj)
    raw_feat = sp.csr_matrix(dataset.features)
    out["adj_indptr"] = raw_adj.indptr.astype(np.int32)
    out["adj_indices"] = raw_adj.indices.astype(np.int32)
    out["adj_data"] = raw_adj.data.astype(np.float32)
    out["feat_indptr"] = raw_feat.indptr.astype(np.int32)
    out["feat_indices"] = raw_feat.indices.astype(np.int32)
    out["feat_data"] = raw_feat.data.astype(np.float32)
    out["feat_shape"] = np.array(raw_feat.shape, dtype=np.int64)
    out["num_labels"] = np.array(dataset.num_labels, dtype=np.int64)

    if normalize_features:
        with np.errstate(divide="ignore"):
            dataset.row_normalize_features()
    dataset.adj_remove_eye()
    with np.errstate(divide="ignore"):
        tensors = vars(dataset.getTensors(getDenseAdj=False, getAdjNormHops=hops_spec))
    feats = tensors["features"]
    out["featn_rows"] = np.asarray(feats.indices)[:, 0].astype(np.int32)
    out["featn_cols"] = np.asarray(feats.indices)[:, 1].astype(np.int32)
    out["featn_vals"] = np.asarray(feats.values).astype(np.float32)
    adj_t = tensors["adj"]
    out["adjre_rows"] = np.asarray(adj_t.indices)[:, 0].astype(np.int32)
    out["adjre_cols"] = np.asarray(adj_t.indices)[:, 1].astype(np.int32)
    out["hops_spec"] = np.array(hops_spec)
    for h, t in enumerate(tensors["adj_hops"]):
        out[f"hop{h}_rows"] = np.asarray(t.indices)[:, 0].astype(np.int32)
        out[f"hop{h}_cols"] = np.asarray(t.indices)[:, 1].astype(np.int32)
        out[f"hop{h}_vals"] = np.asarray(t.values).astype(np.float32)

    for sname, setup in setups.items():
        tf_shim.seed(seed)
        layer_setups = ref_models.parse_network_setup(setup, dataset.num_labels, _dense_units=64,
                                                      _dropout_rate=0.5, parse_preprocessing=True)
        model = ref_h2gcn.H2GCN(layer_setups, l2_regularize_weight=5e-4)
        acts = {}
        logits = model(tensors["adj"], tensors["features"], tensors["adj_hops"], training=False,
                       saveActivations=acts)
        out[f"{sname}/setup"] = np.array(setup)
        wi = 0
        for layer in model.layer_objs:
            for (_, w) in getattr(layer, "weights", []):
                w = np.asarray(w)
                key = "weights/" + sha(w)[:12]   # identical tensors (same seed/shape) are stored once
                out[key] = w
                out[f"{sname}/W{wi}"] = np.array(key)
                wi += 1
        names = []
        for k, v in acts.items():
            if not k.startswith("activations/"):
                continue
            name = k.split("/", 1)[1]
            names.append(name)
            v = np.asarray(v, dtype=np.float32)
            v2 = v.reshape(v.shape[0], -1)
            out[f"{sname}/act/{name}/shape"] = np.array(v.shape, dtype=np.int64)
            out[f"{sname}/act/{name}/rows"] = v2[::ROW_SAMPLE].copy()
            out[f"{sname}/act/{name}/sum"] = np.array([v2.astype(np.float64).sum(), np.abs(v2.astype(np.float64)).sum()])
        out[f"{sname}/act_names"] = np.array(names)
        lg = np.asarray(logits, dtype=np.float32)
        out[f"{sname}/logits_rows"] = lg[::ROW_SAMPLE].copy()
        out[f"{sname}/logits_sum"] = np.array([lg.astype(np.float64).sum(), np.abs(lg.astype(np.float64)).sum()])
        out[f"{sname}/logits_sha"] = np.array(sha(lg))
    return out, tensors


def tiny_graphs():
    """Hand-checkable graphs from SURVEY.md §4: path, star, triangle+tail, isolated vertex, self loops."""
    def sym(n, edges, loops=()):
        a = sp.lil_matrix((n, n), dtype=np.float32)
        for i, j in edges:
            a[i, j] = 1
            a[j, i] = 1
        for i in loops:
            a[i, i] = 1
        return a.tocsr()

    g = {
        "path4": sym(4, [(0, 1), (1, 2), (2, 3)]),
        "star5": sym(5, [(0, 1), (0, 2), (0, 3), (0, 4)]),
        "tri_tail": sym(5, [(0, 1), (1, 2), (0, 2), (2, 3), (3, 4)]),
        "isolated": sym(6, [(0, 1), (1, 2), (3, 4)]),           # vertex 5 isolated; {3,4} has no 2-hop
        "selfloops": sym(5, [(0, 1), (1, 2), (2, 3), (3, 4)], loops=(0, 2)),
    }
    rng = np.random.default_rng(7)
    n, m = 40, 90
    e = rng.integers(0, n, size=(m, 2))
    g["rand40"] = sym(n, [(int(a), int(b)) for a, b in e if a != b], loops=(3,))
    return g


def parse_goldens():
    """Reference parse_network_setup output for every H2GCN/MLP string in the reference's experiment configs."""
    import glob
    import shlex
    strings = set(SETUPS.values())
    for fn in sorted(glob.glob(os.path.join(REF, "experiments/h2gcn/configs/*/h2gcn.json")) +
                     glob.glob(os.path.join(REF, "experiments/h2gcn/configs/*/mlp.json"))):
        for line in json.load(open(fn))["model_args"]:
            toks = shlex.split(line)
            if "--network_setup" in toks:
                strings.add(toks[toks.index("--network_setup") + 1])
    strings |= {"F32-R-D-FO", "M-R-T1-G-V-C1-D-MO", "I-T0-G-V-C0-MO", "M8-E-R-T1-G0-V-L-C1-MO",
                "M8-T1-S1_0_4-MO", "M8-T1-S_2-MO", "M8-Xfoo_bar-MO", "[M8]-[lambda x: x]-MO"}

    def ser(v):
        if isinstance(v, set):
            return {"__set__": sorted(v)}
        if isinstance(v, slice):
            return {"__slice__": [v.start, v.stop, v.step]}
        return v
    res = {}
    for st in sorted(strings):
        conf = ref_models.parse_network_setup(st, 7, _dense_units=64, _dropout_rate=0.5, parse_preprocessing=True)
        res[st] = [[t, {k: ser(v) for k, v in c.items()}] for t, c in conf]
    try:
        ref_models.parse_network_setup("M64-Q", 7)
    except ValueError as e:
        res["__error__M64-Q"] = type(e).__name__
    with open(os.path.join(HERE, "parse_network_setup.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)


def main():
    digests = {}
    parse_goldens()
    # ---- tiny graphs ------------------------------------------------------------------------------------
    rng = np.random.default_rng(11)
    for name, adj in tiny_graphs().items():
        n = adj.shape[0]
        feat = sp.random(n, 12, density=0.5, random_state=np.random.RandomState(5), dtype=np.float32).tolil()
        feat[0, :] = 0  # an all-zero feature row exercises the inf->0 mask of row_normalize_features
        lab = np.eye(3, dtype=np.float64)[rng.integers(0, 3, size=n)]
        out, _ = run_reference(make_dataset(adj, feat, lab), ["1", "2"],
                               {k: SETUPS[k] for k in ("h2gcn2", "h2gcn1", "h2gcn2_hop2only")})
        np.savez_compressed(os.path.join(HERE, f"tiny_{name}.npz"), **out)
        if name in ("tri_tail", "rand40"):
            out, _ = run_reference(make_dataset(adj, feat, lab), ["0", "1,2"], {"h2gcn1": SETUPS["h2gcn1"]})
            np.savez_compressed(os.path.join(HERE, f"tiny_{name}_merged.npz"), **out)

    # ---- Planetoid fixtures that ship inside the reference tree -----------------------------------------
    data_dir = os.path.join(REF, "baselines/gcn/gcn/data")
    for ds in ("cora", "citeseer", "pubmed"):
        dataset = ref_ds.PlanetoidData("ind." + ds, data_dir, val_size=500)
        setups = SETUPS if ds == "cora" else {"h2gcn2": SETUPS["h2gcn2"]}
        out, tensors = run_reference(dataset, ["1", "2"], setups)
        d = {k: sha(out[k]) for k in sorted(out) if out[k].dtype.kind in "if" and out[k].size > 8}
        d["_sizes"] = {"N": int(out["feat_shape"][0]), "nnz1": int(out["hop0_rows"].size),
                       "nnz2": int(out["hop1_rows"].size)}
        digests[ds] = d
        if ds != "pubmed":
            np.savez_compressed(os.path.join(HERE, f"planetoid_{ds}.npz"), **out)
        print(ds, d["_sizes"])
    with open(os.path.join(HERE, "digests.json"), "w") as f:
        json.dump(digests, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
