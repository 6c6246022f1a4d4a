"""Pins the CPU oracle (oracle/h2gcn_oracle.py + the C restatement) to golden vectors produced by EXECUTING the
reference's own Python (tests/golden/make_golden.py).  Integer / fp32-adjacency outputs bit-exact; activations 1e-6."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import cbind
from oracle import h2gcn_oracle as O
from h2gcn_b200.models import parse_network_setup
from tests import util


@pytest.mark.parametrize("name", util.all_golden_names())
def test_precompute_bit_exact(name):
    z = util.load_golden(name)
    adj = O.remove_eye(util.raw_adj(z))
    r, c, _ = O.to_coo_sorted(adj)
    assert np.array_equal(r, z["adjre_rows"]) and np.array_equal(c, z["adjre_cols"])
    hops = O.adj_norm_hops(adj, [str(s) for s in z["hops_spec"]])
    for (gr, gc, gv), (orow, ocol, oval) in zip(util.golden_hops(z), hops):
        assert np.array_equal(orow, gr) and np.array_equal(ocol, gc)
        assert np.array_equal(oval.view(np.uint32), gv.view(np.uint32)), "fp32 adjacency values must be bit-exact"


@pytest.mark.parametrize("name", ["tiny_tri_tail_merged", "tiny_rand40_merged"])
def test_merged_hops(name):
    z = util.load_golden(name)
    adj = O.remove_eye(util.raw_adj(z))
    hops = O.adj_norm_hops(adj, [str(s) for s in z["hops_spec"]])
    assert len(hops) == len(util.golden_hops(z)) == 2
    for (gr, gc, gv), (orow, ocol, oval) in zip(util.golden_hops(z), hops):
        assert np.array_equal(orow, gr) and np.array_equal(ocol, gc) and np.array_equal(oval, gv)


@pytest.mark.parametrize("name", util.all_golden_names())
def test_feature_normalisation(name):
    z = util.load_golden(name)
    with np.errstate(divide="ignore"):
        r, c, v = O.to_coo_sorted(O.row_normalize_features(util.raw_feat(z).tolil()))  # LIL like the loader (:241)
    assert np.array_equal(r, z["featn_rows"]) and np.array_equal(c, z["featn_cols"])
    assert np.array_equal(v, z["featn_vals"])


@pytest.mark.parametrize("name", util.all_golden_names())
def test_forward_activations(name):
    z = util.load_golden(name)
    n = int(z["feat_shape"][0])
    feat = (z["featn_rows"].astype(np.int64), z["featn_cols"].astype(np.int64), z["featn_vals"])
    hops = util.golden_hops(z)
    for setup in util.setups_in(z):
        conf = parse_network_setup(str(z[f"{setup}/setup"]), int(z["num_labels"]), _dense_units=64, _dropout_rate=0.5)
        logits, acts = O.forward(conf, util.weights_of(z, setup), feat, n, hops, return_activations=True)
        names = [str(s) for s in z[f"{setup}/act_names"]]
        assert len(names) == len(acts)
        for nm, a in zip(names, acts):
            a2 = np.asarray(a, dtype=np.float32).reshape(a.shape[0], -1)
            assert tuple(z[f"{setup}/act/{nm}/shape"]) == tuple(np.asarray(a).shape)
            assert util.rel_err(a2[::41], z[f"{setup}/act/{nm}/rows"]) <= 1e-6
            s = z[f"{setup}/act/{nm}/sum"]
            assert abs(a2.astype(np.float64).sum() - s[0]) <= 1e-6 * max(1.0, s[1])
        assert util.rel_err(logits[::41], z[f"{setup}/logits_rows"]) <= 1e-6


def test_pubmed_digest_sizes():
    d = json.load(open(os.path.join(util.GOLDEN, "digests.json")))
    assert d["pubmed"]["_sizes"] == {"N": 19717, "nnz1": 88648, "nnz2": 1075702}
    assert d["cora"]["_sizes"] == {"N": 2708, "nnz1": 10556, "nnz2": 86332}


# ---- the C restatement against the Python oracle ---------------------------------------------------------------
@pytest.mark.parametrize("name", util.all_golden_names())
def test_c_oracle_matches_python(name):
    z = util.load_golden(name)
    n = int(z["feat_shape"][0])
    adj = O.remove_eye(util.raw_adj(z))
    rp2, col2 = cbind.hop2_csr(adj.indptr, adj.indices)
    g = util.golden_hops(z)[1]
    ref_rp, ref_col = util.coo_to_csr(g[0], g[1], n)
    assert np.array_equal(rp2, ref_rp) and np.array_equal(col2, ref_col)
    x = np.random.default_rng(3).standard_normal((n, 8)).astype(np.float32)
    for rows, cols, vals in util.golden_hops(z):
        a = O.spmm_coo(rows, cols, vals, x, n)
        b = cbind.spmm_coo(rows, cols, vals, x, n)
        # np.add.at and the C loop visit the nonzeros in the same order with separate mul and add
        assert np.array_equal(a, b)
    (r1, c1, v1), (r2, c2, v2) = util.golden_hops(z)[:2]
    rp1, cc1 = util.coo_to_csr(r1, c1, n)
    rpp2, cc2 = util.coo_to_csr(r2, c2, n)
    y = cbind.fused_round(rp1, cc1, v1, rpp2, cc2, v2, x)
    assert np.array_equal(y, O.fused_round(util.golden_hops(z)[:2], x))


def test_hand_checked_path4():
    """P4: 0-1-2-3.  Distance-2 pairs are (0,2) and (1,3); degrees 1,2,2,1 -> A1 values 1/sqrt(di dj)."""
    a = sp.csr_matrix(np.array([[0, 1, 0, 0], [1, 0, 1, 0], [0, 1, 0, 1], [0, 0, 1, 0]], dtype=np.float32))
    rings = O.nhood_split(a, 2)
    assert (rings[2].toarray() == np.array([[0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0], [0, 1, 0, 0]])).all()
    (r1, c1, v1), (r2, c2, v2) = O.adj_norm_hops(a)
    assert np.allclose(v1, [2 ** -0.5, 2 ** -0.5, 0.5, 0.5, 2 ** -0.5, 2 ** -0.5])
    assert np.allclose(v2, 1.0) and list(zip(r2, c2)) == [(0, 2), (1, 3), (2, 0), (3, 1)]


def test_early_stop_short_list():
    """A single edge has no distance-2 pairs: nhoodSplit returns [I, P1] and adj_hops[2] raises (reference :571)."""
    a = sp.csr_matrix(np.array([[0, 1], [1, 0]], dtype=np.float32))
    assert len(O.nhood_split(a, 2)) == 2
    with pytest.raises(IndexError):
        O.adj_norm_hops(a, ["1", "2"])
