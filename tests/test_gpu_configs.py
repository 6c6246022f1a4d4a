"""BASELINE.json `configs` as parity cases (SURVEY.md §8d "Concrete synthetic inputs"): the bench line is configs' north-star
point; the others are checked here against the oracle through the public API.

  cfg0  syn-cora proxy: preferential attachment N=1490, m=2, F=1433 sparse features, full H2GCN-2 forward
  cfg1  Cora (planetoid fixture, golden): covered in test_gpu_parity.py::test_forward_matches_reference_activations
  cfg2  syn-products proxy: preferential attachment N=10 000, m=6, d=100 dense features, --no_feature_normalize
  star  uniform N=10 000, |E|=200 000, d=128: test_gpu_parity.py::test_linearity_and_row_scaling_at_full_size
  skew  RMAT-skewed N=10 000, 200 000 edge draws, d=128 (stress row: hub rows, zero-degree rows)
"""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _host_hops(t):
    return [(h.rowptr.cpu().numpy(), h.col.cpu().numpy(), h.values.cpu().numpy()) for h in t.adj_hops]


def _coo(t_hop):
    idx = t_hop.indices.cpu().numpy()
    return idx[:, 0], idx[:, 1], t_hop.values.cpu().numpy()


def test_cfg0_syn_cora_proxy_full_forward(dev):
    from h2gcn_b200.datasets._dataset import GraphData
    from h2gcn_b200.models import parse_network_setup
    from h2gcn_b200.models.H2GCN import H2GCN
    from h2gcn_b200.utils import synth
    from oracle import cbind
    from oracle import h2gcn_oracle as O
    n, F, C = 1490, 1433, 5
    adj = synth.preferential_attachment(n, 2, seed=0)
    feats = sp.random(n, F, density=0.0127, random_state=np.random.RandomState(0), dtype=np.float32)
    feats.data[:] = 1.0                                   # Cora-like binary bag of words
    data = GraphData(adj, feats.tolil(), np.eye(C)[np.arange(n) % C], device=dev)
    with np.errstate(divide="ignore"):
        data.row_normalize_features()
    data.adj_remove_eye()
    t = data.getTensors(getAdjNormHops=["1", "2"])
    rp2, col2 = cbind.hop2_csr(adj.indptr, adj.indices)
    assert np.array_equal(t.adj_hops[1].rowptr.cpu().numpy(), rp2) and np.array_equal(t.adj_hops[1].col.cpu().numpy(), col2)
    setup = "M64-R-T1-G-V-T2-G-V-C1-C2-D0.5-MO"
    model = H2GCN(parse_network_setup(setup, C, _dense_units=64, _dropout_rate=0.5))
    logits = model(t.adj, t.features, t.adj_hops, training=False)
    weights = [w.cpu().numpy() for w in model.trainable_variables]
    ref = O.forward(parse_network_setup(setup, C, _dense_units=64, _dropout_rate=0.5), weights, _coo(t.features), n,
                    [_coo(h) for h in t.adj_hops])
    assert logits.shape == (n, C) and util.rel_err(logits.cpu().numpy(), ref) <= TOL


@pytest.mark.parametrize("name,gen,args,d", [("cfg2 syn-products proxy", "preferential_attachment", (10000, 6), 100),
                                             ("skew RMAT 10k/200k", "rmat_graph", (10000, 200000), 128)])
@pytest.mark.parametrize("mode", ["auto", "csr"])
def test_fused_round_on_config_graphs(dev, name, gen, args, d, mode):
    from h2gcn_b200.datasets._dataset import GraphData
    from h2gcn_b200.ops import HopPlan
    from h2gcn_b200.utils import synth
    from oracle import cbind
    adj = getattr(synth, gen)(*args)
    n = adj.shape[0]
    data = GraphData(adj, sp.identity(n, dtype=np.float32, format="csr"), device=dev)
    t = data.getTensors(getAdjNormHops=["1", "2"])
    rp2, col2 = cbind.hop2_csr(adj.indptr, adj.indices)
    assert np.array_equal(t.adj_hops[1].col.cpu().numpy(), col2), "2-hop pattern bit-exact"
    deg = np.diff(adj.indptr)
    assert np.array_equal(t.adj_hops[0].dinv.cpu().numpy() == 0, deg == 0), "zero-degree mask"
    x = synth.features(n, d, 7)
    y = torch.full((n, 2 * d), float("nan"), device=dev)
    plan = HopPlan(t.adj_hops, mode=mode)
    plan.run(torch.from_numpy(x).to(dev), y, [0, d])
    (rp1, c1, v1), (rpb, c2, v2) = _host_hops(t)
    ref = cbind.fused_round(rp1, c1, v1, rpb, c2, v2, x)
    assert util.rel_err(y.cpu().numpy(), ref) <= TOL, (name, mode, plan.kernel_name)
    assert (y.cpu().numpy()[deg == 0] == 0).all()
