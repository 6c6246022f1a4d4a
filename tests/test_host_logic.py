"""Host-side mirror of the reference interface: model construction from the DSL, index sets, tags, and the fused
program's column layout — all on CPU (no kernel is launched)."""
import numpy as np
import pytest
import torch

from h2gcn_b200.models import Layer, parse_network_setup
from h2gcn_b200.models import _layers as L
from h2gcn_b200.models.H2GCN import H2GCN, _FusedProgram


def test_h2gcn2_layer_objects_match_reference_structure():
    m = H2GCN(parse_network_setup("M64-R-T1-G-V-T2-G-V-C1-C2-D0.5-MO", 7))
    kinds = [type(o).__name__ for o in m.layer_objs]
    assert kinds == ["SparseDense", "ReLU", "GCNLayer", "Flatten", "GCNLayer", "Flatten", "ConcatLayer", "ConcatLayer",
                     "Dropout", "Dense"]
    assert [o.name for o in m.layer_objs] == ["sparse_dense", "re_lu", "gcn_layer", "flatten", "gcn_layer_1",
                                               "flatten_1", "concat_layer", "concat_layer_1", "dropout", "dense"]
    assert m.graph_hops_inds == [2, 4] and m.concat_inds == [6, 7] and m.dropout_inds == [8]
    assert m.output_ind == 9 and m.tagsDict == {1: "1", 3: "2"} and m.embedding_ind is None
    assert L.GCNLayer.SIGNATURE == ["adjhops", "inputs"]


def test_embedding_and_supervised_modifiers():
    m = H2GCN(parse_network_setup("M8-E-R-T1-G0-V-L-C1-MO", 3))
    assert m.embedding_ind == 0 and m.supervised_inds == [3] and m.layer_objs[2].hops == {0}


def test_unknown_layer_type_raises_value_error():
    with pytest.raises(ValueError):
        H2GCN([("Q", {})])
    with pytest.raises(ValueError):
        H2GCN(parse_network_setup("M8-Xfoo_bar-MO", 3))   # experimental layers are not part of the path


class _FakeHop:
    pass


@pytest.mark.parametrize("setup,width,offsets", [
    ("M64-R-T1-G-V-T2-G-V-C1-C2-D0.5-MO", 448, {0: 256, 1: 320, 2: 0}),       # [r2 | r0 | r1]
    ("M64-R-T1-G-V-C1-D0.5-MO", 192, {0: 128, 1: 0}),                          # [r1 | r0]
    ("M16-R-T1-G1-V-T2-G0_1-V-C1-C2-MO", 64, {0: 32, 1: 48, 2: 0}),
    ("M64-R-D0.5-MO", 64, {0: 0}),                                             # MLP
])
def test_fused_program_layout(setup, width, offsets):
    m = H2GCN(parse_network_setup(setup, 7))
    prog = _FusedProgram(m, [_FakeHop(), _FakeHop()], feat_dim=100, n_rows=10, device="cpu")
    assert prog.ok and prog.final_width == width and prog.off == offsets
    assert prog.buf.shape == (10, width)


def test_fused_program_declines_what_it_cannot_lay_out():
    m = H2GCN(parse_network_setup("M64-T0-R-T1-G-V-C0_1-MO", 7))   # pre-ReLU tensor is tagged: cannot fold the ReLU
    assert not _FusedProgram(m, [_FakeHop(), _FakeHop()], 100, 10, "cpu").ok
    m = H2GCN(parse_network_setup("I-T0-G-V-C0-MO", 7))            # dense input path: interpreter only
    assert not _FusedProgram(m, [_FakeHop(), _FakeHop()], 100, 10, "cpu").ok


def test_planetoid_graph_dict_to_adjacency():
    from h2gcn_b200.datasets._dataset import PlanetoidData
    g = {0: [1, 2], 1: [0], 2: [0, 2], 3: []}          # duplicate direction + a self loop + an isolated vertex
    a = PlanetoidData.graphDict2Adj(g).toarray()
    assert (a == np.array([[0, 1, 1, 0], [1, 0, 0, 0], [1, 0, 1, 0], [0, 0, 0, 0]])).all()


def test_glorot_uniform_bounds():
    w = L.glorot_uniform(100, 50, "cpu", torch.Generator().manual_seed(0))
    lim = (6.0 / 150) ** 0.5
    assert w.shape == (100, 50) and float(w.abs().max()) <= lim and float(w.std()) > 0.3 * lim
