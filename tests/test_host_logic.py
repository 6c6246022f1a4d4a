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


# ---- arithmetic codes of the tensor-core path and the numpy model of its int8 operand (no kernel is launched) -----------
def test_splits_codes():
    from h2gcn_b200 import _cabi
    assert _cabi.splits_code(2) == 2 and _cabi.splits_code("3") == 3 and _cabi.splits_code("bf16x2") == 2
    assert _cabi.splits_code("i8x2") == _cabi.H2_SPLITS_I8X2 == 18 and _cabi.splits_code(19) == _cabi.H2_SPLITS_I8X3
    assert _cabi.splits_code(None) == _cabi.splits_code(_cabi.DEFAULT_SPLITS)
    assert set(_cabi.SPLITS_NAME) == {2, 3, 18, 19}
    for bad in (4, "i8x4", "fp8", 1.5):
        with pytest.raises(ValueError):
            _cabi.splits_code(bad)


@pytest.mark.parametrize("pieces,rng_bound", [(2, 32639), (3, 8355711)])
def test_int8_operand_model(pieces, rng_bound):
    """tests/util.py: i8_block_quantize is the model the GPU test compares the int8 kernel with EXACTLY: its own
    invariants — digits range, ROW exponents 0..6 relative to the global maximum, error bound per element."""
    from tests import util
    rng = np.random.default_rng(pieces)
    x = (rng.standard_normal((403, 24)) * np.exp(rng.uniform(-12, 0, size=(403, 1)))).astype(np.float32)
    x[100:104] = 0.0
    deq, step, t = util.i8_block_quantize(x, pieces)
    assert t.min() >= 0 and t.max() <= 6 and len(t) == 403 and (t[100:104] == 0).all()     # one exponent per row
    g = int(np.argmax(np.abs(x).max(axis=1)))
    assert t[g] == 6, "the row holding the global maximum gets the largest exponent"
    q = deq / (step * np.ldexp(1.0, t)[:, None])
    assert np.allclose(q, np.rint(q)) and np.abs(q).max() <= rng_bound
    # |error| <= half a step of the element's row (+ the fp32 rounding of the scaled value)
    bound = 0.5 * step * np.ldexp(1.0, t)[:, None] * (1 + 1e-6) + np.abs(x) * 2.0 ** -23
    assert (np.abs(deq - x) <= bound).all()
    # balanced base-256 digits reproduce q
    qi = np.rint(q).astype(np.int64)
    rest, digits = qi.copy(), []
    for _ in range(pieces):
        dgt = ((rest + 128) & 255) - 128
        digits.append(dgt)
        rest = (rest - dgt) >> 8
    assert (rest == 0).all() and all((dg >= -128).all() and (dg <= 127).all() for dg in digits)
    assert (sum(dg * 256 ** k for k, dg in enumerate(digits)) == qi).all()
    assert (util.i8_block_quantize(np.zeros((8, 4), np.float32), pieces)[0] == 0).all()


def test_bitmap_bit_order_is_a_permutation():
    """bm_bit_pos order 1 (csrc/bitmap_mma.cu): bit(c) = 32*(c/32) + 8*(c%4) + (c%32)/4 — the four columns of operand word
    j = c/4 land 8 bits apart at bit (j % 8) of each byte, so word j = rotate(x, j%8 - t) & (0x01010101 << t)."""
    pos = [32 * (c // 32) + 8 * (c % 4) + (c % 32) // 4 for c in range(64)]
    assert sorted(pos) == list(range(64))
    for c in range(64):
        j, k = c // 4, c % 4
        assert pos[c] // 32 == j // 8 and pos[c] % 32 == 8 * k + j % 8
    for t in range(7):                                   # the expansion identity, on every single-bit row
        for c in range(64):
            x = (1 << pos[c]) >> (32 * (c // 32)) & 0xFFFFFFFF
            j = c // 4
            r = ((j % 8) - t) & 31
            word = ((x >> r) | (x << (32 - r))) & 0xFFFFFFFF & ((0x01010101 << t) & 0xFFFFFFFF)
            assert word == (1 << t) << (8 * (c % 4)), (t, c)
