"""Host logic of the CTA-pair kernel's stream-K schedule (h2gcn_b200/csrc/bm_pair.cu: pair_schedule), checked on the CPU
through the library's self-check hook: every (column group, unit) is covered exactly once and in order, no segment leaves
its row tile or exceeds the int32-accumulator cut, split items own consecutive slots in unit order with one arrival
counter, and a designated finisher is always the last segment of its pair."""
import ctypes

import numpy as np
import pytest


@pytest.fixture(scope="module")
def check():
    from h2gcn_b200 import build
    lib = ctypes.CDLL(build.build())
    fn = lib.h2_debug_pair_schedule_check
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_int32, ctypes.POINTER(ctypes.c_int64), ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)]

    def run(tile_units, ng, n_pairs=74):
        tu = np.ascontiguousarray(tile_units, dtype=np.int64)
        stats = np.zeros(8, dtype=np.int64)
        rc = fn(len(tu), tu.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ng, n_pairs, stats.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
        return rc, dict(zip(["segments", "split_items", "slots", "longest", "max_pair", "min_pair", "finishers", "pairs_used"], stats.tolist()))
    return run


@pytest.mark.parametrize("tiles,per,ng", [(40, 157, 1), (40, 157, 2), (32, 4096, 2), (691, 16384, 2), (2048, 16384, 1), (11, 43, 4),
                                          (3, 5, 1), (1, 1, 1), (1, 100000, 1), (200, 1, 2)])
def test_uniform_tiles(check, tiles, per, ng):
    rc, st = check([per] * tiles, ng)
    assert rc == 0, (rc, st)
    assert st["longest"] <= 2048
    total = tiles * per * ng
    assert st["pairs_used"] <= 74 and st["max_pair"] >= -(-total // 74)
    if total >= 74 * 64:     # enough work: no pair carries more than ~1.5x the mean
        assert st["max_pair"] <= 1.5 * total / 74 + 64, st
    if per > 2048:
        assert st["split_items"] == tiles * ng
    assert st["finishers"] <= st["split_items"]


def test_north_star_shape(check):
    """|V| = 10 000, d = 128: 40 row tiles x 157 units on 74 pairs — every item is split, most pairs have 1-2 segments."""
    rc, st = check([157] * 40, 1)
    assert rc == 0
    assert st["pairs_used"] == 74 and st["split_items"] == 40 and 74 <= st["segments"] <= 74 + 40
    assert st["max_pair"] <= 110 and st["finishers"] >= 30


@pytest.mark.parametrize("seed", range(12))
def test_ragged_and_empty_tiles(check, seed):
    rng = np.random.default_rng(seed)
    tiles = int(rng.integers(1, 300))
    tu = rng.integers(0, int(rng.choice([3, 40, 500, 6000])), size=tiles)
    tu[rng.random(tiles) < 0.2] = 0                       # empty row tiles
    for ng in (1, 2, 4):
        rc, st = check(tu, ng, n_pairs=int(rng.choice([1, 7, 74])))
        assert rc == 0, (seed, ng, rc, st)


def test_cost_knob_and_finisher_switch(check, monkeypatch):
    monkeypatch.setenv("H2_PAIR_COSTS", "0,0,0")
    rc, st0 = check([157] * 40, 1)
    assert rc == 0 and st0["max_pair"] <= 86              # pure unit balance: ceil(6280 / 74) = 85
    monkeypatch.setenv("H2_PAIR_COSTS", "10,24,0")
    monkeypatch.setenv("H2_PAIR_NO_FINISHER", "1")
    rc, st1 = check([157] * 40, 1)
    assert rc == 0 and st1["finishers"] == 0
