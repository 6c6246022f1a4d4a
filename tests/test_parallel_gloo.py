"""Host-side logic of the row-sharded path with world_size 2 over gloo (CPU): nnz-balanced partition, uneven
all-gather of the round input, and that shard-local rounds (computed here by the oracle as the checker) tile the
single-process result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from h2gcn_b200.parallel import all_gather_counts, all_gather_rows, balanced_row_partition, partition_rows
from h2gcn_b200.utils import synth


def test_balanced_partition_properties():
    rng = np.random.default_rng(0)
    w = rng.integers(0, 1000, size=5000)
    w[17] = 400_000  # a hub row
    for world in (1, 2, 3, 4, 8):
        b = balanced_row_partition(w, world)
        assert b[0] == 0 and b[-1] == len(w) and np.all(np.diff(b) >= 0)
        loads = np.array([w[b[q]:b[q + 1]].sum() for q in range(world)])
        assert loads.sum() == w.sum()
        assert loads.max() <= w.sum() / world + w.max()
    assert list(balanced_row_partition(np.zeros(10, dtype=np.int64), 4)) == [0, 2, 5, 7, 10]
    assert list(balanced_row_partition([], 2)) == [0, 0, 0]


def test_partition_prefers_equal_split_when_balanced():
    w = np.full(8000, 100)
    assert list(partition_rows(w, 8)) == [1000 * q for q in range(9)]
    w[5] = 10 ** 6                                  # a hub row: equal split is off by far -> nnz-balanced split
    b = partition_rows(w, 8)
    assert b[1] < 1000 and b[-1] == 8000
    assert list(partition_rows(np.full(10, 3), 4)) == list(balanced_row_partition(np.full(10, 3), 4))   # n % P != 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cbind
        from oracle import h2gcn_oracle as O
        import scipy.sparse as sp
        n, d = 600, 8
        adj = synth.rmat_graph(n, 5000, seed=4)
        x = synth.features(n, d, 0)
        # pass 1 on an equal-rows split, gathered (what ShardedGraph does with h2_hop2_count)
        eq = np.array([(n * q) // world for q in range(world + 1)])
        rp2, col2 = cbind.hop2_csr(adj.indptr, adj.indices)
        deg2_all = np.diff(rp2)
        mine = torch.from_numpy(deg2_all[eq[rank]:eq[rank + 1]].copy())
        deg2 = all_gather_counts(mine, eq).numpy()
        assert np.array_equal(deg2, deg2_all)
        deg1 = np.diff(adj.indptr)
        bounds = balanced_row_partition(deg1 + deg2, world)
        b, e = int(bounds[rank]), int(bounds[rank + 1])
        # the exchange step: uneven all-gather of the input rows
        x_full = torch.empty(n, d)
        all_gather_rows(torch.from_numpy(x[b:e].copy()), x_full, bounds)
        assert np.array_equal(x_full.numpy(), x)
        # the same exchange with bf16 feature rows (BASELINE config 5): a pure copy, bit-identical on every rank
        xb = torch.from_numpy(x).to(torch.bfloat16)
        xb_full = torch.empty(n, d, dtype=torch.bfloat16)
        all_gather_rows(xb[b:e].contiguous(), xb_full, bounds)
        assert torch.equal(xb_full, xb)
        # shard-local round == the same rows of the single-process round (bit-exact: per-row order is P-independent)
        p2 = sp.csr_matrix((np.ones(len(col2)), col2, rp2), shape=(n, n))
        a1, a2 = O.sym_normalize(adj)[0], O.sym_normalize(p2)[0]
        full = cbind.fused_round(a1.indptr, a1.indices, a1.data, a2.indptr, a2.indices, a2.data, x)
        s1, s2 = a1[b:e], a2[b:e]
        part = cbind.fused_round(s1.indptr, s1.indices, s1.data, s2.indptr, s2.indices, s2.data, x_full.numpy())
        assert np.array_equal(part, full[b:e])
        ret[rank] = (b, e, int(s1.nnz + s2.nnz))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    (b0, e0, w0), (b1, e1, w1) = ret[0], ret[1]
    assert b0 == 0 and e0 == b1 and e1 == 600
    assert abs(w0 - w1) <= 0.25 * (w0 + w1), "nnz-balanced, not row-balanced"
