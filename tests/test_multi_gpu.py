"""Row-sharded round on >= 2 GPUs (skipped on a 1-GPU box): NCCL exchange and the peer-memory exchange fused into the
pack kernel give bit-identical shards, and the shards tile the single-process oracle result.
Run on a multi-GPU box:  gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu -q"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, graph_kind, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from h2gcn_b200.parallel import ShardedGraph
        from h2gcn_b200.utils import synth
        from oracle import cbind
        from oracle import h2gcn_oracle as O
        import scipy.sparse as sp
        n, d = (6000, 128) if graph_kind == "uniform" else (5000, 64)
        adj = synth.uniform_graph(n, 90000, seed=1) if graph_kind == "uniform" else synth.rmat_graph(n, 40000, seed=2)
        x = synth.features(n, d, 3)
        outs = {}
        for exchange in ("nccl", "p2p"):
            g = ShardedGraph(adj, rank, world, dev, exchange=exchange)
            assert g.exchange == exchange
            xl = torch.from_numpy(x[g.row_begin:g.row_end]).to(dev)
            y = torch.full((g.n_local, 2 * d), float("nan"), device=dev)
            for _ in range(2):                      # twice: the barriers must let a second round reuse the buffers
                g.round(xl, y, [0, d])
            torch.cuda.synchronize()
            outs[exchange] = y.cpu().numpy()
        assert np.array_equal(outs["nccl"], outs["p2p"]), "the exchange mechanism must not change a single bit"
        rp2, col2 = cbind.hop2_csr(adj.indptr, adj.indices)
        p2 = sp.csr_matrix((np.ones(len(col2)), col2, rp2), shape=(n, n))
        a1, a2 = O.sym_normalize(adj)[0], O.sym_normalize(p2)[0]
        full = cbind.fused_round(a1.indptr, a1.indices, a1.data, a2.indptr, a2.indices, a2.data, x)
        ref = full[g.row_begin:g.row_end]
        err = float(np.abs(outs["p2p"] - ref).max() / np.abs(full).max())
        assert err <= 1e-4, err
        # bf16 rows (BASELINE config 5): shards, gathered copy and output in bf16; within one bf16 rounding of the exact result
        xb = torch.from_numpy(x).to(dev).to(torch.bfloat16)
        yb = torch.full((g.n_local, 2 * d), float("nan"), device=dev, dtype=torch.bfloat16)
        g.round(xb[g.row_begin:g.row_end].contiguous(), yb, [0, d])
        torch.cuda.synchronize()
        x64 = xb.float().cpu().numpy().astype(np.float64)
        refb = np.concatenate([a1[g.row_begin:g.row_end].astype(np.float64) @ x64, a2[g.row_begin:g.row_end].astype(np.float64) @ x64], axis=1)
        errb = np.abs(yb.float().cpu().numpy() - refb)
        assert (errb <= 2.0 ** -8 * np.abs(refb) + 1e-6 * np.abs(refb).max()).all()
        ret[rank] = (g.row_begin, g.row_end, err, g.plan.kernel_name)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("graph_kind", ["uniform", "rmat"])
def test_sharded_round_two_gpus(graph_kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, graph_kind, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    assert ret[0][1] == ret[1][0]
