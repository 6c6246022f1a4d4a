"""bm_pair_kernel on a row shard with > 2^17 columns (accumulator cuts every 2048 units) and two column groups (d = 256)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200 import ops
from h2gcn_b200.utils import synth
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
d = int(sys.argv[3]) if len(sys.argv) > 3 else 256
adj = synth.chung_lu_graph_device(n, 16 * n, gamma=2.5, seed=2, device=dev)
g = ShardedGraph(adj, 0, 1, dev, factored=True, explicit_vals=False, mode="csr")
sl = []
for h in g.hops:
    e = int(h.rowptr[rows].item())
    sl.append(ops.SparseTensor(h.rowptr[:rows + 1].contiguous(), h.col[:e].contiguous(), None, (rows, n), row_begin=0, dinv=h.dinv))
print("n", n, "rows", rows, "nnz", [int(s.rowptr[-1]) for s in sl], flush=True)
for td in (torch.float32, torch.bfloat16):
    x = torch.randn(n, d, device=dev).to(td)
    pc = ops.HopPlan(sl, factored=True, mode="csr")
    pt = ops.HopPlan(sl, factored=True, mode="tensor")
    y0 = torch.empty(rows, 2 * d, device=dev, dtype=td)
    y1 = torch.empty(rows, 2 * d, device=dev, dtype=td)
    pc.run(x, y0, [0, d]); torch.cuda.synchronize(); print("csr ok", flush=True)
    pt.run(x, y1, [0, d]); torch.cuda.synchronize()
    err = (y0.float() - y1.float()).abs().max().item() / y0.float().abs().max().item()
    print(td, pt.kernel_name, "rel diff csr vs tensor %.2e" % err, flush=True)
