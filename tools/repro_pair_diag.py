"""Where does bm_pair_kernel differ from the CSR path on a wide row shard?  Per (256-row tile, 64-feature block) max error."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200 import ops
from h2gcn_b200.utils import synth
dev = torch.device("cuda:0")
n, rows, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
adj = synth.chung_lu_graph_device(n, 16 * n, gamma=2.5, seed=2, device=dev)
g = ShardedGraph(adj, 0, 1, dev, factored=True, explicit_vals=False, mode="csr")
h = g.hops[1]
e = int(h.rowptr[rows].item())
sl = [ops.SparseTensor(h.rowptr[:rows + 1].contiguous(), h.col[:e].contiguous(), None, (rows, n), row_begin=0, dinv=h.dinv)]
x = torch.randn(n, d, device=dev)
y0 = torch.empty(rows, d, device=dev)
ops.HopPlan(sl, factored=True, mode="csr").run(x, y0, [0])
pt = ops.HopPlan(sl, factored=True, mode="tensor")
scale = y0.abs().max().item()
for rep in range(reps):
    y1 = torch.full((rows, d), float("nan"), device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pt.run(x, y1, [0])
    e1.record()
    torch.cuda.synchronize()
    if rep == reps - 1: print("tensor round %.3f ms" % e0.elapsed_time(e1), flush=True)
    err = (y0 - y1).abs() / scale
    err = torch.nan_to_num(err, nan=9.0)
    blk = err.view(rows // 256, 256, d // 64, 64).amax(dim=(1, 3)).cpu().numpy()
    bad = np.argwhere(blk > 1e-5)
    print("rep", rep, "max rel err %.3e" % err.max().item(), "bad (tile, block64):", [tuple(b) for b in bad[:24]], "n_bad", len(bad), flush=True)
    if len(bad):
        t, b = bad[0]
        sub = err[t * 256:(t + 1) * 256, b * 64:(b + 1) * 64]
        rws = (sub.amax(dim=1) > 1e-5).nonzero().flatten().cpu().numpy()
        cls = (sub.amax(dim=0) > 1e-5).nonzero().flatten().cpu().numpy()
        print("   first bad block: rows", rws[:8], "...", len(rws), "cols", cls[:8], "...", len(cls), "ratio y1/y0 sample", (y1[t * 256 + rws[0], b * 64 + cls[0]] / y0[t * 256 + rws[0], b * 64 + cls[0]]).item(), flush=True)
