"""Repeatability of the tensor hop at the north-star point: N reps against the CSR result, per-tile error report."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200 import ops
from h2gcn_b200.utils import synth
dev = torch.device("cuda:0")
d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
g = ShardedGraph(synth.uniform_graph(n, 20 * n, seed=0), 0, 1, dev, mode="csr")
x = torch.from_numpy(synth.features(n, d, 0)).to(dev)
y0 = torch.empty(n, d, device=dev)
ops.HopPlan([g.hops[1]], mode="csr").run(x, y0, [0])
pt = ops.HopPlan([g.hops[1]], mode="tensor")
scale = y0.abs().max().item()
nt = (n + 255) // 256
bad_reps = 0
for rep in range(reps):
    y1 = torch.full((n, d), float("nan"), device=dev)
    pt.run(x, y1, [0])
    torch.cuda.synchronize()
    err = torch.nan_to_num((y0 - y1).abs() / scale, nan=9.0)
    rowerr = err.amax(dim=1).cpu().numpy()
    tiles = sorted({int(r // 256) for r in np.nonzero(rowerr > 1e-5)[0]})
    if tiles:
        bad_reps += 1
        if bad_reps <= 6:
            t = tiles[0]
            rws = np.nonzero(rowerr[t * 256:(t + 1) * 256] > 1e-5)[0]
            print("rep", rep, "max err %.2e" % rowerr.max(), "bad tiles", tiles[:12], "rows in first bad tile: %d..%d (%d)" % (rws.min(), rws.max(), len(rws)), flush=True)
print("d", d, "reps", reps, "bad reps", bad_reps, flush=True)
