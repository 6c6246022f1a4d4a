"""BASELINE config 3 (syn-products proxy: preferential attachment |V| = 10 000, m = 6, d = 100): a few fused rounds, the
target of the ncu capture `profiles/r02_cfg3_round_traffic.csv` (same graph as bench.py's `secondary.cfg3_syn_products_round`)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200.utils import synth
dev = torch.device("cuda:0")
n, d = 10000, 100
g = ShardedGraph(synth.preferential_attachment(n, 6, seed=0), 0, 1, dev)
x = torch.from_numpy(synth.features(n, d, 0)).to(dev)
y = torch.empty(n, 2 * d, device=dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    g.round(x, y, [0, d])
torch.cuda.synchronize()
print("ok", g.plan.kernel_name, "nnz", g.nnz_local)
