"""A few fused rounds at the north-star point, for `ncu --metrics gpu__time_duration.sum` launch lists and traces."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200.utils import synth
dev = torch.device('cuda:0')
n, d = 10000, int(sys.argv[2]) if len(sys.argv) > 2 else 128
splits = sys.argv[1] if len(sys.argv) > 1 else "i8x3"
g = ShardedGraph(synth.uniform_graph(n, 200000, seed=0), 0, 1, dev, splits=splits)
x = torch.from_numpy(synth.features(n, d, 0)).to(dev)
y = torch.empty(n, 2 * d, device=dev)
for _ in range(int(sys.argv[3]) if len(sys.argv) > 3 else 4):
    g.round(x, y, [0, d])
torch.cuda.synchronize()
print("ok", float(y.abs().max()))
