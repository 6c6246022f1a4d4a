"""Times the parts of one fused round separately (CUDA events, L2 flushed before every launch sequence)."""
import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200.ops import HopPlan
from h2gcn_b200.utils import synth
dev = torch.device('cuda:0')
n, d = 10000, 128
splits = sys.argv[1] if len(sys.argv) > 1 else "2"
splits = int(splits) if splits.isdigit() else splits
g = ShardedGraph(synth.uniform_graph(n, 200000, seed=0), 0, 1, dev, splits=splits)
x = torch.from_numpy(synth.features(n, d, 0)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=50, do_flush=True):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(reps):
        if do_flush: flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    return np.median(ts), np.min(ts)
y = torch.empty(n, 2 * d, device=dev); y1 = torch.empty(n, d, device=dev)
full = g.plan
p2 = HopPlan([g.hops[1]], mode="tensor", splits=splits); p1 = HopPlan([g.hops[0]], mode="csr")
print("full round (flush)      med %.1f us  min %.1f us" % timeit(lambda: full.run(x, y, [0, d])))
print("hop2 tensor only (flush) med %.1f us  min %.1f us" % timeit(lambda: p2.run(x, y1, [0])))
print("hop1 csr only (flush)    med %.1f us  min %.1f us" % timeit(lambda: p1.run(x, y1, [0])))
print("full round (warm L2)     med %.1f us  min %.1f us" % timeit(lambda: full.run(x, y, [0, d]), do_flush=False))
print("hop2 tensor (warm L2)    med %.1f us  min %.1f us" % timeit(lambda: p2.run(x, y1, [0]), do_flush=False))
print("hop1 csr (warm L2)       med %.1f us  min %.1f us" % timeit(lambda: p1.run(x, y1, [0]), do_flush=False))
print("empty event pair         med %.1f us  min %.1f us" % timeit(lambda: None, do_flush=False))
