"""Back-to-back (pipelined, 2 caller streams, rotating replicas) rate of: the full round, the tensor hop alone, the CSR hop alone."""
import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200.ops import HopPlan
from h2gcn_b200.utils import synth
dev = torch.device('cuda:0')
n, d, R, K = 10000, 128, 7, 400
adj = synth.uniform_graph(n, 200000, seed=0)
gs = [ShardedGraph(adj, 0, 1, dev) for _ in range(R)]
xs = [torch.from_numpy(synth.features(n, d, 0)).to(dev) for _ in range(R)]
ys = [torch.empty(n, 2 * d, device=dev) for _ in range(R)]
main = torch.cuda.current_stream()
def rate(plans, offs, lanes_n):
    lanes = [torch.cuda.Stream(device=dev) for _ in range(lanes_n)]
    def run(k):
        with torch.cuda.stream(lanes[k % lanes_n]):
            plans[k % R].run(xs[k % R], ys[k % R], offs)
    for k in range(2 * R): run(k)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for l in lanes: l.wait_stream(main)
    for k in range(K): run(k)
    for l in lanes: main.wait_stream(l)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / K
full = [g.plan for g in gs]
p2 = [HopPlan([g.hops[1]], mode="tensor") for g in gs]
p1 = [HopPlan([g.hops[0]], mode="csr") for g in gs]
for lanes_n in (1, 2):
    print("lanes", lanes_n, "full %.1f us  tensor-hop only %.1f us  csr-hop only %.1f us" % (rate(full, [0, d], lanes_n), rate(p2, [d], lanes_n), rate(p1, [0], lanes_n)))
