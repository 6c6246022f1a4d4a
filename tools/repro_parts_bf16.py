"""Row-sharded bf16 round on ONE GPU at a size where the pack kernel walks several items per CTA (second pass from the
gathered copy) and the pair kernel cuts segments every 2048 units: run_parts vs run, bit-identical."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200.utils import synth
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
mode = sys.argv[3] if len(sys.argv) > 3 else "tensor"
adj = synth.chung_lu_graph_device(n, 16 * n, gamma=2.5, seed=2, device=dev)
g = ShardedGraph(adj, 0, 1, dev, factored=True, explicit_vals=False, mode=mode)
print("n", n, "nnz2", g.nnz2_local, "kernel", g.plan.kernel_name, flush=True)
for td in (torch.float32, torch.bfloat16):
    x = torch.randn(n, d, device=dev).to(td)
    y0 = torch.empty(n, 2 * d, device=dev, dtype=td)
    g.plan.run(x, y0, [0, d])
    torch.cuda.synchronize()
    bounds = [0, n // 3 + 5, n]
    parts = [x[bounds[q]:bounds[q + 1]].clone() for q in range(2)]
    xf = torch.full((n, d), float("nan"), device=dev, dtype=td)
    y1 = torch.full((n, 2 * d), float("nan"), device=dev, dtype=td)
    g.plan.run_parts([p.data_ptr() for p in parts], bounds, d, xf, y1, [0, d], d)
    torch.cuda.synchronize()
    print(td, "equal", bool(torch.equal(y0, y1)), "xfull equal", bool(torch.equal(xf, x)), flush=True)
