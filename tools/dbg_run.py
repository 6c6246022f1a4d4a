"""Timeline of bm_mma_kernel (library built with -DH2_BM_TRACE): per CTA clock64 stamps.
slots: 0 start, 1 end, per segment w: 2+6w mma-thread reaches segment, 3+6w acc_empty seen, 4+6w last issue done,
5+6w producer reaches epilogue wait, 6+6w acc_full seen, 7+6w epilogue done."""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200.utils import synth
from h2gcn_b200 import _cabi
dev = torch.device('cuda:0')
adj = synth.uniform_graph(10000, 200000, seed=0)
splits = sys.argv[1] if len(sys.argv) > 1 else "2"
splits = int(splits) if splits.isdigit() else splits
g = ShardedGraph(adj, 0, 1, dev, splits=splits)
x = torch.from_numpy(synth.features(10000, 128, 0)).to(dev); y = torch.empty(10000, 256, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
alone = len(sys.argv) > 2 and sys.argv[2] == "alone"     # only the tensor-core hop (no co-running CSR gather kernel)
if alone:
    from h2gcn_b200.ops import HopPlan
    p2 = HopPlan([g.hops[1]], mode="tensor", splits=splits)
    y1 = torch.empty(10000, 128, device=dev)
for _ in range(3):
    flush.zero_()
    if alone:
        p2.run(x, y1, [0])
    else:
        g.round(x, y, [0, 128])
torch.cuda.synchronize()
lib = ctypes.CDLL(_cabi.SO_PATH)
buf = np.zeros(148 * 32, dtype=np.int64)
lib.h2_debug_read(buf.ctypes.data_as(ctypes.c_void_p))
b = buf.reshape(148, 32)
dur = b[:, 1] - b[:, 0]
print("splits", splits, "CTA duration cycles: min %d median %d max %d" % (dur.min(), np.median(dur), dur.max()))
w = b[:, 26:32]
print("median per CTA: units %d | MMA thread waits full_a %d | producer warp 0 waits full_b %d, empty_a %d, wait::st+arrive %d | "
      "TMA thread waits empty_b %d" % (np.median(w[:, 5]), np.median(w[:, 0]), np.median(w[:, 1]), np.median(w[:, 2]),
                                       np.median(w[:, 3]), np.median(w[:, 4])))
for c in [0, 1, 2, 37, 73, 74, 100, 147]:
    r = b[c]; t0 = r[0]
    segs = []
    for w in range(4):
        s = r[2 + 6 * w: 8 + 6 * w]
        if 8 + 6 * w > 26: break
        if s[0] == 0: break
        segs.append([int(v - t0) for v in s])
    print(c, "end", int(r[1] - t0), segs)

# per-unit timeline of CTA 0 (first 40 units): cycles relative to the CTA start
b2 = np.zeros(8 * 64, dtype=np.int64)
lib.h2_debug_read2(b2.ctypes.data_as(ctypes.c_void_p))
b2 = b2.reshape(8, 64) - b[0, 0]
names = ["TMA: stage free", "prod: B landed", "prod: expanded", "prod: A stage free", "prod: stored", "MMA: A ready", "MMA: issued"]
print("unit " + " | ".join(n.rjust(18) for n in names))
for u in range(40):
    print("%4d " % u + " | ".join(("%18d" % b2[r, u]) if b2[r, u] > 0 else " " * 18 for r in range(7)))
