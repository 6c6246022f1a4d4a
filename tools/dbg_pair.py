"""Wait-time breakdown of bm_pair_kernel (library built with H2_BM_TRACE=1): per CTA clock64 accumulators."""
import os, sys, ctypes, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from h2gcn_b200.parallel import ShardedGraph
from h2gcn_b200.ops import HopPlan
from h2gcn_b200.utils import synth
from h2gcn_b200 import _cabi
dev = torch.device('cuda:0')
splits = sys.argv[1] if len(sys.argv) > 1 else "i8x3"
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
g = ShardedGraph(synth.uniform_graph(10000, 200000, seed=0), 0, 1, dev, splits=splits)
x = torch.from_numpy(synth.features(10000, d, 0)).to(dev)
p2 = HopPlan([g.hops[1]], mode="tensor", splits=splits)
y1 = torch.empty(10000, d, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    p2.run(x, y1, [0])
torch.cuda.synchronize()
lib = ctypes.CDLL(_cabi.SO_PATH)
buf = np.zeros(148 * 16, dtype=np.int64)
lib.h2_debug_read_pair(buf.ctypes.data_as(ctypes.c_void_p))
b = buf.reshape(148, 16)
dur = b[:, 1] - b[:, 0]
names = ["MMA waits full_a", "MMA waits acc_empty", "prod w0 waits full_b", "prod w0 waits empty_a", "prod w0 expand+store",
         "TMA waits empty_b", "epilogue w0", "recv spin", "units", "segments"]
print("splits", splits, "d", d, "CTA duration cycles: min %d median %d max %d" % (dur.min(), np.median(dur), dur.max()))
lead = b[0::2]
for k, nm in enumerate(names):
    col = b[:, 2 + k]
    src = lead if k in (0, 1, 8, 9) else b
    c = src[:, 2 + k]
    print("%-24s median %8d  min %8d  max %8d" % (nm, np.median(c), c.min(), c.max()))
for c in (0, 1, 2, 3, 72, 73, 146, 147):
    print("cta", c, "dur", int(dur[c]), [int(v) for v in b[c, 2:12]])

u = np.zeros(8 * 96, dtype=np.int64)
lib.h2_debug_read_pair_units(u.ctypes.data_as(ctypes.c_void_p))
u = u.reshape(8, 96) - b[0, 0]
names = ["TMA issued", "B landed", "expanded", "A free", "stored", "MMA saw A", "MMA issued"]
print("unit " + " | ".join(n.rjust(10) for n in names))
for k in range(60):
    print("%4d " % k + " | ".join(("%10d" % u[r, k]) if u[r, k] > 0 else " " * 10 for r in range(7)))
