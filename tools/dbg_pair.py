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

if os.environ.get("H2_TRACE_DUMP"):
    np.savetxt(os.environ["H2_TRACE_DUMP"], np.concatenate([dur[:, None], b[:, 2:12]], axis=1), fmt="%d", delimiter=",",
               header="dur,mma_wait_full_a,mma_wait_acc_empty,prod_wait_full_b,prod_wait_empty_a,prod_expand_store,tma_wait_empty_b,epilogue_w0,recv_spin,units,segments")
# per pair: duration (max of its two CTAs), units, segments -> least-squares  dur ~ a + b * units + c * (segments - 1)
pd = np.maximum(dur[0::2], dur[1::2]).astype(np.float64)
pu, ps = lead[:, 10].astype(np.float64), lead[:, 11].astype(np.float64)
ok = pu > 0
A = np.stack([np.ones(ok.sum()), pu[ok], ps[ok] - 1], axis=1)
coef, *_ = np.linalg.lstsq(A, pd[ok], rcond=None)
print("fit: dur = %.0f + %.1f * units + %.0f * (segments - 1);  residual rms %.0f" % (coef[0], coef[1], coef[2], np.sqrt(np.mean((A @ coef - pd[ok]) ** 2))))
for sgs in (1, 2, 3):
    m = ok & (ps == sgs)
    if m.any(): print("pairs with %d segment(s): n %d  units mean %.1f  dur mean %.0f min %.0f max %.0f  epilogue(w0 of leader) mean %.0f" % (sgs, m.sum(), pu[m].mean(), pd[m].mean(), pd[m].min(), pd[m].max(), lead[m, 8].mean()))
u = np.zeros(8 * 96, dtype=np.int64)
lib.h2_debug_read_pair_units(u.ctypes.data_as(ctypes.c_void_p))
u = u.reshape(8, 96) - b[0, 0]
names = ["TMA issued", "B landed", "expanded", "A free", "stored", "MMA saw A", "MMA issued"]
print("unit " + " | ".join(n.rjust(10) for n in names))
for k in range(int(os.environ.get("H2_TRACE_UNITS", "6"))):
    print("%4d " % k + " | ".join(("%10d" % u[r, k]) if u[r, k] > 0 else " " * 10 for r in range(7)))
