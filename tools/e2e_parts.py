"""Where the end-to-end (host-buffer) time goes: raw PCIe copies vs the h2_graph_round_host call."""
import sys, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
dev = torch.device('cuda:0')
n, d = 10000, 128
xh = torch.randn(n, d).pin_memory(); yh = torch.empty(n, 2 * d).pin_memory()
xd = torch.empty(n, d, device=dev); yd = torch.empty(n, 2 * d, device=dev)
def t(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn(); torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / reps
print("H2D 5.12 MB   %.0f us" % t(lambda: xd.copy_(xh, non_blocking=True)))
print("D2H 10.24 MB  %.0f us" % t(lambda: yh.copy_(yd, non_blocking=True)))
print("D2H 2 x strided 5.12 MB  %.0f us" % t(lambda: (yh[:, :d].copy_(yd[:, :d], non_blocking=True), yh[:, d:].copy_(yd[:, d:], non_blocking=True))))
print("sync only %.0f us" % t(lambda: None))
