#!/usr/bin/env python
"""bench.py — edges·featdim/s of the fused 1-hop + 2-hop aggregation round (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...        (N > 1: one rank per GPU, rows sharded)

A "step" is one fused round  Y[:, 0:d] = A1·X, Y[:, d:2d] = A2·X  over the workload named in `config.workload`
(north-star target: uniform random graph |V|=10 000 per GPU, |E|=200 000 per GPU, d=128, fp32, explicit fp32
adjacency values).  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_PER_GPU = 10_000
E_PER_GPU = 200_000
FEAT = 128
L2_FLUSH_BYTES = 256 << 20
L2_BYTES = 126 << 20
DEFAULT_SPLITS = None      # arithmetic of the tensor-core path: None = the library default (h2gcn_b200._cabi.DEFAULT_SPLITS)


_JSON_OUT = sys.stdout


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(n_rows, n_cols, nnz, d, explicit_vals=True):
    """every distinct byte once.  rowptr is int64 here (8 B/row instead of the survey's 4)."""
    return 2 * (n_rows + 1) * 8 + nnz * (4 + (4 if explicit_vals else 0)) + n_cols * d * 4 + 2 * n_rows * d * 4


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz),
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def build_workload(n, n_edges, seed):
    from h2gcn_b200.utils import synth
    return synth.uniform_graph(n, n_edges, seed=seed)


def cpu_baseline(hops_host, x, budget_s=10.0, min_rounds=3, y_gpu=None):
    """Oracle C restatement (OpenMP over rows, every host thread) on full rounds of the SAME workload.  With `y_gpu`
    (the timed path's output for the same input) it also reports the parity of the measured configuration."""
    from oracle import cbind  # the one place bench.py may run oracle/ (cpu_baseline / --impl reference)
    (rp1, c1, v1), (rp2, c2, v2) = hops_host
    n, d = x.shape
    y = np.empty((n, 2 * d), dtype=np.float32)
    cbind.fused_round(rp1, c1, v1, rp2, c2, v2, x, y)  # warm-up
    parity = None
    if y_gpu is not None:
        parity = {"max_abs_err_over_max_abs_ref": float(np.abs(y_gpu.astype(np.float64) - y).max() / np.abs(y).max()),
                  "tolerance": 1e-4, "against": "oracle C port (in-order fp32), same input, full output"}
    rounds, t0 = 0, time.perf_counter()
    while rounds < min_rounds or time.perf_counter() - t0 < budget_s:
        cbind.fused_round(rp1, c1, v1, rp2, c2, v2, x, y)
        rounds += 1
    dt = time.perf_counter() - t0
    nnz = len(c1) + len(c2)
    return {"value": nnz * d * rounds / dt, "unit": "edges*featdim/s", "cores": cbind.max_threads(), "kind": "port",
            "sample": f"{rounds} full fused rounds of the same workload in {dt:.1f} s (C restatement of the TF-CPU "
                      f"functor, OpenMP over rows; TensorFlow is not installable here)", "ms_per_round": 1e3 * dt / rounds,
            "parity_of_timed_path": parity}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port — the reference is TensorFlow, which this image lacks)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cbind
    from oracle import h2gcn_oracle as O
    adj = build_workload(N_PER_GPU, E_PER_GPU, seed=0)
    rp2, col2 = cbind.hop2_csr(adj.indptr, adj.indices)
    import scipy.sparse as sp
    p2 = sp.csr_matrix((np.ones(len(col2), dtype=np.float32), col2, rp2), shape=adj.shape)
    a1 = O.sym_normalize(adj)[0]
    a2 = O.sym_normalize(p2)[0]
    hops = [(a1.indptr.astype(np.int64), a1.indices.astype(np.int32), a1.data.astype(np.float32)),
            (a2.indptr.astype(np.int64), a2.indices.astype(np.int32), a2.data.astype(np.float32))]
    from h2gcn_b200.utils import synth
    x = synth.features(N_PER_GPU, FEAT, 0)
    (rp1, c1, v1), (rpb, c2, v2) = hops
    y = np.empty((N_PER_GPU, 2 * FEAT), dtype=np.float32)
    for _ in range(max(1, args.warmup)):
        cbind.fused_round(rp1, c1, v1, rpb, c2, v2, x, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cbind.fused_round(rp1, c1, v1, rpb, c2, v2, x, y)
    dt = time.perf_counter() - t0
    nnz = len(c1) + len(c2)
    val = nnz * FEAT * args.steps / dt
    line = {"impl": "reference", "metric": "edges*featdim/sec on fused 2-hop SpMM", "value": val,
            "unit": "edges*featdim/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"uniform random graph |V|={N_PER_GPU} |E|={E_PER_GPU} d={FEAT} fp32, seed 0",
                       "nnz1": len(c1), "nnz2": len(c2), "note": "runs once on rank 0 (CPU), independent of --gpus"},
            "cpu_baseline": {"value": val, "unit": "edges*featdim/s", "cores": cbind.max_threads(), "kind": "port",
                             "sample": f"{args.steps} full fused rounds; C restatement of tf.sparse.sparse_dense_matmul "
                                       "(TF-CPU functor order) with OpenMP over rows — TensorFlow itself is absent"},
            "e2e": {"value": val, "unit": "edges*featdim/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--factored", action="store_true", help="index-only CSR (val = dinv_i*dinv_j rebuilt in-kernel)")
    ap.add_argument("--mode", default="auto", choices=["auto", "csr", "tensor"],
                    help="hop storage: auto = by density (dense-ish hops on tcgen05), csr = fp32 gather everywhere")
    ap.add_argument("--streams", type=int, default=2, help="caller streams the independent steps are issued on (round-robin)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"], help="N>1: hop-boundary exchange (p2p = fused into the pack kernel over peer memory)")
    ap.add_argument("--splits", default=DEFAULT_SPLITS, help="arithmetic of the tensor-core path: 2 | 3 (bf16 pieces), i8x2 | i8x3 "
                    "(int8 digits with per-4-row block exponents, exact int32 accumulation)")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON: anything a library writes to fd 1 (NCCL prints its version banner there
    # when NCCL_DEBUG is set) is sent to stderr instead
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    from h2gcn_b200 import _cabi as _c
    args.splits = _c.splits_code(args.splits)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from h2gcn_b200 import _cabi
    from h2gcn_b200.parallel import ShardedGraph
    from h2gcn_b200.utils import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 through torchrun)"
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()

    # ---- workload (untimed set-up: graph, GPU adjacency-power precompute, plan) ------------------------------------
    n = N_PER_GPU * world
    adj = build_workload(n, E_PER_GPU * world, seed=0)
    d = FEAT
    t0 = time.perf_counter()
    g = ShardedGraph(adj, rank, world, dev, factored=args.factored, mode=args.mode, splits=args.splits, exchange=args.exchange)
    torch.cuda.synchronize()
    t_pre = time.perf_counter() - t0
    # R independent replicas of the round's whole working set (graph arrays, X, Y, scratch), visited round-robin, so
    # that consecutive timed steps never find their inputs in the 126 MB L2 ("inputs larger than L2")
    # bitmap of the dense hop + CSR of the sparse hop + X, packed X, partial tiles (~3 x N d 4) + Y
    ws_est = g.n_local * n // 8 + 8 * (g.nnz1_global // world) + 3 * n * d * 4 + 2 * g.n_local * d * 4
    R = max(3, min(8, -(-(2 * L2_BYTES) // max(1, ws_est))))
    R = -(-R // args.streams) * args.streams   # a multiple of the stream count: a replica always runs on the same stream
    graphs = [g] + [ShardedGraph(adj, rank, world, dev, factored=args.factored, mode=args.mode, splits=args.splits, exchange=args.exchange)
                    for _ in range(R - 1)]
    x_full = synth.features(n, d, 0)
    xs = [torch.from_numpy(x_full[g.row_begin:g.row_end]).to(dev) for _ in range(R)]
    if world > 1 and g.exchange == "p2p":        # inputs live in symmetric memory: the exchange is zero-copy
        for r in range(R):
            buf = graphs[r].input_buffer(d)
            buf.copy_(xs[r])
            xs[r] = buf
    ys = [torch.empty(g.n_local, 2 * d, device=dev) for _ in range(R)]
    x_local, y = xs[0], ys[0]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    # consecutive steps are independent rounds (different replicas): issue them on `--streams` caller streams round-robin
    # so that the pack / fix-up of one round can overlap the tensor-core kernel of its neighbour
    main_stream = torch.cuda.current_stream()
    lanes = [torch.cuda.Stream(device=dev) for _ in range(args.streams)] if args.streams > 1 else [main_stream]

    def step(k=0):
        with torch.cuda.stream(lanes[k % len(lanes)]):
            graphs[k % R].round(xs[k % R], ys[k % R], [0, d])

    def fork():
        for ln in lanes:
            ln.wait_stream(main_stream)

    def join():
        for ln in lanes:
            main_stream.wait_stream(ln)

    for k in range(max(args.warmup, R)):
        step(k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---- timed region: EXACTLY K steps back to back between one CUDA-event pair on the launching stream -------------
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = _cabi.launch_count()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    ev0.record()
    fork()
    for k in range(args.steps):
        step(k)
    join()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    launches = _cabi.launch_count() - launches0
    clocks = sampler.finish()
    if world > 1:
        dist.barrier()
    total_ms = float(ev0.elapsed_time(ev1))
    # secondary figure: every step bracketed by its own event pair after an explicit L2 flush (256 MiB write)
    k_fl = min(args.steps, 50)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k_fl)]
    for a, b in ev:
        flush.zero_()
        a.record()
        fork()
        step(0)
        join()
        b.record()
    torch.cuda.synchronize()
    times = np.array([a.elapsed_time(b) for a, b in ev])  # ms

    # secondary figures: the same round with the other arithmetics of the tensor-core path (same replicas-larger-than-L2
    # scheme, fewer steps).  Never allowed to break the headline line.
    other = {}
    if world == 1 and args.mode == "auto":
        for name in ("i8x2", "i8x3", "bf16x2"):
            code = _cabi.splits_code(name)
            if code == args.splits:
                continue
            try:
                alt = [ShardedGraph(adj, rank, world, dev, factored=args.factored, mode=args.mode, splits=code,
                                    exchange=args.exchange) for _ in range(R)]
                ya = torch.empty(g.n_local, 2 * d, device=dev)

                def alt_step(k):
                    with torch.cuda.stream(lanes[k % len(lanes)]):
                        alt[k % R].round(xs[k % R], ys[k % R] if k % R else ya, [0, d])
                for k in range(max(args.warmup, R)):
                    alt_step(k)
                torch.cuda.synchronize()
                k_alt = min(args.steps, 300)
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record()
                fork()
                for k in range(k_alt):
                    alt_step(k)
                join()
                eb.record()
                torch.cuda.synchronize()
                other[_cabi.SPLITS_NAME[code]] = {"ms_per_step": float(ea.elapsed_time(eb)) / k_alt, "steps": k_alt}
                del alt
            except Exception as exc:  # noqa: BLE001
                other[name] = {"error": repr(exc)[:200]}
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
        nnz_t = torch.tensor([g.nnz_local], device=dev, dtype=torch.int64)
        dist.all_reduce(nnz_t)
        nnz_total = int(nnz_t.item())
    else:
        nnz_total = g.nnz_local
    ms_per_step = total_ms / args.steps
    value = nnz_total * d / (ms_per_step * 1e-3)

    # ---- precompute metric: exact-2-hop pattern + normalisation on the GPU, 2-paths/s and nnz2/s -----------
    precompute = None
    if world == 1:
        from h2gcn_b200 import ops
        a = adj.tocsr()
        rp_d = torch.from_numpy(a.indptr.astype(np.int64)).to(dev)
        col_d = torch.from_numpy(a.indices.astype(np.int32)).to(dev)
        two_paths = int((np.diff(a.indptr).astype(np.int64) ** 2).sum())
        for _ in range(2):
            ops.hop2_pattern(rp_d, col_d)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            rp2_d, col2_d = ops.hop2_pattern(rp_d, col_d)        # count -> scan -> (host sync) -> fill
            ops.sym_normalize(rp2_d, col2_d)
        torch.cuda.synchronize()
        t_gpu = (time.perf_counter() - t0) / reps
        precompute = {"what": "nhoodSplit(adj, 2)[2] + SYM normalize on the GPU (h2_hop2_count/fill, h2_sym_normalize), "
                              "host wall clock incl. the one count->alloc->fill sync",
                      "seconds": t_gpu, "two_paths_per_s": two_paths / t_gpu, "nnz2_per_s": int(col2_d.numel()) / t_gpu,
                      "two_paths": two_paths, "nnz2": int(col2_d.numel())}
        if not args.no_cpu_baseline:
            from oracle import cbind
            t0 = time.perf_counter()
            cbind.hop2_csr(a.indptr, a.indices)
            precompute["cpu_port_seconds"] = time.perf_counter() - t0
            precompute["cpu_port_cores"] = cbind.max_threads()

    # ---- e2e: the host-buffer C-ABI call (X host->device, round, Y device->host inside the timed region) -------------
    e2e = None
    if world == 1:
        from h2gcn_b200.ops import HostGraph
        hg = HostGraph(g.hops_host(), g.n_local, n, d_max=d, dinv_host=[h.dinv.cpu().numpy() for h in g.hops],
                       row_begin=g.row_begin, mode=args.mode, splits=args.splits)
        xh = torch.from_numpy(x_full).pin_memory()
        yh = torch.empty(g.n_local, 2 * d).pin_memory()
        for _ in range(3):
            hg.round(xh, yh)
        k_e2e = min(args.steps, 50)
        dt = 0.0
        for _ in range(k_e2e):
            flush.zero_()                    # cold L2 for every call; the flush itself is outside the timed part
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            hg.round(xh, yh)                 # H2D X, fused round, D2H Y, synchronise (inside the C call)
            dt += time.perf_counter() - t0
        e2e = {"value": nnz_total * d * k_e2e / dt, "unit": "edges*featdim/s", "h2d_bytes_per_step": n * d * 4,
               "d2h_bytes_per_step": g.n_local * 2 * d * 4, "ms_per_step": 1e3 * dt / k_e2e, "steps": k_e2e,
               "api": "h2_graph_round_host (adjacency resident, X in / Y out through pinned host buffers, sync per call); host wall clock around each call, L2 flushed before each call outside the timed part"}
        assert float((yh.to(dev) - y).abs().max()) == 0.0, "host-buffer path and device path disagree"
        hg.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    balg = algorithmic_bytes(g.n_local, n, g.nnz_local, d, explicit_vals=not args.factored)
    kern_ms = ms_per_step if world == 1 else float(ev0.elapsed_time(ev1)) / args.steps
    achieved = balg / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and world == 1 and args.mode == "auto":   # measured per arithmetic, for this workload only
        traffic = json.load(open(tp)).get("fused_round_dram_bytes_per_launch", {}).get(_cabi.SPLITS_NAME[args.splits])
    # arithmetic types on the path: fp32 in / out and on the CSR hops; the tensor-core hops multiply 0/1 by int8 digits
    # (int32 accumulation, exact) or by bf16 pieces (fp32 accumulation)
    dtype = "f32" if not g.plan.tensor_idx else ("f32+i8" if args.splits in (_cabi.H2_SPLITS_I8X2, _cabi.H2_SPLITS_I8X3) else "f32+bf16")
    line = {
        "metric": "edges*featdim/sec on fused 2-hop SpMM", "value": value, "unit": "edges*featdim/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"uniform random graph |V|={n} |E|={E_PER_GPU * world} d={d} fp32 "
                               f"({'factored dinv' if args.factored else 'explicit fp32'} adjacency values), seed 0, "
                               f"rows sharded over {world} GPU(s)",
                   "exchange": (g.exchange + (" (all-gather fused into the pack kernel: peer-memory loads over NVLink)"
                                              if g.exchange == "p2p" else " all-gather, then the round")) if world > 1 else None,
                   "n_vertices": n, "nnz1": g.nnz1_global, "nnz2_local": g.nnz2_local, "nnz_local": g.nnz_local,
                   "nnz_total": nnz_total, "max_row_nnz": g.max_row_nnz, "kernel": g.plan.kernel_name,
                   "l2": f"inputs larger than L2: {R} independent replicas of the working set (graph arrays, X, Y, scratch; "
                         f"~{ws_est >> 20} MiB each) visited round-robin; K steps back to back between ONE CUDA-event "
                         f"pair on the launching stream ({args.streams} caller stream(s) forked/joined inside the pair), max over ranks",
                   "ms_per_step_l2_flush_events": float(times.mean()), "ms_per_step_l2_flush_events_min": float(times.min()),
                   "l2_flush_note": f"secondary: {k_fl} steps, each after a {L2_FLUSH_BYTES >> 20} MiB L2-flush write and "
                                    "bracketed by its own event pair (includes ~4 us of event overhead per step)",
                   "precompute_s": t_pre, "wall_s_timed_region": wall,
                   "arithmetic": _cabi.SPLITS_NAME[args.splits], "other_arithmetics": other},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": balg,
                     "kernel_ms": kern_ms, "note": "compulsory bytes of one round (SURVEY §8d, explicit fp32 values, "
                                                    "int64 rowptr) / mean time of one round (ALL its launches: pack, "
                                                    "tcgen05 MMA, fix-up, CSR gather) on rank 0"},
        "clocks": clocks, "gpu_launches": int(launches),
    }
    if e2e is not None:
        line["e2e"] = e2e
    if precompute is not None:
        line["config"]["precompute"] = precompute
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(g.hops_host(), x_full, y_gpu=y.cpu().numpy())
        line["config"]["parity"] = line["cpu_baseline"].pop("parity_of_timed_path")
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
