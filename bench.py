#!/usr/bin/env python
"""bench.py — edges·featdim/s of the fused 1-hop + 2-hop aggregation round (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N ...        (N > 1: one rank per GPU, rows sharded)

A "step" is one fused round  Y[:, 0:d] = A1·X, Y[:, d:2d] = A2·X  over the workload named in `config.workload`:

  N = 1   the north-star point (BASELINE.json metric config): uniform random graph |V| = 10 000, |E| = 200 000, d = 128,
          fp32 in / out, explicit fp32 adjacency values.
  N > 1   BASELINE config 4: Graph500-parameter R-MAT |V| = 2^20, 16 M edge draws, d = 128, rows sharded over the N GPUs
          (strong scaling: the graph is the same for every N > 1), one all-gather of X per round fused into the first
          kernel of the round.  `config.secondary` carries the uniform weak-scaling line of round 1.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
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
RMAT_N = int(os.environ.get("H2_BENCH_RMAT_N", 1 << 20))           # BASELINE config 4 (override only for dry runs)
RMAT_E = int(os.environ.get("H2_BENCH_RMAT_E", 16 * (1 << 20)))
CFG5_N = int(os.environ.get("H2_BENCH_CFG5_N", 4 << 20))           # BASELINE config 5 (override for smaller boxes / dry runs)
CFG5_E = int(os.environ.get("H2_BENCH_CFG5_E", 64 << 20))
CFG5_FEAT = int(os.environ.get("H2_BENCH_CFG5_FEAT", 256))
CFG5_GAMMA = 2.5
L2_FLUSH_BYTES = 256 << 20
L2_BYTES = 126 << 20
DEFAULT_SPLITS = None      # arithmetic of the tensor-core path: None = the library default (i8x3, fp32-equivalent)
METRIC = "edges*featdim/sec on fused 2-hop SpMM"


_JSON_OUT = sys.stdout


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(n_rows, n_cols, nnz, d, explicit_vals=True):
    """every distinct byte once.  rowptr is int64 here (8 B/row instead of the survey's 4)."""
    return 2 * (n_rows + 1) * 8 + nnz * (4 + (4 if explicit_vals else 0)) + n_cols * d * 4 + 2 * n_rows * d * 4


def host_threads():
    """Threads for the CPU arm: every core of the box, stated explicitly — torchrun exports OMP_NUM_THREADS=1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


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


# ----------------------------------------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------------------------------------
def workload_for(world):
    """(kind, n, n_edges, d, description).  The description is the SAME string in both arms (`same_config`)."""
    if world == 1:
        return ("uniform", N_PER_GPU, E_PER_GPU, FEAT,
                f"uniform random graph |V|={N_PER_GPU} |E|={E_PER_GPU} d={FEAT} fp32 (explicit fp32 adjacency values), seed 0")
    return ("rmat", RMAT_N, RMAT_E, FEAT,
            f"R-MAT (a,b,c)=(.57,.19,.19) |V|={RMAT_N} {RMAT_E} edge draws (symmetrised, deduplicated, no self loops) "
            f"d={FEAT} fp32, seed 1, rows sharded over the GPUs (BASELINE config 4)")


def build_workload(kind, n, n_edges):
    from h2gcn_b200.utils import synth
    if kind == "uniform":
        return synth.uniform_graph(n, n_edges, seed=0)
    return synth.rmat_graph(n, n_edges, seed=1)


def sample_hops_cpu(adj, rows, deg2_all):
    """CSR of the sampled ROWS of the normalised hop adjacencies, computed on the host from the definition
    (nhoodSplit / normalize, h2gcn/datasets/_dataset.py:109-158): P1 = A, P2 = bin((A+I)^2) - bin(A+I), values
    fp32(fp64(dinv_i) * dinv_j) with dinv = deg^-1/2 of the hop's own degree vector (inf -> 0)."""
    import scipy.sparse as sp
    n = adj.shape[0]
    a = adj.tocsr()
    ai = (a + sp.identity(n, dtype=np.float32, format="csr")).tocsr()
    sub = ai[rows]
    two = (sub @ ai).tocsr()
    two.data[:] = 1.0
    p2 = (two - two.multiply(sub.astype(bool))).tocsr()     # drop distance <= 1
    p2.eliminate_zeros()
    p2.sort_indices()
    p1 = a[rows].tocsr()
    p1.sort_indices()
    deg1 = np.diff(a.indptr).astype(np.float64)
    out = []
    for p, deg in ((p1, deg1), (p2, np.asarray(deg2_all, dtype=np.float64))):
        with np.errstate(divide="ignore"):
            dinv = np.power(deg, -0.5)
        dinv[np.isinf(dinv)] = 0.0
        counts = np.diff(p.indptr)
        vals = (np.repeat(dinv[rows], counts) * dinv[p.indices]).astype(np.float32)
        out.append((p.indptr.astype(np.int64), p.indices.astype(np.int32), vals))
    return out


def full_shard_check(g, x_all, y, d):
    """EVERY row of the rank's tensor-core hops against the same hops on the fp32 CSR gather kernel (an independent code
    path over the same device arrays), outside the timed region: the row sample above would miss an error confined to one
    128-row half of one tile (cross-CTA hand-over)."""
    import torch
    from h2gcn_b200 import ops
    out = {}
    for h in g.plan.tensor_idx:
        ref = torch.empty(g.n_local, d, device=y.device, dtype=y.dtype)
        ops.HopPlan([g.hops[h]], factored=True, mode="csr").run(x_all, ref, [0])
        got = y[:, h * d:(h + 1) * d].float()
        ref = ref.float()
        scale = float(ref.abs().max())
        err = (got - ref).abs()
        rn = ref.abs().amax(dim=1)
        live = rn > 0
        out[f"hop{h + 1}"] = {"rows": int(g.n_local), "max_abs_err_over_max_abs_ref": float(err.max()) / max(scale, 1e-30),
                             "row_wise_max_rel_err": float((err.amax(dim=1)[live] / rn[live]).max()) if bool(live.any()) else 0.0}
        del ref, got, err
    return out


def pick_sample_rows(adj, budget_nnz, seed=7):
    """Seeded row sample whose 2-hop rows hold about `budget_nnz` entries (estimated from sum of neighbour degrees)."""
    rng = np.random.default_rng(seed)
    n = adj.shape[0]
    deg = np.diff(adj.indptr).astype(np.int64)
    est = np.minimum(np.asarray(adj @ deg.astype(np.float64)).ravel(), n)   # upper estimate of deg2
    order = rng.permutation(n)
    csum = np.cumsum(est[order] + deg[order] + 1)
    k = int(np.searchsorted(csum, budget_nnz)) + 1
    return np.sort(order[:max(16, min(k, 4096))])


def parity_metrics(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    rn = np.abs(ref).max(axis=1)
    en = np.abs(got - ref).max(axis=1)
    live = rn > 0
    return {"max_abs_err_over_max_abs_ref": float(en.max() / max(np.abs(ref).max(), 1e-30)),
            "row_wise_max_rel_err": float((en[live] / rn[live]).max()) if live.any() else 0.0,
            "zero_rows_exact": bool((np.abs(got[~live]) == 0).all()) if (~live).any() else True,
            "tolerance": 1e-4}


def cpu_baseline(hops_host, x, budget_s=10.0, min_rounds=3, y_gpu=None):
    """Oracle C restatement (OpenMP over rows, every host thread) on full rounds of the SAME workload.  With `y_gpu`
    (the timed path's output for the same input) it also reports the parity of the measured configuration."""
    from oracle import cbind  # the one place bench.py may run oracle/ (cpu_baseline / --impl reference)
    (rp1, c1, v1), (rp2, c2, v2) = hops_host
    n, d = x.shape
    thr = host_threads()
    y = np.empty((len(rp1) - 1, 2 * d), dtype=np.float32)
    cbind.fused_round(rp1, c1, v1, rp2, c2, v2, x, y, threads=thr)  # warm-up
    parity = None
    if y_gpu is not None:
        parity = parity_metrics(y_gpu, y)
        parity["against"] = "oracle C port (in-order fp32), same input, full output"
    rounds, t0 = 0, time.perf_counter()
    while rounds < min_rounds or time.perf_counter() - t0 < budget_s:
        cbind.fused_round(rp1, c1, v1, rp2, c2, v2, x, y, threads=thr)
        rounds += 1
    dt = time.perf_counter() - t0
    nnz = len(c1) + len(c2)
    return {"value": nnz * d * rounds / dt, "unit": "edges*featdim/s", "cores": thr, "kind": "port",
            "sample": f"{rounds} full fused rounds of the same workload in {dt:.1f} s (C restatement of the TF-CPU "
                      f"functor, OpenMP over rows; TensorFlow is not installable here)", "ms_per_round": 1e3 * dt / rounds,
            "parity_of_timed_path": parity}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port — the reference is TensorFlow, which this image lacks) on
    the workload the repo arm runs at this N, with every host thread.  N = 1: full rounds.  N > 1 (R-MAT, ~2e10 stored
    2-hop entries): a bounded seeded ROW SAMPLE of the same graph per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cbind
    from h2gcn_b200.utils import synth
    world = args.gpus
    kind, n, n_edges, d, desc = workload_for(world)
    thr = host_threads()
    adj = build_workload(kind, n, n_edges)
    x = synth.features(n, d, 0)
    rp2_all, col2_all = None, None
    if world == 1:
        rp2_all, col2_all = cbind.hop2_csr(adj.indptr, adj.indices, threads=thr)
        deg2 = np.diff(rp2_all)
        rows = np.arange(n)
        sample = "every row: full fused rounds"
    else:
        import ctypes
        rp2 = np.zeros(n + 1, dtype=np.int64)                    # counting pass only: the 2-hop degrees of every vertex
        cbind.lib().oracle_hop2_csr(ctypes.c_int32(n), cbind._p(np.ascontiguousarray(adj.indptr, dtype=np.int32)),
                                    cbind._p(np.ascontiguousarray(adj.indices, dtype=np.int32)), cbind._p(rp2), None,
                                    ctypes.c_int(thr))
        deg2 = np.diff(rp2)
        rows = pick_sample_rows(adj, budget_nnz=6e7)
        sample = f"seeded sample of {len(rows)} rows of the same graph per step (the full 2-hop pattern has {int(deg2.sum())} entries)"
    (rp1, c1, v1), (rpb, c2, v2) = sample_hops_cpu(adj, rows, deg2)
    y = np.empty((len(rows), 2 * d), dtype=np.float32)
    for _ in range(max(1, args.warmup)):
        cbind.fused_round(rp1, c1, v1, rpb, c2, v2, x, y, threads=thr)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cbind.fused_round(rp1, c1, v1, rpb, c2, v2, x, y, threads=thr)
    dt = time.perf_counter() - t0
    nnz = len(c1) + len(c2)
    val = nnz * d * args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": val,
            "unit": "edges*featdim/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "n_vertices": n, "nnz1": int(adj.nnz), "nnz2": int(deg2.sum()),
                       "rows_per_step": int(len(rows)), "nnz_per_step": int(nnz),
                       "note": "runs on rank 0 (CPU); throughput = stored entries of the rows processed per step x d / time"},
            "cpu_baseline": {"value": val, "unit": "edges*featdim/s", "cores": thr, "kind": "port",
                             "sample": f"{args.steps} steps, {sample}; C restatement of tf.sparse.sparse_dense_matmul "
                                       "(TF-CPU functor order) with OpenMP over rows — TensorFlow itself is absent"},
            "e2e": {"value": val, "unit": "edges*featdim/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


# ----------------------------------------------------------------------------------------------------------------------
def timed_bursts(step, fork, join, steps, n_bursts, barrier=None):
    """`n_bursts` bursts of EXACTLY `steps` steps, each between its own CUDA-event pair on the launching stream (forks and
    joins of the caller streams inside the pair).  Returns the per-burst milliseconds."""
    import torch
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_bursts)]
    k0 = 0
    for a, b in ev:
        if barrier is not None:
            barrier()
        a.record()
        fork()
        for k in range(steps):
            step(k0 + k)
        join()
        b.record()
        k0 += steps
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in ev])


def secondary_configs(dev, steps):
    """BASELINE configs 2 and 3 as secondary figures: the full H2GCN-2 forward on the Planetoid Cora fixture (fused launch
    sequence, eager and as ONE CUDA graph) next to the oracle's forward on the host, and one fused round on the
    syn-products proxy (preferential attachment |V| = 10 000, m = 6, d = 100)."""
    import scipy.sparse as sp
    import torch
    from h2gcn_b200.datasets._dataset import GraphData
    from h2gcn_b200.models import parse_network_setup
    from h2gcn_b200.models.H2GCN import H2GCN
    from h2gcn_b200.ops import HopPlan
    from h2gcn_b200.utils import synth
    out = {}
    gold = os.path.join(ROOT, "tests", "golden", "planetoid_cora.npz")
    if os.path.exists(gold):
        z = np.load(gold)
        n = len(z["adj_indptr"]) - 1
        adj = sp.csr_matrix((z["adj_data"], z["adj_indices"], z["adj_indptr"]), shape=(n, n))
        feat = sp.csr_matrix((z["feat_data"], z["feat_indices"], z["feat_indptr"]), shape=tuple(z["feat_shape"]))
        data = GraphData(adj, feat.tolil(), device=dev)
        with np.errstate(divide="ignore"):
            data.row_normalize_features()
        data.adj_remove_eye()
        t = data.getTensors(getAdjNormHops=["1", "2"])
        setup = "M64-R-T1-G-V-T2-G-V-C1-C2-D0.5-MO"
        model = H2GCN(parse_network_setup(setup, int(z["num_labels"]), _dense_units=64, _dropout_rate=0.5))
        logits = model(t.adj, t.features, t.adj_hops)
        prog = model._fused_program(t.features, t.adj_hops)

        def timeit(fn, reps):
            for _ in range(5):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        reps = min(steps, 200)
        eager_ms = timeit(lambda: model(t.adj, t.features, t.adj_hops), reps)
        graph_ms = None
        try:
            graph = prog.capture(t.features)
            graph_ms = timeit(graph.replay, reps)
            same = bool(torch.equal(prog.graph_out, logits))
        except Exception as exc:  # noqa: BLE001
            same = repr(exc)[:160]
        from oracle import h2gcn_oracle as O
        hops = [(h.indices[:, 0].cpu().numpy(), h.indices[:, 1].cpu().numpy(), h.values.cpu().numpy()) for h in t.adj_hops]
        fc = t.features.indices.cpu().numpy()
        W = [w.cpu().numpy() for w in model.trainable_variables]
        conf = parse_network_setup(setup, int(z["num_labels"]), _dense_units=64, _dropout_rate=0.5)
        t0 = time.perf_counter()
        ref = O.forward(conf, W, (fc[:, 0], fc[:, 1], t.features.values.cpu().numpy()), n, hops)
        cpu_ms = 1e3 * (time.perf_counter() - t0)
        out["cfg2_cora_forward"] = {"what": f"H2GCN-2 forward ({setup}) on the Planetoid Cora fixture, N={n}, F={feat.shape[1]}, p=64",
                                    "ms_eager_launch_sequence": eager_ms, "ms_cuda_graph": graph_ms, "graph_equals_eager": same,
                                    "cpu_oracle_ms": cpu_ms, "cpu_kind": "numpy/scipy restatement, 1 thread",
                                    "parity": parity_metrics(logits.cpu().numpy(), ref)}
    a = synth.preferential_attachment(10_000, 6, seed=0)
    n, d = a.shape[0], 100
    t = GraphData(a, sp.identity(n, dtype=np.float32, format="csr"), device=dev).getTensors(getAdjNormHops=["1", "2"])
    plan = HopPlan(t.adj_hops)
    x = torch.from_numpy(synth.features(n, d, 0)).to(dev)
    y = torch.empty(n, 2 * d, device=dev)
    for _ in range(5):
        plan.run(x, y, [0, d])
    reps = min(steps, 200)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.run(x, y, [0, d])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nnz = sum(h.nnz for h in t.adj_hops)
    from oracle import cbind
    h = t.adj_hops
    ref = cbind.fused_round(h[0].rowptr.cpu().numpy(), h[0].col.cpu().numpy(), h[0].values.cpu().numpy(),
                            h[1].rowptr.cpu().numpy(), h[1].col.cpu().numpy(), h[1].values.cpu().numpy(), x.cpu().numpy(),
                            threads=host_threads())
    balg = algorithmic_bytes(n, n, nnz, d)
    out["cfg3_syn_products_round"] = {"what": "fused round on the syn-products proxy (preferential attachment |V|=10000, m=6, d=100; "
                                              "warm L2: the 33 MB working set fits the 126 MB L2)",
                                      "nnz1": h[0].nnz, "nnz2": h[1].nnz, "ms_per_round": ms, "kernel": plan.kernel_name,
                                      "edges_featdim_per_s": nnz * d / (ms * 1e-3),
                                      "roofline_frac_of_measured_hbm": balg / (ms * 1e-3) / 1e9 / peaks()[0],
                                      "parity": parity_metrics(y.cpu().numpy(), ref)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary figures (other arithmetics, cfg2 / cfg3)")
    ap.add_argument("--factored", action="store_true", help="index-only CSR (val = dinv_i*dinv_j rebuilt in-kernel)")
    ap.add_argument("--mode", default="auto", choices=["auto", "csr", "tensor"],
                    help="hop storage: auto = by density (dense-ish hops on tcgen05), csr = fp32 gather everywhere")
    ap.add_argument("--streams", type=int, default=2, help="caller streams the independent steps are issued on (round-robin)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"], help="N>1: hop-boundary exchange (p2p = fused into the pack kernel over peer memory)")
    ap.add_argument("--splits", default=DEFAULT_SPLITS, help="arithmetic of the tensor-core path: i8x3 (default: 3 int8 digits + row "
                    "exponents, fp32-equivalent) | i8x2 | 2 | 3 (bf16 pieces)")
    ap.add_argument("--workload", default="auto", choices=["auto", "uniform", "cfg5"],
                    help="N>1: auto = R-MAT config 4, uniform = the weak-scaling uniform graph of round 1; cfg5 (any N) = BASELINE config 5: "
                         "power-law |V|=4M |E|=64M d=256, bf16 feature rows in / out")
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 through torchrun)"
    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    if args.workload == "cfg5":
        return main_cfg5(args)
    if world == 1:
        return main_single(args)
    return main_multi(args)


# ----------------------------------------------------------------------------------------------------------------------
# N = 1: the north-star point
# ----------------------------------------------------------------------------------------------------------------------
def main_single(args):
    import torch
    from h2gcn_b200 import _cabi
    from h2gcn_b200.parallel import ShardedGraph
    from h2gcn_b200.utils import synth

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _cabi.lib()
    kind, n, n_edges, d, desc = workload_for(1)
    adj = build_workload(kind, n, n_edges)
    t0 = time.perf_counter()
    mk = lambda code: ShardedGraph(adj, 0, 1, dev, factored=args.factored, mode=args.mode, splits=code, exchange=args.exchange)
    g = mk(args.splits)
    torch.cuda.synchronize()
    t_pre = time.perf_counter() - t0
    # R independent replicas of the round's whole working set (graph arrays, X, Y, scratch), visited round-robin, so
    # that consecutive timed steps never find their inputs in the 126 MB L2 ("inputs larger than L2")
    ws_est = g.n_local * n // 8 + 8 * g.nnz1_global + 3 * n * d * 4 + 2 * g.n_local * d * 4
    R = max(3, min(8, -(-(2 * L2_BYTES) // max(1, ws_est))))
    R = -(-R // args.streams) * args.streams   # a multiple of the stream count: a replica always runs on the same stream
    graphs = [g] + [mk(args.splits) for _ in range(R - 1)]
    x_full = synth.features(n, d, 0)
    xs = [torch.from_numpy(x_full).to(dev) for _ in range(R)]
    ys = [torch.empty(n, 2 * d, device=dev) for _ in range(R)]
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    main_stream = torch.cuda.current_stream()
    lanes = [torch.cuda.Stream(device=dev) for _ in range(args.streams)] if args.streams > 1 else [main_stream]

    def fork():
        for ln in lanes:
            ln.wait_stream(main_stream)

    def join():
        for ln in lanes:
            main_stream.wait_stream(ln)

    def make_step(gs, out0=None):
        def step(k=0):
            with torch.cuda.stream(lanes[k % len(lanes)]):
                gs[k % R].round(xs[k % R], ys[k % R] if (out0 is None or k % R) else out0, [0, d])
        return step

    def measure(gs, steps, n_bursts, out0=None):
        """warm-up, then `n_bursts` bursts of `steps` pipelined steps + the one-round latency (flushed L2, one stream)."""
        step = make_step(gs, out0)
        for k in range(max(args.warmup, R)):
            step(k)
        torch.cuda.synchronize()
        bursts = timed_bursts(step, fork, join, steps, n_bursts)
        k_lat = 40
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k_lat)]
        for a, b in ev:
            flush.zero_()
            a.record()
            gs[0].round(xs[0], ys[0] if out0 is None else out0, [0, d])
            b.record()
        torch.cuda.synchronize()
        lat = np.array([a.elapsed_time(b) for a, b in ev])
        return bursts, lat

    # ---- timed region: bursts of EXACTLY K steps back to back, each between one CUDA-event pair ----------------------
    n_bursts = int(min(31, max(5, 30_000 // max(1, args.steps)))) | 1     # odd, so that the median is one real burst
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _cabi.launch_count()
    wall0 = time.perf_counter()
    bursts, lat = measure(graphs, args.steps, n_bursts)
    wall = time.perf_counter() - wall0
    n_launch_total = _cabi.launch_count() - launches0
    clocks = sampler.finish()
    steps_run = max(args.warmup, R) + args.steps * n_bursts + len(lat)
    launches_per_step = n_launch_total / steps_run
    total_ms = float(np.median(bursts))
    ms_per_step = total_ms / args.steps
    nnz_total = g.nnz_local
    value = nnz_total * d / (ms_per_step * 1e-3)
    y_head = ys[0].cpu().numpy()
    x_host_check = x_full

    peak, peak_src = peaks()
    balg = algorithmic_bytes(g.n_local, n, g.nnz_local, d, explicit_vals=not args.factored)

    def roofline_of(ms):
        ach = balg / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak}

    # ---- the other arithmetics of the tensor-core path, same protocol (fewer bursts).  Never allowed to break the line.
    other, other_y = {}, {}
    if not args.no_secondary and args.mode == "auto":
        for name in ("i8x3", "i8x2", "bf16x2"):
            code = _cabi.splits_code(name)
            if code == args.splits:
                continue
            try:
                alt = [mk(code) for _ in range(R)]
                ya = torch.empty(n, 2 * d, device=dev)
                b2, l2 = measure(alt, min(args.steps, 300), 7, out0=ya)
                ms = float(np.median(b2)) / min(args.steps, 300)
                other[name] = {"arithmetic": _cabi.SPLITS_NAME[code], "ms_per_step": ms, "value": nnz_total * d / (ms * 1e-3),
                               "roofline": roofline_of(ms), "ms_per_round_latency": float(np.median(l2)),
                               "bursts": 7, "steps_per_burst": min(args.steps, 300)}
                other_y[name] = ya.cpu().numpy()
                del alt
            except Exception as exc:  # noqa: BLE001
                other[name] = {"error": repr(exc)[:200]}

    # ---- precompute metric: exact-2-hop pattern + normalisation on the GPU, 2-paths/s and nnz2/s -----------
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
        cbind.hop2_csr(a.indptr, a.indices, threads=host_threads())
        precompute["cpu_port_seconds"] = time.perf_counter() - t0
        precompute["cpu_port_cores"] = host_threads()

    # ---- e2e: the host-buffer C-ABI call (X host->device, round, Y device->host inside the timed region) -------------
    from h2gcn_b200.ops import HostGraph
    hg = HostGraph(g.hops_host(), g.n_local, n, d_max=d, dinv_host=[h.dinv.cpu().numpy() for h in g.hops],
                   row_begin=g.row_begin, mode=args.mode, splits=args.splits)
    xh = torch.from_numpy(x_full).pin_memory()
    yh = torch.empty(g.n_local, 2 * d).pin_memory()
    for _ in range(3):
        hg.round(xh, yh)
    k_e2e = min(args.steps, 50)
    dts = []
    for _ in range(k_e2e):
        flush.zero_()                    # cold L2 for every call; the flush itself is outside the timed part
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        hg.round(xh, yh)                 # H2D X, fused round, D2H Y, synchronise (inside the C call)
        dts.append(time.perf_counter() - t0)
    dt_med = float(np.median(dts))
    e2e = {"value": nnz_total * d / dt_med, "unit": "edges*featdim/s", "h2d_bytes_per_step": n * d * 4,
           "d2h_bytes_per_step": g.n_local * 2 * d * 4, "ms_per_step": 1e3 * dt_med, "steps": k_e2e,
           "api": "h2_graph_round_host (adjacency resident, X in / Y out through pinned host buffers, sync per call); host wall "
                  "clock around each call (median), L2 flushed before each call outside the timed part"}
    assert float((yh.to(dev) - ys[0]).abs().max()) == 0.0, "host-buffer path and device path disagree"
    hg.close()

    secondary = {}
    if not args.no_secondary:
        try:
            secondary = secondary_configs(dev, args.steps)
        except Exception as exc:  # noqa: BLE001
            secondary = {"error": repr(exc)[:300]}

    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.mode == "auto":   # measured per arithmetic, for this workload only
        traffic = json.load(open(tp)).get("fused_round_dram_bytes_per_launch", {}).get(_cabi.SPLITS_NAME[args.splits])
    # arithmetic types on the path: fp32 in / out and on the CSR hops; the tensor-core hops multiply 0/1 by int8 digits
    # (int32 accumulation, exact) or by bf16 pieces (fp32 accumulation)
    i8 = args.splits in (_cabi.H2_SPLITS_I8X2, _cabi.H2_SPLITS_I8X3)
    dtype = "f32" if not g.plan.tensor_idx else ("f32+i8" if i8 else "f32+bf16")
    roof = roofline_of(ms_per_step)
    roof.update({"traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": balg, "kernel_ms": ms_per_step,
                 "note": "compulsory bytes of one round (SURVEY §8d, explicit fp32 values, int64 rowptr) / median time of "
                         "one pipelined round (ALL its launches: pack, tcgen05 pair kernel, CSR gather)",
                 "frac_at_one_round_latency": balg / (float(np.median(lat)) * 1e-3) / 1e9 / peak})
    line = {
        "metric": METRIC, "value": value, "unit": "edges*featdim/s",
        "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "ms_per_round_latency": float(np.median(lat)),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "arithmetic": _cabi.SPLITS_NAME[args.splits],
        "value_fp32_equiv": value if args.splits in (_cabi.H2_SPLITS_I8X3, 3) else (other.get("i8x3") or {}).get("value"),
        "config": {"workload": desc,
                   "n_vertices": n, "n_edges": n_edges, "nnz1": g.nnz1_global, "nnz2": g.nnz2_local, "nnz_total": nnz_total,
                   "max_row_nnz": g.max_row_nnz, "kernel": g.plan.kernel_name,
                   "l2": f"inputs larger than L2: {R} independent replicas of the working set (graph arrays, X, Y, scratch; "
                         f"~{ws_est >> 20} MiB each) visited round-robin; steps issued on {args.streams} caller stream(s) "
                         "forked/joined inside each event pair",
                   "timing": f"{n_bursts} bursts of exactly {args.steps} steps, each between one CUDA-event pair on the launching "
                             "stream; ms_per_step = MEDIAN burst / steps (a 20-step burst is ~1 ms: one burst alone is "
                             "dominated by fork/join and pipeline fill)",
                   "burst_ms_per_step": {"median": ms_per_step, "min": float(bursts.min()) / args.steps,
                                         "max": float(bursts.max()) / args.steps, "n": n_bursts},
                   "ms_per_round_latency_note": "ONE round on ONE stream after a 256 MiB L2-flush write, its own event pair "
                                                "(includes ~4 us of event overhead); median of 40 — H2GCN's rounds are dependent, "
                                                "this is what a forward pass pays per round",
                   "ms_per_round_latency_min": float(lat.min()),
                   "precompute_s": t_pre, "wall_s_timed_region": wall,
                   "other_arithmetics": other, "precompute": precompute, "secondary": secondary},
        "roofline": roof,
        "clocks": clocks, "gpu_launches": int(round(launches_per_step * args.steps)),
        "gpu_launches_per_step": launches_per_step,
        "e2e": e2e,
    }
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(g.hops_host(), x_host_check, y_gpu=y_head)
        par = line["cpu_baseline"].pop("parity_of_timed_path")
        from oracle import cbind
        ref = None
        for name, yy in other_y.items():     # parity of every arithmetic that was timed, both metrics
            if ref is None:
                hh = g.hops_host()
                ref = cbind.fused_round(hh[0][0], hh[0][1], hh[0][2], hh[1][0], hh[1][1], hh[1][2], x_host_check, threads=host_threads())
            other[name]["parity"] = parity_metrics(yy, ref)
        line["config"]["parity"] = par
        # repeatability of the timed path: the same round 200 more times into a scratch output, compared BIT for BIT with
        # the output the parity above was computed on (the tensor-core hop's cross-CTA hand-over)
        y_rep = torch.empty_like(ys[0])
        y_first = torch.from_numpy(np.ascontiguousarray(y_head)).to(dev)
        same = 0
        for _ in range(200):
            y_rep.fill_(float("nan"))
            graphs[0].round(xs[0], y_rep, [0, d])
            same += int(torch.equal(y_rep, y_first))
        par["bitwise_repeatable_launches"] = f"{same} of 200"
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


# ----------------------------------------------------------------------------------------------------------------------
# N > 1: BASELINE config 4 (R-MAT 1M / 16M), rows sharded, one fused all-gather per round
# ----------------------------------------------------------------------------------------------------------------------
def main_multi(args):
    import torch
    import torch.distributed as dist
    from h2gcn_b200 import _cabi
    from h2gcn_b200.parallel import ShardedGraph
    from h2gcn_b200.utils import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()
    if args.workload == "uniform":
        kind, n, n_edges, d = "uniform", N_PER_GPU * world, E_PER_GPU * world, FEAT
        desc = f"uniform random graph |V|={n} |E|={n_edges} d={d} fp32, seed 0, rows sharded over {world} GPUs (weak scaling)"
    else:
        kind, n, n_edges, d, desc = workload_for(world)
    adj = build_workload(kind, n, n_edges)
    t0 = time.perf_counter()
    big = kind == "rmat"
    g = ShardedGraph(adj, rank, world, dev, factored=args.factored or big, mode=args.mode, splits=args.splits,
                     exchange=args.exchange, explicit_vals=not big, balance="rows" if big else "auto")
    torch.cuda.synchronize()
    t_pre = time.perf_counter() - t0
    x_full = synth.features(n, d, 0)
    x_local = torch.from_numpy(x_full[g.row_begin:g.row_end]).to(dev)
    if g.exchange == "p2p":        # inputs live in symmetric memory: the exchange is zero-copy
        buf = g.input_buffer(d)
        buf.copy_(x_local)
        x_local = buf
    y = torch.empty(g.n_local, 2 * d, device=dev)

    def step(k=0):
        g.round(x_local, y, [0, d])

    torch.cuda.synchronize()
    dist.barrier()                 # the set-up times differ by seconds between ranks (rank 0 owns the hub rows)
    for k in range(args.warmup):
        step(k)
    torch.cuda.synchronize()
    dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _cabi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        step(k)
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    launches = _cabi.launch_count() - launches0
    clocks = sampler.finish()
    dist.barrier()
    my_ms = float(ev0.elapsed_time(ev1))
    tt = torch.tensor([my_ms], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    cnt = torch.tensor([g.nnz_local, g.nnz2_local], device=dev, dtype=torch.int64)
    dist.all_reduce(cnt)
    nnz_total, nnz2_total = int(cnt[0].item()), int(cnt[1].item())
    ms_per_step = total_ms / args.steps
    value = nnz_total * d / (ms_per_step * 1e-3)

    # ---- e2e at N GPUs: every rank feeds ITS rows of X from pinned host memory and reads ITS rows of Y back -----------
    xh = torch.from_numpy(x_full[g.row_begin:g.row_end].copy()).pin_memory()
    yh = torch.empty(g.n_local, 2 * d).pin_memory()
    k_e2e = min(args.steps, 10)
    for _ in range(2):
        x_local.copy_(xh, non_blocking=True)
        step()
        yh.copy_(y, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        x_local.copy_(xh, non_blocking=True)
        step()
        yh.copy_(y, non_blocking=True)
        torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item()) / k_e2e
    assert float((yh.to(dev) - y).abs().max()) == 0.0

    # ---- parity of THIS run: a seeded sample of rank 0's rows against the oracle C port on the host -------------------
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import cbind
        rng = np.random.default_rng(11)
        deg_all = g.deg2_host
        live = np.nonzero(deg_all[g.row_begin:g.row_end] > 0)[0]
        pick = np.sort(rng.choice(live, size=min(96, len(live)), replace=False)) if len(live) else np.arange(min(96, g.n_local))
        rows = pick + g.row_begin
        (rp1, c1, v1), (rp2, c2, v2) = sample_hops_cpu(adj, rows, deg_all)
        deg2_cpu = np.diff(rp2)
        ref = cbind.fused_round(rp1, c1, v1, rp2, c2, v2, x_full, threads=host_threads())
        got = y[torch.from_numpy(pick).to(dev)].cpu().numpy()
        parity = parity_metrics(got, ref)
        parity.update({"against": "oracle C port on the host: 2-hop rows rebuilt from the definition with scipy, in-order fp32 sums",
                       "rows_checked": int(len(rows)), "entries_checked": int(len(c1) + len(c2)),
                       "deg2_of_sample_matches_gpu": bool(np.array_equal(deg2_cpu, deg_all[rows]))})
    if rank == 0 and parity is not None and g.plan.tensor_idx:
        parity["full_shard_tensor_vs_csr"] = full_shard_check(g, torch.from_numpy(x_full).to(dev), y, d)
    if rank != 0:
        dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    balg = algorithmic_bytes(g.n_local, n, g.nnz_local, d, explicit_vals=True)
    kern_ms = my_ms / args.steps
    achieved = balg / (kern_ms * 1e-3) / 1e9
    i8 = args.splits in (_cabi.H2_SPLITS_I8X2, _cabi.H2_SPLITS_I8X3)
    dtype = "f32" if not g.plan.tensor_idx else ("f32+i8" if i8 else "f32+bf16")
    line = {
        "metric": METRIC, "value": value, "unit": "edges*featdim/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong" if big else "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "arithmetic": _cabi.SPLITS_NAME[args.splits],
        "config": {"workload": desc,
                   "exchange": g.exchange + (" (all-gather fused into the first kernel of the round: peer-memory loads over NVLink)"
                                             if g.exchange == "p2p" else " all-gather, then the round"),
                   "n_vertices": n, "n_edge_draws": n_edges, "nnz1": g.nnz1_global, "nnz2": nnz2_total, "nnz_total": nnz_total,
                   "nnz_local_rank0": g.nnz_local, "rows_rank0": g.n_local, "max_degree": g.max_deg1, "max_degree_2hop": g.max_deg2,
                   "zero_degree_rows": g.zero_deg1, "zero_degree_rows_2hop": g.zero_deg2,
                   "adjacency_values": "factored (dinv_i * dinv_j over the binary pattern; explicit fp32 values of the 2-hop "
                                       "ring would not fit)" if big else "explicit fp32",
                   "kernel": g.plan.kernel_name,
                   "l2": "inputs larger than L2: the rank's hop pattern alone is gigabytes; K steps back to back between ONE "
                         "CUDA-event pair on the launching stream, max over ranks",
                   "precompute_s": t_pre, "wall_s_timed_region": wall, "parity": parity,
                   "note_scaling": "N > 1 runs BASELINE config 4 on a fixed graph (strong scaling among N = 2, 4, 8); the N = 1 line is the "
                                   "north-star point, a different workload" if big else None},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": balg, "kernel_ms": kern_ms,
                     "note": "rank 0: compulsory bytes of its shard's round with the reference's representation (explicit "
                             "fp32 values, SURVEY §8d) / its mean round time, against ONE GPU's measured HBM bandwidth"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": nnz_total * d / e2e_s, "unit": "edges*featdim/s", "h2d_bytes_per_step": int(g.n_local * d * 4),
                "d2h_bytes_per_step": int(g.n_local * 2 * d * 4), "ms_per_step": 1e3 * e2e_s, "steps": k_e2e,
                "api": "per rank: its rows of X from pinned host memory -> ShardedGraph.round (h2_graph_round_parts) -> its rows of Y "
                       "to pinned host memory, synchronise; host wall clock, max over ranks; bytes are rank 0's"},
    }
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()
    dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# BASELINE config 5: power-law |V| = 4M, 64M edge draws, d = 256, bf16 feature rows, rows sharded over the GPUs
# ----------------------------------------------------------------------------------------------------------------------
def main_cfg5(args):
    """2-hop precompute + fused round on the power-law graph with bf16 rows (X shards, gathered copy and Y in bf16, fp32
    accumulation).  One line: precompute seconds, round time, parity of a row sample against the oracle on the host."""
    import torch
    import torch.distributed as dist
    from h2gcn_b200 import _cabi
    from h2gcn_b200.parallel import ShardedGraph
    from h2gcn_b200.utils import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _cabi.lib()
    n, n_edges, d = CFG5_N, CFG5_E, CFG5_FEAT
    desc = (f"power-law (Chung-Lu, exponent {CFG5_GAMMA}) |V|={n} {n_edges} edge draws (symmetrised, deduplicated, no self loops) "
            f"d={d} bf16 feature rows in / out, seed 2, rows sharded over the GPUs (BASELINE config 5)")
    t0 = time.perf_counter()
    adj = synth.chung_lu_graph_device(n, n_edges, gamma=CFG5_GAMMA, seed=2, device=dev)
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    g = ShardedGraph(adj, rank, world, dev, factored=True, mode=args.mode, splits=args.splits, exchange=args.exchange,
                     explicit_vals=False, balance="auto")
    torch.cuda.synchronize()
    t_pre = time.perf_counter() - t0
    sys.stderr.write(f"[cfg5] rank {rank}: rows [{g.row_begin}, {g.row_end}) nnz1+nnz2 {g.nnz_local} max row {g.max_row_nnz} "
                     f"partition {g.balance} exchange {g.exchange} kernel {g.plan.kernel_name} precompute {t_pre:.2f} s\n")
    sys.stderr.flush()
    gen = torch.Generator(device=dev)
    gen.manual_seed(1005)
    x_all = torch.randn(n, d, device=dev, generator=gen).to(torch.bfloat16)          # the same matrix on every rank
    x_local = x_all[g.row_begin:g.row_end].contiguous()
    if world > 1 and g.exchange == "p2p":
        buf = g.input_buffer(d, torch.bfloat16)
        buf.copy_(x_local)
        x_local = buf
    x_host = x_all.float().cpu().numpy() if rank == 0 and not args.no_cpu_baseline else None
    del x_all
    torch.cuda.empty_cache()
    y = torch.empty(g.n_local, 2 * d, device=dev, dtype=torch.bfloat16)

    def step():
        g.round(x_local, y, [0, d])

    def barrier():
        if world > 1:
            dist.barrier()

    torch.cuda.synchronize()
    barrier()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _cabi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - wall0
    launches = _cabi.launch_count() - launches0
    clocks = sampler.finish()
    barrier()
    my_ms = float(ev0.elapsed_time(ev1))
    tt = torch.tensor([my_ms], device=dev, dtype=torch.float64)
    cnt = torch.tensor([g.nnz_local, g.nnz2_local], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt)
    total_ms = float(tt.item())
    nnz_total, nnz2_total = int(cnt[0].item()), int(cnt[1].item())
    ms_per_step = total_ms / args.steps
    value = nnz_total * d / (ms_per_step * 1e-3)

    # e2e: the rank's rows of X from pinned host memory, its rows of Y back
    xh = x_local.cpu().pin_memory()
    yh = torch.empty(g.n_local, 2 * d, dtype=torch.bfloat16).pin_memory()
    k_e2e = min(args.steps, 5)
    for _ in range(1):
        x_local.copy_(xh, non_blocking=True); step(); yh.copy_(y, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        x_local.copy_(xh, non_blocking=True); step(); yh.copy_(y, non_blocking=True)
        torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item()) / k_e2e

    parity = None
    if rank == 0 and x_host is not None:
        from oracle import cbind
        rng = np.random.default_rng(11)
        deg_all = g.deg2_host
        live = np.nonzero(deg_all[g.row_begin:g.row_end] > 0)[0]
        pick = np.sort(rng.choice(live, size=min(64, len(live)), replace=False)) if len(live) else np.arange(min(64, g.n_local))
        rows = pick + g.row_begin
        (rp1, c1, v1), (rp2, c2, v2) = sample_hops_cpu(adj, rows, deg_all)
        ref = cbind.fused_round(rp1, c1, v1, rp2, c2, v2, x_host, threads=host_threads())
        got = y[torch.from_numpy(pick).to(dev)].float().cpu().numpy()
        parity = parity_metrics(got, ref)
        err = np.abs(got.astype(np.float64) - ref)
        parity.update({"tolerance": 2.0 ** -8, "tolerance_note": "bf16 output: every element within one bf16 rounding (2^-8 relative) of the "
                       "oracle's fp32 result on the same bf16 inputs, + 1e-5 of the max for cancelled sums",
                       "within_one_bf16_rounding": bool((err <= 2.0 ** -8 * np.abs(ref) + 1e-5 * np.abs(ref).max()).all()),
                       "against": "oracle C port on the host (in-order fp32 sums) on the bf16-rounded inputs; 2-hop rows rebuilt from the definition with scipy",
                       "rows_checked": int(len(rows)), "entries_checked": int(len(c1) + len(c2)),
                       "deg2_of_sample_matches_gpu": bool(np.array_equal(np.diff(rp2), deg_all[rows]))})
    if rank == 0 and parity is not None and g.plan.tensor_idx:
        parity["full_shard_tensor_vs_csr"] = full_shard_check(g, torch.from_numpy(x_host).to(dev).to(torch.bfloat16), y, d)
        parity["full_shard_note"] = "both sides round to bf16 once: up to one bf16 ulp (2^-7 relative) apart"
    if rank != 0:
        dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    # compulsory bytes of rank 0's shard, bf16 rows: factored CSR (4 B per entry) + X (2 B) + Y (2 x 2 B)
    balg = 2 * (g.n_local + 1) * 8 + g.nnz_local * 4 + n * d * 2 + 2 * g.n_local * d * 2
    kern_ms = my_ms / args.steps
    achieved = balg / (kern_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "edges*featdim/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16 rows, f32 accumulation" + ("" if not g.plan.tensor_idx else " / i8 digits"), "data": "synthetic",
        "config": {"workload": desc, "exchange": g.exchange, "n_vertices": n, "n_edge_draws": n_edges, "nnz1": g.nnz1_global,
                   "nnz2": nnz2_total, "nnz_total": nnz_total, "nnz_local_rank0": g.nnz_local, "rows_rank0": g.n_local,
                   "max_degree": g.max_deg1, "max_degree_2hop": g.max_deg2, "zero_degree_rows": g.zero_deg1,
                   "row_partition": g.balance, "adjacency_values": "factored (dinv_i * dinv_j over the binary pattern)",
                   "kernel": g.plan.kernel_name, "graph_generation_s": t_gen,
                   "precompute_s": t_pre, "precompute_what": "per rank: 2-hop degree count (equal-rows split, all-gathered), hop2 fill + "
                   "SYM normalisation of its rows, round plan; host wall clock",
                   "precompute_nnz2_per_s": nnz2_total / t_pre, "wall_s_timed_region": wall, "parity": parity,
                   "l2": "inputs larger than L2 (the rank's hop pattern alone is gigabytes); K steps between ONE CUDA-event pair, max over ranks"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": balg, "kernel_ms": kern_ms,
                     "note": "rank 0: compulsory bytes of its shard's round (factored CSR indices, bf16 X and Y) / its mean round time, "
                             "against ONE GPU's measured HBM bandwidth"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": nnz_total * d / e2e_s, "unit": "edges*featdim/s", "h2d_bytes_per_step": int(g.n_local * d * 2),
                "d2h_bytes_per_step": int(g.n_local * 2 * d * 2), "ms_per_step": 1e3 * e2e_s, "steps": k_e2e,
                "api": "per rank: its bf16 rows of X from pinned host memory -> ShardedGraph.round (h2_graph_round_parts_ex) -> its bf16 rows "
                       "of Y to pinned host memory, synchronise; host wall clock, max over ranks"},
    }
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
