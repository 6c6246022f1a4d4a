"""1-D row (vertex) sharding of the aggregation round over the GPUs of one box.

One process per GPU.  Rank q owns a contiguous vertex range chosen so that the stored entries of [A1; A2] are balanced
(prefix-sum split, not n/P).  It holds its rows of both hop adjacencies (global column ids), its row slice of the
output, and a full-height gathered copy of the round's input.  The only exchange step is the all-gather of the input
rows at the hop boundary; outputs are row-local, so nothing else is communicated.

The partition / gather logic is backend-agnostic (`torch.distributed`): NCCL on the GPUs, gloo in the CPU tests.
"""
import numpy as np
import torch
import torch.distributed as dist


def balanced_row_partition(weights, world):
    """Contiguous split of rows into `world` ranges with (nearly) equal total weight.
    Returns int64 boundaries b[0..world], b[0]=0, b[world]=len(weights); boundary q is the first row whose prefix
    weight reaches q/world of the total (ties and empty tails handled so that every range is valid)."""
    w = np.asarray(weights, dtype=np.int64)
    n = len(w)
    csum = np.concatenate([[0], np.cumsum(w)])
    total = int(csum[-1])
    bounds = np.zeros(world + 1, dtype=np.int64)
    bounds[world] = n
    for q in range(1, world):
        if total == 0:
            bounds[q] = (n * q) // world
        else:
            bounds[q] = int(np.searchsorted(csum, (total * q) / world, side="left"))
        bounds[q] = min(max(bounds[q], bounds[q - 1]), n)
    return bounds


def all_gather_rows(x_local, x_full, bounds, group=None):
    """x_full[bounds[q]:bounds[q+1]] <- rank q's x_local (uneven shard heights allowed).  In-place into x_full."""
    world = len(bounds) - 1
    if world == 1:
        if x_full.data_ptr() != x_local.data_ptr():
            x_full.copy_(x_local)
        return x_full
    _uneven_all_gather(x_local, x_full, bounds, group)
    return x_full


def _uneven_all_gather(local, full, bounds, group=None):
    """Equal shard heights -> ONE all-gather straight into `full`.  Uneven heights -> one all-gather of shards padded to
    the tallest one, then a compaction copy (ncclAllGather and gloo's allgather both need equal sizes).
    `full[bounds[q]:bounds[q+1]]` receives rank q's rows."""
    world = len(bounds) - 1
    sizes = np.diff(np.asarray(bounds))
    if np.all(sizes == sizes[0]):
        dist.all_gather_into_tensor(full, local.contiguous(), group=group)
        return
    hmax = int(sizes.max())
    tail = tuple(full.shape[1:])
    padded = torch.zeros((hmax,) + tail, dtype=full.dtype, device=full.device)
    padded[:local.shape[0]] = local
    gathered = torch.empty((world * hmax,) + tail, dtype=full.dtype, device=full.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    for q in range(world):
        if sizes[q]:
            full[int(bounds[q]):int(bounds[q + 1])] = gathered[q * hmax:q * hmax + int(sizes[q])]


def partition_rows(weights, world, tolerance=1.05):
    """Row ranges for `world` ranks: the EQUAL split when it is balanced within `tolerance` of the ideal load (then the
    hop-boundary exchange is a single ncclAllGather), else the nnz-balanced prefix-sum split."""
    w = np.asarray(weights, dtype=np.int64)
    n = len(w)
    if n % world == 0 and n > 0:
        eq = np.arange(world + 1, dtype=np.int64) * (n // world)
        loads = np.add.reduceat(w, eq[:-1]) if n else np.zeros(world)
        if loads.max() <= tolerance * max(1.0, w.sum() / world):
            return eq
    return balanced_row_partition(w, world)


def all_gather_counts(local_counts, bounds, group=None):
    """Concatenate per-rank int64 vectors of (uneven) known lengths into the global vector on every rank."""
    world = len(bounds) - 1
    if world == 1:
        return local_counts
    out = torch.empty(int(bounds[-1]), dtype=local_counts.dtype, device=local_counts.device)
    _uneven_all_gather(local_counts, out, bounds, group)
    return out


class ShardedGraph:
    """Row shard of the normalised hop adjacencies [A1, A2] of one graph + its fused-round plan.

    Set-up (once per graph): the full A (4·nnz1 bytes) is replicated on every rank; 2-hop degrees are counted on an
    equal-rows split and all-gathered (they are also the global degree vector D2 the normalisation needs), the
    nnz-balanced boundaries are derived from them, then every rank fills and normalises its own rows."""

    def __init__(self, adj_scipy, rank, world, device, factored=False, group=None, mode="auto", splits=None, exchange="auto",
                 explicit_vals=True, balance="auto"):
        """explicit_vals=False: the fp32 values of the hop adjacencies are never materialised (factored CSR: the kernels
        rebuild dinv_i * dinv_j; implies factored=True) — BASELINE config 4's 2-hop ring has ~2e10 entries.
        balance: how the contiguous row ranges are cut — "nnz" (stored entries of [A1; A2]: what the CSR gather costs),
        "rows" (equal row counts: what the tile-bitmap path costs, whose work is rows x columns whatever the entries),
        "auto" = "rows" when the 2-hop pattern is dense enough for the tensor-core format, else "nnz"."""
        from . import ops
        self.rank, self.world, self.device, self.group = rank, world, device, group
        if not explicit_vals:
            factored = True
        a = adj_scipy.tocsr()
        a.sort_indices()
        n = a.shape[0]
        self.n = n
        rp = torch.from_numpy(a.indptr.astype(np.int64)).to(device)
        col = torch.from_numpy(a.indices.astype(np.int32)).to(device)
        deg1 = rp[1:] - rp[:-1]
        self.nnz1_global = int(col.numel())
        # pass 1: 2-hop degree of an equal-rows slice, gathered to the global D2
        eq = np.array([(n * q) // world for q in range(world + 1)], dtype=np.int64)
        lo, hi = int(eq[rank]), int(eq[rank + 1])
        cnt = torch.empty(hi - lo, dtype=torch.int64, device=device)
        ops.check(ops.lib().h2_hop2_count(n, ops.ptr(rp), ops.ptr(col), lo, hi, ops.ptr(cnt), ops.stream_ptr()))
        deg2 = all_gather_counts(cnt, eq, group)
        self.deg2_host = deg2.cpu().numpy()
        d1 = deg1.cpu().numpy()
        self.max_deg1, self.max_deg2 = int(d1.max(initial=0)), int(self.deg2_host.max(initial=0))
        self.zero_deg1, self.zero_deg2 = int((d1 == 0).sum()), int((self.deg2_host == 0).sum())
        # contiguous row partition
        if balance == "auto":
            dens2 = float(self.deg2_host.sum()) / max(1.0, float(n) * n)
            balance = "rows" if (mode != "csr" and dens2 >= 0.01) else "nnz"
        if balance == "rows":
            self.bounds = eq.copy()
        else:
            self.bounds = partition_rows((deg1 + deg2).cpu().numpy(), world)
        self.balance = balance
        self.row_begin, self.row_end = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.n_local = self.row_end - self.row_begin
        b, e = self.row_begin, self.row_end
        # pass 2: fill + normalise the local rows
        rp2 = ops.exclusive_scan(deg2[b:e].contiguous())
        col2 = torch.empty(int(rp2[-1].item()) if self.n_local else 0, dtype=torch.int32, device=device)
        ops.check(ops.lib().h2_hop2_fill(n, ops.ptr(rp), ops.ptr(col), b, e, ops.ptr(rp2), ops.ptr(col2), ops.stream_ptr()))
        rp1 = (rp[b:e + 1] - rp[b]).contiguous()
        col1 = col[int(rp[b].item()):int(rp[e].item())].contiguous()
        v1, _, d1v = ops.sym_normalize(rp1, col1, n_cols=n, row_begin=b, deg_all=deg1.contiguous(), want_val=explicit_vals)
        v2, _, d2v = ops.sym_normalize(rp2, col2, n_cols=n, row_begin=b, deg_all=deg2.contiguous(), want_val=explicit_vals)
        self.hops = [ops.SparseTensor(rp1, col1, v1, (self.n_local, n), row_begin=b, dinv=d1v),
                     ops.SparseTensor(rp2, col2, v2, (self.n_local, n), row_begin=b, dinv=d2v)]
        self.plan = ops.HopPlan(self.hops, factored=factored, mode=mode, splits=splits)
        self.nnz2_local = int(col2.numel())
        self.nnz_local = int(col1.numel()) + self.nnz2_local
        self.max_row_nnz = int(max(int(deg1[b:e].max().item()) if self.n_local else 0,
                                   int(deg2[b:e].max().item()) if self.n_local else 0))
        self._x_full = {}
        # hop-boundary exchange: "p2p" = the ranks' input shards live in symmetric memory and the first kernel of the
        # round loads them over NVLink (all-gather fused into the pack kernel); "nccl" = ncclAllGather, then the round
        self.exchange = "nccl"
        self._sym = {}
        if world > 1 and exchange in ("auto", "p2p"):
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self._symm_mem = symm_mem
                self.input_buffer(4)            # probe: allocation + rendezvous must work on this box
                self._sym.clear()
                self.exchange = "p2p"
            except Exception as e:              # pragma: no cover - depends on the box
                if exchange == "p2p":
                    raise
                self._exchange_error = repr(e)

    def input_buffer(self, d, dtype=torch.float32):
        """The rank's input shard [n_local, d] for width d (fp32 or bf16 rows), allocated in symmetric memory (peers map
        it over NVLink).  Writing the round input here makes `round` zero-copy; any other tensor is copied in first."""
        buf = self._sym.get((d, dtype))
        if buf is None:
            rows = int(np.diff(self.bounds).max())
            t = self._symm_mem.empty((rows, d), dtype=dtype, device=self.device)
            hdl = self._symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            buf = self._sym[(d, dtype)] = (t, hdl, [int(p) for p in hdl.buffer_ptrs])
        return buf[0][:self.n_local]

    def gathered_input(self, x_local):
        """All-gather of the round input at the hop boundary (the one collective of the path)."""
        if self.world == 1:
            return x_local
        key = (x_local.shape[1], x_local.dtype)
        buf = self._x_full.get(key)
        if buf is None:
            buf = self._x_full[key] = torch.empty(self.n, x_local.shape[1], dtype=x_local.dtype, device=self.device)
        return all_gather_rows(x_local, buf, self.bounds, self.group)

    def round(self, x_local, y_local, offsets, d=None):
        """y_local[:, offsets[h] : +d] = A_h[local rows, :] @ X  with X = all-gather of the ranks' x_local."""
        d = x_local.shape[1] if d is None else d
        if self.world == 1 or self.exchange != "p2p":
            x = self.gathered_input(x_local)
            return self.plan.run(x, y_local, offsets, d=d)
        mine = self.input_buffer(x_local.shape[1], x_local.dtype)
        if mine.data_ptr() != x_local.data_ptr():
            mine.copy_(x_local)
        t, hdl, ptrs = self._sym[(x_local.shape[1], x_local.dtype)]
        key = ("full", x_local.shape[1], x_local.dtype)
        xf = self._x_full.get(key)
        if xf is None:
            xf = self._x_full[key] = torch.empty(self.n, x_local.shape[1], dtype=x_local.dtype, device=self.device)
        hdl.barrier(channel=0)          # every rank's shard is written
        self.plan.run_parts(ptrs, self.bounds, t.stride(0), xf, y_local, offsets, d)
        hdl.barrier(channel=1)          # every rank has finished reading the shards
        return y_local

    def hops_host(self):
        return [(h.rowptr.cpu().numpy(), h.col.cpu().numpy(), h.values.cpu().numpy()) for h in self.hops]
