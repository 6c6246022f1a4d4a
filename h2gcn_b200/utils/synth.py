"""Seeded synthetic graphs for tests and bench.py (SURVEY.md §8d "Concrete synthetic inputs").  Host numpy only.
All graphs are undirected, symmetrised, deduplicated, without self loops; returned as scipy CSR float32 (ones)."""
import numpy as np
import scipy.sparse as sp


def _sym_csr(n, src, dst):
    keep = src != dst
    src, dst = src[keep], dst[keep]
    rows = np.concatenate([src, dst])
    cols = np.concatenate([dst, src])
    m = sp.csr_matrix((np.ones(len(rows), dtype=np.float32), (rows, cols)), shape=(n, n))
    m.sum_duplicates()
    m.data[:] = 1.0
    m.sort_indices()
    return m


def uniform_graph(n, n_edges, seed=0):
    """`n_edges` DISTINCT undirected edges drawn uniformly (the north-star target graph: n=10_000, n_edges=200_000)."""
    rng = np.random.default_rng(seed)
    have = np.empty(0, dtype=np.int64)
    while len(have) < n_edges:
        need = n_edges - len(have)
        a = rng.integers(0, n, size=int(need * 1.1) + 16)
        b = rng.integers(0, n, size=len(a))
        ok = a != b
        lo, hi = np.minimum(a[ok], b[ok]), np.maximum(a[ok], b[ok])
        have = np.unique(np.concatenate([have, lo * n + hi]))
    if len(have) > n_edges:
        have = rng.permutation(have)[:n_edges]
    return _sym_csr(n, have // n, have % n)


def preferential_attachment(n, m, seed=0):
    """Barabasi-Albert style: each new vertex attaches to m earlier vertices chosen by degree (syn-cora / syn-products
    proxies: n=1490,m=2 / n=10_000,m=6)."""
    rng = np.random.default_rng(seed)
    targets = list(range(m))
    repeated = []
    src, dst = [], []
    for v in range(m, n):
        for t in set(targets):
            src.append(v)
            dst.append(t)
        repeated.extend(targets)
        repeated.extend([v] * m)
        targets = [repeated[i] for i in rng.integers(0, len(repeated), size=m)]
    return _sym_csr(n, np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64))


def rmat_graph(scale_n, n_edges, a=0.57, b=0.19, c=0.19, seed=1):
    """R-MAT edge draws (Graph500 parameters) on the next power of two >= scale_n, trimmed to [0, scale_n)."""
    rng = np.random.default_rng(seed)
    bits = int(np.ceil(np.log2(max(2, scale_n))))
    src = np.zeros(n_edges, dtype=np.int64)
    dst = np.zeros(n_edges, dtype=np.int64)
    for _ in range(bits):
        r = rng.random(n_edges)
        src = (src << 1) | (r >= a + b)
        dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c))
    keep = (src < scale_n) & (dst < scale_n)
    return _sym_csr(scale_n, src[keep], dst[keep])


def chung_lu_graph_device(n, n_edges, gamma=2.5, seed=2, device="cuda"):
    """Power-law graph (Chung-Lu: both endpoints of every edge draw are sampled with probability ~ w_i,
    w_i = (i + 64)^(-1/(gamma-1)), so expected degrees follow a power law with exponent gamma; vertex 0 is the largest hub),
    BASELINE config 5's "synthetic power-law |V|=4M |E|=64M".  Built with torch ON THE DEVICE (64 M draws through a
    4 M-entry CDF take minutes in numpy and well under a second here), returned as host scipy CSR like the others.
    Deterministic for a given (n, n_edges, gamma, seed) on the same GPU architecture (Philox)."""
    import torch
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed)
    w = (torch.arange(n, dtype=torch.float64, device=dev) + 64.0) ** (-1.0 / (gamma - 1.0))
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    src = torch.searchsorted(cdf, torch.rand(n_edges, dtype=torch.float64, device=dev, generator=gen), right=True).clamp_(max=n - 1)
    dst = torch.searchsorted(cdf, torch.rand(n_edges, dtype=torch.float64, device=dev, generator=gen), right=True).clamp_(max=n - 1)
    keep = src != dst
    lo, hi = torch.minimum(src[keep], dst[keep]), torch.maximum(src[keep], dst[keep])
    und = torch.unique(lo * n + hi)                                   # distinct undirected edges
    lo, hi = und // n, und % n
    key = torch.sort(torch.cat([lo * n + hi, hi * n + lo]))[0]        # both directions, (row, col) order
    rows, cols = key // n, (key % n).to(torch.int32)
    indptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    indptr[1:] = torch.cumsum(torch.bincount(rows, minlength=n), 0)
    indptr_h, cols_h = indptr.cpu().numpy(), cols.cpu().numpy()
    return sp.csr_matrix((np.ones(len(cols_h), dtype=np.float32), cols_h, indptr_h), shape=(n, n))


def features(n, d, seed=0):
    return np.random.default_rng(seed + 1000).standard_normal((n, d)).astype(np.float32)
