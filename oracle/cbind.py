"""ctypes binding for oracle/h2gcn_oracle.c — TEST INFRASTRUCTURE (see h2gcn_oracle.py header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "h2gcn_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_hop2_csr.restype = ctypes.c_int64
        _lib.oracle_max_threads.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def max_threads():
    return int(lib().oracle_max_threads())


def spmm_coo(rows, cols, vals, b, n_rows):
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.empty((n_rows, b.shape[1]), dtype=np.float32)
    I = ctypes.c_int64
    lib().oracle_spmm_coo_f32(I(len(vals)), _p(rows), _p(cols), _p(vals), _p(b), I(b.shape[1]), I(b.shape[1]),
                              _p(out), I(b.shape[1]), I(n_rows))
    return out


def fused_round(rp1, c1, v1, rp2, c2, v2, x, y=None, off1=0, off2=None, threads=0):
    """Y[:, off1:off1+d] = A1 X ; Y[:, off2:off2+d] = A2 X with all host threads (threads=0) or `threads`."""
    n, d = len(rp1) - 1, x.shape[1]   # rows of the (possibly row-sharded) adjacency; x has one row per COLUMN
    if off2 is None:
        off2 = d
    if y is None:
        y = np.empty((n, 2 * d), dtype=np.float32)
    I = ctypes.c_int64
    a = [np.ascontiguousarray(t, dtype=np.int32) for t in (rp1, c1, rp2, c2)]
    f = [np.ascontiguousarray(t, dtype=np.float32) for t in (v1, v2)]
    assert x.flags.c_contiguous and y.flags.c_contiguous and x.dtype == np.float32 and y.dtype == np.float32
    lib().oracle_fused_round_f32(I(n), _p(a[0]), _p(a[1]), _p(f[0]), _p(a[2]), _p(a[3]), _p(f[1]),
                                 _p(x), I(x.shape[1]), I(d), _p(y), I(y.shape[1]), I(off1), I(off2),
                                 ctypes.c_int(threads))
    return y


def hop2_csr(rowptr, col, threads=0):
    """Exact-distance-2 pattern of a diagonal-free, sorted CSR adjacency: (rowptr2 int64, col2 int32)."""
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    n = len(rowptr) - 1
    rp2 = np.zeros(n + 1, dtype=np.int64)
    nnz2 = lib().oracle_hop2_csr(ctypes.c_int32(n), _p(rowptr), _p(col), _p(rp2), None, ctypes.c_int(threads))
    col2 = np.empty(max(int(nnz2), 1), dtype=np.int32)
    lib().oracle_hop2_csr(ctypes.c_int32(n), _p(rowptr), _p(col), _p(rp2), _p(col2), ctypes.c_int(threads))
    return rp2, col2[:int(nnz2)]
