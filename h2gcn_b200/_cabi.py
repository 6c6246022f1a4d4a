"""ctypes binding of include/h2gcn_b200.h — the only door from Python into the CUDA library.

There is NO CPU fallback: if the shared library is missing `lib()` raises, and every wrapper that computes requires
CUDA tensors.  PyTorch tensors are used purely as device buffers (`.data_ptr()`) and for the current stream.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_lib", "libh2gcn_b200.so")
MAX_HOPS = 8

H2_OK, H2_ERR_INVALID, H2_ERR_ALIGN, H2_ERR_WORKSPACE, H2_ERR_CUDA, H2_ERR_UNSUPPORTED, H2_ERR_INDEX = range(7)

c_i32, c_i64, c_vp, c_sz = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p, ctypes.c_size_t

# `splits` codes of the tensor-core path (include/h2gcn_b200.h): 2 / 3 bf16 pieces, or int8 digits
H2_SPLITS_I8X2, H2_SPLITS_I8X3 = 18, 19
H2_F32, H2_BF16 = 0, 1     # element type codes of the feature matrices of a round
SPLITS = {2: 2, 3: 3, "bf16x2": 2, "bf16x3": 3, "i8x2": H2_SPLITS_I8X2, "i8x3": H2_SPLITS_I8X3,
          H2_SPLITS_I8X2: H2_SPLITS_I8X2, H2_SPLITS_I8X3: H2_SPLITS_I8X3}
SPLITS_NAME = {2: "2 bf16 pieces", 3: "3 bf16 pieces", H2_SPLITS_I8X2: "2 int8 digits + block exponents",
               H2_SPLITS_I8X3: "3 int8 digits + block exponents"}


# Default arithmetic of the tensor-core path: THREE int8 digits with block exponents — 24 significant bits of X', exact
# int32 accumulation, one fp32 rounding in the epilogue: as accurate as the fp32 CSR path (3e-7 of max-abs against the
# fp32 oracle, the reference's own fp32-vs-fp64 error is 2-5e-7), so the model API (forward AND backward) computes at the
# reference's precision.  "i8x2" (16 bits, ~8e-6 norm-wise, rows far below the global maximum lose relative precision) is
# an opt-in: H2GCN_SPLITS=2|3|i8x2|i8x3 or HopPlan(splits=...).
DEFAULT_SPLITS = os.environ.get("H2GCN_SPLITS", "i8x3")


def splits_code(splits=None):
    """Arithmetic of the tensor-core path: None (default) | 2 | 3 | "bf16x2" | "bf16x3" | "i8x2" | "i8x3" -> C-ABI code."""
    if splits is None:
        splits = DEFAULT_SPLITS
    if isinstance(splits, str) and splits.isdigit():
        splits = int(splits)
    try:
        return SPLITS[splits]
    except (KeyError, TypeError):
        raise ValueError(f"unknown splits {splits!r}: one of 2, 3, 'bf16x2', 'bf16x3', 'i8x2', 'i8x3'")


class HopDesc(ctypes.Structure):
    """h2_hop_t"""
    _fields_ = [("rowptr", c_vp), ("col", c_vp), ("val", c_vp), ("dinv", c_vp), ("dinv_row", c_vp),
                ("out_col_off", c_i64), ("in_col_off", c_i64)]


# name -> (restype, argtypes); must list EVERY symbol include/h2gcn_b200.h declares (tests check this).
PROTOTYPES = {
    "h2_abi_version": (ctypes.c_int, []),
    "h2_last_error": (ctypes.c_char_p, []),
    "h2_launch_count": (c_i64, []),
    "h2_remove_eye_count": (ctypes.c_int, [c_i32, c_vp, c_vp, c_vp, c_vp]),
    "h2_remove_eye_fill": (ctypes.c_int, [c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "h2_scan_workspace_bytes": (c_sz, [c_i64]),
    "h2_exclusive_scan_i64": (ctypes.c_int, [c_i64, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "h2_hop2_count": (ctypes.c_int, [c_i32, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp]),
    "h2_hop2_fill": (ctypes.c_int, [c_i32, c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "h2_sym_normalize": (ctypes.c_int, [c_i32, c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "h2_rw_normalize": (ctypes.c_int, [c_i32, c_vp, c_vp, c_vp]),
    "h2_validate_csr": (ctypes.c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "h2_plan_host_bytes": (c_sz, []),
    "h2_plan_dev_bytes": (c_sz, [c_i32, c_i32]),
    "h2_plan_workspace_bytes": (c_sz, [c_i32, c_i32]),
    "h2_plan_build": (ctypes.c_int, [c_i32, c_i32, ctypes.POINTER(HopDesc), c_vp, c_vp, c_vp, c_sz, c_vp]),
    "h2_fused_hops_spmm_f32": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, ctypes.POINTER(HopDesc), c_i32, c_vp, c_i64,
                                              c_vp, c_i64, c_vp]),
    "h2_fused_hops_spmm_ex": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, ctypes.POINTER(HopDesc), c_i32, c_vp, c_i64, c_i32,
                                             c_vp, c_i64, c_i32, c_vp]),
    "h2_bm_host_bytes": (c_sz, []),
    "h2_bm_index_bytes": (c_sz, [c_i32, c_i32]),
    "h2_bm_count": (ctypes.c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_sz, ctypes.POINTER(c_i64), c_vp]),
    "h2_bm_plan_dev_bytes": (c_sz, [c_i32, c_i32, c_i64]),
    "h2_bm_fill": (ctypes.c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "h2_bm_fill_order": (ctypes.c_int, [c_i32, c_i32, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_sz, c_i32, c_vp]),
    "h2_bm_max_width": (c_i32, [c_i32]),
    "h2_bm_xpack_bytes": (c_sz, [c_i32, c_i32, c_i32]),
    "h2_bm_partial_bytes": (c_sz, [c_vp, c_i32, c_i32]),
    "h2_bm_pack_x_f32": (ctypes.c_int, [c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "h2_bm_pack_x_f32_armed": (ctypes.c_int, [c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "h2_bm_spmm_f32": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_sz, c_vp]),
    "h2_sparse_dense_f32": (ctypes.c_int, [c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_i64, c_vp]),
    "h2_dense_f32": (ctypes.c_int, [c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_i32, c_vp, c_i64, c_i64, c_vp]),
    "h2_dense_tc_f32": (ctypes.c_int, [c_i32, c_i32, c_i32, c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, c_vp, c_i32, c_vp, c_i64, c_i64, c_vp]),
    "h2_relu_slice_f32": (ctypes.c_int, [c_i32, c_i32, c_vp, c_i64, c_vp, c_i64, c_i32, c_vp]),
    "h2_graph_create": (ctypes.c_int, [c_i32, c_i32, c_i32, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp),
                                       ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), c_i32, c_i32, c_i32, c_i32,
                                       ctypes.POINTER(c_vp)]),
    "h2_graph_create_device": (ctypes.c_int, [c_i32, c_i32, c_i32, ctypes.POINTER(HopDesc), ctypes.POINTER(c_i64), c_i32, c_i32,
                                              c_i32, ctypes.POINTER(c_vp)]),
    "h2_graph_workspace_bytes": (c_sz, [c_vp, c_i32]),
    "h2_graph_bind_workspace": (ctypes.c_int, [c_vp, c_i32, c_vp, c_sz]),
    "h2_graph_reserve": (ctypes.c_int, [c_vp, c_i32]),
    "h2_graph_formats": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i32)]),
    "h2_graph_round_host": (ctypes.c_int, [c_vp, c_i32, c_vp, c_vp, c_vp]),
    "h2_graph_round": (ctypes.c_int, [c_vp, c_i32, c_vp, c_i64, c_vp, c_i64, ctypes.POINTER(c_i64), c_vp]),
    "h2_graph_round_ex": (ctypes.c_int, [c_vp, c_i32, c_vp, c_i64, c_i32, c_vp, c_i64, c_i32, ctypes.POINTER(c_i64), c_vp]),
    "h2_graph_round_multi": (ctypes.c_int, [c_vp, c_i32, c_vp, c_i64, ctypes.POINTER(c_i64), c_vp, c_i64, ctypes.POINTER(c_i64), c_vp]),
    "h2_graph_round_parts": (ctypes.c_int, [c_vp, c_i32, c_i32, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), c_i64, c_vp, c_i64,
                                            c_vp, c_i64, ctypes.POINTER(c_i64), c_vp]),
    "h2_graph_round_parts_ex": (ctypes.c_int, [c_vp, c_i32, c_i32, ctypes.POINTER(c_vp), ctypes.POINTER(c_i64), c_i64, c_i32, c_vp, c_i64,
                                               c_vp, c_i64, c_i32, ctypes.POINTER(c_i64), c_vp]),
    "h2_sum_slices_f32": (ctypes.c_int, [c_i32, c_i32, c_i32, c_vp, c_i64, c_vp, c_i64, c_i32, c_vp, c_i64, c_vp]),
    "h2_graph_destroy": (ctypes.c_int, [c_vp]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises — loudly — if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise RuntimeError(
                f"h2gcn_b200: CUDA library {SO_PATH} is missing — run `python -m h2gcn_b200.build` "
                "(there is no CPU fallback)")
        handle = ctypes.CDLL(SO_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.h2_abi_version() != 1:
            raise RuntimeError("h2gcn_b200: ABI version mismatch between _cabi.py and the shared library")
        _lib = handle
    return _lib


def check(status):
    """Map a C status to the exception the reference raises at that seam (SURVEY.md §8b error convention)."""
    if status == H2_OK:
        return
    msg = lib().h2_last_error().decode("utf-8", "replace")
    if status == H2_ERR_CUDA:
        raise RuntimeError(f"h2gcn_b200: {msg}")
    raise ValueError(f"h2gcn_b200 (status {status}): {msg}")


def launch_count():
    return int(lib().h2_launch_count())


def ptr(t):
    """Device (or pinned-host) pointer of a torch tensor / None."""
    return None if t is None else t.data_ptr()


def stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("h2gcn_b200: expected a CUDA tensor — the hot path has no CPU fallback")
