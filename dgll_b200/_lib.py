"""ctypes binding of the C-ABI library (``include/dgll_b200.h``).

There is no CPU fallback: if ``libdgll_b200.so`` is missing or a call fails the
caller gets an exception.  Building is explicit (``python -m dgll_b200.build`` or
``__graft_entry__.build()``); importing this module only loads.
"""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint64, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdgll_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, 1, 2, 3
SUM, MEAN, MAX = 0, 1, 2
F32, BF16 = 0, 1
EPI_RELU, EPI_ELU = 1, 2
GAT_SOFTMAX, GAT_EXP_NEG = 0, 1

_P = c_void_p
_I = c_int
_L = c_int64

# name -> (restype, argtypes); mirrors include/dgll_b200.h declaration by declaration
SIGNATURES = {
    "dgllb_version": (_I, []),
    "dgllb_last_error": (c_char_p, []),
    "dgllb_device_info": (_I, [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int64)]),
    "dgllb_launch_count": (_L, []),
    "dgllb_set_option": (_I, [c_char_p, c_char_p]),
    "dgllb_get_option": (_I, [c_char_p, POINTER(c_int)]),
    "dgllb_csr_plan_create": (_I, [_P, _I, _L, _I, _P, POINTER(c_void_p)]),
    "dgllb_csr_plan_info": (_I, [_P, POINTER(c_int64), POINTER(c_int64), POINTER(c_int)]),
    "dgllb_csr_plan_destroy": (None, [_P]),
    "dgllb_spmm_csr": (_I, [_P, _I, _P, _P, _P, _I, _L, _P, _L, _L, _L, _L, _I, _I, _P, _P, _L, _P, _I, _P, _P, _P]),
    "dgllb_spmm_csr_sharded": (_I, [_P, _I, _P, _P, _P, _I, _L, _L, _I, _P, _L, _L, _I, _I, _P]),
    "dgllb_sddmm_csr": (_I, [_P, _I, _P, _P, _L, _P, _L, _P, _L, _I, _P]),
    "dgllb_spmm_max_backward": (_I, [_P, _P, _P, _L, _P, _L, _L, _I, _P]),
    "dgllb_csr_transpose": (_I, [_P, _I, _P, _P, _L, _L, _L, _P, _P, _P, _P, _P]),
    "dgllb_gather_rows": (_I, [_P, _L, _P, _I, _P, _L, _L, _L, _P]),
    "dgllb_gather_rows_cached": (_I, [_P, _L, _P, _L, _P, _P, _P, _P, _P, _L, _L, _L, _P, _P]),
    "dgllb_gather_rows_sharded": (_I, [_P, _I, _L, _L, _P, _I, _P, _L, _L, _L, _P]),
    "dgllb_ipc_export": (_I, [_P, _P, POINTER(c_int64)]),
    "dgllb_ipc_import": (_I, [_P, _L, POINTER(c_void_p)]),
    "dgllb_ipc_release": (_I, [_P, _L]),
    "dgllb_gemm_f32": (_I, [_P, _L, _I, _P, _L, _I, _P, _L, _L, _L, _L, _P, _I, _I, _I, _P]),
    "dgllb_gat_forward": (_I, [_P, _I, _P, _P, _L, _P, _P, _L, _P, _L, _P, _P, _L, _L, _I, _I, c_float, _I, _I, c_float,
                               c_uint64, _P, _P]),
    "dgllb_gat_dropout_mask": (_I, [c_uint64, _L, _I, c_float, _P, _P]),
    "dgllb_gat_backward": (_I, [_P, _I, _P, _P, _P, _P, _P, _L, _P, _P, _L, _P, _L, _P, _P, _P, _L, _P, _L,
                                _P, _P, _L, _P, _L, _L, _I, _I, c_float, _I, c_float, c_uint64, _P, _P, _P]),
    "dgllb_binarize_pack": (_I, [_P, _L, _P, _L, _L, _I, _P]),
    "dgllb_bin_spmm_csr": (_I, [_P, _I, _P, _P, _L, _P, _L, _L, _I, _I, _P, _P]),
    "dgllb_sample_neighbors": (_I, [_P, _I, _P, _P, _I, _L, _I, c_uint64, _P, _P, _P]),
    "dgllb_build_block": (_I, [_P, _L, _P, _P, _L, _P, _P, _P, _P]),
    "dgllb_sample_neighbors_cap": (_I, [_P, _I, _P, _P, _I, _L, _I, c_uint64, _P, _P, _P, _P]),
    "dgllb_build_block_cap": (_I, [_P, _L, _P, _P, _L, _I, _P, _P, _P, _P]),
    "dgllb_csr_slice_rows_ptr": (_I, [_P, _I, _P, _L, _P, _P]),
    "dgllb_csr_slice_rows_fill": (_I, [_P, _I, _P, _P, _P, _L, _P, _P, _P, _P]),
    "dgllb_col_sqsum": (_I, [_P, _P, _L, _L, _I, _P, _P, _P, _P]),
    "dgllb_weighted_choice": (_I, [_P, _P, _L, _I, c_uint64, _P, _P, _P, _P]),
    "dgllb_importance_scale": (_I, [_P, _P, _P, _I, _L, _I, _P, _P]),
    "dgllb_scatter_pos": (_I, [_P, _P, _P, _L, _I, _P]),
    "dgllb_csr_select_cols": (_I, [_P, _P, _P, _L, _L, _P, _P, _P, _P, _P, _P]),
    "launch_gcn_fused_kernel": (None, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I]),
    "launch_gcn_fused_kernel_backward_optimized": (None, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I]),
    "dgllb_gcn_fused_forward": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "dgllb_gcn_fused_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
}

_lib = None


class DgllB200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def lib():
    """The loaded library.  Raises (never falls back) when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "dgll_b200: %s is missing — build it with `python -m dgll_b200.build` "
                "(there is no CPU fallback for the aggregation path)" % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def last_error():
    return lib().dgllb_last_error().decode(errors="replace")


def check(rc, what=""):
    if rc != OK:
        kind = {ERR_INVALID: "invalid argument", ERR_CUDA: "CUDA error", ERR_UNSUPPORTED: "unsupported"}.get(rc, "error %d" % rc)
        raise DgllB200Error("%s: %s: %s" % (what or "dgll_b200", kind, last_error()))


def launch_count():
    return int(lib().dgllb_launch_count())


def set_option(name, value):
    """Tuning option of the library (see ``dgllb_set_option`` in include/dgll_b200.h); ``None`` = default."""
    v = None if value is None else str(value).encode()
    check(lib().dgllb_set_option(name.encode(), v), "set_option")


def get_option(name):
    out = c_int()
    check(lib().dgllb_get_option(name.encode(), ctypes.byref(out)), "get_option")
    return out.value
