"""ctypes binding of ``libdsw.so`` — the only route from Python into the CUDA kernels.

There is no CPU fallback: if the library is missing, or a call is made with a non-CUDA tensor,
the product raises.  (``include/dsw.h`` is the authoritative declaration; the argtypes below
mirror it one to one and ``tests/test_abi_and_ddp.py`` checks that every declared symbol is exported.)
"""
from __future__ import annotations

import ctypes as C
import os
import re

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSW_LIB_PATH") or os.path.join(PKG_DIR, "libdsw.so")  # (override: A/B of two builds)
HEADER_PATH = os.path.join(os.path.dirname(PKG_DIR), "include", "dsw.h")

_i32, _i64, _f32 = C.c_int32, C.c_int64, C.c_float
_ptr, _sz = C.c_void_p, C.c_size_t

# name -> (restype, argtypes); order and meaning as in include/dsw.h
SIGNATURES = {
    "dsw_version": (C.c_int, []),
    "dsw_strerror": (C.c_char_p, [C.c_int]),
    "dsw_last_cuda_error": (C.c_int, []),
    "dsw_last_cuda_error_string": (C.c_char_p, []),
    "dsw_device_count": (C.c_int, []),
    "dsw_plan_create": (C.c_int, [_i32, _i32, _i64, _ptr, _ptr, _ptr, _ptr, C.POINTER(_ptr)]),
    "dsw_plan_destroy": (None, [_ptr]),
    "dsw_plan_shape": (C.c_int, [_ptr, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64), C.POINTER(_i32)]),
    "dsw_plan_operand_bytes": (_i64, [_ptr]),
    "dsw_cheb_fwd_workspace_bytes": (_sz, [_i32] * 5),
    "dsw_cheb_fwd": (C.c_int, [_ptr, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _ptr, _sz, _ptr]),
    "dsw_cheb_terms": (C.c_int, [_ptr, _ptr, _i64, _i64, _ptr, _i32, _i32, _i32, _ptr]),
    "dsw_cheb_fwd_algo": (C.c_int, [_i32] * 3),
    "dsw_cheb_bwd_algo": (C.c_int, [_i32] * 3),
    "dsw_cheb_bwd_workspace_bytes": (_sz, [_i32] * 6),
    "dsw_cheb_bwd": (C.c_int, [_ptr, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr, _sz, _ptr]),
    "dsw_cheb_bwd_ex": (C.c_int, [_ptr, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _ptr, _sz, _ptr]),
    "dsw_cheb_bwd_data_workspace_bytes": (_sz, [_i32] * 5),
    "dsw_cheb_bwd_data": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr, _sz, _ptr]),
    "dsw_cheb_bwd_weight_workspace_bytes": (_sz, [_i32] * 5),
    "dsw_cheb_bwd_weight": (C.c_int, [_ptr, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr, _sz, _ptr]),
    "dsw_linear_workspace_bytes": (_sz, [_i32] * 4),
    "dsw_linear_fwd": (C.c_int, [_ptr, _i64, _i64, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr, _sz, _ptr]),
    "dsw_linear_bwd": (C.c_int, [_ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr, _sz, _ptr]),
    "dsw_linear_rezero_fwd": (C.c_int, [_ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr, _sz, _ptr]),
    "dsw_linear_bwd_acc": (C.c_int, [_ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr, _sz, _ptr]),
    "dsw_linear_rezero_fwd_ld": (C.c_int, [_ptr, _i64, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _i32, _i32, _ptr, _sz, _ptr]),
    "dsw_rezero_fwd": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i64, _ptr]),
    "dsw_rezero_bwd_workspace_bytes": (_sz, []),
    "dsw_rezero_bwd": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _sz, _i64, _ptr]),
    "dsw_wmse_workspace_bytes": (_sz, []),
    "dsw_wmse_fwd": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _sz, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_wmse_bwd": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _ptr]),
    "dsw_wmse_none_fwd": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _ptr]),
    "dsw_wmse_none_bwd": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _ptr]),
    "dsw_spmm_fwd": (C.c_int, [_ptr, _ptr, _i64, _i64, _ptr, _i32, _i32, _ptr]),
    "dsw_spmm_bwd": (C.c_int, [_ptr, _ptr, _i64, _i64, _ptr, _i32, _i32, _ptr]),
    "dsw_spmm_fwd_ex": (C.c_int, [_ptr, _ptr, _i64, _i64, _ptr, _i64, _i64, _ptr, _i64, _i64, _i32, _i32, _ptr]),
    "dsw_spmm_bwd_ex": (C.c_int, [_ptr, _ptr, _i64, _i64, _ptr, _i64, _i64, _ptr, _i64, _i64, _i32, _i32, _ptr]),
    "dsw_maxval_pool_fwd": (C.c_int, [_ptr, _ptr, _i64, _i64, _ptr, _ptr, _ptr, _i32, _i32, _ptr]),
    "dsw_maxval_pool_bwd": (C.c_int, [_ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_scatter_unpool_fwd": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_scatter_unpool_bwd": (C.c_int, [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_nested_maxpool_fwd": (C.c_int, [_ptr, _i64, _i64, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_nested_scatter": (C.c_int, [_ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_nested_gather": (C.c_int, [_ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_nested_avgpool_fwd": (C.c_int, [_ptr, _i64, _i64, _ptr, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_nested_repeat": (C.c_int, [_ptr, _i64, _i64, _ptr, _f32, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_nested_sum": (C.c_int, [_ptr, _i64, _i64, _ptr, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_ar_stack_fwd": (C.c_int, [_ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_ar_stack_bwd": (C.c_int, [_ptr, _ptr, _i32, _i32, _i32, _i32, _i32, _i32, _ptr]),
    "dsw_graph_workspace_bytes": (_sz, [_i32, _i32]),
    "dsw_graph_nnz_capacity": (_i64, [_i32, _i32]),
    "dsw_graph_knn_laplacian": (C.c_int, [_ptr, _i32, _i32, _i32, C.c_double, _i64, _ptr, _ptr, _ptr, C.POINTER(_i64),
                                          C.POINTER(C.c_double), _ptr, _sz, _ptr]),
    "dsw_graph_nested_pool": (C.c_int, [_i32, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr]),
    "dsw_launch_count": (_i64, []),
    "dsw_set_mix_mode": (C.c_int, [C.c_int]),
    "dsw_get_mix_mode": (C.c_int, []),
    "dsw_debug_counters": (C.c_int, [_ptr, C.c_int]),
    "dsw_debug_dense_counters": (C.c_int, [_ptr, C.c_int]),
    "dsw_debug_chain_counters": (C.c_int, [_ptr, C.c_int]),
    "dsw_debug_mix_counters": (C.c_int, [_ptr, C.c_int]),
    "dsw_set_option": (C.c_int, [C.c_int, _i64]),
    "dsw_get_option": (_i64, [C.c_int]),
}

_lib = None


class DswError(RuntimeError):
    pass


def header_symbols():
    """Every function name declared in include/dsw.h."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsw_[a-z0-9_]+)\s*\(", text)))


def load():
    """Load libdsw.so (once).  Raises if it has not been built — there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise DswError(
            f"{LIB_PATH} not found: build it with `python deepsphere-weather_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        if "DSW_LIB_PATH" in os.environ and not hasattr(lib, name):
            continue  # an older build under A/B test; calling the missing entry point raises AttributeError
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    # tuning switches for A/B runs: DSW_OPTIONS="key=value,key=value" (keys: DSW_OPT_* of include/dsw.h)
    for kv in filter(None, os.environ.get("DSW_OPTIONS", "").split(",")):
        k, v = kv.split("=")
        if lib.dsw_set_option(int(k), int(v)) != 0:
            raise DswError(f"DSW_OPTIONS: bad option {kv!r}")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc == 0:
        return
    lib = load()
    msg = lib.dsw_strerror(rc).decode()
    if rc == -5:
        msg += ": " + lib.dsw_last_cuda_error_string().decode()
    raise DswError(f"{what or 'libdsw'} failed ({rc}): {msg}")
