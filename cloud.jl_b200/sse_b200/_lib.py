"""ctypes binding of libsse_b200.so (the CUDA library behind include/sse_b200.h).

There is deliberately no fallback: if the shared library is missing or no CUDA device
is usable, every compute call raises."""
from __future__ import annotations

import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SSE_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libsse_b200.so")

# every symbol include/sse_b200.h declares: name -> (restype, argtypes)
_pd = C.POINTER(C.c_double)
_ppd = C.POINTER(_pd)
_pi64 = C.POINTER(C.c_int64)
_h = C.c_void_p
SYMBOLS = {
    "sse_create": (C.c_int32, [C.POINTER(_abi.sse_config), C.POINTER(_abi.sse_arrays), C.c_int32, C.POINTER(_h)]),
    "sse_destroy": (C.c_int32, [_h]),
    "sse_set_stream": (C.c_int32, [_h, C.c_void_p]),
    "sse_set_kernel_variant": (C.c_int32, [_h, C.c_int32]),
    "sse_get_kernel_variant": (C.c_int32, [_h, C.POINTER(C.c_int32)]),
    "sse_state_alloc": (C.c_int32, [_h, _ppd]),
    "sse_state_free": (C.c_int32, [_h, C.c_void_p]),
    "sse_state_fill": (C.c_int32, [_h, C.c_void_p, C.c_double]),
    "sse_state_upload": (C.c_int32, [_h, C.c_void_p, C.c_void_p]),
    "sse_state_download": (C.c_int32, [_h, C.c_void_p, C.c_void_p]),
    "sse_rhs": (C.c_int32, [_h, C.c_void_p, C.c_void_p, C.c_double]),
    "sse_rhs_host": (C.c_int32, [_h, C.c_void_p, C.c_void_p, C.c_double, C.c_int32]),
    "sse_host_range_plan": (C.c_int32, [_pi64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "sse_host_pin": (C.c_int32, [C.c_void_p, C.c_int64]),
    "sse_host_unpin": (C.c_int32, [C.c_void_p]),
    "sse_rhs_pass_a": (C.c_int32, [_h, C.c_void_p]),
    "sse_rhs_pass_a_range": (C.c_int32, [_h, C.c_void_p, C.c_int64, C.c_int64]),
    "sse_rhs_pass_aux": (C.c_int32, [_h, C.c_void_p, C.c_int64, C.c_int64]),
    "sse_rhs_pass_b": (C.c_int32, [_h, C.c_void_p, C.c_int64, C.c_int64]),
    "sse_halo_configure": (C.c_int32, [_h, _pi64, C.c_int64]),
    "sse_halo_pack": (C.c_int32, [_h, C.c_int32]),
    "sse_halo_send_buffer": (C.c_int32, [_h, _ppd, _pi64]),
    "sse_halo_recv_buffer": (C.c_int32, [_h, C.c_int32, _ppd, _pi64]),
    "sse_halo_unpack": (C.c_int32, [_h, C.c_int32]),
    "sse_comm_unique_id": (C.c_int32, [C.c_void_p]),
    "sse_comm_init": (C.c_int32, [_h, C.c_void_p, C.c_int32, C.c_int32]),
    "sse_comm_init_all": (C.c_int32, [C.POINTER(_h), C.c_int32]),
    "sse_comm_init_local": (C.c_int32, [C.POINTER(_h), C.c_int32]),
    "sse_comm_info": (C.c_int32, [_h, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "sse_halo_plan": (C.c_int32, [_h, C.c_int32, C.POINTER(C.c_int32), _pi64, _pi64, _pi64, C.c_int64]),
    "sse_rhs_multi": (C.c_int32, [C.POINTER(_h), C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_double]),
    "sse_step_ck54_multi": (C.c_int32, [C.POINTER(_h), C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p), C.c_double, C.c_double]),
    "sse_partition_create": (C.c_int32, [_pi64, C.c_int64, C.c_int32, C.POINTER(C.c_int32), C.c_int32, C.c_int32, C.POINTER(_h)]),
    "sse_partition_sizes": (C.c_int32, [_h, _pi64, _pi64, _pi64, C.POINTER(C.c_int32), _pi64]),
    "sse_partition_fill": (C.c_int32, [_h, _pi64, _pi64, C.POINTER(C.c_int32), _pi64, _pi64, _pi64]),
    "sse_partition_destroy": (C.c_int32, [_h]),
    "sse_axpby": (C.c_int32, [_h, C.c_double, C.c_void_p, C.c_double, C.c_void_p]),
    "sse_lsrk_stage": (C.c_int32, [_h, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double]),
    "sse_rhs_lsrk": (C.c_int32, [_h, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]),
    "sse_step_ck54": (C.c_int32, [_h, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double]),
    "sse_set_graph_mode": (C.c_int32, [_h, C.c_int32]),
    "sse_functionals": (C.c_int32, [_h, C.c_void_p, C.c_void_p, _pd]),
    "sse_synchronize": (C.c_int32, [_h]),
    "sse_last_error_string": (C.c_char_p, []),
    "sse_abi_version": (C.c_int32, []),
    "sse_profile_rhs": (C.c_int32, [_h, C.c_void_p, C.c_void_p, C.c_int32, _pd]),
    "sse_launch_count": (C.c_int32, [_h, _pi64]),
    "sse_debug_views": (C.c_int32, [_h, _ppd, _ppd]),
    "sse_plan_selfcheck": (C.c_int32, [C.POINTER(_abi.sse_config), C.POINTER(_abi.sse_arrays), C.POINTER(C.c_int32), _pd]),
    "sse_fp64_peak": (C.c_int32, [C.c_int32, _pd]),
    "sse_geometric_factors": (C.c_int32, [C.POINTER(_abi.sse_geom_config), C.POINTER(_abi.sse_geom_ops), C.c_int32,
                                          C.POINTER(_pd), _pd, _pd, _pd, _pd]),
}

_LIB = None


class SSEError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsse_b200 status {code}: {msg}")
        self.code = code


def load():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.sse_abi_version() != _abi.SSE_ABI_VERSION:
            raise RuntimeError("libsse_b200.so ABI version does not match sse_b200/_abi.py")
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        msg = load().sse_last_error_string()
        raise SSEError(rc, msg.decode() if msg else "")
