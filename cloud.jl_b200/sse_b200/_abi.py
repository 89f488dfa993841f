"""ctypes mirror of include/sse_b200.h (struct layouts and enums only)."""
from __future__ import annotations

import ctypes as C

import numpy as np

SSE_ABI_VERSION = 1

SSE_OK, SSE_ERR_BAD_ARGUMENT, SSE_ERR_UNSUPPORTED, SSE_ERR_CUDA, SSE_ERR_NONFINITE, SSE_ERR_COMM = range(6)
SSE_PDE_ADVECTION, SSE_PDE_ADVECTION_DIFFUSION, SSE_PDE_EULER, SSE_PDE_BURGERS, SSE_PDE_VISCOUS_BURGERS = 0, 1, 2, 3, 4
SSE_FORM_STANDARD_REFERENCE, SSE_FORM_STANDARD_PHYSICAL, SSE_FORM_FLUX_DIFFERENCING = 0, 1, 2
SSE_FLUX_LAX_FRIEDRICHS, SSE_FLUX_CENTRAL, SSE_FLUX_ENTROPY_CONSERVATIVE = 0, 1, 2
SSE_VISCOUS_NONE, SSE_VISCOUS_BR1 = 0, 1
SSE_TWO_POINT_CONSERVATIVE, SSE_TWO_POINT_ENTROPY_CONSERVATIVE = 0, 1
SSE_MASS_WEIGHT_ADJUSTED, SSE_MASS_DIAGONAL, SSE_MASS_CHOLESKY = 0, 1, 2
SSE_V_IDENTITY, SSE_V_DENSE, SSE_V_WARPED = 0, 1, 2

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int64)


class sse_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("d", C.c_int32),
        ("N_c", C.c_int32), ("N_p", C.c_int32), ("N_q", C.c_int32), ("N_f", C.c_int32),
        ("N_fac", C.c_int32), ("p", C.c_int32),
        ("N_e", C.c_int64), ("N_ghost", C.c_int64),
        ("pde", C.c_int32), ("form", C.c_int32), ("inviscid_flux", C.c_int32),
        ("viscous_flux", C.c_int32), ("two_point_flux", C.c_int32), ("mass_solver", C.c_int32),
        ("v_kind", C.c_int32), ("M1d", C.c_int32 * 3),
        ("half_lambda", C.c_double), ("a", C.c_double * 3), ("b", C.c_double), ("gamma", C.c_double),
    ]


class sse_arrays(C.Structure):
    _fields_ = [
        ("V", _pd), ("A", _pd), ("B", _pd), ("C", _pd), ("sigma_i", _pi), ("sigma_o", _pi),
        ("R", _pd), ("W", _pd), ("Bf", _pd), ("D", _pd * 3), ("S", _pd * 3), ("Cfd", _pd),
        ("J_q", _pd), ("Lambda_q", _pd), ("J_f", _pd), ("nJf", _pd), ("nJq", _pd), ("nref", _pd),
        ("VOL", _pd), ("FAC", _pd), ("mapP", _pi),
    ]


SSE_METRIC_EXACT, SSE_METRIC_CURL = 0, 1


class sse_geom_config(C.Structure):
    _fields_ = [("d", C.c_int32), ("N_map", C.c_int32), ("N1", C.c_int32), ("N_q", C.c_int32), ("N_f", C.c_int32),
                ("metric", C.c_int32), ("N_e", C.c_int64)]


class sse_geom_ops(C.Structure):
    _fields_ = [("Drst", _pd * 3), ("Vq", _pd), ("Vf", _pd), ("nrstJ", _pd), ("up", _pd), ("D1", _pd * 3),
                ("Vq1", _pd), ("Vf1", _pd)]


def _ptr(a, ty):
    if a is None:
        return C.cast(None, ty)
    return a.ctypes.data_as(ty)


def fill_arrays(arrays: dict) -> sse_arrays:
    """arrays: name -> flat contiguous numpy buffer in the reference (column-major) memory order."""
    s = sse_arrays()
    for name, ty in sse_arrays._fields_:
        if name in ("D", "S"):
            vals = arrays.get(name) or [None, None, None]
            vals = list(vals) + [None] * (3 - len(vals))
            arr = (_pd * 3)(*[_ptr(v, _pd) for v in vals])
            setattr(s, name, arr)
        else:
            a = arrays.get(name)
            if a is not None:
                want = np.int64 if ty is _pi else np.float64
                assert a.dtype == want and a.flags["C_CONTIGUOUS"], name
            setattr(s, name, _ptr(a, ty))
    return s
