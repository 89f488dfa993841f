"""ctypes wrapper of oracle/libsse_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may import
this module.  It consumes the same SolverImage (config + reference-layout host arrays) that
the product hands to sse_create."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), "cloud.jl_b200"))
from sse_b200 import _abi  # noqa: E402  (struct layouts only)

_LIB = None
_PERF = None


def host_threads() -> int:
    """Host threads this process may use (the affinity mask, not OMP_NUM_THREADS: torch.distributed.run sets that to 1)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def build_perf() -> str:
    """The same source with the flags a user would time a CPU code with (-O3 -march=native, contraction on), for the
    baseline legs of bench.py only -- parity always uses libsse_oracle.so (-ffp-contract=off).  -march=native must be
    resolved on the machine that runs it, so the library is compiled on first use there and keyed by the CPU model."""
    import hashlib
    try:
        model = [l for l in open("/proc/cpuinfo") if l.startswith(("model name", "flags"))][:2]
    except OSError:
        model = []
    key = hashlib.sha1("".join(model).encode()).hexdigest()[:10]
    out = os.path.join(_HERE, "_perf")
    so = os.path.join(out, f"libsse_oracle_perf_{key}.so")
    src = os.path.join(_HERE, "sse_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        os.makedirs(out, exist_ok=True)
        subprocess.check_call(["make", "-C", _HERE, "-B", "perf", f"PERF_SO={so}"], stdout=subprocess.DEVNULL)
    return so


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libsse_oracle.so")
    src = os.path.join(_HERE, "sse_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsse_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def _bind(path):
    L = C.CDLL(path)
    pd = C.POINTER(C.c_double)
    L.sse_oracle_rhs.restype = C.c_int32
    L.sse_oracle_rhs.argtypes = [C.POINTER(_abi.sse_config), C.POINTER(_abi.sse_arrays), pd, pd,
                                 C.c_int32, pd, pd]
    L.sse_oracle_time_rhs.restype = C.c_double
    L.sse_oracle_time_rhs.argtypes = [C.POINTER(_abi.sse_config), C.POINTER(_abi.sse_arrays), pd, pd,
                                      C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    L.sse_oracle_logmean.restype = C.c_double
    L.sse_oracle_logmean.argtypes = [C.c_double, C.c_double]
    L.sse_oracle_inv_logmean.restype = C.c_double
    L.sse_oracle_inv_logmean.argtypes = [C.c_double, C.c_double]
    L.sse_oracle_two_point_flux.argtypes = [C.POINTER(_abi.sse_config), C.c_int32, pd, pd, pd]
    L.sse_oracle_cons_to_entropy.argtypes = [C.POINTER(_abi.sse_config), pd, pd]
    L.sse_oracle_entropy_to_cons.argtypes = [C.POINTER(_abi.sse_config), pd, pd]
    return L


def lib():
    global _LIB
    if _LIB is None:
        _LIB = _bind(build())
    return _LIB


def perf_lib():
    global _PERF
    if _PERF is None:
        _PERF = _bind(build_perf())
    return _PERF


def _pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def rhs(image, u, nthreads: int = 0, return_scratch: bool = False):
    """dudt = semi_discrete_residual!(similar(u), u, solver, t).  u: (N_e, N_c, N_p) C-ordered."""
    cfg = image.cfg
    u = np.ascontiguousarray(u, dtype=np.float64)
    assert u.shape == image.state_shape, (u.shape, image.state_shape)
    dudt = np.zeros_like(u)
    arr = image.c_arrays()
    uq = uf = None
    if return_scratch:
        uq = np.zeros((cfg.N_e, cfg.N_c, cfg.N_q))
        uf = np.zeros((cfg.N_c, cfg.N_e, cfg.N_f))
    rc = lib().sse_oracle_rhs(C.byref(cfg), C.byref(arr), _pd(u), _pd(dudt), nthreads, _pd(uq), _pd(uf))
    if rc != 0:
        raise RuntimeError(f"oracle failed with status {rc}")
    return (dudt, uq, uf) if return_scratch else dudt


def time_rhs(image, u, nthreads: int = 0, reps: int = 3, perf: bool = False):
    """Seconds per RHS (best of reps) and the number of OpenMP threads used.  nthreads = 0 leaves the count to the
    OpenMP runtime (OMP_NUM_THREADS); perf = True times the -O3 -march=native build (build_perf)."""
    cfg = image.cfg
    u = np.ascontiguousarray(u, dtype=np.float64)
    dudt = np.zeros_like(u)
    arr = image.c_arrays()
    used = C.c_int32(0)
    t = (perf_lib() if perf else lib()).sse_oracle_time_rhs(C.byref(cfg), C.byref(arr), _pd(u), _pd(dudt), nthreads, reps, C.byref(used))
    return t, int(used.value), dudt
