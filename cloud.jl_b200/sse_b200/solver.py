"""Host mirror of the reference Solver surface for the path (Python stand-in for the Julia host).

    Solver(...)                      src/Solvers/Solvers.jl:259-376
    semidiscretize(...)              src/Solvers/Solvers.jl:429-452
    semi_discrete_residual!(...)     src/Solvers/Solvers.jl:474-564

`u` and `dudt` are torch CUDA tensors of shape (N_e, N_c, N_p), float64, C-contiguous — the
memory image of the reference's column-major (N_p, N_c, N_e) arrays.  NumPy arrays are accepted
as well; they are staged through pinned host memory and copied to/from the device around the
call (the "host buffers" end-to-end path).  torch only owns memory and streams here; all
arithmetic happens in libsse_b200.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from . import _abi, _lib
from .assembly import (REFERENCE_OPERATOR, PHYSICAL_OPERATOR, SolverImage, SpatialDiscretization,
                       assemble)
from .laws import project_function

# Carpenter & Kennedy (1994) 2N-storage RK4(5) (OrdinaryDiffEq's CarpenterKennedy2N54)
CK54_A = (0.0, -567301805773 / 1357537059087, -2404267990393 / 2016746695238,
          -3550918686646 / 2091501179385, -1275806237668 / 842570457699)
CK54_B = (1432997174477 / 9575080441755, 5161836677717 / 13612068292357,
          1720146321549 / 2090206949498, 3134564353537 / 4481467310338,
          2277821191437 / 14882151754819)
CK54_C = (0.0, 1432997174477 / 9575080441755, 2526269341429 / 6820363962896,
          2006345519317 / 3224310063776, 2802321613138 / 2924317926251)


def _torch():
    import torch
    return torch


class _DevView:
    """__cuda_array_interface__ wrapper of a library-owned device buffer."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 3, "strides": None}


class Solver:
    """One reference `Solver` resident on one GPU (handle of libsse_b200.so)."""

    def __init__(self, image: SolverImage, device: int = 0):
        self.image = image
        self.cfg = image.cfg
        self.device = device
        self._lib = _lib.load()
        self._h = C.c_void_p()
        arr = image.c_arrays()
        _lib.check(self._lib.sse_create(C.byref(image.cfg), C.byref(arr), device, C.byref(self._h)))
        self._pinned = {}

    # -- plumbing ------------------------------------------------------------------------
    def close(self):
        if self._h:
            self._lib.sse_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def state_shape(self) -> Tuple[int, int, int]:
        return self.image.state_shape

    def size(self):
        """Base.size(solver) = (N_p, N_c, N_e) (Solvers.jl:276-285)."""
        return (int(self.cfg.N_p), int(self.cfg.N_c), int(self.cfg.N_e))

    def new_state(self):
        torch = _torch()
        return torch.zeros(self.state_shape, dtype=torch.float64, device=f"cuda:{self.device}")

    def use_current_stream(self):
        torch = _torch()
        s = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self._lib.sse_set_stream(self._h, C.c_void_p(s)))

    @property
    def launches(self) -> int:
        """Kernels launched through this handle, counted by the library at its launch sites (bench accounting)."""
        n = C.c_int64(0)
        _lib.check(self._lib.sse_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def set_kernel_variant(self, v: int):
        _lib.check(self._lib.sse_set_kernel_variant(self._h, v))

    def kernel_variant(self) -> int:
        v = C.c_int32(0)
        _lib.check(self._lib.sse_get_kernel_variant(self._h, C.byref(v)))
        return int(v.value)

    def _check_state(self, x, name):
        torch = _torch()
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
                and tuple(x.shape) == self.state_shape):
            raise ValueError(f"{name}: expected contiguous float64 CUDA tensor of shape {self.state_shape} "
                             "(DimensionMismatch)")
        return C.c_void_p(x.data_ptr())

    # -- the path ------------------------------------------------------------------------
    def rhs(self, dudt, u, t: float = 0.0):
        _lib.check(self._lib.sse_rhs(self._h, self._check_state(u, "u"), self._check_state(dudt, "dudt"), float(t)))
        return dudt

    def pass_a(self, u):
        _lib.check(self._lib.sse_rhs_pass_a(self._h, self._check_state(u, "u")))

    def pass_aux(self, dudt, first, count):
        _lib.check(self._lib.sse_rhs_pass_aux(self._h, self._check_state(dudt, "dudt"), first, count))

    def pass_b(self, dudt, first, count):
        _lib.check(self._lib.sse_rhs_pass_b(self._h, self._check_state(dudt, "dudt"), first, count))

    def pass_a_range(self, u, first, count):
        _lib.check(self._lib.sse_rhs_pass_a_range(self._h, self._check_state(u, "u"), first, count))

    def _chunk_plan(self, chunks: int):
        key = ("plan", chunks)
        if key not in self._pinned:
            self._pinned[key] = chunk_plan(self.image.arrays["mapP"], self.state_shape[0], int(self.cfg.N_f), chunks)
        return self._pinned[key]

    def rhs_host(self, dudt_host, u_host, t: float = 0.0, chunks: int = 0):
        """Residual on HOST buffers (the reference-facing call with `Array` arguments) = one `sse_rhs_host`: the H2D copy
        of u, the kernels and the D2H copy of dudt are pipelined inside the library over `chunks` element ranges (0: the
        library default) on three streams — pass A of a range starts as soon as its slice of u has arrived, pass B of a
        range as soon as pass A has covered its face neighbours (read from mapP), and the download of its dudt overlaps
        the uploads still in flight (full-duplex PCIe).  torch CPU tensors (ideally pinned) are passed directly; NumPy
        arrays are staged through pinned memory."""
        torch = _torch()
        p = self._pinned
        if isinstance(u_host, np.ndarray):
            if "u" not in p:
                p["u"] = torch.empty(self.state_shape, dtype=torch.float64).pin_memory()
                p["du"] = torch.empty(self.state_shape, dtype=torch.float64).pin_memory()
            p["u"].numpy()[...] = u_host
            src, dst = p["u"], p["du"]
        else:
            src, dst = u_host, dudt_host
        for x, name in ((src, "u"), (dst, "dudt")):
            if x.is_cuda or x.dtype != torch.float64 or tuple(x.shape) != tuple(self.state_shape) or not x.is_contiguous():
                raise ValueError(f"{name}: expected a contiguous float64 CPU tensor of shape {self.state_shape}")
        _lib.check(self._lib.sse_rhs_host(self._h, C.c_void_p(src.data_ptr()), C.c_void_p(dst.data_ptr()), float(t), int(chunks)))
        if isinstance(u_host, np.ndarray):
            dudt_host[...] = dst.numpy()
        return dudt_host

    def axpby(self, a, x, b, y):
        _lib.check(self._lib.sse_axpby(self._h, a, self._check_state(x, "x"), b, self._check_state(y, "y")))

    def lsrk_stage(self, u, tmp, dudt, A, B, dt):
        _lib.check(self._lib.sse_lsrk_stage(self._h, self._check_state(u, "u"), self._check_state(tmp, "tmp"),
                                            self._check_state(dudt, "dudt"), A, B, dt))

    def rhs_lsrk(self, u, tmp, dudt, A, B, dt, t=0.0):
        """Residual + one 2N-storage RK stage (fused on the compile-time kernel path)."""
        _lib.check(self._lib.sse_rhs_lsrk(self._h, self._check_state(u, "u"), self._check_state(tmp, "tmp"),
                                          self._check_state(dudt, "dudt"), A, B, dt, float(t)))

    def step_ck54(self, u, tmp, dudt, t, dt):
        _lib.check(self._lib.sse_step_ck54(self._h, self._check_state(u, "u"), self._check_state(tmp, "tmp"),
                                           self._check_state(dudt, "dudt"), t, dt))

    def profile_rhs(self, dudt, u, reps: int = 5) -> np.ndarray:
        """Milliseconds of [pass A, auxiliary pass, pass B first kernel, pass B second kernel] (CUDA events between the kernels)."""
        ms = np.zeros(4)
        _lib.check(self._lib.sse_profile_rhs(self._h, self._check_state(u, "u"), self._check_state(dudt, "dudt"), int(reps),
                                             ms.ctypes.data_as(C.POINTER(C.c_double))))
        return ms

    def set_graph_mode(self, on: bool = True):
        """Replay sse_step_ck54 as one CUDA graph (launch-latency-bound meshes)."""
        _lib.check(self._lib.sse_set_graph_mode(self._h, 1 if on else 0))

    def functionals(self, u, dudt) -> np.ndarray:
        out = np.zeros(int(self.cfg.N_c) + 2)
        _lib.check(self._lib.sse_functionals(self._h, self._check_state(u, "u"), self._check_state(dudt, "dudt"),
                                             out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def synchronize(self):
        _lib.check(self._lib.sse_synchronize(self._h))

    def debug_views(self):
        """(u_q, u_f) scratch as torch tensors: u_q (N_e, N_c, N_q); u_f (N_c, N_f*N_e + ghost)."""
        torch = _torch()
        pq, pf = C.POINTER(C.c_double)(), C.POINTER(C.c_double)()
        _lib.check(self._lib.sse_debug_views(self._h, C.byref(pq), C.byref(pf)))
        c = self.cfg
        nq = int(c.N_e) * int(c.N_c) * int(c.N_q)
        nft = int(c.N_f) * int(c.N_e) + int(c.N_ghost)
        uq = torch.as_tensor(_DevView(C.cast(pq, C.c_void_p).value, nq), device=f"cuda:{self.device}")
        uf = torch.as_tensor(_DevView(C.cast(pf, C.c_void_p).value, nft * int(c.N_c)), device=f"cuda:{self.device}")
        return uq.view(int(c.N_e), int(c.N_c), int(c.N_q)), uf.view(int(c.N_c), nft)

    # -- multi-GPU: the exchange lives in the library (csrc/comm.cu) --------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        _lib.check(_lib.load().sse_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _lib.check(self._lib.sse_comm_init(self._h, buf, int(rank), int(world)))

    @staticmethod
    def comm_init_all(solvers):
        hs = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
        _lib.check(_lib.load().sse_comm_init_all(hs, len(solvers)))

    @staticmethod
    def comm_init_local(solvers):
        hs = (C.c_void_p * len(solvers))(*[s._h for s in solvers])
        _lib.check(_lib.load().sse_comm_init_local(hs, len(solvers)))

    def comm_info(self):
        r, w, v = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        _lib.check(self._lib.sse_comm_info(self._h, C.byref(r), C.byref(w), C.byref(v)))
        return int(r.value), int(w.value), int(v.value)

    def halo_plan(self, mesh):
        """The halo plan of a partitioned mesh (mesh.nbr_ranks / send_idx / recv_off / n_boundary)."""
        nbr = np.asarray(mesh.nbr_ranks or [], dtype=np.int32)
        sc = np.asarray([int(x.size) for x in (mesh.send_idx or [])], dtype=np.int64)
        offs = list(mesh.recv_off or []) + [int(mesh.n_ghost)]
        rc = np.asarray([offs[i + 1] - offs[i] for i in range(nbr.size)], dtype=np.int64)
        idx = (np.concatenate(mesh.send_idx) if nbr.size else np.zeros(0, dtype=np.int64)) + 1
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        p32, p64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        _lib.check(self._lib.sse_halo_plan(self._h, int(nbr.size), nbr.ctypes.data_as(p32), sc.ctypes.data_as(p64),
                                           rc.ctypes.data_as(p64), idx.ctypes.data_as(p64),
                                           int(mesh.N_e - mesh.n_boundary)))

    @staticmethod
    def rhs_multi(solvers, dudts, us, t: float = 0.0):
        """One process driving several GPUs: the residual on all partitions (sse_rhs_multi)."""
        n = len(solvers)
        hs = (C.c_void_p * n)(*[s._h for s in solvers])
        pu = (C.c_void_p * n)(*[s._check_state(u, "u").value for s, u in zip(solvers, us)])
        pd = (C.c_void_p * n)(*[s._check_state(d, "dudt").value for s, d in zip(solvers, dudts)])
        _lib.check(_lib.load().sse_rhs_multi(hs, n, pu, pd, float(t)))
        return dudts

    @staticmethod
    def step_ck54_multi(solvers, us, tmps, dudts, t, dt):
        n = len(solvers)
        hs = (C.c_void_p * n)(*[s._h for s in solvers])
        arr = lambda xs, nm: (C.c_void_p * n)(*[s._check_state(x, nm).value for s, x in zip(solvers, xs)])
        _lib.check(_lib.load().sse_step_ck54_multi(hs, n, arr(us, "u"), arr(tmps, "tmp"), arr(dudts, "dudt"), float(t), float(dt)))

    # -- halo buffers (split form: a caller that runs its own exchange) -----------------------------
    def halo_configure(self, send_index_1based: np.ndarray):
        idx = np.ascontiguousarray(send_index_1based, dtype=np.int64)
        _lib.check(self._lib.sse_halo_configure(self._h, idx.ctypes.data_as(C.POINTER(C.c_int64)), idx.size))

    def halo_buffers(self, which: int = 0):
        torch = _torch()
        p, n = C.POINTER(C.c_double)(), C.c_int64(0)
        _lib.check(self._lib.sse_halo_send_buffer(self._h, C.byref(p), C.byref(n)))
        nv = int(self.cfg.N_c) * (int(self.cfg.d) if which == 1 else 1)
        nsend = (n.value // (int(self.cfg.N_c) * (int(self.cfg.d) if self.image.law.second_order else 1))) * nv
        dev = f"cuda:{self.device}"
        send = torch.as_tensor(_DevView(C.cast(p, C.c_void_p).value, max(nsend, 1)), device=dev)[:nsend]
        _lib.check(self._lib.sse_halo_recv_buffer(self._h, which, C.byref(p), C.byref(n)))
        recv = torch.as_tensor(_DevView(C.cast(p, C.c_void_p).value, max(n.value, 1)), device=dev)[:n.value]
        return send, recv

    def halo_pack(self, which: int = 0):
        _lib.check(self._lib.sse_halo_pack(self._h, which))

    def halo_unpack(self, which: int = 0):
        _lib.check(self._lib.sse_halo_unpack(self._h, which))


def range_plan(mapP_1based, ne: int, nf: int, ranges):
    """For element ranges (a, b) listed in upload order: after[i] = the ranges whose pass B may run once pass A has
    covered ranges[0..i], i.e. once every range holding one of their face neighbours (read from mapP) is through.
    Neighbours beyond the local elements (ghost facet slots of a partitioned mesh) are not range dependencies."""
    nb = (np.asarray(mapP_1based).reshape(-1, nf)[:ne] - 1) // nf
    owner = np.full(ne, -1, dtype=np.int64)
    for i, (a, b) in enumerate(ranges):
        owner[a:b] = i
    ready = []
    for i, (a, b) in enumerate(ranges):
        nbr = nb[a:b].reshape(-1)
        dep = owner[nbr[nbr < ne]]
        ready.append(max(i, int(dep.max()) if dep.size else i))
    return [[k for k in range(len(ranges)) if ready[k] == i] for i in range(len(ranges))]


def chunk_plan(mapP_1based, ne: int, nf: int, chunks: int):
    """Schedule of the pipelined host-buffer residual (Solver.rhs_host): element ranges `bounds`, their upload order
    `up`, and for every upload position the ranges whose pass B becomes runnable there (`after`).  On a slab-ordered
    periodic mesh every range waits for its two neighbours, so three ranges are left when the uploads end."""
    bounds = [ne * c // chunks for c in range(chunks + 1)]
    after = range_plan(mapP_1based, ne, nf, [(bounds[c], bounds[c + 1]) for c in range(chunks)])
    return bounds, list(range(chunks)), after


def fp64_peak(device: int = 0) -> float:
    """Measured register-resident DFMA throughput in FLOP/s (FMA = 2)."""
    out = C.c_double(0.0)
    _lib.check(_lib.load().sse_fp64_peak(device, C.byref(out)))
    return float(out.value)


@dataclass
class ODEProblem:
    """ODEProblem(semi_discrete_residual!, u0, tspan, solver) (Solvers.jl:451)."""
    f: object
    u0: np.ndarray
    tspan: Tuple[float, float]
    p: Solver


def semi_discrete_residual(dudt, u, solver: Solver, t: float = 0.0):
    """semi_discrete_residual!(dudt, u, solver, t) (Solvers.jl:474-564); returns dudt."""
    if isinstance(u, np.ndarray) or not u.is_cuda:
        return solver.rhs_host(dudt, u, t)
    return solver.rhs(dudt, u, t)


def semidiscretize(conservation_law, spatial_discretization: SpatialDiscretization, initial_data, form,
                   tspan, strategy: str = REFERENCE_OPERATOR, mass_matrix_solver: Optional[int] = None,
                   device: int = 0) -> ODEProblem:
    """semidiscretize (Solvers.jl:429-452): project the initial data, build the Solver."""
    sd = spatial_discretization
    u0 = project_function(initial_data, sd.reference_approximation, sd.geometric_factors.J_q, sd.mesh.xyzq)
    image = assemble(conservation_law, sd, form, strategy, mass_matrix_solver)
    return ODEProblem(semi_discrete_residual, u0, tuple(tspan), Solver(image, device))


def solve_ck54(problem: ODEProblem, dt: float, n_steps: int, fused: bool = True):
    """solve(ode, CarpenterKennedy2N54(); dt, adaptive=false) with the state resident on the device
    (test/test_driver.jl:77-83).  Returns the final state as a NumPy array."""
    torch = _torch()
    s = problem.p
    u = torch.from_numpy(np.ascontiguousarray(problem.u0)).to(f"cuda:{s.device}")
    tmp, du = s.new_state(), s.new_state()
    t = problem.tspan[0]
    for _ in range(n_steps):
        if fused:
            s.step_ck54(u, tmp, du, t, dt)
        else:
            for st in range(5):
                s.rhs(du, u, t + CK54_C[st] * dt)
                s.lsrk_stage(u, tmp, du, CK54_A[st], CK54_B[st], dt)
        t += dt
    s.synchronize()
    return u.cpu().numpy()
