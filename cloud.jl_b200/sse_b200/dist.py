"""Element-partitioned residual over N GPUs: one process per GPU, facet-trace halos only.

The reference has no distributed path; the only cross-element coupling of
`semi_discrete_residual!` is the gather of neighbour facet states through `mesh.mapP`
between the two element loops (Solvers.jl:505-511).  On the GPU the whole exchange runs inside
libsse_b200.so (csrc/comm.cu); `DistributedSolver` is the thin caller.  `HaloExchanger` is the
same neighbour exchange on torch tensors, kept for the CPU (gloo) tests of the partition logic
and for callers that run their own exchange through the split entry points.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np


class HaloExchanger:
    """Neighbour exchange of packed facet buffers.  Buffers are variable-fastest [slot][var]
    torch tensors (CPU for gloo, CUDA for NCCL); segment r of the send buffer goes to
    nbr_ranks[r], segment r of the recv buffer comes from it."""

    def __init__(self, nbr_ranks: List[int], send_counts: List[int], recv_counts: List[int], group=None):
        self.nbr, self.sc, self.rc, self.group = list(nbr_ranks), list(send_counts), list(recv_counts), group

    def start(self, send, recv, nvar: int):
        import torch.distributed as dist
        if not self.nbr:
            return []
        ops, so, ro = [], 0, 0
        me = dist.get_rank(self.group)
        for r, ns, nr in zip(self.nbr, self.sc, self.rc):
            s_seg, r_seg = send[so * nvar:(so + ns) * nvar], recv[ro * nvar:(ro + nr) * nvar]
            so, ro = so + ns, ro + nr
            if r == me:                       # periodic wrap onto ourselves (world size 1)
                r_seg.copy_(s_seg)
                continue
            ops.append(dist.P2POp(dist.isend, s_seg, r, self.group))
            ops.append(dist.P2POp(dist.irecv, r_seg, r, self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish(reqs):
        for q in reqs:
            q.wait()


def exchanger_from_mesh(mesh, group=None) -> HaloExchanger:
    sc = [int(s.size) for s in mesh.send_idx]
    offs = list(mesh.recv_off) + [mesh.n_ghost]
    rc = [offs[i + 1] - offs[i] for i in range(len(mesh.nbr_ranks))]
    return HaloExchanger(mesh.nbr_ranks, sc, rc, group)


def partition(mapP_1based, N_f: int, owner, n_parts: int, rank: int) -> dict:
    """Local view of `rank` of an element partition (sse_partition_*: host only, works for any mesh and any owner array):
    elem_gid (0-based here), mapP (0-based, local + ghost numbering, shape (n_local, N_f)), n_interior, n_ghost, nbr_ranks,
    send_count, recv_count, send_idx (0-based indices into the owned facet array)."""
    import ctypes as C
    from . import _lib
    L = _lib.load()
    mp = np.ascontiguousarray(np.asarray(mapP_1based).reshape(-1), dtype=np.int64)
    ow = np.ascontiguousarray(owner, dtype=np.int32)
    ne = ow.size
    h = C.c_void_p()
    p64, p32 = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    _lib.check(L.sse_partition_create(mp.ctypes.data_as(p64), ne, int(N_f), ow.ctypes.data_as(p32), int(n_parts), int(rank), C.byref(h)))
    try:
        nl, ni, ng, ns, nn = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        _lib.check(L.sse_partition_sizes(h, C.byref(nl), C.byref(ni), C.byref(ng), C.byref(nn), C.byref(ns)))
        gid = np.zeros(nl.value, dtype=np.int64)
        mpl = np.zeros(nl.value * N_f, dtype=np.int64)
        nbr = np.zeros(nn.value, dtype=np.int32)
        sc, rc = np.zeros(nn.value, dtype=np.int64), np.zeros(nn.value, dtype=np.int64)
        si = np.zeros(ns.value, dtype=np.int64)
        _lib.check(L.sse_partition_fill(h, gid.ctypes.data_as(p64), mpl.ctypes.data_as(p64), nbr.ctypes.data_as(p32),
                                        sc.ctypes.data_as(p64), rc.ctypes.data_as(p64), si.ctypes.data_as(p64)))
    finally:
        L.sse_partition_destroy(h)
    return {"elem_gid": gid - 1, "mapP": (mpl - 1).reshape(nl.value, N_f), "n_interior": int(ni.value), "n_ghost": int(ng.value),
            "nbr_ranks": nbr.tolist(), "send_count": sc.tolist(), "recv_count": rc.tolist(), "send_idx": si - 1}


class DistributedSolver:
    """Rank-local Solver of an element partition.  The exchange itself lives in the library (csrc/comm.cu: NCCL send/recv on a
    side stream inside sse_rhs / sse_rhs_lsrk / sse_step_ck54 / sse_rhs_host, all-reduce inside sse_functionals); this class
    only forms the communicator -- the NCCL unique id of rank 0 travels through torch.distributed, the host's own launcher --
    and hands the mesh's halo plan to the handle.  `solver` must be built from a partitioned mesh
    (uniform_periodic_mesh(..., part=(rank, world)))."""

    def __init__(self, solver, mesh, group=None):
        import torch.distributed as dist
        self.s, self.mesh, self.group = solver, mesh, group
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if world > 1:
            box = [solver.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            solver.comm_init(box[0], rank, world)
        solver.halo_plan(mesh)
        self.n_int = mesh.N_e - mesh.n_boundary

    def rhs(self, dudt, u, t: float = 0.0):
        return self.s.rhs(dudt, u, t)

    def rhs_host(self, dudt_host, u_host, t: float = 0.0, chunks: int = 0):
        """Residual on this rank's HOST buffers (pinned torch CPU tensors): one sse_rhs_host; the library uploads the
        halo-adjacent elements first so that the halos travel while the interior ranges are uploaded."""
        return self.s.rhs_host(dudt_host, u_host, t, chunks)

    def step_ck54(self, u, tmp, dudt, t, dt):
        return self.s.step_ck54(u, tmp, dudt, t, dt)

    def functionals(self, u, dudt):
        """Global conservation / energy / entropy residuals (all-reduced inside the library)."""
        return self.s.functionals(u, dudt)
