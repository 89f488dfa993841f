"""Element-partitioned residual over N GPUs: one process per GPU, facet-trace halos only.

The reference has no distributed path; the only cross-element coupling of
`semi_discrete_residual!` is the gather of neighbour facet states through `mesh.mapP`
between the two element loops (Solvers.jl:505-511), so the exchange step is: pack the cut
faces' `u_f` after pass A, point-to-point send/recv to the slab neighbours (NCCL over
NVLink via torch.distributed; gloo in the CPU tests), unpack into the ghost slots, run
pass B.  Pass B on interior elements is enqueued before the receive completes so the
transfer hides behind it.
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


class DistributedSolver:
    """Rank-local Solver + halo exchange.  `solver` must be built from a partitioned mesh
    (uniform_periodic_mesh(..., part=(rank, world)))."""

    def __init__(self, solver, mesh, group=None):
        import torch
        self.s, self.mesh, self.group = solver, mesh, group
        send_idx = np.concatenate(mesh.send_idx) if mesh.send_idx else np.zeros(0, dtype=np.int64)
        solver.halo_configure(send_idx + 1)
        self.ex = exchanger_from_mesh(mesh, group)
        self.second = bool(solver.image.law.second_order)
        self.n_int = mesh.N_e - mesh.n_boundary
        self.comm_stream = torch.cuda.Stream(device=solver.device)
        self.nc = int(solver.cfg.N_c)
        self.d = int(solver.cfg.d)

    def _exchange(self, which: int):
        """pack on the compute stream, send/recv on the comm stream; returns an event that marks
        ghost data ready in the recv staging buffer."""
        import torch
        s = self.s
        s.halo_pack(which)
        send, recv = s.halo_buffers(which)
        cur = torch.cuda.current_stream(s.device)
        packed = torch.cuda.Event()
        packed.record(cur)
        with torch.cuda.stream(self.comm_stream):
            self.comm_stream.wait_event(packed)
            reqs = self.ex.start(send, recv, self.nc * (self.d if which else 1))
            HaloExchanger.finish(reqs)
            done = torch.cuda.Event()
            done.record(self.comm_stream)
        return done

    def rhs(self, dudt, u, t: float = 0.0):
        import torch
        s = self.s
        cur = torch.cuda.current_stream(s.device)
        s.pass_a(u)
        done = self._exchange(0)
        if self.second:
            s.pass_aux(dudt, 0, self.n_int)
            cur.wait_event(done)
            s.halo_unpack(0)
            s.pass_aux(dudt, self.n_int, self.mesh.n_boundary)
            done = self._exchange(1)
            s.pass_b(dudt, 0, self.n_int)
            cur.wait_event(done)
            s.halo_unpack(1)
            s.pass_b(dudt, self.n_int, self.mesh.n_boundary)
        else:
            s.pass_b(dudt, 0, self.n_int)                 # interior elements overlap the transfer
            cur.wait_event(done)
            s.halo_unpack(0)
            s.pass_b(dudt, self.n_int, self.mesh.n_boundary)
        return dudt
