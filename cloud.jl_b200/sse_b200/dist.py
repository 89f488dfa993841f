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

    def rhs_host(self, dudt_host, u_host, t: float = 0.0, chunks: int = 8):
        """Residual on this rank's HOST buffers (pinned torch CPU tensors): the boundary elements are uploaded and put
        through pass A first so that the facet halos travel while the interior ranges are uploaded; pass B of an
        interior range starts as soon as pass A has covered its face neighbours (mapP) and its dudt is downloaded while
        later ranges are still arriving; the boundary elements finish after the halo has been unpacked."""
        import torch
        from .solver import range_plan
        s, ne, nb, n_int = self.s, self.mesh.N_e, self.mesh.n_boundary, self.n_int
        if self.second or nb == 0 or n_int < 4 * chunks:
            d_u, d_du = self._host_state()
            d_u.copy_(u_host, non_blocking=True)
            self.rhs(d_du, d_u, t)
            dudt_host.copy_(d_du, non_blocking=True)
            torch.cuda.current_stream(s.device).synchronize()
            return dudt_host
        d_u, d_du = self._host_state()
        if getattr(self, "_plan", None) is None or self._plan[0] != chunks:
            ranges = [(n_int, ne)] + [(n_int * c // chunks, n_int * (c + 1) // chunks) for c in range(chunks)]
            self._plan = (chunks, ranges, range_plan(s.image.arrays["mapP"], ne, int(s.cfg.N_f), ranges))
            self._copy_in, self._copy_out = torch.cuda.Stream(device=s.device), torch.cuda.Stream(device=s.device)
        _, ranges, after = self._plan
        cur = torch.cuda.current_stream(s.device)
        cin, cout = self._copy_in, self._copy_out
        cin.wait_stream(cur)
        cout.wait_stream(cur)

        def pass_b_and_download(a, b):
            s.pass_b(d_du, a, b - a)
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(cout):
                cout.wait_event(ev)
                dudt_host[a:b].copy_(d_du[a:b], non_blocking=True)

        done = None
        for i, (a, b) in enumerate(ranges):
            with torch.cuda.stream(cin):
                d_u[a:b].copy_(u_host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cin)
            cur.wait_event(ev)
            s.pass_a_range(d_u, a, b - a)
            if i == 0:
                done = self._exchange(0)                 # boundary facets are complete: halos travel from here on
            for k in after[i]:
                if k:
                    pass_b_and_download(*ranges[k])
        cur.wait_event(done)
        s.halo_unpack(0)
        pass_b_and_download(*ranges[0])
        cur.wait_stream(cout)
        cur.wait_stream(cin)
        cur.synchronize()
        return dudt_host

    def _host_state(self):
        if getattr(self, "_hs", None) is None:
            self._hs = (self.s.new_state(), self.s.new_state())
        return self._hs
