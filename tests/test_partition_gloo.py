"""Multi-rank halo logic on CPU: world_size 2 (and 4) with the gloo backend.  Every rank builds its
slab of the periodic mesh, fills its owned facet array with a function of the facet-node
coordinates, exchanges the cut faces, and checks that the gather through the rewritten mapP
(local + ghost numbering) sees exactly the neighbour's coincident node."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sse_b200.dist import exchanger_from_mesh
from sse_b200.mesh import ChanWarping, uniform_periodic_mesh
from sse_b200.reference import ModalTensor, reference_approximation


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, elem, d, M, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ra = reference_approximation(ModalTensor(3), elem, mapping_degree=3)
        L = 1.0
        mesh = uniform_periodic_mesh(ra, ((0.0, L),) * d, (M,) * d, ChanWarping(1 / 16, (L,) * d), part=(rank, world))
        nvar = 2
        # "facet state" = smooth periodic functions of the facet-node coordinates
        f = [np.sin(2 * np.pi * sum((m + 1) * mesh.xyzf[m] for m in range(d)) / L),
             np.cos(2 * np.pi * mesh.xyzf[0] / L) * np.cos(2 * np.pi * mesh.xyzf[d - 1] / L)]
        owned = np.stack([a.reshape(-1) for a in f], axis=0)                   # (nvar, N_f*N_e)
        facet = np.concatenate([owned, np.full((nvar, mesh.n_ghost), np.nan)], axis=1)
        send_idx = np.concatenate(mesh.send_idx) if mesh.send_idx else np.zeros(0, dtype=np.int64)
        send = torch.from_numpy(np.ascontiguousarray(owned[:, send_idx].T).reshape(-1))   # [slot][var]
        recv = torch.empty(mesh.n_ghost * nvar, dtype=torch.float64)
        ex = exchanger_from_mesh(mesh)
        ex.finish(ex.start(send, recv, nvar))
        facet[:, owned.shape[1]:] = recv.numpy().reshape(mesh.n_ghost, nvar).T
        got = facet[:, mesh.mapP.reshape(-1)]
        err = float(np.abs(got - owned).max())                                  # neighbour value == own value
        tot = torch.tensor([float(mesh.N_e), float(mesh.n_boundary)])
        dist.all_reduce(tot)
        if rank == 0:
            out.put((err, tot.tolist(), mesh.n_ghost))
        else:
            assert err < 1e-12, err
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("elem,d,M,world", [("Tri", 2, 4, 2), ("Tet", 3, 4, 2), ("Tet", 3, 4, 4)])
def test_halo_exchange_gloo(elem, d, M, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, elem, d, M, q)) for r in range(world)]
    for p in procs:
        p.start()
    err, tot, nghost = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert err < 1e-12
    assert int(tot[0]) == (2 if d == 2 else 6) * M ** d and nghost > 0
