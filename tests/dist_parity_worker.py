"""torchrun worker: element-partitioned residual on WORLD_SIZE GPUs vs the single-domain oracle.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_parity_worker.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), os.path.join(ROOT, "oracle")]
from sse_b200 import cases  # noqa: E402
from sse_b200.dist import DistributedSolver  # noqa: E402
from sse_b200.solver import Solver  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    worst = 0.0
    M3 = 4 if world <= 4 else 8                     # the slab partition needs M divisible by the number of ranks
    for name, kw in (("euler_tgv_3d", dict(M=M3, flux="lf")), ("advection_3d", dict(M=M3, flux="lf")),
                     ("advection_diffusion_2d", dict(M=8)), ("euler_vortex_2d", dict(M=8, flux="ec"))):
        full = cases.BUILDERS[name](**kw)
        u_full = full.u0(seed=0)
        part = cases.BUILDERS[name](part=(rank, world), **kw)
        gid = part.sd.mesh.elem_gid
        s = Solver(part.image(), local)
        s.use_current_stream()
        ds = DistributedSolver(s, part.sd.mesh)
        u = torch.from_numpy(np.ascontiguousarray(u_full[gid])).cuda()
        du = s.new_state()
        for _ in range(2):
            ds.rhs(du, u)
        torch.cuda.synchronize()
        # the pipelined host-buffer call of every rank returns the device residual bit for bit
        hu = u.cpu().pin_memory()
        hdu = torch.empty_like(hu).pin_memory()
        for chunks in (2, 5):
            hdu.zero_()
            ds.rhs_host(hdu, hu, chunks=chunks)
            assert torch.equal(hdu, du.cpu()), (name, rank, chunks)
        # the functionals are all-reduced inside the library: every rank holds the single-domain totals
        fun = ds.functionals(u, du)
        got = [None] * world
        dist.all_gather_object(got, (gid, du.cpu().numpy()))
        if rank == 0:
            import oracle
            ref = oracle.rhs(full.image(), u_full)
            out = np.empty_like(ref)
            for g, d in got:
                out[g] = d
            err = float(np.abs(out - ref).max() / np.abs(ref).max())
            worst = max(worst, err)
            one = Solver(full.image(), local)
            du1 = one.new_state()
            u1 = torch.from_numpy(u_full).cuda()
            one.rhs(du1, u1)
            fun1 = one.functionals(u1, du1)
            one.close()
            scale = float(np.abs(ref).max()) * float(np.prod([b - a for a, b in full.sd.mesh.limits]))
            ferr = float(np.abs(fun - fun1).max() / scale)
            worst = max(worst, ferr if ferr > 1e-11 else 0.0)
            print(f"dist parity {name} world={world}: max rel diff {err:.3e}, functionals (all-reduced) vs single GPU {ferr:.1e} "
                  f"(ghost facets {part.sd.mesh.n_ghost}, boundary elements {part.sd.mesh.n_boundary}/{part.sd.N_e}, "
                  f"NCCL {s.comm_info()[2]})", flush=True)
        s.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        assert worst <= 1e-12, worst
        print("DIST PARITY OK", flush=True)


if __name__ == "__main__":
    main()
