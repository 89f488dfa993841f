"""The element-partitioned residual inside the library (csrc/comm.cu) driven from ONE process, as a single-process Julia
host would drive it: N partitions = N handles, sse_rhs_multi / sse_step_ck54_multi.

* `local`: sse_comm_init_local -- peer copies, no NCCL; two (three) partitions may share one GPU, so the interior/halo
  split, the pack / unpack lists, the ghost numbering of mapP and the BR1 double exchange are covered on a 1-GPU box;
* `nccl`: sse_comm_init_all (ncclCommInitAll), one partition per GPU; needs as many GPUs as partitions.

Reference semantics: the partitioned result must equal the single-domain oracle on the gathered elements (1e-12), because the
only cross-element coupling of semi_discrete_residual! is the mapP gather between its element loops (Solvers.jl:505-511)."""
import numpy as np
import pytest
import torch

import oracle
from sse_b200 import cases
from sse_b200.solver import Solver, solve_ck54, ODEProblem, semi_discrete_residual

pytestmark = pytest.mark.gpu
RTOL = 1.0e-12

CASES = {
    "euler_tgv_3d": dict(M=4, flux="lf"),
    "advection_3d": dict(M=4, flux="lf"),
    "euler_vortex_2d": dict(M=6, flux="ec"),
    "advection_diffusion_2d": dict(M=6),           # BR1: two exchanges per residual
}


def _partitions(name, kw, world, mode):
    ndev = torch.cuda.device_count()
    if mode == "nccl" and ndev < world:
        pytest.skip(f"needs {world} GPUs")
    full = cases.BUILDERS[name](**kw)
    u_full = full.u0(seed=0)
    parts, solvers, us, dus = [], [], [], []
    for r in range(world):
        dev = r if mode == "nccl" else r % ndev
        part = cases.BUILDERS[name](part=(r, world), **kw)
        s = Solver(part.image(), dev)
        parts.append(part)
        solvers.append(s)
        us.append(torch.from_numpy(np.ascontiguousarray(u_full[part.sd.mesh.elem_gid])).to(f"cuda:{dev}"))
        dus.append(s.new_state())
    (Solver.comm_init_all if mode == "nccl" else Solver.comm_init_local)(solvers)
    for s, part in zip(solvers, parts):
        s.halo_plan(part.sd.mesh)
    return full, u_full, parts, solvers, us, dus


def _gather(full_like, parts, xs):
    out = np.empty_like(full_like)
    for part, x in zip(parts, xs):
        out[part.sd.mesh.elem_gid] = x.cpu().numpy()
    return out


@pytest.mark.parametrize("mode", ["local", "nccl"])
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", sorted(CASES))
def test_partitioned_residual_matches_single_domain_oracle(name, world, mode):
    kw = dict(CASES[name])
    if world == 3:
        if name in ("euler_tgv_3d", "advection_3d"):
            pytest.skip("3-D cases are partitioned in two (M = 4)")
    full, u_full, parts, solvers, us, dus = _partitions(name, kw, world, mode)
    ref = oracle.rhs(full.image(), u_full)
    for _ in range(2):                          # twice: the second call reuses ghost buffers that are still being read
        Solver.rhs_multi(solvers, dus, us)
    for s in solvers:
        s.synchronize()
    got = _gather(ref, parts, dus)
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    assert all(p.sd.mesh.n_ghost > 0 and 0 < p.sd.mesh.n_boundary for p in parts)
    for s in solvers:
        s.close()
    assert err <= RTOL, (name, world, mode, err)


@pytest.mark.parametrize("mode", ["local", "nccl"])
def test_partitioned_ck54_step_matches_single_gpu(mode):
    """Three fused CarpenterKennedy2N54 steps on two partitions == the same steps on one handle (test_driver.jl:77-83)."""
    name, kw = "euler_tgv_3d", dict(M=4, flux="ec")
    full, u_full, parts, solvers, us, dus = _partitions(name, kw, 2, mode)
    tmps = [s.new_state() for s in solvers]
    dt, t = 2.0e-3, 0.0
    for _ in range(3):
        Solver.step_ck54_multi(solvers, us, tmps, dus, t, dt)
        t += dt
    for s in solvers:
        s.synchronize()
    got = _gather(u_full, parts, us)
    one = Solver(full.image(), 0)
    ref = solve_ck54(ODEProblem(semi_discrete_residual, u_full, (0.0, 3 * dt), one), dt, 3)
    one.close()
    for s in solvers:
        s.close()
    assert float(np.abs(got - ref).max() / np.abs(ref).max()) <= RTOL


def test_halo_plan_is_checked():
    """An interior range that reads ghosts, receive counts that do not add up to N_ghost and a second communicator are
    refused (SSE_ERR_BAD_ARGUMENT), and a ghosted handle without a plan refuses sse_rhs (SSE_ERR_COMM)."""
    import copy
    from sse_b200._lib import SSEError
    parts = [cases.euler_vortex_2d(M=4, flux="lf", part=(r, 2)) for r in range(2)]
    solvers = [Solver(p.image(), 0) for p in parts]
    s, mesh = solvers[0], parts[0].sd.mesh
    u, du = s.new_state(), s.new_state()
    with pytest.raises(SSEError) as e:
        s.rhs(du, u)
    assert e.value.code == 5
    Solver.comm_init_local(solvers)
    with pytest.raises(SSEError):
        Solver.comm_init_local(solvers)
    bad = copy.copy(mesh)
    bad.n_boundary = 0                                # claims every element is interior
    with pytest.raises(SSEError) as e:
        s.halo_plan(bad)
    assert e.value.code == 1 and "ghost" in str(e.value)
    bad = copy.copy(mesh)
    bad.n_ghost = mesh.n_ghost - 1
    bad.recv_off = list(mesh.recv_off)
    with pytest.raises(SSEError):
        s.halo_plan(bad)
    with pytest.raises(SSEError) as e:                # handles of a local communicator are driven together
        s.halo_plan(mesh)
        s.rhs(du, u)
    assert e.value.code == 5
    for x in solvers:
        x.close()
