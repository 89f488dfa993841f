#!/usr/bin/env python
"""One small residual per kernel family, for compute-sanitizer (memcheck / racecheck / initcheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_cases.py [family ...]

Families: ct (compile-time Euler flux differencing on tets), ct_std (compile-time 3-D advection StandardForm),
tensor (runtime tensor-line kernels, 2-D Euler), generic (row-wise kernels, forced with variant 0), br1 (second-order
PhysicalOperators), step (the fused device-resident CarpenterKennedy2N54 step), functionals, geometry.
Every case is also compared with the oracle, so a sanitizer run doubles as a parity run (the sanitizer only instruments
the CUDA library; the oracle is the checker)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
import oracle  # noqa: E402
from sse_b200 import cases  # noqa: E402
from sse_b200.solver import Solver  # noqa: E402


def run(name, case, variant=1, step=False):
    img, u = case.image(), case.u0(seed=0)
    s = Solver(img, 0)
    s.set_kernel_variant(variant)
    du = s.new_state()
    ud = torch.from_numpy(u).cuda()
    s.rhs(du, ud)
    s.synchronize()
    ref = oracle.rhs(img, u)
    err = float(np.abs(du.cpu().numpy() - ref).max() / np.abs(ref).max())
    extra = ""
    if step:
        tmp = s.new_state()
        s.step_ck54(ud, tmp, du, 0.0, 1e-4)
        s.synchronize()
        f = s.functionals(ud, du)
        extra = f" ck54 step ok, functionals {np.array2string(np.asarray(f), precision=2)}"
    print(f"{name}: variant {s.kernel_variant()}, {case.sd.N_e} elements, rel. diff vs oracle {err:.2e}{extra}", flush=True)
    s.close()
    assert err <= 1e-12, (name, err)


FAMILIES = {
    "ct": lambda: run("ct euler_tgv_3d p4", cases.euler_tgv_3d(M=2, flux="lf")),
    "ct_p3": lambda: run("ct euler_tgv_3d p3", cases.euler_tgv_3d(M=2, p=3, flux="ec")),
    "ct_std": lambda: run("ct_std advection_3d p4", cases.advection_3d(M=2, flux="lf")),
    "tensor": lambda: run("tensor euler_vortex_2d p4", cases.euler_vortex_2d(M=3, flux="lf")),
    "generic": lambda: run("generic euler_tgv_3d p3", cases.euler_tgv_3d(M=2, p=3, flux="lf"), variant=0),
    "generic2d": lambda: run("generic advection_2d", cases.advection_2d(M=3, flux="lf"), variant=0),
    "br1": lambda: run("br1 advection_diffusion_2d", cases.advection_diffusion_2d(M=3)),
    "step": lambda: run("step euler_tgv_3d p4", cases.euler_tgv_3d(M=2, flux="ec"), step=True),
    # round 2: fused advection path at the other compile-time sizes, the stage-fused CK54 kernel at p = 2 / 5, the dense
    # all-pairs kernel (modal and nodal multidimensional schemes), Euler under StandardForm, and the partitioned residual
    "ct_std_p2": lambda: run("ct_std advection_3d p2", cases.advection_3d(M=2, p=2, flux="lf"), step=True),
    "ct_std_p5": lambda: run("ct_std advection_3d p5", cases.advection_3d(M=2, p=5, flux="central")),
    "ct_p2_step": lambda: run("ct euler_tgv_3d p2", cases.euler_tgv_3d(M=2, p=2, flux="lf"), step=True),
    "ct_p5": lambda: run("ct euler_tgv_3d p5", cases.euler_tgv_3d(M=2, p=5, flux="ec")),
    # session 5: the compile-time kernels at N = 7, 8 (other shared-memory plans, one or two CTAs per SM, 352 / 512-thread pair kernel)
    "ct_p6_step": lambda: run("ct euler_tgv_3d p6", cases.euler_tgv_3d(M=2, p=6, flux="lf"), step=True),
    "ct_p7_step": lambda: run("ct euler_tgv_3d p7", cases.euler_tgv_3d(M=2, p=7, flux="ec"), step=True),
    "ct_std_p6": lambda: run("ct_std advection_3d p6", cases.advection_3d(M=2, p=6, flux="central")),
    "ct_std_p7": lambda: run("ct_std advection_3d p7", cases.advection_3d(M=2, p=7, flux="lf"), step=True),
    "dense": lambda: run("dense euler_tgv_3d ModalMulti p2", cases.euler_tgv_3d(M=2, p=2, flux="ec", kind="modal_multi")),
    "dense_nodal": lambda: run("dense euler_vortex_2d NodalMulti p3", cases.euler_vortex_2d(M=3, p=3, flux="lf", kind="nodal_multi")),
    "standard_euler": lambda: run("generic euler_vortex_2d StandardForm", cases.euler_vortex_2d_standard(M=3, p=3, flux="lf")),
    "multi": lambda: run_multi(),
}


def run_multi():
    """Two partitions on one GPU, halos by peer copies (sse_comm_init_local + sse_rhs_multi), against the single-domain oracle."""
    full = cases.euler_tgv_3d(M=4, p=3, flux="lf")
    u_full = full.u0(seed=0)
    parts = [cases.euler_tgv_3d(M=4, p=3, flux="lf", part=(r, 2)) for r in range(2)]
    solvers = [Solver(p.image(), 0) for p in parts]
    Solver.comm_init_local(solvers)
    for s, p in zip(solvers, parts):
        s.halo_plan(p.sd.mesh)
    us = [torch.from_numpy(np.ascontiguousarray(u_full[p.sd.mesh.elem_gid])).cuda() for p in parts]
    dus = [s.new_state() for s in solvers]
    for _ in range(2):
        Solver.rhs_multi(solvers, dus, us)
    for s in solvers:
        s.synchronize()
    ref = oracle.rhs(full.image(), u_full)
    out = np.empty_like(ref)
    for p, d in zip(parts, dus):
        out[p.sd.mesh.elem_gid] = d.cpu().numpy()
    err = float(np.abs(out - ref).max() / np.abs(ref).max())
    print(f"multi (2 partitions, peer copies): rel. diff vs oracle {err:.2e}", flush=True)
    for s in solvers:
        s.close()
    assert err <= 1e-12

if __name__ == "__main__":
    names = sys.argv[1:] or list(FAMILIES)
    for n in names:
        FAMILIES[n]()
    print("sanitize_cases: all families ran")
