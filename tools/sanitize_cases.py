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
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(FAMILIES)
    for n in names:
        FAMILIES[n]()
    print("sanitize_cases: all families ran")
