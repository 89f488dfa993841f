#!/usr/bin/env python
"""RHS throughput of every BASELINE.json configuration on one GPU (the headline config is bench.py's job).

    python tools/bench_configs.py [--big]

Prints one JSON line per configuration: DOF/s, ms per RHS, kernel variant used, max relative difference
against the oracle when the mesh is small enough to run it."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), os.path.join(ROOT, "oracle")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="also run the 3-D configs at ~10^5 elements")
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import torch
    from sse_b200 import cases
    from sse_b200.solver import Solver
    runs = [("config1 advection_2d p4 32x32", lambda: cases.advection_2d(M=32, flux="lf"), True),
            ("config2 euler_vortex_2d p4 32x32", lambda: cases.euler_vortex_2d(M=32, p=4, flux="lf"), True),
            ("config2 euler_vortex_2d p3 32x32", lambda: cases.euler_vortex_2d(M=32, p=3, flux="lf"), True),
            ("config3 advection_diffusion_2d p4 32x32", lambda: cases.advection_diffusion_2d(M=32), True),
            ("config4 advection_3d p4 M=8", lambda: cases.advection_3d(M=8, flux="central"), True),
            ("config5 euler_tgv_3d p4 M=8", lambda: cases.euler_tgv_3d(M=8, flux="lf"), True)]
    # the reference's own notebook configurations: dense multidimensional operators through the generic kernels
    runs += [("euler_3d.ipynb: Euler 3-D ModalMulti p=3 tets M=4 (reference: 31.8 ms/RHS, 1 thread)",
              lambda: cases.euler_tgv_3d(M=4, p=3, flux="ec", kind="modal_multi"), True),
             ("euler_vortex_2d.ipynb: Euler 2-D ModalMulti p=4 tris M=4 (reference: 541 us/RHS)",
              lambda: cases.euler_vortex_2d(M=4, p=4, flux="lf", kind="modal_multi"), True),
             ("advection_3d.ipynb: advection 3-D ModalTensor p=7 tets M=2 (reference: 7.64 ms/RHS)",
              lambda: cases.advection_3d(M=2, p=7, flux="lf"), True)]
    if a.big:
        runs += [("Euler 3-D ModalMulti p=3 tets M=16 (dense operators, generic kernels)",
                  lambda: cases.euler_tgv_3d(M=16, p=3, flux="ec", kind="modal_multi"), False)]
    if a.big:
        runs += [("config1 advection_2d p4 256x256", lambda: cases.advection_2d(M=256, flux="lf"), False),
                 ("config2 euler_vortex_2d p4 256x256", lambda: cases.euler_vortex_2d(M=256, p=4, flux="lf"), False),
                 ("config3 advection_diffusion_2d p4 256x256", lambda: cases.advection_diffusion_2d(M=256), False)]
    if a.big:
        runs += [("config4 advection_3d p4 M=32", lambda: cases.advection_3d(M=32, flux="central"), False),
                 ("config5 euler_tgv_3d p4 M=24", lambda: cases.euler_tgv_3d(M=24, flux="lf"), False)]
    for name, build, check in runs:
        t0 = time.time()
        c = build()
        img, u0 = c.image(), c.u0(seed=0)
        s = Solver(img, 0)
        s.use_current_stream()
        u, du = torch.from_numpy(u0).cuda(), s.new_state()
        for _ in range(3):
            s.rhs(du, u)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            s.rhs(du, u)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        out = {"config": name, "elements": c.sd.N_e, "dof": c.dof, "ms_per_rhs": ms, "dof_per_s": c.dof / (ms * 1e-3),
               "kernel_variant": s.kernel_variant(), "setup_s": round(time.time() - t0, 1),
               "kernel_ms_passA_aux_B1_B2": [round(float(x), 5) for x in s.profile_rhs(du, u)]}
        if check:
            import oracle
            ref = oracle.rhs(img, u0)
            out["max_rel_diff_vs_oracle"] = float(np.abs(du.cpu().numpy() - ref).max() / np.abs(ref).max())
        print(json.dumps(out), flush=True)
        s.close()


if __name__ == "__main__":
    main()
