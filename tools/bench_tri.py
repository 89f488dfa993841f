#!/usr/bin/env python
"""BASELINE config 2 (2-D Euler vortex, p = 4 triangles, flux differencing) at 256 x 256 x 2 = 131 072 elements:
CUDA-event times of pass A / pass B and parity against the oracle on a small mesh, for the library selected by
SSE_B200_LIB (default: in-tree) and the path selected by SSE_TRI_CT (0: runtime tensor-line kernels).

    python tools/bench_tri.py [--M 256] [--p 4] [--steps 20]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), os.path.join(ROOT, "oracle")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=256)
    ap.add_argument("--p", type=int, default=4)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import torch
    import oracle
    from sse_b200 import cases
    from sse_b200.solver import Solver
    out = {"lib": os.path.relpath(os.environ.get("SSE_B200_LIB", "in-tree"), ROOT), "SSE_TRI_CT": os.environ.get("SSE_TRI_CT", "1")}
    c = cases.euler_vortex_2d(M=4, p=a.p, flux="lf")
    img, u = c.image(), c.u0(seed=0)
    s = Solver(img, 0)
    du = s.new_state()
    s.rhs(du, torch.from_numpy(u).cuda())
    s.synchronize()
    ref = oracle.rhs(img, u)
    out["parity_M4"] = float(np.abs(du.cpu().numpy() - ref).max() / np.abs(ref).max())
    out["variant"] = s.kernel_variant()
    s.close()
    c = cases.euler_vortex_2d(M=a.M, p=a.p, flux="lf")
    img = c.image()
    s = Solver(img, 0)
    s.use_current_stream()
    out.update({"elements": int(img.cfg.N_e), "dof": c.dof})
    # "noisy": the L2-projected vortex plus 2e-3 relative noise on every mode (the state of tools/bench_configs.py: most warps
    # leave the first tier of the log-mean); "smooth": the projected vortex itself (what bench.py uses for the headline config)
    for tag, seed in (("noisy", 0), ("smooth", None)):
        u, du = torch.from_numpy(c.u0(seed=seed)).cuda(), s.new_state()
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
        prof = s.profile_rhs(du, u, reps=10)
        out[tag] = {"ms_per_rhs": ms, "dof_per_s": c.dof / (ms * 1e-3), "kernel_ms_passA_aux_B1_B2": [round(float(x), 5) for x in prof],
                    "finite": bool(torch.isfinite(du).all())}
    s.close()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
