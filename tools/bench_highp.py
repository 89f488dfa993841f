#!/usr/bin/env python
"""p = 6, 7 tetrahedra (the reference's examples/advection_3d.ipynb runs ModalTensor(7)): ms per RHS of the compile-time
kernels against the run-time kernels they replace (SSE_CT_NMAX=6 in a second process), parity against the oracle on M = 2.

    python tools/bench_highp.py [--M 8]"""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), os.path.join(ROOT, "oracle")]


def one(a):
    import torch
    import oracle
    from sse_b200 import cases
    from sse_b200.solver import Solver
    runs = (("advection_3d", lambda M, p: cases.advection_3d(M=M, p=p, flux="lf")),
            ("euler_tgv_3d", lambda M, p: cases.euler_tgv_3d(M=M, p=p, flux="ec")))
    for name, mk in (runs[1:] if a.ab else runs):
        for p in ((5, 6, 7) if a.ab else (6, 7)):
            c = mk(2, p)
            img, u = c.image(), c.u0(seed=0)
            s = Solver(img, 0)
            du = s.new_state()
            s.rhs(du, torch.from_numpy(u).cuda())
            s.synchronize()
            ref = oracle.rhs(img, u)
            par = float(np.abs(du.cpu().numpy() - ref).max() / np.abs(ref).max())
            s.close()
            c = mk(a.M, p)
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
            print(json.dumps({"case": name, "p": p, "elements": int(c.sd.N_e), "variant": s.kernel_variant(), "parity_M2": par,
                              "ms_per_rhs": ms, "dof_per_s": c.dof / (ms * 1e-3),
                              "ct_nmax": os.environ.get("SSE_CT_NMAX", "8"),
                              "lib": os.path.basename(os.environ.get("SSE_B200_LIB", "in-tree"))}), flush=True)
            s.close()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--M", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--ab", action="store_true", help="Euler p = 5, 6, 7 with the in-tree library and every lib/variants/*.so (launch shapes)")
    a = ap.parse_args()
    if a.child:
        one(a)
    elif a.ab:
        import glob
        for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "cloud.jl_b200", "lib", "variants", "*.so"))):
            env = dict(os.environ)
            if lib:
                env["SSE_B200_LIB"] = lib
            subprocess.run([sys.executable, __file__, "--child", "--ab", "--M", str(a.M), "--steps", str(a.steps)], env=env, check=False)
    else:
        for nmax in ("8", "6"):
            env = dict(os.environ, SSE_CT_NMAX=nmax)
            subprocess.run([sys.executable, __file__, "--child", "--M", str(a.M), "--steps", str(a.steps)], env=env, check=False)
