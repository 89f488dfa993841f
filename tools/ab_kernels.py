#!/usr/bin/env python
"""A/B of kernel variants: one line per library with parity against the oracle (small meshes) and the
CUDA-event times of pass A / pass B of the headline configuration on a mid-size mesh.

    python tools/ab_kernels.py [--cells 24] [--steps 20] [lib.so ...]

Without arguments it runs the in-tree library and every cloud.jl_b200/lib/variants/*.so (built by
tools/build_variant.sh).  Each library runs in its own process (the path is bound at import)."""
import argparse
import glob
import hashlib
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), os.path.join(ROOT, "oracle"), ROOT]


def one(a):
    import torch
    import oracle
    from sse_b200 import cases
    from sse_b200.solver import Solver
    out = {"lib": os.path.relpath(os.environ.get("SSE_B200_LIB", "in-tree"), ROOT)}
    for name, M, flux in (("tgv_M2_ec", 2, "ec"), ("tgv_M4_lf", 4, "lf")) + ((("tgv_M8_lf", 8, "lf"),) if a.m8 else ()):
        c = cases.euler_tgv_3d(M=M, flux=flux)
        img, u = c.image(), c.u0(seed=0)
        s = Solver(img, 0)
        du = s.new_state()
        s.rhs(du, torch.from_numpy(u).cuda())
        s.synchronize()
        got, ref = du.cpu().numpy(), oracle.rhs(img, u)
        out[name] = float(np.abs(got - ref).max() / np.abs(ref).max())
        out[name + "_sha"] = hashlib.sha1(got.tobytes()).hexdigest()[:10]
        s.close()
    if a.config4:
        c = cases.advection_3d(M=4, flux="central")
        img, u = c.image(), c.u0(seed=0)
        s = Solver(img, 0)
        du = s.new_state()
        s.rhs(du, torch.from_numpy(u).cuda())
        s.synchronize()
        got, ref = du.cpu().numpy(), oracle.rhs(img, u)
        out["adv3d_M4"] = float(np.abs(got - ref).max() / np.abs(ref).max())
        s.close()
    import bench
    case, u0 = bench.build_case(a.cells, "lf", None, device=0)
    s = Solver(case.image(), 0)
    s.use_current_stream()
    u, du = torch.from_numpy(u0).cuda(), s.new_state()
    n_e = case.sd.N_e
    for _ in range(3):
        s.pass_a(u)
        s.pass_b(du, 0, n_e)
    torch.cuda.synchronize()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(a.steps)]
    for e in ev:
        e[0].record()
        s.pass_a(u)
        e[1].record()
        s.pass_b(du, 0, n_e)
        e[2].record()
    torch.cuda.synchronize()
    out["elements"] = n_e
    out["pass_a_ms"] = float(np.median([e[0].elapsed_time(e[1]) for e in ev]))
    out["pass_b_ms"] = float(np.median([e[1].elapsed_time(e[2]) for e in ev]))
    out["rhs_ms"] = float(np.median([e[0].elapsed_time(e[2]) for e in ev]))
    try:
        k = s.profile_rhs(du, u, reps=a.steps)
        out["k_nodal_ms"], out["k_pair_ms"], out["k_project_ms"] = float(k[0]), float(k[2]), float(k[3])
    except Exception as e:       # older library variants have no sse_profile_rhs
        out["profile_error"] = str(e)[:80]
    f = s.functionals(u, du)
    out["conservation"] = float(np.abs(np.asarray(f[:5])).max())
    out["du_sha"] = hashlib.sha1(du.cpu().numpy().tobytes()).hexdigest()[:10]
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--cells", type=int, default=24)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--config4", action="store_true")
    ap.add_argument("--one", action="store_true")
    ap.add_argument("--m8", action="store_true", help="parity on the 3 072-element mesh as well")
    a = ap.parse_args()
    if a.one:
        return one(a)
    libs = a.libs or ([os.path.join(ROOT, "cloud.jl_b200", "lib", "libsse_b200.so")] +
                      sorted(glob.glob(os.path.join(ROOT, "cloud.jl_b200", "lib", "variants", "*.so"))))
    for lib in libs:
        env = dict(os.environ, SSE_B200_LIB=os.path.abspath(lib))
        cmd = [sys.executable, os.path.abspath(__file__), "--one", "--cells", str(a.cells), "--steps", str(a.steps)]
        if a.config4:
            cmd.append("--config4")
        if a.m8:
            cmd.append("--m8")
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
        if r.returncode != 0:
            print(json.dumps({"lib": os.path.relpath(lib, ROOT), "error": r.stderr[-800:]}), flush=True)
        else:
            print(r.stdout.strip().splitlines()[-1], flush=True)


if __name__ == "__main__":
    main()
