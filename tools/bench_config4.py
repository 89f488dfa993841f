#!/usr/bin/env python
"""BASELINE config 4 (3-D linear advection, StandardForm, ModalTensor p = 4 curved tets) on one GPU: parity against the oracle
on a small mesh, then the time of one residual at M^3 x 6 elements for the fused two-kernel path and the three-kernel path
(SSE_ADV_FUSED=0), with the HBM roofline fraction (algorithmic 14 560 B per element).

    python tools/bench_config4.py [M] [steps]        one JSON line per path"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), os.path.join(ROOT, "oracle")]
ALG_BYTES = 14560.0


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    import torch
    import oracle
    from sse_b200 import cases
    from sse_b200.solver import Solver
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    small = [(cases.advection_3d(M=4, flux=f, p=p), f"p{p} {f}") for p, f in ((4, "lf"), (4, "central"), (3, "lf"))]
    big = cases.advection_3d(M=M, flux="lf")
    img_big, u_big = big.image(), big.u0(seed=0)
    for fused in ((1,) if os.environ.get("SSE_C4_FUSED_ONLY") else (1, 0)):
        os.environ["SSE_ADV_FUSED"] = str(fused)
        out = {"path": "fused (k_adv_facets_ct + k_adv_fused_ct)" if fused else "three kernels (k_nodal_ct + k_standard_adv_ct + k_project_ct)"}
        for c, name in small:
            img, u = c.image(), c.u0(seed=0)
            s = Solver(img, 0)
            du = s.new_state()
            ud = torch.from_numpy(u).cuda()
            s.rhs(du, ud)
            s.synchronize()
            ref = oracle.rhs(img, u)
            out["parity " + name] = float(np.abs(du.cpu().numpy() - ref).max() / np.abs(ref).max())
            s.close()
        s = Solver(img_big, 0)
        s.use_current_stream()
        u, du = torch.from_numpy(u_big).cuda(), s.new_state()
        for _ in range(3):
            s.rhs(du, u)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = s.launches
        e0.record()
        for _ in range(steps):
            s.rhs(du, u)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        k = s.profile_rhs(du, u, reps=10)
        f = s.functionals(u, du)
        ne = big.sd.N_e
        out.update({"elements": ne, "dof": big.dof, "ms_per_rhs": ms, "dof_per_s": big.dof / (ms * 1e-3),
                    "launches_per_rhs": (s.launches - l0) / steps, "kernel_ms": [float(x) for x in k],
                    "hbm_algorithmic_gbs": ALG_BYTES * ne / (ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm,
                    "roofline_frac": ALG_BYTES * ne / (ms * 1e-3) / 1e9 / hbm,
                    "conservation": float(abs(f[0]) / max(float(du.abs().max()), 1e-300))})
        print(json.dumps(out), flush=True)
        s.close()


if __name__ == "__main__":
    main()
