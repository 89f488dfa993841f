#!/usr/bin/env python
"""Device-resident CarpenterKennedy2N54 on the headline configuration: time per step (5 stages) of sse_step_ck54 with the
stage-fused kernels (default) and without (SSE_CK54_FUSED=0: residual + stage update per stage), and the difference of the
two states after a few steps.  One JSON line.

    python tools/bench_ck54.py [cells] [steps]"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), ROOT]


def one(cells, steps):
    import torch
    import bench
    from sse_b200.solver import Solver
    case, u0 = bench.build_case(cells, "lf", None, device=0)
    s = Solver(case.image(), 0)
    s.use_current_stream()
    u, tmp, du = torch.from_numpy(u0).cuda(), s.new_state(), s.new_state()
    dt = 1e-4
    for _ in range(2):
        s.step_ck54(u, tmp, du, 0.0, dt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = s.launches
    e0.record()
    for _ in range(steps):
        s.step_ck54(u, tmp, du, 0.0, dt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    import hashlib
    print(json.dumps({"fused": os.environ.get("SSE_CK54_FUSED", "1"), "elements": case.sd.N_e, "ms_per_step": ms,
                      "ms_per_stage": ms / 5, "launches_per_step": (s.launches - l0) / steps,
                      "dof_per_s_per_stage": case.dof / (ms / 5 * 1e-3),
                      "u_sha": hashlib.sha1(u.cpu().numpy().tobytes()).hexdigest()[:12], "u_absmax": float(u.abs().max())}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(int(sys.argv[2]), int(sys.argv[3]))
    else:
        cells = int(sys.argv[1]) if len(sys.argv) > 1 else 56
        steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
        for f in ("1", "0"):
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one", str(cells), str(steps)],
                           env=dict(os.environ, SSE_CK54_FUSED=f), check=False)
