#!/usr/bin/env python
"""Launch-latency-bound meshes (the reference's 2-D examples, BASELINE configs 1-3): time of one device-resident
CarpenterKennedy2N54 step with and without CUDA-graph replay (sse_set_graph_mode).  One JSON line per configuration."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200")]
import torch  # noqa: E402
from sse_b200 import cases  # noqa: E402
from sse_b200.solver import Solver  # noqa: E402

for name, c in (("config1 advection_2d 32x32", cases.advection_2d(M=32, flux="lf")),
                ("config2 euler_vortex_2d 32x32", cases.euler_vortex_2d(M=32, flux="lf")),
                ("config3 advection_diffusion_2d 32x32", cases.advection_diffusion_2d(M=32)),
                ("config5 euler_tgv_3d M=4", cases.euler_tgv_3d(M=4, flux="lf"))):
    img, u0 = c.image(), c.u0(seed=0)
    out = {"config": name, "elements": c.sd.N_e, "dof": c.dof}
    for graph in (0, 1):
        s = Solver(img, 0)
        s.use_current_stream()
        s.set_graph_mode(bool(graph))
        u, tmp, du = torch.from_numpy(u0).cuda(), s.new_state(), s.new_state()
        for _ in range(5):
            s.step_ck54(u, tmp, du, 0.0, 1e-5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = s.launches
        e0.record()
        for _ in range(200):
            s.step_ck54(u, tmp, du, 0.0, 1e-5)
        e1.record()
        torch.cuda.synchronize()
        out["us_per_step_graph" if graph else "us_per_step_launches"] = 1e3 * e0.elapsed_time(e1) / 200
        out["launches_per_step"] = (s.launches - l0) / 200
        s.close()
    out["speedup"] = out["us_per_step_launches"] / out["us_per_step_graph"]
    print(json.dumps(out), flush=True)
