#!/usr/bin/env python
"""Per-kernel times of the 2-D BASELINE configurations at scale (p = 4 triangles, M x M x 2 elements); run under ncu for the
kernel metrics:   python tools/profile_2d.py [M] [config ...]     configs: advection, euler, advdiff"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200")]
import torch
from sse_b200 import cases
from sse_b200.solver import Solver
M = int(sys.argv[1]) if len(sys.argv) > 1 else 128
which = sys.argv[2:] or ["advection", "euler", "advdiff"]
build = {"advection": lambda: cases.advection_2d(M=M, flux="lf"), "euler": lambda: cases.euler_vortex_2d(M=M, p=4, flux="lf"),
         "advdiff": lambda: cases.advection_diffusion_2d(M=M)}
for name in which:
    c = build[name]()
    s = Solver(c.image(), 0)
    u, du = torch.from_numpy(c.u0(seed=0)).cuda(), s.new_state()
    for _ in range(2):
        s.rhs(du, u)
    print(name, "elements", c.sd.N_e, "N_q", int(s.cfg.N_q), "N_f", int(s.cfg.N_f), "variant", s.kernel_variant(),
          "ms [pass A, aux, pass B, -]:", [round(float(x), 4) for x in s.profile_rhs(du, u, reps=5)], flush=True)
    s.close()
