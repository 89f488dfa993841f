#!/usr/bin/env python
"""Run a few residuals of BASELINE config 4 (3-D advection, StandardForm, p = 4 tets) for profiling under ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200")]
import torch  # noqa: E402
from sse_b200 import cases  # noqa: E402
from sse_b200.solver import Solver  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16
c = cases.advection_3d(M=M, flux="central")
s = Solver(c.image(), 0)
u, du = torch.from_numpy(c.u0(seed=0)).cuda(), s.new_state()
for _ in range(4):
    s.rhs(du, u)
torch.cuda.synchronize()
print("elements", c.sd.N_e, "variant", s.kernel_variant())
