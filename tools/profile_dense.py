#!/usr/bin/env python
"""Per-kernel times of the dense multidimensional Euler scheme (ModalMulti p = 3 tets) through the generic kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200")]
import torch
from sse_b200 import cases
from sse_b200.solver import Solver
M = int(sys.argv[1]) if len(sys.argv) > 1 else 16
for kind, seed in (("modal_multi", 0), ("modal_multi", None), ("nodal_multi", 0)):
    c = cases.euler_tgv_3d(M=M, p=3, flux="ec", kind=kind)
    s = Solver(c.image(), 0)
    u, du = torch.from_numpy(c.u0(seed=seed)).cuda(), s.new_state()
    for _ in range(2):
        s.rhs(du, u)
    print(kind, "noisy state" if seed is not None else "smooth state", "elements", c.sd.N_e, "N_q", int(s.cfg.N_q), "N_f", int(s.cfg.N_f), "variant", s.kernel_variant(),
          "ms [pass A, aux, pass B, -]:", [round(float(x), 3) for x in s.profile_rhs(du, u, reps=5)], flush=True)
    s.close()
