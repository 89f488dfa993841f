"""Generates the committed fixtures from the pinned oracle (the reference itself cannot run in this
image: no Julia).  Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), os.path.join(ROOT, "oracle")]
import oracle  # noqa: E402
from sse_b200 import cases  # noqa: E402

c = cases.euler_tgv_3d(M=2, flux="lf")
u = c.u0(seed=11)
du = oracle.rhs(c.image(), u)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "euler_tgv_3d_M2.npz"), u=u, dudt=du)
print("wrote euler_tgv_3d_M2.npz", u.shape, float(np.abs(du).max()))
