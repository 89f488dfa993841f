"""Subprocess of test_element_packing_is_bitwise_neutral: the residual hashes of a few small-element cases.  The packing knobs
(SSE_PACK_ROWS, SSE_THREADS_MIN, SSE_V_SMALL) are read once per process, hence a process per setting."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), os.path.join(ROOT, "oracle")]
import torch  # noqa: E402
from sse_b200 import cases  # noqa: E402
from sse_b200.solver import Solver  # noqa: E402

# element counts that leave a remainder for the one-row launch: 50, 18, 50 elements against 4 rows per CTA
CASES = {"euler_vortex_2d": lambda: cases.euler_vortex_2d(M=5, p=4, flux="lf"),
         "advection_diffusion_2d": lambda: cases.advection_diffusion_2d(M=3),
         "advection_2d": lambda: cases.advection_2d(M=5, flux="lf"),
         "euler_1d_like_standard": lambda: cases.euler_vortex_2d_standard(M=3, p=3, flux="lf")}
out = {}
for name, build in CASES.items():
    c = build()
    s = Solver(c.image(), 0)
    du = s.new_state()
    u = torch.from_numpy(c.u0(seed=0)).cuda()
    s.rhs(du, u)
    # an element range that starts and ends inside a CTA of the packed launch
    du2 = s.new_state()
    s.pass_a(u)
    if hasattr(s, "pass_aux"):
        s.pass_aux(du2, 0, c.sd.N_e)
    s.pass_b(du2, 3, c.sd.N_e - 5)
    s.synchronize()
    a, b = du.cpu().numpy(), du2.cpu().numpy()
    assert np.array_equal(a[3:c.sd.N_e - 2], b[3:c.sd.N_e - 2]), name
    out[name] = hashlib.sha1(a.tobytes()).hexdigest()
    s.close()
print(json.dumps(out))
