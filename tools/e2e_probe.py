#!/usr/bin/env python
"""End-to-end (host-buffer) residual time against the number of pipeline chunks of Solver.rhs_host, full-size config.

    python tools/e2e_probe.py [--cells 56] [--chunks 8 16 32 48 64]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200"), ROOT]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=56)
    ap.add_argument("--chunks", type=int, nargs="*", default=[8, 16, 24, 32, 48, 64])
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    import torch
    import bench
    from sse_b200.solver import Solver
    case, u0 = bench.build_case(a.cells, "lf", None, device=0)
    s = Solver(case.image(), 0)
    s.use_current_stream()
    hu = torch.from_numpy(u0).pin_memory()
    hdu = torch.empty_like(hu).pin_memory()
    u, du = torch.from_numpy(u0).cuda(), s.new_state()
    for _ in range(2):
        s.rhs(du, u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        s.rhs(du, u)
    e1.record()
    torch.cuda.synchronize()
    out = {"elements": case.sd.N_e, "resident_ms": e0.elapsed_time(e1) / a.steps, "e2e_ms": {}}
    for c in a.chunks:
        s.rhs_host(hdu, hu, chunks=c)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.steps):
            s.rhs_host(hdu, hu, chunks=c)
        e1.record()
        torch.cuda.synchronize()
        out["e2e_ms"][c] = e0.elapsed_time(e1) / a.steps
    out["equal"] = bool(torch.equal(hdu, du.cpu()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
