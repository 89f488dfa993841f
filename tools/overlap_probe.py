#!/usr/bin/env python
"""Experiment: run pass A of later element chunks concurrently with pass B of earlier ones (two streams).

    python tools/overlap_probe.py [--cells 24] [--chunks 6]

Pass B of a chunk needs pass A of every chunk holding one of its face neighbours (read from mapP)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "cloud.jl_b200")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=24)
    ap.add_argument("--chunks", type=int, default=6)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    import torch
    from sse_b200 import cases
    from sse_b200.solver import Solver
    c = cases.euler_tgv_3d(M=a.cells, flux="lf")
    img, u0 = c.image(), c.u0(seed=0)
    s = Solver(img, 0)
    ne, nf = s.state_shape[0], int(s.cfg.N_f)
    u, du, du2 = torch.from_numpy(u0).cuda(), s.new_state(), s.new_state()
    C = a.chunks
    bounds = [ne * k // C for k in range(C + 1)]
    owner = np.searchsorted(np.asarray(bounds[1:]), np.arange(ne), side="right")
    nb = (np.asarray(img.arrays["mapP"]).reshape(-1) - 1) // nf          # neighbour element of every facet node
    deps = [sorted(set(owner[nb[bounds[k] * nf:bounds[k + 1] * nf]].tolist())) for k in range(C)]
    print("deps", deps)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()

    def set_stream(st):
        with torch.cuda.stream(st):
            s.use_current_stream()

    def plain():
        set_stream(sa)
        with torch.cuda.stream(sa):
            s.rhs(du, u)

    def overlapped():
        evs = []
        set_stream(sa)
        with torch.cuda.stream(sa):
            for k in range(C):
                s.pass_a_range(u, bounds[k], bounds[k + 1] - bounds[k])
                e = torch.cuda.Event()
                e.record(sa)
                evs.append(e)
        set_stream(sb)
        order = sorted(range(C), key=lambda k: max(deps[k]))
        with torch.cuda.stream(sb):
            for k in order:
                for d in deps[k]:
                    sb.wait_event(evs[d])
                s.pass_b(du2, bounds[k], bounds[k + 1] - bounds[k])
        sa.wait_stream(sb)

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(sa)
        for _ in range(a.steps):
            fn()
        e1.record(sa)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    tp = timeit(plain)
    to = timeit(overlapped)
    print(f"plain {tp:.3f} ms   overlapped({C} chunks) {to:.3f} ms   max diff {float((du - du2).abs().max()):.3e}")


if __name__ == "__main__":
    main()
