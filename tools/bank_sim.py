#!/usr/bin/env python
"""Shared-memory bank-conflict model for 64-bit accesses on sm_100a, as calibrated against ncu's per-instruction
`L1 Wavefronts Shared` of the compile-time kernels: a warp-wide LDS.64 / STS.64 is served in two half-warp wavefronts,
each of which splits further by the largest number of distinct addresses that fall on one of the 16 eight-byte banks.

    python tools/bank_sim.py            # the layouts used by kernels_ct.cuh and the alternatives that were searched

wavefronts(addrs): addrs = 32 double-word addresses (None for inactive lanes)."""
import itertools


def wavefronts(addrs):
    tot = 0
    for h in range(2):
        banks = {}
        for a in addrs[16 * h:16 * h + 16]:
            if a is not None:
                banks.setdefault(a % 16, set()).add(a)
        tot += max((len(s) for s in banks.values()), default=0)
    return tot


N = 5
GROUP_LANES = [divmod(lane, N) if lane // N < 6 else None for lane in range(32)]     # projection kernels: lane = (group, a3)


def partial_layout(RS, SA, SL, interleaved=False):
    """V' partials of one (element, variable) group at l * SL + a3 * SA, groups RS apart: average wavefronts of the stores
    (fixed l, lanes = (group, a3)) and of the reducer loads (lane a3 sums its 7 consecutive l)."""
    st = sum(wavefronts([None if x is None else x[0] * RS + x[1] * SA + l * SL for x in GROUP_LANES]) for l in range(35)) / 35
    ld = 0
    for q in range(7):
        for a in range(5):
            ld += wavefronts([None if x is None else x[0] * RS + a * SA + ((q * 5 + x[1]) if interleaved else (x[1] * 7 + q)) * SL
                              for x in GROUP_LANES])
    return st, ld / 35


def pair_kernel_partner_access(SA=25, SB=5, SC=1):
    """k_fluxdiff_ct: thread = volume node (a, b, c) at a * SA + b * SB + c * SC; wavefronts of reading the cyclic line
    partner (+1 / +2) in each direction and of the thread's own slot."""
    nodes = [(t // 25, (t // 5) % 5, t % 5) for t in range(125)]
    addr = lambda c: c[0] * SA + c[1] * SB + c[2] * SC
    out = []
    for d in range(3):
        for sh in (1, 2):
            tot = 0
            for w in range(4):
                pat = []
                for lane in range(32):
                    t = w * 32 + lane
                    if t >= 125:
                        pat.append(None)
                        continue
                    c = list(nodes[t])
                    c[d] = (c[d] + sh) % 5
                    pat.append(addr(c))
                tot += wavefronts(pat)
            out.append(tot / 4)
    own = sum(wavefronts([addr(nodes[w * 32 + l]) if w * 32 + l < 125 else None for l in range(32)]) for w in range(4)) / 4
    return out, own


def reducer_access(PS):
    """(facet node, variable) reducers of the pair kernel: thread t = (e, x, y) reads stage[e * 5 PS + x * PS + y + 5 i]."""
    tot = 0
    for w in range(4):
        pat = []
        for lane in range(32):
            t = w * 32 + lane
            if t >= 125:
                pat.append(None)
                continue
            e, rjj = divmod(t, 25)
            x, y = divmod(rjj, 5)
            pat.append(e * 5 * PS + x * PS + y)
        tot += wavefronts(pat)
    return tot / 4


if __name__ == "__main__":
    print("V' partials (stores, reducer loads):  l*5 + a3, stride 181 (first half of round 1):", partial_layout(181, 1, 5))
    print("                                      l*5 + 3 a3, stride 191 (now):               ", partial_layout(191, 3, 5))
    print("pair kernel, partner reads per direction/shift and own slot, natural layout:", pair_kernel_partner_access())
    best = min(((sum(r) + o, (sa, sb, sc)) for sa, sb, sc in itertools.product(range(1, 64), range(1, 32), range(1, 8))
                for r, o in [pair_kernel_partner_access(sa, sb, sc)]
                if len({a * sa + b * sb + c * sc for a in range(5) for b in range(5) for c in range(5)}) == 125), default=None)
    print("best linear padding (sum of the seven patterns, strides):", best)
    print("reducer loads: planes 25 apart", reducer_access(25), " planes 37 apart (now)", reducer_access(37))
