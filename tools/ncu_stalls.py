#!/usr/bin/env python
"""Per-source-line warp-stall reasons and shared-memory wavefronts of one kernel from an ncu report.

    ncu -i prof.ncu-rep --page source --csv -k regex:<kernel> > src.csv      # keep the first kernel block only
    python tools/ncu_stalls.py src.csv <mangled kernel name> [top]

Companion of tools/ncu_lines.py (same join of the report's SASS rows with `nvdisasm --print-line-info` of the library
the report was taken with: set SSE_LIB if that is not the in-tree one).  Prints (1) the share of every stall reason,
(2) the lines with the most stall samples and their top reasons, (3) the shared-memory instructions by wavefronts,
with wavefronts per instruction against the ideal (2 for a conflict-free 64-bit access of a full warp)."""
import collections
import csv
import sys

import ncu_lines


def main():
    src_csv, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(open(src_csv)))
    hdr, data = rows[1], rows[2:]
    nxt = [i for i, r in enumerate(data) if r and r[0] == "Kernel Name"]
    if nxt:
        data = data[:nxt[0]]
    seq = ncu_lines.sass_with_lines(kernel)
    assert len(seq) == len(data), f"SASS of the library ({len(seq)}) and of the report ({len(data)}) differ: set SSE_LIB"
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    ie, iw, ii = hdr.index("Instructions Executed"), hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal")
    per, tot = collections.defaultdict(collections.Counter), collections.Counter()
    wf = collections.defaultdict(lambda: [0, 0, 0])
    for (f, ln, s), r in zip(seq, data):
        for i, h in stall_cols:
            v = int(r[i] or 0)
            per[(f, ln)][h] += v
            tot[h] += v
        w = int(r[iw] or 0)
        if w:
            op = s.split()[1] if s.startswith("@") else s.split()[0]
            k = (f, ln, op)
            wf[k][0] += int(r[ie] or 0)
            wf[k][1] += w
            wf[k][2] += int(r[ii] or 0)
    T = max(sum(tot.values()), 1)
    print("stall samples", T, {k.replace("stall_", ""): round(100 * v / T, 1) for k, v in tot.most_common(10)})
    for (f, ln), c in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:top]:
        print(f"{100 * sum(c.values()) / T:5.1f}%  {f}:{ln}  ", {k.replace('stall_', ''): round(100 * v / T, 1) for k, v in c.most_common(4)})
    tw = max(sum(v[1] for v in wf.values()), 1)
    print("shared-memory wavefronts", tw)
    for k, v in sorted(wf.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100 * v[1] / tw:5.1f}%  {k[0]}:{k[1]} {k[2]:8s} inst {v[0]:9d} wavefronts {v[1]:9d} ({v[1] / max(v[0], 1):.2f}/inst, ideal {v[2] / max(v[0], 1):.2f})")


if __name__ == "__main__":
    main()
