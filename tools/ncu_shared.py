#!/usr/bin/env python
"""Shared-memory wavefronts of one kernel of an ncu source page, per source line (joined with the -lineinfo SASS of the
in-tree library like tools/ncu_lines.py):

    ncu -i prof.ncu-rep --page source --csv > prof_src.csv
    python tools/ncu_shared.py prof_src.csv <first csv line of the kernel's section> <mangled kernel name>

Prints, per source line, the LDS/STS warp instructions, their wavefronts, the ideal wavefronts and the global-load
instructions (LDG also pass through the LSU data pipe)."""
import collections
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_lines import sass_with_lines, ROOT  # noqa: E402


def main():
    src_csv, first, kernel = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[first]                       # the line after "Kernel Name"
    data = []
    for r in rows[first + 1:]:
        if r and r[0] == "Kernel Name":
            break
        data.append(r)
    ix = {h: i for i, h in enumerate(hdr)}
    seq = sass_with_lines(kernel)
    assert len(seq) == len(data), (len(seq), len(data))
    agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
    for (f, ln, s), r in zip(seq, data):
        op = s.split()[0] if not s.startswith("@") else s.split()[1]
        n = int(r[ix["Instructions Executed"]] or 0)
        a = agg[(f, ln)]
        if op.startswith("LDS") or op.startswith("STS"):
            a[0] += n
            a[1] += int(r[ix["L1 Wavefronts Shared"]] or 0)
            a[2] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
        elif op.startswith("LDG") or op.startswith("STG") or op.startswith("LDGSTS"):
            a[3] += n
            a[4] += int(r[ix["L1 Tag Requests Global"]] or 0)
    tw = sum(a[1] for a in agg.values())
    ti = sum(a[2] for a in agg.values())
    print(f"shared wavefronts {tw}, ideal {ti} ({100 * ti / max(tw, 1):.1f}%), shared instructions {sum(a[0] for a in agg.values())}, "
          f"global instructions {sum(a[3] for a in agg.values())}, global tag requests {sum(a[4] for a in agg.values())}")
    cache = {}
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1] - kv[1][4]):
        if not (a[0] or a[3]):
            continue
        p = os.path.join(ROOT, "cloud.jl_b200", "csrc", f or "")
        if p not in cache:
            cache[p] = open(p).read().split("\n") if os.path.isfile(p) else []
        text = cache[p][ln - 1].strip()[:90] if ln and ln <= len(cache[p]) else ""
        print(f"{100 * a[1] / max(tw, 1):5.1f}% wf  {a[0]:9d} inst {a[1] / max(a[0], 1):5.2f} wf/inst (ideal {a[2] / max(a[0], 1):4.2f})  glob {a[3]:8d}/{a[4]:9d}  {f}:{ln}  {text}")


if __name__ == "__main__":
    main()
