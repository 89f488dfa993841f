#!/usr/bin/env python
"""Attribute the per-SASS-instruction counters of an ncu report to CUDA source lines.

    ncu -i prof.ncu-rep --page source --csv > prof_src.csv
    python tools/ncu_lines.py prof_src.csv <mangled kernel name> [top]

Uses nvdisasm --print-line-info on the in-tree libsse_b200.so (built with -lineinfo); the SASS
listing of the report and of the cubin are the same instruction sequence."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("SSE_LIB", os.path.join(ROOT, "cloud.jl_b200", "lib", "libsse_b200.so"))


def sass_with_lines(kernel):
    with tempfile.TemporaryDirectory() as td:
        subprocess.check_call(["cuobjdump", "-xelf", "all", LIB], cwd=td, stdout=subprocess.DEVNULL)
        seqs = []
        for cubin in sorted(f for f in os.listdir(td) if f.endswith(".cubin")):
            txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cubin)], capture_output=True,
                                 text=True).stdout.split("\n")
            start = [i for i, l in enumerate(txt) if l.startswith(".text." + kernel + ":")]
            if not start:
                continue
            f = ln = None
            seq = []
            for l in txt[start[0] + 1:]:
                if l.startswith("//---------------------"):
                    break
                m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
                if m:
                    f, ln = os.path.basename(m.group(1)), int(m.group(2))
                    continue
                m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
                if m:
                    seq.append((f, ln, m.group(2)))
            seqs.append(seq)
        return seqs[0] if seqs else []


def main():
    src_csv, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(src_csv)))
    hdr, data = rows[1], rows[2:]
    ie, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    seq = sass_with_lines(kernel)
    assert len(seq) == len(data), (len(seq), len(data))
    inst, smp, pipe = collections.Counter(), collections.Counter(), collections.Counter()
    for (f, ln, s), r in zip(seq, data):
        c = int(r[ie] or 0)
        inst[(f, ln)] += c
        smp[(f, ln)] += int(r[ismp] or 0)
        op = s.split()[0] if not s.startswith("@") else s.split()[1]
        pipe[op.split(".")[0]] += c
    tot, tots = sum(inst.values()), max(sum(smp.values()), 1)
    print(f"total warp instructions {tot}")
    print("opcode mix:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in pipe.most_common(14)))
    cache = {}
    for (f, ln), c in inst.most_common(top):
        p = os.path.join(ROOT, "cloud.jl_b200", "csrc", f or "")
        if p not in cache:
            cache[p] = open(p).read().split("\n") if os.path.isfile(p) else []
        text = cache[p][ln - 1].strip()[:100] if ln and ln <= len(cache[p]) else ""
        print(f"{100 * c / tot:5.1f}% inst {100 * smp[(f, ln)] / tots:5.1f}% smp  {f}:{ln}  {text}")


if __name__ == "__main__":
    main()
