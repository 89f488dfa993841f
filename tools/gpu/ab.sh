#!/bin/bash
# A/B of every library under cloud.jl_b200/lib/variants against the in-tree one ($1 = tag, $2 = cells)
mkdir -p gpurun_out/ab
python tools/ab_kernels.py --cells ${2:-24} --steps 20 --m8 > gpurun_out/ab/$1.jsonl 2> gpurun_out/ab/$1.err
python - <<PY
import json
for l in open("gpurun_out/ab/$1.jsonl"):
    d=json.loads(l)
    if "error" in d: print(d["lib"], "ERROR", d["error"][-300:]); continue
    print("%-50s par %.1e %.1e %.1e  A %.4f  B %.4f  rhs %.4f | nodal %.4f pair %.4f proj %.4f  sha %s" % (d["lib"][-50:], d["tgv_M2_ec"], d["tgv_M4_lf"], d.get("tgv_M8_lf", 0), d["pass_a_ms"], d["pass_b_ms"], d["rhs_ms"], d.get("k_nodal_ms",0), d.get("k_pair_ms",0), d.get("k_project_ms",0), d["du_sha"]))
PY
