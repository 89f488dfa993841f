#!/bin/bash
# A/B of run-time knobs of the in-tree library: tools/gpu/ab_env.sh TAG CELLS "ENV1=a ENV2=b" "ENV1=c" ...   (one run per setting)
tag=$1; cells=$2; shift 2
mkdir -p gpurun_out/ab
: > gpurun_out/ab/$tag.jsonl
for setting in "" "$@"; do
  echo "# $setting" >> gpurun_out/ab/$tag.jsonl
  env $setting python tools/ab_kernels.py --cells $cells --steps 20 cloud.jl_b200/lib/libsse_b200.so >> gpurun_out/ab/$tag.jsonl 2>> gpurun_out/ab/$tag.err
done
python - <<PY
import json
for l in open("gpurun_out/ab/$tag.jsonl"):
    if l.startswith("#"): print(l.strip()); continue
    d=json.loads(l)
    if "error" in d: print(d["lib"], "ERROR", d["error"][-300:]); continue
    print("  par %.1e %.1e  A %.4f  B %.4f  rhs %.4f | nodal %.4f pair %.4f proj %.4f  sha %s" % (d["tgv_M2_ec"], d["tgv_M4_lf"], d["pass_a_ms"], d["pass_b_ms"], d["rhs_ms"], d.get("k_nodal_ms",0), d.get("k_pair_ms",0), d.get("k_project_ms",0), d["du_sha"]))
PY
