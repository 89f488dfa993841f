#!/bin/bash
O=gpurun_out/s4f; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
tail -n 12 $O/gputests.log
python tools/bench_configs.py --big > $O/configs.jsonl 2> $O/configs.err
SSE_TRI_CT=0 python tools/bench_configs.py --big > $O/configs_tri0.jsonl 2>> $O/configs.err
python - <<PY
import json
for f in ("$O/configs.jsonl", "$O/configs_tri0.jsonl"):
  for l in open(f):
    d=json.loads(l)
    if "2d" in d["config"] or "2-D" in d["config"]: print(d["config"][:60], d["elements"], round(d["ms_per_rhs"],4), "%.3e"%d["dof_per_s"], d["kernel_variant"], d.get("max_rel_diff_vs_oracle"), d.get("kernel_ms_passA_aux_B1_B2"))
PY
tail -n 3 $O/configs.err
