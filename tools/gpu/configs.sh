#!/bin/bash
mkdir -p gpurun_out/r2c2
python tools/bench_configs.py --big > gpurun_out/r2c2/configs.jsonl 2> gpurun_out/r2c2/configs.err
python - <<PY
import json
for l in open("gpurun_out/r2c2/configs.jsonl"):
    d=json.loads(l); print(d["config"], d["elements"], round(d["ms_per_rhs"],4), "%.3e"%d["dof_per_s"], d["kernel_variant"], d.get("max_rel_diff_vs_oracle"))
PY
tail -n 3 gpurun_out/r2c2/configs.err
