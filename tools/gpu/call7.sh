#!/bin/bash
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 600 python tools/bench_config4.py 32 20 > $O/config4.jsonl 2> $O/config4.err
tail -3 $O/config4.err
python - <<PY
import json
for l in open("$O/config4.jsonl"):
    d=json.loads(l); print({k:(round(v,4) if isinstance(v,float) else v) for k,v in d.items() if k not in ("dof","elements","hbm_peak_gbs")})
PY
( time python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
tail -6 $O/gputests.log
