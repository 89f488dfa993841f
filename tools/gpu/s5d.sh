#!/bin/bash
O=gpurun_out/s5d; mkdir -p $O
( time python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
grep -n "passed\|failed" $O/gputests.log
timeout 400 python tools/bench_highp.py --ab --M 10 > $O/highp_ab.jsonl 2> $O/highp_ab.err
python - <<PY
import json
for l in open("$O/highp_ab.jsonl"):
    d=json.loads(l); print(d["lib"], d["p"], d["elements"], "%.4f ms"%d["ms_per_rhs"], "%.1e"%d["parity_M2"])
PY
tail -n 3 $O/highp_ab.err
