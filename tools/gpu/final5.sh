#!/bin/bash
# end-of-round evidence on one B200: GPU tests, smoke(), the default bench line, the reference arm, the ncu launch list
mkdir -p gpurun_out/s5z
O=gpurun_out/s5z
( time python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
tail -n 4 $O/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -n 2 $O/smoke.log
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err
python bench.py > $O/bench.json 2> $O/bench.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bench.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["value"], d["roofline_rhs"]["frac"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["gpu_launches"], d["clocks"])
r=json.loads([l for l in open("$O/bench_ref.json") if l.startswith("{")][-1]); print("reference arm", r["value"], r["cpu_baseline"]["cores"], r["ms_per_step"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/bench_under_ncu.log 2>&1
grep -c "k_" $O/launches.csv
python tools/bench_configs.py > $O/configs.jsonl 2> $O/configs.err; tail -n 2 $O/configs.err
timeout 300 python tools/bench_highp.py --M 8 > $O/highp.jsonl 2> $O/highp.err; tail -n 2 $O/highp.err
