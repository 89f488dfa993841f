#!/bin/bash
# triangle kernels, tuned defaults: all GPU tests, config-2 timing (noisy and smooth state, both paths), every BASELINE config, ncu extract
O=gpurun_out/s4e; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
tail -n 4 $O/gputests.log
python tools/bench_tri.py > $O/tri_bench.jsonl 2> $O/tri_bench.err
SSE_TRI_CT=0 python tools/bench_tri.py >> $O/tri_bench.jsonl 2>> $O/tri_bench.err
python tools/bench_tri.py --p 3 >> $O/tri_bench.jsonl 2>> $O/tri_bench.err
python tools/bench_tri.py --M 32 >> $O/tri_bench.jsonl 2>> $O/tri_bench.err
cat $O/tri_bench.jsonl | cut -c1-700; tail -n 3 $O/tri_bench.err
python tools/bench_configs.py --big > $O/configs.jsonl 2> $O/configs.err
python - <<PY
import json
for l in open("$O/configs.jsonl"):
    d=json.loads(l); print(d["config"][:60], d["elements"], round(d["ms_per_rhs"],4), "%.3e"%d["dof_per_s"], d["kernel_variant"], d.get("max_rel_diff_vs_oracle"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tri" -s 4 -c 2 -o $O/tri_prof -f python tools/profile_2d.py 256 euler > $O/ncu.log 2>&1
tail -n 2 $O/ncu.log
