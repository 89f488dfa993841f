#!/bin/bash
mkdir -p gpurun_out/r2p
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2p/gputests.log 2>&1
tail -n 3 gpurun_out/r2p/gputests.log
python tools/bench_configs.py --big 2> gpurun_out/r2p/configs.err | tee gpurun_out/r2p/configs.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    print('  ', d['config'][:45], d['elements'], round(d['ms_per_rhs'],4), '%.3e'%d['dof_per_s'], d['kernel_variant'], d.get('max_rel_diff_vs_oracle'), d['kernel_ms_passA_aux_B1_B2'])
"
