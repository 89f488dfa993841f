#!/bin/bash
# element packing of the generic / tensor-line kernels: GPU parity suite, then the BASELINE configs with and without packing
mkdir -p gpurun_out/r2p
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2p/gputests.log 2>&1
tail -n 3 gpurun_out/r2p/gputests.log
for s in "SSE_PACK_ROWS=8" "SSE_PACK_ROWS=1 SSE_THREADS_MIN=64" "SSE_PACK_ROWS=1" "SSE_PACK_ROWS=4 SSE_THREADS_MIN=64"; do
  echo "# $s"
  env $s python tools/bench_configs.py --big 2> gpurun_out/r2p/configs.err | tee -a gpurun_out/r2p/configs_all.jsonl | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l)
    if '2d' in d['config'] or 'ModalMulti p=4' in d['config']: print('  ', d['config'][:45], d['elements'], round(d['ms_per_rhs'],4), '%.3e'%d['dof_per_s'], d['kernel_variant'], d.get('max_rel_diff_vs_oracle'), d['kernel_ms_passA_aux_B1_B2'])
"
done
