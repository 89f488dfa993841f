#!/bin/bash
# config-4 A/B over the library variants (fused path only)
mkdir -p gpurun_out/ab
: > gpurun_out/ab/$1.jsonl
for lib in cloud.jl_b200/lib/libsse_b200.so cloud.jl_b200/lib/variants/*.so; do
  SSE_C4_FUSED_ONLY=1 SSE_B200_LIB=$PWD/$lib timeout 300 python tools/bench_config4.py ${2:-32} 20 2>> gpurun_out/ab/$1.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%-45s ms %.4f frac %.3f kernels %s parity %.1e' % ('$lib'[-45:], d['ms_per_rhs'], d['roofline_frac'], [round(x,4) for x in d['kernel_ms']], d['parity p4 lf']))
    d['lib']='$lib'; open('gpurun_out/ab/$1.jsonl','a').write(json.dumps(d)+'\n')
"
done
