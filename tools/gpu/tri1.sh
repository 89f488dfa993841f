#!/bin/bash
# first GPU run of the warp-per-element triangle kernels: parity tests, then timing against the tensor-line path
O=gpurun_out/s4a; mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_tri.py -x -q ) > $O/tri_tests.log 2>&1
tail -n 25 $O/tri_tests.log
python tools/bench_tri.py > $O/tri_bench.jsonl 2> $O/tri_bench.err
SSE_TRI_CT=0 python tools/bench_tri.py >> $O/tri_bench.jsonl 2>> $O/tri_bench.err
cat $O/tri_bench.jsonl; tail -n 3 $O/tri_bench.err
