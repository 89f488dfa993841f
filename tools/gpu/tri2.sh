#!/bin/bash
# triangle kernels: parity tests, A/B of library variants, ncu capture with source counters
O=gpurun_out/s4d; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q ) > $O/tri_tests.log 2>&1
tail -n 5 $O/tri_tests.log
python tools/bench_tri.py > $O/tri_bench.jsonl 2> $O/tri_bench.err
for v in cloud.jl_b200/lib/variants/libsse_b200_tri_*.so; do SSE_B200_LIB=$PWD/$v python tools/bench_tri.py >> $O/tri_bench.jsonl 2>> $O/tri_bench.err; done
cat $O/tri_bench.jsonl | cut -c1-400; tail -n 3 $O/tri_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tri" -s 4 -c 2 -o $O/tri_prof -f python tools/profile_2d.py 256 euler > $O/ncu.log 2>&1
tail -n 3 $O/ncu.log
