#!/bin/bash
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
( time python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
tail -5 $O/gputests.log
python tools/bench_ck54.py 56 4 > $O/ck54.jsonl 2> $O/ck54.err
cat $O/ck54.jsonl; tail -2 $O/ck54.err
