#!/bin/bash
mkdir -p gpurun_out/r2h
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2h/gputests.log 2>&1
grep -E "passed|failed|Error" gpurun_out/r2h/gputests.log | tail -n 3
python tools/bench_graph.py > gpurun_out/r2h/graph.jsonl 2> gpurun_out/r2h/graph.err
cat gpurun_out/r2h/graph.jsonl; tail -n 3 gpurun_out/r2h/graph.err
