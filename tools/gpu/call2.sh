#!/bin/bash
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
( time python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
tail -15 $O/gputests.log
python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
tail -3 $O/bench.err
cat $O/bench.json
