#!/bin/bash
O=gpurun_out/s5a; mkdir -p $O
( timeout 600 python -m pytest tests -m gpu -x -q -k "p6 or p7 or cover_p2 or p5" ) > $O/gputests_highp.log 2>&1
tail -n 8 $O/gputests_highp.log
timeout 600 python tools/bench_highp.py --M 8 > $O/highp.jsonl 2> $O/highp.err; cat $O/highp.jsonl; tail -3 $O/highp.err
