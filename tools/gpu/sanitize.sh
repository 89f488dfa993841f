#!/bin/bash
mkdir -p gpurun_out/r2san
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 300 python tools/sanitize_cases.py > gpurun_out/r2san/sanitizer_$tool.log 2>&1
  echo "exit $?" >> gpurun_out/r2san/sanitizer_$tool.log
  tail -n 3 gpurun_out/r2san/sanitizer_$tool.log
done
