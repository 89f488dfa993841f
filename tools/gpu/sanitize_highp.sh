#!/bin/bash
O=gpurun_out/s5san; mkdir -p $O
for tool in racecheck memcheck synccheck; do
  timeout 55 compute-sanitizer --tool $tool --print-limit 100 python tools/sanitize_cases.py ct_p6_step ct_p7_step ct_std_p6 ct_std_p7 > $O/sanitizer_highp_$tool.log 2>&1
  echo "exit $?" >> $O/sanitizer_highp_$tool.log
  tail -n 3 $O/sanitizer_highp_$tool.log
done
