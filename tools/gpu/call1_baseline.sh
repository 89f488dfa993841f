#!/bin/bash
# round 2, GPU call 1: state of the round-1 kernels on this round's box — GPU tests, ncu capture with source, sanitizer
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
( time python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
# ncu: full set + source for the three headline kernels (24 576 elements) and the config-4 kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_nodal_ct|k_fluxdiff_ct|k_project_ct' -s 9 -c 3 \
    -o $O/r2a_prof -f python bench.py --cells 16 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu_headline.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_nodal_ct|k_standard_adv_ct|k_project_ct' -s 6 -c 3 \
    -o $O/r2a_prof_c4 -f python tools/profile_config4.py 16 > $O/ncu_config4.log 2>&1
# compute-sanitizer on one case per kernel family
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > $O/sanitizer_$tool.log 2>&1
  echo "exit $?" >> $O/sanitizer_$tool.log
done
tail -3 $O/gputests.log; cat $O/bench.json | cut -c1-400; tail -5 $O/sanitizer_*.log
