#!/bin/bash
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
( time python -m pytest tests -m gpu -x -q ) > $O/gputests.log 2>&1
tail -5 $O/gputests.log
python bench.py --workload advection_3d --steps 20 --warmup 3 > $O/bench_c4.json 2> $O/bench_c4.err
tail -2 $O/bench_c4.err; cut -c1-1500 $O/bench_c4.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_adv_facets_ct|k_adv_fused_ct' -s 6 -c 2 \
    -o $O/r2f_prof_c4 -f python tools/profile_config4.py 32 > $O/ncu_config4.log 2>&1
tail -2 $O/ncu_config4.log
