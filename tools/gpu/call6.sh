#!/bin/bash
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
cat $O/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline_rhs']['frac'], {k:round(v['ms'],3) for k,v in d['roofline']['kernels'].items()}, d['e2e']['ms_per_step'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_nodal_ct|k_fluxdiff_ct|k_project_ct' -s 9 -c 3 \
    -o $O/r2d_prof -f python bench.py --cells 16 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > $O/ncu_headline.log 2>&1
tail -2 $O/ncu_headline.log | cut -c1-200
