#!/bin/bash
O=gpurun_out/s5c; mkdir -p $O
SSE_STAGE_OPS=${1:-0} timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_aux_physical|k_time_physical|k_nodal_generic" --launch-skip 9 -c 3 -o $O/prof_c3 -f python tools/profile_2d.py 256 advdiff > $O/ncu.log 2>&1
tail -n 3 $O/ncu.log
