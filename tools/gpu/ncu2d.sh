#!/bin/bash
mkdir -p gpurun_out/r2q
ncu --set full --clock-control none --import-source on -c 2 --launch-skip 4 -o gpurun_out/r2q/prof2d -f python tools/profile_2d.py 128 euler > gpurun_out/r2q/ncu.log 2>&1
tail -n 5 gpurun_out/r2q/ncu.log
