#!/bin/bash
# config 3: launch-bounds A/B of the PhysicalOperators kernels (variants built with -DSSE_PHYS_MINB=...); the subset of GPU tests
# that exercises them runs with the best variant when it beats the in-tree library
O=gpurun_out/s5f; mkdir -p $O
: > $O/ab.log
for lib in in-tree cloud.jl_b200/lib/variants/libsse_b200_pm10.so cloud.jl_b200/lib/variants/libsse_b200_pm12.so cloud.jl_b200/lib/variants/libsse_b200_pm16.so; do
  if [ $lib = in-tree ]; then r=$(python tools/profile_2d.py 256 advdiff 2>&1 | tail -n 1); else r=$(SSE_B200_LIB=$PWD/$lib python tools/profile_2d.py 256 advdiff 2>&1 | tail -n 1); fi
  echo "$lib $r" >> $O/ab.log
done
cat $O/ab.log
best=$(python - <<PY
import re
best=None
for l in open("$O/ab.log"):
    m=re.search(r"\[([0-9., ]+)\]\s*$", l)
    if not m: continue
    t=sum(float(x) for x in m.group(1).split(",")[:3])
    if best is None or t<best[0]: best=(t,l.split()[0])
print(best[1])
PY
)
echo "best: $best"
if [ "$best" != in-tree ]; then
  ( SSE_B200_LIB=$PWD/$best timeout 100 python -m pytest tests -m gpu -x -q -k "physical or diffusion or burgers or golden or rhs_matches or packing" ) > $O/gputests_subset.log 2>&1
  tail -n 2 $O/gputests_subset.log
fi
