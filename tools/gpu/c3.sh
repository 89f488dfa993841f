#!/bin/bash
O=gpurun_out/s5b; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q -k "physical or diffusion or burgers or golden or rhs_matches" ) > $O/gputests_subset.log 2>&1
tail -n 5 $O/gputests_subset.log
for so in 1 0; do echo "SSE_STAGE_OPS=$so"; SSE_STAGE_OPS=$so python tools/profile_2d.py 256 advdiff; done > $O/profile2d.log 2>&1; cat $O/profile2d.log
