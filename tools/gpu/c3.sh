#!/bin/bash
O=gpurun_out/s4g; mkdir -p $O
( timeout 900 python -m pytest tests -m gpu -x -q -k "physical or diffusion or burgers or golden or rhs_matches" ) > $O/gputests_subset.log 2>&1
tail -n 5 $O/gputests_subset.log
python tools/profile_2d.py 256 advdiff advection > $O/profile2d.log 2>&1; cat $O/profile2d.log
