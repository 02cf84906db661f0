#!/bin/bash
# A/B of the separate rank-projection launch (LinearSpec.pre_project) on the stage-2/3 layers without task adapters.
out=gpurun_out/r02_linear_ab.txt
: > $out
for c in s2_qkv_fwd s2_fc1_fwd s2_fc2_fwd s2_qkv_bwd s2_fc2_bwd; do
  for pre in 256 1000000; do
    echo -n "pre_min=$pre " >> $out
    MTL_PRE_PROJECT_MIN=$pre python tools/bench_linear.py --case $c --iters 30 >> $out 2>&1
  done
done
cat $out
