#!/bin/bash
out=gpurun_out/r02_trace
mkdir -p $out
for c in s2_qkv_fwd s2_fc1_fwd s2_fc2_bwd; do
  for pre in 256 1000000; do
    MTL_PRE_PROJECT_MIN=$pre MTL_LINEAR_TRACE=$out/trace_${c}_$pre.txt python tools/bench_linear.py --case $c --iters 1 > /dev/null 2>&1
  done
done
python tools/trace_summary.py $out/trace_*.txt > $out/summary.txt 2>&1
cat $out/summary.txt
