#!/bin/bash
out=gpurun_out/trace2
mkdir -p $out
for c in fc1_fwd fc2_bwd s1_fc1_fwd proj_fwd; do
  MTL_LINEAR_TRACE=$out/trace_$c.txt python tools/bench_linear.py --case $c --iters 1 > /dev/null 2>&1
done
python tools/trace_summary.py $out/trace_*.txt > $out/summary.txt 2>&1
cat $out/summary.txt
