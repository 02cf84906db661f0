#!/bin/bash
out=gpurun_out/trace3
mkdir -p $out
for c in s0_fc1_fwd s0_fc2_bwd qkv_fwd s0_fc2_fwd; do
  MTL_LINEAR_TRACE=$out/trace_$c.txt python tools/bench_linear.py --case $c --iters 1 > /dev/null 2>&1
  python tools/bench_linear.py --case $c --iters 20 2>&1 | tail -1
done
python tools/trace_summary.py $out/trace_*.txt > $out/summary.txt 2>&1
cat $out/summary.txt
