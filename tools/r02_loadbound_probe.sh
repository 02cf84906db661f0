#!/bin/bash
# Load-bound multi-stream launches (fc1 backward, fc2 forward at stage 0): time vs the mbarrier wait hint, and a timeline
out=gpurun_out/loadbound
mkdir -p $out
for c in fc1_bwd fc2_fwd s1_fc2_bwd fc1_fwd fc2_bwd; do
  for h in 1000 200 20; do
    echo "hint=$h $(MTL_WAIT_HINT_NS=$h python tools/bench_linear.py --case $c --iters 20 2>&1 | tail -1)"
  done
done > $out/hint_sweep.txt 2>&1
for c in fc1_bwd fc2_fwd; do
  MTL_LINEAR_TRACE=$out/trace_$c.txt python tools/bench_linear.py --case $c --iters 1 > /dev/null 2>&1
done
python tools/trace_summary.py $out/trace_*.txt > $out/summary.txt 2>&1
cat $out/hint_sweep.txt $out/summary.txt
