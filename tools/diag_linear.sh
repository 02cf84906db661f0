#!/bin/bash
# Diagnostics of the fused linear kernel on a B200 (run through gpurun): role timelines (MTL_LINEAR_TRACE), planner
# sweeps (MTL_LINEAR_BN / MTL_LINEAR_SPLITS) and one ncu --set full capture per case. Output under gpurun_out/diag/.
set -u
out=gpurun_out/diag
mkdir -p $out
cases="${CASES:-s2_fc1_fwd s2_fc2_bwd fc1_fwd}"
for c in $cases; do
  python tools/bench_linear.py --case $c --iters 20 > $out/time_$c.json 2>&1
  MTL_LINEAR_TRACE=$out/trace_$c.txt python tools/bench_linear.py --case $c --iters 1 > /dev/null 2>&1
done
for c in s2_fc1_fwd s2_fc2_bwd; do
  for bn in 64 128 192; do
    for sp in 1 2 3 4; do
      echo -n "$c bn=$bn splits=$sp " >> $out/sweep.txt
      MTL_LINEAR_BN=$bn MTL_LINEAR_SPLITS=$sp python tools/bench_linear.py --case $c --iters 20 >> $out/sweep.txt 2>&1
    done
  done
done
if [ "${NCU:-1}" = "1" ]; then
  for c in ${NCU_CASES:-s2_fc1_fwd}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:mtl_linear_kernel --launch-skip 3 -c 1 \
      -f -o $out/ncu_$c python tools/bench_linear.py --case $c --iters 1 > $out/ncu_$c.log 2>&1
  done
fi
