#!/bin/bash
# Quick GPU check used while iterating on kernels: op-level parity, per-shape linear timings, one short bench run.
out=gpurun_out/quick; mkdir -p $out
(timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} 2>&1 | tail -25) > $out/pytest.log; tail -4 $out/pytest.log
for c in ${CASES:-fc1_fwd fc2_fwd proj_fwd qkv_fwd fc2_bwd fc1_bwd s2_fc1_fwd s2_fc2_bwd s3_fc2_bwd}; do
  timeout 120 python tools/bench_linear.py --case $c --iters 20 2>&1 | tail -1 | tee -a $out/linear_times.txt
done
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline --profile-ops $out/ops_profile.json 2>&1 | tail -1 | tee $out/bench.json
