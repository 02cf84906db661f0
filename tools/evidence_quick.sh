#!/bin/bash
# Reduced evidence refresh (tests, bench lines of the three arms, per-op profile, ncu launch list); the ncu --set full
# captures of tools/evidence.sh are left as they are.
set -u
out=gpurun_out/evidence; mkdir -p $out
python -m pytest tests -m gpu -q --timeout 900 > $out/gpu_tests.log 2>&1; tail -3 $out/gpu_tests.log
python bench.py --profile-ops $out/ops_profile.json > $out/bench.log 2> $out/bench.err; tail -1 $out/bench.log > $out/bench.json; cut -c1-300 $out/bench.json
python __graft_entry__.py smoke > $out/smoke.log 2>&1; tail -1 $out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>$out/bench_reference.err; cut -c1-200 $out/bench_reference.json
python bench.py --impl reference-gpu --steps 20 --warmup 5 > $out/bench_reference_gpu.json 2>$out/bench_reference_gpu.err; cut -c1-200 $out/bench_reference_gpu.json
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $out/launches.csv python bench.py --ncu-range --warmup 3 --no-extras > $out/launches.log 2>&1
