#!/bin/bash
# Round evidence on a B200 (run through gpurun): GPU test-suite, bench lines of every arm, per-op profile, ncu launch
# list of ONE step (time + DRAM bytes per launch) and ncu --set full captures of the hot kernels.
# Output: gpurun_out/evidence/; `python tools/evidence_collect.py r02` then copies the summaries into profiles/.
set -u
out=gpurun_out/evidence; mkdir -p $out
python -m pytest tests -m gpu -q --timeout 900 > $out/gpu_tests.log 2>&1; tail -3 $out/gpu_tests.log
python bench.py --profile-ops $out/ops_profile.json > $out/bench.log 2> $out/bench.err; tail -1 $out/bench.log > $out/bench.json; cut -c1-400 $out/bench.json
python __graft_entry__.py smoke > $out/smoke.log 2>&1; tail -1 $out/smoke.log
python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>$out/bench_reference.err; cut -c1-300 $out/bench_reference.json
python bench.py --impl reference-gpu --steps 20 --warmup 5 > $out/bench_reference_gpu.json 2>$out/bench_reference_gpu.err; cut -c1-300 $out/bench_reference_gpu.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --csv --log-file $out/launches.csv python bench.py --ncu-range --warmup 3 --no-extras > $out/launches.log 2>&1
for c in fc1_fwd fc2_bwd proj_fwd s2_fc1_fwd s2_fc2_fwd; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:mtl_linear_kernel --launch-skip 3 -c 1 \
    -f -o $out/ncu_linear_$c python tools/bench_linear.py --case $c --iters 1 > $out/ncu_linear_$c.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rank_project_kernel --launch-skip 3 -c 1 \
  -f -o $out/ncu_rank_project python tools/bench_linear.py --case s2_fc2_fwd --iters 1 > $out/ncu_rank_project.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:xty_umma_kernel --launch-skip 3 -c 1 \
  -f -o $out/ncu_xty_fc1_last python tools/bench_xty.py --case fc1_last --iters 1 > $out/ncu_xty.log 2>&1
for k in win_attn_fwd_kernel win_attn_bwd_kernel ln_bwd_fast_kernel opt_adamw_kernel; do
  timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$k -c 1 \
    -f -o $out/ncu_$k python bench.py --ncu-range --warmup 3 --no-extras > $out/ncu_$k.log 2>&1
done
python tools/bench_attn.py > $out/attn_ab.txt 2>&1; MTL_ATTN_UMMA=1 python tools/bench_attn.py >> $out/attn_ab.txt 2>&1
tools/r02_linear_ab.sh > /dev/null 2>&1; cp gpurun_out/r02_linear_ab.txt $out/linear_ab.txt
ls -la $out | head -60
