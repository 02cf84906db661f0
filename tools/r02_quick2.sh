#!/bin/bash
out=gpurun_out/quick2
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/pytest.log 2>&1; echo "pytest rc=$?" ; tail -5 $out/pytest.log
for v in 128 128; do
  MTL_SUM_IN_PLACE_MIN_K=$v python bench.py --steps 50 --warmup 10 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > $out/bench_$v.json
  python - <<PY
import json
d=json.load(open("$out/bench_$v.json"))
print("SUM_IN_PLACE_MIN_K=$v", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["step_ms_min_median_max"], d.get("host_issue_ms_per_step"), d.get("gpu_launches"))
PY
done
python tools/stager_cost.py 2>&1 | tail -2

