#!/bin/bash
out=gpurun_out/quick3
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/pytest.log 2>&1; echo "pytest rc=$?" ; tail -5 $out/pytest.log
for v in 1 2; do
  python bench.py --steps 50 --warmup 10 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > $out/bench_$v.json
  python - <<PY
import json
d=json.load(open("$out/bench_$v.json"))
print("run $v", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["step_ms_min_median_max"], d.get("host_issue_ms_per_step"), d.get("gpu_launches"))
PY
done
python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline --profile-ops $out/ops_profile.json > /dev/null 2>&1
python tools/linear_families.py $out/ops_profile.json > $out/families.txt 2>&1
python - <<PY
import json
d=json.load(open("$out/ops_profile.json"))
for k,v in d["breakdown"].items(): print(k, v["calls_per_step"], round(v["ms_per_step"],3))
PY
