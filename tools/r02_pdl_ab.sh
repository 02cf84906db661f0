#!/bin/bash
# Programmatic dependent launch on / off: GPU tests with it on, then the bench line both ways
out=gpurun_out/pdl
mkdir -p $out
python -m pytest tests -x -q -m gpu > $out/pytest.log 2>&1; echo "pytest rc=$?" ; tail -3 $out/pytest.log
for v in 1 0 1 0; do
  MTL_PDL=$v python bench.py --steps 50 --warmup 10 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > $out/bench_pdl$v.json
  python - <<PY
import json
d=json.load(open("$out/bench_pdl$v.json"))
print("MTL_PDL=$v", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("host_issue_ms_per_step"))
PY
done
