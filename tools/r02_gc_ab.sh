#!/bin/bash
out=gpurun_out/gc_ab
mkdir -p $out
for v in "" "" "" "" "" "" "" ""; do
  python bench.py --steps 50 --warmup 10 --no-extras --no-cpu-baseline $v 2>/dev/null | tail -1 > $out/b.json
  python - "$v" <<PY
import json,sys
d=json.load(open("$out/b.json"))
print(sys.argv[1] or "freeze", d["value"], d["ms_per_step"], d["resident_steps"], d["e2e"]["step_ms_min_median_max"])
PY
done
