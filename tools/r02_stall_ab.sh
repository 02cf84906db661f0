#!/bin/bash
out=gpurun_out/stall_ab
mkdir -p $out
for v in "" "" "" "" "" "" "" ""; do
  MTL_BENCH_NO_CLOCKS=$v python bench.py --steps 50 --warmup 10 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > $out/b.json
  python - "$v" <<PY
import json,sys
d=json.load(open("$out/b.json"))
print("no_clocks="+sys.argv[1], d["value"], d["ms_per_step"], d["resident_steps"], d["e2e"]["step_ms_min_median_max"])
PY
done
