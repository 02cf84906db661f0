#!/bin/bash
out=gpurun_out/stall_trace
mkdir -p $out
for v in 1 2 3 4 5 6; do
  python bench.py --steps 50 --warmup 10 --no-extras --no-cpu-baseline 2> $out/err_$v.txt | tail -1 > $out/b.json
  python - "$v" <<PY
import json,sys
d=json.load(open("$out/b.json"))
print("run "+sys.argv[1], d["value"], d["ms_per_step"], d["resident_steps"], d["e2e"]["value"], d["e2e"]["step_ms_min_median_max"])
PY

done

