#!/bin/bash
out=gpurun_out/sweep2
mkdir -p $out
run() {
  env "$@" python bench.py --steps 50 --warmup 10 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 > $out/b.json
  python - "$*" <<PY
import json,sys
d=json.load(open("$out/b.json"))
print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["step_ms_min_median_max"])
PY
}
run MTL_NOP=1
run MTL_SUM_IN_PLACE_MIN_K=64
run MTL_PRE_PROJECT_MIN=192
run MTL_PRE_PROJECT_MIN=96
run MTL_NOP=1
