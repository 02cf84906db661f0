#!/bin/bash
for fr in "" ""; do
  MTL_BENCH_FREE_RUN=$fr python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 10 --no-extras 2>/dev/null | tail -1 > gpurun_out/b2.json
  python - "$fr" <<PY
import json,sys
d=json.load(open("gpurun_out/b2.json"))
print("free_run=%s" % sys.argv[1], d["value"], d["ms_per_step"], d["resident_steps"], d["e2e"]["step_ms_min_median_max"])
PY
done
