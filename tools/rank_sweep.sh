#!/bin/bash
# BASELINE.json configs[4] / SURVEY.md §8d C5: Swin-T 448, 4 tasks, batch 32, shared rank r in {4..128}, for task rank 4
# (the shipped YAMLs) and task rank = r (all-equal variant; the packed rank space grows to 320 columns at r = 64 — the
# widest the linear kernel plans — so the all-equal sweep stops there). One JSON line per point.
#   gpurun --timeout 900 -- 'bash tools/rank_sweep.sh gpurun_out/rank_sweep.jsonl'
# then copy the file to profiles/rNN_rank_sweep.jsonl.
set -u
out=${1:-gpurun_out/rank_sweep.jsonl}
mkdir -p "$(dirname "$out")"
: > "$out"
for r in 4 8 16 32 64 128; do
  for rt in 4 $r; do
    [ "$rt" = "$r" ] && [ "$r" = 4 ] && continue      # same point as rt = 4
    [ "$rt" = 128 ] && continue                        # 5 x 128 = 640 columns: beyond the kernel's 320
    timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --r-shared "$r" --r-task "$rt" 2>/dev/null \
      | tail -1 >> "$out" || echo "{\"r_shared\": $r, \"r_task\": $rt, \"failed\": true}" >> "$out"
  done
done
python - "$out" <<'PY'
import json, sys
for line in open(sys.argv[1]):
    d = json.loads(line)
    if "value" in d:
        print(d["config"]["workload"].split(" batch")[0], round(d["value"], 1), "img/s;  linear roofline frac",
              round(d["roofline"]["frac"], 3))
    else:
        print(d)
PY
