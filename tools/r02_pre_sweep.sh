for pre in 96 192 256; do
  MTL_PRE_PROJECT_MIN=$pre python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pre_min=$pre', round(l['ms_per_step'],2), 'e2e', round(l['e2e']['ms_per_step'],2), 'host', l['host_issue_ms_per_step'], l['breakdown_ms_per_step'])"
done
