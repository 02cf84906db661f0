for pre in 256 1000000; do
  MTL_PRE_PROJECT_MIN=$pre python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras --r-shared 128 --r-task 4 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pre_min=$pre', round(l['ms_per_step'],2), 'e2e', round(l['e2e']['ms_per_step'],2), 'host', l['host_issue_ms_per_step'], 'sum', round(sum(l['breakdown_ms_per_step'].values()),2), l['breakdown_ms_per_step'])"
done
