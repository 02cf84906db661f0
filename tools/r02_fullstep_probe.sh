for args in "--amp bf16" "--amp bf16 --torch-optimizer" "--amp fp16" "--amp bf16 --steps 40"; do
  python bench.py --scope full --steps 20 --warmup 5 --no-extras --no-cpu-baseline $args 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$args', round(l['ms_per_step'],2), 'e2e', round(l['e2e']['ms_per_step'],2), 'host', l['host_issue_ms_per_step'])"
done
