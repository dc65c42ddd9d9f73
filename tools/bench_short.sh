python bench.py --steps 20 --warmup 3 --no-secondary 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'launch_us', round(d['roofline']['avg_launch_us'],1))"
