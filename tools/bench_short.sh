#!/bin/bash
# one-line summary of a short headline run: bash tools/bench_short.sh [bench.py args]
python bench.py --steps 12 --warmup 3 --no-secondary "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:round(d[k],3) if isinstance(d[k],float) else d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'launch_us', round(d['roofline']['avg_launch_us'],1), 'iters', round(d['config']['mean_gn_iters'],2), 'single', d.get('single_scan_latency'), 'err', d['config']['median_pos_err_vs_truth_m'])"
