#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -2
python scripts/c1_latency.py 2>&1 | tail -1
for i in 1 2; do python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']; print(round(d['value']), round(d['ms_per_step'], 3), round(r['ms_per_launch'], 3), round(r['alone']['ms_per_launch'], 3), d['e2e']['value'])"; done
python scripts/bench_configs.py 2>&1 | grep "^C[15]"
