#!/bin/bash
for p in 0 1 2 3; do
echo "== prio $p"
RR_SIDE_PRIORITY=$p python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 2 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['e2e']['value'])"
RR_SIDE_PRIORITY=$p python scripts/time_bench_step.py
done
