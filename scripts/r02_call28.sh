#!/bin/bash
for lib in rescut resold rescut resold; do
  export RR_LIB=real_robots_b200/variants/lib_$lib.so
  python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 2 2>/dev/null | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); r = d['roofline']; print('$lib', round(d['value']), round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in r['kernels_ms_per_step'].items()})"
  ACT_SCALE=0 python scripts/time_step.py 2>&1 | tail -1
done
unset RR_LIB
python -m pytest tests -m gpu -q 2>&1 | tail -2
python scripts/c1_latency.py 2>&1 | tail -1
