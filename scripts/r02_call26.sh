#!/bin/bash
for lib in default tail0 tail592 tail2368 default tail0; do
  if [ $lib = default ]; then unset RR_LIB; else export RR_LIB=real_robots_b200/variants/lib_$lib.so; fi
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 2 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$lib', round(d['value']), round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['roofline']['kernels_ms_per_step'].items()})"
done
unset RR_LIB
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "right_after or late_env or mixed_call" 2>&1 | tail -1
