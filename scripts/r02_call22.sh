#!/bin/bash
for lib in default nofence default nofence; do
  if [ $lib = default ]; then unset RR_LIB; else export RR_LIB=real_robots_b200/variants/lib_$lib.so; fi
  python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 2 | tail -1 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$lib', round(d['value']), round(d['ms_per_step'], 3), {k: round(v, 3) for k, v in d['roofline']['kernels_ms_per_step'].items()}, d['roofline']['alone']['ms_per_launch'])"
  ACT_SCALE=0 python scripts/time_step.py
done
unset RR_LIB
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/r02_sanitize_racecheck.log 2>&1; echo "racecheck rc $?"; tail -3 gpurun_out/r02_sanitize_racecheck.log
