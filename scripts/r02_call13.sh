#!/bin/bash
# cost-class ordering of the solver + look-ahead narrowphase: A/B on one B200
mkdir -p gpurun_out
{
for lib in default real_robots_b200/variants/lib_noorder.so real_robots_b200/variants/lib_order2.so; do
  for ea in 0 1; do
    echo "== lib $lib eager $ea"
    if [ $lib = default ]; then export -n RR_LIB; unset RR_LIB; else export RR_LIB=$lib; fi
    RR_EAGER_COLLIDE=$ea python scripts/time_step.py
    RR_EAGER_COLLIDE=$ea python scripts/time_bench_step.py
  done
done
unset RR_LIB
for ea in 0 1; do
  echo "== bench eager $ea"
  RR_EAGER_COLLIDE=$ea python bench.py --steps 60 --warmup 5 --no-cpu-baseline | tail -1
done
} > gpurun_out/r02_call13.log 2>&1
python -m pytest tests -m gpu -q -x > gpurun_out/r02_call13_pytest.log 2>&1; tail -3 gpurun_out/r02_call13_pytest.log
grep -v "^{" gpurun_out/r02_call13.log; grep "^{" gpurun_out/r02_call13.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['e2e']['value'])"
