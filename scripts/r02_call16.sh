#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02_call16_pytest.log 2>&1; tail -5 gpurun_out/r02_call16_pytest.log
{
for cfg in "0 0" "1 1"; do
  set -- $cfg
  echo "== bench heavy $1 eager $2"
  RR_HEAVY_SOLVER=$1 RR_EAGER_COLLIDE=$2 python bench.py --steps 60 --warmup 5 --no-cpu-baseline | tail -1
done
} > gpurun_out/r02_call16.log 2>&1
grep -v "^{" gpurun_out/r02_call16.log; grep "^{" gpurun_out/r02_call16.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['kernels_ms_per_step'], d['e2e']['value'])"
python scripts/time_step.py; ACT_SCALE=0 python scripts/time_step.py
