#!/bin/bash
for lib in default hsm hg37 hg296; do
  if [ $lib = default ]; then unset RR_LIB; else export RR_LIB=real_robots_b200/variants/lib_$lib.so; fi
  ACT_SCALE=0 python scripts/time_step.py
  python scripts/time_bench_step.py
done
unset RR_LIB
RR_HEAVY_SOLVER=0 ACT_SCALE=0 python scripts/time_step.py
RR_HEAVY_SOLVER=0 python scripts/time_bench_step.py
