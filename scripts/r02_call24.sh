#!/bin/bash
for cfg in "1 0" "1 1" "0 0"; do
  set -- $cfg
  echo "== heavy $1 side-priority $2"
  RR_HEAVY_SOLVER=$1 RR_SIDE_PRIORITY=$2 python scripts/time_step.py
  RR_HEAVY_SOLVER=$1 RR_SIDE_PRIORITY=$2 python scripts/time_bench_step.py
done
echo "== sequential"; RR_HEAVY_SOLVER=0 RR_EAGER_COLLIDE=0 python scripts/time_step.py; RR_HEAVY_SOLVER=0 RR_EAGER_COLLIDE=0 python scripts/time_bench_step.py
