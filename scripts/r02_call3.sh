#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_step|k_collide" -s 300 -c 9 --csv --log-file gpurun_out/r02_call10_launches.csv python scripts/time_step.py > gpurun_out/r02_call10_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_step_solve|k_step_dyn" -s 200 -c 2 -o gpurun_out/r02_call10_phys python scripts/time_step.py > gpurun_out/r02_call10_b.log 2>&1
cat gpurun_out/r02_call10_launches.csv | tail -12
