#!/bin/bash
mkdir -p gpurun_out
for v in "$@"; do
  RR_LIB=real_robots_b200/variants/lib_$v.so ncu --set full --clock-control none --import-source on -k regex:"k_raster" -s 3 -c 1 -o gpurun_out/r02_raster_$v python scripts/time_render.py > gpurun_out/r02_raster_ncu_$v.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -5
