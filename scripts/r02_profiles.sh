#!/bin/bash
# round-2 profile capture (B200_PROFILING.md recipe): launch list of the bench command, one --set full capture of the four step kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r02_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_raster|k_step_dyn|k_step_solve|k_collide" -s 192 -c 4 -o gpurun_out/r02_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r02_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
tail -2 gpurun_out/r02_smoke.log; ls -la gpurun_out/r02_full.ncu-rep; tail -c 600 gpurun_out/r02_bench_reference.json
