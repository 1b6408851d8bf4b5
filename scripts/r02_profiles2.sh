#!/bin/bash
# round-2 final profile capture (tree of the last commit) (B200_PROFILING.md recipe) after the solver split / look-ahead narrowphase: launch list of the bench
# command (ncu serialises the kernels: per-launch times are cold and un-overlapped), one --set full capture of the step kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r02c_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_raster|k_step_dyn|k_step_solve|k_collide" -s 194 -c 10 -o gpurun_out/r02c_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r02c_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1
python bench.py > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err
tail -2 gpurun_out/r02c_smoke.log; ls -la gpurun_out/r02c_full.ncu-rep; tail -c 1500 gpurun_out/r02c_bench_n1.json
