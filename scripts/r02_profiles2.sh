#!/bin/bash
# round-2 final profile capture (B200_PROFILING.md recipe) after the solver split / look-ahead narrowphase: launch list of the bench
# command (ncu serialises the kernels: per-launch times are cold and un-overlapped), one --set full capture of the step kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r02b_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_raster|k_step_dyn|k_step_solve|k_collide" -s 320 -c 7 -o gpurun_out/r02b_full python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r02b_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02b_smoke.log 2>&1
python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err
python scripts/bench_configs.py > gpurun_out/r02b_configs.log 2>&1
python scripts/bench_c4.py >> gpurun_out/r02b_configs.log 2>&1
tail -2 gpurun_out/r02b_smoke.log; ls -la gpurun_out/r02b_full.ncu-rep; tail -c 1500 gpurun_out/r02b_bench_n1.json; cat gpurun_out/r02b_configs.log
