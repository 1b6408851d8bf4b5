#!/bin/bash
mkdir -p gpurun_out
python scripts/soak.py > gpurun_out/r02_soak.log 2>&1; tail -3 gpurun_out/r02_soak.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/r02_sanitize_memcheck.log 2>&1; echo "memcheck rc $?"; tail -4 gpurun_out/r02_sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/r02_sanitize_racecheck.log 2>&1; echo "racecheck rc $?"; tail -4 gpurun_out/r02_sanitize_racecheck.log
python scripts/bench_configs.py 2>&1 | grep "^C1"
